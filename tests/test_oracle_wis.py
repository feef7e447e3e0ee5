"""CPU checks of the oracle's WIS / mod-coverage restatements against independent numpy models
(parity unpinned by the reference's own tests; the cited lines are lib-model-wis/src/lib.rs:128-271
and lib-mod-cov/src/lib.rs:76-164)."""
import math

import numpy as np

from cnvetti_b200 import synth


def _seq_dist(a, b):
    y = np.float32(0)
    for s in range(a.size):
        x = np.float32(a[s] - b[s])
        y = np.float32(y + np.float32(x * x))
    return np.float32(np.sqrt(y))


def test_wis_distances_small(oracle):
    X, tids = synth.synth_panel(3, 60, 7, n_contigs=4)
    dist, idx = oracle.wis_distances(X, tids, max_ref_targets=10)
    assert dist.shape == (60, 10)
    for i in (0, 17, 59):
        ents = []
        for j in range(60):
            d = np.float32(np.inf) if tids[i] == tids[j] else _seq_dist(X[i], X[j])
            ents.append((d, j))
        ents.sort()
        assert [e[1] for e in ents[:10]] == idx[i].tolist()
        assert np.array_equal(np.array([e[0] for e in ents[:10]], np.float32), dist[i])


def test_wis_inf_fill_when_few_other_contig_targets(oracle):
    X, _ = synth.synth_panel(4, 12, 5, n_contigs=2)
    tids = np.array([0] * 9 + [1] * 3, np.uint32)
    dist, idx = oracle.wis_distances(X, tids, max_ref_targets=8)
    # rows on contig 0 see only 3 other-contig targets, then +inf entries in index order
    assert np.isinf(dist[0, 3:]).all() and idx[0, 3:].tolist() == [0, 1, 2, 3, 4]
    assert np.isfinite(dist[10, :8]).all()
    # multi-threaded rows agree with single-threaded
    d2, i2 = oracle.wis_distances(X, tids, max_ref_targets=8, nthreads=4)
    assert np.array_equal(dist, d2) and np.array_equal(idx, i2)


def test_wis_prune_calls_filters(oracle):
    X, tids = synth.synth_panel(5, 300, 20, n_contigs=5)
    X[7] *= 3.0  # an outlier target: unreliable / few refs
    dist, idx = oracle.wis_distances(X, tids, max_ref_targets=30)
    thr, ridx, rlen = oracle.wis_prune(dist, idx, z=1.0)
    e = dist[:, 0].astype(np.float64)
    mean = math.fsum(e) / e.size
    sd = math.sqrt(sum((x - mean) ** 2 for x in e) / (e.size - 1))
    assert thr == np.float32(mean + 1.0 * sd)
    for t in (0, 7, 100):
        keep = [int(j) for d, j in zip(dist[t], idx[t]) if d <= thr]
        assert keep == ridx[t, :rlen[t]].tolist()
    calls = oracle.wis_calls(X, ridx, rlen, min_ref_targets=5, z=2.0, rel=0.05)
    t = int(np.argmax(rlen >= 5))
    n = 0
    for s in range(X.shape[1]):
        ref = X[ridx[t, :rlen[t]], s].astype(np.float64)
        m = math.fsum(ref) / ref.size
        v = 0.0
        for r in ref:
            v = v + (r - m) * (r - m)
        sdv = math.sqrt(v / (ref.size - 1))
        x = float(X[t, s])
        if abs(x - m) / sdv > 2.0 and abs(x / m - 1.0) > 0.05:
            n += 1
    assert calls[t] == n
    filt = oracle.wis_filters(rlen, calls, min_ref_targets=5, max_samples_reliable=2)
    assert np.array_equal(filt & 1, (rlen < 5).astype(np.uint8))
    assert np.array_equal((filt >> 1) & 1, (calls > 2).astype(np.uint8))


def test_modcov(oracle):
    X, tids = synth.synth_panel(6, 200, 8, n_contigs=4)
    dist, idx = oracle.wis_distances(X, tids, max_ref_targets=20)
    thr, ridx, rlen = oracle.wis_prune(dist, idx, z=5.64)
    filt = np.zeros(200, np.uint8)
    filt[3] = 1
    rlen[5] = 0
    cv = X[:, 0].copy()
    cv[9] = 0.0
    r = oracle.modcov(cv, filt, ridx, rlen)
    assert r["status"][3] == 0 and r["status"][5] == 0 and np.isnan(r["cv_rel"][3])
    t = 11
    ref = cv[ridx[t, :rlen[t]]].astype(np.float64)
    m = math.fsum(ref) / ref.size
    assert r["ref_mean"][t] == np.float32(m) and r["cv_rel"][t] == np.float32(float(cv[t]) / m)
    assert r["status"][9] == 3 and r["cv2"][9].view(np.uint32) == 0x7F800001  # rel == 0 -> missing CV2
