"""Parity of the CUDA WIS / mod-coverage path (through the C ABI) against the oracle: reference
target sets and their order, distances, threshold, call counts and filter bits bit-exact;
mod-coverage f32 outputs bit-exact except CV2 (f64 log2, rel 1e-6)."""
import numpy as np
import pytest

import cnvetti_b200 as cb
from cnvetti_b200 import model_wis as mw
from cnvetti_b200 import synth
from util import assert_f32_equal_bits, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return cb.Context(0)


def _dev(a):
    import torch
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("T,S,k,ncontig,device,engine", [
    (300, 20, 30, 5, False, mw.ENGINE_EXACT), (257, 7, 100, 3, True, mw.ENGINE_EXACT),
    (12, 5, 8, 2, False, mw.ENGINE_EXACT), (1500, 100, 100, 22, True, mw.ENGINE_EXACT),
    (50, 33, 100, 1, False, mw.ENGINE_EXACT),
])
def test_wis_distances_parity(oracle, ctx, T, S, k, ncontig, device, engine):
    X, tids = synth.synth_panel(T + S, T, S, n_contigs=ncontig)
    odist, oidx = oracle.wis_distances(X, tids, max_ref_targets=k, nthreads=4)
    opts = mw.BuildModelWisOptions(max_ref_targets=k, engine=engine)
    if device:
        dist, idx = mw.compute_distances(_dev(X), _dev(tids), opts, ctx=ctx)
    else:
        dist, idx = mw.compute_distances(X, tids, opts, ctx=ctx)
    assert np.array_equal(to_np(idx).astype(np.uint32), oidx)
    assert_f32_equal_bits(to_np(dist).ravel(), odist.ravel(), "distances")
    # a row block only (multi-GPU sharding unit)
    lo, hi = T // 3, 2 * T // 3
    d2, i2 = mw.compute_distances(X, tids, opts, row_begin=lo, row_end=hi, ctx=ctx)
    assert np.array_equal(i2, oidx[lo:hi]) and np.array_equal(d2.view(np.uint32), odist[lo:hi].view(np.uint32))


@pytest.mark.parametrize("T,S,device", [(400, 24, False), (2000, 100, True)])
def test_wis_build_parity(oracle, ctx, T, S, device):
    X, tids = synth.synth_panel(T, T, S, n_contigs=8)
    rng = np.random.default_rng(T)
    X[rng.integers(0, T, 6)] *= 2.5     # unreliable / isolated targets
    X[5, 3] *= 1.8                      # a single-sample event
    o = dict(filter_z_score=3.0, filter_rel=0.2, min_ref_targets=10, max_ref_targets=40, max_samples_reliable=1)
    odist, oidx = oracle.wis_distances(X, tids, max_ref_targets=o["max_ref_targets"], nthreads=4)
    othr, oridx, orlen = oracle.wis_prune(odist, oidx, z=o["filter_z_score"])
    ocalls = oracle.wis_calls(X, oridx, orlen, o["min_ref_targets"], o["filter_z_score"], o["filter_rel"])
    ofilt = oracle.wis_filters(orlen, ocalls, o["min_ref_targets"], o["max_samples_reliable"])
    opts = mw.BuildModelWisOptions(engine=mw.ENGINE_EXACT, **o)
    m = mw.run(_dev(X), _dev(tids), opts, ctx=ctx) if device else mw.run(X, tids, opts, ctx=ctx)
    assert np.float32(m.threshold).view(np.uint32) == np.float32(othr).view(np.uint32)
    assert_f32_equal_bits(to_np(m.ref_dist).ravel(), odist.ravel(), "distances")
    rl = to_np(m.ref_len).astype(np.uint32)
    assert np.array_equal(rl, orlen)
    ri = to_np(m.ref_idx).astype(np.uint32)
    for t in range(T):
        assert np.array_equal(ri[t, :rl[t]], oridx[t, :orlen[t]]), t
    assert np.array_equal(to_np(m.num_calls).astype(np.uint32), ocalls)
    assert np.array_equal(to_np(m.filt), ofilt)
    assert ocalls.sum() > 0 and (ofilt & 1).any() and (ofilt & 2).any()
    # the separate entry points agree with the one-shot call
    thr = mw.prune_threshold(odist[:, 0].copy(), o["filter_z_score"], ctx=ctx)
    assert np.float32(thr).view(np.uint32) == np.float32(othr).view(np.uint32)
    ridx2, rlen2 = mw.prune_distances(odist, oidx.copy(), thr, ctx=ctx)
    assert np.array_equal(rlen2, orlen)
    calls2 = mw.count_per_probe_calls(X, np.ascontiguousarray(ridx2[10:T - 5]), np.ascontiguousarray(rlen2[10:T - 5]), opts,
                                      row_begin=10, row_end=T - 5, ctx=ctx)
    assert np.array_equal(calls2, ocalls[10:T - 5])
    filt2 = mw.model_filters(orlen, ocalls, opts, ctx=ctx)
    assert np.array_equal(filt2, ofilt)


def test_wis_single_contig_threshold_nan(oracle, ctx):
    X, _ = synth.synth_panel(1, 40, 6, n_contigs=1)
    tids = np.zeros(40, np.uint32)
    m = mw.run(X, tids, mw.BuildModelWisOptions(max_ref_targets=10, engine=mw.ENGINE_EXACT), ctx=ctx)
    odist, oidx = oracle.wis_distances(X, tids, max_ref_targets=10)
    othr, _, orlen = oracle.wis_prune(odist, oidx)
    assert np.isinf(to_np(m.ref_dist)).all() and np.isnan(m.threshold) and np.isnan(othr)
    assert np.array_equal(to_np(m.ref_len), orlen) and (orlen == 0).all()


@pytest.mark.parametrize("device", [False, True])
def test_modcov_parity(oracle, ctx, device):
    T, S = 3000, 12
    X, tids = synth.synth_panel(9, T, S, n_contigs=10)
    odist, oidx = oracle.wis_distances(X, tids, max_ref_targets=50, nthreads=4)
    _, ridx, rlen = oracle.wis_prune(odist, oidx, z=2.0)
    rng = np.random.default_rng(1)
    filt = (rng.random(T) < 0.1).astype(np.uint8)
    rlen[rng.integers(0, T, 20)] = 0
    cv = (X[:, 0] * rng.normal(1, 0.1, T)).astype(np.float32)
    cv[17] = 0.0
    o = oracle.modcov(cv, filt, ridx, rlen)
    if device:
        g = mw.mod_coverage(_dev(cv), _dev(filt), _dev(ridx), _dev(rlen), ctx=ctx)
    else:
        g = mw.mod_coverage(cv, filt, ridx, rlen, ctx=ctx)
    assert np.array_equal(to_np(g["status"]), o["status"])
    for key in ("ref_mean", "ref_sd", "cv_rel", "cvz"):
        assert_f32_equal_bits(g[key], o[key], key)
    a, b = to_np(g["cv2"]), o["cv2"]
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    np.testing.assert_allclose(a[fin], b[fin], rtol=1e-6, atol=1e-7)
    assert (o["status"] == 0).any() and (o["status"] == 3).any()


@pytest.mark.parametrize("T,S,device,k,env", [
    (3000, 12, True, 50, {}), (1037, 70, False, 50, {}), (40, 33, True, 40, {}),    # 16-byte / 4-byte (S % 4 != 0) cp.async, 64-sample chunks
    (1200, 136, True, 100, {}), (1200, 136, True, 100, {"CNVB_PANEL_STAGE": "1"}),   # three chunks, the last one partial
    (1200, 136, True, 100, {"CNVB_PANEL_STAGE": "2"}), (1037, 72, True, 50, {"CNVB_PANEL_STAGE": "2", "CNVB_PANEL_SPL": "1"}),
    (1500, 40, True, 150, {}), (900, 45, True, 150, {}),                             # 32-sample chunks (105 < k <= 199)
    (1037, 72, True, 50, {"CNVB_PANEL_SPL": "1"}),
    (1200, 45, True, 230, {}), (1037, 70, True, 50, {"CNVB_PANEL_STAGE": "0"})])     # the unstaged kernel
def test_modcov_panel_matches_per_sample_oracle(oracle, ctx, monkeypatch, T, S, device, k, env):
    """cnvb_modcov_panel: every column of the sample-major outputs equals the oracle's per-sample
    mod-coverage of that sample (bit-exact; CV2 = log2 within the north_star tolerance), in every form of the
    kernel (reference values staged in shared memory by 16-byte cp.async, 4-byte cp.async or bulk copies, or
    read from global memory)."""
    for key, val in env.items():
        monkeypatch.setenv(key, val)
    X, tids = synth.synth_panel(31 + T, T, S, n_contigs=6)
    odist, oidx = oracle.wis_distances(X, tids, max_ref_targets=k, nthreads=4)
    _, ridx, rlen = oracle.wis_prune(odist, oidx, z=2.0)
    rng = np.random.default_rng(2)
    filt = (rng.random(T) < 0.1).astype(np.uint8)
    rlen[rng.integers(0, T, 5)] = 0
    Y = (X * rng.normal(1, 0.1, X.shape)).astype(np.float32)  # the samples being called
    Y[17 % T, 3] = 0.0
    if S > 42:   # a zero, a subnormal and a negative value in columns the kernel widens with integer instructions
        Y[5, 40] = 0.0   # (samples 32 .. 63 of a chunk)
        Y[9, 41] = np.float32(1e-42)
        Y[11, 42] = -Y[11, 42]
    Y[::7] *= np.float32(1e-15)   # reference sets mixing magnitudes: the plain f64 sum is not exact there
    names = ("ref_mean", "ref_sd", "cv_rel", "cvz", "cv2", "status")
    if device:
        g = mw.mod_coverage_panel(_dev(Y), _dev(filt), _dev(ridx), _dev(rlen), want=names, ctx=ctx)
    else:
        g = mw.mod_coverage_panel(Y, filt, ridx, rlen, want=names, ctx=ctx)
    for s in range(S):
        o = oracle.modcov(np.ascontiguousarray(Y[:, s]), filt, ridx, rlen)
        assert np.array_equal(to_np(g["status"])[s], o["status"])
        for key in ("ref_mean", "ref_sd", "cv_rel", "cvz"):
            assert_f32_equal_bits(to_np(g[key])[s], o[key], "%s[%d]" % (key, s))
        a, b = to_np(g["cv2"])[s], o["cv2"]
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin)
        np.testing.assert_allclose(a[fin], b[fin], rtol=1e-6, atol=1e-7)
    # nullable outputs
    h = mw.mod_coverage_panel(_dev(Y), _dev(filt), _dev(ridx), _dev(rlen), want=("cvz",), ctx=ctx)
    assert list(h) == ["cvz"]
    assert_f32_equal_bits(to_np(h["cvz"]), to_np(g["cvz"]), "cvz only")
    # a row block (a GPU's share of the targets)
    lo, hi = T // 3, T - 5
    h = mw.mod_coverage_panel(Y, filt[lo:hi], ridx[lo:hi], rlen[lo:hi], want=("cvz", "status"), row_begin=lo,
                              row_end=hi, ctx=ctx)
    assert_f32_equal_bits(to_np(h["cvz"]), to_np(g["cvz"])[:, lo:hi], "cvz rows")
    assert np.array_equal(to_np(h["status"]), to_np(g["status"])[:, lo:hi])


@pytest.mark.parametrize("T,S,k,ncontig,device", [(5000, 100, 100, 22, True), (4321, 37, 64, 5, False),
                                                 (9000, 128, 100, 22, True),
                                                 # more than 128 samples: both operands stream through the K-block ring
                                                 (4400, 300, 50, 7, True), (4100, 130, 40, 5, False), (4200, 1000, 30, 4, True)])
def test_wis_tensor_engine_matches_reference_sets(oracle, ctx, T, S, k, ncontig, device):
    """tcgen05 contraction as candidate filter + exact re-scoring: sets, order and distances
    bit-exact against the oracle's f32-sequential distances."""
    X, tids = synth.synth_panel(T * 7 + S, T, S, n_contigs=ncontig)
    X[11] *= 4.0   # a far-away row (large norm, large rounding error bound)
    X[12] = X[13]  # duplicate rows on the same contig and (below) on different contigs: ties
    X[T - 1] = X[0]
    odist, oidx = oracle.wis_distances(X, tids, max_ref_targets=k, nthreads=16)
    opts = mw.BuildModelWisOptions(max_ref_targets=k, engine=mw.ENGINE_TENSOR)
    if device:
        dist, idx = mw.compute_distances(_dev(X), _dev(tids), opts, ctx=ctx)
    else:
        dist, idx = mw.compute_distances(X, tids, opts, ctx=ctx)
    st = mw.last_stats(ctx)
    assert np.array_equal(to_np(idx).astype(np.uint32), oidx)
    assert_f32_equal_bits(to_np(dist).ravel(), odist.ravel(), "distances")
    assert st["rows"] == T and st["tiles"] >= 1
    assert st["fallback_rows"] <= T // 100, st          # the bound, not the fallback, does the work
    assert st["rescored_pairs"] < 6 * T * k, st         # and it is tight: few pairs are re-scored
    # row block (multi-GPU shard) starting off a tile boundary
    lo, hi = 777, 777 + 1300
    d2, i2 = mw.compute_distances(X, tids, opts, row_begin=lo, row_end=hi, ctx=ctx)
    assert np.array_equal(i2, oidx[lo:hi]) and np.array_equal(d2.view(np.uint32), odist[lo:hi].view(np.uint32))


def test_wis_tensor_low_efficiency_rows_many_samples(oracle, ctx):
    """More than 128 samples: the all-ones component is carried in f32 beside the fp16 residual (ortho
    mode).  Targets with a very low or very high capture efficiency -- whose fp16 norms used to be
    dominated by that component, so that the K-proportional error bound let their candidate buffers
    overflow on dense panels -- must stay on the tensor path and match the oracle bit for bit."""
    T, S, k = 6000, 256, 60
    rng = np.random.default_rng(77)
    eff = rng.lognormal(0.0, 1.0, T)            # a wide range: 0.05 .. 20
    eff[:300] = rng.uniform(0.02, 0.06, 300)    # a dense cluster of low-efficiency targets
    X = (eff[:, None] * rng.normal(1.0, 0.08, (T, S)))
    X = (np.maximum(X, 0.0) / X.sum(axis=0, keepdims=True)).astype(np.float32)
    X[5] = X[6]
    tids = (np.arange(T) * 4 // T).astype(np.uint32)
    odist, oidx = oracle.wis_distances(X, tids, max_ref_targets=k, nthreads=16)
    opts = mw.BuildModelWisOptions(max_ref_targets=k, engine=mw.ENGINE_TENSOR)
    dist, idx = mw.compute_distances(_dev(X), _dev(tids), opts, ctx=ctx)
    st = mw.last_stats(ctx)
    assert np.array_equal(to_np(idx).astype(np.uint32), oidx)
    assert_f32_equal_bits(to_np(dist).ravel(), odist.ravel(), "distances")
    assert st["fallback_rows"] == 0, st
    assert st["rescored_pairs"] < 4 * T * k, st


@pytest.mark.parametrize("T,S,device,min_ref,max_rel", [(3000, 40, True, 10, 1), (777, 70, False, 0, 0), (40, 33, True, 3, 4)])
def test_calls_and_modcov_in_one_pass(oracle, ctx, T, S, device, min_ref, max_rel):
    """cnvb_wis_calls_modcov_panel = count_per_probe_calls + FILTER + mod-coverage of the cohort: identical to
    the separate stages (which are held against the oracle elsewhere), bit for bit."""
    X, tids = synth.synth_panel(61 + T, T, S, n_contigs=6)
    rng = np.random.default_rng(T)
    X[rng.integers(0, T, T // 20), rng.integers(0, S, T // 20)] *= 3.0   # outliers: calls
    k = min(T, 50)
    odist, oidx = oracle.wis_distances(X, tids, max_ref_targets=k, nthreads=4)
    _, ridx, rlen = oracle.wis_prune(odist, oidx, z=2.0)
    rlen[rng.integers(0, T, 7)] = 0
    rlen[rng.integers(0, T, 7)] = 2
    opts = mw.BuildModelWisOptions(filter_z_score=2.5, filter_rel=0.1, min_ref_targets=min_ref, max_samples_reliable=max_rel)
    conv = _dev if device else (lambda a: a)
    names = ("ref_mean", "ref_sd", "cv_rel", "cvz", "cv2", "status")
    calls, filt, g = mw.calls_and_mod_coverage_panel(conv(X), conv(ridx), conv(rlen), opts, want=names, ctx=ctx)
    calls0 = mw.count_per_probe_calls(conv(X), conv(ridx), conv(rlen), opts, ctx=ctx)
    filt0 = mw.model_filters(conv(rlen), calls0, opts, ctx=ctx)
    g0 = mw.mod_coverage_panel(conv(X), filt0, conv(ridx), conv(rlen), want=names, ctx=ctx)
    ocalls = oracle.wis_calls(X, ridx, rlen, min_ref, 2.5, 0.1)
    assert np.array_equal(to_np(calls0).astype(np.uint32), ocalls) and ocalls.sum() > 0
    assert np.array_equal(to_np(calls).astype(np.uint32), ocalls)
    assert np.array_equal(to_np(filt), to_np(filt0)) and (to_np(filt) != 0).any() and (to_np(filt) == 0).any()
    for key in names:
        a, b = to_np(g[key]), to_np(g0[key])
        assert np.array_equal(a.view(np.uint8 if key == "status" else np.uint32), b.view(np.uint8 if key == "status" else np.uint32)), key
    lo, hi = T // 4, T - 3
    c2, f2, h = mw.calls_and_mod_coverage_panel(X, ridx[lo:hi], rlen[lo:hi], opts, want=("cvz",), row_begin=lo, row_end=hi, ctx=ctx)
    assert np.array_equal(c2, ocalls[lo:hi]) and np.array_equal(f2, to_np(filt)[lo:hi])
    assert np.array_equal(h["cvz"].view(np.uint32), to_np(g["cvz"])[:, lo:hi].view(np.uint32))


@pytest.mark.parametrize("cluster,xpose", [(1, 0), (1, 1), (2, 1), (4, 0), (4, 1)])
def test_wis_streamed_sweep_variants(oracle, ctx, monkeypatch, cluster, xpose):
    """The forms of the streamed sweep that the defaults do not pick at test sizes -- one CTA per row block,
    clusters of four, transposed M128 x N256 tiles (the default from 10^6 targets) -- against the oracle."""
    monkeypatch.setenv("CNVB_WIS_CLUSTER", str(cluster))
    monkeypatch.setenv("CNVB_WIS_XPOSE", str(xpose))
    T, S, k = 5300, 200, 40
    X, tids = synth.synth_panel(900 + cluster * 10 + xpose, T, S, n_contigs=6)
    X[21] *= 3.0
    X[30] = X[31]
    odist, oidx = oracle.wis_distances(X, tids, max_ref_targets=k, nthreads=16)
    opts = mw.BuildModelWisOptions(max_ref_targets=k, engine=mw.ENGINE_TENSOR)
    dist, idx = mw.compute_distances(_dev(X), _dev(tids), opts, ctx=ctx)
    st = mw.last_stats(ctx)
    assert np.array_equal(to_np(idx).astype(np.uint32), oidx)
    assert_f32_equal_bits(to_np(dist).ravel(), odist.ravel(), "distances")
    assert st["fallback_rows"] == 0, st
    lo, hi = 333, 333 + 2100   # a row block off the tile boundaries
    d2, i2 = mw.compute_distances(X, tids, opts, row_begin=lo, row_end=hi, ctx=ctx)
    assert np.array_equal(i2, oidx[lo:hi]) and np.array_equal(d2.view(np.uint32), odist[lo:hi].view(np.uint32))


@pytest.mark.parametrize("S", [100, 125, 126, 128])
def test_wis_sweep_forms_by_sample_count(oracle, ctx, S):
    """Up to 125 samples the rows' operand lives in tensor memory and the column norms are three extra K entries
    (S + 3 <= 128, wis_resident.cuh); 126 - 128 samples take the streamed sweep with two K blocks -- both select
    the oracle's sets, bit for bit, including the boundary cases either side."""
    T, k = 6100, 40
    X, tids = synth.synth_panel(911 + S, T, S, n_contigs=6)
    X[40] = X[41]
    odist, oidx = oracle.wis_distances(X, tids, max_ref_targets=k, nthreads=16)
    opts = mw.BuildModelWisOptions(max_ref_targets=k, engine=mw.ENGINE_TENSOR)
    dist, idx = mw.compute_distances(_dev(X), _dev(tids), opts, ctx=ctx)
    assert np.array_equal(to_np(idx).astype(np.uint32), oidx)
    assert_f32_equal_bits(to_np(dist).ravel(), odist.ravel(), "distances")
    assert mw.last_stats(ctx)["fallback_rows"] == 0


def test_wis_tensor_single_contig_and_tiny_other_contig(oracle, ctx):
    T, S, k = 4500, 20, 50
    X, _ = synth.synth_panel(77, T, S, n_contigs=1)
    tids = np.zeros(T, np.uint32)
    tids[-30:] = 1   # only 30 other-contig targets for most rows: +inf fill
    odist, oidx = oracle.wis_distances(X, tids, max_ref_targets=k, nthreads=16)
    opts = mw.BuildModelWisOptions(max_ref_targets=k, engine=mw.ENGINE_TENSOR)
    dist, idx = mw.compute_distances(X, tids, opts, ctx=ctx)
    assert np.array_equal(idx, oidx) and np.array_equal(dist.view(np.uint32), odist.view(np.uint32))


@pytest.mark.parametrize("kind", ["flat_projection", "two_clusters", "sparse_shard"])
def test_wis_tensor_windows_adversarial(oracle, ctx, kind):
    """The projection windows prune, they never decide: panels whose all-ones projection carries no
    information (every window must grow to the full width), two far-apart efficiency clusters (windows
    end at the gap) and a sparse row shard (a CTA's rows span many column tiles)."""
    rng = np.random.default_rng(123)
    T, S, k = 4600, 64, 40
    if kind == "flat_projection":
        X = rng.normal(0.0, 1.0, (T, S)).astype(np.float32)
        X -= X.mean(axis=1, keepdims=True)          # all row sums (projections) ~ 0
        X += 5.0
    elif kind == "two_clusters":
        eff = np.where(rng.random(T) < 0.5, 1.0, 7.0)
        X = (eff[:, None] * rng.normal(1.0, 0.05, (T, S))).astype(np.float32)
    else:
        X, _ = synth.synth_panel(99, T, S, n_contigs=8)
    tids = rng.integers(0, 8, T).astype(np.uint32)
    odist, oidx = oracle.wis_distances(X, tids, max_ref_targets=k, nthreads=16)
    opts = mw.BuildModelWisOptions(max_ref_targets=k, engine=mw.ENGINE_TENSOR)
    if kind == "sparse_shard":
        lo, hi = 100, 100 + 460   # 10 % of the rows: their sorted positions are spread over all tiles
        dist, idx = mw.compute_distances(X, tids, opts, row_begin=lo, row_end=hi, ctx=ctx)
        odist, oidx = odist[lo:hi], oidx[lo:hi]
    else:
        dist, idx = mw.compute_distances(X, tids, opts, ctx=ctx)
    assert np.array_equal(to_np(idx).astype(np.uint32), oidx)
    assert_f32_equal_bits(to_np(dist).ravel(), odist.ravel(), "distances")


def test_wis_full_size_engines_agree(oracle, ctx):
    """BASELINE config 3 at full size (T = 200 000, S = 100, k = 100): the tensor engine's reference
    sets for a block of rows equal the exact engine's (every pair in emulated f32-sequential
    arithmetic) and, for a few rows, the oracle's."""
    T, S, k = 200_000, 100, 100
    X, tids = synth.synth_panel(3, T, S)
    dX, dt = _dev(X), _dev(tids)
    full_d, full_i = mw.compute_distances(dX, dt, mw.BuildModelWisOptions(max_ref_targets=k, engine=mw.ENGINE_TENSOR), ctx=ctx)
    st = mw.last_stats(ctx)
    assert st["fallback_rows"] == 0 and st["tiles"] * 256 * 128 < 0.3 * T * T, st
    lo, hi = 123_456, 123_456 + 192
    ex_d, ex_i = mw.compute_distances(dX, dt, mw.BuildModelWisOptions(max_ref_targets=k, engine=mw.ENGINE_EXACT),
                                      row_begin=lo, row_end=hi, ctx=ctx)
    assert np.array_equal(to_np(full_i)[lo:hi], to_np(ex_i))
    assert_f32_equal_bits(to_np(full_d)[lo:hi].ravel(), to_np(ex_d).ravel(), "distances")
    rows = [0, 77_777, T - 1]
    for r in rows:
        od, oi = oracle.wis_distances(X, tids, max_ref_targets=k, row_begin=r, row_end=r + 1, nthreads=16)
        assert np.array_equal(to_np(full_i)[r].astype(np.uint32), oi[0])
        assert_f32_equal_bits(to_np(full_d)[r], od[0], "distances")
    # size-independent properties: ascending distances, no same-contig reference, no duplicates
    d = to_np(full_d)
    i = to_np(full_i).astype(np.int64)
    assert (np.diff(d, axis=1) >= 0).all()
    assert (tids[i] != tids[:, None]).all()
    s = np.sort(i[::97], axis=1)
    assert (np.diff(s, axis=1) > 0).all()
