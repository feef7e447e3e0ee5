"""Committed vectors (tests/golden/path_vectors.npz, script beside it): the CPU restatement must
reproduce them (`-m "not gpu"`), and so must the CUDA path through the C ABI (`-m gpu`)."""
import os

import numpy as np
import pytest

from util import assert_f32_equal_bits, to_np

V = np.load(os.path.join(os.path.dirname(__file__), "golden", "path_vectors.npz"))


def _reads():
    return {k: V["reads_" + k] for k in ("pos", "isize", "flag", "mapq", "cigar_off", "cigar")}


def test_oracle_reproduces_committed_vectors(oracle):
    d = _reads()
    L, wl = int(V["L"]), int(V["wl"])
    rb = oracle.ReadBatch(d["pos"], d["isize"], d["flag"], d["mapq"], d["cigar_off"], d["cigar"])
    oo = oracle.cov_opts(window_length=wl, min_mapq=20)
    counts, proc, skip = oracle.frag_gw(rb, oo, L)
    assert np.array_equal(counts, V["gw_counts"]) and [proc, skip] == V["gw_tally"].tolist()
    tc, tp, tsk = oracle.frag_targets(rb, oracle.cov_opts(window_length=0, min_mapq=20), V["tgt_start"], V["tgt_end"])
    assert np.array_equal(tc, V["tgt_counts"]) and [tp, tsk] == V["tgt_tally"].tolist()
    mean, sd = oracle.coverage(rb, oo, L)
    assert_f32_equal_bits(mean, V["cov_mean"]) and assert_f32_equal_bits(sd, V["cov_sd"])
    gc, gap = oracle.refstats(V["seq"], wl)
    assert_f32_equal_bits(gc, V["gc"])
    assert np.array_equal(gap, V["gap"])
    s, cv, _ = oracle.norm_total(V["lcv"])
    assert s == float(V["n1_sum"])
    assert_f32_equal_bits(cv, V["n1_cv"])
    cv, cv2, _, flags, med = oracle.norm_gc_median(V["n2_lcv"], V["n2_gc"])
    assert np.array_equal(flags, V["n2_flags"])
    assert_f32_equal_bits(med, V["n2_med"])
    assert_f32_equal_bits(cv[(flags & 1) != 0], V["n2_cv"][(flags & 1) != 0])
    dist, idx = oracle.wis_distances(V["wis_X"], V["wis_tid"], max_ref_targets=30)
    assert np.array_equal(idx, V["wis_idx"])
    assert_f32_equal_bits(dist, V["wis_dist"])
    thr, pidx, plen = oracle.wis_prune(dist, idx, 5.64)
    assert np.float32(thr).tobytes() == np.float32(V["wis_thr"]).tobytes() and np.array_equal(plen, V["wis_ref_len"])
    assert np.array_equal(oracle.wis_calls(V["wis_X"], pidx, plen, 10, 5.64, 0.35), V["wis_calls"])
    assert np.array_equal(oracle.hmm_viterbi(V["hmm_cv"], oracle.hmm_tables(0.08)), V["hmm_state"])


@pytest.mark.gpu
def test_cuda_path_reproduces_committed_vectors():
    import cnvetti_b200 as cb
    from cnvetti_b200 import discover, model_wis as mw, piles, pipeline
    from util import soa_batch
    ctx = cb.Context(0)
    d = _reads()
    L, wl = int(V["L"]), int(V["wl"])
    batch = soa_batch(d, device="cuda")
    # A2 + A5 + A6 through the region loop
    opts = cb.CoverageOptions(window_length=wl, min_mapq=20)
    t = pipeline.coverage_run([pipeline.Region("c", L, batch, V["seq"])], opts, ctx=ctx)
    assert np.array_equal(to_np(t.rcv), V["gw_counts"].astype(np.float32))
    assert [t.num_processed[0], t.num_skipped[0]] == V["gw_tally"].tolist()
    assert_f32_equal_bits(t.gc, V["gc"])
    assert np.array_equal(to_np(t.gap).astype(bool), V["gap"].astype(bool))
    assert_f32_equal_bits(t.lcv, V["lcv"])
    assert np.array_equal(to_np(t.ft), V["ft"])
    # A3
    o3 = cb.CoverageOptions(window_length=None, min_mapq=20, considered_regions="TargetRegions")
    t3 = pipeline.coverage_run([pipeline.Region("c", L, batch, None, targets=(V["tgt_start"], V["tgt_end"]))], o3, ctx=ctx)
    assert np.array_equal(to_np(t3.rcv), V["tgt_counts"].astype(np.float32))
    # A4
    o4 = cb.CoverageOptions(window_length=wl, min_mapq=20, count_kind="Coverage")
    t4 = pipeline.coverage_run([pipeline.Region("c", L, batch, None)], o4, ctx=ctx)
    assert_f32_equal_bits(t4.rcv, V["cov_mean"])
    np.testing.assert_allclose(to_np(t4.rcvsd), V["cov_sd"], rtol=1e-6)
    # N1 / N2
    cv, _, s = cb.run_uncorrected_normalization(V["lcv"], None, ctx=ctx)
    assert s == float(V["n1_sum"])
    assert_f32_equal_bits(cv, V["n1_cv"])
    r = cb.run_binned_correction(V["n2_lcv"], V["n2_gc"], None, ctx=ctx)
    assert np.array_equal(to_np(r["flags"]), V["n2_flags"])
    has = (V["n2_flags"] & 1) != 0
    assert_f32_equal_bits(to_np(r["cv"])[has], V["n2_cv"][has])
    # W2-W5, both engines
    for engine in (mw.ENGINE_EXACT, mw.ENGINE_TENSOR):
        m = mw.run(V["wis_X"], V["wis_tid"], mw.BuildModelWisOptions(max_ref_targets=30, engine=engine), ctx=ctx)
        assert np.array_equal(to_np(m.ref_len), V["wis_ref_len"])
        assert_f32_equal_bits(to_np(m.ref_dist), V["wis_dist"])
        assert np.float32(m.threshold).tobytes() == np.float32(V["wis_thr"]).tobytes()
        assert np.array_equal(to_np(m.num_calls), V["wis_calls"]) and np.array_equal(to_np(m.filt), V["wis_filt"])
        for row in range(0, 600, 37):
            n = int(V["wis_ref_len"][row])
            assert np.array_equal(to_np(m.ref_idx)[row, :n], V["wis_ref_idx"][row, :n])
    # M1
    g = mw.mod_coverage(V["wis_X"][:, 0].copy(), (V["wis_filt"] != 0).astype(np.uint8), V["wis_ref_idx"], V["wis_ref_len"], ctx=ctx)
    assert np.array_equal(to_np(g["status"]), V["modcov_status"])
    assert_f32_equal_bits(g["cv_rel"], V["modcov_cv_rel"])
    # H and piles
    st = discover.hmm_segment(V["hmm_cv"], [0, V["hmm_cv"].size], [0.08], ctx=ctx)
    assert np.array_equal(to_np(st), V["hmm_state"])
    p = piles.PileCollector(20, 20, ctx=ctx).collect_piles(batch, L)
    assert np.array_equal(p.start, V["pile_start"]) and np.array_equal(p.end, V["pile_end"]) and np.array_equal(p.size, V["pile_size"])
