"""BASELINE configurations at FULL size against the oracle (not against a re-statement of our own).

  config 1  chr22, 15.2 M reads, 1 kbp windows, MAPQ >= 20: all three aggregator kinds
            (lib-coverage/src/agg.rs:115-198, 301-409, 485-574) -- counts, tallies, means bit-exact,
            the Coverage kind's SD to rel 1e-6
  config 2  chr19-22 of the 30x genome (69 M reads, 500 bp windows) through the region loop +
            MedianGcBinned: counts, GC bits, gap flags, medians, CV bits
            (lib-coverage/src/lib.rs:768-781, lib-normalize/src/correct_binned.rs:25-146)
  config 3  200 000 x 100 panel: rows spread over the panel against the oracle's compute_distances
            (lib-model-wis/src/lib.rs:128-197)
  config 5  (reduced T for the test suite; bench.py --workload cfg5 does it at 3 M x 1000) the same
            with 1000 samples, the streamed tensor sweep
The oracle does 2-4e7 reads/s and ~1e9 sample-pairs/s per thread, so each case stays within seconds.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import cnvetti_b200 as cb
from cnvetti_b200 import model_wis as mw
from cnvetti_b200 import pipeline, synth, synth_torch
from cnvetti_b200.synth import CHR22_LEN, HG38_AUTOSOMES
from util import assert_f32_equal_bits, to_np

pytestmark = pytest.mark.gpu
RTOL = 1e-6


@pytest.fixture(scope="module")
def ctx():
    return cb.Context(0)


@pytest.fixture(scope="module")
def cfg1(oracle):
    """config 1's input on the device and, as BAM-level records, for the oracle"""
    L = CHR22_LEN
    soa = synth_torch.synth_soa_contig(22, L, 7_600_000, "cuda")
    Lc = int(soa["end"].max().item()) + 1000
    d = synth_torch.soa_to_oracle_records(soa)
    reads = oracle.ReadBatch(d["pos"], d["isize"], d["flag"], d["mapq"], d["cigar_off"], d["cigar"])
    batch = cb.ReadBatch(**{k: soa[k].contiguous() for k in ("pos", "end", "mid", "frag_end", "mapq", "flags")})
    return dict(L=Lc, batch=batch, reads=reads, n=len(batch))


def test_cfg1_fragments_genome_wide_full_size(oracle, ctx, cfg1):
    opts = cb.CoverageOptions(window_length=1000, min_mapq=20)
    t = pipeline.coverage_run([pipeline.Region("chr22", cfg1["L"], cfg1["batch"])], opts, ctx=ctx)
    oopts = oracle.cov_opts(window_length=1000, min_mapq=20)
    counts, proc, skip = oracle.frag_gw(cfg1["reads"], oopts, cfg1["L"])
    cov, start, end, frm = oracle.frag_gw_stats(counts, oopts, cfg1["L"])
    lcv, _, ft = oracle.cov_records(cov, None, start, end, frm, False, oopts)
    assert cfg1["n"] == 15_200_000 and proc > 5_000_000
    assert t.num_processed == [proc] and t.num_skipped == [skip]
    assert np.array_equal(to_np(t.rcv).astype(np.int64), counts.astype(np.int64))
    assert_f32_equal_bits(t.rcv, cov, "RCV")
    assert_f32_equal_bits(t.lcv, lcv, "LCV")
    assert_f32_equal_bits(t.frm, frm, "FRM")
    assert np.array_equal(to_np(t.start), start.astype(np.int32)) and np.array_equal(to_np(t.end), end.astype(np.int32))
    assert np.array_equal(to_np(t.ft), ft)


def test_cfg1_fragments_targets_full_size(oracle, ctx, cfg1):
    ts, te = synth.synth_targets(22, CHR22_LEN, 4010)
    opts = cb.CoverageOptions(window_length=None, min_mapq=20, considered_regions="TargetRegions")
    t = pipeline.coverage_run([pipeline.Region("chr22", cfg1["L"], cfg1["batch"], targets=(ts, te))], opts, ctx=ctx)
    oopts = oracle.cov_opts(window_length=0, min_mapq=20)
    counts, proc, skip = oracle.frag_targets(cfg1["reads"], oopts, ts, te)
    assert t.num_processed == [proc] and t.num_skipped == [skip] and counts.sum() > 100_000
    assert np.array_equal(to_np(t.rcv).astype(np.int64), counts.astype(np.int64))
    lcv, _, ft = oracle.cov_records(counts.astype(np.float32), None, ts, te, None, True, oopts)
    assert_f32_equal_bits(t.lcv, lcv, "LCV")
    assert np.array_equal(to_np(t.ft), ft)


def test_cfg1_coverage_kind_full_size(oracle, ctx, cfg1):
    opts = cb.CoverageOptions(window_length=1000, min_mapq=20, count_kind="Coverage")
    t = pipeline.coverage_run([pipeline.Region("chr22", cfg1["L"], cfg1["batch"])], opts, ctx=ctx)
    oopts = oracle.cov_opts(window_length=1000, min_mapq=20)
    mean, sd = oracle.coverage(cfg1["reads"], oopts, cfg1["L"])
    assert mean.size == to_np(t.rcv).size and float(mean.max()) > 20.0
    assert_f32_equal_bits(t.rcv, mean, "RCV (window mean)")
    np.testing.assert_allclose(to_np(t.rcvsd), sd, rtol=RTOL, atol=0)
    assert t.num_processed == [0] and t.num_skipped == [0]  # agg.rs:590-598: the tallies stay 0


def test_cfg2_sample_contigs_full_size(oracle, ctx):
    """chr19..chr22 of config 2 exactly as bench.py generates them (rank 0), region loop + MedianGcBinned
    over these four contigs, against the oracle's coverage + refstats + record assembly + normaliser."""
    wl, minq = 500, 20
    ids = [18, 19, 20, 21]
    regions, oparts = [], []
    for ci in ids:
        L = HG38_AUTOSOMES[ci]
        n_pairs = int(round(L * (900e6 / 2_875_001_522) / 2))
        soa = synth_torch.synth_soa_contig(2 * 1000 + ci, L, n_pairs, "cuda")
        seq = synth_torch.synth_reference_torch(2 * 1000 + ci, L, "cuda")
        batch = cb.ReadBatch(mid=soa["mid"].contiguous(), mapq=soa["mapq"].contiguous(), flags=soa["flags"].contiguous())
        regions.append(pipeline.Region("chr%d" % (ci + 1), L, batch, seq))
        oparts.append((synth_torch.soa_to_oracle_records(soa), L, seq.cpu().numpy()))
        del soa
    opts = cb.CoverageOptions(window_length=wl, min_mapq=minq)
    t = pipeline.normalize_run(pipeline.coverage_run(regions, opts, ctx=ctx), "MedianGcBinned", ctx=ctx)
    oopts = oracle.cov_opts(window_length=wl, min_mapq=minq)
    lcvs, gcs, gaps, rcvs, procs = [], [], [], [], []
    for d, L, seq in oparts:
        reads = oracle.ReadBatch(d["pos"], d["isize"], d["flag"], d["mapq"], d["cigar_off"], d["cigar"])
        counts, proc, skip = oracle.frag_gw(reads, oopts, L)
        cov, start, end, frm = oracle.frag_gw_stats(counts, oopts, L)
        lcv, _, _ = oracle.cov_records(cov, None, start, end, frm, False, oopts)
        gc, gap = oracle.refstats(seq, wl)
        lcvs.append(lcv), gcs.append(gc), gaps.append(gap), rcvs.append(cov), procs.append(proc)
    lcv, gc = np.concatenate(lcvs), np.concatenate(gcs)
    cv, cv2, _, flags, med = oracle.norm_gc_median(lcv, gc)
    assert sum(procs) > 20_000_000 and t.num_processed == procs
    assert_f32_equal_bits(t.rcv, np.concatenate(rcvs), "RCV")
    assert_f32_equal_bits(t.lcv, lcv, "LCV")
    assert_f32_equal_bits(t.gc, gc, "GC")
    assert np.array_equal(to_np(t.gap), np.concatenate(gaps))
    assert np.array_equal(to_np(t.norm_flags), flags)
    assert_f32_equal_bits(t.cv, cv, "CV")
    g2, fin = to_np(t.cv2), np.isfinite(cv2)
    assert np.array_equal(np.isfinite(g2), fin)
    np.testing.assert_allclose(g2[fin], cv2[fin], rtol=RTOL, atol=1e-7)


def _oracle_rows(oracle, X, tids, rows, k):
    """compute_distances of single rows, one oracle call per row on a thread each (ctypes drops the GIL)"""
    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        return list(ex.map(lambda r: oracle.wis_distances(X, tids, k, r, r + 1), rows))


@pytest.mark.parametrize("T,S,n_rows", [(200_000, 100, 48), (150_000, 1000, 16)])
def test_wis_tensor_rows_against_oracle_full_size(oracle, ctx, T, S, n_rows):
    """config 3 at full size / config 5's sample count: the tensor engine's reference sets, order and
    distances of rows spread over the panel (first, last, every contig) against the ORACLE."""
    import torch
    if S == 100:
        X, tids = synth.synth_panel(3, T, S)
        dX, dt = torch.from_numpy(X).cuda(), torch.from_numpy(tids.view(np.int32)).cuda()
    else:
        lcv, dt = synth_torch.synth_cohort_lcv(5, T, 0, S, "cuda")
        cv = torch.empty_like(lcv)
        from cnvetti_b200.normalize import run_uncorrected_normalization
        for s in range(S):
            run_uncorrected_normalization(lcv[s], ctx=ctx, want_sum=False, out=cv[s])
        dX = cv.t().contiguous()
        del lcv, cv
        X, tids = dX.cpu().numpy(), dt.cpu().numpy().astype(np.uint32)
    opts = mw.BuildModelWisOptions(engine=mw.ENGINE_TENSOR)
    dist, idx = mw.compute_distances(dX, dt, opts, ctx=ctx)
    stats = mw.last_stats(ctx)
    assert stats["fallback_rows"] == 0
    rows = sorted(set(np.linspace(0, T - 1, n_rows).astype(np.int64).tolist()))
    got_d, got_i = to_np(dist[rows]), to_np(idx[rows]).astype(np.uint32)
    for q, (od, oi) in enumerate(_oracle_rows(oracle, X, tids, rows, 100)):
        assert np.array_equal(got_i[q], oi[0]), "row %d: reference set / order" % rows[q]
        assert np.array_equal(got_d[q].view(np.uint32), od[0].view(np.uint32)), "row %d: distances" % rows[q]


def test_wis_nan_panel_is_a_reference_panic(ctx):
    """lib-model-wis/src/lib.rs:168: `partial_cmp().unwrap()` panics on a NaN distance -- a panel with a
    missing CV must not come back as a model (both engines, host and device panels)."""
    import torch
    from cnvetti_b200 import _native as N
    for T, engine in ((300, mw.ENGINE_EXACT), (6000, mw.ENGINE_TENSOR), (6000, mw.ENGINE_AUTO)):
        X, tids = synth.synth_panel(11, T, 24, n_contigs=4)
        X[T // 2, 7] = np.nan
        opts = mw.BuildModelWisOptions(max_ref_targets=20, engine=engine)
        for dev in (False, True):
            a = (torch.from_numpy(X).cuda(), torch.from_numpy(tids.view(np.int32)).cuda()) if dev else (X, tids)
            with pytest.raises(N.ReferencePanic) as e:
                mw.compute_distances(a[0], a[1], opts, ctx=ctx)
            assert "lib.rs:168" in str(e.value)
            with pytest.raises(N.ReferencePanic):
                mw.run(a[0], a[1], opts, ctx=ctx)
    # one contig only: no pair is ever evaluated (lib.rs:153-156), the reference does not panic
    X, _ = synth.synth_panel(12, 64, 8, n_contigs=1)
    X[3, 3] = np.nan
    d, _ = mw.compute_distances(X, np.zeros(64, np.uint32), mw.BuildModelWisOptions(max_ref_targets=10), ctx=ctx)
    assert np.isinf(d).all()


def test_wis_infinite_panel_values(oracle, ctx):
    """+inf in one target: its distances are +inf, nothing panics; the same infinity in two targets of
    different contigs: inf - inf = NaN, the reference panics (lib.rs:157-168)."""
    from cnvetti_b200 import _native as N
    X, tids = synth.synth_panel(13, 5000, 16, n_contigs=3)
    X[100, 2] = np.inf
    opts = mw.BuildModelWisOptions(max_ref_targets=20, engine=mw.ENGINE_TENSOR)
    d, i = mw.compute_distances(X, tids, opts, ctx=ctx)
    od, oi = oracle.wis_distances(X, tids, 20, nthreads=os.cpu_count() or 1)
    assert np.array_equal(d.view(np.uint32), od.view(np.uint32)) and np.array_equal(i, oi)
    j = int(np.flatnonzero(tids != tids[100])[0])
    X[j, 2] = np.inf
    with pytest.raises(N.ReferencePanic):
        mw.compute_distances(X, tids, opts, ctx=ctx)
