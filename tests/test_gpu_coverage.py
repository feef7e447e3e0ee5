"""Parity of the CUDA coverage path (through the C ABI) against the oracle.

Integer outputs (fragment counts, tallies, window bounds, filter bits) must be bit-exact; f32
outputs that are a single IEEE operation on exact inputs (window means, LCV, GC ratio, CV of the
GC-binned normaliser) must also be bit-exact; the remaining floating-point outputs are held to the
north_star tolerance rel 1e-6.
"""
import numpy as np
import pytest

import cnvetti_b200 as cb
from cnvetti_b200 import synth
from util import assert_f32_equal_bits, oracle_reads, simple_reads, soa_batch, to_np

pytestmark = pytest.mark.gpu
RTOL = 1e-6  # north_star: normalised depths / z-scores / posteriors within rel 1e-6


@pytest.fixture(scope="module")
def ctx():
    return cb.Context(0)


# ---- A2 FragmentsGenomeWide ---------------------------------------------------------------
@pytest.mark.parametrize("wl,min_mapq,device", [(1000, 20, None), (500, 0, "cuda"), (137, 20, "cuda"),
                                                (1, 30, None), (4096, 0, "cuda")])
def test_frag_gw_parity(oracle, ctx, wl, min_mapq, device):
    L = 2_000_000
    d = synth.synth_bam_records(22, L, 60_001)
    oopts = oracle.cov_opts(window_length=wl, min_mapq=min_mapq)
    counts, proc, skip = oracle.frag_gw(oracle_reads(oracle, d), oopts, L)
    opts = cb.CoverageOptions(window_length=wl, min_mapq=min_mapq)
    agg = cb.FragmentsGenomeWideAggregator(None, opts, L, ctx=ctx)
    agg.put_fetched_records(soa_batch(d, device=device))
    assert agg.num_regions() == counts.size
    assert np.array_equal(to_np(agg.counts()), counts)
    assert agg.num_processed() == proc and agg.num_skipped() == skip
    s = agg.get_all_stats()
    cov, start, end, frm = oracle.frag_gw_stats(counts, oopts, L)
    assert_f32_equal_bits(s.cov, cov, "cov")
    assert np.array_equal(to_np(s.start), start) and np.array_equal(to_np(s.end), end)
    assert_f32_equal_bits(s.frac_removed, frm, "frm")
    assert (to_np(s.mean_mapq) == 60.0).all() and s.cov_sd is None
    # A6 record assembly
    lcv, lcvsd, ft = cb.cov_records(s, opts, False, ctx=ctx)
    olcv, _, oft = oracle.cov_records(cov, None, start, end, frm, False, oopts)
    assert_f32_equal_bits(lcv, olcv, "lcv")
    assert lcvsd is None and np.array_equal(to_np(ft), oft)


@pytest.mark.parametrize("wl,min_mapq", [(8, 0), (16, 127), (16, 128), (500, 200), (32768, 255), (40_000, 1),
                                         (1000, 60)])
def test_frag_gw_packed_filter_adversarial(oracle, ctx, wl, min_mapq):
    """The kernel tests four mapq bytes / two flag words per instruction and forces failing reads
    out of the window arithmetic: every mapq byte value, random flag bits, reads far from their
    neighbours (per-read path), window lengths either side of the multiply-high fast path."""
    L0, L = 1_400_000, 3_000_000  # unpaired centres are pos + end_pos: keep them inside the contig
    d = synth.synth_bam_records(7, L0, 50_000)
    rng = np.random.default_rng(1000 * wl + min_mapq)
    n = d["pos"].size
    d["mapq"] = rng.integers(0, 256, n).astype(np.uint8)
    d["flag"] = (rng.integers(0, 4096, n) & rng.integers(0, 4096, n)).astype(np.uint16)
    far = rng.random(n) < 0.05
    d["pos"] = np.where(far, rng.integers(0, L0 - 1000, n), d["pos"]).astype(np.int32)
    oopts = oracle.cov_opts(window_length=wl, min_mapq=min_mapq)
    counts, proc, skip = oracle.frag_gw(oracle_reads(oracle, d), oopts, L)
    assert proc > 0
    opts = cb.CoverageOptions(window_length=wl, min_mapq=min_mapq)
    for device in (None, "cuda"):
        agg = cb.FragmentsGenomeWideAggregator(None, opts, L, ctx=ctx)
        agg.put_fetched_records(soa_batch(d, device=device))
        assert np.array_equal(to_np(agg.counts()), counts)
        assert agg.num_processed() == proc and agg.num_skipped() == skip


def test_frag_gw_streaming_unaligned_and_piles(oracle, ctx):
    L = 1_500_000
    d = synth.synth_bam_records(5, L, 40_000)
    ps = np.array([10_000, 500_000, 900_123, 1_400_000], np.uint32)
    pe = np.array([12_000, 500_050, 903_456, 1_400_001], np.uint32)
    oopts = oracle.cov_opts(window_length=500, min_mapq=10)
    counts, proc, skip = oracle.frag_gw(oracle_reads(oracle, d), oopts, L, ps, pe)
    assert skip > 0
    opts = cb.CoverageOptions(window_length=500, min_mapq=10)
    n = d["pos"].size
    for device in (None, "cuda"):
        agg = cb.FragmentsGenomeWideAggregator((ps, pe), opts, L, ctx=ctx)
        # three batches with odd boundaries: exercises the scalar (unaligned) kernel and tails
        cuts = [0, 12_345, 12_346, 50_003, n]
        full = soa_batch(d, device=device)
        for a, b in zip(cuts[:-1], cuts[1:]):
            agg.put_fetched_records(cb.ReadBatch(mid=full.mid[a:b], mapq=full.mapq[a:b], flags=full.flags[a:b]))
        agg.put_fetched_records(cb.ReadBatch(mid=full.mid[:0], mapq=full.mapq[:0], flags=full.flags[:0]))
        assert np.array_equal(to_np(agg.counts()), counts)
        assert agg.num_processed() == proc and agg.num_skipped() == skip
        s = agg.get_all_stats(device=device is not None)
        cov, start, end, frm = oracle.frag_gw_stats(counts, oopts, L, ps, pe)
        assert_f32_equal_bits(s.frac_removed, frm, "frm")
        lcv, _, ft = cb.cov_records(s, opts, False, ctx=ctx)
        _, _, oft = oracle.cov_records(cov, None, start, end, frm, False, oopts)
        assert np.array_equal(to_np(ft), oft) and oft.any()


def test_frag_gw_reference_panic(oracle, ctx):
    d = simple_reads([50_000], [50_100])  # unpaired: centre = pos + end_pos, out of range
    with pytest.raises(oracle.OraclePanic):
        oracle.frag_gw(oracle_reads(oracle, d), oracle.cov_opts(window_length=1000), 60_000)
    agg = cb.FragmentsGenomeWideAggregator(None, cb.CoverageOptions(window_length=1000), 60_000, ctx=ctx)
    agg.put_fetched_records(soa_batch(d))
    with pytest.raises(cb.ReferencePanic) as e:
        agg.finish()
    assert "agg.rs:144" in str(e.value)


def test_errors_are_loud(ctx):
    with pytest.raises(cb.CnvbError):
        cb.FragmentsGenomeWideAggregator(None, cb.CoverageOptions(window_length=None), 1000, ctx=ctx)
    agg = cb.CoverageAggregator(cb.CoverageOptions(window_length=100), 1000, ctx=ctx)
    with pytest.raises(cb.CnvbError):
        agg.put_fetched_records(cb.ReadBatch(mid=np.zeros(4, np.int32), mapq=np.zeros(4, np.uint8),
                                             flags=np.zeros(4, np.uint16)))


# ---- A3 FragmentsTargetRegions -------------------------------------------------------------
@pytest.mark.parametrize("seed,overlap,device", [(1, 0.0, None), (2, 0.0, "cuda"), (3, 0.5, "cuda")])
def test_frag_targets_parity(oracle, ctx, seed, overlap, device):
    L = 1_000_000
    d = synth.synth_bam_records(seed, L, 30_000, unpaired_frac=0.1)
    ts, te = synth.synth_targets(seed, L, 1500, overlap_frac=overlap)
    oopts = oracle.cov_opts(min_mapq=0)
    counts, proc, skip = oracle.frag_targets(oracle_reads(oracle, d), oopts, ts, te)
    assert counts.sum() > 1000 and skip > 0
    opts = cb.CoverageOptions(considered_regions="TargetRegions", min_mapq=0)
    agg = cb.FragmentsTargetRegionsAggregator((ts, te), opts, L, ctx=ctx)
    b = soa_batch(d, device=device)
    half = len(b) // 2 + 1
    for sl in (slice(0, half), slice(half, None)):
        agg.put_fetched_records(cb.ReadBatch(pos=b.pos[sl], frag_end=b.frag_end[sl], mapq=b.mapq[sl], flags=b.flags[sl]))
    assert np.array_equal(to_np(agg.counts()), counts)
    assert agg.num_processed() == proc and agg.num_skipped() == skip
    s = agg.get_all_stats()
    assert np.array_equal(to_np(s.start), ts) and np.array_equal(to_np(s.end), te)
    assert s.frac_removed is None and s.cov_sd is None
    lcv, _, ft = cb.cov_records(s, opts, True, ctx=ctx)
    olcv, _, oft = oracle.cov_records(counts.astype(np.float32), None, ts, te, None, True, oopts)
    assert_f32_equal_bits(lcv, olcv, "lcv")
    assert np.array_equal(to_np(ft), oft) and oft.any()


def test_frag_targets_tie_rule(oracle, ctx):
    ts, te = np.array([100, 301], np.uint32), np.array([200, 400], np.uint32)
    d = simple_reads([150], [250])
    d["flag"][:] = 0x3 | 0x40
    d["isize"][:] = 200
    counts, _, _ = oracle.frag_targets(oracle_reads(oracle, d), oracle.cov_opts(), ts, te)
    agg = cb.FragmentsTargetRegionsAggregator((ts, te), cb.CoverageOptions(considered_regions="TargetRegions"), 1000, ctx=ctx)
    agg.put_fetched_records(soa_batch(d))
    assert to_np(agg.counts()).tolist() == counts.tolist() == [1, 0]


# ---- A4 Coverage ------------------------------------------------------------------------------
def _check_coverage(oracle, ctx, d, wl, L, min_mapq=20, device=None):
    oopts = oracle.cov_opts(window_length=wl, min_mapq=min_mapq)
    mean, sd = oracle.coverage(oracle_reads(oracle, d), oopts, L)
    opts = cb.CoverageOptions(window_length=wl, min_mapq=min_mapq, count_kind="Coverage")
    agg = cb.CoverageAggregator(opts, L, ctx=ctx)
    b = soa_batch(d, device=device)
    agg.put_fetched_records(cb.ReadBatch(pos=b.pos, end=b.end, mapq=b.mapq, flags=b.flags))
    s = agg.get_all_stats()
    assert_f32_equal_bits(s.cov, mean, "window mean")           # exact integer sum / wl
    np.testing.assert_allclose(to_np(s.cov_sd), sd, rtol=RTOL, atol=0)
    assert agg.num_processed() == 0 and agg.num_skipped() == 0  # agg.rs: never incremented
    return s, mean, sd


@pytest.mark.parametrize("wl,device", [(1000, None), (500, "cuda"), (137, "cuda"), (10_000, None)])
def test_coverage_parity(oracle, ctx, wl, device):
    L = 1_200_000
    d = synth.synth_bam_records(11, L, 40_000, unpaired_frac=0.05, refskip_frac=0.002)
    s, mean, sd = _check_coverage(oracle, ctx, d, wl, L, device=device)
    assert (mean > 0).sum() > 10
    lcv, lcvsd, ft = cb.cov_records(s, cb.CoverageOptions(window_length=wl), False, ctx=ctx)
    ol, osd, _ = oracle.cov_records(mean, sd, to_np(s.start), to_np(s.end), None, False, oracle.cov_opts(window_length=wl))
    assert_f32_equal_bits(lcv, ol, "lcv")
    np.testing.assert_allclose(to_np(lcvsd), osd, rtol=RTOL)


@pytest.mark.parametrize("seed", range(12))
def test_coverage_state_machine_random(oracle, ctx, seed):
    """sparse coverage: empty windows, single-column windows, last-window corner cases"""
    rng = np.random.default_rng(100 + seed)
    wl = int(rng.integers(2, 40))
    L = int(rng.integers(5, 60)) * wl + int(rng.integers(0, wl))
    n = int(rng.integers(0, 40))
    pos = np.sort(rng.integers(0, max(1, L - 2), n))
    end = np.minimum(pos + rng.integers(1, 3 * wl, n), L)
    flag = rng.choice(np.array([0, 0, 0, 0x400, 0x100, 0x800, 0x200, 0x4, 0x1, 0x3], np.uint16), n)
    d = simple_reads(pos, end, flag, rng.integers(0, 61, n))
    _check_coverage(oracle, ctx, d, wl, L)


def test_coverage_corner_cases(oracle, ctx):
    # exactly the SURVEY A.3 example, one column only, two columns in one window, lone last column
    cases = [
        (np.arange(14), np.full(14, 14), 4, 14),
        ([5], [6], 4, 20),
        ([5], [7], 4, 20),
        ([1, 9], [3, 10], 4, 20),
        ([1, 4], [2, 5], 4, 20),
        ([], [], 7, 50),
        ([0], [50], 7, 50),
    ]
    for pos, end, wl, L in cases:
        _check_coverage(oracle, ctx, simple_reads(pos, end), wl, L, min_mapq=0)


def test_coverage_past_region_end_panics(oracle, ctx):
    d = simple_reads([90], [130])
    with pytest.raises(oracle.OraclePanic):
        oracle.coverage(oracle_reads(oracle, d), oracle.cov_opts(window_length=10), 100)
    agg = cb.CoverageAggregator(cb.CoverageOptions(window_length=10, count_kind="Coverage"), 100, ctx=ctx)
    agg.put_fetched_records(soa_batch(d))
    with pytest.raises(cb.ReferencePanic):
        agg.finish()


# ---- A5 ReferenceStats ---------------------------------------------------------------------------
@pytest.mark.parametrize("wl,off,device", [(1000, 0, None), (500, 0, "cuda"), (500, 3, "cuda"), (16, 1, "cuda"),
                                           (7, 0, None), (333, 5, "cuda"), (100_000, 0, "cuda")])
def test_refstats_parity(oracle, ctx, wl, off, device):
    seq = synth.synth_reference(3, 1_234_567 + off, block=5000)[off:]
    gc, gap = oracle.refstats(seq, wl)
    if device:
        import torch
        full = torch.from_numpy(synth.synth_reference(3, 1_234_567 + off, block=5000)).to(device)
        rs = cb.ReferenceStats.look_at_chars(full[off:], wl, ctx=ctx)   # misaligned device pointer
    else:
        rs = cb.ReferenceStats.look_at_chars(seq, wl, ctx=ctx)
    assert_f32_equal_bits(rs.gc_content, gc, "gc")
    assert np.array_equal(to_np(rs.has_gap), gap)
    if wl <= 1000:
        assert np.isnan(gc).any() and gap.any()


def test_refstats_all_bytes(oracle, ctx):
    seq = np.tile(np.arange(256, dtype=np.uint8), 40)
    for wl in (16, 64, 100):
        gc, gap = oracle.refstats(seq, wl)
        rs = cb.ReferenceStats.look_at_chars(seq, wl, ctx=ctx)
        assert_f32_equal_bits(rs.gc_content, gc, "gc")
        assert np.array_equal(to_np(rs.has_gap), gap)


# ---- N1 / N2 -----------------------------------------------------------------------------------------
def _lcv_gc(seed, n):
    rng = np.random.default_rng(seed)
    gc = rng.beta(8, 11, n).astype(np.float32)
    gc[rng.random(n) < 0.01] = np.frombuffer(np.uint32(0x7F800001).tobytes(), np.float32)[0]
    gc[:3] = [0.0, 1.0, 0.999]
    lcv = (rng.poisson(60 * (0.5 + np.nan_to_num(gc.astype(np.float64)).clip(0, 1)), n) / 500.0).astype(np.float32)
    lcvsd = rng.random(n).astype(np.float32)
    return lcv, gc, lcvsd


@pytest.mark.parametrize("device", [None, "cuda"])
def test_norm_total_parity(oracle, ctx, device):
    lcv, _, lcvsd = _lcv_gc(1, 200_003)
    s, cv, cvsd = oracle.norm_total(lcv, lcvsd)
    if device:
        import torch
        a, b = torch.from_numpy(lcv).to(device), torch.from_numpy(lcvsd).to(device)
    else:
        a, b = lcv, lcvsd
    gcv, gcvsd, gs = cb.run_uncorrected_normalization(a, b, ctx=ctx)
    assert gs == s                      # double-double sum rounds to the exact (Shewchuk) sum
    assert_f32_equal_bits(gcv, cv, "cv")
    assert_f32_equal_bits(gcvsd, cvsd, "cvsd")
    gcv2, none, _ = cb.run_uncorrected_normalization(a, None, ctx=ctx)
    assert none is None
    assert_f32_equal_bits(gcv2, cv, "cv")


@pytest.mark.parametrize("device,n", [(None, 150_001), ("cuda", 150_001), ("cuda", 57), (None, 0)])
def test_norm_gc_median_parity(oracle, ctx, device, n):
    lcv, gc, lcvsd = _lcv_gc(2, max(n, 3))
    lcv, gc, lcvsd = lcv[:n], gc[:n], lcvsd[:n]
    cv, cv2, cvsd, flags, med = oracle.norm_gc_median(lcv, gc, lcvsd)
    if device:
        import torch
        a, g, b = (torch.from_numpy(x).to(device) for x in (lcv, gc, lcvsd))
    else:
        a, g, b = lcv, gc, lcvsd
    r = cb.run_binned_correction(a, g, b, ctx=ctx)
    assert_f32_equal_bits(r["medians"], med, "medians")
    assert np.array_equal(to_np(r["flags"]), flags)
    assert_f32_equal_bits(r["cv"], cv, "cv")
    assert_f32_equal_bits(r["cvsd"], cvsd, "cvsd")
    g2, o2 = to_np(r["cv2"]), cv2
    fin = np.isfinite(o2)
    assert np.array_equal(np.isfinite(g2), fin)
    np.testing.assert_allclose(g2[fin], o2[fin], rtol=RTOL, atol=1e-7)
    if n > 1000:
        assert (flags & oracle.N2_FEW_GC).any() and (flags & oracle.N2_CV).any() and (flags == 0).any()


# ---- BASELINE sizes: size-independent checks against an independent torch model --------------------
def test_full_size_contig_fragments_and_refstats_against_torch_model(ctx):
    """One config-2 contig at full size (chr1: 249 Mbp, 78 M reads, 500 bp windows) -- far beyond what the
    CPU oracle finishes in seconds.  The filter, the window index and the per-window GC fraction are
    restated with plain torch tensor operations (an independent model: no shared code with the kernels)
    and must agree exactly: counts per window, the processed tally, GC bits and gap flags."""
    import torch
    from cnvetti_b200 import synth_torch
    from cnvetti_b200.synth import HG38_AUTOSOMES
    L, wl, minq = HG38_AUTOSOMES[0], 500, 20
    soa = synth_torch.synth_soa_contig(2001, L, int(L * 0.313 / 2), "cuda")
    n_bins = (L + wl - 1) // wl
    agg = cb.FragmentsGenomeWideAggregator(None, cb.CoverageOptions(window_length=wl, min_mapq=minq), L, ctx=ctx)
    agg.put_fetched_records(cb.ReadBatch(mid=soa["mid"].contiguous(), mapq=soa["mapq"].contiguous(),
                                         flags=soa["flags"].contiguous()))
    counts = agg.counts()
    fl = soa["flags"].to(torch.int32) & 0xFFFF
    paired = (fl & 1) != 0
    ok = (soa["mapq"].to(torch.int32) >= minq) & ((fl & (0x100 | 0x200 | 0x400 | 0x800 | 0x1000)) == 0)
    ok &= ~paired | (((fl & 0x2) != 0) & ((fl & 0x2000) == 0))
    model = torch.bincount((soa["mid"][ok] // wl).to(torch.int64), minlength=n_bins)
    assert model.numel() == n_bins and int(ok.sum()) > 20_000_000
    assert torch.equal(torch.as_tensor(counts).to(torch.int64).cuda(), model)
    assert agg.num_processed() == int(ok.sum()) and agg.num_skipped() == 0
    del soa, fl, ok, model
    # reference statistics of the same contig
    seq = synth_torch.synth_reference_torch(2001, L, "cuda")
    rs = cb.ReferenceStats.look_at_chars(seq, wl, ctx=ctx)
    gc, gap = rs.gc_content, rs.has_gap
    pad = n_bins * wl - L
    s = torch.cat([seq, torch.zeros(pad, dtype=torch.uint8, device="cuda")]).view(n_bins, wl)
    up = s & 0xDF
    is_gc = ((up == ord("G")) | (up == ord("C"))).sum(1)
    is_acgt = ((up == ord("A")) | (up == ord("C")) | (up == ord("G")) | (up == ord("T"))).sum(1)
    is_n = (up == ord("N")).sum(1)
    want = is_gc.to(torch.float32) / is_acgt.to(torch.float32)          # one IEEE division of exact integers
    got = torch.as_tensor(gc).cuda()
    have = is_acgt > 0
    assert torch.equal(got[have].view(torch.int32), want[have].view(torch.int32))
    assert bool((got[~have].view(torch.int32) == 0x7F800001).all())
    assert torch.equal(torch.as_tensor(gap).cuda().bool(), is_n > 0)
