"""The native region loop (cnvb_coverage_run = lib_coverage::run's loop, lib.rs:768-781) against the
oracle run region by region, plus the chained normaliser; host and device inputs, all three kinds."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import cnvetti_b200 as cb
from cnvetti_b200 import _native as N
from cnvetti_b200 import pipeline, synth
from util import assert_f32_equal_bits, oracle_reads, simple_reads, soa_batch, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return cb.Context(0)


def _regions(device, with_targets=False):
    import torch
    out, raw = [], []
    for i, L in enumerate((700_123, 310_000, 50_000)):
        d = synth.synth_bam_records(50 + i, L, 9_000 if L > 100_000 else 600)
        seq = synth.synth_reference(60 + i, L)
        tg = synth.synth_targets(70 + i, L, 40 if L > 100_000 else 5) if with_targets else None
        s = torch.from_numpy(seq).to(device) if device else seq
        out.append(pipeline.Region("chr%d" % (i + 1), L, soa_batch(d, device=device), s, targets=tg))
        raw.append((d, seq, tg, L))
    return out, raw


@pytest.mark.parametrize("device,batch", [(None, "0"), ("cuda", "0"), ("cuda", "1")])
def test_region_loop_fragments_gw(oracle, ctx, device, batch, monkeypatch):
    """(batch = 1: the records of all regions of a group in one launch, CNVB_GW_BATCH -- off by default, run.cu.)"""
    monkeypatch.setenv("CNVB_GW_BATCH", batch)
    monkeypatch.setenv("CNVB_GW_GROUPS", "2")
    wl = 500
    regions, raw = _regions(device)
    opts = cb.CoverageOptions(window_length=wl, min_mapq=10)
    t = pipeline.coverage_run(regions, opts, ctx=ctx)
    oopts = oracle.cov_opts(window_length=wl, min_mapq=10)
    lcv_all, gc_all = [], []
    for r, (d, seq, _, L), lo, hi in zip(regions, raw, t.offsets[:-1], t.offsets[1:]):
        counts, proc, skip = oracle.frag_gw(oracle_reads(oracle, d), oopts, L)
        cov, start, end, frm = oracle.frag_gw_stats(counts, oopts, L)
        gc, gap = oracle.refstats(seq, wl)
        lcv, _, ft = oracle.cov_records(cov, None, start, end, frm, False, oopts)
        assert hi - lo == counts.size
        assert_f32_equal_bits(t.rcv[lo:hi], cov, "rcv")
        assert np.array_equal(to_np(t.start[lo:hi]), start) and np.array_equal(to_np(t.end[lo:hi]), end)
        assert_f32_equal_bits(t.frm[lo:hi], frm, "frm")
        assert_f32_equal_bits(t.gc[lo:hi], gc, "gc")
        assert np.array_equal(to_np(t.gap[lo:hi]).astype(bool), gap.astype(bool))
        assert_f32_equal_bits(t.lcv[lo:hi], lcv, "lcv")
        assert np.array_equal(to_np(t.ft[lo:hi]), ft)
        lcv_all.append(lcv)
        gc_all.append(gc)
    assert t.num_processed == [oracle.frag_gw(oracle_reads(oracle, d), oopts, L)[1] for d, _, _, L in raw]
    pipeline.normalize_run(t, "MedianGcBinned", ctx=ctx)
    cv, cv2, _, flags, _ = oracle.norm_gc_median(np.concatenate(lcv_all), np.concatenate(gc_all))
    assert np.array_equal(to_np(t.norm_flags), flags)
    has = (flags & N.N2_CV).astype(bool)
    assert has.sum() > 100
    assert_f32_equal_bits(to_np(t.cv)[has], cv[has], "cv")


def test_region_loop_targets_and_coverage(oracle, ctx):
    regions, raw = _regions("cuda", with_targets=True)
    opts = cb.CoverageOptions(window_length=None, considered_regions="TargetRegions", min_raw_coverage=3)
    t = pipeline.coverage_run(regions, opts, ctx=ctx)
    oopts = oracle.cov_opts(window_length=0, min_raw_coverage=3)
    for (d, _, tg, L), lo, hi in zip(raw, t.offsets[:-1], t.offsets[1:]):
        counts, proc, skip = oracle.frag_targets(oracle_reads(oracle, d), oopts, tg[0], tg[1])
        assert np.array_equal(to_np(t.rcv[lo:hi]), counts.astype(np.float32))
        assert np.array_equal(to_np(t.ft[lo:hi]) & 1, (counts < 3).astype(np.uint8))
    assert t.gc is None and t.frm is None
    # Coverage kind: mean/sd per window
    opts = cb.CoverageOptions(window_length=1000, count_kind="Coverage")
    t = pipeline.coverage_run(regions, opts, ctx=ctx)
    for (d, _, _, L), lo, hi in zip(raw, t.offsets[:-1], t.offsets[1:]):
        mean, sd = oracle.coverage(oracle_reads(oracle, d), oracle.cov_opts(window_length=1000), L)[:2]
        assert_f32_equal_bits(t.rcv[lo:hi], mean, "mean")
        np.testing.assert_allclose(to_np(t.rcvsd[lo:hi]), sd, rtol=1e-6)


def test_region_loop_deferred_equals_synchronous(oracle, ctx):
    """cnvb_coverage_run_async + cnvb_coverage_wait: several runs enqueued without a host wait (normaliser chained
    in stream order) deliver the columns and tallies of the synchronous call; a reference panic surfaces at wait()."""
    import torch
    regions, raw = _regions("cuda")
    for kind_opts in (cb.CoverageOptions(window_length=500, min_mapq=10),
                      cb.CoverageOptions(window_length=1000, count_kind="Coverage")):
        ref = pipeline.coverage_run(regions, kind_opts, ctx=ctx)
        if kind_opts.count_kind == "Fragments":
            pipeline.normalize_run(ref, "MedianGcBinned", ctx=ctx)
        prep = pipeline.PreparedRegions(regions, kind_opts)
        tables = []
        for _ in range(4):
            t = pipeline.coverage_run(prep, kind_opts, ctx=ctx, defer=True)
            assert t.num_processed == [] and t._pending is not None
            if kind_opts.count_kind == "Fragments":
                pipeline.normalize_run(t, "MedianGcBinned", ctx=ctx, want_medians=False)
            tables.append(t)
        for t in tables:
            t.wait()
            assert t.num_processed == ref.num_processed and t.num_skipped == ref.num_skipped
            for col in ("start", "end", "rcv", "rcvsd", "lcv", "ft", "gc", "gap", "cv", "cv2", "norm_flags"):
                a, b = getattr(t, col), getattr(ref, col)
                assert (a is None) == (b is None), col
                if a is not None:
                    assert torch.equal(a.view(torch.uint8), b.view(torch.uint8)), col
            assert t.wait() is t       # idempotent
    # an unpaired read whose centre pos + end_pos lands outside the windows: the reference panics (agg.rs:144)
    d2 = simple_reads([45], [60])
    bad = [pipeline.Region("chrP", 50, soa_batch(d2, device="cuda"), None)]
    t = pipeline.coverage_run(bad, cb.CoverageOptions(window_length=100), ctx=ctx, defer=True)
    with pytest.raises(cb.ReferencePanic, match="region 0"):
        t.wait()
    with pytest.raises(cb.CnvbError, match="device-resident"):
        pipeline.coverage_run([pipeline.Region("chrH", 50, soa_batch(d2), None)], cb.CoverageOptions(window_length=100),
                              ctx=ctx, defer=True)


def test_region_loop_as_a_graph_equals_the_eager_loop(oracle, ctx):
    """cnvb_coverage_plan_*: the loop captured into a CUDA graph and replayed -- the same columns and tallies as the
    eager loop, also after the CONTENTS of the inputs changed between launches (addresses and lengths stay)."""
    import torch
    regions, raw = _regions("cuda")
    opts = cb.CoverageOptions(window_length=500, min_mapq=10)
    eager = pipeline.coverage_run(regions, opts, ctx=ctx)
    plan = pipeline.CoveragePlan(regions, opts, ctx=ctx)
    for _ in range(3):
        plan.launch()
    t = plan.wait()
    assert t.num_processed == eager.num_processed and t.num_skipped == eager.num_skipped
    bits = lambda x: x.view(torch.int32) if x.dtype == torch.float32 else x  # noqa: E731  (missing values are NaNs)
    for name in ("start", "end", "rcv", "lcv", "frm", "gc", "gap", "ft"):
        assert torch.equal(bits(getattr(t, name)), bits(getattr(eager, name))), name
    # other contents in the same buffers: every read of the first region fails the MAPQ filter now
    regions[0].reads.mapq.zero_()
    eager2 = pipeline.coverage_run(regions, opts, ctx=ctx)
    t2 = plan.launch() and plan.wait()
    assert t2.num_processed == eager2.num_processed and t2.num_processed[0] == 0
    assert torch.equal(bits(t2.rcv), bits(eager2.rcv)) and torch.equal(bits(t2.lcv), bits(eager2.lcv))
    plan.close()
    # eager calls keep working on the context after the plan is gone
    again = pipeline.coverage_run(regions, opts, ctx=ctx)
    assert torch.equal(bits(again.rcv), bits(eager2.rcv))


def test_region_loop_plan_refuses_what_it_cannot_capture(oracle, ctx):
    """A plan exists for genome-wide fragment counts without piles: the other kinds size their launches from the
    records, per-region host tables cannot be part of a captured loop.  The plan says so instead of building a graph
    that would go stale, and the context goes on working eagerly."""
    import torch
    regions, raw = _regions("cuda", with_targets=True)
    for opts in (cb.CoverageOptions(window_length=None, considered_regions="TargetRegions", min_raw_coverage=3),
                 cb.CoverageOptions(window_length=1000, count_kind="Coverage")):
        eager = pipeline.coverage_run(regions, opts, ctx=ctx)
        with pytest.raises(N.CnvbError, match="plan"):
            pipeline.CoveragePlan(regions, opts, ctx=ctx)
        again = pipeline.coverage_run(regions, opts, ctx=ctx)
        assert torch.equal(again.rcv.view(torch.int32), eager.rcv.view(torch.int32))
    # piles: host tables per region
    gw = cb.CoverageOptions(window_length=500, min_mapq=10)
    piled = [pipeline.Region(r.name, r.length, r.reads, r.sequence, piles=(np.array([100], np.uint32), np.array([200], np.uint32)))
             for r in regions]
    with pytest.raises(N.CnvbError, match="plan"):
        pipeline.CoveragePlan(piled, gw, ctx=ctx)
    assert pipeline.coverage_run(piled, gw, ctx=ctx).rcv.numel() > 0


def test_region_loop_host_columns_and_panic(oracle, ctx):
    """C ABI called directly with HOST columns; a region that makes the reference panic fails the run."""
    wl = 100
    d0 = simple_reads([10, 250, 900], [60, 300, 950])
    d0["isize"][:] = 40
    d0["flag"][:] = 0x3
    d1 = simple_reads([5], [40])           # unpaired: centre = pos + end_pos = 45 -> bin 0 (fine)
    L = [1000, 50]
    batches = [soa_batch(d0), soa_batch(d1)]
    arr = (N.Region * 2)()
    keep = []
    for i, b in enumerate(batches):
        soa, k = b.c_soa()
        keep.append(k)
        arr[i].region_end = L[i]
        arr[i].reads = soa
    o = cb.CoverageOptions(window_length=wl, min_mapq=0).c_opts()
    total = int(N.lib().cnvb_coverage_rows(C.byref(o), N.FRAGMENTS_GENOME_WIDE, arr, 2))
    assert total == 11
    cols = {k: np.zeros(total, np.float32) for k in ("rcv", "lcv", "frm")}
    cols.update({k: np.zeros(total, np.uint32) for k in ("start", "end")})
    ft = np.zeros(total, np.uint8)
    cc = N.CovColumns(cols["start"].ctypes.data, cols["end"].ctypes.data, cols["rcv"].ctypes.data, None,
                      cols["lcv"].ctypes.data, None, cols["frm"].ctypes.data, None, None, None, ft.ctypes.data)
    nproc = np.zeros(2, np.uint64)
    nskip = np.zeros(2, np.uint64)
    ctx.check(N.lib().cnvb_coverage_run(ctx.handle, C.byref(o), N.FRAGMENTS_GENOME_WIDE, arr, 2, C.byref(cc),
                                        nproc.ctypes.data, nskip.ctypes.data, N.MEM_HOST))
    oo = oracle.cov_opts(window_length=wl, min_mapq=0)
    exp = np.concatenate([oracle.frag_gw(oracle_reads(oracle, d), oo, l)[0] for d, l in zip((d0, d1), L)])
    assert np.array_equal(cols["rcv"], exp.astype(np.float32))
    assert cols["end"][-1] == 50 and cols["start"][10] == 0 and cols["end"][9] == 1000
    assert nproc.tolist() == [3, 1]
    # an unpaired read far right: centre = pos + end_pos lands outside -> reference panics (agg.rs:144)
    d2 = simple_reads([45], [60])          # centre 105 -> window 1 of 1
    soa, k = soa_batch(d2).c_soa()
    arr[1].reads = soa
    rc = N.lib().cnvb_coverage_run(ctx.handle, C.byref(o), N.FRAGMENTS_GENOME_WIDE, arr, 2, C.byref(cc),
                                   nproc.ctypes.data, nskip.ctypes.data, N.MEM_HOST)
    assert rc == N.ERR_REFERENCE_PANIC
    with pytest.raises(cb.ReferencePanic, match="region 1"):
        ctx.check(rc)


def test_quick_wis_build_model_chain(oracle, ctx):
    """quick wis-build-model (coverage on targets -> TotalCovSum -> merge-cov -> build-model-wis) for a
    small cohort, every stage against the oracle chain."""
    from cnvetti_b200 import model_wis as mw
    n_samples, lens = 12, (400_000, 250_000, 150_000)
    tgts = [synth.synth_targets(80 + i, L, 90) for i, L in enumerate(lens)]
    samples, ocols = [], []
    for s in range(n_samples):
        regions, lcvs = [], []
        for i, L in enumerate(lens):
            d = synth.synth_bam_records(1000 + 10 * s + i, L, 5000)
            regions.append(pipeline.Region("chr%d" % (i + 1), L, soa_batch(d, device="cuda"), None, targets=tgts[i]))
            counts, _, _ = oracle.frag_targets(oracle_reads(oracle, d), oracle.cov_opts(window_length=0, min_mapq=0),
                                               tgts[i][0], tgts[i][1])
            lcv, _, _ = oracle.cov_records(counts.astype(np.float32), None, tgts[i][0], tgts[i][1], None, True,
                                           oracle.cov_opts(window_length=0, min_mapq=0))
            lcvs.append(lcv)
        samples.append(regions)
        ocols.append(oracle.norm_total(np.concatenate(lcvs))[1])
    opts = mw.BuildModelWisOptions(max_ref_targets=20, min_ref_targets=3)
    model, X, tids = pipeline.quick_wis_build_model(samples, opts, ctx=ctx)
    oX = np.stack(ocols, axis=1)
    assert_f32_equal_bits(to_np(X), oX, "panel")
    otids = np.concatenate([np.full(len(t[0]), i, np.uint32) for i, t in enumerate(tgts)])
    assert np.array_equal(to_np(tids).astype(np.uint32), otids)
    odist, oidx = oracle.wis_distances(oX, otids, max_ref_targets=20)
    thr, pidx, plen = oracle.wis_prune(odist, oidx, opts.filter_z_score)
    assert np.array_equal(to_np(model.ref_len), plen)
    assert_f32_equal_bits(to_np(model.ref_dist), odist, "distances")
    calls = oracle.wis_calls(oX, pidx, plen, 3, opts.filter_z_score, opts.filter_rel)
    assert np.array_equal(to_np(model.num_calls), calls)


def test_wis_call_cohort_chain(oracle, ctx):
    """BASELINE config 5 in miniature (pipeline.wis_call_cohort): TotalCovSum per sample -> merge-cov ->
    build-model-wis (tensor engine) -> mod-coverage for the whole cohort -> HMM segmentation, every stage
    against the oracle chain; integer outputs and single-rounding f32 outputs bit-exact."""
    import torch
    from cnvetti_b200 import model_wis as mw, synth_torch
    T, S = 4500, 24
    lcv, tids = synth_torch.synth_cohort_lcv(5, T, 0, S, "cuda", n_contigs=5, cnv_per_sample=3)
    opts = mw.BuildModelWisOptions(max_ref_targets=30, min_ref_targets=5)
    timings = {}
    r = pipeline.wis_call_cohort(lcv, tids, wis_options=opts, sigma=0.08, ctx=ctx, timings=timings)
    assert set(timings) >= {"normalize", "merge_cov", "build_model_wis", "mod_coverage", "discover"}
    h = lcv.cpu().numpy()
    oX = np.stack([oracle.norm_total(np.ascontiguousarray(h[s]))[1] for s in range(S)], axis=1)
    assert_f32_equal_bits(to_np(r["panel"]), oX, "panel")
    otids = tids.cpu().numpy().astype(np.uint32)
    odist, oidx = oracle.wis_distances(oX, otids, max_ref_targets=30, nthreads=8)
    thr, pidx, plen = oracle.wis_prune(odist, oidx, opts.filter_z_score)
    m = r["model"]
    assert np.float32(m["threshold"]) == np.float32(thr)
    assert np.array_equal(to_np(m["ref_len"]), plen)
    assert_f32_equal_bits(to_np(m["ref_dist"]), odist, "distances")
    calls = oracle.wis_calls(oX, pidx, plen, 5, opts.filter_z_score, opts.filter_rel)
    filt = oracle.wis_filters(plen, calls, 5, opts.max_samples_reliable)
    assert np.array_equal(to_np(m["num_calls"]), calls) and np.array_equal(to_np(m["filt"]), filt)
    cuts = np.concatenate([[0], np.flatnonzero(np.diff(otids)) + 1, [T]])
    tab = oracle.hmm_tables(0.08)
    n_alt = 0
    for s in range(S):
        o = oracle.modcov(np.ascontiguousarray(oX[:, s]), filt, pidx, plen)
        assert_f32_equal_bits(to_np(r["cv_rel"])[s], o["cv_rel"], "cv_rel[%d]" % s)
        assert_f32_equal_bits(to_np(r["cvz_rows"])[s], o["cvz"], "cvz[%d]" % s)
        for a, b in zip(cuts[:-1], cuts[1:]):
            st = oracle.hmm_viterbi(np.ascontiguousarray(o["cv_rel"][a:b]), tab)
            assert np.array_equal(to_np(r["states"])[s, a:b], st), (s, a, b)
            n_alt += int((st != 2).sum())
    assert n_alt > 0


@pytest.mark.parametrize("n_regions,wl", [(3, 100), (3, 128), (40, 500), (520, 200)])
def test_region_loop_reference_statistics_batched_and_not(oracle, ctx, n_regions, wl):
    """GC / GAP columns of the region loop with device-resident sequences: one batched launch for all regions
    (window length >= 128, at most 512 regions) or region by region otherwise -- both against the oracle,
    with regions of odd lengths and misaligned sequence buffers."""
    import torch
    rng = np.random.default_rng(n_regions * 1000 + wl)
    big = torch.empty(n_regions * 70_000 + 64, dtype=torch.uint8, device="cuda")
    regions, seqs, at = [], [], 0
    for i in range(n_regions):
        L = int(rng.integers(wl + 1, 60_000))
        seq = synth.synth_reference(100 + i, L)
        at += int(rng.integers(0, 16))               # sequences start at arbitrary byte offsets
        view = big[at:at + L]
        view.copy_(torch.from_numpy(seq))
        at += L
        d = synth.synth_bam_records(200 + i, L, 2 if i == 1 else 50)
        regions.append(pipeline.Region("r%d" % i, L, soa_batch(d, device="cuda"), view))
        seqs.append(seq)
    t = pipeline.coverage_run(regions, cb.CoverageOptions(window_length=wl, min_mapq=0), ctx=ctx)
    for seq, lo, hi in zip(seqs, t.offsets[:-1], t.offsets[1:]):
        gc, gap = oracle.refstats(seq, wl)
        assert_f32_equal_bits(t.gc[lo:hi], gc, "gc")
        assert np.array_equal(to_np(t.gap[lo:hi]).astype(bool), gap.astype(bool))


# ---- the communicator and the sharded entry points ---------------------------------------------------------
def test_sharded_entry_points_world_1(ctx):
    """cnvb_comm with one rank needs no NCCL: the sharded normalisers and the sharded model equal the plain
    calls (the multi-rank runs are tools/check_sharded.py and tools/c_abi_sharded.c, below)."""
    import torch
    from cnvetti_b200 import model_wis as mw, parallel
    from cnvetti_b200.normalize import run_binned_correction, run_uncorrected_normalization
    comm = parallel.Communicator(ctx, 0, 1)
    rng = np.random.default_rng(5)
    lcv = torch.from_numpy((rng.poisson(120, 200_000) / 500.0).astype(np.float32)).cuda()
    gc = torch.from_numpy(rng.uniform(0.3, 0.6, 200_000).astype(np.float32)).cuda()
    a = run_binned_correction(lcv, gc, ctx=ctx)
    b = run_binned_correction(lcv, gc, comm=comm)
    assert torch.equal(a["cv"].view(torch.int32), b["cv"].view(torch.int32)) and np.array_equal(a["medians"].view(np.uint32), b["medians"].view(np.uint32))
    assert torch.equal(run_uncorrected_normalization(lcv, ctx=ctx)[0].view(torch.int32),
                       run_uncorrected_normalization(lcv, comm=comm)[0].view(torch.int32))
    X, tids = synth.synth_panel(9, 5000, 32, n_contigs=5)
    dX, dt = torch.from_numpy(X).cuda(), torch.from_numpy(tids.view(np.int32)).cuda()
    o = mw.BuildModelWisOptions(max_ref_targets=30)
    m, rows, T = mw.run_sharded(dX, dt, o, comm)
    f = mw.run(dX, dt, o, ctx=ctx)
    assert rows == (0, 5000) and T == 5000 and np.float32(m.threshold).tobytes() == np.float32(f.threshold).tobytes()
    assert torch.equal(m.ref_dist.view(torch.int32), f.ref_dist.view(torch.int32)) and torch.equal(m.num_calls, f.num_calls)
    assert torch.equal(m.filt, f.filt) and torch.equal(m.ref_len, f.ref_len)
    comm.close()


def test_merge_cov_kernel(ctx):
    """cnvb_merge_cov: columns side by side -> X[T][S] (lib-merge-cov/src/lib.rs:102-232), both input forms,
    ragged tile edges."""
    import torch
    from cnvetti_b200 import pipeline
    for T, S in ((1000, 7), (4097, 33), (130, 100), (1, 1)):
        cv = torch.randn(S, T, device="cuda")
        X = pipeline.merge_cov_panel(cv, ctx=ctx)
        assert torch.equal(X, cv.t().contiguous())
        tables = []
        for s in range(S):
            t = pipeline.CoverageTable()
            t.offsets, t.start, t.end = [0, T], torch.arange(T, device="cuda"), torch.arange(T, device="cuda") + 1
            t.cv = cv[s].clone()
            tables.append(t)
        X2, tids = pipeline.merge_cov(tables, ctx=ctx)
        assert torch.equal(X2, X) and int(tids.max()) == 0


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_c_abi_sharded_from_plain_c(tmp_path):
    """The multi-GPU entry points driven by a C program with one host thread per GPU (no Python, no torch, no
    CUDA calls of its own) -- what a Rust host binding the header does.  2 GPUs when the box has them
    (NCCL inside the library), else the one-rank form."""
    import subprocess
    exe = str(tmp_path / "c_abi_sharded")
    lib_dir = os.path.join(ROOT, "cnvetti_b200")
    subprocess.check_call(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "c_abi_sharded.c"),
                           "-o", exe, "-L", lib_dir, "-lcnvetti_b200", "-Wl,-rpath," + lib_dir, "-lpthread", "-lm"])
    world = 2 if _n_gpus() >= 2 else 1
    r = subprocess.run([exe, str(world)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c_abi_sharded ok" in r.stdout


def test_sharded_parity_under_torchrun():
    """tools/check_sharded.py on 2 GPUs: one genome sharded by contig, the sharded model, the cohort chain --
    all bit-identical to the single-GPU results."""
    import subprocess
    import sys
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs (run by gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "check_sharded.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ALL IDENTICAL" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
