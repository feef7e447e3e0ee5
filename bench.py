#!/usr/bin/env python3
"""bench.py -- throughput of the read-depth hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--scale F]

Workload (N=1 and every N): BASELINE.json configs[1] -- `cnvetti cmd coverage` (Fragments, 500 bp
windows, --reference for GC) + `cmd normalize --normalization MedianGcBinned` on a synthetic 30x
whole-genome sample (chr1-22 hg38 lengths, ~900 M reads).  One "step" = one pass of that path
over the whole sample.  Multi-GPU: one process per GPU, each rank processes its own sample (the
reference's only data parallelism is one coverage+normalize chain per input BAM,
cnvetti/src/quick_wis_build_model.rs:148-184), no data-path collective => weak scaling.

`value`  : reads/s, inputs already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same through the public API with HOST (pinned) SoA buffers: H2D of all inputs and
           D2H of the result table inside the timed region.
`roofline`: dominant kernel frag_bin_gw_kernel, algorithmic 7 B/read (mid i32 + mapq u8 + flags
           u16) over its CUDA-event time inside the timed steps.
`cpu_baseline`: the oracle (restated reference CPU path) on a bounded sample, host cores stated.
`--impl reference`: times that CPU path alone (the reference cannot be built here: no Rust).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WINDOW = 500
MIN_MAPQ = 20
READS_PER_BP = 900e6 / 2_875_001_522      # ~900 M reads over chr1-22
BYTES_PER_READ_GW = 7                     # mid i32 + mapq u8 + flags u16


HMM_KERNELS = ("hmm_viterbi_kernel", "hmm_chunk_matrix_kernel", "hmm_chunk_path_kernel", "hmm_boundary_kernel",
               "hmm_backtrack_kernel", "hmm_states_kernel")
HMM_BYTES_PER_BIN = 11  # SURVEY 8(d) cfg 4: 4 B CV + 5 B back-pointers + 1 B trace-back read + 1 B state


def host_buffer(shape, dtype, torch):
    """Pinned host memory where the box grants it (the e2e leg copies from it every step); pageable
    otherwise -- eight ranks pinning ~9 GB each can exceed what a node allows."""
    try:
        return torch.empty(shape, dtype=dtype, pin_memory=True), True
    except RuntimeError:
        return torch.empty(shape, dtype=dtype), False


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the full genome (tests)")
    ap.add_argument("--quick", action="store_true", help="device-resident leg only (profiling runs)")
    ap.add_argument("--workload", default="coverage", choices=["coverage", "wis", "hmm", "piles", "cfg1", "cfg5"],
                    help="coverage = BASELINE configs[1] (default, the headline); wis = configs[2]")
    ap.add_argument("--targets", type=int, default=None, help="wis / cfg5: number of targets T (200 000 / 3 000 000)")
    ap.add_argument("--samples", type=int, default=None, help="wis / cfg5: number of samples S (100 / 1000)")
    ap.add_argument("--hmm-samples", type=int, default=100, help="hmm: number of samples (22 lanes each)")
    ap.add_argument("--cpu-rows", type=int, default=2000, help="wis: rows of the CPU baseline sample")
    ap.add_argument("--cpu-sample-contigs", type=int, default=4, help="last k contigs form the CPU sample")
    return ap.parse_args()


def contigs(scale):
    from cnvetti_b200.synth import HG38_AUTOSOMES
    return [("chr%d" % (i + 1), max(100_000, int(L * scale))) for i, L in enumerate(HG38_AUTOSOMES)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md).  One
    long-running `nvidia-smi -lms 20` is started early (`arm`, before the warm-up: the tool needs
    ~100 ms to come up); `start`/`stop` bracket the timed region by wall clock and only the samples
    stamped inside it are reported.  A timed region shorter than the sampling period falls back to
    the samples taken under the same load since `arm`, and says so."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None
        self.arm()

    def arm(self):
        if self.proc:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def start(self):
        t_wait = time.time()
        while self.proc and not self.rows and time.time() - t_wait < 0.5:  # the tool is up once it has printed a line
            time.sleep(0.005)
        self.t0 = time.time()

    def _read(self):
        import datetime
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except Exception:
                ts = time.time()
            self.rows.append((ts, r))

    def stop(self):
        self.t1 = time.time()
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)  # let the sample that covers the end of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        inside = [r for ts, r in self.rows if self.t0 is not None and self.t0 - 0.01 <= ts <= self.t1 + 0.01]
        window = "timed region"
        if not inside:
            inside = [r for ts, r in self.rows if ts <= self.t1 + 0.01][-10:]
            window = "timed region shorter than the 20 ms sampling period: last samples under the same load (warm-up)"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic(n_reads, launches_per_step):
    """DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
    (profiles/roofline_traffic.json holds dram bytes per read of the captured launch)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p) and launches_per_step:
        try:
            per_read = float(json.load(open(p))["frag_bin_gw_kernel"]["dram_bytes_per_read"])
            return per_read * n_reads / launches_per_step
        except Exception:
            return None
    return None


def host_bam_decode_leg(n_pairs=1_000_000, contig_len=50_818_468):
    """Host BAM decode, timed on its own (north_star: reported separately from device time): a synthetic
    coordinate-sorted BAM of the config-1 read mix is written, then decoded by the library's BGZF/BAM
    decoder (zlib on all host threads) into pinned SoA columns."""
    import tempfile
    from cnvetti_b200 import bam, synth
    scale = n_pairs / 7_600_000.0
    L = max(100_000, int(contig_len * scale))
    d = synth.synth_bam_records(22, L, n_pairs)
    with tempfile.TemporaryDirectory() as tmp:
        path = synth.write_bam(os.path.join(tmp, "synth.bam"), [("chr22", L)], [d])
        size = os.path.getsize(path)
        best = None
        with bam.BamReader(path) as r:
            for _ in range(3):
                r.decode(min_unclipped=0.6, pinned=True)
                st = r.stats()
                if best is None or st["decode_seconds"] < best["decode_seconds"]:
                    best = st
    return {"reads": best["records"], "file_bytes": size, "uncompressed_bytes": best["uncompressed_bytes"],
            "seconds": best["decode_seconds"], "inflate_seconds": best["inflate_seconds"],
            "parse_seconds": best["parse_seconds"], "reads_per_s": best["records"] / best["decode_seconds"],
            "threads": os.cpu_count(), "what": "BGZF inflate (zlib, level-1 file) + record walk + SoA derivation into "
                                               "pinned columns; best of 3; not part of value / e2e"}


def file_level_leg(n_pairs=1_000_000, contig_len=50_818_468):
    """`cmd coverage` at file level (pipeline.cmd_coverage): synthetic BAM in, coverage VCF out, with the host
    decode, the device part and the VCF write timed apart (second run of two: buffers and pools are warm)."""
    import tempfile
    from cnvetti_b200 import pipeline, synth
    import cnvetti_b200 as cb
    scale = n_pairs / 7_600_000.0
    L = max(100_000, int(contig_len * scale))
    d = synth.synth_bam_records(22, L, n_pairs)
    opts = cb.CoverageOptions(window_length=1000, min_mapq=MIN_MAPQ)
    with tempfile.TemporaryDirectory() as tmp:
        path = synth.write_bam(os.path.join(tmp, "synth.bam"), [("22", L)], [d])
        out = os.path.join(tmp, "cov.vcf.gz")
        timings = None
        for _ in range(2):
            table, timings = pipeline.cmd_coverage(path, out, opts)
        return {"reads": 2 * n_pairs, "windows": table.offsets[-1], "bam_bytes": os.path.getsize(path),
                "vcf_bytes": os.path.getsize(out), "seconds": {k: round(v, 5) for k, v in timings.items()},
                "what": "BAM decode on host threads | region loop on the GPU from pinned columns (incl. H2D) | "
                        "VCF formatting + BGZF on host threads"}


# ---------------------------------------------------------------------------------------------
def cpu_reference_leg(sample, steps, warmup):
    """Times the oracle (restated reference CPU path) on the sample: coverage (Fragments,
    genome-wide) + refstats + record assembly per contig, then MedianGcBinned over the sample's
    windows.  Single thread: `cmd coverage` / `cmd normalize` are single-threaded in the
    reference (regions are processed sequentially, lib-coverage/src/lib.rs:768)."""
    import oracle
    oracle.build()
    oopts = oracle.cov_opts(window_length=WINDOW, min_mapq=MIN_MAPQ)
    prepared = [(oracle.ReadBatch(d["pos"], d["isize"], d["flag"], d["mapq"], d["cigar_off"], d["cigar"]), L, seq)
                for (d, L, seq) in sample]
    n_reads = sum(p[0].n for p in prepared)

    def one():
        lcvs, gcs = [], []
        for reads, L, seq in prepared:
            counts, _, _ = oracle.frag_gw(reads, oopts, L)
            cov, start, end, frm = oracle.frag_gw_stats(counts, oopts, L)
            lcv, _, _ = oracle.cov_records(cov, None, start, end, frm, False, oopts)
            gc, _ = oracle.refstats(seq, WINDOW)
            lcvs.append(lcv)
            gcs.append(gc)
        return oracle.norm_gc_median(np.concatenate(lcvs), np.concatenate(gcs))

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    return n_reads, dt


def bench_wis(args):
    """BASELINE configs[2]: build-model-wis on a synthetic panel (T targets x S samples, k=100).
    One step = compute_distances + prune + per-probe calls + filters (lib_model_wis::run)."""
    import torch
    import cnvetti_b200 as cb
    from cnvetti_b200 import model_wis as mw
    from cnvetti_b200 import synth
    T, S = args.targets or 200_000, args.samples or 100
    dev = "cuda:0"
    ctx = cb.Context(0)
    X, tids = synth.synth_panel(3, T, S)
    dX = torch.from_numpy(X).to(dev)
    dt = torch.from_numpy(tids.view(np.int32)).to(dev)
    opts = mw.BuildModelWisOptions(engine=mw.ENGINE_TENSOR)
    workload = "cfg3: build-model-wis, synthetic panel %d targets x %d samples, k=100 (tensor engine)" % (T, S)

    def step():
        return mw.run(dX, dt, opts, ctx=ctx)

    m = step()
    torch.cuda.synchronize()
    stats = mw.last_stats(ctx)
    sampler = ClockSampler(0)  # armed before the warm-up (nvidia-smi needs ~100 ms to come up)
    for _ in range(max(0, args.warmup - 1)):
        step()
    torch.cuda.synchronize()
    sampler.start()
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    launches = ctx.launch_count - l0
    # per-kernel times: a second, instrumented pass (timing mode serialises what the timed steps overlap)
    ctx.set_timing(True)
    ctx.timing_reset()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    kern = {}
    for kname in ("wis_sweep_kernel", "wis_compact_kernel", "wis_rescore_kernel", "wis_calls_kernel",
                  "wis_row_exact_kernel"):
        kms, cnt = ctx.timing_query(kname)
        kern[kname] = {"ms_per_step": kms / args.steps, "launches_per_step": cnt / args.steps}
    ctx.set_timing(False)
    ctx.timing_reset()
    value = float(T) * T / (ms * 1e-3)
    # e2e: host panel in, model out
    hX = torch.from_numpy(X).pin_memory()
    ht = torch.from_numpy(tids.view(np.int32)).pin_memory()
    k = min(T, opts.max_ref_targets)

    h_idx = torch.empty((T, k), dtype=torch.int32, pin_memory=True)
    h_len = torch.empty(T, dtype=torch.int32, pin_memory=True)
    h_calls = torch.empty(T, dtype=torch.int32, pin_memory=True)
    h_filt = torch.empty(T, dtype=torch.uint8, pin_memory=True)

    def step_e2e():
        mm = mw.run(hX.to(dev, non_blocking=True), ht.to(dev, non_blocking=True), opts, ctx=ctx)
        h_idx.copy_(mm.ref_idx, non_blocking=True)
        h_len.copy_(mm.ref_len, non_blocking=True)
        h_calls.copy_(mm.num_calls, non_blocking=True)
        h_filt.copy_(mm.filt, non_blocking=True)
        torch.cuda.synchronize()

    step_e2e()
    t0 = time.perf_counter()
    esteps = max(1, min(args.steps, 3))
    for _ in range(esteps):
        step_e2e()
    e2e_s = (time.perf_counter() - t0) / esteps
    peak_t = 1406.9
    pfile = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pfile):
        peak_t = float(json.load(open(pfile)).get("bf16_tflops_sustained", peak_t))
    sweep = kern["wis_sweep_kernel"]["ms_per_step"]
    achieved = 2.0 * T * T * S / (sweep * 1e-3) / 1e12 if sweep > 0 else 0.0
    # what the tensor cores executed: visited 256x128 tiles x padded K (projection windows skip the rest)
    tiles = float(stats["tiles"]) if stats else 0.0
    kp = ((S + 63) // 64) * 64
    executed = 2.0 * tiles * 256 * 128 * kp / (sweep * 1e-3) / 1e12 if sweep > 0 else 0.0
    tile_frac = tiles * 256 * 128 / (float(T) * T) if T else 0.0
    cpu = None
    if not args.quick:
        import oracle
        oracle.build()
        rows = min(args.cpu_rows, T)
        ncpu = os.cpu_count() or 1
        t0 = time.perf_counter()
        oracle.wis_distances(X, tids, 100, 0, rows, nthreads=ncpu)
        dt_cpu = time.perf_counter() - t0
        cpu = {"value": rows * float(T) / dt_cpu, "unit": "bin-pairs/s", "cores": ncpu, "kind": "port",
               "sample": "compute_distances for the first %d of %d target rows (all columns), %d threads as rayon "
                         "does; linear in rows" % (rows, T, ncpu)}
    line = {"metric": "WIS bin-pairs/s", "value": value, "unit": "bin-pairs/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (fp16 tensor filter + exact f32 re-score)", "data": "synthetic",
            "config": {"workload": workload, "targets": T, "samples": S, "k": k,
                       "l2": "panel %.0f MB + candidate buffers %.1f GB exceed L2" % (T * S * 4 / 1e6, T * 2048 * 8 / 1e9)},
            "e2e": {"value": float(T) * T / e2e_s, "unit": "bin-pairs/s", "h2d_bytes_per_step": int(X.nbytes + tids.nbytes),
                    "d2h_bytes_per_step": int(T * k * 4 + T * 9), "ms_per_step": e2e_s * 1e3},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "wis_sweep_kernel", "achieved": achieved, "peak": peak_t,
                         "unit": "TFLOP/s", "frac": achieved / peak_t, "traffic": None,
                         "algorithmic_flop_per_pair": 2 * S, "kernel_share_of_step": sweep / ms,
                         "executed_tflops": executed, "tile_fraction_visited": tile_frac,
                         "note": "achieved = 2*T*T*S / sweep time (SURVEY 8(d): the algorithmic count does not "
                                 "change with the method); the sweep visits only tile_fraction_visited of the "
                                 "pair matrix (projection windows), executed_tflops is what the tensor pipe did",
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained"},
            "cpu_baseline": cpu, "clocks": clocks, "kernels": kern, "wis_stats": stats}
    print(json.dumps(line))
    return 0


def bench_hmm(args):
    """BASELINE configs[3]: `cmd discover` on 100 samples x ~2.9 M bins (22 autosomes, 1 kbp windows):
    one lane per (sample, contig).  One step = Viterbi segmentation of all lanes (state per bin)."""
    import torch
    import cnvetti_b200 as cb
    from cnvetti_b200 import discover
    from cnvetti_b200.synth import HG38_AUTOSOMES
    n_samples = args.hmm_samples
    dev = "cuda:0"
    ctx = cb.Context(0)
    lens = [(L + 999) // 1000 for L in HG38_AUTOSOMES]
    lane_len = np.array(lens * n_samples, np.int64)
    lane_off = np.concatenate([[0], np.cumsum(lane_len)]).astype(np.uint64)
    n_bins = int(lane_off[-1])
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    cn = torch.full((n_bins,), 2.0, device=dev)
    # planted segments: ~20 per lane, 5..500 bins, cn in {0,1,3,4}
    nseg = 20 * len(lane_len)
    seg_start = (torch.rand(nseg, generator=g, device=dev) * (n_bins - 600)).long()
    seg_len = (torch.rand(nseg, generator=g, device=dev) * 495).long() + 5
    seg_cn = torch.tensor([0.0, 1.0, 3.0, 4.0], device=dev)[(torch.rand(nseg, generator=g, device=dev) * 4).long().clamp_(0, 3)]
    delta = torch.zeros(n_bins + 1, device=dev)
    delta.index_add_(0, seg_start, seg_cn - 2.0)
    delta.index_add_(0, seg_start + seg_len, -(seg_cn - 2.0))
    cn = (cn + torch.cumsum(delta[:-1], 0)).clamp_(0.0, 4.0).round_()
    mu = cn / 2.0
    cv = (mu + torch.randn(n_bins, generator=g, device=dev) * (0.08 * mu.sqrt() + 0.02)).float().contiguous()
    del delta, cn, mu
    sigma = np.full(len(lane_len), 0.08)
    opts = discover.DiscoverOptions()
    workload = "cfg4: cmd discover (5-state HMM Viterbi), %d samples x %d bins = %d lanes" % (
        n_samples, n_bins // n_samples, len(lane_len))

    def step():
        return discover.hmm_segment(cv, lane_off, sigma, opts, ctx=ctx)

    st = step()
    torch.cuda.synchronize()
    frac_cn2 = float((st == 2).float().mean().item())
    sampler = ClockSampler(0)  # armed before the warm-up (nvidia-smi needs ~100 ms to come up)
    for _ in range(max(0, args.warmup - 1)):
        step()
    torch.cuda.synchronize()
    sampler.start()
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    launches = ctx.launch_count - l0
    # per-kernel times: a second, instrumented pass (timing mode serialises what the timed steps overlap)
    ctx.set_timing(True)
    ctx.timing_reset()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    kern = {}
    for kname in HMM_KERNELS:
        kms, cnt = ctx.timing_query(kname)
        if cnt:
            kern[kname] = {"ms_per_step": kms / args.steps, "launches_per_step": cnt / args.steps}
    ctx.set_timing(False)
    ctx.timing_reset()
    # e2e: pinned host CV in, states out
    hcv = torch.empty(n_bins, dtype=torch.float32, pin_memory=True)
    hcv.copy_(cv)
    hst = torch.empty(n_bins, dtype=torch.uint8, pin_memory=True)

    def step_e2e():
        s_ = discover.hmm_segment(hcv.to(dev, non_blocking=True), lane_off, sigma, opts, ctx=ctx)
        hst.copy_(s_, non_blocking=True)
        torch.cuda.synchronize()

    step_e2e()
    esteps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(esteps):
        step_e2e()
    e2e_s = (time.perf_counter() - t0) / esteps
    peak, peak_src = measured_peak()
    dom = max(kern, key=lambda k_: kern[k_]["ms_per_step"]) if kern else None
    dom_ms = kern[dom]["ms_per_step"] if dom else 0.0
    achieved = HMM_BYTES_PER_BIN * n_bins / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    cpu = None
    if not args.quick:
        import oracle
        from concurrent.futures import ThreadPoolExecutor
        oracle.build()
        ncpu = os.cpu_count() or 1
        nl = min(len(lens), 22)
        host = cv[:int(lane_off[nl])].cpu().numpy()
        tab = oracle.hmm_tables(0.08, opts.sigma0, opts.eps_cn, opts.p_move)
        lanes = [host[int(lane_off[i]):int(lane_off[i + 1])] for i in range(nl)]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(ncpu) as ex:
            list(ex.map(lambda x: oracle.hmm_viterbi(x, tab), lanes))
        dt_cpu = time.perf_counter() - t0
        cpu = {"value": host.size / dt_cpu, "unit": "bins/s", "cores": min(ncpu, nl), "kind": "port",
               "sample": "one sample (22 lanes, %d bins), oracle Viterbi, one lane per thread" % host.size}
    line = {"metric": "HMM bins/s", "value": n_bins / (ms * 1e-3), "unit": "bins/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fixed-point i32 scores from f64 emissions", "data": "synthetic",
            "config": {"workload": workload, "samples": n_samples, "bins": n_bins, "lanes": len(lane_len),
                       "frac_cn2": frac_cn2,
                       "l2": "CV %.1f GB/step exceeds L2" % (n_bins * 4 / 1e9)},
            "e2e": {"value": n_bins / e2e_s, "unit": "bins/s", "h2d_bytes_per_step": n_bins * 4,
                    "d2h_bytes_per_step": n_bins, "ms_per_step": e2e_s * 1e3},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_bin": HMM_BYTES_PER_BIN,
                         "kernel_share_of_step": dom_ms / ms if ms else None},
            "cpu_baseline": cpu, "clocks": clocks, "kernels": kern}
    print(json.dumps(line))
    return 0


def bench_piles(args):
    """`--mask-piles` pass of cmd coverage (lib-coverage/src/piles.rs) on BASELINE configs[0]'s input:
    chr22, ~15 M reads.  One step = collect_piles of the contig (pile list in HBM)."""
    import torch
    import cnvetti_b200 as cb
    from cnvetti_b200 import piles as pl
    from cnvetti_b200 import synth_torch
    from cnvetti_b200.synth import CHR22_LEN
    dev = "cuda:0"
    ctx = cb.Context(0)
    L = int(CHR22_LEN * args.scale)
    n_pairs = int(7_600_000 * args.scale)
    soa = synth_torch.synth_soa_contig(22, L, n_pairs, dev)
    batch = cb.ReadBatch(pos=soa["pos"].contiguous(), end=soa["end"].contiguous(), mapq=soa["mapq"].contiguous(),
                         flags=soa["flags"].contiguous())
    n_reads = len(batch)
    Lc = int(soa["end"].max().item()) + 1
    col = pl.PileCollector(20, MIN_MAPQ, ctx=ctx)
    got = col.collect_piles(batch, Lc)
    sampler = ClockSampler(0)  # armed before the warm-up (nvidia-smi needs ~100 ms to come up)
    for _ in range(max(0, args.warmup - 1)):
        col.collect_piles(batch, Lc)
    torch.cuda.synchronize()
    sampler.start()
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        col.collect_piles(batch, Lc)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    launches = ctx.launch_count - l0
    hb = cb.ReadBatch(**{k: getattr(batch, k).cpu().pin_memory() for k in ("pos", "end", "mapq", "flags")})
    col.collect_piles(hb, Lc)
    t0 = time.perf_counter()
    esteps = max(1, min(args.steps, 3))
    for _ in range(esteps):
        col.collect_piles(hb, Lc)
    e2e_s = (time.perf_counter() - t0) / esteps
    cpu = None
    if not args.quick:
        import oracle
        oracle.build()
        nsub = min(n_reads, 4_000_000)
        d = synth_torch.soa_to_oracle_records(soa, 0, nsub)
        Ls = int(d["pos"].max()) + 1000
        rb = oracle.ReadBatch(d["pos"], d["isize"], d["flag"], d["mapq"], d["cigar_off"], d["cigar"])
        t0 = time.perf_counter()
        depth = oracle.coverage(rb, oracle.cov_opts(window_length=1000, min_mapq=MIN_MAPQ), Ls, debug=True)[3]
        oracle.piles_from_depths(depth[:Ls], 20)
        dt_cpu = time.perf_counter() - t0
        cpu = {"value": nsub / dt_cpu, "unit": "reads/s", "cores": 1, "kind": "port",
               "sample": "first %d reads of the contig: per-base depth pass + base-by-base pile walk, single thread" % nsub}
    peak, peak_src = measured_peak()
    line = {"metric": "aligned reads/s (pile collection)", "value": n_reads / (ms * 1e-3), "unit": "reads/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "cfg1 input (chr22, %d reads): cmd coverage --mask-piles, pile collection" % n_reads,
                       "piles": len(got), "max_pile_size": int(got.size.max()) if len(got) else 0,
                       "l2": "read columns %.0f MB exceed L2" % (n_reads * 11 / 1e6)},
            "e2e": {"value": n_reads / e2e_s, "unit": "reads/s", "h2d_bytes_per_step": n_reads * 11,
                    "d2h_bytes_per_step": len(got) * 12, "ms_per_step": e2e_s * 1e3},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "cub sort/scan + pile kernels", "achieved": n_reads * 11 / (ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": n_reads * 11 / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_read": 11,
                         "note": "whole call (two host syncs for the counts included), not a single kernel"},
            "cpu_baseline": cpu, "clocks": clocks}
    print(json.dumps(line))
    return 0


def bench_cfg1(args):
    """BASELINE configs[0]: one contig (chr22, ~15 M reads), 1 kbp windows, MAPQ >= 20; the three
    aggregator kinds of lib-coverage/src/lib.rs:510-529.  `value` is count kind Coverage (per-base
    depth mean/SD per window, the heaviest kind); the two Fragments kinds are reported beside it."""
    import torch
    import cnvetti_b200 as cb
    from cnvetti_b200 import pipeline, synth, synth_torch
    from cnvetti_b200.synth import CHR22_LEN
    dev = "cuda:0"
    ctx = cb.Context(0)
    L = int(CHR22_LEN * args.scale)
    soa = synth_torch.synth_soa_contig(22, L, int(7_600_000 * args.scale), dev)
    Lc = int(soa["end"].max().item()) + 1000
    batch = cb.ReadBatch(**{k: soa[k].contiguous() for k in ("pos", "end", "mid", "frag_end", "mapq", "flags")})
    n_reads = len(batch)
    ts, te = synth.synth_targets(22, L, int(4000 * args.scale) + 10)
    kinds = {
        "coverage": (cb.CoverageOptions(window_length=1000, min_mapq=MIN_MAPQ, count_kind="Coverage"), None, 11),
        "fragments_gw": (cb.CoverageOptions(window_length=1000, min_mapq=MIN_MAPQ), None, 7),
        "fragments_targets": (cb.CoverageOptions(window_length=None, min_mapq=MIN_MAPQ, considered_regions="TargetRegions"),
                              (ts, te), 11),
    }
    res = {}
    clocks = None
    launches = 0
    for name, (opts, targets, bpr) in kinds.items():
        prep = pipeline.PreparedRegions([pipeline.Region("chr22", Lc, batch, None, targets=targets)], opts)
        for _ in range(max(1, args.warmup)):
            pipeline.coverage_run(prep, opts, ctx=ctx)
        sampler = ClockSampler(0) if name == "coverage" else None
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        l0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            pipeline.coverage_run(prep, opts, ctx=ctx)
        e1.record()
        torch.cuda.synchronize()
        if sampler:
            clocks = sampler.stop()
            launches = ctx.launch_count - l0
        ms = e0.elapsed_time(e1) / args.steps
        res[name] = {"ms_per_step": ms, "reads_per_s": n_reads / (ms * 1e-3), "algorithmic_gbs": n_reads * bpr / (ms * 1e-3) / 1e9}
    # end to end (Coverage kind): pinned host columns in, table out
    opts = kinds["coverage"][0]
    hb = cb.ReadBatch(**{k: getattr(batch, k).cpu().pin_memory() for k in ("pos", "end", "mapq", "flags")})
    hprep = pipeline.PreparedRegions([pipeline.Region("chr22", Lc, hb, None)], opts)

    def step_e2e():
        t = pipeline.coverage_run(hprep, opts, ctx=ctx)
        out = [t.rcv.cpu(), t.rcvsd.cpu(), t.lcv.cpu()]
        torch.cuda.synchronize()
        return out

    step_e2e()
    esteps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(esteps):
        out = step_e2e()
    e2e_s = (time.perf_counter() - t0) / esteps
    cpu = None
    if not args.quick:
        import oracle
        oracle.build()
        nsub = min(n_reads, 4_000_000)
        d = synth_torch.soa_to_oracle_records(soa, 0, nsub)
        Ls = int(d["pos"].max()) + 2000
        rb = oracle.ReadBatch(d["pos"], d["isize"], d["flag"], d["mapq"], d["cigar_off"], d["cigar"])
        t0 = time.perf_counter()
        oracle.coverage(rb, oracle.cov_opts(window_length=1000, min_mapq=MIN_MAPQ), Ls)
        dt_cpu = time.perf_counter() - t0
        cpu = {"value": nsub / dt_cpu, "unit": "reads/s", "cores": 1, "kind": "port",
               "sample": "first %d reads of the contig, CoverageAggregator restatement, single thread" % nsub}
    peak, peak_src = measured_peak()
    cov = res["coverage"]
    line = {"metric": "aligned reads/s binned (count kind Coverage)", "value": cov["reads_per_s"], "unit": "reads/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cov["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 depths, f32 mean/sd", "data": "synthetic",
            "config": {"workload": "cfg1: chr22, %d reads, 1 kbp windows, MAPQ>=%d; Coverage / Fragments GW / Fragments targets"
                                   % (n_reads, MIN_MAPQ), "targets": int(ts.size),
                       "l2": "read columns %.0f MB exceed L2" % (n_reads * 11 / 1e6)},
            "e2e": {"value": n_reads / e2e_s, "unit": "reads/s", "h2d_bytes_per_step": n_reads * 11,
                    "d2h_bytes_per_step": int(sum(o.numel() * 4 for o in out)), "ms_per_step": e2e_s * 1e3},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "coverage_run (depth_prepare + depth_tile + finalize)",
                         "achieved": cov["algorithmic_gbs"], "peak": peak, "unit": "GB/s", "frac": cov["algorithmic_gbs"] / peak,
                         "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_read": 11,
                         "note": "whole region call, not a single kernel"},
            "cpu_baseline": cpu, "clocks": clocks, "kinds": res,
            "host_bam_decode": host_bam_decode_leg() if not args.quick else None,
            "file_level": file_level_leg() if not args.quick else None}
    print(json.dumps(line))
    return 0


def bench_cfg5(args):
    """BASELINE configs[4]: end-to-end WIS germline calling of a synthetic cohort (T targets x S samples):
    TotalCovSum per sample -> merge-cov -> build-model-wis -> mod-coverage for every sample -> HMM.
    Strong scaling: the cohort is fixed; samples are sharded for normalise / segmentation, target rows for
    the model and mod-coverage (one NCCL all-gather of the panel, one all-to-all of the relative coverage)."""
    import torch
    import torch.distributed as dist
    import cnvetti_b200 as cb
    from cnvetti_b200 import model_wis as mw, parallel, pipeline, synth_torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    ctx = cb.Context(local)
    T, S = args.targets or 3_000_000, args.samples or 1000
    ranges = [parallel.shard_rows(S, r, world) for r in range(world)]
    a, b = ranges[rank]
    t_gen = time.perf_counter()
    lcv, tids = synth_torch.synth_cohort_lcv(5, T, a, b, dev)
    torch.cuda.synchronize()

    def log(msg):
        if rank == 0:
            sys.stderr.write("[cfg5 %7.1f s] %s\n" % (time.perf_counter() - t_gen, msg))
            sys.stderr.flush()

    log("cohort generated: %d targets x %d samples" % (T, b - a))
    opts = mw.BuildModelWisOptions(engine=mw.ENGINE_TENSOR)
    workload = ("cfg5: WIS germline calling, synthetic cohort %d targets x %d samples: TotalCovSum, merge-cov, "
                "build-model-wis (k=100, tensor engine), mod-coverage, HMM segmentation" % (T, S))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(timings=None):
        r = pipeline.wis_call_cohort(lcv, tids, sample_ranges=ranges, wis_options=opts, sigma=0.08, ctx=ctx,
                                     timings=timings)
        n_alt = int((r["states"] != 2).sum().item())  # the step's result, read back
        return r, n_alt

    first = {}
    r, n_alt = step(first)
    stats = mw.last_stats(ctx)
    log("first step: stages %s; engine %s; alt bins %d" % ({k_: round(v, 3) for k_, v in first.items()}, stats, n_alt))
    # spot check at full size: 64 rows of this rank's block against the exact f32 engine
    lo, hi = r["rows"]
    m = r["model"]
    pick = torch.linspace(0, hi - lo - 1, 64).long().tolist() if hi > lo else []
    exact_opts = mw.BuildModelWisOptions(engine=mw.ENGINE_EXACT)
    spot_ok = True
    for q in pick:
        d1, i1 = mw.compute_distances(r["panel"], tids, exact_opts, row_begin=lo + q, row_end=lo + q + 1, ctx=ctx)
        spot_ok &= bool(torch.equal(d1[0].view(torch.int32), m["ref_dist"][q].view(torch.int32)))
    del r, m
    log("spot check of %d rows against the exact engine: %s" % (len(pick), spot_ok))
    sampler = ClockSampler(local)
    for _ in range(max(0, args.warmup - 1)):
        step()
    barrier()
    sampler.start()
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = ctx.launch_count - l0
    # stage breakdown (host clock around each stage, one extra step) and the sweep's share
    stages = {}
    step(stages)
    ctx.set_timing(True)
    ctx.timing_reset()
    step()
    torch.cuda.synchronize()
    kern = {}
    for kname in ("wis_sweep_kernel", "wis_compact_kernel", "wis_rescore_kernel", "wis_calls_kernel",
                  "modcov_panel_kernel") + tuple(HMM_KERNELS):
        kms, cnt = ctx.timing_query(kname)
        if cnt:
            kern[kname] = {"ms_per_step": kms, "launches_per_step": cnt}
    ctx.set_timing(False)
    ctx.timing_reset()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peak_t = 1406.9
    pfile = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pfile):
        peak_t = float(json.load(open(pfile)).get("bf16_tflops_sustained", peak_t))
    sweep = kern.get("wis_sweep_kernel", {}).get("ms_per_step", 0.0)
    rows = hi - lo
    kp = ((S + 63) // 64) * 64
    tiles = float(stats["tiles"]) if stats else 0.0
    executed = 2.0 * tiles * 256 * 128 * kp / (sweep * 1e-3) / 1e12 if sweep > 0 else 0.0
    achieved = 2.0 * rows * T * S / (sweep * 1e-3) / 1e12 if sweep > 0 else 0.0
    cpu = None
    e2e = None
    if not args.quick and world == 1:
        # end to end from pinned host memory: LCV columns in, state paths out
        h_in = torch.empty(lcv.shape, dtype=torch.float32, pin_memory=True)
        h_in.copy_(lcv)
        h_out = torch.empty((S, T), dtype=torch.uint8, pin_memory=True)
        del lcv
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d_in = h_in.to(dev, non_blocking=True)
        rr = pipeline.wis_call_cohort(d_in, tids, sample_ranges=ranges, wis_options=opts, sigma=0.08, ctx=ctx)
        h_out.copy_(rr["states"], non_blocking=True)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e = {"value": float(T) * T / e2e_s, "unit": "bin-pairs/s", "h2d_bytes_per_step": int(h_in.numel() * 4),
               "d2h_bytes_per_step": int(h_out.numel()), "ms_per_step": e2e_s * 1e3, "steps": 1}
        # CPU baseline: a bounded sample of the stage that dominates the CPU run
        import oracle
        oracle.build()
        ncpu = os.cpu_count() or 1
        nrows, cols = min(args.cpu_rows, 256), min(T, 200_000)
        pick_c = torch.arange(cols, device=dev) * (T // cols)  # spread over the genome (all contigs)
        Xs = np.ascontiguousarray(rr["panel"][pick_c].cpu().numpy())
        th = tids[pick_c].cpu().numpy().astype(np.uint32)
        del rr, d_in
        t0 = time.perf_counter()
        oracle.wis_distances(Xs, th, 100, 0, nrows, nthreads=ncpu)
        dt_cpu = time.perf_counter() - t0
        cpu = {"value": nrows * float(cols) / dt_cpu, "unit": "bin-pairs/s", "cores": ncpu, "kind": "port",
               "sample": "compute_distances (the stage that dominates the CPU run) for %d target rows x %d columns x %d "
                         "samples, %d threads as rayon does; linear in rows and columns" % (nrows, cols, S, ncpu)}
    line = {"metric": "WIS bin-pairs/s", "value": float(T) * T / (ms * 1e-3), "unit": "bin-pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 (fp16 tensor filter + exact f32 re-score), f64 statistics, i32 HMM scores",
            "data": "synthetic",
            "config": {"workload": workload, "targets": T, "samples": S, "k": 100, "bins_segmented": T * S,
                       "parallelism": "single GPU" if world == 1 else "samples/%d for normalise+HMM, target rows/%d for "
                                      "WIS+mod-coverage" % (world, world),
                       "l2": "panel %.1f GB exceeds L2" % (T * S * 4 / 1e9)},
            "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "wis_sweep_kernel", "achieved": achieved, "peak": peak_t,
                         "unit": "TFLOP/s", "frac": achieved / peak_t, "traffic": None,
                         "executed_tflops": executed, "executed_frac": executed / peak_t,
                         "tile_fraction_visited": tiles * 256 * 128 / (float(rows) * T) if rows else None,
                         "kernel_share_of_step": sweep / ms if ms else None,
                         "note": "achieved = 2*rows*T*S / sweep time (rank 0's row block); executed_tflops = visited "
                                 "256x128 tiles x padded K: what the tensor pipe did",
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained"},
            "cpu_baseline": cpu, "clocks": clocks, "kernels": kern,
            "stages_s": {k_: round(v, 4) for k_, v in stages.items()}, "wis_stats": stats,
            "checks": {"alt_bins": n_alt, "spot_rows_vs_exact_engine": len(pick), "spot_identical": spot_ok}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0 if spot_ok else 1


def main():
    args = parse_args()
    if args.workload == "cfg5" and args.impl == "b200":
        return bench_cfg5(args)
    if args.workload == "cfg1" and args.impl == "b200":
        return bench_cfg1(args)
    if args.workload == "piles" and args.impl == "b200":
        return bench_piles(args)
    if args.workload == "hmm" and args.impl == "b200":
        return bench_hmm(args)
    if args.workload == "wis" and args.impl == "b200":
        return bench_wis(args)
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    if args.impl == "reference" and rank != 0:
        return 0  # rank 0 alone runs the CPU arm

    if not torch.cuda.is_available():
        if args.impl == "reference":
            dev = None
        else:
            print(json.dumps({"error": "no CUDA device: the B200 path has no CPU fallback"}))
            return 2
    else:
        torch.cuda.set_device(local_rank)
        dev = "cuda:%d" % local_rank
    if distributed and args.impl == "b200":
        dist.init_process_group("nccl", device_id=torch.device(dev))

    import cnvetti_b200 as cb
    from cnvetti_b200 import pipeline, synth, synth_torch

    regs = contigs(args.scale)
    n_pairs = [max(1000, int(round(L * READS_PER_BP / 2))) for _, L in regs]
    workload = ("cfg2: cmd coverage (Fragments, %d bp windows, MAPQ>=%d, GC from reference) + normalize "
                "MedianGcBinned, synthetic 30x WGS chr1-22" % (WINDOW, MIN_MAPQ))

    # ---- reference arm: CPU only --------------------------------------------------------------
    k = max(1, min(args.cpu_sample_contigs, len(regs)))
    sample_ids = list(range(len(regs) - k, len(regs)))

    def make_cpu_sample():
        out = []
        for ci in sample_ids:
            L = regs[ci][1]
            gdev = dev or "cpu"
            soa = synth_torch.synth_soa_contig(2 * 1000 + ci + 7919 * rank, L, n_pairs[ci], gdev)
            d = synth_torch.soa_to_oracle_records(soa)
            seq = synth_torch.synth_reference_torch(2 * 1000 + ci, L, gdev).cpu().numpy()
            out.append((d, L, seq))
            del soa
        return out

    if args.impl == "reference":
        sample = make_cpu_sample()
        n_reads, dt = cpu_reference_leg(sample, max(1, args.steps), max(1, min(args.warmup, 1)))
        val = n_reads / dt
        desc = "%s (%.0f M reads) of the workload" % ("+".join(regs[i][0] for i in sample_ids), n_reads / 1e6)
        line = {
            "impl": "reference", "metric": "aligned reads/s binned", "value": val, "unit": "reads/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": {"workload": workload, "sample": desc},
            "cpu_baseline": {"value": val, "unit": "reads/s", "cores": 1, "kind": "port", "sample": desc,
                             "host_cores_available": os.cpu_count()},
            "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "restated reference CPU path (oracle/, plain C -O2); the Rust reference cannot be built "
                    "here; cmd coverage/normalize are single-threaded in the reference",
        }
        print(json.dumps(line))
        return 0

    # ---- B200 arm ---------------------------------------------------------------------------------
    ctx = cb.Context(local_rank)
    opts = cb.CoverageOptions(window_length=WINDOW, min_mapq=MIN_MAPQ)
    t_gen = time.perf_counter()
    dev_regions, host_cols = [], []
    n_reads = 0
    for ci, (name, L) in enumerate(regs):
        soa = synth_torch.synth_soa_contig(2 * 1000 + ci + 7919 * rank, L, n_pairs[ci], dev)
        seq = synth_torch.synth_reference_torch(2 * 1000 + ci, L, dev)
        batch = cb.ReadBatch(mid=soa["mid"].contiguous(), mapq=soa["mapq"].contiguous(), flags=soa["flags"].contiguous())
        dev_regions.append(pipeline.Region(name, L, batch, seq))
        n_reads += len(batch)
        del soa
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t_gen

    prep_dev = pipeline.PreparedRegions(dev_regions, opts)

    def step_device():
        t = pipeline.coverage_run(prep_dev, opts, ctx=ctx)
        return pipeline.normalize_run(t, "MedianGcBinned", ctx=ctx)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # correctness guard on the full-size run: every counted fragment lands in exactly one window
    t = step_device()
    torch.cuda.synchronize()
    assert int(t.rcv.double().sum().item()) == sum(t.num_processed) - sum(t.num_skipped), "checksum"
    n_windows = int(t.rcv.numel())
    sampler = ClockSampler(local_rank)  # armed before the warm-up (nvidia-smi needs ~100 ms to come up)
    for _ in range(max(0, args.warmup - 1)):
        step_device()
    barrier()
    sampler.start()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    ms = e0.elapsed_time(e1) / args.steps
    # per-kernel device time: a second, instrumented pass (events around every launch cost a few
    # microseconds each, so they are kept out of the timed region above).  In timing mode the library
    # keeps all regions on one stream: in the timed steps the regions' kernels overlap on side streams
    # (memory-bound binning beside ALU-bound reference statistics), where a single kernel's duration
    # says nothing about the kernel.
    ctx.set_timing(True)
    ctx.timing_reset()
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i0.record()
    for _ in range(args.steps):
        step_device()
    i1.record()
    torch.cuda.synchronize()
    ms_serial = i0.elapsed_time(i1) / args.steps
    kern = {}
    for kname in ("frag_bin_gw_kernel", "refstats_lut_kernel", "gc_radix_hist_kernel"):
        kms, cnt = ctx.timing_query(kname)
        kern[kname] = {"ms_per_step": kms / args.steps, "launches_per_step": cnt / args.steps}
    ctx.set_timing(False)
    ctx.timing_reset()
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_max = float(tmax.item())
    value = world * n_reads / (ms_max * 1e-3)

    if args.quick:
        if rank == 0:
            print(json.dumps({"metric": "aligned reads/s binned", "value": value, "unit": "reads/s",
                              "n_gpus": world, "ms_per_step": ms_max, "gpu_launches": launches,
                              "ms_per_step_serialised": ms_serial, "kernels": kern, "quick": True}))
        if distributed:
            dist.destroy_process_group()
        return 0

    # ---- end to end: pinned host SoA in, table out -----------------------------------------------------
    host_regions = []
    h2d = 0
    all_pinned = True
    for r in dev_regions:
        cols = {}
        for kcol in ("mid", "mapq", "flags"):
            src = getattr(r.reads, kcol)
            h, pinned = host_buffer(src.shape, src.dtype, torch)
            all_pinned = all_pinned and pinned
            h.copy_(src)
            cols[kcol] = h
            h2d += h.numel() * h.element_size()
        hseq, pinned = host_buffer(r.sequence.shape, torch.uint8, torch)
        all_pinned = all_pinned and pinned
        hseq.copy_(r.sequence)
        h2d += hseq.numel()
        host_regions.append(pipeline.Region(r.name, r.length, cb.ReadBatch(**cols), hseq))
    torch.cuda.synchronize()
    out_host = {kcol: torch.empty(n_windows, dtype=dt, pin_memory=True)
                for kcol, dt in (("cv", torch.float32), ("cv2", torch.float32), ("lcv", torch.float32),
                                 ("rcv", torch.float32), ("gc", torch.float32), ("norm_flags", torch.uint8),
                                 ("gap", torch.uint8))}
    d2h = sum(v.numel() * v.element_size() for v in out_host.values())

    prep_host = pipeline.PreparedRegions(host_regions, opts)

    def step_e2e():
        tt = pipeline.coverage_run(prep_host, opts, ctx=ctx)
        pipeline.normalize_run(tt, "MedianGcBinned", ctx=ctx)
        for kcol, h in out_host.items():
            h.copy_(getattr(tt, kcol), non_blocking=True)
        torch.cuda.synchronize()

    e2e_steps = max(1, min(args.steps, 5))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n_reads / float(te.item())

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    peak, peak_src = measured_peak()
    step_bytes = float(BYTES_PER_READ_GW * n_reads + sum(L for _, L in regs) + n_windows * (5 + 4 + 16 + 8))
    gw = kern["frag_bin_gw_kernel"]
    achieved = BYTES_PER_READ_GW * n_reads / (gw["ms_per_step"] * 1e-3) / 1e9 if gw["ms_per_step"] > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "frag_bin_gw_kernel", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(n_reads, gw["launches_per_step"]), "peak_source": peak_src,
                "algorithmic_bytes_per_read": BYTES_PER_READ_GW,
                "algorithmic_bytes_per_launch": BYTES_PER_READ_GW * n_reads / max(1.0, gw["launches_per_step"]),
                "kernel_share_of_step": gw["ms_per_step"] / ms_serial,
                "kernel_timing": "instrumented pass with the regions serialised on one stream (%.3f ms/step); the "
                                 "timed steps overlap regions on three side streams" % ms_serial,
                "step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms * 1e-3) / 1e9, "unit": "GB/s",
                         "frac": step_bytes / (ms * 1e-3) / 1e9 / peak,
                         "what": "all kernels of one step together: 7 B/read + 1 B/base + 5 B/window refstats out "
                                 "+ 4 B/window counts + normalise 8 B/window x 2 passes + 8 B/window out"}}
    rs = kern["refstats_lut_kernel"]
    total_bp = sum(L for _, L in regs)
    roofline["step"]["algorithmic_bytes"] = step_bytes
    kern["refstats_lut_kernel"]["achieved_gbs"] = total_bp / (rs["ms_per_step"] * 1e-3) / 1e9 if rs["ms_per_step"] > 0 else 0.0

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ------------------------------------------
    cpu = None
    if world == 1:
        sample = make_cpu_sample()
        cn, cdt = cpu_reference_leg(sample, 1, 1)
        cpu = {"value": cn / cdt, "unit": "reads/s", "cores": 1, "kind": "port",
               "sample": "%s (%.0f M reads) of the workload, oracle coverage+refstats+normalize, single thread "
                         "(the reference's cmd coverage/normalize are single-threaded)" %
                         ("+".join(regs[i][0] for i in sample_ids), cn / 1e6),
               "host_cores_available": os.cpu_count()}

    line = {
        "metric": "aligned reads/s binned", "value": value, "unit": "reads/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload, "reads_per_gpu": n_reads, "windows_per_gpu": n_windows,
                   "contigs": len(regs), "window_length": WINDOW, "scale": args.scale,
                   "l2": "inputs (%.1f GB/step) far larger than the 126 MB L2; no flush needed" %
                         ((BYTES_PER_READ_GW * n_reads + total_bp) / 1e9),
                   "parallelism": "one sample per GPU, no collective" if world > 1 else "single GPU"},
        "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": float(te.item()) * 1e3, "steps": e2e_steps,
                "host_memory": "pinned" if all_pinned else "pageable (pinning refused)"},
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "kernels": kern, "gen_seconds": gen_s,
        "host_bam_decode": host_bam_decode_leg() if (rank == 0 and not args.quick) else None,
    }
    print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    # stdout carries exactly one JSON line.  Native libraries write to file descriptor 1 behind Python's
    # back (NCCL prints its version banner there), so fd 1 is pointed at stderr for the run and the
    # real stdout is kept aside for the result line(s) that main() prints.
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_real_stdout, "w", buffering=1)
    rc = main()
    sys.stdout.flush()
    sys.exit(rc)
