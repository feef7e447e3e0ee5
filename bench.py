#!/usr/bin/env python3
"""bench.py -- throughput of the read-depth hot path on B200 (BASELINE.json metric: aligned reads/s
binned; WIS bin-pairs/s; HMM bins/s at 1/2/4/8 B200 vs host CPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ...]

ONE JSON line.  Headline = BASELINE.json configs[1]: `cnvetti cmd coverage` (Fragments, 500 bp windows,
--reference for GC) + `cmd normalize --normalization MedianGcBinned` on a synthetic 30x whole-genome
sample (chr1-22 hg38 lengths, ~900 M reads); one "step" = one pass of that path over the whole sample.
  N = 1   the sample on one GPU.
  N > 1   the SAME sample sharded by contig over the ranks (longest-processing-time assignment,
          lib-coverage/src/lib.rs:768-781: regions are independent), the normaliser's statistics
          exchanged inside the library (4 all-reduces of the 101 x 256 radix histograms, NCCL on the
          context's stream) => strong scaling; `replicas` (one sample per GPU, no collective: the
          reference's own data parallelism, quick_wis_build_model.rs:148-184) is reported beside it.
`value`   reads/s, inputs already resident in HBM (CUDA events, max over ranks).
`e2e`     the same through the public API with HOST (pinned) SoA buffers: H2D of all inputs and D2H
          of the result table inside the timed region.
`roofline` dominant kernel frag_bin_gw_kernel, algorithmic 7 B/read over its
          CUDA-event time.
`cpu_baseline` the oracle (restated reference CPU path) on a bounded sample, host cores stated.
`checks`  chr19-22 of the workload through the GPU path against the oracle (counts, GC bits, medians,
          CV bits): a mismatch fails the run.
`workloads` the other BASELINE metrics in the same run, each a full line of its own (roofline,
          cpu_baseline, e2e, clocks): cfg1 (chr22, three aggregator kinds), cfg3 (build-model-wis
          200 000 x 100), cfg4 (HMM, 100 samples x 2.9 M bins), cfg5 (cohort 3 M x 1000, whole chain).
          `--workload NAME` runs one of them alone (profiling).
`--impl reference` times the restated reference CPU path alone (the Rust reference cannot be built
          here: no cargo / rustc, un-vendored dependencies).
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
               "hmm_backtrack_kernel", "hmm_states_kernel", "hmm_emit_kernel")
HMM_BYTES_PER_BIN = 11  # SURVEY 8(d) cfg 4: 4 B CV + 5 B back-pointers + 1 B trace-back read + 1 B state
WIS_KERNELS = ("wis_sweep_kernel", "wis_compact_kernel", "wis_rescore_kernel", "wis_calls_kernel",
               "wis_row_exact_kernel", "wis_panel_check_kernel", "modcov_panel_kernel", "merge_cov_kernel")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the full genome (tests)")
    ap.add_argument("--quick", action="store_true", help="device-resident legs only (profiling runs)")
    ap.add_argument("--workload", default="all",
                    choices=["all", "coverage", "wis", "hmm", "piles", "cfg1", "cfg5"],
                    help="all = the headline (coverage, BASELINE configs[1]) with the other configs under "
                         "`workloads`; a name = that workload alone")
    ap.add_argument("--targets", type=int, default=None, help="wis / cfg5: number of targets T (200 000 / 3 000 000)")
    ap.add_argument("--samples", type=int, default=None, help="wis / cfg5: number of samples S (100 / 1000)")
    ap.add_argument("--hmm-samples", type=int, default=100, help="hmm: number of samples (22 lanes each)")
    ap.add_argument("--cpu-rows", type=int, default=2000, help="wis: rows of the CPU baseline sample")
    ap.add_argument("--cpu-sample-contigs", type=int, default=4, help="last k contigs form the CPU sample")
    ap.add_argument("--no-replicas", action="store_true", help="N > 1: skip the one-sample-per-GPU side measurement")
    ap.add_argument("--no-graph", action="store_true",
                    help="default workload: eager region loop in the timed steps instead of the CUDA-graph plan")
    return ap.parse_args()


# ---- plumbing -------------------------------------------------------------------------------------------
class Env:
    """rank / device / context / communicator of this process"""

    def __init__(self, args):
        import torch
        import cnvetti_b200 as cb
        from cnvetti_b200 import parallel
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = "cuda:%d" % self.local
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device(self.dev))
            self.dist = dist
        self.ctx = cb.Context(self.local)
        ClockSampler.enabled = self.rank == 0
        self.comm = parallel.Communicator(self.ctx, self.rank, self.world)
        self.t0 = time.perf_counter()

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def log(self, msg):
        if self.rank == 0:
            sys.stderr.write("[bench %7.1f s] %s\n" % (time.perf_counter() - self.t0, msg))
            sys.stderr.flush()

    def free(self):
        """between workloads: hand cached device memory back"""
        import gc
        gc.collect()
        self.ctx.release_memory()
        self.torch.cuda.empty_cache()

    def close(self):
        self.comm.close()
        if self.dist:
            self.dist.destroy_process_group()


def host_buffer(shape, dtype, torch):
    """Pinned host memory where the box grants it (the e2e leg copies from it every step); pageable
    otherwise."""
    try:
        return torch.empty(shape, dtype=dtype, pin_memory=True), True
    except RuntimeError:
        return torch.empty(shape, dtype=dtype), False


def contigs(scale):
    from cnvetti_b200.synth import HG38_AUTOSOMES
    return [("chr%d" % (i + 1), max(100_000, int(L * scale))) for i, L in enumerate(HG38_AUTOSOMES)]


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md's clocks line), sampled
    in-process through NVML every 5 ms: two light queries per sample (SM clock, clocks-event reasons).
    A background `nvidia-smi -lms 20` -- the first version of this sampler -- stalls NCCL collectives on the
    sampled GPU: at 8 ranks the sharded normaliser went from 0.18 to 0.79 ms per step with it
    (gpurun_out/r2_scale_probe_n8.json), so only rank 0 samples, and only these two counters.  `start`/`stop`
    bracket the timed region by wall clock; a region shorter than the period falls back to the samples
    taken under the same load since `arm` (the warm-up), and says so.  Falls back to one `nvidia-smi`
    query after the region when NVML is not importable."""
    enabled = True   # Env switches it off on ranks other than 0
    PERIOD = 0.005

    def __init__(self, index):
        self.index, self.rows, self.t0, self.t1 = index, [], None, None
        self.nvml, self.handle, self.max_mhz, self.thread, self.stop_flag = None, None, None, None, False
        if ClockSampler.enabled:
            self.arm()

    def arm(self):
        if self.thread:
            return
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    why = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    why = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append((time.time(), float(mhz), int(why)))
            except Exception:
                pass
            time.sleep(self.PERIOD)

    def start(self):
        t_wait = time.time()
        while self.thread and not self.rows and time.time() - t_wait < 0.2:
            time.sleep(0.001)
        self.t0 = time.time()

    def stop(self):
        self.t1 = time.time()
        if not ClockSampler.enabled:
            return None
        if not self.thread:
            return self._smi_once()
        time.sleep(2 * self.PERIOD)  # let the sample that covers the end of the region arrive
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        nv = self.nvml
        inside = [r for r in self.rows if self.t0 is not None and self.t0 - 0.002 <= r[0] <= self.t1 + 0.002]
        window = "timed region"
        if not inside:
            inside = [r for r in self.rows if r[0] <= self.t1 + 0.002][-10:]
            window = "timed region shorter than the 5 ms sampling period: last samples under the same load (warm-up)"
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(n for n, b in bits.items() if any(r[2] & b for r in inside))
        sm = [r[1] for r in inside]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                "window": window, "reasons": reasons, "source": "NVML in-process, rank 0, 5 ms period"}

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(r[0]), "sm_max_mhz": float(r[1]), "samples": 1,
                    "window": "one nvidia-smi query right after the timed region (NVML not importable)",
                    "reasons": sorted(n for n, v in zip(names, r[2:6]) if v.strip().lower().startswith("active"))}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def measured_tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p)).get("bf16_tflops_sustained", 1406.9)), "MEASURED_PEAKS.json bf16_tflops_sustained"
    return 1406.9, "fallback 1406.9 TFLOP/s sustained bf16 (B200_PROFILING.md)"


def ncu_traffic(kernel, units_per_launch):
    """DRAM bytes per launch of `kernel` from the COMMITTED ncu --set full capture (profiles/
    roofline_traffic.json holds dram bytes per unit of the captured launch) -- not measured in this run."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p) and units_per_launch:
        try:
            ent = json.load(open(p))[kernel]
            per = float(ent.get("dram_bytes_per_read", ent.get("dram_bytes_per_unit")))
            return per * units_per_launch, "committed ncu capture (profiles/roofline_traffic.json: %.3f DRAM B/unit x units per launch)" % per
        except Exception:
            return None, None
    return None, None


def timed_steps(env, step, steps, warmup, sampler=None):
    """W untimed steps, then K steps between barrier + synchronize on both sides, CUDA events on the
    context's stream (= torch's current stream), max over ranks.  Returns (ms per step, launches, clocks)."""
    torch = env.torch
    for _ in range(max(0, warmup)):
        step()
    env.barrier()
    if sampler:
        sampler.start()
    l0 = env.ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    if hasattr(step, "flush"):
        step.flush()   # (deferred steps: the last tallies are collected inside the timed region)
    e1.record()
    env.barrier()
    clocks = sampler.stop() if sampler else None
    ms = env.max_over_ranks(e0.elapsed_time(e1) / steps)
    return ms, env.ctx.launch_count - l0, clocks


def kernel_times(env, step, steps, names):
    """per-kernel device time: a second, instrumented pass (events around every launch cost a few
    microseconds each, so they stay out of the timed region; in timing mode the library keeps the
    regions on one stream -- a kernel's duration means something only when it has the device to itself)"""
    ctx, torch = env.ctx, env.torch
    ctx.set_timing(True)
    ctx.timing_reset()
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i0.record()
    for _ in range(steps):
        step()
    if hasattr(step, "flush"):
        step.flush()
    i1.record()
    torch.cuda.synchronize()
    kern = {}
    for kname in names:
        kms, cnt = ctx.timing_query(kname)
        if cnt:
            kern[kname] = {"ms_per_step": kms / steps, "launches_per_step": cnt / steps}
    ctx.set_timing(False)
    ctx.timing_reset()
    return kern, i0.elapsed_time(i1) / steps


def host_bam_decode_leg(n_pairs=1_000_000, contig_len=50_818_468, level=6, payload="filler"):
    """Host BAM decode, timed on its own (north_star: reported separately from device time): a synthetic
    coordinate-sorted BAM of the config-1 read mix is written at zlib level `level` (htslib's default is 6),
    then decoded by the library's BGZF/BAM decoder (csrc/hostz.cpp's DEFLATE decoder and CRC-32 on all host threads, zlib
    behind them) into pinned SoA columns.  payload="filler": sequences and qualities are zeros / 0xFF and deflate 15 : 1
    (the decoder's best case: long matches); payload="random": random bases and qualities from eight bins, 2.2 : 1 --
    what a sequencer's file looks like to the inflater (short matches and literals), the number to plan with."""
    import tempfile
    from cnvetti_b200 import bam, synth
    scale = n_pairs / 7_600_000.0
    L = max(100_000, int(contig_len * scale))
    d = synth.synth_bam_records(22, L, n_pairs)
    with tempfile.TemporaryDirectory() as tmp:
        try:
            path = synth.write_bam(os.path.join(tmp, "synth.bam"), [("chr22", L)], [d], level=level, payload=payload)
        except TypeError:
            path = synth.write_bam(os.path.join(tmp, "synth.bam"), [("chr22", L)], [d])
            level = 1
        size = os.path.getsize(path)
        best = None
        with bam.BamReader(path) as r:
            for _ in range(3):
                r.decode(min_unclipped=0.6, pinned=True)
                st = r.stats()
                if best is None or st["decode_seconds"] < best["decode_seconds"]:
                    best = st
    ncpu = os.cpu_count() or 1
    return {"reads": best["records"], "file_bytes": size, "uncompressed_bytes": best["uncompressed_bytes"],
            "zlib_level": level, "payload": payload, "seconds": best["decode_seconds"], "inflate_seconds": best["inflate_seconds"],
            "parse_seconds": best["parse_seconds"], "reads_per_s": best["records"] / best["decode_seconds"],
            "compressed_mb_per_s": size / best["decode_seconds"] / 1e6,
            "compressed_mb_per_s_per_thread": size / best["decode_seconds"] / 1e6 / ncpu,
            "threads": ncpu, "what": "BGZF inflate (own DEFLATE decoder + CLMUL CRC-32, zlib as fallback) + record walk on all threads + SoA "
                                     "derivation into pinned columns; "
                                     "best of 3; not part of value / e2e"}


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
        path = synth.write_bam(os.path.join(tmp, "synth.bam"), [("22", L)], [d], level=6)  # (htslib's default level)
        out = os.path.join(tmp, "cov.vcf.gz")
        timings = None
        for _ in range(2):
            table, timings = pipeline.cmd_coverage(path, out, opts)
        return {"reads": 2 * n_pairs, "windows": table.offsets[-1], "bam_bytes": os.path.getsize(path),
                "vcf_bytes": os.path.getsize(out), "seconds": {k: round(v, 5) for k, v in timings.items()},
                "what": "BAM decode on host threads | region loop on the GPU from pinned columns (incl. H2D) | "
                        "VCF formatting + BGZF on host threads"}


# ---- CPU legs (the only places that execute oracle/) -----------------------------------------------------
def oracle_cfg2(sample):
    """The oracle over the sample's contigs: coverage (Fragments, genome-wide) + refstats + record assembly
    per contig, then MedianGcBinned over the sample's windows.  Returns every column (for the parity
    check) -- single thread: `cmd coverage` / `cmd normalize` are single-threaded in the reference
    (regions are processed sequentially, lib-coverage/src/lib.rs:768)."""
    import oracle
    oopts = oracle.cov_opts(window_length=WINDOW, min_mapq=MIN_MAPQ)
    lcvs, gcs, gaps, rcvs, procs = [], [], [], [], []
    for reads, L, seq in sample:
        counts, proc, _ = oracle.frag_gw(reads, oopts, L)
        cov, start, end, frm = oracle.frag_gw_stats(counts, oopts, L)
        lcv, _, _ = oracle.cov_records(cov, None, start, end, frm, False, oopts)
        gc, gap = oracle.refstats(seq, WINDOW)
        lcvs.append(lcv), gcs.append(gc), gaps.append(gap), rcvs.append(cov), procs.append(proc)
    lcv, gc = np.concatenate(lcvs), np.concatenate(gcs)
    cv, cv2, _, flags, med = oracle.norm_gc_median(lcv, gc)
    return dict(rcv=np.concatenate(rcvs), lcv=lcv, gc=gc, gap=np.concatenate(gaps), cv=cv, cv2=cv2, flags=flags,
                medians=med, processed=procs)


def cpu_reference_leg(sample, steps, warmup):
    import oracle
    oracle.build()
    prepared = [(oracle.ReadBatch(d["pos"], d["isize"], d["flag"], d["mapq"], d["cigar_off"], d["cigar"]), L, seq)
                for (d, L, seq) in sample]
    n_reads = sum(p[0].n for p in prepared)
    out = None
    for _ in range(warmup):
        out = oracle_cfg2(prepared)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = oracle_cfg2(prepared)
    dt = (time.perf_counter() - t0) / steps
    return n_reads, dt, out


def oracle_rows(X, tids, rows, k):
    """compute_distances of single rows, one oracle call per row on a thread each (ctypes drops the GIL)"""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        return list(ex.map(lambda r: oracle.wis_distances(X, tids, k, int(r), int(r) + 1), rows))


def check_rows_against_oracle(X_host, tids_host, rows, got_dist, got_idx, k=100):
    """reference sets, their order and the distance bits of `rows` against the oracle -> (ok, n_rows)"""
    ok = True
    for q, (od, oi) in enumerate(oracle_rows(X_host, tids_host, rows, k)):
        ok &= bool(np.array_equal(got_idx[q].astype(np.uint32), oi[0]))
        ok &= bool(np.array_equal(np.ascontiguousarray(got_dist[q]).view(np.uint32), od[0].view(np.uint32)))
    return ok, len(rows)


# ---- workloads ------------------------------------------------------------------------------------------------
def wis_k_issued(S):
    """K extent of the tcgen05.mma instructions one tile takes: 16 per K16 slab that holds data -- up to 125 samples the
    three fold entries ride along (wis_resident.cuh), beyond that the samples alone (streamed sweep)."""
    return 16 * ((S + 3 + 15) // 16) if S + 3 <= 128 else 16 * ((S + 15) // 16)


def bench_wis(args, env, steps=None, warmup=None):
    """BASELINE configs[2]: build-model-wis on a synthetic panel (T targets x S samples, k=100).
    One step = compute_distances + prune + per-probe calls + filters (lib_model_wis::run).  N > 1: target
    rows sharded over the ranks (cnvb_wis_build_sharded: panel all-gather, nearest-distance all-gather)."""
    torch = env.torch
    from cnvetti_b200 import model_wis as mw
    from cnvetti_b200 import parallel, synth
    steps = steps or args.steps
    warmup = args.warmup if warmup is None else warmup
    T, S = args.targets or 200_000, args.samples or 100
    dev, ctx, world, rank = env.dev, env.ctx, env.world, env.rank
    X, tids = synth.synth_panel(3, T, S)
    lo, hi = parallel.shard_rows(T, rank, world)
    opts = mw.BuildModelWisOptions(engine=mw.ENGINE_TENSOR)
    workload = "cfg3: build-model-wis, synthetic panel %d targets x %d samples, k=100 (tensor engine)" % (T, S)
    if world == 1:
        dX = torch.from_numpy(X).to(dev)
        dt = torch.from_numpy(tids.view(np.int32)).to(dev)

        def step():
            return mw.run(dX, dt, opts, ctx=ctx)
    else:
        dX = torch.from_numpy(X[lo:hi]).to(dev)
        dt = torch.from_numpy(tids[lo:hi].view(np.int32)).to(dev)

        def step():
            return mw.run_sharded(dX, dt, opts, env.comm)[0]

    m = step()
    torch.cuda.synchronize()
    stats = mw.last_stats(ctx)
    # parity at full size: rows of this rank's block against the ORACLE (sets, order, distance bits)
    checks = None
    if not args.quick:
        import oracle
        oracle.build()
        rows = sorted(set(np.linspace(lo, hi - 1, 24).astype(np.int64).tolist()))
        gd = m.ref_dist[[r - lo for r in rows]].cpu().numpy()
        gi = m.ref_idx[[r - lo for r in rows]].cpu().numpy()
        ok, n = check_rows_against_oracle(X, tids, rows, gd, gi)
        checks = {"rows_vs_oracle": n, "identical": ok}
    sampler = ClockSampler(env.local)
    ms, launches, clocks = timed_steps(env, step, steps, max(0, warmup - 1), sampler)
    kern, _ = kernel_times(env, step, min(steps, 5), WIS_KERNELS)
    value = float(T) * T / (ms * 1e-3)
    k = min(T, opts.max_ref_targets)
    e2e = None
    if not args.quick and world == 1:
        hX = torch.from_numpy(X).pin_memory()
        ht = torch.from_numpy(tids.view(np.int32)).pin_memory()
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
        esteps = max(1, min(steps, 3))
        for _ in range(esteps):
            step_e2e()
        e2e_s = (time.perf_counter() - t0) / esteps
        e2e = {"value": float(T) * T / e2e_s, "unit": "bin-pairs/s", "h2d_bytes_per_step": int(X.nbytes + tids.nbytes),
               "d2h_bytes_per_step": int(T * k * 4 + T * 9), "ms_per_step": e2e_s * 1e3}
    peak_t, peak_src = measured_tensor_peak()
    sweep = kern.get("wis_sweep_kernel", {}).get("ms_per_step", 0.0)
    rows_n = hi - lo
    algorithmic = 2.0 * rows_n * T * S / (sweep * 1e-3) / 1e12 if sweep > 0 else 0.0
    # what the tensor cores executed: visited 256x128 tiles x padded K (projection windows skip the rest)
    tiles = float(stats["tiles"]) if stats else 0.0
    kp = wis_k_issued(S)
    executed = 2.0 * tiles * 256 * 128 * kp / (sweep * 1e-3) / 1e12 if sweep > 0 else 0.0
    tile_frac = tiles * 256 * 128 / (float(rows_n) * T) if T and rows_n else 0.0
    cpu = None
    if not args.quick and world == 1:
        import oracle
        rows_c = min(args.cpu_rows, T)
        ncpu = os.cpu_count() or 1
        t0 = time.perf_counter()
        oracle.wis_distances(X, tids, 100, 0, rows_c, nthreads=ncpu)
        dt_cpu = time.perf_counter() - t0
        cpu = {"value": rows_c * float(T) / dt_cpu, "unit": "bin-pairs/s", "cores": ncpu, "kind": "port",
               "sample": "compute_distances for the first %d of %d target rows (all columns), %d threads as rayon "
                         "does; linear in rows" % (rows_c, T, ncpu)}
    del dX, dt, m
    env.free()
    if rank != 0:
        return None
    return {"metric": "WIS bin-pairs/s", "value": value, "unit": "bin-pairs/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f32 (fp16 tensor filter + exact f32 re-score)", "data": "synthetic",
            "config": {"workload": workload, "targets": T, "samples": S, "k": k,
                       "parallelism": "single GPU" if world == 1 else "target rows / %d (cnvb_wis_build_sharded)" % world,
                       "l2": "panel %.0f MB + candidate buffers %.1f GB exceed L2" % (T * S * 4 / 1e6, T * 2048 * 8 / 1e9)},
            "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "wis_sweep_kernel", "achieved": executed, "peak": peak_t,
                         "unit": "TFLOP/s", "frac": executed / peak_t, "traffic": None,
                         "what": "EXECUTED flops: visited 256x128 tiles x K of the MMAs issued (%d: 16 per K16 slab that "
                                 "holds data; round 1 quoted 128, the padded K) x 2 / sweep time -- what the tensor pipe did" % kp,
                         "padded_k_equiv": executed * (((S + 63) // 64) * 64) / kp if kp else None,
                         "algorithmic_equiv": algorithmic, "algorithmic_flop_per_pair": 2 * S,
                         "algorithmic_note": "2*rows*T*S / sweep time (SURVEY 8(d)'s count for the all-pairs "
                                             "contraction); the projection windows skip most tiles, so this exceeds "
                                             "what was executed and is NOT a roofline fraction",
                         "tile_fraction_visited": tile_frac, "kernel_share_of_step": sweep / ms if ms else None,
                         "peak_source": peak_src},
            "cpu_baseline": cpu, "clocks": clocks, "kernels": kern, "wis_stats": stats, "checks": checks}


def bench_hmm(args, env, steps=None, warmup=None):
    """BASELINE configs[3]: `cmd discover` on 100 samples x ~2.9 M bins (22 autosomes, 1 kbp windows):
    one lane per (sample, contig).  One step = Viterbi segmentation of all lanes (state per bin).
    N > 1: the samples are sharded over the ranks (lanes are independent, no collective)."""
    torch = env.torch
    from cnvetti_b200 import discover, parallel
    from cnvetti_b200.synth import HG38_AUTOSOMES
    steps = steps or args.steps
    warmup = args.warmup if warmup is None else warmup
    n_total = args.hmm_samples
    dev, ctx, world, rank = env.dev, env.ctx, env.world, env.rank
    s_lo, s_hi = parallel.shard_rows(n_total, rank, world)
    n_samples = s_hi - s_lo
    lens = [(L + 999) // 1000 for L in HG38_AUTOSOMES]
    bins_per_sample = int(sum(lens))
    lane_len = np.array(lens * n_samples, np.int64)
    lane_off = np.concatenate([[0], np.cumsum(lane_len)]).astype(np.uint64)
    n_bins = int(lane_off[-1])
    g = torch.Generator(device=dev)
    g.manual_seed(4 + 7919 * rank)
    cn = torch.full((n_bins,), 2.0, device=dev)
    nseg = 20 * len(lane_len)  # planted segments: ~20 per lane, 5..500 bins, cn in {0,1,3,4}
    seg_start = (torch.rand(nseg, generator=g, device=dev) * max(1, n_bins - 600)).long()
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
    total_bins = bins_per_sample * n_total
    workload = "cfg4: cmd discover (5-state HMM Viterbi), %d samples x %d bins = %d lanes" % (
        n_total, bins_per_sample, 22 * n_total)

    def step():
        return discover.hmm_segment(cv, lane_off, sigma, opts, ctx=ctx)

    st = step()
    torch.cuda.synchronize()
    frac_cn2 = float((st == 2).float().mean().item())
    # parity at full size: whole lanes of the first sample against the oracle's sequential Viterbi
    checks = None
    if not args.quick and rank == 0:
        import oracle
        oracle.build()
        tab = oracle.hmm_tables(0.08, opts.sigma0, opts.eps_cn, opts.p_move)
        ok, nl, differ, total = True, 0, 0, 0
        for li in (0, 10, 21):
            a, b = int(lane_off[li]), int(lane_off[li + 1])
            lane_cv, lane_st = cv[a:b].cpu().numpy(), st[a:b].cpu().numpy()
            ok &= bool(np.array_equal(oracle.hmm_viterbi(lane_cv, tab), lane_st))
            # quality number: the integer-score spec against the same recurrence in log-space FP64 (north_star's wording)
            differ += int(np.count_nonzero(oracle.hmm_viterbi_f64(lane_cv, tab) != lane_st))
            total += b - a
            nl += 1
        checks = {"lanes_vs_oracle": nl, "identical": ok,
                  "fp64_path_disagreement": {"bins": differ, "of": total, "rate": differ / max(total, 1),
                                             "what": "state calls that differ from a sequential log-space FP64 Viterbi"}}
    sampler = ClockSampler(env.local)
    ms, launches, clocks = timed_steps(env, step, steps, max(0, warmup - 1), sampler)
    kern, _ = kernel_times(env, step, min(steps, 5), HMM_KERNELS)
    e2e = None
    if not args.quick and world == 1:
        hcv = torch.empty(n_bins, dtype=torch.float32, pin_memory=True)
        hcv.copy_(cv)
        hst = torch.empty(n_bins, dtype=torch.uint8, pin_memory=True)

        def step_e2e():
            s_ = discover.hmm_segment(hcv.to(dev, non_blocking=True), lane_off, sigma, opts, ctx=ctx)
            hst.copy_(s_, non_blocking=True)
            torch.cuda.synchronize()

        step_e2e()
        esteps = max(1, min(steps, 3))
        t0 = time.perf_counter()
        for _ in range(esteps):
            step_e2e()
        e2e_s = (time.perf_counter() - t0) / esteps
        e2e = {"value": n_bins / e2e_s, "unit": "bins/s", "h2d_bytes_per_step": n_bins * 4,
               "d2h_bytes_per_step": n_bins, "ms_per_step": e2e_s * 1e3}
    peak, peak_src = measured_peak()
    dom = max(kern, key=lambda k_: kern[k_]["ms_per_step"]) if kern else None
    dom_ms = kern[dom]["ms_per_step"] if dom else 0.0
    kernels_ms = sum(v["ms_per_step"] for v in kern.values())
    cpu = None
    if not args.quick and world == 1:
        import oracle
        from concurrent.futures import ThreadPoolExecutor
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
    del cv, st
    env.free()
    if rank != 0:
        return None
    step_gbs = HMM_BYTES_PER_BIN * n_bins / (ms * 1e-3) / 1e9
    return {"metric": "HMM bins/s", "value": total_bins / (ms * 1e-3), "unit": "bins/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "fixed-point i32 scores from f32 emissions", "data": "synthetic",
            "config": {"workload": workload, "samples": n_total, "bins": total_bins, "lanes": 22 * n_total,
                       "frac_cn2": frac_cn2, "parallelism": "single GPU" if world == 1 else "samples / %d, no collective" % world,
                       "l2": "CV %.1f GB/step exceeds L2" % (n_bins * 4 / 1e9)},
            "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "all HMM kernels of a step (%s dominant)" % dom,
                         "achieved": step_gbs, "peak": peak, "unit": "GB/s",
                         "frac": step_gbs / peak if peak else None, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_bin": HMM_BYTES_PER_BIN,
                         "what": "11 B/bin x bins of this rank / step time (the step IS the kernel chain; three "
                                 "lane groups overlap on side streams)",
                         "dominant_kernel_ms": dom_ms, "kernels_ms_serialised": kernels_ms,
                         "kernel_share_of_step": dom_ms / kernels_ms if kernels_ms else None},
            "cpu_baseline": cpu, "clocks": clocks, "kernels": kern, "checks": checks}


def bench_piles(args, env, steps=None, warmup=None):
    """`--mask-piles` pass of cmd coverage (lib-coverage/src/piles.rs) on BASELINE configs[0]'s input:
    chr22, ~15 M reads.  One step = collect_piles of the contig (pile list in HBM)."""
    torch = env.torch
    import cnvetti_b200 as cb
    from cnvetti_b200 import piles as pl
    from cnvetti_b200 import synth_torch
    from cnvetti_b200.synth import CHR22_LEN
    steps = steps or args.steps
    warmup = args.warmup if warmup is None else warmup
    dev, ctx = env.dev, env.ctx
    L = int(CHR22_LEN * args.scale)
    n_pairs = int(7_600_000 * args.scale)
    soa = synth_torch.synth_soa_contig(22, L, n_pairs, dev)
    batch = cb.ReadBatch(pos=soa["pos"].contiguous(), end=soa["end"].contiguous(), mapq=soa["mapq"].contiguous(),
                         flags=soa["flags"].contiguous())
    n_reads = len(batch)
    Lc = int(soa["end"].max().item()) + 1
    col = pl.PileCollector(20, MIN_MAPQ, ctx=ctx)
    got = col.collect_piles(batch, Lc)
    sampler = ClockSampler(env.local)
    ms, launches, clocks = timed_steps(env, lambda: col.collect_piles(batch, Lc), steps, max(0, warmup - 1), sampler)
    hb = cb.ReadBatch(**{k: getattr(batch, k).cpu().pin_memory() for k in ("pos", "end", "mapq", "flags")})
    col.collect_piles(hb, Lc)
    t0 = time.perf_counter()
    esteps = max(1, min(steps, 3))
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
    del soa, batch
    env.free()
    if env.rank != 0:
        return None
    return {"metric": "aligned reads/s (pile collection)", "value": n_reads / (ms * 1e-3), "unit": "reads/s", "n_gpus": 1,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
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


CFG1_KERNELS = ("frag_bin_gw_kernel", "frag_bin_tgt_kernel", "depth_tile_kernel", "depth_prepare_kernel",
                "depth_ranges_kernel", "window_finalize_kernel", "next_covered_kernel", "depth_region_kernel")


def bench_cfg1(args, env, steps=None, warmup=None):
    """BASELINE configs[0]: one contig (chr22, ~15 M reads), 1 kbp windows, MAPQ >= 20; the three
    aggregator kinds of lib-coverage/src/lib.rs:510-529.  `value` is count kind Coverage (per-base
    depth mean/SD per window, the heaviest kind); the two Fragments kinds are reported beside it, each
    with its own roofline fraction.  A single region does not shard: at N > 1 rank 0 alone runs it."""
    torch = env.torch
    import cnvetti_b200 as cb
    from cnvetti_b200 import pipeline, synth, synth_torch
    from cnvetti_b200.synth import CHR22_LEN
    steps = steps or args.steps
    warmup = args.warmup if warmup is None else warmup
    if env.rank != 0:
        return None
    dev, ctx = env.dev, env.ctx
    L = int(CHR22_LEN * args.scale)
    soa = synth_torch.synth_soa_contig(22, L, int(7_600_000 * args.scale), dev)
    Lc = int(soa["end"].max().item()) + 1000
    batch = cb.ReadBatch(**{k: soa[k].contiguous() for k in ("pos", "end", "mid", "frag_end", "mapq", "flags")})
    n_reads = len(batch)
    ts, te = synth.synth_targets(22, L, int(4000 * args.scale) + 10)
    n_windows = (Lc + 999) // 1000
    kinds = {
        # Coverage 11 B/read + 8 B/base (difference array write + read; it lives in shared memory
        # here, so only the 11 B/read + 8 B/window out reach HBM -- both figures are reported)
        "coverage": (cb.CoverageOptions(window_length=1000, min_mapq=MIN_MAPQ, count_kind="Coverage"), None, 11),
        "fragments_gw": (cb.CoverageOptions(window_length=1000, min_mapq=MIN_MAPQ), None, 7),
        "fragments_targets": (cb.CoverageOptions(window_length=None, min_mapq=MIN_MAPQ, considered_regions="TargetRegions"),
                              (ts, te), 11),
    }
    peak, peak_src = measured_peak()
    res = {}
    clocks = None
    launches = 0
    solo = Env.__new__(Env)  # this workload runs on rank 0 alone: no barrier / reduction over ranks
    solo.__dict__.update(env.__dict__)
    solo.dist = None
    for name, (opts, targets, bpr) in kinds.items():
        prep = pipeline.PreparedRegions([pipeline.Region("chr22", Lc, batch, None, targets=targets)], opts)
        sampler = ClockSampler(env.local) if name == "coverage" else None
        step = lambda: pipeline.coverage_run(prep, opts, ctx=ctx)  # noqa: E731
        ms, ln, ck = timed_steps(solo, step, steps, max(1, warmup), sampler)
        if sampler:
            clocks, launches = ck, ln
        kern, _ = kernel_times(solo, step, min(steps, 5), CFG1_KERNELS)
        # bytes per read of the kind (+ for count kind Coverage 8 B per base for the difference array
        # and 8 B per window out; the array lives in shared memory here, so `hbm_only` counts what reaches HBM)
        alg = n_reads * bpr + (8 * Lc + 8 * n_windows if name == "coverage" else 0)
        dom = max(kern, key=lambda k_: kern[k_]["ms_per_step"]) if kern else None
        dom_ms = kern[dom]["ms_per_step"] if dom else 0.0
        res[name] = {"ms_per_step": ms, "reads_per_s": n_reads / (ms * 1e-3), "launches_per_step": ln / steps,
                     "algorithmic_bytes_per_step": alg,
                     "roofline": {"bound": "hbm", "kernel": dom, "kernel_ms": dom_ms,
                                  "achieved": alg / (dom_ms * 1e-3) / 1e9 if dom_ms else None,
                                  "frac": alg / (dom_ms * 1e-3) / 1e9 / peak if dom_ms else None, "peak": peak,
                                  "hbm_only": {"bytes": n_reads * bpr,
                                               "frac": n_reads * bpr / (dom_ms * 1e-3) / 1e9 / peak if dom_ms else None},
                                  "kernel_share_of_step": dom_ms / ms if ms else None},
                     "whole_call": {"achieved": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / peak,
                                    "what": "every launch of the region call plus the host wait for the tallies"},
                     "kernels": kern}
    cov = res["coverage"]
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
    esteps = max(1, min(steps, 5))
    t0 = time.perf_counter()
    for _ in range(esteps):
        out = step_e2e()
    e2e_s = (time.perf_counter() - t0) / esteps
    cpu = None
    checks = None
    if not args.quick:
        # the whole contig through the oracle: CPU baseline AND parity at full size, all three kinds
        import oracle
        oracle.build()
        d = synth_torch.soa_to_oracle_records(soa)
        rb = oracle.ReadBatch(d["pos"], d["isize"], d["flag"], d["mapq"], d["cigar_off"], d["cigar"])
        oopts = oracle.cov_opts(window_length=1000, min_mapq=MIN_MAPQ)
        t0 = time.perf_counter()
        omean, osd = oracle.coverage(rb, oopts, Lc)
        dt_cov = time.perf_counter() - t0
        t0 = time.perf_counter()
        ocounts, oproc, oskip = oracle.frag_gw(rb, oopts, Lc)
        dt_gw = time.perf_counter() - t0
        t0 = time.perf_counter()
        otc, otp, ots = oracle.frag_targets(rb, oracle.cov_opts(window_length=0, min_mapq=MIN_MAPQ), ts, te)
        dt_tg = time.perf_counter() - t0
        cpu = {"value": n_reads / dt_cov, "unit": "reads/s", "cores": 1, "kind": "port",
               "sample": "the whole contig (%d reads), CoverageAggregator restatement, single thread" % n_reads,
               "fragments_gw_reads_per_s": n_reads / dt_gw, "fragments_targets_reads_per_s": n_reads / dt_tg}
        tables = {}
        for name, (o, targets, _) in kinds.items():
            tables[name] = pipeline.coverage_run([pipeline.Region("chr22", Lc, batch, None, targets=targets)], o, ctx=ctx)
        torch.cuda.synchronize()
        bits = lambda x: np.ascontiguousarray(x.cpu().numpy() if hasattr(x, "cpu") else x, np.float32).view(np.uint32)  # noqa: E731
        c = tables["coverage"]
        sd_g, fin = c.rcvsd.cpu().numpy(), np.isfinite(osd)
        ok_cov = bool(np.array_equal(bits(c.rcv), bits(omean))) and bool(
            np.allclose(sd_g[fin], osd[fin], rtol=1e-6, atol=0) and np.array_equal(np.isfinite(sd_g), fin))
        g = tables["fragments_gw"]
        ok_gw = bool(np.array_equal(g.rcv.cpu().numpy().astype(np.int64), ocounts.astype(np.int64))) and \
            g.num_processed == [oproc] and g.num_skipped == [oskip]
        t = tables["fragments_targets"]
        ok_tg = bool(np.array_equal(t.rcv.cpu().numpy().astype(np.int64), otc.astype(np.int64))) and \
            t.num_processed == [otp] and t.num_skipped == [ots]
        checks = {"what": "the whole contig against the oracle: window means bit-exact + SD rel 1e-6 (Coverage), "
                          "counts and tallies (both Fragments kinds)",
                  "coverage": ok_cov, "fragments_gw": ok_gw, "fragments_targets": ok_tg,
                  "identical": ok_cov and ok_gw and ok_tg}
    line = {"metric": "aligned reads/s binned (count kind Coverage)", "value": cov["reads_per_s"], "unit": "reads/s",
            "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": cov["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 depths, f32 mean/sd", "data": "synthetic",
            "config": {"workload": "cfg1: chr22, %d reads, 1 kbp windows, MAPQ>=%d; Coverage / Fragments GW / Fragments targets"
                                   % (n_reads, MIN_MAPQ), "targets": int(ts.size),
                       "parallelism": "one region: does not shard (rank 0 alone at N > 1)",
                       "l2": "read columns %.0f MB exceed L2" % (n_reads * 11 / 1e6)},
            "e2e": {"value": n_reads / e2e_s, "unit": "reads/s", "h2d_bytes_per_step": n_reads * 11,
                    "d2h_bytes_per_step": int(sum(o.numel() * 4 for o in out)), "ms_per_step": e2e_s * 1e3},
            "gpu_launches": launches,
            "roofline": dict(cov["roofline"], unit="GB/s", traffic=None, peak_source=peak_src,
                             algorithmic_bytes="SURVEY 8(d): 11 B/read + 8 B/base (difference array write + read) + "
                                               "8 B/window out = %d B per launch" % cov["algorithmic_bytes_per_step"],
                             whole_call=cov["whole_call"],
                             note="`kinds` holds the same for both Fragments kinds (7 and 11 B/read)"),
            "cpu_baseline": cpu, "clocks": clocks, "kinds": res, "checks": checks,
            "host_bam_decode": host_bam_decode_leg() if not args.quick else None,
            "file_level": file_level_leg() if not args.quick else None}
    del soa, batch
    env.free()
    return line


def bench_cfg5(args, env, steps=None, warmup=None):
    """BASELINE configs[4]: end-to-end WIS germline calling of a synthetic cohort (T targets x S samples):
    TotalCovSum per sample -> merge-cov -> build-model-wis -> mod-coverage for every sample -> HMM.
    Strong scaling: the cohort is fixed; samples are sharded for normalise / segmentation, target rows for
    the model and mod-coverage (one all-gather of the panel, one all-to-all of the relative coverage; NCCL
    inside the library, on the context's stream)."""
    torch = env.torch
    from cnvetti_b200 import model_wis as mw, parallel, pipeline, synth_torch
    steps = steps or args.steps
    warmup = args.warmup if warmup is None else warmup
    rank, world, dev, ctx = env.rank, env.world, env.dev, env.ctx
    T, S = args.targets or 3_000_000, args.samples or 1000
    ranges = [parallel.shard_rows(S, r, world) for r in range(world)]
    a, b = ranges[rank]
    lcv, tids = synth_torch.synth_cohort_lcv(5, T, a, b, dev)
    torch.cuda.synchronize()
    env.log("cfg5 cohort generated: %d targets x %d samples" % (T, b - a))
    opts = mw.BuildModelWisOptions.quick(engine=mw.ENGINE_TENSOR)
    workload = ("cfg5: WIS germline calling, synthetic cohort %d targets x %d samples: TotalCovSum, merge-cov, "
                "build-model-wis (k=100, tensor engine), mod-coverage, HMM segmentation" % (T, S))

    def step(timings=None):
        r = pipeline.wis_call_cohort(lcv, tids, sample_ranges=ranges, wis_options=opts, sigma=0.08, timings=timings,
                                     comm=env.comm)
        n_alt = int((r["states"] != 2).sum().item())  # the step's result, read back
        return r, n_alt

    first = {}
    r, n_alt = step(first)
    stats = mw.last_stats(ctx)
    env.log("cfg5 first step: stages %s; engine %s; alt bins %d" % ({k_: round(v, 3) for k_, v in first.items()}, stats, n_alt))
    lo, hi = r["rows"]
    m = r["model"]
    # parity at full size: rows of rank 0's block against the ORACLE (reference sets, order, distance bits)
    checks = {"alt_bins": n_alt}
    if not args.quick and rank == 0:
        import oracle
        oracle.build()
        n_rows = 32 if (os.cpu_count() or 1) >= 16 else 8
        rows = sorted(set(np.linspace(lo, hi - 1, n_rows).astype(np.int64).tolist()))
        gd = m["ref_dist"][[q - lo for q in rows]].cpu().numpy()
        gi = m["ref_idx"][[q - lo for q in rows]].cpu().numpy()
        try:
            t0 = time.perf_counter()
            Xh = r["panel"].cpu().numpy()
            th = tids.cpu().numpy().astype(np.uint32)
            ok, n = check_rows_against_oracle(Xh, th, rows, gd, gi)
            checks.update({"rows_vs_oracle": n, "identical": ok, "oracle_seconds": round(time.perf_counter() - t0, 1)})
            del Xh
        except MemoryError:
            checks.update({"rows_vs_oracle": 0, "identical": None, "note": "host memory too small for the panel"})
        env.log("cfg5 oracle rows: %s" % checks)
    del r, m
    sampler = ClockSampler(env.local)
    ms, launches, clocks = timed_steps(env, step, steps, max(0, warmup - 1), sampler)
    stages = {}
    step(stages)  # stage breakdown (host clock around each stage, one extra step)
    kern, _ = kernel_times(env, step, 1, WIS_KERNELS + HMM_KERNELS + ("sum_f32_dd_kernel", "norm_total_map_kernel"))
    peak_t, peak_src = measured_tensor_peak()
    sweep = kern.get("wis_sweep_kernel", {}).get("ms_per_step", 0.0)
    rows_n = hi - lo
    kp = wis_k_issued(S)
    tiles = float(stats["tiles"]) if stats else 0.0
    executed = 2.0 * tiles * 256 * 128 * kp / (sweep * 1e-3) / 1e12 if sweep > 0 else 0.0
    algorithmic = 2.0 * rows_n * T * S / (sweep * 1e-3) / 1e12 if sweep > 0 else 0.0
    cpu = None
    e2e = None
    if not args.quick and world == 1:
        # end to end from pinned host memory: LCV columns in, state paths out
        h_in, _ = host_buffer(lcv.shape, torch.float32, torch)
        h_in.copy_(lcv)
        h_out, _ = host_buffer((S, T), torch.uint8, torch)
        del lcv
        env.free()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d_in = h_in.to(dev, non_blocking=True)
        rr = pipeline.wis_call_cohort(d_in, tids, sample_ranges=ranges, wis_options=opts, sigma=0.08, comm=env.comm)
        h_out.copy_(rr["states"], non_blocking=True)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e = {"value": float(T) * T / e2e_s, "unit": "bin-pairs/s", "h2d_bytes_per_step": int(h_in.numel() * 4),
               "d2h_bytes_per_step": int(h_out.numel()), "ms_per_step": e2e_s * 1e3, "steps": 1}
        # CPU baseline: a bounded sample of the stage that dominates the CPU run
        import oracle
        ncpu = os.cpu_count() or 1
        nrows, cols = min(args.cpu_rows, 256), min(T, 200_000)
        pick_c = torch.arange(cols, device=dev) * (T // cols)  # spread over the genome (all contigs)
        Xs = np.ascontiguousarray(rr["panel"][pick_c].cpu().numpy())
        th = tids[pick_c].cpu().numpy().astype(np.uint32)
        del rr, d_in, h_in, h_out
        t0 = time.perf_counter()
        oracle.wis_distances(Xs, th, 100, 0, nrows, nthreads=ncpu)
        dt_cpu = time.perf_counter() - t0
        cpu = {"value": nrows * float(cols) / dt_cpu, "unit": "bin-pairs/s", "cores": ncpu, "kind": "port",
               "sample": "compute_distances (the stage that dominates the CPU run) for %d target rows x %d columns x %d "
                         "samples, %d threads as rayon does; linear in rows and columns" % (nrows, cols, S, ncpu)}
    else:
        del lcv
    env.free()
    if rank != 0:
        return None
    try:  # the calls + mod-coverage pass: k reference rows' values per (target, sample) once, five f32 + one u8 out
        pk = kern.get("modcov_panel_kernel")
        if pk and pk.get("ms_per_step") and rows_n:
            hbm, _ = measured_peak()
            pbytes = float(rows_n) * S * (100 * 4 + 4 + 21)
            pk["hbm"] = {"algorithmic_bytes": pbytes, "achieved": pbytes / (pk["ms_per_step"] * 1e-3) / 1e9, "unit": "GB/s",
                         "frac": pbytes / (pk["ms_per_step"] * 1e-3) / 1e9 / hbm, "peak": hbm,
                         "what": "rank 0's rows x samples x (100 reference values + the target's own, 4 B each, + 21 B out); "
                                 "bound by latency at 8 warps per SM and the f32 -> f64 conversions, not by HBM (DESIGN 4.9)"}
    except Exception as exc:  # (a side figure must not cost the line)
        env.log("panel kernel figure skipped: %r" % (exc,))
    return {"metric": "WIS bin-pairs/s", "value": float(T) * T / (ms * 1e-3), "unit": "bin-pairs/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 (fp16 tensor filter + exact f32 re-score), f64 statistics, i32 HMM scores",
            "data": "synthetic",
            "config": {"workload": workload, "targets": T, "samples": S, "k": 100, "bins_segmented": T * S,
                       "parallelism": "single GPU" if world == 1 else "samples/%d for normalise+HMM, target rows/%d for "
                                      "WIS+mod-coverage" % (world, world),
                       "l2": "panel %.1f GB exceeds L2" % (T * S * 4 / 1e9)},
            "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "wis_sweep_kernel", "achieved": executed, "peak": peak_t,
                         "unit": "TFLOP/s", "frac": executed / peak_t, "traffic": None,
                         "what": "EXECUTED flops (rank 0): visited 256x128 tiles x K of the MMAs issued (%d) x 2 / sweep time" % kp,
                         "algorithmic_equiv": algorithmic,
                         "tile_fraction_visited": tiles * 256 * 128 / (float(rows_n) * T) if rows_n else None,
                         "kernel_share_of_step": sweep / ms if ms else None, "peak_source": peak_src},
            "cpu_baseline": cpu, "clocks": clocks, "kernels": kern,
            "stages_s": {k_: round(v, 4) for k_, v in stages.items()}, "wis_stats": stats, "checks": checks}


def reference_arm(args):
    """--impl reference: the restated reference CPU path (oracle/) of the headline workload on a bounded
    sample, plus bounded samples of the other BASELINE metrics under `workloads`.  Rank 0 alone runs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from cnvetti_b200 import synth, synth_torch
    import oracle
    oracle.build()
    gdev = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0")) if torch.cuda.is_available() else "cpu"
    regs = contigs(args.scale)
    n_pairs = [max(1000, int(round(L * READS_PER_BP / 2))) for _, L in regs]
    k = max(1, min(args.cpu_sample_contigs, len(regs)))
    sample_ids = list(range(len(regs) - k, len(regs)))
    sample = []
    for ci in sample_ids:
        L = regs[ci][1]
        soa = synth_torch.synth_soa_contig(2 * 1000 + ci, L, n_pairs[ci], gdev)
        sample.append((synth_torch.soa_to_oracle_records(soa), L,
                       synth_torch.synth_reference_torch(2 * 1000 + ci, L, gdev).cpu().numpy()))
        del soa
    n_reads, dt, _ = cpu_reference_leg(sample, max(1, min(args.steps, 3)), 1)
    del sample
    val = n_reads / dt
    desc = "%s (%.0f M reads) of the workload" % ("+".join(regs[i][0] for i in sample_ids), n_reads / 1e6)
    ncpu = os.cpu_count() or 1
    workloads = {}
    if args.workload == "all" and args.scale >= 1.0:
        # cfg3: compute_distances, bounded rows, all host threads (rayon's par_iter over rows)
        T, S, rows = 200_000, 100, 1000
        X, tids = synth.synth_panel(3, T, S)
        t0 = time.perf_counter()
        oracle.wis_distances(X, tids, 100, 0, rows, nthreads=ncpu)
        d3 = time.perf_counter() - t0
        workloads["cfg3"] = {"metric": "WIS bin-pairs/s", "value": rows * float(T) / d3, "unit": "bin-pairs/s",
                             "cores": ncpu, "kind": "port",
                             "sample": "compute_distances, first %d of %d rows x all columns x %d samples" % (rows, T, S)}
        del X
        # cfg4: the oracle's sequential Viterbi, one lane per thread
        from concurrent.futures import ThreadPoolExecutor
        from cnvetti_b200.synth import HG38_AUTOSOMES
        lanes = [synth.synth_cv_lanes(400 + i, (L + 999) // 1000)[0] for i, L in enumerate(HG38_AUTOSOMES)]
        tab = oracle.hmm_tables(0.08)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(ncpu) as ex:
            list(ex.map(lambda x: oracle.hmm_viterbi(x, tab), lanes))
        d4 = time.perf_counter() - t0
        workloads["cfg4"] = {"metric": "HMM bins/s", "value": sum(x.size for x in lanes) / d4, "unit": "bins/s",
                             "cores": min(ncpu, 22), "kind": "port", "sample": "one sample (22 lanes), one lane per thread"}
        # cfg1: CoverageAggregator restatement on the first 4 M reads of chr22
        d = synth.synth_bam_records(22, 13_000_000, 2_000_000)
        rb = oracle.ReadBatch(d["pos"], d["isize"], d["flag"], d["mapq"], d["cigar_off"], d["cigar"])
        t0 = time.perf_counter()
        oracle.coverage(rb, oracle.cov_opts(window_length=1000, min_mapq=MIN_MAPQ), 13_001_000)
        d1 = time.perf_counter() - t0
        workloads["cfg1"] = {"metric": "aligned reads/s binned (count kind Coverage)", "value": rb.n / d1, "unit": "reads/s",
                             "cores": 1, "kind": "port", "sample": "%d reads of a 13 Mbp contig, single thread" % rb.n}
    workload = ("cfg2: cmd coverage (Fragments, %d bp windows, MAPQ>=%d, GC from reference) + normalize "
                "MedianGcBinned, synthetic 30x WGS chr1-22" % (WINDOW, MIN_MAPQ))
    line = {
        "impl": "reference", "metric": "aligned reads/s binned", "value": val, "unit": "reads/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": {"workload": workload, "sample": desc},
        "cpu_baseline": {"value": val, "unit": "reads/s", "cores": 1, "kind": "port", "sample": desc,
                         "host_cores_available": ncpu},
        "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "workloads": workloads,
        "note": "restated reference CPU path (oracle/, plain C -O2); the Rust reference cannot be built "
                "here; cmd coverage/normalize are single-threaded in the reference",
    }
    print(json.dumps(line))
    return 0


def bench_coverage(args, env):
    """The headline: BASELINE configs[1] (see the module docstring)."""
    torch = env.torch
    import cnvetti_b200 as cb
    from cnvetti_b200 import pipeline, synth_torch
    rank, world, dev, ctx, comm = env.rank, env.world, env.dev, env.ctx, env.comm
    regs = contigs(args.scale)
    n_pairs = [max(1000, int(round(L * READS_PER_BP / 2))) for _, L in regs]
    total_reads = 2 * sum(n_pairs)
    total_bp = sum(L for _, L in regs)
    workload = ("cfg2: cmd coverage (Fragments, %d bp windows, MAPQ>=%d, GC from reference) + normalize "
                "MedianGcBinned, synthetic 30x WGS chr1-22" % (WINDOW, MIN_MAPQ))
    opts = cb.CoverageOptions(window_length=WINDOW, min_mapq=MIN_MAPQ)
    owner, per_rank = pipeline.shard_regions([L for _, L in regs], world)
    mine = per_rank[rank]

    def make_regions(ids, seed_shift=0):
        out, n = [], 0
        for ci in ids:
            name, L = regs[ci]
            soa = synth_torch.synth_soa_contig(2 * 1000 + ci + seed_shift, L, n_pairs[ci], dev)
            seq = synth_torch.synth_reference_torch(2 * 1000 + ci, L, dev)
            batch = cb.ReadBatch(mid=soa["mid"].contiguous(), mapq=soa["mapq"].contiguous(), flags=soa["flags"].contiguous())
            out.append(pipeline.Region(name, L, batch, seq))
            n += len(batch)
            del soa
        return out, n

    t_gen = time.perf_counter()
    dev_regions, n_local = make_regions(mine)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t_gen
    prep_dev = pipeline.PreparedRegions(dev_regions, opts)
    shard = comm if world > 1 else None

    # Steps are enqueued without a host wait in between (cnvb_coverage_run_async): the tallies and the
    # reference-panic checks of step i are collected while step i + 1 runs, so at most two steps are in flight
    # and every step still delivers everything the synchronous call does.
    in_flight = []

    def step_device():
        t = pipeline.coverage_run(prep_dev, opts, ctx=ctx, defer=True)
        pipeline.normalize_run(t, "MedianGcBinned", ctx=ctx, comm=shard, want_medians=False)
        in_flight.append(t)
        if len(in_flight) > 1:
            in_flight.pop(0).wait()
        return t

    def flush():
        while in_flight:
            in_flight.pop(0).wait()
    step_device.flush = flush

    # The timed steps replay the region loop as a CUDA graph (cnvb_coverage_plan_*: one launch instead of ~10 API
    # calls per region -- on 8 GPUs, with 2 - 3 small regions per rank, the eager loop is bound by the host's
    # enqueueing); two plans alternate so that the tallies and reference-panic checks of step i are still collected
    # while step i + 1 runs.  The normaliser (with its all-reduces) follows eagerly on the same stream.
    # --no-graph: the eager loop above; the per-kernel timing pass always uses the eager loop.
    plans = []
    if not args.no_graph:
        try:
            plans = [pipeline.CoveragePlan(prep_dev, opts, ctx=ctx) for _ in range(2)]
        except Exception as exc:  # (reported in the line: `graph` false)
            env.log("coverage plan unavailable, eager steps: %s" % exc)
            plans = []
    plan_turn = [0]
    plan_busy = []

    def step_graph():
        p = plans[plan_turn[0] % 2]
        plan_turn[0] += 1
        if p in plan_busy:      # (its previous launch: two steps back)
            plan_busy.remove(p)
            p.wait()
        tt = p.launch()
        pipeline.normalize_run(tt, "MedianGcBinned", ctx=ctx, comm=shard, want_medians=False)
        plan_busy.append(p)
        return tt

    def flush_graph():
        while plan_busy:
            plan_busy.pop(0).wait()
    step_graph.flush = flush_graph
    step_timed = step_graph if plans else step_device

    # correctness guard on the full-size run: every counted fragment lands in exactly one window
    t = step_device()
    flush()
    torch.cuda.synchronize()
    assert int(t.rcv.double().sum().item()) == sum(t.num_processed) - sum(t.num_skipped), "checksum"
    n_windows_local = int(t.rcv.numel())
    n_windows = int(env.sum_over_ranks(n_windows_local))
    sampler = ClockSampler(env.local)  # armed before the warm-up (nvidia-smi needs ~100 ms to come up)
    if plans:  # the planned loop must deliver what the eager loop delivered above
        tg = step_graph()
        flush_graph()
        torch.cuda.synchronize()
        assert tg.num_processed == t.num_processed and tg.num_skipped == t.num_skipped, "planned loop: tallies"
        same = lambda a, b: torch.equal(a.view(torch.int32), b.view(torch.int32))  # noqa: E731  (bits: missing values are NaNs)
        assert same(tg.rcv, t.rcv) and same(tg.cv, t.cv) and torch.equal(tg.norm_flags, t.norm_flags), "planned loop: columns"
    ms_max, launches, clocks = timed_steps(env, step_timed, args.steps, max(0, args.warmup - 1), sampler)
    used_graph = bool(plans)
    for p_ in plans:  # (their aggregators, tables and pinned slots go back before the other legs allocate)
        p_.close()
    plans = []
    kern, ms_serial = kernel_times(env, step_device, args.steps,
                                   ("frag_bin_gw_kernel", "refstats_lut_kernel", "gc_radix_hist_kernel",
                                    "gc_radix_pick_kernel", "norm_gc_map_kernel", "cov_records_kernel"))
    value = total_reads / (ms_max * 1e-3)
    loads = [sum(regs[i][1] for i in lst) for lst in per_rank]

    if args.quick:
        if rank == 0:
            return {"metric": "aligned reads/s binned", "value": value, "unit": "reads/s", "n_gpus": world,
                    "ms_per_step": ms_max, "gpu_launches": launches, "ms_per_step_serialised": ms_serial, "kernels": kern,
                    "quick": True, "scaling": "strong" if world > 1 else "weak", "graph": used_graph}, 0
        return None, 0

    # ---- end to end: pinned host SoA in, table out ---------------------------------------------------------
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
    out_host = {kcol: torch.empty(n_windows_local, dtype=dt_, pin_memory=True)
                for kcol, dt_ in (("cv", torch.float32), ("cv2", torch.float32), ("lcv", torch.float32),
                                  ("rcv", torch.float32), ("gc", torch.float32), ("norm_flags", torch.uint8),
                                  ("gap", torch.uint8))}
    d2h = sum(v.numel() * v.element_size() for v in out_host.values())
    prep_host = pipeline.PreparedRegions(host_regions, opts)

    def step_e2e():
        tt = pipeline.coverage_run(prep_host, opts, ctx=ctx)
        pipeline.normalize_run(tt, "MedianGcBinned", ctx=ctx, comm=shard, want_medians=False)
        for kcol, h in out_host.items():
            h.copy_(getattr(tt, kcol), non_blocking=True)
        torch.cuda.synchronize()

    e2e_steps = max(1, min(args.steps, 5))
    step_e2e()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    env.barrier()
    e2e_s = env.max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    e2e_value = total_reads / e2e_s
    h2d_total = int(env.sum_over_ranks(h2d))
    d2h_total = int(env.sum_over_ranks(d2h))
    del host_regions, prep_host, out_host

    # ---- N > 1: the reference's own data parallelism beside it (one sample per GPU, no collective) ------------
    replicas = None
    if world > 1 and not args.no_replicas:
        del dev_regions, prep_dev, t
        env.free()
        rep_regions, n_rep = make_regions(range(len(regs)), seed_shift=7919 * rank)
        prep_rep = pipeline.PreparedRegions(rep_regions, opts)

        def step_rep():
            tt = pipeline.coverage_run(prep_rep, opts, ctx=ctx)
            return pipeline.normalize_run(tt, "MedianGcBinned", ctx=ctx, want_medians=False)

        rsteps = max(3, min(args.steps, 10))
        ms_rep, _, _ = timed_steps(env, step_rep, rsteps, 2)
        replicas = {"value": world * n_rep / (ms_rep * 1e-3), "unit": "reads/s", "scaling": "weak", "ms_per_step": ms_rep,
                    "steps": rsteps, "reads_per_gpu": n_rep,
                    "what": "one whole-genome sample per GPU, no collective (quick_wis_build_model.rs:148-184)"}
        del rep_regions, prep_rep

    # ---- rank 0: roofline, CPU baseline, parity of the sample contigs against the oracle ----------------------
    rc = 0
    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        step_bytes = float(BYTES_PER_READ_GW * total_reads + total_bp + n_windows * (5 + 4 + 16 + 8))
        gw = kern["frag_bin_gw_kernel"]
        achieved = BYTES_PER_READ_GW * n_local / (gw["ms_per_step"] * 1e-3) / 1e9 if gw["ms_per_step"] > 0 else 0.0
        traffic, traffic_src = ncu_traffic("frag_bin_gw_kernel", n_local / max(1.0, gw["launches_per_step"]))
        roofline = {"bound": "hbm", "kernel": "frag_bin_gw_kernel", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peak_src, "algorithmic_bytes_per_read": BYTES_PER_READ_GW,
                    "algorithmic_bytes_per_launch": BYTES_PER_READ_GW * n_local / max(1.0, gw["launches_per_step"]),
                    "kernel_share_of_step": gw["ms_per_step"] / ms_serial,
                    "kernel_timing": "instrumented pass with the regions serialised on one stream (%.3f ms/step); the "
                                     "timed steps overlap regions on three side streams; rank 0's regions" % ms_serial,
                    "step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_max * 1e-3) / 1e9, "unit": "GB/s",
                             "frac": step_bytes / (ms_max * 1e-3) / 1e9 / (peak * world), "peak": peak * world,
                             "what": "all kernels of one step together, all ranks: 7 B/read + 1 B/base + 5 B/window "
                                     "refstats out + 4 B/window counts + normalise 8 B/window x 2 passes + 8 B/window out"}}
        rs = kern.get("refstats_lut_kernel")
        if rs and rs["ms_per_step"] > 0:
            rs["achieved_gbs"] = sum(regs[i][1] for i in mine) / (rs["ms_per_step"] * 1e-3) / 1e9
        cpu, checks = None, None
        if world == 1:
            k = max(1, min(args.cpu_sample_contigs, len(regs)))
            sample_ids = list(range(len(regs) - k, len(regs)))
            sample = [(synth_torch.soa_to_oracle_records(synth_torch.synth_soa_contig(2 * 1000 + ci, regs[ci][1], n_pairs[ci], dev)),
                       regs[ci][1], synth_torch.synth_reference_torch(2 * 1000 + ci, regs[ci][1], dev).cpu().numpy())
                      for ci in sample_ids]
            cn, cdt, o = cpu_reference_leg(sample, 1, 1)
            del sample
            cpu = {"value": cn / cdt, "unit": "reads/s", "cores": 1, "kind": "port",
                   "sample": "%s (%.0f M reads) of the workload, oracle coverage+refstats+normalize, single thread "
                             "(the reference's cmd coverage/normalize are single-threaded)" %
                             ("+".join(regs[i][0] for i in sample_ids), cn / 1e6),
                   "host_cores_available": os.cpu_count()}
            # the same contigs through the GPU path (region loop + MedianGcBinned over these contigs) vs the oracle
            sub = [dev_regions[i] for i in sample_ids]
            g = pipeline.normalize_run(pipeline.coverage_run(sub, opts, ctx=ctx), "MedianGcBinned", ctx=ctx)
            torch.cuda.synchronize()
            bits = lambda x: np.ascontiguousarray(x.cpu().numpy() if hasattr(x, "cpu") else x, np.float32).view(np.uint32)  # noqa: E731
            cv2_g, fin = g.cv2.cpu().numpy(), np.isfinite(o["cv2"])
            parts = {
                "counts": bool(np.array_equal(bits(g.rcv), bits(o["rcv"]))) and g.num_processed == o["processed"],
                "lcv_bits": bool(np.array_equal(bits(g.lcv), bits(o["lcv"]))),
                "gc_bits": bool(np.array_equal(bits(g.gc), bits(o["gc"]))),
                "gap": bool(np.array_equal(g.gap.cpu().numpy(), o["gap"])),
                "medians_bits": bool(np.array_equal(bits(g.medians), bits(o["medians"]))),
                "cv_bits": bool(np.array_equal(bits(g.cv), bits(o["cv"]))),
                "flags": bool(np.array_equal(g.norm_flags.cpu().numpy(), o["flags"])),
                "cv2_rel_1e-6": bool(np.array_equal(np.isfinite(cv2_g), fin) and
                                     np.allclose(cv2_g[fin], o["cv2"][fin], rtol=1e-6, atol=1e-7)),
            }
            checks = {"what": "%s (%d windows, %.0f M reads) through the region loop + MedianGcBinned against the oracle"
                              % ("+".join(regs[i][0] for i in sample_ids), g.rcv.numel(), cn / 1e6),
                      "identical": all(parts.values()), **parts}
            if not checks["identical"]:
                rc = 1
        line = {
            "metric": "aligned reads/s binned", "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload, "reads": total_reads, "windows": n_windows,
                       "contigs": len(regs), "window_length": WINDOW, "scale": args.scale,
                       "l2": "inputs (%.1f GB/step) far larger than the 126 MB L2; no flush needed" %
                             ((BYTES_PER_READ_GW * total_reads + total_bp) / 1e9),
                       "parallelism": "single GPU" if world == 1 else
                       "ONE sample, contigs sharded over %d GPUs (LPT by length; largest share %.1f %% of the genome, ideal "
                       "%.1f %%); MedianGcBinned exchanges 4 x 103 KB radix histograms (NCCL all-reduce inside the "
                       "library); result columns stay sharded" % (world, 100.0 * max(loads) / total_bp, 100.0 / world),
                       "contigs_of_rank": [[regs[i][0] for i in lst] for lst in per_rank] if world > 1 else None},
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": d2h_total,
                    "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                    "host_memory": "pinned" if all_pinned else "pageable (pinning refused)"},
            "gpu_launches": launches,
            "graph": {"used": used_graph, "what": "the timed steps replay the region loop as a CUDA graph (cnvb_coverage_plan_*), "
                      "two plans alternating; the normaliser follows eagerly; gpu_launches counts the kernels inside the "
                      "replays; --no-graph runs the eager loop"},
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "kernels": kern, "gen_seconds": gen_s, "checks": checks, "replicas": replicas,
            "host_bam_decode": host_bam_decode_leg(),
            "host_bam_decode_realistic": host_bam_decode_leg(n_pairs=200_000, payload="random"),
        }
    return line, rc


def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the B200 path has no CPU fallback"}))
        return 2
    env = Env(args)
    rc = 0
    single = {"wis": bench_wis, "hmm": bench_hmm, "piles": bench_piles, "cfg1": bench_cfg1, "cfg5": bench_cfg5}
    if args.workload in single:
        line = single[args.workload](args, env)
        if line and isinstance(line.get("checks"), dict) and line["checks"].get("identical") is False:
            rc = 1
    else:
        line, rc = bench_coverage(args, env)
        env.free()
        if args.workload == "all" and not args.quick and args.scale >= 1.0:
            # the other BASELINE metrics, each a full line of its own; bounded step counts keep the run short
            subs = {}
            plan = [("cfg1", bench_cfg1, min(args.steps, 20), min(args.warmup, 3)),
                    ("cfg3", bench_wis, min(args.steps, 10), min(args.warmup, 3)),
                    ("cfg4", bench_hmm, min(args.steps, 10), min(args.warmup, 3)),
                    ("cfg5", bench_cfg5, 2, 1)]
            for name, fn, k, w in plan:
                env.log("workload %s" % name)
                try:
                    sub = fn(args, env, steps=k, warmup=w)
                except Exception as ex:  # a failing side workload must not lose the headline
                    sub = {"error": "%s: %s" % (type(ex).__name__, ex)}
                    rc = rc or 1
                    env.free()
                if env.rank == 0:
                    subs[name] = sub
                    if sub and isinstance(sub.get("checks"), dict) and sub["checks"].get("identical") is False:
                        rc = 1
            if line is not None:
                line["workloads"] = subs
    if env.rank == 0 and line is not None:
        if rc:
            line["parity_failed"] = True
        print(json.dumps(line))
    env.close()
    return rc


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
