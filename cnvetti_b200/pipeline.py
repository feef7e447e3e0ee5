"""In-memory mirror of the region loop of `lib_coverage::run` (lib-coverage/src/lib.rs:768-781,
process_region :422-676) followed by `lib_normalize::run` (lib-normalize/src/lib.rs:43-66).

The reference chains these stages through BCF files; here the per-window columns stay in HBM
(one device column per BCF tag) and only the finished table is read back.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

import ctypes as C

from . import _native as N
from .context import buf, default_context
from .coverage import AggregationStats, ReferenceStats, cov_records, make_aggregator
from .normalize import run_binned_correction, run_uncorrected_normalization


@dataclass
class Region:
    """One entry of GenomeRegions (lib-shared/src/regions.rs:15-19) plus its inputs."""
    name: str
    length: int                      # region end == contig length for whole contigs
    reads: object                    # ReadBatch (host or device) in coordinate order
    sequence: Optional[object] = None  # contig bytes (uint8) if --reference was given
    targets: Optional[tuple] = None
    piles: Optional[tuple] = None


@dataclass
class CoverageTable:
    """The per-record columns `cmd coverage` writes (INFO END/GC/GAP, FORMAT RCV/LCV/...), all
    regions concatenated in processing order.  Device tensors unless read back."""
    offsets: List[int] = field(default_factory=list)   # row offset of every region
    names: List[str] = field(default_factory=list)
    start: object = None
    end: object = None
    rcv: object = None
    rcvsd: object = None
    lcv: object = None
    lcvsd: object = None
    frm: object = None
    ft: object = None
    gc: object = None
    gap: object = None
    num_processed: List[int] = field(default_factory=list)
    num_skipped: List[int] = field(default_factory=list)
    # filled by normalize()
    cv: object = None
    cv2: object = None
    cvsd: object = None
    norm_flags: object = None
    medians: object = None
    _pending: object = None          # (handle, ctx, n_regions) of a deferred run

    def wait(self):
        """After coverage_run(..., defer=True): block until the regions' tallies have arrived, fill
        num_processed / num_skipped and raise what the synchronous call would have raised."""
        if self._pending is not None:
            handle, ctx, n = self._pending
            self._pending = None
            nproc = np.zeros(max(1, n), np.uint64)
            nskip = np.zeros(max(1, n), np.uint64)
            ctx.check(N.lib().cnvb_coverage_wait(handle, nproc.ctypes.data, nskip.ctypes.data))
            self.num_processed = [int(x) for x in nproc[:n]]
            self.num_skipped = [int(x) for x in nskip[:n]]
        return self


def _num_regions(options, r):
    if options.considered_regions == "TargetRegions":
        return len(r.targets[0])
    wl = options.window_length
    return (r.length + wl - 1) // wl


_KIND = {("Fragments", "GenomeWide"): N.FRAGMENTS_GENOME_WIDE,
         ("Fragments", "TargetRegions"): N.FRAGMENTS_TARGET_REGIONS,
         ("Coverage", "GenomeWide"): N.COVERAGE}


class PreparedRegions:
    """The regions of one run as the C array `cnvb_coverage_run` takes (built once, reusable while
    the underlying buffers stay alive and in place)."""

    def __init__(self, regions, options):
        key = (options.count_kind, options.considered_regions)
        if key not in _KIND:
            raise ValueError("Invalid combination of coverage/on-target regions")  # lib.rs:528
        self.kind = _KIND[key]
        self.regions = list(regions)
        self.c_opts = options.c_opts()
        self.keep = []
        arr = (N.Region * max(1, len(self.regions)))()
        for i, r in enumerate(self.regions):
            c = arr[i]
            c.region_end = int(r.length)
            soa, keep = r.reads.c_soa()
            c.reads = soa
            self.keep.append(keep)
            if r.sequence is not None and options.window_length and self.kind != N.FRAGMENTS_TARGET_REGIONS:
                p, mem, k = buf(r.sequence, np.uint8)
                c.seq, c.seq_len, c.seq_mem = p, int(k.shape[0]), mem
                self.keep.append(k)
            if r.targets is not None and self.kind == N.FRAGMENTS_TARGET_REGIONS:
                ts = np.ascontiguousarray(r.targets[0], np.uint32)
                te = np.ascontiguousarray(r.targets[1], np.uint32)
                c.tgt_start, c.tgt_end, c.n_tgt = ts.ctypes.data, te.ctypes.data, ts.size
                self.keep += [ts, te]
            if r.piles is not None and self.kind == N.FRAGMENTS_GENOME_WIDE:
                ps = np.ascontiguousarray(r.piles[0], np.uint32)
                pe = np.ascontiguousarray(r.piles[1], np.uint32)
                c.pile_start, c.pile_end, c.n_pile = ps.ctypes.data, pe.ctypes.data, ps.size
                self.keep += [ps, pe]
        self.array = arr
        self.n = len(self.regions)
        self.have_ref = any(arr[i].seq for i in range(self.n))
        self.counts = [int(N.lib().cnvb_coverage_rows(C.byref(self.c_opts), self.kind,
                                                      C.cast(C.byref(arr[i]), C.POINTER(N.Region)), 1))
                       for i in range(self.n)]
        self.offsets = np.concatenate([[0], np.cumsum(self.counts)]).astype(np.int64).tolist()
        self.names = [r.name for r in self.regions]


def coverage_run(regions, options, ctx=None, device=True, defer=False):
    """`cnvetti cmd coverage` on in-memory inputs: the region loop of lib_coverage::run
    (lib.rs:768-781) as one native call (cnvb_coverage_run); columns stay in HBM.
    defer=True (device-resident inputs): nothing waits -- the columns are valid in stream order, the
    tallies and the reference-panic checks arrive with table.wait() (cnvb_coverage_run_async)."""
    import torch
    ctx = ctx or default_context()
    prep = regions if isinstance(regions, PreparedRegions) else PreparedRegions(regions, options)
    dev = "cuda:%d" % ctx.device
    total = prep.offsets[-1]
    t = CoverageTable()
    t.offsets = prep.offsets
    t.names = prep.names
    is_cov = prep.kind == N.COVERAGE
    is_gw = prep.kind == N.FRAGMENTS_GENOME_WIDE
    f32 = lambda: torch.empty(total, dtype=torch.float32, device=dev)
    u8 = lambda: torch.empty(total, dtype=torch.uint8, device=dev)
    i32 = lambda: torch.empty(total, dtype=torch.int32, device=dev)
    t.start, t.end, t.rcv, t.lcv, t.ft = i32(), i32(), f32(), f32(), u8()
    if is_cov:
        t.rcvsd, t.lcvsd = f32(), f32()
    if is_gw:
        t.frm = f32()
    if prep.have_ref:
        t.gc, t.gap = f32(), u8()
    ptr = lambda x: x.data_ptr() if x is not None else None
    cols = N.CovColumns(ptr(t.start), ptr(t.end), ptr(t.rcv), ptr(t.rcvsd), ptr(t.lcv), ptr(t.lcvsd),
                        ptr(t.frm), ptr(t.gc), None, ptr(t.gap), ptr(t.ft))
    if defer:
        handle = C.c_void_p()
        ctx.check(N.lib().cnvb_coverage_run_async(ctx.handle, C.byref(prep.c_opts), prep.kind, prep.array, prep.n,
                                                  C.byref(cols), C.byref(handle)))
        t._pending = (handle, ctx, prep.n)
        return t
    nproc = np.zeros(max(1, prep.n), np.uint64)
    nskip = np.zeros(max(1, prep.n), np.uint64)
    ctx.check(N.lib().cnvb_coverage_run(ctx.handle, C.byref(prep.c_opts), prep.kind, prep.array, prep.n,
                                        C.byref(cols), nproc.ctypes.data, nskip.ctypes.data, N.MEM_DEVICE))
    t.num_processed = [int(x) for x in nproc[:prep.n]]
    t.num_skipped = [int(x) for x in nskip[:prep.n]]
    return t


class CoveragePlan:
    """The region loop over the SAME device buffers again and again, as a CUDA graph (cnvb_coverage_plan_*): for
    pipelines that copy every sample into fixed device columns, where the host's enqueueing -- not the GPU -- bounds
    a run over a few small regions.  `table` holds the output columns (allocated once); launch() enqueues one run
    and returns the table (columns valid in stream order), wait() blocks and fills num_processed / num_skipped.
    Count kind FragmentsGenomeWide, device-resident regions without piles (the library refuses the rest)."""

    def __init__(self, regions, options, ctx=None):
        import torch
        self.ctx = ctx or default_context()
        self.prep = regions if isinstance(regions, PreparedRegions) else PreparedRegions(regions, options)
        prep, dev = self.prep, "cuda:%d" % self.ctx.device
        total = prep.offsets[-1]
        t = CoverageTable()
        t.offsets, t.names = prep.offsets, prep.names
        f32 = lambda: torch.empty(total, dtype=torch.float32, device=dev)  # noqa: E731
        t.start, t.end = (torch.empty(total, dtype=torch.int32, device=dev) for _ in range(2))
        t.rcv, t.lcv, t.ft = f32(), f32(), torch.empty(total, dtype=torch.uint8, device=dev)
        if prep.kind == N.COVERAGE:
            t.rcvsd, t.lcvsd = f32(), f32()
        if prep.kind == N.FRAGMENTS_GENOME_WIDE:
            t.frm = f32()
        if prep.have_ref:
            t.gc, t.gap = f32(), torch.empty(total, dtype=torch.uint8, device=dev)
        ptr = lambda x: x.data_ptr() if x is not None else None  # noqa: E731
        self._cols = N.CovColumns(ptr(t.start), ptr(t.end), ptr(t.rcv), ptr(t.rcvsd), ptr(t.lcv), ptr(t.lcvsd),
                                  ptr(t.frm), ptr(t.gc), None, ptr(t.gap), ptr(t.ft))
        self.table = t
        self._h = C.c_void_p()
        self.ctx.check(N.lib().cnvb_coverage_plan_create(self.ctx.handle, C.byref(prep.c_opts), prep.kind, prep.array,
                                                         prep.n, C.byref(self._cols), C.byref(self._h)))
        self._nproc = np.zeros(max(1, prep.n), np.uint64)
        self._nskip = np.zeros(max(1, prep.n), np.uint64)

    def launch(self):
        self.ctx.check(N.lib().cnvb_coverage_plan_launch(self._h))
        return self.table

    def wait(self):
        self.ctx.check(N.lib().cnvb_coverage_plan_wait(self._h, self._nproc.ctypes.data, self._nskip.ctypes.data))
        self.table.num_processed = [int(x) for x in self._nproc[:self.prep.n]]
        self.table.num_skipped = [int(x) for x in self._nskip[:self.prep.n]]
        return self.table

    def close(self):
        if self._h:
            N.lib().cnvb_coverage_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def normalize_run(table, normalization="MedianGcBinned", ctx=None, comm=None, want_medians=True):
    """`cnvetti cmd normalize` on the table (lib-normalize/src/lib.rs:57-60).  comm: the table holds this
    rank's regions of a sharded run (coverage_run_sharded); the statistics run over all ranks' records."""
    ctx = (comm.ctx if comm is not None else ctx) or default_context()
    if normalization == "TotalCovSum":
        table.cv, table.cvsd, _ = run_uncorrected_normalization(table.lcv, table.lcvsd, ctx=ctx, comm=comm,
                                                                want_sum=want_medians)
    elif normalization == "MedianGcBinned":
        if table.gc is None:
            raise N.CnvbError(N.ERR_INVALID, "Could not access INFO/GC. Did you forget to pass in reference "
                              "when computing coverage so GC content was not computed?")
        r = run_binned_correction(table.lcv, table.gc, table.lcvsd, ctx=ctx, comm=comm, want_medians=want_medians)
        table.cv, table.cv2, table.cvsd, table.norm_flags = r["cv"], r["cv2"], r["cvsd"], r["flags"]
        table.medians = r["medians"]
    else:
        raise ValueError("Unknown normalization")
    return table


def shard_regions(lengths, world):
    """Which rank processes which region of `lib_coverage::run`'s loop (lib-coverage/src/lib.rs:768-781):
    longest-processing-time assignment by region length (reads are ~proportional to it).  Returns
    (owner[region], regions_of_rank[rank]); deterministic, the same on every rank."""
    from . import parallel
    per_rank = parallel.shard_by_weight([int(x) for x in lengths], world)
    owner = [0] * len(lengths)
    for r, lst in enumerate(per_rank):
        for i in lst:
            owner[i] = r
    return owner, per_rank


def coverage_run_sharded(local_regions, options, comm, device=True):
    """The region loop with the regions (contigs) sharded over the ranks of `comm`: this rank processes
    `local_regions` (its share under shard_regions, in genome order) exactly as the single-GPU loop would --
    regions are independent (lib.rs:768-781), no data-path collective.  The table holds the rank's regions;
    normalize_run(table, ..., comm=comm) then normalises against the statistics of ALL ranks' records and
    gather_table() assembles the genome-order table where one is wanted."""
    return coverage_run(local_regions, options, ctx=comm.ctx, device=device)


def gather_table(table, owner, comm, columns=("start", "end", "rcv", "lcv", "gc", "gap", "ft", "cv", "cv2", "norm_flags")):
    """Genome-order columns on every rank from the ranks' sharded tables (what the reference's single
    output file holds).  `owner[i]` = rank of region i (shard_regions); the ranks' tables list their regions
    in genome order.  One all-gather per column; collective."""
    import torch
    world = comm.world
    # rows per region, from every rank (regions are in genome order within a rank)
    mine = [hi - lo for lo, hi in zip(table.offsets[:-1], table.offsets[1:])]
    n_regions = len(owner)
    sizes = torch.zeros(n_regions, dtype=torch.int64, device="cuda:%d" % comm.ctx.device)
    my_ids = [i for i, r in enumerate(owner) if r == comm.rank]
    assert len(my_ids) == len(mine), "table does not hold this rank's regions"
    counts_per_rank = [sum(1 for r in owner if r == q) for q in range(world)]
    packed = torch.tensor(mine, dtype=torch.int64, device=sizes.device)
    allsz = comm.all_gather_rows(packed, counts_per_rank).cpu().tolist()
    pos = [0] * world
    starts = np.concatenate([[0], np.cumsum(counts_per_rank)]).astype(np.int64)
    region_rows = []
    for i in range(n_regions):
        q = owner[i]
        region_rows.append(int(allsz[starts[q] + pos[q]]))
        pos[q] += 1
    rows_per_rank = [sum(region_rows[i] for i in range(n_regions) if owner[i] == q) for q in range(world)]
    # rank-major concatenation -> genome order
    rank_off = np.concatenate([[0], np.cumsum(rows_per_rank)]).astype(np.int64)
    within = [0] * world
    pieces = []
    for i in range(n_regions):
        q = owner[i]
        pieces.append((int(rank_off[q] + within[q]), region_rows[i]))
        within[q] += region_rows[i]
    out = CoverageTable()
    out.offsets = np.concatenate([[0], np.cumsum(region_rows)]).astype(np.int64).tolist()
    for name in columns:
        col = getattr(table, name, None)
        if col is None:
            continue
        full = comm.all_gather_rows(col, rows_per_rank)
        setattr(out, name, torch.cat([full[a:a + n] for a, n in pieces]) if pieces else full)
    return out


def merge_cov(tables, ctx=None):
    """`cnvetti cmd merge-cov` on in-memory tables (lib-merge-cov/src/lib.rs:102-232): the per-sample
    FORMAT/CV columns side by side -> panel X[T][S] (row-major f32) and the contig id of every record
    (ModelInput, lib-model-wis/src/lib.rs:49-62).  All samples must describe the same records."""
    import torch
    ctx = ctx or default_context()
    if not tables:
        raise ValueError("merge_cov: no input")
    first = tables[0]
    for t in tables[1:]:
        if t.offsets != first.offsets or not torch.equal(t.start, first.start) or not torch.equal(t.end, first.end):
            raise N.CnvbError(N.ERR_INVALID, "merge-cov: the inputs do not describe the same regions")
    T, S = int(first.offsets[-1]), len(tables)
    cols = [t.cv.contiguous() for t in tables]
    X = torch.empty((T, S), dtype=torch.float32, device=cols[0].device)
    ptrs = (C.c_void_p * S)(*[c.data_ptr() for c in cols])
    ctx.check(N.lib().cnvb_merge_cov(ctx.handle, ptrs, None, 0, S, T, X.data_ptr()))
    tids = torch.empty(T, dtype=torch.int32, device=X.device)
    for i, (lo, hi) in enumerate(zip(first.offsets[:-1], first.offsets[1:])):
        tids[lo:hi] = i
    return X, tids


def merge_cov_panel(cv_sample_major, ctx=None):
    """The same for a cohort whose columns already sit in one block cv[S][T] (sample-major): -> X[T][S]."""
    import torch
    ctx = ctx or default_context()
    S, T = int(cv_sample_major.shape[0]), int(cv_sample_major.shape[1])
    X = torch.empty((T, S), dtype=torch.float32, device=cv_sample_major.device)
    ctx.check(N.lib().cnvb_merge_cov(ctx.handle, None, cv_sample_major.data_ptr(), T, S, T, X.data_ptr()))
    return X


def quick_wis_build_model(samples, wis_options=None, ctx=None):
    """`cnvetti quick wis-build-model` (cnvetti/src/quick_wis_build_model.rs:133-214) on in-memory
    inputs: per sample cmd coverage (Fragments on TargetRegions, MAPQ >= 0, min_unclipped 0.6; :44-70)
    and cmd normalize (TotalCovSum; :72-80), merge-cov, build-model-wis.  `samples`: one list of
    Region objects (with .targets) per sample.  Returns (model, X, tids)."""
    from .coverage import CoverageOptions
    from . import model_wis
    ctx = ctx or default_context()
    opts = CoverageOptions(window_length=None, count_kind="Fragments", considered_regions="TargetRegions",
                           min_mapq=0, min_unclipped=0.6, min_window_remaining=0.5)
    tables = [normalize_run(coverage_run(regions, opts, ctx=ctx), "TotalCovSum", ctx=ctx) for regions in samples]
    X, tids = merge_cov(tables, ctx=ctx)
    # (the quick command's own preset: max_samples_reliable 8, not the 4 of `cmd build-model-wis`)
    model = model_wis.run(X, tids, wis_options or model_wis.BuildModelWisOptions.quick(), ctx=ctx)
    return model, X, tids


def wis_call_cohort(lcv, tids, sample_ranges=None, wis_options=None, discover_options=None, sigma=0.08,
                    segment=True, ctx=None, group=None, backend=None, timings=None, comm=None):
    """BASELINE config 5, in memory: germline calling of a whole cohort against its own panel --
    per sample `cmd normalize --normalization TotalCovSum`, `merge-cov`, `build-model-wis`,
    per sample `mod-coverage` (the chain of cnvetti/src/quick_wis_build_model.rs:133-214 and
    quick_wis_call.rs:124-190 without the files), then `discover` on the relative coverage.

    `lcv[S_r][T]`: FORMAT/LCV of THIS rank's samples (device tensor, one row per sample);
    `sample_ranges[r] = (s_lo, s_hi)` for every rank (None: single process); `tids[T]` contig id of
    every target (targets in genome order).  Sharding: normalisation by
    sample, then ONE all-gather of the normalised panel; distances / prune / calls / mod-coverage by
    target-row block (one all-gather of T nearest distances for the threshold); an all-to-all brings
    the relative coverage back to sample-major shards for the (sample, contig) HMM lanes.
    `comm` (parallel.Communicator): the exchanges run inside the library (NCCL on the context's stream: the
    samples' normalised columns are gathered straight into one [S][T] block and turned into X[T][S] by the
    merge-cov kernel); without it they go through torch.distributed (`group`).
    `backend` (default: the CUDA library on this rank's GPU) is injectable so that the host logic can
    be exercised with gloo on a CPU-only box.

    Returns dict(model=..., rows=(lo, hi), cv_rel[S_r][T], cvz_rows[S][rows], states[S_r][T] uint8)."""
    import os
    import sys
    import time
    import torch
    from . import model_wis, parallel
    if comm is not None:
        ctx = comm.ctx
    backend = backend or parallel.CudaWisBackend(ctx or default_context())
    wis_options = wis_options or model_wis.BuildModelWisOptions.quick()  # quick_wis_build_model.rs:88-105
    rank, world = (comm.rank, comm.world) if comm is not None else parallel._world(group)
    S_r, T = int(lcv.shape[0]), int(lcv.shape[1])
    if sample_ranges is None:
        sample_ranges = [(0, S_r)]
    S = sample_ranges[-1][1]
    dev = lcv.device

    progress = bool(os.environ.get("CNVB_PROGRESS"))

    def mark(name, t0):
        if progress:
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            sys.stderr.write("[wis_call_cohort] %s done (%.3f s)\n" % (name, time.perf_counter() - t0))
            sys.stderr.flush()
        if timings is not None:
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            timings[name] = timings.get(name, 0.0) + (time.perf_counter() - t0)
        return time.perf_counter()

    t0 = time.perf_counter()
    # cmd normalize, one sample at a time (its sum runs over all targets of the sample); the normalised
    # columns land in this rank's rows of the cohort block [S][T]
    if hasattr(backend, "merge_cov_panel"):
        full = torch.empty((S, T), dtype=lcv.dtype, device=dev)
        cv = full[sample_ranges[rank][0]:sample_ranges[rank][1]]
    else:
        full, cv = None, torch.empty_like(lcv)
    for s in range(S_r):
        backend.normalize_total(lcv[s], cv[s])
    t0 = mark("normalize", t0)
    # merge-cov: the one big exchange, then the panel in ModelInput layout X[T][S]
    if full is not None:
        if comm is not None:
            comm.all_gather_into(full, [b - a for a, b in sample_ranges])
        elif world > 1:
            full = parallel.all_gather_rows(cv, group)
        X = backend.merge_cov_panel(full)
        del full
    else:
        X = parallel.all_gather_rows(cv, group).t().contiguous()  # (injected CPU backends)
    del cv
    t0 = mark("merge_cov", t0)
    lo, hi = parallel.shard_rows(T, rank, world)
    m = parallel.wis_rows_distributed(X, tids, wis_options, lo, hi, backend=backend, group=group, with_modcov=True,
                                      comm=comm)
    t0 = mark("build_model_wis", t0)
    mc = m.pop("modcov")  # (the CUDA backend forms the per-probe calls and these values in one pass)
    if mc is None:
        mc = backend.modcov_panel(X, m["filt"], m["ref_idx"], m["ref_len"], lo, hi)
    t0 = mark("mod_coverage", t0)
    rel = parallel.exchange_rows_to_samples(mc["cv_rel"], sample_ranges, group=group, comm=comm)  # [S_r][T]
    t0 = mark("exchange", t0)
    out = dict(model=m, rows=(lo, hi), cv_rel=rel, cvz_rows=mc["cvz"], panel=X, n_samples=S)
    if segment:
        # (sample, contig) lanes: contig boundaries from the tids of the targets
        t_host = tids.cpu().numpy() if hasattr(tids, "cpu") else np.asarray(tids)
        cuts = np.concatenate([[0], np.flatnonzero(np.diff(t_host)) + 1, [T]]).astype(np.int64)
        lane_off = np.concatenate([s * T + cuts[:-1] for s in range(S_r)] + [[S_r * T]]).astype(np.uint64)
        per = cuts.size - 1  # lanes per sample
        sig = np.full(S_r * per, float(sigma)) if np.isscalar(sigma) else np.asarray(sigma, np.float64)
        states = torch.empty((S_r, T), dtype=torch.uint8, device=dev)
        step = max(1, (1 << 29) // max(T, 1))  # samples per call: at most 2^29 bins each
        for a in range(0, S_r, step):
            b = min(S_r, a + step)
            off = (lane_off[a * per:b * per + 1] - lane_off[a * per]).astype(np.uint64)
            st = backend.segment(rel[a:b].reshape(-1), off, sig[a * per:b * per], discover_options)
            states[a:b] = st.reshape(b - a, T)
        out["states"] = states
        out["lane_off"] = lane_off
        mark("discover", t0)
    return out


# ---- file-level drivers: the two commands of the path with the reference's inputs and outputs -----------
def read_fasta(path):
    """{name: uint8 array} of a FASTA file (`--reference`; the reference reads contigs through an
    indexed reader, lib-coverage/src/reference.rs:58-71 -- same bytes, line breaks removed)."""
    import gzip
    out, name, parts = {}, None, []
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if name is not None:
                    out[name] = np.frombuffer(b"".join(parts), np.uint8)
                name, parts = line[1:].split()[0].decode(), []
            else:
                parts.append(line.rstrip(b"\r\n"))
    if name is not None:
        out[name] = np.frombuffer(b"".join(parts), np.uint8)
    return out


def cmd_coverage(*args, **kwargs):
    """`cnvetti cmd coverage` at file level: see commands.cmd_coverage."""
    from .commands import cmd_coverage as f
    return f(*args, **kwargs)


def cmd_normalize(*args, **kwargs):
    """`cnvetti cmd normalize` at file level: see commands.cmd_normalize."""
    from .commands import cmd_normalize as f
    return f(*args, **kwargs)
