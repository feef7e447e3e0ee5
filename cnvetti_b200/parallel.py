"""Multi-GPU drivers: one process per GPU, `torch.distributed` (NCCL over NVLink on the GPU box,
gloo in CPU tests) for the plumbing.  SURVEY.md section 8(e):

  * coverage / refstats / HMM shard by region (contig) or lane with NO data-path collective;
    only the finished per-window columns are gathered (`gather_columns`);
  * WIS shards by target-row block: all-gather of the panel, local distances for the rank's
    rows, all-gather of one f32 per target (the nearest distance) for the global prune
    threshold, then local prune / per-probe calls / filters.

The compute backend is injectable so that the host logic (partitioning, collectives, result
assembly) can be exercised with gloo on a CPU-only box; the default backend is the CUDA one and
there is no CPU fallback in the product.
"""
import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def shard_rows(n, rank, world):
    """Contiguous, balanced row block [lo, hi) of rank `rank`."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_weight(weights, world):
    """Longest-processing-time assignment of work items (contigs, lanes) to ranks.
    Returns a list of index lists, one per rank; deterministic."""
    order = sorted(range(len(weights)), key=lambda i: (-weights[i], i))
    loads = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda q: (loads[q], q))
        out[r].append(i)
        loads[r] += weights[i]
    for lst in out:
        lst.sort()
    return out


def _world(group=None):
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


class Communicator:
    """cnvb_comm of this rank: NCCL inside the library (include/cnvetti_b200.h "communicator"), on the
    context's stream.  `torch.distributed` (any backend) is used ONCE, to hand rank 0's NCCL unique id to
    the other ranks; a host without torch passes the 128 bytes itself (`unique_id=`)."""

    def __init__(self, ctx, rank=None, world=None, group=None, unique_id=None):
        import ctypes as C
        from . import _native as N
        if rank is None or world is None:
            rank, world = _world(group)
        self.ctx, self.rank, self.world = ctx, int(rank), int(world)
        h = C.c_void_p()
        L = N.lib()
        if self.world == 1:
            ctx.check(L.cnvb_comm_from_nccl(ctx.handle, None, 0, 1, C.byref(h)))
        else:
            ident = (C.c_char * 128)()
            if unique_id is None:
                box = [None]
                if self.rank == 0:
                    if L.cnvb_comm_unique_id(ident) != N.OK:
                        raise N.CnvbError(N.ERR_COMM, L.cnvb_comm_create_error().decode())
                    box = [bytes(ident.raw)]
                dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0,
                                           group=group)
                unique_id = box[0]
            ident.raw = bytes(unique_id)
            ctx.check(L.cnvb_comm_init_rank(ctx.handle, ident, self.rank, self.world, C.byref(h)))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            from . import _native as N
            N.lib().cnvb_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- collectives on device tensors (ctx stream) ---------------------------------------------------
    def all_gather_counts(self, n):
        """every rank's integer -> list (one tiny exchange + host wait)"""
        if self.world == 1:
            return [int(n)]
        t = torch.zeros(self.world, dtype=torch.int64, device="cuda:%d" % self.ctx.device)
        t[self.rank] = int(n)
        self.all_gather_into(t, [1] * self.world)
        return [int(x) for x in t.cpu().tolist()]

    def all_gather_into(self, full, counts):
        """`full` (contiguous, dim 0 = concatenation of the ranks' blocks of counts[r] rows) holds this
        rank's block in place; on return it holds every rank's."""
        import ctypes as C
        from . import _native as N
        row = full[0].numel() * full.element_size() if full.shape[0] else 0
        bytes_ = (C.c_uint64 * self.world)(*[int(c) * row for c in counts])
        off = sum(int(c) for c in counts[:self.rank]) * row
        self.ctx.check(N.lib().cnvb_comm_all_gather_v(self.handle, full.data_ptr() + off, full.data_ptr(), bytes_))
        return full

    def all_gather_rows(self, block, counts=None):
        """blocks of possibly different lengths along dim 0 -> the concatenation in rank order"""
        if self.world == 1:
            return block
        counts = counts or self.all_gather_counts(block.shape[0])
        full = torch.empty((sum(counts),) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
        lo = sum(counts[:self.rank])
        full[lo:lo + block.shape[0]].copy_(block)
        return self.all_gather_into(full, counts)

    def all_to_all(self, send, send_counts, recv, recv_counts):
        """send / recv: contiguous, dim 0 cut into per-rank blocks of send_counts[q] / recv_counts[q] rows"""
        import ctypes as C
        from . import _native as N
        srow = send[0].numel() * send.element_size() if send.shape[0] else 0
        rrow = recv[0].numel() * recv.element_size() if recv.shape[0] else 0
        so = (C.c_uint64 * (self.world + 1))(*np.concatenate([[0], np.cumsum(send_counts)]).astype(np.int64) * srow)
        ro = (C.c_uint64 * (self.world + 1))(*np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64) * rrow)
        self.ctx.check(N.lib().cnvb_comm_all_to_all_v(self.handle, send.data_ptr(), so, recv.data_ptr(), ro))
        return recv


def all_gather_rows(block, group=None):
    """All-gather row blocks of possibly different lengths along dim 0 (sizes are exchanged
    first; NCCL needs equal shapes, so blocks are padded to the longest)."""
    rank, world = _world(group)
    if world == 1:
        return block
    n = torch.tensor([block.shape[0]], dtype=torch.int64, device=block.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
    pad[:block.shape[0]] = block
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


class CudaWisBackend:
    """The product backend: cnvetti_b200.model_wis on the rank's GPU."""

    def __init__(self, ctx=None):
        from . import model_wis
        self.mw = model_wis
        self.ctx = ctx

    def distances(self, X, tids, options, lo, hi):
        return self.mw.compute_distances(X, tids, options, row_begin=lo, row_end=hi, ctx=self.ctx)

    def threshold(self, nearest, z):
        return self.mw.prune_threshold(nearest, z, ctx=self.ctx)

    def prune(self, dist_, idx, thr):
        return self.mw.prune_distances(dist_, idx, thr, ctx=self.ctx)

    def calls(self, X, ref_idx, ref_len, options, lo, hi):
        return self.mw.count_per_probe_calls(X, ref_idx, ref_len, options, row_begin=lo, row_end=hi, ctx=self.ctx)

    def filters(self, ref_len, calls, options):
        return self.mw.model_filters(ref_len, calls, options, ctx=self.ctx)

    # the stages either side of build-model-wis in the cohort pipeline (pipeline.wis_call_cohort)
    def normalize_total(self, lcv_row, out_row):
        from .normalize import run_uncorrected_normalization
        run_uncorrected_normalization(lcv_row, ctx=self.ctx, want_sum=False, out=out_row)

    def modcov_panel(self, X, filt, ref_idx, ref_len, lo, hi):
        return self.mw.mod_coverage_panel(X, filt, ref_idx, ref_len, want=("cv_rel", "cvz"), row_begin=lo,
                                          row_end=hi, ctx=self.ctx)

    def calls_modcov_panel(self, X, ref_idx, ref_len, options, lo, hi):
        """per-probe calls, filters and mod-coverage of the cohort in one pass -> (calls, filt, dict)"""
        return self.mw.calls_and_mod_coverage_panel(X, ref_idx, ref_len, options, want=("cv_rel", "cvz"), row_begin=lo,
                                                    row_end=hi, ctx=self.ctx)

    def merge_cov_panel(self, cv_sample_major):
        from . import pipeline
        return pipeline.merge_cov_panel(cv_sample_major, ctx=self.ctx)

    def segment(self, cv_flat, lane_off, sigma, options):
        from . import discover
        return discover.hmm_segment(cv_flat, lane_off, sigma, options, ctx=self.ctx)


def wis_run_distributed(X_block, tids_block, options, backend=None, group=None):
    """`build-model-wis` with targets sharded by row block.

    Every rank passes ITS rows of the panel (`X_block[rows_r][S]`, `tids_block[rows_r]`, row
    blocks in rank order = target order).  Returns the model of the rank's rows as a dict
    (ref_dist, ref_idx, ref_len, num_calls, filt, threshold, rows=(lo, hi)); ref_idx are global
    target indices."""
    rank, world = _world(group)
    backend = backend or CudaWisBackend()
    X = all_gather_rows(X_block, group)            # panel: 80 MB (cfg 3) / 12 GB (cfg 5) over NVLink
    tids = all_gather_rows(tids_block, group)
    sizes = torch.tensor([X_block.shape[0]], dtype=torch.int64, device=X_block.device)
    if world > 1:
        allsz = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(allsz, sizes, group=group)
        lo = int(sum(int(s.item()) for s in allsz[:rank]))
    else:
        lo = 0
    hi = lo + int(X_block.shape[0])
    return wis_rows_distributed(X, tids, options, lo, hi, backend=backend, group=group)


def wis_rows_distributed(X, tids, options, lo, hi, backend=None, group=None, with_modcov=False, comm=None):
    """The part of `build-model-wis` after the panel exchange: every rank holds the whole panel
    X[T][S] and computes the model of its target rows [lo, hi) (row blocks in rank order).  The
    one exchange is the all-gather of the nearest distance of every target, from which all ranks
    derive the same prune threshold.  with_modcov: the cohort is called against its own model, and a
    backend that can do so forms the per-probe calls and the mod-coverage values in one pass over the
    panel (key "modcov" of the result)."""
    rank, world = (comm.rank, comm.world) if comm is not None else _world(group)
    backend = backend or CudaWisBackend()
    dist_, idx = backend.distances(X, tids, options, lo, hi)
    nearest_local = dist_[:, 0].contiguous() if hasattr(dist_, "contiguous") else np.ascontiguousarray(dist_[:, 0])
    if comm is not None and world > 1:
        T = int(X.shape[0])
        nearest = comm.all_gather_rows(nearest_local, [shard_rows(T, r, world)[1] - shard_rows(T, r, world)[0]
                                                       for r in range(world)])
    elif world > 1:
        nearest_t = nearest_local if torch.is_tensor(nearest_local) else torch.from_numpy(nearest_local)
        nearest_t = nearest_t.to(X.device)
        nearest = all_gather_rows(nearest_t, group)  # one f32 per target
    else:
        nearest = nearest_local
    thr = backend.threshold(nearest, options.filter_z_score)   # identical on every rank
    ref_idx, ref_len = backend.prune(dist_, idx, thr)
    modcov = None
    if with_modcov and hasattr(backend, "calls_modcov_panel"):
        calls, filt, modcov = backend.calls_modcov_panel(X, ref_idx, ref_len, options, lo, hi)
    else:
        calls = backend.calls(X, ref_idx, ref_len, options, lo, hi)
        filt = backend.filters(ref_len, calls, options)
    return dict(ref_dist=dist_, ref_idx=ref_idx, ref_len=ref_len, num_calls=calls, filt=filt,
                threshold=thr, rows=(lo, hi), modcov=modcov)


def exchange_rows_to_samples(block, sample_ranges, group=None, comm=None):
    """All-to-all between the two shardings of a [S][T] table: every rank holds ALL samples for
    ITS target rows (`block[S][rows_r]`, row blocks in rank order) and receives ITS samples
    (`sample_ranges[r] = (s_lo, s_hi)`) for ALL targets -> [S_r][T]."""
    if comm is not None and comm.world > 1:
        # block[S][rows_r] is cut along the samples into contiguous per-rank pieces; from rank q arrive its
        # rows of this rank's samples, [S_r][rows_q], back to back
        rows = comm.all_gather_counts(block.shape[1])
        s_lo, s_hi = sample_ranges[comm.rank]
        n_s = s_hi - s_lo
        send = block.contiguous().view(-1)
        recv = torch.empty(n_s * sum(rows), dtype=block.dtype, device=block.device)
        comm.all_to_all(send, [(b - a) * block.shape[1] for a, b in sample_ranges], recv, [n_s * r for r in rows])
        off = np.concatenate([[0], np.cumsum([n_s * r for r in rows])]).astype(np.int64)
        return torch.cat([recv[off[q]:off[q + 1]].view(n_s, rows[q]) for q in range(comm.world)], dim=1)
    rank, world = _world(group)
    if world == 1:
        return block
    send = [block[a:b].contiguous() for a, b in sample_ranges]
    n_rows = torch.tensor([block.shape[1]], dtype=torch.int64, device=block.device)
    rows = [torch.zeros_like(n_rows) for _ in range(world)]
    dist.all_gather(rows, n_rows, group=group)
    s_lo, s_hi = sample_ranges[rank]
    recv = [torch.empty((s_hi - s_lo, int(r.item())), dtype=block.dtype, device=block.device) for r in rows]
    if dist.get_backend(group) == "gloo":  # (CPU tests: gloo has no all_to_all)
        for src in range(world):
            if src == rank:
                recv[src].copy_(send[src])
        work = []
        for q in range(world):
            if q != rank:
                work.append(dist.isend(send[q], q, group=group))
        for q in range(world):
            if q != rank:
                dist.recv(recv[q], q, group=group)
        for w in work:
            w.wait()
    else:
        dist.all_to_all(recv, send, group=group)
    return torch.cat(recv, dim=1)


def gather_columns(local_columns, owner_of_region, region_sizes, dtype, device, group=None):
    """Assemble per-window columns computed region-by-region on different ranks into the
    genome-order table on every rank.  `local_columns`: {region index: 1-D tensor} for the
    regions this rank owns.  No arithmetic, only broadcasts and a concatenation."""
    rank, world = _world(group)
    out = []
    for r, n in enumerate(region_sizes):
        owner = owner_of_region[r]
        if owner == rank or world == 1:
            t = local_columns[r].contiguous()
        else:
            t = torch.empty(n, dtype=dtype, device=device)
        if world > 1:
            dist.broadcast(t, src=owner, group=group)
        out.append(t)
    return torch.cat(out) if out else torch.empty(0, dtype=dtype, device=device)
