"""Host-side mirror of lib-coverage's operator interface for the hot path.

Names, argument meaning and error behaviour follow the reference:
  CoverageOptions            lib-coverage/src/options.rs:35-90
  BamRecordAggregator        lib-coverage/src/agg.rs:49-64 (trait)
  FragmentsGenomeWideAggregator / FragmentsTargetRegionsAggregator / CoverageAggregator
                             lib-coverage/src/agg.rs:68-257 / 261-444 / 468-600
  AggregationStats           lib-coverage/src/agg.rs:17-32
  ReferenceStats             lib-coverage/src/reference.rs:16-131
Everything computes on the GPU through the C ABI (include/cnvetti_b200.h); nothing here does
arithmetic on the host.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _native as N
from .context import Context, buf, default_context, same_mem


@dataclass
class CoverageOptions:
    """The counting-related members of CoverageOptions (options.rs:55-85), reference defaults
    from cnvetti/src/cli.yaml:132-160."""
    min_mapq: int = 0
    min_unclipped: float = 0.6
    min_window_remaining: float = 0.5
    min_raw_coverage: int = 10
    window_length: Optional[int] = None
    count_kind: str = "Fragments"            # CountKind::{Coverage, Fragments}
    considered_regions: str = "GenomeWide"   # ConsideredRegions::{GenomeWide, TargetRegions}
    plp_flag_mask: int = 0                   # 0 = htslib default mask (see cnvb_cov_opts)

    def c_opts(self):
        return N.CovOpts(int(self.min_mapq), float(np.float32(self.min_unclipped)),
                         int(self.window_length or 0),
                         float(np.float32(self.min_window_remaining)),
                         int(self.min_raw_coverage), int(self.plp_flag_mask))


@dataclass
class ReadBatch:
    """Structure-of-arrays read batch (cnvb_read_soa).  numpy arrays (host) or torch tensors."""
    pos: object = None
    end: object = None
    mid: object = None
    frag_end: object = None
    mapq: object = None
    flags: object = None
    tid: object = None

    def __len__(self):
        for x in (self.mapq, self.flags, self.pos, self.mid):
            if x is not None:
                return int(x.shape[0])
        return 0

    def c_soa(self):
        keep = []
        mems = []
        ptrs = {}
        for name, dt in (("tid", np.int32), ("pos", np.int32), ("end", np.int32),
                         ("mid", np.int32), ("frag_end", np.int32), ("mapq", np.uint8),
                         ("flags", np.uint16)):
            p, m, k = buf(getattr(self, name), dt if name != "flags" else None)
            if name == "flags" and p is not None:
                x = getattr(self, name)
                if hasattr(x, "is_cuda"):
                    if x.element_size() != 2:
                        raise TypeError("flags must be a 16-bit tensor")
                else:
                    p, m, k = buf(x, np.uint16)
            ptrs[name] = p
            keep.append(k)
            mems.append(m)
        soa = N.ReadSoa(len(self), ptrs["tid"], ptrs["pos"], ptrs["end"], ptrs["mid"],
                        ptrs["frag_end"], ptrs["mapq"], ptrs["flags"], same_mem(*mems))
        return soa, keep


def pack_reads(pos, isize, flag, cigar_off, cigar, min_unclipped=0.6):
    """BAM-level fields -> SoA fields exactly as the reference derives them (cnvb_pack_reads).
    Host code of the product (what a BAM decoder thread runs); needs no GPU."""
    pos = np.ascontiguousarray(pos, np.int32)
    isize = np.ascontiguousarray(isize, np.int32)
    flag = np.ascontiguousarray(flag, np.uint16)
    cigar_off = np.ascontiguousarray(cigar_off, np.uint64)
    cigar = np.ascontiguousarray(cigar, np.uint32)
    n = pos.size
    end = np.empty(n, np.int32)
    mid = np.empty(n, np.int32)
    frag_end = np.empty(n, np.int32)
    flags = np.empty(n, np.uint16)
    rc = N.lib().cnvb_pack_reads(n, pos.ctypes.data, isize.ctypes.data, flag.ctypes.data,
                                 cigar_off.ctypes.data, cigar.ctypes.data if cigar.size else None,
                                 float(np.float32(min_unclipped)), end.ctypes.data,
                                 mid.ctypes.data, frag_end.ctypes.data, flags.ctypes.data)
    if rc != N.OK:
        raise N.CnvbError(rc, "cnvb_pack_reads: invalid argument")
    return end, mid, frag_end, flags


@dataclass
class AggregationStats:
    """agg.rs:17-32, for all regions at once (arrays of num_regions)."""
    cov: object
    cov_sd: object          # None = Option::None
    start: object
    end: object
    frac_removed: object    # None = Option::None
    mean_mapq: object


class BamRecordAggregator:
    """agg.rs:49-64.  `put_fetched_records` takes SoA batches instead of a bam::IndexedReader."""
    _kind = None

    def __init__(self, options, contig_length, targets=None, piles=None, ctx=None):
        self.ctx = ctx or default_context()
        self.options = options
        self.contig_length = int(contig_length)
        L = N.lib()
        h = C.c_void_p()
        ts = te = ps = pe = None
        nt = npile = 0
        if targets is not None:
            ts = np.ascontiguousarray(targets[0], np.uint32)
            te = np.ascontiguousarray(targets[1], np.uint32)
            nt = ts.size
        if piles is not None:
            ps = np.ascontiguousarray(piles[0], np.uint32)
            pe = np.ascontiguousarray(piles[1], np.uint32)
            npile = ps.size
        self._keep = (ts, te, ps, pe)
        opts = options.c_opts()
        rc = L.cnvb_cov_create(self.ctx.handle, C.byref(opts), self._kind, self.contig_length,
                               ts.ctypes.data if nt else None, te.ctypes.data if nt else None, nt,
                               ps.ctypes.data if npile else None,
                               pe.ctypes.data if npile else None, npile, C.byref(h))
        self.ctx.check(rc)
        self.handle = h
        self._state = 0

    def put_fetched_records(self, batch):
        soa, keep = batch.c_soa()
        self.ctx.check(N.lib().cnvb_cov_put_records(self.handle, C.byref(soa)))
        self._keep_batch = keep

    def finish_async(self):
        """Enqueue the end-of-stream work; errors surface at the next finish()."""
        if self._state == 0:
            self._state = 1
            self.ctx.check(N.lib().cnvb_cov_finish_async(self.handle))

    def finish(self):
        """End of the fetch()ed stream: waits, then raises what the reference would panic on."""
        if self._state < 2:
            self._state = 2
            self.ctx.check(N.lib().cnvb_cov_finish(self.handle))

    def num_regions(self):
        return int(N.lib().cnvb_cov_num_regions(self.handle))

    def num_processed(self):
        self.finish()
        return int(N.lib().cnvb_cov_num_processed(self.handle))

    def num_skipped(self):
        self.finish()
        return int(N.lib().cnvb_cov_num_skipped(self.handle))

    def counts(self, device=False):
        self.finish_async() if device else self.finish()
        mem = N.MEM_DEVICE if device else N.MEM_HOST
        out = self.ctx.empty(self.num_regions(), np.uint32 if not device else np.int32, mem)
        p, _, _ = buf(out)
        self.ctx.check(N.lib().cnvb_cov_get_counts(self.handle, p, mem))
        return out

    def get_all_stats(self, device=False, out=None):
        """AggregationStats of every region.  device=True keeps everything on the GPU and does
        not synchronise; `out` may provide preallocated (device) arrays to fill."""
        self.finish_async() if device else self.finish()
        n = self.num_regions()
        mem = N.MEM_DEVICE if device else N.MEM_HOST
        idt = np.int32 if device else np.uint32
        if out is not None:
            cov, sd, start, end, frm, mq = (out.cov, out.cov_sd, out.start, out.end,
                                            out.frac_removed, out.mean_mapq)
        else:
            cov = self.ctx.empty(n, np.float32, mem)
            sd = self.ctx.empty(n, np.float32, mem) if self._kind == N.COVERAGE else None
            start = self.ctx.empty(n, idt, mem)
            end = self.ctx.empty(n, idt, mem)
            frm = self.ctx.empty(n, np.float32, mem) if self._kind == N.FRAGMENTS_GENOME_WIDE else None
            mq = self.ctx.empty(n, np.float32, mem)
        ptr = lambda x: buf(x)[0]
        self.ctx.check(N.lib().cnvb_cov_get_stats(self.handle, ptr(cov), ptr(sd), ptr(start),
                                                  ptr(end), ptr(frm), ptr(mq), mem))
        return AggregationStats(cov, sd, start, end, frm, mq)

    def get_stats(self, region_id):
        s = getattr(self, "_all", None)
        if s is None:
            s = self._all = self.get_all_stats()
        return AggregationStats(float(s.cov[region_id]),
                                None if s.cov_sd is None else float(s.cov_sd[region_id]),
                                int(s.start[region_id]), int(s.end[region_id]),
                                None if s.frac_removed is None else float(s.frac_removed[region_id]),
                                float(s.mean_mapq[region_id]))

    def close(self):
        if getattr(self, "handle", None):
            N.lib().cnvb_cov_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FragmentsGenomeWideAggregator(BamRecordAggregator):
    """agg.rs:68-257.  `tree` = pile intervals (starts, ends) or None."""
    _kind = N.FRAGMENTS_GENOME_WIDE

    def __init__(self, tree, options, contig_length, ctx=None):
        BamRecordAggregator.__init__(self, options, contig_length, piles=tree, ctx=ctx)


class FragmentsTargetRegionsAggregator(BamRecordAggregator):
    """agg.rs:261-444.  `target_regions` = (starts, ends), sorted by start."""
    _kind = N.FRAGMENTS_TARGET_REGIONS

    def __init__(self, target_regions, options, contig_length, ctx=None):
        BamRecordAggregator.__init__(self, options, contig_length, targets=target_regions, ctx=ctx)


class CoverageAggregator(BamRecordAggregator):
    """agg.rs:468-600."""
    _kind = N.COVERAGE

    def __init__(self, options, contig_length, ctx=None):
        BamRecordAggregator.__init__(self, options, contig_length, ctx=ctx)


def make_aggregator(options, contig_length, targets=None, piles=None, ctx=None):
    """The match at lib-coverage/src/lib.rs:510-529."""
    key = (options.count_kind, options.considered_regions)
    if key == ("Fragments", "GenomeWide"):
        return FragmentsGenomeWideAggregator(piles, options, contig_length, ctx)
    if key == ("Fragments", "TargetRegions"):
        return FragmentsTargetRegionsAggregator(targets, options, contig_length, ctx)
    if key == ("Coverage", "GenomeWide"):
        return CoverageAggregator(options, contig_length, ctx)
    raise ValueError("Invalid combination of coverage/on-target regions")


class ReferenceStats:
    """reference.rs:16-131: gc_content / has_gap per window."""

    def __init__(self, gc_content, has_gap):
        self.gc_content = gc_content
        self.has_gap = has_gap

    @staticmethod
    def look_at_chars(seq, window_length, ctx=None, out=None):
        """`out` = (gc, has_gap) preallocated arrays in the same memory space as `seq`."""
        ctx = ctx or default_context()
        p, mem, keep = buf(seq, np.uint8)
        n = int(keep.shape[0])
        nb = (n + window_length - 1) // window_length
        gc, gap = out if out is not None else (ctx.empty(nb, np.float32, mem), ctx.empty(nb, np.uint8, mem))
        ctx.check(N.lib().cnvb_refstats(ctx.handle, p, n, int(window_length), buf(gc)[0],
                                        buf(gap)[0], mem))
        return ReferenceStats(gc, gap)


def cov_records(stats, options, is_targets, ctx=None):
    """FORMAT assembly of process_region (lib-coverage/src/lib.rs:623-668):
    returns (lcv, lcvsd or None, ft bits)."""
    ctx = ctx or default_context()
    pc, mem, _ = buf(stats.cov, np.float32)
    n = int(stats.cov.shape[0])
    lcv = ctx.empty(n, np.float32, mem)
    lcvsd = ctx.empty(n, np.float32, mem) if stats.cov_sd is not None else None
    ft = ctx.empty(n, np.uint8, mem)
    opts = options.c_opts()
    ctx.check(N.lib().cnvb_cov_records(ctx.handle, n, pc, buf(stats.cov_sd)[0], buf(stats.start)[0],
                                       buf(stats.end)[0], buf(stats.frac_removed)[0],
                                       int(bool(is_targets)), C.byref(opts), buf(lcv)[0],
                                       buf(lcvsd)[0], buf(ft)[0], mem))
    return lcv, lcvsd, ft
