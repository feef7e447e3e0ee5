"""In-memory mirror of the region loop of `lib_coverage::run` (lib-coverage/src/lib.rs:768-781,
process_region :422-676) followed by `lib_normalize::run` (lib-normalize/src/lib.rs:43-66).

The reference chains these stages through BCF files; here the per-window columns stay in HBM
(one device column per BCF tag) and only the finished table is read back.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _native as N
from .context import default_context
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


def _num_regions(options, r):
    if options.considered_regions == "TargetRegions":
        return len(r.targets[0])
    wl = options.window_length
    return (r.length + wl - 1) // wl


def coverage_run(regions, options, ctx=None, device=True):
    """`cnvetti cmd coverage` on in-memory inputs: one aggregator per region, in order."""
    import torch
    ctx = ctx or default_context()
    dev = "cuda:%d" % ctx.device
    counts = [_num_regions(options, r) for r in regions]
    total = int(sum(counts))
    t = CoverageTable()
    t.offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64).tolist()
    t.names = [r.name for r in regions]
    f32 = lambda: torch.empty(total, dtype=torch.float32, device=dev)
    i32 = lambda: torch.empty(total, dtype=torch.int32, device=dev)
    t.start, t.end, t.rcv, t.lcv = i32(), i32(), f32(), f32()
    t.ft = torch.empty(total, dtype=torch.uint8, device=dev)
    is_cov = options.count_kind == "Coverage"
    is_tgt = options.considered_regions == "TargetRegions"
    if is_cov:
        t.rcvsd, t.lcvsd = f32(), f32()
    if not is_cov and not is_tgt:
        t.frm = f32()
    mq = f32()
    have_ref = any(r.sequence is not None for r in regions) and options.window_length
    if have_ref:
        t.gc = f32()
        t.gap = torch.empty(total, dtype=torch.uint8, device=dev)
    aggs = []
    for r, lo, hi in zip(regions, t.offsets[:-1], t.offsets[1:]):
        sl = slice(lo, hi)
        if have_ref and r.sequence is not None:  # lib.rs:440-454
            seq = r.sequence
            if not (hasattr(seq, "is_cuda") and seq.is_cuda):  # host bytes: one async H2D copy
                seq = torch.as_tensor(seq).to(dev, non_blocking=True)
            ReferenceStats.look_at_chars(seq, options.window_length, ctx=ctx, out=(t.gc[sl], t.gap[sl]))
        agg = make_aggregator(options, r.length, targets=r.targets, piles=r.piles, ctx=ctx)  # lib.rs:510-529
        agg.put_fetched_records(r.reads)                                                    # lib.rs:547
        out = AggregationStats(t.rcv[sl], t.rcvsd[sl] if is_cov else None, t.start[sl], t.end[sl],
                               t.frm[sl] if t.frm is not None else None, mq[sl])
        agg.get_all_stats(device=True, out=out)                                             # lib.rs:560-561
        aggs.append(agg)
    # FORMAT assembly for all rows at once (lib.rs:623-668)
    stats = AggregationStats(t.rcv, t.rcvsd, t.start, t.end, t.frm, mq)
    lcv, lcvsd, ft = cov_records(stats, options, is_tgt, ctx=ctx)
    t.lcv, t.lcvsd, t.ft = lcv, lcvsd, ft
    # the single synchronisation point: tallies + reference-panic checks of every region
    for agg in aggs:
        agg.finish()
        t.num_processed.append(agg.num_processed())
        t.num_skipped.append(agg.num_skipped())
        agg.close()
    return t


def normalize_run(table, normalization="MedianGcBinned", ctx=None):
    """`cnvetti cmd normalize` on the table (lib-normalize/src/lib.rs:57-60)."""
    ctx = ctx or default_context()
    if normalization == "TotalCovSum":
        table.cv, table.cvsd, _ = run_uncorrected_normalization(table.lcv, table.lcvsd, ctx=ctx)
    elif normalization == "MedianGcBinned":
        if table.gc is None:
            raise N.CnvbError(N.ERR_INVALID, "Could not access INFO/GC. Did you forget to pass in reference "
                              "when computing coverage so GC content was not computed?")
        r = run_binned_correction(table.lcv, table.gc, table.lcvsd, ctx=ctx)
        table.cv, table.cv2, table.cvsd, table.norm_flags = r["cv"], r["cv2"], r["cvsd"], r["flags"]
    else:
        raise ValueError("Unknown normalization")
    return table
