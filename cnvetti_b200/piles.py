"""Pile collection for black-listing: mirror of lib-coverage/src/piles.rs (PileCollector) and of
`compute_threshold` / `gather_piles` (lib-coverage/src/lib.rs:265-419)."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native as N
from .context import default_context


@dataclass
class PileInfos:
    """Vec<PileInfo> of one contig as three columns (piles.rs:67-75)."""
    start: np.ndarray
    end: np.ndarray
    size: np.ndarray

    def __len__(self):
        return int(self.start.shape[0])


class PileCollector:
    """piles.rs:24-170: options (pile_max_gap, min_mapq) + collect_piles per contig."""

    def __init__(self, pile_max_gap=20, min_mapq=0, plp_flag_mask=0, ctx=None):
        self.pile_max_gap = int(pile_max_gap)
        self.min_mapq = int(min_mapq)
        self.plp_flag_mask = int(plp_flag_mask)
        self.ctx = ctx or default_context()

    def collect_piles(self, reads, contig_length):
        """`reads`: ReadBatch (host or device) with pos, end, mapq, flags in coordinate order."""
        ctx = self.ctx
        soa, keep = reads.c_soa()
        opts = N.CovOpts(self.min_mapq, 0.0, 0, 0.0, 0, self.plp_flag_mask)
        h = C.c_void_p()
        ctx.check(N.lib().cnvb_piles_collect(ctx.handle, C.byref(soa), C.byref(opts), self.pile_max_gap,
                                             int(contig_length), C.byref(h)))
        try:
            n = int(N.lib().cnvb_piles_count(h))
            out = [np.zeros(n, np.uint32) for _ in range(3)]
            if n:
                ctx.check(N.lib().cnvb_piles_get(h, out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data, N.MEM_HOST))
        finally:
            N.lib().cnvb_piles_destroy(h)
        del keep
        return PileInfos(*out)


def compute_threshold(abs_freq, fdr_cutoff):
    """lib.rs:265-351; raises CnvbError with the reference's bail! text."""
    y = np.ascontiguousarray(abs_freq, np.uint64)
    out = C.c_uint32()
    err = C.create_string_buffer(256)
    rc = N.lib().cnvb_pile_threshold(y.ctypes.data, y.size, float(fdr_cutoff), C.byref(out), err, 256)
    if rc != N.OK:
        raise N.CnvbError(rc, err.value.decode())
    return int(out.value)


def gather_piles(regions, min_mapq=0, pile_max_gap=20, mask_piles_fdr=0.001, ctx=None):
    """lib.rs:353-419: piles of every region, genome-wide size histogram, FDR threshold, and the
    black-listed intervals per region as (start, end) arrays usable as `Region.piles`.
    `regions`: pipeline.Region objects whose reads carry pos, end, mapq, flags."""
    collector = PileCollector(pile_max_gap, min_mapq, ctx=ctx)
    infos = [collector.collect_piles(r.reads, r.length) for r in regions]
    top = max([int(i.size.max()) for i in infos if len(i)] + [0])
    abs_freq = np.zeros(top + 1, np.uint64)
    for i in infos:
        if len(i):
            abs_freq += np.bincount(i.size, minlength=top + 1).astype(np.uint64)
    thresh = compute_threshold(abs_freq, mask_piles_fdr)
    out = []
    for i in infos:
        sel = i.size >= thresh
        out.append((i.start[sel].copy(), i.end[sel].copy()))
    return out, thresh
