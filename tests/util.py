"""Shared helpers for the parity tests."""
import numpy as np

from cnvetti_b200 import ReadBatch, pack_reads


def oracle_reads(orc, d):
    return orc.ReadBatch(d["pos"], d["isize"], d["flag"], d["mapq"], d["cigar_off"], d["cigar"])


def soa_batch(d, min_unclipped=0.6, device=None, lo=0, hi=None):
    """BAM-level dict -> ReadBatch via the product's host packer; optionally on the GPU."""
    end, mid, frag_end, flags = pack_reads(d["pos"], d["isize"], d["flag"], d["cigar_off"],
                                           d["cigar"], min_unclipped)
    cols = dict(pos=d["pos"], end=end, mid=mid, frag_end=frag_end, mapq=d["mapq"], flags=flags)
    cols = {k: np.ascontiguousarray(v[lo:hi]) for k, v in cols.items()}
    if device is not None:
        import torch
        out = {}
        for k, v in cols.items():
            if v.dtype == np.uint16:
                v = v.view(np.int16)
            out[k] = torch.from_numpy(v).to(device)
        cols = out
    return ReadBatch(**cols)


def to_np(x):
    if hasattr(x, "cpu"):
        x = x.cpu().numpy()
    return np.asarray(x)


def f32_bits(a):
    return np.ascontiguousarray(to_np(a), np.float32).view(np.uint32)


def assert_f32_equal_bits(a, b, what=""):
    """bit-for-bit equality (NaN payloads included)"""
    ba, bb = f32_bits(a), f32_bits(b)
    bad = np.nonzero(ba != bb)[0]
    assert bad.size == 0, "%s: %d mismatches, first at %d: %r vs %r" % (
        what, bad.size, bad[0], to_np(a)[bad[0]], to_np(b)[bad[0]])


def simple_reads(pos, end, flag=None, mapq=None):
    pos = np.asarray(pos, np.int32)
    end = np.asarray(end, np.int32)
    n = pos.size
    cig = ((end - pos).astype(np.uint32) << 4)
    return dict(pos=pos, isize=np.zeros(n, np.int32),
                flag=np.zeros(n, np.uint16) if flag is None else np.asarray(flag, np.uint16),
                mapq=np.full(n, 60, np.uint8) if mapq is None else np.asarray(mapq, np.uint8),
                cigar_off=np.arange(n + 1, dtype=np.uint64), cigar=cig)
