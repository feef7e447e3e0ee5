"""Coverage VCF output and input (`cmd coverage` / `cmd normalize` files).

write_coverage_vcf   the table of pipeline.coverage_run / normalize_run -> `.vcf` or `.vcf.gz`, through
                     the library's writer (csrc/vcf.cu; record layout lib-coverage/src/lib.rs:560-673)
read_coverage_vcf    the columns back (what lib-normalize / lib-merge-cov read: FORMAT/LCV, FORMAT/CV,
                     INFO/GC ...), a plain host parser for interchange and for the tests
"""
import ctypes as C
import gzip

import numpy as np

from . import _native as N


def _host(x, dtype):
    if x is None:
        return None
    if hasattr(x, "cpu"):
        x = x.cpu().numpy()
    a = np.ascontiguousarray(x)
    if a.dtype != dtype:
        a = a.view(dtype) if a.dtype.itemsize == np.dtype(dtype).itemsize else a.astype(dtype)
    return a


def write_coverage_vcf(path, table, sample, contigs, is_targets=False, command_line=None, file_date=None, threads=0):
    """`table`: pipeline.CoverageTable (after coverage_run, optionally normalize_run); `contigs`:
    [(name, length)] for the ##contig lines -- table.names must be among them."""
    names = [c[0] for c in contigs]
    lens = np.asarray([c[1] for c in contigs], np.uint32)
    index = {n: i for i, n in enumerate(names)}
    total = table.offsets[-1]
    contig = np.empty(total, np.uint32)
    for name, lo, hi in zip(table.names, table.offsets[:-1], table.offsets[1:]):
        contig[lo:hi] = index[name]
    keep = [contig, lens]
    cols = N.VcfColumns()
    cols.contig = contig.ctypes.data
    for field, attr, dt in (("start", "start", np.int32), ("end", "end", np.int32), ("rcv", "rcv", np.float32),
                            ("rcvsd", "rcvsd", np.float32), ("lcv", "lcv", np.float32), ("lcvsd", "lcvsd", np.float32),
                            ("frm", "frm", np.float32), ("gc", "gc", np.float32), ("gap", "gap", np.uint8),
                            ("ft", "ft", np.uint8), ("cv", "cv", np.float32), ("cv2", "cv2", np.float32),
                            ("cvsd", "cvsd", np.float32), ("norm_flags", "norm_flags", np.uint8)):
        a = _host(getattr(table, attr, None), dt)
        if a is not None:
            assert a.shape[0] == total, field
            keep.append(a)
            setattr(cols, field, a.ctypes.data)
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    opts = N.VcfOpts(str(path).encode(), sample.encode(), len(names), arr, lens.ctypes.data,
                     command_line.encode() if command_line else None, file_date.encode() if file_date else None,
                     1 if is_targets else 0, int(threads))
    err = C.create_string_buffer(512)
    rc = N.lib().cnvb_vcf_write_coverage(C.byref(opts), C.byref(cols), total, err, 512)
    if rc != N.OK:
        raise N.CnvbError(rc, err.value.decode() or "writing %s failed" % path)
    return path


_MISSING = np.frombuffer(np.uint32(N.F32_MISSING_BITS).tobytes(), np.float32)[0]


def read_coverage_vcf(path):
    """-> dict(header=[...], samples=[...], chrom, pos (0-based), id, alt, info={key: list}, format={key: array})
    for single-sample coverage files.  Absent FORMAT values come back as the BCF missing pattern."""
    opener = gzip.open if str(path).endswith(".gz") else open
    header, samples, rows = [], [], []
    with opener(path, "rt") as f:
        for line in f:
            if line.startswith("##"):
                header.append(line.rstrip("\n"))
            elif line.startswith("#"):
                samples = line.rstrip("\n").split("\t")[9:]
            else:
                rows.append(line.rstrip("\n").split("\t"))
    n = len(rows)
    out = dict(header=header, samples=samples, chrom=[r[0] for r in rows], pos=np.array([int(r[1]) - 1 for r in rows], np.int64),
               id=[r[2] for r in rows], alt=[r[4] for r in rows], info={}, format={})
    fkeys = sorted({k for r in rows for k in r[8].split(":")})
    for k in fkeys:
        out["format"][k] = [None] * n if k in ("GT", "FT") else np.full(n, _MISSING, np.float32)
    ikeys = sorted({kv.split("=")[0] for r in rows for kv in r[7].split(";")})
    for k in ikeys:
        out["info"][k] = [None] * n
    for i, r in enumerate(rows):
        for kv in r[7].split(";"):
            k, _, v = kv.partition("=")
            out["info"][k][i] = v if v else True
        for k, v in zip(r[8].split(":"), r[9].split(":")):
            if k in ("GT", "FT"):
                out["format"][k][i] = v
            elif v != ".":
                out["format"][k][i] = np.float32(v)
    return out
