"""The VCF / BCF files that chain the commands of the path, through the library's
htslib-free reader / writer (csrc/bcf.cu): `.bcf` (+ `.csi`), `.vcf.gz` (+ `.tbi`), `.vcf`, exactly as
the file name selects in the reference (lib-coverage/src/lib.rs:174-175, lib-shared/src/bcf_utils.rs:47-82).

VariantFile          a file as columns (what the reference reads record by record through rust-htslib)
write_variant_file   columns -> file, with the index
coverage_header / model_header   the ## lines of lib-coverage/src/lib.rs:67-166 and
                     lib-model-wis/src/lib.rs:280-328 (IDs, Number and Type as there)
write_coverage_file / read_coverage_table   pipeline.CoverageTable <-> file
"""
import ctypes as C
import datetime

import numpy as np

from . import _native as N

MISSING = np.frombuffer(np.uint32(N.F32_MISSING_BITS).tobytes(), np.float32)[0]

# FORMAT keys in the order process_region pushes them (lib-coverage/src/lib.rs:613-664), then what
# lib-normalize (uncorrected.rs:66-82 / correct_binned.rs:105-126) and lib-mod-cov (src/lib.rs:222-240) append
COVERAGE_FORMAT_ORDER = ("RCV", "RCVSD", "LCV", "LCVSD", "FRM", "MQ", "FT", "CV", "CV2", "CVSD", "CVZ")


def _today():
    return datetime.datetime.now(datetime.timezone.utc).strftime("%Y%m%d")


def coverage_header(contigs, command_line=None, file_date=None):
    """build_header of lib-coverage/src/lib.rs:67-166.  Descriptions are free text (the reference's own carry
    line breaks inside the quotes, which htslib truncates); IDs, Number and Type are the contract."""
    lines = ["##fileformat=VCFv4.2",
             "##FILTER=<ID=PASS,Description=\"All filters passed\">",
             "##fileDate=%s" % (file_date or _today()),
             "##cnvetti_cmdCoverageVersion=0.1.0",
             "##cnvetti_cmdCoverageCommand=%s" % (command_line or "cnvetti cmd coverage")]
    lines += ["##contig=<ID=%s,length=%d>" % (n, int(ln)) for n, ln in contigs]
    lines += [
        '##ALT=<ID=WINDOW,Description="Window for read or coverage counting">',
        '##ALT=<ID=TARGET,Description="Target region for fragment counting">',
        '##INFO=<ID=END,Number=1,Type=Integer,Description="Window end">',
        '##INFO=<ID=GC,Number=1,Type=Float,Description="Reference GC fraction of the window">',
        '##INFO=<ID=MEAN_MAPQ,Number=1,Type=Float,Description="Mean MAPQ across samples">',
        '##INFO=<ID=FEW_GCWINDOWS,Number=0,Type=Flag,Description="Too few windows with this GC content for the GC correction">',
        '##INFO=<ID=REF_MEAN,Number=1,Type=Flag,Description="Reference mean coverage">',          # (sic: lib.rs:118)
        '##INFO=<ID=REF_STD_DEV,Number=1,Type=Flag,Description="Reference coverage standard deviation">',
        '##FILTER=<ID=PILE,Description="Piles cover too large a fraction of the window">',
        '##FILTER=<ID=LOWCOV,Description="Too low raw coverage (targeted data)">',
        '##FILTER=<ID=NO_REF,Description="No reference targets for this target">',
        '##INFO=<ID=GAP,Number=0,Type=Flag,Description="Window overlaps N in the reference">',
        '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
        '##FORMAT=<ID=MQ,Number=1,Type=Float,Description="Mean read MAPQ of the region">',
        '##FORMAT=<ID=RCV,Number=1,Type=Float,Description="Raw coverage value">',
        '##FORMAT=<ID=RCVSD,Number=1,Type=Float,Description="Raw coverage standard deviation">',
        '##FORMAT=<ID=LCV,Number=1,Type=Float,Description="Length-normalized coverage value">',
        '##FORMAT=<ID=LCVSD,Number=1,Type=Float,Description="Length-normalized coverage standard deviation">',
        '##FORMAT=<ID=NCV,Number=1,Type=Float,Description="Further normalized coverage value">',
        '##FORMAT=<ID=CV,Number=1,Type=Float,Description="Finalized coverage value">',
        '##FORMAT=<ID=CVSD,Number=1,Type=Float,Description="Finalized coverage standard deviation">',
        '##FORMAT=<ID=CV2,Number=1,Type=Float,Description="Log2 of the finalized coverage value">',
        '##FORMAT=<ID=CVZ,Number=1,Type=Float,Description="Finalized coverage as Z-score of the reference targets">',
        '##FORMAT=<ID=FRM,Number=1,Type=Float,Description="Fraction of the window masked in the sample">',
        '##FORMAT=<ID=FT,Number=1,Type=String,Description="Filter values of the genotype">',
    ]
    return lines


def model_header(contigs, command_line=None, file_date=None):
    """build_header of lib-model-wis/src/lib.rs:280-328 (no samples)."""
    lines = ["##fileformat=VCFv4.2",
             "##FILTER=<ID=PASS,Description=\"All filters passed\">",
             "##fileDate=%s" % (file_date or _today()),
             "##cnvetti_cmdBuildModelWisVersion=0.1.0",
             "##cnvetti_cmdBuildModelWisCommand=%s" % (command_line or "cnvetti cmd build-model-wis")]
    lines += ["##contig=<ID=%s,length=%d>" % (n, int(ln)) for n, ln in contigs]
    lines += [
        '##ALT=<ID=TARGET,Description="Target region for fragment counting">',
        '##INFO=<ID=END,Number=1,Type=Integer,Description="Window end">',
        '##INFO=<ID=NUM_CALLS,Number=1,Type=Integer,Description="Number of samples with calls">',
        '##INFO=<ID=REF_TARGETS,Number=.,Type=String,Description="Target regions used as within-sample reference">',
        '##FILTER=<ID=FEW_REF,Description="Not enough reference targets">',
        '##FILTER=<ID=UNRELIABLE,Description="Called in too many reference panel samples">',
    ]
    return lines


def _np(x, dtype):
    if x is None:
        return None
    if hasattr(x, "cpu"):
        x = x.cpu().numpy()
    a = np.ascontiguousarray(x)
    if a.dtype != dtype:
        a = a.view(dtype) if (a.dtype.itemsize == np.dtype(dtype).itemsize and a.dtype.kind != "f") else a.astype(dtype)
    return a


def write_variant_file(path, header_lines, samples, rid, pos, end, is_targets=False, filters=None, gc=None, gap=None,
                       mean_mapq=None, few_gc=None, ref_mean=None, ref_sd=None, ref_present=None, num_calls=None,
                       ref_off=None, ref_idx=None, fmt=(), ft=None, ft_pos=None, index=True, threads=0):
    """Columns -> `.bcf` / `.vcf.gz` / `.vcf` (csrc/bcf.cu: cnvb_vfile_write).  `fmt`: [(key, values[n][S] or [n],
    present[n] or None)] in record order; `ft`: CNVB_FT_* bits [n][S] written before fmt[ft_pos]."""
    n = int(len(pos))
    S = len(samples)
    keep = []

    def col(x, dtype, shape=None):
        a = _np(x, dtype)
        if a is None:
            return None
        if shape is not None:
            a = np.ascontiguousarray(a.reshape(shape))
        assert a.shape[0] == (shape[0] if shape else n), "column length"
        keep.append(a)
        return a.ctypes.data

    args = N.VfileWriteArgs()
    args.path = str(path).encode()
    text = "\n".join(header_lines) + "\n"
    args.header_text = text.encode()
    args.n_samples = S
    sarr = (C.c_char_p * max(1, S))(*[s.encode() for s in samples])
    args.samples = sarr
    args.n_rows = n
    args.rid, args.pos, args.end = col(rid, np.int32), col(pos, np.int32), col(end, np.int32)
    args.is_targets = 1 if is_targets else 0
    args.filters = col(filters, np.uint8)
    args.gap, args.gc = col(gap, np.uint8), col(gc, np.float32)
    args.mean_mapq = col(mean_mapq, np.float32)
    args.few_gc = col(few_gc, np.uint8)
    args.ref_mean, args.ref_sd, args.ref_present = col(ref_mean, np.float32), col(ref_sd, np.float32), col(ref_present, np.uint8)
    args.num_calls = col(num_calls, np.int32)
    if ref_off is not None:
        ro = _np(ref_off, np.uint64)
        ri = _np(ref_idx, np.uint32)
        assert ro.shape[0] == n + 1 and int(ro[-1]) == ri.shape[0]
        keep += [ro, ri]
        args.ref_off = ro.ctypes.data
        args.ref_idx = ri.ctypes.data if ri.size else np.zeros(1, np.uint32).ctypes.data
    farr = (N.VfileFmt * max(1, len(fmt)))()
    for i, (key, values, present) in enumerate(fmt):
        farr[i].key = key.encode()
        farr[i].values = col(values, np.float32, (n, S) if S else None)
        farr[i].present = col(present, np.uint8)
    args.n_fmt = len(fmt) if S else 0
    args.fmt = farr
    args.ft = col(ft, np.uint8, (n, S)) if (ft is not None and S) else None
    args.ft_pos = len(fmt) if ft_pos is None else int(ft_pos)
    args.write_index = 1 if index else 0
    args.n_threads = int(threads)
    err = C.create_string_buffer(512)
    rc = N.lib().cnvb_vfile_write(C.byref(args), err, 512)
    if rc != N.OK:
        raise N.CnvbError(rc, err.value.decode() or "writing %s failed" % path)
    return path


class VariantFile:
    """A VCF / BCF file as columns (csrc/bcf.cu reader).  Context manager."""

    def __init__(self, path, threads=0):
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = N.lib().cnvb_vfile_open(str(path).encode(), int(threads), C.byref(h), err, 512)
        if rc != N.OK:
            raise N.CnvbError(rc, err.value.decode() or "Could not open BCF file: %s" % path)
        self.handle = h
        self.path = str(path)
        L = N.lib()
        self.n = int(L.cnvb_vfile_num_records(h))
        self.samples = [L.cnvb_vfile_sample(h, i).decode() for i in range(L.cnvb_vfile_num_samples(h))]
        self.contigs = [(L.cnvb_vfile_contig_name(h, i).decode(), int(L.cnvb_vfile_contig_length(h, i)))
                        for i in range(L.cnvb_vfile_num_contigs(h))]
        self.header = L.cnvb_vfile_header_text(h).decode().splitlines()
        self.is_bcf = bool(L.cnvb_vfile_is_bcf(h))
        self._fixed = None

    def close(self):
        if getattr(self, "handle", None):
            N.lib().cnvb_vfile_close(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != N.OK:
            msg = N.lib().cnvb_vfile_last_error(self.handle).decode()
            raise (N.ReferencePanic if rc == N.ERR_REFERENCE_PANIC else N.CnvbError)(rc, msg)

    def fixed(self):
        """dict(rid, pos, end, alt (0 <WINDOW>, 1 <TARGET>, 2 other), filters (VF_* bits), canonical_ids)"""
        if self._fixed is None:
            n = self.n
            rid, pos, end = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.int32)
            alt, filt = np.empty(n, np.uint8), np.empty(n, np.uint8)
            canon = C.c_int(0)
            self._check(N.lib().cnvb_vfile_fixed(self.handle, rid.ctypes.data, pos.ctypes.data, end.ctypes.data,
                                                 alt.ctypes.data, filt.ctypes.data, C.byref(canon)))
            self._fixed = dict(rid=rid, pos=pos, end=end, alt=alt, filters=filt, canonical_ids=bool(canon.value))
        return self._fixed

    def has(self, where, key):
        return any(l.startswith("##%s=<ID=%s," % (where, key)) for l in self.header)

    def info_float(self, key):
        out, present = np.empty(self.n, np.float32), np.empty(self.n, np.uint8)
        self._check(N.lib().cnvb_vfile_info_float(self.handle, key.encode(), out.ctypes.data, present.ctypes.data))
        return out, present

    def info_int(self, key):
        out, present = np.empty(self.n, np.int32), np.empty(self.n, np.uint8)
        self._check(N.lib().cnvb_vfile_info_int(self.handle, key.encode(), out.ctypes.data, present.ctypes.data))
        return out, present

    def info_flag(self, key):
        out = np.empty(self.n, np.uint8)
        self._check(N.lib().cnvb_vfile_info_flag(self.handle, key.encode(), out.ctypes.data))
        return out

    def format_float(self, key):
        """-> (values[n][S], present[n]): absent / missing values carry the BCF missing pattern"""
        S = len(self.samples)
        out, present = np.empty((self.n, S), np.float32), np.empty(self.n, np.uint8)
        self._check(N.lib().cnvb_vfile_format_float(self.handle, key.encode(), out.ctypes.data, present.ctypes.data))
        return out, present

    def format_ft(self):
        S = len(self.samples)
        out, present = np.empty((self.n, S), np.uint8), np.empty(self.n, np.uint8)
        self._check(N.lib().cnvb_vfile_format_ft(self.handle, out.ctypes.data, present.ctypes.data))
        return out, present

    def match_ids(self, other):
        """row of `other` holding the same ID, per record (0xFFFFFFFF: none)"""
        row = np.empty(self.n, np.uint32)
        self._check(N.lib().cnvb_vfile_match_ids(self.handle, other.handle, row.ctypes.data))
        return row

    def ref_targets(self, against=None):
        """INFO/REF_TARGETS as CSR (off[n + 1], idx) over the rows of `against` (default: this file)"""
        off = np.empty(self.n + 1, np.uint64)
        ah = against.handle if against is not None else None
        self._check(N.lib().cnvb_vfile_ref_targets(self.handle, ah, off.ctypes.data, None, 0))
        idx = np.empty(max(1, int(off[-1])), np.uint32)
        self._check(N.lib().cnvb_vfile_ref_targets(self.handle, ah, off.ctypes.data, idx.ctypes.data, idx.size))
        return off, idx[:int(off[-1])]


# ---- pipeline.CoverageTable <-> file --------------------------------------------------------------------
def _table_rid(table, contigs):
    index = {n: i for i, (n, _) in enumerate(contigs)}
    total = table.offsets[-1]
    rid = np.empty(total, np.int32)
    for name, lo, hi in zip(table.names, table.offsets[:-1], table.offsets[1:]):
        if name not in index:
            raise N.CnvbError(N.ERR_INVALID, "Could not find contig %s" % name)   # lib.rs:569-572
        rid[lo:hi] = index[name]
    return rid


def write_coverage_file(path, table, samples, contigs, is_targets=False, header_lines=None, command_line=None,
                        file_date=None, threads=0, extra=None, index=True):
    """The records of `cmd coverage` (+ what `cmd normalize` / `cmd mod-coverage` added) for a single-sample
    table: INFO END [GAP] [GC] MEAN_MAPQ [FEW_GCWINDOWS] [REF_MEAN REF_STD_DEV], FORMAT GT RCV [RCVSD] LCV [LCVSD]
    [FRM] MQ [FT] [CV] [CV2] [CVSD] [CVZ] (lib-coverage/src/lib.rs:560-673).  `samples`: list of sample names (one
    FORMAT column each; the table's values are written for every one of them -- a BAM normally has one)."""
    total = int(table.offsets[-1])
    rid = _table_rid(table, contigs)
    S = len(samples)
    mq = getattr(table, "mean_mapq", None)
    mq = np.full(total, 60.0, np.float32) if mq is None else _np(mq, np.float32)   # agg.rs:241
    nf = _np(getattr(table, "norm_flags", None), np.uint8)

    def rep(x):
        a = _np(x, np.float32)
        return a if S <= 1 else np.repeat(a[:, None], S, axis=1)

    def present(bit, col):
        if col is None:
            return None
        return None if nf is None else ((nf & bit) != 0).astype(np.uint8)

    cols = {"RCV": (table.rcv, None), "RCVSD": (table.rcvsd, None), "LCV": (table.lcv, None), "LCVSD": (table.lcvsd, None),
            "FRM": (table.frm, None), "MQ": (mq, None),
            "CV": (table.cv, present(N.N2_CV, table.cv)), "CV2": (table.cv2, present(N.N2_CV2, table.cv2)),
            "CVSD": (table.cvsd, present(N.N2_CVSD, table.cvsd))}
    if extra:
        cols.update(extra.get("format", {}))
    order = [k for k in COVERAGE_FORMAT_ORDER if k != "FT"]
    if extra and extra.get("format_order"):
        order = list(extra["format_order"])
    fmt = [(k, rep(cols[k][0]), cols[k][1]) for k in order if k in cols and cols[k][0] is not None]
    keys = [f[0] for f in fmt]
    ft_pos = keys.index("MQ") + 1 if "MQ" in keys else len(fmt)
    ft = _np(table.ft, np.uint8)
    if ft is not None and S > 1:
        ft = np.repeat(ft[:, None], S, axis=1)
    few = None if nf is None else ((nf & N.N2_FEW_GCWINDOWS) != 0).astype(np.uint8)
    lines = header_lines or coverage_header(contigs, command_line, file_date)
    kw = dict(extra.get("info", {})) if extra else {}
    return write_variant_file(path, lines, samples, rid, table.start, table.end, is_targets=is_targets,
                              gc=table.gc, gap=table.gap if table.gc is not None else None, mean_mapq=mq, few_gc=few,
                              fmt=fmt, ft=ft if S else None, ft_pos=ft_pos, threads=threads, index=index, **kw)


def read_coverage_table(vf):
    """A single-sample coverage file -> pipeline.CoverageTable (host arrays) -- what lib-normalize / lib-mod-cov
    read: FORMAT/LCV (LCVSD), INFO/GC ...  Records must be grouped by contig (they are, in files of the path)."""
    from .pipeline import CoverageTable
    if len(vf.samples) != 1:
        raise N.CnvbError(N.ERR_INVALID, "Input BCF file must have exactly one sample but had %d" % len(vf.samples))
    fx = vf.fixed()
    t = CoverageTable()
    rid = fx["rid"]
    cuts = np.flatnonzero(np.diff(rid)) + 1 if vf.n else np.zeros(0, np.int64)
    t.offsets = [0] + cuts.tolist() + [vf.n]
    t.names = [vf.contigs[int(rid[o])][0] for o in t.offsets[:-1]] if vf.n else []
    if vf.n == 0:
        t.offsets = [0]
    t.start, t.end = fx["pos"].copy(), fx["end"].copy()

    def f(key):
        if not vf.has("FORMAT", key):
            return None, None
        v, p = vf.format_float(key)
        return (v[:, 0].copy(), p) if p.any() else (None, None)

    t.rcv, _ = f("RCV")
    t.lcv, _ = f("LCV")
    t.rcvsd, _ = f("RCVSD")
    t.lcvsd, _ = f("LCVSD")
    t.frm, _ = f("FRM")
    mq, _ = f("MQ")
    t.mean_mapq = mq
    if vf.has("FORMAT", "FT"):
        ft, p = vf.format_ft()
        t.ft = ft[:, 0].copy()
    else:
        t.ft = np.zeros(vf.n, np.uint8)
    if vf.has("INFO", "GC"):
        gc, p = vf.info_float("GC")
        if p.any():
            gc[p == 0] = MISSING
            t.gc = gc
            t.gap = vf.info_flag("GAP") if vf.has("INFO", "GAP") else np.zeros(vf.n, np.uint8)
    cv, pcv = f("CV")
    if cv is not None:
        cv2, pcv2 = f("CV2")
        cvsd, pcvsd = f("CVSD")
        t.cv, t.cv2, t.cvsd = cv, cv2, cvsd
        few = vf.info_flag("FEW_GCWINDOWS") if vf.has("INFO", "FEW_GCWINDOWS") else np.zeros(vf.n, np.uint8)
        t.norm_flags = ((pcv != 0) * N.N2_CV + (0 if pcv2 is None else (pcv2 != 0) * N.N2_CV2) +
                        (0 if pcvsd is None else (pcvsd != 0) * N.N2_CVSD) + (few != 0) * N.N2_FEW_GCWINDOWS).astype(np.uint8)
    t.is_targets = bool(vf.n) and int(fx["alt"][0]) == 1
    t.filters = fx["filters"]
    return t
