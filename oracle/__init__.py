"""ctypes front-end of the CPU oracle (oracle/cnv_oracle.c).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package; nothing under cnvetti_b200/ does.

The oracle is a plain-C restatement of the reference's arithmetic (the Rust reference cannot be
built here); each C function cites the reference lines it follows.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
F32_MISSING_BITS = 0x7F800001
N2_CV, N2_CV2, N2_CVSD, N2_FEW_GC = 1, 2, 4, 8
HMM_STATES = 5


def build(force=False):
    """Compile liboracle.so with gcc (Makefile in this directory)."""
    src = os.path.join(_HERE, "cnv_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s", "clean", "all"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


class Reads(C.Structure):
    _fields_ = [("n", C.c_uint64), ("pos", C.c_void_p), ("isize", C.c_void_p),
                ("flag", C.c_void_p), ("mapq", C.c_void_p), ("cigar_off", C.c_void_p),
                ("cigar", C.c_void_p)]


class CovOpts(C.Structure):
    _fields_ = [("min_mapq", C.c_uint8), ("min_unclipped", C.c_float),
                ("window_length", C.c_uint32), ("min_window_remaining", C.c_float),
                ("min_raw_coverage", C.c_uint32)]


class HmmTables(C.Structure):
    _fields_ = [("log_start", C.c_double * 5), ("log_trans", (C.c_double * 5) * 5),
                ("mu", C.c_double * 5), ("inv_sd", C.c_double * 5), ("log_norm", C.c_double * 5),
                ("mu_f", C.c_float * 5), ("inv_sd_f", C.c_float * 5), ("log_norm_sc", C.c_float * 5),
                ("sc_start", C.c_int32), ("sc_stay", C.c_int32), ("sc_move", C.c_int32)]


def _declare(L):
    d = C.c_double
    vp = C.c_void_p
    for name in ("sum", "min", "max", "mean", "median", "var", "std_dev", "std_dev_pct",
                 "median_abs_dev", "median_abs_dev_pct", "iqr"):
        f = getattr(L, "orc_stats_" + name)
        f.restype = d
        f.argtypes = [vp, C.c_size_t]
    L.orc_stats_percentile.restype = d
    L.orc_stats_percentile.argtypes = [vp, C.c_size_t, d]
    L.orc_stats_quartiles.restype = None
    L.orc_stats_quartiles.argtypes = [vp, C.c_size_t, vp]
    L.orc_frag_gw.restype = C.c_int
    L.orc_frag_gw.argtypes = [C.POINTER(Reads), C.POINTER(CovOpts), C.c_uint64, vp, vp,
                              C.c_uint32, vp, vp, vp]
    L.orc_frag_gw_stats.restype = None
    L.orc_frag_gw_stats.argtypes = [vp, C.c_uint64, C.c_uint32, vp, vp, C.c_uint32, vp, vp, vp, vp]
    L.orc_frag_targets.restype = C.c_int
    L.orc_frag_targets.argtypes = [C.POINTER(Reads), C.POINTER(CovOpts), vp, vp, C.c_uint32, vp,
                                   vp, vp]
    L.orc_coverage.restype = C.c_int
    L.orc_coverage.argtypes = [C.POINTER(Reads), C.POINTER(CovOpts), C.c_uint64, C.c_uint32, vp,
                               vp, vp, vp]
    L.orc_refstats.restype = None
    L.orc_refstats.argtypes = [vp, C.c_uint64, C.c_uint32, vp, vp]
    L.orc_cov_records.restype = None
    L.orc_cov_records.argtypes = [C.c_uint64, vp, vp, vp, vp, vp, C.c_int, C.POINTER(CovOpts),
                                  vp, vp, vp]
    L.orc_norm_total.restype = d
    L.orc_norm_total.argtypes = [vp, vp, C.c_uint64, vp, vp]
    L.orc_norm_gc_median.restype = None
    L.orc_norm_gc_median.argtypes = [vp, vp, vp, C.c_uint64, vp, vp, vp, vp, vp]
    L.orc_wis_distances.restype = None
    L.orc_wis_distances.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                    C.c_uint32, vp, vp, C.c_int]
    L.orc_wis_prune.restype = C.c_float
    L.orc_wis_prune.argtypes = [vp, vp, C.c_uint32, C.c_uint32, d, vp]
    L.orc_wis_calls.restype = None
    L.orc_wis_calls.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32, C.c_uint32, d, d,
                                vp]
    L.orc_wis_filters.restype = None
    L.orc_wis_filters.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp]
    L.orc_modcov.restype = None
    L.orc_modcov.argtypes = [vp, C.c_uint32, vp, vp, vp, C.c_uint32, vp, vp, vp, vp, vp, vp]
    L.orc_pile_threshold.restype = C.c_int
    L.orc_pile_threshold.argtypes = [vp, C.c_size_t, d, vp]
    L.orc_piles_from_depths.restype = C.c_uint64
    L.orc_piles_from_depths.argtypes = [vp, C.c_uint64, C.c_uint32, vp, vp, vp]
    L.orc_hmm_make_tables.restype = None
    L.orc_hmm_make_tables.argtypes = [d, d, d, d, C.POINTER(HmmTables)]
    L.orc_hmm_viterbi.restype = None
    L.orc_hmm_viterbi.argtypes = [vp, C.c_uint64, C.POINTER(HmmTables), vp]
    L.orc_hmm_viterbi_f64.restype = None
    L.orc_hmm_viterbi_f64.argtypes = [vp, C.c_uint64, C.POINTER(HmmTables), vp]
    L.orc_hmm_emit_scores.restype = None
    L.orc_hmm_emit_scores.argtypes = [C.c_float, C.POINTER(HmmTables), vp]
    L.orc_hmm_posterior.restype = None
    L.orc_hmm_posterior.argtypes = [vp, C.c_uint64, C.POINTER(HmmTables), vp]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class OraclePanic(RuntimeError):
    """The reference would panic on this input (index out of range, unwrap on None ...)."""


# ---- stats --------------------------------------------------------------------------------
def stats(name, xs, *extra):
    xs = _arr(xs, np.float64)
    if name == "quartiles":
        out = np.zeros(3)
        lib().orc_stats_quartiles(_p(xs), xs.size, _p(out))
        return tuple(out.tolist())
    return getattr(lib(), "orc_stats_" + name)(_p(xs), xs.size, *extra)


# ---- reads --------------------------------------------------------------------------------
class ReadBatch:
    """BAM-level records (what the reference sees through rust-htslib)."""

    def __init__(self, pos, isize, flag, mapq, cigar_off, cigar):
        self.pos = _arr(pos, np.int32)
        self.isize = _arr(isize, np.int32)
        self.flag = _arr(flag, np.uint16)
        self.mapq = _arr(mapq, np.uint8)
        self.cigar_off = _arr(cigar_off, np.uint64)
        self.cigar = _arr(cigar, np.uint32)
        self.n = int(self.pos.size)
        assert self.cigar_off.size == self.n + 1
        self.c = Reads(self.n, _p(self.pos).value, _p(self.isize).value, _p(self.flag).value,
                       _p(self.mapq).value, _p(self.cigar_off).value,
                       _p(self.cigar).value if self.cigar.size else None)


def cov_opts(min_mapq=0, min_unclipped=0.6, window_length=1000, min_window_remaining=0.5,
             min_raw_coverage=10):
    return CovOpts(min_mapq, np.float32(min_unclipped), window_length,
                   np.float32(min_window_remaining), min_raw_coverage)


def frag_gw(reads, opts, region_end, pile_start=None, pile_end=None):
    wl = opts.window_length
    nb = (region_end + wl - 1) // wl
    counters = np.zeros(nb, np.uint32)
    proc = np.zeros(1, np.uint32)
    skip = np.zeros(1, np.uint32)
    ps = _arr(pile_start if pile_start is not None else [], np.uint32)
    pe = _arr(pile_end if pile_end is not None else [], np.uint32)
    rc = lib().orc_frag_gw(C.byref(reads.c), C.byref(opts), region_end, _p(ps), _p(pe), ps.size,
                           _p(counters), _p(proc), _p(skip))
    if rc:
        raise OraclePanic("frag_gw")
    return counters, int(proc[0]), int(skip[0])


def frag_gw_stats(counters, opts, region_end, pile_start=None, pile_end=None):
    nb = counters.size
    ps = _arr(pile_start if pile_start is not None else [], np.uint32)
    pe = _arr(pile_end if pile_end is not None else [], np.uint32)
    cov = np.zeros(nb, np.float32)
    start = np.zeros(nb, np.uint32)
    end = np.zeros(nb, np.uint32)
    frm = np.zeros(nb, np.float32)
    lib().orc_frag_gw_stats(_p(counters), region_end, opts.window_length, _p(ps), _p(pe), ps.size,
                            _p(cov), _p(start), _p(end), _p(frm))
    return cov, start, end, frm


def frag_targets(reads, opts, tstart, tend):
    tstart = _arr(tstart, np.uint32)
    tend = _arr(tend, np.uint32)
    counters = np.zeros(tstart.size, np.uint32)
    proc = np.zeros(1, np.uint32)
    skip = np.zeros(1, np.uint32)
    rc = lib().orc_frag_targets(C.byref(reads.c), C.byref(opts), _p(tstart), _p(tend),
                                tstart.size, _p(counters), _p(proc), _p(skip))
    if rc:
        raise OraclePanic("frag_targets")
    return counters, int(proc[0]), int(skip[0])


def coverage(reads, opts, region_end, plp_mask=0x704, debug=False):
    wl = opts.window_length
    nb = (region_end + wl - 1) // wl
    mean = np.zeros(nb, np.float32)
    sd = np.zeros(nb, np.float32)
    cols = np.zeros(nb * wl, np.uint32) if debug else None
    depth = np.zeros(nb * wl, np.uint32) if debug else None
    rc = lib().orc_coverage(C.byref(reads.c), C.byref(opts), region_end, plp_mask, _p(mean),
                            _p(sd), _p(cols), _p(depth))
    if rc:
        raise OraclePanic("coverage")
    return (mean, sd, cols, depth) if debug else (mean, sd)


def refstats(seq, wl):
    seq = _arr(seq, np.uint8)
    nb = (seq.size + wl - 1) // wl
    gc = np.zeros(nb, np.float32)
    gap = np.zeros(nb, np.uint8)
    lib().orc_refstats(_p(seq), seq.size, wl, _p(gc), _p(gap))
    return gc, gap


def cov_records(cov, cov_sd, start, end, frac_removed, is_targets, opts):
    n = cov.size
    lcv = np.zeros(n, np.float32)
    lcvsd = np.zeros(n, np.float32) if cov_sd is not None else None
    ft = np.zeros(n, np.uint8)
    lib().orc_cov_records(n, _p(_arr(cov, np.float32)),
                          _p(_arr(cov_sd, np.float32)) if cov_sd is not None else None,
                          _p(_arr(start, np.uint32)), _p(_arr(end, np.uint32)),
                          _p(_arr(frac_removed, np.float32)) if frac_removed is not None else None,
                          int(is_targets), C.byref(opts), _p(lcv), _p(lcvsd), _p(ft))
    return lcv, lcvsd, ft


def norm_total(lcv, lcvsd=None):
    lcv = _arr(lcv, np.float32)
    n = lcv.size
    cv = np.zeros(n, np.float32)
    cvsd = np.zeros(n, np.float32) if lcvsd is not None else None
    sd_in = _arr(lcvsd, np.float32) if lcvsd is not None else None
    s = lib().orc_norm_total(_p(lcv), _p(sd_in), n, _p(cv), _p(cvsd))
    return s, cv, cvsd


def norm_gc_median(lcv, gc, lcvsd=None):
    lcv = _arr(lcv, np.float32)
    gc = _arr(gc, np.float32)
    n = lcv.size
    cv = np.zeros(n, np.float32)
    cv2 = np.zeros(n, np.float32)
    cvsd = np.zeros(n, np.float32) if lcvsd is not None else None
    sd_in = _arr(lcvsd, np.float32) if lcvsd is not None else None
    flags = np.zeros(n, np.uint8)
    med = np.zeros(101, np.float32)
    lib().orc_norm_gc_median(_p(lcv), _p(sd_in), _p(gc), n, _p(cv), _p(cv2), _p(cvsd), _p(flags),
                             _p(med))
    return cv, cv2, cvsd, flags, med


# ---- WIS ----------------------------------------------------------------------------------
def wis_distances(X, tids, max_ref_targets=100, row_begin=0, row_end=None, nthreads=1):
    X = _arr(X, np.float32)
    tids = _arr(tids, np.uint32)
    T, S = X.shape
    row_end = T if row_end is None else row_end
    k = min(T, max_ref_targets)
    dist = np.zeros((row_end - row_begin, k), np.float32)
    idx = np.zeros((row_end - row_begin, k), np.uint32)
    lib().orc_wis_distances(_p(X), _p(tids), T, S, max_ref_targets, row_begin, row_end, _p(dist),
                            _p(idx), nthreads)
    return dist, idx


def wis_prune(dist, idx, z=5.64):
    dist = _arr(dist, np.float32)
    idx = _arr(idx, np.uint32).copy()
    T, k = dist.shape
    ref_len = np.zeros(T, np.uint32)
    thr = lib().orc_wis_prune(_p(dist), _p(idx), T, k, z, _p(ref_len))
    return thr, idx, ref_len


def wis_calls(X, ref_idx, ref_len, min_ref_targets=10, z=5.64, rel=0.35):
    X = _arr(X, np.float32)
    T, S = X.shape
    ref_idx = _arr(ref_idx, np.uint32)
    ref_len = _arr(ref_len, np.uint32)
    k = ref_idx.shape[1]
    calls = np.zeros(T, np.uint32)
    lib().orc_wis_calls(_p(X), T, S, _p(ref_idx), _p(ref_len), k, min_ref_targets, z, rel,
                        _p(calls))
    return calls


def wis_filters(ref_len, num_calls, min_ref_targets=10, max_samples_reliable=4):
    ref_len = _arr(ref_len, np.uint32)
    num_calls = _arr(num_calls, np.uint32)
    filt = np.zeros(ref_len.size, np.uint8)
    lib().orc_wis_filters(_p(ref_len), _p(num_calls), ref_len.size, min_ref_targets,
                          max_samples_reliable, _p(filt))
    return filt


def modcov(cv, model_filt, ref_idx, ref_len):
    cv = _arr(cv, np.float32)
    T = cv.size
    model_filt = _arr(model_filt, np.uint8)
    ref_idx = _arr(ref_idx, np.uint32)
    ref_len = _arr(ref_len, np.uint32)
    k = ref_idx.shape[1]
    out = [np.zeros(T, np.float32) for _ in range(5)]
    status = np.zeros(T, np.uint8)
    lib().orc_modcov(_p(cv), T, _p(model_filt), _p(ref_idx), _p(ref_len), k,
                     *[_p(o) for o in out], _p(status))
    return dict(ref_mean=out[0], ref_sd=out[1], cv_rel=out[2], cvz=out[3], cv2=out[4],
                status=status)


# ---- piles --------------------------------------------------------------------------------
def pile_threshold(y, fdr):
    y = _arr(y, np.uint64)
    out = np.zeros(1, np.uint32)
    rc = lib().orc_pile_threshold(_p(y), y.size, fdr, _p(out))
    if rc:
        raise OraclePanic("pile_threshold")
    return int(out[0])


def piles_from_depths(depths, pile_max_gap=20):
    depths = _arr(depths, np.uint32)
    n = depths.size
    ps = np.zeros(n, np.uint32)
    pe = np.zeros(n, np.uint32)
    sz = np.zeros(n, np.uint32)
    m = lib().orc_piles_from_depths(_p(depths), n, pile_max_gap, _p(ps), _p(pe), _p(sz))
    return ps[:m].copy(), pe[:m].copy(), sz[:m].copy()


# ---- HMM ----------------------------------------------------------------------------------
def hmm_tables(sigma_s, sigma0=0.02, eps_cn=0.1, p_move=1e-4):
    t = HmmTables()
    lib().orc_hmm_make_tables(sigma_s, sigma0, eps_cn, p_move, C.byref(t))
    return t


def hmm_viterbi(cv, tables):
    cv = _arr(cv, np.float32)
    state = np.zeros(cv.size, np.uint8)
    lib().orc_hmm_viterbi(_p(cv), cv.size, C.byref(tables), _p(state))
    return state


def hmm_viterbi_f64(cv, tables):
    """The recurrence in log-space FP64 (north_star's wording); the product runs on integer scores -- this is for
    the disagreement rate between the two specs."""
    cv = _arr(cv, np.float32)
    state = np.zeros(cv.size, np.uint8)
    lib().orc_hmm_viterbi_f64(_p(cv), cv.size, C.byref(tables), _p(state))
    return state


def hmm_emit_scores(x, tables):
    e = np.zeros(HMM_STATES, np.int32)
    lib().orc_hmm_emit_scores(float(np.float32(x)), C.byref(tables), _p(e))
    return e


def hmm_posterior(cv, tables):
    cv = _arr(cv, np.float32)
    post = np.zeros((cv.size, HMM_STATES), np.float64)
    lib().orc_hmm_posterior(_p(cv), cv.size, C.byref(tables), _p(post))
    return post
