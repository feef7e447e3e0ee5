"""ctypes binding of libcnvetti_b200.so (the C ABI in include/cnvetti_b200.h).

There is no fallback: if the shared library is missing this module raises, and if no CUDA device
is present every compute call fails with CNVB_ERR_CUDA.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcnvetti_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_REFERENCE_PANIC, ERR_STATE, ERR_NOMEM, ERR_COMM = 0, -1, -2, -3, -4, -5, -6
MEM_HOST, MEM_DEVICE = 0, 1
FLAG_CLIPFAIL, FLAG_ISIZE_NONPOS = 0x1000, 0x2000
FRAGMENTS_GENOME_WIDE, FRAGMENTS_TARGET_REGIONS, COVERAGE = 0, 1, 2
FT_LOWCOV, FT_PILE = 1, 2
N2_CV, N2_CV2, N2_CVSD, N2_FEW_GCWINDOWS = 1, 2, 4, 8
F32_MISSING_BITS = 0x7F800001


class ReadSoa(C.Structure):
    _fields_ = [("n", C.c_uint64), ("tid", C.c_void_p), ("pos", C.c_void_p), ("end", C.c_void_p),
                ("mid", C.c_void_p), ("frag_end", C.c_void_p), ("mapq", C.c_void_p),
                ("flags", C.c_void_p), ("mem", C.c_int)]


class CovOpts(C.Structure):
    _fields_ = [("min_mapq", C.c_uint8), ("min_unclipped", C.c_float),
                ("window_length", C.c_uint32), ("min_window_remaining", C.c_float),
                ("min_raw_coverage", C.c_uint32), ("plp_flag_mask", C.c_uint32)]


class Region(C.Structure):
    _fields_ = [("region_end", C.c_uint32), ("reads", ReadSoa), ("seq", C.c_void_p),
                ("seq_len", C.c_uint64), ("seq_mem", C.c_int),
                ("tgt_start", C.c_void_p), ("tgt_end", C.c_void_p), ("n_tgt", C.c_uint32),
                ("pile_start", C.c_void_p), ("pile_end", C.c_void_p), ("n_pile", C.c_uint32)]


class CovColumns(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("start", "end", "rcv", "rcvsd", "lcv", "lcvsd", "frm", "gc",
                                          "mean_mapq", "gap", "ft")]


class VfileFmt(C.Structure):
    _fields_ = [("key", C.c_char_p), ("values", C.c_void_p), ("present", C.c_void_p)]


class VfileWriteArgs(C.Structure):
    _fields_ = [("path", C.c_char_p), ("header_text", C.c_char_p), ("n_samples", C.c_uint32),
                ("samples", C.POINTER(C.c_char_p)), ("n_rows", C.c_uint64), ("rid", C.c_void_p), ("pos", C.c_void_p),
                ("end", C.c_void_p), ("is_targets", C.c_int), ("filters", C.c_void_p), ("gap", C.c_void_p),
                ("gc", C.c_void_p), ("mean_mapq", C.c_void_p), ("few_gc", C.c_void_p), ("ref_mean", C.c_void_p),
                ("ref_sd", C.c_void_p), ("ref_present", C.c_void_p), ("num_calls", C.c_void_p),
                ("ref_off", C.c_void_p), ("ref_idx", C.c_void_p), ("n_fmt", C.c_uint32),
                ("fmt", C.POINTER(VfileFmt)), ("ft", C.c_void_p), ("ft_pos", C.c_uint32), ("write_index", C.c_int),
                ("n_threads", C.c_int)]


VF_PASS, VF_PILE, VF_LOWCOV, VF_NO_REF, VF_FEW_REF, VF_UNRELIABLE, VF_OTHER = 1, 2, 4, 8, 16, 32, 128


class WisOpts(C.Structure):
    _fields_ = [("filter_z_score", C.c_double), ("filter_rel", C.c_double),
                ("min_ref_targets", C.c_uint32), ("max_ref_targets", C.c_uint32),
                ("max_samples_reliable", C.c_uint32), ("engine", C.c_uint32)]


class HmmOpts(C.Structure):
    _fields_ = [("sigma0", C.c_double), ("eps_cn", C.c_double), ("p_move", C.c_double)]


# every symbol include/cnvetti_b200.h declares: name -> (restype, argtypes)
_vp, _u64, _u32, _int = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
SYMBOLS = {
    "cnvb_ctx_create": (_int, [_int, _vp, C.POINTER(_vp)]),
    "cnvb_ctx_destroy": (None, [_vp]),
    "cnvb_ctx_synchronize": (_int, [_vp]),
    "cnvb_last_error": (C.c_char_p, [_vp]),
    "cnvb_create_error": (C.c_char_p, []),
    "cnvb_ctx_launch_count": (_u64, [_vp]),
    "cnvb_version": (C.c_char_p, []),
    "cnvb_ctx_release_memory": (_int, [_vp]),
    "cnvb_ctx_set_timing": (_int, [_vp, _int]),
    "cnvb_ctx_timing_reset": (_int, [_vp]),
    "cnvb_ctx_timing_query": (_int, [_vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(_u64)]),
    "cnvb_pack_reads": (_int, [_u64, _vp, _vp, _vp, _vp, _vp, C.c_float, _vp, _vp, _vp, _vp]),
    "cnvb_bam_open": (_int, [C.c_char_p, _int, C.POINTER(_vp)]),
    "cnvb_bam_error": (C.c_char_p, [_vp]),
    "cnvb_bam_n_refs": (_u32, [_vp]),
    "cnvb_bam_ref_name": (C.c_char_p, [_vp, _u32]),
    "cnvb_bam_ref_len": (_u32, [_vp, _u32]),
    "cnvb_bam_header_text": (_vp, [_vp, C.POINTER(_u64)]),
    "cnvb_bam_decode": (_int, [_vp, C.c_float, _int]),
    "cnvb_bam_fetch": (_int, [_vp, _u32, _u32, _u32, C.c_float, _int]),
    "cnvb_bam_ref_reads": (_int, [_vp, _u32, C.POINTER(ReadSoa)]),
    "cnvb_bam_stats": (_int, [_vp, C.POINTER(_u64), C.POINTER(C.c_double)]),
    "cnvb_bam_close": (None, [_vp]),
    "cnvb_vfile_write": (_int, [C.POINTER(VfileWriteArgs), C.c_char_p, _u64]),
    "cnvb_vfile_open": (_int, [C.c_char_p, _int, C.POINTER(_vp), C.c_char_p, _u64]),
    "cnvb_vfile_close": (None, [_vp]),
    "cnvb_vfile_last_error": (C.c_char_p, [_vp]),
    "cnvb_vfile_num_records": (_u64, [_vp]),
    "cnvb_vfile_num_samples": (_u32, [_vp]),
    "cnvb_vfile_sample": (C.c_char_p, [_vp, _u32]),
    "cnvb_vfile_num_contigs": (_u32, [_vp]),
    "cnvb_vfile_contig_name": (C.c_char_p, [_vp, _u32]),
    "cnvb_vfile_contig_length": (_u32, [_vp, _u32]),
    "cnvb_vfile_header_text": (C.c_char_p, [_vp]),
    "cnvb_vfile_is_bcf": (_int, [_vp]),
    "cnvb_vfile_fixed": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_int)]),
    "cnvb_vfile_info_float": (_int, [_vp, C.c_char_p, _vp, _vp]),
    "cnvb_vfile_info_int": (_int, [_vp, C.c_char_p, _vp, _vp]),
    "cnvb_vfile_info_flag": (_int, [_vp, C.c_char_p, _vp]),
    "cnvb_vfile_format_float": (_int, [_vp, C.c_char_p, _vp, _vp]),
    "cnvb_vfile_format_ft": (_int, [_vp, _vp, _vp]),
    "cnvb_vfile_match_ids": (_int, [_vp, _vp, _vp]),
    "cnvb_vfile_ref_targets": (_int, [_vp, _vp, _vp, _vp, _u64]),
    "cnvb_cov_create": (_int, [_vp, C.POINTER(CovOpts), _int, _u32, _vp, _vp, _u32, _vp, _vp,
                               _u32, C.POINTER(_vp)]),
    "cnvb_cov_put_records": (_int, [_vp, C.POINTER(ReadSoa)]),
    "cnvb_cov_finish": (_int, [_vp]),
    "cnvb_cov_finish_async": (_int, [_vp]),
    "cnvb_cov_num_regions": (_u64, [_vp]),
    "cnvb_cov_num_processed": (_u32, [_vp]),
    "cnvb_cov_num_skipped": (_u32, [_vp]),
    "cnvb_cov_get_stats": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _int]),
    "cnvb_cov_get_counts": (_int, [_vp, _vp, _int]),
    "cnvb_cov_destroy": (None, [_vp]),
    "cnvb_refstats": (_int, [_vp, _vp, _u64, _u32, _vp, _vp, _int]),
    "cnvb_cov_records": (_int, [_vp, _u64, _vp, _vp, _vp, _vp, _vp, _int, C.POINTER(CovOpts),
                                _vp, _vp, _vp, _int]),
    "cnvb_coverage_rows": (_u64, [C.POINTER(CovOpts), _int, C.POINTER(Region), _u32]),
    "cnvb_coverage_run": (_int, [_vp, C.POINTER(CovOpts), _int, C.POINTER(Region), _u32,
                                 C.POINTER(CovColumns), _vp, _vp, _int]),
    "cnvb_coverage_run_async": (_int, [_vp, C.POINTER(CovOpts), _int, C.POINTER(Region), _u32,
                                       C.POINTER(CovColumns), C.POINTER(_vp)]),
    "cnvb_coverage_wait": (_int, [_vp, _vp, _vp]),
    "cnvb_coverage_plan_create": (_int, [_vp, C.POINTER(CovOpts), _int, C.POINTER(Region), _u32,
                                         C.POINTER(CovColumns), C.POINTER(_vp)]),
    "cnvb_coverage_plan_launch": (_int, [_vp]),
    "cnvb_coverage_plan_wait": (_int, [_vp, _vp, _vp]),
    "cnvb_coverage_plan_destroy": (None, [_vp]),
    "cnvb_piles_collect": (_int, [_vp, C.POINTER(ReadSoa), C.POINTER(CovOpts), _u32, _u32, C.POINTER(_vp)]),
    "cnvb_piles_count": (_u64, [_vp]),
    "cnvb_piles_get": (_int, [_vp, _vp, _vp, _vp, _int]),
    "cnvb_piles_destroy": (None, [_vp]),
    "cnvb_pile_threshold": (_int, [_vp, _u64, C.c_double, C.POINTER(_u32), C.c_char_p, _u64]),
    "cnvb_norm_total": (_int, [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _int]),
    "cnvb_norm_gc_median": (_int, [_vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _int]),
    "cnvb_wis_distances": (_int, [_vp, _vp, _vp, _u32, _u32, C.POINTER(WisOpts), _u32, _u32, _vp, _vp, _int]),
    "cnvb_wis_last_stats": (_int, [_vp, _vp]),
    "cnvb_wis_threshold": (_int, [_vp, _vp, _u32, _u32, C.c_double, C.POINTER(C.c_float), _int]),
    "cnvb_wis_prune": (_int, [_vp, _vp, _vp, _u32, _u32, C.c_float, _vp, _int]),
    "cnvb_wis_calls": (_int, [_vp, _vp, _u32, _u32, _vp, _vp, _u32, C.POINTER(WisOpts), _u32, _u32, _vp, _int]),
    "cnvb_wis_filters": (_int, [_vp, _vp, _vp, _u32, C.POINTER(WisOpts), _vp, _int]),
    "cnvb_wis_build": (_int, [_vp, _vp, _vp, _u32, _u32, C.POINTER(WisOpts), _vp, _vp, _vp, _vp, _vp,
                              C.POINTER(C.c_float), _int]),
    "cnvb_comm_unique_id": (_int, [_vp]),
    "cnvb_comm_init_rank": (_int, [_vp, _vp, _int, _int, C.POINTER(_vp)]),
    "cnvb_comm_from_nccl": (_int, [_vp, _vp, _int, _int, C.POINTER(_vp)]),
    "cnvb_comm_destroy": (None, [_vp]),
    "cnvb_comm_rank": (_int, [_vp]),
    "cnvb_comm_world": (_int, [_vp]),
    "cnvb_comm_nccl_version": (_int, [C.POINTER(_int)]),
    "cnvb_comm_create_error": (C.c_char_p, []),
    "cnvb_comm_all_gather_v": (_int, [_vp, _vp, _vp, _vp]),
    "cnvb_comm_all_to_all_v": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "cnvb_comm_all_reduce_u32": (_int, [_vp, _vp, _u64]),
    "cnvb_norm_total_sharded": (_int, [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _int]),
    "cnvb_norm_gc_median_sharded": (_int, [_vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _int]),
    "cnvb_wis_build_sharded": (_int, [_vp, _vp, _vp, _u32, _u32, C.POINTER(WisOpts), _vp, _vp, _vp, _vp, _vp,
                                      C.POINTER(C.c_float), C.POINTER(_u32), C.POINTER(_u32), _int]),
    "cnvb_merge_cov": (_int, [_vp, _vp, _vp, _u64, _u32, _u64, _vp]),
    "cnvb_hmm_segment": (_int, [_vp, _vp, _vp, _u32, _vp, C.POINTER(HmmOpts), _vp, _vp, _int]),
    "cnvb_modcov": (_int, [_vp, _vp, _u32, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _int]),
    "cnvb_modcov_panel": (_int, [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _int]),
    "cnvb_wis_calls_modcov_panel": (_int, [_vp, _vp, _u32, _u32, _vp, _vp, _u32, _vp, _u32, _u32, _vp, _vp, _vp, _vp, _vp,
                                           _vp, _vp, _vp, _int]),
}

_lib = None


def lib():
    """Load the shared library (building it is `python -m cnvetti_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "cnvetti_b200: %s is missing. Build it with `python -m cnvetti_b200.build`; "
                "there is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


class CnvbError(RuntimeError):
    def __init__(self, code, chain):
        self.code = code
        self.chain = chain
        RuntimeError.__init__(self, "error: %s" % chain)


class ReferencePanic(CnvbError):
    """Input on which the reference itself panics (CNVB_ERR_REFERENCE_PANIC)."""
