"""Host-side mirror of lib-model-wis and lib-mod-cov on in-memory panels.

  BuildModelWisOptions   lib-model-wis/src/options.rs:5-30 (defaults cnvetti/src/cli.yaml:370-406)
  compute_distances / prune_distances / count_per_probe_calls / run   lib-model-wis/src/lib.rs
  mod_coverage           lib-mod-cov/src/lib.rs:76-164, 209-240
All arithmetic runs on the GPU through the C ABI.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native as N
from .context import buf, default_context, same_mem

ENGINE_AUTO, ENGINE_TENSOR, ENGINE_EXACT = 0, 1, 2


@dataclass
class BuildModelWisOptions:
    filter_z_score: float = 5.64
    filter_rel: float = 0.35
    min_ref_targets: int = 10
    max_ref_targets: int = 100
    max_samples_reliable: int = 4
    engine: int = ENGINE_AUTO

    @classmethod
    def quick(cls, **kw):
        """The preset `cnvetti quick wis-build-model` hard-codes (cnvetti/src/quick_wis_build_model.rs:
        88-105): as the `cmd build-model-wis` defaults (cli.yaml:370-406) except max_samples_reliable = 8."""
        kw.setdefault("max_samples_reliable", 8)
        return cls(**kw)

    def c_opts(self):
        return N.WisOpts(self.filter_z_score, self.filter_rel, self.min_ref_targets, self.max_ref_targets,
                         self.max_samples_reliable, self.engine)


def _panel(X, tids):
    px, mx, kx = buf(X, np.float32)
    pt, mt, kt = buf(tids, None if hasattr(tids, "is_cuda") else np.uint32)
    T, S = int(kx.shape[0]), int(kx.shape[1])
    return px, pt, same_mem(mx, mt), T, S, (kx, kt)


def _empty2(ctx, rows, k, dtype, mem):
    a = ctx.empty(rows * k, dtype, mem)
    return a.reshape(rows, k)


def compute_distances(X, tids, options=None, row_begin=0, row_end=None, ctx=None):
    """lib.rs:128-197 -> (dist[rows][k] f32, idx[rows][k])."""
    ctx = ctx or default_context()
    options = options or BuildModelWisOptions()
    px, pt, mem, T, S, keep = _panel(X, tids)
    row_end = T if row_end is None else row_end
    k = min(T, options.max_ref_targets)
    rows = row_end - row_begin
    dist = _empty2(ctx, rows, k, np.float32, mem)
    idx = _empty2(ctx, rows, k, np.int32 if mem == N.MEM_DEVICE else np.uint32, mem)
    o = options.c_opts()
    ctx.check(N.lib().cnvb_wis_distances(ctx.handle, px, pt, T, S, C.byref(o), row_begin, row_end,
                                         buf(dist)[0], buf(idx)[0], mem))
    return dist, idx


def last_stats(ctx=None):
    """dict(rows, fallback_rows, rescored_pairs, tiles) of the last distance computation
    (tiles = 256x128 tiles of the pair matrix the tensor sweep visited)."""
    ctx = ctx or default_context()
    out = (C.c_uint64 * 4)()
    ctx.check(N.lib().cnvb_wis_last_stats(ctx.handle, out))
    return dict(rows=int(out[0]), fallback_rows=int(out[1]), rescored_pairs=int(out[2]), tiles=int(out[3]))


def prune_threshold(nearest, z, stride=1, ctx=None):
    """lib.rs:207-216: (mean + z*sd) as f32 of the nearest distances of all targets."""
    ctx = ctx or default_context()
    p, mem, keep = buf(nearest, np.float32)
    T = int(keep.shape[0]) // stride
    thr = C.c_float(0.0)
    ctx.check(N.lib().cnvb_wis_threshold(ctx.handle, p, T, stride, z, C.byref(thr), mem))
    return thr.value


def prune_distances(dist, idx, threshold, ctx=None):
    """lib.rs:221-228 -> (idx compacted in place, ref_len)."""
    ctx = ctx or default_context()
    pd, md, kd = buf(dist, np.float32)
    pi, mi, ki = buf(idx, None if hasattr(idx, "is_cuda") else np.uint32)
    mem = same_mem(md, mi)
    rows, k = int(kd.shape[0]), int(kd.shape[1])
    ref_len = ctx.empty(rows, np.int32 if mem == N.MEM_DEVICE else np.uint32, mem)
    ctx.check(N.lib().cnvb_wis_prune(ctx.handle, pd, pi, rows, k, C.c_float(threshold), buf(ref_len)[0], mem))
    return ki, ref_len


def count_per_probe_calls(X, ref_idx, ref_len, options=None, row_begin=0, row_end=None, ctx=None):
    """lib.rs:235-271."""
    ctx = ctx or default_context()
    options = options or BuildModelWisOptions()
    px, mx, kx = buf(X, np.float32)
    T, S = int(kx.shape[0]), int(kx.shape[1])
    row_end = T if row_end is None else row_end
    pi, mi, ki = buf(ref_idx, None if hasattr(ref_idx, "is_cuda") else np.uint32)
    pl, ml, kl = buf(ref_len, None if hasattr(ref_len, "is_cuda") else np.uint32)
    mem = same_mem(mx, mi, ml)
    k = int(ki.shape[1])
    calls = ctx.empty(row_end - row_begin, np.int32 if mem == N.MEM_DEVICE else np.uint32, mem)
    o = options.c_opts()
    ctx.check(N.lib().cnvb_wis_calls(ctx.handle, px, T, S, pi, pl, k, C.byref(o), row_begin, row_end,
                                     buf(calls)[0], mem))
    return calls


def model_filters(ref_len, num_calls, options=None, ctx=None):
    """FILTER bits of write_output (lib.rs:362-377): 1 FEW_REF, 2 UNRELIABLE."""
    ctx = ctx or default_context()
    options = options or BuildModelWisOptions()
    pl, ml, kl = buf(ref_len, None if hasattr(ref_len, "is_cuda") else np.uint32)
    pc, mc, kc = buf(num_calls, None if hasattr(num_calls, "is_cuda") else np.uint32)
    mem = same_mem(ml, mc)
    filt = ctx.empty(int(kl.shape[0]), np.uint8, mem)
    o = options.c_opts()
    ctx.check(N.lib().cnvb_wis_filters(ctx.handle, pl, pc, int(kl.shape[0]), C.byref(o), buf(filt)[0], mem))
    return filt


@dataclass
class WisModel:
    """What `cmd build-model-wis` writes per target (INFO REF_TARGETS order = ref_idx order)."""
    ref_dist: object
    ref_idx: object
    ref_len: object
    num_calls: object
    filt: object
    threshold: float


def run(X, tids, options=None, ctx=None):
    """lib_model_wis::run (lib.rs:422-447) on an in-memory panel (single GPU)."""
    ctx = ctx or default_context()
    options = options or BuildModelWisOptions()
    px, pt, mem, T, S, keep = _panel(X, tids)
    k = min(T, options.max_ref_targets)
    idt = np.int32 if mem == N.MEM_DEVICE else np.uint32
    dist = _empty2(ctx, T, k, np.float32, mem)
    idx = _empty2(ctx, T, k, idt, mem)
    ref_len = ctx.empty(T, idt, mem)
    calls = ctx.empty(T, idt, mem)
    filt = ctx.empty(T, np.uint8, mem)
    thr = C.c_float(0.0)
    o = options.c_opts()
    ctx.check(N.lib().cnvb_wis_build(ctx.handle, px, pt, T, S, C.byref(o), buf(dist)[0], buf(idx)[0],
                                     buf(ref_len)[0], buf(calls)[0], buf(filt)[0], C.byref(thr), mem))
    return WisModel(dist, idx, ref_len, calls, filt, thr.value)


def run_sharded(X_rows, tid_rows, options, comm):
    """lib_model_wis::run with the target rows sharded over the ranks of `comm` (cnvb_wis_build_sharded):
    every rank passes its rows of the panel (row blocks in rank order) and gets the model of its rows,
    ref_idx in global target indices.  Returns (WisModel, (row_begin, row_end), T).  Collective call."""
    ctx = comm.ctx
    options = options or BuildModelWisOptions()
    px, pt, mem, rows, S, keep = _panel(X_rows, tid_rows)
    # k = min(T, max_ref_targets) is known only after the row counts have been exchanged; panels with fewer
    # targets than max_ref_targets are not what one shards
    k = options.max_ref_targets
    idt = np.int32 if mem == N.MEM_DEVICE else np.uint32
    dist = _empty2(ctx, rows, k, np.float32, mem)
    idx = _empty2(ctx, rows, k, idt, mem)
    ref_len = ctx.empty(rows, idt, mem)
    calls = ctx.empty(rows, idt, mem)
    filt = ctx.empty(rows, np.uint8, mem)
    thr = C.c_float(0.0)
    lo, T = C.c_uint32(0), C.c_uint32(0)
    o = options.c_opts()
    ctx.check(N.lib().cnvb_wis_build_sharded(comm.handle, px, pt, rows, S, C.byref(o), buf(dist)[0], buf(idx)[0],
                                             buf(ref_len)[0], buf(calls)[0], buf(filt)[0], C.byref(thr),
                                             C.byref(lo), C.byref(T), mem))
    if T.value < k:
        raise N.CnvbError(N.ERR_INVALID, "run_sharded: fewer targets (%d) than max_ref_targets (%d)" % (T.value, k))
    return WisModel(dist, idx, ref_len, calls, filt, thr.value), (lo.value, lo.value + rows), T.value


def mod_coverage(cv, model_filt, ref_idx, ref_len, ctx=None):
    """lib-mod-cov: per-target REF_MEAN / REF_STD_DEV / CV / CVZ / CV2 for one sample."""
    ctx = ctx or default_context()
    pc, mc, kc = buf(cv, np.float32)
    pf, mf, kf = buf(model_filt, np.uint8)
    pi, mi, ki = buf(ref_idx, None if hasattr(ref_idx, "is_cuda") else np.uint32)
    pl, ml, kl = buf(ref_len, None if hasattr(ref_len, "is_cuda") else np.uint32)
    mem = same_mem(mc, mf, mi, ml)
    T, k = int(kc.shape[0]), int(ki.shape[1])
    outs = [ctx.empty(T, np.float32, mem) for _ in range(5)]
    status = ctx.empty(T, np.uint8, mem)
    ctx.check(N.lib().cnvb_modcov(ctx.handle, pc, T, pf, pi, pl, k, *[buf(o)[0] for o in outs],
                                  buf(status)[0], mem))
    return dict(ref_mean=outs[0], ref_sd=outs[1], cv_rel=outs[2], cvz=outs[3], cv2=outs[4], status=status)


def mod_coverage_panel(X, model_filt, ref_idx, ref_len, want=("cv_rel", "cvz"), row_begin=0, row_end=None,
                       ctx=None):
    """lib-mod-cov for S samples at once: X[T][S] = FORMAT/CV of the samples side by side, target
    rows [row_begin, row_end) (model_filt / ref_idx / ref_len relative to row_begin).  Returns
    sample-major [S][rows] arrays for the names in `want` (ref_mean, ref_sd, cv_rel, cvz, cv2,
    status); row s equals mod_coverage(X[:, s], ...)[row_begin:row_end]."""
    ctx = ctx or default_context()
    px, mx, kx = buf(X, np.float32)
    pf, mf, kf = buf(model_filt, np.uint8)
    pi, mi, ki = buf(ref_idx, None if hasattr(ref_idx, "is_cuda") else np.uint32)
    pl, ml, kl = buf(ref_len, None if hasattr(ref_len, "is_cuda") else np.uint32)
    mem = same_mem(mx, mf, mi, ml)
    T, S, k = int(kx.shape[0]), int(kx.shape[1]), int(ki.shape[1])
    row_end = T if row_end is None else row_end
    rows = row_end - row_begin
    names = ("ref_mean", "ref_sd", "cv_rel", "cvz", "cv2", "status")
    unknown = set(want) - set(names)
    if unknown:
        raise ValueError("unknown outputs: %s" % sorted(unknown))
    outs = {n: (ctx.empty(S * rows, np.uint8 if n == "status" else np.float32, mem) if n in want else None)
            for n in names}
    ctx.check(N.lib().cnvb_modcov_panel(ctx.handle, px, T, S, row_begin, row_end, pf, pi, pl, k,
                                        *[buf(outs[n])[0] for n in names], mem))
    return {n: v.reshape(S, rows) for n, v in outs.items() if v is not None}


def calls_and_mod_coverage_panel(X, ref_idx, ref_len, options=None, want=("cv_rel", "cvz"), row_begin=0, row_end=None,
                                 ctx=None):
    """count_per_probe_calls + model filters + mod_coverage_panel in one pass over the panel (calling a cohort
    against the model built from it).  Returns (num_calls, filt, {name: [S][rows] array})."""
    ctx = ctx or default_context()
    options = options or BuildModelWisOptions()
    px, mx, kx = buf(X, np.float32)
    pi, mi, ki = buf(ref_idx, None if hasattr(ref_idx, "is_cuda") else np.uint32)
    pl, ml, kl = buf(ref_len, None if hasattr(ref_len, "is_cuda") else np.uint32)
    mem = same_mem(mx, mi, ml)
    T, S, k = int(kx.shape[0]), int(kx.shape[1]), int(ki.shape[1])
    row_end = T if row_end is None else row_end
    rows = row_end - row_begin
    names = ("ref_mean", "ref_sd", "cv_rel", "cvz", "cv2", "status")
    unknown = set(want) - set(names)
    if unknown:
        raise ValueError("unknown outputs: %s" % sorted(unknown))
    calls = ctx.empty(rows, np.int32 if mem == N.MEM_DEVICE else np.uint32, mem)
    filt = ctx.empty(rows, np.uint8, mem)
    outs = {n: (ctx.empty(S * rows, np.uint8 if n == "status" else np.float32, mem) if n in want else None)
            for n in names}
    o = options.c_opts()
    ctx.check(N.lib().cnvb_wis_calls_modcov_panel(ctx.handle, px, T, S, pi, pl, k, C.byref(o), row_begin, row_end,
                                                  buf(calls)[0], buf(filt)[0], *[buf(outs[n])[0] for n in names], mem))
    return calls, filt, {n: v.reshape(S, rows) for n, v in outs.items() if v is not None}
