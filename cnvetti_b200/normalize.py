"""Host-side mirror of lib-normalize (lib-normalize/src/lib.rs:43-66) on in-memory columns."""
import ctypes as C

import numpy as np

from . import _native as N
from .context import buf, default_context


def run_uncorrected_normalization(lcv, lcvsd=None, ctx=None, want_sum=True, out=None, comm=None):
    """Normalization::TotalCovSum (lib-normalize/src/uncorrected.rs:51-96).
    Returns (cv, cvsd or None, lcv_sum).  want_sum=False leaves the sum on the device (no host wait
    for device-resident columns; lcv_sum is None); `out` = preallocated cv.
    comm (parallel.Communicator): the records are sharded over its ranks, the sum runs over all of them
    (one exchange of a double-double pair per rank; collective call)."""
    ctx = (comm.ctx if comm is not None else ctx) or default_context()
    p, mem, keep = buf(lcv, np.float32)
    n = int(keep.shape[0])
    cv = ctx.empty(n, np.float32, mem) if out is None else out
    cvsd = ctx.empty(n, np.float32, mem) if lcvsd is not None else None
    s = C.c_double(0.0)
    if comm is not None:
        ctx.check(N.lib().cnvb_norm_total_sharded(comm.handle, p, buf(lcvsd, np.float32)[0], n, buf(cv)[0],
                                                  buf(cvsd)[0], C.addressof(s) if want_sum else None, mem))
    else:
        ctx.check(N.lib().cnvb_norm_total(ctx.handle, p, buf(lcvsd, np.float32)[0], n, buf(cv)[0],
                                          buf(cvsd)[0], C.addressof(s) if want_sum else None, mem))
    return cv, cvsd, (s.value if want_sum else None)


def run_binned_correction(lcv, gc, lcvsd=None, ctx=None, comm=None, want_medians=True):
    """Normalization::MedianGcBinned (lib-normalize/src/correct_binned.rs:77-146).
    Returns dict(cv, cv2, cvsd, flags, medians); flags bits N2_* say which fields are written.
    comm (parallel.Communicator): the records are sharded over its ranks, the 101 medians run over all of
    them (the radix histograms of each pass are summed over the ranks; collective call).
    want_medians=False skips the read-back of the medians (and with it the host wait)."""
    ctx = (comm.ctx if comm is not None else ctx) or default_context()
    p, mem, keep = buf(lcv, np.float32)
    n = int(keep.shape[0])
    cv = ctx.empty(n, np.float32, mem)
    cv2 = ctx.empty(n, np.float32, mem)
    cvsd = ctx.empty(n, np.float32, mem) if lcvsd is not None else None
    flags = ctx.empty(n, np.uint8, mem)
    med = np.empty(101, np.float32) if want_medians else None
    args = (p, buf(lcvsd, np.float32)[0], buf(gc, np.float32)[0], n, buf(cv)[0], buf(cv2)[0], buf(cvsd)[0],
            buf(flags)[0], med.ctypes.data if want_medians else None, mem)
    if comm is not None:
        ctx.check(N.lib().cnvb_norm_gc_median_sharded(comm.handle, *args))
    else:
        ctx.check(N.lib().cnvb_norm_gc_median(ctx.handle, *args))
    return dict(cv=cv, cv2=cv2, cvsd=cvsd, flags=flags, medians=med)
