"""Device context and buffer plumbing for the C ABI (torch is used for device memory only)."""
import ctypes as C

import numpy as np

from . import _native as N

try:  # torch is plumbing: device memory, streams, torch.distributed
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def buf(x, dtype=None):
    """-> (address, mem, keepalive) for a numpy array (host) or torch tensor (host or cuda)."""
    if x is None:
        return None, None, None
    if _is_torch(x):
        if dtype is not None:
            want = getattr(torch, np.dtype(dtype).name)
            if x.dtype != want:
                raise TypeError("expected %s tensor, got %s" % (want, x.dtype))
        if not x.is_contiguous():
            raise ValueError("tensors passed to cnvetti_b200 must be contiguous")
        mem = N.MEM_DEVICE if x.is_cuda else N.MEM_HOST
        return x.data_ptr(), mem, x
    a = np.ascontiguousarray(x, dtype=dtype)
    return a.ctypes.data, N.MEM_HOST, a


def same_mem(*mems):
    ms = {m for m in mems if m is not None}
    if len(ms) > 1:
        raise ValueError("all buffers of one call must live in the same memory space")
    return ms.pop() if ms else N.MEM_HOST


class Context:
    """One CUDA device + one stream (cnvb_ctx).  Not thread-safe; make one per host thread."""

    def __init__(self, device=0, stream=None):
        L = N.lib()
        h = C.c_void_p()
        if stream is None and torch is not None and torch.cuda.is_available():
            # run on torch's current stream so that torch copies/allocations and our kernels are
            # ordered; handle 0 is the legacy default stream, spelled cudaStreamLegacy (0x1)
            stream = torch.cuda.current_stream(device).cuda_stream or 1
        rc = L.cnvb_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != N.OK:
            raise N.CnvbError(rc, L.cnvb_create_error().decode())
        self.handle = h
        self.device = device

    def check(self, rc):
        if rc == N.OK:
            return
        chain = N.lib().cnvb_last_error(self.handle).decode()
        if rc == N.ERR_REFERENCE_PANIC:
            raise N.ReferencePanic(rc, chain)
        raise N.CnvbError(rc, chain)

    def synchronize(self):
        self.check(N.lib().cnvb_ctx_synchronize(self.handle))

    def release_memory(self):
        """hand the ctx's cached device memory (pool, scratch, staging) back to the driver"""
        self.check(N.lib().cnvb_ctx_release_memory(self.handle))

    @property
    def launch_count(self):
        return int(N.lib().cnvb_ctx_launch_count(self.handle))

    def set_timing(self, on=True):
        self.check(N.lib().cnvb_ctx_set_timing(self.handle, int(on)))

    def timing_reset(self):
        self.check(N.lib().cnvb_ctx_timing_reset(self.handle))

    def timing_query(self, kernel):
        """(total device ms, number of launches) of `kernel` since the last reset."""
        ms = C.c_double(0.0)
        cnt = C.c_uint64(0)
        self.check(N.lib().cnvb_ctx_timing_query(self.handle, kernel.encode(), C.byref(ms), C.byref(cnt)))
        return ms.value, int(cnt.value)

    def close(self):
        if getattr(self, "handle", None):
            N.lib().cnvb_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # device / host allocation helpers for outputs ------------------------------------------
    def empty(self, n, dtype, mem):
        if mem == N.MEM_DEVICE:
            return torch.empty(int(n), dtype=getattr(torch, np.dtype(dtype).name),
                               device="cuda:%d" % self.device)
        return np.empty(int(n), dtype=dtype)


_default = {}


def default_context(device=0):
    ctx = _default.get(device)
    if ctx is None:
        ctx = _default[device] = Context(device)
    return ctx
