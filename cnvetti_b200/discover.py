"""`cnvetti cmd discover` (builder-defined; the reference only has the CLI stub,
cnvetti/src/cli.yaml:457-469, and bails at cnvetti/src/main.rs:135).

5-state copy-number HMM over FORMAT/CV, one lane per (sample, contig); spec in DESIGN.md "HMM".
The recurrences run on the GPU (cnvb_hmm_segment); forming segment records from the state path
is integer run-length bookkeeping done here.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native as N
from .context import buf, default_context

STATES = 5


@dataclass
class DiscoverOptions:
    sigma0: float = 0.02
    eps_cn: float = 0.1
    p_move: float = 1e-4

    def c_opts(self):
        return N.HmmOpts(self.sigma0, self.eps_cn, self.p_move)


def hmm_segment(cv, lane_off, sigma_s, options=None, posterior=False, ctx=None):
    """Viterbi states (uint8, 0..4) for every bin; optionally forward-backward posteriors [n][5]."""
    ctx = ctx or default_context()
    options = options or DiscoverOptions()
    p, mem, keep = buf(cv, np.float32)
    lane_off = np.ascontiguousarray(lane_off, np.uint64)
    sigma_s = np.ascontiguousarray(sigma_s, np.float64)
    n_lanes = lane_off.size - 1
    assert sigma_s.size == n_lanes
    n = int(keep.shape[0])
    state = ctx.empty(n, np.uint8, mem)
    post = ctx.empty(n * STATES, np.float64, mem) if posterior else None
    o = options.c_opts()
    ctx.check(N.lib().cnvb_hmm_segment(ctx.handle, p, lane_off.ctypes.data, n_lanes, sigma_s.ctypes.data,
                                       C.byref(o), buf(state)[0], buf(post)[0], mem))
    if post is not None:
        post = post.reshape(n, STATES)
    return (state, post) if posterior else state


def segments_from_states(state, lane_off):
    """Maximal runs of equal non-CN2 state -> rows (lane, first_bin, end_bin_exclusive, cn);
    one <DEL>/<DUP> record each (bins are relative to the lane start)."""
    state = np.asarray(state.cpu() if hasattr(state, "cpu") else state)
    out = []
    for lane in range(len(lane_off) - 1):
        s = state[int(lane_off[lane]):int(lane_off[lane + 1])]
        if s.size == 0:
            continue
        cut = np.flatnonzero(np.diff(s)) + 1
        starts = np.concatenate([[0], cut])
        ends = np.concatenate([cut, [s.size]])
        for a, b in zip(starts, ends):
            if s[a] != 2:
                out.append((lane, int(a), int(b), int(s[a])))
    return out
