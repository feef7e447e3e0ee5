"""Parity of the CUDA HMM segmentation against the repo's CPU oracle (builder-defined spec: the
reference has no `cmd discover`).  State paths bit-exact; posteriors within rel 1e-6."""
import numpy as np
import pytest

import cnvetti_b200 as cb
from cnvetti_b200 import discover, synth
from util import to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return cb.Context(0)


def _lanes(seed, lengths):
    cvs, cns = [], []
    for i, n in enumerate(lengths):
        cv, cn = synth.synth_cv_lanes(seed * 100 + i, n, n_cnv=max(1, n // 500), missing_frac=0.002)
        cvs.append(cv)
        cns.append(cn)
    off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    return np.concatenate(cvs) if cvs else np.zeros(0, np.float32), np.concatenate(cns) if cns else np.zeros(0), off


@pytest.mark.parametrize("device", [False, True])
def test_viterbi_paths_bit_exact(oracle, ctx, device):
    lengths = [5000, 1, 0, 77, 12345, 3000, 2, 40000] + [900] * 40
    cv, cn, off = _lanes(4, lengths)
    rng = np.random.default_rng(0)
    sig = rng.uniform(0.05, 0.15, len(lengths))
    opts = discover.DiscoverOptions()
    x = cv
    if device:
        import torch
        x = torch.from_numpy(cv).cuda()
    st = to_np(discover.hmm_segment(x, off, sig, opts, ctx=ctx))
    for l, n in enumerate(lengths):
        t = oracle.hmm_tables(float(sig[l]), opts.sigma0, opts.eps_cn, opts.p_move)
        ref = oracle.hmm_viterbi(cv[int(off[l]):int(off[l + 1])], t)
        assert np.array_equal(st[int(off[l]):int(off[l + 1])], ref), l
    # the planted CNVs are found (sanity of the model, not parity)
    l = 4
    seg = st[int(off[l]):int(off[l + 1])]
    truth = cn[int(off[l]):int(off[l + 1])]
    assert (seg == truth).mean() > 0.97
    segs = discover.segments_from_states(st, off)
    assert len(segs) > 10 and all(s[3] != 2 for s in segs)


@pytest.mark.parametrize("p_move", [1e-4, 0.2, 1e-30, 0.05])
def test_viterbi_chunk_edges_and_ties(oracle, ctx, p_move):
    """Lanes around the 64-bin chunk size, constant / missing / outlier observations (long runs of
    equal scores exercise the tie rule), stay == move (p = 0.2) and a vanishing move probability."""
    rng = np.random.default_rng(17)
    missing = np.frombuffer(np.uint32(0x7F800001).tobytes(), np.float32)[0]
    lanes = []
    for n in (1, 2, 63, 64, 65, 127, 128, 129, 191, 192, 193, 1000, 4097):
        lanes.append(rng.normal(1.0, 0.1, n).astype(np.float32))
    lanes.append(np.full(300, 0.75, np.float32))                 # exactly between CN1 and CN2
    lanes.append(np.full(200, missing, np.float32))              # no information at all
    lanes.append(np.zeros(150, np.float32))
    x = rng.normal(1.0, 0.1, 700).astype(np.float32)
    x[100:140] = missing
    x[300] = 1e30
    x[301] = -5.0
    x[400:470] = 0.5
    x[600:] = np.float32(np.inf)
    lanes.append(x)
    q = np.round(rng.normal(1.0, 0.3, 5000) * 4) / 4             # heavily quantised: many exact ties
    lanes.append(q.astype(np.float32))
    lengths = [len(v) for v in lanes]
    off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    cv = np.concatenate(lanes)
    sig = rng.uniform(0.03, 0.2, len(lanes))
    opts = discover.DiscoverOptions(p_move=p_move)
    st = to_np(discover.hmm_segment(cv, off, sig, opts, ctx=ctx))
    for l in range(len(lanes)):
        t = oracle.hmm_tables(float(sig[l]), opts.sigma0, opts.eps_cn, opts.p_move)
        ref = oracle.hmm_viterbi(lanes[l], t)
        assert np.array_equal(st[int(off[l]):int(off[l + 1])], ref), (l, lengths[l])


def test_posteriors_within_tolerance(oracle, ctx):
    lengths = [3000, 500, 8000]
    cv, _, off = _lanes(5, lengths)
    sig = np.array([0.08, 0.1, 0.12])
    opts = discover.DiscoverOptions(p_move=1e-3)
    st, post = discover.hmm_segment(cv, off, sig, opts, posterior=True, ctx=ctx)
    for l in range(len(lengths)):
        t = oracle.hmm_tables(float(sig[l]), opts.sigma0, opts.eps_cn, opts.p_move)
        sl = slice(int(off[l]), int(off[l + 1]))
        ref = oracle.hmm_posterior(cv[sl], t)
        np.testing.assert_allclose(post[sl], ref, rtol=1e-6, atol=1e-300)   # north_star: rel 1e-6 in f64
        assert np.array_equal(st[sl], oracle.hmm_viterbi(cv[sl], t))
    np.testing.assert_allclose(post.sum(axis=1), 1.0, rtol=1e-12)


def test_hmm_argument_errors(ctx):
    with pytest.raises(cb.CnvbError):
        discover.hmm_segment(np.zeros(4, np.float32), [0, 4], [0.1], discover.DiscoverOptions(p_move=0.5), ctx=ctx)


def test_full_size_lanes_match_oracle_and_are_independent(oracle, ctx):
    """BASELINE config 4 at full size (100 samples x 22 contigs at 1 kbp: 2200 lanes, 2.9e8 bins): lanes
    drawn across the range of lengths are recomputed by the oracle (a single lane is milliseconds on the
    CPU) and must be identical; the lanes of a second call on a subset must equal those of the full call
    (a lane's path depends on nothing but the lane)."""
    import torch
    from cnvetti_b200.synth import HG38_AUTOSOMES
    n_samples = 100
    lens = [(L + 999) // 1000 for L in HG38_AUTOSOMES]
    lane_len = np.array(lens * n_samples, np.int64)
    off = np.concatenate([[0], np.cumsum(lane_len)]).astype(np.uint64)
    n = int(off[-1])
    g = torch.Generator(device="cuda")
    g.manual_seed(44)
    cn = torch.full((n,), 2.0, device="cuda")
    nseg = 20 * len(lane_len)
    st = (torch.rand(nseg, generator=g, device="cuda") * (n - 600)).long()
    ln = (torch.rand(nseg, generator=g, device="cuda") * 495).long() + 5
    val = torch.tensor([0.0, 1.0, 3.0, 4.0], device="cuda")[(torch.rand(nseg, generator=g, device="cuda") * 4).long().clamp_(0, 3)]
    delta = torch.zeros(n + 1, device="cuda")
    delta.index_add_(0, st, val - 2.0)
    delta.index_add_(0, st + ln, -(val - 2.0))
    mu = (cn + torch.cumsum(delta[:-1], 0)).clamp_(0.0, 4.0).round_() / 2.0
    cv = (mu + torch.randn(n, generator=g, device="cuda") * (0.08 * mu.sqrt() + 0.02)).float().contiguous()
    sigma = np.full(len(lane_len), 0.08)
    states = discover.hmm_segment(cv, off, sigma, ctx=ctx)
    assert int((states != 2).sum()) > 1_000_000
    tab = oracle.hmm_tables(0.08)
    rng = np.random.default_rng(5)
    pick = sorted(set([0, 21, len(lane_len) - 1] + rng.integers(0, len(lane_len), 30).tolist()))
    for lane in pick:
        a, b = int(off[lane]), int(off[lane + 1])
        want = oracle.hmm_viterbi(cv[a:b].cpu().numpy(), tab)
        assert np.array_equal(states[a:b].cpu().numpy(), want), lane
    # a subset of whole samples in a call of its own
    lo_lane, hi_lane = 22 * 37, 22 * 41
    a, b = int(off[lo_lane]), int(off[hi_lane])
    sub = discover.hmm_segment(cv[a:b].contiguous(), off[lo_lane:hi_lane + 1] - off[lo_lane], sigma[lo_lane:hi_lane], ctx=ctx)
    assert torch.equal(sub, states[a:b])
