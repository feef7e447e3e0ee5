"""Pile collection on the GPU (cnvb_piles_collect, lib-coverage/src/piles.rs) against the oracle's
literal restatement: per-base depth vector of the pileup pass, then the base-by-base walk."""
import numpy as np
import pytest

import cnvetti_b200 as cb
from cnvetti_b200 import piles, pipeline, synth
from util import oracle_reads, simple_reads, soa_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return cb.Context(0)


def _oracle_piles(oracle, d, L, min_mapq, gap):
    wl = 64
    depth = oracle.coverage(oracle_reads(oracle, d), oracle.cov_opts(window_length=wl, min_mapq=min_mapq), L, debug=True)[3]
    return oracle.piles_from_depths(depth[:L], gap)


def _check(oracle, ctx, d, L, min_mapq=0, gap=20, device=None):
    got = piles.PileCollector(gap, min_mapq, ctx=ctx).collect_piles(soa_batch(d, device=device), L)
    es, ee, esz = _oracle_piles(oracle, d, L, min_mapq, gap)
    assert np.array_equal(got.start, es), (got.start[:10], es[:10])
    assert np.array_equal(got.end, ee)
    assert np.array_equal(got.size, esz)
    return got


@pytest.mark.parametrize("gap,min_mapq,device", [(20, 0, None), (0, 20, "cuda"), (150, 30, "cuda"), (1, 0, None)])
def test_piles_parity_synthetic(oracle, ctx, gap, min_mapq, device):
    L = 400_000
    d = synth.synth_bam_records(31, L, 6_000, unpaired_frac=0.05, refskip_frac=0.002)
    L = max(L, int(d["pos"].max()) + 20_000)  # (ref-skip reads of the generator may reach past its nominal length)
    got = _check(oracle, ctx, d, L, min_mapq, gap, device)
    assert len(got) > 20 and got.size.max() >= 2


def test_piles_corner_cases(oracle, ctx):
    L = 1000
    # gap exactly at / just above the limit, nested reads, equal starts, filtered reads splitting a pile
    cases = [
        simple_reads([10, 40], [20, 50]),            # gap 20: joined (end + 20 >= pos)
        simple_reads([10, 41], [20, 50]),            # gap 21: two piles
        simple_reads([10, 12, 12, 12, 30], [100, 20, 40, 13, 35]),
        simple_reads([5, 5, 5], [6, 6, 6]),
        simple_reads([10, 15, 40], [20, 18, 50], flag=[0, 0x400, 0]),
        simple_reads([10, 21, 300], [20, 31, 999], mapq=[60, 3, 60]),
        simple_reads([], []),
        simple_reads([0], [1000]),
    ]
    for d in cases:
        _check(oracle, ctx, d, L, min_mapq=10, gap=20)
    # everything filtered
    got = piles.PileCollector(20, 50, ctx=ctx).collect_piles(soa_batch(simple_reads([1, 2], [5, 6], mapq=[3, 4])), L)
    assert len(got) == 0


def test_piles_random_small(oracle, ctx):
    rng = np.random.default_rng(9)
    for it in range(30):
        L = int(rng.integers(50, 3000))
        n = int(rng.integers(0, 120))
        pos = np.sort(rng.integers(0, L - 1, n))
        end = np.minimum(pos + rng.integers(1, 80, n), L)
        flag = rng.choice(np.array([0, 0, 0, 0x400, 0x100, 0x800, 0x200, 0x4, 0x1, 0x3], np.uint16), n)
        d = simple_reads(pos, end, flag, rng.integers(0, 61, n))
        _check(oracle, ctx, d, L, min_mapq=int(rng.integers(0, 30)), gap=int(rng.integers(0, 40)))


def test_piles_errors(ctx):
    with pytest.raises(cb.ReferencePanic, match="piles.rs"):
        piles.PileCollector(20, 0, ctx=ctx).collect_piles(soa_batch(simple_reads([90], [130])), 100)
    with pytest.raises(cb.CnvbError, match="coordinate-sorted"):
        piles.PileCollector(20, 0, ctx=ctx).collect_piles(soa_batch(simple_reads([50, 10], [60, 20])), 100)


def test_gather_piles_feeds_the_aggregator(oracle, ctx):
    """lib.rs:353-419 end to end: piles per contig -> size histogram -> FDR threshold -> black list,
    then the genome-wide fragment counts with that black list, all against the oracle chain."""
    regions, raw = [], []
    for i, L in enumerate((300_000, 120_000)):
        d = synth.synth_bam_records(40 + i, L, 8_000)
        L = max(L, int(d["pos"].max()) + 20_000)
        regions.append(pipeline.Region("chr%d" % (i + 1), L, soa_batch(d, device="cuda")))
        raw.append((d, L))
    black, thresh = piles.gather_piles(regions, min_mapq=0, pile_max_gap=20, mask_piles_fdr=0.01, ctx=ctx)
    infos = [_oracle_piles(oracle, d, L, 0, 20) for d, L in raw]
    top = max(int(s.max()) for _, _, s in infos)
    hist = np.zeros(top + 1, np.uint64)
    for _, _, s in infos:
        hist += np.bincount(s, minlength=top + 1).astype(np.uint64)
    assert thresh == oracle.pile_threshold(hist, 0.01)
    for (bs, be), (s, e, sz) in zip(black, infos):
        assert np.array_equal(bs, s[sz >= thresh]) and np.array_equal(be, e[sz >= thresh])
    assert sum(len(b[0]) for b in black) > 0
    opts = cb.CoverageOptions(window_length=500, min_mapq=0)
    for r, b, (d, L) in zip(regions, black, raw):
        r.piles = b
    t = pipeline.coverage_run(regions, opts, ctx=ctx)
    oo = oracle.cov_opts(window_length=500, min_mapq=0)
    for (d, L), b, lo, hi, skipped in zip(raw, black, t.offsets[:-1], t.offsets[1:], t.num_skipped):
        counts, proc, skip = oracle.frag_gw(oracle_reads(oracle, d), oo, L, b[0], b[1])
        assert np.array_equal(t.rcv[lo:hi].cpu().numpy(), counts.astype(np.float32)) and skipped == skip
