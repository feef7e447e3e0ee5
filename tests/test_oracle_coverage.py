"""CPU checks of the oracle's aggregator restatements (parity unpinned by the reference's own
tests: the reference has none for these paths) against independent
pure-Python / numpy models of the cited reference lines, and of the product's host packer."""
import numpy as np
import pytest

from cnvetti_b200 import pack_reads, synth


def _reads(oracle, d):
    return oracle.ReadBatch(d["pos"], d["isize"], d["flag"], d["mapq"], d["cigar_off"], d["cigar"])


def _simple_reads(pos, end, flag=None, mapq=None):
    """single-op (M) reads covering [pos, end)"""
    pos = np.asarray(pos, np.int32)
    end = np.asarray(end, np.int32)
    n = pos.size
    cig = ((end - pos).astype(np.uint32) << 4)
    return dict(pos=pos, isize=np.zeros(n, np.int32),
                flag=np.zeros(n, np.uint16) if flag is None else np.asarray(flag, np.uint16),
                mapq=np.full(n, 60, np.uint8) if mapq is None else np.asarray(mapq, np.uint8),
                cigar_off=np.arange(n + 1, dtype=np.uint64), cigar=cig)


def test_coverage_state_machine_example(oracle):
    # wl=4, depth(p)=p+1 for p=0..13
    d = _simple_reads(np.arange(14), np.full(14, 14))
    opts = oracle.cov_opts(window_length=4)
    mean, sd, cols, depth = oracle.coverage(_reads(oracle, d), opts, 14, debug=True)
    assert depth[:14].tolist() == list(range(1, 15))
    assert mean.tolist() == [3.5, 7.5, 11.5, 0.0]
    for w, vec in enumerate(([5, 2, 3, 4], [9, 6, 7, 8], [13, 10, 11, 12])):
        assert sd[w] == np.float32(np.std(np.array(vec, np.float64), ddof=1))
    assert sd[3] == 0.0


def _coverage_model(cols, depth, wl, nb):
    """literal Python transcription of agg.rs:523-573 driven by per-base arrays"""
    mean = np.zeros(nb, np.float32)
    sd = np.zeros(nb, np.float32)
    depths = [0] * wl
    window_id = None
    prev = None

    def push(w):
        nonlocal depths
        v = np.array(depths, np.float64)
        mean[w] = np.float32(v.sum() / wl)
        sd[w] = np.float32(np.sqrt(((v - v.sum() / wl) ** 2).sum() / (wl - 1))) if wl > 1 else 0
        depths = [0] * wl

    for pos in range(len(cols)):
        if cols[pos] == 0:
            continue
        if window_id is not None and prev is not None:
            if window_id != prev:
                push(prev)
                prev = window_id
            window_id = pos // wl
        elif window_id is not None:
            prev = window_id
            window_id = pos // wl
        else:
            window_id = pos // wl
        depths[pos % wl] = int(depth[pos])
    if window_id is not None:
        if prev is None or prev != window_id:
            push(window_id)
    return mean, sd


@pytest.mark.parametrize("seed", range(6))
def test_coverage_vs_python_model(oracle, seed):
    rng = np.random.default_rng(seed)
    wl = int(rng.integers(3, 40))
    L = int(rng.integers(5, 30)) * wl + int(rng.integers(0, wl))
    n = int(rng.integers(0, 60))
    pos = np.sort(rng.integers(0, max(1, L - 2), n))
    end = np.minimum(pos + rng.integers(1, 3 * wl, n), L)
    flag = rng.choice(np.array([0, 0, 0, 0x400, 0x100, 0x800, 0x200, 0x4, 0x1, 0x3], np.uint16), n)
    mapq = rng.integers(0, 61, n)
    d = _simple_reads(pos, end, flag, mapq)
    opts = oracle.cov_opts(window_length=wl, min_mapq=20)
    mean, sd, cols, depth = oracle.coverage(_reads(oracle, d), opts, L, debug=True)
    nb = (L + wl - 1) // wl
    # independent per-base model
    c2 = np.zeros(nb * wl, np.int64)
    d2 = np.zeros(nb * wl, np.int64)
    for i in range(n):
        f = int(flag[i])
        if f & 0x704:
            continue
        c2[pos[i]:end[i]] += 1
        ok = not (f & 0xF00) and mapq[i] >= 20 and (not (f & 1) or (f & 2))
        if ok:
            d2[pos[i]:end[i]] += 1
    assert np.array_equal(cols, c2) and np.array_equal(depth, d2)
    m2, s2 = _coverage_model(c2, d2, wl, nb)
    assert np.array_equal(mean, m2)
    np.testing.assert_allclose(sd, s2, rtol=1e-6)


def _gw_model(d, min_mapq, min_unclipped, wl, region_end):
    nb = (region_end + wl - 1) // wl
    counts = np.zeros(nb, np.uint32)
    proc = 0
    for i in range(d["pos"].size):
        f = int(d["flag"][i])
        if d["mapq"][i] < min_mapq or f & 0xF00:
            continue
        if (f & 1) and not (f & 2):
            continue
        cl = un = 0
        for c in d["cigar"][int(d["cigar_off"][i]):int(d["cigar_off"][i + 1])]:
            op, ln = int(c) & 15, int(c) >> 4
            if op in (0, 1, 7, 8):
                un += ln
            elif op in (4, 5):
                cl += ln
        if np.float32(un) / np.float32(cl + un) < np.float32(min_unclipped):
            continue
        if (f & 1) and d["isize"][i] <= 0:
            continue
        proc += 1
        centre = int(d["pos"][i]) + int(np.trunc(int(d["isize"][i]) / 2))
        counts[centre // wl] += 1
    return counts, proc


def test_frag_gw_vs_python_model(oracle):
    d = synth.synth_bam_records(7, 200_000, 3000)
    opts = oracle.cov_opts(window_length=1000, min_mapq=20)
    counts, proc, skip = oracle.frag_gw(_reads(oracle, d), opts, 200_000)
    c2, p2 = _gw_model(d, 20, 0.6, 1000, 200_000)
    assert np.array_equal(counts, c2) and proc == p2 and skip == 0
    assert counts.sum() == proc


def test_frag_gw_piles(oracle):
    d = synth.synth_bam_records(8, 100_000, 2000)
    opts = oracle.cov_opts(window_length=500)
    ps = np.array([1000, 40_000, 77_777], np.uint32)
    pe = np.array([3000, 40_010, 80_001], np.uint32)
    counts, proc, skip = oracle.frag_gw(_reads(oracle, d), opts, 100_000, ps, pe)
    c0, p0, _ = oracle.frag_gw(_reads(oracle, d), opts, 100_000)
    assert proc == p0 and skip > 0 and counts.sum() + skip == proc
    cov, start, end, frm = oracle.frag_gw_stats(counts, opts, 100_000, ps, pe)
    assert frm[2] == 1.0 and frm[3] == 1.0 and frm[5] == 1.0 and frm[80] == np.float32(10 / 500)
    assert frm[155] == np.float32(223 / 500) and frm[160] == np.float32(1 / 500)
    assert end[-1] == 100_000 and start[1] == 500


def test_frag_gw_unpaired_panics(oracle):
    d = _simple_reads([50_000], [50_100])  # unpaired: centre = pos + end_pos (agg.rs:127-132)
    opts = oracle.cov_opts(window_length=1000)
    with pytest.raises(oracle.OraclePanic):
        oracle.frag_gw(_reads(oracle, d), opts, 60_000)


def test_pack_reads_matches_oracle_derivations(oracle):
    d = synth.synth_bam_records(9, 300_000, 4000, unpaired_frac=0.1, refskip_frac=0.01)
    end, mid, frag_end, flags = pack_reads(d["pos"], d["isize"], d["flag"], d["cigar_off"], d["cigar"], 0.6)
    r = _reads(oracle, d)
    L = oracle.lib()
    import ctypes as C
    L.orc_bam_endpos.restype = C.c_int32
    L.orc_cigar_end_pos.restype = C.c_int32
    for i in range(0, r.n, 37):
        assert end[i] == L.orc_bam_endpos(C.byref(r.c), C.c_uint64(i))
        ep = L.orc_cigar_end_pos(C.byref(r.c), C.c_uint64(i))
        paired = d["flag"][i] & 1
        assert frag_end[i] == (d["pos"][i] + d["isize"][i] if paired else ep)
        assert mid[i] == (d["pos"][i] + int(np.trunc(d["isize"][i] / 2)) if paired else d["pos"][i] + ep)
        assert (flags[i] & 0x0FFF) == d["flag"][i]
        assert bool(flags[i] & 0x2000) == (d["isize"][i] <= 0)


def test_targets_tie_and_overlap(oracle):
    # two targets equidistant from the fragment centre: lowest index wins (documented rule)
    tstart = np.array([100, 301], np.uint32)
    tend = np.array([200, 400], np.uint32)
    d = _simple_reads([150], [250])
    d["flag"][:] = 0x3 | 0x40
    d["isize"][:] = 200          # fragment [150, 350), centre 250: dist 51 / 51
    opts = oracle.cov_opts()
    counts, proc, skip = oracle.frag_targets(_reads(oracle, d), opts, tstart, tend)
    assert counts.tolist() == [1, 0] and proc == 1 and skip == 0
    d["pos"][:] = 500            # no overlap -> skipped
    counts, proc, skip = oracle.frag_targets(_reads(oracle, d), opts, tstart, tend)
    assert counts.sum() == 0 and skip == 1


def test_refstats_small(oracle):
    seq = np.frombuffer(b"ACGTNNacgtRYGGGGCCCCAAAATTTTnnnnXXXX", np.uint8)
    gc, gap = oracle.refstats(seq, 4)
    assert gc[0] == 0.5 and gap[0] == 0
    assert gc[1] == 0.5 and gc[2] == 0.5 and gap[1] == 1       # "NNac": a,c
    assert gc[3] == 1.0 and gc[4] == 1.0 and gc[5] == 0.0 and gc[6] == 0.0
    assert np.isnan(gc[7]) and gap[7] == 1 and np.isnan(gc[8]) and gap[8] == 0
    assert gc[7].view(np.uint32) == 0x7F800001


def test_norm_gc_median_lower_median(oracle):
    # 12 windows in GC bin 40 (even count -> lower median), 9 windows in bin 55 (too few)
    lcv = np.array(list(range(1, 13)) + [5.0] * 9, np.float32)
    gc = np.array([0.405] * 12 + [0.55] * 9, np.float32)
    cv, cv2, cvsd, flags, med = oracle.norm_gc_median(lcv, gc)
    assert med[40] == 6.0 and med[55].view(np.uint32) == 0x7F800001
    assert np.array_equal(cv[:12], lcv[:12] / np.float32(6.0))
    assert (flags[:12] == (oracle.N2_CV | oracle.N2_CV2)).all() and (flags[12:] == oracle.N2_FEW_GC).all()
