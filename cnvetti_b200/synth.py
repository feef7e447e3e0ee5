"""Seeded synthetic inputs of the shapes SURVEY.md section 8(d) names (no network, no real data).

All generators are deterministic in their seed.  BAM-level generators return what a BAM record
carries (pos, insert size, FLAG, MAPQ, CIGAR); `cnvetti_b200.pack_reads` turns them into the SoA
batch the kernels consume.
"""
import numpy as np

HG38_AUTOSOMES = [
    248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
    138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
    83257441, 80373285, 58617616, 64444167, 46709983, 50818468,
]
CHR22_LEN = HG38_AUTOSOMES[21]

CIG_M, CIG_I, CIG_D, CIG_N, CIG_S, CIG_H = 0, 1, 2, 3, 4, 5


def _cnv_density(contig_len, rng, n_cnv):
    """Piecewise-constant relative density with planted CNVs: (starts, ends, factor)."""
    segs = []
    for k in range(n_cnv):
        length = int(rng.integers(200_000, 2_000_000)) if contig_len > 8_000_000 else \
            int(rng.integers(max(2, contig_len // 40), max(3, contig_len // 10)))
        start = int(rng.integers(0, max(1, contig_len - length)))
        segs.append((start, start + length, 0.5 if k % 2 == 0 else 1.5))
    return segs


def synth_bam_records(seed, contig_len, n_pairs, read_len=100, n_cnv=6, unpaired_frac=0.0,
                      refskip_frac=0.0):
    """Coordinate-sorted paired-end records on one contig.

    isize ~ N(350,50) clipped to [120,800]; MAPQ 88% 60 / 7% U[1,59] / 5% 0; 1% dup, 0.5%
    secondary, 0.5% supplementary, 0.3% qcfail, 2% improper pairs; CIGAR 94% <L>M, 4%
    soft-clipped (1..60 bases), 2% one I or D of <= 10.  `unpaired_frac` adds single-end records
    (only meaningful for the Targets / Coverage kinds: genome-wide fragment counting gives
    unpaired reads the centre pos+end_pos, lib-coverage/src/agg.rs:127-132).
    Returns dict(pos, isize, flag, mapq, cigar_off, cigar) of numpy arrays.
    """
    rng = np.random.default_rng(seed)
    segs = _cnv_density(contig_len, rng, n_cnv)
    max_isize = 800
    # fragment starts: rejection sampling against the CNV density
    want = n_pairs
    starts = []
    while want > 0:
        cand = rng.integers(0, max(1, contig_len - max_isize), size=int(want * 1.6) + 16)
        dens = np.ones(cand.size)
        for (s, e, f) in segs:
            dens[(cand >= s) & (cand < e)] = f
        keep = cand[rng.random(cand.size) * 1.5 < dens][:want]
        starts.append(keep)
        want -= keep.size
    fstart = np.concatenate(starts).astype(np.int64)
    n_pairs = fstart.size
    isize = np.clip(np.rint(rng.normal(350, 50, n_pairs)), max(120, read_len), max_isize).astype(np.int64)

    improper = rng.random(n_pairs) < 0.02
    single = rng.random(n_pairs) < unpaired_frac

    def mate_fields(n):
        mapq = np.full(n, 60, np.uint8)
        u = rng.random(n)
        low = (u >= 0.88) & (u < 0.95)
        mapq[low] = rng.integers(1, 60, int(low.sum()))
        mapq[u >= 0.95] = 0
        extra = np.zeros(n, np.uint16)
        extra[rng.random(n) < 0.01] |= 0x400
        extra[rng.random(n) < 0.005] |= 0x100
        extra[rng.random(n) < 0.005] |= 0x800
        extra[rng.random(n) < 0.003] |= 0x200
        return mapq, extra

    mq1, ex1 = mate_fields(n_pairs)
    mq2, ex2 = mate_fields(n_pairs)
    proper = np.where(improper, 0, 0x2).astype(np.uint16)
    flag1 = (0x1 | 0x20 | 0x40) | proper | ex1
    flag2 = (0x1 | 0x10 | 0x80) | proper | ex2
    flag1 = np.where(single, ex1, flag1).astype(np.uint16)  # single-end: no pairing bits
    pos1 = fstart
    pos2 = fstart + isize - read_len
    is1 = np.where(single, 0, isize)
    is2 = -isize
    keep2 = ~single
    pos = np.concatenate([pos1, pos2[keep2]])
    isz = np.concatenate([is1, is2[keep2]])
    flag = np.concatenate([flag1, flag2[keep2]])
    mapq = np.concatenate([mq1, mq2[keep2]])
    n = pos.size

    # CIGARs: up to 3 ops per read
    kind = rng.random(n)
    ops = np.zeros((n, 3), np.uint32)
    nops = np.ones(n, np.int64)
    ops[:, 0] = (read_len << 4) | CIG_M
    sc = (kind >= 0.94) & (kind < 0.98)
    clip = rng.integers(1, min(61, read_len), n).astype(np.uint32)
    left = rng.random(n) < 0.5
    a = sc & left
    ops[a, 0] = (clip[a] << 4) | CIG_S
    ops[a, 1] = ((read_len - clip[a]) << 4) | CIG_M
    b = sc & ~left
    ops[b, 0] = ((read_len - clip[b]) << 4) | CIG_M
    ops[b, 1] = (clip[b] << 4) | CIG_S
    nops[sc] = 2
    indel = kind >= 0.98
    k = rng.integers(1, 11, n).astype(np.uint32)
    first = rng.integers(10, read_len - 20, n).astype(np.uint32)
    ins = indel & (rng.random(n) < 0.5)
    dele = indel & ~ins
    ops[ins, 0] = (first[ins] << 4) | CIG_M
    ops[ins, 1] = (k[ins] << 4) | CIG_I
    ops[ins, 2] = ((read_len - first[ins] - k[ins]) << 4) | CIG_M
    ops[dele, 0] = (first[dele] << 4) | CIG_M
    ops[dele, 1] = (k[dele] << 4) | CIG_D
    ops[dele, 2] = ((read_len - first[dele]) << 4) | CIG_M
    nops[indel] = 3
    if refskip_frac > 0:
        rs = rng.random(n) < refskip_frac
        gap = rng.integers(50, 3000, n).astype(np.uint32)
        ops[rs, 0] = ((read_len // 2) << 4) | CIG_M
        ops[rs, 1] = (gap[rs] << 4) | CIG_N
        ops[rs, 2] = ((read_len - read_len // 2) << 4) | CIG_M
        nops[rs] = 3

    order = np.argsort(pos, kind="stable")
    pos, isz, flag, mapq, ops, nops = pos[order], isz[order], flag[order], mapq[order], ops[order], nops[order]
    cigar_off = np.zeros(n + 1, np.uint64)
    np.cumsum(nops, out=cigar_off[1:])
    mask = np.arange(3)[None, :] < nops[:, None]
    cigar = ops[mask]
    return dict(pos=pos.astype(np.int32), isize=isz.astype(np.int32), flag=flag.astype(np.uint16),
                mapq=mapq.astype(np.uint8), cigar_off=cigar_off, cigar=cigar.astype(np.uint32))


def synth_reference(seed, length, block=10_000, gap_frac=0.02, lowercase_frac=0.3):
    """Synthetic contig sequence: GC fraction per `block` drawn from Beta(8,11), a fraction of
    blocks are N gaps, soft-masked (lower-case) stretches, a sprinkle of IUPAC codes."""
    rng = np.random.default_rng(seed)
    nblk = (length + block - 1) // block
    gcf = rng.beta(8, 11, nblk)
    u = rng.random(length)
    gcb = np.repeat(gcf, block)[:length]
    is_gc = u < gcb
    pick = rng.random(length) < 0.5
    seq = np.where(is_gc, np.where(pick, ord("G"), ord("C")), np.where(pick, ord("A"), ord("T"))).astype(np.uint8)
    lower = np.repeat(rng.random(nblk) < lowercase_frac, block)[:length]
    seq[lower] |= 0x20
    gaps = np.repeat(rng.random(nblk) < gap_frac, block)[:length]
    seq[gaps] = ord("N")
    iupac = rng.random(length) < 1e-4
    seq[iupac] = rng.choice(np.frombuffer(b"RYKMSWnrym", np.uint8), int(iupac.sum()))
    return seq


def synth_targets(seed, contig_len, n_targets, min_len=80, max_len=400, overlap_frac=0.0):
    """Sorted target intervals (exome-like).  `overlap_frac` > 0 makes some neighbours overlap."""
    rng = np.random.default_rng(seed)
    starts = np.sort(rng.integers(0, max(1, contig_len - max_len), n_targets)).astype(np.int64)
    lens = rng.integers(min_len, max_len + 1, n_targets)
    ends = starts + lens
    if overlap_frac <= 0:
        # make disjoint: push starts right where needed
        for i in range(1, n_targets):
            if starts[i] <= ends[i - 1]:
                starts[i] = ends[i - 1] + 1 + int(rng.integers(0, 50))
                ends[i] = starts[i] + lens[i]
        ok = ends <= contig_len
        starts, ends = starts[ok], ends[ok]
    return starts.astype(np.uint32), ends.astype(np.uint32)


def synth_panel(seed, T, S, n_contigs=22):
    """WIS panel X[T][S] (f32) as SURVEY 8(d) config 3: eff_t ~ LogNormal(0,0.4), noise ~ N(1,0.08),
    column-normalised to sum 1 (mimics TotalCovSum); tids from contigs proportional to length."""
    rng = np.random.default_rng(seed)
    eff = rng.lognormal(0.0, 0.4, T)
    X = eff[:, None] * rng.normal(1.0, 0.08, (T, S))
    X = np.maximum(X, 0.0)
    X = X / X.sum(axis=0, keepdims=True)
    w = np.array(HG38_AUTOSOMES[:n_contigs], np.float64)
    bounds = np.floor(np.cumsum(w) / w.sum() * T).astype(np.int64)
    tids = np.searchsorted(bounds, np.arange(T), side="right").astype(np.uint32)
    tids = np.minimum(tids, n_contigs - 1).astype(np.uint32)
    return X.astype(np.float32), tids


def synth_cv_lanes(seed, n_bins, n_cnv=20, noise=0.08, sigma0=0.02, missing_frac=0.001):
    """One normalised coverage lane (config 4): CV ~ N(cn/2, noise*sqrt(cn/2)+sigma0) with planted
    segments cn in {0,1,3,4} of 5..500 bins.  Returns (cv f32, true cn u8)."""
    rng = np.random.default_rng(seed)
    cn = np.full(n_bins, 2, np.uint8)
    for _ in range(n_cnv):
        ln = int(rng.integers(5, 501))
        if n_bins <= ln + 1:
            continue
        s = int(rng.integers(0, n_bins - ln))
        cn[s:s + ln] = rng.choice(np.array([0, 1, 3, 4], np.uint8))
    mu = cn / 2.0
    cv = rng.normal(mu, noise * np.sqrt(mu) + sigma0).astype(np.float32)
    miss = rng.random(n_bins) < missing_frac
    cv[miss] = np.frombuffer(np.uint32(0x7F800001).tobytes(), np.float32)[0]
    return cv, cn


# ---- synthetic BAM files (tests and the host-decode leg of bench.py) ------------------------------
def _bgzf_block(payload, level):
    import struct
    import zlib
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    data = c.compress(payload) + c.flush()
    head = struct.pack("<4BI2BH2BHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, ord("B"), ord("C"), 2, len(data) + 25)
    return head + data + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload))


def _write_bai(path, n_ref, tids, beg, end, vbeg, vend):
    """The .bai of a BAM file from its records (SAM specification 5.2: bins of reg2bin(beg, end) with their chunks,
    16 kbp linear index, virtual offsets).  Test / bench infrastructure: the product READS such an index."""
    import struct
    out = [b"BAI\x01", struct.pack("<i", n_ref)]
    b, e = beg.astype(np.int64), np.maximum(end.astype(np.int64), beg.astype(np.int64) + 1) - 1
    bins = np.zeros(b.size, np.int64)
    done = np.zeros(b.size, bool)
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        m = ~done & ((b >> shift) == (e >> shift))
        bins[m] = base + (b[m] >> shift)
        done |= m
    for r in range(n_ref):
        sel = np.flatnonzero(tids == r)
        if sel.size == 0:
            out.append(struct.pack("<ii", 0, 0))
            continue
        ub = np.unique(bins[sel])
        parts = [struct.pack("<i", ub.size + 1)]
        for bn in ub:
            idx = sel[bins[sel] == bn]
            cut = np.flatnonzero(np.diff(idx) != 1) + 1       # records consecutive in the file share a chunk
            starts = np.concatenate([[0], cut])
            stops = np.concatenate([cut, [idx.size]]) - 1
            parts.append(struct.pack("<Ii", int(bn), starts.size))
            for a, z in zip(starts, stops):
                parts.append(struct.pack("<QQ", int(vbeg[idx[a]]), int(vend[idx[z]])))
        parts.append(struct.pack("<IiQQQQ", 37450, 2, int(vbeg[sel[0]]), int(vend[sel[-1]]), int(sel.size), 0))
        n_intv = int(e[sel].max() >> 14) + 1
        lin = np.full(n_intv, -1, np.int64)
        for i in sel:                                          # first record overlapping each 16 kbp window
            w0, w1 = int(b[i] >> 14), int(e[i] >> 14)
            seg = lin[w0:w1 + 1]
            seg[seg < 0] = int(vbeg[i])
        for w in range(n_intv - 2, -1, -1):
            if lin[w] < 0:
                lin[w] = lin[w + 1]
        parts.append(struct.pack("<i", n_intv))
        parts.append(lin.astype("<u8").tobytes())
        out += parts
    with open(path, "wb") as f:
        f.write(b"".join(out))


def write_bam(path, contigs, records, sample="SYN1", level=1, read_len=None, block_payload=0xff00, index=False,
              payload="filler"):
    """Write a coordinate-sorted BAM file.  `contigs`: [(name, length)]; `records`: one dict per contig (or
    None) with the BAM-level arrays of `synth_bam_records` (pos, isize, flag, mapq, cigar_off, cigar);
    an optional entry with key -1 in a dict `records` holds reads without a reference.  Read names,
    sequences and qualities are fillers (the coverage path never looks at them): zeros and 0xFF, which deflate
    15 : 1 -- payload="random" writes random bases and qualities drawn from eight bins instead (about 3 : 1, what
    a sequencer's BAM file compresses to), the honest input for timing the decoder.  The records are laid
    out with numpy, so tens of millions of reads are practical."""
    import struct
    if isinstance(records, dict):
        extra = records.get(-1)
        records = [records.get(i) for i in range(len(contigs))]
    else:
        extra = None
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % c for c in contigs) + \
           "@RG\tID:rg1\tSM:%s\n" % sample
    head = b"BAM\x01" + struct.pack("<I", len(text)) + text.encode() + struct.pack("<I", len(contigs))
    for name, length in contigs:
        head += struct.pack("<I", len(name) + 1) + name.encode() + b"\x00" + struct.pack("<I", length)
    chunks = [np.frombuffer(head, np.uint8)]
    ix = dict(tid=[], beg=[], end=[], uoff=[], ulen=[])
    stream_at = len(head)
    l_name = 8  # "r" + 6 characters + NUL
    for tid, d in list(enumerate(records)) + [(-1, extra)]:
        if d is None or len(d["pos"]) == 0:
            continue
        n = len(d["pos"])
        coff = np.asarray(d["cigar_off"], np.int64)
        ncig = (coff[1:] - coff[:-1]).astype(np.int64)
        cig = np.asarray(d["cigar"], np.uint32)
        # query length = sum of M I S = X ops (what l_seq must be)
        ops, lens = cig & 0xF, (cig >> 4).astype(np.int64)
        q = np.where(np.isin(ops, (0, 1, 4, 7, 8)), lens, 0)
        cs = np.concatenate([[0], np.cumsum(q)])
        l_seq = (cs[coff[1:]] - cs[coff[:-1]]).astype(np.int64)
        if read_len is not None:
            l_seq = np.full(n, int(read_len), np.int64)
        size = 32 + l_name + 4 * ncig + (l_seq + 1) // 2 + l_seq          # block_size
        off = np.concatenate([[0], np.cumsum(size + 4)]).astype(np.int64)
        if payload == "random":   # (sequence bytes; every fixed field, name, CIGAR and quality is written over it)
            # (a random tile of 1 MiB + 13 bytes repeated: deflate's window is 32 KiB, it cannot see the repeats)
            nib = np.array([1, 2, 4, 8], np.uint8)   # A C G T in the 4-bit code
            two = np.random.default_rng(1000 + tid).integers(0, 16, (1 << 20) + 13, dtype=np.uint8)
            buf = np.resize((nib[two >> 2] << 4) | nib[two & 3], int(off[-1]))
        else:
            buf = np.zeros(int(off[-1]), np.uint8)
        o = off[:-1]

        def put(at, val, nbytes):
            v = np.asarray(val).astype(np.int64)
            for k in range(nbytes):
                buf[o + at + k] = (v >> (8 * k)) & 0xFF

        pos = np.asarray(d["pos"], np.int64)
        put(0, size, 4)
        put(4, np.full(n, tid, np.int64) & 0xFFFFFFFF, 4)
        put(8, pos & 0xFFFFFFFF, 4)
        put(12, np.full(n, l_name), 1)
        put(13, d["mapq"], 1)
        put(14, np.full(n, 4680), 2)                                       # bin: unused here
        put(16, ncig, 2)
        put(18, d["flag"], 2)
        put(20, l_seq, 4)
        put(24, np.full(n, tid, np.int64) & 0xFFFFFFFF, 4)                 # mate on the same contig
        put(28, (pos + np.asarray(d["isize"], np.int64)) & 0xFFFFFFFF, 4)  # (filler)
        put(32, np.asarray(d["isize"], np.int64) & 0xFFFFFFFF, 4)
        idx = np.arange(n)
        buf[o + 36 + l_name - 1] = 0
        buf[o + 36] = ord("r")
        for k in range(6):
            buf[o + 37 + k] = 48 + (idx // 10 ** (5 - k)) % 10
        # CIGAR words: record of every op, position inside its record
        rec_of = np.repeat(np.arange(n), ncig)
        within = np.arange(cig.size) - np.repeat(coff[:-1], ncig)
        base = o[rec_of] + 36 + l_name + 4 * within
        for k in range(4):
            buf[base + k] = (cig >> (8 * k)) & 0xFF
        # qualities 0xFF (missing)
        qbase = o + 36 + l_name + 4 * ncig + (l_seq + 1) // 2
        qrng = np.random.default_rng(2000 + tid)
        qbins = np.array([2, 12, 18, 24, 28, 32, 36, 40], np.uint8)
        qprob = [0.02, 0.02, 0.03, 0.05, 0.08, 0.15, 0.25, 0.40]
        if (l_seq == l_seq[0]).all():
            qv = 0xFF if payload != "random" else np.resize(qbins[qrng.choice(8, size=(1 << 20) + 29, p=qprob)], n * int(l_seq[0]))
            buf[(qbase[:, None] + np.arange(int(l_seq[0]))[None, :]).ravel()] = qv
        else:
            for i in range(n):
                buf[qbase[i]:qbase[i] + l_seq[i]] = 0xFF if payload != "random" else qbins[qrng.choice(8, size=int(l_seq[i]), p=qprob)]
        chunks.append(buf)
        if index and tid >= 0:
            rl = np.where(np.isin(ops, (0, 2, 3, 7, 8)), lens, 0)
            rcs = np.concatenate([[0], np.cumsum(rl)])
            ref_len = (rcs[coff[1:]] - rcs[coff[:-1]]).astype(np.int64)
            unmapped = (np.asarray(d["flag"], np.int64) & 4) != 0
            ref_len = np.where(unmapped | (ncig == 0) | (ref_len == 0), 1, ref_len)   # bam_endpos
            ix["tid"].append(np.full(n, tid, np.int64))
            ix["beg"].append(pos)
            ix["end"].append(pos + ref_len)
            ix["uoff"].append(stream_at + off[:-1])
            ix["ulen"].append(size + 4)
        stream_at += int(off[-1])
    stream = np.concatenate(chunks)
    block_addr = []
    with open(path, "wb") as f:
        for a in range(0, stream.size, block_payload):
            block_addr.append(f.tell())
            f.write(_bgzf_block(stream[a:a + block_payload].tobytes(), level))
        block_addr.append(f.tell())
        f.write(_bgzf_block(b"", level))  # EOF marker
    if index:
        cat = {k: (np.concatenate(v) if v else np.zeros(0, np.int64)) for k, v in ix.items()}
        addr = np.asarray(block_addr, np.int64)
        voff = lambda u: (addr[u // block_payload] << 16) | (u % block_payload)  # noqa: E731
        _write_bai(path + ".bai", len(contigs), cat["tid"], cat["beg"], cat["end"], voff(cat["uoff"]),
                   voff(cat["uoff"] + cat["ulen"]))
    return path
