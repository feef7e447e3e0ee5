"""Device-side synthetic inputs at BASELINE.json sizes (torch is plumbing here: RNG + sort).

Generating 900 M reads with numpy would take minutes; these generators fill HBM directly.  The
records follow the same recipe as `synth.synth_bam_records` but are
produced in the SoA layout the kernels consume, plus the BAM-level `isize` so that a bounded
sample can be handed to the CPU oracle.
"""
import numpy as np
import torch

READ_LEN = 100


def synth_soa_contig(seed, contig_len, n_pairs, device, min_unclipped=0.6):
    """Coordinate-sorted paired-end reads of one contig as device tensors.

    Returns dict(pos, end, mid, frag_end, isize int32; mapq uint8; flags int16 (FLAG | CNVB bits)).
    """
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    n = int(n_pairs)
    fstart = torch.randint(0, max(1, contig_len - 800), (n,), generator=g, device=device, dtype=torch.int32)
    isize = torch.empty(n, device=device).normal_(350.0, 50.0, generator=g).round_().clamp_(120, 800).to(torch.int32)
    improper = torch.rand(n, generator=g, device=device) < 0.02

    def mate(nn):
        u = torch.rand(nn, generator=g, device=device)
        mapq = torch.full((nn,), 60, dtype=torch.uint8, device=device)
        low = (u >= 0.88) & (u < 0.95)
        mapq[low] = torch.randint(1, 60, (int(low.sum()),), generator=g, device=device, dtype=torch.uint8)
        mapq[u >= 0.95] = 0
        fl = torch.zeros(nn, dtype=torch.int32, device=device)
        fl |= (torch.rand(nn, generator=g, device=device) < 0.01).to(torch.int32) * 0x400
        fl |= (torch.rand(nn, generator=g, device=device) < 0.005).to(torch.int32) * 0x100
        fl |= (torch.rand(nn, generator=g, device=device) < 0.005).to(torch.int32) * 0x800
        fl |= (torch.rand(nn, generator=g, device=device) < 0.003).to(torch.int32) * 0x200
        # soft clips: 4% of reads, clip U[1,60]; ratio (100-clip)/100 < min_unclipped fails
        clip = torch.randint(1, 61, (nn,), generator=g, device=device, dtype=torch.int32)
        clipped = torch.rand(nn, generator=g, device=device) < 0.04
        ratio = (READ_LEN - clip).to(torch.float32) / float(READ_LEN)
        clipfail = clipped & (ratio < float(np.float32(min_unclipped)))
        fl |= clipfail.to(torch.int32) * 0x1000
        clip = torch.where(clipped, clip, torch.zeros_like(clip))
        return mapq, fl, clip

    proper = torch.where(improper, torch.zeros_like(isize), torch.full_like(isize, 0x2))
    mq1, fl1, _ = mate(n)
    mq2, fl2, _ = mate(n)
    flag1 = fl1 | proper | (0x1 | 0x20 | 0x40)
    flag2 = fl2 | proper | (0x1 | 0x10 | 0x80) | 0x2000          # mate 2: isize < 0
    pos = torch.cat([fstart, fstart + isize - READ_LEN])
    isz = torch.cat([isize, -isize])
    mid = torch.cat([fstart + isize // 2, fstart + isize - READ_LEN - isize // 2])
    flags = torch.cat([flag1, flag2])
    mapq = torch.cat([mq1, mq2])
    del fstart, isize, improper, mq1, mq2, fl1, fl2, flag1, flag2, proper
    order = torch.argsort(pos, stable=True)
    pos = pos[order]
    out = dict(pos=pos, isize=isz[order], mid=mid[order], mapq=mapq[order],
               flags=flags[order].to(torch.int16))
    out["end"] = pos + READ_LEN
    out["frag_end"] = pos + out["isize"]
    return out


def synth_reference_torch(seed, length, device, block=10_000, gap_frac=0.02, lowercase_frac=0.3):
    """Synthetic contig sequence on the device (uint8): GC per block ~ Beta(8,11), N gaps,
    soft-masked blocks."""
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    nblk = (length + block - 1) // block
    rs = np.random.default_rng(seed)
    gcf = torch.from_numpy(rs.beta(8, 11, nblk).astype(np.float32)).to(device)
    lower = torch.from_numpy(rs.random(nblk) < lowercase_frac).to(device)
    gaps = torch.from_numpy(rs.random(nblk) < gap_frac).to(device)
    seq = torch.empty(length, dtype=torch.uint8, device=device)
    step = 64 * block
    lut = torch.tensor([ord("A"), ord("T"), ord("C"), ord("G")], dtype=torch.uint8, device=device)
    for s in range(0, length, step):
        e = min(length, s + step)
        b0 = s // block
        nb = (e - s + block - 1) // block
        blk = torch.arange(e - s, device=device) // block
        gcb = gcf[b0:b0 + nb][blk]
        is_gc = torch.rand(e - s, generator=g, device=device) < gcb
        pick = torch.rand(e - s, generator=g, device=device) < 0.5
        code = is_gc.to(torch.int64) * 2 + pick.to(torch.int64)
        ch = lut[code]
        ch = torch.where(lower[b0:b0 + nb][blk], ch | 0x20, ch)
        ch = torch.where(gaps[b0:b0 + nb][blk], torch.full_like(ch, ord("N")), ch)
        seq[s:e] = ch
    return seq


def soa_to_oracle_records(soa, lo=0, hi=None):
    """Device SoA slice -> host BAM-level dict for the oracle.  The CIGAR is reconstructed so
    that it reproduces the clip-fail bit and `end`: 100M, or 100S100M for clip-failing reads."""
    sl = slice(lo, hi)
    pos = soa["pos"][sl].cpu().numpy()
    isize = soa["isize"][sl].cpu().numpy()
    fl = soa["flags"][sl].cpu().numpy().view(np.uint16)
    mapq = soa["mapq"][sl].cpu().numpy()
    n = pos.size
    clipfail = (fl & 0x1000) != 0
    nops = np.where(clipfail, 2, 1).astype(np.int64)
    cigar_off = np.zeros(n + 1, np.uint64)
    np.cumsum(nops, out=cigar_off[1:])
    cigar = np.empty(int(cigar_off[-1]), np.uint32)
    first = cigar_off[:-1].astype(np.int64)
    cigar[first] = np.where(clipfail, (READ_LEN << 4) | 4, (READ_LEN << 4) | 0)
    cigar[first[clipfail] + 1] = (READ_LEN << 4) | 0
    return dict(pos=pos, isize=isize, flag=(fl & 0x0FFF).astype(np.uint16), mapq=mapq,
                cigar_off=cigar_off, cigar=cigar)


def synth_cohort_lcv(seed, T, s_lo, s_hi, device, n_contigs=22, cnv_per_sample=20):
    """Raw length-normalised coverage LCV[s_hi - s_lo][T] (f32, sample-major) of a cohort as SURVEY
    8(d) configs 3/5: target efficiency eff_t ~ LogNormal(0, 0.4) shared by all samples, per-sample
    depth factor, noise ~ N(1, 0.08), a few planted CNV segments per sample (x 0.5 / x 1.5).  Sample s
    is generated from (seed, s) alone, so any sharding of the samples yields the same cohort.
    Returns (lcv, tids int32[T])."""
    from .synth import HG38_AUTOSOMES
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    eff = torch.empty(T, device=device).log_normal_(0.0, 0.4, generator=g)
    w = np.array(HG38_AUTOSOMES[:n_contigs], np.float64)
    bounds = np.floor(np.cumsum(w) / w.sum() * T).astype(np.int64)
    tids = np.minimum(np.searchsorted(bounds, np.arange(T), side="right"), n_contigs - 1).astype(np.int32)
    out = torch.empty((s_hi - s_lo, T), dtype=torch.float32, device=device)
    for s in range(s_lo, s_hi):
        g.manual_seed(int(seed) * 1_000_003 + s + 1)
        depth = 20.0 + 20.0 * float(torch.rand(1, generator=g, device=device).item())
        row = out[s - s_lo]
        row.normal_(1.0, 0.08, generator=g)
        row.mul_(eff).mul_(depth).clamp_(min=0.0)
        if cnv_per_sample and T > 600:
            st = (torch.rand(cnv_per_sample, generator=g, device=device) * (T - 600)).long().tolist()
            ln = ((torch.rand(cnv_per_sample, generator=g, device=device) * 495).long() + 5).tolist()
            up = (torch.rand(cnv_per_sample, generator=g, device=device) < 0.5).tolist()
            for a, n, u in zip(st, ln, up):
                row[a:a + n].mul_(1.5 if u else 0.5)
    return out, torch.from_numpy(tids).to(device)
