"""VCF / BCF reader and writer with CSI / TBI index (csrc/bcf.cu, files.py): record layout of
lib-coverage/src/lib.rs:560-673 and lib-model-wis/src/lib.rs:340-419, the containers and index of
lib-shared/src/bcf_utils.rs:47-82.  The binary encoding is checked against an independent BCF2.2 / CSI / tabix
decoder written from the specifications (below), not only by reading back with the library.  Host code: no GPU."""
import gzip
import os
import struct
import zlib

import numpy as np
import pytest

from cnvetti_b200 import _native as N
from cnvetti_b200 import files, pipeline

MISS = np.uint32(N.F32_MISSING_BITS)


def _table(n1=7, n2=5, coverage_kind=False, normalised=True):
    rng = np.random.default_rng(3)
    t = pipeline.CoverageTable()
    t.names, t.offsets = ["chr1", "chr2"], [0, n1, n1 + n2]
    n = n1 + n2
    t.start = np.concatenate([np.arange(n1) * 500, np.arange(n2) * 500]).astype(np.int32)
    t.end = (t.start + 500).astype(np.int32)
    t.end[n1 - 1] = t.start[n1 - 1] + 123            # last window of a contig is shorter
    t.rcv = rng.integers(0, 400, n).astype(np.float32)
    t.lcv = (t.rcv / (t.end - t.start).astype(np.float32)).astype(np.float32)
    if coverage_kind:
        t.rcvsd = rng.random(n).astype(np.float32)
        t.lcvsd = (t.rcvsd / 500).astype(np.float32)
    else:
        t.frm = np.zeros(n, np.float32)
        t.frm[2] = 0.75
    t.ft = np.zeros(n, np.uint8)
    t.ft[2] = N.FT_PILE
    t.ft[3] = N.FT_LOWCOV | N.FT_PILE
    t.gc = rng.random(n).astype(np.float32)
    t.gc[4] = np.frombuffer(MISS.tobytes(), np.float32)[0]
    t.gap = np.zeros(n, np.uint8)
    t.gap[4] = 1
    if normalised:
        t.cv = (t.lcv * 1.7).astype(np.float32)
        t.cv2 = np.log2(np.maximum(t.cv, 1e-6)).astype(np.float32)
        t.norm_flags = np.full(n, N.N2_CV | N.N2_CV2, np.uint8)
        t.norm_flags[5] = N.N2_FEW_GCWINDOWS          # no CV for this window (correct_binned.rs:99-102)
        t.norm_flags[6] = N.N2_CV                     # CV = 0: CV2 missing
    return t


CONTIGS = [("chr1", 3123), ("chr2", 2500), ("chrX", 999)]


# ---- an independent decoder: BGZF, BCF2.2 records, CSI, TBI -----------------------------------------------
def bgzf_blocks(raw):
    """[(file offset, uncompressed bytes)] of a BGZF file"""
    out, at = [], 0
    while at < len(raw):
        assert raw[at:at + 4] == b"\x1f\x8b\x08\x04" and raw[at + 12:at + 14] == b"BC"
        bsize = struct.unpack("<H", raw[at + 16:at + 18])[0] + 1
        data = zlib.decompress(raw[at + 18:at + bsize - 8], -15)
        crc, isize = struct.unpack("<II", raw[at + bsize - 8:at + bsize])
        assert isize == len(data) and crc == zlib.crc32(data)
        out.append((at, data))
        at += bsize
    return out


def typed(buf, at):
    b = buf[at]
    at += 1
    t, ln = b & 0xf, b >> 4
    if ln == 15:
        (ln,), at = typed(buf, at)
    fmt = {1: "b", 2: "h", 3: "i", 5: "f", 7: "c", 0: ""}[t]
    if t == 7:
        return buf[at:at + ln].decode(), at + ln
    if t == 0:
        return (), at
    size = struct.calcsize(fmt)
    raw = buf[at:at + ln * size]
    vals = struct.unpack("<%d%s" % (ln, fmt), raw)
    if t == 5:
        vals = tuple(np.frombuffer(raw, np.uint32).tolist())   # floats as bit patterns
    return vals, at + ln * size


def bcf_decode(data):
    """(header text, [record dict]) from the uncompressed bytes of a BCF file"""
    assert data[:5] == b"BCF\x02\x02"
    l_text = struct.unpack("<I", data[5:9])[0]
    text = data[9:9 + l_text]
    assert text.endswith(b"\x00")
    text = text.rstrip(b"\x00").decode()
    dic = {}
    contigs = {}
    for line in text.splitlines():
        if line.startswith(("##INFO", "##FORMAT", "##FILTER")):
            name = line.split("ID=")[1].split(",")[0]
            dic[int(line.rsplit("IDX=", 1)[1].rstrip(">"))] = name
        if line.startswith("##contig"):
            contigs[int(line.rsplit("IDX=", 1)[1].rstrip(">"))] = line.split("ID=")[1].split(",")[0]
    recs, at = [], 9 + l_text
    while at < len(data):
        l_shared, l_indiv = struct.unpack("<II", data[at:at + 8])
        s = data[at + 8:at + 8 + l_shared]
        chrom, pos, rlen, qual, nai, nfs = struct.unpack("<iiiIII", s[:24])
        n_allele, n_info, n_fmt, n_sample = nai >> 16, nai & 0xffff, nfs >> 24, nfs & 0xffffff
        p = 24
        rid, p = typed(s, p)
        alleles = []
        for _ in range(n_allele):
            a, p = typed(s, p)
            alleles.append(a)
        filt, p = typed(s, p)
        info = []
        for _ in range(n_info):
            (k,), p = typed(s, p)
            v, p = typed(s, p)
            info.append((dic[k], v))
        assert p == l_shared
        ind = data[at + 8 + l_shared:at + 8 + l_shared + l_indiv]
        p, fmt = 0, []
        for _ in range(n_fmt):
            (k,), p = typed(ind, p)
            b = ind[p]
            p += 1
            t, ln = b & 0xf, b >> 4
            assert ln < 15
            size = {1: 1, 2: 2, 3: 4, 5: 4, 7: 1}[t]
            vals = []
            for _s in range(n_sample):
                raw = ind[p:p + ln * size]
                p += ln * size
                vals.append(raw.rstrip(b"\x00").decode() if t == 7 else
                            np.frombuffer(raw, {1: np.int8, 2: np.int16, 3: np.int32, 5: np.uint32}[t]).tolist())
            fmt.append((dic[k], t, vals))
        assert p == l_indiv
        recs.append(dict(chrom=contigs[chrom], pos=pos, rlen=rlen, qual=qual, id=rid, alleles=alleles,
                         filters=[dic[i] for i in filt], info=info, fmt=fmt, n_sample=n_sample, offset=at))
        at += 8 + l_shared + l_indiv
    return text, recs


def reg2bins(beg, end, min_shift, depth):
    end -= 1
    bins, s, t = [], min_shift + depth * 3, 0
    for l in range(depth + 1):
        b, e = t + (beg >> s), t + (end >> s)
        bins += list(range(b, e + 1))
        s -= 3
        t += 1 << (l * 3)
    return bins


def parse_index(path):
    """CSI / TBI -> dict(min_shift, depth, names, refs=[{bin: [(beg, end)]}], loff / ioff)"""
    d = b"".join(x[1] for x in bgzf_blocks(open(path, "rb").read()))
    out = {}
    if d[:4] == b"CSI\x01":
        out["min_shift"], out["depth"], l_aux = struct.unpack("<iii", d[4:16])
        at = 16 + l_aux
        csi = True
    else:
        assert d[:4] == b"TBI\x01"
        n_ref, fmt, cs, cb, ce, meta, skip, l_nm = struct.unpack("<8i", d[4:36])
        assert (fmt, cs, cb, ce, meta, skip) == (2, 1, 2, 0, ord("#"), 0)
        out["names"] = d[36:36 + l_nm].rstrip(b"\x00").decode().split("\x00")
        out["min_shift"], out["depth"] = 14, 5
        at = 36 + l_nm - 4
        csi = False
    if csi:
        n_ref = struct.unpack("<i", d[at:at + 4])[0]
    at += 4
    refs = []
    for _ in range(n_ref):
        n_bin = struct.unpack("<i", d[at:at + 4])[0]
        at += 4
        bins, loff = {}, {}
        for _b in range(n_bin):
            b = struct.unpack("<I", d[at:at + 4])[0]
            at += 4
            if csi:
                loff[b] = struct.unpack("<Q", d[at:at + 8])[0]
                at += 8
            n_chunk = struct.unpack("<i", d[at:at + 4])[0]
            at += 4
            bins[b] = [struct.unpack("<QQ", d[at + 16 * i:at + 16 * i + 16]) for i in range(n_chunk)]
            at += 16 * n_chunk
        ioff = None
        if not csi:
            n_intv = struct.unpack("<i", d[at:at + 4])[0]
            at += 4
            ioff = struct.unpack("<%dQ" % n_intv, d[at:at + 8 * n_intv])
            at += 8 * n_intv
        refs.append(dict(bins=bins, loff=loff, ioff=ioff))
    assert len(d) - at in (0, 8)
    out["refs"] = refs
    return out


# ---- tests ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ext", [".vcf", ".vcf.gz", ".bcf", ".raw"])
def test_coverage_records_round_trip(tmp_path, ext):
    t = _table()
    path = files.write_coverage_file(str(tmp_path / ("cov" + ext)), t, ["NA12878"], CONTIGS,
                                     command_line="cnvetti cmd coverage -i x.bam", file_date="20181115", threads=3)
    vf = files.VariantFile(path, threads=2)
    h = "\n".join(vf.header)
    assert vf.header[0] == "##fileformat=VCFv4.2" and "##fileDate=20181115" in h
    assert "##cnvetti_cmdCoverageCommand=cnvetti cmd coverage -i x.bam" in h and "IDX=" not in h
    assert vf.contigs == CONTIGS and vf.samples == ["NA12878"] and vf.is_bcf == (ext in (".bcf", ".raw"))
    for tag in ("INFO=<ID=END,Number=1,Type=Integer", "INFO=<ID=GC,Number=1,Type=Float", "INFO=<ID=GAP,Number=0,Type=Flag",
                "INFO=<ID=MEAN_MAPQ,Number=1,Type=Float", "INFO=<ID=FEW_GCWINDOWS,Number=0,Type=Flag",
                "FILTER=<ID=PILE", "FILTER=<ID=LOWCOV", "FILTER=<ID=NO_REF", "ALT=<ID=WINDOW", "ALT=<ID=TARGET"):
        assert "##" + tag in h, tag
    for key in ("GT", "MQ", "RCV", "RCVSD", "LCV", "LCVSD", "NCV", "CV", "CVSD", "CV2", "CVZ", "FRM", "FT"):
        assert "##FORMAT=<ID=%s," % key in h, key
    n = t.offsets[-1]
    fx = vf.fixed()
    assert fx["rid"].tolist() == [0] * 7 + [1] * 5 and np.array_equal(fx["pos"], t.start) and np.array_equal(fx["end"], t.end)
    assert (fx["alt"] == 0).all() and (fx["filters"] == 0).all() and fx["canonical_ids"]       # ID chrom:pos+1-end (lib.rs:586-588)
    assert vf.info_flag("GAP").tolist() == t.gap.tolist()
    assert vf.info_flag("FEW_GCWINDOWS").tolist() == [1 if i == 5 else 0 for i in range(n)]
    mq, p = vf.info_float("MEAN_MAPQ")
    assert p.all() and (mq == 60.0).all()
    gc, p = vf.info_float("GC")
    assert p.all() and gc.view(np.uint32)[4] == MISS
    ft, p = vf.format_ft()
    assert ft[:, 0].tolist() == t.ft.tolist() and p.tolist() == (t.ft != 0).astype(int).tolist()   # only when non-empty
    exact = vf.is_bcf
    for key, col in (("RCV", t.rcv), ("LCV", t.lcv), ("FRM", t.frm)):
        v, p = vf.format_float(key)
        assert p.all()
        if exact:
            assert np.array_equal(v[:, 0].view(np.uint32), col.view(np.uint32)), key
        else:
            np.testing.assert_allclose(v[:, 0], col, rtol=1e-5)      # text: six significant digits
    cv, p = vf.format_float("CV")
    has_cv = (t.norm_flags & N.N2_CV) != 0
    assert np.array_equal(p != 0, has_cv) and cv.view(np.uint32)[5, 0] == MISS
    np.testing.assert_allclose(cv[has_cv, 0], t.cv[has_cv], rtol=1e-5)
    cv2, p = vf.format_float("CV2")
    assert np.array_equal(p != 0, (t.norm_flags & N.N2_CV2) != 0) and cv2.view(np.uint32)[6, 0] == MISS
    assert not vf.format_float("RCVSD")[1].any()
    t2 = files.read_coverage_table(vf)
    assert t2.names == t.names and t2.offsets == t.offsets and t2.rcvsd is None and not t2.is_targets
    assert np.array_equal(t2.norm_flags, t.norm_flags) and np.array_equal(t2.gap, t.gap)
    with pytest.raises(N.CnvbError, match="not defined in the header"):
        vf.format_float("XYZ")
    assert os.path.exists(path + ".tbi") == (ext == ".vcf.gz") and os.path.exists(path + ".csi") == (ext == ".bcf")


def test_vcf_text_layout(tmp_path):
    t = _table(coverage_kind=True, normalised=False)
    path = files.write_coverage_file(str(tmp_path / "t.vcf"), t, ["S"], CONTIGS[:2], is_targets=True)
    rows = [l.rstrip("\n").split("\t") for l in open(path) if not l.startswith("#")]
    assert rows[0][:8] == ["chr1", "1", "chr1:1-500", "N", "<TARGET>", ".", ".", rows[0][7]]
    assert rows[0][8] == "GT:RCV:RCVSD:LCV:LCVSD:MQ"                                     # push order of lib.rs:616-657
    assert rows[2][8] == "GT:RCV:RCVSD:LCV:LCVSD:MQ:FT" and rows[2][9].endswith(":60:PILE")
    assert rows[3][9].endswith(":LOWCOV;PILE")                                             # joined with ';' (lib.rs:659-664)
    assert rows[4][7].startswith("END=2500;GAP;GC=.;MEAN_MAPQ=60")                           # GAP before GC (lib.rs:597-606)
    assert rows[6][2] == "chr1:3001-3123"
    header = [l for l in open(path) if l.startswith("#")]
    assert header[-1].rstrip("\n").split("\t")[8:] == ["FORMAT", "S"]
    with pytest.raises(N.CnvbError, match="Could not open"):
        files.write_coverage_file(str(tmp_path / "no" / "dir.vcf"), t, ["S"], CONTIGS[:2])
    with pytest.raises(N.CnvbError, match="Could not find contig"):
        files.write_coverage_file(str(tmp_path / "x.vcf"), t, ["S"], [("chr1", 3123)])
    with pytest.raises(N.CnvbError, match="Could not open BCF file"):
        files.VariantFile(str(tmp_path / "missing.bcf"))


def test_bcf_binary_encoding_against_the_specification(tmp_path):
    """An independent BCF2.2 decoder reads what the writer wrote: field by field."""
    t = _table()
    path = files.write_coverage_file(str(tmp_path / "cov.bcf"), t, ["S1"], CONTIGS, file_date="20200101")
    raw = open(path, "rb").read()
    assert raw[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")   # BGZF EOF block
    data = b"".join(x[1] for x in bgzf_blocks(raw))
    text, recs = bcf_decode(data)
    assert text.splitlines()[0] == "##fileformat=VCFv4.2" and "##FILTER=<ID=PASS,Description=\"All filters passed\",IDX=0>" in text
    assert text.splitlines()[-1] == "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1"
    assert len(recs) == 12
    f32 = lambda x: int(np.float32(x).view(np.uint32))
    for i, r in enumerate(recs):
        assert r["chrom"] == ("chr1" if i < 7 else "chr2") and r["pos"] == t.start[i] and r["rlen"] == t.end[i] - t.start[i]
        assert r["qual"] == int(MISS) and r["alleles"] == ["N", "<WINDOW>"] and r["filters"] == [] and r["n_sample"] == 1
        assert r["id"] == "%s:%d-%d" % (r["chrom"], t.start[i] + 1, t.end[i])
        keys = [k for k, _ in r["info"]]
        want = ["END"] + (["GAP"] if t.gap[i] else []) + ["GC", "MEAN_MAPQ"] + (["FEW_GCWINDOWS"] if i == 5 else [])
        assert keys == want
        info = dict(r["info"])
        assert info["END"] == (int(t.end[i]),) and info["GC"] == (int(t.gc[i:i + 1].view(np.uint32)[0]),) and info["MEAN_MAPQ"] == (f32(60),)
        if t.gap[i]:
            assert info["GAP"] == ()
        fk = [k for k, _, _ in r["fmt"]]
        want = ["GT", "RCV", "LCV", "FRM", "MQ"] + (["FT"] if t.ft[i] else []) + \
               (["CV"] if t.norm_flags[i] & N.N2_CV else []) + (["CV2"] if t.norm_flags[i] & N.N2_CV2 else [])
        assert fk == want, (i, fk)
        fm = {k: (ty, v) for k, ty, v in r["fmt"]}
        assert fm["GT"] == (1, [[0]])                           # push_format_integer(b"GT", &[GT_MISSING]) (lib.rs:613-615)
        assert fm["RCV"] == (5, [[f32(t.rcv[i])]]) and fm["LCV"] == (5, [[int(t.lcv[i:i + 1].view(np.uint32)[0])]])
        if t.ft[i]:
            assert fm["FT"] == (7, ["LOWCOV;PILE" if t.ft[i] == 3 else "PILE"])
    # END = 3123 needs int16, 500 as well; small values int8 (smallest type that holds the value)
    one = data[recs[0]["offset"]:]
    assert one[8 + 24 + 1 + len("chr1:1-500") + 2 + 1 + len("<WINDOW>") + 1:][:5] == bytes([0x11, 1, 0x12]) + struct.pack("<h", 500)


@pytest.mark.parametrize("ext", [".bcf", ".vcf.gz"])
def test_index_finds_every_overlapping_record(tmp_path, ext):
    """CSI (min_shift 14, depth from the longest contig) / TBI: for random regions the chunks of the overlapping
    bins, read from their virtual offsets, contain every overlapping record; linear offsets never overshoot."""
    rng = np.random.default_rng(11)
    contigs = [("1", 248_956_422), ("2", 242_193_529), ("MT", 16_569)]
    t = pipeline.CoverageTable()
    t.names = ["1", "2", "MT"]
    starts, ends = [], []
    for name, L in contigs:
        n = 4000 if L > 1e6 else 30
        s = np.sort(rng.integers(0, L - 2000, n)).astype(np.int32)
        starts.append(s)
        ends.append((s + rng.integers(50, 1500, n)).astype(np.int32))
    t.offsets = np.concatenate([[0], np.cumsum([len(s) for s in starts])]).tolist()
    t.start, t.end = np.concatenate(starts), np.concatenate(ends)
    n = t.offsets[-1]
    t.rcv = rng.integers(0, 100, n).astype(np.float32)
    t.lcv = (t.rcv / (t.end - t.start)).astype(np.float32)
    t.ft = np.zeros(n, np.uint8)
    path = files.write_coverage_file(str(tmp_path / ("x" + ext)), t, ["S"], contigs, is_targets=True, threads=4)
    idx = parse_index(path + (".csi" if ext == ".bcf" else ".tbi"))
    assert idx["min_shift"] == 14 and idx["depth"] == (5 if ext == ".vcf.gz" else 5)   # 249 Mbp + 256 <= 2^29
    if ext == ".vcf.gz":
        assert idx["names"] == ["1", "2", "MT"]
    raw = open(path, "rb").read()
    blocks = bgzf_blocks(raw)
    addr = {a: i for i, (a, _) in enumerate(blocks)}
    ublock = np.concatenate([[0], np.cumsum([len(d) for _, d in blocks])])
    data = b"".join(d for _, d in blocks)
    v2u = lambda v: int(ublock[addr[v >> 16]] + (v & 0xffff))
    if ext == ".bcf":
        _, recs = bcf_decode(data)
        spans = [(v2u_off, r["chrom"], r["pos"], r["pos"] + r["rlen"]) for r in recs for v2u_off in [r["offset"]]]
    else:
        spans, at = [], 0
        for line in data.split(b"\n"):
            if line and not line.startswith(b"#"):
                c = line.split(b"\t")
                end = int(c[7].split(b"END=")[1].split(b";")[0])
                spans.append((at, c[0].decode(), int(c[1]) - 1, end))
            at += len(line) + 1
    meta = ((1 << (3 * idx["depth"] + 3)) - 1) // 7 + 1
    name_to_ref = {c[0]: i for i, c in enumerate(contigs)}
    for ref in range(3):
        bins = idx["refs"][ref]["bins"]
        assert meta in bins and bins[meta][1][0] == len(starts[ref])       # record count in the pseudo-bin
        mine = [s for s in spans if name_to_ref[s[1]] == ref]
        assert v2u(bins[meta][0][0]) == mine[0][0]
        for _ in range(40):
            L = contigs[ref][1]
            a = int(rng.integers(0, L - 1))
            b = min(L, a + int(rng.integers(1, 3_000_000)))
            want = {s[0] for s in mine if s[2] < b and s[3] > a}
            cand = set()
            for bn in reg2bins(a, b, idx["min_shift"], idx["depth"]):
                for cb, ce in bins.get(bn, []):
                    lo, hi = v2u(cb), v2u(ce)
                    cand |= {s[0] for s in mine if lo <= s[0] < hi}
            assert want <= cand, (ref, a, b)
            # linear index: no overlapping record starts before the offset of the window holding `a`
            w = a >> 14
            if ext == ".vcf.gz":
                ioff = idx["refs"][ref]["ioff"]
                if w < len(ioff) and want:
                    assert v2u(ioff[w]) <= min(want)
            elif want:
                for bn in reg2bins(a, b, 14, idx["depth"]):
                    if bn in bins and bn != meta:
                        in_bin = [s[0] for s in mine if s[0] in want and any(v2u(cb) <= s[0] < v2u(ce) for cb, ce in bins[bn])]
                        if in_bin:
                            assert v2u(idx["refs"][ref]["loff"][bn]) <= min(in_bin)


def test_model_file_round_trip(tmp_path):
    """write_output of lib-model-wis (src/lib.rs:340-419): FILTER FEW_REF / UNRELIABLE, INFO END / NUM_CALLS /
    REF_TARGETS (names in distance order), no samples -- and load_target_infos' reading of it (lib-mod-cov :89-164)."""
    contigs = [("chr1", 100_000), ("chr2", 50_000)]
    n = 9
    rid = np.array([0] * 5 + [1] * 4, np.int32)
    pos = (np.arange(n) * 1000 + 7).astype(np.int32)
    end = (pos + 300).astype(np.int32)
    lists = [[5, 6, 8], [7], [], [8, 5, 6, 7], [6], [0, 1, 4], [3, 2], [], [0]]
    off = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.uint64)
    idx = np.array([j for x in lists for j in x], np.uint32)
    filters = np.zeros(n, np.uint8)
    filters[2] = N.VF_FEW_REF
    filters[7] = N.VF_FEW_REF | N.VF_UNRELIABLE
    filters[4] = N.VF_UNRELIABLE
    calls = np.array([0, 1, 0, 130, 40000, 3, 0, 9, 2], np.int32)
    for ext in (".bcf", ".vcf.gz", ".vcf"):
        path = files.write_variant_file(str(tmp_path / ("m" + ext)), files.model_header(contigs, "cmd", "20200101"), [], rid, pos,
                                        end, is_targets=True, filters=filters, num_calls=calls, ref_off=off, ref_idx=idx)
        with files.VariantFile(path) as vf:
            assert vf.samples == [] and vf.n == n and "##cnvetti_cmdBuildModelWisCommand=cmd" in vf.header
            fx = vf.fixed()
            assert (fx["alt"] == 1).all() and fx["filters"].tolist() == filters.tolist() and fx["canonical_ids"]
            assert vf.info_int("NUM_CALLS")[0].tolist() == calls.tolist()
            o2, i2 = vf.ref_targets()
            assert o2.tolist() == off.tolist() and i2.tolist() == idx.tolist()
            assert vf.match_ids(vf).tolist() == list(range(n))
    line = [l for l in open(str(tmp_path / "m.vcf")) if l.startswith("chr1\t8\t")][0].rstrip("\n").split("\t")
    assert len(line) == 8 and line[6] == "." and line[7] == "END=307;NUM_CALLS=0;REF_TARGETS=chr2:5008-5307,chr2:6008-6307,chr2:8008-8307"
    line = [l for l in open(str(tmp_path / "m.vcf")) if l.startswith("chr2\t7008\t")][0].split("\t")
    assert line[6] == "FEW_REF;UNRELIABLE" and line[7].strip() == "END=7307;NUM_CALLS=9"      # empty list: no REF_TARGETS
    _, recs = bcf_decode(b"".join(x[1] for x in bgzf_blocks(open(str(tmp_path / "m.bcf"), "rb").read())))
    assert recs[7]["filters"] == ["FEW_REF", "UNRELIABLE"] and recs[7]["n_sample"] == 0 and recs[7]["fmt"] == []
    assert dict(recs[4]["info"])["NUM_CALLS"] == (40000,) and dict(recs[3]["info"])["NUM_CALLS"] == (130,)
    assert dict(recs[3]["info"])["REF_TARGETS"] == "chr2:8008-8307,chr2:5008-5307,chr2:6008-6307,chr2:7008-7307"
    # names resolved against another file (the sample's coverage file); a name it lacks is the reference's panic
    t = pipeline.CoverageTable()
    t.names, t.offsets = ["chr1", "chr2"], [0, 5, 8]
    keep = np.array([0, 1, 2, 3, 4, 5, 6, 7])
    t.start, t.end = pos[keep], end[keep]
    t.rcv = np.ones(8, np.float32)
    t.lcv = np.ones(8, np.float32)
    t.ft = np.zeros(8, np.uint8)
    cov = files.write_coverage_file(str(tmp_path / "cov.bcf"), t, ["S"], contigs, is_targets=True)
    with files.VariantFile(str(tmp_path / "m.bcf")) as mf, files.VariantFile(cov) as cf:
        assert mf.match_ids(cf).tolist() == list(range(8)) + [0xFFFFFFFF]
        with pytest.raises(N.ReferencePanic, match="Could not find: chr2:8008-8307"):
            mf.ref_targets(against=cf)


def test_multi_sample_file_and_many_blocks(tmp_path):
    """A merged file (lib-merge-cov/src/lib.rs:102-232): S sample columns per FORMAT tag; more rows than one
    formatting batch per thread and many BGZF blocks -- content survives the seams."""
    rng = np.random.default_rng(5)
    n, S = 40_000, 5
    samples = ["s%d" % i for i in range(S)]
    rid = np.zeros(n, np.int32)
    pos = (np.arange(n) * 1000).astype(np.int32)
    end = (pos + 1000).astype(np.int32)
    cv = rng.random((n, S)).astype(np.float32)
    cv.view(np.uint32)[17, 2] = MISS
    lcv = rng.random((n, S)).astype(np.float32)
    present = np.ones(n, np.uint8)
    present[100:200] = 0
    ft = np.zeros((n, S), np.uint8)
    ft[5, 1] = N.FT_PILE
    ft[5, 3] = N.FT_LOWCOV
    for ext in (".bcf", ".vcf.gz"):
        path = files.write_variant_file(str(tmp_path / ("merged" + ext)), files.coverage_header([("1", n * 1000)]), samples, rid,
                                        pos, end, is_targets=True, fmt=[("LCV", lcv, None), ("CV", cv, present)], ft=ft, ft_pos=1,
                                        threads=4)
        with files.VariantFile(path, threads=3) as vf:
            assert vf.samples == samples and vf.n == n
            g, p = vf.format_float("CV")
            assert np.array_equal(p, present)
            if ext == ".bcf":
                assert np.array_equal(g[present != 0].view(np.uint32), cv[present != 0].view(np.uint32))
                assert np.array_equal(vf.format_float("LCV")[0].view(np.uint32), lcv.view(np.uint32))
            else:
                np.testing.assert_allclose(g[present != 0][1:], cv[present != 0][1:], rtol=1e-5)
            assert (g[present == 0].view(np.uint32) == MISS).all() and g.view(np.uint32)[17, 2] == MISS
            f2, fp = vf.format_ft()
            assert np.array_equal(f2, ft) and fp.sum() == 1
    # every BGZF block is checked against its CRC-32 (the library's own inflater first, zlib behind it): a flipped
    # payload bit that still inflates to the right length, and one in the stored checksum, are both refused
    raw = bytearray(open(str(tmp_path / "merged.bcf"), "rb").read())
    bsize = struct.unpack_from("<H", raw, 16)[0] + 1
    for at in (bsize - 8, bsize + 18 + 4000):   # the first block's CRC field; deflate data of the second block
        bad = bytearray(raw)
        bad[at] ^= 0x10
        open(str(tmp_path / "bad.bcf"), "wb").write(bytes(bad))
        with pytest.raises(N.CnvbError, match="inflate|CRC"):
            files.VariantFile(str(tmp_path / "bad.bcf"), threads=2)
    with gzip.open(str(tmp_path / "merged.vcf.gz"), "rt") as f:
        rows = [l for l in f if not l.startswith("#")]
    assert len(rows) == n and rows[5].split("\t")[8] == "GT:LCV:FT:CV"
    assert [c.split(":")[2] for c in rows[5].rstrip("\n").split("\t")[9:]] == [".", "PILE", ".", "LOWCOV", "."]
    assert rows[-1].split("\t")[2] == "1:%d-%d" % ((n - 1) * 1000 + 1, n * 1000)


def test_reader_takes_foreign_layouts(tmp_path):
    """Files not written by this library: explicit IDX in another order, PASS filters, unknown tags, a plain
    (non-BGZF) gzip stream, CRLF-free odd spacing of samples."""
    text = ("##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n"
            "##contig=<ID=chrA,length=5000>\n##contig=<ID=chrB,length=100>\n"
            "##INFO=<ID=FOO,Number=1,Type=String,Description=\"x, with a comma\">\n"
            "##INFO=<ID=END,Number=1,Type=Integer,Description=\"e\">\n##INFO=<ID=GC,Number=1,Type=Float,Description=\"g\">\n"
            "##FILTER=<ID=q10,Description=\"other\">\n##FILTER=<ID=NO_REF,Description=\"n\">\n"
            "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"g\">\n##FORMAT=<ID=CV,Number=1,Type=Float,Description=\"c\">\n"
            "##FORMAT=<ID=XX,Number=2,Type=Integer,Description=\"c\">\n"
            "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ta\tb\n"
            "chrA\t11\tchrA:11-20\tN\t<TARGET>\t.\tPASS\tFOO=bar;END=20;GC=0.25\tGT:XX:CV\t.:1,2:0.5\t./.:3,4:1.5e-3\n"
            "chrA\t31\tweird\tNNN\t<DEL>\t9\tq10;NO_REF\tFOO=b\tGT\t.\t.\n"
            "chrB\t5\tchrB:5-9\tN\t<WINDOW>\t.\t.\tEND=9\tGT:CV\t.:.\t.\n")
    p1 = str(tmp_path / "f.vcf")
    open(p1, "w").write(text)
    p2 = str(tmp_path / "f.vcf.gz")
    with gzip.open(p2, "wt") as f:
        f.write(text)
    for p in (p1, p2):
        with files.VariantFile(p) as vf:
            fx = vf.fixed()
            assert vf.samples == ["a", "b"] and vf.n == 3 and not fx["canonical_ids"]
            assert fx["rid"].tolist() == [0, 0, 1] and fx["pos"].tolist() == [10, 30, 4] and fx["end"].tolist() == [20, 33, 9]
            assert fx["alt"].tolist() == [1, 2, 0]
            assert fx["filters"].tolist() == [N.VF_PASS, N.VF_OTHER | N.VF_NO_REF, 0]
            cv, pr = vf.format_float("CV")
            assert pr.tolist() == [1, 0, 1] and cv[0].tolist() == [0.5, np.float32(1.5e-3)]
            assert (cv[1:].view(np.uint32) == MISS).all()
            gc, pr = vf.info_float("GC")
            assert pr.tolist() == [1, 0, 0] and gc[0] == 0.25
