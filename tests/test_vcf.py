"""Coverage VCF writer (csrc/vcf.cu): record layout of lib-coverage/src/lib.rs:560-673 and the tag set of
:67-166, text and BGZF containers.  Host code: runs without a GPU."""
import gzip
import struct

import numpy as np
import pytest

from cnvetti_b200 import _native as N
from cnvetti_b200 import pipeline, vcf


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
    t.gc[4] = np.frombuffer(np.uint32(N.F32_MISSING_BITS).tobytes(), np.float32)[0]
    t.gap = np.zeros(n, np.uint8)
    t.gap[4] = 1
    if normalised:
        t.cv = (t.lcv * 1.7).astype(np.float32)
        t.cv2 = np.log2(np.maximum(t.cv, 1e-6)).astype(np.float32)
        t.norm_flags = np.full(n, N.N2_CV | N.N2_CV2, np.uint8)
        t.norm_flags[5] = N.N2_FEW_GCWINDOWS          # no CV for this window (correct_binned.rs:99-102)
        t.norm_flags[6] = N.N2_CV                     # CV = 0: CV2 missing
    return t


@pytest.mark.parametrize("ext", [".vcf", ".vcf.gz"])
def test_vcf_records_and_header(tmp_path, ext):
    t = _table()
    contigs = [("chr1", 3123), ("chr2", 2500), ("chrX", 999)]
    path = vcf.write_coverage_vcf(str(tmp_path / ("cov" + ext)), t, "NA12878", contigs, command_line="cnvetti cmd coverage -i x.bam",
                                  file_date="20181115", threads=3)
    r = vcf.read_coverage_vcf(path)
    h = "\n".join(r["header"])
    assert r["header"][0] == "##fileformat=VCFv4.2" and "##fileDate=20181115" in h
    assert "##cnvetti_cmdCoverageCommand=cnvetti cmd coverage -i x.bam" in h
    assert "##contig=<ID=chr2,length=2500>" in h and "##contig=<ID=chrX,length=999>" in h
    for tag in ("INFO=<ID=END,Number=1,Type=Integer", "INFO=<ID=GC,Number=1,Type=Float", "INFO=<ID=GAP,Number=0,Type=Flag",
                "INFO=<ID=MEAN_MAPQ,Number=1,Type=Float", "INFO=<ID=FEW_GCWINDOWS,Number=0,Type=Flag",
                "FILTER=<ID=PILE", "FILTER=<ID=LOWCOV", "FILTER=<ID=NO_REF", "ALT=<ID=WINDOW", "ALT=<ID=TARGET"):
        assert "##" + tag in h, tag
    for key in ("GT", "MQ", "RCV", "RCVSD", "LCV", "LCVSD", "NCV", "CV", "CVSD", "CV2", "CVZ", "FRM", "FT"):
        assert "##FORMAT=<ID=%s," % key in h, key
    assert r["samples"] == ["NA12878"]
    n = t.offsets[-1]
    assert r["chrom"] == ["chr1"] * 7 + ["chr2"] * 5
    assert np.array_equal(r["pos"], t.start) and r["alt"] == ["<WINDOW>"] * n
    assert r["id"][6] == "chr1:3001-3123" and r["id"][7] == "chr2:1-500"          # lib.rs:586-588
    assert [int(x) for x in r["info"]["END"]] == t.end.tolist()
    assert r["info"]["GAP"][4] is True and r["info"]["GAP"][3] is None
    assert r["info"]["GC"][4] == "." and all(x == "60" for x in r["info"]["MEAN_MAPQ"])
    assert r["info"]["FEW_GCWINDOWS"][5] is True and r["info"]["FEW_GCWINDOWS"][4] is None
    f = r["format"]
    assert f["GT"] == ["."] * n
    assert f["FT"][2] == "PILE" and f["FT"][3] == "LOWCOV;PILE" and f["FT"][0] is None   # only when non-empty
    np.testing.assert_allclose(f["RCV"], t.rcv, rtol=1e-5)
    np.testing.assert_allclose(f["LCV"], t.lcv, rtol=1e-5)       # %g: six significant digits
    np.testing.assert_allclose(f["FRM"], t.frm, rtol=1e-5)
    has_cv = (t.norm_flags & N.N2_CV) != 0
    np.testing.assert_allclose(f["CV"][has_cv], t.cv[has_cv], rtol=1e-5)
    miss = np.uint32(N.F32_MISSING_BITS)
    assert f["CV"].view(np.uint32)[5] == miss and f["CV2"].view(np.uint32)[6] == miss and f["CV2"].view(np.uint32)[5] == miss
    assert "RCVSD" not in f
    if ext == ".vcf.gz":  # BGZF: every member carries the BC subfield, the file ends with the 28-byte EOF block
        raw = open(path, "rb").read()
        assert raw[:4] == b"\x1f\x8b\x08\x04" and raw[12:14] == b"BC"
        assert raw[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
        bsize = struct.unpack("<H", raw[16:18])[0] + 1
        assert raw[bsize:bsize + 4] == b"\x1f\x8b\x08\x04" or bsize == len(raw) - 28


def test_vcf_coverage_kind_targets_and_errors(tmp_path):
    t = _table(coverage_kind=True, normalised=False)
    path = vcf.write_coverage_vcf(str(tmp_path / "t.vcf"), t, "S", [("chr1", 3123), ("chr2", 2500)], is_targets=True)
    r = vcf.read_coverage_vcf(path)
    assert r["alt"][0] == "<TARGET>" and "RCVSD" in r["format"] and "FRM" not in r["format"] and "CV" not in r["format"]
    line = [l for l in open(path) if not l.startswith("#")][0].rstrip("\n").split("\t")
    assert line[8] == "GT:RCV:RCVSD:LCV:LCVSD:MQ"                                     # push order of lib.rs:616-657
    with pytest.raises(N.CnvbError, match="BCF"):
        vcf.write_coverage_vcf(str(tmp_path / "t.bcf"), t, "S", [("chr1", 3123), ("chr2", 2500)])
    with pytest.raises(N.CnvbError, match="Could not open"):
        vcf.write_coverage_vcf(str(tmp_path / "no" / "dir.vcf"), t, "S", [("chr1", 3123), ("chr2", 2500)])


def test_vcf_large_file_many_blocks(tmp_path):
    """More rows than one formatting batch thread share and many BGZF blocks: content survives the seams."""
    n = 60_000
    t = pipeline.CoverageTable()
    t.names, t.offsets = ["1"], [0, n]
    t.start = (np.arange(n) * 1000).astype(np.int32)
    t.end = (t.start + 1000).astype(np.int32)
    t.rcv = (np.arange(n) % 977).astype(np.float32)
    t.lcv = (t.rcv / 1000).astype(np.float32)
    path = vcf.write_coverage_vcf(str(tmp_path / "big.vcf.gz"), t, "S", [("1", n * 1000)], threads=4)
    with gzip.open(path, "rt") as f:
        rows = [l for l in f if not l.startswith("#")]
    assert len(rows) == n
    assert rows[0].split("\t")[1] == "1" and rows[-1].split("\t")[2] == "1:%d-%d" % ((n - 1) * 1000 + 1, n * 1000)
    got = np.array([float(l.rstrip("\n").split("\t")[9].split(":")[1]) for l in rows], np.float32)
    assert np.array_equal(got, t.rcv)
