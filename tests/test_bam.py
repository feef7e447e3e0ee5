"""Host BGZF/BAM decoder (csrc/bam.cu): the SoA columns decoded from a BAM file
must equal, bit for bit, the columns the packer derives from the same BAM-level arrays -- i.e. the
file path and the array path feed the kernels identical input.  Pure host code: runs without a GPU."""
import os

import numpy as np
import pytest

import cnvetti_b200 as cb
from cnvetti_b200 import bam, synth


def _expected(d, min_unclipped):
    end, mid, frag_end, flags = cb.pack_reads(d["pos"], d["isize"], d["flag"], d["cigar_off"], d["cigar"], min_unclipped)
    return dict(pos=np.asarray(d["pos"], np.int32), end=end, mid=mid, frag_end=frag_end,
                mapq=np.asarray(d["mapq"], np.uint8), flags=flags)


def _check(batch, exp):
    for k, v in exp.items():
        got = np.asarray(getattr(batch, k))
        assert got.dtype == v.dtype and np.array_equal(got, v), k


@pytest.mark.parametrize("threads,min_unclipped", [(1, 0.6), (4, 0.6), (3, 0.95)])
def test_bam_decode_matches_packer(tmp_path, threads, min_unclipped):
    contigs = [("chr1", 300_000), ("chr2", 200_000), ("chrUn_x", 50_000), ("chr3", 100_000)]
    recs = {0: synth.synth_bam_records(1, 300_000, 30_000), 1: synth.synth_bam_records(2, 200_000, 12_345),
            3: synth.synth_bam_records(3, 100_000, 1),      # chrUn_x stays empty
            -1: synth.synth_bam_records(4, 1000, 17)}       # reads without a reference
    path = synth.write_bam(str(tmp_path / "a.bam"), contigs, recs, sample="NA12878")
    with bam.BamReader(path, threads=threads) as r:
        assert r.references == [c[0] for c in contigs] and r.lengths == [c[1] for c in contigs]
        assert r.sample_names() == ["NA12878"]
        with pytest.raises(bam.BamError):
            r.fetch(0)
        r.decode(min_unclipped=min_unclipped, pinned=False)
        st = r.stats()
        n_all = sum(len(v["pos"]) for v in recs.values())
        assert st["records"] == n_all and st["unplaced"] == len(recs[-1]["pos"])
        assert st["uncompressed_bytes"] > st["compressed_bytes"] > 0 and st["decode_seconds"] > 0
        for tid in (0, 1, 3):
            _check(r.fetch(tid), _expected(recs[tid], min_unclipped))
        assert len(r.fetch("chrUn_x")) == 0
        regs = r.regions()  # default --contig-regex ^(chr)?\d\d?$ (cli.yaml:111): chrUn_x is left out
        assert [g.name for g in regs] == ["chr1", "chr2", "chr3"] and regs[1].length == 200_000
        # sub-region: what fetch(tid, start, end) of an indexed reader yields (overlap semantics)
        full = _expected(recs[0], min_unclipped)
        for a, z in ((0, 1000), (120_000, 121_000), (299_000, 10**9), (5, 6)):
            sel = np.flatnonzero((full["pos"] < z) & (full["end"] > a))
            sub = r.fetch("chr1", a, z)
            assert len(sub) == sel.size
            _check(sub, {k: v[sel] for k, v in full.items()})
        # decoding again (other options) replaces the columns
        r.decode(min_unclipped=0.0, pinned=False)
        _check(r.fetch(1), _expected(recs[1], 0.0))


def test_bam_small_blocks_and_odd_records(tmp_path):
    """Records and the header straddle BGZF blocks (tiny blocks); unpaired, unmapped-placed and
    clipped reads; CIGARs with every operator."""
    pos = np.array([10, 10, 25, 40, 40, 77], np.int32)
    cig, off = [], [0]
    for ops in ([(50, 0)], [(5, 4), (20, 0), (3, 1), (22, 0)], [(10, 5), (30, 0), (1000, 3), (10, 7), (4, 8), (2, 2)],
                [], [(60, 4), (40, 0)], [(3, 6), (50, 0)]):
        cig += [(n << 4) | o for n, o in ops]
        off.append(len(cig))
    d = dict(pos=pos, isize=np.array([300, -300, 0, 0, 120, 5], np.int32),
             flag=np.array([99, 147, 0, 4 | 1 | 8, 1 | 2 | 0x400, 16], np.uint16),
             mapq=np.array([60, 0, 255, 0, 13, 37], np.uint8),
             cigar_off=np.array(off, np.uint64), cigar=np.array(cig, np.uint32))
    contigs = [("%d" % i, 1000 + i) for i in range(1, 41)]  # a header longer than one tiny block
    path = synth.write_bam(str(tmp_path / "b.bam"), contigs, {4: d}, block_payload=97)
    with bam.BamReader(path, threads=2) as r:
        assert len(r.references) == 40 and r.lengths[39] == 1040
        r.decode(min_unclipped=0.6, pinned=False)
        _check(r.fetch(4), _expected(d, 0.6))
        assert all(len(r.fetch(t)) == 0 for t in range(40) if t != 4)


def test_bam_chunk_seams(tmp_path, monkeypatch):
    """The decoder reads the file in chunks of compressed bytes: BGZF blocks cut by a chunk end are
    completed by the next read, records cut by the end of a chunk's last block are carried over, and
    the header may span chunks.  Tiny chunks put a seam every few blocks."""
    contigs = [("chr%d" % i, 150_000) for i in range(1, 30)]  # a 700-byte header
    recs = {0: synth.synth_bam_records(11, 150_000, 4000), 7: synth.synth_bam_records(12, 150_000, 2500),
            28: synth.synth_bam_records(13, 150_000, 1)}
    path = synth.write_bam(str(tmp_path / "seams.bam"), contigs, recs, block_payload=1500)
    expected = {t: _expected(d, 0.6) for t, d in recs.items()}
    for chunk in ("64", "333", "5000", "70001"):
        monkeypatch.setenv("CNVB_BAM_CHUNK", chunk)
        with bam.BamReader(path, threads=3) as r:
            assert len(r.references) == 29
            r.decode(min_unclipped=0.6, pinned=False)
            assert r.stats()["records"] == sum(len(d["pos"]) for d in recs.values())
            for t, exp in expected.items():
                _check(r.fetch(t), exp)
    monkeypatch.delenv("CNVB_BAM_CHUNK")


@pytest.mark.parametrize("env", [{}, {"CNVB_BAM_ZLIB": "1"}, {"CNVB_BAM_SERIAL_WALK": "1"}, {"CNVB_BAM_WALK_SHARE": "64"},
                                 {"CNVB_BAM_WALK_SHARE": "200"}, {"CNVB_BAM_WALK_SHARE": "5000", "CNVB_BAM_CHUNK": "70001"}])
@pytest.mark.parametrize("level,block_payload", [(6, None), (1, 3000), (9, 65280), (0, 20000)])
def test_bam_decoder_paths_agree(tmp_path, monkeypatch, env, level, block_payload):
    """The decoder's own DEFLATE / CRC-32 (hostz.cpp) against zlib behind the same file, and the record walk on all
    threads (guessed record starts, kept only where the walk before them arrives) against the serial walk: the same
    columns whatever the compression level, the block size and the size of the threads' shares (shares smaller
    than a record, records spanning shares)."""
    for key, val in env.items():
        monkeypatch.setenv(key, val)
    contigs = [("chr1", 200_000), ("chr2", 150_000), ("chr3", 90_000)]
    recs = {0: synth.synth_bam_records(21, 200_000, 6000), 1: synth.synth_bam_records(22, 150_000, 3500),
            2: synth.synth_bam_records(23, 90_000, 11), -1: synth.synth_bam_records(24, 1000, 5)}
    kw = {} if block_payload is None else {"block_payload": block_payload}
    path = synth.write_bam(str(tmp_path / "p.bam"), contigs, recs, level=level, **kw)
    with bam.BamReader(path, threads=4) as r:
        r.decode(min_unclipped=0.6, pinned=False)
        assert r.stats()["records"] == sum(len(d["pos"]) for d in recs.values()) and r.stats()["unplaced"] == 10
        for t in (0, 1, 2):
            _check(r.fetch(t), _expected(recs[t], 0.6))


def test_bam_realistic_payload_decodes_to_the_same_columns(tmp_path):
    """synth.write_bam(payload="random") -- random bases and binned qualities, the file the bench times the decoder
    on beside the filler one -- holds the same records: same columns, a compression ratio a sequencer's file has."""
    d = synth.synth_bam_records(31, 400_000, 20_000)
    a = synth.write_bam(str(tmp_path / "filler.bam"), [("chr1", 400_000)], [d], level=6)
    b = synth.write_bam(str(tmp_path / "random.bam"), [("chr1", 400_000)], [d], level=6, payload="random")
    ratio = []
    for path in (a, b):
        with bam.BamReader(path, threads=2) as r:
            r.decode(min_unclipped=0.6, pinned=False)
            st = r.stats()
            ratio.append(st["uncompressed_bytes"] / st["compressed_bytes"])
            _check(r.fetch(0), _expected(d, 0.6))   # (the columns belong to the reader: compared while it is open)
    assert ratio[0] > 8 and 1.5 < ratio[1] < 4


def test_bam_walk_rejects_a_wrong_guess(tmp_path, monkeypatch):
    """The threaded record walk guesses record starts; here the guess MUST be wrong: one record's quality string is
    made of byte patterns that look like a chain of small records, and it is long enough for every share boundary
    to fall inside it.  The share in which that record ends begins its walk at a fake record, the join sees that the
    walk before it ends elsewhere and walks the share again: the columns are those of the serial walk."""
    import struct
    from cnvetti_b200.synth import _bgzf_block

    def record(pos, l_seq, qual=None, mapq=30, flag=0, tlen=0):
        name = b"q\0"
        cigar = struct.pack("<I", (l_seq << 4) | 0)
        body = struct.pack("<iiBBHHHiiii", 0, pos, len(name), mapq, 4680, 1, flag, l_seq, -1, -1, tlen) + name + cigar + \
            bytes((l_seq + 1) // 2) + (qual if qual is not None else b"\xff" * l_seq)
        return struct.pack("<I", len(body)) + body

    fake = struct.pack("<IiiBBHHHiiii", 40, 0, 7, 4, 0, 0, 0, 0, 0, -1, -1, 0) + b"abc\0" + bytes(4)   # 44 bytes, chains on
    assert len(fake) == 44
    L = 44 * 1500
    recs = [record(10 + i, 50) for i in range(20)] + [record(40, L, qual=fake * 1500)] + [record(100 + 3 * i, 60) for i in range(400)]
    text = b"@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:chr1\tLN:100000\n"
    stream = b"BAM\x01" + struct.pack("<I", len(text)) + text + struct.pack("<I", 1) + struct.pack("<I", 5) + b"chr1\0" + \
        struct.pack("<I", 100000) + b"".join(recs)
    path = str(tmp_path / "fake.bam")
    with open(path, "wb") as f:
        for a in range(0, len(stream), 0xff00):
            f.write(_bgzf_block(stream[a:a + 0xff00], 6))
        f.write(_bgzf_block(b"", 6))
    out = {}
    for mode, env in (("serial", {"CNVB_BAM_SERIAL_WALK": "1"}), ("threads", {"CNVB_BAM_WALK_SHARE": "64"})):
        for key in ("CNVB_BAM_SERIAL_WALK", "CNVB_BAM_WALK_SHARE"):
            monkeypatch.delenv(key, raising=False)
        for key, val in env.items():
            monkeypatch.setenv(key, val)
        with bam.BamReader(path, threads=4) as r:
            r.decode(min_unclipped=0.6, pinned=False)
            assert r.stats()["records"] == len(recs)
            got = r.fetch(0)
            out[mode] = {k: np.asarray(getattr(got, k)).copy() for k in ("pos", "end", "mid", "frag_end", "mapq", "flags")}
    assert out["serial"]["pos"].tolist() == [10 + i for i in range(20)] + [40] + [100 + 3 * i for i in range(400)]
    assert out["serial"]["end"][20] == 40 + L
    for k, v in out["serial"].items():
        assert np.array_equal(out["threads"][k], v), k


def test_bam_errors_are_loud(tmp_path):
    with pytest.raises(bam.BamError, match="Could not open"):
        bam.BamReader(str(tmp_path / "missing.bam"))
    p = tmp_path / "text.bam"
    p.write_bytes(b"not a bam file at all, just text that is long enough")
    with pytest.raises(bam.BamError, match="BGZF"):
        bam.BamReader(str(p))
    good = synth.write_bam(str(tmp_path / "c.bam"), [("chr1", 100_000)], [synth.synth_bam_records(5, 100_000, 5000)])
    raw = open(good, "rb").read()
    cut = tmp_path / "cut.bam"
    cut.write_bytes(raw[:len(raw) // 2])
    with bam.BamReader(str(cut)) as r:
        with pytest.raises(bam.BamError, match="truncated"):
            r.decode(pinned=False)
    flip = bytearray(raw)
    flip[len(raw) // 2] ^= 0x55
    bad = tmp_path / "flip.bam"
    bad.write_bytes(bytes(flip))
    with pytest.raises(bam.BamError, match="corrupt"):  # (a small file: found while reading the header)
        with bam.BamReader(str(bad)) as r:
            r.decode(pinned=False)


@pytest.mark.gpu
def test_coverage_from_bam_equals_arrays(tmp_path):
    """`cmd coverage` fed from the file equals the run fed from the arrays (region loop, three kinds)."""
    from cnvetti_b200 import pipeline
    from util import soa_batch, to_np
    ctx = cb.Context(0)
    lens = (400_000, 250_000)
    recs = [synth.synth_bam_records(21 + i, L, 20_000) for i, L in enumerate(lens)]
    path = synth.write_bam(str(tmp_path / "d.bam"), [("chr1", lens[0]), ("chr2", lens[1])], recs)
    for kind in ("Fragments", "Coverage"):
        opts = cb.CoverageOptions(window_length=500, min_mapq=20, count_kind=kind)
        with bam.BamReader(path) as r:
            r.decode(min_unclipped=opts.min_unclipped, pinned=True)
            t1 = pipeline.coverage_run(r.regions(), opts, ctx=ctx)
            ref = [pipeline.Region("chr%d" % (i + 1), L, soa_batch(d)) for i, (L, d) in enumerate(zip(lens, recs))]
            t2 = pipeline.coverage_run(ref, opts, ctx=ctx)
            for col in ("start", "end", "rcv", "lcv", "ft"):
                assert np.array_equal(to_np(getattr(t1, col)), to_np(getattr(t2, col))), (kind, col)
            assert t1.num_processed == t2.num_processed


@pytest.mark.gpu
def test_cmd_coverage_and_normalize_files(tmp_path, oracle):
    """The two commands of the path at file level: BAM (+ FASTA) in, coverage VCF out, normalised VCF out;
    the values in the files are the oracle's (text precision of %g for the floats)."""
    from cnvetti_b200 import files, pipeline
    from util import oracle_reads, to_np as to_np3
    ctx = cb.Context(0)
    lens = (300_000, 200_000)
    names = ("chr1", "chr2")
    recs = [synth.synth_bam_records(31 + i, L, 15_000) for i, L in enumerate(lens)]
    seqs = [synth.synth_reference(41 + i, L) for i, L in enumerate(lens)]
    bam_path = synth.write_bam(str(tmp_path / "in.bam"), list(zip(names, lens)) + [("chrUn", 1000)], recs, sample="S1")
    fa = tmp_path / "ref.fa"
    with open(fa, "wb") as f:
        for n_, s in zip(names, seqs):
            f.write(b">" + n_.encode() + b" synthetic\n")
            for a in range(0, s.size, 60):
                f.write(s[a:a + 60].tobytes() + b"\n")
    opts = cb.CoverageOptions(window_length=500, min_mapq=20)
    out1 = str(tmp_path / "cov.vcf.gz")
    table, timings = pipeline.cmd_coverage(bam_path, out1, opts, reference=str(fa), ctx=ctx)
    assert set(timings) == {"bam_decode", "fasta", "device", "write"} and table.names == ["chr1", "chr2"]
    r = files.VariantFile(out1)
    # only the contigs matching --contig-regex are listed (build_chroms_bam, lib.rs:734-741)
    assert r.samples == ["S1"] and r.contigs == [("chr1", lens[0]), ("chr2", lens[1])]
    assert os.path.exists(out1 + ".tbi")
    oopts = oracle.cov_opts(window_length=500, min_mapq=20)
    lcv_all, gc_all = [], []
    for d, s, L in zip(recs, seqs, lens):
        counts, _, _ = oracle.frag_gw(oracle_reads(oracle, d), oopts, L)
        cov, start, end, frm = oracle.frag_gw_stats(counts, oopts, L)
        lcv_all.append(oracle.cov_records(cov, None, start, end, frm, False, oopts)[0])
        gc_all.append(oracle.refstats(s, 500)[0])
    lcv, gc = np.concatenate(lcv_all), np.concatenate(gc_all)
    file_lcv = r.format_float("LCV")[0][:, 0]
    np.testing.assert_allclose(file_lcv, lcv, rtol=1e-5)      # text container: six significant digits
    file_gc, _ = r.info_float("GC")
    np.testing.assert_allclose(file_gc[np.isfinite(gc)], gc[np.isfinite(gc)], rtol=1e-5)
    assert not np.isfinite(file_gc[~np.isfinite(gc)]).any()
    out2 = str(tmp_path / "norm.vcf")
    pipeline.cmd_normalize(out1, out2, "MedianGcBinned", ctx=ctx)
    r2 = files.VariantFile(out2)
    # the oracle on the values as they stand in the first file (what lib-normalize reads)
    cv, _, _, flags, _ = oracle.norm_gc_median(file_lcv, file_gc)
    has = (flags & 1).astype(bool)
    assert has.sum() > 100
    cv2, pcv = r2.format_float("CV")
    assert np.array_equal(pcv != 0, has)
    np.testing.assert_allclose(cv2[has, 0], cv[has], rtol=1e-5)
    assert np.array_equal(r2.fixed()["pos"], r.fixed()["pos"]) and r2.fixed()["canonical_ids"]
    # binary container: the same chain is bit-exact (BCF stores the f32 values as they are)
    out1b, out2b = str(tmp_path / "cov.bcf"), str(tmp_path / "norm.bcf")
    tb, _ = pipeline.cmd_coverage(bam_path, out1b, opts, reference=str(fa), ctx=ctx)
    assert os.path.exists(out1b + ".csi")
    rb = files.VariantFile(out1b)
    assert rb.is_bcf and np.array_equal(rb.format_float("LCV")[0][:, 0].view(np.uint32), lcv.view(np.uint32))
    assert np.array_equal(rb.info_float("GC")[0].view(np.uint32), gc.view(np.uint32))
    pipeline.cmd_normalize(out1b, out2b, "MedianGcBinned", ctx=ctx)
    ocv, ocv2, _, oflags, _ = oracle.norm_gc_median(lcv, gc)
    rb2 = files.VariantFile(out2b)
    gcv, gp = rb2.format_float("CV")
    assert np.array_equal(gp != 0, (oflags & 1) != 0)
    assert np.array_equal(gcv[gp != 0, 0].view(np.uint32), ocv[gp != 0].view(np.uint32))
    assert np.array_equal(rb2.info_flag("FEW_GCWINDOWS") != 0, (oflags & 8) != 0)
    # --mask-piles: the file-level driver == pile collection + region loop on the decoded regions
    from cnvetti_b200 import piles as piles_mod
    # (a sparse BAM: isolated piles of depth 1 and 2 are what the Poisson fit of compute_threshold needs)
    sparse = [synth.synth_bam_records(51 + i, L, 6_000) for i, L in enumerate(lens)]
    bam_path = synth.write_bam(str(tmp_path / "sparse.bam"), list(zip(names, lens)), sparse, sample="S2")
    opts = cb.CoverageOptions(window_length=500, min_mapq=0)
    out3 = str(tmp_path / "masked.vcf")
    t3, _ = pipeline.cmd_coverage(bam_path, out3, opts, ctx=ctx, mask_piles=True, mask_piles_fdr=0.01)
    with bam.BamReader(bam_path) as rd:
        rd.decode(min_unclipped=opts.min_unclipped, pinned=False)
        regs = rd.regions()
        masks, thresh = piles_mod.gather_piles(regs, min_mapq=opts.min_mapq, mask_piles_fdr=0.01, ctx=ctx)
        for g_, m_ in zip(regs, masks):
            g_.piles = m_
        t4 = pipeline.coverage_run(regs, opts, ctx=ctx)
    assert sum(len(m_[0]) for m_ in masks) > 0
    assert np.array_equal(to_np3(t3.rcv), to_np3(t4.rcv)) and np.array_equal(to_np3(t3.frm), to_np3(t4.frm))
    assert t3.num_skipped == t4.num_skipped and sum(t3.num_skipped) > 0


def test_bam_decode_property_random_records(tmp_path):
    """Randomised records (hypothesis): arbitrary flags / mapq / template lengths, CIGARs of every operator
    mix and length, positions in file order over several contigs -- file path == array path."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    cigar_op = st.tuples(st.integers(0, 1 << 20), st.integers(0, 8))
    record = st.tuples(st.integers(0, 2_000_000), st.integers(-5000, 5000), st.integers(0, 4095), st.integers(0, 255),
                       st.lists(cigar_op, min_size=0, max_size=6))
    counter = [0]

    @settings(max_examples=20, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(st.lists(st.lists(record, min_size=0, max_size=40), min_size=1, max_size=4), st.floats(0.0, 1.0, width=32))
    def run(per_contig, min_unclipped):
        recs = {}
        for tid, rows in enumerate(per_contig):
            if not rows:
                continue
            rows = sorted(rows, key=lambda r: r[0])
            cig, off = [], [0]
            for r in rows:
                cig += [(n << 4) | o for n, o in r[4]]
                off.append(len(cig))
            recs[tid] = dict(pos=np.array([r[0] for r in rows], np.int32), isize=np.array([r[1] for r in rows], np.int32),
                             flag=np.array([r[2] for r in rows], np.uint16), mapq=np.array([r[3] for r in rows], np.uint8),
                             cigar_off=np.array(off, np.uint64), cigar=np.array(cig, np.uint32))
        counter[0] += 1
        path = synth.write_bam(str(tmp_path / ("p%d.bam" % counter[0])), [("c%d" % i, 3_000_000) for i in range(len(per_contig))],
                               recs, read_len=10, block_payload=700)
        with bam.BamReader(path, threads=2) as r:
            r.decode(min_unclipped=min_unclipped, pinned=False)
            assert r.stats()["records"] == sum(len(v["pos"]) for v in recs.values())
            for tid in range(len(per_contig)):
                if tid in recs:
                    _check(r.fetch(tid), _expected(recs[tid], min_unclipped))
                else:
                    assert len(r.fetch(tid)) == 0

    run()


@pytest.mark.parametrize("block_payload,threads", [(0xff00, 4), (1500, 2)])
def test_bam_indexed_fetch_equals_full_decode(tmp_path, block_payload, threads, monkeypatch):
    """cnvb_bam_fetch (bam::IndexedReader::fetch through the .bai, lib-coverage/src/lib.rs:533-543): for random
    intervals the records are exactly those the whole-file decode selects by overlap, in file order; only a part of
    the file is inflated.  Small BGZF blocks and small read chunks put records across every kind of seam."""
    if block_payload < 0xff00:
        monkeypatch.setenv("CNVB_BAM_CHUNK", "4096")
    rng = np.random.default_rng(7)
    contigs = [("chr1", 400_000), ("chr2", 250_000), ("chrE", 90_000), ("chr3", 120_000)]
    recs = {0: synth.synth_bam_records(11, 400_000, 20_000), 1: synth.synth_bam_records(12, 250_000, 9_000),
            3: synth.synth_bam_records(13, 120_000, 3_000), -1: synth.synth_bam_records(14, 1000, 50)}
    path = synth.write_bam(str(tmp_path / "x.bam"), contigs, recs, block_payload=block_payload, index=True, read_len=20)
    with bam.BamReader(path, threads=threads) as full, bam.BamReader(path, threads=threads) as ix:
        assert ix.has_index()
        full.decode(min_unclipped=0.6, pinned=False)
        whole = full.stats()["compressed_bytes"]
        for tid, L in ((0, 400_000), (1, 250_000), (3, 120_000), (2, 90_000)):
            cases = [(0, L), (0, 1), (L - 1, L), (16384, 16385), (16383, 16384), (5, 5)]
            cases += [tuple(sorted(rng.integers(0, L, 2).tolist())) for _ in range(12)]
            for a, b in cases:
                want = full.fetch(tid, a, b)
                got = ix.fetch_indexed(tid, a, b, min_unclipped=0.6)
                assert len(got) == len(want), (tid, a, b, len(got), len(want))
                for k in ("pos", "end", "mid", "frag_end", "mapq", "flags"):
                    assert np.array_equal(np.asarray(getattr(got, k)), np.asarray(getattr(want, k))), (tid, a, b, k)
                if b - a < L // 20 and len(want) and block_payload < 0xff00:
                    assert ix.stats()["compressed_bytes"] < whole // 3, (tid, a, b)   # a seek, not a scan
    # no index: the reference cannot open the file either (lib.rs:531-532)
    plain = synth.write_bam(str(tmp_path / "noidx.bam"), contigs, recs)
    with bam.BamReader(plain) as r:
        assert not r.has_index()
        with pytest.raises(bam.BamError, match="BAI-indexed"):
            r.fetch_indexed(0, 0, 100)
