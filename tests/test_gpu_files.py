"""The commands of the path chained through FILES only (commands.py):
BAM -> cmd coverage -> cmd normalize -> cmd merge-cov -> cmd build-model-wis -> cmd coverage (--wis-model-bcf)
-> cmd mod-coverage, with BCF containers in between -- every value in every file against the oracle chain on the
same reads, bit for bit (BCF stores f32 as it is)."""
import os

import numpy as np
import pytest

import cnvetti_b200 as cb
from cnvetti_b200 import _native as N
from cnvetti_b200 import commands, files, model_wis as mw, synth
from util import assert_f32_equal_bits, oracle_reads

pytestmark = pytest.mark.gpu

LENS = (400_000, 250_000, 150_000)
NAMES = ("chr1", "chr2", "chr3")
MISS = np.uint32(N.F32_MISSING_BITS)


@pytest.fixture(scope="module")
def cohort(tmp_path_factory, oracle):
    """12 BAM files on three contigs + a targets BED; the oracle's normalised column per sample."""
    tmp = tmp_path_factory.mktemp("cohort")
    tgts = [synth.synth_targets(80 + i, L, 90) for i, L in enumerate(LENS)]
    bed = str(tmp / "targets.bed")
    with open(bed, "w") as f:
        for name, (s, e) in zip(NAMES, tgts):
            for a, b in zip(s, e):
                f.write("%s\t%d\t%d\n" % (name, a, b))
    bams, ocols, oraw = [], [], []
    oopts = oracle.cov_opts(window_length=0, min_mapq=0)
    for s in range(12):
        recs, lcvs, raws = [], [], []
        for i, L in enumerate(LENS):
            d = synth.synth_bam_records(1000 + 10 * s + i, L, 5000)
            recs.append(d)
            counts, _, _ = oracle.frag_targets(oracle_reads(oracle, d), oopts, tgts[i][0], tgts[i][1])
            lcv, _, ft = oracle.cov_records(counts.astype(np.float32), None, tgts[i][0], tgts[i][1], None, True, oopts)
            lcvs.append(lcv)
            raws.append(counts.astype(np.float32))
        bams.append(synth.write_bam(str(tmp / ("s%d.bam" % s)), list(zip(NAMES, LENS)) + [("chrUn", 500)], recs, sample="S%d" % s))
        ocols.append(oracle.norm_total(np.concatenate(lcvs))[1])
        oraw.append(np.concatenate(raws))
    return dict(tmp=tmp, bed=bed, bams=bams, tgts=tgts, oX=np.stack(ocols, axis=1), oraw=oraw,
                tids=np.concatenate([np.full(len(t[0]), i, np.uint32) for i, t in enumerate(tgts)]))


def test_files_only_chain_against_the_oracle(cohort, oracle):
    ctx = cb.Context(0)
    tmp, oX, otids = cohort["tmp"], cohort["oX"], cohort["tids"]
    T, S = oX.shape
    model_path, merged_path = str(tmp / "model.bcf"), str(tmp / "merged.bcf")
    # ---- quick wis-build-model: coverage + normalize per BAM, merge-cov, build-model-wis -- through files
    opts = mw.BuildModelWisOptions.quick(max_ref_targets=20, min_ref_targets=3)
    norm = []
    for i, bam in enumerate(cohort["bams"]):
        cov = str(tmp / ("coverage.%d.bcf" % i))
        table, _ = commands.cmd_coverage(bam, cov, commands._quick_coverage_options(), targets_bed=cohort["bed"], ctx=ctx)
        assert table.names == list(NAMES) and os.path.exists(cov + ".csi")
        with files.VariantFile(cov) as vf:
            assert vf.samples == ["S%d" % i] and vf.contigs == list(zip(NAMES, LENS)) and (vf.fixed()["alt"] == 1).all()
            assert_f32_equal_bits(vf.format_float("RCV")[0][:, 0], cohort["oraw"][i], "RCV of sample %d" % i)
            ft, _ = vf.format_ft()
            assert np.array_equal(ft[:, 0] != 0, cohort["oraw"][i] < 10)        # LOWCOV (lib.rs:627-631)
        nrm = str(tmp / ("normalized.%d.bcf" % i))
        commands.cmd_normalize(cov, nrm, "TotalCovSum", ctx=ctx)
        with files.VariantFile(nrm) as vf:
            cv, p = vf.format_float("CV")
            assert p.all() and "##cnvetti_cmdNormalizeVersion=0.1.0" in vf.header
            assert_f32_equal_bits(cv[:, 0], oX[:, i], "CV of sample %d" % i)
        norm.append(nrm)
    commands.cmd_merge_cov(norm, merged_path, ctx=ctx)
    with files.VariantFile(merged_path) as vf:
        assert vf.samples == ["S%d" % i for i in range(S)] and vf.n == T and os.path.exists(merged_path + ".csi")
        assert_f32_equal_bits(vf.format_float("CV")[0], oX, "merged panel")
        assert np.array_equal(vf.fixed()["rid"].astype(np.uint32), otids)
        assert "##cnvetti_cmdMergeCovVersion=0.1.0" in vf.header
    commands.cmd_build_model_wis(merged_path, model_path, opts, ctx=ctx)
    odist, oidx = oracle.wis_distances(oX, otids, max_ref_targets=20)
    thr, pidx, plen = oracle.wis_prune(odist, oidx, opts.filter_z_score)
    calls = oracle.wis_calls(oX, pidx, plen, 3, opts.filter_z_score, opts.filter_rel)
    filt = oracle.wis_filters(plen, calls, 3, opts.max_samples_reliable)
    with files.VariantFile(model_path) as mf:
        assert mf.samples == [] and mf.n == T and mf.fixed()["canonical_ids"]
        assert np.array_equal(mf.info_int("NUM_CALLS")[0].astype(np.uint32), calls)
        want = (((filt & 1) != 0) * N.VF_FEW_REF + ((filt & 2) != 0) * N.VF_UNRELIABLE).astype(np.uint8)
        assert np.array_equal(mf.fixed()["filters"], want) and (want != 0).any() and (want == 0).any()
        off, idx = mf.ref_targets()
        assert np.array_equal(np.diff(off.astype(np.int64)), plen)
        assert np.array_equal(idx, np.concatenate([pidx[t, :plen[t]] for t in range(T)]))     # distance order
    # ---- quick wis-call for two samples: coverage on the model's targets, normalize, mod-coverage
    for s in (0, 7):
        out = str(tmp / ("call.%d.bcf" % s))
        commands.quick_wis_call(cohort["bams"][s], model_path, out, ctx=ctx)
        o = oracle.modcov(np.ascontiguousarray(oX[:, s]), filt, pidx, plen)
        has = (o["status"] & 1) != 0
        assert has.any() and (~has).any()
        with files.VariantFile(out) as vf:
            assert vf.n == T and "##cnvetti_cmdModCoverageVersion=0.1.0" in vf.header and os.path.exists(out + ".csi")
            assert np.array_equal((vf.fixed()["filters"] & N.VF_NO_REF) != 0, ~has)          # lib-mod-cov :241-247
            cv, p = vf.format_float("CV")
            assert p.all()
            assert_f32_equal_bits(cv[has, 0], o["cv_rel"][has], "relative CV")
            assert_f32_equal_bits(cv[~has, 0], oX[~has, s], "CV of records without probe info stays")
            cvz, p = vf.format_float("CVZ")
            assert np.array_equal(p != 0, has)
            assert_f32_equal_bits(cvz[has, 0], o["cvz"][has], "CVZ")
            cv2, p = vf.format_float("CV2")
            assert np.array_equal(p != 0, (o["status"] & 2) != 0)
            fin = (o["status"] & 2) != 0
            miss = cv2.view(np.uint32)[:, 0] == MISS
            np.testing.assert_allclose(cv2[fin & ~miss, 0], o["cv2"][fin & ~miss], rtol=1e-6)   # log2: tolerance of the oracle tests
            rm, p = vf.info_float("REF_MEAN")
            assert np.array_equal(p != 0, has)
            assert_f32_equal_bits(rm[has], o["ref_mean"][has], "REF_MEAN")
            assert_f32_equal_bits(vf.info_float("REF_STD_DEV")[0][has], o["ref_sd"][has], "REF_STD_DEV")


def test_merge_cov_pairs_records_by_position(tmp_path):
    """The synced reader of lib-merge-cov (src/lib.rs:241-250, EXACT pairing): records missing from one input are
    filled with the missing value for its samples; INFO comes from the first input that has the record."""
    from cnvetti_b200 import pipeline
    ctx = cb.Context(0)
    contigs = [("chr1", 10_000), ("chr2", 5_000)]

    def table(starts1, starts2, scale):
        t = pipeline.CoverageTable()
        t.names, t.offsets = ["chr1", "chr2"], [0, len(starts1), len(starts1) + len(starts2)]
        t.start = np.array(list(starts1) + list(starts2), np.int32)
        t.end = (t.start + 100).astype(np.int32)
        t.rcv = (np.arange(t.start.size) + 1).astype(np.float32) * scale
        t.lcv = (t.rcv / 100).astype(np.float32)
        t.ft = np.zeros(t.start.size, np.uint8)
        t.cv = (t.lcv * 3).astype(np.float32)
        return t
    a = table([0, 200, 400], [100, 300], 1.0)
    b = table([200, 400, 600], [300], 10.0)
    pa = files.write_coverage_file(str(tmp_path / "a.bcf"), a, ["A"], contigs, is_targets=True)
    pb = files.write_coverage_file(str(tmp_path / "b.bcf"), b, ["B"], contigs, is_targets=True)
    out = str(tmp_path / "m.vcf.gz")
    commands.cmd_merge_cov([pa, pb], out, ctx=ctx)
    with files.VariantFile(out) as vf:
        fx = vf.fixed()
        assert vf.samples == ["A", "B"] and os.path.exists(out + ".tbi")
        assert fx["rid"].tolist() == [0, 0, 0, 0, 1, 1] and fx["pos"].tolist() == [0, 200, 400, 600, 100, 300]
        cv, p = vf.format_float("CV")
        assert p.all()
        m = cv.view(np.uint32) == MISS
        assert m.tolist() == [[False, True], [False, False], [False, False], [True, False], [False, True], [False, False]]
        np.testing.assert_allclose(cv[1], [a.cv[1], b.cv[0]], rtol=1e-5)
        np.testing.assert_allclose(cv[5], [a.cv[4], b.cv[3]], rtol=1e-5)
    os.remove(pb + ".csi")
    with pytest.raises(N.CnvbError, match="index missing"):
        commands.cmd_merge_cov([pa, pb], out, ctx=ctx)


def test_cmd_coverage_genome_region_through_the_index(tmp_path, oracle):
    """`--genome-region chr:a-b` (lib-coverage/src/lib.rs:707-711): one region (chr, a - 1, b), the reads fetched
    through the .bai, windows counted from 0 to b -- against the oracle on the records htslib's fetch yields."""
    ctx = cb.Context(0)
    L = 400_000
    d = synth.synth_bam_records(77, L, 30_000)
    path = synth.write_bam(str(tmp_path / "r.bam"), [("chr1", L), ("chr2", 1000)], [d, None], sample="R", index=True)
    a, b = 100_000, 200_300
    end, _, _, _ = cb.pack_reads(d["pos"], d["isize"], d["flag"], d["cigar_off"], d["cigar"], 0.6)
    keep = np.flatnonzero((d["pos"] < b) & (end > a))
    coff = np.asarray(d["cigar_off"], np.int64)
    ncig = (coff[1:] - coff[:-1])[keep]
    sub = dict(pos=d["pos"][keep], isize=d["isize"][keep], flag=d["flag"][keep], mapq=d["mapq"][keep],
               cigar_off=np.concatenate([[0], np.cumsum(ncig)]).astype(np.uint64),
               cigar=np.concatenate([d["cigar"][coff[i]:coff[i + 1]] for i in keep]))
    opts = cb.CoverageOptions(window_length=1000, count_kind="Coverage", min_mapq=10)
    out = str(tmp_path / "region.bcf")
    table, timings = commands.cmd_coverage(path, out, opts, genome_region="chr1:100,001-200_300", ctx=ctx)
    mean, sd = oracle.coverage(oracle_reads(oracle, sub), oracle.cov_opts(window_length=1000, min_mapq=10), b)[:2]
    with files.VariantFile(out) as vf:
        assert vf.n == 201 and vf.contigs == [("chr1", L), ("chr2", 1000)]
        assert_f32_equal_bits(vf.format_float("RCV")[0][:, 0], mean, "window means")
        np.testing.assert_allclose(vf.format_float("RCVSD")[0][:, 0], sd, rtol=1e-6)
        assert vf.fixed()["end"][-1] == b                      # the last window ends at the region end
    assert mean[:99].sum() == 0 and mean[101:200].min() > 0    # nothing before the interval was fetched
    with pytest.raises(N.CnvbError, match="not exactly one ':'"):
        commands.cmd_coverage(path, out, opts, genome_region="chr1", ctx=ctx)
