"""File-level mirrors of the commands of the path -- the reference's `pub fn run(&mut Logger, &XxxOptions)`
entry points (SURVEY.md 8(b) seam (ii)) with the same inputs and outputs:

    cmd_coverage          lib_coverage::run        lib-coverage/src/lib.rs:679-795
    cmd_normalize         lib_normalize::run       lib-normalize/src/lib.rs:43-66
    cmd_merge_cov         lib_merge_cov::run       lib-merge-cov/src/lib.rs:235-289
    cmd_build_model_wis   lib_model_wis::run       lib-model-wis/src/lib.rs:422-447
    cmd_mod_coverage      lib_mod_cov::run         lib-mod-cov/src/lib.rs:251-279
    quick_wis_build_model cnvetti quick wis-build-model   cnvetti/src/quick_wis_build_model.rs:133-214
    quick_wis_call        cnvetti quick wis-call          cnvetti/src/quick_wis_call.rs:124-195

Files are read and written by the library's own VCF / BCF code (files.py, csrc/bcf.cu); everything between the
files runs on the GPU through the same calls as the in-memory pipeline (pipeline.py).  The output container and
its index follow the file name (`.bcf` + `.csi`, `.vcf.gz` + `.tbi`, `.vcf`), lib-shared/src/bcf_utils.rs:47-82.
"""
import os
import re
import tempfile
import time

import numpy as np

from . import _native as N
from . import files
from .context import default_context
from .pipeline import CoverageTable, coverage_run, normalize_run, read_fasta


def _dev(ctx):
    return "cuda:%d" % ctx.device


def _to_dev(a, ctx):
    import torch
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(_dev(ctx))


# ---- cmd coverage ----------------------------------------------------------------------------------------
def load_bed_targets(path):
    """{contig: (start[], end[])} of a targets BED file, per contig in file order -- what
    load_bed_regions_from_tabix yields region by region (lib-coverage/src/lib.rs:185-212; a tabix fetch over
    the whole contig returns the lines in file order)."""
    import gzip
    opener = gzip.open if str(path).endswith(".gz") else open
    out = {}
    with opener(path, "rt") as f:
        for line in f:
            if not line.strip() or line.startswith(("#", "track", "browser")):
                continue
            arr = line.rstrip("\n").split("\t")
            if len(arr) < 3:
                raise N.CnvbError(N.ERR_INVALID, "BED file line had too few columns! %d (<3)" % len(arr))
            s, e = out.setdefault(arr[0], ([], []))
            s.append(int(arr[1]))
            e.append(int(arr[2]))
    return {k: (np.asarray(s, np.uint32), np.asarray(e, np.uint32)) for k, (s, e) in out.items()}


def load_model_targets(path, threads=0):
    """{contig: (start[], end[])} from a WIS model file: pos and INFO/END of its records, per contig in file
    order (load_regions_from_wis_model_bcf, lib-coverage/src/lib.rs:215-258)."""
    with files.VariantFile(path, threads=threads) as vf:
        fx = vf.fixed()
        out = {}
        for rid in np.unique(fx["rid"]):
            m = fx["rid"] == rid
            out[vf.contigs[int(rid)][0]] = (fx["pos"][m].astype(np.uint32), fx["end"][m].astype(np.uint32))
        return out


def parse_genome_region(text):
    """`--genome-region chr:a-b` -> (chr, a - 1, b) (GenomeRegions::from_string_list, lib-shared/src/regions.rs:56-90:
    exactly one ':' and one '-', thousands separators ',' and '_' allowed)."""
    parts = text.split(":")
    if len(parts) != 2:
        raise N.CnvbError(N.ERR_INVALID, "Problem parsing region\ncaused by: Could not parse genome region (not exactly one ':')")
    nums = parts[1].split("-")
    if len(nums) != 2:
        raise N.CnvbError(N.ERR_INVALID, "Problem parsing region\ncaused by: Could not parse genome region (not exatly one '-')")
    a, b = (int(x.replace(",", "").replace("_", "")) for x in nums)
    return parts[0], a - 1, b


def cmd_coverage(input_bam, output, options, reference=None, contig_regex=r"^(chr)?\d\d?$", targets_bed=None,
                 wis_model_bcf=None, targets=None, ctx=None, threads=0, command_line=None, mask_piles=False,
                 mask_piles_fdr=0.001, pile_max_gap=20, file_date=None, genome_region=None):
    """`cnvetti cmd coverage --input X.bam --output Y.{bcf,vcf.gz,vcf} [--reference R.fa] [--targets-bed T.bed |
    --wis-model-bcf M.bcf] [--mask-piles] ...`: BAM decode on host threads, with --mask-piles the pile collection
    and its FDR threshold (gather_piles, lib.rs:353-419), the region loop on the GPU, the file written on host
    threads.  Returns (table, timings) with host decode / device / write times apart."""
    import torch
    from . import bam as bam_mod
    ctx = ctx or default_context()
    if options.considered_regions == "TargetRegions" and options.count_kind == "Coverage":
        raise N.CnvbError(N.ERR_INVALID, "Cannot combine target regions with considering coverage. Either consider "
                          "genome-wide coverage or consider fragment counts in target regions.")       # lib.rs:696-704
    timings = {}
    t0 = time.perf_counter()
    with bam_mod.BamReader(input_bam, threads=threads) as reader:
        samples = reader.sample_names()
        region = parse_genome_region(genome_region) if genome_region else None
        region_reads = None
        if region is not None:
            # lib.rs:707-711, 533-543: ONE region (chr, a - 1, b); fetch(tid, a - 1, b) through the .bai -- only the
            # blocks of the interval are inflated.  (Without an index the whole file is decoded and filtered.)
            if region[0] not in reader.references:
                raise N.ReferencePanic(N.ERR_REFERENCE_PANIC, "contig %s is not in the BAM header (the reference unwraps "
                                       "header.tid(), lib-coverage/src/lib.rs:540)" % region[0])
            if reader.has_index():
                region_reads = reader.fetch_indexed(region[0], region[1], region[2], min_unclipped=options.min_unclipped,
                                                    pinned=True)
            else:
                reader.decode(min_unclipped=options.min_unclipped, pinned=False)
                region_reads = reader.fetch(region[0], region[1], region[2])
        else:
            reader.decode(min_unclipped=options.min_unclipped, pinned=True)
        timings["bam_decode"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        seqs = read_fasta(reference) if reference else None
        if targets is None and targets_bed:          # lib.rs:456-469
            targets = load_bed_targets(targets_bed)
        elif targets is None and wis_model_bcf:      # lib.rs:470-487
            targets = load_model_targets(wis_model_bcf, threads=threads)
        timings["fasta"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        if region is not None:
            from .pipeline import Region
            name, _, r_end = region
            seq = None if seqs is None else seqs.get(name)
            wl = options.window_length
            if seq is not None and wl:   # statistics of the contig's windows, as many as the region has (lib.rs:440-454, 560)
                seq = seq[:((r_end + wl - 1) // wl) * wl]
            regions = [Region(name, r_end, region_reads, seq, None if targets is None else targets.get(name))]
        else:
            regions = reader.regions(contig_regex=contig_regex, sequences=seqs, targets=targets)
        if options.considered_regions == "TargetRegions":
            empty = (np.zeros(0, np.uint32), np.zeros(0, np.uint32))
            for r in regions:   # a contig without targets yields an empty GenomeRegions (lib.rs:193-199)
                if r.targets is None:
                    r.targets = empty
        if seqs is not None:
            missing = [r.name for r in regions if r.sequence is None]
            if missing:
                raise N.CnvbError(N.ERR_INVALID, "Could not load reference sequence for %s" % missing[0])
        if mask_piles:  # (lib.rs:755-764: whatever the count kind; only the genome-wide fragment counter uses them)
            from . import piles as piles_mod
            masks, _ = piles_mod.gather_piles(regions, min_mapq=options.min_mapq, pile_max_gap=pile_max_gap,
                                              mask_piles_fdr=mask_piles_fdr, ctx=ctx)
            for r, m in zip(regions, masks):
                r.piles = m
        table = coverage_run(regions, options, ctx=ctx)
        torch.cuda.synchronize()
        timings["device"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        # ##contig lines: the contigs matching --contig-regex (build_chroms_bam, lib.rs:734-741)
        rx = re.compile(contig_regex)
        contigs = [(n, ln) for n, ln in zip(reader.references, reader.lengths) if rx.search(n)]
        files.write_coverage_file(output, table, samples, contigs,
                                  is_targets=options.considered_regions == "TargetRegions", command_line=command_line,
                                  file_date=file_date, threads=threads)
        timings["write"] = time.perf_counter() - t0
    return table, timings


# ---- cmd normalize ---------------------------------------------------------------------------------------
def cmd_normalize(input_path, output, normalization="MedianGcBinned", ctx=None, threads=0, command_line=None):
    """`cnvetti cmd normalize --input X --output Y --normalization TotalCovSum|MedianGcBinned`: FORMAT/LCV (LCVSD)
    and INFO/GC in; the same records plus FORMAT/CV (CV2, CVSD) and INFO/FEW_GCWINDOWS out under the input's header
    extended by the two lines of process_bcf (lib-normalize/src/shared.rs:30-47)."""
    ctx = ctx or default_context()
    with files.VariantFile(input_path, threads=threads) as vf:
        t = files.read_coverage_table(vf)
        if t.lcv is None:
            raise N.CnvbError(N.ERR_INVALID, "Problem getting value for LCV from BCF record")
        host = {k: getattr(t, k) for k in ("lcv", "lcvsd", "gc")}
        t.lcv, t.lcvsd, t.gc = _to_dev(t.lcv, ctx), _to_dev(t.lcvsd, ctx), _to_dev(t.gc, ctx)
        t.cv = t.cv2 = t.cvsd = t.norm_flags = None
        normalize_run(t, normalization, ctx=ctx)
        for k, v in host.items():
            setattr(t, k, v)
        header = vf.header + ["##cnvetti_cmdNormalizeVersion=0.1.0",
                              "##cnvetti_wiseNormalizeCommand=%s" % (command_line or "cnvetti cmd normalize")]
        files.write_coverage_file(output, t, vf.samples, vf.contigs, is_targets=t.is_targets, header_lines=header,
                                  threads=threads)
    return t


# ---- cmd merge-cov ---------------------------------------------------------------------------------------
_FLOAT_KEYS = ("CV", "CV2", "CVSD", "CVZ", "FRM", "LCV", "LCVSD", "MQ", "NCV", "RCV", "RCVSD")  # BTreeSet order (:254-259)


def cmd_merge_cov(inputs, output, ctx=None, threads=0, command_line=None):
    """`cnvetti cmd merge-cov --output Y X1 X2 ...` (merge_files, lib-merge-cov/src/lib.rs:102-232): the records of
    N coverage files paired by position (synced reader, EXACT pairing) into one multi-sample file -- INFO from the
    first file that has the record, every Float FORMAT tag side by side (missing where a file lacks the record or
    the tag), GT the no-call.  The per-tag join of the sample columns runs on the GPU (cnvb_merge_cov)."""
    import ctypes as C
    import torch
    ctx = ctx or default_context()
    if not inputs:
        raise N.CnvbError(N.ERR_INVALID, "merge-cov: no input")
    vfs = [files.VariantFile(p, threads=threads) for p in inputs]
    try:
        for p in inputs:   # reader.set_require_index(true) (:244)
            if str(p).endswith((".bcf", ".vcf.gz")) and not (os.path.exists(str(p) + ".csi") or os.path.exists(str(p) + ".tbi")):
                raise N.CnvbError(N.ERR_INVALID, "Could not open file %s for reading\ncaused by: index missing" % p)
        first = vfs[0]
        samples = [s for vf in vfs for s in vf.samples]
        # union of the records in genome order; key = (contig of the first header, pos, alt kind)
        cid = {n: i for i, (n, _) in enumerate(first.contigs)}
        keys = []
        for vf in vfs:
            fx = vf.fixed()
            try:
                remap = np.asarray([cid[n] for n, _ in vf.contigs], np.int64)
            except KeyError as e:
                raise N.CnvbError(N.ERR_INVALID, "merge-cov: contig %s is not in the first header" % e)
            keys.append((remap[fx["rid"]] << 34) | (fx["pos"].astype(np.int64) << 2) | fx["alt"].astype(np.int64))
        union = np.unique(np.concatenate(keys)) if keys else np.zeros(0, np.int64)
        T = int(union.size)
        rows = [np.searchsorted(union, k) for k in keys]
        src = np.full(T, -1, np.int64)       # first file holding the record (INFO source, :118-127)
        src_row = np.zeros(T, np.int64)
        for i in reversed(range(len(vfs))):
            src[rows[i]] = i
            src_row[rows[i]] = np.arange(rows[i].size)
        rid = (union >> 34).astype(np.int32)
        pos = ((union >> 2) & 0xFFFFFFFF).astype(np.int32)
        alt = (union & 3).astype(np.uint8)

        def gather(getter, dtype, fill):
            out = np.full(T, fill, dtype)
            for i, vf in enumerate(vfs):
                m = src == i
                if m.any():
                    out[m] = getter(vf)[src_row[m]]
            return out

        end = gather(lambda vf: vf.fixed()["end"], np.int32, 0)
        filt = gather(lambda vf: vf.fixed()["filters"], np.uint8, 0)
        info = {}
        if first.has("INFO", "GC"):
            pres = gather(lambda vf: vf.info_float("GC")[1] if vf.has("INFO", "GC") else np.zeros(vf.n, np.uint8), np.uint8, 0)
            if pres.any():
                info["gc"] = gather(lambda vf: vf.info_float("GC")[0] if vf.has("INFO", "GC") else np.full(vf.n, files.MISSING), np.float32, files.MISSING)
                info["gap"] = gather(lambda vf: vf.info_flag("GAP") if vf.has("INFO", "GAP") else np.zeros(vf.n, np.uint8), np.uint8, 0)
        if first.has("INFO", "MEAN_MAPQ"):
            info["mean_mapq"] = gather(lambda vf: vf.info_float("MEAN_MAPQ")[0], np.float32, files.MISSING)
        if first.has("INFO", "FEW_GCWINDOWS"):
            few = gather(lambda vf: vf.info_flag("FEW_GCWINDOWS") if vf.has("INFO", "FEW_GCWINDOWS") else np.zeros(vf.n, np.uint8), np.uint8, 0)
            if few.any():
                info["few_gc"] = few
        # FORMAT floats: one [T][S] block per tag; absent records / tags stay missing (:176-225)
        S = len(samples)
        fmt = []
        present_first = {}
        for key in _FLOAT_KEYS:
            cols, any_present = [], False
            for i, vf in enumerate(vfs):
                Si = len(vf.samples)
                if not vf.has("FORMAT", key):
                    cols += [None] * Si
                    continue
                v, p = vf.format_float(key)
                any_present = any_present or bool(p.any())
                v[p == 0] = files.MISSING
                for s in range(Si):
                    full = np.full(T, files.MISSING, np.float32)
                    full[rows[i]] = v[:, s]
                    cols.append(full)
                if i == 0:
                    present_first[key] = p
            if not any_present:     # dim == 0 -> the tag is skipped (:196-198)
                continue
            dev_cols = [_to_dev(c if c is not None else np.full(T, files.MISSING, np.float32), ctx) for c in cols]
            X = torch.empty((T, S), dtype=torch.float32, device=_dev(ctx))
            ptrs = (C.c_void_p * S)(*[c.data_ptr() for c in dev_cols])
            ctx.check(N.lib().cnvb_merge_cov(ctx.handle, ptrs, None, 0, S, T, X.data_ptr()))
            fmt.append((key, X.cpu().numpy(), None))
        # record order of FORMAT: tags the first file's record holds keep their place, the others follow (the
        # reference updates the translated first record in place); with files of one pipeline this is the
        # coverage order
        order = {k: i for i, k in enumerate(files.COVERAGE_FORMAT_ORDER)}
        fmt.sort(key=lambda f: order.get(f[0], 99))
        ft = None
        if any(vf.has("FORMAT", "FT") for vf in vfs):
            ft = np.zeros((T, S), np.uint8)
            s0 = 0
            for i, vf in enumerate(vfs):
                if vf.has("FORMAT", "FT"):
                    v, _ = vf.format_ft()
                    ft[rows[i], s0:s0 + len(vf.samples)] = v
                s0 += len(vf.samples)
            if not ft.any():
                ft = None
        keys_f = [f[0] for f in fmt]
        header = first.header + ["##cnvetti_cmdMergeCovVersion=0.1.0",
                                 "##cnvetti_cmdMergeCovCommand=%s" % (command_line or "cnvetti cmd merge-cov")]
        files.write_variant_file(output, header, samples, rid, pos, end, is_targets=bool(T) and int(alt[0]) == 1,
                                 filters=filt if filt.any() else None, fmt=fmt, ft=ft,
                                 ft_pos=keys_f.index("MQ") + 1 if "MQ" in keys_f else len(fmt), threads=threads, **info)
        return dict(targets=T, samples=samples)
    finally:
        for vf in vfs:
            vf.close()


# ---- cmd build-model-wis ---------------------------------------------------------------------------------
def cmd_build_model_wis(input_path, output, options=None, ctx=None, threads=0, command_line=None, file_date=None):
    """`cnvetti cmd build-model-wis --input merged.bcf --output model.bcf`: the panel FORMAT/CV[T][S] and the contig
    of every record in (read_coverage_bcf, lib-model-wis/src/lib.rs:65-125), distances / prune / per-probe calls on
    the GPU, the model records out (write_output, :340-419: FILTER FEW_REF / UNRELIABLE, INFO END, NUM_CALLS,
    REF_TARGETS in distance order)."""
    from . import model_wis
    ctx = ctx or default_context()
    options = options or model_wis.BuildModelWisOptions()
    with files.VariantFile(input_path, threads=threads) as vf:
        fx = vf.fixed()
        if not vf.has("FORMAT", "CV"):
            raise N.CnvbError(N.ERR_INVALID, "Could not read FORMAT/CV.")
        X, present = vf.format_float("CV")
        if vf.n and not present.all():
            raise N.CnvbError(N.ERR_INVALID, "Could not read FORMAT/CV.")       # :104-107
        # (a missing value inside a present tag reads as NaN in the reference and panics at :168: the library's
        # panel check reports that case)
        model = model_wis.run(_to_dev(X, ctx), _to_dev(fx["rid"].astype(np.int32), ctx), options, ctx=ctx)
        host = lambda x: x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)  # noqa: E731
        ref_len = host(model.ref_len).astype(np.int64)
        ref_idx = host(model.ref_idx)
        k = ref_idx.shape[1] if ref_idx.ndim == 2 else 0
        off = np.concatenate([[0], np.cumsum(ref_len)]).astype(np.uint64)
        flat = ref_idx[np.arange(k)[None, :] < ref_len[:, None]].astype(np.uint32) if vf.n else np.zeros(0, np.uint32)
        filt_bits = host(model.filt).astype(np.uint8)
        filters = (((filt_bits & 1) != 0) * N.VF_FEW_REF + ((filt_bits & 2) != 0) * N.VF_UNRELIABLE).astype(np.uint8)
        files.write_variant_file(output, files.model_header(vf.contigs, command_line, file_date), [], fx["rid"], fx["pos"],
                                 fx["end"], is_targets=True, filters=filters,
                                 num_calls=host(model.num_calls).astype(np.int32), ref_off=off, ref_idx=flat,
                                 threads=threads)
    return model


# ---- cmd mod-coverage ------------------------------------------------------------------------------------
def cmd_mod_coverage(input_path, input_model, output, ctx=None, threads=0, command_line=None):
    """`cnvetti cmd mod-coverage --input norm.bcf --input-wis-model model.bcf --output out.bcf`
    (lib-mod-cov/src/lib.rs:251-279): for every model record without a filter the mean / SD of the sample's CV
    over its reference targets (load_target_infos, :89-164), then the input records with REF_MEAN, REF_STD_DEV,
    CV (relative), CVZ, CV2 -- or FILTER NO_REF (write_mod_cov, :167-248).  Records are matched by ID, as there."""
    ctx = ctx or default_context()
    with files.VariantFile(input_path, threads=threads) as vf, files.VariantFile(input_model, threads=threads) as mf:
        t = files.read_coverage_table(vf)
        if t.cv is None:
            raise N.CnvbError(N.ERR_INVALID, "Could not access field FORMAT/CV")    # :50-55
        n = vf.n
        mfx = mf.fixed()
        # model records that pass (:101-112), their row in the input, their reference rows in the input
        passing = (mfx["filters"] & ~np.uint8(N.VF_PASS)) == 0
        row_of_model = mf.match_ids(vf)
        off, idx = mf.ref_targets(against=vf)            # a name the input lacks: "Could not find: ..." (:149)
        use = np.flatnonzero(passing)
        if use.size and (row_of_model[use] == 0xFFFFFFFF).any():
            bad = use[row_of_model[use] == 0xFFFFFFFF][0]
            raise N.ReferencePanic(N.ERR_REFERENCE_PANIC, "Could not get norm count for probe: record %d of the model" % (bad + 1))
        # per INPUT row: its reference rows (dense [n][k], k = longest list); rows without a passing model record
        # count as filtered (no probe info -> NO_REF)
        lens = (off[1:] - off[:-1]).astype(np.int64)
        k = int(lens[use].max()) if use.size else 0
        k = max(k, 1)
        ref_len = np.zeros(n, np.uint32)
        ref_idx = np.zeros((n, k), np.uint32)
        model_filt = np.ones(n, np.uint8)
        rows_in = row_of_model[use].astype(np.int64)
        ref_len[rows_in] = lens[use]
        model_filt[rows_in] = 0
        if use.size and int(lens[use].sum()):
            owner = np.repeat(rows_in, lens[use])                       # input row of every listed reference
            start = np.repeat(off[use].astype(np.int64), lens[use])
            flat_pos = np.concatenate([np.arange(int(off[m]), int(off[m + 1])) for m in use])
            ref_idx[owner, flat_pos - start] = idx[flat_pos]
        from . import model_wis
        g = model_wis.mod_coverage(_to_dev(t.cv, ctx), _to_dev(model_filt, ctx), _to_dev(ref_idx.view(np.int32), ctx),
                                   _to_dev(ref_len.view(np.int32), ctx), ctx=ctx)
        status = g["status"]
        outs = {kk: g[kk] for kk in ("ref_mean", "ref_sd", "cv_rel", "cvz", "cv2")}
        st = status.cpu().numpy()
        has = (st & 1) != 0
        h = {k: v.cpu().numpy() for k, v in outs.items()}
        # the output records: the input's, CV replaced in place, CVZ / CV2 appended (bcf_update_format semantics)
        t.cv = np.where(has, h["cv_rel"], t.cv).astype(np.float32)
        nf = t.norm_flags if t.norm_flags is not None else np.full(n, N.N2_CV, np.uint8)
        t.norm_flags = nf
        cv2_present = ((st & 2) != 0).astype(np.uint8)
        prev_cv2 = t.cv2
        extra = {"format": {"CVZ": (h["cvz"], has.astype(np.uint8))},
                 "info": {"ref_mean": h["ref_mean"], "ref_sd": h["ref_sd"], "ref_present": has.astype(np.uint8),
                          "filters": np.where(has, t.filters, t.filters | np.uint8(N.VF_NO_REF)).astype(np.uint8)}}
        if prev_cv2 is not None:   # the input already carries CV2 (MedianGcBinned): updated in place where written
            had = (nf & N.N2_CV2) != 0
            t.cv2 = np.where(cv2_present != 0, h["cv2"], prev_cv2).astype(np.float32)
            t.norm_flags = (nf | (cv2_present * N.N2_CV2)).astype(np.uint8)
            del had
            extra["format_order"] = [k for k in files.COVERAGE_FORMAT_ORDER if k != "FT"]
        else:                      # TotalCovSum input: CV2 is new and follows CVZ (push order :228-240)
            t.cv2 = None
            extra["format"]["CV2"] = (h["cv2"], cv2_present)
            extra["format_order"] = ["RCV", "RCVSD", "LCV", "LCVSD", "FRM", "MQ", "CV", "CVSD", "CVZ", "CV2"]
        header = vf.header + ["##cnvetti_cmdModCoverageVersion=0.1.0",
                              "##cnvetti_cmdModCoverageCommand=%s" % (command_line or "cnvetti cmd mod-coverage")]
        files.write_coverage_file(output, t, vf.samples, vf.contigs, is_targets=t.is_targets, header_lines=header,
                                  threads=threads, extra=extra)
        return dict(has_info=has, **h)


# ---- quick drivers ---------------------------------------------------------------------------------------
def _quick_coverage_options():
    """into_coverage_options of both quick commands (quick_wis_build_model.rs:44-70, quick_wis_call.rs:44-70)."""
    from .coverage import CoverageOptions
    return CoverageOptions(window_length=None, count_kind="Fragments", considered_regions="TargetRegions",
                           min_mapq=0, min_unclipped=0.6, min_window_remaining=0.5, min_raw_coverage=10)


def quick_wis_build_model(inputs, targets_bed, output, output_cov=None, ctx=None, threads=0, workdir=None):
    """`cnvetti quick wis-build-model --targets-bed T.bed --output model.bcf [--output-cov merged.bcf] X1.bam ...`:
    per BAM cmd coverage (fragments on targets) and cmd normalize (TotalCovSum) into a temporary directory, merge-cov,
    build-model-wis with the quick preset (max_samples_reliable 8)."""
    from . import model_wis
    ctx = ctx or default_context()
    with tempfile.TemporaryDirectory(prefix="cnvetti_quick_wis_build_model", dir=workdir) as tmp:
        norm_out = []
        for i, bam in enumerate(inputs):
            cov = os.path.join(tmp, "coverage.%d.bcf" % i)
            cmd_coverage(bam, cov, _quick_coverage_options(), targets_bed=targets_bed, ctx=ctx, threads=threads)
            nrm = os.path.join(tmp, "normalized.%d.bcf" % i)
            cmd_normalize(cov, nrm, "TotalCovSum", ctx=ctx, threads=threads)
            norm_out.append(nrm)
        merged = output_cov or os.path.join(tmp, "normalized.merged.bcf")
        cmd_merge_cov(norm_out, merged, ctx=ctx, threads=threads)
        return cmd_build_model_wis(merged, output, model_wis.BuildModelWisOptions.quick(), ctx=ctx, threads=threads)


def quick_wis_call(input_bam, input_model, output_targets, ctx=None, threads=0, workdir=None):
    """`cnvetti quick wis-call --input X.bam --input-model model.bcf --output-targets out.bcf`: cmd coverage on the
    model's targets (--wis-model-bcf), cmd normalize (TotalCovSum), cmd mod-coverage.  (The reference stops there:
    "Actual calling step has not been implemented yet!", quick_wis_call.rs:190; the IGV export is out of scope.)"""
    ctx = ctx or default_context()
    with tempfile.TemporaryDirectory(prefix="cnvetti_quick_wis_call", dir=workdir) as tmp:
        cov = os.path.join(tmp, "coverage.bcf")
        cmd_coverage(input_bam, cov, _quick_coverage_options(), wis_model_bcf=input_model, ctx=ctx, threads=threads)
        nrm = os.path.join(tmp, "normalized.bcf")
        cmd_normalize(cov, nrm, "TotalCovSum", ctx=ctx, threads=threads)
        return cmd_mod_coverage(nrm, input_model, output_targets, ctx=ctx, threads=threads)


__all__ = ["cmd_coverage", "cmd_normalize", "cmd_merge_cov", "cmd_build_model_wis", "cmd_mod_coverage",
           "quick_wis_build_model", "quick_wis_call", "load_bed_targets", "load_model_targets", "CoverageTable"]
