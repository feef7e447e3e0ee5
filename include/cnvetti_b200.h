/*
 * cnvetti_b200.h -- C ABI of the B200-native read-depth hot path of bihealth/cnvetti.
 *
 * The reference has no FFI; its seams for this path are a Rust trait and four `run` functions
 * (SURVEY.md section 8(b)).  Each entry point below names the reference interface it replaces
 * (paths relative to the reference checkout).  INTEGRATION.md shows the Rust `-sys` binding.
 *
 * Conventions
 *  - plain pointers and sizes; opaque handles; no exceptions cross the ABI;
 *  - every call returns 0 on success or a negative CNVB_ERR_* code; the human-readable error
 *    chain ("what: caused by ...", as error-chain prints in cnvetti/src/main.rs:171-185) is
 *    read with cnvb_last_error(ctx);
 *  - one cnvb_ctx = one CUDA device + one stream + scratch memory.  A ctx (and the handles made
 *    from it) is used by one host thread at a time; different ctxs may be used concurrently
 *    from different threads (the reference runs one `run` per rayon worker,
 *    cnvetti/src/quick_wis_build_model.rs:148-184).  No process-global mutable state;
 *  - `mem` arguments say where caller buffers live: CNVB_MEM_HOST (pageable or pinned host
 *    memory; the library stages through pinned buffers and copies asynchronously) or
 *    CNVB_MEM_DEVICE (device pointers on the ctx's device; zero-copy);
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *    CNVB_ERR_CUDA;
 *  - multi-GPU: one host process (or thread) per GPU, each with its own ctx and a cnvb_comm made
 *    from it; the *_sharded entry points are collective -- every rank calls them, in the same
 *    order, with its own share of the records / target rows.
 */
#ifndef CNVETTI_B200_H
#define CNVETTI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNVB_OK 0
#define CNVB_ERR_INVALID (-1)         /* bad argument                                        */
#define CNVB_ERR_CUDA (-2)            /* CUDA runtime error / no device                      */
#define CNVB_ERR_REFERENCE_PANIC (-3) /* input on which the reference itself panics          */
#define CNVB_ERR_STATE (-4)           /* call order violated (e.g. get_stats before finish)  */
#define CNVB_ERR_NOMEM (-5)
#define CNVB_ERR_COMM (-6)            /* NCCL missing or a collective failed                 */

#define CNVB_MEM_HOST 0
#define CNVB_MEM_DEVICE 1

/* f32::missing() of rust-htslib (bcf_float_missing) */
#define CNVB_F32_MISSING_BITS 0x7F800001u

/* ---- context ---------------------------------------------------------------------------- */
typedef struct cnvb_ctx cnvb_ctx;

/* `stream`: a cudaStream_t to run on (e.g. torch's current stream), or NULL to create one. */
int cnvb_ctx_create(int device, void *stream, cnvb_ctx **out);
void cnvb_ctx_destroy(cnvb_ctx *ctx);
int cnvb_ctx_synchronize(cnvb_ctx *ctx);
/* error chain of the last failing call on this ctx ("" if none); valid until the next call */
const char *cnvb_last_error(const cnvb_ctx *ctx);
/* error text of a failed cnvb_ctx_create (thread-local) */
const char *cnvb_create_error(void);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
uint64_t cnvb_ctx_launch_count(const cnvb_ctx *ctx);
const char *cnvb_version(void);
/* Returns the cached device memory of the ctx (pooled blocks not in use, scratch and staging
 * buffers) to the driver after synchronising the stream; the next calls allocate afresh. */
int cnvb_ctx_release_memory(cnvb_ctx *ctx);
/* Optional instrumentation: while on, every kernel launch is bracketed by CUDA events on the
 * ctx stream; cnvb_ctx_timing_query sums the device time of all launches of `kernel` (e.g.
 * "frag_bin_gw_kernel") since the last reset.  Query and reset synchronise the stream. */
int cnvb_ctx_set_timing(cnvb_ctx *ctx, int on);
int cnvb_ctx_timing_reset(cnvb_ctx *ctx);
int cnvb_ctx_timing_query(cnvb_ctx *ctx, const char *kernel, double *total_ms, uint64_t *count);

/* ---- communicator --------------------------------------------------
 * The reference is one process with rayon threads (lib-model-wis/src/lib.rs:144-193,
 * cnvetti/src/quick_wis_build_model.rs:148-184); here the host keeps that shape per GPU and the
 * exchanges between GPUs happen inside the library: NCCL over NVLink / NVSwitch, resolved at run
 * time (dlopen of libnccl.so.2, or the path in CNVB_NCCL_LIB; nothing is linked), enqueued on the
 * ctx stream.  Either let the library make the NCCL communicator -- rank 0 calls
 * cnvb_comm_unique_id and hands the 128 bytes to the other ranks by any means (a file, MPI, a
 * socket, torch.distributed), then every rank calls cnvb_comm_init_rank -- or wrap an
 * ncclComm_t the host already has (cnvb_comm_from_nccl; not destroyed with the handle).
 * world == 1 needs no NCCL at all. */
typedef struct cnvb_comm cnvb_comm;
typedef struct {
    char bytes[128]; /* an ncclUniqueId */
} cnvb_comm_id;
int cnvb_comm_unique_id(cnvb_comm_id *id);
int cnvb_comm_init_rank(cnvb_ctx *ctx, const cnvb_comm_id *id, int rank, int world, cnvb_comm **out);
int cnvb_comm_from_nccl(cnvb_ctx *ctx, void *nccl_comm, int rank, int world, cnvb_comm **out);
void cnvb_comm_destroy(cnvb_comm *comm);
int cnvb_comm_rank(const cnvb_comm *comm);
int cnvb_comm_world(const cnvb_comm *comm);
/* NCCL_VERSION_CODE of the library that was found; CNVB_ERR_COMM (text in
 * cnvb_comm_create_error, thread-local) if there is none */
int cnvb_comm_nccl_version(int *version);
const char *cnvb_comm_create_error(void);
/* Building blocks of the sharded drivers, for hosts that compose their own (device buffers):
 * all-gather of blocks of different sizes (recv = the blocks back to back in rank order; send may
 * be the rank's own slot of recv), all-to-all (block q of send, [send_off[q], send_off[q+1]), goes
 * to rank q; offsets in bytes, world + 1 entries), in-place sum of u32. */
int cnvb_comm_all_gather_v(cnvb_comm *comm, const void *send, void *recv, const uint64_t *bytes_per_rank);
int cnvb_comm_all_to_all_v(cnvb_comm *comm, const void *send, const uint64_t *send_off, void *recv,
                           const uint64_t *recv_off);
int cnvb_comm_all_reduce_u32(cnvb_comm *comm, uint32_t *buf, uint64_t count);

/* ---- read batches ------------------------------------------------------------------------ */
/* Extra bits the host decoder ORs into the 16-bit FLAG word (BAM uses bits 0..11):          */
#define CNVB_FLAG_CLIPFAIL 0x1000u     /* skip_clipping() is true   (agg.rs:174-193)          */
#define CNVB_FLAG_ISIZE_NONPOS 0x2000u /* insert_size() <= 0        (agg.rs:196-198)          */

/* Structure-of-arrays batch in coordinate order, as the host BAM decoder fills it
 * (north_star: tid, pos, end, mapq, flags, fragment midpoint).  Arrays a kernel does not need
 * may be NULL: GenomeWide reads mid/mapq/flags; TargetRegions pos/frag_end/mapq/flags;
 * Coverage pos/end/mapq/flags. */
typedef struct {
    uint64_t n;
    const int32_t *tid;      /* optional, unused by the kernels (one aggregator = one region) */
    const int32_t *pos;      /* record.pos()                                                  */
    const int32_t *end;      /* htslib bam_endpos(): pos + max(1, reference length)           */
    const int32_t *mid;      /* GenomeWide fragment centre, agg.rs:124-133 (bit pattern of u32)*/
    const int32_t *frag_end; /* TargetRegions interval end, agg.rs:312-320                    */
    const uint8_t *mapq;
    const uint16_t *flags;   /* FLAG | CNVB_FLAG_*                                            */
    int mem;                 /* CNVB_MEM_HOST | CNVB_MEM_DEVICE                               */
} cnvb_read_soa;

/* Host-side packer: BAM-level fields -> SoA fields, exactly as the reference derives them from
 * a bam::Record (agg.rs:124-133, 174-198, 311-321; htslib bam_endpos).  `cigar` is BAM-encoded
 * (len<<4|op), `cigar_off` has n+1 entries.  Pure host code (no device needed); this is what a
 * BGZF/BAM decoder thread calls per block of records. */
int cnvb_pack_reads(uint64_t n, const int32_t *pos, const int32_t *isize, const uint16_t *flag,
                    const uint64_t *cigar_off, const uint32_t *cigar, float min_unclipped,
                    int32_t *out_end, int32_t *out_mid, int32_t *out_frag_end,
                    uint16_t *out_flags);

/* ---- lib-coverage: BamRecordAggregator (lib-coverage/src/agg.rs:49-64) -------------------- */
typedef struct cnvb_cov cnvb_cov;

typedef struct {
    uint8_t min_mapq;           /* CoverageOptions::min_mapq            (options.rs:64)       */
    float min_unclipped;        /* CoverageOptions::min_unclipped       (options.rs:66)       */
    uint32_t window_length;     /* CoverageOptions::window_length       (options.rs:75)       */
    float min_window_remaining; /* CoverageOptions::min_window_remaining (options.rs:69)      */
    uint32_t min_raw_coverage;  /* CoverageOptions::min_raw_coverage    (options.rs:71)       */
    uint32_t plp_flag_mask;     /* flags htslib's pileup drops before the reference's filter;
                                   0 selects the default 0x704 (UNMAP|SECONDARY|QCFAIL|DUP)   */
} cnvb_cov_opts;

/* (count_kind, considered_regions) pairs accepted at lib-coverage/src/lib.rs:510-529 */
#define CNVB_FRAGMENTS_GENOME_WIDE 0    /* FragmentsGenomeWideAggregator   agg.rs:68-257     */
#define CNVB_FRAGMENTS_TARGET_REGIONS 1 /* FragmentsTargetRegionsAggregator agg.rs:261-444   */
#define CNVB_COVERAGE 2                 /* CoverageAggregator              agg.rs:468-600    */

/* = XxxAggregator::new (agg.rs:90, :278, :485).  `region_end` is the `contig_length` argument
 * the reference passes (the region end, lib.rs:514,526).  Targets (sorted by start, as a tabix
 * BED yields them) only for TARGET_REGIONS; piles (disjoint, sorted) only for GENOME_WIDE.
 * Target/pile arrays are host memory and are copied. */
int cnvb_cov_create(cnvb_ctx *ctx, const cnvb_cov_opts *opts, int kind, uint32_t region_end,
                    const uint32_t *tgt_start, const uint32_t *tgt_end, uint32_t n_tgt,
                    const uint32_t *pile_start, const uint32_t *pile_end, uint32_t n_pile,
                    cnvb_cov **out);
/* = put_fetched_records (agg.rs:51): batches arrive in coordinate order; may be called
 * repeatedly.  Asynchronous with respect to the host for CNVB_MEM_DEVICE batches. */
int cnvb_cov_put_records(cnvb_cov *agg, const cnvb_read_soa *batch);
/* end of the fetch()ed stream: flushes, runs the Coverage window pass (incl. the reference's
 * lagging window state machine, agg.rs:523-573) and checks for reference-panic conditions. */
int cnvb_cov_finish(cnvb_cov *agg);
/* Same, but only enqueues the work on the ctx stream: get_stats(CNVB_MEM_DEVICE) may follow at
 * once; tallies and reference-panic conditions are reported by a later cnvb_cov_finish. */
int cnvb_cov_finish_async(cnvb_cov *agg);
uint64_t cnvb_cov_num_regions(const cnvb_cov *agg);     /* agg.rs:57  */
uint32_t cnvb_cov_num_processed(const cnvb_cov *agg);   /* agg.rs:60  (valid after finish)   */
uint32_t cnvb_cov_num_skipped(const cnvb_cov *agg);     /* agg.rs:63  (valid after finish)   */
/* = get_stats for every region id (agg.rs:54, AggregationStats agg.rs:17-32), arrays of
 * num_regions.  cov_sd / frac_removed are the Option<f32> members: pass NULL to skip; asking
 * for one the aggregator kind does not produce is CNVB_ERR_INVALID.  mean_mapq may be NULL. */
int cnvb_cov_get_stats(cnvb_cov *agg, float *cov, float *cov_sd, uint32_t *start, uint32_t *end,
                       float *frac_removed, float *mean_mapq, int mem);
/* raw integer counters (Fragments kinds), for bit-exact checks: u32[num_regions] */
int cnvb_cov_get_counts(cnvb_cov *agg, uint32_t *counts, int mem);
void cnvb_cov_destroy(cnvb_cov *agg);

/* = ReferenceStats::look_at_chars (lib-coverage/src/reference.rs:84-130).
 * gc[ceil(len/wl)] (missing pattern where a window has no ACGT), has_gap[ceil(len/wl)]. */
int cnvb_refstats(cnvb_ctx *ctx, const uint8_t *seq, uint64_t len, uint32_t window_length,
                  float *gc, uint8_t *has_gap, int mem);

/* = FORMAT value assembly in process_region (lib-coverage/src/lib.rs:623-668):
 * LCV = RCV/(end-start), LCVSD, FT bits (CNVB_FT_LOWCOV, CNVB_FT_PILE).
 * cov_sd/lcvsd and frac_removed may be NULL. */
#define CNVB_FT_LOWCOV 1
#define CNVB_FT_PILE 2
int cnvb_cov_records(cnvb_ctx *ctx, uint64_t n, const float *cov, const float *cov_sd,
                     const uint32_t *start, const uint32_t *end, const float *frac_removed,
                     int is_targets, const cnvb_cov_opts *opts, float *lcv, float *lcvsd,
                     uint8_t *ft, int mem);

/* ---- lib-coverage: the region loop --------------------------------------------------------- */
/* One GenomeRegions entry (lib-shared/src/regions.rs:15-19) with its inputs. */
typedef struct {
    uint32_t region_end;       /* end of the region == the `contig_length` the aggregators get  */
    cnvb_read_soa reads;       /* what fetch() yields for the region, coordinate order          */
    const uint8_t *seq;        /* contig bytes for GC/GAP (lib.rs:440-454) or NULL              */
    uint64_t seq_len;
    int seq_mem;               /* CNVB_MEM_HOST | CNVB_MEM_DEVICE                               */
    const uint32_t *tgt_start, *tgt_end; /* TARGET_REGIONS only (host)                          */
    uint32_t n_tgt;
    const uint32_t *pile_start, *pile_end; /* GENOME_WIDE only (host)                           */
    uint32_t n_pile;
} cnvb_region;

/* One column per BCF tag `cmd coverage` writes, rows of all regions concatenated in region
 * order; NULL = not wanted.  start/end/rcv/lcv are required.  rcvsd/lcvsd only for COVERAGE,
 * frm only for GENOME_WIDE, gc/gap (together) only for windowed kinds. */
typedef struct {
    uint32_t *start, *end;
    float *rcv, *rcvsd, *lcv, *lcvsd, *frm, *gc, *mean_mapq;
    uint8_t *gap, *ft;
} cnvb_cov_columns;

/* number of table rows the regions produce (windows or targets) */
uint64_t cnvb_coverage_rows(const cnvb_cov_opts *opts, int kind, const cnvb_region *regions,
                            uint32_t n_regions);
/* = the region loop of lib_coverage::run (lib-coverage/src/lib.rs:768-781) with process_region
 * (:422-676) per region: refstats, aggregator, get_stats, FORMAT assembly -- enqueued back to
 * back, one host wait at the end.  `mem` is where the columns live; num_processed/num_skipped
 * (host, n_regions entries, may be NULL) receive the per-region tallies.  A region whose input
 * makes the reference panic fails the call with CNVB_ERR_REFERENCE_PANIC ("region i caused by"). */
int cnvb_coverage_run(cnvb_ctx *ctx, const cnvb_cov_opts *opts, int kind, const cnvb_region *regions,
                      uint32_t n_regions, const cnvb_cov_columns *out, uint64_t *num_processed,
                      uint64_t *num_skipped, int mem);
/* The same without the host wait, for device-resident inputs and columns: every stage is enqueued and
 * the call returns; the columns are valid in stream order (a following cnvb_norm_* call on the same
 * context may consume them at once).  cnvb_coverage_wait blocks until the regions' tallies have
 * arrived, fills num_processed / num_skipped (n_regions entries each, may be NULL), reports the first
 * region on which the reference would have panicked exactly as cnvb_coverage_run does, and frees the
 * pending state -- it must be called once per successful cnvb_coverage_run_async.  At most 4096
 * regions may be pending per context (the tallies travel through a ring of pinned slots). */
typedef struct cnvb_cov_pending cnvb_cov_pending;
int cnvb_coverage_run_async(cnvb_ctx *ctx, const cnvb_cov_opts *opts, int kind, const cnvb_region *regions,
                            uint32_t n_regions, const cnvb_cov_columns *out, cnvb_cov_pending **pending);
int cnvb_coverage_wait(cnvb_cov_pending *pending, uint64_t *num_processed, uint64_t *num_skipped);
/* The loop of lib_coverage::run (lib.rs:768-781) over the SAME device buffers again and again -- a pipeline that copies
 * every sample into fixed device columns; the strong-scaling step of bench.py -- as a CUDA graph.  plan_create runs
 * the loop once eagerly and once under stream capture (count kind FragmentsGenomeWide, device-resident inputs and
 * columns, no piles: the other kinds size their launches from the records) and instantiates the graph; plan_launch enqueues one run on
 * the context's stream with a single launch and returns; plan_wait blocks until the plan's launches so far are done
 * (not for work enqueued behind them: two plans can alternate, one waited for while the other runs) and reports the
 * tallies and reference-panic checks of the last one as cnvb_coverage_run does.  The plan owns everything
 * the graph touches; the regions' read columns, sequences and the output columns are the caller's and must stay
 * where they are (their CONTENTS may change between launches, their addresses and lengths may not). */
typedef struct cnvb_cov_plan cnvb_cov_plan;
int cnvb_coverage_plan_create(cnvb_ctx *ctx, const cnvb_cov_opts *opts, int kind, const cnvb_region *regions,
                              uint32_t n_regions, const cnvb_cov_columns *out, cnvb_cov_plan **plan);
int cnvb_coverage_plan_launch(cnvb_cov_plan *plan);
int cnvb_coverage_plan_wait(cnvb_cov_plan *plan, uint64_t *num_processed, uint64_t *num_skipped);
void cnvb_coverage_plan_destroy(cnvb_cov_plan *plan);

/* ---- lib-coverage: pile collection for --mask-piles --------------------------------------- */
typedef struct cnvb_piles cnvb_piles;
/* = PileCollector::collect_piles (lib-coverage/src/piles.rs:31-127) for the reads fetch() yields
 * for one contig (coordinate order; pos, end, mapq, flags): maximal runs of covered bases with
 * gaps <= pile_max_gap, size = largest depth.  Depth counts the alignments of a pileup column
 * (htslib's column mask = opts->plp_flag_mask, 0 = default 0x704) that pass piles.rs:45-58 with
 * opts->min_mapq.  region_end = contig length: an alignment reaching past it makes the reference
 * panic (CNVB_ERR_REFERENCE_PANIC).  The result is an opaque list (device memory). */
int cnvb_piles_collect(cnvb_ctx *ctx, const cnvb_read_soa *reads, const cnvb_cov_opts *opts,
                       uint32_t pile_max_gap, uint32_t region_end, cnvb_piles **out);
uint64_t cnvb_piles_count(const cnvb_piles *piles);
/* PileInfo { start, end, size } (piles.rs:67-75) of every pile, in position order; any of the
 * three may be NULL */
int cnvb_piles_get(cnvb_piles *piles, uint32_t *start, uint32_t *end, uint32_t *size, int mem);
void cnvb_piles_destroy(cnvb_piles *piles);
/* = compute_threshold (lib-coverage/src/lib.rs:265-351): depth threshold from the histogram of
 * pile sizes (abs_freq[size] = number of piles, all contigs) at the given FDR.  Pure host
 * arithmetic, no context; on failure the reference's bail! text is copied to err (may be NULL). */
int cnvb_pile_threshold(const uint64_t *abs_freq, uint64_t n, double fdr_cutoff, uint32_t *threshold,
                        char *err, uint64_t err_len);

/* ---- host BGZF/BAM decoder ------------------------------------------
 * Replaces, for the coverage path, rust-htslib's bam::IndexedReader + fetch(tid, 0, len) +
 * read(&mut record) per region (lib-coverage/src/lib.rs:533-547, agg.rs:224-228) together with
 * the per-record derivations of cnvb_pack_reads: one pass over the file, BGZF blocks inflated on
 * n_threads host threads (<= 0: all cores), records turned into per-contig SoA columns in file
 * (= coordinate) order.  The columns of contig tid are what `fetch` on the whole contig yields;
 * records with refID < 0 are counted and dropped.  Pure host code: no context, no device.
 * cnvb_bam_open reads the header only; on failure *out still holds a handle whose
 * cnvb_bam_error() text says why (close it). */
typedef struct cnvb_bam cnvb_bam;
int cnvb_bam_open(const char *path, int n_threads, cnvb_bam **out);
const char *cnvb_bam_error(const cnvb_bam *b);
uint32_t cnvb_bam_n_refs(const cnvb_bam *b);
const char *cnvb_bam_ref_name(const cnvb_bam *b, uint32_t tid);
uint32_t cnvb_bam_ref_len(const cnvb_bam *b, uint32_t tid);
/* SAM header text (@RG SM: = the sample name of the output, lib-coverage/src/lib.rs:36-64) */
const char *cnvb_bam_header_text(const cnvb_bam *b, uint64_t *len);
/* Decode the whole file.  pinned != 0: page-locked columns (cudaHostAlloc; plain memory if no
 * CUDA runtime is usable).  min_unclipped as in cnvb_pack_reads. */
int cnvb_bam_decode(cnvb_bam *b, float min_unclipped, int pinned);
/* The decoded columns of one contig (pointers stay valid until the next decode / close). */
/* = bam::IndexedReader::fetch(tid, start, end) + the record loop (lib-coverage/src/lib.rs:533-547)
 * through the `.bai` beside the file (<path>.bai or <path minus .bam>.bai): only the blocks from the
 * first candidate chunk of the overlapping bins on are read and inflated, reading stops at the first
 * record past the interval; the records whose alignment [pos, bam_endpos) overlaps [start, end)
 * (0-based, half-open) replace the columns of `tid`.  Without an index: CNVB_ERR_INVALID ("Could not
 * open BAI-indexed BAM file", lib.rs:531-532). */
int cnvb_bam_fetch(cnvb_bam *bam, uint32_t tid, uint32_t start, uint32_t end, float min_unclipped, int pinned);
int cnvb_bam_ref_reads(const cnvb_bam *b, uint32_t tid, cnvb_read_soa *out);
/* counts: records decoded | records without reference | compressed bytes | uncompressed bytes;
 * seconds: whole decode | inflate phases | record walk + SoA derivation (wall clock). */
int cnvb_bam_stats(const cnvb_bam *b, uint64_t counts[4], double seconds[3]);
void cnvb_bam_close(cnvb_bam *b);

/* ---- lib-normalize ------------------------------------------------------------------------ */
/* = run_uncorrected_normalization (lib-normalize/src/uncorrected.rs:20-96).  lcvsd/cvsd may be
 * NULL.  *lcv_sum (host, may be NULL) receives the f64 sum. */
int cnvb_norm_total(cnvb_ctx *ctx, const float *lcv, const float *lcvsd, uint64_t n, float *cv,
                    float *cvsd, double *lcv_sum, int mem);
/* = run_binned_correction (lib-normalize/src/correct_binned.rs:25-146).  flags bits say which
 * FORMAT/INFO fields the reference writes for the record.  medians (host, 101 floats) may be
 * NULL.  lcvsd/cvsd may be NULL. */
#define CNVB_N2_CV 1
#define CNVB_N2_CV2 2
#define CNVB_N2_CVSD 4
#define CNVB_N2_FEW_GCWINDOWS 8
int cnvb_norm_gc_median(cnvb_ctx *ctx, const float *lcv, const float *lcvsd, const float *gc,
                        uint64_t n, float *cv, float *cv2, float *cvsd, uint8_t *flags,
                        float *medians, int mem);

/* The two normalisers with the RECORDS sharded over the ranks of `comm` (contigs on different
 * GPUs, lib-coverage/src/lib.rs:768-781): arguments as above for the rank's own records, results
 * for them.  The statistics run over all records of all ranks -- TotalCovSum exchanges one
 * double-double partial sum per rank (uncorrected.rs:20-49), MedianGcBinned sums the 101 x 256
 * radix histograms of the ranks in each of its four passes (correct_binned.rs:25-75) -- and are
 * bit-identical to the single-GPU call on the concatenated records.  Collective. */
int cnvb_norm_total_sharded(cnvb_comm *comm, const float *lcv, const float *lcvsd, uint64_t n, float *cv,
                            float *cvsd, double *lcv_sum, int mem);
int cnvb_norm_gc_median_sharded(cnvb_comm *comm, const float *lcv, const float *lcvsd, const float *gc,
                                uint64_t n, float *cv, float *cv2, float *cvsd, uint8_t *flags,
                                float *medians, int mem);

/* ---- variant files: VCF / BCF reader and writer + index ----------------
 * Replaces what the commands of the path get from rust-htslib's bcf module: the record assembly
 * and output of process_region (lib-coverage/src/lib.rs:560-673) under build_header (:67-166), the
 * read / translate / push / write loops of lib-normalize (src/shared.rs:16-75, uncorrected.rs:
 * 51-96, correct_binned.rs:77-146), lib-merge-cov (src/lib.rs:102-232), lib-model-wis (read_coverage_bcf
 * src/lib.rs:65-125, write_output :340-419) and lib-mod-cov (src/lib.rs:33-60, 89-164, 167-248), and
 * build_index (lib-shared/src/bcf_utils.rs:47-82).  The container follows the file name exactly as
 * there: `.bcf` = BGZF-compressed binary BCF2.2 (+ `.csi`, min_shift 14), `.vcf.gz` = BGZF text
 * (+ `.tbi`), `.vcf` = plain text, any other name = uncompressed BCF; neither of the last two gets an
 * index.  The reader takes all four, telling them apart by content.
 *
 * The schema is the closed tag set of the two headers of the path (coverage: lib-coverage/src/lib.rs:
 * 99-160; model: lib-model-wis/src/lib.rs:308-322): records are columns -- one host array per tag --
 * so that they move to and from the kernels without a per-record object.  Every record of these
 * files has REF N, one symbolic ALT (<WINDOW> | <TARGET>), no QUAL and the ID `chrom:pos+1-end`
 * (lib.rs:586-588, lib-model-wis/src/lib.rs:359-361).  Tags outside the schema are skipped on
 * reading.  Floats are stored exactly in BCF and printed with six significant digits in VCF (as
 * htslib does): only BCF chains the commands bit for bit.  Pure host code (threads for record
 * formatting / parsing and for BGZF blocks). */
#define CNVB_VF_PASS 1        /* FILTER bits of a record (written in this order)                */
#define CNVB_VF_PILE 2
#define CNVB_VF_LOWCOV 4
#define CNVB_VF_NO_REF 8
#define CNVB_VF_FEW_REF 16
#define CNVB_VF_UNRELIABLE 32
#define CNVB_VF_OTHER 128     /* reading: a filter outside the schema                           */
typedef struct {
    const char *key;          /* FORMAT key of a Float tag (RCV, LCV, CV, ...)                  */
    const float *values;      /* [n_rows][n_samples]                                            */
    const uint8_t *present;   /* [n_rows] or NULL (every record carries the key)                */
} cnvb_vfile_fmt;
typedef struct {
    const char *path;
    const char *header_text;  /* the `##` meta lines, '\n'-separated: ##fileformat and the PASS
                                 filter are added when missing; the dictionary (IDX) follows the
                                 order of the lines                                             */
    uint32_t n_samples;
    const char *const *samples;
    uint64_t n_rows;
    const int32_t *rid;       /* index into the header's ##contig lines                         */
    const int32_t *pos;       /* 0-based start                                                  */
    const int32_t *end;       /* INFO/END; also the record length (rlen = end - pos)            */
    int is_targets;           /* ALT <TARGET> instead of <WINDOW>                               */
    const uint8_t *filters;   /* CNVB_VF_* bits per row, or NULL                                */
    /* INFO, in this order; NULL = the tag is not written */
    const uint8_t *gap;       /* flag GAP (only with gc)                                        */
    const float *gc;
    const float *mean_mapq;
    const uint8_t *few_gc;    /* flag FEW_GCWINDOWS                                             */
    const float *ref_mean, *ref_sd;
    const uint8_t *ref_present; /* rows that carry REF_MEAN / REF_STD_DEV (NULL: all)           */
    const int32_t *num_calls;
    const uint64_t *ref_off;  /* REF_TARGETS: CSR over the rows of this table, written as their IDs */
    const uint32_t *ref_idx;
    /* FORMAT: GT (always the no-call) first, then the keys of `fmt` in order, FT before fmt[ft_pos] */
    uint32_t n_fmt;
    const cnvb_vfile_fmt *fmt;
    const uint8_t *ft;        /* [n_rows][n_samples] CNVB_FT_* bits or NULL; a record carries FT
                                 when any sample has a bit set (lib.rs:659-664)                 */
    uint32_t ft_pos;
    int write_index;          /* build_index: .csi beside .bcf, .tbi beside .vcf.gz             */
    int n_threads;            /* <= 0: all cores                                                */
} cnvb_vfile_write_args;
int cnvb_vfile_write(const cnvb_vfile_write_args *args, char *err, uint64_t err_len);

typedef struct cnvb_vfile cnvb_vfile;
int cnvb_vfile_open(const char *path, int n_threads, cnvb_vfile **out, char *err, uint64_t err_len);
void cnvb_vfile_close(cnvb_vfile *f);
const char *cnvb_vfile_last_error(const cnvb_vfile *f);
uint64_t cnvb_vfile_num_records(const cnvb_vfile *f);
uint32_t cnvb_vfile_num_samples(const cnvb_vfile *f);
const char *cnvb_vfile_sample(const cnvb_vfile *f, uint32_t i);
uint32_t cnvb_vfile_num_contigs(const cnvb_vfile *f);
const char *cnvb_vfile_contig_name(const cnvb_vfile *f, uint32_t i);
uint32_t cnvb_vfile_contig_length(const cnvb_vfile *f, uint32_t i);
const char *cnvb_vfile_header_text(const cnvb_vfile *f); /* the ## lines as read (without IDX keys) */
int cnvb_vfile_is_bcf(const cnvb_vfile *f);
/* the fixed columns of every record; any pointer may be NULL.  alt: 0 <WINDOW>, 1 <TARGET>, 2 other;
 * end = INFO/END (pos + length of REF without it); canonical_ids: 1 if every ID is chrom:pos+1-end */
int cnvb_vfile_fixed(cnvb_vfile *f, int32_t *rid, int32_t *pos, int32_t *end, uint8_t *alt, uint8_t *filters,
                     int *canonical_ids);
/* one tag as a column; `present` (nullable) says which records carry it.  Absent values are left as the
 * BCF missing pattern (floats 0x7F800001, ints INT32_MIN); a key the header does not define is
 * CNVB_ERR_INVALID.  Numeric INFO tags yield their first value; FORMAT tags the first value per sample */
int cnvb_vfile_info_float(cnvb_vfile *f, const char *key, float *out, uint8_t *present);
int cnvb_vfile_info_int(cnvb_vfile *f, const char *key, int32_t *out, uint8_t *present);
int cnvb_vfile_info_flag(cnvb_vfile *f, const char *key, uint8_t *out);
int cnvb_vfile_format_float(cnvb_vfile *f, const char *key, float *out, uint8_t *present);
int cnvb_vfile_format_ft(cnvb_vfile *f, uint8_t *out, uint8_t *present);
/* for every record of `f` the row of `other` with the same ID (UINT32_MAX: none) -- the by-name lookups
 * of lib-mod-cov (src/lib.rs:121-123, 146-152) */
int cnvb_vfile_match_ids(cnvb_vfile *f, cnvb_vfile *other, uint32_t *row);
/* INFO/REF_TARGETS of a model file, the names resolved to rows of `against` (the file itself when NULL):
 * off[n + 1], idx[off[n]].  Call with idx = NULL to size: only `off` is filled.  A name `against` does not
 * hold is CNVB_ERR_REFERENCE_PANIC ("Could not find: ...", lib-mod-cov/src/lib.rs:149) */
int cnvb_vfile_ref_targets(cnvb_vfile *f, cnvb_vfile *against, uint64_t *off, uint32_t *idx, uint64_t idx_cap);

/* ---- lib-merge-cov --------------------------------------------------------------------------
 * = merge_files (lib-merge-cov/src/lib.rs:102-232) on in-memory columns: the FORMAT/CV columns of S
 * single-sample tables that describe the same T records, side by side -> X[T][S] row-major, the
 * record layout of the merged file and ModelInput's panel (lib-model-wis/src/lib.rs:49-62).
 * Pass EITHER cv_cols (host array of S device pointers, one column of T floats each) OR
 * cv_sample_major (one device block, column s at cv_sample_major + s * col_stride).  Device
 * memory only; asynchronous on the ctx stream. */
int cnvb_merge_cov(cnvb_ctx *ctx, const float *const *cv_cols, const float *cv_sample_major, uint64_t col_stride,
                   uint32_t S, uint64_t T, float *X);

/* ---- lib-model-wis ------------------------------------------------------------------------ */
typedef struct {
    double filter_z_score;         /* --filter-z-score        5.64  (cli.yaml:370-376)        */
    double filter_rel;             /* --filter-rel            0.35  (cli.yaml:377-383)        */
    uint32_t min_ref_targets;      /* --min-ref-targets       10    (cli.yaml:384-390)        */
    uint32_t max_ref_targets;      /* --max-ref-targets       100   (cli.yaml:391-397)        */
    uint32_t max_samples_reliable; /* --max-samples-reliable  4     (cli.yaml:398-406)        */
    uint32_t engine;               /* CNVB_WIS_ENGINE_*                                       */
} cnvb_wis_opts;
#define CNVB_WIS_ENGINE_AUTO 0   /* tensor for large panels, exact for small ones             */
#define CNVB_WIS_ENGINE_TENSOR 1 /* tcgen05 fp16 contraction as candidate filter + exact
                                    f32-sequential re-scoring of the candidates               */
#define CNVB_WIS_ENGINE_EXACT 2  /* every pair in emulated f32-sequential arithmetic          */

/* = compute_distances (lib-model-wis/src/lib.rs:128-197) for target rows [row_begin,row_end)
 * of the panel X[T][S] (row-major f32, NaN-free, FORMAT/CV of the merged BCF).
 * k = min(T, max_ref_targets).  Outputs [row_end-row_begin][k]: ascending distance, ties by
 * index (the reference's order among equal distances is pdqselect's, which no test pins);
 * same-contig pairs (tid equal, incl. the target itself) have distance +inf exactly as in the
 * reference and appear only when fewer than k other-contig targets exist.
 * The distance is sqrt_f32 of the f32 sum accumulated over samples in order with separate
 * mul and add roundings (lib.rs:157-162) -- bit-exact, whichever engine is used. */
int cnvb_wis_distances(cnvb_ctx *ctx, const float *X, const uint32_t *tid, uint32_t T, uint32_t S,
                       const cnvb_wis_opts *opts, uint32_t row_begin, uint32_t row_end,
                       float *ref_dist, uint32_t *ref_idx, int mem);
/* diagnostics of the last cnvb_wis_distances / cnvb_wis_build on this ctx:
 * out[0] rows computed, out[1] rows that fell back to the exact row kernel, out[2] pairs
 * re-scored exactly (tensor engine), out[3] 256x128 tiles the tensor sweep visited */
int cnvb_wis_last_stats(const cnvb_ctx *ctx, uint64_t out[4]);
/* = threshold of prune_distances (lib.rs:207-216): (mean + z*sd) as f32 over the nearest
 * distance of ALL T targets (f64, exact-sum mean, n-1 SD).  `nearest` = ref_dist[.][0].
 * NaN for T < 2 (the reference returns an empty Vec).  *threshold is host memory. */
int cnvb_wis_threshold(cnvb_ctx *ctx, const float *nearest, uint32_t T, uint32_t stride,
                       double filter_z_score, float *threshold, int mem);
/* = the filter of prune_distances (lib.rs:221-228) on `rows` rows: compacts ref_idx in place,
 * ref_len[r] = number of reference targets kept (distance <= threshold). */
int cnvb_wis_prune(cnvb_ctx *ctx, const float *ref_dist, uint32_t *ref_idx, uint32_t rows, uint32_t k,
                   float threshold, uint32_t *ref_len, int mem);
/* = count_per_probe_calls (lib.rs:235-271) for target rows [row_begin,row_end); ref_idx /
 * ref_len / num_calls are indexed relative to row_begin. */
int cnvb_wis_calls(cnvb_ctx *ctx, const float *X, uint32_t T, uint32_t S, const uint32_t *ref_idx,
                   const uint32_t *ref_len, uint32_t k, const cnvb_wis_opts *opts,
                   uint32_t row_begin, uint32_t row_end, uint32_t *num_calls, int mem);
/* = FILTER column of write_output (lib.rs:362-377) */
#define CNVB_WIS_FEW_REF 1
#define CNVB_WIS_UNRELIABLE 2
int cnvb_wis_filters(cnvb_ctx *ctx, const uint32_t *ref_len, const uint32_t *num_calls, uint32_t rows,
                     const cnvb_wis_opts *opts, uint8_t *filt, int mem);
/* = lib_model_wis::run (lib.rs:422-447) on an in-memory panel, single GPU: distances, prune,
 * per-probe calls, filters.  Outputs as above for all T rows; *threshold is host memory. */
int cnvb_wis_build(cnvb_ctx *ctx, const float *X, const uint32_t *tid, uint32_t T, uint32_t S,
                   const cnvb_wis_opts *opts, float *ref_dist, uint32_t *ref_idx, uint32_t *ref_len,
                   uint32_t *num_calls, uint8_t *filt, float *threshold, int mem);

/* = lib_model_wis::run (lib.rs:422-447) with the target rows sharded over the ranks of `comm`
 * ("cnvb_wis_build(..., n_gpus)").  Every rank passes ITS rows of the panel
 * (X_rows[rows][S], tid_rows[rows]; row blocks in rank order = target order; sizes may differ, a
 * rank may have none) and receives the model of its rows -- outputs as cnvb_wis_build, sized by
 * `rows`, ref_idx holding GLOBAL target indices; *row_begin / *T (host, may be NULL) say where the
 * block sits.  Inside: all-gather of the panel over NVLink, distances for the own rows, all-gather
 * of one nearest distance per target, the same threshold on every rank (*threshold, host), local
 * prune / per-probe calls / filters.  Bit-identical to the single-GPU model.  Collective. */
int cnvb_wis_build_sharded(cnvb_comm *comm, const float *X_rows, const uint32_t *tid_rows, uint32_t rows,
                           uint32_t S, const cnvb_wis_opts *opts, float *ref_dist, uint32_t *ref_idx,
                           uint32_t *ref_len, uint32_t *num_calls, uint8_t *filt, float *threshold,
                           uint32_t *row_begin, uint32_t *T, int mem);

/* ---- lib-mod-cov ---------------------------------------------------------------------------- */
/* = load_target_infos + the FORMAT/INFO values of write_mod_cov (lib-mod-cov/src/lib.rs:76-164,
 * 209-240) for one sample: cv[T] = FORMAT/CV of the sample, model_filt[T] != 0 marks model
 * records carrying a non-PASS filter.  status bit 0: probe info exists (else FILTER NO_REF),
 * bit 1: CV2 written.  Values without probe info are the missing pattern. */
#define CNVB_MODCOV_HAS_INFO 1
#define CNVB_MODCOV_CV2 2
int cnvb_modcov(cnvb_ctx *ctx, const float *cv, uint32_t T, const uint8_t *model_filt,
                const uint32_t *ref_idx, const uint32_t *ref_len, uint32_t k, float *ref_mean,
                float *ref_sd, float *cv_rel, float *cvz, float *cv2, uint8_t *status, int mem);
/* The same for S samples in one call (the per-sample loop of `quick wis-call` over a cohort,
 * cnvetti/src/quick_wis_call.rs:171-181): X[T][S] row-major holds FORMAT/CV of the samples side
 * by side (the merge-cov layout, lib-merge-cov/src/lib.rs:102-232).  Target rows [row_begin,
 * row_end) are computed (a GPU's share when targets are sharded); model_filt / ref_idx / ref_len
 * and the outputs are indexed relative to row_begin.  Outputs are sample-major [S][rows] --
 * row s is exactly what cnvb_modcov writes for cv = X[:, s] -- and each may be NULL. */
int cnvb_modcov_panel(cnvb_ctx *ctx, const float *X, uint32_t T, uint32_t S, uint32_t row_begin,
                      uint32_t row_end, const uint8_t *model_filt, const uint32_t *ref_idx,
                      const uint32_t *ref_len, uint32_t k, float *ref_mean, float *ref_sd,
                      float *cv_rel, float *cvz, float *cv2, uint8_t *status, int mem);
/* count_per_probe_calls (lib-model-wis/src/lib.rs:235-271), the FILTER column (:362-377) and
 * cnvb_modcov_panel in one pass over the panel, for calling a cohort against the model built from
 * it: the two stages gather the same reference rows and form the same mean / sd, and at cohort
 * scale the gathers are the cost.  num_calls / filt are what cnvb_wis_calls + cnvb_wis_filters give
 * for rows [row_begin, row_end); the value outputs (each may be NULL) are those of
 * cnvb_modcov_panel with model_filt = filt. */
int cnvb_wis_calls_modcov_panel(cnvb_ctx *ctx, const float *X, uint32_t T, uint32_t S,
                                const uint32_t *ref_idx, const uint32_t *ref_len, uint32_t k,
                                const cnvb_wis_opts *opts, uint32_t row_begin, uint32_t row_end,
                                uint32_t *num_calls, uint8_t *filt, float *ref_mean, float *ref_sd,
                                float *cv_rel, float *cvz, float *cv2, uint8_t *status, int mem);

/* ---- lib-discover (builder-defined: the reference has no implementation, only the CLI stub
 * cnvetti/src/cli.yaml:457-469 and `bail!("cmd discover not implemented!")`,
 * cnvetti/src/main.rs:135; spec in DESIGN.md "HMM") ------------------------------------------- */
#define CNVB_HMM_STATES 5 /* copy number 0..4 */
typedef struct {
    double sigma0;  /* additive emission SD               (default 0.02) */
    double eps_cn;  /* copy number used for state CN0     (default 0.1)  */
    double p_move;  /* probability of each state change   (default 1e-4) */
} cnvb_hmm_opts;
/* Segmentation of `n_lanes` independent lanes (one per sample and contig).  cv holds the lanes'
 * FORMAT/CV values back to back, lane l = cv[lane_off[l] .. lane_off[l+1]); sigma_s[l] is the
 * lane's noise scale (1.4826 * MAD of CV).  Non-finite CV (missing) emits log-likelihood 0 in
 * every state.  state[i] = Viterbi copy-number state on INTEGER scores in units of 2^-16 nat
 * (transitions llrint(log p * 65536); emissions from four f32 operations, floored at -2^22 and
 * rounded to nearest-even; recurrence max_r(delta[r] + t[r][s]) + e[s], ties -> lowest
 * predecessor, trace-back from the first maximum: DESIGN.md section 6).  Integer (max,+) is
 * associative, so the chunk-parallel evaluation on the GPU and a sequential walk give the same
 * path, bit for bit.  posterior (nullable) = [total bins][5] forward-backward posteriors in f64
 * log space (rel 1e-6).  lane_off and sigma_s are host arrays; cv/state/posterior follow `mem`. */
int cnvb_hmm_segment(cnvb_ctx *ctx, const float *cv, const uint64_t *lane_off, uint32_t n_lanes,
                     const double *sigma_s, const cnvb_hmm_opts *opts, uint8_t *state, double *posterior,
                     int mem);

#ifdef __cplusplus
}
#endif
#endif /* CNVETTI_B200_H */
