/*
 * cnv_oracle.h -- CPU restatement ("oracle") of the reference's read-depth hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The shipped path
 * (cnvetti_b200/) never links, imports or calls anything in this directory.
 *
 * The reference (bihealth/cnvetti v0.1.0, Rust) cannot be compiled in this environment (no
 * cargo/rustc, no htslib), so every function below is a literal plain-C
 * restatement of the cited reference lines.  Parity pinning:
 *   - stats (S1): PINNED by the reference's own 17 known-answer tests
 *     (lib-shared/src/stats.rs:377-907), extracted into tests/golden/stats_kat.json.
 *   - pile threshold: PINNED by lib-coverage/src/lib.rs:797-807 (one vector).
 *   - A2..A6, N1, N2, W2..W5, M1: "parity unpinned" -- the reference holds no test or fixture
 *     for them; they follow the cited source lines statement by statement.
 *   - H (HMM): builder-defined spec; the reference has no implementation at all
 *     (lib-discover/src/main.rs:1-3).
 *
 * All paths cited are relative to /root/reference.
 */
#ifndef CNV_ORACLE_H
#define CNV_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* bcf_float_missing as used by rust-htslib's f32::missing() */
#define ORC_F32_MISSING_BITS 0x7F800001u

/* ---- S1: lib-shared/src/stats.rs:163-305 ---------------------------------------------- */
double orc_stats_sum(const double *x, size_t n);
double orc_stats_min(const double *x, size_t n);
double orc_stats_max(const double *x, size_t n);
double orc_stats_mean(const double *x, size_t n);
double orc_stats_median(const double *x, size_t n);
double orc_stats_var(const double *x, size_t n);
double orc_stats_std_dev(const double *x, size_t n);
double orc_stats_std_dev_pct(const double *x, size_t n);
double orc_stats_median_abs_dev(const double *x, size_t n);
double orc_stats_median_abs_dev_pct(const double *x, size_t n);
double orc_stats_percentile(const double *x, size_t n, double pct);
void orc_stats_quartiles(const double *x, size_t n, double out[3]);
double orc_stats_iqr(const double *x, size_t n);

/* ---- BAM-level records as the reference sees them through rust-htslib ------------------ */
typedef struct {
    uint64_t n;
    const int32_t *pos;         /* record.pos()          */
    const int32_t *isize;       /* record.insert_size()  */
    const uint16_t *flag;       /* BAM FLAG              */
    const uint8_t *mapq;        /* record.mapq()         */
    const uint64_t *cigar_off;  /* n+1 offsets into cigar */
    const uint32_t *cigar;      /* BAM encoding: len<<4 | op (MIDNSHP=X -> 0..8) */
} orc_reads;

typedef struct {
    uint8_t min_mapq;           /* options.min_mapq            (cli.yaml:132-138, default 0)   */
    float min_unclipped;        /* options.min_unclipped       (cli.yaml:139-146, default 0.6) */
    uint32_t window_length;     /* options.window_length                                      */
    float min_window_remaining; /* options.min_window_remaining (default 0.5)                  */
    uint32_t min_raw_coverage;  /* options.min_raw_coverage     (default 10)                   */
} orc_cov_opts;

/* return codes: 0 ok; ORC_PANIC = the reference would panic (index out of range etc.) */
#define ORC_OK 0
#define ORC_PANIC 1

/* absolute end = pos + reference-consuming CIGAR ops (rust-htslib CigarStringView::end_pos) */
int32_t orc_cigar_end_pos(const orc_reads *r, uint64_t i);
/* htslib bam_endpos(): rlen==0 (or unmapped / no CIGAR) counts as 1 */
int32_t orc_bam_endpos(const orc_reads *r, uint64_t i);

/* A2: lib-coverage/src/agg.rs:115-198 (put_bam_record + skip_*), :223-229 */
int orc_frag_gw(const orc_reads *r, const orc_cov_opts *o, uint64_t region_end,
                const uint32_t *pile_start, const uint32_t *pile_end, uint32_t n_pile,
                uint32_t *counters /* ceil(region_end/wl) */, uint32_t *num_processed,
                uint32_t *num_skipped);
/* A2 get_stats: agg.rs:201-219, 231-244 */
void orc_frag_gw_stats(const uint32_t *counters, uint64_t region_end, uint32_t wl,
                       const uint32_t *pile_start, const uint32_t *pile_end, uint32_t n_pile,
                       float *cov, uint32_t *start, uint32_t *end, float *frac_removed);

/* A3: agg.rs:301-361.  Tie rule among equally distant overlapping targets: lowest target index
 * (the reference's order is bio::IntervalTree iteration order -- unpinned). */
int orc_frag_targets(const orc_reads *r, const orc_cov_opts *o, const uint32_t *tstart,
                     const uint32_t *tend, uint32_t n_targets, uint32_t *counters,
                     uint32_t *num_processed, uint32_t *num_skipped);

/* A4: agg.rs:485-574.  plp_mask = flags that keep a read out of htslib's pileup (columns);
 * SURVEY Appendix A.1 documents 0x704 (UNMAP|SECONDARY|QCFAIL|DUP).  Per-base arrays are
 * returned for debugging if non-NULL (length num_bins*wl). */
int orc_coverage(const orc_reads *r, const orc_cov_opts *o, uint64_t region_end,
                 uint32_t plp_mask, float *cov_mean, float *cov_sd, uint32_t *dbg_cols,
                 uint32_t *dbg_depth);

/* A5: lib-coverage/src/reference.rs:84-130 */
void orc_refstats(const uint8_t *seq, uint64_t len, uint32_t wl, float *gc, uint8_t *has_gap);

/* A6: lib-coverage/src/lib.rs:623-668.  cov_sd / frac_removed may be NULL (Option::None).
 * ft bit0 = LOWCOV, bit1 = PILE. */
void orc_cov_records(uint64_t n, const float *cov, const float *cov_sd, const uint32_t *start,
                     const uint32_t *end, const float *frac_removed, int is_targets,
                     const orc_cov_opts *o, float *lcv, float *lcvsd, uint8_t *ft);

/* N1: lib-normalize/src/uncorrected.rs:20-96 */
double orc_norm_total(const float *lcv, const float *lcvsd, uint64_t n, float *cv, float *cvsd);

/* N2: lib-normalize/src/correct_binned.rs:18-146.  CKMS(eps=1e-6).query(0.5) restated as the
 * exact order statistic of 1-based rank ceil(n/2) (quantiles crate, no compression below 5e5
 * items per GC bin).  flags: bit0 CV written, bit1 CV2 written, bit2 CVSD written,
 * bit3 FEW_GCWINDOWS.  medians: 101 entries (missing pattern when count < 10). */
#define ORC_N2_CV 1
#define ORC_N2_CV2 2
#define ORC_N2_CVSD 4
#define ORC_N2_FEW_GC 8
void orc_norm_gc_median(const float *lcv, const float *lcvsd, const float *gc, uint64_t n,
                        float *cv, float *cv2, float *cvsd, uint8_t *flags, float *medians);

/* W2: lib-model-wis/src/lib.rs:128-197.  Selection order = (distance, index) lexicographic
 * (pdqselect residual order is unpinned).  Rows [row_begin,row_end) only, nthreads >= 1. */
void orc_wis_distances(const float *X, const uint32_t *tids, uint32_t T, uint32_t S,
                       uint32_t max_ref_targets, uint32_t row_begin, uint32_t row_end,
                       float *out_dist, uint32_t *out_idx, int nthreads);
/* W3: lib.rs:200-232.  Returns threshold; ref lists compacted in place to ref_len[t] entries */
float orc_wis_prune(const float *dist, uint32_t *idx, uint32_t T, uint32_t k, double z,
                    uint32_t *ref_len);
/* W4: lib.rs:235-271 */
void orc_wis_calls(const float *X, uint32_t T, uint32_t S, const uint32_t *ref_idx,
                   const uint32_t *ref_len, uint32_t k, uint32_t min_ref_targets, double z,
                   double rel, uint32_t *num_calls);
/* W5: lib.rs:362-377.  bit0 FEW_REF, bit1 UNRELIABLE */
void orc_wis_filters(const uint32_t *ref_len, const uint32_t *num_calls, uint32_t T,
                     uint32_t min_ref_targets, uint32_t max_samples_reliable, uint8_t *filt);

/* M1: lib-mod-cov/src/lib.rs:76-164, 209-240.  status bit0 = has probe info (else NO_REF),
 * bit1 = CV2 written. */
void orc_modcov(const float *cv, uint32_t T, const uint8_t *model_filt, const uint32_t *ref_idx,
                const uint32_t *ref_len, uint32_t k, float *ref_mean, float *ref_sd,
                float *cv_rel, float *cvz, float *cv2, uint8_t *status);

/* pile threshold: lib-coverage/src/lib.rs:265-351 */
int orc_pile_threshold(const uint64_t *y, size_t n, double fdr, uint32_t *out);
/* piles_from_depths: lib-coverage/src/piles.rs:78-127; returns number of piles */
uint64_t orc_piles_from_depths(const uint32_t *depths, uint64_t n, uint32_t pile_max_gap,
                               uint32_t *pstart, uint32_t *pend, uint32_t *psize);

/* ---- H: builder-defined HMM (DESIGN.md "HMM spec"); no reference implementation -------- */
#define ORC_HMM_STATES 5
typedef struct {
    double log_start[ORC_HMM_STATES];
    double log_trans[ORC_HMM_STATES][ORC_HMM_STATES]; /* [from][to] */
    double mu[ORC_HMM_STATES];
    double inv_sd[ORC_HMM_STATES];
    double log_norm[ORC_HMM_STATES]; /* -log(sd) - 0.5*log(2*pi) */
    /* Viterbi works on integer scores in units of 2^-16 nat (exactly associative, so any
     * evaluation order gives the same path): f32 emission parameters and integer transitions */
    float mu_f[ORC_HMM_STATES];
    float inv_sd_f[ORC_HMM_STATES];    /* (float)(inv_sd * sqrt(65536 / 2)) */
    float log_norm_sc[ORC_HMM_STATES]; /* (float)(log_norm * 65536) */
    int32_t sc_start, sc_stay, sc_move;  /* llrint(log(.) * 65536) */
} orc_hmm_tables;
#define ORC_HMM_SCALE 65536.0
#define ORC_HMM_ZSCALE 181.01933598375618 /* sqrt(32768) */
#define ORC_HMM_FLOOR 4194304.0f /* emission scores are floored at -64 nat */
/* integer emission scores of one observation (0 for a missing value) */
void orc_hmm_emit_scores(float cv, const orc_hmm_tables *t, int32_t *e);
void orc_hmm_make_tables(double sigma_s, double sigma0, double eps_cn, double p_move,
                         orc_hmm_tables *t);
void orc_hmm_viterbi(const float *cv, uint64_t n, const orc_hmm_tables *t, uint8_t *state);
void orc_hmm_viterbi_f64(const float *cv, uint64_t n, const orc_hmm_tables *t, uint8_t *state);
void orc_hmm_posterior(const float *cv, uint64_t n, const orc_hmm_tables *t, double *post);

#ifdef __cplusplus
}
#endif
#endif
