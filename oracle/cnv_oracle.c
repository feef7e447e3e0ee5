/*
 * cnv_oracle.c -- CPU restatement ("oracle") of the reference's read-depth hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see cnv_oracle.h).  Plain C, compiled -O2 -ffp-contract=off and
 * never with -ffast-math, so that every float operation rounds exactly as the Rust reference's
 * (Rust/LLVM never contracts a*b+c into an FMA and never re-associates).
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 */
#include "cnv_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

/* ======================================================================================== */
/* S1  lib-shared/src/stats.rs                                                              */
/* ======================================================================================== */

/* stats.rs:17-30 local_cmp: NaN sorts above everything */
static int local_cmp(const void *pa, const void *pb) {
    double x = *(const double *)pa, y = *(const double *)pb;
    if (isnan(y)) return -1;
    if (isnan(x)) return 1;
    if (x < y) return -1;
    if (x == y) return 0;
    return 1;
}

static double *sorted_copy(const double *x, size_t n) {
    double *tmp = (double *)malloc((n ? n : 1) * sizeof(double));
    memcpy(tmp, x, n * sizeof(double));
    qsort(tmp, n, sizeof(double), local_cmp); /* stats.rs:32-34 */
    return tmp;
}

/* stats.rs:165-197: Shewchuk grow-expansion; the partials are folded left to right from 0.0 */
double orc_stats_sum(const double *xs, size_t n) {
    size_t cap = 16, len = 0;
    double *partials = (double *)malloc(cap * sizeof(double));
    for (size_t k = 0; k < n; ++k) {
        double x = xs[k];
        size_t j = 0;
        for (size_t i = 0; i < len; ++i) {
            double y = partials[i];
            if (fabs(x) < fabs(y)) {
                double t = x;
                x = y;
                y = t;
            }
            double hi = x + y;
            double lo = y - (hi - x);
            if (lo != 0.0) {
                partials[j] = lo;
                j += 1;
            }
            x = hi;
        }
        if (j >= len) {
            if (len == cap) {
                cap *= 2;
                partials = (double *)realloc(partials, cap * sizeof(double));
            }
            partials[len++] = x;
        } else {
            partials[j] = x;
            len = j + 1;
        }
    }
    double acc = 0.0;
    for (size_t i = 0; i < len; ++i) acc = acc + partials[i];
    free(partials);
    return acc;
}

/* stats.rs:199-207: fold with f64::min / f64::max (NaN-ignoring) */
static double rust_fmin(double p, double q) {
    if (isnan(p)) return q;
    if (isnan(q)) return p;
    return p < q ? p : q;
}
static double rust_fmax(double p, double q) {
    if (isnan(p)) return q;
    if (isnan(q)) return p;
    return p > q ? p : q;
}
double orc_stats_min(const double *x, size_t n) {
    double p = x[0];
    for (size_t i = 0; i < n; ++i) p = rust_fmin(p, x[i]);
    return p;
}
double orc_stats_max(const double *x, size_t n) {
    double p = x[0];
    for (size_t i = 0; i < n; ++i) p = rust_fmax(p, x[i]);
    return p;
}

/* stats.rs:209-212 */
double orc_stats_mean(const double *x, size_t n) { return orc_stats_sum(x, n) / (double)n; }

/* stats.rs:285-305 */
static double percentile_of_sorted(const double *s, size_t n, double pct) {
    if (n == 1) return s[0];
    if (pct == 100.0) return s[n - 1];
    double length = (double)(n - 1);
    double rank = (pct / 100.0) * length;
    double lrank = floor(rank);
    double d = rank - lrank;
    size_t k = (size_t)lrank;
    double lo = s[k];
    double hi = s[k + 1];
    return lo + (hi - lo) * d;
}

/* stats.rs:259-263 */
double orc_stats_percentile(const double *x, size_t n, double pct) {
    double *tmp = sorted_copy(x, n);
    double r = percentile_of_sorted(tmp, n, pct);
    free(tmp);
    return r;
}
/* stats.rs:214-216 */
double orc_stats_median(const double *x, size_t n) { return orc_stats_percentile(x, n, 50.0); }

/* stats.rs:218-234: naive left-to-right sum of squared deviations, n-1 denominator */
double orc_stats_var(const double *x, size_t n) {
    if (n < 2) return 0.0;
    double mean = orc_stats_mean(x, n);
    double v = 0.0;
    for (size_t i = 0; i < n; ++i) {
        double d = x[i] - mean;
        v = v + d * d;
    }
    double denom = (double)(n - 1);
    return v / denom;
}
/* stats.rs:236-238 */
double orc_stats_std_dev(const double *x, size_t n) { return sqrt(orc_stats_var(x, n)); }
/* stats.rs:240-243 */
double orc_stats_std_dev_pct(const double *x, size_t n) {
    return (orc_stats_std_dev(x, n) / orc_stats_mean(x, n)) * 100.0;
}
/* stats.rs:245-252 */
double orc_stats_median_abs_dev(const double *x, size_t n) {
    double med = orc_stats_median(x, n);
    double *ad = (double *)malloc((n ? n : 1) * sizeof(double));
    for (size_t i = 0; i < n; ++i) ad[i] = fabs(med - x[i]);
    double r = orc_stats_median(ad, n) * 1.4826;
    free(ad);
    return r;
}
/* stats.rs:254-257 */
double orc_stats_median_abs_dev_pct(const double *x, size_t n) {
    return (orc_stats_median_abs_dev(x, n) / orc_stats_median(x, n)) * 100.0;
}
/* stats.rs:265-275 */
void orc_stats_quartiles(const double *x, size_t n, double out[3]) {
    double *tmp = sorted_copy(x, n);
    out[0] = percentile_of_sorted(tmp, n, 25.0);
    out[1] = percentile_of_sorted(tmp, n, 50.0);
    out[2] = percentile_of_sorted(tmp, n, 75.0);
    free(tmp);
}
/* stats.rs:277-280 */
double orc_stats_iqr(const double *x, size_t n) {
    double q[3];
    orc_stats_quartiles(x, n, q);
    return q[2] - q[0];
}

/* ======================================================================================== */
/* BAM record helpers                                                                       */
/* ======================================================================================== */

enum { CIG_M = 0, CIG_I, CIG_D, CIG_N, CIG_S, CIG_H, CIG_P, CIG_EQ, CIG_X };

static float f32_missing(void) {
    uint32_t b = ORC_F32_MISSING_BITS;
    float f;
    memcpy(&f, &b, 4);
    return f;
}

static uint32_t cigar_ref_len(const orc_reads *r, uint64_t i) {
    uint32_t rlen = 0;
    for (uint64_t c = r->cigar_off[i]; c < r->cigar_off[i + 1]; ++c) {
        uint32_t op = r->cigar[c] & 0xF, len = r->cigar[c] >> 4;
        if (op == CIG_M || op == CIG_D || op == CIG_N || op == CIG_EQ || op == CIG_X) rlen += len;
    }
    return rlen;
}

int32_t orc_cigar_end_pos(const orc_reads *r, uint64_t i) {
    return r->pos[i] + (int32_t)cigar_ref_len(r, i);
}

int32_t orc_bam_endpos(const orc_reads *r, uint64_t i) {
    uint32_t rlen = 1;
    if (!(r->flag[i] & 0x4) && r->cigar_off[i + 1] > r->cigar_off[i]) {
        rlen = cigar_ref_len(r, i);
        if (rlen == 0) rlen = 1;
    }
    return r->pos[i] + (int32_t)rlen;
}

/* agg.rs:152-154 / 363-365 */
static int skip_mapq(const orc_reads *r, uint64_t i, const orc_cov_opts *o) {
    return r->mapq[i] < o->min_mapq;
}
/* agg.rs:157-162 / 368-373 */
static int skip_flags(const orc_reads *r, uint64_t i) {
    uint16_t f = r->flag[i];
    return (f & 0x100) || (f & 0x800) || (f & 0x400) || (f & 0x200);
}
/* agg.rs:165-171 / 376-382 */
static int skip_discordant(const orc_reads *r, uint64_t i) {
    uint16_t f = r->flag[i];
    if (!(f & 0x1)) return 0;
    return !(f & 0x2);
}
/* agg.rs:174-193 / 385-404: f32 ratio; an empty CIGAR gives NaN < x == false */
static int skip_clipping(const orc_reads *r, uint64_t i, const orc_cov_opts *o) {
    uint32_t num_clipped = 0, num_unclipped = 0;
    for (uint64_t c = r->cigar_off[i]; c < r->cigar_off[i + 1]; ++c) {
        uint32_t op = r->cigar[c] & 0xF, len = r->cigar[c] >> 4;
        if (op == CIG_M || op == CIG_I || op == CIG_EQ || op == CIG_X)
            num_unclipped += len;
        else if (op == CIG_S || op == CIG_H)
            num_clipped += len;
    }
    float ratio = (float)num_unclipped / (float)(num_clipped + num_unclipped);
    return ratio < o->min_unclipped;
}
/* agg.rs:196-198 / 407-409 */
static int skip_paired_and_all_but_leftmost(const orc_reads *r, uint64_t i) {
    return (r->flag[i] & 0x1) && (r->isize[i] <= 0);
}

static int passes_fragment_filters(const orc_reads *r, uint64_t i, const orc_cov_opts *o) {
    return !skip_mapq(r, i, o) && !skip_flags(r, i) && !skip_discordant(r, i) &&
           !skip_clipping(r, i, o) && !skip_paired_and_all_but_leftmost(r, i);
}

/* bio IntervalTree::find(q).next().is_some() over the (disjoint or not) pile intervals */
static int piles_hit(const uint32_t *ps, const uint32_t *pe, uint32_t n, uint32_t qs, uint32_t qe) {
    for (uint32_t k = 0; k < n; ++k)
        if (ps[k] < qe && qs < pe[k]) return 1;
    return 0;
}

/* ======================================================================================== */
/* A2  FragmentsGenomeWideAggregator  (lib-coverage/src/agg.rs:87-257)                      */
/* ======================================================================================== */

int orc_frag_gw(const orc_reads *r, const orc_cov_opts *o, uint64_t region_end,
                const uint32_t *pile_start, const uint32_t *pile_end, uint32_t n_pile,
                uint32_t *counters, uint32_t *num_processed, uint32_t *num_skipped) {
    const uint64_t wl = o->window_length;
    const uint64_t num_bins = (region_end + wl - 1) / wl; /* agg.rs:97 */
    memset(counters, 0, num_bins * sizeof(uint32_t));
    *num_processed = 0;
    *num_skipped = 0;
    for (uint64_t i = 0; i < r->n; ++i) { /* agg.rs:223-229 */
        if (!passes_fragment_filters(r, i, o)) continue; /* agg.rs:116-120 */
        *num_processed += 1;
        /* agg.rs:124-133: i32 arithmetic, `/` truncates toward zero, then `as u32` */
        int32_t c32;
        if (r->flag[i] & 0x1)
            c32 = r->pos[i] + r->isize[i] / 2;
        else
            c32 = r->pos[i] + orc_cigar_end_pos(r, i); /* sic: pos + absolute end */
        uint32_t fragment_center = (uint32_t)c32;
        uint64_t bin = (uint64_t)fragment_center / wl; /* agg.rs:141 */
        if (n_pile == 0 ||
            !piles_hit(pile_start, pile_end, n_pile, fragment_center, fragment_center + 1)) {
            if (bin >= num_bins) return ORC_PANIC; /* mapq_sums[bin] out of range, agg.rs:144 */
            counters[bin] += 1;
        } else {
            *num_skipped += 1; /* agg.rs:146 */
        }
    }
    return ORC_OK;
}

void orc_frag_gw_stats(const uint32_t *counters, uint64_t region_end, uint32_t wl,
                       const uint32_t *pile_start, const uint32_t *pile_end, uint32_t n_pile,
                       float *cov, uint32_t *start, uint32_t *end, float *frac_removed) {
    const uint64_t num_bins = (region_end + wl - 1) / wl; /* agg.rs:246-248 */
    for (uint64_t w = 0; w < num_bins; ++w) {
        /* agg.rs:201-219 get_masked_count: window NOT clipped to the contig end */
        uint32_t ws = (uint32_t)(w * wl), we = (uint32_t)((w + 1) * wl), masked = 0;
        for (uint32_t k = 0; k < n_pile; ++k) {
            if (pile_start[k] < we && ws < pile_end[k]) {
                int32_t e = (int32_t)(we < pile_end[k] ? we : pile_end[k]);
                int32_t s = (int32_t)(ws > pile_start[k] ? ws : pile_start[k]);
                masked += (uint32_t)(e - s);
            }
        }
        cov[w] = (float)counters[w];                                   /* agg.rs:236 */
        start[w] = ws;                                                 /* agg.rs:239 */
        end[w] = (uint32_t)(region_end < (w + 1) * wl ? region_end : (w + 1) * wl); /* :240 */
        frac_removed[w] = (float)masked / (float)wl;                   /* agg.rs:241 */
    }
}

/* ======================================================================================== */
/* A3  FragmentsTargetRegionsAggregator  (agg.rs:261-444)                                   */
/* ======================================================================================== */

/* agg.rs:350-361 */
static uint32_t compute_dist(uint32_t start, uint32_t end, uint32_t pt) {
    if (pt < start) return start - pt;
    if (pt >= end) return pt - end + 1;
    return 0;
}

int orc_frag_targets(const orc_reads *r, const orc_cov_opts *o, const uint32_t *tstart,
                     const uint32_t *tend, uint32_t n_targets, uint32_t *counters,
                     uint32_t *num_processed, uint32_t *num_skipped) {
    memset(counters, 0, (size_t)n_targets * sizeof(uint32_t));
    *num_processed = 0;
    *num_skipped = 0;
    for (uint64_t i = 0; i < r->n; ++i) { /* agg.rs:413-419 */
        if (!passes_fragment_filters(r, i, o)) continue; /* agg.rs:302-306 */
        *num_processed += 1;
        uint32_t begin_pos = (uint32_t)r->pos[i]; /* agg.rs:311 */
        uint32_t end_pos = (r->flag[i] & 0x1) ? (uint32_t)(r->pos[i] + r->isize[i])
                                              : (uint32_t)orc_cigar_end_pos(r, i); /* :312-320 */
        uint32_t fragment_center = (begin_pos + end_pos) / 2; /* agg.rs:321 */
        /* agg.rs:324-339: all intervals intersecting [begin,end), min distance, strict `<`.
         * Iteration here is ascending target index => lowest index wins ties (documented). */
        int have = 0;
        uint32_t best_idx = 0, best_dist = 0;
        for (uint32_t t = 0; t < n_targets; ++t) {
            if (!(tstart[t] < end_pos && begin_pos < tend[t])) continue;
            uint32_t d = compute_dist(tstart[t], tend[t], fragment_center);
            if (!have || d < best_dist) {
                have = 1;
                best_idx = t;
                best_dist = d;
            }
        }
        if (have)
            counters[best_idx] += 1; /* agg.rs:340-342 */
        else
            *num_skipped += 1; /* agg.rs:344 */
    }
    return ORC_OK;
}

/* ======================================================================================== */
/* A4  CoverageAggregator  (agg.rs:468-600)                                                 */
/* ======================================================================================== */

typedef struct {
    uint32_t wl;
    uint64_t num_bins;
    double *depths; /* self.depths (usize -> f64 at push time, agg.rs:504) */
    float *cov_mean, *cov_sd;
} cov_state;

/* agg.rs:503-514 */
static int push_window(cov_state *st, uint64_t window_id) {
    if (window_id >= st->num_bins) return ORC_PANIC; /* self.coverage[window_id] */
    st->cov_mean[window_id] = (float)orc_stats_mean(st->depths, st->wl);
    st->cov_sd[window_id] = (float)orc_stats_std_dev(st->depths, st->wl);
    for (uint32_t j = 0; j < st->wl; ++j) st->depths[j] = 0.0;
    return ORC_OK;
}

int orc_coverage(const orc_reads *r, const orc_cov_opts *o, uint64_t region_end,
                 uint32_t plp_mask, float *cov_mean, float *cov_sd, uint32_t *dbg_cols,
                 uint32_t *dbg_depth) {
    const uint64_t wl = o->window_length;
    const uint64_t num_bins = (region_end + wl - 1) / wl; /* agg.rs:487 */
    const uint64_t L = num_bins * wl;
    int rc = ORC_OK;
    for (uint64_t w = 0; w < num_bins; ++w) cov_mean[w] = 0.0f, cov_sd[w] = 0.0f; /* :456-463 */

    /* htslib pileup engine stand-in: per reference position, the number of alignments in the
     * column (reads not excluded by plp_mask; deletions / ref-skips count as present) and the
     * number of those passing the reference's own filter (agg.rs:548-557). */
    int32_t *dc = (int32_t *)calloc(L + 1, sizeof(int32_t));
    int32_t *dd = (int32_t *)calloc(L + 1, sizeof(int32_t));
    for (uint64_t i = 0; i < r->n; ++i) {
        uint16_t f = r->flag[i];
        if (f & plp_mask) continue;
        int64_t s = r->pos[i], e = orc_bam_endpos(r, i);
        if (s < 0) continue;
        if ((uint64_t)e > L) {
            rc = ORC_PANIC; /* a column past the last window: coverage[window_id] out of range */
            goto done;
        }
        dc[s] += 1;
        dc[e] -= 1;
        int ok = !(f & 0x100) && !(f & 0x400) && !(f & 0x800) && !(f & 0x200) &&
                 (r->mapq[i] >= o->min_mapq) && (!(f & 0x1) || (f & 0x2)); /* agg.rs:550-556 */
        if (ok) {
            dd[s] += 1;
            dd[e] -= 1;
        }
    }

    {
        cov_state st;
        st.wl = (uint32_t)wl;
        st.num_bins = num_bins;
        st.depths = (double *)calloc(wl, sizeof(double));
        st.cov_mean = cov_mean;
        st.cov_sd = cov_sd;
        int have_window = 0, have_prev = 0; /* self.window_id / prev_window_id : Option */
        uint64_t window_id = 0, prev_window_id = 0;
        int64_t cols = 0, depth = 0;
        for (uint64_t pos = 0; pos < L && rc == ORC_OK; ++pos) { /* agg.rs:524-562 */
            cols += dc[pos];
            depth += dd[pos];
            if (dbg_cols) dbg_cols[pos] = (uint32_t)cols;
            if (dbg_depth) dbg_depth[pos] = (uint32_t)depth;
            if (cols <= 0) continue; /* pileup only yields covered positions */
            /* agg.rs:529-542 */
            if (have_window && have_prev) {
                if (window_id != prev_window_id) {
                    rc = push_window(&st, prev_window_id);
                    prev_window_id = window_id;
                }
                window_id = pos / wl;
            } else if (have_window && !have_prev) {
                prev_window_id = window_id;
                have_prev = 1;
                window_id = pos / wl;
            } else {
                window_id = pos / wl;
                have_window = 1;
            }
            st.depths[pos % wl] = (double)depth; /* agg.rs:560 */
        }
        if (rc == ORC_OK && have_window) { /* agg.rs:564-573 */
            if (have_prev) {
                if (prev_window_id != window_id) rc = push_window(&st, window_id);
            } else {
                rc = push_window(&st, window_id);
            }
        }
        free(st.depths);
    }
done:
    free(dc);
    free(dd);
    return rc;
}

/* ======================================================================================== */
/* A5  ReferenceStats::look_at_chars  (lib-coverage/src/reference.rs:84-130)                */
/* ======================================================================================== */

void orc_refstats(const uint8_t *seq, uint64_t len, uint32_t wl, float *gc, uint8_t *has_gap) {
    const uint64_t num_buckets = (len + wl - 1) / wl; /* reference.rs:91 */
    int32_t *gc_count = (int32_t *)calloc(num_buckets ? num_buckets : 1, sizeof(int32_t));
    int32_t *non_n_count = (int32_t *)calloc(num_buckets ? num_buckets : 1, sizeof(int32_t));
    memset(has_gap, 0, num_buckets);
    for (uint64_t i = 0; i < len; ++i) { /* reference.rs:100-116 */
        uint64_t bucket = i / wl;
        switch (seq[i]) {
        case 'g': case 'c': case 'G': case 'C':
            gc_count[bucket] += 1;
            non_n_count[bucket] += 1;
            break;
        case 'a': case 't': case 'A': case 'T':
            non_n_count[bucket] += 1;
            break;
        case 'n': case 'N':
            has_gap[bucket] = 1;
            break;
        default:
            break;
        }
    }
    for (uint64_t b = 0; b < num_buckets; ++b) { /* reference.rs:120-127 */
        if (non_n_count[b] == 0)
            gc[b] = f32_missing();
        else
            gc[b] = (float)gc_count[b] / (float)non_n_count[b];
    }
    free(gc_count);
    free(non_n_count);
}

/* ======================================================================================== */
/* A6  record assembly  (lib-coverage/src/lib.rs:623-668)                                   */
/* ======================================================================================== */

void orc_cov_records(uint64_t n, const float *cov, const float *cov_sd, const uint32_t *start,
                     const uint32_t *end, const float *frac_removed, int is_targets,
                     const orc_cov_opts *o, float *lcv, float *lcvsd, uint8_t *ft) {
    for (uint64_t i = 0; i < n; ++i) {
        uint8_t f = 0;
        if (is_targets && cov[i] < (float)o->min_raw_coverage) f |= 1; /* lib.rs:627-631 */
        float length = (float)(end[i] - start[i]);                     /* lib.rs:637 */
        lcv[i] = cov[i] / length;                                      /* lib.rs:639 */
        if (cov_sd && lcvsd) lcvsd[i] = cov_sd[i] / length;            /* lib.rs:641-645 */
        if (frac_removed && frac_removed[i] > o->min_window_remaining) f |= 2; /* :647-654 */
        ft[i] = f;
    }
}

/* ======================================================================================== */
/* N1  TotalCovSum  (lib-normalize/src/uncorrected.rs:20-96)                                */
/* ======================================================================================== */

double orc_norm_total(const float *lcv, const float *lcvsd, uint64_t n, float *cv, float *cvsd) {
    double *vals = (double *)malloc((n ? n : 1) * sizeof(double));
    for (uint64_t i = 0; i < n; ++i) vals[i] = (double)lcv[i]; /* uncorrected.rs:39-46 */
    double lcv_sum = orc_stats_sum(vals, n);                    /* uncorrected.rs:48 */
    free(vals);
    for (uint64_t i = 0; i < n; ++i) {
        cv[i] = (float)((double)lcv[i] / lcv_sum); /* uncorrected.rs:65-72 */
        if (lcvsd && cvsd) cvsd[i] = (float)((double)lcvsd[i] / lcv_sum); /* :75-81 */
    }
    return lcv_sum;
}

/* ======================================================================================== */
/* N2  MedianGcBinned  (lib-normalize/src/correct_binned.rs:18-146)                         */
/* ======================================================================================== */

static int f32_cmp(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

void orc_norm_gc_median(const float *lcv, const float *lcvsd, const float *gc, uint64_t n,
                        float *cv, float *cv2, float *cvsd, uint8_t *flags, float *medians) {
    enum { NB = 101, MIN_COUNT = 10 }; /* correct_binned.rs:20,36 */
    uint64_t cnt[NB];
    memset(cnt, 0, sizeof(cnt));
    for (uint64_t i = 0; i < n; ++i) { /* correct_binned.rs:47-63 */
        float g = gc[i];
        if (isfinite(g)) {
            float kf = floorf(g * 100.0f);
            if (kf >= 0.0f && kf < (float)NB) cnt[(int)kf] += 1; /* else: processors[gc] panics */
        }
    }
    float *buf[NB];
    uint64_t fill[NB];
    for (int b = 0; b < NB; ++b) {
        buf[b] = (float *)malloc((cnt[b] ? cnt[b] : 1) * sizeof(float));
        fill[b] = 0;
    }
    for (uint64_t i = 0; i < n; ++i) {
        float g = gc[i];
        if (isfinite(g)) {
            float kf = floorf(g * 100.0f);
            if (kf >= 0.0f && kf < (float)NB) buf[(int)kf][fill[(int)kf]++] = lcv[i];
        }
    }
    for (int b = 0; b < NB; ++b) { /* correct_binned.rs:65-74 */
        if (cnt[b] < MIN_COUNT) {
            medians[b] = f32_missing();
        } else {
            /* CKMS(1e-6).query(0.5): element of 1-based rank ceil(n/2) (no compression) */
            qsort(buf[b], cnt[b], sizeof(float), f32_cmp);
            medians[b] = buf[b][(cnt[b] + 1) / 2 - 1];
        }
        free(buf[b]);
    }
    for (uint64_t i = 0; i < n; ++i) { /* correct_binned.rs:82-142 */
        uint8_t f = 0;
        float g = gc[i];
        cv[i] = cv2[i] = f32_missing();
        if (cvsd) cvsd[i] = f32_missing();
        if (isfinite(g)) {
            float kf = floorf(g * 100.0f);
            if (kf >= 0.0f && kf < (float)NB) {
                float median = medians[(int)kf];
                if (!isfinite(median)) {
                    f |= ORC_N2_FEW_GC; /* :99-102 */
                } else {
                    float norm_lcv = lcv[i] / median; /* :104 */
                    cv[i] = norm_lcv;
                    f |= ORC_N2_CV;
                    if (isfinite(norm_lcv)) { /* :108-117 */
                        cv2[i] = (norm_lcv == 0.0f) ? f32_missing() : log2f(norm_lcv);
                        f |= ORC_N2_CV2;
                    }
                    if (lcvsd && cvsd) { /* :119-126 */
                        cvsd[i] = lcvsd[i] / median;
                        f |= ORC_N2_CVSD;
                    }
                }
            }
        }
        flags[i] = f;
    }
}

/* ======================================================================================== */
/* W2  compute_distances  (lib-model-wis/src/lib.rs:128-197)                                */
/* ======================================================================================== */

typedef struct {
    float d;
    uint32_t j;
} dist_ent;

/* (distance, index) lexicographic "less" */
static inline int ent_less(dist_ent a, dist_ent b) {
    return (a.d < b.d) || (a.d == b.d && a.j < b.j);
}

static void heap_sift_down(dist_ent *h, uint32_t n, uint32_t i) { /* max-heap on ent_less */
    for (;;) {
        uint32_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && ent_less(h[m], h[l])) m = l;
        if (r < n && ent_less(h[m], h[r])) m = r;
        if (m == i) return;
        dist_ent t = h[i];
        h[i] = h[m];
        h[m] = t;
        i = m;
    }
}

static int ent_cmp(const void *a, const void *b) {
    dist_ent x = *(const dist_ent *)a, y = *(const dist_ent *)b;
    if (ent_less(x, y)) return -1;
    if (ent_less(y, x)) return 1;
    return 0;
}

typedef struct {
    const float *X;
    const uint32_t *tids;
    uint32_t T, S, size, row_begin, row_end;
    float *out_dist;
    uint32_t *out_idx;
    volatile uint32_t *next_row; /* shared work counter (rows handed out in blocks of 8) */
} wis_job;

static void wis_one_row(const wis_job *J, uint32_t i, dist_ent *heap) {
    const uint32_t T = J->T, S = J->S, size = J->size;
    const float *row = J->X + (size_t)i * S;
    uint32_t hn = 0;
    for (uint32_t j = 0; j < T; ++j) { /* lib.rs:153-165 */
        dist_ent e;
        e.j = j;
        if (J->tids[i] == J->tids[j]) {
            e.d = INFINITY;
        } else {
            const float *row2 = J->X + (size_t)j * S;
            float y = 0.0f;
            for (uint32_t s = 0; s < S; ++s) {
                float x = row[s] - row2[s];
                y += x * x; /* separate f32 mul and add; never an FMA */
            }
            e.d = sqrtf(y);
        }
        /* lib.rs:167-171 select the `size` smallest, then sort them */
        if (hn < size) {
            heap[hn++] = e;
            if (hn == size)
                for (int32_t q = (int32_t)size / 2 - 1; q >= 0; --q)
                    heap_sift_down(heap, size, (uint32_t)q);
        } else if (ent_less(e, heap[0])) {
            heap[0] = e;
            heap_sift_down(heap, size, 0);
        }
    }
    qsort(heap, size, sizeof(dist_ent), ent_cmp);
    for (uint32_t q = 0; q < size; ++q) {
        J->out_dist[(size_t)(i - J->row_begin) * size + q] = heap[q].d;
        J->out_idx[(size_t)(i - J->row_begin) * size + q] = heap[q].j;
    }
}

static void *wis_worker(void *arg) { /* stands in for rayon's par_iter over rows, lib.rs:144 */
    const wis_job *J = (const wis_job *)arg;
    dist_ent *heap = (dist_ent *)malloc((size_t)J->size * sizeof(dist_ent));
    for (;;) {
        uint32_t b = __atomic_fetch_add(J->next_row, 8u, __ATOMIC_RELAXED);
        if (b >= J->row_end) break;
        uint32_t e = b + 8u < J->row_end ? b + 8u : J->row_end;
        for (uint32_t i = b; i < e; ++i) wis_one_row(J, i, heap);
    }
    free(heap);
    return NULL;
}

void orc_wis_distances(const float *X, const uint32_t *tids, uint32_t T, uint32_t S,
                       uint32_t max_ref_targets, uint32_t row_begin, uint32_t row_end,
                       float *out_dist, uint32_t *out_idx, int nthreads) {
    const uint32_t size = T < max_ref_targets ? T : max_ref_targets; /* lib.rs:167 */
    if (size == 0 || row_end <= row_begin) return;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    volatile uint32_t next_row = row_begin;
    wis_job J = {X, tids, T, S, size, row_begin, row_end, out_dist, out_idx, &next_row};
    pthread_t th[256];
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, wis_worker, &J);
    wis_worker(&J);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
}

/* ======================================================================================== */
/* W3  prune_distances  (lib-model-wis/src/lib.rs:200-232)                                  */
/* ======================================================================================== */

float orc_wis_prune(const float *dist, uint32_t *idx, uint32_t T, uint32_t k, double z,
                    uint32_t *ref_len) {
    if (T < 2) { /* lib.rs:211-213: returns an empty Vec */
        for (uint32_t t = 0; t < T; ++t) ref_len[t] = 0;
        return NAN;
    }
    double *err = (double *)malloc((size_t)T * sizeof(double));
    for (uint32_t t = 0; t < T; ++t) err[t] = (double)dist[(size_t)t * k]; /* lib.rs:207-210 */
    double mean = orc_stats_mean(err, T);
    double sd = orc_stats_std_dev(err, T);
    free(err);
    float threshold = (float)(mean + z * sd); /* lib.rs:216 */
    for (uint32_t t = 0; t < T; ++t) {        /* lib.rs:221-228 */
        uint32_t m = 0;
        for (uint32_t q = 0; q < k; ++q)
            if (dist[(size_t)t * k + q] <= threshold) idx[(size_t)t * k + m++] = idx[(size_t)t * k + q];
        ref_len[t] = m;
    }
    return threshold;
}

/* ======================================================================================== */
/* W4  count_per_probe_calls  (lib-model-wis/src/lib.rs:235-271)                            */
/* ======================================================================================== */

void orc_wis_calls(const float *X, uint32_t T, uint32_t S, const uint32_t *ref_idx,
                   const uint32_t *ref_len, uint32_t k, uint32_t min_ref_targets, double z,
                   double rel_thr, uint32_t *num_calls) {
    double *ref_dist = (double *)malloc((size_t)(k ? k : 1) * sizeof(double));
    for (uint32_t t = 0; t < T; ++t) {
        num_calls[t] = 0;
        const uint32_t m = ref_len[t];
        for (uint32_t s = 0; s < S; ++s) {
            for (uint32_t q = 0; q < m; ++q) /* lib.rs:250-253 */
                ref_dist[q] = (double)X[(size_t)ref_idx[(size_t)t * k + q] * S + s];
            if (m < min_ref_targets) continue; /* lib.rs:254-256 */
            double x = (double)X[(size_t)t * S + s];
            double z_score = fabs(x - orc_stats_mean(ref_dist, m)) / orc_stats_std_dev(ref_dist, m);
            double rel = x / orc_stats_mean(ref_dist, m);
            if (z_score > z && fabs(rel - 1.0) > rel_thr) num_calls[t] += 1; /* lib.rs:264-266 */
        }
    }
    free(ref_dist);
}

/* W5  write_output filters (lib-model-wis/src/lib.rs:362-377) */
void orc_wis_filters(const uint32_t *ref_len, const uint32_t *num_calls, uint32_t T,
                     uint32_t min_ref_targets, uint32_t max_samples_reliable, uint8_t *filt) {
    for (uint32_t t = 0; t < T; ++t) {
        uint8_t f = 0;
        if (ref_len[t] < min_ref_targets) f |= 1;
        if (num_calls[t] > max_samples_reliable) f |= 2;
        filt[t] = f;
    }
}

/* ======================================================================================== */
/* M1  mod-coverage  (lib-mod-cov/src/lib.rs:76-164, 209-240)                               */
/* ======================================================================================== */

void orc_modcov(const float *cv, uint32_t T, const uint8_t *model_filt, const uint32_t *ref_idx,
                const uint32_t *ref_len, uint32_t k, float *ref_mean, float *ref_sd,
                float *cv_rel, float *cvz, float *cv2, uint8_t *status) {
    double *ref = (double *)malloc((size_t)(k ? k : 1) * sizeof(double));
    for (uint32_t t = 0; t < T; ++t) {
        status[t] = 0;
        ref_mean[t] = ref_sd[t] = cv_rel[t] = cvz[t] = cv2[t] = f32_missing();
        if (model_filt[t] != 0) continue; /* lib.rs:104-113: any non-PASS filter => skipped */
        const uint32_t m = ref_len[t];
        if (m == 0) continue; /* lib.rs:146 */
        double norm_cov = (double)cv[t]; /* lib.rs:52-55, 121-123 */
        for (uint32_t q = 0; q < m; ++q) ref[q] = (double)cv[ref_idx[(size_t)t * k + q]];
        double mean = orc_stats_mean(ref, m);   /* lib.rs:147 */
        double sd = orc_stats_std_dev(ref, m);  /* lib.rs:148 */
        double rel = norm_cov / mean;           /* lib.rs:78-80 */
        double zs = (norm_cov - mean) / sd;     /* lib.rs:83-85 */
        ref_mean[t] = (float)mean;              /* lib.rs:210-212 */
        ref_sd[t] = (float)sd;                  /* lib.rs:213-215 */
        cv_rel[t] = (float)rel;                 /* lib.rs:217-219 */
        cvz[t] = (float)zs;                     /* lib.rs:220-222 */
        status[t] = 1;
        if (isfinite(rel)) { /* lib.rs:223-232 */
            cv2[t] = (rel == 0.0) ? f32_missing() : (float)log2(rel);
            status[t] |= 2;
        }
    }
    free(ref);
}

/* ======================================================================================== */
/* pile threshold (lib-coverage/src/lib.rs:265-351) and piles (piles.rs:78-127)             */
/* ======================================================================================== */

static double ln_factorial(uint32_t x) { return lgamma((double)x + 1.0); }
/* statrs Poisson::ln_pmf */
static double pois_ln_pmf(double lambda, uint32_t x) {
    return -lambda + (double)x * log(lambda) - ln_factorial(x);
}
/* statrs Poisson::cdf(x) = 1 - P(x+1, lambda) = sum_{i<=x} pmf(i) for integral x */
static double pois_cdf(double lambda, uint32_t x) {
    double acc = 0.0;
    for (uint32_t i = 0; i <= x; ++i) acc += exp(pois_ln_pmf(lambda, i));
    return acc < 1.0 ? acc : 1.0;
}
/* Rust `f64 as u32`: saturating, NaN -> 0 */
static uint32_t f64_as_u32(double v) {
    if (isnan(v)) return 0;
    if (v <= 0.0) return 0;
    if (v >= 4294967295.0) return 4294967295u;
    return (uint32_t)v;
}

int orc_pile_threshold(const uint64_t *y_in, size_t n_in, double fdr_cutoff, uint32_t *out) {
    enum { K_MAX = 20, K_LIM = 21 };
    if (n_in < 3) return ORC_PANIC;      /* bail!("No coverage above 1?") */
    if (y_in[0] != 0) return ORC_PANIC;  /* bail!("Logic error...") */
    size_t new_len = n_in > K_LIM ? n_in : K_LIM;
    uint64_t *y = (uint64_t *)calloc(new_len, sizeof(uint64_t));
    memcpy(y, y_in, n_in * sizeof(uint64_t));
    double y_1 = (double)y[1], y_2 = (double)y[2];
    double lambda = 2.0 * y_2 / y_1;                  /* lib.rs:284 */
    double n = exp(log(y_1) - pois_ln_pmf(lambda, 1)); /* lib.rs:294 */
    double exp_fd[K_LIM], ratios[K_LIM];
    uint64_t obs_d[K_LIM];
    for (uint32_t x = 0; x < K_LIM; ++x) /* lib.rs:298-306 */
        exp_fd[x] = (x == 0) ? 0.0 : n * (1.0 - pois_cdf(lambda, x));
    memset(obs_d, 0, sizeof(obs_d));
    for (uint32_t i = 1; i < K_MAX - 1; ++i) /* lib.rs:311-317 */
        for (uint32_t j = i + 1; j < K_MAX; ++j) obs_d[i] += y[j];
    for (uint32_t i = 0; i < K_LIM; ++i) { /* lib.rs:321-332 */
        double o = (double)obs_d[i];
        ratios[i] = (o != 0.0) ? exp_fd[i] / o : 0.0;
    }
    free(y);
    for (uint32_t i = 1; i < K_LIM; ++i) { /* lib.rs:336-347 */
        if (ratios[i] < fdr_cutoff) {
            uint32_t chosen = i;
            double kv = (double)(chosen + 1); /* k[chosen-1], k = 2..=20 */
            *out = f64_as_u32(kv + (ratios[chosen - 1] - fdr_cutoff) /
                                       ceil(ratios[chosen - 1] - ratios[chosen]));
            return ORC_OK;
        }
    }
    return ORC_PANIC; /* bail!("Could not find appropriate peak threshold") */
}

uint64_t orc_piles_from_depths(const uint32_t *depths, uint64_t n, uint32_t pile_max_gap,
                               uint32_t *pstart, uint32_t *pend, uint32_t *psize) {
    uint64_t m = 0;
    int have = 0;
    uint32_t s = 0, e = 0, sz = 0;
    for (uint64_t p = 0; p < n; ++p) { /* piles.rs:84-117 */
        uint32_t pos = (uint32_t)p, depth = depths[p];
        if (depth > 0) {
            if (have) {
                if (e + pile_max_gap >= pos) {
                    e = pos + 1;
                    sz = depth > sz ? depth : sz;
                } else {
                    if (sz > 0) {
                        pstart[m] = s, pend[m] = e, psize[m] = sz;
                        ++m;
                    }
                    s = pos, e = pos + 1, sz = depth;
                }
            } else {
                have = 1;
                s = pos, e = pos + 1, sz = depth;
            }
        }
    }
    if (have && sz > 0) { /* piles.rs:120-124 */
        pstart[m] = s, pend[m] = e, psize[m] = sz;
        ++m;
    }
    return m;
}

/* ======================================================================================== */
/* H  builder-defined 5-state HMM (no reference implementation; DESIGN.md "HMM spec")       */
/* ======================================================================================== */

void orc_hmm_make_tables(double sigma_s, double sigma0, double eps_cn, double p_move,
                         orc_hmm_tables *t) {
    const int K = ORC_HMM_STATES;
    for (int c = 0; c < K; ++c) {
        double cn = (double)c > eps_cn ? (double)c : eps_cn;
        t->mu[c] = cn / 2.0;
        double sd = sigma_s * sqrt(t->mu[c]) + sigma0;
        t->inv_sd[c] = 1.0 / sd;
        t->log_norm[c] = -log(sd) - 0.5 * log(2.0 * M_PI);
        t->log_start[c] = log(1.0 / (double)K);
        for (int d = 0; d < K; ++d)
            t->log_trans[c][d] = (c == d) ? log(1.0 - (double)(K - 1) * p_move) : log(p_move);
        t->mu_f[c] = (float)t->mu[c];
        t->inv_sd_f[c] = (float)(t->inv_sd[c] * ORC_HMM_ZSCALE); /* folds the 1/2 and the score unit into z */
        t->log_norm_sc[c] = (float)(t->log_norm[c] * ORC_HMM_SCALE);
    }
    t->sc_start = (int32_t)llrint(t->log_start[0] * ORC_HMM_SCALE);
    t->sc_stay = (int32_t)llrint(t->log_trans[0][0] * ORC_HMM_SCALE);
    t->sc_move = (int32_t)llrint(t->log_trans[0][1] * ORC_HMM_SCALE);
}

/* Integer emission score: four f32 operations, one rounding each (no contraction),
 * floored at -64 nat, rounded to the nearest integer (ties to even) by the 1.5 * 2^23 trick,
 * which is exact for |score| <= 2^22. */
void orc_hmm_emit_scores(float cv, const orc_hmm_tables *t, int32_t *e) {
    for (int s = 0; s < ORC_HMM_STATES; ++s) {
        if (!isfinite(cv)) {
            e[s] = 0;
            continue;
        }
        float t1 = cv - t->mu_f[s];
        float z = t1 * t->inv_sd_f[s];
        float q = z * z;
        float ll = t->log_norm_sc[s] - q;
        if (!(ll >= -ORC_HMM_FLOOR)) ll = -ORC_HMM_FLOOR;
        float r = ll + 12582912.0f;
        uint32_t bits;
        memcpy(&bits, &r, 4);
        e[s] = (int32_t)(bits - 0x4B400000u);
    }
}

/* emission log-likelihood; operation order is part of the spec (each step rounds once) */
static inline void hmm_emit(double x, int missing, const orc_hmm_tables *t, double *e) {
    for (int s = 0; s < ORC_HMM_STATES; ++s) {
        if (missing) {
            e[s] = 0.0;
        } else {
            double t1 = x - t->mu[s];
            double z = t1 * t->inv_sd[s];
            double q = z * z;
            double h = 0.5 * q;
            e[s] = t->log_norm[s] - h;
        }
    }
}

void orc_hmm_viterbi(const float *cv, uint64_t n, const orc_hmm_tables *t, uint8_t *state) {
    const int K = ORC_HMM_STATES;
    if (n == 0) return;
    uint8_t *bp = (uint8_t *)malloc(n * K);
    int64_t d[ORC_HMM_STATES], nd[ORC_HMM_STATES];
    int32_t e[ORC_HMM_STATES];
    orc_hmm_emit_scores(cv[0], t, e);
    for (int s = 0; s < K; ++s) d[s] = (int64_t)t->sc_start + e[s];
    for (uint64_t b = 1; b < n; ++b) {
        orc_hmm_emit_scores(cv[b], t, e);
        for (int s = 0; s < K; ++s) {
            int64_t best = d[0] + (s == 0 ? t->sc_stay : t->sc_move);
            int arg = 0;
            for (int r = 1; r < K; ++r) {
                int64_t c = d[r] + (r == s ? t->sc_stay : t->sc_move);
                if (c > best) { /* strict: ties keep the lowest state */
                    best = c;
                    arg = r;
                }
            }
            nd[s] = best + e[s];
            bp[b * K + s] = (uint8_t)arg;
        }
        for (int s = 0; s < K; ++s) d[s] = nd[s];
    }
    int arg = 0;
    for (int s = 1; s < K; ++s)
        if (d[s] > d[arg]) arg = s;
    for (uint64_t b = n; b-- > 0;) {
        state[b] = (uint8_t)arg;
        if (b > 0) arg = bp[b * K + arg];
    }
    free(bp);
}

/* The same recurrence in log-space FP64, strictly sequential -- what BASELINE.json's north_star words for the
 * segmentation ("log-space FP64 with the reference's tie-breaking").  The product runs on the integer scores
 * above (exactly associative, hence chunk-parallel); this function exists so that the rate at which the two
 * specs disagree can be reported (bench.py, cfg4: `fp64_path_disagreement`).  Emissions as in hmm_emit, no floor;
 * ties keep the lowest predecessor / the first maximum, as above. */
void orc_hmm_viterbi_f64(const float *cv, uint64_t n, const orc_hmm_tables *t, uint8_t *state) {
    const int K = ORC_HMM_STATES;
    if (n == 0) return;
    uint8_t *bp = (uint8_t *)malloc(n * K);
    double d[ORC_HMM_STATES], nd[ORC_HMM_STATES], e[ORC_HMM_STATES];
    hmm_emit((double)cv[0], !isfinite(cv[0]), t, e);
    for (int s = 0; s < K; ++s) d[s] = t->log_start[s] + e[s];
    for (uint64_t b = 1; b < n; ++b) {
        hmm_emit((double)cv[b], !isfinite(cv[b]), t, e);
        for (int s = 0; s < K; ++s) {
            double best = d[0] + t->log_trans[0][s];
            int arg = 0;
            for (int r = 1; r < K; ++r) {
                double c = d[r] + t->log_trans[r][s];
                if (c > best) {
                    best = c;
                    arg = r;
                }
            }
            nd[s] = best + e[s];
            bp[b * K + s] = (uint8_t)arg;
        }
        for (int s = 0; s < K; ++s) d[s] = nd[s];
    }
    int arg = 0;
    for (int s = 1; s < K; ++s)
        if (d[s] > d[arg]) arg = s;
    for (uint64_t b = n; b-- > 0;) {
        state[b] = (uint8_t)arg;
        if (b > 0) arg = bp[b * K + arg];
    }
    free(bp);
}

static inline double lse5(const double *v) {
    double m = v[0];
    for (int s = 1; s < ORC_HMM_STATES; ++s)
        if (v[s] > m) m = v[s];
    if (!(m > -INFINITY)) return m;
    double acc = 0.0;
    for (int s = 0; s < ORC_HMM_STATES; ++s) acc += exp(v[s] - m);
    return m + log(acc);
}

void orc_hmm_posterior(const float *cv, uint64_t n, const orc_hmm_tables *t, double *post) {
    const int K = ORC_HMM_STATES;
    if (n == 0) return;
    double *alpha = (double *)malloc(n * K * sizeof(double));
    double e[ORC_HMM_STATES], v[ORC_HMM_STATES], beta[ORC_HMM_STATES], nb[ORC_HMM_STATES];
    hmm_emit((double)cv[0], !isfinite(cv[0]), t, e);
    for (int s = 0; s < K; ++s) alpha[s] = t->log_start[s] + e[s];
    for (uint64_t b = 1; b < n; ++b) {
        hmm_emit((double)cv[b], !isfinite(cv[b]), t, e);
        for (int s = 0; s < K; ++s) {
            for (int r = 0; r < K; ++r) v[r] = alpha[(b - 1) * K + r] + t->log_trans[r][s];
            alpha[b * K + s] = e[s] + lse5(v);
        }
    }
    for (int s = 0; s < K; ++s) beta[s] = 0.0;
    for (uint64_t b = n; b-- > 0;) {
        for (int s = 0; s < K; ++s) v[s] = alpha[b * K + s] + beta[s];
        double norm = lse5(v);
        for (int s = 0; s < K; ++s) post[b * K + s] = exp(v[s] - norm);
        if (b > 0) {
            hmm_emit((double)cv[b], !isfinite(cv[b]), t, e);
            for (int r = 0; r < K; ++r) {
                for (int s = 0; s < K; ++s) v[s] = t->log_trans[r][s] + e[s] + beta[s];
                nb[r] = lse5(v);
            }
            for (int r = 0; r < K; ++r) beta[r] = nb[r];
        }
    }
    free(alpha);
}
