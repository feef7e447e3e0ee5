// coverage.cu -- the three BamRecordAggregator implementations of lib-coverage as sm_100a kernels.
//
//   FragmentsGenomeWide   (lib-coverage/src/agg.rs:68-257)   frag_bin_gw_kernel
//   FragmentsTargetRegions(lib-coverage/src/agg.rs:261-444)  frag_bin_tgt_kernel   (coverage_tgt.cuh)
//   Coverage              (lib-coverage/src/agg.rs:468-600)  depth_tile_kernel ... (coverage_depth.cuh)
//
// All three are HBM-streaming integer kernels: 128-bit coalesced loads of the SoA read batch,
// per-read predicate evaluation in registers, and warp-aggregated reductions so that the number
// of memory atomics is one per (warp, distinct bin) instead of one per read.
#include <new>
#include <vector>

#include "common.cuh"

using namespace cnvb;

namespace {

constexpr uint32_t kDefaultPlpMask = 0x704;  // UNMAP | SECONDARY | QCFAIL | DUP

// ---- read filter (fragment kinds), agg.rs:116-120 ----------------------------------------
// fails if: mapq < min | secondary|supplementary|duplicate|qcfail | clip ratio too low |
//           paired && (!proper || insert_size <= 0)
__device__ __forceinline__ bool frag_pass(uint32_t f, uint32_t mapq, uint32_t min_mapq) {
    const uint32_t bad_always = 0x100u | 0x800u | 0x400u | 0x200u | CNVB_FLAG_CLIPFAIL;
    const uint32_t t = f ^ 0x2u;  // bit 1 now means "not a proper pair"
    const uint32_t bad_paired = (f & 1u) ? (t & (0x2u | CNVB_FLAG_ISIZE_NONPOS)) : 0u;
    return (mapq >= min_mapq) && ((f & bad_always) == 0u) && (bad_paired == 0u);
}

// Add `cnt` to counters[bin] for every lane, with one RED per distinct bin in the warp.
// Must be called by all 32 lanes (converged).  Coordinate-sorted input makes the loop run once
// for almost every warp.
__device__ __forceinline__ void warp_bin_add(uint32_t *__restrict__ counters, uint32_t bin,
                                             uint32_t cnt) {
    const uint32_t lane = lane_id();
    unsigned todo = __ballot_sync(0xffffffffu, cnt != 0u);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const uint32_t b = __shfl_sync(0xffffffffu, bin, leader);
        const unsigned same = __ballot_sync(0xffffffffu, bin == b) & todo;
        const uint32_t v = ((same >> lane) & 1u) ? cnt : 0u;
        const uint32_t total = __reduce_add_sync(0xffffffffu, v);
        if ((int)lane == leader) atomicAdd(&counters[b], total);
        todo &= ~same;
    }
}

// pile mask: is `c` inside one of the disjoint, sorted intervals?  (agg.rs:142)
__device__ __forceinline__ bool pile_hit(const uint32_t *__restrict__ ps,
                                         const uint32_t *__restrict__ pe, uint32_t n, uint32_t c) {
    // last interval with start <= c
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t m = (lo + hi) >> 1;
        if (__ldg(ps + m) <= c) lo = m + 1; else hi = m;
    }
    return lo > 0 && c < __ldg(pe + lo - 1);
}

struct GwArgs {
    const int32_t *mid;
    const uint8_t *mapq;
    const uint16_t *flags;
    uint64_t n;
    uint32_t *counters;
    uint32_t num_bins;
    FastDiv div;
    uint32_t magic_up;  // ceil(2^32 / wl)
    uint32_t wl_small;  // wl <= 32768: magic_up division is exact below 4*wl
    uint32_t min_mapq;
    const uint32_t *pile_start, *pile_end;
    uint32_t n_pile;
    uint32_t *stats;  // [0] processed, [1] skipped, [2] out-of-range centres
};

// One element: returns weight (0/1) and bin; updates the per-thread tallies.
template <bool PILES>
__device__ __forceinline__ uint32_t gw_weight(const GwArgs &a, uint32_t mid, uint32_t mapq,
                                              uint32_t f, uint32_t &n_proc, uint32_t &n_skip) {
    if (!frag_pass(f, mapq, a.min_mapq)) return 0u;
    n_proc += 1;
    if (PILES) {
        if (pile_hit(a.pile_start, a.pile_end, a.n_pile, mid)) {
            n_skip += 1;
            return 0u;
        }
    }
    return 1u;
}

// ---- A2: FragmentsGenomeWideAggregator::put_bam_record over a whole batch ------------------
// Each lane owns 4 consecutive reads per step: mid as one int4 (128-bit), mapq as one u32,
// flags as one uint2.  Algorithmic traffic 7 B/read.
//
// The kernel is instruction-issue bound (7 bytes per read is very little data), so the per-read
// work is cut to the filter plus three compares.  Coordinate-sorted input means the 128 reads of
// one warp step fall into a handful of adjacent windows: the warp takes the window A of its
// first counting read as anchor and every lane classifies its reads into the four windows
// A-1 .. A+2 with three unsigned compares, accumulating four 8-bit counters packed in one
// register (32 lanes x 4 reads = 128 <= 255, no carry between bytes).  ONE warp REDUX then yields
// all four window totals and lanes 0..3 issue at most one RED each.  Reads outside the four
// windows (possible for any input, rare for sorted reads) take a per-read path.
// fail bits of two packed 16-bit flag words at once (see frag_pass): non-zero half = read fails
__device__ __forceinline__ uint32_t frag_fail2(uint32_t F) {
    const uint32_t G = F ^ 0x00020002u;                          // bit 1: not a proper pair
    const uint32_t T = (F & 0x00010001u) * 0x2002u + 0x1F001F00u;  // paired ? 0x3F02 : 0x1F00
    return G & T;
}

// 1 << s with PTX semantics: shift amounts of 32 or more give 0 (no branch, no select)
__device__ __forceinline__ uint32_t shl_clamp1(uint32_t s) {
    uint32_t r;
    asm("shl.b32 %0, 1, %1;" : "=r"(r) : "r"(s));
    return r;
}

// Pass bits of four consecutive reads, as bit 7 of byte e (0x80808080-masked):
//   * mapq >= minq for the four packed bytes at once: t = (mq | H) - (minq4 & ~H) cannot borrow
//     between bytes and has bit 7 = (low 7 bits of mapq >= low 7 bits of minq); the top bits
//     decide when they differ;
//   * frag_fail2 leaves a non-zero 16-bit half (<= 0x3F02) for a failing read: + 0x7FFF carries
//     it into bit 15 of the half without touching the neighbour; PRMT gathers the four bits.
__device__ __forceinline__ uint32_t frag_pass4(uint32_t mq, uint2 fl, uint32_t minq4) {
    constexpr uint32_t H = 0x80808080u;
    const uint32_t t = (mq | H) - (minq4 & ~H);
    const uint32_t d = mq ^ minq4;
    const uint32_t ge = (~d & t) | (d & mq);
    const uint32_t nz0 = frag_fail2(fl.x) + 0x7fff7fffu, nz1 = frag_fail2(fl.y) + 0x7fff7fffu;
    const uint32_t fail = __byte_perm(nz0, nz1, 0x7531);
    return ge & ~fail & H;
}

template <bool PILES, bool WL_SMALL, int Q>
__device__ __forceinline__ void gw_region(const GwArgs &a) {
    // A warp step covers Q consecutive groups of 32 quads (Q x 128 reads): lane l owns quad
    // qb + 32 u + l of group u, so every load instruction of the warp is one contiguous
    // 512 / 128 / 256-byte run.  The per-step warp work (anchor, division, REDUX, counter REDs)
    // is paid once per Q x 128 reads; the per-read work is integer arithmetic on packed words
    // without predicates (the earlier predicate-per-read form ran out of predicate registers).
    const uint32_t nq = (uint32_t)(a.n >> 2);  // full quads (a launch covers < 2^32 reads)
    const uint32_t stride = gridDim.x * blockDim.x * Q;
    const int4 *__restrict__ mid4 = reinterpret_cast<const int4 *>(a.mid);
    const uint32_t *__restrict__ mq4 = reinterpret_cast<const uint32_t *>(a.mapq);
    const uint2 *__restrict__ fl4 = reinterpret_cast<const uint2 *>(a.flags);
    const uint32_t lane = lane_id();
    const uint32_t wl = a.div.d;
    const uint32_t minq4 = a.min_mapq * 0x01010101u;  // (launcher: min_mapq <= 255)
    constexpr uint32_t kFailFlags = 0x01000100u;  // "secondary" in both halves: never counted
    uint32_t n_proc = 0, n_skip = 0, n_oob = 0;

    // warp-uniform trip count so that the warp collectives below stay converged
    uint32_t qb = (blockIdx.x * blockDim.x + threadIdx.x - lane) * Q;
    // software pipeline: the loads of step i+1 are in flight while step i is classified
    int4 m[Q];
    uint32_t mq[Q];
    uint2 fl[Q];
#pragma unroll
    for (int u = 0; u < Q; ++u) {
        m[u] = make_int4(0, 0, 0, 0);
        mq[u] = 0;
        fl[u] = make_uint2(kFailFlags, kFailFlags);
        const uint32_t i = qb + 32u * u + lane;
        if (i < nq) {
            m[u] = ld_stream(mid4 + i);
            mq[u] = ld_stream(mq4 + i);
            fl[u] = ld_stream(fl4 + i);
        }
    }
    for (; qb < nq; qb += stride) {
        uint32_t mids[Q][4], pass[Q];
        const uint32_t nb = qb + stride;
#pragma unroll
        for (int u = 0; u < Q; ++u) {
            mids[u][0] = (uint32_t)m[u].x;
            mids[u][1] = (uint32_t)m[u].y;
            mids[u][2] = (uint32_t)m[u].z;
            mids[u][3] = (uint32_t)m[u].w;
            pass[u] = frag_pass4(mq[u], fl[u], minq4);
            const uint32_t nx = nb + 32u * u + lane;
            fl[u] = make_uint2(kFailFlags, kFailFlags);
            if (nx < nq && nb >= stride) {  // (second test: no wrap-around of the 32-bit index)
                m[u] = ld_stream(mid4 + nx);
                mq[u] = ld_stream(mq4 + nx);
                fl[u] = ld_stream(fl4 + nx);
            }
            if (PILES) {
                n_proc += __popc(pass[u]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if ((pass[u] >> (8 * e + 7) & 1u) && pile_hit(a.pile_start, a.pile_end, a.n_pile, mids[u][e])) {
                        pass[u] &= ~(0x80u << (8 * e));
                        n_skip += 1;
                    }
                }
            }
        }
        // anchor window: the smallest first-of-quad counting read of the step's first group;
        // if none of those 32 reads counts the anchor is far away and the step takes the
        // per-read path below
        uint32_t any_pass = pass[0];
#pragma unroll
        for (int u = 1; u < Q; ++u) any_pass |= pass[u];
        if (!__any_sync(0xffffffffu, any_pass != 0u)) continue;
        const uint32_t amid = __reduce_min_sync(0xffffffffu, (pass[0] & 0x80u) ? mids[0][0] : 0xffffffffu);
        if (amid >= 0x80000000u) {
            // No usable anchor (none of the 32 candidates counts, or a negative centre): window
            // offsets relative to a base near 2^32 would be ambiguous modulo 2^32, so every counting
            // read of the step takes the per-read path.  (Never happens for sorted, sane input.)
#pragma unroll
            for (int u = 0; u < Q; ++u) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (pass[u] >> (8 * e + 7) & 1u) {
                        if (!PILES) n_proc += 1;
                        const uint32_t b = fastdiv(mids[u][e], a.div);
                        if (b < a.num_bins) atomicAdd(&a.counters[b], 1u); else n_oob += 1;
                    }
                }
            }
            continue;
        }
        const uint32_t A = fastdiv(amid, a.div);
        // base < 2^31 (or -wl for A == 0): r = mid - base is the true difference or >= 2^31
        const uint32_t base = (A - 1u) * wl;
        uint32_t acc[Q], n_out = 0u;
#pragma unroll
        for (int u = 0; u < Q; ++u) {
            acc[u] = 0u;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                // window offset k = (mid - base) / wl by a multiply-high with the rounded-UP magic:
                // exact for r < 4*wl whenever 4*wl^2 <= 2^32, and any r >= 4*wl still yields k >= 4.
                // A failing read gets bit 31 of r set (k >= 2^31 / wl >= 4); 8 k < 2^32 for wl >= 16,
                // so its shift amount is >= 32 and it contributes nothing.
                const uint32_t failbit = ~(pass[u] << (24 - 8 * e)) & 0x80000000u;
                const uint32_t r = (mids[u][e] - base) | failbit;
                if (WL_SMALL) {
                    acc[u] += shl_clamp1(__umulhi(r, a.magic_up) << 3);
                } else {
                    const uint32_t k = fastdiv(r, a.div);
                    acc[u] += (k < 4u && !failbit) ? (1u << (8u * k)) : 0u;
                }
            }
            // counting reads of the quad that fell outside the four windows
            n_out += __popc(pass[u]) - ((acc[u] * 0x01010101u) >> 24);
        }
        // 8-bit fields: one group adds at most 128 to a field over the warp, two could reach 256,
        // so every group has its own REDUX
        uint32_t cnt4 = 0u, tot_all = 0u;
#pragma unroll
        for (int u = 0; u < Q; ++u) {
            const uint32_t tot = __reduce_add_sync(0xffffffffu, acc[u]);
            cnt4 += (tot >> (8u * (lane & 3u))) & 0xffu;
            tot_all += (tot * 0x01010101u) >> 24;
        }
        // without piles every read passing the filter is counted: tally per warp, not per read
        if (!PILES && lane == 0u) n_proc += tot_all;
        if (lane < 4u) {
            const uint32_t bin = A - 1u + lane;
            if (cnt4) {
                if (bin < a.num_bins) atomicAdd(&a.counters[bin], cnt4); else n_oob += cnt4;
            }
        }
        if (__any_sync(0xffffffffu, n_out != 0u)) {  // reads far from the anchor: one RED each
#pragma unroll
            for (int u = 0; u < Q; ++u) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (n_out != 0u && (pass[u] >> (8 * e + 7) & 1u)) {
                        const uint32_t r = mids[u][e] - base;
                        const uint32_t k = WL_SMALL ? __umulhi(r, a.magic_up) : fastdiv(r, a.div);
                        if (k >= 4u) {
                            if (!PILES) n_proc += 1;
                            const uint32_t b = fastdiv(mids[u][e], a.div);
                            if (b < a.num_bins) atomicAdd(&a.counters[b], 1u); else n_oob += 1;
                        }
                    }
                }
            }
        }
    }
    // tail (n % 4 reads): one thread, direct atomics
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (uint64_t i = (uint64_t)nq << 2; i < a.n; ++i) {
            const uint32_t mid = (uint32_t)a.mid[i];
            if (gw_weight<PILES>(a, mid, a.mapq[i], a.flags[i], n_proc, n_skip)) {
                const uint32_t b = fastdiv(mid, a.div);
                if (b >= a.num_bins) n_oob += 1; else atomicAdd(&a.counters[b], 1u);
            }
        }
    }
    n_proc = __reduce_add_sync(0xffffffffu, n_proc);
    n_skip = __reduce_add_sync(0xffffffffu, n_skip);
    n_oob = __reduce_add_sync(0xffffffffu, n_oob);
    if (lane == 0) {
        if (n_proc) atomicAdd(&a.stats[0], n_proc);
        if (n_skip) atomicAdd(&a.stats[1], n_skip);
        if (n_oob) atomicAdd(&a.stats[2], n_oob);
    }
}

template <bool PILES, bool WL_SMALL, int Q>
__global__ void __launch_bounds__(256, 4) frag_bin_gw_kernel(GwArgs a) {
    gw_region<PILES, WL_SMALL, Q>(a);
}

// All regions of a run in ONE launch (the loop of lib_coverage::run, lib-coverage/src/lib.rs:768-781, over
// device-resident read columns): every CTA takes its grid-stride share of region 0, then of region 1, ... --
// no barrier in between, so the warps flow from one region into the next and the launch pays ONE fill and one
// tail instead of one per region (fit over the 22 contigs of config 2: ~5 us each, a tenth of the binning
// time).  The regions' arguments travel in the kernel parameters (constant bank, indexed by the region).
constexpr uint32_t kGwBatchMax = 40;  // x 88 bytes of arguments: inside the 4 KB of kernel parameters
struct GwBatch {
    uint32_t n;
    GwArgs r[kGwBatchMax];
};
template <bool PILES, bool WL_SMALL, int Q>
__global__ void __launch_bounds__(256, 4) frag_bin_gw_batch_kernel(const __grid_constant__ GwBatch b) {
    for (uint32_t i = 0; i < b.n; ++i) gw_region<PILES, WL_SMALL, Q>(b.r[i]);
}

// Same operator for SoA pointers that are not 16/4/8-byte aligned: scalar coalesced loads.
template <bool PILES>
__global__ void __launch_bounds__(256, 4) frag_bin_gw_scalar_kernel(GwArgs a) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t n_proc = 0, n_skip = 0, n_oob = 0;
    const uint64_t first = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t ib = first - lane_id(); ib < a.n; ib += stride) {
        const uint64_t i = ib + lane_id();
        uint32_t cnt = 0, b = 0;
        if (i < a.n) {
            const uint32_t mid = (uint32_t)a.mid[i];
            cnt = gw_weight<PILES>(a, mid, a.mapq[i], a.flags[i], n_proc, n_skip);
            b = fastdiv(mid, a.div);
            if (cnt && b >= a.num_bins) {
                n_oob += 1;
                cnt = 0;
            }
        }
        warp_bin_add(a.counters, b, cnt);
    }
    n_proc = __reduce_add_sync(0xffffffffu, n_proc);
    n_skip = __reduce_add_sync(0xffffffffu, n_skip);
    n_oob = __reduce_add_sync(0xffffffffu, n_oob);
    if (lane_id() == 0) {
        if (n_proc) atomicAdd(&a.stats[0], n_proc);
        if (n_skip) atomicAdd(&a.stats[1], n_skip);
        if (n_oob) atomicAdd(&a.stats[2], n_oob);
    }
}

// ---- get_stats for the window kinds (agg.rs:201-219, 231-244, 576-587) ----------------------
__global__ void window_stats_kernel(const uint32_t *__restrict__ counters, uint32_t num_bins,
                                    uint32_t wl, uint32_t region_end,
                                    const uint32_t *__restrict__ ps,
                                    const uint32_t *__restrict__ pe, uint32_t n_pile,
                                    float *__restrict__ cov, uint32_t *__restrict__ start,
                                    uint32_t *__restrict__ end, float *__restrict__ frm,
                                    float *__restrict__ mean_mapq) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= num_bins) return;
    const uint64_t ws64 = (uint64_t)w * wl, we64 = ws64 + wl;
    const uint32_t ws = (uint32_t)ws64, we = (uint32_t)we64;
    if (cov) cov[w] = (float)counters[w];
    if (start) start[w] = ws;
    if (end) end[w] = we64 < region_end ? we : region_end;
    if (mean_mapq) mean_mapq[w] = 60.0f;  // agg.rs:242 "TODO: actually compute"
    if (frm) {
        uint32_t masked = 0;
        if (n_pile) {
            // first interval whose end > ws (ends are sorted because piles are disjoint+sorted)
            uint32_t lo = 0, hi = n_pile;
            while (lo < hi) {
                const uint32_t m = (lo + hi) >> 1;
                if (pe[m] <= ws) lo = m + 1; else hi = m;
            }
            for (uint32_t k = lo; k < n_pile && ps[k] < we; ++k) {
                const uint32_t e = min(we, pe[k]), s = max(ws, ps[k]);
                masked += e - s;
            }
        }
        frm[w] = (float)masked / (float)wl;
    }
}

// get_stats for targets (agg.rs:421-431)
__global__ void target_stats_kernel(const uint32_t *__restrict__ counters, uint32_t n,
                                    const uint32_t *__restrict__ ts,
                                    const uint32_t *__restrict__ te, float *__restrict__ cov,
                                    uint32_t *__restrict__ start, uint32_t *__restrict__ end,
                                    float *__restrict__ mean_mapq) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (cov) cov[t] = (float)counters[t];
    if (start) start[t] = ts[t];
    if (end) end[t] = te[t];
    if (mean_mapq) mean_mapq[t] = 60.0f;
}

}  // namespace

#include "coverage_tgt.cuh"
#include "coverage_depth.cuh"

// ---------------------------------------------------------------------------------------------
struct cnvb_cov {
    cnvb_ctx *ctx = nullptr;
    cnvb_cov_opts opts{};
    int kind = 0;
    uint32_t region_end = 0;
    uint64_t num_regions = 0;
    int state = 0;  // 0 open, 1 finish enqueued, 2 finished (tallies on the host, checked)
    int finish_rc = 0;
    cudaEvent_t done = nullptr;
    uint32_t *h_tally = nullptr;  // pinned slot in ctx->tally_slots
    uint32_t h_stats[4] = {0, 0, 0, 0};
    // device state
    uint32_t *d_counters = nullptr;  // u32[num_regions]
    uint32_t *d_stats = nullptr;     // u32[4]
    uint32_t *d_tgt_start = nullptr, *d_tgt_end = nullptr, *d_tgt_pmax = nullptr;
    uint32_t n_tgt = 0;
    uint32_t *d_pile_start = nullptr, *d_pile_end = nullptr;
    uint32_t n_pile = 0;
    // Coverage kind: reads are kept resident until finish()
    DepthState depth;
};

static void cov_free(cnvb_cov *a) {
    if (!a) return;
    cnvb_ctx *ctx = a->ctx;
    pool_release(ctx, a->d_counters);
    pool_release(ctx, a->d_stats);
    pool_release(ctx, a->d_tgt_start);
    pool_release(ctx, a->d_tgt_end);
    pool_release(ctx, a->d_tgt_pmax);
    pool_release(ctx, a->d_pile_start);
    pool_release(ctx, a->d_pile_end);
    a->depth.release(ctx);
    if (a->done) ctx->event_pool.push_back(a->done);
    delete a;
}

extern "C" int cnvb_cov_create(cnvb_ctx *ctx, const cnvb_cov_opts *opts, int kind,
                               uint32_t region_end, const uint32_t *tgt_start,
                               const uint32_t *tgt_end, uint32_t n_tgt,
                               const uint32_t *pile_start, const uint32_t *pile_end,
                               uint32_t n_pile, cnvb_cov **out) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (!opts || !out) return fail(ctx, CNVB_ERR_INVALID, "cnvb_cov_create: null argument");
    *out = nullptr;
    if (kind != CNVB_FRAGMENTS_GENOME_WIDE && kind != CNVB_FRAGMENTS_TARGET_REGIONS &&
        kind != CNVB_COVERAGE)
        return fail(ctx, CNVB_ERR_INVALID, "Invalid combination of coverage/on-target regions");
    if (kind != CNVB_FRAGMENTS_TARGET_REGIONS && opts->window_length == 0)
        return fail(ctx, CNVB_ERR_INVALID, "Window length must be set here");
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    cnvb_cov *a = new (std::nothrow) cnvb_cov();
    if (!a) return fail(ctx, CNVB_ERR_NOMEM, "out of host memory");
    a->ctx = ctx;
    a->opts = *opts;
    if (a->opts.plp_flag_mask == 0) a->opts.plp_flag_mask = kDefaultPlpMask;
    a->kind = kind;
    a->region_end = region_end;
    a->h_tally = ctx->tally_slots + 4 * (ctx->tally_next++ % kTallySlots);
    if (ctx->capturing) {  // (a plan's aggregators outlive any number of other calls: their own slot)
        a->h_tally = static_cast<uint32_t *>(capture_pin(ctx, nullptr, 16));
        if (!a->h_tally) {
            delete a;
            return fail(ctx, CNVB_ERR_NOMEM, "pinned tally slot");
        }
    }
    if (kind == CNVB_FRAGMENTS_TARGET_REGIONS) {
        a->num_regions = n_tgt;
        a->n_tgt = n_tgt;
    } else {
        const uint64_t wl = opts->window_length;
        a->num_regions = ((uint64_t)region_end + wl - 1) / wl;  // agg.rs:97, :487
    }
    auto bail = [&](int rc) {
        cov_free(a);
        return rc;
    };
    cudaError_t e;
    const size_t nreg = a->num_regions ? a->num_regions : 1;
    if ((e = pool_alloc_t(ctx, nreg, &a->d_counters)) != cudaSuccess ||
        (e = pool_alloc_t(ctx, 4, &a->d_stats)) != cudaSuccess)
        return bail(fail_cuda(ctx, e, "allocating aggregator counters"));
    cudaMemsetAsync(a->d_counters, 0, nreg * sizeof(uint32_t), ctx->stream);
    cudaMemsetAsync(a->d_stats, 0, 4 * sizeof(uint32_t), ctx->stream);
    bool uploaded = false;
    auto upload = [&](uint32_t **dst, const uint32_t *src, uint32_t n) -> cudaError_t {
        cudaError_t err = pool_alloc_t(ctx, n, dst);
        if (err != cudaSuccess) return err;
        if (n) err = cudaMemcpyAsync(*dst, src, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
        uploaded = true;
        return err;
    };
    if (kind == CNVB_FRAGMENTS_TARGET_REGIONS) {
        if (n_tgt && (!tgt_start || !tgt_end))
            return bail(fail(ctx, CNVB_ERR_INVALID, "target arrays missing"));
        for (uint32_t t = 0; t < n_tgt; ++t) {
            if (tgt_start[t] >= tgt_end[t])
                return bail(fail(ctx, CNVB_ERR_REFERENCE_PANIC,
                                 "target %u is empty or reversed (IntervalTree::insert panics)", t));
            if (t && tgt_start[t] < tgt_start[t - 1])
                return bail(fail(ctx, CNVB_ERR_INVALID, "targets must be sorted by start"));
        }
        // running maximum of target ends: lower bound of the overlap search (coverage_tgt.cuh)
        std::vector<uint32_t> pmax(n_tgt ? n_tgt : 1);
        uint32_t run = 0;
        for (uint32_t t = 0; t < n_tgt; ++t) {
            run = tgt_end[t] > run ? tgt_end[t] : run;
            pmax[t] = run;
        }
        if ((e = upload(&a->d_tgt_start, tgt_start, n_tgt)) != cudaSuccess ||
            (e = upload(&a->d_tgt_end, tgt_end, n_tgt)) != cudaSuccess ||
            (e = upload(&a->d_tgt_pmax, pmax.data(), n_tgt)) != cudaSuccess ||
            (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)  // pmax dies with this scope
            return bail(fail_cuda(ctx, e, "uploading targets"));
    }
    if (kind == CNVB_FRAGMENTS_GENOME_WIDE && n_pile) {
        if (!pile_start || !pile_end) return bail(fail(ctx, CNVB_ERR_INVALID, "pile arrays missing"));
        for (uint32_t k = 0; k < n_pile; ++k) {
            if (pile_start[k] >= pile_end[k] || (k && pile_start[k] < pile_end[k - 1]))
                return bail(fail(ctx, CNVB_ERR_INVALID, "piles must be non-empty, disjoint and sorted"));
        }
        if ((e = upload(&a->d_pile_start, pile_start, n_pile)) != cudaSuccess ||
            (e = upload(&a->d_pile_end, pile_end, n_pile)) != cudaSuccess)
            return bail(fail_cuda(ctx, e, "uploading piles"));
        a->n_pile = n_pile;
    }
    // caller-owned host arrays may be freed after we return
    if (uploaded && ctx->capturing)
        return bail(fail(ctx, CNVB_ERR_INVALID, "a coverage plan takes regions without per-region host tables (targets, piles)"));
    if (uploaded && (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
        return bail(fail_cuda(ctx, e, "initialising aggregator"));
    *out = a;
    return CNVB_OK;
}

extern "C" void cnvb_cov_destroy(cnvb_cov *a) {
    if (!a) return;
    cudaSetDevice(a->ctx->device);
    // pooled blocks are reused in stream order, so no synchronisation is needed here
    cov_free(a);
}

namespace {

// Copies one SoA column chunk to device staging if the batch is in host memory.
template <typename T>
int stage_column(cnvb_ctx *ctx, const T *src, uint64_t n, int mem, T *dst_stage, const T **out) {
    if (mem == CNVB_MEM_DEVICE) {
        *out = src;
        return 0;
    }
    CNVB_CUDA(ctx, cudaMemcpyAsync(dst_stage, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream),
              "host-to-device copy of a read column");
    *out = dst_stage;
    return 0;
}

constexpr uint64_t kChunkReads = 32ull << 20;  // host batches are streamed in 32 Mi-read chunks

// everything of the kernel arguments that comes from the aggregator (the read columns are the caller's)
void gw_fill_args(const cnvb_cov *a, GwArgs *g) {
    g->counters = a->d_counters;
    g->num_bins = (uint32_t)a->num_regions;
    g->div = make_fastdiv(a->opts.window_length);
    g->magic_up = (uint32_t)(((1ull << 32) + a->opts.window_length - 1) / a->opts.window_length);
    g->wl_small = a->opts.window_length >= 16 && a->opts.window_length <= 32768;
    g->min_mapq = a->opts.min_mapq;
    g->pile_start = a->d_pile_start;
    g->pile_end = a->d_pile_end;
    g->n_pile = a->n_pile;
    g->stats = a->d_stats;
}
bool gw_aligned(const GwArgs &g) {
    return ((uintptr_t)g.mid % 16 == 0) && ((uintptr_t)g.mapq % 4 == 0) && ((uintptr_t)g.flags % 8 == 0) &&
           g.min_mapq <= 255u;  // (packed byte compare)
}

int put_gw(cnvb_cov *a, const cnvb_read_soa *b) {
    cnvb_ctx *ctx = a->ctx;
    if (!b->mid || !b->mapq || !b->flags)
        return fail(ctx, CNVB_ERR_INVALID, "GenomeWide batches need mid, mapq and flags");
    constexpr uint64_t kMaxLaunchReads = 1ull << 31;
    const uint64_t chunk = b->mem == CNVB_MEM_DEVICE ? (b->n < kMaxLaunchReads ? b->n : kMaxLaunchReads)
                                                     : (b->n < kChunkReads ? b->n : kChunkReads);
    int32_t *s_mid = nullptr;
    uint8_t *s_mq = nullptr;
    uint16_t *s_fl = nullptr;
    if (b->mem != CNVB_MEM_DEVICE) {
        CNVB_TRY(ensure_stage(ctx, Carver::need({chunk * 4, chunk * 1, chunk * 2})));
        Carver c(ctx->stage);
        s_mid = c.take<int32_t>(chunk);
        s_mq = c.take<uint8_t>(chunk);
        s_fl = c.take<uint16_t>(chunk);
    }
    for (uint64_t off = 0; off < b->n; off += chunk) {
        const uint64_t n = (b->n - off) < chunk ? (b->n - off) : chunk;
        GwArgs g{};
        CNVB_TRY(stage_column(ctx, b->mid + off, n, b->mem, s_mid, &g.mid));
        CNVB_TRY(stage_column(ctx, b->mapq + off, n, b->mem, s_mq, &g.mapq));
        CNVB_TRY(stage_column(ctx, b->flags + off, n, b->mem, s_fl, &g.flags));
        g.n = n;
        gw_fill_args(a, &g);
        const bool aligned = gw_aligned(g);
        const uint64_t work = aligned ? (n + 3) / 4 : n;
        const uint64_t work_blocks = (work + 255) / 256;
        static const int gw_quads = [] {  // quads per lane and step (tuning knob; 2 measured best)
            const char *e = getenv("CNVB_GW_QUADS");
            return e && atoi(e) == 1 ? 1 : 2;
        }();
        {
            KernelSpan span(ctx, "frag_bin_gw_kernel");
            if (aligned) {
#define CNVB_GW_LAUNCH_Q(P, S, Q) \
    frag_bin_gw_kernel<P, S, Q><<<persistent_grid(ctx, frag_bin_gw_kernel<P, S, Q>, 256, 0, (work_blocks + Q - 1) / Q), 256, 0, ctx->stream>>>(g)
#define CNVB_GW_LAUNCH(P, S) \
    do { if (gw_quads == 2) CNVB_GW_LAUNCH_Q(P, S, 2); else CNVB_GW_LAUNCH_Q(P, S, 1); } while (0)
                if (a->n_pile) {
                    if (g.wl_small) CNVB_GW_LAUNCH(true, true); else CNVB_GW_LAUNCH(true, false);
                } else {
                    if (g.wl_small) CNVB_GW_LAUNCH(false, true); else CNVB_GW_LAUNCH(false, false);
                }
#undef CNVB_GW_LAUNCH
#undef CNVB_GW_LAUNCH_Q
            } else {
                if (a->n_pile)
                    frag_bin_gw_scalar_kernel<true><<<persistent_grid(ctx, frag_bin_gw_scalar_kernel<true>, 256, 0, work_blocks), 256, 0, ctx->stream>>>(g);
                else
                    frag_bin_gw_scalar_kernel<false><<<persistent_grid(ctx, frag_bin_gw_scalar_kernel<false>, 256, 0, work_blocks), 256, 0, ctx->stream>>>(g);
            }
        }
        CNVB_LAUNCHED(ctx, "frag_bin_gw_kernel");
    }
    return CNVB_OK;
}

}  // namespace

// FragmentsGenomeWide, all regions of a run at once: aggs[i] takes batches[i] (device-resident columns).
// Returns 0 when the records are on their way (ONE launch per kGwBatchMax regions on the context's stream),
// 1 when the batch form does not apply -- the caller then feeds the aggregators one by one -- or an error.
int cnvb::gw_batch_put(cnvb_ctx *ctx, cnvb_cov *const *aggs, const cnvb_read_soa *batches, size_t batch_stride,
                       uint32_t n, int ctas_per_sm) {
    if (n < 2) return 1;
    std::vector<GwArgs> args(n);
    bool piles = false;
    for (uint32_t i = 0; i < n; ++i) {
        const cnvb_cov *a = aggs[i];
        const cnvb_read_soa *b = reinterpret_cast<const cnvb_read_soa *>(reinterpret_cast<const char *>(batches) + i * batch_stride);
        if (!a || a->kind != CNVB_FRAGMENTS_GENOME_WIDE || a->state != 0 || b->mem != CNVB_MEM_DEVICE) return 1;
        if (b->n >= (1ull << 31) || (b->n && (!b->mid || !b->mapq || !b->flags))) return 1;
        if (a->opts.window_length != aggs[0]->opts.window_length || a->opts.min_mapq != aggs[0]->opts.min_mapq) return 1;
        GwArgs &g = args[i];
        g = GwArgs{};
        g.mid = b->mid;
        g.mapq = b->mapq;
        g.flags = b->flags;
        g.n = b->n;
        gw_fill_args(a, &g);
        if (b->n && !gw_aligned(g)) return 1;
        piles = piles || a->n_pile != 0;
    }
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    const bool wl_small = args[0].wl_small != 0;
    for (uint32_t first = 0; first < n; first += kGwBatchMax) {
        GwBatch batch{};
        batch.n = std::min<uint32_t>(kGwBatchMax, n - first);
        uint64_t quads = 0;
        for (uint32_t i = 0; i < batch.n; ++i) {
            batch.r[i] = args[first + i];
            quads = std::max<uint64_t>(quads, (batch.r[i].n + 3) / 4);
        }
        const uint64_t work_blocks = (((quads + 255) / 256) + 1) / 2;  // (two quads per lane and step)
        {
            KernelSpan span(ctx, "frag_bin_gw_kernel");
            // (a persistent grid must be RESIDENT as a whole: beside another persistent kernel -- the reference
            // statistics on their side stream -- the caller asks for fewer CTAs per SM than the kernel could have alone)
            const uint64_t cap = ctas_per_sm > 0 ? (uint64_t)ctas_per_sm * (uint64_t)ctx->sm_count : ~0ull;
#define CNVB_GW_BATCH_LAUNCH(P, S) \
    frag_bin_gw_batch_kernel<P, S, 2><<<(unsigned)std::min<uint64_t>(cap, persistent_grid(ctx, frag_bin_gw_batch_kernel<P, S, 2>, 256, 0, work_blocks)), 256, 0, ctx->stream>>>(batch)
            if (piles) {
                if (wl_small) CNVB_GW_BATCH_LAUNCH(true, true); else CNVB_GW_BATCH_LAUNCH(true, false);
            } else {
                if (wl_small) CNVB_GW_BATCH_LAUNCH(false, true); else CNVB_GW_BATCH_LAUNCH(false, false);
            }
#undef CNVB_GW_BATCH_LAUNCH
        }
        CNVB_LAUNCHED(ctx, "frag_bin_gw_kernel");
    }
    return 0;
}

extern "C" int cnvb_cov_put_records(cnvb_cov *a, const cnvb_read_soa *b) {
    if (!a) return CNVB_ERR_INVALID;
    cnvb_ctx *ctx = a->ctx;
    if (!b) return fail(ctx, CNVB_ERR_INVALID, "cnvb_cov_put_records: null batch");
    if (a->state != 0) return fail(ctx, CNVB_ERR_STATE, "put_records after finish");
    if (b->mem != CNVB_MEM_HOST && b->mem != CNVB_MEM_DEVICE)
        return fail(ctx, CNVB_ERR_INVALID, "batch.mem must be CNVB_MEM_HOST or CNVB_MEM_DEVICE");
    if (b->n == 0) return CNVB_OK;
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    switch (a->kind) {
    case CNVB_FRAGMENTS_GENOME_WIDE: return put_gw(a, b);
    case CNVB_FRAGMENTS_TARGET_REGIONS:
        return put_targets(ctx, b, a->d_tgt_start, a->d_tgt_end, a->d_tgt_pmax, a->n_tgt,
                           a->opts.min_mapq, a->d_counters, a->d_stats);
    default: return a->depth.append(ctx, b);
    }
}

// Enqueue the end-of-stream work without waiting for it (results stay in stream order).
extern "C" int cnvb_cov_finish_async(cnvb_cov *a) {
    if (!a) return CNVB_ERR_INVALID;
    cnvb_ctx *ctx = a->ctx;
    if (a->state != 0) return CNVB_OK;
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    if (a->kind == CNVB_COVERAGE)
        CNVB_TRY(a->depth.finish(ctx, a->opts, a->region_end, (uint32_t)a->num_regions, a->d_stats));
    CNVB_CUDA(ctx, cudaMemcpyAsync(a->h_tally, a->d_stats, 16, cudaMemcpyDeviceToHost, ctx->stream),
              "reading aggregator tallies");
    if (!a->done) a->done = KernelSpan::get_event(ctx);
    CNVB_CUDA(ctx, cudaEventRecord(a->done, ctx->stream), "recording aggregator completion");
    a->state = 1;
    return CNVB_OK;
}

extern "C" int cnvb_cov_finish(cnvb_cov *a) {
    if (!a) return CNVB_ERR_INVALID;
    cnvb_ctx *ctx = a->ctx;
    if (a->state == 2) return a->finish_rc;
    CNVB_TRY(cnvb_cov_finish_async(a));
    CNVB_CUDA(ctx, cudaEventSynchronize(a->done), "finishing aggregator");
    return cov_collect(a);
}

int cnvb::cov_collect(cnvb_cov *a) {
    if (!a) return CNVB_ERR_INVALID;
    cnvb_ctx *ctx = a->ctx;
    memcpy(a->h_stats, a->h_tally, 16);
    a->state = 2;
    a->finish_rc = CNVB_OK;
    if (a->h_stats[3])
        a->finish_rc = fail(ctx, CNVB_ERR_INVALID, "records are not coordinate-sorted (%u inversions)",
                            a->h_stats[3]);
    else if (a->h_stats[2] && a->kind == CNVB_COVERAGE)
        a->finish_rc = fail(ctx, CNVB_ERR_REFERENCE_PANIC,
                            "%u alignment(s) reach past the last window; the reference panics indexing "
                            "self.coverage (lib-coverage/src/agg.rs:506)", a->h_stats[2]);
    else if (a->h_stats[2])
        a->finish_rc = fail(ctx, CNVB_ERR_REFERENCE_PANIC,
                            "%u fragment centre(s) fall outside the windows of the region; the reference "
                            "panics at mapq_sums[bin] (lib-coverage/src/agg.rs:144); unpaired reads in "
                            "genome-wide mode get centre pos+end_pos (agg.rs:127-132)", a->h_stats[2]);
    return a->finish_rc;
}

extern "C" uint64_t cnvb_cov_num_regions(const cnvb_cov *a) { return a ? a->num_regions : 0; }
extern "C" uint32_t cnvb_cov_num_processed(const cnvb_cov *a) { return a ? a->h_stats[0] : 0; }
extern "C" uint32_t cnvb_cov_num_skipped(const cnvb_cov *a) { return a ? a->h_stats[1] : 0; }

extern "C" int cnvb_cov_get_counts(cnvb_cov *a, uint32_t *counts, int mem) {
    if (!a) return CNVB_ERR_INVALID;
    cnvb_ctx *ctx = a->ctx;
    if (a->state == 0) return fail(ctx, CNVB_ERR_STATE, "get_counts before finish");
    if (a->kind == CNVB_COVERAGE) return fail(ctx, CNVB_ERR_INVALID, "Coverage has no fragment counts");
    if (!counts) return fail(ctx, CNVB_ERR_INVALID, "null output");
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    CNVB_CUDA(ctx, cudaMemcpyAsync(counts, a->d_counters, a->num_regions * sizeof(uint32_t),
                                   mem == CNVB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                   ctx->stream), "copying counters");
    if (mem != CNVB_MEM_DEVICE) CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "copying counters");
    return CNVB_OK;
}

extern "C" int cnvb_cov_get_stats(cnvb_cov *a, float *cov, float *cov_sd, uint32_t *start,
                                  uint32_t *end, float *frac_removed, float *mean_mapq, int mem) {
    if (!a) return CNVB_ERR_INVALID;
    cnvb_ctx *ctx = a->ctx;
    if (a->state == 0) return fail(ctx, CNVB_ERR_STATE, "get_stats before finish");
    if (cov_sd && a->kind != CNVB_COVERAGE)
        return fail(ctx, CNVB_ERR_INVALID, "cov_sd is None for fragment aggregators (agg.rs:237,424)");
    if (frac_removed && a->kind != CNVB_FRAGMENTS_GENOME_WIDE)
        return fail(ctx, CNVB_ERR_INVALID, "frac_removed is None for this aggregator (agg.rs:428,584)");
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    const uint64_t n = a->num_regions;
    if (n == 0) return CNVB_OK;
    float *d_cov = cov, *d_sd = cov_sd, *d_frm = frac_removed, *d_mq = mean_mapq;
    uint32_t *d_start = start, *d_end = end;
    if (mem != CNVB_MEM_DEVICE) {
        CNVB_TRY(ensure_scratch(ctx, Carver::need({n * 4, n * 4, n * 4, n * 4, n * 4, n * 4})));
        Carver c(ctx->scratch);
        d_cov = cov ? c.take<float>(n) : nullptr;
        d_sd = cov_sd ? c.take<float>(n) : nullptr;
        d_start = start ? c.take<uint32_t>(n) : nullptr;
        d_end = end ? c.take<uint32_t>(n) : nullptr;
        d_frm = frac_removed ? c.take<float>(n) : nullptr;
        d_mq = mean_mapq ? c.take<float>(n) : nullptr;
    }
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (a->kind == CNVB_FRAGMENTS_TARGET_REGIONS) {
        target_stats_kernel<<<blocks, 256, 0, ctx->stream>>>(a->d_counters, (uint32_t)n, a->d_tgt_start,
                                                             a->d_tgt_end, d_cov, d_start, d_end, d_mq);
        CNVB_LAUNCHED(ctx, "target_stats_kernel");
    } else if (a->kind == CNVB_FRAGMENTS_GENOME_WIDE) {
        window_stats_kernel<<<blocks, 256, 0, ctx->stream>>>(
            a->d_counters, (uint32_t)n, a->opts.window_length, a->region_end, a->d_pile_start,
            a->d_pile_end, a->n_pile, d_cov, d_start, d_end, d_frm, d_mq);
        CNVB_LAUNCHED(ctx, "window_stats_kernel");
    } else {
        CNVB_TRY(a->depth.get_stats(ctx, a->opts.window_length, a->region_end, (uint32_t)n, d_cov, d_sd,
                                    d_start, d_end, d_mq));
    }
    if (mem != CNVB_MEM_DEVICE) {
        auto back = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
            return dst ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream) : cudaSuccess;
        };
        cudaError_t e;
        if ((e = back(cov, d_cov, n * 4)) != cudaSuccess || (e = back(cov_sd, d_sd, n * 4)) != cudaSuccess ||
            (e = back(start, d_start, n * 4)) != cudaSuccess || (e = back(end, d_end, n * 4)) != cudaSuccess ||
            (e = back(frac_removed, d_frm, n * 4)) != cudaSuccess ||
            (e = back(mean_mapq, d_mq, n * 4)) != cudaSuccess)
            return fail_cuda(ctx, e, "copying aggregation stats to the host");
        CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "copying aggregation stats to the host");
    }
    return CNVB_OK;
}

// ---- A6: FORMAT value assembly (lib-coverage/src/lib.rs:623-668) ----------------------------
namespace {
__global__ void cov_records_kernel(uint64_t n, const float *__restrict__ cov,
                                   const float *__restrict__ cov_sd,
                                   const uint32_t *__restrict__ start,
                                   const uint32_t *__restrict__ end,
                                   const float *__restrict__ frm, int is_targets,
                                   float min_raw_coverage, float min_window_remaining,
                                   float *__restrict__ lcv, float *__restrict__ lcvsd,
                                   uint8_t *__restrict__ ft) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float c = cov[i];
    const float length = (float)(end[i] - start[i]);             // lib.rs:637
    uint8_t f = 0;
    if (is_targets && c < min_raw_coverage) f |= CNVB_FT_LOWCOV;  // lib.rs:627-631
    lcv[i] = __fdiv_rn(c, length);                               // lib.rs:639
    if (cov_sd && lcvsd) lcvsd[i] = __fdiv_rn(cov_sd[i], length); // lib.rs:641-645
    if (frm && frm[i] > min_window_remaining) f |= CNVB_FT_PILE;  // lib.rs:647-654
    if (ft) ft[i] = f;
}
}  // namespace

extern "C" int cnvb_cov_records(cnvb_ctx *ctx, uint64_t n, const float *cov, const float *cov_sd,
                                const uint32_t *start, const uint32_t *end, const float *frac_removed,
                                int is_targets, const cnvb_cov_opts *opts, float *lcv, float *lcvsd,
                                uint8_t *ft, int mem) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (!opts || (n && (!cov || !start || !end || !lcv)))
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_cov_records: null argument");
    if (n == 0) return CNVB_OK;
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    const float *d_cov = cov, *d_sd = cov_sd, *d_frm = frac_removed;
    const uint32_t *d_start = start, *d_end = end;
    float *d_lcv = lcv, *d_lcvsd = lcvsd;
    uint8_t *d_ft = ft;
    if (mem != CNVB_MEM_DEVICE) {
        CNVB_TRY(ensure_scratch(ctx, Carver::need({n * 4, n * 4, n * 4, n * 4, n * 4, n * 4, n * 4, n})));
        Carver c(ctx->scratch);
        auto up = [&](const void *src, size_t bytes, auto *dst) -> cudaError_t {
            return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
        };
        float *t_cov = c.take<float>(n), *t_sd = c.take<float>(n), *t_frm = c.take<float>(n);
        uint32_t *t_start = c.take<uint32_t>(n), *t_end = c.take<uint32_t>(n);
        d_lcv = c.take<float>(n);
        d_lcvsd = c.take<float>(n);
        d_ft = c.take<uint8_t>(n);
        cudaError_t e = up(cov, n * 4, t_cov);
        if (e == cudaSuccess && cov_sd) e = up(cov_sd, n * 4, t_sd);
        if (e == cudaSuccess && frac_removed) e = up(frac_removed, n * 4, t_frm);
        if (e == cudaSuccess) e = up(start, n * 4, t_start);
        if (e == cudaSuccess) e = up(end, n * 4, t_end);
        if (e != cudaSuccess) return fail_cuda(ctx, e, "uploading stats");
        d_cov = t_cov;
        d_sd = cov_sd ? t_sd : nullptr;
        d_frm = frac_removed ? t_frm : nullptr;
        d_start = t_start;
        d_end = t_end;
        if (!lcvsd) d_lcvsd = nullptr;
    }
    cov_records_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
        n, d_cov, d_sd, d_start, d_end, d_frm, is_targets, (float)opts->min_raw_coverage,
        opts->min_window_remaining, d_lcv, d_lcvsd, d_ft);
    CNVB_LAUNCHED(ctx, "cov_records_kernel");
    if (mem != CNVB_MEM_DEVICE) {
        CNVB_CUDA(ctx, cudaMemcpyAsync(lcv, d_lcv, n * 4, cudaMemcpyDeviceToHost, ctx->stream), "copy lcv");
        if (lcvsd && cov_sd)
            CNVB_CUDA(ctx, cudaMemcpyAsync(lcvsd, d_lcvsd, n * 4, cudaMemcpyDeviceToHost, ctx->stream), "copy lcvsd");
        if (ft) CNVB_CUDA(ctx, cudaMemcpyAsync(ft, d_ft, n, cudaMemcpyDeviceToHost, ctx->stream), "copy ft");
        CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "cov_records");
    }
    return CNVB_OK;
}
