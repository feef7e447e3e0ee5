// coverage_depth.cuh -- CoverageAggregator (lib-coverage/src/agg.rs:468-600).
// Included by coverage.cu only.
//
// The reference walks htslib's pileup one reference base at a time.  Here base-level depth is a
// difference array + prefix scan that never leaves the SM: each CTA owns a tile of kTile
// reference positions, scatters +1/-1 events of the reads overlapping the tile into a
// shared-memory difference array, scans it with warp shuffles and reduces straight to
// per-window integer moments (sum d, sum d^2, number of pileup columns, first column).  Neither
// the difference array nor the depth array ever touches HBM.
//
// One 32-bit word carries both counts the reference needs per base:
//   bits 16..31  "cols"  alignments htslib's pileup reports (flag & plp_mask == 0)
//   bits  0..15  "depth" of those, the ones passing agg.rs:548-557
// (sums of +-0x10001 wrap modulo 2^32 but every prefix is a valid pair while cols < 65536;
//  htslib itself caps pileup depth at 8000.)
//
// The reference's window state machine lags by one pileup column (agg.rs:523-573; SURVEY.md
// Appendix A.3).  Its net effect on window W's depth vector is closed-form:
//   * W's own first column f_W is missing unless W is the first covered window,
//   * index (f_N mod wl) is overwritten with depth(f_N), N = next covered window,
//   * the last covered window is never pushed (stays 0/0) when it has >= 2 columns; if it has
//     exactly one, it receives the previous covered window's pending vector and that window
//     stays 0/0 instead.
// window_finalize_kernel applies exactly that to the integer moments.
#pragma once

namespace {

constexpr uint32_t kTile = 8192;  // reference positions per CTA (32 KB of shared memory)
constexpr uint32_t kNone = 0xffffffffu;

// reference filter inside the pileup column, agg.rs:548-557
__device__ __forceinline__ uint32_t depth_pass(uint32_t f, uint32_t mapq, uint32_t min_mapq) {
    return ((f & (0x100u | 0x400u | 0x800u | 0x200u)) == 0u) && (mapq >= min_mapq) &&
           (!(f & 1u) || (f & 2u));
}

__device__ __forceinline__ uint64_t lower_bound_i32(const int32_t *__restrict__ a, uint64_t n, int64_t key) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t m = (lo + hi) >> 1;
        if ((int64_t)__ldg(a + m) < key) lo = m + 1; else hi = m;
    }
    return lo;
}

struct DepthArgs {
    const int32_t *pos, *end;
    const uint8_t *mapq;
    const uint16_t *flags;
    uint64_t n;
    uint32_t plp_mask, min_mapq, wl, num_bins;
    uint64_t L;  // num_bins * wl
    // per window
    unsigned long long *S, *Q, *first;
    uint32_t *ncol, *next;
    // scalars: [0] first covered window, [1] last covered, [2] covered window before the last,
    //          [3] max alignment span
    uint32_t *scal;
    uint32_t *stats;  // [2] out-of-range alignments, [3] sortedness violations
    float *mean, *sd;
    unsigned long long *tile_lo, *tile_hi;  // [tiles] alignments that can overlap the tile: [lo, hi)
    uint32_t *blk_min, *blk_top;            // [window blocks] smallest covered window | [2 x] the two largest
};

// pass 0: maximum span, bounds and order checks (four alignments per thread and step when the
// columns are 16 / 8-byte aligned, which pooled and torch allocations are)
__global__ void depth_prepare_kernel(DepthArgs a) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t span = 0, oob = 0, unsorted = 0;
    auto one = [&](int32_t prev, int32_t p, int32_t e, uint32_t f, bool has_prev) {
        if (has_prev && prev > p) unsorted += 1;
        if ((f & a.plp_mask) == 0u && p >= 0) {
            if (e > p) span = max(span, (uint32_t)(e - p));
            if ((uint64_t)(int64_t)e > a.L) oob += 1;
        }
    };
    const bool vec = ((reinterpret_cast<uintptr_t>(a.pos) | reinterpret_cast<uintptr_t>(a.end)) & 15u) == 0u &&
                     (reinterpret_cast<uintptr_t>(a.flags) & 7u) == 0u;
    const uint64_t nq = vec ? a.n >> 2 : 0;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
        const int4 p = ld_stream(reinterpret_cast<const int4 *>(a.pos) + q);
        const int4 e = ld_stream(reinterpret_cast<const int4 *>(a.end) + q);
        const uint2 f = ld_stream(reinterpret_cast<const uint2 *>(a.flags) + q);
        const int32_t prev = q ? a.pos[4 * q - 1] : 0;
        one(prev, p.x, e.x, f.x & 0xffffu, q != 0);
        one(p.x, p.y, e.y, f.x >> 16, true);
        one(p.y, p.z, e.z, f.y & 0xffffu, true);
        one(p.z, p.w, e.w, f.y >> 16, true);
    }
    for (uint64_t i = (nq << 2) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride)
        one(i ? a.pos[i - 1] : 0, a.pos[i], a.end[i], a.flags[i], i != 0);
    span = __reduce_max_sync(0xffffffffu, span);
    oob = __reduce_add_sync(0xffffffffu, oob);
    unsorted = __reduce_add_sync(0xffffffffu, unsorted);
    if (lane_id() == 0) {
        if (span) atomicMax(&a.scal[3], span);
        if (oob) atomicAdd(&a.stats[2], oob);
        if (unsorted) atomicAdd(&a.stats[3], unsorted);
    }
}

// pass 0b: which alignments can overlap a tile -- one thread per tile, so that the latency of the
// two binary searches overlaps across tiles instead of stalling every tile CTA at its start
__global__ void depth_ranges_kernel(DepthArgs a, uint64_t tiles) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tiles) return;
    const uint64_t t0 = t * kTile;
    const uint64_t t1 = (t0 + kTile < a.L) ? t0 + kTile : a.L;
    a.tile_lo[t] = lower_bound_i32(a.pos, a.n, (int64_t)t0 - (int64_t)a.scal[3] + 1);
    a.tile_hi[t] = lower_bound_i32(a.pos, a.n, (int64_t)t1);
}

// pass 1: one CTA per tile of kTile positions
__global__ void __launch_bounds__(256) depth_tile_kernel(DepthArgs a) {
    __shared__ __align__(16) uint32_t d[kTile];
    __shared__ uint32_t s_warp_tot[8];
    __shared__ uint32_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t t0 = (uint64_t)blockIdx.x * kTile;
    const uint64_t t1 = (t0 + kTile < a.L) ? t0 + kTile : a.L;
    const uint64_t r_lo = a.tile_lo[blockIdx.x], r_hi = a.tile_hi[blockIdx.x];

    for (uint32_t i = tid; i < kTile / 4; i += 256) reinterpret_cast<uint4 *>(d)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) s_carry = 0u;
    __syncthreads();
    // events of the alignments overlapping the tile
    uint32_t carry = 0u;
    auto event = [&](uint32_t f, int64_t p, int64_t e, uint32_t mq) {
        if ((f & a.plp_mask) == 0u && p >= 0 && e > (int64_t)t0 && e > p) {
            const uint32_t v = 0x10000u + depth_pass(f, mq, a.min_mapq);
            if (p < (int64_t)t0) carry += v; else atomicAdd(&d[p - t0], v);
            if (e < (int64_t)t1) atomicAdd(&d[e - t0], 0u - v);
        }
    };
    const bool vec = ((reinterpret_cast<uintptr_t>(a.pos) | reinterpret_cast<uintptr_t>(a.end)) & 15u) == 0u &&
                     (reinterpret_cast<uintptr_t>(a.flags) & 7u) == 0u && (reinterpret_cast<uintptr_t>(a.mapq) & 3u) == 0u;
    if (vec) {
        // four alignments per thread and step (128-bit loads), two steps in flight: a tile's ~2500 alignments
        // are three dependent trips to HBM instead of ten
        const uint64_t nq_full = a.n >> 2;  // whole quads of the columns; the last n % 4 alignments go one by one
        const uint64_t q_lo = r_lo >> 2, q_hi = min((r_hi + 3) >> 2, nq_full);
        if (tid == 0)
            for (uint64_t i = max(r_lo, nq_full << 2); i < r_hi; ++i) event(a.flags[i], a.pos[i], a.end[i], a.mapq[i]);
        for (uint64_t q = q_lo + tid; q < q_hi; q += 512) {
            const uint64_t q2 = q + 256;
            const bool two = q2 < q_hi;
            const int4 p0 = ld_stream(reinterpret_cast<const int4 *>(a.pos) + q);
            const int4 e0 = ld_stream(reinterpret_cast<const int4 *>(a.end) + q);
            const uint2 f0 = ld_stream(reinterpret_cast<const uint2 *>(a.flags) + q);
            const uint32_t m0 = ld_stream(reinterpret_cast<const uint32_t *>(a.mapq) + q);
            int4 p1 = make_int4(0, 0, 0, 0), e1 = make_int4(0, 0, 0, 0);
            uint2 f1 = make_uint2(0u, 0u);
            uint32_t m1 = 0u;
            if (two) {
                p1 = ld_stream(reinterpret_cast<const int4 *>(a.pos) + q2);
                e1 = ld_stream(reinterpret_cast<const int4 *>(a.end) + q2);
                f1 = ld_stream(reinterpret_cast<const uint2 *>(a.flags) + q2);
                m1 = ld_stream(reinterpret_cast<const uint32_t *>(a.mapq) + q2);
            }
            auto quad = [&](uint64_t qq, const int4 &p, const int4 &e, const uint2 &f, uint32_t m) {
                const uint64_t i0 = qq << 2;  // (the first and the last quad may reach outside [r_lo, r_hi))
                if (i0 >= r_lo && i0 < r_hi) event(f.x & 0xffffu, p.x, e.x, m & 0xffu);
                if (i0 + 1 >= r_lo && i0 + 1 < r_hi) event(f.x >> 16, p.y, e.y, (m >> 8) & 0xffu);
                if (i0 + 2 >= r_lo && i0 + 2 < r_hi) event(f.y & 0xffffu, p.z, e.z, (m >> 16) & 0xffu);
                if (i0 + 3 >= r_lo && i0 + 3 < r_hi) event(f.y >> 16, p.w, e.w, m >> 24);
            };
            quad(q, p0, e0, f0, m0);
            if (two) quad(q2, p1, e1, f1, m1);
        }
    } else {
        for (uint64_t i = r_lo + tid; i < r_hi; i += 256) event(a.flags[i], a.pos[i], a.end[i], a.mapq[i]);
    }
    carry = __reduce_add_sync(0xffffffffu, carry);
    if (lane == 0 && carry) atomicAdd(&s_carry, carry);
    __syncthreads();
    // prefix scan in place: warp w owns d[w*1024 .. w*1024+1023]; a lane takes four consecutive
    // values per step (one 128-bit access, conflict-free), the warp scans the 32 lane totals
    const uint32_t seg = warp * 1024u;
    uint4 *d4 = reinterpret_cast<uint4 *>(d + seg);
    uint32_t tot = 0u;
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) {
        const uint4 v = d4[k * 32u + lane];
        tot += v.x + v.y + v.z + v.w;
    }
    tot = __reduce_add_sync(0xffffffffu, tot);
    if (lane == 0) s_warp_tot[warp] = tot;
    __syncthreads();
    uint32_t run = s_carry;
    for (uint32_t u = 0; u < warp; ++u) run += s_warp_tot[u];
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) {
        uint4 v = d4[k * 32u + lane];
        v.y += v.x;
        v.z += v.y;
        v.w += v.z;
        uint32_t x = v.w;  // lane total -> inclusive scan over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((int)lane >= o) x += y;
        }
        const uint32_t base = run + x - v.w;  // everything before this lane's four values
        v.x += base;
        v.y += base;
        v.z += base;
        v.w += base;
        d4[k * 32u + lane] = v;
        run += __shfl_sync(0xffffffffu, x, 31);
    }
    __syncthreads();
    // integer moments of every window intersecting the tile (one warp per window, strided)
    if (t1 <= t0) return;
    const uint64_t w_first = t0 / a.wl, w_last = (t1 - 1) / a.wl;
    for (uint64_t W = w_first + warp; W <= w_last; W += 8) {
        const uint64_t lo = max(t0, W * a.wl), hi = min(t1, (W + 1) * a.wl);
        unsigned long long S = 0ull, Q = 0ull, key = ~0ull;
        uint32_t nc = 0u;
        for (uint64_t p = lo + lane; p < hi; p += 32) {
            const uint32_t v = d[p - t0];
            const uint32_t dep = v & 0xffffu;
            S += dep;
            Q += (unsigned long long)dep * dep;
            if (v >> 16) {
                nc += 1;
                const unsigned long long kk = ((unsigned long long)p << 32) | v;
                key = kk < key ? kk : key;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            S += __shfl_xor_sync(0xffffffffu, S, o);
            Q += __shfl_xor_sync(0xffffffffu, Q, o);
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o);
            key = ok < key ? ok : key;
        }
        nc = __reduce_add_sync(0xffffffffu, nc);
        if (lane == 0) {
            if (S) atomicAdd(&a.S[W], S);
            if (Q) atomicAdd(&a.Q[W], Q);
            if (nc) {
                atomicAdd(&a.ncol[W], nc);
                atomicMin(&a.first[W], key);
            }
        }
    }
}

// pass 2: next covered window for every window, plus first / last / second-to-last covered window.
// A reverse exclusive min-scan of (covered ? W : none) over the windows: 1024 windows per CTA,
// (2a) per-CTA summaries, (2b) carry from the CTAs to the right + scan inside the CTA.
__global__ void __launch_bounds__(1024) covered_blocks_kernel(DepthArgs a) {
    __shared__ uint32_t s_min[32], s_t1[32], s_t2[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t W = blockIdx.x * 1024u + tid;
    const bool cov = W < a.num_bins && a.ncol[W] > 0u;
    uint32_t mn = cov ? W : kNone;
    // two largest covered windows as (t1 > t2), kNone = absent; kept as W + 1 so that 0 = absent
    uint32_t t1 = cov ? W + 1u : 0u, t2 = 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        const uint32_t o1 = __shfl_xor_sync(0xffffffffu, t1, o), o2 = __shfl_xor_sync(0xffffffffu, t2, o);
        const uint32_t hi = max(t1, o1), lo = max(min(t1, o1), max(t2, o2));  // (distinct windows per lane)
        t1 = hi;
        t2 = lo;
    }
    if (lane == 0) {
        s_min[warp] = mn;
        s_t1[warp] = t1;
        s_t2[warp] = t2;
    }
    __syncthreads();
    if (warp == 0) {
        mn = s_min[lane];
        t1 = s_t1[lane];
        t2 = s_t2[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            const uint32_t o1 = __shfl_xor_sync(0xffffffffu, t1, o), o2 = __shfl_xor_sync(0xffffffffu, t2, o);
            const uint32_t hi = max(t1, o1), lo = max(min(t1, o1), max(t2, o2));
            t1 = hi;
            t2 = lo;
        }
        if (lane == 0) {
            a.blk_min[blockIdx.x] = mn;
            a.blk_top[2 * blockIdx.x] = t1;
            a.blk_top[2 * blockIdx.x + 1] = t2;
        }
    }
}

__global__ void __launch_bounds__(1024) next_covered_kernel(DepthArgs a) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_red[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t nblk = gridDim.x;
    // carry: smallest covered window in the CTAs to the right
    uint32_t carry = kNone;
    for (uint32_t bb = blockIdx.x + 1u + tid; bb < nblk; bb += 1024u) carry = min(carry, a.blk_min[bb]);
    carry = __reduce_min_sync(0xffffffffu, carry);
    if (lane == 0) s_red[warp] = carry;
    __syncthreads();
    carry = s_red[0];
    for (uint32_t u = 1; u < 32; ++u) carry = min(carry, s_red[u]);
    // inside the CTA walk from the right end: thread 0 holds the right-most window of the chunk
    const int64_t W = (int64_t)blockIdx.x * 1024 + 1023 - (int64_t)tid;
    const bool cov = W < (int64_t)a.num_bins && a.ncol[W < (int64_t)a.num_bins ? W : 0] > 0u;
    uint32_t v = cov ? (uint32_t)W : kNone;  // inclusive min over this and earlier threads
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)lane >= o) v = min(v, y);
    }
    if (lane == 31) s_w[warp] = v;
    __syncthreads();
    uint32_t pre = carry;
    for (uint32_t u = 0; u < warp; ++u) pre = min(pre, s_w[u]);
    uint32_t excl = __shfl_up_sync(0xffffffffu, v, 1);  // windows strictly to the right
    if (lane == 0) excl = kNone;
    excl = min(excl, pre);
    if (W < (int64_t)a.num_bins) a.next[W] = excl;
    if (blockIdx.x == 0 && warp == 0) {
        // first covered window = smallest summary; last and the one before it = the two largest
        uint32_t mn = kNone, t1 = 0u, t2 = 0u;
        for (uint32_t bb = lane; bb < nblk; bb += 32u) {
            mn = min(mn, a.blk_min[bb]);
            const uint32_t o1 = a.blk_top[2 * bb], o2 = a.blk_top[2 * bb + 1];
            const uint32_t hi = max(t1, o1), lo = max(min(t1, o1), max(t2, o2));
            t1 = hi;
            t2 = lo;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            const uint32_t o1 = __shfl_xor_sync(0xffffffffu, t1, o), o2 = __shfl_xor_sync(0xffffffffu, t2, o);
            const uint32_t hi = max(t1, o1), lo = max(min(t1, o1), max(t2, o2));
            t1 = hi;
            t2 = lo;
        }
        if (lane == 0) {
            a.scal[0] = mn;
            a.scal[1] = t1 ? t1 - 1u : kNone;
            a.scal[2] = t2 ? t2 - 1u : kNone;
        }
    }
}

__device__ __forceinline__ uint32_t depth_at(const DepthArgs &a, uint64_t q, uint32_t max_span) {
    const uint64_t lo = lower_bound_i32(a.pos, a.n, (int64_t)q - (int64_t)max_span + 1);
    const uint64_t hi = lower_bound_i32(a.pos, a.n, (int64_t)q + 1);
    uint32_t dep = 0;
    for (uint64_t i = lo; i < hi; ++i) {
        const uint32_t f = a.flags[i];
        const int64_t p = a.pos[i], e = a.end[i];
        if ((f & a.plp_mask) == 0u && p >= 0 && p <= (int64_t)q && (int64_t)q < e)
            dep += depth_pass(f, a.mapq[i], a.min_mapq);
    }
    return dep;
}

// The depth vector the reference pushes for covered window W (see header comment), as moments.
__device__ void pushed_moments(const DepthArgs &a, uint32_t W, uint32_t first_cov, uint32_t max_span,
                               long long &S, long long &Q) {
    S = (long long)a.S[W];
    Q = (long long)a.Q[W];
    const unsigned long long kW = a.first[W];
    if (W != first_cov) {
        const long long d1 = (long long)(kW & 0xffffull);
        S -= d1;
        Q -= d1 * d1;
    }
    const uint32_t N = a.next[W];
    if (N != kNone) {
        const unsigned long long kN = a.first[N];
        const uint64_t fN = kN >> 32;
        const long long dN = (long long)(kN & 0xffffull);
        const uint64_t q = (uint64_t)W * a.wl + fN % a.wl;
        long long cur = 0;
        if (!(W != first_cov && q == (kW >> 32))) cur = depth_at(a, q, max_span);
        S += dN - cur;
        Q += dN * dN - cur * cur;
    }
}

// push_window (agg.rs:503-514): mean and sample SD over exactly wl entries, from integer moments
__device__ void moments_to_stats(long long S, long long Q, uint32_t wl, float &mean, float &sd) {
    const double n = (double)wl;
    mean = (float)((double)S / n);  // Stats::mean of integers: the exact sum, one division
    if (wl < 2) {
        sd = 0.0f;
        return;
    }
    double num;  // sum (x - mean)^2 * wl  == Q*wl - S^2
    const unsigned long long uq = (unsigned long long)Q, us = (unsigned long long)S;
    if (__umul64hi(uq, (unsigned long long)wl) == 0ull && __umul64hi(us, us) == 0ull) {
        const unsigned long long qa = uq * wl, sb = us * us;
        num = qa >= sb ? (double)(qa - sb) : 0.0;
    } else {
        num = ((double)Q - (double)S * ((double)S / n)) * n;
        if (num < 0.0) num = 0.0;
    }
    const double var = num / (n * (n - 1.0));
    sd = (float)sqrt(var);
}

// pass 3: one thread per window
__global__ void window_finalize_kernel(DepthArgs a) {
    const uint32_t W = blockIdx.x * blockDim.x + threadIdx.x;
    if (W >= a.num_bins) return;
    float mean = 0.0f, sd = 0.0f;  // CoverageBin::new(), agg.rs:456-463
    const uint32_t first_cov = a.scal[0], last_cov = a.scal[1], prev_cov = a.scal[2];
    const uint32_t max_span = a.scal[3];
    if (a.ncol[W] > 0u) {
        long long S, Q;
        if (W == last_cov) {
            const bool single = a.ncol[W] == 1u;
            if (first_cov == last_cov) {
                if (single) {  // prev_window_id == None at the end: push_window(window_id)
                    pushed_moments(a, W, first_cov, max_span, S, Q);
                    moments_to_stats(S, Q, a.wl, mean, sd);
                }
            } else if (single) {  // pending vector of the previous covered window lands here
                pushed_moments(a, prev_cov, first_cov, max_span, S, Q);
                moments_to_stats(S, Q, a.wl, mean, sd);
            }
        } else if (!(a.next[W] == last_cov && a.ncol[last_cov] == 1u)) {
            pushed_moments(a, W, first_cov, max_span, S, Q);
            moments_to_stats(S, Q, a.wl, mean, sd);
        }
    }
    a.mean[W] = mean;
    a.sd[W] = sd;
}

__global__ void coverage_stats_kernel(const float *__restrict__ mean, const float *__restrict__ sdv,
                                      uint32_t num_bins, uint32_t wl, uint32_t region_end,
                                      float *__restrict__ cov, float *__restrict__ cov_sd,
                                      uint32_t *__restrict__ start, uint32_t *__restrict__ end,
                                      float *__restrict__ mean_mapq) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= num_bins) return;
    const uint64_t ws = (uint64_t)w * wl, we = ws + wl;
    if (cov) cov[w] = mean[w];
    if (cov_sd) cov_sd[w] = sdv[w];
    if (start) start[w] = (uint32_t)ws;
    if (end) end[w] = we < region_end ? (uint32_t)we : region_end;  // agg.rs:583
    if (mean_mapq) mean_mapq[w] = 60.0f;                           // agg.rs:585
}

struct DepthState {
    int32_t *pos = nullptr, *end = nullptr;
    uint8_t *mapq = nullptr;
    uint16_t *flags = nullptr;
    uint64_t n = 0, cap = 0;
    bool owns = true;
    unsigned long long *S = nullptr, *Q = nullptr, *first = nullptr;
    uint32_t *ncol = nullptr, *next = nullptr, *scal = nullptr;
    float *mean = nullptr, *sd = nullptr;
    unsigned long long *tile_lo = nullptr, *tile_hi = nullptr;
    uint32_t *blk_min = nullptr, *blk_top = nullptr;

    void release(cnvb_ctx *ctx) {
        pool_release(ctx, tile_lo);
        pool_release(ctx, tile_hi);
        pool_release(ctx, blk_min);
        pool_release(ctx, blk_top);
        tile_lo = tile_hi = nullptr;
        blk_min = blk_top = nullptr;
        if (owns) {
            pool_release(ctx, pos);
            pool_release(ctx, end);
            pool_release(ctx, mapq);
            pool_release(ctx, flags);
        }
        pool_release(ctx, S);
        pool_release(ctx, Q);
        pool_release(ctx, first);
        pool_release(ctx, ncol);
        pool_release(ctx, next);
        pool_release(ctx, scal);
        pool_release(ctx, mean);
        pool_release(ctx, sd);
        pos = end = nullptr;
        mapq = nullptr;
        flags = nullptr;
        S = Q = first = nullptr;
        ncol = next = scal = nullptr;
        mean = sd = nullptr;
        n = cap = 0;
    }

    int reserve(cnvb_ctx *ctx, uint64_t want) {
        if (want <= cap) return 0;
        uint64_t ncap = cap ? cap * 2 : want;
        if (ncap < want) ncap = want;
        int32_t *np = nullptr, *ne = nullptr;
        uint8_t *nm = nullptr;
        uint16_t *nf = nullptr;
        cudaError_t e;
        if ((e = pool_alloc_t(ctx, ncap, &np)) != cudaSuccess || (e = pool_alloc_t(ctx, ncap, &ne)) != cudaSuccess ||
            (e = pool_alloc_t(ctx, ncap, &nm)) != cudaSuccess || (e = pool_alloc_t(ctx, ncap, &nf)) != cudaSuccess) {
            pool_release(ctx, np);
            pool_release(ctx, ne);
            pool_release(ctx, nm);
            pool_release(ctx, nf);
            return fail_cuda(ctx, e, "growing the resident read store");
        }
        if (n) {
            cudaMemcpyAsync(np, pos, n * 4, cudaMemcpyDeviceToDevice, ctx->stream);
            cudaMemcpyAsync(ne, end, n * 4, cudaMemcpyDeviceToDevice, ctx->stream);
            cudaMemcpyAsync(nm, mapq, n, cudaMemcpyDeviceToDevice, ctx->stream);
            cudaMemcpyAsync(nf, flags, n * 2, cudaMemcpyDeviceToDevice, ctx->stream);
        }
        if (owns) {  // stream-ordered reuse: the copies above are enqueued before any new owner
            pool_release(ctx, pos);
            pool_release(ctx, end);
            pool_release(ctx, mapq);
            pool_release(ctx, flags);
        }
        pos = np;
        end = ne;
        mapq = nm;
        flags = nf;
        cap = ncap;
        owns = true;
        return 0;
    }

    // Coverage needs every alignment overlapping a tile, so batches are kept resident until
    // finish().  A single device-resident batch is used in place (zero-copy).
    int append(cnvb_ctx *ctx, const cnvb_read_soa *b) {
        if (!b->pos || !b->end || !b->mapq || !b->flags)
            return fail(ctx, CNVB_ERR_INVALID, "Coverage batches need pos, end, mapq and flags");
        if (n == 0 && cap == 0 && b->mem == CNVB_MEM_DEVICE) {
            pos = const_cast<int32_t *>(b->pos);
            end = const_cast<int32_t *>(b->end);
            mapq = const_cast<uint8_t *>(b->mapq);
            flags = const_cast<uint16_t *>(b->flags);
            n = b->n;
            cap = b->n;
            owns = false;
            return 0;
        }
        if (!owns) {  // second batch after a borrowed one: take a private copy first
            const int32_t *bp = pos, *be = end;
            const uint8_t *bm = mapq;
            const uint16_t *bf = flags;
            const uint64_t bn = n;
            pos = end = nullptr;
            mapq = nullptr;
            flags = nullptr;
            n = cap = 0;
            owns = true;
            CNVB_TRY(reserve(ctx, bn + b->n));
            cudaMemcpyAsync(pos, bp, bn * 4, cudaMemcpyDeviceToDevice, ctx->stream);
            cudaMemcpyAsync(end, be, bn * 4, cudaMemcpyDeviceToDevice, ctx->stream);
            cudaMemcpyAsync(mapq, bm, bn, cudaMemcpyDeviceToDevice, ctx->stream);
            cudaMemcpyAsync(flags, bf, bn * 2, cudaMemcpyDeviceToDevice, ctx->stream);
            n = bn;
        }
        CNVB_TRY(reserve(ctx, n + b->n));
        const cudaMemcpyKind k = b->mem == CNVB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        CNVB_CUDA(ctx, cudaMemcpyAsync(pos + n, b->pos, b->n * 4, k, ctx->stream), "copying pos");
        CNVB_CUDA(ctx, cudaMemcpyAsync(end + n, b->end, b->n * 4, k, ctx->stream), "copying end");
        CNVB_CUDA(ctx, cudaMemcpyAsync(mapq + n, b->mapq, b->n, k, ctx->stream), "copying mapq");
        CNVB_CUDA(ctx, cudaMemcpyAsync(flags + n, b->flags, b->n * 2, k, ctx->stream), "copying flags");
        n += b->n;
        return 0;
    }

    int finish(cnvb_ctx *ctx, const cnvb_cov_opts &o, uint32_t region_end, uint32_t num_bins,
               uint32_t *d_stats) {
        (void)region_end;
        const size_t nb = num_bins ? num_bins : 1;
        cudaError_t e;
        if ((e = pool_alloc_t(ctx, nb, &S)) != cudaSuccess || (e = pool_alloc_t(ctx, nb, &Q)) != cudaSuccess ||
            (e = pool_alloc_t(ctx, nb, &first)) != cudaSuccess || (e = pool_alloc_t(ctx, nb, &ncol)) != cudaSuccess ||
            (e = pool_alloc_t(ctx, nb, &next)) != cudaSuccess || (e = pool_alloc_t(ctx, 4, &scal)) != cudaSuccess ||
            (e = pool_alloc_t(ctx, nb, &mean)) != cudaSuccess || (e = pool_alloc_t(ctx, nb, &sd)) != cudaSuccess)
            return fail_cuda(ctx, e, "allocating window moments");
        const uint64_t n_tiles = ((uint64_t)num_bins * o.window_length + kTile - 1) / kTile;
        const uint32_t n_wblk = (num_bins + 1023u) / 1024u;
        if ((e = pool_alloc_t(ctx, n_tiles ? n_tiles : 1, &tile_lo)) != cudaSuccess ||
            (e = pool_alloc_t(ctx, n_tiles ? n_tiles : 1, &tile_hi)) != cudaSuccess ||
            (e = pool_alloc_t(ctx, n_wblk ? n_wblk : 1, &blk_min)) != cudaSuccess ||
            (e = pool_alloc_t(ctx, 2 * (n_wblk ? n_wblk : 1), &blk_top)) != cudaSuccess)
            return fail_cuda(ctx, e, "allocating window moments");
        cudaMemsetAsync(S, 0, nb * 8, ctx->stream);
        cudaMemsetAsync(Q, 0, nb * 8, ctx->stream);
        cudaMemsetAsync(first, 0xff, nb * 8, ctx->stream);
        cudaMemsetAsync(ncol, 0, nb * 4, ctx->stream);
        cudaMemsetAsync(scal, 0, 16, ctx->stream);
        DepthArgs a{};
        a.pos = pos;
        a.end = end;
        a.mapq = mapq;
        a.flags = flags;
        a.n = n;
        a.plp_mask = o.plp_flag_mask;
        a.min_mapq = o.min_mapq;
        a.wl = o.window_length;
        a.num_bins = num_bins;
        a.L = (uint64_t)num_bins * o.window_length;
        a.S = S;
        a.Q = Q;
        a.first = first;
        a.ncol = ncol;
        a.next = next;
        a.scal = scal;
        a.stats = d_stats;
        a.mean = mean;
        a.sd = sd;
        a.tile_lo = tile_lo;
        a.tile_hi = tile_hi;
        a.blk_min = blk_min;
        a.blk_top = blk_top;
        if (num_bins == 0) return 0;
        if (n) {
            uint64_t blocks = (n + 255) / 256;
            const uint64_t mb = (uint64_t)ctx->sm_count * 8;
            if (blocks > mb) blocks = mb;
            depth_prepare_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(a);
            CNVB_LAUNCHED(ctx, "depth_prepare_kernel");
            const uint64_t tiles = (a.L + kTile - 1) / kTile;
            depth_ranges_kernel<<<(unsigned)((tiles + 127) / 128), 128, 0, ctx->stream>>>(a, tiles);
            CNVB_LAUNCHED(ctx, "depth_ranges_kernel");
            {
                KernelSpan span(ctx, "depth_tile_kernel");
                depth_tile_kernel<<<(unsigned)tiles, 256, 0, ctx->stream>>>(a);
            }
            CNVB_LAUNCHED(ctx, "depth_tile_kernel");
        }
        covered_blocks_kernel<<<n_wblk, 1024, 0, ctx->stream>>>(a);
        CNVB_LAUNCHED(ctx, "covered_blocks_kernel");
        next_covered_kernel<<<n_wblk, 1024, 0, ctx->stream>>>(a);
        CNVB_LAUNCHED(ctx, "next_covered_kernel");
        window_finalize_kernel<<<(num_bins + 127) / 128, 128, 0, ctx->stream>>>(a);
        CNVB_LAUNCHED(ctx, "window_finalize_kernel");
        // (sortedness, a precondition of put_fetched_records, is checked from the tallies in
        //  cnvb_cov_finish)
        return 0;
    }

    int get_stats(cnvb_ctx *ctx, uint32_t wl, uint32_t region_end, uint32_t num_bins, float *cov,
                  float *cov_sd, uint32_t *start, uint32_t *endp, float *mean_mapq) {
        coverage_stats_kernel<<<(num_bins + 255) / 256, 256, 0, ctx->stream>>>(
            mean, sd, num_bins, wl, region_end, cov, cov_sd, start, endp, mean_mapq);
        CNVB_LAUNCHED(ctx, "coverage_stats_kernel");
        return 0;
    }
};

}  // namespace
