// coverage_tgt.cuh -- FragmentsTargetRegionsAggregator (lib-coverage/src/agg.rs:261-444).
// Included by coverage.cu only.
//
// Per passing read: fragment interval [begin, end) = [pos, frag_end), centre (begin+end)/2 (u32),
// among targets intersecting the interval choose the one with the smallest compute_dist
// (agg.rs:350-361); strict `<` in ascending target index => lowest index wins ties (the
// reference's tie order is bio::IntervalTree's iteration order, which no test pins).
//
// Targets are sorted by start, so the intersecting set lies in [lo, ub):
//   ub = first target with start >= end           (binary search on starts)
//   lo = first target with running-max-end > begin (binary search on the prefix maximum of ends)
// For disjoint targets that range IS the intersecting set (1-3 targets for exome data).
#pragma once

namespace {

struct TgtArgs {
    const int32_t *pos;
    const int32_t *frag_end;
    const uint8_t *mapq;
    const uint16_t *flags;
    uint64_t n;
    const uint32_t *ts, *te, *pmax;
    uint32_t n_tgt;
    uint32_t min_mapq;
    uint32_t *counters;
    uint32_t *stats;
};

__device__ __forceinline__ uint32_t tgt_dist(uint32_t s, uint32_t e, uint32_t pt) {
    if (pt < s) return s - pt;
    if (pt >= e) return pt - e + 1u;
    return 0u;
}

// target table loads: read-only global path, or plain loads when the table sits in shared memory
template <bool SM>
__device__ __forceinline__ uint32_t ldt(const uint32_t *p) {
    if (SM) return *p;
    return __ldg(p);
}

// returns target index or 0xffffffff
template <bool SM = false>
__device__ __forceinline__ uint32_t tgt_best(const TgtArgs &a, uint32_t begin, uint32_t end) {
    const uint32_t centre = (begin + end) >> 1;  // agg.rs:321 (u32)
    uint32_t lo = 0, hi = a.n_tgt;
    while (lo < hi) {  // first t with pmax[t] > begin
        const uint32_t m = (lo + hi) >> 1;
        if (ldt<SM>(a.pmax + m) > begin) hi = m; else lo = m + 1;
    }
    uint32_t ub = lo;
    hi = a.n_tgt;
    while (ub < hi) {  // first t with ts[t] >= end
        const uint32_t m = (ub + hi) >> 1;
        if (ldt<SM>(a.ts + m) >= end) hi = m; else ub = m + 1;
    }
    uint32_t best = 0xffffffffu, best_d = 0xffffffffu;
    for (uint32_t t = lo; t < ub; ++t) {
        const uint32_t s = ldt<SM>(a.ts + t), e = ldt<SM>(a.te + t);
        if (!(begin < e)) continue;  // (s < end holds for every t < ub)
        const uint32_t d = tgt_dist(s, e, centre);
        if (best == 0xffffffffu || d < best_d) {
            best = t;
            best_d = d;
            if (d == 0u) break;  // nothing later can be strictly smaller
        }
    }
    return best;
}

// First index of a sorted array with arr[i] > key (GT) or arr[i] >= key (!GT), found by the whole
// warp: 32 probes per step, three steps cover 32768 targets.  `key` must be warp-uniform.
template <bool GT, bool SM = false>
__device__ __forceinline__ uint32_t warp_first(const uint32_t *__restrict__ arr, uint32_t n, uint32_t key) {
    const uint32_t lane = lane_id();
    uint32_t lo = 0, hi = n;  // the answer lies in [lo, hi]
    while (hi - lo > 32u) {
        const uint32_t step = (hi - lo + 31u) / 32u;
        const uint32_t idx = lo + (lane + 1u) * step - 1u;
        bool p = true;
        if (idx < hi) {
            const uint32_t v = ldt<SM>(arr + idx);
            p = GT ? v > key : v >= key;
        }
        const unsigned b = __ballot_sync(0xffffffffu, p);
        if (b == 0u) {
            lo = hi;
            break;
        }
        const uint32_t f = __ffs(b) - 1u;
        hi = min(hi, lo + (f + 1u) * step - 1u);
        lo = lo + f * step;
    }
    const uint32_t idx = lo + lane;
    bool p = true;
    if (idx < hi) {
        const uint32_t v = ldt<SM>(arr + idx);
        p = GT ? v > key : v >= key;
    }
    const unsigned b = __ballot_sync(0xffffffffu, p);
    return b ? lo + (uint32_t)__ffs(b) - 1u : hi;
}

// The same with a starting guess: coordinate-sorted reads handed to a warp in consecutive steps move the
// answer by a few targets at most, so one probe round over [hint - 1, hint + 31) settles it (lane 0 checks
// that the answer is not below the hint); anything else falls back to the full search.
template <bool GT, bool SM>
__device__ __forceinline__ uint32_t warp_first_hint(const uint32_t *__restrict__ arr, uint32_t n, uint32_t key, uint32_t hint) {
    const uint32_t lane = lane_id();
    if (hint > n) hint = n;
    const uint32_t idx = hint + lane - 1u;  // lane 0: hint - 1
    bool p = lane != 0u;                    // beyond the end counts as "true"; hint == 0: nothing below
    if ((lane != 0u || hint != 0u) && idx < n) {
        const uint32_t v = ldt<SM>(arr + idx);
        p = GT ? v > key : v >= key;
    }
    const unsigned b = __ballot_sync(0xffffffffu, p);
    if ((b & 1u) == 0u && (b >> 1) != 0u) return hint + (uint32_t)__ffs(b >> 1) - 1u;
    return warp_first<GT, SM>(arr, n, key);
}

// the scan of tgt_best over a candidate range that is a superset of the intersecting targets
template <bool SM = false>
__device__ __forceinline__ uint32_t tgt_best_in(const TgtArgs &a, uint32_t begin, uint32_t end, uint32_t lo, uint32_t ub) {
    const uint32_t centre = (begin + end) >> 1;  // agg.rs:321 (u32)
    uint32_t best = 0xffffffffu, best_d = 0xffffffffu;
    for (uint32_t t = lo; t < ub; ++t) {
        const uint32_t s = ldt<SM>(a.ts + t), e = ldt<SM>(a.te + t);
        if (!(begin < e) || !(s < end)) continue;  // does not intersect [begin, end)
        const uint32_t d = tgt_dist(s, e, centre);
        if (best == 0xffffffffu || d < best_d) {
            best = t;
            best_d = d;
            if (d == 0u) break;  // nothing later can be strictly smaller
        }
    }
    return best;
}

// SM: the target table of the region (running max of ends | starts | ends, 12 B per target) is staged in
// shared memory by every CTA -- the two cooperative searches and the candidate scans of a warp step are a
// chain of ~8 dependent table reads, ~30 cycles each from shared memory against ~250 from L2.  The read
// columns of the NEXT step are requested before the current one is processed (VEC), so the trip to HBM
// overlaps the searches.
template <bool VEC, bool SM>
__global__ void __launch_bounds__(256, SM ? 3 : 4) frag_bin_tgt_kernel(TgtArgs a) {
    constexpr int E = VEC ? 4 : 1;
    extern __shared__ __align__(16) uint32_t s_tab[];
    if (SM) {
        const uint32_t n = a.n_tgt;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            s_tab[i] = __ldg(a.pmax + i);
            s_tab[n + i] = __ldg(a.ts + i);
            s_tab[2 * n + i] = __ldg(a.te + i);
        }
        __syncthreads();
        a.pmax = s_tab;
        a.ts = s_tab + n;
        a.te = s_tab + 2 * n;
    }
    const uint64_t nitems = VEC ? (a.n >> 2) : a.n;
    // every warp takes a CONTIGUOUS run of steps (32 items each): the target range of one step is the
    // starting guess of the next
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_steps = (nitems + 31) >> 5, per_warp = (n_steps + n_warps - 1) / n_warps;
    const uint64_t warp_id = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t stride = 32;
    const uint64_t first = warp_id * per_warp * 32 + lane_id();
    const uint64_t w_end = min(nitems, (warp_id + 1) * per_warp * 32);
    uint32_t n_proc = 0, n_skip = 0, hint_lo = 0, hint_ub = 0;
    int4 np = make_int4(0, 0, 0, 0), nfe = make_int4(0, 0, 0, 0);
    uint32_t nm = 0;
    uint2 nf = make_uint2(0, 0);
    if (VEC && first < w_end) {
        np = ld_stream(reinterpret_cast<const int4 *>(a.pos) + first);
        nfe = ld_stream(reinterpret_cast<const int4 *>(a.frag_end) + first);
        nm = ld_stream(reinterpret_cast<const uint32_t *>(a.mapq) + first);
        nf = ld_stream(reinterpret_cast<const uint2 *>(a.flags) + first);
    }
    for (uint64_t qb = first - lane_id(); qb < w_end; qb += stride) {
        const uint64_t q = qb + lane_id();
        const bool valid = q < w_end;
        uint32_t begin[E], fend[E], mq[E], fl[E];
        if (VEC) {
            const int4 p = np, fe = nfe;
            const uint32_t m = nm;
            const uint2 f = nf;
            if (q + stride < w_end) {  // the next step's columns
                np = ld_stream(reinterpret_cast<const int4 *>(a.pos) + q + stride);
                nfe = ld_stream(reinterpret_cast<const int4 *>(a.frag_end) + q + stride);
                nm = ld_stream(reinterpret_cast<const uint32_t *>(a.mapq) + q + stride);
                nf = ld_stream(reinterpret_cast<const uint2 *>(a.flags) + q + stride);
            }
            const uint32_t pp[4] = {(uint32_t)p.x, (uint32_t)p.y, (uint32_t)p.z, (uint32_t)p.w};
            const uint32_t ee[4] = {(uint32_t)fe.x, (uint32_t)fe.y, (uint32_t)fe.z, (uint32_t)fe.w};
            const uint32_t ff[4] = {f.x & 0xffffu, f.x >> 16, f.y & 0xffffu, f.y >> 16};
#pragma unroll
            for (int e = 0; e < E; ++e) {
                begin[e] = pp[e];
                fend[e] = ee[e];
                mq[e] = (m >> (8 * e)) & 0xffu;
                fl[e] = ff[e];
            }
        } else {
            begin[0] = valid ? (uint32_t)a.pos[q] : 0u;
            fend[0] = valid ? (uint32_t)a.frag_end[q] : 0u;
            mq[0] = valid ? a.mapq[q] : 0u;
            fl[0] = valid ? a.flags[q] : 0u;
        }
        // Coordinate-sorted reads: the 32 E fragments of this warp step lie within a few hundred
        // bases, so ONE cooperative search per bound (over the smallest begin / the largest end of
        // the warp) yields a candidate range that holds every fragment's intersecting targets --
        // usually one to four of them -- instead of two 12-step binary searches per read.
        bool pass[E];
        uint32_t bmin = 0xffffffffu, emax = 0u;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            pass[e] = valid && frag_pass(fl[e], mq[e], a.min_mapq);
            if (pass[e]) {
                bmin = min(bmin, begin[e]);
                emax = max(emax, fend[e]);
            }
        }
        bmin = __reduce_min_sync(0xffffffffu, bmin);
        emax = __reduce_max_sync(0xffffffffu, emax);
        if (emax == 0u && bmin == 0xffffffffu) continue;  // no fragment passes the filter (warp-uniform)
        const uint32_t lo_w = warp_first_hint<true, SM>(a.pmax, a.n_tgt, bmin, hint_lo);   // first target whose running max end > begin
        const uint32_t ub_w = warp_first_hint<false, SM>(a.ts, a.n_tgt, emax, hint_ub);    // first target starting at or after the end
        hint_lo = lo_w;
        hint_ub = ub_w;
        const bool narrow = ub_w <= lo_w + 16u;  // otherwise (unsorted input, giant fragments): per-read search
#pragma unroll
        for (int e = 0; e < E; ++e) {
            uint32_t best = 0xffffffffu, cnt = 0;
            if (pass[e]) {
                n_proc += 1;
                best = narrow ? tgt_best_in<SM>(a, begin[e], fend[e], lo_w, ub_w) : tgt_best<SM>(a, begin[e], fend[e]);
                if (best == 0xffffffffu) n_skip += 1; else cnt = 1;  // agg.rs:340-345
            }
            warp_bin_add(a.counters, best, cnt);
        }
    }
    if (VEC && blockIdx.x == 0 && threadIdx.x == 0) {
        for (uint64_t i = (a.n >> 2) << 2; i < a.n; ++i) {
            if (frag_pass(a.flags[i], a.mapq[i], a.min_mapq)) {
                n_proc += 1;
                const uint32_t best = tgt_best<SM>(a, (uint32_t)a.pos[i], (uint32_t)a.frag_end[i]);
                if (best == 0xffffffffu) n_skip += 1; else atomicAdd(&a.counters[best], 1u);
            }
        }
    }
    n_proc = __reduce_add_sync(0xffffffffu, n_proc);
    n_skip = __reduce_add_sync(0xffffffffu, n_skip);
    if (lane_id() == 0) {
        if (n_proc) atomicAdd(&a.stats[0], n_proc);
        if (n_skip) atomicAdd(&a.stats[1], n_skip);
    }
}

int put_targets(cnvb_ctx *ctx, const cnvb_read_soa *b, const uint32_t *d_ts, const uint32_t *d_te,
                const uint32_t *d_pmax, uint32_t n_tgt, uint32_t min_mapq, uint32_t *d_counters,
                uint32_t *d_stats) {
    if (!b->pos || !b->frag_end || !b->mapq || !b->flags)
        return fail(ctx, CNVB_ERR_INVALID, "TargetRegions batches need pos, frag_end, mapq and flags");
    constexpr uint64_t kChunk = 32ull << 20;
    const uint64_t chunk = b->mem == CNVB_MEM_DEVICE ? b->n : (b->n < kChunk ? b->n : kChunk);
    int32_t *s_pos = nullptr, *s_fe = nullptr;
    uint8_t *s_mq = nullptr;
    uint16_t *s_fl = nullptr;
    if (b->mem != CNVB_MEM_DEVICE) {
        CNVB_TRY(ensure_stage(ctx, Carver::need({chunk * 4, chunk * 4, chunk, chunk * 2})));
        Carver c(ctx->stage);
        s_pos = c.take<int32_t>(chunk);
        s_fe = c.take<int32_t>(chunk);
        s_mq = c.take<uint8_t>(chunk);
        s_fl = c.take<uint16_t>(chunk);
    }
    auto stage = [&](auto *dst, const auto *src, uint64_t n, auto **out) -> int {
        if (b->mem == CNVB_MEM_DEVICE) {
            *out = src;
            return 0;
        }
        CNVB_CUDA(ctx, cudaMemcpyAsync(dst, src, n * sizeof(*src), cudaMemcpyHostToDevice, ctx->stream),
                  "host-to-device copy of a read column");
        *out = dst;
        return 0;
    };
    for (uint64_t off = 0; off < b->n; off += chunk) {
        const uint64_t n = (b->n - off) < chunk ? (b->n - off) : chunk;
        TgtArgs g{};
        CNVB_TRY(stage(s_pos, b->pos + off, n, &g.pos));
        CNVB_TRY(stage(s_fe, b->frag_end + off, n, &g.frag_end));
        CNVB_TRY(stage(s_mq, b->mapq + off, n, &g.mapq));
        CNVB_TRY(stage(s_fl, b->flags + off, n, &g.flags));
        g.n = n;
        g.ts = d_ts;
        g.te = d_te;
        g.pmax = d_pmax;
        g.n_tgt = n_tgt;
        g.min_mapq = min_mapq;
        g.counters = d_counters;
        g.stats = d_stats;
        const bool aligned = ((uintptr_t)g.pos % 16 == 0) && ((uintptr_t)g.frag_end % 16 == 0) &&
                             ((uintptr_t)g.mapq % 4 == 0) && ((uintptr_t)g.flags % 8 == 0);
        const uint64_t work = aligned ? (n + 3) / 4 : n;
        const uint64_t work_blocks = (work + 255) / 256;
        // table in shared memory: up to 6144 targets (72 KB, three CTAs per SM)
        // (off unless CNVB_TGT_SMEM=1: measured on config 1, 4010 targets -- 74 against 66 us; the 48 KB copy per
        // CTA costs more than the shorter table latency returns once the searches start from a hint)
        static const bool smem_on = [] {
            const char *e = getenv("CNVB_TGT_SMEM");
            return e && atoi(e) == 1;
        }();
        const bool in_smem = smem_on && n_tgt > 0 && n_tgt <= 6144 && n >= (1u << 16);
        const size_t tab_bytes = in_smem ? (size_t)n_tgt * 12 : 0;
        {
            KernelSpan span(ctx, "frag_bin_tgt_kernel");
#define CNVB_TGT_LAUNCH(V, M)                                                                                             \
    do {                                                                                                                 \
        auto kern = frag_bin_tgt_kernel<V, M>;                                                                           \
        if (M) CNVB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 6144 * 12), "smem opt-in"); \
        kern<<<persistent_grid(ctx, kern, 256, M ? 6144 * 12 : 0, work_blocks), 256, tab_bytes, ctx->stream>>>(g);        \
    } while (0)
            if (aligned && in_smem) CNVB_TGT_LAUNCH(true, true);
            else if (aligned) CNVB_TGT_LAUNCH(true, false);
            else if (in_smem) CNVB_TGT_LAUNCH(false, true);
            else CNVB_TGT_LAUNCH(false, false);
#undef CNVB_TGT_LAUNCH
        }
        CNVB_LAUNCHED(ctx, "frag_bin_tgt_kernel");
    }
    return CNVB_OK;
}

}  // namespace
