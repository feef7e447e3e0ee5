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

// returns target index or 0xffffffff
__device__ __forceinline__ uint32_t tgt_best(const TgtArgs &a, uint32_t begin, uint32_t end) {
    const uint32_t centre = (begin + end) >> 1;  // agg.rs:321 (u32)
    uint32_t lo = 0, hi = a.n_tgt;
    while (lo < hi) {  // first t with pmax[t] > begin
        const uint32_t m = (lo + hi) >> 1;
        if (__ldg(a.pmax + m) > begin) hi = m; else lo = m + 1;
    }
    uint32_t ub = lo;
    hi = a.n_tgt;
    while (ub < hi) {  // first t with ts[t] >= end
        const uint32_t m = (ub + hi) >> 1;
        if (__ldg(a.ts + m) >= end) hi = m; else ub = m + 1;
    }
    uint32_t best = 0xffffffffu, best_d = 0xffffffffu;
    for (uint32_t t = lo; t < ub; ++t) {
        const uint32_t s = __ldg(a.ts + t), e = __ldg(a.te + t);
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
template <bool GT>
__device__ __forceinline__ uint32_t warp_first(const uint32_t *__restrict__ arr, uint32_t n, uint32_t key) {
    const uint32_t lane = lane_id();
    uint32_t lo = 0, hi = n;  // the answer lies in [lo, hi]
    while (hi - lo > 32u) {
        const uint32_t step = (hi - lo + 31u) / 32u;
        const uint32_t idx = lo + (lane + 1u) * step - 1u;
        bool p = true;
        if (idx < hi) {
            const uint32_t v = __ldg(arr + idx);
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
        const uint32_t v = __ldg(arr + idx);
        p = GT ? v > key : v >= key;
    }
    const unsigned b = __ballot_sync(0xffffffffu, p);
    return b ? lo + (uint32_t)__ffs(b) - 1u : hi;
}

// the scan of tgt_best over a candidate range that is a superset of the intersecting targets
__device__ __forceinline__ uint32_t tgt_best_in(const TgtArgs &a, uint32_t begin, uint32_t end, uint32_t lo, uint32_t ub) {
    const uint32_t centre = (begin + end) >> 1;  // agg.rs:321 (u32)
    uint32_t best = 0xffffffffu, best_d = 0xffffffffu;
    for (uint32_t t = lo; t < ub; ++t) {
        const uint32_t s = __ldg(a.ts + t), e = __ldg(a.te + t);
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

template <bool VEC>
__global__ void __launch_bounds__(256, 4) frag_bin_tgt_kernel(TgtArgs a) {
    constexpr int E = VEC ? 4 : 1;
    const uint64_t nitems = VEC ? (a.n >> 2) : a.n;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t n_proc = 0, n_skip = 0;
    const uint64_t first = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t qb = first - lane_id(); qb < nitems; qb += stride) {
        const uint64_t q = qb + lane_id();
        const bool valid = q < nitems;
        uint32_t begin[E], fend[E], mq[E], fl[E];
        if (VEC) {
            int4 p = make_int4(0, 0, 0, 0), fe = make_int4(0, 0, 0, 0);
            uint32_t m = 0;
            uint2 f = make_uint2(0, 0);
            if (valid) {
                p = ld_stream(reinterpret_cast<const int4 *>(a.pos) + q);
                fe = ld_stream(reinterpret_cast<const int4 *>(a.frag_end) + q);
                m = ld_stream(reinterpret_cast<const uint32_t *>(a.mapq) + q);
                f = ld_stream(reinterpret_cast<const uint2 *>(a.flags) + q);
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
        const uint32_t lo_w = warp_first<true>(a.pmax, a.n_tgt, bmin);    // first target whose running max end > begin
        const uint32_t ub_w = warp_first<false>(a.ts, a.n_tgt, emax);     // first target starting at or after the end
        const bool narrow = ub_w <= lo_w + 16u;  // otherwise (unsorted input, giant fragments): per-read search
#pragma unroll
        for (int e = 0; e < E; ++e) {
            uint32_t best = 0xffffffffu, cnt = 0;
            if (pass[e]) {
                n_proc += 1;
                best = narrow ? tgt_best_in(a, begin[e], fend[e], lo_w, ub_w) : tgt_best(a, begin[e], fend[e]);
                if (best == 0xffffffffu) n_skip += 1; else cnt = 1;  // agg.rs:340-345
            }
            warp_bin_add(a.counters, best, cnt);
        }
    }
    if (VEC && blockIdx.x == 0 && threadIdx.x == 0) {
        for (uint64_t i = (a.n >> 2) << 2; i < a.n; ++i) {
            if (frag_pass(a.flags[i], a.mapq[i], a.min_mapq)) {
                n_proc += 1;
                const uint32_t best = tgt_best(a, (uint32_t)a.pos[i], (uint32_t)a.frag_end[i]);
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
        {
            KernelSpan span(ctx, "frag_bin_tgt_kernel");
            if (aligned)
                frag_bin_tgt_kernel<true><<<persistent_grid(ctx, frag_bin_tgt_kernel<true>, 256, 0, work_blocks), 256, 0, ctx->stream>>>(g);
            else
                frag_bin_tgt_kernel<false><<<persistent_grid(ctx, frag_bin_tgt_kernel<false>, 256, 0, work_blocks), 256, 0, ctx->stream>>>(g);
        }
        CNVB_LAUNCHED(ctx, "frag_bin_tgt_kernel");
    }
    return CNVB_OK;
}

}  // namespace
