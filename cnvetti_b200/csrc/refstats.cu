// refstats.cu -- ReferenceStats::look_at_chars (lib-coverage/src/reference.rs:84-130).
//
// Per window: gc = #[GCgc] / #[ACGTacgt] (f32, missing if no ACGT), has_gap = any [Nn].
// HBM-streaming byte kernel, algorithmic traffic 1 B/base in, 5 B/window out.
//
// A byte-at-a-time (or SWAR) classifier is instruction-issue bound long before HBM is busy
// (round-1 ncu: 0.8 warp instructions per byte, 12 % DRAM throughput).  The kernels classify TWO
// bytes per shared-memory lookup: a 65536-entry u16 table (128 KB, built once per context, copied by
// every CTA; one persistent 1024-thread CTA per SM) maps a byte pair to its packed class counts
//   bits 0..6 #[GCgc]   bits 7..13 #[ACGTacgt]   bits 14.. #[Nn]   (entries hold 0..2 per field)
// with   [CGcg] <=> (b & 0xDB) == 0x43    [Aa] <=> (b & 0xDF) == 0x41
//        [Tt]   <=> (b & 0xDF) == 0x54    [Nn] <=> (b & 0xDF) == 0x4E
// so a lane turns 16 bytes (one 128-bit load) into counts with 8 LDS.U16 + 8 adds.
//   wl >= 128: refstats_lane_kernel -- G lanes walk one window's chunks in turn (see there);
//   shorter windows / sequences >= 4 GB: refstats_lut_kernel -- a group of G lanes (16 G >= wl + 30)
//   reads the aligned chunks overlapping the window once, masking the bytes of the neighbours.
// No counters in HBM, no atomics, no memset, no finalize pass, bit-reproducible.
#include "common.cuh"

using namespace cnvb;

namespace {

__device__ __forceinline__ uint32_t byte_mask_below(int k) {  // bytes with index < k (k in 0..4)
    return k >= 4 ? 0xFFFFFFFFu : ((1u << (8 * k)) - 1u);
}

struct RefArgs {
    const uint8_t *seq;
    uint64_t len;
    uint32_t wl;
};

__device__ __forceinline__ uint32_t ref_class(uint32_t b) {
    const uint32_t isgc = (b & 0xDBu) == 0x43u, x = b & 0xDFu;
    const uint32_t isacgt = isgc | (x == 0x41u) | (x == 0x54u);
    const uint32_t isn = x == 0x4Eu;
    return isgc | (isacgt << 7) | (isn << 14);  // 7-bit fields: a lane may add up 63 pairs before unpacking
}

__device__ __forceinline__ void ref_store(float *__restrict__ gc, uint8_t *__restrict__ has_gap, uint64_t W,
                                          uint32_t g, uint32_t c, uint32_t n) {
    // reference.rs:120-127: i32 counts -> f32 ratio, f32::missing() if the window has no ACGT
    gc[W] = c == 0 ? f32_missing() : __fdiv_rn((float)g, (float)c);
    has_gap[W] = n != 0;
}

// The two table indices of a 32-bit word (two byte pairs), in place; a bijection on each 16-bit half.
// The shared-memory bank of an entry is (index >> 1) & 31.  Genomic text uses about ten byte values, so
// indexing by (b1 << 8 | b0) would put all lookups of a warp into the five banks that b0 alone selects.
// A, C, G, T differ in bits 1 and 2; the index keeps b0's bits 1, 2 and mixes b1's bits 1, 2 into bits
// 3, 4: the 16 upper-case pairs (and the 16 lower-case ones) get 16 different banks, lanes with the same
// pair share one address (broadcast).  (b0's case bit is bank bit 4 already; mixing in b1's too costs
// two more instructions per word and only matters where a warp reads both cases at once.)
__device__ __forceinline__ uint32_t lut_swizzle(uint32_t w) { return w ^ ((w >> 6) & 0x00180018u); }
__device__ __forceinline__ uint32_t lut_index(uint32_t pair) { return lut_swizzle(pair) & 0xffffu; }

// the table is built once per context in HBM (L2-resident afterwards); every CTA copies its 128 KB
// with 16-byte loads instead of classifying 65536 pairs again at every launch
__global__ void refstats_build_lut_kernel(uint16_t *__restrict__ lut) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 65536u) lut[lut_index(i)] = (uint16_t)(ref_class(i & 0xffu) + ref_class(i >> 8));
}
// 128 KB per CTA and launch: one thread issues four 32 KB bulk async copies (TMA engine, L2 -> shared memory)
// and the CTA waits on their mbarrier -- 1024 threads fetching 8 x 16 bytes each took ~8 us of exposed
// latency per launch, a fifth of the 22-launch step.
__device__ __forceinline__ void load_lut(uint16_t *s_lut, const uint16_t *__restrict__ lut) {
    __shared__ __align__(8) uint64_t s_bar;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_lut);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(65536u * 2u) : "memory");
#pragma unroll
        for (uint32_t part = 0; part < 4; ++part)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             dst + part * 32768u),
                         "l"(reinterpret_cast<const char *>(lut) + part * 32768u), "r"(32768u), "r"(bar)
                         : "memory");
    }
    __syncthreads();  // the barrier is initialised before anybody polls it
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar)
            : "memory");
    } while (!ok);
}

// Lane-group variant (wl >= 128): G lanes own one window and walk its interior 16-byte chunks in turn,
// so the inner loop is nothing but load -> eight table lookups -> add.  The two edge chunks of a window
// (whose foreign bytes must be masked) are peeled out of the loop: lane 0 of the group takes the first,
// lane 1 the last, once per window.  G >= 2 so that the lanes of a group read whole 32-byte sectors;
// larger G only serves to fill the machine when a contig has few windows.
//
// (Round-1 experiments on the index arithmetic -- halves by PRMT, doubling and adds as IMAD on the FMA
// pipe, right shifts as IMAD.HI -- cut the ALU-pipe instructions by up to 30 % and left the time unchanged
// (IMAD.HI / IMAD.WIDE made it worse): after the edge peeling the kernel waits on its loads, not on a pipe.)
__device__ __forceinline__ uint32_t lut_counts(const uint16_t *s_lut, const uint32_t (&w)[4]) {
    uint32_t t = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t x = lut_swizzle(w[j]);
        t += (uint32_t)s_lut[x & 0xffffu] + (uint32_t)s_lut[x >> 16];
    }
    return t;  // 7-bit fields, at most 16 each
}

constexpr int kInFlight = 4;  // 16-byte loads per lane in flight (8 measured the same)

// counts of one window for lane `sub` of its group: interior chunks in turn, lanes 0 / 1 also an edge chunk
template <int G>
__device__ __forceinline__ void ref_window_counts(const uint16_t *s_lut, const int4 *__restrict__ base4, uint32_t os,
                                                  uint32_t oe, uint32_t sub, uint32_t &g, uint32_t &c, uint32_t &n) {
    const uint32_t cf = os >> 4, cl = (oe - 1u) >> 4;
    // edge chunks: issued first, consumed last
    const uint32_t ek = sub == 0 ? cf : cl;
    const bool has_edge = sub == 0 || (sub == 1 && cl != cf);
    int4 edge = make_int4(0, 0, 0, 0);
    if (has_edge) edge = __ldg(base4 + ek);
    // the G lanes take the interior chunks of the window in turn (together they read G x 16 contiguous
    // bytes); kInFlight loads per lane are in flight before the first is consumed -- with one, the 1024
    // threads of an SM keep only 16 KB in flight, less than half of what HBM latency needs
    for (uint32_t k = cf + 1u + sub; k < cl; k += kInFlight * G) {
        int4 raw[kInFlight];
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) {
            const uint32_t kk = k + u * G;
            raw[u] = kk < cl ? __ldg(base4 + kk) : make_int4(0, 0, 0, 0);  // (0x00 is in no class)
        }
        uint32_t acc = 0;  // packed 7-bit fields: kInFlight x 16 <= 127
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) {
            const uint32_t w[4] = {(uint32_t)raw[u].x, (uint32_t)raw[u].y, (uint32_t)raw[u].z, (uint32_t)raw[u].w};
            acc += lut_counts(s_lut, w);
        }
        g += acc & 127u;
        c += (acc >> 7) & 127u;
        n += acc >> 14;
    }
    if (has_edge) {  // zero the bytes of the neighbouring windows
        uint32_t w[4] = {(uint32_t)edge.x, (uint32_t)edge.y, (uint32_t)edge.z, (uint32_t)edge.w};
        const uint32_t c0 = ek << 4;
        const int lo = os > c0 ? (int)(os - c0) : 0, hi = oe < c0 + 16u ? (int)(oe - c0) : 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = min(max(lo - 4 * j, 0), 4), y = min(max(hi - 4 * j, 0), 4);
            w[j] &= byte_mask_below(y) & ~byte_mask_below(x);
        }
        const uint32_t acc = lut_counts(s_lut, w);
        g += acc & 127u;
        c += (acc >> 7) & 127u;
        n += acc >> 14;
    }
}

template <int G>
__global__ void __launch_bounds__(1024, 1) refstats_lane_kernel(RefArgs a, uint32_t num_bins,
                                                                const uint16_t *__restrict__ lut,
                                                                float *__restrict__ gc, uint8_t *__restrict__ has_gap) {
    extern __shared__ __align__(16) uint16_t s_lut[];  // [65536]
    load_lut(s_lut, lut);
    const uint32_t lane = lane_id(), sub = lane % G;
    const uint32_t n_groups = gridDim.x * (blockDim.x / G);
    const uintptr_t addr = reinterpret_cast<uintptr_t>(a.seq);
    const uint32_t mis = (uint32_t)(addr & 15u);
    const int4 *__restrict__ base4 = reinterpret_cast<const int4 *>(addr - mis);  // 16-byte aligned
    const uint32_t wl = a.wl, end_off = (uint32_t)a.len + mis;  // (the host guarantees len + 16 < 2^32)
    // warp-uniform trip count so that the group shuffles below stay converged
    const uint32_t first = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const uint32_t warp_first = (blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) / G;
    for (uint32_t W0 = warp_first, W = first; W0 < num_bins; W0 += n_groups, W += n_groups) {
        uint32_t g = 0, c = 0, n = 0;
        if (W < num_bins) {
            const uint32_t os = W * wl + mis, oe = min(os + wl, end_off);
            ref_window_counts<G>(s_lut, base4, os, oe, sub, g, c, n);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            g += __shfl_xor_sync(0xffffffffu, g, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
            n += __shfl_xor_sync(0xffffffffu, n, o);
        }
        if (W < num_bins && sub == 0) ref_store(gc, has_gap, W, g, c, n);
    }
}

// The same for all regions of a run in ONE launch.  A launch of this kernel costs ~12 us before the first
// byte is counted (1024-thread CTAs with 128 KB of shared memory: the SMs switch their L1 / shared-memory
// split against the binning kernels'), which over the 22 contig launches of a genome was a third of the
// reference-statistics time.  Windows are numbered through all regions; a lane group finds its region by
// binary search in a shared-memory copy of the region table.
struct RefRegion {
    const uint8_t *seq;
    uint32_t len;
    uint32_t first;    // number of windows in the regions before this one
    uint64_t out_off;  // row of the region's first window in the output columns
};
constexpr uint32_t kMaxBatchRegions = 512;

template <int G>
__global__ void __launch_bounds__(1024, 1) refstats_batch_kernel(const RefRegion *__restrict__ regs, uint32_t n_regs,
                                                                 uint32_t wl, uint32_t total_bins,
                                                                 const uint16_t *__restrict__ lut, float *__restrict__ gc,
                                                                 uint8_t *__restrict__ has_gap) {
    extern __shared__ __align__(16) uint16_t s_lut[];  // [65536]
    __shared__ RefRegion s_regs[kMaxBatchRegions];
    for (uint32_t i = threadIdx.x; i < n_regs; i += blockDim.x) s_regs[i] = regs[i];
    load_lut(s_lut, lut);  // (its __syncthreads covers the region table too)
    const uint32_t lane = lane_id(), sub = lane % G;
    const uint32_t n_groups = gridDim.x * (blockDim.x / G);
    const uint32_t first = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const uint32_t warp_first = (blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) / G;
    for (uint32_t W0 = warp_first, W = first; W0 < total_bins; W0 += n_groups, W += n_groups) {
        uint32_t g = 0, c = 0, n = 0;
        uint64_t out = 0;
        if (W < total_bins) {
            uint32_t lo = 0, hi = n_regs;  // last region with first <= W
            while (hi - lo > 1u) {
                const uint32_t mid = (lo + hi) >> 1;
                if (s_regs[mid].first <= W) lo = mid; else hi = mid;
            }
            const RefRegion r = s_regs[lo];
            const uint32_t Wl = W - r.first;
            out = r.out_off + Wl;
            const uintptr_t addr = reinterpret_cast<uintptr_t>(r.seq);
            const uint32_t mis = (uint32_t)(addr & 15u);
            const int4 *__restrict__ base4 = reinterpret_cast<const int4 *>(addr - mis);
            const uint32_t os = Wl * wl + mis, oe = min(os + wl, r.len + mis);
            ref_window_counts<G>(s_lut, base4, os, oe, sub, g, c, n);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            g += __shfl_xor_sync(0xffffffffu, g, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
            n += __shfl_xor_sync(0xffffffffu, n, o);
        }
        if (W < total_bins && sub == 0) ref_store(gc, has_gap, out, g, c, n);
    }
}

struct RefWin {  // one window as seen by a lane group
    uintptr_t as, ae, a0;  // absolute byte addresses: window start / end, first aligned chunk
    uint32_t nchunks;      // aligned 16-byte chunks overlapping the window (0 = idle group)
    uint64_t W;
};

__device__ __forceinline__ RefWin ref_window(const RefArgs &a, uint64_t W, uint64_t num_bins) {
    RefWin r;
    r.W = W;
    const uintptr_t base = reinterpret_cast<uintptr_t>(a.seq);
    const uint64_t s = W * a.wl;
    const uint64_t e = s + a.wl < a.len ? s + a.wl : a.len;
    r.as = base + s;
    r.ae = base + e;
    r.a0 = r.as & ~(uintptr_t)15;
    r.nchunks = W < num_bins ? (uint32_t)((r.ae - r.a0 + 15) >> 4) : 0u;
    return r;
}

// General variant: a group of G lanes per window (short windows, or sequences >= 4 GB).  Chunks
// that stick out of the window are masked in registers.
template <int G>
__global__ void __launch_bounds__(1024, 1) refstats_lut_kernel(RefArgs a, uint64_t num_bins,
                                                               const uint16_t *__restrict__ lut,
                                                               float *__restrict__ gc,
                                                               uint8_t *__restrict__ has_gap) {
    extern __shared__ __align__(16) uint16_t s_lut[];  // [65536]
    load_lut(s_lut, lut);
    constexpr uint32_t kGroups = 32 / G;
    const uint32_t lane = lane_id(), sub = lane % G, grp = lane / G;
    const uint64_t warp_global = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t step = (uint64_t)gridDim.x * (blockDim.x >> 5) * kGroups;
    for (uint64_t W0 = warp_global * kGroups; W0 < num_bins; W0 += step) {
        const RefWin cur = ref_window(a, W0 + grp, num_bins);
        uint32_t g = 0, c = 0, n = 0;
        for (uint32_t k = sub; k < cur.nchunks; k += G) {
            const uintptr_t ca = cur.a0 + ((uintptr_t)k << 4);
            const int4 raw = ld_stream(reinterpret_cast<const int4 *>(ca));
            uint32_t w[4] = {(uint32_t)raw.x, (uint32_t)raw.y, (uint32_t)raw.z, (uint32_t)raw.w};
            if (ca < cur.as || ca + 16 > cur.ae) {  // first / last chunk: zero the foreign bytes
                const int lo = ca < cur.as ? (int)(cur.as - ca) : 0;
                const int hi = ca + 16 > cur.ae ? (int)(cur.ae - ca) : 16;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int x = min(max(lo - 4 * j, 0), 4), y = min(max(hi - 4 * j, 0), 4);
                    w[j] &= byte_mask_below(y) & ~byte_mask_below(x);
                }
            }
            uint32_t sum = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t x = lut_swizzle(w[j]);
                sum += (uint32_t)s_lut[x & 0xffffu] + (uint32_t)s_lut[x >> 16];
            }
            g += sum & 127u;
            c += (sum >> 7) & 127u;
            n += sum >> 14;
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            g += __shfl_xor_sync(0xffffffffu, g, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
            n += __shfl_xor_sync(0xffffffffu, n, o);
        }
        if (cur.nchunks && sub == 0) ref_store(gc, has_gap, cur.W, g, c, n);
    }
}

}  // namespace

int cnvb::refstats_prepare(cnvb_ctx *ctx) {
    if (ctx->ref_lut) return 0;
    CNVB_CUDA(ctx, cudaMalloc(&ctx->ref_lut, 65536 * sizeof(uint16_t)), "allocating the classification table");
    refstats_build_lut_kernel<<<256, 256, 0, ctx->stream>>>(static_cast<uint16_t *>(ctx->ref_lut));
    CNVB_LAUNCHED(ctx, "refstats_build_lut_kernel");
    return 0;
}

extern "C" int cnvb_refstats(cnvb_ctx *ctx, const uint8_t *seq, uint64_t len, uint32_t window_length,
                             float *gc, uint8_t *has_gap, int mem) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (window_length == 0) return fail(ctx, CNVB_ERR_INVALID, "window length must be positive");
    if (len && (!seq || !gc || !has_gap)) return fail(ctx, CNVB_ERR_INVALID, "cnvb_refstats: null argument");
    const uint64_t nb = (len + window_length - 1) / window_length;  // reference.rs:91
    if (nb == 0) return CNVB_OK;
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    float *d_gc = gc;
    uint8_t *d_gap = has_gap;
    const uint8_t *d_seq = seq;
    if (mem != CNVB_MEM_DEVICE) {
        CNVB_TRY(ensure_scratch(ctx, Carver::need({nb * 4, nb})));
        Carver c(ctx->scratch);
        d_gc = c.take<float>(nb);
        d_gap = c.take<uint8_t>(nb);
        CNVB_TRY(ensure_stage(ctx, len + 16));
        CNVB_CUDA(ctx, cudaMemcpyAsync(ctx->stage, seq, len, cudaMemcpyHostToDevice, ctx->stream),
                  "host-to-device copy of the reference sequence");
        d_seq = static_cast<const uint8_t *>(ctx->stage);
    }
    CNVB_TRY(refstats_prepare(ctx));
    const uint16_t *d_lut = static_cast<const uint16_t *>(ctx->ref_lut);
    RefArgs a{};
    a.seq = d_seq;
    a.len = len;
    a.wl = window_length;
    {
        // lane group size: 16*G bytes must cover a window plus up to 15+15 bytes of misalignment
        int G = 2;
        while (G < 32 && (uint64_t)16 * G < (uint64_t)window_length + 30) G <<= 1;
        const size_t smem = 65536 * sizeof(uint16_t);
        const uint64_t groups = 32 / G;
        uint64_t blocks = (nb + 32 * groups - 1) / (32 * groups);
        if (blocks > (uint64_t)ctx->sm_count) blocks = ctx->sm_count;  // persistent: one CTA per SM
        KernelSpan span(ctx, "refstats_lut_kernel");
        if (window_length >= 128 && len < 0xFFFFFF00ull) {
            // lanes per window: 1 when the contig has enough windows to fill the machine, more otherwise
            // (every lane keeps at least four chunks)
            int GL = 2;
            while (GL < 32 && nb * (uint64_t)GL < (uint64_t)ctx->sm_count * 1024 && (uint64_t)window_length / (2 * GL) >= 32) GL <<= 1;
            blocks = (nb * GL + 1023) / 1024;
            if (blocks > (uint64_t)ctx->sm_count) blocks = ctx->sm_count;
#define CNVB_REF_LANE(GG)                                                                                              \
    case GG:                                                                                                           \
        CNVB_CUDA(ctx, cudaFuncSetAttribute(refstats_lane_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                            (int)smem), "smem opt-in");                                                \
        refstats_lane_kernel<GG><<<(unsigned)blocks, 1024, smem, ctx->stream>>>(a, (uint32_t)nb, d_lut, d_gc, d_gap);   \
        break;
            switch (GL) {
                CNVB_REF_LANE(2)
                CNVB_REF_LANE(4)
                CNVB_REF_LANE(8)
                CNVB_REF_LANE(16)
                CNVB_REF_LANE(32)
            }
#undef CNVB_REF_LANE
        } else {
#define CNVB_REF_LAUNCH(GG)                                                                          \
    case GG:                                                                                         \
        CNVB_CUDA(ctx, cudaFuncSetAttribute(refstats_lut_kernel<GG>,                                 \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                  "smem opt-in");                                                                    \
        refstats_lut_kernel<GG><<<(unsigned)blocks, 1024, smem, ctx->stream>>>(a, nb, d_lut, d_gc, d_gap);  \
        break;
            switch (G) {
                CNVB_REF_LAUNCH(2)
                CNVB_REF_LAUNCH(4)
                CNVB_REF_LAUNCH(8)
                CNVB_REF_LAUNCH(16)
                CNVB_REF_LAUNCH(32)
            }
#undef CNVB_REF_LAUNCH
        }
    }
    CNVB_LAUNCHED(ctx, "refstats_lut_kernel");
    if (mem != CNVB_MEM_DEVICE) {
        CNVB_CUDA(ctx, cudaMemcpyAsync(gc, d_gc, nb * 4, cudaMemcpyDeviceToHost, ctx->stream), "copy gc");
        CNVB_CUDA(ctx, cudaMemcpyAsync(has_gap, d_gap, nb, cudaMemcpyDeviceToHost, ctx->stream), "copy has_gap");
        CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "refstats");
    }
    return CNVB_OK;
}


// All regions of a run in one launch (device-resident sequences; window length >= 128, each region below
// 4 GB).  out_off[i] = row of region i's first window in gc / has_gap.  Returns 1 if the batch form does
// not apply (the caller then goes region by region), 0 on success, a negative error otherwise.
int cnvb::refstats_batch_device(cnvb_ctx *ctx, uint32_t n, const uint8_t *const *seq, const uint64_t *len,
                                const uint64_t *out_off, uint32_t window_length, float *gc, uint8_t *has_gap,
                                bool shares_sms, std::vector<void *> *defer_release) {
    if (n == 0) return 0;
    if (window_length < 128 || n > kMaxBatchRegions) return 1;
    std::vector<RefRegion> regs(n);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (len[i] >= 0xFFFFFF00ull) return 1;
        regs[i].seq = seq[i];
        regs[i].len = (uint32_t)len[i];
        regs[i].first = (uint32_t)total;
        regs[i].out_off = out_off[i];
        total += (len[i] + window_length - 1) / window_length;
    }
    if (total == 0) return 0;
    if (total >= 0xFFFFFF00ull) return 1;
    CNVB_TRY(refstats_prepare(ctx));
    RefRegion *d_regs = nullptr;
    cudaError_t e = pool_alloc_t(ctx, (size_t)n, &d_regs);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "allocating the region table");
    // (pageable source: the copy is staged by the runtime before the call returns, so `regs` may go; under a
    // plan's capture the copy is a graph node that reads its source at every replay: pinned, owned by the plan)
    const void *src = regs.data();
    if (ctx->capturing) {
        src = capture_pin(ctx, regs.data(), n * sizeof(RefRegion));
        if (!src) {
            pool_release(ctx, d_regs);
            return fail(ctx, CNVB_ERR_NOMEM, "pinned region table");
        }
    }
    e = cudaMemcpyAsync(d_regs, src, n * sizeof(RefRegion), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) {
        pool_release(ctx, d_regs);
        return fail_cuda(ctx, e, "uploading the region table");
    }
    const size_t smem = 65536 * sizeof(uint16_t);
    int GL = 2;
    while (GL < 32 && total * (uint64_t)GL < (uint64_t)ctx->sm_count * 1024 && (uint64_t)window_length / (2 * GL) >= 32) GL <<= 1;
    // Threads per CTA (one CTA per SM: the table takes 128 KB).  Alone, 1024 threads are fastest (0.60 ms per
    // config-2 step against 0.70 with 512).  Beside the binning kernel (the region loop's side streams), 1024
    // threads x 56 registers leave no room for a binning CTA on the SM and the two kernels merely take
    // turns; with 512 two binning CTAs fit and the ALU-bound table lookups run under the HBM-bound binning:
    // 1.82 -> 1.72 ms per step (768: 1.79, 256: 1.81).  CNVB_REF_THREADS overrides.
    static const unsigned ref_threads_env = [] {
        const char *e = getenv("CNVB_REF_THREADS");
        const int v = e ? atoi(e) : 0;
        return v == 256 || v == 512 || v == 768 || v == 1024 ? (unsigned)v : 0u;
    }();
    const unsigned ref_threads = ref_threads_env ? ref_threads_env : (shares_sms ? 512u : 1024u);
    uint64_t blocks = (total * GL + ref_threads - 1) / ref_threads;
    if (blocks > (uint64_t)ctx->sm_count) blocks = ctx->sm_count;
    const uint16_t *d_lut = static_cast<const uint16_t *>(ctx->ref_lut);
    {
        KernelSpan span(ctx, "refstats_lut_kernel");
#define CNVB_REF_BATCH(GG)                                                                                             \
    case GG:                                                                                                           \
        e = cudaFuncSetAttribute(refstats_batch_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        if (e == cudaSuccess)                                                                                          \
            refstats_batch_kernel<GG><<<(unsigned)blocks, ref_threads, smem, ctx->stream>>>(d_regs, n, window_length,   \
                                                                                     (uint32_t)total, d_lut, gc, has_gap); \
        break;
        switch (GL) {
            CNVB_REF_BATCH(2)
            CNVB_REF_BATCH(4)
            CNVB_REF_BATCH(8)
            CNVB_REF_BATCH(16)
            CNVB_REF_BATCH(32)
        }
#undef CNVB_REF_BATCH
    }
    // The region loop runs this launch on a side stream of its own: the table then stays with the caller
    // until the streams have joined, so that no aggregator of another stream is handed the block (the
    // pool would order that stream behind this whole launch).  Otherwise: stream-ordered reuse.
    if (defer_release)
        defer_release->push_back(d_regs);
    else
        pool_release(ctx, d_regs);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "smem opt-in");
    CNVB_LAUNCHED(ctx, "refstats_lut_kernel");
    return 0;
}
