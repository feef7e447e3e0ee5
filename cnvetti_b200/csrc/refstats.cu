// refstats.cu -- ReferenceStats::look_at_chars (lib-coverage/src/reference.rs:84-130).
//
// Per window: gc = #[GCgc] / #[ACGTacgt] (f32, missing if no ACGT), has_gap = any [Nn].
// HBM-streaming byte kernel, algorithmic traffic 1 B/base in, 5 B/window out.
//
// A byte-at-a-time (or SWAR) classifier is instruction-issue bound long before HBM is busy
// (round-1 ncu: 0.8 warp instructions per byte, 12 % DRAM throughput).  This kernel classifies
// TWO bytes per shared-memory lookup: a 65536-entry u16 table (128 KB, built in the CTA's shared
// memory at kernel start; one persistent 1024-thread CTA per SM) maps a byte pair to its packed
// class counts
//   bits 0..4 #[GCgc]   bits 5..9 #[ACGTacgt]   bits 10..14 #[Nn]
// with   [CGcg] <=> (b & 0xDB) == 0x43    [Aa] <=> (b & 0xDF) == 0x41
//        [Tt]   <=> (b & 0xDF) == 0x54    [Nn] <=> (b & 0xDF) == 0x4E
// so a lane turns its 16 bytes (one 128-bit load) into counts with 8 LDS.U16 + 7 adds.  A group
// of G lanes (power of two, 16*G >= wl + 30) owns one window per pass: the lanes read the
// aligned 16-byte chunks overlapping the window (coalesced), bytes outside the window are zeroed
// (0x00 belongs to no class), the group reduces with REDUX / shuffles and its first lane writes
// gc and has_gap directly -- no counters in HBM, no atomics, no memset, no finalize pass,
// bit-reproducible.
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
    return isgc | (isacgt << 5) | (isn << 10);
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

template <int G>
__global__ void __launch_bounds__(1024, 1) refstats_lut_kernel(RefArgs a, uint64_t num_bins,
                                                               float *__restrict__ gc,
                                                               uint8_t *__restrict__ has_gap) {
    extern __shared__ uint16_t s_lut[];  // [65536]
    for (uint32_t i = threadIdx.x; i < 65536u; i += blockDim.x)
        s_lut[i] = (uint16_t)(ref_class(i & 0xffu) + ref_class(i >> 8));
    __syncthreads();
    constexpr uint32_t kGroups = 32 / G;
    const uint32_t lane = lane_id(), sub = lane % G, grp = lane / G;
    const uint64_t warp_global = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t step = (uint64_t)gridDim.x * (blockDim.x >> 5) * kGroups;
    uint64_t W0 = warp_global * kGroups;
    // software pipeline: the first chunk of the next window is in flight while this one is counted
    RefWin nxt = ref_window(a, W0 + grp, num_bins);
    int4 nraw = make_int4(0, 0, 0, 0);
    if (sub < nxt.nchunks) nraw = ld_stream(reinterpret_cast<const int4 *>(nxt.a0 + ((uintptr_t)sub << 4)));
    for (; W0 < num_bins; W0 += step) {
        const RefWin cur = nxt;
        int4 raw = nraw;
        nxt = ref_window(a, W0 + step + grp, num_bins);
        if (sub < nxt.nchunks) nraw = ld_stream(reinterpret_cast<const int4 *>(nxt.a0 + ((uintptr_t)sub << 4)));
        uint32_t g = 0, c = 0, n = 0;
        for (uint32_t k = sub; k < cur.nchunks; k += G) {
            const uintptr_t ca = cur.a0 + ((uintptr_t)k << 4);
            if (k != sub) raw = ld_stream(reinterpret_cast<const int4 *>(ca));
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
            for (int j = 0; j < 4; ++j) sum += (uint32_t)s_lut[w[j] & 0xffffu] + (uint32_t)s_lut[w[j] >> 16];
            g += sum & 31u;
            c += (sum >> 5) & 31u;
            n += sum >> 10;
        }
        if (G == 32) {
            g = __reduce_add_sync(0xffffffffu, g);
            c = __reduce_add_sync(0xffffffffu, c);
            n = __reduce_add_sync(0xffffffffu, n);
        } else {
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                g += __shfl_xor_sync(0xffffffffu, g, o);
                c += __shfl_xor_sync(0xffffffffu, c, o);
                n += __shfl_xor_sync(0xffffffffu, n, o);
            }
        }
        if (cur.nchunks && sub == 0) {
            // reference.rs:120-127: i32 counts -> f32 ratio, f32::missing() if no ACGT
            gc[cur.W] = c == 0 ? f32_missing() : __fdiv_rn((float)g, (float)c);
            has_gap[cur.W] = n != 0;
        }
    }
}

}  // namespace

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
#define CNVB_REF_LAUNCH(GG)                                                                          \
    case GG:                                                                                         \
        CNVB_CUDA(ctx, cudaFuncSetAttribute(refstats_lut_kernel<GG>,                                 \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                  "smem opt-in");                                                                    \
        refstats_lut_kernel<GG><<<(unsigned)blocks, 1024, smem, ctx->stream>>>(a, nb, d_gc, d_gap);  \
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
    CNVB_LAUNCHED(ctx, "refstats_lut_kernel");
    if (mem != CNVB_MEM_DEVICE) {
        CNVB_CUDA(ctx, cudaMemcpyAsync(gc, d_gc, nb * 4, cudaMemcpyDeviceToHost, ctx->stream), "copy gc");
        CNVB_CUDA(ctx, cudaMemcpyAsync(has_gap, d_gap, nb, cudaMemcpyDeviceToHost, ctx->stream), "copy has_gap");
        CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "refstats");
    }
    return CNVB_OK;
}
