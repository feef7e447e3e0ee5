// refstats.cu -- ReferenceStats::look_at_chars (lib-coverage/src/reference.rs:84-130).
//
// Per window: gc = #[GCgc] / #[ACGTacgt] (f32, missing if no ACGT), has_gap = any [Nn].
// HBM-streaming byte kernel, 1 B/base: every lane loads 16 sequence bytes with one 128-bit load
// and classifies them with SWAR byte tests (no per-byte branches, no lookup table):
//   [CGcg] <=> (b & 0xDB) == 0x43      [Aa] <=> (b & 0xDF) == 0x41
//   [Tt]   <=> (b & 0xDF) == 0x54      [Nn] <=> (b & 0xDF) == 0x4E
// A lane's 16 bytes belong to at most two windows (wl >= 16); per-lane counts are combined with
// one warp REDUX per distinct window and leave the SM as one RED per (warp step, window).
#include "common.cuh"

using namespace cnvb;

namespace {

// 0x80 in every byte of y that is zero (exact, no cross-byte carries)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t y) {
    const uint32_t t = (y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | y) & 0x80808080u;
}

__device__ __forceinline__ uint32_t byte_mask_below(int k) {  // bytes with index < k (k in 0..4)
    return k >= 4 ? 0xFFFFFFFFu : ((1u << (8 * k)) - 1u);
}

// number of flagged bytes with chunk index in [lo, hi)
__device__ __forceinline__ uint32_t count_flags(const uint32_t f[4], int lo, int hi) {
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int a = min(max(lo - 4 * j, 0), 4), b = min(max(hi - 4 * j, 0), 4);
        const uint32_t m = byte_mask_below(b) & ~byte_mask_below(a);
        c |= (f[j] & m) >> (7 - j);
    }
    return __popc(c);
}
__device__ __forceinline__ uint32_t count_flags_all(const uint32_t f[4]) {
    return __popc((f[0] >> 7) | (f[1] >> 6) | (f[2] >> 5) | (f[3] >> 4));
}

struct RefArgs {
    const uint8_t *seq;  // original pointer
    uint64_t len;
    uint32_t wl;
    FastDiv div;  // only valid while len < 2^32; 64-bit division otherwise
    uint32_t *gc_cnt, *acgt_cnt, *n_cnt;
};

__device__ __forceinline__ void warp_window_add(const RefArgs &a, uint32_t win, uint32_t packed) {
    const uint32_t lane = lane_id();
    unsigned todo = __ballot_sync(0xffffffffu, packed != 0u);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const uint32_t w = __shfl_sync(0xffffffffu, win, leader);
        const unsigned same = __ballot_sync(0xffffffffu, win == w) & todo;
        const uint32_t v = ((same >> lane) & 1u) ? packed : 0u;
        const uint32_t tot = __reduce_add_sync(0xffffffffu, v);
        if ((int)lane == leader) {
            const uint32_t g = tot & 0x3ffu, c = (tot >> 10) & 0x3ffu, n = tot >> 20;
            if (g) atomicAdd(&a.gc_cnt[w], g);
            if (c) atomicAdd(&a.acgt_cnt[w], c);
            if (n) atomicAdd(&a.n_cnt[w], n);
        }
        todo &= ~same;
    }
}

// requires wl >= 16
__global__ void __launch_bounds__(256, 4) refstats_count_kernel(RefArgs a) {
    const uintptr_t addr = reinterpret_cast<uintptr_t>(a.seq);
    const uint32_t mis = (uint32_t)(addr & 15u);  // bytes before seq[0] in its 16-byte line
    const uint4 *base = reinterpret_cast<const uint4 *>(addr - mis);
    const uint64_t nchunks = (a.len + mis + 15) / 16;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t first = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t cb = first - lane_id(); cb < nchunks; cb += stride) {
        const uint64_t c = cb + lane_id();
        uint32_t winA = 0, packA = 0, winB = 0, packB = 0;
        if (c < nchunks) {
            const int4 raw = ld_stream(reinterpret_cast<const int4 *>(base + c));
            const uint32_t w[4] = {(uint32_t)raw.x, (uint32_t)raw.y, (uint32_t)raw.z, (uint32_t)raw.w};
            uint32_t fgc[4], facgt[4], fn[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t x = w[j] & 0xDFDFDFDFu;
                fgc[j] = zero_bytes((w[j] & 0xDBDBDBDBu) ^ 0x43434343u);
                facgt[j] = fgc[j] | zero_bytes(x ^ 0x41414141u) | zero_bytes(x ^ 0x54545454u);
                fn[j] = zero_bytes(x ^ 0x4E4E4E4Eu);
            }
            // sequence offsets covered by this chunk: [o, o+16) clipped to [0, len)
            const int64_t o = (int64_t)(c * 16) - mis;
            const int lo = o < 0 ? (int)(-o) : 0;
            const int hi = (uint64_t)(o + 16) > a.len ? (int)((int64_t)a.len - o) : 16;
            const uint64_t p0 = (uint64_t)(o + lo);
            const uint64_t wA = a.len <= 0xFFFFFFFFull ? fastdiv((uint32_t)p0, a.div) : p0 / a.wl;
            const uint64_t boundary = (wA + 1) * a.wl;  // first offset of the next window
            winA = (uint32_t)wA;
            winB = winA + 1;
            if (lo == 0 && hi == 16 && boundary >= (uint64_t)(o + 16)) {
                packA = count_flags_all(fgc) | (count_flags_all(facgt) << 10) | (count_flags_all(fn) << 20);
            } else {
                const int split = boundary < (uint64_t)(o + hi) ? (int)(boundary - o) : hi;
                packA = count_flags(fgc, lo, split) | (count_flags(facgt, lo, split) << 10) |
                        (count_flags(fn, lo, split) << 20);
                packB = count_flags(fgc, split, hi) | (count_flags(facgt, split, hi) << 10) |
                        (count_flags(fn, split, hi) << 20);
            }
        }
        warp_window_add(a, winA, packA);
        if (__any_sync(0xffffffffu, packB != 0u)) warp_window_add(a, winB, packB);
    }
}

// any window length (used for wl < 16): one thread per window
__global__ void refstats_small_kernel(RefArgs a, uint64_t num_bins) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= num_bins) return;
    const uint64_t lo = w * a.wl, hi = min(lo + a.wl, a.len);
    uint32_t g = 0, c = 0, n = 0;
    for (uint64_t i = lo; i < hi; ++i) {
        const uint32_t b = a.seq[i];
        const uint32_t isgc = (b & 0xDBu) == 0x43u, x = b & 0xDFu;
        g += isgc;
        c += isgc | (x == 0x41u) | (x == 0x54u);
        n += x == 0x4Eu;
    }
    a.gc_cnt[w] = g;
    a.acgt_cnt[w] = c;
    a.n_cnt[w] = n;
}

__global__ void refstats_finalize_kernel(const uint32_t *__restrict__ gc_cnt,
                                         const uint32_t *__restrict__ acgt_cnt,
                                         const uint32_t *__restrict__ n_cnt, uint64_t num_bins,
                                         float *__restrict__ gc, uint8_t *__restrict__ has_gap) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= num_bins) return;
    const uint32_t c = acgt_cnt[w];
    // reference.rs:120-127: i32 counts -> f32 ratio, f32::missing() if the window has no ACGT
    gc[w] = c == 0 ? f32_missing() : __fdiv_rn((float)gc_cnt[w], (float)c);
    has_gap[w] = n_cnt[w] != 0;
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
    CNVB_TRY(ensure_scratch(ctx, Carver::need({nb * 4, nb * 4, nb * 4, nb * 4, nb})));
    Carver c(ctx->scratch);
    uint32_t *cnt = c.take<uint32_t>(nb * 3);
    float *d_gc = gc;
    uint8_t *d_gap = has_gap;
    const uint8_t *d_seq = seq;
    if (mem != CNVB_MEM_DEVICE) {
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
    a.div = make_fastdiv(window_length);
    a.gc_cnt = cnt;
    a.acgt_cnt = cnt + nb;
    a.n_cnt = cnt + 2 * nb;
    if (window_length >= 16) {
        CNVB_CUDA(ctx, cudaMemsetAsync(cnt, 0, nb * 12, ctx->stream), "clearing window counters");
        const uint64_t nchunks = (len + 15 + 15) / 16;
        uint64_t blocks = (nchunks + 255) / 256;
        const uint64_t mb = (uint64_t)ctx->sm_count * 8;
        if (blocks > mb) blocks = mb;
        {
            KernelSpan span(ctx, "refstats_count_kernel");
            refstats_count_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(a);
        }
        CNVB_LAUNCHED(ctx, "refstats_count_kernel");
    } else {
        refstats_small_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, ctx->stream>>>(a, nb);
        CNVB_LAUNCHED(ctx, "refstats_small_kernel");
    }
    refstats_finalize_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, ctx->stream>>>(
        a.gc_cnt, a.acgt_cnt, a.n_cnt, nb, d_gc, d_gap);
    CNVB_LAUNCHED(ctx, "refstats_finalize_kernel");
    if (mem != CNVB_MEM_DEVICE) {
        CNVB_CUDA(ctx, cudaMemcpyAsync(gc, d_gc, nb * 4, cudaMemcpyDeviceToHost, ctx->stream), "copy gc");
        CNVB_CUDA(ctx, cudaMemcpyAsync(has_gap, d_gap, nb, cudaMemcpyDeviceToHost, ctx->stream), "copy has_gap");
        CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "refstats");
    }
    return CNVB_OK;
}
