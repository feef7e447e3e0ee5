// wis_tensor.cuh -- tensor-core engine of compute_distances (lib-model-wis/src/lib.rs:128-197).
// Included by wis.cu only.
//
// All-pairs distances over the panel are a norms-minus-2*H*H^T contraction.  The reference's
// distance is an f32 sequential sum, so the tensor cores are used as a FILTER with a rigorous
// error bound and every surviving pair is re-evaluated exactly (wis_ref_distance):
//
//   prep     centre the panel (distances are translation invariant), scale by a power of two,
//            round to fp16: H[T][Kp].  Per row keep n_t = |h_t|^2 and delta_t = |u_t - h_t|
//            (the actual rounding error norm), so that for the real distance D and the fp16
//            distance d_h:  | D_ij - d_h,ij | <= delta_i + delta_j   (triangle inequality).
//   sweep    one CTA per 256 target rows, column tiles of 128 streamed by TMA through an mbarrier ring,
//            tcgen05.mma (cta_group::1, K16, fp16 x fp16 -> fp32) into TMEM accumulators, 16 epilogue warps
//            read them back with tcgen05.ld and test  n'_j - 2 g_ij <= theta_i + kappa_i smax, appending the
//            survivors to the row's candidate buffer in HBM.  Two kernels: up to 125 samples the rows' operand
//            stays in shared memory and the column norms ride in the contraction (wis_resident.cuh); beyond
//            that both operands stream K block by K block (wis_sweep_kernel below: CTA pairs with TMA
//            multicast, transposed M128 x N256 tiles for large panels).
//   order    rows are sorted by their projection p_t on the all-ones direction (the panel's first
//            principal direction: targets differ mostly by capture efficiency).  A projection is
//            1-Lipschitz, |p_i - p_j| <= D_ij, so a column can only be among the k nearest of row
//            i if |p_j - p_i| <= V_i (k-th smallest upper bound seen so far): in sorted order
//            every block of 256 rows needs one contiguous WINDOW of column tiles.
//   compact  between sweeps over windows that grow geometrically around the block's own
//            position (x4 per round, never beyond what V_i still allows) a warp per row selects
//            the k-th smallest candidate value and tightens theta'_i with the bound below.
//   rescore  exact f32-sequential distance for every candidate, (distance, index) top-k.
//
// Bound (scaled units; everything per PAIR so that a few large-norm rows do not loosen it for
// all): the fp16 squared distance d_h^2 and its tensor-core value differ by at most
// eps_ij = 2 c_acc s_i s_j + c_rnd (n_i + n_j)  (s = sqrt(n); fp32 accumulation over Kp terms and
// the f32 epilogue), the real distance D and d_h by at most delta_i + delta_j, the reference's
// f32 distance and D by a relative g.  Hence per pair
//   U_ij = (1+g)(sqrt(d2 + eps_ij) + delta_i + delta_j) >= D_ref >= (1-g)(sqrt(d2 - eps_ij) - delta_i - delta_j) = L_ij.
// With V = k-th smallest U_ij seen so far (>= the final k-th reference distance) a pair can only
// be among the top k if L_ij <= V.  The compaction kernel applies exactly that test; the sweep
// uses the cheaper sufficient form (delta_j <= rho s_j, rho = max_j delta_j / s_j)
//   n'_j - 2 g_ij - kappa_i s_j <= theta_i,   n'_j = n_j (1 - rho^2 - c_rnd),
//   kappa_i = 2 c_acc s_i + 2 rho V'_i,  theta_i = V'_i^2 - n_i (1 - c_rnd),  V'_i = V/(1-g) + delta_i,
// i.e. two FFMAs and a compare per pair.  Rows whose buffer overflows fall back to the exact row
// kernel, so the result never depends on the bound being tight.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

namespace {

constexpr uint32_t kTileRows = 256;   // target rows per CTA (two M=128 accumulators)
constexpr uint32_t kTileCols = 128;   // columns per B tile (UMMA N)
constexpr uint32_t kKBlock = 64;      // fp16 elements per 128-byte swizzle row
constexpr uint32_t kMaxKBlocks = 2;   // A stays resident: Kp <= 128
constexpr uint32_t kStages = 8;       // B ring: 8 x 16 KB
constexpr uint32_t kTileBytes = 128 * kKBlock * 2;  // 16 KB: 128 rows x 64 fp16
constexpr uint32_t kCandCap = 2048;   // candidate buffer entries per row
constexpr uint32_t kMaxTensorS = 4096;  // samples (K dimension) the tensor engine takes
constexpr uint32_t kMaxTensorK = 384;  // 32 x 24 first-window candidates must hold k after same-contig drops
constexpr uint32_t kEpiWarps = 16;     // 4 per TMEM lane quarter, each a quarter of the tile's columns
constexpr uint32_t kSweepThreads = (4 + kEpiWarps) * 32;

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// the same load delivered to every CTA of `mask` (same CTA-relative offsets, each CTA's own barrier)
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap *map, uint32_t bar, int32_t c0, int32_t c1,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
        "%4}], [%2], %5;" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// arrives on the barrier at the same offset in every CTA of `mask` when the MMAs issued so far have retired
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major, 128B-swizzled operand tile (rows of 64 fp16 = 128 B, 8-row atoms of 1024 B):
// start address >> 4 | LBO 1 | SBO 1024 B | descriptor version 1 (sm_100) | SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// kind::f16: D f32 (bit 4), A/B f16 K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | ((kTileCols >> 3) << 17) | ((128u >> 4) << 24);
// transposed tile: M = the 128 columns of the B tile, N = the 256 rows of the block
constexpr uint32_t kIdescX = (1u << 4) | ((kTileRows >> 3) << 17) | ((kTileCols >> 4) << 24);

// ---- prep -----------------------------------------------------------------------------------------
// column sums (f64) and the largest magnitude: a warp walks rows, lanes hold the columns lane + 32 q of
// the 128-column slice blockIdx.y
__global__ void __launch_bounds__(256) wis_colsum_kernel(const float *__restrict__ X, uint32_t T, uint32_t S,
                                                         double *__restrict__ colsum, float *__restrict__ absmax) {
    constexpr uint32_t kQ = 4;
    __shared__ double s_sum[128];
    for (uint32_t s = threadIdx.x; s < 128u; s += blockDim.x) s_sum[s] = 0.0;
    __syncthreads();
    const uint32_t lane = lane_id(), n_warps = gridDim.x * (blockDim.x >> 5), c0 = blockIdx.y * 128u;
    double acc[kQ];
    float mx = 0.0f;
#pragma unroll
    for (uint32_t q = 0; q < kQ; ++q) acc[q] = 0.0;
    for (uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < T; t += n_warps) {
#pragma unroll
        for (uint32_t q = 0; q < kQ; ++q) {
            const uint32_t c = c0 + lane + 32 * q;
            if (c < S) {
                const float v = X[(size_t)t * S + c];
                acc[q] += (double)v;
                mx = fmaxf(mx, fabsf(v));
            }
        }
    }
#pragma unroll
    for (uint32_t q = 0; q < kQ; ++q)
        if (c0 + lane + 32 * q < S) atomicAdd(&s_sum[lane + 32 * q], acc[q]);
    __syncthreads();
    for (uint32_t c = threadIdx.x; c < 128u && c0 + c < S; c += blockDim.x) atomicAdd(&colsum[c0 + c], s_sum[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) atomicMax(reinterpret_cast<int *>(absmax), __float_as_int(mx));  // non-negative floats order as ints
}

// the centre that is subtracted from every row (any vector works: distances are translation invariant)
__global__ void wis_cmean_kernel(const double *__restrict__ colsum, uint32_t T, uint32_t S, float *__restrict__ cmean) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < S) cmean[c] = (float)(colsum[c] / (double)T);
}

struct WisPrep {
    float scale;      // power of two
    float inv_scale;
};

// one warp per row: projection of the centred, scaled row on ones/sqrt(S), in f64 from the raw
// values (so |p_i - p_j| <= scale |x_i - x_j| holds up to the f64 summation error, covered by
// the window slack); written as an order-preserving u32 key of the f32-rounded value
__device__ __forceinline__ uint32_t f32_key(float v) {
    const uint32_t b = __float_as_uint(v);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float key_f32(uint32_t u) {
    return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

__global__ void wis_proj_kernel(const float *__restrict__ X, uint32_t T, uint32_t S, const float *__restrict__ cmean,
                                float scale, uint32_t *__restrict__ pkey, uint32_t *__restrict__ ident,
                                float *__restrict__ pabs_max) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (warp >= T) return;
    double acc = 0.0;
    for (uint32_t s = lane; s < S; s += 32) acc += (double)X[(size_t)warp * S + s] - (double)cmean[s];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        const float p = (float)(acc * (double)scale / sqrt((double)S));
        pkey[warp] = f32_key(p);
        ident[warp] = warp;
        atomicMax(reinterpret_cast<int *>(pabs_max), __float_as_int(fabsf(p)));
    }
}

// one warp per SORTED position: centre, scale, round the row perm[pos] to fp16; n_t (of the fp16
// row) and delta_t in f64; contig id and projection in sorted order.
//
// ortho (panels with more than 128 samples): the component along the all-ones direction is taken out
// of the fp16 row and carried as the f32 projection q_t instead, u_t = q_t 1 + h_t + r_t with
// 1 = ones / sqrt(S).  The sweep then evaluates  d^2 = (q_i - q_j)^2 + |h_i - h_j|^2, the first term in
// f32 in the epilogue.  Why: the accumulation error of the tensor dot product is bounded by
// c_acc s_i s_j with c_acc proportional to K; targets far from the panel mean relative to their noise
// (low capture efficiency) have norms s dominated by exactly this component, and at K = 1024 the
// bound let so many columns pass that their candidate buffers overflowed (6 % of the rows of a
// 10^6 x 1000 panel fell back to the exact kernel).  Without it s is the norm of the noise alone.
// The rounded h_t is not exactly orthogonal to 1: |u_i - u_j|^2 has the cross term
// 2 (q_i - q_j)(e_i - e_j), e_t = 1 . h_t, which moves the distance by at most 2 (|e_i| + |e_j|);
// delta_t = |r_t| + 2 |e_t| covers it (all measured in f64 from the raw values).
__global__ void wis_pack_kernel(const float *__restrict__ X, const uint32_t *__restrict__ tid, uint32_t T, uint32_t S,
                                uint32_t Kp, const float *__restrict__ cmean, float scale,
                                const uint32_t *__restrict__ perm, const uint32_t *__restrict__ pkey_sorted,
                                __half *__restrict__ H, float *__restrict__ nrm, float *__restrict__ delta,
                                uint32_t *__restrict__ tid_sorted, float *__restrict__ psorted,
                                float *__restrict__ maxes, int ortho, __half *__restrict__ HR) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (warp >= T) return;
    const uint32_t src = perm[warp];
    const float q = key_f32(pkey_sorted[warp]);
    const double inv_sqrt_s = 1.0 / sqrt((double)S);
    const double qc = ortho ? (double)q * inv_sqrt_s : 0.0;  // the component removed from every sample
    const float qcf = (float)qc;
    double n = 0.0, e = 0.0, eps = 0.0;
    for (uint32_t s = lane; s < Kp; s += 32) {
        __half h = __float2half_rn(0.0f);
        if (s < S) {
            const float u = __fsub_rn(__fmul_rn(__fsub_rn(X[(size_t)src * S + s], cmean[s]), scale), qcf);
            h = __float2half_rn(u);
            const double hf = (double)__half2float(h);
            n += hf * hf;
            eps += hf;
            // against the exact value (scale is a power of two; f64 rounding is far below the f32 / fp16 ones)
            const double r = ((double)X[(size_t)src * S + s] - (double)cmean[s]) * (double)scale - qc - hf;
            e += r * r;
        }
        H[(size_t)warp * Kp + s] = h;
        // resident sweep: the rows' operand, -2 h (exact: |h| < 2^11) and the three fold entries behind the samples
        if (HR) HR[(size_t)warp * Kp + s] = s < S ? __float2half_rn(-2.0f * __half2float(h)) : __float2half_rn(s < S + 3u ? 32768.0f : 0.0f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n += __shfl_xor_sync(0xffffffffu, n, o);
        e += __shfl_xor_sync(0xffffffffu, e, o);
        eps += __shfl_xor_sync(0xffffffffu, eps, o);
    }
    if (lane == 0) {
        const float nf = __double2float_ru(n);
        // rounding error norm, inflated for the f32 rounding of the centred value itself
        double d = sqrt(e) * (1.0 + 1.0 / 1024.0) + sqrt(n) * (1.0 / 8388608.0);
        if (ortho) d += 2.0 * fabs(eps) * inv_sqrt_s * (1.0 + 1.0 / 1024.0);
        const float df = __double2float_ru(d);
        nrm[warp] = nf;
        delta[warp] = df;
        tid_sorted[warp] = tid[src];
        psorted[warp] = q;
        const float rho = __double2float_ru((double)df / fmax(sqrt(n), 1.0));  // delta_j <= rho * max(s_j, 1)
        atomicMax(reinterpret_cast<int *>(maxes), __float_as_int(rho));
    }
}

// row subset [row_begin, row_end): flags over sorted positions, and the gather of their fp16 rows
__global__ void wis_rowflag_kernel(const uint32_t *__restrict__ perm, uint32_t T, uint32_t row_begin, uint32_t row_end,
                                   uint8_t *__restrict__ flag) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < T) flag[j] = (perm[j] >= row_begin && perm[j] < row_end) ? 1 : 0;
}
__global__ void wis_gather_rows_kernel(const __half *__restrict__ H, uint32_t Kp, const uint32_t *__restrict__ rows_a,
                                       uint32_t rows, uint32_t rows_pad, __half *__restrict__ HA) {
    const uint32_t r = blockIdx.x;
    const uint32_t words = Kp / 2;  // Kp is a multiple of 64
    const uint32_t *src = r < rows ? reinterpret_cast<const uint32_t *>(H + (size_t)rows_a[r] * Kp) : nullptr;
    uint32_t *dst = reinterpret_cast<uint32_t *>(HA + (size_t)r * Kp);
    for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src ? src[w] : 0u;
    (void)rows_pad;
}

// per-column operands of the sweep test: n'_j (+inf beyond T: never a candidate) and s_j
struct ColMeta {  // per target, sorted order: everything the bounds need in one 16-byte load
    float nrm, delta, s;
    uint32_t tid;
};

__global__ void wis_cols_kernel(const float *__restrict__ nrm, const float *__restrict__ delta,
                                const uint32_t *__restrict__ tid_sorted, uint32_t T, uint32_t Tpad,
                                const float *__restrict__ maxes, double c_rnd, float *__restrict__ nprime,
                                float *__restrict__ snorm, ColMeta *__restrict__ meta) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Tpad) return;
    if (j >= T) {
        nprime[j] = __int_as_float(0x7f800000);
        snorm[j] = 0.0f;
        meta[j] = ColMeta{0.0f, 0.0f, 0.0f, 0xFFFFFFFFu};
        return;
    }
    const double rho = (double)maxes[0], n = (double)nrm[j];
    nprime[j] = __double2float_rd(n * (1.0 - rho * rho - c_rnd));
    const float sj = __double2float_ru(fmax(sqrt(n), 1.0));
    snorm[j] = sj;
    meta[j] = ColMeta{nrm[j], delta[j], sj, tid_sorted[j]};
}

__global__ void wis_tile_smax_kernel(const float *__restrict__ snorm, uint32_t n_tiles, float *__restrict__ tile_smax) {
    const uint32_t tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = lane_id();
    if (tile >= n_tiles) return;
    float m = 0.0f;
    for (uint32_t c = lane; c < kTileCols; c += 32) m = fmaxf(m, snorm[tile * kTileCols + c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) tile_smax[tile] = m;
}

__global__ void wis_fill_kernel(float *__restrict__ p, uint32_t lo, uint32_t hi, float v) {
    const uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < hi) p[i] = v;
}

// ---- the error bounds in f32 with directed rounding (upper bounds round up, lower bounds down) --------
struct BoundConsts {
    float up_infl;       // ortho mode: stored values are rounded-down sums v + (q_i - q_j)^2; the exact sum is at
                         // most v (1 + up_infl) + up_infl (n_i + n_j)   (1 otherwise: no inflation)
    float c_acc2;        // 2 c_acc, rounded up
    float c_rnd;         // rounded up
    float one_plus_g;    // 1 + g, rounded up
    float inv_1mg;       // 1 / (1 - g), rounded up
    float one_m_crnd;    // 1 - c_rnd, rounded down
    const float *maxes;  // [0] rho
};

// >= D_ref for the pair behind a stored value v = n'_j - 2 g  (U_ij of the header comment)
__device__ __forceinline__ float pair_upper(float v, const ColMeta &mi, const ColMeta &mj, float rho2c, const BoundConsts &c) {
    const float nn = __fadd_ru(mi.nrm, mj.nrm);
    // (ortho: |v| <= the exact value + n_i + n_j, and the exact one exceeds v by at most 2^-21 of that)
    // (v = +inf for the padding columns of the first window: no 0 * inf)
    const float vu = c.up_infl > 0.0f ? __fadd_ru(v, __fmul_ru(c.up_infl, __fadd_ru(fabsf(v), nn))) : v;
    const float d2 = __fadd_ru(__fadd_ru(vu, __fmul_ru(mj.nrm, rho2c)), mi.nrm);
    const float eps = __fadd_ru(__fmul_ru(__fmul_ru(c.c_acc2, mi.s), mj.s), __fmul_ru(c.c_rnd, nn));
    const float sq = __fsqrt_ru(fmaxf(__fadd_ru(d2, eps), 0.0f));
    return __fmul_ru(c.one_plus_g, __fadd_ru(__fadd_ru(sq, mi.delta), mj.delta));
}
// can the pair still be among the k nearest?  L_ij <= V, written without a sqrt: d2 - eps <= (V' + delta_j)^2
__device__ __forceinline__ bool pair_alive(float v, const ColMeta &mi, const ColMeta &mj, float rho2c_rd, float vg,
                                           const BoundConsts &c) {
    const float d2 = __fadd_rd(__fadd_rd(v, __fmul_rd(mj.nrm, rho2c_rd)), mi.nrm);
    const float eps = __fadd_ru(__fmul_ru(__fmul_ru(c.c_acc2, mi.s), mj.s), __fmul_ru(c.c_rnd, __fadd_ru(mi.nrm, mj.nrm)));
    const float w = __fadd_ru(vg, mj.delta);
    return __fsub_rd(d2, eps) <= __fmul_ru(w, w);
}
// V (>= the row's final k-th reference distance) -> V' for the windows, theta and kappa for the sweep test
__device__ __forceinline__ void bound_to_thresholds(float V, const ColMeta &mi, float rho, const BoundConsts &c, float &vg,
                                                    float &theta, float &kappa) {
    vg = __fadd_ru(__fmul_ru(V, c.inv_1mg), mi.delta);
    theta = __fsub_ru(__fmul_ru(__fmul_ru(vg, vg), 1.0f + 1.0f / 4194304.0f), __fmul_rd(mi.nrm, c.one_m_crnd));
    kappa = __fmul_ru(__fadd_ru(__fmul_ru(c.c_acc2, mi.s), __fmul_ru(__fmul_ru(2.0f, rho), vg)), 1.0f + 1.0f / 1048576.0f);
}

// ---- sweep ------------------------------------------------------------------------------------------
struct SweepArgs {
    const float *nprime;     // [Tpad] n'_j (+inf beyond T), sorted column order
    const float *tile_smax;  // [n column tiles] max_j s_j over the tile
    const float *psorted;    // [Tpad] projections q_j in sorted order (STREAM mode: the (q_i - q_j)^2 term)
    const float *thr;        // [rows] theta_i (+inf until k candidates are known)
    const float *kap;        // [rows] kappa_i
    const uint4 *rng;        // [row blocks] column tiles of this round: [x, y) and [z, w)
    uint32_t T, rows;
    // first round only
    const uint32_t *tid_sorted;  // [Tpad] contig of every column (0xFFFFFFFF beyond T)
    const uint32_t *rows_a;  // [rows] sorted position of every row (nullptr: identity)
    uint32_t k_mmas;         // K16 slabs per tile that hold data: ceil(S / 16)
    uint32_t n_kblocks;      // ceil(Kp / 64)
    uint2 *cand;             // [rows][kCandCap] (bits of v = n'_j - 2 g, sorted column position): one 8-byte store per candidate
    uint32_t *cand_cnt;      // [rows]
    uint32_t dbg;            // CNVB_WIS_DBG (timing experiments only; results are wrong), bits: 1 skip the epilogue's
                             // TMEM reads and tests, 2 skip the MMAs, 4 skip the append's stores, 8 skip the appends
};

// STREAM: more than kMaxKBlocks K-blocks (S > 128): the A operand does not fit beside the B ring, so
// both operands stream through kStreamStages stages of (A rows 0..127 | A rows 128..255 | B) K-blocks.
constexpr uint32_t kStreamStages = 4;  // x 48 KB
constexpr int kDefaultCluster = 2;      // CTAs per cluster of the streamed sweep

// CLUSTER (with STREAM): two CTAs of a cluster share one block of 256 rows and take its column tiles in
// turn.  The streamed sweep is bound by L2 -> SM traffic (48 KB of operands per 256 x 128 x 64 MMA block =
// 87 flop/B against ~6300 B/clk of L2 bandwidth: measured 3.6-3.7 kflop/clk/SM, 45 % of the tensor
// peak); with the A K-block of the shared rows fetched once per pair -- each CTA loads one half and
// TMA-multicasts it into both shared memories -- a CTA moves 32 KB per block (131 flop/B).  A stage is
// refilled when BOTH consumers are done with it: the MMA issuers commit to the stage's empty barrier in
// both CTAs (multicast commit), which therefore counts two arrivals.  Candidate slots come from the
// global per-row counter (two CTAs append to the same rows).
template <bool FIRST, int CL, bool XPOSE>
__global__ void __launch_bounds__(kSweepThreads, 1)
wis_sweep_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const SweepArgs a) {
    constexpr bool CLUSTER = CL > 1;
    static_assert(CL == 1 || CL == 2 || CL == 4, "clusters of 2 or 4 CTAs");
    static_assert(!XPOSE || !FIRST, "the transposed tile exists for the later rounds");
    const uint32_t crank = CLUSTER ? cluster_ctarank() : 0u;
    constexpr bool XP_OWN = XPOSE && CL == 1;  // transposed tiles, one CTA per row block: the rows' counters live in shared memory
    const uint32_t blk = blockIdx.x / CL;
    // this block's share of the round: two tile ranges either side of what it has swept before
    const uint4 rng = a.rng[blk];
    const uint32_t n_lo = rng.y - rng.x, n_all = n_lo + (rng.w - rng.z);
    if (n_all == 0) return;  // (both CTAs of a cluster)
    // CLUSTER: tile i of this CTA is tile 2 i + rank of the block; both CTAs make `steps` passes through the
    // operand ring (the second one without a tile of its own in the last pass when the count is odd)
    const uint32_t n_tiles = (n_all + CL - 1u - crank) / CL;
    const uint32_t steps = (n_all + CL - 1u) / CL;
    const uint32_t rng_x = rng.x, rng_z = rng.z;
#define CNVB_TILE_G(g) ((g) < n_lo ? rng_x + (g) : rng_z + ((g) - n_lo))
#define CNVB_TILE_AT(it) CNVB_TILE_G((uint32_t)CL * (it) + crank)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint32_t s_cnt[kTileRows];  // candidates collected so far per row of this CTA
    __shared__ __align__(16) float s_col[kEpiWarps][64];  // per-warp column operands (n'_j | q_j)
    __shared__ __align__(16) float4 s_row[XPOSE ? kTileRows : 1];      // XPOSE: per-row operands (q_i, theta_i, kappa_i)
    __shared__ __align__(16) float2 s_rq[XPOSE ? kEpiWarps : 1][64];   // XPOSE: per warp and tile, (q_i, theta_i + kappa_i smax)
    // carve: A tiles [n_kblocks][2 row halves] x 16 KB (the halves of a K-block adjacent: one N = 256 operand) | B ring | barriers
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base;
    const uint32_t b_base = a_base + 2 * kMaxKBlocks * kTileBytes;
    const uint32_t bar_base = b_base + kStages * kTileBytes;
    const uint32_t bar_a_full = bar_base;
    const uint32_t bar_b_full = bar_base + 8;                 // kStages
    const uint32_t bar_b_empty = bar_b_full + 8 * kStages;    // kStages
    const uint32_t bar_acc_full = bar_b_empty + 8 * kStages;  // 2
    const uint32_t bar_acc_empty = bar_acc_full + 16;         // 2
    const uint32_t tmem_slot = bar_acc_empty + 16;
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t row0 = blk * kTileRows;  // position in the sorted row list
    if ((!CLUSTER && !XPOSE) || XP_OWN)
        for (uint32_t t = threadIdx.x; t < kTileRows; t += blockDim.x)
            s_cnt[t] = row0 + t < a.rows ? a.cand_cnt[row0 + t] : 0u;
    if (XPOSE)  // row operands (projection, theta, kappa): constant for the whole launch, read back as broadcasts
        for (uint32_t t = threadIdx.x; t < kTileRows; t += blockDim.x) {
            const uint32_t row = row0 + t;
            const bool live = row < a.rows;
            s_row[t] = make_float4(live ? a.psorted[a.rows_a ? a.rows_a[row] : row] : 0.0f,
                                   live ? a.thr[row] : -__int_as_float(0x7f800000), live ? a.kap[row] : 0.0f, 0.0f);
        }

    if (warp == 0 && lane == 0) {
        mbar_init(bar_a_full, 1);
        for (uint32_t s = 0; s < kStages; ++s) {
            mbar_init(bar_b_full + 8 * s, 1);
            mbar_init(bar_b_empty + 8 * s, CL);  // CLUSTER: the MMAs of all CTAs of the cluster release a stage
        }
        for (uint32_t s = 0; s < 2; ++s) {
            mbar_init(bar_acc_full + 8 * s, 1);
            mbar_init(bar_acc_empty + 8 * s, kEpiWarps);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: 4 accumulators x 128 columns (2 row halves x 2 stages)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CLUSTER) cluster_sync_all();  // the peer's barriers exist before anything is sent to them
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            constexpr uint32_t n_stages = kStreamStages;
            uint32_t st = 0, ph = 0;
            for (uint32_t it = 0; it < steps; ++it) {
                const bool have = it < n_tiles;  // (CLUSTER: the last pass may only serve the peer)
                const int32_t j0 = have ? (int32_t)(CNVB_TILE_AT(it) * kTileCols) : 0;
                for (uint32_t kb = 0; kb < a.n_kblocks; ++kb) {
                    mbar_wait(bar_b_empty + 8 * st, ph ^ 1u);
                    const uint32_t sb = a_base + st * 3u * kTileBytes;
                    mbar_expect_tx(bar_b_full + 8 * st, (have ? 3u : 2u) * kTileBytes);
                    if (CLUSTER) {  // this CTA's share of the rows (256 / CL of them), into all CTAs of the cluster
                        tma_load_2d_mc(sb + crank * (2u * kTileBytes / CL), &tmap_a, bar_b_full + 8 * st,
                                       (int32_t)(kb * kKBlock), (int32_t)(row0 + crank * (256u / CL)),
                                       (uint16_t)((1u << CL) - 1u));
                    } else {
                        tma_load_2d(sb, &tmap_a, bar_b_full + 8 * st, (int32_t)(kb * kKBlock), (int32_t)row0);
                        tma_load_2d(sb + kTileBytes, &tmap_a, bar_b_full + 8 * st, (int32_t)(kb * kKBlock), (int32_t)(row0 + 128));
                    }
                    if (have) tma_load_2d(sb + 2u * kTileBytes, &tmap_b, bar_b_full + 8 * st, (int32_t)(kb * kKBlock), j0);
                
                    if (++st == n_stages) {
                        st = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t n_stages = kStreamStages;
            uint32_t st = 0, ph = 0, as = 0, aph = 0;
            // descriptors differ from this base only in the start-address field (16-byte units):
            // +2 per K16 slab inside a 128-byte swizzle row, + bytes / 16 per tile or stage
            const uint64_t sdesc0 = umma_desc(a_base);
            for (uint32_t it = 0; it < steps; ++it) {
                const bool have = it < n_tiles;
                if (have) {
                    mbar_wait(bar_acc_empty + 8 * as, aph ^ 1u);
                    tc_fence_after();
                }
                for (uint32_t kb = 0; kb < a.n_kblocks; ++kb) {
                    mbar_wait(bar_b_full + 8 * st, ph);
                    tc_fence_after();
                    const uint64_t sd = sdesc0 + (uint64_t)(st * (3u * kTileBytes >> 4));
                    if (have && XPOSE) {
                        // one M128 x N256 MMA per K16 slab: the B tile's 128 columns on the M axis, the block's
                        // 256 rows (two adjacent 16 KB tiles) on the N axis -- 12 KB of operands per 128-cycle
                        // slot instead of 8 KB per 64-cycle slot
                        const uint32_t d_tmem = tmem_base + as * kTileRows;
#pragma unroll
                        for (uint32_t kk = 0; kk < kKBlock / 16; ++kk) {
                            if (kb * (kKBlock / 16) + kk < a.k_mmas)
                                tc_mma_f16(d_tmem, sd + (uint64_t)(2u * (kTileBytes >> 4)) + 2u * kk, sd + 2u * kk, kIdescX,
                                           (kb | kk) != 0u);
                        }
                    } else if (have) {
#pragma unroll
                        for (uint32_t h = 0; h < 2; ++h) {
                            const uint32_t d_tmem = tmem_base + (as * 2 + h) * kTileCols;
#pragma unroll
                            for (uint32_t kk = 0; kk < kKBlock / 16; ++kk) {
                                if (kb * (kKBlock / 16) + kk < a.k_mmas)
                                    tc_mma_f16(d_tmem, sd + (uint64_t)(h * (kTileBytes >> 4)) + 2u * kk,
                                               sd + (uint64_t)(2u * (kTileBytes >> 4)) + 2u * kk, kIdesc, (kb | kk) != 0u);
                            }
                        }
                    }
                    // frees the stage (in both CTAs: the peer's producer writes into this one's too)
                    if (CLUSTER) tc_commit_mc(bar_b_empty + 8 * st, (uint16_t)((1u << CL) - 1u)); else tc_commit(bar_b_empty + 8 * st);
                    if (++st == n_stages) {
                        st = 0;
                        ph ^= 1u;
                    }
                }
            
                if (have) {
                    tc_commit(bar_acc_full + 8 * as);  // accumulators of this tile complete
                    if (++as == 2) {
                        as = 0;
                        aph ^= 1u;
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: 16 warps.  Warp e reads TMEM lane quarter q = e & 3 (rows row0 + 32 q + lane
        // and + 128) and the 32 columns [32 (e >> 2), +32) of both accumulators.  With four warps per
        // scheduler the TMEM / L1 latencies of one warp hide behind the compares of the others. =====
        const uint32_t e = warp - 4, q = e & 3u, cpart = e >> 2;
        if (XPOSE) {
            // Transposed tile: TMEM lane = column of the tile (this thread's column j = 32 q + lane), TMEM
            // column = row of the block; warp e takes the rows [64 cpart, 64 cpart + 64).  The column operands
            // are two registers per thread; the row operands come from shared memory as broadcasts.
            uint32_t as = 0, aph = 0;
            // this warp's 64 rows: two per lane, kept in registers; per tile the lanes publish (q_i, theta_i +
            // kappa_i smax) to the warp's strip, and the pairs of two rows come back as one broadcast LDS.128
            const float4 ro0 = s_row[cpart * 64u + lane], ro1 = s_row[cpart * 64u + 32u + lane];
            for (uint32_t it = 0; it < n_tiles; ++it) {
                const uint32_t tile = CNVB_TILE_AT(it);
                const uint32_t j = tile * kTileCols + q * 32u + lane;
                const float nj = __ldg(a.nprime + j), pj = __ldg(a.psorted + j);
                const float smax = __ldg(a.tile_smax + tile);
                s_rq[e][lane] = make_float2(ro0.x, fmaf(ro0.z, smax, ro0.y));
                s_rq[e][32u + lane] = make_float2(ro1.x, fmaf(ro1.z, smax, ro1.y));
                __syncwarp();
                if (it + 1 < n_tiles) {
                    const uint32_t jn = CNVB_TILE_AT(it + 1) * kTileCols + q * 32u;
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.nprime + jn));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.psorted + jn));
                }
                mbar_wait(bar_acc_full + 8 * as, aph);
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t r0 = cpart * 64u + (uint32_t)h * 32u;  // first row of this batch within the block
                    if (a.dbg & 1u) continue;
                    uint32_t r[32];
                    tc_ld32(tmem_base + ((q * 32u) << 16) + as * kTileRows + r0, r);
                    uint32_t gmask = 0u;
#pragma unroll
                    for (int c = 0; c < 32; c += 2) {
                        // (q, rhs) of rows c and c + 1: same address for all lanes
                        const float4 two = *reinterpret_cast<const float4 *>(&s_rq[e][32 * h + c]);
                        const float d0 = __fsub_rz(pj, two.x), d1 = __fsub_rz(pj, two.z);
                        const float w0 = fmaf(-2.0f, __uint_as_float(r[c]), nj), w1 = fmaf(-2.0f, __uint_as_float(r[c + 1]), nj);
                        const float v0 = __fmaf_rd(d0, d0, w0);
                        const float v1 = __fmaf_rd(d1, d1, w1);
                        r[c] = __float_as_uint(v0);
                        r[c + 1] = __float_as_uint(v1);
                        gmask |= (v0 <= two.y || v1 <= two.w) ? (1u << (c >> 2)) : 0u;
                    }
                    const uint32_t wmask = (a.dbg & 8u) ? 0u : __reduce_or_sync(0xffffffffu, gmask);
                    if (wmask) {  // survivors are rare: the append code runs only for row groups in which some column passes
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) {
                            if (wmask & (1u << g4)) {
#pragma unroll
                                for (int e4 = 0; e4 < 4; ++e4) {
                                    // one row, this warp's 32 columns: the lanes that pass take consecutive slots
                                    // (ONE counter update per warp and row, the stores of a row coalesce)
                                    const int c = 4 * g4 + e4;
                                    const float v = __uint_as_float(r[c]);
                                    const bool pass = v <= s_rq[e][32 * h + c].y && j < a.T;  // (rows beyond the end: theta = -inf)
                                    const uint32_t bal = __ballot_sync(0xffffffffu, pass);
                                    if (bal) {
                                        const uint32_t row = row0 + r0 + c;
                                        uint32_t base = 0u;
                                        if (lane == 0)
                                            base = XP_OWN ? atomicAdd(&s_cnt[r0 + c], (uint32_t)__popc(bal))
                                                          : atomicAdd(&a.cand_cnt[row], (uint32_t)__popc(bal));
                                        base = __shfl_sync(0xffffffffu, base, 0);
                                        const uint32_t slot = base + (uint32_t)__popc(bal & ((1u << lane) - 1u));
                                        if (pass && slot < kCandCap && !(a.dbg & 4u)) {  // beyond: overflow, the row falls back
                                            a.cand[(size_t)row * kCandCap + slot] = make_uint2(__float_as_uint(v), j);
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_acc_empty + 8 * as);
                if (++as == 2) {
                    as = 0;
                    aph ^= 1u;
                }
            }
        } else {
        uint32_t rloc[2];
        float thr[2], kap[2], qrow[2];
        uint32_t trow[2];  // FIRST: contig of the row (a value no column has for rows beyond the end)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            rloc[h] = h * 128 + q * 32 + lane;  // row within the CTA's 256
            const uint32_t row = row0 + rloc[h];
            const bool live = row < a.rows;
            thr[h] = live ? a.thr[row] : -__int_as_float(0x7f800000);
            kap[h] = live ? a.kap[row] : 0.0f;
            trow[h] = 0xFFFFFFFEu;
            qrow[h] = 0.0f;
            if (FIRST && live) trow[h] = a.tid_sorted[a.rows_a ? a.rows_a[row] : row];
            if (live) qrow[h] = a.psorted[a.rows_a ? a.rows_a[row] : row];
        }
        uint32_t as = 0, aph = 0;
        for (uint32_t it = 0; it < n_tiles; ++it) {
            const uint32_t tile = CNVB_TILE_AT(it);
            const uint32_t j0 = tile * kTileCols + cpart * 32u;
            // column operands of this warp's 32 columns (same addresses for all lanes: L1 broadcast)
            float nj[32];
            // ortho mode needs two operands per column; 64 registers for them beside the 32 accumulator
            // values would spill.  Each lane fetches one column's pair (two coalesced 128-byte loads, before
            // the accumulator wait) into the warp's own shared-memory strip, and the values come back four
            // at a time as broadcast LDS.128 while the accumulators are consumed.
            s_col[e][lane] = __ldg(a.nprime + j0 + lane);
            s_col[e][32 + lane] = __ldg(a.psorted + j0 + lane);
            __syncwarp();
        
            const float smax = __ldg(a.tile_smax + tile);
            if (it + 1 < n_tiles) {  // the next tile's column operands: in L1 by the time they are needed
                const uint32_t jn = CNVB_TILE_AT(it + 1) * kTileCols + cpart * 32u;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a.nprime + jn));
                if (FIRST) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.tid_sorted + jn));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a.psorted + jn));
            }
            uint32_t tj[FIRST ? 32 : 1];
            if (FIRST) {
                const uint4 *tj4 = reinterpret_cast<const uint4 *>(a.tid_sorted + j0);
#pragma unroll
                for (int g4 = 0; g4 < (FIRST ? 8 : 0); ++g4) {
                    const uint4 v = __ldg(tj4 + g4);
                    tj[4 * g4 + 0] = v.x;
                    tj[4 * g4 + 1] = v.y;
                    tj[4 * g4 + 2] = v.z;
                    tj[4 * g4 + 3] = v.w;
                }
            }
            mbar_wait(bar_acc_full + 8 * as, aph);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (a.dbg & 1u) continue;
                uint32_t r[32];
                tc_ld32(tmem_base + ((q * 32u) << 16) + (as * 2 + h) * kTileCols + cpart * 32u, r);
                if (h == 1) {
                    // the tile's values are in registers: the accumulator stage goes back to the MMA issuer now,
                    // not after the tests and appends of this half
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_acc_empty + 8 * as);
                }
                // ortho mode: r[c] <- n'_j - 2 g + (q_j - q_i)^2.  The difference is rounded towards zero
                // and the FMA down, so the value never exceeds the exact one: a column that passes the
                // exact test passes this one, and the stored value is a lower bound (the compaction
                // inflates it for the upper bound).
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4) {
                    const float4 nv = *reinterpret_cast<const float4 *>(&s_col[e][4 * g4]);
                    const float4 pv = *reinterpret_cast<const float4 *>(&s_col[e][32 + 4 * g4]);
                    const float d0 = __fsub_rz(pv.x, qrow[h]), d1 = __fsub_rz(pv.y, qrow[h]);
                    const float d2 = __fsub_rz(pv.z, qrow[h]), d3 = __fsub_rz(pv.w, qrow[h]);
                    r[4 * g4 + 0] = __float_as_uint(__fmaf_rd(d0, d0, fmaf(-2.0f, __uint_as_float(r[4 * g4 + 0]), nv.x)));
                    r[4 * g4 + 1] = __float_as_uint(__fmaf_rd(d1, d1, fmaf(-2.0f, __uint_as_float(r[4 * g4 + 1]), nv.y)));
                    r[4 * g4 + 2] = __float_as_uint(__fmaf_rd(d2, d2, fmaf(-2.0f, __uint_as_float(r[4 * g4 + 2]), nv.z)));
                    r[4 * g4 + 3] = __float_as_uint(__fmaf_rd(d3, d3, fmaf(-2.0f, __uint_as_float(r[4 * g4 + 3]), nv.w)));
                }
            
                // the values under test, v_c = n'_j - 2 g (already formed in ortho mode)
                float v[32];
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    v[c] = __uint_as_float(r[c]);
                const uint32_t row = row0 + rloc[h];
                uint2 *cand = a.cand + (size_t)row * kCandCap;  // (rows beyond a.rows never pass: thr = -inf)
                const float inf = __int_as_float(0x7f800000);
                if (FIRST) {  // (same-contig columns are no candidates)
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] = tj[c] == trow[h] ? inf : v[c];
                }
                // minima of every four columns (FMNMX3 + FMNMX each) and of the row's 32 values
                float gm[8];
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4)
                    gm[g4] = fminf(fminf(v[4 * g4], v[4 * g4 + 1]), fminf(v[4 * g4 + 2], v[4 * g4 + 3]));
                const float m = fminf(fminf(fminf(gm[0], gm[1]), fminf(gm[2], gm[3])), fminf(fminf(gm[4], gm[5]), fminf(gm[6], gm[7])));
                if (FIRST) {
                    // First window, no bound yet: of every four columns the nearest one is stored -- 8 per row and
                    // warp, 32 per row and tile.  ANY k distinct columns give a valid bound (the k-th smallest
                    // upper bound among them); the lists are thrown away after it has been derived and the window
                    // is swept again with the bound.  (Minima of larger groups over a wider window were tried --
                    // one of 32 over 0.4 k tiles: the stored minima then come from tiles far away in projection,
                    // where even the nearest column is far, and the bound is useless.)
                    if (row < a.rows && m < inf) {
                        uint32_t pick = 0u, have = 0u;
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) {
                            const uint32_t w = v[4 * g4] == gm[g4] ? 0u : (v[4 * g4 + 1] == gm[g4] ? 1u : (v[4 * g4 + 2] == gm[g4] ? 2u : 3u));
                            pick |= w << (2 * g4);
                            have |= (gm[g4] < inf ? 1u : 0u) << g4;
                        }
                        const uint32_t cnt = (uint32_t)__popc(have);
                        uint32_t slot = CLUSTER ? atomicAdd(&a.cand_cnt[row], cnt) : atomicAdd(&s_cnt[rloc[h]], cnt);
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) {
                            if (have & (1u << g4)) {
                                if (slot < kCandCap) {
                                    cand[slot] = make_uint2(__float_as_uint(gm[g4]), j0 + 4u * g4 + ((pick >> (2 * g4)) & 3u));
                                }
                                ++slot;
                            }
                        }
                    }
                    continue;
                }
                // keep iff  n'_j - 2 g - kappa_i s_j <= theta_i;  sufficient: n'_j - 2 g <= theta_i + kappa_i smax.
                // Survivors are rare in the later rounds: the row's minimum decides, one vote per warp; the append
                // code runs only where some row of the warp has a column that passes, and there only for the groups
                // of four columns in which one does.
                const float rhs = fmaf(kap[h], smax, thr[h]);
                if (__any_sync(0xffffffffu, m <= rhs) && !(a.dbg & 8u)) {
                    uint32_t gmask = 0u;
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) gmask |= gm[g4] <= rhs ? (1u << g4) : 0u;
                    const uint32_t wmask = __reduce_or_sync(0xffffffffu, gmask);
                    // (rows without a bound yet have rhs = +inf: the padding columns, n'_j = +inf, must not pass)
                    const uint32_t valid = j0 + 32u <= a.T ? 0xFFFFFFFFu : (j0 < a.T ? (1u << (a.T - j0)) - 1u : 0u);
                    uint32_t mask = 0u;
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) {
                        if (wmask & (1u << g4)) {  // (warp-uniform)
#pragma unroll
                            for (int e4 = 0; e4 < 4; ++e4) mask |= v[4 * g4 + e4] <= rhs ? (1u << (4 * g4 + e4)) : 0u;
                        }
                    }
                    mask &= valid;
                    // one counter update per row, warp and tile: the row's passing columns take consecutive slots
                    const uint32_t cnt = (uint32_t)__popc(mask);
                    uint32_t slot = 0u;
                    if (cnt) slot = CLUSTER ? atomicAdd(&a.cand_cnt[row], cnt) : atomicAdd(&s_cnt[rloc[h]], cnt);
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) {
                        if (wmask & (1u << g4)) {
#pragma unroll
                            for (int e4 = 0; e4 < 4; ++e4) {
                                const int c = 4 * g4 + e4;
                                if (mask & (1u << c)) {
                                    if (slot < kCandCap && !(a.dbg & 4u)) {  // beyond: overflow, the row falls back
                                        cand[slot] = make_uint2(__float_as_uint(v[c]), j0 + (uint32_t)c);
                                    }
                                    ++slot;
                                }
                            }
                        }
                    }
                }
            }
            if (a.dbg & 1u) {  // (the halves were skipped: release the stage here)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_acc_empty + 8 * as);
            }
            if (++as == 2) {
                as = 0;
                aph ^= 1u;
            }
        }
        }  // !XPOSE
    }
    tc_fence_before();
    __syncthreads();
    if ((!CLUSTER && !XPOSE) || XP_OWN)
        for (uint32_t t = threadIdx.x; t < kTileRows; t += blockDim.x)
            if (row0 + t < a.rows) a.cand_cnt[row0 + t] = s_cnt[t];
    if (CLUSTER) cluster_sync_all();  // no CTA leaves while its peer may still send to it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

#undef CNVB_TILE_AT
#undef CNVB_TILE_G

// ---- windows: which column tiles a block of rows still has to see --------------------------------
struct WindowArgs {
    const float *psorted;     // [T] projections, ascending
    const uint32_t *rows_a;   // [rows] sorted position of every row (nullptr: identity)
    const float *vg;          // [rows] V'_i = V_i / (1 - g) + delta_i (+inf: no bound yet)
    const uint32_t *overflow; // [rows] rows that fell back to the exact kernel need nothing
    uint32_t T, n_tiles, rows;
    uint32_t round, w0, grow; // round 0: own tiles +- w0; later: at most `grow` more tiles per side
    float slack;              // projection rounding slack (absolute)
    uint4 *rng;               // [row blocks] out: tile ranges [x, y) and [z, w) of this round
    uint2 *done;              // [row blocks] in/out: tiles swept so far [x, y)
    unsigned long long *tiles_swept;
};

__device__ __forceinline__ uint32_t lower_bound_f32(const float *p, uint32_t n, float v) {  // first idx with p[idx] >= v
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (p[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ uint32_t upper_bound_f32(const float *p, uint32_t n, float v) {  // first idx with p[idx] > v
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (p[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kTileRows) wis_window_kernel(WindowArgs a) {
    __shared__ float s_lo[kTileRows / 32], s_hi[kTileRows / 32];
    const uint32_t r = blockIdx.x * kTileRows + threadIdx.x;
    const float inf = __int_as_float(0x7f800000);
    float lo = inf, hi = -inf;
    if (r < a.rows && !(a.round > 0 && a.overflow[r])) {
        const float p = a.psorted[a.rows_a ? a.rows_a[r] : r];
        float w = 0.0f;
        if (a.round > 0) w = __fadd_ru(__fmul_ru(a.vg[r], 1.0f + 1.0f / 1048576.0f), a.slack);  // inf stays inf
        lo = __fsub_rd(p, w);
        hi = __fadd_ru(p, w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane_id() == 0) {
        s_lo[threadIdx.x >> 5] = lo;
        s_hi[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (uint32_t w = 1; w < kTileRows / 32; ++w) {
        lo = fminf(lo, s_lo[w]);
        hi = fmaxf(hi, s_hi[w]);
    }
    uint4 out = make_uint4(0, 0, 0, 0);
    uint2 dn = a.round ? a.done[blockIdx.x] : make_uint2(0, 0);
    if (lo <= hi) {  // some row of the block still takes part
        const uint32_t req_lo = lower_bound_f32(a.psorted, a.T, lo) / kTileCols;
        const uint32_t ub = upper_bound_f32(a.psorted, a.T, hi);
        const uint32_t req_hi = min(a.n_tiles, (ub + kTileCols - 1) / kTileCols);
        if (a.round == 0) {
            // 2 (1 + w0) tiles centred on the block, for the first bound only (rows of a sparse
            // subset see an off-centre window, which merely loosens that bound); nothing is stored
            // in this round, so no tile counts as swept yet
            const uint32_t c = min((req_lo + req_hi) >> 1, a.n_tiles), half = 1u + a.w0;
            const uint32_t size = min(a.n_tiles, 2u * half);
            out.x = c > half ? c - half : 0u;
            if (out.x + size > a.n_tiles) out.x = a.n_tiles - size;
            out.y = out.x + size;
            dn = make_uint2(c, c);
        } else {
            const uint32_t lim_lo = dn.x > a.grow ? dn.x - a.grow : 0u;
            const uint32_t lim_hi = min(a.n_tiles, dn.y + a.grow);
            const uint32_t new_lo = min(dn.x, max(req_lo, lim_lo));
            const uint32_t new_hi = max(dn.y, min(req_hi, lim_hi));
            out = make_uint4(new_lo, dn.x, dn.y, new_hi);
            dn = make_uint2(new_lo, new_hi);
        }
    }
    a.rng[blockIdx.x] = out;
    a.done[blockIdx.x] = dn;
    const uint32_t n = (out.y - out.x) + (out.w - out.z);
    if (n) atomicAdd(a.tiles_swept, (unsigned long long)n);
}

// ---- compaction: k-th smallest candidate value per row, new threshold -----------------------------
struct CompactArgs {
    uint2 *cand;             // (bits of the stored v = n'_j - 2 g, sorted column position)
    uint32_t *cand_cnt;
    float *thr, *kap;
    uint32_t *overflow;      // [rows] sticky flag
    uint32_t *seen;          // [rows] list length after the row's last compaction (unchanged lists are skipped)
    float *vg;               // [rows] out: V'_i for the window kernel
    const uint32_t *rows_a;  // [rows] sorted position of every row (nullptr: identity)
    const ColMeta *meta;     // [Tpad] sorted order
    uint32_t rows, k;
    uint32_t discard;        // first round: derive the bound, then drop the list (the window is swept again)
    unsigned long long *max_len;  // longest list any compaction left behind
    BoundConsts bc;
};

constexpr uint32_t kCompactWarps = 8;
constexpr uint32_t kKeyRegs = 16;  // candidates per lane whose keys stay in registers (lists up to 512)

// the bin (of 256, warp-private in shared memory) that holds the want-th smallest entry; returns the
// bin and leaves in `want` the rank inside it
__device__ __forceinline__ uint32_t hist_find(const uint32_t *hist, uint32_t &want) {
    const uint32_t lane = lane_id();
    uint32_t c[8], sum = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        c[t] = hist[8 * lane + t];
        sum += c[t];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += u;
    }
    const uint32_t owner = __ffs(__ballot_sync(0xffffffffu, incl >= want)) - 1;  // want <= total by construction
    uint32_t bin = 0, rest = 0;
    if (lane == owner) {
        uint32_t w = want - (incl - sum);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (w > c[t] && t < 7) {
                w -= c[t];
            } else {
                bin = 8 * lane + t;
                rest = w;
                break;
            }
        }
    }
    bin = __shfl_sync(0xffffffffu, bin, owner);
    want = __shfl_sync(0xffffffffu, rest, owner);
    return bin;
}

// One warp per row: upper bounds U_ij of all stored pairs, V = (an upper bound of) the k-th smallest of
// them by a two-level 256-bin histogram over the key range, then only the pairs with L_ij <= V stay.
// (six CTAs per SM: the kernel waits on its gathers -- ncu: 11 - 13 long-scoreboard stalls per issue at 44 % occupancy
// with the 64 registers ptxas takes unasked; at 40 registers and 52 bytes of spills: 2.29 -> 1.93 ms at config 3)
__global__ void __launch_bounds__(kCompactWarps * 32, 6) wis_compact_kernel(CompactArgs a) {
    __shared__ uint32_t s_hist[kCompactWarps][256];
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (r >= a.rows) return;
    const uint32_t n = a.cand_cnt[r];
    if (n > kCandCap) {
        if (lane == 0) {
            a.overflow[r] = 1u;
            a.cand_cnt[r] = 0u;
            a.seen[r] = 0u;
            a.thr[r] = -__int_as_float(0x7f800000);  // stop collecting: the row is recomputed exactly
            a.kap[r] = 0.0f;
        }
        return;
    }
    if (a.overflow[r] || n == a.seen[r]) return;  // nothing new since the last compaction
    uint32_t *hist = s_hist[threadIdx.x >> 5];
    uint2 *cand = a.cand + (size_t)r * kCandCap;
    const ColMeta mi = a.meta[a.rows_a ? a.rows_a[r] : r];
    const float rho = a.bc.maxes[0];
    const float rho2c = __fadd_ru(__fmul_ru(rho, rho), a.bc.c_rnd), rho2c_rd = __fadd_rd(__fmul_rd(rho, rho), a.bc.c_rnd);
    // key of candidate c: ordered bits of U_ij, all ones for same-contig pairs (never candidates)
    auto key_of = [&](uint32_t c) -> uint32_t {
        const uint2 en = cand[c];
        const ColMeta mj = a.meta[en.y];
        if (mj.tid == mi.tid) return 0xFFFFFFFFu;
        return __float_as_uint(pair_upper(__uint_as_float(en.x), mi, mj, rho2c, a.bc)) | 0x80000000u;  // U >= 0
    };
    uint32_t key[kKeyRegs];
    uint32_t kmin = 0xFFFFFFFFu, kmax = 0u, n_valid = 0;
#pragma unroll
    for (uint32_t e = 0; e < kKeyRegs; ++e) {
        const uint32_t c = e * 32 + lane;
        key[e] = c < n ? key_of(c) : 0xFFFFFFFFu;
        if (key[e] != 0xFFFFFFFFu) {
            kmin = min(kmin, key[e]);
            kmax = max(kmax, key[e]);
            n_valid += 1;
        }
    }
    for (uint32_t c = kKeyRegs * 32 + lane; c < n; c += 32) {  // long lists: keys are recomputed per pass
        const uint32_t kk = key_of(c);
        if (kk != 0xFFFFFFFFu) {
            kmin = min(kmin, kk);
            kmax = max(kmax, kk);
            n_valid += 1;
        }
    }
    n_valid = __reduce_add_sync(0xffffffffu, n_valid);
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    const bool have_bound = n_valid >= a.k;  // otherwise: only drop the same-contig entries
    float vg = __int_as_float(0x7f800000), theta = vg, kappa = 0.0f;
    if (have_bound) {
        // quantise the key range to 16 bits, select in two 8-bit levels
        const uint32_t span = kmax - kmin;
        const uint32_t sh = span >> 16 ? (32u - (uint32_t)__clz(span)) - 16u : 0u;
        uint32_t want = a.k;
        for (uint32_t b = lane; b < 256; b += 32) hist[b] = 0u;
        __syncwarp();
#pragma unroll
        for (uint32_t e = 0; e < kKeyRegs; ++e)
            if (key[e] != 0xFFFFFFFFu) atomicAdd(&hist[((key[e] - kmin) >> sh) >> 8], 1u);
        for (uint32_t c = kKeyRegs * 32 + lane; c < n; c += 32) {
            const uint32_t kk = key_of(c);
            if (kk != 0xFFFFFFFFu) atomicAdd(&hist[((kk - kmin) >> sh) >> 8], 1u);
        }
        __syncwarp();
        const uint32_t b1 = hist_find(hist, want);
        __syncwarp();
        for (uint32_t b = lane; b < 256; b += 32) hist[b] = 0u;
        __syncwarp();
#pragma unroll
        for (uint32_t e = 0; e < kKeyRegs; ++e) {
            const uint32_t qk = (key[e] - kmin) >> sh;
            if (key[e] != 0xFFFFFFFFu && (qk >> 8) == b1) atomicAdd(&hist[qk & 255u], 1u);
        }
        for (uint32_t c = kKeyRegs * 32 + lane; c < n; c += 32) {
            const uint32_t kk = key_of(c), qk = (kk - kmin) >> sh;
            if (kk != 0xFFFFFFFFu && (qk >> 8) == b1) atomicAdd(&hist[qk & 255u], 1u);
        }
        __syncwarp();
        const uint32_t b2 = hist_find(hist, want);
        __syncwarp();
        // largest key of the selected fine bin: >= the k-th smallest U, hence >= the row's final k-th distance
        const uint32_t vkey = min(kmax, kmin + ((((b1 << 8) | b2) + 1u) << sh) - 1u);
        bound_to_thresholds(__uint_as_float(vkey & 0x7FFFFFFFu), mi, rho, a.bc, vg, theta, kappa);
    }
    // keep exactly the pairs that can still be among the top k
    uint32_t kept = 0;
    for (uint32_t base = 0; base < (a.discard ? 0u : n); base += 32) {
        const uint32_t c = base + lane;
        float v = 0.0f;
        uint32_t j = 0;
        bool keep = false;
        if (c < n) {
            const uint2 en = cand[c];
            v = __uint_as_float(en.x);
            j = en.y;
            const ColMeta mj = a.meta[j];
            keep = mj.tid != mi.tid && (!have_bound || pair_alive(v, mi, mj, rho2c_rd, vg, a.bc));
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const uint32_t pos = kept + __popc(m & ((1u << lane) - 1u));
        __syncwarp();
        if (keep) {
            cand[pos] = make_uint2(__float_as_uint(v), j);  // pos <= c: compaction in place, front to back
        }
        kept += __popc(m);
        __syncwarp();
    }
    if (lane == 0) {
        a.cand_cnt[r] = kept;
        a.seen[r] = kept;
        if (kept > 64u) atomicMax(a.max_len, (unsigned long long)kept);  // (at least 64 is assumed anyway)
        if (have_bound) {
            a.thr[r] = theta;
            a.kap[r] = kappa;
            a.vg[r] = vg;
        }
    }
}

// ---- exact re-scoring and (distance, index) top-k ------------------------------------------------
constexpr uint32_t kRescoreWarps = 4;
constexpr uint32_t kStageF4 = 7;               // float4 per staged row chunk (28 samples)
constexpr uint32_t kStageStride = kStageF4 * 4;  // 28 words: float4 reads of 8 consecutive rows hit 8 distinct bank groups

// One CTA per row.  STAGED: every warp copies 32 candidate rows chunk by chunk into shared memory
// with coalesced 16-byte loads (a lane walking its own row costs a 32-byte sector per 16 bytes and is
// bound by L1 throughput), then lane c accumulates candidate c in the reference's order.
template <bool STAGED>
__global__ void __launch_bounds__(kRescoreWarps * 32)
wis_rescore_kernel(const float *__restrict__ X, const uint32_t *__restrict__ tid, uint32_t T, uint32_t S, uint32_t k,
                   uint32_t row_begin, uint32_t rows, const uint32_t *__restrict__ perm,
                   const uint32_t *__restrict__ rows_a, const uint2 *__restrict__ cand,
                   const uint32_t *__restrict__ cand_cnt, const uint32_t *__restrict__ overflow, uint32_t key_cap,
                   float *__restrict__ out_dist, uint32_t *__restrict__ out_idx) {
    extern __shared__ __align__(16) uint8_t s_raw[];
    // [key_cap] u64 keys (distance bits << 32 | target index) | [S4] target row | STAGED: [warps][32][28] stage
    unsigned long long *s_key = reinterpret_cast<unsigned long long *>(s_raw);
    float *s_xi = reinterpret_cast<float *>(s_raw + (size_t)key_cap * 8);
    const uint32_t S4 = (S + 3u) & ~3u;
    float *s_stage = s_xi + S4;
    const uint32_t r = blockIdx.x;
    if (overflow[r]) return;  // recomputed by the exact row kernel
    const uint32_t i = perm[rows_a ? rows_a[r] : r];  // target index of this row
    const uint32_t n = cand_cnt[r];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    for (uint32_t s = threadIdx.x; s < S; s += blockDim.x) s_xi[s] = X[(size_t)i * S + s];
    __syncthreads();
    if (STAGED) {  // S % 4 == 0 and X 16-byte aligned
        float *stage = s_stage + (size_t)warp * 32 * kStageStride;
        const uint32_t row_f4 = S >> 2;
        for (uint32_t c0 = warp * 32; c0 < n; c0 += kRescoreWarps * 32) {
            const uint32_t c = c0 + lane;
            const uint32_t j = c < n ? perm[cand[(size_t)r * kCandCap + c].y] : i;  // (padding lanes: any valid row)
            float y = 0.0f;
            for (uint32_t f0 = 0; f0 < row_f4; f0 += kStageF4) {
                const uint32_t nf = min(kStageF4, row_f4 - f0);
                __syncwarp();
                for (uint32_t e = lane; e < 32 * nf; e += 32) {
                    const uint32_t cc = e / nf, f = e - cc * nf;
                    const uint32_t jj = __shfl_sync(0xffffffffu, j, cc);
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(X + (size_t)jj * S) + f0 + f);
                    *reinterpret_cast<float4 *>(stage + cc * kStageStride + 4 * f) = v;
                }
                __syncwarp();
                const float4 *mine = reinterpret_cast<const float4 *>(stage + lane * kStageStride);
                const float4 *xi4 = reinterpret_cast<const float4 *>(s_xi) + f0;
                for (uint32_t f = 0; f < nf; ++f) {
                    const float4 p = xi4[f], q4 = mine[f];
                    float x = __fsub_rn(p.x, q4.x);
                    y = __fadd_rn(y, __fmul_rn(x, x));
                    x = __fsub_rn(p.y, q4.y);
                    y = __fadd_rn(y, __fmul_rn(x, x));
                    x = __fsub_rn(p.z, q4.z);
                    y = __fadd_rn(y, __fmul_rn(x, x));
                    x = __fsub_rn(p.w, q4.w);
                    y = __fadd_rn(y, __fmul_rn(x, x));
                }
            }
            if (c < n) s_key[c] = ((unsigned long long)__float_as_uint(__fsqrt_rn(y)) << 32) | j;
        }
    } else {
        for (uint32_t c = threadIdx.x; c < n; c += blockDim.x) {
            const uint32_t j = perm[cand[(size_t)r * kCandCap + c].y];
            s_key[c] = ((unsigned long long)__float_as_uint(wis_ref_distance(s_xi, X + (size_t)j * S, S, false)) << 32) | j;
        }
    }
    __syncthreads();
    // (distance, index) rank of every candidate: distances are non-negative, so their bit patterns
    // order like the values and one 64-bit compare implements the documented tie rule
    const size_t o = (size_t)(i - row_begin) * k;
    for (uint32_t c = threadIdx.x; c < n; c += blockDim.x) {
        const unsigned long long mine = s_key[c];
        uint32_t rank = 0;
        for (uint32_t e = 0; e < n; ++e) rank += s_key[e] < mine ? 1u : 0u;
        if (rank < k) {
            out_dist[o + rank] = __uint_as_float((uint32_t)(mine >> 32));
            out_idx[o + rank] = (uint32_t)mine;
        }
    }
    if (threadIdx.x == 0 && n < k) {  // +inf entries of the reference, in index order
        const uint32_t ti = tid[i];
        uint32_t w = n;
        for (uint32_t j = 0; j < T && w < k; ++j) {
            if (tid[j] == ti) {
                out_dist[o + w] = __int_as_float(0x7f800000);
                out_idx[o + w] = j;
                ++w;
            }
        }
    }
}

__global__ void wis_overflow_list_kernel(const uint32_t *__restrict__ overflow, const uint32_t *__restrict__ cand_cnt,
                                         uint32_t rows, const uint32_t *__restrict__ perm,
                                         const uint32_t *__restrict__ rows_a, uint32_t *__restrict__ list,
                                         uint32_t *__restrict__ count, unsigned long long *__restrict__ n_rescored) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t c = 0;
    if (r < rows) {
        if (overflow[r]) list[atomicAdd(count, 1u)] = perm[rows_a ? rows_a[r] : r]; else c = cand_cnt[r];
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane_id() == 0 && c) atomicAdd(n_rescored, (unsigned long long)c);
}

}  // namespace
#include "wis_resident.cuh"
namespace {

// ---- host orchestration ------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int make_tmap(cnvb_ctx *ctx, void *H, uint64_t T, uint64_t Kp, CUtensorMap *out, uint32_t box_rows = 128) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || !p) return fail(ctx, CNVB_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
        fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    const cuuint64_t dims[2] = {Kp, T};
    const cuuint64_t strides[1] = {Kp * 2};
    const cuuint32_t box[2] = {kKBlock, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, H, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, CNVB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

int run_rows_exact(cnvb_ctx *ctx, const float *dX, const uint32_t *dtid, uint32_t T, uint32_t S, uint32_t k,
                   const uint32_t *d_row_list, uint32_t n_rows, uint32_t row_begin, float *d_dist, uint32_t *d_idx,
                   uint32_t *d_nan_flag);

struct PoolGuard {  // releases pooled blocks at scope exit
    cnvb_ctx *ctx;
    std::vector<void *> ptrs;
    explicit PoolGuard(cnvb_ctx *c) : ctx(c) {}
    template <typename T>
    cudaError_t alloc(size_t n, T **out) {
        cudaError_t e = pool_alloc_t(ctx, n, out);
        if (e == cudaSuccess) ptrs.push_back(*out);
        return e;
    }
    ~PoolGuard() {
        for (void *p : ptrs) pool_release(ctx, p);
    }
};

int wis_distances_tensor(cnvb_ctx *ctx, const float *dX, const uint32_t *dtid, uint32_t T, uint32_t S, uint32_t k,
                         uint32_t row_begin, uint32_t row_end, float *d_dist, uint32_t *d_idx) {
    const uint32_t rows = row_end - row_begin;
    if (rows == 0) return 0;
    // up to 125 samples: the resident sweep (rows in tensor memory, column norms folded into the contraction;
    // wis_resident.cuh).  CNVB_WIS_RES = 0 selects the shared-memory form it replaced.
    const bool res = S + 3u <= kResK && [] {
        const char *e = getenv("CNVB_WIS_RES");
        return !(e && atoi(e) == 0);
    }();
    const uint32_t Kp = res ? kResK : ((S + kKBlock - 1) / kKBlock) * kKBlock;
    if (Kp > kMaxTensorS)
        return fail(ctx, CNVB_ERR_INVALID, "tensor engine: more than %u samples; use the exact engine", kMaxTensorS);
    const bool stream_a = !res;  // more than 125 samples: both operands stream through the K-block ring
    if (k > kMaxTensorK) return fail(ctx, CNVB_ERR_INVALID, "tensor engine: max_ref_targets > %u", kMaxTensorK);
    const uint32_t n_tiles = (T + kTileCols - 1) / kTileCols;
    const uint32_t Tpad = n_tiles * kTileCols;
    const bool subset = rows != T;
    const uint32_t row_blocks = (rows + kTileRows - 1) / kTileRows;
    const uint32_t rows_pad = row_blocks * kTileRows;
    PoolGuard g(ctx);
    double *colsum = nullptr;
    float *scal = nullptr;  // [0] absmax | [1] rho | [2] max |p|
    __half *H = nullptr, *HA = nullptr, *HR = nullptr;
    float *nrm = nullptr, *delta = nullptr, *thr = nullptr, *kap = nullptr, *vg = nullptr,
          *nprime = nullptr, *snorm = nullptr, *tile_smax = nullptr, *psorted = nullptr;
    uint2 *cand = nullptr;
    uint32_t *cand_cnt = nullptr, *overflow = nullptr, *seen = nullptr, *ovf_list = nullptr,
             *pkey = nullptr, *pkey_sorted = nullptr, *ident = nullptr, *perm = nullptr, *tid_sorted = nullptr,
             *rows_a = nullptr, *n_sel = nullptr;
    uint8_t *rflag = nullptr;
    float *cmean = nullptr;
    ColMeta *meta = nullptr;
    uint4 *rng = nullptr;
    uint2 *done = nullptr;
    unsigned long long *counters = nullptr;  // [0] tiles swept | [1] pairs re-scored | [2] longest list
    void *cub_tmp = nullptr;
    size_t sort_bytes = 0, sel_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, pkey, pkey_sorted, ident, perm, (int)T, 0, 32, ctx->stream);
    if (subset)
        cub::DeviceSelect::Flagged(nullptr, sel_bytes, cub::CountingInputIterator<uint32_t>(0), rflag, rows_a, n_sel, (int)T,
                                   ctx->stream);
    cudaError_t e;
    if ((e = g.alloc(S, &colsum)) != cudaSuccess || (e = g.alloc(4, &scal)) != cudaSuccess ||
        (e = g.alloc((size_t)Tpad * Kp, &H)) != cudaSuccess || (e = g.alloc(Tpad, &nrm)) != cudaSuccess ||
        (e = g.alloc(Tpad, &delta)) != cudaSuccess || (e = g.alloc(rows, &thr)) != cudaSuccess ||
        (e = g.alloc(rows, &kap)) != cudaSuccess || (e = g.alloc(rows, &vg)) != cudaSuccess ||
        (e = g.alloc(Tpad, &nprime)) != cudaSuccess || (e = g.alloc(Tpad, &snorm)) != cudaSuccess ||
        (e = g.alloc(n_tiles, &tile_smax)) != cudaSuccess || (e = g.alloc(Tpad, &psorted)) != cudaSuccess ||
        (e = g.alloc((size_t)rows * kCandCap, &cand)) != cudaSuccess || (e = g.alloc(rows, &cand_cnt)) != cudaSuccess ||
        (e = g.alloc(rows, &overflow)) != cudaSuccess || (e = g.alloc(rows, &seen)) != cudaSuccess ||
        (e = g.alloc(rows + 1, &ovf_list)) != cudaSuccess || (e = g.alloc(T, &pkey)) != cudaSuccess ||
        (e = g.alloc(T, &pkey_sorted)) != cudaSuccess || (e = g.alloc(T, &ident)) != cudaSuccess ||
        (e = g.alloc(T, &perm)) != cudaSuccess || (e = g.alloc(Tpad, &tid_sorted)) != cudaSuccess ||
        (e = g.alloc(row_blocks, &rng)) != cudaSuccess || (e = g.alloc(row_blocks, &done)) != cudaSuccess ||
        (e = g.alloc(4, &counters)) != cudaSuccess || (e = g.alloc(S, &cmean)) != cudaSuccess ||
        (e = g.alloc(Tpad, &meta)) != cudaSuccess ||
        (e = g.alloc(std::max(sort_bytes, sel_bytes) + 16, reinterpret_cast<uint8_t **>(&cub_tmp))) != cudaSuccess)
        return fail_cuda(ctx, e, "allocating the tensor engine's work space");
    if (subset && ((e = g.alloc(rows, &rows_a)) != cudaSuccess || (e = g.alloc(T, &rflag)) != cudaSuccess ||
                   (e = g.alloc(1, &n_sel)) != cudaSuccess ||
                   (e = g.alloc((size_t)rows_pad * Kp, &HA)) != cudaSuccess))
        return fail_cuda(ctx, e, "allocating the tensor engine's work space");
    if (res && (e = g.alloc((size_t)Tpad * Kp, &HR)) != cudaSuccess)
        return fail_cuda(ctx, e, "allocating the tensor engine's work space");
    cudaStream_t st = ctx->stream;
    cudaMemsetAsync(colsum, 0, S * sizeof(double), st);
    cudaMemsetAsync(scal, 0, 4 * sizeof(float), st);
    cudaMemsetAsync(cand_cnt, 0, rows * 4, st);
    cudaMemsetAsync(overflow, 0, rows * 4, st);
    cudaMemsetAsync(seen, 0, rows * 4, st);
    cudaMemsetAsync(ovf_list, 0, (rows + 1) * 4, st);
    cudaMemsetAsync(counters, 0, 32, st);
    wis_colsum_kernel<<<dim3(std::min<uint32_t>((T + 7) / 8, (uint32_t)ctx->sm_count * 8), (S + 127) / 128), 256, 0, st>>>(
        dX, T, S, colsum, scal);
    CNVB_LAUNCHED(ctx, "wis_colsum_kernel");
    float absmax = 0.0f;
    CNVB_CUDA(ctx, cudaMemcpyAsync(&absmax, scal, 4, cudaMemcpyDeviceToHost, st), "reading the panel range");
    CNVB_CUDA(ctx, cudaStreamSynchronize(st), "reading the panel range");
    // |x - c| <= 2 absmax; map that to about 2^13 so that fp16 keeps 11 significant bits
    int ex = 0;
    frexpf(absmax > 0.0f ? 2.0f * absmax : 1.0f, &ex);
    // (resident sweep: 2^11, so that the norms n'_j / 2^15 and the rows' -2 h stay inside the fp16 range)
    const float scale = ldexpf(1.0f, (res ? 11 : 13) - ex);
    wis_cmean_kernel<<<(S + 127) / 128, 128, 0, st>>>(colsum, T, S, cmean);
    CNVB_LAUNCHED(ctx, "wis_cmean_kernel");
    // sort the targets by their projection on the all-ones direction
    wis_proj_kernel<<<(T + 7) / 8, 256, 0, st>>>(dX, T, S, cmean, scale, pkey, ident, scal + 2);
    CNVB_LAUNCHED(ctx, "wis_proj_kernel");
    CNVB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(cub_tmp, sort_bytes, pkey, pkey_sorted, ident, perm, (int)T, 0, 32, st),
              "sorting the projections");
    ctx->launches += 1;
    if (Tpad > T) cudaMemsetAsync(H + (size_t)T * Kp, 0, (size_t)(Tpad - T) * Kp * 2, st);
    wis_pack_kernel<<<(T + 7) / 8, 256, 0, st>>>(dX, dtid, T, S, Kp, cmean, scale, perm, pkey_sorted, H, nrm, delta,
                                                 tid_sorted, psorted, scal + 1, stream_a ? 1 : 0, HR);
    CNVB_LAUNCHED(ctx, "wis_pack_kernel");
    const double c_acc = (double)Kp / 8388608.0;  // fp32 accumulation of Kp exact products, truncating adds
    // f32 rounding of norms and of the epilogue FFMAs; resident sweep: the norm is an addend of the tensor
    // accumulation (error up to c_acc n_j), carried by the same terms of the bounds
    const double c_rnd = 8.0 / 16777216.0 + (res ? c_acc : 0.0);
    if (Tpad > T) cudaMemsetAsync(tid_sorted + T, 0xFF, (size_t)(Tpad - T) * 4, st);
    if (Tpad > T) cudaMemsetAsync(psorted + T, 0, (size_t)(Tpad - T) * 4, st);  // (n'_j = +inf there)
    wis_cols_kernel<<<(Tpad + 255) / 256, 256, 0, st>>>(nrm, delta, tid_sorted, T, Tpad, scal + 1, c_rnd, nprime, snorm, meta);
    CNVB_LAUNCHED(ctx, "wis_cols_kernel");
    wis_tile_smax_kernel<<<(n_tiles + 7) / 8, 256, 0, st>>>(snorm, n_tiles, tile_smax);
    CNVB_LAUNCHED(ctx, "wis_tile_smax_kernel");
    if (res) {
        if (Tpad > T) cudaMemsetAsync(HR + (size_t)T * Kp, 0, (size_t)(Tpad - T) * Kp * 2, st);
        wis_fold_cols_kernel<<<(Tpad + 255) / 256, 256, 0, st>>>(nprime, T, Tpad, S, H);
        CNVB_LAUNCHED(ctx, "wis_fold_cols_kernel");
    }
    wis_fill_kernel<<<(rows + 255) / 256, 256, 0, st>>>(thr, 0, rows, __builtin_inff());
    CNVB_LAUNCHED(ctx, "wis_fill_kernel");
    wis_fill_kernel<<<(rows + 255) / 256, 256, 0, st>>>(vg, 0, rows, __builtin_inff());
    CNVB_LAUNCHED(ctx, "wis_fill_kernel");
    cudaMemsetAsync(kap, 0, rows * 4, st);
    if (subset) {  // the rows of this call, in sorted order, and their fp16 rows as the A operand
        wis_rowflag_kernel<<<(T + 255) / 256, 256, 0, st>>>(perm, T, row_begin, row_end, rflag);
        CNVB_LAUNCHED(ctx, "wis_rowflag_kernel");
        CNVB_CUDA(ctx, cub::DeviceSelect::Flagged(cub_tmp, sel_bytes, cub::CountingInputIterator<uint32_t>(0), rflag, rows_a,
                                                  n_sel, (int)T, st), "selecting the rows");
        ctx->launches += 1;
        wis_gather_rows_kernel<<<rows_pad, 64, 0, st>>>(res ? HR : H, Kp, rows_a, rows, rows_pad, HA);
        CNVB_LAUNCHED(ctx, "wis_gather_rows_kernel");
    }
    float pabs = 0.0f;
    CNVB_CUDA(ctx, cudaMemcpyAsync(&pabs, scal + 2, 4, cudaMemcpyDeviceToHost, st), "reading the projection range");
    CNVB_CUDA(ctx, cudaStreamSynchronize(st), "reading the projection range");

    CUtensorMap tmap_b, tmap_a;
    CNVB_TRY(make_tmap(ctx, H, Tpad, Kp, &tmap_b));
    if (subset) CNVB_TRY(make_tmap(ctx, HA, rows_pad, Kp, &tmap_a));
    else if (res) CNVB_TRY(make_tmap(ctx, HR, Tpad, Kp, &tmap_a));  // (resident sweep: the rows' operand is its own matrix)
    else tmap_a = tmap_b;

    const size_t smem = 2 * kMaxKBlocks * kTileBytes + kStages * kTileBytes + 512 + 1024;
    // streamed sweep: the CTAs of a cluster share their rows' operand through TMA multicast
    // (CNVB_WIS_CLUSTER = 1: one CTA per block, 2 or 4: CTAs per cluster)
    const int cluster_env = [] {  // (read at every call: the tests switch it)
        const char *e = getenv("CNVB_WIS_CLUSTER");
        const int v = e ? atoi(e) : -1;
        return v == 0 ? 1 : v;
    }();
    const int cl = !stream_a ? 1 : (cluster_env == 1 || cluster_env == 2 || cluster_env == 4 ? cluster_env : kDefaultCluster);
    if (cl == 4) CNVB_TRY(make_tmap(ctx, subset ? (void *)HA : (void *)H, subset ? rows_pad : Tpad, Kp, &tmap_a, 64));  // quarter loads
    // Later rounds of the streamed sweep on transposed tiles (one M128 x N256 MMA per K16 slab, rows on the N
    // axis; see the kernel).  Measured at both ends: 300 000 x 1000 targets, sweeps of ~10 ms at full clock:
    // 40.3 against 36.5 ms (its epilogue fetches the row operands per pair); 3 000 000 x 1000, seconds under
    // the power cap: 3.40 against 3.63 s (886 TFLOP/s executed; a third less shared-memory operand traffic per
    // flop).  So it is used for large panels; CNVB_WIS_XPOSE = 0 / 1 forces the choice.
    const int xpose_env = [] {
        const char *e = getenv("CNVB_WIS_XPOSE");
        return e ? (atoi(e) != 0 ? 1 : 0) : -1;
    }();
    const bool xp = stream_a && (xpose_env >= 0 ? xpose_env == 1 : T >= 1000000u);
    auto sweep_first = cl == 4 ? wis_sweep_kernel<true, 4, false>
                     : cl == 2 ? wis_sweep_kernel<true, 2, false>
                               : wis_sweep_kernel<true, 1, false>;
    auto sweep_next = cl == 4 ? (xp ? wis_sweep_kernel<false, 4, true> : wis_sweep_kernel<false, 4, false>)
                    : cl == 2 ? (xp ? wis_sweep_kernel<false, 2, true> : wis_sweep_kernel<false, 2, false>)
                               : (xp ? wis_sweep_kernel<false, 1, true> : wis_sweep_kernel<false, 1, false>);
    CNVB_CUDA(ctx, cudaFuncSetAttribute(sweep_first, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem opt-in");
    CNVB_CUDA(ctx, cudaFuncSetAttribute(sweep_next, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem opt-in");
    BoundConsts bc{};
    {
        const double g_ref = (S + 4.0) / 16777216.0;  // the reference's own f32 rounding (sequential sum + sqrt)
        auto up = [](double v) { return std::nextafterf((float)v, INFINITY); };
        auto down = [](double v) { return std::nextafterf((float)v, -INFINITY); };
        bc.up_infl = stream_a ? up(1.0 / 1048576.0) : 0.0f;
        bc.c_acc2 = up(2.0 * c_acc);
        bc.c_rnd = up(c_rnd);
        bc.one_plus_g = up(1.0 + g_ref);
        bc.inv_1mg = up(1.0 / (1.0 - g_ref));
        bc.one_m_crnd = down(1.0 - c_rnd);
        bc.maxes = scal + 1;
    }
    SweepArgs sa{};
    sa.nprime = nprime;
    sa.tile_smax = tile_smax;
    sa.psorted = psorted;
    sa.thr = thr;
    sa.kap = kap;
    sa.rng = rng;
    sa.T = T;
    sa.rows = rows;
    sa.tid_sorted = tid_sorted;
    sa.rows_a = rows_a;
    sa.k_mmas = ((res ? S + 3u : S) + 15) / 16;
    sa.n_kblocks = Kp / kKBlock;
    sa.cand = cand;
    sa.cand_cnt = cand_cnt;
    {
        const char *e = getenv("CNVB_WIS_DBG");
        sa.dbg = e ? (uint32_t)atoi(e) : 0u;
    }
    const size_t smem_res = (size_t)(4 + kResStages) * kTileBytes + 512 + 1024;
    if (res) {
        CNVB_CUDA(ctx, cudaFuncSetAttribute(wis_sweep_res_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_res),
                  "smem opt-in");
        CNVB_CUDA(ctx, cudaFuncSetAttribute(wis_sweep_res_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_res),
                  "smem opt-in");
    }
    auto launch_sweep = [&](decltype(sweep_first) kern) -> cudaError_t {
        if (res) {
            if (kern == sweep_first)
                wis_sweep_res_kernel<true><<<row_blocks, kSweepThreads, smem_res, st>>>(tmap_a, tmap_b, sa);
            else
                wis_sweep_res_kernel<false><<<row_blocks, kSweepThreads, smem_res, st>>>(tmap_a, tmap_b, sa);
            return cudaGetLastError();
        }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)cl * row_blocks);
        cfg.blockDim = dim3(kSweepThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cl;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kern, tmap_a, tmap_b, sa);
    };
    // CNVB_WIS_TRACE = 1 (diagnostics, stderr): per round the tiles swept, the candidates the sweep left in the
    // buffers (before the compaction) and the time of the launch, each bracketed by a host synchronisation
    const bool trace = [] {
        const char *e = getenv("CNVB_WIS_TRACE");
        return e && atoi(e) != 0;
    }();
    std::vector<uint32_t> trace_cnt;
    cudaEvent_t tr_a = nullptr, tr_b = nullptr;
    if (trace) {
        cudaEventCreate(&tr_a);
        cudaEventCreate(&tr_b);
    }
    auto trace_begin = [&]() {
        if (trace) cudaEventRecord(tr_a, st);
    };
    auto trace_end = [&](const char *what, uint32_t round, unsigned long long tiles) {
        if (!trace) return;
        cudaEventRecord(tr_b, st);
        cudaEventSynchronize(tr_b);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, tr_a, tr_b);
        trace_cnt.resize(rows);
        cudaMemcpy(trace_cnt.data(), cand_cnt, (size_t)rows * 4, cudaMemcpyDeviceToHost);
        unsigned long long sum = 0, mx = 0;
        for (uint32_t c : trace_cnt) {
            sum += c;
            mx = std::max<unsigned long long>(mx, c);
        }
        fprintf(stderr, "[wis trace] round %u %-8s %8.3f ms  tiles %9llu  candidates in buffers %11llu (%.1f per row, max %llu)\n",
                round, what, ms, tiles, sum, (double)sum / rows, mx);
        if (tiles) {  // how the blocks' tile counts spread over the SMs when CTAs are dispatched in index order
            std::vector<uint4> h_rng(row_blocks);
            cudaMemcpy(h_rng.data(), rng, (size_t)row_blocks * sizeof(uint4), cudaMemcpyDeviceToHost);
            std::vector<unsigned long long> sm((size_t)ctx->sm_count, 0ull);
            unsigned long long big = 0, total = 0;
            for (const uint4 &r4 : h_rng) {
                const unsigned long long n = (r4.y - r4.x) + (r4.w - r4.z);
                big = std::max(big, n);
                total += n;
                *std::min_element(sm.begin(), sm.end()) += n;
            }
            const unsigned long long span = *std::max_element(sm.begin(), sm.end());
            fprintf(stderr, "[wis trace]          blocks %u, largest %llu tiles, in-order makespan %llu tiles = %.2f x the mean per SM\n",
                    row_blocks, big, span, (double)span * ctx->sm_count / std::max<unsigned long long>(total, 1));
        }
    };
    CompactArgs ca{};
    ca.cand = cand;
    ca.cand_cnt = cand_cnt;
    ca.thr = thr;
    ca.kap = kap;
    ca.overflow = overflow;
    ca.seen = seen;
    ca.vg = vg;
    ca.rows_a = rows_a;
    ca.meta = meta;
    ca.max_len = counters + 2;
    ca.rows = rows;
    ca.k = k;
    ca.bc = bc;
    WindowArgs wa{};
    wa.psorted = psorted;
    wa.rows_a = rows_a;
    wa.vg = vg;
    wa.overflow = overflow;
    wa.T = T;
    wa.n_tiles = n_tiles;
    wa.rows = rows;
    // first window: 2 (1 + w0) tiles; every (tile, column quarter) group stores the nearest of each four columns,
    // so 32 x tiles candidates per row, of which at least k must remain after the same-contig ones.  The window
    // is sized for 3.2 k candidates: the k-th smallest of them is then about the 9 % quantile of the window's
    // distances
    wa.w0 = std::min<uint32_t>(11, std::max<uint32_t>(2, (16 * k / 5 + 63) / 64 - 1));
    const uint32_t tiles0 = std::min<uint32_t>(n_tiles, 2 * (1 + wa.w0));
    // f32 rounding of the projections (2 x half an ulp of the largest) plus their f64 summation error
    wa.slack = pabs * (1.0f / 4194304.0f) + 1e-30f;
    wa.rng = rng;
    wa.done = done;
    wa.tiles_swept = counters;

    // round 0: a first bound per row from the window around its own position; the lists are dropped
    // again.  (Too few columns for that -- tiny panels -- : the first real round collects everything.)
    wa.round = 0;
    wa.grow = 0;
    wis_window_kernel<<<row_blocks, kTileRows, 0, st>>>(wa);
    CNVB_LAUNCHED(ctx, "wis_window_kernel");
    if (32 * tiles0 >= 5 * k / 4) {
        {
            KernelSpan span(ctx, "wis_sweep_kernel");
            trace_begin();
            CNVB_CUDA(ctx, launch_sweep(sweep_first), "launching the first sweep");
            trace_end("sweep", 0, (unsigned long long)tiles0 * row_blocks);
        }
        CNVB_LAUNCHED(ctx, "wis_sweep_kernel");
        ca.discard = 1;
        {
            KernelSpan span(ctx, "wis_compact_kernel");
            trace_begin();
            wis_compact_kernel<<<(rows + kCompactWarps - 1) / kCompactWarps, kCompactWarps * 32, 0, st>>>(ca);
            trace_end("compact", 0, 0);
        }
        CNVB_LAUNCHED(ctx, "wis_compact_kernel");
        ca.discard = 0;
    }
    // rounds 1..: windows reach 4^r (1 + w0) tiles either side of the block's own position, never
    // beyond what the rows' bounds still allow; stop when a round has nothing left to sweep
    unsigned long long tiles_before = 0;
    uint32_t half_prev = 0, half = 4 * (tiles0 / 2 ? tiles0 / 2 : 1);
    for (uint32_t round = 1;; ++round) {
        wa.round = round;
        wa.grow = half - half_prev;
        wis_window_kernel<<<row_blocks, kTileRows, 0, st>>>(wa);
        CNVB_LAUNCHED(ctx, "wis_window_kernel");
        unsigned long long tiles_now = 0;
        CNVB_CUDA(ctx, cudaMemcpyAsync(&tiles_now, counters, 8, cudaMemcpyDeviceToHost, st), "reading the round size");
        CNVB_CUDA(ctx, cudaStreamSynchronize(st), "reading the round size");
        if (tiles_now == tiles_before && round > 1) break;
        const unsigned long long tiles_prev_round = tiles_before;
        tiles_before = tiles_now;
        {
            {
                KernelSpan span(ctx, "wis_sweep_kernel");
                trace_begin();
                CNVB_CUDA(ctx, launch_sweep(sweep_next), "launching the sweep");
                trace_end("sweep", round, tiles_now - tiles_prev_round);
            }
            CNVB_LAUNCHED(ctx, "wis_sweep_kernel");
        }
        {
            KernelSpan span(ctx, "wis_compact_kernel");
            trace_begin();
            wis_compact_kernel<<<(rows + kCompactWarps - 1) / kCompactWarps, kCompactWarps * 32, 0, st>>>(ca);
            trace_end("compact", round, 0);
        }
        CNVB_LAUNCHED(ctx, "wis_compact_kernel");
        if (half_prev >= n_tiles) break;  // the last round was allowed to reach every tile
        half_prev = half;
        half *= 4;
    }
    const bool staged = (S & 3u) == 0u && (reinterpret_cast<uintptr_t>(dX) & 15u) == 0u;
    // the longest list decides the key buffer (more CTAs per SM when the lists are short, as they are)
    unsigned long long h_maxn = 0;
    CNVB_CUDA(ctx, cudaMemcpyAsync(&h_maxn, counters + 2, 8, cudaMemcpyDeviceToHost, st), "reading the longest list");
    CNVB_CUDA(ctx, cudaStreamSynchronize(st), "reading the longest list");
    const uint32_t key_cap = (uint32_t)std::min<unsigned long long>(kCandCap, (std::max<unsigned long long>(h_maxn, 64) + 1) & ~1ull);
    const size_t rsmem = (size_t)key_cap * 8 + (((size_t)S + 3) & ~(size_t)3) * 4 +
                         (staged ? (size_t)kRescoreWarps * 32 * kStageStride * 4 : 0);
    {
        auto kern = staged ? wis_rescore_kernel<true> : wis_rescore_kernel<false>;
        CNVB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem), "smem opt-in");
        KernelSpan span(ctx, "wis_rescore_kernel");
        kern<<<rows, kRescoreWarps * 32, rsmem, st>>>(dX, dtid, T, S, k, row_begin, rows, perm, rows_a, cand, cand_cnt,
                                                      overflow, key_cap, d_dist, d_idx);
    }
    CNVB_LAUNCHED(ctx, "wis_rescore_kernel");
    // rows whose candidate buffer overflowed: exact row kernel
    wis_overflow_list_kernel<<<(rows + 255) / 256, 256, 0, st>>>(overflow, cand_cnt, rows, perm, rows_a, ovf_list + 1,
                                                                 ovf_list, counters + 1);
    CNVB_LAUNCHED(ctx, "wis_overflow_list_kernel");
    uint32_t n_ovf = 0;
    unsigned long long h_cnt[2] = {0, 0};
    CNVB_CUDA(ctx, cudaMemcpyAsync(&n_ovf, ovf_list, 4, cudaMemcpyDeviceToHost, st), "reading the fallback count");
    CNVB_CUDA(ctx, cudaMemcpyAsync(h_cnt, counters, 16, cudaMemcpyDeviceToHost, st), "reading the engine counters");
    CNVB_CUDA(ctx, cudaStreamSynchronize(st), "tensor engine");
    ctx->wis_stats[1] = n_ovf;
    ctx->wis_stats[2] = h_cnt[1];
    ctx->wis_stats[3] = h_cnt[0];  // (includes the first window, swept twice)
    if (n_ovf) CNVB_TRY(run_rows_exact(ctx, dX, dtid, T, S, k, ovf_list + 1, n_ovf, row_begin, d_dist, d_idx, nullptr));
    CNVB_CUDA(ctx, cudaStreamSynchronize(st), "tensor engine");  // work space is released on return
    return 0;
}

}  // namespace
