// normalize.cu -- lib-normalize: TotalCovSum and MedianGcBinned as sm_100a kernels.
//
//   N1  run_uncorrected_normalization (lib-normalize/src/uncorrected.rs:20-96)
//       S = Stats::sum(LCV as f64)  -> double-double (2 x f64) tree reduction, fixed order
//       CV = ((lcv as f64) / S) as f32
//   N2  run_binned_correction (lib-normalize/src/correct_binned.rs:25-146)
//       key = floor(GC*100) (f32), 101 bins; median = order statistic of rank ceil(n/2) of the
//       bin's LCV values (what CKMS(1e-6).query(0.5) returns below 5e5 items per bin), found by a
//       4-pass MSB radix select run for all 101 bins at once; CV = lcv/median, CV2 = log2(CV).
#include "comm.cuh"
#include "dd.cuh"

using namespace cnvb;

namespace {

__global__ void __launch_bounds__(256) sum_f32_dd_kernel(const float *__restrict__ x, uint64_t n,
                                                         double *__restrict__ partial /* 2*grid */) {
    dd acc{0.0, 0.0};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        acc = dd_add_d(acc, (double)x[i]);
    acc = block_reduce_dd(acc);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = acc.hi;
        partial[2 * blockIdx.x + 1] = acc.lo;
    }
}

__global__ void __launch_bounds__(256) sum_dd_final_kernel(const double *__restrict__ partial, uint32_t m,
                                                           double *__restrict__ out) {
    dd acc{0.0, 0.0};
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) acc = dd_add(acc, dd{partial[2 * i], partial[2 * i + 1]});
    acc = block_reduce_dd(acc);
    if (threadIdx.x == 0) {
        out[0] = __dadd_rn(acc.hi, acc.lo);
        out[1] = acc.hi;  // the unrounded pair: what a rank contributes when the records are sharded
        out[2] = acc.lo;
    }
}

// sharded records: the ranks' (hi, lo) pairs folded in rank order -- the same on every rank
__global__ void sum_dd_ranks_kernel(const double *__restrict__ pairs, int world, double *__restrict__ out) {
    dd acc{0.0, 0.0};
    for (int r = 0; r < world; ++r) acc = dd_add(acc, dd{pairs[2 * r], pairs[2 * r + 1]});
    out[0] = __dadd_rn(acc.hi, acc.lo);
}

__global__ void norm_total_map_kernel(const float *__restrict__ lcv, const float *__restrict__ lcvsd,
                                      uint64_t n, const double *__restrict__ sum,
                                      float *__restrict__ cv, float *__restrict__ cvsd) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double s = sum[0];
    cv[i] = (float)__ddiv_rn((double)lcv[i], s);                          // uncorrected.rs:65-72
    if (lcvsd && cvsd) cvsd[i] = (float)__ddiv_rn((double)lcvsd[i], s);   // uncorrected.rs:75-81
}

// ---- N2 -------------------------------------------------------------------------------------
constexpr int kGcBins = 101;      // correct_binned.rs:36
constexpr uint32_t kMinCount = 10;  // correct_binned.rs:20
constexpr uint32_t kNoKey = 0xffffffffu;

__device__ __forceinline__ uint32_t gc_key(float gc) {  // correct_binned.rs:55-56, 96-97
    if (!isfinite(gc)) return kNoKey;
    const float k = floorf(__fmul_rn(gc, 100.0f));
    if (!(k >= 0.0f && k < (float)kGcBins)) return kNoKey;  // the reference indexes out of range
    return (uint32_t)k;
}
// order-preserving map f32 -> u32
__device__ __forceinline__ uint32_t f32_ordered(float v) {
    const uint32_t b = __float_as_uint(v);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float f32_unordered(uint32_t u) {
    const uint32_t b = u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu);
    return __uint_as_float(b);
}

struct SelectState {        // one per GC bin
    uint32_t prefix;        // high bits of the median found so far
    uint32_t rank;          // 1-based rank still to be located inside the prefix class
    uint32_t active;        // 0 if the bin has < 10 values (median missing)
};

// one radix pass: histogram of the next 8 bits among values whose higher bits match prefix[g]
template <int PASS>
__global__ void __launch_bounds__(512) gc_radix_hist_kernel(const float *__restrict__ lcv,
                                                            const float *__restrict__ gc, uint64_t n,
                                                            const SelectState *__restrict__ st,
                                                            uint32_t *__restrict__ hist /*[101][256]*/) {
    extern __shared__ uint32_t s_hist[];  // [101][256]
    for (uint32_t i = threadIdx.x; i < kGcBins * 256; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    constexpr int shift = 24 - 8 * PASS;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    // four elements per thread and step: the eight loads are in flight together (with one element per
    // step 1024 threads keep 8 KB per SM in flight and the pass runs at 1 TB/s)
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        float gv[4], lv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t i = i0 + u * stride;
            gv[u] = i < n ? gc[i] : __int_as_float(0x7fc00000);  // NaN: no key
            lv[u] = i < n ? lcv[i] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t g = gc_key(gv[u]);
            if (g == kNoKey) continue;
            const uint32_t uo = f32_ordered(lv[u]);
            if constexpr (PASS > 0) {
                const SelectState s = st[g];
                if (!s.active || (uo >> (shift + 8)) != s.prefix) continue;
            }
            atomicAdd(&s_hist[g * 256u + ((uo >> shift) & 0xffu)], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kGcBins * 256; i += blockDim.x) {
        const uint32_t v = s_hist[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// after a pass: per GC bin, pick the digit holding the wanted rank (one warp per bin)
template <int PASS>
__global__ void gc_radix_pick_kernel(uint32_t *__restrict__ hist, SelectState *__restrict__ st,
                                     float *__restrict__ medians) {
    const uint32_t g = blockIdx.x, lane = threadIdx.x;  // 32 threads
    uint32_t h[8], tot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        h[k] = hist[g * 256u + lane * 8u + k];
        tot += h[k];
    }
    // exclusive prefix of the lane totals
    uint32_t incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += y;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    SelectState s;
    if (PASS == 0) {
        s.prefix = 0;
        s.active = total >= kMinCount;          // correct_binned.rs:68
        s.rank = (total + 1) / 2;               // CKMS query(0.5): 1-based rank ceil(n/2)
    } else {
        s = st[g];
    }
    const uint32_t excl = incl - tot;
    bool mine = s.active && s.rank > excl && s.rank <= incl;
    if (mine) {
        uint32_t cum = excl;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (s.rank <= cum + h[k]) {
                s.prefix = (s.prefix << 8) | (lane * 8u + k);
                s.rank -= cum;
                break;
            }
            cum += h[k];
        }
        st[g] = s;
        if (PASS == 3) medians[g] = f32_unordered(s.prefix);
    }
    if (!s.active && lane == 0) {
        st[g] = s;
        medians[g] = f32_missing();
    }
    // clear for the next pass
#pragma unroll
    for (int k = 0; k < 8; ++k) hist[g * 256u + lane * 8u + k] = 0u;
}

__global__ void norm_gc_map_kernel(const float *__restrict__ lcv, const float *__restrict__ lcvsd,
                                   const float *__restrict__ gc, uint64_t n,
                                   const float *__restrict__ medians, float *__restrict__ cv,
                                   float *__restrict__ cv2, float *__restrict__ cvsd,
                                   uint8_t *__restrict__ flags) {
    __shared__ float s_med[kGcBins];
    for (uint32_t i = threadIdx.x; i < kGcBins; i += blockDim.x) s_med[i] = medians[i];
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t f = 0;
    float o_cv = f32_missing(), o_cv2 = f32_missing(), o_sd = f32_missing();
    const uint32_t g = gc_key(gc[i]);
    if (g != kNoKey) {
        const float median = s_med[g];
        if (!isfinite(median)) {
            f |= CNVB_N2_FEW_GCWINDOWS;  // correct_binned.rs:99-102
        } else {
            const float norm = __fdiv_rn(lcv[i], median);  // :104
            o_cv = norm;
            f |= CNVB_N2_CV;
            if (isfinite(norm)) {  // :108-117
                o_cv2 = norm == 0.0f ? f32_missing() : log2f(norm);
                f |= CNVB_N2_CV2;
            }
            if (lcvsd && cvsd) {  // :119-126
                o_sd = __fdiv_rn(lcvsd[i], median);
                f |= CNVB_N2_CVSD;
            }
        }
    }
    cv[i] = o_cv;
    cv2[i] = o_cv2;
    if (cvsd) cvsd[i] = o_sd;
    flags[i] = f;
}

}  // namespace

namespace {
int norm_total_impl(cnvb_ctx *ctx, cnvb_comm *comm, const float *lcv, const float *lcvsd, uint64_t n, float *cv,
                    float *cvsd, double *lcv_sum, int mem) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (n && (!lcv || !cv)) return fail(ctx, CNVB_ERR_INVALID, "cnvb_norm_total: null argument");
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    const bool host = mem != CNVB_MEM_DEVICE;
    const bool sd = lcvsd && cvsd;
    uint32_t grid = (uint32_t)((n + 255) / 256);
    const uint32_t maxg = (uint32_t)ctx->sm_count * 4;
    if (grid > maxg) grid = maxg;
    if (grid == 0) grid = 1;
    const size_t nn = n ? n : 1;
    const int world = comm_world(comm);
    CNVB_TRY(ensure_scratch(ctx, Carver::need({(size_t)grid * 16, 32, (size_t)world * 16, nn * 4, nn * 4, nn * 4, nn * 4})));
    Carver c(ctx->scratch);
    double *partial = c.take<double>((size_t)grid * 2);
    double *d_sum = c.take<double>(4);
    double *d_pairs = c.take<double>((size_t)world * 2);
    const float *d_lcv = lcv, *d_lcvsd = sd ? lcvsd : nullptr;
    float *d_cv = cv, *d_cvsd = sd ? cvsd : nullptr;
    if (host) {
        float *t_lcv = c.take<float>(nn), *t_sd = c.take<float>(nn);
        d_cv = c.take<float>(nn);
        d_cvsd = sd ? c.take<float>(nn) : nullptr;
        CNVB_CUDA(ctx, cudaMemcpyAsync(t_lcv, lcv, n * 4, cudaMemcpyHostToDevice, ctx->stream), "upload lcv");
        if (sd) CNVB_CUDA(ctx, cudaMemcpyAsync(t_sd, lcvsd, n * 4, cudaMemcpyHostToDevice, ctx->stream), "upload lcvsd");
        d_lcv = t_lcv;
        d_lcvsd = sd ? t_sd : nullptr;
    }
    sum_f32_dd_kernel<<<grid, 256, 0, ctx->stream>>>(d_lcv, n, partial);
    CNVB_LAUNCHED(ctx, "sum_f32_dd_kernel");
    sum_dd_final_kernel<<<1, 256, 0, ctx->stream>>>(partial, grid, d_sum);
    CNVB_LAUNCHED(ctx, "sum_dd_final_kernel");
    if (world > 1) {  // uncorrected.rs:20-49 runs over ALL records: one exchange of a (hi, lo) pair per rank
        std::vector<uint64_t> bytes((size_t)world, 16);
        CNVB_TRY(comm_all_gather_v(comm, d_sum + 1, d_pairs, bytes.data()));
        sum_dd_ranks_kernel<<<1, 1, 0, ctx->stream>>>(d_pairs, world, d_sum);
        CNVB_LAUNCHED(ctx, "sum_dd_ranks_kernel");
    }
    if (n) {
        norm_total_map_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_lcv, d_lcvsd, n, d_sum, d_cv, d_cvsd);
        CNVB_LAUNCHED(ctx, "norm_total_map_kernel");
    }
    if (host) {
        CNVB_CUDA(ctx, cudaMemcpyAsync(cv, d_cv, n * 4, cudaMemcpyDeviceToHost, ctx->stream), "download cv");
        if (sd) CNVB_CUDA(ctx, cudaMemcpyAsync(cvsd, d_cvsd, n * 4, cudaMemcpyDeviceToHost, ctx->stream), "download cvsd");
    }
    if (lcv_sum) CNVB_CUDA(ctx, cudaMemcpyAsync(lcv_sum, d_sum, 8, cudaMemcpyDeviceToHost, ctx->stream), "download sum");
    if (host || lcv_sum) CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "norm_total");
    return CNVB_OK;
}

int norm_gc_median_impl(cnvb_ctx *ctx, cnvb_comm *comm, const float *lcv, const float *lcvsd, const float *gc,
                        uint64_t n, float *cv, float *cv2, float *cvsd, uint8_t *flags, float *medians, int mem) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (n && (!lcv || !gc || !cv || !cv2 || !flags))
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_norm_gc_median: null argument");
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    const bool host = mem != CNVB_MEM_DEVICE;
    const bool sd = lcvsd && cvsd;
    const size_t nn = n ? n : 1;
    CNVB_TRY(ensure_scratch(ctx, Carver::need({kGcBins * 256 * 4, kGcBins * sizeof(SelectState), kGcBins * 4,
                                               nn * 4, nn * 4, nn * 4, nn * 4, nn * 4, nn * 4, nn})));
    Carver c(ctx->scratch);
    uint32_t *hist = c.take<uint32_t>(kGcBins * 256);
    SelectState *st = c.take<SelectState>(kGcBins);
    float *d_med = c.take<float>(kGcBins);
    const float *d_lcv = lcv, *d_lcvsd = sd ? lcvsd : nullptr, *d_gc = gc;
    float *d_cv = cv, *d_cv2 = cv2, *d_cvsd = sd ? cvsd : nullptr;
    uint8_t *d_flags = flags;
    if (host) {
        float *t_lcv = c.take<float>(nn), *t_sd = c.take<float>(nn), *t_gc = c.take<float>(nn);
        d_cv = c.take<float>(nn);
        d_cv2 = c.take<float>(nn);
        d_cvsd = sd ? c.take<float>(nn) : nullptr;
        d_flags = c.take<uint8_t>(nn);
        CNVB_CUDA(ctx, cudaMemcpyAsync(t_lcv, lcv, n * 4, cudaMemcpyHostToDevice, ctx->stream), "upload lcv");
        CNVB_CUDA(ctx, cudaMemcpyAsync(t_gc, gc, n * 4, cudaMemcpyHostToDevice, ctx->stream), "upload gc");
        if (sd) CNVB_CUDA(ctx, cudaMemcpyAsync(t_sd, lcvsd, n * 4, cudaMemcpyHostToDevice, ctx->stream), "upload lcvsd");
        d_lcv = t_lcv;
        d_gc = t_gc;
        d_lcvsd = sd ? t_sd : nullptr;
    }
    CNVB_CUDA(ctx, cudaMemsetAsync(hist, 0, kGcBins * 256 * 4, ctx->stream), "clearing histograms");
    const size_t smem = kGcBins * 256 * 4;
    static_assert(kGcBins * 256 * 4 <= 227 * 1024, "histogram must fit in shared memory");
    CNVB_CUDA(ctx, cudaFuncSetAttribute(gc_radix_hist_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem opt-in");
    CNVB_CUDA(ctx, cudaFuncSetAttribute(gc_radix_hist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem opt-in");
    CNVB_CUDA(ctx, cudaFuncSetAttribute(gc_radix_hist_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem opt-in");
    CNVB_CUDA(ctx, cudaFuncSetAttribute(gc_radix_hist_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem opt-in");
    uint32_t grid = (uint32_t)((n + 511) / 512);
    const uint32_t maxg = (uint32_t)ctx->sm_count * 2;
    if (grid > maxg) grid = maxg;
    if (grid == 0) grid = 1;
#define CNVB_RADIX_PASS(P)                                                                          \
    {                                                                                               \
        KernelSpan span(ctx, "gc_radix_hist_kernel");                                               \
        gc_radix_hist_kernel<P><<<grid, 512, smem, ctx->stream>>>(d_lcv, d_gc, n, st, hist);        \
    }                                                                                               \
    CNVB_LAUNCHED(ctx, "gc_radix_hist_kernel");                                                     \
    /* sharded records: the medians run over ALL of them (correct_binned.rs:25-75) -- the histograms  \
       of the ranks are summed, after which every rank picks the same digit */                       \
    CNVB_TRY(comm_all_reduce_u32(comm, hist, kGcBins * 256));                                        \
    gc_radix_pick_kernel<P><<<kGcBins, 32, 0, ctx->stream>>>(hist, st, d_med);                      \
    CNVB_LAUNCHED(ctx, "gc_radix_pick_kernel");
    CNVB_RADIX_PASS(0)
    CNVB_RADIX_PASS(1)
    CNVB_RADIX_PASS(2)
    CNVB_RADIX_PASS(3)
#undef CNVB_RADIX_PASS
    if (n) {
        norm_gc_map_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_lcv, d_lcvsd, d_gc, n, d_med, d_cv,
                                                                                d_cv2, d_cvsd, d_flags);
        CNVB_LAUNCHED(ctx, "norm_gc_map_kernel");
    }
    if (host) {
        CNVB_CUDA(ctx, cudaMemcpyAsync(cv, d_cv, n * 4, cudaMemcpyDeviceToHost, ctx->stream), "download cv");
        CNVB_CUDA(ctx, cudaMemcpyAsync(cv2, d_cv2, n * 4, cudaMemcpyDeviceToHost, ctx->stream), "download cv2");
        if (sd) CNVB_CUDA(ctx, cudaMemcpyAsync(cvsd, d_cvsd, n * 4, cudaMemcpyDeviceToHost, ctx->stream), "download cvsd");
        CNVB_CUDA(ctx, cudaMemcpyAsync(flags, d_flags, n, cudaMemcpyDeviceToHost, ctx->stream), "download flags");
    }
    if (medians) CNVB_CUDA(ctx, cudaMemcpyAsync(medians, d_med, kGcBins * 4, cudaMemcpyDeviceToHost, ctx->stream), "download medians");
    if (host || medians) CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "norm_gc_median");
    return CNVB_OK;
}
}  // namespace

extern "C" int cnvb_norm_total(cnvb_ctx *ctx, const float *lcv, const float *lcvsd, uint64_t n, float *cv,
                               float *cvsd, double *lcv_sum, int mem) {
    return norm_total_impl(ctx, nullptr, lcv, lcvsd, n, cv, cvsd, lcv_sum, mem);
}

extern "C" int cnvb_norm_gc_median(cnvb_ctx *ctx, const float *lcv, const float *lcvsd, const float *gc,
                                   uint64_t n, float *cv, float *cv2, float *cvsd, uint8_t *flags,
                                   float *medians, int mem) {
    return norm_gc_median_impl(ctx, nullptr, lcv, lcvsd, gc, n, cv, cv2, cvsd, flags, medians, mem);
}

extern "C" int cnvb_norm_total_sharded(cnvb_comm *comm, const float *lcv, const float *lcvsd, uint64_t n, float *cv,
                                       float *cvsd, double *lcv_sum, int mem) {
    if (!comm) return CNVB_ERR_INVALID;
    return norm_total_impl(comm->ctx, comm, lcv, lcvsd, n, cv, cvsd, lcv_sum, mem);
}

extern "C" int cnvb_norm_gc_median_sharded(cnvb_comm *comm, const float *lcv, const float *lcvsd, const float *gc,
                                           uint64_t n, float *cv, float *cv2, float *cvsd, uint8_t *flags,
                                           float *medians, int mem) {
    if (!comm) return CNVB_ERR_INVALID;
    return norm_gc_median_impl(comm->ctx, comm, lcv, lcvsd, gc, n, cv, cv2, cvsd, flags, medians, mem);
}
