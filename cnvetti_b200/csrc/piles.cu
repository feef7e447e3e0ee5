// piles.cu -- pile collection for black-listing (lib-coverage/src/piles.rs:31-127, the masking
// pass `--mask-piles` of lib-coverage/src/lib.rs:353-419) and the FDR depth threshold
// (lib.rs:265-351).
//
// The reference materialises a per-base depth vector of the whole contig from a second pileup pass
// and walks it base by base.  A pile is a maximal set of covered bases whose gaps are at most
// pile_max_gap, its size the largest depth inside.  Both follow from the coordinate-sorted reads
// directly, without any per-base array:
//   * a read (after the pileup filter) starts a new pile iff its start lies more than pile_max_gap
//     beyond the furthest end seen so far  -> prefix maximum of the ends + head flags + prefix sum;
//   * depth only rises at read starts, and at the start of read i it is
//     (i + 1) - #{ends <= pos_i}                -> one sort of the ends + a binary search per read;
//     the pile's size is the maximum of that over its reads (atomicMax into the pile record).
// Scans, selection and the sort are CUB device primitives (plumbing); the filter, head-flag and
// emit kernels are ours.  Everything is integer work: bit-exact against the checker.
#include <cmath>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include "common.cuh"

using namespace cnvb;

struct cnvb_piles {
    cnvb_ctx *ctx = nullptr;
    uint64_t n = 0;
    uint32_t *d_start = nullptr, *d_end = nullptr, *d_size = nullptr;
};

namespace {

constexpr uint32_t kDefaultPlpMask = 0x704u;  // htslib pileup: UNMAP | SECONDARY | QCFAIL | DUP

struct PileArgs {
    const int32_t *pos, *end;
    const uint8_t *mapq;
    const uint16_t *flags;
    uint64_t n;
    uint32_t plp_mask, min_mapq, region_end;
    uint8_t *keep;
    uint32_t *stats;  // [0] reads reaching past the contig, [1] order inversions
};

// piles.rs:45-58 (the filter inside a pileup column) after htslib's own column mask
__global__ void pile_flag_kernel(PileArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const uint32_t f = a.flags[i];
    const int32_t p = a.pos[i], e = a.end[i];
    const bool keep = (f & a.plp_mask) == 0u && ((f & (0x100u | 0x400u | 0x800u | 0x200u)) == 0u) &&
                      (uint32_t)a.mapq[i] >= a.min_mapq && (!(f & 1u) || (f & 2u)) && p >= 0 && e > p;
    a.keep[i] = keep ? 1 : 0;
    if (keep && (uint32_t)e > a.region_end) atomicAdd(&a.stats[0], 1u);  // result[pos] out of bounds (piles.rs:60)
    if (i > 0 && a.pos[i - 1] > p) atomicAdd(&a.stats[1], 1u);
}

__global__ void pile_head_kernel(const uint32_t *__restrict__ pos, const uint32_t *__restrict__ pmax, uint64_t n,
                                 uint32_t gap, uint32_t *__restrict__ head) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // piles.rs:91: the pile is extended while end + pile_max_gap >= pos
    head[i] = (i == 0 || (uint64_t)pos[i] > (uint64_t)pmax[i - 1] + gap) ? 1u : 0u;
}

__device__ __forceinline__ uint64_t upper_bound_u32(const uint32_t *__restrict__ a, uint64_t n, uint32_t key) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t m = (lo + hi) >> 1;
        if (__ldg(a + m) <= key) lo = m + 1; else hi = m;
    }
    return lo;
}

// Pile ids are non-decreasing along the reads, so the 256 reads of a CTA touch at most 256
// consecutive piles: the maxima are combined in shared memory first and every (CTA, pile) pair
// issues one global atomic -- a deep-coverage contig is ONE pile, and 15 M atomics on one address
// would serialise.
__global__ void __launch_bounds__(256) pile_emit_kernel(const uint32_t *__restrict__ pos, const uint32_t *__restrict__ pmax,
                                                        const uint32_t *__restrict__ head, const uint32_t *__restrict__ pid1,
                                                        const uint32_t *__restrict__ ends_sorted, uint64_t n,
                                                        uint32_t *__restrict__ start, uint32_t *__restrict__ end,
                                                        uint32_t *__restrict__ size) {
    __shared__ uint32_t s_max[256];
    __shared__ uint32_t s_first;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    s_max[threadIdx.x] = 0u;
    if (threadIdx.x == 0) s_first = pid1[(uint64_t)blockIdx.x * blockDim.x] - 1u;
    __syncthreads();
    if (i < n) {
        const uint32_t pile = pid1[i] - 1u;
        const uint32_t p = pos[i];
        // reads 0..i have started; those whose end is <= p have ended (ties in p: the last one sees them all)
        const uint32_t depth = (uint32_t)(i + 1 - upper_bound_u32(ends_sorted, n, p));
        atomicMax(&s_max[pile - s_first], depth);
        if (head[i]) start[pile] = p;
        if (i + 1 == n || head[i + 1]) end[pile] = pmax[i];
    }
    __syncthreads();
    if (s_max[threadIdx.x]) atomicMax(&size[s_first + threadIdx.x], s_max[threadIdx.x]);
}

struct MaxOp {
    __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

void piles_free(cnvb_piles *h) {
    if (!h) return;
    pool_release(h->ctx, h->d_start);
    pool_release(h->ctx, h->d_end);
    pool_release(h->ctx, h->d_size);
    delete h;
}

}  // namespace

extern "C" int cnvb_piles_collect(cnvb_ctx *ctx, const cnvb_read_soa *reads, const cnvb_cov_opts *opts,
                                  uint32_t pile_max_gap, uint32_t region_end, cnvb_piles **out) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (!reads || !opts || !out) return fail(ctx, CNVB_ERR_INVALID, "cnvb_piles_collect: null argument");
    *out = nullptr;
    const uint64_t n = reads->n;
    if (n && (!reads->pos || !reads->end || !reads->mapq || !reads->flags))
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_piles_collect needs pos, end, mapq and flags");
    if (n >= 0x7FFFFFFFull) return fail(ctx, CNVB_ERR_INVALID, "cnvb_piles_collect: at most 2^31 - 2 reads per call");
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    cnvb_piles *h = new (std::nothrow) cnvb_piles();
    if (!h) return fail(ctx, CNVB_ERR_NOMEM, "out of host memory");
    h->ctx = ctx;
    if (n == 0) {
        *out = h;
        return CNVB_OK;
    }
    struct Guard {
        cnvb_ctx *c;
        std::vector<void *> p;
        ~Guard() {
            for (void *x : p) pool_release(c, x);
        }
    } g{ctx, {}};
    auto alloc = [&](size_t bytes, auto **dst) -> cudaError_t {
        void *p = nullptr;
        cudaError_t e = pool_alloc(ctx, bytes ? bytes : 1, &p);
        if (e == cudaSuccess) {
            g.p.push_back(p);
            *dst = static_cast<std::remove_pointer_t<decltype(dst)>>(p);
        }
        return e;
    };
    auto bail = [&](int rc) {
        piles_free(h);
        return rc;
    };
    cudaStream_t st = ctx->stream;
    cudaError_t e;
    // inputs on the device
    const int32_t *d_pos = reads->pos, *d_end = reads->end;
    const uint8_t *d_mapq = reads->mapq;
    const uint16_t *d_flags = reads->flags;
    if (reads->mem != CNVB_MEM_DEVICE) {
        int32_t *tp = nullptr, *te = nullptr;
        uint8_t *tm = nullptr;
        uint16_t *tf = nullptr;
        if ((e = alloc(n * 4, &tp)) != cudaSuccess || (e = alloc(n * 4, &te)) != cudaSuccess ||
            (e = alloc(n, &tm)) != cudaSuccess || (e = alloc(n * 2, &tf)) != cudaSuccess)
            return bail(fail_cuda(ctx, e, "allocating the read batch"));
        if ((e = cudaMemcpyAsync(tp, reads->pos, n * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
            (e = cudaMemcpyAsync(te, reads->end, n * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
            (e = cudaMemcpyAsync(tm, reads->mapq, n, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
            (e = cudaMemcpyAsync(tf, reads->flags, n * 2, cudaMemcpyHostToDevice, st)) != cudaSuccess)
            return bail(fail_cuda(ctx, e, "uploading the read batch"));
        d_pos = tp;
        d_end = te;
        d_mapq = tm;
        d_flags = tf;
    }
    uint8_t *keep = nullptr;
    uint32_t *stats = nullptr, *pos_c = nullptr, *end_c = nullptr, *pmax = nullptr, *head = nullptr, *pid1 = nullptr,
             *ends_sorted = nullptr, *d_nsel = nullptr;
    void *tmp = nullptr;
    size_t b_sel = 0, b_scan = 0, b_sum = 0, b_sort = 0;
    cub::DeviceSelect::Flagged(nullptr, b_sel, pos_c, keep, pos_c, d_nsel, (int)n, st);
    cub::DeviceScan::InclusiveScan(nullptr, b_scan, end_c, pmax, MaxOp(), (int)n, st);
    cub::DeviceScan::InclusiveSum(nullptr, b_sum, head, pid1, (int)n, st);
    cub::DeviceRadixSort::SortKeys(nullptr, b_sort, end_c, ends_sorted, (int)n, 0, 32, st);
    const size_t b_tmp = std::max(std::max(b_sel, b_scan), std::max(b_sum, b_sort)) + 256;
    if ((e = alloc(n, &keep)) != cudaSuccess || (e = alloc(16, &stats)) != cudaSuccess ||
        (e = alloc(n * 4, &pos_c)) != cudaSuccess || (e = alloc(n * 4, &end_c)) != cudaSuccess ||
        (e = alloc(n * 4, &pmax)) != cudaSuccess || (e = alloc(n * 4, &head)) != cudaSuccess ||
        (e = alloc(n * 4, &pid1)) != cudaSuccess || (e = alloc(n * 4, &ends_sorted)) != cudaSuccess ||
        (e = alloc(8, &d_nsel)) != cudaSuccess || (e = alloc(b_tmp, reinterpret_cast<uint8_t **>(&tmp))) != cudaSuccess)
        return bail(fail_cuda(ctx, e, "allocating the pile work space"));
    cudaMemsetAsync(stats, 0, 16, st);
    PileArgs a{};
    a.pos = d_pos;
    a.end = d_end;
    a.mapq = d_mapq;
    a.flags = d_flags;
    a.n = n;
    a.plp_mask = opts->plp_flag_mask ? opts->plp_flag_mask : kDefaultPlpMask;
    a.min_mapq = opts->min_mapq;
    a.region_end = region_end;
    a.keep = keep;
    a.stats = stats;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    {
        KernelSpan span(ctx, "pile_flag_kernel");
        pile_flag_kernel<<<blocks, 256, 0, st>>>(a);
    }
    if (cudaGetLastError() != cudaSuccess) return bail(fail(ctx, CNVB_ERR_CUDA, "launch of pile_flag_kernel"));
    ctx->launches += 1;
    size_t bytes = b_tmp;
    if ((e = cub::DeviceSelect::Flagged(tmp, bytes, reinterpret_cast<const uint32_t *>(d_pos), keep, pos_c, d_nsel, (int)n,
                                        st)) != cudaSuccess)
        return bail(fail_cuda(ctx, e, "selecting the pileup reads"));
    bytes = b_tmp;
    if ((e = cub::DeviceSelect::Flagged(tmp, bytes, reinterpret_cast<const uint32_t *>(d_end), keep, end_c, d_nsel, (int)n,
                                        st)) != cudaSuccess)
        return bail(fail_cuda(ctx, e, "selecting the pileup reads"));
    ctx->launches += 2;
    uint32_t h_stats[4] = {0, 0, 0, 0};
    uint32_t n_c = 0;
    if ((e = cudaMemcpyAsync(&n_c, d_nsel, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(h_stats, stats, 16, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaStreamSynchronize(st)) != cudaSuccess)
        return bail(fail_cuda(ctx, e, "counting the pileup reads"));
    if (h_stats[1])
        return bail(fail(ctx, CNVB_ERR_INVALID, "records are not coordinate-sorted (%u inversions)", h_stats[1]));
    if (h_stats[0])
        return bail(fail(ctx, CNVB_ERR_REFERENCE_PANIC,
                         "%u alignment(s) reach past the contig end; the reference panics indexing the depth "
                         "vector (lib-coverage/src/piles.rs:60)", h_stats[0]));
    if (n_c == 0) {
        *out = h;
        return CNVB_OK;
    }
    bytes = b_tmp;
    if ((e = cub::DeviceScan::InclusiveScan(tmp, bytes, end_c, pmax, MaxOp(), (int)n_c, st)) != cudaSuccess)
        return bail(fail_cuda(ctx, e, "prefix maximum of the read ends"));
    const unsigned cblocks = (unsigned)(((uint64_t)n_c + 255) / 256);
    {
        KernelSpan span(ctx, "pile_head_kernel");
        pile_head_kernel<<<cblocks, 256, 0, st>>>(pos_c, pmax, n_c, pile_max_gap, head);
    }
    bytes = b_tmp;
    if ((e = cub::DeviceScan::InclusiveSum(tmp, bytes, head, pid1, (int)n_c, st)) != cudaSuccess)
        return bail(fail_cuda(ctx, e, "numbering the piles"));
    bytes = b_tmp;
    if ((e = cub::DeviceRadixSort::SortKeys(tmp, bytes, end_c, ends_sorted, (int)n_c, 0, 32, st)) != cudaSuccess)
        return bail(fail_cuda(ctx, e, "sorting the read ends"));
    ctx->launches += 4;
    uint32_t n_piles = 0;
    if ((e = cudaMemcpyAsync(&n_piles, pid1 + (n_c - 1), 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaStreamSynchronize(st)) != cudaSuccess)
        return bail(fail_cuda(ctx, e, "counting the piles"));
    if ((e = pool_alloc_t(ctx, n_piles, &h->d_start)) != cudaSuccess || (e = pool_alloc_t(ctx, n_piles, &h->d_end)) != cudaSuccess ||
        (e = pool_alloc_t(ctx, n_piles, &h->d_size)) != cudaSuccess)
        return bail(fail_cuda(ctx, e, "allocating the piles"));
    h->n = n_piles;
    cudaMemsetAsync(h->d_size, 0, (size_t)n_piles * 4, st);
    {
        KernelSpan span(ctx, "pile_emit_kernel");
        pile_emit_kernel<<<cblocks, 256, 0, st>>>(pos_c, pmax, head, pid1, ends_sorted, n_c, h->d_start, h->d_end, h->d_size);
    }
    if (cudaGetLastError() != cudaSuccess) return bail(fail(ctx, CNVB_ERR_CUDA, "launch of pile_emit_kernel"));
    ctx->launches += 1;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return bail(fail_cuda(ctx, e, "collecting piles"));
    *out = h;
    return CNVB_OK;
}

extern "C" uint64_t cnvb_piles_count(const cnvb_piles *h) { return h ? h->n : 0; }

extern "C" int cnvb_piles_get(cnvb_piles *h, uint32_t *start, uint32_t *end, uint32_t *size, int mem) {
    if (!h) return CNVB_ERR_INVALID;
    cnvb_ctx *ctx = h->ctx;
    if (h->n == 0) return CNVB_OK;
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    const cudaMemcpyKind kind = mem == CNVB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (start) CNVB_CUDA(ctx, cudaMemcpyAsync(start, h->d_start, h->n * 4, kind, ctx->stream), "copying piles");
    if (end) CNVB_CUDA(ctx, cudaMemcpyAsync(end, h->d_end, h->n * 4, kind, ctx->stream), "copying piles");
    if (size) CNVB_CUDA(ctx, cudaMemcpyAsync(size, h->d_size, h->n * 4, kind, ctx->stream), "copying piles");
    if (mem != CNVB_MEM_DEVICE) CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "copying piles");
    return CNVB_OK;
}

extern "C" void cnvb_piles_destroy(cnvb_piles *h) { piles_free(h); }

// ---- compute_threshold (lib-coverage/src/lib.rs:265-351): pure host arithmetic ---------------------
namespace {
// statrs Poisson::ln_pmf / cdf for integral arguments
double poisson_ln_pmf(double lambda, unsigned x) { return -lambda + (double)x * std::log(lambda) - std::lgamma((double)x + 1.0); }
double poisson_cdf(double lambda, unsigned x) {
    double acc = 0.0;
    for (unsigned i = 0; i <= x; ++i) acc += std::exp(poisson_ln_pmf(lambda, i));
    return acc < 1.0 ? acc : 1.0;
}
uint32_t rust_f64_as_u32(double v) {  // saturating cast, NaN -> 0
    if (!(v > 0.0)) return 0u;
    if (v >= 4294967295.0) return 4294967295u;
    return (uint32_t)v;
}
}  // namespace

extern "C" int cnvb_pile_threshold(const uint64_t *abs_freq, uint64_t n, double fdr_cutoff, uint32_t *threshold,
                                   char *err, uint64_t err_len) {
    auto say = [&](const char *msg) {
        if (err && err_len) snprintf(err, (size_t)err_len, "%s", msg);
    };
    if (!abs_freq || !threshold) {
        say("cnvb_pile_threshold: null argument");
        return CNVB_ERR_INVALID;
    }
    constexpr unsigned k_max = 20, k_lim = k_max + 1;
    if (n < 3) {  // lib.rs:270-271
        say("No coverage above 1?");
        return CNVB_ERR_INVALID;
    }
    if (abs_freq[0] != 0) {  // lib.rs:272-273
        say("Logic error, y[0] must be == 0");
        return CNVB_ERR_INVALID;
    }
    std::vector<uint64_t> y(abs_freq, abs_freq + n);
    if (y.size() < k_lim) y.resize(k_lim, 0);
    const double y1 = (double)y[1], y2 = (double)y[2];
    const double lambda = 2.0 * y2 / y1;                                   // lib.rs:284
    if (!(lambda > 0.0) || !std::isfinite(lambda)) {                        // Poisson::new fails (lib.rs:288-293)
        say("Could not construct Poisson distribution");
        return CNVB_ERR_INVALID;
    }
    const double nn = std::exp(std::log(y1) - poisson_ln_pmf(lambda, 1));  // lib.rs:294
    double exp_fd[k_lim], ratios[k_lim];
    uint64_t obs[k_lim] = {0};
    for (unsigned x = 0; x < k_lim; ++x) exp_fd[x] = x == 0 ? 0.0 : nn * (1.0 - poisson_cdf(lambda, x));  // :298-306
    for (unsigned i = 1; i < k_max - 1; ++i)                                                               // :311-317
        for (unsigned j = i + 1; j < k_max; ++j) obs[i] += y[j];
    for (unsigned i = 0; i < k_lim; ++i) ratios[i] = obs[i] ? exp_fd[i] / (double)obs[i] : 0.0;            // :321-332
    for (unsigned i = 1; i < k_lim; ++i) {                                                                 // :336-347
        if (ratios[i] < fdr_cutoff) {
            const double k_val = (double)(i + 1);  // k = 2 ..= 20, k[chosen - 1]
            *threshold = rust_f64_as_u32(k_val + (ratios[i - 1] - fdr_cutoff) / std::ceil(ratios[i - 1] - ratios[i]));
            return CNVB_OK;
        }
    }
    say("Could not find appropriate peak threshold");
    return CNVB_ERR_INVALID;
}
