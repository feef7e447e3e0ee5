// merge.cu -- lib-merge-cov on in-memory columns: N single-sample FORMAT/CV columns side by side ->
// the multi-sample record layout (lib-merge-cov/src/lib.rs:102-232: the synced reader walks the samples'
// files in lockstep and concatenates the FORMAT arrays of each record), which is also ModelInput's
// panel X[T][S] (lib-model-wis/src/lib.rs:49-62, 65-125).
//
// HBM bound, 8 B per value (4 read + 4 written).  A CTA moves a tile of 64 targets x 32 samples through
// shared memory: column reads are 256-byte runs along the targets, row writes 128-byte runs along the
// samples; the persistent grid walks the tiles target-major so that the writes of neighbouring CTAs
// fall into the same rows.
#include "common.cuh"

using namespace cnvb;

namespace {

constexpr int kTileT = 64, kTileS = 32;

__global__ void __launch_bounds__(256) merge_cov_kernel(const float *const *__restrict__ cols, const float *__restrict__ base,
                                                        uint64_t col_stride, uint32_t S, uint64_t T, float *__restrict__ X) {
    __shared__ float tile[kTileS][kTileT + 1];
    const uint64_t tiles_t = (T + kTileT - 1) / kTileT;
    const uint32_t tiles_s = (S + kTileS - 1) / kTileS;
    const uint64_t n_tiles = tiles_t * tiles_s;
    const uint32_t tx = threadIdx.x & 63u, ty = threadIdx.x >> 6;  // read: 64 targets x 4 samples per step
    const uint32_t wx = threadIdx.x & 31u, wy = threadIdx.x >> 5;  // write: 32 samples x 8 targets per step
    for (uint64_t tile_id = blockIdx.x; tile_id < n_tiles; tile_id += gridDim.x) {
        const uint64_t t0 = (tile_id / tiles_s) * kTileT;
        const uint32_t s0 = (uint32_t)(tile_id % tiles_s) * kTileS;
#pragma unroll
        for (uint32_t q = 0; q < kTileS / 4; ++q) {
            const uint32_t s = s0 + ty + 4 * q;
            const uint64_t t = t0 + tx;
            if (s < S && t < T) {
                const float *c = cols ? cols[s] : base + (uint64_t)s * col_stride;
                tile[ty + 4 * q][tx] = __ldg(c + t);
            }
        }
        __syncthreads();
#pragma unroll
        for (uint32_t q = 0; q < kTileT / 8; ++q) {
            const uint64_t t = t0 + wy + 8 * q;
            const uint32_t s = s0 + wx;
            if (s < S && t < T) X[t * S + s] = tile[wx][wy + 8 * q];
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int cnvb_merge_cov(cnvb_ctx *ctx, const float *const *cv_cols, const float *cv_sample_major, uint64_t col_stride,
                              uint32_t S, uint64_t T, float *X) {
    if (!ctx) return CNVB_ERR_INVALID;
    if ((!cv_cols && !cv_sample_major) || (cv_cols && cv_sample_major) || !X || S == 0)
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_merge_cov: pass either the column pointers or one sample-major block");
    if (T == 0) return CNVB_OK;
    if (cv_sample_major && col_stride < T) return fail(ctx, CNVB_ERR_INVALID, "cnvb_merge_cov: column stride below T");
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    const float **d_cols = nullptr;
    if (cv_cols) {  // the pointer table travels to the device (pageable source: staged before the call returns)
        cudaError_t e = pool_alloc_t(ctx, (size_t)S, &d_cols);
        if (e != cudaSuccess) return fail_cuda(ctx, e, "allocating the column table");
        e = cudaMemcpyAsync(d_cols, cv_cols, (size_t)S * sizeof(float *), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) {
            pool_release(ctx, d_cols);
            return fail_cuda(ctx, e, "uploading the column table");
        }
    }
    const uint64_t n_tiles = ((T + kTileT - 1) / kTileT) * ((S + kTileS - 1) / kTileS);
    const unsigned grid = persistent_grid(ctx, merge_cov_kernel, 256, 0, n_tiles);
    {
        KernelSpan span(ctx, "merge_cov_kernel");
        merge_cov_kernel<<<grid, 256, 0, ctx->stream>>>(d_cols, cv_sample_major, col_stride, S, T, X);
    }
    if (d_cols) pool_release(ctx, d_cols);  // stream-ordered reuse
    CNVB_LAUNCHED(ctx, "merge_cov_kernel");
    return CNVB_OK;
}
