// dd.cuh -- double-double (2 x f64) arithmetic for reproducible, effectively exact reductions.
//
// The reference sums with an exact (Shewchuk) accumulator (lib-shared/src/stats.rs:165-197) and
// returns the correctly rounded total; a ~106-bit double-double accumulation rounds to the same
// f64 whatever the reduction order, which is what makes the GPU results independent of grid shape.
#pragma once

namespace {

// ---- double-double arithmetic (no FMA contraction: explicit _rn intrinsics) -----------------
struct dd {
    double hi, lo;
};
__device__ __forceinline__ dd two_sum(double a, double b) {
    const double s = __dadd_rn(a, b);
    const double bb = __dsub_rn(s, a);
    const double e = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
    return {s, e};
}
__device__ __forceinline__ dd fast_two_sum(double a, double b) {  // |a| >= |b|
    const double s = __dadd_rn(a, b);
    return {s, __dsub_rn(b, __dsub_rn(s, a))};
}
__device__ __forceinline__ dd dd_add(dd x, dd y) {
    dd s = two_sum(x.hi, y.hi);
    s.lo = __dadd_rn(s.lo, __dadd_rn(x.lo, y.lo));
    return fast_two_sum(s.hi, s.lo);
}
__device__ __forceinline__ dd dd_add_d(dd x, double y) {
    dd s = two_sum(x.hi, y);
    s.lo = __dadd_rn(s.lo, x.lo);
    return fast_two_sum(s.hi, s.lo);
}
__device__ __forceinline__ dd dd_shfl_xor(dd v, int o) {
    return {__shfl_xor_sync(0xffffffffu, v.hi, o), __shfl_xor_sync(0xffffffffu, v.lo, o)};
}

// block-level dd reduction; result valid in thread 0
__device__ dd block_reduce_dd(dd v) {
    __shared__ double s_hi[32], s_lo[32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = dd_add(v, dd_shfl_xor(v, o));
    if (lane == 0) {
        s_hi[warp] = v.hi;
        s_lo[warp] = v.lo;
    }
    __syncthreads();
    if (warp == 0) {
        v = lane < nw ? dd{s_hi[lane], s_lo[lane]} : dd{0.0, 0.0};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = dd_add(v, dd_shfl_xor(v, o));
    }
    return v;
}

}  // namespace
