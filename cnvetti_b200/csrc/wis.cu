// wis.cu -- lib-model-wis (WISExome-style within-sample model) and lib-mod-cov on sm_100a.
//
//   W2 compute_distances      lib-model-wis/src/lib.rs:128-197
//   W3 prune_distances        lib-model-wis/src/lib.rs:200-232
//   W4 count_per_probe_calls  lib-model-wis/src/lib.rs:235-271
//   W5 write_output filters   lib-model-wis/src/lib.rs:362-377
//   M1 load_target_infos / write_mod_cov values   lib-mod-cov/src/lib.rs:76-164, 209-240
//
// The reference distance is an f32 sum accumulated over the samples IN ORDER with separate
// multiply and add roundings, then an f32 sqrt (lib.rs:157-162).  Every distance that is
// reported or compared here is computed exactly that way (__fsub_rn/__fmul_rn/__fadd_rn/
// __fsqrt_rn: no FMA contraction, no re-association), so reference-target sets and their order
// are bit-exact.  The tensor-core engine (wis_tensor.cuh) only decides WHICH pairs get that
// exact evaluation; it never contributes a reported number.
#include <algorithm>
#include <vector>

#include "comm.cuh"
#include "dd.cuh"

using namespace cnvb;

namespace {

constexpr uint32_t kMaxK = 1024;  // max_ref_targets supported by the selection kernels

// exact reference distance between two panel rows (lib.rs:157-162)
__device__ __forceinline__ float wis_ref_distance(const float *__restrict__ a, const float *__restrict__ b,
                                                  uint32_t S, bool vec4) {
    float y = 0.0f;
    if (vec4) {
        const float4 *a4 = reinterpret_cast<const float4 *>(a);
        const float4 *b4 = reinterpret_cast<const float4 *>(b);
        for (uint32_t s = 0; s < (S >> 2); ++s) {
            const float4 p = a4[s], q = __ldg(b4 + s);
            float x = __fsub_rn(p.x, q.x);
            y = __fadd_rn(y, __fmul_rn(x, x));
            x = __fsub_rn(p.y, q.y);
            y = __fadd_rn(y, __fmul_rn(x, x));
            x = __fsub_rn(p.z, q.z);
            y = __fadd_rn(y, __fmul_rn(x, x));
            x = __fsub_rn(p.w, q.w);
            y = __fadd_rn(y, __fmul_rn(x, x));
        }
    } else {
        for (uint32_t s = 0; s < S; ++s) {
            const float x = __fsub_rn(a[s], __ldg(b + s));
            y = __fadd_rn(y, __fmul_rn(x, x));
        }
    }
    return __fsqrt_rn(y);
}

// (distance, index) lexicographic order: the documented tie rule
__device__ __forceinline__ bool ent_less(float da, uint32_t ja, float db, uint32_t jb) {
    return da < db || (da == db && ja < jb);
}

// ---- exact engine: one CTA per target row ---------------------------------------------------------
// Streams every other-contig row past the target, keeps the k smallest (distance, index) pairs
// in shared memory (threshold + buffer, bitonic compaction), independent of the column order.
// Used for small panels and as the fallback of the tensor engine.
constexpr int kRowThreads = 256;
constexpr uint32_t kRowBuf = 2048;
constexpr uint32_t kRowChunk = 1024;  // columns per pass: at most this many appends

__device__ void bitonic_sort_pairs(float *d, uint32_t *j, uint32_t n /* power of two */) {
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const uint32_t lo = (t / stride) * (stride << 1) + (t % stride), hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const float dl = d[lo], dh = d[hi];
                const uint32_t jl = j[lo], jh = j[hi];
                const bool swap = up ? ent_less(dh, jh, dl, jl) : ent_less(dl, jl, dh, jh);
                if (swap) {
                    d[lo] = dh;
                    d[hi] = dl;
                    j[lo] = jh;
                    j[hi] = jl;
                }
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kRowThreads) wis_row_exact_kernel(
    const float *__restrict__ X, const uint32_t *__restrict__ tid, uint32_t T, uint32_t S, uint32_t k,
    const uint32_t *__restrict__ row_list, uint32_t row_begin, float *__restrict__ out_dist,
    uint32_t *__restrict__ out_idx, uint32_t *__restrict__ nan_flag) {
    extern __shared__ float s_xi[];  // [S]
    __shared__ float s_d[kRowBuf];
    __shared__ uint32_t s_j[kRowBuf];
    __shared__ uint32_t s_cnt;
    __shared__ float s_thr;
    const uint32_t i = row_list ? row_list[blockIdx.x] : row_begin + blockIdx.x;
    const uint32_t ti = tid[i];
    const bool vec4 = (S & 3u) == 0u && (reinterpret_cast<uintptr_t>(X) & 15u) == 0u;
    for (uint32_t s = threadIdx.x; s < S; s += blockDim.x) s_xi[s] = X[(size_t)i * S + s];
    if (threadIdx.x == 0) {
        s_cnt = 0;
        s_thr = __int_as_float(0x7f800000);  // +inf
    }
    __syncthreads();
    for (uint32_t base = 0; base < T; base += kRowChunk) {
        const float thr = s_thr;
#pragma unroll
        for (uint32_t q = 0; q < kRowChunk / kRowThreads; ++q) {
            const uint32_t j = base + q * kRowThreads + threadIdx.x;
            if (j < T && tid[j] != ti) {  // same contig (and j == i): +inf in the reference
                const float d = wis_ref_distance(s_xi, X + (size_t)j * S, S, vec4);
                // a NaN distance: the reference's partial_cmp().unwrap() panics (lib.rs:168)
                if (d != d && nan_flag) *nan_flag = 1u;
                // (a computed +inf -- an infinite CV, or an overflowing sum -- ranks with the same-contig +inf
                // entries by index: left to the fill-up below)
                if (d <= thr && d < __int_as_float(0x7f800000)) {
                    const uint32_t slot = atomicAdd(&s_cnt, 1u);
                    s_d[slot] = d;  // slot < kRowBuf: the buffer is compacted below kRowChunk first
                    s_j[slot] = j;
                }
            }
        }
        __syncthreads();
        // Block-uniform decision: every thread takes its snapshot of the counter between two barriers, so no
        // warp can start the next chunk's appends while another still evaluates the condition (the branch
        // contains barriers of its own).
        const uint32_t n_now = s_cnt;
        __syncthreads();
        if (n_now > kRowBuf - kRowChunk) {  // next pass could overflow: keep the k smallest
            const uint32_t n = n_now;
            for (uint32_t t = n + threadIdx.x; t < kRowBuf; t += blockDim.x) {
                s_d[t] = __int_as_float(0x7f800000);
                s_j[t] = 0xffffffffu;
            }
            bitonic_sort_pairs(s_d, s_j, kRowBuf);
            if (threadIdx.x == 0) {
                s_cnt = n < k ? n : k;
                if (n >= k) s_thr = s_d[k - 1];
            }
            __syncthreads();
        }
    }
    const uint32_t n = s_cnt;
    for (uint32_t t = n + threadIdx.x; t < kRowBuf; t += blockDim.x) {
        s_d[t] = __int_as_float(0x7f800000);
        s_j[t] = 0xffffffffu;
    }
    bitonic_sort_pairs(s_d, s_j, kRowBuf);
    const size_t o = (size_t)(i - row_begin) * k;
    const uint32_t m = n < k ? n : k;
    for (uint32_t t = threadIdx.x; t < m; t += blockDim.x) {
        out_dist[o + t] = s_d[t];
        out_idx[o + t] = s_j[t];
    }
    // fewer than k finite distances: the reference's +inf entries (same contig, lib.rs:153-156, or a
    // distance that is +inf) fill up, in index order
    if (threadIdx.x == 0 && m < k) {
        uint32_t w = m;
        for (uint32_t j = 0; j < T && w < k; ++j) {
            if (tid[j] == ti || wis_ref_distance(s_xi, X + (size_t)j * S, S, vec4) == __int_as_float(0x7f800000)) {
                out_dist[o + w] = __int_as_float(0x7f800000);
                out_idx[o + w] = j;
                ++w;
            }
        }
    }
}

// ---- W3 ----------------------------------------------------------------------------------------
// threshold = (mean + z * sd) as f32 of the nearest distances (f64; exact-sum mean, n-1 SD).
// Two grid-wide reductions in double-double; the block that finishes last folds the partials.
constexpr uint32_t kThrBlocks = 128, kThrThreads = 256;

__device__ dd block_reduce_dd_all(dd v) {  // result valid in warp 0
    __shared__ double s_hi[32], s_lo[32];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = dd_add(v, dd_shfl_xor(v, o));
    __syncthreads();
    if (lane == 0) {
        s_hi[warp] = v.hi;
        s_lo[warp] = v.lo;
    }
    __syncthreads();
    v = lane < nw ? dd{s_hi[lane], s_lo[lane]} : dd{0.0, 0.0};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = dd_add(v, dd_shfl_xor(v, o));
    return v;
}

// PASS 0: sum of x -> ws[0] = mean.  PASS 1: sum of (x - mean)^2 -> out[0] = threshold.
// ws: [0] mean | [1] ticket (as u64) | [2 ..] per-block partials (hi, lo)
template <int PASS>
__global__ void __launch_bounds__(kThrThreads) wis_threshold_kernel(const float *__restrict__ nearest, uint32_t T,
                                                                    uint32_t stride, double z, double *__restrict__ ws,
                                                                    float *__restrict__ out) {
    __shared__ bool s_last;
    const double mean = PASS ? ws[0] : 0.0;
    dd acc{0.0, 0.0};
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
        double x = (double)nearest[(size_t)t * stride];
        if (PASS) {
            x = __dsub_rn(x, mean);
            x = __dmul_rn(x, x);
        }
        acc = dd_add_d(acc, x);
    }
    acc = block_reduce_dd_all(acc);
    unsigned long long *ticket = reinterpret_cast<unsigned long long *>(ws + 1);
    if (threadIdx.x == 0) {
        ws[2 + 2 * blockIdx.x] = acc.hi;
        ws[3 + 2 * blockIdx.x] = acc.lo;
        __threadfence();
        s_last = atomicAdd(ticket, 1ull) == (unsigned long long)(gridDim.x * (PASS + 1)) - 1ull;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    dd tot{0.0, 0.0};
    for (uint32_t b = threadIdx.x; b < gridDim.x; b += blockDim.x)
        tot = dd_add(tot, dd{((volatile double *)ws)[2 + 2 * b], ((volatile double *)ws)[3 + 2 * b]});
    tot = block_reduce_dd_all(tot);
    if (threadIdx.x == 0) {
        if (!PASS) {
            ws[0] = __ddiv_rn(__dadd_rn(tot.hi, tot.lo), (double)T);  // Stats::mean
        } else {
            const double var = __ddiv_rn(__dadd_rn(tot.hi, tot.lo), (double)(T - 1));  // stats.rs:231-232
            const double sd = __dsqrt_rn(var);
            out[0] = (float)__dadd_rn(mean, __dmul_rn(z, sd));  // lib.rs:216
        }
    }
}

// one warp per row: keep the entries with distance <= threshold, in order (lib.rs:221-228)
__global__ void __launch_bounds__(256) wis_prune_kernel(const float *__restrict__ dist, uint32_t *__restrict__ idx,
                                                        uint32_t rows, uint32_t k, float thr,
                                                        uint32_t *__restrict__ ref_len) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (r >= rows) return;
    uint32_t m = 0;
    for (uint32_t base = 0; base < k; base += 32) {
        const uint32_t q = base + lane;
        const bool keep = q < k && dist[(size_t)r * k + q] <= thr;
        const uint32_t j = q < k ? idx[(size_t)r * k + q] : 0u;
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) idx[(size_t)r * k + m + __popc(mask & ((1u << lane) - 1u))] = j;  // target slot <= q
        m += __popc(mask);
        __syncwarp();
    }
    if (lane == 0) ref_len[r] = m;
}

// ---- W4 ----------------------------------------------------------------------------------------
// A CTA takes kCallsGroup targets; its threads walk the flattened (target, sample) items so that all
// lanes work whatever S is, and neighbouring lanes read neighbouring samples of the same reference row.
// Mean = correctly rounded exact sum / m (Stats::sum, stats.rs:165-197): when the addends' exponents
// lie close enough together a plain f64 sum of m f32 values is exact, otherwise double-double.
constexpr uint32_t kCallsGroup = 32, kCallsThreads = 256;

__global__ void __launch_bounds__(kCallsThreads) wis_calls_kernel(const float *__restrict__ X, uint32_t S,
                                                                  const uint32_t *__restrict__ ref_idx,
                                                                  const uint32_t *__restrict__ ref_len, uint32_t k,
                                                                  uint32_t min_ref, double zthr, double rel_thr,
                                                                  uint32_t row_begin, uint32_t rows,
                                                                  uint32_t *__restrict__ num_calls) {
    extern __shared__ uint32_t s_ref[];  // [kCallsGroup][k]
    __shared__ uint32_t s_len[kCallsGroup], s_calls[kCallsGroup];
    const uint32_t r0 = blockIdx.x * kCallsGroup;
    const uint32_t groups = min(kCallsGroup, rows - r0);
    for (uint32_t g = threadIdx.x; g < kCallsGroup; g += blockDim.x) {
        s_len[g] = g < groups ? ref_len[r0 + g] : 0u;
        s_calls[g] = 0u;
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < groups * k; e += blockDim.x) {
        const uint32_t g = e / k, q = e - g * k;
        if (q < s_len[g]) s_ref[e] = ref_idx[(size_t)(r0 + g) * k + q];
    }
    __syncthreads();
    for (uint32_t item = threadIdx.x; item < groups * S; item += blockDim.x) {
        const uint32_t g = item / S, s = item - g * S;
        const uint32_t m = s_len[g];
        if (m < min_ref) continue;  // lib.rs:254-256
        const uint32_t *ref = s_ref + g * k;
        double sum = 0.0;
        uint32_t amax = 0u, amin = 0x7FFFFFFFu;
        for (uint32_t q = 0; q < m; ++q) {
            const float xf = __ldg(X + (size_t)ref[q] * S + s);
            const uint32_t bits = __float_as_uint(xf) & 0x7FFFFFFFu;
            amax = max(amax, bits);
            amin = min(amin, bits ? bits : 0x7FFFFFFFu);
            sum = __dadd_rn(sum, (double)xf);
        }
        // every addend is a multiple of 2^(emin - 23) and every partial sum is below 2^(emax + 1 + log2 m):
        // no rounding happened if that span fits 53 bits
        const int span = (int)(amax >> 23) - (int)max(amin >> 23, 1u);
        const int room = 29 - (32 - __clz((int)m - 1));  // 29 - ceil(log2 m)
        if (amax != 0u && span > room) {
            dd acc{0.0, 0.0};
            for (uint32_t q = 0; q < m; ++q) acc = dd_add_d(acc, (double)__ldg(X + (size_t)ref[q] * S + s));
            sum = __dadd_rn(acc.hi, acc.lo);
        }
        const double mean = __ddiv_rn(sum, (double)m);
        double v = 0.0;  // stats.rs:218-234: naive left-to-right, same order as the reference
        for (uint32_t q = 0; q < m; ++q) {
            const double d = __dsub_rn((double)__ldg(X + (size_t)ref[q] * S + s), mean);
            v = __dadd_rn(v, __dmul_rn(d, d));
        }
        const double var = m < 2 ? 0.0 : __ddiv_rn(v, (double)(m - 1));
        const double sd = __dsqrt_rn(var);
        const double x = (double)X[(size_t)(row_begin + r0 + g) * S + s];
        const double zs = __ddiv_rn(fabs(__dsub_rn(x, mean)), sd);  // lib.rs:260
        const double rel = __ddiv_rn(x, mean);                      // lib.rs:262
        if (zs > zthr && fabs(__dsub_rn(rel, 1.0)) > rel_thr) atomicAdd(&s_calls[g], 1u);  // lib.rs:264-266
    }
    __syncthreads();
    for (uint32_t g = threadIdx.x; g < groups; g += blockDim.x) num_calls[r0 + g] = s_calls[g];
}

__global__ void wis_filters_kernel(const uint32_t *__restrict__ ref_len, const uint32_t *__restrict__ num_calls,
                                   uint32_t rows, uint32_t min_ref, uint32_t max_rel, uint8_t *__restrict__ filt) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    uint8_t f = 0;
    if (ref_len[r] < min_ref) f |= CNVB_WIS_FEW_REF;        // lib.rs:362-369
    if (num_calls[r] > max_rel) f |= CNVB_WIS_UNRELIABLE;   // lib.rs:370-377
    filt[r] = f;
}

// ---- M1 ----------------------------------------------------------------------------------------
__global__ void modcov_kernel(const float *__restrict__ cv, uint32_t T, const uint8_t *__restrict__ model_filt,
                              const uint32_t *__restrict__ ref_idx, const uint32_t *__restrict__ ref_len,
                              uint32_t k, float *__restrict__ ref_mean, float *__restrict__ ref_sd,
                              float *__restrict__ cv_rel, float *__restrict__ cvz, float *__restrict__ cv2,
                              uint8_t *__restrict__ status) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    uint8_t st = 0;
    float o_mean = f32_missing(), o_sd = f32_missing(), o_rel = f32_missing(), o_z = f32_missing(),
          o_cv2 = f32_missing();
    const uint32_t m = ref_len[t];
    if (model_filt[t] == 0 && m > 0) {  // lib.rs:104-113 (PASS records only), :146 (non-empty)
        const uint32_t *ri = ref_idx + (size_t)t * k;
        dd acc{0.0, 0.0};
        for (uint32_t q = 0; q < m; ++q) acc = dd_add_d(acc, (double)__ldg(cv + ri[q]));
        const double mean = __ddiv_rn(__dadd_rn(acc.hi, acc.lo), (double)m);  // lib.rs:147
        double v = 0.0;
        for (uint32_t q = 0; q < m; ++q) {
            const double d = __dsub_rn((double)__ldg(cv + ri[q]), mean);
            v = __dadd_rn(v, __dmul_rn(d, d));
        }
        const double sd = __dsqrt_rn(m < 2 ? 0.0 : __ddiv_rn(v, (double)(m - 1)));  // lib.rs:148
        const double x = (double)cv[t];
        const double rel = __ddiv_rn(x, mean);                    // lib.rs:78-80
        const double zs = __ddiv_rn(__dsub_rn(x, mean), sd);      // lib.rs:83-85
        o_mean = (float)mean;
        o_sd = (float)sd;
        o_rel = (float)rel;
        o_z = (float)zs;
        st = CNVB_MODCOV_HAS_INFO;
        if (isfinite(rel)) {  // lib.rs:223-232
            o_cv2 = rel == 0.0 ? f32_missing() : (float)log2(rel);
            st |= CNVB_MODCOV_CV2;
        }
    }
    ref_mean[t] = o_mean;
    ref_sd[t] = o_sd;
    cv_rel[t] = o_rel;
    cvz[t] = o_z;
    cv2[t] = o_cv2;
    status[t] = st;
}


// (double)x of a NORMAL x (exponent field 1 .. 254) with integer instructions: |x| >> 3 puts exponent and mantissa
// in place in the high word (as the high half of |x| * 2^29, with the addend moving the exponent bias from 127 to
// 1023 in the same multiply-add), the low word is x << 29, the sign is OR-ed back.
__device__ __forceinline__ double widen_normal_f32(uint32_t b) {
    uint32_t hi;
    asm("mad.hi.u32 %0, %1, 0x20000000, 0x38000000;" : "=r"(hi) : "r"(b & 0x7FFFFFFFu));
    return __hiloint2double((int)(hi | (b & 0x80000000u)), (int)(b << 29));
}

// M1 for a batch of samples at once: X[T][S] holds FORMAT/CV of S samples side by side (the merge-cov
// layout); outputs are sample-major [S][T], i.e. one contiguous per-sample column each, as the per-sample
// call writes them (and as the segmentation reads them).  A CTA takes 32 targets and walks the samples in chunks
// of 32 x SPL; warp w handles targets w, w+8, .. of a chunk, a lane the samples lane, lane + 32, .. of it
// (neighbouring lanes read neighbouring samples of the same reference row: whole 128-byte lines), results are
// transposed through shared memory so that the stores run along the target axis.  Arithmetic and its order
// are those of modcov_kernel.
//
// Staged forms: the values of a chunk's samples in the target's reference targets are copied into shared memory
// ONCE and the passes (sum, the double-double fallback, variance) read them from there.  At config 5 the unstaged
// kernel was bound by HBM (ncu: 3.2 TB read in 0.64 s = 5.1 TB/s, 13 % L2 hits): every pass fetched the 100 rows'
// segments again, 2.7 x the 1.2 TB that have to move once.  A warp owns k x 32 x SPL floats.
//   kPanelAsync16  cp.async of 16 bytes per lane: a warp instruction moves two (SPL = 1: four) rows' segments.  Needs
//                  16-byte aligned segments: S % 4 == 0.  The default where it applies.
//   kPanelAsync    cp.async of 4 bytes per lane, sample and row (any S).
//   kPanelBulk     one bulk copy (cp.async.bulk) per reference row and warp, counted by the warp's mbarrier.  The
//                  copy takes uniform operands, so the lanes' copies are issued one at a time behind R2UR (ncu: 61
//                  cycles per copy, 28 % of the warps' time) -- measured slower than the 16-byte cp.async, kept
//                  under test (CNVB_PANEL_STAGE=2).
// A chunk is 64 samples wide at the reference's default k: two independent f64 chains per lane.  The eight warps an
// SM holds (shared memory: 205 KB of values at k = 100) leave the loops bound by latency and by the conversion
// pipe: F2F.F64.F32 issues once per 8 cycles and SM quarter (ncu: the loops advance one reference target per 36 - 46
// cycles, two warps x two conversions x 8).  The odd columns of a lane are therefore widened f32 -> f64 with integer
// instructions (same bits for normal numbers: sign, exponent + 896, mantissa << 29; a column in which the first
// pass sees a zero, a subnormal, an infinity or a NaN is redone with F2F), the even ones on the conversion pipe.
// SPL = 2 up to k = 105 (205 KB at k = 100), SPL = 1 up to k = 199, the unstaged form beyond.  The staging area is
// reused for the transposed outputs of the chunk.  Arithmetic and its order are unchanged: identical bits.
enum PanelMode { kPanelPlain = 0, kPanelAsync = 1, kPanelBulk = 2, kPanelAsync16 = 3 };
constexpr uint32_t kPanelGroup = 32, kPanelThreads = 256, kPanelWarps = kPanelThreads / 32;
constexpr size_t panel_out_bytes(uint32_t spl) { return (size_t)5 * 32 * spl * (kPanelGroup + 1) * sizeof(float); }

struct PanelOut {
    float *ref_mean, *ref_sd, *cv_rel, *cvz, *cv2;
    uint8_t *status;
};

// CALLS: the per-probe call counts of build-model-wis (count_per_probe_calls, lib.rs:235-271) come out of the
// same pass -- both stages gather the same reference rows and form the same mean / sd, and at cohort scale
// the gathers (1.2 TB at config 5) are all there is.  The model filters are not known yet in that case
// (UNRELIABLE depends on the counts): values are written for every target with enough reference targets
// and modcov_mask_kernel blanks the filtered ones afterwards.
struct PanelCalls {
    uint32_t min_ref;
    double zthr, rel_thr;
    uint32_t *num_calls;
};

template <bool CALLS, int MODE, int SPL>
__global__ void __launch_bounds__(kPanelThreads)
    modcov_panel_kernel(const float *__restrict__ X, uint32_t S, uint32_t row_begin, uint32_t rows,
                        const uint8_t *__restrict__ model_filt, const uint32_t *__restrict__ ref_idx,
                        const uint32_t *__restrict__ ref_len, uint32_t k, PanelOut o, PanelCalls pc) {
    constexpr bool STAGED = MODE != kPanelPlain;
    constexpr uint32_t CH = 32 * SPL;                         // samples of a chunk
    constexpr uint32_t NIT = kPanelGroup / kPanelWarps;       // targets of a warp per chunk
    extern __shared__ __align__(128) uint32_t s_ref[];        // [kPanelGroup][k], then (staged) the values [warp][k][CH]
    __shared__ uint32_t s_len[kPanelGroup], s_calls[kPanelGroup];
    __shared__ float s_out_own[STAGED ? 1 : 5 * CH * (kPanelGroup + 1)];
    __shared__ uint8_t s_st[CH][kPanelGroup + 4];
    __shared__ __align__(8) uint64_t s_bar[kPanelWarps];
    float *const stage = reinterpret_cast<float *>(s_ref + kPanelGroup * k);
    // [value][sample in chunk][target]; staged: over the staged values, which are dead once the chunk's targets are done
    float(*const s_out)[CH][kPanelGroup + 1] =
        reinterpret_cast<float(*)[CH][kPanelGroup + 1]>(STAGED ? stage : s_out_own);
    const uint32_t r0 = blockIdx.x * kPanelGroup;  // relative to row_begin, like every per-row argument
    const uint32_t groups = min(kPanelGroup, rows - r0);
    for (uint32_t g = threadIdx.x; g < kPanelGroup; g += blockDim.x) {
        if (CALLS) {  // lib.rs:254-256: targets with fewer than min_ref_targets reference targets take no part
            const uint32_t m = g < groups ? ref_len[r0 + g] : 0u;
            s_len[g] = (m >= pc.min_ref) ? m : 0u;
            s_calls[g] = 0u;
        } else {
            s_len[g] = (g < groups && model_filt[r0 + g] == 0) ? ref_len[r0 + g] : 0u;  // lib.rs:104-113, :146
        }
    }
    if (MODE == kPanelBulk && threadIdx.x == 0) {
        for (uint32_t w = 0; w < kPanelWarps; ++w)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_bar[w])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < groups * k; e += blockDim.x) {
        const uint32_t g = e / k, q = e - g * k;
        if (q < s_len[g]) s_ref[e] = ref_idx[(size_t)(r0 + g) * k + q];
    }
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const float *const mine = stage + (size_t)warp * k * CH + lane;  // this lane's first column of the warp's values
    const uint32_t mine_s = (uint32_t)__cvta_generic_to_shared(mine);
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar[warp]);
    uint32_t phase = 0;
    for (uint32_t s0 = 0; s0 < S; s0 += CH) {
        // a warp's targets of this chunk; the results wait in registers because the transposition buffer lies
        // over the staged values of the other warps
        float res[NIT][SPL][5];
        uint8_t res_st[NIT][SPL];
        uint32_t sm[SPL];  // the lane's samples; past the end of the panel (last chunk): computed on whatever the
                           // staging area holds (unstaged: on the last sample) and discarded
#pragma unroll
        for (int j = 0; j < SPL; ++j) sm[j] = s0 + 32 * j + lane;
#pragma unroll
        for (uint32_t it = 0; it < NIT; ++it) {
            const uint32_t g = warp + it * kPanelWarps;
            if (g >= groups) continue;
            const uint32_t m = s_len[g];
            const uint32_t *ref = s_ref + g * k;
#pragma unroll
            for (int j = 0; j < SPL; ++j) {
#pragma unroll
                for (int v = 0; v < 5; ++v) res[it][j][v] = f32_missing();
                res_st[it][j] = 0;
            }
            double o_zd[SPL], o_reld[SPL];  // CALLS: the f64 values behind cvz / cv_rel
#pragma unroll
            for (int j = 0; j < SPL; ++j) o_zd[j] = 0.0, o_reld[j] = 1.0;
            if (m > 0) {
                if (MODE == kPanelBulk) {
                    const uint32_t bytes = min(CH, S - s0) * 4u;
                    if (lane == 0)
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(m * bytes) : "memory");
                    __syncwarp();  // every lane is done with the previous target's values
                    for (uint32_t q = lane; q < m; q += 32)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                         mine_s - lane * 4u + q * (CH * 4u)),
                                     "l"(X + (size_t)ref[q] * S + s0), "r"(bytes), "r"(bar)
                                     : "memory");
                    uint32_t ok;
                    do {
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t"
                            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                            "selp.u32 %0, 1, 0, p;\n\t}"
                            : "=r"(ok)
                            : "r"(bar), "r"(phase)
                            : "memory");
                    } while (!ok);
                    phase ^= 1u;
                } else if (MODE == kPanelAsync16) {
                    constexpr uint32_t LPR = CH / 4, RPI = 32 / LPR;  // lanes per row segment, rows per instruction
                    const uint32_t col = (lane % LPR) * 4u;
                    __syncwarp();  // every lane is done with the previous target's values
                    if (col < min(CH, S - s0))
                        for (uint32_t q = lane / LPR; q < m; q += RPI)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(mine_s - lane * 4u + (q * CH + col) * 4u),
                                         "l"(X + (size_t)ref[q] * S + s0 + col)
                                         : "memory");
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    __syncwarp();  // the lanes read one another's copies
                } else if (MODE == kPanelAsync) {
                    for (uint32_t q = 0; q < m; ++q) {
#pragma unroll
                        for (int j = 0; j < SPL; ++j)
                            if (sm[j] < S)
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(mine_s + (q * CH + 32 * j) * 4u),
                                             "l"(X + (size_t)ref[q] * S + sm[j])
                                             : "memory");
                    }
                    asm volatile("cp.async.wait_all;" ::: "memory");  // a lane reads back its own copies only
                }
                auto xat = [&](uint32_t q, int j) -> float {
                    return STAGED ? mine[q * CH + 32 * j] : __ldg(X + (size_t)ref[q] * S + min(sm[j], S - 1));
                };
                // exact sum (Stats::sum): a plain f64 sum of m f32 values is exact when their exponents lie
                // close together (see wis_calls_kernel); double-double otherwise
                double sum[SPL];
                uint32_t amax[SPL], amin[SPL];  // of |x| as integers; amin over ALL values here, zeros included
#pragma unroll
                for (int j = 0; j < SPL; ++j) sum[j] = 0.0, amax[j] = 0u, amin[j] = 0x7FFFFFFFu;
#pragma unroll 4
                for (uint32_t q = 0; q < m; ++q) {
#pragma unroll
                    for (int j = 0; j < SPL; ++j) {
                        const float xf = xat(q, j);
                        const uint32_t bits = __float_as_uint(xf) & 0x7FFFFFFFu;
                        amax[j] = max(amax[j], bits);
                        amin[j] = min(amin[j], bits);
                        // odd columns widen on the integer pipes, even ones on the conversion pipe (redone below
                        // if the column turns out to hold a zero, a subnormal, an infinity or a NaN)
                        sum[j] = __dadd_rn(sum[j], (STAGED && (j & 1)) ? widen_normal_f32(__float_as_uint(xf)) : (double)xf);
                    }
                }
                const int room = 29 - (32 - __clz((int)m - 1));  // 29 - ceil(log2 m)
                double mean[SPL], v[SPL];
                bool normal = true;  // only normal numbers in the lane's odd columns: exponent fields 1 .. 254
#pragma unroll
                for (int j = 0; j < SPL; ++j) {
                    const bool normal_j = !(j & 1) || (amin[j] >= 0x00800000u && amax[j] < 0x7F800000u);
                    normal = normal && normal_j;
                    if (STAGED && !normal_j) {
                        sum[j] = 0.0;
                        for (uint32_t q = 0; q < m; ++q) sum[j] = __dadd_rn(sum[j], (double)xat(q, j));
                    }
                    if (amin[j] == 0u) {  // the span is taken over the non-zero values
                        amin[j] = 0x7FFFFFFFu;
                        for (uint32_t q = 0; q < m; ++q) {
                            const uint32_t bits = __float_as_uint(xat(q, j)) & 0x7FFFFFFFu;
                            amin[j] = min(amin[j], bits ? bits : 0x7FFFFFFFu);
                        }
                    }
                    const int span = (int)(amax[j] >> 23) - (int)max(amin[j] >> 23, 1u);
                    if (amax[j] != 0u && span > room) {
                        dd acc{0.0, 0.0};
                        for (uint32_t q = 0; q < m; ++q) acc = dd_add_d(acc, (double)xat(q, j));
                        sum[j] = __dadd_rn(acc.hi, acc.lo);
                    }
                    mean[j] = __ddiv_rn(sum[j], (double)m);  // lib.rs:147
                    v[j] = 0.0;
                }
                if (STAGED && normal) {
#pragma unroll 4
                    for (uint32_t q = 0; q < m; ++q) {
#pragma unroll
                        for (int j = 0; j < SPL; ++j) {
                            const float xf = xat(q, j);
                            const double d = __dsub_rn((j & 1) ? widen_normal_f32(__float_as_uint(xf)) : (double)xf, mean[j]);
                            v[j] = __dadd_rn(v[j], __dmul_rn(d, d));
                        }
                    }
                } else {
#pragma unroll 4
                    for (uint32_t q = 0; q < m; ++q) {
#pragma unroll
                        for (int j = 0; j < SPL; ++j) {
                            const double d = __dsub_rn((double)xat(q, j), mean[j]);
                            v[j] = __dadd_rn(v[j], __dmul_rn(d, d));
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < SPL; ++j) {
                    if (sm[j] >= S) continue;
                    const double sd = __dsqrt_rn(m < 2 ? 0.0 : __ddiv_rn(v[j], (double)(m - 1)));  // lib.rs:148
                    const double x = (double)X[(size_t)(row_begin + r0 + g) * S + sm[j]];
                    const double rel = __ddiv_rn(x, mean[j]);                // lib.rs:78-80
                    const double zs = __ddiv_rn(__dsub_rn(x, mean[j]), sd);  // lib.rs:83-85
                    res[it][j][0] = (float)mean[j];
                    res[it][j][1] = (float)sd;
                    res[it][j][2] = (float)rel;
                    res[it][j][3] = (float)zs;
                    o_zd[j] = zs;
                    o_reld[j] = rel;
                    uint8_t st = CNVB_MODCOV_HAS_INFO;
                    if (isfinite(rel)) {  // lib.rs:223-232
                        res[it][j][4] = rel == 0.0 ? f32_missing() : (float)log2(rel);
                        st |= CNVB_MODCOV_CV2;
                    }
                    res_st[it][j] = st;
                }
            }
            if (CALLS) {  // lib.rs:260-266 (|x - mean| / sd = |(x - mean) / sd| bit for bit)
                uint32_t n_hit = 0;
#pragma unroll
                for (int j = 0; j < SPL; ++j) {
                    const bool hit = (res_st[it][j] & CNVB_MODCOV_HAS_INFO) && fabs(o_zd[j]) > pc.zthr &&
                                     fabs(__dsub_rn(o_reld[j], 1.0)) > pc.rel_thr;
                    n_hit += __popc(__ballot_sync(0xffffffffu, hit));
                }
                if (lane == 0 && n_hit) atomicAdd(&s_calls[g], n_hit);
            }
        }
        if (STAGED) __syncthreads();
#pragma unroll
        for (uint32_t it = 0; it < NIT; ++it) {
            const uint32_t g = warp + it * kPanelWarps;
            if (g >= groups) continue;
#pragma unroll
            for (int j = 0; j < SPL; ++j) {
#pragma unroll
                for (int v = 0; v < 5; ++v) s_out[v][32 * j + lane][g] = res[it][j][v];
                s_st[32 * j + lane][g] = res_st[it][j];
            }
        }
        __syncthreads();
        float *const outs[5] = {o.ref_mean, o.ref_sd, o.cv_rel, o.cvz, o.cv2};
        for (uint32_t e = threadIdx.x; e < CH * kPanelGroup; e += blockDim.x) {
            const uint32_t sl = e / kPanelGroup, g = e % kPanelGroup;
            if (s0 + sl < S && g < groups) {
                const size_t at = (size_t)(s0 + sl) * rows + r0 + g;
#pragma unroll
                for (int v = 0; v < 5; ++v)
                    if (outs[v]) outs[v][at] = s_out[v][sl][g];
                if (o.status) o.status[at] = s_st[sl][g];
            }
        }
        // the next chunk's bulk copies (async proxy) overwrite what was written and read here (generic proxy)
        if (MODE == kPanelBulk) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
    if (CALLS)
        for (uint32_t g = threadIdx.x; g < groups; g += blockDim.x) pc.num_calls[r0 + g] = s_calls[g];
}

// after the fused pass: targets whose model record carries a filter have no probe info (lib.rs:104-113)
__global__ void modcov_mask_kernel(const uint8_t *__restrict__ filt, uint32_t rows, uint64_t n, PanelOut o) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || filt[i % rows] == 0) return;
    float *const outs[5] = {o.ref_mean, o.ref_sd, o.cv_rel, o.cvz, o.cv2};
#pragma unroll
    for (int v = 0; v < 5; ++v)
        if (outs[v]) outs[v][i] = f32_missing();
    if (o.status) o.status[i] = 0;
}

}  // namespace

#include "wis_tensor.cuh"

// ---- host side ------------------------------------------------------------------------------------
namespace {

// device view of a caller buffer: uploads host memory into pooled device memory
template <typename T>
struct DevIn {
    cnvb_ctx *ctx;
    const T *p = nullptr;
    T *owned = nullptr;
    int init(cnvb_ctx *c, const T *src, size_t n, int mem, const char *what) {
        ctx = c;
        if (mem == CNVB_MEM_DEVICE || !src) {
            p = src;
            return 0;
        }
        cudaError_t e = pool_alloc_t(ctx, n, &owned);
        if (e != cudaSuccess) return fail_cuda(ctx, e, what);
        e = cudaMemcpyAsync(owned, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) return fail_cuda(ctx, e, what);
        p = owned;
        return 0;
    }
    ~DevIn() {
        if (owned) pool_release(ctx, owned);
    }
};
template <typename T>
struct DevOut {
    cnvb_ctx *ctx;
    T *p = nullptr;
    T *host = nullptr;
    T *owned = nullptr;
    size_t n = 0;
    int init(cnvb_ctx *c, T *dst, size_t count, int mem, const char *what) {
        ctx = c;
        n = count;
        if (mem == CNVB_MEM_DEVICE || !dst) {
            p = dst;
            return 0;
        }
        cudaError_t e = pool_alloc_t(ctx, count, &owned);
        if (e != cudaSuccess) return fail_cuda(ctx, e, what);
        p = owned;
        host = dst;
        return 0;
    }
    int download() {  // enqueue the copy back (caller synchronises)
        if (host && n)
            CNVB_CUDA(ctx, cudaMemcpyAsync(host, owned, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream),
                      "copying results to the host");
        return 0;
    }
    ~DevOut() {
        if (owned) pool_release(ctx, owned);
    }
};

int check_wis_opts(cnvb_ctx *ctx, const cnvb_wis_opts *o, uint32_t T) {
    if (!o) return fail(ctx, CNVB_ERR_INVALID, "null options");
    if (o->max_ref_targets == 0)
        return fail(ctx, CNVB_ERR_REFERENCE_PANIC, "max_ref_targets == 0: the reference panics at xs[0] (lib.rs:208)");
    const uint32_t k = T < o->max_ref_targets ? T : o->max_ref_targets;
    if (k > kMaxK) return fail(ctx, CNVB_ERR_INVALID, "max_ref_targets > %u is not supported", kMaxK);
    return 0;
}

int run_rows_exact(cnvb_ctx *ctx, const float *dX, const uint32_t *dtid, uint32_t T, uint32_t S, uint32_t k,
                   const uint32_t *d_row_list, uint32_t n_rows, uint32_t row_begin, float *d_dist, uint32_t *d_idx,
                   uint32_t *d_nan_flag) {
    if (n_rows == 0) return 0;
    const size_t smem = (size_t)S * sizeof(float);
    if (smem > 200 * 1024) return fail(ctx, CNVB_ERR_INVALID, "more than 51200 samples are not supported");
    CNVB_CUDA(ctx, cudaFuncSetAttribute(wis_row_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
              "smem opt-in");
    KernelSpan span(ctx, "wis_row_exact_kernel");
    wis_row_exact_kernel<<<n_rows, kRowThreads, smem, ctx->stream>>>(dX, dtid, T, S, k, d_row_list, row_begin, d_dist,
                                                                     d_idx, d_nan_flag);
    CNVB_LAUNCHED(ctx, "wis_row_exact_kernel");
    return 0;
}

// W1: the reference's ModelInput is NaN-free by assumption -- a NaN distance makes `partial_cmp().unwrap()`
// panic (lib-model-wis/src/lib.rs:168).  One pass over the panel: the largest |x| by bit pattern (NaN patterns
// order above +inf) and whether the targets lie on more than one contig (same-contig pairs are never
// evaluated, lib.rs:153-156).
__global__ void __launch_bounds__(256) wis_panel_check_kernel(const float *__restrict__ X, const uint32_t *__restrict__ tid,
                                                              uint32_t T, uint64_t n, uint32_t *__restrict__ out) {
    int mx = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(X) & 15u) == 0u) {
        const int4 *X4 = reinterpret_cast<const int4 *>(X);
        const uint64_t n4 = n >> 2;
        for (uint64_t i = g; i < n4; i += stride) {
            const int4 v = ld_stream(X4 + i);
            mx = max(max(mx, v.x & 0x7fffffff), max(max(v.y & 0x7fffffff, v.z & 0x7fffffff), v.w & 0x7fffffff));
        }
        for (uint64_t i = (n4 << 2) + g; i < n; i += stride) mx = max(mx, __float_as_int(X[i]) & 0x7fffffff);
    } else {
        for (uint64_t i = g; i < n; i += stride) mx = max(mx, __float_as_int(X[i]) & 0x7fffffff);
    }
    bool other = false;
    const uint32_t t0 = tid[0];
    for (uint64_t i = g; i < T; i += stride) other |= tid[i] != t0;
    mx = __reduce_max_sync(0xffffffffu, mx);
    other = __any_sync(0xffffffffu, other);
    if (lane_id() == 0) {
        if (mx) atomicMax(reinterpret_cast<int *>(out), mx);
        if (other) out[1] = 1u;
    }
}

int wis_distances_dev(cnvb_ctx *ctx, const float *dX, const uint32_t *dtid, uint32_t T, uint32_t S,
                      const cnvb_wis_opts *o, uint32_t row_begin, uint32_t row_end, float *d_dist, uint32_t *d_idx) {
    const uint32_t k = T < o->max_ref_targets ? T : o->max_ref_targets;
    uint32_t engine = o->engine;
    if (engine == CNVB_WIS_ENGINE_AUTO)  // the tensor engine pays off for large panels
        engine = (T >= 4096 && S <= kMaxTensorS && k <= kMaxTensorK) ? CNVB_WIS_ENGINE_TENSOR
                                                                             : CNVB_WIS_ENGINE_EXACT;
    ctx->wis_stats[0] = row_end - row_begin;
    ctx->wis_stats[1] = ctx->wis_stats[2] = ctx->wis_stats[3] = 0;
    if (row_end == row_begin) return 0;
    uint32_t *d_chk = nullptr;  // [0] max |x| pattern | [1] more than one contig | [2] NaN distance seen
    cudaError_t e = pool_alloc_t(ctx, 4, &d_chk);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "allocating the panel check");
    struct Release {
        cnvb_ctx *c;
        void *p;
        ~Release() { pool_release(c, p); }
    } release{ctx, d_chk};
    uint32_t chk[4] = {0, 0, 0, 0};
    CNVB_CUDA(ctx, cudaMemsetAsync(d_chk, 0, 16, ctx->stream), "panel check");
    {
        const uint64_t n = (uint64_t)T * S;
        const unsigned blocks = (unsigned)std::min<uint64_t>((n / 4 + 255) / 256 + 1, (uint64_t)ctx->sm_count * 8);
        wis_panel_check_kernel<<<blocks, 256, 0, ctx->stream>>>(dX, dtid, T, n, d_chk);
        CNVB_LAUNCHED(ctx, "wis_panel_check_kernel");
    }
    CNVB_CUDA(ctx, cudaMemcpyAsync(chk, d_chk, 8, cudaMemcpyDeviceToHost, ctx->stream), "panel check");
    CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "panel check");
    const bool has_nan = chk[0] > 0x7f800000u, has_inf = chk[0] == 0x7f800000u;
    if (has_nan && chk[1])
        return fail(ctx, CNVB_ERR_REFERENCE_PANIC,
                    "the panel holds a NaN (a missing FORMAT/CV): called `Option::unwrap()` on a `None` value, "
                    "dist.partial_cmp() in compute_distances (lib-model-wis/src/lib.rs:168)");
    // a non-finite panel has no fp16 image: every pair in the reference's own arithmetic (inf - inf = NaN panics there)
    if (has_nan || has_inf) engine = CNVB_WIS_ENGINE_EXACT;
    if (engine == CNVB_WIS_ENGINE_TENSOR)
        return wis_distances_tensor(ctx, dX, dtid, T, S, k, row_begin, row_end, d_dist, d_idx);
    CNVB_TRY(run_rows_exact(ctx, dX, dtid, T, S, k, nullptr, row_end - row_begin, row_begin, d_dist, d_idx,
                            has_inf ? d_chk + 2 : nullptr));
    if (has_inf) {
        CNVB_CUDA(ctx, cudaMemcpyAsync(chk + 2, d_chk + 2, 4, cudaMemcpyDeviceToHost, ctx->stream), "panel check");
        CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "panel check");
        if (chk[2])
            return fail(ctx, CNVB_ERR_REFERENCE_PANIC,
                        "two targets hold the same infinity in one sample (inf - inf = NaN): called `Option::unwrap()` "
                        "on a `None` value, dist.partial_cmp() in compute_distances (lib-model-wis/src/lib.rs:168)");
    }
    return 0;
}

}  // namespace

extern "C" int cnvb_wis_last_stats(const cnvb_ctx *ctx, uint64_t out[4]) {
    if (!ctx || !out) return CNVB_ERR_INVALID;
    for (int i = 0; i < 4; ++i) out[i] = ctx->wis_stats[i];
    return CNVB_OK;
}

extern "C" int cnvb_wis_distances(cnvb_ctx *ctx, const float *X, const uint32_t *tid, uint32_t T, uint32_t S,
                                  const cnvb_wis_opts *opts, uint32_t row_begin, uint32_t row_end, float *ref_dist,
                                  uint32_t *ref_idx, int mem) {
    if (!ctx) return CNVB_ERR_INVALID;
    CNVB_TRY(check_wis_opts(ctx, opts, T));
    if (!X || !tid || !ref_dist || !ref_idx || row_end > T || row_begin > row_end || S == 0)
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_wis_distances: bad argument");
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    const uint32_t k = T < opts->max_ref_targets ? T : opts->max_ref_targets;
    const size_t rows = row_end - row_begin;
    DevIn<float> dX;
    DevIn<uint32_t> dt;
    DevOut<float> od;
    DevOut<uint32_t> oi;
    CNVB_TRY(dX.init(ctx, X, (size_t)T * S, mem, "uploading the panel"));
    CNVB_TRY(dt.init(ctx, tid, T, mem, "uploading contig ids"));
    CNVB_TRY(od.init(ctx, ref_dist, rows * k, mem, "allocating distances"));
    CNVB_TRY(oi.init(ctx, ref_idx, rows * k, mem, "allocating indices"));
    CNVB_TRY(wis_distances_dev(ctx, dX.p, dt.p, T, S, opts, row_begin, row_end, od.p, oi.p));
    CNVB_TRY(od.download());
    CNVB_TRY(oi.download());
    if (mem != CNVB_MEM_DEVICE) CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "wis distances");
    return CNVB_OK;
}

namespace {
int wis_threshold_dev(cnvb_ctx *ctx, const float *d_nearest, uint32_t T, uint32_t stride, double z, float *thr_host) {
    if (T < 2) {  // lib.rs:211-213: prune_distances returns an empty Vec
        *thr_host = __builtin_nanf("");
        return 0;
    }
    CNVB_TRY(ensure_scratch(ctx, 8192));
    double *ws = static_cast<double *>(ctx->scratch);  // mean | ticket | block partials
    float *d_thr = reinterpret_cast<float *>(static_cast<char *>(ctx->scratch) + 4096 + 2048);
    CNVB_CUDA(ctx, cudaMemsetAsync(ws, 0, 16, ctx->stream), "threshold");
    const unsigned blocks = std::min<unsigned>(kThrBlocks, (T + kThrThreads - 1) / kThrThreads);
    wis_threshold_kernel<0><<<blocks, kThrThreads, 0, ctx->stream>>>(d_nearest, T, stride, z, ws, d_thr);
    CNVB_LAUNCHED(ctx, "wis_threshold_kernel");
    wis_threshold_kernel<1><<<blocks, kThrThreads, 0, ctx->stream>>>(d_nearest, T, stride, z, ws, d_thr);
    CNVB_LAUNCHED(ctx, "wis_threshold_kernel");
    CNVB_CUDA(ctx, cudaMemcpyAsync(thr_host, d_thr, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream), "threshold");
    CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "threshold");
    return 0;
}
}  // namespace

extern "C" int cnvb_wis_threshold(cnvb_ctx *ctx, const float *nearest, uint32_t T, uint32_t stride,
                                  double filter_z_score, float *threshold, int mem) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (!threshold || (T && !nearest) || stride == 0) return fail(ctx, CNVB_ERR_INVALID, "cnvb_wis_threshold: bad argument");
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    DevIn<float> dn;
    CNVB_TRY(dn.init(ctx, nearest, (size_t)T * stride, mem, "uploading nearest distances"));
    return wis_threshold_dev(ctx, dn.p, T, stride, filter_z_score, threshold);
}

extern "C" int cnvb_wis_prune(cnvb_ctx *ctx, const float *ref_dist, uint32_t *ref_idx, uint32_t rows, uint32_t k,
                              float threshold, uint32_t *ref_len, int mem) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (rows && (!ref_dist || !ref_idx || !ref_len)) return fail(ctx, CNVB_ERR_INVALID, "cnvb_wis_prune: null argument");
    if (rows == 0) return CNVB_OK;
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    DevIn<float> dd_;
    DevIn<uint32_t> di_in;
    DevOut<uint32_t> di, dl;
    CNVB_TRY(dd_.init(ctx, ref_dist, (size_t)rows * k, mem, "uploading distances"));
    CNVB_TRY(di.init(ctx, ref_idx, (size_t)rows * k, mem, "allocating indices"));
    CNVB_TRY(dl.init(ctx, ref_len, rows, mem, "allocating lengths"));
    if (mem != CNVB_MEM_DEVICE)
        CNVB_CUDA(ctx, cudaMemcpyAsync(di.p, ref_idx, (size_t)rows * k * 4, cudaMemcpyHostToDevice, ctx->stream),
                  "uploading indices");
    wis_prune_kernel<<<(rows + 7) / 8, 256, 0, ctx->stream>>>(dd_.p, di.p, rows, k, threshold, dl.p);
    CNVB_LAUNCHED(ctx, "wis_prune_kernel");
    CNVB_TRY(di.download());
    CNVB_TRY(dl.download());
    if (mem != CNVB_MEM_DEVICE) CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "wis prune");
    return CNVB_OK;
}

namespace {
int wis_calls_dev(cnvb_ctx *ctx, const float *dX, uint32_t S, const uint32_t *d_idx, const uint32_t *d_len, uint32_t k,
                  const cnvb_wis_opts *o, uint32_t row_begin, uint32_t rows, uint32_t *d_calls) {
    if (rows == 0) return 0;
    const size_t smem = (size_t)kCallsGroup * k * sizeof(uint32_t);
    CNVB_CUDA(ctx, cudaFuncSetAttribute(wis_calls_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
              "smem opt-in");
    KernelSpan span(ctx, "wis_calls_kernel");
    wis_calls_kernel<<<(rows + kCallsGroup - 1) / kCallsGroup, kCallsThreads, smem, ctx->stream>>>(
        dX, S, d_idx, d_len, k, o->min_ref_targets, o->filter_z_score, o->filter_rel, row_begin, rows, d_calls);
    CNVB_LAUNCHED(ctx, "wis_calls_kernel");
    return 0;
}
}  // namespace

extern "C" int cnvb_wis_calls(cnvb_ctx *ctx, const float *X, uint32_t T, uint32_t S, const uint32_t *ref_idx,
                              const uint32_t *ref_len, uint32_t k, const cnvb_wis_opts *opts, uint32_t row_begin,
                              uint32_t row_end, uint32_t *num_calls, int mem) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (!opts || !X || !ref_idx || !ref_len || !num_calls || row_end > T || row_begin > row_end || k > kMaxK)
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_wis_calls: bad argument");
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    const uint32_t rows = row_end - row_begin;
    DevIn<float> dX;
    DevIn<uint32_t> di, dl;
    DevOut<uint32_t> dc;
    CNVB_TRY(dX.init(ctx, X, (size_t)T * S, mem, "uploading the panel"));
    CNVB_TRY(di.init(ctx, ref_idx, (size_t)rows * k, mem, "uploading indices"));
    CNVB_TRY(dl.init(ctx, ref_len, rows, mem, "uploading lengths"));
    CNVB_TRY(dc.init(ctx, num_calls, rows, mem, "allocating call counts"));
    CNVB_TRY(wis_calls_dev(ctx, dX.p, S, di.p, dl.p, k, opts, row_begin, rows, dc.p));
    CNVB_TRY(dc.download());
    if (mem != CNVB_MEM_DEVICE) CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "wis calls");
    return CNVB_OK;
}

extern "C" int cnvb_wis_filters(cnvb_ctx *ctx, const uint32_t *ref_len, const uint32_t *num_calls, uint32_t rows,
                                const cnvb_wis_opts *opts, uint8_t *filt, int mem) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (!opts || (rows && (!ref_len || !num_calls || !filt))) return fail(ctx, CNVB_ERR_INVALID, "cnvb_wis_filters: null argument");
    if (rows == 0) return CNVB_OK;
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    DevIn<uint32_t> dl, dc;
    DevOut<uint8_t> df;
    CNVB_TRY(dl.init(ctx, ref_len, rows, mem, "uploading lengths"));
    CNVB_TRY(dc.init(ctx, num_calls, rows, mem, "uploading call counts"));
    CNVB_TRY(df.init(ctx, filt, rows, mem, "allocating filters"));
    wis_filters_kernel<<<(rows + 255) / 256, 256, 0, ctx->stream>>>(dl.p, dc.p, rows, opts->min_ref_targets,
                                                                    opts->max_samples_reliable, df.p);
    CNVB_LAUNCHED(ctx, "wis_filters_kernel");
    CNVB_TRY(df.download());
    if (mem != CNVB_MEM_DEVICE) CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "wis filters");
    return CNVB_OK;
}

extern "C" int cnvb_wis_build(cnvb_ctx *ctx, const float *X, const uint32_t *tid, uint32_t T, uint32_t S,
                              const cnvb_wis_opts *opts, float *ref_dist, uint32_t *ref_idx, uint32_t *ref_len,
                              uint32_t *num_calls, uint8_t *filt, float *threshold, int mem) {
    if (!ctx) return CNVB_ERR_INVALID;
    CNVB_TRY(check_wis_opts(ctx, opts, T));
    if (!X || !tid || !ref_dist || !ref_idx || !ref_len || !num_calls || !filt || !threshold || S == 0 || T == 0)
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_wis_build: bad argument");
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    const uint32_t k = T < opts->max_ref_targets ? T : opts->max_ref_targets;
    DevIn<float> dX;
    DevIn<uint32_t> dt;
    DevOut<float> od;
    DevOut<uint32_t> oi, ol, oc;
    DevOut<uint8_t> of;
    CNVB_TRY(dX.init(ctx, X, (size_t)T * S, mem, "uploading the panel"));
    CNVB_TRY(dt.init(ctx, tid, T, mem, "uploading contig ids"));
    CNVB_TRY(od.init(ctx, ref_dist, (size_t)T * k, mem, "allocating distances"));
    CNVB_TRY(oi.init(ctx, ref_idx, (size_t)T * k, mem, "allocating indices"));
    CNVB_TRY(ol.init(ctx, ref_len, T, mem, "allocating lengths"));
    CNVB_TRY(oc.init(ctx, num_calls, T, mem, "allocating call counts"));
    CNVB_TRY(of.init(ctx, filt, T, mem, "allocating filters"));
    CNVB_TRY(wis_distances_dev(ctx, dX.p, dt.p, T, S, opts, 0, T, od.p, oi.p));         // lib.rs:427
    CNVB_TRY(od.download());  // distances are reported un-pruned (ascending)
    CNVB_TRY(wis_threshold_dev(ctx, od.p, T, k, opts->filter_z_score, threshold));       // lib.rs:207-216
    wis_prune_kernel<<<(T + 7) / 8, 256, 0, ctx->stream>>>(od.p, oi.p, T, k, *threshold, ol.p);  // :221-228
    CNVB_LAUNCHED(ctx, "wis_prune_kernel");
    CNVB_TRY(wis_calls_dev(ctx, dX.p, S, oi.p, ol.p, k, opts, 0, T, oc.p));              // lib.rs:429-436
    wis_filters_kernel<<<(T + 255) / 256, 256, 0, ctx->stream>>>(ol.p, oc.p, T, opts->min_ref_targets,
                                                                 opts->max_samples_reliable, of.p);
    CNVB_LAUNCHED(ctx, "wis_filters_kernel");
    CNVB_TRY(oi.download());
    CNVB_TRY(ol.download());
    CNVB_TRY(oc.download());
    CNVB_TRY(of.download());
    if (mem != CNVB_MEM_DEVICE) CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "wis build");
    return CNVB_OK;
}

// ---- build-model-wis with the targets sharded by row block over the ranks of a communicator ------------
namespace {
__global__ void wis_nearest_kernel(const float *__restrict__ dist, uint32_t rows, uint32_t k, float *__restrict__ nearest) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) nearest[r] = dist[(size_t)r * k];
}
}  // namespace

extern "C" int cnvb_wis_build_sharded(cnvb_comm *comm, const float *X_rows, const uint32_t *tid_rows, uint32_t rows,
                                      uint32_t S, const cnvb_wis_opts *opts, float *ref_dist, uint32_t *ref_idx,
                                      uint32_t *ref_len, uint32_t *num_calls, uint8_t *filt, float *threshold,
                                      uint32_t *row_begin_out, uint32_t *T_out, int mem) {
    if (!comm) return CNVB_ERR_INVALID;
    cnvb_ctx *ctx = comm->ctx;
    if (!opts || !threshold || S == 0 || (rows && (!X_rows || !tid_rows || !ref_dist || !ref_idx || !ref_len || !num_calls || !filt)))
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_wis_build_sharded: bad argument");
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    const int world = comm->world, me = comm->rank;
    // row counts of all ranks (row blocks in rank order = target order)
    uint32_t *d_counts = nullptr;
    cudaError_t e = pool_alloc_t(ctx, (size_t)world, &d_counts);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "allocating the row counts");
    PoolGuard guard{ctx};
    guard.ptrs.push_back(d_counts);
    std::vector<uint32_t> counts((size_t)world, 0);
    {
        std::vector<uint64_t> four((size_t)world, 4);
        CNVB_CUDA(ctx, cudaMemcpyAsync(d_counts + me, &rows, 4, cudaMemcpyHostToDevice, ctx->stream), "row counts");
        CNVB_TRY(comm_all_gather_v(comm, d_counts + me, d_counts, four.data()));
        CNVB_CUDA(ctx, cudaMemcpyAsync(counts.data(), d_counts, (size_t)world * 4, cudaMemcpyDeviceToHost, ctx->stream), "row counts");
        CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "row counts");
    }
    uint64_t T64 = 0, lo64 = 0;
    for (int r = 0; r < world; ++r) {
        if (r == me) lo64 = T64;
        T64 += counts[(size_t)r];
    }
    if (T64 == 0 || T64 > 0xFFFFFFF0ull) return fail(ctx, CNVB_ERR_INVALID, "cnvb_wis_build_sharded: %llu targets", (unsigned long long)T64);
    const uint32_t T = (uint32_t)T64, lo = (uint32_t)lo64, hi = lo + rows;
    if (row_begin_out) *row_begin_out = lo;
    if (T_out) *T_out = T;
    CNVB_TRY(check_wis_opts(ctx, opts, T));
    const uint32_t k = T < opts->max_ref_targets ? T : opts->max_ref_targets;
    // the ONE big exchange: every rank needs the whole panel (80 MB at config 3, 12 GB at config 5)
    float *dX = nullptr;
    uint32_t *dt = nullptr;
    float *d_nearest = nullptr;
    if ((e = pool_alloc_t(ctx, (size_t)T * S, &dX)) != cudaSuccess) return fail_cuda(ctx, e, "allocating the panel");
    guard.ptrs.push_back(dX);
    if ((e = pool_alloc_t(ctx, (size_t)T, &dt)) != cudaSuccess) return fail_cuda(ctx, e, "allocating contig ids");
    guard.ptrs.push_back(dt);
    if ((e = pool_alloc_t(ctx, (size_t)T, &d_nearest)) != cudaSuccess) return fail_cuda(ctx, e, "allocating nearest distances");
    guard.ptrs.push_back(d_nearest);
    std::vector<uint64_t> bx((size_t)world), bt((size_t)world);
    for (int r = 0; r < world; ++r) {
        bx[(size_t)r] = (uint64_t)counts[(size_t)r] * S * 4;
        bt[(size_t)r] = (uint64_t)counts[(size_t)r] * 4;
    }
    const cudaMemcpyKind up = mem == CNVB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (rows) {
        CNVB_CUDA(ctx, cudaMemcpyAsync(dX + (size_t)lo * S, X_rows, (size_t)rows * S * 4, up, ctx->stream), "placing the panel rows");
        CNVB_CUDA(ctx, cudaMemcpyAsync(dt + lo, tid_rows, (size_t)rows * 4, up, ctx->stream), "placing contig ids");
    }
    CNVB_TRY(comm_all_gather_v(comm, dX + (size_t)lo * S, dX, bx.data()));
    CNVB_TRY(comm_all_gather_v(comm, dt + lo, dt, bt.data()));
    DevOut<float> od;
    DevOut<uint32_t> oi, ol, oc;
    DevOut<uint8_t> of;
    CNVB_TRY(od.init(ctx, ref_dist, (size_t)rows * k, mem, "allocating distances"));
    CNVB_TRY(oi.init(ctx, ref_idx, (size_t)rows * k, mem, "allocating indices"));
    CNVB_TRY(ol.init(ctx, ref_len, rows, mem, "allocating lengths"));
    CNVB_TRY(oc.init(ctx, num_calls, rows, mem, "allocating call counts"));
    CNVB_TRY(of.init(ctx, filt, rows, mem, "allocating filters"));
    CNVB_TRY(wis_distances_dev(ctx, dX, dt, T, S, opts, lo, hi, od.p, oi.p));                 // lib.rs:427, own rows
    CNVB_TRY(od.download());
    // prune threshold (lib.rs:207-216) runs over the nearest distance of ALL targets: one f32 per target
    if (rows) {
        wis_nearest_kernel<<<(rows + 255) / 256, 256, 0, ctx->stream>>>(od.p, rows, k, d_nearest + lo);
        CNVB_LAUNCHED(ctx, "wis_nearest_kernel");
    }
    CNVB_TRY(comm_all_gather_v(comm, d_nearest + lo, d_nearest, bt.data()));
    CNVB_TRY(wis_threshold_dev(ctx, d_nearest, T, 1, opts->filter_z_score, threshold));        // same bits on every rank
    if (rows) {
        wis_prune_kernel<<<(rows + 7) / 8, 256, 0, ctx->stream>>>(od.p, oi.p, rows, k, *threshold, ol.p);  // :221-228
        CNVB_LAUNCHED(ctx, "wis_prune_kernel");
        CNVB_TRY(wis_calls_dev(ctx, dX, S, oi.p, ol.p, k, opts, lo, rows, oc.p));               // lib.rs:429-436
        wis_filters_kernel<<<(rows + 255) / 256, 256, 0, ctx->stream>>>(ol.p, oc.p, rows, opts->min_ref_targets,
                                                                        opts->max_samples_reliable, of.p);
        CNVB_LAUNCHED(ctx, "wis_filters_kernel");
    }
    CNVB_TRY(oi.download());
    CNVB_TRY(ol.download());
    CNVB_TRY(oc.download());
    CNVB_TRY(of.download());
    if (mem != CNVB_MEM_DEVICE) CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "wis build (sharded)");
    return CNVB_OK;
}

namespace {
// launch shape of modcov_panel_kernel: staged with 64-sample chunks, staged with 32-sample chunks, or unstaged;
// 16-byte copies where the segments are 16-byte aligned.  CNVB_PANEL_STAGE=0 / 1 / 2 force the unstaged / the 4-byte
// cp.async / the bulk-copy form, CNVB_PANEL_SPL=1 the 32-sample chunks (A/B measurements and the tests of the fallbacks).
struct PanelLaunch {
    size_t smem;
    int mode, spl;
};
PanelLaunch panel_launch(uint32_t k, const float *X, uint32_t S) {
    const size_t ref = (size_t)kPanelGroup * k * sizeof(uint32_t);
    const size_t limit = 227 * 1024 - 3072;  // less the kernel's static arrays
    const char *e = getenv("CNVB_PANEL_STAGE");
    const int want = e ? atoi(e) : (int)kPanelAsync16;
    e = getenv("CNVB_PANEL_SPL");
    const int max_spl = e ? atoi(e) : 2;
    if (want != kPanelPlain)
        for (int spl = 2; spl >= 1; --spl) {
            const size_t stage = std::max((size_t)k * 32 * spl * kPanelWarps * sizeof(float), panel_out_bytes(spl));
            if (spl > max_spl || ref + stage > limit) continue;
            const bool aligned = S % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15u) == 0;
            return {ref + stage, (want == kPanelAsync || !aligned) ? (int)kPanelAsync : want == kPanelBulk ? (int)kPanelBulk : (int)kPanelAsync16, spl};
        }
    return {ref, kPanelPlain, 1};
}
template <bool CALLS, int SPL>
auto panel_kernel_spl(int mode) {
    return mode == kPanelAsync16 ? modcov_panel_kernel<CALLS, kPanelAsync16, SPL>
           : mode == kPanelBulk  ? modcov_panel_kernel<CALLS, kPanelBulk, SPL>
                                 : modcov_panel_kernel<CALLS, kPanelAsync, SPL>;
}
template <bool CALLS>
auto panel_kernel(const PanelLaunch &pl) {
    if (pl.mode == kPanelPlain) return modcov_panel_kernel<CALLS, kPanelPlain, 1>;
    return pl.spl == 2 ? panel_kernel_spl<CALLS, 2>(pl.mode) : panel_kernel_spl<CALLS, 1>(pl.mode);
}
}  // namespace

extern "C" int cnvb_modcov_panel(cnvb_ctx *ctx, const float *X, uint32_t T, uint32_t S, uint32_t row_begin,
                                 uint32_t row_end, const uint8_t *model_filt, const uint32_t *ref_idx,
                                 const uint32_t *ref_len, uint32_t k, float *ref_mean, float *ref_sd, float *cv_rel,
                                 float *cvz, float *cv2, uint8_t *status, int mem) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (row_end > T || row_begin > row_end) return fail(ctx, CNVB_ERR_INVALID, "cnvb_modcov_panel: bad row range");
    const uint32_t rows = row_end - row_begin;
    if (rows && S && (!X || !model_filt || !ref_idx || !ref_len))
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_modcov_panel: null argument");
    if (rows == 0 || S == 0) return CNVB_OK;
    if (k > kMaxK) return fail(ctx, CNVB_ERR_INVALID, "more than %u reference targets per target are not supported", kMaxK);
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    const size_t n = (size_t)rows * S;
    DevIn<float> dx;
    DevIn<uint8_t> dmf;
    DevIn<uint32_t> di, dl;
    DevOut<float> o1, o2, o3, o4, o5;
    DevOut<uint8_t> os;
    CNVB_TRY(dx.init(ctx, X, (size_t)T * S, mem, "uploading the CV panel"));
    CNVB_TRY(dmf.init(ctx, model_filt, rows, mem, "uploading model filters"));
    CNVB_TRY(di.init(ctx, ref_idx, (size_t)rows * k, mem, "uploading reference targets"));
    CNVB_TRY(dl.init(ctx, ref_len, rows, mem, "uploading reference target counts"));
    CNVB_TRY(o1.init(ctx, ref_mean, n, mem, "allocating outputs"));
    CNVB_TRY(o2.init(ctx, ref_sd, n, mem, "allocating outputs"));
    CNVB_TRY(o3.init(ctx, cv_rel, n, mem, "allocating outputs"));
    CNVB_TRY(o4.init(ctx, cvz, n, mem, "allocating outputs"));
    CNVB_TRY(o5.init(ctx, cv2, n, mem, "allocating outputs"));
    CNVB_TRY(os.init(ctx, status, n, mem, "allocating outputs"));
    {
        KernelSpan span(ctx, "modcov_panel_kernel");
        const PanelOut po{o1.p, o2.p, o3.p, o4.p, o5.p, os.p};
        const PanelLaunch pl = panel_launch(k, dx.p, S);
        auto kern = panel_kernel<false>(pl);
        if (pl.smem > 24 * 1024)
            CNVB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem), "smem opt-in");
        kern<<<(rows + kPanelGroup - 1) / kPanelGroup, kPanelThreads, pl.smem, ctx->stream>>>(
            dx.p, S, row_begin, rows, dmf.p, di.p, dl.p, k, po, PanelCalls{});
    }
    CNVB_LAUNCHED(ctx, "modcov_panel_kernel");
    CNVB_TRY(o1.download());
    CNVB_TRY(o2.download());
    CNVB_TRY(o3.download());
    CNVB_TRY(o4.download());
    CNVB_TRY(o5.download());
    CNVB_TRY(os.download());
    if (mem != CNVB_MEM_DEVICE) CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "modcov panel");
    return CNVB_OK;
}

extern "C" int cnvb_wis_calls_modcov_panel(cnvb_ctx *ctx, const float *X, uint32_t T, uint32_t S, const uint32_t *ref_idx,
                                           const uint32_t *ref_len, uint32_t k, const cnvb_wis_opts *opts,
                                           uint32_t row_begin, uint32_t row_end, uint32_t *num_calls, uint8_t *filt,
                                           float *ref_mean, float *ref_sd, float *cv_rel, float *cvz, float *cv2,
                                           uint8_t *status, int mem) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (!opts || row_end > T || row_begin > row_end || k > kMaxK)
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_wis_calls_modcov_panel: bad argument");
    const uint32_t rows = row_end - row_begin;
    if (rows && S && (!X || !ref_idx || !ref_len || !num_calls || !filt))
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_wis_calls_modcov_panel: null argument");
    if (rows == 0) return CNVB_OK;
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    const size_t n = (size_t)rows * S;
    DevIn<float> dx;
    DevIn<uint32_t> di, dl;
    DevOut<uint32_t> oc;
    DevOut<uint8_t> of, os;
    DevOut<float> o1, o2, o3, o4, o5;
    CNVB_TRY(dx.init(ctx, X, (size_t)T * S, mem, "uploading the CV panel"));
    CNVB_TRY(di.init(ctx, ref_idx, (size_t)rows * k, mem, "uploading reference targets"));
    CNVB_TRY(dl.init(ctx, ref_len, rows, mem, "uploading reference target counts"));
    CNVB_TRY(oc.init(ctx, num_calls, rows, mem, "allocating outputs"));
    CNVB_TRY(of.init(ctx, filt, rows, mem, "allocating outputs"));
    CNVB_TRY(o1.init(ctx, ref_mean, n, mem, "allocating outputs"));
    CNVB_TRY(o2.init(ctx, ref_sd, n, mem, "allocating outputs"));
    CNVB_TRY(o3.init(ctx, cv_rel, n, mem, "allocating outputs"));
    CNVB_TRY(o4.init(ctx, cvz, n, mem, "allocating outputs"));
    CNVB_TRY(o5.init(ctx, cv2, n, mem, "allocating outputs"));
    CNVB_TRY(os.init(ctx, status, n, mem, "allocating outputs"));
    const PanelOut po{o1.p, o2.p, o3.p, o4.p, o5.p, os.p};
    if (S == 0) {
        CNVB_CUDA(ctx, cudaMemsetAsync(oc.p, 0, rows * sizeof(uint32_t), ctx->stream), "clearing the call counts");
    } else {
        KernelSpan span(ctx, "modcov_panel_kernel");
        const PanelLaunch pl = panel_launch(k, dx.p, S);
        auto kern = panel_kernel<true>(pl);
        if (pl.smem > 24 * 1024)
            CNVB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem), "smem opt-in");
        const PanelCalls pc{opts->min_ref_targets, opts->filter_z_score, opts->filter_rel, oc.p};
        kern<<<(rows + kPanelGroup - 1) / kPanelGroup, kPanelThreads, pl.smem, ctx->stream>>>(
            dx.p, S, row_begin, rows, nullptr, di.p, dl.p, k, po, pc);
    }
    CNVB_LAUNCHED(ctx, "modcov_panel_kernel");
    wis_filters_kernel<<<(rows + 255) / 256, 256, 0, ctx->stream>>>(dl.p, oc.p, rows, opts->min_ref_targets,
                                                                    opts->max_samples_reliable, of.p);
    CNVB_LAUNCHED(ctx, "wis_filters_kernel");
    if (n) {
        modcov_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(of.p, rows, n, po);
        CNVB_LAUNCHED(ctx, "modcov_mask_kernel");
    }
    CNVB_TRY(oc.download());
    CNVB_TRY(of.download());
    CNVB_TRY(o1.download());
    CNVB_TRY(o2.download());
    CNVB_TRY(o3.download());
    CNVB_TRY(o4.download());
    CNVB_TRY(o5.download());
    CNVB_TRY(os.download());
    if (mem != CNVB_MEM_DEVICE) CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "calls + modcov panel");
    return CNVB_OK;
}

extern "C" int cnvb_modcov(cnvb_ctx *ctx, const float *cv, uint32_t T, const uint8_t *model_filt,
                           const uint32_t *ref_idx, const uint32_t *ref_len, uint32_t k, float *ref_mean,
                           float *ref_sd, float *cv_rel, float *cvz, float *cv2, uint8_t *status, int mem) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (T && (!cv || !model_filt || !ref_idx || !ref_len || !ref_mean || !ref_sd || !cv_rel || !cvz || !cv2 || !status))
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_modcov: null argument");
    if (T == 0) return CNVB_OK;
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    DevIn<float> dcv;
    DevIn<uint8_t> dmf;
    DevIn<uint32_t> di, dl;
    DevOut<float> o1, o2, o3, o4, o5;
    DevOut<uint8_t> os;
    CNVB_TRY(dcv.init(ctx, cv, T, mem, "uploading CV"));
    CNVB_TRY(dmf.init(ctx, model_filt, T, mem, "uploading model filters"));
    CNVB_TRY(di.init(ctx, ref_idx, (size_t)T * k, mem, "uploading reference targets"));
    CNVB_TRY(dl.init(ctx, ref_len, T, mem, "uploading reference target counts"));
    CNVB_TRY(o1.init(ctx, ref_mean, T, mem, "allocating outputs"));
    CNVB_TRY(o2.init(ctx, ref_sd, T, mem, "allocating outputs"));
    CNVB_TRY(o3.init(ctx, cv_rel, T, mem, "allocating outputs"));
    CNVB_TRY(o4.init(ctx, cvz, T, mem, "allocating outputs"));
    CNVB_TRY(o5.init(ctx, cv2, T, mem, "allocating outputs"));
    CNVB_TRY(os.init(ctx, status, T, mem, "allocating outputs"));
    {
        KernelSpan span(ctx, "modcov_kernel");
        modcov_kernel<<<(T + 127) / 128, 128, 0, ctx->stream>>>(dcv.p, T, dmf.p, di.p, dl.p, k, o1.p, o2.p, o3.p, o4.p,
                                                                o5.p, os.p);
    }
    CNVB_LAUNCHED(ctx, "modcov_kernel");
    CNVB_TRY(o1.download());
    CNVB_TRY(o2.download());
    CNVB_TRY(o3.download());
    CNVB_TRY(o4.download());
    CNVB_TRY(o5.download());
    CNVB_TRY(os.download());
    if (mem != CNVB_MEM_DEVICE) CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "modcov");
    return CNVB_OK;
}
