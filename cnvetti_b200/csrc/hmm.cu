// hmm.cu -- `cmd discover`: 5-state copy-number HMM segmentation (builder-defined spec).
//
// The reference ships no implementation (lib-discover/src/main.rs:1-3 is "Hello, world!";
// cnvetti/src/main.rs:135 bails), so the model is specified in DESIGN.md and checked against the
// repo's own CPU checker (the test-only C restatement of this spec):
//   states CN0..CN4; emission CV ~ Normal(mu_c, sd_c), mu_c = max(c, eps_cn)/2,
//   sd_c = sigma_s sqrt(mu_c) + sigma0; transitions stay 1-4p / move p (p <= 0.2); uniform start.
//
// Viterbi runs on INTEGER scores (units of 2^-16 nat; emissions from five f32 operations, floored
// at -64 nat): integer (max, +) is exactly associative, so the recurrence can be cut into chunks
// and evaluated in parallel along a lane and still give, bit for bit, the path of the sequential
// checker (argmax ties -> lowest state):
//   A  chunk_matrix   thread per 64-bin chunk: the 5x5 (max,+) transfer matrix of the chunk
//                     (five recurrences, one per entry state; O(K) per step thanks to the
//                     stay/move structure: max_r(d_r + t_rs) = max(d_s + stay, max_r d_r + move))
//   B  boundary       thread per lane: 5-vector times matrix along the chunks -> the exact scores
//                     (up to a common constant) entering every chunk
//   C  chunk_path     thread per chunk: the recurrence again from the true entry scores, packed
//                     15-bit back-pointers per bin, and the map "state at chunk end -> state at the
//                     end of the previous chunk"
//   D  backtrack      thread per lane: end state of every chunk, last chunk first
//   E  states         thread per chunk: trace back inside the chunk, one state byte per bin
// CV, back-pointers and states move between HBM and the per-thread rows through shared-memory
// tiles so that every global access is coalesced.  Per bin: 4 B CV read twice, 1 B back-pointer
// written and read, 1 B state written (+ ~2 B of per-chunk data).
// Posteriors (optional) stay a per-lane f64 forward-backward pass.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "common.cuh"

using namespace cnvb;

namespace {

constexpr int K = CNVB_HMM_STATES;
constexpr uint32_t kChunk = 64;          // bins per chunk
constexpr uint32_t kChunkThreads = 128;  // chunks per CTA
constexpr uint32_t kCvStride = kChunk + 1;   // words: conflict-free rows
constexpr uint32_t kBpStride = kChunk + 4;   // bytes (17 words)
constexpr uint32_t kStStride = kChunk + 4;   // bytes (17 words)
constexpr uint32_t kMatStride = 28;          // ints per stored matrix (25 used; 16-byte aligned rows)

struct HmmTables {  // per lane, posterior pass (f64)
    double log_start[K], log_trans[K][K], mu[K], inv_sd[K], log_norm[K];
};
struct VitTables {  // per lane, integer-score Viterbi
    float mu_f[K], inv_sd_f[K], log_norm_sc[K];  // inv_sd_f = (float)(inv_sd * sqrt(32768)), log_norm_sc = (float)(log_norm * 65536)
    int32_t sc_start, sc_stay, sc_move;
};

__device__ __forceinline__ void hmm_emit(float cvf, const HmmTables &t, double (&e)[K]) {
    if (!isfinite(cvf)) {
#pragma unroll
        for (int s = 0; s < K; ++s) e[s] = 0.0;
        return;
    }
    const double x = (double)cvf;
#pragma unroll
    for (int s = 0; s < K; ++s) {
        const double t1 = __dsub_rn(x, t.mu[s]);
        const double z = __dmul_rn(t1, t.inv_sd[s]);
        const double q = __dmul_rn(z, z);
        const double h = __dmul_rn(0.5, q);
        e[s] = __dsub_rn(t.log_norm[s], h);
    }
}

// the f32 part of the tables, in registers
struct EmitF {
    float mu[K], inv_sd[K], ln[K];
};
// integer emission scores: four f32 roundings, floor at -2^22, round to nearest even by the
// 1.5 * 2^23 trick (exact for |score| <= 2^22); 0 for a missing observation
__device__ __forceinline__ void emit_scores(float x, const EmitF &t, int (&e)[K]) {
    const bool ok = isfinite(x);
#pragma unroll
    for (int s = 0; s < K; ++s) {
        const float t1 = __fsub_rn(x, t.mu[s]);
        const float z = __fmul_rn(t1, t.inv_sd[s]);
        const float q = __fmul_rn(z, z);
        const float ll = fmaxf(__fsub_rn(t.ln[s], q), -4194304.0f);  // (x finite: ll is a number or -inf)
        e[s] = ok ? (int)(__float_as_uint(__fadd_rn(ll, 12582912.0f)) - 0x4B400000u) : 0;
    }
}

struct VitArgs {
    const float *cv;
    const unsigned long long *lane_off;   // [n_lanes + 1] bins
    const unsigned long long *chunk_off;  // [n_lanes + 1] chunks
    const VitTables *tables;              // [n_lanes]
    uint32_t n_lanes;
    unsigned long long n_chunks;
    // the lanes [lane_lo, lane_hi) = chunks [chunk_lo, chunk_hi) this launch works on (lane groups run
    // their five-kernel chains on different streams)
    uint32_t lane_lo, lane_hi;
    unsigned long long chunk_lo, chunk_hi;
    int *mat;             // [n_chunks][kMatStride]
    int *din;             // [n_chunks][K] scores entering the chunk (normalised)
    uint16_t *org;        // [n_chunks] chunk-end state -> state at the end of the previous chunk
    uint8_t *end_state;   // [n_chunks]
    uint8_t *fin_state;   // [n_lanes]
    uint8_t *bp;          // [bins] best predecessor overall (3 bits) | per state: moved there? (5 bits)
    uint8_t *state;       // [bins]
};

struct ChunkPos {
    uint32_t lane, k, len;
    unsigned long long b0;
};
// chunk index -> (lane, chunk within the lane, first bin, length); lanes hold >= 1 chunk each or none
__device__ __forceinline__ ChunkPos locate_chunk(const VitArgs &a, unsigned long long chunk) {
    uint32_t lo = 0, hi = a.n_lanes;  // last lane with chunk_off[lane] <= chunk
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a.chunk_off[mid] <= chunk) lo = mid; else hi = mid;
    }
    ChunkPos p;
    p.lane = lo;
    p.k = (uint32_t)(chunk - a.chunk_off[lo]);
    const unsigned long long l0 = a.lane_off[lo], l1 = a.lane_off[lo + 1];
    p.b0 = l0 + (unsigned long long)p.k * kChunk;
    const unsigned long long rest = l1 - p.b0;
    p.len = rest < kChunk ? (uint32_t)rest : kChunk;
    return p;
}

// warp-cooperative copy of the CV values of the warp's 32 chunks into the CTA tile
__device__ __forceinline__ void load_cv_tile(const VitArgs &a, const ChunkPos &p, bool valid, float *tile) {
    const uint32_t lane = lane_id(), wbase = (threadIdx.x & ~31u);
    for (uint32_t j = 0; j < 32; ++j) {
        const unsigned long long b0 = __shfl_sync(0xffffffffu, p.b0, j);
        const uint32_t len = __shfl_sync(0xffffffffu, valid ? p.len : 0u, j);
        float *row = tile + (wbase + j) * kCvStride;
        if (lane < len) row[lane] = a.cv[b0 + lane];
        if (lane + 32 < len) row[lane + 32] = a.cv[b0 + lane + 32];
    }
    __syncwarp();
}

__device__ __forceinline__ EmitF load_emit(const VitTables &t) {
    EmitF f;
#pragma unroll
    for (int s = 0; s < K; ++s) {
        f.mu[s] = t.mu_f[s];
        f.inv_sd[s] = t.inv_sd_f[s];
        f.ln[s] = t.log_norm_sc[s];
    }
    return f;
}

// ---- A: transfer matrix of every chunk ---------------------------------------------------------
__global__ void __launch_bounds__(kChunkThreads) hmm_chunk_matrix_kernel(VitArgs a) {
    extern __shared__ float s_cv[];
    const unsigned long long chunk = a.chunk_lo + (unsigned long long)blockIdx.x * kChunkThreads + threadIdx.x;
    const bool valid = chunk < a.chunk_hi;
    ChunkPos p{0, 0, 0, 0};
    if (valid) p = locate_chunk(a, chunk);
    load_cv_tile(a, p, valid, s_cv);
    if (!valid) return;
    const VitTables &tab = a.tables[p.lane];
    const EmitF em = load_emit(tab);
    const int stay = tab.sc_stay, move = tab.sc_move;
    const float *x = s_cv + threadIdx.x * kCvStride;
    int d[K][K], e[K];
    emit_scores(x[0], em, e);
    if (p.k == 0) {
        // the lane starts here: row 0 carries the start distribution, the other rows are copies far
        // below it (a constant offset survives every (max,+) step), so B can use one rule for all
#pragma unroll
        for (int c = 0; c < K; ++c) d[0][c] = tab.sc_start + e[c];
#pragma unroll
        for (int r = 1; r < K; ++r)
#pragma unroll
            for (int c = 0; c < K; ++c) d[r][c] = d[0][c] - (1 << 29);
    } else {
#pragma unroll
        for (int r = 0; r < K; ++r)
#pragma unroll
            for (int c = 0; c < K; ++c) d[r][c] = (r == c ? stay : move) + e[c];
    }
    // Once the best paths from all five entry states have merged -- with data that speaks for one state a move
    // is paid once and staying wrong is paid per bin, so after two or three bins -- the matrix has rank one in
    // (max,+): d[r][c] = a[r] + d[0][c].  A product of a rank-one matrix with anything stays rank one, so from
    // there on only row 0 is carried (13 instead of 65 instructions per bin beside the 35 of the emissions) and
    // the rows are put back at the end: the same integers, (max,+) being exactly associative.  The test costs 35
    // instructions and runs after bins 3, 7, 15 and 31; a warp switches when all its chunks have merged (either
    // form gives the same matrix, so the vote is about speed only).
    bool merged = false;
    int arow[K];
#pragma unroll
    for (int r = 0; r < K; ++r) arow[r] = 0;
    for (uint32_t i = 1; i < p.len; ++i) {
        emit_scores(x[i], em, e);
        if (merged) {
            const int m = __vimax3_s32(__vimax3_s32(d[0][0], d[0][1], d[0][2]), d[0][3], d[0][4]) + move;
#pragma unroll
            for (int c = 0; c < K; ++c) d[0][c] = __viaddmax_s32(d[0][c], stay, m) + e[c];
        } else {
#pragma unroll
            for (int r = 0; r < K; ++r) {  // DPX: three-way max and fused add-max
                const int m = __vimax3_s32(__vimax3_s32(d[r][0], d[r][1], d[r][2]), d[r][3], d[r][4]) + move;
#pragma unroll
                for (int c = 0; c < K; ++c) d[r][c] = __viaddmax_s32(d[r][c], stay, m) + e[c];
            }
        }
        if ((i & 15u) == 0u) {  // keep the scores small: a common constant changes no decision
            int g = d[0][0];
#pragma unroll
            for (int c = 1; c < K; ++c) g = max(g, d[0][c]);
#pragma unroll
            for (int r = 0; r < K; ++r)
#pragma unroll
                for (int c = 0; c < K; ++c) d[r][c] -= g;  // (merged: rows 1.. are dead, the compiler drops them)
        }
        if (!merged && (i == 3u || i == 7u || i == 15u || i == 31u)) {
            bool one = true;
#pragma unroll
            for (int r = 1; r < K; ++r)
#pragma unroll
                for (int c = 1; c < K; ++c) one = one && (d[r][c] - d[0][c]) == (d[r][0] - d[0][0]);
            if (__all_sync(__activemask(), one)) {
                merged = true;
#pragma unroll
                for (int r = 1; r < K; ++r) arow[r] = d[r][0] - d[0][0];
            }
        }
    }
    if (merged) {
#pragma unroll
        for (int r = 1; r < K; ++r)
#pragma unroll
            for (int c = 0; c < K; ++c) d[r][c] = arow[r] + d[0][c];
    }
    {
        int g = d[0][0];
#pragma unroll
        for (int c = 1; c < K; ++c) g = max(g, d[0][c]);
        int *out = a.mat + chunk * kMatStride;
        int flat[kMatStride];
#pragma unroll
        for (int r = 0; r < K; ++r)
#pragma unroll
            for (int c = 0; c < K; ++c) flat[r * K + c] = d[r][c] - g;
        flat[25] = flat[26] = flat[27] = 0;
#pragma unroll
        for (int q = 0; q < (int)kMatStride / 4; ++q)
            reinterpret_cast<int4 *>(out)[q] = make_int4(flat[4 * q], flat[4 * q + 1], flat[4 * q + 2], flat[4 * q + 3]);
    }
}

// ---- B: scores entering every chunk ------------------------------------------------------------
// One warp per lane.  The recurrence along the chunks is serial, so the warp fetches the matrices 32
// chunks at a time with coalesced loads (the next batch is in flight while the current one is
// applied); every lane then walks the batch redundantly out of shared memory and lane j keeps the
// vector entering chunk j, so the results leave with one coalesced store per batch.
constexpr uint32_t kLaneWarps = 4;

__global__ void __launch_bounds__(kLaneWarps * 32) hmm_boundary_kernel(VitArgs a) {
    __shared__ int4 s_m[kLaneWarps][32 * kMatStride / 4];
    const uint32_t lane_idx = a.lane_lo + blockIdx.x * kLaneWarps + (threadIdx.x >> 5), lane = lane_id();
    if (lane_idx >= a.lane_hi) return;
    const unsigned long long c0 = a.chunk_off[lane_idx], c1 = a.chunk_off[lane_idx + 1];
    if (c1 == c0) return;
    int4 *sm = s_m[threadIdx.x >> 5];
    constexpr uint32_t kQ = kMatStride / 4;  // int4 per matrix = int4 per lane and batch
    int4 pre[kQ];
    auto fetch = [&](unsigned long long base) {
        const uint32_t n = (uint32_t)min((unsigned long long)32, c1 - base);
        const int4 *src = reinterpret_cast<const int4 *>(a.mat + base * kMatStride);
#pragma unroll
        for (uint32_t q = 0; q < kQ; ++q) {
            const uint32_t idx = q * 32 + lane;
            pre[q] = idx < n * kQ ? src[idx] : make_int4(0, 0, 0, 0);
        }
    };
    fetch(c0);
    int dv[K] = {0, 0, 0, 0, 0};
    for (unsigned long long base = c0; base < c1; base += 32) {
        const uint32_t n = (uint32_t)min((unsigned long long)32, c1 - base);
        __syncwarp();
#pragma unroll
        for (uint32_t q = 0; q < kQ; ++q) sm[q * 32 + lane] = pre[q];
        __syncwarp();
        if (base + 32 < c1) fetch(base + 32);
        int mine[K] = {0, 0, 0, 0, 0};
        for (uint32_t j = 0; j < n; ++j) {
            if (lane == j) {
#pragma unroll
                for (int c = 0; c < K; ++c) mine[c] = dv[c];
            }
            int m[kMatStride];
#pragma unroll
            for (uint32_t q = 0; q < kQ; ++q) {
                const int4 v = sm[j * kQ + q];
                m[4 * q] = v.x;
                m[4 * q + 1] = v.y;
                m[4 * q + 2] = v.z;
                m[4 * q + 3] = v.w;
            }
            int nd[K];
#pragma unroll
            for (int c = 0; c < K; ++c) {
                int best = dv[0] + m[c];
#pragma unroll
                for (int r = 1; r < K; ++r) best = max(best, dv[r] + m[r * K + c]);
                nd[c] = best;
            }
            int g = nd[0];
#pragma unroll
            for (int c = 1; c < K; ++c) g = max(g, nd[c]);
#pragma unroll
            for (int c = 0; c < K; ++c) dv[c] = nd[c] - g;
        }
        if (lane < n) {
#pragma unroll
            for (int c = 0; c < K; ++c) a.din[(base + lane) * K + c] = mine[c];
        }
    }
    if (lane == 0) {
        int arg = 0, top = dv[0];
#pragma unroll
        for (int c = 1; c < K; ++c) {
            if (dv[c] > top) {  // strict: ties keep the lowest state
                top = dv[c];
                arg = c;
            }
        }
        a.fin_state[lane_idx] = (uint8_t)arg;
    }
}

// ---- C: back-pointers of every chunk from its true entry scores ------------------------------
// The best predecessor of state c is either c itself (stay) or the best state overall (r*), so a
// bin's back-pointers fit one byte: r* in bits 0..2, "moved" flags of the five states in bits 3..7.
__global__ void __launch_bounds__(kChunkThreads) hmm_chunk_path_kernel(VitArgs a) {
    extern __shared__ float s_cv[];
    uint8_t *s_bp = reinterpret_cast<uint8_t *>(s_cv + kChunkThreads * kCvStride);
    const unsigned long long chunk = a.chunk_lo + (unsigned long long)blockIdx.x * kChunkThreads + threadIdx.x;
    const bool valid = chunk < a.chunk_hi;
    ChunkPos p{0, 0, 0, 0};
    if (valid) p = locate_chunk(a, chunk);
    load_cv_tile(a, p, valid, s_cv);
    if (valid) {
        const VitTables &tab = a.tables[p.lane];
        const EmitF em = load_emit(tab);
        const int stay = tab.sc_stay, move = tab.sc_move;
        const float *x = s_cv + threadIdx.x * kCvStride;
        uint8_t *bp = s_bp + threadIdx.x * kBpStride;
        int d[K], e[K];
        uint32_t o[K];  // state at the end of the previous chunk that a path ending in c came from
#pragma unroll
        for (int c = 0; c < K; ++c) o[c] = c;
        uint32_t first = 0;
        if (p.k == 0) {
            emit_scores(x[0], em, e);
#pragma unroll
            for (int c = 0; c < K; ++c) d[c] = tab.sc_start + e[c];
            bp[0] = 0;
            first = 1;
        } else {
#pragma unroll
            for (int c = 0; c < K; ++c) d[c] = a.din[chunk * K + c];
        }
        for (uint32_t i = first; i < p.len; ++i) {
            emit_scores(x[i], em, e);
            int m = d[0];
            uint32_t rstar = 0;
#pragma unroll
            for (int c = 1; c < K; ++c) {  // lowest state among the best predecessors: ties keep the earlier one
                bool keep;
                m = __vibmax_s32(m, d[c], &keep);
                rstar = keep ? rstar : (uint32_t)c;
            }
            const int mv = m + move;
            // origin of the best predecessor; in the first bin of a later chunk the predecessors ARE the origins
            uint32_t orig = o[0];
#pragma unroll
            for (int c = 1; c < K; ++c) orig = rstar == (uint32_t)c ? o[c] : orig;
            uint32_t packed = rstar;
#pragma unroll
            for (int c = 0; c < K; ++c) {
                const int sv = d[c] + stay;
                // staying wins unless moving is strictly better, or equally good from a lower state
                const bool moved = sv < mv + ((uint32_t)c > rstar ? 1 : 0);
                d[c] = max(sv, mv) + e[c];
                packed |= moved ? (8u << c) : 0u;
                o[c] = moved ? orig : o[c];
            }
            bp[i] = (uint8_t)packed;
            if ((i & 15u) == 15u) {
                const int g = __vimax3_s32(__vimax3_s32(d[0], d[1], d[2]), d[3], d[4]);
#pragma unroll
                for (int c = 0; c < K; ++c) d[c] -= g;
            }
        }
        a.org[chunk] = (uint16_t)(o[0] | (o[1] << 3) | (o[2] << 6) | (o[3] << 9) | (o[4] << 12));
    }
    __syncwarp();
    {   // coalesced copy of the warp's back-pointer rows to HBM
        const uint32_t lane = lane_id(), wbase = (threadIdx.x & ~31u);
        for (uint32_t j = 0; j < 32; ++j) {
            const unsigned long long b0 = __shfl_sync(0xffffffffu, p.b0, j);
            const uint32_t len = __shfl_sync(0xffffffffu, valid ? p.len : 0u, j);
            const uint8_t *row = s_bp + (wbase + j) * kBpStride;
            if (lane < len) a.bp[b0 + lane] = row[lane];
            if (lane + 32 < len) a.bp[b0 + lane + 32] = row[lane + 32];
        }
    }
}

// ---- D: end state of every chunk ---------------------------------------------------------------
// One warp per lane, 32 chunks per coalesced load, the serial walk through shuffles.
__global__ void __launch_bounds__(kLaneWarps * 32) hmm_backtrack_kernel(VitArgs a) {
    const uint32_t lane_idx = a.lane_lo + blockIdx.x * kLaneWarps + (threadIdx.x >> 5), lane = lane_id();
    if (lane_idx >= a.lane_hi) return;
    const unsigned long long c0 = a.chunk_off[lane_idx], c1 = a.chunk_off[lane_idx + 1];
    if (c1 == c0) return;
    uint32_t s = a.fin_state[lane_idx];
    unsigned long long hi = c1;
    unsigned long long lo = hi - c0 > 32 ? hi - 32 : c0;
    uint32_t o = lane < (uint32_t)(hi - lo) ? a.org[lo + lane] : 0u;
    while (true) {
        const uint32_t n = (uint32_t)(hi - lo);
        const unsigned long long nhi = lo, nlo = nhi - c0 > 32 ? nhi - 32 : c0;
        uint32_t o_next = 0u;
        if (nhi > c0 && lane < (uint32_t)(nhi - nlo)) o_next = a.org[nlo + lane];  // in flight during the walk
        uint32_t mine = 0;
        for (uint32_t j = n; j-- > 0;) {
            if (lane == j) mine = s;
            s = (__shfl_sync(0xffffffffu, o, j) >> (3 * s)) & 7u;
        }
        if (lane < n) a.end_state[lo + lane] = (uint8_t)mine;
        if (nhi == c0) break;
        hi = nhi;
        lo = nlo;
        o = o_next;
    }
}

// ---- E: states inside every chunk ----------------------------------------------------------------
__global__ void __launch_bounds__(kChunkThreads) hmm_states_kernel(VitArgs a) {
    extern __shared__ float s_dyn[];
    uint8_t *s_bp = reinterpret_cast<uint8_t *>(s_dyn);
    uint8_t *s_st = s_bp + kChunkThreads * kBpStride;
    const unsigned long long chunk = a.chunk_lo + (unsigned long long)blockIdx.x * kChunkThreads + threadIdx.x;
    const bool valid = chunk < a.chunk_hi;
    ChunkPos p{0, 0, 0, 0};
    if (valid) p = locate_chunk(a, chunk);
    const uint32_t lane = lane_id(), wbase = (threadIdx.x & ~31u);
    for (uint32_t j = 0; j < 32; ++j) {
        const unsigned long long b0 = __shfl_sync(0xffffffffu, p.b0, j);
        const uint32_t len = __shfl_sync(0xffffffffu, valid ? p.len : 0u, j);
        uint8_t *row = s_bp + (wbase + j) * kBpStride;
        if (lane < len) row[lane] = a.bp[b0 + lane];
        if (lane + 32 < len) row[lane + 32] = a.bp[b0 + lane + 32];
    }
    __syncwarp();
    if (valid) {
        const uint8_t *bp = s_bp + threadIdx.x * kBpStride;
        uint8_t *st = s_st + threadIdx.x * kStStride;
        uint32_t s = a.end_state[chunk];
        for (uint32_t i = p.len; i-- > 0;) {
            st[i] = (uint8_t)s;
            const uint32_t b = bp[i];
            if (i > 0) s = ((b >> (3 + s)) & 1u) ? (b & 7u) : s;
        }
    }
    __syncwarp();
    for (uint32_t j = 0; j < 32; ++j) {
        const unsigned long long b0 = __shfl_sync(0xffffffffu, p.b0, j);
        const uint32_t len = __shfl_sync(0xffffffffu, valid ? p.len : 0u, j);
        const uint8_t *row = s_st + (wbase + j) * kStStride;
        if (lane < len) a.state[b0 + lane] = row[lane];
        if (lane + 32 < len) a.state[b0 + lane + 32] = row[lane + 32];
    }
}

struct HmmArgs {  // posterior pass
    const float *cv;
    const unsigned long long *lane_off;  // device copy
    const uint32_t *order;               // lanes, longest first
    const HmmTables *tables;             // [n_lanes]
    uint32_t n_lanes;
    double *alpha;                       // [total bins][K]
    double *post;
};

__device__ __forceinline__ double lse5(const double (&v)[K]) {
    double m = v[0];
#pragma unroll
    for (int s = 1; s < K; ++s)
        if (v[s] > m) m = v[s];
    if (!(m > -INFINITY)) return m;
    double acc = 0.0;
#pragma unroll
    for (int s = 0; s < K; ++s) acc += exp(v[s] - m);
    return m + log(acc);
}

__global__ void __launch_bounds__(32) hmm_posterior_kernel(HmmArgs a) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= a.n_lanes) return;
    const uint32_t lane = a.order[slot];
    const unsigned long long b0 = a.lane_off[lane], b1 = a.lane_off[lane + 1];
    if (b1 <= b0) return;
    const HmmTables t = a.tables[lane];
    double e[K], v[K], al[K];
    hmm_emit(a.cv[b0], t, e);
#pragma unroll
    for (int s = 0; s < K; ++s) {
        al[s] = t.log_start[s] + e[s];
        a.alpha[b0 * K + s] = al[s];
    }
    for (unsigned long long b = b0 + 1; b < b1; ++b) {
        hmm_emit(a.cv[b], t, e);
        double na[K];
#pragma unroll
        for (int s = 0; s < K; ++s) {
#pragma unroll
            for (int r = 0; r < K; ++r) v[r] = al[r] + t.log_trans[r][s];
            na[s] = e[s] + lse5(v);
        }
#pragma unroll
        for (int s = 0; s < K; ++s) {
            al[s] = na[s];
            a.alpha[b * K + s] = na[s];
        }
    }
    double beta[K];
#pragma unroll
    for (int s = 0; s < K; ++s) beta[s] = 0.0;
    for (unsigned long long b = b1; b-- > b0;) {
#pragma unroll
        for (int s = 0; s < K; ++s) v[s] = a.alpha[b * K + s] + beta[s];
        const double norm = lse5(v);
#pragma unroll
        for (int s = 0; s < K; ++s) a.post[b * K + s] = exp(v[s] - norm);
        if (b > b0) {
            hmm_emit(a.cv[b], t, e);
            double nb[K];
#pragma unroll
            for (int r = 0; r < K; ++r) {
#pragma unroll
                for (int s = 0; s < K; ++s) v[s] = t.log_trans[r][s] + e[s] + beta[s];
                nb[r] = lse5(v);
            }
#pragma unroll
            for (int r = 0; r < K; ++r) beta[r] = nb[r];
        }
    }
}

// same arithmetic, in the same order, as the CPU checker's table builder (host libm on both sides)
void make_tables(double sigma_s, const cnvb_hmm_opts &o, HmmTables &t, VitTables &v) {
    for (int c = 0; c < K; ++c) {
        const double cn = (double)c > o.eps_cn ? (double)c : o.eps_cn;
        t.mu[c] = cn / 2.0;
        const double sd = sigma_s * std::sqrt(t.mu[c]) + o.sigma0;
        t.inv_sd[c] = 1.0 / sd;
        t.log_norm[c] = -std::log(sd) - 0.5 * std::log(2.0 * M_PI);
        t.log_start[c] = std::log(1.0 / (double)K);
        for (int d = 0; d < K; ++d)
            t.log_trans[c][d] = (c == d) ? std::log(1.0 - (double)(K - 1) * o.p_move) : std::log(o.p_move);
        v.mu_f[c] = (float)t.mu[c];
        v.inv_sd_f[c] = (float)(t.inv_sd[c] * 181.01933598375618);  // sqrt(32768): folds the 1/2 and the score unit
        v.log_norm_sc[c] = (float)(t.log_norm[c] * 65536.0);
    }
    v.sc_start = (int32_t)std::llrint(t.log_start[0] * 65536.0);
    v.sc_stay = (int32_t)std::llrint(t.log_trans[0][0] * 65536.0);
    v.sc_move = (int32_t)std::llrint(t.log_trans[0][1] * 65536.0);
}

}  // namespace

extern "C" int cnvb_hmm_segment(cnvb_ctx *ctx, const float *cv, const uint64_t *lane_off, uint32_t n_lanes,
                                const double *sigma_s, const cnvb_hmm_opts *opts, uint8_t *state, double *posterior,
                                int mem) {
    if (!ctx) return CNVB_ERR_INVALID;
    if (!opts || !lane_off || !sigma_s || (n_lanes && (!cv || !state)))
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_hmm_segment: null argument");
    if (n_lanes == 0) return CNVB_OK;
    // p_move <= 0.2: staying must be at least as likely as any single move (the O(K) recurrence relies on it)
    if (!(opts->p_move > 0.0) || !(opts->p_move <= 0.2) || !(opts->sigma0 >= 0.0) || !(opts->eps_cn > 0.0))
        return fail(ctx, CNVB_ERR_INVALID, "cnvb_hmm_segment: invalid model parameters (need 0 < p_move <= 0.2, "
                                           "sigma0 >= 0, eps_cn > 0)");
    for (uint32_t l = 0; l < n_lanes; ++l) {
        if (lane_off[l + 1] < lane_off[l]) return fail(ctx, CNVB_ERR_INVALID, "lane offsets must be non-decreasing");
        if (!(sigma_s[l] >= 0.0) || !std::isfinite(sigma_s[l]) || !(sigma_s[l] + opts->sigma0 > 0.0))
            return fail(ctx, CNVB_ERR_INVALID, "lane %u: invalid noise level", l);
    }
    CNVB_CUDA(ctx, cudaSetDevice(ctx->device), "selecting device");
    const uint64_t base = lane_off[0], total = lane_off[n_lanes] - base;
    if (total == 0) return CNVB_OK;
    std::vector<HmmTables> tabs(n_lanes);
    std::vector<VitTables> vtabs(n_lanes);
    for (uint32_t l = 0; l < n_lanes; ++l) make_tables(sigma_s[l], *opts, tabs[l], vtabs[l]);
    std::vector<unsigned long long> offs(n_lanes + 1), coffs(n_lanes + 1);
    coffs[0] = 0;
    for (uint32_t l = 0; l <= n_lanes; ++l) offs[l] = lane_off[l] - base;
    for (uint32_t l = 0; l < n_lanes; ++l) coffs[l + 1] = coffs[l] + (offs[l + 1] - offs[l] + kChunk - 1) / kChunk;
    const unsigned long long n_chunks = coffs[n_lanes];

    struct Guard {
        cnvb_ctx *c;
        std::vector<void *> p;
        ~Guard() {
            for (void *x : p) pool_release(c, x);
        }
    } g{ctx, {}};
    auto alloc = [&](size_t bytes, void **out) -> cudaError_t {
        cudaError_t e = pool_alloc(ctx, bytes, out);
        if (e == cudaSuccess) g.p.push_back(*out);
        return e;
    };
    void *d_off = nullptr, *d_coff = nullptr, *d_tabs = nullptr, *d_bp = nullptr, *d_cv = nullptr, *d_state = nullptr,
         *d_mat = nullptr, *d_din = nullptr, *d_org = nullptr, *d_end = nullptr, *d_fin = nullptr;
    cudaError_t e;
    if ((e = alloc((n_lanes + 1) * 8, &d_off)) != cudaSuccess || (e = alloc((n_lanes + 1) * 8, &d_coff)) != cudaSuccess ||
        (e = alloc(n_lanes * sizeof(VitTables), &d_tabs)) != cudaSuccess || (e = alloc(total, &d_bp)) != cudaSuccess ||
        (e = alloc(n_chunks * kMatStride * 4, &d_mat)) != cudaSuccess || (e = alloc(n_chunks * K * 4, &d_din)) != cudaSuccess ||
        (e = alloc(n_chunks * 2, &d_org)) != cudaSuccess || (e = alloc(n_chunks, &d_end)) != cudaSuccess ||
        (e = alloc(n_lanes, &d_fin)) != cudaSuccess)
        return fail_cuda(ctx, e, "allocating HMM work space");
    const bool host = mem != CNVB_MEM_DEVICE;
    const float *cvp = cv + base;
    uint8_t *stp = state + base;
    double *pop = posterior ? posterior + base * K : nullptr;
    if (host) {
        if ((e = alloc(total * 4, &d_cv)) != cudaSuccess || (e = alloc(total, &d_state)) != cudaSuccess)
            return fail_cuda(ctx, e, "allocating HMM work space");
        CNVB_CUDA(ctx, cudaMemcpyAsync(d_cv, cvp, total * 4, cudaMemcpyHostToDevice, ctx->stream), "uploading CV");
    }
    CNVB_CUDA(ctx, cudaMemcpyAsync(d_off, offs.data(), (n_lanes + 1) * 8, cudaMemcpyHostToDevice, ctx->stream), "lanes");
    CNVB_CUDA(ctx, cudaMemcpyAsync(d_coff, coffs.data(), (n_lanes + 1) * 8, cudaMemcpyHostToDevice, ctx->stream), "lanes");
    CNVB_CUDA(ctx, cudaMemcpyAsync(d_tabs, vtabs.data(), n_lanes * sizeof(VitTables), cudaMemcpyHostToDevice, ctx->stream),
              "tables");
    VitArgs va{};
    va.cv = host ? static_cast<const float *>(d_cv) : cvp;
    va.lane_off = static_cast<const unsigned long long *>(d_off);
    va.chunk_off = static_cast<const unsigned long long *>(d_coff);
    va.tables = static_cast<const VitTables *>(d_tabs);
    va.n_lanes = n_lanes;
    va.n_chunks = n_chunks;
    va.mat = static_cast<int *>(d_mat);
    va.din = static_cast<int *>(d_din);
    va.org = static_cast<uint16_t *>(d_org);
    va.end_state = static_cast<uint8_t *>(d_end);
    va.fin_state = static_cast<uint8_t *>(d_fin);
    va.bp = static_cast<uint8_t *>(d_bp);
    va.state = host ? static_cast<uint8_t *>(d_state) : stp;
    const size_t smem_a = (size_t)kChunkThreads * kCvStride * 4;
    const size_t smem_c = smem_a + (size_t)kChunkThreads * kBpStride;
    const size_t smem_e = (size_t)kChunkThreads * kBpStride + (size_t)kChunkThreads * kStStride;
    CNVB_CUDA(ctx, cudaFuncSetAttribute(hmm_chunk_path_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c),
              "smem opt-in");
    // Lane groups on side streams: the per-lane scans (B, D) are latency-bound with one warp per lane
    // and the chunk kernels (A, C) are issue-bound, so the chains of different groups overlap well.
    // Groups are contiguous lane ranges with about the same number of chunks.  (Timing mode: one group.)
    constexpr int kGroups = 3;
    int n_groups = (ctx->timing || n_lanes < 2 * kGroups) ? 1 : kGroups;
    if (n_groups > 1) {
        for (int q = 0; q < kGroups && n_groups > 1; ++q) {
            if (!ctx->side[q] && cudaStreamCreateWithFlags(&ctx->side[q], cudaStreamNonBlocking) != cudaSuccess) n_groups = 1;
            if (n_groups > 1 && !ctx->side_join[q] &&
                cudaEventCreateWithFlags(&ctx->side_join[q], cudaEventDisableTiming) != cudaSuccess)
                n_groups = 1;
        }
        if (n_groups > 1 && !ctx->side_fork && cudaEventCreateWithFlags(&ctx->side_fork, cudaEventDisableTiming) != cudaSuccess)
            n_groups = 1;
    }
    cudaStream_t main_stream = ctx->stream;
    if (n_groups > 1) {
        CNVB_CUDA(ctx, cudaEventRecord(ctx->side_fork, main_stream), "forking streams");
        for (int q = 0; q < n_groups; ++q) CNVB_CUDA(ctx, cudaStreamWaitEvent(ctx->side[q], ctx->side_fork, 0), "forking streams");
    }
    uint32_t lane_cursor = 0;
    for (int q = 0; q < n_groups; ++q) {
        // lanes of group q: up to the lane where the chunk count passes (q + 1) / n_groups of the total
        uint32_t l_hi = n_lanes;
        if (q + 1 < n_groups) {
            const unsigned long long want = n_chunks * (unsigned long long)(q + 1) / (unsigned long long)n_groups;
            l_hi = lane_cursor;
            while (l_hi < n_lanes && coffs[l_hi + 1] <= want) ++l_hi;
        }
        va.lane_lo = lane_cursor;
        va.lane_hi = l_hi;
        va.chunk_lo = coffs[lane_cursor];
        va.chunk_hi = coffs[l_hi];
        lane_cursor = l_hi;
        const unsigned long long gc = va.chunk_hi - va.chunk_lo;
        if (gc == 0) continue;
        cudaStream_t st = n_groups > 1 ? ctx->side[q] : main_stream;
        const unsigned cblocks = (unsigned)((gc + kChunkThreads - 1) / kChunkThreads);
        const unsigned lblocks = (va.lane_hi - va.lane_lo + kLaneWarps - 1) / kLaneWarps;
        ctx->stream = st;  // (KernelSpan records on the launching stream)
        {
            KernelSpan span(ctx, "hmm_chunk_matrix_kernel");
            hmm_chunk_matrix_kernel<<<cblocks, kChunkThreads, smem_a, st>>>(va);
        }
        {
            KernelSpan span(ctx, "hmm_boundary_kernel");
            hmm_boundary_kernel<<<lblocks, kLaneWarps * 32, 0, st>>>(va);
        }
        {
            KernelSpan span(ctx, "hmm_chunk_path_kernel");
            hmm_chunk_path_kernel<<<cblocks, kChunkThreads, smem_c, st>>>(va);
        }
        {
            KernelSpan span(ctx, "hmm_backtrack_kernel");
            hmm_backtrack_kernel<<<lblocks, kLaneWarps * 32, 0, st>>>(va);
        }
        {
            KernelSpan span(ctx, "hmm_states_kernel");
            hmm_states_kernel<<<cblocks, kChunkThreads, smem_e, st>>>(va);
        }
        ctx->stream = main_stream;
        ctx->launches += 5;
        if (cudaGetLastError() != cudaSuccess) return fail(ctx, CNVB_ERR_CUDA, "launch of the HMM kernels");
    }
    if (n_groups > 1) {
        for (int q = 0; q < n_groups; ++q) {
            CNVB_CUDA(ctx, cudaEventRecord(ctx->side_join[q], ctx->side[q]), "joining streams");
            CNVB_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->side_join[q], 0), "joining streams");
        }
    }
    const unsigned pblocks = (n_lanes + 31) / 32;
    if (posterior) {
        std::vector<uint32_t> order(n_lanes);
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
            return lane_off[x + 1] - lane_off[x] > lane_off[y + 1] - lane_off[y];
        });
        void *d_order = nullptr, *d_alpha = nullptr, *d_post = nullptr, *d_ptabs = nullptr;
        if ((e = alloc(n_lanes * 4, &d_order)) != cudaSuccess || (e = alloc(total * K * 8, &d_alpha)) != cudaSuccess ||
            (e = alloc(n_lanes * sizeof(HmmTables), &d_ptabs)) != cudaSuccess)
            return fail_cuda(ctx, e, "allocating forward variables");
        CNVB_CUDA(ctx, cudaMemcpyAsync(d_ptabs, tabs.data(), n_lanes * sizeof(HmmTables), cudaMemcpyHostToDevice, ctx->stream),
                  "tables");
        if (host && (e = alloc(total * K * 8, &d_post)) != cudaSuccess) return fail_cuda(ctx, e, "allocating posteriors");
        CNVB_CUDA(ctx, cudaMemcpyAsync(d_order, order.data(), n_lanes * 4, cudaMemcpyHostToDevice, ctx->stream), "lanes");
        HmmArgs a{};
        a.cv = va.cv;
        a.lane_off = va.lane_off;
        a.order = static_cast<const uint32_t *>(d_order);
        a.tables = static_cast<const HmmTables *>(d_ptabs);
        a.n_lanes = n_lanes;
        a.alpha = static_cast<double *>(d_alpha);
        a.post = host ? static_cast<double *>(d_post) : pop;
        {
            KernelSpan span(ctx, "hmm_posterior_kernel");
            hmm_posterior_kernel<<<pblocks, 32, 0, ctx->stream>>>(a);
        }
        CNVB_LAUNCHED(ctx, "hmm_posterior_kernel");
        if (host)
            CNVB_CUDA(ctx, cudaMemcpyAsync(pop, d_post, total * K * 8, cudaMemcpyDeviceToHost, ctx->stream),
                      "downloading posteriors");
        CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "HMM posteriors");  // `order` is read by the copy above
    }
    if (host) CNVB_CUDA(ctx, cudaMemcpyAsync(stp, d_state, total, cudaMemcpyDeviceToHost, ctx->stream), "downloading states");
    // host-side lane tables / offsets are read by the copies above: wait before they go out of scope
    CNVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream), "HMM segmentation");
    return CNVB_OK;
}
