// wis_resident.cuh -- the sweep for panels of up to 125 samples (S + 3 <= 128): the block's 256 rows are
// the A operand of the MMAs and live in TENSOR MEMORY for the whole launch.  Included by wis_tensor.cuh.
//
// What the measurements of the shared-memory form (wis_sweep_kernel<.., STREAM = false, ..>, removed) said:
//   * an SS-mode M128 x N128 x K16 tcgen05.mma reads 8 KB of operands from shared memory per 64-cycle slot --
//     the SM's whole 128 B/clk, with the TMA writes of the next tile on top: 135 cycles per MMA measured,
//     1890 per 256 x 128 tile (14 MMAs) with the epilogue switched off;
//   * the epilogue (tcgen05.ld, one FFMA with n'_j and one compare per pair, group predicates) was 5160 warp
//     instructions per tile = 1290 issue cycles per scheduler at best, 2100 measured alone;
//   * both share two accumulator stages of a whole tile each, so every stall of one side reaches the other:
//     2660 cycles per tile in rounds that append nothing, 4000 - 17 000 in the rounds that append.
// Here:
//   * TS-mode MMAs: A = the rows, from TMEM (written once per CTA with tcgen05.st: lane = row, 32-bit column c =
//     K elements 2 c, 2 c + 1), B = the column tile from shared memory, 4 KB per 64-cycle slot -- half the
//     shared-memory bandwidth, and all 192 KB of it is B ring (12 stages = 6 tiles in flight);
//   * the column norms are FOLDED into the contraction: rows hold (-2 h_i | C, C, C), columns
//     (h_j | hi, mid, lo) with hi + mid + lo = n'_j / C exactly (three fp16 hold the 24 bits of an f32; C = 2^15
//     keeps them in range), so the accumulator IS the value under test, v_ij = n'_j - 2 g_ij.  No column operands
//     in the epilogue: a tree of three-input minima over the 32 values of a row (16 FMNMX3 / FMNMX), ONE compare
//     with the row's threshold, one vote per warp.  The products C x part are exact; the fp32 accumulation sees
//     addends of total magnitude 2 s_i s_j + n_j instead of 2 s_i s_j, which the bounds carry as c_acc n_j
//     (BoundConsts::c_rnd is raised by c_acc on this path);
//   * the unit of the accumulator ring is a HALF tile (128 rows x 128 columns, 128 TMEM columns), three of them
//     beside the 128 columns of the operand: while the epilogue warps hold one, the tensor core fills the next
//     and has a third to go on to.  An epilogue warp gives its stage back as soon as tcgen05.ld has delivered the
//     values to its registers -- before the tests and appends.
#pragma once

namespace {

constexpr uint32_t kResStages = 12;   // B ring: 12 x 16 KB (one 128-column x 64-K block each)
constexpr uint32_t kResAcc = 3;       // accumulator ring: half tiles of 128 TMEM columns
constexpr uint32_t kResACols = 64;    // TMEM columns of one row half's operand: 128 K elements, two per column
constexpr uint32_t kResK = 128;       // padded K of this path
constexpr float kFoldC = 32768.0f;    // the rows' three fold entries; the columns' are n'_j / C in three parts

__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// the fold entries of the column operand: n'_j / C in three fp16 parts whose sum never exceeds it
// (+inf beyond T: such a column never passes a test)
__global__ void wis_fold_cols_kernel(const float *__restrict__ nprime, uint32_t T, uint32_t Tpad, uint32_t S,
                                     __half *__restrict__ H) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Tpad) return;
    __half *row = H + (size_t)j * kResK + S;
    if (j >= T) {
        row[0] = __ushort_as_half((unsigned short)0x7C00u);
        row[1] = __float2half_rn(0.0f);
        row[2] = __float2half_rn(0.0f);
        return;
    }
    const float x = nprime[j] * (1.0f / kFoldC);  // exact: a power of two
    const __half hi = __float2half_rd(x);         // (rounded down: the remainders stay non-negative)
    const float r1 = __fsub_rn(x, __half2float(hi));  // exact
    const __half mid = __float2half_rd(r1);
    const float r2 = __fsub_rn(r1, __half2float(mid));  // exact
    row[0] = hi;
    row[1] = mid;
    row[2] = __float2half_rd(r2);
}

template <bool FIRST>
__global__ void __launch_bounds__(kSweepThreads, 1)
wis_sweep_res_kernel(const __grid_constant__ CUtensorMap tmap_b, const SweepArgs a) {
    const uint32_t blk = blockIdx.x;
    // this block's share of the round: two tile ranges either side of what it has swept before
    const uint4 rng = a.rng[blk];
    const uint32_t n_lo = rng.y - rng.x, n_tiles = n_lo + (rng.w - rng.z);
    if (n_tiles == 0) return;
    const uint32_t rng_x = rng.x, rng_z = rng.z;
#define CNVB_RTILE(it) ((it) < n_lo ? rng_x + (it) : rng_z + ((it) - n_lo))
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint32_t s_cnt[kTileRows];  // candidates collected so far per row of this CTA
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_base = smem_base;
    const uint32_t bar_base = b_base + kResStages * kTileBytes;
    const uint32_t bar_b_full = bar_base;                          // kResStages
    const uint32_t bar_b_empty = bar_b_full + 8 * kResStages;      // kResStages
    const uint32_t bar_acc_full = bar_b_empty + 8 * kResStages;    // kResAcc
    const uint32_t bar_acc_empty = bar_acc_full + 8 * kResAcc;     // kResAcc
    const uint32_t tmem_slot = bar_acc_empty + 8 * kResAcc;
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t row0 = blk * kTileRows;  // position in the sorted row list
    for (uint32_t t = threadIdx.x; t < kTileRows; t += blockDim.x)
        s_cnt[t] = row0 + t < a.rows ? a.cand_cnt[row0 + t] : 0u;
    if (warp == 0 && lane == 0) {
        for (uint32_t s = 0; s < kResStages; ++s) {
            mbar_init(bar_b_full + 8 * s, 1);
            mbar_init(bar_b_empty + 8 * s, 1);
        }
        for (uint32_t s = 0; s < kResAcc; ++s) {
            mbar_init(bar_acc_full + 8 * s, 1);
            mbar_init(bar_acc_empty + 8 * s, kEpiWarps);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: 2 x 64 operand columns | 3 x 128 accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const uint32_t acc_base = tmem_base + 2u * kResACols;

    // TMA producer (one thread), first part: as many tiles as the empty ring takes are on their way while the
    // operand is written (no waits here: this thread has to reach the barrier below)
    uint32_t p_st = 0, p_ph = 0, p_it = 0;
    auto produce = [&](uint32_t it_end) {
        for (; p_it < it_end; ++p_it) {
            const int32_t j0 = (int32_t)(CNVB_RTILE(p_it) * kTileCols);
            for (uint32_t kb = 0; kb < 2; ++kb) {
                mbar_wait(bar_b_empty + 8 * p_st, p_ph ^ 1u);
                mbar_expect_tx(bar_b_full + 8 * p_st, kTileBytes);
                tma_load_2d(b_base + p_st * kTileBytes, &tmap_b, bar_b_full + 8 * p_st, (int32_t)(kb * kKBlock), j0);
                if (++p_st == kResStages) {
                    p_st = 0;
                    p_ph ^= 1u;
                }
            }
        }
    };
    if (warp == 0 && lane == 0) produce(min(n_tiles, kResStages / 2u));
    if (warp >= 4) {
        // the rows' operand: epilogue warp e writes TMEM lane quarter e & 3 of row half (e >> 2) & 1, the 32
        // columns (64 K elements) e >> 3 -- every thread its own row, 128 bytes of it
        const uint32_t e = warp - 4, q = e & 3u, h = (e >> 2) & 1u, kh = e >> 3;
        const uint32_t row = row0 + h * 128u + q * 32u + lane;
        uint32_t w[32];
        if (row < a.rows) {
            const uint4 *src = reinterpret_cast<const uint4 *>(a.HA + (size_t)(a.rows_a ? a.rows_a[row] : row) * kResK + kh * 64u);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint4 x = __ldg(src + i);
                w[4 * i + 0] = x.x;
                w[4 * i + 1] = x.y;
                w[4 * i + 2] = x.z;
                w[4 * i + 3] = x.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) w[i] = 0u;
        }
        tc_st32(tmem_base + ((q * 32u) << 16) + h * kResACols + kh * 32u, w);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        if (lane == 0) produce(n_tiles);  // ===== TMA producer, the rest =====
    } else if (warp == 1) {
        // ===== MMA issuer (one thread): per tile and row half seven (ceil((S + 3) / 16)) M128 x N128 x K16 MMAs =====
        if (lane == 0) {
            uint32_t st = 0, ph = 0, as = 0, aph = 0;
            const uint64_t bdesc0 = umma_desc(b_base);
            for (uint32_t it = 0; it < n_tiles; ++it) {
#pragma unroll
                for (uint32_t h = 0; h < 2; ++h) {
                    mbar_wait(bar_acc_empty + 8 * as, aph ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = acc_base + as * kTileCols;
#pragma unroll
                    for (uint32_t kb = 0; kb < 2; ++kb) {
                        const uint32_t sb = st + kb;  // (kResStages is even: a tile's two stages never wrap apart)
                        if (h == 0) {
                            mbar_wait(bar_b_full + 8 * sb, ph);
                            tc_fence_after();
                        }
                        const uint64_t bdesc = bdesc0 + (uint64_t)(sb * (kTileBytes >> 4));
#pragma unroll
                        for (uint32_t kk = 0; kk < kKBlock / 16; ++kk) {
                            if (kb * (kKBlock / 16) + kk < a.k_mmas && !(a.dbg & 2u))
                                tc_mma_f16_ts(d_tmem, tmem_base + h * kResACols + (kb * (kKBlock / 16) + kk) * 8u, bdesc + 2u * kk,
                                              kIdesc, (kb | kk) != 0u);
                        }
                        if (h == 1) tc_commit(bar_b_empty + 8 * sb);  // frees the B stage when these MMAs retire
                    }
                    tc_commit(bar_acc_full + 8 * as);  // this half tile's accumulators complete
                    if (++as == kResAcc) {
                        as = 0;
                        aph ^= 1u;
                    }
                }
                st += 2;
                if (st == kResStages) {
                    st = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: 16 warps.  Warp e reads TMEM lane quarter q = e & 3 (rows 32 q + lane of the half in
        // the stage) and the 32 columns [32 (e >> 2), +32) of every half tile. =====
        const uint32_t e = warp - 4, q = e & 3u, cpart = e >> 2;
        float thr[2], kap[2];
        uint32_t trow[2];  // FIRST: contig of the row (a value no column has for rows beyond the end)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t row = row0 + h * 128 + q * 32 + lane;
            const bool live = row < a.rows;
            thr[h] = live ? a.thr[row] : -__int_as_float(0x7f800000);
            kap[h] = live ? a.kap[row] : 0.0f;
            trow[h] = 0xFFFFFFFEu;
            if (FIRST && live) trow[h] = a.tid_sorted[a.rows_a ? a.rows_a[row] : row];
        }
        const float inf = __int_as_float(0x7f800000);
        uint32_t as = 0, aph = 0;
        for (uint32_t it = 0; it < n_tiles; ++it) {
            const uint32_t tile = CNVB_RTILE(it);
            const uint32_t j0 = tile * kTileCols + cpart * 32u;
            const float smax = __ldg(a.tile_smax + tile);
            uint32_t tj[FIRST ? 32 : 1];
            if (FIRST) {  // contigs of this warp's 32 columns (same addresses for all lanes: L1 broadcast)
                const uint4 *tj4 = reinterpret_cast<const uint4 *>(a.tid_sorted + j0);
#pragma unroll
                for (int g4 = 0; g4 < (FIRST ? 8 : 0); ++g4) {
                    const uint4 t4 = __ldg(tj4 + g4);
                    tj[4 * g4 + 0] = t4.x;
                    tj[4 * g4 + 1] = t4.y;
                    tj[4 * g4 + 2] = t4.z;
                    tj[4 * g4 + 3] = t4.w;
                }
                if (it + 1 < n_tiles)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.tid_sorted + CNVB_RTILE(it + 1) * kTileCols + cpart * 32u));
            }
            // (rows without a bound yet have rhs = +inf: the padding columns, v = +inf, must not pass)
            const uint32_t valid = j0 + 32u <= a.T ? 0xFFFFFFFFu : (j0 < a.T ? (1u << (a.T - j0)) - 1u : 0u);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                mbar_wait(bar_acc_full + 8 * as, aph);
                tc_fence_after();
                uint32_t r[32];
                if (!(a.dbg & 1u)) tc_ld32(acc_base + ((q * 32u) << 16) + as * kTileCols + cpart * 32u, r);
                // the values are in registers: the stage goes back to the MMA issuer before the tests and appends
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_acc_empty + 8 * as);
                if (++as == kResAcc) {
                    as = 0;
                    aph ^= 1u;
                }
                if (a.dbg & 1u) continue;
                float v[32];  // v_c = n'_j - 2 g_ij, straight from the accumulator
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(r[c]);
                const uint32_t rloc = (uint32_t)h * 128u + q * 32u + lane;
                const uint32_t row = row0 + rloc;
                uint2 *cand = a.cand + (size_t)row * kCandCap;  // (rows beyond a.rows never pass: thr = -inf)
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
                        // (v = +inf beyond T: padding columns are never a group's minimum)
                        uint32_t slot = atomicAdd(&s_cnt[rloc], (uint32_t)__popc(have));
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
                    uint32_t mask = 0u;
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) {
                        if (wmask & (1u << g4)) {  // (warp-uniform)
#pragma unroll
                            for (int e4 = 0; e4 < 4; ++e4) mask |= v[4 * g4 + e4] <= rhs ? (1u << (4 * g4 + e4)) : 0u;
                        }
                    }
                    mask &= valid;
                    // one counter update per row, warp and half tile: the row's passing columns take consecutive slots
                    const uint32_t cnt = (uint32_t)__popc(mask);
                    uint32_t slot = 0u;
                    if (cnt) slot = atomicAdd(&s_cnt[rloc], cnt);
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
        }
    }
    tc_fence_before();
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < kTileRows; t += blockDim.x)
        if (row0 + t < a.rows) a.cand_cnt[row0 + t] = s_cnt[t];
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

#undef CNVB_RTILE

}  // namespace
