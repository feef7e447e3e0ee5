// wis_resident.cuh -- the sweep for panels of up to 125 samples (S + 3 <= 128): the block's 256 rows stay in
// shared memory for the whole launch, the column tiles stream past.  Included by wis_tensor.cuh.
//
// What the measurements of the first form of this kernel said (per-round trace, CNVB_WIS_TRACE / CNVB_WIS_DBG;
// 200 000 targets x 100 samples):
//   * its epilogue (tcgen05.ld, one FFMA with n'_j and one compare per pair, group predicates) was 5160 warp
//     instructions per 256 x 128 tile = 1290 issue cycles per scheduler at best, 2100 measured alone, against
//     1890 for the 14 MMAs alone (135 cycles per M128 x N128 x K16 instruction, SS mode);
//   * MMA issuer and epilogue shared two accumulator stages of a whole tile each, so every stall of one side
//     reached the other: 2660 cycles per tile in rounds that append nothing;
//   * the rounds that append were bound by the stores: every candidate cost two scattered 4-byte stores (value,
//     index), and an SM retires ONE lane store to a distinct line per clock: 1.4 ms for 88 M candidates.
// Here:
//   * the column norms are FOLDED into the contraction: rows hold (-2 h_i | C, C, C), columns
//     (h_j | hi, mid, lo) with hi + mid + lo = n'_j / C exactly (three fp16 hold the 24 bits of an f32; C = 2^15
//     keeps them in range), so the accumulator IS the value under test, v_ij = n'_j - 2 g_ij.  No column operands
//     in the epilogue: a tree of three-input minima over the 32 values of a row (FMNMX3 / FMNMX), ONE compare
//     with the row's threshold, one vote per warp.  The products C x part are exact; the fp32 accumulation sees
//     addends of total magnitude 2 s_i s_j + n_j instead of 2 s_i s_j, which the bounds carry as c_acc n_j
//     (BoundConsts::c_rnd is raised by c_acc on this path);
//   * the unit of the accumulator ring is a HALF tile (128 rows x 128 columns, 128 TMEM columns), four of them:
//     while the epilogue warps hold one, the tensor core fills the next and has two more to go on to.  An
//     epilogue warp gives its stage back as soon as tcgen05.ld has delivered the values to its registers --
//     before the tests and appends;
//   * candidates are (value, index) pairs, one 8-byte store each; the first window stores two per 16-byte store.
// An MMA of this kind costs about 64 + N / 2 cycles whatever feeds it: 133 per M128 x N128 x K16 from shared or
// tensor memory once consecutive MMAs go to different accumulators, 192 per M128 x N256 x K16.  The transposed
// M128 x N256 tile (1460 cycles per tile for the MMAs alone against 1890) was built on this kernel too and is
// SLOWER in the sweep: a thread then owns a column and meets 32 rows with 32 different thresholds per tcgen05.ld,
// and its tcgen05.ld traffic does not overlap the N = 256 MMAs -- 2700 - 2950 cycles per tile with both, the sum
// of the two alone (1460 + 1490 - 1640), against 2150 - 2500 here.
// Tried and measured slower: the rows' operand in TENSOR MEMORY (TS-mode tcgen05.mma, written once per CTA with
// tcgen05.st; bit-exact): 2240 - 2400 cycles per tile for the MMAs alone against 1890 - 2020 from shared memory --
// the operand reads compete with the accumulators' read-modify-write for the TMEM ports.
#pragma once

namespace {

constexpr uint32_t kResStages = 8;    // B ring: 8 x 16 KB (one 128-column x 64-K block each)
constexpr uint32_t kResAcc = 4;       // accumulator ring: half tiles of 128 TMEM columns
constexpr uint32_t kResK = 128;       // padded K of this path
constexpr float kFoldC = 32768.0f;    // the rows' three fold entries; the columns' are n'_j / C in three parts

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
wis_sweep_res_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const SweepArgs a) {
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
    const uint32_t a_base = smem_base;  // the rows' operand: [K block][row half] x 16 KB
    const uint32_t b_base = a_base + 4u * kTileBytes;
    const uint32_t bar_base = b_base + kResStages * kTileBytes;
    const uint32_t bar_a_full = bar_base + 8 * (2 * kResStages + 2 * kResAcc) + 8;
    const uint32_t bar_b_full = bar_base;                          // kResStages
    const uint32_t bar_b_empty = bar_b_full + 8 * kResStages;      // kResStages
    const uint32_t bar_acc_full = bar_b_empty + 8 * kResStages;    // kResAcc
    const uint32_t bar_acc_empty = bar_acc_full + 8 * kResAcc;     // kResAcc
    const uint32_t tmem_slot = bar_acc_empty + 8 * kResAcc;
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t row0 = blk * kTileRows;  // position in the sorted row list
    const uint32_t n_kb = (a.k_mmas + 3u) / 4u;  // K blocks of 64 that hold data (S + 3 <= 64: one)
    for (uint32_t t = threadIdx.x; t < kTileRows; t += blockDim.x)
        s_cnt[t] = row0 + t < a.rows ? a.cand_cnt[row0 + t] : 0u;
    if (warp == 0 && lane == 0) {
        mbar_init(bar_a_full, 1);
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
    if (warp == 1) {  // TMEM: 4 x 128 accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const uint32_t acc_base = tmem_base;

    if (warp == 0) {
        // ===== TMA producer (one thread): the rows' operand once, then the column tiles =====
        if (lane == 0) {
            mbar_expect_tx(bar_a_full, 2u * n_kb * kTileBytes);
            for (uint32_t kb = 0; kb < n_kb; ++kb)
                for (uint32_t h = 0; h < 2; ++h)
                    tma_load_2d(a_base + (kb * 2u + h) * kTileBytes, &tmap_a, bar_a_full, (int32_t)(kb * kKBlock), (int32_t)(row0 + h * 128));
            uint32_t st = 0, ph = 0;
            for (uint32_t it = 0; it < n_tiles; ++it) {
                const int32_t j0 = (int32_t)(CNVB_RTILE(it) * kTileCols);
                for (uint32_t kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(bar_b_empty + 8 * st, ph ^ 1u);
                    mbar_expect_tx(bar_b_full + 8 * st, kTileBytes);
                    tma_load_2d(b_base + st * kTileBytes, &tmap_b, bar_b_full + 8 * st, (int32_t)(kb * kKBlock), j0);
                    if (++st == kResStages) {
                        st = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread): per tile and row half seven (ceil((S + 3) / 16)) M128 x N128 x K16 MMAs =====
        if (lane == 0) {
            // The two halves of a tile alternate MMA by MMA: an MMA into the accumulator the previous one wrote
            // waits for it (seven in a row: 160 cycles each; alternating: 133), so a tile takes two stages of
            // the ring at a time and both complete together.
            uint32_t st = 0, ph = 0, u = 0;
            const uint64_t bdesc0 = umma_desc(b_base), adesc0 = umma_desc(a_base);
            mbar_wait(bar_a_full, 0);
            tc_fence_after();
            for (uint32_t it = 0; it < n_tiles; ++it) {
                const uint32_t as0 = u % kResAcc, as1 = (u + 1u) % kResAcc, aph = (u / kResAcc) & 1u;  // (kResAcc is even: same lap)
                u += 2u;
                mbar_wait(bar_acc_empty + 8 * as0, aph ^ 1u);
                mbar_wait(bar_acc_empty + 8 * as1, aph ^ 1u);
                tc_fence_after();
#pragma unroll
                for (uint32_t kb = 0; kb < 2; ++kb) {
                    if (kb < n_kb) {
                        mbar_wait(bar_b_full + 8 * st, ph);
                        tc_fence_after();
                        const uint64_t bdesc = bdesc0 + (uint64_t)(st * (kTileBytes >> 4));
#pragma unroll
                        for (uint32_t kk = 0; kk < kKBlock / 16; ++kk) {
#pragma unroll
                            for (uint32_t h = 0; h < 2; ++h) {
                                if (kb * (kKBlock / 16) + kk < a.k_mmas && !(a.dbg & 2u))
                                    tc_mma_f16(acc_base + (h ? as1 : as0) * kTileCols,
                                               adesc0 + (uint64_t)((kb * 2u + h) * (kTileBytes >> 4)) + 2u * kk, bdesc + 2u * kk, kIdesc,
                                               (kb | kk) != 0u);
                            }
                        }
                        tc_commit(bar_b_empty + 8 * st);  // frees the B stage when these MMAs retire
                        if (++st == kResStages) {
                            st = 0;
                            ph ^= 1u;
                        }
                    }
                }
                tc_commit(bar_acc_full + 8 * as0);  // both half tiles' accumulators complete
                tc_commit(bar_acc_full + 8 * as1);
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: 16 warps.  Warp e reads TMEM lane quarter q = e & 3 (rows 32 q + lane of the half in
        // the stage) and the 32 columns [32 (e >> 2), +32) of every half tile. =====
        const uint32_t e = warp - 4, q = e & 3u, cpart = e >> 2;
        float thr[2], kap[2];
        uint32_t trow[2];  // FIRST: contig of the row (a value no column has for rows beyond the end)
        uint32_t own[2];   // the row's own column (sorted position): a same-contig index for padding entries
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t row = row0 + h * 128 + q * 32 + lane;
            const bool live = row < a.rows;
            thr[h] = live ? a.thr[row] : -__int_as_float(0x7f800000);
            kap[h] = live ? a.kap[row] : 0.0f;
            trow[h] = 0xFFFFFFFEu;
            own[h] = live ? (a.rows_a ? a.rows_a[row] : row) : 0u;
            if (FIRST && live) trow[h] = a.tid_sorted[own[h]];
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
                    // Every (row, warp, half tile) stores exactly eight entries -- groups without a finite minimum as
                    // padding (the row's own column: same contig, dropped by the compaction) -- so the slots are
                    // multiples of eight and two entries go out per 16-byte store.
                    if (row < a.rows) {
                        uint32_t pick = 0u;
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) {
                            const uint32_t w = v[4 * g4] == gm[g4] ? 0u : (v[4 * g4 + 1] == gm[g4] ? 1u : (v[4 * g4 + 2] == gm[g4] ? 2u : 3u));
                            pick |= w << (2 * g4);
                        }
                        const uint32_t slot = atomicAdd(&s_cnt[rloc], 8u);
                        if (slot + 8u <= kCandCap) {
                            uint4 *dst = reinterpret_cast<uint4 *>(cand + slot);
#pragma unroll
                            for (int g2 = 0; g2 < 4; ++g2) {
                                const int ga = 2 * g2, gb = 2 * g2 + 1;
                                const uint32_t ja = gm[ga] < inf ? j0 + 4u * ga + ((pick >> (2 * ga)) & 3u) : own[h];
                                const uint32_t jb = gm[gb] < inf ? j0 + 4u * gb + ((pick >> (2 * gb)) & 3u) : own[h];
                                dst[g2] = make_uint4(__float_as_uint(gm[ga]), ja, __float_as_uint(gm[gb]), jb);
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
