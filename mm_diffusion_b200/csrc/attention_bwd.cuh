// Flash-style attention BACKWARD on tcgen05/TMEM for sm_100a (training; gradient-guided
// sampling row a19).  Adjoint of attention.cuh for SingleModalQKVAttention / QKVAttention
// (mm_diffusion/multimodal_unet.py:212-244, :507-564) including the Random-Shift windows.
//
// With P = softmax(Q K^T / sqrt(d)), O = P V, delta = rowsum(dO o O), lse = log2-sum-exp of the scaled logits
// (written by the forward kernels):
//     dV = P^T dO          dP = dO V^T          dS = P o (dP - delta) / sqrt(d)
//     dQ = dS K            dK = dS^T Q
// Two instantiations of ONE kernel, each recomputing the logits (no atomics, deterministic):
//   KV = false  "dQ pass":  X tile = 128 queries, walks the key window Y.   S = Q_X K_Y^T, dP = dO_X V_Y^T,
//                           acc1 += dS K_Y                                  -> dQ
//   KV = true   "dKV pass": X tile = 128 keys, walks the window of queries Y that attend them.
//                           S^T = K_X Q_Y^T, dP^T = V_X dO_Y^T, acc1 += dS^T Q_Y -> dK, acc2 += P^T dO_Y -> dV
// Query block i attends key blocks (i + shift + j) mod n, j < win  <=>  key block k is attended by query blocks
// (k - shift - win + 1 + j) mod n, j < win: both walks are a contiguous token range modulo the per-sample count.
// Warp roles as in the forward: warps 0-3 softmax/gradient math (thread = X row), warp 4 TMA, warp 5 tcgen05.mma issue.
#pragma once
#include "attention.cuh"
#include "common.cuh"

namespace mmd {

struct alignas(64) AttnBwdParams {
    CUtensorMap x1_map, x2_map, y1_map, y2_map;   // 2-D [rows][ld] fp16, box (64 cols, 128 rows), SWIZZLE_128B
    int x1_col0, x2_col0, y1_col0, y2_col0;       // column of head 0 inside each view
    act_t* out1; int out1_ld; int out1_col0;      // acc1 -> dQ (KV = false) | dK (KV = true)
    act_t* out2; int out2_ld; int out2_col0;      // acc2 -> dV (KV = true)
    const float* lse;                             // [heads][stat_ld], indexed by absolute QUERY row
    const float* delta;                           // [heads][stat_ld]
    long long stat_ld;
    long long q_rows_total;
    int B, heads, n_blocks;
    int x_blk, x_per_batch;
    int y_blk, y_per_batch;
    int win;
    const int* shift_ptr;
    int x_tiles;
    float scale_log2;                             // d^-1/2 * log2(e)
    float rs;                                     // d^-1/2
};

constexpr int AB_THREADS = 320;   // warps 0-7 gradient math (two column halves x four lane quarters), warp 8 TMA, warp 9 MMA

template <int D>
struct AttnBwdSmem {
    static constexpr int NCH = (D + 63) / 64;
    static constexpr int TILE = NCH * 128 * 128;
    static constexpr int YST = (D == 64) ? 2 : 1;      // Y stages: the next Y tile streams in under the current one's math
    static constexpr int X1_OFF = 0;
    static constexpr int X2_OFF = TILE;
    static constexpr int Y_OFF = 2 * TILE;             // per stage: Y1 | Y2
    static constexpr int DS_OFF = Y_OFF + YST * 2 * TILE;   // 128 x 128 fp16 (two 64-column chunks)
    static constexpr int P_OFF = DS_OFF + 32768;
    static constexpr int STAT_OFF = P_OFF + 32768;     // lse[128] delta[128] floats (KV pass: per Y column)
    static constexpr int BAR_OFF = STAT_OFF + 1024;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
    static_assert(TOTAL <= 232448, "attention backward shared memory budget");
};

template <int D, bool KV>
__global__ void __launch_bounds__(AB_THREADS, 1) attn_bwd_kernel(const __grid_constant__ AttnBwdParams p) {
    using S = AttnBwdSmem<D>;
    constexpr int YST = S::YST;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* x_full = bars;        // X1 + X2 landed
    uint64_t* y_full = bars + 1;    // [2] Y1 + Y2 of a stage landed
    uint64_t* y_empty = bars + 3;   // [2] accumulate MMAs that read the stage retired (also frees dS / P)
    uint64_t* s_full = bars + 5;    // S and dP of the current tile are in TMEM
    uint64_t* p_ready = bars + 6;   // 256 arrivals: dS (and P) written, S / dP consumed
    uint64_t* acc_full = bars + 7;  // all accumulate MMAs retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    float* s_lse = reinterpret_cast<float*>(smem + S::STAT_OFF);
    float* s_delta = s_lse + 128;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- work decode (mirrors attn_decode with the roles of the two sides chosen by KV)
    int idx = blockIdx.x;
    const int xt = idx % p.x_tiles; idx /= p.x_tiles;
    const int head = idx % p.heads; idx /= p.heads;
    const int blk = idx % p.n_blocks;
    const int b = idx / p.n_blocks;
    const int x_row0 = b * p.x_per_batch + blk * p.x_blk + xt * 128;
    const int x_valid = min(128, p.x_blk - xt * 128);
    const int shift = p.shift_ptr ? *p.shift_ptr : 0;
    int start_blk;
    if (!KV) start_blk = (blk + shift) % p.n_blocks;
    else start_blk = ((blk - shift - p.win + 1) % p.n_blocks + p.n_blocks) % p.n_blocks;
    const int start = start_blk * p.y_blk;
    const int len = p.win * p.y_blk;
    const int len0 = min(len, p.y_per_batch - start);
    const int seg_row0 = b * p.y_per_batch + start, seg_row1 = b * p.y_per_batch;
    const int t0n = (len0 + 127) / 128;
    const int T = t0n + (len - len0 + 127) / 128;
    auto y_tile = [&](int t, int& row, int& valid) {
        if (t < t0n) { row = seg_row0 + t * 128; valid = min(128, len0 - t * 128); }
        else { const int u = t - t0n; row = seg_row1 + u * 128; valid = min(128, (len - len0) - u * 128); }
    };

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&p.x1_map);
        tma_prefetch_desc(&p.x2_map);
        tma_prefetch_desc(&p.y1_map);
        tma_prefetch_desc(&p.y2_map);
        mbar_init(x_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 1); }
        mbar_init(s_full, 1);
        mbar_init(p_ready, 256);
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 9) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base;            // 128 columns
    const uint32_t tmem_DP = tmem_base + 128;     // 128 columns
    const uint32_t tmem_A1 = tmem_base + 256;     // D columns
    const uint32_t tmem_A2 = tmem_base + 384;     // D columns (KV)

    if (warp == 8) {
        // ===================== TMA producer (uniform warp, elected lane issues; see gemm.cuh) =====================
        {
            if (elect_one()) {
                mbar_expect_tx(x_full, 2 * S::TILE);
                for (int ch = 0; ch < S::NCH; ++ch) {
                    tma_load_2d(smem + S::X1_OFF + ch * 16384, &p.x1_map, x_full, p.x1_col0 + head * D + ch * 64, x_row0);
                    tma_load_2d(smem + S::X2_OFF + ch * 16384, &p.x2_map, x_full, p.x2_col0 + head * D + ch * 64, x_row0);
                }
            }
            __syncwarp();
            for (int t = 0; t < T; ++t) {
                const int st = t % YST;
                const uint32_t use = static_cast<uint32_t>(t / YST);
                int yrow, yvalid;
                y_tile(t, yrow, yvalid);
                mbar_wait(&y_empty[st], (use & 1) ^ 1);
                uint8_t* y1 = smem + S::Y_OFF + st * 2 * S::TILE;
                if (elect_one()) {
                    mbar_expect_tx(&y_full[st], 2 * S::TILE);
                    for (int ch = 0; ch < S::NCH; ++ch) {
                        tma_load_2d(y1 + ch * 16384, &p.y1_map, &y_full[st], p.y1_col0 + head * D + ch * 64, yrow);
                        tma_load_2d(y1 + S::TILE + ch * 16384, &p.y2_map, &y_full[st], p.y2_col0 + head * D + ch * 64, yrow);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer (uniform warp, elected lane issues) =====================
        {
            constexpr uint32_t idesc_s = umma_idesc_f16(128, 128, 0, 0);
            constexpr uint32_t idesc_acc = umma_idesc_f16(128, D, 0, 1);   // B (Y tile) is MN-major
            const uint32_t x1 = smem_u32(smem + S::X1_OFF), x2 = smem_u32(smem + S::X2_OFF);
            const uint64_t dsd0 = umma_desc_sw128(smem_u32(smem + S::DS_OFF), 16, 1024);
            const uint64_t pd0 = umma_desc_sw128(smem_u32(smem + S::P_OFF), 16, 1024);
            auto issue_s = [&](int t) {   // S = X1 Y1^T, dP = X2 Y2^T of tile t
                const int st = t % YST;
                mbar_wait(&y_full[st], (t / YST) & 1);
                tc_fence_after();
                const uint32_t y1 = smem_u32(smem + S::Y_OFF + st * 2 * S::TILE), y2 = y1 + S::TILE;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks) {
                        const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
                        umma_f16_ss(tmem_S, umma_desc_sw128(x1 + off, 16, 1024), umma_desc_sw128(y1 + off, 16, 1024), idesc_s,
                                    ks != 0 ? 1u : 0u);
                    }
#pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks) {
                        const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
                        umma_f16_ss(tmem_DP, umma_desc_sw128(x2 + off, 16, 1024), umma_desc_sw128(y2 + off, 16, 1024), idesc_s,
                                    ks != 0 ? 1u : 0u);
                    }
                    umma_commit(s_full);
                }
                __syncwarp();
            };
            mbar_wait(x_full, 0);
            issue_s(0);
            for (int t = 0; t < T; ++t) {
                const int st = t % YST;
                int yrow, yvalid;
                y_tile(t, yrow, yvalid);
                mbar_wait(p_ready, t & 1);
                tc_fence_after();
                const uint32_t y1 = smem_u32(smem + S::Y_OFF + st * 2 * S::TILE);
                const uint64_t y1m = umma_desc_sw128(y1, 128 * 128, 1024);   // MN-major views of the Y tiles
                const uint64_t y2m = umma_desc_sw128(y1 + S::TILE, 128 * 128, 1024);
                const int nks = (yvalid + 15) >> 4;
                if (elect_one()) {
                    for (int ks = 0; ks < nks; ++ks) {
                        const uint64_t aoff = static_cast<uint64_t>((ks >> 2) * (16384 >> 4) + (ks & 3) * 2);
                        umma_f16_ss(tmem_A1, dsd0 + aoff, y1m + static_cast<uint64_t>(ks) * (2048 >> 4), idesc_acc, (t | ks) != 0 ? 1u : 0u);
                        if (KV)
                            umma_f16_ss(tmem_A2, pd0 + aoff, y2m + static_cast<uint64_t>(ks) * (2048 >> 4), idesc_acc, (t | ks) != 0 ? 1u : 0u);
                    }
                    umma_commit(&y_empty[st]);
                    if (t == T - 1) umma_commit(acc_full);
                }
                __syncwarp();
                if (t != T - 1) issue_s(t + 1);   // S / dP are free (p_ready(t)); queued right behind the accumulate MMAs
            }
        }
    } else {
        // ===================== gradient math (thread = X row, half of the Y columns) =====================
        const int quarter = warp & 3, half = warp >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        const float* lse_h = p.lse + static_cast<size_t>(head) * p.stat_ld;
        const float* delta_h = p.delta + static_cast<size_t>(head) * p.stat_ld;
        float my_lse = 0.f, my_delta = 0.f;
        if (!KV) {
            const long long qr = min(static_cast<long long>(x_row0 + row), p.q_rows_total - 1);
            my_lse = lse_h[qr];
            my_delta = delta_h[qr];
        }
        uint8_t* ds_chunk = smem + S::DS_OFF + half * 16384;   // this half's 64-column chunk
        uint8_t* p_chunk = smem + S::P_OFF + half * 16384;
        for (int t = 0; t < T; ++t) {
            int yrow, yvalid;
            y_tile(t, yrow, yvalid);
            if (KV) {
                // per-column (query) statistics of this Y tile; the previous tile's readers are past their p_ready arrival
                named_bar_sync(1, 256);
                if (half == 0) {
                    const long long qr = min(static_cast<long long>(yrow + row), p.q_rows_total - 1);
                    s_lse[row] = lse_h[qr];
                    s_delta[row] = delta_h[qr];
                }
                named_bar_sync(1, 256);
            }
            mbar_wait(s_full, t & 1);
            tc_fence_after();
            // dS / P buffers are free once the accumulate MMAs of the previous tile have retired
            if (t > 0) mbar_wait(&y_empty[(t - 1) % YST], ((t - 1) / YST) & 1);
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int c = half * 2 + cc;   // 32-column chunk of the tile
                uint32_t sv[32], dv[32];
                tmem_ld32(tmem_S + lane_base + c * 32, sv);
                tmem_ld32(tmem_DP + lane_base + c * 32, dv);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 dpk, ppk;
                    uint32_t* dw = reinterpret_cast<uint32_t*>(&dpk);
                    uint32_t* pw = reinterpret_cast<uint32_t*>(&ppk);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int col = c * 32 + j * 8 + 2 * k;
                        float l0, l1, d0, d1;
                        if (KV) { l0 = s_lse[col]; l1 = s_lse[col + 1]; d0 = s_delta[col]; d1 = s_delta[col + 1]; }
                        else { l0 = l1 = my_lse; d0 = d1 = my_delta; }
                        float p0 = ex2_fast(fmaf(__uint_as_float(sv[j * 8 + 2 * k]), p.scale_log2, -l0));
                        float p1 = ex2_fast(fmaf(__uint_as_float(sv[j * 8 + 2 * k + 1]), p.scale_log2, -l1));
                        if (col >= yvalid) p0 = 0.f;
                        if (col + 1 >= yvalid) p1 = 0.f;
                        const float g0 = p0 * (__uint_as_float(dv[j * 8 + 2 * k]) - d0) * p.rs;
                        const float g1 = p1 * (__uint_as_float(dv[j * 8 + 2 * k + 1]) - d1) * p.rs;
                        const __half2 hg = __floats2half2_rn(g0, g1);
                        dw[k] = *reinterpret_cast<const uint32_t*>(&hg);
                        if (KV) {
                            const __half2 hp = __floats2half2_rn(p0, p1);
                            pw[k] = *reinterpret_cast<const uint32_t*>(&hp);
                        }
                    }
                    *reinterpret_cast<uint4*>(ds_chunk + sw128_off(row, cc * 4 + j)) = dpk;
                    if (KV) *reinterpret_cast<uint4*>(p_chunk + sw128_off(row, cc * 4 + j)) = ppk;
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_ready);
        }
        // ---- epilogue: accumulators -> fp16 gradients (each half stores D / 2 columns of its row)
        mbar_wait(acc_full, 0);
        tc_fence_after();
        constexpr int HC = D / 2;
#pragma unroll 1
        for (int which = 0; which < (KV ? 2 : 1); ++which) {
            const uint32_t t_acc = (which == 0 ? tmem_A1 : tmem_A2) + lane_base + half * HC;
            act_t* obase = (which == 0) ? p.out1 : p.out2;
            const int old = (which == 0) ? p.out1_ld : p.out2_ld;
            const int ocol = (which == 0) ? p.out1_col0 : p.out2_col0;
            act_t* orow = obase + static_cast<size_t>(x_row0 + row) * old + ocol + head * D + half * HC;
#pragma unroll
            for (int c = 0; c < HC / 16; ++c) {
                uint32_t v[16];
                tmem_ld16(t_acc + c * 16, v);
                tmem_ld_wait();
                if (row < x_valid) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        uint4 pk;
                        __half2* ph2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            ph2[k] = __floats2half2_rn(__uint_as_float(v[j * 8 + 2 * k]), __uint_as_float(v[j * 8 + 2 * k + 1]));
                        *reinterpret_cast<uint4*>(orow + c * 16 + j * 8) = pk;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        __syncwarp();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace mmd
