// Weight-gradient kernel of the implicit-GEMM convolutions (training backward):
//
//   dW[n, (tap, src, c)] = sum_{token} dY[token, n] * A_src[token + delta(tap), c]
//
// i.e. the reduction runs over the TOKEN axis, which is the slow (row) axis of both channels-last operands, so
// both tcgen05 operands are MN-major: one 128-token TMA box of dY is the A operand (M = output channels),
// the tap-shifted box of the activation is the B operand (N = input channels), 16 tokens per UMMA K-step.
// The geometry (tensor-map rank, token box, taps as coordinate offsets, out-of-bound zero fill = "same"
// padding) is exactly the forward kernel's (gemm.cuh), so a forward GemmProblem describes its own wgrad.
//
// Work item = (128-row tile of n) x (128-column block of one (tap, source)) x (token split); a CTA accumulates its
// token range in TMEM (fp32) and adds the tile into the packed fp32 gradient [n][k_total] with vector atomics
// (plain stores when there is one split).  Persistent, warp-specialised: warp 0 TMA, warp 1 MMA, warps 2-5 epilogue.
// Reference semantics: torch.autograd of F.conv{1,2,3}d as used by VideoConv / AudioConv
// (mm_diffusion/multimodal_unet.py:68-131).
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace mmd {

constexpr int WG_THREADS = 192;
constexpr int WG_STAGES = 3;
constexpr int WG_DY_BYTES = 2 * GEMM_BM * 128;   // two 64-channel chunks of 128 tokens
constexpr int WG_A_BYTES = 2 * GEMM_BM * 128;
constexpr int WG_STAGE_BYTES = WG_DY_BYTES + WG_A_BYTES;
constexpr int WG_ONES_BYTES = 2048;   // 16 tokens x 128 B of fp16 ones: B operand of the bias-gradient MMA (any layout)
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + WG_ONES_BYTES + 256 + 1024;

struct alignas(64) WgradParams {
    CUtensorMap dy_map;               // dY [tokens][n], same token geometry as the activations, box (64, box[0..3])
    CUtensorMap a_map[GEMM_MAX_SRC];  // activation sources, box (64, box[0..3])
    int n_src;
    int src_chunks[GEMM_MAX_SRC];
    int rank;
    int box[4];
    int ntile[4];
    int n_taps;
    int tap[GEMM_MAX_TAPS][3];
    int m_tiles;          // 128-token tiles
    int n_tiles;          // 128-row tiles of n
    int n;                // valid output channels
    int blocks_per_tap;   // 128-column blocks per tap (sum over sources of ceil(chunks / 2))
    int total_chunks;     // 64-channel chunks per tap (all sources)
    int splits;           // token splits
    float* dw;            // [n][ld] fp32, packed column order k = (tap * total_chunks + chunk) * 64 + c
    long long ld;
    float* db;            // [n] fp32 bias gradient = column sums of dY (null = off), accumulated with atomics
};

struct WgradItem {
    int n_tile, tap, src, chunk0, ncols, col0, m_begin, m_end;
    bool bias;   // this item also reduces dY over its tokens (first column block of every (n tile, split))
};

MMD_DEVINL WgradItem wgrad_decode(const WgradParams& p, int item) {
    WgradItem w;
    // split-major item order: CTAs that run at the same time work on the SAME token range for different taps / column
    // blocks / n tiles, so the dY and activation tiles they share are served by L2 instead of being re-read from HBM
    const int per_split = p.n_tiles * p.blocks_per_tap * p.n_taps;
    const int split = item / per_split;
    int r = item % per_split;
    const int cb = r % (p.blocks_per_tap * p.n_taps);
    w.n_tile = r / (p.blocks_per_tap * p.n_taps);
    w.bias = (p.db != nullptr) && (cb == 0);
    w.tap = cb / p.blocks_per_tap;
    int b = cb % p.blocks_per_tap;
    int chunk_base = 0;
    w.src = 0;
    w.chunk0 = 0;
    w.ncols = 128;
    for (int s = 0; s < p.n_src; ++s) {
        const int nb = (p.src_chunks[s] + 1) >> 1;
        if (b < nb) {
            w.src = s;
            w.chunk0 = 2 * b;
            w.ncols = (p.src_chunks[s] - 2 * b >= 2) ? 128 : 64;
            break;
        }
        b -= nb;
        chunk_base += p.src_chunks[s];
    }
    w.col0 = (w.tap * p.total_chunks + chunk_base + w.chunk0) * GEMM_BK;
    w.m_begin = static_cast<int>(static_cast<long long>(p.m_tiles) * split / p.splits);
    w.m_end = static_cast<int>(static_cast<long long>(p.m_tiles) * (split + 1) / p.splits);
    return w;
}

MMD_DEVINL void wgrad_tile_origin(const WgradParams& p, int m_idx, int* c /*[5]*/) {
    int r = m_idx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = r % p.ntile[i];
        r /= p.ntile[i];
        c[i + 1] = t * p.box[i];
    }
    c[0] = 0;
}

__global__ void __launch_bounds__(WG_THREADS, 1) conv_wgrad_kernel(const __grid_constant__ WgradParams p, int n_items) {
    extern __shared__ uint8_t wg_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wg_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ones = smem + WG_STAGES * WG_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ones + WG_ONES_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + WG_STAGES;
    uint64_t* tfull_bar = bars + 2 * WG_STAGES;
    uint64_t* tempty_bar = bars + 2 * WG_STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&p.dy_map);
        for (int s = 0; s < p.n_src; ++s) tma_prefetch_desc(&p.a_map[s]);
        for (int i = 0; i < WG_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 4);
        }
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < WG_ONES_BYTES / 16; i += WG_THREADS)
        reinterpret_cast<uint4*>(ones)[i] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async_smem();
    if (warp == 1) tmem_alloc(tmem_slot, 512);   // two accumulators [0,128) [128,256) + their bias columns [256,272) [272,288)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer (uniform warp, elected lane issues; see gemm.cuh) =================
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const WgradItem w = wgrad_decode(p, item);
                const uint32_t bytes = WG_DY_BYTES + w.ncols * (GEMM_BM * 2);
                for (int m = w.m_begin; m < w.m_end; ++m) {
                    int org[5], c[5];
                    wgrad_tile_origin(p, m, org);
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* dst = smem + stage * WG_STAGE_BYTES;
                    if (elect_one()) {
                        mbar_expect_tx(&full_bar[stage], bytes);
                        c[1] = org[1]; c[2] = org[2]; c[3] = org[3]; c[4] = org[4];
                        for (int h = 0; h < 2; ++h) {
                            c[0] = w.n_tile * 128 + h * 64;
                            tma_load_nd(p.rank, dst + h * (GEMM_BM * 128), &p.dy_map, &full_bar[stage], c);
                        }
                        c[1] = org[1] + p.tap[w.tap][0];
                        c[2] = org[2] + p.tap[w.tap][1];
                        c[3] = org[3] + p.tap[w.tap][2];
                        for (int h = 0; h < w.ncols / 64; ++h) {
                            c[0] = (w.chunk0 + h) * GEMM_BK;
                            tma_load_nd(p.rank, dst + WG_DY_BYTES + h * (GEMM_BM * 128), &p.a_map[w.src], &full_bar[stage], c);
                        }
                    }
                    __syncwarp();
                    if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (uniform warp, elected lane issues) =================
        {
            constexpr uint32_t idesc128 = umma_idesc_f16(128, 128, 1, 1);
            constexpr uint32_t idesc64 = umma_idesc_f16(128, 64, 1, 1);
            constexpr uint32_t idesc16 = umma_idesc_f16(128, 16, 1, 1);
            const uint64_t ones_d = umma_desc_sw128(smem_u32(ones), GEMM_BM * 128, 1024);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const WgradItem w = wgrad_decode(p, item);
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 128;
                const uint32_t idesc = (w.ncols == 128) ? idesc128 : idesc64;
                bool first = true;
                for (int m = w.m_begin; m < w.m_end; ++m) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + stage * WG_STAGE_BYTES);
                    // MN-major operands: 16 tokens per K-step = two 8-row groups (SBO 1024 B); the second 64-channel
                    // chunk of the M / N extent sits one [128 x 128 B] unit further (LBO)
                    const uint64_t ad0 = umma_desc_sw128(base, GEMM_BM * 128, 1024);
                    const uint64_t bd0 = umma_desc_sw128(base + WG_DY_BYTES, GEMM_BM * 128, 1024);
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < GEMM_BM / 16; ++ks) {
                            const uint32_t accum = (first && ks == 0) ? 0u : 1u;
                            umma_f16_ss(d_tmem, ad0 + ks * (2048 >> 4), bd0 + ks * (2048 >> 4), idesc, accum);
                            if (w.bias) umma_f16_ss(tmem_base + 256 + acc * 16, ad0 + ks * (2048 >> 4), ones_d, idesc16, accum);
                        }
                        umma_commit(&empty_bar[stage]);
                        if (m == w.m_end - 1) umma_commit(&tfull_bar[acc]);
                    }
                    __syncwarp();
                    first = false;
                    if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ================= epilogue: TMEM -> fp32 gradient =================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const WgradItem w = wgrad_decode(p, item);
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 128;
            const int n_row = w.n_tile * 128 + row;
            const bool empty_range = (w.m_end <= w.m_begin);   // nothing was accumulated: TMEM holds stale data
            float* drow = p.dw + static_cast<size_t>(n_row) * p.ld + w.col0;
            for (int cc = 0; cc < w.ncols / 32; ++cc) {
                uint32_t v[32];
                tmem_ld32(t_addr + cc * 32, v);
                tmem_ld_wait();
                if (n_row < p.n && !empty_range) {
                    if (p.splits == 1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<float4*>(drow + cc * 32 + j * 4) =
                                make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                            __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            atomicAdd(reinterpret_cast<float4*>(drow + cc * 32 + j * 4),
                                      make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                  __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
                    }
                }
            }
            if (w.bias) {
                uint32_t bv[16];
                tmem_ld16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + 256 + acc * 16, bv);
                tmem_ld_wait();
                if (n_row < p.n && !empty_range) atomicAdd(&p.db[n_row], __uint_as_float(bv[0]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        __syncwarp();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------
// Layout helpers around the packed gradient.
// ---------------------------------------------------------------------------
// g[co][ci][t] += scale * dwpk[co * ld + col_off + t * Ci + ci]   (inverse of pack_weight_kernel)
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwpk, float* __restrict__ g, int Co, int Ci, int T,
                                    long long ld, long long col_off, float scale, const float* __restrict__ gscale) {
    if (gscale) scale *= gscale[1];
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(Co) * Ci * T;
    if (idx >= total) return;
    const int t = static_cast<int>(idx % T);
    const long long r = idx / T;
    const int ci = static_cast<int>(r % Ci);
    const int co = static_cast<int>(r / Ci);
    g[idx] += scale * dwpk[co * ld + col_off + static_cast<long long>(t) * Ci + ci];
}

// Transposed pack for the data gradient: dst[c][t * Co + co] = w[co][c_lo + c][t]  (fp32 [Co][Ci][T] -> fp16 [Cs][T*Co]).
// With the taps negated, conv_gemm_kernel over dY with these weights computes dX of source channels [c_lo, c_lo + Cs).
__global__ void pack_weight_t_kernel(const float* __restrict__ w, act_t* __restrict__ dst, int Co, int Ci, int T, int c_lo,
                                     int Cs, long long ld, long long col_off) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(Cs) * T * Co;
    if (idx >= total) return;
    const int co = static_cast<int>(idx % Co);
    const long long r = idx / Co;
    const int t = static_cast<int>(r % T);
    const int c = static_cast<int>(r / T);
    dst[c * ld + col_off + static_cast<long long>(t) * Co + co] =
        __float2half_rn(w[(static_cast<long long>(co) * Ci + c_lo + c) * T + t]);
}

// dst[k][co] = w[co][ci][t] with k = t * Ci + ci: transpose of the stem's packed [Co][64] matrix (rows >= T * Ci stay zero)
__global__ void pack_stem_t_kernel(const float* __restrict__ w, act_t* __restrict__ dst, int Co, int Ci, int T, long long ld) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(Co) * Ci * T) return;
    const int t = static_cast<int>(idx % T);
    const long long r = idx / T;
    const int ci = static_cast<int>(r % Ci);
    const int co = static_cast<int>(r / Ci);
    dst[(static_cast<long long>(t) * Ci + ci) * ld + co] = __float2half_rn(w[idx]);
}

}  // namespace mmd
