// Implicit-GEMM convolution kernel for sm_100a (tcgen05 + TMEM + TMA).
//
// One kernel covers every conv family of MultimodalUNet.forward
// (reference: mm_diffusion/multimodal_unet.py:68-131 VideoConv/AudioConv):
//   out[token, n] = bias[n] + sum_{tap, src, c} A_src[token + delta(tap), c] * W[n, (tap, src, c)]
// on channels-last fp16 activations.  "same" zero padding comes for free from TMA
// out-of-bound zero fill: a tap is just a coordinate offset on the A tensor map.
// Several A sources share one K loop, which gives (a) the channel concat of the
// U-Net skip connections (multimodal_unet.py:1093-1094) without materialising it and
// (b) residual / skip-conv accumulation (multimodal_unet.py:482-483) as extra K
// segments (identity weights for an identity skip).
#pragma once
#include "common.cuh"

namespace mmd {

constexpr int GEMM_BM = 128;       // tokens per tile (UMMA M)
constexpr int GEMM_BK = 64;        // channels per k-iteration (one 128-byte swizzle atom)
constexpr int GEMM_MAX_SRC = 4;
constexpr int GEMM_MAX_TAPS = 27;
constexpr int GEMM_THREADS = 192;  // warp0 TMA, warp1 MMA, warps2-5 epilogue

struct alignas(64) GemmParams {
    CUtensorMap a_map[GEMM_MAX_SRC];  // activation sources, rank `rank`, box (64, box[0..3])
    CUtensorMap b_map;                // packed weights [N, K_total] (K-major), box (64, BN)
    CUtensorMap o_map;                // output, same geometry as A (out_mode 0)
    int n_src;
    int src_chunks[GEMM_MAX_SRC];     // 64-channel chunks per source
    int rank;                         // tensor-map rank of A/O (2..5)
    int box[4];                       // box extents of coordinates 1..4 (product = 128)
    int ntile[4];                     // tiles along coordinates 1..4
    int n_taps;
    int tap[GEMM_MAX_TAPS][3];        // coordinate deltas on coordinates 1..3
    int m_tiles, n_tiles;
    const float* bias;                // [n_tiles*BN] fp32 (padded)
    // out_mode 1: fp32 strided scatter (network heads write NCHW fp32 directly)
    int out_mode;
    float* out_f32;
    long long ostride[4];
    long long ostride_c;
    int dims[4];
    int n_valid;
    // fused GroupNorm statistics of the (fp16-rounded) output: per (domain, group) sum / sum of squares
    // accumulated into double slots [domain][32][2]; domain = base(tile) + row / stats_rows
    double* stats;          // null = off
    int stats_cpg;          // channels per group = N / 32
    int stats_rows;         // rows of one domain inside a tile (64 or 128)
    int stats_mul[4];       // domain base = sum_i origin[i+1] * stats_mul[i] / stats_div
    int stats_div;
    int stats_valid_coord;  // >= 0: rows >= dims[c] - origin[c+1] of the tile are padding (ragged last tile)
};

template <int BN>
struct GemmSmem {
    static constexpr int A_BYTES = GEMM_BM * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // epilogue staging: two buffers of HALF columns each, rotated per half-tile, so the TMA store of one half
    // drains while the next half is being converted (a single buffer serialises store latency with the epilogue)
    static constexpr int HALF = (BN >= 128) ? 128 : BN;          // columns per staging buffer / TMA store group
    static constexpr int NHALF = (BN >= 64) ? BN / HALF : 0;
    static constexpr int OUT_BUF = (BN >= 64) ? (HALF / 64) * GEMM_BM * 128 : 0;
    static constexpr int OUT_BYTES = 2 * OUT_BUF;
    static constexpr int STAGES = (BN >= 256) ? 3 : ((BN >= 128) ? 5 : 8);
    static constexpr int BAR_BYTES = 256 + 2 * 32 * 2 * 4;   // barriers + GroupNorm bins [2 domains][32 groups][2]
    static constexpr int TOTAL = STAGES * STAGE_BYTES + OUT_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : ((2 * BN <= 64) ? 64 : ((2 * BN <= 128) ? 128 : ((2 * BN <= 256) ? 256 : 512)));
};

MMD_DEVINL void gemm_tile_origin(const GemmParams& p, int m_idx, int* c /*[5]*/) {
    int r = m_idx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int t = r % p.ntile[i];
        r /= p.ntile[i];
        c[i + 1] = t * p.box[i];
    }
    c[0] = 0;
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1) conv_gemm_kernel(const __grid_constant__ GemmParams p) {
    using S = GemmSmem<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    uint8_t* out_stage = smem + S::STAGES * S::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + S::OUT_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + S::STAGES;
    uint64_t* tfull_bar = bars + 2 * S::STAGES;
    uint64_t* tempty_bar = bars + 2 * S::STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::STAGES + 4);
    float* gn_bins = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = p.m_tiles * p.n_tiles;
    int total_chunks = 0;
    for (int s = 0; s < p.n_src; ++s) total_chunks += p.src_chunks[s];
    const int num_kb = p.n_taps * total_chunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.n_src; ++s) tma_prefetch_desc(&p.a_map[s]);
        tma_prefetch_desc(&p.b_map);
        if (p.out_mode == 0) tma_prefetch_desc(&p.o_map);
        for (int i = 0; i < S::STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, S::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer (one thread) =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m_idx = tile / p.n_tiles;
                const int n_idx = tile - m_idx * p.n_tiles;
                int org[5];
                gemm_tile_origin(p, m_idx, org);
                int kb = 0;
                for (int t = 0; t < p.n_taps; ++t) {
                    int c[5];
                    c[1] = org[1] + p.tap[t][0];
                    c[2] = org[2] + p.tap[t][1];
                    c[3] = org[3] + p.tap[t][2];
                    c[4] = org[4];
                    for (int s = 0; s < p.n_src; ++s) {
                        for (int ch = 0; ch < p.src_chunks[s]; ++ch, ++kb) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            uint8_t* a_dst = stage_base + stage * S::STAGE_BYTES;
                            uint8_t* b_dst = a_dst + S::A_BYTES;
                            mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
                            c[0] = ch * GEMM_BK;
                            tma_load_nd(p.rank, a_dst, &p.a_map[s], &full_bar[stage], c);
                            tma_load_2d(b_dst, &p.b_map, &full_bar[stage], kb * GEMM_BK, n_idx * BN);
                            if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(GEMM_BM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(stage_base + stage * S::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + S::A_BYTES;
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        const uint64_t ad = umma_desc_sw128(a_addr + k * 32, 16, 1024);
                        const uint64_t bd = umma_desc_sw128(b_addr + k * 32, 16, 1024);
                        umma_f16_ss(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);
            }
        }
    } else {
        // ================= epilogue (4 warps, thread = accumulator row) =================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const bool leader = (threadIdx.x == 64);
        int it = 0;
        uint32_t obuf_sel = 0;   // staging buffer rotation (leader's bulk-group order matches it)
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int m_idx = tile / p.n_tiles;
            const int n_idx = tile - m_idx * p.n_tiles;
            int org[5];
            gemm_tile_origin(p, m_idx, org);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
            const float* bias = p.bias + n_idx * BN;

            if constexpr (BN >= 64) {
                constexpr int HALF = S::HALF;
                if (p.stats != nullptr) gn_bins[threadIdx.x - 64] = 0.f;
                int valid_rows = GEMM_BM;
                if (p.stats != nullptr && p.stats_valid_coord >= 0)
                    valid_rows = min(GEMM_BM, p.dims[p.stats_valid_coord] - org[p.stats_valid_coord + 1]);
#pragma unroll 1
                for (int hf = 0; hf < S::NHALF; ++hf) {
                    uint8_t* obuf = out_stage + (obuf_sel & 1) * S::OUT_BUF;
                    ++obuf_sel;
                    if (leader) tma_store_wait_read1();  // the store issued two halves ago has drained this buffer
                    named_bar_sync(1, 128);
#pragma unroll 1
                    for (int cc = 0; cc < HALF / 32; ++cc) {
                        uint32_t v[32];
                        tmem_ld32(t_addr + hf * HALF + cc * 32, v);
                        tmem_ld_wait();
                        uint8_t* chunk = obuf + (cc >> 1) * (GEMM_BM * 128);
                        const float* bcol = bias + hf * HALF + cc * 32;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bcol + j * 8));
                            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bcol + j * 8 + 4));
                            __half2 h0 = __floats2half2_rn(__uint_as_float(v[j * 8 + 0]) + b0.x, __uint_as_float(v[j * 8 + 1]) + b0.y);
                            __half2 h1 = __floats2half2_rn(__uint_as_float(v[j * 8 + 2]) + b0.z, __uint_as_float(v[j * 8 + 3]) + b0.w);
                            __half2 h2 = __floats2half2_rn(__uint_as_float(v[j * 8 + 4]) + b1.x, __uint_as_float(v[j * 8 + 5]) + b1.y);
                            __half2 h3 = __floats2half2_rn(__uint_as_float(v[j * 8 + 6]) + b1.z, __uint_as_float(v[j * 8 + 7]) + b1.w);
                            uint4 pk;
                            pk.x = *reinterpret_cast<uint32_t*>(&h0);
                            pk.y = *reinterpret_cast<uint32_t*>(&h1);
                            pk.z = *reinterpret_cast<uint32_t*>(&h2);
                            pk.w = *reinterpret_cast<uint32_t*>(&h3);
                            *reinterpret_cast<uint4*>(chunk + sw128_off(row, (cc & 1) * 4 + j)) = pk;
                        }
                    }
                    if (hf == S::NHALF - 1) {   // all accumulator columns have been read: hand the TMEM stage back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(1, 128);
                    if (leader) {
                        int c[5] = {0, org[1], org[2], org[3], org[4]};
#pragma unroll
                        for (int ch = 0; ch < HALF / 64; ++ch) {
                            c[0] = n_idx * BN + hf * HALF + ch * 64;
                            tma_store_nd(p.rank, &p.o_map, obuf + ch * (GEMM_BM * 128), c);
                        }
                        tma_store_commit();
                    }
                    if (p.stats != nullptr) {
                        // column sums of the staged fp16 half-tile: thread = (8-column octet, run of NOCT rows)
                        constexpr int NOCT = HALF / 8;
                        const int et = threadIdx.x - 64;
                        const int oct = et % NOCT;
                        const int r_begin = (et / NOCT) * NOCT;
                        float sm[8], sq[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) { sm[i] = 0.f; sq[i] = 0.f; }
                        const uint8_t* chunk = obuf + (oct >> 3) * (GEMM_BM * 128);
#pragma unroll 4
                        for (int rr = 0; rr < NOCT; ++rr) {
                            const int r = r_begin + rr;
                            if (r < valid_rows) {
                                const uint4 raw = *reinterpret_cast<const uint4*>(chunk + sw128_off(r, oct & 7));
                                const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float2 f = __half22float2(h[i]);
                                    sm[2 * i] += f.x; sq[2 * i] = fmaf(f.x, f.x, sq[2 * i]);
                                    sm[2 * i + 1] += f.y; sq[2 * i + 1] = fmaf(f.y, f.y, sq[2 * i + 1]);
                                }
                            }
                        }
                        const int col0 = n_idx * BN + hf * HALF + oct * 8;
                        const int g0 = (n_idx * BN) / p.stats_cpg;
                        const int dl = r_begin / p.stats_rows;   // 0 or 1: domain inside the tile
                        float* bins = gn_bins + dl * 64;
                        int g = col0 / p.stats_cpg;
                        float as = 0.f, aq = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int gi = (col0 + i) / p.stats_cpg;
                            if (gi != g) {
                                atomicAdd(&bins[2 * (g - g0)], as);
                                atomicAdd(&bins[2 * (g - g0) + 1], aq);
                                g = gi; as = 0.f; aq = 0.f;
                            }
                            as += sm[i]; aq += sq[i];
                        }
                        atomicAdd(&bins[2 * (g - g0)], as);
                        atomicAdd(&bins[2 * (g - g0) + 1], aq);
                    }
                }
                if (p.stats != nullptr) {
                    named_bar_sync(1, 128);
                    const int et = threadIdx.x - 64;
                    const int g0 = (n_idx * BN) / p.stats_cpg;
                    const int dom_base = (org[1] * p.stats_mul[0] + org[2] * p.stats_mul[1] + org[3] * p.stats_mul[2] +
                                          org[4] * p.stats_mul[3]) / p.stats_div;
                    const int bdl = et >> 6, bg = (et & 63) >> 1;
                    const float val = gn_bins[et];
                    if (g0 + bg < 32 && (bdl == 0 || p.stats_rows < GEMM_BM) && val != 0.f)
                        atomicAdd(&p.stats[(static_cast<size_t>(dom_base + bdl) * 32 + (g0 + bg)) * 2 + (et & 1)],
                                  static_cast<double>(val));
                }
            } else {
                // narrow-N head: fp32 scatter, row -> token coordinates via the box decomposition
                uint32_t v[16];
                tmem_ld16(t_addr, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                int r = row;
                long long off = 0;
                bool ok = true;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int li = r % p.box[i];
                    r /= p.box[i];
                    const int ci = org[i + 1] + li;
                    ok = ok && (ci < p.dims[i]);
                    off += static_cast<long long>(ci) * p.ostride[i];
                }
                if (ok) {
                    for (int n = 0; n < p.n_valid; ++n)
                        p.out_f32[off + n * p.ostride_c] = __uint_as_float(v[n]) + __ldg(bias + n);
                }
            }
        }
        if constexpr (BN >= 64) {
            if (leader) tma_store_wait_all0();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, S::TMEM_COLS);
    }
}

}  // namespace mmd
