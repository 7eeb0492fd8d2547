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
//
// Shared-memory budget (227 KB): the mainloop is bound by how many operand bytes are in flight against the
// ~0.7 us L2 latency, not by the tensor pipe, so everything that is not pipeline stages is
// kept small.  Two epilogue shapes (template parameter OC = staged output columns per chunk, double buffered):
//   OC=64  (K-heavy GEMMs, the mainloop hides the epilogue): 2 x 16 KB staging -> BN=128: 6 x 32 KB stages, BN=256: 4 x 48 KB
//   OC=128 (short-K GEMMs, store/epilogue bound: fewer barrier rounds): 2 x 32 KB -> BN=128: 5 stages, BN=256: 3 stages
//   BN=64: 8 x 24 KB, BN=16 (heads, no staging): 8 x 18 KB
#pragma once
#include "common.cuh"

namespace mmd {

constexpr int GEMM_BM = 128;       // tokens per tile (UMMA M)
constexpr int GEMM_BK = 64;        // channels per k-iteration (one 128-byte swizzle atom)
constexpr int GEMM_MAX_SRC = 4;
constexpr int GEMM_MAX_TAPS = 27;
// warp0 TMA, warp1 MMA, then EG epilogue warpgroups of four warps (EG = 2: the two groups take alternate tiles of the
// CTA, one per TMEM accumulator stage, so two tiles drain at the same time — the short-K GEMMs are bound by the
// epilogue, not by the tensor pipe or the loads), then (XF) four warps that transform the A operand (fused GroupNorm apply)
constexpr int gemm_threads(int eg, bool xf) { return 64 + 128 * eg + (xf ? 128 : 0); }
constexpr int GEMM_XF_MAXC = 512;  // widest source the fused GroupNorm apply supports

// Division of a non-negative 31-bit index by a launch constant (tile -> (row block, column block) -> box origin runs
// once per tile in every role; the hardware has no integer divide).  q = (n * mul) >> (31 + shift), exact for n < 2^31.
struct FastDiv {
    uint32_t mul, shift, d;   // d == 1: identity (mul unused)
    __host__ void set(uint32_t div) {
        d = div < 1 ? 1 : div;
        shift = 0;
        while ((1u << shift) < d) ++shift;
        mul = d == 1 ? 0u : static_cast<uint32_t>(((1ull << (31 + shift)) + d - 1) / d);
    }
    __device__ __forceinline__ int div(int n) const {
        return d == 1 ? n : static_cast<int>(__umulhi(static_cast<uint32_t>(n), mul) >> (shift - 1));
    }
};

struct alignas(64) GemmParams {
    CUtensorMap a_map[GEMM_MAX_SRC];  // activation sources, always rank 5 (unit extents past `rank`), box (64, box[0..3])
    CUtensorMap b_map;                // packed weights [N, K_total] (K-major), box (64, BN)
    CUtensorMap o_map;                // output, same geometry as A (out_mode 0)
    CUtensorMap o32_map;              // rank-2 outputs: the same tensor with a (64 columns, 32 rows) box: one store per epilogue warp
    int wstore;                       // 1: every epilogue warp stores its own 32-row band (no cross-warp barrier per chunk)
    int n_src;
    int src_chunks[GEMM_MAX_SRC];     // 64-channel chunks per source
    int rank;                         // tensor-map rank of A/O (2..5)
    int box[4];                       // box extents of coordinates 1..4 (product = 128)
    int ntile[4];                     // tiles along coordinates 1..4
    int n_taps;
    int tap[GEMM_MAX_TAPS][3];        // coordinate deltas on coordinates 1..3
    int m_tiles, n_tiles;             // 128-token tiles, BN-column tiles
    FastDiv fd_ntile[4], fd_ntiles, fd_stats;   // dividers by ntile[i], n_tiles, stats_div
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
    int stats_cpg;          // channels per group = N / 32 (multiple of 4)
    int stats_rows;         // rows of one domain inside a tile (64 or 128)
    int stats_mul[4];       // domain base = sum_i origin[i+1] * stats_mul[i] / stats_div
    int stats_div;
    int stats_valid_coord;  // >= 0: rows >= dims[c] - origin[c+1] of the tile are padding (ragged last tile)
    // optional second output of the same reduction at 4-channel granularity: [domain][stats_q_n = N / 4][2] doubles, so a
    // GroupNorm over a channel CONCAT that contains this tensor can fold its groups from the producers' partial sums
    // instead of re-reading the tensors (groups of (c1 + c2) / 32 channels do not line up with either source's groups)
    double* stats_q;
    int stats_q_n;
    // fused GroupNorm apply on the A operand of source 0 (XF kernels, pointwise GEMMs):
    //   a <- act(gn(a) * (1 + scale) + shift), done in shared memory between the TMA landing and the MMA
    //   (reference: nn.py:22-33 + multimodal_unet.py:459-470 out_layers / :284,664 attention norms)
    const double* xf_sums;  // statistics slots [domains * xf_nsub][32][2]; null = off
    const float* xf_gamma;
    const float* xf_beta;
    const float* xf_film;   // [batch][xf_film_ld]: scale at [0,C), shift at [C,2C); may be null
    int xf_film_ld;
    int xf_dom_per_batch;   // FiLM row = domain / xf_dom_per_batch
    int xf_c, xf_nsub, xf_silu;
    double xf_inv_n;        // 1 / (rows * channels-per-group) the statistics of one domain cover
    int xf_rows;            // rows of one domain inside a tile (64 or 128)
    int xf_mul[4];          // domain base = sum_i origin[i+1] * xf_mul[i] / xf_div
    int xf_div;
    // L2 prefetch distance of the A operand in tiles of this CTA's tile sequence (0 = off): short-K GEMMs are bound by
    // the bytes in flight against DRAM latency, not by the tensor pipe; prefetching the tiles this CTA will load a
    // few iterations from now turns the pipeline's TMA loads into L2 hits.
    int pf_tiles;
    // measurement only (MMD_GEMM_DBG, results are wrong when set): ablation switches that locate the bound of a shape —
    // 1: no global store of the output, 2: no TMEM -> shared conversion, 4: no MMA issue, 8: no A loads, 16: no B loads,
    // 32: no column-sum pass of the fused statistics, 64: no fold + atomics of the fused statistics
    int dbg;
};

// MT = 128-token row blocks per CTA tile.  MT = 2: one weight tile feeds two accumulators (256 tokens x BN per k-block),
// which cuts the L2 -> shared-memory bytes per flop by 25 % (BN 128) / 33 % (BN 256): at 128 x BN tiles the K loop of
// every conv family sits on the L2 delivery rate (~43 B/clk/SM with all SMs loading), not on the tensor pipe.
template <int BN, int OC, bool XF = false, int EG = 1, int MT = 1>
struct GemmSmem {
    static constexpr int A_ONE = GEMM_BM * 128;
    static constexpr int A_BYTES = MT * A_ONE;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // epilogue staging: two buffers of OC columns (OC/64 swizzled 16 KB units each), rotated per chunk, so the TMA
    // store of one chunk drains while the next is converted
    static constexpr int NCHUNK = (BN >= 64) ? BN / OC : 0;
    static constexpr int UNITS = OC / 64;
    static constexpr int OUT_BUF = UNITS * GEMM_BM * 128;
    static constexpr int OUT_BYTES = (BN >= 64) ? EG * 2 * OUT_BUF : 0;   // per epilogue warpgroup
    // barriers (256 B) | GroupNorm partials [BN/64 units][4 bands][16 quads][2] floats
    static constexpr int GN_ONE = (BN >= 64) ? (BN / 64) * 4 * 16 * 2 * 4 : 0;
    static constexpr int GN_BYTES = EG * GN_ONE;
    // XF: per-channel affine table [2 domains][a|b][GEMM_XF_MAXC] floats + group mean / rstd [2][32][2]
    // this tile's bias [BN] floats (epilogue reads it from shared memory: the global loads sat on its critical path)
    static constexpr int BIAS_OFF = 256 + GN_BYTES;
    static constexpr int BIAS_BYTES = (BN >= 64) ? EG * 2 * BN * 4 : 0;   // per group, double buffered by tile parity
    static constexpr int XF_OFF = BIAS_OFF + BIAS_BYTES;
    static constexpr int XF_BYTES = XF ? (2 * 2 * GEMM_XF_MAXC * 4 + 512) : 0;
    static constexpr int BAR_BYTES = 256 + GN_BYTES + BIAS_BYTES + XF_BYTES;
    static constexpr int LIMIT = 232448;   // 227 KB
    static constexpr int FIT = (LIMIT - OUT_BYTES - BAR_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = FIT > 8 ? 8 : FIT;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + OUT_BYTES + BAR_BYTES;   // base is 1024-aligned (checked)
    static_assert(BN < 64 || (OC % 64 == 0 && BN % OC == 0), "staging chunk");
    static_assert(STAGES >= (MT == 2 ? 2 : 3) && TOTAL <= LIMIT, "shared memory budget");
    static_assert(3 * STAGES + 4 <= 30, "barrier block");
    static_assert(MT == 1 || (MT == 2 && EG == 2 && !XF && BN >= 128), "256-token tiles: one epilogue warpgroup per row block");
    // accumulator stages: double buffered unless two 256-column accumulators already fill the 512 TMEM columns
    static constexpr int NACC = (2 * MT * BN <= 512) ? 2 : 1;
    static constexpr int ACC_COLS = NACC * MT * BN;
    static constexpr int TMEM_COLS = (ACC_COLS <= 32) ? 32 : ((ACC_COLS <= 64) ? 64 : ((ACC_COLS <= 128) ? 128 : ((ACC_COLS <= 256) ? 256 : 512)));
};

MMD_DEVINL void gemm_tile_origin(const GemmParams& p, int m_idx, int* c /*[5]*/) {
    int r = m_idx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = p.fd_ntile[i].div(r);
        c[i + 1] = (r - q * p.ntile[i]) * p.box[i];
        r = q;
    }
    c[0] = 0;
}

// Ablation switches (GemmParams::dbg, MMD_GEMM_DBG) exist only in builds with -DMMD_GEMM_ABLATION
// (MMD_NVCC_EXTRA=-DMMD_GEMM_ABLATION MMD_LIB_OUT=... python -m mm_diffusion_b200.build; load with MMD_LIB=...): the
// production library does not carry their loads and branches in the role loops.
#ifdef MMD_GEMM_ABLATION
#define GEMM_DBG(p) ((p).dbg)
#else
#define GEMM_DBG(p) 0
#endif

template <int BN, int OC, bool XF, int EG, int MT>
__global__ void __launch_bounds__(gemm_threads(EG, XF), 1) conv_gemm_kernel(const __grid_constant__ GemmParams p) {
    using S = GemmSmem<BN, OC, XF, EG, MT>;
    static_assert(EG == 1 || EG == 2, "one or two epilogue warpgroups");
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* stage_base = smem;
    uint8_t* out_stage = smem + S::STAGES * S::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + S::OUT_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + S::STAGES;
    uint64_t* tfull_bar = bars + 2 * S::STAGES;
    uint64_t* tempty_bar = bars + 2 * S::STAGES + 2;
    uint64_t* xf_bar = bars + 2 * S::STAGES + 4;   // [STAGES] (XF only): A tile transformed, the MMA may read it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S::STAGES + 4);
    float* gn_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
    float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + S::BIAS_OFF);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = ((p.m_tiles + MT - 1) / MT) * p.n_tiles;   // CTA tiles: MT row blocks x one column block
    // Tile schedule of this CTA.  Plain kernels: strided (tile = blockIdx.x + i * gridDim.x): the N tiles of one row block
    // run at the same time on neighbouring CTAs and share the A tile through L2.  XF kernels: a contiguous range, so a CTA
    // stays inside one GroupNorm domain for many tiles and rebuilds its affine table once or twice instead of per tile.
    const int tile_begin = XF ? static_cast<int>(static_cast<long long>(blockIdx.x) * total_tiles / gridDim.x) : static_cast<int>(blockIdx.x);
    const int tile_end = XF ? static_cast<int>(static_cast<long long>(blockIdx.x + 1) * total_tiles / gridDim.x) : total_tiles;
    const int tile_step = XF ? 1 : static_cast<int>(gridDim.x);
    int total_chunks = 0;
    for (int s = 0; s < p.n_src; ++s) total_chunks += p.src_chunks[s];
    const int num_kb = p.n_taps * total_chunks;

    pdl_trigger();   // the next kernel of the stream may start its own prologue from here on
    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) {   // the 128-byte swizzle atoms need a 1024-byte aligned base
            printf("conv_gemm_kernel: dynamic shared memory base not 1024-byte aligned\n");
            __trap();
        }
        for (int s = 0; s < p.n_src; ++s) tma_prefetch_desc(&p.a_map[s]);
        tma_prefetch_desc(&p.b_map);
        if (p.out_mode == 0) tma_prefetch_desc(&p.o_map);
        for (int i = 0; i < S::STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
            if constexpr (XF) mbar_init(&xf_bar[i], 4);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 4 * MT);   // MT = 2: both epilogue warpgroups read every accumulator stage
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, S::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp != 0) pdl_wait();   // the producer warp waits after it has requested the first weight tiles

    if (warp == 0) {
        // ================= TMA producer =================
        // Same shape as the MMA issuer below: the whole warp walks the loops in uniform control flow, one elected lane
        // issues the expect_tx + TMA instructions (no per-instruction ELECT / BRA.U.ANY loops in the SASS).
        {
            // Weights are constants of the plan, activations come from the previous kernel: the B halves of the first
            // pipeline stages are requested before the grid dependency resolves, the A halves after it.
            int pre = 0;
            if (tile_begin < tile_end) {
                const int n_idx0 = (tile_begin - p.fd_ntiles.div(tile_begin) * p.n_tiles);
                pre = (GEMM_DBG(p) & 24) ? 0 : min(S::STAGES, num_kb);
                // (an incomplete last row-block pair loads one A block only)
                const int nv0 = min(MT, p.m_tiles - p.fd_ntiles.div(tile_begin) * MT);
                if (elect_one()) {
                    for (int kb = 0; kb < pre; ++kb) {
                        mbar_expect_tx(&full_bar[kb], S::B_BYTES + nv0 * S::A_ONE);
                        tma_load_2d(stage_base + kb * S::STAGE_BYTES + S::A_BYTES, &p.b_map, &full_bar[kb], kb * GEMM_BK, n_idx0 * BN);
                    }
                }
                __syncwarp();
            }
            pdl_wait();
            // A tiles of the next pf_tiles-1 tiles of this CTA go to L2 right away; inside the loop every k-block load
            // is paired with the prefetch of the same k-block pf_tiles tiles ahead
            const int pf = p.pf_tiles;
            if (MT == 1 && pf > 1 && lane == 0) {
                for (int d = 1; d < pf; ++d) {
                    const int tile = tile_begin + d * tile_step;
                    if (tile >= tile_end) break;
                    int org[5];
                    gemm_tile_origin(p, p.fd_ntiles.div(tile), org);
                    for (int t = 0; t < p.n_taps; ++t) {
                        int c[5] = {0, org[1] + p.tap[t][0], org[2] + p.tap[t][1], org[3] + p.tap[t][2], org[4]};
                        for (int s = 0; s < p.n_src; ++s)
                            for (int ch = 0; ch < p.src_chunks[s]; ++ch) {
                                c[0] = ch * GEMM_BK;
                                tma_prefetch_nd(5, &p.a_map[s], c);
                            }
                    }
                }
            }
            __syncwarp();
            int stage = 0;
            uint32_t phase = 0;
            int gk = 0;   // k-blocks issued by this CTA (the first `pre` already have their B half in flight)
            for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
                const int mp = p.fd_ntiles.div(tile);
                const int n_idx = tile - mp * p.n_tiles;
                const int m_idx = mp * MT;
                const int nv = min(MT, p.m_tiles - m_idx);   // row blocks of this tile that exist
                int org[5];
                gemm_tile_origin(p, m_idx, org);
                int org1[5] = {0, 0, 0, 0, 0};
                if (MT == 2 && nv == 2) gemm_tile_origin(p, m_idx + 1, org1);
                const int pf_tile = tile + pf * tile_step;
                const bool pf_on = MT == 1 && pf > 0 && pf_tile < tile_end;
                int porg[5] = {0, 0, 0, 0, 0};
                if (pf_on) gemm_tile_origin(p, p.fd_ntiles.div(pf_tile), porg);
                int kb = 0;
                for (int t = 0; t < p.n_taps; ++t) {
                    int c[5];
                    c[1] = org[1] + p.tap[t][0];
                    c[2] = org[2] + p.tap[t][1];
                    c[3] = org[3] + p.tap[t][2];
                    c[4] = org[4];
                    int pc[5] = {0, porg[1] + p.tap[t][0], porg[2] + p.tap[t][1], porg[3] + p.tap[t][2], porg[4]};
                    for (int s = 0; s < p.n_src; ++s) {
                        for (int ch = 0; ch < p.src_chunks[s]; ++ch, ++kb, ++gk) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            uint8_t* a_dst = stage_base + stage * S::STAGE_BYTES;
                            c[0] = ch * GEMM_BK;
                            if (elect_one()) {
                                if (pf_on) {
                                    pc[0] = ch * GEMM_BK;
                                    tma_prefetch_nd(5, &p.a_map[s], pc);
                                }
                                if (GEMM_DBG(p) & 24) {   // ablation: drop the A and / or B loads
                                    const uint32_t bytes = ((GEMM_DBG(p) & 8) ? 0 : nv * S::A_ONE) + ((GEMM_DBG(p) & 16) ? 0 : S::B_BYTES);
                                    if (bytes == 0) { mbar_arrive(&full_bar[stage]); }
                                    else {
                                        mbar_expect_tx(&full_bar[stage], bytes);
                                        if (!(GEMM_DBG(p) & 8)) {
                                            tma_load_5d(a_dst, &p.a_map[s], &full_bar[stage], c[0], c[1], c[2], c[3], c[4]);
                                            if (MT == 2 && nv == 2) {
                                                const int c1[5] = {c[0], org1[1] + p.tap[t][0], org1[2] + p.tap[t][1], org1[3] + p.tap[t][2], org1[4]};
                                                tma_load_5d(a_dst + S::A_ONE, &p.a_map[s], &full_bar[stage], c1[0], c1[1], c1[2], c1[3], c1[4]);
                                            }
                                        }
                                        if (!(GEMM_DBG(p) & 16)) tma_load_2d(a_dst + S::A_BYTES, &p.b_map, &full_bar[stage], kb * GEMM_BK, n_idx * BN);
                                    }
                                } else {
                                    if (gk >= pre) mbar_expect_tx(&full_bar[stage], S::B_BYTES + nv * S::A_ONE);
                                    tma_load_5d(a_dst, &p.a_map[s], &full_bar[stage], c[0], c[1], c[2], c[3], c[4]);
                                    if (MT == 2 && nv == 2) {
                                        const int c1[5] = {c[0], org1[1] + p.tap[t][0], org1[2] + p.tap[t][1], org1[3] + p.tap[t][2], org1[4]};
                                        tma_load_5d(a_dst + S::A_ONE, &p.a_map[s], &full_bar[stage], c1[0], c1[1], c1[2], c1[3], c1[4]);
                                    }
                                    if (gk >= pre) tma_load_2d(a_dst + S::A_BYTES, &p.b_map, &full_bar[stage], kb * GEMM_BK, n_idx * BN);
                                }
                            }
                            __syncwarp();
                            if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp walks the tile / k-block loops in uniform control flow (all lanes wait on the barriers); one
        // elected lane issues the tcgen05 instructions.  Issuing from a lane-guarded branch (`if (lane == 0)`) makes
        // ptxas wrap every UTCHMMA / UTCBAR in an ELECT + BRA.U.ANY loop over the active lanes with R2UR moves of each
        // operand: ~75 dependent instructions per k-block on one warp, longer than the 256 cycles the tensor pipe needs
        // for a 128 x 128 x 64 block (measured: the k-block loop alone, no loads / MMAs / epilogue, cost ~500 cycles).
        {
            constexpr uint32_t idesc = umma_idesc_f16(GEMM_BM, BN, 0, 0);
            const uint64_t a_desc0 = umma_desc_sw128(smem_u32(stage_base), 16, 1024);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it) {
                const int acc = (S::NACC == 2) ? (it & 1) : 0;
                const uint32_t acc_phase = (S::NACC == 2) ? ((it >> 1) & 1) : (it & 1);
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (MT * BN);
                const bool two = MT == 2 && p.fd_ntiles.div(tile) * MT + 1 < p.m_tiles;   // second row block exists
                for (int kb = 0; kb < num_kb; ++kb) {
                    // XF kernels: every k-block is handed over by the transform warps (they wait for the TMA, transform the
                    // blocks of source 0 and pass the others through), so all roles follow the ring in lock-step
                    mbar_wait(XF ? &xf_bar[stage] : &full_bar[stage], phase);
                    tc_fence_after();
                    // descriptors: stage 0's plus the stage offset in the 16-byte-unit start-address field (shared memory
                    // is < 256 KB, so the 14-bit field cannot carry); the 16-element k-steps bump it by +32 bytes = +2
                    const uint64_t ad0 = a_desc0 + static_cast<uint64_t>(stage * (S::STAGE_BYTES >> 4));
                    const uint64_t bd0 = ad0 + (S::A_BYTES >> 4);
                    if (elect_one()) {
                        if (!(GEMM_DBG(p) & 4)) {
#pragma unroll
                            for (int k = 0; k < GEMM_BK / 16; ++k)
                                umma_f16_ss(d_tmem, ad0 + 2 * k, bd0 + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                            if (two) {   // second row block of the tile: same weight tile, its own accumulator columns
                                constexpr uint64_t a1 = S::A_ONE >> 4;
#pragma unroll
                                for (int k = 0; k < GEMM_BK / 16; ++k)
                                    umma_f16_ss(d_tmem + BN, ad0 + a1 + 2 * k, bd0 + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                            }
                        }
                        umma_commit(&empty_bar[stage]);
                        if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
                    }
                    __syncwarp();
                    if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (XF && warp >= 2 + 4 * EG) {
        // ================= A-operand transform (4 warps): fused GroupNorm apply (+FiLM, +SiLU) on source 0 =================
        // thread = (16-byte column unit `oct` of the 128-byte row, group of 8 rows `rg`): a warp touches four complete
        // rows per access (conflict-free in the 128-byte swizzle), and every thread needs just 8 channels of the affine
        // table per k-block.  y = act(x * a[c] + b[c]) in fp32, stored back in place; SiLU = h + h tanh(h), h = y / 2
        // (one MUFU op per element; the 1/2 is folded into the table).
        if (p.xf_sums != nullptr) {
            float* xf_coef = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + S::XF_OFF);   // [dl][a|b][c]
            float* xf_gstat = xf_coef + 2 * 2 * GEMM_XF_MAXC;                                          // [dl][32][mean|rstd]
            const int tt = threadIdx.x - (64 + 128 * EG);   // 0..127
            const int oct = tt & 7, rg = tt >> 3;
            const int C = p.xf_c;
            const int cpg = C / 32;
            const int ndom = (p.xf_rows < GEMM_BM) ? 2 : 1;
            const int dl = (rg * 8) / p.xf_rows;
            const int xf_chunks = p.src_chunks[0];
            const bool do_silu = p.xf_silu != 0;
            const float pre = do_silu ? 0.5f : 1.0f;
            int stage = 0;
            uint32_t phase = 0;
            int cached_dom = -1;
            for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
                const int m_idx = p.fd_ntiles.div(tile);
                int org[5];
                gemm_tile_origin(p, m_idx, org);
                const int dom_base = (org[1] * p.xf_mul[0] + org[2] * p.xf_mul[1] + org[3] * p.xf_mul[2] +
                                      org[4] * p.xf_mul[3]) / p.xf_div;
                if (dom_base != cached_dom) {   // (re)build the per-channel affine table of this tile's domain(s)
                    named_bar_sync(2, 128);
                    if (tt < 32 * ndom) {
                        const int d = tt >> 5, g = tt & 31;
                        double su = 0.0, sq = 0.0;
                        for (int k = 0; k < p.xf_nsub; ++k) {
                            const double* sl = p.xf_sums + (static_cast<size_t>(dom_base + d) * p.xf_nsub + k) * 64;
                            su += sl[2 * g];
                            sq += sl[2 * g + 1];
                        }
                        const double mean = su * p.xf_inv_n;
                        double var = sq * p.xf_inv_n - mean * mean;
                        if (var < 0) var = 0;
                        xf_gstat[(d * 32 + g) * 2] = static_cast<float>(mean);
                        xf_gstat[(d * 32 + g) * 2 + 1] = rsqrtf(static_cast<float>(var) + 1e-5f);
                    }
                    named_bar_sync(2, 128);
                    for (int i = tt; i < ndom * C; i += 128) {
                        const int d = i / C, c = i - d * C;
                        const int g = c / cpg;
                        float a = xf_gstat[(d * 32 + g) * 2 + 1] * __ldg(p.xf_gamma + c);
                        float b = __ldg(p.xf_beta + c) - xf_gstat[(d * 32 + g) * 2] * a;
                        if (p.xf_film != nullptr) {
                            const float* fb = p.xf_film + static_cast<size_t>((dom_base + d) / p.xf_dom_per_batch) * p.xf_film_ld;
                            const float sc = 1.f + __ldg(fb + c);
                            a *= sc;
                            b = b * sc + __ldg(fb + C + c);
                        }
                        xf_coef[(d * 2 + 0) * GEMM_XF_MAXC + c] = a * pre;
                        xf_coef[(d * 2 + 1) * GEMM_XF_MAXC + c] = b * pre;
                    }
                    named_bar_sync(2, 128);
                    cached_dom = dom_base;
                }
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (kb < xf_chunks) {
                        const float* ca = xf_coef + (dl * 2 + 0) * GEMM_XF_MAXC + kb * GEMM_BK + oct * 8;
                        const float* cb = xf_coef + (dl * 2 + 1) * GEMM_XF_MAXC + kb * GEMM_BK + oct * 8;
                        const float4 a0 = *reinterpret_cast<const float4*>(ca), a1 = *reinterpret_cast<const float4*>(ca + 4);
                        const float4 b0 = *reinterpret_cast<const float4*>(cb), b1 = *reinterpret_cast<const float4*>(cb + 4);
                        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        mbar_wait(&full_bar[stage], phase);
                        uint8_t* a_tile = stage_base + stage * S::STAGE_BYTES;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            uint4* ptr = reinterpret_cast<uint4*>(a_tile + sw128_off(rg * 8 + j, oct));
                            uint4 raw = *ptr;
                            __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float2 f = __half22float2(h[k]);
                                float u0 = fmaf(f.x, av[2 * k], bv[2 * k]);
                                float u1 = fmaf(f.y, av[2 * k + 1], bv[2 * k + 1]);
                                if (do_silu) {
                                    float t0, t1;
                                    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
                                    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
                                    u0 = fmaf(u0, t0, u0);
                                    u1 = fmaf(u1, t1, u1);
                                }
                                h[k] = __floats2half2_rn(u0, u1);
                            }
                            *ptr = raw;
                        }
                        fence_proxy_async_smem();
                    } else {
                        // untransformed block (skip / residual segment): passed through.  A parity wait is only meaningful
                        // for a waiter that follows every phase of a stage — skipping blocks here either lets the warp run
                        // ahead of a fill that has not started or fall two phases behind one that has — so every block is
                        // waited for and handed on.
                        mbar_wait(&full_bar[stage], phase);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&xf_bar[stage]);
                    if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ================= epilogue (4 warps, thread = accumulator row) =================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int eg = (warp - 2) >> 2;            // epilogue warpgroup: with EG = 2 group g drains accumulator stage g
        const int et = (threadIdx.x - 64) & 127;   // 0..127 inside the group
        const bool lead_warp = (et >> 5) == 0;     // its elected lane issues the TMA stores of the group
        const int ebar = 1 + 2 * eg;               // named barrier of this group (2 belongs to the transform warps)
        uint8_t* const out_stage_g = out_stage + eg * 2 * S::OUT_BUF;
        float* const gn_part_g = gn_part + eg * (S::GN_ONE / 4);
        float* const bias_g0 = bias_s + eg * 2 * BN;
        int tcount = 0;   // tiles this group has processed (bias buffer parity)
        int it = 0;
        uint32_t obuf_sel = 0;   // staging buffer rotation (leader's bulk-group order matches it)
        for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it) {
            // EG = 2, MT = 1: the groups take alternate tiles; MT = 2: group g drains row block g of every tile
            if (MT == 1 && EG == 2 && (it & 1) != eg) continue;
            const int acc = (S::NACC == 2) ? (it & 1) : 0;
            const uint32_t acc_phase = (S::NACC == 2) ? ((it >> 1) & 1) : (it & 1);
            const int mp = p.fd_ntiles.div(tile);
            const int n_idx = tile - mp * p.n_tiles;
            const int m_idx = mp * MT + (MT == 2 ? eg : 0);
            if (MT == 2 && m_idx >= p.m_tiles) {   // missing second row block of the last pair: keep the barrier protocol going
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                continue;
            }
            int org[5];
            gemm_tile_origin(p, m_idx, org);
            const float* bias = p.bias + n_idx * BN;
            // the bias row is double buffered by tile parity: with per-warp stores a fast warp may write the next tile's row
            // while a slow one still converts this tile (the tile's one barrier keeps them within a tile of each other)
            float* const bias_g = bias_g0 + (tcount & 1) * BN;
            ++tcount;
            // this tile's bias: requested before the wait for the accumulator so its latency hides behind the mainloop
            float bias_pre[(BN + 127) / 128];
            if constexpr (BN >= 64) {
#pragma unroll
                for (int i = 0; i < (BN + 127) / 128; ++i) bias_pre[i] = (et + i * 128 < BN) ? __ldg(bias + et + i * 128) : 0.f;
            }
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * (MT * BN) + (MT == 2 ? eg * BN : 0);

            if constexpr (BN >= 64) {
                int valid_rows = GEMM_BM;
                if (p.stats != nullptr && p.stats_valid_coord >= 0)
                    valid_rows = min(GEMM_BM, p.dims[p.stats_valid_coord] - org[p.stats_valid_coord + 1]);
#pragma unroll 1
                for (int cc = 0; cc < S::NCHUNK; ++cc) {
                    uint8_t* obuf = out_stage_g + (obuf_sel & 1) * S::OUT_BUF;
                    ++obuf_sel;
                    // wstore (token-matrix outputs): each warp owns the 32-row band of the staging buffers it converts and stores,
                    // so a chunk needs no cross-warp barrier at all (one per tile remains, for the shared bias row)
                    const bool ws = p.wstore != 0;
                    if (ws || lead_warp) {
                        if (elect_one()) tma_store_wait_read1();  // the store issued two chunks ago has drained this buffer
                        __syncwarp();
                    }
                    if (cc == 0) {   // (readers of the previous tile's bias are past that tile's last barrier)
#pragma unroll
                        for (int i = 0; i < (BN + 127) / 128; ++i)
                            if (et + i * 128 < BN) bias_g[et + i * 128] = bias_pre[i];
                    }
                    if (!ws || cc == 0) named_bar_sync(ebar, 128);
                    // accumulator -> fp16 staging, 32 columns at a time with the next TMEM load already in flight
                    uint32_t va[32], vb[32];
                    if (!(GEMM_DBG(p) & 2)) tmem_ld32(t_addr + cc * OC, va);
#pragma unroll
                    for (int l = 0; l < OC / 32; ++l) {
                        if (GEMM_DBG(p) & 2) break;
                        uint32_t* v = (l & 1) ? vb : va;
                        tmem_ld_wait();
                        if (l + 1 < OC / 32) tmem_ld32(t_addr + cc * OC + (l + 1) * 32, (l & 1) ? va : vb);
                        uint8_t* unit_base = obuf + (l >> 1) * (GEMM_BM * 128);
                        const float* bcol = bias_g + cc * OC + l * 32;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 b0 = *reinterpret_cast<const float4*>(bcol + q * 8);
                            const float4 b1 = *reinterpret_cast<const float4*>(bcol + q * 8 + 4);
                            __half2 h0 = __floats2half2_rn(__uint_as_float(v[q * 8 + 0]) + b0.x, __uint_as_float(v[q * 8 + 1]) + b0.y);
                            __half2 h1 = __floats2half2_rn(__uint_as_float(v[q * 8 + 2]) + b0.z, __uint_as_float(v[q * 8 + 3]) + b0.w);
                            __half2 h2 = __floats2half2_rn(__uint_as_float(v[q * 8 + 4]) + b1.x, __uint_as_float(v[q * 8 + 5]) + b1.y);
                            __half2 h3 = __floats2half2_rn(__uint_as_float(v[q * 8 + 6]) + b1.z, __uint_as_float(v[q * 8 + 7]) + b1.w);
                            uint4 pk;
                            pk.x = *reinterpret_cast<uint32_t*>(&h0);
                            pk.y = *reinterpret_cast<uint32_t*>(&h1);
                            pk.z = *reinterpret_cast<uint32_t*>(&h2);
                            pk.w = *reinterpret_cast<uint32_t*>(&h3);
                            *reinterpret_cast<uint4*>(unit_base + sw128_off(row, (l & 1) * 4 + q)) = pk;
                        }
                    }
                    if (cc == S::NCHUNK - 1) {   // all accumulator columns have been read: hand the TMEM stage back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                    }
                    fence_proxy_async_smem();
                    if (ws) {
                        __syncwarp();
                        if (!(GEMM_DBG(p) & 1)) {
                            if (elect_one()) {
                                const int band_rows = quad * 32;   // the rows this warp converted: its TMEM lane quarter
#pragma unroll
                                for (int u = 0; u < S::UNITS; ++u)
                                    tma_store_5d(&p.o32_map, obuf + u * (GEMM_BM * 128) + band_rows * 128, n_idx * BN + cc * OC + u * 64,
                                                 org[1] + band_rows, 0, 0, 0);
                                tma_store_commit();
                            }
                            __syncwarp();
                        }
                    } else {
                    named_bar_sync(ebar, 128);
                    if (lead_warp && !(GEMM_DBG(p) & 1)) {
                        if (elect_one()) {
                            int c[5] = {0, org[1], org[2], org[3], org[4]};
#pragma unroll
                            for (int u = 0; u < S::UNITS; ++u) {
                                c[0] = n_idx * BN + cc * OC + u * 64;
                                tma_store_5d(&p.o_map, obuf + u * (GEMM_BM * 128), c[0], c[1], c[2], c[3], c[4]);
                            }
                            tma_store_commit();
                        }
                        __syncwarp();
                    }
                    }
                    if (p.stats != nullptr && !(GEMM_DBG(p) & 32)) {
                        // Column sums of the staged fp16 chunk without atomics, one 64-column unit at a time: lane & 15 =
                        // 4-column quad (8 bytes of a 128-byte row), the two half-warps take 16 rows each of the warp's
                        // 32-row band; per (band, quad) partials are folded into groups by the write-out pass below.
                        const int band = quad, quad4 = lane & 15, half = lane >> 4;   // the warp's own rows (no cross-warp read)
                        const int unit = quad4 >> 1, sub = (quad4 & 1) * 8;
#pragma unroll 1
                        for (int u = 0; u < S::UNITS; ++u) {
                            const uint8_t* ub = obuf + u * (GEMM_BM * 128);
                            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
                            // all 16 loads first, then the arithmetic: a per-row `if (r < valid_rows)` makes every row its own
                            // reconvergence block with the shared-memory latency exposed (measured 13.6 us of a 39.5 us
                            // launch); rows past a ragged edge are masked by value instead (the staging rows exist either way)
                            uint2 raw[16];
#pragma unroll
                            for (int rr = 0; rr < 16; ++rr)
                                raw[rr] = *reinterpret_cast<const uint2*>(ub + sw128_off(band * 32 + half * 16 + rr, unit) + sub);
#pragma unroll
                            for (int rr = 0; rr < 16; ++rr) {
                                const bool ok = band * 32 + half * 16 + rr < valid_rows;
                                const uint32_t lo = ok ? raw[rr].x : 0u, hi = ok ? raw[rr].y : 0u;
                                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&lo));
                                const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi));
                                s0 += a.x; q0 = fmaf(a.x, a.x, q0);
                                s1 += a.y; q1 = fmaf(a.y, a.y, q1);
                                s2 += b.x; q2 = fmaf(b.x, b.x, q2);
                                s3 += b.y; q3 = fmaf(b.y, b.y, q3);
                            }
                            float su = (s0 + s1) + (s2 + s3), sq = (q0 + q1) + (q2 + q3);
                            su += __shfl_xor_sync(0xffffffffu, su, 16);
                            sq += __shfl_xor_sync(0xffffffffu, sq, 16);
                            if (half == 0) {   // gn_part[64-column unit of the tile][band][quad][2]
                                float* part = gn_part_g + (((cc * S::UNITS + u) * 4 + band) * 16 + quad4) * 2;
                                part[0] = su;
                                part[1] = sq;
                            }
                        }
                    }
                }
                if (p.stats != nullptr && !(GEMM_DBG(p) & 64)) {
                    named_bar_sync(ebar, 128);
                    // thread = (domain-in-tile, local group, statistic): fold bands x quads of that group
                    const int cpg = p.stats_cpg;                 // multiple of 4
                    const int qpg = cpg >> 2;                    // quads per group
                    const int ndom = (p.stats_rows < GEMM_BM) ? 2 : 1;
                    const int bands_per_dom = 4 / ndom;
                    const int col_base = n_idx * BN;
                    const int g_first = col_base / cpg;
                    const int groups_tile = (col_base + BN - 1) / cpg - g_first + 1;   // groups intersecting this tile
                    const int dom_base = p.fd_stats.div(org[1] * p.stats_mul[0] + org[2] * p.stats_mul[1] + org[3] * p.stats_mul[2] +
                                                        org[4] * p.stats_mul[3]);
                    for (int item = et; item < ndom * groups_tile * 2; item += 128) {
                        const int st = item & 1;
                        const int gl = (item >> 1) % groups_tile;
                        const int dl = (item >> 1) / groups_tile;
                        const int g = g_first + gl;
                        if (g >= 32) continue;
                        // quads of group g inside this tile: global quad index = column / 4
                        const int q_lo = max(g * qpg, col_base >> 2), q_hi = min((g + 1) * qpg, (col_base + BN) >> 2);
                        float a = 0.f;
                        for (int q = q_lo; q < q_hi; ++q) {
                            const int ql = q - (col_base >> 2);           // 0 .. BN/4-1
                            const int uq = ql >> 4, qq = ql & 15;
                            for (int b = 0; b < bands_per_dom; ++b)
                                a += gn_part_g[((uq * 4 + dl * bands_per_dom + b) * 16 + qq) * 2 + st];
                        }
                        if (q_hi > q_lo)
                            atomicAdd(&p.stats[(static_cast<size_t>(dom_base + dl) * 32 + g) * 2 + st], static_cast<double>(a));
                    }
                    if (p.stats_q != nullptr) {
                        constexpr int QT = BN / 4;   // quads of this tile (power of two)
                        for (int item = et; item < ndom * QT * 2; item += 128) {
                            const int st = item & 1;
                            const int ql = (item >> 1) & (QT - 1);
                            const int dl = (item >> 1) / QT;
                            const int uq = ql >> 4, qq = ql & 15;
                            float a = 0.f;
                            for (int b = 0; b < bands_per_dom; ++b)
                                a += gn_part_g[((uq * 4 + dl * bands_per_dom + b) * 16 + qq) * 2 + st];
                            const int qg = (col_base >> 2) + ql;
                            if (qg < p.stats_q_n)
                                atomicAdd(&p.stats_q[(static_cast<size_t>(dom_base + dl) * p.stats_q_n + qg) * 2 + st], static_cast<double>(a));
                        }
                    }
                    // (the next tile's named barriers order these reads before gn_part is rewritten)
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
            if (lead_warp || p.wstore != 0) {
                if (elect_one()) tma_store_wait_all0();
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        __syncwarp();
        tmem_dealloc(tmem_base, S::TMEM_COLS);
    }
}

}  // namespace mmd
