// Flash-style attention core on tcgen05/TMEM for sm_100a.
//
// Covers (reference: mm_diffusion/multimodal_unet.py)
//   * SingleModalQKVAttention.forward :221-240  (spatial video self-attention, audio self-attention)
//   * QKVAttention.forward            :507-564  with CrossAttentionBlock.attention_index :614-647
//     (Random-Shift cross-modal attention: query block i attends the key blocks
//      (i + shift + j) mod n_blocks, j < win)
// softmax(q k^T / sqrt(d)) v with fp32 logits/softmax, fp16 operands.
//
// One CTA = one 128-row query tile of one (sample, block, head).  Warp roles:
//   warps 0-3  softmax + output accumulation (thread = query row)
//   warp  4    TMA producer (Q once, K/V tiles double buffered)
//   warp  5    tcgen05.mma issuer: S = Q K^T into TMEM (double buffered), O_t = P V into TMEM
// The key window of the cross-modal case is a contiguous token range modulo the
// per-sample key count, i.e. at most two contiguous segments, walked in 128-key tiles.
#pragma once
#include "common.cuh"

namespace mmd {

constexpr int ATT_THREADS = 192;
constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 128;

struct alignas(64) AttnParams {
    CUtensorMap q_map, k_map, v_map;  // 2-D views [rows][cols], box (64 cols, 128 rows), SWIZZLE_128B
    act_t* out;                       // [q rows][out_ld]; head h writes columns [h*d, (h+1)*d)
    int out_ld;
    int B, heads;
    int q_col0, k_col0, v_col0;       // column of head 0 inside each view
    int n_blocks;                     // query blocks (frames / audio segments) per sample
    int q_blk, q_per_batch;           // query rows per block / per sample
    int k_blk, k_per_batch;           // key rows per block / per sample (wrap modulus)
    int win;                          // window, in blocks
    const int* shift_ptr;             // device scalar (random window shift), may be null
    int q_tiles;                      // ceil(q_blk / 128)
    float scale_log2;                 // d^-1/2 * log2(e)
    // training only: log2-sum-exp of the scaled logits per (head, query row), lse[head * lse_ld + row]; the backward
    // kernels rebuild P = exp2(s * scale_log2 - lse) from it (null = not written)
    float* lse;
    long long lse_ld;
    // -DMMD_ATTN_TRACE builds only: clock64 stamps of the first CTAs' roles per key tile (tools/gpu_attn_trace.py)
    long long* trace;
};

// Timeline instrumentation of attention64_kernel (compiled out unless -DMMD_ATTN_TRACE): role 0 = softmax warp 0,
// 1 = MMA warp, 2 = TMA warp; up to 8 events per (CTA, role, tile); CTAs 0..7, tiles 0..31.
#ifdef MMD_ATTN_TRACE
#define ATT_TRACE(p, role, g, ev)                                                                              \
    do {                                                                                                       \
        if ((p).trace != nullptr && blockIdx.x < 8 && (g) < 32 && (threadIdx.x & 31) == 0)                       \
            (p).trace[((static_cast<int>(blockIdx.x) * 3 + (role)) * 32 + (g)) * 8 + (ev)] = clock64();          \
    } while (0)
#else
#define ATT_TRACE(p, role, g, ev) do { } while (0)
#endif

struct AttnWork {
    int q_row0, q_valid;
    int seg_row[2], seg_len[2];  // absolute key row start and length of the <=2 contiguous key segments
    int n_tiles;
    int head;
};

MMD_DEVINL AttnWork attn_decode(const AttnParams& p, int idx) {
    AttnWork w;
    const int qt = idx % p.q_tiles; idx /= p.q_tiles;
    w.head = idx % p.heads; idx /= p.heads;
    const int blk = idx % p.n_blocks;
    const int b = idx / p.n_blocks;
    w.q_row0 = b * p.q_per_batch + blk * p.q_blk + qt * ATT_BQ;
    w.q_valid = min(ATT_BQ, p.q_blk - qt * ATT_BQ);
    const int shift = p.shift_ptr ? *p.shift_ptr : 0;
    const int start = ((blk + shift) % p.n_blocks) * p.k_blk;
    const int len = p.win * p.k_blk;
    const int len0 = min(len, p.k_per_batch - start);
    w.seg_row[0] = b * p.k_per_batch + start;
    w.seg_len[0] = len0;
    w.seg_row[1] = b * p.k_per_batch;
    w.seg_len[1] = len - len0;
    w.n_tiles = (len0 + ATT_BKV - 1) / ATT_BKV + (len - len0 + ATT_BKV - 1) / ATT_BKV;
    return w;
}
// tile t -> (absolute key row, valid keys)
MMD_DEVINL void attn_tile(const AttnWork& w, int t, int& row, int& valid) {
    const int t0 = (w.seg_len[0] + ATT_BKV - 1) / ATT_BKV;
    if (t < t0) {
        row = w.seg_row[0] + t * ATT_BKV;
        valid = min(ATT_BKV, w.seg_len[0] - t * ATT_BKV);
    } else {
        const int u = t - t0;
        row = w.seg_row[1] + u * ATT_BKV;
        valid = min(ATT_BKV, w.seg_len[1] - u * ATT_BKV);
    }
}

// D = 64 runs two CTAs per SM (one S buffer, one V stage, 256 TMEM columns): the second CTA's MMAs fill the
// tensor pipe while the first is in its softmax.  D = 96 / 128 keep the deeper single-CTA configuration.
template <int D>
struct AttnCfg {
    static constexpr int SBUF = (D == 64) ? 1 : 2;      // S accumulators in TMEM
    static constexpr int VST = (D == 64) ? 1 : 2;       // V stages in shared memory
    static constexpr int TMEM_COLS = (D == 64) ? 256 : 512;
    static constexpr int O_COL = SBUF * 128;            // TMEM column of the O tile
};

template <int D>
struct AttnSmem {
    static constexpr int NCH = (D + 63) / 64;          // 64-column boxes per operand tile
    static constexpr int TILE = NCH * ATT_BQ * 128;    // bytes of one Q/K/V tile
    static constexpr int Q_OFF = 0;
    static constexpr int K_OFF = TILE;
    static constexpr int V_OFF = K_OFF + 2 * TILE;
    static constexpr int P_OFF = V_OFF + AttnCfg<D>::VST * TILE;
    static constexpr int P_BYTES = 2 * ATT_BQ * 128;
    static constexpr int BAR_OFF = P_OFF + P_BYTES;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

MMD_DEVINL float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int D>
__global__ void __launch_bounds__(ATT_THREADS, (D == 64) ? 2 : 1) attention_kernel(const __grid_constant__ AttnParams p) {
    using S = AttnSmem<D>;
    using Cfg = AttnCfg<D>;
    constexpr int SBUF = Cfg::SBUF;
    constexpr int VST = Cfg::VST;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* q_full = bars;        // 1
    uint64_t* k_full = bars + 1;    // 2
    uint64_t* k_empty = bars + 3;   // 2
    uint64_t* v_full = bars + 5;    // 2
    uint64_t* v_empty = bars + 7;   // 2
    uint64_t* s_full = bars + 9;    // 2
    uint64_t* p_ready = bars + 11;  // 1 (128 arrivals)
    uint64_t* o_full = bars + 12;   // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const AttnWork w = attn_decode(p, blockIdx.x);
    const int T = w.n_tiles;

    pdl_trigger();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&p.q_map);
        tma_prefetch_desc(&p.k_map);
        tma_prefetch_desc(&p.v_map);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1);
        }
        mbar_init(p_ready, 128);
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // q/k/v come from the previous kernel
    const uint32_t tmem_S = tmem_base;                // SBUF 128-column buffers
    const uint32_t tmem_O = tmem_base + Cfg::O_COL;   // D columns

    if (warp == 4) {
        // ===================== TMA producer (uniform warp, elected lane issues; see gemm.cuh) =====================
        {
            if (elect_one()) {
                mbar_expect_tx(q_full, S::TILE);
                for (int ch = 0; ch < S::NCH; ++ch)
                    tma_load_2d(smem + S::Q_OFF + ch * (ATT_BQ * 128), &p.q_map, q_full, p.q_col0 + w.head * D + ch * 64,
                                w.q_row0);
            }
            __syncwarp();
            for (int t = 0; t < T; ++t) {
                const int st = t & 1;
                const uint32_t ph = (t >> 1) & 1;
                int krow, kvalid;
                attn_tile(w, t, krow, kvalid);
                mbar_wait(&k_empty[st], ph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&k_full[st], S::TILE);
                    for (int ch = 0; ch < S::NCH; ++ch)
                        tma_load_2d(smem + S::K_OFF + st * S::TILE + ch * (ATT_BKV * 128), &p.k_map, &k_full[st],
                                    p.k_col0 + w.head * D + ch * 64, krow);
                }
                __syncwarp();
                const int vs = (VST == 2) ? st : 0;
                const uint32_t vph = (VST == 2) ? ph : static_cast<uint32_t>(t & 1);
                mbar_wait(&v_empty[vs], vph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&v_full[vs], S::TILE);
                    for (int ch = 0; ch < S::NCH; ++ch)
                        tma_load_2d(smem + S::V_OFF + vs * S::TILE + ch * (ATT_BKV * 128), &p.v_map, &v_full[vs],
                                    p.v_col0 + w.head * D + ch * 64, krow);
                }
                __syncwarp();
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer (uniform warp, elected lane issues) =====================
        {
            constexpr uint32_t idesc_qk = umma_idesc_f16(ATT_BQ, ATT_BKV, 0, 0);
            constexpr uint32_t idesc_pv = umma_idesc_f16(ATT_BQ, D, 0, 1);  // B (=V) is MN-major
            const uint32_t q_addr = smem_u32(smem + S::Q_OFF);
            const uint32_t p_addr = smem_u32(smem + S::P_OFF);
            auto issue_qk = [&](int t) {
                const int st = t & 1;
                const uint32_t ph = (t >> 1) & 1;
                mbar_wait(&k_full[st], ph);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(smem + S::K_OFF + st * S::TILE);
                const int sb = (SBUF == 2) ? st : 0;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks) {
                        const uint32_t off = (ks >> 2) * (ATT_BQ * 128) + (ks & 3) * 32;
                        umma_f16_ss(tmem_S + sb * 128, umma_desc_sw128(q_addr + off, 16, 1024),
                                    umma_desc_sw128(k_addr + off, 16, 1024), idesc_qk, ks != 0 ? 1u : 0u);
                    }
                    umma_commit(&k_empty[st]);
                    umma_commit(&s_full[sb]);
                }
                __syncwarp();
            };
            mbar_wait(q_full, 0);
            tc_fence_after();
            issue_qk(0);
            for (int t = 0; t < T; ++t) {
                if (SBUF == 2 && t + 1 < T) issue_qk(t + 1);   // look-ahead into the other S buffer
                const int st = t & 1;
                const uint32_t ph = (t >> 1) & 1;
                const int vs = (VST == 2) ? st : 0;
                const uint32_t vph = (VST == 2) ? ph : static_cast<uint32_t>(t & 1);
                int krow, kvalid;
                attn_tile(w, t, krow, kvalid);
                mbar_wait(p_ready, t & 1);       // softmax t is done with S (and P is written)
                mbar_wait(&v_full[vs], vph);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(smem + S::V_OFF + vs * S::TILE);
                const int nks = (kvalid + 15) >> 4;
                if (elect_one()) {
                    for (int ks = 0; ks < nks; ++ks) {
                        const uint32_t poff = (ks >> 2) * (ATT_BQ * 128) + (ks & 3) * 32;
                        umma_f16_ss(tmem_O, umma_desc_sw128(p_addr + poff, 16, 1024),
                                    umma_desc_sw128(v_addr + ks * 2048, ATT_BKV * 128, 1024), idesc_pv, ks != 0 ? 1u : 0u);
                    }
                    umma_commit(&v_empty[vs]);
                    umma_commit(o_full);
                }
                __syncwarp();
                if (SBUF == 1 && t + 1 < T) issue_qk(t + 1);   // single S buffer: next logits after this tile's PV
            }
        }
    } else {
        // ===================== softmax / accumulate (thread = row) =====================
        const int row = warp * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
        float o_acc[D];
#pragma unroll
        for (int i = 0; i < D; ++i) o_acc[i] = 0.f;
        float m_run = -INFINITY;  // running max of scaled logits (log2 domain)
        float l_run = 0.f;
        float alpha_prev = 1.f;
        uint8_t* p_smem = smem + S::P_OFF;
        for (int t = 0; t < T; ++t) {
            const int sb = (SBUF == 2) ? (t & 1) : 0;
            const uint32_t sph = (SBUF == 2) ? ((t >> 1) & 1) : (t & 1);
            int krow, kvalid;
            attn_tile(w, t, krow, kvalid);
            mbar_wait(&s_full[sb], sph);
            tc_fence_after();
            const uint32_t s_addr = tmem_S + sb * 128 + lane_base;
            const bool full_tile = (kvalid == ATT_BKV);
            float psum = 0.f;
            float m_new, alpha;
            // pass 1: row max (logits stay in TMEM; the read bandwidth is not the limiter)
            float mx = -INFINITY;
            if (full_tile) {
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[32];
                    tmem_ld32(s_addr + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    if (c * 32 >= kvalid) break;
                    uint32_t v[32];
                    tmem_ld32(s_addr + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i < kvalid) mx = fmaxf(mx, __uint_as_float(v[i]));
                }
            }
            m_new = fmaxf(m_run, mx * p.scale_log2);
            alpha = ex2_fast(m_run - m_new);  // first tile: exp2(-inf) = 0
            // fold in the previous tile's PV result (also proves P / O buffers are free again)
            if (t > 0) {
                mbar_wait(o_full, (t - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < D / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld32(tmem_O + lane_base + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = o_acc[c * 32 + i] * alpha_prev + __uint_as_float(v[i]);
                }
            }
            // pass 2: probabilities -> smem (K-major SW128, two 64-key chunks), row sum
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                const bool live = full_tile || (c * 32 < kvalid);
                if (live) {
                    tmem_ld32(s_addr + c * 32, v);
                    tmem_ld_wait();
                }
                uint8_t* chunk = p_smem + (c >> 1) * (ATT_BQ * 128);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 pk;
                    uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int col = c * 32 + j * 8 + 2 * k;
                        float e0 = 0.f, e1 = 0.f;
                        if (full_tile || col < kvalid) e0 = ex2_fast(fmaf(__uint_as_float(v[j * 8 + 2 * k]), p.scale_log2, -m_new));
                        if (full_tile || col + 1 < kvalid) e1 = ex2_fast(fmaf(__uint_as_float(v[j * 8 + 2 * k + 1]), p.scale_log2, -m_new));
                        psum += e0 + e1;
                        const __half2 h = __floats2half2_rn(e0, e1);
                        pw[k] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    *reinterpret_cast<uint4*>(chunk + sw128_off(row, (c & 1) * 4 + j)) = pk;
                }
            }
            l_run = l_run * alpha + psum;
            m_run = m_new;
            alpha_prev = alpha;
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_ready);
        }
        // last PV result
        mbar_wait(o_full, (T - 1) & 1);
        tc_fence_after();
        const float inv_l = 1.f / l_run;
        if (p.lse != nullptr && row < w.q_valid)
            p.lse[static_cast<size_t>(w.head) * p.lse_ld + w.q_row0 + row] = m_run + log2f(l_run);
        act_t* orow = p.out + static_cast<size_t>(w.q_row0 + row) * p.out_ld + w.head * D;
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
            uint32_t v[32];
            tmem_ld32(tmem_O + lane_base + c * 32, v);
            tmem_ld_wait();
            if (row < w.q_valid) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 pk;
                    __half2* ph2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int i = c * 32 + j * 8 + 2 * k;
                        const float a0 = (o_acc[i] * alpha_prev + __uint_as_float(v[j * 8 + 2 * k])) * inv_l;
                        const float a1 = (o_acc[i + 1] * alpha_prev + __uint_as_float(v[j * 8 + 2 * k + 1])) * inv_l;
                        ph2[k] = __floats2half2_rn(a0, a1);
                    }
                    *reinterpret_cast<uint4*>(orow + c * 32 + j * 8) = pk;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        __syncwarp();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}


// ===========================================================================
// attention64_kernel — head_dim 64 specialisation (all cross-modal sites and the 32x32 spatial self-attention,
// i.e. > 90 % of the attention FLOPs).  Differences to the generic kernel, all aimed at the softmax warps, which
// are the limiter at d = 64 (MUFU 16 ex2/clk/SM vs 512 tensor clocks per 128x128 tile):
//   * O and the row sums l accumulate in TMEM across KV tiles (P·V and P·1 MMAs with accumulate), so the per-tile
//     "fold previous O into registers" step disappears; O is rescaled in place only when the running maximum
//     moves by more than 2^8 (warp-uniform decision), which after the first tiles is rare;
//   * the row sum is taken by the tensor core from the same rounded fp16 P that feeds P·V (no FADD chain, no
//     second conversion); full 128-key tiles run a mask-free instantiation of both passes;
//   * logits are read from TMEM with the next chunk's load in flight.
// ===========================================================================
struct Attn64Smem {
    static constexpr int Q_OFF = 0;
    static constexpr int K_OFF = 16384;                 // 2 stages
    static constexpr int V_OFF = K_OFF + 2 * 16384;     // 1 stage
    static constexpr int P_OFF = V_OFF + 16384;         // 128 x 128 fp16, two 64-key chunks
    static constexpr int ONES_OFF = P_OFF + 32768;      // K-major ones tile [16][128]: two chunks of 16 rows x 128 B
    static constexpr int BAR_OFF = ONES_OFF + 4096;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
    static constexpr int TMEM_COLS = 256;               // S 0..127 | O 128..191 | l 192..207
};

MMD_DEVINL void tmem_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
MMD_DEVINL void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
MMD_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// exp2 of two fp32 logits through one MUFU op: pack to half2, ex2.approx.f16x2
MMD_DEVINL uint32_t ex2_h2(float lo, float hi) {
    const __half2 x = __floats2half2_rn(lo, hi);
    uint32_t r;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(*reinterpret_cast<const uint32_t*>(&x)));
    return r;
}

// exp2 on the FMA / ALU pipes (no MUFU): Cody-Waite split x = r + f, r = round(x), f in [-0.5, 0.5]; cubic minimax of 2^f
// (max relative error 7.5e-5, an order below the fp16 rounding of the probability it becomes) scaled by 2^r through the
// exponent field.  The exp unit (16 ex2/clk/SM) is the limiter of the d = 64 softmax; a fraction of the elements takes
// this path so both pipes work (the split is a template parameter of the kernels).
MMD_DEVINL float ex2_poly(float x) {
    x = fmaxf(x, -126.0f);
    const float xi = x + 12582912.0f;            // 1.5 * 2^23: the integer part lands in the low mantissa bits
    const float f = x - (xi - 12582912.0f);
    float pl = fmaf(0.05517165f, f, 0.24261112f);
    pl = fmaf(pl, f, 0.69326099f);
    pl = fmaf(pl, f, 0.99992807f);
    return __int_as_float(__float_as_int(pl) + (__float_as_int(xi) << 23));
}

// Row maximum of a 128-column logit tile in TMEM (thread = row); the next 32-column load is in flight while
// the current one is reduced.  FULL tiles carry no masking code at all.
template <bool FULL>
MMD_DEVINL float attn64_rowmax(uint32_t s_addr, int kvalid) {
    uint32_t va[32], vb[32];
    float mx = -INFINITY;
    tmem_ld32(s_addr, va);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (!FULL && c * 32 >= kvalid) break;   // ragged tile: 32-column chunks past the last valid key are never touched
        uint32_t* cur = (c & 1) ? vb : va;
        tmem_ld_wait();
        if (c < 3 && (FULL || (c + 1) * 32 < kvalid)) tmem_ld32(s_addr + (c + 1) * 32, (c & 1) ? va : vb);
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (FULL || c * 32 + i < kvalid) mx = fmaxf(mx, __uint_as_float(cur[i]));
    }
    return mx;
}

// p = exp2(s * scale - m) as fp16 into the swizzled P tile (two 64-key chunks).
// PQ of every 4 consecutive elements take ex2_poly instead of the MUFU op (0, 1 or 2).
// Returns the largest exponent argument s * scale - m of the valid columns: the streaming path of the kernels uses the
// running maximum of EARLIER tiles as m and only checks afterwards that nothing came near the fp16 range of P.
// wait_bar (optional): barrier that frees the P tile (P.V of the previous key tile retired); it is waited for only right
// before the first shared-memory store, so the TMEM load and the first 32 columns' exponentials overlap that P.V.
template <bool FULL, int PQ>
MMD_DEVINL float attn64_write_p(uint32_t s_addr, int kvalid, float scale_log2, float nm, uint8_t* p_smem, int row,
                                uint64_t* wait_bar = nullptr, uint32_t wait_parity = 0) {
    float amax = -INFINITY;
    uint32_t va[32], vb[32];
    tmem_ld32(s_addr, va);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        // ragged tile: the P.V / P.1 MMAs read ceil(kvalid / 16) 16-key steps only, all inside the chunks written here
        if (!FULL && c * 32 >= kvalid) break;
        uint32_t* cur = (c & 1) ? vb : va;
        tmem_ld_wait();
        if (c < 3 && (FULL || (c + 1) * 32 < kvalid)) tmem_ld32(s_addr + (c + 1) * 32, (c & 1) ? va : vb);
        uint8_t* chunk = p_smem + (c >> 1) * (ATT_BQ * 128);
        uint4 pks[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t* pw = reinterpret_cast<uint32_t*>(&pks[j]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int col = c * 32 + j * 8 + 2 * k;
                const float a0 = fmaf(__uint_as_float(cur[j * 8 + 2 * k]), scale_log2, nm);
                const float a1 = fmaf(__uint_as_float(cur[j * 8 + 2 * k + 1]), scale_log2, nm);
                float e0 = ex2_fast(a0);
                float e1 = (PQ == 2 || (PQ == 1 && (k & 1))) ? ex2_poly(a1) : ex2_fast(a1);
                if (!FULL) {
                    if (col >= kvalid) e0 = 0.f; else amax = fmaxf(amax, a0);
                    if (col + 1 >= kvalid) e1 = 0.f; else amax = fmaxf(amax, a1);
                } else {
                    amax = fmaxf(amax, fmaxf(a0, a1));
                }
                const __half2 h = __floats2half2_rn(e0, e1);
                pw[k] = *reinterpret_cast<const uint32_t*>(&h);
            }
        }
        if (c == 0 && wait_bar != nullptr) {
            mbar_wait(wait_bar, wait_parity);
            tc_fence_after();
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(chunk + sw128_off(row, (c & 1) * 4 + j)) = pks[j];
    }
    return amax;
}

// Rescale the O / l accumulators of this thread's row in TMEM by alpha (the running maximum moved).
MMD_DEVINL void attn64_rescale(uint32_t tmem_O, uint32_t tmem_L, uint32_t lane_base, float alpha) {
    uint32_t o[32];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        tmem_ld32(tmem_O + lane_base + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st32(tmem_O + lane_base + c * 32, o);
    }
    tmem_ld16(tmem_L + lane_base, o);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
    tmem_st16(tmem_L + lane_base, o);
    tmem_st_wait();
}

// Largest exponent the streaming pass may produce: P is fp16 (max 65504 = 2^15.999), sums and O are fp32.
constexpr float ATT_STREAM_LIMIT = 14.0f;

template <int PQ>
__global__ void __launch_bounds__(ATT_THREADS, 2) attention64_kernel(const __grid_constant__ AttnParams p, int n_items) {
    // Persistent: a CTA walks work items blockIdx.x, +gridDim.x, ... ; the TMA warp runs ahead into the next item
    // (Q as soon as the last Q·K^T of the current item has been issued, K/V as stages free up), so the per-item
    // start-up latency (Q/K fetch, barrier set-up, TMEM allocation) is paid once per CTA instead of once per item.
    using S = Attn64Smem;
    constexpr int D = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* q_full = bars;        // 1
    uint64_t* q_empty = bars + 1;   // 1
    uint64_t* k_full = bars + 2;    // 2
    uint64_t* k_empty = bars + 4;   // 2
    uint64_t* v_full = bars + 6;    // 1
    uint64_t* v_empty = bars + 7;   // 1
    uint64_t* s_full = bars + 8;    // 1
    uint64_t* p_ready = bars + 9;   // 1 (128 arrivals)
    uint64_t* o_full = bars + 10;   // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    pdl_trigger();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&p.q_map);
        tma_prefetch_desc(&p.k_map);
        tma_prefetch_desc(&p.v_map);
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
        mbar_init(v_full, 1);
        mbar_init(v_empty, 1);
        mbar_init(s_full, 1);
        mbar_init(p_ready, 128);
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    // constant ones tile for the row-sum MMA
    for (int i = threadIdx.x; i < 4096 / 16; i += ATT_THREADS)
        reinterpret_cast<uint4*>(smem + S::ONES_OFF)[i] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async_smem();
    if (warp == 5) tmem_alloc(tmem_slot, S::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // q/k/v come from the previous kernel
    const uint32_t tmem_S = tmem_base;
    const uint32_t tmem_O = tmem_base + 128;
    const uint32_t tmem_L = tmem_base + 192;

    if (warp == 4) {
        // ===================== TMA producer =====================
        // whole warp in uniform control flow (all lanes wait), one elected lane issues: a lane-guarded branch makes
        // ptxas wrap every TMA / tcgen05 instruction in an ELECT + BRA.U.ANY loop over the active lanes (gemm.cuh)
        {
            int g = 0;   // KV tiles issued so far (all items)
            int it = 0;  // items started
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const AttnWork w = attn_decode(p, item);
                mbar_wait(q_empty, (it & 1) ^ 1);   // last Q·K^T of the previous item has been issued and retired
                if (elect_one()) {
                    mbar_expect_tx(q_full, 16384);
                    tma_load_2d(smem + S::Q_OFF, &p.q_map, q_full, p.q_col0 + w.head * D, w.q_row0);
                }
                __syncwarp();
                for (int t = 0; t < w.n_tiles; ++t, ++g) {
                    const int st = g & 1;
                    int krow, kvalid;
                    attn_tile(w, t, krow, kvalid);
                    mbar_wait(&k_empty[st], ((g >> 1) & 1) ^ 1);
                    ATT_TRACE(p, 2, g, 0);   // K stage free
                    if (elect_one()) {
                        mbar_expect_tx(&k_full[st], 16384);
                        tma_load_2d(smem + S::K_OFF + st * 16384, &p.k_map, &k_full[st], p.k_col0 + w.head * D, krow);
                    }
                    __syncwarp();
                    mbar_wait(v_empty, (g & 1) ^ 1);
                    ATT_TRACE(p, 2, g, 1);   // V stage free
                    if (elect_one()) {
                        mbar_expect_tx(v_full, 16384);
                        tma_load_2d(smem + S::V_OFF, &p.v_map, v_full, p.v_col0 + w.head * D, krow);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer (uniform warp, elected lane) =====================
        {
            constexpr uint32_t idesc_qk = umma_idesc_f16(ATT_BQ, ATT_BKV, 0, 0);
            constexpr uint32_t idesc_pv = umma_idesc_f16(ATT_BQ, D, 0, 1);   // V is MN-major
            constexpr uint32_t idesc_l = umma_idesc_f16(ATT_BQ, 16, 0, 0);   // P x ones^T
            const uint64_t qd0 = umma_desc_sw128(smem_u32(smem + S::Q_OFF), 16, 1024);
            const uint64_t kd0 = umma_desc_sw128(smem_u32(smem + S::K_OFF), 16, 1024);
            const uint64_t pd0 = umma_desc_sw128(smem_u32(smem + S::P_OFF), 16, 1024);
            const uint64_t vd0 = umma_desc_sw128(smem_u32(smem + S::V_OFF), ATT_BKV * 128, 1024);   // MN-major V
            const uint64_t od0 = umma_desc_sw128(smem_u32(smem + S::ONES_OFF), 16, 1024);
            int gq = 0;   // Q·K^T tiles issued
            int itq = 0;  // items whose first Q·K^T has been issued
            // issues Q·K^T of tile t of the item `w` (first tile waits for that item's Q; last tile releases Q)
            auto issue_qk = [&](const AttnWork& w, int t) {
                if (t == 0) {
                    mbar_wait(q_full, itq & 1);
                    ++itq;
                }
                const int st = gq & 1;
                mbar_wait(&k_full[st], (gq >> 1) & 1);
                ATT_TRACE(p, 1, gq, 0);   // K tile landed (Q.K^T of tile gq can issue)
                tc_fence_after();
                const uint64_t kd = kd0 + static_cast<uint64_t>(st) * (16384 >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks)
                        umma_f16_ss(tmem_S, qd0 + 2 * ks, kd + 2 * ks, idesc_qk, ks != 0 ? 1u : 0u);
                    umma_commit(&k_empty[st]);
                    if (t == w.n_tiles - 1) umma_commit(q_empty);
                    umma_commit(s_full);
                }
                __syncwarp();
                ++gq;
            };
            int g = 0;
            bool first = true;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const AttnWork w = attn_decode(p, item);
                if (first) { issue_qk(w, 0); first = false; }
                for (int t = 0; t < w.n_tiles; ++t, ++g) {
                    int krow, kvalid;
                    attn_tile(w, t, krow, kvalid);
                    mbar_wait(p_ready, g & 1);
                    ATT_TRACE(p, 1, g, 1);   // softmax of tile g done
                    // The logits of the NEXT tile go first: S is free as soon as the softmax of this tile has read it, and the
                    // softmax warps can start on them while P.V / P.1 of this tile are still being issued and executed
                    // (measured: issuing the 16 small-N MMAs of P.V / P.1 alone takes ~840 cycles)
                    // (Within an item only: the first logits of the NEXT item wait for its Q tile, which must not hold up
                    // the last P.V of this one — measured 7400 -> 9600 cycles per item boundary when it did.)
                    if (t + 1 < w.n_tiles) issue_qk(w, t + 1);
                    mbar_wait(v_full, g & 1);
                    ATT_TRACE(p, 1, g, 2);   // V tile landed
                    tc_fence_after();
                    const int nks = (kvalid + 15) >> 4;
                    if (elect_one()) {
                        // descriptors are base + compile-time increments (16-byte units)
                        if (nks == ATT_BKV / 16) {
                            // full tile: straight-line issue with compile-time descriptor increments
#pragma unroll
                            for (int ks = 0; ks < ATT_BKV / 16; ++ks)
                                umma_f16_ss(tmem_O, pd0 + ((ks >> 2) * (ATT_BQ * 128 >> 4) + (ks & 3) * 2), vd0 + ks * (2048 >> 4),
                                            idesc_pv, (t | ks) != 0 ? 1u : 0u);
#pragma unroll
                            for (int ks = 0; ks < ATT_BKV / 16; ++ks)
                                umma_f16_ss(tmem_L, pd0 + ((ks >> 2) * (ATT_BQ * 128 >> 4) + (ks & 3) * 2),
                                            od0 + ((ks >> 2) * (2048 >> 4) + (ks & 3) * 2), idesc_l, (t | ks) != 0 ? 1u : 0u);
                        } else {
                            // ragged tile: running descriptors (a 64-column chunk boundary after ks = 3)
                            uint64_t pd = pd0, vd = vd0;
                            for (int ks = 0; ks < nks; ++ks) {
                                umma_f16_ss(tmem_O, pd, vd, idesc_pv, (t | ks) != 0 ? 1u : 0u);
                                vd += 2048 >> 4;
                                pd += (ks == 3) ? (ATT_BQ * 128 >> 4) - 6 : 2;
                            }
                            uint64_t od = od0;
                            pd = pd0;
                            for (int ks = 0; ks < nks; ++ks) {
                                umma_f16_ss(tmem_L, pd, od, idesc_l, (t | ks) != 0 ? 1u : 0u);
                                pd += (ks == 3) ? (ATT_BQ * 128 >> 4) - 6 : 2;
                                od += (ks == 3) ? (2048 >> 4) - 6 : 2;
                            }
                        }
                        umma_commit(v_empty);
                        umma_commit(o_full);
                    }
                    __syncwarp();
                    ATT_TRACE(p, 1, g, 3);   // P.V of tile g issued
                    if (t + 1 == w.n_tiles && item + static_cast<int>(gridDim.x) < n_items) {
                        const AttnWork wn = attn_decode(p, item + gridDim.x);
                        issue_qk(wn, 0);
                    }
                }
            }
        }
    } else {
        // ===================== softmax warps (thread = query row) =====================
        const int row = warp * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
        const uint32_t s_addr = tmem_S + lane_base;
        uint8_t* p_smem = smem + S::P_OFF;
        int g = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const AttnWork w = attn_decode(p, item);
            const int T = w.n_tiles;
            float m_used = 0.f;
            // ragged query blocks (400 / 100 / 25 audio tokens per segment): a warp whose 32 rows all lie past the block
            // skips the whole softmax (the exp unit is the limiter) and only keeps the barrier protocol going; its
            // P / O rows are garbage that is never stored
            const bool warp_active = warp * 32 < w.q_valid;
            for (int t = 0; t < T; ++t, ++g) {
                int krow, kvalid;
                attn_tile(w, t, krow, kvalid);
                mbar_wait(s_full, g & 1);
                if (warp == 0) ATT_TRACE(p, 0, g, 0);   // logits of tile g ready
                tc_fence_after();
                if (!warp_active) {
                    if (t > 0) {
                        mbar_wait(o_full, (g - 1) & 1);
                        tc_fence_after();
                    }
                    tc_fence_before();
                    mbar_arrive(p_ready);
                    continue;
                }
                const bool full_tile = (kvalid == ATT_BKV);
                // First tile of an item: the true row maximum has to be known before any exponential (a max pass, then the
                // probability pass).  Later tiles: ONE streaming pass over the logits (TMEM reads, 64 B/clk/SM, cost as much
                // as the exponentials) — the probabilities are taken against the running maximum of the EARLIER tiles (P is
                // fp16 and O / l are fp32, so exponents up to 2^14 are harmless), and only if a row overshoots that range
                // (rare after the first tile) the accumulators are rescaled and the pass is repeated.
                if (t == 0) {
                    const float mx = full_tile ? attn64_rowmax<true>(s_addr, kvalid) : attn64_rowmax<false>(s_addr, kvalid);
                    m_used = mx * p.scale_log2;
                }
                // t >= 1: P (and, for a fix-up, O) may only be touched once P·V of the previous tile has retired; the wait
                // sits inside the first pass, right before its first store
#pragma unroll 1
                for (int attempt = 0; attempt < 2; ++attempt) {
                    uint64_t* wb = (t > 0 && attempt == 0) ? o_full : nullptr;
                    const uint32_t wp = static_cast<uint32_t>((g - 1) & 1);
                    const float amax = full_tile ? attn64_write_p<true, PQ>(s_addr, kvalid, p.scale_log2, -m_used, p_smem, row, wb, wp)
                                                 : attn64_write_p<false, PQ>(s_addr, kvalid, p.scale_log2, -m_used, p_smem, row, wb, wp);
                    if (warp == 0 && attempt == 0) ATT_TRACE(p, 0, g, 1);   // first pass done (P.V of tile g - 1 retired inside it)
                    if (t == 0 || attempt == 1 || !__any_sync(0xffffffffu, amax > ATT_STREAM_LIMIT)) break;
                    const float m_new = m_used + fmaxf(amax, 0.f);
                    attn64_rescale(tmem_O, tmem_L, lane_base, ex2_fast(m_used - m_new));
                    m_used = m_new;
                }
                fence_proxy_async_smem();
                tc_fence_before();
                if (warp == 0) ATT_TRACE(p, 0, g, 2);   // probabilities of tile g written
                mbar_arrive(p_ready);
            }
            // ---- item epilogue: O / l -> global (the next item's first P·V waits for our next p_ready arrival)
            mbar_wait(o_full, (g - 1) & 1);
            tc_fence_after();
            uint32_t lv[16];
            tmem_ld16(tmem_L + lane_base, lv);
            tmem_ld_wait();
            const float inv_l = 1.f / __uint_as_float(lv[0]);
            if (p.lse != nullptr && row < w.q_valid)
                p.lse[static_cast<size_t>(w.head) * p.lse_ld + w.q_row0 + row] = m_used + log2f(__uint_as_float(lv[0]));
            act_t* orow = p.out + static_cast<size_t>(w.q_row0 + row) * p.out_ld + w.head * D;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem_O + lane_base + c * 32, v);
                tmem_ld_wait();
                if (row < w.q_valid) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 pk;
                        __half2* ph2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            ph2[k] = __floats2half2_rn(__uint_as_float(v[j * 8 + 2 * k]) * inv_l, __uint_as_float(v[j * 8 + 2 * k + 1]) * inv_l);
                        *reinterpret_cast<uint4*>(orow + c * 32 + j * 8) = pk;
                    }
                }
            }
            tc_fence_before();   // O / l reads are complete before this thread's next p_ready arrival
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        __syncwarp();
        tmem_dealloc(tmem_base, S::TMEM_COLS);
    }
}


// ---- attention64t_kernel: P never touches shared memory.  The probabilities go back to TENSOR memory (fp16, two per
// 32-bit column, 64 columns next to S / O) and P.V reads its A operand from there (tcgen05.mma with a TMEM A operand);
// the row sums are fp32 registers.  Per 128-key tile this removes the 32 KB of P stores and the 64 KB the P.V / P.1 MMAs read
// back: the two CTAs of an SM were spending ~1400 of ~2200 cycles per tile on shared-memory bandwidth.
struct Attn64tSmem {
    static constexpr int Q_OFF = 0;
    static constexpr int K_OFF = 16384;                 // 2 stages
    static constexpr int V_OFF = K_OFF + 2 * 16384;     // 2 stages (the P tile's 32 KB are free)
    static constexpr int BAR_OFF = V_OFF + 2 * 16384;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
    static constexpr int TMEM_COLS = 256;               // S 0..127 | O 128..191 | P 192..255
};

// p = exp2(s * scale - m) for the thread's row: fp16 pairs into TMEM columns p_addr .. +63, sum of the (rounded) values
// into psum.  wait_bar: barrier that frees P / O (P.V of the previous tile retired), waited for before the first store.
template <bool FULL>
MMD_DEVINL float attn64_write_p_tmem(uint32_t s_addr, uint32_t p_addr, int kvalid, float scale_log2, float nm, float& psum,
                                     uint64_t* wait_bar, uint32_t wait_parity) {
    float amax = -INFINITY;
    float s0 = 0.f, s1 = 0.f;
    uint32_t va[32], vb[32];
    tmem_ld32(s_addr, va);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (!FULL && c * 32 >= kvalid) break;
        uint32_t* cur = (c & 1) ? vb : va;
        tmem_ld_wait();
        if (c < 3 && (FULL || (c + 1) * 32 < kvalid)) tmem_ld32(s_addr + (c + 1) * 32, (c & 1) ? va : vb);
        uint32_t pw[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int col = c * 32 + 2 * k;
            const float a0 = fmaf(__uint_as_float(cur[2 * k]), scale_log2, nm);
            const float a1 = fmaf(__uint_as_float(cur[2 * k + 1]), scale_log2, nm);
            float e0 = ex2_fast(a0);
            float e1 = ex2_fast(a1);
            if (!FULL) {
                if (col >= kvalid) e0 = 0.f; else amax = fmaxf(amax, a0);
                if (col + 1 >= kvalid) e1 = 0.f; else amax = fmaxf(amax, a1);
            } else {
                amax = fmaxf(amax, fmaxf(a0, a1));
            }
            const __half2 h = __floats2half2_rn(e0, e1);
            const float2 hr = __half22float2(h);
            s0 += hr.x;
            s1 += hr.y;
            pw[k] = *reinterpret_cast<const uint32_t*>(&h);
        }
        if (c == 0 && wait_bar != nullptr) {
            mbar_wait(wait_bar, wait_parity);
            tc_fence_after();
        }
        tmem_st16(p_addr + c * 16, pw);
    }
    psum = s0 + s1;
    return amax;
}

MMD_DEVINL void attn64_rescale_o(uint32_t tmem_O, uint32_t lane_base, float alpha) {
    uint32_t o[32];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        tmem_ld32(tmem_O + lane_base + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st32(tmem_O + lane_base + c * 32, o);
    }
    tmem_st_wait();
}

template <int PQ>
__global__ void __launch_bounds__(ATT_THREADS, 2) attention64t_kernel(const __grid_constant__ AttnParams p, int n_items) {
    // Persistent: a CTA walks work items blockIdx.x, +gridDim.x, ... ; the TMA warp runs ahead into the next item
    // (Q as soon as the last Q·K^T of the current item has been issued, K/V as stages free up), so the per-item
    // start-up latency (Q/K fetch, barrier set-up, TMEM allocation) is paid once per CTA instead of once per item.
    using S = Attn64tSmem;
    constexpr int D = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* q_full = bars;        // 1
    uint64_t* q_empty = bars + 1;   // 1
    uint64_t* k_full = bars + 2;    // 2
    uint64_t* k_empty = bars + 4;   // 2
    uint64_t* v_full = bars + 6;    // 2
    uint64_t* v_empty = bars + 8;   // 2
    uint64_t* s_full = bars + 10;   // 1
    uint64_t* p_ready = bars + 11;  // 1 (128 arrivals)
    uint64_t* o_full = bars + 12;   // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    pdl_trigger();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&p.q_map);
        tma_prefetch_desc(&p.k_map);
        tma_prefetch_desc(&p.v_map);
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        mbar_init(s_full, 1);
        mbar_init(p_ready, 128);
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, S::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // q/k/v come from the previous kernel
    const uint32_t tmem_S = tmem_base;
    const uint32_t tmem_O = tmem_base + 128;
    const uint32_t tmem_P = tmem_base + 192;   // P (fp16, two per column): 64 columns

    if (warp == 4) {
        // ===================== TMA producer =====================
        // whole warp in uniform control flow (all lanes wait), one elected lane issues: a lane-guarded branch makes
        // ptxas wrap every TMA / tcgen05 instruction in an ELECT + BRA.U.ANY loop over the active lanes (gemm.cuh)
        {
            int g = 0;   // KV tiles issued so far (all items)
            int it = 0;  // items started
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const AttnWork w = attn_decode(p, item);
                mbar_wait(q_empty, (it & 1) ^ 1);   // last Q·K^T of the previous item has been issued and retired
                if (elect_one()) {
                    mbar_expect_tx(q_full, 16384);
                    tma_load_2d(smem + S::Q_OFF, &p.q_map, q_full, p.q_col0 + w.head * D, w.q_row0);
                }
                __syncwarp();
                for (int t = 0; t < w.n_tiles; ++t, ++g) {
                    const int st = g & 1;
                    int krow, kvalid;
                    attn_tile(w, t, krow, kvalid);
                    mbar_wait(&k_empty[st], ((g >> 1) & 1) ^ 1);
                    ATT_TRACE(p, 2, g, 0);   // K stage free
                    if (elect_one()) {
                        mbar_expect_tx(&k_full[st], 16384);
                        tma_load_2d(smem + S::K_OFF + st * 16384, &p.k_map, &k_full[st], p.k_col0 + w.head * D, krow);
                    }
                    __syncwarp();
                    mbar_wait(&v_empty[st], ((g >> 1) & 1) ^ 1);
                    ATT_TRACE(p, 2, g, 1);   // V stage free
                    if (elect_one()) {
                        mbar_expect_tx(&v_full[st], 16384);
                        tma_load_2d(smem + S::V_OFF + st * 16384, &p.v_map, &v_full[st], p.v_col0 + w.head * D, krow);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer (uniform warp, elected lane) =====================
        {
            constexpr uint32_t idesc_qk = umma_idesc_f16(ATT_BQ, ATT_BKV, 0, 0);
            constexpr uint32_t idesc_pv = umma_idesc_f16(ATT_BQ, D, 0, 1);   // V is MN-major
            const uint64_t qd0 = umma_desc_sw128(smem_u32(smem + S::Q_OFF), 16, 1024);
            const uint64_t kd0 = umma_desc_sw128(smem_u32(smem + S::K_OFF), 16, 1024);
            const uint64_t vd0 = umma_desc_sw128(smem_u32(smem + S::V_OFF), ATT_BKV * 128, 1024);   // MN-major V
            int gq = 0;   // Q·K^T tiles issued
            int itq = 0;  // items whose first Q·K^T has been issued
            // issues Q·K^T of tile t of the item `w` (first tile waits for that item's Q; last tile releases Q)
            auto issue_qk = [&](const AttnWork& w, int t) {
                if (t == 0) {
                    mbar_wait(q_full, itq & 1);
                    ++itq;
                }
                const int st = gq & 1;
                mbar_wait(&k_full[st], (gq >> 1) & 1);
                ATT_TRACE(p, 1, gq, 0);   // K tile landed (Q.K^T of tile gq can issue)
                tc_fence_after();
                const uint64_t kd = kd0 + static_cast<uint64_t>(st) * (16384 >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks)
                        umma_f16_ss(tmem_S, qd0 + 2 * ks, kd + 2 * ks, idesc_qk, ks != 0 ? 1u : 0u);
                    umma_commit(&k_empty[st]);
                    if (t == w.n_tiles - 1) umma_commit(q_empty);
                    umma_commit(s_full);
                }
                __syncwarp();
                ++gq;
            };
            int g = 0;
            bool first = true;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const AttnWork w = attn_decode(p, item);
                if (first) { issue_qk(w, 0); first = false; }
                for (int t = 0; t < w.n_tiles; ++t, ++g) {
                    int krow, kvalid;
                    attn_tile(w, t, krow, kvalid);
                    mbar_wait(p_ready, g & 1);
                    ATT_TRACE(p, 1, g, 1);   // softmax of tile g done
                    // The logits of the NEXT tile go first: S is free as soon as the softmax of this tile has read it, and the
                    // softmax warps can start on them while P.V / P.1 of this tile are still being issued and executed
                    // (measured: issuing the 16 small-N MMAs of P.V / P.1 alone takes ~840 cycles)
                    // (Within an item only: the first logits of the NEXT item wait for its Q tile, which must not hold up
                    // the last P.V of this one — measured 7400 -> 9600 cycles per item boundary when it did.)
                    if (t + 1 < w.n_tiles) issue_qk(w, t + 1);
                    const int vst = g & 1;
                    mbar_wait(&v_full[vst], (g >> 1) & 1);
                    ATT_TRACE(p, 1, g, 2);   // V tile landed
                    tc_fence_after();
                    const int nks = (kvalid + 15) >> 4;
                    if (elect_one()) {
                        // P is the A operand straight from tensor memory: 8 columns (16 keys) per K step
                        const uint64_t vd = vd0 + static_cast<uint64_t>(vst) * (16384 >> 4);
                        if (nks == ATT_BKV / 16) {
#pragma unroll
                            for (int ks = 0; ks < ATT_BKV / 16; ++ks)
                                umma_f16_ts(tmem_O, tmem_P + ks * 8, vd + ks * (2048 >> 4), idesc_pv, (t | ks) != 0 ? 1u : 0u);
                        } else {
                            for (int ks = 0; ks < nks; ++ks)
                                umma_f16_ts(tmem_O, tmem_P + ks * 8, vd + ks * (2048 >> 4), idesc_pv, (t | ks) != 0 ? 1u : 0u);
                        }
                        umma_commit(&v_empty[vst]);
                        umma_commit(o_full);
                    }
                    __syncwarp();
                    ATT_TRACE(p, 1, g, 3);   // P.V of tile g issued
                    if (t + 1 == w.n_tiles && item + static_cast<int>(gridDim.x) < n_items) {
                        const AttnWork wn = attn_decode(p, item + gridDim.x);
                        issue_qk(wn, 0);
                    }
                }
            }
        }
    } else {
        // ===================== softmax warps (thread = query row) =====================
        const int row = warp * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
        const uint32_t s_addr = tmem_S + lane_base;
        int g = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const AttnWork w = attn_decode(p, item);
            const int T = w.n_tiles;
            float m_used = 0.f;
            float l_run = 0.f;   // row sum of the probabilities (fp32 registers; the scale of m_used)
            // ragged query blocks (400 / 100 / 25 audio tokens per segment): a warp whose 32 rows all lie past the block
            // skips the whole softmax (the exp unit is the limiter) and only keeps the barrier protocol going; its
            // P / O rows are garbage that is never stored
            const bool warp_active = warp * 32 < w.q_valid;
            for (int t = 0; t < T; ++t, ++g) {
                int krow, kvalid;
                attn_tile(w, t, krow, kvalid);
                mbar_wait(s_full, g & 1);
                if (warp == 0) ATT_TRACE(p, 0, g, 0);   // logits of tile g ready
                tc_fence_after();
                if (!warp_active) {
                    if (t > 0) {
                        mbar_wait(o_full, (g - 1) & 1);
                        tc_fence_after();
                    }
                    tc_fence_before();
                    mbar_arrive(p_ready);
                    continue;
                }
                const bool full_tile = (kvalid == ATT_BKV);
                // First tile of an item: the true row maximum has to be known before any exponential (a max pass, then the
                // probability pass).  Later tiles: ONE streaming pass over the logits (TMEM reads, 64 B/clk/SM, cost as much
                // as the exponentials) — the probabilities are taken against the running maximum of the EARLIER tiles (P is
                // fp16 and O / l are fp32, so exponents up to 2^14 are harmless), and only if a row overshoots that range
                // (rare after the first tile) the accumulators are rescaled and the pass is repeated.
                if (t == 0) {
                    const float mx = full_tile ? attn64_rowmax<true>(s_addr, kvalid) : attn64_rowmax<false>(s_addr, kvalid);
                    m_used = mx * p.scale_log2;
                }
                // t >= 1: P (and, for a fix-up, O) may only be touched once P·V of the previous tile has retired; the wait
                // sits inside the first pass, right before its first store
                float tile_sum = 0.f;
#pragma unroll 1
                for (int attempt = 0; attempt < 2; ++attempt) {
                    uint64_t* wb = (t > 0 && attempt == 0) ? o_full : nullptr;
                    const uint32_t wp = static_cast<uint32_t>((g - 1) & 1);
                    const float amax = full_tile ? attn64_write_p_tmem<true>(s_addr, tmem_P + lane_base, kvalid, p.scale_log2, -m_used, tile_sum, wb, wp)
                                                 : attn64_write_p_tmem<false>(s_addr, tmem_P + lane_base, kvalid, p.scale_log2, -m_used, tile_sum, wb, wp);
                    if (warp == 0 && attempt == 0) ATT_TRACE(p, 0, g, 1);   // first pass done (P.V of tile g - 1 retired inside it)
                    if (t == 0 || attempt == 1 || !__any_sync(0xffffffffu, amax > ATT_STREAM_LIMIT)) break;
                    const float m_new = m_used + fmaxf(amax, 0.f);
                    const float alpha = ex2_fast(m_used - m_new);
                    attn64_rescale_o(tmem_O, lane_base, alpha);
                    l_run *= alpha;
                    m_used = m_new;
                }
                l_run += tile_sum;
                tmem_st_wait();
                tc_fence_before();
                if (warp == 0) ATT_TRACE(p, 0, g, 2);   // probabilities of tile g written
                mbar_arrive(p_ready);
            }
            // ---- item epilogue: O / l -> global (the next item's first P·V waits for our next p_ready arrival)
            mbar_wait(o_full, (g - 1) & 1);
            tc_fence_after();
            const float inv_l = 1.f / l_run;
            if (p.lse != nullptr && row < w.q_valid)
                p.lse[static_cast<size_t>(w.head) * p.lse_ld + w.q_row0 + row] = m_used + log2f(l_run);
            act_t* orow = p.out + static_cast<size_t>(w.q_row0 + row) * p.out_ld + w.head * D;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem_O + lane_base + c * 32, v);
                tmem_ld_wait();
                if (row < w.q_valid) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 pk;
                        __half2* ph2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            ph2[k] = __floats2half2_rn(__uint_as_float(v[j * 8 + 2 * k]) * inv_l, __uint_as_float(v[j * 8 + 2 * k + 1]) * inv_l);
                        *reinterpret_cast<uint4*>(orow + c * 32 + j * 8) = pk;
                    }
                }
            }
            tc_fence_before();   // O / l reads are complete before this thread's next p_ready arrival
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        __syncwarp();
        tmem_dealloc(tmem_base, S::TMEM_COLS);
    }
}



// ---- column-half variants for attention64h_kernel (two threads per query row): chunks c0, c0 + 1 of the four 32-column
// chunks of a logit tile; chunk pair c0 / 2 is one 64-key chunk of P.
template <bool FULL>
MMD_DEVINL float attn64_rowmax_part(uint32_t s_addr, int kvalid, int c0) {
    // one 32-column chunk in registers at a time: two CTAs x 320 threads leave 102 registers per thread
    uint32_t v[32];
    float mx = -INFINITY;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        const int c = c0 + cc;
        if (!FULL && c * 32 >= kvalid) break;
        tmem_ld32(s_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (FULL || c * 32 + i < kvalid) mx = fmaxf(mx, __uint_as_float(v[i]));
    }
    return mx;
}

// p = exp2(s * scale - m) for this thread's 64 columns as fp16 into its 64-key chunk of P.  Two logits share ONE exp-unit
// operation: the fp32 arguments are packed to half2 (the conversion the fp16 P needs anyway) and go through
// ex2.approx.f16x2, whose result IS the packed P word.  The argument rounding costs at most 2^-11 relative in the exponent
// (0.03 % of the probability for the dominant arguments in [-1, 0], up to 0.5 % for probabilities below 2^-8).
// Returns the largest exponent argument (from the raw maximum: scale > 0).
template <bool FULL>
MMD_DEVINL float attn64_write_p_part(uint32_t s_addr, int kvalid, float scale_log2, float nm, uint8_t* p_smem, int row, int c0) {
    float smax = -INFINITY;
    uint32_t v[32];
    uint8_t* chunk = p_smem + (c0 >> 1) * (ATT_BQ * 128);
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        const int c = c0 + cc;
        if (!FULL && c * 32 >= kvalid) break;
        tmem_ld32(s_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 pk;
            uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int col = c * 32 + j * 8 + 2 * k;
                float s0 = __uint_as_float(v[j * 8 + 2 * k]), s1 = __uint_as_float(v[j * 8 + 2 * k + 1]);
                if (!FULL) {   // masked keys: exponent -inf -> probability 0, and out of the maximum
                    if (col >= kvalid) s0 = -INFINITY;
                    if (col + 1 >= kvalid) s1 = -INFINITY;
                }
                smax = fmaxf(smax, fmaxf(s0, s1));
                pw[k] = ex2_h2(fmaf(s0, scale_log2, nm), fmaf(s1, scale_log2, nm));
            }
            *reinterpret_cast<uint4*>(chunk + sw128_off(row, cc * 4 + j)) = pk;
        }
    }
    return fmaf(smax, scale_log2, nm);
}

// half 0 rescales O columns 0..31 and the row sums, half 1 O columns 32..63
MMD_DEVINL void attn64_rescale_part(uint32_t tmem_O, uint32_t tmem_L, uint32_t lane_base, float alpha, int hf) {
    uint32_t o[32];
    tmem_ld32(tmem_O + lane_base + hf * 32, o);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
    tmem_st32(tmem_O + lane_base + hf * 32, o);
    if (hf == 0) {
        tmem_ld16(tmem_L + lane_base, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st16(tmem_L + lane_base, o);
    }
    tmem_st_wait();
}

constexpr int ATT64H_THREADS = 320;   // warps 0-7 softmax (row quarter x column half), warp 8 TMA, warp 9 MMA
struct Attn64hSmem : Attn64Smem {
    static constexpr int XCH_OFF = Attn64Smem::BAR_OFF + 256;   // exchange between the column halves: row maxima 2 x 2 x 128 floats, row sums 2 x 128
    static constexpr int TOTAL = XCH_OFF + 3072 + 1024;
};

// attention64h_kernel: attention64_kernel with EIGHT softmax warps per CTA (two threads per query row, each owning 64 of
// the 128 logit columns = one 64-key chunk of P); everything else (TMA / MMA warps, barrier protocol, TMEM layout, two
// CTAs per SM) is the same.
template <int PQ>
__global__ void __launch_bounds__(ATT64H_THREADS, 2) attention64h_kernel(const __grid_constant__ AttnParams p, int n_items) {
    // Persistent: a CTA walks work items blockIdx.x, +gridDim.x, ... ; the TMA warp runs ahead into the next item
    // (Q as soon as the last Q·K^T of the current item has been issued, K/V as stages free up), so the per-item
    // start-up latency (Q/K fetch, barrier set-up, TMEM allocation) is paid once per CTA instead of once per item.
    using S = Attn64hSmem;
    constexpr int D = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* q_full = bars;        // 1
    uint64_t* q_empty = bars + 1;   // 1
    uint64_t* k_full = bars + 2;    // 2
    uint64_t* k_empty = bars + 4;   // 2
    uint64_t* v_full = bars + 6;    // 1
    uint64_t* v_empty = bars + 7;   // 1
    uint64_t* s_full = bars + 8;    // 1
    uint64_t* p_ready = bars + 9;   // 1 (256 arrivals)
    uint64_t* o_full = bars + 10;   // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    pdl_trigger();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&p.q_map);
        tma_prefetch_desc(&p.k_map);
        tma_prefetch_desc(&p.v_map);
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
        mbar_init(v_full, 1);
        mbar_init(v_empty, 1);
        mbar_init(s_full, 1);
        mbar_init(p_ready, 256);
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    // constant ones tile for the row-sum MMA
    for (int i = threadIdx.x; i < 4096 / 16; i += ATT64H_THREADS)
        reinterpret_cast<uint4*>(smem + S::ONES_OFF)[i] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async_smem();
    if (warp == 9) tmem_alloc(tmem_slot, S::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // q/k/v come from the previous kernel
    const uint32_t tmem_S = tmem_base;
    const uint32_t tmem_O = tmem_base + 128;
    const uint32_t tmem_L = tmem_base + 192;

    if (warp == 8) {
        // ===================== TMA producer =====================
        // whole warp in uniform control flow (all lanes wait), one elected lane issues: a lane-guarded branch makes
        // ptxas wrap every TMA / tcgen05 instruction in an ELECT + BRA.U.ANY loop over the active lanes (gemm.cuh)
        {
            int g = 0;   // KV tiles issued so far (all items)
            int it = 0;  // items started
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const AttnWork w = attn_decode(p, item);
                mbar_wait(q_empty, (it & 1) ^ 1);   // last Q·K^T of the previous item has been issued and retired
                if (elect_one()) {
                    mbar_expect_tx(q_full, 16384);
                    tma_load_2d(smem + S::Q_OFF, &p.q_map, q_full, p.q_col0 + w.head * D, w.q_row0);
                }
                __syncwarp();
                for (int t = 0; t < w.n_tiles; ++t, ++g) {
                    const int st = g & 1;
                    int krow, kvalid;
                    attn_tile(w, t, krow, kvalid);
                    mbar_wait(&k_empty[st], ((g >> 1) & 1) ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&k_full[st], 16384);
                        tma_load_2d(smem + S::K_OFF + st * 16384, &p.k_map, &k_full[st], p.k_col0 + w.head * D, krow);
                    }
                    __syncwarp();
                    mbar_wait(v_empty, (g & 1) ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(v_full, 16384);
                        tma_load_2d(smem + S::V_OFF, &p.v_map, v_full, p.v_col0 + w.head * D, krow);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer (uniform warp, elected lane) =====================
        {
            constexpr uint32_t idesc_qk = umma_idesc_f16(ATT_BQ, ATT_BKV, 0, 0);
            constexpr uint32_t idesc_pv = umma_idesc_f16(ATT_BQ, D, 0, 1);   // V is MN-major
            constexpr uint32_t idesc_l = umma_idesc_f16(ATT_BQ, 16, 0, 0);   // P x ones^T
            const uint64_t qd0 = umma_desc_sw128(smem_u32(smem + S::Q_OFF), 16, 1024);
            const uint64_t kd0 = umma_desc_sw128(smem_u32(smem + S::K_OFF), 16, 1024);
            const uint64_t pd0 = umma_desc_sw128(smem_u32(smem + S::P_OFF), 16, 1024);
            const uint64_t vd0 = umma_desc_sw128(smem_u32(smem + S::V_OFF), ATT_BKV * 128, 1024);   // MN-major V
            const uint64_t od0 = umma_desc_sw128(smem_u32(smem + S::ONES_OFF), 16, 1024);
            int gq = 0;   // Q·K^T tiles issued
            int itq = 0;  // items whose first Q·K^T has been issued
            // issues Q·K^T of tile t of the item `w` (first tile waits for that item's Q; last tile releases Q)
            auto issue_qk = [&](const AttnWork& w, int t) {
                if (t == 0) {
                    mbar_wait(q_full, itq & 1);
                    ++itq;
                }
                const int st = gq & 1;
                mbar_wait(&k_full[st], (gq >> 1) & 1);
                tc_fence_after();
                const uint64_t kd = kd0 + static_cast<uint64_t>(st) * (16384 >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks)
                        umma_f16_ss(tmem_S, qd0 + 2 * ks, kd + 2 * ks, idesc_qk, ks != 0 ? 1u : 0u);
                    umma_commit(&k_empty[st]);
                    if (t == w.n_tiles - 1) umma_commit(q_empty);
                    umma_commit(s_full);
                }
                __syncwarp();
                ++gq;
            };
            int g = 0;
            bool first = true;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const AttnWork w = attn_decode(p, item);
                if (first) { issue_qk(w, 0); first = false; }
                for (int t = 0; t < w.n_tiles; ++t, ++g) {
                    int krow, kvalid;
                    attn_tile(w, t, krow, kvalid);
                    mbar_wait(p_ready, g & 1);
                    mbar_wait(v_full, g & 1);
                    tc_fence_after();
                    const int nks = (kvalid + 15) >> 4;
                    if (elect_one()) {
                        // descriptors are base + compile-time increments (16-byte units)
                        if (nks == ATT_BKV / 16) {
                            // full tile: straight-line issue with compile-time descriptor increments
#pragma unroll
                            for (int ks = 0; ks < ATT_BKV / 16; ++ks)
                                umma_f16_ss(tmem_O, pd0 + ((ks >> 2) * (ATT_BQ * 128 >> 4) + (ks & 3) * 2), vd0 + ks * (2048 >> 4),
                                            idesc_pv, (t | ks) != 0 ? 1u : 0u);
#pragma unroll
                            for (int ks = 0; ks < ATT_BKV / 16; ++ks)
                                umma_f16_ss(tmem_L, pd0 + ((ks >> 2) * (ATT_BQ * 128 >> 4) + (ks & 3) * 2),
                                            od0 + ((ks >> 2) * (2048 >> 4) + (ks & 3) * 2), idesc_l, (t | ks) != 0 ? 1u : 0u);
                        } else {
                            // ragged tile: running descriptors (a 64-column chunk boundary after ks = 3)
                            uint64_t pd = pd0, vd = vd0;
                            for (int ks = 0; ks < nks; ++ks) {
                                umma_f16_ss(tmem_O, pd, vd, idesc_pv, (t | ks) != 0 ? 1u : 0u);
                                vd += 2048 >> 4;
                                pd += (ks == 3) ? (ATT_BQ * 128 >> 4) - 6 : 2;
                            }
                            uint64_t od = od0;
                            pd = pd0;
                            for (int ks = 0; ks < nks; ++ks) {
                                umma_f16_ss(tmem_L, pd, od, idesc_l, (t | ks) != 0 ? 1u : 0u);
                                pd += (ks == 3) ? (ATT_BQ * 128 >> 4) - 6 : 2;
                                od += (ks == 3) ? (2048 >> 4) - 6 : 2;
                            }
                        }
                        umma_commit(v_empty);
                        umma_commit(o_full);
                    }
                    __syncwarp();
                    // next logits: same item, or the first tile of the next item (S is free: softmax of this tile is done)
                    if (t + 1 < w.n_tiles) {
                        issue_qk(w, t + 1);
                    } else if (item + static_cast<int>(gridDim.x) < n_items) {
                        const AttnWork wn = attn_decode(p, item + gridDim.x);
                        issue_qk(wn, 0);
                    }
                }
            }
        }
    } else {
        // ===================== softmax warps (two threads per query row: column halves) =====================
        // warp w: rows 32 (w & 3) .. +31 (its TMEM lane quarter), logit columns 64 (w >> 2) .. +63 = one 64-key chunk of P.
        // The softmax of one CTA is a dependent chain per warp (TMEM load -> scale -> exp -> pack -> store) and was latency
        // bound with four warps per CTA; eight warps double the independent chains per scheduler.
        const int q4 = warp & 3, hf = warp >> 2;
        const int row = q4 * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
        const uint32_t s_addr = tmem_S + lane_base;
        uint8_t* p_smem = smem + S::P_OFF;
        float* xch = reinterpret_cast<float*>(smem + S::XCH_OFF);   // [tile parity][half][128 rows]
        int g = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const AttnWork w = attn_decode(p, item);
            const int T = w.n_tiles;
            float m_used = 0.f;
            // ragged query blocks: a row quarter past the block skips the softmax (both of its warps) and only keeps
            // the barrier protocol going; its P / O rows are garbage that is never stored
            const bool warp_active = q4 * 32 < w.q_valid;
            for (int t = 0; t < T; ++t, ++g) {
                int krow, kvalid;
                attn_tile(w, t, krow, kvalid);
                mbar_wait(s_full, g & 1);
                tc_fence_after();
                if (!warp_active) {
                    if (t > 0) {
                        mbar_wait(o_full, (g - 1) & 1);
                        tc_fence_after();
                    }
                    tc_fence_before();
                    mbar_arrive(p_ready);
                    continue;
                }
                const bool full_tile = (kvalid == ATT_BKV);
                float* mine = xch + ((g & 1) * 2 + hf) * 128 + row;
                const float* other = xch + ((g & 1) * 2 + (hf ^ 1)) * 128 + row;
                // First tile of an item: true row maximum first (the two halves exchange theirs), then the probabilities.
                // Later tiles: one streaming pass against the running maximum of the earlier tiles; the halves exchange the
                // largest exponent they produced and, if a row overshot the fp16 range of P, both rescale and repeat.
                if (t == 0) {
                    const float mx = full_tile ? attn64_rowmax_part<true>(s_addr, kvalid, 2 * hf) : attn64_rowmax_part<false>(s_addr, kvalid, 2 * hf);
                    *mine = mx;
                    named_bar_sync(1 + q4, 64);
                    m_used = fmaxf(mx, *other) * p.scale_log2;
                } else {
                    mbar_wait(o_full, (g - 1) & 1);   // P·V of the previous tile is done: P and O may be touched
                    tc_fence_after();
                }
#pragma unroll 1
                for (int attempt = 0; attempt < 2; ++attempt) {
                    const float amax = full_tile ? attn64_write_p_part<true>(s_addr, kvalid, p.scale_log2, -m_used, p_smem, row, 2 * hf)
                                                 : attn64_write_p_part<false>(s_addr, kvalid, p.scale_log2, -m_used, p_smem, row, 2 * hf);
                    if (t == 0 || attempt == 1) break;
                    *mine = amax;
                    named_bar_sync(1 + q4, 64);
                    const float arow = fmaxf(amax, *other);
                    if (!__any_sync(0xffffffffu, arow > ATT_STREAM_LIMIT)) break;
                    const float m_new = m_used + fmaxf(arow, 0.f);
                    attn64_rescale_part(tmem_O, tmem_L, lane_base, ex2_fast(m_used - m_new), hf);
                    m_used = m_new;
                }
                fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(p_ready);
            }
            // ---- item epilogue: O / l -> global (the next item's first P·V waits for our next p_ready arrival)
            mbar_wait(o_full, (g - 1) & 1);
            tc_fence_after();
            uint32_t lv[16];
            tmem_ld16(tmem_L + lane_base, lv);
            tmem_ld_wait();
            const float inv_l = 1.f / __uint_as_float(lv[0]);
            if (hf == 0 && p.lse != nullptr && row < w.q_valid)
                p.lse[static_cast<size_t>(w.head) * p.lse_ld + w.q_row0 + row] = m_used + log2f(__uint_as_float(lv[0]));
            act_t* orow = p.out + static_cast<size_t>(w.q_row0 + row) * p.out_ld + w.head * D;
            {
                const int c = hf;   // this half's 32 output columns
                uint32_t v[32];
                tmem_ld32(tmem_O + lane_base + c * 32, v);
                tmem_ld_wait();
                if (row < w.q_valid) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 pk;
                        __half2* ph2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            ph2[k] = __floats2half2_rn(__uint_as_float(v[j * 8 + 2 * k]) * inv_l, __uint_as_float(v[j * 8 + 2 * k + 1]) * inv_l);
                        *reinterpret_cast<uint4*>(orow + c * 32 + j * 8) = pk;
                    }
                }
            }
            tc_fence_before();   // O / l reads are complete before this thread's next p_ready arrival
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        __syncwarp();
        tmem_dealloc(tmem_base, S::TMEM_COLS);
    }
}




// ---- attention64th_kernel: attention64t_kernel (P in tensor memory) with EIGHT softmax warps, two threads per query
// row (column halves, as attention64h_kernel): the softmax phase is the long part of a CTA's per-tile cycle.
struct Attn64thSmem : Attn64tSmem {
    static constexpr int XCH_OFF = Attn64tSmem::BAR_OFF + 256;   // exchange between the column halves (3 KB)
    static constexpr int TOTAL = XCH_OFF + 3072 + 1024;
};

// this thread's 64 logit columns (chunks c0, c0 + 1) -> P columns 16 c0 .. +31 in tensor memory; psum = sum of the rounded values
// H2: the two exponentials of a packed pair through ONE exp-unit operation (ex2.approx.f16x2 on the half2-packed arguments;
// its result is the packed P word) instead of two fp32 ones.
template <bool FULL, bool H2 = false>
MMD_DEVINL float attn64_write_p_tmem_part(uint32_t s_addr, uint32_t p_addr, int kvalid, float scale_log2, float nm, int c0,
                                          float& psum, uint64_t* wait_bar, uint32_t wait_parity) {
    float amax = -INFINITY;
    float s0 = 0.f, s1 = 0.f;
    uint32_t v[32];
    bool waited = (wait_bar == nullptr);
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        const int c = c0 + cc;
        if (!FULL && c * 32 >= kvalid) break;
        tmem_ld32(s_addr + c * 32, v);
        tmem_ld_wait();
        uint32_t pw[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int col = c * 32 + 2 * k;
            float a0 = fmaf(__uint_as_float(v[2 * k]), scale_log2, nm);
            float a1 = fmaf(__uint_as_float(v[2 * k + 1]), scale_log2, nm);
            if (H2) {
                if (!FULL) {   // masked keys: exponent -inf -> probability 0, kept out of the maximum
                    if (col >= kvalid) a0 = -INFINITY;
                    if (col + 1 >= kvalid) a1 = -INFINITY;
                }
                amax = fmaxf(amax, fmaxf(a0, a1));
                pw[k] = ex2_h2(a0, a1);
                const float2 hr = __half22float2(*reinterpret_cast<const __half2*>(&pw[k]));
                s0 += hr.x;
                s1 += hr.y;
            } else {
                float e0 = ex2_fast(a0);
                float e1 = ex2_fast(a1);
                if (!FULL) {
                    if (col >= kvalid) e0 = 0.f; else amax = fmaxf(amax, a0);
                    if (col + 1 >= kvalid) e1 = 0.f; else amax = fmaxf(amax, a1);
                } else {
                    amax = fmaxf(amax, fmaxf(a0, a1));
                }
                const __half2 h = __floats2half2_rn(e0, e1);
                const float2 hr = __half22float2(h);
                s0 += hr.x;
                s1 += hr.y;
                pw[k] = *reinterpret_cast<const uint32_t*>(&h);
            }
        }
        if (!waited) {
            mbar_wait(wait_bar, wait_parity);
            tc_fence_after();
            waited = true;
        }
        tmem_st16(p_addr + c * 16, pw);
    }
    if (!waited) {   // (a half without valid columns still has to observe the barrier before the caller touches O)
        mbar_wait(wait_bar, wait_parity);
        tc_fence_after();
    }
    psum = s0 + s1;
    return amax;
}

MMD_DEVINL void attn64_rescale_o_part(uint32_t tmem_O, uint32_t lane_base, float alpha, int hf) {
    uint32_t o[32];
    tmem_ld32(tmem_O + lane_base + hf * 32, o);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
    tmem_st32(tmem_O + lane_base + hf * 32, o);
    tmem_st_wait();
}

template <int PQ>
__global__ void __launch_bounds__(ATT64H_THREADS, 2) attention64th_kernel(const __grid_constant__ AttnParams p, int n_items) {
    // Persistent: a CTA walks work items blockIdx.x, +gridDim.x, ... ; the TMA warp runs ahead into the next item
    // (Q as soon as the last Q·K^T of the current item has been issued, K/V as stages free up), so the per-item
    // start-up latency (Q/K fetch, barrier set-up, TMEM allocation) is paid once per CTA instead of once per item.
    using S = Attn64thSmem;
    constexpr int D = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* q_full = bars;        // 1
    uint64_t* q_empty = bars + 1;   // 1
    uint64_t* k_full = bars + 2;    // 2
    uint64_t* k_empty = bars + 4;   // 2
    uint64_t* v_full = bars + 6;    // 2
    uint64_t* v_empty = bars + 8;   // 2
    uint64_t* s_full = bars + 10;   // 1
    uint64_t* p_ready = bars + 11;  // 1 (256 arrivals)
    uint64_t* o_full = bars + 12;   // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    pdl_trigger();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&p.q_map);
        tma_prefetch_desc(&p.k_map);
        tma_prefetch_desc(&p.v_map);
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        mbar_init(s_full, 1);
        mbar_init(p_ready, 256);
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    if (warp == 9) tmem_alloc(tmem_slot, S::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // q/k/v come from the previous kernel
    const uint32_t tmem_S = tmem_base;
    const uint32_t tmem_O = tmem_base + 128;
    const uint32_t tmem_P = tmem_base + 192;   // P (fp16, two per column): 64 columns

    if (warp == 8) {
        // ===================== TMA producer =====================
        // whole warp in uniform control flow (all lanes wait), one elected lane issues: a lane-guarded branch makes
        // ptxas wrap every TMA / tcgen05 instruction in an ELECT + BRA.U.ANY loop over the active lanes (gemm.cuh)
        {
            int g = 0;   // KV tiles issued so far (all items)
            int it = 0;  // items started
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const AttnWork w = attn_decode(p, item);
                mbar_wait(q_empty, (it & 1) ^ 1);   // last Q·K^T of the previous item has been issued and retired
                if (elect_one()) {
                    mbar_expect_tx(q_full, 16384);
                    tma_load_2d(smem + S::Q_OFF, &p.q_map, q_full, p.q_col0 + w.head * D, w.q_row0);
                }
                __syncwarp();
                for (int t = 0; t < w.n_tiles; ++t, ++g) {
                    const int st = g & 1;
                    int krow, kvalid;
                    attn_tile(w, t, krow, kvalid);
                    mbar_wait(&k_empty[st], ((g >> 1) & 1) ^ 1);
                    ATT_TRACE(p, 2, g, 0);   // K stage free
                    if (elect_one()) {
                        mbar_expect_tx(&k_full[st], 16384);
                        tma_load_2d(smem + S::K_OFF + st * 16384, &p.k_map, &k_full[st], p.k_col0 + w.head * D, krow);
                    }
                    __syncwarp();
                    mbar_wait(&v_empty[st], ((g >> 1) & 1) ^ 1);
                    ATT_TRACE(p, 2, g, 1);   // V stage free
                    if (elect_one()) {
                        mbar_expect_tx(&v_full[st], 16384);
                        tma_load_2d(smem + S::V_OFF + st * 16384, &p.v_map, &v_full[st], p.v_col0 + w.head * D, krow);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer (uniform warp, elected lane) =====================
        {
            constexpr uint32_t idesc_qk = umma_idesc_f16(ATT_BQ, ATT_BKV, 0, 0);
            constexpr uint32_t idesc_pv = umma_idesc_f16(ATT_BQ, D, 0, 1);   // V is MN-major
            const uint64_t qd0 = umma_desc_sw128(smem_u32(smem + S::Q_OFF), 16, 1024);
            const uint64_t kd0 = umma_desc_sw128(smem_u32(smem + S::K_OFF), 16, 1024);
            const uint64_t vd0 = umma_desc_sw128(smem_u32(smem + S::V_OFF), ATT_BKV * 128, 1024);   // MN-major V
            int gq = 0;   // Q·K^T tiles issued
            int itq = 0;  // items whose first Q·K^T has been issued
            // issues Q·K^T of tile t of the item `w` (first tile waits for that item's Q; last tile releases Q)
            auto issue_qk = [&](const AttnWork& w, int t) {
                if (t == 0) {
                    mbar_wait(q_full, itq & 1);
                    ++itq;
                }
                const int st = gq & 1;
                mbar_wait(&k_full[st], (gq >> 1) & 1);
                ATT_TRACE(p, 1, gq, 0);   // K tile landed (Q.K^T of tile gq can issue)
                tc_fence_after();
                const uint64_t kd = kd0 + static_cast<uint64_t>(st) * (16384 >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks)
                        umma_f16_ss(tmem_S, qd0 + 2 * ks, kd + 2 * ks, idesc_qk, ks != 0 ? 1u : 0u);
                    umma_commit(&k_empty[st]);
                    if (t == w.n_tiles - 1) umma_commit(q_empty);
                    umma_commit(s_full);
                }
                __syncwarp();
                ++gq;
            };
            int g = 0;
            bool first = true;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const AttnWork w = attn_decode(p, item);
                if (first) { issue_qk(w, 0); first = false; }
                for (int t = 0; t < w.n_tiles; ++t, ++g) {
                    int krow, kvalid;
                    attn_tile(w, t, krow, kvalid);
                    mbar_wait(p_ready, g & 1);
                    ATT_TRACE(p, 1, g, 1);   // softmax of tile g done
                    // The logits of the NEXT tile go first: S is free as soon as the softmax of this tile has read it, and the
                    // softmax warps can start on them while P.V / P.1 of this tile are still being issued and executed
                    // (measured: issuing the 16 small-N MMAs of P.V / P.1 alone takes ~840 cycles)
                    // (Within an item only: the first logits of the NEXT item wait for its Q tile, which must not hold up
                    // the last P.V of this one — measured 7400 -> 9600 cycles per item boundary when it did.)
                    if (t + 1 < w.n_tiles) issue_qk(w, t + 1);
                    const int vst = g & 1;
                    mbar_wait(&v_full[vst], (g >> 1) & 1);
                    ATT_TRACE(p, 1, g, 2);   // V tile landed
                    tc_fence_after();
                    const int nks = (kvalid + 15) >> 4;
                    if (elect_one()) {
                        // P is the A operand straight from tensor memory: 8 columns (16 keys) per K step
                        const uint64_t vd = vd0 + static_cast<uint64_t>(vst) * (16384 >> 4);
                        if (nks == ATT_BKV / 16) {
#pragma unroll
                            for (int ks = 0; ks < ATT_BKV / 16; ++ks)
                                umma_f16_ts(tmem_O, tmem_P + ks * 8, vd + ks * (2048 >> 4), idesc_pv, (t | ks) != 0 ? 1u : 0u);
                        } else {
                            for (int ks = 0; ks < nks; ++ks)
                                umma_f16_ts(tmem_O, tmem_P + ks * 8, vd + ks * (2048 >> 4), idesc_pv, (t | ks) != 0 ? 1u : 0u);
                        }
                        umma_commit(&v_empty[vst]);
                        umma_commit(o_full);
                    }
                    __syncwarp();
                    ATT_TRACE(p, 1, g, 3);   // P.V of tile g issued
                    if (t + 1 == w.n_tiles && item + static_cast<int>(gridDim.x) < n_items) {
                        const AttnWork wn = attn_decode(p, item + gridDim.x);
                        issue_qk(wn, 0);
                    }
                }
            }
        }
    } else {
        // ===================== softmax warps (two threads per query row: column halves) =====================
        // warp w: rows 32 (w & 3) .. +31 (its TMEM lane quarter), logit columns 64 (w >> 2) .. +63 -> P columns 32 (w >> 2) .. +31
        const int q4 = warp & 3, hf = warp >> 2;
        const int row = q4 * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
        const uint32_t s_addr = tmem_S + lane_base;
        float* xch = reinterpret_cast<float*>(smem + S::XCH_OFF);   // [tile parity][half][128 rows] + row sums [half][128]
        int g = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const AttnWork w = attn_decode(p, item);
            const int T = w.n_tiles;
            float m_used = 0.f;
            float l_run = 0.f;   // this thread's share of the row sum (its 64 columns of every tile)
            const bool warp_active = q4 * 32 < w.q_valid;
            for (int t = 0; t < T; ++t, ++g) {
                int krow, kvalid;
                attn_tile(w, t, krow, kvalid);
                mbar_wait(s_full, g & 1);
                if (warp == 0) ATT_TRACE(p, 0, g, 0);   // logits of tile g ready
                tc_fence_after();
                if (!warp_active) {
                    if (t > 0) {
                        mbar_wait(o_full, (g - 1) & 1);
                        tc_fence_after();
                    }
                    tc_fence_before();
                    mbar_arrive(p_ready);
                    continue;
                }
                const bool full_tile = (kvalid == ATT_BKV);
                float* mine = xch + ((g & 1) * 2 + hf) * 128 + row;
                const float* other = xch + ((g & 1) * 2 + (hf ^ 1)) * 128 + row;
                if (t == 0) {
                    const float mx = full_tile ? attn64_rowmax_part<true>(s_addr, kvalid, 2 * hf) : attn64_rowmax_part<false>(s_addr, kvalid, 2 * hf);
                    *mine = mx;
                    named_bar_sync(1 + q4, 64);
                    m_used = fmaxf(mx, *other) * p.scale_log2;
                }
                float tile_sum = 0.f;
#pragma unroll 1
                for (int attempt = 0; attempt < 2; ++attempt) {
                    uint64_t* wb = (t > 0 && attempt == 0) ? o_full : nullptr;
                    const uint32_t wp = static_cast<uint32_t>((g - 1) & 1);
                    const float amax = full_tile ? attn64_write_p_tmem_part<true, PQ == 1>(s_addr, tmem_P + lane_base, kvalid, p.scale_log2, -m_used, 2 * hf, tile_sum, wb, wp)
                                                 : attn64_write_p_tmem_part<false, PQ == 1>(s_addr, tmem_P + lane_base, kvalid, p.scale_log2, -m_used, 2 * hf, tile_sum, wb, wp);
                    if (warp == 0 && attempt == 0) ATT_TRACE(p, 0, g, 1);
                    if (t == 0 || attempt == 1) break;
                    *mine = amax;
                    named_bar_sync(1 + q4, 64);
                    const float arow = fmaxf(amax, *other);
                    if (!__any_sync(0xffffffffu, arow > ATT_STREAM_LIMIT)) break;
                    const float m_new = m_used + fmaxf(arow, 0.f);
                    const float alpha = ex2_fast(m_used - m_new);
                    attn64_rescale_o_part(tmem_O, lane_base, alpha, hf);
                    l_run *= alpha;
                    m_used = m_new;
                }
                l_run += tile_sum;
                tmem_st_wait();
                tc_fence_before();
                if (warp == 0) ATT_TRACE(p, 0, g, 2);   // probabilities of tile g written
                mbar_arrive(p_ready);
            }
            // ---- item epilogue: O / l -> global
            mbar_wait(o_full, (g - 1) & 1);
            tc_fence_after();
            xch[512 + hf * 128 + row] = l_run;
            named_bar_sync(1 + q4, 64);
            const float l_tot = l_run + xch[512 + (hf ^ 1) * 128 + row];
            const float inv_l = 1.f / l_tot;
            if (hf == 0 && p.lse != nullptr && row < w.q_valid)
                p.lse[static_cast<size_t>(w.head) * p.lse_ld + w.q_row0 + row] = m_used + log2f(l_tot);
            act_t* orow = p.out + static_cast<size_t>(w.q_row0 + row) * p.out_ld + w.head * D;
            {
                const int c = hf;   // this half's 32 output columns
                uint32_t v[32];
                tmem_ld32(tmem_O + lane_base + c * 32, v);
                tmem_ld_wait();
                if (row < w.q_valid) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 pk;
                        __half2* ph2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            ph2[k] = __floats2half2_rn(__uint_as_float(v[j * 8 + 2 * k]) * inv_l, __uint_as_float(v[j * 8 + 2 * k + 1]) * inv_l);
                        *reinterpret_cast<uint4*>(orow + c * 32 + j * 8) = pk;
                    }
                }
            }
            tc_fence_before();   // O reads are complete before this thread's next p_ready arrival
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        __syncwarp();
        tmem_dealloc(tmem_base, S::TMEM_COLS);
    }
}





// ===========================================================================
// attention64x2_kernel — head_dim 64 with TWO 128-row query tiles per CTA sharing every K/V tile.
// One CTA per SM (all 512 TMEM columns: S0 S1 O0 O1 P0 P1; ~100 KB shared memory for Q x 2, K x 2, V x 2): two softmax
// groups of four warps each own one query tile, one TMA warp, one MMA warp.  While group 0 is in its softmax the tensor
// core runs group 1's Q·K^T / P·V and vice versa, and every K/V tile is fetched once for 256 query rows.  P lives in tensor
// memory and the row sums in registers as in attention64t_kernel (shared helpers); within an item the next tile's logits
// of a group are issued before its P·V.  Items whose query block has no second tile (<= 128 rows) run with group 1 idle.
// ===========================================================================
struct Attn64x2Smem {
    static constexpr int Q_OFF = 0;                       // two query tiles
    static constexpr int K_OFF = 2 * 16384;               // 2 stages
    static constexpr int V_OFF = K_OFF + 2 * 16384;       // 2 stages
    static constexpr int BAR_OFF = V_OFF + 2 * 16384;     // (P lives in tensor memory, the row sums in registers)
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
    static constexpr int TMEM_COLS = 512;                 // S0 0 | S1 128 | O0 256 | O1 320 | P0 384 | P1 448
};
constexpr int ATT2_THREADS = 384;   // warps 0-3: softmax of tile 0, 4-7: softmax of tile 1, 8: TMA, 9: MMA, 10-11: idle
// Three warpgroups so the register file can be re-split (setmaxnreg): the softmax threads hold a whole 128-column logit
// row in registers, the producer warpgroup needs next to nothing.
constexpr int ATT2_REGS_SOFTMAX = 208;
constexpr int ATT2_REGS_PRODUCER = 88;

struct AttnWork2 {
    AttnWork w;      // tile 0 of the pair (q_row0, key segments, head)
    int qv[2];       // valid rows of the two tiles (qv[1] == 0: single-tile item)
};
MMD_DEVINL AttnWork2 attn_decode2(const AttnParams& p, int q_pairs, int idx) {
    const int qp = idx % q_pairs;
    const int rest = idx / q_pairs;
    AttnWork2 r;
    r.w = attn_decode(p, rest * p.q_tiles + 2 * qp);
    r.qv[0] = r.w.q_valid;
    r.qv[1] = max(0, min(ATT_BQ, p.q_blk - (2 * qp + 1) * ATT_BQ));
    return r;
}

template <int PQ>
__global__ void __launch_bounds__(ATT2_THREADS, 1) attention64x2_kernel(const __grid_constant__ AttnParams p, int n_items, int q_pairs) {
    using S = Attn64x2Smem;
    constexpr int D = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* q_full = bars;         // 1
    uint64_t* q_empty = bars + 1;    // 1
    uint64_t* k_full = bars + 2;     // 2
    uint64_t* k_empty = bars + 4;    // 2
    uint64_t* v_full = bars + 6;     // 2
    uint64_t* v_empty = bars + 8;    // 2 (two arrivals: one per query tile's P·V)
    uint64_t* s_full = bars + 10;    // 2 (per group)
    uint64_t* p_ready = bars + 12;   // 2 (128 arrivals each)
    uint64_t* o_full = bars + 14;    // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    pdl_trigger();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&p.q_map);
        tma_prefetch_desc(&p.k_map);
        tma_prefetch_desc(&p.v_map);
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 2);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_ready[i], 128);
            mbar_init(&o_full[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 9) tmem_alloc(tmem_slot, S::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // q/k/v come from the previous kernel

    if (warp == 8) {
        // ===================== TMA producer =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ATT2_REGS_PRODUCER));
        {   // uniform warp, elected lane issues (see gemm.cuh)
            int n = 0;    // K/V tiles issued so far (all items)
            int it = 0;   // items started
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const AttnWork2 w2 = attn_decode2(p, q_pairs, item);
                const AttnWork& w = w2.w;
                const bool two = w2.qv[1] > 0;
                mbar_wait(q_empty, (it & 1) ^ 1);   // the last Q·K^T of the previous item has retired
                if (elect_one()) {
                    mbar_expect_tx(q_full, two ? 32768 : 16384);
                    tma_load_2d(smem + S::Q_OFF, &p.q_map, q_full, p.q_col0 + w.head * D, w.q_row0);
                    if (two) tma_load_2d(smem + S::Q_OFF + 16384, &p.q_map, q_full, p.q_col0 + w.head * D, w.q_row0 + ATT_BQ);
                }
                __syncwarp();
                for (int t = 0; t < w.n_tiles; ++t, ++n) {
                    const int st = n & 1;
                    const uint32_t ph = (n >> 1) & 1;
                    int krow, kvalid;
                    attn_tile(w, t, krow, kvalid);
                    mbar_wait(&k_empty[st], ph ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&k_full[st], 16384);
                        tma_load_2d(smem + S::K_OFF + st * 16384, &p.k_map, &k_full[st], p.k_col0 + w.head * D, krow);
                    }
                    __syncwarp();
                    mbar_wait(&v_empty[st], ph ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&v_full[st], 16384);
                        tma_load_2d(smem + S::V_OFF + st * 16384, &p.v_map, &v_full[st], p.v_col0 + w.head * D, krow);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ATT2_REGS_PRODUCER));
        {   // uniform warp, elected lane issues
            constexpr uint32_t idesc_qk = umma_idesc_f16(ATT_BQ, ATT_BKV, 0, 0);
            const uint64_t qd0 = umma_desc_sw128(smem_u32(smem + S::Q_OFF), 16, 1024);
            const uint64_t kd0 = umma_desc_sw128(smem_u32(smem + S::K_OFF), 16, 1024);
            const uint64_t vd0 = umma_desc_sw128(smem_u32(smem + S::V_OFF), ATT_BKV * 128, 1024);   // MN-major V
            constexpr uint32_t idesc_pv = umma_idesc_f16(ATT_BQ, D, 0, 1);   // V is MN-major
            int itq = 0;            // items whose Q has been waited for
            int gs[2] = {0, 0};     // tiles processed per group (phases of p_ready / o_full; s_full runs one ahead)
            // Q·K^T of K tile `kidx` (global tile counter) of an item for group grp; `last_of_tile`: no further Q·K^T
            // reads this K stage; `last_of_item`: nor this item's Q
            auto issue_qk = [&](int grp, int kidx, bool first_of_item, bool last_of_tile, bool last_of_item) {
                if (first_of_item) {
                    mbar_wait(q_full, itq & 1);
                    ++itq;
                }
                const int st = kidx & 1;
                mbar_wait(&k_full[st], (kidx >> 1) & 1);
                tc_fence_after();
                const uint64_t qd = qd0 + static_cast<uint64_t>(grp) * (16384 >> 4);
                const uint64_t kd = kd0 + static_cast<uint64_t>(st) * (16384 >> 4);
                const uint32_t tS = tmem_base + grp * 128;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks) umma_f16_ss(tS, qd + 2 * ks, kd + 2 * ks, idesc_qk, ks != 0 ? 1u : 0u);
                    if (last_of_tile) umma_commit(&k_empty[st]);
                    if (last_of_tile && last_of_item) umma_commit(q_empty);
                    umma_commit(&s_full[grp]);
                }
                __syncwarp();
            };
            int n = 0;   // global index of the current K/V tile
            bool have = static_cast<int>(blockIdx.x) < n_items;
            AttnWork2 cur{};
            if (have) {
                cur = attn_decode2(p, q_pairs, blockIdx.x);
                const bool two = cur.qv[1] > 0;
                issue_qk(0, 0, true, !two, cur.w.n_tiles == 1);
                if (two) issue_qk(1, 0, false, true, cur.w.n_tiles == 1);
            }
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const bool two = cur.qv[1] > 0;
                const int T = cur.w.n_tiles;
                const bool has_next = item + static_cast<int>(gridDim.x) < n_items;
                AttnWork2 nxt{};
                if (has_next) nxt = attn_decode2(p, q_pairs, item + gridDim.x);
                const bool next_two = has_next && nxt.qv[1] > 0;
                for (int t = 0; t < T; ++t, ++n) {
                    int krow, kvalid;
                    attn_tile(cur.w, t, krow, kvalid);
                    const int vst = n & 1;
                    const uint32_t vph = (n >> 1) & 1;
                    const uint64_t vd = vd0 + static_cast<uint64_t>(vst) * (16384 >> 4);
                    for (int grp = 0; grp < (two ? 2 : 1); ++grp) {
                        mbar_wait(&p_ready[grp], gs[grp] & 1);
                        // next logits of this group first (S is free once its softmax has read it), then P.V — inside an
                        // item only: the next item's first logits wait for its Q tiles and go after the last P.V
                        if (t + 1 < T) issue_qk(grp, n + 1, false, grp == (two ? 1 : 0), t + 2 == T);
                        mbar_wait(&v_full[vst], vph);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t tO = tmem_base + 256 + grp * 64, tP = tmem_base + 384 + grp * 64;
                            const int nks = (kvalid + 15) >> 4;
                            for (int ks = 0; ks < nks; ++ks)
                                umma_f16_ts(tO, tP + ks * 8, vd + static_cast<uint64_t>(ks) * (2048 >> 4), idesc_pv, (t | ks) != 0 ? 1u : 0u);
                            umma_commit(&v_empty[vst]);
                            umma_commit(&o_full[grp]);
                        }
                        __syncwarp();
                        ++gs[grp];
                        if (t + 1 == T && has_next && (grp == 0 || next_two))
                            issue_qk(grp, n + 1, grp == 0, grp == (next_two ? 1 : 0), nxt.w.n_tiles == 1);
                    }
                    if (!two) {
                        if (elect_one()) umma_commit(&v_empty[vst]);   // the idle group's arrival
                        __syncwarp();
                        if (t + 1 == T && next_two) issue_qk(1, n + 1, false, true, nxt.w.n_tiles == 1);
                    }
                }
                cur = nxt;
            }
        }
    } else if (warp >= 10) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ATT2_REGS_PRODUCER));   // idle half of the producer warpgroup
    } else {
        // ===================== softmax groups (thread = query row of the group's tile) =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(ATT2_REGS_SOFTMAX));
        const int grp = warp >> 2;
        const int wg = warp & 3;
        const int row = wg * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(wg * 32) << 16;
        const uint32_t tmem_S = tmem_base + grp * 128;
        const uint32_t tmem_O = tmem_base + 256 + grp * 64;
        const uint32_t tmem_P = tmem_base + 384 + grp * 64;
        const uint32_t s_addr = tmem_S + lane_base;
        uint64_t* sf = &s_full[grp];
        uint64_t* pr = &p_ready[grp];
        uint64_t* of = &o_full[grp];
        int g = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const AttnWork2 w2 = attn_decode2(p, q_pairs, item);
            const AttnWork& w = w2.w;
            const int q_valid = w2.qv[grp];
            if (q_valid <= 0) continue;   // single-tile item: group 1 sits it out (the MMA thread skips it too)
            const int q_row0 = w.q_row0 + grp * ATT_BQ;
            const int T = w.n_tiles;
            float m_used = 0.f;
            float l_run = 0.f;   // row sum of the probabilities (registers)
            const bool warp_active = wg * 32 < q_valid;
            for (int t = 0; t < T; ++t, ++g) {
                int krow, kvalid;
                attn_tile(w, t, krow, kvalid);
                mbar_wait(sf, g & 1);
                tc_fence_after();
                if (!warp_active) {
                    if (t > 0) {
                        mbar_wait(of, (g - 1) & 1);
                        tc_fence_after();
                    }
                    tc_fence_before();
                    mbar_arrive(pr);
                    continue;
                }
                const bool full_tile = (kvalid == ATT_BKV);
                // first tile: max pass + probability pass; later tiles: one streaming pass against the running maximum of the
                // earlier tiles, repeated after a rescale in the rare case a row overshoots (see attention64_kernel)
                if (t == 0) {
                    const float mx = full_tile ? attn64_rowmax<true>(s_addr, kvalid) : attn64_rowmax<false>(s_addr, kvalid);
                    m_used = mx * p.scale_log2;
                }
                // (t >= 1: P·V of the previous tile must have retired before P / O are touched; waited for inside the first
                // pass, right before its first tensor-memory store)
                float tile_sum = 0.f;
#pragma unroll 1
                for (int attempt = 0; attempt < 2; ++attempt) {
                    uint64_t* wb = (t > 0 && attempt == 0) ? of : nullptr;
                    const uint32_t wp = static_cast<uint32_t>((g - 1) & 1);
                    const float amax = full_tile ? attn64_write_p_tmem<true>(s_addr, tmem_P + lane_base, kvalid, p.scale_log2, -m_used, tile_sum, wb, wp)
                                                 : attn64_write_p_tmem<false>(s_addr, tmem_P + lane_base, kvalid, p.scale_log2, -m_used, tile_sum, wb, wp);
                    if (t == 0 || attempt == 1 || !__any_sync(0xffffffffu, amax > ATT_STREAM_LIMIT)) break;
                    const float m_new = m_used + fmaxf(amax, 0.f);
                    const float alpha = ex2_fast(m_used - m_new);
                    attn64_rescale_o(tmem_O, lane_base, alpha);
                    l_run *= alpha;
                    m_used = m_new;
                }
                l_run += tile_sum;
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(pr);
            }
            // ---- item epilogue: O / l -> global
            mbar_wait(of, (g - 1) & 1);
            tc_fence_after();
            if (warp_active) {
                const float inv_l = 1.f / l_run;
                if (p.lse != nullptr && row < q_valid)
                    p.lse[static_cast<size_t>(w.head) * p.lse_ld + q_row0 + row] = m_used + log2f(l_run);
                act_t* orow = p.out + static_cast<size_t>(q_row0 + row) * p.out_ld + w.head * D;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    tmem_ld32(tmem_O + lane_base + c * 32, v);
                    tmem_ld_wait();
                    if (row < q_valid) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 pk;
                            __half2* ph2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                ph2[k] = __floats2half2_rn(__uint_as_float(v[j * 8 + 2 * k]) * inv_l, __uint_as_float(v[j * 8 + 2 * k + 1]) * inv_l);
                            *reinterpret_cast<uint4*>(orow + c * 32 + j * 8) = pk;
                        }
                    }
                }
            }
            tc_fence_before();   // O / l reads are complete before this thread's next p_ready arrival
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        __syncwarp();
        tmem_dealloc(tmem_base, S::TMEM_COLS);
    }
}

}  // namespace mmd
