// Implicit-GEMM convolution on a CTA PAIR (tcgen05.mma.cta_group::2) for sm_100a.
//
// Same problem, parameters and epilogue as conv_gemm_kernel (gemm.cuh), but one 256-token x BN tile is computed by the two
// CTAs of a cluster: each CTA stages ITS 128 token rows of A and HALF of the weight tile (BN/2 rows), the leader CTA's
// single MMA thread issues M = 256 instructions that read A and B from both CTAs' shared memory, and each CTA ends up with
// its own 128 x BN accumulator in its own TMEM.  Why: with one CTA per tile the tensor core reads (128 + BN) x 16 operand
// elements from shared memory per K-step and the TMA writes the same bytes — at BN = 256 that is 96 KB per k-block against
// the 128 B/clk shared-memory port, i.e. the short-K GEMMs of the network were bound by shared-memory bandwidth (measured:
// neither an L2 prefetch nor deeper staging moved them).  The pair halves the weight bytes each SM stages and reads
// (64 KB per k-block at BN = 256) and fits one more pipeline stage.
//
// Cross-CTA protocol (r = %cluster_ctarank, leader = rank 0):
//   full[s]    leader's barrier, 2 arrivals: each CTA's producer arms it with its own stage bytes (the peer through a
//              remote arrive.expect_tx) and both CTAs' TMA loads complete_tx on it (.cta_group::2 loads)
//   empty[s]   one per CTA, released by the leader's tcgen05.commit multicast to both CTAs
//   tfull[a]   one per CTA, same multicast commit after the last k-block of a tile
//   tempty[a]  leader's barrier, 8 arrivals: the four epilogue warps of BOTH CTAs (the peer's through remote arrives)
#pragma once
#include "gemm.cuh"

namespace mmd {

MMD_DEVINL uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
MMD_DEVINL void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
MMD_DEVINL uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
MMD_DEVINL void mbar_arrive_cluster(uint32_t caddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
MMD_DEVINL void mbar_expect_tx_cluster(uint32_t caddr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(caddr), "r"(bytes) : "memory");
}
// TMA loads whose completion is signalled on a barrier that may live in the peer CTA (cta_group::2)
MMD_DEVINL void tma2_load_nd(int rank, void* dst, const CUtensorMap* m, uint32_t bar_caddr, const int* c) {
    const uint32_t d = smem_u32(dst);
    const uint64_t mp = reinterpret_cast<uint64_t>(m);
    switch (rank) {
        case 2:
            asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(d),
                         "l"(mp), "r"(bar_caddr), "r"(c[0]), "r"(c[1]) : "memory");
            break;
        case 3:
            asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(d),
                         "l"(mp), "r"(bar_caddr), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory");
            break;
        case 4:
            asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(d),
                         "l"(mp), "r"(bar_caddr), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory");
            break;
        default:
            asm volatile("cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(d),
                         "l"(mp), "r"(bar_caddr), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
            break;
    }
}
MMD_DEVINL void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
MMD_DEVINL void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
MMD_DEVINL void umma2_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once all previously issued MMAs have retired
MMD_DEVINL void umma2_commit_both(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
}

template <int BN, int OC>
struct Gemm2Smem {
    static constexpr int A_BYTES = GEMM_BM * 128;          // this CTA's 128 token rows
    static constexpr int B_BYTES = (BN / 2) * 128;         // this CTA's half of the weight tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NCHUNK = BN / OC;
    static constexpr int UNITS = OC / 64;
    static constexpr int OUT_BUF = UNITS * GEMM_BM * 128;
    static constexpr int OUT_BYTES = 2 * OUT_BUF;
    static constexpr int GN_BYTES = (BN / 64) * 4 * 16 * 2 * 4;
    static constexpr int BIAS_OFF = 256 + GN_BYTES;
    static constexpr int BIAS_BYTES = BN * 4;
    static constexpr int BAR_BYTES = 256 + GN_BYTES + BIAS_BYTES;
    static constexpr int LIMIT = 232448;   // 227 KB
    static constexpr int FIT = (LIMIT - OUT_BYTES - BAR_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = FIT > 8 ? 8 : FIT;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + OUT_BYTES + BAR_BYTES;
    static_assert(BN == 128 || BN == 256, "pair tiles are 256 x 128 or 256 x 256");
    static_assert(OC % 64 == 0 && BN % OC == 0, "staging chunk");
    static_assert(STAGES >= 3 && TOTAL <= LIMIT, "shared memory budget");
    static_assert(2 * STAGES + 4 <= 30, "barrier block");
    static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
};

template <int BN, int OC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1) conv_gemm2_kernel(const __grid_constant__ GemmParams p) {
    using S = Gemm2Smem<BN, OC>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* stage_base = smem;
    uint8_t* out_stage = smem + S::STAGES * S::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + S::OUT_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + S::STAGES;
    uint64_t* tfull_bar = bars + 2 * S::STAGES;
    uint64_t* tempty_bar = bars + 2 * S::STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::STAGES + 4);
    float* gn_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
    float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + S::BIAS_OFF);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool lead_cta = cta_rank == 0;
    const int n_clusters = static_cast<int>(gridDim.x >> 1);
    const int cid = static_cast<int>(blockIdx.x >> 1);
    const int m_pairs = (p.m_tiles + 1) >> 1;
    const int total_pt = m_pairs * p.n_tiles;     // pair tiles: 256 tokens x BN
    int total_chunks = 0;
    for (int s = 0; s < p.n_src; ++s) total_chunks += p.src_chunks[s];
    const int num_kb = p.n_taps * total_chunks;

    pdl_trigger();
    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) {
            printf("conv_gemm2_kernel: dynamic shared memory base not 1024-byte aligned\n");
            __trap();
        }
        for (int s = 0; s < p.n_src; ++s) tma_prefetch_desc(&p.a_map[s]);
        tma_prefetch_desc(&p.b2_map);
        tma_prefetch_desc(&p.o_map);
        for (int i = 0; i < S::STAGES; ++i) {
            mbar_init(&full_bar[i], 2);    // one arming arrival per CTA of the pair (the leader's copy is the live one)
            mbar_init(&empty_bar[i], 1);   // the leader's MMA commit, multicast to both CTAs
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 8);  // four epilogue warps of each CTA (the leader's copy is the live one)
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc2(tmem_slot, S::TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();    // barrier initialisation and the paired TMEM allocation are visible to both CTAs
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    // a tile whose row block lies past the end (odd number of row blocks: the last pair is half empty) is loaded as
    // all-out-of-bounds (zero fill, the byte counts still complete) and neither stored nor counted in the statistics
    auto tile_of = [&](int pt, int& m_idx, int& n_idx, bool& valid) {
        const int mp = pt / p.n_tiles;
        n_idx = pt - mp * p.n_tiles;
        m_idx = 2 * mp + static_cast<int>(cta_rank);
        valid = m_idx < p.m_tiles;
    };

    if (warp == 0) {
        // ================= TMA producer (one thread per CTA) =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int pt = cid; pt < total_pt; pt += n_clusters) {
                int m_idx, n_idx;
                bool valid;
                tile_of(pt, m_idx, n_idx, valid);
                int org[5];
                gemm_tile_origin(p, valid ? m_idx : 0, org);
                if (!valid) org[p.rank - 1] = 0x3fffffff / 2;   // outermost coordinate far outside: the whole box is zero-filled
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
                            const uint32_t full_lead = mapa_u32(smem_u32(&full_bar[stage]), 0);
                            mbar_expect_tx_cluster(full_lead, S::STAGE_BYTES);
                            c[0] = ch * GEMM_BK;
                            tma2_load_nd(p.rank, a_dst, &p.a_map[s], full_lead, c);
                            const int cb[2] = {kb * GEMM_BK, n_idx * BN + static_cast<int>(cta_rank) * (BN / 2)};
                            tma2_load_nd(2, a_dst + S::A_BYTES, &p.b2_map, full_lead, cb);
                            if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one thread of the leader CTA) =================
        if (lead_cta && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(2 * GEMM_BM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int pt = cid; pt < total_pt; pt += n_clusters, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(stage_base + stage * S::STAGE_BYTES);
                    const uint64_t ad0 = umma_desc_sw128(a_addr, 16, 1024);
                    const uint64_t bd0 = umma_desc_sw128(a_addr + S::A_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k)
                        umma2_f16_ss(d_tmem, ad0 + 2 * k, bd0 + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma2_commit_both(&empty_bar[stage]);
                    if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
                }
                umma2_commit_both(&tfull_bar[acc]);
            }
        }
    } else {
        // ================= epilogue (4 warps per CTA, thread = accumulator row of this CTA's half) =================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int et = threadIdx.x - 64;   // 0..127
        const bool leader = (et == 0);
        auto release_acc = [&](int acc) {
            if (lead_cta) mbar_arrive(&tempty_bar[acc]);
            else mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
        };
        int it = 0;
        uint32_t obuf_sel = 0;
        for (int pt = cid; pt < total_pt; pt += n_clusters, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            int m_idx, n_idx;
            bool valid;
            tile_of(pt, m_idx, n_idx, valid);
            int org[5];
            gemm_tile_origin(p, valid ? m_idx : 0, org);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            if (!valid) {   // empty half of the last pair: hand the accumulator back, nothing to store
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(acc);
                continue;
            }
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
            const float* bias = p.bias + n_idx * BN;
            {
                int valid_rows = GEMM_BM;
                if (p.stats != nullptr && p.stats_valid_coord >= 0)
                    valid_rows = min(GEMM_BM, p.dims[p.stats_valid_coord] - org[p.stats_valid_coord + 1]);
#pragma unroll 1
                for (int cc = 0; cc < S::NCHUNK; ++cc) {
                    uint8_t* obuf = out_stage + (obuf_sel & 1) * S::OUT_BUF;
                    ++obuf_sel;
                    if (leader) tma_store_wait_read1();  // the store issued two chunks ago has drained this buffer
                    if (cc == 0) {   // (readers of the previous tile's bias are past that tile's last barrier)
#pragma unroll
                        for (int i = et; i < BN; i += 128) bias_s[i] = __ldg(bias + i);
                    }
                    named_bar_sync(1, 128);
                    // accumulator -> fp16 staging, 32 columns at a time with the next TMEM load already in flight
                    uint32_t va[32], vb[32];
                    tmem_ld32(t_addr + cc * OC, va);
#pragma unroll
                    for (int l = 0; l < OC / 32; ++l) {
                        uint32_t* v = (l & 1) ? vb : va;
                        tmem_ld_wait();
                        if (l + 1 < OC / 32) tmem_ld32(t_addr + cc * OC + (l + 1) * 32, (l & 1) ? va : vb);
                        uint8_t* unit_base = obuf + (l >> 1) * (GEMM_BM * 128);
                        const float* bcol = bias_s + cc * OC + l * 32;
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
                        if (lane == 0) release_acc(acc);
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(1, 128);
                    if (leader) {
                        int c[5] = {0, org[1], org[2], org[3], org[4]};
#pragma unroll
                        for (int u = 0; u < S::UNITS; ++u) {
                            c[0] = n_idx * BN + cc * OC + u * 64;
                            tma_store_nd(p.rank, &p.o_map, obuf + u * (GEMM_BM * 128), c);
                        }
                        tma_store_commit();
                    }
                    if (p.stats != nullptr) {
                        // Column sums of the staged fp16 chunk without atomics, one 64-column unit at a time: lane & 15 =
                        // 4-column quad (8 bytes of a 128-byte row), the two half-warps take 16 rows each of the warp's
                        // 32-row band; per (band, quad) partials are folded into groups by the write-out pass below.
                        const int band = et >> 5, quad4 = lane & 15, half = lane >> 4;
                        const int unit = quad4 >> 1, sub = (quad4 & 1) * 8;
#pragma unroll 1
                        for (int u = 0; u < S::UNITS; ++u) {
                            const uint8_t* ub = obuf + u * (GEMM_BM * 128);
                            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll 8
                            for (int rr = 0; rr < 16; ++rr) {
                                const int r = band * 32 + half * 16 + rr;
                                if (r < valid_rows) {
                                    const uint2 raw = *reinterpret_cast<const uint2*>(ub + sw128_off(r, unit) + sub);
                                    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
                                    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
                                    s0 += a.x; q0 = fmaf(a.x, a.x, q0);
                                    s1 += a.y; q1 = fmaf(a.y, a.y, q1);
                                    s2 += b.x; q2 = fmaf(b.x, b.x, q2);
                                    s3 += b.y; q3 = fmaf(b.y, b.y, q3);
                                }
                            }
                            float su = (s0 + s1) + (s2 + s3), sq = (q0 + q1) + (q2 + q3);
                            su += __shfl_xor_sync(0xffffffffu, su, 16);
                            sq += __shfl_xor_sync(0xffffffffu, sq, 16);
                            if (half == 0) {   // gn_part[64-column unit of the tile][band][quad][2]
                                float* part = gn_part + (((cc * S::UNITS + u) * 4 + band) * 16 + quad4) * 2;
                                part[0] = su;
                                part[1] = sq;
                            }
                        }
                    }
                }
                if (p.stats != nullptr) {
                    named_bar_sync(1, 128);
                    // thread = (domain-in-tile, local group, statistic): fold bands x quads of that group
                    const int cpg = p.stats_cpg;                 // multiple of 4
                    const int qpg = cpg >> 2;                    // quads per group
                    const int ndom = (p.stats_rows < GEMM_BM) ? 2 : 1;
                    const int bands_per_dom = 4 / ndom;
                    const int col_base = n_idx * BN;
                    const int g_first = col_base / cpg;
                    const int groups_tile = (col_base + BN - 1) / cpg - g_first + 1;   // groups intersecting this tile
                    const int dom_base = (org[1] * p.stats_mul[0] + org[2] * p.stats_mul[1] + org[3] * p.stats_mul[2] +
                                          org[4] * p.stats_mul[3]) / p.stats_div;
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
                                a += gn_part[((uq * 4 + dl * bands_per_dom + b) * 16 + qq) * 2 + st];
                        }
                        if (q_hi > q_lo)
                            atomicAdd(&p.stats[(static_cast<size_t>(dom_base + dl) * 32 + g) * 2 + st], static_cast<double>(a));
                    }
                    // (the next tile's named barriers order these reads before gn_part is rewritten)
                }
            }
        }
        if (leader) tma_store_wait_all0();
    }

    tc_fence_before();
    cluster_sync_all();   // the peer may still read this CTA's shared memory / signal its barriers until here
    if (warp == 1) {
        tc_fence_after();
        __syncwarp();
        tmem_dealloc2(tmem_base, S::TMEM_COLS);
    }
}

}  // namespace mmd
