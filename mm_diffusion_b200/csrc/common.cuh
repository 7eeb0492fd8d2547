// Shared device-side primitives for the sm_100a kernels of the MM-Diffusion
// denoising hot path: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (UMMA +
// TMEM) wrappers and the UMMA descriptor builders.  Hand-written inline PTX;
// nothing here comes from the reference (it has no native code).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmd {

// Activation / weight storage type.  The reference's production setting is
// fp16 convs + fp32 GroupNorm/softmax; we keep fp16 storage
// with fp32 accumulation everywhere.
using act_t = __half;

#define MMD_DEVINL __device__ __forceinline__

// Bounded wait: a mis-programmed pipeline must trap instead of hanging the box.
#ifndef MMD_WAIT_CYCLES
#define MMD_WAIT_CYCLES (6000000000LL)  // ~3 s at 2 GHz
#endif

MMD_DEVINL uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

MMD_DEVINL uint32_t lane_id() {
    uint32_t l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

MMD_DEVINL bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ----------------------------------------------------------------- mbarrier
MMD_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
MMD_DEVINL void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
MMD_DEVINL void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
MMD_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
MMD_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
MMD_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint
// expires) instead of spinning — a spinning single-thread role (TMA / MMA issuer) otherwise steals issue slots
// from the math warps that share its scheduler (measured: ~45 % of all executed instructions were poll loops).
MMD_DEVINL bool mbar_try_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
// MMD_WAIT_MODE (compile time): 0 = one plain try_wait, then try_wait with a 20 us suspend-time hint (ptxas lowers the
// hint to NANOSLEEP.SYNCS + PHASECHK); 1 = plain try_wait in a loop (the hardware-blocking form with its own short time
// limit, as CUTLASS waits); 2 = plain try_wait loop with a 32 ns nanosleep back-off between attempts.
#ifndef MMD_WAIT_MODE
#define MMD_WAIT_MODE 0
#endif
MMD_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
#if MMD_WAIT_MODE == 0
    while (!mbar_try_wait_sleep(bar, parity, 20000u)) {
#else
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
#if MMD_WAIT_MODE == 2
        __nanosleep(32);
#endif
        if ((++spins & 1023u) != 0) continue;
#endif
        if (clock64() - t0 > MMD_WAIT_CYCLES) {
            printf("mmd: mbarrier timeout block %d thread %d bar@%u parity %u\n", blockIdx.x, threadIdx.x, smem_u32(bar), parity);
            __trap();
        }
    }
}

// ---------------------------------------------------------------------- TMA
MMD_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

MMD_DEVINL void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
MMD_DEVINL void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
MMD_DEVINL void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
MMD_DEVINL void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
// rank-dispatched load; c[0] is the innermost (channel) coordinate.
MMD_DEVINL void tma_load_nd(int rank, void* dst, const CUtensorMap* m, uint64_t* bar, const int* c) {
    switch (rank) {
        case 2: tma_load_2d(dst, m, bar, c[0], c[1]); break;
        case 3: tma_load_3d(dst, m, bar, c[0], c[1], c[2]); break;
        case 4: tma_load_4d(dst, m, bar, c[0], c[1], c[2], c[3]); break;
        default: tma_load_5d(dst, m, bar, c[0], c[1], c[2], c[3], c[4]); break;
    }
}

// L2 prefetch of a tile (no shared-memory destination, no completion tracking)
MMD_DEVINL void tma_prefetch_nd(int rank, const CUtensorMap* m, const int* c) {
    const uint64_t mp = reinterpret_cast<uint64_t>(m);
    switch (rank) {
        case 2: asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(mp), "r"(c[0]), "r"(c[1]) : "memory"); break;
        case 3: asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(mp), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory"); break;
        case 4: asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(mp), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory"); break;
        default: asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];" ::"l"(mp), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory"); break;
    }
}

MMD_DEVINL void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
MMD_DEVINL void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
MMD_DEVINL void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
MMD_DEVINL void tma_store_5d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
MMD_DEVINL void tma_store_nd(int rank, const CUtensorMap* m, const void* src, const int* c) {
    switch (rank) {
        case 2: tma_store_2d(m, src, c[0], c[1]); break;
        case 3: tma_store_3d(m, src, c[0], c[1], c[2]); break;
        case 4: tma_store_4d(m, src, c[0], c[1], c[2], c[3]); break;
        default: tma_store_5d(m, src, c[0], c[1], c[2], c[3], c[4]); break;
    }
}
MMD_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
MMD_DEVINL void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
MMD_DEVINL void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
MMD_DEVINL void tma_store_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------ tcgen05
MMD_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
MMD_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Whole-warp call.  Writes the TMEM base address to *dst_smem.
MMD_DEVINL void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
MMD_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; single-thread issue.
MMD_DEVINL void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand from tensor memory (K-major, 16-bit values packed two per 32-bit column: 8 columns per K = 16 step), B from a
// shared-memory descriptor.
MMD_DEVINL void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives on the mbarrier when all previously issued MMAs of this thread retire.
MMD_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp gets lane
// (warp%4)*32+i, columns [col, col+32).
MMD_DEVINL void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
MMD_DEVINL void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
MMD_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, SWIZZLE_128B canonical layouts
// (bit layout: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
//  layout_type [61,64) with SWIZZLE_128B = 2).
MMD_DEVINL uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// kind::f16 instruction descriptor: fp32 accumulate, fp16 (fmt 0) or bf16 (fmt 1)
// operands, M x N tile, operand majors (0 = K-major, 1 = MN-major).
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, int a_mn_major, int b_mn_major, int fmt = 0) {
    return (1u << 4) | (static_cast<uint32_t>(fmt) << 7) | (static_cast<uint32_t>(fmt) << 10) |
           (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16) |
           (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// Byte offset of 16-byte chunk `c16` (0..7) of row `r` inside a [rows x 128 B]
// SWIZZLE_128B tile (tile base 1024-B aligned).
MMD_DEVINL uint32_t sw128_off(uint32_t r, uint32_t c16) { return r * 128u + ((c16 ^ (r & 7u)) << 4); }

// Programmatic dependent launch: the next kernel of the stream may start its prologue while this one drains;
// every kernel that is launched with the PDL attribute calls pdl_wait() before its first global access.
// (Both are no-ops for a kernel launched without the attribute.)
MMD_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
MMD_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

MMD_DEVINL void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

MMD_DEVINL float silu_f(float x) { return x / (1.0f + __expf(-x)); }

// ------------------------------------------------------------------ dropout
// Counter-based Philox4x32-10 (Salmon et al., SC'11): 128 random bits per (key, counter).  The training path uses it
// for nn.Dropout (multimodal_unet.py:376,384): the mask of an element depends only on (seed, site, element index), so
// the backward regenerates it instead of storing it.
MMD_DEVINL uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
// Dropout parameters of one training forward, in device memory (the captured graphs read them at replay):
// {seed lo, seed hi, drop threshold on 16-bit uniforms (0 = dropout off), float bits of 1 / (1 - p)}.
struct DropState { uint32_t seed_lo, seed_hi, thresh16, scale_bits; };
// Keep mask (bit i = keep channel c0 + i) of the 8 consecutive elements starting at element index 8 * idx8 of dropout
// site `site`: element kept iff its 16-bit uniform >= thresh16, i.e. P(drop) = thresh16 / 65536.
MMD_DEVINL uint32_t dropout_keep8(const DropState& d, uint32_t site, unsigned long long idx8) {
    const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(idx8), static_cast<uint32_t>(idx8 >> 32), site, 0x6d6d64u),
                                  make_uint2(d.seed_lo, d.seed_hi));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    uint32_t keep = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        keep |= ((w[i] & 0xFFFFu) >= d.thresh16 ? 1u : 0u) << (2 * i);
        keep |= ((w[i] >> 16) >= d.thresh16 ? 1u : 0u) << (2 * i + 1);
    }
    return keep;
}

}  // namespace mmd
