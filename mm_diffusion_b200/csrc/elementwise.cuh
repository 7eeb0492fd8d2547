// HBM-bound helper kernels of the denoising step (channels-last fp16 activations):
// GroupNorm statistics / apply (+SiLU, +FiLM), pooling / nearest upsampling,
// im2col of the 3- and 1-channel network inputs, timestep-embedding MLPs,
// the 16-token temporal self-attention core and the fused p_sample tail.
// Each kernel cites the reference lines whose arithmetic it restates.
#pragma once
#include "common.cuh"

namespace mmd {

// ---------------------------------------------------------------------------
// GroupNorm statistics.  Reference: nn.py:16-33 (GroupNorm32: 32 groups, eps 1e-5,
// statistics in fp32 over (C/32 channels x all positions of the domain)).
// Input is the channel-concatenation of up to two row-major sources
// (U-Net skip concat, multimodal_unet.py:1093-1094).  A "domain" is R consecutive
// rows; sums go to double accumulators [NS][32][2].
// ---------------------------------------------------------------------------
struct GnSrc {
    const act_t* x1; int c1; int ld1;
    const act_t* x2; int c2; int ld2;
};

// SiLU through one MUFU op: x * sigmoid(x) = 0.5 x (1 + tanh(x/2)); tanh.approx.f32 abs error ~5e-4, i.e. below
// the fp16 rounding of the stored result.
MMD_DEVINL float silu_fast(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return 0.5f * x * (1.0f + t);
}

constexpr int GN_UNROLL = 4;

__global__ void __launch_bounds__(256) gn_stats_kernel(GnSrc s, int R, int rows_per_block, double* __restrict__ sums) {
    pdl_trigger();
    pdl_wait();
    const int C = s.c1 + s.c2;
    const int cpg = C / 32;
    const int vpr = C / 8;  // 16-byte vectors per row
    const int ns = blockIdx.y;
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(R, r0 + rows_per_block);
    const int rows_per_pass = blockDim.x / vpr;
    const int vec = threadIdx.x % vpr;
    const int rsub = threadIdx.x / vpr;
    // double accumulators in shared memory too: the order of the atomic additions then only moves the sums at the 1e-16
    // level, far below the fp32 mean / rstd they turn into, so the forward is bit-reproducible in practice
    __shared__ double sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = 0.0;
    __syncthreads();
    if (rsub < rows_per_pass) {
        float sm[8], sq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { sm[i] = 0.f; sq[i] = 0.f; }
        const int c0 = vec * 8;
        const act_t* base;
        int ld;
        if (c0 < s.c1) { base = s.x1 + c0; ld = s.ld1; } else { base = s.x2 + (c0 - s.c1); ld = s.ld2; }
        base += static_cast<size_t>(ns) * R * ld;
        for (int r = r0 + rsub; r < r1; r += GN_UNROLL * rows_per_pass) {
            uint4 raw[GN_UNROLL];
#pragma unroll
            for (int u = 0; u < GN_UNROLL; ++u) {
                const int rr = r + u * rows_per_pass;
                raw[u] = (rr < r1) ? __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(rr) * ld)) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < GN_UNROLL; ++u) {
                const __half2* h = reinterpret_cast<const __half2*>(&raw[u]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(h[i]);
                    sm[2 * i] += f.x; sq[2 * i] = fmaf(f.x, f.x, sq[2 * i]);
                    sm[2 * i + 1] += f.y; sq[2 * i + 1] = fmaf(f.y, f.y, sq[2 * i + 1]);
                }
            }
        }
        // fold the 8 channels into their (<= 4) groups before touching shared memory
        int g = c0 / cpg;
        float as = 0.f, aq = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int gi = (c0 + i) / cpg;
            if (gi != g) {
                atomicAdd(&sh[2 * g], static_cast<double>(as));
                atomicAdd(&sh[2 * g + 1], static_cast<double>(aq));
                g = gi; as = 0.f; aq = 0.f;
            }
            as += sm[i]; aq += sq[i];
        }
        atomicAdd(&sh[2 * g], static_cast<double>(as));
        atomicAdd(&sh[2 * g + 1], static_cast<double>(aq));
    }
    __syncthreads();
    if (threadIdx.x < 64) atomicAdd(&sums[static_cast<size_t>(ns) * 64 + threadIdx.x], sh[threadIdx.x]);
}

// ---------------------------------------------------------------------------
// GroupNorm statistics of a channel concat [x1 (c1) | x2 (c2)] from the per-4-channel partial sums the producers of x1 and
// x2 left behind (GemmParams::stats_q): sums[d][g][2] = sum over the nsub slots of domain d and the quads of group g.
// Replaces a full statistics pass over both tensors (the decoder's skip-concat norms, multimodal_unet.py:1093-1094 + :338).
// ---------------------------------------------------------------------------
__global__ void gn_fold_quads_kernel(const double* __restrict__ q1, int c1, const double* __restrict__ q2, int c2, int nsub,
                                     double* __restrict__ sums) {
    pdl_trigger();
    pdl_wait();
    const int d = blockIdx.x;
    const int g = threadIdx.x >> 1, st = threadIdx.x & 1;   // 64 threads
    const int n1 = c1 / 4, n2 = c2 / 4;
    const int qpg = (n1 + n2) / 32;
    double acc = 0.0;
    for (int j = g * qpg; j < (g + 1) * qpg; ++j) {
        const double* src = j < n1 ? q1 + (static_cast<size_t>(d) * nsub * n1 + j) * 2 : q2 + (static_cast<size_t>(d) * nsub * n2 + (j - n1)) * 2;
        const size_t stride = static_cast<size_t>(j < n1 ? n1 : n2) * 2;
        for (int k = 0; k < nsub; ++k) acc += src[k * stride + st];
    }
    sums[(static_cast<size_t>(d) * 32 + g) * 2 + st] = acc;
}

// ---------------------------------------------------------------------------
// GroupNorm apply: y = act( gn(x) * (1 + scale) + shift ), act = SiLU or identity.
// Reference: nn.py:22-33; multimodal_unet.py:338-347 (in_layers: GN -> SiLU),
// :459-470 (out_layers: GN * (1+scale) + shift -> SiLU), :284/:664 (attention norm, no SiLU).
// film points at [B][film_ld] floats with scale at [0,C) and shift at [C,2C)
// (th.chunk order, multimodal_unet.py:462,468).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_apply_kernel(GnSrc s, int R, int rows_per_block, const double* __restrict__ sums,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ film, int film_ld, int ns_per_batch, int do_silu,
                                act_t* __restrict__ y, int nsub, long long stat_rows,
                                const DropState* __restrict__ drop, uint32_t drop_site) {
    pdl_trigger();
    // gamma / beta are constants of the plan: requested before the grid dependency resolves, so their latency overlaps the
    // previous kernel's tail (the statistics and the FiLM vector are produced inside the step and are read after the wait)
    const int C_pre = s.c1 + s.c2;
    float g_pre[4], b_pre[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = threadIdx.x + i * 256;
        g_pre[i] = c < C_pre ? __ldg(gamma + c) : 0.f;
        b_pre[i] = c < C_pre ? __ldg(beta + c) : 0.f;
    }
    pdl_wait();
    // nn.Dropout after the SiLU of the ResBlock out_layers (multimodal_unet.py:376,384), training forwards only
    DropState ds{0u, 0u, 0u, 0u};
    if (drop != nullptr) ds = *drop;
    const bool dropping = ds.thresh16 != 0u;
    const float drop_scale = __uint_as_float(ds.scale_bits);
    extern __shared__ float coef[];  // [2][C] then [64] group mean / rstd
    const int C = s.c1 + s.c2;
    const int cpg = C / 32;
    const int vpr = C / 8;
    const int ns = blockIdx.y;
    float* gstat = coef + 2 * C;
    // Everything the block needs from global memory is requested up front, so the three dependent latencies of the
    // straightforward order (statistics -> FiLM vector -> first rows) overlap: small tensors are latency bound here.
    const int rows_per_pass = blockDim.x / vpr;
    const int vec = threadIdx.x % vpr;
    const int rsub = threadIdx.x / vpr;
    const bool streams = rsub < rows_per_pass;
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(R, r0 + rows_per_block);
    const int c0 = vec * 8;
    const act_t* base;
    int ld;
    if (c0 < s.c1) { base = s.x1 + c0; ld = s.ld1; } else { base = s.x2 + (c0 - s.c1); ld = s.ld2; }
    base += static_cast<size_t>(ns) * R * ld;
    int r = r0 + rsub;
    uint4 raw[GN_UNROLL];
    bool have = false;
    if (streams && r < r1) {
#pragma unroll
        for (int u = 0; u < GN_UNROLL; ++u) {
            const int rr = r + u * rows_per_pass;
            if (rr < r1) raw[u] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(rr) * ld));
        }
        have = true;
    }
    float sc_pre[4], sh_pre[4];
    if (film != nullptr) {
        const float* fb = film + static_cast<size_t>(ns / ns_per_batch) * film_ld;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = threadIdx.x + i * 256;
            sc_pre[i] = c < C ? 1.f + fb[c] : 1.f;
            sh_pre[i] = c < C ? fb[C + c] : 0.f;
        }
    }
    if (threadIdx.x < 32) {  // the only double-precision arithmetic: 32 groups (x nsub partial slots)
        const double inv_n = 1.0 / (static_cast<double>(stat_rows) * cpg);
        double su = 0.0, sq = 0.0;
        for (int k = 0; k < nsub; ++k) {
            su += sums[(static_cast<size_t>(ns) * nsub + k) * 64 + 2 * threadIdx.x];
            sq += sums[(static_cast<size_t>(ns) * nsub + k) * 64 + 2 * threadIdx.x + 1];
        }
        const double mean = su * inv_n;
        double var = sq * inv_n - mean * mean;
        if (var < 0) var = 0;
        gstat[2 * threadIdx.x] = static_cast<float>(mean);
        gstat[2 * threadIdx.x + 1] = rsqrtf(static_cast<float>(var) + 1e-5f);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = threadIdx.x + i * 256;
        if (c < C) {
            const int g = c / cpg;
            float a = gstat[2 * g + 1] * g_pre[i];
            float b = b_pre[i] - gstat[2 * g] * a;
            if (film != nullptr) {
                a *= sc_pre[i];
                b = b * sc_pre[i] + sh_pre[i];
            }
            coef[c] = a;
            coef[C + c] = b;
        }
    }
    for (int c = threadIdx.x + 4 * 256; c < C; c += 256) {   // wider than 1024 channels (not in this network)
        const int g = c / cpg;
        float a = gstat[2 * g + 1] * gamma[c];
        float b = beta[c] - gstat[2 * g] * a;
        if (film != nullptr) {
            const float* fb = film + static_cast<size_t>(ns / ns_per_batch) * film_ld;
            const float sc = 1.f + fb[c];
            a *= sc;
            b = b * sc + fb[C + c];
        }
        coef[c] = a;
        coef[C + c] = b;
    }
    __syncthreads();
    if (!streams) return;
    act_t* ybase = y + static_cast<size_t>(ns) * R * C + c0;
    float ca[8], cb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ca[i] = coef[c0 + i]; cb[i] = coef[C + c0 + i]; }
    for (; r < r1; r += GN_UNROLL * rows_per_pass) {
        if (!have) {
#pragma unroll
            for (int u = 0; u < GN_UNROLL; ++u) {
                const int rr = r + u * rows_per_pass;
                if (rr < r1) raw[u] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(rr) * ld));
            }
        }
        have = false;
#pragma unroll
        for (int u = 0; u < GN_UNROLL; ++u) {
            const int rr = r + u * rows_per_pass;
            if (rr >= r1) break;
            const __half2* h = reinterpret_cast<const __half2*>(&raw[u]);
            uint4 outv;
            __half2* o = reinterpret_cast<__half2*>(&outv);
            uint32_t keep = 0xFFu;
            if (dropping)
                keep = dropout_keep8(ds, drop_site, ((static_cast<unsigned long long>(ns) * R + rr) * C + c0) >> 3);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = __half22float2(h[k]);
                float u0 = fmaf(f.x, ca[2 * k], cb[2 * k]);
                float u1 = fmaf(f.y, ca[2 * k + 1], cb[2 * k + 1]);
                if (do_silu) { u0 = silu_fast(u0); u1 = silu_fast(u1); }
                if (dropping) {
                    u0 = ((keep >> (2 * k)) & 1u) ? u0 * drop_scale : 0.f;
                    u1 = ((keep >> (2 * k + 1)) & 1u) ? u1 * drop_scale : 0.f;
                }
                o[k] = __floats2half2_rn(u0, u1);
            }
            *reinterpret_cast<uint4*>(ybase + static_cast<size_t>(rr) * C) = outv;
        }
    }
}

// ---------------------------------------------------------------------------
// GroupNorm whose domain is one pixel across the F frames (temporal attention norm:
// input rearranged "(b h w) c f", multimodal_unet.py:489-490 -> stats over C/32 x F).
// x, y: [B][F][P][C].  One thread per (pixel, group).
// ---------------------------------------------------------------------------
// Register-resident: a thread loads its whole (F x C/32) domain with wide independent loads (all in flight
// at once), reduces and normalises from registers.  Consecutive threads own consecutive groups of one pixel,
// so a warp reads / writes complete channel rows.
template <int HALVES> struct GnVec;
template <> struct GnVec<8> { using type = uint4; };
template <> struct GnVec<4> { using type = uint2; };
template <> struct GnVec<2> { using type = uint32_t; };

template <int CPG, int F_>
__global__ void __launch_bounds__(128) gn_temporal_kernel(const act_t* __restrict__ x, act_t* __restrict__ y,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          int B, int P, int C) {
    pdl_trigger();
    pdl_wait();
    constexpr int VEC = (CPG % 8 == 0) ? 8 : ((CPG % 4 == 0) ? 4 : 2);
    constexpr int NV = CPG / VEC;
    using V = typename GnVec<VEC>::type;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(B) * P * 32;
    if (idx >= total) return;
    const int g = static_cast<int>(idx % 32);
    const long long bp = idx / 32;
    const int p = static_cast<int>(bp % P);
    const int b = static_cast<int>(bp / P);
    const size_t fstride = static_cast<size_t>(P) * C;
    const size_t base = (static_cast<size_t>(b) * F_ * P + p) * C + g * CPG;
    V v[F_][NV];
#pragma unroll
    for (int f = 0; f < F_; ++f)
#pragma unroll
        for (int k = 0; k < NV; ++k) v[f][k] = __ldg(reinterpret_cast<const V*>(x + base + f * fstride) + k);
    float su = 0.f, ss = 0.f;
#pragma unroll
    for (int f = 0; f < F_; ++f)
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const __half2* h = reinterpret_cast<const __half2*>(&v[f][k]);
#pragma unroll
            for (int i = 0; i < VEC / 2; ++i) {
                const float2 t = __half22float2(h[i]);
                su += t.x + t.y;
                ss = fmaf(t.x, t.x, fmaf(t.y, t.y, ss));
            }
        }
    const float inv_n = 1.f / (F_ * CPG);
    const float mean = su * inv_n;
    const float var = fmaxf(ss * inv_n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + 1e-5f);
    float ca[CPG], cb[CPG];
#pragma unroll
    for (int i = 0; i < CPG; ++i) {
        ca[i] = rstd * gamma[g * CPG + i];
        cb[i] = beta[g * CPG + i] - mean * ca[i];
    }
#pragma unroll
    for (int f = 0; f < F_; ++f)
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            V o;
            const __half2* h = reinterpret_cast<const __half2*>(&v[f][k]);
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int i = 0; i < VEC / 2; ++i) {
                const float2 t = __half22float2(h[i]);
                const int c = k * VEC + 2 * i;
                oh[i] = __floats2half2_rn(fmaf(t.x, ca[c], cb[c]), fmaf(t.y, ca[c + 1], cb[c + 1]));
            }
            *(reinterpret_cast<V*>(y + base + f * fstride) + k) = o;
        }
}

// ---------------------------------------------------------------------------
// Resampling (multimodal_unet.py:133-208): AvgPool3d (1,2,2) / AvgPool1d(4) and
// nearest x(1,2,2) / x4, channels-last.  mode 0: video pool, 1: audio pool,
// 2: video up, 3: audio up.  Thread = one 8-channel vector of one output token.
// ---------------------------------------------------------------------------
__global__ void resample_kernel(const act_t* __restrict__ x, act_t* __restrict__ y, int mode, int N, int H, int W,
                                int C) {
    pdl_trigger();
    pdl_wait();
    // video: x [N][H][W][C]; audio: x [N][H(=L)][C] with W unused.  One 16-byte vector of the output per thread; the index
    // math is 32-bit unsigned (the launcher rejects tensors of 2^31 vectors or more): 64-bit divisions by run-time values
    // cost more instructions than the data movement itself.
    const uint32_t vpr = static_cast<uint32_t>(C) / 8;
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t total;
    if (mode == 0) total = static_cast<uint32_t>(N) * (H / 2) * (W / 2) * vpr;
    else if (mode == 1) total = static_cast<uint32_t>(N) * (H / 4) * vpr;
    else if (mode == 2) total = static_cast<uint32_t>(N) * (H * 2) * (W * 2) * vpr;
    else total = static_cast<uint32_t>(N) * (H * 4) * vpr;
    if (idx >= total) return;
    const uint32_t v = idx % vpr;
    uint32_t t = idx / vpr;
    if (mode == 0 || mode == 1) {
        const uint4* src[4];
        if (mode == 0) {
            const uint32_t w2 = W / 2, h2 = H / 2;
            const uint32_t ow = t % w2; t /= w2;
            const uint32_t oh = t % h2;
            const uint32_t n = t / h2;
            const act_t* b0 = x + ((static_cast<size_t>(n) * H + 2 * oh) * W + 2 * ow) * C + v * 8;
            src[0] = reinterpret_cast<const uint4*>(b0);
            src[1] = reinterpret_cast<const uint4*>(b0 + C);
            src[2] = reinterpret_cast<const uint4*>(b0 + static_cast<size_t>(W) * C);
            src[3] = reinterpret_cast<const uint4*>(b0 + static_cast<size_t>(W) * C + C);
        } else {
            const uint32_t l4 = H / 4;
            const uint32_t ol = t % l4;
            const uint32_t n = t / l4;
            const act_t* b0 = x + (static_cast<size_t>(n) * H + 4 * ol) * C + v * 8;
            for (int k = 0; k < 4; ++k) src[k] = reinterpret_cast<const uint4*>(b0 + static_cast<size_t>(k) * C);
        }
        uint4 raw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) raw[k] = __ldg(src[k]);
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __half2* h = reinterpret_cast<const __half2*>(&raw[k]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(h[i]);
                acc[2 * i] += f.x;
                acc[2 * i + 1] += f.y;
            }
        }
        uint4 o;
        __half2* oh2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) oh2[i] = __floats2half2_rn(acc[2 * i] * 0.25f, acc[2 * i + 1] * 0.25f);
        reinterpret_cast<uint4*>(y)[idx] = o;
    } else {
        const act_t* b0;
        if (mode == 2) {
            const uint32_t w2 = W * 2, h2 = H * 2;
            const uint32_t ow = t % w2; t /= w2;
            const uint32_t oh = t % h2;
            const uint32_t n = t / h2;
            b0 = x + ((static_cast<size_t>(n) * H + oh / 2) * W + ow / 2) * C + v * 8;
        } else {
            const uint32_t l4 = H * 4;
            const uint32_t ol = t % l4;
            const uint32_t n = t / l4;
            b0 = x + (static_cast<size_t>(n) * H + ol / 4) * C + v * 8;
        }
        reinterpret_cast<uint4*>(y)[idx] = __ldg(reinterpret_cast<const uint4*>(b0));
    }
}

// ---------------------------------------------------------------------------
// Narrow 3x3x3 video head as "pointwise GEMM + gather" (inference plans).  The implicit-GEMM form re-reads every
// activation row for 27 taps (1.8 GB through L2 for 3 output channels: 206 us per step at B = 4); instead ONE pointwise GEMM
// computes the 27 x Co per-tap partial products of every token, Y[tok][tap * 4 + c] = sum_k W[c][k][tap] * A[tok][k]
// (27 * 4 = 108 -> 128 columns), and this kernel sums the 27 neighbours: out[pos][c] = b[c] + sum_tap Y[pos + d(tap)][tap][c]
// with "same" zero padding.  Reference: multimodal_unet.py:1003-1007 (video_out), :68-106 (VideoConv, conv_type '3d').
// ---------------------------------------------------------------------------
__global__ void pack_head_taps_kernel(const float* __restrict__ w, act_t* __restrict__ dst, int Co, int Ci, int T) {
    // w [Co][Ci][T] fp32 -> dst [128][Ci] fp16, row = tap * 4 + c (rows past T * 4 and c >= Co are zero)
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 128 * Ci) return;
    const int row = idx / Ci, k = idx - row * Ci;
    const int tap = row >> 2, c = row & 3;
    const float v = (tap < T && c < Co) ? w[(static_cast<size_t>(c) * Ci + k) * T + tap] : 0.f;
    dst[idx] = __float2half(v);
}

__global__ void __launch_bounds__(256) head_gather3d_kernel(const act_t* __restrict__ y, const float* __restrict__ bias,
                                                            float* __restrict__ out, int B, int F, int H, int W, int Co) {
    pdl_trigger();
    pdl_wait();
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = static_cast<uint32_t>(B) * F * H * W;
    if (idx >= total) return;
    const int w0 = idx % W;
    uint32_t t = idx / W;
    const int h0 = t % H; t /= H;
    const int f0 = t % F;
    const int b0 = t / F;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
        const int ff = f0 + kt - 1;
        if (ff < 0 || ff >= F) continue;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int hh = h0 + ky - 1;
            if (hh < 0 || hh >= H) continue;
            const act_t* rowp = y + ((static_cast<size_t>(b0) * F + ff) * H + hh) * static_cast<size_t>(W) * 128;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ww = w0 + kx - 1;
                if (ww < 0 || ww >= W) continue;
                const int tap = (kt * 3 + ky) * 3 + kx;
                const uint2 raw = __ldg(reinterpret_cast<const uint2*>(rowp + static_cast<size_t>(ww) * 128 + tap * 4));
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
                const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
                acc[0] += a.x; acc[1] += a.y; acc[2] += c.x; acc[3] += c.y;
            }
        }
    }
    const size_t hw = static_cast<size_t>(H) * W;
    float* o = out + (static_cast<size_t>(b0) * F + f0) * Co * hw + static_cast<size_t>(h0) * W + w0;
    for (int c = 0; c < Co; ++c) o[c * hw] = acc[c] + __ldg(bias + c);
}

// ---------------------------------------------------------------------------
// im2col of the raw network inputs (InitialBlock, multimodal_unet.py:680-694):
// video fp32 [BF][3][H][W] -> A [BF*H*W][64] with k = (ky*3+kx)*3 + c, zero padded;
// audio fp32 [B][1][L]     -> A [B*L][64]    with k = tap.
// ---------------------------------------------------------------------------
__global__ void im2col_video_kernel(const float* __restrict__ x, act_t* __restrict__ a, int BF, int Cin, int H, int W) {
    pdl_trigger();
    pdl_wait();
    // k -> (dy, dx, c) table once per block (the per-element divisions by run-time values cost more than the loads);
    // 32-bit indices (the launcher's tensors are far below 2^31 elements)
    __shared__ int tab[64];   // (dy + 1) | (dx + 1) << 2 | c << 4, or -1 past the 9 * Cin real columns
    if (threadIdx.x < 64) {
        const int k = threadIdx.x;
        int e = -1;
        if (k < 9 * Cin) {
            const int tap = k / Cin, c = k - tap * Cin;
            e = (tap / 3) | ((tap % 3) << 2) | (c << 4);
        }
        tab[k] = e;
    }
    __syncthreads();
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = static_cast<uint32_t>(BF) * H * W * 8;  // 8 vectors of 8 per token
    if (idx >= total) return;
    const int v = static_cast<int>(idx & 7);
    uint32_t t = idx >> 3;
    const int w = static_cast<int>(t % W); t /= W;
    const int h = static_cast<int>(t % H);
    const uint32_t n = t / H;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (v * 8 < 9 * Cin) {   // vectors past the real columns are all padding
        __half* oh = reinterpret_cast<__half*>(&o);
        const float* xn = x + static_cast<size_t>(n) * Cin * H * W;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tab[v * 8 + i];
            float val = 0.f;
            if (e >= 0) {
                const int yy = h + (e & 3) - 1, xx = w + ((e >> 2) & 3) - 1, c = e >> 4;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) val = __ldg(xn + (static_cast<size_t>(c) * H + yy) * W + xx);
            }
            oh[i] = __float2half_rn(val);
        }
    }
    reinterpret_cast<uint4*>(a)[idx] = o;
}

__global__ void im2col_audio_kernel(const float* __restrict__ x, act_t* __restrict__ a, int B, int Cin, int L) {
    pdl_trigger();
    pdl_wait();
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(B) * L * 8;
    if (idx >= total) return;
    const int v = static_cast<int>(idx & 7);
    long long t = idx >> 3;
    const int l = static_cast<int>(t % L);
    const long long n = t / L;
    uint4 o;
    __half* oh = reinterpret_cast<__half*>(&o);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = v * 8 + i;
        float val = 0.f;
        if (k < 3 * Cin) {
            const int tap = k / Cin, c = k % Cin;
            const int ll = l + tap - 1;
            if (ll >= 0 && ll < L) val = x[(n * Cin + c) * L + ll];
        }
        oh[i] = __float2half_rn(val);
    }
    reinterpret_cast<uint4*>(a)[idx] = o;
}

// ---------------------------------------------------------------------------
// Timestep embedding + time_embed MLP (nn.py:192-210, multimodal_unet.py:791-795,1075)
// emb[b] = W2 * silu(W1 * [cos(t f_i), sin(t f_i)] + b1) + b2, all fp32; then stores silu(emb)
// because every consumer is emb_layers = Linear(SiLU(emb)) (multimodal_unet.py:366-372).
// One block per sample, blockDim = dim (128).
// ---------------------------------------------------------------------------
__global__ void time_embed_kernel(const float* __restrict__ t, const float* __restrict__ w1, const float* __restrict__ b1,
                                  const float* __restrict__ w2, const float* __restrict__ b2, int dim,
                                  float* __restrict__ emb_out, float* __restrict__ silu_emb_out) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sh[];  // [2][dim]
    float* e0 = sh;
    float* e1 = sh + dim;
    const int b = blockIdx.x;
    const int i = threadIdx.x;
    const int half = dim / 2;
    const float tv = t[b];
    if (i < dim) {
        const int j = (i < half) ? i : i - half;
        const float freq = expf(-logf(10000.0f) * static_cast<float>(j) / static_cast<float>(half));
        const float arg = tv * freq;
        e0[i] = (i < half) ? cosf(arg) : sinf(arg);
    }
    __syncthreads();
    if (i < dim) {
        float acc = b1[i];
        for (int k = 0; k < dim; ++k) acc += w1[i * dim + k] * e0[k];
        e1[i] = silu_f(acc);
    }
    __syncthreads();
    if (i < dim) {
        float acc = b2[i];
        for (int k = 0; k < dim; ++k) acc += w2[i * dim + k] * e1[k];
        emb_out[b * dim + i] = acc;
        silu_emb_out[b * dim + i] = silu_f(acc);
    }
}

// All ResBlock emb_layers at once: out[b][j] = bias[j] + sum_k W[j][k] * silu_emb[b][k]
// (rows of all 28 emb_layers.1 Linear layers stacked).  One warp per row j.
__global__ void emb_layers_kernel(const float* __restrict__ silu_emb, const float* __restrict__ w,
                                  const float* __restrict__ bias, int B, int dim, int rows, float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= rows) return;
    for (int b = 0; b < B; ++b) {
        float acc = 0.f;
        for (int k = lane; k < dim; k += 32) acc += w[static_cast<size_t>(j) * dim + k] * silu_emb[b * dim + k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[static_cast<size_t>(b) * rows + j] = acc + bias[j];
    }
}

// ---------------------------------------------------------------------------
// Temporal self-attention core: sequences of F (8 or 16) tokens per (sample, pixel), heads of width
// d = C/heads (SingleModalQKVAttention, multimodal_unet.py:221-240; fp32 softmax).
// qkv: [B][F][P][3C] (q | k | v on channels, head h = channels [h*d,(h+1)*d)), out: [B][F][P][C].
// Bandwidth-bound (8 FLOP/B): one warp per (b, pixel, head); rows are staged with
// cp.async, the two 16x16 products run on mma.sync.m16n8k16 (legacy tensor path is plenty here:
// the kernel moves ~4C*2 bytes per token and does 64*C FLOPs per token).
// ---------------------------------------------------------------------------
MMD_DEVINL void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
MMD_DEVINL void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
MMD_DEVINL void mma_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
MMD_DEVINL void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
MMD_DEVINL uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int TATT_WARPS = 4;

template <int F_>
__global__ void __launch_bounds__(TATT_WARPS * 32) temporal_attn_kernel(const act_t* __restrict__ qkv, act_t* __restrict__ out,
                                                                         int B, int P, int C, int heads) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) uint8_t tsm[];
    const int d = C / heads;
    const int pitch = d + 8;  // halves; 16-byte aligned rows, conflict-free ldmatrix
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    act_t* sq = reinterpret_cast<act_t*>(tsm) + static_cast<size_t>(warp) * 3 * 16 * pitch;
    act_t* sk = sq + 16 * pitch;
    act_t* sv = sk + 16 * pitch;
    if (F_ < 16) {  // rows F_..15 of the 16-row tiles stay zero
        for (int i = lane; i < 3 * 16 * pitch / 8; i += 32) reinterpret_cast<uint4*>(sq)[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
    }
    const int vpr = d / 8;
    const long long total = static_cast<long long>(B) * P * heads;
    const float scale_log2 = 1.4426950408889634f * rsqrtf(static_cast<float>(d));
    const uint32_t sq_a = smem_u32(sq), sk_a = smem_u32(sk), sv_a = smem_u32(sv);
    for (long long item = static_cast<long long>(blockIdx.x) * TATT_WARPS + warp; item < total;
         item += static_cast<long long>(gridDim.x) * TATT_WARPS) {
        const int h = static_cast<int>(item % heads);
        const long long bp = item / heads;
        const int p = static_cast<int>(bp % P);
        const int b = static_cast<int>(bp / P);
        for (int i = lane; i < 3 * F_ * vpr; i += 32) {
            const int m = i / (F_ * vpr);
            const int rem = i - m * (F_ * vpr);
            const int f = rem / vpr, v = rem - f * vpr;
            const act_t* src = qkv + ((static_cast<size_t>(b) * F_ + f) * P + p) * (3 * static_cast<size_t>(C)) + m * C + h * d + v * 8;
            cp_async16(sq_a + ((m * 16 + f) * pitch + v * 8) * 2, src);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        // ---- S = Q K^T (16 x 16, two n-tiles of 8 keys)
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
        const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;
        const int lcol = (lane >> 4) * 8;
        const int krow = (lane & 7) + (lane >> 4) * 8;
        const int kcol = ((lane >> 3) & 1) * 8;
        for (int k0 = 0; k0 < d; k0 += 16) {
            uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
            ldmatrix_x4(sq_a + (lrow * pitch + k0 + lcol) * 2, a0, a1, a2, a3);
            ldmatrix_x4(sk_a + (krow * pitch + k0 + kcol) * 2, b0, b1, b2, b3);
            mma_16816(s0, a0, a1, a2, a3, b0, b1);
            mma_16816(s1, a0, a1, a2, a3, b2, b3);
        }
        // ---- softmax over the 16 keys of rows (lane/4) and (lane/4 + 8)
        const int c0 = (lane & 3) * 2;
        float pr[2][4];  // [n-tile][c0..c3]
#pragma unroll
        for (int j = 0; j < 4; ++j) { pr[0][j] = s0[j]; pr[1][j] = s1[j]; }
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (t * 8 + c0 + (j & 1) >= F_) pr[t][j] = -INFINITY;
        float mx0 = fmaxf(fmaxf(pr[0][0], pr[0][1]), fmaxf(pr[1][0], pr[1][1]));
        float mx1 = fmaxf(fmaxf(pr[0][2], pr[0][3]), fmaxf(pr[1][2], pr[1][3]));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            pr[t][0] = exp2f((pr[t][0] - mx0) * scale_log2); pr[t][1] = exp2f((pr[t][1] - mx0) * scale_log2);
            pr[t][2] = exp2f((pr[t][2] - mx1) * scale_log2); pr[t][3] = exp2f((pr[t][3] - mx1) * scale_log2);
            sum0 += pr[t][0] + pr[t][1];
            sum1 += pr[t][2] + pr[t][3];
        }
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
        const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
        const uint32_t pa0 = pack_h2(pr[0][0] * inv0, pr[0][1] * inv0);
        const uint32_t pa1 = pack_h2(pr[0][2] * inv1, pr[0][3] * inv1);
        const uint32_t pa2 = pack_h2(pr[1][0] * inv0, pr[1][1] * inv0);
        const uint32_t pa3 = pack_h2(pr[1][2] * inv1, pr[1][3] * inv1);
        // ---- O = P V, 16 output channels per iteration; staged into the (now free) Q tile
        __syncwarp();
        const int r = lane >> 2;
        for (int n0 = 0; n0 < d; n0 += 16) {
            uint32_t v0, v1, v2, v3;
            ldmatrix_x4_trans(sv_a + (lrow * pitch + n0 + lcol) * 2, v0, v1, v2, v3);
            float oa[4] = {0.f, 0.f, 0.f, 0.f}, ob[4] = {0.f, 0.f, 0.f, 0.f};
            mma_16816(oa, pa0, pa1, pa2, pa3, v0, v1);
            mma_16816(ob, pa0, pa1, pa2, pa3, v2, v3);
            *reinterpret_cast<uint32_t*>(sq + r * pitch + n0 + c0) = pack_h2(oa[0], oa[1]);
            *reinterpret_cast<uint32_t*>(sq + (r + 8) * pitch + n0 + c0) = pack_h2(oa[2], oa[3]);
            *reinterpret_cast<uint32_t*>(sq + r * pitch + n0 + 8 + c0) = pack_h2(ob[0], ob[1]);
            *reinterpret_cast<uint32_t*>(sq + (r + 8) * pitch + n0 + 8 + c0) = pack_h2(ob[2], ob[3]);
        }
        __syncwarp();
        for (int i = lane; i < F_ * vpr; i += 32) {
            const int f = i / vpr, v = i - f * vpr;
            *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(b) * F_ + f) * P + p) * C + h * d + v * 8) =
                *reinterpret_cast<const uint4*>(sq + f * pitch + v * 8);
        }
        __syncwarp();
        if (F_ < 16) {  // restore the zero rows of Q clobbered by the staging writes
            for (int i = lane; i < (16 - F_) * pitch / 8; i += 32)
                reinterpret_cast<uint4*>(sq + F_ * pitch)[i] = make_uint4(0, 0, 0, 0);
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------
// Fused p_sample tail (multimodal_gaussian_diffusion.py:345-350, 215-218, 453-470):
//   x0 = clamp(a_t x - b_t eps, -1, 1) (clamp optional); mean = c1_t x0 + c2_t x;
//   sample = mean + nz_t * sigma_t * z.  coef: [B][6] = {a, b, c1, c2, sigma, nonzero}.
// ---------------------------------------------------------------------------
__global__ void p_sample_tail_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                     const float* __restrict__ z, const float* __restrict__ coef, long long per_sample,
                                     long long total, int clip, float* __restrict__ sample,
                                     float* __restrict__ pred_xstart) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float* c = coef + (i / per_sample) * 6;
        const float xv = x[i];
        float x0 = c[0] * xv - c[1] * eps[i];
        if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
        const float mean = c[2] * x0 + c[3] * xv;
        sample[i] = mean + c[5] * c[4] * z[i];
        if (pred_xstart) pred_xstart[i] = x0;
    }
}

// q_sample (multimodal_gaussian_diffusion.py:187-205): y = a[b] * x0 + s[b] * noise.  coef [B][2].
__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                const float* __restrict__ coef, long long per_sample, long long total,
                                float* __restrict__ y) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float* c = coef + (i / per_sample) * 2;
        y[i] = c[0] * x0[i] + c[1] * noise[i];
    }
}

// Sample epilogue of the sampling scripts (py_scripts/multimodal_sample_sr.py:159-163): video [N, C, H, W] fp32 in
// [-1, 1] (N = batch x frames) -> uint8 [N, H, W, C] = ((x + 1) * 127.5).clamp(0, 255).to(uint8) with the permute to
// channels-last folded in.  Thread = one output pixel (C bytes); reads are coalesced per channel plane.
__global__ void sample_epilogue_kernel(const float* __restrict__ x, unsigned char* __restrict__ out, long long n_img, int C,
                                       int HW) {
    const long long total = n_img * HW;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long img = i / HW;
        const int px = static_cast<int>(i - img * HW);
        const float* src = x + (img * C) * HW + px;
        unsigned char* dst = out + i * C;
        for (int c = 0; c < C; ++c) {
            const float v = fminf(fmaxf((src[static_cast<long long>(c) * HW] + 1.0f) * 127.5f, 0.0f), 255.0f);
            dst[c] = static_cast<unsigned char>(v);   // truncation, like torch's float -> uint8 cast
        }
    }
}

// ---------------------------------------------------------------------------
// DPM-Solver state updates (multimodal_dpm_solver_plus.py:532-1036): every first/second/third-order update, and the
// x0 conversion of data_prediction_fn (:419-440), is x_t = sum_i c_i * tensor_i with step-wide scalars c_i.
// ---------------------------------------------------------------------------
struct LinCombArgs {
    const float* src[4];
    float coef[4];
    int n;
};
__global__ void lincomb_kernel(LinCombArgs a, long long n4, long long total, float* __restrict__ out) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (t < a.n) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(a.src[t]) + i);
                const float c = a.coef[t];
                acc.x = fmaf(c, v.x, acc.x); acc.y = fmaf(c, v.y, acc.y); acc.z = fmaf(c, v.z, acc.z); acc.w = fmaf(c, v.w, acc.w);
            }
        }
        reinterpret_cast<float4*>(out)[i] = acc;
    }
    // tail (numel % 4)
    for (long long i = n4 * 4 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        float acc = 0.f;
        for (int t = 0; t < a.n; ++t) acc = fmaf(a.coef[t], a.src[t][i], acc);
        out[i] = acc;
    }
}

// Dynamic thresholding tail (data_prediction_fn :431-438): x0 = clamp(x0, -s_b, s_b) / (s_b / max_val), s_b per sample.
__global__ void dpm_threshold_kernel(float* __restrict__ x0, const float* __restrict__ s, long long per_sample, long long total,
                                     float max_val) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float sb = s[i / per_sample];
        x0[i] = fminf(fmaxf(x0[i], -sb), sb) / (sb / max_val);
    }
}

// Adaptive-step error (dpm_solver_adaptive :1134-1138): per sample sum of ((hi - lo) / max(atol, rtol*max(|lo|,|prev|)))^2.
__global__ void dpm_error_kernel(const float* __restrict__ hi, const float* __restrict__ lo, const float* __restrict__ prev,
                                 long long per_sample, float atol, float rtol, double* __restrict__ out) {
    const int b = blockIdx.y;
    const float* h = hi + b * per_sample;
    const float* l = lo + b * per_sample;
    const float* p = prev + b * per_sample;
    float acc = 0.f;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < per_sample;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float d = fmaxf(atol, rtol * fmaxf(fabsf(l[i]), fabsf(p[i])));
        const float e = (h[i] - l[i]) / d;
        acc = fmaf(e, e, acc);
    }
    __shared__ float red[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) atomicAdd(&out[b], static_cast<double>(v));
    }
}

// ---------------------------------------------------------------------------
// Weight repacking (once per weight load): fp32 conv weight [Co][Ci][T] -> fp16 [Co][ld]
// at column col_off with k = t*Ci + ci  (K order of conv_gemm_kernel: tap-major).
// ---------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, act_t* __restrict__ dst, int Co, int Ci, int T,
                                   long long ld, long long col_off) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(Co) * Ci * T;
    if (idx >= total) return;
    const int t = static_cast<int>(idx % T);
    const long long r = idx / T;
    const int ci = static_cast<int>(r % Ci);
    const int co = static_cast<int>(r / Ci);
    dst[co * ld + col_off + static_cast<long long>(t) * Ci + ci] = __float2half_rn(w[idx]);
}
__global__ void pack_identity_kernel(act_t* __restrict__ dst, int C, long long ld, long long col_off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) dst[i * ld + col_off + i] = __float2half_rn(1.0f);
}
__global__ void add_vec_kernel(float* __restrict__ dst, const float* __restrict__ src, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}
__global__ void cast_f32_to_f16_kernel(const float* __restrict__ s, act_t* __restrict__ d, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) d[i] = __float2half_rn(s[i]);
}
__global__ void cast_f16_to_f32_kernel(const act_t* __restrict__ s, float* __restrict__ d, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) d[i] = __half2float(s[i]);
}

}  // namespace mmd
