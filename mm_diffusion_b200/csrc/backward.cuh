// CUDA-core kernels of the training backward: GroupNorm (+FiLM, +SiLU) backward,
// per-pixel temporal GroupNorm / temporal attention backward, resampling adjoints, gradient scaling, the narrow
// output heads and the time-embedding MLPs.  All activation gradients are channels-last fp16 carrying one global
// power-of-two scale (device scalar `gscale[0]`, inverse in `gscale[1]`) chosen from the incoming output gradient so
// fp16 keeps its range; parameter gradients are accumulated UNSCALED in fp32.
// Reference semantics: torch.autograd through the modules cited per kernel.
#pragma once
#include "common.cuh"
#include "elementwise.cuh"

namespace mmd {

MMD_DEVINL float dsilu_f(float u) {
    const float s = 1.0f / (1.0f + __expf(-u));
    return s * (1.0f + u * (1.0f - s));
}
// silu'(u) through one MUFU op (sigmoid = 0.5 (1 + tanh(u / 2)), tanh.approx abs error ~5e-4: below the fp16 rounding
// of the gradients it multiplies)
MMD_DEVINL float dsilu_fast(float u) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * u));
    const float s = fmaf(0.5f, t, 0.5f);
    return s * fmaf(u, 1.0f - s, 1.0f);
}

// ---------------------------------------------------------------------------
// Gradient scale: gscale = {s, 1/s}, s = 2^floor(log2(16 / amax |dY|)) over both modalities.
// ---------------------------------------------------------------------------
__global__ void absmax_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ out_bits) {
    float m = 0.f;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float v = fabsf(x[i]);
        if (v == v && v < 3.0e38f) m = fmaxf(m, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));   // non-negative floats order like their bit patterns
}
__global__ void make_gscale_kernel(const unsigned int* __restrict__ amax_bits, float* __restrict__ gscale) {
    const float a = __uint_as_float(*amax_bits);
    float s = 1.0f;
    if (a > 0.f) s = exp2f(floorf(log2f(16.0f / a)));
    s = fminf(fmaxf(s, 1.0f / 16777216.0f), 1.0e9f);
    gscale[0] = s;
    gscale[1] = 1.0f / s;
}

// y (+)= x for fp16 gradients (residual / identity paths).
__global__ void grad_add_kernel(const act_t* __restrict__ x, act_t* __restrict__ y, long long n8, int accumulate) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        uint4 a = __ldg(reinterpret_cast<const uint4*>(x) + i);
        if (accumulate) {
            const uint4 b = reinterpret_cast<const uint4*>(y)[i];
            __half2* ha = reinterpret_cast<__half2*>(&a);
            const __half2* hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
            for (int k = 0; k < 4; ++k) ha[k] = __hadd2(ha[k], hb[k]);
        }
        reinterpret_cast<uint4*>(y)[i] = a;
    }
}
// strided variant: rows x C block of x (ld ldx) added into y (ld ldy); C % 8 == 0
__global__ void grad_add2d_kernel(const act_t* __restrict__ x, long long ldx, act_t* __restrict__ y, long long ldy, long long rows,
                                  int C, int accumulate) {
    const int vpr = C / 8;
    const long long total = rows * vpr;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / vpr;
        const int v = static_cast<int>(i - r * vpr);
        uint4 a = __ldg(reinterpret_cast<const uint4*>(x + r * ldx + v * 8));
        uint4* dst = reinterpret_cast<uint4*>(y + r * ldy + v * 8);
        if (accumulate) {
            const uint4 b = *dst;
            __half2* ha = reinterpret_cast<__half2*>(&a);
            const __half2* hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
            for (int k = 0; k < 4; ++k) ha[k] = __hadd2(ha[k], hb[k]);
        }
        *dst = a;
    }
}

// ---------------------------------------------------------------------------
// GroupNorm32 (+FiLM, +SiLU) backward.  Forward (gn_apply_kernel): with xh = (x - mean_g) * rstd_g,
//   u = (xh * gamma + beta) * (1 + sc) + sh,  y = silu(u) | u.
// Pass A (reduce): T[ns][c] = { sum_rows du, sum_rows du * xh } with du = dy * silu'(u).
// Pass B (apply):  k_c = (1 + sc) gamma_c;  S1_g = sum_c k_c T1_c;  S2_g = sum_c k_c T2_c;  m = rows * C/32
//   dx = rstd_g * (k_c du - S1_g / m - xh S2_g / m)            (torch.nn.functional.group_norm backward)
// Finalize: dgamma_c += sum_ns (1 + sc) T2;  dbeta_c += sum_ns (1 + sc) T1;
//           dsc[b][c] += sum_{ns of b} (gamma T2 + beta T1);  dsh[b][c] += sum T1      (nn.py:16-33, multimodal_unet.py:459-470)
// ---------------------------------------------------------------------------
struct GnBwdArgs {
    GnSrc s;
    int R;                 // rows per domain
    int rows_per_block;
    const double* sums;    // forward statistics [ns * nsub][32][2]
    int nsub;
    long long stat_rows;
    const float* gamma;
    const float* beta;
    const float* film;     // [B][film_ld] or null
    int film_ld;
    int ns_per_batch;
    int do_silu;
    const act_t* dy;       // [ns * R][C]
    float* T;              // [ns][C][2]
    const DropState* drop; // dropout applied to the forward output (null = none): dy is masked / scaled on the fly
    uint32_t drop_site;
};

// shared layout: a[C] b[C] rs[C] mr[C] (u = x a + b, xh = x rs - mr) | gstat[64]
MMD_DEVINL void gn_bwd_prologue(const GnBwdArgs& g, int ns, float* coef, float* gstat) {
    const int C = g.s.c1 + g.s.c2;
    const int cpg = C / 32;
    if (threadIdx.x < 32) {
        const double inv_n = 1.0 / (static_cast<double>(g.stat_rows) * cpg);
        double su = 0.0, sq = 0.0;
        for (int k = 0; k < g.nsub; ++k) {
            su += g.sums[(static_cast<size_t>(ns) * g.nsub + k) * 64 + 2 * threadIdx.x];
            sq += g.sums[(static_cast<size_t>(ns) * g.nsub + k) * 64 + 2 * threadIdx.x + 1];
        }
        const double mean = su * inv_n;
        double var = sq * inv_n - mean * mean;
        if (var < 0) var = 0;
        gstat[2 * threadIdx.x] = static_cast<float>(mean);
        gstat[2 * threadIdx.x + 1] = rsqrtf(static_cast<float>(var) + 1e-5f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int grp = c / cpg;
        const float rstd = gstat[2 * grp + 1], mean = gstat[2 * grp];
        float a = rstd * g.gamma[c];
        float b = g.beta[c] - mean * a;
        if (g.film != nullptr) {
            const float* fb = g.film + static_cast<size_t>(ns / g.ns_per_batch) * g.film_ld;
            const float sc = 1.f + fb[c];
            a *= sc;
            b = b * sc + fb[C + c];
        }
        coef[c] = a;
        coef[C + c] = b;
        coef[2 * C + c] = rstd;
        coef[3 * C + c] = mean * rstd;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256, 2) gn_bwd_reduce_kernel(GnBwdArgs g) {
    extern __shared__ float gsh[];   // coef[4C] | gstat[64] | red[2C]
    const int C = g.s.c1 + g.s.c2;
    float* coef = gsh;
    float* gstat = gsh + 4 * C;
    float* red = gstat + 64;
    const int ns = blockIdx.y;
    DropState ds{0u, 0u, 0u, 0u};
    if (g.drop != nullptr) ds = *g.drop;
    const bool dropping = ds.thresh16 != 0u;
    const float drop_scale = __uint_as_float(ds.scale_bits);
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
    gn_bwd_prologue(g, ns, coef, gstat);
    const int vpr = C / 8;
    const int rows_per_pass = blockDim.x / vpr;
    const int vec = threadIdx.x % vpr, rsub = threadIdx.x / vpr;
    if (rsub < rows_per_pass) {
        const int r0 = blockIdx.x * g.rows_per_block;
        const int r1 = min(g.R, r0 + g.rows_per_block);
        const int c0 = vec * 8;
        const act_t* base;
        int ld;
        if (c0 < g.s.c1) { base = g.s.x1 + c0; ld = g.s.ld1; } else { base = g.s.x2 + (c0 - g.s.c1); ld = g.s.ld2; }
        base += static_cast<size_t>(ns) * g.R * ld;
        const act_t* dyb = g.dy + static_cast<size_t>(ns) * g.R * C + c0;
        float ca[8], cb[8], crs[8], cmr[8];   // this thread's 8 channels: u = x a + b, xh = x rs - mr
#pragma unroll
        for (int i = 0; i < 8; ++i) { ca[i] = coef[c0 + i]; cb[i] = coef[C + c0 + i]; crs[i] = coef[2 * C + c0 + i]; cmr[i] = coef[3 * C + c0 + i]; }
        float t1[8], t2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { t1[i] = 0.f; t2[i] = 0.f; }
        for (int r = r0 + rsub; r < r1; r += GN_UNROLL * rows_per_pass) {
            uint4 xr[GN_UNROLL], dr[GN_UNROLL];
#pragma unroll
            for (int u = 0; u < GN_UNROLL; ++u) {
                const int rr = r + u * rows_per_pass;
                if (rr < r1) {
                    xr[u] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(rr) * ld));
                    dr[u] = __ldg(reinterpret_cast<const uint4*>(dyb + static_cast<size_t>(rr) * C));
                } else {
                    xr[u] = make_uint4(0, 0, 0, 0);
                    dr[u] = make_uint4(0, 0, 0, 0);   // dy = 0: contributes nothing
                }
            }
#pragma unroll
            for (int u = 0; u < GN_UNROLL; ++u) {
                const __half* xh = reinterpret_cast<const __half*>(&xr[u]);
                const __half* dh = reinterpret_cast<const __half*>(&dr[u]);
                uint32_t keep = 0xFFu;
                if (dropping) {
                    const int rr = r + u * rows_per_pass;
                    keep = dropout_keep8(ds, g.drop_site, ((static_cast<unsigned long long>(ns) * g.R + rr) * C + c0) >> 3);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float x = __half2float(xh[i]);
                    float du = __half2float(dh[i]);
                    if (dropping) du = ((keep >> i) & 1u) ? du * drop_scale : 0.f;
                    if (g.do_silu) du *= dsilu_fast(fmaf(x, ca[i], cb[i]));
                    const float xn = fmaf(x, crs[i], -cmr[i]);
                    t1[i] += du;
                    t2[i] = fmaf(du, xn, t2[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            atomicAdd(&red[2 * (c0 + i)], t1[i]);
            atomicAdd(&red[2 * (c0 + i) + 1], t2[i]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&g.T[static_cast<size_t>(ns) * 2 * C + i], red[i]);
}

struct GnBwdOut {
    act_t* dx1; int ld1; int acc1;   // gradient of source 1 ([ns*R][ld1], columns [0, c1))
    act_t* dx2; int ld2; int acc2;   // gradient of source 2 (nullable)
};

__global__ void __launch_bounds__(256, 2) gn_bwd_apply_kernel(GnBwdArgs g, GnBwdOut o) {
    extern __shared__ float gsh[];   // coef[4C] | gstat[64] | p[C] | qr[64]
    const int C = g.s.c1 + g.s.c2;
    const int cpg = C / 32;
    float* coef = gsh;
    float* gstat = gsh + 4 * C;
    float* pc = gstat + 64;
    float* qr = pc + C;
    const int ns = blockIdx.y;
    DropState ds{0u, 0u, 0u, 0u};
    if (g.drop != nullptr) ds = *g.drop;
    const bool dropping = ds.thresh16 != 0u;
    const float drop_scale = __uint_as_float(ds.scale_bits);
    gn_bwd_prologue(g, ns, coef, gstat);
    const float* T = g.T + static_cast<size_t>(ns) * 2 * C;
    if (threadIdx.x < 32) {
        const int grp = threadIdx.x;
        float s1 = 0.f, s2 = 0.f;
        for (int c = grp * cpg; c < (grp + 1) * cpg; ++c) {
            float k = g.gamma[c];
            if (g.film != nullptr) k *= 1.f + g.film[static_cast<size_t>(ns / g.ns_per_batch) * g.film_ld + c];
            s1 = fmaf(k, T[2 * c], s1);
            s2 = fmaf(k, T[2 * c + 1], s2);
        }
        const float inv_m = 1.0f / (static_cast<float>(g.R) * cpg);
        const float rstd = gstat[2 * grp + 1];
        qr[2 * grp] = rstd * s1 * inv_m;
        qr[2 * grp + 1] = rstd * s2 * inv_m;
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float k = g.gamma[c];
        if (g.film != nullptr) k *= 1.f + g.film[static_cast<size_t>(ns / g.ns_per_batch) * g.film_ld + c];
        pc[c] = k * coef[2 * C + c];
    }
    __syncthreads();
    const int vpr = C / 8;
    const int rows_per_pass = blockDim.x / vpr;
    const int vec = threadIdx.x % vpr, rsub = threadIdx.x / vpr;
    if (rsub >= rows_per_pass) return;
    const int r0 = blockIdx.x * g.rows_per_block;
    const int r1 = min(g.R, r0 + g.rows_per_block);
    const int c0 = vec * 8;
    const act_t* base;
    int ld;
    act_t* dbase;
    int dld, dacc;
    if (c0 < g.s.c1) {
        base = g.s.x1 + c0; ld = g.s.ld1;
        dbase = o.dx1 + c0; dld = o.ld1; dacc = o.acc1;
    } else {
        base = g.s.x2 + (c0 - g.s.c1); ld = g.s.ld2;
        dbase = o.dx2 ? o.dx2 + (c0 - g.s.c1) : nullptr; dld = o.ld2; dacc = o.acc2;
    }
    if (dbase == nullptr) return;
    base += static_cast<size_t>(ns) * g.R * ld;
    dbase += static_cast<size_t>(ns) * g.R * dld;
    const act_t* dyb = g.dy + static_cast<size_t>(ns) * g.R * C + c0;
    // dx = p du - q - xh r with xh = x rs - mr  ==>  dx = p du + x (-rs r) + (mr r - q): five coefficient rows in registers
    float ca[8], cb[8], cp[8], cx[8], cc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = c0 + i;
        const int grp = c / cpg;
        ca[i] = coef[c]; cb[i] = coef[C + c];
        cp[i] = pc[c];
        cx[i] = -coef[2 * C + c] * qr[2 * grp + 1];
        cc[i] = coef[3 * C + c] * qr[2 * grp + 1] - qr[2 * grp];
    }
    for (int r = r0 + rsub; r < r1; r += GN_UNROLL * rows_per_pass) {
        uint4 xr[GN_UNROLL], dr[GN_UNROLL], pv[GN_UNROLL];
#pragma unroll
        for (int u = 0; u < GN_UNROLL; ++u) {
            const int rr = r + u * rows_per_pass;
            if (rr < r1) {
                xr[u] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(rr) * ld));
                dr[u] = __ldg(reinterpret_cast<const uint4*>(dyb + static_cast<size_t>(rr) * C));
                if (dacc) pv[u] = *reinterpret_cast<const uint4*>(dbase + static_cast<size_t>(rr) * dld);
            }
        }
#pragma unroll
        for (int u = 0; u < GN_UNROLL; ++u) {
            const int rr = r + u * rows_per_pass;
            if (rr >= r1) break;
            const __half* xh = reinterpret_cast<const __half*>(&xr[u]);
            const __half* dh = reinterpret_cast<const __half*>(&dr[u]);
            const __half* ph = reinterpret_cast<const __half*>(&pv[u]);
            uint4 outv;
            __half* oh = reinterpret_cast<__half*>(&outv);
            uint32_t keep = 0xFFu;
            if (dropping) keep = dropout_keep8(ds, g.drop_site, ((static_cast<unsigned long long>(ns) * g.R + rr) * C + c0) >> 3);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float x = __half2float(xh[i]);
                float du = __half2float(dh[i]);
                if (dropping) du = ((keep >> i) & 1u) ? du * drop_scale : 0.f;
                if (g.do_silu) du *= dsilu_fast(fmaf(x, ca[i], cb[i]));
                float dx = fmaf(cp[i], du, fmaf(x, cx[i], cc[i]));
                if (dacc) dx += __half2float(ph[i]);
                oh[i] = __float2half_rn(dx);
            }
            *reinterpret_cast<uint4*>(dbase + static_cast<size_t>(rr) * dld) = outv;
        }
    }
}

// one thread per channel; dfilm [B][film_ld] (scale at c, shift at C + c), accumulated with atomics (both modalities
// of a ResBlock share one emb_layers output); parameter gradients are unscaled by gscale[1].
__global__ void gn_bwd_finalize_kernel(const float* __restrict__ T, int ns, int C, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ film, int film_ld,
                                       int ns_per_batch, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ dfilm, const float* __restrict__ gscale) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float inv = gscale ? gscale[1] : 1.0f;
    float dg = 0.f, db = 0.f;
    const float gm = gamma[c], bt = beta[c];
    for (int d = 0; d < ns; ++d) {
        const float t1 = T[(static_cast<size_t>(d) * C + c) * 2], t2 = T[(static_cast<size_t>(d) * C + c) * 2 + 1];
        float k = 1.f;
        if (film != nullptr) {
            const int b = d / ns_per_batch;
            k += film[static_cast<size_t>(b) * film_ld + c];
            atomicAdd(&dfilm[static_cast<size_t>(b) * film_ld + c], gm * t2 + bt * t1);
            atomicAdd(&dfilm[static_cast<size_t>(b) * film_ld + C + c], t1);
        }
        dg = fmaf(k, t2, dg);
        db = fmaf(k, t1, db);
    }
    atomicAdd(&dgamma[c], dg * inv);
    atomicAdd(&dbeta[c], db * inv);
}

// ---------------------------------------------------------------------------
// Per-pixel temporal GroupNorm backward (forward: gn_temporal_kernel; domain = F frames x C/32 channels of a pixel).
// One thread per (pixel, group); dgamma / dbeta via shared-memory then global atomics.
// ---------------------------------------------------------------------------
template <int CPG, int F_>
__global__ void __launch_bounds__(128) gn_temporal_bwd_kernel(const act_t* __restrict__ x, const act_t* __restrict__ dy,
                                                              act_t* __restrict__ dx, const float* __restrict__ gamma,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int P,
                                                              int C, const float* __restrict__ gscale) {
    extern __shared__ float tsh[];   // [2][C]
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) tsh[i] = 0.f;
    __syncthreads();
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(B) * P * 32;
    if (idx < total) {
        const int g = static_cast<int>(idx % 32);
        const long long bp = idx / 32;
        const int p = static_cast<int>(bp % P);
        const int b = static_cast<int>(bp / P);
        const size_t fstride = static_cast<size_t>(P) * C;
        const size_t base = (static_cast<size_t>(b) * F_ * P + p) * C + g * CPG;
        float su = 0.f, ss = 0.f;
        for (int f = 0; f < F_; ++f)
#pragma unroll
            for (int i = 0; i < CPG; ++i) {
                const float v = __half2float(x[base + f * fstride + i]);
                su += v;
                ss = fmaf(v, v, ss);
            }
        const float inv_n = 1.f / (F_ * CPG);
        const float mean = su * inv_n;
        const float rstd = rsqrtf(fmaxf(ss * inv_n - mean * mean, 0.f) + 1e-5f);
        float s1 = 0.f, s2 = 0.f;
        float dg[CPG], db[CPG];
#pragma unroll
        for (int i = 0; i < CPG; ++i) { dg[i] = 0.f; db[i] = 0.f; }
        for (int f = 0; f < F_; ++f)
#pragma unroll
            for (int i = 0; i < CPG; ++i) {
                const float xn = (__half2float(x[base + f * fstride + i]) - mean) * rstd;
                const float d = __half2float(dy[base + f * fstride + i]);
                const float dxh = d * gamma[g * CPG + i];
                s1 += dxh;
                s2 = fmaf(dxh, xn, s2);
                dg[i] = fmaf(d, xn, dg[i]);
                db[i] += d;
            }
        s1 *= inv_n;
        s2 *= inv_n;
        for (int f = 0; f < F_; ++f)
#pragma unroll
            for (int i = 0; i < CPG; ++i) {
                const float xn = (__half2float(x[base + f * fstride + i]) - mean) * rstd;
                const float d = __half2float(dy[base + f * fstride + i]);
                dx[base + f * fstride + i] = __float2half_rn(rstd * (d * gamma[g * CPG + i] - s1 - xn * s2));
            }
#pragma unroll
        for (int i = 0; i < CPG; ++i) {
            atomicAdd(&tsh[g * CPG + i], dg[i]);
            atomicAdd(&tsh[C + g * CPG + i], db[i]);
        }
    }
    __syncthreads();
    const float inv = gscale ? gscale[1] : 1.0f;
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        atomicAdd(&dgamma[i], tsh[i] * inv);
        atomicAdd(&dbeta[i], tsh[C + i] * inv);
    }
}

// ---------------------------------------------------------------------------
// Temporal self-attention backward (forward: temporal_attn_kernel).  One warp per (sample, pixel, head):
//   P = softmax(Q K^T / sqrt(d)); dV = P^T dO; dP = dO V^T; dS = P o (dP - rowsum(P o dP)) / sqrt(d);
//   dQ = dS K; dK = dS^T Q.   qkv / dqkv: [B][F][P][3C]; d_out: [B][F][P][C].
// ---------------------------------------------------------------------------
constexpr int TAB_WARPS = 4;
template <int F_>
__global__ void __launch_bounds__(TAB_WARPS * 32) temporal_attn_bwd_kernel(const act_t* __restrict__ qkv, const act_t* __restrict__ d_out,
                                                                           act_t* __restrict__ dqkv, int B, int P, int C, int heads) {
    extern __shared__ __align__(16) uint8_t tbs[];
    const int d = C / heads;
    const int pitch = d + 2;   // halves
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = static_cast<size_t>(4) * 16 * pitch * sizeof(act_t) + 2 * 16 * 17 * sizeof(float);
    uint8_t* wb = tbs + warp * per_warp;
    act_t* sq = reinterpret_cast<act_t*>(wb);
    act_t* sk = sq + 16 * pitch;
    act_t* sv = sk + 16 * pitch;
    act_t* sdo = sv + 16 * pitch;
    float* sp = reinterpret_cast<float*>(sdo + 16 * pitch);   // P  [16][17]
    float* sds = sp + 16 * 17;                                // dS [16][17]
    const float rs = rsqrtf(static_cast<float>(d));
    const long long total = static_cast<long long>(B) * P * heads;
    for (long long item = static_cast<long long>(blockIdx.x) * TAB_WARPS + warp; item < total;
         item += static_cast<long long>(gridDim.x) * TAB_WARPS) {
        const int h = static_cast<int>(item % heads);
        const long long bp = item / heads;
        const int p = static_cast<int>(bp % P);
        const int b = static_cast<int>(bp / P);
        for (int i = lane; i < 16 * (d / 2); i += 32) {
            const int f = i / (d / 2), c2 = i - f * (d / 2);
            uint32_t q = 0, k = 0, v = 0, o = 0;
            if (f < F_) {
                const size_t tok = (static_cast<size_t>(b) * F_ + f) * P + p;
                const act_t* src = qkv + tok * (3 * static_cast<size_t>(C)) + h * d + 2 * c2;
                q = *reinterpret_cast<const uint32_t*>(src);
                k = *reinterpret_cast<const uint32_t*>(src + C);
                v = *reinterpret_cast<const uint32_t*>(src + 2 * C);
                o = *reinterpret_cast<const uint32_t*>(d_out + tok * C + h * d + 2 * c2);
            }
            *reinterpret_cast<uint32_t*>(sq + f * pitch + 2 * c2) = q;
            *reinterpret_cast<uint32_t*>(sk + f * pitch + 2 * c2) = k;
            *reinterpret_cast<uint32_t*>(sv + f * pitch + 2 * c2) = v;
            *reinterpret_cast<uint32_t*>(sdo + f * pitch + 2 * c2) = o;
        }
        __syncwarp();
        // lane -> row i = lane / 2, columns j0 .. j0 + 7
        const int i = lane >> 1, j0 = (lane & 1) * 8;
        float s[8], dp[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] = 0.f; dp[j] = 0.f; }
        for (int c = 0; c < d; c += 2) {
            const float2 qv = __half22float2(*reinterpret_cast<const __half2*>(sq + i * pitch + c));
            const float2 ov = __half22float2(*reinterpret_cast<const __half2*>(sdo + i * pitch + c));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 kv = __half22float2(*reinterpret_cast<const __half2*>(sk + (j0 + j) * pitch + c));
                const float2 vv = __half22float2(*reinterpret_cast<const __half2*>(sv + (j0 + j) * pitch + c));
                s[j] = fmaf(qv.x, kv.x, fmaf(qv.y, kv.y, s[j]));
                dp[j] = fmaf(ov.x, vv.x, fmaf(ov.y, vv.y, dp[j]));
            }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j] = (j0 + j < F_) ? s[j] * rs : -INFINITY;
            mx = fmaxf(mx, s[j]);
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j] = (j0 + j < F_) ? __expf(s[j] - mx) : 0.f;
            sum += s[j];
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        const float inv = 1.f / sum;
        float delta = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j] *= inv;
            delta = fmaf(s[j], dp[j], delta);
        }
        delta += __shfl_xor_sync(0xffffffffu, delta, 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool live = (i < F_) && (j0 + j < F_);
            sp[i * 17 + j0 + j] = live ? s[j] : 0.f;
            sds[i * 17 + j0 + j] = live ? s[j] * (dp[j] - delta) * rs : 0.f;
        }
        __syncwarp();
        // outputs: lane owns channel pairs c2 = lane, lane + 32, ...
        for (int c2 = lane; c2 < d / 2; c2 += 32) {
            float2 qc[16], kc[16], oc[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                qc[r] = __half22float2(*reinterpret_cast<const __half2*>(sq + r * pitch + 2 * c2));
                kc[r] = __half22float2(*reinterpret_cast<const __half2*>(sk + r * pitch + 2 * c2));
                oc[r] = __half22float2(*reinterpret_cast<const __half2*>(sdo + r * pitch + 2 * c2));
            }
#pragma unroll 1
            for (int r = 0; r < F_; ++r) {
                float2 dq = make_float2(0.f, 0.f), dk = make_float2(0.f, 0.f), dv = make_float2(0.f, 0.f);
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const float a = sds[r * 17 + t];    // dS[r][t]
                    const float bt = sds[t * 17 + r];   // dS[t][r]
                    const float pt = sp[t * 17 + r];    // P[t][r]
                    dq.x = fmaf(a, kc[t].x, dq.x); dq.y = fmaf(a, kc[t].y, dq.y);
                    dk.x = fmaf(bt, qc[t].x, dk.x); dk.y = fmaf(bt, qc[t].y, dk.y);
                    dv.x = fmaf(pt, oc[t].x, dv.x); dv.y = fmaf(pt, oc[t].y, dv.y);
                }
                const size_t tok = (static_cast<size_t>(b) * F_ + r) * P + p;
                act_t* dst = dqkv + tok * (3 * static_cast<size_t>(C)) + h * d + 2 * c2;
                *reinterpret_cast<__half2*>(dst) = __floats2half2_rn(dq.x, dq.y);
                *reinterpret_cast<__half2*>(dst + C) = __floats2half2_rn(dk.x, dk.y);
                *reinterpret_cast<__half2*>(dst + 2 * C) = __floats2half2_rn(dv.x, dv.y);
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// Resampling adjoints (forward: resample_kernel).  mode 0/1: gradient of avg-pool = 0.25 * nearest expansion of dy;
// mode 2/3: gradient of nearest upsampling = sum over the 4 children of dy.  (N, H, W) are the FORWARD INPUT extents.
// dx (+)= adjoint(dy).
// ---------------------------------------------------------------------------
__global__ void resample_bwd_kernel(const act_t* __restrict__ dy, act_t* __restrict__ dx, int mode, int N, int H, int W, int C,
                                    int accumulate) {
    const int vpr = C / 8;
    const long long total = (mode == 0 || mode == 2) ? static_cast<long long>(N) * H * W * vpr : static_cast<long long>(N) * H * vpr;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int v = static_cast<int>(idx % vpr);
    long long t = idx / vpr;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    auto add = [&](const act_t* src, float sc) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(src));
        const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(h[i]);
            acc[2 * i] = fmaf(sc, f.x, acc[2 * i]);
            acc[2 * i + 1] = fmaf(sc, f.y, acc[2 * i + 1]);
        }
    };
    if (mode == 0) {          // x [N][H][W][C], dy [N][H/2][W/2][C]
        const int w = static_cast<int>(t % W); t /= W;
        const int h = static_cast<int>(t % H);
        const long long n = t / H;
        add(dy + ((n * (H / 2) + h / 2) * (W / 2) + w / 2) * static_cast<long long>(C) + v * 8, 0.25f);
    } else if (mode == 1) {   // x [N][L][C], dy [N][L/4][C]
        const int l = static_cast<int>(t % H);
        const long long n = t / H;
        add(dy + (n * (H / 4) + l / 4) * static_cast<long long>(C) + v * 8, 0.25f);
    } else if (mode == 2) {   // x [N][H][W][C], dy [N][2H][2W][C]
        const int w = static_cast<int>(t % W); t /= W;
        const int h = static_cast<int>(t % H);
        const long long n = t / H;
        const act_t* b0 = dy + ((n * (2 * H) + 2 * h) * (2 * W) + 2 * w) * static_cast<long long>(C) + v * 8;
        add(b0, 1.f); add(b0 + C, 1.f);
        add(b0 + static_cast<long long>(2 * W) * C, 1.f); add(b0 + static_cast<long long>(2 * W) * C + C, 1.f);
    } else {                  // x [N][L][C], dy [N][4L][C]
        const int l = static_cast<int>(t % H);
        const long long n = t / H;
        const act_t* b0 = dy + (n * (4LL * H) + 4LL * l) * C + v * 8;
        for (int k = 0; k < 4; ++k) add(b0 + static_cast<long long>(k) * C, 1.f);
    }
    uint4* dst = reinterpret_cast<uint4*>(dx) + idx;
    if (accumulate) {
        const uint4 prev = *dst;
        const __half2* h = reinterpret_cast<const __half2*>(&prev);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(h[i]);
            acc[2 * i] += f.x;
            acc[2 * i + 1] += f.y;
        }
    }
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) oh[i] = __floats2half2_rn(acc[2 * i], acc[2 * i + 1]);
    *dst = o;
}

// ---------------------------------------------------------------------------
// Narrow output heads (video_out Conv3d 3x3x3 -> 3 channels, audio_out Conv1d k3 -> 1 channel; forward: conv_gemm BN = 16
// with an fp32 strided scatter).  dout is the fp32 gradient in the API layout; token coordinates follow the forward
// geometry (dims[0..3], innermost first); `gscale[0]` scales it into the fp16 gradient range.  Both adjoints run on the
// tensor-core kernels (see head_im2col_bwd_kernel below).
// ---------------------------------------------------------------------------
struct HeadGeom {
    int ncoord;                 // token coordinates used (1..4)
    int dims[4];
    long long ostride[4];       // fp32 layout strides per coordinate
    long long ostride_c;        // per output channel
    int n_out;                  // <= 4
    int n_taps;                 // <= 27
    int tap[27][3];
    int C;                      // input channels (multiple of 8)
};

// Tensor-core path of the head adjoints: G[tok][t * n_out + n] = s * dout[n][tok - delta(t)] (fp16, zero outside the
// tensor and in the padding columns).  Then dX = G * Wt (forward implicit-GEMM kernel, one tap) and
// dW[(t, n)][c] = sum_tok G[tok][(t, n)] * A[tok][c] (conv_wgrad_kernel): both adjoints become plain token GEMMs.
__global__ void head_im2col_bwd_kernel(HeadGeom g, const float* __restrict__ dout, act_t* __restrict__ G, int ldG, long long tokens,
                                       const float* __restrict__ gscale) {
    const int vpr = ldG / 8;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= tokens * vpr) return;
    const int v = static_cast<int>(idx % vpr);
    long long tok = idx / vpr;
    int co[4];
    long long r = tok;
    for (int i = 0; i < 4; ++i) { co[i] = static_cast<int>(r % g.dims[i]); r /= g.dims[i]; }
    const float s = gscale ? gscale[0] : 1.0f;
    const int terms = g.n_taps * g.n_out;
    uint4 o;
    __half* oh = reinterpret_cast<__half*>(&o);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = v * 8 + i;
        float val = 0.f;
        if (k < terms) {
            const int t = k / g.n_out, n = k - t * g.n_out;
            const int c0 = co[0] - g.tap[t][0], c1 = co[1] - g.tap[t][1], c2 = co[2] - g.tap[t][2];
            if (c0 >= 0 && c0 < g.dims[0] && c1 >= 0 && c1 < g.dims[1] && c2 >= 0 && c2 < g.dims[2])
                val = s * __ldg(dout + c0 * g.ostride[0] + c1 * g.ostride[1] + c2 * g.ostride[2] + co[3] * g.ostride[3] + n * g.ostride_c);
        }
        oh[i] = __float2half_rn(val);
    }
    reinterpret_cast<uint4*>(G)[idx] = o;
}
// dst[c][t * n_out + n] = w[n][c][t]  (fp32 [n_out][C][T] -> fp16 [C][ld]; padding columns stay zero)
__global__ void pack_head_t_kernel(const float* __restrict__ w, act_t* __restrict__ dst, int n_out, int C, int T, int ld) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_out * C * T) return;
    const int t = idx % T;
    const int r = idx / T;
    const int c = r % C, n = r / C;
    dst[static_cast<size_t>(c) * ld + t * n_out + n] = __float2half_rn(w[idx]);
}
// g[n][c][t] += inv_s * dwpk[(t * n_out + n)][c]
__global__ void unpack_head_wgrad_kernel(const float* __restrict__ dwpk, float* __restrict__ gw, int n_out, int C, int T,
                                         const float* __restrict__ gscale) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_out * C * T) return;
    const int t = idx % T;
    const int r = idx / T;
    const int c = r % C, n = r / C;
    gw[idx] += (gscale ? gscale[1] : 1.0f) * dwpk[static_cast<size_t>(t * n_out + n) * C + c];
}
// db[n] += sum over all tokens of dout[n][tok] (unscaled fp32); grid.y = n
__global__ void __launch_bounds__(256) head_bias_kernel(HeadGeom g, const float* __restrict__ dout, float* __restrict__ db, long long tokens) {
    const int n = blockIdx.y;
    float acc = 0.f;
    for (long long tok = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; tok < tokens;
         tok += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long r = tok, off = 0;
        for (int i = 0; i < 4; ++i) { off += (r % g.dims[i]) * g.ostride[i]; r /= g.dims[i]; }
        acc += dout[off + n * g.ostride_c];
    }
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        atomicAdd(&db[n], t);
    }
}

// ---------------------------------------------------------------------------
// Stem adjoints (forward: im2col_{video,audio}_kernel + K = 64 GEMM).  dcol: fp16 [tokens][64] (scaled), dx: fp32 input
// gradient in the API layout (unscaled), needed only for gradient-guided sampling (multimodal_gaussian_diffusion.py:722-819).
// ---------------------------------------------------------------------------
__global__ void col2im_video_kernel(const act_t* __restrict__ dcol, float* __restrict__ dx, int BF, int Cin, int H, int W,
                                    const float* __restrict__ gscale) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(BF) * Cin * H * W;
    if (idx >= total) return;
    const int w = static_cast<int>(idx % W);
    long long t = idx / W;
    const int h = static_cast<int>(t % H); t /= H;
    const int c = static_cast<int>(t % Cin);
    const long long n = t / Cin;
    float acc = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
        // output token (yy, xx) read input (yy + tap/3 - 1, xx + tap%3 - 1) = (h, w)
        const int yy = h - (tap / 3 - 1), xx = w - (tap % 3 - 1);
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        acc += __half2float(dcol[((n * H + yy) * W + xx) * 64 + tap * Cin + c]);
    }
    dx[idx] = acc * (gscale ? gscale[1] : 1.0f);
}
__global__ void col2im_audio_kernel(const act_t* __restrict__ dcol, float* __restrict__ dx, int B, int Cin, int L,
                                    const float* __restrict__ gscale) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(B) * Cin * L;
    if (idx >= total) return;
    const int l = static_cast<int>(idx % L);
    long long t = idx / L;
    const int c = static_cast<int>(t % Cin);
    const long long n = t / Cin;
    float acc = 0.f;
    for (int tap = 0; tap < 3; ++tap) {
        const int ll = l - (tap - 1);
        if (ll < 0 || ll >= L) continue;
        acc += __half2float(dcol[(n * L + ll) * 64 + tap * Cin + c]);
    }
    dx[idx] = acc * (gscale ? gscale[1] : 1.0f);
}

// ---------------------------------------------------------------------------
// emb_layers / time_embed backward (forward: emb_layers_kernel, time_embed_kernel; all fp32).
//   demb_all [B][rows] (scaled by s) -> dW_all[j][k] += inv_s * sum_b demb[b][j] silu_emb[b][k]; db_all[j] += inv_s * sum_b demb[b][j]
//   dsilu[b][k] = sum_j demb[b][j] W[j][k]  (still scaled)
// ---------------------------------------------------------------------------
__global__ void emb_layers_bwd_w_kernel(const float* __restrict__ demb, const float* __restrict__ silu_emb, int B, int dim, int rows,
                                        float* __restrict__ dw, float* __restrict__ db, const float* __restrict__ gscale) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(rows) * dim) return;
    const int k = static_cast<int>(idx % dim);
    const int j = static_cast<int>(idx / dim);
    const float inv = gscale ? gscale[1] : 1.0f;
    float acc = 0.f, bacc = 0.f;
    for (int b = 0; b < B; ++b) {
        const float d = demb[static_cast<size_t>(b) * rows + j];
        acc = fmaf(d, silu_emb[b * dim + k], acc);
        bacc += d;
    }
    dw[idx] += acc * inv;
    if (k == 0) db[j] += bacc * inv;
}
// one block per (sample, k-chunk of 32): dsilu[b][k] = sum_j demb[b][j] * w[j][k]
__global__ void __launch_bounds__(256) emb_layers_bwd_x_kernel(const float* __restrict__ demb, const float* __restrict__ w, int dim, int rows,
                                                               float* __restrict__ dsilu) {
    const int b = blockIdx.y;
    const int k = blockIdx.x * 32 + (threadIdx.x & 31);
    const int jl = threadIdx.x >> 5;   // 8 row lanes
    float acc = 0.f;
    if (k < dim)
        for (int j = jl; j < rows; j += 8) acc = fmaf(demb[static_cast<size_t>(b) * rows + j], w[static_cast<size_t>(j) * dim + k], acc);
    __shared__ float red[8][33];
    red[jl][threadIdx.x & 31] = acc;
    __syncthreads();
    if (jl == 0 && k < dim) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
        dsilu[b * dim + k] = t;
    }
}
// time_embed MLP backward, one block per sample (blockDim = dim); recomputes the forward activations.
//   e0 = sinusoid(t); z1 = W1 e0 + b1; e1 = silu(z1); emb = W2 e1 + b2; consumers see silu(emb).
//   partial gradients are written per sample to scratch and reduced by time_embed_bwd_reduce_kernel.
__global__ void time_embed_bwd_kernel(const float* __restrict__ t, const float* __restrict__ w1, const float* __restrict__ b1,
                                      const float* __restrict__ w2, const float* __restrict__ b2, int dim,
                                      const float* __restrict__ dsilu, float* __restrict__ scratch /*[B][4][dim]: e0, e1, demb, dz1*/) {
    extern __shared__ float sh[];  // e0[dim] e1[dim] z1[dim] demb[dim]
    float* e0 = sh;
    float* e1 = sh + dim;
    float* z1 = sh + 2 * dim;
    float* de = sh + 3 * dim;
    const int b = blockIdx.x, i = threadIdx.x;
    const int half = dim / 2;
    const float tv = t[b];
    if (i < dim) {
        const int j = (i < half) ? i : i - half;
        const float freq = expf(-logf(10000.0f) * static_cast<float>(j) / static_cast<float>(half));
        e0[i] = (i < half) ? cosf(tv * freq) : sinf(tv * freq);
    }
    __syncthreads();
    if (i < dim) {
        float acc = b1[i];
        for (int k = 0; k < dim; ++k) acc += w1[i * dim + k] * e0[k];
        z1[i] = acc;
        e1[i] = silu_f(acc);
    }
    __syncthreads();
    if (i < dim) {
        float acc = b2[i];
        for (int k = 0; k < dim; ++k) acc += w2[i * dim + k] * e1[k];
        de[i] = dsilu[b * dim + i] * dsilu_f(acc);   // d emb
    }
    __syncthreads();
    if (i < dim) {
        float acc = 0.f;   // d e1[i] = sum_j demb[j] W2[j][i]
        for (int j = 0; j < dim; ++j) acc += de[j] * w2[j * dim + i];
        const float dz = acc * dsilu_f(z1[i]);
        float* sc = scratch + static_cast<size_t>(b) * 4 * dim;
        sc[i] = e0[i];
        sc[dim + i] = e1[i];
        sc[2 * dim + i] = de[i];
        sc[3 * dim + i] = dz;
    }
}
// dW2[j][k] += inv * sum_b demb[b][j] e1[b][k]; db2[j]; dW1[j][k] += inv * sum_b dz1[b][j] e0[b][k]; db1[j]
__global__ void time_embed_bwd_reduce_kernel(const float* __restrict__ scratch, int B, int dim, float* __restrict__ dw1,
                                             float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2,
                                             const float* __restrict__ gscale) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= dim * dim) return;
    const int k = idx % dim, j = idx / dim;
    const float inv = gscale ? gscale[1] : 1.0f;
    float a1 = 0.f, a2 = 0.f, c1 = 0.f, c2 = 0.f;
    for (int b = 0; b < B; ++b) {
        const float* sc = scratch + static_cast<size_t>(b) * 4 * dim;
        a2 = fmaf(sc[2 * dim + j], sc[dim + k], a2);
        a1 = fmaf(sc[3 * dim + j], sc[k], a1);
        c2 += sc[2 * dim + j];
        c1 += sc[3 * dim + j];
    }
    dw1[idx] += a1 * inv;
    dw2[idx] += a2 * inv;
    if (k == 0) { db1[j] += c1 * inv; db2[j] += c2 * inv; }
}

// delta[h][row] = sum_c dO[row][h*d + c] * O[row][h*d + c]  (flash-attention backward pre-pass); one warp per (row, head)
__global__ void attn_delta_kernel(const act_t* __restrict__ d_out, const act_t* __restrict__ out, long long rows, int C, int heads,
                                  float* __restrict__ delta, long long delta_ld) {
    const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= rows * heads) return;
    const int h = static_cast<int>(wid % heads);
    const long long row = wid / heads;
    const int d = C / heads;
    float acc = 0.f;
    for (int c = lane * 2; c < d; c += 64) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(d_out + row * C + h * d + c));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(out + row * C + h * d + c));
        acc = fmaf(a.x, b.x, fmaf(a.y, b.y, acc));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) delta[h * delta_ld + row] = acc;
}

// column sums of an fp32 [B][rows] slice etc. are done by the kernels above; generic scaled fp32 accumulate:
__global__ void axpy_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, const float* __restrict__ gscale) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i] * (gscale ? gscale[1] : 1.0f);
}

}  // namespace mmd
