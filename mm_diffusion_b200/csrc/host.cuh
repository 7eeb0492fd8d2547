// Host-side plumbing shared by ops.cu / model.cu: error reporting, TMA descriptor
// encoding, conv-GEMM / attention problem descriptions and their launchers.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mmdiff.h"
#include "attention.cuh"
#include "attention_bwd.cuh"
#include "backward.cuh"
#include "common.cuh"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "wgrad.cuh"

namespace mmd {

// ------------------------------------------------------------------ errors
std::string& last_error_ref();
int fail(int code, const char* fmt, ...);

#define MMD_CUDA_OK(expr)                                                                          \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) return ::mmd::fail(MMD_ECUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define MMD_TRY(expr)             \
    do {                          \
        int _r = (expr);          \
        if (_r != MMD_OK) return _r; \
    } while (0)

int num_sms();

// ------------------------------------------- programmatic dependent launch
// Inside a PdlScope, a kernel launched on one of the scope's streams right after another kernel of ours carries the
// programmatic-stream-serialization attribute: its CTAs may become resident (barrier init, TMEM allocation,
// descriptor prefetch) while the predecessor drains, and block in griddepcontrol.wait until it has completed.
// Any other operation on the stream (memset, copy, event wait) disarms the next launch.  MMD_NO_PDL=1 disables it.
struct PdlState {
    bool active = false;
    cudaStream_t streams[2] = {nullptr, nullptr};
    bool armed[2] = {false, false};
};
PdlState& pdl_state();
struct PdlScope {
    PdlScope(cudaStream_t s0, cudaStream_t s1);
    ~PdlScope();
};
void pdl_break(cudaStream_t st);   // a non-kernel operation went onto `st`
void pdl_break_all();

template <typename... KArgs, typename... Args>
cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    PdlState& ps = pdl_state();
    int idx = -1;
    if (ps.active) idx = (st == ps.streams[0]) ? 0 : ((st == ps.streams[1]) ? 1 : -1);
    if (idx >= 0 && ps.armed[idx]) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
    if (idx >= 0) ps.armed[idx] = (e == cudaSuccess);
    return e;
}

// --------------------------------------------------------------- TMA maps
// dims[0] is the innermost (contiguous) extent; strides_bytes[i] is the stride of dims[i+1].
int encode_tmap(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box);

// ------------------------------------------------------------- conv-GEMM
struct ConvGeom {
    int rank = 2;                      // including channel coordinate
    long long dims[4] = {1, 1, 1, 1};  // token coordinates 1..4 (innermost first)
    int box[4] = {1, 1, 1, 1};
    long long tokens() const { return dims[0] * dims[1] * dims[2] * dims[3]; }
};
// Fill box[] so that the product is exactly 128 (power-of-two greedy, last coordinate takes the rest).
void geom_fill_box(ConvGeom& g);

struct GemmProblem {
    ConvGeom g;
    int n_src = 0;
    const act_t* src[GEMM_MAX_SRC] = {};
    int src_c[GEMM_MAX_SRC] = {};
    int n_taps = 1;
    int taps[GEMM_MAX_TAPS][3] = {};
    const act_t* w = nullptr;  // packed [n_pad][k_total], k = tap*c_total + c
    const float* bias = nullptr;  // [n_pad]
    int n = 0;       // logical output channels
    int bn = 128;    // tile N: 128, 64 or 16
    act_t* out = nullptr;     // [tokens][n]
    float* out_f32 = nullptr; // bn == 16 scatter
    long long ostride[4] = {};
    long long ostride_c = 0;
    // fused GroupNorm statistics of the output (see GemmParams)
    double* stats = nullptr;
    int stats_rows = 128;
    int stats_mul[4] = {0, 0, 0, 0};
    int stats_div = 1;
    int stats_valid_coord = -1;
    double* stats_q = nullptr;   // optional per-4-channel partial sums [domain][n / 4][2]
    // fused GroupNorm apply (+FiLM, +SiLU) on source 0 (pointwise GEMMs; see GemmParams::xf_*)
    const double* xf_sums = nullptr;
    const float* xf_gamma = nullptr;
    const float* xf_beta = nullptr;
    const float* xf_film = nullptr;
    int xf_film_ld = 0, xf_dom_per_batch = 1, xf_nsub = 1, xf_silu = 0;
    long long xf_stat_rows = 0;   // rows the statistics of one domain cover
    int xf_rows = 128;            // rows of one domain inside a tile (64 or 128)
    int xf_mul[4] = {0, 0, 0, 0};
    int xf_div = 1;
    long long k_total() const {
        long long c = 0;
        for (int i = 0; i < n_src; ++i) c += src_c[i];
        return c * n_taps;
    }
    int n_pad() const { return (n + bn - 1) / bn * bn; }
};
int pick_bn(int n);
int pick_oc(int bn, long long num_kb);
int build_gemm(const GemmProblem& pr, GemmParams* out);
int launch_gemm(const GemmParams& p, int bn, cudaStream_t st);
int gemm_init_attrs();

// ------------------------------------------------------------- attention
struct AttnProblem {
    const act_t* q; int q_ld; int q_col0; long long q_rows;
    const act_t* k; int k_ld; int k_col0; long long k_rows;
    const act_t* v; int v_ld; int v_col0;
    act_t* out; int out_ld;
    int B, heads, d;
    int n_blocks, q_blk, k_blk, win;
    const int* shift_dev;  // nullable
};
int build_attn(const AttnProblem& pr, AttnParams* out);
int launch_attn(const AttnParams& p, int d, cudaStream_t st);

// ------------------------------------------------------------ elementwise
int launch_gn_stats(const GnSrc& s, int ns, int rows, double* sums, cudaStream_t st, bool zero_sums = true);
// sums: [ns * nsub][32][2]; the statistics of domain i are the sum of its nsub consecutive slots, taken over
// stat_rows rows in total (0 = rows; differs for nearest-upsampled inputs whose statistics come from the source).
int launch_gn_apply(const GnSrc& s, int ns, int rows, const double* sums, const float* gamma, const float* beta,
                    const float* film, int film_ld, int ns_per_batch, int silu, act_t* y, cudaStream_t st,
                    int nsub = 1, long long stat_rows = 0, const DropState* drop = nullptr, uint32_t drop_site = 0);
// nn.Dropout parameters of a training forward -> device DropState (p == 0 disables); keep mask export for the tests
int launch_set_dropout(DropState* dev, float p, unsigned long long seed, cudaStream_t st);
int launch_dropout_mask(const DropState* dev, uint32_t site, long long elems, unsigned char* keep, cudaStream_t st);
int launch_gn_temporal(const act_t* x, act_t* y, const float* gamma, const float* beta, int B, int F, int P, int C,
                       cudaStream_t st);
int launch_resample(const act_t* x, act_t* y, int mode, int n, int h, int w, int c, cudaStream_t st);
int launch_temporal_attn(const act_t* qkv, act_t* out, int B, int F, int P, int C, int heads, cudaStream_t st);
int launch_pack_weight(const float* w, act_t* dst, int co, int ci, int t, long long ld, long long col_off, cudaStream_t st);

// ------------------------------------------------------------- backward
// Weight gradient of a conv-GEMM: same geometry / sources / taps as the forward GemmProblem; dy is the fp16 output
// gradient [tokens][n]; dw is the packed fp32 gradient [n][ld] (column order of the forward's packed weights), zeroed
// by the caller.
struct WgradProblem {
    ConvGeom g;
    int n_src = 0;
    const act_t* src[GEMM_MAX_SRC] = {};
    int src_c[GEMM_MAX_SRC] = {};
    int n_taps = 1;
    int taps[GEMM_MAX_TAPS][3] = {};
    const act_t* dy = nullptr;
    int n = 0;
    float* dw = nullptr;
    long long ld = 0;
    float* db = nullptr;   // optional fp32 [n] column sums of dy (bias gradient), zeroed by the caller
};
int build_wgrad(const WgradProblem& pr, WgradParams* out, int* n_items);
int launch_wgrad(const WgradParams& p, int n_items, cudaStream_t st);
int launch_unpack_wgrad(const float* dwpk, float* g, int co, int ci, int t, long long ld, long long col_off, float scale,
                        cudaStream_t st, const float* gscale = nullptr);
int launch_pack_weight_t(const float* w, act_t* dst, int co, int ci, int t, int c_lo, int cs, long long ld, long long col_off,
                         cudaStream_t st);
int launch_grad_add(const act_t* x, act_t* y, long long n, int accumulate, cudaStream_t st);
int launch_grad_add2d(const act_t* x, long long ldx, act_t* y, long long ldy, long long rows, int C, int accumulate, cudaStream_t st);

// Attention backward: `fwd` describes the forward call (out = forward output O); gradients land in column ranges of
// row-major fp16 matrices like q / k / v themselves.  lse / delta: [heads][stat_ld] floats (lse from the forward).
struct AttnBwdOut {
    act_t* dq; int dq_ld; int dq_col0;
    act_t* dk; int dk_ld; int dk_col0;
    act_t* dv; int dv_ld; int dv_col0;
};
int build_attn_bwd(const AttnProblem& fwd, const act_t* d_out, int d_out_ld, const float* lse, const float* delta,
                   long long stat_ld, const AttnBwdOut& o, AttnBwdParams* pq, AttnBwdParams* pkv);
int launch_attn_bwd(const AttnBwdParams& pq, const AttnBwdParams& pkv, int d, cudaStream_t st);
int launch_attn_delta(const act_t* d_out, const act_t* out, long long rows, int C, int heads, float* delta, long long delta_ld,
                      cudaStream_t st);
int launch_attn_lse(const AttnParams& p, int d, float* lse, long long lse_ld, cudaStream_t st);   // forward + lse output

// GroupNorm backward (three launches: reduce, apply, finalize).  T: fp32 [ns][C][2] scratch zeroed by the caller.
struct GnBwdProblem {
    GnSrc s;
    int ns = 0, rows = 0;
    const double* sums = nullptr;
    int nsub = 1;
    long long stat_rows = 0;
    const float* gamma = nullptr;
    const float* beta = nullptr;
    const float* film = nullptr;
    int film_ld = 0, ns_per_batch = 1, silu = 0;
    const act_t* dy = nullptr;
    float* T = nullptr;
    GnBwdOut out{};
    float* dgamma = nullptr;
    float* dbeta = nullptr;
    float* dfilm = nullptr;
    const float* gscale = nullptr;
    const DropState* drop = nullptr;   // dropout of the forward output (training): regenerated from (seed, site, index)
    uint32_t drop_site = 0;
};
int launch_gn_bwd(const GnBwdProblem& pr, cudaStream_t st);
int launch_gn_temporal_bwd(const act_t* x, const act_t* dy, act_t* dx, const float* gamma, float* dgamma, float* dbeta, int B,
                           int F, int P, int C, const float* gscale, cudaStream_t st);
int launch_temporal_attn_bwd(const act_t* qkv, const act_t* d_out, act_t* dqkv, int B, int F, int P, int C, int heads,
                             cudaStream_t st);
int launch_resample_bwd(const act_t* dy, act_t* dx, int mode, int n, int h, int w, int c, int accumulate, cudaStream_t st);
// Head adjoints on the tensor-core kernels: G (fp16 [tokens][ldG]) = shifted, scaled copy of dout; dX = G Wt (forward
// implicit-GEMM kernel, token GEMM), dW = G^T A (conv_wgrad_kernel), bias gradient by a strided fp32 reduction.
struct HeadBwdPlan {
    HeadGeom hg{};
    long long tokens = 0;
    int ldG = 0, bn = 0, items = 0;
    act_t* G = nullptr;        // [tokens][ldG] scratch
    float* dwpk = nullptr;     // [ldG][C] fp32 scratch
    GemmParams gemm;           // dX = G * wt^T   (unused when dx == nullptr)
    WgradParams wgrad;
    bool want_dx = false;
};
inline int head_ld(int terms) { return terms <= 64 ? 64 : 128; }
int build_head_bwd(const HeadGeom& hg, const act_t* x, act_t* G, const act_t* wt /*[C][ldG]*/, const float* zero_bias, act_t* dx,
                   float* dwpk, HeadBwdPlan* out);
int run_head_bwd(const HeadBwdPlan& hp, const float* dout, float* dw, float* db, const float* gscale, cudaStream_t st);
int launch_pack_head_t(const float* w, act_t* wt, int n_out, int C, int T, int ldG, cudaStream_t st);

}  // namespace mmd
