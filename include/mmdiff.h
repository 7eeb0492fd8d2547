/* libmmdiff — C ABI of the B200-native MM-Diffusion denoising hot path.
 *
 * The reference (researchmm/MM-Diffusion) has no FFI of its own: its boundary is the Python
 * API.  This header is the C-ABI the Python shim (mm_diffusion_b200/) binds
 * with ctypes; every entry point names the reference interface it stands in for
 * (paths relative to the reference repo).  Conventions: plain pointers and sizes only, every
 * call returns 0 on success or a negative MMD_E* code (message via mmd_last_error()), the
 * caller owns every buffer, device pointers are CUDA device memory of the current device,
 * `stream` is a cudaStream_t passed as void*.  Handles are not thread-safe; distinct handles are.
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef MMDIFF_H_
#define MMDIFF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMD_OK 0
#define MMD_EINVAL (-1)   /* bad argument / unsupported configuration */
#define MMD_ECUDA (-2)    /* CUDA runtime or driver error */
#define MMD_ESTATE (-3)   /* call order violated (e.g. forward before weights are complete) */
#define MMD_ENOTFOUND (-4)

#define MMD_MAX_LEVELS 8

/* Mirror of the constructor arguments of MultimodalUNet
 * (mm_diffusion/multimodal_unet.py:737-764) after create_model's string parsing
 * (mm_diffusion/multimodal_script_util.py:156-201). */
typedef struct MmdConfig {
    int video_f, video_c, video_h, video_w; /* video_size [F,C,H,W] */
    int audio_c, audio_l;                   /* audio_size [C,L] */
    int model_channels;
    int video_out_channels, audio_out_channels;
    int num_res_blocks;
    int n_levels;
    int channel_mult[MMD_MAX_LEVELS];
    int num_heads;          /* heads of the in-ResBlock self-attention (always used, :410-419) */
    int num_head_channels;  /* cross-attention head width, or -1 -> num_heads (:591-598) */
    int n_cross;
    int cross_attention_resolutions[MMD_MAX_LEVELS];
    int cross_attention_windows[MMD_MAX_LEVELS];
    int cross_attention_shift; /* bool */
    int n_video_attn;
    int video_attention_resolutions[MMD_MAX_LEVELS];
    int n_audio_attn;
    int audio_attention_resolutions[MMD_MAX_LEVELS];
    int max_batch;
} MmdConfig;

typedef struct MmdModel MmdModel;

const char* mmd_last_error(void);
const char* mmd_version(void);

/* ---- model lifecycle: stands in for MultimodalUNet.__init__ / load_state_dict
 *      (multimodal_unet.py:737-1012, :1033-1054; key schema SURVEY.md App. F) ---- */
int mmd_model_create(const MmdConfig* cfg, MmdModel** out);
int mmd_model_destroy(MmdModel* m);
/* Parameter inventory in the reference's registration order. */
int mmd_model_num_params(const MmdModel* m);
int mmd_model_param_info(const MmdModel* m, int index, const char** name, int* ndim, int64_t shape[5]);
/* Upload one parameter (fp32, contiguous, device pointer, reference layout [Cout,Cin,k...]).
 * The library repacks into its own fp16 K-major layout; the caller keeps ownership of `data`. */
int mmd_model_set_param(MmdModel* m, const char* name, const float* data, int64_t numel, void* stream);
/* All parameters in one call: `flat` (device, fp32) holds every tensor at the float offset mmd_model_param_offset()
 * reports for it (the layout of the flat gradient buffer of mmd_model_backward); n_floats == mmd_model_param_floats().
 * The Python shim keeps its nn.Parameters as views of such a buffer, so an optimizer step (fp16_util.py:64-74 copies
 * the masters back per tensor in the reference) reaches the library as ONE device copy + one repack graph launch. */
int mmd_model_set_params_flat(MmdModel* m, const float* flat, int64_t n_floats, void* stream);
/* Number of cross-attention blocks that draw a random window shift per forward
 * (CrossAttentionBlock.attention_index, multimodal_unet.py:619-622), in execution order,
 * and the inclusive upper bound F - window of each draw. */
int mmd_model_num_shifts(const MmdModel* m);
int mmd_model_shift_bound(const MmdModel* m, int index);
/* Device bytes the model needs for a batch of `batch` samples (activations + scratch). */
size_t mmd_model_workspace_bytes(const MmdModel* m, int batch);
int mmd_model_num_launches(const MmdModel* m, int batch);

/* ---- MultimodalUNet.forward (multimodal_unet.py:1058-1101) ----
 * video_in  [B,F,C,H,W] fp32, audio_in [B,C,L] fp32, timesteps [B] fp32 (all device);
 * shifts: host array of mmd_model_num_shifts() ints (the values random.randint would have returned);
 * video_out [B,F,Cout,H,W] fp32, audio_out [B,Cout,L] fp32 (device). */
int mmd_model_forward(MmdModel* m, int batch, const float* video_in, const float* audio_in, const float* timesteps,
                      const int32_t* shifts, float* video_out, float* audio_out, void* stream);

/* ---- training / input gradients: MultimodalUNet.forward under autograd (multimodal_gaussian_diffusion.py:1141 training
 *      losses, :815 gradient-guided conditional sampling).  mmd_model_forward_train is mmd_model_forward with every
 *      intermediate kept (no CUDA graph); mmd_model_backward consumes the gradients of its two outputs (fp32, same
 *      layouts) and writes
 *        param_grads  fp32 [mmd_model_param_floats()], parameter i at mmd_model_param_offset(i), reference layout
 *                     [Cout,Cin,k...] (overwritten, not accumulated; may be NULL),
 *        d_video_in / d_audio_in  fp32 gradients wrt the inputs (may be NULL).
 *      One backward per forward_train; activation gradients are fp16 with one internal power-of-two scale. ---- */
int mmd_model_forward_train(MmdModel* m, int batch, const float* video_in, const float* audio_in, const float* timesteps,
                            const int32_t* shifts, float* video_out, float* audio_out, void* stream);
int mmd_model_backward(MmdModel* m, int batch, const float* d_video_out, const float* d_audio_out, float* param_grads,
                       float* d_video_in, float* d_audio_in, void* stream);
/* nn.Dropout of the ResBlock out_layers (multimodal_unet.py:370-386; `--dropout 0.1` in ssh_scripts/multimodal_train.sh:4):
 * probability and 64-bit seed used by the NEXT mmd_model_forward_train calls (p = 0, the default, disables dropout).
 * An element is kept iff Philox4x32-10(seed, site, element index) >= p and scaled by 1 / (1 - p); mmd_model_backward
 * regenerates the same mask.  Inference (mmd_model_forward) never drops. */
int mmd_model_set_dropout(MmdModel* m, float p, uint64_t seed);
/* Training-forward counter of the plan for `batch` (0 = none yet).  A backward belongs to the forward after which this
 * value was read; a later forward at the same batch size overwrites the kept activations (callers compare). */
int64_t mmd_model_train_generation(const MmdModel* m, int batch);
/* Dropout sites of the training plan for `batch` in execution order (per ResBlock: video out_layers, then audio), and
 * the keep mask the LAST training forward used at site `index`: uint8 [rows][channels], channels-last like the
 * activation (video rows = B*F*H*W, audio rows = B*L); modality 0 = video, 1 = audio.  (Parity tests hand the masks to
 * the oracle.) */
int mmd_model_num_dropout_sites(const MmdModel* m, int batch);
int mmd_model_dropout_site(const MmdModel* m, int batch, int index, int64_t* rows, int* channels, int* modality);
int mmd_model_dropout_mask(const MmdModel* m, int batch, int index, unsigned char* keep, void* stream);

size_t mmd_model_train_workspace_bytes(const MmdModel* m, int batch);
int64_t mmd_model_param_offset(const MmdModel* m, int index);
int64_t mmd_model_param_floats(const MmdModel* m);
int mmd_model_num_backward_launches(const MmdModel* m, int batch);
/* measurement hooks of the backward plan (per-step device time, kernel-family tag of a step) */
int mmd_model_profile_backward(MmdModel* m, int batch, int reps, float* ms, int cap, void* stream);
const char* mmd_model_backward_step_kind(const MmdModel* m, int batch, int index);

/* ---- measurement hooks (bench.py): per-launch device time of the plan for `batch` (mean of `reps` un-graphed
 *      executions, CUDA events on `stream`) and each step's kernel family / algorithmic FLOPs / bytes
 *      (DESIGN.md §kernels).  mmd_model_profile returns the step count (>0) or a negative error. ---- */
int mmd_model_profile(MmdModel* m, int batch, int reps, float* ms, int cap, void* stream);
int mmd_model_step_info(const MmdModel* m, int batch, int index, const char** kind, double* flops, double* bytes,
                        int* kernels);

/* ---- sampler tail: GaussianDiffusion.p_sample after the model call
 *      (multimodal_gaussian_diffusion.py:292-350, :453-470): per element
 *      x0 = clip(a*x - b*eps); mean = c1*x0 + c2*x; sample = mean + nz*sigma*z.
 *      coef: device [B][6] fp32 = {a,b,c1,c2,sigma,nz}; pred_xstart may be NULL. ---- */
int mmd_p_sample_tail(const float* x, const float* eps, const float* noise, const float* coef, int batch,
                      int64_t per_sample, int clip_denoised, float* sample, float* pred_xstart, void* stream);
/* GaussianDiffusion.q_sample (multimodal_gaussian_diffusion.py:187-205); coef device [B][2]. */
int mmd_q_sample(const float* x_start, const float* noise, const float* coef, int batch, int64_t per_sample, float* out,
                 void* stream);

/* ---- sample epilogue of the sampling scripts (py_scripts/multimodal_sample_sr.py:159-163):
 *      video fp32 [n_images][channels][hw] in [-1, 1] -> uint8 [n_images][hw][channels] =
 *      ((x + 1) * 127.5).clamp(0, 255).to(uint8) with the permute(0, 1, 3, 4, 2) folded in (n_images = batch x frames). ---- */
int mmd_sample_epilogue(const float* video, unsigned char* out, int64_t n_images, int channels, int64_t hw, void* stream);

/* ---- DPM-Solver state arithmetic (DPM_Solver, multimodal_dpm_solver_plus.py:373-1298) ----
 * Every solver update and the eps -> x0 conversion is a linear combination of at most four fp32 tensors with
 * step-wide scalar coefficients (:532-1036, :419-430): out = sum_i coef[i] * src[i].  src: host array of device
 * pointers; coef: host array.  out may alias a source. */
int mmd_lincomb(int n_terms, const float* const* src, const float* coef, int64_t numel, float* out, void* stream);
/* Dynamic thresholding tail (:431-438): x0 = clamp(x0, -s[b], s[b]) / (s[b] / max_val), in place; s device [B]. */
int mmd_dpm_threshold(float* x0, const float* s, int batch, int64_t per_sample, float max_val, void* stream);
/* Adaptive step-size error (:1134-1138): out[b] = sum_i ((hi-lo) / max(atol, rtol*max(|lo|,|prev|)))^2, device double [B]. */
int mmd_dpm_error_sq(const float* hi, const float* lo, const float* prev, int batch, int64_t per_sample, float atol,
                     float rtol, double* out, void* stream);

/* ---- operator-level entry points (used by the parity tests; same kernels the model plan launches) ---- */

/* GroupNorm32 (+SiLU, +FiLM) on channels-last fp16: x [ns*rows, c1(+c2)] -> y.  nn.py:16-33. */
int mmd_op_group_norm(const void* x1, int c1, const void* x2, int c2, int ns, int rows, const float* gamma,
                      const float* beta, const float* film, int film_ld, int ns_per_batch, int silu, void* y,
                      void* stream);
int mmd_op_group_norm_temporal(const void* x, void* y, const float* gamma, const float* beta, int B, int F, int P, int C,
                               void* stream);
/* mode 0 video avg-pool(1,2,2), 1 audio avg-pool(4), 2 video nearest x2, 3 audio nearest x4. */
int mmd_op_resample(const void* x, void* y, int mode, int n, int h, int w, int c, void* stream);

/* Implicit-GEMM convolution (VideoConv / AudioConv, multimodal_unet.py:68-131) on channels-last fp16.
 * geometry: rank-1 token coordinates (innermost first) with extents dims[] and box[] (product 128);
 * taps: n_taps x 3 coordinate deltas; sources are concatenated along channels.
 * weight: fp32 [n][c_total][n_taps] (reference layout, flattened kernel dims), bias fp32 [n].
 * out: fp16 [tokens][n] (out_f32 == NULL) or fp32 scatter with strides (heads).
 * gn_sums (optional): GroupNorm(32) statistics of the output, reduced in the GEMM epilogue and ADDED to
 *   double [domains][32][2] = (sum, sum of squares) per (domain, group); the caller zeroes it.  gn_rows = tokens of
 *   one domain.  rank 2: domain = token / gn_rows (gn_rows 64 or a multiple of 128); rank 3 (L,B): one domain per
 *   sample (gn_rows == dims[0]); rank 4 (P,F,B): one domain per (b,f) (gn_rows == dims[0] >= 64).  Needs n % 128 == 0.
 *   (normalization() statistics of multimodal_unet.py's GroupNorm32 consumers, nn.py:93-100.) */
typedef struct MmdConvDesc {
    int rank;          /* 2..5 including the channel coordinate */
    int64_t dims[4];
    int box[4];
    int n_src;
    const void* src[4];
    int src_channels[4];
    int n_taps;
    int taps[27][3];
    const float* weight;
    const float* bias;
    int n;
    void* out;
    float* out_f32;
    int64_t ostride[4];
    int64_t ostride_c;
    double* gn_sums;
    int64_t gn_rows;
} MmdConvDesc;
int mmd_op_conv(const MmdConvDesc* d, void* stream);

/* Measurement entry (tools/gpu_gemm_micro.py): mmd_op_conv, then the same packed problem launched `reps` more times back to
   back on `stream` between two CUDA events; *us_per_launch = average device time per launch in microseconds (weight
   packing excluded).  Blocks until the launches are done.  No reference counterpart. */
int mmd_op_conv_timed(const MmdConvDesc* d, int reps, float* us_per_launch, void* stream);
/* Pointwise convolution (n_taps == 1) whose source 0 is first normalised: conv(act(GroupNorm32(src0) * (1 + scale) +
 * shift) ++ src1..) — the ResBlock out_layers (multimodal_unet.py:459-470) and the attention norms (:284, :664) with
 * the GroupNorm apply done on the GEMM's A operand in shared memory (no normalised tensor in HBM).  ns domains:
 * rank 2: tokens / ns consecutive rows each (64 or a multiple of 128); rank 3 (L,B): one per sample (ns == dims[1]).
 * film: [ns / ns_per_batch][film_ld] with scale at [0,C) and shift at [C,2C), or NULL. */
int mmd_op_conv_gn(const MmdConvDesc* d, const float* gamma, const float* beta, const float* film, int film_ld, int ns,
                   int ns_per_batch, int silu, void* stream);

/* Attention core.  q/k/v are column ranges of row-major fp16 matrices.
 * Query block i of sample b attends key blocks (i+shift+j) mod n_blocks, j<win
 * (QKVAttention.forward multimodal_unet.py:507-564; SingleModalQKVAttention :221-240 with win=1, shift=0). */
typedef struct MmdAttnDesc {
    const void* q; int q_ld; int q_col0; int64_t q_rows;
    const void* k; int k_ld; int k_col0; int64_t k_rows;
    const void* v; int v_ld; int v_col0;
    void* out; int out_ld;
    int batch, heads, head_dim;
    int n_blocks, q_blk, k_blk, win, shift;
} MmdAttnDesc;
int mmd_op_attention(const MmdAttnDesc* d, void* stream);
/* 16-token temporal attention: qkv [B,F,P,3C] -> out [B,F,P,C]. */
int mmd_op_temporal_attention(const void* qkv, void* out, int B, int F, int P, int C, int heads, void* stream);

/* ---- operator-level backward entry points (training; parity tests of the backward
 *      kernels against torch.autograd of the reference ops; same kernels the model's backward plan launches) ---- */

/* Weight / bias gradient of mmd_op_conv: same descriptor (geometry, sources, taps, n; weight/bias/out unused);
 * dy fp16 [tokens][n]; dweight fp32 [n][c_total][n_taps] and dbias fp32 [n] are ACCUMULATED into.
 * (autograd of F.conv1d/2d/3d in VideoConv / AudioConv, multimodal_unet.py:68-131.) */
int mmd_op_conv_wgrad(const MmdConvDesc* d, const void* dy, float* dweight, float* dbias, void* stream);
/* Data gradient wrt source `src_index`: dx fp16 [tokens][src_channels[src_index]] = conv of dy with the transposed,
 * tap-mirrored weights (runs on the forward implicit-GEMM kernel).  Needs n % 64 == 0. */
int mmd_op_conv_dgrad(const MmdConvDesc* d, const void* dy, int src_index, void* dx, void* stream);
/* GroupNorm32 (+FiLM, +SiLU) backward of mmd_op_group_norm (nn.py:16-33, multimodal_unet.py:459-470): dy fp16
 * [ns*rows][c1+c2]; dx1 [..][c1], dx2 [..][c2] (may be NULL when c2 == 0); dgamma/dbeta [C] and dfilm [B][film_ld]
 * (scale grads at [c], shift grads at [C + c]) are ACCUMULATED into (dfilm may be NULL without film). */
int mmd_op_group_norm_bwd(const void* x1, int c1, const void* x2, int c2, int ns, int rows, const float* gamma, const float* beta,
                          const float* film, int film_ld, int ns_per_batch, int silu, const void* dy, void* dx1, void* dx2,
                          float* dgamma, float* dbeta, float* dfilm, void* stream);
int mmd_op_group_norm_temporal_bwd(const void* x, const void* dy, void* dx, const float* gamma, float* dgamma, float* dbeta, int B,
                                   int F, int P, int C, void* stream);
/* Adjoint of mmd_op_resample; (n, h, w, c) are the FORWARD INPUT extents, dy has the forward output shape. */
int mmd_op_resample_bwd(const void* dy, void* dx, int mode, int n, int h, int w, int c, void* stream);
int mmd_op_temporal_attention_bwd(const void* qkv, const void* d_out, void* dqkv, int B, int F, int P, int C, int heads, void* stream);
/* Attention forward (writes d->out and lse fp32 [heads][q_rows]) followed by its backward: d_out fp16 [q_rows][heads*d];
 * dq / dk / dv are written into column ranges (start dq_col0 / dk_col0 / dv_col0, head h at + h*d) of row-major fp16
 * matrices with leading dimension grad_ld (QKVAttention / SingleModalQKVAttention autograd, multimodal_unet.py:212-244, 507-564). */
int mmd_op_attention_fwd_bwd(const MmdAttnDesc* d, const void* d_out, float* lse, void* dq, void* dk, void* dv, int grad_ld,
                             int dq_col0, int dk_col0, int dv_col0, void* stream);
/* Narrow output heads (video_out / audio_out, multimodal_unet.py:1003-1012): descriptor as for the forward head call
 * (fp32 strided output layout); dout fp32 in that layout; dx fp16 [tokens][C] (NULL = skip); dweight fp32
 * [n][C][n_taps] / dbias [n] ACCUMULATED into (NULL = skip). */
int mmd_op_head_bwd(const MmdConvDesc* d, const float* dout, void* dx, float* dweight, float* dbias, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMDIFF_H_ */
