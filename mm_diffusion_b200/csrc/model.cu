// MultimodalUNet as a static launch plan over the sm_100a kernels (C-ABI: mmd_model_*).
//
// Restates the reference's model assembly (mm_diffusion/multimodal_unet.py:737-1012 constructor,
// :1058-1101 forward, :434-495 ResBlock, :655-678 CrossAttentionBlock) as one topology walk that
//   (create)  registers every parameter under the reference's name / shape / order
//             and lays out the repacked fp16 weights, and
//   (plan)    emits, for a batch size, the ordered list of kernel launches on channels-last fp16
//             activations [B,F,H,W,C] / [B,L,C].
// The plan is captured into a CUDA graph; per-call inputs (x, t, window shifts) go through
// fixed staging buffers so the graph is replayable.
#include <algorithm>
#include <functional>
#include <map>
#include <memory>
#include <unordered_map>

#include <nvtx3/nvToolsExt.h>

#include "host.cuh"

namespace mmd {

// NVTX range per plan step (named by kernel family) on the eager paths (MMD_NO_GRAPH=1 / profiling hooks): nsys and ncu
// --nvtx captures can filter by family.  Graph replays carry no ranges (the graph is one node to the tools).
struct NvtxStep {
    static bool on() { static const bool v = [] { const char* e = getenv("MMD_NVTX"); return !(e && e[0] == '0'); }(); return v; }
    explicit NvtxStep(const char* name) { if (on()) nvtxRangePushA(name); }
    ~NvtxStep() { if (on()) nvtxRangePop(); }
};

struct ShiftArgs { int n; int v[64]; };
__global__ void set_shifts_kernel(ShiftArgs a, int* dst) {
    if (threadIdx.x < a.n) dst[threadIdx.x] = a.v[threadIdx.x];
}

using LaunchFn = std::function<int(cudaStream_t)>;

struct ParamInfo {
    std::string name;
    std::vector<int64_t> shape;
    size_t offset = 0;  // floats, into the fp32 arena
    int64_t numel = 0;
    bool set = false;
};

struct Seg { int w; int ci; int T; };   // one weight tensor of a packed GEMM: param index, input channels, taps

struct PackedConv {
    std::vector<Seg> segs;        // conv segments in K order (identity columns follow them)
    std::vector<int> biases;      // bias parameters summed into the packed bias
    size_t w_off = 0;     // halves, into packed arena
    size_t b_off = 0;     // floats, into packed bias arena
    int n = 0, n_pad = 0, bn = 128;
    long long k_total = 0;
    int identity_c = 0;   // trailing identity columns (residual folded into the GEMM)
};

struct Arena {
    uint8_t* base = nullptr;
    size_t cap = 0, top = 0, peak = 0;
    uintptr_t fake = 0;   // dry runs hand out distinct non-null fake addresses per arena (the backward tape keys on them)
    void* take(size_t bytes) {
        size_t at = (top + 1023) & ~size_t(1023);
        top = at + bytes;
        peak = std::max(peak, top);
        return base ? base + at : reinterpret_cast<void*>(fake + at);  // dry run when base == nullptr
    }
};

struct StepInfo {
    std::string kind;     // kernel family tag
    double flops = 0;     // algorithmic FLOPs (multiply-add = 2)
    double bytes = 0;     // algorithmic bytes (inputs + outputs + weights, fp16 activations)
    int kernels = 1;      // kernel launches inside the step
    int stream = 0;       // 0 = video / main branch, 1 = audio branch, -1 = cross-stream sync point
};

struct Plan {
    int B = 0;
    std::vector<LaunchFn> steps;
    std::vector<StepInfo> info;
    uint8_t* ws = nullptr;
    size_t ws_bytes = 0;
    float *in_video = nullptr, *in_audio = nullptr, *t_dev = nullptr, *out_video = nullptr, *out_audio = nullptr;
    int* shifts_dev = nullptr;
    cudaGraphExec_t graph = nullptr;
    cudaGraphExec_t bwd_graph = nullptr;   // training plans: the backward launch list
    std::vector<cudaEvent_t> events;   // fork / join events of the two-branch capture
    // ---- training plan only (every intermediate kept; backward launch list built from the tape)
    bool train = false;
    std::vector<LaunchFn> bwd_steps;
    std::vector<std::string> bwd_kind;   // kernel family tag per backward step (profiling)
    std::vector<LaunchFn> tpack_ops;   // transposed weight packs for the data gradients (re-run when weights change)
    bool tpack_dirty = true;
    cudaGraphExec_t tpack_graph = nullptr;   // the same ops as one multi-branch graph (one launch per parameter update)
    float *d_out_video = nullptr, *d_out_audio = nullptr;   // staged output gradients (fp32, API layout)
    float *d_in_video = nullptr, *d_in_audio = nullptr;     // input gradients (fp32, API layout)
    float* gscale = nullptr;            // {s, 1/s}
    unsigned int* amax_bits = nullptr;
    float* g32 = nullptr;               // parameter gradients, same offsets as MmdModel::w32
    float* d_emb_all = nullptr;         // [B][emb_rows]
    act_t* wpk_t = nullptr;             // transposed packs
    float* zero_bias = nullptr;
    bool fwd_done = false;
    long long generation = 0;           // bumped by every training forward: a backward must match the forward it follows
    // nn.Dropout of the ResBlock out_layers (multimodal_unet.py:376,384): parameters of the current training forward in
    // device memory (read by the captured graphs), and the sites in execution order (tests export their masks)
    DropState* drop_dev = nullptr;
    struct DropSite { uint32_t site; long long rows; int C; int modality; };
    std::vector<DropSite> drop_sites;
    ~Plan() {
        if (graph) cudaGraphExecDestroy(graph);
        if (bwd_graph) cudaGraphExecDestroy(bwd_graph);
        if (tpack_graph) cudaGraphExecDestroy(tpack_graph);
        for (auto e : events) cudaEventDestroy(e);
        if (ws) cudaFree(ws);
        if (wpk_t) cudaFree(wpk_t);
    }
};

// One forward launch as the backward needs to see it.
struct TapeOp {
    enum Kind { GEMM, GN, GN_TEMPORAL, TATTN, ATTN, RESAMPLE } kind = GEMM;
    // GEMM
    std::string tag;
    ConvGeom g;
    std::vector<std::pair<const act_t*, int>> srcs;
    std::vector<std::array<int, 3>> taps;
    const PackedConv* pc = nullptr;
    act_t* out = nullptr;
    float* out_f32 = nullptr;
    long long ostride[4] = {0, 0, 0, 0};
    long long ostride_c = 0;
    int stem = 0;             // 1 video im2col stem, 2 audio im2col stem (source is the im2col buffer)
    // GN
    const act_t *x1 = nullptr, *x2 = nullptr;
    int c1 = 0, c2 = 0, ns = 0, rows = 0, gn_g = -1, gn_b = -1, emb_row0 = -1, silu = 0, nsub = 1, ns_per_batch = 1;
    const double* sums = nullptr;
    long long stat_rows = 0;
    act_t* y = nullptr;
    int drop_site = -1;       // >= 0: the forward applied dropout to y (site id of the Philox stream)
    // GN_TEMPORAL / TATTN / RESAMPLE
    const act_t* in = nullptr;
    int B = 0, F = 0, P = 0, C = 0, heads = 0, mode = 0, n_ = 0, h_ = 0, w_ = 0;
    // ATTN
    AttnProblem ap{};
    float* lse = nullptr;
};

}  // namespace mmd

using namespace mmd;

struct MmdModel {
    MmdConfig cfg{};
    std::vector<ParamInfo> params;
    std::unordered_map<std::string, int> param_index;
    float* w32 = nullptr;      // fp32 copies of all parameters
    size_t w32_floats = 0;
    std::map<std::string, PackedConv> packs;
    std::vector<LaunchFn> pack_ops;
    cudaGraphExec_t pack_graph = nullptr;   // pack_ops as one multi-branch graph: a parameter update costs one launch
    act_t* wpk = nullptr;      // packed fp16 weights
    size_t wpk_halves = 0;
    float* bpk = nullptr;      // packed fp32 biases + stacked emb weights
    size_t bpk_floats = 0;
    bool dirty = true;
    float drop_p = 0.f;                    // dropout of the next training forward (mmd_model_set_dropout)
    unsigned long long drop_seed = 0;
    std::vector<int> shift_bounds;  // per cross block in execution order; -1 = no draw
    int emb_rows = 0;               // stacked emb_layers rows
    size_t emb_w_off = 0, emb_b_off = 0;
    std::map<int, std::unique_ptr<Plan>> plans;
    std::map<int, std::unique_ptr<Plan>> train_plans;
    bool use_graph = true;
    bool two_streams = true;   // MMD_ONE_STREAM=1 captures everything on one stream
    cudaStream_t cap_stream = nullptr, cap_stream2 = nullptr;
    ~MmdModel() {
        plans.clear();
        train_plans.clear();
        if (pack_graph) cudaGraphExecDestroy(pack_graph);
        if (w32) cudaFree(w32);
        if (wpk) cudaFree(wpk);
        if (bpk) cudaFree(bpk);
        if (cap_stream) cudaStreamDestroy(cap_stream);
        if (cap_stream2) cudaStreamDestroy(cap_stream2);
    }
};

namespace mmd {

// ---------------------------------------------------------------------------
// Topology walk.  mode CREATE: registers params + packed layouts (no device work).
//                 mode PLAN  : emits launches into plan (weights/arenas must exist).
// ---------------------------------------------------------------------------
// Fused GroupNorm statistics attached to a tensor by its producer: `slots` [B*nsub][32][2] doubles
// (video: one slot per frame, nsub = F; audio: one per sample), `rows` = rows per sample the sums cover.
struct Stat {
    double* slots = nullptr; int nsub = 1; long long rows = 0;
    double* qslots = nullptr;   // per-4-channel partial sums of the same slots ([slot][C / 4][2]) for concat GroupNorms
};
struct VT { act_t* p; int C, H, W; Stat st; };   // video [B,F,H,W,C]
struct AT { act_t* p; int C, L; Stat st; };      // audio [B,L,C]
// A GroupNorm whose apply is deferred into the A path of the pointwise GEMM that consumes it (inference plans):
// kind 1 = rank-2 token geometry (domain = token / rows), 3 = audio geometry (L,B) (domain = sample).
struct XfSpec {
    bool on = false;
    int kind = 0;
    const double* sums = nullptr;
    int nsub = 1;
    long long stat_rows = 0;
    int gn_g = -1, gn_b = -1;
    const float* film = nullptr;
    int ns_per_batch = 1, silu = 0, ns = 0, rows = 0;
};

struct Walker {
    MmdModel& m;
    bool create;
    Plan* plan = nullptr;
    int B = 1;
    Arena persist, scratch, scratch_a, stats;   // scratch per branch (video / audio run concurrently); stats: GroupNorm
                                                // accumulators, zeroed once per forward
    // training: nothing is recycled (scratch marks are ignored), every launch is taped, activation gradients live in
    // `grads`, per-op backward temporaries in `bscratch` (reset per op: the backward runs on one stream)
    bool train = false;
    Arena grads, bscratch, tpack;
    std::vector<TapeOp> tape;
    std::map<const act_t*, std::pair<act_t*, bool>> grad_map;   // activation -> (gradient buffer, written yet)
    struct EmbBlk { int w, b, row0, rows; };
    std::vector<EmbBlk> emb_blocks;
    int te_idx[4] = {-1, -1, -1, -1};
    void release(Arena& a, size_t mark) { if (!train) a.top = mark; }
    int cur = 0;                                // branch being emitted: 0 video (main stream), 1 audio
    Arena& S() { return cur == 1 ? scratch_a : scratch; }
    int err = MMD_OK;
    size_t w32_top = 0, wpk_top = 0, bpk_top = 0;
    int shift_slot = 0;
    int drop_site_top = 0;
    int emb_row_top = 0;
    bool wide_n = true;         // MMD_NO_BN256=1 keeps 128-wide GEMM tiles
    bool fuse_stats = true;     // MMD_NO_FUSED_STATS=1 keeps every GroupNorm on the standalone statistics kernel
    int xf_mask = 0;            // MMD_XF bitmask: GroupNorm apply folded into the consumer GEMM's A path for
                                // 1 = ResBlock out_layers, 2 = self-attention norms, 4 = cross-attention norms.
                                // Off by default: measured on B200 the standalone apply pass it removes (3.2 -> 1.7 ms of
                                // ungraphed kernel time) is more than paid back by the slower consumer GEMMs inside the
                                // two-branch step graph (12.52 ms -> 13.02 ms per step with every family on); DESIGN.md §6.
    float* emb_all = nullptr;   // [B][emb_rows]
    float* silu_emb = nullptr;  // [B][E]

    Walker(MmdModel& mm, bool c) : m(mm), create(c) {
        persist.fake = uintptr_t(1) << 40; scratch.fake = uintptr_t(2) << 40; scratch_a.fake = uintptr_t(3) << 40;
        stats.fake = uintptr_t(4) << 40; grads.fake = uintptr_t(5) << 40; bscratch.fake = uintptr_t(6) << 40;
        tpack.fake = uintptr_t(7) << 40;
        const char* e = getenv("MMD_NO_FUSED_STATS");
        fuse_stats = !(e && e[0] == '1');
        const char* w = getenv("MMD_NO_BN256");
        wide_n = !(w && w[0] == '1');
        const char* x = getenv("MMD_XF");
        if (x) xf_mask = atoi(x);
        const char* qs = getenv("MMD_NO_QUAD_STATS");
        quad_stats = fuse_stats && !(qs && qs[0] == '1');
    }
    double* stat_slots_video() { return static_cast<double*>(stats.take(sizeof(double) * 64 * B * F())); }
    double* stat_slots_audio() { return static_cast<double*>(stats.take(sizeof(double) * 64 * B)); }
    // block outputs also leave per-4-channel partial sums: the decoder's skip-concat GroupNorms fold their statistics
    // from the two producers instead of re-reading both tensors (MMD_NO_QUAD_STATS=1 keeps the standalone pass)
    bool quad_stats = true;
    double* quad_slots_video(int C) { return quad_stats ? static_cast<double*>(stats.take(sizeof(double) * 2 * (C / 4) * B * F())) : nullptr; }
    double* quad_slots_audio(int C) { return quad_stats ? static_cast<double*>(stats.take(sizeof(double) * 2 * (C / 4) * B)) : nullptr; }
    // fused statistics need the 64- or 128-row domain runs the GEMM epilogue reduces over
    // (the epilogue reduction works on 128-column half tiles and 4-channel quads: C % 128 == 0)
    bool can_fuse_video(int hw, int c) const { return fuse_stats && hw >= 64 && c % 128 == 0; }
    bool can_fuse_audio(int L, int c) const { return fuse_stats && L >= 128 && c % 128 == 0; }   // one sample per 128-row tile
    const MmdConfig& cfg() const { return m.cfg; }
    int F() const { return m.cfg.video_f; }

    bool bad() const { return err != MMD_OK; }
    void set_err(int e) { if (err == MMD_OK) err = e; }

    // ---------------- parameters
    int reg(const std::string& name, std::vector<int64_t> shape) {
        if (create) {
            ParamInfo pi;
            pi.name = name;
            pi.shape = shape;
            pi.numel = 1;
            for (auto s : shape) pi.numel *= s;
            pi.offset = w32_top;
            w32_top += (pi.numel + 3) & ~size_t(3);
            m.param_index[name] = static_cast<int>(m.params.size());
            m.params.push_back(pi);
            return static_cast<int>(m.params.size()) - 1;
        }
        auto it = m.param_index.find(name);
        if (it == m.param_index.end()) { set_err(fail(MMD_ENOTFOUND, "internal: param %s", name.c_str())); return 0; }
        return it->second;
    }
    const float* pf(int idx) const { return m.w32 + m.params[idx].offset; }
    struct ConvP { int w, b; };
    ConvP reg_conv(const std::string& p, int o, int i, std::vector<int64_t> k) {
        std::vector<int64_t> ws = {o, i};
        ws.insert(ws.end(), k.begin(), k.end());
        ConvP r;
        r.w = reg(p + ".weight", ws);
        r.b = reg(p + ".bias", {o});
        return r;
    }
    struct GnP { int g, b; };
    GnP reg_gn(const std::string& p, int c) {
        GnP r;
        r.g = reg(p + ".GroupNorm.weight", {c});
        r.b = reg(p + ".GroupNorm.bias", {c});
        return r;
    }

    // ---------------- packed weights
    const PackedConv* pack(const std::string& key, int n, const std::vector<Seg>& segs, int identity_c,
                           const std::vector<int>& biases, int force_bn = 0, long long min_k = 0) {
        if (!create) {
            auto it = m.packs.find(key);
            if (it == m.packs.end()) { set_err(fail(MMD_ENOTFOUND, "internal: pack %s", key.c_str())); return nullptr; }
            return &it->second;
        }
        PackedConv pc;
        pc.segs = segs;
        pc.biases = biases;
        pc.n = n;
        pc.bn = force_bn ? force_bn : pick_bn(n);
        pc.n_pad = (n + pc.bn - 1) / pc.bn * pc.bn;
        long long k = 0;
        for (auto& s : segs) k += static_cast<long long>(s.ci) * s.T;
        k += identity_c;
        k = std::max(k, min_k);
        pc.k_total = k;
        pc.identity_c = identity_c;
        pc.w_off = wpk_top;
        wpk_top += (static_cast<size_t>(pc.n_pad) * k + 63) & ~size_t(63);
        pc.b_off = bpk_top;
        bpk_top += (pc.n_pad + 63) & ~size_t(63);
        MmdModel* mp = &m;
        std::vector<Seg> segs_c = segs;
        std::vector<int> bias_c = biases;
        m.pack_ops.push_back([mp, pc, segs_c, identity_c, bias_c](cudaStream_t st) -> int {
            act_t* dst = mp->wpk + pc.w_off;
            long long col = 0;
            for (auto& s : segs_c) {
                MMD_TRY(launch_pack_weight(mp->w32 + mp->params[s.w].offset, dst, pc.n, s.ci, s.T, pc.k_total, col, st));
                col += static_cast<long long>(s.ci) * s.T;
            }
            if (identity_c > 0) {
                pack_identity_kernel<<<(identity_c + 255) / 256, 256, 0, st>>>(dst, identity_c, pc.k_total, col);
                MMD_CUDA_OK(cudaGetLastError());
            }
            float* bdst = mp->bpk + pc.b_off;
            MMD_CUDA_OK(cudaMemsetAsync(bdst, 0, sizeof(float) * pc.n_pad, st));
            for (int bi : bias_c) {
                add_vec_kernel<<<(pc.n + 255) / 256, 256, 0, st>>>(bdst, mp->w32 + mp->params[bi].offset, pc.n);
                MMD_CUDA_OK(cudaGetLastError());
            }
            return MMD_OK;
        });
        m.packs[key] = pc;
        return &m.packs[key];
    }

    // tap-major pack of a narrow conv head: [128 = T x 4 rows][ci] fp16 (pack_head_taps_kernel), no bias (the gather adds it)
    const PackedConv* pack_head_taps(const std::string& key, int w_param, int co, int ci, int T) {
        if (!create) {
            auto it = m.packs.find(key);
            if (it == m.packs.end()) { set_err(fail(MMD_ENOTFOUND, "internal: pack %s", key.c_str())); return nullptr; }
            return &it->second;
        }
        if (T * 4 > 128) { set_err(fail(MMD_EINVAL, "head taps %d", T)); return nullptr; }
        PackedConv pc;
        pc.segs = {};
        pc.biases = {};
        pc.n = 128;
        pc.bn = 128;
        pc.n_pad = 128;
        pc.k_total = ci;
        pc.identity_c = 0;
        pc.w_off = wpk_top;
        wpk_top += (static_cast<size_t>(128) * ci + 63) & ~size_t(63);
        pc.b_off = bpk_top;
        bpk_top += 128;
        MmdModel* mp = &m;
        m.pack_ops.push_back([mp, pc, w_param, co, ci, T](cudaStream_t st) -> int {
            pack_head_taps_kernel<<<(128 * ci + 255) / 256, 256, 0, st>>>(mp->w32 + mp->params[w_param].offset, mp->wpk + pc.w_off, co, ci, T);
            MMD_CUDA_OK(cudaGetLastError());
            MMD_CUDA_OK(cudaMemsetAsync(mp->bpk + pc.b_off, 0, sizeof(float) * 128, st));
            return MMD_OK;
        });
        m.packs[key] = pc;
        return &m.packs[key];
    }
    static bool head_taps_on() {
        static const bool on = [] { const char* e = getenv("MMD_HEAD_TAPS"); return !(e && e[0] == '0'); }();
        return on;
    }

    // ---------------- activations
    act_t* alloc_p(size_t elems) { return static_cast<act_t*>(persist.take(elems * sizeof(act_t))); }
    act_t* alloc_s(size_t elems) { return static_cast<act_t*>(S().take(elems * sizeof(act_t))); }
    void* alloc_s_bytes(size_t bytes) { return S().take(bytes); }
    size_t vtok(const VT& v) const { return static_cast<size_t>(B) * F() * v.H * v.W; }
    size_t atok(const AT& a) const { return static_cast<size_t>(B) * a.L; }
    bool emitting() const { return plan != nullptr && persist.base != nullptr; }

    void push(LaunchFn fn, const char* kind, double flops, double bytes, int kernels = 1) {
        if (!emitting()) return;
        plan->steps.push_back(std::move(fn));
        StepInfo si;
        si.kind = kind; si.flops = flops; si.bytes = bytes; si.kernels = kernels; si.stream = cur;
        plan->info.push_back(si);
    }
    // both branches wait for each other (graph edges; a no-op when the plan runs on one stream)
    void sync_branches() {
        if (!emitting()) return;
        plan->steps.push_back([](cudaStream_t) { return MMD_OK; });
        StepInfo si;
        si.kind = "sync"; si.kernels = 0; si.stream = -1;
        plan->info.push_back(si);
    }

    // ---------------- emitters
    // stat_kind: 0 none, 1 video tokens (2-D geometry, one domain per frame of `stat_hw` rows),
    //            2 video temporal geometry (P,F,B), 3 audio geometry (L,B)
    void emit_gemm(const char* tag, const ConvGeom& g, const std::vector<std::pair<const act_t*, int>>& srcs,
                   const std::vector<std::array<int, 3>>& taps, const PackedConv* pc, act_t* out, float* out_f32 = nullptr,
                   const long long* ostride = nullptr, long long ostride_c = 0, double* stat_slots = nullptr,
                   int stat_kind = 0, int stat_hw = 0, int stem = 0, const XfSpec* xf = nullptr, double* stat_q = nullptr) {
        if (train && pc) {
            TapeOp op;
            op.kind = TapeOp::GEMM;
            op.tag = tag; op.g = g; op.srcs = srcs; op.taps = taps; op.pc = pc; op.out = out; op.out_f32 = out_f32;
            if (ostride) for (int i = 0; i < 4; ++i) op.ostride[i] = ostride[i];
            op.ostride_c = ostride_c;
            op.stem = stem;
            tape.push_back(op);
        }
        if (!emitting() || bad() || !pc) return;
        GemmProblem pr;
        pr.g = g;
        pr.n_src = static_cast<int>(srcs.size());
        long long ctot = 0;
        for (size_t i = 0; i < srcs.size(); ++i) { pr.src[i] = srcs[i].first; pr.src_c[i] = srcs[i].second; ctot += srcs[i].second; }
        pr.n_taps = static_cast<int>(taps.size());
        for (size_t t = 0; t < taps.size(); ++t) for (int j = 0; j < 3; ++j) pr.taps[t][j] = taps[t][j];
        if (ctot * pr.n_taps != pc->k_total) { set_err(fail(MMD_EINVAL, "internal: K mismatch %lld vs %lld", ctot * pr.n_taps, pc->k_total)); return; }
        pr.w = m.wpk + pc->w_off;
        pr.bias = m.bpk + pc->b_off;
        pr.n = pc->n;
        pr.bn = pc->bn;
        // wide-N tiles halve the A re-reads and the shared-memory traffic per FLOP; only when the tile count
        // still fills the machine a couple of times over
        {
            long long mt = 1;
            for (int i = 0; i < 4; ++i) mt *= (g.dims[i] + g.box[i] - 1) / g.box[i];
            if (wide_n && pc->bn == 128 && pc->n % 256 == 0 && mt * (pc->n / 256) >= 2LL * num_sms()) pr.bn = 256;
        }
        pr.out = out;
        pr.out_f32 = out_f32;
        if (ostride) for (int i = 0; i < 4; ++i) pr.ostride[i] = ostride[i];
        pr.ostride_c = ostride_c;
        if (stat_slots && stat_kind == 1) {
            pr.stats = stat_slots; pr.stats_rows = std::min(stat_hw, 128); pr.stats_mul[0] = 1; pr.stats_div = stat_hw;
        } else if (stat_slots && stat_kind == 2) {
            pr.stats = stat_slots; pr.stats_rows = g.box[0]; pr.stats_mul[1] = 1; pr.stats_mul[2] = F(); pr.stats_div = 1;
        } else if (stat_slots && stat_kind == 3) {
            pr.stats = stat_slots; pr.stats_rows = 128; pr.stats_mul[1] = 1; pr.stats_div = 1; pr.stats_valid_coord = 0;
        }
        if (pr.stats) pr.stats_q = stat_q;
        if (xf && xf->on) {
            pr.xf_sums = xf->sums; pr.xf_gamma = pf(xf->gn_g); pr.xf_beta = pf(xf->gn_b); pr.xf_film = xf->film;
            pr.xf_film_ld = m.emb_rows; pr.xf_dom_per_batch = xf->ns_per_batch; pr.xf_nsub = xf->nsub; pr.xf_silu = xf->silu;
            pr.xf_stat_rows = xf->stat_rows;
            if (xf->kind == 1) { pr.xf_rows = std::min(xf->rows, 128); pr.xf_mul[0] = 1; pr.xf_div = xf->rows; }
            else { pr.xf_rows = 128; pr.xf_mul[1] = 1; pr.xf_div = 1; }
        }
        auto gp = std::make_shared<GemmParams>();
        int r = build_gemm(pr, gp.get());
        if (r != MMD_OK) { set_err(r); return; }
        const int bn = pr.bn;
        const double tokens = static_cast<double>(g.tokens());
        const double k_alg = static_cast<double>(pc->k_total - pc->identity_c);
        const double flops = 2.0 * tokens * k_alg * pc->n;
        const double bytes = 2.0 * (tokens * static_cast<double>(ctot) + tokens * pc->n * (out_f32 ? 2.0 : 1.0) + k_alg * pc->n);
        push([gp, bn](cudaStream_t st) { return launch_gemm(*gp, bn, st); }, tag, flops, bytes);
    }
    static ConvGeom geom2(long long tokens) { ConvGeom g; g.rank = 2; g.dims[0] = tokens; geom_fill_box(g); return g; }
    ConvGeom geom_spatial(const VT& v) const { ConvGeom g; g.rank = 4; g.dims[0] = v.W; g.dims[1] = v.H; g.dims[2] = static_cast<long long>(B) * F(); geom_fill_box(g); return g; }
    ConvGeom geom_temporal(const VT& v) const { ConvGeom g; g.rank = 4; g.dims[0] = static_cast<long long>(v.H) * v.W; g.dims[1] = F(); g.dims[2] = B; geom_fill_box(g); return g; }
    ConvGeom geom_audio(const AT& a) const { ConvGeom g; g.rank = 3; g.dims[0] = a.L; g.dims[1] = B; geom_fill_box(g); return g; }

    // GroupNorm over `ns` domains of `rows` rows on the concat of (x1,c1),(x2,c2)
    // `st` (optional): statistics already accumulated by the producer of x1 (single-source only);
    // per_frame selects one slot per domain instead of the nsub slots of a sample.
    // xf (optional, inference plans): the consumer is a pointwise GEMM of geometry xf_kind that can apply the norm on
    // its A operand; then nothing is written here (only the statistics, if the producer did not fuse them) and the
    // raw input is returned for the GEMM to read.
    act_t* emit_gn(const act_t* x1, int c1, const act_t* x2, int c2, int ns, int rows, GnP gn, const float* film,
                   int ns_per_batch, int silu, const Stat* st = nullptr, bool per_frame = false, XfSpec* xf = nullptr,
                   int xf_kind = 0, int drop_modality = -1, const Stat* st2 = nullptr) {
        const int C = c1 + c2;
        // training plans: the out_layers' Dropout rides on this kernel (mask from a counter-based generator, so the
        // backward regenerates it); inference plans never drop
        const int drop_site = (train && drop_modality >= 0) ? drop_site_top++ : -1;
        const DropState* drop_dev = (drop_site >= 0 && plan) ? plan->drop_dev : nullptr;
        if (drop_site >= 0 && emitting())
            plan->drop_sites.push_back(Plan::DropSite{static_cast<uint32_t>(drop_site), static_cast<long long>(ns) * rows, C, drop_modality});
        const bool fused = fuse_stats && st && st->slots && !x2;
        const bool defer = xf && xf_kind != 0 && !train && !x2 && C % 64 == 0 && C <= GEMM_XF_MAXC &&
                           (xf_kind == 3 ? (ns == B && rows >= 128) : (rows == 64 || rows % 128 == 0));
        if (defer) {
            double* sums = fused ? st->slots : static_cast<double*>(stats.take(sizeof(double) * 64 * ns));
            xf->on = true; xf->kind = xf_kind; xf->sums = sums; xf->gn_g = gn.g; xf->gn_b = gn.b; xf->film = film;
            xf->ns_per_batch = ns_per_batch; xf->silu = silu; xf->ns = ns; xf->rows = rows;
            xf->nsub = fused ? (per_frame ? 1 : st->nsub) : 1;
            xf->stat_rows = fused ? (per_frame ? st->rows / st->nsub : st->rows) : rows;
            if (!fused && emitting() && !bad()) {
                GnSrc s{x1, c1, c1, nullptr, 0, 0};
                push([=](cudaStream_t stx) -> int { return launch_gn_stats(s, ns, rows, sums, stx, false); },
                     "group_norm", 0.0, 2.0 * ns * static_cast<double>(rows) * C, 1);
            }
            return const_cast<act_t*>(x1);
        }
        act_t* y = alloc_s(static_cast<size_t>(ns) * rows * C);
        double* sums = fused ? st->slots : static_cast<double*>(stats.take(sizeof(double) * 64 * ns));
        if (train) {
            TapeOp op;
            op.kind = TapeOp::GN;
            op.x1 = x1; op.c1 = c1; op.x2 = x2; op.c2 = c2; op.ns = ns; op.rows = rows; op.gn_g = gn.g; op.gn_b = gn.b;
            op.emb_row0 = film ? static_cast<int>(film - emb_all) : -1;
            op.silu = silu; op.ns_per_batch = ns_per_batch; op.sums = sums; op.y = y;
            op.nsub = fused ? (per_frame ? 1 : st->nsub) : 1;
            op.stat_rows = fused ? (per_frame ? st->rows / st->nsub : st->rows) : rows;
            op.drop_site = drop_site;
            tape.push_back(op);
        }
        if (!emitting() || bad()) return y;
        if (fused) {
            GnSrc s{x1, c1, c1, nullptr, 0, 0};
            const float* gamma = pf(gn.g);
            const float* beta = pf(gn.b);
            const int film_ld = m.emb_rows;
            const int nsub = per_frame ? 1 : st->nsub;
            const long long srows = per_frame ? st->rows / st->nsub : st->rows;
            const uint32_t dsite = drop_site >= 0 ? static_cast<uint32_t>(drop_site) : 0u;
            push([=](cudaStream_t stx) -> int {
                return launch_gn_apply(s, ns, rows, sums, gamma, beta, film, film_ld, ns_per_batch, silu, y, stx, nsub, srows,
                                       drop_dev, dsite);
            }, "group_norm", 0.0, 2.0 * 2.0 * ns * static_cast<double>(rows) * C, 1);
            return y;
        }
        GnSrc s{x1, c1, c1, x2, c2, c2};
        const float* gamma = pf(gn.g);
        const float* beta = pf(gn.b);
        const int film_ld = m.emb_rows;
        const uint32_t dsite = drop_site >= 0 ? static_cast<uint32_t>(drop_site) : 0u;
        // channel concat whose two producers left per-4-channel partial sums: the group statistics are a fold of those
        // (a few KB) instead of a pass over both tensors
        if (x2 && st && st2 && st->qslots && st2->qslots && st->nsub == st2->nsub && st->rows == rows && st2->rows == rows &&
            c1 % 4 == 0 && c2 % 4 == 0 && (c1 + c2) % 128 == 0) {
            const double* q1 = st->qslots;
            const double* q2 = st2->qslots;
            const int nsub = st->nsub;
            push([=](cudaStream_t stx) -> int {
                MMD_CUDA_OK(launch_kernel(gn_fold_quads_kernel, static_cast<unsigned>(ns), 64, 0, stx, q1, c1, q2, c2, nsub, sums));
                return launch_gn_apply(s, ns, rows, sums, gamma, beta, film, film_ld, ns_per_batch, silu, y, stx, 1, 0, drop_dev, dsite);
            }, "group_norm", 0.0, 2.0 * ns * static_cast<double>(rows) * C + 2.0 * ns * static_cast<double>(rows) * C, 2);
            return y;
        }
        push([=](cudaStream_t st) -> int {
            MMD_TRY(launch_gn_stats(s, ns, rows, sums, st, false));
            return launch_gn_apply(s, ns, rows, sums, gamma, beta, film, film_ld, ns_per_batch, silu, y, st, 1, 0, drop_dev, dsite);
        }, "group_norm", 0.0, 2.0 * 2.0 * ns * static_cast<double>(rows) * C, 2);
        return y;
    }

    void emit_attn(const act_t* q, int q_ld, int q_col0, long long q_rows, const act_t* k, int k_ld, int k_col0,
                   long long k_rows, const act_t* v, int v_col0, act_t* out, int out_ld, int heads, int d, int n_blocks,
                   int q_blk, int k_blk, int win, const int* shift_dev) {
        AttnProblem pr{q, q_ld, q_col0, q_rows, k, k_ld, k_col0, k_rows, v, k_ld, v_col0, out, out_ld,
                       B, heads, d, n_blocks, q_blk, k_blk, win, shift_dev};
        float* lse = nullptr;
        if (train) {
            lse = static_cast<float*>(persist.take(sizeof(float) * heads * (q_rows + 128)));
            TapeOp op;
            op.kind = TapeOp::ATTN;
            op.ap = pr; op.lse = lse;
            tape.push_back(op);
        }
        if (!emitting() || bad()) return;
        auto ap = std::make_shared<AttnParams>();
        int r = build_attn(pr, ap.get());
        if (r != MMD_OK) { set_err(r); return; }
        ap->lse = lse;
        ap->lse_ld = q_rows;
        const double fl = 4.0 * static_cast<double>(q_rows) * (static_cast<double>(win) * k_blk) * d * heads;
        const double by = 2.0 * (2.0 * q_rows + 2.0 * k_rows) * d * heads;
        push([ap, d](cudaStream_t st) { return launch_attn(*ap, d, st); }, win == 1 && shift_dev == nullptr && q_blk == k_blk ? "self_attention" : "cross_attention", fl, by);
    }

    // ---------------- network pieces
    // SingleModalAtten over sequences: kind 0 spatial (per frame), 1 temporal (per pixel), 2 audio
    act_t* self_attention(const std::string& p, const act_t* x, int C, int kind, const VT* vt, const AT* at,
                          const Stat* in_st, Stat* out_st) {
        GnP gn = reg_gn(p + ".norm", C);
        ConvP qkv = reg_conv(p + ".qkv", 3 * C, C, {1});
        ConvP proj = reg_conv(p + ".proj_out", C, C, {1});
        const int heads = cfg().num_heads;
        const int d = C / heads;
        if (create) {
            if (C % heads != 0) set_err(fail(MMD_EINVAL, "channels %d not divisible by num_heads %d", C, heads));
            if (kind != 1 && d != 64 && d != 96 && d != 128)
                set_err(fail(MMD_EINVAL, "self-attention head dim %d unsupported by the tcgen05 kernel (64/96/128)", d));
        }
        const PackedConv* pq = pack(p + ".qkv", 3 * C, {{qkv.w, C, 1}}, 0, {qkv.b});
        const PackedConv* pp = pack(p + ".proj+res", C, {{proj.w, C, 1}}, C, {proj.b});
        if (create) return nullptr;
        const size_t tokens = vt ? vtok(*vt) : atok(*at);
        act_t* out = alloc_p(tokens * C);
        cur = (kind == 2) ? 1 : 0;
        const size_t mark = S().top;
        act_t* xn;
        XfSpec xf;
        XfSpec* xfp = (xf_mask & 2) ? &xf : nullptr;
        if (kind == 0) {
            xn = emit_gn(x, C, nullptr, 0, B * F(), vt->H * vt->W, gn, nullptr, 1, 0, in_st, true, xfp, 1);
        } else if (kind == 2) {
            xn = emit_gn(x, C, nullptr, 0, B, at->L, gn, nullptr, 1, 0, in_st, false, xfp, 3);
        } else {
            xn = alloc_s(tokens * C);
            if (train) {
                TapeOp op;
                op.kind = TapeOp::GN_TEMPORAL;
                op.in = x; op.y = xn; op.gn_g = gn.g; op.gn_b = gn.b; op.B = B; op.F = F(); op.P = vt->H * vt->W; op.C = C;
                tape.push_back(op);
            }
            if (emitting()) {
                const float* gamma = pf(gn.g);
                const float* beta = pf(gn.b);
                const int Bc = B, Fc = F(), P = vt->H * vt->W;
                push([=](cudaStream_t st) { return launch_gn_temporal(x, xn, gamma, beta, Bc, Fc, P, C, st); },
                     "group_norm", 0.0, 2.0 * 2.0 * static_cast<double>(tokens) * C);
            }
        }
        act_t* qkvb = alloc_s(tokens * 3 * C);
        AT aview{nullptr, C, at ? at->L : 0};
        const ConvGeom gtok = at ? geom_audio(aview) : geom2(static_cast<long long>(tokens));
        emit_gemm("conv1x1_qkv", gtok, {{xn, C}}, {{0, 0, 0}}, pq, qkvb, nullptr, nullptr, 0, nullptr, 0, 0, 0, &xf);
        act_t* o = alloc_s(tokens * C);
        if (kind == 0) {
            const int hw = vt->H * vt->W;
            emit_attn(qkvb, 3 * C, 0, tokens, qkvb, 3 * C, C, tokens, qkvb, 2 * C, o, C, heads, d, F(), hw, hw, 1, nullptr);
        } else if (kind == 2) {
            emit_attn(qkvb, 3 * C, 0, tokens, qkvb, 3 * C, C, tokens, qkvb, 2 * C, o, C, heads, d, 1, at->L, at->L, 1, nullptr);
        } else {
            if (train) {
                TapeOp op;
                op.kind = TapeOp::TATTN;
                op.in = qkvb; op.y = o; op.B = B; op.F = F(); op.P = vt->H * vt->W; op.C = C; op.heads = heads;
                tape.push_back(op);
            }
            const int Bc = B, Fc = F(), P = vt->H * vt->W;
            if (emitting())
            push([=](cudaStream_t st) { return launch_temporal_attn(qkvb, o, Bc, Fc, P, C, heads, st); },
                 "temporal_attention", 4.0 * static_cast<double>(tokens) * Fc * C, 2.0 * 4.0 * static_cast<double>(tokens) * C);
        }
        // statistics of the block output for its consumer's GroupNorm (the spatial output only feeds the
        // per-pixel temporal norm, which computes its own)
        double* slots = nullptr;
        double* qsl = nullptr;
        int skind = 0, shw = 0;
        if (out_st) *out_st = Stat{};
        if (kind == 1 && can_fuse_video(vt->H * vt->W, C)) {
            slots = stat_slots_video(); skind = 1; shw = vt->H * vt->W;
            qsl = quad_slots_video(C);
            if (out_st) *out_st = Stat{slots, F(), static_cast<long long>(F()) * shw, qsl};
        } else if (kind == 2 && can_fuse_audio(at->L, C)) {
            slots = stat_slots_audio(); skind = 3;
            qsl = quad_slots_audio(C);
            if (out_st) *out_st = Stat{slots, 1, at->L, qsl};
        }
        emit_gemm("conv1x1_proj", gtok, {{o, C}, {x, C}}, {{0, 0, 0}}, pp, out, nullptr, nullptr, 0, slots, skind, shw, 0, nullptr, qsl);
        release(S(), mark);
        cur = 0;
        return out;
    }

    void res_block(const std::string& p, VT& v, const act_t* v2, int vc2, AT& a, const act_t* a2, int ac2, int cout,
                   int dilation, bool up, bool down, bool vattn, bool aattn, const Stat* v2st = nullptr, const Stat* a2st = nullptr) {
        const int cin = v.C + vc2;
        const int E = cfg().model_channels;
        // registration order = module definition order in ResBlock.__init__ (multimodal_unet.py:338-419)
        GnP vin_gn = reg_gn(p + ".video_in_layers.0", cin);
        ConvP vsp = reg_conv(p + ".video_in_layers.2.video_conv_spatial", cout, cin, {3, 3});
        ConvP vtp = reg_conv(p + ".video_in_layers.2.video_conv_temporal", cout, cout, {3});
        GnP ain_gn = reg_gn(p + ".audio_in_layers.0", cin);
        ConvP aconv = reg_conv(p + ".audio_in_layers.2.audio_conv", cout, cin, {3});
        int emb_w = reg(p + ".emb_layers.1.weight", {2 * cout, E});
        int emb_b = reg(p + ".emb_layers.1.bias", {2 * cout});
        GnP vout_gn = reg_gn(p + ".video_out_layers.0", cout);
        ConvP vout = reg_conv(p + ".video_out_layers.3.video_conv", cout, cout, {1, 1, 1});
        GnP aout_gn = reg_gn(p + ".audio_out_layers.0", cout);
        ConvP aout = reg_conv(p + ".audio_out_layers.3.audio_conv", cout, cout, {1});
        ConvP vskip{-1, -1}, askip{-1, -1};
        if (cin != cout) {
            vskip = reg_conv(p + ".video_skip_connection.video_conv", cout, cin, {1, 1, 1});
            askip = reg_conv(p + ".audio_skip_connection.audio_conv", cout, cin, {1});
        }
        // stacked emb_layers rows
        const int emb_row0 = emb_row_top;
        emb_row_top += 2 * cout;
        if (train) emb_blocks.push_back(EmbBlk{emb_w, emb_b, emb_row0, 2 * cout});
        if (create) {
            MmdModel* mp = &m;
            const int rows = 2 * cout;
            m.pack_ops.push_back([mp, emb_w, emb_b, emb_row0, rows, E](cudaStream_t st) -> int {
                MMD_CUDA_OK(cudaMemcpyAsync(mp->bpk + mp->emb_w_off + static_cast<size_t>(emb_row0) * E,
                                            mp->w32 + mp->params[emb_w].offset, sizeof(float) * rows * E,
                                            cudaMemcpyDeviceToDevice, st));
                MMD_CUDA_OK(cudaMemcpyAsync(mp->bpk + mp->emb_b_off + emb_row0, mp->w32 + mp->params[emb_b].offset,
                                            sizeof(float) * rows, cudaMemcpyDeviceToDevice, st));
                return MMD_OK;
            });
        }
        const PackedConv* p_vsp = pack(p + ".v_spatial", cout, {{vsp.w, cin, 9}}, 0, {vsp.b});
        const PackedConv* p_vtp = pack(p + ".v_temporal", cout, {{vtp.w, cout, 3}}, 0, {vtp.b});
        const PackedConv* p_ac = pack(p + ".a_conv", cout, {{aconv.w, cin, 3}}, 0, {aconv.b});
        const PackedConv *p_vo, *p_ao;
        if (cin != cout) {
            p_vo = pack(p + ".v_out+skip", cout, {{vout.w, cout, 1}, {vskip.w, cin, 1}}, 0, {vout.b, vskip.b});
            p_ao = pack(p + ".a_out+skip", cout, {{aout.w, cout, 1}, {askip.w, cin, 1}}, 0, {aout.b, askip.b});
        } else {
            p_vo = pack(p + ".v_out+res", cout, {{vout.w, cout, 1}}, cin, {vout.b});
            p_ao = pack(p + ".a_out+res", cout, {{aout.w, cout, 1}}, cin, {aout.b});
        }

        VT vo{nullptr, cout, v.H, v.W};
        AT ao{nullptr, cout, a.L};
        if (!create) {
            if ((up || down) && (v2 || a2)) { set_err(fail(MMD_EINVAL, "internal: resampling block with concat input")); return; }
            const float* film = emb_all ? emb_all + emb_row0 : nullptr;
            const int Fr = F();
            // ---------------- video branch
            {
                if (down) { vo.H = v.H / 2; vo.W = v.W / 2; }
                if (up) { vo.H = v.H * 2; vo.W = v.W * 2; }
                vo.p = alloc_p(static_cast<size_t>(B) * Fr * vo.H * vo.W * cout);
                cur = 0;
                const size_t mark = S().top;
                const int hw = v.H * v.W;
                act_t* h0 = emit_gn(v.p, v.C, v2, vc2, B, Fr * hw, vin_gn, nullptr, 1, 1, &v.st, false, nullptr, 0, -1, v2st);
                act_t* u = alloc_s(vtok(v) * cout);
                VT vin{h0, cin, v.H, v.W};
                std::vector<std::array<int, 3>> taps9;
                for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) taps9.push_back({kx - 1, ky - 1, 0});
                emit_gemm("conv3x3_spatial", geom_spatial(vin), {{h0, cin}}, taps9, p_vsp, u);
                act_t* h1 = alloc_s(vtok(v) * cout);
                // GN2 statistics from the temporal conv's epilogue (nearest upsampling keeps them; pooling does not)
                Stat h1st;
                if (!down && can_fuse_video(hw, cout)) h1st = Stat{stat_slots_video(), Fr, static_cast<long long>(Fr) * hw};
                emit_gemm("conv_temporal", geom_temporal(vin), {{u, cout}}, {{0, -1, 0}, {0, 0, 0}, {0, 1, 0}}, p_vtp, h1,
                          nullptr, nullptr, 0, h1st.slots, 2, 0);
                const act_t* xs = v.p;   // skip-path input (single source when resampling)
                if (up || down) {
                    act_t* h1r = alloc_s(static_cast<size_t>(B) * Fr * vo.H * vo.W * cout);
                    act_t* xr = alloc_s(static_cast<size_t>(B) * Fr * vo.H * vo.W * v.C);
                    if (train) {
                        TapeOp op;
                        op.kind = TapeOp::RESAMPLE;
                        op.mode = down ? 0 : 2; op.n_ = B * Fr; op.h_ = v.H; op.w_ = v.W;
                        op.in = h1; op.y = h1r; op.C = cout;
                        tape.push_back(op);
                        op.in = v.p; op.y = xr; op.C = v.C;
                        tape.push_back(op);
                    }
                    if (emitting()) {
                        const int mode = down ? 0 : 2;
                        const int n = B * Fr, H = v.H, W = v.W, c_h = cout, c_x = v.C;
                        const act_t* xin = v.p;
                        const double elems = static_cast<double>(n) * H * W * (c_h + c_x);
                        push([=](cudaStream_t st) -> int {
                            MMD_TRY(launch_resample(h1, h1r, mode, n, H, W, c_h, st));
                            return launch_resample(xin, xr, mode, n, H, W, c_x, st);
                        }, "resample", 0.0, 2.0 * elems * (down ? 1.25 : 5.0), 2);
                    }
                    h1 = h1r;
                    xs = xr;
                }
                const int hwo = vo.H * vo.W;
                XfSpec xfv;
                act_t* h2 = emit_gn(h1, cout, nullptr, 0, B, Fr * hwo, vout_gn, film, 1, 1, &h1st, false,
                                    (xf_mask & 1) ? &xfv : nullptr, 1, 0);
                std::vector<std::pair<const act_t*, int>> srcs = {{h2, cout}, {xs, v.C}};
                if (v2) srcs.push_back({v2, vc2});
                if (can_fuse_video(hwo, cout)) vo.st = Stat{stat_slots_video(), Fr, static_cast<long long>(Fr) * hwo, quad_slots_video(cout)};
                emit_gemm("conv1x1_out", geom2(static_cast<long long>(B) * Fr * hwo), srcs, {{0, 0, 0}}, p_vo, vo.p,
                          nullptr, nullptr, 0, vo.st.slots, 1, hwo, 0, &xfv, vo.st.qslots);
                release(S(), mark);
            }
            // ---------------- audio branch
            {
                if (down) ao.L = a.L / 4;
                if (up) ao.L = a.L * 4;
                ao.p = alloc_p(static_cast<size_t>(B) * ao.L * cout);
                cur = 1;
                const size_t mark = S().top;
                act_t* h0 = emit_gn(a.p, a.C, a2, ac2, B, a.L, ain_gn, nullptr, 1, 1, &a.st, false, nullptr, 0, -1, a2st);
                act_t* h1 = alloc_s(atok(a) * cout);
                AT ain{h0, cin, a.L};
                Stat h1st;
                if (!down && can_fuse_audio(a.L, cout)) h1st = Stat{stat_slots_audio(), 1, a.L};
                emit_gemm("conv_audio_k3", geom_audio(ain), {{h0, cin}}, {{-dilation, 0, 0}, {0, 0, 0}, {dilation, 0, 0}}, p_ac, h1,
                          nullptr, nullptr, 0, h1st.slots, 3, 0);
                const act_t* xs = a.p;
                if (up || down) {
                    act_t* h1r = alloc_s(static_cast<size_t>(B) * ao.L * cout);
                    act_t* xr = alloc_s(static_cast<size_t>(B) * ao.L * a.C);
                    if (train) {
                        TapeOp op;
                        op.kind = TapeOp::RESAMPLE;
                        op.mode = down ? 1 : 3; op.n_ = B; op.h_ = a.L; op.w_ = 1;
                        op.in = h1; op.y = h1r; op.C = cout;
                        tape.push_back(op);
                        op.in = a.p; op.y = xr; op.C = a.C;
                        tape.push_back(op);
                    }
                    if (emitting()) {
                        const int mode = down ? 1 : 3;
                        const int n = B, L = a.L, c_h = cout, c_x = a.C;
                        const act_t* xin = a.p;
                        const double elems = static_cast<double>(n) * L * (c_h + c_x);
                        push([=](cudaStream_t st) -> int {
                            MMD_TRY(launch_resample(h1, h1r, mode, n, L, 1, c_h, st));
                            return launch_resample(xin, xr, mode, n, L, 1, c_x, st);
                        }, "resample", 0.0, 2.0 * elems * (down ? 1.25 : 5.0), 2);
                    }
                    h1 = h1r;
                    xs = xr;
                }
                XfSpec xfa;
                act_t* h2 = emit_gn(h1, cout, nullptr, 0, B, ao.L, aout_gn, film, 1, 1, &h1st, false,
                                    (xf_mask & 1) ? &xfa : nullptr, 3, 1);
                std::vector<std::pair<const act_t*, int>> srcs = {{h2, cout}, {xs, a.C}};
                if (a2) srcs.push_back({a2, ac2});
                if (can_fuse_audio(ao.L, cout)) ao.st = Stat{stat_slots_audio(), 1, ao.L, quad_slots_audio(cout)};
                emit_gemm("conv1x1_out", geom_audio(ao), srcs, {{0, 0, 0}}, p_ao, ao.p, nullptr, nullptr, 0, ao.st.slots, 3, 0,
                          0, &xfa, ao.st.qslots);
                release(S(), mark);
                cur = 0;
            }
        }
        // ---------------- in-block self attention (multimodal_unet.py:485-493)
        if (vattn) {
            Stat tst;
            act_t* s = self_attention(p + ".spatial_attention_block", vo.p, cout, 0, &vo, nullptr, &vo.st, nullptr);
            act_t* t = self_attention(p + ".temporal_attention_block", s, cout, 1, &vo, nullptr, nullptr, &tst);
            vo.p = t;
            vo.st = tst;
        }
        if (aattn) {
            Stat ast;
            act_t* s = self_attention(p + ".audio_attention_block", ao.p, cout, 2, nullptr, &ao, &ao.st, &ast);
            ao.p = s;
            ao.st = ast;
        }
        v = vo;
        a = ao;
    }

    void cross_block(const std::string& p, VT& v, AT& a, int window, bool shift) {
        const int C = v.C;
        GnP vn = reg_gn(p + ".v_norm", C);
        GnP an = reg_gn(p + ".a_norm", C);
        ConvP vq = reg_conv(p + ".v_qkv", 3 * C, C, {1});
        ConvP aq = reg_conv(p + ".a_qkv", 3 * C, C, {1});
        ConvP vp = reg_conv(p + ".video_proj_out.video_conv", C, C, {1, 1, 1});
        ConvP apj = reg_conv(p + ".audio_proj_out.audio_conv", C, C, {1});
        const int heads = cfg().num_head_channels == -1 ? cfg().num_heads : C / cfg().num_head_channels;
        const int d = C / std::max(heads, 1);
        const int slot = shift_slot++;
        if (create) {
            if (heads < 1 || C % heads != 0 || (d != 64 && d != 96 && d != 128))
                set_err(fail(MMD_EINVAL, "cross-attention head dim %d unsupported (64/96/128)", d));
            if (window < 1 || window > F()) set_err(fail(MMD_EINVAL, "cross-attention window %d vs %d frames", window, F()));
            m.shift_bounds.push_back(shift ? F() - window : -1);
        }
        const PackedConv* p_vq = pack(p + ".v_qkv", 3 * C, {{vq.w, C, 1}}, 0, {vq.b});
        const PackedConv* p_aq = pack(p + ".a_qkv", 3 * C, {{aq.w, C, 1}}, 0, {aq.b});
        const PackedConv* p_vp = pack(p + ".v_proj+res", C, {{vp.w, C, 1}}, C, {vp.b});
        const PackedConv* p_ap = pack(p + ".a_proj+res", C, {{apj.w, C, 1}}, C, {apj.b});
        if (create) return;
        const int Fr = F();
        const int hw = v.H * v.W;
        if (a.L % Fr != 0) { set_err(fail(MMD_EINVAL, "audio length %d not divisible by %d frames (unsupported remainder segment)", a.L, Fr)); return; }
        const int apf = a.L / Fr;
        const size_t vt = vtok(v), at = atok(a);
        act_t* vout = alloc_p(vt * C);
        act_t* aout = alloc_p(at * C);
        const size_t mark_v = scratch.top, mark_a = scratch_a.top;
        cur = 0;
        XfSpec xfv, xfa;
        act_t* vnrm = emit_gn(v.p, C, nullptr, 0, B, Fr * hw, vn, nullptr, 1, 0, &v.st, false, (xf_mask & 4) ? &xfv : nullptr, 1);
        act_t* vqkv = alloc_s(vt * 3 * C);
        emit_gemm("conv1x1_qkv", geom2(static_cast<long long>(vt)), {{vnrm, C}}, {{0, 0, 0}}, p_vq, vqkv, nullptr, nullptr, 0,
                  nullptr, 0, 0, 0, &xfv);
        act_t* ov = alloc_s(vt * C);
        cur = 1;
        act_t* anrm = emit_gn(a.p, C, nullptr, 0, B, a.L, an, nullptr, 1, 0, &a.st, false, (xf_mask & 4) ? &xfa : nullptr, 3);
        act_t* aqkv = alloc_s(at * 3 * C);
        emit_gemm("conv1x1_qkv", geom_audio(a), {{anrm, C}}, {{0, 0, 0}}, p_aq, aqkv, nullptr, nullptr, 0, nullptr, 0, 0, 0,
                  &xfa);
        act_t* oa = alloc_s(at * C);
        const int* sdev = (shift && plan) ? plan->shifts_dev + slot : nullptr;
        // video queries attend audio keys/values and vice versa (multimodal_unet.py:530-559): each branch needs the
        // other's qkv projection, and nobody may recycle its qkv scratch before both attention kernels are done
        sync_branches();
        cur = 0;
        emit_attn(vqkv, 3 * C, 0, vt, aqkv, 3 * C, C, at, aqkv, 2 * C, ov, C, heads, d, Fr, hw, apf, window, sdev);
        cur = 1;
        emit_attn(aqkv, 3 * C, 0, at, vqkv, 3 * C, C, vt, vqkv, 2 * C, oa, C, heads, d, Fr, apf, hw, window, sdev);
        sync_branches();
        cur = 0;
        Stat vst, ast;
        if (can_fuse_video(hw, C)) vst = Stat{stat_slots_video(), Fr, static_cast<long long>(Fr) * hw, quad_slots_video(C)};
        if (can_fuse_audio(a.L, C)) ast = Stat{stat_slots_audio(), 1, a.L, quad_slots_audio(C)};
        emit_gemm("conv1x1_proj", geom2(static_cast<long long>(vt)), {{ov, C}, {v.p, C}}, {{0, 0, 0}}, p_vp, vout,
                  nullptr, nullptr, 0, vst.slots, 1, hw, 0, nullptr, vst.qslots);
        cur = 1;
        emit_gemm("conv1x1_proj", geom_audio(a), {{oa, C}, {a.p, C}}, {{0, 0, 0}}, p_ap, aout, nullptr, nullptr, 0,
                  ast.slots, 3, 0, 0, nullptr, ast.qslots);
        cur = 0;
        release(scratch, mark_v);
        release(scratch_a, mark_a);
        v.p = vout;
        a.p = aout;
        v.st = vst;
        a.st = ast;
    }

    static bool contains(const int* arr, int n, int v) {
        for (int i = 0; i < n; ++i) if (arr[i] == v) return true;
        return false;
    }
    static int index_of(const int* arr, int n, int v) {
        for (int i = 0; i < n; ++i) if (arr[i] == v) return i;
        return -1;
    }

    void walk() {
        const MmdConfig& c = cfg();
        const int mc = c.model_channels;
        const int E = mc;
        // ---- time embedding (multimodal_unet.py:791-795, 1075)
        int te0w = reg("time_embed.0.weight", {E, mc}), te0b = reg("time_embed.0.bias", {E});
        int te2w = reg("time_embed.2.weight", {E, E}), te2b = reg("time_embed.2.bias", {E});
        float* emb = nullptr;
        te_idx[0] = te0w; te_idx[1] = te0b; te_idx[2] = te2w; te_idx[3] = te2b;
        if (!create && emitting()) {
            uint8_t* sbase = stats.base;
            const size_t sbytes = stats.cap;
            push([=](cudaStream_t st) -> int {
                MMD_CUDA_OK(cudaMemsetAsync(sbase, 0, sbytes, st));
                pdl_break(st);
                return MMD_OK;
            }, "memset", 0.0, static_cast<double>(sbytes), 0);
        }
        if (!create) {
            emb = static_cast<float*>(persist.take(sizeof(float) * B * E));
            silu_emb = static_cast<float*>(persist.take(sizeof(float) * B * E));
            emb_all = static_cast<float*>(persist.take(sizeof(float) * B * std::max(m.emb_rows, 1)));
            if (emitting()) {
                const float *w1 = pf(te0w), *b1 = pf(te0b), *w2 = pf(te2w), *b2 = pf(te2b);
                const float* tdev = plan->t_dev;
                float* se = silu_emb;
                float* ea = emb_all;
                const int Bc = B, rows = m.emb_rows;
                const float* ew = m.bpk + m.emb_w_off;
                const float* eb = m.bpk + m.emb_b_off;
                push([=](cudaStream_t st) -> int {
                    MMD_CUDA_OK(launch_kernel(time_embed_kernel, Bc, E, 2 * E * sizeof(float), st, tdev, w1, b1, w2, b2, E, emb, se));
                    MMD_CUDA_OK(launch_kernel(emb_layers_kernel, (rows + 7) / 8, 256, 0, st, se, ew, eb, Bc, E, rows, ea));
                    return MMD_OK;
                }, "time_embed", 2.0 * Bc * (2.0 * E * E + static_cast<double>(rows) * E), 4.0 * (static_cast<double>(rows) * E + 2.0 * E * E), 2);
                sync_branches();   // the statistics memset and the FiLM table precede both branches
            }
        }
        // ---- input blocks
        int ch = c.channel_mult[0] * mc;
        std::vector<int> chans = {ch};
        VT v{nullptr, ch, c.video_h, c.video_w};
        AT a{nullptr, ch, c.audio_l};
        std::vector<VT> vstack;
        std::vector<AT> astack;
        {   // InitialBlock (multimodal_unet.py:680-694)
            const std::string p = "input_blocks.0.0";
            ConvP vsp = reg_conv(p + ".video_conv.video_conv_spatial", ch, c.video_c, {3, 3});
            ConvP vtp = reg_conv(p + ".video_conv.video_conv_temporal", ch, ch, {3});
            ConvP ac = reg_conv(p + ".audio_conv.audio_conv", ch, c.audio_c, {3});
            if (create && (9 * c.video_c > 64 || 3 * c.audio_c > 64))
                set_err(fail(MMD_EINVAL, "input channels too wide for the im2col stem (video_c %d audio_c %d)", c.video_c, c.audio_c));
            const PackedConv* p_sp = pack(p + ".v_spatial", ch, {{vsp.w, c.video_c, 9}}, 0, {vsp.b}, 0, 64);
            const PackedConv* p_tp = pack(p + ".v_temporal", ch, {{vtp.w, ch, 3}}, 0, {vtp.b});
            const PackedConv* p_ac = pack(p + ".a_conv", ch, {{ac.w, c.audio_c, 3}}, 0, {ac.b}, 0, 64);
            if (!create) {
                v.p = alloc_p(vtok(v) * ch);
                a.p = alloc_p(atok(a) * ch);
                const int hw0 = v.H * v.W;
                {   // video branch
                    cur = 0;
                    const size_t mark = S().top;
                    act_t* colv = alloc_s(vtok(v) * 64);
                    act_t* u = alloc_s(vtok(v) * ch);
                    if (emitting()) {
                        const float* vin = plan->in_video;
                        const int BF = B * F(), Cv = c.video_c, H = c.video_h, W = c.video_w;
                        push([=](cudaStream_t st) -> int {
                            const long long tv = static_cast<long long>(BF) * H * W * 8;
                            MMD_CUDA_OK(launch_kernel(im2col_video_kernel, static_cast<unsigned>((tv + 255) / 256), 256, 0, st, vin, colv,
                                                      BF, Cv, H, W));
                            return MMD_OK;
                        }, "im2col", 0.0, (4.0 * Cv + 128.0) * static_cast<double>(BF) * H * W, 1);
                    }
                    emit_gemm("conv_stem", geom2(static_cast<long long>(vtok(v))), {{colv, 64}}, {{0, 0, 0}}, p_sp, u, nullptr, nullptr, 0, nullptr, 0, 0, 1);
                    if (can_fuse_video(hw0, ch)) v.st = Stat{stat_slots_video(), F(), static_cast<long long>(F()) * hw0, quad_slots_video(ch)};
                    emit_gemm("conv_temporal", geom_temporal(v), {{u, ch}}, {{0, -1, 0}, {0, 0, 0}, {0, 1, 0}}, p_tp, v.p,
                              nullptr, nullptr, 0, v.st.slots, 2, 0, 0, nullptr, v.st.qslots);
                    release(S(), mark);
                }
                {   // audio branch
                    cur = 1;
                    const size_t mark = S().top;
                    act_t* cola = alloc_s(atok(a) * 64);
                    if (emitting()) {
                        const float* ain = plan->in_audio;
                        const int Bc = B, Ca = c.audio_c, L = c.audio_l;
                        push([=](cudaStream_t st) -> int {
                            const long long ta = static_cast<long long>(Bc) * L * 8;
                            MMD_CUDA_OK(launch_kernel(im2col_audio_kernel, static_cast<unsigned>((ta + 255) / 256), 256, 0, st, ain, cola,
                                                      Bc, Ca, L));
                            return MMD_OK;
                        }, "im2col", 0.0, (4.0 * Ca + 128.0) * static_cast<double>(Bc) * L, 1);
                    }
                    if (can_fuse_audio(a.L, ch)) a.st = Stat{stat_slots_audio(), 1, a.L, quad_slots_audio(ch)};
                    emit_gemm("conv_stem", geom_audio(a), {{cola, 64}}, {{0, 0, 0}}, p_ac, a.p, nullptr, nullptr, 0, a.st.slots, 3, 0, 2,
                              nullptr, a.st.qslots);
                    release(S(), mark);
                    cur = 0;
                }
            }
            vstack.push_back(v);
            astack.push_back(a);
        }
        int ds = 1, dil = 1, idx = 1;
        for (int level = 0; level < c.n_levels; ++level) {
            const int mult = c.channel_mult[level];
            for (int r = 0; r < c.num_res_blocks; ++r) {
                const std::string p = "input_blocks." + std::to_string(idx);
                const int cout = mult * mc;
                res_block(p + ".0", v, nullptr, 0, a, nullptr, 0, cout, 1 << (dil % 10), false, false,
                          contains(c.video_attention_resolutions, c.n_video_attn, ds),
                          contains(c.audio_attention_resolutions, c.n_audio_attn, ds));
                ++dil;
                ch = cout;
                if (contains(c.cross_attention_resolutions, c.n_cross, ds)) {
                    const int wi = index_of(c.cross_attention_resolutions, c.n_cross, ds);
                    cross_block(p + ".1", v, a, c.cross_attention_windows[wi], c.cross_attention_shift != 0);
                }
                vstack.push_back(v);
                astack.push_back(a);
                chans.push_back(ch);
                ++idx;
                if (bad()) return;
            }
            if (level != c.n_levels - 1) {
                const std::string p = "input_blocks." + std::to_string(idx);
                res_block(p + ".0", v, nullptr, 0, a, nullptr, 0, ch, 1 << (dil % 10), false, true, false, false);
                ++dil;
                vstack.push_back(v);
                astack.push_back(a);
                chans.push_back(ch);
                ds *= 2;
                ++idx;
            }
        }
        // ---- middle (multimodal_unet.py:875-941)
        const int mid_dil = 1 << (dil % 10);
        const bool three_part = c.n_cross == 3 && c.cross_attention_windows[0] == 1 && c.cross_attention_windows[1] == 4 &&
                                c.cross_attention_windows[2] == 8;
        res_block("middle_blocks.0", v, nullptr, 0, a, nullptr, 0, ch, mid_dil, false, false, true, true);
        if (three_part) {
            cross_block("middle_blocks.1", v, a, c.video_f, false);
            res_block("middle_blocks.2", v, nullptr, 0, a, nullptr, 0, ch, mid_dil, false, false, true, true);
        } else {
            res_block("middle_blocks.1", v, nullptr, 0, a, nullptr, 0, ch, mid_dil, false, false, true, true);
        }
        if (bad()) return;
        // ---- output blocks (multimodal_unet.py:944-1000, 1092-1095)
        dil -= 1;
        idx = 0;
        for (int level = c.n_levels - 1; level >= 0; --level) {
            const int mult = c.channel_mult[level];
            for (int i = 0; i <= c.num_res_blocks; ++i) {
                const std::string p = "output_blocks." + std::to_string(idx);
                const int ich = chans.back();
                chans.pop_back();
                VT vs = vstack.back(); vstack.pop_back();
                AT as = astack.back(); astack.pop_back();
                if (!create && (vs.H != v.H || vs.W != v.W || as.L != a.L || vs.C != ich)) {
                    set_err(fail(MMD_EINVAL, "skip shape mismatch at %s (resblock_updown=False style configs are unsupported)", p.c_str()));
                    return;
                }
                const int cout = mc * mult;
                int sub = 0;
                res_block(p + "." + std::to_string(sub++), v, vs.p, ich, a, as.p, ich, cout, 1 << (dil % 10), false, false,
                          contains(c.video_attention_resolutions, c.n_video_attn, ds),
                          contains(c.audio_attention_resolutions, c.n_audio_attn, ds), &vs.st, &as.st);
                dil -= 1;
                ch = cout;
                if (contains(c.cross_attention_resolutions, c.n_cross, ds)) {
                    const int wi = index_of(c.cross_attention_resolutions, c.n_cross, ds);
                    cross_block(p + "." + std::to_string(sub++), v, a, c.cross_attention_windows[wi], c.cross_attention_shift != 0);
                }
                if (level > 0 && i == c.num_res_blocks) {
                    res_block(p + "." + std::to_string(sub++), v, nullptr, 0, a, nullptr, 0, ch, 1 << (dil % 10), true, false, false, false);
                    ds /= 2;
                }
                ++idx;
                if (bad()) return;
            }
        }
        // ---- heads (multimodal_unet.py:1003-1012, 1097-1098); audio_out is registered first
        const int ch0 = c.channel_mult[0] * mc;
        GnP agn = reg_gn("audio_out.0", ch);
        ConvP ahead = reg_conv("audio_out.2.audio_conv", c.audio_out_channels, ch0, {3});
        GnP vgn = reg_gn("video_out.0", ch);
        ConvP vhead = reg_conv("video_out.2.video_conv", c.video_out_channels, ch0, {3, 3, 3});
        if (create && (ch != ch0 || c.video_out_channels > 16 || c.audio_out_channels > 16))
            set_err(fail(MMD_EINVAL, "head configuration unsupported (ch %d vs %d)", ch, ch0));
        const PackedConv* p_ah = pack("audio_out.head", c.audio_out_channels, {{ahead.w, ch0, 3}}, 0, {ahead.b}, 16);
        const PackedConv* p_vh = pack("video_out.head", c.video_out_channels, {{vhead.w, ch0, 27}}, 0, {vhead.b}, 16);
        // inference plans run the video head as one pointwise GEMM over the 27 x 4 per-tap columns + a neighbour gather
        // (head_gather3d_kernel); the tap-major weight pack lives next to the implicit-GEMM pack the training plans use
        const bool head_taps_ok = c.video_out_channels <= 4 && ch0 % 64 == 0;
        const PackedConv* p_vt = head_taps_ok ? pack_head_taps("video_out.head_taps", vhead.w, c.video_out_channels, ch0, 27) : nullptr;
        if (!create) {
            const size_t mark = scratch.top, mark_a = scratch_a.top;
            const int Fr = F();
            cur = 0;
            act_t* hv = emit_gn(v.p, ch, nullptr, 0, B, Fr * v.H * v.W, vgn, nullptr, 1, 1, &v.st, false);
            ConvGeom g5; g5.rank = 5; g5.dims[0] = v.W; g5.dims[1] = v.H; g5.dims[2] = Fr; g5.dims[3] = B; geom_fill_box(g5);
            std::vector<std::array<int, 3>> taps27;
            for (int kt = 0; kt < 3; ++kt) for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) taps27.push_back({kx - 1, ky - 1, kt - 1});
            const long long Co = c.video_out_channels, HW = static_cast<long long>(v.H) * v.W;
            const long long os_v[4] = {1, v.W, Co * HW, Fr * Co * HW};
            if (!train && p_vt && head_taps_on()) {
                const size_t tokens = vtok(v);
                act_t* ytap = alloc_s(tokens * 128);
                ConvGeom g2; g2.rank = 2; g2.dims[0] = static_cast<long long>(tokens); geom_fill_box(g2);
                emit_gemm("conv_head", g2, {{hv, ch}}, {{0, 0, 0}}, p_vt, ytap);
                if (emitting() && !bad()) {
                    const float* hb = pf(vhead.b);
                    float* ov = plan->out_video;
                    const int Bc = B, Hc = v.H, Wc = v.W, Coc = c.video_out_channels;
                    push([=](cudaStream_t st) -> int {
                        const unsigned blocks = static_cast<unsigned>((static_cast<size_t>(Bc) * Fr * Hc * Wc + 255) / 256);
                        MMD_CUDA_OK(launch_kernel(head_gather3d_kernel, blocks, 256, 0, st, static_cast<const act_t*>(ytap), hb, ov, Bc, Fr, Hc, Wc, Coc));
                        return MMD_OK;
                    }, "conv_head", 0.0, 2.0 * tokens * 128 + 4.0 * tokens * Coc, 1);
                }
            } else {
                emit_gemm("conv_head", g5, {{hv, ch}}, taps27, p_vh, nullptr, plan ? plan->out_video : nullptr, os_v, HW);
            }
            cur = 1;
            act_t* ha = emit_gn(a.p, ch, nullptr, 0, B, a.L, agn, nullptr, 1, 1, &a.st, false);
            const long long Ca = c.audio_out_channels;
            const long long os_a[4] = {1, Ca * a.L, 0, 0};
            emit_gemm("conv_head", geom_audio(a), {{ha, ch}}, {{-1, 0, 0}, {0, 0, 0}, {1, 0, 0}}, p_ah, nullptr, plan ? plan->out_audio : nullptr, os_a, a.L);
            cur = 0;
            sync_branches();   // join: the caller's stream continues after both heads
            release(scratch, mark);
            release(scratch_a, mark_a);
        }
        if (create) m.emb_rows = emb_row_top;
    }

    // =====================================================================================================
    // Backward launch list from the tape (training plans).  Runs in the dry pass too (it sizes the gradient,
    // temporary and transposed-pack arenas); closures are only built when emitting.
    // =====================================================================================================
    float* g32_of(int param) const { return plan->g32 + m.params[param].offset; }
    act_t* grad_buf(const act_t* x, size_t elems) {
        auto it = grad_map.find(x);
        if (it != grad_map.end()) return it->second.first;
        act_t* gbuf = static_cast<act_t*>(grads.take(elems * sizeof(act_t)));
        grad_map[x] = {gbuf, false};
        return gbuf;
    }
    bool grad_written(const act_t* x) const {
        auto it = grad_map.find(x);
        return it != grad_map.end() && it->second.second;
    }
    void mark_written(const act_t* x) { grad_map[x].second = true; }
    void bpush(LaunchFn fn, const std::string& kind) {
        if (!emitting()) return;
        plan->bwd_steps.push_back(std::move(fn));
        plan->bwd_kind.push_back(kind);
    }
    // gradient of an op output; every forward tensor has a consumer, so it must have been produced already
    act_t* need_grad(const act_t* y, const char* what) {
        if (!grad_written(y)) { set_err(fail(MMD_ESTATE, "internal: backward of %s has no incoming gradient", what)); return nullptr; }
        return grad_map[y].first;
    }

    void bwd_gemm(const TapeOp& op) {
        const PackedConv* pc = op.pc;
        const long long tokens = op.g.tokens();
        const int n = pc->n;
        const float* gs = plan ? plan->gscale : nullptr;
        if (op.tag == "conv_head") {   // narrow fp32 heads: im2col of the API-layout gradient, then the tensor-core GEMM / wgrad kernels
            const bool video = op.g.rank == 5;
            const act_t* x = op.srcs[0].first;
            const int C = op.srcs[0].second;
            act_t* dx = grad_buf(x, static_cast<size_t>(tokens) * C);
            if (grad_written(x)) { set_err(fail(MMD_ESTATE, "internal: head input gradient already written")); return; }
            mark_written(x);
            const int terms = static_cast<int>(op.taps.size()) * n;
            if (terms > 128 || C % 64 != 0) { set_err(fail(MMD_EINVAL, "head backward: %d taps x %d outputs over %d channels unsupported", static_cast<int>(op.taps.size()), n, C)); return; }
            const int ldG = head_ld(terms);
            const size_t hmark = bscratch.top;
            act_t* G = static_cast<act_t*>(bscratch.take(sizeof(act_t) * static_cast<size_t>(tokens) * ldG));
            float* dwpk = static_cast<float*>(bscratch.take(sizeof(float) * static_cast<size_t>(ldG) * C));
            act_t* wt = static_cast<act_t*>(tpack.take(sizeof(act_t) * static_cast<size_t>(C) * ldG));
            release_b(hmark);
            if (!emitting()) return;
            HeadGeom hg{};
            hg.ncoord = op.g.rank - 1;
            for (int i = 0; i < 4; ++i) { hg.dims[i] = static_cast<int>(op.g.dims[i]); hg.ostride[i] = op.ostride[i]; }
            hg.ostride_c = op.ostride_c;
            hg.n_out = n;
            hg.n_taps = static_cast<int>(op.taps.size());
            for (size_t t = 0; t < op.taps.size(); ++t) for (int j = 0; j < 3; ++j) hg.tap[t][j] = op.taps[t][j];
            hg.C = C;
            const float* dout = video ? plan->d_out_video : plan->d_out_audio;
            const float* w = pf(pc->segs[0].w);
            float* dw = g32_of(pc->segs[0].w);
            float* db = g32_of(pc->biases[0]);
            // tensor-core path shared with mmd_op_head_bwd: G = shifted, scaled fp16 copy of dout; dX = G Wt, dW = G^T A
            auto hp = std::make_shared<HeadBwdPlan>();
            int r = build_head_bwd(hg, x, G, wt, plan->zero_bias, dx, dwpk, hp.get());
            if (r != MMD_OK) { set_err(r); return; }
            const int n_out = n, T = hg.n_taps;
            plan->tpack_ops.push_back([=](cudaStream_t st) { return launch_pack_head_t(w, wt, n_out, C, T, ldG, st); });
            bpush([=](cudaStream_t st) { return run_head_bwd(*hp, dout, dw, db, gs, st); }, "head_bwd");
            return;
        }
        act_t* dy = need_grad(op.out, op.tag.c_str());
        if (!dy) return;
        const size_t bmark = bscratch.top;
        // ---- sources: conv segments first, identity (residual) columns after them
        long long conv_cols = 0;
        for (auto& sg : pc->segs) conv_cols += sg.ci;
        if (op.stem) conv_cols = op.srcs[0].second;
        struct SrcInfo { const act_t* p; int c; long long off; bool conv; };
        std::vector<SrcInfo> si;
        long long a = 0;
        for (auto& sc : op.srcs) {
            const bool conv = a < conv_cols;
            if (conv && a + sc.second > conv_cols) { set_err(fail(MMD_ESTATE, "internal: source straddles conv/identity columns")); return; }
            si.push_back(SrcInfo{sc.first, sc.second, conv ? a : a - conv_cols, conv});
            a += sc.second;
        }
        // ---- bias gradients: column sums of dY
        if (pc->biases.size() > 2) { set_err(fail(MMD_ESTATE, "internal: more than two biases in one GEMM")); return; }
        // (reduced inside the wgrad kernel by an extra ones-operand MMA on the dY tiles it already holds)
        // ---- weight gradients: one tcgen05 wgrad over the conv sources, unpacked per parameter
        {
            WgradProblem wp;
            wp.g = op.g;
            for (auto& x : si) if (x.conv) { wp.src[wp.n_src] = x.p; wp.src_c[wp.n_src] = x.c; ++wp.n_src; }
            wp.n_taps = static_cast<int>(op.taps.size());
            for (size_t t = 0; t < op.taps.size(); ++t) for (int j = 0; j < 3; ++j) wp.taps[t][j] = op.taps[t][j];
            wp.dy = dy;
            wp.n = n;
            wp.ld = conv_cols * wp.n_taps;
            const size_t dw_floats = static_cast<size_t>(n) * wp.ld;
            const size_t dw_bytes = sizeof(float) * (dw_floats + n);
            float* dwpk = static_cast<float*>(bscratch.take(dw_bytes));
            wp.dw = dwpk;
            float* dbpk = dwpk + dw_floats;
            if (!pc->biases.empty()) wp.db = dbpk;
            if (emitting()) {
                std::vector<float*> bdst;
                for (int bi : pc->biases) bdst.push_back(g32_of(bi));
                auto wpar = std::make_shared<WgradParams>();
                int items = 0;
                int r = build_wgrad(wp, wpar.get(), &items);
                if (r != MMD_OK) { set_err(r); return; }
                std::vector<Seg> segs = pc->segs;
                std::vector<float*> gdst;
                for (auto& sg : segs) gdst.push_back(g32_of(sg.w));
                const long long ld = wp.ld;
                bpush([=](cudaStream_t st) -> int {
                    MMD_CUDA_OK(cudaMemsetAsync(dwpk, 0, dw_bytes, st));
                    MMD_TRY(launch_wgrad(*wpar, items, st));
                    long long col = 0;
                    for (size_t i = 0; i < segs.size(); ++i) {
                        MMD_TRY(launch_unpack_wgrad(dwpk, gdst[i], n, segs[i].ci, segs[i].T, ld, col, 1.0f, st, gs));
                        col += static_cast<long long>(segs[i].ci) * segs[i].T;
                    }
                    for (float* bd : bdst) axpy_f32_kernel<<<(n + 255) / 256, 256, 0, st>>>(dbpk, bd, n, gs);
                    MMD_CUDA_OK(cudaGetLastError());
                    return MMD_OK;
                }, "wgrad:" + op.tag);
            }
        }
        // ---- data gradients
        if (op.stem) {
            // input gradient (gradient-guided sampling): dcol = dY * Wpk, then the im2col adjoint into the fp32 API layout
            const int Ci = (op.stem == 1) ? cfg().video_c : cfg().audio_c;
            const int T = pc->segs[0].T;
            act_t* wt = static_cast<act_t*>(tpack.take(sizeof(act_t) * 64 * static_cast<size_t>(n)));
            act_t* dcol = static_cast<act_t*>(bscratch.take(sizeof(act_t) * static_cast<size_t>(tokens) * 64));
            if (emitting()) {
                const float* w32p = pf(pc->segs[0].w);
                const int Co = n;
                plan->tpack_ops.push_back([=](cudaStream_t st) -> int {
                    MMD_CUDA_OK(cudaMemsetAsync(wt, 0, sizeof(act_t) * 64 * static_cast<size_t>(Co), st));
                    const long long total = static_cast<long long>(Co) * Ci * T;
                    pack_stem_t_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(w32p, wt, Co, Ci, T, Co);
                    MMD_CUDA_OK(cudaGetLastError());
                    return MMD_OK;
                });
                GemmProblem pr;
                pr.g = op.g;
                pr.n_src = 1; pr.src[0] = dy; pr.src_c[0] = n;
                pr.n_taps = 1;
                pr.w = wt; pr.bias = plan->zero_bias; pr.n = 64; pr.bn = 64; pr.out = dcol;
                auto gp = std::make_shared<GemmParams>();
                int r = build_gemm(pr, gp.get());
                if (r != MMD_OK) { set_err(r); return; }
                const int Bc = B, Fr = F(), H = cfg().video_h, W = cfg().video_w, L = cfg().audio_l;
                float* dxin = (op.stem == 1) ? plan->d_in_video : plan->d_in_audio;
                const int stem = op.stem;
                bpush([=](cudaStream_t st) -> int {
                    MMD_TRY(launch_gemm(*gp, 64, st));
                    pdl_break(st);
                    if (stem == 1) {
                        const long long total = static_cast<long long>(Bc) * Fr * Ci * H * W;
                        col2im_video_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(dcol, dxin, Bc * Fr, Ci, H, W, gs);
                    } else {
                        const long long total = static_cast<long long>(Bc) * Ci * L;
                        col2im_audio_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(dcol, dxin, Bc, Ci, L, gs);
                    }
                    MMD_CUDA_OK(cudaGetLastError());
                    return MMD_OK;
                }, "dgrad:stem");
            }
            release_b(bmark);
            return;
        }
        long long seg_start = 0;
        for (auto& x : si) {
            act_t* gx = grad_buf(x.p, static_cast<size_t>(tokens) * x.c);
            const bool acc = grad_written(x.p);
            mark_written(x.p);
            if (!x.conv) {   // residual / identity columns: dX += dY[:, off : off + c]
                const long long off = x.off;
                const int c = x.c;
                bpush([=](cudaStream_t st) { return launch_grad_add2d(dy + off, n, gx, c, tokens, c, acc ? 1 : 0, st); }, "residual_add");
                continue;
            }
            // segment that holds this source's channel range
            const Seg* sg = nullptr;
            seg_start = 0;
            for (auto& cand : pc->segs) {
                if (x.off >= seg_start && x.off + x.c <= seg_start + cand.ci) { sg = &cand; break; }
                seg_start += cand.ci;
            }
            if (!sg) { set_err(fail(MMD_ESTATE, "internal: source does not fit a weight segment")); return; }
            const int T = sg->T, Ci = sg->ci, c_lo = static_cast<int>(x.off - seg_start), cs = x.c;
            if (n % 64 != 0 || cs % 64 != 0) { set_err(fail(MMD_EINVAL, "backward: channels %d -> %d not multiples of 64", cs, n)); return; }
            const long long kt = static_cast<long long>(T) * n;
            act_t* wt = static_cast<act_t*>(tpack.take(sizeof(act_t) * static_cast<size_t>(cs) * kt));
            act_t* tmp = acc ? static_cast<act_t*>(bscratch.take(sizeof(act_t) * static_cast<size_t>(tokens) * cs)) : nullptr;
            if (emitting()) {
                const float* w32p = pf(sg->w);
                const int Co = n;
                plan->tpack_ops.push_back([=](cudaStream_t st) { return launch_pack_weight_t(w32p, wt, Co, Ci, T, c_lo, cs, kt, 0, st); });
                GemmProblem pr;
                pr.g = op.g;
                pr.n_src = 1; pr.src[0] = dy; pr.src_c[0] = n;
                pr.n_taps = static_cast<int>(op.taps.size());
                for (size_t t = 0; t < op.taps.size(); ++t) for (int j = 0; j < 3; ++j) pr.taps[t][j] = -op.taps[t][j];
                pr.w = wt; pr.bias = plan->zero_bias; pr.n = cs; pr.bn = pick_bn(cs);
                {
                    long long mt = 1;
                    for (int i = 0; i < 4; ++i) mt *= (op.g.dims[i] + op.g.box[i] - 1) / op.g.box[i];
                    if (wide_n && pr.bn == 128 && cs % 256 == 0 && mt * (cs / 256) >= 2LL * num_sms()) pr.bn = 256;
                }
                pr.out = acc ? tmp : gx;
                auto gp = std::make_shared<GemmParams>();
                int r = build_gemm(pr, gp.get());
                if (r != MMD_OK) { set_err(r); return; }
                const int bn = pr.bn;
                const long long elems = tokens * cs;
                bpush([=](cudaStream_t st) -> int {
                    MMD_TRY(launch_gemm(*gp, bn, st));
                    pdl_break(st);
                    if (acc) return launch_grad_add(tmp, gx, elems, 1, st);
                    return MMD_OK;
                }, "dgrad:" + op.tag);
            }
        }
        release_b(bmark);
    }
    void release_b(size_t mark) { bscratch.top = mark; }

    void bwd_gn(const TapeOp& op) {
        const int C = op.c1 + op.c2;
        act_t* dy = need_grad(op.y, "group_norm");
        if (!dy) return;
        const size_t per = static_cast<size_t>(op.ns) * op.rows;
        act_t* g1 = grad_buf(op.x1, per * op.c1);
        const bool a1 = grad_written(op.x1);
        mark_written(op.x1);
        act_t* g2 = nullptr;
        bool a2 = false;
        if (op.x2) { g2 = grad_buf(op.x2, per * op.c2); a2 = grad_written(op.x2); mark_written(op.x2); }
        const size_t bmark = bscratch.top;
        const size_t t_bytes = sizeof(float) * 2 * C * op.ns;
        float* T = static_cast<float*>(bscratch.take(t_bytes));
        if (emitting()) {
            GnBwdProblem pr;
            pr.s = GnSrc{op.x1, op.c1, op.c1, op.x2, op.c2, op.c2};
            pr.ns = op.ns; pr.rows = op.rows; pr.sums = op.sums; pr.nsub = op.nsub; pr.stat_rows = op.stat_rows;
            pr.gamma = pf(op.gn_g); pr.beta = pf(op.gn_b);
            pr.film = op.emb_row0 >= 0 ? emb_all + op.emb_row0 : nullptr;
            pr.film_ld = m.emb_rows; pr.ns_per_batch = op.ns_per_batch; pr.silu = op.silu;
            pr.dy = dy; pr.T = T;
            pr.out = GnBwdOut{g1, op.c1, a1 ? 1 : 0, g2, op.c2, a2 ? 1 : 0};
            pr.dgamma = g32_of(op.gn_g); pr.dbeta = g32_of(op.gn_b);
            pr.dfilm = op.emb_row0 >= 0 ? plan->d_emb_all + op.emb_row0 : nullptr;
            pr.gscale = plan->gscale;
            if (op.drop_site >= 0) { pr.drop = plan->drop_dev; pr.drop_site = static_cast<uint32_t>(op.drop_site); }
            bpush([=](cudaStream_t st) -> int {
                MMD_CUDA_OK(cudaMemsetAsync(T, 0, t_bytes, st));
                return launch_gn_bwd(pr, st);
            }, "group_norm_bwd");
        }
        release_b(bmark);
    }

    void bwd_gn_temporal(const TapeOp& op) {
        act_t* dy = need_grad(op.y, "temporal group_norm");
        if (!dy) return;
        const size_t elems = static_cast<size_t>(op.B) * op.F * op.P * op.C;
        act_t* gx = grad_buf(op.in, elems);
        const bool acc = grad_written(op.in);
        mark_written(op.in);
        const size_t bmark = bscratch.top;
        act_t* tmp = acc ? static_cast<act_t*>(bscratch.take(elems * sizeof(act_t))) : nullptr;
        if (emitting()) {
            const act_t* x = op.in;
            const float* gamma = pf(op.gn_g);
            float* dg = g32_of(op.gn_g);
            float* db = g32_of(op.gn_b);
            const float* gs = plan->gscale;
            const int Bc = op.B, Fc = op.F, P = op.P, C = op.C;
            bpush([=](cudaStream_t st) -> int {
                MMD_TRY(launch_gn_temporal_bwd(x, dy, acc ? tmp : gx, gamma, dg, db, Bc, Fc, P, C, gs, st));
                if (acc) return launch_grad_add(tmp, gx, static_cast<long long>(elems), 1, st);
                return MMD_OK;
            }, "gn_temporal_bwd");
        }
        release_b(bmark);
    }

    void bwd_tattn(const TapeOp& op) {
        act_t* d_out = need_grad(op.y, "temporal attention");
        if (!d_out) return;
        const size_t tokens = static_cast<size_t>(op.B) * op.F * op.P;
        act_t* dqkv = grad_buf(op.in, tokens * 3 * op.C);
        if (grad_written(op.in)) { set_err(fail(MMD_ESTATE, "internal: qkv gradient already written")); return; }
        mark_written(op.in);
        if (!emitting()) return;
        const act_t* qkv = op.in;
        const int Bc = op.B, Fc = op.F, P = op.P, C = op.C, heads = op.heads;
        bpush([=](cudaStream_t st) { return launch_temporal_attn_bwd(qkv, d_out, dqkv, Bc, Fc, P, C, heads, st); }, "temporal_attention_bwd");
    }

    void bwd_attn(const TapeOp& op) {
        const AttnProblem& ap = op.ap;
        act_t* d_out = need_grad(ap.out, "attention");
        if (!d_out) return;
        const int C = ap.heads * ap.d;
        act_t* gq = grad_buf(ap.q, static_cast<size_t>(ap.q_rows) * ap.q_ld);
        act_t* gk = grad_buf(ap.k, static_cast<size_t>(ap.k_rows) * ap.k_ld);
        mark_written(ap.q);
        mark_written(ap.k);
        const size_t bmark = bscratch.top;
        float* delta = static_cast<float*>(bscratch.take(sizeof(float) * ap.heads * (ap.q_rows + 128)));
        if (emitting()) {
            auto pq = std::make_shared<AttnBwdParams>();
            auto pkv = std::make_shared<AttnBwdParams>();
            AttnBwdOut o{gq, ap.q_ld, ap.q_col0, gk, ap.k_ld, ap.k_col0, gk, ap.k_ld, ap.v_col0};
            int r = build_attn_bwd(ap, d_out, C, op.lse, delta, ap.q_rows, o, pq.get(), pkv.get());
            if (r != MMD_OK) { set_err(r); return; }
            const act_t* out = ap.out;
            const long long q_rows = ap.q_rows;
            const int heads = ap.heads, d = ap.d;
            bpush([=](cudaStream_t st) -> int {
                MMD_TRY(launch_attn_delta(d_out, out, q_rows, C, heads, delta, q_rows, st));
                return launch_attn_bwd(*pq, *pkv, d, st);
            }, (ap.win == 1 && ap.shift_dev == nullptr && ap.q_blk == ap.k_blk) ? "self_attention_bwd" : "cross_attention_bwd");
        }
        release_b(bmark);
    }

    void bwd_resample(const TapeOp& op) {
        act_t* dy = need_grad(op.y, "resample");
        if (!dy) return;
        const size_t in_elems = (op.mode == 0 || op.mode == 2) ? static_cast<size_t>(op.n_) * op.h_ * op.w_ * op.C
                                                                : static_cast<size_t>(op.n_) * op.h_ * op.C;
        act_t* gx = grad_buf(op.in, in_elems);
        const bool acc = grad_written(op.in);
        mark_written(op.in);
        if (!emitting()) return;
        const int mode = op.mode, n = op.n_, h = op.h_, w = op.w_, c = op.C;
        bpush([=](cudaStream_t st) { return launch_resample_bwd(dy, gx, mode, n, h, w, c, acc ? 1 : 0, st); }, "resample_bwd");
    }

    void emit_backward() {
        if (bad()) return;
        for (auto it = tape.rbegin(); it != tape.rend() && !bad(); ++it) {
            switch (it->kind) {
                case TapeOp::GEMM: bwd_gemm(*it); break;
                case TapeOp::GN: bwd_gn(*it); break;
                case TapeOp::GN_TEMPORAL: bwd_gn_temporal(*it); break;
                case TapeOp::TATTN: bwd_tattn(*it); break;
                case TapeOp::ATTN: bwd_attn(*it); break;
                case TapeOp::RESAMPLE: bwd_resample(*it); break;
            }
        }
        if (bad()) return;
        // ---- FiLM table -> emb_layers -> time_embed MLP (all fp32)
        const int E = cfg().model_channels, rows = m.emb_rows;
        const size_t bmark = bscratch.top;
        float* dw_stack = static_cast<float*>(bscratch.take(sizeof(float) * (static_cast<size_t>(rows) * E + rows)));
        float* dsilu = static_cast<float*>(bscratch.take(sizeof(float) * B * E));
        float* te_scratch = static_cast<float*>(bscratch.take(sizeof(float) * B * 4 * E));
        if (emitting()) {
            const float* demb = plan->d_emb_all;
            const float* se = silu_emb;
            const float* ew = m.bpk + m.emb_w_off;
            const float* gs = plan->gscale;
            const int Bc = B;
            float* db_stack = dw_stack + static_cast<size_t>(rows) * E;
            std::vector<EmbBlk> blks = emb_blocks;
            std::vector<float*> gw, gb;
            for (auto& eb : blks) { gw.push_back(g32_of(eb.w)); gb.push_back(g32_of(eb.b)); }
            const float *w1 = pf(te_idx[0]), *b1 = pf(te_idx[1]), *w2 = pf(te_idx[2]), *b2 = pf(te_idx[3]);
            float *dw1 = g32_of(te_idx[0]), *db1 = g32_of(te_idx[1]), *dw2 = g32_of(te_idx[2]), *db2 = g32_of(te_idx[3]);
            const float* tdev = plan->t_dev;
            bpush([=](cudaStream_t st) -> int {
                MMD_CUDA_OK(cudaMemsetAsync(dw_stack, 0, sizeof(float) * (static_cast<size_t>(rows) * E + rows), st));
                const long long tot = static_cast<long long>(rows) * E;
                emb_layers_bwd_w_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, st>>>(demb, se, Bc, E, rows, dw_stack, db_stack, gs);
                MMD_CUDA_OK(cudaGetLastError());
                for (size_t i = 0; i < blks.size(); ++i) {
                    const long long nw = static_cast<long long>(blks[i].rows) * E;
                    axpy_f32_kernel<<<static_cast<unsigned>((nw + 255) / 256), 256, 0, st>>>(dw_stack + static_cast<size_t>(blks[i].row0) * E, gw[i], nw, nullptr);
                    axpy_f32_kernel<<<(blks[i].rows + 255) / 256, 256, 0, st>>>(db_stack + blks[i].row0, gb[i], blks[i].rows, nullptr);
                }
                MMD_CUDA_OK(cudaGetLastError());
                emb_layers_bwd_x_kernel<<<dim3((E + 31) / 32, Bc), 256, 0, st>>>(demb, ew, E, rows, dsilu);
                MMD_CUDA_OK(cudaGetLastError());
                time_embed_bwd_kernel<<<Bc, E, 4 * E * sizeof(float), st>>>(tdev, w1, b1, w2, b2, E, dsilu, te_scratch);
                MMD_CUDA_OK(cudaGetLastError());
                time_embed_bwd_reduce_kernel<<<(E * E + 255) / 256, 256, 0, st>>>(te_scratch, Bc, E, dw1, db1, dw2, db2, gs);
                MMD_CUDA_OK(cudaGetLastError());
                return MMD_OK;
            }, "time_embed_bwd");
        }
        release_b(bmark);
    }
};

// Device memory is allocated on first use so that the parameter inventory (names / shapes / shift bounds)
// can be queried on a machine without a GPU.
static int ensure_device(MmdModel* m) {
    if (m->w32) return MMD_OK;
    MMD_CUDA_OK(cudaMalloc(&m->w32, sizeof(float) * std::max<size_t>(m->w32_floats, 4)));
    MMD_CUDA_OK(cudaMalloc(&m->wpk, sizeof(act_t) * std::max<size_t>(m->wpk_halves, 8)));
    MMD_CUDA_OK(cudaMalloc(&m->bpk, sizeof(float) * std::max<size_t>(m->bpk_floats, 4)));
    MMD_CUDA_OK(cudaMemset(m->wpk, 0, sizeof(act_t) * std::max<size_t>(m->wpk_halves, 8)));
    MMD_CUDA_OK(cudaMemset(m->bpk, 0, sizeof(float) * std::max<size_t>(m->bpk_floats, 4)));
    MMD_CUDA_OK(cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking));
    MMD_CUDA_OK(cudaStreamCreateWithFlags(&m->cap_stream2, cudaStreamNonBlocking));
    return MMD_OK;
}

static int validate_cfg(const MmdConfig& c) {
    if (c.n_levels < 1 || c.n_levels > MMD_MAX_LEVELS || c.n_cross < 0 || c.n_cross > MMD_MAX_LEVELS ||
        c.n_video_attn < 0 || c.n_video_attn > MMD_MAX_LEVELS || c.n_audio_attn < 0 || c.n_audio_attn > MMD_MAX_LEVELS)
        return fail(MMD_EINVAL, "config list lengths out of range");
    if (c.model_channels % 64 != 0) return fail(MMD_EINVAL, "model_channels %d must be a multiple of 64 (64-channel K chunks)", c.model_channels);
    if (c.video_f != 8 && c.video_f != 16) return fail(MMD_EINVAL, "video frames %d unsupported (8 or 16)", c.video_f);
    if (c.max_batch < 1) return fail(MMD_EINVAL, "max_batch");
    const int down = 1 << (c.n_levels - 1);
    if (c.video_h % down || c.video_w % down) return fail(MMD_EINVAL, "video size not divisible by %d", down);
    long long adown = 1;
    for (int i = 1; i < c.n_levels; ++i) adown *= 4;
    if (c.audio_l % adown) return fail(MMD_EINVAL, "audio length not divisible by %lld", adown);
    if ((c.video_w & (c.video_w - 1)) || (c.video_h & (c.video_h - 1))) return fail(MMD_EINVAL, "video H/W must be powers of two");
    // QKVAttention lets the last audio segment absorb the remainder when the audio length at a cross-attention level is
    // not a multiple of the frame count (multimodal_unet.py:547-548, 644-645).  The window kernels here take equal
    // segments (every shipped configuration: 25600 / 4^l is a multiple of 16), so such a configuration is rejected when
    // the model is created instead of at the first forward.
    {
        long long L = c.audio_l;
        for (int level = 0, ds = 1; level < c.n_levels; ++level, ds *= 2) {
            bool cross = false;
            for (int i = 0; i < c.n_cross; ++i) cross = cross || c.cross_attention_resolutions[i] == ds;
            if (cross && L % c.video_f != 0)
                return fail(MMD_EINVAL, "audio length %lld at cross-attention resolution %d is not a multiple of the %d frames: "
                            "the reference's remainder segment (multimodal_unet.py:547-548) is not supported", L, ds, c.video_f);
            L /= 4;
        }
    }
    return MMD_OK;
}

static int build_plan(MmdModel* m, int B, Plan** out) {
    auto it = m->plans.find(B);
    if (it != m->plans.end()) { *out = it->second.get(); return MMD_OK; }
    if (B < 1 || B > m->cfg.max_batch) return fail(MMD_EINVAL, "batch %d outside [1, %d]", B, m->cfg.max_batch);
    const MmdConfig& c = m->cfg;
    // pass 1: dry run to size the arenas
    Walker dry(*m, false);
    dry.B = B;
    dry.walk();
    if (dry.bad()) return dry.err;
    auto plan = std::make_unique<Plan>();
    plan->B = B;
    const size_t vin = sizeof(float) * B * c.video_f * c.video_c * c.video_h * c.video_w;
    const size_t ain = sizeof(float) * B * c.audio_c * c.audio_l;
    const size_t vout = sizeof(float) * B * c.video_f * c.video_out_channels * c.video_h * c.video_w;
    const size_t aout = sizeof(float) * B * c.audio_out_channels * c.audio_l;
    auto al = [](size_t x) { return (x + 1023) & ~size_t(1023); };
    const size_t io_bytes = al(vin) + al(ain) + al(vout) + al(aout) + al(sizeof(float) * B) + al(sizeof(int) * 64);
    const size_t p_bytes = al(dry.persist.peak) + 1024, s_bytes = al(dry.scratch.peak) + 1024;
    const size_t sa_bytes = al(dry.scratch_a.peak) + 1024;
    const size_t st_bytes = al(dry.stats.peak) + 1024;
    plan->ws_bytes = io_bytes + p_bytes + s_bytes + sa_bytes + st_bytes;
    MMD_CUDA_OK(cudaMalloc(&plan->ws, plan->ws_bytes));
    MMD_CUDA_OK(cudaMemset(plan->ws, 0, plan->ws_bytes));
    uint8_t* q = plan->ws;
    plan->in_video = reinterpret_cast<float*>(q); q += al(vin);
    plan->in_audio = reinterpret_cast<float*>(q); q += al(ain);
    plan->out_video = reinterpret_cast<float*>(q); q += al(vout);
    plan->out_audio = reinterpret_cast<float*>(q); q += al(aout);
    plan->t_dev = reinterpret_cast<float*>(q); q += al(sizeof(float) * B);
    plan->shifts_dev = reinterpret_cast<int*>(q); q += al(sizeof(int) * 64);
    Walker w(*m, false);
    w.B = B;
    w.plan = plan.get();
    w.persist.base = q; w.persist.cap = p_bytes;
    w.scratch.base = q + p_bytes; w.scratch.cap = s_bytes;
    w.scratch_a.base = q + p_bytes + s_bytes; w.scratch_a.cap = sa_bytes;
    w.stats.base = q + p_bytes + s_bytes + sa_bytes; w.stats.cap = st_bytes;
    w.walk();
    if (w.bad()) return w.err;
    if (w.persist.peak > p_bytes || w.scratch.peak > s_bytes || w.scratch_a.peak > sa_bytes || w.stats.peak > st_bytes)
        return fail(MMD_ESTATE, "internal: arena overflow");
    *out = plan.get();
    m->plans[B] = std::move(plan);
    return MMD_OK;
}


// Training plan: same topology walk with nothing recycled, plus the backward launch list built from the tape.
static int build_train_plan(MmdModel* m, int B, Plan** out) {
    auto it = m->train_plans.find(B);
    if (it != m->train_plans.end()) { *out = it->second.get(); return MMD_OK; }
    if (B < 1 || B > m->cfg.max_batch) return fail(MMD_EINVAL, "batch %d outside [1, %d]", B, m->cfg.max_batch);
    const MmdConfig& c = m->cfg;
    Walker dry(*m, false);
    dry.B = B;
    dry.train = true;
    dry.walk();
    dry.emit_backward();
    if (dry.bad()) return dry.err;
    auto plan = std::make_unique<Plan>();
    plan->B = B;
    plan->train = true;
    const size_t vin = sizeof(float) * B * c.video_f * c.video_c * c.video_h * c.video_w;
    const size_t ain = sizeof(float) * B * c.audio_c * c.audio_l;
    const size_t vout = sizeof(float) * B * c.video_f * c.video_out_channels * c.video_h * c.video_w;
    const size_t aout = sizeof(float) * B * c.audio_out_channels * c.audio_l;
    auto al = [](size_t x) { return (x + 1023) & ~size_t(1023); };
    const size_t g32_bytes = al(sizeof(float) * std::max<size_t>(m->w32_floats, 4));
    const size_t demb_bytes = al(sizeof(float) * B * std::max(m->emb_rows, 1));
    const size_t zb_bytes = al(sizeof(float) * 4096);
    const size_t io_bytes = 2 * (al(vin) + al(ain) + al(vout) + al(aout)) + al(sizeof(float) * B) + al(sizeof(int) * 64) + 1024 +
                            g32_bytes + demb_bytes + zb_bytes;
    const size_t p_bytes = al(dry.persist.peak) + 1024, s_bytes = al(dry.scratch.peak) + 1024;
    const size_t sa_bytes = al(dry.scratch_a.peak) + 1024, st_bytes = al(dry.stats.peak) + 1024;
    const size_t g_bytes = al(dry.grads.peak) + 1024, b_bytes = al(dry.bscratch.peak) + 1024;
    const size_t t_bytes = al(dry.tpack.peak) + 1024;
    plan->ws_bytes = io_bytes + p_bytes + s_bytes + sa_bytes + st_bytes + g_bytes + b_bytes;
    if (cudaMalloc(&plan->ws, plan->ws_bytes) != cudaSuccess) {
        // a training plan keeps every intermediate (tens of GB): before giving up, drop the cached plans of other batch
        // sizes (e.g. the ragged last batch of an epoch) and try once more
        cudaGetLastError();
        plan->ws = nullptr;
        MMD_CUDA_OK(cudaDeviceSynchronize());
        m->train_plans.clear();
        m->plans.clear();
        if (cudaMalloc(&plan->ws, plan->ws_bytes) != cudaSuccess) {
            cudaGetLastError();
            plan->ws = nullptr;
            return fail(MMD_ECUDA, "training plan for batch %d needs %.1f GB of device memory (kept activations + gradients)", B,
                        plan->ws_bytes / 1e9);
        }
    }
    MMD_CUDA_OK(cudaMemset(plan->ws, 0, plan->ws_bytes));
    MMD_CUDA_OK(cudaMalloc(&plan->wpk_t, t_bytes));
    MMD_CUDA_OK(cudaMemset(plan->wpk_t, 0, t_bytes));
    uint8_t* q = plan->ws;
    plan->in_video = reinterpret_cast<float*>(q); q += al(vin);
    plan->in_audio = reinterpret_cast<float*>(q); q += al(ain);
    plan->out_video = reinterpret_cast<float*>(q); q += al(vout);
    plan->out_audio = reinterpret_cast<float*>(q); q += al(aout);
    plan->d_in_video = reinterpret_cast<float*>(q); q += al(vin);
    plan->d_in_audio = reinterpret_cast<float*>(q); q += al(ain);
    plan->d_out_video = reinterpret_cast<float*>(q); q += al(vout);
    plan->d_out_audio = reinterpret_cast<float*>(q); q += al(aout);
    plan->t_dev = reinterpret_cast<float*>(q); q += al(sizeof(float) * B);
    plan->shifts_dev = reinterpret_cast<int*>(q); q += al(sizeof(int) * 64);
    plan->gscale = reinterpret_cast<float*>(q);
    plan->amax_bits = reinterpret_cast<unsigned int*>(q + 64);
    plan->drop_dev = reinterpret_cast<DropState*>(q + 128); q += 1024;
    plan->g32 = reinterpret_cast<float*>(q); q += g32_bytes;
    plan->d_emb_all = reinterpret_cast<float*>(q); q += demb_bytes;
    plan->zero_bias = reinterpret_cast<float*>(q); q += zb_bytes;
    Walker w(*m, false);
    w.B = B;
    w.train = true;
    w.plan = plan.get();
    w.persist.base = q; w.persist.cap = p_bytes; q += p_bytes;
    w.scratch.base = q; w.scratch.cap = s_bytes; q += s_bytes;
    w.scratch_a.base = q; w.scratch_a.cap = sa_bytes; q += sa_bytes;
    w.stats.base = q; w.stats.cap = st_bytes; q += st_bytes;
    w.grads.base = q; w.grads.cap = g_bytes; q += g_bytes;
    w.bscratch.base = q; w.bscratch.cap = b_bytes; q += b_bytes;
    w.tpack.base = reinterpret_cast<uint8_t*>(plan->wpk_t); w.tpack.cap = t_bytes;
    w.walk();
    w.emit_backward();
    if (w.bad()) return w.err;
    if (w.persist.peak > p_bytes || w.scratch.peak > s_bytes || w.scratch_a.peak > sa_bytes || w.stats.peak > st_bytes ||
        w.grads.peak > g_bytes || w.bscratch.peak > b_bytes || w.tpack.peak > t_bytes)
        return fail(MMD_ESTATE, "internal: training arena overflow");
    *out = plan.get();
    m->train_plans[B] = std::move(plan);
    return MMD_OK;
}

static int stage_inputs(MmdModel* m, Plan* plan, int batch, const float* video_in, const float* audio_in, const float* timesteps,
                        const int32_t* shifts, cudaStream_t st) {
    const MmdConfig& c = m->cfg;
    const size_t vin = sizeof(float) * batch * c.video_f * c.video_c * c.video_h * c.video_w;
    const size_t ain = sizeof(float) * batch * c.audio_c * c.audio_l;
    MMD_CUDA_OK(cudaMemcpyAsync(plan->in_video, video_in, vin, cudaMemcpyDeviceToDevice, st));
    MMD_CUDA_OK(cudaMemcpyAsync(plan->in_audio, audio_in, ain, cudaMemcpyDeviceToDevice, st));
    MMD_CUDA_OK(cudaMemcpyAsync(plan->t_dev, timesteps, sizeof(float) * batch, cudaMemcpyDeviceToDevice, st));
    ShiftArgs sa{};
    sa.n = static_cast<int>(m->shift_bounds.size());
    for (int i = 0; i < sa.n; ++i) {
        int v = shifts ? shifts[i] : 0;
        const int hi = m->shift_bounds[i];
        if (hi < 0) v = 0;
        else if (v < 0 || v > hi) return fail(MMD_EINVAL, "shift %d of block %d outside [0, %d]", v, i, hi);
        sa.v[i] = v;
    }
    if (sa.n > 0) {
        set_shifts_kernel<<<1, 64, 0, st>>>(sa, plan->shifts_dev);
        MMD_CUDA_OK(cudaGetLastError());
    }
    return MMD_OK;
}

static size_t dry_workspace(MmdModel* m, int B) {
    Walker dry(*m, false);
    dry.B = B;
    dry.walk();
    if (dry.bad()) return 0;
    return dry.persist.peak + dry.scratch.peak + dry.scratch_a.peak + dry.stats.peak + (1 << 20);
}

// Capture the forward launch list of a plan into a CUDA graph.  Two-branch capture: video / main steps on cap_stream,
// audio steps on cap_stream2; "sync" steps make the branches wait for each other (graph edges), so the small audio
// kernels overlap the video ones.
static int capture_forward_graph(MmdModel* m, Plan* plan, cudaStream_t st) {
    // pack ops (if any) ran on `st`; make sure the capture stream sees a quiescent device
    MMD_CUDA_OK(cudaStreamSynchronize(st));
    cudaGraph_t g = nullptr;
    const bool two = m->two_streams;
    auto new_event = [&]() -> cudaEvent_t {
        cudaEvent_t ev = nullptr;
        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        plan->events.push_back(ev);
        return ev;
    };
    MMD_CUDA_OK(cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal));
    PdlScope pdl(m->cap_stream, two ? m->cap_stream2 : nullptr);
    int r = MMD_OK;
    cudaError_t ce = cudaSuccess;
    auto cross_sync = [&]() {
        pdl_break_all();   // the kernels after a join have two predecessors: plain launches
        cudaEvent_t e0 = new_event(), e1 = new_event();
        ce = cudaEventRecord(e0, m->cap_stream);
        if (ce == cudaSuccess) ce = cudaEventRecord(e1, m->cap_stream2);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(m->cap_stream, e1, 0);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(m->cap_stream2, e0, 0);
    };
    if (two) {   // fork: pull the second stream into the capture
        cudaEvent_t e0 = new_event();
        ce = cudaEventRecord(e0, m->cap_stream);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(m->cap_stream2, e0, 0);
    }
    for (size_t i = 0; i < plan->steps.size() && r == MMD_OK && ce == cudaSuccess; ++i) {
        const int sid = plan->info[i].stream;
        if (sid < 0) { if (two) cross_sync(); continue; }
        r = plan->steps[i]((two && sid == 1) ? m->cap_stream2 : m->cap_stream);
    }
    if (two && ce == cudaSuccess) {   // join (the plan ends with a sync step, this only closes the fork)
        cudaEvent_t e1 = new_event();
        ce = cudaEventRecord(e1, m->cap_stream2);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(m->cap_stream, e1, 0);
    }
    cudaError_t e = cudaStreamEndCapture(m->cap_stream, &g);
    if (ce != cudaSuccess) { if (g) cudaGraphDestroy(g); return fail(MMD_ECUDA, "two-branch capture: %s", cudaGetErrorString(ce)); }
    if (r != MMD_OK) { if (g) cudaGraphDestroy(g); return r; }
    if (e != cudaSuccess) return fail(MMD_ECUDA, "graph capture: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&plan->graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(MMD_ECUDA, "graph instantiate: %s", cudaGetErrorString(e));
    return MMD_OK;
}

// Backward launch list of a training plan as one graph (single branch: the tape order is a valid serial schedule).
static int capture_backward_graph(MmdModel* m, Plan* plan, cudaStream_t st) {
    MMD_CUDA_OK(cudaStreamSynchronize(st));
    cudaGraph_t g = nullptr;
    MMD_CUDA_OK(cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal));
    int r = MMD_OK;
    for (size_t i = 0; i < plan->bwd_steps.size() && r == MMD_OK; ++i) r = plan->bwd_steps[i](m->cap_stream);
    cudaError_t e = cudaStreamEndCapture(m->cap_stream, &g);
    if (r != MMD_OK) { if (g) cudaGraphDestroy(g); return r; }
    if (e != cudaSuccess) return fail(MMD_ECUDA, "backward graph capture: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&plan->bwd_graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(MMD_ECUDA, "backward graph instantiate: %s", cudaGetErrorString(e));
    return MMD_OK;
}

// A list of small independent device operations (weight repacks: distinct destinations, fixed pointers) as ONE graph
// with `branches` parallel chains: a parameter update then costs one graph launch instead of thousands of stream
// launches (~2 600 per update for the production network), and the chains overlap on the device.
static int capture_parallel_ops(MmdModel* m, const std::vector<LaunchFn>& ops, cudaStream_t st, cudaGraphExec_t* out,
                                int branches = 8) {
    MMD_CUDA_OK(cudaStreamSynchronize(st));
    std::vector<cudaStream_t> side(branches - 1, nullptr);
    std::vector<cudaEvent_t> evs;
    auto new_event = [&]() { cudaEvent_t e = nullptr; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); evs.push_back(e); return e; };
    for (auto& s : side) MMD_CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaGraph_t g = nullptr;
    int r = MMD_OK;
    cudaError_t ce = cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal);
    if (ce == cudaSuccess) {
        cudaEvent_t fork = new_event();
        ce = cudaEventRecord(fork, m->cap_stream);
        for (auto& s : side) if (ce == cudaSuccess) ce = cudaStreamWaitEvent(s, fork, 0);
        for (size_t i = 0; i < ops.size() && r == MMD_OK && ce == cudaSuccess; ++i) {
            const size_t b = i % branches;
            r = ops[i](b == 0 ? m->cap_stream : side[b - 1]);
        }
        for (auto& s : side) {
            if (ce != cudaSuccess) break;
            cudaEvent_t j = new_event();
            ce = cudaEventRecord(j, s);
            if (ce == cudaSuccess) ce = cudaStreamWaitEvent(m->cap_stream, j, 0);
        }
        const cudaError_t e = cudaStreamEndCapture(m->cap_stream, &g);
        if (ce == cudaSuccess) ce = e;
    }
    for (auto& s : side) if (s) cudaStreamDestroy(s);
    for (auto& e : evs) cudaEventDestroy(e);
    if (r != MMD_OK) { if (g) cudaGraphDestroy(g); return r; }
    if (ce != cudaSuccess) { if (g) cudaGraphDestroy(g); cudaGetLastError(); return fail(MMD_ECUDA, "repack graph capture: %s", cudaGetErrorString(ce)); }
    ce = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    if (ce != cudaSuccess) return fail(MMD_ECUDA, "repack graph instantiate: %s", cudaGetErrorString(ce));
    return MMD_OK;
}

static bool repack_graphs_enabled() {
    static const bool on = [] { const char* e = getenv("MMD_REPACK_GRAPH"); return !(e && e[0] == '0'); }();
    return on;
}

// Re-derive every packed weight layout after a parameter update: the forward's K-major fp16 packs now, the transposed
// packs of the training plans lazily at their next backward (whichever entry point notices the update first).
static int repack_if_dirty(MmdModel* m, cudaStream_t st) {
    if (!m->dirty) return MMD_OK;
    if (m->use_graph && repack_graphs_enabled()) {
        if (!m->pack_graph) MMD_TRY(capture_parallel_ops(m, m->pack_ops, st, &m->pack_graph));
        MMD_CUDA_OK(cudaGraphLaunch(m->pack_graph, st));
    } else {
        for (auto& op : m->pack_ops) MMD_TRY(op(st));
    }
    m->dirty = false;
    for (auto& tp : m->train_plans) tp.second->tpack_dirty = true;
    return MMD_OK;
}

static bool train_graphs_enabled() {
    static const bool on = [] { const char* e = getenv("MMD_TRAIN_GRAPH"); return !(e && e[0] == '0'); }();
    return on;
}

}  // namespace mmd


// ===========================================================================
extern "C" {

int mmd_model_create(const MmdConfig* cfg, MmdModel** out) {
    if (!cfg || !out) return fail(MMD_EINVAL, "null argument");
    MMD_TRY(validate_cfg(*cfg));
    auto m = std::make_unique<MmdModel>();
    m->cfg = *cfg;
    const char* ng = getenv("MMD_NO_GRAPH");
    m->use_graph = !(ng && ng[0] == '1');
    const char* os = getenv("MMD_ONE_STREAM");
    m->two_streams = !(os && os[0] == '1');
    Walker w(*m, true);
    w.walk();
    if (w.bad()) return w.err;
    if (m->shift_bounds.size() > 64) return fail(MMD_EINVAL, "too many cross-attention blocks");
    // stacked emb_layers weights/biases live in the fp32 packed arena
    m->emb_w_off = w.bpk_top;
    w.bpk_top += static_cast<size_t>(m->emb_rows) * cfg->model_channels;
    m->emb_b_off = w.bpk_top;
    w.bpk_top += m->emb_rows;
    m->w32_floats = w.w32_top;
    m->wpk_halves = w.wpk_top;
    m->bpk_floats = w.bpk_top;
    *out = m.release();
    return MMD_OK;
}

int mmd_model_destroy(MmdModel* m) {
    delete m;
    return MMD_OK;
}

int mmd_model_num_params(const MmdModel* m) { return m ? static_cast<int>(m->params.size()) : 0; }

int mmd_model_param_info(const MmdModel* m, int index, const char** name, int* ndim, int64_t shape[5]) {
    if (!m || index < 0 || index >= static_cast<int>(m->params.size())) return fail(MMD_EINVAL, "param index %d", index);
    const ParamInfo& p = m->params[index];
    if (name) *name = p.name.c_str();
    if (ndim) *ndim = static_cast<int>(p.shape.size());
    if (shape) for (size_t i = 0; i < p.shape.size() && i < 5; ++i) shape[i] = p.shape[i];
    return MMD_OK;
}

int mmd_model_set_param(MmdModel* m, const char* name, const float* data, int64_t numel, void* stream) {
    if (!m || !name || !data) return fail(MMD_EINVAL, "null argument");
    auto it = m->param_index.find(name);
    if (it == m->param_index.end()) return fail(MMD_ENOTFOUND, "unknown parameter %s", name);
    MMD_TRY(ensure_device(m));
    ParamInfo& p = m->params[it->second];
    if (numel != p.numel) return fail(MMD_EINVAL, "parameter %s: %lld elements given, %lld expected", name, (long long)numel, (long long)p.numel);
    MMD_CUDA_OK(cudaMemcpyAsync(m->w32 + p.offset, data, sizeof(float) * numel, cudaMemcpyDeviceToDevice,
                                static_cast<cudaStream_t>(stream)));
    p.set = true;
    m->dirty = true;
    return MMD_OK;
}

/* Every parameter at once: `flat` holds the parameters at the float offsets of mmd_model_param_offset (the layout of
 * the flat gradient buffer), `n_floats` == mmd_model_param_floats.  One device copy instead of one call per tensor. */
int mmd_model_set_params_flat(MmdModel* m, const float* flat, int64_t n_floats, void* stream) {
    if (!m || !flat) return fail(MMD_EINVAL, "null argument");
    if (n_floats != static_cast<int64_t>(m->w32_floats))
        return fail(MMD_EINVAL, "flat parameter buffer has %lld floats, %lld expected", (long long)n_floats, (long long)m->w32_floats);
    MMD_TRY(ensure_device(m));
    MMD_CUDA_OK(cudaMemcpyAsync(m->w32, flat, sizeof(float) * m->w32_floats, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    for (auto& p : m->params) p.set = true;
    m->dirty = true;
    return MMD_OK;
}

int mmd_model_num_shifts(const MmdModel* m) { return m ? static_cast<int>(m->shift_bounds.size()) : 0; }
int mmd_model_shift_bound(const MmdModel* m, int index) {
    if (!m || index < 0 || index >= static_cast<int>(m->shift_bounds.size())) return -1;
    return m->shift_bounds[index];
}

size_t mmd_model_workspace_bytes(const MmdModel* m, int batch) {
    if (!m) return 0;
    return dry_workspace(const_cast<MmdModel*>(m), batch);
}

/* Device bytes of a training plan (kept activations + activation gradients + backward temporaries + transposed weight
 * packs); host-only dry run of the forward walk AND the backward tape, 0 on error (see mmd_last_error). */
size_t mmd_model_train_workspace_bytes(const MmdModel* m, int batch) {
    if (!m) return 0;
    Walker dry(*const_cast<MmdModel*>(m), false);
    dry.B = batch;
    dry.train = true;
    dry.walk();
    dry.emit_backward();
    if (dry.bad()) return 0;
    return dry.persist.peak + dry.scratch.peak + dry.scratch_a.peak + dry.stats.peak + dry.grads.peak + dry.bscratch.peak +
           dry.tpack.peak + sizeof(float) * m->w32_floats + (1 << 20);
}

int mmd_model_num_launches(const MmdModel* m, int batch) {
    if (!m) return 0;
    auto it = m->plans.find(batch);
    if (it != m->plans.end()) return static_cast<int>(it->second->steps.size());
    auto tt = m->train_plans.find(batch);   // only a training plan exists at this batch: same forward launch list
    return tt == m->train_plans.end() ? 0 : static_cast<int>(tt->second->steps.size());
}

// Per-launch device timing of one forward (no graph): fills ms[i] for plan step i with the mean over `reps`
// back-to-back executions measured with CUDA events on `stream`.  The plan for `batch` must already exist
// (call mmd_model_forward once first).  Returns the number of steps.
int mmd_model_profile(MmdModel* m, int batch, int reps, float* ms, int cap, void* stream) {
    if (!m || !ms) return fail(MMD_EINVAL, "null argument");
    auto it = m->plans.find(batch);
    if (it == m->plans.end()) return fail(MMD_ESTATE, "no plan for batch %d yet (run a forward first)", batch);
    Plan* plan = it->second.get();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n = static_cast<int>(plan->steps.size());
    if (cap < n) return fail(MMD_EINVAL, "profile buffer too small (%d < %d)", cap, n);
    if (reps < 1) reps = 1;
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) MMD_CUDA_OK(cudaEventCreate(&e));
    for (int i = 0; i < n; ++i) ms[i] = 0.f;
    int r = MMD_OK;
    for (int rep = 0; rep < reps && r == MMD_OK; ++rep) {
        MMD_CUDA_OK(cudaEventRecord(ev[0], st));
        for (int i = 0; i < n; ++i) {
            {
                NvtxStep range(plan->info[i].kind.c_str());
                r = plan->steps[i](st);
            }
            if (r != MMD_OK) break;
            MMD_CUDA_OK(cudaEventRecord(ev[i + 1], st));
        }
        MMD_CUDA_OK(cudaStreamSynchronize(st));
        for (int i = 0; i < n && r == MMD_OK; ++i) {
            float t = 0.f;
            MMD_CUDA_OK(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
            ms[i] += t / reps;
        }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    return r == MMD_OK ? n : r;
}

int mmd_model_step_info(const MmdModel* m, int batch, int index, const char** kind, double* flops, double* bytes, int* kernels) {
    if (!m) return fail(MMD_EINVAL, "null argument");
    auto it = m->plans.find(batch);
    const Plan* plan = nullptr;
    if (it != m->plans.end()) plan = it->second.get();
    else {
        auto tt = m->train_plans.find(batch);
        if (tt == m->train_plans.end()) return fail(MMD_ESTATE, "no plan for batch %d", batch);
        plan = tt->second.get();
    }
    if (index < 0 || index >= static_cast<int>(plan->info.size())) return fail(MMD_EINVAL, "step index %d", index);
    const StepInfo& si = plan->info[index];
    if (kind) *kind = si.kind.c_str();
    if (flops) *flops = si.flops;
    if (bytes) *bytes = si.bytes;
    if (kernels) *kernels = si.kernels;
    return MMD_OK;
}

int mmd_model_forward(MmdModel* m, int batch, const float* video_in, const float* audio_in, const float* timesteps,
                      const int32_t* shifts, float* video_out, float* audio_out, void* stream) {
    if (!m || !video_in || !audio_in || !timesteps || !video_out || !audio_out) return fail(MMD_EINVAL, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MMD_TRY(ensure_device(m));
    for (auto& p : m->params)
        if (!p.set) return fail(MMD_ESTATE, "parameter %s was never set", p.name.c_str());
    MMD_TRY(repack_if_dirty(m, st));
    Plan* plan = nullptr;
    MMD_TRY(build_plan(m, batch, &plan));
    const MmdConfig& c = m->cfg;
    const size_t vin = sizeof(float) * batch * c.video_f * c.video_c * c.video_h * c.video_w;
    const size_t ain = sizeof(float) * batch * c.audio_c * c.audio_l;
    const size_t vout = sizeof(float) * batch * c.video_f * c.video_out_channels * c.video_h * c.video_w;
    const size_t aout = sizeof(float) * batch * c.audio_out_channels * c.audio_l;
    MMD_CUDA_OK(cudaMemcpyAsync(plan->in_video, video_in, vin, cudaMemcpyDeviceToDevice, st));
    MMD_CUDA_OK(cudaMemcpyAsync(plan->in_audio, audio_in, ain, cudaMemcpyDeviceToDevice, st));
    MMD_CUDA_OK(cudaMemcpyAsync(plan->t_dev, timesteps, sizeof(float) * batch, cudaMemcpyDeviceToDevice, st));
    ShiftArgs sa{};
    sa.n = static_cast<int>(m->shift_bounds.size());
    for (int i = 0; i < sa.n; ++i) {
        int v = shifts ? shifts[i] : 0;
        const int hi = m->shift_bounds[i];
        if (hi < 0) v = 0;
        else if (v < 0 || v > hi) return fail(MMD_EINVAL, "shift %d of block %d outside [0, %d]", v, i, hi);
        sa.v[i] = v;
    }
    if (sa.n > 0) {
        set_shifts_kernel<<<1, 64, 0, st>>>(sa, plan->shifts_dev);
        MMD_CUDA_OK(cudaGetLastError());
    }
    if (m->use_graph) {
        if (!plan->graph) MMD_TRY(capture_forward_graph(m, plan, st));
        MMD_CUDA_OK(cudaGraphLaunch(plan->graph, st));
    } else {
        PdlScope pdl(st, nullptr);
        for (size_t i = 0; i < plan->steps.size(); ++i) {
            NvtxStep range(plan->info[i].kind.c_str());
            MMD_TRY(plan->steps[i](st));
        }
    }
    MMD_CUDA_OK(cudaMemcpyAsync(video_out, plan->out_video, vout, cudaMemcpyDeviceToDevice, st));
    MMD_CUDA_OK(cudaMemcpyAsync(audio_out, plan->out_audio, aout, cudaMemcpyDeviceToDevice, st));
    return MMD_OK;
}

// ---- training: forward that keeps every intermediate + backward ----
int mmd_model_forward_train(MmdModel* m, int batch, const float* video_in, const float* audio_in, const float* timesteps,
                            const int32_t* shifts, float* video_out, float* audio_out, void* stream) {
    if (!m || !video_in || !audio_in || !timesteps || !video_out || !audio_out) return fail(MMD_EINVAL, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MMD_TRY(ensure_device(m));
    for (auto& p : m->params)
        if (!p.set) return fail(MMD_ESTATE, "parameter %s was never set", p.name.c_str());
    MMD_TRY(repack_if_dirty(m, st));
    Plan* plan = nullptr;
    MMD_TRY(build_train_plan(m, batch, &plan));
    MMD_TRY(stage_inputs(m, plan, batch, video_in, audio_in, timesteps, shifts, st));
    MMD_TRY(launch_set_dropout(plan->drop_dev, m->drop_p, m->drop_seed, st));
    if (m->use_graph && train_graphs_enabled()) {
        if (!plan->graph) MMD_TRY(capture_forward_graph(m, plan, st));
        MMD_CUDA_OK(cudaGraphLaunch(plan->graph, st));
    } else {
        PdlScope pdl(st, nullptr);
        for (size_t i = 0; i < plan->steps.size(); ++i) {
            NvtxStep range(plan->info[i].kind.c_str());
            MMD_TRY(plan->steps[i](st));
        }
    }
    const MmdConfig& c = m->cfg;
    const size_t vout = sizeof(float) * batch * c.video_f * c.video_out_channels * c.video_h * c.video_w;
    const size_t aout = sizeof(float) * batch * c.audio_out_channels * c.audio_l;
    MMD_CUDA_OK(cudaMemcpyAsync(video_out, plan->out_video, vout, cudaMemcpyDeviceToDevice, st));
    MMD_CUDA_OK(cudaMemcpyAsync(audio_out, plan->out_audio, aout, cudaMemcpyDeviceToDevice, st));
    plan->fwd_done = true;
    ++plan->generation;
    return MMD_OK;
}

/* nn.Dropout of the ResBlock out_layers (multimodal_unet.py:376,384; --dropout 0.1 in ssh_scripts/multimodal_train.sh):
 * probability and seed of the NEXT mmd_model_forward_train calls (p = 0, the default, disables it).  The mask of an
 * element is Philox4x32-10(seed, site, element index) >= p, regenerated by the backward. */
int mmd_model_set_dropout(MmdModel* m, float p, uint64_t seed) {
    if (!m) return fail(MMD_EINVAL, "null model");
    if (!(p >= 0.f) || p >= 1.f) return fail(MMD_EINVAL, "dropout probability %f out of [0, 1)", p);
    m->drop_p = p;
    m->drop_seed = seed;
    return MMD_OK;
}
/* Training-forward counter of the plan for `batch` (0 = none yet): a backward belongs to the forward that returned the
 * same value (unet._UNetFunction checks it; a second forward before the backward overwrites the kept activations). */
int64_t mmd_model_train_generation(const MmdModel* m, int batch) {
    if (!m) return 0;
    auto it = m->train_plans.find(batch);
    return it == m->train_plans.end() ? 0 : static_cast<int64_t>(it->second->generation);
}
/* Dropout sites of the training plan in execution order (video then audio out_layers of every ResBlock) and the keep
 * mask (uint8 [rows][channels], channels-last like the activation) the LAST training forward used at site `index`. */
int mmd_model_num_dropout_sites(const MmdModel* m, int batch) {
    if (!m) return 0;
    auto it = m->train_plans.find(batch);
    return it == m->train_plans.end() ? 0 : static_cast<int>(it->second->drop_sites.size());
}
int mmd_model_dropout_site(const MmdModel* m, int batch, int index, int64_t* rows, int* channels, int* modality) {
    if (!m) return fail(MMD_EINVAL, "null model");
    auto it = m->train_plans.find(batch);
    if (it == m->train_plans.end() || index < 0 || index >= static_cast<int>(it->second->drop_sites.size()))
        return fail(MMD_EINVAL, "dropout site %d of batch %d does not exist", index, batch);
    const auto& s = it->second->drop_sites[index];
    if (rows) *rows = s.rows;
    if (channels) *channels = s.C;
    if (modality) *modality = s.modality;
    return MMD_OK;
}
int mmd_model_dropout_mask(const MmdModel* m, int batch, int index, unsigned char* keep, void* stream) {
    if (!m || !keep) return fail(MMD_EINVAL, "null argument");
    auto it = m->train_plans.find(batch);
    if (it == m->train_plans.end() || index < 0 || index >= static_cast<int>(it->second->drop_sites.size()))
        return fail(MMD_EINVAL, "dropout site %d of batch %d does not exist", index, batch);
    const auto& s = it->second->drop_sites[index];
    return launch_dropout_mask(it->second->drop_dev, s.site, s.rows * s.C, keep, static_cast<cudaStream_t>(stream));
}

int mmd_model_backward(MmdModel* m, int batch, const float* d_video_out, const float* d_audio_out, float* param_grads,
                       float* d_video_in, float* d_audio_in, void* stream) {
    if (!m || !d_video_out || !d_audio_out) return fail(MMD_EINVAL, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto it = m->train_plans.find(batch);
    if (it == m->train_plans.end() || !it->second->fwd_done)
        return fail(MMD_ESTATE, "mmd_model_backward: no mmd_model_forward_train at batch %d precedes this call", batch);
    Plan* plan = it->second.get();
    const MmdConfig& c = m->cfg;
    const long long vout_n = static_cast<long long>(batch) * c.video_f * c.video_out_channels * c.video_h * c.video_w;
    const long long aout_n = static_cast<long long>(batch) * c.audio_out_channels * c.audio_l;
    const size_t vin = sizeof(float) * batch * c.video_f * c.video_c * c.video_h * c.video_w;
    const size_t ain = sizeof(float) * batch * c.audio_c * c.audio_l;
    MMD_CUDA_OK(cudaMemcpyAsync(plan->d_out_video, d_video_out, sizeof(float) * vout_n, cudaMemcpyDeviceToDevice, st));
    MMD_CUDA_OK(cudaMemcpyAsync(plan->d_out_audio, d_audio_out, sizeof(float) * aout_n, cudaMemcpyDeviceToDevice, st));
    MMD_CUDA_OK(cudaMemsetAsync(plan->amax_bits, 0, sizeof(unsigned int), st));
    MMD_CUDA_OK(cudaMemsetAsync(plan->g32, 0, sizeof(float) * m->w32_floats, st));
    MMD_CUDA_OK(cudaMemsetAsync(plan->d_emb_all, 0, sizeof(float) * batch * std::max(m->emb_rows, 1), st));
    const int blocks = 2 * num_sms();
    absmax_kernel<<<blocks, 256, 0, st>>>(plan->d_out_video, vout_n, plan->amax_bits);
    absmax_kernel<<<blocks, 256, 0, st>>>(plan->d_out_audio, aout_n, plan->amax_bits);
    make_gscale_kernel<<<1, 1, 0, st>>>(plan->amax_bits, plan->gscale);
    MMD_CUDA_OK(cudaGetLastError());
    if (plan->tpack_dirty) {
        if (m->use_graph && repack_graphs_enabled()) {
            if (!plan->tpack_graph) MMD_TRY(capture_parallel_ops(m, plan->tpack_ops, st, &plan->tpack_graph));
            MMD_CUDA_OK(cudaGraphLaunch(plan->tpack_graph, st));
        } else {
            for (auto& op : plan->tpack_ops) MMD_TRY(op(st));
        }
        plan->tpack_dirty = false;
    }
    if (m->use_graph && train_graphs_enabled()) {
        if (!plan->bwd_graph) MMD_TRY(capture_backward_graph(m, plan, st));
        MMD_CUDA_OK(cudaGraphLaunch(plan->bwd_graph, st));
    } else {
        for (size_t i = 0; i < plan->bwd_steps.size(); ++i) {
            NvtxStep range(plan->bwd_kind[i].c_str());
            MMD_TRY(plan->bwd_steps[i](st));
        }
    }
    if (param_grads)
        MMD_CUDA_OK(cudaMemcpyAsync(param_grads, plan->g32, sizeof(float) * m->w32_floats, cudaMemcpyDeviceToDevice, st));
    if (d_video_in) MMD_CUDA_OK(cudaMemcpyAsync(d_video_in, plan->d_in_video, vin, cudaMemcpyDeviceToDevice, st));
    if (d_audio_in) MMD_CUDA_OK(cudaMemcpyAsync(d_audio_in, plan->d_in_audio, ain, cudaMemcpyDeviceToDevice, st));
    plan->fwd_done = false;
    return MMD_OK;
}

/* Offset (in floats) of parameter `index` inside the flat gradient buffer of mmd_model_backward, and its total length. */
int64_t mmd_model_param_offset(const MmdModel* m, int index) {
    if (!m || index < 0 || index >= static_cast<int>(m->params.size())) return -1;
    return static_cast<int64_t>(m->params[index].offset);
}
int64_t mmd_model_param_floats(const MmdModel* m) { return m ? static_cast<int64_t>(m->w32_floats) : 0; }
int mmd_model_num_backward_launches(const MmdModel* m, int batch) {
    if (!m) return 0;
    auto it = m->train_plans.find(batch);
    return it == m->train_plans.end() ? 0 : static_cast<int>(it->second->bwd_steps.size());
}

// Per-step device time of the backward plan (mean of `reps` executions, CUDA events on `stream`); a forward_train at
// this batch must have run.  Gradients accumulate across repetitions: profiling only.  Returns the step count.
int mmd_model_profile_backward(MmdModel* m, int batch, int reps, float* ms, int cap, void* stream) {
    if (!m || !ms) return fail(MMD_EINVAL, "null argument");
    auto it = m->train_plans.find(batch);
    if (it == m->train_plans.end()) return fail(MMD_ESTATE, "no training plan for batch %d yet", batch);
    Plan* plan = it->second.get();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n = static_cast<int>(plan->bwd_steps.size());
    if (cap < n) return fail(MMD_EINVAL, "profile buffer too small (%d < %d)", cap, n);
    if (reps < 1) reps = 1;
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) MMD_CUDA_OK(cudaEventCreate(&e));
    for (int i = 0; i < n; ++i) ms[i] = 0.f;
    int r = MMD_OK;
    for (int rep = 0; rep < reps && r == MMD_OK; ++rep) {
        MMD_CUDA_OK(cudaEventRecord(ev[0], st));
        for (int i = 0; i < n; ++i) {
            r = plan->bwd_steps[i](st);
            if (r != MMD_OK) break;
            MMD_CUDA_OK(cudaEventRecord(ev[i + 1], st));
        }
        MMD_CUDA_OK(cudaStreamSynchronize(st));
        for (int i = 0; i < n && r == MMD_OK; ++i) {
            float t = 0.f;
            MMD_CUDA_OK(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
            ms[i] += t / reps;
        }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    return r == MMD_OK ? n : r;
}
const char* mmd_model_backward_step_kind(const MmdModel* m, int batch, int index) {
    if (!m) return "";
    auto it = m->train_plans.find(batch);
    if (it == m->train_plans.end() || index < 0 || index >= static_cast<int>(it->second->bwd_kind.size())) return "";
    return it->second->bwd_kind[index].c_str();
}

}  // extern "C"
