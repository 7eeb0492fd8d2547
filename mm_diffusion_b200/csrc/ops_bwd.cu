// Host launchers + operator-level C ABI of the training-backward kernels (wgrad.cuh, attention_bwd.cuh, backward.cuh).
#include "host.cuh"

namespace mmd {

// ------------------------------------------------------------- conv wgrad
int build_wgrad(const WgradProblem& pr, WgradParams* out, int* n_items) {
    WgradParams& p = *out;
    memset(&p, 0, sizeof(p));
    const ConvGeom& g = pr.g;
    if (g.rank < 2 || g.rank > 5) return fail(MMD_EINVAL, "wgrad geometry rank %d", g.rank);
    if (static_cast<long long>(g.box[0]) * g.box[1] * g.box[2] * g.box[3] != GEMM_BM)
        return fail(MMD_EINVAL, "wgrad box product must be %d", GEMM_BM);
    if (pr.n_src < 1 || pr.n_src > GEMM_MAX_SRC || pr.n_taps < 1 || pr.n_taps > GEMM_MAX_TAPS)
        return fail(MMD_EINVAL, "wgrad sources/taps out of range");
    if (pr.n % 8 != 0) return fail(MMD_EINVAL, "wgrad output channels %d not a multiple of 8", pr.n);
    p.n_src = pr.n_src;
    p.rank = g.rank;
    p.n_taps = pr.n_taps;
    long long m_tiles = 1;
    for (int i = 0; i < 4; ++i) {
        p.box[i] = g.box[i];
        p.ntile[i] = static_cast<int>((g.dims[i] + g.box[i] - 1) / g.box[i]);
        m_tiles *= p.ntile[i];
    }
    for (int t = 0; t < pr.n_taps; ++t)
        for (int j = 0; j < 3; ++j) p.tap[t][j] = pr.taps[t][j];
    uint64_t dims[5], str[4];
    uint32_t box[5];
    int total_chunks = 0, blocks = 0;
    for (int s = 0; s < pr.n_src; ++s) {
        if (pr.src_c[s] % GEMM_BK != 0) return fail(MMD_EINVAL, "wgrad source channels %d not a multiple of 64", pr.src_c[s]);
        p.src_chunks[s] = pr.src_c[s] / GEMM_BK;
        total_chunks += p.src_chunks[s];
        blocks += (p.src_chunks[s] + 1) / 2;
        dims[0] = pr.src_c[s];
        box[0] = GEMM_BK;
        uint64_t pitch = static_cast<uint64_t>(pr.src_c[s]) * sizeof(act_t);
        for (int i = 1; i < g.rank; ++i) {
            dims[i] = g.dims[i - 1];
            box[i] = g.box[i - 1];
            str[i - 1] = pitch;
            pitch *= g.dims[i - 1];
        }
        MMD_TRY(encode_tmap(&p.a_map[s], pr.src[s], g.rank, dims, str, box));
    }
    {
        dims[0] = pr.n;
        box[0] = GEMM_BK;
        uint64_t pitch = static_cast<uint64_t>(pr.n) * sizeof(act_t);
        for (int i = 1; i < g.rank; ++i) {
            dims[i] = g.dims[i - 1];
            box[i] = g.box[i - 1];
            str[i - 1] = pitch;
            pitch *= g.dims[i - 1];
        }
        MMD_TRY(encode_tmap(&p.dy_map, pr.dy, g.rank, dims, str, box));
    }
    p.m_tiles = static_cast<int>(m_tiles);
    p.n_tiles = (pr.n + 127) / 128;
    p.n = pr.n;
    p.blocks_per_tap = blocks;
    p.total_chunks = total_chunks;
    if (pr.ld != static_cast<long long>(total_chunks) * GEMM_BK * pr.n_taps)
        return fail(MMD_EINVAL, "wgrad: packed leading dimension %lld != %lld", pr.ld, static_cast<long long>(total_chunks) * GEMM_BK * pr.n_taps);
    const long long base_items = static_cast<long long>(p.n_tiles) * blocks * pr.n_taps;
    long long splits = (2LL * num_sms() + base_items - 1) / base_items;
    splits = std::max<long long>(1, std::min<long long>(splits, m_tiles));
    // keep at least ~4 token tiles per item so the TMEM drain / atomics stay a small part of an item
    splits = std::max<long long>(1, std::min<long long>(splits, std::max<long long>(1, m_tiles / 4)));
    p.splits = static_cast<int>(splits);
    p.dw = pr.dw;
    p.ld = pr.ld;
    p.db = pr.db;
    *n_items = static_cast<int>(base_items * splits);
    return MMD_OK;
}

int launch_wgrad(const WgradParams& p, int n_items, cudaStream_t st) {
    static bool done = false;
    if (!done) {
        MMD_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
        done = true;
    }
    const int grid = std::min(n_items, num_sms());
    conv_wgrad_kernel<<<grid, WG_THREADS, WG_SMEM, st>>>(p, n_items);
    MMD_CUDA_OK(cudaGetLastError());
    pdl_break(st);
    return MMD_OK;
}

int launch_unpack_wgrad(const float* dwpk, float* g, int co, int ci, int t, long long ld, long long col_off, float scale,
                        cudaStream_t st, const float* gscale) {
    const long long total = static_cast<long long>(co) * ci * t;
    unpack_wgrad_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(dwpk, g, co, ci, t, ld, col_off, scale, gscale);
    MMD_CUDA_OK(cudaGetLastError());
    pdl_break(st);
    return MMD_OK;
}

int launch_pack_weight_t(const float* w, act_t* dst, int co, int ci, int t, int c_lo, int cs, long long ld, long long col_off,
                         cudaStream_t st) {
    const long long total = static_cast<long long>(cs) * t * co;
    pack_weight_t_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(w, dst, co, ci, t, c_lo, cs, ld, col_off);
    MMD_CUDA_OK(cudaGetLastError());
    pdl_break(st);
    return MMD_OK;
}

int launch_grad_add(const act_t* x, act_t* y, long long n, int accumulate, cudaStream_t st) {
    if (n % 8 != 0) return fail(MMD_EINVAL, "grad_add: element count %lld not a multiple of 8", n);
    const long long n8 = n / 8;
    const unsigned grid = static_cast<unsigned>(std::max<long long>(1, std::min<long long>((n8 + 255) / 256, 16LL * num_sms())));
    grad_add_kernel<<<grid, 256, 0, st>>>(x, y, n8, accumulate);
    MMD_CUDA_OK(cudaGetLastError());
    pdl_break(st);
    return MMD_OK;
}

int launch_grad_add2d(const act_t* x, long long ldx, act_t* y, long long ldy, long long rows, int C, int accumulate, cudaStream_t st) {
    if (C % 8 != 0 || ldx % 8 != 0 || ldy % 8 != 0) return fail(MMD_EINVAL, "grad_add2d: widths must be multiples of 8");
    const long long total = rows * (C / 8);
    const unsigned grid = static_cast<unsigned>(std::max<long long>(1, std::min<long long>((total + 255) / 256, 16LL * num_sms())));
    grad_add2d_kernel<<<grid, 256, 0, st>>>(x, ldx, y, ldy, rows, C, accumulate);
    MMD_CUDA_OK(cudaGetLastError());
    pdl_break(st);
    return MMD_OK;
}

// ------------------------------------------------------------- attention backward
int launch_attn_lse(const AttnParams& p, int d, float* lse, long long lse_ld, cudaStream_t st) {
    AttnParams q = p;
    q.lse = lse;
    q.lse_ld = lse_ld;
    return launch_attn(q, d, st);
}

int launch_attn_delta(const act_t* d_out, const act_t* out, long long rows, int C, int heads, float* delta, long long delta_ld,
                      cudaStream_t st) {
    const long long warps = rows * heads;
    const unsigned grid = static_cast<unsigned>((warps * 32 + 255) / 256);
    attn_delta_kernel<<<grid, 256, 0, st>>>(d_out, out, rows, C, heads, delta, delta_ld);
    MMD_CUDA_OK(cudaGetLastError());
    pdl_break(st);
    return MMD_OK;
}

static int attn_bwd_map(CUtensorMap* m, const act_t* base, int ld, long long rows) {
    uint64_t dims[2] = {static_cast<uint64_t>(ld), static_cast<uint64_t>(rows)};
    uint64_t str[1] = {static_cast<uint64_t>(ld) * sizeof(act_t)};
    uint32_t box[2] = {64, 128};
    return encode_tmap(m, base, 2, dims, str, box);
}

int build_attn_bwd(const AttnProblem& f, const act_t* d_out, int d_out_ld, const float* lse, const float* delta,
                   long long stat_ld, const AttnBwdOut& o, AttnBwdParams* pq, AttnBwdParams* pkv) {
    if (f.d != 64 && f.d != 96 && f.d != 128) return fail(MMD_EINVAL, "attention head_dim %d unsupported (64/96/128)", f.d);
    if (f.q_blk <= 0 || f.k_blk <= 0 || f.win < 1 || f.win > f.n_blocks) return fail(MMD_EINVAL, "attention block geometry");
    const float rs = 1.0f / sqrtf(static_cast<float>(f.d));
    for (int pass = 0; pass < 2; ++pass) {
        AttnBwdParams& p = pass == 0 ? *pq : *pkv;
        memset(&p, 0, sizeof(p));
        p.lse = lse; p.delta = delta; p.stat_ld = stat_ld; p.q_rows_total = f.q_rows;
        p.B = f.B; p.heads = f.heads; p.n_blocks = f.n_blocks; p.win = f.win; p.shift_ptr = f.shift_dev;
        p.scale_log2 = 1.4426950408889634f * rs;
        p.rs = rs;
        if (pass == 0) {   // X = queries: X1 = Q, X2 = dO; Y1 = K, Y2 = V
            MMD_TRY(attn_bwd_map(&p.x1_map, f.q, f.q_ld, f.q_rows));
            MMD_TRY(attn_bwd_map(&p.x2_map, d_out, d_out_ld, f.q_rows));
            MMD_TRY(attn_bwd_map(&p.y1_map, f.k, f.k_ld, f.k_rows));
            MMD_TRY(attn_bwd_map(&p.y2_map, f.v, f.v_ld, f.k_rows));
            p.x1_col0 = f.q_col0; p.x2_col0 = 0; p.y1_col0 = f.k_col0; p.y2_col0 = f.v_col0;
            p.out1 = o.dq; p.out1_ld = o.dq_ld; p.out1_col0 = o.dq_col0;
            p.x_blk = f.q_blk; p.x_per_batch = f.q_blk * f.n_blocks;
            p.y_blk = f.k_blk; p.y_per_batch = f.k_blk * f.n_blocks;
        } else {           // X = keys: X1 = K, X2 = V; Y1 = Q, Y2 = dO
            MMD_TRY(attn_bwd_map(&p.x1_map, f.k, f.k_ld, f.k_rows));
            MMD_TRY(attn_bwd_map(&p.x2_map, f.v, f.v_ld, f.k_rows));
            MMD_TRY(attn_bwd_map(&p.y1_map, f.q, f.q_ld, f.q_rows));
            MMD_TRY(attn_bwd_map(&p.y2_map, d_out, d_out_ld, f.q_rows));
            p.x1_col0 = f.k_col0; p.x2_col0 = f.v_col0; p.y1_col0 = f.q_col0; p.y2_col0 = 0;
            p.out1 = o.dk; p.out1_ld = o.dk_ld; p.out1_col0 = o.dk_col0;
            p.out2 = o.dv; p.out2_ld = o.dv_ld; p.out2_col0 = o.dv_col0;
            p.x_blk = f.k_blk; p.x_per_batch = f.k_blk * f.n_blocks;
            p.y_blk = f.q_blk; p.y_per_batch = f.q_blk * f.n_blocks;
        }
        p.x_tiles = (p.x_blk + 127) / 128;
    }
    return MMD_OK;
}

template <int D>
static int attn_bwd_launch_d(const AttnBwdParams& pq, const AttnBwdParams& pkv, cudaStream_t st) {
    static bool done = false;
    if (!done) {
        MMD_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnBwdSmem<D>::TOTAL));
        MMD_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnBwdSmem<D>::TOTAL));
        done = true;
    }
    const int gq = pq.B * pq.n_blocks * pq.heads * pq.x_tiles;
    const int gk = pkv.B * pkv.n_blocks * pkv.heads * pkv.x_tiles;
    attn_bwd_kernel<D, false><<<gq, AB_THREADS, AttnBwdSmem<D>::TOTAL, st>>>(pq);
    MMD_CUDA_OK(cudaGetLastError());
    attn_bwd_kernel<D, true><<<gk, AB_THREADS, AttnBwdSmem<D>::TOTAL, st>>>(pkv);
    MMD_CUDA_OK(cudaGetLastError());
    pdl_break(st);
    return MMD_OK;
}

int launch_attn_bwd(const AttnBwdParams& pq, const AttnBwdParams& pkv, int d, cudaStream_t st) {
    if (d == 64) return attn_bwd_launch_d<64>(pq, pkv, st);
    if (d == 96) return attn_bwd_launch_d<96>(pq, pkv, st);
    if (d == 128) return attn_bwd_launch_d<128>(pq, pkv, st);
    return fail(MMD_EINVAL, "attention backward head_dim %d", d);
}

// ------------------------------------------------------------- GroupNorm backward
static int gn_bwd_rows_per_block(int ns, int rows, int C) {
    const int rows_per_pass = std::max(1, 256 / (C / 8));
    const int target_blocks = 4 * num_sms();
    const int per_domain = std::max(1, target_blocks / std::max(1, ns));
    int rpb = (rows + per_domain - 1) / per_domain;
    rpb = std::max(rpb, 2 * GN_UNROLL * rows_per_pass);
    return std::min(rpb, rows);
}

int launch_gn_bwd(const GnBwdProblem& pr, cudaStream_t st) {
    const int C = pr.s.c1 + pr.s.c2;
    if (C % 32 != 0 || pr.s.c1 % 8 != 0 || C / 8 > 256) return fail(MMD_EINVAL, "group norm backward channels %d unsupported", C);
    GnBwdArgs a{};
    a.s = pr.s;
    a.R = pr.rows;
    a.rows_per_block = gn_bwd_rows_per_block(pr.ns, pr.rows, C);
    a.sums = pr.sums;
    a.nsub = pr.nsub;
    a.stat_rows = pr.stat_rows > 0 ? pr.stat_rows : pr.rows;
    a.gamma = pr.gamma; a.beta = pr.beta; a.film = pr.film; a.film_ld = pr.film_ld;
    a.ns_per_batch = pr.ns_per_batch > 0 ? pr.ns_per_batch : 1;
    a.do_silu = pr.silu;
    a.dy = pr.dy;
    a.T = pr.T;
    a.drop = pr.drop;
    a.drop_site = pr.drop_site;
    dim3 grid((pr.rows + a.rows_per_block - 1) / a.rows_per_block, pr.ns);
    const size_t sm_a = (6 * C + 64) * sizeof(float), sm_b = (5 * C + 128) * sizeof(float);
    gn_bwd_reduce_kernel<<<grid, 256, sm_a, st>>>(a);
    MMD_CUDA_OK(cudaGetLastError());
    gn_bwd_apply_kernel<<<grid, 256, sm_b, st>>>(a, pr.out);
    MMD_CUDA_OK(cudaGetLastError());
    if (pr.dgamma) {
        gn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(pr.T, pr.ns, C, pr.gamma, pr.beta, pr.film, pr.film_ld, a.ns_per_batch,
                                                                pr.dgamma, pr.dbeta, pr.dfilm, pr.gscale);
        MMD_CUDA_OK(cudaGetLastError());
    }
    pdl_break(st);
    return MMD_OK;
}

template <int CPG>
static int gn_temporal_bwd_cpg(const act_t* x, const act_t* dy, act_t* dx, const float* gamma, float* dgamma, float* dbeta, int B,
                               int F, int P, int C, const float* gscale, cudaStream_t st) {
    const long long total = static_cast<long long>(B) * P * 32;
    const unsigned grid = static_cast<unsigned>((total + 127) / 128);
    const size_t sm = 2 * C * sizeof(float);
    if (F == 16) gn_temporal_bwd_kernel<CPG, 16><<<grid, 128, sm, st>>>(x, dy, dx, gamma, dgamma, dbeta, B, P, C, gscale);
    else if (F == 8) gn_temporal_bwd_kernel<CPG, 8><<<grid, 128, sm, st>>>(x, dy, dx, gamma, dgamma, dbeta, B, P, C, gscale);
    else return fail(MMD_EINVAL, "temporal group norm backward supports 8 or 16 frames, got %d", F);
    MMD_CUDA_OK(cudaGetLastError());
    pdl_break(st);
    return MMD_OK;
}

int launch_gn_temporal_bwd(const act_t* x, const act_t* dy, act_t* dx, const float* gamma, float* dgamma, float* dbeta, int B,
                           int F, int P, int C, const float* gscale, cudaStream_t st) {
    switch (C / 32) {
        case 2: return gn_temporal_bwd_cpg<2>(x, dy, dx, gamma, dgamma, dbeta, B, F, P, C, gscale, st);
        case 4: return gn_temporal_bwd_cpg<4>(x, dy, dx, gamma, dgamma, dbeta, B, F, P, C, gscale, st);
        case 8: return gn_temporal_bwd_cpg<8>(x, dy, dx, gamma, dgamma, dbeta, B, F, P, C, gscale, st);
        case 12: return gn_temporal_bwd_cpg<12>(x, dy, dx, gamma, dgamma, dbeta, B, F, P, C, gscale, st);
        case 16: return gn_temporal_bwd_cpg<16>(x, dy, dx, gamma, dgamma, dbeta, B, F, P, C, gscale, st);
        default: return fail(MMD_EINVAL, "temporal group norm backward channels %d unsupported", C);
    }
}

int launch_temporal_attn_bwd(const act_t* qkv, const act_t* d_out, act_t* dqkv, int B, int F, int P, int C, int heads,
                             cudaStream_t st) {
    const int d = C / heads;
    if (d % 2 != 0 || C % heads != 0) return fail(MMD_EINVAL, "temporal attention backward head dim %d", d);
    const size_t per_warp = static_cast<size_t>(4) * 16 * (d + 2) * sizeof(act_t) + 2 * 16 * 17 * sizeof(float);
    const size_t smem = per_warp * TAB_WARPS;
    if (smem > 200 * 1024) return fail(MMD_EINVAL, "temporal attention backward head dim %d too large", d);
    static bool attr_done = false;
    if (!attr_done) {
        MMD_CUDA_OK(cudaFuncSetAttribute(temporal_attn_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MMD_CUDA_OK(cudaFuncSetAttribute(temporal_attn_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    const long long items = static_cast<long long>(B) * P * heads;
    const unsigned grid = static_cast<unsigned>(std::min<long long>((items + TAB_WARPS - 1) / TAB_WARPS, 8LL * num_sms()));
    if (F == 16) temporal_attn_bwd_kernel<16><<<grid, TAB_WARPS * 32, smem, st>>>(qkv, d_out, dqkv, B, P, C, heads);
    else if (F == 8) temporal_attn_bwd_kernel<8><<<grid, TAB_WARPS * 32, smem, st>>>(qkv, d_out, dqkv, B, P, C, heads);
    else return fail(MMD_EINVAL, "temporal attention backward supports F in {8,16}, got %d", F);
    MMD_CUDA_OK(cudaGetLastError());
    pdl_break(st);
    return MMD_OK;
}

int launch_resample_bwd(const act_t* dy, act_t* dx, int mode, int n, int h, int w, int c, int accumulate, cudaStream_t st) {
    if (mode < 0 || mode > 3 || c % 8 != 0) return fail(MMD_EINVAL, "resample backward mode %d / channels %d", mode, c);
    const long long total = (mode == 0 || mode == 2) ? static_cast<long long>(n) * h * w * (c / 8) : static_cast<long long>(n) * h * (c / 8);
    resample_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(dy, dx, mode, n, h, w, c, accumulate);
    MMD_CUDA_OK(cudaGetLastError());
    pdl_break(st);
    return MMD_OK;
}

static long long head_tokens(const HeadGeom& g) {
    long long t = 1;
    for (int i = 0; i < 4; ++i) t *= g.dims[i];
    return t;
}

int launch_pack_head_t(const float* w, act_t* wt, int n_out, int C, int T, int ldG, cudaStream_t st) {
    MMD_CUDA_OK(cudaMemsetAsync(wt, 0, sizeof(act_t) * static_cast<size_t>(C) * ldG, st));
    pack_head_t_kernel<<<(n_out * C * T + 255) / 256, 256, 0, st>>>(w, wt, n_out, C, T, ldG);
    MMD_CUDA_OK(cudaGetLastError());
    pdl_break(st);
    return MMD_OK;
}

int build_head_bwd(const HeadGeom& hg, const act_t* x, act_t* G, const act_t* wt, const float* zero_bias, act_t* dx, float* dwpk,
                   HeadBwdPlan* out) {
    HeadBwdPlan& hp = *out;
    const int terms = hg.n_taps * hg.n_out;
    if (terms > 128 || hg.C % 64 != 0)
        return fail(MMD_EINVAL, "head backward: %d taps x %d outputs over %d channels unsupported", hg.n_taps, hg.n_out, hg.C);
    hp.hg = hg;
    hp.tokens = head_tokens(hg);
    hp.ldG = head_ld(terms);
    hp.G = G;
    hp.dwpk = dwpk;
    hp.want_dx = dx != nullptr;
    ConvGeom g2;
    g2.rank = 2;
    g2.dims[0] = hp.tokens;
    geom_fill_box(g2);
    if (dx) {
        GemmProblem pr;
        pr.g = g2;
        pr.n_src = 1; pr.src[0] = G; pr.src_c[0] = hp.ldG;
        pr.n_taps = 1;
        pr.w = wt; pr.bias = zero_bias; pr.n = hg.C; pr.bn = pick_bn(hg.C); pr.out = dx;
        MMD_TRY(build_gemm(pr, &hp.gemm));
        hp.bn = pr.bn;
    }
    WgradProblem wp;
    wp.g = g2;
    wp.n_src = 1; wp.src[0] = x; wp.src_c[0] = hg.C;
    wp.n_taps = 1;
    wp.dy = G; wp.n = hp.ldG; wp.dw = dwpk; wp.ld = hg.C;
    return build_wgrad(wp, &hp.wgrad, &hp.items);
}

int run_head_bwd(const HeadBwdPlan& hp, const float* dout, float* dw, float* db, const float* gscale, cudaStream_t st) {
    const HeadGeom& hg = hp.hg;
    const long long total = hp.tokens * (hp.ldG / 8);
    head_im2col_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(hg, dout, hp.G, hp.ldG, hp.tokens, gscale);
    if (db)
        head_bias_kernel<<<dim3(static_cast<unsigned>(std::min<long long>((hp.tokens + 255) / 256, 2LL * num_sms())), hg.n_out), 256, 0, st>>>(
            hg, dout, db, hp.tokens);
    MMD_CUDA_OK(cudaGetLastError());
    pdl_break(st);
    if (hp.want_dx) {
        MMD_TRY(launch_gemm(hp.gemm, hp.bn, st));
        pdl_break(st);
    }
    if (dw) {
        MMD_CUDA_OK(cudaMemsetAsync(hp.dwpk, 0, sizeof(float) * static_cast<size_t>(hp.ldG) * hg.C, st));
        MMD_TRY(launch_wgrad(hp.wgrad, hp.items, st));
        const int n = hg.n_out * hg.C * hg.n_taps;
        unpack_head_wgrad_kernel<<<(n + 255) / 256, 256, 0, st>>>(hp.dwpk, dw, hg.n_out, hg.C, hg.n_taps, gscale);
        MMD_CUDA_OK(cudaGetLastError());
    }
    return MMD_OK;
}

}  // namespace mmd

// ===========================================================================
//                 operator-level C ABI (parity tests of the backward kernels)
// ===========================================================================
using namespace mmd;

static void fill_geom(const MmdConvDesc* d, ConvGeom& g) {
    g.rank = d->rank;
    for (int i = 0; i < 4; ++i) { g.dims[i] = d->dims[i] > 0 ? d->dims[i] : 1; g.box[i] = d->box[i] > 0 ? d->box[i] : 1; }
    if (d->box[0] <= 0) geom_fill_box(g);
}

extern "C" {

int mmd_op_conv_wgrad(const MmdConvDesc* d, const void* dy, float* dweight, float* dbias, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!d || !dy || !dweight) return fail(MMD_EINVAL, "conv_wgrad: null argument");
    WgradProblem pr;
    fill_geom(d, pr.g);
    pr.n_src = d->n_src;
    int ctot = 0;
    for (int s = 0; s < d->n_src && s < GEMM_MAX_SRC; ++s) {
        pr.src[s] = static_cast<const act_t*>(d->src[s]);
        pr.src_c[s] = d->src_channels[s];
        ctot += d->src_channels[s];
    }
    pr.n_taps = d->n_taps;
    for (int t = 0; t < d->n_taps && t < GEMM_MAX_TAPS; ++t)
        for (int j = 0; j < 3; ++j) pr.taps[t][j] = d->taps[t][j];
    pr.dy = static_cast<const act_t*>(dy);
    pr.n = d->n;
    pr.ld = static_cast<long long>(ctot) * d->n_taps;
    float* dwpk = nullptr;
    const size_t dw_floats = static_cast<size_t>(pr.ld) * d->n;
    MMD_CUDA_OK(cudaMallocAsync(&dwpk, sizeof(float) * (dw_floats + d->n), st));
    MMD_CUDA_OK(cudaMemsetAsync(dwpk, 0, sizeof(float) * (dw_floats + d->n), st));
    pr.dw = dwpk;
    if (dbias) pr.db = dwpk + dw_floats;   // bias gradient reduced by the wgrad kernel itself (ones-operand MMA)
    WgradParams wp;
    int items = 0;
    int r = build_wgrad(pr, &wp, &items);
    if (r == MMD_OK) r = launch_wgrad(wp, items, st);
    // caller's dweight is [n][ctot][taps] (reference layout), accumulated into
    if (r == MMD_OK) r = launch_unpack_wgrad(dwpk, dweight, d->n, ctot, d->n_taps, pr.ld, 0, 1.0f, st);
    if (r == MMD_OK && dbias) {
        axpy_f32_kernel<<<(d->n + 255) / 256, 256, 0, st>>>(dwpk + dw_floats, dbias, d->n, nullptr);
        if (cudaGetLastError() != cudaSuccess) r = fail(MMD_ECUDA, "bias gradient accumulate");
    }
    cudaFreeAsync(dwpk, st);
    return r;
}

// Data gradient of source `src_index` through the forward kernel with transposed weights and negated taps.
int mmd_op_conv_dgrad(const MmdConvDesc* d, const void* dy, int src_index, void* dx, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!d || !dy || !dx || src_index < 0 || src_index >= d->n_src) return fail(MMD_EINVAL, "conv_dgrad: bad argument");
    if (d->n % 64 != 0) return fail(MMD_EINVAL, "conv_dgrad: output channels %d not a multiple of 64", d->n);
    GemmProblem pr;
    fill_geom(d, pr.g);
    int ctot = 0, c_lo = 0;
    for (int s = 0; s < d->n_src; ++s) {
        if (s == src_index) c_lo = ctot;
        ctot += d->src_channels[s];
    }
    const int cs = d->src_channels[src_index];
    pr.n_src = 1;
    pr.src[0] = static_cast<const act_t*>(dy);
    pr.src_c[0] = d->n;
    pr.n_taps = d->n_taps;
    for (int t = 0; t < d->n_taps; ++t)
        for (int j = 0; j < 3; ++j) pr.taps[t][j] = -d->taps[t][j];
    pr.n = cs;
    pr.bn = pick_bn(cs);
    pr.out = static_cast<act_t*>(dx);
    const long long kt = pr.k_total();
    const int npad = pr.n_pad();
    act_t* wp = nullptr;
    float* bp = nullptr;
    MMD_CUDA_OK(cudaMallocAsync(&wp, sizeof(act_t) * kt * npad, st));
    MMD_CUDA_OK(cudaMallocAsync(&bp, sizeof(float) * npad, st));
    MMD_CUDA_OK(cudaMemsetAsync(wp, 0, sizeof(act_t) * kt * npad, st));
    MMD_CUDA_OK(cudaMemsetAsync(bp, 0, sizeof(float) * npad, st));
    int r = launch_pack_weight_t(d->weight, wp, d->n, ctot, d->n_taps, c_lo, cs, kt, 0, st);
    pr.w = wp;
    pr.bias = bp;
    GemmParams gp;
    if (r == MMD_OK) r = build_gemm(pr, &gp);
    if (r == MMD_OK) r = launch_gemm(gp, pr.bn, st);
    cudaFreeAsync(wp, st);
    cudaFreeAsync(bp, st);
    return r;
}

int mmd_op_group_norm_bwd(const void* x1, int c1, const void* x2, int c2, int ns, int rows, const float* gamma, const float* beta,
                          const float* film, int film_ld, int ns_per_batch, int silu, const void* dy, void* dx1, void* dx2,
                          float* dgamma, float* dbeta, float* dfilm, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int C = c1 + c2;
    GnSrc s{static_cast<const act_t*>(x1), c1, c1, static_cast<const act_t*>(x2), c2, c2};
    double* sums = nullptr;
    float* T = nullptr;
    MMD_CUDA_OK(cudaMallocAsync(&sums, sizeof(double) * 64 * ns, st));
    MMD_CUDA_OK(cudaMallocAsync(&T, sizeof(float) * 2 * C * ns, st));
    MMD_CUDA_OK(cudaMemsetAsync(T, 0, sizeof(float) * 2 * C * ns, st));
    int r = launch_gn_stats(s, ns, rows, sums, st);
    GnBwdProblem pr;
    pr.s = s; pr.ns = ns; pr.rows = rows; pr.sums = sums; pr.nsub = 1; pr.stat_rows = rows;
    pr.gamma = gamma; pr.beta = beta; pr.film = film; pr.film_ld = film_ld; pr.ns_per_batch = ns_per_batch > 0 ? ns_per_batch : 1;
    pr.silu = silu; pr.dy = static_cast<const act_t*>(dy); pr.T = T;
    pr.out = GnBwdOut{static_cast<act_t*>(dx1), c1, 0, static_cast<act_t*>(dx2), c2, 0};
    pr.dgamma = dgamma; pr.dbeta = dbeta; pr.dfilm = dfilm; pr.gscale = nullptr;
    if (r == MMD_OK) r = launch_gn_bwd(pr, st);
    cudaFreeAsync(sums, st);
    cudaFreeAsync(T, st);
    return r;
}

int mmd_op_group_norm_temporal_bwd(const void* x, const void* dy, void* dx, const float* gamma, float* dgamma, float* dbeta, int B,
                                   int F, int P, int C, void* stream) {
    return launch_gn_temporal_bwd(static_cast<const act_t*>(x), static_cast<const act_t*>(dy), static_cast<act_t*>(dx), gamma, dgamma,
                                  dbeta, B, F, P, C, nullptr, static_cast<cudaStream_t>(stream));
}

int mmd_op_resample_bwd(const void* dy, void* dx, int mode, int n, int h, int w, int c, void* stream) {
    return launch_resample_bwd(static_cast<const act_t*>(dy), static_cast<act_t*>(dx), mode, n, h, w, c, 0,
                               static_cast<cudaStream_t>(stream));
}

int mmd_op_temporal_attention_bwd(const void* qkv, const void* d_out, void* dqkv, int B, int F, int P, int C, int heads, void* stream) {
    return launch_temporal_attn_bwd(static_cast<const act_t*>(qkv), static_cast<const act_t*>(d_out), static_cast<act_t*>(dqkv), B, F,
                                    P, C, heads, static_cast<cudaStream_t>(stream));
}

// Forward (writes d->out and the per-row log-sum-exp) then backward of the attention core.  lse: fp32 [heads][q_rows].
int mmd_op_attention_fwd_bwd(const MmdAttnDesc* d, const void* d_out, float* lse, void* dq, void* dk, void* dv, int grad_ld,
                             int dq_col0, int dk_col0, int dv_col0, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!d || !d_out || !lse || !dq || !dk || !dv) return fail(MMD_EINVAL, "attention_fwd_bwd: null argument");
    int* shift_dev = nullptr;
    float* delta = nullptr;
    MMD_CUDA_OK(cudaMallocAsync(&shift_dev, sizeof(int), st));
    MMD_CUDA_OK(cudaMemcpyAsync(shift_dev, &d->shift, sizeof(int), cudaMemcpyHostToDevice, st));
    MMD_CUDA_OK(cudaMallocAsync(&delta, sizeof(float) * d->heads * d->q_rows, st));
    AttnProblem pr{static_cast<const act_t*>(d->q), d->q_ld, d->q_col0, d->q_rows,
                   static_cast<const act_t*>(d->k), d->k_ld, d->k_col0, d->k_rows,
                   static_cast<const act_t*>(d->v), d->v_ld, d->v_col0,
                   static_cast<act_t*>(d->out), d->out_ld,
                   d->batch, d->heads, d->head_dim, d->n_blocks, d->q_blk, d->k_blk, d->win, shift_dev};
    AttnParams ap;
    int r = build_attn(pr, &ap);
    if (r == MMD_OK) r = launch_attn_lse(ap, d->head_dim, lse, d->q_rows, st);
    const int C = d->heads * d->head_dim;
    if (r == MMD_OK && d->out_ld != C) r = fail(MMD_EINVAL, "attention_fwd_bwd: out_ld must equal heads * head_dim");
    if (r == MMD_OK) r = launch_attn_delta(static_cast<const act_t*>(d_out), static_cast<const act_t*>(d->out), d->q_rows, C, d->heads,
                                           delta, d->q_rows, st);
    AttnBwdOut o{static_cast<act_t*>(dq), grad_ld, dq_col0, static_cast<act_t*>(dk), grad_ld, dk_col0, static_cast<act_t*>(dv), grad_ld, dv_col0};
    AttnBwdParams pq, pkv;
    if (r == MMD_OK) r = build_attn_bwd(pr, static_cast<const act_t*>(d_out), C, lse, delta, d->q_rows, o, &pq, &pkv);
    if (r == MMD_OK) r = launch_attn_bwd(pq, pkv, d->head_dim, st);
    cudaFreeAsync(shift_dev, st);
    cudaFreeAsync(delta, st);
    return r;
}

// Narrow-head adjoints: geometry from the forward descriptor (out_f32 layout strides); dx fp16 [tokens][C] (scale 1).
// Same tensor-core path as the model's backward plan (build_head_bwd / run_head_bwd).
int mmd_op_head_bwd(const MmdConvDesc* d, const float* dout, void* dx, float* dweight, float* dbias, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!d || !dout || d->n_src != 1 || !d->weight) return fail(MMD_EINVAL, "head_bwd: bad argument");
    HeadGeom g{};
    g.ncoord = d->rank - 1;
    for (int i = 0; i < 4; ++i) { g.dims[i] = d->dims[i] > 0 ? static_cast<int>(d->dims[i]) : 1; g.ostride[i] = d->ostride[i]; }
    g.ostride_c = d->ostride_c;
    g.n_out = d->n;
    g.n_taps = d->n_taps;
    for (int t = 0; t < d->n_taps; ++t)
        for (int j = 0; j < 3; ++j) g.tap[t][j] = d->taps[t][j];
    g.C = d->src_channels[0];
    const int ldG = head_ld(g.n_taps * g.n_out);
    long long tokens = 1;
    for (int i = 0; i < 4; ++i) tokens *= g.dims[i];
    act_t *G = nullptr, *wt = nullptr;
    float *dwpk = nullptr, *zb = nullptr;
    MMD_CUDA_OK(cudaMallocAsync(&G, sizeof(act_t) * static_cast<size_t>(tokens) * ldG, st));
    MMD_CUDA_OK(cudaMallocAsync(&wt, sizeof(act_t) * static_cast<size_t>(g.C) * ldG, st));
    MMD_CUDA_OK(cudaMallocAsync(&dwpk, sizeof(float) * static_cast<size_t>(ldG) * g.C, st));
    MMD_CUDA_OK(cudaMallocAsync(&zb, sizeof(float) * 1024, st));
    MMD_CUDA_OK(cudaMemsetAsync(zb, 0, sizeof(float) * 1024, st));
    int r = launch_pack_head_t(d->weight, wt, g.n_out, g.C, g.n_taps, ldG, st);
    HeadBwdPlan hp;
    if (r == MMD_OK) r = build_head_bwd(g, static_cast<const act_t*>(d->src[0]), G, wt, zb, static_cast<act_t*>(dx), dwpk, &hp);
    if (r == MMD_OK) r = run_head_bwd(hp, dout, dweight, dbias, nullptr, st);
    cudaFreeAsync(G, st);
    cudaFreeAsync(wt, st);
    cudaFreeAsync(dwpk, st);
    cudaFreeAsync(zb, st);
    return r;
}

}  // extern "C"
