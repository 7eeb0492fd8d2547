"""Operator parity on the GPU: each sm_100a kernel vs a plain PyTorch fp32 statement of the
reference op it replaces (reference lines cited in mm_diffusion_b200/ops.py).  fp16 storage,
fp32 accumulate -> tolerance: rel-L2 <= 2e-3 for convs / norms, 3e-3 for attention."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a = a.float()
    b = b.float()
    return (a - b).norm().item() / max(b.norm().item(), 1e-12)


@pytest.fixture(scope="module")
def ops():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from mm_diffusion_b200 import ops as o
    return o


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


@pytest.mark.parametrize("m,cins,cout", [(256, [128], 128), (1000, [64], 64), (4096, [256, 128], 256),
                                          (300, [128, 64, 64], 384), (16384, [512], 1536),
                                          (65536, [128], 512), (40000, [192, 64], 256),
                                          (90000, [64], 128), (80000, [128, 64], 128)])   # last two: > 4 waves of tiles, even / odd tile count
def test_conv_pointwise(ops, m, cins, cout):
    xs = [_rand(m, c, seed=i).half() for i, c in enumerate(cins)]
    w = _rand(cout, sum(cins), scale=0.05, seed=7)
    b = _rand(cout, seed=8)
    y = ops.conv_pointwise(xs, w, b)
    ref = torch.cat([x.float() for x in xs], dim=1) @ w.t() + b
    assert rel_l2(y, ref) < 2e-3


def _ref_gn(x, gamma, beta, ns, film=None, ns_per_batch=1, silu=False):
    """GroupNorm32 (+FiLM, +SiLU) of nn.py:16-33 / multimodal_unet.py:459-470 on [ns * rows, C] fp32."""
    c = x.shape[-1]
    xd = x.float().reshape(ns, -1, c).permute(0, 2, 1)
    y = F.group_norm(xd, 32, gamma, beta, eps=1e-5)
    if film is not None:
        fb = film.float().repeat_interleave(ns_per_batch, dim=0)
        y = y * (1 + fb[:, :c, None]) + fb[:, c:2 * c, None]
    if silu:
        y = F.silu(y)
    return y.permute(0, 2, 1).reshape(x.shape)


@pytest.mark.parametrize("m,c,cskip,cout,ns,film,silu", [
    (4096, 128, 0, 128, 4, True, True),        # out_layers, one domain per tile, BN=128
    (2048, 64, 64, 64, 2, True, True),         # small-config widths (BN=64), skip segment untouched
    (65536, 256, 256, 256, 4, True, True),     # BN=256, residual / skip K-segment after the normalised one
    (16384, 384, 0, 1152, 64, False, False),   # attention norm -> qkv, per-frame domains of 256 rows, 9 N tiles
    (4096, 512, 0, 1536, 64, False, False),    # 64-row domains: two domains per tile
    (40960, 128, 128, 128, 5, True, True),     # many tiles per CTA (table rebuilt when the domain changes)
    (32768, 512, 1024, 512, 2, True, True),    # widest source, long skip segment
])
def test_conv_pointwise_fused_gn_apply(ops, m, c, cskip, cout, ns, film, silu):
    """GroupNorm apply (+FiLM, +SiLU) folded into the GEMM's A path == norm then conv in fp32."""
    x = (_rand(m, c, seed=31) * 1.5 + 0.3).half()
    srcs = [x]
    if cskip:
        srcs.append(_rand(m, cskip, seed=32).half())
    gamma = 1 + 0.2 * _rand(c, seed=33)
    beta = 0.2 * _rand(c, seed=34)
    fb = 0.3 * _rand(ns, 2 * c + 6, seed=35) if film else None   # film_ld > 2C like the stacked emb table
    w = _rand(cout, c + cskip, scale=0.05, seed=36)
    b = _rand(cout, seed=37)
    y = ops.conv_pointwise_gn(srcs, w, b, gamma, beta, ns, film=fb, silu=silu)
    h = _ref_gn(x, gamma, beta, ns, film=fb, silu=silu)
    ref = torch.cat([h] + [s_.float() for s_ in srcs[1:]], dim=1) @ w.t() + b
    assert rel_l2(y, ref) < 2e-3


@pytest.mark.parametrize("b,l,c,cout,silu", [(2, 1600, 128, 128, True), (3, 400, 512, 1536, False), (4, 400, 256, 256, True)])
def test_conv_audio_pointwise_fused_gn_apply(ops, b, l, c, cout, silu):
    """audio geometry (L,B): one domain per sample, ragged last tile (L % 128 != 0), fused output statistics on top"""
    x = (_rand(b, l, c, seed=41) + 0.5).half()
    gamma = 1 + 0.2 * _rand(c, seed=42)
    beta = 0.2 * _rand(c, seed=43)
    fb = 0.3 * _rand(b, 2 * c, seed=44) if silu else None
    w = _rand(cout, c, scale=0.05, seed=45)
    bias = _rand(cout, seed=46)
    sums = torch.zeros(b, 32, 2, dtype=torch.float64, device="cuda")
    y = ops.conv_pointwise_gn([x], w, bias, gamma, beta, b, film=fb, silu=silu, gn_sums=sums, gn_rows=l)
    h = _ref_gn(x.reshape(b * l, c), gamma, beta, b, film=fb, silu=silu)
    ref = (h @ w.t() + bias).reshape(b, l, cout)
    assert rel_l2(y, ref) < 2e-3
    _check_sums(sums, y, b)


def _ref_gn_sums(y, domains):
    """(sum, sum of squares) per (domain, group of C/32 channels) of the fp16 result the kernel stored."""
    c = y.shape[-1]
    yd = y.double().reshape(domains, -1, 32, c // 32)
    return torch.stack([yd.sum(dim=(1, 3)), (yd * yd).sum(dim=(1, 3))], dim=-1)


def _check_sums(got, y, domains):
    ref = _ref_gn_sums(y, domains)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, err


@pytest.mark.parametrize("m,cin,cout,rows", [(1024, 128, 128, 128), (1024, 128, 128, 64), (4096, 64, 256, 1024),
                                             (8192, 128, 384, 4096), (65536, 128, 512, 4096), (40960, 256, 256, 64),
                                             (81920, 128, 128, 4096), (77824, 64, 128, 64)])
def test_conv_pointwise_fused_gn_statistics(ops, m, cin, cout, rows):
    """GEMM-epilogue GroupNorm statistics (128- and 256-wide tiles, 1 or 2 domains per tile) == sums over the output."""
    x = _rand(m, cin, seed=21).half()
    w = _rand(cout, cin, scale=0.05, seed=22)
    b = _rand(cout, seed=23)
    sums = torch.zeros(m // rows, 32, 2, dtype=torch.float64, device="cuda")
    y = ops.conv_pointwise([x], w, b, gn_sums=sums, gn_rows=rows)
    assert rel_l2(y, x.float() @ w.t() + b) < 2e-3
    _check_sums(sums, y, m // rows)


@pytest.mark.parametrize("b,f,p,c", [(2, 16, 256, 128), (1, 16, 64, 128), (1, 8, 1024, 256)])
def test_conv_temporal_fused_gn_statistics(ops, b, f, p, c):
    x = _rand(b, f, p, c, seed=4).half()
    wt = _rand(c, c, 3, scale=0.05, seed=5)
    bias = _rand(c, seed=6)
    sums = torch.zeros(b * f, 32, 2, dtype=torch.float64, device="cuda")
    y = ops.conv_temporal(x, wt, bias, gn_sums=sums)
    _check_sums(sums, y, b * f)


@pytest.mark.parametrize("b,l,ci,co,dil", [(2, 1600, 128, 128, 1), (3, 400, 128, 256, 4), (1, 25600, 128, 128, 2)])
def test_conv_audio_fused_gn_statistics(ops, b, l, ci, co, dil):
    """ragged last tile (L % 128 != 0): padding rows must not enter the statistics"""
    x = _rand(b, l, ci, seed=9).half()
    wt = _rand(co, ci, 3, scale=0.05, seed=10)
    bias = _rand(co, seed=11)
    sums = torch.zeros(b, 32, 2, dtype=torch.float64, device="cuda")
    y = ops.conv_audio(x, wt, bias, dil, gn_sums=sums)
    _check_sums(sums, y, b)


@pytest.mark.parametrize("n,h,w,ci,co", [(2, 64, 64, 128, 128), (3, 32, 32, 256, 128), (5, 16, 16, 64, 64),
                                          (16, 8, 8, 128, 256), (8, 4, 4, 64, 128), (20, 64, 64, 64, 128)])
def test_conv_spatial(ops, n, h, w, ci, co):
    x = _rand(n, h, w, ci, seed=1).half()
    wt = _rand(co, ci, 3, 3, scale=0.03, seed=2)
    b = _rand(co, seed=3)
    y = ops.conv_spatial(x, wt, b)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt, b, padding=1).permute(0, 2, 3, 1)
    assert rel_l2(y, ref) < 2e-3


@pytest.mark.parametrize("b,f,p,c", [(2, 16, 256, 128), (1, 16, 64, 128), (3, 8, 16, 64), (1, 16, 4096, 128)])
def test_conv_temporal(ops, b, f, p, c):
    x = _rand(b, f, p, c, seed=4).half()
    wt = _rand(c, c, 3, scale=0.05, seed=5)
    bias = _rand(c, seed=6)
    y = ops.conv_temporal(x, wt, bias)
    xr = x.float().permute(0, 2, 3, 1).reshape(b * p, c, f)
    ref = F.conv1d(xr, wt, bias, padding=1).reshape(b, p, c, f).permute(0, 3, 1, 2)
    assert rel_l2(y, ref) < 2e-3


@pytest.mark.parametrize("b,l,ci,co,dil", [(2, 1600, 128, 128, 1), (1, 400, 128, 256, 512), (3, 400, 64, 64, 4),
                                            (1, 25600, 128, 128, 2), (2, 100, 128, 128, 64)])
def test_conv_audio(ops, b, l, ci, co, dil):
    x = _rand(b, l, ci, seed=9).half()
    wt = _rand(co, ci, 3, scale=0.05, seed=10)
    bias = _rand(co, seed=11)
    y = ops.conv_audio(x, wt, bias, dil)
    ref = F.conv1d(x.float().permute(0, 2, 1), wt, bias, padding=dil, dilation=dil).permute(0, 2, 1)
    assert rel_l2(y, ref) < 2e-3


def test_conv_heads(ops):
    x = _rand(2, 8, 16, 16, 128, seed=12).half()
    wt = _rand(3, 128, 3, 3, 3, scale=0.05, seed=13)
    bias = _rand(3, seed=14)
    y = ops.conv3d_head(x, wt, bias)
    ref = F.conv3d(x.float().permute(0, 4, 1, 2, 3), wt, bias, padding=1).permute(0, 2, 1, 3, 4)
    assert rel_l2(y, ref) < 2e-3
    xa = _rand(2, 1000, 128, seed=15).half()
    wa = _rand(1, 128, 3, scale=0.05, seed=16)
    ba = _rand(1, seed=17)
    ya = ops.conv1d_head(xa, wa, ba)
    refa = F.conv1d(xa.float().permute(0, 2, 1), wa, ba, padding=1)
    assert rel_l2(ya, refa) < 2e-3


@pytest.mark.parametrize("ns,rows,c1,c2,silu,film", [(2, 4096, 128, 0, True, False), (3, 400, 384, 0, False, False),
                                                     (2, 1024, 512, 384, True, True), (32, 64, 256, 0, False, False),
                                                     (2, 100, 64, 64, True, True)])
def test_group_norm(ops, ns, rows, c1, c2, silu, film):
    x1 = (_rand(ns * rows, c1, seed=20) * 2 + 0.5).half()
    x2 = (_rand(ns * rows, c2, seed=21) - 0.3).half() if c2 else None
    C = c1 + c2
    gamma = _rand(C, seed=22) * 0.5 + 1
    beta = _rand(C, seed=23)
    fm = _rand(ns, 2 * C, seed=24) * 0.5 if film else None
    y = ops.group_norm(x1, gamma, beta, ns, x2=x2, film=fm, ns_per_batch=1, silu=silu)
    xx = x1.float() if x2 is None else torch.cat([x1.float(), x2.float()], dim=1)
    xr = xx.reshape(ns, rows, C).permute(0, 2, 1)
    ref = F.group_norm(xr, 32, gamma, beta, eps=1e-5)
    if film:
        ref = ref * (1 + fm[:, :C, None]) + fm[:, C:, None]
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1).reshape(ns * rows, C)
    assert rel_l2(y, ref) < 2e-3


def test_group_norm_temporal(ops):
    B, Fr, P, C = 2, 16, 64, 256
    x = (_rand(B, Fr, P, C, seed=30) * 1.5 + 0.2).half()
    gamma = _rand(C, seed=31) * 0.5 + 1
    beta = _rand(C, seed=32)
    y = ops.group_norm_temporal(x, gamma, beta)
    xr = x.float().permute(0, 2, 3, 1).reshape(B * P, C, Fr)
    ref = F.group_norm(xr, 32, gamma, beta, eps=1e-5).reshape(B, P, C, Fr).permute(0, 3, 1, 2)
    assert rel_l2(y, ref) < 2e-3


def test_resample(ops):
    x = _rand(3, 16, 16, 128, seed=40).half()
    y = ops.resample(x, "vpool")
    ref = F.avg_pool2d(x.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert rel_l2(y, ref) < 1e-3
    y = ops.resample(x, "vup")
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert rel_l2(y, ref) == 0
    a = _rand(2, 400, 64, seed=41).half()
    y = ops.resample(a, "apool")
    ref = F.avg_pool1d(a.float().permute(0, 2, 1), 4).permute(0, 2, 1)
    assert rel_l2(y, ref) < 1e-3
    y = ops.resample(a, "aup")
    ref = F.interpolate(a.float().permute(0, 2, 1), scale_factor=4, mode="nearest").permute(0, 2, 1)
    assert rel_l2(y, ref) == 0


def _ref_attention(q, k, v, B, heads, d, n_blocks, q_blk, k_blk, win, shift):
    """q [B*n_blocks*q_blk, heads*d], k/v [B*n_blocks*k_blk, heads*d] fp32 -> out like q."""
    out = torch.empty_like(q)
    qb = q.reshape(B, n_blocks, q_blk, heads, d)
    kb = k.reshape(B, n_blocks * k_blk, heads, d)
    vb = v.reshape(B, n_blocks * k_blk, heads, d)
    ob = out.reshape(B, n_blocks, q_blk, heads, d)
    tot = n_blocks * k_blk
    for i in range(n_blocks):
        idx = (torch.arange(win * k_blk, device=q.device) + (i + shift) * k_blk) % tot
        kk = kb[:, idx]  # [B, nk, h, d]
        vv = vb[:, idx]
        s = torch.einsum("bqhd,bkhd->bhqk", qb[:, i], kk) / math.sqrt(d)
        p = torch.softmax(s, dim=-1)
        ob[:, i] = torch.einsum("bhqk,bkhd->bqhd", p, vv)
    return out


@pytest.mark.parametrize("B,heads,d,n_blocks,q_blk,k_blk,win,shift", [
    (1, 4, 64, 2, 256, 256, 1, 0),     # spatial self, 16x16
    (2, 4, 64, 16, 1024, 400, 1, 5),   # cross video->audio ds2
    (1, 4, 64, 16, 400, 1024, 1, 15),  # cross audio->video ds2
    (1, 6, 64, 16, 100, 256, 4, 11),   # cross a->v ds4, wrapping window
    (2, 8, 64, 16, 25, 64, 8, 8),      # cross a->v ds8
    (1, 8, 64, 16, 64, 25, 16, 0),     # middle block, full window
    (1, 4, 96, 4, 256, 256, 1, 0),     # spatial self d=96
    (1, 4, 128, 16, 64, 64, 1, 0),     # spatial self d=128
    (2, 4, 128, 1, 400, 400, 1, 0),    # audio self
])
def test_attention(ops, B, heads, d, n_blocks, q_blk, k_blk, win, shift):
    C = heads * d
    nq = B * n_blocks * q_blk
    nk = B * n_blocks * k_blk
    qm = _rand(nq, 3 * C, seed=50).half()
    km = qm if q_blk == k_blk and win == 1 and shift == 0 else _rand(nk, 3 * C, seed=51).half()
    y = ops.attention(qm, km, km, 0, C, 2 * C, B, heads, d, n_blocks, q_blk, k_blk, win, shift)
    ref = _ref_attention(qm[:, :C].float(), km[:, C:2 * C].float(), km[:, 2 * C:].float(), B, heads, d, n_blocks,
                         q_blk, k_blk, win, shift)
    assert rel_l2(y, ref) < 3e-3


@pytest.mark.parametrize("B,Fr,P,C,heads", [(2, 16, 64, 512, 4), (1, 16, 256, 384, 4), (1, 8, 16, 64, 1)])
def test_temporal_attention(ops, B, Fr, P, C, heads):
    d = C // heads
    qkv = _rand(B, Fr, P, 3 * C, seed=60).half()
    y = ops.temporal_attention(qkv, heads)
    q, k, v = qkv.float().split(C, dim=-1)
    q = q.reshape(B, Fr, P, heads, d)
    k = k.reshape(B, Fr, P, heads, d)
    v = v.reshape(B, Fr, P, heads, d)
    s = torch.einsum("bfphd,bgphd->bphfg", q, k) / math.sqrt(d)
    p = torch.softmax(s, dim=-1)
    ref = torch.einsum("bphfg,bgphd->bfphd", p, v).reshape(B, Fr, P, C)
    assert rel_l2(y, ref) < 2e-3
