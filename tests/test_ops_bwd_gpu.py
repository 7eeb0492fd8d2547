"""Backward-operator parity on the GPU: each sm_100a backward kernel vs torch.autograd (fp32) of the reference op it
differentiates (reference lines cited in mm_diffusion_b200/ops.py / include/mmdiff.h).  fp16 storage of activations
and activation gradients, fp32 accumulation -> tolerance rel-L2 <= 3e-3 (5e-3 for attention gradients)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 3e-3


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


def _report(name, errs, tol=TOL):
    msg = ", ".join(f"{k}={v:.2e}" for k, v in errs.items())
    print(f"[bwd] {name}: {msg}")
    bad = {k: v for k, v in errs.items() if not (v < tol)}
    assert not bad, f"{name}: {bad} (tolerance {tol})"


# ------------------------------------------------------------------ conv weight gradients (tcgen05, MN-major operands)
@pytest.mark.parametrize("m,cins,cout", [(256, [128], 128), (4096, [64], 64), (4096, [256, 128], 256),
                                          (1000, [128, 64, 64], 384), (65536, [128], 512), (40000, [192], 128)])
def test_conv_wgrad_pointwise(ops, m, cins, cout):
    xs = [_rand(m, c, seed=i).half() for i, c in enumerate(cins)]
    dy = _rand(m, cout, scale=0.5, seed=9).half()
    dw, db = ops.conv_wgrad(xs, dy, 2, [m], [(0, 0, 0)])
    ref_w = dy.float().t() @ torch.cat([x.float() for x in xs], dim=1)
    _report(f"wgrad_pointwise m={m} {cins}->{cout}", {"dw": rel_l2(dw[:, :, 0], ref_w), "db": rel_l2(db, dy.float().sum(0))})


@pytest.mark.parametrize("n,h,w,ci,co", [(2, 16, 16, 128, 128), (4, 32, 32, 64, 128), (1, 8, 8, 256, 384), (16, 64, 64, 128, 128)])
def test_conv_wgrad_spatial(ops, n, h, w, ci, co):
    x = _rand(n, h, w, ci, seed=1).half()
    dy = _rand(n, h, w, co, scale=0.5, seed=2).half()
    taps = [(kx - 1, ky - 1, 0) for ky in range(3) for kx in range(3)]
    dw, db = ops.conv_wgrad([x], dy, 4, [w, h, n], taps)
    wt = torch.zeros(co, ci, 3, 3, device="cuda", requires_grad=True)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, None, padding=1)
    (gw,) = torch.autograd.grad(y, wt, dy.float().permute(0, 3, 1, 2))
    _report(f"wgrad_spatial {n}x{h}x{w} {ci}->{co}", {"dw": rel_l2(dw, gw.reshape(co, ci, 9)), "db": rel_l2(db, dy.float().sum((0, 1, 2)))})


@pytest.mark.parametrize("b,f,p,c", [(2, 16, 256, 128), (1, 8, 64, 64), (1, 16, 1024, 256)])
def test_conv_wgrad_temporal(ops, b, f, p, c):
    x = _rand(b, f, p, c, seed=4).half()
    dy = _rand(b, f, p, c, scale=0.5, seed=5).half()
    dw, _ = ops.conv_wgrad([x], dy, 4, [p, f, b], [(0, k - 1, 0) for k in range(3)])
    wt = torch.zeros(c, c, 3, device="cuda", requires_grad=True)
    xr = x.float().permute(0, 2, 3, 1).reshape(b * p, c, f)
    y = F.conv1d(xr, wt, None, padding=1)
    (gw,) = torch.autograd.grad(y, wt, dy.float().permute(0, 2, 3, 1).reshape(b * p, c, f))
    _report(f"wgrad_temporal {b}x{f}x{p}x{c}", {"dw": rel_l2(dw, gw)})


@pytest.mark.parametrize("b,l,ci,co,dil", [(2, 1600, 128, 128, 1), (1, 400, 128, 256, 512), (3, 400, 64, 64, 4), (2, 100, 128, 128, 64)])
def test_conv_wgrad_audio(ops, b, l, ci, co, dil):
    x = _rand(b, l, ci, seed=9).half()
    dy = _rand(b, l, co, scale=0.5, seed=10).half()
    dw, _ = ops.conv_wgrad([x], dy, 3, [l, b], [((k - 1) * dil, 0, 0) for k in range(3)])
    wt = torch.zeros(co, ci, 3, device="cuda", requires_grad=True)
    y = F.conv1d(x.float().permute(0, 2, 1), wt, None, padding=dil, dilation=dil)
    (gw,) = torch.autograd.grad(y, wt, dy.float().permute(0, 2, 1))
    _report(f"wgrad_audio {b}x{l} {ci}->{co} dil {dil}", {"dw": rel_l2(dw, gw)})


# ------------------------------------------------------------------ conv data gradients (forward kernel, transposed weights)
def test_conv_dgrad(ops):
    errs = {}
    # spatial 3x3
    n, h, w, ci, co = 2, 16, 16, 128, 256
    wt = _rand(co, ci, 3, 3, scale=0.03, seed=2)
    dy = _rand(n, h, w, co, scale=0.5, seed=3).half()
    taps = [(kx - 1, ky - 1, 0) for ky in range(3) for kx in range(3)]
    dx = ops.conv_dgrad([ci], dy, wt.reshape(co, ci, 9), 4, [w, h, n], taps)
    x = torch.zeros(n, ci, h, w, device="cuda", requires_grad=True)
    (gx,) = torch.autograd.grad(F.conv2d(x, wt, None, padding=1), x, dy.float().permute(0, 3, 1, 2))
    errs["spatial"] = rel_l2(dx, gx.permute(0, 2, 3, 1))
    # temporal k3
    b, f, p, c = 2, 8, 64, 128
    wt = _rand(c, c, 3, scale=0.05, seed=5)
    dy = _rand(b, f, p, c, scale=0.5, seed=6).half()
    dx = ops.conv_dgrad([c], dy, wt, 4, [p, f, b], [(0, k - 1, 0) for k in range(3)])
    x = torch.zeros(b * p, c, f, device="cuda", requires_grad=True)
    (gx,) = torch.autograd.grad(F.conv1d(x, wt, None, padding=1), x, dy.float().permute(0, 2, 3, 1).reshape(b * p, c, f))
    errs["temporal"] = rel_l2(dx, gx.reshape(b, p, c, f).permute(0, 3, 1, 2))
    # dilated audio k3, ragged length
    b, l, ci, co, dil = 2, 400, 128, 64, 8
    wt = _rand(co, ci, 3, scale=0.05, seed=7)
    dy = _rand(b, l, co, scale=0.5, seed=8).half()
    dx = ops.conv_dgrad([ci], dy, wt, 3, [l, b], [((k - 1) * dil, 0, 0) for k in range(3)])
    x = torch.zeros(b, ci, l, device="cuda", requires_grad=True)
    (gx,) = torch.autograd.grad(F.conv1d(x, wt, None, padding=dil, dilation=dil), x, dy.float().permute(0, 2, 1))
    errs["audio"] = rel_l2(dx, gx.permute(0, 2, 1))
    # pointwise, second of two sources
    m, cins, co = 3000, [128, 64], 256
    wt = _rand(co, sum(cins), scale=0.05, seed=9)
    dy = _rand(m, co, scale=0.5, seed=10).half()
    dx = ops.conv_dgrad(cins, dy, wt.reshape(co, sum(cins), 1), 2, [m], [(0, 0, 0)], src_index=1)
    errs["pointwise_src1"] = rel_l2(dx, dy.float() @ wt[:, 128:])
    _report("conv_dgrad", errs)


# ------------------------------------------------------------------ GroupNorm (+FiLM, +SiLU)
@pytest.mark.parametrize("ns,rows,c1,c2,silu,film", [(2, 4096, 128, 0, True, False), (3, 400, 384, 0, False, False),
                                                     (2, 1024, 512, 384, True, True), (32, 64, 256, 0, False, False),
                                                     (2, 100, 64, 64, True, True)])
def test_group_norm_bwd(ops, ns, rows, c1, c2, silu, film):
    x1 = (_rand(ns * rows, c1, seed=20) * 2 + 0.5).half()
    x2 = (_rand(ns * rows, c2, seed=21) - 0.3).half() if c2 else None
    C = c1 + c2
    gamma = (_rand(C, seed=22) * 0.5 + 1).requires_grad_(True)
    beta = _rand(C, seed=23).requires_grad_(True)
    fm = (_rand(ns, 2 * C, seed=24) * 0.5).requires_grad_(True) if film else None
    dy = _rand(ns * rows, C, scale=0.5, seed=25).half()
    dx1, dx2, dg, db, dfilm = ops.group_norm_bwd(x1, gamma, beta, ns, dy, x2=x2, film=fm, ns_per_batch=1, silu=silu)
    xx = (x1.float() if x2 is None else torch.cat([x1.float(), x2.float()], dim=1)).requires_grad_(True)
    xr = xx.reshape(ns, rows, C).permute(0, 2, 1)
    ref = F.group_norm(xr, 32, gamma, beta, eps=1e-5)
    if film:
        ref = ref * (1 + fm[:, :C, None]) + fm[:, C:, None]
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1).reshape(ns * rows, C)
    inputs = [xx, gamma, beta] + ([fm] if film else [])
    grads = torch.autograd.grad(ref, inputs, dy.float())
    errs = {"dx1": rel_l2(dx1, grads[0][:, :c1]), "dgamma": rel_l2(dg, grads[1]), "dbeta": rel_l2(db, grads[2])}
    if c2:
        errs["dx2"] = rel_l2(dx2, grads[0][:, c1:])
    if film:
        errs["dfilm"] = rel_l2(dfilm, grads[3])
    _report(f"group_norm_bwd ns={ns} rows={rows} C={c1}+{c2} silu={silu} film={film}", errs)


def test_group_norm_temporal_bwd(ops):
    B, Fr, P, C = 2, 16, 64, 256
    x = (_rand(B, Fr, P, C, seed=30) * 1.5 + 0.2).half()
    gamma = (_rand(C, seed=31) * 0.5 + 1).requires_grad_(True)
    beta = _rand(C, seed=32).requires_grad_(True)
    dy = _rand(B, Fr, P, C, scale=0.5, seed=33).half()
    dx, dg, db = ops.group_norm_temporal_bwd(x, gamma, dy)
    xf = x.float().requires_grad_(True)
    xr = xf.permute(0, 2, 3, 1).reshape(B * P, C, Fr)
    ref = F.group_norm(xr, 32, gamma, beta, eps=1e-5).reshape(B, P, C, Fr).permute(0, 3, 1, 2)
    gx, gg, gb = torch.autograd.grad(ref, [xf, gamma, beta], dy.float())
    _report("group_norm_temporal_bwd", {"dx": rel_l2(dx, gx), "dgamma": rel_l2(dg, gg), "dbeta": rel_l2(db, gb)})


def test_resample_bwd(ops):
    errs = {}
    for mode, shape in (("vpool", (3, 16, 16, 128)), ("vup", (3, 8, 8, 64)), ("apool", (2, 400, 64)), ("aup", (2, 100, 128))):
        x = _rand(*shape, seed=40).half()
        xf = x.float().requires_grad_(True)
        if mode == "vpool":
            y = F.avg_pool2d(xf.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
        elif mode == "vup":
            y = F.interpolate(xf.permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
        elif mode == "apool":
            y = F.avg_pool1d(xf.permute(0, 2, 1), 4).permute(0, 2, 1)
        else:
            y = F.interpolate(xf.permute(0, 2, 1), scale_factor=4, mode="nearest").permute(0, 2, 1)
        dy = _rand(*y.shape, seed=41).half()
        (gx,) = torch.autograd.grad(y, xf, dy.float())
        errs[mode] = rel_l2(ops.resample_bwd(dy.contiguous(), mode, shape), gx)
    _report("resample_bwd", errs, tol=1e-3)


# ------------------------------------------------------------------ attention
@pytest.mark.parametrize("B,Fr,P,C,heads", [(2, 16, 64, 512, 4), (1, 16, 256, 384, 4), (1, 8, 16, 64, 1)])
def test_temporal_attention_bwd(ops, B, Fr, P, C, heads):
    d = C // heads
    qkv = _rand(B, Fr, P, 3 * C, seed=60).half()
    d_out = _rand(B, Fr, P, C, scale=0.5, seed=61).half()
    dqkv = ops.temporal_attention_bwd(qkv, d_out, heads)
    qf = qkv.float().requires_grad_(True)
    q, k, v = qf.split(C, dim=-1)
    q = q.reshape(B, Fr, P, heads, d)
    k = k.reshape(B, Fr, P, heads, d)
    v = v.reshape(B, Fr, P, heads, d)
    s = torch.einsum("bfphd,bgphd->bphfg", q, k) / math.sqrt(d)
    ref = torch.einsum("bphfg,bgphd->bfphd", torch.softmax(s, dim=-1), v).reshape(B, Fr, P, C)
    (g,) = torch.autograd.grad(ref, qf, d_out.float())
    _report(f"temporal_attention_bwd {B}x{Fr}x{P}x{C}", {"dq": rel_l2(dqkv[..., :C], g[..., :C]), "dk": rel_l2(dqkv[..., C:2 * C], g[..., C:2 * C]),
                                                          "dv": rel_l2(dqkv[..., 2 * C:], g[..., 2 * C:])})


def _ref_attention(q, k, v, B, heads, d, n_blocks, q_blk, k_blk, win, shift):
    qb = q.reshape(B, n_blocks, q_blk, heads, d)
    kb = k.reshape(B, n_blocks * k_blk, heads, d)
    vb = v.reshape(B, n_blocks * k_blk, heads, d)
    tot = n_blocks * k_blk
    outs, lses = [], []
    for i in range(n_blocks):
        idx = (torch.arange(win * k_blk, device=q.device) + (i + shift) * k_blk) % tot
        s = torch.einsum("bqhd,bkhd->bhqk", qb[:, i], kb[:, idx]) / math.sqrt(d)
        lses.append(torch.logsumexp(s, dim=-1) * 1.4426950408889634)   # [B, h, q] in log2 units
        outs.append(torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, dim=-1), vb[:, idx]))
    out = torch.stack(outs, dim=1).reshape(B * n_blocks * q_blk, heads * d)
    lse = torch.stack(lses, dim=2).permute(1, 0, 2, 3).reshape(heads, B * n_blocks * q_blk)
    return out, lse


@pytest.mark.parametrize("B,heads,d,n_blocks,q_blk,k_blk,win,shift", [
    (1, 4, 64, 2, 256, 256, 1, 0),     # spatial self, 16x16
    (1, 2, 64, 16, 1024, 400, 1, 5),   # cross video->audio ds2
    (1, 2, 64, 16, 400, 1024, 1, 15),  # cross audio->video ds2
    (1, 6, 64, 16, 100, 256, 4, 11),   # cross a->v ds4, wrapping window
    (2, 8, 64, 16, 25, 64, 8, 8),      # cross a->v ds8
    (1, 8, 64, 16, 64, 25, 16, 0),     # middle block, full window
    (1, 4, 96, 4, 256, 256, 1, 0),     # spatial self d=96
    (1, 4, 128, 16, 64, 64, 1, 0),     # spatial self d=128
    (2, 4, 128, 1, 400, 400, 1, 0),    # audio self
])
def test_attention_bwd(ops, B, heads, d, n_blocks, q_blk, k_blk, win, shift):
    C = heads * d
    nq = B * n_blocks * q_blk
    nk = B * n_blocks * k_blk
    qm = _rand(nq, 3 * C, seed=50).half()
    same = q_blk == k_blk and win == 1 and shift == 0
    km = qm if same else _rand(nk, 3 * C, seed=51).half()
    d_out = _rand(nq, C, scale=0.5, seed=52).half()
    out, lse, dq, dk, dv = ops.attention_fwd_bwd(qm, km, km, 0, C, 2 * C, B, heads, d, n_blocks, q_blk, k_blk, d_out, win, shift)
    qf = qm[:, :C].float().requires_grad_(True)
    kf = km[:, C:2 * C].float().requires_grad_(True)
    vf = km[:, 2 * C:].float().requires_grad_(True)
    ref, ref_lse = _ref_attention(qf, kf, vf, B, heads, d, n_blocks, q_blk, k_blk, win, shift)
    gq, gk, gv = torch.autograd.grad(ref, [qf, kf, vf], d_out.float())
    errs = {"out": rel_l2(out, ref), "lse": (lse - ref_lse).abs().max().item() / max(ref_lse.abs().max().item(), 1.0),
            "dq": rel_l2(dq, gq), "dk": rel_l2(dk, gk), "dv": rel_l2(dv, gv)}
    _report(f"attention_bwd B={B} h={heads} d={d} blocks={n_blocks} q={q_blk} k={k_blk} win={win} shift={shift}", errs, tol=5e-3)


# ------------------------------------------------------------------ narrow heads
def test_head_bwd(ops):
    errs = {}
    B, Fr, H, W, C, n = 2, 8, 16, 16, 128, 3
    x = _rand(B, Fr, H, W, C, seed=12).half()
    wt = _rand(n, C, 3, 3, 3, scale=0.05, seed=13).requires_grad_(True)
    dout = _rand(B, Fr, n, H, W, scale=0.5, seed=14)
    taps = [(kx - 1, ky - 1, kt - 1) for kt in range(3) for ky in range(3) for kx in range(3)]
    dx, dw, db = ops.head_bwd(x, wt, dout, 5, [W, H, Fr, B], taps, [1, W, n * H * W, Fr * n * H * W], H * W)
    xf = x.float().requires_grad_(True)
    y = F.conv3d(xf.permute(0, 4, 1, 2, 3), wt, None, padding=1).permute(0, 2, 1, 3, 4)
    gx, gw = torch.autograd.grad(y, [xf, wt], dout)
    errs.update(v_dx=rel_l2(dx, gx), v_dw=rel_l2(dw, gw.reshape(n, C, 27)), v_db=rel_l2(db, dout.sum((0, 1, 3, 4))))
    xa = _rand(2, 1000, 128, seed=15).half()
    wa = _rand(1, 128, 3, scale=0.05, seed=16).requires_grad_(True)
    da = _rand(2, 1, 1000, scale=0.5, seed=17)
    dx, dw, db = ops.head_bwd(xa, wa, da, 3, [1000, 2], [(k - 1, 0, 0) for k in range(3)], [1, 1000], 1000)
    xf = xa.float().requires_grad_(True)
    gx, gw = torch.autograd.grad(F.conv1d(xf.permute(0, 2, 1), wa, None, padding=1), [xf, wa], da)
    errs.update(a_dx=rel_l2(dx, gx), a_dw=rel_l2(dw, gw), a_db=rel_l2(db, da.sum((0, 2))))
    _report("head_bwd", errs)
