"""Operator-level Python wrappers over the C-ABI (tensors are channels-last fp16 CUDA tensors).

These mirror the reference's leaf modules so the parity tests read like tests of
the reference's own ops:
  conv_*            <-> VideoConv / AudioConv      (mm_diffusion/multimodal_unet.py:68-131)
  group_norm        <-> GroupNorm32 (+SiLU, +FiLM) (mm_diffusion/nn.py:16-33)
  attention         <-> QKVAttention / SingleModalQKVAttention (multimodal_unet.py:212-244, 498-564)
  resample          <-> Upsample / Downsample      (multimodal_unet.py:133-208)
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import MmdAttnDesc, MmdConvDesc, check, current_stream_ptr, ptr


def _conv(srcs, weight, bias, n, rank, dims, taps, out=None, out_f32=None, ostride=None, ostride_c=0, gn_sums=None,
          gn_rows=0, gn_in=None, timed_reps=0):
    lib = _lib.load()
    d = MmdConvDesc()
    d.rank = rank
    for i in range(4):
        d.dims[i] = dims[i] if i < len(dims) else 1
        d.box[i] = 0  # let the library pick the 128-token box
    d.n_src = len(srcs)
    keep = []
    for i, s in enumerate(srcs):
        assert s.dtype == torch.float16 and s.is_contiguous() and s.is_cuda
        d.src[i] = s.data_ptr()
        d.src_channels[i] = s.shape[-1]
        keep.append(s)
    d.n_taps = len(taps)
    for t, tp in enumerate(taps):
        for j in range(3):
            d.taps[t][j] = int(tp[j])
    w = weight.detach().to(torch.float32).contiguous()
    b = None if bias is None else bias.detach().to(torch.float32).contiguous()
    d.weight = w.data_ptr()
    d.bias = ptr(b)
    d.n = n
    d.out = ptr(out)
    d.out_f32 = ptr(out_f32)
    if ostride is not None:
        for i in range(4):
            d.ostride[i] = ostride[i] if i < len(ostride) else 0
    d.ostride_c = ostride_c
    if gn_sums is not None:
        assert gn_sums.dtype == torch.float64 and gn_sums.is_contiguous() and gn_sums.is_cuda
        d.gn_sums = gn_sums.data_ptr()
        d.gn_rows = gn_rows
    if gn_in is not None:   # GroupNorm32 (+FiLM, +SiLU) of source 0 applied on the GEMM's A operand
        g = gn_in["gamma"].detach().float().contiguous()
        b2 = gn_in["beta"].detach().float().contiguous()
        f = None if gn_in.get("film") is None else gn_in["film"].detach().float().contiguous()
        check(lib.mmd_op_conv_gn(C.byref(d), g.data_ptr(), b2.data_ptr(), ptr(f), 0 if f is None else f.shape[-1],
                                 int(gn_in["ns"]), int(gn_in.get("ns_per_batch", 1)), int(bool(gn_in.get("silu", False))),
                                 current_stream_ptr()))
    elif timed_reps > 0:   # measurement: average device microseconds per launch over timed_reps back-to-back launches
        us = C.c_float(0.0)
        check(lib.mmd_op_conv_timed(C.byref(d), int(timed_reps), C.byref(us), current_stream_ptr()))
        return us.value
    else:
        check(lib.mmd_op_conv(C.byref(d), current_stream_ptr()))
    return out if out is not None else out_f32


def conv_pointwise(srcs, weight, bias, gn_sums=None, gn_rows=0):
    """1x1(x1) conv over token matrices [M, C_i] (channel-concatenated sources) -> [M, Cout] fp16.

    gn_sums (float64 [M / gn_rows, 32, 2], zeroed by the caller) receives the GroupNorm(32) sum / sum of squares of
    the output per run of gn_rows tokens, reduced inside the GEMM epilogue."""
    m = srcs[0].numel() // srcs[0].shape[-1]
    n = weight.shape[0]
    out = torch.empty(srcs[0].shape[:-1] + (n,), dtype=torch.float16, device=srcs[0].device)
    w = weight.reshape(n, -1, 1)
    return _conv([s.reshape(m, s.shape[-1]) for s in srcs], w, bias, n, 2, [m], [(0, 0, 0)], out=out, gn_sums=gn_sums,
                 gn_rows=gn_rows)


def conv_pointwise_gn(srcs, weight, bias, gamma, beta, ns, film=None, ns_per_batch=1, silu=False, gn_sums=None, gn_rows=0):
    """conv_pointwise whose first source is GroupNorm32-normalised (+FiLM, +SiLU) on the fly: the ResBlock out_layers /
    attention-norm -> 1x1 conv pairs (multimodal_unet.py:459-470, :284, :664) without the normalised tensor in HBM.
    srcs[0]: [M, C] (ns domains of M / ns consecutive rows) or [B, L, C] (one domain per sample, ns == B)."""
    n = weight.shape[0]
    out = torch.empty(srcs[0].shape[:-1] + (n,), dtype=torch.float16, device=srcs[0].device)
    w = weight.reshape(n, -1, 1)
    gn_in = {"gamma": gamma, "beta": beta, "ns": ns, "film": film, "ns_per_batch": ns_per_batch, "silu": silu}
    if srcs[0].dim() == 3:
        B, L, _ = srcs[0].shape
        return _conv(list(srcs), w, bias, n, 3, [L, B], [(0, 0, 0)], out=out, gn_sums=gn_sums, gn_rows=gn_rows, gn_in=gn_in)
    m = srcs[0].shape[0]
    return _conv(list(srcs), w, bias, n, 2, [m], [(0, 0, 0)], out=out, gn_sums=gn_sums, gn_rows=gn_rows, gn_in=gn_in)


def conv_spatial(x, weight, bias):
    """3x3 'same' conv per frame: x [N,H,W,C] fp16, weight [Co,Ci,3,3] -> [N,H,W,Co]."""
    N, H, W, Ci = x.shape
    n = weight.shape[0]
    out = torch.empty((N, H, W, n), dtype=torch.float16, device=x.device)
    taps = [(kx - 1, ky - 1, 0) for ky in range(3) for kx in range(3)]
    return _conv([x], weight.reshape(n, Ci, 9), bias, n, 4, [W, H, N], taps, out=out)


def conv_temporal(x, weight, bias, gn_sums=None):
    """k=3 'same' conv along frames: x [B,F,P,C], weight [Co,Ci,3] -> [B,F,P,Co].

    gn_sums: optional float64 [B*F, 32, 2] fused output statistics (one domain per frame)."""
    B, F, P, Ci = x.shape
    n = weight.shape[0]
    out = torch.empty((B, F, P, n), dtype=torch.float16, device=x.device)
    taps = [(0, k - 1, 0) for k in range(3)]
    return _conv([x], weight.reshape(n, Ci, 3), bias, n, 4, [P, F, B], taps, out=out, gn_sums=gn_sums, gn_rows=P)


def conv_audio(x, weight, bias, dilation=1, gn_sums=None):
    """k=3 dilated 'same' conv: x [B,L,C], weight [Co,Ci,3] -> [B,L,Co].

    gn_sums: optional float64 [B, 32, 2] fused output statistics (one domain per sample)."""
    B, L, Ci = x.shape
    n = weight.shape[0]
    out = torch.empty((B, L, n), dtype=torch.float16, device=x.device)
    taps = [((k - 1) * dilation, 0, 0) for k in range(3)]
    return _conv([x], weight.reshape(n, Ci, 3), bias, n, 3, [L, B], taps, out=out, gn_sums=gn_sums, gn_rows=L)


def conv3d_head(x, weight, bias):
    """3x3x3 'same' conv to a few channels, fp32 NCHW output: x [B,F,H,W,C] -> [B,F,Co,H,W] fp32."""
    B, F, H, W, Ci = x.shape
    n = weight.shape[0]
    out = torch.empty((B, F, n, H, W), dtype=torch.float32, device=x.device)
    taps = [(kx - 1, ky - 1, kt - 1) for kt in range(3) for ky in range(3) for kx in range(3)]
    return _conv([x], weight.reshape(n, Ci, 27), bias, n, 5, [W, H, F, B], taps, out_f32=out,
                 ostride=[1, W, n * H * W, F * n * H * W], ostride_c=H * W)


def conv1d_head(x, weight, bias):
    """k=3 'same' conv to a few channels, fp32 NCL output: x [B,L,C] -> [B,Co,L] fp32."""
    B, L, Ci = x.shape
    n = weight.shape[0]
    out = torch.empty((B, n, L), dtype=torch.float32, device=x.device)
    taps = [(k - 1, 0, 0) for k in range(3)]
    return _conv([x], weight.reshape(n, Ci, 3), bias, n, 3, [L, B], taps, out_f32=out, ostride=[1, n * L],
                 ostride_c=L)


def group_norm(x, gamma, beta, ns, x2=None, film=None, ns_per_batch=1, silu=False):
    """GroupNorm(32) over `ns` domains of equal row count on [rows_total, C] (optionally concat of x and x2)."""
    lib = _lib.load()
    c1 = x.shape[-1]
    c2 = 0 if x2 is None else x2.shape[-1]
    rows_total = x.numel() // c1
    rows = rows_total // ns
    y = torch.empty(x.shape[:-1] + (c1 + c2,), dtype=torch.float16, device=x.device)
    g = gamma.detach().float().contiguous()
    b = beta.detach().float().contiguous()
    f = None if film is None else film.detach().float().contiguous()
    check(lib.mmd_op_group_norm(x.data_ptr(), c1, ptr(x2), c2, ns, rows, g.data_ptr(), b.data_ptr(), ptr(f),
                                0 if f is None else f.shape[-1], ns_per_batch, int(silu), y.data_ptr(),
                                current_stream_ptr()))
    return y


def group_norm_temporal(x, gamma, beta):
    """x [B,F,P,C]; statistics per (b, pixel, group) over F x C/32 (temporal attention norm)."""
    lib = _lib.load()
    B, F, P, Cc = x.shape
    y = torch.empty_like(x)
    g = gamma.detach().float().contiguous()
    b = beta.detach().float().contiguous()
    check(lib.mmd_op_group_norm_temporal(x.data_ptr(), y.data_ptr(), g.data_ptr(), b.data_ptr(), B, F, P, Cc,
                                         current_stream_ptr()))
    return y


def resample(x, mode):
    """mode: 'vpool' [N,H,W,C]->[N,H/2,W/2,C]; 'apool' [N,L,C]->[N,L/4,C]; 'vup' x2; 'aup' x4."""
    lib = _lib.load()
    if mode in ("vpool", "vup"):
        N, H, W, Cc = x.shape
        shape = (N, H // 2, W // 2, Cc) if mode == "vpool" else (N, H * 2, W * 2, Cc)
        m = 0 if mode == "vpool" else 2
    else:
        N, H, Cc = x.shape
        W = 1
        shape = (N, H // 4, Cc) if mode == "apool" else (N, H * 4, Cc)
        m = 1 if mode == "apool" else 3
    y = torch.empty(shape, dtype=torch.float16, device=x.device)
    check(lib.mmd_op_resample(x.data_ptr(), y.data_ptr(), m, N, H, W, Cc, current_stream_ptr()))
    return y


def attention(q_mat, k_mat, v_mat, q_col0, k_col0, v_col0, batch, heads, head_dim, n_blocks, q_blk, k_blk, win=1,
              shift=0):
    """Windowed attention over row-major fp16 matrices (column ranges select q/k/v and heads).

    Query block i (q_blk rows) of sample b attends key blocks (i+shift+j) mod n_blocks, j < win.
    Returns [q_rows, heads*head_dim] fp16."""
    lib = _lib.load()
    d = MmdAttnDesc()
    d.q, d.q_ld, d.q_col0, d.q_rows = q_mat.data_ptr(), q_mat.shape[1], q_col0, q_mat.shape[0]
    d.k, d.k_ld, d.k_col0, d.k_rows = k_mat.data_ptr(), k_mat.shape[1], k_col0, k_mat.shape[0]
    d.v, d.v_ld, d.v_col0 = v_mat.data_ptr(), v_mat.shape[1], v_col0
    out = torch.empty((q_mat.shape[0], heads * head_dim), dtype=torch.float16, device=q_mat.device)
    d.out, d.out_ld = out.data_ptr(), heads * head_dim
    d.batch, d.heads, d.head_dim = batch, heads, head_dim
    d.n_blocks, d.q_blk, d.k_blk, d.win, d.shift = n_blocks, q_blk, k_blk, win, shift
    check(lib.mmd_op_attention(C.byref(d), current_stream_ptr()))
    return out


def temporal_attention(qkv, heads):
    """qkv [B,F,P,3C] -> [B,F,P,C]: attention over the F axis per pixel."""
    lib = _lib.load()
    B, F, P, C3 = qkv.shape
    Cc = C3 // 3
    out = torch.empty((B, F, P, Cc), dtype=torch.float16, device=qkv.device)
    check(lib.mmd_op_temporal_attention(qkv.data_ptr(), out.data_ptr(), B, F, P, Cc, heads, current_stream_ptr()))
    return out


# --------------------------------------------------------------------------- backward operators (training)
def _conv_desc(srcs, n, rank, dims, taps, weight=None, out_f32_strides=None, ostride_c=0):
    d = MmdConvDesc()
    d.rank = rank
    for i in range(4):
        d.dims[i] = dims[i] if i < len(dims) else 1
        d.box[i] = 0
    d.n_src = len(srcs)
    for i, s in enumerate(srcs):
        assert s.dtype == torch.float16 and s.is_contiguous() and s.is_cuda
        d.src[i] = s.data_ptr()
        d.src_channels[i] = s.shape[-1]
    d.n_taps = len(taps)
    for t, tp in enumerate(taps):
        for j in range(3):
            d.taps[t][j] = int(tp[j])
    d.n = n
    if weight is not None:
        d.weight = weight.data_ptr()
    if out_f32_strides is not None:
        for i in range(4):
            d.ostride[i] = out_f32_strides[i] if i < len(out_f32_strides) else 0
    d.ostride_c = ostride_c
    return d


def conv_wgrad(srcs, dy, rank, dims, taps):
    """Weight / bias gradient of a conv over channels-last fp16 sources: returns (dW fp32 [n, c_total, n_taps], db [n])."""
    lib = _lib.load()
    n = dy.shape[-1]
    ctot = sum(s.shape[-1] for s in srcs)
    dw = torch.zeros((n, ctot, len(taps)), dtype=torch.float32, device=dy.device)
    db = torch.zeros((n,), dtype=torch.float32, device=dy.device)
    d = _conv_desc(srcs, n, rank, dims, taps)
    check(lib.mmd_op_conv_wgrad(C.byref(d), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), current_stream_ptr()))
    return dw, db


def conv_dgrad(src_channels, dy, weight, rank, dims, taps, src_index=0):
    """Data gradient wrt source `src_index`: weight fp32 [n, c_total, n_taps]; returns fp16 [..., c_src]."""
    lib = _lib.load()
    n = dy.shape[-1]
    w = weight.detach().to(torch.float32).contiguous()
    d = MmdConvDesc()
    d.rank = rank
    for i in range(4):
        d.dims[i] = dims[i] if i < len(dims) else 1
        d.box[i] = 0
    d.n_src = len(src_channels)
    for i, c in enumerate(src_channels):
        d.src_channels[i] = c
    d.n_taps = len(taps)
    for t, tp in enumerate(taps):
        for j in range(3):
            d.taps[t][j] = int(tp[j])
    d.n = n
    d.weight = w.data_ptr()
    dx = torch.empty(dy.shape[:-1] + (src_channels[src_index],), dtype=torch.float16, device=dy.device)
    check(lib.mmd_op_conv_dgrad(C.byref(d), dy.data_ptr(), src_index, dx.data_ptr(), current_stream_ptr()))
    return dx


def group_norm_bwd(x, gamma, beta, ns, dy, x2=None, film=None, ns_per_batch=1, silu=False):
    """Backward of group_norm(): returns (dx, dx2 | None, dgamma, dbeta, dfilm | None)."""
    lib = _lib.load()
    c1 = x.shape[-1]
    c2 = 0 if x2 is None else x2.shape[-1]
    Cc = c1 + c2
    rows = (x.numel() // c1) // ns
    g = gamma.detach().float().contiguous()
    b = beta.detach().float().contiguous()
    f = None if film is None else film.detach().float().contiguous()
    dx = torch.empty_like(x)
    dx2 = None if x2 is None else torch.empty_like(x2)
    dg = torch.zeros(Cc, dtype=torch.float32, device=x.device)
    db = torch.zeros(Cc, dtype=torch.float32, device=x.device)
    dfilm = None if f is None else torch.zeros_like(f)
    check(lib.mmd_op_group_norm_bwd(x.data_ptr(), c1, ptr(x2), c2, ns, rows, g.data_ptr(), b.data_ptr(), ptr(f),
                                    0 if f is None else f.shape[-1], ns_per_batch, int(silu), dy.data_ptr(), dx.data_ptr(),
                                    ptr(dx2), dg.data_ptr(), db.data_ptr(), ptr(dfilm), current_stream_ptr()))
    return dx, dx2, dg, db, dfilm


def group_norm_temporal_bwd(x, gamma, dy):
    lib = _lib.load()
    B, F, P, Cc = x.shape
    g = gamma.detach().float().contiguous()
    dx = torch.empty_like(x)
    dg = torch.zeros(Cc, dtype=torch.float32, device=x.device)
    db = torch.zeros(Cc, dtype=torch.float32, device=x.device)
    check(lib.mmd_op_group_norm_temporal_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), g.data_ptr(), dg.data_ptr(),
                                             db.data_ptr(), B, F, P, Cc, current_stream_ptr()))
    return dx, dg, db


def resample_bwd(dy, mode, in_shape):
    """Adjoint of resample(x, mode) for x of shape in_shape."""
    lib = _lib.load()
    if mode in ("vpool", "vup"):
        N, H, W, Cc = in_shape
        m = 0 if mode == "vpool" else 2
    else:
        N, H, Cc = in_shape
        W = 1
        m = 1 if mode == "apool" else 3
    dx = torch.empty(in_shape, dtype=torch.float16, device=dy.device)
    check(lib.mmd_op_resample_bwd(dy.data_ptr(), dx.data_ptr(), m, N, H, W, Cc, current_stream_ptr()))
    return dx


def temporal_attention_bwd(qkv, d_out, heads):
    lib = _lib.load()
    B, F, P, C3 = qkv.shape
    dqkv = torch.empty_like(qkv)
    check(lib.mmd_op_temporal_attention_bwd(qkv.data_ptr(), d_out.data_ptr(), dqkv.data_ptr(), B, F, P, C3 // 3, heads,
                                            current_stream_ptr()))
    return dqkv


def attention_fwd_bwd(q_mat, k_mat, v_mat, q_col0, k_col0, v_col0, batch, heads, head_dim, n_blocks, q_blk, k_blk, d_out,
                      win=1, shift=0):
    """Forward + backward of attention(): returns (out, lse [heads, q_rows], dq [q_rows, C], dk [k_rows, C], dv [k_rows, C])."""
    lib = _lib.load()
    Cc = heads * head_dim
    d = MmdAttnDesc()
    d.q, d.q_ld, d.q_col0, d.q_rows = q_mat.data_ptr(), q_mat.shape[1], q_col0, q_mat.shape[0]
    d.k, d.k_ld, d.k_col0, d.k_rows = k_mat.data_ptr(), k_mat.shape[1], k_col0, k_mat.shape[0]
    d.v, d.v_ld, d.v_col0 = v_mat.data_ptr(), v_mat.shape[1], v_col0
    out = torch.empty((q_mat.shape[0], Cc), dtype=torch.float16, device=q_mat.device)
    d.out, d.out_ld = out.data_ptr(), Cc
    d.batch, d.heads, d.head_dim = batch, heads, head_dim
    d.n_blocks, d.q_blk, d.k_blk, d.win, d.shift = n_blocks, q_blk, k_blk, win, shift
    lse = torch.zeros((heads, q_mat.shape[0]), dtype=torch.float32, device=q_mat.device)
    dq = torch.zeros((q_mat.shape[0], Cc), dtype=torch.float16, device=q_mat.device)
    dk = torch.zeros((k_mat.shape[0], Cc), dtype=torch.float16, device=q_mat.device)
    dv = torch.zeros((k_mat.shape[0], Cc), dtype=torch.float16, device=q_mat.device)
    # dq / dk / dv are separate matrices here (leading dimension C, column 0)
    check(lib.mmd_op_attention_fwd_bwd(C.byref(d), d_out.data_ptr(), lse.data_ptr(), dq.data_ptr(), dk.data_ptr(),
                                       dv.data_ptr(), Cc, 0, 0, 0, current_stream_ptr()))
    return out, lse, dq, dk, dv


def head_bwd(x, weight, dout, rank, dims, taps, ostride, ostride_c, need_dx=True):
    """Adjoint of conv3d_head / conv1d_head: returns (dx fp16 | None, dW fp32 [n, C, taps], db [n])."""
    lib = _lib.load()
    n = weight.shape[0]
    w = weight.detach().to(torch.float32).reshape(n, x.shape[-1], len(taps)).contiguous()
    d = _conv_desc([x], n, rank, dims, taps, weight=w, out_f32_strides=ostride, ostride_c=ostride_c)
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.zeros_like(w)
    db = torch.zeros(n, dtype=torch.float32, device=x.device)
    check(lib.mmd_op_head_bwd(C.byref(d), dout.data_ptr(), ptr(dx), dw.data_ptr(), db.data_ptr(), current_stream_ptr()))
    return dx, dw, db
