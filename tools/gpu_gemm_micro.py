"""Back-to-back launch timing of the implicit-GEMM convolution on the production step's own shapes (B = 4, 16 frames,
64x64 video, channels 128/256/384/512 at 64/32/16/8 pixels), through the C ABI's measurement entry mmd_op_conv_timed.
Prints one line per shape: microseconds per launch, TFLOP/s, algorithmic GB/s and the time the tensor pipe / HBM alone
would need.  Kernel variants are selected by the library's environment switches (MMD_EG, MMD_GEMM_DBG, ...), one process
per variant:   MMD_EG=0 python tools/gpu_gemm_micro.py [reps]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mm_diffusion_b200 import ops

REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 200
PEAK_TF, PEAK_GBS = 2250.0, 7000.0
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)


def rnd(*shape):
    return (torch.randn(*shape, device=dev, generator=g) * 0.5).half()


def wgt(*shape):
    return torch.randn(*shape, device=dev, generator=g) * 0.05


def run(name, fn, m, k, n):
    """k = taps x input channels (GEMM K); algorithmic bytes = A read once + output written once + weights, fp16."""
    taps = 9 if name.startswith("3x3") else (3 if name.startswith("k3") else 1)
    us = fn()
    flops = 2.0 * m * k * n
    byts = 2.0 * (m * k / taps + m * n + k * n)
    rec = {"shape": name, "M": m, "K": k, "N": n, "us": round(us, 2), "tflops": round(flops / us * 1e-6, 1),
           "gbs": round(byts / us * 1e-3, 1), "us_tensor_peak": round(flops / PEAK_TF * 1e-6, 2),
           "us_hbm_peak": round(byts / PEAK_GBS * 1e-3, 2)}
    print(json.dumps(rec), flush=True)


B, F = 4, 16
N = B * F
LEVELS = [(64, 128), (32, 256), (16, 384), (8, 512)]
for res, c in LEVELS:
    m = N * res * res
    x4 = rnd(N, res, res, c)
    taps9 = [(kx - 1, ky - 1, 0) for ky in range(3) for kx in range(3)]
    w = wgt(c, c, 9)
    b = wgt(c)
    out = torch.empty((N, res, res, c), dtype=torch.float16, device=dev)
    run(f"3x3_r{res}_c{c}", lambda: ops._conv([x4], w, b, c, 4, [res, res, N], taps9, out=out, timed_reps=REPS), m, 9 * c, c)
    x5 = x4.view(B, F, res * res, c)
    w3 = wgt(c, c, 3)
    out5 = torch.empty((B, F, res * res, c), dtype=torch.float16, device=dev)
    taps3 = [(0, kk - 1, 0) for kk in range(3)]
    run(f"k3t_r{res}_c{c}", lambda: ops._conv([x5], w3, b, c, 4, [res * res, F, B], taps3, out=out5, timed_reps=REPS), m, 3 * c, c)
    xm = x4.view(m, c)
    w1 = wgt(c, c, 1)
    outm = torch.empty((m, c), dtype=torch.float16, device=dev)
    run(f"1x1_r{res}_c{c}", lambda: ops._conv([xm], w1, b, c, 2, [m], [(0, 0, 0)], out=outm, timed_reps=REPS), m, c, c)
    sums = torch.zeros((N, 32, 2), dtype=torch.float64, device=dev)
    run(f"1x1+stats_r{res}_c{c}", lambda: ops._conv([xm], w1, b, c, 2, [m], [(0, 0, 0)], out=outm, gn_sums=sums, gn_rows=res * res,
                                                   timed_reps=REPS), m, c, c)
    if res <= 32:
        wq = wgt(3 * c, c, 1)
        bq = wgt(3 * c)
        outq = torch.empty((m, 3 * c), dtype=torch.float16, device=dev)
        run(f"qkv_r{res}_c{c}", lambda: ops._conv([xm], wq, bq, 3 * c, 2, [m], [(0, 0, 0)], out=outq, timed_reps=REPS), m, c, 3 * c)
    # decoder-side 1x1 over the channel concat of two sources (skip connection)
    xs = rnd(m, c)
    w2 = wgt(c, 2 * c, 1)
    run(f"1x1cat_r{res}_c{c}", lambda: ops._conv([xm, xs], w2, b, c, 2, [m], [(0, 0, 0)], out=outm, timed_reps=REPS), m, 2 * c, c)
# audio: k = 3 dilated conv over [B, L, C]
for L, c in [(25600, 128), (6400, 256), (1600, 384), (400, 512)]:
    xa = rnd(B, L, c)
    wa = wgt(c, c, 3)
    ba = wgt(c)
    outa = torch.empty((B, L, c), dtype=torch.float16, device=dev)
    tapsa = [((kk - 1) * 2, 0, 0) for kk in range(3)]
    run(f"k3a_L{L}_c{c}", lambda: ops._conv([xa], wa, ba, c, 3, [L, B], tapsa, out=outa, timed_reps=REPS), B * L, 3 * c, c)
