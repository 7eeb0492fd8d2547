"""Timeline of attention64_kernel from its clock64 stamps (library built with -DMMD_ATTN_TRACE, loaded through MMD_LIB):
one self-attention launch of the production 32x32 level (B = 4, 16 frames, 4 heads of 64, 1024 tokens per frame), then the
per-tile stamps of the first CTAs, in clocks relative to each CTA's first logits-ready stamp.

    MMD_NVCC_EXTRA=-DMMD_ATTN_TRACE MMD_LIB_OUT=mm_diffusion_b200/libmmdiff_trace.so python -m mm_diffusion_b200.build --force
    MMD_LIB=$PWD/mm_diffusion_b200/libmmdiff_trace.so python tools/gpu_attn_trace.py
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mm_diffusion_b200 import _lib, ops

B, heads, d, n_blocks, blk = 4, 4, 64, 16, 1024
Cc = heads * d
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B * n_blocks * blk, 3 * Cc, device="cuda", generator=g) * 0.5).half()
for _ in range(3):
    ops.attention(qkv, qkv, qkv, 0, Cc, 2 * Cc, B, heads, d, n_blocks, blk, blk, 1, 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.attention(qkv, qkv, qkv, 0, Cc, 2 * Cc, B, heads, d, n_blocks, blk, blk, 1, 0)
e1.record()
torch.cuda.synchronize()
print(f"self-attention 32x32 level (B=4): {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch (incl. op-level set-up)")
# cross-attention shape of the same level: 1024 video tokens per frame against 400 audio tokens per segment
kv = (torch.randn(B * n_blocks * 400, 3 * Cc, device="cuda", generator=g) * 0.5).half()
for _ in range(3):
    ops.attention(qkv, kv, kv, 0, Cc, 2 * Cc, B, heads, d, n_blocks, blk, 400, 1, 3)
e0.record()
for _ in range(20):
    ops.attention(qkv, kv, kv, 0, Cc, 2 * Cc, B, heads, d, n_blocks, blk, 400, 1, 3)
e1.record()
torch.cuda.synchronize()
print(f"cross-attention video->audio (1024 x 400 per block): {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
if not hasattr(C.CDLL(_lib.LIB_PATH), "mmd_attn_trace_dump"):
    sys.exit(0)
ops.attention(qkv, qkv, qkv, 0, Cc, 2 * Cc, B, heads, d, n_blocks, blk, blk, 1, 0)
torch.cuda.synchronize()
lib = _lib.load()
fn = getattr(C.CDLL(_lib.LIB_PATH), "mmd_attn_trace_dump")
buf = np.zeros(8 * 3 * 32 * 8, dtype=np.int64)
assert fn(buf.ctypes.data_as(C.c_void_p)) == 0
t = buf.reshape(8, 3, 32, 8)
names = {0: ["S ready", "PV(g-1) retired", "P written"], 1: ["K landed", "softmax done", "V landed", "PV issued"],
         2: ["K stage free", "V stage free"]}
for cta in range(2):
    t0 = t[cta, 0, 0, 0]
    print(f"== CTA {cta} (clocks relative to its first logits-ready stamp)")
    print("tile | softmax: S-ready  PVprev-done  P-written | MMA: K-landed  softmax-done  V-landed  PV-issued | TMA: K-free  V-free")
    for gi in range(2, 20):
        r = lambda role, ev: (int(t[cta, role, gi, ev] - t0) if t[cta, role, gi, ev] else -1)
        print(f"{gi:4d} | {r(0,0):8d} {r(0,1):8d} {r(0,2):8d} | {r(1,0):8d} {r(1,1):8d} {r(1,2):8d} {r(1,3):8d} | {r(2,0):8d} {r(2,1):8d}")
    d_s = np.diff(t[cta, 0, 2:20, 0])
    print("tile period (S-ready to S-ready):", d_s.tolist())
    soft = (t[cta, 0, 2:20, 2] - t[cta, 0, 2:20, 0]).tolist()
    print("softmax phase (S-ready -> P-written):", soft)
    gap = (t[cta, 0, 3:20, 0] - t[cta, 0, 2:19, 2]).tolist()
    print("P-written -> next S-ready:", gap)
