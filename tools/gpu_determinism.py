"""Run-to-run determinism probe of the individual kernels and the whole forward (GPU)."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mm_diffusion_b200 import ops
from tests.util_golden import build_b200_model, cfg_of, load_golden, rel_l2
from oracle.mmdiff_oracle import synthetic_state_dict, draw_shifts

def rnd(*s, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*s, generator=g) * scale).cuda()

def same(name, f, n=4):
    outs = [f().clone() for _ in range(n)]
    torch.cuda.synchronize()
    diffs = [(o.float() - outs[0].float()).abs().max().item() for o in outs[1:]]
    print(f"{name:28s} max|diff| over {n-1} reruns: {max(diffs):.3e}", flush=True)

x = rnd(16384, 256, seed=1).half(); w = rnd(256, 256, seed=2, scale=0.05); b = rnd(256, seed=3)
same("conv_pointwise", lambda: ops.conv_pointwise([x], w, b))
xs = rnd(8, 32, 32, 256, seed=4).half(); ws = rnd(256, 256, 3, 3, seed=5, scale=0.03)
same("conv_spatial", lambda: ops.conv_spatial(xs, ws, b))
xt = rnd(2, 16, 256, 128, seed=6).half(); wt = rnd(128, 128, 3, seed=7, scale=0.05); bt = rnd(128, seed=8)
same("conv_temporal", lambda: ops.conv_temporal(xt, wt, bt))
xa = rnd(2, 6400, 256, seed=9).half(); wa = rnd(256, 256, 3, seed=10, scale=0.05)
same("conv_audio", lambda: ops.conv_audio(xa, wa, b, 16))
g1 = (rnd(2 * 16384, 256, seed=11) + 0.5).half(); gam = rnd(256, seed=12) * 0.3 + 1; bet = rnd(256, seed=13)
same("group_norm", lambda: ops.group_norm(g1, gam, bet, 2, silu=True))
qm = rnd(2 * 16 * 256, 3 * 384, seed=14).half()
same("attention_self_d96", lambda: ops.attention(qm, qm, qm, 0, 384, 768, 2, 4, 96, 16, 256, 256))
qv = rnd(16 * 1024, 768, seed=15).half(); ka = rnd(16 * 400, 768, seed=16).half()
same("attention_cross", lambda: ops.attention(qv, ka, ka, 0, 256, 512, 1, 4, 64, 16, 1024, 400, 1, 5))
qt = rnd(2, 16, 64, 3 * 512, seed=17).half()
same("temporal_attention", lambda: ops.temporal_attention(qt, 4))
gt = rnd(2, 16, 64, 256, seed=18).half()
same("group_norm_temporal", lambda: ops.group_norm_temporal(gt, gam, bet))

fx = load_golden("small"); cfg = cfg_of(fx)
model = build_b200_model(cfg, synthetic_state_dict(cfg, seed=0))
g = torch.Generator().manual_seed(3)
v = torch.randn(2, *cfg.video_size, generator=g).cuda(); a = torch.randn(2, *cfg.audio_size, generator=g).cuda()
t = torch.tensor([100, 800]).cuda(); sh = draw_shifts(cfg, random.Random(0))
with torch.no_grad():
    outs = [model(v, a, t, shifts=sh) for _ in range(4)]
print("small forward rerun rel-L2:", [f"{rel_l2(o[0], outs[0][0]):.2e}" for o in outs[1:]], flush=True)
with torch.no_grad():
    steps = model.profile(2, reps=1)
print("plan steps:", len(steps), "kernels:", sum(s['kernels'] for s in steps))
