"""Production training step (multimodal_training_losses forward + backward) on one B200: finiteness, timing with CUDA
events, gradient norm.  usage: python tools/gpu_train_check.py [B] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
model, diffusion = bench.build_b200(torch.device("cuda"))
model.convert_to_fp32()
model.train()
g = torch.Generator().manual_seed(0)
x0 = {"video": torch.randn(B, *bench.VIDEO_SIZE, generator=g).clamp(-1, 1).cuda(),
      "audio": torch.randn(B, *bench.AUDIO_SIZE, generator=g).clamp(-1, 1).cuda()}
t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(99)).cuda()
times = []
for i in range(K + 2):
    model.zero_grad(set_to_none=True)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    terms = diffusion.multimodal_training_losses(model, x0, t)
    loss = terms["loss"].mean()
    e1.record()
    loss.backward()
    e2.record()
    torch.cuda.synchronize()
    times.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
gn = torch.sqrt(sum((p.grad.float() ** 2).sum() for p in model.parameters())).item()
fin = all(torch.isfinite(p.grad).all().item() for p in model.parameters())
fw = sorted(x[0] for x in times[2:])[len(times[2:]) // 2]
bw = sorted(x[1] for x in times[2:])[len(times[2:]) // 2]
print(f"B={B} loss {loss.item():.5f} grad-norm {gn:.4e} finite {fin} fwd {fw:.2f} ms bwd {bw:.2f} ms "
      f"-> {B / ((fw + bw) * 1e-3):.1f} sample-steps/s; mem {torch.cuda.max_memory_allocated() / 1e9:.1f} GB torch "
      f"+ train plan; bwd launches {model.num_backward_launches(B)}")
# per-family backward profile (gradients are garbage after this: profiling only)
model.zero_grad(set_to_none=True)
terms = diffusion.multimodal_training_losses(model, x0, t)
torch.cuda.synchronize()
steps = model.profile_backward(B, reps=2)
fam = {}
for s in steps:
    k = s["kind"].split(":")[0] if not s["kind"].startswith(("wgrad", "dgrad")) else s["kind"]
    f = fam.setdefault(k, [0.0, 0])
    f[0] += s["ms"]; f[1] += 1
tot = sum(v[0] for v in fam.values())
print(f"backward profile (un-graphed, per-step events): total {tot:.2f} ms over {len(steps)} steps")
for k, (ms, n) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:32s} {ms:8.3f} ms {n:4d} steps {100 * ms / tot:5.1f}%")
