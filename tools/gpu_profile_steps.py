"""Per-launch profile of one production forward (un-graphed, CUDA events): writes gpurun_out/steps_b{B}.txt."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model, diffusion = bench.build_b200(torch.device("cuda"))
x = {"video": torch.randn(B, *bench.VIDEO_SIZE).cuda(), "audio": torch.randn(B, *bench.AUDIO_SIZE).cuda()}
t = torch.full((B,), 500, device="cuda", dtype=torch.long)
with torch.no_grad():
    for _ in range(3):
        diffusion.p_sample(model, x, t)
    torch.cuda.synchronize()
    steps = model.profile(B, reps=5)
os.makedirs("gpurun_out", exist_ok=True)
with open(f"gpurun_out/steps_b{B}.txt", "w") as f:
    f.write(f"# B={B} total {sum(s['ms'] for s in steps):.3f} ms over {len(steps)} steps\n")
    for i, s in enumerate(steps):
        ms = max(s["ms"], 1e-6)
        f.write(f"{i:4d} {s['kind']:20s} {s['ms']*1e3:9.1f} us  {s['flops']/1e9:9.2f} GF {s['bytes']/1e6:9.2f} MB  "
                f"{s['flops']/ms/1e9:8.1f} TF/s {s['bytes']/ms/1e6:8.1f} GB/s\n")
print(open(f"gpurun_out/steps_b{B}.txt").read()[:200])
