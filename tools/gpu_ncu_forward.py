"""One production p_sample step (B from argv, no CUDA graph) for ncu captures."""
import os, sys
os.environ["MMD_NO_GRAPH"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model, diffusion = bench.build_b200(torch.device("cuda"))
x = {"video": torch.randn(B, *bench.VIDEO_SIZE).cuda(), "audio": torch.randn(B, *bench.AUDIO_SIZE).cuda()}
t = torch.full((B,), 500, device="cuda", dtype=torch.long)
with torch.no_grad():
    diffusion.p_sample(model, x, t)
torch.cuda.synchronize()
print("done")
