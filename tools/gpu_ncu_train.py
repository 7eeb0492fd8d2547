"""One production training step (forward + backward, B from argv) for ncu captures of the backward kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model, diffusion = bench.build_b200(torch.device("cuda"))
model.convert_to_fp32()
model.train()
g = torch.Generator().manual_seed(0)
x0 = {"video": torch.randn(B, *bench.VIDEO_SIZE, generator=g).clamp(-1, 1).cuda(),
      "audio": torch.randn(B, *bench.AUDIO_SIZE, generator=g).clamp(-1, 1).cuda()}
t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(99)).cuda()
loss = diffusion.multimodal_training_losses(model, x0, t)["loss"].mean()
loss.backward()
torch.cuda.synchronize()
print("done", loss.item())
