"""Training backward parity on the GPU: gradients of multimodal_training_losses through
the sm_100a backward (dgrad on the forward implicit-GEMM kernel, tcgen05 wgrad, flash-attention backward, GroupNorm /
resampling / head / time-embedding adjoints) against torch.autograd through the CPU oracle (PyTorch fp32 restatement
of the reference, pinned to reference goldens) on the same seeded inputs, weights and window shifts.
Tolerance (fp16 activations and activation gradients, fp32 accumulation, vs fp32): global rel-L2 over all 300+
parameter gradients <= 3e-2, every large tensor's own rel-L2 <= 6e-2; loss rel-err <= 1e-2."""
import random

import pytest
import torch

from oracle.mmdiff_oracle import DiffusionOracle, draw_shifts, synthetic_state_dict
from tests.util_golden import build_b200_model, cfg_of, load_golden, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from mm_diffusion_b200.script_util import create_gaussian_diffusion
    fx = load_golden("small")
    cfg = cfg_of(fx)
    sd = synthetic_state_dict(cfg, seed=fx["weight_seed"])
    model = build_b200_model(cfg, sd)
    return fx, cfg, sd, model, create_gaussian_diffusion()


def _data(cfg, B, seed):
    g = torch.Generator().manual_seed(seed)
    x0 = {"video": torch.randn(B, *cfg.video_size, generator=g).clamp(-1, 1),
          "audio": torch.randn(B, *cfg.audio_size, generator=g).clamp(-1, 1)}
    noise = {"video": torch.randn(B, *cfg.video_size, generator=g), "audio": torch.randn(B, *cfg.audio_size, generator=g)}
    return x0, noise


def _oracle_grads(cfg, sd, x0, noise, t, shifts, wrt_inputs=False):
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    orc = DiffusionOracle(1000)
    if wrt_inputs:
        xv = x0["video"].clone().requires_grad_(True)
        xa = x0["audio"].clone().requires_grad_(True)
        terms = orc.training_losses(sdg, cfg, {"video": xv, "audio": xa}, t, noise, shifts)
        terms["loss"].mean().backward()
        return terms, {k: v.grad for k, v in sdg.items()}, xv.grad, xa.grad
    terms = orc.training_losses(sdg, cfg, x0, t, noise, shifts)
    terms["loss"].mean().backward()
    return terms, {k: v.grad for k, v in sdg.items()}, None, None


def test_training_losses_backward_matches_oracle(setup):
    fx, cfg, sd, model, diffusion = setup
    B = 2
    x0, noise = _data(cfg, B, 321)
    t = torch.tensor([700, 31])
    shifts = draw_shifts(cfg, random.Random(9))
    ref_terms, ref_grads, _, _ = _oracle_grads(cfg, sd, x0, noise, t, shifts)

    model.train()
    model.zero_grad(set_to_none=True)
    random.seed(9)   # the model draws the same shifts from the global RNG
    terms = diffusion.multimodal_training_losses(model, {k: v.cuda() for k, v in x0.items()}, t.cuda(),
                                                 noise={k: v.cuda() for k, v in noise.items()})
    loss = terms["loss"].mean()
    loss.backward()
    torch.cuda.synchronize()
    model.eval()
    lerr = abs(loss.item() - ref_terms["loss"].mean().item()) / abs(ref_terms["loss"].mean().item())
    num = den = 0.0
    rows = []
    missing = []
    for name, p in model.named_parameters():
        g_ref = ref_grads[name]
        if p.grad is None:
            missing.append(name)
            continue
        g = p.grad.detach().float().cpu()
        num += (g - g_ref).double().pow(2).sum().item()
        den += g_ref.double().pow(2).sum().item()
        rows.append((rel_l2(g, g_ref), g_ref.norm().item(), name))
    glob = (num / max(den, 1e-30)) ** 0.5
    rows.sort(reverse=True)
    print(f"[bwd-model] loss rel-err {lerr:.2e}; global grad rel-L2 {glob:.3e} over {len(rows)} tensors; fwd launches "
          f"{model.num_launches(B)}, bwd launches {model.num_backward_launches(B)}")
    for e, nrm, name in rows[:12]:
        print(f"[bwd-model]   worst: {e:.3e}  |g_ref| {nrm:.3e}  {name}")
    assert not missing, f"parameters without gradient: {missing[:5]}"
    assert lerr < 1e-2
    assert glob < 3e-2, glob
    big = [r for r in rows if r[1] > 1e-3 * (den ** 0.5)]
    assert all(e < 6e-2 for e, _, _ in big), [r for r in big if r[0] >= 6e-2][:5]


def test_input_gradients_match_oracle(setup):
    """d loss / d inputs (gradient-guided conditional sampling differentiates wrt the target modality's x_t)."""
    fx, cfg, sd, model, diffusion = setup
    B = 2
    x0, noise = _data(cfg, B, 77)
    t = torch.tensor([400, 5])
    shifts = draw_shifts(cfg, random.Random(4))
    orc = DiffusionOracle(1000)
    xv = x0["video"].clone().requires_grad_(True)
    xa = x0["audio"].clone().requires_grad_(True)
    from oracle.mmdiff_oracle import unet_forward
    ev, ea = unet_forward(sd, cfg, xv, xa, t, shifts)
    ((ev - noise["video"]) ** 2).mean().add(((ea - noise["audio"]) ** 2).mean()).backward()

    model.eval()
    for p in model.parameters():
        p.requires_grad_(False)
    try:
        v = x0["video"].cuda().requires_grad_(True)
        a = x0["audio"].cuda().requires_grad_(True)
        vo, ao = model(v, a, t.cuda(), shifts=shifts)
        loss = ((vo.float() - noise["video"].cuda()) ** 2).mean() + ((ao.float() - noise["audio"].cuda()) ** 2).mean()
        gv, ga = torch.autograd.grad(loss, [v, a])
    finally:
        for p in model.parameters():
            p.requires_grad_(True)
    errs = {"d_video": rel_l2(gv, xv.grad), "d_audio": rel_l2(ga, xa.grad)}
    print("[bwd-model] input gradients:", {k: f"{e:.3e}" for k, e in errs.items()})
    assert all(e < 5e-2 for e in errs.values()), errs


def test_gradient_guided_step_runs(setup):
    """One step of conditional_p_sample_loop with class_scale > 0 (reference :722-819): finite, right shapes, and the
    conditioned modality equals q_sample of the condition."""
    fx, cfg, sd, model, diffusion = setup
    from mm_diffusion_b200.script_util import create_gaussian_diffusion
    short = create_gaussian_diffusion(timestep_respacing="2")
    B = 2
    shape = {"video": (B, *cfg.video_size), "audio": (B, *cfg.audio_size)}
    g = torch.Generator().manual_seed(5)
    cond_audio = (0.1 * torch.randn(B, *cfg.audio_size, generator=g)).cuda()
    model.eval()
    out = short.conditional_p_sample_loop(model, shape, use_fp16=False, model_kwargs={"audio": cond_audio}, progress=False,
                                          class_scale=3.0, device=torch.device("cuda"))
    assert out["video"].shape == shape["video"] and out["audio"].shape == shape["audio"]
    assert torch.isfinite(out["video"]).all() and torch.isfinite(out["audio"]).all()


def test_production_backward_matches_oracle():
    """Production topology (133.7 M parameters, 16x3x64x64 + 25600, every kernel instantiation of the real network:
    256/384/512 channels, head dims 64/96/128, windows 1/4/8/16) at batch 1: parameter gradients vs oracle autograd
    on the host cores.  Same tolerances as the small configuration."""
    fx = load_golden("production")
    cfg = cfg_of(fx)
    sd = synthetic_state_dict(cfg, seed=fx["weight_seed"])
    x0, noise = _data(cfg, 1, 99)
    t = torch.tensor([500])
    shifts = draw_shifts(cfg, random.Random(3))
    torch.set_num_threads(min(32, torch.get_num_threads()))
    ref_terms, ref_grads, _, _ = _oracle_grads(cfg, sd, x0, noise, t, shifts)
    from mm_diffusion_b200.script_util import create_gaussian_diffusion
    diffusion = create_gaussian_diffusion()
    model = build_b200_model(cfg, sd)
    model.train()
    random.seed(3)
    terms = diffusion.multimodal_training_losses(model, {k: v.cuda() for k, v in x0.items()}, t.cuda(),
                                                 noise={k: v.cuda() for k, v in noise.items()})
    loss = terms["loss"].mean()
    loss.backward()
    torch.cuda.synchronize()
    lerr = abs(loss.item() - ref_terms["loss"].mean().item()) / abs(ref_terms["loss"].mean().item())
    num = den = 0.0
    rows = []
    for name, p in model.named_parameters():
        g_ref = ref_grads[name]
        assert p.grad is not None, name
        g = p.grad.detach().float().cpu()
        assert torch.isfinite(g).all(), name
        num += (g - g_ref).double().pow(2).sum().item()
        den += g_ref.double().pow(2).sum().item()
        rows.append((rel_l2(g, g_ref), g_ref.norm().item(), name))
    glob = (num / max(den, 1e-30)) ** 0.5
    rows.sort(reverse=True)
    print(f"[bwd-model] production: loss rel-err {lerr:.2e}; global grad rel-L2 {glob:.3e} over {len(rows)} tensors; "
          f"bwd launches {model.num_backward_launches(1)}")
    for e, nrm, name in rows[:8]:
        print(f"[bwd-model]   worst: {e:.3e}  |g_ref| {nrm:.3e}  {name}")
    assert lerr < 1e-2
    assert glob < 3e-2, glob
    big = [r for r in rows if r[1] > 1e-3 * (den ** 0.5)]
    assert all(e < 6e-2 for e, _, _ in big), [r for r in big if r[0] >= 6e-2][:5]


def test_parameter_update_reaches_every_packed_layout(setup):
    """A training loop changes the parameters every step: after an in-place update (and an interleaved no-grad forward,
    as in EMA / sampling during training) the next step must use the new weights in the forward packs AND in the
    transposed packs of the data gradients.  Gradients after the update vs oracle autograd at the updated weights."""
    fx, cfg, sd, model, diffusion = setup
    B = 2
    x0, noise = _data(cfg, B, 55)
    t = torch.tensor([250, 900])
    shifts = draw_shifts(cfg, random.Random(6))
    model.train()
    model.zero_grad(set_to_none=True)
    random.seed(6)
    diffusion.multimodal_training_losses(model, {k: v.cuda() for k, v in x0.items()}, t.cuda(),
                                         noise={k: v.cuda() for k, v in noise.items()})["loss"].mean().backward()
    g = torch.Generator().manual_seed(8)
    new_sd = {}
    with torch.no_grad():
        for name, p in model.named_parameters():   # a large, structured update (a plain SGD step would be tiny)
            delta = 0.02 * torch.randn(p.shape, generator=g)
            p.add_(delta.cuda())
            new_sd[name] = sd[name] + delta
        model.eval()
        model(x0["video"].cuda(), x0["audio"].cuda(), t.cuda(), shifts=shifts)   # no-grad forward sees the update first
    try:
        ref_terms, ref_grads, _, _ = _oracle_grads(cfg, new_sd, x0, noise, t, shifts)
        model.train()
        model.zero_grad(set_to_none=True)
        random.seed(6)
        loss = diffusion.multimodal_training_losses(model, {k: v.cuda() for k, v in x0.items()}, t.cuda(),
                                                    noise={k: v.cuda() for k, v in noise.items()})["loss"].mean()
        loss.backward()
        torch.cuda.synchronize()
        num = sum((p.grad.float().cpu() - ref_grads[n]).double().pow(2).sum().item() for n, p in model.named_parameters())
        den = sum(ref_grads[n].double().pow(2).sum().item() for n, _ in model.named_parameters())
        glob = (num / den) ** 0.5
        lerr = abs(loss.item() - ref_terms["loss"].mean().item()) / abs(ref_terms["loss"].mean().item())
        print(f"[bwd-model] after parameter update: loss rel-err {lerr:.2e}, global grad rel-L2 {glob:.3e}")
        assert lerr < 1e-2 and glob < 3e-2, (lerr, glob)
    finally:
        model.eval()
        model.load_state_dict(sd, strict=True)   # the module-scoped fixture is shared


def test_training_with_dropout_matches_oracle_with_exported_masks(setup):
    """nn.Dropout(p=0.1) of the ResBlock out_layers (multimodal_unet.py:370-386; the shipped training flags,
    ssh_scripts/multimodal_train.sh:4): the keep-mask statistic is p +- 1 %, masks differ between sites and forwards,
    and with the masks of the forward exported to the oracle the loss and every parameter gradient agree with
    torch.autograd (same tolerances as the dropout-free test)."""
    fx, cfg, sd, model, diffusion = setup
    B = 2
    x0, noise = _data(cfg, B, 555)
    t = torch.tensor([600, 120])
    shifts = draw_shifts(cfg, random.Random(11))
    model.train()
    model.dropout = 0.1
    try:
        model.zero_grad(set_to_none=True)
        random.seed(11)
        torch.manual_seed(2024)
        terms = diffusion.multimodal_training_losses(model, {k: v.cuda() for k, v in x0.items()}, t.cuda(),
                                                     noise={k: v.cuda() for k, v in noise.items()})
        loss = terms["loss"].mean()
        masks = [(mv.cpu(), ma.cpu()) for mv, ma in model.dropout_masks(B)]
        loss.backward()
        torch.cuda.synchronize()
        grads = {n: p.grad.detach().float().cpu() for n, p in model.named_parameters()}
        # a second forward draws different masks
        random.seed(11)
        with torch.enable_grad():
            diffusion.multimodal_training_losses(model, {k: v.cuda() for k, v in x0.items()}, t.cuda(),
                                                 noise={k: v.cuda() for k, v in noise.items()})
        masks2 = [(mv.cpu(), ma.cpu()) for mv, ma in model.dropout_masks(B)]
    finally:
        model.dropout = 0
        model.eval()
    kept = sum(int(mv.sum()) + int(ma.sum()) for mv, ma in masks)
    total = sum(mv.numel() + ma.numel() for mv, ma in masks)
    drop_rate = 1.0 - kept / total
    per_site = [1.0 - float(mv.float().mean()) for mv, _ in masks] + [1.0 - float(ma.float().mean()) for _, ma in masks]
    print(f"[dropout] {len(masks)} ResBlocks, {total} elements, drop rate {drop_rate:.5f}, per-site range "
          f"{min(per_site):.4f}..{max(per_site):.4f}")
    assert abs(drop_rate - 0.1) < 0.001                     # p +- 1 %
    assert all(abs(r - 0.1) < 0.02 for r in per_site)       # every site, small ones included
    assert any((a[0] != b[0]).any() for a, b in zip(masks, masks2)), "masks did not change between forwards"
    same_shape = [(i, j) for i in range(len(masks)) for j in range(i + 1, len(masks)) if masks[i][0].shape == masks[j][0].shape]
    assert same_shape and all((masks[i][0] != masks[j][0]).any() for i, j in same_shape), "two sites share a mask"

    thresh = int(0.1 * 65536.0 + 0.5)
    drop = {"scale": 1.0 / (1.0 - thresh / 65536.0), "masks": masks}
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref_terms = DiffusionOracle(1000).training_losses(sdg, cfg, x0, t, noise, shifts, dropout=drop)
    ref_terms["loss"].mean().backward()
    lerr = abs(loss.item() - ref_terms["loss"].mean().item()) / abs(ref_terms["loss"].mean().item())
    num = sum((grads[n] - sdg[n].grad).double().pow(2).sum().item() for n in grads)
    den = sum(sdg[n].grad.double().pow(2).sum().item() for n in grads)
    glob = (num / den) ** 0.5
    print(f"[dropout] loss rel-err {lerr:.2e}; global gradient rel-L2 vs oracle autograd with the same masks {glob:.3e}")
    assert lerr < 1e-2
    assert glob < 3e-2, glob


def test_stale_forward_is_rejected(setup):
    """One tape per batch size: a backward whose forward was overwritten by a later differentiable forward must fail
    loudly instead of differentiating the wrong activations."""
    from mm_diffusion_b200._lib import MmdError
    fx, cfg, sd, model, diffusion = setup
    B = 2
    x0, noise = _data(cfg, B, 99)
    t = torch.tensor([10, 900]).cuda()
    model.train()
    try:
        a = diffusion.multimodal_training_losses(model, {k: v.cuda() for k, v in x0.items()}, t)["loss"].mean()
        b = diffusion.multimodal_training_losses(model, {k: v.cuda() for k, v in x0.items()}, t)["loss"].mean()
        with pytest.raises((MmdError, RuntimeError), match="stale forward"):
            a.backward()
        b.backward()   # the latest forward is still differentiable
    finally:
        model.zero_grad(set_to_none=True)
        model.eval()


def test_flat_buffer_trainer_step_on_device():
    """SURVEY.md §8 row f1 on the GPU: one training step through mm_diffusion_b200.fp16_util.MixedPrecisionTrainer (fp16
    loss-scaling protocol) — the sm_100a backward writes the flat gradient buffer the trainer's single master parameter
    aliases, AdamW steps it, and the next forward must run on the updated weights.  Checked against (a) a per-tensor AdamW
    step taken from the same gradients and (b) the oracle forward at the trainer's checkpoint dict."""
    from mm_diffusion_b200 import fp16_util as ours
    from mm_diffusion_b200.script_util import create_gaussian_diffusion
    from oracle.mmdiff_oracle import unet_forward
    fx = load_golden("small")
    cfg = cfg_of(fx)
    sd = synthetic_state_dict(cfg, seed=fx["weight_seed"])
    model = build_b200_model(cfg, sd).train()     # its own instance: the trainer switches it to flat gradients
    diffusion = create_gaussian_diffusion()
    trainer = ours.MixedPrecisionTrainer(model=model, use_fp16=True, fp16_scale_growth=1e-3)
    lr, wd, eps = 1e-3, 0.01, 1e-8
    opt = torch.optim.AdamW(trainer.master_params, lr=lr, weight_decay=wd, eps=eps)
    B = 2
    x0, noise = _data(cfg, B, 91)
    t = torch.tensor([100, 800])
    trainer.zero_grad()
    random.seed(12)
    loss = diffusion.multimodal_training_losses(model, {k: v.cuda() for k, v in x0.items()}, t.cuda(),
                                                noise={k: v.cuda() for k, v in noise.items()})["loss"].mean()
    trainer.backward(loss)
    torch.cuda.synchronize()
    scale = 2.0 ** trainer.lg_loss_scale
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    grads = {n: (p.grad.detach().float() / scale).clone() for n, p in model.named_parameters()}
    assert all(torch.isfinite(g).all() for g in grads.values())
    assert model.flat_grad is not None and all(p.grad.data_ptr() >= model.flat_grad.data_ptr() for p in model.parameters())
    assert trainer.optimize(opt) is True
    assert abs(trainer.lg_loss_scale - (ours.INITIAL_LOG_LOSS_SCALE + 1e-3)) < 1e-9
    # (a) first AdamW step from zero moments: p <- p (1 - lr wd) - lr g / (|g| + eps)
    worst = 0.0
    for n, p in model.named_parameters():
        g = grads[n]
        want = before[n] * (1 - lr * wd) - lr * g / (g.abs() + eps)
        worst = max(worst, (p.detach() - want).abs().max().item())
    assert worst < 2e-6, worst
    # (b) the library's packed weights follow: forward at the new weights vs the oracle at the trainer's checkpoint
    ckpt = {k: v.detach().float().cpu() for k, v in trainer.master_params_to_state_dict(trainer.master_params).items()}
    assert list(ckpt) == list(sd)
    moved = max((ckpt[k] - sd[k]).abs().max().item() for k in sd)
    assert moved > 5e-4, moved
    g2 = torch.Generator().manual_seed(17)
    v = torch.randn(B, *cfg.video_size, generator=g2)
    a = torch.randn(B, *cfg.audio_size, generator=g2)
    shifts = draw_shifts(cfg, random.Random(4))
    model.eval()
    with torch.no_grad():
        ov, oa = unet_forward(ckpt, cfg, v, a, t, shifts)
        ev, ea = model(v.cuda(), a.cuda(), t.cuda(), shifts=shifts)
        o0v, _ = unet_forward(sd, cfg, v, a, t, shifts)
    rv, ra = rel_l2(ev, ov), rel_l2(ea, oa)
    print(f"[flat trainer] AdamW step max |dp| err {worst:.2e}; forward at updated weights rel-L2 video {rv:.2e} audio {ra:.2e}; "
          f"old-vs-new output {rel_l2(o0v, ov):.2e}")
    assert rv < 3e-3 and ra < 3e-3, (rv, ra)
    assert rel_l2(o0v, ov) > 3 * rv    # the update is visible well above the parity tolerance
