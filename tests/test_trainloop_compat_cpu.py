"""CPU: the reference's own MixedPrecisionTrainer (mm_diffusion/fp16_util.py, driven by TrainLoop in
multimodal_train_util.py) on the drop-in MultimodalUNet shim: master-parameter flattening by registration order
(fp16_util.py:81-93), loss-scaled optimisation, master -> model copies and the checkpoint key schema
(master_params_to_state_dict) must work unchanged, in fp32 and in the use_fp16 configuration.
Build-container test: it imports the UNMODIFIED reference from /root/reference and is skipped where that is absent."""
import os
import sys
import types

import pytest
import torch

REF = os.environ.get("MMD_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "mm_diffusion")), reason="reference checkout not present")


def _import_fp16_util():
    for name, attrs in (("mpi4py", {"MPI": types.SimpleNamespace(COMM_WORLD=None)}), ("blobfile", {})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from mm_diffusion import fp16_util
    return fp16_util


@pytest.mark.parametrize("use_fp16", [False, True])
def test_reference_mixed_precision_trainer_drives_the_shim(use_fp16):
    fp16_util = _import_fp16_util()
    from oracle.make_golden import SMALL
    from oracle.mmdiff_oracle import synthetic_state_dict
    from tests.util_golden import build_b200_model
    sd = synthetic_state_dict(SMALL, seed=0)
    model = build_b200_model(SMALL, sd, device="cpu")
    model.train()
    trainer = fp16_util.MixedPrecisionTrainer(model=model, use_fp16=use_fp16, fp16_scale_growth=1e-3)
    assert model.dtype == (torch.float16 if use_fp16 else torch.float32)   # convert_to_fp16() was called by the trainer
    opt = torch.optim.AdamW(trainer.master_params, lr=1e-2, weight_decay=0.0)
    g = torch.Generator().manual_seed(3)
    scale = 2.0 ** trainer.lg_loss_scale if use_fp16 else 1.0
    before = {k: v.detach().clone() for k, v in model.named_parameters()}
    for p in model.parameters():   # what loss.backward() through the sm_100a path leaves behind: fp32 .grad on every parameter
        p.grad = torch.randn(p.shape, generator=g) * 1e-3 * scale
    assert trainer.optimize(opt) is True
    changed = sum(int(not torch.equal(before[k], v.detach())) for k, v in model.named_parameters())
    assert changed == len(before), f"only {changed} of {len(before)} parameters were updated"
    # the update magnitude is lr-sized (Adam), i.e. the loss scale was divided out again
    k0 = "input_blocks.0.0.video_conv.video_conv_spatial.weight"
    step = (dict(model.named_parameters())[k0].detach() - before[k0]).abs().max().item()
    assert 1e-3 < step < 5e-2, step
    # checkpoint schema (TrainLoop.save, multimodal_train_util.py:357-360): same keys and shapes as the model's state_dict
    ckpt = trainer.master_params_to_state_dict(trainer.master_params)
    own = model.state_dict()
    assert list(ckpt.keys()) == list(own.keys())
    assert all(tuple(ckpt[k].shape) == tuple(own[k].shape) for k in own)
    # and back: state_dict_to_master_params (resume path, :117-127)
    masters = trainer.state_dict_to_master_params(ckpt)
    assert sum(m.numel() for m in masters) == sum(p.numel() for p in model.parameters())
    trainer.zero_grad()
    assert all(p.grad is None or float(p.grad.abs().sum()) == 0.0 for p in model.parameters())


def test_script_util_surface_matches_reference():
    """model_and_diffusion_defaults(): same keys, same default values; argparse helpers behave the same; the factory
    accepts exactly the reference's keyword set (every script builds its parser from these)."""
    _import_fp16_util()
    import argparse
    import inspect
    from mm_diffusion import multimodal_script_util as ref
    from mm_diffusion_b200 import script_util as ours
    rd, od = ref.model_and_diffusion_defaults(), ours.model_and_diffusion_defaults()
    assert list(rd.keys()) == list(od.keys())
    assert rd == od
    assert set(inspect.signature(ref.create_model_and_diffusion).parameters) == \
        set(inspect.signature(ours.create_model_and_diffusion).parameters)
    pr, po = argparse.ArgumentParser(), argparse.ArgumentParser()
    ref.add_dict_to_argparser(pr, rd)
    ours.add_dict_to_argparser(po, od)
    argv = ["--use_fp16", "False", "--video_size", "16,3,64,64", "--cross_attention_windows", "1,4,8", "--num_channels", "128"]
    assert vars(pr.parse_args(argv)) == vars(po.parse_args(argv))
    assert ref.args_to_dict(pr.parse_args(argv), rd.keys()) == ours.args_to_dict(po.parse_args(argv), od.keys())
    for v in ("yes", "True", "0", "n"):
        assert ref.str2bool(v) == ours.str2bool(v)


def test_flat_buffer_trainer_matches_reference_trainer_step():
    """mm_diffusion_b200.fp16_util.MixedPrecisionTrainer (one flat master parameter, one host sync
    per step) takes the same optimizer / EMA step as the reference's trainer from the same gradients, for fp32 and for
    the fp16 loss-scaling protocol (overflow -> skip + lower scale), and writes reference-schema checkpoints.
    Host-only: gradients are injected where the sm_100a backward would have written them."""
    import copy
    import sys
    import types
    from tests.util_golden import build_b200_model, cfg_of, load_golden
    from mm_diffusion_b200 import fp16_util as ours
    if "mpi4py" not in sys.modules:
        m = types.ModuleType("mpi4py"); m.MPI = types.SimpleNamespace(COMM_WORLD=None); sys.modules["mpi4py"] = m
        sys.modules.setdefault("blobfile", types.ModuleType("blobfile"))
    ref_root = "/root/reference" if __import__("os").path.isdir("/root/reference/mm_diffusion") else \
        __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.dirname(__file__)), "baseline", "_ref")
    if not __import__("os").path.isdir(__import__("os").path.join(ref_root, "mm_diffusion")):
        pytest.skip("reference not available")
    sys.path.insert(0, ref_root)
    for k in [k for k in sys.modules if k == "mm_diffusion" or k.startswith("mm_diffusion.")]:
        del sys.modules[k]   # make sure the unmodified reference modules are the ones imported below
    from mm_diffusion import fp16_util as ref_fp16
    from mm_diffusion.nn import update_ema as ref_update_ema

    cfg = cfg_of(load_golden("small"))
    for use_fp16 in (False, True):
        torch.manual_seed(0)
        model_a = build_b200_model(cfg, device="cpu").train()    # driven by the reference trainer (autograd-style .grad tensors)
        model_b = copy.deepcopy(model_a)                          # driven by the flat trainer
        ta = ref_fp16.MixedPrecisionTrainer(model=model_a, use_fp16=use_fp16, fp16_scale_growth=1e-3)
        tb = ours.MixedPrecisionTrainer(model=model_b, use_fp16=use_fp16, fp16_scale_growth=1e-3)
        assert len(tb.master_params) == 1 and tb.master_params[0].numel() >= sum(p.numel() for p in model_b.parameters())
        oa = torch.optim.AdamW(ta.master_params, lr=1e-3, weight_decay=0.01)
        ob = torch.optim.AdamW(tb.master_params, lr=1e-3, weight_decay=0.01)
        ema_a, ema_b = copy.deepcopy(ta.master_params), copy.deepcopy(tb.master_params)
        g = torch.Generator().manual_seed(3)
        for step, overflow in enumerate([False, True, False]):
            scale = 2 ** ta.lg_loss_scale if use_fp16 else 1.0
            assert not use_fp16 or ta.lg_loss_scale == tb.lg_loss_scale
            ta.zero_grad(); tb.zero_grad()
            buf, views = model_b._grad_buffer(torch.device("cpu"), model_b._flat_params.numel())
            buf.zero_()
            for pa, pb, vb in zip(model_a.parameters(), model_b.parameters(), views):
                gr = torch.randn(pa.shape, generator=g) * 1e-2 * scale
                if overflow and use_fp16 and pa.ndim > 1:
                    gr[..., 0] = float("inf")
                pa.grad = gr.clone()
                vb.copy_(gr)
                pb.grad = vb          # what _UNetFunction.backward does in flat-gradient mode
            took_a, took_b = ta.optimize(oa), tb.optimize(ob)
            assert took_a == took_b == (not (overflow and use_fp16))
            if took_a:
                ref_update_ema(ema_a, ta.master_params, rate=0.9)
                ours.update_ema(ema_b, tb.master_params, rate=0.9)
        sa = ta.master_params_to_state_dict(ta.master_params)
        sb = tb.master_params_to_state_dict(tb.master_params)
        assert list(sa) == list(sb)
        worst = max((sa[k].float() - sb[k].float()).abs().max().item() for k in sa)
        assert worst < 1e-6, worst
        ea = ta.master_params_to_state_dict(ema_a)
        eb = tb.master_params_to_state_dict(ema_b)
        assert max((ea[k].float() - eb[k].float()).abs().max().item() for k in ea) < 1e-6
        # the model sees the update without any copy: its parameters are views of the master buffer
        assert all(torch.equal(p.detach(), sb[n]) for n, p in model_b.named_parameters())
        # round trip through the reference checkpoint schema
        back = tb.state_dict_to_master_params(sb)
        assert torch.equal(back[0].detach(), tb.master_params[0].detach())
