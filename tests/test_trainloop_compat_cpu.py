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
