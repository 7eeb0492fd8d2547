"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/mmdiff.h declares,
the MultimodalUNet shim reproduces the reference's parameter inventory, factories / schedule tables match the
oracle (pinned to the reference), and the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import random
import re

import numpy as np
import pytest
import torch

from oracle.mmdiff_oracle import DiffusionOracle, UNetConfig, param_shapes, shift_bounds
from tests.util_golden import build_b200_model, cfg_of, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mm_diffusion_b200 import _lib
    header = open(os.path.join(ROOT, "include", "mmdiff.h")).read()
    declared = set(re.findall(r"\b(mmd_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in mmdiff.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert b"sm_100a" in _lib.load().mmd_version()


@pytest.mark.parametrize("name", ["small", "production"])
def test_shim_parameter_inventory_matches_reference(name):
    fx = load_golden(name)
    cfg = cfg_of(fx)
    model = build_b200_model(cfg, device="cpu")
    mine = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    ref = [(k, tuple(s)) for k, s in fx["state_dict_keys"]]
    assert mine == ref
    assert [n for n, _ in model.named_parameters()] == [k for k, _ in ref]  # registration order (DDP / fp16 masters)
    assert model.shift_bounds == shift_bounds(cfg)
    # the draws consume Python's global RNG exactly like the reference forward did
    random.seed(7)
    drawn = [s for s, b in zip(model.draw_shifts(), model.shift_bounds) if b >= 0]
    assert drawn == [v for _, _, v in fx["forward_shift_draws"]]


def test_state_dict_roundtrip_and_tolerant_loader():
    cfg = cfg_of(load_golden("small"))
    a = build_b200_model(cfg, device="cpu")
    b = build_b200_model(cfg, device="cpu")
    sd = {k: v.clone() for k, v in a.state_dict().items()}
    key = "input_blocks.1.0.video_in_layers.2.video_conv_spatial.weight"
    sd_bad = dict(sd)
    sd_bad[key] = torch.zeros(3, 3)  # shape-mismatched entry is dropped by load_state_dict_ (reference :1033-1054)
    b.load_state_dict_(sd_bad, is_strict=False)
    for k, v in b.state_dict().items():
        if k != key:
            assert torch.equal(v, sd[k])
    # zero-initialised sites of the reference
    assert a.state_dict()["video_out.2.video_conv.weight"].abs().sum() == 0
    assert a.state_dict()["middle_blocks.1.video_proj_out.video_conv.bias"].abs().sum() == 0
    assert a.state_dict()["input_blocks.1.0.video_in_layers.0.GroupNorm.weight"].min() == 1


def test_no_cpu_fallback():
    from mm_diffusion_b200._lib import MmdError
    cfg = cfg_of(load_golden("small"))
    model = build_b200_model(cfg, device="cpu")
    v = torch.zeros(1, *cfg.video_size)
    a = torch.zeros(1, *cfg.audio_size)
    with pytest.raises(MmdError):
        with torch.no_grad():
            model(v, a, torch.zeros(1))


def test_unsupported_configs_are_rejected():
    from mm_diffusion_b200._lib import MmdError
    from mm_diffusion_b200.unet import MultimodalUNet
    with pytest.raises(MmdError):  # head dim 32 has no tcgen05 instantiation
        MultimodalUNet([8, 3, 16, 16], [1, 2048], 64, 3, 1, 1, [1], [1], True, [1], [-1], channel_mult=(1,),
                       num_heads=2, num_head_channels=64, use_scale_shift_norm=True)
    with pytest.raises(NotImplementedError):
        MultimodalUNet([8, 3, 16, 16], [1, 2048], 64, 3, 1, 1, [1], [1], True, [1], [-1], channel_mult=(1,),
                       num_heads=1, num_head_channels=64, use_scale_shift_norm=False)
    # audio length not a multiple of the frame count at a cross-attention level: the reference's remainder segment
    # (multimodal_unet.py:547-548) is rejected at construction, not at the first forward
    with pytest.raises(MmdError, match="remainder segment"):
        MultimodalUNet([8, 3, 16, 16], [1, 2052], 64, 3, 1, 1, [1], [1], True, [1], [-1], channel_mult=(1,),
                       num_heads=1, num_head_channels=64, use_scale_shift_norm=True)


def test_factories_and_schedule_tables_match_oracle():
    from mm_diffusion_b200 import script_util as su
    d = su.model_and_diffusion_defaults()
    assert set(d) >= {"video_size", "audio_size", "cross_attention_windows", "timestep_respacing", "use_fp16"}
    diff = su.create_gaussian_diffusion()
    o = DiffusionOracle(1000)
    for attr in ("betas", "alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                 "posterior_mean_coef1", "posterior_mean_coef2", "posterior_variance"):
        assert np.array_equal(getattr(diff, attr), getattr(o, attr)), attr
    assert np.array_equal(diff._model_log_variance, o.fixed_large_log_variance)
    assert diff.timestep_map == list(range(1000)) and diff.num_timesteps == 1000
    spaced = su.create_gaussian_diffusion(timestep_respacing="ddim25")
    assert spaced.num_timesteps == 25 and spaced.timestep_map[:3] == [0, 40, 80]
    from mm_diffusion_b200.respace import space_timesteps
    assert space_timesteps(300, [10, 15, 20]) == space_timesteps(300, "10,15,20")
    assert len(space_timesteps(300, [10, 15, 20])) == 45
    import argparse
    p = argparse.ArgumentParser()
    su.add_dict_to_argparser(p, d)
    ns = p.parse_args(["--use_fp16", "True", "--num_channels", "64"])
    assert ns.use_fp16 is True and ns.num_channels == 64


def test_generic_cpu_diffusion_math_matches_oracle():
    """The un-fused (torch) statements in gaussian_diffusion.py — used off the hot path — agree with the oracle."""
    from mm_diffusion_b200 import script_util as su
    diff = su.create_gaussian_diffusion()
    o = DiffusionOracle(1000)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 5, generator=g)
    eps = torch.randn(2, 3, 5, generator=g)
    z = torch.randn(2, 3, 5, generator=g)
    t = torch.tensor([0, 731])
    assert torch.allclose(diff.q_sample(x, t, noise=eps), o.q_sample(x, t, eps), atol=1e-6)
    x0 = diff._predict_xstart_from_eps(x, t, eps).clamp(-1, 1)
    mean, _, _ = diff.q_posterior_mean_variance(x0, x, t)
    sample = mean + (t != 0).float().view(-1, 1, 1) * torch.exp(0.5 * diff._gather(diff._model_log_variance, t, x)) * z
    ref_sample, ref_x0 = o.p_sample_tail(x, eps, t, z)
    assert torch.allclose(sample, ref_sample, atol=1e-6) and torch.allclose(x0, ref_x0, atol=1e-6)


def test_compat_install_aliases_reference_module_names():
    import sys
    from mm_diffusion_b200 import compat
    saved = {k: sys.modules.get(k) for k in compat._MAP}
    try:
        names = compat.install()
        assert set(names) == set(compat._MAP)
        import importlib
        su = importlib.import_module("mm_diffusion.multimodal_script_util")
        assert su.create_model_and_diffusion.__module__ == "mm_diffusion_b200.script_util"
        un = importlib.import_module("mm_diffusion.multimodal_unet")
        assert un.MultimodalUNet.__module__ == "mm_diffusion_b200.unet"
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_training_plan_dry_run_is_consistent():
    """Host-only dry run of the training plan: the forward walk plus the backward tape (every tensor's gradient has a
    producer before its consumer's backward, segments / sources line up) for the small and the production topology."""
    import ctypes as C
    from mm_diffusion_b200 import _lib
    from oracle.make_golden import PRODUCTION, SMALL
    from tests.util_golden import build_b200_model
    lib = _lib.load()
    for cfg in (SMALL, PRODUCTION):
        m = build_b200_model(cfg, None, device="cpu")
        h = C.c_void_p()
        _lib.check(lib.mmd_model_create(C.byref(m._cfg), C.byref(h)))
        try:
            fwd = lib.mmd_model_workspace_bytes(h, 2)
            train = lib.mmd_model_train_workspace_bytes(h, 2)
            assert train > 0, lib.mmd_last_error()
            assert train > fwd   # kept activations + gradients
            assert lib.mmd_model_param_floats(h) >= sum(p.numel() for p in m.parameters())
            offs = [lib.mmd_model_param_offset(h, i) for i in range(lib.mmd_model_num_params(h))]
            assert offs == sorted(offs) and offs[0] == 0
        finally:
            lib.mmd_model_destroy(h)


def test_flat_parameter_storage_keeps_the_module_surface():
    """The shim re-homes its nn.Parameters as views of one flat fp32 buffer laid out like the library's arena (one device
    copy per optimizer step instead of one call per tensor): names / shapes / values, state_dict round trips, optimizer
    updates and deepcopy must behave like ordinary parameters (host-only: the handle is created without a device)."""
    import copy
    import ctypes as C
    from mm_diffusion_b200 import _lib
    cfg = cfg_of(load_golden("small"))
    model = build_b200_model(cfg, device="cpu")
    lib = _lib.load()
    h = C.c_void_p()
    _lib.check(lib.mmd_model_create(C.byref(model._cfg), C.byref(h)))
    model._handle, model._handle_device = h, torch.device("cpu")
    before = {k: v.clone() for k, v in model.state_dict().items()}
    params_before = list(model.parameters())
    assert model._flatten_parameters(torch.device("cpu")) and model._flat_params_ok()
    assert all(a is b for a, b in zip(params_before, model.parameters()))           # same Parameter objects
    after = model.state_dict()
    assert list(after) == list(before) and all(torch.equal(after[k], before[k]) for k in before)
    # the flat buffer has the library's layout: parameter i lives at mmd_model_param_offset(i)
    offs = model._param_offsets()
    flat = model._flat_params
    assert flat.numel() == lib.mmd_model_param_floats(h)
    for p, off in list(zip(model.parameters(), offs))[::37]:
        assert torch.equal(flat[off:off + p.numel()].view(p.shape), p.detach())
    # an optimizer step updates the buffer in place and bumps the version sum the forward checks
    v0 = sum(p._version for p in model.parameters())
    opt = torch.optim.SGD(model.parameters(), lr=0.5)
    for p in model.parameters():
        p.grad = torch.ones_like(p)
    snap = flat.clone()
    opt.step()
    assert sum(p._version for p in model.parameters()) > v0
    p0, off0 = next(iter(model.parameters())), offs[0]
    assert torch.allclose(flat[off0:off0 + p0.numel()], snap[off0:off0 + p0.numel()] - 0.5)
    # load_state_dict copies in place (aliasing kept); deepcopy yields independent tensors and is detected
    model.load_state_dict(before)
    model._plist = list(model.parameters())   # (the forward rebuilds this cache lazily)
    assert model._flat_params_ok() and torch.equal(next(iter(model.parameters())).detach(), before[next(iter(before))])
    clone = copy.deepcopy(model)   # the library handle stays behind; the copy's parameters are independent tensors
    assert clone._handle is None and clone._flat_params is None
    assert all(torch.equal(a, b) and a.data_ptr() != b.data_ptr() for a, b in zip(clone.parameters(), model.parameters()))
    model._handle = None
    lib.mmd_model_destroy(h)


def test_bench_clock_sampler_windows_to_the_timed_region():
    """bench.py starts nvidia-smi before the warm-up (its first sample takes longer than a short timed region) and counts
    only the samples after mark(); a region without a sample falls back to the nearest one and says so."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class _Proc:
        def terminate(self): pass
        def wait(self, timeout=None): return 0
        def kill(self): pass

    line = "0, {sm}, 1965, {pw}, 0x0000000000000004, Not Active, Not Active, Not Active, {cap}"
    s = bench.ClockSampler(0)
    s.proc = _Proc()
    s.lines = [line.format(sm=1200, pw=150.0, cap="Not Active")]          # warm-up sample: must not count
    s.mark()
    s.lines += [line.format(sm=1900, pw=600.0, cap="Active"), line.format(sm=1920, pw=610.0, cap="Active")]
    out = s.stop()
    assert out["samples"] == 2 and out["sm_mhz"] == 1910.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and "note" not in out
    s2 = bench.ClockSampler(0)
    s2.proc = _Proc()
    s2.lines = [line.format(sm=1800, pw=300.0, cap="Not Active")]
    s2.mark()                                                              # nothing arrives inside the region
    out2 = s2.stop()
    assert out2["samples"] == 1 and out2["sm_mhz"] == 1800.0 and "note" in out2
