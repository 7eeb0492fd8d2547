#!/usr/bin/env python
"""Run one of the reference's UNCHANGED scripts (py_scripts/*.py) with this repository's sm_100a hot path underneath.

    python tools/run_reference_script.py [--stock] [--synthetic-data] py_scripts/multimodal_sample_sr.py -- <script flags>

What happens:
  * sys.path = [tools/ref_shims, <reference root>, repo root]: the stand-ins answer `import mpi4py / blobfile / moviepy /
    wandb` (absent from this image, none of them on the denoising path);
  * unless --stock: mm_diffusion_b200.compat.install() aliases the five hot-path modules under the reference's names
    before the script imports them, so `create_model_and_diffusion`, `p_sample_loop`, `DPM_Solver`, `TrainLoop(model=…)`
    run on libmmdiff.so.  --stock runs the reference untouched (PyTorch eager) — the GPU baseline arm;
  * `mm_diffusion.evaluator` (FVD / AudioCLIP metrics: tensorflow, chainer, ignite — OUT of scope, not in the image) is
    replaced by a stub whose eval_multimodal raises if a script ever reaches it (the scripts only do when --ref_path exists);
  * --synthetic-data replaces `mm_diffusion.multimodal_datasets.load_data` (mp4 decoding through PyAV / moviepy, data loading
    is OUT of scope) by a generator of Landscape-shaped random batches, so multimodal_train.py runs without a dataset.
The script itself is executed with runpy under __main__, byte for byte as shipped.

Reference root: $MMD_REFERENCE, else baseline/_ref (tools/install_reference.py), else /root/reference.
"""
from __future__ import annotations

import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root() -> str:
    for cand in (os.environ.get("MMD_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "mm_diffusion")):
            return cand
    raise SystemExit("no reference checkout found (set MMD_REFERENCE or run tools/install_reference.py)")


def prepare(stock: bool = False, synthetic_data: bool = False) -> str:
    ref = reference_root()
    shims = os.path.join(ROOT, "tools", "ref_shims")
    for p in (ROOT, ref, shims):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    if not stock:
        from mm_diffusion_b200 import compat
        compat.install()
    ev = types.ModuleType("mm_diffusion.evaluator")

    def eval_multimodal(*a, **k):
        raise RuntimeError("mm_diffusion.evaluator is stubbed: FVD / AudioCLIP evaluation is outside the denoising path")
    ev.eval_multimodal = eval_multimodal
    sys.modules["mm_diffusion.evaluator"] = ev
    if synthetic_data:
        ds = types.ModuleType("mm_diffusion.multimodal_datasets")

        def load_data(*, data_dir=None, batch_size, video_size, audio_size, **_):
            import torch
            g = torch.Generator().manual_seed(1234 + int(os.environ.get("RANK", "0")))
            while True:
                yield (torch.randn(batch_size, *video_size, generator=g).clamp(-1, 1),
                       torch.randn(batch_size, *audio_size, generator=g).clamp(-1, 1))
        ds.load_data = load_data
        sys.modules["mm_diffusion.multimodal_datasets"] = ds
    return ref


def main():
    argv = sys.argv[1:]
    stock = "--stock" in argv
    synth = "--synthetic-data" in argv
    argv = [a for a in argv if a not in ("--stock", "--synthetic-data")]
    if not argv:
        raise SystemExit(__doc__)
    script = argv[0]
    rest = argv[1:]
    if rest and rest[0] == "--":
        rest = rest[1:]
    ref = prepare(stock, synth)
    path = script if os.path.isabs(script) else os.path.join(ref, script)
    sys.argv = [path] + rest
    try:
        runpy.run_path(path, run_name="__main__")
    finally:
        if os.environ.get("MMD_REPORT_NATIVE"):
            lib = sys.modules.get("mm_diffusion_b200._lib")
            print(f"[mmd] native library loaded: {bool(lib is not None and getattr(lib, '_lib', None) is not None)}", flush=True)


if __name__ == "__main__":
    main()
