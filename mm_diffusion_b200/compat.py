"""Drop-in installation under the reference's module names.

`install()` registers this package's modules in sys.modules as
    mm_diffusion.multimodal_unet / multimodal_gaussian_diffusion / multimodal_respace / multimodal_script_util /
    multimodal_dpm_solver_plus
so the reference's unchanged scripts (py_scripts/multimodal_sample_sr.py, multimodal_train.py, ...) and its
remaining modules (dist_util, logger, ...) pick up the B200-native hot path.
The reference checkout must be importable (sys.path) for the modules this package does not replace; call
install() before anything imports them.  See INTEGRATION.md.
"""
from __future__ import annotations

import importlib
import sys

_MAP = {
    "mm_diffusion.multimodal_unet": "mm_diffusion_b200.unet",
    "mm_diffusion.multimodal_gaussian_diffusion": "mm_diffusion_b200.gaussian_diffusion",
    "mm_diffusion.multimodal_respace": "mm_diffusion_b200.respace",
    "mm_diffusion.multimodal_script_util": "mm_diffusion_b200.script_util",
    "mm_diffusion.multimodal_dpm_solver_plus": "mm_diffusion_b200.dpm_solver",
}


_TAIL = {"mm_diffusion.fp16_util": "mm_diffusion_b200.fp16_util"}


def install(force: bool = True, optimizer_tail: bool = False):
    """Alias the hot-path modules under the reference's names; returns the list of names installed.
    optimizer_tail=True also aliases `mm_diffusion.fp16_util`: TrainLoop then builds the flat-buffer
    MixedPrecisionTrainer (one master parameter = the model's flat parameter buffer, one host sync per step)."""
    done = []
    mapping = dict(_MAP)
    if optimizer_tail:
        mapping.update(_TAIL)
    for ref_name, ours in mapping.items():
        if ref_name in sys.modules and not force:
            continue
        sys.modules[ref_name] = importlib.import_module(ours)
        done.append(ref_name)
    pkg = sys.modules.get("mm_diffusion")
    if pkg is not None:  # package already imported: also patch the attributes
        for ref_name in done:
            setattr(pkg, ref_name.split(".")[-1], sys.modules[ref_name])
    return done
