"""CPU: the built library really is Blackwell-native — the tensor-core kernels of the path contain tcgen05.mma
(SASS UTCHMMA), TMA loads (UTMALDG) and TMEM loads (LDTM) for sm_100a, read from libmmdiff.so with cuobjdump
(B200_PROFILING.md: the mnemonics that prove tcgen05 / TMA).  No GPU needed."""
import re
import shutil
import subprocess

import pytest

from mm_diffusion_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass_by_function():
    try:
        out = subprocess.run([CUOBJDUMP, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300)
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    assert out.returncode == 0, out.stderr[:500]
    assert "sm_100a" in out.stdout or "SM100a" in out.stdout.replace("_", "") or "sm_100" in out.stdout
    funcs = {}
    for chunk in re.split(r"\n\s*Function : ", out.stdout)[1:]:
        name, _, body = chunk.partition("\n")
        funcs[name.strip()] = body
    return funcs


@pytest.mark.parametrize("kernel,needs", [
    ("conv_gemm_kernel", ("UTCHMMA", "UTMALDG", "LDTM", "UTMASTG")),   # implicit-GEMM conv: tcgen05.mma, TMA load / store, TMEM epilogue
    ("conv_wgrad_kernel", ("UTCHMMA", "UTMALDG", "LDTM")),             # weight gradient
    ("attention64_kernel", ("UTCHMMA", "UTMALDG", "LDTM", "STTM")),    # flash attention d = 64 (O rescale in TMEM)
    ("attention64x2_kernel", ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "USETMAXREG")),   # two query tiles per CTA, register re-split
    ("attention64h_kernel", ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "MUFU.EX2.F16")),   # eight softmax warps, f16x2 exponentials
    ("attention64t_kernel", ("UTCHMMA tmem", "UTMALDG", "LDTM", "STTM")),   # P.V with the A operand (P) in tensor memory
    ("attention64th_kernel", ("UTCHMMA tmem", "UTMALDG", "LDTM", "STTM")),  # the same with eight softmax warps (production)
    ("attention_kernel", ("UTCHMMA", "UTMALDG", "LDTM")),
    ("attn_bwd_kernel", ("UTCHMMA", "UTMALDG", "LDTM")),               # flash attention backward
])
def test_tensor_core_kernels_use_tcgen05_and_tma(sass_by_function, kernel, needs):
    def is_instance(name):
        if kernel not in name or "temporal" in name:
            return False
        return not (kernel == "attention_kernel" and "attention64" in name)
    bodies = {n: b for n, b in sass_by_function.items() if is_instance(n)}
    assert bodies, f"{kernel} not found in the library"
    for name, body in bodies.items():
        req = [m for m in needs if not (m == "UTMASTG" and "ILi16E" in name)]   # the BN = 16 head variant scatters fp32 with plain stores
        missing = [m for m in req if m not in body]
        assert not missing, f"{name}: SASS lacks {missing}"
        assert "HMMA.16816" not in body, f"{name} contains legacy mma.sync"


def test_fused_groupnorm_gemm_variant_exists(sass_by_function):
    """The XF instantiations (GroupNorm apply on the A operand in shared memory) are in the library next to the plain ones,
    and carry the tanh (SiLU) of the fused out_layers path."""
    xf = {n: b for n, b in sass_by_function.items() if "conv_gemm_kernel" in n and "ELb1E" in n}
    plain = {n: b for n, b in sass_by_function.items() if "conv_gemm_kernel" in n and "ELb0E" in n}
    assert len(xf) >= 5 and len(plain) >= 6, (list(xf), list(plain))
    for name, body in xf.items():
        assert "MUFU.TANH" in body and "UTCHMMA" in body, name
    assert all("MUFU.TANH" not in b for b in plain.values())


@pytest.mark.parametrize("kernel", ["conv_gemm_kernel", "attention64_kernel", "attention64t_kernel", "attention64th_kernel", "attention64h_kernel", "attention64x2_kernel", "attention_kernel",
                                    "conv_wgrad_kernel", "attn_bwd_kernel"])
def test_single_lane_issue_has_no_waterfall_loops(sass_by_function, kernel):
    """TMA / tcgen05 instructions are issued by an elect.sync lane of a warp in uniform control flow.  Issued from an
    `if (lane == 0)` branch instead, ptxas wraps each UTCHMMA / UTCBAR / UTMALDG in an ELECT ... BRA.U.ANY loop over the
    active lanes (measured: the MMA warp then falls behind the tensor pipe) — none may come back."""
    bodies = {n: b for n, b in sass_by_function.items()
              if kernel in n and "temporal" not in n and not (kernel == "attention_kernel" and "attention64" in n)}
    assert bodies, kernel
    for name, body in bodies.items():
        assert "BRA.U.ANY" not in body, f"{name}: per-instruction ELECT / BRA.U.ANY loop is back"
        assert "ELECT" in body, name
