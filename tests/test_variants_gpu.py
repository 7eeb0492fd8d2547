"""GPU: the switchable kernel variants stay correct.  The library reads its environment switches once per process, so each
variant runs the operator / forward parity tests in a subprocess with the switch set."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env,select", [
    ({"MMD_ATTN_TMEM": "1"}, "attention or forward_small"),             # P in tensor memory, four softmax warps (attention64t_kernel)
    ({"MMD_ATTN_TMEM": "0"}, "attention or forward_small"),             # probabilities through shared memory (attention64_kernel)
    ({"MMD_ATTN_TMEM": "0", "MMD_ATTN_SPLIT": "1"}, "attention or forward_small"),   # eight softmax warps, f16x2 exponentials
    ({"MMD_ATTN_PAIR": "1"}, "attention or forward_small"),             # two query tiles per CTA at every d = 64 site
    ({"MMD_ATTN_PAIR": "0"}, "attention or forward_small or production"),  # ... at none
    ({"MMD_XF": "7"}, "forward_small or forward_production"),           # GroupNorm apply on the GEMM A operand
    ({"MMD_EG": "0", "MMD_MT": "0"}, "conv or forward_small"),          # one epilogue warpgroup, 128-token tiles only
    ({"MMD_WSTORE": "1"}, "conv or forward_small or production"),       # per-warp output stores for token-matrix GEMMs
    ({"MMD_NO_GRAPH": "1", "MMD_NO_PDL": "1"}, "forward_small"),        # eager launches without programmatic dependent launch
])
def test_variant_passes_the_parity_tests(env, select):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_ops_gpu.py", "tests/test_forward_gpu.py", "-m", "gpu", "-q", "-x",
                        "-p", "no:cacheprovider", "-k", select], cwd=ROOT, env=e, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, f"{env}: {tail}"
    assert " passed" in r.stdout and "no tests ran" not in r.stdout, tail
