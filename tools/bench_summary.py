"""One-line (or per-family) summary of a bench.py JSON line: python tools/bench_summary.py file.json [label] [--families]"""
import json
import sys

path = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else path
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
except Exception as e:  # noqa: BLE001
    print(f"{label}: no bench line ({e})")
    sys.exit(0)
f = d.get("families", {})
g = lambda k: f.get(k, {}).get("ms", float("nan"))
print(f"{label}: ms/step {d.get('ms_per_step')} e2e {d.get('e2e', {}).get('ms_per_step')} ungraphed {d.get('forward_ms_ungraphed')} "
      f"launches {d.get('launches_per_step')} | gn {g('group_norm'):.3f} 3x3 {g('conv3x3_spatial'):.3f} qkv {g('conv1x1_qkv'):.3f} "
      f"out {g('conv1x1_out'):.3f} proj {g('conv1x1_proj'):.3f} tconv {g('conv_temporal'):.3f} audio {g('conv_audio_k3'):.3f} "
      f"xattn {g('cross_attention'):.3f} sattn {g('self_attention'):.3f} tattn {g('temporal_attention'):.3f} "
      f"clk {d.get('clocks', {}).get('sm_mhz') if d.get('clocks') else None} finite {d.get('finite')}")
if "--families" in sys.argv:
    for k, v in f.items():
        print(f"  {k:20s} {v['ms']:8.3f} ms  {v['launches']:4d} launches  {v['tflops']:8.1f} TF/s  {v['gbs']:8.1f} GB/s")
