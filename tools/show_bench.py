#!/usr/bin/env python
"""Pretty-prints bench.py JSON lines."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e)
        continue
    print(f"== {f}")
    for k in ("impl", "value", "ms_per_step", "layer_tflops_per_gpu", "layer_frac_of_bf16_peak", "gpu_launches", "clocks",
              "cpu_baseline"):
        if k in d:
            print(f"   {k}: {d[k]}")
    if d.get("e2e"):
        print("   e2e:", {k: v for k, v in d["e2e"].items() if k != "how"})
    if d.get("roofline"):
        r = d["roofline"]
        print(f"   roofline: {r['achieved']:.1f} {r['unit']} frac {r['frac']:.3f} ms/launch {r.get('ms_per_launch', 0):.4f} traffic {r['traffic']}")
    for k, v in sorted((d.get("kernels") or {}).items(), key=lambda kv: -kv[1]["ms"]):
        print(f"      {k:20s} {v['ms']:.4f} ms  x{v['calls_per_step']}")
