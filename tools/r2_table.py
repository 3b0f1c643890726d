#!/usr/bin/env python
"""Builds the 'Measured on B200' table of DESIGN.md / README.md from the bench lines under profiles/ (round 2)."""
import json
import os
import sys

P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")


def load(name):
    f = os.path.join(P, name)
    if not os.path.isfile(f):
        return None
    try:
        return json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        return None


def k(v):
    return f"{v / 1e3:.1f} k" if v < 1e6 else f"{v / 1e6:.2f} M"


rows = []


def prefill(name, label):
    d = load(name)
    if not d:
        return
    r = d.get("roofline") or {}
    c = d.get("clocks") or {}
    rows.append((label, f"{k(d['value'])} tok/s ({d['ms_per_step']:.2f} ms/step, {d.get('ms_per_layer', d['ms_per_step']):.2f} ms/layer); "
                        f"e2e {k(d['e2e']['value'])} tok/s",
                 f"{100 * d['layer_frac_of_bf16_peak']:.1f} % of burst / {100 * d.get('layer_frac_of_bf16_sustained', 0):.1f} % of sustained; "
                 f"SwiGLU GEMM {r.get('achieved', 0):.0f} TF/s = {100 * r.get('frac', 0):.1f} % of its peak; SM {c.get('sm_mhz')} MHz",
                 name))


prefill("r2_bench_c3_n1.json", "**c3** (default): 32 layers + final norm, global batch 64 x 1485 tokens, N = 1")
prefill("r2_bench_c3_n1_final.json", "c3, N = 1 again on the final code of the round (another box: 1.07 GHz under the cap)")
prefill("r2_bench_c3_n2.json", "c3, N = 2 (32 samples per GPU)")
prefill("r2_bench_c3_n4.json", "c3, N = 4 (16 samples per GPU)")
prefill("r2_bench_c3_n8.json", "c3, N = 8 (8 samples per GPU)")
prefill("r2_bench_c3_n8_final.json", "c3, N = 8 again on the final code of the round (another box: 1.17 GHz under the cap)")
prefill("r2_bench_c4_n1.json", "c4: 32 layers, global batch 16 x 2564 tokens, N = 1")
prefill("r2_bench_c4_n1_final.json", "c4, N = 1 again on the final code of the round")
prefill("r2_bench_c4_n8.json", "c4, N = 8 (2 samples per GPU)")
prefill("r2_bench_c4_n8_final.json", "c4, N = 8 again on the final code of the round")
prefill("r2_bench_c2.json", "c2: ONE layer, 8 x 1485 tokens, N = 1 (burst clocks)")
prefill("r2_bench_c2_lora64.json", "c2 + LoRA r = 64 on all ten Linears")
prefill("r2_bench_stack32_c2.json", "32-layer stack over 8 samples (= the c3 shard of one GPU at N = 8), N = 1")
d = load("r2_bench_decode.json")
if d:
    r = d["roofline"]
    rows.append(("decode step after a c2 prefill: 32 layers, batch 8, ~1500 cached positions, CUDA graph",
                 f"{d['value']:.0f} tok/s ({d['ms_per_step']:.2f} ms per token step, {d['launches_per_step']} launches); e2e {d['e2e']['value']:.0f} tok/s",
                 f"HBM-bound: {r['achieved']:.0f} GB/s = {100 * r['frac']:.1f} % of the measured {r['peak']:.0f} GB/s "
                 f"({r['weight_bytes'] / 1e9:.2f} GB weights + {r['kv_bytes'] / 1e9:.2f} GB K/V per step)", "r2_bench_decode.json"))
for name, label in (("r2_bench_train32_auto_n1.json", "**c5**: LoRA r = 64 training step, 32 layers, 8 x 1485 tokens, N = 1, default (`recompute = 'auto'`: every layer's activations fit and are kept in HBM)"),
                    ("r2_bench_train32_auto_n8.json", "c5, default mode, N = 8 (bucketed LoRA-grad all-reduce overlapped with the backward)"),
                    ("r2_bench_train32_n1.json", "c5, every layer checkpointed like the reference (`--recompute 1`), N = 1"),
                    ("r2_bench_train32_partial_n1.json", "c5, auto with the budget forced down to 40 GB (`VEX_TRAIN_KEEP_RESERVE_GB=140`): 7 layers kept, 25 checkpointed"),
                    ("r2_bench_train32_n2.json", "c5, N = 2, checkpointed"), ("r2_bench_train32_n8.json", "c5, N = 8, checkpointed")):
    d = load(name)
    if d:
        ex = d.get("allreduce_exposed_ms")
        rows.append((label, f"{k(d['value'])} tok/s ({d['ms_per_step']:.1f} ms/step = {d['ms_per_step'] / d['config']['layers']:.2f} ms/layer)"
                            + (f"; all-reduce of {d['allreduce_bytes'] / 1e6:.0f} MB: {ex:+.1f} ms vs the same step without collectives" if ex is not None else ""),
                     f"{100 * d['step_frac_of_bf16_peak']:.1f} % of burst / {100 * d.get('step_frac_of_bf16_sustained', 0):.1f} % of sustained (algorithmic fwd + recompute + bwd FLOP); peak memory {d['peak_mem_gb']:.1f} GB", name))
d = load("r2_bench_ref.json")
if d:
    rows.append(("reference arm / CPU baseline: the oracle port, fp32 eager, 1 sample through 2 layers scaled to 32",
                 f"{d['value']:.0f} tok/s on {d['cpu_baseline']['cores']} host threads", "—", "r2_bench_ref.json"))
print("| workload | result | of measured peak | bench line (`profiles/`) |\n|---|---|---|---|")
for r in rows:
    print("| " + " | ".join(r) + " |")
