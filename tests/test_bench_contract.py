"""bench.py output contract, checked on the arm that runs without a GPU (--impl reference: the CPU restatement of the
reference layer on a bounded sample): exactly ONE line on stdout, valid JSON, the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["metric"] == "visual-expert prefill tokens/s" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_strong_scaling_shards_cover_the_global_batch(monkeypatch):
    """bench.py's default workload (c3) splits a FIXED global batch of 64 samples by contiguous sample ranges
    (sharding.shard_batch, the reference's DistributedSamplerWrapper, data/datamodule.py:104-111): at every N the shards
    partition the batch, the per-rank token counts add up to the global count, and rank r's ids equal the r-th slice
    of the global id tensors."""
    import argparse
    import importlib.util
    import torch
    spec = importlib.util.spec_from_file_location("vex_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from mmmm_b200.inputs import make_ids
    monkeypatch.setattr(bench, "H", 8)   # tiny hidden size: only the sharding logic is under test
    for wl, gb, seq in (("c3", 64, 1485), ("c4", 16, 2564)):
        args = argparse.Namespace(workload=wl, layers=0, lora=0, graph=1)
        tt_g, pos_g, pm_g = make_ids(gb, bench.WORKLOADS[wl]["nv"], bench.WORKLOADS[wl]["nt"])
        for world in (1, 2, 4, 8):
            total, seen = 0, 0
            for rank in range(world):
                monkeypatch.setenv("WORLD_SIZE", str(world))
                monkeypatch.setenv("RANK", str(rank))
                layers, b, gbatch, nv, nt, scaling, (lo, hi) = bench.workload(args)
                assert (layers, gbatch, scaling, b) == (32, gb, "strong", gb // world) and hi - lo == b
                inp, global_tokens = bench.make_shard_inputs(args)
                assert inp.hidden_states.shape == (b, seq, 8) and global_tokens == int(pm_g.sum())
                assert torch.equal(inp.token_type_ids, tt_g[lo:hi]) and torch.equal(inp.padding_mask, pm_g[lo:hi])
                assert torch.equal(inp.position_ids, pos_g[lo:hi])
                total += inp.num_valid_tokens
                seen += b
                cfg = bench.config_dict(args, inp.num_valid_tokens)
                assert cfg["global_batch"] == gb and cfg["samples_per_gpu"] == b and "sharded by sample" in cfg["workload"]
            assert seen == gb and total == int(pm_g.sum())
    # the single-layer workload stays per-GPU ("weak")
    monkeypatch.setenv("WORLD_SIZE", "4")
    monkeypatch.setenv("RANK", "3")
    layers, b, gbatch, *_rest = bench.workload(argparse.Namespace(workload="c2", layers=0, lora=0, graph=1))
    assert (layers, b, gbatch, _rest[2]) == (1, 8, 32, "weak")
