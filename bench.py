#!/usr/bin/env python
"""Benchmark of the visual-expert decoder hot path (BASELINE.json: prefill tokens/s, % of bf16 tensor peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c4|c2|c4s] [--lora R]

Workloads (BASELINE.json configs):
  c3 (default)  configs[2]: the full 32-layer visual-expert decoder (+ final norm), GLOBAL batch 64 x (1225 vision + 256
                text tokens) sharded by sample over the N ranks (64 / 32 / 16 / 8 samples per GPU at N = 1 / 2 / 4 / 8) --
                "scaling": "strong", `value` = global tokens/s.  This is the configuration north_star's ">= 7x from 1
                to 8 GPUs" is defined on, and it fits one GPU (25.9 GB of weights + ~8 GB of activations).
  c4            configs[3]: same stack, global batch 16 x (2048 vision + 512 text), strong scaling.
  c2            configs[1]: ONE layer, batch 8 x (1225 + 256) per GPU ("weak": every rank runs its own batch).
  c4s           one c4 shard (2 samples per GPU), weak.
One "step" = one prefill forward of the workload's stack over the rank's shard.  Samples are independent on this
path (block-diagonal attention, row-wise everything else): no data-path collective; ranks only meet in the
barrier + max-over-ranks of the timing.  Rank 0 prints ONE JSON line.

  value        : tokens/s with inputs resident in HBM (one CUDA-graph replay per step, CUDA events, max over ranks)
  e2e          : same metric through mmmm_b200.graph.PipelinedHostPrefill with HOST (pinned) inputs, H2D + D2H inside
                 the timed region
  roofline     : dominant kernel (the SwiGLU gate/up grouped GEMM) against the measured bf16 peak
  cpu_baseline : the oracle restatement of the reference layer (fp32 eager PyTorch) on the host cores,
                 bounded sample, rank 0 at N = 1 only
--impl reference times that CPU path as the reference arm (the reference itself cannot travel to the
GPU box: it is eager PyTorch importing packages absent from this image; oracle/oracle_layer.py is
proven bit-identical to it in tests/test_oracle.py); it honours --steps / --warmup.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, I, HEADS = 4096, 11008, 32
WORKLOADS = {
    # layers, samples (global batch if strong scaling, per GPU if weak), vision tokens, text tokens
    "c3": dict(layers=32, batch=64, nv=1225, nt=256, scaling="strong"),   # BASELINE.json configs[2]
    "c4": dict(layers=32, batch=16, nv=2048, nt=512, scaling="strong"),   # configs[3]
    "c2": dict(layers=1, batch=8, nv=1225, nt=256, scaling="weak"),       # configs[1]
    "c4s": dict(layers=1, batch=2, nv=2048, nt=512, scaling="weak"),      # one configs[3] shard
}
GEMM_FLOP_PER_TOKEN = 2 * (H * 3 * H + H * H + 3 * H * I)  # 404 750 336


def workload(args):
    """(layers, samples on this rank's GPU, global samples, nv, nt, scaling, [lo, hi) of the global batch)."""
    from mmmm_b200.sharding import shard_range
    w = WORKLOADS[args.workload]
    world = max(int(os.environ.get("WORLD_SIZE", "1")), 1)
    rank = int(os.environ.get("RANK", "0"))
    layers = args.layers or w["layers"]
    if w["scaling"] == "strong":  # rank g takes the contiguous samples [g*B/G, (g+1)*B/G) (datamodule.py:104-111)
        lo, hi = shard_range(w["batch"], rank, world)
        return layers, hi - lo, w["batch"], w["nv"], w["nt"], "strong", (lo, hi)
    return layers, w["batch"], w["batch"] * world, w["nv"], w["nt"], "weak", (rank * w["batch"], (rank + 1) * w["batch"])


def peaks():
    p = dict(bf16_tflops=1590.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(f):
        try:
            d = json.load(open(f))
            p = dict(bf16_tflops=float(d["bf16_tflops"]), hbm_gbs=float(d["hbm_gbs"]),
                     bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", 0)), source="MEASURED_PEAKS.json")
        except Exception:
            pass
    return p


_RESULT_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout
    when the process group comes up), so file descriptor 1 is pointed at stderr for the rest of the run and the result
    line goes to a private duplicate of the original stdout."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: str):
    out = _RESULT_OUT or sys.stdout
    out.write(line + "\n")
    out.flush()


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md clocks line).  Sampled in-process through
    NVML (pynvml: no start-up delay, so even a 100 ms timed region gets samples); `nvidia-smi -lms` is the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    PERIOD_S = 0.005

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.nvml, self._stop = index, [], None, None, threading.Event()
        self.thread = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        idx = self.index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:  # NVML enumerates physical devices
            try:
                idx = int(vis.split(",")[self.index])
            except ValueError:
                pass
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            # first queries in the calling thread: NVML loads its entry points lazily (hundreds of ms on a fresh box,
            # longer than a 120 ms timed region), so the polling thread must start warm
            self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)
            self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            time.sleep(1.0)  # nvidia-smi needs a moment before its first line
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        n, h = self.nvml, self.handle
        bits = (("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown),
                ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown),
                ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap))
        mx = None
        try:
            mx = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
        except Exception:
            pass
        while not self._stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                try:
                    pw = n.nvmlDeviceGetPowerUsage(h) / 1e3
                except Exception:
                    pw = 0.0
                self.rows.append([str(self.index), sm, mx, pw] + ["Active" if mask & b else "Not Active" for _, b in bits])
            except Exception:
                pass
            self._stop.wait(self.PERIOD_S)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Index of the next sample: bracket the timed region with two marks to select its samples."""
        return len(self.rows)

    def stop(self, lo: int = 0, hi: int | None = None):
        if self.nvml is None and self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvml and nvidia-smi unavailable"])
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
        in_window = hi is None or hi > lo
        rows = self.rows[lo:hi] if in_window else self.rows  # no sample inside the window: fall back to the whole run
        sm, mx, reasons, power = [], None, set(), []
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2]) if r[2] is not None else mx
                power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if str(v).lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=mx, reasons=["no samples"])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                    window="timed region" if in_window else "whole run (no sample fell inside the timed region)",
                    power_w_max=max(power) if power else None,
                    source="nvml" if self.nvml is not None else "nvidia-smi")


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_forward_timer(args, steps: int, warmup: int):
    """fp32 eager forward of ONE sample of the workload through the oracle (reference restatement), all host threads.
    A 32-layer workload is sampled with TWO distinct layers (BASELINE configs[0]'s depth: 32 fp32 layers are 52 GB of
    host memory and ~11 s per step) and scaled to the stack's depth; a single-layer workload runs its one layer.
    Returns (tokens of the sample, per-step seconds FOR THE FULL STACK DEPTH, description)."""
    from oracle import oracle_layer as O
    from mmmm_b200.inputs import make_inputs
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would pin the CPU
    # arm to one core)
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        pass
    w = WORKLOADS[args.workload]
    layers = args.layers or w["layers"]
    n_run = min(layers, 2)
    ws = [O.random_weights(H, I, HEADS, seed=i, dtype=torch.float32) for i in range(n_run)]
    lora = [O.random_lora(H, I, r=args.lora, dtype=torch.float32) for _ in range(n_run)] if args.lora else None
    inp = make_inputs(1, w["nv"], w["nt"], H, seed=0, dtype=torch.float32)
    scale = layers / n_run
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.decoder_stack(ws, inp.hidden_states, inp.token_type_ids, inp.position_ids, inp.padding_mask,
                            num_heads=HEADS, lora=lora)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt * scale)
    what = (f"1 of {w['batch']} samples ({inp.num_valid_tokens} tokens: {w['nv']} vision + {w['nt']} text), fp32 eager "
            f"oracle, {n_run} layer(s) timed" + (f", scaled x{scale:g} to the {layers}-layer stack" if scale != 1 else ""))
    return inp.num_valid_tokens, times, what


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    tokens, times, what = cpu_forward_timer(args, steps, warmup)
    ms = 1e3 * sum(times) / len(times)
    val = tokens / (ms / 1e3)
    cores = torch.get_num_threads()
    layers, b, gb, nv, nt, scaling, _ = workload(args)
    emit(json.dumps({
        "impl": "reference", "metric": "visual-expert prefill tokens/s", "value": val, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": config_dict(args, 0),
        "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": cores, "kind": "port",
                         "sample": what + f"; mean of {steps} steps after {warmup} warm-up(s)"},
        "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = CPU eager forward of the reference decoder (oracle port, bit-identical to the "
                "unmodified reference in tests/test_oracle.py; xformers attention replaced by the equivalent "
                "block-diagonal causal softmax); rank 0 only, one CPU process whatever N is",
    }))


def config_dict(args, tokens_per_gpu):
    layers, b, gb, nv, nt, scaling, _ = workload(args)
    world = max(int(os.environ.get("WORLD_SIZE", "1")), 1)
    batch = (f"global batch {gb} x ({nv} vision + {nt} text tokens) sharded by sample, {b} per GPU" if scaling == "strong"
             else f"batch {b} x ({nv} vision + {nt} text tokens) per GPU")
    return {"workload": f"{args.workload}: {layers} visual-expert layer(s) (hidden {H}, {HEADS} heads, I {I})"
                        f"{' + final RMSNorm' if layers > 1 else ''} bf16 prefill, {batch}",
            "layers": layers, "cuda_graph": bool(args.graph), "global_batch": gb,
            "samples_per_gpu": b, "seq_len": 1 + nv + 2 + 1 + nt, "tokens_per_gpu": tokens_per_gpu,
            "lora_r": args.lora, "parallelism": f"dp{world} (samples sharded, no collective)",
            "residual_stream": "expert-sorted across layers (decoder_stack_forward)" if layers > 1 else "flat [B, L, H] (drop-in layer call)",
            "l2": "per-step working set (0.81 GB weights per layer + >0.5 GB activations) exceeds the 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------------------------- GPU arm
def make_gpu_layer(device, lora_r: int, seed: int = 0, lora_dropout: float = 0.0):
    """Random-init layer with the reference's init (Linear ~ N(0, 0.02)), generated on the device."""
    from mmmm_b200.modeling_cogvlm import CogVLMDecoderLayer, VexConfig
    from mmmm_b200.peft_compat import attach_mock_lora
    with torch.device("meta"):
        layer = CogVLMDecoderLayer(VexConfig(hidden_size=H, intermediate_size=I, num_attention_heads=HEADS))
    layer = layer.to_empty(device=device).to(torch.bfloat16)
    g = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():
        for m in layer.modules():
            if isinstance(m, torch.nn.Linear):
                m.weight.copy_(torch.randn(m.weight.shape, generator=g, device=device, dtype=torch.float32) * 0.02)
        for n in (layer.input_layernorm, layer.post_attention_layernorm):
            n.weight.copy_(1 + 0.1 * torch.randn(H, generator=g, device=device))
        rot = layer.self_attn.rotary_emb
        rot.inv_freq = (1.0 / (rot.base ** (torch.arange(0, 128, 2, device=device) / 128))).to(torch.bfloat16)
    if lora_r:
        torch.manual_seed(seed)
        attach_mock_lora(layer, r=lora_r, lora_alpha=8, b_std=0.02, lora_dropout=lora_dropout)
    return layer.eval()


def ncu_traffic(key: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (not live)."""
    f = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(f)).get(key)
    except Exception:
        return None


def make_norm(device, seed: int = 1234):
    from mmmm_b200.modeling_cogvlm import RMSNorm
    n = RMSNorm(H).to(device)
    g = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():
        n.weight.copy_(1 + 0.1 * torch.randn(H, generator=g, device=device))
    return n


def make_shard_inputs(args):
    """This rank's shard of the synthetic batch (host tensors) and the GLOBAL valid-token count.  The id tensors of
    the whole global batch are built identically on every rank and split with sharding.shard_batch (contiguous
    samples per rank, like the reference's DistributedSamplerWrapper, data/datamodule.py:104-111); hidden states are
    generated per sample (seed = global sample index) for the local samples only."""
    from mmmm_b200.inputs import LayerInputs, make_ids
    from mmmm_b200.sharding import shard_batch
    layers, b, gb, nv, nt, scaling, (lo, hi) = workload(args)
    world = max(int(os.environ.get("WORLD_SIZE", "1")), 1)
    rank = int(os.environ.get("RANK", "0"))
    tt, pos, pm = make_ids(gb, nv, nt)
    global_tokens = int(pm.sum())
    if scaling == "strong":
        tt, pos, pm = shard_batch((tt, pos, pm), rank, world)
    else:
        tt, pos, pm = tt[lo:hi], pos[lo:hi], pm[lo:hi]
    hs = torch.empty(hi - lo, tt.shape[1], H, dtype=torch.bfloat16)
    for i in range(hi - lo):
        g = torch.Generator().manual_seed(1000 + lo + i)
        hs[i] = torch.randn(tt.shape[1], H, generator=g).to(torch.bfloat16)
    return LayerInputs(hs, tt.contiguous(), pos.contiguous(), pm.contiguous()), global_tokens


def run_ours(args):
    import torch.distributed as dist
    from mmmm_b200 import ops  # noqa: F401  (loads libvex.so; raises if it is missing)
    from mmmm_b200 import instrument
    from mmmm_b200._lib import lib
    from mmmm_b200.graph import GraphedPrefill
    from mmmm_b200.modeling_cogvlm import decoder_stack_forward
    from mmmm_b200.plan import GLOBAL_PLAN_CACHE
    from mmmm_b200.sharding import bind_to_gpu_numa_node, max_over_ranks

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else {"numa_node": None}
    if lib().vex_device_check() != 0:
        raise SystemExit("no sm_100 device: the visual-expert kernels only run on B200")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    nl, b, gb, nv, nt, scaling, _ = workload(args)
    stack = nl > 1
    layers = [make_gpu_layer(dev, args.lora, seed=i) for i in range(nl)]
    final_norm = make_norm(dev) if stack else None
    host, total_tokens = make_shard_inputs(args)
    tokens = host.num_valid_tokens  # this rank's
    inp = host.to(dev)

    def forward(hs, tt, pos, pm):
        if stack:  # the full decoder: sorted residual stream across layers + final masked norm
            plan = GLOBAL_PLAN_CACHE.get(tt, pm)
            return decoder_stack_forward(layers, final_norm, hs, plan, pos)[0]
        return layers[0](hs, token_type_ids=tt, position_ids=pos, padding_mask=pm)[0]

    graphed = None
    if args.graph:
        graphed = GraphedPrefill(layers, inp.hidden_states, inp.token_type_ids, inp.position_ids, inp.padding_mask,
                                 final_norm=final_norm)
        step = graphed.replay
    else:
        def step():
            return forward(inp.hidden_states, inp.token_type_ids, inp.position_ids, inp.padding_mask)

    # ---- e2e: HOST (pinned) inputs in, host output back; H2D + D2H inside the timed region.  Two requests are kept
    # in flight the way a serving loop would, so the copies of step i+1 / i-1 overlap the kernels of step i:
    # mmmm_b200.graph.PipelinedHostPrefill (one captured graph per slot) by default; with --no-graph the eager
    # module call on three streams (host-dispatch bound on slow host cores).
    pin = lambda t: t.pin_memory()
    h_host, tt_host, pos_host, pm_host = map(pin, (host.hidden_states, host.token_type_ids, host.position_ids,
                                                   host.padding_mask))
    if args.graph:
        from mmmm_b200.graph import PipelinedHostPrefill
        pipe = PipelinedHostPrefill(layers, h_host, tt_host, pos_host, pm_host, final_norm=final_norm, depth=2,
                                    device=dev)
        out_host = pipe.out_host
        e2e_how = "PipelinedHostPrefill.submit on pinned host inputs; 2 requests in flight (H2D / graph replay / D2H streams)"

        def step_e2e():
            pipe.submit(h_host, tt_host, pos_host, pm_host)
    else:
        out_host = [torch.empty_like(h_host).pin_memory() for _ in range(2)]
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        dev_in = [tuple(torch.empty_like(t, device=dev) for t in (h_host, tt_host, pos_host, pm_host))
                  for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        e2e_state = {"i": 0}
        e2e_how = "forward(...) on pinned host inputs; 2 requests in flight (H2D / compute / D2H streams)"

        def step_e2e():
            i = e2e_state["i"]
            k = i & 1
            cur = torch.cuda.current_stream()
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_free[k])        # compute of step i-2 has consumed this input buffer
                for d, src in zip(dev_in[k], (h_host, tt_host, pos_host, pm_host)):
                    d.copy_(src, non_blocking=True)
                ev_in[k].record(s_in)
            cur.wait_event(ev_in[k])
            out = forward(*dev_in[k])
            ev_free[k].record(cur)
            ev_out[k].record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_out[k])
                out_host[k].copy_(out, non_blocking=True)
            out.record_stream(s_out)
            e2e_state["i"] = i + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    marks = {}

    def timed(fn, steps, warmup, sampler=None):
        with torch.no_grad():
            for _ in range(warmup):
                fn()
            barrier()
            if sampler is not None:
                marks["lo"] = sampler.mark()
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            instrument.reset()
            ranged = sampler is not None and os.environ.get("VEX_PROFILER_RANGE") == "1"
            if ranged:  # `ncu --profile-from-start off`: the launch list covers exactly the timed region
                torch.cuda.cudart().cudaProfilerStart()
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            launches = instrument.launches()
            barrier()
            if ranged:
                torch.cuda.cudart().cudaProfilerStop()
            wall_ms = (time.perf_counter() - t0) * 1e3
            if sampler is not None:
                marks["hi"] = sampler.mark()
        ms = max(e0.elapsed_time(e1), 0.0)
        return ms, wall_ms, launches

    # per-launch device times (CUDA events on the launching stream) of eager steps.  Single layer: taken first, so the
    # dominant kernel is timed alone at burst clocks (denominator: the burst peak).  Stack: the kernel is timed inside
    # a long step (32 layers back to back under the power cap; denominator: the sustained peak).
    kernels = None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # started early: the poller is warm long before the timed region
        with torch.no_grad():
            eager_step = lambda: forward(inp.hidden_states, inp.token_type_ids, inp.position_ids, inp.padding_mask)
            kernels = instrument.profile(eager_step, iters=2 if stack else 5)
    barrier()
    time.sleep(0.5)

    ms_total, _, launches = timed(step, args.steps, args.warmup, sampler if rank == 0 else None)
    clocks = sampler.stop(marks.get("lo", 0), marks.get("hi")) if rank == 0 else None
    ms_step = max_over_ranks(ms_total, dev) / args.steps
    if args.graph:  # launches inside a replayed graph are not visible to the Python counter: count one eager step
        with torch.no_grad():
            GLOBAL_PLAN_CACHE.clear()
            instrument.reset()
            forward(inp.hidden_states, inp.token_type_ids, inp.position_ids, inp.padding_mask)
            launches = instrument.launches() * args.steps
            torch.cuda.synchronize()
    e2e_steps = max(4, args.steps // 2)
    # the e2e region ends when the last D2H has landed: use the wall clock around a full synchronize
    _, wall_e2e, _ = timed(step_e2e, e2e_steps, 3)
    ms_e2e = max_over_ranks(wall_e2e, dev) / e2e_steps

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    sustained = pk.get("bf16_tflops_sustained") or pk["bf16_tflops"]
    value = total_tokens / (ms_step / 1e3)
    seq = 1 + nv + 2 + 1 + nt
    # per-GPU algorithmic work of one step (rank 0's shard; shards are equal for the BASELINE batch sizes)
    flop_gemm = tokens * GEMM_FLOP_PER_TOKEN * nl
    flop_attn = b * 4 * HEADS * 128 * seq * (seq + 1) // 2 * nl
    flop_lora = tokens * 2 * args.lora * 69888 * nl if args.lora else 0
    tf_layer = (flop_gemm + flop_attn + flop_lora) / (ms_step / 1e3) / 1e12
    roof = None
    if kernels and "gemm_swiglu" in kernels:
        gu = kernels["gemm_swiglu"]
        flop = 2.0 * tokens * H * 2 * I + (2.0 * tokens * args.lora * 2 * I if args.lora else 0)
        ms_launch = gu["ms"] / max(gu["calls_per_step"], 1)
        ach = flop / (ms_launch / 1e3) / 1e12
        peak = sustained if stack else pk["bf16_tflops"]
        roof = {"kernel": "k3_grouped_gemm_pair (SwiGLU gate/up)", "bound": "tensor", "achieved": ach,
                "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": ncu_traffic(f"gemm_swiglu_{args.workload}_b{b}") or ncu_traffic(f"gemm_swiglu_{args.workload}"),
                "traffic_source": "profiles/traffic.json (ncu --set full dram__bytes of one launch of this shape; not live)",
                "peak_source": pk["source"] + (" (sustained figure: the kernel is timed inside a 32-layer step)" if stack
                                               else " (burst figure: the kernel is timed alone)"),
                "frac_of_burst": ach / pk["bf16_tflops"], "frac_of_sustained": ach / sustained,
                "flop_per_launch": flop, "ms_per_launch": ms_launch}
    cpu = None
    if world == 1 and not args.no_cpu:
        ctok, ctimes, what = cpu_forward_timer(args, 3, 1)
        best = min(ctimes)
        cpu = {"value": ctok / best, "unit": "tokens/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": what + "; best of 3 after 1 warm-up"}
    h2d = sum(t.numel() * t.element_size() for t in (h_host, tt_host, pos_host, pm_host))
    d2h = out_host[0].numel() * out_host[0].element_size()
    cfg = config_dict(args, tokens)
    emit(json.dumps({
        "metric": "visual-expert prefill tokens/s", "value": value, "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": cfg,
        "tokens_per_s_per_gpu": value / world, "ms_per_layer": ms_step / nl,
        "layer_tflops_per_gpu": tf_layer, "layer_frac_of_bf16_peak": tf_layer / pk["bf16_tflops"],
        "layer_frac_of_bf16_sustained": tf_layer / sustained,
        "roofline": roof, "kernels": kernels, "cpu_baseline": cpu,
        "e2e": {"value": total_tokens / (ms_e2e / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e,
                "how": e2e_how},
        "gpu_launches": launches, "clocks": clocks, "peaks": pk, "host_numa": numa,
    }))
    if world > 1:
        dist.destroy_process_group()


def run_train(args):
    """BASELINE config 5: LoRA (r = --lora, default 64) forward + backward of an N-layer visual-expert decoder with
    the LoRA-gradient all-reduce (NCCL) -- one step = fwd + self-checkpointed bwd + all-reduce over one batch.
    Forward, recompute and backward all run on the native kernels (mmmm_b200/training.py); the all-reduce is issued
    per layer from inside the backward on a side stream (training.BucketedGradReducer), so only the last layer's
    collective is exposed.  --allreduce post times the round-1 post-backward reducer instead (A/B)."""
    import torch.distributed as dist
    from mmmm_b200 import instrument
    from mmmm_b200.sharding import max_over_ranks
    from mmmm_b200.training import BucketedGradReducer, LoraGradReducer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    r = args.lora or 64
    n_layers, b, gb, nv, nt, scaling, _ = workload(args)
    args.layers = n_layers
    layers = [make_gpu_layer(dev, r, seed=i, lora_dropout=args.lora_dropout).train() for i in range(args.layers)]
    for l in layers:
        l.recompute = "auto" if args.recompute < 0 else bool(args.recompute)
    params = [p for l in layers for p in l.parameters() if p.requires_grad]
    overlapped = args.allreduce == "overlap"
    reducer = (BucketedGradReducer(layers, layers_per_collective=args.ar_group, tail_layers=args.ar_tail) if overlapped
               else LoraGradReducer(params))
    host, total_tokens = make_shard_inputs(args)
    inp = host.to(dev)
    tokens = int(inp.padding_mask.sum())
    proj = torch.randn_like(inp.hidden_states)
    ar_ms = []

    def step():
        h = inp.hidden_states.detach().requires_grad_(True)
        x = h
        if overlapped:
            reducer.zero()
        for layer in layers:
            x = layer(x, token_type_ids=inp.token_type_ids, position_ids=inp.position_ids,
                      padding_mask=inp.padding_mask)[0]
        loss = (x.float() * proj.float()).mean()
        if not overlapped:
            for p in params:
                p.grad = None
        loss.backward()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if overlapped:
            reducer.finish()   # joins the side stream: what is left of the last layer's collective
        else:
            reducer.reduce()
        e1.record()
        ar_ms.append((e0, e1))
        return loss

    for _ in range(args.warmup):
        step()
    kernels = instrument.profile(step, iters=2)  # every rank: the step contains the all-reduce (a collective)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ar_ms.clear()
    instrument.reset()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lo_mark = sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    launches = instrument.launches()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop(lo_mark, sampler.mark()) if rank == 0 else None
    ms_step = max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    allreduce_ms = sum(a.elapsed_time(b_) for a, b_ in ar_ms) / max(len(ar_ms), 1)
    # cost of the collective inside the step: the same step with the process group's collectives skipped
    ms_nocomm = None
    if world > 1 and overlapped:
        saved = reducer._world
        reducer._world = lambda: 1
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        dist.barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            step()
        f1.record()
        torch.cuda.synchronize()
        ms_nocomm = max_over_ranks(f0.elapsed_time(f1), dev) / args.steps
        reducer._world = saved
    if rank == 0:
        total = total_tokens
        seq = 1 + nv + 2 + 1 + nt
        # algorithmic FLOP of one step: forward + recompute (up to the down projection) + input-gradient GEMMs
        # (base weights frozen) + LoRA forward/dgrad/wgrad + attention forward x2 and backward (5 GEMM-equivalents
        # of the causal score matrix per direction; the two-kernel backward executes 7)
        g_fwd = tokens * GEMM_FLOP_PER_TOKEN
        g_down = tokens * 2 * H * I
        lora_f = tokens * 2 * r * 69888
        attn_f = b * 4 * HEADS * 128 * seq * (seq + 1) // 2
        from mmmm_b200.training import KEEP_BUDGET
        asked = KEEP_BUDGET.granted + KEEP_BUDGET.refused
        # fraction of layer forwards that checkpointed (auto: the ones the memory budget refused to keep)
        rec = (KEEP_BUDGET.refused / asked if asked else 1.0) if args.recompute < 0 else (1 if args.recompute else 0)
        flop = args.layers * ((g_fwd + lora_f) + rec * (g_fwd - g_down + lora_f) + (g_fwd + lora_f) + 2 * lora_f
                              + (1 + rec) * attn_f + 2.5 * attn_f)
        pk = peaks()
        exposed = None if ms_nocomm is None else ms_step - ms_nocomm
        busbw = None
        if world > 1:  # ring all-reduce moves 2 (N-1)/N of the buffer per rank
            t = (exposed if exposed and exposed > 0 else allreduce_ms) / 1e3
            busbw = reducer.nbytes * 2 * (world - 1) / world / max(t, 1e-9) / 1e9
        emit(json.dumps({
            "metric": "visual-expert LoRA train tokens/s", "value": total / (ms_step / 1e3), "unit": "tokens/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(config_dict(args, tokens), lora_r=r, lora_dropout=args.lora_dropout,
                           recompute=("auto" if args.recompute < 0 else bool(args.recompute)),
                           recomputed_layer_fraction=rec,
                           mode=("train: fwd + recompute + bwd + LoRA-grad allreduce" if rec == 1 else
                                 "train: fwd (activations kept in HBM) + bwd + LoRA-grad allreduce" if rec == 0 else
                                 "train: fwd (activations kept while they fit) + partial recompute + bwd + LoRA-grad "
                                 "allreduce"),
                           allreduce=((f"bucket views, NCCL AVG per {args.ar_group} layer(s) (the group of layer 0: "
                                       f"{args.ar_tail or args.ar_group}) on a side stream during the backward"
                                       if args.ar_group else "bucket views, ONE NCCL AVG over the flat "
                                       "bucket after the backward") if overlapped else "post-backward, packed (round 1)")),
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
            "allreduce_tail_ms": allreduce_ms, "allreduce_bytes": reducer.nbytes,
            "ms_per_step_without_collectives": ms_nocomm, "allreduce_exposed_ms": exposed,
            "allreduce_busbw_gbs_over_exposed_time": busbw,
            "trainable_params": (reducer.acc if overlapped else reducer.flat).numel(),
            "gpu_launches": launches, "clocks": clocks, "kernels": kernels,
            "step_tflops_per_gpu": flop / (ms_step / 1e3) / 1e12,
            "step_frac_of_bf16_peak": flop / (ms_step / 1e3) / 1e12 / pk["bf16_tflops"],
            "step_frac_of_bf16_sustained": flop / (ms_step / 1e3) / 1e12 / (pk.get("bf16_tflops_sustained") or pk["bf16_tflops"]),
            "peaks": pk,
            "note": "forward, recompute and backward all run on the native sm_100a kernels (K1-K9); only the "
                    "LoRA-gradient all-reduce is NCCL",
        }))
    if world > 1:
        dist.destroy_process_group()


def run_decode(args):
    """generation after a prefill.  One step = ONE new token for every sample of the batch through the
    full stack (q_len == 1: language-expert weights only, K / V appended in place by the QKV epilogue, K4d over the
    pre-allocated cache, final norm), replayed as one CUDA graph per token (kv_cache.StaticKVCache).  HBM-bound: per
    step every language-expert weight (404.75 MB per layer) and the live K / V of every sample are read once."""
    import torch.distributed as dist
    from types import SimpleNamespace
    from mmmm_b200 import instrument
    from mmmm_b200.kv_cache import StaticKVCache
    from mmmm_b200.modeling_cogvlm import decoder_stack_forward
    from mmmm_b200.plan import GLOBAL_PLAN_CACHE
    from mmmm_b200.sharding import max_over_ranks

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nl, b, gb, nv, nt, scaling, _ = workload(args)
    layers = [make_gpu_layer(dev, args.lora, seed=i) for i in range(nl)]
    model = SimpleNamespace(layers=layers, norm=make_norm(dev))
    host, _ = make_shard_inputs(args)
    inp = host.to(dev)
    L = inp.padding_mask.shape[1]
    n_new = args.warmup + args.steps + 8
    cache = StaticKVCache(nl, b, HEADS, L + n_new, dev)
    with torch.no_grad():
        plan = GLOBAL_PLAN_CACHE.get(inp.token_type_ids, inp.padding_mask)
        decoder_stack_forward(layers, model.norm, inp.hidden_states, plan, inp.position_ids, use_cache=True,
                              kv_out=cache.layers)
    cache.start(inp.padding_mask)
    x = torch.randn(b, 1, H, device=dev).to(torch.bfloat16)
    x_host, out_host = x.cpu().pin_memory(), torch.empty(b, 1, H, dtype=torch.bfloat16).pin_memory()
    pos = inp.position_ids.max(dim=1, keepdim=True).values + 1
    state = {"i": 0}

    def step():
        out = cache.step(model, x, pos + state["i"], graph=bool(args.graph))
        state["i"] += 1
        return out

    def step_e2e():  # host token embedding in, host hidden state out, every step (what a sampling loop on the host does)
        x.copy_(x_host, non_blocking=True)
        out_host.copy_(step(), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    kernels = None
    with torch.no_grad():
        instrument.reset()
        cache.step(model, x, pos, graph=False)       # one eager step: launch count + per-kernel times
        launches_per_step = instrument.launches()
        state["i"] = 1
        kernels = instrument.profile(lambda: (cache.step(model, x, pos + 1, graph=False), None)[1], iters=2) if rank == 0 else None
        state["i"] = 4 if rank == 0 else 1
        cache.host_len = L + state["i"]
        cache.past_len.fill_(cache.host_len)
    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lo = sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_timed = args.steps - 2
    for _ in range(n_timed):
        step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop(lo, sampler.mark()) if rank == 0 else None
    ms_step = max_over_ranks(e0.elapsed_time(e1), dev) / n_timed
    t0 = time.perf_counter()
    for _ in range(2):
        step_e2e()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3, dev) / 2
    if rank == 0:
        pk = peaks()
        kv_len = L + args.warmup + args.steps // 2
        w_bytes = nl * 2 * (H * 3 * H + H * H + 3 * H * I)           # language expert, bf16
        kv_bytes = nl * 2 * b * kv_len * H * 2                        # K and V of every sample, once per layer
        gbs = (w_bytes + kv_bytes) / (ms_step / 1e3) / 1e9
        total = b * world
        emit(json.dumps({
            "metric": "visual-expert decode tokens/s", "value": total / (ms_step / 1e3), "unit": "tokens/s",
            "n_gpus": world, "steps": n_timed, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"decode after a {args.workload} prefill: {nl} layers, batch {b} per GPU, ~{kv_len} cached "
                                   f"positions per sample, 1 new token per sample per step",
                       "layers": nl, "samples_per_gpu": b, "cuda_graph": bool(args.graph), "lora_r": args.lora,
                       "l2": "per-step weights (12.95 GB) exceed the 126 MB L2; no flush needed"},
            "roofline": {"kernel": "whole decode step (weight + KV streaming)", "bound": "hbm", "achieved": gbs,
                         "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"], "traffic": None,
                         "bytes_per_step": w_bytes + kv_bytes, "weight_bytes": w_bytes, "kv_bytes": kv_bytes,
                         "peak_source": pk["source"]},
            "kernels": kernels,
            "e2e": {"value": total / (ms_e2e / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": b * H * 2,
                    "d2h_bytes_per_step": b * H * 2, "ms_per_step": ms_e2e,
                    "how": "host embedding in, graph replay, host hidden state out, synchronised every token"},
            "gpu_launches": launches_per_step * n_timed, "launches_per_step": launches_per_step, "clocks": clocks,
            "peaks": pk,
        }))
    if world > 1:
        dist.destroy_process_group()


def run_vision(args):
    """the EVA2-CLIP-E vision encoder in front of the decoder (63 layers of 1792 = 16 x 112, MLP 15360,
    GLU projector to 4096 / 11008) over a batch of 490 x 490 images (35 x 35 patches of 14 + class token = 1226
    tokens per image, the BASELINE vision-token count), one step = EVA2CLIPModel.forward over the batch."""
    import torch.distributed as dist
    from types import SimpleNamespace
    from mmmm_b200 import instrument
    from mmmm_b200.sharding import max_over_ranks
    from mmmm_b200.visual import EVA2CLIPModel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    C, heads, Iv, nl, side, patch = 1792, 16, 15360, args.vision_layers, 490, 14
    b = workload(args)[1]
    vc = dict(hidden_size=C, num_heads=heads, intermediate_size=Iv, num_hidden_layers=nl, layer_norm_eps=1e-6,
              in_channels=3, patch_size=(1, patch, patch), pos_embed_shape=(1, side // patch, side // patch),
              hidden_act="gelu")
    with torch.device("meta"):
        model = EVA2CLIPModel(SimpleNamespace(hidden_size=H, intermediate_size=I, vision_config=vc))
    model = model.to_empty(device=dev).to(torch.bfloat16).eval()
    g = torch.Generator(device=dev).manual_seed(0)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("layernorm.weight") or n.endswith("norm1.weight"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g, device=dev))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g, device=dev))
    images = [torch.randn(3, 1, side, side, generator=g, device=dev).to(torch.bfloat16) for _ in range(b)]
    ps, pool = [(1, patch, patch)] * b, [(1, 1, 1)] * b
    host_images = [im.cpu().pin_memory() for im in images]

    def step():
        return model(images, ps, pool)

    def step_e2e():
        out = model([h.to(dev, non_blocking=True) for h in host_images], ps, pool)
        return [o.cpu() for o in out]

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step()
        kernels = instrument.profile(step, iters=3) if rank == 0 else None
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        instrument.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        launches = instrument.launches()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = sampler.stop() if rank == 0 else None
        ms_step = max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
        step_e2e()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_e2e = max(2, args.steps // 4)
        for _ in range(n_e2e):
            step_e2e()
        torch.cuda.synchronize()
        ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3, dev) / n_e2e
    if rank == 0:
        L = 1 + (side // patch) ** 2
        tokens = b * L
        hd = C // heads
        flop_layer = tokens * 2 * (C * 3 * C + C * C + 2 * C * Iv) + b * 4 * heads * hd * L * L
        flop_glu = b * (L - 1) * 2 * (C * H + 3 * H * I)
        flop_patch = b * (L - 1) * 2 * 3 * patch * patch * C
        flop = nl * flop_layer + flop_glu + flop_patch
        pk = peaks()
        tf = flop / (ms_step / 1e3) / 1e12
        roof = None
        if kernels and "gemm_plain_gelu" in kernels:
            k = kernels["gemm_plain_gelu"]
            ms_launch = k["ms"] / max(k["calls_per_step"], 1)
            ach = 2.0 * tokens * C * Iv / (ms_launch / 1e3) / 1e12
            roof = {"kernel": "k3_grouped_gemm_pair (fc1 + bias + GELU epilogue)", "bound": "tensor", "achieved": ach,
                    "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"], "traffic": None,
                    "peak_source": pk["source"] + " (burst figure)", "ms_per_launch": ms_launch}
        total = tokens * world
        emit(json.dumps({
            "metric": "vision-encoder prefill tokens/s", "value": total / (ms_step / 1e3), "unit": "tokens/s",
            "images_per_s": b * world / (ms_step / 1e3), "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"EVA2-CLIP-E vision encoder ({nl} layers, hidden {C} = {heads} x {hd}, MLP {Iv}) + GLU "
                                   f"projector ({H} / {I}), batch {b} x {side}x{side} image ({L} tokens) per GPU",
                       "vision_layers": nl, "images_per_gpu": b, "tokens_per_gpu": tokens,
                       "parallelism": f"dp{world} (images sharded, no collective)",
                       "l2": "per-step weights (136 MB per layer) and activations exceed the 126 MB L2; no flush needed"},
            "encoder_tflops_per_gpu": tf, "encoder_frac_of_bf16_peak": tf / pk["bf16_tflops"],
            "roofline": roof, "kernels": kernels,
            "e2e": {"value": total / (ms_e2e / 1e3), "unit": "tokens/s",
                    "h2d_bytes_per_step": sum(h.numel() * 2 for h in host_images),
                    "d2h_bytes_per_step": b * (L + 1) * H * 2, "ms_per_step": ms_e2e,
                    "how": "model(images) with pinned host images, features copied back to the host"},
            "gpu_launches": launches, "clocks": clocks, "peaks": pk,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--lora", type=int, default=0, help="LoRA rank on all ten Linears (0 = frozen weights only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--layers", type=int, default=0, help="decoder layers per step (default: the workload's: 32 for c3 / c4, 1 for c2)")
    ap.add_argument("--train", action="store_true", help="config 5: LoRA fwd+bwd training step + grad all-reduce")
    ap.add_argument("--lora-dropout", type=float, default=0.0, help="--train: lora_dropout (the reference uses 0.05)")
    ap.add_argument("--recompute", type=int, default=-1, help="--train: 1 = checkpoint each layer like the reference "
                    "(save the input, recompute in backward); 0 = keep the activations in HBM (no recompute pass); "
                    "-1 (default, the product default) = keep them while they fit the memory budget, checkpoint the rest")
    ap.add_argument("--allreduce", default="overlap", choices=["overlap", "post"], help="--train: bucket-view reducer "
                    "(default; gradients accumulate straight into the flat bucket) or the round-1 pack / unpack reducer")
    ap.add_argument("--ar-tail", type=int, default=2, help="--train: layers in the group that holds layer 0 (its "
                    "collective is the one the backward cannot hide); 0 = like the other groups")
    ap.add_argument("--ar-group", type=int, default=8, help="--train: layers per NCCL collective, issued on a side stream "
                    "during the backward (default 8; 0 = ONE collective over the whole bucket after the backward)")
    ap.add_argument("--graph", type=int, default=1, help="1: replay the forward as one CUDA graph (default), 0: eager")
    ap.add_argument("--decode", action="store_true", help="SURVEY 8(f)-2: time graphed decode steps after a prefill")
    ap.add_argument("--vision", action="store_true", help="SURVEY 8(f)-4: time the EVA2-CLIP-E vision encoder instead")
    ap.add_argument("--vision-layers", type=int, default=63)
    args = ap.parse_args()
    claim_stdout()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.decode:
        run_decode(args)
    elif args.vision:
        run_vision(args)
    elif args.train:
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
