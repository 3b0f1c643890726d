#!/usr/bin/env python
"""Per-kernel timing on the GPU box (CUDA events, L2-sized working sets): attention implementations and the
grouped GEMM at the BASELINE config-2 shapes.  Prints one JSON line per measurement."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmmm_b200 import ops  # noqa: E402
from mmmm_b200.inputs import make_ids  # noqa: E402
from mmmm_b200.plan import build_plan  # noqa: E402


def timeit(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_attention(B=8, nv=1225, nt=256, heads=32, only=None):
    tt, pos, pm = make_ids(B, nv, nt)
    plan = build_plan(tt.cuda(), pm.cuda())
    L = tt.shape[1]
    cap = B * L
    qkv = torch.randn(cap, 3 * heads * 128, device="cuda").bfloat16()
    out = torch.empty(cap, heads * 128, device="cuda", dtype=torch.bfloat16)
    flop = B * 4 * heads * 128 * L * (L + 1) / 2
    res = {}
    impls = only or ("mma", "tc1", "tc2-tmem", "tc2", "tc3")
    for impl in impls:
        os.environ["VEX_ATTN_P"] = impl.split("-")[1] if "-" in impl else "early"
        out.zero_()
        if impl == "tc3":
            attend = ops.attention
        else:  # superseded kernels: libvex_baselines.so (csrc/baselines/), not part of the product library
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "helpers"))
            import baselines
            attend = lambda *a, _i=impl.split("-")[0]: baselines.attention(_i, *a)
        try:
            ms = timeit(lambda: attend(qkv, plan.cu_seqlens, B, L, heads, plan.token_to_sorted, out, 128 ** -0.5))
        except Exception as e:  # keep measuring the other implementations
            print(json.dumps({"kernel": f"attention_{impl}", "error": str(e)[:200]}))
            continue
        res[impl] = out.clone()
        print(json.dumps({"kernel": f"attention_{impl}", "B": B, "L": L, "ms": ms, "tflops": flop / ms / 1e9}))
    for impl in impls:
        if impl != "mma" and impl in res and "mma" in res:
            d = (res["mma"].float() - res[impl].float()).abs().max().item()
            print(json.dumps({f"attention_mma_vs_{impl}_maxabs": d}))
    os.environ.pop("VEX_ATTN_P", None)


def bench_gemm(Tv=9808, Tl=2072):
    H, I = 4096, 11008
    cap = Tv + Tl
    counts = torch.tensor([Tv, Tl, cap, 0], dtype=torch.int32, device="cuda")
    mk = lambda *s: (torch.randn(*s, device="cuda") * 0.02).bfloat16()
    x = mk(cap, H)
    for name, N, K, mode in (("qkv", 3 * H, H, ops.EPI_PLAIN), ("dense", H, H, ops.EPI_PLAIN),
                             ("gateup_swiglu", I, H, ops.EPI_SWIGLU), ("down", H, I, ops.EPI_PLAIN)):
        a = x if K == H else mk(cap, K)
        w = [mk(N, K), mk(N, K) if mode == ops.EPI_SWIGLU else None, mk(N, K), mk(N, K) if mode == ops.EPI_SWIGLU else None]
        out = torch.empty(cap, N, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: ops.grouped_gemm_fused(a, w, out, counts, mode, None, None, [None, None], [None] * 4, 0,
                                                   [], 0, False, 1.0))
        flop = 2.0 * cap * N * K * (2 if mode == ops.EPI_SWIGLU else 1)
        ref = timeit(lambda: torch.matmul(a, w[0].T))
        print(json.dumps({"kernel": f"gemm_{name}", "M": cap, "N": N, "K": K, "ms": ms, "tflops": flop / ms / 1e9,
                          "cublas_single_expert_ms": ref, "cublas_tflops": 2.0 * cap * N * K / ref / 1e9}))


def bench_rowwise(T=11880, H=4096, I=11008):
    """HBM-bound kernels at config-2 sizes: achieved GB/s = algorithmic bytes / time."""
    dev = "cuda"
    n = torch.tensor([T], dtype=torch.int32, device=dev)
    perm = torch.randperm(T, device=dev).int()
    x = torch.randn(T, H, device=dev).bfloat16()
    w = torch.ones(H, device=dev, dtype=torch.bfloat16)
    y = torch.empty_like(x)
    big = [torch.randn(T, I, device=dev).bfloat16() for _ in range(5)]
    dw = torch.zeros(H, device=dev, dtype=torch.float32)
    cases = {
        "k2_rmsnorm_gather": (lambda: ops.rmsnorm_gather(x, w, 1e-6, perm, n, y), 2 * T * H * 2),
        "k5_silu_mul": (lambda: ops.silu_mul(big[0], big[1], n, big[2]), 3 * T * I * 2),
        "k6_residual_scatter": (lambda: ops.residual_scatter(x, y, perm, n, y), 3 * T * H * 2),
        "k7_gather_rows": (lambda: ops.gather_rows(x, perm, n, y), 2 * T * H * 2),
        "k7_silu_mul_backward": (lambda: ops.silu_mul_backward(big[0], big[1], big[2], n, big[3], big[4]), 5 * T * I * 2),
        "k7_rmsnorm_backward": (lambda: ops.rmsnorm_backward(x, x, perm, w, 1e-6, x, None, y, perm, dw, n), 4 * T * H * 2),
    }
    for name, (fn, nbytes) in cases.items():
        ms = timeit(fn, iters=50)
        print(json.dumps({"kernel": name, "rows": T, "ms": ms, "GBps": nbytes / ms / 1e6}))


def bench_lm_head(B=8, L=1485, H=4096, V=32008, frac=0.17):
    """Fused lm_head + weighted CE (forward + backward to the hidden states) against the reference's way of
    computing it -- fp32 logits for every position, then F.cross_entropy on the masked rows -- in eager PyTorch on
    the same GPU (a comparator, not a fallback)."""
    import torch.nn.functional as F
    from mmmm_b200.lm_head import fused_lm_head_loss
    g = torch.Generator(device="cuda").manual_seed(0)
    h = torch.randn(B, L, H, device="cuda", generator=g).bfloat16()
    lin = torch.nn.Linear(H, V, bias=False, device="cuda", dtype=torch.bfloat16)
    lin.weight.requires_grad_(False)
    labels = torch.randint(0, V, (B, L), device="cuda", generator=g)
    labels[torch.rand(B, L, device="cuda", generator=g) >= frac] = -100
    wt = (0.5 + torch.rand(B, L, device="cuda", generator=g)).bfloat16()
    n_lab = int((labels != -100).sum())

    def ours():
        x = h.detach().requires_grad_(True)
        fused_lm_head_loss(x, lin, labels, wt).backward()
        return x.grad

    def eager():
        x = h.detach().requires_grad_(True)
        # eager PyTorch the way the reference computes it (modeling_cogvlm.py:610-627, 701-706); the oracle package is
        # test infrastructure and is not imported by tools
        logits = F.linear(x, lin.weight).float().view(-1, V)
        lab = labels.view(-1)
        mask = lab != -100
        ce = F.cross_entropy(logits, lab, reduction="none")
        (torch.dot(ce[mask], wt.float().view(-1)[mask]) / mask.sum()).backward()
        return x.grad

    ms_o = timeit(ours, iters=10, warmup=3)
    torch.cuda.reset_peak_memory_stats()
    ours()
    mem_o = torch.cuda.max_memory_allocated()
    ms_e = timeit(eager, iters=5, warmup=2)
    torch.cuda.reset_peak_memory_stats()
    eager()
    mem_e = torch.cuda.max_memory_allocated()
    flop = 3 * 2.0 * n_lab * H * V  # logits forward, logits recompute, dgrad over the labelled rows
    print(json.dumps({"kernel": "lm_head_ce_fwd_bwd", "positions": B * L, "labelled": n_lab, "V": V, "ms": ms_o,
                      "tflops_labelled_rows": flop / ms_o / 1e9, "peak_mem_GB": mem_o / 1e9,
                      "eager_all_positions_ms": ms_e, "eager_peak_mem_GB": mem_e / 1e9, "speedup": ms_e / ms_o}))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "attention"):
        bench_attention()
        bench_attention(B=2, nv=2048, nt=512)
    if which.startswith("attention:"):  # one implementation (ncu captures): attention:tc2, attention:tc2-smem, ...
        bench_attention(only=(which.split(":", 1)[1],))
    if which in ("all", "gemm"):
        bench_gemm()
    if which in ("all", "rowwise"):
        bench_rowwise()
    if which in ("all", "lmhead"):
        bench_lm_head()
