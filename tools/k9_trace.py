#!/usr/bin/env python
"""Timing experiment for the attention-backward dK/dV kernel: `build` (here) compiles mmmm_b200/libvex_trace_k9.so with
-DVEX_ATTN_TRACE (cycle stamps of softmax warp 0 of CTA (0,0,0): the key block with the most query steps); `run` (GPU box)
runs forward + backward at c2 shapes and prints the per-step phase durations."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "mmmm_b200", "libvex_trace_k9.so")


def build():
    from mmmm_b200 import build as b
    objdir = os.path.join(b.HERE, "build_trace")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in b.CFLAGS if f not in ("-Xptxas", "-v")]
    objs = []
    for src in b.sources():
        obj = os.path.join(objdir, src[:-3] + ".o")
        extra = ["-DVEX_ATTN_TRACE"] if src == "k9_attention_bwd.cu" else []
        subprocess.run([b.NVCC, *b.ARCH_FLAGS, *flags, *extra, "-c", os.path.join(b.CSRC, src), "-o", obj], check=True)
        objs.append(obj)
    subprocess.run([b.NVCC, *b.ARCH_FLAGS, "-shared", "-Xcompiler", "-fPIC", "-o", LIB, *objs, "-cudart", "static"],
                   check=True)
    print(LIB)


def run():
    os.environ["VEX_LIB_PATH"] = LIB
    import torch
    from mmmm_b200 import ops
    from mmmm_b200._lib import lib
    from tools.bench_kernels import timeit
    B, heads, L = 8, 32, 1485
    H = heads * 128
    cap = B * L
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(cap, 3 * H, generator=g).bfloat16().cuda()
    d_tok = torch.randn(cap, H, generator=g).bfloat16().cuda()
    cu = (torch.arange(B + 1) * L).int().cuda()
    ident = torch.arange(cap).int().cuda()
    pos = torch.arange(L).repeat(B).cuda()
    cos = torch.randn(2048, 128).bfloat16().cuda()
    sin = torch.randn(2048, 128).bfloat16().cuda()
    out = torch.zeros(cap, H, dtype=torch.bfloat16).cuda()
    lse = torch.zeros(heads, cap, dtype=torch.float32).cuda()
    ops.attention_train(qkv, cu, B, L, heads, ident, out, 128 ** -0.5, lse)
    dqkv = torch.zeros(cap, 3 * H, dtype=torch.bfloat16).cuda()
    delta = torch.empty(heads, cap, dtype=torch.float32).cuda()
    call = lambda: ops.attention_backward(qkv, out, d_tok, lse, delta, cu, ident, ident, pos, cos, sin, B, L, heads, dqkv,
                                          128 ** -0.5)
    Lb = lib()
    Lb.vex_debug_k9_trace.argtypes = [ctypes.c_void_p]
    assert Lb.vex_debug_k9_trace(0) == 0
    print(f"attention backward (delta + dkdv + dq): {timeit(call) * 1e3:.1f} us")
    if os.environ.get("VEX_K9_IMPL", "persistent") != "grid":
        return run_persistent(Lb, call)
    buf = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
    assert Lb.vex_debug_k9_trace(buf.data_ptr()) == 0
    call()
    torch.cuda.synchronize()
    t = buf.cpu().view(64, 8)
    t0 = int(t[63, 0])
    print(f"CTA (0,0,0): prologue to first step {int(t[63, 1]) - t0}, steps end @ {int(t[63, 2]) - t0}, "
          f"acc_full wait {int(t[63, 3]) - int(t[63, 2])}, epilogue {int(t[63, 4]) - int(t[63, 3])}, "
          f"total {int(t[63, 4]) - t0}")
    print(" step @start | stat+bar  wait_S  ld+compute  wait_p_empty  store+arrive | total")
    for s in range(24):
        r = [int(v) for v in t[s, :6]]
        if r[5] == 0:
            break
        print(f"  {s:2d} @ {r[0] - t0:7d} | " + " ".join(f"{r[i + 1] - r[i]:7d}" for i in range(5)) + f" | {r[5] - r[0]:6d}")


def run_persistent(Lb, call):
    """Stamps of CTA 0 of the persistent dQ kernel (k9_attn_bwd_dq_p): softmax warp 4 (group 0: even global steps), the MMA
    warp and the producer thread, per global step."""
    import torch
    buf = torch.zeros(256 * 16, dtype=torch.int64, device="cuda")
    assert Lb.vex_debug_k9_trace(buf.data_ptr()) == 0
    call()
    torch.cuda.synchronize()
    t = buf.cpu().view(256, 16)
    rows = [g for g in range(256) if int(t[g, 8]) or int(t[g, 0])]
    t0 = min(int(v) for v in t[rows][:, [0, 8]].flatten() if int(v))
    print("dQ persistent, CTA 0.  softmax (group 0, even steps): wait_S  pass0  pass1  st+wait  arrive | MMA: issueS  wait_P  issue_dQ | producer slot-free @")
    for g in rows[:80]:
        r = [int(v) for v in t[g]]
        sm = "   ".join(f"{r[i + 1] - r[i]:6d}" for i in range(5)) if r[0] else " " * 42
        mma = "   ".join(f"{r[i + 1] - r[i]:6d}" for i in range(8, 11)) if r[8] else ""
        print(f"  g={g:3d} sm@{(r[0] - t0) if r[0] else 0:8d} | {sm} | mma@{(r[8] - t0) if r[8] else 0:8d} {mma} | prod@{(r[12] - t0) if r[12] else 0:8d}"
              + (f"  ITEM END, epilogue {r[7] - r[6]}" if r[6] and r[7] else ""))


if __name__ == "__main__":
    build() if sys.argv[1:] == ["build"] else run()
