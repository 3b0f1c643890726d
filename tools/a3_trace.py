#!/usr/bin/env python
"""Timing experiment for the persistent attention kernel: `build` (here) compiles mmmm_b200/libvex_trace_a3.so with
-DVEX_ATTN_TRACE (cycle stamps of tile A's softmax warp 0 in CTA 0, trace pointer held in a register); `run` (GPU box)
runs the c2 attention and prints the phase durations of the first key blocks / work items of that CTA."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "mmmm_b200", "libvex_trace_a3.so")


def build():
    from mmmm_b200 import build as b
    objdir = os.path.join(b.HERE, "build_trace")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in b.CFLAGS if f not in ("-Xptxas", "-v")]
    objs = []
    for src in b.sources():
        obj = os.path.join(objdir, src[:-3] + ".o")
        extra = ["-DVEX_ATTN_TRACE"] if src == "k4_attention_tc3.cu" else []
        subprocess.run([b.NVCC, *b.ARCH_FLAGS, *flags, *extra, "-c", os.path.join(b.CSRC, src), "-o", obj], check=True)
        objs.append(obj)
    subprocess.run([b.NVCC, *b.ARCH_FLAGS, "-shared", "-Xcompiler", "-fPIC", "-o", LIB, *objs, "-cudart", "static"],
                   check=True)
    print(LIB)


def run():
    os.environ["VEX_LIB_PATH"] = LIB
    os.environ["VEX_ATTN_IMPL"] = "tc3"
    import torch
    from mmmm_b200 import ops
    from mmmm_b200._lib import lib
    from mmmm_b200.plan import build_plan
    from tools.bench_kernels import make_ids, timeit
    B, heads = 8, 32
    tt, pos, pm = make_ids(B, 1225, 256)
    plan = build_plan(tt.cuda(), pm.cuda())
    L = tt.shape[1]
    cap = B * L
    qkv = torch.randn(cap, 3 * heads * 128, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(cap, heads * 128, device="cuda", dtype=torch.bfloat16)
    call = lambda: ops.attention(qkv, plan.cu_seqlens, B, L, heads, plan.token_to_sorted, out, 128 ** -0.5)
    Lb = lib()
    Lb.vex_debug_a3_trace.argtypes = [ctypes.c_void_p]
    assert Lb.vex_debug_a3_trace(0) == 0
    print(f"attention tc3: {timeit(call) * 1e3:.1f} us")
    buf = torch.zeros(128 * 8, dtype=torch.int64, device="cuda")
    assert Lb.vex_debug_a3_trace(buf.data_ptr()) == 0
    call()
    torch.cuda.synchronize()
    t = buf.cpu().view(128, 8)
    t0 = int(t[0, 0])
    print(" blk @start | wait_S  ld+s_free  max  exp  wait_pv  store+arrive | total   (E = last block of an item; gap to "
          "the next block = wait PV + epilogue + next item's first wait)")
    for c in range(40):
        r = [int(v) for v in t[c, :8]]
        if r[6] == 0:
            break
        nxt = int(t[c + 1, 0]) if c + 1 < 128 else 0
        tail = f"  E gap {nxt - r[6]:6d}" if r[7] and nxt else ""
        print(f"  {c:3d} @ {r[0] - t0:7d} | " + " ".join(f"{r[i + 1] - r[i]:6d}" for i in range(6)) + f" | {r[6] - r[0]:6d}{tail}")


if __name__ == "__main__":
    build() if sys.argv[1:] == ["build"] else run()
