#!/usr/bin/env python
"""Timing experiments for the two-tile attention kernel.  `build` (here) compiles experiment libraries
mmmm_b200/libvex_trace_e<E>.so with -DVEX_ATTN_TRACE -DA2_EMU=<E> (cycle stamps of one CTA's tile-A / tile-B softmax
warp 0 and of the MMA issuer; E of every 8 exp2 pairs on the FMA pipe); `run` (GPU box) times the c2 attention with each
library and schedule variant and prints the per-block phase durations of the heaviest CTA."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
EMUS = (0, 1, 2, 3, 4, 5)


def lib_path(e):
    return os.path.join(ROOT, "mmmm_b200", f"libvex_trace_e{e}.so")


def build():
    from mmmm_b200 import build as b
    objdir = os.path.join(b.HERE, "build_trace")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in b.CFLAGS if f not in ("-Xptxas", "-v")]
    objs = []
    for src in b.sources():
        if src == "k4_attention_tc2.cu":
            continue
        obj = os.path.join(objdir, src[:-3] + ".o")
        if not os.path.isfile(obj) or os.path.getmtime(obj) < os.path.getmtime(os.path.join(b.CSRC, src)):
            subprocess.run([b.NVCC, *b.ARCH_FLAGS, *flags, "-c", os.path.join(b.CSRC, src), "-o", obj], check=True)
        objs.append(obj)
    for e in EMUS:
        obj = os.path.join(objdir, f"k4_attention_tc2_e{e}.o")
        subprocess.run([b.NVCC, *b.ARCH_FLAGS, *flags, "-DVEX_ATTN_TRACE", f"-DA2_EMU={e}", "-c",
                        os.path.join(b.CSRC, "k4_attention_tc2.cu"), "-o", obj], check=True)
        subprocess.run([b.NVCC, *b.ARCH_FLAGS, "-shared", "-Xcompiler", "-fPIC", "-o", lib_path(e), *objs, obj,
                        "-cudart", "static"], check=True)
        print(lib_path(e))


def run_one(e, variants, verbose):
    os.environ["VEX_LIB_PATH"] = lib_path(e)
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
    buf = torch.zeros(4 * 64 * 8, dtype=torch.int64, device="cuda")
    Lb = lib()
    Lb.vex_debug_attn_trace.argtypes = [ctypes.c_void_p]
    os.environ["VEX_ATTN_IMPL"] = "tc2"
    call = lambda: ops.attention(qkv, plan.cu_seqlens, B, L, heads, plan.token_to_sorted, out, 128 ** -0.5)
    ref = None
    for variant in variants:
        os.environ["VEX_ATTN_P"] = variant
        assert Lb.vex_debug_attn_trace(0) == 0
        ms = timeit(call)
        if ref is None and e == 0:
            ref = out.clone()
        buf.zero_()
        assert Lb.vex_debug_attn_trace(buf.data_ptr()) == 0
        call()
        torch.cuda.synchronize()
        t = buf.cpu().view(4, 64, 8)
        mid = []
        for role in (0, 1):
            rows = [[int(v) for v in t[role, j, :6]] for j in range(2, 9)]
            d = [[r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[5] - r[0]] for r in rows]
            mid.append([sum(c) // len(c) for c in zip(*d)])
        print(f"emu={e} {variant:6s} {ms * 1e3:7.1f} us | A: wait_S {mid[0][0]:5d} ld+max {mid[0][1]:5d} exp {mid[0][2]:5d} "
              f"wait_pv {mid[0][3]:5d} store {mid[0][4]:5d} block {mid[0][5]:5d} | B block {mid[1][5]:5d}", flush=True)
        if verbose and variant in ("early", "token"):
            t0 = int(t[0, 0, 0])
            print("   j: issuer A top | s_free wait, S issue, PV (p_full wait + issue) || issuer B ... || softmax A start, B start")
            for j in range(1, 10):
                ra = [int(v) for v in t[2, j]]
                rb = [int(v) for v in t[3, j]]
                print(f"   {j:2d} @ {ra[0] - t0:6d} | {ra[1] - ra[0]:5d} {ra[2] - ra[1]:5d} {ra[3] - ra[2]:5d} || @ {rb[0] - t0:6d} | "
                      f"{rb[1] - rb[0]:5d} {rb[2] - rb[1]:5d} {rb[3] - rb[2]:5d} || A @ {int(t[0, j, 0]) - t0:6d}  B @ {int(t[1, j, 0]) - t0:6d}")

if __name__ == "__main__":
    if sys.argv[1:] == ["build"]:
        build()
    elif sys.argv[1] == "one":
        run_one(int(sys.argv[2]), sys.argv[3].split(","), len(sys.argv) > 4)
    else:
        for e in EMUS:
            subprocess.run([sys.executable, __file__, "one", str(e), "early,tmem"])
