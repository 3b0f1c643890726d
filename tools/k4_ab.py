#!/usr/bin/env python
"""A/B timing of the K4 attention kernel: `build` (here) compiles mmmm_b200/libvex_k4_old.so with k4_attention_tc3.cu taken
from a git revision (default HEAD); `run` (GPU box) times ops.attention at the c2 and c4-shard shapes under the current
libvex.so and the old one in alternating sub-processes (min / median of CUDA-event timings, L2 flushed between calls)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OLD = os.path.join(ROOT, "mmmm_b200", "libvex_k4_old.so")


def build(rev="HEAD"):
    from mmmm_b200 import build as b
    objdir = os.path.join(b.HERE, "build_k4_old")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in b.CFLAGS if f not in ("-Xptxas", "-v")]
    tmp = os.path.join(objdir, "old_k4_attention_tc3.cu")
    with open(tmp, "w") as f:
        f.write(subprocess.run(["git", "show", f"{rev}:mmmm_b200/csrc/k4_attention_tc3.cu"], cwd=ROOT, check=True,
                               capture_output=True, text=True).stdout)
    try:
        objs = []
        for src in b.sources():
            obj = os.path.join(objdir, src[:-3] + ".o")
            path = tmp if src == "k4_attention_tc3.cu" else os.path.join(b.CSRC, src)
            subprocess.run([b.NVCC, *b.ARCH_FLAGS, *flags, "-I", b.CSRC, "-c", path, "-o", obj], check=True)
            objs.append(obj)
        subprocess.run([b.NVCC, *b.ARCH_FLAGS, "-shared", "-Xcompiler", "-fPIC", "-o", OLD, *objs, "-cudart", "static"],
                       check=True)
    finally:
        os.remove(tmp)
    print(OLD)


def one():
    import torch
    from mmmm_b200 import ops
    from mmmm_b200.plan import build_plan
    from tools.bench_kernels import make_ids
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, B, nv, nt in (("c2", 8, 1225, 256), ("c4x2", 2, 2304, 256)):
        heads = 32
        tt, pos, pm = make_ids(B, nv, nt)
        plan = build_plan(tt.cuda(), pm.cuda())
        L = tt.shape[1]
        cap = B * L
        qkv = torch.randn(cap, 3 * heads * 128, device="cuda", dtype=torch.bfloat16)
        out = torch.empty(cap, heads * 128, device="cuda", dtype=torch.bfloat16)
        call = lambda: ops.attention(qkv, plan.cu_seqlens, B, L, heads, plan.token_to_sorted, out, 128 ** -0.5)
        for _ in range(5):
            call()
        ts = []
        for _ in range(40):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); call(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        print(f"   {name}: min {ts[0]:.1f} us  median {ts[len(ts) // 2]:.1f} us", flush=True)


def run():
    for rep in range(3):
        for tag, libpath in (("new", None), ("old", OLD)):
            env = dict(os.environ)
            if libpath:
                env["VEX_LIB_PATH"] = libpath
            print(tag, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, check=True)


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "run"
    if cmd == "build":
        build(*sys.argv[2:])
    elif cmd == "one":
        one()
    else:
        run()
