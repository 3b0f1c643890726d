#!/usr/bin/env python
"""Where does the end-to-end (host buffers) step go?  PCIe H2D / D2H / both-way rates for the c2 hidden-state buffer,
then the double-buffered e2e loop with the eager module call and with two CUDA-graph instances."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mmmm_b200.graph import GraphedPrefill  # noqa: E402
from mmmm_b200.inputs import make_inputs  # noqa: E402


def wall(fn, n):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3


def main():
    dev = torch.device("cuda", 0)
    b, nv, nt = bench.WORKLOADS["c2"]
    host = make_inputs(b, nv, nt, bench.H, seed=0)
    h = host.hidden_states.pin_memory()
    o = torch.empty_like(h).pin_memory()
    d_in, d_out = torch.empty_like(h, device=dev), torch.randn(h.shape, device=dev).to(h.dtype)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    nbytes = h.numel() * h.element_size()

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            o.copy_(d_out, non_blocking=True)

    for f in (h2d, d2h):
        f()
    r = {"bytes": nbytes, "h2d_gbs": nbytes / wall(h2d, 10) / 1e6, "d2h_gbs": nbytes / wall(d2h, 10) / 1e6,
         "both_ms": wall(lambda: (h2d(), d2h()), 10)}
    r["both_gbs_each"] = nbytes / r["both_ms"] / 1e6
    print(json.dumps(r))

    layer = bench.make_gpu_layer(dev, 0)
    ids = [t.pin_memory() for t in (host.token_type_ids, host.position_ids, host.padding_mask)]
    inp = host.to(dev)
    graphs = [GraphedPrefill([layer], inp.hidden_states, inp.token_type_ids, inp.position_ids, inp.padding_mask)
              for _ in range(2)]
    outs = [torch.empty_like(h).pin_memory() for _ in range(2)]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    st = {"i": 0}

    def step_graph():
        i = st["i"]
        k = i & 1
        g = graphs[k]
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(s_in):
            if i >= 2:
                s_in.wait_event(ev_free[k])
            g.hidden_states.copy_(h, non_blocking=True)
            g.token_type_ids.copy_(ids[0], non_blocking=True)
            g.position_ids.copy_(ids[1], non_blocking=True)
            g.padding_mask.copy_(ids[2], non_blocking=True)
            ev_in[k].record(s_in)
        cur.wait_event(ev_in[k])
        if i >= 2:
            cur.wait_event(ev_copied[k])  # the graph's output buffer of step i-2 has been read out
        g.replay()
        ev_free[k].record(cur)
        ev_done[k].record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done[k])
            outs[k].copy_(g.output, non_blocking=True)
            ev_copied[k].record(s_out)
        st["i"] = i + 1

    with torch.no_grad():
        for _ in range(4):
            step_graph()
        ms = wall(step_graph, 20)
    print(json.dumps({"e2e_graph_ms": ms, "tok_s": host.num_valid_tokens / ms * 1e3}))
    with torch.no_grad():
        ms1 = wall(graphs[0].replay, 20)
    print(json.dumps({"graph_only_ms": ms1}))


if __name__ == "__main__":
    main()
