"""Training-step variant of the visual-expert layer (BASELINE config 5): LoRA forward + backward.

Forward = the fused sm_100a path (``visual_expert_layer_forward``).  The layer checkpoints itself the way the
reference trains (non-reentrant gradient checkpointing is always on, mmmm/models/mmmm.py:287-291): only the
layer input is saved; the backward re-runs the native forward up to the down projection keeping the
intermediates, then propagates with the native backward kernels

    d_out -> gather to sorted rows (K7) -> down_proj dgrad (+LoRA) [K3, MN-major B] -> SwiGLU backward (K7)
          -> gate/up dgrad (+LoRA, accumulated) -> RMSNorm backward + residual add (K7) -> dense dgrad (+LoRA,
          scattered to token order) -> attention backward + rotary adjoint (K9) -> QKV dgrad (+LoRA)
          -> RMSNorm backward + residual add, scattered to [B, L] (K7)

and the ten LoRA weight gradients with K8 (tcgen05, reduction over tokens).  Gradients are produced for the layer
input, every active ``lora_A`` / ``lora_B`` and the (modules_to_save) RMSNorm weights; the base Linear weights are
frozen, as under PEFT.  No host synchronisation anywhere: expert row counts stay on the device.

What autograd would do on the reference is restated here per op; ``tests/test_layer_gpu.py`` checks every gradient
against torch.autograd over the oracle layer.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .peft_compat import LinearSpec

HEAD_DIM = 128


def _bf16(t: torch.Tensor) -> torch.Tensor:
    from .modeling_cogvlm import _bf16 as cast
    return cast(t)


def _base_weight(t: torch.Tensor) -> torch.Tensor:
    from .modeling_cogvlm import _base_weight as chk
    return chk(t)


class _GradSink:
    """fp32 accumulators for the trainable tensors of one layer backward (K8 / K7 add into them atomically).  With a
    ``BucketedGradReducer`` attached to the layer the accumulators ARE that layer's segment of the reducer's flat fp32
    buffer (no per-step allocation, no pack copy); otherwise they are fresh zero tensors."""

    def __init__(self, reducer=None):
        self.items: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.reducer = reducer

    def buffer(self, p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if p is None or not p.requires_grad:
            return None
        if id(p) not in self.items:
            buf = self.reducer.accumulator(p) if self.reducer is not None else None
            if buf is None:
                buf = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
            self.items[id(p)] = (p, buf)
        return self.items[id(p)][1]

    def result(self):
        if self.reducer is not None:  # the reducer owns the gradients (it installs p.grad views itself)
            return [(p, g if g.dtype == p.dtype else g.to(p.dtype)) for p, g in self.items.values()
                    if not self.reducer.owns(p)]
        return [(p, g if g.dtype == p.dtype else g.to(p.dtype)) for p, g in self.items.values()]


def _routed_linear_backward(dy: torch.Tensor, x: torch.Tensor, t: Optional[torch.Tensor],
                            pair: Tuple[LinearSpec, LinearSpec], counts: torch.Tensor, out: torch.Tensor,
                            sink: _GradSink, *, accumulate: bool = False, row_map: Optional[torch.Tensor] = None,
                            x_dropped: Optional[torch.Tensor] = None, dropout_seed: int = 0):
    """Backward of the routed (vision / language) Linear + LoRA: y = x W_e^T + T B_e^T with T = s drop(x) A_e^T.
        dx = dy W_e + mask/(1-p) * (dT A_e)   (dT = s dy B_e)        dB_e = dy^T T        dA_e = dT^T drop(x)
    ``dy`` / ``x`` / ``t`` are expert-sorted [cap, *]; ``out`` receives dx (through ``row_map`` / accumulated).
    Without dropout the LoRA term rides in the main GEMM as a K-extension; with dropout (``x_dropped`` = the
    forward's dropped copy of x) it is a second, mask-applying accumulate pass (VEX_EPI_DROPOUT_ACC)."""
    sv, sl = pair
    both = sl.lora_A is not None
    p_drop = sv.dropout if sv.lora_A is not None else 0.0
    dt, r, lora_a = None, 0, [None, None]
    if sv.lora_A is not None:
        r = sv.r
        dt = torch.empty(dy.shape[0], r, dtype=torch.bfloat16, device=dy.device)
        # dT = s * dy . B  (B stored [out, r]: the small-N transposed GEMM; language rows stay unused without an adapter)
        ops.grouped_gemm_dgrad(dy, [_bf16(sv.lora_B), _bf16(sl.lora_B) if both else None], dt, counts, False, None, None,
                               [None, None], 0, not both, float(sv.scaling))
        lora_a = [_bf16(sv.lora_A), _bf16(sl.lora_A) if both else None]
        gB = (sink.buffer(sv.lora_B), sink.buffer(sl.lora_B) if both else None)
        gA = (sink.buffer(sv.lora_A), sink.buffer(sl.lora_A) if both else None)
        if gB[0] is not None or gB[1] is not None:
            ops.lora_wgrad(dy, t, gB[0], gB[1], False, counts)
        if gA[0] is not None or gA[1] is not None:
            ops.lora_wgrad(x_dropped if p_drop > 0 else x, dt, gA[0], gA[1], True, counts)
    if p_drop > 0:
        ops.grouped_gemm_dgrad(dy, [_base_weight(sv.weight), _base_weight(sl.weight)], out, counts, accumulate, row_map, None,
                               [None, None], 0, False, 1.0)
        ops.grouped_gemm_dgrad(dt, lora_a, out, counts, True, row_map, None, [None, None], 0, not both, 1.0,
                               p_drop, dropout_seed)
    else:
        ops.grouped_gemm_dgrad(dy, [_base_weight(sv.weight), _base_weight(sl.weight)], out, counts, accumulate, row_map, dt, lora_a,
                               r, False, 1.0)


def layer_backward(layer, plan, position_ids: torch.Tensor, hidden_states: torch.Tensor, d_out: torch.Tensor,
                   dropout_seed: Optional[int] = None, keep: Optional[Dict] = None):
    """Returns (d_hidden [B, L, H], [(param, grad), ...]) -- see the module docstring.  ``dropout_seed``: the seed the
    forward used for the LoRA dropout masks (the recompute and the dgrad epilogues regenerate them from it).
    ``keep``: the intermediates of an activation-keeping forward (``layer.recompute = False``); None = recompute them
    from ``hidden_states`` (the reference's checkpointing granularity)."""
    from .modeling_cogvlm import dropout_stream_seed, visual_expert_layer_forward
    B, L, H = hidden_states.shape
    cap = B * L
    attn, mlp = layer.self_attn, layer.mlp
    heads = attn.num_heads
    I = mlp.vision_mlp.intermediate_size
    dev = hidden_states.device
    if keep is None:
        keep = {}
        with torch.no_grad():
            visual_expert_layer_forward(layer, hidden_states, plan, position_ids, keep=keep, dropout_seed=dropout_seed)
    specs = keep["specs"]
    seed = keep.get("dropout_seed") or 0
    drop = lambda nm, k: dict(x_dropped=keep.get(nm + "_xd"), dropout_seed=dropout_stream_seed(seed, k))
    counts, s2f, n_valid = plan.counts, plan.sorted_to_flat, plan.n_valid
    new = lambda *shape: torch.empty(*shape, dtype=torch.bfloat16, device=dev)
    reducer = getattr(layer, "_vex_grad_reducer", None)
    sink = _GradSink(reducer)
    hf = hidden_states.view(cap, H)
    dof = d_out.view(cap, H)

    dy = new(cap, H)                                                   # d_out in expert-sorted order
    ops.gather_rows(dof, s2f, n_valid, dy)
    # ---- MLP block ----
    dact = new(cap, I)
    _routed_linear_backward(dy, keep["act"], keep["t_down"], specs["down"], counts, dact, sink, **drop("down", 4))
    dg, du = new(cap, I), new(cap, I)
    ops.silu_mul_backward(dact, keep["gate"], keep["up"], n_valid, dg, du)
    del dact
    dxn2 = new(cap, H)
    _routed_linear_backward(dg, keep["xn2"], keep["t_gate"], specs["gate"], counts, dxn2, sink, **drop("gate", 2))
    _routed_linear_backward(du, keep["xn2"], keep["t_up"], specs["up"], counts, dxn2, sink, accumulate=True,
                            **drop("up", 3))
    del dg, du
    ln1, ln2 = keep["ln1"], keep["ln2"]
    dh1 = new(cap, H)                                                  # grad w.r.t. h1 rows (sorted): dy + norm branch
    ops.rmsnorm_backward(dxn2, keep["h1"].view(cap, H), s2f, ln2.weight.detach(), ln2.variance_epsilon, dy, None, dh1,
                         None, sink.buffer(ln2.weight), n_valid)
    # ---- attention block ----
    dctx_tok = new(cap, H)                                             # token order (K9 zeroes its tail rows)
    _routed_linear_backward(dh1, keep["ctx"], keep["t_dense"], specs["dense"], counts, dctx_tok, sink,
                            row_map=plan.sorted_to_token, **drop("dense", 1))
    dqkv = new(cap, 3 * H)                                             # d(pre-rotary q | k | v), sorted order
    delta = torch.empty(heads, cap, dtype=torch.float32, device=dev)
    ops.attention_backward(keep["qkv"], keep["ctx"], dctx_tok, keep["lse"], delta, plan.cu_seqlens,
                           plan.token_to_sorted, plan.token_to_flat, position_ids.reshape(-1), keep["cos"], keep["sin"],
                           B, L, heads, dqkv, HEAD_DIM ** -0.5)
    dxn1 = dxn2                                                        # reuse
    _routed_linear_backward(dqkv, keep["xn1"], keep["t_qkv"], specs["qkv"], counts, dxn1, sink, **drop("qkv", 0))
    d_hidden = torch.empty_like(d_out)
    ops.copy_padded_rows(dof, plan.flat_to_sorted, d_hidden.view(cap, H))   # padded rows: identity path only
    ops.rmsnorm_backward(dxn1, hf, s2f, ln1.weight.detach(), ln1.variance_epsilon, dh1, None, d_hidden.view(cap, H), s2f,
                         sink.buffer(ln1.weight), n_valid)
    if reducer is not None:  # this layer's gradients are complete: cast its segment, maybe hand a layer group to NCCL
        reducer.layer_done(layer)
    return d_hidden, sink.result()


# --------------------------------------------------------------------------------------------------
# autograd plumbing
# --------------------------------------------------------------------------------------------------
def trainable_tensors(layer) -> List[torch.Tensor]:
    """The parameters of the layer that currently require grad (LoRA A/B, norm copies), in a fixed order."""
    return [p for p in layer.parameters() if p.requires_grad]


class _LayerFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layer, plan, position_ids, hidden_states, *trainables):
        from .modeling_cogvlm import next_dropout_seed, visual_expert_layer_forward
        ctx.dropout_seed = next_dropout_seed()  # the backward's recompute must regenerate the same LoRA dropout masks
        ctx.keep = None
        mode = recompute_default(layer)
        if mode == "auto":
            rows = hidden_states.shape[0] * hidden_states.shape[1]
            token = KEEP_BUDGET.take(hidden_states.device, keep_bytes_estimate(layer, rows, hidden_states.shape[2]))
            if token is not None:
                ctx.keep = {"continue_forward": True, "_budget": token}
        elif not mode:
            # B200 has the HBM to keep one layer's intermediates (~1.5 GB per 8 x 1485 tokens, 47 GB for 32 layers):
            # the backward then starts from them instead of re-running the forward
            ctx.keep = {"continue_forward": True}
        with torch.no_grad():
            out, _ = visual_expert_layer_forward(layer, hidden_states, plan, position_ids, keep=ctx.keep,
                                                 dropout_seed=ctx.dropout_seed)
        ctx.layer, ctx.plan, ctx.position_ids = layer, plan, position_ids
        ctx.trainables = trainables
        ctx.save_for_backward(hidden_states)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (hidden_states,) = ctx.saved_tensors
        with torch.no_grad():
            d_hidden, pg = layer_backward(ctx.layer, ctx.plan, ctx.position_ids, hidden_states, d_out.contiguous(),
                                          ctx.dropout_seed, keep=ctx.keep)
            ctx.keep = None  # release the kept activations as soon as the layer's backward is done
        by_id = {id(p): g for p, g in pg}
        tr = tuple(by_id.get(id(p)) for p in ctx.trainables)
        return (None, None, None, d_hidden if ctx.needs_input_grad[3] else None, *tr)


def recompute_default(layer):
    """How a layer's training forward treats its intermediates: ``True`` -- save only the layer input and recompute in
    the backward, like the reference's always-on gradient checkpointing (mmmm/models/mmmm.py:287-291); ``False`` -- keep
    them in HBM (no recompute pass, ~1.3x faster step); ``"auto"`` (default) -- keep them while they fit (`KeepBudget`).
    ``layer.recompute`` overrides the environment (VEX_TRAIN_RECOMPUTE = 1 | 0 | auto)."""
    import os
    flag = getattr(layer, "recompute", None)
    if flag is None:
        flag = {"1": True, "0": False}.get(os.environ.get("VEX_TRAIN_RECOMPUTE", "auto"), "auto")
    return flag if flag == "auto" else bool(flag)


def keep_bytes_estimate(layer, rows: int, hidden: int) -> int:
    """Upper bound of what an activation-keeping forward leaves in HBM for the backward: per token the normed input,
    qkv (3H), the attention context, h1 and its norm, the LoRA T rows (< H), gate / up / their product (3I), bf16."""
    try:
        inter = _base_weight(layer.mlp.language_mlp.down_proj).shape[1]
    except Exception:  # a layer without the reference's module tree: assume the 7B ratio
        inter = (hidden * 43) // 16
    return rows * (8 * hidden + 3 * inter) * 2


class KeepBudget:
    """Memory-adaptive checkpointing.  B200 has the HBM to keep a 32-layer step's intermediates (47 GB at 8 x 1485
    tokens), which the A100-era reference could not; but a longer batch may not fit, and running out of memory in layer
    27 is not an acceptable failure mode for a drop-in.  So every layer forward asks for its estimate: granted while
    free device memory (driver-free + the caching allocator's free blocks, read once when no layer is outstanding) minus a
    reserve (VEX_TRAIN_KEEP_RESERVE_GB, default 10 % of the device + 8 GB for optimizer state / fragmentation) covers
    it, and returned when the layer's backward (or the death of its autograd node) releases the activations.  Layers that
    are refused checkpoint themselves like the reference, so a step degrades layer by layer instead of failing."""

    def __init__(self):
        self._state = {}
        self.granted = self.refused = 0   # cumulative, for reporting (bench.py --train)

    @staticmethod
    def _available(device) -> int:
        import os
        free, total = torch.cuda.mem_get_info(device)
        cached = torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
        env = os.environ.get("VEX_TRAIN_KEEP_RESERVE_GB")
        reserve = int(float(env) * 2 ** 30) if env else total // 10 + 8 * 2 ** 30
        return free + cached - reserve

    def take(self, device, nbytes: int):
        """A token (returned to the budget when it dies) if ``nbytes`` more may be kept on ``device``, else None."""
        st = self._state.setdefault(torch.device(device).index or 0, {"out": 0, "avail": 0})
        if st["out"] == 0:
            st["avail"] = self._available(device)
        if st["avail"] < nbytes:
            self.refused += 1
            return None
        st["avail"] -= nbytes
        st["out"] += 1
        self.granted += 1
        return _KeepToken(st, nbytes)


class _KeepToken:
    def __init__(self, st, nbytes):
        self._st, self._n = st, nbytes

    def __del__(self):
        self._st["avail"] += self._n
        self._st["out"] -= 1


KEEP_BUDGET = KeepBudget()


def layer_forward_train(layer, hidden_states: torch.Tensor, plan, position_ids: torch.Tensor) -> torch.Tensor:
    tr = trainable_tensors(layer)
    for name, p in layer.named_parameters():
        if p.requires_grad and not ("lora_" in name or "layernorm" in name):
            raise NotImplementedError(f"only LoRA adapters and RMSNorm weights are trainable in the fused layer "
                                      f"(base weights are frozen under PEFT); got requires_grad on {name}")
    return _LayerFunction.apply(layer, plan, position_ids, hidden_states, *tr)


# --------------------------------------------------------------------------------------------------
# LoRA-gradient all-reduce (the only collective of the path)
# --------------------------------------------------------------------------------------------------
class BucketedGradReducer:
    """Data-parallel averaging of the trainable gradients (LoRA A / B + modules_to_save norm copies) -- the role of
    DDP's bucketed reducer in the reference (conf/phase-vlm/fit.yaml:11-15, ``gradient_as_bucket_view``), restricted to
    the adapter tensors.

    * ONE flat fp32 accumulation buffer, laid out layer by layer.  The backward kernels (K8 weight gradients, K7 norm
      gradients: fp32 atomics) accumulate straight into a layer's segment -- no per-step gradient tensors, no pack.
    * ``p.grad`` of every trainable tensor is a VIEW into the flat communication buffer (bucket view): no unpack.
      bf16 parameters are reduced in bf16 (each layer's segment is cast once, right after its backward: half the bytes
      on the wire, what DDP moves for bf16 params); fp32 parameters are reduced in place.
    * Collectives are NCCL ``AVG`` over contiguous ranges of ``layers_per_collective`` layers, issued on a side stream
      as soon as the range's last layer has finished its backward; ``finish()`` reduces whatever is left and joins.
    How many collectives: the forward / backward kernels of this path are PERSISTENT with one CTA per SM and a static
    tile assignment, so a concurrently resident NCCL kernel does not "fill gaps" -- the SMs it occupies delay the CTAs
    of the next GEMM and stretch that whole launch, and every collective also costs a cast + two stream joins.
    Measured on 2 B200 (round 2, 32 layers, r = 64, 573 MB of bf16 gradients, ~445 ms step): one collective per LAYER
    (32 per step) cost tens of ms of step time; groups of 8 layers (4 collectives, 143 MB each, the first three hidden
    behind the backward of the layers below) and one collective over the whole bucket both land within +-2 ms of the
    step without any collective (profiles/r2_train_allreduce.md).  Default: 8 layers per collective, the layer-group
    scheme of SURVEY 8(e); ``layers_per_collective=0`` = one collective after the backward."""

    def __init__(self, layers, process_group=None, layers_per_collective: Optional[int] = 8, tail_layers: int = 2):
        self.group = process_group
        self.layers = list(layers)
        self.per = layers_per_collective if layers_per_collective and layers_per_collective > 0 else len(self.layers)
        self.per = max(1, min(self.per, len(self.layers)))
        # Layer groups, in layer order.  The backward runs from the last layer to the first, so the group that contains
        # layer 0 is the only one whose collective nothing can hide: it gets ``tail_layers`` layers (2 of 32: 36 MB
        # instead of 143 MB), the groups above it ``layers_per_collective`` each.
        n = len(self.layers)
        bounds = [0]
        if self.per < n and 0 < tail_layers < self.per:
            bounds.append(tail_layers)
        while bounds[-1] < n:
            bounds.append(min(n, bounds[-1] + self.per))
        self.groups = [range(a, b) for a, b in zip(bounds[:-1], bounds[1:])]
        self.group_of = [g for g, members in enumerate(self.groups) for _ in members]
        self.segments: Dict[int, Tuple[int, int]] = {}   # id(layer) -> [lo, hi) in elements
        self.index: Dict[int, int] = {}                  # id(layer) -> position in self.layers
        self.slots: Dict[int, Tuple[int, int]] = {}      # id(param) -> [lo, hi)
        self.params: List[torch.Tensor] = []
        off = 0
        for i, layer in enumerate(self.layers):
            lo = off
            for p in trainable_tensors(layer):
                if id(p) in self.slots:
                    continue
                self.slots[id(p)] = (off, off + p.numel())
                self.params.append(p)
                off += (p.numel() + 3) // 4 * 4          # keep every slot 16-byte aligned
            self.segments[id(layer)] = (lo, off)
            self.index[id(layer)] = i
            layer._vex_grad_reducer = self
        dev = self.params[0].device if self.params else "cpu"
        self.acc = torch.zeros(off, dtype=torch.float32, device=dev)
        dtypes = {p.dtype for p in self.params}
        self.comm_dtype = torch.float32 if dtypes != {torch.bfloat16} else torch.bfloat16
        self.comm = self.acc if self.comm_dtype == torch.float32 else torch.zeros(off, dtype=torch.bfloat16, device=dev)
        self.stream = torch.cuda.Stream() if self.acc.is_cuda else None
        self._views = {}
        for p in self.params:
            lo, hi = self.slots[id(p)]
            src = self.comm if p.dtype == self.comm_dtype else None
            self._views[id(p)] = None if src is None else src[lo:hi].view_as(p)
        self._done = set()       # layer positions whose backward has finished this step
        self._reduced = set()    # group numbers already handed to NCCL this step
        self._pending = False

    @property
    def nbytes(self) -> int:
        return self.comm.numel() * self.comm.element_size()

    def owns(self, p: torch.Tensor) -> bool:
        return id(p) in self.slots

    def accumulator(self, p: torch.Tensor) -> Optional[torch.Tensor]:
        s = self.slots.get(id(p))
        return None if s is None else self.acc[s[0]:s[1]].view_as(p)

    def _world(self) -> int:
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.group)

    def _group_range(self, g: int) -> Tuple[int, int, range]:
        members = self.groups[g]
        lo = self.segments[id(self.layers[members[0]])][0]
        hi = self.segments[id(self.layers[members[-1]])][1]
        return lo, hi, members

    def _reduce_range(self, lo: int, hi: int, side: bool) -> None:
        import torch.distributed as dist
        world = self._world()
        if world == 1 or hi == lo:
            return
        if self.stream is None:  # CPU (gloo tests): synchronous, gloo has no AVG
            dist.all_reduce(self.comm[lo:hi], group=self.group)
            self.comm[lo:hi].div_(world)
            return
        if side:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                dist.all_reduce(self.comm[lo:hi], op=dist.ReduceOp.AVG, group=self.group)
            self._pending = True
        else:
            dist.all_reduce(self.comm[lo:hi], op=dist.ReduceOp.AVG, group=self.group)

    def layer_done(self, layer) -> None:
        """Called from the layer's autograd node when its gradients are complete."""
        lo, hi = self.segments[id(layer)]
        if hi > lo and self.comm is not self.acc:
            self.comm[lo:hi].copy_(self.acc[lo:hi])      # one cast of the segment, in stream order behind the backward
        i = self.index[id(layer)]
        self._done.add(i)
        g = self.group_of[i]
        glo, ghi, members = self._group_range(g)
        n_groups = len(self.groups)
        if n_groups > 1 and g not in self._reduced and all(m in self._done for m in members):
            self._reduce_range(glo, ghi, side=True)      # overlaps the backward of the layers below
            self._reduced.add(g)

    def finish(self) -> None:
        """Reduces what has not been reduced yet, joins the side stream and publishes ``p.grad`` (bucket views)."""
        for g in range(len(self.groups)):
            if g not in self._reduced:
                lo, hi, _ = self._group_range(g)
                self._reduce_range(lo, hi, side=False)
        if self.stream is not None and self._pending:
            torch.cuda.current_stream().wait_stream(self.stream)
            self._pending = False
        self._done.clear()
        self._reduced.clear()
        for p in self.params:
            v = self._views[id(p)]
            if v is None:  # mixed dtypes: this parameter is not in the communication dtype
                lo, hi = self.slots[id(p)]
                v = self.comm[lo:hi].view_as(p).to(p.dtype)
            p.grad = v

    def zero(self) -> None:
        self.acc.zero_()


class LoraGradReducer:
    """Post-backward variant (round 1): packs ``p.grad`` of every trainable tensor into one flat buffer, all-reduces
    it in chunks on a side stream and unpacks.  Kept for callers that produce gradients outside the fused layer; the
    training step uses ``BucketedGradReducer`` (overlapped, no pack / unpack)."""

    def __init__(self, params: List[torch.Tensor], chunk_bytes: int = 64 << 20, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        n = sum(p.numel() for p in self.params)
        dtype = self.params[0].dtype if self.params else torch.float32
        dev = self.params[0].device if self.params else "cpu"
        self.flat = torch.zeros(n, dtype=dtype, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.chunk = max(1, chunk_bytes // self.flat.element_size())
        self.stream = torch.cuda.Stream() if self.flat.is_cuda else None

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def reduce(self) -> None:
        import torch.distributed as dist
        for p, v in zip(self.params, self.views):
            if p.grad is not None:
                v.copy_(p.grad)
            else:
                v.zero_()
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        world = dist.get_world_size(self.group)
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                for lo in range(0, self.flat.numel(), self.chunk):
                    dist.all_reduce(self.flat[lo:lo + self.chunk], group=self.group)
                self.flat.div_(world)
            torch.cuda.current_stream().wait_stream(self.stream)
        else:
            dist.all_reduce(self.flat, group=self.group)
            self.flat.div_(world)
        for p, v in zip(self.params, self.views):
            if p.grad is not None:
                p.grad.copy_(v)
            else:
                p.grad = v.clone()
