"""TEST INFRASTRUCTURE -- a plain-PyTorch restatement of the layer backward (cuBLAS matmuls, SDPA attention
backward) over the expert-sorted buffers of the native forward.  It was the interim backward of the training step
before the native kernels (K3 dgrad, K7, K8, K9) existed; it is kept only as a second checker for
tests/test_layer_gpu.py next to the oracle's autograd.  Nothing under mmmm_b200/ imports it."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from mmmm_b200.peft_compat import LinearSpec

HEAD_DIM = 128


# --------------------------------------------------------------------------------------------------
# pieces of the backward (row-wise formulas over expert-sorted rows)
# --------------------------------------------------------------------------------------------------
def _rmsnorm_bwd(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor, eps: float):
    """y = w * (x * rsqrt(mean(x^2) + eps)); returns (dx, dw) in fp32 math."""
    xf, dyf, wf = x.float(), dy.float(), w.float()
    inv = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    xhat = xf * inv
    dyw = dyf * wf
    dx = inv * (dyw - xhat * (dyw * xhat).mean(-1, keepdim=True))
    dw = (dyf * xhat).sum(0)
    return dx.to(x.dtype), dw


def _linear_bwd(dy: torch.Tensor, x: torch.Tensor, t: Optional[torch.Tensor], spec: LinearSpec):
    """y = x W^T + (s x A^T) B^T, T = s x A^T (bf16, as the forward produced it).
    Returns dx, (dA, dB) or None."""
    dx = dy @ spec.weight.detach()
    grads = None
    if spec.lora_A is not None:
        A, Bm, s = spec.lora_A.detach().to(dy.dtype), spec.lora_B.detach().to(dy.dtype), spec.scaling
        dT = dy @ Bm                                    # [n, r]
        dB = dy.float().t() @ t.float()                 # [out, r]   (T already carries the scaling)
        dA = (dT.float().t() @ x.float()) * s           # [r, in]
        dx = dx + (dT @ A) * s
        grads = (dA, dB)
    return dx, grads


def _rotate_half_t(z: torch.Tensor) -> torch.Tensor:
    """Transpose of rotate_half: R(x) = cat(-x2, x1)  =>  R^T(z) = cat(z2, -z1)."""
    half = z.shape[-1] // 2
    return torch.cat((z[..., half:], -z[..., :half]), dim=-1)


def layer_backward(layer, plan, position_ids: torch.Tensor, hidden_states: torch.Tensor, d_out: torch.Tensor):
    """Returns (d_hidden [B, L, H], {param tensor id -> grad}) -- see the module docstring."""
    from mmmm_b200.modeling_cogvlm import visual_expert_layer_forward
    B, L, H = hidden_states.shape
    cap = B * L
    attn = layer.self_attn
    heads = attn.num_heads
    keep: Dict = {}
    with torch.no_grad():
        visual_expert_layer_forward(layer, hidden_states, plan, position_ids, keep=keep)
    Tv, Tl = plan.counts[:2].tolist()   # interim: host sync (the native kernels read the counts on the device)
    T = Tv + Tl
    seg = ((0, Tv), (Tv, T))
    s2f = plan.sorted_to_flat[:T].long()
    s2t = plan.sorted_to_token[:T].long()
    specs = keep["specs"]
    grads: Dict[int, torch.Tensor] = {}
    params: Dict[int, torch.Tensor] = {}

    def add_grad(p: torch.Tensor, g: torch.Tensor):
        if p.requires_grad:
            params[id(p)] = p
            grads[id(p)] = grads[id(p)] + g.to(p.dtype) if id(p) in grads else g.to(p.dtype)

    def routed_linear_bwd(dy, x, t, pair):
        dx = torch.empty(T, pair[0].weight.shape[1], dtype=dy.dtype, device=dy.device)
        for (lo, hi), spec in zip(seg, pair):
            if hi == lo:
                continue
            dxe, g = _linear_bwd(dy[lo:hi], x[lo:hi], None if t is None else t[lo:hi], spec)
            dx[lo:hi] = dxe
            if g is not None:
                add_grad(spec.lora_A, g[0])
                add_grad(spec.lora_B, g[1])
        return dx

    hf = hidden_states.view(cap, H)
    dof = d_out.reshape(cap, H)
    dy = dof[s2f]                                                      # [T, H] sorted
    # ---- MLP block ----
    dact = routed_linear_bwd(dy, keep["act"][:T], keep["t_down"], specs["down"])
    g, u = keep["gate"][:T].float(), keep["up"][:T].float()
    sig = torch.sigmoid(g)
    silu = (g * sig).to(dy.dtype).float()                              # forward rounds silu(g) to bf16
    dactf = dact.float()
    du = (dactf * silu).to(dy.dtype)
    dg = (dactf * u * (sig * (1 + g * (1 - sig)))).to(dy.dtype)
    xn2 = keep["xn2"][:T]
    dxn2 = routed_linear_bwd(dg, xn2, keep["t_gate"], specs["gate"]) + \
        routed_linear_bwd(du, xn2, keep["t_up"], specs["up"])
    h1s = keep["h1"].view(cap, H)[s2f]
    dx2, dw2 = _rmsnorm_bwd(dxn2, h1s, keep["ln2"].weight.detach(), keep["ln2"].variance_epsilon)
    add_grad(keep["ln2"].weight, dw2)
    dh1 = dy + dx2                                                     # grad w.r.t. h1 rows (sorted)
    # ---- attention block ----
    dctx = routed_linear_bwd(dh1, keep["ctx"][:T], keep["t_dense"], specs["dense"])
    qkv = keep["qkv"][:T].view(T, 3, heads, HEAD_DIM)                  # token order
    dctx_tok = torch.empty_like(dctx)
    dctx_tok[s2t] = dctx
    dctx_tok = dctx_tok.view(T, heads, HEAD_DIM)
    dqkv = torch.empty(T, 3, heads, HEAD_DIM, dtype=dy.dtype, device=dy.device)
    cu = plan.cu_seqlens.tolist()
    for b in range(B):                                                 # interim: SDPA backward per sample
        lo, hi = cu[b], cu[b + 1]
        if hi == lo:
            continue
        with torch.enable_grad():
            q, k, v = (qkv[lo:hi, i].transpose(0, 1).unsqueeze(0).detach().requires_grad_(True) for i in range(3))
            o = F.scaled_dot_product_attention(q, k, v, is_causal=True)
            gq, gk, gv = torch.autograd.grad(o, (q, k, v), dctx_tok[lo:hi].transpose(0, 1).unsqueeze(0))
        for i, gi in enumerate((gq, gk, gv)):
            dqkv[lo:hi, i] = gi[0].transpose(0, 1)
    # rotary backward on q and k (token order): y = x c + R(x) s  =>  dx = dy c + R^T(dy s)
    pos = position_ids.reshape(-1)[plan.token_to_flat[:T].long()]
    c = keep["cos"][pos].unsqueeze(1)
    s_ = keep["sin"][pos].unsqueeze(1)
    dqk = dqkv[:, :2]
    dqkv[:, :2] = (dqk * c.unsqueeze(1)) + _rotate_half_t(dqk * s_.unsqueeze(1))
    dqkv_sorted = dqkv.view(T, 3 * H)[s2t]                             # back to expert-sorted rows
    dxn1 = routed_linear_bwd(dqkv_sorted, keep["xn1"][:T], keep["t_qkv"], specs["qkv"])
    dx1, dw1 = _rmsnorm_bwd(dxn1, hf[s2f], keep["ln1"].weight.detach(), keep["ln1"].variance_epsilon)
    add_grad(keep["ln1"].weight, dw1)
    d_hidden = d_out.clone().view(cap, H)                              # residual path (and padded rows)
    d_hidden[s2f] = (dh1 + dx1).to(d_hidden.dtype)
    return d_hidden.view(B, L, H), [(params[i], grads[i]) for i in grads]


