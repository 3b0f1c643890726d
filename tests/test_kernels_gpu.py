"""GPU parity tests of the individual kernels, called through the C ABI (ctypes -> libvex.so) via the
registered torch.library ops, against the oracle (oracle/oracle_layer.py) or a plain fp32 restatement
of the same op on identical seeded inputs.  Integer/index work is bit-exact; floating point within the
tolerance written in each test."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle_layer as O  # noqa: E402


def _ops():
    from mmmm_b200 import ops
    return ops


def _plan(tt, pm):
    from mmmm_b200.plan import build_plan
    return build_plan(tt.cuda(), pm.cuda())


def _rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


# ------------------------------------------------------------------------------------------ K1
def _routing_cases(golden_dir):
    from mmmm_b200.inputs import make_ids
    g = torch.load(os.path.join(golden_dir, "routing.pt"), weights_only=False)
    cases = [(c["token_type_ids"], c["padding_mask"]) for c in g["routing"] if c["token_type_ids"].shape[1] > 1]
    gen = torch.Generator().manual_seed(0)
    for B, L in [(1, 2), (2, 257), (7, 1000), (64, 1485), (3, 5000)]:
        tt = (torch.rand(B, L, generator=gen) < 0.7).long()
        pm = torch.rand(B, L, generator=gen) < 0.85
        cases.append((tt, pm))
    tt, _, pm = make_ids(8, 1225, 256, ragged=True, seed=4)
    cases.append((tt, pm))
    tt, _, pm = make_ids(2, 2048, 512)
    cases.append((tt, pm))
    cases.append((torch.ones(2, 9, dtype=torch.long), torch.zeros(2, 9, dtype=torch.bool)))  # nothing valid
    cases.append((torch.ones(2, 9, dtype=torch.long), torch.ones(2, 9, dtype=torch.bool)))   # all vision-typed
    return cases


def test_k1_partition_bit_exact(golden_dir):
    for tt, pm in _routing_cases(golden_dir):
        want = O.routing_plan(tt, pm)
        plan = _plan(tt, pm)
        c = plan.counts.cpu().tolist()
        assert c[0] == want.vision_idx.numel() and c[1] == want.language_idx.numel()
        assert c[2] == want.valid_idx.numel()
        assert c[3] == int(pm.sum(1).max())
        assert torch.equal(plan.vision_idx().cpu(), want.vision_idx)
        assert torch.equal(plan.language_idx().cpu(), want.language_idx)
        assert torch.equal(plan.valid_idx().cpu(), want.valid_idx)
        assert torch.equal(plan.cu_seqlens.cpu().long(), want.cu_seqlens)
        T, n = c[2], tt.numel()
        s2f, f2s = plan.sorted_to_flat.cpu().long(), plan.flat_to_sorted.cpu().long()
        s2t, t2s, t2f = plan.sorted_to_token.cpu().long(), plan.token_to_sorted.cpu().long(), plan.token_to_flat.cpu().long()
        assert (s2f[T:] == -1).all() and (s2t[T:] == -1).all() and (t2s[T:] == -1).all() and (t2f[T:] == -1).all()
        assert torch.equal(f2s[s2f[:T]], torch.arange(T))               # inverse permutations
        assert torch.equal(t2s[s2t[:T]], torch.arange(T))
        assert torch.equal(t2f[s2t[:T]], s2f[:T])
        assert (f2s[~pm.reshape(-1)] == -1).all() and int((f2s >= 0).sum()) == T
        # masks with the reference's meaning
        from mmmm_b200.modeling_cogvlm import get_expert_mask
        from mmmm_b200.plan import GLOBAL_PLAN_CACHE
        GLOBAL_PLAN_CACHE.clear()
        v, l = get_expert_mask(tt.cuda(), pm.cuda())
        wv, wl = O.expert_masks(tt, pm)
        assert torch.equal(v.cpu(), wv) and torch.equal(l.cpu(), wl)


# ------------------------------------------------------------------------------------------ K2
@pytest.mark.parametrize("H", [256, 1024, 4096])
@pytest.mark.parametrize("wdtype", [torch.bfloat16, torch.float32])
def test_k2_rmsnorm_gather(H, wdtype):
    ops = _ops()
    g = torch.Generator().manual_seed(H)
    x = (torch.randn(300, H, generator=g) * 3).bfloat16()
    w = (1 + 0.2 * torch.randn(H, generator=g)).to(wdtype)
    src = torch.randperm(300, generator=g)[:211].int()
    n = torch.tensor([200], dtype=torch.int32)
    out = torch.full((211, H), 7.0, dtype=torch.bfloat16).cuda()
    ops.rmsnorm_gather(x.cuda(), w.cuda(), 1e-6, src.cuda(), n.cuda(), out)
    want = O.rms_norm(x[src[:200].long()], w, 1e-6)
    got = out.cpu()
    # identical formula; only the fp32 summation order differs -> at most 1 bf16 ulp on a few elements
    torch.testing.assert_close(got[:200].float(), want.float(), rtol=8e-3, atol=1e-5)
    assert (got[:200] != want).float().mean() < 0.02
    assert (got[200:] == 7.0).all()  # rows past the device-side count are untouched


# ------------------------------------------------------------------------------------------ K5 / K6
def test_k5_silu_mul():
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    gate = (torch.randn(130, 11008, generator=g) * 2).bfloat16()
    up = torch.randn(130, 11008, generator=g).bfloat16()
    out = torch.zeros_like(gate).cuda()
    ops.silu_mul(gate.cuda(), up.cuda(), torch.tensor([97], dtype=torch.int32).cuda(), out)
    want = torch.nn.functional.silu(gate[:97]) * up[:97]
    torch.testing.assert_close(out[:97].cpu().float(), want.float(), rtol=8e-3, atol=1e-6)
    assert (out[97:] == 0).all()


def test_k6_residual_scatter_and_copy_padded():
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    H, n_flat, n = 512, 90, 61
    y = torch.randn(n_flat, H, generator=g).bfloat16()
    res = torch.randn(n_flat, H, generator=g).bfloat16()
    dst = torch.randperm(n_flat, generator=g)[:n_flat].int()
    out = torch.zeros(n_flat, H, dtype=torch.bfloat16).cuda()
    ops.residual_scatter(y.cuda(), res.cuda(), dst.cuda(), torch.tensor([n], dtype=torch.int32).cuda(), out)
    want = torch.zeros(n_flat, H, dtype=torch.bfloat16)
    want[dst[:n].long()] = res[dst[:n].long()] + y[:n]
    assert torch.equal(out.cpu(), want)  # one bf16 add: bit-exact
    f2s = torch.full((n_flat,), -1, dtype=torch.int32)
    f2s[dst[:n].long()] = torch.arange(n, dtype=torch.int32)
    ops.copy_padded_rows(res.cuda(), f2s.cuda(), out)
    want[f2s < 0] = res[f2s < 0]
    assert torch.equal(out.cpu(), want)


# ------------------------------------------------------------------------------------------ K7 (backward, row-wise)
def test_k7_gather_rows():
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(97, 4096, generator=g).bfloat16()
    src = torch.randperm(97, generator=g)[:80].int()
    out = torch.full((80, 4096), 5.0, dtype=torch.bfloat16).cuda()
    ops.gather_rows(x.cuda(), src.cuda(), torch.tensor([61], dtype=torch.int32).cuda(), out)
    assert torch.equal(out[:61].cpu(), x[src[:61].long()])
    assert (out[61:] == 5.0).all()


def test_k7_dropout_rows_mask_properties():
    """LoRA input dropout: Bernoulli(1 - p) keep mask from a counter-based hash -- reproducible per seed, different
    across seeds, kept values scaled by 1 / (1 - p) with one bf16 rounding, rows past the live count untouched."""
    ops = _ops()
    g = torch.Generator().manual_seed(13)
    rows, n, K, p = 700, 650, 4096, 0.05
    x = torch.randn(rows, K, generator=g).bfloat16().cuda()
    nr = torch.tensor([n], dtype=torch.int32).cuda()
    out = torch.full_like(x, 9.0)
    ops.dropout_rows(x, nr, out, p, 1234567890123)
    keep = out[:n] != 0
    frac = float(keep.float().mean())
    assert abs(frac - (1 - p)) < 2e-3, frac
    want = (x[:n].float() * (1.0 / (1.0 - p))).bfloat16()
    assert torch.equal(out[:n][keep], want[keep])
    assert (out[n:] == 9.0).all()
    # per-row and per-column keep rates stay close to 1 - p (no stripes)
    assert float((keep.float().mean(0) - (1 - p)).abs().max()) < 0.05
    assert float((keep.float().mean(1) - (1 - p)).abs().max()) < 0.03
    out2 = torch.empty_like(x)
    ops.dropout_rows(x, nr, out2, p, 1234567890123)
    assert torch.equal(out2[:n], out[:n])
    ops.dropout_rows(x, nr, out2, p, 1234567890124)
    agree = float(((out2[:n] != 0) == keep).float().mean())
    assert abs(agree - ((1 - p) ** 2 + p ** 2)) < 5e-3, agree        # independent masks
    ops.dropout_rows(x, nr, out2, 0.0, 5)
    assert torch.equal(out2[:n], x[:n])


@pytest.mark.parametrize("single", [False, True])
def test_k3_dgrad_dropout_accumulate_epilogue(single, gemm_kernel):
    """out += mask / (1 - p) * (dT . A): the mask the epilogue regenerates must be the one vex_dropout_rows drew."""
    ops = _ops()
    Tv, Tl, N, r, p, seed = 300, 0 if single else 170, 768, 64, 0.25, 987654321987
    g = torch.Generator().manual_seed(14)
    T = Tv + Tl
    cap = T + 6
    dt = torch.randn(cap, r, generator=g).bfloat16().cuda()
    Av = (torch.randn(r, N, generator=g) * 0.2).bfloat16().cuda()
    Al = (torch.randn(r, N, generator=g) * 0.2).bfloat16().cuda()
    base = torch.randn(cap, N, generator=g).bfloat16().cuda()
    counts = torch.tensor([Tv, Tl, T, 0], dtype=torch.int32).cuda()
    ones = torch.ones(cap, N, dtype=torch.bfloat16).cuda()
    mask = torch.zeros_like(ones)
    ops.dropout_rows(ones, counts[2:3], mask, p, seed)                 # 0 or 1 / (1 - p)
    out = base.clone()
    ops.grouped_gemm_dgrad(dt, [Av, None if single else Al], out, counts, True, None, None, [None, None], 0, single,
                           1.0, p, seed)
    want = base.float().clone()
    want[:T] += (mask[:T].float() * _ref_dgrad(dt, Av, Al, Tv, Tl)[:T]).bfloat16().float()
    torch.testing.assert_close(out.float(), want, rtol=1e-2, atol=3e-2)


def test_k7_silu_mul_backward_vs_autograd():
    """Adjoint of the eager bf16 act_fn(gate) * up (modeling_cogvlm.py:55) against torch.autograd on CPU."""
    ops = _ops()
    g = torch.Generator().manual_seed(12)
    n, I = 97, 11008
    gate = (torch.randn(130, I, generator=g) * 2).bfloat16()
    up = torch.randn(130, I, generator=g).bfloat16()
    dact = torch.randn(130, I, generator=g).bfloat16()
    ga, ua = gate[:n].clone().requires_grad_(True), up[:n].clone().requires_grad_(True)
    (torch.nn.functional.silu(ga) * ua).backward(dact[:n])
    dg = torch.zeros_like(gate).cuda()
    du = torch.zeros_like(gate).cuda()
    ops.silu_mul_backward(dact.cuda(), gate.cuda(), up.cuda(), torch.tensor([n], dtype=torch.int32).cuda(), dg, du)
    torch.testing.assert_close(du[:n].cpu().float(), ua.grad.float(), rtol=1.6e-2, atol=1e-5)
    torch.testing.assert_close(dg[:n].cpu().float(), ga.grad.float(), rtol=1.6e-2, atol=2e-3)
    assert (dg[n:] == 0).all() and (du[n:] == 0).all()


@pytest.mark.parametrize("H", [256, 1024, 2048, 4096])
@pytest.mark.parametrize("wdtype", [torch.bfloat16, torch.float32])
def test_k7_rmsnorm_backward_vs_autograd(H, wdtype):
    """Adjoint of RMSNorm.forward (:36-41) incl. gather of x, the fused residual-gradient add, the scatter of dx and
    the fp32 weight gradient, against torch.autograd over the oracle's rms_norm."""
    ops = _ops()
    g = torch.Generator().manual_seed(H)
    n_src, n_rows, cap = 300, 200, 211
    x = (torch.randn(n_src, H, generator=g) * 3).bfloat16()
    w = (1 + 0.2 * torch.randn(H, generator=g)).to(wdtype)
    dy = torch.randn(cap, H, generator=g).bfloat16()
    add = torch.randn(cap, H, generator=g).bfloat16()
    xmap = torch.randperm(n_src, generator=g)[:cap].int()
    dmap = torch.randperm(n_src, generator=g)[:cap].int()
    xa = x[xmap[:n_rows].long()].clone().requires_grad_(True)
    wa = w.clone().requires_grad_(True)
    O.rms_norm(xa, wa, 1e-6).backward(dy[:n_rows])
    dx = torch.full((n_src, H), 7.0, dtype=torch.bfloat16).cuda()
    dw = torch.zeros(H, dtype=torch.float32).cuda()
    ops.rmsnorm_backward(dy.cuda(), x.cuda(), xmap.cuda(), w.cuda(), 1e-6, add.cuda(), None, dx, dmap.cuda(), dw,
                         torch.tensor([n_rows], dtype=torch.int32).cuda())
    want = torch.full((n_src, H), 7.0)
    want[dmap[:n_rows].long()] = xa.grad.float() + add[:n_rows].float()
    torch.testing.assert_close(dx.cpu().float(), want, rtol=1.6e-2, atol=2e-2)
    torch.testing.assert_close(dw.cpu(), wa.grad.float(), rtol=2e-2, atol=0.15)
    # accumulating into dweight, no add, identity maps
    dx2 = torch.zeros(cap, H, dtype=torch.bfloat16).cuda()
    ops.rmsnorm_backward(dy.cuda(), x[xmap.long()].contiguous().cuda(), None, w.cuda(), 1e-6, None, None, dx2, None, dw,
                         torch.tensor([n_rows], dtype=torch.int32).cuda())
    torch.testing.assert_close(dx2[:n_rows].cpu().float(), xa.grad.float(), rtol=1.6e-2, atol=2e-2)
    torch.testing.assert_close(dw.cpu(), 2 * wa.grad.float(), rtol=2e-2, atol=0.3)
    assert (dx2[n_rows:] == 0).all()


# ------------------------------------------------------------------------------------------ K3
@pytest.fixture(params=["pair", "single"], autouse=False)
def gemm_kernel(request, monkeypatch):
    """Both tile kernels: the CTA-pair (cta_group::2) default and the single-CTA variant."""
    monkeypatch.setenv("VEX_GEMM_PAIR", "1" if request.param == "pair" else "0")
    return request.param


def _gemm_problem(Tv, Tl, N, K, seed=0, cap_extra=5):
    g = torch.Generator().manual_seed(seed)
    cap = Tv + Tl + cap_extra
    a = torch.randn(cap, K, generator=g).bfloat16().cuda()
    wv = (torch.randn(N, K, generator=g) * 0.05).bfloat16().cuda()
    wl = (torch.randn(N, K, generator=g) * 0.05).bfloat16().cuda()
    counts = torch.tensor([Tv, Tl, Tv + Tl, 0], dtype=torch.int32).cuda()
    return a, wv, wl, counts, cap


def _ref_linear(a, wv, wl, Tv, Tl):
    out = torch.zeros(a.shape[0], wv.shape[0], dtype=torch.float32, device=a.device)
    out[:Tv] = a[:Tv].float() @ wv.float().T
    out[Tv:Tv + Tl] = a[Tv:Tv + Tl].float() @ wl.float().T
    return out


GEMM_SHAPES = [  # Tv, Tl, N, K
    (300, 100, 512, 256), (128, 128, 256, 64), (1, 1, 256, 128), (0, 77, 256, 192), (77, 0, 768, 256),
    (129, 255, 384, 200), (1000, 517, 1024, 1024), (2500, 300, 4096, 512),
]


@pytest.mark.parametrize("Tv,Tl,N,K", GEMM_SHAPES)
def test_k3_plain_grouped(Tv, Tl, N, K, gemm_kernel):
    ops = _ops()
    a, wv, wl, counts, cap = _gemm_problem(Tv, Tl, N, K, seed=N + K)
    out = torch.full((cap, N), 3.0, dtype=torch.bfloat16).cuda()
    ops.grouped_gemm(a, wv, wl, out, counts, None, 1.0)
    want = _ref_linear(a, wv, wl, Tv, Tl)
    T = Tv + Tl
    torch.testing.assert_close(out[:T].float(), want[:T], rtol=1e-2, atol=1e-2)
    assert (out[T:] == 3.0).all()  # masked rows untouched


def test_k3_small_n_lora_a_shape():
    """N = 64 path (the LoRA A projection) incl. alpha and single-expert mode."""
    ops = _ops()
    a, wv, wl, counts, cap = _gemm_problem(333, 140, 64, 1024, seed=5)
    out = torch.zeros(cap, 64, dtype=torch.bfloat16).cuda()
    ops.grouped_gemm(a, wv, wl, out, counts, None, 0.5)
    want = _ref_linear(a, wv, wl, 333, 140) * 0.5
    torch.testing.assert_close(out[:473].float(), want[:473], rtol=1e-2, atol=1e-2)
    out.zero_()
    ops.grouped_gemm(a, wv, None, out, counts, None, 1.0)
    torch.testing.assert_close(out[:333].float(), want[:333] * 2, rtol=1e-2, atol=1e-2)
    assert (out[333:] == 0).all()


def test_k3_scatter_and_residual(gemm_kernel):
    ops = _ops()
    Tv, Tl, N, K = 260, 140, 512, 320
    a, wv, wl, counts, cap = _gemm_problem(Tv, Tl, N, K, seed=9, cap_extra=40)
    g = torch.Generator().manual_seed(3)
    rmap = torch.randperm(cap, generator=g).int().cuda()
    res = torch.randn(cap, N, generator=g).bfloat16().cuda()
    out = torch.zeros(cap, N, dtype=torch.bfloat16).cuda()
    ops.grouped_gemm_fused(a, [wv, None, wl, None], out, counts, ops.EPI_RESIDUAL, rmap, res, [None, None],
                           [None] * 4, 0, [], 0, False, 1.0)
    T = Tv + Tl
    lin = _ref_linear(a, wv, wl, Tv, Tl)[:T].bfloat16()
    want = torch.zeros_like(out)
    want[rmap[:T].long()] = lin + res[rmap[:T].long()]
    torch.testing.assert_close(out.float(), want.float(), rtol=1e-2, atol=2e-2)
    # in place (residual == out)
    out2 = res.clone()
    ops.grouped_gemm_fused(a, [wv, None, wl, None], out2, counts, ops.EPI_RESIDUAL, rmap, None, [None, None],
                           [None] * 4, 0, [], 0, False, 1.0)
    want2 = res.clone()
    want2[rmap[:T].long()] = want[rmap[:T].long()]
    torch.testing.assert_close(out2.float(), want2.float(), rtol=1e-2, atol=2e-2)


@pytest.mark.parametrize("r", [64, 16])
def test_k3_lora_k_extension(r, gemm_kernel):
    ops = _ops()
    Tv, Tl, N, K = 200, 150, 768, 256
    a, wv, wl, counts, cap = _gemm_problem(Tv, Tl, N, K, seed=21)
    g = torch.Generator().manual_seed(4)
    Av, Al = [(torch.randn(r, K, generator=g) * 0.1).bfloat16().cuda() for _ in range(2)]
    Bv, Bl = [(torch.randn(N, r, generator=g) * 0.1).bfloat16().cuda() for _ in range(2)]
    t = torch.zeros(cap, r, dtype=torch.bfloat16).cuda()
    ops.grouped_gemm(a, Av, Al, t, counts, None, 1.0)
    t_ref = _ref_linear(a, Av, Al, Tv, Tl).bfloat16()
    torch.testing.assert_close(t[:350].float(), t_ref[:350].float(), rtol=1e-2, atol=1e-2)
    out = torch.zeros(cap, N, dtype=torch.bfloat16).cuda()
    ops.grouped_gemm_fused(a, [wv, None, wl, None], out, counts, ops.EPI_PLAIN, None, None, [t, None],
                           [Bv, None, Bl, None], r, [], 0, False, 1.0)
    want = _ref_linear(a, wv, wl, Tv, Tl) + _ref_linear(t_ref, Bv, Bl, Tv, Tl)
    torch.testing.assert_close(out[:350].float(), want[:350], rtol=1e-2, atol=2e-2)
    # vision-only adapter (lora_lang = False)
    out.zero_()
    ops.grouped_gemm_fused(a, [wv, None, wl, None], out, counts, ops.EPI_PLAIN, None, None, [t, None],
                           [Bv, None, None, None], r, [], 0, False, 1.0)
    want = _ref_linear(a, wv, wl, Tv, Tl)
    want[:Tv] += t_ref[:Tv].float() @ Bv.float().T
    torch.testing.assert_close(out[:350].float(), want[:350], rtol=1e-2, atol=2e-2)


@pytest.mark.parametrize("lora", [False, True])
def test_k3_swiglu(lora, gemm_kernel):
    ops = _ops()
    Tv, Tl, I, K, r = 300, 180, 640, 256, 64
    g = torch.Generator().manual_seed(33)
    cap = Tv + Tl + 9
    a = torch.randn(cap, K, generator=g).bfloat16().cuda()
    mk = lambda *s, sc=0.08: (torch.randn(*s, generator=g) * sc).bfloat16().cuda()
    gv, uv, gl, ul = mk(I, K), mk(I, K), mk(I, K), mk(I, K)
    counts = torch.tensor([Tv, Tl, Tv + Tl, 0], dtype=torch.int32).cuda()
    T = Tv + Tl
    gate = _ref_linear(a, gv, gl, Tv, Tl)
    up = _ref_linear(a, uv, ul, Tv, Tl)
    lt, lb, rr = [None, None], [None] * 4, 0
    if lora:
        Ag = [mk(r, K), mk(r, K)]
        Au = [mk(r, K), mk(r, K)]
        Bg = [mk(I, r), mk(I, r)]
        Bu = [mk(I, r), mk(I, r)]
        tg = torch.zeros(cap, r, dtype=torch.bfloat16).cuda()
        tu = torch.zeros(cap, r, dtype=torch.bfloat16).cuda()
        ops.grouped_gemm(a, Ag[0], Ag[1], tg, counts, None, 1.0)
        ops.grouped_gemm(a, Au[0], Au[1], tu, counts, None, 1.0)
        gate = gate + _ref_linear(tg, Bg[0], Bg[1], Tv, Tl)
        up = up + _ref_linear(tu, Bu[0], Bu[1], Tv, Tl)
        lt, lb, rr = [tg, tu], [Bg[0], Bu[0], Bg[1], Bu[1]], r
    out = torch.zeros(cap, I, dtype=torch.bfloat16).cuda()
    ops.grouped_gemm_fused(a, [gv, uv, gl, ul], out, counts, ops.EPI_SWIGLU, None, None, lt, lb, rr, [], 0,
                           False, 1.0)
    want = torch.nn.functional.silu(gate.bfloat16()) * up.bfloat16()
    torch.testing.assert_close(out[:T].float(), want[:T].float(), rtol=2e-2, atol=2e-2)
    assert (out[T:] == 0).all()


def test_k3_rope_epilogue_scatter_to_token_order(gemm_kernel):
    from mmmm_b200.inputs import make_ids
    ops = _ops()
    heads, Hd = 2, 256
    tt, pos, pm = make_ids(3, 40, 300, ragged=True, seed=8)   # positions > 256: bf16-table collapse region
    plan = _plan(tt, pm)
    B, L = tt.shape
    cap = B * L
    g = torch.Generator().manual_seed(6)
    x = torch.randn(cap, Hd, generator=g).bfloat16().cuda()       # sorted-order activations
    wv = (torch.randn(3 * Hd, Hd, generator=g) * 0.06).bfloat16().cuda()
    wl = (torch.randn(3 * Hd, Hd, generator=g) * 0.06).bfloat16().cuda()
    cos, sin = O.rotary_tables(O.default_inv_freq(128).bfloat16(), 512)
    out = torch.zeros(cap, 3 * Hd, dtype=torch.bfloat16).cuda()
    ops.grouped_gemm_fused(x, [wv, None, wl, None], out, plan.counts, ops.EPI_ROPE, plan.sorted_to_token, None,
                           [None, None], [None] * 4, 0,
                           [cos.cuda().contiguous(), sin.cuda().contiguous(), pos.reshape(-1).cuda(), plan.sorted_to_flat],
                           2 * Hd, False, 1.0)
    Tv, Tl, T = plan.counts.cpu().tolist()[:3]
    lin = _ref_linear(x, wv, wl, Tv, Tl)[:T].bfloat16().cpu()     # sorted order, eager-bf16 Linear output
    s2f = plan.sorted_to_flat.cpu().long()[:T]
    s2t = plan.sorted_to_token.cpu().long()[:T]
    p = pos.reshape(-1)[s2f]
    q, k, v = lin.split(Hd, dim=-1)
    qh, kh = q.view(1, T, heads, 128).permute(0, 2, 1, 3), k.view(1, T, heads, 128).permute(0, 2, 1, 3)
    qr, kr = O.apply_rotary(qh, kh, cos, sin, p[None])
    want = torch.cat([qr.permute(0, 2, 1, 3).reshape(T, Hd), kr.permute(0, 2, 1, 3).reshape(T, Hd), v], dim=-1)
    got = out.cpu()[s2t]
    torch.testing.assert_close(got.float(), want.float(), rtol=2e-2, atol=2e-2)


# ------------------------------------------------------------------------------------------ K3 backward (dgrad)
def _ref_dgrad(dy, wv, wl, Tv, Tl):
    """dX = dY . W with W as stored [out, in] (autograd of F.linear w.r.t. its input), per expert segment."""
    out = torch.zeros(dy.shape[0], wv.shape[1], dtype=torch.float32, device=dy.device)
    out[:Tv] = dy[:Tv].float() @ wv.float()
    out[Tv:Tv + Tl] = dy[Tv:Tv + Tl].float() @ wl.float()
    return out


DGRAD_SHAPES = [  # Tv, Tl, N (= in features), K (= out features)
    (300, 100, 512, 256), (1, 1, 256, 128), (0, 77, 256, 192), (77, 0, 768, 256), (129, 255, 320, 448),
    (1000, 517, 1024, 1024), (700, 300, 2752, 512),
]


@pytest.mark.parametrize("Tv,Tl,N,K", DGRAD_SHAPES)
def test_k3_dgrad_transposed_weights(Tv, Tl, N, K, gemm_kernel):
    """MN-major B operand: the nn.Linear weight [out = K, in = N] is read as stored."""
    ops = _ops()
    g = torch.Generator().manual_seed(N * 7 + K)
    cap = Tv + Tl + 5
    dy = torch.randn(cap, K, generator=g).bfloat16().cuda()
    wv = (torch.randn(K, N, generator=g) * 0.05).bfloat16().cuda()
    wl = (torch.randn(K, N, generator=g) * 0.05).bfloat16().cuda()
    counts = torch.tensor([Tv, Tl, Tv + Tl, 0], dtype=torch.int32).cuda()
    T = Tv + Tl
    out = torch.full((cap, N), 3.0, dtype=torch.bfloat16).cuda()
    ops.grouped_gemm_dgrad(dy, [wv, wl], out, counts, False, None, None, [None, None], 0, False, 1.0)
    want = _ref_dgrad(dy, wv, wl, Tv, Tl)
    torch.testing.assert_close(out[:T].float(), want[:T], rtol=1e-2, atol=1e-2)
    assert (out[T:] == 3.0).all()
    # accumulate in place + scatter through a row map
    gmap = torch.randperm(cap, generator=g).int().cuda()
    base = torch.randn(cap, N, generator=g).bfloat16().cuda()
    out2 = base.clone()
    ops.grouped_gemm_dgrad(dy, [wv, wl], out2, counts, True, gmap, None, [None, None], 0, False, 1.0)
    want2 = base.clone().float()
    want2[gmap[:T].long()] += want[:T].bfloat16().float()
    torch.testing.assert_close(out2.float(), want2, rtol=1e-2, atol=2e-2)


@pytest.mark.parametrize("r", [64, 16])
def test_k3_dgrad_lora(r, gemm_kernel):
    """dX = dY . W + (s dY . B) . A: the small-N transposed GEMM (dT) and the transposed K-extension."""
    ops = _ops()
    Tv, Tl, N, K, s = 200, 150, 768, 512, 0.5
    g = torch.Generator().manual_seed(77 + r)
    cap = Tv + Tl + 3
    T = Tv + Tl
    dy = torch.randn(cap, K, generator=g).bfloat16().cuda()
    mk = lambda *sh, sc=0.1: (torch.randn(*sh, generator=g) * sc).bfloat16().cuda()
    wv, wl = mk(K, N, sc=0.05), mk(K, N, sc=0.05)
    Av, Al = mk(r, N), mk(r, N)          # lora_A [r, in]
    Bv, Bl = mk(K, r), mk(K, r)          # lora_B [out, r]
    counts = torch.tensor([Tv, Tl, T, 0], dtype=torch.int32).cuda()
    dt = torch.zeros(cap, r, dtype=torch.bfloat16).cuda()
    ops.grouped_gemm_dgrad(dy, [Bv, Bl], dt, counts, False, None, None, [None, None], 0, False, s)
    dt_ref = (_ref_dgrad(dy, Bv, Bl, Tv, Tl) * s).bfloat16()
    torch.testing.assert_close(dt[:T].float(), dt_ref[:T].float(), rtol=1e-2, atol=1e-2)
    assert (dt[T:] == 0).all()
    out = torch.zeros(cap, N, dtype=torch.bfloat16).cuda()
    ops.grouped_gemm_dgrad(dy, [wv, wl], out, counts, False, None, dt, [Av, Al], r, False, 1.0)
    want = _ref_dgrad(dy, wv, wl, Tv, Tl) + _ref_dgrad(dt_ref, Av, Al, Tv, Tl)
    torch.testing.assert_close(out[:T].float(), want[:T], rtol=1e-2, atol=2e-2)
    # vision-only adapter
    out.zero_()
    ops.grouped_gemm_dgrad(dy, [wv, wl], out, counts, False, None, dt, [Av, None], r, False, 1.0)
    want = _ref_dgrad(dy, wv, wl, Tv, Tl)
    want[:Tv] += dt_ref[:Tv].float() @ Av.float()
    torch.testing.assert_close(out[:T].float(), want[:T], rtol=1e-2, atol=2e-2)


# ------------------------------------------------------------------------------------------ K8 (LoRA wgrad)
@pytest.mark.parametrize("Tv,Tl,F,r", [(300, 100, 256, 64), (1, 1, 128, 64), (0, 77, 384, 16), (77, 0, 128, 64),
                                       (2500, 1100, 1024, 64), (1024, 64, 640, 32), (5000, 3000, 11008, 64)])
def test_k8_lora_wgrad(Tv, Tl, F, r):
    """out_e = x_e^T . y_e over each expert's rows (tcgen05, both operands MN-major), fp32; rows of the other
    expert / past the live count (NaN here) must not leak in."""
    ops = _ops()
    g = torch.Generator().manual_seed(F + r + Tv)
    T = Tv + Tl
    cap = T + 70
    x = torch.randn(cap, F, generator=g).bfloat16()
    y = torch.randn(cap, r, generator=g).bfloat16()
    x[T:] = float("nan")
    y[T:] = float("nan")
    counts = torch.tensor([Tv, Tl, T, 0], dtype=torch.int32).cuda()
    xc, yc = x.cuda(), y.cuda()
    want_v = xc[:Tv].float().T @ yc[:Tv].float()
    want_l = xc[Tv:T].float().T @ yc[Tv:T].float()
    tol = dict(rtol=2e-3, atol=2e-3 * max(T, 1) ** 0.5)
    ov = torch.zeros(F, r, dtype=torch.float32).cuda()
    ol = torch.zeros(F, r, dtype=torch.float32).cuda()
    ops.lora_wgrad(xc, yc, ov, ol, False, counts)
    torch.testing.assert_close(ov, want_v, **tol)
    torch.testing.assert_close(ol, want_l, **tol)
    # transposed store (dA layout), accumulating, language adapter absent
    ot = torch.ones(r, F, dtype=torch.float32).cuda()
    ops.lora_wgrad(xc, yc, ot, None, True, counts)
    torch.testing.assert_close(ot, want_v.T + 1, **tol)


# ------------------------------------------------------------------------------------------ K4
@pytest.mark.parametrize("impl", ["tc3", "tc2", "tc2-smem", "tc2-tmem", "tc2-token", "tc1", "mma"])
@pytest.mark.parametrize("lens", [[1], [64], [65, 3, 128], [129, 128, 127], [300, 17, 1, 255], [257, 256, 255, 384],
                                  [1357], [700, 1485]])
def test_k4_attention_vs_oracle(lens, impl, monkeypatch):
    """Every implementation -- the two-tile tcgen05 kernel in its four schedules (default: early S, P in shared memory;
    P in shared memory with the serial schedule; P in TMEM; early S with exp turns), the one-tile tcgen05 kernel and
    the mma.sync baseline -- against the oracle."""
    monkeypatch.setenv("VEX_ATTN_P", impl.split("-")[1] if "-" in impl else "early")
    ops = _ops()
    attend = ops.attention if impl == "tc3" else _baseline_attention(impl.split("-")[0])
    heads = 3
    B, Lmax = len(lens), max(lens)
    g = torch.Generator().manual_seed(sum(lens))
    pm = torch.zeros(B, Lmax, dtype=torch.bool)
    for b, n in enumerate(lens):
        pm[b, :n] = True
    q, k, v = [torch.randn(B, heads, Lmax, 128, generator=g).bfloat16() for _ in range(3)]
    want = O.attention(q, k, v, pm)                     # [B, heads, L, 128]
    T = sum(lens)
    # token-order packed qkv [T, 3, heads, 128]
    tok = lambda t: t.permute(0, 2, 1, 3)[pm]            # [T, heads, 128]
    qkv = torch.stack([tok(q), tok(k), tok(v)], dim=1).reshape(T, 3 * heads * 128).contiguous()
    qkv_buf = torch.full((B * Lmax, 3 * heads * 128), float("nan"), dtype=torch.bfloat16)  # tail rows are garbage
    qkv_buf[:T] = qkv
    cu = torch.zeros(B + 1, dtype=torch.int32)
    cu[1:] = torch.tensor(lens).cumsum(0)
    rmap = torch.randperm(B * Lmax, generator=g).int()
    out = torch.zeros(B * Lmax, heads * 128, dtype=torch.bfloat16).cuda()
    attend(qkv_buf.cuda(), cu.cuda(), B, Lmax, heads, rmap.cuda(), out, 128 ** -0.5)
    got = out.cpu()[rmap[:T].long()]
    torch.testing.assert_close(got.float(), tok(want).reshape(T, heads * 128).float(), rtol=2e-2, atol=2e-2)


def _baseline_attention(impl):
    """The superseded attention kernels live in libvex_baselines.so (csrc/baselines/), outside the product library."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "vex_test_baselines", os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers", "baselines.py"))
    baselines = _BASELINES.setdefault("mod", importlib.util.module_from_spec(spec))
    if not hasattr(baselines, "attention"):
        spec.loader.exec_module(baselines)
    return lambda *a: baselines.attention(impl, *a)


_BASELINES = {}


@pytest.mark.parametrize("impl", ["tc3", "tc2"])
def test_k4_attention_many_items_per_cta(impl, monkeypatch):
    """512 work items (8 ragged samples x 16 heads x 4 query-tile pairs, some pairs past the sample's end, some with an
    inactive second tile) on at most 148 persistent CTAs: every CTA walks several items, so the running barrier phases,
    the Q / O hand-over between items and the scheduler ring of the persistent kernel are exercised; checked against
    the oracle like the small cases."""
    monkeypatch.delenv("VEX_ATTN_P", raising=False)
    ops = _ops()
    attend = ops.attention if impl == "tc3" else _baseline_attention(impl)
    heads, lens = 16, [700, 130, 1, 513, 257, 1024, 64, 385]
    B, Lmax = len(lens), max(lens)
    g = torch.Generator().manual_seed(77)
    pm = torch.zeros(B, Lmax, dtype=torch.bool)
    for b, n in enumerate(lens):
        pm[b, :n] = True
    q, k, v = [torch.randn(B, heads, Lmax, 128, generator=g).bfloat16() for _ in range(3)]
    want = O.attention(q, k, v, pm)
    T = sum(lens)
    tok = lambda t: t.permute(0, 2, 1, 3)[pm]
    qkv_buf = torch.full((B * Lmax, 3 * heads * 128), float("nan"), dtype=torch.bfloat16)
    qkv_buf[:T] = torch.stack([tok(q), tok(k), tok(v)], dim=1).reshape(T, 3 * heads * 128)
    cu = torch.zeros(B + 1, dtype=torch.int32)
    cu[1:] = torch.tensor(lens).cumsum(0)
    rmap = torch.randperm(B * Lmax, generator=g).int()
    out = torch.zeros(B * Lmax, heads * 128, dtype=torch.bfloat16).cuda()
    qkv_dev, cu_dev, rmap_dev = qkv_buf.cuda(), cu.cuda(), rmap.cuda()
    for _ in range(3):  # repeated launches reuse / rotate the work counters
        out.zero_()
        attend(qkv_dev, cu_dev, B, Lmax, heads, rmap_dev, out, 128 ** -0.5)
        got = out.cpu()[rmap[:T].long()]
        torch.testing.assert_close(got.float(), tok(want).reshape(T, heads * 128).float(), rtol=2e-2, atol=2e-2)
    untouched = torch.ones(B * Lmax, dtype=torch.bool)
    untouched[rmap[:T].long()] = False
    assert not out.cpu()[untouched].any()  # rows that belong to no live token are never written


# ------------------------------------------------------------------------------------------ K9 (attention backward)
@pytest.mark.parametrize("lens", [[1], [64], [65, 3, 128], [129, 128, 127], [300, 17, 1, 255], [700, 1485]])
def test_k9_attention_backward_vs_autograd(lens):
    """dq / dk / dv w.r.t. the PRE-rotary projections (attention + rotary adjoints fused, scattered to sorted order)
    against torch.autograd over the oracle's apply_rotary + attention in fp32."""
    ops = _ops()
    heads = 3
    H = heads * 128
    B, Lmax = len(lens), max(lens)
    T = sum(lens)
    cap = B * Lmax
    g = torch.Generator().manual_seed(sum(lens) + 1)
    pm = torch.zeros(B, Lmax, dtype=torch.bool)
    for b, n in enumerate(lens):
        pm[b, :n] = True
    pos = torch.randint(0, 300, (B, Lmax), generator=g)
    cos, sin = O.rotary_tables(O.default_inv_freq(128).bfloat16(), 512)
    q0, k0, v0 = [torch.randn(B, heads, Lmax, 128, generator=g).bfloat16() for _ in range(3)]
    d_out = torch.randn(B, heads, Lmax, 128, generator=g).bfloat16()
    # ---- reference: fp32 autograd on the GPU over the oracle functions
    qa, ka, va = [t.cuda().float().requires_grad_(True) for t in (q0, k0, v0)]
    qr, kr = O.apply_rotary(qa, ka, cos.cuda().float(), sin.cuda().float(), pos.cuda())
    out_ref = O.attention(qr, kr, va, pm.cuda())
    (out_ref * (d_out.cuda().float() * pm.cuda()[:, None, :, None])).sum().backward()
    tok = lambda t: t.permute(0, 2, 1, 3)[pm.to(t.device)].reshape(T, H)        # [T, heads*128] token order
    want = torch.cat([tok(qa.grad), tok(ka.grad), tok(va.grad)], dim=-1).cpu()
    # ---- ours: forward (rotated q/k in bf16 like the QKV epilogue produces them) with lse, then backward
    qr16, kr16 = O.apply_rotary(q0, k0, cos, sin, pos)
    qkv = torch.full((cap + 0, 3 * H), float("nan"), dtype=torch.bfloat16)
    qkv[:T] = torch.cat([tok(qr16), tok(kr16), tok(v0)], dim=-1)
    cu = torch.zeros(B + 1, dtype=torch.int32)
    cu[1:] = torch.tensor(lens).cumsum(0)
    t2s = torch.randperm(cap, generator=g).int()                                 # arbitrary token -> sorted row map
    t2f = torch.full((cap,), -1, dtype=torch.int32)
    t2f[:T] = torch.nonzero(pm.reshape(-1))[:, 0].int()
    qkv_c = qkv.cuda()
    out_sorted = torch.zeros(cap, H, dtype=torch.bfloat16).cuda()
    lse = torch.zeros(heads, cap, dtype=torch.float32).cuda()
    ops.attention_train(qkv_c, cu.cuda(), B, Lmax, heads, t2s.cuda(), out_sorted, 128 ** -0.5, lse)
    got_out = out_sorted.cpu()[t2s[:T].long()]
    torch.testing.assert_close(got_out.float(), tok(out_ref.detach()).cpu(), rtol=2e-2, atol=2e-2)
    d_tok = torch.full((cap, H), float("nan"), dtype=torch.bfloat16)            # tail rows are garbage
    d_tok[:T] = tok(d_out)
    dqkv = torch.zeros(cap, 3 * H, dtype=torch.bfloat16).cuda()
    delta = torch.empty(heads, cap, dtype=torch.float32).cuda()
    ops.attention_backward(qkv_c, out_sorted, d_tok.cuda(), lse, delta, cu.cuda(), t2s.cuda(), t2f.cuda(),
                           pos.reshape(-1).cuda(), cos.cuda().contiguous(), sin.cuda().contiguous(), B, Lmax, heads,
                           dqkv, 128 ** -0.5)
    got = dqkv.cpu()[t2s[:T].long()].float()
    # dq / dk vanish identically for single-token samples (dP == delta): floor the denominators with dv's scale
    floor_max, floor_fro = 0.05 * float(want.abs().max()), 0.05 * float(want[:, 2 * H:].norm())
    report = {}
    for name, sl in (("dq", slice(0, H)), ("dk", slice(H, 2 * H)), ("dv", slice(2 * H, 3 * H))):
        w, gt = want[:, sl], got[:, sl]
        err = float((gt - w).abs().max() / max(float(w.abs().max()), floor_max))
        fro = float((gt - w).norm() / max(float(w.norm()), floor_fro))
        report[name] = (round(err, 4), round(fro, 4))
    assert all(e <= 2e-2 and f <= 1.5e-2 for e, f in report.values()), report
    # rows that belong to no token stay untouched
    untouched = torch.ones(cap, dtype=torch.bool)
    untouched[t2s[:T].long()] = False
    assert (dqkv.cpu()[untouched] == 0).all()


# ------------------------------------------------------------------------------------------ K12 (decode skinny GEMM) / K4d
@pytest.mark.parametrize("B", [1, 3, 8, 19, 32])
@pytest.mark.parametrize("r", [0, 64])
def test_k12_decode_gemm_vs_fp32_and_k3(B, r):
    """The weight-streaming decode GEMM (mma.sync fragments loaded straight from global memory, K permuted inside
    32-element chunks) against fp32 matmuls with the eager-bf16 rounding points, for every epilogue the decode step uses
    (PLAIN with alpha, RESIDUAL separate / in place, SWIGLU, ROPE + KV-cache append) with and without the LoRA
    K-extension, and against K3 on the same inputs."""
    ops = _ops()
    g = torch.Generator().manual_seed(100 + B + r)
    H, I, heads = 512, 1408, 4
    rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).bfloat16().cuda()
    counts = torch.tensor([B, 0, B, 1], dtype=torch.int32).cuda()
    x = rnd(B, H)
    lin = lambda a, w: (a.float() @ w.float().T)
    def lora(a, A, Bm):
        return 0 if r == 0 else (a.float() @ A.float().T).bfloat16().float() @ Bm.float().T
    def t_of(a, A):
        if r == 0:
            return None
        t = torch.empty(B, r, dtype=torch.bfloat16).cuda()
        ops.grouped_gemm_raw(a, [A], t, counts, ops.EPI_PLAIN, single_expert=True, alpha=1.0, skinny=True)
        torch.testing.assert_close(t.float(), (a.float() @ A.float().T), rtol=2e-2, atol=2e-2)
        return t
    close = lambda got, want: torch.testing.assert_close(got.float(), want.float(), rtol=2e-2, atol=2e-2)

    # PLAIN with alpha
    w = rnd(768, H, sc=0.06)
    out = torch.zeros(B, 768, dtype=torch.bfloat16).cuda()
    ops.grouped_gemm_raw(x, [w], out, counts, ops.EPI_PLAIN, single_expert=True, alpha=0.5, skinny=True)
    close(out, 0.5 * lin(x, w))
    # RESIDUAL (+ LoRA), separate and in place, vs K3
    A, Bm = (rnd(r, H, sc=0.05), rnd(H, r, sc=0.05)) if r else (None, None)
    wd, res = rnd(H, H, sc=0.06), rnd(B, H)
    t = t_of(x, A)
    kw = dict(single_expert=True, lora_t=[t, None], lora_b=[Bm, None, None, None], lora_r=r, residual=res)
    o_s, o_k = torch.zeros(B, H, dtype=torch.bfloat16).cuda(), torch.zeros(B, H, dtype=torch.bfloat16).cuda()
    ops.grouped_gemm_raw(x, [wd], o_s, counts, ops.EPI_RESIDUAL, skinny=True, **kw)
    ops.grouped_gemm_raw(x, [wd], o_k, counts, ops.EPI_RESIDUAL, skinny=False, **kw)
    want = (lin(x, wd) + lora(x, A, Bm)).bfloat16().float() + res.float()
    close(o_s, want), close(o_s, o_k)
    inpl = res.clone()
    ops.grouped_gemm_raw(x, [wd], inpl, counts, ops.EPI_RESIDUAL, skinny=True, **dict(kw, residual=None))
    assert torch.equal(inpl, o_s)
    # SWIGLU (+ LoRA on gate and up)
    wg, wu = rnd(I, H, sc=0.06), rnd(I, H, sc=0.06)
    Ag, Bg, Au, Bu = (rnd(r, H, sc=0.05), rnd(I, r, sc=0.05), rnd(r, H, sc=0.05), rnd(I, r, sc=0.05)) if r else [None] * 4
    tg, tu = t_of(x, Ag), t_of(x, Au)
    act = torch.zeros(B, I, dtype=torch.bfloat16).cuda()
    ops.grouped_gemm_raw(x, [wg, wu], act, counts, ops.EPI_SWIGLU, single_expert=True, lora_t=[tg, tu],
                         lora_b=[Bg, Bu, None, None], lora_r=r, skinny=True)
    gt = (lin(x, wg) + lora(x, Ag, Bg)).bfloat16().float()
    up = (lin(x, wu) + lora(x, Au, Bu)).bfloat16().float()
    close(act, torch.nn.functional.silu(gt).bfloat16().float() * up)
    # ROPE + KV append at a device-side position
    wq = rnd(3 * H, H, sc=0.06)
    cos, sin = O.rotary_tables(O.default_inv_freq(128).bfloat16(), 600)
    pos = torch.randint(0, 600, (B,), generator=g)
    cap, at = 40, 17
    kc = torch.zeros(B, heads, cap, 128, dtype=torch.bfloat16).cuda()
    vc = torch.zeros_like(kc)
    kv_pos = torch.tensor([at], dtype=torch.int32).cuda()
    ident = torch.arange(B, dtype=torch.int32).cuda()
    qkv = torch.zeros(B, 3 * H, dtype=torch.bfloat16).cuda()
    Aq, Bq = (rnd(r, H, sc=0.05), rnd(3 * H, r, sc=0.05)) if r else (None, None)
    tq = t_of(x, Aq)
    ops.grouped_gemm_raw(x, [wq], qkv, counts, ops.EPI_ROPE, single_expert=True, lora_t=[tq, None],
                         lora_b=[Bq, None, None, None], lora_r=r,
                         rope=(cos.cuda().contiguous(), sin.cuda().contiguous(), pos.cuda(), ident), rope_cols=2 * H,
                         kv=(kc, vc, 1, kv_pos), skinny=True)
    l3 = (lin(x, wq) + lora(x, Aq, Bq)).bfloat16().cpu()
    q, k, v = l3.split(H, dim=-1)
    qh, kh = q.view(1, B, heads, 128).permute(0, 2, 1, 3), k.view(1, B, heads, 128).permute(0, 2, 1, 3)
    qr, kr = O.apply_rotary(qh, kh, cos, sin, pos[None])
    want = torch.cat([qr.permute(0, 2, 1, 3).reshape(B, H), kr.permute(0, 2, 1, 3).reshape(B, H), v], dim=-1)
    close(qkv.cpu(), want)
    assert torch.equal(kc[:, :, at].reshape(B, H), qkv[:, H:2 * H]) and torch.equal(vc[:, :, at].reshape(B, H), qkv[:, 2 * H:])
    kc[:, :, at] = 0
    vc[:, :, at] = 0
    assert not kc.any() and not vc.any()          # nothing else in the cache was touched


@pytest.mark.parametrize("B,L,cap", [(1, 700, 800), (2, 63, 64), (8, 1500, 1600), (40, 130, 130), (3, 4100, 4200)])
def test_k4d_cluster_split_decode_attention_vs_oracle(B, L, cap):
    """Decode attention over the pre-allocated cache with the positions of one (sample, head) split over a thread-block
    cluster (1 / 2 / 4 / 8 CTAs, exact softmax through distributed shared memory) against the oracle's generation
    branch; the device-side length counter selects the live prefix of the cache."""
    ops = _ops()
    heads = 4
    g = torch.Generator().manual_seed(B * 1000 + L)
    q = torch.randn(B, heads, 1, 128, generator=g).bfloat16()
    k = torch.randn(B, heads, L, 128, generator=g).bfloat16()
    v = torch.randn(B, heads, L, 128, generator=g).bfloat16()
    mask = torch.rand(B, L, generator=g) > 0.2
    mask[:, -1] = True
    want = O.attention_decode(q, k, v, mask)                     # [B, heads, 1, 128]
    kc = torch.full((B, heads, cap, 128), float("nan"), dtype=torch.bfloat16)
    vc = torch.full((B, heads, cap, 128), float("nan"), dtype=torch.bfloat16)
    kc[:, :, :L], vc[:, :, :L] = k, v
    mfull = torch.ones(B, cap, dtype=torch.bool)
    mfull[:, :L] = mask
    out = torch.zeros(B, heads * 128, dtype=torch.bfloat16).cuda()
    kv_len = torch.tensor([L - 1], dtype=torch.int32).cuda()      # positions cached before this step
    ops.attention_decode_cache(q.reshape(B, heads * 128).cuda(), kc.cuda(), vc.cuda(), mfull.cuda(), kv_len, out,
                               128 ** -0.5)
    torch.testing.assert_close(out.cpu().float(), want.reshape(B, heads * 128).float(), rtol=2e-2, atol=2e-2)
