"""The oracle (oracle/oracle_layer.py) against fixtures produced by the unmodified reference
(tests/golden/, written by oracle/make_golden.py) and, where /root/reference exists, against the
live reference itself (bit-exact in fp32 on CPU)."""
import os

import pytest
import torch

from oracle import oracle_layer as O
from oracle import reference_loader as RL


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _weights(case, dtype):
    w = {k: v.to(dtype) for k, v in case["weights"].items()}
    w["self_attn.rotary_emb.inv_freq"] = case["inv_freq"].to(dtype)
    return w


def test_routing_known_answers(golden_dir):
    g = _load(golden_dir, "routing.pt")
    for case in g["routing"]:
        v, l = O.expert_masks(case["token_type_ids"], case["padding_mask"])
        assert torch.equal(v, case["vision"]) and torch.equal(l, case["language"])
    # SURVEY 8(c)(1) literal vectors
    c0 = g["routing"][0]
    assert c0["vision"].int().tolist() == [[0, 1, 1, 1, 0, 0, 0, 0], [0, 1, 1, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 1, 1, 0]]
    assert c0["language"].int().tolist() == [[1, 0, 0, 0, 1, 1, 1, 1], [1, 0, 0, 1, 1, 1, 0, 0], [0, 1, 1, 1, 1, 0, 0, 0]]
    assert g["c1_counts"] == dict(vision=1226, language=131, total=1357)
    assert g["build_position_ids"]["y"].tolist() == [[0, 1, 2, 2, 3, 4, 5, 6]]


def test_routing_plan_matches_boolean_index_order(golden_dir):
    g = _load(golden_dir, "routing.pt")
    for case in g["routing"]:
        tt, pm = case["token_type_ids"], case["padding_mask"]
        if tt.shape[1] == 1:
            continue
        plan = O.routing_plan(tt, pm)
        flat = torch.arange(tt.numel()).view_as(tt)
        assert torch.equal(plan.vision_idx, flat[case["vision"]])
        assert torch.equal(plan.language_idx, flat[case["language"]])
        assert torch.equal(plan.valid_idx, flat[pm])
        assert plan.cu_seqlens[-1] == pm.sum()


def test_bf16_arange_quirk(golden_dir):
    g = _load(golden_dir, "routing.pt")
    want = [250, 251, 252, 253, 254, 255, 256, 256, 258, 260, 260, 260, 262, 264, 264, 264, 266, 268, 268, 268]
    assert g["bf16_arange_250_270"].tolist() == want
    cos, _ = O.rotary_tables(O.default_inv_freq(128).to(torch.bfloat16), 270)
    assert cos.dtype == torch.bfloat16
    assert torch.equal(cos[256], cos[257]) and torch.equal(cos[259], cos[260])


@pytest.mark.parametrize("name", ["layer_tiny.pt", "layer_longpos.pt", "layer_pos600.pt"])
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_oracle_matches_golden_layer(golden_dir, name, prec):
    case = _load(golden_dir, name)
    dtype = torch.float32 if prec == "fp32" else torch.bfloat16
    cfg = case["config"]
    out, (k, v) = O.decoder_layer(
        _weights(case, dtype), case["hidden_states"].to(dtype), case["token_type_ids"], case["position_ids"],
        case["padding_mask"], num_heads=cfg["num_heads"], rms_norm_eps=cfg["rms_norm_eps"], use_cache=True)
    pm = case["padding_mask"]
    ref = case[prec]
    # same ATen ops on the same shapes -> bit-exact on the same CPU; allow 1e-4 across machines (fp32)
    tol = dict(rtol=0, atol=1e-4) if prec == "fp32" else dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(out[pm].float(), ref["out"][pm].float(), **tol)
    pmh = pm[:, None, :].expand(k.shape[:3])
    torch.testing.assert_close(k[pmh].float(), ref["k"][pmh].float(), **tol)
    torch.testing.assert_close(v[pmh].float(), ref["v"][pmh].float(), **tol)
    cos, sin = O.rotary_tables(case["inv_freq"].to(dtype), ref["cos"].shape[0])
    assert torch.equal(cos, ref["cos"]) and torch.equal(sin, ref["sin"])


@pytest.mark.skipif(not RL.reference_available(), reason="/root/reference not present on this machine")
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_oracle_bit_exact_vs_live_reference(dtype):
    from mmmm_b200.inputs import make_ids
    hidden, heads, inter = 256, 2, 320
    layer, cfg = RL.make_reference_layer(hidden, inter, heads, dtype=dtype, seed=5)
    tt, pos, pm = make_ids(3, 17, 12, ragged=True, seed=5)
    h = torch.randn(3, tt.shape[1], hidden, generator=torch.Generator().manual_seed(9)).to(dtype)
    with torch.no_grad():
        ref_out, ref_kv = layer(h, token_type_ids=tt, position_ids=pos, padding_mask=pm, use_cache=True)
    w = {k: v for k, v in layer.state_dict().items()}
    out, (k, v) = O.decoder_layer(w, h, tt, pos, pm, num_heads=heads, rms_norm_eps=cfg.rms_norm_eps, use_cache=True)
    assert torch.equal(out[pm], ref_out[pm])
    assert torch.equal(k, ref_kv[0]) and torch.equal(v, ref_kv[1])


@pytest.mark.skipif(not RL.reference_available(), reason="/root/reference not present on this machine")
def test_stack_vs_live_reference_llm_forward():
    """The caller loop (modeling_cogvlm.py:547-573) through the real CogVLMModel.llm_forward."""
    from mmmm_b200.inputs import make_ids
    M = RL.load_reference()
    cfg = M.CogVLMConfig(hidden_size=256, intermediate_size=256, num_attention_heads=2, num_hidden_layers=2,
                         vocab_size=64, vision_config={})
    cfg.lora_lang = True
    torch.manual_seed(3)
    model = M.CogVLMModel(cfg).eval()
    tt, pos, pm = make_ids(2, 9, 6, ragged=True, seed=1)
    emb = torch.randn(2, tt.shape[1], 256)
    with torch.no_grad():
        ref = model.llm_forward(inputs_embeds=emb, token_type_ids=tt, position_ids=pos,
                                attention_mask=pm.long(), use_cache=False, return_dict=True).last_hidden_state
    lw = [dict(l.state_dict()) for l in model.layers]
    out = O.decoder_stack(lw, emb, tt, pos, pm, num_heads=2, rms_norm_eps=cfg.rms_norm_eps,
                          final_norm_weight=model.norm.weight.detach())
    assert torch.equal(out[pm], ref[pm])


def test_lora_restatement_is_additive():
    H, I = 256, 256
    w = O.random_weights(H, I, 2, seed=0)
    lora = O.random_lora(H, I, r=16, seed=1)
    x = torch.randn(5, H)
    p = "self_attn.vision_expert_dense"
    y0 = O.linear(x, w[p + ".weight"])
    y1 = O.linear(x, w[p + ".weight"], lora[p])
    torch.testing.assert_close(y1 - y0, x @ lora[p].A.T @ lora[p].B.T, rtol=1e-4, atol=1e-5)


@pytest.mark.skipif(not RL.reference_available(), reason="/root/reference not present on this machine")
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_oracle_decode_steps_bit_exact_vs_live_reference(dtype):
    """Prefill with use_cache, then two q_len == 1 steps through the generation branch (:129-141, :258-262)."""
    from mmmm_b200.inputs import make_ids
    hidden, heads, inter = 256, 2, 320
    layer, cfg = RL.make_reference_layer(hidden, inter, heads, dtype=dtype, seed=6)
    tt, pos, pm = make_ids(3, 11, 9, ragged=True, seed=6)
    g = torch.Generator().manual_seed(10)
    h = torch.randn(3, tt.shape[1], hidden, generator=g).to(dtype)
    w = dict(layer.state_dict())
    with torch.no_grad():
        _, ref_kv = layer(h, token_type_ids=tt, position_ids=pos, padding_mask=pm, use_cache=True)
    _, kv = O.decoder_layer(w, h, tt, pos, pm, num_heads=heads, rms_norm_eps=cfg.rms_norm_eps, use_cache=True)
    mask = pm.clone()
    next_pos = pos.max(dim=1, keepdim=True).values + 1
    for step in range(2):
        x = torch.randn(3, 1, hidden, generator=g).to(dtype)
        mask = torch.cat([mask, torch.ones(3, 1, dtype=torch.bool)], dim=1)
        tt1 = torch.zeros(3, 1, dtype=torch.long)
        with torch.no_grad():
            ref_out, ref_kv = layer(x, token_type_ids=tt1, position_ids=next_pos + step, padding_mask=mask,
                                    past_key_value=ref_kv, use_cache=True)
        out, kv = O.decoder_layer(w, x, tt1, next_pos + step, mask, num_heads=heads,
                                  rms_norm_eps=cfg.rms_norm_eps, use_cache=True, past_key_value=kv)
        assert torch.equal(out, ref_out)
        assert torch.equal(kv[0], ref_kv[0]) and torch.equal(kv[1], ref_kv[1])


def test_lm_head_ce_vs_golden_and_live_reference(golden_dir):
    """the oracle's lm_head + _sample_weighted_ce restatement against the fixture written by the
    unmodified reference function, and bit-exact against the live function where /root/reference exists."""
    g = _load(golden_dir, "lm_head_ce.pt")
    for prec, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        h, w = g["hidden_states"].to(dt), g["lm_head_weight"].to(dt)
        got_w = O.lm_head_loss(h, w, g["labels"], g["weight"])
        got_p = O.lm_head_loss(h, w, g["labels"], None)
        assert torch.equal(got_w, g["loss"][prec]["weighted"]) and torch.equal(got_p, g["loss"][prec]["plain"])
    if RL.reference_available():
        M = RL.load_reference()
        logits = torch.nn.functional.linear(g["hidden_states"].float(), g["lm_head_weight"].float())
        for wt in (g["weight"], None):
            assert torch.equal(M._sample_weighted_ce(logits, g["labels"], wt), O.sample_weighted_ce(logits, g["labels"], wt))
    # all labels ignored -> NaN, like the reference (0 / 0)
    lab = torch.full_like(g["labels"], -100)
    assert torch.isnan(O.lm_head_loss(g["hidden_states"].float(), g["lm_head_weight"].float(), lab, g["weight"]))
