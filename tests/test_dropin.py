"""CPU-side checks of the drop-in boundary: module tree / state-dict keys equal the
reference's, parameters move over without copies, PEFT-style wrappers are read through, errors are loud."""
import pytest
import torch
from torch import nn

from mmmm_b200.modeling_cogvlm import CogVLMDecoderLayer, VexConfig, swap_decoder_layers
from mmmm_b200.peft_compat import MockLoraLinear, attach_mock_lora, resolve_linear, resolve_norm
from oracle import reference_loader as RL

EXPECTED_KEYS = {
    "self_attn.rotary_emb.inv_freq": (64,),
    "self_attn.vision_expert_query_key_value.weight": (768, 256),
    "self_attn.vision_expert_dense.weight": (256, 256),
    "self_attn.language_expert_query_key_value.weight": (768, 256),
    "self_attn.language_expert_dense.weight": (256, 256),
    "mlp.language_mlp.gate_proj.weight": (320, 256),
    "mlp.language_mlp.up_proj.weight": (320, 256),
    "mlp.language_mlp.down_proj.weight": (256, 320),
    "mlp.vision_mlp.gate_proj.weight": (320, 256),
    "mlp.vision_mlp.up_proj.weight": (320, 256),
    "mlp.vision_mlp.down_proj.weight": (256, 320),
    "input_layernorm.weight": (256,),
    "post_attention_layernorm.weight": (256,),
}
CFG = dict(hidden_size=256, intermediate_size=320, num_attention_heads=2)


def test_state_dict_keys_and_shapes():
    layer = CogVLMDecoderLayer(VexConfig(**CFG))
    got = {k: tuple(v.shape) for k, v in layer.state_dict().items()}
    assert got == EXPECTED_KEYS
    assert type(layer).__name__ == "CogVLMDecoderLayer"          # _no_split_modules (:347)
    assert type(layer.input_layernorm.weight).__name__ == "NoWeightDecayParameter"


@pytest.mark.skipif(not RL.reference_available(), reason="/root/reference not present")
def test_same_tree_as_live_reference_and_swap():
    M = RL.load_reference()
    cfg = M.CogVLMConfig(num_hidden_layers=2, vocab_size=32, vision_config={}, **CFG)
    cfg.lora_lang = True
    model = M.CogVLMModel(cfg)
    ref_layer = model.layers[0]
    ours = CogVLMDecoderLayer(cfg)                                  # accepts the reference's config object
    assert {k: tuple(v.shape) for k, v in ours.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in ref_layer.state_dict().items()}
    assert [n for n, _ in ours.named_modules()] == [n for n, _ in ref_layer.named_modules()]
    ours.load_state_dict(ref_layer.state_dict(), strict=True)
    keys_before = list(model.state_dict().keys())
    ptrs = {k: v.data_ptr() for k, v in model.state_dict().items()}
    swap_decoder_layers(model)
    assert all(isinstance(l, CogVLMDecoderLayer) for l in model.layers)
    assert list(model.state_dict().keys()) == keys_before
    assert {k: v.data_ptr() for k, v in model.state_dict().items()} == ptrs   # moved, not copied
    # the LoRA target hooks select the same module names as the reference's (mmmm/utils.py:19-43 semantics)
    t_attn, _ = ours.self_attn.get_lora_modules("model.layers.0.self_attn")
    assert sorted(t_attn) == sorted(f"model.layers.0.self_attn.{n}" for n in (
        "vision_expert_query_key_value", "vision_expert_dense", "language_expert_query_key_value",
        "language_expert_dense"))
    t_mlp, _ = ours.mlp.get_lora_modules("mlp")
    assert len(t_mlp) == 6
    ours.self_attn.config.lora_lang = False
    assert len(ours.self_attn.get_lora_modules("a")[0]) == 2 and len(ours.mlp.get_lora_modules("m")[0]) == 3
    ours.self_attn.config.lora_lang = True


def test_mock_peft_wrapping_and_resolver():
    layer = CogVLMDecoderLayer(VexConfig(**CFG))
    base_w = layer.self_attn.vision_expert_dense.weight
    attach_mock_lora(layer, r=16, b_std=0.02)
    keys = set(layer.state_dict().keys())
    # PEFT's in-module key names
    assert "self_attn.vision_expert_dense.base_layer.weight" in keys
    assert "self_attn.vision_expert_dense.lora_A.default.weight" in keys
    assert "self_attn.vision_expert_dense.lora_B.default.weight" in keys
    assert "input_layernorm.modules_to_save.default.weight" in keys
    assert "input_layernorm.original_module.weight" in keys
    spec = resolve_linear(layer.self_attn.vision_expert_dense)
    assert spec.weight is base_w and spec.r == 16 and spec.scaling == 8 / 4
    assert spec.lora_A.shape == (16, 256) and spec.lora_B.shape == (256, 16)
    m = layer.self_attn.vision_expert_dense
    m.disable_adapters = True
    assert resolve_linear(m).lora_A is None
    m.disable_adapters, m.merged = False, True
    assert resolve_linear(m).lora_A is None
    m.merged = False
    m.active_adapters = ["default", "other"]
    m.lora_A["other"], m.lora_B["other"] = nn.Linear(256, 16, bias=False), nn.Linear(16, 256, bias=False)
    with pytest.raises(NotImplementedError):
        resolve_linear(m)
    wrapped = layer.input_layernorm
    assert resolve_norm(wrapped) is wrapped.modules_to_save["default"]
    wrapped.disable_adapters = True
    assert resolve_norm(wrapped) is wrapped.original_module
    # the mock computes PEFT's formula (used by bench/tests as the eager LoRA reference)
    lin = nn.Linear(8, 4, bias=False)
    ml = MockLoraLinear(lin, r=8, b_std=0.1)
    x = torch.randn(3, 8)
    want = lin(x) + (x @ ml.lora_A["default"].weight.T @ ml.lora_B["default"].weight.T) * ml.scaling["default"]
    torch.testing.assert_close(ml(x), want)


def test_errors_are_loud_on_cpu():
    layer = CogVLMDecoderLayer(VexConfig(**CFG))
    h = torch.zeros(1, 4, 256, dtype=torch.bfloat16)
    ids = torch.zeros(1, 4, dtype=torch.long)
    pm = torch.ones(1, 4, dtype=torch.bool)
    with pytest.raises(ValueError, match="no CPU path"):
        layer(h, ids, ids, pm)
    with pytest.raises(NotImplementedError, match="past_key_value"):
        layer(h, ids, ids, pm, past_key_value=(h, h))
    with pytest.raises(TypeError):
        layer(h)
    with pytest.raises(ValueError, match="head_dim"):
        CogVLMDecoderLayer(VexConfig(hidden_size=256, intermediate_size=256, num_attention_heads=4))


def test_rotary_module_keeps_reference_table_semantics():
    from oracle import oracle_layer as O
    from mmmm_b200.modeling_cogvlm import RotaryEmbedding
    rot = RotaryEmbedding(128, max_position_embeddings=64)
    cos, sin = rot.tables(300, torch.device("cpu"), torch.float32)
    c_ref, s_ref = O.rotary_tables(O.default_inv_freq(128), 300)
    assert torch.equal(cos[:300], c_ref) and torch.equal(sin[:300], s_ref)
    rot = rot.to(torch.bfloat16)                                   # bf16-true: table rebuilt in bf16 (quirk 2)
    cos, _ = rot.tables(300, torch.device("cpu"), torch.bfloat16)
    c_ref, _ = O.rotary_tables(O.default_inv_freq(128).bfloat16(), 300)
    assert cos.dtype == torch.bfloat16 and torch.equal(cos[:300], c_ref)
    assert torch.equal(cos[256], cos[257])
    c3, s3 = rot(torch.zeros(1, dtype=torch.bfloat16), seq_len=10)  # reference forward signature
    assert c3.shape == (10, 1, 128)


def test_plan_cache_key_tolerates_inference_tensors():
    """Inference tensors (torch.inference_mode -- how scripts/evaluate/models/mmmm.py:132 calls generate) have no
    version counter; the plan-cache key must not touch `_version` on them."""
    import torch
    from mmmm_b200.plan import _version_of
    a = torch.zeros(2, 3, dtype=torch.long)
    v0 = _version_of(a)
    a.add_(1)
    assert _version_of(a) != v0                       # normal tensors: in-place edits invalidate the plan
    with torch.inference_mode():
        b = torch.zeros(2, 3, dtype=torch.long)
    key = _version_of(b)                              # would raise "Inference tensors do not track version counter"
    assert key[0] == "inference" and key == _version_of(b)


def test_kv_cache_view_capacity_detection():
    """Tuple-cache decode appends in place only into views of a buffer with room left (modeling_cogvlm._cache_capacity)."""
    import torch
    from mmmm_b200.modeling_cogvlm import _cache_capacity, _full_cache_view
    kv = torch.zeros(2, 3, 4, 10, 128)                # [k|v, B, heads, capacity, 128]
    k, v = kv[0][:, :, :7], kv[1][:, :, :7]
    assert _cache_capacity(k) == 10 and _cache_capacity(v) == 10           # v starts at a storage offset
    full = _full_cache_view(v, 10)
    assert full.shape == (3, 4, 10, 128) and full.data_ptr() == kv[1].data_ptr()
    assert _cache_capacity(k.contiguous()) == 7                             # exactly full: no room -> re-allocate
    assert _cache_capacity(torch.zeros(3, 4, 7, 64)) == 0                   # wrong head_dim
    assert _cache_capacity(kv[0].transpose(1, 2)[:, :, :3]) == 0            # not the cache layout


def test_next_position_ids_rule():
    """mmmm.py:354-365 (+1 per step) and :380-384 (keep_position: the token after <bop> and an <eop> keep the previous
    position)."""
    from mmmm_b200.kv_cache import next_position_ids
    BOP, EOP = 7, 8
    pos = torch.tensor([[3, 4], [3, 4], [3, 4], [3, 4]])
    ids = torch.tensor([[5, 6, 9],      # ordinary token                 -> 5
                        [5, BOP, 9],    # the token right after <bop>    -> 4
                        [5, 6, EOP],    # <eop> itself                   -> 4
                        [5, BOP, EOP]])  # both conditions: still one back -> 4
    # the reference, literally: cat([pos, pos[:, -1:] + 1]); pos[:, -1] -= keep; pos[:, -1:]
    ref = torch.cat([pos, pos[:, -1:] + 1], dim=1)
    keep = (ids[:, -2] == BOP) | (ids[:, -1] == EOP)
    ref[:, -1] -= keep.long()
    got = next_position_ids(pos, ids, BOP, EOP)
    assert got.shape == (4, 1) and torch.equal(got, ref[:, -1:]) and got[:, 0].tolist() == [5, 4, 4, 4]


def test_static_cache_reorder_matches_reference_reorder():
    """StaticKVCache.reorder == the reference's _reorder_cache (modeling_cogvlm.py:782-788) on the tuple view, in place."""
    from mmmm_b200.kv_cache import StaticKVCache
    cache = StaticKVCache(n_layers=2, batch=3, heads=2, capacity=6, device="cpu", dtype=torch.float32)
    cache._kv.copy_(torch.arange(cache._kv.numel(), dtype=torch.float32).view_as(cache._kv))
    pm = torch.tensor([[1, 1, 1, 0], [1, 1, 0, 0], [1, 1, 1, 1]], dtype=torch.bool)
    cache.start(pm)
    before = tuple((k.clone(), v.clone()) for k, v in cache.views())
    ptrs = [k.data_ptr() for k, _ in cache.layers]
    beam = torch.tensor([2, 0, 0])
    cache.reorder(beam)
    want = tuple(tuple(t.index_select(0, beam) for t in layer) for layer in before)
    for (k, v), (wk, wv) in zip(cache.views(), want):
        assert torch.equal(k, wk) and torch.equal(v, wv)
    assert torch.equal(cache.mask[:, :4], pm.index_select(0, beam)) and bool(cache.mask[:, 4:].all())
    assert [k.data_ptr() for k, _ in cache.layers] == ptrs


def test_keep_budget_policy(monkeypatch):
    """Memory-adaptive checkpointing (training.KeepBudget): activations are kept while the device has room beyond the
    reserve, layers past that are refused (they checkpoint like the reference, mmmm.py:287-291), and the budget comes
    back when a layer's token dies (its backward ran or its autograd node was dropped)."""
    import gc
    from mmmm_b200 import training as T
    GB = 2 ** 30
    monkeypatch.setenv("VEX_TRAIN_KEEP_RESERVE_GB", "10")
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda device=None: (30 * GB, 180 * GB))
    monkeypatch.setattr(torch.cuda, "memory_reserved", lambda device=None: 6 * GB)
    monkeypatch.setattr(torch.cuda, "memory_allocated", lambda device=None: 4 * GB)
    budget = T.KeepBudget()                       # available = 30 + (6 - 4) - 10 = 22 GB
    dev = torch.device("cuda", 0)
    tokens = [budget.take(dev, 5 * GB) for _ in range(5)]
    assert [t is not None for t in tokens] == [True, True, True, True, False]
    assert (budget.granted, budget.refused) == (4, 1)
    tokens[0] = None                              # one layer's backward ran
    gc.collect()
    again = budget.take(dev, 5 * GB)
    refused = budget.take(dev, 5 * GB)
    assert again is not None and refused is None
    del again
    # everything released: the next request re-reads the device state
    tokens.clear()
    gc.collect()
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda device=None: (12 * GB, 180 * GB))
    assert budget.take(dev, 5 * GB) is None       # 12 + 2 - 10 = 4 GB

    class L:  # the policy switch: attribute beats environment, default is "auto"
        pass
    layer = L()
    monkeypatch.delenv("VEX_TRAIN_RECOMPUTE", raising=False)
    assert T.recompute_default(layer) == "auto"
    monkeypatch.setenv("VEX_TRAIN_RECOMPUTE", "1")
    assert T.recompute_default(layer) is True
    layer.recompute = False
    assert T.recompute_default(layer) is False
    assert T.keep_bytes_estimate(layer, 1000, 4096) == 1000 * (8 * 4096 + 3 * 11008) * 2


def test_bucketed_reducer_group_boundaries():
    """Layer groups of the overlapped LoRA-grad all-reduce (conf/phase-vlm/fit.yaml:11-15 bucket views): the
    group that holds layer 0 -- the last one the backward finishes, the only collective nothing can hide -- is short."""
    from mmmm_b200.training import BucketedGradReducer
    layers = [torch.nn.Linear(4, 4, bias=False) for _ in range(32)]
    as_ranges = lambda r: [(g.start, g.stop) for g in r.groups]
    assert as_ranges(BucketedGradReducer(layers)) == [(0, 2), (2, 10), (10, 18), (18, 26), (26, 32)]
    assert as_ranges(BucketedGradReducer(layers, layers_per_collective=8, tail_layers=0)) == [(0, 8), (8, 16), (16, 24), (24, 32)]
    assert as_ranges(BucketedGradReducer(layers, layers_per_collective=0)) == [(0, 32)]
    assert as_ranges(BucketedGradReducer(layers[:3], layers_per_collective=1)) == [(0, 1), (1, 2), (2, 3)]
    r = BucketedGradReducer(layers)
    assert r.group_of[:3] == [0, 0, 1] and r.group_of[-1] == 4 and len(r.group_of) == 32
