"""The vision-encoder oracle (oracle/oracle_vision.py) against the fixture produced by the unmodified reference
(tests/golden/vision_tiny.pt, written by oracle/make_golden.py) and, where /root/reference exists, against the live
reference (bit-exact on CPU in fp32 and bf16); plus the host logic of the product module (mmmm_b200/visual.py):
state-dict keys, packing plan, head-slot padding, loud failure on CPU tensors."""
import os

import pytest
import torch

from oracle import oracle_vision as OV
from oracle import reference_loader as RL


def _case(golden_dir):
    return torch.load(os.path.join(golden_dir, "vision_tiny.pt"), weights_only=False)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_oracle_matches_golden(golden_dir, prec):
    c = _case(golden_dir)
    dt = torch.float32 if prec == "fp32" else torch.bfloat16
    cfg = OV.VisionConfig(**c["config"])
    w = {k: v.to(dt) for k, v in c["weights"].items()}
    imgs = [i.to(dt) for i in c["images"]]
    x, seqlens, grids = OV.patch_embedding(w, imgs, c["patch"])
    assert seqlens == c[prec]["seqlens"] and grids == c[prec]["grids"]
    assert torch.equal(x, c[prec]["patch_embedding"])
    feats = OV.eva2clip(w, imgs, c["patch"], c["pool"], cfg)
    for got, want in zip(feats, c[prec]["features"]):
        assert got.shape == want.shape
        # same ATen kernels on the same values: bit-exact on the machine that wrote the fixture, 1e-4 elsewhere
        torch.testing.assert_close(got.float(), want.float(), rtol=1e-4 if prec == "fp32" else 2e-2,
                                   atol=1e-5 if prec == "fp32" else 2e-2)


def test_golden_shapes_known_answers(golden_dir):
    c = _case(golden_dir)
    # (D/pd, H/ph, W/pw) grids, 1 + n tokens per image, pooled feature counts + boi/eoi
    assert c["fp32"]["grids"] == [(1, 4, 4), (1, 3, 5), (2, 4, 4), (4, 8, 4)]
    assert c["fp32"]["seqlens"] == [17, 16, 33, 129]
    assert [f.shape[1] for f in c["fp32"]["features"]] == [16 + 2, 15 + 2, 4 + 2, 32 + 2]


@pytest.mark.skipif(not RL.reference_available(), reason="needs /root/reference")
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_oracle_equals_live_reference(dt):
    cfg = OV.VisionConfig(hidden_size=256, num_heads=2, intermediate_size=384, num_hidden_layers=2,
                          patch_size=(4, 8, 8), pos_embed_shape=(2, 4, 4), lm_hidden_size=256, lm_intermediate_size=384)
    w = OV.random_vision_weights(cfg, seed=3)
    model = RL.make_reference_vision(cfg, w, dtype=dt)
    imgs = [i.to(dt) for i in OV.random_images([(4, 32, 32), (2, 16, 48), (8, 40, 32)], seed=5)]
    ps, pool = [(4, 8, 8), (2, 8, 8), (4, 8, 8)], [(1, 1, 1), (1, 2, 2), (2, 1, 2)]
    with torch.no_grad():
        ref = model(imgs, ps, pool)
    got = OV.eva2clip({k: v.to(dt) for k, v in w.items()}, imgs, ps, pool, cfg)
    for a, b in zip(ref, got):
        assert torch.equal(a, b)
    # the feature scatter of CogVLMModel.forward (modeling_cogvlm.py:450-453), restated
    emb = torch.randn(3, 40, 256).to(dt)
    want = emb.clone()
    for i, f in enumerate(ref):
        want[i, 1:1 + f.shape[1]] = f[0]
    assert torch.equal(OV.scatter_image_features(emb, got), want)


@pytest.mark.skipif(not RL.reference_available(), reason="needs /root/reference")
def test_oracle_equals_live_reference_head_dim_112():
    """EVA2-CLIP-E geometry: 1792 = 16 heads x 112 (one layer, narrow MLP to stay small)."""
    cfg = OV.VisionConfig(hidden_size=1792, num_heads=16, intermediate_size=256, num_hidden_layers=1,
                          patch_size=(1, 14, 14), pos_embed_shape=(1, 3, 3), lm_hidden_size=256, lm_intermediate_size=256)
    w = OV.random_vision_weights(cfg, seed=4)
    model = RL.make_reference_vision(cfg, w)
    imgs = OV.random_images([(1, 42, 42), (1, 28, 56)], seed=6)
    ps, pool = [(1, 14, 14)] * 2, [(1, 1, 1)] * 2
    with torch.no_grad():
        ref = model(imgs, ps, pool)
    for a, b in zip(ref, OV.eva2clip(w, imgs, ps, pool, cfg)):
        assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------ product host logic
def _tiny_model(**over):
    from types import SimpleNamespace
    from mmmm_b200.visual import EVA2CLIPModel
    vc = dict(hidden_size=256, num_heads=2, intermediate_size=256, num_hidden_layers=2, layer_norm_eps=1e-6,
              in_channels=3, patch_size=(4, 8, 8), pos_embed_shape=(2, 4, 4), hidden_act="gelu")
    vc.update(over)
    return EVA2CLIPModel(SimpleNamespace(hidden_size=256, intermediate_size=256, vision_config=vc))


def test_state_dict_keys_match_reference_layout(golden_dir):
    model = _tiny_model()
    keys = set(model.state_dict().keys())
    assert keys == set(_case(golden_dir)["weights"].keys())  # the fixture's keys are the reference's state dict
    if RL.reference_available():
        cfg = OV.VisionConfig(hidden_size=256, num_heads=2, intermediate_size=256, num_hidden_layers=2,
                              patch_size=(4, 8, 8), pos_embed_shape=(2, 4, 4), lm_hidden_size=256,
                              lm_intermediate_size=256)
        ref = RL.make_reference_vision(cfg, OV.random_vision_weights(cfg, seed=1))
        ref_sd = ref.state_dict()
        assert keys == set(ref_sd.keys())
        ours = model.state_dict()
        assert all(ours[k].shape == ref_sd[k].shape for k in keys)
    # bare-parameter spelling of the ParameterWrapper children (mmmm/utils.py:71-77) loads too
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd["patch_embedding.cls_embedding"] = sd.pop("patch_embedding.cls_embedding.weight") + 1
    model.load_state_dict(sd)
    assert float(model.patch_embedding.cls_embedding.weight[0, 0]) == 1.0


def test_plan_matches_oracle_packing(golden_dir):
    from mmmm_b200.visual import VisionPlan
    c = _case(golden_dir)
    shapes = [(3, *s) for s in c["shapes"]]
    plan = VisionPlan(shapes, c["patch"], c["pool"], "cpu")
    assert [1 + im.n for im in plan.images] == c["fp32"]["seqlens"]
    assert [im.grid for im in plan.images] == c["fp32"]["grids"]
    assert [im.m + 2 for im in plan.images] == [f.shape[1] for f in c["fp32"]["features"]]
    assert plan.cu_seqlens.tolist() == [0, 17, 33, 66, 195] and plan.T == 195 and plan.max_len == 129
    # every non-class packed row is the target of exactly one patch row; class rows of none
    hit = torch.zeros(plan.T, dtype=torch.int32)
    for g in plan.groups.values():
        assert g["count"][0] == g["rows"] == g["row_map"].numel()
        hit[g["row_map"].long()] += 1
    cls = torch.tensor([im.start for im in plan.images])
    assert hit[cls].sum() == 0 and hit.sum() == plan.T - plan.B and hit.max() == 1
    # three distinct patch sizes -> three GEMM groups; images 0 and 2 share one
    assert sorted(len(g["images"]) for g in plan.groups.values()) == [1, 1, 2]
    fmap = plan.feature_row_map([100, 200, 300, 400], "cpu")
    assert fmap.numel() == plan.M == 16 + 15 + 4 + 32 and fmap[:16].tolist() == list(range(100, 116))


def test_head_slot_padding_preserves_attention():
    """Zero-padded 128-wide head slots give the same attention + dense output as the 112-wide heads."""
    from mmmm_b200.visual import _pad_heads_cols, _pad_heads_rows
    g = torch.Generator().manual_seed(0)
    heads, hd, C, T = 2, 112, 224, 9
    x = torch.randn(1, T, C, generator=g)
    w = {"a.query_key_value.weight": torch.randn(3 * C, C, generator=g) * 0.05,
         "a.query_key_value.bias": torch.randn(3 * C, generator=g) * 0.05,
         "a.dense.weight": torch.randn(C, C, generator=g) * 0.05, "a.dense.bias": torch.randn(C, generator=g) * 0.05}
    want = OV.attention(w, "a.", x, [4, 5], heads)
    wq, bq = _pad_heads_rows(w["a.query_key_value.weight"], heads, hd), _pad_heads_rows(w["a.query_key_value.bias"], heads, hd)
    wd = _pad_heads_cols(w["a.dense.weight"], heads, hd)
    assert wq.shape == (3 * heads * 128, C) and wd.shape == (C, heads * 128)
    qkv = torch.nn.functional.linear(x, wq, bq).reshape(1, T, 3, heads, 128).permute(2, 0, 1, 3, 4)
    ctx = OV.blockdiag_attention(qkv[0], qkv[1], qkv[2], [4, 5], hd ** -0.5)
    got = torch.nn.functional.linear(ctx.reshape(1, T, -1), wd, w["a.dense.bias"])
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)


def test_cpu_tensors_raise():
    model = _tiny_model().to(torch.bfloat16)
    img = [torch.zeros(3, 4, 32, 32, dtype=torch.bfloat16)]
    with torch.no_grad(), pytest.raises(ValueError, match="CUDA"):
        model(img, [(4, 8, 8)], [(1, 1, 1)])
    with pytest.raises(NotImplementedError, match="autograd"):
        model(img, [(4, 8, 8)], [(1, 1, 1)])
    with pytest.raises(ValueError, match="head_dim"):
        _tiny_model(hidden_size=512, num_heads=2)


def test_patch_embedding_checkpoint_inflation_matches_reference():
    """PatchEmbedding._load_from_state_dict (visual.py:38-57): a pretrained 2-D EVA2-CLIP position embedding
    [1 + h*w, C] is split / resampled / repeated along depth on load.  Checked against the live reference module when
    /root/reference exists (same state dict in, same parameters out), and structurally everywhere."""
    g = torch.Generator().manual_seed(3)
    C = 256
    pt = torch.randn(1 + 3 * 3, C, generator=g)                    # pretrained 3 x 3 grid + class row
    model = _tiny_model(pt_pos_embed_shape=(3, 3))
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd["patch_embedding.position_embedding.weight"] = pt.clone()
    del sd["patch_embedding.cls_pos_embed.weight"]
    model.load_state_dict(sd)
    pe = model.patch_embedding.position_embedding.weight
    assert pe.shape == (1, C, 2, 4, 4)
    assert torch.equal(model.patch_embedding.cls_pos_embed.weight, pt[0:1])
    assert torch.equal(pe[:, :, 0], pe[:, :, 1])                    # repeated along depth
    # same grid: no resampling, pure reshape
    same = _tiny_model(pt_pos_embed_shape=(4, 4))
    pt4 = torch.randn(1 + 16, C, generator=g)
    sd4 = {k: v.clone() for k, v in same.state_dict().items()}
    sd4["patch_embedding.position_embedding.weight"] = pt4.clone()
    del sd4["patch_embedding.cls_pos_embed.weight"]
    same.load_state_dict(sd4)
    assert torch.equal(same.patch_embedding.position_embedding.weight[0, :, 0], pt4[1:].reshape(4, 4, C).permute(2, 0, 1))
    if RL.reference_available():
        cfg = OV.VisionConfig(hidden_size=256, num_heads=2, intermediate_size=256, num_hidden_layers=2,
                              patch_size=(4, 8, 8), pos_embed_shape=(2, 4, 4), lm_hidden_size=256,
                              lm_intermediate_size=256)
        ref = RL.make_reference_vision(cfg, OV.random_vision_weights(cfg, seed=1))
        ref.patch_embedding.pt_pos_embed_shape = (3, 3)
        rsd = {k: v.clone() for k, v in ref.state_dict().items()}
        rsd["patch_embedding.position_embedding.weight"] = pt.clone()
        del rsd["patch_embedding.cls_pos_embed.weight"]
        ref.load_state_dict(rsd)
        assert torch.equal(ref.patch_embedding.position_embedding.weight, pe)
        assert torch.equal(ref.patch_embedding.cls_pos_embed.weight, model.patch_embedding.cls_pos_embed.weight)
