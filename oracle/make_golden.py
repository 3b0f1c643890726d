"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.pt from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden

The reference has no golden vectors of its own for this path, so the
fixtures are outputs of the reference code itself, loaded through ``oracle/reference_loader.py``.
Weights and activations are rounded to bf16-representable values so that the same fixture serves the
fp32 and the bf16 checks.  Everything is CPU and seeded; the files are small (a few MB).
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import reference_loader as RL  # noqa: E402
from mmmm_b200.inputs import make_ids  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def routing_vectors(M):
    """Known-answer routing cases run through the real ``get_expert_mask`` (modeling_cogvlm.py:58-70)."""
    cases = []
    tt = torch.tensor([[0, 1, 1, 1, 1, 0, 0, 0], [0, 1, 1, 1, 0, 0, 0, 0], [1, 1, 0, 1, 0, 1, 1, 1]])
    pm = torch.tensor([[1] * 8, [1, 1, 1, 1, 1, 1, 0, 0], [1, 1, 1, 1, 1, 1, 1, 0]], dtype=torch.bool)
    cases.append((tt, pm))
    # L == 1: padding ignored
    cases.append((torch.tensor([[1], [0]]), torch.tensor([[False], [True]])))
    # non-contiguous (arbitrary) padding pattern, all-vision row, all-padded row
    g = torch.Generator().manual_seed(7)
    tt = (torch.rand(5, 37, generator=g) < 0.6).long()
    pm = torch.rand(5, 37, generator=g) < 0.8
    tt[3] = 1
    pm[4] = False
    cases.append((tt, pm))
    # realistic layout, ragged
    tt, _, pm = make_ids(4, 20, 9, ragged=True, seed=3)
    cases.append((tt, pm))
    out = []
    for tt, pm in cases:
        v, l = M.get_expert_mask(tt, pm)
        out.append(dict(token_type_ids=tt, padding_mask=pm, vision=v, language=l))
    return out


def layer_case(M, *, hidden, heads, inter, batch, nv, nt, seed, with_lora=False):
    layer, cfg = RL.make_reference_layer(hidden, inter, heads, dtype=torch.float32, seed=seed)
    # bf16-representable weights so fp32 and bf16 runs share the fixture
    with torch.no_grad():
        for p in layer.parameters():
            p.copy_(p.to(torch.bfloat16).float())
    tt, pos, pm = make_ids(batch, nv, nt, ragged=True, seed=seed)
    g = torch.Generator().manual_seed(seed + 100)
    h = torch.randn(batch, tt.shape[1], hidden, generator=g).to(torch.bfloat16).float()
    weights = {k: v.detach().clone() for k, v in layer.state_dict().items()}

    def run(dtype):
        lay = layer.to(dtype)
        lay.self_attn.rotary_emb.max_seq_len_cached = 0  # rebuild the table in this dtype (quirk 2)
        with torch.no_grad():
            out, present = lay(h.to(dtype), token_type_ids=tt, position_ids=pos, padding_mask=pm, use_cache=True)
        cos = lay.self_attn.rotary_emb.cos_cached[:, 0].clone()
        sin = lay.self_attn.rotary_emb.sin_cached[:, 0].clone()
        return dict(out=out.clone(), k=present[0].clone(), v=present[1].clone(), cos=cos, sin=sin)

    res32 = run(torch.float32)
    res16 = run(torch.bfloat16)
    layer.to(torch.float32)
    # padded rows of the reference output are torch.empty garbage (modeling_cogvlm.py:277): blank them
    for r in (res32, res16):
        r["out"][~pm] = 0
    return dict(
        config=dict(hidden_size=hidden, num_heads=heads, intermediate_size=inter, rms_norm_eps=cfg.rms_norm_eps),
        weights={k: v.to(torch.bfloat16) for k, v in weights.items() if "inv_freq" not in k},
        inv_freq=weights["self_attn.rotary_emb.inv_freq"],
        hidden_states=h.to(torch.bfloat16), token_type_ids=tt, position_ids=pos, padding_mask=pm,
        fp32=res32, bf16=res16,
    )


VISION_TINY = dict(hidden_size=256, num_heads=2, intermediate_size=256, num_hidden_layers=2, patch_size=(4, 8, 8),
                   pos_embed_shape=(2, 4, 4), lm_hidden_size=256, lm_intermediate_size=256)
VISION_TINY_IMAGES = dict(shapes=[(4, 32, 32), (1, 24, 40), (8, 32, 32), (8, 64, 32)],
                          patch=[(4, 8, 8), (1, 8, 8), (4, 8, 8), (2, 8, 8)],
                          pool=[(1, 1, 1), (1, 1, 1), (2, 2, 2), (2, 2, 1)])


def vision_case():
    """EVA2CLIPModel.forward (visual.py:191-208) of the UNMODIFIED reference on a tiny config: four images with
    different depths / patch depths (depth-reduced kernels, resampled position embeddings) and pool sizes."""
    from oracle import oracle_vision as OV
    cfg = OV.VisionConfig(**VISION_TINY)
    w = OV.random_vision_weights(cfg, seed=21)
    imgs = OV.random_images(VISION_TINY_IMAGES["shapes"], seed=22)
    model = RL.make_reference_vision(cfg, w)
    res = {}
    for prec, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        m = model.to(dt)
        with torch.no_grad():
            x, mask, shapes = m.patch_embedding([i.to(dt) for i in imgs], VISION_TINY_IMAGES["patch"])
            feats = m([i.to(dt) for i in imgs], VISION_TINY_IMAGES["patch"], VISION_TINY_IMAGES["pool"])
        res[prec] = dict(patch_embedding=x.clone(), seqlens=list(mask.seqlens), grids=[tuple(s) for s in shapes],
                         features=[f.clone() for f in feats])
    return dict(config=VISION_TINY, weights={k: v.to(torch.bfloat16) for k, v in w.items()},
                images=[i.to(torch.bfloat16) for i in imgs], **VISION_TINY_IMAGES, **res)


def main():
    os.makedirs(OUT, exist_ok=True)
    M = RL.load_reference()
    torch.manual_seed(0)
    misc = dict(
        routing=routing_vectors(M),
        build_position_ids=dict(x=torch.tensor([[0, 1, 1, 1, 1, 0, 0, 0]]),
                                y=M.build_position_ids(torch.tensor([[0, 1, 1, 1, 1, 0, 0, 0]]))),
        bf16_arange_250_270=torch.arange(270, dtype=torch.bfloat16)[250:270].float(),
    )
    # counts at config 1 (1225 vision + 128 text): SURVEY 8(c)(2)
    tt, _, pm = make_ids(1, 1225, 128)
    v, l = M.get_expert_mask(tt, pm)
    misc["c1_counts"] = dict(vision=int(v.sum()), language=int(l.sum()), total=int(pm.sum()))
    torch.save(misc, os.path.join(OUT, "routing.pt"))

    tiny = layer_case(M, hidden=256, heads=2, inter=384, batch=3, nv=20, nt=14, seed=11)
    torch.save(tiny, os.path.join(OUT, "layer_tiny.pt"))
    # long positions (> 256) to exercise the bf16 arange collapse of the rotary table
    longp = layer_case(M, hidden=256, heads=2, inter=256, batch=2, nv=8, nt=300, seed=12)
    torch.save(longp, os.path.join(OUT, "layer_longpos.pt"))
    # the ragged text lengths of layer_longpos top out at position 253; this one is ragged over [300, 600] text tokens,
    # so every sample runs through the collapsed region of the bf16-built table (positions 257 .. ~600: bf16 arange
    # steps of 2, then 4 above 512) -- the rotary quirk pinned at LAYER level by reference output
    pos600 = layer_case(M, hidden=256, heads=2, inter=256, batch=2, nv=8, nt=600, seed=17)
    assert int(pos600["position_ids"].max()) >= 512
    torch.save(pos600, os.path.join(OUT, "layer_pos600.pt"))
    # lm_head + _sample_weighted_ce (modeling_cogvlm.py:610-627, :701-706) through the unmodified reference function
    g = torch.Generator().manual_seed(13)
    Bc, Lc, Hc, Vc = 3, 37, 64, 333
    hid = torch.randn(Bc, Lc, Hc, generator=g).bfloat16()
    wlm = (torch.randn(Vc, Hc, generator=g) * 0.3).bfloat16()
    labels = torch.randint(0, Vc, (Bc, Lc), generator=g)
    labels[torch.rand(Bc, Lc, generator=g) < 0.6] = -100
    wt = (0.5 + torch.rand(Bc, Lc, generator=g)).bfloat16()
    ce = {}
    for prec, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        logits = torch.nn.functional.linear(hid.to(dt), wlm.to(dt)).float()
        ce[prec] = dict(weighted=M._sample_weighted_ce(logits, labels, wt), plain=M._sample_weighted_ce(logits, labels, None))
    torch.save(dict(hidden_states=hid, lm_head_weight=wlm, labels=labels, weight=wt, loss=ce),
               os.path.join(OUT, "lm_head_ce.pt"))
    torch.save(vision_case(), os.path.join(OUT, "vision_tiny.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
