"""CPU, build container only: the oracle against the reference's own code EXECUTED IN PLACE from /root/reference
(AST-extracted, never copied), and against transformers.CLIPVisionModel (the graph the reference's training twin uses,
training/image_encoder.py:39,68). Skipped where /root/reference is absent (the GPU box)."""
import ast
import math
import os

import numpy as np
import pytest
import torch

REF = os.environ.get("DTP_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "trt_inference")),
                                reason="reference tree not present")


def extract(path, names):
    tree = ast.parse(open(path).read())
    ns = {"np": np, "torch": torch, "math": math}
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns


@pytest.mark.parametrize("S", [3, 7, 20, 30])
def test_ddim_against_reference_class(S):
    from oracle.ddim import DDIM
    ns = extract(os.path.join(REF, "trt_inference/utilities.py"), {"DDIMScheduler"})
    ref = ns["DDIMScheduler"](device="cpu", num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012)
    ref.set_timesteps(S)
    ref.configure()
    o = DDIM()
    o.set_timesteps(S)
    assert torch.equal(ref.timesteps, o.timesteps)
    assert torch.equal(ref.alphas_cumprod, o.alphas_cumprod)
    g = torch.Generator().manual_seed(S)
    x, eps = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    for idx in range(S):
        assert torch.equal(ref.step(eps, x, idx, ref.timesteps[idx]), o.step(eps, x, idx))


def test_pos_encoding_and_patches_against_reference():
    from oracle import image_encoder as ie
    ns = extract(os.path.join(REF, "trt_inference/image_encoder.py"), {"positional_encoding_2d", "get_image_patches"})
    for hid, hw in ((768, 3), (768, 2), (128, 3), (64, 1)):
        assert torch.equal(ns["positional_encoding_2d"](hid, hw, hw), ie.positional_encoding_2d(hid, hw, hw))
    img = torch.rand(1, 3, 224, 224)
    mine = ie.preprocess_patches(img * 0 + img)  # 224 input: no bicubic step
    mean = torch.tensor(ie.CLIP_MEAN)[None, :, None, None]
    std = torch.tensor(ie.CLIP_STD)[None, :, None, None]
    norm = (img - mean) / std
    ref = []
    for ps in (224, 112, 74):
        p = ns["get_image_patches"](norm, ps)
        ref.append(torch.nn.functional.interpolate(p, size=(224, 224), mode="bilinear", align_corners=False)
                   if ps != 224 else p)
    assert torch.equal(torch.cat(ref), mine)


def test_add_extra_context_matches_unfold_dilation():
    """kornia is absent: restate its documented 'unfold' engine (geodesic border -1e4, origin k//2) literally and compare
    with the oracle's padded max-pool (handler.py:25-33)."""
    from oracle.pipeline import add_extra_context
    g = torch.Generator().manual_seed(0)
    mask = (torch.rand(2, 1, 24, 24, generator=g) > 0.9).float()
    src, masked = torch.rand(1, 3, 24, 24, generator=g), torch.rand(2, 3, 24, 24, generator=g)
    for pad in (1, 4, 7, 10):
        o = pad // 2
        padded = torch.nn.functional.pad(mask, (o, pad - o - 1, o, pad - o - 1), value=-1e4)
        win = padded.unfold(2, pad, 1).unfold(3, pad, 1)
        dil = win.reshape(2, 1, 24, 24, -1).max(-1).values
        hint = 1 - dil
        ref_img, ref_mask = masked + src * hint, torch.clamp(mask + hint, 0, 1)
        got_img, got_mask = add_extra_context(src, masked, mask, pad)
        assert torch.equal(ref_img, got_img) and torch.equal(ref_mask, got_mask)


def test_clip_visual_against_transformers():
    """openai-CLIP visual tower with proj=None == transformers CLIPVisionModel.pooler_output (random init, small)."""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    from diffusiontexturepainting_b200.weights import EncoderConfig
    from oracle import image_encoder as ie
    cfg = EncoderConfig(width=64, layers=2, heads=2, mlp=128)
    hf = CLIPVisionModel(CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                                          num_attention_heads=2, image_size=224, patch_size=32,
                                          hidden_act="quick_gelu")).eval()
    m = hf.vision_model
    sd = {"clip.visual.conv1.weight": m.embeddings.patch_embedding.weight,
          "clip.visual.class_embedding": m.embeddings.class_embedding,
          "clip.visual.positional_embedding": m.embeddings.position_embedding.weight,
          "clip.visual.ln_pre.weight": m.pre_layrnorm.weight, "clip.visual.ln_pre.bias": m.pre_layrnorm.bias,
          "clip.visual.ln_post.weight": m.post_layernorm.weight, "clip.visual.ln_post.bias": m.post_layernorm.bias}
    for i, layer in enumerate(m.encoder.layers):
        b = f"clip.visual.transformer.resblocks.{i}"
        a = layer.self_attn
        sd[f"{b}.attn.in_proj_weight"] = torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight])
        sd[f"{b}.attn.in_proj_bias"] = torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias])
        sd[f"{b}.attn.out_proj.weight"], sd[f"{b}.attn.out_proj.bias"] = a.out_proj.weight, a.out_proj.bias
        sd[f"{b}.ln_1.weight"], sd[f"{b}.ln_1.bias"] = layer.layer_norm1.weight, layer.layer_norm1.bias
        sd[f"{b}.ln_2.weight"], sd[f"{b}.ln_2.bias"] = layer.layer_norm2.weight, layer.layer_norm2.bias
        sd[f"{b}.mlp.c_fc.weight"], sd[f"{b}.mlp.c_fc.bias"] = layer.mlp.fc1.weight, layer.mlp.fc1.bias
        sd[f"{b}.mlp.c_proj.weight"], sd[f"{b}.mlp.c_proj.bias"] = layer.mlp.fc2.weight, layer.mlp.fc2.bias
    sd = {k: v.detach() for k, v in sd.items()}
    x = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        ref = hf(pixel_values=x).pooler_output
        got = ie.clip_visual(sd, cfg, x)
    assert torch.allclose(ref, got, atol=2e-5, rtol=1e-4), (ref - got).abs().max()
