"""CPU tests of the real-checkpoint ingestion: key spellings of the files the reference reads
(diffusers VAE attention old / new names, `unet.`-prefixed LoRA factors, the training-side image_encoder.pth whose CLIP tower
is a transformers.CLIPVisionModel) are mapped onto the inventory with every shape checked."""
import os

import pytest
import torch

from diffusiontexturepainting_b200 import checkpoints as ck
from diffusiontexturepainting_b200 import weights as W


def tiny(seed):
    return W.synth_model(W.tiny_config(), seed)


def to_new_vae_spelling(sd):
    out = {}
    for k, v in sd.items():
        for old, new in (("query", "to_q"), ("key", "to_k"), ("value", "to_v"), ("proj_attn", "to_out.0")):
            if f".attentions.0.{old}." in k:
                k = k.replace(f".{old}.", f".{new}.")
                if k.endswith(".weight"):
                    v = v[:, :, None, None]  # exported as a 1x1 convolution
                break
        out[k] = v
    return out


def to_hf_clip_spelling(enc_sd):
    """Inverse of hf_clip_vision_to_openai: what the training-side ConditionPatchEncoder.state_dict() looks like."""
    out = {}
    inv = {"class_embedding": "embeddings.class_embedding", "conv1.weight": "embeddings.patch_embedding.weight",
           "positional_embedding": "embeddings.position_embedding.weight", "ln_pre.weight": "pre_layrnorm.weight",
           "ln_pre.bias": "pre_layrnorm.bias", "ln_post.weight": "post_layernorm.weight", "ln_post.bias": "post_layernorm.bias"}
    lay = {"ln_1": "layer_norm1", "ln_2": "layer_norm2", "attn.out_proj": "self_attn.out_proj", "mlp.c_fc": "mlp.fc1",
           "mlp.c_proj": "mlp.fc2"}
    for k, v in enc_sd.items():
        if not k.startswith("clip.visual."):
            out[k] = v
            continue
        r = k[len("clip.visual."):]
        if r in inv:
            out["clip.vision_model." + inv[r]] = v
            continue
        _, _, i, rest = r.split(".", 3)
        base = f"clip.vision_model.encoder.layers.{i}."
        if rest.startswith("attn.in_proj_"):
            wb = rest.split("_")[-1]
            q, kk, vv = v.chunk(3, 0)
            out[base + f"self_attn.q_proj.{wb}"], out[base + f"self_attn.k_proj.{wb}"] = q, kk
            out[base + f"self_attn.v_proj.{wb}"] = vv
        else:
            name, wb = rest.rsplit(".", 1)
            out[base + f"{lay[name]}.{wb}"] = v
    out["clip.vision_model.embeddings.position_ids"] = torch.arange(50)[None]
    out["pos_emb"] = torch.zeros(1, 14, 8)  # non-persistent buffers that older torch versions still wrote
    return out


def write_files(tmp, u, v, e):
    from safetensors.torch import save_file
    os.makedirs(tmp / "unet"), os.makedirs(tmp / "vae")
    base = {k: t.contiguous() for k, t in u.items() if ".processor." not in k}
    save_file(base, str(tmp / "unet" / "diffusion_pytorch_model.safetensors"))
    torch.save(to_new_vae_spelling(v), str(tmp / "vae" / "diffusion_pytorch_model.bin"))
    torch.save({"unet." + k: t for k, t in u.items() if ".processor." in k}, str(tmp / "lora.bin"))
    torch.save(to_hf_clip_spelling(e), str(tmp / "image_encoder.pth"))


def test_real_checkpoint_spellings_load_completely(tmp_path):
    u0, v0, e0 = tiny(1)          # what the engine starts from (synthetic)
    u1, v1, e1 = tiny(2)          # "real" weights, written in the files' spellings
    write_files(tmp_path, u1, v1, e1)
    reps = ck.load_real_checkpoints(u0, v0, e0, str(tmp_path), str(tmp_path / "lora.bin"),
                                    str(tmp_path / "image_encoder.pth"), strict=True)
    assert len(reps) == 4 and all(r.ok for r in reps), [r.summary() for r in reps]
    for got, want in ((u0, u1), (v0, v1), (e0, e1)):
        assert got.keys() == want.keys()
        for k in want:
            assert torch.equal(got[k], want[k].float()), k
    # unexpected keys are reported, not fatal (position_ids is dropped silently, pos_emb filtered)
    assert reps[3].unexpected == []
    # LoRA merged == the reference's W + up @ down on the loaded factors
    merged = W.merge_lora(u0)
    k = next(k for k in u1 if k.endswith("attn1.processor.to_q_lora.down.weight"))
    base = k.split(".processor.")[0]
    want = u1[base + ".to_q.weight"] + u1[k.replace(".down.", ".up.")] @ u1[k]
    assert torch.allclose(merged[base + ".to_q.weight"], want)


def test_hf_clip_mapping_matches_transformers_module_keys():
    """The mapping is exercised on the state dict of an actual transformers.CLIPVisionModel (random init, small)."""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    hf = CLIPVisionModel(CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                                          image_size=224, patch_size=32, hidden_act="quick_gelu"))
    sd = {"clip." + k: v for k, v in hf.state_dict().items()}
    out = ck.hf_clip_vision_to_openai(sd)
    cfg = W.EncoderConfig(width=64, layers=2, heads=2, mlp=128)
    want = {k: s for k, s in W.encoder_param_shapes(cfg).items() if k.startswith("clip.visual.")}
    assert set(out) == set(want)
    for k, shape in want.items():
        assert tuple(out[k].shape) == tuple(shape), k
    a = hf.vision_model.encoder.layers[1].self_attn
    assert torch.equal(out["clip.visual.transformer.resblocks.1.attn.in_proj_weight"],
                       torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight]))


def test_strict_mode_rejects_partial_or_misshapen_files(tmp_path):
    u0, v0, e0 = tiny(1)
    u1, v1, e1 = tiny(2)
    write_files(tmp_path, u1, v1, e1)
    # drop one VAE tensor and break one shape
    p = str(tmp_path / "vae" / "diffusion_pytorch_model.bin")
    sd = torch.load(p)
    victim = next(k for k in sd if k.endswith("conv_in.weight"))
    del sd[victim]
    bad = next(k for k in sd if k.endswith("conv_out.bias"))
    sd[bad] = torch.zeros(sd[bad].numel() + 1)
    torch.save(sd, p)
    with pytest.raises(ck.CheckpointError, match="missing|mismatch"):
        ck.load_real_checkpoints(u0, v0, e0, str(tmp_path), None, None, strict=True)
    u0, v0, e0 = tiny(1)
    reps = ck.load_real_checkpoints(u0, v0, e0, str(tmp_path), None, None, strict=False)
    vae = [r for r in reps if r.what.startswith("vae")][0]
    assert victim in vae.missing and vae.mismatched[0][0] == bad and not vae.ok
    assert torch.equal(v0[victim], tiny(1)[1][victim])      # untouched inventory tensor keeps its synthetic value


def test_absent_files_leave_the_inventory_untouched(tmp_path):
    u0, v0, e0 = tiny(1)
    assert ck.load_real_checkpoints(u0, v0, e0, str(tmp_path / "nowhere"), "/no/lora.bin", "/no/enc.pth") == []
    u1, _, _ = tiny(1)
    assert all(torch.equal(u0[k], u1[k]) for k in u1)


def test_missing_checkpoints_raise_unless_synthetic_weights_are_requested(tmp_path, monkeypatch, capsys):
    """the default server construction must not silently serve seeded random weights (the reference crashes in
    from_pretrained / torch.load when a file is absent, models.py:796-813, trt_model.py:58)."""
    from diffusiontexturepainting_b200 import weights as W
    from diffusiontexturepainting_b200.checkpoints import CheckpointError
    from diffusiontexturepainting_b200.stable_diffusion_pipeline import load_state_dicts
    monkeypatch.setenv("DTP_HF_DIR", str(tmp_path / "nothing"))
    monkeypatch.setenv("DTP_IMAGE_ENCODER", str(tmp_path / "none.pth"))
    monkeypatch.delenv("DTP_SYNTHETIC_WEIGHTS", raising=False)
    cfg = W.ModelConfig(unet=W.tiny_config().unet, vae=W.tiny_config().vae, enc=W.tiny_config().enc, name="sd15-inpaint")
    with pytest.raises(CheckpointError, match="DTP_SYNTHETIC_WEIGHTS"):
        load_state_dicts(cfg, str(tmp_path / "lora.bin"))
    monkeypatch.setenv("DTP_SYNTHETIC_WEIGHTS", "1")
    u, v, e = load_state_dicts(cfg, str(tmp_path / "lora.bin"))
    assert "SYNTHETIC" in capsys.readouterr().out and len(u) and len(v) and len(e)
    # a real UNet next to a missing LoRA file: the LoRA merge must be a no-op, not random
    (tmp_path / "hf" / "unet").mkdir(parents=True)
    real = {k: t for k, t in W.synth_state_dict(W.unet_param_shapes(cfg.unet), 99).items() if ".processor." not in k}
    torch.save(real, tmp_path / "hf" / "unet" / "diffusion_pytorch_model.bin")
    monkeypatch.setenv("DTP_HF_DIR", str(tmp_path / "hf"))
    u, _, _ = load_state_dicts(cfg, str(tmp_path / "lora.bin"))
    merged = W.merge_lora(u)
    k = next(k for k in real if k.endswith("attn1.to_q.weight"))
    assert torch.equal(merged[k], real[k])
    # the narrow test configuration is synthetic by definition
    monkeypatch.delenv("DTP_SYNTHETIC_WEIGHTS", raising=False)
    assert len(load_state_dicts(W.tiny_config())[0])
