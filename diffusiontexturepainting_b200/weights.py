"""Model configurations, parameter inventories (diffusers / openai-CLIP key names), seeded synthetic weights, LoRA merge
and packing of weights into the layouts the sm_100a kernels consume.

The reference loads `runwayml/stable-diffusion-inpainting` (diffusers 0.12.0 layout, trt_inference/models.py:795-813,
1036-1041), fuses rank-4 LoRA into to_q/to_k/to_v/to_out.0 of all attention modules (models.py:1042-1093) and loads
`image_encoder.pth` into ConditionPatchEncoder (trt_inference/trt_model.py:57-61). No checkpoints exist offline, so the
same key inventory is filled with seeded, variance-preserving random values; real checkpoints drop in by key name.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Tuple

import torch


# ----------------------------------------------------------------------------------------------------------------
# configurations
# ----------------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class UNetConfig:
    in_channels: int = 9
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    down_attention: Tuple[bool, ...] = (True, True, True, False)
    heads: int = 8  # diffusers `attention_head_dim=8` is the head COUNT for SD-1.x
    cross_dim: int = 768
    groups: int = 32
    lora_rank: int = 4

    @property
    def time_dim(self):
        return self.block_out_channels[0] * 4


@dataclass(frozen=True)
class VAEConfig:
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    latent_channels: int = 4
    groups: int = 32


@dataclass(frozen=True)
class EncoderConfig:
    width: int = 768  # CLIP ViT-B/32 visual tower
    layers: int = 12
    heads: int = 12
    mlp: int = 3072
    tower_layers: int = 4  # three patch towers of BasicTransformerBlock(hid, 4 heads, gelu, attention_bias)
    tower_heads: int = 4
    cross_dim: int = 768
    num_patches: Tuple[int, ...] = (1, 4, 9)


@dataclass(frozen=True)
class ModelConfig:
    unet: UNetConfig = field(default_factory=UNetConfig)
    vae: VAEConfig = field(default_factory=VAEConfig)
    enc: EncoderConfig = field(default_factory=EncoderConfig)
    name: str = "sd15-inpaint"


def sd15_config() -> ModelConfig:
    return ModelConfig()


def tiny_config() -> ModelConfig:
    """Same graph, narrow channels: for CPU-side tests and golden vectors (not a benchmark configuration)."""
    return ModelConfig(
        unet=UNetConfig(block_out_channels=(64, 128, 256, 256), heads=4, cross_dim=128),
        vae=VAEConfig(block_out_channels=(64, 64, 128, 128)),
        enc=EncoderConfig(width=128, layers=2, heads=2, mlp=256, tower_layers=2, tower_heads=4, cross_dim=128),
        name="tiny",
    )


# ----------------------------------------------------------------------------------------------------------------
# parameter inventories
# ----------------------------------------------------------------------------------------------------------------
def _resnet(shapes, p, cin, cout, temb):
    shapes[f"{p}.norm1.weight"] = (cin,)
    shapes[f"{p}.norm1.bias"] = (cin,)
    shapes[f"{p}.conv1.weight"] = (cout, cin, 3, 3)
    shapes[f"{p}.conv1.bias"] = (cout,)
    if temb:
        shapes[f"{p}.time_emb_proj.weight"] = (cout, temb)
        shapes[f"{p}.time_emb_proj.bias"] = (cout,)
    shapes[f"{p}.norm2.weight"] = (cout,)
    shapes[f"{p}.norm2.bias"] = (cout,)
    shapes[f"{p}.conv2.weight"] = (cout, cout, 3, 3)
    shapes[f"{p}.conv2.bias"] = (cout,)
    if cin != cout:
        shapes[f"{p}.conv_shortcut.weight"] = (cout, cin, 1, 1)
        shapes[f"{p}.conv_shortcut.bias"] = (cout,)


def _transformer2d(shapes, p, c, cross, rank):
    shapes[f"{p}.norm.weight"] = (c,)
    shapes[f"{p}.norm.bias"] = (c,)
    shapes[f"{p}.proj_in.weight"] = (c, c, 1, 1)
    shapes[f"{p}.proj_in.bias"] = (c,)
    b = f"{p}.transformer_blocks.0"
    for n in ("norm1", "norm2", "norm3"):
        shapes[f"{b}.{n}.weight"] = (c,)
        shapes[f"{b}.{n}.bias"] = (c,)
    for attn, kv in (("attn1", c), ("attn2", cross)):
        shapes[f"{b}.{attn}.to_q.weight"] = (c, c)
        shapes[f"{b}.{attn}.to_k.weight"] = (c, kv)
        shapes[f"{b}.{attn}.to_v.weight"] = (c, kv)
        shapes[f"{b}.{attn}.to_out.0.weight"] = (c, c)
        shapes[f"{b}.{attn}.to_out.0.bias"] = (c,)
        if rank:
            for nm, cin in (("q", c), ("k", kv), ("v", kv), ("out", c)):
                shapes[f"{b}.{attn}.processor.to_{nm}_lora.down.weight"] = (rank, cin)
                shapes[f"{b}.{attn}.processor.to_{nm}_lora.up.weight"] = (c, rank)
    shapes[f"{b}.ff.net.0.proj.weight"] = (8 * c, c)
    shapes[f"{b}.ff.net.0.proj.bias"] = (8 * c,)
    shapes[f"{b}.ff.net.2.weight"] = (c, 4 * c)
    shapes[f"{b}.ff.net.2.bias"] = (c,)
    shapes[f"{p}.proj_out.weight"] = (c, c, 1, 1)
    shapes[f"{p}.proj_out.bias"] = (c,)


def unet_param_shapes(cfg: UNetConfig) -> Dict[str, tuple]:
    """diffusers-0.12.0 UNet2DConditionModel state-dict keys (+ LoRAAttnProcessor keys after load_attn_procs)."""
    s: Dict[str, tuple] = {}
    ch = cfg.block_out_channels
    T = cfg.time_dim
    s["conv_in.weight"] = (ch[0], cfg.in_channels, 3, 3)
    s["conv_in.bias"] = (ch[0],)
    s["time_embedding.linear_1.weight"] = (T, ch[0])
    s["time_embedding.linear_1.bias"] = (T,)
    s["time_embedding.linear_2.weight"] = (T, T)
    s["time_embedding.linear_2.bias"] = (T,)
    skip = [ch[0]]
    cin = ch[0]
    for i, cout in enumerate(ch):
        for j in range(cfg.layers_per_block):
            _resnet(s, f"down_blocks.{i}.resnets.{j}", cin, cout, T)
            if cfg.down_attention[i]:
                _transformer2d(s, f"down_blocks.{i}.attentions.{j}", cout, cfg.cross_dim, cfg.lora_rank)
            cin = cout
            skip.append(cout)
        if i != len(ch) - 1:
            s[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            s[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (cout,)
            skip.append(cout)
    mid = ch[-1]
    _resnet(s, "mid_block.resnets.0", mid, mid, T)
    _transformer2d(s, "mid_block.attentions.0", mid, cfg.cross_dim, cfg.lora_rank)
    _resnet(s, "mid_block.resnets.1", mid, mid, T)
    rev = list(reversed(ch))
    up_attn = list(reversed(cfg.down_attention))
    cin = mid
    for i, cout in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            sk = skip.pop()
            _resnet(s, f"up_blocks.{i}.resnets.{j}", cin + sk, cout, T)
            if up_attn[i]:
                _transformer2d(s, f"up_blocks.{i}.attentions.{j}", cout, cfg.cross_dim, cfg.lora_rank)
            cin = cout
        if i != len(rev) - 1:
            s[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            s[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (cout,)
    s["conv_norm_out.weight"] = (ch[0],)
    s["conv_norm_out.bias"] = (ch[0],)
    s["conv_out.weight"] = (cfg.out_channels, ch[0], 3, 3)
    s["conv_out.bias"] = (cfg.out_channels,)
    return s


def _vae_attn(s, p, c):
    s[f"{p}.group_norm.weight"] = (c,)
    s[f"{p}.group_norm.bias"] = (c,)
    for n in ("query", "key", "value", "proj_attn"):
        s[f"{p}.{n}.weight"] = (c, c)
        s[f"{p}.{n}.bias"] = (c,)


def vae_param_shapes(cfg: VAEConfig) -> Dict[str, tuple]:
    """diffusers-0.12.0 AutoencoderKL state-dict keys."""
    s: Dict[str, tuple] = {}
    ch = cfg.block_out_channels
    L = cfg.latent_channels
    s["encoder.conv_in.weight"] = (ch[0], 3, 3, 3)
    s["encoder.conv_in.bias"] = (ch[0],)
    cin = ch[0]
    for i, cout in enumerate(ch):
        for j in range(cfg.layers_per_block):
            _resnet(s, f"encoder.down_blocks.{i}.resnets.{j}", cin, cout, 0)
            cin = cout
        if i != len(ch) - 1:
            s[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            s[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"] = (cout,)
    c = ch[-1]
    _resnet(s, "encoder.mid_block.resnets.0", c, c, 0)
    _vae_attn(s, "encoder.mid_block.attentions.0", c)
    _resnet(s, "encoder.mid_block.resnets.1", c, c, 0)
    s["encoder.conv_norm_out.weight"] = (c,)
    s["encoder.conv_norm_out.bias"] = (c,)
    s["encoder.conv_out.weight"] = (2 * L, c, 3, 3)
    s["encoder.conv_out.bias"] = (2 * L,)
    s["quant_conv.weight"] = (2 * L, 2 * L, 1, 1)
    s["quant_conv.bias"] = (2 * L,)
    s["post_quant_conv.weight"] = (L, L, 1, 1)
    s["post_quant_conv.bias"] = (L,)
    s["decoder.conv_in.weight"] = (c, L, 3, 3)
    s["decoder.conv_in.bias"] = (c,)
    _resnet(s, "decoder.mid_block.resnets.0", c, c, 0)
    _vae_attn(s, "decoder.mid_block.attentions.0", c)
    _resnet(s, "decoder.mid_block.resnets.1", c, c, 0)
    rev = list(reversed(ch))
    cin = c
    for i, cout in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            _resnet(s, f"decoder.up_blocks.{i}.resnets.{j}", cin, cout, 0)
            cin = cout
        if i != len(rev) - 1:
            s[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            s[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"] = (cout,)
    s["decoder.conv_norm_out.weight"] = (ch[0],)
    s["decoder.conv_norm_out.bias"] = (ch[0],)
    s["decoder.conv_out.weight"] = (3, ch[0], 3, 3)
    s["decoder.conv_out.bias"] = (3,)
    return s


def encoder_param_shapes(cfg: EncoderConfig) -> Dict[str, tuple]:
    """ConditionPatchEncoder state-dict keys (trt_inference/image_encoder.py:43-76) with the openai-CLIP visual tower
    (`clip.visual.*`, proj removed)."""
    s: Dict[str, tuple] = {}
    w = cfg.width
    v = "clip.visual"
    s[f"{v}.conv1.weight"] = (w, 3, 32, 32)
    s[f"{v}.class_embedding"] = (w,)
    s[f"{v}.positional_embedding"] = (50, w)
    for n in ("ln_pre", "ln_post"):
        s[f"{v}.{n}.weight"] = (w,)
        s[f"{v}.{n}.bias"] = (w,)
    for i in range(cfg.layers):
        b = f"{v}.transformer.resblocks.{i}"
        for n in ("ln_1", "ln_2"):
            s[f"{b}.{n}.weight"] = (w,)
            s[f"{b}.{n}.bias"] = (w,)
        s[f"{b}.attn.in_proj_weight"] = (3 * w, w)
        s[f"{b}.attn.in_proj_bias"] = (3 * w,)
        s[f"{b}.attn.out_proj.weight"] = (w, w)
        s[f"{b}.attn.out_proj.bias"] = (w,)
        s[f"{b}.mlp.c_fc.weight"] = (cfg.mlp, w)
        s[f"{b}.mlp.c_fc.bias"] = (cfg.mlp,)
        s[f"{b}.mlp.c_proj.weight"] = (w, cfg.mlp)
        s[f"{b}.mlp.c_proj.bias"] = (w,)
    for tower in ("l", "m", "s"):
        for i in range(cfg.tower_layers):
            b = f"{tower}_patch_encoder_layers.{i}"
            for n in ("norm1", "norm3"):
                s[f"{b}.{n}.weight"] = (w,)
                s[f"{b}.{n}.bias"] = (w,)
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                s[f"{b}.attn1.{n}.weight"] = (w, w)
                s[f"{b}.attn1.{n}.bias"] = (w,)
            s[f"{b}.ff.net.0.proj.weight"] = (4 * w, w)
            s[f"{b}.ff.net.0.proj.bias"] = (4 * w,)
            s[f"{b}.ff.net.2.weight"] = (w, 4 * w)
            s[f"{b}.ff.net.2.bias"] = (w,)
    s["final_layer_norm.weight"] = (w,)
    s["final_layer_norm.bias"] = (w,)
    s["proj_out.weight"] = (cfg.cross_dim, w)
    s["proj_out.bias"] = (cfg.cross_dim,)
    s["uncond_vector"] = (1, sum(cfg.num_patches), cfg.cross_dim)
    return s


# ----------------------------------------------------------------------------------------------------------------
# synthetic weights
# ----------------------------------------------------------------------------------------------------------------
_RESIDUAL_TAILS = (".conv2.weight", ".to_out.0.weight", ".ff.net.2.weight", ".proj_out.weight", ".proj_attn.weight",
                   ".attn.out_proj.weight", ".mlp.c_proj.weight")


def synth_state_dict(shapes: Dict[str, tuple], seed: int) -> Dict[str, torch.Tensor]:
    """Seeded variance-preserving init: weights ~ N(0, 1/fan_in) (residual-branch tails x0.3 so that 20 denoising steps of
    a random network stay inside fp16 range), norm gains 1 + 0.1 N, biases 0.02 N, LoRA factors 0.02 N / 0.1 N."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in shapes.items():
        if "_lora.down" in name:
            t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(shape[1]))
        elif "_lora.up" in name:
            t = torch.randn(shape, generator=g) * 0.02
        elif name.endswith("norm.weight") or any(name.endswith(f"{n}.weight") for n in (
                "norm1", "norm2", "norm3", "conv_norm_out", "group_norm", "ln_1", "ln_2", "ln_pre", "ln_post",
                "final_layer_norm")):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias") or name.endswith("in_proj_bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif name.endswith("class_embedding") or name.endswith("positional_embedding"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif name == "uncond_vector":
            t = torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
            if name.endswith(_RESIDUAL_TAILS) and not name.startswith("proj_out"):
                t = t * 0.3
        sd[name] = t
    return sd


def synth_model(cfg: ModelConfig, seed: int = 20240726):
    """(unet_sd incl. LoRA factors, vae_sd, encoder_sd) in fp32 with diffusers / openai key names."""
    return (synth_state_dict(unet_param_shapes(cfg.unet), seed),
            synth_state_dict(vae_param_shapes(cfg.vae), seed + 1),
            synth_state_dict(encoder_param_shapes(cfg.enc), seed + 2))


def merge_lora(unet_sd: Dict[str, torch.Tensor], scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """W <- W + scale * up @ down for to_q/to_k/to_v/to_out.0 of every attention module; returns a dict without the
    processor keys (trt_inference/models.py:1046-1093 does this in place before ONNX export)."""
    out = {k: v for k, v in unet_sd.items() if ".processor." not in k}
    for k, v in unet_sd.items():
        if k.endswith("_lora.down.weight"):
            base, nm = k.split(".processor.to_")
            nm = nm.split("_lora")[0]
            up = unet_sd[k.replace(".down.", ".up.")]
            target = f"{base}.to_out.0.weight" if nm == "out" else f"{base}.to_{nm}.weight"
            out[target] = out[target] + scale * (up.float() @ v.float()).to(out[target].dtype)
    return out


def round_fp16(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """fp32 state dict whose matrix weights hold exactly the values the fp16 engines see (biases / norm affine stay fp32)."""
    out = {}
    for k, v in sd.items():
        out[k] = v.half().float() if (v.dim() >= 2 and k != "uncond_vector") else v.float()
    return out


# ----------------------------------------------------------------------------------------------------------------
# packing for the CUDA runtime: fp16 [N, K] matrices (conv: k = (ky*3+kx)*Cin_pad + c), fp32 vectors
# ----------------------------------------------------------------------------------------------------------------
def _pad64(c):
    return (c + 63) // 64 * 64


def pack_conv3x3(w: torch.Tensor) -> torch.Tensor:
    cout, cin = w.shape[:2]
    cp = _pad64(cin)
    t = torch.zeros(cout, 3, 3, cp, dtype=torch.float32)
    t[..., :cin] = w.float().permute(0, 2, 3, 1)
    return t.reshape(cout, 9 * cp).half().contiguous()


def geglu_perm(c4: int) -> torch.Tensor:
    """Row permutation that interleaves the value / gate halves of ff.net.0.proj per 32-row chunk (16 values | 16 gates),
    matching the EPI_GEGLU epilogue of the contraction kernel."""
    idx = torch.arange(c4).view(-1, 16)
    return torch.cat([idx, idx + c4], dim=1).reshape(-1)


def pack_unet(unet_sd_merged: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Merged (LoRA-free) diffusers UNet state dict -> runtime tensors. Matrix weights must already be fp16-representable
    or are rounded here."""
    sd = unet_sd_merged
    out: Dict[str, torch.Tensor] = {}
    done = set()
    for k, v in sd.items():
        if k in done:
            continue
        if k.endswith(".attn1.to_q.weight"):
            b = k[:-len("to_q.weight")]
            out[b + "to_qkv.weight"] = torch.cat([sd[b + "to_q.weight"], sd[b + "to_k.weight"], sd[b + "to_v.weight"]],
                                                 0).half().contiguous()
            done.update({b + "to_q.weight", b + "to_k.weight", b + "to_v.weight"})
        elif k.endswith(".attn2.to_k.weight"):
            b = k[:-len("to_k.weight")]
            out[b + "to_kv.weight"] = torch.cat([sd[b + "to_k.weight"], sd[b + "to_v.weight"]], 0).half().contiguous()
            done.update({b + "to_k.weight", b + "to_v.weight"})
        elif k.endswith(".attn1.to_k.weight") or k.endswith(".attn1.to_v.weight") or k.endswith(".attn2.to_v.weight"):
            continue
        elif k.endswith("ff.net.0.proj.weight"):
            perm = geglu_perm(v.shape[0] // 2)
            out[k] = v[perm].half().contiguous()
            bk = k[:-len("weight")] + "bias"
            out[bk] = sd[bk][perm].float().contiguous()
        elif k.endswith("ff.net.0.proj.bias"):
            continue
        elif v.dim() == 4 and v.shape[-1] == 3:
            out[k] = pack_conv3x3(v)
        elif v.dim() == 4:
            out[k] = v.reshape(v.shape[0], v.shape[1]).half().contiguous()
        elif v.dim() == 2:
            out[k] = v.half().contiguous()
        else:
            out[k] = v.float().contiguous()
    return out


def pack_vae(vae_sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    for k, v in vae_sd.items():
        if k.endswith(".query.weight"):
            b = k[:-len("query.weight")]
            out[b + "qkv.weight"] = torch.cat([vae_sd[b + "query.weight"], vae_sd[b + "key.weight"],
                                               vae_sd[b + "value.weight"]], 0).half().contiguous()
            out[b + "qkv.bias"] = torch.cat([vae_sd[b + "query.bias"], vae_sd[b + "key.bias"],
                                             vae_sd[b + "value.bias"]], 0).float().contiguous()
        elif any(k.endswith(f".{n}.{t}") for n in ("query", "key", "value") for t in ("weight", "bias")):
            continue
        elif k == "post_quant_conv.weight":
            # 1x1 conv on the 4 latent channels; K padded to 8 so that activation rows are 16-byte aligned for TMA
            t = torch.zeros(v.shape[0], 8, dtype=torch.float32)
            t[:, :v.shape[1]] = v.reshape(v.shape[0], v.shape[1]).float()
            out[k] = t.half().contiguous()
        elif v.dim() == 4 and v.shape[-1] == 3:
            out[k] = pack_conv3x3(v)
        elif v.dim() == 4:
            out[k] = v.reshape(v.shape[0], v.shape[1]).half().contiguous()
        elif v.dim() == 2:
            out[k] = v.half().contiguous()
        else:
            out[k] = v.float().contiguous()
    return out


def positional_encoding_2d(channels: int, height: int, width: int) -> torch.Tensor:
    """2-D sinusoidal table of ConditionPatchEncoder (trt_inference/image_encoder.py:20-31): x on the first half of the
    channels, y on the second half, sin on even / cos on odd channel indices."""
    pe = torch.zeros(channels, height, width)
    d = channels // 2
    inv = torch.pow(10000.0, -torch.arange(0.0, d, 2) / d)
    xs = torch.arange(0.0, width)[None, :] * inv[:, None]   # (d/2, W)
    ys = torch.arange(0.0, height)[None, :] * inv[:, None]  # (d/2, H)
    pe[0:d:2] = torch.sin(xs)[:, None, :].expand(-1, height, -1)
    pe[1:d:2] = torch.cos(xs)[:, None, :].expand(-1, height, -1)
    pe[d::2] = torch.sin(ys)[:, :, None].expand(-1, -1, width)
    pe[d + 1::2] = torch.cos(ys)[:, :, None].expand(-1, -1, width)
    return pe


def patch_pos_emb(hid: int, num_patches=(1, 4, 9)) -> torch.Tensor:
    """(1, 14, hid) table: each (C, s, s) grid is REINTERPRETED as (s*s, C) (a raw view, not a permute) exactly as the
    reference does (image_encoder.py:54-56) — the model was trained with that scramble."""
    parts = []
    for n in num_patches:
        s = int(round(n ** 0.5))
        parts.append(positional_encoding_2d(hid, s, s).reshape(1, n, hid))
    return torch.cat(parts, dim=1)


def pack_encoder(enc_sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    for k, v in enc_sd.items():
        if k.endswith("conv1.weight"):
            out[k] = v.reshape(v.shape[0], -1).half().contiguous()  # [width, 3*32*32], k = c*1024 + dy*32 + dx
        elif k.endswith(".attn1.to_q.weight"):
            b = k[:-len("to_q.weight")]
            out[b + "to_qkv.weight"] = torch.cat([enc_sd[b + "to_q.weight"], enc_sd[b + "to_k.weight"],
                                                  enc_sd[b + "to_v.weight"]], 0).half().contiguous()
            out[b + "to_qkv.bias"] = torch.cat([enc_sd[b + "to_q.bias"], enc_sd[b + "to_k.bias"],
                                                enc_sd[b + "to_v.bias"]], 0).float().contiguous()
        elif any(k.endswith(f".attn1.{n}.{t}") for n in ("to_q", "to_k", "to_v") for t in ("weight", "bias")):
            continue
        elif k == "uncond_vector":
            out[k] = v.float().contiguous()
        elif v.dim() == 2 and not k.endswith("positional_embedding"):
            out[k] = v.half().contiguous()
        else:
            out[k] = v.float().contiguous()
    return out
