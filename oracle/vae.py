"""ORACLE (test infrastructure): diffusers-0.12.0 AutoencoderKL encoder / decoder with the
reference's wrappers: TorchVAEEncoder = encode(x).latent_dist.sample() (trt_inference/models.py:1328-1335) and
vae.forward = decode (models.py:1237-1244)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .unet import resnet


def attention_block(sd, p, x, groups):
    """AttentionBlock (single head, GN eps 1e-6, fp32 softmax, residual)."""
    B, C, H, W = x.shape
    h = F.group_norm(x, groups, sd[f"{p}.group_norm.weight"], sd[f"{p}.group_norm.bias"], 1e-6)
    h = h.view(B, C, H * W).transpose(1, 2)
    q = F.linear(h, sd[f"{p}.query.weight"], sd[f"{p}.query.bias"])
    k = F.linear(h, sd[f"{p}.key.weight"], sd[f"{p}.key.bias"])
    v = F.linear(h, sd[f"{p}.value.weight"], sd[f"{p}.value.bias"])
    s = (q @ k.transpose(-1, -2)) * (C ** -0.5)
    a = torch.softmax(s.float(), dim=-1).to(q.dtype)
    o = F.linear(a @ v, sd[f"{p}.proj_attn.weight"], sd[f"{p}.proj_attn.bias"])
    return o.transpose(1, 2).reshape(B, C, H, W) + x


def encode_moments(sd, cfg, x):
    """Encoder + quant_conv -> (B, 2*latent, h, w) moments [mean | logvar]."""
    ch = cfg.block_out_channels
    g = cfg.groups
    h = F.conv2d(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    for i in range(len(ch)):
        for j in range(cfg.layers_per_block):
            h = resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, None, g, 1e-6)
        if i != len(ch) - 1:
            h = F.pad(h, (0, 1, 0, 1))
            h = F.conv2d(h, sd[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"],
                         sd[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"], stride=2)
    h = resnet(sd, "encoder.mid_block.resnets.0", h, None, g, 1e-6)
    h = attention_block(sd, "encoder.mid_block.attentions.0", h, g)
    h = resnet(sd, "encoder.mid_block.resnets.1", h, None, g, 1e-6)
    h = F.silu(F.group_norm(h, g, sd["encoder.conv_norm_out.weight"], sd["encoder.conv_norm_out.bias"], 1e-6))
    h = F.conv2d(h, sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


def encode_sample(sd, cfg, x, noise=None):
    """DiagonalGaussianDistribution.sample(): mean + exp(0.5*clamp(logvar,-30,20)) * noise. noise=None -> mode (the
    deterministic parity setting; the reference draws unseeded in-engine noise, models.py:1334-1335)."""
    mom = encode_moments(sd, cfg, x)
    mean, logvar = mom.chunk(2, dim=1)
    if noise is None:
        return mean
    std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
    return mean + std * noise


def decode(sd, cfg, z):
    ch = cfg.block_out_channels
    g = cfg.groups
    h = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    h = F.conv2d(h, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    h = resnet(sd, "decoder.mid_block.resnets.0", h, None, g, 1e-6)
    h = attention_block(sd, "decoder.mid_block.attentions.0", h, g)
    h = resnet(sd, "decoder.mid_block.resnets.1", h, None, g, 1e-6)
    for i in range(len(ch)):
        for j in range(cfg.layers_per_block + 1):
            h = resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, None, g, 1e-6)
        if i != len(ch) - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"],
                         sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
    h = F.silu(F.group_norm(h, g, sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"], 1e-6))
    return F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)
