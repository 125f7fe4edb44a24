"""ORACLE (test infrastructure): diffusers-0.12.0 UNet2DConditionModel (SD-1.5 inpainting config) as functions over a
state dict. Graph per SURVEY.md Appendix A.1; I/O contract per trt_inference/models.py:1097-1139:
sample (3B,9,h,w), timestep scalar, encoder_hidden_states (3B,14,768) -> (3B,4,h,w)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def resnet(sd, p, x, temb, groups, eps):
    """ResnetBlock2D: GN->SiLU->conv1 (+ time_emb_proj(SiLU(temb))) -> GN->SiLU->conv2 ; + shortcut."""
    h = F.silu(F.group_norm(x, groups, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], eps))
    h = F.conv2d(h, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=1)
    if temb is not None:
        t = F.linear(F.silu(temb), sd[f"{p}.time_emb_proj.weight"], sd[f"{p}.time_emb_proj.bias"])
        h = h + t[:, :, None, None]
    h = F.silu(F.group_norm(h, groups, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], eps))
    h = F.conv2d(h, sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1)
    if f"{p}.conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[f"{p}.conv_shortcut.weight"], sd[f"{p}.conv_shortcut.bias"])
    return x + h


def _lora(sd, p, nm, x):
    k = f"{p}.processor.to_{nm}_lora.down.weight"
    if k in sd:  # un-merged LoRAAttnProcessor path (scale 1.0): used to cross-check the merge
        return F.linear(F.linear(x, sd[k]), sd[f"{p}.processor.to_{nm}_lora.up.weight"])
    return 0.0


def attention(sd, p, x, ctx, heads):
    """CrossAttention (diffusers 0.12): to_q/k/v without bias, to_out.0 with bias, softmax in fp32."""
    ctx = x if ctx is None else ctx
    q = F.linear(x, sd[f"{p}.to_q.weight"]) + _lora(sd, p, "q", x)
    k = F.linear(ctx, sd[f"{p}.to_k.weight"]) + _lora(sd, p, "k", ctx)
    v = F.linear(ctx, sd[f"{p}.to_v.weight"]) + _lora(sd, p, "v", ctx)
    B, n, C = q.shape
    d = C // heads
    q = q.view(B, n, heads, d).transpose(1, 2)
    k = k.view(B, -1, heads, d).transpose(1, 2)
    v = v.view(B, -1, heads, d).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (d ** -0.5)
    a = torch.softmax(s.float(), dim=-1).to(q.dtype)
    o = (a @ v).transpose(1, 2).reshape(B, n, C)
    return F.linear(o, sd[f"{p}.to_out.0.weight"], sd[f"{p}.to_out.0.bias"]) + _lora(sd, p, "out", o)


def transformer2d(sd, p, x, ctx, heads, groups):
    """Transformer2DModel with one BasicTransformerBlock (GEGLU feed-forward)."""
    B, C, H, W = x.shape
    res = x
    h = F.group_norm(x, groups, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], 1e-6)
    h = F.conv2d(h, sd[f"{p}.proj_in.weight"], sd[f"{p}.proj_in.bias"])
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    b = f"{p}.transformer_blocks.0"
    h = h + attention(sd, f"{b}.attn1", F.layer_norm(h, (C,), sd[f"{b}.norm1.weight"], sd[f"{b}.norm1.bias"]), None,
                      heads)
    h = h + attention(sd, f"{b}.attn2", F.layer_norm(h, (C,), sd[f"{b}.norm2.weight"], sd[f"{b}.norm2.bias"]), ctx,
                      heads)
    n3 = F.layer_norm(h, (C,), sd[f"{b}.norm3.weight"], sd[f"{b}.norm3.bias"])
    proj = F.linear(n3, sd[f"{b}.ff.net.0.proj.weight"], sd[f"{b}.ff.net.0.proj.bias"])
    a, gate = proj.chunk(2, dim=-1)
    h = h + F.linear(a * F.gelu(gate), sd[f"{b}.ff.net.2.weight"], sd[f"{b}.ff.net.2.bias"])
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    h = F.conv2d(h, sd[f"{p}.proj_out.weight"], sd[f"{p}.proj_out.bias"])
    return h + res


def unet_forward(sd, cfg, sample, timestep, ctx, taps=None):
    """UNet2DConditionModel.forward. `taps` (dict) optionally receives named intermediate activations."""
    ch = cfg.block_out_channels
    g = cfg.groups
    t = torch.as_tensor(timestep, dtype=torch.float32, device=sample.device).reshape(-1)
    t = t.expand(sample.shape[0])
    temb = timestep_embedding(t, ch[0]).to(sample.dtype)
    temb = F.linear(temb, sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"])
    temb = F.linear(F.silu(temb), sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])
    x = F.conv2d(sample, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    if taps is not None:
        taps["conv_in"] = x
    skips = [x]
    for i in range(len(ch)):
        for j in range(cfg.layers_per_block):
            x = resnet(sd, f"down_blocks.{i}.resnets.{j}", x, temb, g, 1e-5)
            if cfg.down_attention[i]:
                x = transformer2d(sd, f"down_blocks.{i}.attentions.{j}", x, ctx, cfg.heads, g)
            skips.append(x)
            if taps is not None:
                taps[f"down.{i}.{j}"] = x
        if i != len(ch) - 1:
            x = F.conv2d(x, sd[f"down_blocks.{i}.downsamplers.0.conv.weight"],
                         sd[f"down_blocks.{i}.downsamplers.0.conv.bias"], stride=2, padding=1)
            skips.append(x)
    x = resnet(sd, "mid_block.resnets.0", x, temb, g, 1e-5)
    x = transformer2d(sd, "mid_block.attentions.0", x, ctx, cfg.heads, g)
    x = resnet(sd, "mid_block.resnets.1", x, temb, g, 1e-5)
    if taps is not None:
        taps["mid"] = x
    up_attn = list(reversed(cfg.down_attention))
    for i in range(len(ch)):
        for j in range(cfg.layers_per_block + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet(sd, f"up_blocks.{i}.resnets.{j}", x, temb, g, 1e-5)
            if up_attn[i]:
                x = transformer2d(sd, f"up_blocks.{i}.attentions.{j}", x, ctx, cfg.heads, g)
        if i != len(ch) - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = F.conv2d(x, sd[f"up_blocks.{i}.upsamplers.0.conv.weight"],
                         sd[f"up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
        if taps is not None:
            taps[f"up.{i}"] = x
    x = F.silu(F.group_norm(x, g, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], 1e-5))
    return F.conv2d(x, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)
