"""ORACLE CROSS-CHECKS (test infrastructure): the same graphs as oracle/unet.py and oracle/vae.py, but built from code that
was NOT written for this repository, so that the functional oracle is pinned to something external:

  * AutoencoderKL encoder  == transformers' ChameleonVQVAEEncoder (the taming / latent-diffusion `Encoder` the diffusers
    class derives from: conv_in, [ResnetBlock x2, pad(0,1,0,1)+stride-2 conv] x3, [ResnetBlock x2], mid block with one
    single-head attention, GroupNorm(eps 1e-6) + swish + conv_out). The diffusers-0.12.0 state-dict keys are renamed onto it
    and loaded with strict=True (trt_inference/models.py:1328-1335 wraps exactly this encoder).
  * AutoencoderKL decoder  == transformers' JanusVQVAEDecoder with its extra per-level attention lists emptied (the
    latent-diffusion `Decoder`: conv_in, mid, [ResnetBlock x3, nearest x2 + conv] levels, norm_out, conv_out;
    models.py:1237-1244).
  * UNet2DConditionModel   == a torch.nn module tree (nn.GroupNorm / nn.Conv2d / nn.LayerNorm / nn.Linear and
    F.multi_head_attention_forward for both attention layers) whose parameter names ARE the diffusers-0.12.0 key inventory:
    the merged state dict must load with strict=True (models.py:1036-1095 loads the same keys into the real class).
  * kornia.morphology.dilation with a flat pad x pad kernel (handler.py:25-33) == scipy.ndimage.maximum_filter with a
    constant -1e4 border.

When `diffusers` itself is importable (it is not in the build image), tests/golden/make_diffusers_golden.py writes
tests/golden/diffusers_tiny.npz from the real classes and tests/test_oracle_independent.py consumes it."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------------
# VAE: diffusers key names -> transformers' latent-diffusion encoder / decoder modules
# ----------------------------------------------------------------------------------------------------------------
def _rename_resnet(dst, src, sd, out):
    for n in ("norm1", "conv1", "norm2", "conv2"):
        for t in ("weight", "bias"):
            out[f"{dst}.{n}.{t}"] = sd[f"{src}.{n}.{t}"]
    if f"{src}.conv_shortcut.weight" in sd:
        out[f"{dst}.nin_shortcut.weight"] = sd[f"{src}.conv_shortcut.weight"]
        out[f"{dst}.nin_shortcut.bias"] = sd[f"{src}.conv_shortcut.bias"]


def _rename_attn(dst, src, sd, out):
    out[f"{dst}.norm.weight"] = sd[f"{src}.group_norm.weight"]
    out[f"{dst}.norm.bias"] = sd[f"{src}.group_norm.bias"]
    for a, b in (("q", "query"), ("k", "key"), ("v", "value"), ("proj_out", "proj_attn")):
        w = sd[f"{src}.{b}.weight"]
        out[f"{dst}.{a}.weight"] = w.reshape(w.shape[0], w.shape[1], 1, 1)  # Linear -> 1x1 conv
        out[f"{dst}.{a}.bias"] = sd[f"{src}.{b}.bias"]


def hf_vae_encoder(vae_sd, cfg):
    """transformers ChameleonVQVAEEncoder carrying the diffusers AutoencoderKL encoder weights (strict load)."""
    from transformers.models.chameleon.configuration_chameleon import ChameleonVQVAEConfig
    from transformers.models.chameleon.modeling_chameleon import ChameleonVQVAEEncoder
    ch = cfg.block_out_channels
    hc = ChameleonVQVAEConfig(double_latent=True, latent_channels=cfg.latent_channels, in_channels=3,
                              base_channels=ch[0], channel_multiplier=[c // ch[0] for c in ch],
                              num_res_blocks=cfg.layers_per_block, attn_resolutions=None, attn_type="vanilla",
                              dropout=0.0, resolution=512)
    enc = ChameleonVQVAEEncoder(hc).eval()
    out = {}
    for t in ("weight", "bias"):
        out[f"conv_in.{t}"] = vae_sd[f"encoder.conv_in.{t}"]
        out[f"norm_out.{t}"] = vae_sd[f"encoder.conv_norm_out.{t}"]
        out[f"conv_out.{t}"] = vae_sd[f"encoder.conv_out.{t}"]
    for i in range(len(ch)):
        for j in range(cfg.layers_per_block):
            _rename_resnet(f"down.{i}.block.{j}", f"encoder.down_blocks.{i}.resnets.{j}", vae_sd, out)
        if i != len(ch) - 1:
            for t in ("weight", "bias"):
                out[f"down.{i}.downsample.conv.{t}"] = vae_sd[f"encoder.down_blocks.{i}.downsamplers.0.conv.{t}"]
    _rename_resnet("mid.block_1", "encoder.mid_block.resnets.0", vae_sd, out)
    _rename_attn("mid.attn_1", "encoder.mid_block.attentions.0", vae_sd, out)
    _rename_resnet("mid.block_2", "encoder.mid_block.resnets.1", vae_sd, out)
    enc.load_state_dict(out, strict=True)
    return enc


def hf_vae_decoder(vae_sd, cfg):
    """transformers JanusVQVAEDecoder (per-level attention lists emptied) carrying the AutoencoderKL decoder weights."""
    from transformers.models.janus.configuration_janus import JanusVQVAEConfig
    from transformers.models.janus.modeling_janus import JanusVQVAEDecoder
    ch = cfg.block_out_channels
    hc = JanusVQVAEConfig(latent_channels=cfg.latent_channels, base_channels=ch[0],
                          channel_multiplier=[c // ch[0] for c in ch], num_res_blocks=cfg.layers_per_block,
                          dropout=0.0, out_channels=3, in_channels=3)
    dec = JanusVQVAEDecoder(hc).eval()
    for lvl in dec.up:
        lvl.attn = nn.ModuleList()  # AutoencoderKL has attention in the mid block only
    out = {}
    for t in ("weight", "bias"):
        out[f"conv_in.{t}"] = vae_sd[f"decoder.conv_in.{t}"]
        out[f"norm_out.{t}"] = vae_sd[f"decoder.conv_norm_out.{t}"]
        out[f"conv_out.{t}"] = vae_sd[f"decoder.conv_out.{t}"]
    _rename_resnet("mid.block_1", "decoder.mid_block.resnets.0", vae_sd, out)
    _rename_attn("mid.attn_1", "decoder.mid_block.attentions.0", vae_sd, out)
    _rename_resnet("mid.block_2", "decoder.mid_block.resnets.1", vae_sd, out)
    for i in range(len(ch)):  # Janus stores the levels lowest resolution first, like diffusers' up_blocks
        for j in range(cfg.layers_per_block + 1):
            _rename_resnet(f"up.{i}.block.{j}", f"decoder.up_blocks.{i}.resnets.{j}", vae_sd, out)
        if i != len(ch) - 1:
            for t in ("weight", "bias"):
                out[f"up.{i}.upsample.conv.{t}"] = vae_sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.{t}"]
    dec.load_state_dict(out, strict=True)
    return dec


def hf_vae_encode_moments(vae_sd, cfg, x):
    h = hf_vae_encoder(vae_sd, cfg)(x.clone())
    return F.conv2d(h, vae_sd["quant_conv.weight"], vae_sd["quant_conv.bias"])


def hf_vae_decode(vae_sd, cfg, z):
    h = F.conv2d(z, vae_sd["post_quant_conv.weight"], vae_sd["post_quant_conv.bias"])
    return hf_vae_decoder(vae_sd, cfg)(h)


# ----------------------------------------------------------------------------------------------------------------
# UNet2DConditionModel as a torch.nn module tree with the diffusers-0.12.0 parameter names
# ----------------------------------------------------------------------------------------------------------------
class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        if cin != cout:
            self.conv_shortcut = nn.Conv2d(cin, cout, 1)

    def forward(self, x, temb):
        h = self.conv1(self.nonlinearity(self.norm1(x)))
        h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.nonlinearity(self.norm2(h)))
        if hasattr(self, "conv_shortcut"):
            x = self.conv_shortcut(x)
        return x + h


class CrossAttention(nn.Module):
    """Projections are nn.Linear (diffusers names); the attention itself is torch's own multi-head implementation."""

    def __init__(self, c, kv, heads):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(c, c, bias=False)
        self.to_k = nn.Linear(kv, c, bias=False)
        self.to_v = nn.Linear(kv, c, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        C = x.shape[-1]
        out, _ = F.multi_head_attention_forward(
            x.transpose(0, 1), ctx.transpose(0, 1), ctx.transpose(0, 1), C, self.heads,
            in_proj_weight=None, in_proj_bias=None, bias_k=None, bias_v=None, add_zero_attn=False, dropout_p=0.0,
            out_proj_weight=self.to_out[0].weight, out_proj_bias=self.to_out[0].bias, training=False,
            need_weights=False, use_separate_proj_weight=True, q_proj_weight=self.to_q.weight,
            k_proj_weight=self.to_k.weight, v_proj_weight=self.to_v.weight)
        return out.transpose(0, 1)


class GEGLU(nn.Module):
    def __init__(self, c, inner):
        super().__init__()
        self.proj = nn.Linear(c, 2 * inner)

    def forward(self, x):
        a, gate = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(c, 4 * c), nn.Dropout(0.0), nn.Linear(4 * c, c)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, c, heads, cross):
        super().__init__()
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(c), nn.LayerNorm(c), nn.LayerNorm(c)
        self.attn1 = CrossAttention(c, c, heads)
        self.attn2 = CrossAttention(c, cross, heads)
        self.ff = FeedForward(c)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        return x + self.ff(self.norm3(x))


class Transformer2DModel(nn.Module):
    def __init__(self, c, heads, cross, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.proj_in = nn.Conv2d(c, c, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(c, heads, cross)])
        self.proj_out = nn.Conv2d(c, c, 1)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        h = self.proj_in(self.norm(x)).permute(0, 2, 3, 1).reshape(B, H * W, C)
        h = self.transformer_blocks[0](h, ctx)
        return self.proj_out(h.reshape(B, H, W, C).permute(0, 3, 1, 2)) + x


class _Sampler(nn.Module):
    def __init__(self, c, stride):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=stride, padding=1)


class _Block(nn.Module):
    def __init__(self, resnets, attentions, down=None, up=None):
        super().__init__()
        self.resnets = nn.ModuleList(resnets)
        if attentions:
            self.attentions = nn.ModuleList(attentions)
        if down is not None:
            self.downsamplers = nn.ModuleList([down])
        if up is not None:
            self.upsamplers = nn.ModuleList([up])


class _TimeEmbedding(nn.Module):
    def __init__(self, c, t):
        super().__init__()
        self.linear_1 = nn.Linear(c, t)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(t, t)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


class UNet2DConditionTwin(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        ch, g, T = cfg.block_out_channels, cfg.groups, cfg.time_dim
        self.cfg = cfg
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = _TimeEmbedding(ch[0], T)
        tf = lambda c: Transformer2DModel(c, cfg.heads, cfg.cross_dim, g)
        skip, cin, downs = [ch[0]], ch[0], []
        for i, cout in enumerate(ch):
            res, att = [], []
            for _ in range(cfg.layers_per_block):
                res.append(ResnetBlock2D(cin, cout, T, g, 1e-5))
                if cfg.down_attention[i]:
                    att.append(tf(cout))
                cin = cout
                skip.append(cout)
            last = i == len(ch) - 1
            downs.append(_Block(res, att, down=None if last else _Sampler(cout, 2)))
            if not last:
                skip.append(cout)
        self.down_blocks = nn.ModuleList(downs)
        mid = ch[-1]
        self.mid_block = _Block([ResnetBlock2D(mid, mid, T, g, 1e-5), ResnetBlock2D(mid, mid, T, g, 1e-5)], [tf(mid)])
        ups, cin = [], mid
        up_attn = list(reversed(cfg.down_attention))
        for i, cout in enumerate(reversed(ch)):
            res, att = [], []
            for _ in range(cfg.layers_per_block + 1):
                res.append(ResnetBlock2D(cin + skip.pop(), cout, T, g, 1e-5))
                if up_attn[i]:
                    att.append(tf(cout))
                cin = cout
            ups.append(_Block(res, att, up=None if i == len(ch) - 1 else _Sampler(cout, 1)))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=1e-5)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    @staticmethod
    def sinusoid(t, dim):
        # Timesteps(num_channels, flip_sin_to_cos=True, downscale_freq_shift=0)
        half = dim // 2
        f = torch.exp(torch.arange(half, dtype=torch.float32, device=t.device) * (-math.log(10000.0) / half))
        a = t[:, None].float() * f[None]
        return torch.cat([a.cos(), a.sin()], dim=-1)

    def forward(self, sample, timestep, ctx):
        t = torch.as_tensor(timestep, dtype=torch.float32, device=sample.device).reshape(-1).expand(sample.shape[0])
        temb = self.time_embedding(self.sinusoid(t, self.cfg.block_out_channels[0]).to(sample.dtype))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            for j, r in enumerate(blk.resnets):
                x = r(x, temb)
                if hasattr(blk, "attentions"):
                    x = blk.attentions[j](x, ctx)
                skips.append(x)
            if hasattr(blk, "downsamplers"):
                x = blk.downsamplers[0].conv(x)
                skips.append(x)
        x = self.mid_block.resnets[0](x, temb)
        x = self.mid_block.attentions[0](x, ctx)
        x = self.mid_block.resnets[1](x, temb)
        for blk in self.up_blocks:
            for j, r in enumerate(blk.resnets):
                x = r(torch.cat([x, skips.pop()], dim=1), temb)
                if hasattr(blk, "attentions"):
                    x = blk.attentions[j](x, ctx)
            if hasattr(blk, "upsamplers"):
                x = blk.upsamplers[0].conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return self.conv_out(F.silu(self.conv_norm_out(x)))


def unet_twin(unet_sd_merged, cfg):
    m = UNet2DConditionTwin(cfg).eval()
    m.load_state_dict(unet_sd_merged, strict=True)  # the diffusers key inventory, nothing missing, nothing extra
    return m


# ----------------------------------------------------------------------------------------------------------------
# kornia.morphology.dilation(mask, ones(pad, pad)) through scipy
# ----------------------------------------------------------------------------------------------------------------
def scipy_flat_dilation(mask, pad):
    """(B,1,H,W) -> (B,1,H,W): kornia pads with -max_val = -1e4 ('geodesic' border) and takes the window maximum with the
    structuring element's origin at pad // 2; scipy's centred maximum_filter has the same window for odd and even sizes."""
    import numpy as np
    from scipy import ndimage
    m = mask.detach().cpu().numpy()
    out = np.empty_like(m)
    for b in range(m.shape[0]):
        out[b, 0] = ndimage.maximum_filter(m[b, 0], size=(pad, pad), mode="constant", cval=-1e4)
    return torch.from_numpy(out)
