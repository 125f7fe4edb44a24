"""ORACLE (test infrastructure): ConditionPatchEncoder (trt_inference/image_encoder.py:43-115) with the openai-CLIP
ViT-B/32 visual tower (proj=None, image_encoder.py:49-50) restated over a state dict, plus the brush pre-processing
(image_encoder.py:100-115, handler.py:36-45)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)  # image_encoder.py:75
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)  # image_encoder.py:76


def positional_encoding_2d(channels, height, width):
    """image_encoder.py:20-31 (sin/cos of x on the first half of the channels, of y on the second half)."""
    pe = torch.zeros(channels, height, width)
    d = channels // 2
    freq = 1.0 / (10000.0 ** (torch.arange(0.0, d, 2) / d))
    x = torch.arange(0.0, width)[:, None]
    y = torch.arange(0.0, height)[:, None]
    pe[0:d:2] = torch.sin(x * freq).t()[:, None, :]
    pe[1:d:2] = torch.cos(x * freq).t()[:, None, :]
    pe[d::2] = torch.sin(y * freq).t()[:, :, None]
    pe[d + 1::2] = torch.cos(y * freq).t()[:, :, None]
    return pe


def patch_pos_emb(hid, num_patches=(1, 4, 9)):
    """image_encoder.py:54-56: (C,H,W).view(1, HW, C) — a raw reinterpretation (not a permute); kept bit-for-bit."""
    parts = [positional_encoding_2d(hid, int(math.sqrt(n)), int(math.sqrt(n))).reshape(1, n, hid) for n in num_patches]
    return torch.cat(parts, dim=1)


def crop_resize_square(image, width):
    """handler.py:36-45: CenterCrop(min side) + Resize(width) (bilinear, no antialias: torchvision 0.15 tensor default)."""
    H, W = image.shape[-2:]
    m = min(H, W)
    top, left = int(round((H - m) / 2.0)), int(round((W - m) / 2.0))
    img = image[..., top:top + m, left:left + m]
    if m != width:
        img = F.interpolate(img[None] if img.dim() == 3 else img, size=(width, width), mode="bilinear",
                            align_corners=False)
        img = img[0] if image.dim() == 3 else img
    return img


def preprocess_patches(image):
    """image_encoder.py:100-115: (1,3,R,R) in [0,1] -> (14,3,224,224) normalised multi-scale patches."""
    if image.shape[-1] != 224 or image.shape[-2] != 224:
        image = F.interpolate(image, (224, 224), mode="bicubic", align_corners=True, antialias=False)
    mean = torch.tensor(CLIP_MEAN, device=image.device)[None, :, None, None]
    std = torch.tensor(CLIP_STD, device=image.device)[None, :, None, None]
    image = (image - mean) / std
    img = image[0]
    out = []
    for n in (1, 4, 9):
        ps = 224 // int(math.sqrt(n))
        p = img.unfold(1, ps, ps).unfold(2, ps, ps).permute(1, 2, 0, 3, 4).reshape(-1, 3, ps, ps)
        if ps != 224:
            p = F.interpolate(p, size=(224, 224), mode="bilinear", align_corners=False)
        out.append(p)
    return torch.cat(out, dim=0)


def _mha(x, w_in, b_in, w_out, b_out, heads):
    B, n, C = x.shape
    qkv = F.linear(x, w_in, b_in)
    q, k, v = qkv.chunk(3, dim=-1)
    d = C // heads
    q = q.view(B, n, heads, d).transpose(1, 2)
    k = k.view(B, n, heads, d).transpose(1, 2)
    v = v.view(B, n, heads, d).transpose(1, 2)
    a = torch.softmax((q @ k.transpose(-1, -2)) * (d ** -0.5), dim=-1)
    return F.linear((a @ v).transpose(1, 2).reshape(B, n, C), w_out, b_out)


def clip_visual(sd, cfg, images, prefix="clip.visual"):
    """openai CLIP VisionTransformer.forward with proj=None: (N,3,224,224) -> (N,width) = ln_post(CLS)."""
    v = prefix
    w = cfg.width
    x = F.conv2d(images, sd[f"{v}.conv1.weight"], stride=32)
    x = x.reshape(x.shape[0], w, -1).permute(0, 2, 1)
    cls = sd[f"{v}.class_embedding"].to(x.dtype)[None, None].expand(x.shape[0], 1, w)
    x = torch.cat([cls, x], dim=1) + sd[f"{v}.positional_embedding"].to(x.dtype)
    x = F.layer_norm(x, (w,), sd[f"{v}.ln_pre.weight"], sd[f"{v}.ln_pre.bias"])
    for i in range(cfg.layers):
        b = f"{v}.transformer.resblocks.{i}"
        h = F.layer_norm(x, (w,), sd[f"{b}.ln_1.weight"], sd[f"{b}.ln_1.bias"])
        x = x + _mha(h, sd[f"{b}.attn.in_proj_weight"], sd[f"{b}.attn.in_proj_bias"], sd[f"{b}.attn.out_proj.weight"],
                     sd[f"{b}.attn.out_proj.bias"], cfg.heads)
        h = F.layer_norm(x, (w,), sd[f"{b}.ln_2.weight"], sd[f"{b}.ln_2.bias"])
        h = F.linear(h, sd[f"{b}.mlp.c_fc.weight"], sd[f"{b}.mlp.c_fc.bias"])
        h = h * torch.sigmoid(1.702 * h)  # QuickGELU
        x = x + F.linear(h, sd[f"{b}.mlp.c_proj.weight"], sd[f"{b}.mlp.c_proj.bias"])
    return F.layer_norm(x[:, 0, :], (w,), sd[f"{v}.ln_post.weight"], sd[f"{v}.ln_post.bias"])


def _tower_block(sd, b, x, heads):
    """BasicTransformerBlock(dim, heads, activation_fn='gelu', attention_bias=True), no cross-attention."""
    C = x.shape[-1]
    h = F.layer_norm(x, (C,), sd[f"{b}.norm1.weight"], sd[f"{b}.norm1.bias"])
    w_in = torch.cat([sd[f"{b}.attn1.to_q.weight"], sd[f"{b}.attn1.to_k.weight"], sd[f"{b}.attn1.to_v.weight"]], 0)
    b_in = torch.cat([sd[f"{b}.attn1.to_q.bias"], sd[f"{b}.attn1.to_k.bias"], sd[f"{b}.attn1.to_v.bias"]], 0)
    x = x + _mha(h, w_in, b_in, sd[f"{b}.attn1.to_out.0.weight"], sd[f"{b}.attn1.to_out.0.bias"], heads)
    h = F.layer_norm(x, (C,), sd[f"{b}.norm3.weight"], sd[f"{b}.norm3.bias"])
    h = F.gelu(F.linear(h, sd[f"{b}.ff.net.0.proj.weight"], sd[f"{b}.ff.net.0.proj.bias"]))
    return x + F.linear(h, sd[f"{b}.ff.net.2.weight"], sd[f"{b}.ff.net.2.bias"])


def encoder_forward(sd, cfg, patches):
    """ConditionPatchEncoder.forward (image_encoder.py:78-98): (14,3,224,224) -> ((1,14,cross), uncond (1,14,cross))."""
    w = cfg.width
    tok = clip_visual(sd, cfg, patches).float().view(1, -1, w) + patch_pos_emb(w, cfg.num_patches).to(patches.device)
    l, m, s = cfg.num_patches
    parts = [tok[:, :l], tok[:, l:l + m], tok[:, l + m:]]
    outs = []
    for name, h in zip(("l", "m", "s"), parts):
        for i in range(cfg.tower_layers):
            h = _tower_block(sd, f"{name}_patch_encoder_layers.{i}", h, cfg.tower_heads)
        outs.append(h)
    h = torch.cat(outs, dim=1)
    h = F.layer_norm(h, (w,), sd["final_layer_norm.weight"], sd["final_layer_norm.bias"])
    h = F.linear(h, sd["proj_out.weight"], sd["proj_out.bias"])
    return h, sd["uncond_vector"]


def encode_image(sd, cfg, image):
    """ConditionPatchEncoder.encode_image (image_encoder.py:106-115)."""
    return encoder_forward(sd, cfg, preprocess_patches(image))
