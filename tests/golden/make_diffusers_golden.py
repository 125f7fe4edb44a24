"""Writes tests/golden/diffusers_tiny.npz from the REAL diffusers classes the reference loads
(trt_inference/models.py:1036-1095 UNet2DConditionModel, :1237-1244 / :1328-1335 AutoencoderKL), instantiated from config
at the tiny widths with this repository's seeded synthetic state dicts (strict load). Run it where `diffusers` is
installed (the build image has no network and no diffusers; tests/test_oracle_independent.py reports the missing fixture):

    pip install diffusers==0.12.0 && python tests/golden/make_diffusers_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    from diffusers import AutoencoderKL, UNet2DConditionModel

    from diffusiontexturepainting_b200 import weights as W
    cfg = W.tiny_config()
    u, v, _ = W.synth_model(cfg)
    u = W.merge_lora(u)
    uc, vc = cfg.unet, cfg.vae
    unet = UNet2DConditionModel(
        in_channels=uc.in_channels, out_channels=uc.out_channels, block_out_channels=uc.block_out_channels,
        layers_per_block=uc.layers_per_block, attention_head_dim=uc.heads, cross_attention_dim=uc.cross_dim,
        norm_num_groups=uc.groups,
        down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
        up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D")).eval()
    unet.load_state_dict(u, strict=True)
    vae = AutoencoderKL(in_channels=3, out_channels=3, latent_channels=vc.latent_channels,
                        block_out_channels=vc.block_out_channels, layers_per_block=vc.layers_per_block,
                        norm_num_groups=vc.groups, down_block_types=("DownEncoderBlock2D",) * 4,
                        up_block_types=("UpDecoderBlock2D",) * 4).eval()
    vae.load_state_dict(v, strict=True)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 9, 16, 16, generator=g)
    ctx = torch.randn(3, 14, uc.cross_dim, generator=g)
    vx = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    vz = torch.randn(2, 4, 8, 8, generator=g)
    with torch.inference_mode():
        out = unet(x, 501.0, encoder_hidden_states=ctx).sample
        moments = vae.quant_conv(vae.encoder(vx))
        dec = vae.decode(vz).sample
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "diffusers_tiny.npz"),
                        unet_x=x.numpy(), unet_t=np.float32(501.0), unet_ctx=ctx.numpy(), unet_out=out.numpy(),
                        vae_x=vx.numpy(), vae_moments=moments.numpy(), vae_z=vz.numpy(), vae_dec=dec.numpy())
    print("wrote diffusers_tiny.npz")


if __name__ == "__main__":
    main()
