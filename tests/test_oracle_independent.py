"""Pins the functional oracle (oracle/unet.py, oracle/vae.py, oracle/pipeline.add_extra_context) to implementations that
were not written for this repository (oracle/independent.py): transformers' latent-diffusion VQ-VAE encoder / decoder
modules carrying the AutoencoderKL weights (strict load of the renamed diffusers keys), a torch.nn module tree with the
diffusers-0.12.0 UNet parameter names (strict load, torch's own multi-head attention), and scipy's maximum filter for
kornia's flat dilation. If a real `diffusers` fixture exists (tests/golden/diffusers_tiny.npz, written by
tests/golden/make_diffusers_golden.py where diffusers is installed) it is checked too; otherwise that is reported."""
import os

import pytest
import torch

from diffusiontexturepainting_b200 import weights as W
from oracle import independent as ind
from oracle import pipeline as op
from oracle import unet as un
from oracle import vae as va

HERE = os.path.dirname(os.path.abspath(__file__))


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def gen(s):
    return torch.Generator().manual_seed(s)


@pytest.fixture(scope="module")
def tiny():
    cfg = W.tiny_config()
    u, v, e = W.synth_model(cfg)
    return cfg, W.merge_lora(u), v


def test_vae_encoder_matches_transformers_ldm_encoder(tiny):
    cfg, _, v = tiny
    x = torch.rand(2, 3, 64, 64, generator=gen(1)) * 2 - 1
    with torch.inference_mode():
        ref = ind.hf_vae_encode_moments(v, cfg.vae, x)
        got = va.encode_moments(v, cfg.vae, x)
    assert got.shape == ref.shape == (2, 8, 8, 8)
    assert rel(got, ref) < 2e-5, rel(got, ref)


def test_vae_decoder_matches_transformers_ldm_decoder(tiny):
    cfg, _, v = tiny
    z = torch.randn(2, 4, 8, 8, generator=gen(2))
    with torch.inference_mode():
        ref = ind.hf_vae_decode(v, cfg.vae, z)
        got = va.decode(v, cfg.vae, z)
    assert got.shape == ref.shape == (2, 3, 64, 64)
    assert rel(got, ref) < 2e-5, rel(got, ref)


def test_vae_full_width_blocks_match_transformers(tiny):
    """SD-1.x widths (512 channels, one 512-wide attention head) on the mid block only: cheap but exercises the real shapes."""
    cfg = W.sd15_config().vae
    shapes = {k: s for k, s in W.vae_param_shapes(cfg).items() if k.startswith("decoder.mid_block")}
    sd = W.synth_state_dict(shapes, 5)
    x = torch.randn(1, 512, 8, 8, generator=gen(3))
    from transformers.models.janus.configuration_janus import JanusVQVAEConfig
    from transformers.models.janus.modeling_janus import JanusVQVAEMidBlock
    mid = JanusVQVAEMidBlock(JanusVQVAEConfig(dropout=0.0), 512).eval()
    out = {}
    ind._rename_resnet("block_1", "decoder.mid_block.resnets.0", sd, out)
    ind._rename_attn("attn_1", "decoder.mid_block.attentions.0", sd, out)
    ind._rename_resnet("block_2", "decoder.mid_block.resnets.1", sd, out)
    mid.load_state_dict(out, strict=True)
    with torch.inference_mode():
        ref = mid(x.clone())
        h = va.resnet(sd, "decoder.mid_block.resnets.0", x, None, 32, 1e-6)
        h = va.attention_block(sd, "decoder.mid_block.attentions.0", h, 32)
        got = va.resnet(sd, "decoder.mid_block.resnets.1", h, None, 32, 1e-6)
    assert rel(got, ref) < 2e-5


def test_unet_matches_torch_nn_twin_with_diffusers_keys(tiny):
    cfg, u, _ = tiny
    twin = ind.unet_twin(u, cfg.unet)  # strict=True: oracle key inventory == module tree
    assert sum(p.numel() for p in twin.parameters()) == sum(t.numel() for t in u.values())
    x = torch.randn(3, 9, 16, 16, generator=gen(4))
    ctx = torch.randn(3, 14, cfg.unet.cross_dim, generator=gen(5))
    with torch.inference_mode():
        for t in (901.0, 1.0):
            ref = twin(x, t, ctx)
            got = un.unet_forward(u, cfg.unet, x, t, ctx)
            assert got.shape == ref.shape == (3, 4, 16, 16)
            assert rel(got, ref) < 2e-5, (t, rel(got, ref))


def test_unet_sd15_key_inventory_loads_strict_into_twin():
    """Full SD-1.5-inpaint widths: parameter names / shapes of the inventory the engine consumes == the module tree built from
    torch.nn layers (859.5 M parameters; meta device, no arithmetic)."""
    cfg = W.sd15_config().unet
    with torch.device("meta"):
        twin = ind.UNet2DConditionTwin(cfg)
    shapes = {k: s for k, s in W.unet_param_shapes(cfg).items() if ".processor." not in k}
    got = {k: tuple(p.shape) for k, p in twin.state_dict().items()}
    assert got == {k: tuple(s) for k, s in shapes.items()}
    n = sum(p.numel() for p in twin.parameters())
    assert abs(n - 859.5e6) < 0.1e6, n


@pytest.mark.parametrize("pad", [1, 2, 7, 21, 40, 150])
def test_flat_dilation_matches_scipy(pad):
    R = 48
    m = (torch.rand(2, 1, R, R, generator=gen(pad)) > 0.93).float()
    img = torch.rand(2, 3, R, R, generator=gen(pad + 1)) * 2 - 1
    brush = torch.rand(1, 3, R, R, generator=gen(pad + 2)) * 2 - 1
    ctx_img, ctx_mask = op.add_extra_context(brush, img * m, m, pad=pad)
    hint = 1 - ind.scipy_flat_dilation(m, pad)
    assert torch.equal(ctx_mask, torch.clamp(m + hint, 0, 1))
    assert torch.equal(ctx_img, img * m + brush * hint)


def test_real_diffusers_fixture_if_present(tiny):
    path = os.path.join(HERE, "golden", "diffusers_tiny.npz")
    if not os.path.exists(path):
        pytest.skip("UNPINNED vs the real diffusers 0.12 classes: diffusers is not installable in this image; run "
                    "tests/golden/make_diffusers_golden.py where it is to create tests/golden/diffusers_tiny.npz")
    import numpy as np
    cfg, u, v = tiny
    d = np.load(path)
    t = lambda k: torch.from_numpy(d[k])
    with torch.inference_mode():
        assert rel(un.unet_forward(u, cfg.unet, t("unet_x"), float(d["unet_t"]), t("unet_ctx")), t("unet_out")) < 1e-4
        assert rel(va.encode_moments(v, cfg.vae, t("vae_x")), t("vae_moments")) < 1e-4
        assert rel(va.decode(v, cfg.vae, t("vae_z")), t("vae_dec")) < 1e-4
