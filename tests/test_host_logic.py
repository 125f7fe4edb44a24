"""CPU: host-side logic of the product (weight inventory, LoRA merge, packing layouts, schedule tables, façade plumbing)
and the C-ABI surface (library loads and exports every symbol include/dtp.h declares; no compute calls without a GPU)."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

from diffusiontexturepainting_b200 import weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parameter_counts_match_sd15():
    cfg = W.sd15_config()
    n_unet = sum(math.prod(s) for k, s in W.unet_param_shapes(cfg.unet).items() if "lora" not in k)
    n_vae = sum(math.prod(s) for s in W.vae_param_shapes(cfg.vae).values())
    n_clip = sum(math.prod(s) for k, s in W.encoder_param_shapes(cfg.enc).items() if k.startswith("clip."))
    assert n_unet == 859_535_364  # SD-1.5 inpainting UNet (9 input channels)
    assert n_vae == 83_653_863
    assert n_clip == 87_456_000  # ViT-B/32 visual tower without the projection
    n_lora = sum(1 for k in W.unet_param_shapes(cfg.unet) if k.endswith("_lora.down.weight"))
    assert n_lora == 32 * 4  # 32 attention modules x (q, k, v, out): train_texture_inpaint_lora.py:411-433


def test_lora_merge_equals_residual():
    from oracle import unet as un
    cfg = W.tiny_config()
    u = W.synth_state_dict(W.unet_param_shapes(cfg.unet), 3)
    m = W.merge_lora(u)
    assert not any(".processor." in k for k in m)
    g = torch.Generator().manual_seed(0)
    x, ctx = torch.randn(3, 9, 8, 8, generator=g), torch.randn(3, 14, cfg.unet.cross_dim, generator=g)
    a = un.unet_forward(u, cfg.unet, x, 501, ctx)
    b = un.unet_forward(m, cfg.unet, x, 501, ctx)
    assert torch.allclose(a, b, atol=2e-5)
    k = "down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k"
    d = m[k + ".weight"] - u[k + ".weight"]
    lo = u[k.replace(".to_k", ".processor.to_k_lora.up.weight")] @ u[k.replace(".to_k", ".processor.to_k_lora.down.weight")]
    assert torch.allclose(d, lo, atol=1e-6) and d.abs().max() > 0


def test_packing_layouts():
    cfg = W.tiny_config()
    u, v, e = W.synth_model(cfg, 5)
    pu = W.pack_unet(W.merge_lora(u))
    # conv3x3: [Cout, 9*Cin_pad], k = (ky*3+kx)*Cin_pad + c, zero padded channels
    w = u["conv_in.weight"]
    p = pu["conv_in.weight"].float().view(w.shape[0], 3, 3, 64)
    assert torch.equal(p[..., :9], w.half().float().permute(0, 2, 3, 1)) and p[..., 9:].abs().max() == 0
    # fused QKV / KV
    b = "down_blocks.1.attentions.0.transformer_blocks.0"
    C = cfg.unet.block_out_channels[1]
    assert pu[f"{b}.attn1.to_qkv.weight"].shape == (3 * C, C)
    assert pu[f"{b}.attn2.to_kv.weight"].shape == (2 * C, cfg.unet.cross_dim)
    # GEGLU interleave: chunk of 32 rows = 16 value rows + their 16 gate rows
    wf = W.merge_lora(u)[f"{b}.ff.net.0.proj.weight"].half()
    pf = pu[f"{b}.ff.net.0.proj.weight"]
    assert torch.equal(pf[0:16], wf[0:16]) and torch.equal(pf[16:32], wf[4 * C:4 * C + 16])
    assert torch.equal(pf[32:48], wf[16:32])
    bf = u[f"{b}.ff.net.0.proj.bias"]
    assert torch.equal(pu[f"{b}.ff.net.0.proj.bias"][16:32], bf[4 * C:4 * C + 16])
    assert all(t.dtype in (torch.float16, torch.float32) for t in pu.values())
    assert not any(k.endswith((".to_k.weight", ".to_v.weight")) for k in pu)
    pv = W.pack_vae(v)
    assert pv["post_quant_conv.weight"].shape == (4, 8) and pv["post_quant_conv.weight"][:, 4:].abs().max() == 0
    assert pv["encoder.mid_block.attentions.0.qkv.weight"].shape[0] == 3 * cfg.vae.block_out_channels[-1]
    pe = W.pack_encoder(e)
    assert pe["clip.visual.conv1.weight"].shape == (cfg.enc.width, 3072)
    assert pe["l_patch_encoder_layers.0.attn1.to_qkv.bias"].shape == (3 * cfg.enc.width,)
    assert pe["uncond_vector"].dtype == torch.float32


def test_round_fp16_only_touches_matrices():
    cfg = W.tiny_config()
    u = W.synth_state_dict(W.unet_param_shapes(cfg.unet), 1)
    r = W.round_fp16(u)
    assert torch.equal(r["conv_in.bias"], u["conv_in.bias"])
    assert torch.equal(r["conv_in.weight"], u["conv_in.weight"].half().float())


def test_facade_settings_are_cast_from_numpy_scalars():
    """Wire settings are numpy scalars (server_io.py:105-119); with numpy >= 2 `1000 // np.uint8(20)` overflows and
    `np.uint8(0) - 1` wraps, so the façade casts on entry."""
    from diffusiontexturepainting_b200.trt_model import TRTConditionalInpainter
    s = TRTConditionalInpainter._settings(dict(steps=np.uint8(20), context_pad=np.uint8(150), tg_steps=np.uint8(0),
                                               width=np.uint16(256), cfg_weight=np.float32(2.0),
                                               tg_weight=np.float32(1.0)))
    assert s == dict(steps=20, context_pad=150, tg_steps=0, cfg_weight=2.0, tg_weight=1.0)
    assert all(type(v) in (int, float) for v in s.values())
    from diffusiontexturepainting_b200.scheduler import DDIMScheduler
    d = DDIMScheduler(device="cpu")
    d.set_timesteps(np.uint8(20))
    d.configure()
    assert int(d.timesteps[0]) == 951 and len(d.evaluation_schedule(1)[0]) == 19


def test_crop_resize_square_and_patches():
    from diffusiontexturepainting_b200.trt_model import crop_resize_square
    from oracle import image_encoder as ie
    img = torch.rand(3, 90, 130, generator=torch.Generator().manual_seed(0))
    a = crop_resize_square(img, 64)
    assert a.shape == (3, 64, 64) and torch.equal(a, ie.crop_resize_square(img, 64))
    assert torch.equal(crop_resize_square(img[:, :64, :64], 64), img[:, :64, :64])


def test_shard_range_partitions():
    from diffusiontexturepainting_b200.parallel import shard_range
    for n in (1, 7, 8, 32, 33):
        for w in (1, 2, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1


def test_library_exports_every_declared_symbol():
    from diffusiontexturepainting_b200 import _native as nat
    from diffusiontexturepainting_b200 import build
    lib_path = build.build()
    L = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, "include", "dtp.h")).read()
    declared = set(re.findall(r"\b(dtp_[a-z0-9_]+)\s*\(", header))
    declared -= {"dtp_engine"}
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/dtp.h but not exported"
    for name in nat.exported_symbols():
        assert name in declared, f"{name} bound in _native.py but not declared in include/dtp.h"


def test_no_gpu_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from diffusiontexturepainting_b200 import _native as nat
    from diffusiontexturepainting_b200.engine import Engine
    with pytest.raises(nat.DtpError):
        Engine(W.tiny_config())


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "diffusiontexturepainting_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn


def test_resize_fold_algebra():
    """The two resize folds of the launch plan, restated in torch on the CPU (the kernels are checked on the GPU in
    tests/test_gpu_ops.py): (a) conv3x3(nearest_upsample_2x(x)) == four parity-class 2x2 convolutions over x with the taps
    that land on the same source pixel pre-summed (kernels.cu upconv_fold_weights_kernel, gemm_setup_upconv2x);
    (b) a stride-2 3x3 convolution reads tap (ky, kx) from the (row parity, column parity) view of the input shifted by
    floor((k - pad) / 2) (gemm_setup_conv3x3_s2), for both paddings the model uses."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 5, 6, 8, generator=g, dtype=torch.float64)
    w = torch.randn(7, 5, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1)
    taps = {(0, 0): [0], (0, 1): [1, 2], (1, 0): [0, 1], (1, 1): [2]}  # S(parity, s): 3x3 taps that share source offset s-1+parity
    out = torch.zeros_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))
    H, W = x.shape[2:]
    for py in range(2):
        for px in range(2):
            acc = 0
            for sy in range(2):
                for sx in range(2):
                    weff = sum(w[:, :, ky, kx] for ky in taps[(py, sy)] for kx in taps[(px, sx)])
                    oy, ox = sy - 1 + py, sx - 1 + px  # source offset of this tap
                    src = xp[:, :, 1 + oy:1 + oy + H, 1 + ox:1 + ox + W]
                    acc = acc + torch.einsum("oc,nchw->nohw", weff, src)
            out[:, :, py::2, px::2] = acc
    assert torch.allclose(out, ref, atol=1e-12)
    for pad in (1, 0):
        if pad:
            ref2 = F.conv2d(x, w, stride=2, padding=1)
        else:
            ref2 = F.conv2d(F.pad(x, (0, 1, 0, 1)), w, stride=2)
        Ho, Wo = H // 2, W // 2
        views = {(py, px): F.pad(x[:, :, py::2, px::2], (1, 1, 1, 1)) for py in range(2) for px in range(2)}  # zero fill = TMA OOB
        acc = 0
        for ky in range(3):
            for kx in range(3):
                oy, ox = ky - pad, kx - pad
                py, px = oy & 1, ox & 1
                sy, sx = (oy - py) >> 1, (ox - px) >> 1
                src = views[(py, px)][:, :, 1 + sy:1 + sy + Ho, 1 + sx:1 + sx + Wo]
                acc = acc + torch.einsum("oc,nchw->nohw", w[:, :, ky, kx], src)
        assert torch.allclose(acc, ref2, atol=1e-12)


def test_header_is_plain_c_and_a_c_host_gets_a_clean_error_without_a_gpu(tmp_path):
    """The boundary is a C ABI: include/dtp.h must compile as C99 (and C++17) on its own, a C host must link against the
    library, and on a machine without an sm_100 device dtp_create must FAIL WITH A MESSAGE (no crash, no silent fallback)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    inc = os.path.join(ROOT, "include")
    for std, lang in (("-std=c99", "c"), ("-std=c++17", "c++")):
        r = subprocess.run([cc, "-x", lang, std, "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-I", inc,
                            os.path.join(inc, "dtp.h")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    from diffusiontexturepainting_b200 import _native as nat
    lib = nat.library_path() if hasattr(nat, "library_path") else os.path.join(ROOT, "diffusiontexturepainting_b200", "libdtp_sm100.so")
    if not os.path.exists(lib):
        pytest.skip("library not built")
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "dtp.h"
int main(void) {
    dtp_config cfg;
    dtp_handle* h = NULL;
    int rc;
    memset(&cfg, 0, sizeof cfg);
    cfg.unet_in_channels = 9; cfg.unet_out_channels = 4;
    cfg.unet_block_out[0] = 64; cfg.unet_block_out[1] = 128; cfg.unet_block_out[2] = 256; cfg.unet_block_out[3] = 256;
    cfg.unet_down_attn[0] = cfg.unet_down_attn[1] = cfg.unet_down_attn[2] = 1;
    cfg.unet_layers_per_block = 2; cfg.unet_heads = 4; cfg.unet_cross_dim = 128; cfg.groups = 32;
    cfg.vae_block_out[0] = 64; cfg.vae_block_out[1] = 64; cfg.vae_block_out[2] = 128; cfg.vae_block_out[3] = 128;
    cfg.vae_layers_per_block = 2; cfg.vae_latent = 4;
    cfg.enc_width = 128; cfg.enc_layers = 2; cfg.enc_heads = 2; cfg.enc_mlp = 256; cfg.enc_tower_layers = 2;
    cfg.enc_tower_heads = 4; cfg.enc_cross_dim = 128; cfg.enc_tokens = 14; cfg.arena_bytes = 1u << 20;
    rc = dtp_create(&cfg, &h);
    printf("rc=%d handle=%s msg=%s\n", rc, h ? "set" : "null", dtp_last_error(h));
    if (h) dtp_destroy(h);
    return 0;
}
''')
    exe = tmp_path / "host"
    r = subprocess.run([cc, "-std=c99", "-I", inc, str(src), "-o", str(exe), lib, "-Wl,-rpath," + os.path.dirname(lib)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout, r.stderr)
    if not torch.cuda.is_available():
        assert "rc=-" in r.stdout and "handle=null" in r.stdout, r.stdout
        assert "msg=" in r.stdout and len(r.stdout.split("msg=")[1].strip()) > 8, r.stdout  # a reason, not an empty string
