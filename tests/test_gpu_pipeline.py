"""Stage-level and end-to-end parity of the native stamp path (through the C-ABI) against the oracle — the fp32
PyTorch restatement of the reference pipeline — on identical seeded inputs and fp16-representable weights.

Tolerances: UNet single forward rel-L2 <= 1e-2; VAE encode / decode <= 1e-2; image encoder <= 1e-2;
end-to-end stamp (SURVEY.md §8c as written): PSNR >= 40 dB on generate_raw (BEFORE the alpha composite) and >= 99.9 % of the
pixels within 2/255 against the fp32 oracle; uint8 wire output <= 2 LSB on >= 99.9 % of the bytes. The benchmarked
configurations are covered at full model size: C2 (512 x 512, steps = 20 -> 19 evaluations, and the strict 20) and C3 per
GPU (256 x 256, B = 4, 10 evaluations)."""
import json
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

DEV = "cuda"
_LOG = {}


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def log(name, **kv):
    _LOG[name] = kv
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_metrics.json"), "w") as f:
            json.dump(_LOG, f, indent=1)
    except OSError:
        pass
    print(name, kv)


class Bundle:
    def __init__(self, cfg):
        from diffusiontexturepainting_b200 import weights as W
        from diffusiontexturepainting_b200.engine import Engine
        from oracle.pipeline import OraclePipeline
        self.W = W
        self.cfg = cfg
        u, v, e = W.synth_model(cfg)
        self.sds = (u, v, e)
        um = W.round_fp16(W.merge_lora(u))
        to = lambda sd: {k: t.to(DEV) for k, t in sd.items()}
        self.oracle_sds = (to(um), to(W.round_fp16(v)), to(W.round_fp16(e)))
        self.engine = Engine(cfg, 0, arena_bytes=6 << 30)
        self.engine.load_state_dicts(u, v, e)
        self.OraclePipeline = OraclePipeline

    def oracle(self, res):
        return self.OraclePipeline(self.cfg, *self.oracle_sds, res)


@pytest.fixture(scope="module")
def tiny():
    from diffusiontexturepainting_b200 import weights as W
    return Bundle(W.tiny_config())


@pytest.fixture(scope="module")
def full():
    from diffusiontexturepainting_b200 import weights as W
    return Bundle(W.sd15_config())


def gen(seed):
    return torch.Generator().manual_seed(seed)


def smooth_image(seed, c, r):
    x = torch.rand(1, c, r // 8 + 1, r // 8 + 1, generator=gen(seed))
    return torch.nn.functional.interpolate(x, size=(r, r), mode="bilinear", align_corners=True)[0].clamp(0, 1)


def make_canvas(B, R, seed=2):
    canvas = torch.stack([torch.cat([smooth_image(seed + i, 3, R), torch.zeros(1, R, R)]) for i in range(B)])
    canvas[:, 3, :int(0.4 * R)] = 1.0
    return canvas


def check_vae(b, R, B=2, name="tiny"):
    from oracle import vae as va
    x = (torch.stack([smooth_image(10 + i, 3, R) for i in range(B)]) * 2 - 1).to(DEV)
    noise = torch.randn(B, 4, R // 8, R // 8, generator=gen(5)).to(DEV)
    ora = b.oracle(R)
    for nz, tag in ((None, "mode"), (noise, "sampled")):
        got = b.engine.vae_encode(x, nz)
        ref = ora.vae_encode(x, nz)
        e = rel_l2(got, ref)
        log(f"{name}.vae_encode.{tag}.R{R}", rel_l2=e)
        assert e < 1e-2
    z = torch.randn(B, 4, R // 8, R // 8, generator=gen(6)).to(DEV) * 0.18215 * 4
    got = b.engine.vae_decode(z)
    ref = (va.decode(b.oracle_sds[1], b.cfg.vae, z / 0.18215) / 2 + 0.5).clamp(0, 1)
    e = rel_l2(got, ref)
    log(f"{name}.vae_decode.R{R}", rel_l2=e, max_abs=(got - ref).abs().max().item())
    assert e < 1e-2


def check_unet(b, R, B=1, name="tiny", n_steps=4, steps=(0, 3)):
    from oracle import unet as un
    from oracle.ddim import DDIM
    h = R // 8
    sample = torch.randn(3 * B, 9, h, h, generator=gen(7)).to(DEV)
    emb = torch.randn(1, 14, b.cfg.unet.cross_dim, generator=gen(8)).to(DEV)
    unc = torch.randn(1, 14, b.cfg.unet.cross_dim, generator=gen(9)).to(DEV)
    b.engine.set_condition(emb[0], unc[0])
    d = DDIM()
    d.set_timesteps(n_steps)
    ts = [float(t) for t in d.timesteps]
    b.engine.set_schedule(ts, [0.5] * n_steps, [0.6] * n_steps, 2.0, 1.0, n_steps)
    ctx = torch.cat([unc.expand(B, -1, -1), emb.expand(B, -1, -1), emb.expand(B, -1, -1)]).half().float()
    for step in steps:
        got = b.engine.unet_forward(sample, step)
        ref = un.unet_forward(b.oracle_sds[0], b.cfg.unet, sample.half().float(), ts[step], ctx)
        e = rel_l2(got, ref)
        log(f"{name}.unet.R{R}.B{B}.step{step}", rel_l2=e, ref_std=ref.std().item())
        assert e < 1e-2


def check_encoder(b, name):
    from oracle import image_encoder as ie
    patches = torch.randn(14, 3, 224, 224, generator=gen(11)).to(DEV)
    got = b.engine.encode_patches(patches)
    ref, _ = ie.encoder_forward(b.oracle_sds[2], b.cfg.enc, patches.half().float())
    e = rel_l2(got, ref[0])
    log(f"{name}.image_encoder", rel_l2=e)
    assert e < 1e-2


def check_e2e(b, R, steps, B=1, name="tiny", psnr_min=40.0, frac_min=0.999, strict=False, pad=None, u8=True):
    """generate_raw (before the composite, where every pixel is generated) against the fp32 oracle: PSNR and the fraction
    of pixels within 2/255; then the composited float output and the uint8 wire output (handler.py:55-56 truncation)."""
    from diffusiontexturepainting_b200.trt_model import TRTConditionalInpainter
    model = TRTConditionalInpainter(R, device=0, model_config=b.cfg, state_dicts=b.sds, max_batch_size=B)
    model.pipeline.sample_posterior = False
    model.pipeline.strict_schedule = strict
    brush = smooth_image(1, 3, R + 16)
    model.set_brush(brush)
    ora = b.oracle(R)
    ora.set_brush(brush)
    e_emb = rel_l2(model.conditioning[0], ora.conditioning[0])
    canvas = make_canvas(B, R)
    lat = torch.randn(B, 4, R // 8, R // 8, generator=gen(42))
    settings = dict(steps=steps, context_pad=pad if pad is not None else R // 2, tg_steps=steps, width=R, cfg_weight=2.0,
                    tg_weight=1.0)
    with torch.inference_mode():
        ref_raw = ora.generate_raw(canvas, lat.to(DEV), strict=strict, **settings).cpu()
    got_raw = model.generate_raw(canvas, init_latents=lat, **settings).cpu()
    mse = ((got_raw - ref_raw) ** 2).mean().item()
    psnr = 10 * math.log10(1.0 / max(mse, 1e-20))
    frac = ((got_raw - ref_raw).abs() <= 2.0 / 255).float().mean().item()
    got = model.generate(canvas, init_latents=lat, **settings).cpu()
    alpha = canvas[:, 3:]
    ref = canvas[:, :3] * alpha + ref_raw * (1 - alpha)
    comp_max = (got - ref).abs().max().item()
    n_eval = steps if strict else steps - 1
    tag = f"{name}.e2e.R{R}.S{steps}{'strict' if strict else ''}.B{B}"
    rec = dict(psnr_raw=psnr, frac_within_2_255_raw=frac, emb_rel_l2=e_emb, unet_evaluations=n_eval,
               composite_max_abs=comp_max, launches=model.engine.counter("launches"))
    if u8:
        canvas_u8 = (canvas.permute(0, 2, 3, 1) * 255).to(torch.uint8)
        out_u8 = model.stamp_u8(canvas_u8, init_latents=lat, **settings).cpu()
        cq = canvas_u8.permute(0, 3, 1, 2).float() / 255
        with torch.inference_mode():
            ref_q = ora.generate(cq, lat.to(DEV), strict=strict, **settings).cpu()
        ref_u8 = (ref_q * 255).to(torch.uint8).permute(0, 2, 3, 1)
        d8 = (out_u8.int() - ref_u8.int()).abs()
        rec.update(u8_max_lsb=int(d8.max()), u8_frac_within_2=(d8 <= 2).float().mean().item())
    log(tag, **rec)
    assert torch.isfinite(got_raw).all() and torch.isfinite(got).all()
    assert e_emb < 1e-2
    assert psnr >= psnr_min, rec
    assert frac >= frac_min, rec
    if u8:
        assert rec["u8_frac_within_2"] >= frac_min, rec
    model.pipeline.teardown()
    del model
    torch.cuda.empty_cache()
    return got, ref


def test_tiny_vae(tiny):
    check_vae(tiny, 64)
    check_vae(tiny, 128, B=1)


def test_tiny_unet(tiny):
    check_unet(tiny, 64)
    check_unet(tiny, 128, B=2)


def test_tiny_encoder(tiny):
    check_encoder(tiny, "tiny")


def test_tiny_e2e(tiny):
    check_e2e(tiny, 64, 4)
    check_e2e(tiny, 128, 6, B=2)


def test_full_unet(full):
    check_unet(full, 64, name="sd15")
    check_unet(full, 256, name="sd15")


def test_full_fused_cross_attention_matches_the_two_kernel_path(full):
    """cross_attn_kernel (scores + softmax + output in one launch; needs 8 heads, so only the full model uses it) against
    the two-contraction path on the same UNet evaluation: levels with 4096 / 1024 / 256 / 64 rows per sample group (R = 512)
    and ragged tiles (R = 64: 64 / 16 / 4 / 1 rows)."""
    h_sizes = (512, 64)
    for R in h_sizes:
        h = R // 8
        sample = torch.randn(3, 9, h, h, generator=gen(17)).to(DEV)
        emb = torch.randn(14, full.cfg.unet.cross_dim, generator=gen(18)).to(DEV)
        full.engine.set_condition(emb, emb * 0.3)
        full.engine.set_schedule([501.0], [0.5], [0.6], 2.0, 1.0, 1)
        outs = []
        for on in (1, 0):
            full.engine.set_option("fuse_cross", on)
            outs.append(full.engine.unet_forward(sample, 0).clone())
        full.engine.set_option("fuse_cross", 1)
        e = rel_l2(outs[0], outs[1])
        log(f"sd15.cross_fused_vs_split.R{R}", rel_l2=e)
        assert e < 2e-3


def test_full_vae(full):
    check_vae(full, 128, B=1, name="sd15")


def test_full_encoder(full):
    check_encoder(full, "sd15")


def test_full_e2e(full):
    check_e2e(full, 128, 5, name="sd15")


# ---- the benchmarked configurations at full model size (BASELINE.json configs 2 and 3; inpaint_pipeline.py:52-153) ----
def test_full_unet_c2_512(full):
    """one UNet evaluation at 512 x 512 / B = 1 (12 288 rows at level 0): first and last entry of a 20-entry schedule; this
    is where the measured tile table (320-wide pair tiles, in-kernel split-K) is exercised"""
    check_unet(full, 512, name="sd15", n_steps=20, steps=(0, 19))


def test_full_unet_c3_256_b4(full):
    check_unet(full, 256, B=4, name="sd15", n_steps=10, steps=(0, 9))


def test_full_vae_c2_512(full):
    check_vae(full, 512, B=2, name="sd15")


def test_full_vae_c3_256(full):
    check_vae(full, 256, B=2, name="sd15")


def test_full_e2e_c2_reference_semantics(full):
    """512 x 512, steps = 20 through the facade = 19 evaluations (t_start = 1), context_pad 150: the server's call"""
    check_e2e(full, 512, 20, name="sd15", pad=150)


def test_full_e2e_c2_strict(full):
    """512 x 512, all 20 evaluations: the configuration bench.py times"""
    check_e2e(full, 512, 20, name="sd15", strict=True, pad=150, u8=False)


def test_full_e2e_c3_per_gpu(full):
    """256 x 256, B = 4 stamps, 10 evaluations (strict): the per-GPU share of BASELINE config 3"""
    check_e2e(full, 256, 10, B=4, name="sd15", strict=True, pad=150)


def test_full_branch_dedup_matches_three_branch_evaluation(full):
    """The uncond and the cond branch share their UNet sample input (inpaint_pipeline.py:116-138), so the layers in front of
    the first cross-attention run once for both (option dedup_branches; needs the fused cross-attention kernel, i.e. the
    8-head full model). Against the plain three-branch plan on the same stamps, B = 1 and B = 2; dtp_unet_forward with a
    caller-supplied sample tensor must NOT de-duplicate (its three groups are arbitrary)."""
    from diffusiontexturepainting_b200.trt_model import TRTConditionalInpainter
    steps = 4
    # R = 128: the level-0 block has <= 1024 rows, so norm3 is folded too and the cross-attention kernel also emits the row
    # statistics of its three-group output (separate buffer from the two-group statistics it reads)
    for R, B in ((256, 1), (256, 2), (128, 1)):
        model = TRTConditionalInpainter(R, device=0, model_config=full.cfg, state_dicts=full.sds, max_batch_size=B)
        model.pipeline.sample_posterior = False
        model.pipeline.strict_schedule = True
        model.set_brush(smooth_image(1, 3, R))
        canvas = make_canvas(B, R)
        lat = torch.randn(B, 4, R // 8, R // 8, generator=gen(42))
        settings = dict(steps=steps, context_pad=R // 3, tg_steps=2, width=R, cfg_weight=2.0, tg_weight=1.0)
        eng = model.engine
        outs, ops = {}, {}
        for on in (1, 0):
            eng.set_option("dedup_branches", on)
            outs[on] = model.generate_raw(canvas, init_latents=lat, **settings).clone()
            ops[on] = eng.counter("unet_plan_ops")
        eng.set_option("dedup_branches", 1)
        assert ops[1] == ops[0] + 1  # the residual of the first transformer block is expanded to three groups
        e = rel_l2(outs[1], outs[0])
        log(f"sd15.branch_dedup.R{R}.B{B}", rel_l2=e)
        assert e < 2e-3
        if B == 1 and R == 256:
            h = R // 8
            sample = torch.randn(3, 9, h, h, generator=gen(17)).to(DEV)  # three DIFFERENT groups
            a = eng.unet_forward(sample, 0).clone()
            assert eng.counter("unet_plan_ops") == ops[0]
            eng.set_option("dedup_branches", 0)
            bb = eng.unet_forward(sample, 0).clone()
            eng.set_option("dedup_branches", 1)
            assert torch.equal(a, bb)
        model.pipeline.teardown()
        del model
        torch.cuda.empty_cache()


def test_graph_replay_and_options_are_equivalent(tiny):
    """CUDA-graph replay, eager launch order, unfolded cross-attention and the materialised-score attention path must all
    produce the same stamp (bit-identical for graph vs eager; <= 2e-3 rel-L2 across algorithmic variants)."""
    from diffusiontexturepainting_b200.trt_model import TRTConditionalInpainter
    R, steps, B = 128, 4, 2
    model = TRTConditionalInpainter(R, device=0, model_config=tiny.cfg, state_dicts=tiny.sds, max_batch_size=B)
    model.pipeline.sample_posterior = False
    model.set_brush(smooth_image(1, 3, R))
    canvas = make_canvas(B, R)
    lat = torch.randn(B, 4, R // 8, R // 8, generator=gen(42))
    settings = dict(steps=steps, context_pad=30, tg_steps=2, width=R, cfg_weight=2.5, tg_weight=0.7)
    eng = model.engine
    outs = {}
    eng.set_option("graph", 0)
    outs["eager"] = model.generate(canvas, init_latents=lat, **settings).clone()
    eng.set_option("graph", 1)
    for i in range(3):  # warm (eager), capture, replay
        outs[f"graph{i}"] = model.generate(canvas, init_latents=lat, **settings).clone()
    assert eng.counter("graph_launches") >= 2
    for i in range(3):
        assert torch.equal(outs["eager"], outs[f"graph{i}"])
    eng.set_option("fold_cross", 0)
    outs["unfolded"] = model.generate(canvas, init_latents=lat, **settings).clone()
    eng.set_option("fold_cross", 1)
    eng.set_option("flash", 0)
    outs["noflash"] = model.generate(canvas, init_latents=lat, **settings).clone()
    eng.set_option("flash", 1)
    eng.set_option("fold_ln", 0)   # standalone LayerNorm kernels instead of the LayerNorm folded into QKV / scores / FF1
    outs["ln_kernels"] = model.generate(canvas, init_latents=lat, **settings).clone()
    n_unfolded = eng.counter("unet_plan_ops")
    eng.set_option("fold_ln", 1)
    outs["ln_folded"] = model.generate(canvas, init_latents=lat, **settings).clone()
    assert eng.counter("unet_plan_ops") < n_unfolded  # three launches fewer per transformer block
    eng.set_option("fuse_shortcut", 0)   # conv_shortcut as its own 1x1 contraction + residual read in conv2
    outs["shortcut_separate"] = model.generate(canvas, init_latents=lat, **settings).clone()
    eng.set_option("fuse_shortcut", 1)
    eng.set_option("fuse_ff_out", 0)   # ff.net.2 and proj_out as two contractions
    outs["ff_out_separate"] = model.generate(canvas, init_latents=lat, **settings).clone()
    eng.set_option("fuse_ff_out", 1)
    eng.set_option("fuse_cross", 0)   # cross-attention as two contractions (scores with softmax epilogue, output)
    outs["cross_two_kernels"] = model.generate(canvas, init_latents=lat, **settings).clone()
    eng.set_option("fuse_cross", 1)
    model.generate(canvas, init_latents=lat, **settings)  # defaults again: at this size no upsample is folded
    n_ops = eng.counter("unet_plan_ops")
    eng.set_option("fold_upsample_rows", 0)   # nearest-2x upsample folded into its convolution at every level
    outs["upsample_folded"] = model.generate(canvas, init_latents=lat, **settings).clone()
    assert eng.counter("unet_plan_ops") == n_ops - 3  # three upsample launches fewer
    eng.set_option("fold_upsample", 0)        # ... and at none
    outs["upsample_separate"] = model.generate(canvas, init_latents=lat, **settings).clone()
    eng.set_option("fold_upsample", 1)
    eng.set_option("fold_upsample_rows", 3072)
    eng.set_option("fold_downsample", 0)      # stride-2 convolutions as a gather kernel + linear contraction
    outs["downsample_im2col"] = model.generate(canvas, init_latents=lat, **settings).clone()
    assert eng.counter("unet_plan_ops") == n_ops + 3
    eng.set_option("fold_downsample", 1)
    eng.set_option("splitk_f16", 0)           # fp32 partials in the in-kernel split-K reduction
    outs["splitk_f32_partials"] = model.generate(canvas, init_latents=lat, **settings).clone()
    eng.set_option("splitk_f16", 1)
    for k in ("unfolded", "noflash", "ln_kernels", "ln_folded", "shortcut_separate", "ff_out_separate", "cross_two_kernels",
              "upsample_folded", "upsample_separate", "downsample_im2col", "splitk_f32_partials"):
        e = rel_l2(outs[k], outs["eager"])
        log(f"tiny.variant.{k}", rel_l2=e)
        assert e < 2e-3
    # u8 fast path == float path composited and truncated (handler.py:55-56)
    u8 = model.stamp_u8((canvas.permute(0, 2, 3, 1) * 255).to(torch.uint8), init_latents=lat, **settings)
    canvas_q = (canvas.permute(0, 2, 3, 1) * 255).to(torch.uint8).permute(0, 3, 1, 2).float() / 255
    ref = model.generate(canvas_q, init_latents=lat, **settings)
    ref_u8 = (ref * 255).to(torch.uint8).permute(0, 2, 3, 1)
    diff = (u8.int() - ref_u8.int()).abs().max().item()
    log("tiny.stamp_u8.max_lsb_diff", diff=diff)
    assert diff <= 1
    model.pipeline.teardown()


def test_reference_step_semantics_and_strict_schedule(tiny):
    """steps = S runs S - 1 evaluations (t_start = 1 quirk, stable_diffusion_pipeline.py:348-355); strict runs S; tg is
    forced to zero from evaluation index tg_steps on (stable_diffusion_pipeline.py:419-420). Checked against the oracle."""
    from diffusiontexturepainting_b200.trt_model import TRTConditionalInpainter
    R = 64
    model = TRTConditionalInpainter(R, device=0, model_config=tiny.cfg, state_dicts=tiny.sds)
    model.pipeline.sample_posterior = False
    brush = smooth_image(1, 3, R)
    model.set_brush(brush)
    ora = tiny.oracle(R)
    ora.set_brush(brush)
    canvas = make_canvas(1, R)
    lat = torch.randn(1, 4, R // 8, R // 8, generator=gen(3))
    for strict in (False, True):
        for steps, tg_steps in ((3, 1), (1, 1), (5, 0)):
            model.pipeline.strict_schedule = strict
            s = dict(steps=steps, context_pad=10, tg_steps=tg_steps, width=R, cfg_weight=2.0, tg_weight=1.5)
            got = model.generate_raw(canvas, init_latents=lat, **s).cpu()
            ref = ora.generate_raw(canvas, lat.to(DEV), strict=strict, **s).cpu()
            e = rel_l2(got, ref)
            log(f"tiny.schedule.strict{int(strict)}.S{steps}.tg{tg_steps}", rel_l2=e)
            assert e < 5e-3
    model.pipeline.teardown()
