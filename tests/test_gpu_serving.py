"""GPU tests of the serving helpers with the real engine: the stamp micro-batcher coalescing
concurrent stamps into one generate(B > 1) while set_brush interleaves on another thread, the arena growing on demand when
the model was built the way run.py builds it (no max_batch_size), the per-stage timers, and a checkpoint round trip
(`save_attn_procs`-style LoRA file + training-side image_encoder.pth) into the engine."""
import os
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def gen(s):
    return torch.Generator().manual_seed(s)


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def tiny_model():
    from diffusiontexturepainting_b200 import weights as W
    from diffusiontexturepainting_b200.trt_model import TRTConditionalInpainter
    cfg = W.tiny_config()
    sds = W.synth_model(cfg)
    m = TRTConditionalInpainter(64, device=0, model_config=cfg, state_dicts=sds)  # max_batch_size defaults to 1
    m.pipeline.sample_posterior = False
    yield m, cfg, sds
    m.pipeline.teardown()


def test_batcher_with_the_real_model_and_interleaved_set_brush(tiny_model):
    from diffusiontexturepainting_b200.serving import StampBatcher
    from diffusiontexturepainting_b200.testdata import make_canvas, smooth_image
    model, _, _ = tiny_model
    R = 64
    settings = dict(steps=3, context_pad=20, tg_steps=3, width=R, cfg_weight=2.0, tg_weight=1.0)
    brush_a, brush_b = smooth_image(1, 3, R), smooth_image(5, 3, R)
    canvases = make_canvas(6, R)
    lat = torch.randn(6, 4, R // 8, R // 8, generator=gen(1))

    # reference results, one stamp at a time, per brush (fixed latents through a wrapper so runs are comparable)
    def run_single(brush, i):
        model.set_brush(brush)
        return model.generate(canvases[i:i + 1], init_latents=lat[i:i + 1], **settings)[0].cpu()
    ref_a = [run_single(brush_a, i) for i in range(6)]

    rows = {}

    def generate(canv, **kw):  # the batcher hands over cat(canvases); find their rows to pick the matching latents
        idx = [next(j for j in range(6) if torch.equal(canv[i].cpu(), canvases[j])) for i in range(canv.shape[0])]
        rows.setdefault("batches", []).append(idx)
        return model.generate(canv, init_latents=lat[idx], **kw)

    model.set_brush(brush_a)
    with StampBatcher(generate, max_batch=8, max_wait_ms=100, brush_key=lambda: model.brush_generation,
                      lock=model.lock) as b:
        futs = [b.submit(canvases[i], settings) for i in range(6)]
        outs = [f.result(timeout=120).cpu() for f in futs]
        assert max(len(x) for x in rows["batches"]) > 1  # coalesced: the arena (sized for B = 1) grew on demand
        for i in range(6):
            assert rel_l2(outs[i], ref_a[i]) < 2e-3, i
        # a brush switch from another thread while stamps are in flight: every stamp either completes under the brush it
        # was submitted with, or fails loudly - never a torn or wrong brush
        futs = [b.submit(canvases[i], settings) for i in range(3)]
        t = threading.Thread(target=model.set_brush, args=(brush_b,))
        t.start()
        t.join()
        late = [b.submit(canvases[i], settings) for i in range(3)]
        ref_b = None
        for i, f in enumerate(futs):
            try:
                o = f.result(timeout=120).cpu()
                assert rel_l2(o, ref_a[i]) < 2e-3
            except RuntimeError as e:
                assert "brush changed" in str(e)
        outs_b = [f.result(timeout=120).cpu() for f in late]
    ref_b = [run_single(brush_b, i) for i in range(3)]
    for i in range(3):
        assert rel_l2(outs_b[i], ref_b[i]) < 2e-3, i
    assert model.engine.counter("arena_bytes") > 0


def test_stage_timers_and_summary(tiny_model, capsys):
    from diffusiontexturepainting_b200.testdata import make_canvas, smooth_image
    model, _, _ = tiny_model
    R = 64
    model.set_brush(smooth_image(1, 3, R))
    settings = dict(steps=4, context_pad=20, tg_steps=4, width=R, cfg_weight=2.0, tg_weight=1.0)
    u8 = (make_canvas(1, R).permute(0, 2, 3, 1) * 255).to(torch.uint8)
    model.enable_stage_timers(True)
    model.stamp_u8(u8, **settings)
    torch.cuda.synchronize()
    t = model.stage_times_ms()
    model.print_summary()
    model.enable_stage_timers(False)
    assert t["unet"][1] == 3 and t["vae_encoder"][1] == 1 and t["vae"][1] == 1 and t["canvas_preprocess"][1] == 1
    assert all(ms > 0 for ms, n in t.values() if n)
    assert "Pipeline" in capsys.readouterr().out


def test_checkpoint_files_round_trip_into_the_engine(tmp_path, monkeypatch):
    """Files in the formats the reference reads (diffusers-layout UNet / VAE, `save_attn_procs` LoRA, training-side
    image_encoder.pth with the transformers CLIP spelling) -> load_state_dicts -> engine: the UNet evaluation equals the one
    of an engine fed the same tensors directly."""
    from diffusiontexturepainting_b200 import weights as W
    from diffusiontexturepainting_b200.engine import Engine
    from diffusiontexturepainting_b200.stable_diffusion_pipeline import load_state_dicts
    tc = W.tiny_config()
    cfg = W.ModelConfig(unet=tc.unet, vae=tc.vae, enc=tc.enc, name="sd15-inpaint")
    u, v, e = W.synth_model(cfg, 77)
    hf = tmp_path / "hf"
    (hf / "unet").mkdir(parents=True)
    (hf / "vae").mkdir(parents=True)
    torch.save({k: t for k, t in u.items() if ".processor." not in k}, hf / "unet" / "diffusion_pytorch_model.bin")
    torch.save(v, hf / "vae" / "diffusion_pytorch_model.bin")
    torch.save({"unet." + k: t for k, t in u.items() if ".processor." in k}, tmp_path / "pytorch_lora_weights.bin")
    # training-side spelling of the CLIP tower (transformers.CLIPVisionModel keys)
    enc_file = {}
    w = cfg.enc.width
    for k, t in e.items():
        if not k.startswith("clip.visual."):
            enc_file[k] = t
            continue
        r = k[len("clip.visual."):]
        simple = {"class_embedding": "embeddings.class_embedding", "conv1.weight": "embeddings.patch_embedding.weight",
                  "positional_embedding": "embeddings.position_embedding.weight", "ln_pre.weight": "pre_layrnorm.weight",
                  "ln_pre.bias": "pre_layrnorm.bias", "ln_post.weight": "post_layernorm.weight",
                  "ln_post.bias": "post_layernorm.bias"}
        if r in simple:
            enc_file["clip.vision_model." + simple[r]] = t
            continue
        parts = r.split(".")
        i, name = parts[2], ".".join(parts[3:])
        base = f"clip.vision_model.encoder.layers.{i}."
        lm = {"ln_1": "layer_norm1", "ln_2": "layer_norm2", "attn.out_proj": "self_attn.out_proj", "mlp.c_fc": "mlp.fc1",
              "mlp.c_proj": "mlp.fc2"}
        if name.startswith("attn.in_proj_"):
            wb = name[len("attn.in_proj_"):]
            for j, n in enumerate(("q", "k", "v")):
                enc_file[base + f"self_attn.{n}_proj.{wb}"] = t[j * w:(j + 1) * w]
        else:
            stem, wb = name.rsplit(".", 1)
            enc_file[base + lm[stem] + "." + wb] = t
    torch.save(enc_file, tmp_path / "image_encoder.pth")
    monkeypatch.setenv("DTP_HF_DIR", str(hf))
    monkeypatch.setenv("DTP_IMAGE_ENCODER", str(tmp_path / "image_encoder.pth"))
    monkeypatch.delenv("DTP_SYNTHETIC_WEIGHTS", raising=False)
    lu, lv, le = load_state_dicts(cfg, str(tmp_path / "pytorch_lora_weights.bin"), seed=1)  # seed 1: nothing may survive
    for a, b in ((lu, u), (lv, v), (le, e)):
        assert set(a) == set(b) and all(torch.equal(a[k], b[k].float()) for k in b)
    outs = []
    for sds in ((lu, lv, le), (u, v, e)):
        eng = Engine(cfg, 0, arena_bytes=1 << 30)
        eng.load_state_dicts(*sds)
        emb = torch.randn(14, cfg.unet.cross_dim, generator=gen(2)).to(DEV)
        eng.set_condition(emb, emb * 0.5)
        eng.set_schedule([501.0], [0.5], [0.6], 2.0, 1.0, 1)
        outs.append(eng.unet_forward(torch.randn(3, 9, 8, 8, generator=gen(3)).to(DEV), 0).clone())
        eng.close()
    assert torch.equal(outs[0], outs[1])
