"""GPU half of the drop-in test: real NEW_BRUSH and NEW_STAMP wire frames go through the reference handler's call sequence
(handler.py:92-123 — the UNMODIFIED handler when the reference tree and the stubs are available, else tests/handler_twin.py,
which tests/test_dropin_handler.py proves identical on the CPU) into TRTConditionalInpainter(256) exactly as run.py:30 builds
it (positional resolution only; the model draws its own latents from the seed-42 CUDA generator and samples the VAE
posterior). The response bytes are compared with the oracle run on the same random streams: <= 2 LSB on >= 99.9 % of the
bytes."""
import numpy as np
import pytest
import torch

import handler_twin as twin

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_wire_frames_through_the_handler_sequence(monkeypatch):
    from diffusiontexturepainting_b200 import trt_model as tm
    from diffusiontexturepainting_b200 import weights as W
    from oracle.pipeline import OraclePipeline
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = W.tiny_config()
    sds = W.synth_model(cfg)
    R = 256
    # run.py:30 calls TRTConditionalInpainter(256) with nothing else: the weights come from the loader; here the tiny
    # synthetic inventory is injected through the loader hook so the constructor call stays the reference's
    monkeypatch.setattr(tm, "_default_model", lambda: (cfg, sds), raising=False)
    model = tm.TRTConditionalInpainter(R)
    assert model.resolution() == R and model.device() == 0

    settings = dict(steps=8, context_pad=150, cfg_weight=2.0, tg_weight=1.0, tg_steps=8)
    brush_frame, stamp_frame = twin.synthetic_frames(R, seed=3, **settings)
    responses = [twin.handle_binary_request(model, f) for f in (brush_frame, stamp_frame)]

    # oracle on the same random streams: generator(42) for the initial latents, generator(43) for the posterior noise
    g_lat = torch.Generator(device=DEV).manual_seed(42)
    g_noise = torch.Generator(device=DEV).manual_seed(43)
    to = lambda sd: {k: t.to(DEV) for k, t in sd.items()}
    ora = OraclePipeline(cfg, to(W.round_fp16(W.merge_lora(sds[0]))), to(W.round_fp16(sds[1])), to(W.round_fp16(sds[2])),
                         R)
    brush = twin.np_to_torch(twin.binary_to_image(brush_frame, 14)[..., :3])
    ora.set_brush(brush)
    h = R // 8
    mask = torch.zeros(1, 1, R, R, device=DEV)
    mask[..., :R // 2, :R // 2] = 1
    contexts = [torch.cat([ora.image.float(), mask], dim=1),
                twin.np_to_torch(twin.binary_to_image(stamp_frame, 14)).unsqueeze(0).to(DEV)]
    kinds = (twin.RETURN_PREVIEW, twin.RETURN_STAMP)
    for resp, ctx, kind in zip(responses, contexts, kinds):
        lat = torch.randn((1, 4, h, h), device=DEV, dtype=torch.float32, generator=g_lat)
        nz = torch.randn((2, 4, h, h), device=DEV, dtype=torch.float32, generator=g_noise)
        with torch.inference_mode():
            ref = ora.generate(ctx, lat, vae_noise=(nz[:1], nz[1:]), **dict(settings, width=R)).cpu()
        ref_u8 = twin.torch_to_np(ref[0])
        assert resp[0] == kind
        got = twin.binary_to_image(resp, 1)
        assert got.shape == (R, R, 3) and len(resp) == 1 + 12 + R * R * 3
        d = np.abs(got.astype(np.int32) - ref_u8.astype(np.int32))
        frac = float((d <= 2).mean())
        print(f"dropin.{kind}: max LSB diff {d.max()}, within 2 LSB {frac:.5f}")
        assert frac >= 0.999, (d.max(), frac)
    model.pipeline.teardown()
