"""One small end-to-end stamp for compute-sanitizer (memcheck / racecheck / synccheck): tiny configuration, B = 2,
128 x 128, 3 evaluations, eager launches (no graph) so every kernel instantiation the tiny model uses runs under the tool:
single-CTA and CTA-pair contraction tiles, the in-kernel split-K reduction (tickets, fp16 partials), the folded upsample and
stride-2 convolutions, flash attention, both GroupNorm kernels (per-group and per-sample barrier), canvas pre-process,
composite. FULL=1 runs one stamp of the full-width model instead (fused cross-attention kernel, branch de-duplication,
320-wide pair tiles), which the 4-head tiny configuration does not reach.
    compute-sanitizer --tool memcheck  python profiles/sanitize_stamp.py
    compute-sanitizer --tool racecheck python profiles/sanitize_stamp.py
    FULL=1 R=64 compute-sanitizer --tool memcheck python profiles/sanitize_stamp.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusiontexturepainting_b200 import weights as W  # noqa: E402
from diffusiontexturepainting_b200.testdata import make_canvas, smooth_image  # noqa: E402
from diffusiontexturepainting_b200.trt_model import TRTConditionalInpainter  # noqa: E402

FULL = os.environ.get("FULL", "0") == "1"  # SD-1.5 widths (8 heads): the fused cross-attention kernel, branch de-duplication,
R, B, S = int(os.environ.get("R", "128")), (1 if FULL else 2), (2 if FULL else 3)  # 320-wide pair tiles; one stamp, two evaluations
cfg = W.sd15_config() if FULL else W.tiny_config()
model = TRTConditionalInpainter(R, device=0, model_config=cfg, state_dicts=W.synth_model(cfg), max_batch_size=B)
model.pipeline.sample_posterior = False
model.pipeline.strict_schedule = True
model.engine.set_option("graph", 0)
model.engine.set_option("fold_upsample_rows", 0)  # the folded upsample convolutions too (off at this size by default)
model.set_brush(smooth_image(1, 3, R))
canvas = make_canvas(B, R)
lat = torch.randn(B, 4, R // 8, R // 8, generator=torch.Generator().manual_seed(42))
out = model.generate(canvas, init_latents=lat, steps=S, context_pad=40, tg_steps=S, width=R, cfg_weight=2.0, tg_weight=1.0)
torch.cuda.synchronize()
u8 = model.stamp_u8((canvas.permute(0, 2, 3, 1) * 255).to(torch.uint8), init_latents=lat, steps=S, context_pad=40,
                    tg_steps=S, width=R, cfg_weight=2.0, tg_weight=1.0)
torch.cuda.synchronize()
assert torch.isfinite(out).all()
print("sanitize_stamp: ok", tuple(out.shape), tuple(u8.shape), model.engine.counter("launches"), "launches")
