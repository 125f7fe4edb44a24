"""BASELINE config C4: VAE encode + decode only, 512x512, batch 16 (one encode + one decode per image).
    python profiles/vae_bench.py [--batch 16] [--resolution 512]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusiontexturepainting_b200 import weights as W  # noqa: E402
from diffusiontexturepainting_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--resolution", type=int, default=512)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
B, R = a.batch, a.resolution
cfg = W.sd15_config()
eng = Engine(cfg, 0, arena_bytes=int(os.environ.get("DTP_ARENA_BYTES", 40 << 30)))
eng.load_state_dicts(*W.synth_model(cfg))
g = torch.Generator().manual_seed(0)
x = (torch.rand(B, 3, R, R, generator=g) * 2 - 1).cuda()
z = (torch.randn(B, 4, R // 8, R // 8, generator=g) * 0.18215).cuda()
for _ in range(2):
    lat = eng.vae_encode(x)
    img = eng.vae_decode(z)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
enc_ms = dec_ms = 0.0
for _ in range(a.steps):
    e[0].record()
    lat = eng.vae_encode(x)
    e[1].record()
    img = eng.vae_decode(z)
    e[2].record()
    torch.cuda.synchronize()
    enc_ms += e[0].elapsed_time(e[1])
    dec_ms += e[1].elapsed_time(e[2])
enc_ms /= a.steps
dec_ms /= a.steps
GF_E = {512: 1116.7, 256: 272.7}[R]
GF_D = {512: 2514.5, 256: 622.2}[R]
print(json.dumps({"config": f"C4: VAE encode+decode only, {R}x{R}, batch {B}", "encode_ms": enc_ms, "decode_ms": dec_ms,
                  "encode_tflops": B * GF_E / enc_ms, "decode_tflops": B * GF_D / dec_ms,
                  "images_per_s": B / ((enc_ms + dec_ms) / 1e3), "finite": bool(torch.isfinite(img).all()),
                  "arena_peak_GiB": eng.counter("arena_peak") / 2 ** 30}))
