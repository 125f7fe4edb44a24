"""Deterministic synthetic stamp inputs shared by the tests, the smoke test and the benchmark: brush and
canvas RGB = low-pass uniform noise in [0,1]; alpha = top 40 % of rows known (a deterministic stand-in for the
draw-down masks of training/mask_generator.py:22-71)."""
import torch


def smooth_image(seed, c, r):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(1, c, r // 8 + 1, r // 8 + 1, generator=g)
    return torch.nn.functional.interpolate(x, size=(r, r), mode="bilinear", align_corners=True)[0].clamp(0, 1)


def make_canvas(B, R, seed=2):
    canvas = torch.stack([torch.cat([smooth_image(seed + i, 3, R), torch.zeros(1, R, R)]) for i in range(B)])
    canvas[:, 3, :int(0.4 * R)] = 1.0
    return canvas
