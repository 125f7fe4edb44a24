"""Upsample2D (nearest 2x + conv3x3) as one folded contraction (dtp_op_upconv2x) vs the two-kernel path (upsample kernel + 3x3
contraction), over the shapes of the UNet up path (512 / 256 stamps, 3 and 12 samples) and of the VAE decoder; every tile width.
    python profiles/upconv_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusiontexturepainting_b200 import _native as nat  # noqa: E402

L = nat.lib()
dev = "cuda"
PAIR = 0x1000
SHAPES = [  # (Nimg, H, W, C)  (input resolution)
    (3, 32, 32, 640), (3, 16, 16, 1280), (3, 8, 8, 1280), (3, 16, 16, 640), (12, 16, 16, 640), (12, 8, 8, 1280),
    (12, 4, 4, 1280), (1, 64, 64, 512), (1, 128, 128, 512), (1, 256, 256, 256),
]
BNS = [0, 128, 160, 192, 256, 128 | PAIR, 256 | PAIR, 320 | PAIR]


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(3):
            g.replay()
        e1.record(side)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * reps)


for n, H, W, C in SHAPES:
    torch.manual_seed(0)
    x = torch.randn(n, H, W, C, device=dev).half()
    w = (torch.randn(C, 9 * C, device=dev) * (9 * C) ** -0.5).half()
    b = torch.randn(C, device=dev)
    wst = torch.empty(4 * C, 4 * C, device=dev, dtype=torch.float16)
    up = torch.empty(n, 2 * H, 2 * W, C, device=dev, dtype=torch.float16)
    out = torch.empty(n, 2 * H, 2 * W, C, device=dev, dtype=torch.float16)
    out2 = torch.empty_like(out)

    def two():
        nat.check_op(L.dtp_op_upsample2x(nat.ptr(x), n, H, W, C, nat.ptr(up), nat.stream_ptr()), "up")
        nat.check_op(L.dtp_op_conv3x3(nat.ptr(up), C, None, 0, n, 2 * H, 2 * W, nat.ptr(w), C, nat.ptr(b), None, 0, nat.ptr(out2),
                                      C, 0, 1.0, 0, 0, 1, nat.stream_ptr()), "conv")
    t2 = timed(two)
    line = "N=%2d in=%3dx%3d C=%4d  two-kernel %7.2f us |" % (n, H, W, C, t2)
    for BN in BNS:
        if (BN & 0xfff) == 320 and C % 320:
            continue

        # fold the weights once (Wt given), then time the contraction alone (Wt = NULL: wstack is already folded)
        nat.check_op(L.dtp_op_upconv2x(nat.ptr(x), C, n, H, W, nat.ptr(w), C, nat.ptr(b), nat.ptr(wst), nat.ptr(out), BN,
                                       nat.stream_ptr()), "upconv")

        def one():
            nat.check_op(L.dtp_op_upconv2x(nat.ptr(x), C, n, H, W, None, C, nat.ptr(b), nat.ptr(wst), nat.ptr(out), BN,
                                           nat.stream_ptr()), "upconv")
        try:
            t1 = timed(one)
        except Exception as exc:  # noqa: BLE001
            line += " %s:ERR" % (BN & 0xfff)
            print(exc)
            continue
        err = ((out.float() - out2.float()).norm() / out2.float().norm()).item()
        line += " %d%s: %.2f (%.0e)" % (BN & 0xfff, "p" if BN & PAIR else "", t1, err)
    print(line, flush=True)
