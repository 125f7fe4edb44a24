"""Weight-streaming experiment: small-M contractions (the 8x8 / 16x16 UNet levels) with COLD weights (a ring of distinct
weight buffers larger than L2), row-major vs 64x64-blocked weight layout, over (BN, split-K).
    python profiles/weight_stream_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusiontexturepainting_b200 import _native as nat  # noqa: E402

L = nat.lib()
dev = "cuda"
W_BLOCKED = 1 << 10


def block(w):
    N, K = w.shape
    return w.view(N // 64, 64, K // 64, 64).permute(0, 2, 1, 3).contiguous()


def run(M, N, K, BN, sp, blocked, nbuf):
    A = torch.randn(M, K, device=dev).half()
    Ws = [torch.randn(N, K, device=dev).half() for _ in range(nbuf)]
    Wb = [block(w) if blocked else w for w in Ws]
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    flags = W_BLOCKED if blocked else 0

    def call(i):
        nat.check_op(L.dtp_op_linear(nat.ptr(A), K, K, None, 0, 0, M, nat.ptr(Wb[i % nbuf]), K, N, None, None, 0,
                                     nat.ptr(out), N, flags, 1.0, 0, BN, sp, nat.stream_ptr()))
    for i in range(nbuf):
        call(i)
    torch.cuda.synchronize()
    ref = A.float() @ Ws[(nbuf - 1) % nbuf].float().t()
    err = ((out.float() - ref).norm() / ref.norm()).item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3 * nbuf
    e0.record()
    for i in range(reps):
        call(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    return us, err


if __name__ == "__main__":
    for (M, N, K) in [(192, 1280, 11520), (768, 1280, 11520), (192, 1280, 1280), (768, 1280, 5120)]:
        wbytes = N * K * 2
        nbuf = max(2, int(400e6 // wbytes))
        for blocked in (0, 1):
            res = []
            for BN in (64, 128, 256):
                for sp in (1, 2, 4, 6, 8, 12, 16):
                    if (K // 64) // sp < 2:
                        continue
                    us, err = run(M, N, K, BN, sp, blocked, nbuf)
                    res.append((us, BN, sp, err))
            res.sort()
            best = ", ".join(f"BN={b} sp={s}: {u:.1f}us" for u, b, s, _ in res[:4])
            print(f"M={M} N={N} K={K} cold({nbuf} bufs) blocked={blocked}: {best}  | max err {max(r[3] for r in res):.1e} "
                  f"| weights at HBM peak {wbytes / 6.5e6:.1f}us", flush=True)
