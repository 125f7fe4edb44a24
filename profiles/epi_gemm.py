"""Epilogue-bound contractions (short K, wide N) launched repeatedly - target for `ncu --set full`.
    python profiles/epi_gemm.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusiontexturepainting_b200 import _native as nat  # noqa: E402

L = nat.lib()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for (M, N, K, BN, flags) in [(12288, 2560, 320, 256, 8), (12288, 960, 320, 256, 0), (12288, 320, 320, 128, 0), (3072, 5120, 640, 256, 8)]:
    A = torch.randn(M, K, device="cuda").half()
    W = torch.randn(N, K, device="cuda").half()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").half() if not flags else None
    out = torch.empty(M, N // 2 if flags & 8 else N, device="cuda", dtype=torch.float16)

    def call():
        nat.check_op(L.dtp_op_linear(nat.ptr(A), K, K, None, 0, 0, M, nat.ptr(W), K, N, nat.ptr(bias), nat.ptr(res) if res is not None else None,
                                     N if res is not None else 0, nat.ptr(out), out.shape[1], flags, 1.0, 0, BN, 1, nat.stream_ptr()))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    ob = out.numel() * 2 + (res.numel() * 2 if res is not None else 0) + A.numel() * 2
    print(f"M={M} N={N} K={K} BN={BN} flags={flags}: {us:.1f} us  {2.0*M*N*K/us/1e6:.0f} TFLOP/s  {ob/us/1e6:.2f} TB/s (A + residual + out)", flush=True)
