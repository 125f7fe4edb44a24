"""Contraction kernel micro-benchmark: times one problem shape per line with CUDA events and, with --stages, dumps the
per-CTA globaltimer checkpoints (setup / first load / mainloop / epilogue / teardown).
    python profiles/gemm_microbench.py [--stages]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusiontexturepainting_b200 import _native as nat  # noqa: E402

L = nat.lib()
L.dtp_ops_set_debug_buffer.argtypes = [C.c_void_p]
L.dtp_ops_set_debug_buffer.restype = None
dev = "cuda"


def run_linear(M, N, K, BN, splits=1, flags=0, reps=20, stages=False):
    A = torch.randn(M, K, device=dev).half()
    W = torch.randn(N, K, device=dev).half()
    out = torch.empty(M, N // 2 if flags & 8 else N, device=dev, dtype=torch.float16)
    bias = torch.randn(N, device=dev)
    nct = ((M + 127) // 128) * ((N + max(BN, 32) - 1) // max(BN, 32)) * splits
    dbg = torch.zeros(nct, 8, dtype=torch.int64, device=dev)

    def call():
        rc = L.dtp_op_linear(nat.ptr(A), K, K, None, 0, 0, M, nat.ptr(W), K, N, nat.ptr(bias), None, 0, nat.ptr(out),
                             out.shape[1], flags, 1.0, 0, BN, splits, nat.stream_ptr())
        nat.check_op(rc)

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
    tf = 2.0 * M * N * K / us / 1e6
    line = f"linear M={M} N={N} K={K} BN={BN} splits={splits} flags={flags}: {us:8.1f} us  {tf:7.1f} TFLOP/s  ctas={nct}"
    if stages:
        L.dtp_ops_set_debug_buffer(C.c_void_p(dbg.data_ptr()))
        call()
        torch.cuda.synchronize()
        L.dtp_ops_set_debug_buffer(None)
        d = dbg.cpu().double()
        t0 = d[:, 0].min()
        rel = (d[:, :7] - d[:, :1])
        names = ["setup", "first_load", "mma_issued", "acc_ready", "epilogue", "teardown"]
        cols = [rel[:, 1], rel[:, 2] - rel[:, 1], rel[:, 3] - rel[:, 2], rel[:, 4] - rel[:, 3], rel[:, 5] - rel[:, 4],
                rel[:, 6] - rel[:, 5]]
        line += "\n    per-CTA ns (median): " + ", ".join(f"{n}={float(c.median()):.0f}" for n, c in zip(names, cols))
        line += f"; cta lifetime median={float(rel[:, 6].median()):.0f} ns; kernel span={float(d[:, 6].max() - t0):.0f} ns"
        line += f"; start spread={float(d[:, 0].max() - t0):.0f} ns"
    print(line, flush=True)


if __name__ == "__main__":
    st = "--stages" in sys.argv
    for (M, N, K, BN) in [(12288, 320, 320, 256), (12288, 320, 320, 160), (12288, 320, 320, 64), (12288, 320, 1280, 160),
                          (12288, 320, 2880, 160), (12288, 320, 2880, 128), (768, 1280, 1280, 64), (768, 1280, 11520, 128),
                          (12288, 2560, 320, 256), (4096, 4096, 64, 256), (8192, 8192, 8192, 256),
                          (8192, 8192, 8192, 192), (8192, 8192, 8192, 128)]:
        run_linear(M, N, K, BN, stages=st)
    run_linear(12288, 2560, 320, 256, flags=8, stages=st)
    run_linear(192, 1280, 11520, 64, splits=7, stages=st)
