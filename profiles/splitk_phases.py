"""Phase timers of the in-kernel split-K reduction (second checkpoint bank of the contraction kernel).
    python profiles/splitk_phases.py"""
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


def run(M, N, K, BN, sp):
    A = torch.randn(M, K, device="cuda").half()
    W = torch.randn(N, K, device="cuda").half()
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    nct = ((M + 127) // 128) * ((N + BN - 1) // BN) * sp
    dbg = torch.zeros(2 * nct, 8, dtype=torch.int64, device="cuda")

    def call():
        nat.check_op(L.dtp_op_linear(nat.ptr(A), K, K, None, 0, 0, M, nat.ptr(W), K, N, None, None, 0, nat.ptr(out), N, 0, 1.0,
                                     0, BN, sp, nat.stream_ptr()))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    L.dtp_ops_set_debug_buffer(C.c_void_p(dbg.data_ptr()))
    call()
    torch.cuda.synchronize()
    L.dtp_ops_set_debug_buffer(None)
    d = dbg.cpu().double()
    a, b = d[:nct], d[nct:]
    t0 = a[:, 0].min()
    names1 = ["start", "setup", "first_mma", "mma_done", "acc_seen", "epi_done", "exit"]
    names2 = ["partials_stored", "fenced", "bar", "group_complete", "slices_landed", "summed"]
    print(f"M={M} N={N} K={K} BN={BN} sp={sp} ctas={nct}")
    print("   median ns:", ", ".join(f"{n}={float((a[:, i] - t0).median()):.0f}" for i, n in enumerate(names1)))
    print("   median ns:", ", ".join(f"{n}={float((b[:, i] - t0).median()):.0f}" for i, n in enumerate(names2)))


for cfg in [(3072, 640, 5760, 256, 2), (768, 1280, 11520, 256, 4), (192, 1280, 11520, 256, 14), (192, 1280, 11520, 128, 7)]:
    run(*cfg)
