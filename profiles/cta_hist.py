"""Per-CTA checkpoint distribution of one contraction launch (percentiles over CTAs).
    python profiles/cta_hist.py M N K BN splits"""
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
    dbg = torch.zeros(2 * 148 + 8, 8, dtype=torch.int64, device="cuda")

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
    d = dbg[:148].cpu().double()
    d = d[d[:, 0] > 0]
    t0 = d[:, 0].min()
    rel = (d[:, :7] - t0) / 1e3
    names = ["start", "setup", "first_mma", "mma_done", "acc_seen", "epi_done", "exit"]
    print(f"M={M} N={N} K={K} BN={BN & 0xfff}{'p' if BN & 0x1000 else ''} sp={sp}: {d.shape[0]} CTAs reporting")
    for i, n in enumerate(names):
        col = rel[:, i]
        col = col[col > -1e6]
        q = torch.quantile(col, torch.tensor([0.0, 0.25, 0.5, 0.75, 1.0], dtype=torch.double))
        print(f"   {n:10s} min/25/50/75/max us: " + " ".join(f"{float(x):7.2f}" for x in q))


if __name__ == "__main__":
    M, N, K, BN, sp = (int(x, 0) for x in sys.argv[1:6])
    run(M, N, K, BN, sp)
