"""Weight-streaming shapes with COLD weights (ring of buffers > L2) and the per-CTA globaltimer checkpoints of the
contraction kernel.   python profiles/ws_stages.py"""
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
    nbuf = max(2, int(400e6 // (N * K * 2)))
    A = torch.randn(M, K, device="cuda").half()
    Ws = [torch.randn(N, K, device="cuda").half() for _ in range(nbuf)]
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    nct = ((M + 127) // 128) * ((N + BN - 1) // BN) * sp
    dbg = torch.zeros(max(nct, 148), 8, dtype=torch.int64, device="cuda")

    def call(i):
        nat.check_op(L.dtp_op_linear(nat.ptr(A), K, K, None, 0, 0, M, nat.ptr(Ws[i % nbuf]), K, N, None, None, 0,
                                     nat.ptr(out), N, 0, 1.0, 0, BN, sp, nat.stream_ptr()))
    for i in range(nbuf):
        call(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3 * nbuf
    e0.record()
    for i in range(reps):
        call(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    L.dtp_ops_set_debug_buffer(C.c_void_p(dbg.data_ptr()))
    call(1)
    torch.cuda.synchronize()
    L.dtp_ops_set_debug_buffer(None)
    d = dbg[:min(nct, 148)].cpu().double()
    t0 = d[:, 0].min()
    rel = d[:, :7] - t0
    med = [float(rel[:, i].median()) for i in range(7)]
    mx = [float(rel[:, i].max()) for i in range(7)]
    print(f"M={M} N={N} K={K} BN={BN} sp={sp}: {us:.1f} us/launch (incl. finalize), {N*K*2/us/1e6:.2f} TB/s weights; ctas={nct}")
    print("   median ns since first CTA start: start/setup/first_mma/mma_done/acc_seen/epi_done/exit =", [int(x) for x in med])
    print("   max    ns:", [int(x) for x in mx])


if __name__ == "__main__":
    for cfg in [(192, 1280, 11520, 128, 7), (192, 1280, 11520, 256, 8), (192, 1280, 11520, 64, 7), (192, 1280, 11520, 128, 1),
                (768, 1280, 11520, 256, 4), (192, 1280, 1280, 64, 1), (12288, 320, 320, 128, 1)]:
        run(*cfg)
