"""A/B of the in-kernel split-K reduction with fp32 vs fp16 partials (DTP_SPLITK_F16=0/1, read once per process): the weight-
streaming shapes of a 512x512 stamp, cold weights (ring of buffers > L2), 20 launches per CUDA graph.
    for v in 0 1; do DTP_SPLITK_F16=$v python profiles/splitk_f16_ab.py; done"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusiontexturepainting_b200 import _native as nat  # noqa: E402

L = nat.lib()
SHAPES = [(192, 1280, 11520, 192, 10), (192, 1280, 23040, 256, 14), (768, 1280, 11520, 320 | 0x1000, 6), (768, 1280, 6400, 160, 3),
          (3072, 640, 5760, 256, 2), (3072, 640, 11520, 320 | 0x1000, 3)]
print("DTP_SPLITK_F16=%s" % os.environ.get("DTP_SPLITK_F16", "1"))
for M, N, K, BN, sp in SHAPES:
    nbuf = max(2, int(400e6 // (N * K * 2)))
    torch.manual_seed(0)
    A = torch.randn(M, K, device="cuda").half()
    Ws = [(torch.randn(N, K, device="cuda") * K ** -0.5).half() for _ in range(nbuf)]
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    it = [0]

    def call():
        w = Ws[it[0] % nbuf]
        it[0] += 1
        nat.check_op(L.dtp_op_linear(nat.ptr(A), K, K, None, 0, 0, M, nat.ptr(w), K, N, None, None, 0, nat.ptr(out), N, 0, 1.0, 0,
                                     BN, sp, nat.stream_ptr()), "linear")
    call()
    torch.cuda.synchronize()
    ref = A.float() @ Ws[0].float().t()
    err = ((out.float() - ref).norm() / ref.norm()).item()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(20):
                call()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(5):
            g.replay()
        e1.record(side)
    torch.cuda.synchronize()
    print("M=%5d N=%5d K=%6d BN=%d splits=%2d  %.2f us  rel_l2=%.2e" % (M, N, K, BN & 0xfff, sp, e0.elapsed_time(e1) * 10, err), flush=True)
