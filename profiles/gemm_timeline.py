"""Per-CTA phase timeline of the contraction kernel from its globaltimer checkpoints (dtp_ops_set_debug_buffer):
0 start, 1 setup done (barriers, TMEM), 2 first operands landed, 3 all MMAs issued, 4 first accumulator ready,
5 epilogue of the last tile done, 6 TMEM released.   python profiles/gemm_timeline.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusiontexturepainting_b200 import _native as nat  # noqa: E402

L = nat.lib()
dev = "cuda"
dbg = torch.zeros(4096 * 8, dtype=torch.int64, device=dev)
BN = int(os.environ.get("BN", "0"))
SHAPES = [(12288, 320, 320, 0), (3072, 640, 640, 0), (768, 1280, 1280, 0), (12288, 960, 320, 0), (12288, 2560, 320, 8),
          (192, 1280, 1280, 0), (192, 1280, 11520, 0), (768, 1280, 11520, 0)]
for (M, N, K, flags) in SHAPES:
    A = torch.randn(M, K, device=dev).half()
    Wt = (torch.randn(N, K, device=dev) * 0.05).half()
    bias = torch.randn(N, device=dev)
    Nout = N // 2 if flags & 8 else N
    out = torch.empty(M, Nout, device=dev, dtype=torch.float16)

    def call():
        nat.check_op(L.dtp_op_linear(nat.ptr(A), K, K, None, 0, 0, M, nat.ptr(Wt), K, N, nat.ptr(bias), None, 0,
                                     nat.ptr(out), Nout, flags, 1.0, 0, BN, 0, nat.stream_ptr()), "linear")
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    err = float("nan")
    if not flags & 8:
        ref = A.float() @ Wt.float().t() + bias
        err = ((out.float() - ref).norm() / ref.norm()).item()
    # device-bound period: 50 launches captured into one CUDA graph (PDL edges kept); an eager Python loop is CPU-bound
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for _ in range(50):
                call()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(4):
            graph.replay()
        e1.record(side)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 200
    L.dtp_ops_set_debug_buffer(nat.ptr(dbg))
    dbg.zero_()
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    L.dtp_ops_set_debug_buffer(None)
    t = dbg.view(-1, 8).cpu()
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = (t[:, :7] - t0).float() / 1e3
    names = ["start", "setup", "operands", "mma issued", "acc ready", "epilogue", "released"]
    print("linear M=%d N=%d K=%d flags=%#x BN=%d light=%s: %.1f us per launch in a graph (%.0f TFLOP/s), CTAs=%d rel=%.1e"
          % (M, N, K, flags, BN, os.environ.get("DTP_GEMM_LIGHT", "12"), us, 2.0 * M * N * K / us / 1e6, t.shape[0], err))
    print("   " + "  ".join("%s %.2f/%.2f" % (names[k], rel[:, k].median(), rel[:, k].max()) for k in range(7)), flush=True)
