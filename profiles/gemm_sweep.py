"""Empirical (BN, split-K) sweep of the contraction kernel over the UNet / VAE problem shapes of the benchmark
workload; prints the best configuration per shape next to the heuristic's pick (gemm_pick_config).
    python profiles/gemm_sweep.py > gpurun_out/sweep.log"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusiontexturepainting_b200 import _native as nat  # noqa: E402

L = nat.lib()
dev = "cuda"


def timeit(mk_call, nbuf):
    """Device-bound time per launch: `reps` launches (cycling through nbuf weight buffers so weights stay cold, as in a
    stamp where 1.7 GB of weights stream per UNet evaluation) captured into one CUDA graph and replayed 3 times."""
    reps = max(nbuf, 12)
    for i in range(nbuf):
        mk_call(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for i in range(reps):
                mk_call(i)
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(3):
            graph.replay()
        e1.record(side)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * reps)


def ring(nbytes):
    return max(1, min(24, int(300e6 // nbytes)))


PAIR = 0x1000  # GEMM_BN_PAIR: run as CTA pairs (cta_group::2)
SPLITS = (1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16)


def configs(M, N, K):
    """(BN, split-K) candidates; pair mode doubles the list for problems with at least two 128-row tiles. Split-K only where
    the whole grid is co-resident (the in-kernel reduction) or nearly so."""
    mt = (M + 127) // 128
    for BN in (32, 64, 128, 160, 192, 256, 320):
        if BN > 64 and BN - 32 >= N:
            continue
        if BN == 32 and N > 64:
            continue
        if BN == 320 and (N % 320 or mt < 2):
            continue
        gn = (N + BN - 1) // BN
        for pair in ((PAIR,) if BN == 320 else (0, PAIR) if (mt >= 2 and BN >= 64) else (0,)):
            for sp in SPLITS:
                if sp > 1 and ((K // 64) // sp < 2 or mt * gn * sp > 160):
                    continue
                yield BN | pair, sp


def sweep_linear(M, N, K, flags=0):
    nbuf = ring(N * K * 2)
    A = torch.randn(M, K, device=dev).half()
    Ws = [torch.randn(N, K, device=dev).half() for _ in range(nbuf)]
    out = torch.empty(M, N // 2 if flags & 8 else N, device=dev, dtype=torch.float16)
    bias = torch.randn(N, device=dev)
    res = {}

    def mk(BN, sp):
        def f(i):
            nat.check_op(L.dtp_op_linear(nat.ptr(A), K, K, None, 0, 0, M, nat.ptr(Ws[i % nbuf]), K, N, nat.ptr(bias), None, 0,
                                         nat.ptr(out), out.shape[1], flags, 1.0, 0, BN, sp, nat.stream_ptr()))
        return f
    for BN, sp in configs(M, N, K):
        res[(BN, sp)] = timeit(mk(BN, sp), nbuf)
    auto = timeit(mk(0, 1), nbuf)
    report(f"linear M={M} N={N} K={K} flags={flags}", res, auto, 2.0 * M * N * K)


def sweep_conv(n, H, Wd, cin, cout):
    nbuf = ring(cout * 9 * cin * 2)
    x = torch.randn(n, H, Wd, cin, device=dev).half()
    ws = [torch.randn(cout, 9 * cin, device=dev).half() for _ in range(nbuf)]
    out = torch.empty(n, H, Wd, cout, device=dev, dtype=torch.float16)
    bias = torch.randn(cout, device=dev)
    res = {}

    def mk(BN, sp):
        def f(i):
            nat.check_op(L.dtp_op_conv3x3(nat.ptr(x), cin, None, 0, n, H, Wd, nat.ptr(ws[i % nbuf]), cout, nat.ptr(bias), None, 0,
                                          nat.ptr(out), cout, 0, 1.0, 0, BN, sp, nat.stream_ptr()))
        return f
    for BN, sp in configs(n * H * Wd, cout, 9 * cin):
        res[(BN, sp)] = timeit(mk(BN, sp), nbuf)
    auto = timeit(mk(0, 1), nbuf)
    report(f"conv3x3 n={n} {H}x{Wd} cin={cin} cout={cout} (M={n*H*Wd} K={9*cin})", res, auto,
           2.0 * n * H * Wd * cout * 9 * cin)


def sweep_conv_sc(n, H, Wd, c2, cs, cout):
    """conv2 + fused 1x1 shortcut (dtp_op_conv3x3_shortcut): K = 9 * c2 + cs"""
    nbuf = ring(cout * (9 * c2 + cs) * 2)
    x = torch.randn(n, H, Wd, c2, device=dev).half()
    sx = torch.randn(n, H, Wd, cs, device=dev).half()
    ws = [torch.randn(cout, 9 * c2 + cs, device=dev).half() for _ in range(nbuf)]
    out = torch.empty(n, H, Wd, cout, device=dev, dtype=torch.float16)
    bias = torch.randn(cout, device=dev)
    res = {}

    def mk(BN, sp):
        def f(i):
            nat.check_op(L.dtp_op_conv3x3_shortcut(nat.ptr(x), c2, nat.ptr(sx), cs, None, 0, n, H, Wd, nat.ptr(ws[i % nbuf]), cout,
                                                   nat.ptr(bias), nat.ptr(out), BN, sp, nat.stream_ptr()))
        return f
    for BN, sp in configs(n * H * Wd, cout, 9 * c2 + cs):
        res[(BN, sp)] = timeit(mk(BN, sp), nbuf)
    auto = timeit(mk(0, 1), nbuf)
    report(f"conv3x3sc n={n} {H}x{Wd} c2={c2} cs={cs} cout={cout} (M={n*H*Wd} K={9*c2+cs})", res, auto,
           2.0 * n * H * Wd * cout * (9 * c2 + cs))


CSV = open(os.path.join(ROOT, "gpurun_out", "sweep_all.csv"), "a") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None


def report(name, res, auto, flops):
    if CSV:
        for (bn, sp), v in sorted(res.items()):
            CSV.write(f"{name};{bn};{sp};{v:.2f}\n")
        CSV.write(f"{name};0;0;{auto:.2f}\n")
        CSV.flush()
    best = sorted(res.items(), key=lambda kv: kv[1])[:5]
    s = ", ".join(f"BN={k[0] & 0xfff}{'p' if k[0] & PAIR else ''} sp={k[1]}: {v:.1f}us" for k, v in best)
    print(f"{name}: auto {auto:.1f}us ({flops / auto / 1e6:.0f} TF) | best {s} ({flops / best[0][1] / 1e6:.0f} TF)",
          flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "linear"):
        for shp in [(12288, 320, 320), (12288, 960, 320), (12288, 320, 1280), (12288, 320, 640), (3072, 640, 640),
                    (3072, 1920, 640), (3072, 640, 2560), (3072, 640, 1280), (3072, 320, 2880), (768, 1280, 1280),
                    (768, 3840, 1280), (768, 1280, 5120), (768, 1280, 2560), (768, 640, 5760), (192, 1280, 1280),
                    (192, 1280, 2560), (192, 1280, 5120), (192, 1280, 11520), (192, 3840, 1280)]:
            sweep_linear(*shp)
        for shp in [(12288, 2560, 320), (3072, 5120, 640), (768, 10240, 1280), (192, 10240, 1280)]:
            sweep_linear(*shp, flags=8)
    if which in ("all", "conv"):
        for shp in [(3, 64, 64, 320, 320), (3, 64, 64, 640, 320), (3, 64, 64, 960, 320), (3, 64, 64, 640, 640),
                    (3, 32, 32, 320, 640), (3, 32, 32, 640, 640), (3, 32, 32, 960, 640), (3, 32, 32, 1280, 640),
                    (3, 32, 32, 1920, 640), (3, 32, 32, 1280, 1280), (3, 16, 16, 640, 1280), (3, 16, 16, 1280, 1280),
                    (3, 16, 16, 1920, 1280), (3, 16, 16, 2560, 1280), (3, 8, 8, 1280, 1280), (3, 8, 8, 2560, 1280)]:
            sweep_conv(*shp)
    if which in ("all", "r256"):
        for shp in [(3072, 320, 320), (3072, 960, 320), (3072, 320, 1280), (768, 640, 640), (768, 1920, 640),
                    (768, 640, 2560), (48, 1280, 1280), (48, 1280, 5120)]:
            sweep_linear(*shp)
        for shp in [(3072, 2560, 320), (768, 5120, 640), (48, 10240, 1280)]:
            sweep_linear(*shp, flags=8)
        for shp in [(3, 32, 32, 320, 320), (3, 32, 32, 640, 320), (3, 32, 32, 960, 320), (3, 16, 16, 640, 640),
                    (3, 16, 16, 1280, 640), (3, 16, 16, 960, 640), (3, 4, 4, 1280, 1280), (3, 4, 4, 2560, 1280)]:
            sweep_conv(*shp)
    if which in ("all", "extra"):  # remaining shapes of the 512x512 stamp plan (UNet skip widths, conv_in, VAE levels)
        for shp in [(12288, 320, 960), (3072, 640, 1920), (768, 1280, 1920), (3072, 640, 960), (768, 1280, 640),
                    (3072, 640, 320)]:
            sweep_linear(*shp)
        for shp in [(3, 64, 64, 64, 320), (1, 512, 512, 128, 128), (1, 128, 128, 512, 512), (1, 256, 256, 256, 256),
                    (1, 64, 64, 512, 512), (2, 64, 64, 512, 512), (1, 256, 256, 512, 256), (1, 256, 256, 512, 512),
                    (1, 512, 512, 256, 256)]:
            sweep_conv(*shp)
    if which in ("all", "r2"):  # shapes introduced by the round-2 fusions: ff.net.2 + proj_out (K = 5C), conv2 + conv_shortcut
        for shp in [(12288, 320, 1600), (3072, 640, 3200), (768, 1280, 6400), (192, 1280, 6400)]:
            sweep_linear(*shp)
        for shp in [(3, 64, 64, 320, 960, 320), (3, 64, 64, 320, 640, 320), (3, 32, 32, 640, 320, 640),
                    (3, 32, 32, 640, 1920, 640), (3, 32, 32, 640, 1280, 640), (3, 32, 32, 640, 960, 640),
                    (3, 16, 16, 1280, 640, 1280), (3, 16, 16, 1280, 2560, 1280), (3, 16, 16, 1280, 1920, 1280),
                    (3, 8, 8, 1280, 2560, 1280)]:
            sweep_conv_sc(*shp)
    if which == "splitk":  # every stamp shape with at most 24 row tiles (the split-K candidates), after a change of the reduction
        for shp in [(3072, 640, 640), (3072, 1920, 640), (3072, 640, 2560), (3072, 640, 1280), (3072, 320, 2880), (768, 1280, 1280),
                    (768, 3840, 1280), (768, 1280, 5120), (768, 1280, 2560), (768, 640, 5760), (192, 1280, 1280), (192, 1280, 2560),
                    (192, 1280, 5120), (192, 1280, 11520), (192, 3840, 1280), (3072, 640, 1920), (768, 1280, 1920), (3072, 640, 960),
                    (768, 1280, 640), (3072, 640, 320), (3072, 640, 3200), (768, 1280, 6400), (192, 1280, 6400)]:
            sweep_linear(*shp)
        for shp in [(3072, 5120, 640), (768, 10240, 1280), (192, 10240, 1280)]:
            sweep_linear(*shp, flags=8)
        for shp in [(3, 32, 32, 320, 640), (3, 32, 32, 640, 640), (3, 32, 32, 960, 640), (3, 32, 32, 1280, 640), (3, 32, 32, 1920, 640),
                    (3, 32, 32, 1280, 1280), (3, 16, 16, 640, 1280), (3, 16, 16, 1280, 1280), (3, 16, 16, 1920, 1280),
                    (3, 16, 16, 2560, 1280), (3, 8, 8, 1280, 1280), (3, 8, 8, 2560, 1280)]:
            sweep_conv(*shp)
        for shp in [(3, 32, 32, 640, 320, 640), (3, 32, 32, 640, 1920, 640), (3, 32, 32, 640, 1280, 640), (3, 32, 32, 640, 960, 640),
                    (3, 16, 16, 1280, 640, 1280), (3, 16, 16, 1280, 2560, 1280), (3, 16, 16, 1280, 1920, 1280),
                    (3, 8, 8, 1280, 2560, 1280)]:
            sweep_conv_sc(*shp)
    if which == "dedup":  # two sample groups at level 0 (layers in front of the first cross-attention, dedup_branches)
        for shp in [(8192, 320, 320), (8192, 960, 320)]:
            sweep_linear(*shp)
        sweep_conv(2, 64, 64, 320, 320)
    if which in ("all", "vae"):
        for shp in [(2, 512, 512, 128, 128), (2, 256, 256, 256, 256), (2, 128, 128, 512, 512), (1, 512, 512, 256, 128)]:
            sweep_conv(*shp)
