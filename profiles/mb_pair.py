"""Single-CTA vs CTA-pair contraction on the UNet conv-like shapes, with per-CTA stage timers (run twice: DTP_CLUSTER=0/1)."""
import sys
sys.path.insert(0, 'profiles')
import gemm_microbench as g
for (M, N, K, BN, sp) in [(3072, 640, 5760, 256, 1), (3072, 640, 5760, 256, 2), (3072, 640, 5760, 128, 1), (768, 1280, 11520, 256, 4),
                          (768, 1280, 11520, 256, 5), (192, 1280, 11520, 256, 8), (192, 1280, 11520, 256, 14), (192, 1280, 11520, 128, 7),
                          (12288, 320, 2880, 160, 1), (12288, 640, 5760, 256, 1)]:
    g.run_linear(M, N, K, BN, splits=sp, stages=True)
