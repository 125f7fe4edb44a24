"""Launch one contraction shape repeatedly (cold weights: ring of buffers > L2) — target for `ncu --set full`.
    python profiles/one_gemm.py M N K BN splits [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusiontexturepainting_b200 import _native as nat  # noqa: E402

L = nat.lib()
M, N, K, BN, sp = (int(x) for x in sys.argv[1:6])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 20
nbuf = max(2, int(400e6 // (N * K * 2)))
A = torch.randn(M, K, device="cuda").half()
Ws = [torch.randn(N, K, device="cuda").half() for _ in range(nbuf)]
out = torch.empty(M, N, device="cuda", dtype=torch.float16)
for i in range(reps):
    nat.check_op(L.dtp_op_linear(nat.ptr(A), K, K, None, 0, 0, M, nat.ptr(Ws[i % nbuf]), K, N, None, None, 0, nat.ptr(out),
                                 N, 0, 1.0, 0, BN, sp, nat.stream_ptr()))
torch.cuda.synchronize()
print("done")
