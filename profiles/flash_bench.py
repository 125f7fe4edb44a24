"""Self-attention kernel alone at the UNet shapes: CUDA-event time per launch (graph-free, back to back).
    python profiles/flash_bench.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from diffusiontexturepainting_b200 import _native as nat  # noqa: E402

L = nat.lib()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for (seq, heads, d, batch) in [(4096, 8, 40, 3), (1024, 8, 80, 3), (256, 8, 160, 3)]:
    C = heads * d
    qkv = torch.randn(batch, seq, 3 * C, device="cuda").half()
    out = torch.zeros(batch, seq, C, device="cuda", dtype=torch.float16)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]

    def call():
        nat.check_op(L.dtp_op_flash_attn(nat.ptr(q), nat.ptr(k), nat.ptr(v), 3 * C, seq * 3 * C, nat.ptr(out), C, seq * C,
                                         seq, heads, d, batch, nat.stream_ptr()), "flash_attn")
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
    fl = 4.0 * seq * seq * d * heads * batch
    print(f"flash seq={seq} heads={heads} d={d} batch={batch}: {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s (algorithmic)", flush=True)
