"""Parity of every CUDA kernel family (called through the C-ABI operator entry points of include/dtp.h) against a
plain PyTorch fp32 evaluation of the same op on the same fp16-rounded inputs.

Tolerances: single conv / GEMM (fp16 in, fp32 accumulate, fp16 out) rel-L2 <= 2e-3; norms <= 2e-3;
attention <= 3e-3; element-wise scheduler / canvas kernels bit-exact in fp32."""
import math

import pytest
import torch
import torch.nn.functional as F

from diffusiontexturepainting_b200 import _native as nat

pytestmark = pytest.mark.gpu

DEV = "cuda"


def rel_l2(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def h(x):
    return x.half().contiguous()


def linear_op(A, W, bias=None, residual=None, flags=0, alpha=1.0, BN=0, splits=1, A1=None, out_f32=False):
    L = nat.lib()
    M, K0 = A.shape
    K1 = A1.shape[1] if A1 is not None else 0
    N = W.shape[0]
    n_out = N // 2 if flags & 8 else N
    out = torch.empty(M, n_out, device=DEV, dtype=torch.float32 if out_f32 else torch.float16)
    rc = L.dtp_op_linear(nat.ptr(A), A.stride(0), K0, nat.ptr(A1), A1.stride(0) if A1 is not None else 0, K1, M,
                         nat.ptr(W), W.stride(0), N, nat.ptr(bias), nat.ptr(residual),
                         residual.stride(0) if residual is not None else 0, nat.ptr(out), n_out, flags, alpha, 0, BN,
                         splits, nat.stream_ptr())
    nat.check_op(rc, "dtp_op_linear")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("BN", [0, 32, 64, 128, 160, 192, 256])
@pytest.mark.parametrize("M,N,K", [(300, 200, 320), (128, 256, 64), (1000, 520, 1280), (7, 40, 72)])
def test_linear_plain(M, N, K, BN):
    A, W, b = h(rnd(M, K)), h(rnd(N, K, scale=K ** -0.5)), rnd(N)
    out = linear_op(A, W, bias=b, BN=BN)
    ref = A.float() @ W.float().t() + b
    assert rel_l2(out, ref) < 2e-3


@pytest.mark.parametrize("splits", [2, 5])
def test_linear_splitk_residual(splits):
    M, N, K = 192, 1280, 2560
    A, W, b, r = h(rnd(M, K)), h(rnd(N, K, scale=K ** -0.5)), rnd(N), h(rnd(M, N))
    out = linear_op(A, W, bias=b, residual=r, BN=128, splits=splits)
    ref = A.float() @ W.float().t() + b + r.float()
    assert rel_l2(out, ref) < 2e-3


def test_linear_dual_source():
    M, N, K0, K1 = 500, 320, 640, 320
    A0, A1, W = h(rnd(M, K0)), h(rnd(M, K1, seed=3)), h(rnd(N, K0 + K1, scale=(K0 + K1) ** -0.5))
    out = linear_op(A0, W, A1=A1)
    ref = torch.cat([A0, A1], 1).float() @ W.float().t()
    assert rel_l2(out, ref) < 2e-3


@pytest.mark.parametrize("flag,fn", [(1, lambda x: F.gelu(x)), (2, lambda x: x * torch.sigmoid(1.702 * x)),
                                     (4, lambda x: F.silu(x))])
def test_linear_activations(flag, fn):
    M, N, K = 260, 384, 256
    A, W, b = h(rnd(M, K)), h(rnd(N, K, scale=K ** -0.5)), rnd(N)
    out = linear_op(A, W, bias=b, flags=flag)
    assert rel_l2(out, fn(A.float() @ W.float().t() + b)) < 2e-3


@pytest.mark.parametrize("splits", [1, 3])
def test_linear_geglu(splits):
    # weight rows interleaved per 32-column chunk: 16 value rows then their 16 gate rows
    M, C = 300, 320
    K = C
    A = h(rnd(M, K))
    Wfull, bfull = rnd(8 * C, K, scale=K ** -0.5), rnd(8 * C)
    ref_h = A.float() @ h(Wfull).float().t() + bfull
    ref = ref_h[:, :4 * C] * F.gelu(ref_h[:, 4 * C:])
    idx = torch.arange(4 * C, device=DEV).view(-1, 16)
    perm = torch.cat([idx, idx + 4 * C], dim=1).reshape(-1)
    out = linear_op(A, h(Wfull[perm]), bias=bfull[perm].contiguous(), flags=8, BN=128, splits=splits)
    assert out.shape == (M, 4 * C)
    assert rel_l2(out, ref) < 2e-3


def test_linear_f32_out_and_alpha():
    M, N, K = 130, 96, 128
    A, W = h(rnd(M, K)), h(rnd(N, K, scale=K ** -0.5))
    out = linear_op(A, W, flags=128, alpha=0.25, out_f32=True)
    assert rel_l2(out, 0.25 * (A.float() @ W.float().t())) < 1e-5


def conv_ref(x_nhwc, w_packed, cin, bias=None):
    cout = w_packed.shape[0]
    w = w_packed.float().view(cout, 3, 3, cin).permute(0, 3, 1, 2)
    y = F.conv2d(x_nhwc.float().permute(0, 3, 1, 2), w, bias, padding=1)
    return y.permute(0, 2, 3, 1).contiguous()


def conv_op(x0, w, bias=None, x1=None, residual=None, flags=0, BN=0, splits=1, cout=None, f32_nchw=False):
    L = nat.lib()
    n, H, W_, c0 = x0.shape
    c1 = x1.shape[3] if x1 is not None else 0
    cout = w.shape[0]
    if f32_nchw:
        out = torch.empty(n, cout, H, W_, device=DEV, dtype=torch.float32)
        flags |= 32
    else:
        out = torch.empty(n, H, W_, cout, device=DEV, dtype=torch.float16)
    rc = L.dtp_op_conv3x3(nat.ptr(x0), c0, nat.ptr(x1), c1, n, H, W_, nat.ptr(w), cout, nat.ptr(bias),
                          nat.ptr(residual), cout, nat.ptr(out), cout, flags, 1.0, H * W_, BN, splits,
                          nat.stream_ptr())
    nat.check_op(rc, "dtp_op_conv3x3")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("n,H,W,cin,cout", [(3, 16, 16, 64, 128), (3, 8, 8, 128, 64), (2, 64, 64, 64, 320),
                                            (1, 256, 256, 64, 32), (3, 4, 4, 320, 320), (5, 1, 1, 128, 128),
                                            (1, 48, 48, 64, 64), (3, 2, 2, 64, 64)])
def test_conv3x3(n, H, W, cin, cout):
    x = h(rnd(n, H, W, cin))
    w = h(rnd(cout, 9 * cin, scale=(9 * cin) ** -0.5))
    b = rnd(cout)
    out = conv_op(x, w, bias=b)
    assert rel_l2(out, conv_ref(x, w, cin, b)) < 2e-3


PAIR = 0x1000  # GEMM_BN_PAIR: CTA-pair (cta_group::2) mode, selected per op through the BN argument


@pytest.mark.parametrize("splits,BN", [(1, 128), (4, 128), (3, 64), (2, 256), (1, 160), (2, 192), (1, 128 | PAIR),
                                       (2, 256 | PAIR), (3, 64 | PAIR), (5, 160 | PAIR)])
def test_conv3x3_dual_source_residual_splitk(splits, BN):
    n, H, W, c0, c1, cout = 3, 8, 8, 128, 64, 256
    x0, x1 = h(rnd(n, H, W, c0)), h(rnd(n, H, W, c1, seed=5))
    w = h(rnd(cout, 9 * (c0 + c1), scale=(9 * (c0 + c1)) ** -0.5))
    b, r = rnd(cout), h(rnd(n, H, W, cout, seed=9))
    out = conv_op(x0, w, bias=b, x1=x1, residual=r, BN=BN, splits=splits)
    ref = conv_ref(torch.cat([x0, x1], 3), w, c0 + c1, b) + r.float()
    assert rel_l2(out, ref) < 2e-3


@pytest.mark.parametrize("n,H,W,cin,cout,splits", [(3, 16, 16, 128, 320, 1), (3, 32, 32, 64, 640, 1), (3, 8, 8, 256, 640, 4),
                                                   (1, 64, 64, 64, 320, 1), (5, 16, 16, 64, 320, 1)])
def test_conv3x3_wide_pair_tile(n, H, W, cin, cout, splits):
    """BN = 320 CTA-pair tiles (two N = 160 MMAs per k-step into one 320-column accumulator), odd m-tile counts included."""
    x = h(rnd(n, H, W, cin))
    w = h(rnd(cout, 9 * cin, scale=(9 * cin) ** -0.5))
    b, r = rnd(cout), h(rnd(n, H, W, cout, seed=9))
    out = conv_op(x, w, bias=b, residual=r, BN=320 | PAIR, splits=splits)
    assert rel_l2(out, conv_ref(x, w, cin, b) + r.float()) < 2e-3


@pytest.mark.parametrize("n,H,W,c2,cs0,cs1,cout,BN,splits", [(3, 16, 16, 128, 64, 0, 128, 0, 1), (3, 8, 8, 256, 128, 128, 256, 128, 3),
                                                              (3, 32, 32, 64, 64, 128, 320, 320 | PAIR, 1), (2, 16, 16, 128, 192, 0, 128, 64, 2),
                                                              (3, 64, 64, 320, 320, 320, 320, 0, 1)])
def test_conv3x3_with_fused_shortcut(n, H, W, c2, cs0, cs1, cout, BN, splits):
    """conv2 + conv_shortcut of a ResnetBlock2D as ONE contraction (shortcut = extra k-blocks at the centre tap)"""
    L = nat.lib()
    x = h(rnd(n, H, W, c2))
    s0 = h(rnd(n, H, W, cs0, seed=1))
    s1 = h(rnd(n, H, W, cs1, seed=2)) if cs1 else None
    w2 = h(rnd(cout, 9 * c2, scale=(9 * c2) ** -0.5, seed=3))
    ws = h(rnd(cout, cs0 + cs1, scale=(cs0 + cs1) ** -0.5, seed=4))
    wt = torch.cat([w2, ws], 1).contiguous()
    bias = rnd(cout, seed=5)
    out = torch.empty(n, H, W, cout, device=DEV, dtype=torch.float16)
    rc = L.dtp_op_conv3x3_shortcut(nat.ptr(x), c2, nat.ptr(s0), cs0, nat.ptr(s1), cs1, n, H, W, nat.ptr(wt), cout,
                                   nat.ptr(bias), nat.ptr(out), BN, splits, nat.stream_ptr())
    nat.check_op(rc, "conv3x3_shortcut")
    torch.cuda.synchronize()
    s = torch.cat([s0, s1], 3) if cs1 else s0
    ref = conv_ref(x, w2, c2) + s.float() @ ws.float().t() + bias
    assert rel_l2(out, ref) < 2e-3


@pytest.mark.parametrize("n,H,W,cin,cout,BN", [(3, 8, 8, 64, 128, 0), (3, 16, 16, 128, 320, 320 | PAIR), (2, 32, 32, 64, 192, 128),
                                               (1, 16, 16, 192, 64, 64), (3, 32, 32, 640, 640, 0), (5, 8, 8, 128, 320, 256 | PAIR),
                                               (1, 64, 64, 512, 512, 0)])
def test_upsample_conv_folded(n, H, W, cin, cout, BN):
    """Upsample2D = nearest 2x + conv3x3 as ONE contraction over the half-resolution input: four parity classes of 2x2 taps
    with pre-summed weights (reference graph: resize folded, models.py:128-186)"""
    L = nat.lib()
    x = h(rnd(n, H, W, cin))
    w = h(rnd(cout, 9 * cin, scale=(9 * cin) ** -0.5, seed=3))
    bias = rnd(cout, seed=5)
    wst = torch.empty(4 * cout, 4 * cin, device=DEV, dtype=torch.float16)
    out = torch.full((n, 2 * H, 2 * W, cout), float("nan"), device=DEV, dtype=torch.float16)
    rc = L.dtp_op_upconv2x(nat.ptr(x), cin, n, H, W, nat.ptr(w), cout, nat.ptr(bias), nat.ptr(wst), nat.ptr(out), BN,
                           nat.stream_ptr())
    nat.check_op(rc, "upconv2x")
    torch.cuda.synchronize()
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    ref = conv_ref(up, w, cin, bias)
    assert rel_l2(out, ref) < 2e-3
    # against the two-kernel path of the same library (upsample kernel + 3x3 contraction): same fp16 inputs, fp32 accumulate
    up16 = torch.empty(n, 2 * H, 2 * W, cin, device=DEV, dtype=torch.float16)
    nat.check_op(L.dtp_op_upsample2x(nat.ptr(x), n, H, W, cin, nat.ptr(up16), nat.stream_ptr()), "upsample2x")
    two = conv_op(up16, w, bias=bias)
    assert rel_l2(out, two) < 1.5e-3


@pytest.mark.parametrize("n,H,W,cin,cout,pad_lo,BN,splits", [(3, 16, 16, 64, 128, 1, 0, 1), (3, 64, 64, 320, 320, 1, 0, 1),
                                                             (2, 32, 32, 128, 128, 0, 128, 1), (3, 8, 8, 256, 640, 1, 320 | PAIR, 4),
                                                             (1, 64, 64, 128, 128, 0, 64, 2), (5, 4, 4, 128, 192, 1, 0, 3)])
def test_conv3x3_stride2_from_parity_views(n, H, W, cin, cout, pad_lo, BN, splits):
    """Downsample2D without an im2col buffer: tap (ky, kx) is a box of the (row parity, column parity) view of the input"""
    L = nat.lib()
    x = h(rnd(n, H, W, cin))
    w = h(rnd(cout, 9 * cin, scale=(9 * cin) ** -0.5, seed=3))
    bias = rnd(cout, seed=5)
    out = torch.full((n, H // 2, W // 2, cout), float("nan"), device=DEV, dtype=torch.float16)
    rc = L.dtp_op_conv3x3_s2(nat.ptr(x), cin, n, H, W, nat.ptr(w), cout, nat.ptr(bias), pad_lo, nat.ptr(out), BN, splits,
                             nat.stream_ptr())
    nat.check_op(rc, "conv3x3_s2")
    torch.cuda.synchronize()
    xn = x.float().permute(0, 3, 1, 2)
    wn = w.float().view(cout, 3, 3, cin).permute(0, 3, 1, 2)
    if pad_lo:
        ref = F.conv2d(xn, wn, bias, stride=2, padding=1)
    else:
        ref = F.conv2d(F.pad(xn, (0, 1, 0, 1)), wn, bias, stride=2)
    assert rel_l2(out, ref.permute(0, 2, 3, 1)) < 2e-3


def test_conv3x3_small_cout_f32_nchw():
    n, H, W, cin, cout = 3, 16, 16, 320, 4
    x, w, b = h(rnd(n, H, W, cin)), h(rnd(cout, 9 * cin, scale=(9 * cin) ** -0.5)), rnd(cout)
    out = conv_op(x, w, bias=b, f32_nchw=True)
    ref = conv_ref(x, w, cin, b).permute(0, 3, 1, 2)
    assert rel_l2(out, ref) < 1e-3


@pytest.mark.parametrize("seq,heads,d", [(256, 8, 40), (200, 4, 80), (128, 2, 160), (64, 1, 512)])
def test_bmm_attention_pair(seq, heads, d):
    """scores = Q K^T / sqrt(d) (K-major B) then out = P V with V consumed MN-major."""
    L = nat.lib()
    B, C = 3, heads * d
    qkv = h(rnd(B, seq, 3 * C))
    S = torch.empty(B, heads, seq, seq, device=DEV, dtype=torch.float16)
    scale = 1.0 / math.sqrt(d)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    rc = L.dtp_op_bmm(nat.ptr(q), 3 * C, d, seq * 3 * C, nat.ptr(k), 3 * C, d, seq * 3 * C, 0, seq, seq, d, heads, B,
                      nat.ptr(S), seq, seq * seq, heads * seq * seq, scale, 0, 0, nat.stream_ptr())
    nat.check_op(rc, "bmm qk")
    torch.cuda.synchronize()
    qh = q.float().view(B, seq, heads, d).permute(0, 2, 1, 3)
    kh = k.float().view(B, seq, heads, d).permute(0, 2, 1, 3)
    vh = v.float().view(B, seq, heads, d).permute(0, 2, 1, 3)
    Sref = qh @ kh.transpose(-1, -2) * scale
    assert rel_l2(S, Sref) < 2e-3
    P = torch.softmax(Sref, -1).half().contiguous()
    O = torch.empty(B, seq, C, device=DEV, dtype=torch.float16)
    rc = L.dtp_op_bmm(nat.ptr(P), seq, seq * seq, heads * seq * seq, nat.ptr(v), 3 * C, d, seq * 3 * C, 1, seq, d, seq,
                      heads, B, nat.ptr(O), C, d, seq * C, 1.0, 0, 0, nat.stream_ptr())
    nat.check_op(rc, "bmm pv")
    torch.cuda.synchronize()
    Oref = (P.float() @ vh).permute(0, 2, 1, 3).reshape(B, seq, C)
    assert rel_l2(O, Oref) < 2e-3


@pytest.mark.parametrize("seq,heads,d,batch", [(4096, 8, 40, 3), (1024, 8, 80, 3), (256, 8, 160, 3), (200, 4, 40, 2),
                                                  (129, 2, 80, 1), (128, 1, 64, 1), (1000, 3, 192, 1), (384, 4, 16, 2),
                                                  (2304, 8, 40, 1)])
def test_flash_attention(seq, heads, d, batch):
    """tcgen05 flash attention vs fp32 softmax(QK^T/sqrt(d))V on the same fp16 inputs; tolerance 3e-3 rel-L2."""
    L = nat.lib()
    C = heads * d
    qkv = h(rnd(batch, seq, 3 * C))
    out = torch.zeros(batch, seq, C, device=DEV, dtype=torch.float16)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    nat.check_op(L.dtp_op_flash_attn(nat.ptr(q), nat.ptr(k), nat.ptr(v), 3 * C, seq * 3 * C, nat.ptr(out), C, seq * C,
                                     seq, heads, d, batch, nat.stream_ptr()), "flash_attn")
    torch.cuda.synchronize()
    qh = q.float().view(batch, seq, heads, d).transpose(1, 2)
    kh = k.float().view(batch, seq, heads, d).transpose(1, 2)
    vh = v.float().view(batch, seq, heads, d).transpose(1, 2)
    ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(batch, seq, C)
    assert rel_l2(out, ref) < 3e-3
    # peaked scores exercise the lazy rescaling path (running max grows block after block)
    if seq >= 256:
        qkv2 = qkv.clone()
        qkv2[..., :C] = (qkv2[..., :C].float() * 6).half()
        q, k, v = qkv2[..., :C], qkv2[..., C:2 * C], qkv2[..., 2 * C:]
        nat.check_op(L.dtp_op_flash_attn(nat.ptr(q), nat.ptr(k), nat.ptr(v), 3 * C, seq * 3 * C, nat.ptr(out), C,
                                         seq * C, seq, heads, d, batch, nat.stream_ptr()), "flash_attn")
        torch.cuda.synchronize()
        qh = q.float().view(batch, seq, heads, d).transpose(1, 2)
        ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(batch, seq, C)
        assert rel_l2(out, ref) < 3e-3


@pytest.mark.parametrize("n,hw,c0,c1,silu", [(3, 256, 320, 0, 1), (3, 64, 1280, 640, 1), (2, 1024, 128, 0, 0),
                                             (3, 16, 640, 320, 1), (1, 4096, 512, 0, 1), (3, 1, 1280, 1280, 1),
                                             (3, 4096, 320, 0, 1), (3, 4096, 640, 320, 1), (2, 1024, 1280, 640, 0),
                                             (2, 16384, 128, 0, 1)])
def test_groupnorm(n, hw, c0, c1, silu):
    _groupnorm_case(n, hw, c0, c1, silu)


@pytest.mark.parametrize("n,hw,c,silu", [(1, 262144, 128, 1), (2, 262144, 128, 1), (1, 262144, 256, 1), (2, 65536, 256, 1),
                                         (2, 16384, 512, 0), (3, 4096, 320, 1), (12, 1024, 320, 1)])
def test_groupnorm_vae_scale(n, hw, c, silu):
    """VAE GroupNorms of the 512 x 512 stamp: rows = n * hw = 262 144 / 524 288 (C2), plus the C3-per-GPU UNet shape."""
    _groupnorm_case(n, hw, c, 0, silu)


def _groupnorm_case(n, hw, c0, c1, silu):
    L = nat.lib()
    x0 = h(rnd(n, hw, c0) * 2 + 0.5)
    x1 = h(rnd(n, hw, c1, seed=2)) if c1 else None
    C = c0 + c1
    g, b = rnd(C) * 0.2 + 1, rnd(C, seed=4) * 0.2
    out = torch.empty(n, hw, C, device=DEV, dtype=torch.float16)
    rc = L.dtp_op_groupnorm(nat.ptr(x0), c0, nat.ptr(x1), c1, n, hw, 32, nat.ptr(g), nat.ptr(b), 1e-5, silu,
                            nat.ptr(out), nat.stream_ptr())
    nat.check_op(rc, "groupnorm")
    torch.cuda.synchronize()
    x = torch.cat([x0, x1], 2) if c1 else x0
    ref = F.group_norm(x.float().permute(0, 2, 1), 32, g, b, 1e-5).permute(0, 2, 1)
    if silu:
        ref = F.silu(ref)
    assert rel_l2(out, ref) < 2e-3


@pytest.mark.parametrize("rows,C", [(1000, 320), (77, 1280), (14, 768), (4096, 640)])
def test_layernorm(rows, C):
    L = nat.lib()
    x = h(rnd(rows, C) * 3 + 1)
    g, b = rnd(C) * 0.2 + 1, rnd(C, seed=4) * 0.2
    out = torch.empty_like(x)
    nat.check_op(L.dtp_op_layernorm(nat.ptr(x), rows, C, nat.ptr(g), nat.ptr(b), 1e-5, nat.ptr(out), nat.stream_ptr()))
    torch.cuda.synchronize()
    assert rel_l2(out, F.layer_norm(x.float(), (C,), g, b, 1e-5)) < 2e-3


@pytest.mark.parametrize("rows,cols", [(333, 4096), (50, 50), (9, 14), (1024, 1024)])
def test_softmax(rows, cols):
    L = nat.lib()
    x = h(rnd(rows, cols) * 4)
    ref = torch.softmax(x.float(), -1)
    nat.check_op(L.dtp_op_softmax(nat.ptr(x), rows, cols, cols, nat.stream_ptr()))
    torch.cuda.synchronize()
    assert rel_l2(x, ref) < 2e-3


@pytest.mark.parametrize("nq,nkv,heads,d,batch", [(1024, 14, 8, 40, 3), (50, 50, 12, 64, 14), (9, 9, 4, 192, 1),
                                                  (64, 64, 8, 160, 3), (16, 16, 8, 80, 6), (1, 1, 4, 192, 1)])
def test_attn_small(nq, nkv, heads, d, batch):
    L = nat.lib()
    C = heads * d
    q = h(rnd(batch, nq, C))
    kvb = 2 if nkv == 14 else batch
    k, v = h(rnd(kvb, nkv, C, seed=1)), h(rnd(kvb, nkv, C, seed=2))
    idx = torch.tensor([0 if i < batch // 3 else 1 for i in range(batch)], device=DEV, dtype=torch.int32) \
        if nkv == 14 else None
    out = torch.empty_like(q)
    scale = d ** -0.5
    nat.check_op(L.dtp_op_attn_small(nat.ptr(q), C, nat.ptr(k), C, nat.ptr(v), C, nat.ptr(out), C, nq, nkv, heads, d,
                                     batch, nq * C, nkv * C, nq * C, nat.ptr(idx), scale, nat.stream_ptr()))
    torch.cuda.synchronize()
    kk = k[idx.long()] if idx is not None else k
    vv = v[idx.long()] if idx is not None else v
    qh = q.float().view(batch, nq, heads, d).transpose(1, 2)
    kh = kk.float().view(batch, nkv, heads, d).transpose(1, 2)
    vh = vv.float().view(batch, nkv, heads, d).transpose(1, 2)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * scale, -1) @ vh).transpose(1, 2).reshape(batch, nq, C)
    assert rel_l2(out, ref) < 3e-3


def test_upsample_and_im2col():
    L = nat.lib()
    n, H, W, C = 2, 6, 10, 64
    x = h(rnd(n, H, W, C))
    up = torch.empty(n, 2 * H, 2 * W, C, device=DEV, dtype=torch.float16)
    nat.check_op(L.dtp_op_upsample2x(nat.ptr(x), n, H, W, C, nat.ptr(up), nat.stream_ptr()))
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert torch.equal(up.float(), ref)
    for pad_lo, Hin in ((1, 8), (0, 8)):
        x = h(rnd(n, Hin, Hin, C))
        Ho = Hin // 2
        col = torch.empty(n, Ho, Ho, 9 * C, device=DEV, dtype=torch.float16)
        nat.check_op(L.dtp_op_im2col_s2(nat.ptr(x), n, Hin, Hin, C, pad_lo, Ho, Ho, nat.ptr(col), nat.stream_ptr()))
        torch.cuda.synchronize()
        xp = x.float().permute(0, 3, 1, 2)
        xp = F.pad(xp, (1, 1, 1, 1)) if pad_lo else F.pad(xp, (0, 1, 0, 1))
        un = F.unfold(xp, 3, stride=2).view(n, C, 9, Ho, Ho).permute(0, 3, 4, 2, 1).reshape(n, Ho, Ho, 9 * C)
        assert torch.equal(col.float(), un)


def test_ddim_step_bit_exact():
    L = nat.lib()
    B, chw = 2, 4 * 32 * 32
    eps3, lat = rnd(3 * B, chw), rnd(B, chw, seed=7)
    out = torch.empty_like(lat)
    cfg, tg, a_t, a_p = 2.0, 1.0, 0.27499884366989136, 0.3455579876899719
    nat.check_op(L.dtp_op_ddim_step(nat.ptr(eps3), nat.ptr(lat), nat.ptr(out), B, chw, cfg, tg, a_t, a_p,
                                    nat.stream_ptr()))
    torch.cuda.synchronize()
    # reference arithmetic restated with torch CPU fp32 tensors in the reference's operation order
    e3, x = eps3.cpu(), lat.cpu()
    eu, ec, et = e3.chunk(3)
    e = eu + cfg * (ec - eu) + tg * (et - ec)
    at, ap = torch.tensor(a_t, dtype=torch.float32), torch.tensor(a_p, dtype=torch.float32)
    x0 = (x - (1 - at) ** 0.5 * e) / at ** 0.5
    ref = ap ** 0.5 * x0 + (1 - ap) ** 0.5 * e
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("B,R,pad", [(2, 64, 21), (1, 256, 150), (2, 512, 150), (1, 512, 255), (1, 128, 1), (1, 64, 200)])
def test_canvas_preprocess_and_composite(B, R, pad):
    """pad = 150 at 256 / 512 is the server operating point (manager.py:104-110); 255 is the wire maximum (uint8 header
    field, server_io.py:105-119); pad > R covers windows wider than the image."""
    L = nat.lib()
    canvas = torch.rand(B, 4, R, R, generator=torch.Generator().manual_seed(1))
    canvas[:, 3] = (canvas[:, 3] > 0.97).float()
    canvas[B - 1, 3, :20] = 1.0
    brush = torch.rand(1, 3, R, R, generator=torch.Generator().manual_seed(2))
    cd, bd = canvas.to(DEV), brush.to(DEV)
    mi, m, ci, cm = (torch.empty(B, c, R, R, device=DEV) for c in (3, 1, 3, 1))
    scratch = torch.empty(B, R, R, device=DEV)
    nat.check_op(L.dtp_op_canvas_preprocess(nat.ptr(cd), nat.ptr(bd), B, R, pad, nat.ptr(mi), nat.ptr(m), nat.ptr(ci),
                                            nat.ptr(cm), nat.ptr(scratch), nat.stream_ptr()))
    torch.cuda.synchronize()
    images = canvas[:, :3] * 2 - 1.0
    masks = canvas[:, 3:]
    masked = images * masks
    lo, hi = pad // 2, pad - pad // 2 - 1
    # flat window maximum == row maximum of column maxima (exact), evaluated on the device to keep pad = 150 @ 512 fast
    dil = F.max_pool2d(F.pad(masks.to(DEV), (lo, hi, 0, 0), value=-1e4), (1, pad), stride=1)
    dil = F.max_pool2d(F.pad(dil, (0, 0, lo, hi), value=-1e4), (pad, 1), stride=1).cpu()
    hint = 1 - dil
    ctx_img = masked + (brush * 2 - 1) * hint
    ctx_mask = torch.clamp(masks + hint, min=0, max=1)
    assert torch.equal(mi.cpu(), masked) and torch.equal(m.cpu(), 1 - masks)
    assert torch.equal(ci.cpu(), ctx_img) and torch.equal(cm.cpu(), 1 - ctx_mask)
    raw = torch.rand(B, 3, R, R, generator=torch.Generator().manual_seed(3)).to(DEV)
    of, ou = torch.empty(B, 3, R, R, device=DEV), torch.empty(B, R, R, 3, device=DEV, dtype=torch.uint8)
    nat.check_op(L.dtp_op_composite(nat.ptr(cd), nat.ptr(raw), B, R, nat.ptr(of), nat.ptr(ou), nat.stream_ptr()))
    torch.cuda.synchronize()
    ref = canvas[:, :3] * masks + raw.cpu() * (1 - masks)
    assert torch.equal(of.cpu(), ref)
    assert torch.equal(ou.cpu(), (ref * 255).to(torch.uint8).permute(0, 2, 3, 1))
