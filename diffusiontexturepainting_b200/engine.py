"""Python handle around the C-ABI engine (include/dtp.h). Device memory, streams and the RNG are PyTorch's; all arithmetic
of the stamp path runs in libdtp_sm100.so. There is no fallback: a missing library or a non-Blackwell device raises."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _native as nat
from . import weights as W


class DtpConfig(C.Structure):
    _fields_ = [
        ("unet_in_channels", C.c_int), ("unet_out_channels", C.c_int),
        ("unet_block_out", C.c_int * 4), ("unet_down_attn", C.c_int * 4),
        ("unet_layers_per_block", C.c_int), ("unet_heads", C.c_int), ("unet_cross_dim", C.c_int),
        ("groups", C.c_int),
        ("vae_block_out", C.c_int * 4), ("vae_layers_per_block", C.c_int), ("vae_latent", C.c_int),
        ("enc_width", C.c_int), ("enc_layers", C.c_int), ("enc_heads", C.c_int), ("enc_mlp", C.c_int),
        ("enc_tower_layers", C.c_int), ("enc_tower_heads", C.c_int), ("enc_cross_dim", C.c_int),
        ("enc_tokens", C.c_int),
        ("arena_bytes", C.c_ulonglong),
    ]


def _c_config(cfg: W.ModelConfig, arena_bytes: int) -> DtpConfig:
    c = DtpConfig()
    u, v, e = cfg.unet, cfg.vae, cfg.enc
    c.unet_in_channels, c.unet_out_channels = u.in_channels, u.out_channels
    c.unet_block_out = (C.c_int * 4)(*u.block_out_channels)
    c.unet_down_attn = (C.c_int * 4)(*[int(x) for x in u.down_attention])
    c.unet_layers_per_block, c.unet_heads, c.unet_cross_dim = u.layers_per_block, u.heads, u.cross_dim
    c.groups = u.groups
    c.vae_block_out = (C.c_int * 4)(*v.block_out_channels)
    c.vae_layers_per_block, c.vae_latent = v.layers_per_block, v.latent_channels
    c.enc_width, c.enc_layers, c.enc_heads, c.enc_mlp = e.width, e.layers, e.heads, e.mlp
    c.enc_tower_layers, c.enc_tower_heads, c.enc_cross_dim = e.tower_layers, e.tower_heads, e.cross_dim
    c.enc_tokens = sum(e.num_patches)
    c.arena_bytes = int(arena_bytes)
    return c


class Engine:
    """One engine per process / GPU. Tensors handed to the methods must be contiguous CUDA tensors on `device`."""

    def __init__(self, cfg: W.ModelConfig, device: int = 0, arena_bytes: int = 0):
        if not torch.cuda.is_available():
            raise nat.DtpError("no CUDA device: the stamp path has no CPU fallback")
        self.cfg = cfg
        self.device = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        torch.cuda.set_device(self.device)
        self._lib = nat.lib()
        self._h = C.c_void_p()
        cc = _c_config(cfg, arena_bytes)
        rc = self._lib.dtp_create(C.byref(cc), C.byref(self._h))
        if rc != 0:
            raise nat.DtpError(f"dtp_create failed ({rc}): {self._lib.dtp_last_error(None).decode()}")
        self.tokens = sum(cfg.enc.num_patches)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.dtp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise nat.DtpError(f"{what} failed ({rc}): {self._lib.dtp_last_error(self._h).decode()}")

    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ------------------------------------------------------------------ weights
    def load_packed(self, prefix: str, packed: Dict[str, torch.Tensor]):
        for name, t in packed.items():
            t = t.detach().contiguous()  # host or device memory: the engine copies with cudaMemcpyDefault
            if t.dtype == torch.float16:
                dt = 1
            elif t.dtype == torch.float32:
                dt = 0
            else:
                raise nat.DtpError(f"tensor {name}: unsupported dtype {t.dtype}")
            shape = (C.c_longlong * max(t.dim(), 1))(*(list(t.shape) or [1]))
            self._check(self._lib.dtp_set_tensor(self._h, (prefix + name).encode(), C.c_void_p(t.data_ptr()), shape,
                                                 max(t.dim(), 1), dt), f"dtp_set_tensor({prefix}{name})")

    def load_state_dicts(self, unet_sd, vae_sd, enc_sd, merge_lora: bool = True):
        """diffusers / openai-CLIP keyed fp32 state dicts -> LoRA merge (models.py:1046-1093) -> packed upload."""
        if merge_lora:
            unet_sd = W.merge_lora(unet_sd)
        self.load_packed("unet.", W.pack_unet(unet_sd))
        self.load_packed("vae.", W.pack_vae(vae_sd))
        enc = W.pack_encoder(enc_sd)
        enc["pos_emb"] = W.patch_pos_emb(self.cfg.enc.width, self.cfg.enc.num_patches).reshape(-1, self.cfg.enc.width)
        self.load_packed("enc.", enc)
        self._check(self._lib.dtp_finalize_weights(self._h), "dtp_finalize_weights")

    def load_prepacked(self, packed: Dict[str, torch.Tensor]):
        """Already packed + prefixed tensors (e.g. received through parallel.broadcast_packed)."""
        self.load_packed("", packed)
        self._check(self._lib.dtp_finalize_weights(self._h), "dtp_finalize_weights")

    # ------------------------------------------------------------------ per brush
    def encode_patches(self, patches: torch.Tensor) -> torch.Tensor:
        patches = patches.to(self.device, torch.float32).contiguous()
        assert patches.shape == (self.tokens, 3, 224, 224), patches.shape
        out = torch.empty(self.tokens, self.cfg.enc.cross_dim, device=self.device, dtype=torch.float32)
        self._check(self._lib.dtp_encode_patches(self._h, nat.ptr(patches), nat.ptr(out), self._stream(), None),
                    "dtp_encode_patches")
        return out

    def set_condition(self, emb: torch.Tensor, uncond: torch.Tensor):
        emb = emb.to(self.device, torch.float32).reshape(self.tokens, -1).contiguous()
        uncond = uncond.to(self.device, torch.float32).reshape(self.tokens, -1).contiguous()
        self._check(self._lib.dtp_set_condition(self._h, nat.ptr(emb), nat.ptr(uncond), self._stream()),
                    "dtp_set_condition")
        torch.cuda.current_stream().synchronize()  # emb / uncond are temporaries of this call

    # ------------------------------------------------------------------ per settings change
    def set_schedule(self, timesteps, alpha_t, alpha_prev, cfg: float, tg: float, tg_steps: int):
        n = len(timesteps)
        arr = lambda v: (C.c_float * max(n, 1))(*[float(x) for x in v])
        self._check(self._lib.dtp_set_schedule(self._h, n, arr(timesteps), arr(alpha_t), arr(alpha_prev), float(cfg),
                                               float(tg), int(tg_steps)), "dtp_set_schedule")

    # ------------------------------------------------------------------ per stamp
    def infer(self, masked_img, mask, ctx_img, ctx_mask, init_latents, vae_noise=None, out=None):
        B, _, R, _ = masked_img.shape
        args = [t.to(self.device, torch.float32).contiguous() for t in (masked_img, mask, ctx_img, ctx_mask, init_latents)]
        vn = vae_noise.to(self.device, torch.float32).contiguous() if vae_noise is not None else None
        if out is None:
            out = torch.empty(B, 3, R, R, device=self.device, dtype=torch.float32)
        self._check(self._lib.dtp_infer(self._h, B, R, *[nat.ptr(a) for a in args], nat.ptr(vn), nat.ptr(out),
                                        self._stream()), "dtp_infer")
        self._keep = (args, vn)  # borrowed until the stream drains
        return out

    def stamp(self, canvas, brush, pad, init_latents, vae_noise=None, composite=True, out_f32=None, out_u8=None):
        B, _, R, _ = canvas.shape
        canvas = canvas.to(self.device, torch.float32).contiguous()
        brush = brush.to(self.device, torch.float32).contiguous()
        init_latents = init_latents.to(self.device, torch.float32).contiguous()
        vn = vae_noise.to(self.device, torch.float32).contiguous() if vae_noise is not None else None
        if out_f32 is None and out_u8 is None:
            out_f32 = torch.empty(B, 3, R, R, device=self.device, dtype=torch.float32)
        self._check(self._lib.dtp_stamp(self._h, B, R, nat.ptr(canvas), nat.ptr(brush), int(pad), nat.ptr(init_latents),
                                        nat.ptr(vn), int(bool(composite)), nat.ptr(out_f32), nat.ptr(out_u8),
                                        self._stream()), "dtp_stamp")
        self._keep = (canvas, brush, init_latents, vn)
        return out_f32 if out_f32 is not None else out_u8

    # ------------------------------------------------------------------ stages
    def vae_encode(self, images, noise=None):
        images = images.to(self.device, torch.float32).contiguous()
        Nb, _, R, _ = images.shape
        noise = noise.to(self.device, torch.float32).contiguous() if noise is not None else None
        out = torch.empty(Nb, self.cfg.vae.latent_channels, R // 8, R // 8, device=self.device, dtype=torch.float32)
        self._check(self._lib.dtp_vae_encode(self._h, Nb, R, nat.ptr(images), nat.ptr(noise), nat.ptr(out),
                                             self._stream()), "dtp_vae_encode")
        self._keep = (images, noise)
        return out

    def vae_decode(self, latents):
        latents = latents.to(self.device, torch.float32).contiguous()
        B, _, h, _ = latents.shape
        out = torch.empty(B, 3, h * 8, h * 8, device=self.device, dtype=torch.float32)
        self._check(self._lib.dtp_vae_decode(self._h, B, h * 8, nat.ptr(latents), nat.ptr(out), self._stream()),
                    "dtp_vae_decode")
        self._keep = (latents,)
        return out

    def unet_forward(self, sample, step: int):
        sample = sample.to(self.device, torch.float32).contiguous()
        Bz, _, h, _ = sample.shape
        out = torch.empty(Bz, self.cfg.unet.out_channels, h, h, device=self.device, dtype=torch.float32)
        self._check(self._lib.dtp_unet_forward(self._h, Bz // 3, h * 8, nat.ptr(sample), None, None, int(step),
                                               nat.ptr(out), self._stream()), "dtp_unet_forward")
        self._keep = (sample,)
        return out

    def counter(self, name: str) -> int:
        return int(self._lib.dtp_get_counter(self._h, name.encode()))

    def set_option(self, name: str, value: int):
        self._check(self._lib.dtp_set_option(self._h, name.encode(), int(value)), f"dtp_set_option({name})")

    def profile_dump(self, path: str):
        self._check(self._lib.dtp_profile_dump(self._h, path.encode()), "dtp_profile_dump")

    KINDS = ("other", "contraction", "groupnorm", "layernorm", "softmax", "attn_small", "flash_attn")

    def profile(self):
        """{kind: (device microseconds, launches)} accumulated since set_option('profile', 1)."""
        return {k: (self.counter(f"prof_us_{i}"), self.counter(f"prof_n_{i}")) for i, k in enumerate(self.KINDS)}


def arena_estimate(cfg: W.ModelConfig, batch: int, resolution: int) -> int:
    """Generous bound on the transient activation slab for one (batch, resolution) working point."""
    h = resolution // 8
    bz = 3 * batch
    c0 = cfg.unet.block_out_channels[0]
    seq = h * h
    unet = bz * seq * c0 * 2 * 40 + (bz * cfg.unet.heads * seq * seq * 2 if seq > 64 else 0)
    v0 = cfg.vae.block_out_channels[0]
    vae = 2 * batch * resolution * resolution * v0 * 2 * 14
    vae_attn = 2 * batch * seq * seq * 2 * 2
    return int(max(unet, vae + vae_attn, 1 << 28) * 1.25) + (1 << 28)
