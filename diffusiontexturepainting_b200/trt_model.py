"""Drop-in for trt_inference/trt_model.py: TRTConditionalInpainter with the reference's constructor, set_brush,
generate_raw and (via the base class contract) generate — the surface trt_inference/handler.py drives
(handler.py:66-110). No TensorRT anywhere: the name is kept so that `from trt_model import TRTConditionalInpainter`
in run.py:21 keeps working."""
from __future__ import annotations

import threading
import time

import torch
import torch.nn.functional as F

from . import _native as nat
from . import weights as W
from .image_encoder import ConditionPatchEncoder
from .inpaint_pipeline import InpaintPipeline
from .model_base import ConditionalInpainterBase
from .serving import BrushCache


def crop_resize_square(image, width):
    """handler.py:36-45: CenterCrop(min side) then Resize(width) (bilinear; antialias off as in the reference container's
    torchvision 0.15 tensor path)."""
    H, W_ = image.shape[-2:]
    m = min(H, W_)
    if width is None or width <= 0:
        width = m
    top, left = int(round((H - m) / 2.0)), int(round((W_ - m) / 2.0))
    img = image[..., top:top + m, left:left + m]
    if m != width:
        lead = img.dim() == 3
        img = F.interpolate(img[None] if lead else img, size=(width, width), mode="bilinear", align_corners=False)
        img = img[0] if lead else img
    return img


def _default_model():
    """(model_config, state_dicts) used when the constructor is called the way run.py:30 calls it —
    TRTConditionalInpainter(256), nothing else. (None, None) = SD-1.5-inpaint from the reference's checkpoint paths."""
    return None, None


class TRTConditionalInpainter(ConditionalInpainterBase):
    def __init__(self, resolution, device=0, model_config=None, state_dicts=None, max_batch_size=1, verbose=False,
                 preloaded=None):
        super().__init__()
        self.verbose = verbose
        # One model instance is shared by every websocket connection (run.py:30,39) and the engine handle is not
        # thread-safe: set_brush / generate / stamp_u8 serialise on this lock (the reference serialises on the IOLoop;
        # serving.StampBatcher calls generate from a worker thread).
        self.lock = threading.RLock()
        self.brush_generation = 0
        if model_config is None and state_dicts is None and preloaded is None:
            model_config, state_dicts = _default_model()
        self.pipeline = InpaintPipeline(
            scheduler="DDIM", guidance_scale=2, denoising_steps=20, texture_guidance_steps=20, version="1.5",
            hf_token="", verbose=False, nvtx_profile=False, max_batch_size=16, device=device,
            model_config=model_config, state_dicts=None)
        cfg = self.pipeline.model_config
        sds = state_dicts
        uncond = None
        if preloaded is not None:  # (packed + prefixed tensor dict, uncond_vector): multi-GPU broadcast path
            self.pipeline._prepacked, uncond = preloaded
        else:
            if sds is None:
                from .stable_diffusion_pipeline import load_state_dicts
                sds = load_state_dicts(cfg, "/workspace/checkpoints/pytorch_lora_weights.bin")
            self.pipeline._state_dicts = sds
            uncond = sds[2]["uncond_vector"]
        self.pipeline.loadEngines("/workspace/engine", "/workspace/onnx", 16, opt_batch_size=max_batch_size,
                                  opt_image_height=resolution, opt_image_width=resolution, text_maxlen=14,
                                  lora_path="/workspace/checkpoints/pytorch_lora_weights.bin",
                                  timing_cache="./timing.cache")
        self.pipeline.loadResources(resolution, resolution, batch_size=1, seed=42)
        self._arena_batch = max(1, int(max_batch_size))
        self.image_encoder = ConditionPatchEncoder(self.pipeline.engine, cfg.enc.num_patches)
        self.image_encoder.uncond_vector = uncond.float().to(self.pipeline.device)
        self._resolution = resolution
        self.conditioning = None
        self.image = None
        self._device = device
        # brush history of the client holds <= 10 brushes; set to None to re-encode on every switch
        self.brush_cache = BrushCache(10)

    def device(self):
        return self._device

    def resolution(self):
        return self._resolution

    @property
    def engine(self):
        return self.pipeline.engine

    def set_brush(self, image):
        """image: 3 x H x W float32 0..1 (trt_model.py:79-88)."""
        with self.lock:
            key = BrushCache.key(image, self.resolution()) if self.brush_cache is not None else None
            hit = self.brush_cache.get(key) if key is not None else None
            if hit is not None:
                self.image, self.conditioning = hit
            else:
                self.image = crop_resize_square(image, width=self.resolution()).unsqueeze(0) \
                    .to(self.pipeline.device).float().contiguous()
                self.conditioning = self.image_encoder.encode_image(self.image)
                if key is not None:
                    self.brush_cache.put(key, (self.image, self.conditioning))
            self.pipeline.set_condition(*self.conditioning)
            self.brush_generation += 1

    def _ensure_batch(self, B):
        """The activation arena is sized for `max_batch_size` at construction (run.py passes none -> 1); a larger coalesced
        batch (serving.StampBatcher) grows it once instead of failing with 'activation arena exhausted'."""
        if B > self._arena_batch:
            from .engine import arena_estimate
            need = arena_estimate(self.pipeline.model_config, B, self.resolution())
            self.engine.set_option("arena_mib", (need >> 20) + 1)
            self._arena_batch = B

    @staticmethod
    def _settings(settings):
        # wire values are numpy scalars (server_io.py:105-119)
        return dict(steps=int(settings["steps"]), context_pad=int(settings["context_pad"]),
                    tg_steps=int(settings["tg_steps"]), cfg_weight=float(settings["cfg_weight"]),
                    tg_weight=float(settings["tg_weight"]))

    def preprocess_canvas(self, canvas, pad):
        """trt_model.py:103-109 + handler.py:25-33 in one pair of kernels (separable flat dilation)."""
        B, _, R, _ = canvas.shape
        dev = canvas.device
        mi, ci = torch.empty(B, 3, R, R, device=dev), torch.empty(B, 3, R, R, device=dev)
        m, cm = torch.empty(B, 1, R, R, device=dev), torch.empty(B, 1, R, R, device=dev)
        scratch = torch.empty(B, R, R, device=dev)
        nat.check_op(nat.lib().dtp_op_canvas_preprocess(nat.ptr(canvas), nat.ptr(self.image), B, R, int(pad),
                                                        nat.ptr(mi), nat.ptr(m), nat.ptr(ci), nat.ptr(cm),
                                                        nat.ptr(scratch), nat.stream_ptr()), "canvas_preprocess")
        return mi, m, ci, cm

    def generate_raw(self, canvas, init_latents=None, vae_noise=None, **settings):
        """canvas: B x 4 x res x res float32 0..1 -> B x 3 x res x res float32 0..1 (trt_model.py:90-121)."""
        with self.lock:
            if self.conditioning is None:
                raise RuntimeError("set_brush must be called before generate")
            s = self._settings(settings)
            canvas = canvas.to(self.pipeline.device, torch.float32).contiguous()
            self._ensure_batch(canvas.shape[0])
            masked_images, masks, context_masked_image, context_mask = self.preprocess_canvas(canvas, s["context_pad"])
            self.pipeline.update_infer_settings(denoising_steps=s["steps"], guidance_scale=s["cfg_weight"],
                                                texture_guidance_scale=s["tg_weight"],
                                                texture_guidance_steps=s["tg_steps"])
            start = time.time()
            image_embeds, negative_embeds = self.conditioning
            result = self.pipeline.infer(prompt=image_embeds, negative_prompt=negative_embeds,
                                         input_image=masked_images, mask_image=masks,
                                         context_masked_image=context_masked_image, context_mask=context_mask,
                                         image_width=self.resolution(), image_height=self.resolution(),
                                         init_latents=init_latents, vae_noise=vae_noise)
            if self.verbose:
                torch.cuda.synchronize()
                print("Inference time:", time.time() - start)
            return result

    def generate(self, canvas, init_latents=None, vae_noise=None, **settings):
        """model_base.py:51-58 with the alpha composite as one kernel."""
        with self.lock:
            canvas = canvas.to(self.pipeline.device, torch.float32).contiguous()
            result = self.generate_raw(canvas, init_latents=init_latents, vae_noise=vae_noise, **settings)
            B, _, R, _ = canvas.shape
            out = torch.empty_like(result)
            nat.check_op(nat.lib().dtp_op_composite(nat.ptr(canvas), nat.ptr(result), B, R, nat.ptr(out), None,
                                                    nat.stream_ptr()), "composite")
            return out

    def stamp_u8(self, canvas_u8_hwc, init_latents=None, vae_noise=None, **settings):
        """Fast path for the websocket handler: uint8 HWC RGBA canvas (host or device) in, uint8 HWC RGB
        stamp out, everything between in one dtp_stamp call (pre-process, infer, composite, x255 truncation)."""
        with self.lock:
            if self.conditioning is None:
                raise RuntimeError("set_brush must be called before stamp_u8")
            s = self._settings(settings)
            dev = self.pipeline.device
            c = canvas_u8_hwc if torch.is_tensor(canvas_u8_hwc) else torch.from_numpy(canvas_u8_hwc)
            if c.dim() == 3:
                c = c.unsqueeze(0)
            canvas = (c.to(dev, non_blocking=True).to(torch.float32).permute(0, 3, 1, 2) / 255).contiguous()
            B, _, R, _ = canvas.shape
            h = R // 8
            self._ensure_batch(B)
            if not self.pipeline._same_condition(*self.conditioning):
                self.pipeline.set_condition(*self.conditioning)
            self.pipeline.update_infer_settings(s["steps"], s["cfg_weight"], s["tg_weight"], s["tg_steps"])
            self.pipeline._push_schedule(1.0)
            if init_latents is None:
                init_latents = self.pipeline.initialize_latents(B, 4, h, h)
            if vae_noise is None and self.pipeline.sample_posterior:
                vae_noise = torch.randn((2 * B, 4, h, h), device=dev, dtype=torch.float32,
                                        generator=self.pipeline.noise_generator)
            out = torch.empty(B, R, R, 3, device=dev, dtype=torch.uint8)
            self.engine.stamp(canvas, self.image, s["context_pad"], init_latents, vae_noise, composite=True, out_u8=out)
            return out

    # stable_diffusion_pipeline.py:486-503 (print_summary): per-stage device time of the last stamps run with
    # enable_stage_timers(True) — the reference's cudart events around 'vae_encoder' / 'unet' / 'vae'
    STAGES = ("canvas_preprocess", "vae_encoder", "unet", "latent_step", "vae", "composite")

    def enable_stage_timers(self, on=True, nvtx=False):
        self.engine.set_option("stage_timers", int(bool(on)))
        self.engine.set_option("nvtx", int(bool(nvtx)))

    def stage_times_ms(self):
        return {n: (self.engine.counter(f"stage_us_{i}") / 1e3, self.engine.counter(f"stage_n_{i}"))
                for i, n in enumerate(self.STAGES)}

    def print_summary(self):
        t = self.stage_times_ms()
        print("|------------|--------------|")
        print("| {:^10} | {:^12} |".format("Module", "Latency"))
        print("|------------|--------------|")
        for name, (ms, n) in t.items():
            print("| {:^10} | {:>9.2f} ms |".format(name[:10] + (" x %d" % n if n > 1 else ""), ms))
        print("|------------|--------------|")
        print("| {:^10} | {:>9.2f} ms |".format("Pipeline", sum(v[0] for v in t.values())))
        print("|------------|--------------|")
