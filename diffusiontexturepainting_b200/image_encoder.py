"""Drop-in for trt_inference/image_encoder.py: ConditionPatchEncoder.encode_image / preprocess_image with the reference's
semantics; the CLIP ViT-B/32 visual tower, the three patch towers, final LayerNorm and projection run in the native
engine (dtp_encode_patches). The brush pre-processing (one bicubic resize, unfold, bilinear patch up-sampling — per
brush, not per stamp) uses torch's resampling ops on the device as plumbing."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def get_image_patches(image, patch_size):
    """Non-overlapping patch_size tiles in row-major order; the remainder (224 - 3*74 = 2 px) is dropped."""
    if image.dim() == 4:
        image = image.squeeze(0)
    c, H, W = image.shape
    ny, nx = H // patch_size, W // patch_size
    tiles = image[:, :ny * patch_size, :nx * patch_size].reshape(c, ny, patch_size, nx, patch_size)
    return tiles.permute(1, 3, 0, 2, 4).reshape(ny * nx, c, patch_size, patch_size)


class ConditionPatchEncoder:
    def __init__(self, engine, num_patches=(1, 4, 9)):
        self.engine = engine
        self.num_patches = tuple(num_patches)
        self.total_patches = sum(num_patches)
        dev = engine.device
        self.mean = torch.tensor(CLIP_MEAN, device=dev)
        self.std = torch.tensor(CLIP_STD, device=dev)
        self.uncond_vector = None  # (1, 14, cross) — set by the owner from the weight inventory

    def preprocess_image(self, image):
        # image_encoder.py:100-104
        if image.shape[-1] != 224 or image.shape[-2] != 224:
            image = F.interpolate(image, (224, 224), mode="bicubic", align_corners=True, antialias=False)
        return (image - self.mean[None, :, None, None]) / self.std[None, :, None, None]

    def make_patches(self, image):
        # image_encoder.py:106-113: patch sizes 224, 112, 74 -> Resize(224) (bilinear) -> (14,3,224,224)
        image = self.preprocess_image(image)
        out = []
        for n in self.num_patches:
            ps = 224 // int(math.sqrt(n))
            p = get_image_patches(image, ps)
            if ps != 224:
                p = F.interpolate(p, size=(224, 224), mode="bilinear", align_corners=False)
            out.append(p)
        return torch.cat(out, dim=0).contiguous()

    def encode_image(self, image):
        """(1,3,R,R) in [0,1] on the device -> (image_embeds (1,14,cross), uncond_vector (1,14,cross)), float32."""
        emb = self.engine.encode_patches(self.make_patches(image.float()))
        return emb.unsqueeze(0), self.uncond_vector
