"""Weight ingestion for real checkpoints. Pure host logic: reads the files the reference reads and
maps their key spellings onto the inventory of `weights.py`, checking every shape. The formats:

* `runwayml/stable-diffusion-inpainting` in diffusers layout: `{unet,vae}/diffusion_pytorch_model.{safetensors,bin}`
  (trt_inference/models.py:795-813, loaded through diffusers 0.12.0). VAE attention projections are spelled
  `query / key / value / proj_attn` in that release and `to_q / to_k / to_v / to_out.0` in files re-saved by diffusers
  >= 0.17 (the training side pins 0.17, training/requirements.txt:1); some exports keep them as 1x1 convolutions.
* `pytorch_lora_weights.bin`: `<attn>.processor.to_{q,k,v,out}_lora.{down,up}.weight`, written by
  `unet.save_attn_procs` (training/train_texture_inpaint_lora.py:787); loaders of newer diffusers prefix the keys with `unet.`.
* `image_encoder.pth`: `state_dict()` of the TRAINING `ConditionPatchEncoder` (train_texture_inpaint_lora.py:789), whose CLIP
  tower is `transformers.CLIPVisionModel` (`clip.vision_model.*`, training/image_encoder.py:39). The inference module is built
  on openai-CLIP (`clip.visual.*`, trt_inference/image_encoder.py:49-50) and loads the file with `strict=False`
  (trt_model.py:58-59), i.e. it silently keeps the downloaded openai weights for the tower. Both spellings name the same
  pretrained ViT-B/32, so here the tower is taken from the file: `clip.vision_model.*` is rewritten to `clip.visual.*`
  (q/k/v projections concatenated into `in_proj`), which needs no download.
"""
from __future__ import annotations

import dataclasses
import os
import re
from typing import Dict, List, Optional, Tuple

import torch


@dataclasses.dataclass
class IngestReport:
    what: str
    loaded: List[str] = dataclasses.field(default_factory=list)
    missing: List[str] = dataclasses.field(default_factory=list)       # inventory keys the file does not provide
    unexpected: List[str] = dataclasses.field(default_factory=list)    # file keys the inventory does not know
    mismatched: List[Tuple[str, tuple, tuple]] = dataclasses.field(default_factory=list)  # (key, file shape, wanted)

    @property
    def ok(self) -> bool:
        return not self.missing and not self.mismatched

    def summary(self) -> str:
        s = f"{self.what}: {len(self.loaded)} tensors loaded"
        if self.missing:
            s += f", {len(self.missing)} missing (e.g. {self.missing[0]})"
        if self.unexpected:
            s += f", {len(self.unexpected)} unexpected (e.g. {self.unexpected[0]})"
        if self.mismatched:
            k, a, b = self.mismatched[0]
            s += f", {len(self.mismatched)} shape mismatches (e.g. {k}: file {a} vs inventory {b})"
        return s


class CheckpointError(RuntimeError):
    pass


# ----------------------------------------------------------------------------------------------------------------
# key normalisation
# ----------------------------------------------------------------------------------------------------------------
_VAE_ATTN_NEW_TO_OLD = {"to_q": "query", "to_k": "key", "to_v": "value", "to_out.0": "proj_attn"}


def normalize_vae_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """diffusers >= 0.17 attention spellings -> the 0.12.0 ones of the inventory; 1x1-conv projections -> matrices."""
    out = {}
    for k, v in sd.items():
        m = re.match(r"(.*\.attentions\.\d+)\.(to_q|to_k|to_v|to_out\.0)\.(weight|bias)$", k)
        if m:
            k = f"{m.group(1)}.{_VAE_ATTN_NEW_TO_OLD[m.group(2)]}.{m.group(3)}"
        if re.search(r"\.attentions\.\d+\.(query|key|value|proj_attn)\.weight$", k) and v.dim() == 4 and v.shape[2:] == (1, 1):
            v = v[:, :, 0, 0]
        out[k] = v
    return out


def normalize_unet_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Strip the `unet.` prefix newer LoRA writers add; proj_in / proj_out saved as linear layers (use_linear_projection
    exports) are reshaped to the 1x1-conv form of the SD-1.5 inventory."""
    out = {}
    for k, v in sd.items():
        if k.startswith("unet."):
            k = k[5:]
        if re.search(r"attentions\.\d+\.proj_(in|out)\.weight$", k) and v.dim() == 2:
            v = v[:, :, None, None]
        out[k] = v
    return out


def hf_clip_vision_to_openai(sd: Dict[str, torch.Tensor], prefix: str = "clip.vision_model.",
                             out_prefix: str = "clip.visual.") -> Dict[str, torch.Tensor]:
    """transformers.CLIPVisionModel keys -> openai-CLIP visual keys (same tensors; q/k/v concatenated as in_proj).
    Keys outside `prefix` pass through unchanged."""
    out: Dict[str, torch.Tensor] = {}
    qkv: Dict[Tuple[int, str], Dict[str, torch.Tensor]] = {}
    simple = {"embeddings.class_embedding": "class_embedding", "embeddings.patch_embedding.weight": "conv1.weight",
              "embeddings.position_embedding.weight": "positional_embedding", "pre_layrnorm.weight": "ln_pre.weight",
              "pre_layrnorm.bias": "ln_pre.bias", "post_layernorm.weight": "ln_post.weight",
              "post_layernorm.bias": "ln_post.bias"}
    layer_map = {"layer_norm1": "ln_1", "layer_norm2": "ln_2", "self_attn.out_proj": "attn.out_proj", "mlp.fc1": "mlp.c_fc",
                 "mlp.fc2": "mlp.c_proj"}
    for k, v in sd.items():
        if not k.startswith(prefix):
            out[k] = v
            continue
        r = k[len(prefix):]
        if r in simple:
            out[out_prefix + simple[r]] = v
            continue
        if r == "embeddings.position_ids":
            continue
        m = re.match(r"encoder\.layers\.(\d+)\.(.*)\.(weight|bias)$", r)
        if not m:
            out[k] = v  # reported as unexpected by the ingestion
            continue
        i, name, wb = int(m.group(1)), m.group(2), m.group(3)
        if name in layer_map:
            out[f"{out_prefix}transformer.resblocks.{i}.{layer_map[name]}.{wb}"] = v
        elif name in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj"):
            qkv.setdefault((i, wb), {})[name[-6]] = v
        else:
            out[k] = v
    for (i, wb), parts in qkv.items():
        if set(parts) != {"q", "k", "v"}:
            raise CheckpointError(f"CLIP layer {i}: incomplete q/k/v projection ({sorted(parts)})")
        out[f"{out_prefix}transformer.resblocks.{i}.attn.in_proj_{wb}"] = torch.cat([parts["q"], parts["k"], parts["v"]], 0)
    return out


# ----------------------------------------------------------------------------------------------------------------
# ingestion
# ----------------------------------------------------------------------------------------------------------------
def ingest(target: Dict[str, torch.Tensor], source: Dict[str, torch.Tensor], what: str,
           required: Optional[callable] = None) -> IngestReport:
    """Copy every tensor of `source` whose key and shape match the inventory `target` (in place, as float32). Keys of
    `target` selected by `required` (default: all) that the file lacks are reported as missing."""
    rep = IngestReport(what)
    for k, v in source.items():
        if not torch.is_tensor(v):
            continue
        if k not in target:
            rep.unexpected.append(k)
            continue
        if tuple(v.shape) != tuple(target[k].shape):
            rep.mismatched.append((k, tuple(v.shape), tuple(target[k].shape)))
            continue
        target[k] = v.detach().to(torch.float32)
        rep.loaded.append(k)
    have = set(rep.loaded)
    rep.missing = [k for k in target if k not in have and (required is None or required(k))]
    return rep


def _read(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    sd = torch.load(path, map_location="cpu", weights_only=True)
    if isinstance(sd, dict) and "state_dict" in sd and isinstance(sd["state_dict"], dict):
        sd = sd["state_dict"]
    return sd


def _first_existing(paths) -> Optional[str]:
    for p in paths:
        if p and os.path.exists(p):
            return p
    return None


def load_real_checkpoints(unet_sd, vae_sd, enc_sd, hf_dir: str, lora_path: Optional[str], encoder_path: Optional[str],
                          strict: bool = True) -> List[IngestReport]:
    """Overwrite the (synthetic) inventories with whatever real checkpoint files exist at the reference's paths. A file
    that exists must load completely when `strict` (every inventory tensor of that component present, every shape equal):
    a half-loaded network would silently produce wrong textures. Absent files leave the inventory untouched."""
    reports: List[IngestReport] = []

    def finish(rep: IngestReport):
        reports.append(rep)
        if strict and not rep.ok:
            raise CheckpointError(rep.summary())

    p = _first_existing(os.path.join(hf_dir, "unet", f) for f in ("diffusion_pytorch_model.safetensors",
                                                                 "diffusion_pytorch_model.bin"))
    if p:
        finish(ingest(unet_sd, normalize_unet_keys(_read(p)), f"unet <- {p}", required=lambda k: ".processor." not in k))
    p = _first_existing(os.path.join(hf_dir, "vae", f) for f in ("diffusion_pytorch_model.safetensors",
                                                                "diffusion_pytorch_model.bin"))
    if p:
        finish(ingest(vae_sd, normalize_vae_keys(_read(p)), f"vae <- {p}"))
    if lora_path and os.path.exists(lora_path):
        finish(ingest(unet_sd, normalize_unet_keys(_read(lora_path)), f"lora <- {lora_path}",
                      required=lambda k: ".processor." in k))
    if encoder_path and os.path.exists(encoder_path):
        src = hf_clip_vision_to_openai(_read(encoder_path))
        src = {k: v for k, v in src.items() if k not in ("pos_emb", "mean", "std")}
        # files written from the openai-CLIP based inference module have no tower at all when saved with the tower
        # excluded; then the tower keys are simply missing and strict mode says so
        finish(ingest(enc_sd, src, f"image encoder <- {encoder_path}"))
    return reports
