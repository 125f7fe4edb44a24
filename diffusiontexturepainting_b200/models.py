"""Drop-in for trt_inference/models.py. The reference module describes the three networks for ONNX export and TensorRT
build (UNet :1017-1234 incl. the LoRA fuse :1042-1093, VAE :1237-1320, VAEEncoder :1338-1420) and rewrites the ONNX
graphs to use plugin ops (Optimizer :25-792). Here the networks are launch plans inside libdtp_sm100.so
(csrc/runtime.cu); what remains on the host is the weight inventory, the LoRA merge and the packing, re-exported from
weights.py under names that mirror the reference's factory functions."""
from .weights import (ModelConfig, UNetConfig, VAEConfig, EncoderConfig, sd15_config, tiny_config, unet_param_shapes,
                      vae_param_shapes, encoder_param_shapes, synth_model, merge_lora, pack_unet, pack_vae,
                      pack_encoder)


def make_UNet(cfg=None):
    return unet_param_shapes((cfg or sd15_config()).unet)


def make_VAE(cfg=None):
    return {k: v for k, v in vae_param_shapes((cfg or sd15_config()).vae).items()
            if k.startswith(("decoder.", "post_quant_conv."))}


def make_VAEEncoder(cfg=None):
    return {k: v for k, v in vae_param_shapes((cfg or sd15_config()).vae).items()
            if k.startswith(("encoder.", "quant_conv."))}
