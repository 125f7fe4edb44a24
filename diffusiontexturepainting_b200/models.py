"""Drop-in for trt_inference/models.py. The reference module describes the three networks for ONNX export and TensorRT
build (UNet :1017-1234 incl. the LoRA fuse :1042-1093, VAE :1237-1320, VAEEncoder :1338-1420) and rewrites the ONNX
graphs to use plugin ops (Optimizer :25-792). Here the networks are launch plans inside libdtp_sm100.so
(csrc/runtime.cu); what remains on the host is the weight inventory, the LoRA merge and the packing in weights.py, whose
names this module re-exports so that `import models` keeps resolving for third-party scripts."""
from .weights import (ModelConfig, UNetConfig, VAEConfig, EncoderConfig, sd15_config, tiny_config,  # noqa: F401
                      unet_param_shapes, vae_param_shapes, encoder_param_shapes, synth_model, merge_lora, pack_unet,
                      pack_vae, pack_encoder)
