"""ORACLE — TEST INFRASTRUCTURE ONLY.

Plain-PyTorch (CPU fp32 by default) restatement of the reference's brush-stamp path
(nv-tlabs/DiffusionTexturePainting, trt_inference/): the diffusers-0.12.0 module graphs the reference exports to
TensorRT (UNet2DConditionModel SD-1.5-inpaint + merged LoRA, AutoencoderKL, ConditionPatchEncoder with the CLIP
ViT-B/32 visual tower) driven by the reference's own loop logic (inpaint_pipeline.py:52-153,
stable_diffusion_pipeline.py:340-355,407-484, utilities.py:370-529, trt_model.py:90-121, handler.py:25-60,
model_base.py:51-58). Every function cites the reference file:line it follows.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package, and only
as the checker or the timed CPU baseline — never as the shipped path. The product (diffusiontexturepainting_b200/) never
imports it.

PARITY PINNING STATUS: the reference ships no tests, golden vectors or fixtures for this path, and its
third-party arithmetic (diffusers 0.12.0, openai-CLIP, kornia, TensorRT 8.6) is absent from /root/reference and from
this image. What IS pinned against the reference's own executable code (tests/test_oracle_reference.py, run in the build
container where /root/reference exists; vectors committed under tests/golden/ by tests/golden/make_golden.py):
  * DDIMScheduler tables and step()  (utilities.py:370-529, class executed in place via AST extraction);
  * initialize_timesteps t_start quirk (stable_diffusion_pipeline.py:348-355);
  * positional_encoding_2d / get_image_patches / pos_emb view-scramble (image_encoder.py:20-56);
  * the wire codec (server_io.py) round trip;
  * the CLIP visual tower against transformers.CLIPVisionModel (same graph as openai-CLIP with proj=None, as the
    reference's training/image_encoder.py:39,68 relies on).
The UNet / VAE graphs come from the absent third-party diffusers==0.12.0. They are pinned to code
that was not written for this repository (oracle/independent.py, tests/test_oracle_independent.py):
  * AutoencoderKL encoder == transformers' ChameleonVQVAEEncoder and decoder == transformers' JanusVQVAEDecoder (both are
    the latent-diffusion / taming Encoder / Decoder that diffusers' class derives from) with the diffusers keys renamed
    and loaded strict=True: whole-network agreement <= 2e-5 rel-L2 in fp32;
  * UNet2DConditionModel == a torch.nn module tree (nn.GroupNorm / Conv2d / LayerNorm / Linear,
    F.multi_head_attention_forward) whose parameter names are the diffusers-0.12.0 inventory (strict load, 859.5 M
    parameters at SD-1.5 widths): agreement <= 2e-5 rel-L2;
  * kornia's flat dilation == scipy.ndimage.maximum_filter (bit-exact).
STILL UNPINNED: agreement with the real diffusers classes themselves (tests/golden/make_diffusers_golden.py writes the
fixture where diffusers is installed; the test reports its absence) and with the TensorRT engines' fp16 numerics.
"""
