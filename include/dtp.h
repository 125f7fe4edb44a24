/*
 * dtp.h — C-ABI of libdtp_sm100.so: the sm_100a brush-stamp inpainting path of Diffusion Texture Painting.
 *
 * This is the boundary a maintainer of nv-tlabs/DiffusionTexturePainting binds (ctypes from Python; see
 * INTEGRATION.md) in place of the TensorRT engine layer of trt_inference/:
 *
 *   reference interface (file:line, relative to the reference repo)             entry point here
 *   -------------------------------------------------------------------------   --------------------------------
 *   Engine.load/activate/allocate_buffers   trt_inference/utilities.py:221-250   dtp_create, dtp_set_tensor,
 *   StableDiffusionPipeline.loadEngines     stable_diffusion_pipeline.py:189-334   dtp_finalize_weights
 *   UNet.get_model LoRA merge               trt_inference/models.py:1042-1093     (host side, weights.py) -> dtp_set_tensor
 *   ConditionPatchEncoder.forward           trt_inference/image_encoder.py:78-98  dtp_encode_patches
 *   text_embeddings = cat([neg,prompt,prompt]) inpaint_pipeline.py:140            dtp_set_condition
 *   InpaintPipeline.update_infer_settings   inpaint_pipeline.py:39-50             dtp_set_schedule
 *   InpaintPipeline.infer                   inpaint_pipeline.py:52-153            dtp_infer
 *   TRTConditionalInpainter.generate_raw    trt_inference/trt_model.py:90-121     dtp_stamp (pre-process + infer [+ composite])
 *   encode_image -> Engine.infer('vae_encoder')  stable_diffusion_pipeline.py:464-474   dtp_vae_encode
 *   Engine.infer('unet')                    stable_diffusion_pipeline.py:436-441  dtp_unet_forward
 *   scheduler.step + guidance               stable_diffusion_pipeline.py:449-455, utilities.py:441-522   dtp_op_ddim_step
 *   decode_latent -> Engine.infer('vae')    stable_diffusion_pipeline.py:476-484  dtp_vae_decode
 *   add_extra_context (kornia dilation)     trt_inference/handler.py:25-33        dtp_op_canvas_preprocess
 *   ConditionalInpainterBase.generate       trt_inference/model_base.py:51-58     dtp_op_composite
 *
 * Conventions: every function returns 0 on success and a negative code on failure (never throws across the ABI);
 * dtp_last_error()/dtp_ops_last_error() return a human-readable reason. All pointers are BORROWED DEVICE pointers
 * unless the name says "host"; they must stay valid until the stream reaches the end of the call. Work is enqueued
 * on the caller's stream (pass torch.cuda.current_stream().cuda_stream); no hidden synchronisation unless stated.
 * A handle is bound to the CUDA device current at dtp_create and is not thread-safe (one handle per GPU / process).
 * "f16" buffers are IEEE binary16; activations are NHWC (rows = pixels, row length = channels).
 */
#ifndef DTP_H_
#define DTP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------------------
 * epilogue flags of the contraction entry points (same values as dtp::GemmFlags)
 * ---------------------------------------------------------------------------------------------------------- */
#define DTP_EPI_GELU (1 << 0)
#define DTP_EPI_QUICKGELU (1 << 1)
#define DTP_EPI_SILU (1 << 2)
#define DTP_EPI_GEGLU (1 << 3)
#define DTP_EPI_BIAS_M (1 << 4)
#define DTP_EPI_OUT_F32_NCHW (1 << 5)
#define DTP_EPI_IMG01 (1 << 6)
#define DTP_EPI_OUT_F32 (1 << 7)

/* ------------------------------------------------------------------------------------------------------------
 * operator entry points (one kernel family per call; parity tests and roofline isolation)
 * ---------------------------------------------------------------------------------------------------------- */
const char* dtp_ops_last_error(void);
/* tuning aid: per-CTA globaltimer checkpoints of the contraction kernel (see csrc/capi_ops.cu); NULL disables */
void dtp_ops_set_debug_buffer(long long* dbg);

/* out[M,N] = epilogue(alpha * [A0 | A1][M,K0+K1] * Wt[N,K0+K1]^T + bias) (+ residual). BN<=0: pick tile/split.
 * BN | 0x1000 runs the problem as CTA pairs (cta_group::2 MMA over two m-tiles, each CTA fetching half of the weight tile). */
int dtp_op_linear(const void* A0, int lda0, int K0, const void* A1, int lda1, int K1, int M, const void* Wt, int ldw,
                  int N, const float* bias, const void* residual, int ldr, void* out, int ldc, int flags, float alpha,
                  int hw_out, int BN, int splits, void* stream);
/* 3x3 stride-1 pad-1 conv over NHWC sources (Nimg,H,W,C0) [| (Nimg,H,W,C1)]; Wt [Cout, 9*(C0+C1)], k = tap*C + c */
int dtp_op_conv3x3(const void* A0, int C0, const void* A1, int C1, int Nimg, int H, int W, const void* Wt, int Cout,
                   const float* bias, const void* residual, int ldr, void* out, int ldc, int flags, float alpha,
                   int hw_out, int BN, int splits, void* stream);
/* ResnetBlock2D tail in one contraction: out = conv3x3(A0) + conv1x1([S0 | S1]) + bias. A0 (Nimg,H,W,C0), shortcut sources
 * S0 (Nimg,H,W,CS0) and optional S1 (Nimg,H,W,CS1) NHWC f16; Wt [Cout, 9*C0 + CS0 + CS1] (conv2 rows with the conv_shortcut
 * rows appended along K), bias = conv2.bias + conv_shortcut.bias (diffusers ResnetBlock2D.forward: output = shortcut(x) + h) */
int dtp_op_conv3x3_shortcut(const void* A0, int C0, const void* S0, int CS0, const void* S1, int CS1, int Nimg, int H, int W,
                            const void* Wt, int Cout, const float* bias, void* out, int BN, int splits, void* stream);
/* Downsample2D: conv 3x3 stride 2 straight from the NHWC input (four parity-view tensor maps, no im2col buffer). A (Nimg,H,W,C)
 * f16, H and W even; pad_lo = 1: padding 1 (UNet), 0: F.pad (0,1,0,1) then no padding (VAE encoder); out (Nimg,H/2,W/2,Cout). */
int dtp_op_conv3x3_s2(const void* A, int C, int Nimg, int H, int W, const void* Wt, int Cout, const float* bias, int pad_lo,
                      void* out, int BN, int splits, void* stream);
/* Upsample2D (diffusers: F.interpolate(scale_factor=2, mode="nearest") then conv 3x3 pad 1; the reference's graph folds the
 * resize, models.py:128-186) as ONE contraction over the half-resolution input: four output-parity classes, each a 2x2
 * convolution with pre-summed taps (4/9 of the multiply-adds, no upsampled tensor). A (Nimg,H,W,C) NHWC f16, Wt [Cout, 9*C]
 * (the conv's packed weights; NULL = wstack already holds the folded weights of an earlier call), wstack: the folded weights
 * [4*Cout, 4*C] f16, out (Nimg,2H,2W,Cout) f16. */
int dtp_op_upconv2x(const void* A, int C, int Nimg, int H, int W, const void* Wt, int Cout, const float* bias, void* wstack,
                    void* out, int BN, void* stream);
/* batched D_z = alpha * A_z * B_z^T over z = (z1 < nz1, z2 < nz2); b_mn: B_z given as [K,N] row-major */
int dtp_op_bmm(const void* A, int lda, long long a_zs1, long long a_zs2, const void* B, int ldb, long long b_zs1,
               long long b_zs2, int b_mn, int M, int N, int K, int nz1, int nz2, void* out, int ldc, long long out_zs1,
               long long out_zs2, float alpha, int flags, int BN, void* stream);
int dtp_op_groupnorm(const void* x0, int C0, const void* x1, int C1, int Nimg, int HW, int groups, const float* gamma,
                     const float* beta, float eps, int silu, void* out, void* stream);
int dtp_op_layernorm(const void* x, int rows, int C, const float* gamma, const float* beta, float eps, void* out,
                     void* stream);
int dtp_op_softmax(void* x, long long rows, int cols, int ld, void* stream);
int dtp_op_attn_small(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int nq,
                      int nkv, int heads, int d, int batch, long long q_bs, long long kv_bs, long long o_bs,
                      const int* kv_index, float scale, void* stream);
/* tcgen05 flash self-attention over packed rows: q/k/v point at head 0 (row stride ld, samples bs elements apart);
 * out[b][row][head*d + c] = softmax(q k^T / sqrt(d)) v. d % 8 == 0, d <= 192. */
int dtp_op_flash_attn(const void* q, const void* k, const void* v, int ld, long long bs, void* out, int ldo,
                      long long o_bs, int seq, int heads, int d, int batch, void* stream);
int dtp_op_upsample2x(const void* x, int Nimg, int H, int W, int C, void* out, void* stream);
int dtp_op_im2col_s2(const void* x, int Nimg, int H, int W, int C, int pad_lo, int Ho, int Wo, void* out, void* stream);
/* eps3 (3B,chw) f32 [uncond|cond|tg]; DDIM eta=0 step with the reference's operation order */
int dtp_op_ddim_step(const float* eps3, const float* latents_in, float* latents_out, int B, int chw, float cfg, float tg,
                     float alpha_t, float alpha_prev, void* stream);
int dtp_op_pack_unet_input(const float* latents, const float* mask3, const float* masked3, int B, int hw, void* out,
                           void* stream);
int dtp_op_nchw_to_nhwc_pad(const float* x, int Nimg, int C, int HW, int Cpad, float divisor, void* out, void* stream);
/* canvas (B,4,R,R) f32 0..1, brush (1,3,R,R) f32 0..1 -> masked image, mask (1 = generate), context pair. scratch: B*R*R floats */
int dtp_op_canvas_preprocess(const float* canvas, const float* brush, int B, int R, int pad, float* masked_img,
                             float* mask, float* ctx_img, float* ctx_mask, float* scratch, void* stream);
/* rgb*alpha + raw*(1-alpha); optional uint8 HWC copy (truncating x255 as handler.py:55-56) */
int dtp_op_composite(const float* canvas, const float* raw, int B, int R, float* out_f32, unsigned char* out_u8hwc,
                     void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * pipeline entry points
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct dtp_config {
    /* UNet2DConditionModel (SD-1.5 inpainting: 9 in, 4 out, [320,640,1280,1280], 2 layers, 8 heads, ctx 768) */
    int unet_in_channels, unet_out_channels;
    int unet_block_out[4];
    int unet_down_attn[4];
    int unet_layers_per_block;
    int unet_heads;
    int unet_cross_dim;
    int groups; /* GroupNorm groups (32) for UNet and VAE */
    /* AutoencoderKL ([128,256,512,512], 2 layers, 4 latent channels) */
    int vae_block_out[4];
    int vae_layers_per_block;
    int vae_latent;
    /* ConditionPatchEncoder: CLIP ViT-B/32 visual (768, 12 layers, 12 heads, mlp 3072) + 3 towers of 4 blocks x 4 heads */
    int enc_width, enc_layers, enc_heads, enc_mlp, enc_tower_layers, enc_tower_heads, enc_cross_dim;
    int enc_tokens; /* 14 = 1 + 4 + 9 patches */
    unsigned long long arena_bytes; /* transient activation slab; 0 = 8 GiB */
} dtp_config;

typedef struct dtp_engine dtp_handle;

int dtp_create(const dtp_config* cfg, dtp_handle** out);
void dtp_destroy(dtp_handle* h);
const char* dtp_last_error(dtp_handle* h);

/* Hand one named tensor (host OR device memory, dtype 0 = f32, 1 = f16) to the engine; it is copied (cudaMemcpyDefault). Names and layouts
 * are those produced by diffusiontexturepainting_b200/weights.py pack_unet / pack_vae / pack_encoder with the prefixes
 * "unet.", "vae.", "enc.". Replaces Engine.load + the ONNX/plan caches (stable_diffusion_pipeline.py:263-334). */
int dtp_set_tensor(dtp_handle* h, const char* name, const void* host_ptr, const long long* shape, int ndim, int dtype);
int dtp_finalize_weights(dtp_handle* h);

/* ConditionPatchEncoder.forward (image_encoder.py:78-98): patches (14,3,224,224) f32 normalised -> emb (14, cross) f32 */
int dtp_encode_patches(dtp_handle* h, const float* patches, float* emb_out, void* stream, void* reserved);
/* text_embeddings = cat([negative, prompt, prompt]).half() (inpaint_pipeline.py:140) + cross-attention K/V of all
 * transformer layers projected once for both embeddings. emb / uncond: (14, cross) f32. */
int dtp_set_condition(dtp_handle* h, const float* emb, const float* uncond, void* stream);
/* Explicit evaluation schedule (host arrays of length n): timestep, alpha_cumprod[t], alpha_cumprod[t_prev] per UNet
 * evaluation; cfg / tg weights; tg is forced to 0 from evaluation index tg_steps on
 * (stable_diffusion_pipeline.py:419-420). The reference's `steps = S` semantics (S-1 evaluations) is produced by the
 * Python façade (inpaint_pipeline.py), a strict S-evaluation schedule by passing all S entries. */
int dtp_set_schedule(dtp_handle* h, int n, const float* timesteps, const float* alpha_t, const float* alpha_prev,
                     float cfg, float tg, int tg_steps);
/* InpaintPipeline.infer: all tensors f32 NCHW on the device. masked_img/ctx_img (B,3,R,R) in [-1,1]; mask/ctx_mask
 * (B,1,R,R) with 1 = generate; init_latents (B,4,R/8,R/8); vae_noise (2B,4,R/8,R/8) or NULL (posterior mode);
 * out_images (B,3,R,R) in [0,1]. */
int dtp_infer(dtp_handle* h, int B, int R, const float* masked_img, const float* mask, const float* ctx_img,
              const float* ctx_mask, const float* init_latents, const float* vae_noise, float* out_images, void* stream);
/* TRTConditionalInpainter.generate_raw (composite = 0) / ConditionalInpainterBase.generate (composite = 1):
 * canvas (B,4,R,R) f32 0..1, brush (1,3,R,R) f32 0..1. out_f32 (B,3,R,R) and/or out_u8 (B,R,R,3), either nullable. */
int dtp_stamp(dtp_handle* h, int B, int R, const float* canvas, const float* brush, int pad, const float* init_latents,
              const float* vae_noise, int composite, float* out_f32, unsigned char* out_u8, void* stream);
/* stage entry points (roofline isolation and per-stage parity) */
int dtp_vae_encode(dtp_handle* h, int Nb, int R, const float* images, const float* noise, float* latents_out,
                   void* stream); /* latents = 0.18215 * sample */
int dtp_vae_decode(dtp_handle* h, int B, int R, const float* latents, float* images_out,
                   void* stream); /* images = clamp(decode(latents / 0.18215)/2 + 0.5, 0, 1) */
/* one UNet evaluation at schedule index `step`: sample (3B,9,h,h) f32 -> eps (3B,4,h,h) f32 */
int dtp_unet_forward(dtp_handle* h, int B, int R, const float* sample, const float* reserved0, const float* reserved1,
                     int step, float* eps_out, void* stream);
/* counters: "launches" (kernels launched by this engine so far), "stamps", "arena_peak", "arena_bytes", "graph_launches",
 * "unet_plan_ops", "device", "stage_us_<k>" / "stage_n_<k>" (k = 0..5: canvas_preprocess, vae_encoder, unet, latent_step,
 * vae, composite; device time and count accumulated since dtp_set_option("stage_timers", 1)).
 * options: "graph" (CUDA-graph replay of a stamp, default 1), "fold_cross", "fuse_cross" (image-token cross-attention as one
 * launch), "fold_ln" (LayerNorm folded into the consuming contraction), "fold_ln_ff_rows", "fuse_shortcut" (conv_shortcut inside
 * conv2), "fuse_ff_out" (feed-forward output projection folded into the transformer's proj_out), "dedup_branches" (stamp path: the layers in front of the first cross-attention run once for the
 * uncond and the cond branch, whose sample inputs are identical), "splitk_f16" (fp16 partials in the in-kernel split-K reduction;
 * process-wide), "fold_downsample" (stride-2 convolutions
 * without an im2col buffer), "fold_upsample" (nearest-2x upsample
 * folded into its 3x3 convolution; "fold_upsample_rows": smallest output pixel count it applies to), "flash", "profile" (per-op events),
 * "stage_timers", "nvtx" (NVTX ranges per stage), "arena_mib" (grow the activation arena to at least this size). */
long long dtp_get_counter(dtp_handle* h, const char* name);
int dtp_set_option(dtp_handle* h, const char* name, int value);
/* after dtp_set_option(h, "profile", 1): CSV of device time per distinct op (label, calls, microseconds), slowest first */
int dtp_profile_dump(dtp_handle* h, const char* path);

#ifdef __cplusplus
}
#endif
#endif /* DTP_H_ */
