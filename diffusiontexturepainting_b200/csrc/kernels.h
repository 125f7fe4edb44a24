// HBM-bound kernels of the stamp path (norms, softmax, small attention, resampling, scheduler step, canvas pre/post).
// Activations are NHWC fp16 (rows = pixels, row length = channels); latents / images at the boundary are NCHW fp32
// exactly as the reference façade hands them over (trt_inference/inpaint_pipeline.py:52-153).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace dtp {

const char* kernels_last_error();
// debug aid: per-CTA globaltimer checkpoints (8 slots per CTA) of the single-launch GroupNorm kernel; nullptr = off
void kernels_set_debug(long long* dbg);

// GroupNorm (+SiLU) over one or two NHWC sources (channel concat, never materialised before the norm).
// stats_ws: float[Nimg * chunks * groups * 2 + Nimg * groups * 2]
int gn_num_chunks(int HW, int C);
size_t gn_ws_floats(int Nimg, int HW, int C, int groups);
// folds conv3x3(nearest_upsample_2x(.)) weights [Cout, 9*C] into the stacked parity-class weights [4*Cout, 4*C] (gemm_tc.h)
int launch_upconv_fold_weights(const __half* W, int Cout, int C, __half* Wst, cudaStream_t st);
int launch_groupnorm(const __half* x0, int C0, const __half* x1, int C1, int Nimg, int HW, int groups,
                     const float* gamma, const float* beta, float eps, int silu, __half* out, float* stats_ws,
                     cudaStream_t st, int* launches = nullptr);  // *launches: kernels enqueued (1 or 2)

int launch_layernorm(const __half* x, int rows, int C, const float* gamma, const float* beta, float eps, __half* out,
                     cudaStream_t st);
// in-place row softmax (fp32 math) over fp16 rows
int launch_softmax_rows(__half* x, long long rows, int cols, int ld, cudaStream_t st);

// attention with a short key/value sequence (nkv <= 64) held in shared memory: cross-attention on the 14 image tokens,
// CLIP (50 tokens), patch towers (1/4/9 tokens) and UNet self-attention at <= 8x8 latents.
// q: [batch][nq][ldq] with head h at columns h*d..; k/v: [kvbatch][nkv][ld*]; kv_index (device, nullable) maps batch ->
// kvbatch.
int launch_attn_small(const __half* q, int ldq, const __half* k, int ldk, const __half* v, int ldv, __half* out, int ldo,
                      int nq, int nkv, int heads, int d, int batch, long long q_bs, long long kv_bs, long long o_bs,
                      const int* kv_index, float scale, cudaStream_t st);

// tcgen05 flash self-attention over packed rows: q/k/v point at the first head's columns, row stride ld, samples bs
// elements apart; out[b][row][head*d + c]. Head dim d % 8 == 0, d <= 192.
struct FlashOp {
    CUtensorMap mq, mk, mv;
    int seq, heads, d, batch;
    __half* out;
    int ldo;
    long long o_bs;
};
int flash_attn_setup(FlashOp* op, const __half* q, const __half* k, const __half* v, int ld, long long bs, __half* out,
                     int ldo, long long o_bs, int seq, int heads, int d, int batch);
int flash_attn_launch(const FlashOp* op, cudaStream_t st);
const char* flash_last_error();

// Fused image-token cross-attention of a transformer block (cross_attn.cu): h += softmax_T(LN2(h) Wscore_z^T) Wout_z^T + bias,
// one launch instead of the score / output contraction pair. h -> hout [3 * rows_z, C] (hout == h: in place); wscore [3][128][C] (LayerNorm gamma
// folded), wout [3][C][128] per-brush operands; ln_stats: per-(32-column chunk, row) partial sums of h (gemm_tc.h)
struct CrossOp {
    CUtensorMap mapH, mapW1, mapW2;
    int rows_z, C, T;
    const float2* ln_stats;
    const float *ln_colsum, *sbias, *obias;
    const __half* h;
    __half* hout;
    int csplit;
    float2* stats_out;
    int in_fold;  // row groups 0 and 1 read the same input rows (h has two row groups, hout three)
};
int cross_attn_setup(CrossOp* op, const __half* h, __half* hout, int rows_z, int C, int T, const __half* wscore,
                     const __half* wout, const float2* ln_stats, const float* ln_colsum, const float* sbias, const float* obias,
                     float2* stats_out, int in_fold = 0);
int cross_attn_launch(const CrossOp* op, cudaStream_t st);
const char* cross_last_error();

// [cond | texture-guidance] sample groups (Bs samples each, per_sample halfs per sample) -> [uncond = cond | cond | texture-guidance]
int launch_expand_branches(const __half* in, __half* out, int Bs, long long per_sample, cudaStream_t st);
int launch_upsample2x(const __half* x, int Nimg, int H, int W, int C, __half* out, cudaStream_t st);
// stride-2 3x3 gather: out[(n,oy,ox)][tap*C + c] = x[n, 2*oy+ky-pad_lo, 2*ox+kx-pad_lo, c] (0 outside)
int launch_im2col_s2(const __half* x, int Nimg, int H, int W, int C, int pad_lo, int Ho, int Wo, __half* out,
                     cudaStream_t st);
// stride-32 32x32 patch gather for the CLIP patch embedding: x fp32 [N,3,224,224] -> fp16 [N*49, 3072] (k = c*1024+dy*32+dx)
int launch_patchify32(const float* x, int Nimg, __half* out, cudaStream_t st);

// eps3: (3B, chw) fp32 = [uncond | cond | texture-guidance]; DDIM eta=0 epsilon-prediction step
// (stable_diffusion_pipeline.py:449-455, utilities.py:441-522)
int launch_guidance_ddim(const float* eps3, const float* latents_in, float* latents_out, int B, int chw, float cfg,
                         float tg, float alpha_t, float alpha_prev, cudaStream_t st);
// (3B, hw, 64) fp16 NHWC <- [latents(4) | mask(1) | masked latents(4) | zero pad] (stable_diffusion_pipeline.py:423-427)
int launch_pack_unet_input(const float* latents, const float* mask3, const float* masked3, int B, int hw, __half* out,
                           cudaStream_t st);
// fp32 NCHW -> fp16 NHWC with channel padding; out = x / divisor
int launch_nchw_to_nhwc_pad(const float* x, int Nimg, int C, int HW, int Cpad, float divisor, __half* out,
                            cudaStream_t st);
// latent = 0.18215 * (mean + exp(0.5*clamp(logvar,-30,20)) * noise); moments fp32 NCHW (B,8,hw); noise nullable
int launch_vae_sample(const float* moments, const float* noise, int B, int hw, float scale, float* out, cudaStream_t st);
// nearest downsample of an fp32 (B,1,R,R) mask to (B,1,R/f,R/f): out[i,j] = in[f*i, f*j] (inpaint_pipeline.py:114-115)
int launch_mask_nearest(const float* in, int B, int R, int f, float* out, cudaStream_t st);

// K19: trt_model.py:103-109 + handler.py:25-33. scratch: float[B*R*R]
int launch_canvas_preprocess(const float* canvas, const float* brush, int B, int R, int pad, float* masked_img,
                             float* mask, float* ctx_img, float* ctx_mask, float* scratch, cudaStream_t st);
// K18: model_base.py:56-58 (+ handler.py:55-56 when out_u8hwc != null)
int launch_composite(const float* canvas, const float* raw, int B, int R, float* out_f32, unsigned char* out_u8hwc,
                     cudaStream_t st);

// Folded LayerNorm (gemm_tc.h): W'[n,k] = fp16(W[n,k] * gamma[k]); colsum[n] = sum_k W'[n,k]; bias_out[n] = bias_in[n] (or 0)
// + sum_k W[n,k] * beta[k]. One warp per weight row; load time only.
int launch_ln_fold_weights(const __half* W, int N, int K, int ldw, const float* gamma, const float* beta,
                           const float* bias_in, __half* Wout, float* colsum, float* bias_out, cudaStream_t st);
// Folded cross-attention scores with a folded LayerNorm, per brush: for every row n = h * 16 + j of the (3, heads * 16, C)
// score operand: colsum[slot][n] = sum_c Wscore[slot][n][c]; bias[slot][n] = scale * K_slot[j, head h] . qbeta[head h]
// (qbeta = Wq beta, the LayerNorm shift seen through the query projection); rows of the pad tokens j >= T give 0.
int launch_cross_ln_finish(const __half* wscore, const __half* kv, const float* qbeta, int C, int heads, int T,
                           float scale, float* colsum, float* bias, cudaStream_t st);

// misc small helpers used by the runtime
int launch_add_rows_bcast(__half* x, const float* add, long long rows, int C, int period, cudaStream_t st);  // x[r,:] += add[r % period,:]
int launch_f32_to_f16(const float* x, __half* out, long long n, cudaStream_t st);
int launch_f16_to_f32(const __half* x, float* out, long long n, cudaStream_t st);
int launch_timestep_embedding(const float* timesteps, int n, int dim, __half* out, cudaStream_t st);
int launch_copy_f32(const float* src, float* dst, long long n, cudaStream_t st);
// out[r, :] = concat rows: CLIP class token + patch tokens + positional embedding
int launch_clip_embed(const __half* patch_tok, const float* cls, const float* pos, int Nimg, int C, __half* out,
                      cudaStream_t st);
int launch_gather_rows(const __half* x, int ld, const int* rows_idx, int nrows, int C, __half* out, cudaStream_t st);

}  // namespace dtp
