// tcgen05 tensor-core contraction for the stamp path (sm_100a only).
//
//   D[M,N] = alpha * A[M,K] * W[N,K]^T (+ bias) -> activation -> (+ residual)      fp16 operands, fp32 TMEM accumulation
//
// One kernel family serves every dense contraction of the UNet / VAE / image encoder:
//   * linear layers and 1x1 convs (A = activation rows; NHWC == (pixels, C));
//   * 3x3 stride-1 pad-1 convs as implicit GEMM: the A tile of tap (ky,kx) is a shifted 4-D TMA box of the NHWC
//     activation; out-of-bounds pixels are zero-filled by TMA, which IS the conv zero padding;
//   * skip-concat inputs (torch.cat([h, skip], 1) in the diffusers up blocks) read as two sources, never
//     materialised: the channel blocks of a tap come first from source 0, then from source 1;
//   * batched products for attention (scores = Q K^T, out = P V) with (head, sample) batch coordinates carried as
//     TMA dimensions 2 and 3, V consumed as an MN-major B operand (no transpose pass).
// Replaces the TensorRT conv/GEMM tactics behind trt_inference/utilities.py:252 (Engine.infer) for the graphs
// described by trt_inference/models.py:1017-1420.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dtp {

enum GemmFlags : int {
    EPI_NONE = 0,
    EPI_GELU = 1 << 0,          // erf GELU on (acc*alpha + bias)
    EPI_QUICKGELU = 1 << 1,     // x * sigmoid(1.702 x)   (CLIP MLP)
    EPI_SILU = 1 << 2,          // x * sigmoid(x)
    EPI_GEGLU = 1 << 3,         // out[:, j] = a_j * gelu(g_j); weight rows interleaved per 32-column chunk (16 a | 16 g)
    EPI_BIAS_M = 1 << 4,        // bias indexed by output row
    EPI_OUT_F32_NCHW = 1 << 5,  // fp32 output, NCHW with hw_out pixels per image
    EPI_IMG01 = 1 << 6,         // out = clamp(x / 2 + 0.5, 0, 1)   (inpaint_pipeline.py:148)
    EPI_OUT_F32 = 1 << 7,       // fp32 row-major output
    GEMM_B_MN = 1 << 8,         // B operand is MN-major in global memory: B[K, N] row-major (V of attention)
    GEMM_HINT_CL2 = 1 << 11,    // heuristic hint only: the op will run as 2-CTA clusters sharing the weight tile by TMA multicast
    GEMM_W_BLOCKED = 1 << 10,   // weights stored as [N/64][K/64][64][64] tiles (8 KB contiguous per 64x64 block: DRAM-burst friendly)
    EPI_SOFTMAX16 = 1 << 9,     // row softmax over the first `aux` columns of every 16-column group (folded cross-attention)
};

// OR-ed into a BN argument / returned by gemm_pick_config in *BN: run the op as CTA pairs (cta_group::2, one 256-row MMA per
// pair of m-tiles, each CTA fetching half of the weight tile). Ignored for problems with a single m-tile.
constexpr int GEMM_BN_PAIR = 0x1000;

struct GemmParams {
    int M, N;          // output rows per batch entry / output columns of the contraction
    int num_kb;        // number of 64-wide k blocks (all taps)
    int mode;          // 0 = linear, 1 = conv3x3 (stride 1, pad 1)
    int H, W, Nimg;    // conv: spatial extent (input == output) and image count
    int bw, bh, bn;    // conv: pixel extents of the A box
    int tiles_x, tiles_y;
    int rows_valid;    // conv: bw*bh*bn (<= 128) rows of the tile that map to pixels
    int cblocks0;      // 64-channel blocks (per tap) that come from source 0
    int cblocks;       // 64-channel blocks per tap (source 0 + source 1); linear: == num_kb
    int sc_blocks0;    // conv3x3 + fused 1x1 shortcut: 64-channel blocks of the shortcut's first source (after the 9 taps)
    int sc_blocks;     //   ... of both shortcut sources; num_kb = 9 * cblocks + sc_blocks
    int splits;        // split-K factor; > 1 writes fp32 partials to workspace; the last-arriving CTA of a tile sums them and
                       // runs the epilogue (tile_counters), or a finalize kernel does (tile_counters == nullptr)
    int nz1, nz2;      // batch extents (heads, samples); grid.z = nz1 * nz2 * splits
    int a_batched;     // A map uses (z1,z2) as coords 2,3
    int b_batched;     // B map uses (z1,z2) as coords 2,3
    long long out_zs1, out_zs2;  // output offset (elements) per z1 / z2
    long long res_zs1, res_zs2;  // residual offset (elements) per z1 / z2
    int aux;           // EPI_SOFTMAX16: valid columns per group
    int flags;         // GemmFlags
    int hw_out;        // EPI_OUT_F32_NCHW: pixels per image
    int ldc;           // output row stride (elements)
    int ldr;           // residual row stride (elements)
    float alpha;       // multiplies acc before bias
    const float* bias;        // [N] fp32 (or [M] with EPI_BIAS_M), nullable
    const __half* residual;   // [M, ldr] fp16, nullable
    void* out;                // fp16 [M, ldc] (default) / fp32 variants
    float* workspace;         // [batch*splits, Mpad, N] fp32
    int* tile_counters;       // split-K arrival tickets, one per (batch, n-tile, m-tile); nullptr -> separate finalize launch
    long long* dbg;           // optional: per-CTA globaltimer checkpoints [ctas][8] (tuning aid), nullable
    int grid_m, grid_n, total_tiles;  // filled at launch: tile grid of the persistent scheduler
    int n_last;        // > 0: the last n-tile is ragged; its MMAs use N = n_last (multiple of 16) and its weight box comes from
                       // the op's mapBL (n_last rows; pair mode n_last / 2), so a narrow tail tile costs what it computes
    int dbg_mode;      // tuning aid (DTP_EPI_DEBUG): 1 = skip global stores, 2 = skip TMEM loads
    int* err_flag;     // mapped host flag of the owning KernelCtx: set when a cross-CTA wait gives up (kctx.h)
    // ---- LayerNorm folded into the contraction (trt_inference/models.py:304-365 fuses LN as one plugin op in front of the
    // GEMM; here it disappears into it): with W' = W diag(gamma) the normalised product is
    //     LN(x) W^T = rstd_r * (x W'^T - mean_r * colsum(W')) + W beta,
    // so the kernel multiplies the RAW rows and the epilogue applies out = rstd_r * (acc - mean_r * ln_colsum[n]) + bias[n].
    // Row statistics arrive as per-(32-column chunk, row) partial (sum, sum of squares) written by the producer of x.
    const float2* ln_stats;   // [ln_chunks][ln_rows], row index = (z2 * nz1 + z1) * M + row; nullptr = no folded LayerNorm
    const float* ln_colsum;   // [N] (+ z1 * bias_zs1): row sums of the fp16 gamma-scaled weights
    int ln_chunks, ln_rows;
    float ln_inv_c, ln_eps;
    // ---- row statistics of THIS op's fp16 output for a following folded LayerNorm: one float2 per (32-column chunk, row)
    float2* stats_out;        // [N / 32][stats_rows]; needs N % 32 == 0 and a plain fp16 row-major output
    int stats_rows;
    long long bias_zs1;       // bias / ln_colsum offset (elements) per batch coordinate z1 (folded cross-attention scores)
    // ---- nearest-2x upsample folded into the 3x3 convolution that follows it (gemm_setup_upconv2x): z1 = output parity class
    // (py, px); the class's four 2x2 taps read the HALF-resolution input at (y + sy - 1 + py, x + sx - 1 + px); weight rows of
    // class z1 start at z1 * N; output pixel = (2y + py, 2x + px) of the (2H, 2W) image
    int ws_half;   // in-kernel split-K reduction: partials travel as fp16 (set at launch; DTP_SPLITK_F16=0 keeps fp32)
    int up2;
    // ---- 3x3 stride-2 convolution without an im2col buffer (gemm_setup_conv3x3_s2): the four tensor maps mapA0..mapA3 are the
    // (row parity, column parity) views of the input at half resolution; tap (ky, kx) reads input pixel
    // (2y + ky - down_pad, 2x + kx - down_pad) = view ((ky - down_pad) & 1, (kx - down_pad) & 1) at (y + floor(.../2), x + ...)
    int down2;     // 1 = on
    int down_pad;  // 1: symmetric padding 1 (UNet Downsample2D), 0: F.pad (0,1,0,1) (VAE encoder)
};

struct GemmOp {
    CUtensorMap mapA0, mapA1, mapB;
    CUtensorMap mapA2, mapA3;  // conv3x3 with a fused 1x1 shortcut: sources of the shortcut (centre tap); == mapA0 otherwise
    CUtensorMap mapBh;  // B map with a BN/2-row box: each CTA of a CTA pair fetches its half of the weight tile
    CUtensorMap mapBL;  // B map of the ragged last n-tile (box rows = n_last, or n_last/2 in pair mode); == mapB / mapBh if none
    int cluster;        // 1 or 2
    GemmParams p;
    int BN;      // 32, 64, 128, 160, 192 or 256
    int grid_m;  // number of 128-row tiles
    int* tile_counters;  // split-K tickets of the KernelCtx current at setup time (nullptr: finalize kernel)
    int light;   // 1: two-CTAs-per-SM configuration (short K loops, BN <= 128); set by gemm_launch from the problem shape
};

// Tensor-map helper (driver entry point resolved at run time; no link-time libcuda dependency). 0 on success.
int make_map_4d(CUtensorMap* m, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                const uint32_t box[4]);

// Linear problem: A0 [M, K0] (row stride lda0) and optional A1 [M, K1] (K0 % 64 == 0 when A1 is used);
// Wt [N, K0+K1] row-major (row stride ldw). Strides in elements, multiples of 8.
// In-kernel split-K reduction: partials as fp16 (default; each partial rounded once, summed in fp32: 18.2 -> 17.0 us at
// 192 x 1280 x 11520 / 10 splits, op error 2.1e-4 -> 2.9e-4) or fp32. Process-wide; takes effect at the next launch.
void gemm_set_splitk_half(int on);
// conv3x3(nearest_upsample_2x(x)) as four parity-class 2x2 convolutions over x itself (4/9 of the multiply-adds, no upsampled
// tensor): A [Nimg, H, W, C] fp16 NHWC (C % 64 == 0), Wstack [4 * Cout, 4 * C] from launch_upconv_fold_weights, output
// [Nimg, 2H, 2W, ldc]. One launch, batch coordinate z1 = parity class.
int gemm_setup_upconv2x(GemmOp* op, const __half* A, int C, int Nimg, int H, int W, const __half* Wstack, int Cout, int BN);
// 3x3 stride-2 convolution (Downsample2D) straight from the NHWC input [Nimg, H, W, C] (H, W even, C % 64 == 0): output
// [Nimg, H/2, W/2, Cout]; Wt [Cout, 9*C] as for gemm_setup_conv3x3. pad_lo as in GemmParams::down_pad.
int gemm_setup_conv3x3_s2(GemmOp* op, const __half* A, int C, int Nimg, int H, int W, const __half* Wt, int Cout, int pad_lo,
                          int BN, int splits);
int gemm_setup_linear(GemmOp* op, const __half* A0, int lda0, int K0, const __half* A1, int lda1, int K1, int M,
                      const __half* Wt, int ldw, int N, int BN, int splits, int w_blocked = 0);
// 3x3/s1/p1 conv over NHWC activations: sources (Nimg,H,W,C0) and optional (Nimg,H,W,C1), C0,C1 % 64 == 0;
// Wt [Cout, 9*(C0+C1)] with k = (ky*3+kx)*(C0+C1) + c.
// Optional fused 1x1 convolution over other sources S0 (Nimg,H,W,CS0) [| S1 (Nimg,H,W,CS1)] of the same spatial extent
// (the conv_shortcut of a ResnetBlock2D, computed as extra k-blocks at the centre tap instead of a separate launch plus a
// residual read): Wt is then [Cout, 9*(C0+C1) + CS0 + CS1] with the shortcut weights appended along K.
int gemm_setup_conv3x3(GemmOp* op, const __half* A0, int C0, const __half* A1, int C1, int Nimg, int H, int W,
                       const __half* Wt, int Cout, int BN, int splits, int w_blocked = 0, const __half* S0 = nullptr,
                       int CS0 = 0, const __half* S1 = nullptr, int CS1 = 0);
// Batched product over (z1 in [0,nz1), z2 in [0,nz2)):  D_z[M,N] = A_z[M,K] * B_z^T
//   A_z = A + z1*a_zs1 + z2*a_zs2 (row stride lda);  K-major B_z[N,K] = B + z1*b_zs1 + z2*b_zs2 (row stride ldb),
//   or with b_mn: B_z[K,N] row-major (row stride ldb).  All strides in elements, multiples of 8.
int gemm_setup_batched(GemmOp* op, const __half* A, int lda, long long a_zs1, long long a_zs2, const __half* B, int ldb,
                       long long b_zs1, long long b_zs2, int b_mn, int M, int N, int K, int nz1, int nz2, int BN);

int gemm_launch(const GemmOp* op, cudaStream_t stream);
// kernels gemm_launch enqueues for this op: 1, or 2 when a split-K problem needs the separate finalize kernel
int gemm_num_launches(const GemmOp* op);
// allocates the split-K arrival tickets (called by the setup functions; never during graph capture)
void gemm_prepare_splitk();
size_t gemm_workspace_bytes(const GemmOp* op);
// heuristic tile / split selection for a problem with `mtiles` 128-row tiles
void gemm_pick_config(int mtiles, int N, int num_kb, int flags, int* BN, int* splits);
// true (and *BN / *splits set) when the measured table holds this exact problem shape
bool gemm_tuned_config(int mtiles, int N, int num_kb, int flags, int* BN, int* splits);
const char* gemm_last_error();
// true when linear / conv problems with >= 2 m-tiles run as 2-CTA clusters (DTP_CLUSTER=0 disables)
bool gemm_cluster_enabled();

}  // namespace dtp
