// Fused image-token cross-attention of a UNet transformer block (sm_100a): ONE launch for
//     h += softmax_14( LN2(h) (scale K_h Wq_h)^T ) (V_h Wo_h^T)^T + b_o
// i.e. the two skinny contractions that round 1 ran as `cross_scores` (M x C by C x 128, softmax over the 14 image tokens of
// each head in the epilogue) and `cross_out` (M x 128 by 128 x C, + bias + residual). Both operands are prepared per brush by
// the condition plan (the query / output projections are folded into the keys / values, runtime.cu), the LayerNorm in front
// is folded into the score contraction (gemm_tc.h), so a transformer block's cross-attention becomes: raw rows of h in,
// rows of h out. Replaces the fMHCA plugin + its Q / out projections of the reference graph (trt_inference/models.py:422-465,
// 520-592, enabled :1160). The probabilities never leave the SM: stage 1 accumulates the 128 scores of 128 rows in TMEM,
// the epilogue warps normalise (LayerNorm fold), soft-max each 16-column head group and write P as an fp16 K-major operand
// tile into shared memory, stage 2 multiplies it with the folded value operand chunk by chunk (<= 256 output columns).
//
// CTA = 320 threads: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-9 epilogue (two per TMEM lane quarter,
// alternate 32-column chunks). One CTA per (128-row tile, sample group z in [uncond | cond | texture-guidance]) and, at the
// narrow levels, per 256-column output chunk (each CTA then recomputes the tile's scores: cheap, and the SMs are idle anyway).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>

#include "common.cuh"
#include "gemm_tc.h"
#include "kctx.h"
#include "kernels.h"

namespace dtp {

static char g_cross_err[256] = "";
const char* cross_last_error() { return g_cross_err; }

struct CrossParams {
    int rows_z;        // rows per sample group (M of each batch entry)
    int C;             // channels
    int T;             // valid tokens per 16-column head group
    int n_chunks;      // ceil(C / 256)
    int kb1;           // C / 64: k-blocks of the score contraction
    // folded LayerNorm of the score contraction (gemm_tc.h): per-(32-column chunk, row) partial sums of h
    const float2* ln_stats;
    int ln_chunks, ln_rows;
    float ln_inv_c, ln_eps;
    const float* ln_colsum;  // [3][128]
    const float* sbias;      // [3][128]  W_score beta (LayerNorm shift seen through the folded query projection)
    const float* obias;      // [C] attn2.to_out.0.bias
    const __half* h;         // [3 * rows_z, C]: rows in (score operand and residual); in_fold: [2 * rows_z, C]
    int in_fold;             // 1: row groups 0 and 1 share their input rows (group z reads input group max(z - 1, 0)); out of place only
    __half* hout;            // result rows; == h (in place) only when csplit == 1
    int csplit;              // the output chunks of a row tile are spread over `csplit` CTAs (grid.z), each recomputing the
                             // tile's scores: parallelism for the narrow levels (6 / 3 row tiles at 16x16 / 8x8 latents)
    float2* stats_out;       // optional row statistics of the result for a following folded LayerNorm
    int stats_rows;
};

__device__ __forceinline__ void cx_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) __trap();
    }
}
__device__ __forceinline__ void cx_epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

constexpr int CX_STAGES = 4;                 // stage-1 ring: (A 128 x 64 | W_score 128 x 64) fp16 = 32 KB per stage
constexpr int CX_STAGE_BYTES = 32768;
constexpr int CX_P_BYTES = 32768;            // P: two [128 x 64] fp16 sub-tiles, 128-byte swizzled
constexpr int CX_W2_BYTES = 65536;           // one W_out chunk: two [256 x 64] sub-tiles; two chunks alias the stage-1 ring
constexpr int CX_SMEM = CX_STAGES * CX_STAGE_BYTES + CX_P_BYTES + 1024 /*colsum|sbias*/ + 1024 /*obias chunk*/ + 256 + 1024;

__global__ void __launch_bounds__(320, 1)
    cross_attn_kernel(const __grid_constant__ CUtensorMap mapH, const __grid_constant__ CUtensorMap mapW1,
                      const __grid_constant__ CUtensorMap mapW2, const __grid_constant__ CrossParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;                                   // stage 1: CX_STAGES x 32 KB; stage 2: two W_out chunks x 64 KB
    uint8_t* sP = ring + CX_STAGES * CX_STAGE_BYTES;
    float* scs = reinterpret_cast<float*>(sP + CX_P_BYTES);  // [128] colsum | [128] score bias
    float* sob = scs + 256;                                  // [256] output bias of the current chunk
    uint64_t* bars = reinterpret_cast<uint64_t*>(sob + 256);
    uint64_t* full1 = bars;                  // [CX_STAGES]
    uint64_t* empty1 = full1 + CX_STAGES;    // [CX_STAGES]
    uint64_t* s_full = empty1 + CX_STAGES;   // scores complete in TMEM
    uint64_t* p_full = s_full + 1;           // P written to shared memory (8 warp arrivals)
    uint64_t* w2_full = p_full + 1;          // [2]
    uint64_t* w2_empty = w2_full + 2;        // [2]
    uint64_t* acc_full = w2_empty + 2;       // stage-2 accumulator ready
    uint64_t* acc_empty = acc_full + 1;      // stage-2 accumulator drained (8 warp arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x, z = blockIdx.y;
    const int zi = p.in_fold ? max(z - 1, 0) : z;        // input row group
    const int row0 = zi * p.rows_z + m_tile * 128;       // first row of this tile in h (operand, residual, LayerNorm statistics)
    const int row0_out = z * p.rows_z + m_tile * 128;    // ... in hout / stats_out

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapH);
        tma_prefetch_desc(&mapW1);
        tma_prefetch_desc(&mapW2);
        for (int s = 0; s < CX_STAGES; ++s) {
            mbar_init(&full1[s], 1);
            mbar_init(&empty1[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(p_full, 8);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&w2_full[s], 1);
            mbar_init(&w2_empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 8);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);  // scores [0,128) | stage-2 accumulator [128, 384)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer ----------------
            // the per-brush score operand is not written by any kernel of the stamp: its first tiles go out before the wait
            const int pre = min(CX_STAGES, p.kb1);
            for (int s = 0; s < pre; ++s) {
                mbar_arrive_expect_tx(&full1[s], CX_STAGE_BYTES);
                tma_load_4d(ring + s * CX_STAGE_BYTES + 16384, &mapW1, &full1[s], s * 64, 0, z, 0);
            }
            pdl_wait();
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < p.kb1; ++kb) {
                if (kb >= pre) {
                    cx_wait(&empty1[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full1[stage], CX_STAGE_BYTES);
                    tma_load_4d(ring + stage * CX_STAGE_BYTES + 16384, &mapW1, &full1[stage], kb * 64, 0, z, 0);
                }
                tma_load_4d(ring + stage * CX_STAGE_BYTES, &mapH, &full1[stage], kb * 64, row0, 0, 0);
                if (++stage == CX_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            // stage 2: the ring is free once every stage-1 MMA has retired (s_full); chunk c lands in half (c & 1) of it
            cx_wait(s_full, 0);
            for (int c = blockIdx.z, ci = 0; c < p.n_chunks; c += p.csplit, ++ci) {
                const int buf = ci & 1;
                if (ci >= 2) cx_wait(&w2_empty[buf], ((ci >> 1) - 1) & 1);
                mbar_arrive_expect_tx(&w2_full[buf], CX_W2_BYTES);
                tma_load_4d(ring + buf * CX_W2_BYTES, &mapW2, &w2_full[buf], 0, c * 256, z, 0);
                tma_load_4d(ring + buf * CX_W2_BYTES + 32768, &mapW2, &w2_full[buf], 64, c * 256, z, 0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer ----------------
            const uint32_t idesc_s = umma_idesc_f16(128, 128, 0, 0);
            const uint32_t idesc_o = umma_idesc_f16(128, 256, 0, 0);
            pdl_wait();
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < p.kb1; ++kb) {
                cx_wait(&full1[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(ring + stage * CX_STAGE_BYTES);
                const uint64_t adesc = umma_desc_k_sw128(sa), bdesc = umma_desc_k_sw128(sa + 16384);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tmem_base, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc_s,
                             (kb > 0 || k > 0) ? 1u : 0u);
                umma_commit(&empty1[stage]);
                if (++stage == CX_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(s_full);
            cx_wait(p_full, 0);
            tc_fence_after();
            const uint32_t p_addr = smem_u32(sP);
            for (int c = blockIdx.z, ci = 0; c < p.n_chunks; c += p.csplit, ++ci) {
                const int buf = ci & 1;
                cx_wait(&w2_full[buf], (ci >> 1) & 1);
                if (ci > 0) cx_wait(acc_empty, (ci - 1) & 1);
                tc_fence_after();
                const uint32_t w_addr = smem_u32(ring + buf * CX_W2_BYTES);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t off = static_cast<uint32_t>((kk >> 2) * 16384 + (kk & 3) * 32);   // P: 16 KB sub-tiles
                    const uint32_t woff = static_cast<uint32_t>((kk >> 2) * 32768 + (kk & 3) * 32);  // W_out: 32 KB sub-tiles
                    umma_f16(tmem_base + 128, umma_desc_k_sw128(p_addr + off), umma_desc_k_sw128(w_addr + woff), idesc_o,
                             kk > 0 ? 1u : 0u);
                }
                umma_commit(&w2_empty[buf]);
                umma_commit(acc_full);
            }
        }
    } else {
        // ---------------- epilogue warps ----------------
        const int q = warp & 3;
        const int r = q * 32 + lane;          // row of the tile
        const int half = (warp - 2) >> 2;     // 0: even 32-column chunks, 1: odd ones
        const int et = threadIdx.x - 64;
        const bool row_ok = m_tile * 128 + r < p.rows_z;
        const long long grow = static_cast<long long>(row0) + r;
        const long long grow_out = static_cast<long long>(row0_out) + r;
        const uint32_t t_lane = static_cast<uint32_t>(q * 32) << 16;
        // per-brush vectors: independent of the predecessor kernel
        if (et < 128) {
            scs[et] = __ldg(p.ln_colsum + z * 128 + et);
            scs[128 + et] = __ldg(p.sbias + z * 128 + et);
        }
        pdl_wait();
        // LayerNorm coefficients of this row (partials written by the producer of h)
        float ln_r = 1.0f, ln_nm = 0.0f;
        if (row_ok) {
            const float2* sp = p.ln_stats + grow;
            float s = 0.0f, qq = 0.0f;
            for (int ch = 0; ch < p.ln_chunks; ch += 10) {
                float2 t[10];
#pragma unroll
                for (int u = 0; u < 10; ++u)
                    t[u] = (ch + u < p.ln_chunks) ? __ldcg(sp + static_cast<long long>(ch + u) * p.ln_rows) : make_float2(0.0f, 0.0f);
#pragma unroll
                for (int u = 0; u < 10; ++u) {
                    s += t[u].x;
                    qq += t[u].y;
                }
            }
            const float mean = s * p.ln_inv_c;
            const float var = fmaxf(qq * p.ln_inv_c - mean * mean, 0.0f);
            ln_r = rsqrtf(var + p.ln_eps);
            ln_nm = -ln_r * mean;
        }
        cx_epi_sync();  // scs visible
        // ---- stage-1 epilogue: scores -> probabilities -> P tile (this warp: chunks `half` and `half + 2` of four)
        cx_wait(s_full, 0);
        tc_fence_after();
        const uint32_t prow = smem_u32(sP + r * 128);
        const uint32_t sw16 = static_cast<uint32_t>(r & 7) << 4;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int cc = 32 * (half + 2 * it);
            uint32_t raw[32];
            tmem_ld_32x32(tmem_base + t_lane + static_cast<uint32_t>(cc), raw);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaf(__uint_as_float(raw[i]), ln_r, fmaf(ln_nm, scs[cc + i], scs[128 + cc + i]));
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                float m = -INFINITY;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (i < p.T) m = fmaxf(m, v[16 * g + i]);
                float sum = 0.0f;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float e = (i < p.T) ? __expf(v[16 * g + i] - m) : 0.0f;
                    v[16 * g + i] = e;
                    sum += e;
                }
                const float inv = __fdividef(1.0f, sum);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[16 * g + i] *= inv;
            }
            // 32 columns = 64 bytes = four 16-byte pieces of the row's 128-byte swizzle atom in sub-tile cc / 64
            const uint32_t base = prow + static_cast<uint32_t>(cc >> 6) * 16384u;
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
                const uint32_t piece = static_cast<uint32_t>(((cc & 63) >> 3) + pc);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base ^ ((piece << 4) ^ sw16)),
                             "r"(pack_half2(v[8 * pc + 0], v[8 * pc + 1])), "r"(pack_half2(v[8 * pc + 2], v[8 * pc + 3])),
                             "r"(pack_half2(v[8 * pc + 4], v[8 * pc + 5])), "r"(pack_half2(v[8 * pc + 6], v[8 * pc + 7]))
                             : "memory");
            }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        // ---- stage-2 epilogue: chunk by chunk (bias + residual, optional row statistics), in place into h
        const __half* hrow = p.h + grow * p.C;
        __half* orow = p.hout + grow_out * p.C;
        for (int c = blockIdx.z, ci = 0; c < p.n_chunks; c += p.csplit, ++ci) {
            const int n0 = c * 256;
            cx_epi_sync();  // previous chunk's readers are done with sob
            sob[et] = (n0 + et < p.C) ? __ldg(p.obias + n0 + et) : 0.0f;
            cx_epi_sync();
            cx_wait(acc_full, ci & 1);
            tc_fence_after();
#pragma unroll 1
            for (int it = 0; it < 4; ++it) {
                const int cc = 32 * (half + 2 * it);
                if (n0 + cc >= p.C) break;
                uint32_t raw[32];
                tmem_ld_32x32(tmem_base + t_lane + 128u + static_cast<uint32_t>(cc), raw);
                tmem_ld_wait();
                if (row_ok) {
                    float v[32];
                    const uint4* rp = reinterpret_cast<const uint4*>(hrow + n0 + cc);
                    uint4 rr[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) rr[i] = rp[i];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const __half2* h2 = reinterpret_cast<const __half2*>(&rr[i]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f = __half22float2(h2[j]);
                            v[8 * i + 2 * j] = __uint_as_float(raw[8 * i + 2 * j]) + sob[cc + 8 * i + 2 * j] + f.x;
                            v[8 * i + 2 * j + 1] = __uint_as_float(raw[8 * i + 2 * j + 1]) + sob[cc + 8 * i + 2 * j + 1] + f.y;
                        }
                    }
                    if (p.stats_out != nullptr) {
                        float s0 = 0.0f, q0 = 0.0f;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            s0 += v[i];
                            q0 = fmaf(v[i], v[i], q0);
                        }
                        p.stats_out[static_cast<long long>((n0 + cc) >> 5) * p.stats_rows + grow_out] = make_float2(s0, q0);
                    }
                    uint4* op = reinterpret_cast<uint4*>(orow + n0 + cc);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 u;
                        u.x = pack_half2(v[8 * i + 0], v[8 * i + 1]);
                        u.y = pack_half2(v[8 * i + 2], v[8 * i + 3]);
                        u.z = pack_half2(v[8 * i + 4], v[8 * i + 5]);
                        u.w = pack_half2(v[8 * i + 6], v[8 * i + 7]);
                        op[i] = u;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// h / hout: [3 * rows_z, C] fp16 (hout == h: in place, no chunk split); wscore: [3][128][C]; wout: [3][C][128]
int cross_attn_setup(CrossOp* op, const __half* h, __half* hout, int rows_z, int C, int T, const __half* wscore,
                     const __half* wout, const float2* ln_stats, const float* ln_colsum, const float* sbias, const float* obias,
                     float2* stats_out, int in_fold) {
    if ((C % 64) != 0 || C < 64 || T < 1 || T > 16 || rows_z < 1) {
        snprintf(g_cross_err, sizeof(g_cross_err), "cross_attn: unsupported shape rows=%d C=%d T=%d", rows_z, C, T);
        return -1;
    }
    if (in_fold && hout == h) {
        snprintf(g_cross_err, sizeof(g_cross_err), "cross_attn: shared input row groups need an out-of-place result");
        return -1;
    }
    op->in_fold = in_fold ? 1 : 0;
    const uint64_t rows = (in_fold ? 2ull : 3ull) * rows_z;
    {
        uint64_t dims[4] = {(uint64_t)C, rows, 1, 1};
        uint64_t st[3] = {(uint64_t)C * 2, (uint64_t)C * 2 * rows, (uint64_t)C * 2 * rows};
        uint32_t box[4] = {64, 128, 1, 1};
        if (make_map_4d(&op->mapH, h, dims, st, box)) return -2;
    }
    {
        uint64_t dims[4] = {(uint64_t)C, 128, 3, 1};
        uint64_t st[3] = {(uint64_t)C * 2, (uint64_t)C * 2 * 128, (uint64_t)C * 2 * 128 * 3};
        uint32_t box[4] = {64, 128, 1, 1};
        if (make_map_4d(&op->mapW1, wscore, dims, st, box)) return -2;
    }
    {
        uint64_t dims[4] = {128, (uint64_t)C, 3, 1};
        uint64_t st[3] = {256, (uint64_t)256 * C, (uint64_t)256 * C * 3};
        uint32_t box[4] = {64, 256, 1, 1};
        if (make_map_4d(&op->mapW2, wout, dims, st, box)) return -2;
    }
    op->rows_z = rows_z;
    op->C = C;
    op->T = T;
    op->ln_stats = ln_stats;
    op->ln_colsum = ln_colsum;
    op->sbias = sbias;
    op->obias = obias;
    op->h = h;
    op->hout = hout;
    op->stats_out = stats_out;
    // spread the output chunks over CTAs while the grid stays within one wave; in place only without the split (a CTA of
    // another chunk would otherwise read rows this one has already overwritten)
    const int tiles = ((rows_z + 127) / 128) * 3, n_chunks = (C + 255) / 256;
    op->csplit = 1;
    if (hout != h)
        while (op->csplit < n_chunks && tiles * (op->csplit + 1) <= 148) ++op->csplit;
    return 0;
}

int cross_attn_launch(const CrossOp* op, cudaStream_t st) {
    static bool attr_set_dev[kMaxDevices] = {false};
    bool& attr_set = attr_set_dev[kctx_device()];
    if (!attr_set) {
        if (cudaFuncSetAttribute(cross_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CX_SMEM) != cudaSuccess) {
            snprintf(g_cross_err, sizeof(g_cross_err), "cross_attn: cudaFuncSetAttribute: %s", cudaGetErrorString(cudaGetLastError()));
            return -1;
        }
        attr_set = true;
    }
    CrossParams p;
    p.rows_z = op->rows_z;
    p.C = op->C;
    p.T = op->T;
    p.n_chunks = (op->C + 255) / 256;
    p.kb1 = op->C / 64;
    p.ln_stats = op->ln_stats;
    p.ln_chunks = op->C / 32;
    p.ln_rows = (op->in_fold ? 2 : 3) * op->rows_z;
    p.in_fold = op->in_fold;
    p.ln_inv_c = 1.0f / static_cast<float>(op->C);
    p.ln_eps = 1e-5f;
    p.ln_colsum = op->ln_colsum;
    p.sbias = op->sbias;
    p.obias = op->obias;
    p.h = op->h;
    p.hout = op->hout;
    p.csplit = op->csplit;
    p.stats_out = op->stats_out;
    p.stats_rows = 3 * op->rows_z;
    const dim3 grid((op->rows_z + 127) / 128, 3, op->csplit);
    if (launch_k(cross_attn_kernel, grid, dim3(320), CX_SMEM, st, op->mapH, op->mapW1, op->mapW2, p) != cudaSuccess ||
        cudaGetLastError() != cudaSuccess) {
        snprintf(g_cross_err, sizeof(g_cross_err), "cross_attn launch failed");
        return -1;
    }
    return 0;
}

}  // namespace dtp
