// Flash self-attention for the UNet transformer blocks on tcgen05 (sm_100a): softmax(Q K^T / sqrt(d)) V with the score
// tile living only in TMEM / registers (never in HBM). Replaces TensorRT's fused MHA for attn1 of the 16
// BasicTransformerBlocks (trt_inference/models.py:1158 keeps the native fMHA path; graph: SURVEY.md Appendix A.1).
//
// One CTA = one (sample, head, 128-query tile). Warp 0 (one lane) feeds Q once and the K / V tiles of each 128-key block
// by TMA; warp 1 allocates TMEM and (one lane) issues  S = Q K^T  (fp32 in TMEM columns [0,128)) and  O += P V  (columns
// [128, 128+dN)); warps 2..5 (one thread per query row) run the online softmax: two passes over S with tcgen05.ld, P
// written as fp16 into a 128-byte-swizzled shared tile that is the A operand of the second MMA, lazy rescaling of O
// (only when the running maximum grows by more than 2^8, exact because the same stale maximum scales P and the row sum).
// Head dims 40 / 80 / 160 are zero-padded to 64 / 128 / 192 by TMA out-of-bounds fill; only ceil(d/16) k-steps are issued.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "kctx.h"
#include "gemm_tc.h"
#include "kernels.h"

namespace dtp {

struct FlashParams {
    int seq_q, seq_kv, heads, batch, d;
    int dN;           // accumulator columns of O: d rounded up to 16
    int ksteps_qk;    // ceil(d / 16)
    float scale_log2;  // softmax scale * log2(e)
    __half* out;
    int ldo;
    long long o_bs;
};

__device__ __forceinline__ void fa_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) __trap();
    }
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// packed fp32x2 arithmetic (FFMA2 / FADD2 on sm_100): two lanes per instruction; nvcc does not form these on its own
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
    unsigned long long av, bv, cv, dv;
    asm("mov.b64 %0, {%1, %2};" : "=l"(av) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cv) : "f"(c));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(dv) : "l"(av), "l"(bv), "l"(cv));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(dv));
}
__device__ __forceinline__ void fadd2(float& acc0, float& acc1, float a0, float a1) {
    unsigned long long av, cv, dv;
    asm("mov.b64 %0, {%1, %2};" : "=l"(av) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(cv) : "f"(acc0), "f"(acc1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(dv) : "l"(av), "l"(cv));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc0), "=f"(acc1) : "l"(dv));
}
// three-input maximum (FMNMX3 on sm_100): halves the instruction count of the row-maximum pass
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ void softmax_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// DCH: 64-wide chunks of the (padded) head dim; KV_STAGES: ring depth of the K and V tiles
template <int DCH, int KV_STAGES>
__global__ void __launch_bounds__(192, DCH == 1 ? 2 : 1)
    flash_attn_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                      const __grid_constant__ CUtensorMap mapV, const __grid_constant__ FlashParams p) {
    constexpr int TILE_BYTES = DCH * 16384;  // one [128 rows x DCH*64] fp16 operand tile
    constexpr int P_BYTES = 2 * 16384;       // [128 q x 128 kv] fp16
    constexpr int TM_COLS = (128 + DCH * 64 <= 256) ? 256 : 512;
    constexpr uint32_t O_COL = 128;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + TILE_BYTES;
    uint8_t* sV = sK + KV_STAGES * TILE_BYTES;
    uint8_t* sP = sV + KV_STAGES * TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = k_full + KV_STAGES;
    uint64_t* v_full = k_empty + KV_STAGES;
    uint64_t* v_empty = v_full + KV_STAGES;
    uint64_t* s_full = v_empty + KV_STAGES;
    uint64_t* s_empty = s_full + 1;
    uint64_t* p_full = s_empty + 1;
    uint64_t* o_done = p_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, head = blockIdx.y, b = blockIdx.z;
    const int nblk = (p.seq_kv + 127) / 128;

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapQ);
        tma_prefetch_desc(&mapK);
        tma_prefetch_desc(&mapV);
        mbar_init(q_full, 1);
        for (int s = 0; s < KV_STAGES; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_empty, 4);
        mbar_init(p_full, 4);
        mbar_init(o_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // q / k / v are produced by the preceding projection kernel

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------ TMA producer ------------------------------
            mbar_arrive_expect_tx(q_full, TILE_BYTES);
#pragma unroll
            for (int c = 0; c < DCH; ++c) tma_load_4d(sQ + c * 16384, &mapQ, q_full, c * 64, q0, head, b);
            int st = 0;
            uint32_t ph = 0;
            for (int j = 0; j < nblk; ++j) {
                fa_wait(&k_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
#pragma unroll
                for (int c = 0; c < DCH; ++c)
                    tma_load_4d(sK + st * TILE_BYTES + c * 16384, &mapK, &k_full[st], c * 64, j * 128, head, b);
                fa_wait(&v_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
#pragma unroll
                for (int c = 0; c < DCH; ++c)
                    tma_load_4d(sV + st * TILE_BYTES + c * 16384, &mapV, &v_full[st], c * 64, j * 128, head, b);
                if (++st == KV_STAGES) {
                    st = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------ MMA issuer ------------------------------
            const uint32_t idesc_s = umma_idesc_f16(128, 128, 0, 0);
            const uint32_t idesc_o = umma_idesc_f16(128, p.dN, 0, 1);
            const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
            fa_wait(q_full, 0);
            int st = 0;
            uint32_t ph = 0;
            for (int j = 0; j < nblk; ++j) {
                // S = Q K_j^T
                fa_wait(&k_full[st], ph);
                fa_wait(s_empty, (j & 1) ^ 1);  // softmax has consumed the previous S
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sK + st * TILE_BYTES);
                for (int kk = 0; kk < p.ksteps_qk; ++kk) {
                    const uint32_t off = static_cast<uint32_t>((kk >> 2) * 16384 + (kk & 3) * 32);
                    umma_f16(tmem_base, umma_desc_k_sw128(q_addr + off), umma_desc_k_sw128(k_addr + off), idesc_s,
                             kk > 0 ? 1u : 0u);
                }
                umma_commit(&k_empty[st]);
                umma_commit(s_full);
                // O += P_j V_j
                fa_wait(&v_full[st], ph);
                fa_wait(p_full, j & 1);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(sV + st * TILE_BYTES);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t poff = static_cast<uint32_t>((kk >> 2) * 16384 + (kk & 3) * 32);
                    umma_f16(tmem_base + O_COL, umma_desc_k_sw128(p_addr + poff),
                             umma_desc_mn_sw128(v_addr + static_cast<uint32_t>(kk) * 2048u, 16384), idesc_o,
                             (j > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&v_empty[st]);
                umma_commit(o_done);
                if (++st == KV_STAGES) {
                    st = 0;
                    ph ^= 1;
                }
            }
        }
    } else {
        // ------------------------------ softmax / correction / epilogue: one thread per query row ------------------------------
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t t_lane = static_cast<uint32_t>(q * 32) << 16;
        float m_used = -INFINITY, l = 0.0f;  // m_used in the log2 domain (score * scale * log2 e)
        uint8_t* prow = sP + r * 128;
        const int sw = r & 7;
        const float sc = p.scale_log2;
        for (int j = 0; j < nblk; ++j) {
            fa_wait(s_full, j & 1);
            tc_fence_after();
            const int kv_valid = min(128, p.seq_kv - j * 128);
            // pass 1: block maximum of the raw scores (four 32-column TMEM reads in flight together)
            float bm = -INFINITY;
            {
                uint32_t a[32], b[32];
                tmem_ld_32x32(tmem_base + t_lane + 0, a);
                tmem_ld_32x32(tmem_base + t_lane + 32, b);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i < kv_valid) bm = fmaxf(bm, __uint_as_float(a[i]));
                    if (32 + i < kv_valid) bm = fmaxf(bm, __uint_as_float(b[i]));
                }
                tmem_ld_32x32(tmem_base + t_lane + 64, a);
                tmem_ld_32x32(tmem_base + t_lane + 96, b);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (64 + i < kv_valid) bm = fmaxf(bm, __uint_as_float(a[i]));
                    if (96 + i < kv_valid) bm = fmaxf(bm, __uint_as_float(b[i]));
                }
            }
            bm *= sc;  // scale > 0: max commutes with the scaling
            // P (smem) and O (TMEM) are free once the previous block's second MMA has retired
            if (j > 0) fa_wait(o_done, (j - 1) & 1);
            tc_fence_after();
            bool need = false;
            float factor = 1.0f;
            if (j == 0) {
                m_used = bm;
            } else if (bm > m_used + 8.0f) {
                need = true;
                factor = ex2_approx(m_used - bm);
                m_used = bm;
            }
            if (__any_sync(0xffffffffu, need)) {
                l *= factor;
                for (int c = 0; c < p.dN; c += 32) {
                    uint32_t o[32];
                    tmem_ld_32x32(tmem_base + t_lane + O_COL + c, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
                    tmem_st_32x32(tmem_base + t_lane + O_COL + c, o);
                }
                tmem_st_wait();
            }
            // pass 2: P = exp2(s * scale - m): one FFMA + one MUFU per element; fp16 P into the swizzled A tile
            const float neg_m = -m_used;
            float l0 = 0.0f, l1 = 0.0f;
#pragma unroll 1
            for (int hc = 0; hc < 128; hc += 64) {
                uint32_t raw[64];
                tmem_ld_32x32(tmem_base + t_lane + hc, *reinterpret_cast<uint32_t(*)[32]>(&raw[0]));
                tmem_ld_32x32(tmem_base + t_lane + hc + 32, *reinterpret_cast<uint32_t(*)[32]>(&raw[32]));
                tmem_ld_wait();
                if (hc == 64) {
                    // the score tile is fully in registers: release S to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_empty);
                }
                uint8_t* sub = prow + (hc >> 6) * 16384;
#pragma unroll
                for (int c = 0; c < 64; c += 8) {
                    float e[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        e[i] = ex2_approx(fmaf(__uint_as_float(raw[c + i]), sc, neg_m));
                        if (hc + c + i >= kv_valid) e[i] = 0.0f;
                    }
                    l0 += (e[0] + e[1]) + (e[2] + e[3]);
                    l1 += (e[4] + e[5]) + (e[6] + e[7]);
                    __half2 h0 = __floats2half2_rn(e[0], e[1]), h1 = __floats2half2_rn(e[2], e[3]);
                    __half2 h2 = __floats2half2_rn(e[4], e[5]), h3 = __floats2half2_rn(e[6], e[7]);
                    uint4 u;
                    u.x = *reinterpret_cast<uint32_t*>(&h0);
                    u.y = *reinterpret_cast<uint32_t*>(&h1);
                    u.z = *reinterpret_cast<uint32_t*>(&h2);
                    u.w = *reinterpret_cast<uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(sub + (((c >> 3) ^ sw) << 4)) = u;
                }
            }
            l += l0 + l1;
            // P visible to the async proxy; O rescaled
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // epilogue: O / l -> fp16
        fa_wait(o_done, (nblk - 1) & 1);
        tc_fence_after();
        const float inv = 1.0f / l;
        const int row = q0 + r;
        __half* orow = p.out + b * p.o_bs + static_cast<long long>(row) * p.ldo + head * p.d;
        for (int c = 0; c < p.dN; c += 32) {
            uint32_t raw[32];
            tmem_ld_32x32(tmem_base + t_lane + O_COL + c, raw);
            tmem_ld_wait();
            if (row < p.seq_q) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    if (c + i + 8 <= p.d) {
                        __half2 h0 = __floats2half2_rn(__uint_as_float(raw[i]) * inv, __uint_as_float(raw[i + 1]) * inv);
                        __half2 h1 = __floats2half2_rn(__uint_as_float(raw[i + 2]) * inv, __uint_as_float(raw[i + 3]) * inv);
                        __half2 h2 = __floats2half2_rn(__uint_as_float(raw[i + 4]) * inv, __uint_as_float(raw[i + 5]) * inv);
                        __half2 h3 = __floats2half2_rn(__uint_as_float(raw[i + 6]) * inv, __uint_as_float(raw[i + 7]) * inv);
                        uint4 u;
                        u.x = *reinterpret_cast<uint32_t*>(&h0);
                        u.y = *reinterpret_cast<uint32_t*>(&h1);
                        u.z = *reinterpret_cast<uint32_t*>(&h2);
                        u.w = *reinterpret_cast<uint32_t*>(&h3);
                        *reinterpret_cast<uint4*>(orow + c + i) = u;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
// Two-warpgroup variant (256 queries per CTA) for head dims <= 128: K / V tiles are fetched once for two query tiles, and
// while one warpgroup runs its softmax the tensor pipe works on the other one's S = Q K^T / O += P V, so the MUFU-bound
// softmax and the MMAs overlap inside one CTA (TMEM: S0 | S1 | O0 | O1 = 512 columns).
// Warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: softmax of query tile 0, warps 6-9: softmax of query tile 1.
// ------------------------------------------------------------------------------------------------------------
template <int DCH, int KS>
__global__ void __launch_bounds__(384, 1)
    flash_attn2_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                       const __grid_constant__ CUtensorMap mapV, const __grid_constant__ FlashParams p) {
    constexpr int TILE_BYTES = DCH * 16384;
    constexpr int P_BYTES = 2 * 16384;
    constexpr uint32_t O_COL0 = 256, O_STRIDE = DCH * 64;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                       // [2][TILE]
    uint8_t* sK = sQ + 2 * TILE_BYTES;        // [KS][TILE]
    uint8_t* sV = sK + KS * TILE_BYTES;       // [KS][TILE]
    uint8_t* sP = sV + KS * TILE_BYTES;       // [2][P]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * P_BYTES);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = k_full + KS;
    uint64_t* v_full = k_empty + KS;
    uint64_t* v_empty = v_full + KS;
    uint64_t* s_full = v_empty + KS;   // [2]
    uint64_t* s_empty = s_full + 2;    // [2]
    uint64_t* p_full = s_empty + 2;    // [2]
    uint64_t* o_done = p_full + 2;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 256, head = blockIdx.y, b = blockIdx.z;
    const int nblk = (p.seq_kv + 127) / 128;

    pdl_launch_dependents();
    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&mapQ);
        tma_prefetch_desc(&mapK);
        tma_prefetch_desc(&mapV);
        mbar_init(q_full, 1);
        for (int s = 0; s < KS; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 1);
        }
        for (int w = 0; w < 2; ++w) {
            mbar_init(&s_full[w], 1);
            mbar_init(&s_empty[w], 4);
            mbar_init(&p_full[w], 4);
            mbar_init(&o_done[w], 1);
        }
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    // warps 0-3 / 4-7: softmax warpgroups of query tile 0 / 1 (224 registers each: the 128 scores of a row stay in
    // registers); warps 8-11: control warpgroup (8 = TMA producer, 9 = MMA issuer) shrunk to 40 registers
    if (warp >= 8) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
      if (warp == 8) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, 2 * TILE_BYTES);
#pragma unroll
            for (int w = 0; w < 2; ++w)
#pragma unroll
                for (int c = 0; c < DCH; ++c)
                    tma_load_4d(sQ + w * TILE_BYTES + c * 16384, &mapQ, q_full, c * 64, q0 + w * 128, head, b);
            int st = 0;
            uint32_t ph = 0;
            for (int j = 0; j < nblk; ++j) {
                fa_wait(&k_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
#pragma unroll
                for (int c = 0; c < DCH; ++c)
                    tma_load_4d(sK + st * TILE_BYTES + c * 16384, &mapK, &k_full[st], c * 64, j * 128, head, b);
                fa_wait(&v_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
#pragma unroll
                for (int c = 0; c < DCH; ++c)
                    tma_load_4d(sV + st * TILE_BYTES + c * 16384, &mapV, &v_full[st], c * 64, j * 128, head, b);
                if (++st == KS) {
                    st = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_f16(128, 128, 0, 0);
            const uint32_t idesc_o = umma_idesc_f16(128, p.dN, 0, 1);
            const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
            auto issue_s = [&](int w, uint32_t k_addr) {
                for (int kk = 0; kk < p.ksteps_qk; ++kk) {
                    const uint32_t off = static_cast<uint32_t>((kk >> 2) * 16384 + (kk & 3) * 32);
                    umma_f16(tmem_base + w * 128, umma_desc_k_sw128(q_addr + w * TILE_BYTES + off),
                             umma_desc_k_sw128(k_addr + off), idesc_s, kk > 0 ? 1u : 0u);
                }
                umma_commit(&s_full[w]);
            };
            auto issue_pv = [&](int w, uint32_t v_addr, int j) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t poff = static_cast<uint32_t>((kk >> 2) * 16384 + (kk & 3) * 32);
                    umma_f16(tmem_base + O_COL0 + w * O_STRIDE, umma_desc_k_sw128(p_addr + w * P_BYTES + poff),
                             umma_desc_mn_sw128(v_addr + static_cast<uint32_t>(kk) * 2048u, 16384), idesc_o,
                             (j > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&o_done[w]);
            };
            fa_wait(q_full, 0);
            int kst = 0;
            uint32_t kph = 0;  // K ring position of block j + 1 (the one whose scores are issued inside iteration j)
            int vst = 0;
            uint32_t vph = 0;
            // prologue: scores of block 0 for both query tiles
            fa_wait(&k_full[0], 0);
            tc_fence_after();
            issue_s(0, smem_u32(sK));
            issue_s(1, smem_u32(sK));
            umma_commit(&k_empty[0]);
            if (++kst == KS) {
                kst = 0;
                kph ^= 1;
            }
            for (int j = 0; j < nblk; ++j) {
                // The scores of block j + 1 only need the S columns back (s_empty: the softmax warps hold block j in
                // registers), not P_j: issue them FIRST, so they are ready long before the warpgroup finishes its
                // exponentials, and only then the two P V products of block j.
                if (j + 1 < nblk) {
                    fa_wait(&k_full[kst], kph);
                    const uint32_t k_addr = smem_u32(sK + kst * TILE_BYTES);
                    fa_wait(&s_empty[0], j & 1);
                    tc_fence_after();
                    issue_s(0, k_addr);
                    fa_wait(&s_empty[1], j & 1);
                    tc_fence_after();
                    issue_s(1, k_addr);
                    umma_commit(&k_empty[kst]);
                    if (++kst == KS) {
                        kst = 0;
                        kph ^= 1;
                    }
                }
                fa_wait(&v_full[vst], vph);
                const uint32_t v_addr = smem_u32(sV + vst * TILE_BYTES);
                fa_wait(&p_full[0], j & 1);
                tc_fence_after();
                issue_pv(0, v_addr, j);
                fa_wait(&p_full[1], j & 1);
                tc_fence_after();
                issue_pv(1, v_addr, j);
                umma_commit(&v_empty[vst]);
                if (++vst == KS) {
                    vst = 0;
                    vph ^= 1;
                }
            }
        }
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const int w = warp >> 2;  // query tile / warpgroup
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t t_lane = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t s_col = tmem_base + t_lane + w * 128;
        const uint32_t o_col = tmem_base + t_lane + O_COL0 + w * O_STRIDE;
        float m_used = -INFINITY, l = 0.0f;
        uint8_t* prow = sP + w * P_BYTES + r * 128;
        const int sw = r & 7;
        const uint32_t prow_s = smem_u32(prow);  // 128-byte aligned: XOR with a value < 128 only touches the piece bits
        const uint32_t sw16 = static_cast<uint32_t>(sw) << 4;
        const float sc = p.scale_log2;
        for (int j = 0; j < nblk; ++j) {
            fa_wait(&s_full[w], j & 1);
            tc_fence_after();
            const int kv_valid = min(128, p.seq_kv - j * 128);
            // TMEM -> register bandwidth (~64 B/clk/SM) is the scarce resource of this loop: read the 128 scores of the
            // row exactly once, keep them in registers for both the maximum and the exponentials, release S at once.
            uint32_t raw[128];
#pragma unroll
            for (int c = 0; c < 128; c += 32) tmem_ld_32x32(s_col + c, *reinterpret_cast<uint32_t(*)[32]>(&raw[c]));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[w]);
            float bm = -INFINITY;
            if (kv_valid == 128) {
                float b0 = -INFINITY, b1 = -INFINITY, b2 = -INFINITY, b3 = -INFINITY;
#pragma unroll
                for (int i = 0; i < 128; i += 8) {
                    b0 = max3(b0, __uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
                    b1 = max3(b1, __uint_as_float(raw[i + 2]), __uint_as_float(raw[i + 3]));
                    b2 = max3(b2, __uint_as_float(raw[i + 4]), __uint_as_float(raw[i + 5]));
                    b3 = max3(b3, __uint_as_float(raw[i + 6]), __uint_as_float(raw[i + 7]));
                }
                bm = max3(max3(b0, b1, b2), b3, -INFINITY);
            } else {
#pragma unroll
                for (int i = 0; i < 128; ++i) {
                    if (i >= kv_valid) raw[i] = 0xff800000u;  // -inf: exp2 -> 0
                    bm = fmaxf(bm, __uint_as_float(raw[i]));
                }
            }
            bm *= sc;
            bool need = false;
            float factor = 1.0f;
            if (j == 0) {
                m_used = bm;
            } else if (bm > m_used + 8.0f) {
                need = true;
                factor = ex2_approx(m_used - bm);
                m_used = bm;
            }
            // exponentials first, packed to fp16 in registers: they depend on neither O nor the P buffer, so the latency of
            // the previous block's P V product (o_done) hides behind the MUFU work instead of stalling the warp in front of it
            const float neg_m = -m_used;
            float l0 = 0.0f, l1 = 0.0f;
            uint32_t pk[64];
#pragma unroll
            for (int c = 0; c < 128; c += 8) {
                float e[8];
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    float x0, x1;
                    ffma2(x0, x1, __uint_as_float(raw[c + i]), __uint_as_float(raw[c + i + 1]), sc, neg_m);
                    e[i] = ex2_approx(x0);
                    e[i + 1] = ex2_approx(x1);
                }
                fadd2(l0, l1, e[0], e[1]);
                fadd2(l0, l1, e[2], e[3]);
                fadd2(l0, l1, e[4], e[5]);
                fadd2(l0, l1, e[6], e[7]);
                pk[(c >> 1) + 0] = pack_half2(e[0], e[1]);
                pk[(c >> 1) + 1] = pack_half2(e[2], e[3]);
                pk[(c >> 1) + 2] = pack_half2(e[4], e[5]);
                pk[(c >> 1) + 3] = pack_half2(e[6], e[7]);
            }
            if (j > 0) fa_wait(&o_done[w], (j - 1) & 1);  // P buffer free again, O_{j-1} final
            tc_fence_after();
            if (__any_sync(0xffffffffu, need)) {
                l *= factor;
                for (int c = 0; c < p.dN; c += 32) {
                    uint32_t o[32];
                    tmem_ld_32x32(o_col + c, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
                    tmem_st_32x32(o_col + c, o);
                }
                tmem_st_wait();
            }
            l += l0 + l1;
            // explicit 32-bit shared-window addresses: (piece ^ sw) << 4 == (piece << 4) ^ (sw << 4), one LOP3 per store
#pragma unroll
            for (int c = 0; c < 128; c += 8)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"((prow_s + (c >> 6) * 16384) ^ ((((c & 63) >> 3) << 4) ^ sw16)),
                             "r"(pk[(c >> 1) + 0]), "r"(pk[(c >> 1) + 1]), "r"(pk[(c >> 1) + 2]), "r"(pk[(c >> 1) + 3])
                             : "memory");
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[w]);
        }
        fa_wait(&o_done[w], (nblk - 1) & 1);
        tc_fence_after();
        const float inv = 1.0f / l;
        const int row = q0 + w * 128 + r;
        __half* orow = p.out + b * p.o_bs + static_cast<long long>(row) * p.ldo + head * p.d;
        for (int c = 0; c < p.dN; c += 32) {
            uint32_t raw[32];
            tmem_ld_32x32(o_col + c, raw);
            tmem_ld_wait();
            if (row < p.seq_q) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    if (c + i + 8 <= p.d) {
                        __half2 h0 = __floats2half2_rn(__uint_as_float(raw[i]) * inv, __uint_as_float(raw[i + 1]) * inv);
                        __half2 h1 = __floats2half2_rn(__uint_as_float(raw[i + 2]) * inv, __uint_as_float(raw[i + 3]) * inv);
                        __half2 h2 = __floats2half2_rn(__uint_as_float(raw[i + 4]) * inv, __uint_as_float(raw[i + 5]) * inv);
                        __half2 h3 = __floats2half2_rn(__uint_as_float(raw[i + 6]) * inv, __uint_as_float(raw[i + 7]) * inv);
                        uint4 u;
                        u.x = *reinterpret_cast<uint32_t*>(&h0);
                        u.y = *reinterpret_cast<uint32_t*>(&h1);
                        u.z = *reinterpret_cast<uint32_t*>(&h2);
                        u.w = *reinterpret_cast<uint32_t*>(&h3);
                        *reinterpret_cast<uint4*>(orow + c + i) = u;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, 512);
}

template <int DCH, int KS>
static int launch_flash2(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const FlashParams& p,
                         cudaStream_t st) {
    constexpr int SMEM = DCH * 16384 * (2 + 2 * KS) + 4 * 16384 + (9 + 4 * KS) * 8 + 16 + 1024;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    static bool attr_set_dev[kMaxDevices] = {false};
    bool& attr_set = attr_set_dev[kctx_device()];
    if (!attr_set) {
        if (cudaFuncSetAttribute(flash_attn2_kernel<DCH, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) !=
            cudaSuccess)
            return -1;
        attr_set = true;
    }
    dim3 grid((p.seq_q + 255) / 256, p.heads, p.batch);
    return launch_k(flash_attn2_kernel<DCH, KS>, grid, dim3(384), SMEM, st, mq, mk, mv, p) == cudaSuccess ? 0 : -1;
}

// ------------------------------------------------------------------------------------------------------------
// Variant 3: one softmax warpgroup (128 queries per CTA) with TWO score buffers in TMEM: S_{j+1} = Q K_{j+1}^T is issued
// while the warpgroup is still exponentiating block j, so the tensor pipe never sits on the softmax critical path; the
// 128 scores of a row are read from TMEM once and stay in registers (setmaxnreg).  TMEM: S[0] | S[1] | O (<= 160 columns).
// Warps 0-3: softmax, warp 4: TMA producer, warp 5: MMA issuer, warps 6-7 idle (complete the control warpgroup).
// ------------------------------------------------------------------------------------------------------------
template <int DCH, int KS>
__global__ void __launch_bounds__(256, 1)
    flash_attn3_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                       const __grid_constant__ CUtensorMap mapV, const __grid_constant__ FlashParams p) {
    constexpr int TILE_BYTES = DCH * 16384;
    constexpr int P_BYTES = 2 * 16384;
    constexpr uint32_t O_COL = 256;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + TILE_BYTES;
    uint8_t* sV = sK + KS * TILE_BYTES;
    uint8_t* sP = sV + KS * TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = k_full + KS;
    uint64_t* v_full = k_empty + KS;
    uint64_t* v_empty = v_full + KS;
    uint64_t* s_full = v_empty + KS;  // [2]
    uint64_t* s_free = s_full + 2;    // [2]
    uint64_t* p_full = s_free + 2;
    uint64_t* o_done = p_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, head = blockIdx.y, b = blockIdx.z;
    const int nblk = (p.seq_kv + 127) / 128;

    pdl_launch_dependents();
    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&mapQ);
        tma_prefetch_desc(&mapK);
        tma_prefetch_desc(&mapV);
        mbar_init(q_full, 1);
        for (int s = 0; s < KS; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 4);
        }
        mbar_init(p_full, 4);
        mbar_init(o_done, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp >= 4) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
      if (warp == 4) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, TILE_BYTES);
#pragma unroll
            for (int c = 0; c < DCH; ++c) tma_load_4d(sQ + c * 16384, &mapQ, q_full, c * 64, q0, head, b);
            int st = 0;
            uint32_t ph = 0;
            for (int j = 0; j < nblk; ++j) {
                fa_wait(&k_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
#pragma unroll
                for (int c = 0; c < DCH; ++c)
                    tma_load_4d(sK + st * TILE_BYTES + c * 16384, &mapK, &k_full[st], c * 64, j * 128, head, b);
                fa_wait(&v_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
#pragma unroll
                for (int c = 0; c < DCH; ++c)
                    tma_load_4d(sV + st * TILE_BYTES + c * 16384, &mapV, &v_full[st], c * 64, j * 128, head, b);
                if (++st == KS) {
                    st = 0;
                    ph ^= 1;
                }
            }
        }
      } else if (warp == 5) {
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_f16(128, 128, 0, 0);
            const uint32_t idesc_o = umma_idesc_f16(128, p.dN, 0, 1);
            const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
            fa_wait(q_full, 0);
            int kst = 0, vst = 0;
            uint32_t kph = 0, vph = 0;
            // scores of block jj into buffer jj & 1 (the softmax must have copied block jj - 2 out of that buffer)
            auto issue_s = [&](int jj) {
                fa_wait(&k_full[kst], kph);
                if (jj >= 2) fa_wait(&s_free[jj & 1], ((jj >> 1) - 1) & 1);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sK + kst * TILE_BYTES);
                for (int kk = 0; kk < p.ksteps_qk; ++kk) {
                    const uint32_t off = static_cast<uint32_t>((kk >> 2) * 16384 + (kk & 3) * 32);
                    umma_f16(tmem_base + (jj & 1) * 128, umma_desc_k_sw128(q_addr + off), umma_desc_k_sw128(k_addr + off),
                             idesc_s, kk > 0 ? 1u : 0u);
                }
                umma_commit(&k_empty[kst]);
                umma_commit(&s_full[jj & 1]);
                if (++kst == KS) {
                    kst = 0;
                    kph ^= 1;
                }
            };
            issue_s(0);
            for (int j = 0; j < nblk; ++j) {
                if (j + 1 < nblk) issue_s(j + 1);  // overlaps the softmax of block j
                fa_wait(&v_full[vst], vph);
                fa_wait(p_full, j & 1);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(sV + vst * TILE_BYTES);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t poff = static_cast<uint32_t>((kk >> 2) * 16384 + (kk & 3) * 32);
                    umma_f16(tmem_base + O_COL, umma_desc_k_sw128(p_addr + poff),
                             umma_desc_mn_sw128(v_addr + static_cast<uint32_t>(kk) * 2048u, 16384), idesc_o,
                             (j > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&v_empty[vst]);
                umma_commit(o_done);
                if (++vst == KS) {
                    vst = 0;
                    vph ^= 1;
                }
            }
        }
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t t_lane = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t o_col = tmem_base + t_lane + O_COL;
        float m_used = -INFINITY, l = 0.0f;
        uint8_t* prow = sP + r * 128;
        const int sw = r & 7;
        const float sc = p.scale_log2;
        for (int j = 0; j < nblk; ++j) {
            fa_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            const int kv_valid = min(128, p.seq_kv - j * 128);
            const uint32_t s_col = tmem_base + t_lane + (j & 1) * 128;
            uint32_t raw[128];
#pragma unroll
            for (int c = 0; c < 128; c += 32) tmem_ld_32x32(s_col + c, *reinterpret_cast<uint32_t(*)[32]>(&raw[c]));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[j & 1]);
            float bm = -INFINITY;
            if (kv_valid == 128) {
                float b0 = -INFINITY, b1 = -INFINITY, b2 = -INFINITY, b3 = -INFINITY;
#pragma unroll
                for (int i = 0; i < 128; i += 8) {
                    b0 = max3(b0, __uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
                    b1 = max3(b1, __uint_as_float(raw[i + 2]), __uint_as_float(raw[i + 3]));
                    b2 = max3(b2, __uint_as_float(raw[i + 4]), __uint_as_float(raw[i + 5]));
                    b3 = max3(b3, __uint_as_float(raw[i + 6]), __uint_as_float(raw[i + 7]));
                }
                bm = max3(max3(b0, b1, b2), b3, -INFINITY);
            } else {
#pragma unroll
                for (int i = 0; i < 128; ++i) {
                    if (i >= kv_valid) raw[i] = 0xff800000u;
                    bm = fmaxf(bm, __uint_as_float(raw[i]));
                }
            }
            bm *= sc;
            if (j > 0) fa_wait(o_done, (j - 1) & 1);
            tc_fence_after();
            bool need = false;
            float factor = 1.0f;
            if (j == 0) {
                m_used = bm;
            } else if (bm > m_used + 8.0f) {
                need = true;
                factor = ex2_approx(m_used - bm);
                m_used = bm;
            }
            if (__any_sync(0xffffffffu, need)) {
                l *= factor;
                for (int c = 0; c < p.dN; c += 32) {
                    uint32_t o[32];
                    tmem_ld_32x32(o_col + c, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
                    tmem_st_32x32(o_col + c, o);
                }
                tmem_st_wait();
            }
            const float neg_m = -m_used;
            float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
            for (int c = 0; c < 128; c += 8) {
                float e[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) e[i] = ex2_approx(fmaf(__uint_as_float(raw[c + i]), sc, neg_m));
                l0 += (e[0] + e[1]) + (e[2] + e[3]);
                l1 += (e[4] + e[5]) + (e[6] + e[7]);
                __half2 h0 = __floats2half2_rn(e[0], e[1]), h1 = __floats2half2_rn(e[2], e[3]);
                __half2 h2 = __floats2half2_rn(e[4], e[5]), h3 = __floats2half2_rn(e[6], e[7]);
                uint4 u;
                u.x = *reinterpret_cast<uint32_t*>(&h0);
                u.y = *reinterpret_cast<uint32_t*>(&h1);
                u.z = *reinterpret_cast<uint32_t*>(&h2);
                u.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(prow + (c >> 6) * 16384 + ((((c & 63) >> 3) ^ sw) << 4)) = u;
            }
            l += l0 + l1;
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        fa_wait(o_done, (nblk - 1) & 1);
        tc_fence_after();
        const float inv = 1.0f / l;
        const int row = q0 + r;
        __half* orow = p.out + b * p.o_bs + static_cast<long long>(row) * p.ldo + head * p.d;
        for (int c = 0; c < p.dN; c += 32) {
            uint32_t raw[32];
            tmem_ld_32x32(o_col + c, raw);
            tmem_ld_wait();
            if (row < p.seq_q) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    if (c + i + 8 <= p.d) {
                        __half2 h0 = __floats2half2_rn(__uint_as_float(raw[i]) * inv, __uint_as_float(raw[i + 1]) * inv);
                        __half2 h1 = __floats2half2_rn(__uint_as_float(raw[i + 2]) * inv, __uint_as_float(raw[i + 3]) * inv);
                        __half2 h2 = __floats2half2_rn(__uint_as_float(raw[i + 4]) * inv, __uint_as_float(raw[i + 5]) * inv);
                        __half2 h3 = __floats2half2_rn(__uint_as_float(raw[i + 6]) * inv, __uint_as_float(raw[i + 7]) * inv);
                        uint4 u;
                        u.x = *reinterpret_cast<uint32_t*>(&h0);
                        u.y = *reinterpret_cast<uint32_t*>(&h1);
                        u.z = *reinterpret_cast<uint32_t*>(&h2);
                        u.w = *reinterpret_cast<uint32_t*>(&h3);
                        *reinterpret_cast<uint4*>(orow + c + i) = u;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 512);
}

template <int DCH, int KS>
static int launch_flash3(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const FlashParams& p,
                         cudaStream_t st) {
    constexpr int SMEM = DCH * 16384 * (1 + 2 * KS) + 2 * 16384 + (8 + 4 * KS) * 8 + 16 + 1024;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    static bool attr_set_dev[kMaxDevices] = {false};
    bool& attr_set = attr_set_dev[kctx_device()];
    if (!attr_set) {
        if (cudaFuncSetAttribute(flash_attn3_kernel<DCH, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) !=
            cudaSuccess)
            return -1;
        attr_set = true;
    }
    dim3 grid((p.seq_q + 127) / 128, p.heads, p.batch);
    return launch_k(flash_attn3_kernel<DCH, KS>, grid, dim3(256), SMEM, st, mq, mk, mv, p) == cudaSuccess ? 0 : -1;
}

template <int DCH, int KV_STAGES>
static int launch_flash(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const FlashParams& p,
                        cudaStream_t st) {
    constexpr int SMEM = DCH * 16384 * (1 + 2 * KV_STAGES) + 2 * 16384 + (5 + 4 * KV_STAGES) * 8 + 16 + 1024;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    static bool attr_set_dev[kMaxDevices] = {false};
    bool& attr_set = attr_set_dev[kctx_device()];
    if (!attr_set) {
        if (cudaFuncSetAttribute(flash_attn_kernel<DCH, KV_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) !=
            cudaSuccess)
            return -1;
        attr_set = true;
    }
    dim3 grid((p.seq_q + 127) / 128, p.heads, p.batch);
    launch_k(flash_attn_kernel<DCH, KV_STAGES>, dim3(grid), dim3(192), SMEM, st, mq, mk, mv, p);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

static char g_fa_err[256] = "";
const char* flash_last_error() { return g_fa_err; }

int flash_attn_setup(FlashOp* op, const __half* q, const __half* k, const __half* v, int ld, long long bs, __half* out,
                     int ldo, long long o_bs, int seq, int heads, int d, int batch) {
    if ((d % 8) != 0 || d > 192 || (ld % 8) != 0) {
        snprintf(g_fa_err, sizeof(g_fa_err), "flash attention: unsupported head dim %d / row stride %d", d, ld);
        return -1;
    }
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)seq, (uint64_t)heads, (uint64_t)batch};
    uint64_t st[3] = {(uint64_t)ld * 2, (uint64_t)d * 2, (uint64_t)bs * 2};
    if (batch == 1) st[2] = (uint64_t)ld * 2 * seq;
    uint32_t box[4] = {64, 128, 1, 1};
    if (make_map_4d(&op->mq, q, dims, st, box) || make_map_4d(&op->mk, k, dims, st, box) ||
        make_map_4d(&op->mv, v, dims, st, box)) {
        snprintf(g_fa_err, sizeof(g_fa_err), "flash attention tensor map: %s", gemm_last_error());
        return -1;
    }
    op->seq = seq;
    op->heads = heads;
    op->d = d;
    op->batch = batch;
    op->out = out;
    op->ldo = ldo;
    op->o_bs = o_bs;
    return 0;
}

int flash_attn_launch(const FlashOp* op, cudaStream_t st) {
    FlashParams p;
    p.seq_q = p.seq_kv = op->seq;
    p.heads = op->heads;
    p.batch = op->batch;
    p.d = op->d;
    p.dN = (op->d + 15) / 16 * 16;
    p.ksteps_qk = (op->d + 15) / 16;
    p.scale_log2 = (1.0f / sqrtf(static_cast<float>(op->d))) * 1.4426950408889634f;
    p.out = op->out;
    p.ldo = op->ldo;
    p.o_bs = op->o_bs;
    int r;
    static const bool two_wg = []() {
        const char* e = getenv("DTP_FLASH2");
        return !(e && e[0] == '0');
    }();
    static const int variant = []() {
        const char* e = getenv("DTP_FLASH_VARIANT");
        return e ? atoi(e) : 2;
    }();
    if (variant == 3 && op->seq >= 128 && op->d <= 64)
        r = launch_flash3<1, 2>(op->mq, op->mk, op->mv, p, st);
    else if (variant == 3 && op->seq >= 128 && op->d <= 128)
        r = launch_flash3<2, 2>(op->mq, op->mk, op->mv, p, st);
    else if (variant == 3 && op->seq >= 128)
        r = launch_flash3<3, 1>(op->mq, op->mk, op->mv, p, st);
    else if (two_wg && op->seq >= 256 && op->d <= 64)
        r = launch_flash2<1, 2>(op->mq, op->mk, op->mv, p, st);
    else if (two_wg && op->seq >= 256 && op->d <= 128)
        r = launch_flash2<2, 1>(op->mq, op->mk, op->mv, p, st);
    else if (op->d <= 64)
        r = launch_flash<1, 1>(op->mq, op->mk, op->mv, p, st);
    else if (op->d <= 128)
        r = launch_flash<2, 2>(op->mq, op->mk, op->mv, p, st);
    else
        r = launch_flash<3, 1>(op->mq, op->mk, op->mv, p, st);
    if (r) snprintf(g_fa_err, sizeof(g_fa_err), "flash attention launch: %s", cudaGetErrorString(cudaGetLastError()));
    return r;
}

}  // namespace dtp
