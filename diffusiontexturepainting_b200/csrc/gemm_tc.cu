// tcgen05 / TMEM / TMA contraction kernel for sm_100a. See gemm_tc.h for the problem statement.
//
// CTA = 192 threads: warp 0 (one lane) is the TMA producer, warp 1 allocates TMEM and (one lane) issues
// tcgen05.mma, warps 2..5 are the epilogue (each owns the 32 TMEM lanes of its warp-id % 4 quarter).
// Operand tiles are 128 x 64 (A) and BN x 64 (B) fp16, 128-byte swizzled, STAGES-deep mbarrier ring.
#include "gemm_tc.h"

#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace dtp {

static char g_gemm_err[256] = "";
const char* gemm_last_error() { return g_gemm_err; }

// ------------------------------------------------------------------------------------------------------------
// epilogue (shared by the main kernel and the split-K finalize kernel)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void epilogue_store32(const GemmParams& p, long long out_off, int row, int col0,
                                                 float (&v)[32]) {
    const int N = p.N;
    const int ncols = min(32, N - col0);
    if (ncols <= 0) return;
    const int flags = p.flags;
    const float rowbias = (p.bias != nullptr && (flags & EPI_BIAS_M)) ? __ldg(p.bias + row) : 0.0f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        float x = v[i] * p.alpha;
        if (p.bias != nullptr) {
            if (flags & EPI_BIAS_M)
                x += rowbias;
            else if (i < ncols)
                x += __ldg(p.bias + col0 + i);
        }
        if (flags & EPI_GELU) x = gelu_erf_f(x);
        if (flags & EPI_QUICKGELU) x = quick_gelu_f(x);
        if (flags & EPI_SILU) x = silu_f(x);
        v[i] = x;
    }
    if (flags & EPI_GEGLU) {
        // columns [0,16) hold a_j, [16,32) hold the gates g_j of the same 16 outputs
        const int oc0 = col0 >> 1;
        __half* out = reinterpret_cast<__half*>(p.out) + out_off + static_cast<long long>(row) * p.ldc + oc0;
        __align__(16) __half h[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) h[i] = __float2half_rn(v[i] * gelu_erf_f(v[16 + i]));
        if ((p.ldc & 7) == 0) {
            reinterpret_cast<uint4*>(out)[0] = reinterpret_cast<const uint4*>(h)[0];
            reinterpret_cast<uint4*>(out)[1] = reinterpret_cast<const uint4*>(h)[1];
        } else {
            for (int i = 0; i < 16; ++i) out[i] = h[i];
        }
        return;
    }
    if (p.residual != nullptr) {
        const __half* r = p.residual + static_cast<long long>(row) * p.ldr + col0;
        if (ncols == 32 && (p.ldr & 7) == 0) {
            __align__(16) __half h[32];
#pragma unroll
            for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(h)[i] = __ldg(reinterpret_cast<const uint4*>(r) + i);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += __half2float(h[i]);
        } else {
            for (int i = 0; i < ncols; ++i) v[i] += __half2float(r[i]);
        }
    }
    if (flags & EPI_IMG01) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fminf(fmaxf(v[i] * 0.5f + 0.5f, 0.0f), 1.0f);
    }
    if (flags & EPI_OUT_F32_NCHW) {
        float* out = reinterpret_cast<float*>(p.out) + out_off;
        const int img = row / p.hw_out, pix = row - img * p.hw_out;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < ncols) out[(static_cast<long long>(img) * N + col0 + i) * p.hw_out + pix] = v[i];
        return;
    }
    if (flags & EPI_OUT_F32) {
        float* out = reinterpret_cast<float*>(p.out) + out_off + static_cast<long long>(row) * p.ldc + col0;
        if (ncols == 32 && (p.ldc & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                reinterpret_cast<float4*>(out)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
            for (int i = 0; i < ncols; ++i) out[i] = v[i];
        }
        return;
    }
    __half* out = reinterpret_cast<__half*>(p.out) + out_off + static_cast<long long>(row) * p.ldc + col0;
    if (ncols == 32 && (p.ldc & 7) == 0 && ((out_off & 7) == 0)) {
        __align__(16) __half h[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) h[i] = __float2half_rn(v[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(h)[i];
    } else {
        for (int i = 0; i < ncols; ++i) out[i] = __float2half_rn(v[i]);
    }
}

// bounded mbarrier wait: a descriptor / byte-count bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("dtp gemm: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
                   threadIdx.x);
            __trap();
        }
    }
}

__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define DBG_MARK(slot)                                                                                         \
    do {                                                                                                       \
        if (p.dbg != nullptr)                                                                                  \
            p.dbg[(static_cast<long long>(blockIdx.z) * gridDim.y * gridDim.x + blockIdx.y * gridDim.x +      \
                   blockIdx.x) * 8 + (slot)] = gtime();                                                        \
    } while (0)

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                   const __grid_constant__ CUtensorMap mapB, const GemmParams p) {
    constexpr int A_BYTES = 128 * 128;
    constexpr int B_BYTES = BN * 128;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int TM_COLS = BN < 32 ? 32 : BN;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int zsplit = blockIdx.z % p.splits;
    const int zb = blockIdx.z / p.splits;
    const int z1 = zb % p.nz1, z2 = zb / p.nz1;
    const int kb_begin = static_cast<int>(static_cast<long long>(zsplit) * p.num_kb / p.splits);
    const int kb_end = static_cast<int>(static_cast<long long>(zsplit + 1) * p.num_kb / p.splits);
    const int m_tile = blockIdx.x;
    const int n0 = blockIdx.y * BN;
    const bool b_mn = (p.flags & GEMM_B_MN) != 0;

    int x0 = 0, y0 = 0, img0 = 0;
    if (p.mode == 1) {
        const int tx = m_tile % p.tiles_x;
        const int ty = (m_tile / p.tiles_x) % p.tiles_y;
        const int tn = m_tile / (p.tiles_x * p.tiles_y);
        x0 = tx * p.bw;
        y0 = ty * p.bh;
        img0 = tn * p.bn;
    }

    if (threadIdx.x == 0) DBG_MARK(0);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapA1);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) DBG_MARK(1);

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------ TMA producer ------------------------------
            const uint32_t a_bytes = (p.mode == 1) ? static_cast<uint32_t>(p.rows_valid) * 128u : A_BYTES;
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * STAGE_BYTES;
                uint8_t* sb = sa + A_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], a_bytes + B_BYTES);
                if (p.mode == 1) {
                    const int tap = kb / p.cblocks;
                    const int cb = kb - tap * p.cblocks;
                    const int ky = tap / 3, kx = tap - ky * 3;
                    if (cb < p.cblocks0)
                        tma_load_4d(sa, &mapA0, &full_bar[stage], cb * 64, x0 + kx - 1, y0 + ky - 1, img0);
                    else
                        tma_load_4d(sa, &mapA1, &full_bar[stage], (cb - p.cblocks0) * 64, x0 + kx - 1, y0 + ky - 1,
                                    img0);
                } else {
                    const int za = p.a_batched ? z1 : 0, zb2 = p.a_batched ? z2 : 0;
                    if (kb < p.cblocks0)
                        tma_load_4d(sa, &mapA0, &full_bar[stage], kb * 64, m_tile * 128, za, zb2);
                    else
                        tma_load_4d(sa, &mapA1, &full_bar[stage], (kb - p.cblocks0) * 64, m_tile * 128, za, zb2);
                }
                const int bz1 = p.b_batched ? z1 : 0, bz2 = p.b_batched ? z2 : 0;
                if (!b_mn) {
                    tma_load_4d(sb, &mapB, &full_bar[stage], kb * 64, n0, bz1, bz2);
                } else {
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j)
                        tma_load_4d(sb + j * 8192, &mapB, &full_bar[stage], n0 + j * 64, kb * 64, bz1, bz2);
                }
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------ MMA issuer ------------------------------
            const uint32_t idesc = umma_idesc_f16(128, BN, 0, b_mn ? 1 : 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                mbar_wait_bounded(&full_bar[stage], phase);
                if (kb == kb_begin) DBG_MARK(2);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                const uint32_t sb = sa + A_BYTES;
                const uint64_t adesc = umma_desc_k_sw128(sa);
                const uint64_t bdesc = b_mn ? umma_desc_mn_sw128(sb, 8192) : umma_desc_k_sw128(sb);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // advance 16 K-elements: 32 B inside the swizzle atom (K-major) / two 8-row groups (MN-major)
                    const uint64_t ad = adesc + static_cast<uint64_t>(k * 2);
                    const uint64_t bd = bdesc + static_cast<uint64_t>(b_mn ? k * 128 : k * 2);
                    umma_f16(tmem_base, ad, bd, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(tmem_full_bar);
            DBG_MARK(3);
        }
    } else {
        // ------------------------------ epilogue warps ------------------------------
        const int q = warp & 3;
        const int r = q * 32 + lane;
        int row = -1;
        if (p.mode == 1) {
            if (r < p.rows_valid) {
                const int w = r % p.bw;
                const int hh = (r / p.bw) % p.bh;
                const int nn = r / (p.bw * p.bh);
                const int img = img0 + nn;
                if (img < p.Nimg) row = (img * p.H + y0 + hh) * p.W + x0 + w;
            }
        } else {
            const int m = m_tile * 128 + r;
            if (m < p.M) row = m;
        }
        const long long out_off = static_cast<long long>(z1) * p.out_zs1 + static_cast<long long>(z2) * p.out_zs2;
        mbar_wait_bounded(tmem_full_bar, 0);
        tc_fence_after();
        if (threadIdx.x == 64) DBG_MARK(4);
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            if (n0 + c >= p.N) break;
            uint32_t raw[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c), raw);
            tmem_ld_wait();
            if (row >= 0) {
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                if (p.splits > 1) {
                    float* ws = p.workspace +
                                (static_cast<long long>(blockIdx.z) * p.M + row) * static_cast<long long>(p.N) + n0 + c;
                    const int ncols = min(32, p.N - (n0 + c));
                    if (ncols == 32 && (p.N & 3) == 0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            reinterpret_cast<float4*>(ws)[i] =
                                make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    } else {
                        for (int i = 0; i < ncols; ++i) ws[i] = v[i];
                    }
                } else {
                    epilogue_store32(p, out_off, row, n0 + c, v);
                }
            }
        }
    }
    if (threadIdx.x == 64) DBG_MARK(5);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TM_COLS);
    if (threadIdx.x == 32) DBG_MARK(6);
}

// sums split-K partials and runs the epilogue; one thread per (row, 32-column chunk)
__global__ void __launch_bounds__(256) gemm_splitk_finalize_kernel(const GemmParams p) {
    const int chunks = (p.N + 31) / 32;
    const long long total = static_cast<long long>(p.nz1) * p.nz2 * p.M * chunks;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int chunk = static_cast<int>(idx % chunks);
    const long long rz = idx / chunks;
    const int row = static_cast<int>(rz % p.M);
    const int zb = static_cast<int>(rz / p.M);
    const int col0 = chunk * 32;
    const int ncols = min(32, p.N - col0);
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.0f;
    for (int s = 0; s < p.splits; ++s) {
        const float* ws =
            p.workspace + (static_cast<long long>(zb * p.splits + s) * p.M + row) * static_cast<long long>(p.N) + col0;
        if (ncols == 32 && (p.N & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(ws) + i);
                v[4 * i] += t.x;
                v[4 * i + 1] += t.y;
                v[4 * i + 2] += t.z;
                v[4 * i + 3] += t.w;
            }
        } else {
            for (int i = 0; i < ncols; ++i) v[i] += ws[i];
        }
    }
    const int z1 = zb % p.nz1, z2 = zb / p.nz1;
    const long long out_off = static_cast<long long>(z1) * p.out_zs1 + static_cast<long long>(z2) * p.out_zs2;
    epilogue_store32(p, out_off, row, col0, v);
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "cuTensorMapEncodeTiled entry point unavailable (%s)",
                 cudaGetErrorString(e));
        return nullptr;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    return fn;
}

int make_map_4d(CUtensorMap* m, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                const uint32_t box[4]) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return -1;
    cuuint64_t gd[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t gs[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
    cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "tensor map base %p not 16-byte aligned", base);
        return -2;
    }
    for (int i = 0; i < 3; ++i)
        if (gs[i] % 16 != 0) {
            snprintf(g_gemm_err, sizeof(g_gemm_err), "tensor map stride[%d]=%llu not a multiple of 16", i,
                     (unsigned long long)gs[i]);
            return -3;
        }
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gd, gs, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_gemm_err, sizeof(g_gemm_err),
                 "cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu %llu strides %llu %llu %llu box %u %u %u %u",
                 (int)r, (unsigned long long)gd[0], (unsigned long long)gd[1], (unsigned long long)gd[2],
                 (unsigned long long)gd[3], (unsigned long long)gs[0], (unsigned long long)gs[1],
                 (unsigned long long)gs[2], bx[0], bx[1], bx[2], bx[3]);
        return -4;
    }
    return 0;
}

static void params_defaults(GemmParams& p) {
    memset(&p, 0, sizeof(p));
    p.splits = 1;
    p.nz1 = 1;
    p.nz2 = 1;
    p.alpha = 1.0f;
}

static int fix_bn(int BN) { return (BN == 32 || BN == 64 || BN == 128 || BN == 256) ? BN : 128; }

// 2-D (K inner, rows outer) map expressed as 4-D with unit batch dims
static int map_rows(CUtensorMap* m, const __half* base, uint64_t K, uint64_t rows, uint64_t ld, uint32_t box_rows) {
    uint64_t dims[4] = {K, rows, 1, 1};
    uint64_t st[3] = {ld * 2, ld * 2 * rows, ld * 2 * rows};
    uint32_t box[4] = {64, box_rows, 1, 1};
    return make_map_4d(m, base, dims, st, box);
}

int gemm_setup_linear(GemmOp* op, const __half* A0, int lda0, int K0, const __half* A1, int lda1, int K1, int M,
                      const __half* Wt, int ldw, int N, int BN, int splits) {
    params_defaults(op->p);
    GemmParams& p = op->p;
    BN = fix_bn(BN);
    if (A1 != nullptr && (K0 % 64) != 0) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "dual-source linear needs K0 %% 64 == 0 (K0=%d)", K0);
        return -10;
    }
    const int kb0 = (K0 + 63) / 64, kb1 = A1 ? (K1 + 63) / 64 : 0;
    p.M = M;
    p.N = N;
    p.num_kb = kb0 + kb1;
    p.mode = 0;
    p.cblocks0 = kb0;
    p.cblocks = kb0 + kb1;
    p.splits = splits < 1 ? 1 : (splits > p.num_kb ? p.num_kb : splits);
    p.ldc = N;
    op->BN = BN;
    op->grid_m = (M + 127) / 128;
    int r = map_rows(&op->mapA0, A0, K0, M, lda0, 128);
    if (r) return r;
    if (A1) {
        r = map_rows(&op->mapA1, A1, K1, M, lda1, 128);
        if (r) return r;
    } else {
        op->mapA1 = op->mapA0;
    }
    return map_rows(&op->mapB, Wt, K0 + (A1 ? K1 : 0), N, ldw, BN);
}

static int largest_divisor_le(int n, int cap) {
    int best = 1;
    for (int d = 1; d <= n && d <= cap; ++d)
        if (n % d == 0) best = d;
    return best;
}

int gemm_setup_conv3x3(GemmOp* op, const __half* A0, int C0, const __half* A1, int C1, int Nimg, int H, int W,
                       const __half* Wt, int Cout, int BN, int splits) {
    params_defaults(op->p);
    GemmParams& p = op->p;
    BN = fix_bn(BN);
    if ((C0 % 64) != 0 || (A1 != nullptr && (C1 % 64) != 0)) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "conv3x3 needs channel counts %% 64 == 0 (C0=%d C1=%d)", C0, C1);
        return -11;
    }
    const int C = C0 + (A1 ? C1 : 0);
    p.M = Nimg * H * W;
    p.N = Cout;
    p.mode = 1;
    p.H = H;
    p.W = W;
    p.Nimg = Nimg;
    p.bw = largest_divisor_le(W, 128);
    p.bh = largest_divisor_le(H, 128 / p.bw);
    p.bn = (p.bh == H) ? (128 / (p.bw * p.bh)) : 1;
    if (p.bn > Nimg) p.bn = Nimg;
    if (p.bn < 1) p.bn = 1;
    p.tiles_x = W / p.bw;
    p.tiles_y = H / p.bh;
    p.rows_valid = p.bw * p.bh * p.bn;
    p.cblocks0 = C0 / 64;
    p.cblocks = C / 64;
    p.num_kb = 9 * p.cblocks;
    p.splits = splits < 1 ? 1 : (splits > p.num_kb ? p.num_kb : splits);
    p.ldc = Cout;
    op->BN = BN;
    op->grid_m = p.tiles_x * p.tiles_y * ((Nimg + p.bn - 1) / p.bn);
    {
        uint64_t dims[4] = {(uint64_t)C0, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        uint64_t st[3] = {(uint64_t)C0 * 2, (uint64_t)C0 * 2 * W, (uint64_t)C0 * 2 * W * H};
        uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
        int r = make_map_4d(&op->mapA0, A0, dims, st, box);
        if (r) return r;
    }
    if (A1) {
        uint64_t dims[4] = {(uint64_t)C1, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        uint64_t st[3] = {(uint64_t)C1 * 2, (uint64_t)C1 * 2 * W, (uint64_t)C1 * 2 * W * H};
        uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
        int r = make_map_4d(&op->mapA1, A1, dims, st, box);
        if (r) return r;
    } else {
        op->mapA1 = op->mapA0;
    }
    return map_rows(&op->mapB, Wt, (uint64_t)9 * C, Cout, (uint64_t)9 * C, BN);
}

int gemm_setup_batched(GemmOp* op, const __half* A, int lda, long long a_zs1, long long a_zs2, const __half* B, int ldb,
                       long long b_zs1, long long b_zs2, int b_mn, int M, int N, int K, int nz1, int nz2, int BN) {
    params_defaults(op->p);
    GemmParams& p = op->p;
    BN = fix_bn(BN);
    if (b_mn && BN < 64) BN = 64;
    p.M = M;
    p.N = N;
    p.num_kb = (K + 63) / 64;
    p.mode = 0;
    p.cblocks0 = p.num_kb;
    p.cblocks = p.num_kb;
    p.nz1 = nz1;
    p.nz2 = nz2;
    p.a_batched = 1;
    p.b_batched = 1;
    p.ldc = N;
    if (b_mn) p.flags |= GEMM_B_MN;
    op->BN = BN;
    op->grid_m = (M + 127) / 128;
    // size-1 batch dims still need a legal (multiple of 16 B) stride
    auto zstride = [](long long s, uint64_t fallback) -> uint64_t { return s > 0 ? (uint64_t)s * 2 : fallback; };
    {
        uint64_t dims[4] = {(uint64_t)K, (uint64_t)M, (uint64_t)nz1, (uint64_t)nz2};
        uint64_t st[3] = {(uint64_t)lda * 2, zstride(a_zs1, (uint64_t)lda * 2 * M), zstride(a_zs2, (uint64_t)lda * 2 * M)};
        uint32_t box[4] = {64, 128, 1, 1};
        int r = make_map_4d(&op->mapA0, A, dims, st, box);
        if (r) return r;
        op->mapA1 = op->mapA0;
    }
    if (!b_mn) {
        uint64_t dims[4] = {(uint64_t)K, (uint64_t)N, (uint64_t)nz1, (uint64_t)nz2};
        uint64_t st[3] = {(uint64_t)ldb * 2, zstride(b_zs1, (uint64_t)ldb * 2 * N), zstride(b_zs2, (uint64_t)ldb * 2 * N)};
        uint32_t box[4] = {64, (uint32_t)BN, 1, 1};
        return make_map_4d(&op->mapB, B, dims, st, box);
    } else {
        uint64_t dims[4] = {(uint64_t)N, (uint64_t)K, (uint64_t)nz1, (uint64_t)nz2};
        uint64_t st[3] = {(uint64_t)ldb * 2, zstride(b_zs1, (uint64_t)ldb * 2 * K), zstride(b_zs2, (uint64_t)ldb * 2 * K)};
        uint32_t box[4] = {64, 64, 1, 1};
        return make_map_4d(&op->mapB, B, dims, st, box);
    }
}

size_t gemm_workspace_bytes(const GemmOp* op) {
    const GemmParams& p = op->p;
    if (p.splits <= 1) return 0;
    return static_cast<size_t>(p.nz1) * p.nz2 * p.splits * p.M * static_cast<size_t>(p.N) * sizeof(float);
}

void gemm_pick_config(int mtiles, int N, int num_kb, int flags, int* BN, int* splits) {
    const int kSMs = 148;
    int bn = 32;
    if (N > 32) bn = 64;
    if (N > 64) bn = 128;
    if (N > 128) {
        // prefer the 256-wide tile (half the A re-reads per flop) once it still fills the machine
        const int t256 = mtiles * ((N + 255) / 256);
        bn = (t256 >= kSMs) ? 256 : 128;
        if (bn == 128 && mtiles * ((N + 127) / 128) < kSMs / 2 && (flags & GEMM_B_MN) == 0) bn = 64;
    }
    const int tiles = mtiles * ((N + bn - 1) / bn);
    int sp = 1;
    if (tiles < kSMs && num_kb >= 8) {
        sp = kSMs / tiles;
        if (sp > num_kb / 4) sp = num_kb / 4;
        if (sp > 16) sp = 16;
        if (sp < 1) sp = 1;
    }
    *BN = bn;
    *splits = sp;
}

template <int BN, int STAGES>
static int launch_cfg(const GemmOp* op, cudaStream_t stream) {
    constexpr int STAGE_BYTES = 128 * 128 + BN * 128;
    constexpr int SMEM = STAGES * STAGE_BYTES + (2 * STAGES + 1) * 8 + 16 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e =
            cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) {
            snprintf(g_gemm_err, sizeof(g_gemm_err), "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return -20;
        }
        attr_set = true;
    }
    const GemmParams& p = op->p;
    dim3 grid(op->grid_m, (p.N + BN - 1) / BN, p.nz1 * p.nz2 * p.splits);
    gemm_tc_kernel<BN, STAGES><<<grid, 192, SMEM, stream>>>(op->mapA0, op->mapA1, op->mapB, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "gemm launch: %s", cudaGetErrorString(e));
        return -21;
    }
    return 0;
}

int gemm_launch(const GemmOp* op, cudaStream_t stream) {
    const GemmParams& p = op->p;
    if (p.splits > 1 && p.workspace == nullptr) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "split-K gemm without workspace");
        return -22;
    }
    int r;
    switch (op->BN) {
        case 32: r = launch_cfg<32, 8>(op, stream); break;
        case 64: r = launch_cfg<64, 8>(op, stream); break;
        case 128: r = launch_cfg<128, 6>(op, stream); break;
        default: r = launch_cfg<256, 4>(op, stream); break;
    }
    if (r) return r;
    if (p.splits > 1) {
        const int chunks = (p.N + 31) / 32;
        const long long total = static_cast<long long>(p.nz1) * p.nz2 * p.M * chunks;
        const int blocks = static_cast<int>((total + 255) / 256);
        gemm_splitk_finalize_kernel<<<blocks, 256, 0, stream>>>(p);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            snprintf(g_gemm_err, sizeof(g_gemm_err), "finalize launch: %s", cudaGetErrorString(e));
            return -23;
        }
    }
    return 0;
}

}  // namespace dtp
