// tcgen05 / TMEM / TMA contraction kernel for sm_100a. See gemm_tc.h for the problem statement.
//
// CTA = 320 threads: warp 0 (one lane) is the TMA producer, warp 1 allocates TMEM and (one lane) issues tcgen05.mma,
// warps 2..9 are the epilogue (two per TMEM lane quarter, warp-id % 4, taking alternate 32-column chunks of the tile).
// Operand tiles are 128 x 64 (A) and BN x 64 (B) fp16, 128-byte swizzled, STAGES-deep mbarrier ring; persistent tile loop
// with two TMEM accumulators. CL = 2: CTA pairs (cta_group::2, one 256-row MMA per pair, each CTA fetching its 128 rows of
// A and half of the weight tile); BN = 320 exists in pair mode only (two N = 160 MMAs per k-step, one accumulator).
// Split-K partials are reduced inside the kernel when the host can guarantee one tile per CTA on a co-resident grid.
#include "gemm_tc.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kctx.h"

namespace dtp {

static char g_gemm_err[256] = "";
const char* gemm_last_error() { return g_gemm_err; }

// ------------------------------------------------------------------------------------------------------------
// epilogue: one thread owns one output row and 32 consecutive columns of it (the tcgen05.ld 32x32b shape).
// Shared by the main kernel and the split-K finalize kernel. No runtime-indexed register arrays: everything below
// unrolls to compile-time register indices (the first version kept v[] in local memory and spent ~6 us per chunk).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float quick_gelu_fast(float x) { return __fdividef(x, 1.0f + __expf(-1.702f * x)); }

__device__ __forceinline__ void store_half_chunk(__half* out, const float (&v)[32], int ncols, bool vec_ok) {
    if (ncols == 32 && vec_ok && aligned32(out)) {
        // 64 bytes per lane as two full-sector stores
#pragma unroll
        for (int i = 0; i < 2; ++i)
            st_global_v8(out + 16 * i, pack_half2(v[16 * i + 0], v[16 * i + 1]), pack_half2(v[16 * i + 2], v[16 * i + 3]),
                         pack_half2(v[16 * i + 4], v[16 * i + 5]), pack_half2(v[16 * i + 6], v[16 * i + 7]),
                         pack_half2(v[16 * i + 8], v[16 * i + 9]), pack_half2(v[16 * i + 10], v[16 * i + 11]),
                         pack_half2(v[16 * i + 12], v[16 * i + 13]), pack_half2(v[16 * i + 14], v[16 * i + 15]));
    } else if (ncols == 32 && vec_ok) {
        uint4* o = reinterpret_cast<uint4*>(out);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = pack_half2(v[8 * i + 0], v[8 * i + 1]);
            u.y = pack_half2(v[8 * i + 2], v[8 * i + 3]);
            u.z = pack_half2(v[8 * i + 4], v[8 * i + 5]);
            u.w = pack_half2(v[8 * i + 6], v[8 * i + 7]);
            o[i] = u;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < ncols) out[i] = __float2half_rn(v[i]);
    }
}

// fp32 chunk (split-K partials, fp32 outputs): 128 bytes per lane as four full-sector stores
__device__ __forceinline__ void store_f32_chunk(float* out, const float (&v)[32], int ncols, bool vec_ok) {
    if (ncols == 32 && vec_ok && aligned32(out)) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            st_global_v8f(out + 8 * i, v[8 * i], v[8 * i + 1], v[8 * i + 2], v[8 * i + 3], v[8 * i + 4], v[8 * i + 5], v[8 * i + 6],
                          v[8 * i + 7]);
    } else if (ncols == 32 && vec_ok) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            reinterpret_cast<float4*>(out)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < ncols) out[i] = v[i];
    }
}

// folded LayerNorm: (rstd, -rstd * mean) of A row `grow` from the producer's per-chunk partial sums (gemm_tc.h)
__device__ __forceinline__ void ln_row_coef(const GemmParams& p, long long grow, float& ln_r, float& ln_nm) {
    const float2* sp = p.ln_stats + grow;
    float s = 0.0f, q = 0.0f;
    // the warp issues in order: every load of a batch must be issued before the first add that consumes one, or each
    // partial costs a full L2 round trip (C / 32 = 10 / 20 / 40 partials per row at the UNet widths)
    for (int ch = 0; ch < p.ln_chunks; ch += 10) {
        float2 t[10];
#pragma unroll
        for (int u = 0; u < 10; ++u)
            t[u] = (ch + u < p.ln_chunks) ? __ldcg(sp + static_cast<long long>(ch + u) * p.ln_rows) : make_float2(0.0f, 0.0f);
#pragma unroll
        for (int u = 0; u < 10; ++u) {
            s += t[u].x;
            q += t[u].y;
        }
    }
    const float mean = s * p.ln_inv_c;
    const float var = fmaxf(q * p.ln_inv_c - mean * mean, 0.0f);
    ln_r = rsqrtf(var + p.ln_eps);
    ln_nm = -ln_r * mean;
}
// (sum, sum of squares) of 32 output values for a following folded LayerNorm. Taken before the fp16 rounding of the store:
// the rounding errors are zero-mean and 2^-11 relative, far below what the statistics resolve, and the conversion round
// trip would double the instruction count of this (exposed) part of the epilogue.
__device__ __forceinline__ void emit_row_stats(const GemmParams& p, long long grow, int col0, const float (&v)[32]) {
    float s0 = 0.0f, s1 = 0.0f, q0 = 0.0f, q1 = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
        s0 += v[i];
        s1 += v[i + 1];
        q0 = fmaf(v[i], v[i], q0);
        q1 = fmaf(v[i + 1], v[i + 1], q1);
    }
    p.stats_out[static_cast<long long>(col0 >> 5) * p.stats_rows + grow] = make_float2(s0 + s1, q0 + q1);
}

// bias_chunk: 32 floats for columns col0.. (shared memory in the main kernel, global in the finalize kernel), or nullptr.
// grow: row index in the statistics buffers ((z2 * nz1 + z1) * M + row); colsum_chunk: like bias_chunk, for the folded LayerNorm
__device__ __forceinline__ void epilogue_store32(const GemmParams& p, long long out_off, long long res_off, int row,
                                                 int col0, float (&v)[32], const float* bias_chunk,
                                                 long long grow = 0, const float* colsum_chunk = nullptr,
                                                 bool feat = true) {
    const int N = p.N;
    const int ncols = min(32, N - col0);
    if (ncols <= 0) return;
    const int flags = p.flags;
    const float alpha = p.alpha;
    // residual loads first: their latency overlaps the arithmetic below
    const bool has_res = p.residual != nullptr;
    uint4 rraw[4];
    const __half* rp = has_res ? p.residual + res_off + static_cast<long long>(row) * p.ldr + col0 : nullptr;
    const bool res_vec = has_res && ncols == 32 && (p.ldr & 7) == 0;
    if (res_vec) {
        if (aligned32(rp)) {
            uint32_t t0[8], t1[8];
            ld_global_nc_v8(rp, t0);
            ld_global_nc_v8(rp + 16, t1);
            rraw[0] = make_uint4(t0[0], t0[1], t0[2], t0[3]);
            rraw[1] = make_uint4(t0[4], t0[5], t0[6], t0[7]);
            rraw[2] = make_uint4(t1[0], t1[1], t1[2], t1[3]);
            rraw[3] = make_uint4(t1[4], t1[5], t1[6], t1[7]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) rraw[i] = __ldg(reinterpret_cast<const uint4*>(rp) + i);
        }
    }
    if (feat && p.ln_stats != nullptr) {
        float ln_r, ln_nm;
        ln_row_coef(p, grow, ln_r, ln_nm);
#pragma unroll
        for (int i = 0; i < 32; ++i)
            v[i] = fmaf(v[i], ln_r, fmaf(ln_nm, colsum_chunk[i], bias_chunk != nullptr ? bias_chunk[i] : 0.0f));
    } else if (flags & EPI_BIAS_M) {
        const float rb = p.bias ? __ldg(p.bias + row) : 0.0f;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], alpha, rb);
    } else if (bias_chunk != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], alpha, bias_chunk[i]);
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= alpha;
    }
    if (flags & EPI_SOFTMAX16) {
        // folded cross-attention: each 16-column group holds the scores of one head against the (<= 16) image tokens
        const int nv = p.aux;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            float m = -INFINITY;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < nv) m = fmaxf(m, v[16 * g + i]);
            float sum = 0.0f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float e = (i < nv) ? __expf(v[16 * g + i] - m) : 0.0f;
                v[16 * g + i] = e;
                sum += e;
            }
            const float inv = __fdividef(1.0f, sum);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[16 * g + i] *= inv;
        }
    }
    if (flags & EPI_GEGLU) {
        // columns [0,16) hold a_j, [16,32) hold the gates g_j of the same 16 outputs
        __half* out = reinterpret_cast<__half*>(p.out) + out_off + static_cast<long long>(row) * p.ldc + (col0 >> 1);
        float o[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = v[i] * gelu_erf_f(v[16 + i]);
        if ((p.ldc & 7) == 0 && aligned32(out)) {
            st_global_v8(out, pack_half2(o[0], o[1]), pack_half2(o[2], o[3]), pack_half2(o[4], o[5]), pack_half2(o[6], o[7]),
                         pack_half2(o[8], o[9]), pack_half2(o[10], o[11]), pack_half2(o[12], o[13]), pack_half2(o[14], o[15]));
        } else if ((p.ldc & 7) == 0) {
            uint4* d = reinterpret_cast<uint4*>(out);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                uint4 u;
                u.x = pack_half2(o[8 * i + 0], o[8 * i + 1]);
                u.y = pack_half2(o[8 * i + 2], o[8 * i + 3]);
                u.z = pack_half2(o[8 * i + 4], o[8 * i + 5]);
                u.w = pack_half2(o[8 * i + 6], o[8 * i + 7]);
                d[i] = u;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) out[i] = __float2half_rn(o[i]);
        }
        return;
    }
    if (flags & (EPI_GELU | EPI_QUICKGELU | EPI_SILU)) {
        if (flags & EPI_GELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_erf_f(v[i]);
        } else if (flags & EPI_QUICKGELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = quick_gelu_fast(v[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = silu_fast(v[i]);
        }
    }
    if (has_res) {
        if (res_vec) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const __half2* h2 = reinterpret_cast<const __half2*>(&rraw[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(h2[j]);
                    v[8 * i + 2 * j] += f.x;
                    v[8 * i + 2 * j + 1] += f.y;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < ncols) v[i] += __half2float(rp[i]);
        }
    }
    if (flags & EPI_IMG01) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fminf(fmaxf(v[i] * 0.5f + 0.5f, 0.0f), 1.0f);
    }
    if (flags & EPI_OUT_F32_NCHW) {
        float* out = reinterpret_cast<float*>(p.out) + out_off;
        const int img = row / p.hw_out, pix = row - img * p.hw_out;
        float* o = out + (static_cast<long long>(img) * N + col0) * p.hw_out + pix;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < ncols) o[static_cast<long long>(i) * p.hw_out] = v[i];
        return;
    }
    if (flags & EPI_OUT_F32) {
        float* out = reinterpret_cast<float*>(p.out) + out_off + static_cast<long long>(row) * p.ldc + col0;
        store_f32_chunk(out, v, ncols, (p.ldc & 3) == 0 && (out_off & 3) == 0);
        return;
    }
    __half* out = reinterpret_cast<__half*>(p.out) + out_off + static_cast<long long>(row) * p.ldc + col0;
    store_half_chunk(out, v, ncols, (p.ldc & 7) == 0 && (out_off & 7) == 0);
    if (feat && p.stats_out != nullptr) emit_row_stats(p, grow, col0, v);
}

// bounded mbarrier wait: a descriptor / byte-count bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) __trap();
    }
}

__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// second bank of checkpoints (rows gridDim.x .. 2*gridDim.x-1 of the buffer): phases of the in-kernel split-K reduction
#define DBG_MARK2(slot)                                                                                  \
    do {                                                                                                 \
        if (p.dbg != nullptr) p.dbg[static_cast<long long>(gridDim.x + blockIdx.x) * 8 + (slot)] = gtime(); \
    } while (0)
#define DBG_MARK(slot)                                                                     \
    do {                                                                                   \
        if (p.dbg != nullptr) p.dbg[static_cast<long long>(blockIdx.x) * 8 + (slot)] = gtime(); \
    } while (0)

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// operand box load: in pair mode (CL == 2) the bytes are counted on the leader CTA's barrier
template <int CL>
__device__ __forceinline__ void tma4(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    if (CL == 2)
        tma_load_4d_pair(smem, m, bar, c0, c1, c2, c3);
    else
        tma_load_4d(smem, m, bar, c0, c1, c2, c3);
}

struct TileCoord {
    int m_tile, n_tile, n0, z1, z2, zsplit, zb, kb_begin, kb_end, x0, y0, img0;
};
// CL = 2: `t` indexes a PAIR of adjacent m-tiles handled by the two CTAs of a cluster (same n-tile, same k-range)
template <int BN, int CL>
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int t, int rank) {
    TileCoord c;
    const int gm = (CL == 2) ? (p.grid_m + 1) / 2 : p.grid_m;
    c.m_tile = (CL == 2) ? 2 * (t % gm) + rank : t % gm;
    int r = t / gm;
    const int n_tile = r % p.grid_n;
    r /= p.grid_n;
    c.zsplit = r % p.splits;
    c.zb = r / p.splits;
    c.z1 = c.zb % p.nz1;
    c.z2 = c.zb / p.nz1;
    c.n0 = n_tile * BN;
    c.n_tile = n_tile;
    c.kb_begin = static_cast<int>(static_cast<long long>(c.zsplit) * p.num_kb / p.splits);
    c.kb_end = static_cast<int>(static_cast<long long>(c.zsplit + 1) * p.num_kb / p.splits);
    c.x0 = c.y0 = c.img0 = 0;
    if (p.mode == 1) {
        const int tx = c.m_tile % p.tiles_x;
        const int ty = (c.m_tile / p.tiles_x) % p.tiles_y;
        const int tn = c.m_tile / (p.tiles_x * p.tiles_y);
        c.x0 = tx * p.bw;
        c.y0 = ty * p.bh;
        c.img0 = tn * p.bn;
    }
    return c;
}

__device__ __forceinline__ int tile_group(const GemmParams& p, const TileCoord& c) {
    return (c.zb * p.grid_n + c.n_tile) * p.grid_m + c.m_tile;
}
// Tile of persistent wave k for CTA (pair) `slot` of `nslots`: waves alternate direction, so the CTAs that got the cheap
// tail tiles of one wave (tiles are ordered n-tile-major: full-width tiles first, the ragged last n-tile at the end) get
// the first tiles of the next one. -1 past the end.
__device__ __forceinline__ int tile_of_wave(int k, int slot, int nslots, int total) {
    const int t = k * nslots + ((k & 1) ? nslots - 1 - slot : slot);
    return t < total ? t : -1;
}
// output row (NHWC pixel index / matrix row) of tile-local row r, or -1 when r is padding
__device__ __forceinline__ int tile_row(const GemmParams& p, const TileCoord& c, int r) {
    if (p.mode == 1) {
        if (r >= p.rows_valid) return -1;
        const int w = r % p.bw;
        const int hh = (r / p.bw) % p.bh;
        const int nn = r / (p.bw * p.bh);
        const int img = c.img0 + nn;
        if (img >= p.Nimg) return -1;
        if (p.up2) return (img * 2 * p.H + 2 * (c.y0 + hh) + (c.z1 >> 1)) * 2 * p.W + 2 * (c.x0 + w) + (c.z1 & 1);
        return (img * p.H + c.y0 + hh) * p.W + c.x0 + w;
    }
    const int m = c.m_tile * 128 + r;
    return (m < p.M) ? m : -1;
}

// Persistent kernel: grid = min(tiles, SMs); every role walks the same static tile sequence t = blockIdx.x + i*gridDim.x.
// Two TMEM accumulators (BN columns each) let the epilogue of tile i overlap the mainloop of tile i+1.
// OCC = 2: light configuration (few stages, BN <= 128 -> <= 256 TMEM columns, <= 102 registers) so that two CTAs share an SM:
// the next kernel's CTAs become resident while this one drains (PDL overlap instead of a kernel-boundary bubble) and two
// tiles interleave their load latency / mainloop / epilogue on one SM.
// FEAT = 1: the instantiation that can fold a LayerNorm into its epilogue and emit row statistics (gemm_tc.h). Kept apart
// from the plain one because the epilogue sits at the register cap: the extra live values cost spills in EVERY op when
// compiled into one kernel (measured: +16 us on an unrelated split-K convolution).
template <int BN, int STAGES, int CL, int OCC = 1, int FEAT = 0>
__global__ void __launch_bounds__(320, OCC)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                   const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapBL,
                   const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                   const __grid_constant__ GemmParams p) {
    // CL == 2: CTA pair (cta_group::2). One MMA spans both SMs (M = 256); each CTA's shared memory holds its own 128 rows
    // of A and HALF of the B tile (rows [rank*BN/2, +BN/2)), which halves the shared-memory traffic per CTA - the limit of
    // the single-CTA mainloop (TMA write + MMA read of A and B exceed 128 B/clk for wide tiles).
    constexpr int A_BYTES = 128 * 128;
    constexpr int B_BYTES = (BN / CL) * 128;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int TM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    // BN = 320 (CTA pairs only): one tile = two MMAs of N = 160 per k-step on the same A tile, a single 320-column
    // accumulator (no double buffering); the weight tile is two 160-row sub-tiles, each split between the CTAs of the pair.
    constexpr int NSUB = BN > 256 ? 2 : 1;
    constexpr int BNS = BN / NSUB;
    constexpr int NACC = (2 * BN <= 512) ? 2 : 1;
    static_assert(NSUB == 1 || CL == 2, "tiles wider than 256 columns exist in pair mode only");
    constexpr int B_CHUNKS = (BN + 63) / 64;  // MN-major B: 64-column boxes

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
    uint64_t* peer_full_bar = tmem_empty_bar + 2;   // [STAGES] spare; [0] = completion of the split-K slice copies
    uint64_t* red_bar = peer_full_bar;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_full_bar + STAGES);
    int* ticket_slot = reinterpret_cast<int*>(tmem_slot + 1);  // split-K arrival ticket of the current tile (epilogue warps)
    float* sbias = reinterpret_cast<float*>(tmem_slot + 4);  // [BN]
    float* scolsum = sbias + BN;                              // [BN] folded LayerNorm: row sums of the scaled weights

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const bool b_mn = (p.flags & GEMM_B_MN) != 0;
    const bool w_blocked = (p.flags & GEMM_W_BLOCKED) != 0;
    const int total_tiles = p.total_tiles;  // CL == 2: number of m-tile pairs
    const int rank = (CL == 2) ? static_cast<int>(cluster_ctarank()) : 0;
    const int tile0 = (CL == 2) ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int tstep = (CL == 2) ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

    pdl_launch_dependents();  // the next kernel of the stream may start its prologue now
    if (threadIdx.x == 0) DBG_MARK(0);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapA1);
        if (p.sc_blocks > 0 || p.down2) {
            tma_prefetch_desc(&mapA2);
            tma_prefetch_desc(&mapA3);
        }
        tma_prefetch_desc(&mapB);
        if (p.n_last > 0) tma_prefetch_desc(&mapBL);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&peer_full_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 8 * CL);  // one arrival per epilogue warp (of both CTAs in pair mode)
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (CL == 2)
            tmem_alloc2(tmem_slot, TM_COLS);
        else
            tmem_alloc(tmem_slot, TM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();  // the peer's barriers must exist before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) DBG_MARK(1);

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------ TMA producer ------------------------------
            const uint32_t a_bytes = (p.mode == 1) ? static_cast<uint32_t>(p.rows_valid) * 128u : A_BYTES;
            const uint32_t b_bytes = b_mn ? static_cast<uint32_t>(B_CHUNKS) * 8192u : B_BYTES;
            int stage = 0;
            uint32_t phase = 0;
            // Weights (the B operand of linear / conv problems) are never written by a kernel of the stream: their first
            // STAGES tiles are requested BEFORE waiting for the predecessor kernel, hiding the DRAM latency of
            // weight-streaming layers behind the previous kernel's tail. Activations (A) are only touched after the wait.
            int pre = 0;
            if (!p.b_batched && !b_mn && tile0 < total_tiles) {
                TileCoord c0 = decode_tile<BN, CL>(p, tile0, rank);
                if (p.up2) c0.n0 += c0.z1 * p.N;  // weight rows of the parity class
                const bool ragged0 = p.n_last > 0 && c0.n_tile == p.grid_n - 1;
                const CUtensorMap* mB0 = ragged0 ? &mapBL : &mapB;
                const uint32_t b0_bytes = ragged0 ? static_cast<uint32_t>(p.n_last / CL) * 128u : b_bytes;
                pre = min(STAGES, c0.kb_end - c0.kb_begin);
                for (int s = 0; s < pre; ++s) {
                    // pair mode: the leader's barrier counts the operand bytes of both CTAs
                    if (CL == 1 || rank == 0) mbar_arrive_expect_tx(&full_bar[s], CL * (a_bytes + b0_bytes));
                    uint8_t* sbp = smem + s * STAGE_BYTES + A_BYTES;
                    if (CL == 2) {
#pragma unroll
                        for (int j = 0; j < NSUB; ++j)
                            tma_load_4d_pair(sbp + j * (BNS / CL) * 128, mB0, &full_bar[s], (c0.kb_begin + s) * 64,
                                             NSUB == 1 ? c0.n0 + rank * ((ragged0 ? p.n_last : BN) / CL)
                                                       : c0.n0 + j * BNS + rank * (BNS / CL),
                                             0, 0);
                    }
                    else if (w_blocked)
                        tma_load_4d(sbp, &mapB, &full_bar[s], 0, 0, c0.kb_begin + s, c0.n0 >> 6);
                    else
                        tma_load_4d(sbp, mB0, &full_bar[s], (c0.kb_begin + s) * 64, c0.n0, 0, 0);
                }
            }
            pdl_wait();
            for (int wave = 0;; ++wave) {
                const int t = tile_of_wave(wave, tile0, tstep, total_tiles);
                if (t < 0) break;
                const TileCoord c = decode_tile<BN, CL>(p, t, rank);
                const int za = p.a_batched ? c.z1 : 0, za2 = p.a_batched ? c.z2 : 0;
                const int bz1 = p.b_batched ? c.z1 : 0, bz2 = p.b_batched ? c.z2 : 0;
                const bool ragged = p.n_last > 0 && c.n_tile == p.grid_n - 1;
                const CUtensorMap* mB = ragged ? &mapBL : &mapB;
                const uint32_t tile_b_bytes = ragged ? static_cast<uint32_t>(p.n_last / CL) * 128u : b_bytes;
                const int wrow0 = c.n0 + (p.up2 ? c.z1 * p.N : 0);  // first weight row of the tile
                const int brow = wrow0 + rank * ((ragged ? p.n_last : BN) / CL);
                for (int kb = c.kb_begin; kb < c.kb_end; ++kb) {
                    const bool prefetched = (wave == 0) && (kb - c.kb_begin) < pre;
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    if (!prefetched) {
                        mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
                        if (CL == 1 || rank == 0) mbar_arrive_expect_tx(&full_bar[stage], CL * (a_bytes + tile_b_bytes));
                    }
                    if (p.mode == 1) {
                        const int tap = kb / p.cblocks;
                        const int cb = kb - tap * p.cblocks;
                        int ky = tap / 3, kx = tap - ky * 3;
                        if (p.up2) {  // 2x2 taps of the parity class, offsets (s - 1 + parity)
                            ky = (tap >> 1) + (c.z1 >> 1);
                            kx = (tap & 1) + (c.z1 & 1);
                        }
                        if (p.down2) {
                            const int oy = ky - p.down_pad, ox = kx - p.down_pad;
                            const int py = oy & 1, px = ox & 1;
                            const CUtensorMap* mv = py ? (px ? &mapA3 : &mapA2) : (px ? &mapA1 : &mapA0);
                            tma4<CL>(sa, mv, &full_bar[stage], cb * 64, c.x0 + ((ox - px) >> 1), c.y0 + ((oy - py) >> 1), c.img0);
                        } else if (tap >= 9) {
                            // fused 1x1 shortcut: unshifted boxes of the shortcut's own sources
                            const int sb = kb - 9 * p.cblocks;
                            if (sb < p.sc_blocks0)
                                tma4<CL>(sa, &mapA2, &full_bar[stage], sb * 64, c.x0, c.y0, c.img0);
                            else
                                tma4<CL>(sa, &mapA3, &full_bar[stage], (sb - p.sc_blocks0) * 64, c.x0, c.y0, c.img0);
                        } else if (cb < p.cblocks0)
                            tma4<CL>(sa, &mapA0, &full_bar[stage], cb * 64, c.x0 + kx - 1, c.y0 + ky - 1, c.img0);
                        else
                            tma4<CL>(sa, &mapA1, &full_bar[stage], (cb - p.cblocks0) * 64, c.x0 + kx - 1, c.y0 + ky - 1,
                                     c.img0);
                    } else {
                        if (kb < p.cblocks0)
                            tma4<CL>(sa, &mapA0, &full_bar[stage], kb * 64, c.m_tile * 128, za, za2);
                        else
                            tma4<CL>(sa, &mapA1, &full_bar[stage], (kb - p.cblocks0) * 64, c.m_tile * 128, za, za2);
                    }
                    if (prefetched) {
                        // B tile of this stage is already in flight
                    } else if (CL == 2) {
#pragma unroll
                        for (int j = 0; j < NSUB; ++j)
                            tma_load_4d_pair(sb + j * (BNS / CL) * 128, mB, &full_bar[stage], kb * 64,
                                             NSUB == 1 ? brow : wrow0 + j * BNS + rank * (BNS / CL), 0, 0);
                    } else if (w_blocked) {
                        tma_load_4d(sb, &mapB, &full_bar[stage], 0, 0, kb, c.n0 >> 6);
                    } else if (!b_mn) {
                        tma_load_4d(sb, mB, &full_bar[stage], kb * 64, wrow0, bz1, bz2);
                    } else {
#pragma unroll
                        for (int j = 0; j < B_CHUNKS; ++j)
                            tma_load_4d(sb + j * 8192, &mapB, &full_bar[stage], c.n0 + j * 64, kb * 64, bz1, bz2);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (CL == 2 && rank == 1) {
            // peer CTA of a pair: its operands are consumed by the leader's MMAs (its TMA loads report to the leader's barriers)
        } else if (lane == 0) {
            // ------------------------------ MMA issuer (leader CTA in pair mode) ------------------------------
            const uint32_t idesc_full = umma_idesc_f16(128 * CL, BNS, 0, b_mn ? 1 : 0);
            const uint32_t idesc_last = umma_idesc_f16(128 * CL, p.n_last > 0 ? p.n_last : BN, 0, b_mn ? 1 : 0);
            pdl_wait();
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            bool first = true;
            for (int wave = 0;; ++wave) {
                const int t = tile_of_wave(wave, tile0, tstep, total_tiles);
                if (t < 0) break;
                const TileCoord c = decode_tile<BN, CL>(p, t, rank);
                const uint32_t idesc = (p.n_last > 0 && c.n_tile == p.grid_n - 1) ? idesc_last : idesc_full;
                mbar_wait_bounded(&tmem_empty_bar[acc], acc_phase ^ 1);  // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
                for (int kb = c.kb_begin; kb < c.kb_end; ++kb) {
                    mbar_wait_bounded(&full_bar[stage], phase);
                    if (first) {
                        DBG_MARK(2);
                        first = false;
                    }
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
                        if (CL == 2) {
#pragma unroll
                            for (int j = 0; j < NSUB; ++j)
                                umma_f16_2cta(tmem_d + static_cast<uint32_t>(j * BNS),
                                              ad, bd + static_cast<uint64_t>((j * (BNS / CL) * 128) >> 4), idesc,
                                              (kb > c.kb_begin || k > 0) ? 1u : 0u);
                        } else
                            umma_f16(tmem_d, ad, bd, idesc, (kb > c.kb_begin || k > 0) ? 1u : 0u);
                    }
                    if (CL == 2)
                        umma_commit2_mc(&empty_bar[stage], 3);  // frees the stage in both CTAs of the pair
                    else
                        umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (CL == 2)
                    umma_commit2_mc(&tmem_full_bar[acc], 3);  // wakes the epilogue warps of both CTAs
                else
                    umma_commit(&tmem_full_bar[acc]);
                if (++acc == NACC) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
            DBG_MARK(3);
        }
    } else {
        // ------------------------------ epilogue warps ------------------------------
        // eight epilogue warps: two per TMEM lane quarter (warp-id % 4), taking alternate 32-column chunks of the tile
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int et = threadIdx.x - 64;  // 0..255
        const int half = (warp - 2) >> 2;  // 0: even chunks, 1: odd chunks
        pdl_wait();
        int acc = 0;
        uint32_t acc_phase = 0;
        const bool col_bias = p.bias != nullptr && (p.flags & EPI_BIAS_M) == 0;
        const bool ln_on = FEAT != 0 && p.ln_stats != nullptr;
        const bool fused_reduce = p.splits > 1 && p.tile_counters != nullptr;
        // 32-byte aligned full-width fp16 rows (st.global.v8 / 16-byte residual loads) and an epilogue the fast path covers
        const bool fast_epi =
            OCC == 1 && p.splits == 1 && p.dbg_mode == 0 && (p.N & 31) == 0 && (p.ldc & 15) == 0 &&
            (reinterpret_cast<uintptr_t>(p.out) & 31) == 0 && ((p.out_zs1 | p.out_zs2) & 15) == 0 &&
            ((p.flags & ~(GEMM_B_MN | GEMM_W_BLOCKED | GEMM_HINT_CL2)) == 0
                 ? (p.residual == nullptr || ((p.ldr & 7) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0 &&
                                              ((p.res_zs1 | p.res_zs2) & 7) == 0))
                 : ((p.flags & ~(GEMM_B_MN | GEMM_W_BLOCKED | GEMM_HINT_CL2)) == EPI_GEGLU && p.residual == nullptr &&
                    (p.ldc & 31) == 0));
        for (int wave = 0;; ++wave) {
            const int t = tile_of_wave(wave, tile0, tstep, total_tiles);
            if (t < 0) break;
            const TileCoord c = decode_tile<BN, CL>(p, t, rank);
            const int row = tile_row(p, c, r);
            const long long out_off = static_cast<long long>(c.z1) * p.out_zs1 + static_cast<long long>(c.z2) * p.out_zs2;
            const long long res_off = static_cast<long long>(c.z1) * p.res_zs1 + static_cast<long long>(c.z2) * p.res_zs2;
            // row index in the statistics buffers of a folded LayerNorm (batch entries are blocks of M rows in z order)
            const long long grow = static_cast<long long>(c.z2 * p.nz1 + c.z1) * p.M + (row >= 0 ? row : 0);
            if (col_bias || ln_on) {
                epi_bar_sync();  // previous tile's readers are done with sbias
                const long long boff = static_cast<long long>(c.z1) * p.bias_zs1 + c.n0;
                for (int i = et; i < BN; i += 256) {
                    sbias[i] = (col_bias && c.n0 + i < p.N) ? __ldg(p.bias + boff + i) : 0.0f;
                    if (ln_on) scolsum[i] = (c.n0 + i < p.N) ? __ldg(p.ln_colsum + boff + i) : 0.0f;
                }
                epi_bar_sync();
            }
            // chunks this warp owns: cc = 32*half, 32*half + 64, ... (bounded by the tile width and by N)
            int n_mine = 0;
            for (int cc = 32 * half; cc < BN && c.n0 + cc < p.N; cc += 64) ++n_mine;
            const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
            if (fast_epi) {
                // ---- fast path (fp16 row-major output, bias, optional residual, or GEGLU; every chunk full and 16-byte
                // aligned): the residual of ALL chunks is requested before waiting for the accumulator, so its L2 latency
                // hides behind the mainloop, and the TMEM read of chunk i+1 is in flight while chunk i is processed.
                constexpr int IT = (BN + 63) / 64;
                constexpr int RD = IT < 2 ? IT : 2;  // residual chunks in flight (a ring: chunk it + RD is requested once chunk it is consumed)
                uint4 rres[RD][4];
                const bool has_res = p.residual != nullptr;
                const __half* rrow = has_res && row >= 0
                                         ? p.residual + res_off + static_cast<long long>(row) * p.ldr + c.n0 + 32 * half
                                         : nullptr;
#pragma unroll
                for (int it = 0; it < RD; ++it) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        rres[it][i] = (rrow != nullptr && it < n_mine) ? __ldg(reinterpret_cast<const uint4*>(rrow + 64 * it) + i)
                                                                       : make_uint4(0u, 0u, 0u, 0u);
                }
                float ln_r = 1.0f, ln_nm = 0.0f;
                if (ln_on && row >= 0) ln_row_coef(p, grow, ln_r, ln_nm);  // L2 latency hides behind the mainloop
                mbar_wait_bounded(&tmem_full_bar[acc], acc_phase);
                tc_fence_after();
                if (wave == 0 && et == 0) DBG_MARK(4);
                // wide tiles keep one chunk of scores in registers (register budget: 168 per thread), narrow ones two
                constexpr int RB = IT <= 2 ? 2 : 1;
                uint32_t raw[RB][32];
                if (n_mine > 0) tmem_ld_32x32(tmem_acc + static_cast<uint32_t>(32 * half), raw[0]);
#pragma unroll
                for (int it = 0; it < IT; ++it) {
                    if (it < n_mine) {
                        const int cc = 32 * half + 64 * it;
                        tmem_ld_wait();
                        if (RB == 2 && it + 1 < IT && it + 1 < n_mine)
                            tmem_ld_32x32(tmem_acc + static_cast<uint32_t>(cc + 64), raw[(it + 1) % RB]);
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[it % RB][i]);
                        if (RB == 1 && it + 1 < IT && it + 1 < n_mine)
                            tmem_ld_32x32(tmem_acc + static_cast<uint32_t>(cc + 64), raw[0]);
                        const float* bc = sbias + cc;
                        if (ln_on) {
                            const float* cs = scolsum + cc;
                            const f32x2 r2 = pk2(ln_r), nm2 = pk2(ln_nm);
#pragma unroll
                            for (int i = 0; i < 32; i += 2)
                                upk2(fma2(pk2(v[i], v[i + 1]), r2, fma2(nm2, pk2(cs[i], cs[i + 1]), pk2(bc[i], bc[i + 1]))), v[i],
                                     v[i + 1]);
                        } else if (col_bias) {
                            const f32x2 al2 = pk2(p.alpha);
#pragma unroll
                            for (int i = 0; i < 32; i += 2)
                                upk2(fma2(pk2(v[i], v[i + 1]), al2, pk2(bc[i], bc[i + 1])), v[i], v[i + 1]);
                        } else {
                            const f32x2 al2 = pk2(p.alpha);
#pragma unroll
                            for (int i = 0; i < 32; i += 2) upk2(mul2(pk2(v[i], v[i + 1]), al2), v[i], v[i + 1]);
                        }
                        if (row >= 0) {
                            if (p.flags & EPI_GEGLU) {
                                uint32_t o[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    float o0, o1;
                                    upk2(geglu2(pk2(v[2 * i], v[2 * i + 1]), pk2(v[16 + 2 * i], v[17 + 2 * i])), o0, o1);
                                    o[i] = pack_half2(o0, o1);
                                }
                                __half* out = reinterpret_cast<__half*>(p.out) + out_off + static_cast<long long>(row) * p.ldc +
                                              ((c.n0 + cc) >> 1);
                                st_global_v8(out, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
                            } else {
                                if (has_res) {
#pragma unroll
                                    for (int i = 0; i < 4; ++i) {
                                        const __half2* h2 = reinterpret_cast<const __half2*>(&rres[it % RD][i]);
#pragma unroll
                                        for (int j = 0; j < 4; ++j) {
                                            const float2 f = __half22float2(h2[j]);
                                            upk2(add2(pk2(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]), pk2(f.x, f.y)), v[8 * i + 2 * j],
                                                 v[8 * i + 2 * j + 1]);
                                        }
                                    }
                                    if (it + RD < IT && it + RD < n_mine) {
#pragma unroll
                                        for (int i = 0; i < 4; ++i)
                                            rres[it % RD][i] = __ldg(reinterpret_cast<const uint4*>(rrow + 64 * (it + RD)) + i);
                                    }
                                }
                                if (FEAT != 0 && p.stats_out != nullptr) emit_row_stats(p, grow, c.n0 + cc, v);
                                __half* out = reinterpret_cast<__half*>(p.out) + out_off + static_cast<long long>(row) * p.ldc +
                                              c.n0 + cc;
#pragma unroll
                                for (int i = 0; i < 2; ++i)
                                    st_global_v8(out + 16 * i, pack_half2(v[16 * i + 0], v[16 * i + 1]),
                                                 pack_half2(v[16 * i + 2], v[16 * i + 3]), pack_half2(v[16 * i + 4], v[16 * i + 5]),
                                                 pack_half2(v[16 * i + 6], v[16 * i + 7]), pack_half2(v[16 * i + 8], v[16 * i + 9]),
                                                 pack_half2(v[16 * i + 10], v[16 * i + 11]), pack_half2(v[16 * i + 12], v[16 * i + 13]),
                                                 pack_half2(v[16 * i + 14], v[16 * i + 15]));
                            }
                        }
                    }
                }
                // every TMEM read of this warp has completed (the last tmem_ld_wait above): hand the accumulator back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CL == 2 && rank == 1)
                        mbar_arrive_remote(&tmem_empty_bar[acc], 0);
                    else
                        mbar_arrive(&tmem_empty_bar[acc]);
                }
                if (++acc == NACC) {
                    acc = 0;
                    acc_phase ^= 1;
                }
                continue;
            }
            mbar_wait_bounded(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            if (wave == 0 && et == 0) DBG_MARK(4);
            if (n_mine == 0) {
                // nothing to read for this warp in this tile: still hand the accumulator back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CL == 2 && rank == 1)
                        mbar_arrive_remote(&tmem_empty_bar[acc], 0);
                    else
                        mbar_arrive(&tmem_empty_bar[acc]);
                }
            }
#pragma unroll 1
            for (int cc = 32 * half, it = 0; it < n_mine; cc += 64, ++it) {
                uint32_t raw[32];
                if (p.dbg_mode & 2) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) raw[i] = 0;
                } else {
                    tmem_ld_32x32(tmem_acc + static_cast<uint32_t>(cc), raw);
                    tmem_ld_wait();
                }
                if (it == n_mine - 1) {
                    // last chunk of this tile is in registers: hand the accumulator back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (CL == 2 && rank == 1)
                            mbar_arrive_remote(&tmem_empty_bar[acc], 0);
                        else
                            mbar_arrive(&tmem_empty_bar[acc]);
                    }
                }
                if (fused_reduce) {
                    // tile-contiguous partial [group][split][128 rows][BN] (padding rows are never read back into an output)
                    if (row >= 0) {
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                        const long long widx = ((static_cast<long long>(tile_group(p, c)) * p.splits + c.zsplit) * 128 + r) * BN + cc;
                        if (p.ws_half)
                            store_half_chunk(reinterpret_cast<__half*>(p.workspace) + widx, v, 32, true);
                        else
                            store_f32_chunk(p.workspace + widx, v, 32, true);
                    }
                } else if (row >= 0 && !((p.dbg_mode & 1) && raw[0] != 0x7fc01234u)) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                    if (p.splits > 1) {
                        float* ws = p.workspace +
                                    (static_cast<long long>(c.zb * p.splits + c.zsplit) * p.M + row) *
                                        static_cast<long long>(p.N) +
                                    c.n0 + cc;
                        const int ncols = min(32, p.N - (c.n0 + cc));
                        store_f32_chunk(ws, v, ncols, (p.N & 3) == 0);
                    } else {
                        epilogue_store32(p, out_off, res_off, row, c.n0 + cc, v, col_bias ? sbias + cc : nullptr, grow,
                                         scolsum + cc, FEAT != 0);
                    }
                }
            }
            if (fused_reduce) {
                // Split-K without a second launch (host guarantees one tile per CTA and a fully co-resident grid, so waiting
                // for the other CTAs of the split group cannot deadlock): every CTA publishes its fp32 partial, waits until
                // the group's `splits` partials are visible, then reduces ITS row slice of the tile in split order
                // (bit-reproducible) and runs the epilogue for that slice.
                int* cnt = p.tile_counters + 2 * tile_group(p, c);
                const int r0 = (c.zsplit * 128) / p.splits, r1 = ((c.zsplit + 1) * 128) / p.splits;
                const uint32_t esz = p.ws_half ? 2u : 4u;  // partials in fp16 halve the L2 write burst and the gather
                const uint32_t slice_bytes = static_cast<uint32_t>(r1 - r0) * BN * esz;
                const uint8_t* wsg = reinterpret_cast<const uint8_t*>(p.workspace) +
                                     static_cast<long long>(tile_group(p, c)) * p.splits * 128 * BN * esz;
                // operand stages are idle (this CTA has no further tile): [splits][rows of my slice][BN] staging, then the sums
                float* stage_f = reinterpret_cast<float*>(smem);
                float* ssum = reinterpret_cast<float*>(smem + static_cast<size_t>(p.splits) * slice_bytes);
                if (et == 0) DBG_MARK2(0);
                fence_proxy_async_all();  // my partial is read by the peers' bulk copies (async proxy)
                __threadfence();
                if (et == 0) DBG_MARK2(1);
                epi_bar_sync();
                if (et == 0) {
                    DBG_MARK2(2);
                    atomicAdd(cnt, 1);
                    const long long t0 = clock64();
                    while (ld_acquire_gpu(cnt) < p.splits) {
                        if (clock64() - t0 > 4000000000LL) {  // a CTA of the split group never became resident
                            if (p.err_flag != nullptr) *p.err_flag = 2;
                            break;
                        }
                    }
                    DBG_MARK2(3);
                    fence_proxy_async_all();
                    // all of the group's slices in flight at once: one bulk copy per split
                    mbar_arrive_expect_tx(red_bar, static_cast<uint32_t>(p.splits) * slice_bytes);
                    for (int sidx = 0; sidx < p.splits; ++sidx)
                        bulk_load_1d(reinterpret_cast<uint8_t*>(stage_f) + static_cast<size_t>(sidx) * slice_bytes,
                                     wsg + (static_cast<long long>(sidx) * 128 + r0) * BN * esz, slice_bytes, red_bar);
                }
                mbar_wait_bounded(red_bar, 0);
                if (et == 0) DBG_MARK2(4);
                constexpr int P4 = BN / 4;  // 16-byte pieces per tile row
                const int npieces = (r1 - r0) * P4;
                for (int i = et; i < npieces; i += 256) {
                    float4 a4;
                    if (p.ws_half) {
                        const uint2* src = reinterpret_cast<const uint2*>(stage_f) + i;
                        a4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        for (int sidx = 0; sidx < p.splits; ++sidx) {
                            const uint2 t = src[static_cast<size_t>(sidx) * npieces];
                            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
                            const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
                            a4.x += lo.x;
                            a4.y += lo.y;
                            a4.z += hi.x;
                            a4.w += hi.y;
                        }
                    } else {
                        const float4* src = reinterpret_cast<const float4*>(stage_f) + i;
                        a4 = src[0];
                        for (int sidx = 1; sidx < p.splits; ++sidx) {
                            const float4 t4 = src[static_cast<size_t>(sidx) * npieces];
                            a4.x += t4.x;
                            a4.y += t4.y;
                            a4.z += t4.z;
                            a4.w += t4.w;
                        }
                    }
                    // unit u = (row, 32-column chunk) occupies 128 B; its 16-B pieces are XOR-swizzled by the unit index
                    const int rl = i / P4, c4 = i - rl * P4;
                    const int u = rl * (BN / 32) + (c4 >> 3);
                    *reinterpret_cast<float4*>(ssum + u * 32 + (((c4 & 7) ^ (u & 7)) << 2)) = a4;
                }
                epi_bar_sync();
                if (et == 0) {
                    DBG_MARK2(5);
                    // all CTAs of the group have passed their wait before the last one gets here: safe to clear both slots
                    if (atomicAdd(cnt + 1, 1) == p.splits - 1) {
                        cnt[0] = 0;
                        cnt[1] = 0;
                    }
                }
                for (int u = et; u < (r1 - r0) * (BN / 32); u += 256) {
                    const int rl = u / (BN / 32), ch = u - rl * (BN / 32);
                    const int orow = tile_row(p, c, r0 + rl);
                    const int col0 = c.n0 + ch * 32;
                    if (orow < 0 || col0 >= p.N) continue;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 t4 = *reinterpret_cast<const float4*>(ssum + u * 32 + ((j ^ (u & 7)) << 2));
                        v[4 * j] = t4.x;
                        v[4 * j + 1] = t4.y;
                        v[4 * j + 2] = t4.z;
                        v[4 * j + 3] = t4.w;
                    }
                    epilogue_store32(p, out_off, res_off, orow, col0, v, col_bias ? sbias + ch * 32 : nullptr,
                                     static_cast<long long>(c.z2 * p.nz1 + c.z1) * p.M + orow, scolsum + ch * 32, FEAT != 0);
                }
            }
            if (++acc == NACC) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (et == 0) DBG_MARK(5);
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();  // no CTA may exit while its peer can still arrive on its barriers / read its operands
    if (warp == 1) {
        if (CL == 2)
            tmem_dealloc2(tmem_base, TM_COLS);
        else
            tmem_dealloc(tmem_base, TM_COLS);
    }
    if (threadIdx.x == 32) DBG_MARK(6);
}

// sums split-K partials and runs the epilogue; one thread per (row, 32-column chunk)
__global__ void __launch_bounds__(256) gemm_splitk_finalize_kernel(const __grid_constant__ GemmParams p) {
    pdl_enter();
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
        if (ncols == 32 && (p.N & 3) == 0 && aligned32(ws)) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t t[8];
                ld_global_nc_v8(ws + 8 * i, t);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[8 * i + j] += __uint_as_float(t[j]);
            }
        } else if (ncols == 32 && (p.N & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(ws) + i);
                v[4 * i] += t.x;
                v[4 * i + 1] += t.y;
                v[4 * i + 2] += t.z;
                v[4 * i + 3] += t.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < ncols) v[i] += ws[i];
        }
    }
    float b[32], cs[32];
    const bool col_bias = p.bias != nullptr && (p.flags & EPI_BIAS_M) == 0;
    const int z1 = zb % p.nz1, z2 = zb / p.nz1;
    const long long boff = static_cast<long long>(z1) * p.bias_zs1 + col0;
    if (col_bias) {
#pragma unroll
        for (int i = 0; i < 32; ++i) b[i] = (i < ncols) ? __ldg(p.bias + boff + i) : 0.0f;
    }
    if (p.ln_stats != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; ++i) cs[i] = (i < ncols) ? __ldg(p.ln_colsum + boff + i) : 0.0f;
    }
    const long long out_off = static_cast<long long>(z1) * p.out_zs1 + static_cast<long long>(z2) * p.out_zs2;
    const long long res_off = static_cast<long long>(z1) * p.res_zs1 + static_cast<long long>(z2) * p.res_zs2;
    epilogue_store32(p, out_off, res_off, row, col0, v, col_bias ? b : nullptr, static_cast<long long>(zb) * p.M + row, cs);
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

bool gemm_cluster_enabled() {
    // CTA-pair (cta_group::2) mode is functionally verified (all operator and pipeline parity tests pass with
    // DTP_CLUSTER=1) but measured ~2x slower per k-block than the single-CTA mainloop in round 1 (profiles/README.md,
    // item 10), so it is opt-in until that is understood.
    static const bool on = []() {
        const char* e = getenv("DTP_CLUSTER");
        return e && e[0] == '1';
    }();
    return on;
}

static void bind_ctx(GemmOp* op);
static void params_defaults(GemmParams& p) {
    memset(&p, 0, sizeof(p));
    p.splits = 1;
    p.nz1 = 1;
    p.nz2 = 1;
    p.alpha = 1.0f;
}

static int fix_bn(int BN) {
    return (BN == 32 || BN == 64 || BN == 128 || BN == 160 || BN == 192 || BN == 256 || BN == 320) ? BN : 128;
}
// BN = 320 exists for CTA pairs on problems whose N is a multiple of 320 (the UNet widths); anything else runs 256-wide
static int legal_bn(int BN, bool pair_possible, int N) { return (BN == 320 && !(pair_possible && N % 320 == 0)) ? 256 : BN; }

// 2-D (K inner, rows outer) map expressed as 4-D with unit batch dims
static int map_rows(CUtensorMap* m, const __half* base, uint64_t K, uint64_t rows, uint64_t ld, uint32_t box_rows) {
    uint64_t dims[4] = {K, rows, 1, 1};
    uint64_t st[3] = {ld * 2, ld * 2 * rows, ld * 2 * rows};
    uint32_t box[4] = {64, box_rows, 1, 1};
    return make_map_4d(m, base, dims, st, box);
}

// weights in [N/64][K/64][64][64] tiles: dims (k_in, n_in, k_blk, n_blk)
static int map_blocked(CUtensorMap* m, const __half* base, uint64_t K, uint64_t N, uint32_t BN) {
    uint64_t dims[4] = {64, 64, K / 64, (N + 63) / 64};
    uint64_t st[3] = {128, 8192, 8192 * (K / 64)};
    uint32_t box[4] = {64, 64, 1, BN / 64};
    return make_map_4d(m, base, dims, st, box);
}

// Ragged last n-tile (N not a multiple of BN): its own weight box of n_last (pair mode: n_last / 2) rows
static int setup_ragged(GemmOp* op, const __half* Wt, uint64_t K, uint64_t N, uint64_t ldw, int BN) {
    GemmParams& p = op->p;
    p.n_last = 0;
    op->mapBL = (op->cluster == 2) ? op->mapBh : op->mapB;
    const int rows_last = static_cast<int>(N % BN);
    if (rows_last == 0) return 0;
    const int n_last = (rows_last + 15) & ~15;
    if (n_last >= BN) return 0;
    if (map_rows(&op->mapBL, Wt, K, N, ldw, static_cast<uint32_t>(op->cluster == 2 ? n_last / 2 : n_last))) return -14;
    p.n_last = n_last;
    return 0;
}

int gemm_setup_linear(GemmOp* op, const __half* A0, int lda0, int K0, const __half* A1, int lda1, int K1, int M,
                      const __half* Wt, int ldw, int N, int BN, int splits, int w_blocked) {
    params_defaults(op->p);
    GemmParams& p = op->p;
    const bool want_pair = (BN & GEMM_BN_PAIR) != 0 || gemm_cluster_enabled();
    BN = legal_bn(fix_bn(BN & ~GEMM_BN_PAIR), want_pair && !w_blocked && M > 128, N);
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
    bind_ctx(op);
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
    op->mapA2 = op->mapA0;
    op->mapA3 = op->mapA0;
    op->cluster = 1;
    if (!w_blocked && want_pair && op->grid_m >= 2 && BN >= 32) {
        if (map_rows(&op->mapBh, Wt, K0 + (A1 ? K1 : 0), N, ldw, BN > 256 ? BN / 4 : BN / 2)) return -13;
        op->cluster = 2;
    }
    if (w_blocked) {
        const int K = K0 + (A1 ? K1 : 0);
        if ((K % 64) != 0 || (N % 64) != 0 || (BN % 64) != 0) {
            snprintf(g_gemm_err, sizeof(g_gemm_err), "blocked weights need K, N, BN %% 64 == 0 (K=%d N=%d BN=%d)", K, N, BN);
            return -12;
        }
        p.flags |= GEMM_W_BLOCKED;
        r = map_blocked(&op->mapB, Wt, K, N, BN);
        op->mapBL = op->mapB;
        return r;
    }
    if (BN > 256)
        op->mapB = op->mapBh;  // pair-only tile width: the single-CTA box does not exist (and would exceed 256 rows)
    else
        r = map_rows(&op->mapB, Wt, K0 + (A1 ? K1 : 0), N, ldw, BN);
    if (r) return r;
    return setup_ragged(op, Wt, K0 + (A1 ? K1 : 0), N, ldw, BN);
}

static int largest_divisor_le(int n, int cap) {
    int best = 1;
    for (int d = 1; d <= n && d <= cap; ++d)
        if (n % d == 0) best = d;
    return best;
}

int gemm_setup_conv3x3(GemmOp* op, const __half* A0, int C0, const __half* A1, int C1, int Nimg, int H, int W,
                       const __half* Wt, int Cout, int BN, int splits, int w_blocked, const __half* S0, int CS0,
                       const __half* S1, int CS1) {
    params_defaults(op->p);
    GemmParams& p = op->p;
    const bool want_pair = (BN & GEMM_BN_PAIR) != 0 || gemm_cluster_enabled();
    BN = fix_bn(BN & ~GEMM_BN_PAIR);
    if (BN == 320 && !(want_pair && !w_blocked && Cout % 320 == 0)) BN = 256;  // (single m-tile problems: checked below)
    if ((C0 % 64) != 0 || (A1 != nullptr && (C1 % 64) != 0)) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "conv3x3 needs channel counts %% 64 == 0 (C0=%d C1=%d)", C0, C1);
        return -11;
    }
    const int C = C0 + (A1 ? C1 : 0);
    if (S0 == nullptr) CS0 = 0;
    if (S1 == nullptr) CS1 = 0;
    if ((CS0 % 64) != 0 || (CS1 % 64) != 0 || (S1 != nullptr && S0 == nullptr) || ((S0 != nullptr) && w_blocked)) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "conv3x3 + shortcut needs shortcut channel counts %% 64 == 0 (CS0=%d CS1=%d)", CS0, CS1);
        return -11;
    }
    const int KT = 9 * C + CS0 + CS1;  // weight row length
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
    p.sc_blocks0 = CS0 / 64;
    p.sc_blocks = (CS0 + CS1) / 64;
    p.num_kb = 9 * p.cblocks + p.sc_blocks;
    p.splits = splits < 1 ? 1 : (splits > p.num_kb ? p.num_kb : splits);
    bind_ctx(op);
    p.ldc = Cout;
    op->BN = BN;
    op->grid_m = p.tiles_x * p.tiles_y * ((Nimg + p.bn - 1) / p.bn);
    if (BN == 320 && op->grid_m < 2) {
        BN = 256;
        op->BN = 256;
    }
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
    op->mapA2 = op->mapA0;
    op->mapA3 = op->mapA0;
    const __half* srcs[2] = {S0, S1};
    const int cs[2] = {CS0, CS1};
    for (int i = 0; i < 2; ++i) {
        if (srcs[i] == nullptr) continue;
        uint64_t dims[4] = {(uint64_t)cs[i], (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        uint64_t st[3] = {(uint64_t)cs[i] * 2, (uint64_t)cs[i] * 2 * W, (uint64_t)cs[i] * 2 * W * H};
        uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
        int r = make_map_4d(i == 0 ? &op->mapA2 : &op->mapA3, srcs[i], dims, st, box);
        if (r) return r;
    }
    op->cluster = 1;
    if (!w_blocked && want_pair && op->grid_m >= 2 && BN >= 32) {
        if (map_rows(&op->mapBh, Wt, (uint64_t)KT, Cout, (uint64_t)KT, BN > 256 ? BN / 4 : BN / 2)) return -13;
        op->cluster = 2;
    }
    if (w_blocked) {
        if ((Cout % 64) != 0 || (BN % 64) != 0) {
            snprintf(g_gemm_err, sizeof(g_gemm_err), "blocked conv weights need Cout, BN %% 64 == 0 (Cout=%d BN=%d)", Cout, BN);
            return -12;
        }
        p.flags |= GEMM_W_BLOCKED;
        const int rb = map_blocked(&op->mapB, Wt, (uint64_t)9 * C, Cout, BN);
        op->mapBL = op->mapB;
        return rb;
    }
    int rm = 0;
    if (BN > 256)
        op->mapB = op->mapBh;
    else
        rm = map_rows(&op->mapB, Wt, (uint64_t)KT, Cout, (uint64_t)KT, BN);
    if (rm) return rm;
    return setup_ragged(op, Wt, (uint64_t)KT, Cout, (uint64_t)KT, BN);
}

int gemm_setup_conv3x3_s2(GemmOp* op, const __half* A, int C, int Nimg, int H, int W, const __half* Wt, int Cout, int pad_lo,
                          int BN, int splits) {
    params_defaults(op->p);
    GemmParams& p = op->p;
    const bool want_pair = (BN & GEMM_BN_PAIR) != 0 || gemm_cluster_enabled();
    BN = fix_bn(BN & ~GEMM_BN_PAIR);
    if (BN == 320 && !(want_pair && Cout % 320 == 0)) BN = 256;
    if ((C % 64) != 0 || (H & 1) || (W & 1)) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "conv3x3 stride 2 needs C %% 64 == 0 and even H, W (C=%d H=%d W=%d)", C, H, W);
        return -11;
    }
    const int Ho = H / 2, Wo = W / 2;
    const int KT = 9 * C;
    p.M = Nimg * Ho * Wo;
    p.N = Cout;
    p.mode = 1;
    p.down2 = 1;
    p.down_pad = pad_lo ? 1 : 0;
    p.H = Ho;
    p.W = Wo;
    p.Nimg = Nimg;
    p.bw = largest_divisor_le(Wo, 128);
    p.bh = largest_divisor_le(Ho, 128 / p.bw);
    p.bn = (p.bh == Ho) ? (128 / (p.bw * p.bh)) : 1;
    if (p.bn > Nimg) p.bn = Nimg;
    if (p.bn < 1) p.bn = 1;
    p.tiles_x = Wo / p.bw;
    p.tiles_y = Ho / p.bh;
    p.rows_valid = p.bw * p.bh * p.bn;
    p.cblocks0 = C / 64;
    p.cblocks = C / 64;
    p.num_kb = 9 * p.cblocks;
    p.splits = splits < 1 ? 1 : (splits > p.num_kb ? p.num_kb : splits);
    bind_ctx(op);
    p.ldc = Cout;
    op->BN = BN;
    op->grid_m = p.tiles_x * p.tiles_y * ((Nimg + p.bn - 1) / p.bn);
    if (BN == 320 && op->grid_m < 2) {
        BN = 256;
        op->BN = 256;
    }
    CUtensorMap* views[4] = {&op->mapA0, &op->mapA1, &op->mapA2, &op->mapA3};
    for (int v = 0; v < 4; ++v) {  // view (py, px): pixels (2y + py, 2x + px)
        const int py = v >> 1, px = v & 1;
        uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)Nimg};
        uint64_t st[3] = {(uint64_t)C * 4, (uint64_t)C * 4 * W, (uint64_t)C * 2 * W * H};
        uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
        int r = make_map_4d(views[v], A + (static_cast<size_t>(py) * W + px) * C, dims, st, box);
        if (r) return r;
    }
    op->cluster = 1;
    if (want_pair && op->grid_m >= 2 && BN >= 32) {
        if (map_rows(&op->mapBh, Wt, (uint64_t)KT, Cout, (uint64_t)KT, BN > 256 ? BN / 4 : BN / 2)) return -13;
        op->cluster = 2;
    }
    int rm = 0;
    if (BN > 256)
        op->mapB = op->mapBh;
    else
        rm = map_rows(&op->mapB, Wt, (uint64_t)KT, Cout, (uint64_t)KT, BN);
    if (rm) return rm;
    return setup_ragged(op, Wt, (uint64_t)KT, Cout, (uint64_t)KT, BN);
}

int gemm_setup_upconv2x(GemmOp* op, const __half* A, int C, int Nimg, int H, int W, const __half* Wstack, int Cout, int BN) {
    params_defaults(op->p);
    GemmParams& p = op->p;
    const bool want_pair = (BN & GEMM_BN_PAIR) != 0 || gemm_cluster_enabled();
    BN = fix_bn(BN & ~GEMM_BN_PAIR);
    if (BN == 320 && !(want_pair && Cout % 320 == 0)) BN = 256;
    if ((C % 64) != 0) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "upconv2x needs C %% 64 == 0 (C=%d)", C);
        return -11;
    }
    if ((Cout % BN) != 0 && (Cout % 16) != 0) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "upconv2x needs Cout %% 16 == 0 (Cout=%d)", Cout);
        return -11;
    }
    const int KT = 4 * C;
    p.M = Nimg * H * W;  // rows per parity class
    p.N = Cout;
    p.mode = 1;
    p.up2 = 1;
    p.nz1 = 4;
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
    p.cblocks0 = C / 64;
    p.cblocks = C / 64;
    p.num_kb = 4 * p.cblocks;
    p.splits = 1;
    bind_ctx(op);
    p.ldc = Cout;
    op->BN = BN;
    op->grid_m = p.tiles_x * p.tiles_y * ((Nimg + p.bn - 1) / p.bn);
    if (BN == 320 && op->grid_m < 2) {
        BN = 256;
        op->BN = 256;
    }
    {
        uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        uint64_t st[3] = {(uint64_t)C * 2, (uint64_t)C * 2 * W, (uint64_t)C * 2 * W * H};
        uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
        int r = make_map_4d(&op->mapA0, A, dims, st, box);
        if (r) return r;
    }
    op->mapA1 = op->mapA0;
    op->mapA2 = op->mapA0;
    op->mapA3 = op->mapA0;
    op->cluster = 1;
    const uint64_t wrows = 4ull * Cout;
    if (want_pair && op->grid_m >= 2 && BN >= 32) {
        if (map_rows(&op->mapBh, Wstack, (uint64_t)KT, wrows, (uint64_t)KT, BN > 256 ? BN / 4 : BN / 2)) return -13;
        op->cluster = 2;
    }
    int rm = 0;
    if (BN > 256)
        op->mapB = op->mapBh;
    else
        rm = map_rows(&op->mapB, Wstack, (uint64_t)KT, wrows, (uint64_t)KT, BN);
    if (rm) return rm;
    // ragged last n-tile: its box may run into the next class's rows; those columns are >= N and never stored
    p.n_last = 0;
    op->mapBL = (op->cluster == 2) ? op->mapBh : op->mapB;
    const int rows_last = Cout % BN;
    if (rows_last != 0) {
        const int n_last = (rows_last + 15) & ~15;
        if (n_last < BN) {
            if (map_rows(&op->mapBL, Wstack, (uint64_t)KT, wrows, (uint64_t)KT, static_cast<uint32_t>(op->cluster == 2 ? n_last / 2 : n_last)))
                return -14;
            p.n_last = n_last;
        }
    }
    return 0;
}

int gemm_setup_batched(GemmOp* op, const __half* A, int lda, long long a_zs1, long long a_zs2, const __half* B, int ldb,
                       long long b_zs1, long long b_zs2, int b_mn, int M, int N, int K, int nz1, int nz2, int BN) {
    params_defaults(op->p);
    GemmParams& p = op->p;
    BN = fix_bn(BN & ~GEMM_BN_PAIR);
    if (BN == 320) BN = 256;
    if (b_mn && BN < 64) BN = 64;
    if (b_mn && BN == 160) BN = 192;  // MN-major B is loaded in 64-column boxes
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
    op->cluster = 1;
    bind_ctx(op);
    // size-1 batch dims still need a legal (multiple of 16 B) stride
    auto zstride = [](long long s, uint64_t fallback) -> uint64_t { return s > 0 ? (uint64_t)s * 2 : fallback; };
    {
        uint64_t dims[4] = {(uint64_t)K, (uint64_t)M, (uint64_t)nz1, (uint64_t)nz2};
        uint64_t st[3] = {(uint64_t)lda * 2, zstride(a_zs1, (uint64_t)lda * 2 * M), zstride(a_zs2, (uint64_t)lda * 2 * M)};
        uint32_t box[4] = {64, 128, 1, 1};
        int r = make_map_4d(&op->mapA0, A, dims, st, box);
        if (r) return r;
        op->mapA1 = op->mapA0;
        op->mapA2 = op->mapA0;
        op->mapA3 = op->mapA0;
    }
    if (!b_mn) {
        uint64_t dims[4] = {(uint64_t)K, (uint64_t)N, (uint64_t)nz1, (uint64_t)nz2};
        uint64_t st[3] = {(uint64_t)ldb * 2, zstride(b_zs1, (uint64_t)ldb * 2 * N), zstride(b_zs2, (uint64_t)ldb * 2 * N)};
        uint32_t box[4] = {64, (uint32_t)BN, 1, 1};
        const int rk = make_map_4d(&op->mapB, B, dims, st, box);
        op->mapBL = op->mapB;
        return rk;
    } else {
        uint64_t dims[4] = {(uint64_t)N, (uint64_t)K, (uint64_t)nz1, (uint64_t)nz2};
        uint64_t st[3] = {(uint64_t)ldb * 2, zstride(b_zs1, (uint64_t)ldb * 2 * K), zstride(b_zs2, (uint64_t)ldb * 2 * K)};
        uint32_t box[4] = {64, 64, 1, 1};
        const int rk = make_map_4d(&op->mapB, B, dims, st, box);
        op->mapBL = op->mapB;
        return rk;
    }
}

size_t gemm_workspace_bytes(const GemmOp* op) {
    const GemmParams& p = op->p;
    if (p.splits <= 1) return 0;
    // the larger of the two partial layouts: row-major [batch*splits][M][N] (separate finalize kernel) and tile-contiguous
    // [tile group][split][128][BN] (in-kernel reduction)
    const size_t flat = static_cast<size_t>(p.nz1) * p.nz2 * p.splits * p.M * static_cast<size_t>(p.N);
    const size_t tiled = static_cast<size_t>(p.nz1) * p.nz2 * p.splits * op->grid_m * ((p.N + op->BN - 1) / op->BN) * 128 *
                         static_cast<size_t>(op->BN);
    return (flat > tiled ? flat : tiled) * sizeof(float);
}

// Measured configurations for the problem shapes of the benchmark workloads (profiles/make_gemm_table.py). DTP_GEMM_TABLE=0
// falls back to the cost model everywhere.
struct TunedCfg {
    int mtiles, N, num_kb, geglu, BN, splits;
};
static const TunedCfg g_tuned[] = {
#include "gemm_tuned.inc"
};
static bool tuned_lookup(int mtiles, int N, int num_kb, int flags, int* BN, int* splits) {
    static const int on = []() {
        const char* e = getenv("DTP_GEMM_TABLE");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    if (!on || (flags & (GEMM_B_MN | GEMM_W_BLOCKED | EPI_SOFTMAX16 | GEMM_HINT_CL2))) return false;
    const int geglu = (flags & EPI_GEGLU) ? 1 : 0;
    for (const TunedCfg& t : g_tuned)
        if (t.mtiles == mtiles && t.N == N && t.num_kb == num_kb && t.geglu == geglu) {
            *BN = t.BN;
            *splits = t.splits;
            return true;
        }
    return false;
}

bool gemm_tuned_config(int mtiles, int N, int num_kb, int flags, int* BN, int* splits) {
    return tuned_lookup(mtiles, N, num_kb, flags, BN, splits);
}

void gemm_pick_config(int mtiles, int N, int num_kb, int flags, int* BN, int* splits) {
    if (tuned_lookup(mtiles, N, num_kb, flags, BN, splits)) return;
    // Cost model (SM cycles) of the persistent kernel, searched over (BN, split-K):
    //   per k-block the tensor pipe needs 2*BN cycles (128 x BN x 64 MACs at 4096 MAC/clk/SM) and the TMA feed
    //   (16 KB of A + 128*BN B of B) moves ~64 B/clk per SM (measured); chip-wide the L2 -> SM fabric sustains
    //   ~7000 B/clk (measured ~14 TB/s on an 8192^3 problem), which bounds kernels whose tiles re-read A / B many times;
    //   the epilogue of a tile (~510 clk per 32 columns) overlaps the next tile's mainloop;
    //   split-K adds an fp32 partial write + a finalize pass (launch ~4900 clk + bytes at ~4200 B/clk).
    // Constants fitted to the CUDA-graph-timed sweep in profiles/gemm_sweep_r1.csv (regret 2.2 % over its 62 shapes).
    const int kSMs = 148;
    static const int cand_k[] = {32, 64, 128, 160, 192, 256};
    static const int cand_mn[] = {64, 128, 192, 256};
    const bool mn = (flags & (GEMM_B_MN | GEMM_W_BLOCKED)) != 0;
    const int* cand = mn ? cand_mn : cand_k;
    const int ncand = mn ? 4 : 6;
    double best_cost = -1.0;
    int best = cand[ncand - 1], best_sp = 1;
    const double rows = 128.0 * mtiles;
    for (int i = 0; i < ncand; ++i) {
        const int bn = cand[i];
        if (bn >= 64 && bn - 32 >= ((N + 31) / 32) * 32) continue;  // mostly padding
        const long long gn = (N + bn - 1) / bn;
        const double t_mma = 2.0 * bn;
        const double t_tma = (16384.0 + ((flags & GEMM_HINT_CL2) ? 64.0 : 128.0) * bn) / 61.0;
        const double t_kb = t_mma > t_tma ? t_mma : t_tma;
        const double t_epi = (bn / 32) * 510.0;
        const int max_sp = (flags & GEMM_B_MN) ? 1 : 16;
        for (int sp = 1; sp <= max_sp; ++sp) {
            if (sp > 1 && num_kb / sp < 4) break;
            const long long tiles = static_cast<long long>(mtiles) * gn * sp;
            const double kb_per = static_cast<double>(num_kb) / sp;
            const double t_main = kb_per * t_kb;
            const double t_epi_eff = sp > 1 ? 1800.0 : t_epi;
            const double t_tile = (t_main > t_epi_eff ? t_main : t_epi_eff) + 800.0;
            const double per_cta = static_cast<double>((tiles + kSMs - 1) / kSMs);
            double cost = per_cta * t_tile + t_epi_eff + 7900.0;
            const double l2_bytes = static_cast<double>(mtiles) * gn * num_kb *
                                    (16384.0 + ((flags & GEMM_HINT_CL2) ? 64.0 : 128.0) * bn);
            const double t_l2 = l2_bytes / 7000.0;
            if (t_l2 > cost) cost = t_l2;
            if (sp > 1) cost += 4900.0 + rows * N * 4.0 * (sp + 1) / 4200.0;
            if (best_cost < 0 || cost < best_cost * 0.97 || (cost <= best_cost && bn > best && sp <= best_sp)) {
                best_cost = cost;
                best = bn;
                best_sp = sp;
            }
        }
    }
    *BN = best;
    *splits = best_sp;
}

// Arrival tickets of the fused split-K reduction: zeroed once, self-resetting afterwards (the last arriver of a tile clears
// its slot), shared by all launches of a stream. They live in the KernelCtx of the engine that builds the op (kctx.h), are
// captured into the op by the setup functions and therefore never allocated during graph capture.
static int* tile_counters() {
    static const int on = []() {
        const char* e = getenv("DTP_SPLITK_FUSED");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    if (!on) return nullptr;
    KernelCtx* c = kctx_current();
    return c ? c->tile_counters : nullptr;
}
void gemm_prepare_splitk() { (void)tile_counters(); }
static void bind_ctx(GemmOp* op) {
    KernelCtx* c = kctx_current();
    op->tile_counters = op->p.splits > 1 ? tile_counters() : nullptr;
    op->p.err_flag = c ? c->err_flag_dev : nullptr;
}
static int num_sms();
static int max_pair_clusters();
// fp16 partials for the in-kernel split-K reduction (each partial is rounded once; the sum stays fp32): on by default
static int g_splitk_half = -1;  // -1: from the environment (DTP_SPLITK_F16, default on)
void gemm_set_splitk_half(int on) { g_splitk_half = on ? 1 : 0; }
static bool splitk_half() {
    if (g_splitk_half < 0) {
        const char* e = getenv("DTP_SPLITK_F16");
        g_splitk_half = (e && e[0] == '0') ? 0 : 1;
    }
    return g_splitk_half != 0;
}
// The in-kernel split-K reduction waits for the other CTAs of a tile's split group, so it is only used when every CTA owns
// exactly one tile and the whole grid is co-resident (tiles <= SMs; CTA pairs: <= the co-resident cluster count).
static bool splitk_fused(const GemmOp* op) {
    const GemmParams& p = op->p;
    if (p.splits <= 1 || op->tile_counters == nullptr) return false;
    const long long gn = (p.N + op->BN - 1) / op->BN;
    const long long groups = static_cast<long long>(op->grid_m) * gn * p.nz1 * p.nz2;
    if (2 * groups > kTileCounterSlots) return false;
    // slice staging of the reduction must fit the idle operand stages: 128 * BN * (2 | 4) B of partial slices + the fp32 sums of
    // the CTA's own slice; 320-wide tiles with fp32 partials only fit from 4 splits up, with fp16 partials always
    if (op->BN > 256 && p.splits < 4 && !splitk_half()) return false;
    if (op->cluster == 2) {
        const long long pair_tiles = static_cast<long long>((op->grid_m + 1) / 2) * gn * p.nz1 * p.nz2 * p.splits;
        return pair_tiles <= max_pair_clusters();
    }
    return groups * p.splits <= num_sms();
}
int gemm_num_launches(const GemmOp* op) { return (op->p.splits > 1 && !splitk_fused(op)) ? 2 : 1; }

static int num_sms() {
    KernelCtx* c = kctx_current();
    return c ? c->sms : 148;
}

// cudaFuncSetAttribute applies to the current device's instance of the kernel: one flag per device ordinal
struct PerDeviceFlag {
    bool set[kMaxDevices] = {false};
    bool& here() { return set[kctx_device()]; }
};

template <int BN, int STAGES>
static int launch_light(const GemmOp* op, cudaStream_t stream) {
    constexpr int SMEM = STAGES * (128 * 128 + BN * 128) + (3 * STAGES + 4) * 8 + 16 + BN * 8 + 1024;
    static_assert(SMEM <= 113 * 1024 && 2 * BN <= 256, "two CTAs per SM");
    static PerDeviceFlag attr_flag;
    bool& attr_set = attr_flag.here();
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) {
            snprintf(g_gemm_err, sizeof(g_gemm_err), "cudaFuncSetAttribute (light): %s", cudaGetErrorString(cudaGetLastError()));
            return -20;
        }
        attr_set = true;
    }
    GemmParams p = op->p;
    p.dbg_mode = 0;
    p.grid_m = op->grid_m;
    p.grid_n = (p.N + BN - 1) / BN;
    const long long tiles = static_cast<long long>(p.grid_m) * p.grid_n * p.nz1 * p.nz2 * p.splits;
    if (tiles > 0x7fffffffLL) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "too many tiles");
        return -24;
    }
    p.total_tiles = static_cast<int>(tiles);
    const int cap = 2 * num_sms();
    p.tile_counters = (tiles <= cap && splitk_fused(op)) ? op->tile_counters : nullptr;
    p.ws_half = (p.tile_counters != nullptr && splitk_half()) ? 1 : 0;
    const int grid = static_cast<int>(tiles < cap ? tiles : cap);
    cudaError_t e = launch_k(gemm_tc_kernel<BN, STAGES, 1, 2>, dim3(grid), dim3(320), SMEM, stream, op->mapA0, op->mapA1, op->mapB, op->mapBL, op->mapA2, op->mapA3, p);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "gemm launch (light): %s", cudaGetErrorString(e));
        return -21;
    }
    return 0;
}

template <int BN, int STAGES, int STAGES2, int FEAT>
static int launch_cfg(const GemmOp* op, cudaStream_t stream) {
    constexpr int SMEM = STAGES * (128 * 128 + BN * 128) + (3 * STAGES + 4) * 8 + 16 + BN * 8 + 1024;
    constexpr int SMEM2 = STAGES2 * (128 * 128 + BN * 64) + (3 * STAGES2 + 4) * 8 + 16 + BN * 8 + 1024;
    static_assert(SMEM <= 227 * 1024 && SMEM2 <= 227 * 1024, "shared memory budget");
    static PerDeviceFlag attr_flag;
    bool& attr_set = attr_flag.here();
    if (!attr_set) {
        cudaError_t e =
            cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, 1, 1, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES2, 2, 1, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2);
        if (e != cudaSuccess) {
            snprintf(g_gemm_err, sizeof(g_gemm_err), "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return -20;
        }
        attr_set = true;
    }
    GemmParams p = op->p;
    static const int dbgmode = []() {
        const char* e = getenv("DTP_EPI_DEBUG");
        return e ? atoi(e) : 0;
    }();
    p.dbg_mode = dbgmode;
    p.grid_m = op->grid_m;
    p.grid_n = (p.N + BN - 1) / BN;
    const int cl = op->cluster == 2 ? 2 : 1;
    const long long gm = cl == 2 ? (p.grid_m + 1) / 2 : p.grid_m;
    const long long tiles = gm * p.grid_n * p.nz1 * p.nz2 * p.splits;  // cluster mode: pairs of m-tiles
    if (tiles > 0x7fffffffLL) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "too many tiles");
        return -24;
    }
    p.total_tiles = static_cast<int>(tiles);
    p.tile_counters = splitk_fused(op) ? op->tile_counters : nullptr;
    p.ws_half = (p.tile_counters != nullptr && splitk_half()) ? 1 : 0;
    cudaError_t e;
    if (cl == 2) {
        const int max_clusters = num_sms() / 2;  // (fused split-K additionally requires tiles <= max_pair_clusters())
        const int nclusters = static_cast<int>(tiles < max_clusters ? tiles : max_clusters);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * nclusters);
        cfg.blockDim = dim3(320);
        cfg.dynamicSmemBytes = SMEM2;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = 2;
        attr[1].val.clusterDim.y = 1;
        attr[1].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 2;
        e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES2, 2, 1, FEAT>, op->mapA0, op->mapA1, op->mapBh, op->mapBL, op->mapA2, op->mapA3, p);
    } else {
        const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
        e = launch_k(gemm_tc_kernel<BN, STAGES, 1, 1, FEAT>, dim3(grid), dim3(320), SMEM, stream, op->mapA0, op->mapA1, op->mapB, op->mapBL, op->mapA2, op->mapA3, p);
    }
    if (e != cudaSuccess) e = cudaGetLastError();
    else e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "gemm launch: %s", cudaGetErrorString(e));
        return -21;
    }
    return 0;
}

// co-resident 2-CTA clusters of the pair-mode kernel (one CTA per SM for every instantiation; queried on one of them)
static int max_pair_clusters() {
    KernelCtx* kc = kctx_current();
    int local = -1;
    int& n = kc ? kc->max_pair_clusters : local;
    if (n < 0) {
        constexpr int SMEM2 = 8 * (128 * 128 + 128 * 64) + (3 * 8 + 4) * 8 + 16 + 128 * 8 + 1024;
        cudaFuncSetAttribute(gemm_tc_kernel<128, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * (num_sms() / 2));
        cfg.blockDim = dim3(320);
        cfg.dynamicSmemBytes = SMEM2;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int c = 0;
        if (cudaOccupancyMaxActiveClusters(&c, gemm_tc_kernel<128, 8, 2>, &cfg) != cudaSuccess) {
            cudaGetLastError();
            c = 0;
        }
        n = c;
    }
    return n;
}

// BN = 320: CTA pairs only (a single-CTA 128 x 320 tile would need 56 KB per stage)
template <int BN, int STAGES2, int FEAT>
static int launch_pair_only(const GemmOp* op, cudaStream_t stream) {
    constexpr int SMEM2 = STAGES2 * (128 * 128 + BN * 64) + (3 * STAGES2 + 4) * 8 + 16 + BN * 8 + 1024;
    static_assert(SMEM2 <= 227 * 1024, "shared memory budget");
    static PerDeviceFlag attr_flag;
    bool& attr_set = attr_flag.here();
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES2, 2, 1, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2) != cudaSuccess) {
            snprintf(g_gemm_err, sizeof(g_gemm_err), "cudaFuncSetAttribute (pair): %s", cudaGetErrorString(cudaGetLastError()));
            return -20;
        }
        attr_set = true;
    }
    if (op->cluster != 2) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "BN=%d needs pair mode", BN);
        return -25;
    }
    GemmParams p = op->p;
    p.dbg_mode = 0;
    p.grid_m = op->grid_m;
    p.grid_n = (p.N + BN - 1) / BN;
    const long long tiles = static_cast<long long>((p.grid_m + 1) / 2) * p.grid_n * p.nz1 * p.nz2 * p.splits;
    if (tiles > 0x7fffffffLL) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "too many tiles");
        return -24;
    }
    p.total_tiles = static_cast<int>(tiles);
    p.tile_counters = splitk_fused(op) ? op->tile_counters : nullptr;
    p.ws_half = (p.tile_counters != nullptr && splitk_half()) ? 1 : 0;
    const int max_clusters = num_sms() / 2;
    const int nclusters = static_cast<int>(tiles < max_clusters ? tiles : max_clusters);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * nclusters);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = SMEM2;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES2, 2, 1, FEAT>, op->mapA0, op->mapA1, op->mapBh, op->mapBL, op->mapA2, op->mapA3, p);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_gemm_err, sizeof(g_gemm_err), "gemm launch (pair): %s", cudaGetErrorString(e));
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
    // light configuration: DTP_GEMM_LIGHT = max k-blocks per tile for which it is used. Off by default: measured neutral
    // on isolated short-K launches and 1.5 % slower over a whole stamp (profiles/README.md).
    static const int light_max_kb = []() {
        const char* e = getenv("DTP_GEMM_LIGHT");
        return e ? atoi(e) : 0;
    }();
    const int kb_per_tile = (p.num_kb + p.splits - 1) / (p.splits > 0 ? p.splits : 1);
    const bool light = light_max_kb > 0 && op->cluster != 2 && op->BN <= 128 && kb_per_tile <= light_max_kb &&
                       !(p.flags & GEMM_W_BLOCKED);
    const bool feat = p.ln_stats != nullptr || p.stats_out != nullptr;  // folded LayerNorm / row statistics instantiation
    int r;
    if (light && !feat) {
        switch (op->BN) {
            case 32: r = launch_light<32, 4>(op, stream); break;
            case 64: r = launch_light<64, 4>(op, stream); break;
            default: r = launch_light<128, 3>(op, stream); break;
        }
    } else if (feat) {
        switch (op->BN) {
            case 32: r = launch_cfg<32, 8, 8, 1>(op, stream); break;
            case 64: r = launch_cfg<64, 8, 8, 1>(op, stream); break;
            case 128: r = launch_cfg<128, 6, 8, 1>(op, stream); break;
            case 160: r = launch_cfg<160, 5, 7, 1>(op, stream); break;
            case 192: r = launch_cfg<192, 5, 7, 1>(op, stream); break;
            case 320: r = launch_pair_only<320, 6, 1>(op, stream); break;
            default: r = launch_cfg<256, 4, 6, 1>(op, stream); break;
        }
    } else {
        switch (op->BN) {
            case 32: r = launch_cfg<32, 8, 8, 0>(op, stream); break;
            case 64: r = launch_cfg<64, 8, 8, 0>(op, stream); break;
            case 128: r = launch_cfg<128, 6, 8, 0>(op, stream); break;
            case 160: r = launch_cfg<160, 5, 7, 0>(op, stream); break;
            case 192: r = launch_cfg<192, 5, 7, 0>(op, stream); break;
            case 320: r = launch_pair_only<320, 6, 0>(op, stream); break;
            default: r = launch_cfg<256, 4, 6, 0>(op, stream); break;
        }
    }
    if (r) return r;
    if (p.splits > 1 && !splitk_fused(op)) {
        const int chunks = (p.N + 31) / 32;
        const long long total = static_cast<long long>(p.nz1) * p.nz2 * p.M * chunks;
        const int blocks = static_cast<int>((total + 255) / 256);
        launch_k(gemm_splitk_finalize_kernel, dim3(blocks), dim3(256), 0, stream, p);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            snprintf(g_gemm_err, sizeof(g_gemm_err), "finalize launch: %s", cudaGetErrorString(e));
            return -23;
        }
    }
    return 0;
}

}  // namespace dtp
