// HBM-bound kernels of the stamp path. See kernels.h.
#include "kernels.h"

#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "kctx.h"

namespace dtp {

static char g_kerr[256] = "";
const char* kernels_last_error() { return g_kerr; }

static long long* g_kdbg = nullptr;  // per-CTA globaltimer checkpoints of the single-launch GroupNorm (debug aid)
void kernels_set_debug(long long* dbg) { g_kdbg = dbg; }
__device__ __forceinline__ long long kgtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define KDBG(slot)                                                                                                     \
    do {                                                                                                               \
        if (dbg != nullptr && threadIdx.x == 0)                                                                        \
            dbg[(static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (slot)] = kgtime();               \
    } while (0)

static int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_kerr, sizeof(g_kerr), "%s: %s", what, cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

struct alignas(16) Half8 {
    __half2 h[4];
};
__device__ __forceinline__ Half8 ld8(const __half* p) {
    Half8 r;
    *reinterpret_cast<uint4*>(&r) = __ldg(reinterpret_cast<const uint4*>(p));
    return r;
}
__device__ __forceinline__ void st8(__half* p, const Half8& v) {
    *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&v);
}
__device__ __forceinline__ void unpack8(const Half8& v, float (&f)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __half22float2(v.h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ Half8 pack8(const float (&f)[8]) {
    Half8 v;
#pragma unroll
    for (int i = 0; i < 4; ++i) v.h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    return v;
}

// ------------------------------------------------------------------------------------------------------------
// GroupNorm (+SiLU) over one or two NHWC sources.
//   large tensors: (1) per-chunk partial sums; the last CTA of a sample (ticket counter) reduces them in a fixed order in
//                  double and emits per-channel (scale, shift); (2) apply.          -> 2 launches, deterministic
//   small tensors: one CTA per (sample, group) keeps its elements in registers: stats + apply in a single launch.
// ------------------------------------------------------------------------------------------------------------
int gn_num_chunks(int HW, int C) {
    long long elems = static_cast<long long>(HW) * C;
    long long chunks = elems / 16384;
    if (chunks < 1) chunks = 1;
    if (chunks > 512) chunks = 512;
    if (chunks > HW) chunks = HW;
    return static_cast<int>(chunks);
}
size_t gn_ws_floats(int Nimg, int HW, int C, int groups) {
    int chunks = gn_num_chunks(HW, C);
    if (chunks < 160) chunks = 160;  // the single-launch path publishes one partial per CTA (<= #SMs per sample)
    return static_cast<size_t>(Nimg) * chunks * groups * 2 + static_cast<size_t>(Nimg) * C * 2 + 64;
}

// self-resetting per-sample tickets (stream-ordered reuse) live in the current KernelCtx (kctx.h)
static int gn_counters(int Nimg, int** out) {
    KernelCtx* c = kctx_current();
    if (!c || Nimg > kGnSamples) return -1;
    *out = c->gn_counters;
    return 0;
}

// grid (chunks, Nimg); block = CV * rows_per_iter threads, CV = C / 8
__global__ void gn_stats_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW,
                                int groups, int chunks, float* __restrict__ partial, int* __restrict__ counters,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                float* __restrict__ coef) {
    pdl_enter();
    extern __shared__ float2 sm_acc[];  // [rows_per_iter][C]
    __shared__ int s_ticket;
    const int C = C0 + C1;
    const int CV = C >> 3;
    const int cv = threadIdx.x % CV;
    const int prow = threadIdx.x / CV;
    const int rows_per_iter = blockDim.x / CV;
    const int n = blockIdx.y;
    const int ppc = (HW + chunks - 1) / chunks;
    const int p0 = blockIdx.x * ppc;
    const int p1 = min(HW, p0 + ppc);
    const int c = cv * 8;
    const __half* src;
    int ldc;
    if (c < C0) {
        src = x0 + static_cast<long long>(n) * HW * C0 + c;
        ldc = C0;
    } else {
        src = x1 + static_cast<long long>(n) * HW * C1 + (c - C0);
        ldc = C1;
    }
    float s[8], ss[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = ss[e] = 0.0f;
    int p = p0 + prow;
    // four independent 16-byte loads in flight per thread
    for (; p + 3 * rows_per_iter < p1; p += 4 * rows_per_iter) {
        Half8 h[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) h[u] = ld8(src + static_cast<long long>(p + u * rows_per_iter) * ldc);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float f[8];
            unpack8(h[u], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                s[e] += f[e];
                ss[e] = fmaf(f[e], f[e], ss[e]);
            }
        }
    }
    for (; p < p1; p += rows_per_iter) {
        float f[8];
        unpack8(ld8(src + static_cast<long long>(p) * ldc), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            s[e] += f[e];
            ss[e] = fmaf(f[e], f[e], ss[e]);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) sm_acc[prow * C + c + e] = make_float2(s[e], ss[e]);
    __syncthreads();
    const int cpg = C / groups;
    if (threadIdx.x < groups) {
        float a = 0.0f, b = 0.0f;
        for (int r = 0; r < rows_per_iter; ++r)
            for (int cc = 0; cc < cpg; ++cc) {
                const float2 t = sm_acc[r * C + threadIdx.x * cpg + cc];
                a += t.x;
                b += t.y;
            }
        float* dst = partial + ((static_cast<long long>(n) * chunks + blockIdx.x) * groups + threadIdx.x) * 2;
        dst[0] = a;
        dst[1] = b;
    }
    // last CTA of this sample folds the partials (fixed order -> bitwise reproducible) into per-channel scale / shift
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(&counters[n], 1);
    __syncthreads();
    if (s_ticket != chunks - 1) return;
    __threadfence();
    // The whole CTA folds the partials (round 1 used one thread per group: 512 chunks at 8 loads in flight were 64 serial L2
    // round trips, ~60 us of the 95 us this kernel took on the 512 x 512 VAE tensors): nl threads per group take the
    // chunks l, l + nl, ... in order, then thread g adds the nl sub-sums in order - still a fixed summation order.
    double* dsum = reinterpret_cast<double*>(sm_acc);  // [groups][nl][2]; sm_acc is dead (read above, barrier passed)
    const int nl = max(1, min(static_cast<int>(blockDim.x) / groups, 32));
    {
        const int g = threadIdx.x / nl, l = threadIdx.x - g * nl;
        if (g < groups) {
            double a = 0.0, b = 0.0;
            const float2* pp = reinterpret_cast<const float2*>(partial) + static_cast<long long>(n) * chunks * groups + g;
            int ch = l;
            for (; ch + 3 * nl < chunks; ch += 4 * nl) {
                float2 t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) t[u] = __ldcg(pp + static_cast<long long>(ch + u * nl) * groups);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    a += static_cast<double>(t[u].x);
                    b += static_cast<double>(t[u].y);
                }
            }
            for (; ch < chunks; ch += nl) {
                const float2 t = __ldcg(pp + static_cast<long long>(ch) * groups);
                a += static_cast<double>(t.x);
                b += static_cast<double>(t.y);
            }
            dsum[(g * nl + l) * 2] = a;
            dsum[(g * nl + l) * 2 + 1] = b;
        }
    }
    __syncthreads();
    double mean_d = 0.0, rstd_d = 0.0;
    if (threadIdx.x < groups) {
        double a = 0.0, b = 0.0;
        for (int l = 0; l < nl; ++l) {
            a += dsum[(threadIdx.x * nl + l) * 2];
            b += dsum[(threadIdx.x * nl + l) * 2 + 1];
        }
        const double count = static_cast<double>(HW) * cpg;
        mean_d = a / count;
        double var = b / count - mean_d * mean_d;
        if (var < 0.0) var = 0.0;
        rstd_d = 1.0 / sqrt(var + static_cast<double>(eps));
    }
    __syncthreads();  // dsum fully consumed before stat (same storage) is written
    float* stat = reinterpret_cast<float*>(sm_acc);  // [groups][2] mean, rstd
    if (threadIdx.x < groups) {
        stat[2 * threadIdx.x] = static_cast<float>(mean_d);
        stat[2 * threadIdx.x + 1] = static_cast<float>(rstd_d);
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
        const int g = ch / cpg;
        const float sc = stat[2 * g + 1] * __ldg(gamma + ch);
        coef[(static_cast<long long>(n) * C + ch) * 2] = sc;
        coef[(static_cast<long long>(n) * C + ch) * 2 + 1] = __ldg(beta + ch) - stat[2 * g] * sc;
    }
    if (threadIdx.x == 0) counters[n] = 0;
}

__global__ void __launch_bounds__(256)
    gn_apply_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW,
                    long long total_vec, const float* __restrict__ coef, int silu, __half* __restrict__ out) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total_vec) return;
    const int C = C0 + C1;
    const int CV = C >> 3;
    const int cv = static_cast<int>(i % CV);
    const long long pix = i / CV;  // n*HW + p
    const int n = static_cast<int>(pix / HW);
    const int c = cv * 8;
    float f[8];
    if (c < C0)
        unpack8(ld8(x0 + pix * C0 + c), f);
    else
        unpack8(ld8(x1 + pix * C1 + (c - C0)), f);
    const float4* cf = reinterpret_cast<const float4*>(coef + (static_cast<long long>(n) * C + c) * 2);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float4 t = __ldg(cf + e);  // (scale, shift) of two channels
        float y0 = fmaf(f[2 * e], t.x, t.y), y1 = fmaf(f[2 * e + 1], t.z, t.w);
        if (silu) {
            y0 = silu_f(y0);
            y1 = silu_f(y1);
        }
        f[2 * e] = y0;
        f[2 * e + 1] = y1;
    }
    st8(out + pix * C + c, pack8(f));
}

// one CTA per (group, sample); every thread keeps up to EPT half2 pairs of the group's HW x cpg elements in registers
template <int EPT>
__global__ void __launch_bounds__(256)
    gn_group_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW, int groups,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
                    __half* __restrict__ out) {
    pdl_enter();
    __shared__ float red[2][8];
    __shared__ float s_mean, s_rstd;
    const int C = C0 + C1;
    const int cpg = C / groups;
    const int g = blockIdx.x, n = blockIdx.y;
    const int hp = cpg >> 1;           // half2 pairs per pixel
    const int total = HW * hp;         // pairs in this group
    const int c_base = g * cpg;
    float2 v[EPT];
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int i = threadIdx.x + k * 256;
        v[k] = make_float2(0.0f, 0.0f);
        if (i < total) {
            const int p = i / hp, c = c_base + 2 * (i - p * hp);
            const __half* src = (c < C0) ? x0 + (static_cast<long long>(n) * HW + p) * C0 + c
                                         : x1 + (static_cast<long long>(n) * HW + p) * C1 + (c - C0);
            v[k] = __half22float2(*reinterpret_cast<const __half2*>(src));
            s += v[k].x + v[k].y;
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    s = warp_sum(s);
    if (lane == 0) red[0][warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < 8; ++w) t += red[0][w];
        s_mean = t / static_cast<float>(2 * total);
    }
    __syncthreads();
    const float mean = s_mean;
    float ss = 0.0f;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int i = threadIdx.x + k * 256;
        if (i < total) {
            const float a = v[k].x - mean, b = v[k].y - mean;
            ss += a * a + b * b;
        }
    }
    ss = warp_sum(ss);
    if (lane == 0) red[1][warp] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < 8; ++w) t += red[1][w];
        s_rstd = rsqrtf(t / static_cast<float>(2 * total) + eps);
    }
    __syncthreads();
    const float rstd = s_rstd;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int i = threadIdx.x + k * 256;
        if (i < total) {
            const int p = i / hp, c = c_base + 2 * (i - p * hp);
            float y0 = (v[k].x - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
            float y1 = (v[k].y - mean) * rstd * __ldg(gamma + c + 1) + __ldg(beta + c + 1);
            if (silu) {
                y0 = silu_f(y0);
                y1 = silu_f(y1);
            }
            *reinterpret_cast<__half2*>(out + (static_cast<long long>(n) * HW + p) * C + c) = __floats2half2_rn(y0, y1);
        }
    }
}

// Register-resident GroupNorm, one CTA per (group, sample): groups are independent, so a CTA that owns a whole group needs
// no cross-CTA barrier at all - load once (every load of a thread in flight together), two in-CTA reductions (mean, then
// centred variance: the two-pass form), normalise out of registers. V halfs per vector (cpg % V == 0), at most EPT vectors per
// thread; the block size is a multiple of the vectors per pixel, so a thread's channel slot (and its gamma / beta, fetched
// before the dependency wait) is fixed.
template <int V>
struct HalfVec;
template <>
struct HalfVec<8> {
    using T = uint4;
};
template <>
struct HalfVec<4> {
    using T = uint2;
};
template <>
struct HalfVec<2> {
    using T = uint32_t;
};

__device__ __forceinline__ float gn_block_sum(float v, float* red, int warp, int lane) {
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (lane < static_cast<int>(blockDim.x >> 5)) ? red[lane] : 0.0f;
    t = warp_sum(t);  // every warp folds the same NW values in the same order
    __syncthreads();
    return t;
}

template <int V, int EPT>
__global__ void __launch_bounds__(960)
    gn_group2_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW, int groups,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
                     __half* __restrict__ out) {
    using VT = typename HalfVec<V>::T;
    __shared__ float red[32];
    const int C = C0 + C1;
    const int cpg = C / groups;
    const int vp = cpg / V;  // vectors per pixel of this group
    const int g = blockIdx.x, n = blockIdx.y;
    const int j = threadIdx.x % vp, prow = threadIdx.x / vp, rows_per_iter = blockDim.x / vp;
    const int c = g * cpg + j * V;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float ga[V], be[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        ga[e] = __ldg(gamma + c + e);
        be[e] = __ldg(beta + c + e);
    }
    const __half* src;
    int ldc;
    if (c < C0) {
        src = x0 + static_cast<long long>(n) * HW * C0 + c;
        ldc = C0;
    } else {
        src = x1 + static_cast<long long>(n) * HW * C1 + (c - C0);
        ldc = C1;
    }
    pdl_wait();
    VT v[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int p = prow + k * rows_per_iter;
        if (p < HW) v[k] = *reinterpret_cast<const VT*>(src + static_cast<long long>(p) * ldc);
    }
    pdl_launch_dependents();
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        if (prow + k * rows_per_iter < HW) {
            const __half2* h = reinterpret_cast<const __half2*>(&v[k]);
#pragma unroll
            for (int e = 0; e < V / 2; ++e) {
                const float2 f = __half22float2(h[e]);
                s += f.x + f.y;
            }
        }
    }
    const float inv_count = 1.0f / (static_cast<float>(HW) * static_cast<float>(cpg));
    const float mean = gn_block_sum(s, red, warp, lane) * inv_count;
    float ss = 0.0f;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        if (prow + k * rows_per_iter < HW) {
            const __half2* h = reinterpret_cast<const __half2*>(&v[k]);
#pragma unroll
            for (int e = 0; e < V / 2; ++e) {
                const float2 f = __half22float2(h[e]);
                const float a = f.x - mean, b = f.y - mean;
                ss = fmaf(a, a, ss);
                ss = fmaf(b, b, ss);
            }
        }
    }
    const float var = gn_block_sum(ss, red, warp, lane) * inv_count + eps;
    float rstd = rsqrtf(var);
    rstd = rstd * (1.5f - 0.5f * var * rstd * rstd);
    float sc[V], sh[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        sc[e] = rstd * ga[e];
        sh[e] = be[e] - mean * sc[e];
    }
    __half* dst = out + static_cast<long long>(n) * HW * C + c;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int p = prow + k * rows_per_iter;
        if (p < HW) {
            VT o;
            const __half2* h = reinterpret_cast<const __half2*>(&v[k]);
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int e = 0; e < V / 2; ++e) {
                const float2 f = __half22float2(h[e]);
                float y0 = fmaf(f.x, sc[2 * e], sh[2 * e]), y1 = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]);
                if (silu) {
                    y0 = silu_f(y0);
                    y1 = silu_f(y1);
                }
                oh[e] = __floats2half2_rn(y0, y1);
            }
            *reinterpret_cast<VT*>(dst + static_cast<long long>(p) * C) = o;
        }
    }
}

template <int V>
static int launch_gn_group2(int ept, dim3 grid, int threads, cudaStream_t st, const __half* x0, int C0, const __half* x1, int C1,
                            int HW, int groups, const float* gamma, const float* beta, float eps, int silu, __half* out) {
#define GN_G2(E) launch_k(gn_group2_kernel<V, E>, grid, dim3(threads), 0, st, x0, C0, x1, C1, HW, groups, gamma, beta, eps, silu, out)
    // instantiations hold at most 22 data registers per thread (ept * V / 2)
    if (ept <= 2)
        GN_G2(2);
    else if (ept <= 5)
        GN_G2(5);
    else if constexpr (V <= 4) {
        if (ept <= 11)
            GN_G2(11);
        else if constexpr (V == 2) {
            if (ept <= 16)
                GN_G2(16);
            else
                GN_G2(22);
        }
    }
#undef GN_G2
    return check_launch("gn_group2");
}

// 0: never, 1: by the measured rule, 2: whenever the shape allows (benchmarking)
static int gn_group2_mode() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DTP_GN_GROUP");
        v = e ? atoi(e) : 1;
    }
    return v;
}

// Single-launch GroupNorm for tensors that fit the SMs' shared memory (every UNet GroupNorm at B <= 4): the CTAs of a sample
// keep their row slab in smem, publish per-group partial sums, meet at a per-sample barrier (all CTAs are co-resident:
// grid <= #SMs, one CTA per SM), fold the partials in a fixed order and normalise straight out of smem. The tensor is
// read from L2/HBM once and there is one launch instead of two.
//   barrier state per sample: one monotonically growing arrival counter (see GN_BAR_QUANTUM); replaying the same launch
//   (CUDA graph) needs no host-side reset.
static int gn_barrier_state(int Nimg, unsigned** out) {
    KernelCtx* c = kctx_current();
    if (!c || Nimg > kGnSamples) return -1;
    *out = c->gn_barrier;
    return 0;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// grid (cps, Nimg); block = CV * rows_per_iter threads; dynamic smem = slab [rpc][C] fp16 | acc [rows_per_iter][C] float2
// Barrier: every launch adds exactly GN_BAR_QUANTUM to the sample's counter (CTA 0 adds the remainder), so the counter
// is a multiple of the quantum between launches and a CTA derives its target from the value its own arrival returned.
constexpr unsigned GN_BAR_QUANTUM = 256;
__global__ void __launch_bounds__(512)
    gn_fused_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW, int groups,
                    int rpc, float* __restrict__ partial, unsigned* __restrict__ barrier,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
                    __half* __restrict__ out, long long* __restrict__ dbg, int* __restrict__ err_flag) {
    extern __shared__ __align__(16) unsigned char gn_smem[];
    __shared__ float s_stat[256][2];
    const int C = C0 + C1;
    const int CV = C >> 3;
    const int cv = threadIdx.x % CV;
    const int prow = threadIdx.x / CV;
    const int rows_per_iter = blockDim.x / CV;
    const int n = blockIdx.y;
    const int cps = gridDim.x;
    const int p0 = blockIdx.x * rpc;
    const int p1 = min(HW, p0 + rpc);
    const int c = cv * 8;
    __half* slab = reinterpret_cast<__half*>(gn_smem);
    float2* acc = reinterpret_cast<float2*>(gn_smem + static_cast<size_t>(rpc) * C * sizeof(__half));
    const __half* src;
    int ldc;
    if (c < C0) {
        src = x0 + static_cast<long long>(n) * HW * C0 + c;
        ldc = C0;
    } else {
        src = x1 + static_cast<long long>(n) * HW * C1 + (c - C0);
        ldc = C1;
    }
    KDBG(0);
    // affine parameters do not depend on the producer: fetch them before the dependency wait
    float ga[8], be[8];
    {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
        ga[0] = g0.x, ga[1] = g0.y, ga[2] = g0.z, ga[3] = g0.w, ga[4] = g1.x, ga[5] = g1.y, ga[6] = g1.z, ga[7] = g1.w;
        be[0] = b0.x, be[1] = b0.y, be[2] = b0.z, be[3] = b0.w, be[4] = b1.x, be[5] = b1.y, be[6] = b1.z, be[7] = b1.w;
    }
    pdl_wait();
    KDBG(1);
    float s[8], ss[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = ss[e] = 0.0f;
    int p = p0 + prow;
    for (; p + 7 * rows_per_iter < p1; p += 8 * rows_per_iter) {  // eight 16-byte loads in flight per thread
        Half8 h[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) h[u] = ld8(src + static_cast<long long>(p + u * rows_per_iter) * ldc);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            st8(slab + static_cast<size_t>(p + u * rows_per_iter - p0) * C + c, h[u]);
            float f[8];
            unpack8(h[u], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                s[e] += f[e];
                ss[e] = fmaf(f[e], f[e], ss[e]);
            }
        }
    }
    for (; p < p1; p += rows_per_iter) {
        const Half8 h = ld8(src + static_cast<long long>(p) * ldc);
        st8(slab + static_cast<size_t>(p - p0) * C + c, h);
        float f[8];
        unpack8(h, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            s[e] += f[e];
            ss[e] = fmaf(f[e], f[e], ss[e]);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; e += 2)  // 16-byte stores: consecutive threads are 64 B apart, no bank conflicts beyond the minimum
        *reinterpret_cast<float4*>(&acc[prow * C + c + e]) = make_float4(s[e], ss[e], s[e + 1], ss[e + 1]);
    __syncthreads();
    KDBG(2);
    // per-group sums of this CTA: eight lanes per group walk the group's [rows_per_iter][cpg] accumulators, then a fixed
    // butterfly (bitwise reproducible); the block size is a multiple of 32, so every warp is complete
    const int cpg = C / groups;
    const int lane8 = threadIdx.x & 7, slot = threadIdx.x >> 3, nslots = blockDim.x >> 3;
    for (int gb = 0; gb < groups; gb += nslots) {
        const int g = gb + slot;
        float a = 0.0f, b = 0.0f;
        if (g < groups) {
            const int cnt = rows_per_iter * cpg;
            for (int e = lane8; e < cnt; e += 8) {
                const int r = e / cpg, cc = e - r * cpg;
                const float2 t = acc[r * C + g * cpg + cc];
                a += t.x;
                b += t.y;
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (g < groups && lane8 == 0)
            reinterpret_cast<float2*>(partial)[(static_cast<long long>(n) * cps + blockIdx.x) * groups + g] = make_float2(a, b);
    }
    // per-sample barrier (thread 0 fences on behalf of the CTA: the bar.sync orders the partial stores before it)
    __syncthreads();
    KDBG(3);
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned w = blockIdx.x == 0 ? GN_BAR_QUANTUM - static_cast<unsigned>(cps - 1) : 1u;
        const unsigned prev = atomicAdd(barrier + n, w);
        const unsigned target = (prev / GN_BAR_QUANTUM + 1u) * GN_BAR_QUANTUM;
        if (prev + w != target) {
            const long long t0 = clock64();
            while (static_cast<int>(ld_acquire_u32(barrier + n) - target) < 0) {
                if (clock64() - t0 > 4000000000LL) {  // ~2 s: a CTA of this grid never became resident (GPU shared?)
                    if (err_flag != nullptr) *err_flag = 1;  // reported by the engine; the context stays alive
                    break;
                }
            }
        }
        __threadfence();
    }
    __syncthreads();
    KDBG(4);
    pdl_launch_dependents();
    // fold the partials of all CTAs of this sample: eight lanes per group, lane l sums partials l, l+8, ... in order (all
    // loads of a lane in flight together), then the same fixed butterfly
    {
        const float inv_count = 1.0f / (static_cast<float>(HW) * static_cast<float>(cpg));
        for (int gb = 0; gb < groups; gb += nslots) {
            const int g = gb + slot;
            double a = 0.0, b = 0.0;
            if (g < groups) {
                const float2* pp = reinterpret_cast<const float2*>(partial) + static_cast<long long>(n) * cps * groups + g;
                for (int j = lane8; j < cps; j += 64) {
                    float2 t[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        t[u] = (j + 8 * u < cps) ? __ldcg(pp + static_cast<long long>(j + 8 * u) * groups) : make_float2(0.0f, 0.0f);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        a += static_cast<double>(t[u].x);
                        b += static_cast<double>(t[u].y);
                    }
                }
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            if (g < groups && lane8 == 0) {
                const double mean = a * static_cast<double>(inv_count);
                double var = b * static_cast<double>(inv_count) - mean * mean;
                const float v = fmaxf(static_cast<float>(var), 0.0f) + eps;
                float r = rsqrtf(v);
                r = r * (1.5f - 0.5f * v * r * r);  // one Newton step: full fp32 accuracy
                s_stat[g][0] = static_cast<float>(mean);
                s_stat[g][1] = r;
            }
        }
    }
    __syncthreads();
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int g = (c + e) / cpg;
        sc[e] = s_stat[g][1] * ga[e];
        sh[e] = be[e] - s_stat[g][0] * sc[e];
    }
    KDBG(5);
    __half* dst = out + static_cast<long long>(n) * HW * C + c;
    for (p = p0 + prow; p < p1; p += rows_per_iter) {
        float f[8];
        unpack8(*reinterpret_cast<const Half8*>(slab + static_cast<size_t>(p - p0) * C + c), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float y = fmaf(f[e], sc[e], sh[e]);
            if (silu) y = silu_f(y);
            f[e] = y;
        }
        st8(dst + static_cast<long long>(p) * C, pack8(f));
    }
    KDBG(6);
}

static int gn_fused_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DTP_GN_FUSED");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v;
}
static int gn_sm_count() {
    KernelCtx* c = kctx_current();
    return c ? c->sms : 148;
}

int launch_groupnorm(const __half* x0, int C0, const __half* x1, int C1, int Nimg, int HW, int groups,
                     const float* gamma, const float* beta, float eps, int silu, __half* out, float* stats_ws,
                     cudaStream_t st, int* launches) {
    if (launches) *launches = 1;
    if (x1 == nullptr) C1 = 0;
    const int C = C0 + C1;
    if ((C0 % 8) || (C1 % 8) || (C % groups) || groups > 256) {
        snprintf(g_kerr, sizeof(g_kerr), "groupnorm: unsupported channels C0=%d C1=%d groups=%d", C0, C1, groups);
        return -1;
    }
    const int CV = C / 8;
    if (CV > 1024) {
        snprintf(g_kerr, sizeof(g_kerr), "groupnorm: C=%d too large", C);
        return -1;
    }
    const int cpg = C / groups;
    if (gn_group2_mode() != 0) {
        // vector width: the widest that divides the group and keeps every vector inside one source
        const int V = (cpg % 8 == 0) ? 8 : (cpg % 4 == 0) ? 4 : (cpg % 2 == 0) ? 2 : 0;
        const int vp = V ? cpg / V : 0;
        if (V && vp <= 30 && (C0 % V) == 0) {
            int lcm = vp;  // block size: a multiple of the vectors per pixel and of the warp size
            while (lcm % 32) lcm += vp;
            const long long total = static_cast<long long>(HW) * vp;
            int threads = static_cast<int>(((total + lcm - 1) / lcm) * lcm);
            if (threads > 960) threads = (960 / lcm) * lcm;
            const int ept = threads > 0 ? static_cast<int>((HW + threads / vp - 1) / (threads / vp)) : 1 << 30;
            const bool fits = threads > 0 && ept * V <= 44;  // data registers per thread (64-register cap at 960 threads)
            // measured (profiles/gn_bench.py, B200): the barrier-free kernel wins while a group is small enough that one
            // CTA streams it faster than the multi-CTA kernel's publish / barrier / fold chain (~3.5 us) costs
            const int sms = gn_sm_count();
            const long long waves = (static_cast<long long>(Nimg) * groups + sms - 1) / sms;
            const bool rule = static_cast<long long>(HW) * cpg * waves <= 24576;
            if (fits && (gn_group2_mode() == 2 || rule)) {
                const dim3 grid(groups, Nimg);
                if (V == 8) return launch_gn_group2<8>(ept, grid, threads, st, x0, C0, x1, C1, HW, groups, gamma, beta, eps, silu, out);
                if (V == 4) return launch_gn_group2<4>(ept, grid, threads, st, x0, C0, x1, C1, HW, groups, gamma, beta, eps, silu, out);
                return launch_gn_group2<2>(ept, grid, threads, st, x0, C0, x1, C1, HW, groups, gamma, beta, eps, silu, out);
            }
        }
    }
    if (gn_fused_enabled() && CV <= 512 && Nimg <= gn_sm_count() && Nimg <= kGnSamples) {
        // block = CV * rows_per_iter threads, a multiple of 32 (the group reductions use full-warp shuffles)
        int gcd = CV, t32 = 32;
        while (t32) {
            const int r = gcd % t32;
            gcd = t32;
            t32 = r;
        }
        const int mult = 32 / gcd;
        int rows_per_iter = (512 / CV) / mult * mult;
        const int threads = CV * rows_per_iter;
        int cps = gn_sm_count() / Nimg;
        // keep at least ~8 KB of rows per CTA and no more CTAs than row groups
        int min_rows = (8192 + C * 2 - 1) / (C * 2);
        if (min_rows < rows_per_iter) min_rows = rows_per_iter;
        if (cps > (HW + min_rows - 1) / min_rows) cps = (HW + min_rows - 1) / min_rows;
        if (cps < 1) cps = 1;
        const int rpc = (HW + cps - 1) / cps;
        cps = (HW + rpc - 1) / rpc;
        const size_t smem = static_cast<size_t>(rpc) * C * sizeof(__half) + static_cast<size_t>(rows_per_iter) * C * sizeof(float2);
        if (rows_per_iter >= 1 && smem <= 200 * 1024 && groups <= 256) {
            static bool attr_set_dev[kMaxDevices] = {false};
            bool& attr_set = attr_set_dev[kctx_device()];
            if (!attr_set) {
                if (cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
                    return check_launch("gn_fused attr");
                attr_set = true;
            }
            unsigned* barrier = nullptr;
            if (gn_barrier_state(Nimg, &barrier)) {
                snprintf(g_kerr, sizeof(g_kerr), "groupnorm: barrier allocation failed");
                return -1;
            }
            launch_k(gn_fused_kernel, dim3(cps, Nimg), dim3(threads), smem, st, x0, C0, x1, C1, HW, groups, rpc, stats_ws, barrier,
                     gamma, beta, eps, silu, out, g_kdbg, kctx_current()->err_flag_dev);
            return check_launch("gn_fused");
        }
    }
    const long long pairs = static_cast<long long>(HW) * (cpg / 2);
    if ((cpg % 2) == 0 && pairs <= 256 * 12) {
        dim3 grid(groups, Nimg);
        if (pairs <= 256 * 4)
            launch_k(gn_group_kernel<4>, dim3(grid), dim3(256), 0, st, x0, C0, x1, C1, HW, groups, gamma, beta, eps, silu, out);
        else
            launch_k(gn_group_kernel<12>, dim3(grid), dim3(256), 0, st, x0, C0, x1, C1, HW, groups, gamma, beta, eps, silu, out);
        return check_launch("gn_group");
    }
    const int chunks = gn_num_chunks(HW, C);
    int rows_per_iter = 256 / CV;
    if (rows_per_iter < 1) rows_per_iter = 1;
    int threads = CV * rows_per_iter;
    if (threads < groups) {
        rows_per_iter = (groups + CV - 1) / CV;
        threads = CV * rows_per_iter;
    }
    int* counters = nullptr;
    if (gn_counters(Nimg, &counters)) {
        snprintf(g_kerr, sizeof(g_kerr), "groupnorm: counter allocation failed");
        return -1;
    }
    float* partial = stats_ws;
    float* coef = stats_ws + ((static_cast<size_t>(Nimg) * chunks * groups * 2 + 3) & ~size_t(3));
    size_t smem = static_cast<size_t>(rows_per_iter) * C * sizeof(float2);
    if (smem < static_cast<size_t>(groups) * 8) smem = static_cast<size_t>(groups) * 8;
    launch_k(gn_stats_kernel, dim3(dim3(chunks, Nimg)), dim3(threads), smem, st, x0, C0, x1, C1, HW, groups, chunks, partial, counters,
                                                               gamma, beta, eps, coef);
    if (check_launch("gn_stats")) return -1;
    if (launches) *launches = 2;
    const long long total_vec = static_cast<long long>(Nimg) * HW * CV;
    launch_k(gn_apply_kernel, dim3(static_cast<unsigned>((total_vec + 255) / 256)), dim3(256), 0, st, x0, C0, x1, C1, HW, total_vec, coef,
                                                                                   silu, out);
    return check_launch("gn_apply");
}

// Weights of conv3x3(nearest_upsample_2x(x)) as four parity-class 2x2 convolutions over x (gemm_setup_upconv2x):
//   Wst[(py*2+px) * Cout + o][(sy*2+sx) * C + c] = sum over the 3x3 taps (ky, kx) that land on input offset
//   (sy - 1 + py, sx - 1 + px), i.e. ky in S(py, sy), kx in S(px, sx) with S(0,0) = {0}, S(0,1) = {1,2}, S(1,0) = {0,1},
//   S(1,1) = {2}; summed in fp32, rounded to fp16 once.
__global__ void __launch_bounds__(256) upconv_fold_weights_kernel(const __half* __restrict__ W, int Cout, int C,
                                                                  __half* __restrict__ Wst) {
    const long long total = 16LL * Cout * C;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = static_cast<int>(i % C);
    long long r = i / C;
    const int s = static_cast<int>(r & 3);
    r >>= 2;
    const int o = static_cast<int>(r % Cout);
    const int cls = static_cast<int>(r / Cout);
    const int py = cls >> 1, px = cls & 1, sy = s >> 1, sx = s & 1;
    const int ky0 = (py == 0) ? (sy == 0 ? 0 : 1) : (sy == 0 ? 0 : 2), ky1 = (py == 0) ? (sy == 0 ? 0 : 2) : (sy == 0 ? 1 : 2);
    const int kx0 = (px == 0) ? (sx == 0 ? 0 : 1) : (sx == 0 ? 0 : 2), kx1 = (px == 0) ? (sx == 0 ? 0 : 2) : (sx == 0 ? 1 : 2);
    const __half* w = W + static_cast<long long>(o) * 9 * C + c;
    float acc = 0.0f;
    for (int ky = ky0; ky <= ky1; ++ky)
        for (int kx = kx0; kx <= kx1; ++kx) acc += __half2float(w[(ky * 3 + kx) * C]);
    Wst[(static_cast<long long>(cls) * Cout + o) * 4 * C + s * C + c] = __float2half_rn(acc);
}
int launch_upconv_fold_weights(const __half* W, int Cout, int C, __half* Wst, cudaStream_t st) {
    const long long total = 16LL * Cout * C;
    upconv_fold_weights_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(W, Cout, C, Wst);
    return check_launch("upconv_fold_weights");
}

// ------------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (C <= 2048), two-pass variance
// ------------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, int rows, int C,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, __half* __restrict__ out) {
    pdl_enter();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int CV = C >> 3;
    const __half* src = x + static_cast<long long>(row) * C;
    float f[NV][8];
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int v = lane + j * 32;
        if (v < CV) {
            unpack8(ld8(src + v * 8), f[j]);
#pragma unroll
            for (int e = 0; e < 8; ++e) s += f[j][e];
        }
    }
    const float mean = warp_sum(s) / static_cast<float>(C);
    float ss = 0.0f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int v = lane + j * 32;
        if (v < CV) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float d = f[j][e] - mean;
                ss += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(C) + eps);
    __half* dst = out + static_cast<long long>(row) * C;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int v = lane + j * 32;
        if (v < CV) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8) + 1);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8) + 1);
            const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] = fmaf((f[j][e] - mean) * rstd, gg[e], bb[e]);
            st8(dst + v * 8, pack8(y));
        }
    }
}

int launch_layernorm(const __half* x, int rows, int C, const float* gamma, const float* beta, float eps, __half* out,
                     cudaStream_t st) {
    if ((C % 8) || C > 2048) {
        snprintf(g_kerr, sizeof(g_kerr), "layernorm: unsupported C=%d", C);
        return -1;
    }
    if (rows <= 0) return 0;
    const int nv = (C / 8 + 31) / 32;
    const dim3 grid((rows + 7) / 8), block(256);
    if (nv <= 1)
        launch_k(layernorm_kernel<1>, grid, block, 0, st, x, rows, C, gamma, beta, eps, out);
    else if (nv <= 2)
        launch_k(layernorm_kernel<2>, grid, block, 0, st, x, rows, C, gamma, beta, eps, out);
    else if (nv <= 3)
        launch_k(layernorm_kernel<3>, grid, block, 0, st, x, rows, C, gamma, beta, eps, out);
    else if (nv <= 5)
        launch_k(layernorm_kernel<5>, grid, block, 0, st, x, rows, C, gamma, beta, eps, out);
    else
        launch_k(layernorm_kernel<8>, grid, block, 0, st, x, rows, C, gamma, beta, eps, out);
    return check_launch("layernorm");
}

// ------------------------------------------------------------------------------------------------------------
// folded LayerNorm: weight preparation (load time / per brush)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ln_fold_weights_kernel(const __half* __restrict__ W, int N, int K, int ldw,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const float* __restrict__ bias_in, __half* __restrict__ Wout,
                                                             float* __restrict__ colsum, float* __restrict__ bias_out) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    float cs = 0.0f, bs = 0.0f;
    for (int k = lane; k < K; k += 32) {
        const float w = __half2float(W[static_cast<long long>(n) * ldw + k]);
        if (Wout != nullptr) {  // (gamma == nullptr with Wout == nullptr: plain bias_out = bias_in + W beta)
            const __half wp = __float2half_rn(w * __ldg(gamma + k));
            Wout[static_cast<long long>(n) * ldw + k] = wp;
            cs += __half2float(wp);
        }
        bs = fmaf(w, __ldg(beta + k), bs);
    }
    cs = warp_sum(cs);
    bs = warp_sum(bs);
    if (lane == 0) {
        if (colsum) colsum[n] = cs;
        if (bias_out) bias_out[n] = bs + (bias_in ? bias_in[n] : 0.0f);
    }
}

int launch_ln_fold_weights(const __half* W, int N, int K, int ldw, const float* gamma, const float* beta,
                           const float* bias_in, __half* Wout, float* colsum, float* bias_out, cudaStream_t st) {
    ln_fold_weights_kernel<<<(N + 7) / 8, 256, 0, st>>>(W, N, K, ldw, gamma, beta, bias_in, Wout, colsum, bias_out);
    return check_launch("ln_fold_weights");
}

// grid (heads * 16, 3): one warp per score row
__global__ void __launch_bounds__(32) cross_ln_finish_kernel(const __half* __restrict__ wscore, const __half* __restrict__ kv,
                                                            const float* __restrict__ qbeta, int C, int heads, int T,
                                                            float scale, float* __restrict__ colsum,
                                                            float* __restrict__ bias) {
    pdl_enter();
    const int n = blockIdx.x, slot = blockIdx.y, lane = threadIdx.x;
    const int HP = heads * 16, h = n >> 4, j = n & 15, d = C / heads;
    const __half* w = wscore + (static_cast<long long>(slot) * HP + n) * C;
    float cs = 0.0f;
    for (int c = lane; c < C; c += 32) cs += __half2float(w[c]);
    cs = warp_sum(cs);
    float bs = 0.0f;
    if (j < T) {
        const int ctx = slot == 0 ? 0 : 1;  // [uncond | cond | cond] (inpaint_pipeline.py:140)
        const __half* k = kv + (static_cast<long long>(ctx) * T + j) * 2 * C + h * d;
        for (int i = lane; i < d; i += 32) bs = fmaf(__half2float(k[i]), __ldg(qbeta + h * d + i), bs);
    }
    bs = warp_sum(bs);
    if (lane == 0) {
        colsum[slot * HP + n] = cs;
        bias[slot * HP + n] = scale * bs;
    }
}

int launch_cross_ln_finish(const __half* wscore, const __half* kv, const float* qbeta, int C, int heads, int T,
                           float scale, float* colsum, float* bias, cudaStream_t st) {
    launch_k(cross_ln_finish_kernel, dim3(heads * 16, 3), dim3(32), 0, st, wscore, kv, qbeta, C, heads, T, scale, colsum, bias);
    return check_launch("cross_ln_finish");
}

// ------------------------------------------------------------------------------------------------------------
// row softmax, in place
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(__half* __restrict__ x, long long rows, int cols, int ld) {
    pdl_enter();
    const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    __half* p = x + row * ld;
    const bool vec = ((cols & 7) == 0) && ((ld & 7) == 0);
    float m = -INFINITY;
    if (vec) {
        for (int v = lane; v < (cols >> 3); v += 32) {
            float f[8];
            unpack8(*reinterpret_cast<const Half8*>(p + v * 8), f);
#pragma unroll
            for (int e = 0; e < 8; ++e) m = fmaxf(m, f[e]);
        }
    } else {
        for (int c = lane; c < cols; c += 32) m = fmaxf(m, __half2float(p[c]));
    }
    m = warp_max(m);
    float s = 0.0f;
    if (vec) {
        for (int v = lane; v < (cols >> 3); v += 32) {
            float f[8];
            unpack8(*reinterpret_cast<const Half8*>(p + v * 8), f);
#pragma unroll
            for (int e = 0; e < 8; ++e) s += __expf(f[e] - m);
        }
    } else {
        for (int c = lane; c < cols; c += 32) s += __expf(__half2float(p[c]) - m);
    }
    const float inv = 1.0f / warp_sum(s);
    if (vec) {
        for (int v = lane; v < (cols >> 3); v += 32) {
            float f[8];
            unpack8(*reinterpret_cast<const Half8*>(p + v * 8), f);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __expf(f[e] - m) * inv;
            st8(p + v * 8, pack8(f));
        }
    } else {
        for (int c = lane; c < cols; c += 32) p[c] = __float2half_rn(__expf(__half2float(p[c]) - m) * inv);
    }
}

int launch_softmax_rows(__half* x, long long rows, int cols, int ld, cudaStream_t st) {
    if (rows <= 0) return 0;
    launch_k(softmax_rows_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, st, x, rows, cols, ld);
    return check_launch("softmax_rows");
}

// ------------------------------------------------------------------------------------------------------------
// attention with nkv <= 64 keys resident in shared memory
// ------------------------------------------------------------------------------------------------------------
constexpr int ATTN_ROWS_PER_BLOCK = 32;
// grid (ceil(nq / 32), heads, batch); block 128
__global__ void __launch_bounds__(128)
    attn_small_kernel(const __half* __restrict__ q, int ldq, const __half* __restrict__ k, int ldk,
                      const __half* __restrict__ v, int ldv, __half* __restrict__ out, int ldo, int nq, int nkv, int d,
                      long long q_bs, long long kv_bs, long long o_bs, const int* __restrict__ kv_index, float scale,
                      int kpitch) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char sm_raw[];
    __half* Ks = reinterpret_cast<__half*>(sm_raw);  // [nkv][kpitch]
    __half* Vs = Ks + nkv * kpitch;                   // [nkv][d]
    float* qs = reinterpret_cast<float*>(Vs + ((nkv * d + 7) & ~7));  // [4][d]
    const int h = blockIdx.y, b = blockIdx.z;
    const int kvb = kv_index ? kv_index[b] : b;
    const __half* kb = k + kvb * kv_bs + h * d;
    const __half* vb = v + kvb * kv_bs + h * d;
    for (int i = threadIdx.x; i < nkv * (d >> 1); i += blockDim.x) {
        const int j = i / (d >> 1), c2 = i - j * (d >> 1);
        reinterpret_cast<__half2*>(Ks + j * kpitch)[c2] =
            __ldg(reinterpret_cast<const __half2*>(kb + static_cast<long long>(j) * ldk) + c2);
        reinterpret_cast<__half2*>(Vs + j * d)[c2] =
            __ldg(reinterpret_cast<const __half2*>(vb + static_cast<long long>(j) * ldv) + c2);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* qw = qs + warp * d;
    const int row_end = min(nq, (blockIdx.x + 1) * ATTN_ROWS_PER_BLOCK);
    for (int row = blockIdx.x * ATTN_ROWS_PER_BLOCK + warp; row < row_end; row += 4) {
        const __half* qr = q + b * q_bs + static_cast<long long>(row) * ldq + h * d;
        __syncwarp();
        for (int c = lane; c < d; c += 32) qw[c] = __half2float(qr[c]) * scale;
        __syncwarp();
        float s0 = -INFINITY, s1 = -INFINITY;
        if (lane < nkv) {
            float a = 0.0f;
            const __half2* kr = reinterpret_cast<const __half2*>(Ks + lane * kpitch);
            for (int c2 = 0; c2 < (d >> 1); ++c2) {
                const float2 t = __half22float2(kr[c2]);
                a += qw[2 * c2] * t.x + qw[2 * c2 + 1] * t.y;
            }
            s0 = a;
        }
        if (lane + 32 < nkv) {
            float a = 0.0f;
            const __half2* kr = reinterpret_cast<const __half2*>(Ks + (lane + 32) * kpitch);
            for (int c2 = 0; c2 < (d >> 1); ++c2) {
                const float2 t = __half22float2(kr[c2]);
                a += qw[2 * c2] * t.x + qw[2 * c2 + 1] * t.y;
            }
            s1 = a;
        }
        const float m = warp_max(fmaxf(s0, s1));
        const float e0 = (lane < nkv) ? __expf(s0 - m) : 0.0f;
        const float e1 = (lane + 32 < nkv) ? __expf(s1 - m) : 0.0f;
        const float inv = 1.0f / warp_sum(e0 + e1);
        const float p0 = e0 * inv, p1 = e1 * inv;
        __half* orow = out + b * o_bs + static_cast<long long>(row) * ldo + h * d;
        for (int c0 = 0; c0 < d; c0 += 32) {
            const int c = c0 + lane;
            float acc = 0.0f;
            for (int j = 0; j < nkv; ++j) {
                const float pj = __shfl_sync(0xffffffffu, (j < 32) ? p0 : p1, j & 31);
                if (c < d) acc += pj * __half2float(Vs[j * d + c]);
            }
            if (c < d) orow[c] = __float2half_rn(acc);
        }
    }
}

// Short key sequences (nkv <= 16: the 14 image tokens of the cross-attention): one thread per (query row, head). K and V
// of the head sit in shared memory (every lane of a warp reads the same key -> broadcast), scores and probabilities stay
// in registers, q is streamed in 16-byte pieces.
template <int NKV_MAX>
__global__ void __launch_bounds__(128)
    attn_tiny_kv_kernel(const __half* __restrict__ q, int ldq, const __half* __restrict__ k, int ldk,
                        const __half* __restrict__ v, int ldv, __half* __restrict__ out, int ldo, int nq, int nkv, int d,
                        long long q_bs, long long kv_bs, long long o_bs, const int* __restrict__ kv_index, float scale) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char sm_raw[];
    __half* Ks = reinterpret_cast<__half*>(sm_raw);  // [nkv][d]
    __half* Vs = Ks + nkv * d;                        // [nkv][d]
    const int h = blockIdx.y, b = blockIdx.z;
    const int kvb = kv_index ? kv_index[b] : b;
    const __half* kb = k + kvb * kv_bs + h * d;
    const __half* vb = v + kvb * kv_bs + h * d;
    const int dv = d >> 3;
    for (int i = threadIdx.x; i < nkv * dv; i += blockDim.x) {
        const int j = i / dv, c = i - j * dv;
        reinterpret_cast<uint4*>(Ks)[i] = __ldg(reinterpret_cast<const uint4*>(kb + static_cast<long long>(j) * ldk) + c);
        reinterpret_cast<uint4*>(Vs)[i] = __ldg(reinterpret_cast<const uint4*>(vb + static_cast<long long>(j) * ldv) + c);
    }
    __syncthreads();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nq) return;
    const __half* qr = q + b * q_bs + static_cast<long long>(row) * ldq + h * d;
    float s[NKV_MAX];
#pragma unroll
    for (int j = 0; j < NKV_MAX; ++j) s[j] = 0.0f;
    for (int c = 0; c < dv; ++c) {
        float qf[8];
        unpack8(ld8(qr + c * 8), qf);
#pragma unroll
        for (int j = 0; j < NKV_MAX; ++j) {
            if (j < nkv) {
                float kf[8];
                unpack8(*reinterpret_cast<const Half8*>(Ks + (j * dv + c) * 8), kf);
#pragma unroll
                for (int e = 0; e < 8; ++e) s[j] = fmaf(qf[e], kf[e], s[j]);
            }
        }
    }
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < NKV_MAX; ++j)
        if (j < nkv) m = fmaxf(m, s[j] * scale);
    float l = 0.0f;
#pragma unroll
    for (int j = 0; j < NKV_MAX; ++j) {
        s[j] = (j < nkv) ? __expf(s[j] * scale - m) : 0.0f;
        l += s[j];
    }
    const float inv = 1.0f / l;
    __half* orow = out + b * o_bs + static_cast<long long>(row) * ldo + h * d;
    for (int c = 0; c < dv; ++c) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = 0.0f;
#pragma unroll
        for (int j = 0; j < NKV_MAX; ++j) {
            if (j < nkv) {
                float vf[8];
                unpack8(*reinterpret_cast<const Half8*>(Vs + (j * dv + c) * 8), vf);
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = fmaf(s[j], vf[e], o[e]);
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] *= inv;
        st8(orow + c * 8, pack8(o));
    }
}

int launch_attn_small(const __half* q, int ldq, const __half* k, int ldk, const __half* v, int ldv, __half* out, int ldo,
                      int nq, int nkv, int heads, int d, int batch, long long q_bs, long long kv_bs, long long o_bs,
                      const int* kv_index, float scale, cudaStream_t st) {
    if (nkv > 64 || nkv < 1 || (d & 1) || (ldk & 1) || (ldv & 1)) {
        snprintf(g_kerr, sizeof(g_kerr), "attn_small: unsupported nkv=%d d=%d", nkv, d);
        return -1;
    }
    if (nkv <= 16 && (d % 8) == 0 && (ldq % 8) == 0 && (ldk % 8) == 0 && (ldv % 8) == 0 && (ldo % 8) == 0 && nq >= 64) {
        const size_t smem_t = static_cast<size_t>(nkv) * d * 2 * 2;
        if (smem_t <= 48 * 1024) {
            dim3 grid_t((nq + 127) / 128, heads, batch);
            launch_k(attn_tiny_kv_kernel<16>, dim3(grid_t), dim3(128), smem_t, st, q, ldq, k, ldk, v, ldv, out, ldo, nq, nkv, d, q_bs,
                                                                 kv_bs, o_bs, kv_index, scale);
            return check_launch("attn_tiny_kv");
        }
    }
    const int kpitch = d + ((((d >> 1) & 1) == 0) ? 2 : 0);
    const size_t smem = static_cast<size_t>(nkv) * kpitch * 2 + ((static_cast<size_t>(nkv) * d + 7) & ~size_t(7)) * 2 +
                        4 * static_cast<size_t>(d) * 4;
    static size_t smem_set = 48 * 1024;
    if (smem > smem_set) {
        if (cudaFuncSetAttribute(attn_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess) {
            snprintf(g_kerr, sizeof(g_kerr), "attn_small: smem %zu too large", smem);
            return -1;
        }
        smem_set = smem;
    }
    dim3 grid((nq + ATTN_ROWS_PER_BLOCK - 1) / ATTN_ROWS_PER_BLOCK, heads, batch);
    launch_k(attn_small_kernel, dim3(grid), dim3(128), smem, st, q, ldq, k, ldk, v, ldv, out, ldo, nq, nkv, d, q_bs, kv_bs, o_bs, kv_index,
                                              scale, kpitch);
    return check_launch("attn_small");
}

// ------------------------------------------------------------------------------------------------------------
// resampling / gathers
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    upsample2x_kernel(const __half* __restrict__ x, int H, int W, int CV, long long total, __half* __restrict__ out) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int cv = static_cast<int>(i % CV);
    long long t = i / CV;
    const int ox = static_cast<int>(t % (2 * W));
    t /= (2 * W);
    const int oy = static_cast<int>(t % (2 * H));
    const long long n = t / (2 * H);
    const uint4 val = __ldg(reinterpret_cast<const uint4*>(x) + ((n * H + (oy >> 1)) * W + (ox >> 1)) * CV + cv);
    reinterpret_cast<uint4*>(out)[i] = val;
}
int launch_upsample2x(const __half* x, int Nimg, int H, int W, int C, __half* out, cudaStream_t st) {
    if (C % 8) {
        snprintf(g_kerr, sizeof(g_kerr), "upsample2x: C=%d", C);
        return -1;
    }
    const long long total = static_cast<long long>(Nimg) * 4 * H * W * (C / 8);
    launch_k(upsample2x_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, x, H, W, C / 8, total, out);
    return check_launch("upsample2x");
}

__global__ void __launch_bounds__(256) im2col_s2_kernel(const __half* __restrict__ x, int H, int W, int CV, int pad_lo,
                                                        int Ho, int Wo, long long total, __half* __restrict__ out) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int cv = static_cast<int>(i % CV);
    long long t = i / CV;
    const int tap = static_cast<int>(t % 9);
    t /= 9;
    const int ox = static_cast<int>(t % Wo);
    t /= Wo;
    const int oy = static_cast<int>(t % Ho);
    const long long n = t / Ho;
    const int iy = 2 * oy + tap / 3 - pad_lo, ix = 2 * ox + tap % 3 - pad_lo;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
        val = __ldg(reinterpret_cast<const uint4*>(x) + ((n * H + iy) * W + ix) * CV + cv);
    reinterpret_cast<uint4*>(out)[i] = val;
}
int launch_im2col_s2(const __half* x, int Nimg, int H, int W, int C, int pad_lo, int Ho, int Wo, __half* out,
                     cudaStream_t st) {
    if (C % 8) {
        snprintf(g_kerr, sizeof(g_kerr), "im2col_s2: C=%d", C);
        return -1;
    }
    const long long total = static_cast<long long>(Nimg) * Ho * Wo * 9 * (C / 8);
    launch_k(im2col_s2_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, x, H, W, C / 8, pad_lo, Ho, Wo, total,
                                                                                out);
    return check_launch("im2col_s2");
}

__global__ void __launch_bounds__(256) patchify32_kernel(const float* __restrict__ x, long long total,
                                                         __half* __restrict__ out) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int kk = static_cast<int>(i % 3072);
    const long long t = i / 3072;
    const int patch = static_cast<int>(t % 49);
    const long long n = t / 49;
    const int c = kk >> 10, dy = (kk >> 5) & 31, dx = kk & 31;
    const int py = patch / 7, px = patch - py * 7;
    out[i] = __float2half_rn(__ldg(x + ((n * 3 + c) * 224 + py * 32 + dy) * 224 + px * 32 + dx));
}
int launch_patchify32(const float* x, int Nimg, __half* out, cudaStream_t st) {
    const long long total = static_cast<long long>(Nimg) * 49 * 3072;
    launch_k(patchify32_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, x, total, out);
    return check_launch("patchify32");
}

// ------------------------------------------------------------------------------------------------------------
// guidance + DDIM step (fp32, IEEE operation order of the reference; no fused multiply-add)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    guidance_ddim_kernel(const float* __restrict__ eps3, const float* __restrict__ lat, float* __restrict__ out,
                         long long n, float cfg, float tg, float sqrt_beta_t, float sqrt_alpha_t, float sqrt_alpha_prev,
                         float sqrt_beta_prev) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float eu = eps3[i], ec = eps3[n + i], et = eps3[2 * n + i];
    // eps = eu + cfg*(ec-eu) + tg*(et-ec)        (stable_diffusion_pipeline.py:449-451)
    float e = __fadd_rn(eu, __fmul_rn(cfg, __fsub_rn(ec, eu)));
    e = __fadd_rn(e, __fmul_rn(tg, __fsub_rn(et, ec)));
    // x0 = (x - sqrt(1-a_t) e) / sqrt(a_t);  x' = sqrt(a_prev) x0 + sqrt(1-a_prev) e     (utilities.py:477-505)
    const float x0 = __fdiv_rn(__fsub_rn(lat[i], __fmul_rn(sqrt_beta_t, e)), sqrt_alpha_t);
    out[i] = __fadd_rn(__fmul_rn(sqrt_alpha_prev, x0), __fmul_rn(sqrt_beta_prev, e));
}
int launch_guidance_ddim(const float* eps3, const float* latents_in, float* latents_out, int B, int chw, float cfg,
                         float tg, float alpha_t, float alpha_prev, cudaStream_t st) {
    const long long n = static_cast<long long>(B) * chw;
    const float beta_t = 1.0f - alpha_t;
    const float beta_prev = 1.0f - alpha_prev;  // (1 - a_prev - std_dev^2) with eta = 0
    launch_k(guidance_ddim_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, 
        eps3, latents_in, latents_out, n, cfg, tg, sqrtf(beta_t), sqrtf(alpha_t), sqrtf(alpha_prev), sqrtf(beta_prev));
    return check_launch("guidance_ddim");
}

__global__ void __launch_bounds__(256) expand_branches_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long grp_vec) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // vector of the 3-group output
    if (i >= 3 * grp_vec) return;
    out[i] = in[i < grp_vec ? i : i - grp_vec];
}
int launch_expand_branches(const __half* in, __half* out, int Bs, long long per_sample, cudaStream_t st) {
    const long long grp = static_cast<long long>(Bs) * per_sample;
    if ((grp & 7) || (reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
        snprintf(g_kerr, sizeof(g_kerr), "expand_branches: group size / pointers must be 16-byte multiples");
        return -1;
    }
    const long long grp_vec = grp / 8;
    launch_k(expand_branches_kernel, dim3(static_cast<unsigned>((3 * grp_vec + 255) / 256)), dim3(256), 0, st,
             reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), grp_vec);
    return check_launch("expand_branches");
}

// ------------------------------------------------------------------------------------------------------------
// layout packers
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    pack_unet_input_kernel(const float* __restrict__ lat, const float* __restrict__ mask3,
                           const float* __restrict__ masked3, int B, int hw, __half* __restrict__ out) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // (sample, pixel)
    if (i >= static_cast<long long>(3) * B * hw) return;
    const int s = static_cast<int>(i / hw), p = static_cast<int>(i % hw);
    const int b = s % B;
    __align__(16) __half h[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) h[c] = __float2half_rn(0.0f);
#pragma unroll
    for (int c = 0; c < 4; ++c) h[c] = __float2half_rn(lat[(static_cast<long long>(b) * 4 + c) * hw + p]);
    h[4] = __float2half_rn(mask3[static_cast<long long>(s) * hw + p]);
#pragma unroll
    for (int c = 0; c < 4; ++c) h[5 + c] = __float2half_rn(masked3[(static_cast<long long>(s) * 4 + c) * hw + p]);
    uint4* dst = reinterpret_cast<uint4*>(out + i * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = reinterpret_cast<const uint4*>(h)[j];
}
int launch_pack_unet_input(const float* latents, const float* mask3, const float* masked3, int B, int hw, __half* out,
                           cudaStream_t st) {
    const long long n = static_cast<long long>(3) * B * hw;
    launch_k(pack_unet_input_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, latents, mask3, masked3, B, hw, out);
    return check_launch("pack_unet_input");
}

__global__ void __launch_bounds__(256) nchw_to_nhwc_pad_kernel(const float* __restrict__ x, int C, int HW, int Cpad,
                                                               float divisor, long long total,
                                                               __half* __restrict__ out) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // (n, pixel, cpad/8)
    if (i >= total) return;
    const int CV = Cpad >> 3;
    const int cv = static_cast<int>(i % CV);
    const long long t = i / CV;
    const int p = static_cast<int>(t % HW);
    const long long n = t / HW;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = cv * 8 + e;
        f[e] = (c < C) ? __fdiv_rn(__ldg(x + (n * C + c) * HW + p), divisor) : 0.0f;
    }
    st8(out + i * 8, pack8(f));
}
int launch_nchw_to_nhwc_pad(const float* x, int Nimg, int C, int HW, int Cpad, float divisor, __half* out,
                            cudaStream_t st) {
    const long long total = static_cast<long long>(Nimg) * HW * (Cpad / 8);
    launch_k(nchw_to_nhwc_pad_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, x, C, HW, Cpad, divisor, total,
                                                                                       out);
    return check_launch("nchw_to_nhwc_pad");
}

__global__ void __launch_bounds__(256) vae_sample_kernel(const float* __restrict__ mom, const float* __restrict__ noise,
                                                         int hw, long long total, float scale, float* __restrict__ out) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // (b, c<4, p)
    if (i >= total) return;
    const int p = static_cast<int>(i % hw);
    const long long t = i / hw;
    const int c = static_cast<int>(t % 4);
    const long long b = t / 4;
    const float mean = mom[(b * 8 + c) * hw + p];
    float z = mean;
    if (noise != nullptr) {
        float logvar = mom[(b * 8 + 4 + c) * hw + p];
        logvar = fminf(fmaxf(logvar, -30.0f), 20.0f);
        z = __fadd_rn(mean, __fmul_rn(expf(0.5f * logvar), noise[i]));
    }
    out[i] = __fmul_rn(scale, z);
}
int launch_vae_sample(const float* moments, const float* noise, int B, int hw, float scale, float* out,
                      cudaStream_t st) {
    const long long total = static_cast<long long>(B) * 4 * hw;
    launch_k(vae_sample_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, moments, noise, hw, total, scale, out);
    return check_launch("vae_sample");
}

__global__ void __launch_bounds__(256) mask_nearest_kernel(const float* __restrict__ in, int R, int f, long long total,
                                                           float* __restrict__ out) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int r = R / f;
    const int x = static_cast<int>(i % r);
    const long long t = i / r;
    const int y = static_cast<int>(t % r);
    const long long b = t / r;
    out[i] = in[(b * R + static_cast<long long>(y) * f) * R + static_cast<long long>(x) * f];
}
int launch_mask_nearest(const float* in, int B, int R, int f, float* out, cudaStream_t st) {
    const long long total = static_cast<long long>(B) * (R / f) * (R / f);
    launch_k(mask_nearest_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, in, R, f, total, out);
    return check_launch("mask_nearest");
}

// ------------------------------------------------------------------------------------------------------------
// canvas pre-process (K19) and post-process (K18)
// ------------------------------------------------------------------------------------------------------------
// pass 1: row-wise running max of alpha over x in [x - pad/2, x + pad - pad/2 - 1]
__global__ void __launch_bounds__(256) dilate_rows_kernel(const float* __restrict__ canvas, int R, int pad,
                                                          long long total, float* __restrict__ scratch) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // (b, y, x)
    if (i >= total) return;
    const int x = static_cast<int>(i % R);
    const long long t = i / R;
    const int y = static_cast<int>(t % R);
    const long long b = t / R;
    const float* a = canvas + ((b * 4 + 3) * R + y) * R;
    const int lo = max(0, x - pad / 2), hi = min(R - 1, x + pad - pad / 2 - 1);
    float m = -1e4f;
    for (int xx = lo; xx <= hi; ++xx) m = fmaxf(m, __ldg(a + xx));
    scratch[i] = m;
}
__global__ void __launch_bounds__(256)
    canvas_finish_kernel(const float* __restrict__ canvas, const float* __restrict__ brush,
                         const float* __restrict__ scratch, int R, int pad, long long total,
                         float* __restrict__ masked_img, float* __restrict__ mask, float* __restrict__ ctx_img,
                         float* __restrict__ ctx_mask) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // (b, y, x)
    if (i >= total) return;
    const int x = static_cast<int>(i % R);
    const long long t = i / R;
    const int y = static_cast<int>(t % R);
    const long long b = t / R;
    const long long plane = static_cast<long long>(R) * R;
    const int lo = max(0, y - pad / 2), hi = min(R - 1, y + pad - pad / 2 - 1);
    float dil = -1e4f;
    for (int yy = lo; yy <= hi; ++yy) dil = fmaxf(dil, __ldg(scratch + (b * R + yy) * R + x));
    const float alpha = canvas[(b * 4 + 3) * plane + static_cast<long long>(y) * R + x];
    const float hint = __fsub_rn(1.0f, dil);
    const long long pix = static_cast<long long>(y) * R + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float img = __fsub_rn(__fmul_rn(canvas[(b * 4 + c) * plane + pix], 2.0f), 1.0f);
        const float mi = __fmul_rn(img, alpha);
        const float bs = __fsub_rn(__fmul_rn(__ldg(brush + c * plane + pix), 2.0f), 1.0f);
        masked_img[(b * 3 + c) * plane + pix] = mi;
        ctx_img[(b * 3 + c) * plane + pix] = __fadd_rn(mi, __fmul_rn(bs, hint));
    }
    mask[b * plane + pix] = __fsub_rn(1.0f, alpha);
    const float cm = fminf(fmaxf(__fadd_rn(alpha, hint), 0.0f), 1.0f);
    ctx_mask[b * plane + pix] = __fsub_rn(1.0f, cm);
}
int launch_canvas_preprocess(const float* canvas, const float* brush, int B, int R, int pad, float* masked_img,
                             float* mask, float* ctx_img, float* ctx_mask, float* scratch, cudaStream_t st) {
    if (pad < 1) pad = 1;
    const long long total = static_cast<long long>(B) * R * R;
    const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
    launch_k(dilate_rows_kernel, dim3(blocks), dim3(256), 0, st, canvas, R, pad, total, scratch);
    if (check_launch("dilate_rows")) return -1;
    launch_k(canvas_finish_kernel, dim3(blocks), dim3(256), 0, st, canvas, brush, scratch, R, pad, total, masked_img, mask, ctx_img,
                                                ctx_mask);
    return check_launch("canvas_finish");
}

__global__ void __launch_bounds__(256) composite_kernel(const float* __restrict__ canvas, const float* __restrict__ raw,
                                                        int R, long long total, float* __restrict__ out_f32,
                                                        unsigned char* __restrict__ out_u8) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // (b, pix)
    if (i >= total) return;
    const long long plane = static_cast<long long>(R) * R;
    const long long pix = i % plane, b = i / plane;
    const float alpha = canvas[(b * 4 + 3) * plane + pix];
    const float ia = __fsub_rn(1.0f, alpha);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = __fadd_rn(__fmul_rn(canvas[(b * 4 + c) * plane + pix], alpha),
                                  __fmul_rn(raw[(b * 3 + c) * plane + pix], ia));
        if (out_f32) out_f32[(b * 3 + c) * plane + pix] = v;
        if (out_u8) out_u8[(b * plane + pix) * 3 + c] = static_cast<unsigned char>(__fmul_rn(v, 255.0f));
    }
}
int launch_composite(const float* canvas, const float* raw, int B, int R, float* out_f32, unsigned char* out_u8hwc,
                     cudaStream_t st) {
    const long long total = static_cast<long long>(B) * R * R;
    launch_k(composite_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, canvas, raw, R, total, out_f32,
                                                                                out_u8hwc);
    return check_launch("composite");
}

// ------------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) add_rows_bcast_kernel(__half* __restrict__ x, const float* __restrict__ add,
                                                             long long total, int C, int period) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = static_cast<int>(i % C);
    const long long r = i / C;
    x[i] = __float2half_rn(__half2float(x[i]) + add[(r % period) * C + c]);
}
int launch_add_rows_bcast(__half* x, const float* add, long long rows, int C, int period, cudaStream_t st) {
    const long long total = rows * C;
    launch_k(add_rows_bcast_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, x, add, total, C, period);
    return check_launch("add_rows_bcast");
}
__global__ void __launch_bounds__(256) f32_to_f16_kernel(const float* __restrict__ x, __half* __restrict__ out,
                                                         long long n) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2half_rn(x[i]);
}
int launch_f32_to_f16(const float* x, __half* out, long long n, cudaStream_t st) {
    if (n <= 0) return 0;
    launch_k(f32_to_f16_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, x, out, n);
    return check_launch("f32_to_f16");
}
__global__ void __launch_bounds__(256) f16_to_f32_kernel(const __half* __restrict__ x, float* __restrict__ out,
                                                         long long n) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __half2float(x[i]);
}
int launch_f16_to_f32(const __half* x, float* out, long long n, cudaStream_t st) {
    if (n <= 0) return 0;
    launch_k(f16_to_f32_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, x, out, n);
    return check_launch("f16_to_f32");
}
__global__ void __launch_bounds__(256) copy_f32_kernel(const float* __restrict__ s, float* __restrict__ d, long long n) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) d[i] = s[i];
}
int launch_copy_f32(const float* src, float* dst, long long n, cudaStream_t st) {
    if (n <= 0) return 0;
    launch_k(copy_f32_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, src, dst, n);
    return check_launch("copy_f32");
}

// diffusers Timesteps(dim, flip_sin_to_cos=True, freq_shift=0): emb = [cos(t f_k) | sin(t f_k)], f_k = exp(-ln(1e4) k / (dim/2))
__global__ void timestep_embedding_kernel(const float* __restrict__ ts, int n, int dim, __half* __restrict__ out) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * dim) return;
    const int r = i / dim, c = i % dim;
    const int half_dim = dim / 2;
    const int k = c % half_dim;
    const float f = expf(-logf(10000.0f) * static_cast<float>(k) / static_cast<float>(half_dim));
    const float a = ts[r] * f;
    out[i] = __float2half_rn(c < half_dim ? cosf(a) : sinf(a));
}
int launch_timestep_embedding(const float* timesteps, int n, int dim, __half* out, cudaStream_t st) {
    launch_k(timestep_embedding_kernel, dim3((n * dim + 255) / 256), dim3(256), 0, st, timesteps, n, dim, out);
    return check_launch("timestep_embedding");
}

// CLIP token assembly: out[n, 0, :] = cls + pos[0]; out[n, 1+j, :] = patch_tok[n*49+j] + pos[1+j]
__global__ void __launch_bounds__(256) clip_embed_kernel(const __half* __restrict__ tok, const float* __restrict__ cls,
                                                         const float* __restrict__ pos, int C, long long total,
                                                         __half* __restrict__ out) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = static_cast<int>(i % C);
    const long long r = i / C;
    const int t = static_cast<int>(r % 50);
    const long long n = r / 50;
    const float base = (t == 0) ? cls[c] : __half2float(tok[(n * 49 + (t - 1)) * C + c]);
    out[i] = __float2half_rn(base + pos[t * C + c]);
}
int launch_clip_embed(const __half* patch_tok, const float* cls, const float* pos, int Nimg, int C, __half* out,
                      cudaStream_t st) {
    const long long total = static_cast<long long>(Nimg) * 50 * C;
    launch_k(clip_embed_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, patch_tok, cls, pos, C, total, out);
    return check_launch("clip_embed");
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const __half* __restrict__ x, int ld,
                                                          const int* __restrict__ idx, int C, long long total,
                                                          __half* __restrict__ out) {
    pdl_enter();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = static_cast<int>(i % C);
    const long long r = i / C;
    out[i] = x[static_cast<long long>(idx[r]) * ld + c];
}
int launch_gather_rows(const __half* x, int ld, const int* rows_idx, int nrows, int C, __half* out, cudaStream_t st) {
    const long long total = static_cast<long long>(nrows) * C;
    launch_k(gather_rows_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, x, ld, rows_idx, C, total, out);
    return check_launch("gather_rows");
}

}  // namespace dtp
