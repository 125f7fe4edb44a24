// Native runtime of the stamp path. See runtime.h.
#include "runtime.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <nvtx3/nvToolsExt.h>

#include "kernels.h"

namespace dtp {

// ------------------------------------------------------------------------------------------------------------
// Arena
// ------------------------------------------------------------------------------------------------------------
void* Arena::alloc(size_t bytes) {
    bytes = (bytes + 1023) & ~size_t(1023);
    if (bytes == 0) bytes = 1024;
    for (auto it = free_.begin(); it != free_.end(); ++it) {
        if (it->second >= bytes) {
            const size_t off = it->first, sz = it->second;
            free_.erase(it);
            if (sz > bytes) free_[off + bytes] = sz - bytes;
            live_[off] = bytes;
            used_ += bytes;
            peak_ = std::max(peak_, off + bytes);
            return base_ + off;
        }
    }
    return nullptr;
}

void Arena::release(void* p) {
    if (!p) return;
    const size_t off = static_cast<size_t>(static_cast<char*>(p) - base_);
    auto it = live_.find(off);
    if (it == live_.end()) return;
    size_t sz = it->second;
    live_.erase(it);
    used_ -= sz;
    size_t start = off;
    auto nx = free_.lower_bound(off);
    if (nx != free_.end() && nx->first == off + sz) {
        sz += nx->second;
        nx = free_.erase(nx);
    }
    if (nx != free_.begin()) {
        auto pv = std::prev(nx);
        if (pv->first + pv->second == off) {
            start = pv->first;
            sz += pv->second;
            free_.erase(pv);
        }
    }
    free_[start] = sz;
}

cudaEvent_t Profiler::get() {
    if (next == pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        pool.push_back(e);
    }
    return pool[next++];
}
void Profiler::collect() {
    if (recs.empty()) return;
    cudaEventSynchronize(recs.back().b);
    for (auto& r : recs) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        us[r.kind] += 1000.0 * ms;
        n[r.kind] += r.launches;
        if (r.label && !r.label->empty()) {
            auto& e = by_label[*r.label];
            e.first += 1000.0 * ms;
            e.second += 1;
        }
    }
    recs.clear();
    next = 0;
}
void Profiler::reset() {
    collect();
    for (int i = 0; i < K_NUM; ++i) {
        us[i] = 0;
        n[i] = 0;
    }
    by_label.clear();
}

// Measurement aid (dtp_set_option("debug_skip_kinds", mask)): ops whose kind bit is set are not launched, so the in-graph
// cost of a kernel family is the difference of two stamp times. Results are garbage while it is non-zero.
static int g_debug_skip_kinds = 0;
// same, per op label: dtp_set_option("debug_skip_label", h) leaves out the ops whose label hashes to h (31-bit FNV-1a of the
// label text as dtp_profile_dump prints it; profiles/ablate_labels.py)
static int g_debug_skip_label = 0;
static int label_hash(const std::string& s) {
    unsigned h = 2166136261u;
    for (unsigned char c : s) h = (h ^ c) * 16777619u;
    h &= 0x7fffffffu;
    return h ? static_cast<int>(h) : 1;
}

int Plan::run(cudaStream_t st, long long* launch_counter, Profiler* prof) const {
    const bool p = prof && prof->on;
    for (size_t i = 0; i < ops.size(); ++i) {
        if (g_debug_skip_kinds && i < kinds.size() && ((g_debug_skip_kinds >> kinds[i]) & 1)) continue;
        if (g_debug_skip_label && i < labels.size() && label_hash(labels[i]) == g_debug_skip_label) continue;
        cudaEvent_t a = nullptr, b = nullptr;
        if (p) {
            a = prof->get();
            b = prof->get();
            cudaEventRecord(a, st);
        }
        const int r = ops[i](st);
        if (r < 0) return r;
        if (p) {
            cudaEventRecord(b, st);
            prof->recs.push_back({i < kinds.size() ? kinds[i] : 0, r, a, b, i < labels.size() ? &labels[i] : nullptr});
        }
        if (launch_counter) *launch_counter += r;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// Builder: appends launch closures to a plan, carving activations from the arena
// ------------------------------------------------------------------------------------------------------------
struct Lin {
    const __half* A0 = nullptr;
    int lda0 = 0, K0 = 0;
    const __half* A1 = nullptr;
    int lda1 = 0, K1 = 0;
    long long M = 0;
    const __half* W = nullptr;
    int ldw = 0, N = 0;
    const float* bias = nullptr;
    const __half* res = nullptr;
    int ldr = 0;
    void* out = nullptr;
    int ldc = 0;
    int flags = 0;
    float alpha = 1.0f;
    int hw_out = 0;
    // folded LayerNorm on the A rows (gemm_tc.h) / row statistics of the output for the next folded LayerNorm
    const float2* ln_stats = nullptr;
    const float* ln_colsum = nullptr;
    int ln_C = 0;
    float2* stats_out = nullptr;
    int tune_kb = 0;  // > 0: look the tile / split configuration up for this many k-blocks (a measured neighbour shape)
};

struct Builder {
    Engine& e;
    Plan& plan;
    std::string prefix;
    bool ok = true;

    Builder(Engine& eng, Plan& p, const std::string& pre) : e(eng), plan(p), prefix(pre) {}

    void fail(const std::string& m) {
        if (ok) e.err_ = m;
        ok = false;
    }
    const WT* wt(const std::string& name) {
        const WT* t = e.find(prefix + name);
        if (!t) fail("missing weight tensor '" + prefix + name + "'");
        return t;
    }
    const __half* W16(const std::string& name, int* rows = nullptr, int* cols = nullptr) {
        const WT* t = wt(name);
        if (!t) return nullptr;
        if (t->dtype != 1 || t->shape.size() != 2) {
            fail("weight '" + prefix + name + "' must be a 2-D f16 matrix");
            return nullptr;
        }
        if (rows) *rows = static_cast<int>(t->shape[0]);
        if (cols) *cols = static_cast<int>(t->shape[1]);
        return static_cast<const __half*>(t->dev);
    }
    const float* F32(const std::string& name) {
        const WT* t = wt(name);
        if (!t) return nullptr;
        if (t->dtype != 0) {
            fail("weight '" + prefix + name + "' must be f32");
            return nullptr;
        }
        return static_cast<const float*>(t->dev);
    }
    Act alloc(int N, int H, int W, int C) {
        Act a;
        a.N = N;
        a.H = H;
        a.W = W;
        a.C = C;
        a.p = static_cast<__half*>(e.arena_.alloc(a.bytes()));
        if (!a.p) {
            char buf[160];
            snprintf(buf, sizeof(buf), "activation arena exhausted (need %zu more bytes, capacity %zu): raise arena_bytes",
                     a.bytes(), e.arena_.capacity());
            fail(buf);
        }
        return a;
    }
    Act like(const Act& a, int C) { return alloc(a.N, a.H, a.W, C); }
    void* raw(size_t bytes) {
        void* p = e.arena_.alloc(bytes);
        if (!p) fail("activation arena exhausted: raise arena_bytes");
        return p;
    }
    void release(Act& a) {
        e.arena_.release(a.p);
        a.p = nullptr;
    }
    void release_raw(void* p) { e.arena_.release(p); }

    void push_gemm(GemmOp op, const char* what = "linear") {
        e.ws_needed_ = std::max(e.ws_needed_, gemm_workspace_bytes(&op));
        Engine* eng = &e;
        char lab[192];
        snprintf(lab, sizeof(lab), "%s M=%d N=%d K=%d z=%d BN=%d splits=%d grid=%dx%d flags=0x%x", what, op.p.M, op.p.N,
                 op.p.num_kb * 64, op.p.nz1 * op.p.nz2, op.BN, op.p.splits, op.grid_m, (op.p.N + op.BN - 1) / op.BN,
                 op.p.flags);
        const std::string label = lab;
        plan.add(K_GEMM, [op, eng](cudaStream_t st) mutable -> int {
            op.p.workspace = eng->ws_;
            if (gemm_launch(&op, st)) {
                eng->err_ = std::string("contraction launch failed: ") + gemm_last_error();
                return -1;
            }
            return gemm_num_launches(&op);
        }, label);
    }

    void linear(const Lin& a) {
        if (!ok) return;
        GemmOp op;
        int BN, splits;
        const int kb = (a.K0 + 63) / 64 + (a.A1 ? (a.K1 + 63) / 64 : 0);
        const int mt = static_cast<int>((a.M + 127) / 128);
        const int hint = a.flags | ((mt >= 2 && gemm_cluster_enabled()) ? GEMM_HINT_CL2 : 0);
        if (!gemm_tuned_config(mt, a.N, kb, hint, &BN, &splits))  // exact shape measured? else a measured neighbour / the model
            gemm_pick_config(mt, a.N, a.tune_kb > 0 ? a.tune_kb : kb, hint, &BN, &splits);
        if (gemm_setup_linear(&op, a.A0, a.lda0, a.K0, a.A1, a.lda1, a.K1, static_cast<int>(a.M), a.W, a.ldw, a.N, BN,
                              splits)) {
            fail(std::string("linear setup: ") + gemm_last_error());
            return;
        }
        op.p.bias = a.bias;
        op.p.residual = a.res;
        op.p.ldr = a.ldr;
        op.p.out = a.out;
        op.p.ldc = a.ldc > 0 ? a.ldc : a.N;
        op.p.flags |= a.flags;
        op.p.alpha = a.alpha;
        op.p.hw_out = a.hw_out;
        if (a.ln_stats) {
            op.p.ln_stats = a.ln_stats;
            op.p.ln_colsum = a.ln_colsum;
            op.p.ln_chunks = a.ln_C / 32;
            op.p.ln_rows = static_cast<int>(a.M);
            op.p.ln_inv_c = 1.0f / static_cast<float>(a.ln_C);
            op.p.ln_eps = 1e-5f;
        }
        if (a.stats_out) {
            op.p.stats_out = a.stats_out;
            op.p.stats_rows = static_cast<int>(a.M);
        }
        push_gemm(op);
    }

    struct Bmm {
        const __half* A = nullptr;
        int lda = 0;
        long long a_zs1 = 0, a_zs2 = 0;
        const __half* B = nullptr;
        int ldb = 0;
        long long b_zs1 = 0, b_zs2 = 0;
        int b_mn = 0;
        int M = 0, N = 0, K = 0, nz1 = 1, nz2 = 1;
        void* out = nullptr;
        int ldc = 0;
        long long out_zs1 = 0, out_zs2 = 0;
        const float* bias = nullptr;
        const __half* res = nullptr;
        int ldr = 0;
        long long res_zs1 = 0, res_zs2 = 0;
        int flags = 0;
        float alpha = 1.0f;
        int aux = 0;
        const char* label = "bmm";
        const float2* ln_stats = nullptr;  // folded LayerNorm (rows of all batch entries: nz1 * M)
        const float* ln_colsum = nullptr;
        int ln_C = 0;
        long long bias_zs1 = 0;
        float2* stats_out = nullptr;
    };
    void bmm(const Bmm& a) {
        if (!ok) return;
        GemmOp op;
        int BN, sp;
        gemm_pick_config(((a.M + 127) / 128) * a.nz1 * a.nz2, a.N, (a.K + 63) / 64, a.b_mn ? GEMM_B_MN : 0, &BN, &sp);
        if (gemm_setup_batched(&op, a.A, a.lda, a.a_zs1, a.a_zs2, a.B, a.ldb, a.b_zs1, a.b_zs2, a.b_mn, a.M, a.N, a.K, a.nz1,
                               a.nz2, BN)) {
            fail(std::string("batched contraction setup: ") + gemm_last_error());
            return;
        }
        op.p.out = a.out;
        op.p.ldc = a.ldc;
        op.p.out_zs1 = a.out_zs1;
        op.p.out_zs2 = a.out_zs2;
        op.p.bias = a.bias;
        op.p.residual = a.res;
        op.p.ldr = a.ldr;
        op.p.res_zs1 = a.res_zs1;
        op.p.res_zs2 = a.res_zs2;
        op.p.flags |= a.flags;
        op.p.alpha = a.alpha;
        op.p.aux = a.aux;
        op.p.bias_zs1 = a.bias_zs1;
        if (a.ln_stats) {
            op.p.ln_stats = a.ln_stats;
            op.p.ln_colsum = a.ln_colsum;
            op.p.ln_chunks = a.ln_C / 32;
            op.p.ln_rows = a.M * a.nz1 * a.nz2;
            op.p.ln_inv_c = 1.0f / static_cast<float>(a.ln_C);
            op.p.ln_eps = 1e-5f;
        }
        if (a.stats_out) {
            op.p.stats_out = a.stats_out;
            op.p.stats_rows = a.M * a.nz1 * a.nz2;
        }
        push_gemm(op, a.label);
    }

    // 3x3 stride-1 pad-1 conv over one or two NHWC sources; w_name: [Cout, 9*(C0+C1)] f16
    void conv3x3_into(const Act& a0, const Act& a1, const std::string& w_name, const float* bias, const __half* res,
                      int ldr, void* out, int ldc, int flags, int hw_out) {
        if (!ok) return;
        int cout = 0, kk = 0;
        const __half* W = W16(w_name, &cout, &kk);
        if (!W) return;
        conv3x3_raw(a0, a1, W, cout, kk, prefix + w_name, bias, res, ldr, out, ldc, flags, hw_out, Act{}, Act{});
    }
    // same with an explicit weight matrix [cout, kk] and an optional fused 1x1 shortcut over s0 [| s1] (gemm_tc.h)
    void conv3x3_raw(const Act& a0, const Act& a1, const __half* W, int cout, int kk, const std::string& what,
                     const float* bias, const __half* res, int ldr, void* out, int ldc, int flags, int hw_out,
                     const Act& s0, const Act& s1) {
        if (!ok) return;
        const int C = a0.C + (a1.p ? a1.C : 0);
        const int CS0 = s0.p ? s0.C : 0, CS1 = s1.p ? s1.C : 0;
        if (kk != 9 * C + CS0 + CS1) {
            fail("conv weight '" + what + "' has K=" + std::to_string(kk) + ", expected " +
                 std::to_string(9 * C + CS0 + CS1));
            return;
        }
        GemmOp probe, op;
        if (gemm_setup_conv3x3(&probe, a0.p, a0.C, a1.p, a1.p ? a1.C : 0, a0.N, a0.H, a0.W, W, cout, 128, 1, 0, s0.p, CS0,
                               s1.p, CS1)) {
            fail(std::string("conv setup: ") + gemm_last_error());
            return;
        }
        int BN, splits;
        gemm_pick_config(probe.grid_m, cout, probe.p.num_kb,
                         flags | ((probe.grid_m >= 2 && gemm_cluster_enabled()) ? GEMM_HINT_CL2 : 0), &BN, &splits);
        if (CS0 > 0 && !gemm_tuned_config(probe.grid_m, cout, probe.p.num_kb, flags, &BN, &splits)) {
            // no measured entry for the longer K of this fused shortcut: take the tile / split of the plain convolution
            int bn0 = 0, sp0 = 0;
            gemm_pick_config(probe.grid_m, cout, 9 * (C / 64),
                             flags | ((probe.grid_m >= 2 && gemm_cluster_enabled()) ? GEMM_HINT_CL2 : 0), &bn0, &sp0);
            BN = bn0;
            splits = sp0;
        }
        if (gemm_setup_conv3x3(&op, a0.p, a0.C, a1.p, a1.p ? a1.C : 0, a0.N, a0.H, a0.W, W, cout, BN, splits, 0, s0.p, CS0,
                               s1.p, CS1)) {
            fail(std::string("conv setup: ") + gemm_last_error());
            return;
        }
        op.p.bias = bias;
        op.p.residual = res;
        op.p.ldr = ldr;
        op.p.out = out;
        op.p.ldc = ldc > 0 ? ldc : cout;
        op.p.flags |= flags;
        op.p.hw_out = hw_out;
        push_gemm(op, CS0 > 0 ? "conv3x3+shortcut" : "conv3x3");
    }
    Act conv3x3(const Act& a0, const Act& a1, const std::string& p, const float* bias_override, const __half* res) {
        int cout = 0;
        W16(p + ".weight", &cout, nullptr);
        if (!ok) return Act{};
        Act out = like(a0, cout);
        if (!ok) return out;
        conv3x3_into(a0, a1, p + ".weight", bias_override ? bias_override : F32(p + ".bias"), res, cout, out.p, cout, 0,
                     0);
        return out;
    }

    Act groupnorm(const Act& a0, const Act& a1, const std::string& p, float eps, int silu) {
        if (!ok) return Act{};
        const int C = a0.C + (a1.p ? a1.C : 0);
        Act out = like(a0, C);
        const int HW = a0.H * a0.W;
        const int groups = e.cfg_.groups;
        float* ws = static_cast<float*>(raw(gn_ws_floats(a0.N, HW, C, groups) * sizeof(float)));
        const float* g = F32(p + ".weight");
        const float* bt = F32(p + ".bias");
        if (!ok) return out;
        Engine* eng = &e;
        const __half *x0 = a0.p, *x1 = a1.p;
        const int C0 = a0.C, C1 = a1.p ? a1.C : 0, Nimg = a0.N;
        __half* o = out.p;
        plan.add(K_GROUPNORM, [=](cudaStream_t st) -> int {
            int n_launched = 1;
            if (launch_groupnorm(x0, C0, x1, C1, Nimg, HW, groups, g, bt, eps, silu, o, ws, st, &n_launched)) {
                eng->err_ = kernels_last_error();
                return -1;
            }
            return n_launched;
        }, "groupnorm rows=" + std::to_string(static_cast<long long>(Nimg) * HW) + " C=" + std::to_string(C0 + C1));
        release_raw(ws);  // stream-ordered: the next consumer of this slab runs after the norm
        return out;
    }

    void layernorm(const Act& x, const std::string& p, __half* out) {
        if (!ok) return;
        const float* g = F32(p + ".weight");
        const float* bt = F32(p + ".bias");
        if (!ok) return;
        Engine* eng = &e;
        const __half* xp = x.p;
        const int rows = static_cast<int>(x.rows()), C = x.C;
        plan.add(K_LAYERNORM, [=](cudaStream_t st) -> int {
            if (launch_layernorm(xp, rows, C, g, bt, 1e-5f, out, st)) {
                eng->err_ = kernels_last_error();
                return -1;
            }
            return 1;
        }, "layernorm rows=" + std::to_string(rows) + " C=" + std::to_string(C));
    }

    // q/k/v given as pointers with row strides; batch entries `seq` rows apart
    void attention(const __half* q, int ldq, const __half* k, int ldk, const __half* v, int ldv, __half* out, int ldo,
                   int nq, int nkv, int heads, int d, int batch, long long q_bs, long long kv_bs, long long o_bs,
                   const int* kv_index) {
        if (!ok) return;
        Engine* eng = &e;
        const float scale = 1.0f / sqrtf(static_cast<float>(d));
        const bool flash_ok = kv_index == nullptr && d <= 192 && (d % 8) == 0 && nq == nkv && nq >= 16 && ldq == ldk &&
                              ldk == ldv && q_bs == kv_bs && e.opt_flash_;
        if (nkv <= 64 && !flash_ok) {
            plan.add(K_ATTN_SMALL, [=](cudaStream_t st) -> int {
                if (launch_attn_small(q, ldq, k, ldk, v, ldv, out, ldo, nq, nkv, heads, d, batch, q_bs, kv_bs, o_bs,
                                      kv_index, scale, st)) {
                    eng->err_ = kernels_last_error();
                    return -1;
                }
                return 1;
            }, "attn_small nq=" + std::to_string(nq) + " nkv=" + std::to_string(nkv) + " heads=" + std::to_string(heads) +
                   " d=" + std::to_string(d) + " batch=" + std::to_string(batch));
            return;
        }
        if (kv_index) {
            fail("indexed key/value batches are only supported for short sequences");
            return;
        }
        if (flash_ok) {
            // tcgen05 flash attention: scores stay in TMEM
            FlashOp fop;
            if (flash_attn_setup(&fop, q, k, v, ldq, q_bs, out, ldo, o_bs, nq, heads, d, batch)) {
                fail(flash_last_error());
                return;
            }
            plan.add(K_FLASH, [fop, eng](cudaStream_t st) -> int {
                if (flash_attn_launch(&fop, st)) {
                    eng->err_ = flash_last_error();
                    return -1;
                }
                return 1;
            }, "flash_attn seq=" + std::to_string(nq) + " heads=" + std::to_string(heads) + " d=" + std::to_string(d) +
                   " batch=" + std::to_string(batch));
            return;
        }
        // scores = scale * Q K^T (fp16, materialised), row softmax, out = P V with V consumed MN-major
        const long long srows = static_cast<long long>(batch) * heads * nq;
        const int ldS = (nkv + 7) & ~7;
        __half* S = static_cast<__half*>(raw(static_cast<size_t>(srows) * ldS * sizeof(__half)));
        if (!ok) return;
        GemmOp qk, pv;
        int BN, sp;
        const int mt = ((nq + 127) / 128) * heads * batch;
        gemm_pick_config(mt, nkv, (d + 63) / 64, 0, &BN, &sp);
        if (gemm_setup_batched(&qk, q, ldq, d, q_bs, k, ldk, d, kv_bs, 0, nq, nkv, d, heads, batch, BN)) {
            fail(std::string("attention QK setup: ") + gemm_last_error());
            return;
        }
        qk.p.out = S;
        qk.p.ldc = ldS;
        qk.p.out_zs1 = static_cast<long long>(nq) * ldS;
        qk.p.out_zs2 = static_cast<long long>(heads) * nq * ldS;
        qk.p.alpha = scale;
        push_gemm(qk, "attn_qk");
        plan.add(K_SOFTMAX, [=](cudaStream_t st) -> int {
            if (launch_softmax_rows(S, srows, nkv, ldS, st)) {
                eng->err_ = kernels_last_error();
                return -1;
            }
            return 1;
        }, "softmax rows=" + std::to_string(srows) + " cols=" + std::to_string(nkv));
        gemm_pick_config(mt, d, (nkv + 63) / 64, GEMM_B_MN, &BN, &sp);
        if (gemm_setup_batched(&pv, S, ldS, static_cast<long long>(nq) * ldS, static_cast<long long>(heads) * nq * ldS,
                               v, ldv, d, kv_bs, 1, nq, d, nkv, heads, batch, BN)) {
            fail(std::string("attention PV setup: ") + gemm_last_error());
            return;
        }
        pv.p.out = out;
        pv.p.ldc = ldo;
        pv.p.out_zs1 = d;
        pv.p.out_zs2 = o_bs;
        push_gemm(pv, "attn_pv");
        release_raw(S);
    }

    // Upsample2D (nearest 2x, then conv 3x3): one contraction over the half-resolution input when the fold is on
    Act upsample_conv(const Act& x, const std::string& p) {
        if (!ok) return Act{};
        int cout = 0, kk = 0;
        const __half* W = W16(p + ".weight", &cout, &kk);
        if (!ok) return Act{};
        const long long rows_out = 4LL * x.N * x.H * x.W;
        if (!e.opt_fold_upsample_ || (x.C % 64) != 0 || kk != 9 * x.C || rows_out < e.opt_fold_upsample_rows_) {
            Act up = upsample2x(x);
            Act y = conv3x3(up, Act{}, p, nullptr, nullptr);
            release(up);
            return y;
        }
        const __half* Wst = e.upconv_weights(prefix + p, W, cout, x.C);
        const float* bias = F32(p + ".bias");
        if (!Wst || !ok) {
            ok = false;
            return Act{};
        }
        Act out = alloc(x.N, 2 * x.H, 2 * x.W, cout);
        if (!ok) return out;
        GemmOp probe, op;
        if (gemm_setup_upconv2x(&probe, x.p, x.C, x.N, x.H, x.W, Wst, cout, 128)) {
            fail(std::string("upconv setup: ") + gemm_last_error());
            return out;
        }
        int BN, splits;
        gemm_pick_config(4 * probe.grid_m, cout, probe.p.num_kb,
                         (probe.grid_m >= 2 && gemm_cluster_enabled()) ? GEMM_HINT_CL2 : 0, &BN, &splits);
        // measured (profiles/upconv_bench.py): 256-wide tiles everywhere, as CTA pairs once there are several waves of them
        BN = (4 * probe.grid_m >= 512 && cout % 256 == 0) ? (256 | GEMM_BN_PAIR) : 256;
        if (gemm_setup_upconv2x(&op, x.p, x.C, x.N, x.H, x.W, Wst, cout, BN)) {
            fail(std::string("upconv setup: ") + gemm_last_error());
            return out;
        }
        op.p.bias = bias;
        op.p.out = out.p;
        op.p.ldc = cout;
        push_gemm(op, "upconv2x");
        return out;
    }

    Act upsample2x(const Act& x) {
        if (!ok) return Act{};
        Act out = alloc(x.N, 2 * x.H, 2 * x.W, x.C);
        if (!ok) return out;
        Engine* eng = &e;
        const __half* xp = x.p;
        __half* o = out.p;
        const int N = x.N, H = x.H, W = x.W, C = x.C;
        plan.add(K_OTHER, [=](cudaStream_t st) -> int {
            if (launch_upsample2x(xp, N, H, W, C, o, st)) {
                eng->err_ = kernels_last_error();
                return -1;
            }
            return 1;
        });
        return out;
    }

    // stride-2 3x3 conv: gather + linear. pad_lo = 1 (UNet, symmetric pad 1) or 0 (VAE, F.pad (0,1,0,1))
    Act downsample(const Act& x, const std::string& p, int pad_lo) {
        if (!ok) return Act{};
        const int Ho = x.H / 2, Wo = x.W / 2;
        int cout = 0, kk = 0;
        const __half* W = W16(p + ".weight", &cout, &kk);
        if (!ok) return Act{};
        if (kk != 9 * x.C) {
            fail("downsample weight '" + prefix + p + "' has unexpected K");
            return Act{};
        }
        if (e.opt_fold_downsample_ && (x.C % 64) == 0 && (x.H % 2) == 0 && (x.W % 2) == 0) {
            // the nine taps straight from the input through four parity-view tensor maps: no im2col buffer, no gather launch
            Act out = alloc(x.N, Ho, Wo, cout);
            if (!ok) return out;
            GemmOp probe, op;
            if (gemm_setup_conv3x3_s2(&probe, x.p, x.C, x.N, x.H, x.W, W, cout, pad_lo, 128, 1)) {
                fail(std::string("stride-2 conv setup: ") + gemm_last_error());
                return out;
            }
            int BN, splits;
            const int hint = (probe.grid_m >= 2 && gemm_cluster_enabled()) ? GEMM_HINT_CL2 : 0;
            if (!gemm_tuned_config(probe.grid_m, cout, probe.p.num_kb, hint, &BN, &splits))
                gemm_pick_config(probe.grid_m, cout, probe.p.num_kb, hint, &BN, &splits);
            if (gemm_setup_conv3x3_s2(&op, x.p, x.C, x.N, x.H, x.W, W, cout, pad_lo, BN, splits)) {
                fail(std::string("stride-2 conv setup: ") + gemm_last_error());
                return out;
            }
            op.p.bias = F32(p + ".bias");
            op.p.out = out.p;
            op.p.ldc = cout;
            push_gemm(op, "conv3x3s2");
            return out;
        }
        Act col = alloc(x.N, Ho, Wo, 9 * x.C);
        Engine* eng = &e;
        const __half* xp = x.p;
        __half* cp = col.p;
        const int N = x.N, H = x.H, Wd = x.W, C = x.C;
        plan.add(K_OTHER, [=](cudaStream_t st) -> int {
            if (launch_im2col_s2(xp, N, H, Wd, C, pad_lo, Ho, Wo, cp, st)) {
                eng->err_ = kernels_last_error();
                return -1;
            }
            return 1;
        });
        Act out = alloc(x.N, Ho, Wo, cout);
        if (!ok) return out;
        Lin l;
        l.A0 = col.p;
        l.lda0 = 9 * x.C;
        l.K0 = 9 * x.C;
        l.M = col.rows();
        l.W = W;
        l.ldw = kk;
        l.N = cout;
        l.bias = F32(p + ".bias");
        l.out = out.p;
        linear(l);
        release(col);
        return out;
    }

    // ResnetBlock2D; `skip` is the second source of a channel concat (p == nullptr when absent).
    Act resnet(const std::string& p, const Act& x, const Act& skip, float eps, const float* conv1_bias_override) {
        if (!ok) return Act{};
        const int cin = x.C + (skip.p ? skip.C : 0);
        Act n1 = groupnorm(x, skip, p + ".norm1", eps, 1);
        Act h = conv3x3(n1, Act{}, p + ".conv1", conv1_bias_override, nullptr);
        release(n1);
        Act n2 = groupnorm(h, Act{}, p + ".norm2", eps, 1);
        release(h);
        if (!ok) return Act{};
        const int cout = n2.C;
        Act sc;
        const __half* scp = x.p;
        if (e.find(prefix + p + ".conv_shortcut.weight") && e.opt_fuse_shortcut_ && (x.C % 64) == 0 &&
            (!skip.p || (skip.C % 64) == 0)) {
            // conv_shortcut (1x1 over x [| skip]) rides along conv2 as extra k-blocks at the centre tap: one launch and one
            // fp16 round trip of the shortcut activation less, and the sum is taken in the fp32 accumulator
            const Engine::FusedShortcut* f = e.fused_shortcut(prefix + p, cout, cin);
            if (!f) {
                fail(e.err_);
                return Act{};
            }
            Act out = like(x, cout);
            if (ok)
                conv3x3_raw(n2, Act{}, f->W, cout, 9 * cout + cin, prefix + p + ".conv2+conv_shortcut", f->bias, nullptr, 0,
                            out.p, cout, 0, 0, x, skip);
            release(n2);
            return out;
        }
        if (e.find(prefix + p + ".conv_shortcut.weight")) {
            sc = like(x, cout);
            int r = 0, c = 0;
            Lin l;
            l.A0 = x.p;
            l.lda0 = x.C;
            l.K0 = x.C;
            if (skip.p) {
                l.A1 = skip.p;
                l.lda1 = skip.C;
                l.K1 = skip.C;
            }
            l.M = x.rows();
            l.W = W16(p + ".conv_shortcut.weight", &r, &c);
            l.ldw = c;
            l.N = cout;
            l.bias = F32(p + ".conv_shortcut.bias");
            l.out = sc.p;
            if (ok && c != cin) fail("shortcut weight '" + prefix + p + "' has unexpected K");
            linear(l);
            scp = sc.p;
        } else if (cin != cout) {
            fail("resnet '" + prefix + p + "' changes channels but has no conv_shortcut");
        }
        Act out = conv3x3(n2, Act{}, p + ".conv2", nullptr, scp);
        release(n2);
        if (sc.p) release(sc);
        return out;
    }
};

// ------------------------------------------------------------------------------------------------------------
// Engine: weights
// ------------------------------------------------------------------------------------------------------------
Engine::Engine(const dtp_config& cfg) : cfg_(cfg) {
    if (cfg_.arena_bytes == 0) cfg_.arena_bytes = 8ull << 30;
    if (cfg_.enc_tokens <= 0) cfg_.enc_tokens = 14;
    device_ = kctx_device();
    kctx_ = kctx_create();  // on the device current at dtp_create (include/dtp.h: a handle is bound to that device)
    if (!kctx_) err_ = "allocation of the kernel synchronisation state failed";
}

Engine::~Engine() {
    cudaDeviceSynchronize();
    for (cudaEvent_t e : stage_pool_) cudaEventDestroy(e);
    for (cudaEvent_t e : prof_.pool) cudaEventDestroy(e);
    if (gstream_) cudaStreamDestroy(gstream_);
    if (gev_in_) cudaEventDestroy(gev_in_);
    if (gev_out_) cudaEventDestroy(gev_out_);
    kctx_destroy(kctx_);
    if (g_infer_.exec) cudaGraphExecDestroy(g_infer_.exec);
    if (g_stamp_.exec) cudaGraphExecDestroy(g_stamp_.exec);
    for (auto& kv : w_)
        if (kv.second.dev) cudaFree(kv.second.dev);
    for (void* p : persistent_) cudaFree(p);
    if (arena_base_) cudaFree(arena_base_);
    if (ws_) cudaFree(ws_);
}

const WT* Engine::find(const std::string& name) {
    auto it = w_.find(name);
    return it == w_.end() ? nullptr : &it->second;
}

void* Engine::persistent(size_t bytes, bool zero) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        err_ = "cudaMalloc(" + std::to_string(bytes) + ") failed";
        return nullptr;
    }
    if (zero) cudaMemset(p, 0, bytes);
    persistent_.push_back(p);
    return p;
}

int Engine::ensure_arena() {
    if (arena_base_) return 0;
    if (cudaMalloc(&arena_base_, cfg_.arena_bytes) != cudaSuccess)
        return fail("cudaMalloc of the " + std::to_string(cfg_.arena_bytes >> 20) + " MiB activation arena failed");
    arena_.init(arena_base_, cfg_.arena_bytes);
    return 0;
}

int Engine::grow_arena(size_t bytes) {
    if (bytes <= cfg_.arena_bytes && arena_base_) return 0;
    cudaDeviceSynchronize();
    if (arena_base_) cudaFree(arena_base_);
    arena_base_ = nullptr;
    if (bytes > cfg_.arena_bytes) cfg_.arena_bytes = bytes;
    // every plan holds pointers into the old slab
    unet_plan_.clear();
    vae_enc_plan_.clear();
    vae_dec_plan_.clear();
    enc_plan_.clear();
    cond_plan_.clear();
    g_infer_.key.clear();
    g_stamp_.key.clear();
    return ensure_arena();
}

const Engine::FusedShortcut* Engine::fused_shortcut(const std::string& prefix, int cout, int cin) {
    auto it = fused_sc_.find(prefix);
    if (it != fused_sc_.end()) return &it->second;
    const WT* w2 = find(prefix + ".conv2.weight");
    const WT* ws = find(prefix + ".conv_shortcut.weight");
    const WT* b2 = find(prefix + ".conv2.bias");
    const WT* bs = find(prefix + ".conv_shortcut.bias");
    if (!w2 || !ws || !b2 || !bs || w2->dtype != 1 || ws->dtype != 1 || w2->shape.size() != 2 || ws->shape.size() != 2 ||
        w2->shape[0] != cout || w2->shape[1] != 9LL * cout || ws->shape[0] != cout || ws->shape[1] != cin ||
        b2->host.size() != static_cast<size_t>(cout) || bs->host.size() != static_cast<size_t>(cout)) {
        fail("fused shortcut: unexpected conv2 / conv_shortcut tensors under '" + prefix + "'");
        return nullptr;
    }
    const size_t K2 = 9 * static_cast<size_t>(cout), KT = K2 + cin;
    FusedShortcut f;
    f.W = static_cast<__half*>(persistent(static_cast<size_t>(cout) * KT * 2, false));
    f.bias = static_cast<float*>(persistent(static_cast<size_t>(cout) * 4, false));
    if (!f.W || !f.bias) return nullptr;
    std::vector<float> bsum(cout);
    for (int i = 0; i < cout; ++i) bsum[i] = b2->host[i] + bs->host[i];
    if (cudaMemcpy2D(f.W, KT * 2, w2->dev, K2 * 2, K2 * 2, cout, cudaMemcpyDeviceToDevice) != cudaSuccess ||
        cudaMemcpy2D(f.W + K2, KT * 2, ws->dev, static_cast<size_t>(cin) * 2, static_cast<size_t>(cin) * 2, cout,
                     cudaMemcpyDeviceToDevice) != cudaSuccess ||
        cudaMemcpy(f.bias, bsum.data(), static_cast<size_t>(cout) * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        fail("fused shortcut: weight concatenation failed");
        return nullptr;
    }
    return &fused_sc_.emplace(prefix, f).first->second;
}

const __half* Engine::upconv_weights(const std::string& key, const __half* W, int cout, int cin) {
    auto it = upconv_w_.find(key);
    if (it != upconv_w_.end()) return it->second;
    __half* wst = static_cast<__half*>(persistent(16ull * cout * cin * sizeof(__half), false));
    if (!wst) return nullptr;
    if (launch_upconv_fold_weights(W, cout, cin, wst, nullptr) || cudaDeviceSynchronize() != cudaSuccess) {
        fail(std::string("upsample-conv weight fold failed: ") + kernels_last_error());
        return nullptr;
    }
    upconv_w_.emplace(key, wst);
    return wst;
}

int Engine::check_device_error() {
    const int v = kctx_take_error(kctx_);
    if (v == 0) return 0;
    return fail(std::string("a cross-CTA wait (") + (v == 1 ? "GroupNorm barrier" : "split-K ticket") +
                ") timed out in an earlier call: the kernel's CTAs were not co-resident (is the GPU shared with another "
                "process or stream?); the results of that call are invalid");
}

void Engine::stage_begin(int stage, cudaStream_t st) {
    if (opt_nvtx_) {
        static const char* names[ST_NUM] = {"canvas_preprocess", "vae_encoder", "unet", "latent_step", "vae", "composite"};
        nvtxRangePushA(names[stage]);
    }
    if (!opt_stage_timers_) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
    while (stage_next_ + 2 > stage_pool_.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        stage_pool_.push_back(e);
    }
    StageRec r{stage, stage_pool_[stage_next_], stage_pool_[stage_next_ + 1]};
    stage_next_ += 2;
    cudaEventRecord(r.a, st);
    stage_recs_.push_back(r);
}

void Engine::stage_end(int stage, cudaStream_t st) {
    if (opt_nvtx_) nvtxRangePop();
    if (!opt_stage_timers_) return;
    for (size_t i = stage_recs_.size(); i-- > 0;)
        if (stage_recs_[i].stage == stage) {
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone)
                cudaEventRecord(stage_recs_[i].b, st);
            return;
        }
}

void Engine::stage_collect() {
    if (stage_recs_.empty()) return;
    cudaEventSynchronize(stage_recs_.back().b);
    for (auto& r : stage_recs_) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            stage_us_[r.stage] += 1000.0 * ms;
            stage_n_[r.stage] += 1;
        } else {
            cudaGetLastError();
        }
    }
    stage_recs_.clear();
    stage_next_ = 0;
}

int Engine::ensure_ws() {
    if (ws_needed_ <= ws_bytes_) return 0;
    if (ws_) {
        cudaDeviceSynchronize();
        cudaFree(ws_);
    }
    ws_ = nullptr;
    ws_bytes_ = 0;
    const size_t want = ws_needed_ + (ws_needed_ >> 2);
    if (cudaMalloc(&ws_, want) != cudaSuccess) return fail("cudaMalloc of the split-K workspace failed");
    ws_bytes_ = want;
    return 0;
}

int Engine::set_tensor(const char* name, const void* host, const int64_t* shape, int ndim, int dtype) {
    if (!name || !host || ndim < 0 || ndim > 8 || (dtype != 0 && dtype != 1)) return fail("set_tensor: bad arguments");
    WT t;
    t.dtype = dtype;
    t.shape.assign(shape, shape + ndim);
    const size_t bytes = static_cast<size_t>(t.numel()) * (dtype == 0 ? 4 : 2);
    if (cudaMalloc(&t.dev, bytes ? bytes : 16) != cudaSuccess)
        return fail(std::string("set_tensor: cudaMalloc failed for ") + name);
    if (cudaMemcpy(t.dev, host, bytes, cudaMemcpyDefault) != cudaSuccess) {
        cudaFree(t.dev);
        return fail(std::string("set_tensor: copy failed for ") + name);
    }
    if (dtype == 0 && t.numel() <= (1 << 22)) {
        t.host.resize(t.numel());
        cudaMemcpy(t.host.data(), host, bytes, cudaMemcpyDefault);  // host or device source (UVA)
    }
    auto it = w_.find(name);
    if (it != w_.end()) {
        cudaFree(it->second.dev);
        w_.erase(it);
    }
    w_.emplace(name, std::move(t));
    finalized_ = false;
    return 0;
}

// UNet traversal shared by finalize (inventory) and the plan builder: calls back for each resnet / transformer prefix.
template <typename FR, typename FT>
static void walk_unet(const dtp_config& c, FR on_resnet, FT on_transformer) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < c.unet_layers_per_block; ++j) {
            on_resnet("down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j));
            if (c.unet_down_attn[i]) on_transformer("down_blocks." + std::to_string(i) + ".attentions." + std::to_string(j));
        }
    on_resnet(std::string("mid_block.resnets.0"));
    on_transformer(std::string("mid_block.attentions.0"));
    on_resnet(std::string("mid_block.resnets.1"));
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < c.unet_layers_per_block + 1; ++j) {
            on_resnet("up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j));
            if (c.unet_down_attn[3 - i]) on_transformer("up_blocks." + std::to_string(i) + ".attentions." + std::to_string(j));
        }
}

int Engine::finalize_weights() {
    if (finalized_) return 0;
    // inventory of UNet resnets (time-embedding row layout) and transformers (cross-attention K/V slots)
    resnets_.clear();
    tf_names_.clear();
    temb_total_ = 0;
    bool missing = false;
    std::string first_missing;
    walk_unet(
        cfg_,
        [&](const std::string& p) {
            const WT* b = find("unet." + p + ".conv1.bias");
            const WT* tb = find("unet." + p + ".time_emb_proj.bias");
            if (!b || !tb || b->host.empty() || tb->host.size() != b->host.size()) {
                if (!missing) first_missing = "unet." + p + ".{conv1,time_emb_proj}.bias";
                missing = true;
                return;
            }
            resnets_.push_back({p, temb_total_});
            temb_total_ += static_cast<int>(b->host.size());
        },
        [&](const std::string& p) { tf_names_.push_back(p); });
    if (missing) return fail("finalize_weights: missing " + first_missing);
    // combined bias conv1.bias + time_emb_proj.bias (the per-step projection is added by the schedule tables)
    std::vector<float> comb(temb_total_);
    for (auto& r : resnets_) {
        const WT* b = find("unet." + r.first + ".conv1.bias");
        const WT* tb = find("unet." + r.first + ".time_emb_proj.bias");
        for (size_t i = 0; i < b->host.size(); ++i) comb[r.second + i] = b->host[i] + tb->host[i];
    }
    {
        WT t;
        t.dtype = 0;
        t.shape = {temb_total_};
        if (cudaMalloc(&t.dev, comb.size() * 4) != cudaSuccess) return fail("finalize_weights: cudaMalloc failed");
        cudaMemcpy(t.dev, comb.data(), comb.size() * 4, cudaMemcpyHostToDevice);
        auto it = w_.find("unet.__temb_bias_comb");
        if (it != w_.end()) {
            cudaFree(it->second.dev);
            w_.erase(it);
        }
        w_.emplace("unet.__temb_bias_comb", std::move(t));
    }
    temb_cur_ = static_cast<float*>(persistent(static_cast<size_t>(temb_total_) * 4, true));
    if (!temb_cur_) return -1;
    cross_kv_.assign(tf_names_.size(), nullptr);
    for (size_t i = 0; i < tf_names_.size(); ++i) {
        const WT* kvw = find("unet." + tf_names_[i] + ".transformer_blocks.0.attn2.to_kv.weight");
        if (!kvw) return fail("finalize_weights: missing unet." + tf_names_[i] + ".transformer_blocks.0.attn2.to_kv.weight");
        cross_kv_[i] = static_cast<__half*>(persistent(static_cast<size_t>(2 * cfg_.enc_tokens) * kvw->shape[0] * 2, true));
        if (!cross_kv_[i]) return -1;
    }
    wscore_.assign(tf_names_.size(), nullptr);
    wout_.assign(tf_names_.size(), nullptr);
    for (size_t i = 0; i < tf_names_.size(); ++i) {
        const WT* qw = find("unet." + tf_names_[i] + ".transformer_blocks.0.attn2.to_q.weight");
        if (!qw) return fail("finalize_weights: missing unet." + tf_names_[i] + ".transformer_blocks.0.attn2.to_q.weight");
        const size_t C = static_cast<size_t>(qw->shape[0]), HP = static_cast<size_t>(cfg_.unet_heads) * 16;
        wscore_[i] = static_cast<__half*>(persistent(3 * HP * C * 2, true));
        wout_[i] = static_cast<__half*>(persistent(3 * C * HP * 2, true));
        if (!wscore_[i] || !wout_[i]) return -1;
    }
    ctx_ = static_cast<__half*>(persistent(static_cast<size_t>(2 * cfg_.enc_tokens) * cfg_.unet_cross_dim * 2, true));
    ctx_f32_ = static_cast<float*>(persistent(static_cast<size_t>(2 * cfg_.enc_tokens) * cfg_.unet_cross_dim * 4, true));
    if (!ctx_ || !ctx_f32_) return -1;
    unet_plan_.clear();
    vae_enc_plan_.clear();
    vae_dec_plan_.clear();
    enc_plan_.clear();
    cond_plan_.clear();
    temb_dirty_ = true;
    cond_set_ = false;
    fused_sc_.clear();
    upconv_w_.clear();
    if (prepare_ln_fold()) return -1;
    if (prepare_ff_out()) return -1;
    finalized_ = true;
    return 0;
}

// [Wp W2 | Wp] and Wp b2 + bp per transformer layer (runtime.h, FfOut); the product runs on the contraction kernel itself
int Engine::prepare_ff_out() {
    ffo_.clear();
    if (ensure_arena()) return -1;
    std::vector<FfOut> v(tf_names_.size());
    Plan plan;
    Builder b(*this, plan, "unet.");
    for (size_t i = 0; i < tf_names_.size(); ++i) {
        const std::string t = tf_names_[i] + ".transformer_blocks.0";
        int C = 0, k2 = 0, cp = 0, kp = 0;
        const __half* W2 = b.W16(t + ".ff.net.2.weight", &C, &k2);
        const __half* Wp = b.W16(tf_names_[i] + ".proj_out.weight", &cp, &kp);
        const float* b2 = b.F32(t + ".ff.net.2.bias");
        const float* bp = b.F32(tf_names_[i] + ".proj_out.bias");
        if (!b.ok) {
            b.ok = true;  // a checkpoint without these tensors simply keeps the two-launch tail
            return 0;
        }
        if (k2 != 4 * C || cp != C || kp != C || (C % 8) != 0) return 0;
        FfOut& f = v[i];
        f.W = static_cast<__half*>(persistent(static_cast<size_t>(C) * 5 * C * 2, false));
        f.bias = static_cast<float*>(persistent(static_cast<size_t>(C) * 4, false));
        if (!f.W || !f.bias) return -1;
        Builder::Bmm g;  // (Wp W2)[n, k4] = sum_c Wp[n, c] W2[c, k4]: W2 consumed as an MN-major B operand
        g.A = Wp; g.lda = C; g.B = W2; g.ldb = 4 * C; g.b_mn = 1;
        g.M = C; g.N = 4 * C; g.K = C; g.out = f.W; g.ldc = 5 * C; g.label = "ff_out_fold";
        b.bmm(g);
        if (!b.ok) return -1;
        if (cudaMemcpy2DAsync(f.W + 4 * C, static_cast<size_t>(5) * C * 2, Wp, static_cast<size_t>(C) * 2,
                              static_cast<size_t>(C) * 2, C, cudaMemcpyDeviceToDevice, 0) != cudaSuccess)
            return fail("ff_out fold: copy failed");
        if (launch_ln_fold_weights(Wp, C, C, C, nullptr, b2, bp, nullptr, nullptr, f.bias, 0))
            return fail(std::string("ff_out fold: ") + kernels_last_error());
    }
    if (ensure_ws()) return -1;
    if (plan.run(0, &launches_, nullptr)) return -1;
    if (cudaStreamSynchronize(0) != cudaSuccess) return fail("ff_out fold: device error");
    ffo_ = std::move(v);
    return 0;
}

// gamma-scaled copies of the weights behind the three LayerNorms of every transformer block, their row sums and W beta
int Engine::prepare_ln_fold() {
    ln_.clear();
    std::vector<LnFold> v(tf_names_.size());
    const int heads = cfg_.unet_heads, HP = heads * 16;
    for (size_t i = 0; i < tf_names_.size(); ++i) {
        const std::string t = "unet." + tf_names_[i] + ".transformer_blocks.0";
        const WT* qkv = find(t + ".attn1.to_qkv.weight");
        const WT* ff1 = find(t + ".ff.net.0.proj.weight");
        const WT* ff1b = find(t + ".ff.net.0.proj.bias");
        const WT* q = find(t + ".attn2.to_q.weight");
        const WT *g1 = find(t + ".norm1.weight"), *b1 = find(t + ".norm1.bias");
        const WT *g2 = find(t + ".norm2.weight"), *b2 = find(t + ".norm2.bias");
        const WT *g3 = find(t + ".norm3.weight"), *b3 = find(t + ".norm3.bias");
        if (!qkv || !ff1 || !ff1b || !q || !g1 || !b1 || !g2 || !b2 || !g3 || !b3) return 0;  // folding stays off
        const int C = static_cast<int>(q->shape[0]);
        if (qkv->dtype != 1 || ff1->dtype != 1 || q->dtype != 1 || qkv->shape[1] != C || ff1->shape[1] != C) return 0;
        LnFold& f = v[i];
        const size_t n3 = static_cast<size_t>(qkv->shape[0]), n8 = static_cast<size_t>(ff1->shape[0]);
        f.qkv_w = static_cast<__half*>(persistent(n3 * C * 2, false));
        f.qkv_cs = static_cast<float*>(persistent(n3 * 4, true));
        f.qkv_b = static_cast<float*>(persistent(n3 * 4, true));
        f.ff1_w = static_cast<__half*>(persistent(n8 * C * 2, false));
        f.ff1_cs = static_cast<float*>(persistent(n8 * 4, true));
        f.ff1_b = static_cast<float*>(persistent(n8 * 4, true));
        f.q_w = static_cast<__half*>(persistent(static_cast<size_t>(C) * C * 2, false));
        f.q_beta = static_cast<float*>(persistent(static_cast<size_t>(C) * 4, true));
        f.wscore = static_cast<__half*>(persistent(static_cast<size_t>(3) * HP * C * 2, true));
        f.ws_cs = static_cast<float*>(persistent(static_cast<size_t>(3) * HP * 4, true));
        f.ws_b = static_cast<float*>(persistent(static_cast<size_t>(3) * HP * 4, true));
        if (!f.qkv_w || !f.qkv_cs || !f.qkv_b || !f.ff1_w || !f.ff1_cs || !f.ff1_b || !f.q_w || !f.q_beta || !f.wscore ||
            !f.ws_cs || !f.ws_b)
            return -1;
        auto H = [](const WT* w) { return static_cast<const __half*>(w->dev); };
        auto F = [](const WT* w) { return static_cast<const float*>(w->dev); };
        if (launch_ln_fold_weights(H(qkv), static_cast<int>(n3), C, C, F(g1), F(b1), nullptr, f.qkv_w, f.qkv_cs, f.qkv_b, 0) ||
            launch_ln_fold_weights(H(ff1), static_cast<int>(n8), C, C, F(g3), F(b3), F(ff1b), f.ff1_w, f.ff1_cs, f.ff1_b, 0) ||
            launch_ln_fold_weights(H(q), C, C, C, F(g2), F(b2), nullptr, f.q_w, nullptr, f.q_beta, 0))
            return fail(std::string("LayerNorm weight folding failed: ") + kernels_last_error());
        launches_ += 3;
    }
    if (cudaStreamSynchronize(0) != cudaSuccess) return fail("LayerNorm weight folding: device error");
    ln_ = std::move(v);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// UNet plan
// ------------------------------------------------------------------------------------------------------------
// dedup: the uncond and the cond branch of a stamp see the same UNet sample input and differ only in their cross-attention
// context (inpaint_pipeline.py:116-138: mask = [m, m, ctx_m], masked_latents = [l, l, ctx_l], embeddings = [neg, prompt,
// prompt]), so everything in front of the FIRST cross-attention of the network is computed once for both: x then holds the
// samples [cond | texture-guidance] only, the fused cross-attention kernel reads row group max(z - 1, 0) and writes all three
// groups, and the block returns [uncond | cond | texture-guidance]. Exact (the two branches were bitwise equal before).
static Act transformer2d(Builder& b, Engine& e, const dtp_config& cfg, const std::string& p, const Act& x, __half* kv,
                         const int* kv_index, bool dedup = false) {
    if (!b.ok) return Act{};
    const int C = x.C, heads = cfg.unet_heads, d = C / heads;
    const int seq = x.H * x.W, batch = x.N;
    const long long rows = x.rows();
    const int batch_out = dedup ? batch / 2 * 3 : batch;
    const long long rows_out = dedup ? rows / 2 * 3 : rows;
    const std::string t = p + ".transformer_blocks.0";
    const int tf = e.tf_index(p);
    // The three LayerNorms of the block are folded into the contractions that consume them (gemm_tc.h; the reference fuses
    // each into one plugin op in front of the GEMM, models.py:304-365): the producer of h emits per-row partial sums, the
    // consumer multiplies the raw rows by gamma-scaled weights and normalises in its epilogue. Needs C % 32 == 0.
    const bool fl = e.fold_ln() && (C % 32) == 0 && e.fold_cross();
    // The GEGLU contraction is epilogue-bound at the wide levels (6.5 tiles of 128 x 256 per CTA at 12 288 rows): there the
    // folded epilogue costs more than the LayerNorm kernel it removes (measured +12 us vs -10.5 us), so norm3 is folded
    // only below fold_ln_ff_rows rows (norm1 / norm2 are folded everywhere: QKV +6 us / scores +2.5 us vs -10.5 us each)
    const bool fl3 = fl && rows_out <= e.fold_ln_ff_rows();
    float2* st = fl ? static_cast<float2*>(b.raw(static_cast<size_t>(rows) * (C / 32) * sizeof(float2))) : nullptr;
    // dedup: the row statistics behind the cross-attention are laid out for three row groups
    float2* st3 = (dedup && fl3) ? static_cast<float2*>(b.raw(static_cast<size_t>(rows_out) * (C / 32) * sizeof(float2))) : st;
    Act x3 = x;  // residual of the block's output
    if (dedup && b.ok) {
        x3 = b.alloc(batch_out, x.H, x.W, C);
        if (!b.ok) return Act{};
        Engine* eng = &e;
        const __half* src = x.p;
        __half* dst = x3.p;
        const int Bs = batch / 2;
        const long long per_sample = static_cast<long long>(seq) * C;
        b.plan.add(K_OTHER, [=](cudaStream_t s) -> int {
            if (launch_expand_branches(src, dst, Bs, per_sample, s)) {
                eng->set_error(kernels_last_error());
                return -1;
            }
            return 1;
        }, "expand_branches");
    }
    Act n = b.groupnorm(x, Act{}, p + ".norm", 1e-6f, 0);
    Act h = b.like(x, C);
    int wr = 0, wc = 0;
    {
        Lin l;
        l.A0 = n.p; l.lda0 = C; l.K0 = C; l.M = rows;
        l.W = b.W16(p + ".proj_in.weight", &wr, &wc); l.ldw = C; l.N = C;
        l.bias = b.F32(p + ".proj_in.bias"); l.out = h.p; l.stats_out = st;
        b.linear(l);
    }
    b.release(n);
    // self-attention
    Act tmp;
    if (!fl) {
        tmp = b.like(x, C);
        b.layernorm(h, t + ".norm1", tmp.p);
    }
    Act qkv = b.like(x, 3 * C);
    {
        Lin l;
        l.A0 = fl ? h.p : tmp.p; l.lda0 = C; l.K0 = C; l.M = rows;
        l.W = fl ? e.ln(tf).qkv_w : b.W16(t + ".attn1.to_qkv.weight"); l.ldw = C; l.N = 3 * C; l.out = qkv.p;
        if (fl) {
            l.ln_stats = st; l.ln_colsum = e.ln(tf).qkv_cs; l.ln_C = C; l.bias = e.ln(tf).qkv_b;
        }
        b.linear(l);
    }
    if (!fl) b.release(tmp);
    Act att = b.like(x, C);
    if (b.ok)
        b.attention(qkv.p, 3 * C, qkv.p + C, 3 * C, qkv.p + 2 * C, 3 * C, att.p, C, seq, seq, heads, d, batch,
                    static_cast<long long>(seq) * 3 * C, static_cast<long long>(seq) * 3 * C,
                    static_cast<long long>(seq) * C, nullptr);
    b.release(qkv);
    {
        Lin l;
        l.A0 = att.p; l.lda0 = C; l.K0 = C; l.M = rows;
        l.W = b.W16(t + ".attn1.to_out.0.weight"); l.ldw = C; l.N = C;
        l.bias = b.F32(t + ".attn1.to_out.0.bias"); l.res = h.p; l.ldr = C; l.out = h.p; l.stats_out = st;
        b.linear(l);
    }
    b.release(att);
    // cross-attention on the image tokens. K / V are constant per brush, so the query projection is folded into the keys
    // (scores = LN2(h) (scale K_h Wq_h)^T) and the output projection into the values (out = P (V_h Wo_h^T)), both
    // prepared by the condition plan: two skinny contractions, softmax over the 14 tokens in the first one's epilogue.
    // Sample groups [uncond | cond | texture-guidance] pick their context through the batch coordinate.
    if (!fl) {
        tmp = b.like(x, C);
        b.layernorm(h, t + ".norm2", tmp.p);
    }
    const int T = cfg.enc_tokens;
    if (fl && e.fuse_cross() && heads * 16 == 128 && (C % 64) == 0) {
        // score and output contraction as ONE kernel: the probabilities stay in shared memory (cross_attn.cu)
        CrossOp cop;
        const long long grp_rows = dedup ? rows / 2 : rows / 3;
        // few row tiles (narrow levels): out of place, so that the output chunks can be spread over CTAs
        const bool split = dedup || (((grp_rows + 127) / 128) * 3 * 2 <= 148 && C > 256);
        Act h2 = split ? b.alloc(batch_out, x.H, x.W, C) : h;
        if (!b.ok) return Act{};
        if (cross_attn_setup(&cop, h.p, h2.p, static_cast<int>(grp_rows), C, T, e.ln(tf).wscore, e.wout(tf), st,
                             e.ln(tf).ws_cs, e.ln(tf).ws_b, b.F32(t + ".attn2.to_out.0.bias"), fl3 ? st3 : nullptr,
                             dedup ? 1 : 0)) {
            b.fail(std::string("cross attention setup: ") + cross_last_error());
        } else if (b.ok) {
            Engine* eng = &e;
            b.plan.add(K_GEMM, [cop, eng](cudaStream_t s) -> int {
                if (cross_attn_launch(&cop, s)) {
                    eng->set_error(cross_last_error());
                    return -1;
                }
                return 1;
            }, "cross_attn rows=" + std::to_string(grp_rows) + " C=" + std::to_string(C) + " z=3 csplit=" +
                   std::to_string(cop.csplit));
        }
        if (split) {
            b.release(h);
            h = h2;
        }
    } else if (dedup) {
        b.fail("branch de-duplication needs the fused cross-attention kernel");
        return Act{};
    } else if (e.fold_cross()) {
        const int HP = heads * 16;
        const long long grp_rows = rows / 3;
        Act P = b.like(x, HP);
        if (b.ok) {
            Builder::Bmm g;
            g.A = fl ? h.p : tmp.p; g.lda = C; g.a_zs1 = grp_rows * C;
            g.B = fl ? e.ln(tf).wscore : e.wscore(tf); g.ldb = C; g.b_zs1 = static_cast<long long>(HP) * C;
            g.M = static_cast<int>(grp_rows); g.N = HP; g.K = C; g.nz1 = 3;
            g.out = P.p; g.ldc = HP; g.out_zs1 = grp_rows * HP;
            g.flags = EPI_SOFTMAX16; g.aux = T; g.label = "cross_scores";
            if (fl) {
                g.ln_stats = st; g.ln_colsum = e.ln(tf).ws_cs; g.ln_C = C; g.bias = e.ln(tf).ws_b; g.bias_zs1 = HP;
            }
            b.bmm(g);
        }
        if (!fl) b.release(tmp);
        if (b.ok) {
            Builder::Bmm g;
            g.A = P.p; g.lda = HP; g.a_zs1 = grp_rows * HP;
            g.B = e.wout(tf); g.ldb = HP; g.b_zs1 = static_cast<long long>(C) * HP;
            g.M = static_cast<int>(grp_rows); g.N = C; g.K = HP; g.nz1 = 3;
            g.out = h.p; g.ldc = C; g.out_zs1 = grp_rows * C;
            g.bias = b.F32(t + ".attn2.to_out.0.bias");
            g.res = h.p; g.ldr = C; g.res_zs1 = grp_rows * C; g.label = "cross_out"; g.stats_out = fl3 ? st : nullptr;
            b.bmm(g);
        }
        b.release(P);
    } else {
        Act q = b.like(x, C);
        {
            Lin l;
            l.A0 = tmp.p; l.lda0 = C; l.K0 = C; l.M = rows;
            l.W = b.W16(t + ".attn2.to_q.weight"); l.ldw = C; l.N = C; l.out = q.p;
            b.linear(l);
        }
        b.release(tmp);
        att = b.like(x, C);
        if (b.ok)
            b.attention(q.p, C, kv, 2 * C, kv + C, 2 * C, att.p, C, seq, T, heads, d, batch,
                        static_cast<long long>(seq) * C, static_cast<long long>(T) * 2 * C,
                        static_cast<long long>(seq) * C, kv_index);
        b.release(q);
        {
            Lin l;
            l.A0 = att.p; l.lda0 = C; l.K0 = C; l.M = rows;
            l.W = b.W16(t + ".attn2.to_out.0.weight"); l.ldw = C; l.N = C;
            l.bias = b.F32(t + ".attn2.to_out.0.bias"); l.res = h.p; l.ldr = C; l.out = h.p;
            b.linear(l);
        }
        b.release(att);
    }
    // GEGLU feed-forward (from here on every tensor has the output's sample count)
    if (!fl3) {
        tmp = b.like(h, C);
        b.layernorm(h, t + ".norm3", tmp.p);
    }
    Act g = b.like(h, 4 * C);
    {
        Lin l;
        l.A0 = fl3 ? h.p : tmp.p; l.lda0 = C; l.K0 = C; l.M = rows_out;
        l.W = fl3 ? e.ln(tf).ff1_w : b.W16(t + ".ff.net.0.proj.weight"); l.ldw = C; l.N = 8 * C;
        l.bias = fl3 ? e.ln(tf).ff1_b : b.F32(t + ".ff.net.0.proj.bias"); l.out = g.p; l.ldc = 4 * C; l.flags = EPI_GEGLU;
        if (fl3) {
            l.ln_stats = st3; l.ln_colsum = e.ln(tf).ff1_cs; l.ln_C = C;
        }
        b.linear(l);
    }
    if (!fl3) b.release(tmp);
    Act out = b.like(h, C);
    if (e.fuse_ff_out()) {
        // ff.net.2 + residual + proj_out + residual as ONE contraction over [g | h] (runtime.h, FfOut)
        Lin l;
        l.A0 = g.p; l.lda0 = 4 * C; l.K0 = 4 * C; l.A1 = h.p; l.lda1 = C; l.K1 = C; l.M = rows_out;
        l.W = e.ffo(tf).W; l.ldw = 5 * C; l.N = C; l.bias = e.ffo(tf).bias;
        l.res = x3.p; l.ldr = C; l.out = out.p; l.tune_kb = 4 * C / 64;
        b.linear(l);
        b.release(g);
    } else {
        {
            Lin l;
            l.A0 = g.p; l.lda0 = 4 * C; l.K0 = 4 * C; l.M = rows_out;
            l.W = b.W16(t + ".ff.net.2.weight"); l.ldw = 4 * C; l.N = C;
            l.bias = b.F32(t + ".ff.net.2.bias"); l.res = h.p; l.ldr = C; l.out = h.p;
            b.linear(l);
        }
        b.release(g);
        {
            Lin l;
            l.A0 = h.p; l.lda0 = C; l.K0 = C; l.M = rows_out;
            l.W = b.W16(p + ".proj_out.weight"); l.ldw = C; l.N = C;
            l.bias = b.F32(p + ".proj_out.bias"); l.res = x3.p; l.ldr = C; l.out = out.p;
            b.linear(l);
        }
    }
    b.release(h);
    if (dedup) b.release(x3);
    if (st3 && st3 != st) b.release_raw(st3);
    if (st) b.release_raw(st);
    return out;
}

// shared_input: the sample groups [uncond | cond] of the evaluation are known to be identical (the stamp path packs them from
// the same latents / mask / masked-image latents); false for dtp_unet_forward with a caller-supplied sample tensor
int Engine::build_unet_plan(int B, int R, bool shared_input) {
    const bool want_dedup = shared_input && opt_dedup_branches_ != 0;
    if (unet_plan_.key_a == B && unet_plan_.key_b == R && unet_plan_dedup_ == want_dedup) return 0;
    if (R % 8) return fail("resolution must be a multiple of 8");
    if (ensure_arena()) return -1;
    const int Bz = 3 * B, h = R / 8;
    const size_t need = static_cast<size_t>(Bz) * h * h;
    if (need > unet_io_cap_) {
        unet_in_ = static_cast<__half*>(persistent(need * 64 * 2, true));
        unet_eps_ = static_cast<float*>(persistent(need * cfg_.unet_out_channels * 4, true));
        if (!unet_in_ || !unet_eps_) return -1;
        unet_io_cap_ = need;
    }
    if (kv_index_B_ != B) {
        std::vector<int> idx(Bz);
        for (int i = 0; i < Bz; ++i) idx[i] = i < B ? 0 : 1;
        kv_index_ = static_cast<int*>(persistent(Bz * sizeof(int), false));
        if (!kv_index_) return -1;
        cudaMemcpy(kv_index_, idx.data(), Bz * sizeof(int), cudaMemcpyHostToDevice);
        kv_index_B_ = B;
    }
    unet_plan_.clear();
    arena_.reset();
    Builder b(*this, unet_plan_, "unet.");
    // op 0: select the time-embedding bias row of the step being evaluated
    unet_plan_.add(K_OTHER, [this](cudaStream_t st) -> int {
        if (launch_copy_f32(temb_all_ + static_cast<size_t>(cur_step_) * temb_total_, temb_cur_, temb_total_, st)) {
            err_ = kernels_last_error();
            return -1;
        }
        return 1;
    });
    std::map<std::string, int> temb_off, tf_idx;
    for (auto& r : resnets_) temb_off[r.first] = r.second;
    for (size_t i = 0; i < tf_names_.size(); ++i) tf_idx[tf_names_[i]] = static_cast<int>(i);
    auto res = [&](const std::string& p, const Act& x, const Act& skip) {
        return b.resnet(p, x, skip, 1e-5f, temb_cur_ + temb_off[p]);
    };
    auto tf = [&](const std::string& p, const Act& x) {
        return transformer2d(b, *this, cfg_, p, x, cross_kv_[tf_idx[p]], kv_index_);
    };
    Act in;
    in.p = unet_in_; in.N = Bz; in.H = h; in.W = h; in.C = 64;
    Act x = b.conv3x3(in, Act{}, "conv_in", nullptr, nullptr);
    std::vector<Act> skips;
    skips.push_back(x);
    for (int i = 0; i < 4 && b.ok; ++i) {
        for (int j = 0; j < cfg_.unet_layers_per_block && b.ok; ++j) {
            const std::string rp = "down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j);
            // first resnet + first self-attention: once for the uncond and the cond branch (transformer2d, dedup)
            const bool dd = i == 0 && j == 0 && want_dedup && cfg_.unet_down_attn[0] && fold_ln() && fold_cross() &&
                            fuse_cross() && cfg_.unet_heads * 16 == 128 && (x.C % 64) == 0;
            Act xin = x;
            if (dd) {
                xin.p = x.p + static_cast<size_t>(B) * h * h * x.C;  // samples [cond | texture-guidance] of conv_in's output
                xin.N = 2 * B;
            }
            Act y = res(rp, xin, Act{});
            if (cfg_.unet_down_attn[i]) {
                const std::string ap = "down_blocks." + std::to_string(i) + ".attentions." + std::to_string(j);
                Act z = dd ? transformer2d(b, *this, cfg_, ap, y, cross_kv_[tf_idx[ap]], kv_index_, true) : tf(ap, y);
                b.release(y);
                y = z;
            }
            x = y;
            skips.push_back(x);
        }
        if (i != 3 && b.ok) {
            x = b.downsample(x, "down_blocks." + std::to_string(i) + ".downsamplers.0.conv", 1);
            skips.push_back(x);
        }
    }
    if (b.ok) {
        Act y = res("mid_block.resnets.0", x, Act{});
        Act z = tf("mid_block.attentions.0", y);
        b.release(y);
        x = res("mid_block.resnets.1", z, Act{});  // the last down activation stays alive as a skip
        b.release(z);
    }
    for (int i = 0; i < 4 && b.ok; ++i) {
        for (int j = 0; j < cfg_.unet_layers_per_block + 1 && b.ok; ++j) {
            Act skip = skips.back();
            skips.pop_back();
            Act y = res("up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), x, skip);
            b.release(x);
            b.release(skip);
            if (cfg_.unet_down_attn[3 - i]) {
                Act z = tf("up_blocks." + std::to_string(i) + ".attentions." + std::to_string(j), y);
                b.release(y);
                y = z;
            }
            x = y;
        }
        if (i != 3 && b.ok) {
            Act up = b.upsample_conv(x, "up_blocks." + std::to_string(i) + ".upsamplers.0.conv");
            b.release(x);
            x = up;
        }
    }
    if (b.ok) {
        Act n = b.groupnorm(x, Act{}, "conv_norm_out", 1e-5f, 1);
        b.release(x);
        b.conv3x3_into(n, Act{}, "conv_out.weight", b.F32("conv_out.bias"), nullptr, 0, unet_eps_,
                       cfg_.unet_out_channels, EPI_OUT_F32_NCHW, h * h);
        b.release(n);
    }
    if (!b.ok) {
        unet_plan_.clear();
        return -1;
    }
    unet_plan_.key_a = B;
    unet_plan_.key_b = R;
    unet_plan_dedup_ = want_dedup;
    return ensure_ws();
}

// ------------------------------------------------------------------------------------------------------------
// VAE plans
// ------------------------------------------------------------------------------------------------------------
static Act vae_attention(Builder& b, const std::string& p, const Act& x) {
    if (!b.ok) return Act{};
    const int C = x.C, seq = x.H * x.W;
    const long long rows = x.rows();
    Act n = b.groupnorm(x, Act{}, p + ".group_norm", 1e-6f, 0);
    Act qkv = b.like(x, 3 * C);
    {
        Lin l;
        l.A0 = n.p; l.lda0 = C; l.K0 = C; l.M = rows;
        l.W = b.W16(p + ".qkv.weight"); l.ldw = C; l.N = 3 * C; l.bias = b.F32(p + ".qkv.bias"); l.out = qkv.p;
        b.linear(l);
    }
    b.release(n);
    Act att = b.like(x, C);
    if (b.ok)
        b.attention(qkv.p, 3 * C, qkv.p + C, 3 * C, qkv.p + 2 * C, 3 * C, att.p, C, seq, seq, 1, C, x.N,
                    static_cast<long long>(seq) * 3 * C, static_cast<long long>(seq) * 3 * C,
                    static_cast<long long>(seq) * C, nullptr);
    b.release(qkv);
    Act out = b.like(x, C);
    {
        Lin l;
        l.A0 = att.p; l.lda0 = C; l.K0 = C; l.M = rows;
        l.W = b.W16(p + ".proj_attn.weight"); l.ldw = C; l.N = C; l.bias = b.F32(p + ".proj_attn.bias");
        l.res = x.p; l.ldr = C; l.out = out.p;
        b.linear(l);
    }
    b.release(att);
    return out;
}

int Engine::build_vae_enc_plan(int Nb, int R) {
    if (vae_enc_plan_.key_a == Nb && vae_enc_plan_.key_b == R) return 0;
    if (R % 8) return fail("resolution must be a multiple of 8");
    if (ensure_arena()) return -1;
    const int h = R / 8;
    const size_t need = static_cast<size_t>(Nb) * R * R;
    if (need > vae_enc_cap_) {
        vae_enc_in_ = static_cast<__half*>(persistent(need * 64 * 2, true));
        vae_moments_ = static_cast<float*>(persistent(static_cast<size_t>(Nb) * 2 * cfg_.vae_latent * h * h * 4, true));
        if (!vae_enc_in_ || !vae_moments_) return -1;
        vae_enc_cap_ = need;
    }
    vae_enc_plan_.clear();
    arena_.reset();
    Builder b(*this, vae_enc_plan_, "vae.");
    Act in;
    in.p = vae_enc_in_; in.N = Nb; in.H = R; in.W = R; in.C = 64;
    Act x = b.conv3x3(in, Act{}, "encoder.conv_in", nullptr, nullptr);
    for (int i = 0; i < 4 && b.ok; ++i) {
        for (int j = 0; j < cfg_.vae_layers_per_block && b.ok; ++j) {
            Act y = b.resnet("encoder.down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), x, Act{},
                             1e-6f, nullptr);
            b.release(x);
            x = y;
        }
        if (i != 3 && b.ok) {
            Act y = b.downsample(x, "encoder.down_blocks." + std::to_string(i) + ".downsamplers.0.conv", 0);
            b.release(x);
            x = y;
        }
    }
    if (b.ok) {
        Act y = b.resnet("encoder.mid_block.resnets.0", x, Act{}, 1e-6f, nullptr);
        b.release(x);
        Act z = vae_attention(b, "encoder.mid_block.attentions.0", y);
        b.release(y);
        x = b.resnet("encoder.mid_block.resnets.1", z, Act{}, 1e-6f, nullptr);
        b.release(z);
        Act n = b.groupnorm(x, Act{}, "encoder.conv_norm_out", 1e-6f, 1);
        b.release(x);
        Act mo = b.conv3x3(n, Act{}, "encoder.conv_out", nullptr, nullptr);  // (Nb, h, h, 2*latent) f16
        b.release(n);
        Lin l;
        const int L2 = 2 * cfg_.vae_latent;
        l.A0 = mo.p; l.lda0 = L2; l.K0 = L2; l.M = mo.rows();
        l.W = b.W16("quant_conv.weight"); l.ldw = L2; l.N = L2; l.bias = b.F32("quant_conv.bias");
        l.out = vae_moments_; l.flags = EPI_OUT_F32_NCHW; l.hw_out = h * h;
        b.linear(l);
        b.release(mo);
    }
    if (!b.ok) {
        vae_enc_plan_.clear();
        return -1;
    }
    vae_enc_plan_.key_a = Nb;
    vae_enc_plan_.key_b = R;
    return ensure_ws();
}

int Engine::build_vae_dec_plan(int B, int R) {
    if (vae_dec_plan_.key_a == B && vae_dec_plan_.key_b == R) return 0;
    if (R % 8) return fail("resolution must be a multiple of 8");
    if (ensure_arena()) return -1;
    const int h = R / 8;
    const size_t need = static_cast<size_t>(B) * R * R;
    if (need > vae_dec_cap_) {
        vae_dec_in_ = static_cast<__half*>(persistent(static_cast<size_t>(B) * h * h * 8 * 2, true));
        vae_dec_out_ = static_cast<float*>(persistent(need * 3 * 4, true));
        if (!vae_dec_in_ || !vae_dec_out_) return -1;
        vae_dec_cap_ = need;
    }
    vae_dec_plan_.clear();
    arena_.reset();
    Builder b(*this, vae_dec_plan_, "vae.");
    // post_quant_conv (1x1, latent -> latent) writes the first channels of a zero-padded 64-channel NHWC tensor
    __half* pq = static_cast<__half*>(persistent(static_cast<size_t>(B) * h * h * 64 * 2, true));
    if (!pq) return -1;
    {
        Lin l;
        l.A0 = vae_dec_in_; l.lda0 = 8; l.K0 = 8; l.M = static_cast<long long>(B) * h * h;
        l.W = b.W16("post_quant_conv.weight"); l.ldw = 8; l.N = cfg_.vae_latent; l.bias = b.F32("post_quant_conv.bias");
        l.out = pq; l.ldc = 64;
        b.linear(l);
    }
    Act in;
    in.p = pq; in.N = B; in.H = h; in.W = h; in.C = 64;
    Act x = b.conv3x3(in, Act{}, "decoder.conv_in", nullptr, nullptr);
    if (b.ok) {
        Act y = b.resnet("decoder.mid_block.resnets.0", x, Act{}, 1e-6f, nullptr);
        b.release(x);
        Act z = vae_attention(b, "decoder.mid_block.attentions.0", y);
        b.release(y);
        x = b.resnet("decoder.mid_block.resnets.1", z, Act{}, 1e-6f, nullptr);
        b.release(z);
    }
    for (int i = 0; i < 4 && b.ok; ++i) {
        for (int j = 0; j < cfg_.vae_layers_per_block + 1 && b.ok; ++j) {
            Act y = b.resnet("decoder.up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), x, Act{}, 1e-6f,
                             nullptr);
            b.release(x);
            x = y;
        }
        if (i != 3 && b.ok) {
            Act up = b.upsample_conv(x, "decoder.up_blocks." + std::to_string(i) + ".upsamplers.0.conv");
            b.release(x);
            x = up;
        }
    }
    if (b.ok) {
        Act n = b.groupnorm(x, Act{}, "decoder.conv_norm_out", 1e-6f, 1);
        b.release(x);
        b.conv3x3_into(n, Act{}, "decoder.conv_out.weight", b.F32("decoder.conv_out.bias"), nullptr, 0, vae_dec_out_, 3,
                       EPI_OUT_F32_NCHW | EPI_IMG01, R * R);
        b.release(n);
    }
    if (!b.ok) {
        vae_dec_plan_.clear();
        return -1;
    }
    vae_dec_plan_.key_a = B;
    vae_dec_plan_.key_b = R;
    return ensure_ws();
}

// ------------------------------------------------------------------------------------------------------------
// image encoder plan (CLIP ViT-B/32 visual tower + three patch towers), fixed shapes
// ------------------------------------------------------------------------------------------------------------
int Engine::build_encoder_plan() {
    if (enc_plan_.key_a == 1) return 0;
    if (ensure_arena()) return -1;
    const int T = cfg_.enc_tokens, w = cfg_.enc_width;
    if (!enc_in_copy_) {
        enc_in_copy_ = static_cast<float*>(persistent(static_cast<size_t>(T) * 3 * 224 * 224 * 4, true));
        enc_out_ = static_cast<float*>(persistent(static_cast<size_t>(T) * cfg_.enc_cross_dim * 4, true));
        if (!enc_in_copy_ || !enc_out_) return -1;
    }
    enc_plan_.clear();
    arena_.reset();
    Builder b(*this, enc_plan_, "enc.");
    Engine* eng = this;
    const std::string v = "clip.visual";
    Act pat = b.alloc(T, 49, 1, 3072);
    if (b.ok) {
        const float* src = enc_in_copy_;
        __half* dst = pat.p;
        enc_plan_.add(K_OTHER, [=](cudaStream_t st) -> int {
            if (launch_patchify32(src, T, dst, st)) {
                eng->err_ = kernels_last_error();
                return -1;
            }
            return 1;
        });
    }
    Act tok = b.alloc(T, 49, 1, w);
    {
        Lin l;
        l.A0 = pat.p; l.lda0 = 3072; l.K0 = 3072; l.M = static_cast<long long>(T) * 49;
        l.W = b.W16(v + ".conv1.weight"); l.ldw = 3072; l.N = w; l.out = tok.p;
        b.linear(l);
    }
    b.release(pat);
    Act x = b.alloc(T, 50, 1, w);
    if (b.ok) {
        const float* cls = b.F32(v + ".class_embedding");
        const float* pos = b.F32(v + ".positional_embedding");
        const __half* tp = tok.p;
        __half* xp = x.p;
        enc_plan_.add(K_OTHER, [=](cudaStream_t st) -> int {
            if (launch_clip_embed(tp, cls, pos, T, w, xp, st)) {
                eng->err_ = kernels_last_error();
                return -1;
            }
            return 1;
        });
    }
    b.release(tok);
    Act tmp = b.like(x, w);
    b.layernorm(x, v + ".ln_pre", tmp.p);
    b.release(x);
    x = tmp;
    const long long rows = x.rows();
    const int heads = cfg_.enc_heads, d = w / heads;
    for (int i = 0; i < cfg_.enc_layers && b.ok; ++i) {
        const std::string r = v + ".transformer.resblocks." + std::to_string(i);
        Act t = b.like(x, w);
        b.layernorm(x, r + ".ln_1", t.p);
        Act qkv = b.like(x, 3 * w);
        {
            Lin l;
            l.A0 = t.p; l.lda0 = w; l.K0 = w; l.M = rows;
            l.W = b.W16(r + ".attn.in_proj_weight"); l.ldw = w; l.N = 3 * w; l.bias = b.F32(r + ".attn.in_proj_bias");
            l.out = qkv.p;
            b.linear(l);
        }
        if (b.ok)
            b.attention(qkv.p, 3 * w, qkv.p + w, 3 * w, qkv.p + 2 * w, 3 * w, t.p, w, 50, 50, heads, d, T, 50LL * 3 * w,
                        50LL * 3 * w, 50LL * w, nullptr);
        b.release(qkv);
        {
            Lin l;
            l.A0 = t.p; l.lda0 = w; l.K0 = w; l.M = rows;
            l.W = b.W16(r + ".attn.out_proj.weight"); l.ldw = w; l.N = w; l.bias = b.F32(r + ".attn.out_proj.bias");
            l.res = x.p; l.ldr = w; l.out = x.p;
            b.linear(l);
        }
        b.layernorm(x, r + ".ln_2", t.p);
        Act m = b.like(x, cfg_.enc_mlp);
        {
            Lin l;
            l.A0 = t.p; l.lda0 = w; l.K0 = w; l.M = rows;
            l.W = b.W16(r + ".mlp.c_fc.weight"); l.ldw = w; l.N = cfg_.enc_mlp; l.bias = b.F32(r + ".mlp.c_fc.bias");
            l.out = m.p; l.flags = EPI_QUICKGELU;
            b.linear(l);
        }
        b.release(t);
        {
            Lin l;
            l.A0 = m.p; l.lda0 = cfg_.enc_mlp; l.K0 = cfg_.enc_mlp; l.M = rows;
            l.W = b.W16(r + ".mlp.c_proj.weight"); l.ldw = cfg_.enc_mlp; l.N = w; l.bias = b.F32(r + ".mlp.c_proj.bias");
            l.res = x.p; l.ldr = w; l.out = x.p;
            b.linear(l);
        }
        b.release(m);
    }
    // ln_post on the class tokens (row n*50), then + pos_emb (the reference's view-scrambled table, uploaded as enc.pos_emb)
    Act cls = b.alloc(T, 1, 1, w);
    Act lat = b.alloc(T, 1, 1, w);
    if (b.ok) {
        std::vector<int> idx(T);
        for (int i = 0; i < T; ++i) idx[i] = i * 50;
        int* didx = static_cast<int*>(persistent(T * sizeof(int), false));
        if (!didx) return -1;
        cudaMemcpy(didx, idx.data(), T * sizeof(int), cudaMemcpyHostToDevice);
        const __half* xp = x.p;
        __half* cp = cls.p;
        enc_plan_.add(K_OTHER, [=](cudaStream_t st) -> int {
            if (launch_gather_rows(xp, w, didx, T, w, cp, st)) {
                eng->err_ = kernels_last_error();
                return -1;
            }
            return 1;
        });
        b.layernorm(cls, v + ".ln_post", lat.p);
        const float* pe = b.F32("pos_emb");
        __half* lp = lat.p;
        if (b.ok)
            enc_plan_.add(K_OTHER, [=](cudaStream_t st) -> int {
                if (launch_add_rows_bcast(lp, pe, T, w, T, st)) {
                    eng->err_ = kernels_last_error();
                    return -1;
                }
                return 1;
            });
    }
    b.release(x);
    b.release(cls);
    // three towers over token ranges [0,1), [1,5), [5,14) (image_encoder.py:83-92)
    const char* tower_names[3] = {"l", "m", "s"};
    const int th = cfg_.enc_tower_heads, td = w / th;
    int row0 = 0;
    for (int tw = 0; tw < 3 && b.ok; ++tw) {
        const int n = (tw == 0) ? 1 : (tw == 1 ? 4 : 9);
        Act xs;
        xs.p = lat.p + static_cast<long long>(row0) * w; xs.N = 1; xs.H = n; xs.W = 1; xs.C = w;
        for (int i = 0; i < cfg_.enc_tower_layers && b.ok; ++i) {
            const std::string r = std::string(tower_names[tw]) + "_patch_encoder_layers." + std::to_string(i);
            Act t = b.like(xs, w);
            b.layernorm(xs, r + ".norm1", t.p);
            Act qkv = b.like(xs, 3 * w);
            {
                Lin l;
                l.A0 = t.p; l.lda0 = w; l.K0 = w; l.M = n;
                l.W = b.W16(r + ".attn1.to_qkv.weight"); l.ldw = w; l.N = 3 * w; l.bias = b.F32(r + ".attn1.to_qkv.bias");
                l.out = qkv.p;
                b.linear(l);
            }
            if (b.ok)
                b.attention(qkv.p, 3 * w, qkv.p + w, 3 * w, qkv.p + 2 * w, 3 * w, t.p, w, n, n, th, td, 1, 0, 0, 0,
                            nullptr);
            b.release(qkv);
            {
                Lin l;
                l.A0 = t.p; l.lda0 = w; l.K0 = w; l.M = n;
                l.W = b.W16(r + ".attn1.to_out.0.weight"); l.ldw = w; l.N = w; l.bias = b.F32(r + ".attn1.to_out.0.bias");
                l.res = xs.p; l.ldr = w; l.out = xs.p;
                b.linear(l);
            }
            b.layernorm(xs, r + ".norm3", t.p);
            Act m = b.like(xs, 4 * w);
            {
                Lin l;
                l.A0 = t.p; l.lda0 = w; l.K0 = w; l.M = n;
                l.W = b.W16(r + ".ff.net.0.proj.weight"); l.ldw = w; l.N = 4 * w; l.bias = b.F32(r + ".ff.net.0.proj.bias");
                l.out = m.p; l.flags = EPI_GELU;
                b.linear(l);
            }
            b.release(t);
            {
                Lin l;
                l.A0 = m.p; l.lda0 = 4 * w; l.K0 = 4 * w; l.M = n;
                l.W = b.W16(r + ".ff.net.2.weight"); l.ldw = 4 * w; l.N = w; l.bias = b.F32(r + ".ff.net.2.bias");
                l.res = xs.p; l.ldr = w; l.out = xs.p;
                b.linear(l);
            }
            b.release(m);
        }
        row0 += n;
    }
    if (b.ok) {
        Act t = b.like(lat, w);
        b.layernorm(lat, "final_layer_norm", t.p);
        Lin l;
        l.A0 = t.p; l.lda0 = w; l.K0 = w; l.M = T;
        l.W = b.W16("proj_out.weight"); l.ldw = w; l.N = cfg_.enc_cross_dim; l.bias = b.F32("proj_out.bias");
        l.out = enc_out_; l.flags = EPI_OUT_F32;
        b.linear(l);
        b.release(t);
    }
    b.release(lat);
    if (!b.ok) {
        enc_plan_.clear();
        return -1;
    }
    enc_plan_.key_a = 1;
    return ensure_ws();
}

int Engine::encode_patches(const float* patches, float* emb_out, cudaStream_t st) {
    if (!finalized_ && finalize_weights()) return -1;
    if (build_encoder_plan()) return -1;
    const size_t n = static_cast<size_t>(cfg_.enc_tokens) * 3 * 224 * 224;
    if (cudaMemcpyAsync(enc_in_copy_, patches, n * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return fail("encode_patches: input copy failed");
    if (enc_plan_.run(st, &launches_, &prof_)) return -1;
    if (cudaMemcpyAsync(emb_out, enc_out_, static_cast<size_t>(cfg_.enc_tokens) * cfg_.enc_cross_dim * 4,
                        cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return fail("encode_patches: output copy failed");
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// condition: fp16 cast of [uncond | cond] and the cross-attention K/V of every transformer layer
// ------------------------------------------------------------------------------------------------------------
int Engine::build_cond_plan() {
    if (cond_plan_.key_a == 1) return 0;
    cond_plan_.clear();
    Builder b(*this, cond_plan_, "unet.");
    const int T = cfg_.enc_tokens, D = cfg_.unet_cross_dim;
    Engine* eng = this;
    cond_plan_.add(K_OTHER, [=](cudaStream_t st) -> int {
        if (launch_f32_to_f16(eng->ctx_f32_, eng->ctx_, static_cast<long long>(2) * T * D, st)) {
            eng->err_ = kernels_last_error();
            return -1;
        }
        return 1;
    });
    for (size_t i = 0; i < tf_names_.size() && b.ok; ++i) {
        int r = 0, c = 0;
        Lin l;
        l.A0 = ctx_; l.lda0 = D; l.K0 = D; l.M = 2 * T;
        l.W = b.W16(tf_names_[i] + ".transformer_blocks.0.attn2.to_kv.weight", &r, &c);
        l.ldw = D; l.N = r; l.out = cross_kv_[i];
        if (b.ok && c != D) b.fail("attn2.to_kv weight has unexpected K");
        b.linear(l);
    }
    // folded cross-attention operands per layer and context slot (0 = uncond, 1 = 2 = cond)
    const int heads = cfg_.unet_heads, HP = heads * 16;
    for (size_t i = 0; i < tf_names_.size() && b.ok; ++i) {
        const std::string t = tf_names_[i] + ".transformer_blocks.0";
        int C = 0, kq = 0;
        const __half* Wq = b.W16(t + ".attn2.to_q.weight", &C, &kq);
        const __half* Wo = b.W16(t + ".attn2.to_out.0.weight");
        if (!b.ok) break;
        const int d = C / heads;
        const float scale = 1.0f / sqrtf(static_cast<float>(d));
        for (int slot = 0; slot < 3 && b.ok; ++slot) {
            const int c = slot == 0 ? 0 : 1;
            const __half* Kc = cross_kv_[i] + static_cast<size_t>(c) * T * 2 * C;
            const __half* Vc = Kc + C;
            {   // Wscore[slot][h*16 + j][:] = scale * K_c,h[j,:] Wq_h      (B = Wq_h consumed MN-major)
                Builder::Bmm g;
                g.A = Kc; g.lda = 2 * C; g.a_zs1 = d;
                g.B = Wq; g.ldb = C; g.b_zs1 = static_cast<long long>(d) * C; g.b_mn = 1;
                g.M = T; g.N = C; g.K = d; g.nz1 = heads;
                g.out = wscore_[i] + static_cast<size_t>(slot) * HP * C; g.ldc = C; g.out_zs1 = 16LL * C;
                g.alpha = scale; g.label = "fold_scores";
                b.bmm(g);
            }
            if (!ln_.empty()) {  // same operand through Wq diag(norm2.weight): the LayerNorm-folded score contraction
                Builder::Bmm g;
                g.A = Kc; g.lda = 2 * C; g.a_zs1 = d;
                g.B = ln_[i].q_w; g.ldb = C; g.b_zs1 = static_cast<long long>(d) * C; g.b_mn = 1;
                g.M = T; g.N = C; g.K = d; g.nz1 = heads;
                g.out = ln_[i].wscore + static_cast<size_t>(slot) * HP * C; g.ldc = C; g.out_zs1 = 16LL * C;
                g.alpha = scale; g.label = "fold_scores_ln";
                b.bmm(g);
            }
            {   // Wout[slot][:, h*16 + j] = Wo[:, h*d:(h+1)*d] V_c,h[j,:]^T
                Builder::Bmm g;
                g.A = Wo; g.lda = C; g.a_zs1 = d;
                g.B = Vc; g.ldb = 2 * C; g.b_zs1 = d;
                g.M = C; g.N = T; g.K = d; g.nz1 = heads;
                g.out = wout_[i] + static_cast<size_t>(slot) * C * HP; g.ldc = HP; g.out_zs1 = 16;
                g.label = "fold_out";
                b.bmm(g);
            }
        }
    }
    for (size_t i = 0; i < tf_names_.size() && b.ok && !ln_.empty(); ++i) {
        int C = 0;
        b.W16(tf_names_[i] + ".transformer_blocks.0.attn2.to_q.weight", &C, nullptr);
        if (!b.ok) break;
        const LnFold f = ln_[i];
        const __half* kvp = cross_kv_[i];
        const float scale = 1.0f / sqrtf(static_cast<float>(C / heads));
        cond_plan_.add(K_OTHER, [=](cudaStream_t st) -> int {
            if (launch_cross_ln_finish(f.wscore, kvp, f.q_beta, C, heads, T, scale, f.ws_cs, f.ws_b, st)) {
                eng->err_ = kernels_last_error();
                return -1;
            }
            return 1;
        });
    }
    if (!b.ok) {
        cond_plan_.clear();
        return -1;
    }
    cond_plan_.key_a = 1;
    return ensure_ws();
}

int Engine::set_condition(const float* emb, const float* uncond, cudaStream_t st) {
    if (!finalized_ && finalize_weights()) return -1;
    if (build_cond_plan()) return -1;
    const size_t n = static_cast<size_t>(cfg_.enc_tokens) * cfg_.unet_cross_dim;
    if (cudaMemcpyAsync(ctx_f32_, uncond, n * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(ctx_f32_ + n, emb, n * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return fail("set_condition: copy failed");
    if (cond_plan_.run(st, &launches_, &prof_)) return -1;
    cond_set_ = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// schedule: per-evaluation time-embedding bias rows
// ------------------------------------------------------------------------------------------------------------
int Engine::set_schedule(int n, const float* ts, const float* a_t, const float* a_prev, float cfg, float tg,
                         int tg_steps) {
    if (n < 0 || n > 1000) return fail("set_schedule: bad evaluation count");
    const bool same = (n == n_steps_) && std::equal(ts, ts + n, ts_.begin());
    ts_.assign(ts, ts + n);
    a_t_.assign(a_t, a_t + n);
    a_prev_.assign(a_prev, a_prev + n);
    n_steps_ = n;
    cfg_w_ = cfg;
    tg_w_ = tg;
    tg_steps_ = tg_steps;
    if (!same) temb_dirty_ = true;
    return 0;
}

int Engine::build_temb_tables(cudaStream_t st) {
    if (!temb_dirty_) return 0;
    if (!finalized_ && finalize_weights()) return -1;
    const int n = n_steps_;
    if (n == 0) {
        temb_dirty_ = false;
        return 0;
    }
    if (n > temb_cap_steps_) {
        temb_all_ = static_cast<float*>(persistent(static_cast<size_t>(n) * temb_total_ * 4, true));
        if (!temb_all_) return -1;
        temb_cap_steps_ = n;
    }
    const int c0 = cfg_.unet_block_out[0], Tdim = 4 * c0;
    float* d_ts = nullptr;
    __half *e16 = nullptr, *t1 = nullptr, *t2 = nullptr;
    if (cudaMalloc(&d_ts, n * 4) != cudaSuccess || cudaMalloc(&e16, static_cast<size_t>(n) * c0 * 2) != cudaSuccess ||
        cudaMalloc(&t1, static_cast<size_t>(n) * Tdim * 2) != cudaSuccess ||
        cudaMalloc(&t2, static_cast<size_t>(n) * Tdim * 2) != cudaSuccess)
        return fail("schedule tables: cudaMalloc failed");
    cudaMemcpyAsync(d_ts, ts_.data(), n * 4, cudaMemcpyHostToDevice, st);
    Plan p;
    Builder b(*this, p, "unet.");
    Engine* eng = this;
    p.add(K_OTHER, [=](cudaStream_t s) -> int {
        if (launch_timestep_embedding(d_ts, n, c0, e16, s)) {
            eng->err_ = kernels_last_error();
            return -1;
        }
        return 1;
    });
    {
        Lin l;
        l.A0 = e16; l.lda0 = c0; l.K0 = c0; l.M = n;
        l.W = b.W16("time_embedding.linear_1.weight"); l.ldw = c0; l.N = Tdim;
        l.bias = b.F32("time_embedding.linear_1.bias"); l.out = t1; l.flags = EPI_SILU;
        b.linear(l);
    }
    {
        // every consumer applies SiLU to temb before its projection (ResnetBlock2D), so fold it here
        Lin l;
        l.A0 = t1; l.lda0 = Tdim; l.K0 = Tdim; l.M = n;
        l.W = b.W16("time_embedding.linear_2.weight"); l.ldw = Tdim; l.N = Tdim;
        l.bias = b.F32("time_embedding.linear_2.bias"); l.out = t2; l.flags = EPI_SILU;
        b.linear(l);
    }
    const float* comb = b.F32("__temb_bias_comb");
    for (auto& r : resnets_) {
        if (!b.ok) break;
        int rr = 0, cc = 0;
        Lin l;
        l.A0 = t2; l.lda0 = Tdim; l.K0 = Tdim; l.M = n;
        l.W = b.W16(r.first + ".time_emb_proj.weight", &rr, &cc); l.ldw = Tdim; l.N = rr;
        l.bias = comb + r.second; l.out = temb_all_ + r.second; l.ldc = temb_total_; l.flags = EPI_OUT_F32;
        b.linear(l);
    }
    int rc = b.ok ? 0 : -1;
    if (rc == 0) rc = ensure_ws();
    if (rc == 0) rc = p.run(st, &launches_, &prof_);
    cudaStreamSynchronize(st);
    cudaFree(d_ts);
    cudaFree(e16);
    cudaFree(t1);
    cudaFree(t2);
    if (rc == 0) temb_dirty_ = false;
    return rc;
}

// ------------------------------------------------------------------------------------------------------------
// stage entry points and the stamp driver
// ------------------------------------------------------------------------------------------------------------
#define KCHECK(call)                       \
    do {                                   \
        if ((call) != 0) {                 \
            err_ = kernels_last_error();   \
            return -1;                     \
        }                                  \
        ++launches_;                       \
    } while (0)

int Engine::vae_encode(int Nb, int R, const float* images, const float* noise, float* latents_out, cudaStream_t st) {
    if (!finalized_ && finalize_weights()) return -1;
    if (build_vae_enc_plan(Nb, R)) return -1;
    const int h = R / 8;
    KCHECK(launch_nchw_to_nhwc_pad(images, Nb, 3, R * R, 64, 1.0f, vae_enc_in_, st));
    if (vae_enc_plan_.run(st, &launches_, &prof_)) return -1;
    KCHECK(launch_vae_sample(vae_moments_, noise, Nb, h * h, 0.18215f, latents_out, st));
    return 0;
}

int Engine::vae_decode(int B, int R, const float* latents, float* images_out, cudaStream_t st) {
    if (!finalized_ && finalize_weights()) return -1;
    if (build_vae_dec_plan(B, R)) return -1;
    const int h = R / 8;
    KCHECK(launch_nchw_to_nhwc_pad(latents, B, cfg_.vae_latent, h * h, 8, 0.18215f, vae_dec_in_, st));
    if (vae_dec_plan_.run(st, &launches_, &prof_)) return -1;
    if (images_out != vae_dec_out_)
        KCHECK(launch_copy_f32(vae_dec_out_, images_out, static_cast<long long>(B) * 3 * R * R, st));
    return 0;
}

int Engine::unet_forward(int B, int R, const float* sample, const float* latents, const float* mask3,
                         const float* masked3, int step, float* eps_out, cudaStream_t st) {
    if (!finalized_ && finalize_weights()) return -1;
    if (!cond_set_) return fail("unet_forward: call dtp_set_condition first");
    if (step < 0 || step >= n_steps_) return fail("unet_forward: step outside the schedule");
    if (build_temb_tables(st)) return -1;
    if (build_unet_plan(B, R, sample == nullptr)) return -1;
    const int h = R / 8, Bz = 3 * B;
    if (sample) {
        KCHECK(launch_nchw_to_nhwc_pad(sample, Bz, cfg_.unet_in_channels, h * h, 64, 1.0f, unet_in_, st));
    } else {
        KCHECK(launch_pack_unet_input(latents, mask3, masked3, B, h * h, unet_in_, st));
    }
    cur_step_ = step;
    if (unet_plan_.run(st, &launches_, &prof_)) return -1;
    if (eps_out && eps_out != unet_eps_)
        KCHECK(launch_copy_f32(unet_eps_, eps_out, static_cast<long long>(Bz) * cfg_.unet_out_channels * h * h, st));
    return 0;
}

int Engine::ensure_io(int B, int R) {
    if (static_cast<size_t>(B) <= io_cap_B_ && static_cast<size_t>(R) <= io_cap_R_) return 0;
    const size_t b = std::max<size_t>(B, io_cap_B_), r = std::max<size_t>(R, io_cap_R_);
    const size_t hw = (r / 8) * (r / 8);
    lat_ = static_cast<float*>(persistent(b * 4 * hw * 4, true));
    mask3_ = static_cast<float*>(persistent(3 * b * hw * 4, true));
    masked3_ = static_cast<float*>(persistent(3 * b * 4 * hw * 4, true));
    lat2_ = static_cast<float*>(persistent(2 * b * 4 * hw * 4, true));
    // canvas pre-process outputs: masked (3), mask (1), ctx (3), ctx mask (1), scratch (1), raw result (3), images x2 (6)
    pre_ = static_cast<float*>(persistent(b * 18 * r * r * 4, true));
    // staging: canvas (4), brush (3, one image), init latents, vae noise, out f32 (3); out u8 (3 bytes / pixel)
    stage_ = static_cast<float*>(persistent((b * 7 * r * r + 3 * r * r + b * 4 * hw + 2 * b * 4 * hw) * 4, true));
    stage_u8_ = static_cast<unsigned char*>(persistent(b * 3 * r * r, true));
    if (!lat_ || !mask3_ || !masked3_ || !lat2_ || !pre_ || !stage_ || !stage_u8_) return -1;
    g_infer_.key.clear();
    g_stamp_.key.clear();
    io_cap_B_ = b;
    io_cap_R_ = r;
    return 0;
}

std::string Engine::schedule_key() const {
    char buf[160];
    double h = 0.0;
    for (int i = 0; i < n_steps_; ++i) h = h * 1.000001 + ts_[i] * (i + 1) + a_t_[i] * 7.0 + a_prev_[i] * 13.0;
    snprintf(buf, sizeof(buf), "n%d c%.9g t%.9g s%d h%.17g", n_steps_, cfg_w_, tg_w_, tg_steps_, h);
    return buf;
}

int Engine::run_graphed(GraphSlot& slot, const std::string& key, const std::function<int(cudaStream_t)>& body,
                        cudaStream_t st) {
    if (!opt_graph_ || prof_.on || opt_stage_timers_) return body(st);
    if (slot.key != key) {
        if (slot.exec) cudaGraphExecDestroy(slot.exec);
        slot.exec = nullptr;
        slot.key = key;
        slot.warm = 0;
    }
    if (slot.warm < 1) {  // first call with this key runs eagerly: builds plans, tables, kernel attributes
        ++slot.warm;
        return body(st);
    }
    // capture and replay on an engine-owned stream (the caller's may be the legacy default stream, which cannot be captured);
    // events order it after the caller's staging copies and the caller's stream after the replay
    if (!gstream_) {
        if (cudaStreamCreateWithFlags(&gstream_, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&gev_in_, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&gev_out_, cudaEventDisableTiming) != cudaSuccess)
            return fail("graph stream / event creation failed");
    }
    if (!slot.exec) {
        const long long before = launches_;
        cudaStreamSynchronize(st);
        if (cudaStreamBeginCapture(gstream_, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
            cudaGetLastError();
            return body(st);
        }
        const int rc = body(gstream_);
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(gstream_, &g);
        if (rc != 0 || e != cudaSuccess || g == nullptr) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            if (rc != 0) return rc;
            return fail(std::string("graph capture failed: ") + cudaGetErrorString(e));
        }
        slot.launches = launches_ - before;
        launches_ = before;
        if (cudaGraphInstantiate(&slot.exec, g, 0) != cudaSuccess) {
            cudaGraphDestroy(g);
            slot.exec = nullptr;
            return fail("cudaGraphInstantiate failed");
        }
        cudaGraphDestroy(g);
    }
    if (cudaEventRecord(gev_in_, st) != cudaSuccess || cudaStreamWaitEvent(gstream_, gev_in_, 0) != cudaSuccess)
        return fail("graph stream ordering failed");
    if (cudaGraphLaunch(slot.exec, gstream_) != cudaSuccess) return fail("cudaGraphLaunch failed");
    if (cudaEventRecord(gev_out_, gstream_) != cudaSuccess || cudaStreamWaitEvent(st, gev_out_, 0) != cudaSuccess)
        return fail("graph stream ordering failed");
    launches_ += slot.launches;
    ++graph_launches_;
    return 0;
}

int Engine::infer(int B, int R, const float* masked_img, const float* mask, const float* ctx_img, const float* ctx_mask,
                  const float* init_latents, const float* vae_noise, float* out_images, cudaStream_t st) {
    if (B < 1 || R < 8 || (R % 8)) return fail("infer: bad batch / resolution");
    if (!finalized_ && finalize_weights()) return -1;
    if (!cond_set_) return fail("infer: call dtp_set_condition first");
    if (ensure_io(B, R)) return -1;
    if (!opt_graph_ || prof_.on || opt_stage_timers_)
        return infer_body(B, R, masked_img, mask, ctx_img, ctx_mask, init_latents, vae_noise, out_images, st);
    // stage the caller's tensors at fixed addresses so the captured graph can be replayed
    const size_t plane = static_cast<size_t>(B) * R * R, hw = static_cast<size_t>(R / 8) * (R / 8);
    float* s_masked = pre_;
    float* s_mask = pre_ + 3 * plane;
    float* s_ctx = pre_ + 4 * plane;
    float* s_cmask = pre_ + 7 * plane;
    float* s_out = pre_ + 9 * plane;
    float* s_lat = stage_ + 7 * plane + 3 * static_cast<size_t>(R) * R;
    float* s_noise = s_lat + static_cast<size_t>(B) * 4 * hw;
    cudaMemcpyAsync(s_masked, masked_img, 3 * plane * 4, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(s_mask, mask, plane * 4, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(s_ctx, ctx_img, 3 * plane * 4, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(s_cmask, ctx_mask, plane * 4, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(s_lat, init_latents, static_cast<size_t>(B) * 4 * hw * 4, cudaMemcpyDeviceToDevice, st);
    if (vae_noise) cudaMemcpyAsync(s_noise, vae_noise, static_cast<size_t>(2) * B * 4 * hw * 4, cudaMemcpyDeviceToDevice, st);
    const std::string key = "i B" + std::to_string(B) + " R" + std::to_string(R) + (vae_noise ? " n1 " : " n0 ") +
                            schedule_key();
    const float* nz = vae_noise ? s_noise : nullptr;
    if (run_graphed(g_infer_, key,
                    [=](cudaStream_t s) { return infer_body(B, R, s_masked, s_mask, s_ctx, s_cmask, s_lat, nz, s_out, s); },
                    st))
        return -1;
    if (cudaMemcpyAsync(out_images, s_out, 3 * plane * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return fail("infer: output copy failed");
    return 0;
}

int Engine::infer_body(int B, int R, const float* masked_img, const float* mask, const float* ctx_img,
                       const float* ctx_mask, const float* init_latents, const float* vae_noise, float* out_images,
                       cudaStream_t st) {
    if (B < 1 || R < 8 || (R % 8)) return fail("infer: bad batch / resolution");
    if (!finalized_ && finalize_weights()) return -1;
    if (!cond_set_) return fail("infer: call dtp_set_condition first");
    if (ensure_io(B, R)) return -1;
    if (build_temb_tables(st)) return -1;
    const int h = R / 8, hw = h * h;
    const size_t img = static_cast<size_t>(B) * 3 * R * R;
    // both VAE encodes (masked image, context image) as one batch of 2B (inpaint_pipeline.py:125-126)
    float* both = pre_ + static_cast<size_t>(B) * 12 * R * R;
    stage_begin(ST_VAE_ENC, st);
    KCHECK(launch_copy_f32(masked_img, both, img, st));
    KCHECK(launch_copy_f32(ctx_img, both + img, img, st));
    if (vae_encode(2 * B, R, both, vae_noise, lat2_, st)) return -1;
    stage_end(ST_VAE_ENC, st);
    // masks: nearest /8, stacked [mask, mask, ctx_mask]; masked latents [ml, ml, cml] (inpaint_pipeline.py:114-116,136)
    KCHECK(launch_mask_nearest(mask, B, R, 8, mask3_, st));
    KCHECK(launch_copy_f32(mask3_, mask3_ + static_cast<size_t>(B) * hw, static_cast<long long>(B) * hw, st));
    KCHECK(launch_mask_nearest(ctx_mask, B, R, 8, mask3_ + static_cast<size_t>(2) * B * hw, st));
    const long long l4 = static_cast<long long>(B) * 4 * hw;
    KCHECK(launch_copy_f32(lat2_, masked3_, l4, st));
    KCHECK(launch_copy_f32(lat2_, masked3_ + l4, l4, st));
    KCHECK(launch_copy_f32(lat2_ + l4, masked3_ + 2 * l4, l4, st));
    KCHECK(launch_copy_f32(init_latents, lat_, l4, st));
    // denoising loop (stable_diffusion_pipeline.py:407-462)
    for (int i = 0; i < n_steps_; ++i) {
        const float tg = (i > tg_steps_ - 1) ? 0.0f : tg_w_;
        stage_begin(ST_UNET, st);
        if (unet_forward(B, R, nullptr, lat_, mask3_, masked3_, i, nullptr, st)) return -1;
        stage_end(ST_UNET, st);
        stage_begin(ST_DDIM, st);
        KCHECK(launch_guidance_ddim(unet_eps_, lat_, lat_, B, 4 * hw, cfg_w_, tg, a_t_[i], a_prev_[i], st));
        stage_end(ST_DDIM, st);
    }
    stage_begin(ST_VAE_DEC, st);
    if (vae_decode(B, R, lat_, out_images, st)) return -1;
    stage_end(ST_VAE_DEC, st);
    ++stamps_;
    return 0;
}

int Engine::stamp(int B, int R, const float* canvas, const float* brush, int pad, const float* init_latents,
                  const float* vae_noise, int composite, float* out_f32, unsigned char* out_u8, cudaStream_t st) {
    if (B < 1 || R < 8 || (R % 8)) return fail("stamp: bad batch / resolution");
    if (!finalized_ && finalize_weights()) return -1;
    if (!cond_set_) return fail("stamp: call dtp_set_condition first");
    if (ensure_io(B, R)) return -1;
    if (!opt_graph_ || prof_.on || opt_stage_timers_)
        return stamp_body(B, R, canvas, brush, pad, init_latents, vae_noise, composite, out_f32, out_u8, st);
    const size_t plane = static_cast<size_t>(B) * R * R, hw = static_cast<size_t>(R / 8) * (R / 8);
    float* s_canvas = stage_;
    float* s_out = stage_ + 4 * plane;
    float* s_brush = stage_ + 7 * plane;
    float* s_lat = s_brush + 3 * static_cast<size_t>(R) * R;
    float* s_noise = s_lat + static_cast<size_t>(B) * 4 * hw;
    cudaMemcpyAsync(s_canvas, canvas, 4 * plane * 4, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(s_brush, brush, static_cast<size_t>(3) * R * R * 4, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(s_lat, init_latents, static_cast<size_t>(B) * 4 * hw * 4, cudaMemcpyDeviceToDevice, st);
    if (vae_noise) cudaMemcpyAsync(s_noise, vae_noise, static_cast<size_t>(2) * B * 4 * hw * 4, cudaMemcpyDeviceToDevice, st);
    const std::string key = "s B" + std::to_string(B) + " R" + std::to_string(R) + " p" + std::to_string(pad) +
                            (vae_noise ? " n1" : " n0") + (composite ? " c1" : " c0") + (out_f32 ? " f1" : " f0") +
                            (out_u8 ? " u1 " : " u0 ") + schedule_key();
    const float* nz = vae_noise ? s_noise : nullptr;
    float* of = out_f32 ? s_out : nullptr;
    unsigned char* ou = out_u8 ? stage_u8_ : nullptr;
    if (run_graphed(g_stamp_, key,
                    [=](cudaStream_t s) { return stamp_body(B, R, s_canvas, s_brush, pad, s_lat, nz, composite, of, ou, s); },
                    st))
        return -1;
    if (out_f32 && cudaMemcpyAsync(out_f32, s_out, 3 * plane * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return fail("stamp: output copy failed");
    if (out_u8 && cudaMemcpyAsync(out_u8, stage_u8_, 3 * plane, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return fail("stamp: output copy failed");
    return 0;
}

int Engine::stamp_body(int B, int R, const float* canvas, const float* brush, int pad, const float* init_latents,
                       const float* vae_noise, int composite, float* out_f32, unsigned char* out_u8, cudaStream_t st) {
    if (B < 1 || R < 8 || (R % 8)) return fail("stamp: bad batch / resolution");
    if (ensure_io(B, R)) return -1;
    const size_t plane = static_cast<size_t>(B) * R * R;
    float* masked = pre_;
    float* mask = pre_ + 3 * plane;
    float* ctx = pre_ + 4 * plane;
    float* cmask = pre_ + 7 * plane;
    float* scratch = pre_ + 8 * plane;
    float* raw = pre_ + 9 * plane;
    stage_begin(ST_PRE, st);
    if (launch_canvas_preprocess(canvas, brush, B, R, pad, masked, mask, ctx, cmask, scratch, st)) {
        err_ = kernels_last_error();
        return -1;
    }
    launches_ += 2;
    stage_end(ST_PRE, st);
    float* dst = (composite || out_f32 == nullptr) ? raw : out_f32;
    if (infer_body(B, R, masked, mask, ctx, cmask, init_latents, vae_noise, dst, st)) return -1;
    if (composite || out_u8) {
        if (composite) {
            stage_begin(ST_POST, st);
            KCHECK(launch_composite(canvas, raw, B, R, out_f32, out_u8, st));
            stage_end(ST_POST, st);
        } else {
            return fail("stamp: uint8 output is only produced together with compositing");
        }
    }
    return 0;
}

long long Engine::counter(const char* name) const {
    const std::string n = name ? name : "";
    if (n == "launches") return launches_;
    if (n == "stamps") return stamps_;
    if (n == "arena_peak") return static_cast<long long>(arena_.peak());
    if (n == "arena_bytes") return static_cast<long long>(cfg_.arena_bytes);
    if (n == "graph_launches") return graph_launches_;
    if (n == "unet_plan_ops") return static_cast<long long>(unet_plan_.ops.size());
    if (n == "ws_bytes") return static_cast<long long>(ws_bytes_);
    if (n == "device") return device_;
    if (n.rfind("stage_us_", 0) == 0 || n.rfind("stage_n_", 0) == 0) {
        const_cast<Engine*>(this)->stage_collect();
        const bool is_us = n[6] == 'u';
        const int k = atoi(n.c_str() + (is_us ? 9 : 8));
        if (k < 0 || k >= ST_NUM) return -1;
        return is_us ? static_cast<long long>(stage_us_[k]) : stage_n_[k];
    }
    if (n.rfind("prof_us_", 0) == 0 || n.rfind("prof_n_", 0) == 0) {
        const_cast<Profiler&>(prof_).collect();
        const bool is_us = n[5] == 'u';
        const int k = atoi(n.c_str() + (is_us ? 8 : 7));
        if (k < 0 || k >= K_NUM) return -1;
        return is_us ? static_cast<long long>(prof_.us[k]) : prof_.n[k];
    }
    return -1;
}

int Engine::profile_dump(const char* path) {
    prof_.collect();
    FILE* f = fopen(path, "w");
    if (!f) return fail(std::string("profile_dump: cannot open ") + path);
    std::vector<std::pair<std::string, std::pair<double, long long>>> v(prof_.by_label.begin(), prof_.by_label.end());
    std::sort(v.begin(), v.end(), [](const auto& a, const auto& b) { return a.second.first > b.second.first; });
    fprintf(f, "total_us,calls,avg_us,op\n");
    for (auto& e : v)
        fprintf(f, "%.1f,%lld,%.2f,%s\n", e.second.first, e.second.second, e.second.first / e.second.second,
                e.first.c_str());
    fclose(f);
    return 0;
}

int Engine::set_option(const char* name, int value) {
    const std::string n = name ? name : "";
    if (n == "sync_check") {
        opt_sync_check_ = value;
        return 0;
    }
    if (n == "graph") {
        opt_graph_ = value;
        return 0;
    }
    if (n == "fold_cross") {
        opt_fold_cross_ = value;
        unet_plan_.clear();
        g_infer_.key.clear();
        g_stamp_.key.clear();
        return 0;
    }
    if (n == "fuse_cross") {
        opt_fuse_cross_ = value;
        unet_plan_.clear();
        g_infer_.key.clear();
        g_stamp_.key.clear();
        return 0;
    }
    if (n == "fuse_ff_out") {
        opt_fuse_ff_out_ = value;
        unet_plan_.clear();
        g_infer_.key.clear();
        g_stamp_.key.clear();
        return 0;
    }
    if (n == "splitk_f16") {
        gemm_set_splitk_half(value);
        g_infer_.key.clear();
        g_stamp_.key.clear();
        return 0;
    }
    if (n == "dedup_branches") {
        opt_dedup_branches_ = value;
        unet_plan_.clear();
        g_infer_.key.clear();
        g_stamp_.key.clear();
        return 0;
    }
    if (n == "fold_downsample") {
        opt_fold_downsample_ = value;
        unet_plan_.clear();
        vae_enc_plan_.clear();
        g_infer_.key.clear();
        g_stamp_.key.clear();
        return 0;
    }
    if (n == "fold_upsample" || n == "fold_upsample_rows") {
        (n == "fold_upsample" ? opt_fold_upsample_ : opt_fold_upsample_rows_) = value;
        unet_plan_.clear();
        vae_dec_plan_.clear();
        g_infer_.key.clear();
        g_stamp_.key.clear();
        return 0;
    }
    if (n == "fuse_shortcut") {
        opt_fuse_shortcut_ = value;
        unet_plan_.clear();
        vae_enc_plan_.clear();
        vae_dec_plan_.clear();
        g_infer_.key.clear();
        g_stamp_.key.clear();
        return 0;
    }
    if (n == "fold_ln_ff_rows") {
        opt_fold_ln_ff_rows_ = value;
        unet_plan_.clear();
        g_infer_.key.clear();
        g_stamp_.key.clear();
        return 0;
    }
    if (n == "fold_ln") {
        opt_fold_ln_ = value;
        unet_plan_.clear();
        g_infer_.key.clear();
        g_stamp_.key.clear();
        return 0;
    }
    if (n == "flash") {
        opt_flash_ = value;
        unet_plan_.clear();
        g_infer_.key.clear();
        g_stamp_.key.clear();
        return 0;
    }
    if (n == "debug_skip_kinds") {
        g_debug_skip_kinds = value;
        g_infer_.key.clear();
        g_stamp_.key.clear();
        return 0;
    }
    if (n == "debug_skip_label") {
        g_debug_skip_label = value;
        g_infer_.key.clear();
        g_stamp_.key.clear();
        return 0;
    }
    if (n == "profile") {
        prof_.reset();
        prof_.on = value != 0;
        return 0;
    }
    if (n == "stage_timers") {
        stage_collect();
        for (int i = 0; i < ST_NUM; ++i) {
            stage_us_[i] = 0;
            stage_n_[i] = 0;
        }
        opt_stage_timers_ = value;
        return 0;
    }
    if (n == "nvtx") {
        opt_nvtx_ = value;
        return 0;
    }
    if (n == "arena_mib") return grow_arena(static_cast<size_t>(value) << 20);
    return fail("unknown option " + n);
}

}  // namespace dtp

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
using dtp::Engine;
struct dtp_engine {
    Engine* e;
};
static char g_create_err[256] = "";

// Every entry point runs on the device the handle was created on, with the handle's kernel context current (kctx.h).
struct ApiScope {
    int prev_dev = -1;
    dtp::KernelCtxScope k;
    explicit ApiScope(Engine* e) : k(e->kctx()) {
        if (cudaGetDevice(&prev_dev) != cudaSuccess) prev_dev = -1;
        if (prev_dev != e->device()) cudaSetDevice(e->device());
        else prev_dev = -1;
    }
    ~ApiScope() {
        if (prev_dev >= 0) cudaSetDevice(prev_dev);
    }
};
#define API_ENTER(h)                         \
    if (!(h) || !(h)->e) return -1;          \
    ApiScope api_scope_((h)->e);             \
    if ((h)->e->check_device_error()) return -5

extern "C" {

int dtp_create(const dtp_config* cfg, dtp_handle** out) {
    if (!cfg || !out) return -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        snprintf(g_create_err, sizeof(g_create_err), "no CUDA device");
        return -2;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    if (prop.major != 10) {
        snprintf(g_create_err, sizeof(g_create_err), "device sm_%d%d is not sm_100: this library is Blackwell-only",
                 prop.major, prop.minor);
        return -3;
    }
    dtp_handle* h = new dtp_engine;
    h->e = new Engine(*cfg);
    if (!h->e->ok()) {
        snprintf(g_create_err, sizeof(g_create_err), "%s", h->e->last_error());
        delete h->e;
        delete h;
        return -4;
    }
    *out = h;
    return 0;
}
void dtp_destroy(dtp_handle* h) {
    if (!h) return;
    {
        ApiScope scope(h->e);
        delete h->e;
        h->e = nullptr;
    }
    delete h;
}
const char* dtp_last_error(dtp_handle* h) { return h ? h->e->last_error() : g_create_err; }
int dtp_set_tensor(dtp_handle* h, const char* name, const void* host_ptr, const long long* shape, int ndim, int dtype) {
    API_ENTER(h);
    return h->e->set_tensor(name, host_ptr, reinterpret_cast<const int64_t*>(shape), ndim, dtype);
}
int dtp_finalize_weights(dtp_handle* h) {
    API_ENTER(h);
    return h->e->finalize_weights();
}
int dtp_encode_patches(dtp_handle* h, const float* patches, float* emb_out, void* stream, void*) {
    API_ENTER(h);
    return h->e->encode_patches(patches, emb_out, (cudaStream_t)stream);
}
int dtp_set_condition(dtp_handle* h, const float* emb, const float* uncond, void* stream) {
    API_ENTER(h);
    return h->e->set_condition(emb, uncond, (cudaStream_t)stream);
}
int dtp_set_schedule(dtp_handle* h, int n, const float* timesteps, const float* alpha_t, const float* alpha_prev,
                     float cfg, float tg, int tg_steps) {
    API_ENTER(h);
    return h->e->set_schedule(n, timesteps, alpha_t, alpha_prev, cfg, tg, tg_steps);
}
int dtp_infer(dtp_handle* h, int B, int R, const float* masked_img, const float* mask, const float* ctx_img,
              const float* ctx_mask, const float* init_latents, const float* vae_noise, float* out_images, void* stream) {
    API_ENTER(h);
    return h->e->infer(B, R, masked_img, mask, ctx_img, ctx_mask, init_latents, vae_noise, out_images,
                       (cudaStream_t)stream);
}
int dtp_stamp(dtp_handle* h, int B, int R, const float* canvas, const float* brush, int pad, const float* init_latents,
              const float* vae_noise, int composite, float* out_f32, unsigned char* out_u8, void* stream) {
    API_ENTER(h);
    return h->e->stamp(B, R, canvas, brush, pad, init_latents, vae_noise, composite, out_f32, out_u8,
                       (cudaStream_t)stream);
}
int dtp_vae_encode(dtp_handle* h, int Nb, int R, const float* images, const float* noise, float* latents_out,
                   void* stream) {
    API_ENTER(h);
    return h->e->vae_encode(Nb, R, images, noise, latents_out, (cudaStream_t)stream);
}
int dtp_vae_decode(dtp_handle* h, int B, int R, const float* latents, float* images_out, void* stream) {
    API_ENTER(h);
    return h->e->vae_decode(B, R, latents, images_out, (cudaStream_t)stream);
}
int dtp_unet_forward(dtp_handle* h, int B, int R, const float* sample, const float*, const float*, int step,
                     float* eps_out, void* stream) {
    API_ENTER(h);
    return h->e->unet_forward(B, R, sample, nullptr, nullptr, nullptr, step, eps_out, (cudaStream_t)stream);
}
long long dtp_get_counter(dtp_handle* h, const char* name) {
    if (!h || !h->e) return -1;
    ApiScope scope(h->e);
    return h->e->counter(name);
}
int dtp_set_option(dtp_handle* h, const char* name, int value) {
    API_ENTER(h);
    return h->e->set_option(name, value);
}
int dtp_profile_dump(dtp_handle* h, const char* path) {
    API_ENTER(h);
    return h->e->profile_dump(path);
}

}  // extern "C"
