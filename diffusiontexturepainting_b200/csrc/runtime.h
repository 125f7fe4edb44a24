// Native runtime of the stamp path: weight store, activation arena, launch plans (UNet / VAE / image encoder) and the
// per-stamp driver. Replaces the TensorRT engine layer of the reference (trt_inference/utilities.py:54-264 Engine,
// stable_diffusion_pipeline.py:189-338 loadEngines/runEngine) and the graphs of trt_inference/models.py.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/dtp.h"
#include "gemm_tc.h"
#include "kctx.h"

namespace dtp {

struct WT {  // one named weight tensor
    void* dev = nullptr;
    std::vector<int64_t> shape;
    int dtype = 0;  // 0 = f32, 1 = f16
    std::vector<float> host;  // host copy of f32 tensors (small: biases, norm affine, embeddings)
    int64_t numel() const {
        int64_t n = 1;
        for (auto s : shape) n *= s;
        return n;
    }
};

// First-fit offset allocator over one device slab; plans carve their transient activations from it at build time.
class Arena {
   public:
    void init(char* base, size_t cap) {
        base_ = base;
        cap_ = cap;
        reset();
    }
    void reset() {
        free_.clear();
        free_[0] = cap_;
        used_ = peak_ = 0;
    }
    void* alloc(size_t bytes);
    void release(void* p);
    size_t peak() const { return peak_; }
    size_t capacity() const { return cap_; }

   private:
    char* base_ = nullptr;
    size_t cap_ = 0, used_ = 0, peak_ = 0;
    std::map<size_t, size_t> free_;                 // offset -> size
    std::unordered_map<size_t, size_t> live_;       // offset -> size
};

struct Act {  // NHWC fp16 activation: rows = N*H*W pixels, C channels
    __half* p = nullptr;
    int N = 0, H = 0, W = 0, C = 0;
    long long rows() const { return static_cast<long long>(N) * H * W; }
    size_t bytes() const { return static_cast<size_t>(rows()) * C * sizeof(__half); }
};

enum OpKind { K_OTHER = 0, K_GEMM = 1, K_GROUPNORM = 2, K_LAYERNORM = 3, K_SOFTMAX = 4, K_ATTN_SMALL = 5, K_FLASH = 6, K_NUM = 7 };

// Optional per-op device timing (CUDA events on the launching stream), enabled with dtp_set_option("profile", 1).
struct Profiler {
    bool on = false;
    std::vector<cudaEvent_t> pool;
    struct Rec {
        int kind;
        int launches;
        cudaEvent_t a, b;
        const std::string* label;
    };
    std::map<std::string, std::pair<double, long long>> by_label;  // label -> (us, calls)
    std::vector<Rec> recs;
    size_t next = 0;
    double us[K_NUM] = {0};
    long long n[K_NUM] = {0};
    cudaEvent_t get();
    void collect();
    void reset();
};

struct Plan {
    std::vector<std::function<int(cudaStream_t)>> ops;
    std::vector<int> kinds;
    std::vector<std::string> labels;
    int key_a = -1, key_b = -1;  // (batch, resolution) the plan was built for
    int run(cudaStream_t st, long long* launch_counter, Profiler* prof = nullptr) const;
    void add(int kind, std::function<int(cudaStream_t)> f, const std::string& label = std::string()) {
        ops.push_back(std::move(f));
        kinds.push_back(kind);
        labels.push_back(label);
    }
    void clear() {
        ops.clear();
        kinds.clear();
        labels.clear();
        key_a = key_b = -1;
    }
};

// LayerNorm folded into the consuming contraction (gemm_tc.h): per transformer layer, prepared at finalize_weights
// (gamma-scaled weights, their row sums, W beta) and per brush (the folded cross-attention score operand)
struct LnFold {
    __half* qkv_w = nullptr;   // (3C, C)  to_qkv diag(norm1.weight)
    float *qkv_cs = nullptr, *qkv_b = nullptr;
    __half* ff1_w = nullptr;   // (8C, C)  ff.net.0.proj diag(norm3.weight), GEGLU row order
    float *ff1_cs = nullptr, *ff1_b = nullptr;
    __half* q_w = nullptr;     // (C, C)   attn2.to_q diag(norm2.weight)
    float* q_beta = nullptr;   // (C)      attn2.to_q norm2.bias
    __half* wscore = nullptr;  // (3, heads*16, C) per brush: scale * K_h (Wq diag(gamma))_h
    float *ws_cs = nullptr, *ws_b = nullptr;  // (3, heads*16)
};

// ff.net.2 and proj_out of a transformer block are two linear maps with nothing but a residual between them:
//   out = (g W2^T + b2 + h) Wp^T + bp + x = [g | h] [Wp W2 | Wp]^T + (Wp b2 + bp) + x
// -> one contraction over K = 4C + C with the product weights prepared at finalize_weights (per transformer layer)
struct FfOut {
    __half* W = nullptr;  // (C, 5C)
    float* bias = nullptr;
};

class Engine {
   public:
    explicit Engine(const dtp_config& cfg);
    ~Engine();

    int set_tensor(const char* name, const void* host, const int64_t* shape, int ndim, int dtype);
    int finalize_weights();
    int encode_patches(const float* patches, float* emb_out, cudaStream_t st);
    int set_condition(const float* emb, const float* uncond, cudaStream_t st);
    int set_schedule(int n, const float* ts, const float* a_t, const float* a_prev, float cfg, float tg, int tg_steps);
    int infer(int B, int R, const float* masked_img, const float* mask, const float* ctx_img, const float* ctx_mask,
              const float* init_latents, const float* vae_noise, float* out_images, cudaStream_t st);
    int stamp(int B, int R, const float* canvas, const float* brush, int pad, const float* init_latents,
              const float* vae_noise, int composite, float* out_f32, unsigned char* out_u8, cudaStream_t st);
    int infer_body(int B, int R, const float* masked_img, const float* mask, const float* ctx_img,
                   const float* ctx_mask, const float* init_latents, const float* vae_noise, float* out_images,
                   cudaStream_t st);
    int stamp_body(int B, int R, const float* canvas, const float* brush, int pad, const float* init_latents,
                   const float* vae_noise, int composite, float* out_f32, unsigned char* out_u8, cudaStream_t st);
    int vae_encode(int Nb, int R, const float* images, const float* noise, float* latents_out, cudaStream_t st);
    int vae_decode(int B, int R, const float* latents, float* images_out, cudaStream_t st);
    int unet_forward(int B, int R, const float* sample, const float* latents, const float* mask3, const float* masked3,
                     int step, float* eps_out, cudaStream_t st);
    long long counter(const char* name) const;
    int set_option(const char* name, int value);
    int profile_dump(const char* path);
    const char* last_error() const { return err_.c_str(); }
    void set_error(const std::string& m) { err_ = m; }
    KernelCtx* kctx() const { return kctx_; }
    int device() const { return device_; }
    bool ok() const { return kctx_ != nullptr; }
    // called at the top of every C-ABI entry point: reports a cross-CTA wait that gave up during an earlier call
    int check_device_error();
    int grow_arena(size_t bytes);
    bool fold_cross() const { return opt_fold_cross_ != 0; }
    // conv2 and conv_shortcut of a ResnetBlock2D concatenated along K ([cout, 9*cout + cin], bias = conv2.bias +
    // conv_shortcut.bias): built on first use, cached per resnet prefix
    struct FusedShortcut {
        __half* W = nullptr;
        float* bias = nullptr;
    };
    const FusedShortcut* fused_shortcut(const std::string& prefix, int cout, int cin);
    // Upsample2D conv weights folded into the four parity-class 2x2 convolutions ([4*cout, 4*cin], kernels.h
    // launch_upconv_fold_weights): built on first use, cached per conv prefix
    const __half* upconv_weights(const std::string& key, const __half* W, int cout, int cin);
    bool fold_ln() const { return opt_fold_ln_ != 0 && !ln_.empty(); }
    const LnFold& ln(int i) const { return ln_[i]; }
    bool fuse_cross() const { return opt_fuse_cross_ != 0; }
    bool fuse_ff_out() const { return opt_fuse_ff_out_ != 0 && !ffo_.empty(); }
    const FfOut& ffo(int i) const { return ffo_[i]; }
    long long fold_ln_ff_rows() const { return opt_fold_ln_ff_rows_; }
    int tf_index(const std::string& prefix) const {
        for (size_t i = 0; i < tf_names_.size(); ++i)
            if (tf_names_[i] == prefix) return static_cast<int>(i);
        return 0;
    }
    const __half* wscore(int i) const { return wscore_[i]; }
    const __half* wout(int i) const { return wout_[i]; }

   private:
    friend struct Builder;
    int fail(const std::string& msg) {
        err_ = msg;
        return -1;
    }
    const WT* find(const std::string& name);
    void* persistent(size_t bytes, bool zero);
    int ensure_arena();
    int build_unet_plan(int B, int R, bool shared_input);
    bool unet_plan_dedup_ = false;
    int build_vae_enc_plan(int Nb, int R);
    int build_vae_dec_plan(int B, int R);
    int build_encoder_plan();
    int build_cond_plan();
    int build_temb_tables(cudaStream_t st);
    int ensure_ws();
    int ensure_io(int B, int R);

    dtp_config cfg_;
    std::unordered_map<std::string, WT> w_;
    std::vector<void*> persistent_;
    bool finalized_ = false;
    std::string err_;

    char* arena_base_ = nullptr;
    Arena arena_;
    float* ws_ = nullptr;  // split-K workspace shared by all contractions (stream-ordered reuse)
    size_t ws_bytes_ = 0, ws_needed_ = 0;

    Plan unet_plan_, vae_enc_plan_, vae_dec_plan_, enc_plan_, cond_plan_;
    // plan I/O (persistent)
    __half* unet_in_ = nullptr;      // (3B, h, h, 64) packed sample
    float* unet_eps_ = nullptr;      // (3B, 4, h, h)
    __half* vae_enc_in_ = nullptr;   // (Nb, R, R, 64)
    float* vae_moments_ = nullptr;   // (Nb, 8, h, h)
    __half* vae_dec_in_ = nullptr;   // (B, h, h, 8)
    float* vae_dec_out_ = nullptr;   // (B, 3, R, R)
    size_t unet_io_cap_ = 0, vae_enc_cap_ = 0, vae_dec_cap_ = 0;
    int* kv_index_ = nullptr;
    int kv_index_B_ = -1;
    // per-infer scratch (persistent, sized by ensure_io)
    float *lat_ = nullptr, *mask3_ = nullptr, *masked3_ = nullptr, *lat2_ = nullptr, *pre_ = nullptr;
    size_t io_cap_B_ = 0, io_cap_R_ = 0;

    // encoder plan I/O
    float* enc_patches_ = nullptr;  // borrowed per call
    float* enc_in_copy_ = nullptr;  // (14,3,224,224) staging so the plan has a fixed address
    float* enc_out_ = nullptr;      // (14, cross) f32

    // condition
    __half* ctx_ = nullptr;  // (2*14, cross) f16: rows 0..13 uncond, 14..27 cond
    float* ctx_f32_ = nullptr;
    std::vector<__half*> cross_kv_;  // per transformer layer: (28, 2C)
    std::vector<__half*> wscore_;    // per layer: (3, heads*16, C)  scale * K_h Wq_h, zero rows for the pad tokens
    std::vector<__half*> wout_;      // per layer: (3, C, heads*16)  Wo_h V_h^T
    int opt_fold_cross_ = 1;
    std::vector<LnFold> ln_;
    std::vector<FfOut> ffo_;
    int opt_fuse_ff_out_ = 1;
    int opt_fuse_cross_ = 1;  // score + output contraction of the folded cross-attention as one kernel (cross_attn.cu)
    int prepare_ff_out();
    std::unordered_map<std::string, FusedShortcut> fused_sc_;
    int opt_fuse_shortcut_ = 1;
    std::unordered_map<std::string, __half*> upconv_w_;
    int opt_dedup_branches_ = 1;         // layers in front of the first cross-attention once for the uncond and the cond branch
    int opt_fold_downsample_ = 1;        // stride-2 convolutions read the input through parity-view tensor maps (no im2col buffer)
    int opt_fold_upsample_ = 1;          // nearest-2x upsample folded into its convolution (gemm_setup_upconv2x)
    int opt_fold_upsample_rows_ = 3072;  // ... for outputs of at least this many pixels (below, the 16/9 larger folded weights cost more than the multiply-adds save)
    int opt_fold_ln_ = 1;
    int opt_fold_ln_ff_rows_ = 1024;  // norm3 -> FF1 folding only for activations of at most this many rows
    int prepare_ln_fold();
    std::vector<std::string> tf_names_;  // transformer prefixes in execution order
    bool cond_set_ = false;

    // schedule
    int n_steps_ = 0, tg_steps_ = 0;
    float cfg_w_ = 2.0f, tg_w_ = 0.0f;
    std::vector<float> ts_, a_t_, a_prev_;
    float* temb_all_ = nullptr;   // (n_steps, temb_total) f32: conv1.bias + time_emb_proj(SiLU(temb_s)) per resnet
    float* temb_cur_ = nullptr;   // (temb_total) row of the step being evaluated
    int temb_total_ = 0;
    int temb_cap_steps_ = 0;
    std::vector<std::pair<std::string, int>> resnets_;  // UNet resnet prefix -> offset into the temb row
    int cur_step_ = 0;
    bool temb_dirty_ = true;

    // CUDA-graph replay of a whole stamp (all launches of pre-process, 2x VAE encode, N UNet evaluations, decode)
    struct GraphSlot {
        std::string key;
        cudaGraphExec_t exec = nullptr;
        int warm = 0;
        long long launches = 0;
    };
    GraphSlot g_infer_, g_stamp_;
    cudaStream_t gstream_ = nullptr;
    cudaEvent_t gev_in_ = nullptr, gev_out_ = nullptr;
    int run_graphed(GraphSlot& slot, const std::string& key, const std::function<int(cudaStream_t)>& body,
                    cudaStream_t st);
    std::string schedule_key() const;
    float* stage_ = nullptr;           // fixed-address staging of caller inputs / outputs for graph replay
    unsigned char* stage_u8_ = nullptr;
    int opt_graph_ = 1;
    long long graph_launches_ = 0;

    long long launches_ = 0, stamps_ = 0;
    // per-stage device timers (the reference's cudart events + NVTX ranges 'vae_encoder' / 'unet' / 'vae',
    // stable_diffusion_pipeline.py:146-149,358-366,486-503); dtp_set_option("stage_timers" / "nvtx", 1), eager mode only
    enum Stage { ST_PRE = 0, ST_VAE_ENC, ST_UNET, ST_DDIM, ST_VAE_DEC, ST_POST, ST_NUM };
    struct StageRec {
        int stage;
        cudaEvent_t a, b;
    };
    std::vector<StageRec> stage_recs_;
    std::vector<cudaEvent_t> stage_pool_;
    size_t stage_next_ = 0;
    double stage_us_[ST_NUM] = {0};
    long long stage_n_[ST_NUM] = {0};
    int opt_stage_timers_ = 0, opt_nvtx_ = 0;
    void stage_begin(int stage, cudaStream_t st);
    void stage_end(int stage, cudaStream_t st);
    void stage_collect();
    Profiler prof_;
    KernelCtx* kctx_ = nullptr;  // cross-CTA synchronisation state of this engine's kernels (kctx.h)
    int device_ = 0;
    int opt_sync_check_ = 0;
    int opt_flash_ = 1;
};

}  // namespace dtp
