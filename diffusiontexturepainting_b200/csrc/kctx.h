// Per-engine device-side state of the kernels that synchronise across CTAs (split-K arrival tickets of the contraction
// kernel, per-sample barrier / tickets of the GroupNorm kernels) plus the device facts their launch logic needs.
//
// Round 1 kept these as process-global statics allocated on whichever device called first; two engines (or an engine and the
// operator entry points) on different devices then dereferenced another device's pointers, and two engines on one device
// shared ticket slots. Now every Engine owns a KernelCtx (created with the engine, on the engine's device, never during
// graph capture) and makes it current for the duration of each API call; the operator entry points of capi_ops.cu use a
// lazily created per-device default context. A handle remains single-stream / not thread-safe (include/dtp.h).
//
// err_flag: mapped pinned host int. A cross-CTA wait that exceeds its bound (a CTA of the grid never became resident because
// the GPU is shared with another client) sets it and gives up instead of __trap()-ing the context; the engine reports it at
// the next API call.
#pragma once
#include <cuda_runtime.h>

namespace dtp {

struct KernelCtx {
    int device = -1;
    int sms = 148;
    int max_pair_clusters = -1;      // co-resident 2-CTA clusters of the pair-mode contraction kernel (lazy)
    int* tile_counters = nullptr;    // [kTileCounterSlots] split-K tickets, zero between launches
    unsigned* gn_barrier = nullptr;  // [kGnSamples * 2]
    int* gn_counters = nullptr;      // [kGnSamples]
    int* err_flag_host = nullptr;    // pinned + mapped
    int* err_flag_dev = nullptr;     // device alias of err_flag_host
};

constexpr int kTileCounterSlots = 1 << 16;
constexpr int kGnSamples = 1024;
constexpr int kMaxDevices = 64;

KernelCtx* kctx_create();               // on the current device; nullptr on allocation failure
void kctx_destroy(KernelCtx* c);
KernelCtx* kctx_current();              // the calling thread's current context, else the current device's default one
void kctx_set_current(KernelCtx* c);    // nullptr: back to the per-device default
int kctx_take_error(KernelCtx* c);      // returns and clears the error flag (0 = none)

struct KernelCtxScope {  // RAII: Engine API calls
    KernelCtx* prev;
    explicit KernelCtxScope(KernelCtx* c);
    ~KernelCtxScope();
};

// device ordinal of the calling thread (for per-device "attribute already set" flags)
int kctx_device();

}  // namespace dtp
