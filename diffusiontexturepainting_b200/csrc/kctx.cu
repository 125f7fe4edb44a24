// See kctx.h.
#include "kctx.h"

#include <mutex>

namespace dtp {

static thread_local KernelCtx* t_current = nullptr;
static KernelCtx* g_default[kMaxDevices] = {nullptr};
static std::mutex g_default_mu;

int kctx_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        dev = 0;
    }
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

KernelCtx* kctx_create() {
    KernelCtx* c = new KernelCtx;
    c->device = kctx_device();
    if (cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, c->device) != cudaSuccess || c->sms <= 0) c->sms = 148;
    bool ok = cudaMalloc(&c->tile_counters, kTileCounterSlots * sizeof(int)) == cudaSuccess &&
              cudaMalloc(&c->gn_barrier, kGnSamples * 2 * sizeof(unsigned)) == cudaSuccess &&
              cudaMalloc(&c->gn_counters, kGnSamples * sizeof(int)) == cudaSuccess &&
              cudaHostAlloc(&c->err_flag_host, sizeof(int), cudaHostAllocMapped) == cudaSuccess;
    if (ok) {
        *c->err_flag_host = 0;
        ok = cudaHostGetDevicePointer(&c->err_flag_dev, c->err_flag_host, 0) == cudaSuccess &&
             cudaMemset(c->tile_counters, 0, kTileCounterSlots * sizeof(int)) == cudaSuccess &&
             cudaMemset(c->gn_barrier, 0, kGnSamples * 2 * sizeof(unsigned)) == cudaSuccess &&
             cudaMemset(c->gn_counters, 0, kGnSamples * sizeof(int)) == cudaSuccess;
    }
    if (!ok) {
        cudaGetLastError();
        kctx_destroy(c);
        return nullptr;
    }
    return c;
}

void kctx_destroy(KernelCtx* c) {
    if (!c) return;
    if (t_current == c) t_current = nullptr;
    if (c->tile_counters) cudaFree(c->tile_counters);
    if (c->gn_barrier) cudaFree(c->gn_barrier);
    if (c->gn_counters) cudaFree(c->gn_counters);
    if (c->err_flag_host) cudaFreeHost(c->err_flag_host);
    delete c;
}

KernelCtx* kctx_current() {
    if (t_current) return t_current;
    const int dev = kctx_device();
    std::lock_guard<std::mutex> lk(g_default_mu);
    if (!g_default[dev]) g_default[dev] = kctx_create();
    return g_default[dev];
}

void kctx_set_current(KernelCtx* c) { t_current = c; }

int kctx_take_error(KernelCtx* c) {
    if (!c || !c->err_flag_host) return 0;
    volatile int* f = c->err_flag_host;
    const int v = *f;
    if (v) *f = 0;
    return v;
}

KernelCtxScope::KernelCtxScope(KernelCtx* c) : prev(t_current) { t_current = c; }
KernelCtxScope::~KernelCtxScope() { t_current = prev; }

}  // namespace dtp
