// dsc_device_cuda.cpp -- CUDA runtime implementation of dsc_device.h.
// Errors follow the library convention: message on stderr, exit(EXIT_FAILURE).
#include "dsc_device.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CUDA_OK(call)                                                                        \
    do {                                                                                     \
        const cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                             \
            fprintf(stderr, "dsc(cuda): %s failed: %s\n", #call, cudaGetErrorString(e_));    \
            exit(EXIT_FAILURE);                                                              \
        }                                                                                    \
    } while (0)

namespace dscdev {

struct Event { cudaEvent_t ev; };

namespace {
cudaStream_t g_streams[3] = {nullptr, nullptr, nullptr};
bool g_streams_ready = false;
size_t g_arena_calls = 0;
std::vector<Event *> g_free_events;
char g_name[256] = "";

void ensure_streams() {
    if (g_streams_ready) return;
    for (auto &s : g_streams) CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    g_streams_ready = true;
}
}  // namespace

int device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void set_device(int ordinal) { CUDA_OK(cudaSetDevice(ordinal)); }

size_t free_memory() {
    size_t fr = 0, tot = 0;
    CUDA_OK(cudaMemGetInfo(&fr, &tot));
    return fr;
}

const char *device_name() {
    cudaDeviceProp p;
    int dev = 0;
    CUDA_OK(cudaGetDevice(&dev));
    CUDA_OK(cudaGetDeviceProperties(&p, dev));
    snprintf(g_name, sizeof(g_name), "%s", p.name);
    return g_name;
}

void *arena_alloc(size_t bytes) {
    void *p = nullptr;
    CUDA_OK(cudaMalloc(&p, bytes));
    ++g_arena_calls;
    return p;
}

void arena_free(void *p) { if (p) CUDA_OK(cudaFree(p)); }
size_t arena_alloc_calls() { return g_arena_calls; }

bool host_pin(void *p, size_t bytes) {
    if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return false; }
    return true;
}
void host_unpin(void *p) { if (cudaHostUnregister(p) != cudaSuccess) cudaGetLastError(); }

void *stream(int which) { ensure_streams(); return g_streams[which]; }
void stream_sync(int which) { ensure_streams(); CUDA_OK(cudaStreamSynchronize(g_streams[which])); }
void sync_all() { if (g_streams_ready) for (auto s : g_streams) CUDA_OK(cudaStreamSynchronize(s)); }

void copy_h2d(void *dst, const void *src, size_t bytes, int which) {
    ensure_streams();
    CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_streams[which]));
}
void copy_d2h(void *dst, const void *src, size_t bytes, int which) {
    ensure_streams();
    CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_streams[which]));
}

void copy_d2h_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, int which) {
    ensure_streams();
    CUDA_OK(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToHost, g_streams[which]));
}

Event *event_record(int which) {
    ensure_streams();
    Event *e;
    if (!g_free_events.empty()) { e = g_free_events.back(); g_free_events.pop_back(); }
    else { e = new Event; CUDA_OK(cudaEventCreate(&e->ev)); }
    CUDA_OK(cudaEventRecord(e->ev, g_streams[which]));
    return e;
}
void stream_wait(int which, Event *e) { ensure_streams(); CUDA_OK(cudaStreamWaitEvent(g_streams[which], e->ev, 0)); }
float event_ms(Event *a, Event *b) {
    float ms = 0.f;
    CUDA_OK(cudaEventSynchronize(b->ev));
    CUDA_OK(cudaEventElapsedTime(&ms, a->ev, b->ev));
    return ms;
}
void event_wait(Event *e) { CUDA_OK(cudaEventSynchronize(e->ev)); }
void event_release(Event *e) { if (e) g_free_events.push_back(e); }

}  // namespace dscdev
