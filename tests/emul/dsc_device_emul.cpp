// dsc_device_emul.cpp -- host-memory stand-in for dsc_device.h.  TEST INFRASTRUCTURE ONLY
// (tests/emul/libdsc_emul.so); see cuda_shim.h.  "Device" memory is malloc'ed host memory,
// copies are memcpy, streams are synchronous.
#include "dsc_device.h"

#include <chrono>
#include <cstdlib>
#include <cstring>

namespace dscdev {
struct Event { double t; };
namespace {
size_t g_calls = 0;
double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
}
int device_count() { return getenv("DSC_EMUL_NO_DEVICE") ? 0 : 1; }
void set_device(int) {}
size_t free_memory() { return (size_t)8 << 30; }
const char *device_name() { return "pthread-emulated device (tests only)"; }
void *arena_alloc(size_t bytes) { ++g_calls; void *p = nullptr; if (posix_memalign(&p, 4096, bytes)) return nullptr; return p; }
void arena_free(void *p) { free(p); }
size_t arena_alloc_calls() { return g_calls; }
bool host_pin(void *, size_t) { return true; }
void host_unpin(void *) {}
void *stream(int) { return nullptr; }
void stream_sync(int) {}
void sync_all() {}
void copy_h2d(void *d, const void *s, size_t n, int) { memcpy(d, s, n); }
void copy_d2h(void *d, const void *s, size_t n, int) { memcpy(d, s, n); }
void copy_d2h_2d(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t rows, int) {
    for (size_t r = 0; r < rows; ++r) memcpy((char *)d + r * dp, (const char *)s + r * sp, w);
}
Event *event_record(int) { return new Event{now_ms()}; }
void stream_wait(int, Event *) {}
float event_ms(Event *a, Event *b) { return (float)(b->t - a->t); }
void event_wait(Event *) {}
void event_release(Event *e) { delete e; }
}  // namespace dscdev
