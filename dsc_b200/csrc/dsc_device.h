// dsc_device.h -- the handful of CUDA runtime services the host runtime needs, behind plain
// functions so that tests/emul can substitute host memory for them (dsc_device_emul.cpp).
// The product implementation is dsc_device_cuda.cpp.
#pragma once

#include <cstddef>

namespace dscdev {

struct Event;   // opaque

// Number of usable devices (0: the context becomes host-only and every FFT entry point aborts).
int device_count();
void set_device(int ordinal);
size_t free_memory();
const char *device_name();

// The ONE device allocation of a context, and its release.
void *arena_alloc(size_t bytes);
void arena_free(void *p);
size_t arena_alloc_calls();

// Page-lock an existing host range so copies run at full PCIe rate and asynchronously.
bool host_pin(void *p, size_t bytes);
void host_unpin(void *p);

// Streams: 0 = compute, 1 = upload, 2 = download.
void *stream(int which);
void stream_sync(int which);
void sync_all();

void copy_h2d(void *dst_dev, const void *src_host, size_t bytes, int which_stream);
void copy_d2h(void *dst_host, const void *src_dev, size_t bytes, int which_stream);
// rows x width_bytes sub-matrix (pitches in bytes), device -> host
void copy_d2h_2d(void *dst_host, size_t dst_pitch, const void *src_dev, size_t src_pitch,
                 size_t width_bytes, size_t rows, int which_stream);

// Cross-stream ordering and timing.
Event *event_record(int which_stream);          // from a small recycled pool
void stream_wait(int which_stream, Event *e);
float event_ms(Event *start, Event *stop);      // both must have completed
void event_wait(Event *e);                      // block the host until the event has completed
void event_release(Event *e);

}  // namespace dscdev
