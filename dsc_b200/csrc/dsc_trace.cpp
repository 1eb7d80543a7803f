// dsc_trace.cpp -- Chrome / Perfetto trace recorder.
//
// Keeps the reference's contract (/root/reference/dsc/src/dsc_tracing.cpp:25-46, 260-280):
// a fixed array of DSC_MAX_TRACES records made at init, 'B'/'E' events with CLOCK_MONOTONIC
// microsecond timestamps, pid, the calling thread as tid, silently dropping records once full,
// and a dump that is one JSON array of {"name","cat","ph","ts","pid","tid","args"}.
// New here: device spans.  A kernel (or copy) is bracketed by two events on its stream; at dump
// time each pair becomes a complete ('X') event with "dur", placed on a synthetic tid per
// stream and anchored to the host clock through an event recorded when recording was enabled.
#include "dsc_runtime.h"

#include <cstdarg>
#include <cstring>
#include <ctime>
#include <pthread.h>
#include <unistd.h>

namespace {

constexpr int NAME_MAX_ = 48, CAT_MAX_ = 16, ARGS_MAX_ = 512;   // two 4-D tensor descriptions plus the op's own fields (about 400 bytes)

struct record {
    char name[NAME_MAX_];
    char cat[CAT_MAX_];
    char args[ARGS_MAX_];
    u64 ts;                    // microseconds
    u64 tid;
    int pid;
    char phase;                // 'B', 'E' or 'X' (device span, resolved at dump time)
    int stream_id;
    dscdev::Event *start, *stop;
};

struct state {
    record *records;
    u64 count, capacity;
    bool recording;
    dscdev::Event *anchor;     // device-time origin ...
    u64 anchor_us;             // ... and the host time it corresponds to
};

state *g = nullptr;

u64 now_us() noexcept {
    timespec ts{};
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (u64) ts.tv_sec * 1000000ULL + (u64) ts.tv_nsec / 1000ULL;
}

record *next_record() noexcept {
    if (g == nullptr || g->count >= g->capacity) return nullptr;   // full: drop, like the reference
    return &g->records[g->count++];
}

void copy_str(char *dst, int cap, const char *src) noexcept {
    if (src == nullptr) { dst[0] = '\0'; return; }
    strncpy(dst, src, (size_t) cap - 1);
    dst[cap - 1] = '\0';
}

}  // namespace

void dsc_trace_init(const u64 max_traces) noexcept {
#if defined(DSC_ENABLE_TRACING)
    if (g != nullptr) return;
    g = (state *) calloc(1, sizeof(state));
    g->records = (record *) malloc(max_traces * sizeof(record));
    DSC_ASSERT(g->records != nullptr);
    g->capacity = max_traces;
#else
    DSC_UNUSED(max_traces);
#endif
}

void dsc_trace_shutdown() noexcept {
    if (g == nullptr) return;
    dsc_trace_clear();
    free(g->records);
    free(g);
    g = nullptr;
}

bool dsc_trace_recording() noexcept { return g != nullptr && g->recording; }

void dsc_trace_set_recording(const bool on) noexcept {
    if (g == nullptr) return;
    if (on && !g->recording && dscdev::device_count() > 0) {
        // tie the device timeline to the host clock once per recording session
        dscdev::sync_all();
        if (g->anchor) dscdev::event_release(g->anchor);
        g->anchor = dscdev::event_record(0);
        dscdev::stream_sync(0);
        g->anchor_us = now_us();
    }
    g->recording = on;
}

void dsc_trace_event(const char phase, const char *name, const char *cat, const char *args_json) noexcept {
    record *r = next_record();
    if (r == nullptr) return;
    copy_str(r->name, NAME_MAX_, name);
    copy_str(r->cat, CAT_MAX_, cat);
    copy_str(r->args, ARGS_MAX_, args_json);
    r->ts = now_us();
    r->pid = (int) getpid();
    r->tid = (u64) pthread_self();
    r->phase = phase;
    r->stream_id = -1;
    r->start = r->stop = nullptr;
}

void dsc_trace_gpu_span(const char *name, const char *cat, const int stream_id,
                        dscdev::Event *start, dscdev::Event *stop, const char *args_json) noexcept {
    record *r = next_record();
    if (r == nullptr) {
        dscdev::event_release(start);
        dscdev::event_release(stop);
        return;
    }
    copy_str(r->name, NAME_MAX_, name);
    copy_str(r->cat, CAT_MAX_, cat);
    copy_str(r->args, ARGS_MAX_, args_json);
    r->ts = 0;
    r->pid = (int) getpid();
    r->tid = 0;
    r->phase = 'X';
    r->stream_id = stream_id;
    r->start = start;
    r->stop = stop;
}

int dsc_trace_describe_tensor(char *dst, const int cap, const dsc_tensor *x) noexcept {
    // {"shape": "[4, 64]", "dtype": "f32", "backend": "CUDA", "addr": "0x.."}; 1-D shapes are bare
    // numbers, as in the reference (dsc_tracing.cpp:48-60, 86-96)
    char shape[64];
    int o = 0;
    if (x->n_dim > 1) {
        o += snprintf(shape + o, sizeof(shape) - o, "\"[");
        for (int i = 0; i < x->n_dim; ++i)
            o += snprintf(shape + o, sizeof(shape) - o, "%d%s", x->shape[DSC_MAX_DIMS - x->n_dim + i],
                          i < x->n_dim - 1 ? ", " : "");
        snprintf(shape + o, sizeof(shape) - o, "]\"");
    } else {
        snprintf(shape, sizeof(shape), "%d", x->shape[DSC_MAX_DIMS - 1]);
    }
    return snprintf(dst, (size_t) cap, "{\"shape\": %s, \"dtype\": \"%s\", \"backend\": \"%s\", \"addr\": \"0x%zx\"}",
                    shape, DSC_DTYPE_NAMES[x->dtype], DSC_BACKEND_NAMES[x->backend], (size_t) x->data);
}

void dsc_trace_dump(const char *filename) noexcept {
    if (g == nullptr) return;
    dscdev::sync_all();   // device spans must have completed before their events are read
    FILE *f = fopen(filename, "wt");
    DSC_ASSERT(f != nullptr);
    fprintf(f, "[\n");
    for (u64 i = 0; i < g->count; ++i) {
        const record *r = &g->records[i];
        if (r->phase == 'X') {
            const double start_ms = g->anchor ? (double) dscdev::event_ms(g->anchor, r->start) : 0.0;
            const double dur_ms = (double) dscdev::event_ms(r->start, r->stop);
            fprintf(f, "\t{\"name\": \"%s\", \"cat\": \"%s\", \"ph\": \"X\", \"ts\": %.3f, \"dur\": %.3f, \"pid\": %d, \"tid\": %d",
                    r->name, r->cat, (double) g->anchor_us + start_ms * 1e3, dur_ms * 1e3, r->pid, 1000000 + r->stream_id);
        } else {
            fprintf(f, "\t{\"name\": \"%s\", \"cat\": \"%s\", \"ph\": \"%c\", \"ts\": %lu, \"pid\": %d, \"tid\": %lu",
                    r->name, r->cat, r->phase, (unsigned long) r->ts, r->pid, (unsigned long) r->tid);
        }
        if (r->args[0] != '\0') fprintf(f, ", \"args\": %s", r->args);
        fprintf(f, "}%s\n", i + 1 < g->count ? "," : "");
    }
    fprintf(f, "]");
    fclose(f);
    DSC_LOG_INFO("exported Perfetto-compatible traces to \"%s\"", filename);
}

void dsc_trace_clear() noexcept {
    if (g == nullptr) return;
    for (u64 i = 0; i < g->count; ++i) {
        if (g->records[i].phase == 'X') {
            dscdev::event_release(g->records[i].start);
            dscdev::event_release(g->records[i].stop);
        }
    }
    g->count = 0;
}
