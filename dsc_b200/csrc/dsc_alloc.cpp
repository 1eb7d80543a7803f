// dsc_alloc.cpp -- dsc_range_alloc: best-fit allocation with address-ordered coalescing over
// an abstract byte range, bookkeeping in a node pool made once at init.
//
// Same policy as the reference's general-purpose allocator (best fit, split, coalesce with
// both neighbours on free: /root/reference/dsc/src/dsc_allocator.cpp:49-197) but with the
// metadata out of band, so one implementation serves the host arena and the device arena.
#include "dsc_runtime.h"

void dsc_range_alloc::init(const usize bytes, const usize granule_, const int max_nodes) noexcept {
    granule = granule_;
    capacity = bytes / granule * granule;
    capacity_nodes = max_nodes;
    nodes = (node *) malloc(sizeof(node) * (usize) max_nodes);
    DSC_ASSERT(nodes != nullptr);
    reset();
}

void dsc_range_alloc::destroy() noexcept {
    free(nodes);
    nodes = nullptr;
}

void dsc_range_alloc::reset() noexcept {
    // slot 0 describes the whole range as one free block; the rest form the unused stack
    nodes[0] = node{0, capacity, -1, -1, false};
    head = 0;
    free_nodes = -1;
    for (int i = capacity_nodes - 1; i >= 1; --i) {
        nodes[i].used = false;
        nodes[i].next = free_nodes;
        free_nodes = i;
    }
    used = 0;
}

int dsc_range_alloc::take_node() noexcept {
    if (free_nodes < 0) DSC_LOG_FATAL("allocator node pool exhausted (%d blocks)", capacity_nodes);
    const int id = free_nodes;
    free_nodes = nodes[id].next;
    return id;
}

void dsc_range_alloc::give_node(const int id) noexcept {
    nodes[id].used = false;
    nodes[id].size = 0;
    nodes[id].next = free_nodes;
    free_nodes = id;
}

int dsc_range_alloc::alloc(const usize bytes) noexcept {
    const usize need = DSC_ALIGN(DSC_MAX(bytes, (usize) 1), granule);
    int best = -1;
    for (int i = head; i >= 0; i = nodes[i].next) {
        if (!nodes[i].used && nodes[i].size >= need && (best < 0 || nodes[i].size < nodes[best].size)) {
            best = i;
            if (nodes[i].size == need) break;
        }
    }
    if (best < 0) return -1;
    node &b = nodes[best];
    if (b.size > need) {
        // split: the tail stays free
        const int t = take_node();
        nodes[t] = node{b.off + need, nodes[best].size - need, best, nodes[best].next, false};
        if (nodes[best].next >= 0) nodes[nodes[best].next].prev = t;
        nodes[best].next = t;
        nodes[best].size = need;
    }
    nodes[best].used = true;
    used += need;
    return best;
}

void dsc_range_alloc::release(const int id) noexcept {
    if (id < 0 || id >= capacity_nodes || !nodes[id].used) return;   // double free: ignore
    nodes[id].used = false;
    used -= nodes[id].size;
    // merge with the next block
    const int nx = nodes[id].next;
    if (nx >= 0 && !nodes[nx].used) {
        nodes[id].size += nodes[nx].size;
        nodes[id].next = nodes[nx].next;
        if (nodes[nx].next >= 0) nodes[nodes[nx].next].prev = id;
        give_node(nx);
    }
    // merge into the previous block
    const int pv = nodes[id].prev;
    if (pv >= 0 && !nodes[pv].used) {
        nodes[pv].size += nodes[id].size;
        nodes[pv].next = nodes[id].next;
        if (nodes[id].next >= 0) nodes[nodes[id].next].prev = pv;
        give_node(id);
    }
}
