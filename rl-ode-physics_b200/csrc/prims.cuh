// prims.cuh -- device-wide exclusive scan and stable LSD radix sort used by the broadphase,
// manifold compaction and colour ordering.  Hand-written (no CUB): the sort is a warp-chunk
// counting sort per 8-bit digit with __match_any_sync ranking, so equal keys keep their input
// order and every result is deterministic (no atomics decide an output position).
#pragma once

#include "engine.h"

namespace ob {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS; // 4096

struct ScanWorkspace {
    int *sums[3] = {nullptr, nullptr, nullptr};
    size_t cap[3] = {0, 0, 0};
};

// exclusive scan of in[0..n) into out (out may alias in). n = n_dev ? min(*n_dev, n_max) : n_max.
// If total != nullptr the grand total is written there.
void scan_exclusive(const int *in, int *out, long n_max, const int *n_dev, int *total, ScanWorkspace &ws,
                    cudaStream_t st);
void scan_workspace_free(ScanWorkspace &ws);

constexpr int SORT_CHUNK = 512; // items per warp chunk

struct SortWorkspace {
    int *hist = nullptr; // 256 * nchunks
    size_t cap = 0;
    uint32_t *keys_tmp = nullptr;
    int *vals_tmp = nullptr;
    size_t cap_items = 0;
    ScanWorkspace scan;
};

// stable sort of (keys, vals) by the low `bits` bits of keys, n = n_dev ? *n_dev : n_max items.
// Results end in keys / vals (ping-pong through the workspace handled inside).
void sort_pairs(uint32_t *keys, int *vals, long n_max, const int *n_dev, int bits, SortWorkspace &ws,
                cudaStream_t st);
void sort_workspace_free(SortWorkspace &ws);

} // namespace ob
