// prims.cu -- exclusive scan + stable radix sort (see prims.cuh).
#include "prims.cuh"

#include <string.h>

#include <mutex>
#include <unordered_map>

namespace ob {

// ------------------------------------------------------------------------------ guarded device memory
// (engine.h: ob_malloc / ob_free; ODE_B200_DEBUG_GUARD=1)

namespace {
constexpr size_t GUARD = 256; // cudaMalloc's own alignment, so guarded pointers stay as aligned as plain ones
constexpr unsigned char GUARD_BYTE = 0xA5;
struct GuardRec {
    size_t bytes;
    const char *file;
    int line, device;
};
std::mutex g_guard_mu;
std::unordered_map<void *, GuardRec> g_guard_recs;
bool guards_on() {
    static int v = -1;
    if (v < 0) {
        const char *s = getenv("ODE_B200_DEBUG_GUARD");
        v = (s && s[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
} // namespace

cudaError_t guard_malloc(void **p, size_t bytes, const char *file, int line) {
    if (!guards_on()) return cudaMalloc(p, bytes);
    const size_t body = (bytes + GUARD - 1) / GUARD * GUARD; // the tail band starts at the next 256-byte boundary ...
    unsigned char *base = nullptr;
    cudaError_t err = cudaMalloc(&base, body + 2 * GUARD);
    if (err != cudaSuccess) return err;
    // ... and the slack between the requested size and that boundary carries the pattern too
    if ((err = cudaMemset(base, GUARD_BYTE, GUARD)) != cudaSuccess) return err;
    if ((err = cudaMemset(base + GUARD + bytes, GUARD_BYTE, body - bytes + GUARD)) != cudaSuccess) return err;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_guard_mu);
        g_guard_recs[base + GUARD] = GuardRec{bytes, file, line, dev};
    }
    *p = base + GUARD;
    return cudaSuccess;
}

cudaError_t guard_free(void *p) {
    if (!p || !guards_on()) return cudaFree(p);
    {
        std::lock_guard<std::mutex> lk(g_guard_mu);
        g_guard_recs.erase(p);
    }
    return cudaFree(static_cast<unsigned char *>(p) - GUARD);
}

int guard_check(int verbose) {
    if (!guards_on()) return -1;
    std::lock_guard<std::mutex> lk(g_guard_mu);
    int cur = 0, bad = 0;
    cudaGetDevice(&cur);
    std::vector<unsigned char> host;
    for (const auto &kv : g_guard_recs) {
        const GuardRec &r = kv.second;
        unsigned char *user = static_cast<unsigned char *>(kv.first);
        const size_t body = (r.bytes + GUARD - 1) / GUARD * GUARD, tail = body - r.bytes + GUARD;
        cudaSetDevice(r.device);
        cudaDeviceSynchronize();
        host.resize(GUARD + tail);
        if (cudaMemcpy(host.data(), user - GUARD, GUARD, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(host.data() + GUARD, user + r.bytes, tail, cudaMemcpyDeviceToHost) != cudaSuccess) {
            if (verbose) fprintf(stderr, "libode_b200: guard check could not read %s:%d\n", r.file, r.line);
            bad++;
            continue;
        }
        long first_before = -1, first_after = -1;
        for (size_t i = 0; i < GUARD; i++)
            if (host[i] != GUARD_BYTE) { first_before = (long)(GUARD - i); break; }
        for (size_t i = 0; i < tail; i++)
            if (host[GUARD + i] != GUARD_BYTE) { first_after = (long)i; break; }
        if (first_before >= 0 || first_after >= 0) {
            bad++;
            if (verbose)
                fprintf(stderr, "libode_b200: guard band damaged around the %zu-byte allocation of %s:%d (%ld bytes before it, %ld bytes past its end)\n",
                        r.bytes, r.file, r.line, first_before, first_after);
        }
    }
    cudaSetDevice(cur);
    return bad;
}

// ------------------------------------------------------------------------------------------ scan

__device__ __forceinline__ int warp_scan_incl(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// exclusive scan across the block; *total = block sum. smem: 32 ints. All threads must call.
__device__ __forceinline__ int block_scan_excl(int v, int *smem, int *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = warp_scan_incl(v);
    if (lane == 31) smem[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = (lane < nw) ? smem[lane] : 0;
        int wi = warp_scan_incl(w);
        smem[lane] = wi - w; // exclusive offsets of warps
        if (lane == 31) smem[32] = wi;
    }
    __syncthreads();
    int res = incl - v + smem[wid];
    *total = smem[32];
    __syncthreads();
    return res;
}

__device__ __forceinline__ long scan_len(long n_max, const int *n_dev) {
    if (!n_dev) return n_max;
    long n = *n_dev;
    return n < n_max ? n : n_max;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const int *__restrict__ in, long n_max,
                                                               const int *__restrict__ n_dev,
                                                               int *__restrict__ sums) {
    __shared__ int sm[33];
    const long n = scan_len(n_max, n_dev);
    const long base = (long)blockIdx.x * SCAN_TILE;
    int v = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        long i = base + (long)j * SCAN_THREADS + threadIdx.x;
        if (i < n) v += in[i];
    }
    int tot;
    block_scan_excl(v, sm, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const int *__restrict__ in, int *__restrict__ out,
                                                              long n_max, const int *__restrict__ n_dev,
                                                              const int *__restrict__ sums) {
    __shared__ int sm[33];
    const long n = scan_len(n_max, n_dev);
    const long base = (long)blockIdx.x * SCAN_TILE;
    if (base >= n) return;
    int carry = sums[blockIdx.x];
#pragma unroll 1
    for (int j = 0; j < SCAN_ITEMS; j++) {
        long i = base + (long)j * SCAN_THREADS + threadIdx.x;
        int v = (i < n) ? in[i] : 0;
        int tot;
        int ex = block_scan_excl(v, sm, &tot);
        if (i < n) out[i] = carry + ex;
        carry += tot;
    }
}

// one block scans everything (used for short arrays and for the block-sum levels)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_single(const int *__restrict__ in, int *__restrict__ out,
                                                               long n_max, const int *__restrict__ n_dev,
                                                               int *__restrict__ total) {
    // every thread owns SI consecutive elements per pass (scanned in registers), so a pass of the block covers 8192 elements
    // with one block scan: the 14 k per-block class sums of a 2048-env broadphase take 2 passes instead of 28 (16 -> 4 us)
    constexpr int SI = 16;
    __shared__ int sm[33];
    const long n = scan_len(n_max, n_dev);
    int carry = 0;
    for (long b = 0; b < n; b += (long)SCAN_THREADS * SI) {
        const long i0 = b + (long)threadIdx.x * SI;
        int v[SI];
        int s = 0;
#pragma unroll
        for (int k = 0; k < SI; k++) v[k] = (i0 + k < n) ? in[i0 + k] : 0;
#pragma unroll
        for (int k = 0; k < SI; k++) { const int t = v[k]; v[k] = s; s += t; }
        int tot;
        const int ex = block_scan_excl(s, sm, &tot);
#pragma unroll
        for (int k = 0; k < SI; k++)
            if (i0 + k < n) out[i0 + k] = carry + ex + v[k];
        carry += tot;
    }
    if (total && threadIdx.x == 0) *total = carry;
}

static void ws_ensure(ScanWorkspace &ws, int level, size_t n) {
    if (ws.cap[level] >= n) return;
    if (ws.sums[level]) OB_CUDA(ob_free(ws.sums[level]));
    size_t cap = n + n / 2 + 64;
    OB_CUDA(ob_malloc(&ws.sums[level], cap * sizeof(int)));
    ws.cap[level] = cap;
}

static void scan_level(const int *in, int *out, long n_max, const int *n_dev, int *total, ScanWorkspace &ws,
                       int level, cudaStream_t st) {
    const long single_limit = 4 * SCAN_TILE;
    if (n_max <= single_limit || level >= 3) {
        k_scan_single<<<1, SCAN_THREADS, 0, st>>>(in, out, n_max, n_dev, total);
        OB_CHECK_KERNEL("k_scan_single", st);
        return;
    }
    long nb = (n_max + SCAN_TILE - 1) / SCAN_TILE;
    ws_ensure(ws, level, (size_t)nb);
    int *sums = ws.sums[level];
    k_scan_reduce<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n_max, n_dev, sums);
    OB_CHECK_KERNEL("k_scan_reduce", st);
    scan_level(sums, sums, nb, nullptr, total, ws, level + 1, st);
    k_scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, n_max, n_dev, sums);
    OB_CHECK_KERNEL("k_scan_apply", st);
}

void scan_exclusive(const int *in, int *out, long n_max, const int *n_dev, int *total, ScanWorkspace &ws,
                    cudaStream_t st) {
    if (n_max <= 0) {
        if (total) OB_CUDA(cudaMemsetAsync(total, 0, sizeof(int), st));
        return;
    }
    scan_level(in, out, n_max, n_dev, total, ws, 0, st);
}

void scan_workspace_free(ScanWorkspace &ws) {
    for (int i = 0; i < 3; i++) {
        if (ws.sums[i]) ob_free(ws.sums[i]);
        ws.sums[i] = nullptr;
        ws.cap[i] = 0;
    }
}

// ------------------------------------------------------------------------------------ radix sort

constexpr int SORT_WARPS = 8;

__global__ void __launch_bounds__(SORT_WARPS * 32) k_sort_hist(const uint32_t *__restrict__ keys, long n_max,
                                                                const int *__restrict__ n_dev, int shift,
                                                                int nchunks, int *__restrict__ hist) {
    __shared__ int sh[SORT_WARPS][256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long n = scan_len(n_max, n_dev);
    const int chunk = blockIdx.x * SORT_WARPS + wid;
    for (int d = lane; d < 256; d += 32) sh[wid][d] = 0;
    __syncwarp();
    if (chunk < nchunks) {
        const long base = (long)chunk * SORT_CHUNK;
        for (int r = 0; r < SORT_CHUNK / 32; r++) {
            long i = base + r * 32 + lane;
            if (i < n) atomicAdd(&sh[wid][(keys[i] >> shift) & 255u], 1);
        }
        __syncwarp();
        for (int d = lane; d < 256; d += 32) hist[(long)d * nchunks + chunk] = sh[wid][d];
    }
}

__global__ void __launch_bounds__(SORT_WARPS * 32) k_sort_scatter(const uint32_t *__restrict__ keys,
                                                                   const int *__restrict__ vals,
                                                                   uint32_t *__restrict__ keys_out,
                                                                   int *__restrict__ vals_out, long n_max,
                                                                   const int *__restrict__ n_dev, int shift,
                                                                   int nchunks, const int *__restrict__ hist) {
    __shared__ int base[SORT_WARPS][256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long n = scan_len(n_max, n_dev);
    const int chunk = blockIdx.x * SORT_WARPS + wid;
    if (chunk >= nchunks) return;
    for (int d = lane; d < 256; d += 32) base[wid][d] = hist[(long)d * nchunks + chunk];
    __syncwarp();
    const long cbase = (long)chunk * SORT_CHUNK;
    const unsigned lt = (1u << lane) - 1u;
    for (int r = 0; r < SORT_CHUNK / 32; r++) {
        long i = cbase + r * 32 + lane;
        bool valid = i < n;
        unsigned act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            uint32_t k = keys[i];
            int v = vals[i];
            unsigned d = (k >> shift) & 255u;
            unsigned m = __match_any_sync(act, d);
            int rank = __popc(m & lt);
            int pos = base[wid][d] + rank;
            __syncwarp(act);
            if (rank == 0) base[wid][d] += __popc(m);
            __syncwarp(act);
            keys_out[pos] = k;
            vals_out[pos] = v;
        }
    }
}

void sort_pairs(uint32_t *keys, int *vals, long n_max, const int *n_dev, int bits, SortWorkspace &ws,
                cudaStream_t st) {
    if (n_max <= 0) return;
    const int passes = (bits + 7) / 8;
    const int nchunks = (int)((n_max + SORT_CHUNK - 1) / SORT_CHUNK);
    size_t need = (size_t)256 * nchunks;
    if (ws.cap < need) {
        if (ws.hist) OB_CUDA(ob_free(ws.hist));
        ws.cap = need + need / 2;
        OB_CUDA(ob_malloc(&ws.hist, ws.cap * sizeof(int)));
    }
    if (ws.cap_items < (size_t)n_max) {
        if (ws.keys_tmp) OB_CUDA(ob_free(ws.keys_tmp));
        if (ws.vals_tmp) OB_CUDA(ob_free(ws.vals_tmp));
        ws.cap_items = (size_t)n_max + (size_t)n_max / 2;
        OB_CUDA(ob_malloc(&ws.keys_tmp, ws.cap_items * sizeof(uint32_t)));
        OB_CUDA(ob_malloc(&ws.vals_tmp, ws.cap_items * sizeof(int)));
    }
    uint32_t *kin = keys, *kout = ws.keys_tmp;
    int *vin = vals, *vout = ws.vals_tmp;
    const unsigned nblk = (unsigned)((nchunks + SORT_WARPS - 1) / SORT_WARPS);
    for (int p = 0; p < passes; p++) {
        k_sort_hist<<<nblk, SORT_WARPS * 32, 0, st>>>(kin, n_max, n_dev, p * 8, nchunks, ws.hist);
        OB_CHECK_KERNEL("k_sort_hist", st);
        scan_exclusive(ws.hist, ws.hist, (long)need, nullptr, nullptr, ws.scan, st);
        k_sort_scatter<<<nblk, SORT_WARPS * 32, 0, st>>>(kin, vin, kout, vout, n_max, n_dev, p * 8, nchunks,
                                                         ws.hist);
        OB_CHECK_KERNEL("k_sort_scatter", st);
        uint32_t *tk = kin; kin = kout; kout = tk;
        int *tv = vin; vin = vout; vout = tv;
    }
    if (kin != keys) {
        OB_CUDA(cudaMemcpyAsync(keys, kin, (size_t)n_max * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        OB_CUDA(cudaMemcpyAsync(vals, vin, (size_t)n_max * sizeof(int), cudaMemcpyDeviceToDevice, st));
    }
}

void sort_workspace_free(SortWorkspace &ws) {
    if (ws.hist) ob_free(ws.hist);
    if (ws.keys_tmp) ob_free(ws.keys_tmp);
    if (ws.vals_tmp) ob_free(ws.vals_tmp);
    ws.hist = nullptr; ws.keys_tmp = nullptr; ws.vals_tmp = nullptr;
    ws.cap = 0; ws.cap_items = 0;
    scan_workspace_free(ws.scan);
}

} // namespace ob
