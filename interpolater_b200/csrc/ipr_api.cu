// interpolater_b200 — C ABI implementation: device-resident plan, kernel scheduling, and the
// one-shot host-buffer entry points with the multi-GPU scheduler (cell ranges per device).
// See include/interpolater_b200.h for the contract and the reference lines each call replaces.
#include "interpolater_b200.h"   // include/ (build adds -I)

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <chrono>
#include <utility>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define IPR_HAVE_SSE2 1
#endif
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "ipr_kernels.cuh"

using namespace ipr;

namespace {

void set_err(char* err, size_t errlen, const char* fmt, ...) {
    if (!err || errlen == 0) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, errlen, fmt, ap);
    va_end(ap);
}

#define IPR_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            set_err(err, errlen, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, \
                    __LINE__, cudaGetErrorString(e_));                                          \
            return (e_ == cudaErrorMemoryAllocation) ? IPR_ENOMEM : IPR_ECUDA;                  \
        }                                                                                       \
    } while (0)

inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

// IPR_TRACE=1: phase timestamps of the host-buffer path on stderr
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    Trace() : on(getenv("IPR_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[ipr trace] %8.2f ms  %s\n", ms, what);
    }
};


// No C++ exception may cross the C ABI (the caller is R's .Call or ctypes): host allocations of the
// staging vectors are the realistic source.
template <class F>
int no_throw(char* err, size_t errlen, F&& body) {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        set_err(err, errlen, "out of host memory");
        return IPR_ENOMEM;
    } catch (const std::exception& e) {
        set_err(err, errlen, "internal error: %s", e.what());
        return IPR_ECUDA;
    }
}

// ---------------------------------------------------------------------------------------------
// Device buffer pool.  cudaMalloc / cudaFree of the multi-GB work buffers (A-operand cache, output
// ring, B operand) was measured to cost anything from 15 to 600 ms per host-buffer call depending
// on the allocator's mood; freed buffers are therefore kept per device and reused by exact size.
// ipr_release_cached_memory() (or memory pressure) returns them to the driver.
// ---------------------------------------------------------------------------------------------
struct DevicePool {
    std::mutex mu;
    std::map<void*, std::pair<int, size_t>> live;              // ptr -> (device, bytes)
    std::multimap<std::pair<int, size_t>, void*> idle;         // (device, bytes) -> ptr
    size_t idle_bytes = 0;
    void drain_locked() {
        int cur = 0;
        cudaGetDevice(&cur);
        for (auto& kv : idle) {
            cudaSetDevice(kv.first.first);
            cudaFree(kv.second);
        }
        idle.clear();
        idle_bytes = 0;
        cudaSetDevice(cur);
    }
};
DevicePool& pool() {
    static DevicePool p;
    return p;
}
size_t pool_limit_bytes() {
    static const size_t v = [] {
        const char* e = getenv("IPR_POOL_MAX_GB");
        return (size_t)(e ? atoll(e) : 64) << 30;
    }();
    return v;
}

// streams are recycled per device as well (creation / destruction occasionally takes tens of ms on a
// shared host); a stream is only returned after it has been synchronised
struct StreamCache {
    std::mutex mu;
    std::multimap<int, cudaStream_t> idle;
};
StreamCache& stream_cache() {
    static StreamCache c;
    return c;
}
cudaError_t acquire_stream(cudaStream_t* out) {
    int dev = 0;
    cudaGetDevice(&dev);
    {
        StreamCache& C = stream_cache();
        std::lock_guard<std::mutex> lk(C.mu);
        auto it = C.idle.find(dev);
        if (it != C.idle.end()) {
            *out = it->second;
            C.idle.erase(it);
            return cudaSuccess;
        }
    }
    return cudaStreamCreateWithFlags(out, cudaStreamNonBlocking);
}
void release_stream(cudaStream_t s) {
    if (!s) return;
    int dev = 0;
    cudaGetDevice(&dev);
    StreamCache& C = stream_cache();
    std::lock_guard<std::mutex> lk(C.mu);
    if (C.idle.count(dev) < 16) C.idle.insert({dev, s});
    else cudaStreamDestroy(s);
}

cudaError_t dev_malloc_raw(void** out, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    DevicePool& P = pool();
    std::lock_guard<std::mutex> lk(P.mu);
    auto it = P.idle.find({dev, bytes});
    if (it != P.idle.end()) {
        *out = it->second;
        P.idle_bytes -= bytes;
        P.idle.erase(it);
        P.live[*out] = {dev, bytes};
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {  // give cached buffers back and retry once
        cudaGetLastError();
        P.drain_locked();
        e = cudaMalloc(out, bytes);
    }
    if (e == cudaSuccess) P.live[*out] = {dev, bytes};
    return e;
}
template <class T>
cudaError_t dev_malloc(T** out, size_t bytes) {
    return dev_malloc_raw(reinterpret_cast<void**>(out), bytes);
}
void dev_free(void* ptr) {
    if (!ptr) return;
    DevicePool& P = pool();
    std::lock_guard<std::mutex> lk(P.mu);
    auto it = P.live.find(ptr);
    if (it == P.live.end()) {
        cudaFree(ptr);
        return;
    }
    const int dev = it->second.first;
    const size_t bytes = it->second.second;
    P.live.erase(it);
    // small buffers are kept too: a plan frees ~20 of them, and cudaFree was traced at up to 0.5 s for
    // the lot on a busy host (it synchronises the device and unmaps), 0.2-0.5 s out of a 0.57 s call
    if (P.idle_bytes + bytes <= pool_limit_bytes() && P.idle.size() < 4096) {
        P.idle.insert({{dev, bytes}, ptr});
        P.idle_bytes += bytes;
    } else {
        cudaFree(ptr);
    }
}

inline double canon_zero(double v) { return v == 0.0 ? 0.0 : v; }  // -0 -> +0

// Z-order (Morton) sort of points on a 65536^2 lattice spanning their bounding box
std::vector<int> morton_order(const double* x, const double* y, int n, int xdiv = 1) {
    double x0 = x[0], x1 = x[0], y0 = y[0], y1 = y[0];
    for (int i = 1; i < n; i++) {
        x0 = std::min(x0, x[i]); x1 = std::max(x1, x[i]);
        y0 = std::min(y0, y[i]); y1 = std::max(y1, y[i]);
    }
    const double span = std::max(std::max(x1 - x0, y1 - y0), 1e-300);
    auto spread = [](uint32_t v) {
        uint64_t r = v & 0xFFFF;
        r = (r | (r << 8)) & 0x00FF00FF;
        r = (r | (r << 4)) & 0x0F0F0F0F;
        r = (r | (r << 2)) & 0x33333333;
        r = (r | (r << 1)) & 0x55555555;
        return r;
    };
    std::vector<std::pair<uint64_t, int>> key(n);
    for (int i = 0; i < n; i++) {
        const uint32_t ix = (uint32_t)std::min(65535.0, (x[i] - x0) / span * 65535.0) / (uint32_t)xdiv;
        const uint32_t iy = (uint32_t)std::min(65535.0, (y[i] - y0) / span * 65535.0);
        key[i] = {spread(ix) | (spread(iy) << 1), i};
    }
    std::sort(key.begin(), key.end());
    std::vector<int> perm(n);
    for (int i = 0; i < n; i++) perm[i] = key[i].second;
    return perm;
}

}  // namespace

struct ipr_plan {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    long long n_cells = 0, c_pad = 0;
    int n_st = 0, s_pad = 0, n_dates = 0, n_gran = 0, n_stages = 0;
    size_t z_doubles = 0;  // size of the granule-block arrays
    double2* d_cell = nullptr;
    double2* d_st = nullptr;
    double* d_obs = nullptr;  // raw, column-major n_dates x n_st (ld = n_dates)
    double* d_z0 = nullptr;
    double* d_mask = nullptr;
    uint8_t* d_flag = nullptr;
    // Kriging fallback (ipr_plan_idw_kriging): per-date constant-layer flags and values, built on first use
    uint8_t* d_kflag = nullptr;
    double* d_kconst = nullptr;
    bool kflags_ready = false;
    std::vector<uint8_t> h_has_na_raw;   // any NA on the date, before the sparse-NA reclassification of IDW()
    uint8_t* d_has_na = nullptr;
    int* d_counters = nullptr;  // [0] zero pairs, [1] non-finite obs
    int2* d_pairs = nullptr;
    int pair_capacity = 0;
    // zero-distance CSR (device) built once per geometry
    int n_fix = 0;
    int* d_fix_cell = nullptr;
    int* d_fix_off = nullptr;
    int* d_fix_st = nullptr;
    long long zero_pairs = 0;
    std::vector<uint8_t> h_has_na;     // per date: the date's granule needs the masked contraction
    // sparse-NA correction (granules with few missing observations are contracted unmasked)
    int* d_na_n = nullptr;
    int* d_na_idx = nullptr;
    uint8_t* d_date_mode = nullptr;    // 1: denominator of this date is corrected by idw_sparse_na_fix_kernel
    std::vector<uint8_t> h_date_mode;
    bool any_sparse_fix = false;
    bool obs_ready = false;
    long long launches = 0;
    // error-free int8 contraction of the availability mask (IPR_MASK_INT8=0 disables): digit planes of the
    // weights, byte mask, list of (cell, date) pairs that need a direct FP64 evaluation
    bool int8_mask = false;
    int n_kblocks = 0;
    double* d_wmax = nullptr;
    unsigned char* d_wq = nullptr;   // digit planes, UMMA B-operand image per (32-cell group, k-block)
    unsigned char* d_mask8 = nullptr;
    int2* d_fb_list = nullptr;
    int* d_fb_count = nullptr;
    int fb_cap = 1 << 22;
    int wq_mode = -1;
    double wq_param = 0.0;
    int wq_tile0 = -1;
    // optional post-pass (rounding, polygon mask) applied to every finished layer
    int post_round_mode = 0;
    double post_p10 = 1.0;
    uint8_t* d_cell_keep = nullptr;
    // A-operand cache (weights in DMMA fragment order) for a chunk of cell tiles + per-cell sums
    double* d_wfrag = nullptr;
    double* d_wsum = nullptr;
    int* d_stage_list = nullptr;  // per cell tile of the chunk: stages with any non-zero weight
    int* d_n_active = nullptr;
    int* d_perm = nullptr;        // spatial (Morton) order of the stations: sorted position -> input index
    double4* d_stage_box = nullptr;  // bounding box of each station stage (Cressman culling in weights_kernel)
    int* d_run_perm = nullptr;    // Morton order of the 8-cell runs: 8 consecutive entries = one compact cell tile
    std::vector<double2> h_run_xy;   // first cell of every run; d_run_perm is built from it on the first Cressman call
    bool run_perm_ready = false;
    std::map<int, int*> run_perm_alt;   // the same with x compressed by 4 / 16 / 64: flatter tiles (ensure_run_perm_alt)
    unsigned long long* d_active_total = nullptr;
    unsigned long long stage_tiles_total = 0;  // (cell tile, stage) blocks the dense product would visit
    int w_chunk_tiles = 0;   // capacity of d_wfrag in cell tiles
    int w_mode = -1;         // what the cache currently holds: weight mode, parameter, first tile
    double w_param = 0.0;
    int w_tile0 = -1;
    // Cressman: per-CTA scratch of the persistent kernel (station list + A panel of the CTA's current cell tile)
    double* d_cr_panel = nullptr;
    int* d_cr_list = nullptr;
    int* d_cr_counter = nullptr;
    double st_bbox_area = 0.0;   // of the stations' bounding box: station density for the Cressman tile-shape choice
    int cr_grid = 0;
    // optional per-phase event timing (ipr_plan_profile)
    bool prof_on = false;
    int phase_radius = -1;   // Cressman: radius index of the contraction being launched
    struct Span { int id; cudaEvent_t e0, e1; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> ev_pool;
};

namespace {

// RAII span of one profiling phase on the plan stream (no-op unless ipr_plan_profile enabled it)
struct PhaseScope {
    ipr_plan* p;
    cudaEvent_t e1 = nullptr;
    PhaseScope(ipr_plan* plan, int id) : p(plan) {
        if (!p->prof_on) return;
        cudaEvent_t ev[2] = {nullptr, nullptr};
        for (auto& e : ev) {
            if (!p->ev_pool.empty()) {
                e = p->ev_pool.back();
                p->ev_pool.pop_back();
            } else if (cudaEventCreate(&e) != cudaSuccess) {
                e = nullptr;
            }
        }
        if (!ev[0] || !ev[1]) return;
        cudaEventRecord(ev[0], p->stream);
        e1 = ev[1];
        p->spans.push_back({id, ev[0], ev[1]});
    }
    void finish() {
        if (e1) cudaEventRecord(e1, p->stream);
        e1 = nullptr;
    }
    ~PhaseScope() { finish(); }
};

// date tiles of one launch group: [MASKED][NP] -> start dates
struct TileSets {
    std::vector<int> u[9];  // unmasked, NP = 2, 4, 6, 8 used
    std::vector<int> m[5];  // masked,   NP = 2, 4 used
};
void build_tiles(const ipr_plan* p, int t0, int t1, bool use_mask, TileSets& ts, bool raw_na = false);

template <int NP, bool MASKED, int KSTEPS = kKSteps>
int launch_contract(ipr_plan* p, ContractArgs& a, const std::vector<int>& tiles, char* err, size_t errlen) {
    using L = SmemLayout<MASKED, KSTEPS>;
    static std::atomic<bool> configured[64];   // zero-initialised; racing first calls both set the attribute
    auto kern = contract_kernel<NP, MASKED, false, KSTEPS>;
    PhaseScope ps(p, p->phase_radius >= 0 ? IPR_PHASE_RADIUS0 + std::min(p->phase_radius, 23)
                                          : (a.run_if ? IPR_PHASE_FIXUP : IPR_PHASE_CONTRACT));
    if (!configured[p->device & 63]) {
        IPR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
        configured[p->device & 63] = true;
    }
    for (size_t i0 = 0; i0 < tiles.size(); i0 += kMaxTilesPerLaunch) {
        const int n = (int)std::min<size_t>(kMaxTilesPerLaunch, tiles.size() - i0);
        a.tiles.n = n;
        for (int i = 0; i < n; i++) a.tiles.t0[i] = tiles[i0 + i];
        const long long blocks = (long long)a.n_cell_tiles * n;
        kern<<<(unsigned)blocks, kThreads, L::kTotal, p->stream>>>(a);
        IPR_CUDA(cudaGetLastError());
        p->launches++;
    }
    return IPR_OK;
}

// Cressman cell tiles: 8 Morton-consecutive runs of 8 consecutive cells, i.e. compact 2-D patches
// instead of 64 x 1 strips, so that fewer station stages reach any cell of a tile.
int ensure_run_perm(ipr_plan* p, char* err, size_t errlen) {
    if (p->run_perm_ready) return IPR_OK;
    const int n_runs = (int)p->h_run_xy.size();
    std::vector<double> rx(n_runs), ry(n_runs);
    for (int i = 0; i < n_runs; i++) {
        rx[i] = p->h_run_xy[i].x;
        ry[i] = p->h_run_xy[i].y;
    }
    const std::vector<int> run_perm = morton_order(rx.data(), ry.data(), n_runs);
    IPR_CUDA(cudaMemcpyAsync(p->d_run_perm, run_perm.data(), sizeof(int) * run_perm.size(), cudaMemcpyHostToDevice,
                             p->stream));
    IPR_CUDA(cudaStreamSynchronize(p->stream));   // the staging vector goes out of scope
    p->run_perm_ready = true;
    return IPR_OK;
}

// Flatter cell tiles for the store-bound radii: the Morton sort with x divided by `xdiv` groups 2 runs x 4 rows
// (xdiv 4: 128-byte pieces of a layer row per tile), 4 runs x 2 rows (16: 256 B) or 8 runs x 1 row (64: 512 B)
// instead of 1 run x 8 rows (64 B).  Longer pieces are what the HBM write stream wants; squarer tiles see fewer
// stations.  Built on first use, kept for the plan's lifetime.
// The cells of a raster arrive row by row (terra cell order).  When the rows are whole numbers of 8-cell runs, the
// tiles are cut as exact rectangles of `rows` x `wruns` runs, aligned in x, so that every tile writes aligned
// 64 * wruns-byte pieces of a layer row; a Morton sort of the runs (the general case: arbitrary cell sets) cuts its
// blocks at lattice positions unrelated to the run boundaries, and a tile then mixes shorter pieces (measured with
// tools/store_pattern_probe.cu: aligned 128-byte pieces stream at 4.7 TB/s, 64-byte ones at 3.4 TB/s).
// Returns false when the cells are not such a raster.
bool regular_run_perm(const ipr_plan* p, int rows, int wruns, std::vector<int>& perm) {
    const int n_runs = (int)p->h_run_xy.size();
    if (p->n_cells % 8 != 0) return false;
    const int n_real = (int)(p->n_cells / 8);
    int rpr = 1;
    while (rpr < n_real && p->h_run_xy[rpr].y == p->h_run_xy[0].y) rpr++;
    if (rpr >= n_real || n_real % rpr != 0) return false;
    const int nrow = n_real / rpr;
    for (int j = 0; j < nrow; j++) {          // every row: one y, the x pattern of row 0 (first cells of the runs)
        const double y = p->h_run_xy[(size_t)j * rpr].y;
        if (j > 0 && y == p->h_run_xy[(size_t)(j - 1) * rpr].y) return false;
        for (int i = 0; i < rpr; i++)
            if (p->h_run_xy[(size_t)j * rpr + i].y != y || p->h_run_xy[(size_t)j * rpr + i].x != p->h_run_xy[i].x) return false;
    }
    perm.clear();
    perm.reserve(n_runs);
    std::vector<char> used(n_runs, 0);
    for (int rb = 0; rb + rows <= nrow; rb += rows)
        for (int cb = 0; cb + wruns <= rpr; cb += wruns)
            for (int i = 0; i < rows; i++)
                for (int j = 0; j < wruns; j++) {
                    const int run = (rb + i) * rpr + cb + j;
                    perm.push_back(run);
                    used[run] = 1;
                }
    for (int r = 0; r < n_runs; r++)          // ragged right / bottom edges and the padding runs
        if (!used[r]) perm.push_back(r);
    return true;
}

int ensure_run_perm_alt(ipr_plan* p, int xdiv, const int** d_perm, char* err, size_t errlen) {
    auto it = p->run_perm_alt.find(xdiv);
    if (it == p->run_perm_alt.end()) {
        const int n_runs = (int)p->h_run_xy.size();
        const int wruns = xdiv >= 64 ? 8 : xdiv >= 16 ? 4 : xdiv >= 4 ? 2 : 1;
        std::vector<int> perm;
        if ((getenv("IPR_CR_MORTON") && atoi(getenv("IPR_CR_MORTON")) != 0) || !regular_run_perm(p, 8 / wruns, wruns, perm)) {
            if (xdiv <= 1) {
                *d_perm = p->d_run_perm;
                return ensure_run_perm(p, err, errlen);
            }
            std::vector<double> rx(n_runs), ry(n_runs);
            for (int i = 0; i < n_runs; i++) {
                rx[i] = p->h_run_xy[i].x;
                ry[i] = p->h_run_xy[i].y;
            }
            perm = morton_order(rx.data(), ry.data(), n_runs, xdiv);
        }
        int* d = nullptr;
        IPR_CUDA(dev_malloc(&d, sizeof(int) * perm.size()));
        it = p->run_perm_alt.emplace(xdiv, d).first;
        IPR_CUDA(cudaMemcpyAsync(d, perm.data(), sizeof(int) * perm.size(), cudaMemcpyHostToDevice, p->stream));
        IPR_CUDA(cudaStreamSynchronize(p->stream));   // the staging vector goes out of scope
    }
    *d_perm = it->second;
    return IPR_OK;
}

// Makes sure the A-operand cache holds the weights of `mode`/`wp` for the cell tiles starting at
// tile0 (regenerated only when something changed: the geometry is fixed for the plan's lifetime).
int ensure_weights(ipr_plan* p, int mode, const WeightParams& wp, int tile0, int n_tiles, char* err, size_t errlen) {
    const double key = (mode == kIdwEvenP) ? (double)wp.n : wp.a;
    if (p->w_mode == mode && p->w_param == key && p->w_tile0 == tile0) return IPR_OK;
    WeightArgs w{};
    w.cell_xy = p->d_cell;
    w.st_xy = p->d_st;
    w.wfrag = p->d_wfrag;
    w.wsum = p->d_wsum;
    w.wmax = p->d_wmax;
    w.stage_list = p->d_stage_list;
    w.n_active = p->d_n_active;
    w.active_total = p->d_active_total;
    w.run_perm = (mode == kCressman) ? p->d_run_perm : nullptr;
    w.stage_box = (mode == kCressman) ? p->d_stage_box : nullptr;
    // Cressman lists (and counts) half stages
    p->stage_tiles_total += (unsigned long long)n_tiles * p->n_stages * (mode == kCressman ? kSubStages : 1);
    w.cell_tile0 = tile0;
    w.n_cell_tiles = n_tiles;
    w.n_stages = p->n_stages;
    w.n_st = p->n_st;
    w.wp = wp;
    PhaseScope ps(p, IPR_PHASE_WEIGHTS);
    switch (mode) {
        case kIdwP2: weights_kernel<kIdwP2><<<n_tiles, kThreads, 0, p->stream>>>(w); break;
        case kIdwEvenP: weights_kernel<kIdwEvenP><<<n_tiles, kThreads, 0, p->stream>>>(w); break;
        case kIdwGenericP: weights_kernel<kIdwGenericP><<<n_tiles, kThreads, 0, p->stream>>>(w); break;
        case kIdwKriging: weights_kernel<kIdwKriging><<<n_tiles, kThreads, 0, p->stream>>>(w); break;
        default: weights_kernel<kCressman><<<n_tiles, kThreads, 0, p->stream>>>(w); break;
    }
    IPR_CUDA(cudaGetLastError());
    p->launches++;
    p->w_mode = mode;
    p->w_param = key;
    p->w_tile0 = tile0;
    return IPR_OK;
}

// bytes of the digit planes for `tiles` 64-cell tiles: 7 bytes per (cell, padded station)
size_t wq_bytes(const ipr_plan* p, long long tiles) {
    return (size_t)tiles * (kCellsPerCta / kUCells) * p->n_kblocks * kUBBytes;
}

// Lazily sizes the A-operand cache: the whole grid if it fits in half of the free device memory
// (IPR_WCACHE_MB overrides), otherwise the largest chunk of cell tiles that does.
int ensure_wcache(ipr_plan* p, bool need_wq, char* err, size_t errlen) {
    if (p->d_wfrag) {
        // digit planes wanted later than the weights (NA data arrived after an NA-free run): same chunk
        // size; if they do not fit, this plan keeps both GEMMs on the FP64 pipe
        if (need_wq && !p->d_wq &&
            dev_malloc(&p->d_wq, wq_bytes(p, p->w_chunk_tiles)) != cudaSuccess) {
            cudaGetLastError();
            p->d_wq = nullptr;
            p->int8_mask = false;
        }
        return IPR_OK;
    }
    const long long total_tiles = p->c_pad / kCellsPerCta;
    const size_t tile_bytes = (size_t)p->n_stages * kWStageDoubles * sizeof(double);
    size_t free_b = 0, total_b = 0;
    IPR_CUDA(cudaMemGetInfo(&free_b, &total_b));
    size_t budget = free_b / (need_wq ? 3 : 2);   // the digit planes take as much again as the weights
    if (const char* e = getenv("IPR_WCACHE_MB")) budget = (size_t)atoll(e) << 20;
    long long tiles = std::min<long long>(total_tiles, (long long)(budget / tile_bytes));
    if (tiles < 1) {
        set_err(err, errlen, "not enough device memory for the weight cache (%zu bytes per cell tile)", tile_bytes);
        return IPR_ENOMEM;
    }
    cudaError_t e = dev_malloc(&p->d_wfrag, tile_bytes * tiles);
    if (e == cudaSuccess) e = dev_malloc(&p->d_stage_list, sizeof(int) * (size_t)tiles * p->n_stages * kSubStages);
    if (e == cudaSuccess) e = dev_malloc(&p->d_n_active, sizeof(int) * (size_t)tiles);
    if (e == cudaSuccess && need_wq)
        e = dev_malloc(&p->d_wq, wq_bytes(p, tiles));
    if (e != cudaSuccess) {  // leave the plan without a cache, so that a later call starts over
        dev_free(p->d_wfrag);
        dev_free(p->d_stage_list);
        dev_free(p->d_n_active);
        dev_free(p->d_wq);
        p->d_wfrag = nullptr;
        p->d_stage_list = nullptr;
        p->d_n_active = nullptr;
        p->d_wq = nullptr;
        IPR_CUDA(e);
    }
    p->w_chunk_tiles = (int)tiles;
    return IPR_OK;
}

template <int NP>
int launch_contract_denom(ipr_plan* p, ContractArgs& a, const std::vector<int>& tiles, char* err, size_t errlen) {
    using L = SmemLayout<false>;
    static std::atomic<bool> configured[64];   // zero-initialised; racing first calls both set the attribute
    auto kern = contract_kernel<NP, false, true>;
    PhaseScope ps(p, IPR_PHASE_CONTRACT);
    if (!configured[p->device & 63]) {
        IPR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
        configured[p->device & 63] = true;
    }
    for (size_t i0 = 0; i0 < tiles.size(); i0 += kMaxTilesPerLaunch) {
        const int n = (int)std::min<size_t>(kMaxTilesPerLaunch, tiles.size() - i0);
        a.tiles.n = n;
        for (int i = 0; i < n; i++) a.tiles.t0[i] = tiles[i0 + i];
        kern<<<(unsigned)((long long)a.n_cell_tiles * n), kThreads, L::kTotal, p->stream>>>(a);
        IPR_CUDA(cudaGetLastError());
        p->launches++;
    }
    return IPR_OK;
}

// Masked granules of [t0, t1) for the current cell-tile chunk: digit planes of the weights (cached
// like the weights), integer mask contraction (tcgen05 kind::i8) into the output buffer, numerator GEMM
// dividing by it, direct re-evaluation of the listed ill-scaled pairs.  If that list overflows (weights so
// peaked that a missing nearest station leaves a denominator below 2^-10 of the row scale on millions of
// pairs, e.g. large p with many NA), the launch group falls back by itself to the FP64 masked GEMM: the
// same tiles are launched again with both accumulators on the DMMA pipe, and those CTAs exit at once
// unless the device-side overflow flag is set — no host round trip, no failed call.
int run_int8_masked(ipr_plan* p, int mode, const WeightParams& wp, ContractArgs& a, const TileSets& ts, int t0, int t1,
                    char* err, size_t errlen) {
    const double key = (mode == kIdwEvenP) ? (double)wp.n : wp.a;
    const int n_groups = a.n_cell_tiles * (kCellsPerCta / kUCells);
    if (!(p->wq_mode == mode && p->wq_param == key && p->wq_tile0 == a.cell_tile0)) {
        WeightArgs w{};
        w.cell_xy = p->d_cell;
        w.st_xy = p->d_st;
        w.cell_tile0 = a.cell_tile0;
        w.n_cell_tiles = a.n_cell_tiles;
        w.n_stages = p->n_stages;
        w.wp = wp;
        PhaseScope ps(p, IPR_PHASE_SLICE);
        if (mode == kIdwP2)
            slice_weights_kernel<kIdwP2><<<n_groups, 256, 0, p->stream>>>(w, p->d_wmax, p->d_wq, p->n_kblocks, p->n_st);
        else if (mode == kIdwEvenP)
            slice_weights_kernel<kIdwEvenP><<<n_groups, 256, 0, p->stream>>>(w, p->d_wmax, p->d_wq, p->n_kblocks, p->n_st);
        else
            slice_weights_kernel<kIdwGenericP><<<n_groups, 256, 0, p->stream>>>(w, p->d_wmax, p->d_wq, p->n_kblocks,
                                                                                p->n_st);
        IPR_CUDA(cudaGetLastError());
        p->launches++;
        p->wq_mode = mode;
        p->wq_param = key;
        p->wq_tile0 = a.cell_tile0;
    }
    // granules (128 dates) covered by the masked half tiles, with the width needed inside the window
    std::vector<int> gran[9];
    std::vector<int> all_gran;
    {
        std::vector<int> starts;
        for (int np : {4, 2})
            for (int h : ts.m[np]) starts.push_back(h - h % kGran);
        std::sort(starts.begin(), starts.end());
        starts.erase(std::unique(starts.begin(), starts.end()), starts.end());
        for (int b : starts) {
            const int used = std::min(b + kGran, t1) - b;
            gran[used > 96 ? 8 : used > 64 ? 6 : used > 32 ? 4 : 2].push_back(b);
        }
        all_gran = starts;
    }
    // [0] list length, [1] overflow flag of this launch group
    IPR_CUDA(cudaMemsetAsync(p->d_fb_count, 0, 2 * sizeof(int), p->stream));
    static std::atomic<int> n_sm[64];   // zero-initialised; racing first calls write the same value
    if (!n_sm[p->device & 63]) {
        IPR_CUDA(cudaFuncSetAttribute(denom_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kUSmemBytes));
        int v = 0;
        IPR_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, p->device));
        n_sm[p->device & 63] = v;
    }
    DenomArgs d{};
    d.wq = p->d_wq;
    d.mask8 = p->d_mask8;
    d.wmax = p->d_wmax;
    d.date_flag = p->d_flag;
    d.out = a.out;
    d.ld_out = a.ld_out;
    d.n_cells = p->n_cells;
    d.cell_group0 = a.cell_tile0 * (kCellsPerCta / kUCells);
    d.n_cell_groups = n_groups;
    d.n_kblocks = p->n_kblocks;
    d.t_lo = t0;
    d.t_hi = t1;
    d.fb_list = p->d_fb_list;
    d.fb_count = p->d_fb_count;
    d.fb_cap = p->fb_cap;
    int rc;
    for (size_t i0 = 0; i0 < all_gran.size(); i0 += kMaxTilesPerLaunch) {
        PhaseScope ps(p, IPR_PHASE_MASK_GEMM);
        const int n = (int)std::min<size_t>(kMaxTilesPerLaunch, all_gran.size() - i0);
        d.tiles.n = n;
        for (int i = 0; i < n; i++) d.tiles.t0[i] = all_gran[i0 + i];
        // persistent: one CTA per SM walks the (cell group, granule) tiles
        const long long tiles = (long long)n_groups * n;
        const unsigned grid = (unsigned)std::min<long long>(tiles, n_sm[p->device & 63]);
        denom_umma_kernel<<<grid, kUThreads, kUSmemBytes, p->stream>>>(d);
        IPR_CUDA(cudaGetLastError());
        p->launches++;
    }
    if (!gran[8].empty() && (rc = launch_contract_denom<8>(p, a, gran[8], err, errlen))) return rc;
    if (!gran[6].empty() && (rc = launch_contract_denom<6>(p, a, gran[6], err, errlen))) return rc;
    if (!gran[4].empty() && (rc = launch_contract_denom<4>(p, a, gran[4], err, errlen))) return rc;
    if (!gran[2].empty() && (rc = launch_contract_denom<2>(p, a, gran[2], err, errlen))) return rc;
    {
        // direct FP64 re-evaluation of the listed pairs; launched unconditionally (no host wait)
#define IPR_DF_ARGS p->d_fb_list, p->d_fb_count, p->fb_cap, p->d_fb_count + 1, p->d_cell, p->d_st, p->n_st, p->d_perm, \
                    p->d_obs, (long long)p->n_dates, wp, t0, a.out, a.ld_out
        const unsigned nb = 296;
        PhaseScope ps(p, IPR_PHASE_FIXUP);
        if (mode == kIdwP2) idw_direct_fix_kernel<kIdwP2><<<nb, 256, 0, p->stream>>>(IPR_DF_ARGS);
        else if (mode == kIdwEvenP) idw_direct_fix_kernel<kIdwEvenP><<<nb, 256, 0, p->stream>>>(IPR_DF_ARGS);
        else idw_direct_fix_kernel<kIdwGenericP><<<nb, 256, 0, p->stream>>>(IPR_DF_ARGS);
#undef IPR_DF_ARGS
        IPR_CUDA(cudaGetLastError());
        p->launches++;
    }
    // fallback: both GEMMs on the FP64 pipe for the same tiles, executed only if the list overflowed
    a.run_if = p->d_fb_count + 1;
    if (!ts.m[4].empty() && (rc = launch_contract<4, true>(p, a, ts.m[4], err, errlen))) return rc;
    if (!ts.m[2].empty() && (rc = launch_contract<2, true>(p, a, ts.m[2], err, errlen))) return rc;
    a.run_if = nullptr;
    return IPR_OK;
}

// One weight function over the date range: loops over cell-tile chunks of the A-operand cache
int run_contraction(ipr_plan* p, int mode, const WeightParams& wp, bool idw, int t0, int t1, double* d_out,
                    long long ld_out, char* err, size_t errlen) {
    // Kriging fallback: a co-located station weighs 1e20, so its absence on a date can neither be subtracted
    // afterwards (sparse-NA correction) nor sliced into 8-bit digits next to weights 1e-30 times smaller: every
    // granule with an NA goes through the FP64 masked GEMM
    const bool krig = (mode == kIdwKriging);
    TileSets ts;
    build_tiles(p, t0, t1, idw, ts, krig);
    const bool masked_here = idw && (!ts.m[4].empty() || !ts.m[2].empty());
    int rc = ensure_wcache(p, masked_here && p->int8_mask && !krig, err, errlen);
    if (rc) return rc;
    if (mode == kCressman && (rc = ensure_run_perm(p, err, errlen))) return rc;
    const int total_tiles = (int)(p->c_pad / kCellsPerCta);
    for (int tile0 = 0; tile0 < total_tiles; tile0 += p->w_chunk_tiles) {
        const int n_tiles = std::min(p->w_chunk_tiles, total_tiles - tile0);
        if ((rc = ensure_weights(p, mode, wp, tile0, n_tiles, err, errlen))) return rc;
        ContractArgs a{};
        a.wfrag = p->d_wfrag;
        a.wsum = p->d_wsum;
        a.stage_list = p->d_stage_list;
        a.n_active = p->d_n_active;
        a.run_perm = (mode == kCressman) ? p->d_run_perm : nullptr;
        a.z0 = p->d_z0;
        a.mask = p->d_mask;
        a.date_flag = krig ? p->d_kflag : idw ? p->d_flag : nullptr;
        a.date_const = krig ? p->d_kconst : nullptr;
        a.out = d_out;
        a.ld_out = ld_out;
        a.n_cells = p->n_cells;
        a.cell_tile0 = tile0;
        a.n_cell_tiles = n_tiles;
        a.n_stages = p->n_stages;
        a.t_lo = t0;
        a.t_hi = t1;
        // masked granules through the error-free int8 mask contraction (every chunk of cell tiles)
        const bool use_int8 = masked_here && p->int8_mask && p->d_wq != nullptr && !krig;
        if (use_int8 && (rc = run_int8_masked(p, mode, wp, a, ts, t0, t1, err, errlen))) return rc;
        if (mode == kCressman) {   // half stages (8 stations): finer active lists
            constexpr int K = kCressmanKSteps;
            if (!ts.u[8].empty() && (rc = launch_contract<8, false, K>(p, a, ts.u[8], err, errlen))) return rc;
            if (!ts.u[6].empty() && (rc = launch_contract<6, false, K>(p, a, ts.u[6], err, errlen))) return rc;
            if (!ts.u[4].empty() && (rc = launch_contract<4, false, K>(p, a, ts.u[4], err, errlen))) return rc;
            if (!ts.u[2].empty() && (rc = launch_contract<2, false, K>(p, a, ts.u[2], err, errlen))) return rc;
            continue;
        }
        if (!ts.u[8].empty() && (rc = launch_contract<8, false>(p, a, ts.u[8], err, errlen))) return rc;
        if (!ts.u[6].empty() && (rc = launch_contract<6, false>(p, a, ts.u[6], err, errlen))) return rc;
        if (!ts.u[4].empty() && (rc = launch_contract<4, false>(p, a, ts.u[4], err, errlen))) return rc;
        if (!ts.u[2].empty() && (rc = launch_contract<2, false>(p, a, ts.u[2], err, errlen))) return rc;
        if (use_int8) continue;   // the masked granules of this chunk are done
        if (!ts.m[4].empty() && (rc = launch_contract<4, true>(p, a, ts.m[4], err, errlen))) return rc;
        if (!ts.m[2].empty() && (rc = launch_contract<2, true>(p, a, ts.m[2], err, errlen))) return rc;
    }
    return IPR_OK;
}

// Split [t0, t1) into date tiles on granule boundaries (multiples of 128 dates, absolute).  A
// granule with any NA inside the window (when `use_mask`) becomes two masked 64-date tiles, the
// others one 128-date tile; the width actually computed is trimmed to the window in steps of
// 32 dates (NP = 2, 4, 6, 8), so the ragged last granule does not pay for empty columns.
void build_tiles(const ipr_plan* p, int t0, int t1, bool use_mask, TileSets& ts, bool raw_na) {
    const std::vector<uint8_t>& has_na = raw_na ? p->h_has_na_raw : p->h_has_na;
    for (int b = t0 - t0 % kGran; b < t1; b += kGran) {
        const int lo = std::max(b, t0), hi = std::min(b + kGran, t1);
        bool na = false;
        if (use_mask)
            for (int t = lo; t < hi; t++) na |= (has_na[t] != 0);
        if (na) {
            for (int h = b; h < b + kGran; h += 64) {
                if (h >= hi || h + 64 <= lo) continue;
                const int used = std::min(hi, h + 64) - h;      // columns of this half that are needed
                ts.m[used > 32 ? 4 : 2].push_back(h);
            }
        } else {
            const int used = hi - b;
            ts.u[used > 96 ? 8 : used > 64 ? 6 : used > 32 ? 4 : 2].push_back(b);
        }
    }
}

// IPR_CRESSMAN_LEGACY=1: the round-1 path (weights pass per radius + one CTA per (cell tile, date tile)), kept for A/B
bool cressman_legacy() {
    static const bool v = getenv("IPR_CRESSMAN_LEGACY") && atoi(getenv("IPR_CRESSMAN_LEGACY")) != 0;
    return v;
}

// Cressman, one radius: the fused persistent kernel (station lists + weights + contraction), one launch per
// group of at most kMaxTilesPerLaunch date tiles.
template <int NP, int KS>
int launch_cressman(ipr_plan* p, CressmanArgs& c, const std::vector<int>& tiles, char* err, size_t errlen) {
    static std::atomic<bool> configured[64];
    auto kern = cressman_kernel<NP, KS>;
    if (!configured[p->device & 63]) {
        IPR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CrLayout<KS>::kTotal));
        configured[p->device & 63] = true;
    }
    for (size_t i0 = 0; i0 < tiles.size(); i0 += kMaxTilesPerLaunch) {
        const int n = (int)std::min<size_t>(kMaxTilesPerLaunch, tiles.size() - i0);
        c.tiles.n = n;
        for (int i = 0; i < n; i++) c.tiles.t0[i] = tiles[i0 + i];
        IPR_CUDA(cudaMemsetAsync(p->d_cr_counter, 0, sizeof(int), p->stream));
        kern<<<(unsigned)std::min(c.n_cell_tiles, p->cr_grid), kThreads, CrLayout<KS>::kTotal, p->stream>>>(c);
        IPR_CUDA(cudaGetLastError());
        p->launches++;
    }
    return IPR_OK;
}

int run_cressman_radius(ipr_plan* p, double radius_m, int t0, int t1, double* d_out, long long ld_out, char* err,
                        size_t errlen) {
    // Cell-tile shape by the expected length of a tile's station list (n_st x disc / bounding box).  Short lists:
    // the kernel is bound by the 29 GB store stream of the stack, which wants long contiguous pieces of a layer
    // row per tile (measured on cfg 4 at 25 km: 14.1 / 12.2 / 10.1 / 9.0 ms for 64 / 128 / 256 / 512-byte pieces);
    // long lists: bound by the DMMA pipe, which wants square tiles (fewest stations in reach of any cell).
    const double est_list = p->st_bbox_area > 0.0
                                ? std::min((double)p->n_st, p->n_st * 3.141592653589793 * radius_m * radius_m / p->st_bbox_area)
                                : (double)p->n_st;
    int xdiv = est_list < 10.0 ? 64 : est_list < 40.0 ? 4 : 1;
    if (const char* e = getenv("IPR_CR_XDIV")) xdiv = atoi(e);   // 1, 4, 16, 64: override for experiments
    const int* d_perm = nullptr;
    int rc = ensure_run_perm_alt(p, xdiv, &d_perm, err, errlen);
    if (rc) return rc;
    const int list_cap = (int)round_up(p->n_st, kCrListPad);
    if (!p->d_cr_panel) {
        int n_sm = 0;
        IPR_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, p->device));
        p->cr_grid = 2 * n_sm;   // two resident CTAs per SM
        IPR_CUDA(dev_malloc(&p->d_cr_panel, sizeof(double) * (size_t)p->cr_grid * list_cap * kCellsPerCta));
        IPR_CUDA(dev_malloc(&p->d_cr_list, sizeof(int) * (size_t)p->cr_grid * list_cap));
        IPR_CUDA(dev_malloc(&p->d_cr_counter, sizeof(int) * 32));
    }
    TileSets ts;
    build_tiles(p, t0, t1, false, ts);
    std::vector<int> tiles;
    int np = 2;
    for (int w : {2, 4, 6, 8})
        if (!ts.u[w].empty()) {
            np = w;
            tiles.insert(tiles.end(), ts.u[w].begin(), ts.u[w].end());
        }
    std::sort(tiles.begin(), tiles.end());
    CressmanArgs c{};
    c.cell_xy = p->d_cell;
    c.st_xy = p->d_st;
    c.group_box = p->d_stage_box;
    c.run_perm = d_perm;
    c.z0 = p->d_z0;
    c.out = d_out;
    c.ld_out = ld_out;
    c.n_cells = p->n_cells;
    c.n_cell_tiles = (int)(p->c_pad / kCellsPerCta);
    c.n_st = p->n_st;
    c.s_pad = p->s_pad;
    c.n_groups = p->n_stages * kSubStages;
    c.r = radius_m;
    c.r2 = radius_m * radius_m;
    c.t_lo = t0;
    c.t_hi = t1;
    c.panel = p->d_cr_panel;
    c.list = p->d_cr_list;
    c.list_cap = list_cap;
    c.active_total = p->d_active_total;
    c.tile_counter = p->d_cr_counter;
    // per-phase cycle counters of the kernel: only in a build with -DIPR_CR_PROFILE=1 (they cost 16 registers);
    // single device, single thread — a debugging aid, not part of the product path
    static unsigned long long* d_prof = nullptr;
    if (IPR_CR_PROFILE && getenv("IPR_CR_PROF")) {
        if (!d_prof) IPR_CUDA(cudaMalloc(&d_prof, 64));
        IPR_CUDA(cudaMemsetAsync(d_prof, 0, 64, p->stream));
        c.prof = d_prof;
    }
    const size_t n_launch = (tiles.size() + kMaxTilesPerLaunch - 1) / kMaxTilesPerLaunch;
    p->stage_tiles_total += (unsigned long long)c.n_cell_tiles * c.n_groups * n_launch;
    PhaseScope ps(p, IPR_PHASE_RADIUS0 + std::min(std::max(p->phase_radius, 0), 23));
    // the widest tile of the call decides the instantiation; narrower (ragged) tiles mask their stores
    // (16-station stages, KS = 4, were measured too: 57.1 against 56.5 ms at 200 km, 20.5 against 19.7 at 100 km)
    switch (np) {
        case 8: rc = launch_cressman<8, 2>(p, c, tiles, err, errlen); break;
        case 6: rc = launch_cressman<6, 2>(p, c, tiles, err, errlen); break;
        case 4: rc = launch_cressman<4, 2>(p, c, tiles, err, errlen); break;
        default: rc = launch_cressman<2, 2>(p, c, tiles, err, errlen); break;
    }
    if (c.prof && rc == IPR_OK) {
        unsigned long long h[8];
        IPR_CUDA(cudaStreamSynchronize(p->stream));
        IPR_CUDA(cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost));
        const double per = 1.0 / ((double)std::min(c.n_cell_tiles, p->cr_grid) * kConsumerWarps);
        fprintf(stderr, "[cr prof] R=%.0f km, cycles per warp: prepare %.0f  pre-wait %.0f  wait %.0f  mma %.0f  release %.0f  epilogue %.0f  (duty: empty-wait %.0f  gather issue %.0f)\n",
                radius_m / 1000.0, h[0] * per, h[1] * per, h[2] * per, h[3] * per, h[4] * per, h[5] * per, h[6] * per, h[7] * per);
    }
    return rc;
}

int run_post(ipr_plan* p, int t0, int t1, double* d_out, long long ld_out, char* err, size_t errlen) {
    if (p->post_round_mode == 0 && !p->d_cell_keep) return IPR_OK;
    dim3 grid((unsigned)((p->n_cells + 255) / 256), (unsigned)std::min(t1 - t0, 64));
    PhaseScope ps(p, IPR_PHASE_POST);
    post_kernel<<<grid, 256, 0, p->stream>>>(d_out, ld_out, p->n_cells, t1 - t0, p->post_round_mode, p->post_p10,
                                             p->d_cell_keep);
    IPR_CUDA(cudaGetLastError());
    p->launches++;
    return IPR_OK;
}

int check_range(const ipr_plan* p, int t0, int t1, const double* d_out, long long ld_out, char* err, size_t errlen) {
    if (!p) {
        set_err(err, errlen, "plan is NULL");
        return IPR_EINVAL;
    }
    if (!p->obs_ready) {
        set_err(err, errlen, "ipr_plan_set_obs must be called before interpolating");
        return IPR_EINVAL;
    }
    if (t0 < 0 || t1 > p->n_dates || t0 > t1) {
        set_err(err, errlen, "date range [%d, %d) outside [0, %d)", t0, t1, p->n_dates);
        return IPR_EINVAL;
    }
    if (!d_out || ld_out < p->n_cells) {
        set_err(err, errlen, "output pointer is NULL or ld_out < n_cells");
        return IPR_EINVAL;
    }
    return IPR_OK;
}

int scan_zero_pairs(ipr_plan* p, char* err, size_t errlen) {
    IPR_CUDA(cudaMemsetAsync(p->d_counters, 0, sizeof(int), p->stream));
    const int threads = 256;
    const unsigned blocks = (unsigned)((p->n_cells + threads - 1) / threads);
    zero_pairs_kernel<<<blocks, threads, 0, p->stream>>>(p->d_cell, p->n_cells, p->d_st, p->n_st, p->d_perm, p->d_pairs,
                                                         p->pair_capacity, p->d_counters);
    IPR_CUDA(cudaGetLastError());
    p->launches++;
    int n = 0;
    IPR_CUDA(cudaMemcpyAsync(&n, p->d_counters, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    IPR_CUDA(cudaStreamSynchronize(p->stream));
    if (n > p->pair_capacity) {
        set_err(err, errlen, "%d zero-distance (cell, station) pairs exceed the supported %d", n, p->pair_capacity);
        return IPR_EUNSUPPORTED;
    }
    p->zero_pairs = n;
    p->n_fix = 0;
    if (n == 0) return IPR_OK;
    std::vector<int2> h(n);
    IPR_CUDA(cudaMemcpy(h.data(), p->d_pairs, sizeof(int2) * n, cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end(), [](const int2& a, const int2& b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
    std::vector<int> cell, off, st;
    for (int i = 0; i < n; i++) {
        if (i == 0 || h[i].x != h[i - 1].x) {
            cell.push_back(h[i].x);
            off.push_back(i);
        }
        st.push_back(h[i].y);
    }
    off.push_back(n);
    p->n_fix = (int)cell.size();
    dev_free(p->d_fix_cell);
    dev_free(p->d_fix_off);
    dev_free(p->d_fix_st);
    p->d_fix_cell = p->d_fix_off = p->d_fix_st = nullptr;
    IPR_CUDA(dev_malloc(&p->d_fix_cell, sizeof(int) * cell.size()));
    IPR_CUDA(dev_malloc(&p->d_fix_off, sizeof(int) * off.size()));
    IPR_CUDA(dev_malloc(&p->d_fix_st, sizeof(int) * st.size()));
    IPR_CUDA(cudaMemcpy(p->d_fix_cell, cell.data(), sizeof(int) * cell.size(), cudaMemcpyHostToDevice));
    IPR_CUDA(cudaMemcpy(p->d_fix_off, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice));
    IPR_CUDA(cudaMemcpy(p->d_fix_st, st.data(), sizeof(int) * st.size(), cudaMemcpyHostToDevice));
    return IPR_OK;
}

int prepare_obs(ipr_plan* p, char* err, size_t errlen) {
    IPR_CUDA(cudaMemsetAsync(p->d_counters + 1, 0, sizeof(int), p->stream));
    dim3 block(32, 8);
    dim3 grid((unsigned)((p->n_dates + 31) / 32));
    std::unique_ptr<PhaseScope> ps(new PhaseScope(p, IPR_PHASE_PREP));
    prep_obs_kernel<<<grid, block, 0, p->stream>>>(p->d_obs, p->n_dates, p->n_dates, p->n_st, p->d_perm, p->d_z0, p->d_mask,
                                                   p->n_stages, p->d_flag, p->d_has_na, p->d_counters + 1);
    IPR_CUDA(cudaGetLastError());
    p->launches++;
    {
        const int warps_per_block = 8;
        na_lists_kernel<<<(p->n_dates + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, p->stream>>>(
            p->d_obs, p->n_dates, p->n_dates, p->n_st, p->d_perm, p->d_na_n, p->d_na_idx);
        IPR_CUDA(cudaGetLastError());
        p->launches++;
    }
    if (p->int8_mask) {
        const long long total = (long long)p->n_gran * p->n_kblocks * 2 * kGran;
        prep_mask8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, p->stream>>>(
            p->d_obs, p->n_dates, p->n_dates, p->n_st, p->d_perm, p->d_mask8, p->n_kblocks, p->n_gran);
        IPR_CUDA(cudaGetLastError());
        p->launches++;
    }
    ps.reset();
    int bad = 0;
    p->h_has_na.resize(p->n_dates);
    std::vector<int> na_n(p->n_dates);
    IPR_CUDA(cudaMemcpyAsync(p->h_has_na.data(), p->d_has_na, p->n_dates, cudaMemcpyDeviceToHost, p->stream));
    IPR_CUDA(cudaMemcpyAsync(na_n.data(), p->d_na_n, sizeof(int) * p->n_dates, cudaMemcpyDeviceToHost, p->stream));
    IPR_CUDA(cudaMemcpyAsync(&bad, p->d_counters + 1, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    IPR_CUDA(cudaStreamSynchronize(p->stream));
    if (bad) {
        set_err(err, errlen, "obs holds %d infinite value(s); only finite observations or NA are supported", bad);
        return IPR_EUNSUPPORTED;
    }
    // Per 128-date granule: few missing observations (<= 5 % of the entries, <= kNaCap per date) ->
    // unmasked contraction + denominator correction of the affected dates; otherwise masked GEMM.
    // (regenerating a missing weight costs ~10 FP64 instructions against 1 FMA per entry of the
    // masked GEMM, so the break-even is near 10 % NA.)
    p->h_has_na_raw = p->h_has_na;
    p->kflags_ready = false;
    p->h_date_mode.assign(p->n_dates, 0);
    p->any_sparse_fix = false;
    const bool allow_sparse = !(getenv("IPR_SPARSE_NA") && atoi(getenv("IPR_SPARSE_NA")) == 0);
    for (int g0 = 0; g0 < p->n_dates; g0 += kGran) {
        const int g1 = std::min(p->n_dates, g0 + kGran);
        long long total = 0;
        int worst = 0;
        for (int t = g0; t < g1; t++) {
            total += na_n[t];
            worst = std::max(worst, na_n[t]);
        }
        if (total == 0 || !allow_sparse || worst > kNaCap || total * 20 > (long long)(g1 - g0) * p->n_st) continue;
        for (int t = g0; t < g1; t++) {
            p->h_has_na[t] = 0;
            if (na_n[t] > 0) {
                p->h_date_mode[t] = 1;
                p->any_sparse_fix = true;
            }
        }
    }
    IPR_CUDA(cudaMemcpyAsync(p->d_date_mode, p->h_date_mode.data(), p->n_dates, cudaMemcpyHostToDevice, p->stream));
    IPR_CUDA(cudaStreamSynchronize(p->stream));
    p->obs_ready = true;
    return IPR_OK;
}

}  // namespace

extern "C" {

int ipr_version(void) { return 100; }

int ipr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return IPR_ECUDA;
    int ok = 0;
    for (int i = 0; i < n; i++) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10) ok++;
    }
    return ok;
}

double ipr_na_real(void) {
    double v;
    const unsigned long long b = kNaRealBits;
    memcpy(&v, &b, 8);
    return v;
}

void* ipr_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) return nullptr;
    return p;
}
void ipr_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

static int plan_create_impl(ipr_plan** plan, int device, const double* cell_x, const double* cell_y, int64_t n_cells,
                            const double* st_x, const double* st_y, int32_t n_st, int32_t n_dates, char* err,
                            size_t errlen) {
    if (!plan || !cell_x || !cell_y || !st_x || !st_y) {
        set_err(err, errlen, "NULL argument");
        return IPR_EINVAL;
    }
    if (n_cells <= 0 || n_st <= 0 || n_dates <= 0) {
        set_err(err, errlen, "n_cells, n_st and n_dates must be positive");
        return IPR_EINVAL;
    }
    if (n_cells > 2000000000LL) {
        set_err(err, errlen, "n_cells > 2e9 is not supported");
        return IPR_EUNSUPPORTED;
    }
    // device count and compute capability are queried once per process: cudaGetDeviceProperties was traced at
    // 2-25 ms per call (IPR_TRACE), the whole variable part of a 540 ms host-buffer call
    static std::mutex dev_mu;
    static int ndev = -1;
    static int cc[64][2];
    {
        std::lock_guard<std::mutex> lk(dev_mu);
        if (ndev < 0) {
            int n = 0;
            IPR_CUDA(cudaGetDeviceCount(&n));
            for (int i = 0; i < n && i < 64; i++) {
                IPR_CUDA(cudaDeviceGetAttribute(&cc[i][0], cudaDevAttrComputeCapabilityMajor, i));
                IPR_CUDA(cudaDeviceGetAttribute(&cc[i][1], cudaDevAttrComputeCapabilityMinor, i));
            }
            ndev = std::min(n, 64);
        }
    }
    if (device < 0 || device >= ndev) {
        set_err(err, errlen, "device %d not available (%d visible)", device, ndev);
        return IPR_ECUDA;
    }
    if (cc[device][0] != 10) {
        set_err(err, errlen, "device %d is sm_%d%d; this engine is built for sm_100a only (no fallback)", device,
                cc[device][0], cc[device][1]);
        return IPR_EUNSUPPORTED;
    }
    IPR_CUDA(cudaSetDevice(device));
    Trace tr;
    // destroyed on every early return (error codes and exceptions alike)
    std::unique_ptr<ipr_plan, void (*)(ipr_plan*)> holder(new ipr_plan(), ipr_plan_destroy);
    ipr_plan* p = holder.get();
    p->device = device;
    p->n_cells = n_cells;
    p->c_pad = round_up(n_cells, kCellsPerCta);
    p->n_st = n_st;
    p->s_pad = (int)round_up(n_st, kKT);
    p->n_dates = n_dates;
    p->n_gran = (int)((n_dates + kGran - 1) / kGran) + 1;  // +1 spare granule of zero padding
    p->n_stages = p->s_pad / kKT;
    p->z_doubles = (size_t)p->n_gran * p->n_stages * kBlkDoubles;
    p->pair_capacity = std::max(4 * n_st, 65536);
    IPR_CUDA(acquire_stream(&p->stream));
    IPR_CUDA(cudaEventCreate(&p->ev0));
    IPR_CUDA(cudaEventCreate(&p->ev1));
    IPR_CUDA(dev_malloc(&p->d_cell, sizeof(double2) * p->c_pad));
    // padded so that both the 16-station stages and the 32-station integer k-blocks stay in bounds
    const size_t st_alloc = (size_t)std::max<long long>(p->s_pad + kKT, round_up((p->s_pad + kQKb - 1) / kQKb, kQKbPad) * kQKb);
    IPR_CUDA(dev_malloc(&p->d_st, sizeof(double2) * st_alloc));
    IPR_CUDA(dev_malloc(&p->d_obs, sizeof(double) * (size_t)n_dates * n_st));
    IPR_CUDA(dev_malloc(&p->d_z0, sizeof(double) * p->z_doubles));
    IPR_CUDA(dev_malloc(&p->d_mask, sizeof(double) * p->z_doubles));
    IPR_CUDA(dev_malloc(&p->d_flag, (size_t)p->n_gran * kGran));
    IPR_CUDA(dev_malloc(&p->d_has_na, (size_t)p->n_gran * kGran));
    IPR_CUDA(dev_malloc(&p->d_na_n, sizeof(int) * (size_t)n_dates));
    IPR_CUDA(dev_malloc(&p->d_na_idx, sizeof(int) * (size_t)n_dates * kNaCap));
    IPR_CUDA(dev_malloc(&p->d_date_mode, (size_t)n_dates));
    IPR_CUDA(dev_malloc(&p->d_counters, sizeof(int) * 4));
    IPR_CUDA(dev_malloc(&p->d_pairs, sizeof(int2) * p->pair_capacity));
    IPR_CUDA(dev_malloc(&p->d_wsum, sizeof(double) * p->c_pad));
    IPR_CUDA(dev_malloc(&p->d_wmax, sizeof(double) * p->c_pad));
    p->n_kblocks = (int)round_up((p->s_pad + kQKb - 1) / kQKb, kQKbPad);   // padded: whole pipeline stages
    // default on; the error bound of the digit planes is derived for at most 2^16 stations
    p->int8_mask = !(getenv("IPR_MASK_INT8") && atoi(getenv("IPR_MASK_INT8")) == 0) && n_st <= 65536;
    if (p->int8_mask) {
        IPR_CUDA(dev_malloc(&p->d_mask8, (size_t)p->n_gran * p->n_kblocks * kUABytes));
        if (const char* e = getenv("IPR_FB_CAP")) p->fb_cap = std::max(1, atoi(e));   // tests: force the overflow path
        IPR_CUDA(dev_malloc(&p->d_fb_list, sizeof(int2) * (size_t)p->fb_cap));
        IPR_CUDA(dev_malloc(&p->d_fb_count, sizeof(int) * 2));   // [0] list length, [1] overflow flag of the launch group
        IPR_CUDA(cudaMemsetAsync(p->d_fb_count, 0, sizeof(int) * 2, p->stream));
    }
    IPR_CUDA(dev_malloc(&p->d_perm, sizeof(int) * n_st));
    IPR_CUDA(dev_malloc(&p->d_run_perm, sizeof(int) * (size_t)(p->c_pad / 8)));
    IPR_CUDA(dev_malloc(&p->d_stage_box, sizeof(double4) * (size_t)p->n_stages * kSubStages));
    IPR_CUDA(dev_malloc(&p->d_active_total, sizeof(unsigned long long)));
    IPR_CUDA(cudaMemsetAsync(p->d_active_total, 0, sizeof(unsigned long long), p->stream));
    IPR_CUDA(cudaMemsetAsync(p->d_z0, 0, sizeof(double) * p->z_doubles, p->stream));
    IPR_CUDA(cudaMemsetAsync(p->d_mask, 0, sizeof(double) * p->z_doubles, p->stream));
    IPR_CUDA(cudaMemsetAsync(p->d_flag, 0, (size_t)p->n_gran * kGran, p->stream));
    IPR_CUDA(cudaMemsetAsync(p->d_counters, 0, sizeof(int) * 4, p->stream));
    tr.mark("(plan) buffers acquired, memsets queued");
    {
        std::vector<double2> h(p->c_pad);
        for (long long i = 0; i < p->c_pad; i++) {
            const long long k = std::min<long long>(i, n_cells - 1);
            h[i] = make_double2(canon_zero(cell_x[k]), canon_zero(cell_y[k]));
            if (!std::isfinite(h[i].x) || !std::isfinite(h[i].y)) {
                set_err(err, errlen, "cell coordinate %lld is not finite", k);
                return IPR_EINVAL;
            }
        }
        tr.mark("(plan) cell coordinates staged on the host");
        IPR_CUDA(cudaMemcpyAsync(p->d_cell, h.data(), sizeof(double2) * p->c_pad, cudaMemcpyHostToDevice, p->stream));
        // cell tiles for Cressman are 8 Morton-consecutive runs of 8 consecutive cells (ensure_run_perm);
        // the sort costs ~10 ms per 10^6 cells on the host, so it waits for the first Cressman call
        p->h_run_xy.resize((size_t)(p->c_pad / 8));
        for (size_t i = 0; i < p->h_run_xy.size(); i++) p->h_run_xy[i] = h[i * 8];
        IPR_CUDA(cudaStreamSynchronize(p->stream));
        tr.mark("(plan) cell coordinates uploaded (stream synchronised)");
        for (int i = 0; i < n_st; i++) {
            if (!std::isfinite(st_x[i]) || !std::isfinite(st_y[i])) {
                set_err(err, errlen, "station coordinate %d is not finite", i);
                return IPR_EINVAL;
            }
        }
        {
            double x0 = st_x[0], x1 = st_x[0], y0 = st_y[0], y1 = st_y[0];
            for (int i = 1; i < n_st; i++) {
                x0 = std::min(x0, st_x[i]); x1 = std::max(x1, st_x[i]);
                y0 = std::min(y0, st_y[i]); y1 = std::max(y1, st_y[i]);
            }
            p->st_bbox_area = (x1 - x0) * (y1 - y0);
        }
        // stations in Morton order: a pipeline stage (kKT consecutive stations) is then a compact
        // blob, so that for Cressman most (cell tile, stage) blocks lie wholly beyond the radius and
        // are skipped.  The permutation only changes the order of summation.
        std::vector<int> perm = morton_order(st_x, st_y, n_st);
        std::vector<double2> hs(st_alloc, make_double2(kPadCoord, kPadCoord));
        for (int i = 0; i < n_st; i++) hs[i] = make_double2(canon_zero(st_x[perm[i]]), canon_zero(st_y[perm[i]]));
        IPR_CUDA(cudaMemcpyAsync(p->d_perm, perm.data(), sizeof(int) * n_st, cudaMemcpyHostToDevice, p->stream));
        // bounding box of every half stage (the granularity of Cressman's active lists)
        constexpr int kHalf = kKT / kSubStages;
        std::vector<double4> boxes((size_t)p->n_stages * kSubStages);
        for (size_t hb = 0; hb < boxes.size(); hb++) {
            double4 b = make_double4(hs[hb * kHalf].x, hs[hb * kHalf].y, hs[hb * kHalf].x, hs[hb * kHalf].y);
            for (int k = 1; k < kHalf; k++) {
                const double2 c = hs[hb * kHalf + k];
                b.x = std::min(b.x, c.x);
                b.y = std::min(b.y, c.y);
                b.z = std::max(b.z, c.x);
                b.w = std::max(b.w, c.y);
            }
            boxes[hb] = b;
        }
        IPR_CUDA(cudaMemcpyAsync(p->d_stage_box, boxes.data(), sizeof(double4) * boxes.size(), cudaMemcpyHostToDevice,
                                  p->stream));
        IPR_CUDA(cudaStreamSynchronize(p->stream));
        IPR_CUDA(cudaMemcpyAsync(p->d_st, hs.data(), sizeof(double2) * hs.size(), cudaMemcpyHostToDevice, p->stream));
        IPR_CUDA(cudaStreamSynchronize(p->stream));  // staging vectors go out of scope
    }
    tr.mark("(plan) stations sorted and uploaded");
    const int rc = scan_zero_pairs(p, err, errlen);
    if (rc != IPR_OK) return rc;
    tr.mark("(plan) zero-distance scan done");
    *plan = holder.release();
    return IPR_OK;
}

int ipr_plan_create(ipr_plan** plan, int device, const double* cell_x, const double* cell_y, int64_t n_cells,
                    const double* st_x, const double* st_y, int32_t n_st, int32_t n_dates, char* err, size_t errlen) {
    return no_throw(err, errlen, [&] {
        return plan_create_impl(plan, device, cell_x, cell_y, n_cells, st_x, st_y, n_st, n_dates, err, errlen);
    });
}

void ipr_plan_destroy(ipr_plan* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    dev_free(p->d_cell);
    dev_free(p->d_st);
    dev_free(p->d_obs);
    dev_free(p->d_z0);
    dev_free(p->d_mask);
    dev_free(p->d_flag);
    dev_free(p->d_kflag);
    dev_free(p->d_kconst);
    dev_free(p->d_has_na);
    dev_free(p->d_na_n);
    dev_free(p->d_na_idx);
    dev_free(p->d_date_mode);
    dev_free(p->d_counters);
    dev_free(p->d_pairs);
    dev_free(p->d_fix_cell);
    dev_free(p->d_fix_off);
    dev_free(p->d_fix_st);
    dev_free(p->d_wfrag);
    dev_free(p->d_wsum);
    dev_free(p->d_wmax);
    dev_free(p->d_wq);
    dev_free(p->d_mask8);
    dev_free(p->d_fb_list);
    dev_free(p->d_fb_count);
    dev_free(p->d_stage_list);
    dev_free(p->d_n_active);
    dev_free(p->d_perm);
    dev_free(p->d_run_perm);
    for (auto& kv : p->run_perm_alt) dev_free(kv.second);
    dev_free(p->d_stage_box);
    dev_free(p->d_cell_keep);
    dev_free(p->d_active_total);
    dev_free(p->d_cr_panel);
    dev_free(p->d_cr_list);
    dev_free(p->d_cr_counter);
    for (auto& sp : p->spans) {
        cudaEventDestroy(sp.e0);
        cudaEventDestroy(sp.e1);
    }
    for (auto e : p->ev_pool) cudaEventDestroy(e);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->stream) release_stream(p->stream);   // synchronised above
    delete p;
}

// obs on host with an explicit leading dimension (column stride, >= n_dates): used by the
// multi-GPU scheduler to hand each device its date block of every station's series
static int plan_set_obs_ld(ipr_plan* p, const double* obs, long long ld, char* err, size_t errlen) {
    if (!p || !obs) {
        set_err(err, errlen, "NULL argument");
        return IPR_EINVAL;
    }
    IPR_CUDA(cudaSetDevice(p->device));
    IPR_CUDA(cudaMemcpy2DAsync(p->d_obs, sizeof(double) * p->n_dates, obs, sizeof(double) * ld,
                               sizeof(double) * p->n_dates, p->n_st, cudaMemcpyHostToDevice, p->stream));
    return prepare_obs(p, err, errlen);
}

int ipr_plan_set_obs(ipr_plan* p, const double* obs, char* err, size_t errlen) {
    return no_throw(err, errlen, [&] { return plan_set_obs_ld(p, obs, p ? p->n_dates : 0, err, errlen); });
}

int ipr_plan_set_obs_device(ipr_plan* p, const double* d_obs, char* err, size_t errlen) {
    if (!p || !d_obs) {
        set_err(err, errlen, "NULL argument");
        return IPR_EINVAL;
    }
    IPR_CUDA(cudaSetDevice(p->device));
    IPR_CUDA(cudaMemcpyAsync(p->d_obs, d_obs, sizeof(double) * (size_t)p->n_dates * p->n_st, cudaMemcpyDeviceToDevice,
                             p->stream));
    return no_throw(err, errlen, [&] { return prepare_obs(p, err, errlen); });
}

static int plan_idw_impl(ipr_plan* p, double pw, int32_t t0, int32_t t1, double* d_out, int64_t ld_out, char* err,
                         size_t errlen) {
    int rc = check_range(p, t0, t1, d_out, ld_out, err, errlen);
    if (rc) return rc;
    if (!(pw > 0) || !std::isfinite(pw)) {
        set_err(err, errlen, "p must be a positive finite number");
        return IPR_EINVAL;
    }
    if (t0 == t1) return IPR_OK;
    IPR_CUDA(cudaSetDevice(p->device));
    WeightParams wp{};
    int mode;
    const double half = pw / 2.0;
    if (pw == 2.0) {
        mode = kIdwP2;
    } else if (half == std::floor(half) && half <= 16.0) {
        mode = kIdwEvenP;
        wp.n = (int)half;
    } else {
        mode = kIdwGenericP;
        wp.a = -half;
    }
    if ((rc = run_contraction(p, mode, wp, true, t0, t1, d_out, ld_out, err, errlen))) return rc;
    bool fix_here = false;
    if (p->any_sparse_fix)
        for (int t = t0; t < t1 && !fix_here; t++) fix_here = (p->h_date_mode[t] == 1);
    PhaseScope ps_fix(p, IPR_PHASE_FIXUP);
    if (fix_here) {
        const int threads = 256;
        for (int tb = t0; tb < t1; tb += 32768) {
            const int te = std::min(t1, tb + 32768);
            dim3 grid((unsigned)((p->n_cells + threads - 1) / threads), (unsigned)(te - tb));
            double* o = d_out + (size_t)(tb - t0) * ld_out;
#define IPR_FIX_ARGS p->d_cell, p->d_st, p->n_st, p->d_wsum, p->n_cells, p->d_na_n, p->d_na_idx, p->d_date_mode, \
                     p->d_flag, tb, te, wp, o, ld_out
            if (mode == kIdwP2) idw_sparse_na_fix_kernel<kIdwP2><<<grid, threads, 0, p->stream>>>(IPR_FIX_ARGS);
            else if (mode == kIdwEvenP) idw_sparse_na_fix_kernel<kIdwEvenP><<<grid, threads, 0, p->stream>>>(IPR_FIX_ARGS);
            else idw_sparse_na_fix_kernel<kIdwGenericP><<<grid, threads, 0, p->stream>>>(IPR_FIX_ARGS);
#undef IPR_FIX_ARGS
            IPR_CUDA(cudaGetLastError());
            p->launches++;
        }
    }
    for (int f0 = 0; f0 < p->n_fix; f0 += 65535) {   // gridDim.y <= 65535
        dim3 grid((unsigned)((t1 - t0 + 127) / 128), (unsigned)std::min(65535, p->n_fix - f0));
        idw_zero_fix_kernel<<<grid, 128, 0, p->stream>>>(p->d_fix_cell, p->d_fix_off, p->d_fix_st, f0, p->n_fix, p->d_obs,
                                                         p->n_dates, p->d_flag, t0, t1, d_out, ld_out);
        IPR_CUDA(cudaGetLastError());
        p->launches++;
    }
    ps_fix.finish();
    return run_post(p, t0, t1, d_out, ld_out, err, errlen);
}

int ipr_plan_idw(ipr_plan* p, double pw, int32_t t0, int32_t t1, double* d_out, int64_t ld_out, char* err,
                 size_t errlen) {
    return no_throw(err, errlen, [&] { return plan_idw_impl(p, pw, t0, t1, d_out, ld_out, err, errlen); });
}

// Kriging_Ordinary's IDW fallback over [t0, t1): R/Kriging_Ordinary.R:406-423 (constant layers) and :465-476
// (the weighted mean with d == 0 -> 1e-10).  Same contraction as IDW(); no zero-distance running mean, no
// early-out on all-zero dates, never NA.
static int plan_idw_kriging_impl(ipr_plan* p, int32_t t0, int32_t t1, double* d_out, int64_t ld_out, char* err,
                                 size_t errlen) {
    int rc = check_range(p, t0, t1, d_out, ld_out, err, errlen);
    if (rc) return rc;
    if (t0 == t1) return IPR_OK;
    IPR_CUDA(cudaSetDevice(p->device));
    if (!p->kflags_ready) {
        if (!p->d_kflag) {
            IPR_CUDA(dev_malloc(&p->d_kflag, (size_t)p->n_gran * kGran));
            IPR_CUDA(dev_malloc(&p->d_kconst, sizeof(double) * (size_t)p->n_gran * kGran));
        }
        PhaseScope ps(p, IPR_PHASE_PREP);
        kriging_flags_kernel<<<(p->n_dates + 7) / 8, 256, 0, p->stream>>>(p->d_obs, p->n_dates, p->n_dates, p->n_st,
                                                                         p->d_kflag, p->d_kconst);
        IPR_CUDA(cudaGetLastError());
        p->launches++;
        p->kflags_ready = true;
    }
    WeightParams wp{};
    if ((rc = run_contraction(p, kIdwKriging, wp, true, t0, t1, d_out, ld_out, err, errlen))) return rc;
    return run_post(p, t0, t1, d_out, ld_out, err, errlen);
}

int ipr_plan_idw_kriging(ipr_plan* p, int32_t t0, int32_t t1, double* d_out, int64_t ld_out, char* err, size_t errlen) {
    return no_throw(err, errlen, [&] { return plan_idw_kriging_impl(p, t0, t1, d_out, ld_out, err, errlen); });
}

// Ordinary kriging over [t0, t1) (R/Kriging_Ordinary.R:405-508).  model[t] (host arrays indexed by absolute date):
// 0 = the date falls back to IDW (flat variogram, singular or ill-conditioned Gamma: decided in R, :428-462),
// 1..4 = exponential / spherical / gaussian / linear with nugget[t], sill[t], range[t].  The fallback pass writes
// every date of the range (IDW fallback + the constant layers of the prelude); the kriged dates are then
// overwritten: one (n+1)^2 solve per date, then the cell x station variogram pass.
static int plan_kriging_impl(ipr_plan* p, const int32_t* model, const double* nugget, const double* sill,
                             const double* range, int32_t t0, int32_t t1, double* d_out, int64_t ld_out, char* err,
                             size_t errlen) {
    if (!model || !nugget || !sill || !range) {
        set_err(err, errlen, "NULL argument");
        return IPR_EINVAL;
    }
    int rc = check_range(p, t0, t1, d_out, ld_out, err, errlen);
    if (rc) return rc;
    if (t0 == t1) return IPR_OK;
    // the post pass (rounding / mask) must run once, after the kriged dates are in place
    const int post_mode = p->post_round_mode;
    uint8_t* const post_keep = p->d_cell_keep;
    p->post_round_mode = 0;
    p->d_cell_keep = nullptr;
    rc = plan_idw_kriging_impl(p, t0, t1, d_out, ld_out, err, errlen);
    p->post_round_mode = post_mode;
    p->d_cell_keep = post_keep;
    if (rc) return rc;
    std::vector<int> list;
    for (int t = t0; t < t1; t++) {
        if (model[t] == 0) continue;
        if (model[t] < 1 || model[t] > 4 || !(range[t] > 0.0) || !std::isfinite(nugget[t]) || !std::isfinite(sill[t]) ||
            !std::isfinite(range[t])) {
            set_err(err, errlen, "date %d: variogram model must be 0..4 with finite nugget / sill and range > 0", t);
            return IPR_EINVAL;
        }
        list.push_back(t);
    }
    if (!list.empty()) {
        const int n1 = p->n_st + 1;
        int n_sm = 0;
        IPR_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, p->device));
        const size_t per_cta = sizeof(double) * (size_t)n1 * (n1 + 1);
        const int grid = (int)std::max<size_t>(1, std::min<size_t>({list.size(), (size_t)n_sm, (size_t)(8ull << 30) / per_cta}));
        const size_t n_list = list.size();
        double *d_scratch = nullptr, *d_alpha = nullptr, *d_par = nullptr, *d_fb = nullptr;
        int *d_list = nullptr, *d_avail = nullptr, *d_meta = nullptr, *d_model = nullptr;
        struct Free {
            std::vector<void*> v;
            ~Free() { for (void* q : v) dev_free(q); }
        } fr;
        auto grab = [&](auto** ptr, size_t bytes) -> cudaError_t {
            const cudaError_t e = dev_malloc(ptr, bytes);
            if (e == cudaSuccess) fr.v.push_back(*ptr);
            return e;
        };
        IPR_CUDA(grab(&d_scratch, per_cta * grid));
        IPR_CUDA(grab(&d_alpha, sizeof(double) * n_list * n1));
        IPR_CUDA(grab(&d_avail, sizeof(int) * n_list * p->n_st));
        IPR_CUDA(grab(&d_meta, sizeof(int) * n_list * 4));
        IPR_CUDA(grab(&d_fb, sizeof(double) * n_list));
        IPR_CUDA(grab(&d_list, sizeof(int) * n_list));
        IPR_CUDA(grab(&d_par, sizeof(double) * 3 * p->n_dates));
        IPR_CUDA(grab(&d_model, sizeof(int) * p->n_dates));
        IPR_CUDA(cudaMemcpyAsync(d_list, list.data(), sizeof(int) * n_list, cudaMemcpyHostToDevice, p->stream));
        IPR_CUDA(cudaMemcpyAsync(d_par, nugget, sizeof(double) * p->n_dates, cudaMemcpyHostToDevice, p->stream));
        IPR_CUDA(cudaMemcpyAsync(d_par + p->n_dates, sill, sizeof(double) * p->n_dates, cudaMemcpyHostToDevice, p->stream));
        IPR_CUDA(cudaMemcpyAsync(d_par + 2 * (size_t)p->n_dates, range, sizeof(double) * p->n_dates, cudaMemcpyHostToDevice,
                                 p->stream));
        IPR_CUDA(cudaMemcpyAsync(d_model, model, sizeof(int) * p->n_dates, cudaMemcpyHostToDevice, p->stream));
        IPR_CUDA(cudaMemsetAsync(d_meta, 0, sizeof(int) * n_list * 4, p->stream));
        KrigingArgs k{};
        k.dates = d_list;
        k.n_list = (int)n_list;
        k.nugget = d_par;
        k.sill = d_par + p->n_dates;
        k.range = d_par + 2 * (size_t)p->n_dates;
        k.model = d_model;
        k.kflag = p->d_kflag;
        k.obs = p->d_obs;
        k.ld_obs = p->n_dates;
        k.perm = p->d_perm;
        k.st_xy = p->d_st;
        k.n_st = p->n_st;
        k.scratch = d_scratch;
        k.alpha = d_alpha;
        k.avail = d_avail;
        k.meta = d_meta;
        k.fallback_value = d_fb;
        {
            PhaseScope ps(p, IPR_PHASE_FIXUP);
            kriging_solve_kernel<<<grid, kKrigThreads, 0, p->stream>>>(k);
            IPR_CUDA(cudaGetLastError());
            p->launches++;
        }
        {
            PhaseScope ps(p, IPR_PHASE_CONTRACT);
            for (size_t l0 = 0; l0 < n_list; l0 += 32768) {   // gridDim.y <= 65535
                KrigingArgs kk = k;
                kk.dates = d_list + l0;
                kk.alpha = d_alpha + l0 * n1;
                kk.avail = d_avail + l0 * p->n_st;
                kk.meta = d_meta + l0 * 4;
                kk.fallback_value = d_fb + l0;
                dim3 g((unsigned)((p->n_cells + 255) / 256), (unsigned)std::min<size_t>(32768, n_list - l0));
                kriging_predict_kernel<<<g, 256, 0, p->stream>>>(kk, p->d_cell, p->n_cells, t0, d_out, ld_out);
                IPR_CUDA(cudaGetLastError());
                p->launches++;
            }
        }
        // the scratch goes back to the pool below; the host arrays were copied from pageable memory
        IPR_CUDA(cudaStreamSynchronize(p->stream));
    }
    return run_post(p, t0, t1, d_out, ld_out, err, errlen);
}

int ipr_plan_kriging(ipr_plan* p, const int32_t* model, const double* nugget, const double* sill, const double* range,
                     int32_t t0, int32_t t1, double* d_out, int64_t ld_out, char* err, size_t errlen) {
    return no_throw(err, errlen,
                    [&] { return plan_kriging_impl(p, model, nugget, sill, range, t0, t1, d_out, ld_out, err, errlen); });
}

int ipr_plan_set_post(ipr_plan* p, int32_t round_mode, int32_t n_round, const uint8_t* cell_keep, char* err,
                      size_t errlen) {
    if (!p) {
        set_err(err, errlen, "plan is NULL");
        return IPR_EINVAL;
    }
    if (round_mode < 0 || round_mode > 2 || (round_mode != 0 && (n_round < 0 || n_round > 15))) {
        set_err(err, errlen, "round_mode must be 0, 1 or 2 and n_round in [0, 15]");
        return IPR_EINVAL;
    }
    IPR_CUDA(cudaSetDevice(p->device));
    p->post_round_mode = round_mode;
    p->post_p10 = std::pow(10.0, (double)n_round);
    if (cell_keep) {
        if (!p->d_cell_keep) IPR_CUDA(dev_malloc(&p->d_cell_keep, (size_t)p->n_cells));
        IPR_CUDA(cudaMemcpyAsync(p->d_cell_keep, cell_keep, (size_t)p->n_cells, cudaMemcpyHostToDevice, p->stream));
        IPR_CUDA(cudaStreamSynchronize(p->stream));
    } else if (p->d_cell_keep) {
        dev_free(p->d_cell_keep);
        p->d_cell_keep = nullptr;
    }
    return IPR_OK;
}

static int plan_cressman_impl(ipr_plan* p, const double* radii_m, int32_t n_radii, int32_t t0, int32_t t1,
                              double* d_out, int64_t ld_out, int64_t stack_stride, char* err, size_t errlen) {
    int rc = check_range(p, t0, t1, d_out, ld_out, err, errlen);
    if (rc) return rc;
    if (!radii_m || n_radii <= 0) {
        set_err(err, errlen, "radii_m must hold at least one radius");
        return IPR_EINVAL;
    }
    for (int i = 0; i < n_radii; i++) {
        if (!(radii_m[i] >= 0) || !std::isfinite(radii_m[i])) {
            set_err(err, errlen, "radius %d is not a finite non-negative number", i);
            return IPR_EINVAL;
        }
    }
    if (t0 == t1) return IPR_OK;
    IPR_CUDA(cudaSetDevice(p->device));
    for (int ri = 0; ri < n_radii; ri++) {
        WeightParams wp{};
        wp.a = radii_m[ri] * radii_m[ri];
        p->phase_radius = ri;
        if (cressman_legacy())
            rc = run_contraction(p, kCressman, wp, false, t0, t1, d_out + (size_t)ri * stack_stride, ld_out, err, errlen);
        else
            rc = run_cressman_radius(p, radii_m[ri], t0, t1, d_out + (size_t)ri * stack_stride, ld_out, err, errlen);
        p->phase_radius = -1;
        if (rc) return rc;
        if ((rc = run_post(p, t0, t1, d_out + (size_t)ri * stack_stride, ld_out, err, errlen))) return rc;
    }
    return IPR_OK;
}

int ipr_plan_cressman(ipr_plan* p, const double* radii_m, int32_t n_radii, int32_t t0, int32_t t1, double* d_out,
                      int64_t ld_out, int64_t stack_stride, char* err, size_t errlen) {
    return no_throw(err, errlen, [&] {
        return plan_cressman_impl(p, radii_m, n_radii, t0, t1, d_out, ld_out, stack_stride, err, errlen);
    });
}

int ipr_plan_stage_blocks(ipr_plan* p, uint64_t* executed, uint64_t* dense) {
    if (!p || !executed || !dense) return IPR_EINVAL;
    cudaSetDevice(p->device);
    unsigned long long v = 0;
    if (cudaMemcpyAsync(&v, p->d_active_total, sizeof v, cudaMemcpyDeviceToHost, p->stream) != cudaSuccess ||
        cudaStreamSynchronize(p->stream) != cudaSuccess)
        return IPR_ECUDA;
    *executed = v;
    *dense = p->stage_tiles_total;
    return IPR_OK;
}

void ipr_release_cached_memory(void) {
    DevicePool& P = pool();
    std::lock_guard<std::mutex> lk(P.mu);
    P.drain_locked();
}

void ipr_plan_reset_cache(ipr_plan* p) {
    if (p) {
        p->w_mode = -1;
        p->wq_mode = -1;
    }
}

int ipr_plan_sync(ipr_plan* p) {
    if (!p) return IPR_EINVAL;
    cudaSetDevice(p->device);
    if (cudaStreamSynchronize(p->stream) != cudaSuccess) return IPR_ECUDA;
    return IPR_OK;
}
int ipr_plan_timer_start(ipr_plan* p) {
    if (!p) return IPR_EINVAL;
    cudaSetDevice(p->device);
    return cudaEventRecord(p->ev0, p->stream) == cudaSuccess ? IPR_OK : IPR_ECUDA;
}
int ipr_plan_timer_stop(ipr_plan* p, float* ms) {
    if (!p || !ms) return IPR_EINVAL;
    cudaSetDevice(p->device);
    if (cudaEventRecord(p->ev1, p->stream) != cudaSuccess) return IPR_ECUDA;
    if (cudaEventSynchronize(p->ev1) != cudaSuccess) return IPR_ECUDA;
    return cudaEventElapsedTime(ms, p->ev0, p->ev1) == cudaSuccess ? IPR_OK : IPR_ECUDA;
}
int ipr_plan_profile(ipr_plan* p, int enable) {
    if (!p) return IPR_EINVAL;
    p->prof_on = enable != 0;
    return IPR_OK;
}
int ipr_plan_profile_read(ipr_plan* p, double* ms, int n) {
    if (!p || !ms || n < IPR_N_PHASES) return IPR_EINVAL;
    cudaSetDevice(p->device);
    if (cudaStreamSynchronize(p->stream) != cudaSuccess) return IPR_ECUDA;
    for (int i = 0; i < n; i++) ms[i] = 0.0;
    for (auto& sp : p->spans) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, sp.e0, sp.e1) == cudaSuccess && sp.id >= 0 && sp.id < IPR_N_PHASES) ms[sp.id] += t;
        p->ev_pool.push_back(sp.e0);
        p->ev_pool.push_back(sp.e1);
    }
    p->spans.clear();
    return IPR_OK;
}
int64_t ipr_plan_launch_count(const ipr_plan* p) { return p ? p->launches : 0; }
int64_t ipr_plan_zero_pairs(const ipr_plan* p) { return p ? p->zero_pairs : 0; }
void* ipr_plan_stream(const ipr_plan* p) { return p ? (void*)p->stream : nullptr; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// One-shot host-buffer entry points: multi-GPU scheduler.
// Each device gets a contiguous range of cells [c0, c1) on 64-cell tile boundaries (no inter-GPU
// exchange); inside a device the dates are processed in 128-date chunks through a ring
// of device buffers so that the D2H copy of chunk i overlaps the kernels of chunk i+1.
// ---------------------------------------------------------------------------------------------
namespace {

struct HostJob {
    const double *cell_x, *cell_y, *st_x, *st_y, *obs;
    long long n_cells;
    int n_st, n_dates;
    bool cressman;
    double p;
    const double* radii;
    int n_radii;
    double* out;
    bool out_pinned = false;  // caller's output is page-locked: D2H lands in it directly
    int round_mode = 0, n_round = 0;
    const uint8_t* cell_keep = nullptr;
    bool kriging = false;                         // IDW fallback of Kriging_Ordinary instead of IDW()
    const int32_t* kr_model = nullptr;            // with kriging: per-date variogram (ipr_kriging_f64), else fallback only
    const double *kr_nugget = nullptr, *kr_sill = nullptr, *kr_range = nullptr;
    int n_job_devices = 1;                        // devices sharing the host's cores / memory bandwidth in this job
    // asynchronous jobs (ipr_job_*): cooperative cancellation and progress in work items (date chunks)
    std::atomic<int>* cancel = nullptr;
    std::atomic<long long>* done = nullptr;
};

constexpr int kRing = 3;

// Work items of one device: (stack, date chunk), stack outermost so that the A-operand cache is generated
// once per radius.  A chunk is one 128-date granule (1 GB per 10^6 cells): D2H starts after the first
// granule and only the last granule's copy is left when the kernels finish.
struct WorkItem { int stack, t0, t1; };
std::vector<WorkItem> make_items(int n_stacks, int nd) {
    std::vector<WorkItem> items;
    for (int s = 0; s < n_stacks; s++)
        for (int t = 0; t < nd; t += kGran) {
            const int te = std::min(nd, t + kGran);
            // the very first granule goes out as 32 + 96 dates: the copy engine starts after a quarter of
            // a granule's kernels instead of a whole one (nothing else can overlap with that wait)
            if (s == 0 && t == 0 && te > 32) {
                items.push_back({s, 0, 32});
                items.push_back({s, 32, te});
            } else {
                items.push_back({s, t, te});
            }
        }
    return items;
}

// ---------------------------------------------------------------------------------------------
// Host topology.  The D2H stream of a multi-GPU job is bound by the HOST side (round 1: 8 x 29 GB in
// 2.6 s = 91 GB/s for the box against 57 GB/s for one GPU alone), so the threads that touch a device's
// share of the output — its scheduler thread and the memcpy fan-out of the pageable path — run on the
// CPUs sysfs reports as local to that device, and pinned memory they allocate is first touched there.
// IPR_NUMA=0 disables all of it.  No libnuma dependency: sysfs + sched_setaffinity + set_mempolicy.
// ---------------------------------------------------------------------------------------------
bool numa_enabled() {
    static const bool on = !(getenv("IPR_NUMA") && atoi(getenv("IPR_NUMA")) == 0);
    return on;
}

std::vector<int> parse_cpulist(const char* s) {
    std::vector<int> cpus;
    while (*s) {
        char* end = nullptr;
        long lo = strtol(s, &end, 10);
        if (end == s) break;
        long hi = lo;
        if (*end == '-') {
            s = end + 1;
            hi = strtol(s, &end, 10);
        }
        for (long c = lo; c <= hi && c < 4096; c++) cpus.push_back((int)c);
        s = (*end == ',') ? end + 1 : end;
        if (*end != ',') break;
    }
    return cpus;
}

// CPUs local to a CUDA device that this process may run on; empty when unknown or when they are all the
// CPUs anyway (single-node hosts: nothing to bind)
std::vector<int> device_local_cpus_uncached(int device);
std::vector<int> device_local_cpus(int device) {
    static std::mutex mu;
    static std::map<int, std::vector<int>> cache;   // sysfs + cudaDeviceGetPCIBusId once per device
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(device);
    if (it == cache.end()) it = cache.emplace(device, device_local_cpus_uncached(device)).first;
    return it->second;
}
std::vector<int> device_local_cpus_uncached(int device) {
    std::vector<int> out;
    if (!numa_enabled()) return out;
    char bdf[32] = "";
    if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, device) != cudaSuccess) {
        cudaGetLastError();
        return out;
    }
    for (char* c = bdf; *c; c++) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bdf);
    FILE* f = fopen(path, "r");
    if (!f) return out;
    char buf[4096] = "";
    const bool ok = fgets(buf, sizeof buf, f) != nullptr;
    fclose(f);
    if (!ok) return out;
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return out;
    for (int c : parse_cpulist(buf))
        if (c < CPU_SETSIZE && CPU_ISSET(c, &allowed)) out.push_back(c);
    if ((int)out.size() == CPU_COUNT(&allowed)) out.clear();
    return out;
}

void bind_this_thread(const std::vector<int>& cpus) {
    if (cpus.empty()) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c : cpus)
        if (c < CPU_SETSIZE) CPU_SET(c, &set);
    pthread_setaffinity_np(pthread_self(), sizeof set, &set);   // best effort
}

// RAII: restores the calling thread's affinity mask (the one-device path runs on the caller's thread)
struct AffinityScope {
    cpu_set_t saved;
    bool active = false;
    explicit AffinityScope(const std::vector<int>& cpus) {
        if (cpus.empty() || pthread_getaffinity_np(pthread_self(), sizeof saved, &saved) != 0) return;
        active = true;
        bind_this_thread(cpus);
    }
    ~AffinityScope() {
        if (active) pthread_setaffinity_np(pthread_self(), sizeof saved, &saved);
    }
};

// ---------------------------------------------------------------------------------------------
// Pageable destination (what R allocates): D2H goes through a small ring of pinned staging slots
// and a drain thread fans every slot out to the caller's memory with parallel memcpy.  Pinning a
// fresh 29 GB matrix with cudaHostRegister instead was measured at 7.2 s (page faults taken one by
// one inside the driver); first-touching it from many threads is what makes this path fast.
// ---------------------------------------------------------------------------------------------
// Staging slot -> the caller's pageable matrix with non-temporal stores.  An ordinary store to a line that is not
// in cache first READS the line (read for ownership): with the GPU's DMA writing the next slot at PCIe rate at
// the same time, that fourth stream of memory traffic is what held the pageable path at 720-820 ms per 29.2 GB
// against 550 ms for a pinned result.  glibc's memcpy only switches to streaming stores above a threshold tied to
// the shared cache size (hundreds of MB on these hosts), far above the 16 MiB a drain thread copies per slot.
// IPR_NT_COPY=0 restores memcpy.
inline bool nt_copy_enabled() {
    static const bool v = !(getenv("IPR_NT_COPY") && atoi(getenv("IPR_NT_COPY")) == 0);
    return v;
}
inline void copy_to_pageable(char* dst, const char* src, size_t n) {
#ifdef IPR_HAVE_SSE2
    if (nt_copy_enabled() && n >= 4096) {
        size_t head = (64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63;
        memcpy(dst, src, head);
        dst += head;
        src += head;
        n -= head;
        const size_t lines = n / 64;
        for (size_t i = 0; i < lines; i++) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 64 * i));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 64 * i + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 64 * i + 32));
            const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 64 * i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 64 * i), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 64 * i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 64 * i + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 64 * i + 48), d);
        }
        _mm_sfence();
        memcpy(dst + lines * 64, src + lines * 64, n - lines * 64);
        return;
    }
#endif
    memcpy(dst, src, n);
}

class PageableSink {
public:
    static constexpr size_t kSlotBytes = 256u << 20;
    static constexpr int kSlots = 4;

    ~PageableSink() { finish(); }

    // n_threads: memcpy fan-out of one slot (the host's cores are shared by the sinks of all devices of the
    // job); cpus: where those threads (and the drain thread) run — the device's local CPUs, or empty
    int start(unsigned n_threads, std::vector<int> cpus) {
        cpus_ = std::move(cpus);
        for (int i = 0; i < kSlots; i++) {
            {   // pinned slots are recycled process-wide (cudaHostAlloc is slow), one set per sink
                std::lock_guard<std::mutex> lk(pinned_mu());
                if (!pinned_free().empty()) {
                    slot_[i] = pinned_free().back();
                    pinned_free().pop_back();
                }
            }
            if (!slot_[i] && cudaHostAlloc(&slot_[i], kSlotBytes, cudaHostAllocPortable) != cudaSuccess) {
                slot_[i] = nullptr;
                cudaGetLastError();
                return IPR_ENOMEM;
            }
            if (cudaEventCreateWithFlags(&ev_[i], cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess)
                return IPR_ECUDA;
        }
        n_threads_ = std::max(1u, n_threads);
        drain_ = std::thread([this] {
            bind_this_thread(cpus_);   // the fan-out threads inherit the mask
            drain_loop();
        });
        started_ = true;
        return IPR_OK;
    }

    // enqueue device -> host copy of `rows` rows of `width` bytes (contiguous on the device) to pageable
    // memory with row pitch `pitch`, on `stream`
    int push2d(const char* d_src, char* dst, size_t width, size_t rows, size_t pitch, cudaStream_t stream) {
        if (width == pitch || rows == 1) return push(d_src, dst, width * rows, 0, 0, stream);
        if (width > kSlotBytes) {  // a single row larger than a staging slot: row by row
            for (size_t r = 0; r < rows; r++) {
                const int rc = push(d_src + r * width, dst + r * pitch, width, 0, 0, stream);
                if (rc) return rc;
            }
            return IPR_OK;
        }
        const size_t rows_per = kSlotBytes / width;
        for (size_t r0 = 0; r0 < rows; r0 += rows_per) {
            const size_t nr = std::min(rows_per, rows - r0);
            const int rc = push(d_src + r0 * width, dst + r0 * pitch, width * nr, width, pitch, stream);
            if (rc) return rc;
        }
        return IPR_OK;
    }

    // `bytes` contiguous on the device; on the host either contiguous (width == 0) or rows of `width`
    // bytes at `pitch` (then bytes is a multiple of width and fits one slot)
    int push(const char* d_src, char* dst, size_t bytes, size_t width, size_t pitch, cudaStream_t stream) {
        for (size_t off = 0; off < bytes; off += kSlotBytes) {
            const size_t n = std::min(kSlotBytes, bytes - off);
            const long long id = issued_;
            const int k = (int)(id % kSlots);
            {   // wait until the slot's previous contents have been drained
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return drained_ >= id - kSlots + 1 || failed_; });
                if (failed_) return IPR_ECUDA;
            }
            if (cudaMemcpyAsync(slot_[k], d_src + off, n, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
                cudaEventRecord(ev_[k], stream) != cudaSuccess)
                return IPR_ECUDA;
            {
                std::lock_guard<std::mutex> lk(mu_);
                work_.push_back({dst + off, n, width, pitch});
                issued_++;
            }
            cv_.notify_all();
        }
        return IPR_OK;
    }

    int finish() {
        if (!started_) {
            for (int i = 0; i < kSlots; i++)
                if (ev_[i]) {
                    cudaEventDestroy(ev_[i]);
                    ev_[i] = nullptr;
                }
            release_slots();
            return IPR_OK;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            done_ = true;
        }
        cv_.notify_all();
        if (drain_.joinable()) drain_.join();
        for (int i = 0; i < kSlots; i++)
            if (ev_[i]) {
                cudaEventDestroy(ev_[i]);
                ev_[i] = nullptr;
            }
        release_slots();
        started_ = false;
        return failed_ ? IPR_ECUDA : IPR_OK;
    }

private:
    struct Work { char* dst; size_t bytes, width, pitch; };
    static std::mutex& pinned_mu() { static std::mutex m; return m; }
    static std::vector<void*>& pinned_free() { static std::vector<void*> v; return v; }
    void release_slots() {
        std::lock_guard<std::mutex> lk(pinned_mu());
        for (int i = 0; i < kSlots; i++)
            if (slot_[i]) {
                pinned_free().push_back(slot_[i]);
                slot_[i] = nullptr;
            }
    }

    void drain_loop() {
        for (long long id = 0;; id++) {
            Work w;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return issued_ > id || done_; });
                if (issued_ <= id) return;
                w = work_[id];
            }
            const int k = (int)(id % kSlots);
            if (cudaEventSynchronize(ev_[k]) != cudaSuccess) {
                std::lock_guard<std::mutex> lk(mu_);
                failed_ = true;
                cv_.notify_all();
                return;
            }
            // parallel first-touch + copy
            const char* src = static_cast<const char*>(slot_[k]);
            std::vector<std::thread> th;
            if (w.width == 0) {
                const size_t per = (w.bytes / n_threads_ + 4095) & ~size_t(4095);
                for (unsigned t = 0; t < n_threads_; t++) {
                    const size_t lo = std::min(w.bytes, (size_t)t * per), hi = std::min(w.bytes, lo + per);
                    if (hi > lo) th.emplace_back([=] { copy_to_pageable(w.dst + lo, src + lo, hi - lo); });
                }
            } else {
                const size_t rows = w.bytes / w.width;
                const size_t per = (rows + n_threads_ - 1) / n_threads_;
                for (unsigned t = 0; t < n_threads_; t++) {
                    const size_t lo = std::min(rows, (size_t)t * per), hi = std::min(rows, lo + per);
                    if (hi > lo)
                        th.emplace_back([=] {
                            for (size_t r = lo; r < hi; r++) copy_to_pageable(w.dst + r * w.pitch, src + r * w.width, w.width);
                        });
                }
            }
            for (auto& x : th) x.join();
            {
                std::lock_guard<std::mutex> lk(mu_);
                drained_ = id + 1;
            }
            cv_.notify_all();
        }
    }

    void* slot_[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    std::thread drain_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<Work> work_;
    long long issued_ = 0, drained_ = 0;
    bool done_ = false, failed_ = false, started_ = false;
    unsigned n_threads_ = 1;
    std::vector<int> cpus_;
};

// A device's column range of the caller's pinned matrix is written layer by layer with 1-D copies rather than one
// pitched 2-D copy per chunk: measured from one process on 8 GPUs (29.2 GB): 310 ms row-wise (94 GB/s, the box's
// host-side ceiling) against 334 ms pitched.  IPR_D2H_ROWS=0 restores cudaMemcpy2DAsync.
bool rowwise_d2h() {
    static const bool v = !(getenv("IPR_D2H_ROWS") && atoi(getenv("IPR_D2H_ROWS")) == 0);
    return v;
}

// One device's share of a host-buffer job: cells [c0, c1) for every date (and every stack).
int run_device_block(const HostJob& job, int device, long long c0, long long c1, char* err, size_t errlen) {
    if (c0 >= c1) return IPR_OK;
    Trace tr;
    const int nd = job.n_dates;
    const long long nc = c1 - c0;
    // this thread and the host-side copy threads of this device stay on the device's local CPUs
    const std::vector<int> local_cpus = device_local_cpus(device);
    AffinityScope affinity(local_cpus);
    ipr_plan* plan = nullptr;
    int rc = ipr_plan_create(&plan, device, job.cell_x + c0, job.cell_y + c0, nc, job.st_x, job.st_y, job.n_st, nd, err,
                             errlen);
    if (rc) return rc;
    struct Guard {
        ipr_plan* p;
        double* ring[kRing] = {nullptr, nullptr, nullptr};
        cudaEvent_t done[kRing] = {nullptr, nullptr, nullptr}, computed[kRing] = {nullptr, nullptr, nullptr};
        cudaStream_t copy = nullptr;
        ~Guard() {
            Trace td;
            cudaSetDevice(p->device);
            if (p->stream) cudaStreamSynchronize(p->stream);   // kernels may still write the ring on an error return
            if (copy) {
                cudaStreamSynchronize(copy);
                release_stream(copy);   // synchronised just above
            }
            for (int i = 0; i < kRing; i++) {
                dev_free(ring[i]);
                if (done[i]) cudaEventDestroy(done[i]);
                if (computed[i]) cudaEventDestroy(computed[i]);
            }
            td.mark("(teardown) ring freed");
            ipr_plan_destroy(p);
            td.mark("(teardown) plan destroyed");
        }
    } g{plan};
    if ((job.round_mode != 0 || job.cell_keep) &&
        (rc = ipr_plan_set_post(plan, job.round_mode, job.n_round, job.cell_keep ? job.cell_keep + c0 : nullptr, err,
                                errlen)))
        return rc;
    tr.mark("plan created (allocations, coordinate upload, zero-distance scan)");
    rc = plan_set_obs_ld(plan, job.obs, job.n_dates, err, errlen);
    if (rc) return rc;
    tr.mark("obs uploaded and prepared");
    const int n_stacks = job.cressman ? job.n_radii : 1;
    const std::vector<WorkItem> items = make_items(n_stacks, nd);
    const size_t slot_doubles = (size_t)nc * std::min(kGran, (int)round_up(nd, 16));
    IPR_CUDA(acquire_stream(&g.copy));
    PageableSink sink;
    if (!job.out_pinned) {
        // the host's cores are shared by the sinks of all devices of the job
        const unsigned cores = local_cpus.empty() ? std::max(1u, std::thread::hardware_concurrency())
                                                  : (unsigned)local_cpus.size();
        const unsigned share = local_cpus.empty() ? cores / (unsigned)std::max(1, job.n_job_devices) : cores;
        unsigned fan = std::max(2u, std::min(16u, share));
        if (const char* e = getenv("IPR_DRAIN_THREADS")) fan = (unsigned)std::max(1, atoi(e));
        if ((rc = sink.start(fan, local_cpus))) {
            set_err(err, errlen, "could not set up the pinned staging ring for a pageable output buffer");
            return rc;
        }
    }
    const int n_ring = std::min<int>(kRing, (int)items.size());
    for (int i = 0; i < n_ring; i++) {
        IPR_CUDA(dev_malloc(&g.ring[i], sizeof(double) * slot_doubles));
        IPR_CUDA(cudaEventCreateWithFlags(&g.done[i], cudaEventDisableTiming));
        IPR_CUDA(cudaEventCreateWithFlags(&g.computed[i], cudaEventDisableTiming));
    }
    tr.mark("ring allocated");
    bool cancelled = false;
    long long reported = 0;
    // the host polls the slot's event (sleeping 100 us between queries: no interrupt-driven blocking sync on the
    // copy stream); IPR_PACE=0 queues everything at once behind stream waits (no progress, cancel only at the start)
    const bool pace_on_host = !(getenv("IPR_PACE") && atoi(getenv("IPR_PACE")) == 0);
    for (size_t ii = 0; ii < items.size(); ii++) {
        const WorkItem& it = items[ii];
        const int slot = (int)(ii % n_ring);
        if ((int)ii >= n_ring) {
            // the slot's previous chunk has reached the host: the GPU still holds the chunks in between, so
            // waiting here on the host costs no bubble, and it is where a job is cancelled / reports progress
            if (pace_on_host) {
                cudaError_t q;
                while ((q = cudaEventQuery(g.done[slot])) == cudaErrorNotReady)
                    std::this_thread::sleep_for(std::chrono::microseconds(100));
                IPR_CUDA(q);
                if (job.done) job.done->fetch_add(1);
                reported++;
            } else {
                IPR_CUDA(cudaStreamWaitEvent(plan->stream, g.done[slot], 0));
            }
        }
        if (job.cancel && job.cancel->load() != 0) {
            cancelled = true;
            break;
        }
        if (job.cressman)
            rc = ipr_plan_cressman(plan, job.radii + it.stack, 1, it.t0, it.t1, g.ring[slot], nc, 0, err, errlen);
        else
            rc = job.kriging ? (job.kr_model ? ipr_plan_kriging(plan, job.kr_model, job.kr_nugget, job.kr_sill, job.kr_range,
                                                                it.t0, it.t1, g.ring[slot], nc, err, errlen)
                                             : ipr_plan_idw_kriging(plan, it.t0, it.t1, g.ring[slot], nc, err, errlen))
                             : ipr_plan_idw(plan, job.p, it.t0, it.t1, g.ring[slot], nc, err, errlen);
        if (rc) return rc;
        IPR_CUDA(cudaEventRecord(g.computed[slot], plan->stream));
        IPR_CUDA(cudaStreamWaitEvent(g.copy, g.computed[slot], 0));
        // layer t of the caller's matrix is a row of n_cells doubles; this device owns columns [c0, c1)
        double* dst = job.out + (size_t)it.stack * job.n_cells * job.n_dates + (size_t)it.t0 * job.n_cells + c0;
        const size_t width = sizeof(double) * (size_t)nc, pitch = sizeof(double) * (size_t)job.n_cells;
        const size_t rows = (size_t)(it.t1 - it.t0);
        if (job.out_pinned) {
            if (width == pitch)
                IPR_CUDA(cudaMemcpyAsync(dst, g.ring[slot], width * rows, cudaMemcpyDeviceToHost, g.copy));
            else if (rowwise_d2h())
                for (size_t rr = 0; rr < rows; rr++)
                    IPR_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(dst) + rr * pitch,
                                             reinterpret_cast<const char*>(g.ring[slot]) + rr * width, width,
                                             cudaMemcpyDeviceToHost, g.copy));
            else
                IPR_CUDA(cudaMemcpy2DAsync(dst, pitch, g.ring[slot], width, width, rows, cudaMemcpyDeviceToHost, g.copy));
        } else if ((rc = sink.push2d(reinterpret_cast<const char*>(g.ring[slot]), reinterpret_cast<char*>(dst), width,
                                     rows, pitch, g.copy))) {
            set_err(err, errlen, "staged device-to-host copy failed");
            return rc;
        }
        IPR_CUDA(cudaEventRecord(g.done[slot], g.copy));
    }
    tr.mark(cancelled ? "cancelled" : "all chunks submitted");
    if ((rc = ipr_plan_sync(plan))) {
        set_err(err, errlen, "stream synchronisation failed");
        return rc;
    }
    tr.mark("compute stream drained");
    IPR_CUDA(cudaStreamSynchronize(g.copy));
    tr.mark("copy stream drained");
    if ((rc = sink.finish())) {
        set_err(err, errlen, "staged device-to-host copy failed");
        return rc;
    }
    tr.mark("host staging drained");
    if (cancelled) {
        set_err(err, errlen, "cancelled");
        return IPR_ECANCELLED;
    }
    if (job.done) job.done->fetch_add((long long)items.size() - reported);
    return IPR_OK;
}

// contiguous cell ranges on 64-cell tile boundaries, equal to within one tile (ipr_cell_split_bounds)
void cell_split(long long n_cells, int G, long long* bounds) {
    const long long n_tiles = (n_cells + kCellsPerCta - 1) / kCellsPerCta;
    for (int g = 0; g <= G; g++) bounds[g] = std::min<long long>(n_cells, n_tiles * g / G * kCellsPerCta);
    bounds[G] = n_cells;
}

int run_host_job(const HostJob& job, const int* devices, int n_devices, char* err, size_t errlen) {
    if (!job.cell_x || !job.cell_y || !job.st_x || !job.st_y || !job.obs || !job.out) {
        set_err(err, errlen, "NULL argument");
        return IPR_EINVAL;
    }
    if (job.n_cells <= 0 || job.n_st <= 0 || job.n_dates <= 0) {
        set_err(err, errlen, "n_cells, n_st and n_dates must be positive");
        return IPR_EINVAL;
    }
    Trace tj;
    std::vector<int> devs;
    if (!devices || n_devices <= 0) devs.push_back(0);
    else devs.assign(devices, devices + n_devices);
    const int G = (int)devs.size();
    // is the caller's output page-locked (then D2H lands in it directly), or pageable (staged)?
    HostJob jobx = job;
    {
        cudaPointerAttributes attr;
        cudaError_t e = cudaPointerGetAttributes(&attr, job.out);
        if (e != cudaSuccess) cudaGetLastError();
        jobx.out_pinned = (e == cudaSuccess && attr.type == cudaMemoryTypeHost);
    }
    if (!jobx.out_pinned && getenv("IPR_THP") && atoi(getenv("IPR_THP")) != 0) {
        // opt-in: transparent huge pages for a fresh pageable result.  Measured on the B200 box (THP = madvise,
        // 29.2 GB never-touched matrix): 1.39 s with the hint against 0.95 s without — zeroing 2 MB pages inside
        // the fault handler serialises what the 4 KB faults of the 16 staging threads spread out — so it is off
        const size_t bytes = sizeof(double) * (size_t)job.n_cells * job.n_dates * (job.cressman ? job.n_radii : 1);
        const uintptr_t lo = (reinterpret_cast<uintptr_t>(job.out) + 4095) & ~uintptr_t(4095);
        const uintptr_t hi = (reinterpret_cast<uintptr_t>(job.out) + bytes) & ~uintptr_t(4095);
        if (hi > lo) madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);
    }
    // Every output element depends only on its own cell and date, so the devices split the
    // CELLS into contiguous ranges on 64-cell tile boundaries: nothing is computed twice (a date split
    // would make every device regenerate all cell x station weights), the shares are equal to within
    // one tile, and each device writes its own column range of every layer.  No inter-GPU exchange.
    std::vector<long long> bounds(G + 1, 0);
    cell_split(job.n_cells, G, bounds.data());
    jobx.n_job_devices = G;
    std::vector<int> rcs(G, IPR_OK);
    std::vector<std::string> errs(G, std::string(256, '\0'));
    auto run_block = [&](int g) {
        char* e = &errs[g][0];
        const size_t n = errs[g].size();
        rcs[g] = no_throw(e, n, [&] { return run_device_block(jobx, devs[g], bounds[g], bounds[g + 1], e, n); });
        if (rcs[g] != IPR_OK && jobx.cancel) jobx.cancel->store(1);   // no point in finishing the other shares
    };
    if (G == 1) {
        run_block(0);
    } else {
        std::vector<std::thread> th;
        for (int g = 0; g < G; g++) th.emplace_back(run_block, g);
        for (auto& t : th) t.join();
    }
    tj.mark("(job) all device blocks done");
    // a real failure on any device wins over the IPR_ECANCELLED the others report after the job's flag was raised
    int first = -1;
    for (int g = 0; g < G; g++)
        if (rcs[g] != IPR_OK && (first < 0 || (rcs[first] == IPR_ECANCELLED && rcs[g] != IPR_ECANCELLED))) first = g;
    if (first >= 0) {
        set_err(err, errlen, "device %d: %s", devs[first], errs[first].c_str());
        return rcs[first];
    }
    return IPR_OK;
}

}  // namespace

// asynchronous job: the one-shot call on a worker thread, polled by the host language's main thread
struct ipr_job {
    HostJob job;
    std::vector<int> devs;
    std::vector<double> radii;        // copied: the caller's radii vector need not outlive the start call
    std::atomic<int> cancel{0};
    std::atomic<long long> done{0};
    long long total = 0;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    bool finished = false;
    int rc = IPR_OK;
    char err[512] = "";
};

namespace {

int job_start(ipr_job** out, HostJob job, const int* devices, int n_devices, char* err, size_t errlen) {
    if (!out) {
        set_err(err, errlen, "NULL argument");
        return IPR_EINVAL;
    }
    *out = nullptr;
    return no_throw(err, errlen, [&] {
        std::unique_ptr<ipr_job> j(new ipr_job());
        j->job = job;
        if (job.cressman) {
            j->radii.assign(job.radii, job.radii + job.n_radii);
            j->job.radii = j->radii.data();
        }
        if (devices && n_devices > 0) j->devs.assign(devices, devices + n_devices);
        else j->devs.push_back(0);
        j->job.cancel = &j->cancel;
        j->job.done = &j->done;
        const int n_stacks = job.cressman ? job.n_radii : 1;
        j->total = (long long)make_items(n_stacks, std::max(job.n_dates, 1)).size() * (long long)j->devs.size();
        ipr_job* p = j.get();
        p->th = std::thread([p] {
            const int rc = no_throw(p->err, sizeof p->err, [&] {
                return run_host_job(p->job, p->devs.data(), (int)p->devs.size(), p->err, sizeof p->err);
            });
            std::lock_guard<std::mutex> lk(p->mu);
            p->rc = rc;
            p->finished = true;
            p->cv.notify_all();
        });
        *out = j.release();
        return IPR_OK;
    });
}

}  // namespace

extern "C" {

int ipr_job_start_idw(ipr_job** job, const double* cell_x, const double* cell_y, int64_t n_cells, const double* st_x,
                      const double* st_y, int32_t n_st, const double* obs, int32_t n_dates, double p,
                      int32_t round_mode, int32_t n_round, const uint8_t* cell_keep, const int* devices, int n_devices,
                      double* out, char* err, size_t errlen) {
    HostJob hj{cell_x, cell_y, st_x, st_y, obs, n_cells, n_st, n_dates, false, p, nullptr, 0, out};
    hj.round_mode = round_mode;
    hj.n_round = n_round;
    hj.cell_keep = cell_keep;
    return job_start(job, hj, devices, n_devices, err, errlen);
}

int ipr_job_start_cressman(ipr_job** job, const double* cell_x, const double* cell_y, int64_t n_cells,
                           const double* st_x, const double* st_y, int32_t n_st, const double* obs, int32_t n_dates,
                           const double* radii_m, int32_t n_radii, int32_t round_mode, int32_t n_round,
                           const uint8_t* cell_keep, const int* devices, int n_devices, double* out, char* err,
                           size_t errlen) {
    if (!radii_m || n_radii <= 0) {
        set_err(err, errlen, "radii_m must hold at least one radius");
        return IPR_EINVAL;
    }
    HostJob hj{cell_x, cell_y, st_x, st_y, obs, n_cells, n_st, n_dates, true, 2.0, radii_m, n_radii, out};
    hj.round_mode = round_mode;
    hj.n_round = n_round;
    hj.cell_keep = cell_keep;
    return job_start(job, hj, devices, n_devices, err, errlen);
}

int ipr_job_wait(ipr_job* job, int timeout_ms, int64_t* done, int64_t* total) {
    if (!job) return IPR_EINVAL;
    bool fin;
    {
        std::unique_lock<std::mutex> lk(job->mu);
        if (timeout_ms < 0) job->cv.wait(lk, [&] { return job->finished; });
        else job->cv.wait_for(lk, std::chrono::milliseconds(timeout_ms), [&] { return job->finished; });
        fin = job->finished;
    }
    if (done) *done = std::min<long long>(job->done.load(), job->total);
    if (total) *total = job->total;
    return fin ? 1 : 0;
}

void ipr_job_cancel(ipr_job* job) {
    if (job) job->cancel.store(1);
}

int ipr_job_finish(ipr_job* job, char* err, size_t errlen) {
    if (!job) return IPR_EINVAL;
    if (job->th.joinable()) job->th.join();
    const int rc = job->rc;
    if (rc != IPR_OK) set_err(err, errlen, "%s", job->err);
    delete job;
    return rc;
}

void ipr_cell_split_bounds(int64_t n_cells, int n_devices, int64_t* bounds) {
    if (!bounds || n_devices < 1) return;
    std::vector<long long> b((size_t)n_devices + 1, 0);
    cell_split(std::max<long long>(n_cells, 0), n_devices, b.data());
    for (int g = 0; g <= n_devices; g++) bounds[g] = b[g];
}

int ipr_idw_f64(const double* cell_x, const double* cell_y, int64_t n_cells, const double* st_x, const double* st_y,
                int32_t n_st, const double* obs, int32_t n_dates, double p, const int* devices, int n_devices,
                double* out, char* err, size_t errlen) {
    HostJob job{cell_x, cell_y, st_x, st_y, obs, n_cells, n_st, n_dates, false, p, nullptr, 0, out};
    return no_throw(err, errlen, [&] { return run_host_job(job, devices, n_devices, err, errlen); });
}

int ipr_idw_kriging_f64(const double* cell_x, const double* cell_y, int64_t n_cells, const double* st_x,
                        const double* st_y, int32_t n_st, const double* obs, int32_t n_dates, const int* devices,
                        int n_devices, double* out, char* err, size_t errlen) {
    HostJob job{cell_x, cell_y, st_x, st_y, obs, n_cells, n_st, n_dates, false, 2.0, nullptr, 0, out};
    job.kriging = true;
    return no_throw(err, errlen, [&] { return run_host_job(job, devices, n_devices, err, errlen); });
}

int ipr_kriging_f64(const double* cell_x, const double* cell_y, int64_t n_cells, const double* st_x, const double* st_y,
                    int32_t n_st, const double* obs, int32_t n_dates, const int32_t* model, const double* nugget,
                    const double* sill, const double* range, const int* devices, int n_devices, double* out, char* err,
                    size_t errlen) {
    if (!model || !nugget || !sill || !range) {
        set_err(err, errlen, "NULL argument");
        return IPR_EINVAL;
    }
    HostJob job{cell_x, cell_y, st_x, st_y, obs, n_cells, n_st, n_dates, false, 2.0, nullptr, 0, out};
    job.kriging = true;
    job.kr_model = model;
    job.kr_nugget = nugget;
    job.kr_sill = sill;
    job.kr_range = range;
    return no_throw(err, errlen, [&] { return run_host_job(job, devices, n_devices, err, errlen); });
}

int ipr_idw_post_f64(const double* cell_x, const double* cell_y, int64_t n_cells, const double* st_x,
                     const double* st_y, int32_t n_st, const double* obs, int32_t n_dates, double p, int32_t round_mode,
                     int32_t n_round, const uint8_t* cell_keep, const int* devices, int n_devices, double* out,
                     char* err, size_t errlen) {
    HostJob job{cell_x, cell_y, st_x, st_y, obs, n_cells, n_st, n_dates, false, p, nullptr, 0, out};
    job.round_mode = round_mode;
    job.n_round = n_round;
    job.cell_keep = cell_keep;
    return no_throw(err, errlen, [&] { return run_host_job(job, devices, n_devices, err, errlen); });
}

int ipr_cressman_post_f64(const double* cell_x, const double* cell_y, int64_t n_cells, const double* st_x,
                          const double* st_y, int32_t n_st, const double* obs, int32_t n_dates,
                          const double* radii_m, int32_t n_radii, int32_t round_mode, int32_t n_round,
                          const uint8_t* cell_keep, const int* devices, int n_devices, double* out, char* err,
                          size_t errlen) {
    if (!radii_m || n_radii <= 0) {
        set_err(err, errlen, "radii_m must hold at least one radius");
        return IPR_EINVAL;
    }
    HostJob job{cell_x, cell_y, st_x, st_y, obs, n_cells, n_st, n_dates, true, 2.0, radii_m, n_radii, out};
    job.round_mode = round_mode;
    job.n_round = n_round;
    job.cell_keep = cell_keep;
    return no_throw(err, errlen, [&] { return run_host_job(job, devices, n_devices, err, errlen); });
}

int ipr_cressman_f64(const double* cell_x, const double* cell_y, int64_t n_cells, const double* st_x,
                     const double* st_y, int32_t n_st, const double* obs, int32_t n_dates, const double* radii_m,
                     int32_t n_radii, const int* devices, int n_devices, double* out, char* err, size_t errlen) {
    if (!radii_m || n_radii <= 0) {
        set_err(err, errlen, "radii_m must hold at least one radius");
        return IPR_EINVAL;
    }
    HostJob job{cell_x, cell_y, st_x, st_y, obs, n_cells, n_st, n_dates, true, 2.0, radii_m, n_radii, out};
    return no_throw(err, errlen, [&] { return run_host_job(job, devices, n_devices, err, errlen); });
}

}  // extern "C"
