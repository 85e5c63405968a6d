// interpolater_b200 — sm_100a kernels for the IDW / Cressman cell x station x date contraction.
//
// Math (reference: R/IDW.R:260-342, R/Cressman.R:279-321 — restated in oracle/reference_port.py):
//   IDW       out[c,t] = sum_s w(c,s) z0[s,t] / sum_s w(c,s) m[s,t],  w = d(c,s)^-p
//   Cressman  out[c,t] = sum_s w(c,s) z0[s,t] / sum_s w(c,s),         w = (R^2-d^2)/(R^2+d^2), 0 if d > R
// i.e. a dense FP64 GEMM  [cells x stations] * [stations x dates]  on the FP64 tensor pipe
// (mma.sync m8n8k4 -> SASS DMMA.8x8x4, the only FP64 tensor op sm_100 has; tcgen05 has no f64 kind).
//
// Two kernels:
//   weights_kernel   generates the A operand ONCE per (geometry, weight function): every thread
//                    computes its weights directly in the DMMA A-fragment register layout
//                    (a[row = lane/4][k = lane%4]) and streams them to HBM in fragment order, plus
//                    the per-cell weight sum.  (Generating weights inside the contraction costs
//                    ~10 FP64-pipe instructions per weight per 256 dates — the FP64 pipe is the
//                    roofline here, HBM is idle — so the weights are traded to HBM: +0.3 % time for
//                    the generation pass, -7 % in the contraction.)
//   contract_kernel  C tile = 64 cells x 128 dates per CTA (64 dates x {numerator, mask} when the
//                    granule holds NA), 8 warps x (8 cells x all dates), 64 accumulator registers
//                    per thread -> 128 registers, TWO CTAs per SM: one CTA's prologue / epilogue /
//                    barrier waits hide under the other's DMMAs (measured 92.5 % -> 95.5 % of peak
//                    against one 256-date CTA per SM).  A fragments: one coalesced LDG.64 per
//                    lane per k-step, prefetched a stage ahead.  B (z0 / mask granule blocks, see
//                    below): L2 -> shared memory with cp.async.bulk (TMA bulk engine, UBLKCP) through
//                    a kStages-deep mbarrier pipeline; a slot is refilled by the last warp that
//                    releases it, so no warp ever blocks on "slot empty".  Shared rows are padded
//                    by 4 doubles: the LDS.128 B-fragment reads are bank-conflict free.
//                    Epilogue: normalise, NA rule, per-date early-outs, store (cells contiguous).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ipr {

constexpr int kConsumerWarps = 8;
constexpr int kThreads = kConsumerWarps * 32;
constexpr int kCellsPerCta = kConsumerWarps * 8;     // 64
#ifndef IPR_KT
#define IPR_KT 16
#endif
constexpr int kKT = IPR_KT;                          // stations per pipeline stage
constexpr int kKSteps = kKT / 4;                     // DMMA k-steps per stage
constexpr int kMaxTilesPerLaunch = 56;
// Device layout of the B operand ("granule blocks"): dates are cut into granules of 128, stations
// into pipeline stages of kKT; block (g, stage) = [kKT rows][kGranLd doubles] is ONE contiguous
// chunk that is already in the padded shared-memory layout, so a stage is refilled by one bulk
// copy per granule (33 row-sized copies per stage were measured to cap the kernel at 70 % of peak).
constexpr int kGran = 128;
constexpr int kGranLd = kGran + 4;
constexpr int kBlkDoubles = kKT * kGranLd;
constexpr int kBlkBytes = kBlkDoubles * 8;
// A operand cache: weights in fragment order, [cell tile][stage][k-step][warp][lane]
constexpr int kWStageDoubles = kKSteps * kThreads;
// Cressman walks the station axis in HALF stages (2 k-steps = 8 stations): both operand layouts
// are k-step-major inside a stage, so a half stage is a contiguous half of the same A chunk / B
// block and the finer list skips more of the exactly-zero product (measured -12 % on cfg 4; IDW,
// which visits every stage anyway, lost 2.6 % with 8-station stages and keeps 16).
constexpr int kSubStages = 2;
constexpr int kCressmanKSteps = kKSteps / kSubStages;
constexpr double kPadCoord = 1e150;                  // padded stations: d^2 ~ 2e300 -> weight ~ 0
constexpr unsigned long long kNaRealBits = 0x7FF00000000007A2ull;

enum WeightMode : int {
    kIdwP2 = 0,       // w = 1/d^2
    kIdwEvenP = 1,    // w = (1/d^2)^n, p = 2n
    kIdwGenericP = 2, // w = pow(d^2, -p/2)
    kCressman = 3,    // w = (R^2 - d^2)/(R^2 + d^2), 0 when d^2 > R^2
    kIdwKriging = 4   // Kriging_Ordinary's IDW fallback: w = 1/d^2 with d == 0 replaced by 1e-10 (no fix-up)
};

struct WeightParams {
    double a;   // kIdwGenericP: -p/2 ; kCressman: R^2
    int n;      // kIdwEvenP: p/2
};

struct TileList {
    int n;
    int t0[kMaxTilesPerLaunch];   // absolute start date of each tile (multiple of 128; of 64 when masked)
};

struct ContractArgs {
    const double* wfrag;      // A operand cache for cell tiles [cell_tile0, cell_tile0 + n_cell_tiles)
    const double* wsum;       // [c_pad] per-cell sum of weights over all stations
    const int* stage_list;    // [n_cell_tiles][kSubStages * n_stages] (half) stages holding any non-zero weight,
                              // ascending, in units of the instantiation's KSTEPS k-steps
    const int* n_active;      // [n_cell_tiles] length of each list
    const int* run_perm;      // nullptr, or [c_pad/8]: warp w of cell tile T owns the 8 consecutive cells
                              // of run run_perm[8T + w] (spatially compact tiles for Cressman)
    const double* z0;         // granule blocks [n_gran][n_stages][kBlkDoubles]
    const double* mask;       // same layout (masked variants only)
    const uint8_t* date_flag; // IDW: 0 normal, 1 all-NA layer, 2 exact-zero layer, 3 constant layer date_const[t]
                              // (Kriging fallback); nullptr for Cressman
    const double* date_const; // [n_dates] value of the flag-3 layers, or nullptr
    double* out;              // layer t at out + (t - t_lo) * ld_out
    long long ld_out;
    long long n_cells;
    int cell_tile0, n_cell_tiles;
    int n_stages;
    int t_lo, t_hi;           // store window [t_lo, t_hi)
    const int* run_if;        // nullptr, or a device flag: the launch is a no-op unless it is non-zero (fallback
                              // of the int8 mask path, decided on the device)
    TileList tiles;
};

struct WeightArgs {
    const double2* cell_xy;   // [c_pad]
    const double2* st_xy;     // [s_pad]
    double* wfrag;            // chunk base
    double* wsum;             // [c_pad]
    double* wmax;             // [c_pad] largest weight of the cell (row scale of the int8 slices)
    int* stage_list;          // [n_cell_tiles][kSubStages * n_stages]
    int* n_active;            // [n_cell_tiles]
    const int* run_perm;      // see ContractArgs
    const double4* stage_box; // nullptr, or [kSubStages * n_stages] bounding box (xmin, ymin, xmax, ymax) of each
                              // half stage's stations: Cressman skips those farther than R from the tile's box
    unsigned long long* active_total;  // cumulative count of active (cell tile, stage) blocks
    int cell_tile0, n_cell_tiles;
    int n_stages;
    int n_st;                 // real stations; positions >= n_st are padding and carry exactly zero weight
    WeightParams wp;
};

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (TMA bulk engine)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// D(8x8) += A(8x4, row) * B(4x8, col), FP64 tensor core (SASS: DMMA.8x8x4)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
// 1/x: MUFU.RCP64H seed + 2 Newton-Raphson steps (measured bit-identical to IEEE division on the
// probe's 3e5 samples, profiles/r01_fp64_probe.log); x finite, > 0
__device__ __forceinline__ double rcp_nr2(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

template <int WMODE>
__device__ __forceinline__ double weight_from_d2(double d2, const WeightParams& wp) {
    if (WMODE == kCressman) {
        const double num = wp.a - d2;
        const double den = wp.a + d2;
        const double w = num * rcp_nr2(den);
        return d2 > wp.a ? 0.0 : w;
    } else {
        // exact zero distance: excluded here (weight 0) and patched by idw_zero_fix_kernel,
        // R/IDW.R:310-326.  The test is on the bit pattern so it stays off the FP64 pipe.
        const bool zero = ((__double2hiint(d2) | __double2loint(d2)) == 0);
        double w;
        if (WMODE == kIdwKriging) {
            // R/Kriging_Ordinary.R:470-471: distances[distances == 0] <- 1e-10; weights <- 1 / distances^2
            return rcp_nr2(zero ? 1e-10 * 1e-10 : d2);
        } else if (WMODE == kIdwP2) {
            w = rcp_nr2(d2);
        } else if (WMODE == kIdwEvenP) {
            const double r = rcp_nr2(d2);
            w = r;
            for (int i = 1; i < wp.n; i++) w *= r;
        } else {
            w = pow(d2, wp.a);
        }
        return zero ? 0.0 : w;
    }
}

template <bool MASKED, int KSTEPS = kKSteps>
struct SmemLayout {
    static constexpr int kBlocks = MASKED ? 2 : 1;   // granule blocks per stage (z0 [+ mask])
    // pipeline depth (two CTAs share the SM's 227 KB); half stages: same bytes in flight
    static constexpr int kStages = MASKED ? 3 : (KSTEPS == kKSteps ? 5 : 8);
    static constexpr int kPartDoubles = KSTEPS * 4 * kGranLd;   // one (half) block: KSTEPS*4 station rows
    static constexpr int kPartBytes = kPartDoubles * 8;
    static constexpr int kStageBytes = kBlocks * kPartBytes;
    static constexpr int kBarOff = kStageBytes * kStages;
    static constexpr int kTotal = kBarOff + kStages * 8 + kStages * 4 + 16;
    static constexpr uint32_t kTxBytes = kBlocks * kPartBytes;
};

// Refill one pipeline stage: one bulk copy (TMA bulk engine) per granule block.  Called by one lane.
// it = position in the tile's active-stage list (selects the slot), sid = the stage itself.
template <bool MASKED, int KSTEPS>
__device__ __forceinline__ void produce_stage(const ContractArgs& args, unsigned char* smem, uint64_t* full_bar,
                                              int it, int sid, int gran0) {
    using L = SmemLayout<MASKED, KSTEPS>;
    const int slot = it % L::kStages;
    unsigned char* stage = smem + slot * L::kStageBytes;
    // generic-proxy reads of this slot (all warps, ordered by the release counter) precede the
    // async-proxy writes issued below
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&full_bar[slot], L::kTxBytes);
    // blocks of a granule are contiguous and rows are k-step-major: (half) stage sid starts sid parts in
    const size_t b0 = (size_t)gran0 * args.n_stages * kBlkDoubles + (size_t)sid * L::kPartDoubles;
    bulk_g2s(stage, args.z0 + b0, L::kPartBytes, &full_bar[slot]);
    if (MASKED) bulk_g2s(stage + L::kPartBytes, args.mask + b0, L::kPartBytes, &full_bar[slot]);
}

// ---------------------------------------------------------------------------------------------
// A operand: weights in DMMA fragment order + per-cell weight sums.  One CTA per cell tile.
// ---------------------------------------------------------------------------------------------
template <int WMODE>
__global__ void __launch_bounds__(kThreads) weights_kernel(const WeightArgs args) {
    constexpr int kChunkStages = 32;               // stages staged in shared memory at a time
    constexpr int kChunk = kChunkStages * kKT;     // 1024 stations, 16 KB
    __shared__ double2 s_st[kChunk];
    __shared__ int s_n_active;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q = lane & 3;   // station inside the k-step (A column)
    const int r = lane >> 2;  // cell row of the fragment
    const int ct = args.cell_tile0 + blockIdx.x;
    const long long run = args.run_perm ? (long long)__ldg(args.run_perm + (size_t)ct * 8 + warp) : (long long)ct * 8 + warp;
    const long long cell = run * 8 + r;
    const double2 cxy = args.cell_xy[cell];
    // Cressman records (and culls) half stages; IDW whole ones
    constexpr int SUB = (WMODE == kCressman) ? kSubStages : 1;
    constexpr int KS = kKSteps / SUB;
    double* dst = args.wfrag + (size_t)blockIdx.x * args.n_stages * kWStageDoubles + threadIdx.x;
    int* list = args.stage_list + (size_t)blockIdx.x * args.n_stages * kSubStages;
    double wsum[4] = {0.0, 0.0, 0.0, 0.0};
    double wmx = 0.0;
    if (threadIdx.x == 0) s_n_active = 0;
    // bounding box of the tile's cells (Cressman culling)
    __shared__ double s_box[kConsumerWarps][4];
    double bx0 = cxy.x, bx1 = cxy.x, by0 = cxy.y, by1 = cxy.y;
    if (WMODE == kCressman && args.stage_box != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bx0 = fmin(bx0, __shfl_xor_sync(0xffffffffu, bx0, o));
            bx1 = fmax(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
            by0 = fmin(by0, __shfl_xor_sync(0xffffffffu, by0, o));
            by1 = fmax(by1, __shfl_xor_sync(0xffffffffu, by1, o));
        }
        if (lane == 0) {
            s_box[warp][0] = bx0;
            s_box[warp][1] = bx1;
            s_box[warp][2] = by0;
            s_box[warp][3] = by1;
        }
        __syncthreads();
#pragma unroll
        for (int w8 = 0; w8 < kConsumerWarps; w8++) {
            bx0 = fmin(bx0, s_box[w8][0]);
            bx1 = fmax(bx1, s_box[w8][1]);
            by0 = fmin(by0, s_box[w8][2]);
            by1 = fmax(by1, s_box[w8][3]);
        }
    }
    for (int st0 = 0; st0 < args.n_stages; st0 += kChunkStages) {
        const int ns = min(kChunkStages, args.n_stages - st0);
        __syncthreads();
        for (int i = threadIdx.x; i < ns * kKT; i += kThreads) s_st[i] = args.st_xy[st0 * kKT + i];
        __syncthreads();
        for (int sl = 0; sl < ns; sl++) {
#pragma unroll
            for (int h = 0; h < SUB; h++) {
                const int sub = (st0 + sl) * SUB + h;
                if (WMODE == kCressman && args.stage_box != nullptr) {
                    // every station of the (half) stage farther than R from every cell of the tile
                    // (box-to-box distance, with a safety margin): all weights are exactly 0
                    const double4 b = args.stage_box[sub];
                    const double gx = fmax(0.0, fmax(b.x - bx1, bx0 - b.z));
                    const double gy = fmax(0.0, fmax(b.y - by1, by0 - b.w));
                    if (gx * gx + gy * gy > args.wp.a * (1.0 + 1e-9)) continue;
                }
                double w[KS];
                bool any = false;
                // independent reciprocal chains overlap; fixed summation order per k-step residue
#pragma unroll
                for (int kk = 0; kk < KS; kk++) {
                    const int ks = h * KS + kk;      // k-step inside the 16-station stage
                    const double2 s = s_st[sl * kKT + ks * 4 + q];
                    const double dx = cxy.x - s.x, dy = cxy.y - s.y;
                    w[kk] = weight_from_d2<WMODE>(fma(dy, dy, dx * dx), args.wp);
                    // padding is given weight 0 explicitly: at (1e150, 1e150) a small generic p would not underflow
                    if ((st0 + sl) * kKT + ks * 4 + q >= args.n_st) w[kk] = 0.0;
                    wsum[ks & 3] += w[kk];
                    wmx = fmax(wmx, w[kk]);
                    any |= (w[kk] != 0.0);
                }
                // a (half) stage whose 64 x 4*KS weights are all zero (Cressman: every station beyond R) is
                // not stored and will be skipped by the contraction
                if (__syncthreads_or(any)) {
                    double* d = dst + (size_t)(st0 + sl) * kWStageDoubles + (size_t)h * KS * kThreads;
#pragma unroll
                    for (int kk = 0; kk < KS; kk++) __stcs(d + kk * kThreads, w[kk]);
                    if (threadIdx.x == 0) list[s_n_active++] = sub;
                }
            }
        }
    }
    double tot = (wsum[0] + wsum[1]) + (wsum[2] + wsum[3]);
    // lanes of a quad share the cell; their partial sums cover disjoint stations
    tot += __shfl_xor_sync(0xffffffffu, tot, 1);
    tot += __shfl_xor_sync(0xffffffffu, tot, 2);
    wmx = fmax(wmx, __shfl_xor_sync(0xffffffffu, wmx, 1));
    wmx = fmax(wmx, __shfl_xor_sync(0xffffffffu, wmx, 2));
    if (q == 0) {
        args.wsum[cell] = tot;
        args.wmax[cell] = wmx;
    }
    if (threadIdx.x == 0) {
        args.n_active[blockIdx.x] = s_n_active;
        atomicAdd(args.active_total, (unsigned long long)s_n_active);
    }
}

// ---------------------------------------------------------------------------------------------
// DMMA contraction + normalise epilogue
//   NP     : number of 16-date column groups computed (tile width 16*NP dates: 128 for a full
//            unmasked granule, 64 for a masked half granule, less for the ragged last granule)
//   MASKED : also contract the availability mask (IDW with NA in the granule: 4*S flop/cell-day)
// grid.x = n_cell_tiles * n_tiles, date tile fastest: CTAs that run together share the A tile
// through L2 and the whole B operand (tens of MB) stays L2-resident.
// ---------------------------------------------------------------------------------------------
//   KSTEPS : k-steps per pipeline stage: kKSteps (16 stations), or kCressmanKSteps (8) for Cressman's finer lists
template <int NP, bool MASKED, bool DENOM_IN_OUT = false, int KSTEPS = kKSteps>
__global__ void __launch_bounds__(kThreads, 2) contract_kernel(const ContractArgs args) {
    static_assert(!(MASKED && DENOM_IN_OUT), "the denominator comes either from the mask GEMM or from the output buffer");
    using L = SmemLayout<MASKED, KSTEPS>;
    constexpr int kStages = L::kStages;
    constexpr int kStageW = KSTEPS * kThreads;   // doubles of the A cache per (half) stage
    static_assert(NP >= 1 && NP <= (MASKED ? 4 : 8), "64 accumulator registers per thread");
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
    // release counters: a slot is refilled by whichever warp is the LAST to finish reading it, so
    // no warp ever blocks on "slot empty"
    int* release_cnt = reinterpret_cast<int*>(full_bar + kStages);

    if (args.run_if != nullptr && *args.run_if == 0) return;   // uniform for the whole grid
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tile_idx = blockIdx.x % args.tiles.n;
    const int ct_local = blockIdx.x / args.tiles.n;
    const int tile_t0 = args.tiles.t0[tile_idx];
    const int gran0 = tile_t0 / kGran;
    const int col0 = tile_t0 % kGran;   // 0, or 64 for the second masked half of a granule
    // the tile's list of stages with any non-zero weight (all of them for IDW; for Cressman only
    // the stations near the tile: the rest of the dense product is exactly zero and is skipped)
    const int* list = args.stage_list + (size_t)ct_local * args.n_stages * kSubStages;
    const int n_stages = args.n_active[ct_local];

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(&full_bar[s], 1);
            release_cnt[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int it = 0; it < kStages && it < n_stages; it++)
            produce_stage<MASKED, KSTEPS>(args, smem, full_bar, it, __ldg(list + it), gran0);
    }

    const int q = lane & 3;   // k index inside a k-step (A column / B row)
    const int r = lane >> 2;  // row of the 8x8 fragment (cell) / B column
    const long long ct = args.cell_tile0 + ct_local;
    const long long run = args.run_perm ? (long long)__ldg(args.run_perm + (size_t)ct * 8 + warp) : ct * 8 + warp;
    const long long cell = run * 8 + r;
    const double* wsrc = args.wfrag + (size_t)ct_local * args.n_stages * kWStageDoubles + threadIdx.x;

    double acc[2 * NP][2];
    double accm[MASKED ? 2 * NP : 1][2];
#pragma unroll
    for (int i = 0; i < 2 * NP; i++) {
        acc[i][0] = 0.0;
        acc[i][1] = 0.0;
        if (MASKED) {
            accm[i][0] = 0.0;
            accm[i][1] = 0.0;
        }
    }

    double w_cur[KSTEPS];
    if (n_stages > 0) {
        const double* w0 = wsrc + (size_t)__ldg(list) * kStageW;
#pragma unroll
        for (int kk = 0; kk < KSTEPS; kk++) w_cur[kk] = __ldg(w0 + kk * kThreads);
    }

    for (int it = 0; it < n_stages; it++) {
        const int slot = it % kStages;
        const int j = it / kStages;
        const double* zs = reinterpret_cast<const double*>(smem + slot * L::kStageBytes);

        // A fragments of the next stage: in flight underneath this stage's DMMAs
        double w_next[KSTEPS];
        {
            const int itn = (it + 1 < n_stages) ? it + 1 : it;
            const double* wn = wsrc + (size_t)__ldg(list + itn) * kStageW;
#pragma unroll
            for (int kk = 0; kk < KSTEPS; kk++) w_next[kk] = __ldg(wn + kk * kThreads);
        }
        mbar_wait(&full_bar[slot], j & 1);
#pragma unroll
        for (int kk = 0; kk < KSTEPS; kk++) {
            const double a = w_cur[kk];
            const double* zrow = zs + (kk * 4 + q) * kGranLd + col0 + 2 * r;
            const double* mrow = zrow + L::kPartDoubles;  // the mask part follows the z0 part
#pragma unroll
            for (int i = 0; i < NP; i++) {
                // one LDS.128 = the thread's B elements of two adjacent column groups -> two DMMAs
                const double2 b = *reinterpret_cast<const double2*>(zrow + 16 * i);
                dmma884(acc[2 * i][0], acc[2 * i][1], a, b.x);
                dmma884(acc[2 * i + 1][0], acc[2 * i + 1][1], a, b.y);
                if (MASKED) {
                    const double2 m = *reinterpret_cast<const double2*>(mrow + 16 * i);
                    dmma884(accm[2 * i][0], accm[2 * i][1], a, m.x);
                    dmma884(accm[2 * i + 1][0], accm[2 * i + 1][1], a, m.y);
                }
            }
        }
        // release the slot; the last warp to do so refills it with stage it + kStages
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(&release_cnt[slot], 1) == kConsumerWarps - 1) {
                release_cnt[slot] = 0;
                __threadfence_block();
                if (it + kStages < n_stages)
                    produce_stage<MASKED, KSTEPS>(args, smem, full_bar, it + kStages, __ldg(list + it + kStages), gran0);
            }
        }
#pragma unroll
        for (int kk = 0; kk < KSTEPS; kk++) w_cur[kk] = w_next[kk];
    }

    // ===== epilogue: normalise, NA rule, per-date early-outs, store =====
    const double na = __longlong_as_double((long long)kNaRealBits);
    if (cell >= args.n_cells) return;
    const double wsum = args.wsum[cell];
    const double inv_wsum = 1.0 / wsum;
    // thread holds dates tile_t0 + 16 i + 4 q + {0,1,2,3} for its cell:
    //   frag 2i   -> dates +0 (c0), +2 (c1);  frag 2i+1 -> dates +1 (c0), +3 (c1)
    // Running pointers and a tile-level bounds test: the per-element form (bounds test, 64-bit multiply) was ~25
    // dependent instructions per store (measured on cressman_kernel, where it was 72 % of the 25 km launch).
    const long long ld = args.ld_out;
    double* pe = args.out + (long long)(tile_t0 - args.t_lo + 4 * q) * ld + cell;
    const bool whole = tile_t0 >= args.t_lo && tile_t0 + 16 * NP <= args.t_hi;
#pragma unroll
    for (int i = 0; i < NP; i++) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int t = tile_t0 + 16 * i + 4 * q + e;
            const bool in_window = whole || (t >= args.t_lo && t < args.t_hi);
            const int f = 2 * i + (e & 1);
            const int c = e >> 1;
            const double numer = acc[f][c];
            double* optr = pe + e * ld;
            // DENOM_IN_OUT: the per-date denominator was left in the output buffer by
            // denom_umma_kernel (error-free int8 mask contraction) and is replaced by the result
            double denom = wsum;
            if (MASKED) denom = accm[f][c];
            else if (DENOM_IN_OUT) denom = in_window ? *optr : 1.0;
            // date-independent denominator: one true division per thread (inv_wsum), then multiplies
            double v = (MASKED || DENOM_IN_OUT) ? numer / denom : numer * inv_wsum;
            // R/IDW.R:332,339 ; R/Cressman.R:317 : !is.finite(v) | denom == 0 -> NA
            v = (isfinite(v) && denom != 0.0) ? v : na;
            if (args.date_flag != nullptr && in_window) {
                const uint8_t fl = args.date_flag[t];
                if (fl == 1) v = na;        // R/IDW.R:299-301 all-NA date
                else if (fl == 2) v = 0.0;  // R/IDW.R:294-297 all |obs| < 1e-12
                else if (fl == 3) v = args.date_const[t];   // R/Kriging_Ordinary.R:410-423 constant layer
            }
            // streaming store: the output is written once and never re-read by this kernel, so it must
            // not displace the L2-resident B operand and the A tiles shared by concurrent CTAs
            if (in_window) __stcs(optr, v);
        }
        pe += 16 * ld;
    }
}

// ---------------------------------------------------------------------------------------------
// Cressman: one persistent kernel per radius (weights, station lists and contraction fused).
//
// w = (R^2 - d^2)/(R^2 + d^2) vanishes beyond R, so a 64-cell tile (a compact ~8 x 8 km patch, see
// run_perm) only sees the few stations within R of one of its cells: 0.2 % (25 km) to 10 % (200 km) of
// the 2,000 stations of cfg 4.  A CTA owns a cell tile for ALL date tiles of the call:
//   prepare   (once per tile)  candidate 8-station groups by box distance (one group per thread), then
//             the exact per-station test "within R of any cell of the tile" (one station per thread),
//             compacted in ascending station order into the tile's station list; the A panel
//             [k-step][warp][lane] of exactly those stations and the row sums are generated from the
//             list — in the reference's own arithmetic (dist = sqrt(dx^2 + dy^2), dist_sq = dist * dist,
//             w = (R2 - dist_sq) / (R2 + dist_sq), w = 0 where dist > R; R/Cressman.R:279-300), so the
//             ill-conditioned weights next to the cut-off are the reference's bit for bit — and parked
//             in a per-CTA scratch that stays L2-resident (296 CTAs x <= 140 KB);
//   contract  the B operand of a stage is GATHERED: eight listed station rows (128 dates = 1056 B each,
//             contiguous in the granule-block layout) by 16-byte cp.async (LDGSTS), landing on the slot's `full`
//             mbarrier; the pipeline keeps running across the date tiles of the cell tile, so that the next
//             tile's rows arrive under the current tile's epilogue stores.  A slot is released by one
//             mbarrier.arrive per warp on its `empty` barrier and refilled by the warp whose turn it is
//             (rotating duty), kLag stages after its release.
// Cell tiles are handed out by an atomic counter (dynamic scheduling), their shape (8 x 8 ... 1 x 64 cells) is
// the host's choice per radius (run_perm).  Against the round-1 path (Morton-blob half stages, one CTA per
// (cell tile, date tile), one weights pass per radius) this executes 1.05x instead of 1.4-3x the in-range
// products, removes 450,000 CTA lifecycles per radius, and has no separate weights pass at all.
// ---------------------------------------------------------------------------------------------
struct CressmanArgs {
    const double2* cell_xy;      // [c_pad]
    const double2* st_xy;        // [>= s_pad] stations in Morton order
    const double4* group_box;    // [n_groups] bounding boxes of the 8-station groups
    const int* run_perm;         // [c_pad/8] compact cell tiles (8 Morton-consecutive runs of 8 cells); nullptr: 64 x 1 strips
    const double* z0;            // row of (granule g, station s) at ((g * s_pad + s) * kGranLd)
    double* out;                 // layer t at out + (t - t_lo) * ld_out
    long long ld_out, n_cells;
    int n_cell_tiles;
    int n_st, s_pad, n_groups;
    double r, r2;                // radius in metres and its square (R2 <- R * R on the host, as the reference)
    int t_lo, t_hi;
    double* panel;               // per CTA: list_cap * 64 doubles
    int* list;                   // per CTA: list_cap ints
    int list_cap;                // round_up(n_st, kCrListPad)
    unsigned long long* active_total;   // += 8-station stages executed per (cell tile)
    int* tile_counter;           // zeroed before the launch: cell tiles are handed out in order, one atomic each
    unsigned long long* prof;    // debugging (IPR_CR_PROF=1): cycles per phase summed over warps, else nullptr
    TileList tiles;
};

// stations per pipeline stage: 8 (KS = 2 k-steps, 8 slots) for the short lists of the small radii, where padding a
// list to a whole stage is a visible share of the work; 16 (KS = 4, 5 slots) for long lists, where the per-stage
// bookkeeping (barrier wait, slot release, refill) is what keeps the DMMA pipe from saturating
constexpr int kCrRowBytes = kGranLd * 8;               // 1056
constexpr int kCrListPad = 16;                         // list capacity granule (the larger stage)
#ifndef IPR_CR_PROFILE
#define IPR_CR_PROFILE 0   // 1: per-phase cycle counters in cressman_kernel (IPR_CR_PROF=1 prints them); costs registers
#endif
#ifndef IPR_CR_STAGES2
#define IPR_CR_STAGES2 12
#endif
#ifndef IPR_CR_STAGES4
#define IPR_CR_STAGES4 6
#endif
template <int KS>
struct CrLayout {
    static constexpr int kStages = (KS == 2) ? IPR_CR_STAGES2 : IPR_CR_STAGES4;   // ~100 KB in flight per CTA, two CTAs per SM
    static constexpr int kStageBytes = KS * 4 * kCrRowBytes;
    static constexpr int kBarOff = kStageBytes * kStages;
    static constexpr int kTotal = kBarOff + kStages * 8 + 16;
};

// the reference's weight, operation by operation (no FMA contraction)
__device__ __forceinline__ double cressman_weight_exact(double2 c, double2 s, double r, double r2) {
    const double dx = c.x - s.x, dy = c.y - s.y;
    const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));   // terra::distance, planar
    const double d2 = __dmul_rn(d, d);                                              // dist_sq <- dist_mat * dist_mat
    const double w = __ddiv_rn(__dsub_rn(r2, d2), __dadd_rn(r2, d2));
    return d > r ? 0.0 : w;                                                         // w[dist_mat > R] <- 0
}

// block-wide ordered compaction step: every thread passes `flag`; returns this thread's rank among the
// flagged threads of the block (ascending thread order) and the block total.  Two barriers.
__device__ __forceinline__ int block_rank(bool flag, int* s_warp_cnt, int& total) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_warp_cnt[warp] = __popc(m);
    __syncthreads();
    int before = 0, tot = 0;
#pragma unroll
    for (int w8 = 0; w8 < kConsumerWarps; w8++) {
        const int c = s_warp_cnt[w8];
        before += (w8 < warp) ? c : 0;
        tot += c;
    }
    __syncthreads();
    total = tot;
    return before + __popc(m & ((1u << lane) - 1u));
}

template <int NP, int KS>
__global__ void __launch_bounds__(kThreads, 2) cressman_kernel(const CressmanArgs args) {
    using L = CrLayout<KS>;
    constexpr int kCrKSteps = KS;
    constexpr int kCrRows = KS * 4;
    constexpr int kLag = L::kStages / 2;   // refill the slot released kLag stages ago: the warps drift apart by a
                                           // few stages, and the warp on duty should not have to wait for the slowest
    constexpr int kStages = L::kStages;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
    __shared__ uint64_t empty_bar[kStages];
    __shared__ double2 s_cell[kCellsPerCta];
    __shared__ double s_box[kConsumerWarps][4];
    __shared__ int s_warp_cnt[kConsumerWarps];
    // the tile's station list, kept next to the pipeline when it fits: the warp that refills a slot reads its
    // eight station ids from here — from L2 (the global copy) that load sat on the refill path, and since the
    // refiller is by construction the slowest warp its ~1 us set the pace of the whole CTA
    constexpr int kListSmem = 2048;
    __shared__ int s_list[kListSmem];
    // prepare-phase scratch aliases the (idle) pipeline slots
    int* s_cand = reinterpret_cast<int*>(smem);                 // candidate 8-station groups
    constexpr int kCandCap = 4096;                              // groups per round (32,768 stations)
    double2* s_stxy = reinterpret_cast<double2*>(smem + kCandCap * 4);   // listed stations' coordinates, by chunk
    constexpr int kStChunk = 2048;
    static_assert(kCandCap * 4 + kStChunk * 16 <= L::kBarOff, "prepare scratch must fit in the pipeline slots");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q = lane & 3;
    const int r = lane >> 2;
    double* panel = args.panel + (size_t)blockIdx.x * args.list_cap * kCellsPerCta + threadIdx.x;
    int* list = args.list + (size_t)blockIdx.x * args.list_cap;
    const double r2_margin = args.r2 * (1.0 + 1e-9);
    const double na = __longlong_as_double((long long)kNaRealBits);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(&full_bar[s], 32);   // one deferred arrival per lane of the refilling warp
            mbar_init(&empty_bar[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned it_base = 0;   // pipeline stages consumed so far by this CTA (slot and parity bookkeeping)

    long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    __shared__ int s_ct;
    for (;;) {
        // cell tiles are handed out dynamically (their cost varies with the station density around them; a static
        // round-robin left a tail of ~3 % at the end of the launch)
        if (threadIdx.x == 0) s_ct = atomicAdd(args.tile_counter, 1);
        __syncthreads();
        const int ct = s_ct;
        if (ct >= args.n_cell_tiles) break;
        long long pt0 = (IPR_CR_PROFILE && args.prof) ? clock64() : 0;
        // ================= prepare: station list, A panel, row sums =================
        const long long run = args.run_perm ? (long long)__ldg(args.run_perm + (size_t)ct * 8 + warp) : (long long)ct * 8 + warp;
        const long long cell = run * 8 + r;
        const double2 cxy = args.cell_xy[cell];
        if (q == 0) s_cell[warp * 8 + r] = cxy;
        double bx0 = cxy.x, bx1 = cxy.x, by0 = cxy.y, by1 = cxy.y;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bx0 = fmin(bx0, __shfl_xor_sync(0xffffffffu, bx0, o));
            bx1 = fmax(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
            by0 = fmin(by0, __shfl_xor_sync(0xffffffffu, by0, o));
            by1 = fmax(by1, __shfl_xor_sync(0xffffffffu, by1, o));
        }
        if (lane == 0) {
            s_box[warp][0] = bx0;
            s_box[warp][1] = bx1;
            s_box[warp][2] = by0;
            s_box[warp][3] = by1;
        }
        __syncthreads();
#pragma unroll
        for (int w8 = 0; w8 < kConsumerWarps; w8++) {
            bx0 = fmin(bx0, s_box[w8][0]);
            bx1 = fmax(bx1, s_box[w8][1]);
            by0 = fmin(by0, s_box[w8][2]);
            by1 = fmax(by1, s_box[w8][3]);
        }
        int n_list = 0;
        for (int g0 = 0; g0 < args.n_groups; g0 += kCandCap) {
            // candidate groups of this round: box-to-box distance, one group per thread
            int n_cand = 0;
            const int g1 = min(args.n_groups, g0 + kCandCap);
            for (int base = g0; base < g1; base += kThreads) {
                const int g = base + threadIdx.x;
                bool near = false;
                if (g < g1) {
                    const double4 b = args.group_box[g];
                    const double gx = fmax(0.0, fmax(b.x - bx1, bx0 - b.z));
                    const double gy = fmax(0.0, fmax(b.y - by1, by0 - b.w));
                    near = fma(gy, gy, gx * gx) <= r2_margin;
                }
                int tot;
                const int rank = block_rank(near, s_warp_cnt, tot);
                if (near) s_cand[n_cand + rank] = g;
                n_cand += tot;
            }
            __syncthreads();
            // exact test, one station per thread: within R of any cell of the tile
            for (int base = 0; base < n_cand * 8; base += kThreads) {
                const int i = base + threadIdx.x;
                int s = -1;
                bool in = false;
                if (i < n_cand * 8) {
                    s = s_cand[i >> 3] * 8 + (i & 7);
                    if (s < args.n_st) {
                        const double2 st = __ldg(&args.st_xy[s]);
                        double dmin = 1e300;
#pragma unroll 8
                        for (int c = 0; c < kCellsPerCta; c++) {
                            const double dx = s_cell[c].x - st.x, dy = s_cell[c].y - st.y;
                            dmin = fmin(dmin, fma(dy, dy, dx * dx));
                        }
                        in = dmin <= r2_margin;
                    }
                }
                int tot;
                const int rank = block_rank(in, s_warp_cnt, tot);
                if (in) list[n_list + rank] = s;
                n_list += tot;
            }
            __syncthreads();   // s_cand is rewritten by the next round
        }
        // pad to whole 8-station stages (position >= n_list carries weight 0; any valid row will do)
        const int n_pad = (n_list + kCrRows - 1) / kCrRows * kCrRows;
        if ((int)threadIdx.x < n_pad - n_list) list[n_list + threadIdx.x] = 0;
        __syncthreads();       // the list is read by other threads below (and by the producer lanes)
        const bool list_in_smem = n_pad <= kListSmem;
        if (list_in_smem)
            for (int i = threadIdx.x; i < n_pad; i += kThreads) s_list[i] = list[i];
        double wsum4[4] = {0.0, 0.0, 0.0, 0.0};
        for (int c0 = 0; c0 < n_pad; c0 += kStChunk) {
            const int nc = min(kStChunk, n_pad - c0);
            for (int i = threadIdx.x; i < nc; i += kThreads) s_stxy[i] = __ldg(&args.st_xy[list[c0 + i]]);
            __syncthreads();
#pragma unroll 2
            for (int k = 0; k < nc / 4; k++) {
                const int pos = c0 + 4 * k + q;
                double w = cressman_weight_exact(cxy, s_stxy[4 * k + q], args.r, args.r2);
                if (pos >= n_list) w = 0.0;
                panel[(size_t)(c0 / 4 + k) * kThreads] = w;
                wsum4[k & 3] += w;
            }
            __syncthreads();
        }
        double wsum = (wsum4[0] + wsum4[1]) + (wsum4[2] + wsum4[3]);
        wsum += __shfl_xor_sync(0xffffffffu, wsum, 1);
        wsum += __shfl_xor_sync(0xffffffffu, wsum, 2);
        const int ns = n_pad / kCrRows;                  // stages per date tile
        if (threadIdx.x == 0 && ns > 0) atomicAdd(args.active_total, (unsigned long long)ns);

        // ================= contraction over the date tiles, one continuous pipeline =================
        const int total = ns * args.tiles.n;
        // producer: the calling warp gathers the stage's eight station rows with 16-byte asynchronous copies
        // (LDGSTS): four lanes per row, 64 contiguous bytes per row and instruction.  (Eight 1 KB bulk copies per
        // stage starved the pipeline — ncu: 23 % of the warp samples spinning on the full barrier, DMMA pipe 70 %
        // active; the bulk-copy engine wants few large requests.)  Every lane's copies arrive on the slot's
        // barrier when they have landed.
        auto produce = [&](int jt, int i, int slot) {
#pragma unroll
            for (int r0 = 0; r0 < kCrRows; r0 += 8) {
                const int row = r0 + (lane >> 2);
                const int s = list_in_smem ? s_list[i * kCrRows + row] : list[i * kCrRows + row];
                const char* src = reinterpret_cast<const char*>(
                                      args.z0 + ((size_t)(args.tiles.t0[jt] / kGran) * args.s_pad + s) * kGranLd) + (lane & 3) * 16;
                const uint32_t dst = smem_u32(smem + slot * L::kStageBytes + row * kCrRowBytes) + (lane & 3) * 16;
#pragma unroll
                for (int j = 0; j < (kCrRowBytes / 16 + 3) / 4; j++) {
                    if ((lane & 3) + 4 * j < kCrRowBytes / 16)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + j * 64), "l"(src + j * 64) : "memory");
                }
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full_bar[slot])) : "memory");
        };
        if (warp == 0)
            for (int lin = 0; lin < kStages && lin < total; lin++)
                produce(lin / ns, lin % ns, (int)((it_base + (unsigned)lin) % kStages));
        // running pipeline state (no division or modulo in the stage loop): slot / parity of the current stage,
        // the (date tile, stage) that the refill duty of the current stage would gather, and whose turn it is
        int slot = (int)(it_base % kStages);
        unsigned par = (it_base / kStages) & 1u;
        int pj = (kStages - kLag) / max(ns, 1), pi = (kStages - kLag) % max(ns, 1);
        // stages until this warp's next turn at the refill duty (a turn before stage kLag has nothing to refill)
        int to_duty = (int)(((unsigned)kLag + (unsigned)warp - it_base) & (kConsumerWarps - 1));
        if (to_duty < kLag) to_duty += kConsumerWarps;

        double w_cur[kCrKSteps];
        if (ns > 0) {
#pragma unroll
            for (int kk = 0; kk < kCrKSteps; kk++) w_cur[kk] = panel[(size_t)kk * kThreads];
        }
        const double inv_wsum = 1.0 / wsum;
        if (IPR_CR_PROFILE && args.prof) { const long long t = clock64(); pf[0] += t - pt0; pt0 = t; }
        for (int jt = 0; jt < args.tiles.n; jt++) {
            const int tile_t0 = args.tiles.t0[jt];
            double acc[2 * NP][2];
#pragma unroll
            for (int i = 0; i < 2 * NP; i++) {
                acc[i][0] = 0.0;
                acc[i][1] = 0.0;
            }
            for (int i = 0; i < ns; i++) {
                const double* zs = reinterpret_cast<const double*>(smem + slot * L::kStageBytes);
                double w_next[kCrKSteps];
                {
                    // (two stages ahead was measured too: 88.9 against 87.2 ms per cfg-4 step)
                    const int in = (i + 1 < ns) ? i + 1 : 0;     // the first stage of the next date tile
#pragma unroll
                    for (int kk = 0; kk < kCrKSteps; kk++) w_next[kk] = panel[(size_t)(in * kCrKSteps + kk) * kThreads];
                }
                // refill duty, rotating over the warps: the warp whose turn it is waits until every warp has released
                // the slot of stage lin - kLag and gathers stage lin - kLag + kStages into it.  (The last warp to
                // release used to refill: the slowest warp got the extra work every time, fell further behind, and
                // the other SMSPs' DMMA pipes idled behind the 8-stage window — 14 % of the samples on this wait.)
                if (to_duty == 0 && pj < args.tiles.n) {
                    // the slot released kLag stages ago: kLag = kStages / 2 slots away, on the other parity if that wraps
                    const int slot_r = slot >= kLag ? slot - kLag : slot + (kStages - kLag);
                    const unsigned par_r = slot >= kLag ? par : par ^ 1u;
                    long long ta = (IPR_CR_PROFILE && args.prof) ? clock64() : 0;
                    mbar_wait(&empty_bar[slot_r], par_r);
                    if (IPR_CR_PROFILE && args.prof) { const long long t = clock64(); pf[6] += t - ta; ta = t; }
                    produce(pj, pi, slot_r);
                    if (IPR_CR_PROFILE && args.prof) { const long long t = clock64(); pf[7] += t - ta; }
                }
                to_duty = (to_duty == 0) ? kConsumerWarps - 1 : to_duty - 1;
                if (++pi == ns) {
                    pi = 0;
                    pj++;
                }
                if (IPR_CR_PROFILE && args.prof) { const long long t = clock64(); pf[1] += t - pt0; pt0 = t; }
                mbar_wait(&full_bar[slot], par);
                if (IPR_CR_PROFILE && args.prof) { const long long t = clock64(); pf[2] += t - pt0; pt0 = t; }
#pragma unroll
                for (int kk = 0; kk < kCrKSteps; kk++) {
                    const double a = w_cur[kk];
                    const double* zrow = zs + (kk * 4 + q) * kGranLd + 2 * r;
#pragma unroll
                    for (int n = 0; n < NP; n++) {
                        const double2 b = *reinterpret_cast<const double2*>(zrow + 16 * n);
                        dmma884(acc[2 * n][0], acc[2 * n][1], a, b.x);
                        dmma884(acc[2 * n + 1][0], acc[2 * n + 1][1], a, b.y);
                    }
                }
                __syncwarp();
                if (IPR_CR_PROFILE && args.prof) { const long long t = clock64(); pf[3] += t - pt0; pt0 = t; }
                if (lane == 0) mbar_arrive(&empty_bar[slot]);   // this warp's reads of the slot are done (release)
#pragma unroll
                for (int kk = 0; kk < kCrKSteps; kk++) w_cur[kk] = w_next[kk];
                if (++slot == kStages) {
                    slot = 0;
                    par ^= 1u;
                }
                if (IPR_CR_PROFILE && args.prof) { const long long t = clock64(); pf[4] += t - pt0; pt0 = t; }
            }
            // ---- epilogue of this date tile: normalise, NA rule (R/Cressman.R:316-317), store ----
            // A thread holds dates 16 n + 4 q + {0..3} of its cell.  Branch-free and on running pointers: with a
            // bounds test, a 64-bit multiply and a policy switch per element this was ~25 dependent instructions
            // per store, 6,000 cycles per date tile, three quarters of the 25 km kernel.
            if (cell < args.n_cells) {
                const long long ld = args.ld_out;
                double* pe = args.out + (long long)(tile_t0 - args.t_lo + 4 * q) * ld + cell;
                if (tile_t0 >= args.t_lo && tile_t0 + 16 * NP <= args.t_hi) {
#pragma unroll
                    for (int n = 0; n < NP; n++) {
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            double v = acc[2 * n + (e & 1)][e >> 1] * inv_wsum;
                            v = (isfinite(v) && wsum != 0.0) ? v : na;
                            __stcs(pe + e * ld, v);
                        }
                        pe += 16 * ld;
                    }
                } else {   // first / last tile of the window, or a ragged granule
#pragma unroll
                    for (int n = 0; n < NP; n++) {
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const int t = tile_t0 + 16 * n + 4 * q + e;
                            double v = acc[2 * n + (e & 1)][e >> 1] * inv_wsum;
                            v = (isfinite(v) && wsum != 0.0) ? v : na;
                            if (t >= args.t_lo && t < args.t_hi) __stcs(pe + e * ld, v);
                        }
                        pe += 16 * ld;
                    }
                }
            }
            if (IPR_CR_PROFILE && args.prof) { const long long t = clock64(); pf[5] += t - pt0; pt0 = t; }
        }
        it_base += (unsigned)total;
        __syncthreads();   // every warp is done with the slots, the list and the panel of this tile
    }
    if (IPR_CR_PROFILE && args.prof && lane == 0)
        for (int i = 0; i < 8; i++) atomicAdd(args.prof + i, (unsigned long long)pf[i]);
}

// ---------------------------------------------------------------------------------------------
// Kriging_Ordinary's per-date prelude (R/Kriging_Ordinary.R:406-423): fewer than two available stations ->
// the layer is their mean (or 0 when there is none); all available values equal -> that value.  One warp
// per date; flag 3 + the constant, else flag 0.
// ---------------------------------------------------------------------------------------------
__global__ void kriging_flags_kernel(const double* __restrict__ obs, long long ld_obs, int n_dates, int n_st,
                                     uint8_t* __restrict__ flag, double* __restrict__ value) {
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= n_dates) return;
    int cnt = 0;
    double lo = 1e308 * 10.0, hi = -1e308 * 10.0;   // +inf / -inf
    for (int s = lane; s < n_st; s += 32) {
        const double v = obs[(long long)s * ld_obs + t];
        if (v == v) {
            cnt++;
            lo = fmin(lo, v);
            hi = fmax(hi, v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) {
        const bool constant = (cnt < 2) || (lo == hi);
        flag[t] = constant ? 3 : 0;
        value[t] = (cnt == 0) ? 0.0 : lo;   // one station: its value is the mean; all equal: the value
    }
}

// ---------------------------------------------------------------------------------------------
// obs preparation: NA -> 0 / mask split into the padded station-major layout + per-date flags
//   block = (32 dates, 8 station stripes)
// ---------------------------------------------------------------------------------------------
__global__ void prep_obs_kernel(const double* __restrict__ obs, long long ld_obs, int n_dates, int n_st,
                                const int* __restrict__ perm, double* __restrict__ z0, double* __restrict__ mask, int n_stages,
                                uint8_t* __restrict__ date_flag, uint8_t* __restrict__ date_has_na,
                                int* __restrict__ nonfinite_count) {
    __shared__ int s_cnt[8][32];
    __shared__ int s_big[8][32];
    const int t = blockIdx.x * 32 + threadIdx.x;
    int cnt = 0, big = 0, bad = 0;
    if (t < n_dates) {
        const size_t gbase = (size_t)(t / kGran) * n_stages * kBlkDoubles + (t % kGran);
        for (int s = threadIdx.y; s < n_st; s += 8) {
            // row s of the B operand is the station that the spatial sort put at position s
            const double v = obs[(long long)perm[s] * ld_obs + t];
            const bool isna = (v != v);
            const size_t off = gbase + (size_t)(s / kKT) * kBlkDoubles + (s % kKT) * kGranLd;
            z0[off] = isna ? 0.0 : v;
            mask[off] = isna ? 0.0 : 1.0;
            if (!isna) {
                cnt++;
                if (!(fabs(v) < 1e-12)) big = 1;
                if (isinf(v)) bad++;
            }
        }
    }
    s_cnt[threadIdx.y][threadIdx.x] = cnt;
    s_big[threadIdx.y][threadIdx.x] = big;
    if (bad) atomicAdd(nonfinite_count, bad);
    __syncthreads();
    if (threadIdx.y == 0 && t < n_dates) {
        int c = 0, b = 0;
        for (int y = 0; y < 8; y++) {
            c += s_cnt[y][threadIdx.x];
            b |= s_big[y][threadIdx.x];
        }
        date_flag[t] = (c == 0) ? 1 : (b ? 0 : 2);
        date_has_na[t] = (c < n_st) ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// exact zero-distance (cell, station) pairs: bitwise coordinate equality (coordinates are
// canonicalised to +0 on upload), one thread per cell, stations broadcast from shared memory.
// ---------------------------------------------------------------------------------------------
__global__ void zero_pairs_kernel(const double2* __restrict__ cell_xy, long long n_cells,
                                  const double2* __restrict__ st_xy, int n_st, const int* __restrict__ perm,
                                  int2* __restrict__ pairs,
                                  int capacity, int* __restrict__ n_pairs) {
    __shared__ double2 s_st[256];
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double2 cxy = make_double2(0, 0);
    if (c < n_cells) cxy = cell_xy[c];
    const long long cxb = __double_as_longlong(cxy.x), cyb = __double_as_longlong(cxy.y);
    for (int s0 = 0; s0 < n_st; s0 += 256) {
        __syncthreads();
        if (s0 + (int)threadIdx.x < n_st) s_st[threadIdx.x] = st_xy[s0 + threadIdx.x];
        __syncthreads();
        const int n = min(256, n_st - s0);
        if (c < n_cells) {
            for (int k = 0; k < n; k++) {
                const double2 s = s_st[k];
                if (__double_as_longlong(s.x) == cxb && __double_as_longlong(s.y) == cyb) {
                    const int slot = atomicAdd(n_pairs, 1);
                    if (slot < capacity) pairs[slot] = make_int2((int)c, perm[s0 + k]);  // original station index
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// IDW zero-distance fix-up (R/IDW.R:312-326): a cell co-located with stations takes the running
// mean of their available observations, in station order; dates flagged by the early-outs and
// dates where none of the co-located stations reports keep the contraction's value.
//   fix_cell[i], fix_off[i]..fix_off[i+1] -> fix_st[] (ascending station index)
// ---------------------------------------------------------------------------------------------
__global__ void idw_zero_fix_kernel(const int* __restrict__ fix_cell, const int* __restrict__ fix_off,
                                    const int* __restrict__ fix_st, int fix0, int n_fix, const double* __restrict__ obs,
                                    long long ld_obs, const uint8_t* __restrict__ date_flag, int t_lo, int t_hi,
                                    double* __restrict__ out, long long ld_out) {
    const int t = t_lo + blockIdx.x * blockDim.x + threadIdx.x;
    const int i = fix0 + blockIdx.y;   // gridDim.y is capped at 65535: the host launches the list in slices
    if (t >= t_hi || i >= n_fix) return;
    if (date_flag[t] != 0) return;
    double pred = 0.0;
    int hit = 0;
    for (int k = fix_off[i]; k < fix_off[i + 1]; k++) {
        const double o = obs[(long long)fix_st[k] * ld_obs + t];
        if (o != o) continue;  // station not available on this date
        pred = (hit == 0) ? o : (pred * (double)hit + o) / (double)(hit + 1);
        hit++;
    }
    if (hit > 0) out[(long long)(t - t_lo) * ld_out + fix_cell[i]] = pred;
}

// ---------------------------------------------------------------------------------------------
// Sparse-NA IDW.  When a 128-date granule holds only a few missing observations, the second GEMM
// (weights x availability mask) is not worth its 2*S flop per cell-day: the granule is contracted
// unmasked (numerator / wsum) and the denominator of the few affected dates is corrected,
//     out = numer / (wsum - sum_{s NA on t} w(c,s)),
// with the missing stations' weights regenerated on the fly (~10 FP64 instructions each, i.e.
// 10 * NA-fraction of the cost of the masked GEMM).  Lists are built in station order by one warp
// per date, so the summation order is deterministic.
// ---------------------------------------------------------------------------------------------
constexpr int kNaCap = 256;   // most NA stations per date the correction handles

__global__ void na_lists_kernel(const double* __restrict__ obs, long long ld_obs, int n_dates, int n_st,
                                const int* __restrict__ perm, int* __restrict__ na_n, int* __restrict__ na_idx) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= n_dates) return;
    int count = 0;
    for (int s0 = 0; s0 < n_st; s0 += 32) {
        const int s = s0 + lane;
        bool isna = false;
        if (s < n_st) {
            const double v = obs[(long long)perm[s] * ld_obs + t];
            isna = (v != v);
        }
        const unsigned m = __ballot_sync(0xffffffffu, isna);
        if (isna) {
            const int pos = count + __popc(m & ((1u << lane) - 1u));
            if (pos < kNaCap) na_idx[(size_t)t * kNaCap + pos] = s;
        }
        count += __popc(m);
    }
    if (lane == 0) na_n[t] = count;
}

template <int WMODE>
__global__ void idw_sparse_na_fix_kernel(const double2* __restrict__ cell_xy, const double2* __restrict__ st_xy,
                                         int n_st, const double* __restrict__ wsum, long long n_cells,
                                         const int* __restrict__ na_n, const int* __restrict__ na_idx,
                                         const uint8_t* __restrict__ date_mode, const uint8_t* __restrict__ date_flag,
                                         int t_lo, int t_hi, WeightParams wp, double* __restrict__ out,
                                         long long ld_out) {
    __shared__ double2 s_na[kNaCap];
    __shared__ int s_idx[kNaCap];
    const int t = t_lo + blockIdx.y;
    if (t >= t_hi || date_mode[t] != 1 || date_flag[t] != 0) return;   // uniform per block
    const int n = na_n[t];
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const int s = na_idx[(size_t)t * kNaCap + k];
        s_idx[k] = s;
        s_na[k] = st_xy[s];
    }
    __syncthreads();
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const double2 cxy = cell_xy[c];
    double corr = 0.0;
    for (int k = 0; k < n; k++) {
        const double dx = cxy.x - s_na[k].x, dy = cxy.y - s_na[k].y;
        corr += weight_from_d2<WMODE>(fma(dy, dy, dx * dx), wp);
    }
    const double ws = wsum[c];
    double denom = ws - corr;
    if (corr > 0.99 * ws) {
        // the missing stations carry almost all of the weight: the subtraction would cancel, so sum
        // the weights of the available stations directly (rare: a missing station next to the cell)
        denom = 0.0;
        int k = 0;
        for (int s = 0; s < n_st; s++) {
            if (k < n && s_idx[k] == s) {
                k++;
                continue;
            }
            const double2 sc = st_xy[s];
            const double dx = cxy.x - sc.x, dy = cxy.y - sc.y;
            denom += weight_from_d2<WMODE>(fma(dy, dy, dx * dx), wp);
        }
    }
    double* ptr = out + (long long)(t - t_lo) * ld_out + c;
    double v = (*ptr) * ws / denom;     // numer = (numer / wsum) * wsum
    if (!isfinite(v) || denom == 0.0) v = __longlong_as_double((long long)kNaRealBits);
    *ptr = v;
}

// ---------------------------------------------------------------------------------------------
// Error-free int8 contraction of the availability mask (IDW with many NA) on the 5th-generation
// tensor cores: tcgen05.mma kind::i8, accumulators in TMEM.
//   denom[c,t] = sum_s w(c,s) m(s,t),  m in {0,1}
// is the second GEMM of the masked path.  m is exact in 8 bits, so the weights are cut into
// kQSlices 8-bit digits below a per-cell power-of-two scale, w = 2^e_c * sum_i d_i 256^-(i+1); each
// digit plane contracted with the byte mask gives exact int32 sums (at most S * 255 < 2^31 for the
// S <= 65 536 stations the path is enabled for) and the planes are recombined exactly in two int64
// limbs and one FP64 rounding.  Truncation below 2^(e_c-56) bounds the error by S * 2^-56 of the
// scale; (cell, date) pairs whose denominator is so small against the scale (< 2^-10: the cell's
// dominant station is missing) that this could exceed 1e-9 of it are listed and recomputed directly
// by idw_direct_fix_kernel (a list that overflows switches the launch group to the FP64 masked GEMM).
//
// One MMA computes D[128 dates x 224] += A[128 dates x 32 stations] * B[224 x 32 stations]^T with
// A = the byte mask of a granule and B = the seven digit planes of 32 cells (N = plane * 32 + cell):
// every plane reuses the same mask tile, which is read ONCE.  Measured on B200 (tools/umma_i8_probe.cu):
// 112 cycles per M128 N224 K32 MMA from shared memory = 4.54 POP/s (N <= 128 is bound at 80 cycles by
// the shared-memory read of A), against 1.14 POP/s for the legacy mma.sync IMMA path of round 1.
// Both operands live in HBM as the canonical no-swizzle K-major shared-memory image (core matrix =
// 8 rows x 16 bytes contiguous; SBO = 256 B between 8-row groups, LBO = 128 B between the two 16-byte
// K halves), so a pipeline stage is filled by two plain 1-D bulk copies and needs no tensor map.
// ---------------------------------------------------------------------------------------------
constexpr int kQSlices = 7;
constexpr int kQKb = 32;                                   // stations per integer k-block (K of kind::i8)
constexpr int kQKbPad = 4;                                 // n_kblocks is padded to a multiple of this
constexpr int kUCells = 32;                                // cells per MMA tile
constexpr int kUN = kQSlices * kUCells;                    // 224 = N of one MMA
constexpr int kUABytes = kGran * kQKb;                     // 4096: mask image per (granule, k-block)
constexpr int kUBBytes = kUN * kQKb;                       // 7168: digit image per (32-cell group, k-block)
constexpr int kUKbs = 2;                                   // k-blocks per pipeline stage
constexpr int kUStages = 8;
constexpr int kUStageBytes = kUKbs * (kUABytes + kUBBytes);   // 22 528
constexpr int kUThreads = 192;                             // warp 0 producer, warp 1 MMA issuer, warps 2-5 epilogue
constexpr int kUTmemCols = 512;                            // two accumulator buffers of 224 columns (at 0 and 256)
constexpr int kUBarBytes = (2 * kUStages + 4) * 8 + 16;
// epilogue staging: every epilogue warp transposes its 32 dates x 32 cells through shared memory (rows padded to 33
// doubles) so that a store instruction writes 32 consecutive cells of ONE layer (256 contiguous bytes) instead of
// 16 bytes in each of 32 layers
constexpr int kUStageLd = kUCells + 1;
constexpr int kUStagingOff = kUStages * kUStageBytes + ((kUBarBytes + 15) / 16) * 16;
constexpr int kUSmemBytes = kUStagingOff + 4 * 32 * kUStageLd * 8;
constexpr double kQDirectBelow = 0x1p-10;                  // denominators below this share of the scale are re-evaluated in FP64
static_assert(kQKbPad % kUKbs == 0, "whole pipeline stages");

struct DenomArgs {
    const unsigned char* wq;     // [n_cell_groups][n_kblocks][kUBBytes] digit planes (chunk-relative)
    const unsigned char* mask8;  // [n_gran][n_kblocks][kUABytes]
    const double* wmax;          // [c_pad]
    const uint8_t* date_flag;
    double* out;
    long long ld_out, n_cells;
    int cell_group0, n_cell_groups, n_kblocks;   // 32-cell groups; cell_group0 is absolute
    int t_lo, t_hi;
    int2* fb_list;               // (cell, date) pairs to recompute directly
    int* fb_count;
    int fb_cap;
    TileList tiles;              // granule start dates
};

// power-of-two scale of a cell's weights: 2^e > wmax (0 when the row has no weight at all)
__device__ __forceinline__ int row_exponent(double wmax) {
    return (wmax > 0.0 && isfinite(wmax)) ? ilogb(wmax) + 1 : 0;
}

// byte offset of (row, k) inside one k-block image of the canonical K-major layout
__host__ __device__ constexpr int umma_kmajor_offset(int row, int k) {
    return (row >> 3) * 256 + (k >> 4) * 128 + (row & 7) * 16 + (k & 15);
}

// digit planes of the weights as the B-operand image.  One CTA per 32-cell group; thread = (cell,
// part), the parts stride over the 16-station halves of the k-blocks (a 16-byte row of a core matrix).
template <int WMODE>
__global__ void __launch_bounds__(256) slice_weights_kernel(const WeightArgs args, const double* __restrict__ wmax,
                                                            unsigned char* __restrict__ wq, int n_kblocks, int n_st) {
    const int cl = threadIdx.x & 31;
    const int part = threadIdx.x >> 5;
    const long long cell = ((long long)args.cell_tile0 * (kCellsPerCta / kUCells) + blockIdx.x) * kUCells + cl;
    const double2 c = args.cell_xy[cell];
    const double sc = scalbn(1.0, -row_exponent(wmax[cell]));
    unsigned char* dst = wq + (size_t)blockIdx.x * n_kblocks * kUBBytes + umma_kmajor_offset(cl, 0);
    for (int hk = part; hk < 2 * n_kblocks; hk += 8) {
        // x = w / 2^e in [0, 1); its first 56 fraction bits, taken at once: x * 2^56 is exact and the
        // conversion truncates, exactly like peeling seven base-256 digits one by one
        unsigned long long u[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int s = hk * 16 + i;
            const double2 st = __ldg(&args.st_xy[s]);   // uniform across the warp: broadcast
            const double dx = c.x - st.x, dy = c.y - st.y;
            const double w = weight_from_d2<WMODE>(fma(dy, dy, dx * dx), args.wp) * sc;
            u[i] = (s < n_st) ? __double2ull_rz(w * 0x1p56) : 0ull;   // padded stations carry no weight
        }
#pragma unroll
        for (int sl = 0; sl < kQSlices; sl++) {
            unsigned r[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < 16; j++)
                r[j >> 2] |= (unsigned)((u[j] >> (48 - 8 * sl)) & 255ull) << (8 * (j & 3));
            // row n = sl * 32 + cl of the k-block image, K half hk & 1
            *reinterpret_cast<uint4*>(dst + (size_t)(hk >> 1) * kUBBytes + sl * (kUCells / 8) * 256 + (hk & 1) * 128) =
                make_uint4(r[0], r[1], r[2], r[3]);
        }
    }
}

// availability mask as bytes, A-operand image: one thread = 16 stations of one date (a core-matrix row)
__global__ void prep_mask8_kernel(const double* __restrict__ obs, long long ld_obs, int n_dates, int n_st,
                                  const int* __restrict__ perm, unsigned char* __restrict__ mask8, int n_kblocks,
                                  int n_gran) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n_gran * n_kblocks * 2 * kGran;
    if (idx >= total) return;
    const int tl = (int)(idx & (kGran - 1));      // date inside the granule: consecutive lanes, coalesced obs reads
    const int h = (int)((idx >> 7) & 1);
    const long long blk = idx >> 8;               // granule * n_kblocks + kb
    const int kb = (int)(blk % n_kblocks);
    const int gran = (int)(blk / n_kblocks);
    const int t = gran * kGran + tl;
    unsigned r[4] = {0u, 0u, 0u, 0u};
    if (t < n_dates) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int k = kb * kQKb + h * 16 + i;
            if (k < n_st) {
                const double v = obs[(long long)perm[k] * ld_obs + t];
                if (v == v) r[i >> 2] |= 1u << (8 * (i & 3));
            }
        }
    }
    *reinterpret_cast<uint4*>(mask8 + (size_t)blk * kUABytes + umma_kmajor_offset(tl, h * 16)) =
        make_uint4(r[0], r[1], r[2], r[3]);
}

// ---- tcgen05 helpers ----
// shared-memory matrix descriptor: no swizzle, K-major, version 1 (Blackwell); offsets in 16-byte units
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) |
           ((uint64_t)1 << 46);
}
// instruction descriptor of kind::i8: D = s32, A = B = u8, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_i8(int m, int n) {
    return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier when every MMA issued so far by this thread has completed (implies
// tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Persistent kernel, one CTA per SM.  Tile = (32-cell group, 128-date granule), granule fastest so that
// concurrently running CTAs share the group's digit planes through L2.  Warp 0 streams the operands
// through an 8-deep mbarrier pipeline, one lane of warp 1 issues the MMAs into one of two TMEM
// accumulator buffers, warps 2-5 (one per TMEM lane quarter) drain the other buffer: thread = date,
// 7 planes x 32 cells of int32 sums -> 32 denominators.
__global__ void __launch_bounds__(kUThreads, 1) denom_umma_kernel(const DenomArgs args) {
    extern __shared__ __align__(1024) unsigned char usm[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(usm + kUStages * kUStageBytes);
    uint64_t* empty_bar = full_bar + kUStages;
    uint64_t* tfull_bar = empty_bar + kUStages;    // accumulator buffer ready for the epilogue
    uint64_t* tempty_bar = tfull_bar + 2;          // accumulator buffer drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nks = args.n_kblocks / kUKbs;
    const int n_tiles = args.n_cell_groups * args.tiles.n;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kUStages; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kUTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int cg = tile / args.tiles.n;
                const int gran = args.tiles.t0[tile % args.tiles.n] / kGran;
                const unsigned char* a_src = args.mask8 + (size_t)gran * args.n_kblocks * kUABytes;
                const unsigned char* b_src = args.wq + (size_t)cg * args.n_kblocks * kUBBytes;
                for (int ks = 0; ks < nks; ks++, it++) {
                    const int s = it % kUStages;
                    mbar_wait(&empty_bar[s], ((it / kUStages) & 1) ^ 1);
                    unsigned char* stage = usm + s * kUStageBytes;
                    mbar_expect_tx(&full_bar[s], kUStageBytes);
                    bulk_g2s(stage, a_src + (size_t)ks * kUKbs * kUABytes, kUKbs * kUABytes, &full_bar[s]);
                    bulk_g2s(stage + kUKbs * kUABytes, b_src + (size_t)ks * kUKbs * kUBBytes, kUKbs * kUBBytes,
                             &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(kGran, kUN);
            int it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, tl++) {
                const int buf = tl & 1;
                mbar_wait(&tempty_bar[buf], ((tl >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tacc = tbase + (uint32_t)buf * 256u;
                for (int ks = 0; ks < nks; ks++, it++) {
                    const int s = it % kUStages;
                    mbar_wait(&full_bar[s], (it / kUStages) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(usm + s * kUStageBytes);
#pragma unroll
                    for (int u = 0; u < kUKbs; u++)
                        umma_i8(tacc, umma_smem_desc(sa + u * kUABytes), umma_smem_desc(sa + kUKbs * kUABytes + u * kUBBytes),
                                idesc, (ks | u) ? 1u : 0u);
                    umma_commit(&empty_bar[s]);   // the stage may be refilled once these MMAs have read it
                }
                umma_commit(&tfull_bar[buf]);
            }
        }
    } else {
        const int quarter = warp & 3;              // TMEM lanes [32 quarter, 32 quarter + 32) belong to this warp
        const int row = quarter * 32 + lane;       // date inside the granule
        double* stg = reinterpret_cast<double*>(usm + kUStagingOff) + quarter * 32 * kUStageLd;
        int tl = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, tl++) {
            const int buf = tl & 1;
            const int cg = tile / args.tiles.n;
            const int t = args.tiles.t0[tile % args.tiles.n] + row;
            const long long cell0 = (long long)(args.cell_group0 + cg) * kUCells;
            const double scale_l = scalbn(1.0, row_exponent(args.wmax[cell0 + lane]));   // lane <-> cell of the group
            const bool t_ok = (t >= args.t_lo && t < args.t_hi);
            const bool listable = t_ok && args.date_flag[t] == 0;
            mbar_wait(&tfull_bar[buf], (tl >> 1) & 1);
            tc_fence_after();
            const uint32_t tacc = tbase + ((uint32_t)(quarter * 32) << 16) + (uint32_t)buf * 256u;
#pragma unroll 1
            for (int c8 = 0; c8 < kUCells / 8; c8++) {
                uint32_t v[kQSlices][8];
#pragma unroll
                for (int p = 0; p < kQSlices; p++) tmem_ld8(tacc + p * kUCells + c8 * 8, v[p]);
                tmem_ld_wait();
                if (c8 == kUCells / 8 - 1) {
                    // last read of this accumulator buffer: hand it back to the MMA issuer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[buf]);
                }
                double d[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    // exact recombination: plane sums < 2^24, limbs < 2^41 and < 2^49, one rounding at the end
                    const long long hi = ((long long)v[0][i] << 16) + ((long long)v[1][i] << 8) + (long long)v[2][i];
                    const long long lo = ((long long)v[3][i] << 24) + ((long long)v[4][i] << 16) +
                                         ((long long)v[5][i] << 8) + (long long)v[6][i];
                    d[i] = fma((double)hi, 0x1p-24, (double)lo * 0x1p-56);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const double sc = __shfl_sync(0xffffffffu, scale_l, c8 * 8 + i);
                    const long long cell = cell0 + c8 * 8 + i;
                    // truncation error < S * 2^-56 of the scale <= 2^-40 for S <= 2^16 stations (the host
                    // disables this path beyond): guaranteed below 1e-9 of d only while d >= 2^-10
                    if (d[i] < kQDirectBelow && listable && cell < args.n_cells) {
                        const int slot = atomicAdd(args.fb_count, 1);
                        if (slot < args.fb_cap) args.fb_list[slot] = make_int2((int)cell, t);
                    }
                    d[i] *= sc;
                }
#pragma unroll
                for (int i = 0; i < 8; i++) stg[lane * kUStageLd + c8 * 8 + i] = d[i];
            }
            __syncwarp();
            // transposed write-out: lane <-> cell, one layer per instruction
            {
                const int t_first = t - lane;                       // date of TMEM lane 0 of this warp
                const bool cell_ok = cell0 + lane < args.n_cells;
                double* o = args.out + (long long)(t_first - args.t_lo) * args.ld_out + cell0 + lane;
#pragma unroll 8
                for (int rr = 0; rr < 32; rr++) {
                    const int tr = t_first + rr;
                    if (cell_ok && tr >= args.t_lo && tr < args.t_hi) o[(long long)rr * args.ld_out] = stg[rr * kUStageLd + lane];
                }
            }
            __syncwarp();   // the staging tile is rewritten by the next accumulator tile
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(kUTmemCols) : "memory");
    }
}

// direct FP64 evaluation of listed (cell, date) pairs: one warp each
template <int WMODE>
__global__ void idw_direct_fix_kernel(const int2* __restrict__ list, const int* __restrict__ count, int cap,
                                      int* __restrict__ overflow,
                                      const double2* __restrict__ cell_xy, const double2* __restrict__ st_xy,
                                      int n_st, const int* __restrict__ perm, const double* __restrict__ obs,
                                      long long ld_obs, WeightParams wp, int t_lo, double* __restrict__ out,
                                      long long ld_out) {
    // grid-stride over the list (usually empty: the launch is unconditional so that the host never
    // has to wait for the count); a list that overflowed raises the sticky flag instead
    if (*count > cap && threadIdx.x == 0 && blockIdx.x == 0) *overflow = 1;
    const int n = min(*count, cap);
    const int lane = threadIdx.x & 31;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += n_warps) {
    const int2 it = list[w];
    const double2 c = cell_xy[it.x];
    double num = 0.0, den = 0.0;
    for (int s = lane; s < n_st; s += 32) {
        const double z = obs[(long long)perm[s] * ld_obs + it.y];
        if (z != z) continue;
        const double2 sc = st_xy[s];
        const double dx = c.x - sc.x, dy = c.y - sc.y;
        const double wgt = weight_from_d2<WMODE>(fma(dy, dy, dx * dx), wp);
        num = fma(wgt, z, num);
        den += wgt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        num += __shfl_xor_sync(0xffffffffu, num, o);
        den += __shfl_xor_sync(0xffffffffu, den, o);
    }
    if (lane == 0) {
        double v = num / den;
        if (!isfinite(v) || den == 0.0) v = __longlong_as_double((long long)kNaRealBits);
        out[(long long)(it.y - t_lo) * ld_out + it.x] = v;
    }
    }
}

// ---------------------------------------------------------------------------------------------
// Optional post-pass on finished layers (the reference's steps right after the hot loop, which would
// otherwise cost host passes over the whole stack): rounding to n digits — mode 1 = R's round()
// as used by terra::app(x, function(v) round(v, n)) at R/IDW.R:356-358 (nearest of the floor/ceil
// candidates at 10^-n, ties to the even candidate); mode 2 = terra's round(SpatRaster, n) at
// R/Cressman.R:353-355 (std::round(v * 10^n) / 10^n) — and the polygon mask (cells whose centre is
// outside the study area become NA, R/IDW.R:376-381).  Elementwise, HBM-bound.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double round_r_digits(double x, double p10) {
    if (!isfinite(x) || x == 0.0) return x;
    const double ax = fabs(x);
    const double x10 = p10 * ax;
    const double i10 = floor(x10);
    const double xd = i10 / p10;
    const double xu = ceil(x10) / p10;
    const double du = xu - ax, dd = ax - xd;
    const bool even = (fmod(i10, 2.0) == 0.0);
    const double r = (dd < du || (dd == du && even)) ? xd : xu;
    return copysign(r, x);
}
__device__ __forceinline__ double round_terra_digits(double x, double p10) {
    if (!isfinite(x)) return x;
    return round(x * p10) / p10;
}

__global__ void post_kernel(double* __restrict__ out, long long ld_out, long long n_cells, int n_layers,
                            int round_mode, double p10, const uint8_t* __restrict__ cell_keep) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const bool keep = (cell_keep == nullptr) || (cell_keep[c] != 0);
    const double na = __longlong_as_double((long long)kNaRealBits);
    for (int t = blockIdx.y; t < n_layers; t += gridDim.y) {
        double* ptr = out + (long long)t * ld_out + c;
        if (!keep) {
            *ptr = na;
            continue;
        }
        if (round_mode == 0) continue;
        const double v = *ptr;
        *ptr = (round_mode == 1) ? round_r_digits(v, p10) : round_terra_digits(v, p10);
    }
}

// ---------------------------------------------------------------------------------------------
// Ordinary kriging: the kriging branch of perform_kriging, R/Kriging_Ordinary.R:428-497.
//
// The reference solves Gamma w = [gamma(d(c, .)); 1] for EVERY cell (solve() inside apply, :489) and predicts
// max(0, sum(w[1:n] * v)).  Gamma (the variogram matrix of the n stations available on the date, with the Lagrange
// row and column) does not depend on the cell and is symmetric, so
//     sum(w[1:n] v) = [v; 0]' Gamma^-1 [gamma(c); 1] = alpha' [gamma(c); 1],   alpha = Gamma^-1 [v; 0]:
// ONE (n+1)^2 factorisation and solve per date (kriging_solve_kernel, one CTA per date, LU with partial pivoting
// on [Gamma | v] in an L2-resident scratch), then a cell x station pass that evaluates the date's variogram model
// on the fly (kriging_predict_kernel).  The variogram parameters are fitted per date in R and differ from date to
// date, so there is no shared A operand and no GEMM here: both kernels run on the FP64 vector pipe.
// ---------------------------------------------------------------------------------------------
enum VariogramModel : int { kVarioExponential = 1, kVarioSpherical = 2, kVarioGaussian = 3, kVarioLinear = 4 };

// variogram_models(), R/Kriging_Ordinary.R:262-275
__device__ __forceinline__ double variogram_value(double h, double nugget, double sill, double range, int model) {
    switch (model) {
        case kVarioExponential: return nugget + sill * (1.0 - exp(-3.0 * h / range));
        case kVarioSpherical: {
            const double u = h / range;
            return h <= range ? nugget + sill * (1.5 * h / range - 0.5 * (u * u * u)) : nugget + sill;
        }
        case kVarioGaussian: {
            const double u = h / range;
            return nugget + sill * (1.0 - exp(-3.0 * (u * u)));
        }
        default: return h <= range ? nugget + sill * h / range : nugget + sill;
    }
}

struct KrigingArgs {
    const int* dates;            // [n_list] absolute dates to krige
    int n_list;
    const double* nugget;        // [n_dates], indexed by absolute date
    const double* sill;
    const double* range;
    const int* model;
    const uint8_t* kflag;        // [n_dates] 3: constant layer (prelude), never kriged
    const double* obs;           // column-major n_dates x n_st, input station order
    long long ld_obs;
    const int* perm;             // sorted position -> input station
    const double2* st_xy;        // stations in sorted order
    int n_st;
    double* scratch;             // per CTA: (n_st + 1) x (n_st + 2) doubles, column-major, ld = n_st + 1
    double* alpha;               // per list entry: n_st + 1 doubles (solution, packed over the available stations)
    int* avail;                  // per list entry: n_st ints (sorted positions of the available stations)
    int* meta;                   // per list entry: [0] n available, [1] status (0 solved, 1 skipped, 2 singular), [2..3] unused
    double* fallback_value;      // per list entry: mean of the available values (solve() failed, :494)
};

constexpr int kKrigThreads = 1024;
constexpr int kKrigNB = 16;         // panel width of the blocked elimination

__global__ void __launch_bounds__(kKrigThreads) kriging_solve_kernel(const KrigingArgs a) {
    __shared__ int s_n;
    __shared__ int s_warp_cnt[kKrigThreads / 32];
    __shared__ double s_red_v[kKrigThreads / 32];
    __shared__ int s_red_i[kKrigThreads / 32];
    __shared__ int s_piv;
    __shared__ double s_pivval;
    __shared__ double s_lt[kKrigNB * 128];   // L21 tile [p][row]
    __shared__ double s_ut[kKrigNB * 64];    // U12 tile [p][col]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = a.n_st + 1;
    double* G = a.scratch + (size_t)blockIdx.x * ld * (ld + 1);
    for (int li = blockIdx.x; li < a.n_list; li += gridDim.x) {
        const int t = a.dates[li];
        int* avail = a.avail + (size_t)li * a.n_st;
        int* meta = a.meta + (size_t)li * 4;
        double* alpha = a.alpha + (size_t)li * ld;
        if (a.kflag[t] == 3) {                       // constant layer of the prelude: nothing to solve
            if (tid == 0) meta[1] = 1;
            continue;
        }
        // ---- available stations, ascending sorted position (ordered block compaction) ----
        if (tid == 0) s_n = 0;
        __syncthreads();
        for (int base = 0; base < a.n_st; base += kKrigThreads) {
            const int s = base + tid;
            bool ok = false;
            if (s < a.n_st) {
                const double v = a.obs[(long long)a.perm[s] * a.ld_obs + t];
                ok = (v == v);
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (lane == 0) s_warp_cnt[warp] = __popc(m);
            __syncthreads();
            int before = s_n, tot = 0;
            for (int w = 0; w < kKrigThreads / 32; w++) {
                if (w < warp) before += s_warp_cnt[w];
                tot += s_warp_cnt[w];
            }
            if (ok) avail[before + __popc(m & ((1u << lane) - 1u))] = s;
            __syncthreads();
            if (tid == 0) s_n += tot;
            __syncthreads();
        }
        const int n = s_n;
        const int m1 = n + 1;                        // order of the system
        const double nug = a.nugget[t], sil = a.sill[t], rng = a.range[t];
        const int model = a.model[t];
        __syncthreads();
        // ---- [Gamma | b]: gamma(h_ij) off the diagonal, the nugget on it (matrix(nugget, n+1, n+1), :431-435),
        //      ones in the last row / column, 0 in the corner; b = [v; 0] ----
        for (long long e = tid; e < (long long)m1 * (m1 + 1); e += kKrigThreads) {
            const int i = (int)(e % m1), j = (int)(e / m1);
            double g;
            if (j == m1) {
                g = (i < n) ? a.obs[(long long)a.perm[avail[i]] * a.ld_obs + t] : 0.0;
            } else if (i == n || j == n) {
                g = (i == n && j == n) ? 0.0 : 1.0;
            } else if (i == j) {
                g = nug;
            } else {
                const double2 p = a.st_xy[avail[i]], q = a.st_xy[avail[j]];
                const double dx = p.x - q.x, dy = p.y - q.y;
                g = variogram_value(sqrt(dx * dx + dy * dy), nug, sil, rng, model);
            }
            G[(size_t)j * ld + i] = g;
        }
        __syncthreads();
        // ---- blocked Gaussian elimination with partial pivoting on the augmented matrix (right-looking, panels of
        //      kKrigNB columns).  Every entry outside the panel is updated by ONE ascending chain of FMAs starting
        //      from the entry itself — in the row-block solve (U12) and in the trailing update alike — so two
        //      identical rows (coincident stations) go through bit-identical operations and leave an exactly zero
        //      row behind, which is how the reference's solve() detects the singular system ----
        bool singular = false;
        const int ncols = m1 + 1;                    // columns of [Gamma | b]
        for (int k0 = 0; k0 < m1 && !singular; k0 += kKrigNB) {
            const int nb = min(kKrigNB, m1 - k0);
            // 1. panel: columns k0 .. k0+nb-1, updates confined to the panel, row swaps over all remaining columns
            for (int k = k0; k < k0 + nb; k++) {
                double best = -1.0;
                int bi = k;
                for (int i = k + tid; i < m1; i += kKrigThreads) {
                    const double v = fabs(G[(size_t)k * ld + i]);
                    if (v > best) {
                        best = v;
                        bi = i;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) {
                        best = ov;
                        bi = oi;
                    }
                }
                if (lane == 0) {
                    s_red_v[warp] = best;
                    s_red_i[warp] = bi;
                }
                __syncthreads();
                if (tid == 0) {
                    double bv = s_red_v[0];
                    int bb = s_red_i[0];
                    for (int w = 1; w < kKrigThreads / 32; w++)
                        if (s_red_v[w] > bv || (s_red_v[w] == bv && s_red_i[w] < bb)) {
                            bv = s_red_v[w];
                            bb = s_red_i[w];
                        }
                    s_piv = bb;
                    s_pivval = bv;
                }
                __syncthreads();
                if (!(s_pivval > 0.0)) {             // exactly singular: solve() raises, the reference takes the mean
                    singular = true;
                    break;
                }
                const int pr = s_piv;
                if (pr != k)
                    for (int j = k0 + tid; j < ncols; j += kKrigThreads) {
                        const double x = G[(size_t)j * ld + k];
                        G[(size_t)j * ld + k] = G[(size_t)j * ld + pr];
                        G[(size_t)j * ld + pr] = x;
                    }
                __syncthreads();
                // multipliers by true division (LAPACK's dgetf2): x / x is exactly 1, x * (1 / x) need not be
                const double piv = G[(size_t)k * ld + k];
                for (int i = k + 1 + tid; i < m1; i += kKrigThreads) G[(size_t)k * ld + i] /= piv;
                __syncthreads();
                // the rest of the panel: a warp per column, lanes down the rows
                for (int j = k + 1 + warp; j < k0 + nb; j += kKrigThreads / 32) {
                    double* col = G + (size_t)j * ld;
                    const double ukj = col[k];
                    const double* lcol = G + (size_t)k * ld;
                    for (int i = k + 1 + lane; i < m1; i += 32) col[i] = fma(-lcol[i], ukj, col[i]);
                }
                __syncthreads();
            }
            if (singular) break;
            const int c0 = k0 + nb;                  // first column / row beyond the panel
            if (c0 >= ncols) break;
            // 2. row block: U12 = L11^-1 A12 (unit lower triangular), one thread per column, ascending chains
            for (int j = c0 + tid; j < ncols; j += kKrigThreads) {
                double* col = G + (size_t)j * ld;
                for (int qq = 1; qq < nb; qq++) {
                    double val = col[k0 + qq];
                    for (int pp = 0; pp < qq; pp++) val = fma(-G[(size_t)(k0 + pp) * ld + k0 + qq], col[k0 + pp], val);
                    col[k0 + qq] = val;
                }
            }
            __syncthreads();
            // 3. trailing update A22 -= L21 U12 in 128-row x 64-column tiles: L21 and U12 tiles in shared memory,
            //    a thread owns rows tx, tx + 64 and four columns; the chain over the panel is ascending per entry
            const int tx = tid & 63, ty = tid >> 6;  // 64 x 16
            for (int r0 = c0; r0 < m1; r0 += 128)
                for (int j0 = c0; j0 < ncols; j0 += 64) {
                    __syncthreads();
                    for (int e = tid; e < nb * 128; e += kKrigThreads) {        // Ls[p][row]
                        const int pp = e >> 7, rr = e & 127;
                        s_lt[e] = (r0 + rr < m1) ? G[(size_t)(k0 + pp) * ld + r0 + rr] : 0.0;
                    }
                    for (int e = tid; e < nb * 64; e += kKrigThreads) {         // Us[p][col]
                        const int cc = e / nb, pp = e - cc * nb;                // consecutive threads walk down a column
                        s_ut[pp * 64 + cc] = (j0 + cc < ncols) ? G[(size_t)(j0 + cc) * ld + k0 + pp] : 0.0;
                    }
                    __syncthreads();
                    double v[2][4];
#pragma unroll
                    for (int a2 = 0; a2 < 2; a2++)
#pragma unroll
                        for (int b4 = 0; b4 < 4; b4++) {
                            const int i = r0 + tx + 64 * a2, j = j0 + 4 * ty + b4;
                            v[a2][b4] = (i < m1 && j < ncols) ? G[(size_t)j * ld + i] : 0.0;
                        }
                    for (int pp = 0; pp < nb; pp++) {
                        const double l0 = s_lt[pp * 128 + tx], l1 = s_lt[pp * 128 + tx + 64];
#pragma unroll
                        for (int b4 = 0; b4 < 4; b4++) {
                            const double u = s_ut[pp * 64 + 4 * ty + b4];
                            v[0][b4] = fma(-l0, u, v[0][b4]);
                            v[1][b4] = fma(-l1, u, v[1][b4]);
                        }
                    }
#pragma unroll
                    for (int a2 = 0; a2 < 2; a2++)
#pragma unroll
                        for (int b4 = 0; b4 < 4; b4++) {
                            const int i = r0 + tx + 64 * a2, j = j0 + 4 * ty + b4;
                            if (i < m1 && j < ncols) G[(size_t)j * ld + i] = v[a2][b4];
                        }
                }
            __syncthreads();
        }
        if (singular) {
            // mean(available_values), :494 — sequential sum in station order like R's mean()
            if (tid == 0) {
                double sum = 0.0;
                for (int i = 0; i < n; i++) sum += a.obs[(long long)a.perm[avail[i]] * a.ld_obs + t];
                a.fallback_value[li] = sum / (double)n;
                meta[0] = n;
                meta[1] = 2;
            }
            __syncthreads();
            continue;
        }
        // ---- back substitution (column-oriented: no reductions) ----
        double* b = G + (size_t)m1 * ld;
        for (int k = m1 - 1; k >= 0; k--) {
            if (tid == 0) b[k] /= G[(size_t)k * ld + k];
            __syncthreads();
            const double xk = b[k];
            for (int i = tid; i < k; i += kKrigThreads) b[i] = fma(-G[(size_t)k * ld + i], xk, b[i]);
            __syncthreads();
        }
        for (int i = tid; i < m1; i += kKrigThreads) alpha[i] = b[i];
        if (tid == 0) {
            meta[0] = n;
            meta[1] = 0;
        }
        __syncthreads();
    }
}

// out[c, t] = max(0, alpha_n + sum_i alpha_i * gamma_t(d(c, s_i)))   (:478-491); a date whose system was singular
// gets mean(available_values) (:494); constant-layer dates are left as the fallback pass wrote them.
__global__ void __launch_bounds__(256) kriging_predict_kernel(const KrigingArgs a, const double2* __restrict__ cell_xy,
                                                               long long n_cells, int t_lo, double* __restrict__ out,
                                                               long long ld_out) {
    constexpr int kTile = 256;
    __shared__ double2 s_xy[kTile];
    __shared__ double s_al[kTile];
    const int li = blockIdx.y;
    const int t = a.dates[li];
    const int* meta = a.meta + (size_t)li * 4;
    const int status = meta[1];
    if (status == 1) return;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double* optr = out + (long long)(t - t_lo) * ld_out + c;
    if (status == 2) {
        if (c < n_cells) *optr = a.fallback_value[li];
        return;
    }
    const int n = meta[0];
    const int ld = a.n_st + 1;
    const double* alpha = a.alpha + (size_t)li * ld;
    const int* avail = a.avail + (size_t)li * a.n_st;
    const double nug = a.nugget[t], sil = a.sill[t], rng = a.range[t];
    const int model = a.model[t];
    const double2 cxy = (c < n_cells) ? cell_xy[c] : make_double2(0.0, 0.0);
    double acc = 0.0;
    for (int s0 = 0; s0 < n; s0 += kTile) {
        const int m = min(kTile, n - s0);
        __syncthreads();
        if ((int)threadIdx.x < m) {
            s_xy[threadIdx.x] = a.st_xy[avail[s0 + threadIdx.x]];
            s_al[threadIdx.x] = alpha[s0 + threadIdx.x];
        }
        __syncthreads();
        for (int i = 0; i < m; i++) {
            const double dx = cxy.x - s_xy[i].x, dy = cxy.y - s_xy[i].y;
            acc = fma(s_al[i], variogram_value(sqrt(dx * dx + dy * dy), nug, sil, rng, model), acc);
        }
    }
    if (c < n_cells) *optr = fmax(0.0, acc + alpha[n]);
}

}  // namespace ipr
