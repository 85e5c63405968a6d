// interpolater_b200 — sm_100a kernels for the IDW / Cressman cell x station x date contraction.
//
// Math (reference: R/IDW.R:260-342, R/Cressman.R:279-321 — restated in oracle/reference_port.py):
//   IDW       out[c,t] = sum_s w(c,s) z0[s,t] / sum_s w(c,s) m[s,t],  w = d(c,s)^-p
//   Cressman  out[c,t] = sum_s w(c,s) z0[s,t] / sum_s w(c,s),         w = (R^2-d^2)/(R^2+d^2), 0 if d > R
// i.e. a dense FP64 GEMM  [cells x stations] * [stations x dates]  whose A operand (the weights)
// is never materialised: every thread generates its A fragment element in registers from the cell
// and station coordinates, directly in the DMMA.8x8x4 register layout
// (a[row = lane/4][k = lane%4]), and feeds it to 2*NP accumulators fragments covering 16*NP dates.
//
// B operand (z0 / mask rows, station-major, dates contiguous) is streamed L2 -> shared memory
// with cp.async.bulk (TMA bulk engine, UBLKCP) through a kStages-deep mbarrier full/empty
// pipeline; the refill of a stage is issued by one warp per iteration, in rotation.  Shared rows are padded by 4 doubles so that the
// LDS.128 B-fragment reads (4 stations x 2 adjacent dates per quarter-warp) are conflict-free.
//
// CTA = 8 warps (8 cells each); tile = 64 cells x 16*NP dates;
// accumulators 8*NP doubles/thread (NP=16 -> 128 registers).  One CTA per SM.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ipr {

constexpr int kConsumerWarps = 8;
constexpr int kThreads = kConsumerWarps * 32;
constexpr int kCellsPerCta = kConsumerWarps * 8;     // 64
constexpr int kKT = 16;                              // stations per pipeline stage
constexpr int kStages = 4;
constexpr int kMaxTilesPerLaunch = 56;
constexpr double kPadCoord = 1e150;                  // padded stations: d^2 ~ 2e300 -> weight ~ 0
constexpr unsigned long long kNaRealBits = 0x7FF00000000007A2ull;

enum WeightMode : int {
    kIdwP2 = 0,       // w = 1/d^2
    kIdwEvenP = 1,    // w = (1/d^2)^n, p = 2n
    kIdwGenericP = 2, // w = pow(d^2, -p/2)
    kCressman = 3     // w = (R^2 - d^2)/(R^2 + d^2), 0 when d^2 > R^2
};

struct WeightParams {
    double a;   // kIdwGenericP: -p/2 ; kCressman: R^2
    int n;      // kIdwEvenP: p/2
};

struct TileList {
    int n;
    int t0[kMaxTilesPerLaunch];   // absolute start date of each tile (multiple of 16)
};

struct ContractArgs {
    const double2* cell_xy;   // [C_pad]
    const double2* st_xy;     // [S_pad + kKT]
    const double* z0;         // [S_pad][ldz]
    const double* mask;       // [S_pad][ldz] (masked variants only)
    const uint8_t* date_flag; // [>= ldz] 0 normal, 1 all-NA layer, 2 exact-zero layer (IDW only)
    double* out;              // layer t at out + (t - t_lo) * ld_out
    long long ld_out;
    long long n_cells;
    int s_pad;
    int ldz;
    int t_lo, t_hi;           // store window [t_lo, t_hi)
    WeightParams wp;
    TileList tiles;
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
        if (WMODE == kIdwP2) {
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

template <int NP, bool MASKED>
struct SmemLayout {
    static constexpr int kW = 16 * NP;                // dates per tile
    static constexpr int kLd = kW + 4;                // padded row (doubles)
    static constexpr int kRowBytes = kW * 8;
    static constexpr int kZBytes = kKT * kLd * 8;
    static constexpr int kCoordOff = kZBytes * (MASKED ? 2 : 1);
    static constexpr int kStageBytes = kCoordOff + kKT * 16;
    static constexpr int kBarOff = kStageBytes * kStages;
    static constexpr int kTotal = kBarOff + 2 * kStages * 8;
    static constexpr uint32_t kTxBytes = kKT * kRowBytes * (MASKED ? 2 : 1) + kKT * 16;
};


// Refill one pipeline stage: z0 rows by lanes 0..15, mask rows by lanes 16..31, and the station
// coordinates of the FOLLOWING stage (they travel one stage ahead of their z rows so that the
// next stage's weights can be generated underneath this stage's DMMAs).  Called by a whole warp.
template <int NP, bool MASKED>
__device__ __forceinline__ void produce_stage(const ContractArgs& args, unsigned char* smem, uint64_t* full_bar,
                                              uint64_t* empty_bar, int it, int tile_t0, int lane) {
    using L = SmemLayout<NP, MASKED>;
    const int slot = it % kStages;
    const int j = it / kStages;
    unsigned char* stage = smem + slot * L::kStageBytes;
    if (lane == 0) {
        mbar_wait(&empty_bar[slot], (j & 1) ^ 1);
        mbar_expect_tx(&full_bar[slot], L::kTxBytes);
    }
    __syncwarp();
    const int row = lane & 15;
    const size_t goff = (size_t)(it * kKT + row) * args.ldz + tile_t0;
    if (lane < 16) {
        bulk_g2s(stage + row * L::kLd * 8, args.z0 + goff, L::kRowBytes, &full_bar[slot]);
    } else if (MASKED) {
        bulk_g2s(stage + L::kZBytes + row * L::kLd * 8, args.mask + goff, L::kRowBytes, &full_bar[slot]);
    }
    if (lane == 0) {
        bulk_g2s(stage + L::kCoordOff, args.st_xy + (size_t)(it + 1) * kKT, kKT * 16, &full_bar[slot]);
    }
}

// ---------------------------------------------------------------------------------------------
// Fused weight-generation + DMMA contraction + normalise epilogue
//   NP     : number of 16-date column groups per tile (tile width 16*NP dates)
//   MASKED : also contract the availability mask (IDW with NA in the tile: 4*S flop/cell-day)
//   WMODE  : weight function
// ---------------------------------------------------------------------------------------------
template <int NP, bool MASKED, int WMODE>
__global__ void __launch_bounds__(kThreads, 1) contract_kernel(const ContractArgs args) {
    using L = SmemLayout<NP, MASKED>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
    uint64_t* empty_bar = full_bar + kStages;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tile_t0 = args.tiles.t0[blockIdx.y];
    const int n_stages = args.s_pad / kKT;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // prologue: the first kStages-1 stages
    if (warp == 0) {
        for (int it = 0; it < kStages - 1 && it < n_stages; it++)
            produce_stage<NP, MASKED>(args, smem, full_bar, empty_bar, it, tile_t0, lane);
    }

    const int q = lane & 3;   // k index inside a k-step (A column / B row)
    const int r = lane >> 2;  // row of the 8x8 fragment (cell) / B column
    const long long cell = (long long)blockIdx.x * kCellsPerCta + warp * 8 + r;
    const double2 cxy = args.cell_xy[cell];

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
    double wsum = 0.0;

    double w_cur[kKT / 4];
#pragma unroll
    for (int kk = 0; kk < kKT / 4; kk++) {
        const double2 s = args.st_xy[kk * 4 + q];
        const double dx = cxy.x - s.x, dy = cxy.y - s.y;
        w_cur[kk] = weight_from_d2<WMODE>(fma(dy, dy, dx * dx), args.wp);
    }

    for (int it = 0; it < n_stages; it++) {
        const int slot = it % kStages;
        const int j = it / kStages;
        const unsigned char* stage = smem + slot * L::kStageBytes;
        const double* zs = reinterpret_cast<const double*>(stage);
        const double* ms = reinterpret_cast<const double*>(stage + L::kZBytes);
        const double2* cs = reinterpret_cast<const double2*>(stage + L::kCoordOff);
        {
            // refill the slot freed by iteration it-1 with stage it+kStages-1; warps take turns
            const int pit = it + kStages - 1;
            if (pit < n_stages && warp == (it & (kConsumerWarps - 1)))
                produce_stage<NP, MASKED>(args, smem, full_bar, empty_bar, pit, tile_t0, lane);
        }
        mbar_wait(&full_bar[slot], j & 1);

        double w_next[kKT / 4];
#pragma unroll
        for (int kk = 0; kk < kKT / 4; kk++) {
            const double2 s = cs[kk * 4 + q];
            const double dx = cxy.x - s.x, dy = cxy.y - s.y;
            w_next[kk] = weight_from_d2<WMODE>(fma(dy, dy, dx * dx), args.wp);
        }
#pragma unroll
        for (int kk = 0; kk < kKT / 4; kk++) {
            const double a = w_cur[kk];
            wsum += a;
            const double* zrow = zs + (kk * 4 + q) * L::kLd + 2 * r;
            const double* mrow = ms + (kk * 4 + q) * L::kLd + 2 * r;
#pragma unroll
            for (int i = 0; i < NP; i++) {
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
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[slot]);
#pragma unroll
        for (int kk = 0; kk < kKT / 4; kk++) w_cur[kk] = w_next[kk];
    }

    // ===== epilogue: normalise, NA rule, per-date early-outs, store =====
    // lanes of a quad share the cell row; their partial weight sums cover disjoint stations
    wsum += __shfl_xor_sync(0xffffffffu, wsum, 1);
    wsum += __shfl_xor_sync(0xffffffffu, wsum, 2);
    const double na = __longlong_as_double((long long)kNaRealBits);
    if (cell >= args.n_cells) return;
#pragma unroll
    for (int i = 0; i < NP; i++) {
        // thread holds dates tile_t0 + 16 i + 4 q + {0,1,2,3} for its cell:
        //   frag 2i   -> dates +0 (c0), +2 (c1);  frag 2i+1 -> dates +1 (c0), +3 (c1)
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int t = tile_t0 + 16 * i + 4 * q + e;
            if (t < args.t_lo || t >= args.t_hi) continue;
            const int f = 2 * i + (e & 1);
            const int c = e >> 1;
            const double numer = acc[f][c];
            const double denom = MASKED ? accm[f][c] : wsum;
            double v = numer / denom;
            // R/IDW.R:332,339 ; R/Cressman.R:317 : !is.finite(v) | denom == 0 -> NA
            if (!isfinite(v) || denom == 0.0) v = na;
            if (WMODE != kCressman) {
                const uint8_t fl = args.date_flag[t];
                if (fl == 1) v = na;        // R/IDW.R:299-301 all-NA date
                else if (fl == 2) v = 0.0;  // R/IDW.R:294-297 all |obs| < 1e-12
            }
            args.out[(long long)(t - args.t_lo) * args.ld_out + cell] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// obs preparation: NA -> 0 / mask split into the padded station-major layout + per-date flags
//   block = (32 dates, 8 station stripes)
// ---------------------------------------------------------------------------------------------
__global__ void prep_obs_kernel(const double* __restrict__ obs, long long ld_obs, int n_dates, int n_st,
                                double* __restrict__ z0, double* __restrict__ mask, int ldz,
                                uint8_t* __restrict__ date_flag, uint8_t* __restrict__ date_has_na,
                                int* __restrict__ nonfinite_count) {
    __shared__ int s_cnt[8][32];
    __shared__ int s_big[8][32];
    const int t = blockIdx.x * 32 + threadIdx.x;
    int cnt = 0, big = 0, bad = 0;
    if (t < n_dates) {
        for (int s = threadIdx.y; s < n_st; s += 8) {
            const double v = obs[(long long)s * ld_obs + t];
            const bool isna = (v != v);
            z0[(long long)s * ldz + t] = isna ? 0.0 : v;
            mask[(long long)s * ldz + t] = isna ? 0.0 : 1.0;
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
                                  const double2* __restrict__ st_xy, int n_st, int2* __restrict__ pairs,
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
                    if (slot < capacity) pairs[slot] = make_int2((int)c, s0 + k);
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
                                    const int* __restrict__ fix_st, int n_fix, const double* __restrict__ obs,
                                    long long ld_obs, const uint8_t* __restrict__ date_flag, int t_lo, int t_hi,
                                    double* __restrict__ out, long long ld_out) {
    const int t = t_lo + blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
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

}  // namespace ipr
