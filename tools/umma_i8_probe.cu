// Probe for the Blackwell-native integer tensor path: tcgen05.mma kind::i8 (u8 x u8 -> s32, accumulators
// in TMEM), operands staged in shared memory in the canonical no-swizzle K-major layout by plain 1-D bulk
// copies (the operands are pre-arranged in HBM as the shared-memory image).
//
//   1. correctness of the shared-memory / instruction descriptors and of the TMEM accumulator layout
//      (D row = TMEM lane, D column = TMEM column) against a host reference, for the two candidate
//      core-matrix orders and both readings of the LBO / SBO fields;
//   2. issue rate of M=128 x N x K=32 MMAs from shared memory for several N (cycles per MMA), which
//      sizes the mask-GEMM kernel of the engine (csrc/ipr_kernels.cuh, denom_umma_kernel).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/umma_i8_probe tools/umma_i8_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at line %d: %s\n", cudaGetErrorName(e_), __LINE__, cudaGetErrorString(e_)); \
            exit(2);                                                                            \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared-memory matrix descriptor, no swizzle, K-major: core matrix = 8 rows x 16 bytes (128 B contiguous)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version 1 (Blackwell)
    return d;
}
// instruction descriptor: kind::i8, u8 x u8 -> s32, both operands K-major
__host__ __device__ inline uint32_t idesc_i8(int M, int N) {
    uint32_t d = 0;
    d |= 2u << 4;                    // c_format = S32
    d |= 0u << 7;                    // a_format = U8
    d |= 0u << 10;                   // b_format = U8
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct ProbeArgs {
    const unsigned char* a;   // [n_kb][128 x 32 bytes] canonical image
    const unsigned char* b;   // [n_kb][N x 32 bytes] canonical image
    int* d;                   // [128][N] row-major result
    int N, n_kb;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    int reps;                 // > 0: timing mode, issue `reps` rounds of the n_kb MMAs
    long long* cycles;
};

// one CTA, 6 warps: warp 0 loads + issues, warp 1 allocates TMEM, warps 2..5 read the accumulator back
__global__ void __launch_bounds__(192, 1) probe_kernel(const ProbeArgs args) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full_bar, done_bar;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = args.N, n_kb = args.n_kb;
    const uint32_t a_bytes = 128 * 32, b_bytes = (uint32_t)N * 32;
    unsigned char* sa = smem;
    unsigned char* sb = smem + (size_t)n_kb * a_bytes;
    if (threadIdx.x == 0) {
        mbar_init(&full_bar, 1);
        mbar_init(&done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base;
    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&full_bar, n_kb * (a_bytes + b_bytes));
            bulk_g2s(sa, args.a, n_kb * a_bytes, &full_bar);
            bulk_g2s(sb, args.b, n_kb * b_bytes, &full_bar);
            mbar_wait(&full_bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = idesc_i8(128, N);
            const int rounds = args.reps > 0 ? args.reps : 1;
            const long long c0 = clock64();
            for (int rep = 0; rep < rounds; rep++)
                for (int kb = 0; kb < n_kb; kb++) {
                    const uint64_t ad = smem_desc(smem_u32(sa + (size_t)kb * a_bytes), args.a_lbo, args.a_sbo);
                    const uint64_t bd = smem_desc(smem_u32(sb + (size_t)kb * b_bytes), args.b_lbo, args.b_sbo);
                    umma_i8(tbase, ad, bd, idesc, (rep | kb) ? 1u : 0u);
                }
            umma_commit(&done_bar);
            mbar_wait(&done_bar, 0);
            const long long c1 = clock64();
            if (args.cycles) args.cycles[blockIdx.x] = c1 - c0;
        }
        __syncwarp();
    }
    if (warp >= 2) {
        mbar_wait(&done_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3;             // the TMEM lane quarter this warp may read
        const int row = quarter * 32 + lane;
        if (args.reps == 0 && blockIdx.x == 0)
            for (int c0 = 0; c0 < N; c0 += 8) {
                uint32_t r[8];
                tmem_ld8(tbase + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
                for (int i = 0; i < 8; i++) args.d[(size_t)row * N + c0 + i] = (int)r[i];
            }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256) : "memory");
    }
}

// host packing of a [rows][n_kb * 32] u8 matrix (row-major, K contiguous) into the canonical image
//   order 0: per k-block [row group j][k-half h][8 rows][16 B]   (SBO = 256, LBO = 128)
//   order 1: per k-block [k-half h][row group j][8 rows][16 B]   (SBO = 128, LBO = rows/8 * 128)
static std::vector<unsigned char> pack(const std::vector<unsigned char>& m, int rows, int n_kb, int order) {
    std::vector<unsigned char> out((size_t)rows * n_kb * 32);
    for (int kb = 0; kb < n_kb; kb++)
        for (int r = 0; r < rows; r++)
            for (int k = 0; k < 32; k++) {
                const int j = r / 8, rr = r % 8, h = k / 16, b = k % 16;
                size_t off = (size_t)kb * rows * 32;
                if (order == 0) off += (size_t)j * 256 + h * 128 + rr * 16 + b;
                else off += (size_t)h * (rows / 8) * 128 + (size_t)j * 128 + rr * 16 + b;
                out[off] = m[(size_t)r * n_kb * 32 + kb * 32 + k];
            }
    return out;
}

int main(int argc, char** argv) {
    const int timing_only = (argc > 1 && atoi(argv[1]) == 1);
    int dev_count = 0;
    CK(cudaGetDeviceCount(&dev_count));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (!timing_only) {
        for (int N : {224, 64, 128, 256}) {
            const int n_kb = 3, K = n_kb * 32;
            std::vector<unsigned char> A((size_t)128 * K), B((size_t)N * K);
            srand(1234 + N);
            for (auto& v : A) v = (unsigned char)(rand() & 1);
            for (auto& v : B) v = (unsigned char)(rand() & 255);
            std::vector<int> ref((size_t)128 * N, 0);
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < N; n++) {
                    int s = 0;
                    for (int k = 0; k < K; k++) s += (int)A[(size_t)m * K + k] * (int)B[(size_t)n * K + k];
                    ref[(size_t)m * N + n] = s;
                }
            for (int order = 0; order < 2; order++)
                for (int swap = 0; swap < 2; swap++) {
                    auto pa = pack(A, 128, n_kb, order), pb = pack(B, N, n_kb, order);
                    unsigned char *da, *db;
                    int* dd;
                    CK(cudaMalloc(&da, pa.size()));
                    CK(cudaMalloc(&db, pb.size()));
                    CK(cudaMalloc(&dd, sizeof(int) * 128 * N));
                    CK(cudaMemcpy(da, pa.data(), pa.size(), cudaMemcpyHostToDevice));
                    CK(cudaMemcpy(db, pb.data(), pb.size(), cudaMemcpyHostToDevice));
                    CK(cudaMemset(dd, 0xFF, sizeof(int) * 128 * N));
                    ProbeArgs pa_{};
                    pa_.a = da;
                    pa_.b = db;
                    pa_.d = dd;
                    pa_.N = N;
                    pa_.n_kb = n_kb;
                    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
                    if (order == 0) {
                        a_lbo = 128; a_sbo = 256; b_lbo = 128; b_sbo = 256;
                    } else {
                        a_lbo = 16 * 128; a_sbo = 128; b_lbo = (N / 8) * 128; b_sbo = 128;
                    }
                    if (swap) {
                        std::swap(a_lbo, a_sbo);
                        std::swap(b_lbo, b_sbo);
                    }
                    pa_.a_lbo = a_lbo; pa_.a_sbo = a_sbo; pa_.b_lbo = b_lbo; pa_.b_sbo = b_sbo;
                    pa_.reps = 0;
                    pa_.cycles = nullptr;
                    probe_kernel<<<1, 192, n_kb * (128 + N) * 32 + 1024>>>(pa_);
                    CK(cudaGetLastError());
                    CK(cudaDeviceSynchronize());
                    std::vector<int> got((size_t)128 * N);
                    CK(cudaMemcpy(got.data(), dd, sizeof(int) * 128 * N, cudaMemcpyDeviceToHost));
                    long long bad = 0;
                    for (size_t i = 0; i < got.size(); i++) bad += (got[i] != ref[i]);
                    printf("N=%3d order=%d swap=%d  a(lbo=%u,sbo=%u) b(lbo=%u,sbo=%u): %lld / %zu mismatches  [d00=%d ref=%d, d(5,7)=%d ref=%d]\n",
                           N, order, swap, a_lbo, a_sbo, b_lbo, b_sbo, bad, got.size(), got[0], ref[0], got[5 * N + 7],
                           ref[5 * N + 7]);
                    cudaFree(da);
                    cudaFree(db);
                    cudaFree(dd);
                }
        }
    }
    // issue rate: all SMs, each CTA re-issues the same n_kb MMAs `reps` times
    for (int N : {64, 128, 224, 256}) {
        const int n_kb = 8, reps = 2000;
        unsigned char *da, *db;
        long long* dc;
        CK(cudaMalloc(&da, (size_t)n_kb * 128 * 32));
        CK(cudaMalloc(&db, (size_t)n_kb * N * 32));
        CK(cudaMalloc(&dc, sizeof(long long) * 148));
        CK(cudaMemset(da, 1, (size_t)n_kb * 128 * 32));
        CK(cudaMemset(db, 0, (size_t)n_kb * N * 32));
        ProbeArgs p{};
        p.a = da; p.b = db; p.d = nullptr; p.N = N; p.n_kb = n_kb;
        p.a_lbo = 128; p.a_sbo = 256; p.b_lbo = 128; p.b_sbo = 256;
        p.reps = reps;
        p.cycles = dc;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        probe_kernel<<<148, 192, n_kb * (128 + N) * 32 + 1024>>>(p);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        probe_kernel<<<148, 192, n_kb * (128 + N) * 32 + 1024>>>(p);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        long long cyc[148];
        CK(cudaMemcpy(cyc, dc, sizeof(cyc), cudaMemcpyDeviceToHost));
        const double n_mma = (double)reps * n_kb;
        const double ops = 2.0 * 128 * N * 32 * n_mma * 148;
        printf("rate N=%3d: %.1f cycles/MMA (SM 0), %.3f ms, %.0f TOP/s dense int8\n", N, (double)cyc[0] / n_mma, ms,
               ops / (ms * 1e-3) * 1e-12);
        cudaFree(da);
        cudaFree(db);
        cudaFree(dc);
    }
    return 0;
}
