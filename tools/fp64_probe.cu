// FP64 pipe probe for B200 (sm_100a): measures DFMA, DMMA (m8n8k4 / m16n8k8 / m16n8k16)
// throughput and reciprocal variants, so the roofline denominator for the IDW/Cressman
// contraction is a measured number (MEASURED_PEAKS.json has no FP64 entry).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// NACC independent 8x8 accumulator fragments per warp; a/b fragments reused (register operands only)
template <int NACC>
__global__ void k_dmma884(double* out, int iters, double a0, double b0) {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c0[i] = threadIdx.x; c1[i] = i; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma1688(double* out, int iters, double a0, double b0) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    double a[4] = {a0, a0 + 1e-9, a0 + 2e-9, a0 + threadIdx.x * 1e-9};
    double b[2] = {b0, b0 * 0.5};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma1688(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma16816(double* out, int iters, double a0, double b0) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = a0 + (i + threadIdx.x) * 1e-9;
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = b0 * (1 + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA + DFMA co-issue: does weight generation (DFMA) steal DMMA slots, i.e. are they one pipe?
template <int NACC, int NF>
__global__ void k_mix(double* out, int iters, double a0, double b0) {
    double c0[NACC], c1[NACC], f[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c0[i] = threadIdx.x; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < NF; i++) f[i] = threadIdx.x + i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c0[i], c1[i], a, b);
#pragma unroll
        for (int i = 0; i < NF; i++) f[i] = fma(f[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < NF; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// legacy warp-level tensor paths (mma.sync): headroom for error-free low-precision emulation of
// the 0/1 mask contraction (INT8 / BF16 slices with exact accumulation)
template <int NACC>
__global__ void k_imma(int* out, int iters, int a0, int b0) {
    int c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    unsigned a[4] = {(unsigned)a0, (unsigned)a0 + 1, (unsigned)a0 + 2, (unsigned)a0 + threadIdx.x};
    unsigned b[2] = {(unsigned)b0, (unsigned)b0 * 3};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_hmma_bf16(float* out, int iters, unsigned a0, unsigned b0) {
    float c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    unsigned a[4] = {a0, a0 + 1, a0 + 2, a0 + threadIdx.x};
    unsigned b[2] = {b0, b0 * 3};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// do DMMA and IMMA overlap when issued by different warps of one SM? (fused masked-IDW design)
__global__ void k_dmma_imma(double* out, int iters_d, int iters_i, int imma_warps) {
    const int warp = threadIdx.x >> 5;
    if (warp < 8) {
        double c0[16], c1[16];
#pragma unroll
        for (int i = 0; i < 16; i++) { c0[i] = threadIdx.x; c1[i] = i; }
        double a = 1.0000001 + threadIdx.x * 1e-9, b = 1e-9;
        for (int it = 0; it < iters_d; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) dmma884(c0[i], c1[i], a, b);
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) s += c0[i] + c1[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (warp < 8 + imma_warps) {
        int c[8][4];
#pragma unroll
        for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
        unsigned a[4] = {3u, 4u, 5u, threadIdx.x}, b[2] = {5u, 15u};
        for (int it = 0; it < iters_i; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
        int s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

// reciprocal variants -------------------------------------------------------
__device__ __forceinline__ double rcp_mufu2(double x) {  // MUFU.RCP64H seed + 2 Newton steps
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ double rcp_mufu3(double x) {
    double r = rcp_mufu2(x);
    double e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double rcp_f32seed(double x) {  // f32 rcp seed + 2 Newton steps
    float xf = (float)x;
    float rf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(xf));
    double r = (double)rf;
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
template <int MODE>
__global__ void k_rcp(double* out, const double* in, int n, int reps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = in[i], acc = 0;
    for (int r = 0; r < reps; r++) {
        double y;
        if (MODE == 0) y = 1.0 / x;
        else if (MODE == 1) y = __drcp_rn(x);
        else if (MODE == 2) y = rcp_mufu2(x);
        else if (MODE == 3) y = rcp_mufu3(x);
        else y = rcp_f32seed(x);
        acc += y;
        x += 1.0;
    }
    out[i] = acc;
}

template <typename F>
static float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s sm_%d%d SMs %d clock %d kHz\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.clockRate);
    int nsm = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 16 * 1024));
    const int iters = 20000;
    // ---- DFMA
    for (int wps : {4, 8, 16, 32}) {
        int threads = 256, blocks = nsm * wps * 32 / threads;
        float ms = time_ms([&] { k_dfma<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double flops = 2.0 * 16 * iters * (double)blocks * threads;
        printf("DFMA nacc=16 warps/SM=%2d : %.3f ms  %.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
    }
    // ---- DMMA m8n8k4
    for (int wps : {4, 8, 16, 32}) {
        int threads = 128, blocks = nsm * wps * 32 / threads;
        float ms = time_ms([&] { k_dmma884<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double flops = 2.0 * 256 * 16 * iters * (double)blocks * threads / 32;
        printf("DMMA m8n8k4 nacc=16 warps/SM=%2d : %.3f ms  %.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
    }
    for (int wps : {8}) {
        int threads = 128, blocks = nsm * wps * 32 / threads;
        float ms = time_ms([&] { k_dmma884<32><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double flops = 2.0 * 256 * 32 * iters * (double)blocks * threads / 32;
        printf("DMMA m8n8k4 nacc=32 warps/SM=%2d : %.3f ms  %.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
        ms = time_ms([&] { k_dmma884<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        flops = 2.0 * 256 * 4 * iters * (double)blocks * threads / 32;
        printf("DMMA m8n8k4 nacc=4  warps/SM=%2d : %.3f ms  %.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
        ms = time_ms([&] { k_dmma884<1><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        flops = 2.0 * 256 * 1 * iters * (double)blocks * threads / 32;
        printf("DMMA m8n8k4 nacc=1  warps/SM=%2d : %.3f ms  %.2f TFLOP/s (latency-bound: %.1f ns/dmma)\n", wps, ms, flops / ms * 1e-9, ms * 1e6 / iters);
    }
    for (int wps : {4, 8, 16}) {
        int threads = 128, blocks = nsm * wps * 32 / threads;
        float ms = time_ms([&] { k_dmma1688<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        double flops = 2.0 * 16 * 8 * 8 * 8 * iters * (double)blocks * threads / 32;
        printf("DMMA m16n8k8 nacc=8 warps/SM=%2d : %.3f ms  %.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
        ms = time_ms([&] { k_dmma16816<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        flops = 2.0 * 16 * 8 * 16 * 8 * iters * (double)blocks * threads / 32;
        printf("DMMA m16n8k16 nacc=8 warps/SM=%2d : %.3f ms  %.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
    }
    // ---- mix: 16 DMMA + NF DFMA per iteration
    {
        int wps = 8, threads = 128, blocks = nsm * wps * 32 / threads;
        float ms0 = time_ms([&] { k_mix<16, 0><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        float ms8 = time_ms([&] { k_mix<16, 8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        float ms16 = time_ms([&] { k_mix<16, 16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        float ms32 = time_ms([&] { k_mix<16, 32><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("MIX 16 DMMA + {0,8,16,32} DFMA per iter, 8 warps/SM: %.3f %.3f %.3f %.3f ms\n", ms0, ms8, ms16, ms32);
    }
    // ---- legacy mma.sync INT8 / BF16 (warp-level tensor path)
    for (int wps : {8, 16, 32}) {
        int threads = 128, blocks = nsm * wps * 32 / threads;
        float ms = time_ms([&] { k_imma<8><<<blocks, threads>>>((int*)out, iters, 3, 5); });
        double ops = 2.0 * 16 * 8 * 32 * 8 * iters * (double)blocks * threads / 32;
        printf("IMMA m16n8k32 s8 nacc=8 warps/SM=%2d : %.3f ms  %.1f TOP/s\n", wps, ms, ops / ms * 1e-9);
        ms = time_ms([&] { k_hmma_bf16<8><<<blocks, threads>>>((float*)out, iters, 0x3f803f80u, 0x3f803f80u); });
        ops = 2.0 * 16 * 8 * 16 * 8 * iters * (double)blocks * threads / 32;
        printf("HMMA m16n8k16 bf16 nacc=8 warps/SM=%2d : %.3f ms  %.1f TFLOP/s\n", wps, ms, ops / ms * 1e-9);
    }
    // ---- DMMA (8 warps) with and without IMMA (4 warps) on the same SM
    {
        const int it_d = 20000;
        // IMMA work sized to ~25 % of the DMMA time when run alone: 16 DMMA/iter * 16 cyc * 2 warps per SMSP = 512 cyc/iter;
        // 8 IMMA/iter * 8.4 cyc * 1 warp per SMSP = 67 cyc/iter -> iters_i = 0.25 * 512 / 67 * it_d
        const int it_i = (int)(0.25 * 512.0 / 67.0 * it_d);
        float t_d = time_ms([&] { k_dmma_imma<<<nsm, 384>>>(out, it_d, 0, 0); });
        float t_i = time_ms([&] { k_dmma_imma<<<nsm, 384>>>(out, 0, it_i, 4); });
        float t_b = time_ms([&] { k_dmma_imma<<<nsm, 384>>>(out, it_d, it_i, 4); });
        printf("DMMA+IMMA co-issue (8+4 warps/SM): DMMA alone %.3f ms, IMMA alone %.3f ms, both %.3f ms (sum %.3f)\n",
               t_d, t_i, t_b, t_d + t_i);
    }
    // ---- reciprocal variants: accuracy and cost
    {
        int n = nsm * 2048;
        double* h = (double*)malloc(sizeof(double) * n);
        srand(1);
        for (int i = 0; i < n; i++) h[i] = exp(log(1e2) + (log(1e13) - log(1e2)) * (rand() / (double)RAND_MAX));
        double* din; CK(cudaMalloc(&din, sizeof(double) * n));
        CK(cudaMemcpy(din, h, sizeof(double) * n, cudaMemcpyHostToDevice));
        double* ho = (double*)malloc(sizeof(double) * n);
        const char* names[5] = {"1.0/x", "__drcp_rn", "mufu+2newton", "mufu+3newton", "f32seed+2newton"};
        for (int mode = 0; mode < 5; mode++) {
            auto launch = [&](int reps) {
                switch (mode) {
                    case 0: k_rcp<0><<<n / 256, 256>>>(out, din, n, reps); break;
                    case 1: k_rcp<1><<<n / 256, 256>>>(out, din, n, reps); break;
                    case 2: k_rcp<2><<<n / 256, 256>>>(out, din, n, reps); break;
                    case 3: k_rcp<3><<<n / 256, 256>>>(out, din, n, reps); break;
                    default: k_rcp<4><<<n / 256, 256>>>(out, din, n, reps); break;
                }
            };
            launch(1); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(ho, out, sizeof(double) * n, cudaMemcpyDeviceToHost));
            double maxrel = 0;
            for (int i = 0; i < n; i++) { double ref = 1.0 / h[i]; double r = fabs(ho[i] - ref) / ref; if (r > maxrel) maxrel = r; }
            float ms = time_ms([&] { launch(2000); });
            printf("RCP %-16s max rel err %.3e  %.3f ms for %d x 2000 (%.2f G rcp/s)\n", names[mode], maxrel, ms, n, n * 2000.0 / ms * 1e-6);
        }
    }
    // ---- pinned host <-> device bandwidth
    {
        size_t bytes = (size_t)1 << 30;
        void *hp, *dp; CK(cudaMallocHost(&hp, bytes)); CK(cudaMalloc(&dp, bytes));
        float h2d = time_ms([&] { CK(cudaMemcpyAsync(dp, hp, bytes, cudaMemcpyHostToDevice)); }, 3);
        float d2h = time_ms([&] { CK(cudaMemcpyAsync(hp, dp, bytes, cudaMemcpyDeviceToHost)); }, 3);
        printf("PCIe pinned 1 GiB: H2D %.1f GB/s  D2H %.1f GB/s\n", bytes / h2d * 1e-6, bytes / d2h * 1e-6);
    }
    return 0;
}
