// What does the store pattern of a cell x date tile cost by itself?  (nvcc -arch=sm_100a -O3 -o store_pattern_probe ...)
// Writes a 3,650-layer x 1,000,000-cell FP64 stack (29.2 GB) tile by tile like cressman_kernel's epilogue — 296 CTAs
// of 8 warps, each CTA walking the 29 date tiles (128 dates) of its 64-cell tile, streaming 8-byte stores, a store
// instruction = 8 cells x 4 dates — for different shapes of the 64-cell tile and two warp mappings:
//   mapping 0: a warp owns 8 cells x 128 dates            (the kernel's: two warps share a 128-byte line)
//   mapping 1: a warp owns 16 cells x 64 dates            (both halves of a line from the same warp, back to back)
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 2) probe(double* out, long long ld, int ncol, int rows_per_tile, int mapping,
                                                int n_tiles, int n_dt, int* counter) {
    __shared__ int s_ct;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane & 3, r = lane >> 2;
    const int cells_per_row = 64 / rows_per_tile;            // 64, 32, 16 or 8
    const int tiles_per_row = ncol / cells_per_row;
    for (;;) {
        if (threadIdx.x == 0) s_ct = atomicAdd(counter, 1);
        __syncthreads();
        const int ct = s_ct;
        __syncthreads();
        if (ct >= n_tiles) break;
        const int trow = (ct / tiles_per_row) * rows_per_tile, tcol = (ct % tiles_per_row) * cells_per_row;
        for (int jt = 0; jt < n_dt; jt++) {
            if (mapping == 0) {
                const int cell_in_tile = warp * 8 + r;       // 8 runs of 8 cells, row-major inside the tile
                const long long cell = (long long)(trow + cell_in_tile / cells_per_row) * ncol + tcol + cell_in_tile % cells_per_row;
                double* p = out + (long long)(jt * 128 + 4 * q) * ld + cell;
#pragma unroll
                for (int n = 0; n < 8; n++) {
#pragma unroll
                    for (int e = 0; e < 4; e++) __stcs(p + e * ld, 1.0 + n);
                    p += 16 * ld;
                }
            } else {
                const int pair = warp & 3, half = warp >> 2; // 4 pairs of runs (16 cells), 2 date halves
#pragma unroll
                for (int n = 0; n < 4; n++) {
#pragma unroll
                    for (int e = 0; e < 4; e++) {
#pragma unroll
                        for (int mb = 0; mb < 2; mb++) {
                            const int cell_in_tile = pair * 16 + mb * 8 + r;
                            const long long cell = (long long)(trow + cell_in_tile / cells_per_row) * ncol + tcol + cell_in_tile % cells_per_row;
                            __stcs(out + (long long)(jt * 128 + half * 64 + 16 * n + 4 * q + e) * ld + cell, 1.0 + n);
                        }
                    }
                }
            }
        }
    }
}

int main() {
    const int ncol = 1000, nrow = 1000, D = 3650, n_dt = 28;   // 28 full date tiles (3,584 dates)
    const long long C = (long long)ncol * nrow;
    double* out;
    int* counter;
    cudaMalloc(&out, sizeof(double) * C * D);
    cudaMalloc(&counter, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int shapes[4] = {1, 2, 4, 8};   // rows per tile: 1 x 64, 2 x 32, 4 x 16, 8 x 8 (ncol must divide evenly: use 960 cols)
    for (int mapping = 0; mapping < 2; mapping++)
        for (int si = 0; si < 4; si++) {
            const int rpt = shapes[si];
            if (mapping == 1 && 64 / rpt < 16) continue;
            const int use_col = 960;                         // multiple of 64
            const int n_tiles = (nrow / rpt) * (use_col / (64 / rpt));
            float best = 1e9f;
            for (int it = 0; it < 3; it++) {
                cudaMemset(counter, 0, 4);
                cudaEventRecord(e0);
                probe<<<296, 256>>>(out, C, use_col, rpt, mapping, n_tiles, n_dt, counter);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const double bytes = 8.0 * use_col * nrow * 128.0 * n_dt;
            printf("mapping %d, tile %d rows x %2d cells (%3d-byte pieces): %.2f ms, %.0f GB/s  (%s)\n", mapping, rpt, 64 / rpt,
                   64 / rpt * 8, best, bytes / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
