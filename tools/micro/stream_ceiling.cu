// Micro-benchmark: how fast can ONE B200 stream 768 MB through the same TMA ring the kernels use,
// with a trivial body (sum), and with plain coalesced loads?  Gives the ceiling for the roofline
// discussion in DESIGN.md.    nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_ceiling stream_ceiling.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "../../fitnoise_b200/csrc/fn_tile.cuh"
using namespace fn;

__global__ void __launch_bounds__(256) ring_kernel(TileSrc src, int stages, long long n_tiles, double* out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[FN_MAX_STAGES];
    const TileLayout L = tile_layout(src);
    tile_prologue(smem_raw, bars, stages, src, L, n_tiles);
    __syncthreads();
    double acc = 0.0;
    const int per = src.genes * src.m;
    tile_loop(smem_raw, bars, stages, src, L, n_tiles, false, [&](unsigned char* st, long long tile) {
        const double* y = (const double*)st;
        for (int i = threadIdx.x; i < per; i += blockDim.x) acc += y[i];
    });
    if (acc == 12345.678) out[0] = acc;
}

__global__ void __launch_bounds__(256) ldg_kernel(const double2* y, long long n2, double* out) {
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 v = __ldg(&y[i]);
        acc += v.x + v.y;
    }
    if (acc == 12345.678) out[0] = acc;
}

int main() {
    const long long n = 1000000, n_pad = 1000448; const int m = 96;
    double *y, *out; uint8_t* status;
    cudaMalloc(&y, n_pad * m * 8); cudaMalloc(&out, 8); cudaMalloc(&status, n_pad);
    cudaMemset(y, 0, n_pad * m * 8); cudaMemset(status, 1, n_pad);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = (double)n * (m * 8 + 1);
    int genes_opts[] = {32, 64, 128};
    for (int gi = 0; gi < 3; ++gi) for (int stages = 2; stages <= 4; ++stages) for (int ctas = 1; ctas <= 4; ++ctas) {
        const int genes = genes_opts[gi];
        TileSrc src; src.y = y; src.aux = nullptr; src.gs = nullptr; src.cst = nullptr; src.status = status; src.m = m; src.genes = genes; src.row_stride = m; src.l2_keep = 0;
        const size_t smem = tile_stage_bytes(genes, m, false, false, false) * stages;
        if (smem * ctas > 220000) continue;
        cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const long long n_tiles = (n + genes - 1) / genes;
        const int grid = 148 * ctas;
        float best = 1e9;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            ring_kernel<<<grid, 256, smem>>>(src, stages, n_tiles, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 1 && ms < best) best = ms;
        }
        printf("ring genes %3d stages %d ctas %d : %.4f ms %.0f GB/s  (%s)\n", genes, stages, ctas, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    for (int mult = 2; mult <= 16; mult *= 2) {
        float best = 1e9;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            ldg_kernel<<<148 * mult, 256>>>((const double2*)y, n * m / 2, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 1 && ms < best) best = ms;
        }
        printf("ldg  ctas/SM %2d : %.4f ms %.0f GB/s\n", mult, best, (double)n * m * 8 / best / 1e6);
    }
    return 0;
}
