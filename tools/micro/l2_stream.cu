// Micro-benchmark: does a shard that fits the 126 MB L2 stream faster than HBM when it is swept again and
// again (the optimiser loop of the strong-scaling case: 96 MB per GPU, one sweep per evaluation)?
// One persistent kernel makes PASSES sweeps over W MB with (a) the TMA ring of fn_tile.cuh, every CTA
// re-reading its own tiles, (b) plain coalesced 16-byte loads; prints GB/s per sweep.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_stream l2_stream.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "../../fitnoise_b200/csrc/fn_tile.cuh"
using namespace fn;

__global__ void __launch_bounds__(256) ring_passes(TileSrc src, int stages, long long n_tiles, int passes, double* out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[FN_MAX_STAGES];
    const TileLayout L = tile_layout(src);
    TileRing ring;
    ring_setup(ring, bars, stages, n_tiles, true);
    __syncthreads();
    ring.begin(smem_raw, bars, src, L);
    double acc = 0.0;
    const int per = src.genes * src.m;
    for (int p = 0; p < passes; ++p)
        ring.pass(smem_raw, bars, src, L, false, [&](unsigned char* st, long long tile) {
            const double* y = (const double*)st;
            for (int i = threadIdx.x; i < per; i += blockDim.x) acc += y[i];
        });
    ring.drain(bars);
    if (acc == 12345.678) out[0] = acc;
}

__global__ void __launch_bounds__(256) ldg_passes(const double2* y, long long n2, int passes, double* out) {
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
            double2 v;
            asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(y + i));
            acc += v.x + v.y;
        }
    if (acc == 12345.678) out[0] = acc;
}

int main() {
    const int m = 96;
    const long long n_max = 1000448;
    double *y, *out; uint8_t* status;
    cudaMalloc(&y, n_max * m * 8); cudaMalloc(&out, 8); cudaMalloc(&status, n_max);
    cudaMemset(y, 0, n_max * m * 8); cudaMemset(status, 1, n_max);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const long long sizes[] = {31250, 62500, 93750, 125000, 156250, 250000, 1000000};
    for (int keep = 0; keep <= 1; ++keep)
        for (long long n : sizes) {
            const int genes = 64, stages = 2, ctas = 2, passes = 20;
            TileSrc src; src.y = y; src.aux = nullptr; src.gs = nullptr; src.cst = nullptr; src.status = status;
            src.m = m; src.genes = genes; src.row_stride = m; src.l2_keep = keep;
            const size_t smem = tile_stage_bytes(genes, m, false, false, false) * stages;
            cudaFuncSetAttribute(ring_passes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            const long long n_tiles = (n + genes - 1) / genes;
            float t[2];
            for (int k = 0; k < 2; ++k) {            // passes and 2 x passes: the difference is free of launch cost and first touch
                float best = 1e9;
                for (int rep = 0; rep < 4; ++rep) {
                    cudaEventRecord(e0);
                    ring_passes<<<148 * ctas, 256, smem>>>(src, stages, n_tiles, passes * (k + 1), out);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
                }
                t[k] = best;
            }
            const double per_pass_ms = (t[1] - t[0]) / passes;
            printf("tma ring  evict_last %d  %7lld genes %6.1f MB : %.4f ms/sweep  %.0f GB/s  (%s)\n", keep, n, n * m * 8 / 1048576.0,
                   per_pass_ms, (double)n * (m * 8 + 1) / per_pass_ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    for (long long n : sizes) {
        const int passes = 20;
        float t[2];
        for (int k = 0; k < 2; ++k) {
            float best = 1e9;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                ldg_passes<<<148 * 8, 256>>>((const double2*)y, n * m / 2, passes * (k + 1), out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            t[k] = best;
        }
        const double per_pass_ms = (t[1] - t[0]) / passes;
        printf("ld.cg 16B               %7lld genes %6.1f MB : %.4f ms/sweep  %.0f GB/s\n", n, n * m * 8 / 1048576.0, per_pass_ms,
               (double)n * m * 8 / per_pass_ms / 1e6);
    }
    return 0;
}
