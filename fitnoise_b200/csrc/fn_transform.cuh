// Moderated log transform of count data (fitnoise.transform, reference fittransform.py:65-219).
//
// y_gj = ln(P_j(x_gj)) / (order ln 2),  P_j(x) = x^order + a_0j x^(order-1) + ... + a_(order-1)j
// (Transform_polynomial._apply_transform, fittransform.py:190-194).  The objective
// ln(var_h1) - ln(var_h0) (fittransform.py:103-121) and its gradient need, over all genes,
//   s_j  = sum_g y_gj                      Syy = sum_g |y_g|^2        Stt = sum_g |Q' y_g|^2
//   A_ij = sum_g y_gj w_gj(i)   B_ij = sum_g w_gj(i)   D_ij = sum_g (Q Q' y_g)_j w_gj(i),
//   w_gj(i) = d y_gj / d a_ij = x_gj^(order-1-i) / (P_j(x_gj) order ln 2)
// with Q the orthonormal basis of the design; the host combines them (fn_api.cu).  One pass over x.
//
// Mapping: a warp per gene; lane l owns samples l, l+32, ... (K per lane), read straight from
// global memory (each read of a warp is one contiguous 256-byte run), kept in registers for the
// second sweep; the per-sample sums live in the owning lane's registers for the whole kernel.
// The design basis sits in shared memory (m x 8, zero padded) so that the register budget allows
// two CTAs per SM: the kernel is bound by the fp64 pipe (one log and one division per count).
#pragma once
#include <cuda_runtime.h>
#include "fn_kernels.cuh"

namespace fn {

#define TR_MAX_ORDER 3
#define TR_MAXC 8             // design columns of the transform objective

struct TransformParams {
    long long n;
    int m, c, order;
    const double* x;        // n x m counts
    const double* q;        // m x c orthonormal design basis (device)
    const double* a;        // order x m polynomial coefficients (device)
    GridReduce gr;          // nacc = 2 + m (1 + 3 order): [Syy, Stt, s[m], A[order][m], B[order][m], D[order][m]]
    StepVars sv;
};

// K = samples per lane = ceil(m / 32); C = design columns (template width)
template <int K, int ORDER>
__global__ void __launch_bounds__(256, (K * ORDER > 9 ? 1 : 2)) transform_kernel(const TransformParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* vals = (double*)smem_raw;                       // nacc
    double* scratch = vals + (2 + p.m * (1 + 3 * ORDER));   // warps x per-lane values, reused
    double* qs = scratch + 32 * (blockDim.x >> 5);          // (32 K) x TR_MAXC design basis, zero padded
    const int m = p.m, c = p.c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const double inv_l = 1.0 / (ORDER * 0.6931471805599453094);

    for (int e = threadIdx.x; e < 32 * K * TR_MAXC; e += blockDim.x) {
        const int j = e / TR_MAXC, t = e % TR_MAXC;
        qs[e] = (j < m && t < c) ? p.q[j * c + t] : 0.0;
    }
    __syncthreads();
    double a[K][ORDER];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = lane + 32 * k;
#pragma unroll
        for (int i = 0; i < ORDER; ++i) a[k][i] = j < m ? p.a[i * m + j] : 1.0;
    }
    double s[K], A[K][ORDER], B[K][ORDER], D[K][ORDER];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        s[k] = 0.0;
#pragma unroll
        for (int i = 0; i < ORDER; ++i) { A[k][i] = 0.0; B[k][i] = 0.0; D[k][i] = 0.0; }
    }
    double syy = 0.0, stt = 0.0;

    const long long gwarp = (long long)blockIdx.x * nwarp + warp, stride = (long long)gridDim.x * nwarp;
    for (long long g = gwarp; g < p.n; g += stride) {
        double x[K], y[K], pinv[K];
        double t[TR_MAXC];
#pragma unroll
        for (int q = 0; q < TR_MAXC; ++q) t[q] = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j = lane + 32 * k;
            x[k] = j < m ? __ldg(&p.x[g * m + j]) : 0.0;
            double poly = 1.0;
#pragma unroll
            for (int i = 0; i < ORDER; ++i) poly = poly * x[k] + a[k][i];
            pinv[k] = 1.0 / poly;
            y[k] = j < m ? log(poly) * inv_l : 0.0;
            syy += y[k] * y[k];
            const double* qr = qs + (size_t)j * TR_MAXC;
#pragma unroll
            for (int q = 0; q < TR_MAXC; ++q) if (q < c) t[q] += qr[q] * y[k];
        }
#pragma unroll
        for (int q = 0; q < TR_MAXC; ++q) {
            if (q < c) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) t[q] += __shfl_xor_sync(0xffffffffu, t[q], off);
                if (lane == 0) stt += t[q] * t[q];
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double proj = 0.0;                 // (Q Q' y_g)_j
            const double* qr = qs + (size_t)(lane + 32 * k) * TR_MAXC;
#pragma unroll
            for (int q = 0; q < TR_MAXC; ++q) if (q < c) proj += qr[q] * t[q];
            s[k] += y[k];
            double w = pinv[k] * inv_l;        // i = ORDER-1: x^0
#pragma unroll
            for (int i = ORDER - 1; i >= 0; --i) {
                const int j = lane + 32 * k;
                const double wj = j < m ? w : 0.0;
                A[k][i] += y[k] * wj; B[k][i] += wj; D[k][i] += proj * wj;
                w *= x[k];
            }
        }
    }

    // CTA reduction: the same lane of every warp owns the same samples
    const int nacc = 2 + m * (1 + 3 * ORDER);
    auto warp_scalar = [&](double v) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        return v;
    };
    syy = warp_scalar(syy);
    stt = warp_scalar(stt);
    if (lane == 0) { scratch[warp] = syy; scratch[nwarp + warp] = stt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int w = 0; w < nwarp; ++w) { s0 += scratch[w]; s1 += scratch[nwarp + w]; }
        vals[0] = s0; vals[1] = s1;
    }
    __syncthreads();
    auto column = [&](double v, int k, int slot) {          // slot: 0 = s, 1+i = A_i, 1+ORDER+i = B_i, ...
        scratch[warp * 32 + lane] = v;
        __syncthreads();
        const int j = lane + 32 * k;
        if (warp == 0 && j < m) {
            double acc = 0.0;
            for (int w = 0; w < nwarp; ++w) acc += scratch[w * 32 + lane];
            vals[2 + slot * m + j] = acc;
        }
        __syncthreads();
    };
#pragma unroll
    for (int k = 0; k < K; ++k) {
        column(s[k], k, 0);
#pragma unroll
        for (int i = 0; i < ORDER; ++i) {
            column(A[k][i], k, 1 + i);
            column(B[k][i], k, 1 + ORDER + i);
            column(D[k][i], k, 1 + 2 * ORDER + i);
        }
    }
    (void)nacc;
    grid_reduce(p.gr, p.sv, vals);
}

// y = ln(P(x)) / (order ln 2) - column mean   (fittransform.py:161-163)
__global__ void transform_apply_kernel(long long n, int m, int order, const double* x, const double* a,
                                       const double* col_mean, double* y) {
    const long long total = n * m, stride = (long long)gridDim.x * blockDim.x;
    const double inv_l = 1.0 / (order * 0.6931471805599453094);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int j = (int)(e % m);
        double poly = 1.0;
        for (int i = 0; i < order; ++i) poly = poly * x[e] + a[i * m + j];
        y[e] = log(poly) * inv_l - col_mean[j];
    }
}

}  // namespace fn
