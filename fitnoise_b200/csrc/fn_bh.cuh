// K4: Benjamini-Hochberg q-values on the device (replaces p_adjust.py:3-21, two numpy argsorts).
//
//   key_i  = bit pattern of p_i (NaN -> +inf, read back as 1.0); non-negative doubles order like their bit patterns
//   sort   = stable LSD radix sort of (key, index), 8 passes of 8 bits: per-CTA digit histograms,
//            one scan, stable scatter (warp match + per-warp digit offsets)
//   v_r    = (p_(r) * n) / (r + 1)  with the reference's operation order (p_adjust.py:14-16),
//            un-fused multiply and divide, so q is BIT-EXACT given p
//   q_(r)  = min(1, min_{s >= r} v_s): suffix-min scan (min is exact, any association order)
//   q[index_(r)] = q_(r)
// Ties need no care: equal p get equal q whatever their order (cumulative min), as in the reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fn {

#define BH_THREADS 256
#define BH_ITEMS 8                          // keys per thread per CTA chunk
#define BH_CHUNK (BH_THREADS * BH_ITEMS)

__global__ void bh_keys_kernel(long long n, const double* p, uint64_t* keys, uint32_t* idx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // NaN counts as p = 1.0 (p_adjust.py:7).  Its key is +inf so that the sorted order doubles as
    // the top-table order (ascending p, stable, missing p last, like R's order()); bh_key_p() maps
    // the key back to 1.0.  Entries with p = 1 get q = 1 wherever they rank among themselves.
    const double v = p[i];
    keys[i] = (v != v) ? 0x7ff0000000000000ull : (uint64_t)__double_as_longlong(v);
    idx[i] = (uint32_t)i;
}

__device__ __forceinline__ double bh_key_p(uint64_t key) {
    return key == 0x7ff0000000000000ull ? 1.0 : __longlong_as_double((long long)key);
}

// hist[d * nblocks + b] = number of keys of CTA b's chunk whose digit is d
__global__ void __launch_bounds__(BH_THREADS) bh_hist_kernel(long long n, const uint64_t* keys, int shift,
                                                             uint32_t* hist, int nblocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * BH_CHUNK;
    for (int it = 0; it < BH_ITEMS; ++it) {
        const long long i = base + it * BH_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xff], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// Exclusive scan of hist (256 * nblocks entries, digit-major) in place, in three small launches:
// per-segment sums, a one-CTA scan of the segment sums, per-segment scan with the segment's offset.
#define BH_SEG (BH_THREADS * BH_ITEMS)

__device__ __forceinline__ uint32_t bh_block_exclusive_scan(uint32_t v, uint32_t* warp_tot /* shared, >= 32 */, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nwarp ? warp_tot[lane] : 0;
        uint32_t winc = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, off);
            if (lane >= off) winc += t;
        }
        warp_tot[lane] = winc - w;                 // exclusive offset of each warp
        if (lane == 31) warp_tot[32] = winc;       // block total
    }
    __syncthreads();
    total = warp_tot[32];
    return warp_tot[warp] + inc - v;
}

__global__ void __launch_bounds__(BH_THREADS) bh_segsum_kernel(const uint32_t* hist, long long total, uint32_t* seg_sum) {
    __shared__ uint32_t red[BH_THREADS];
    const long long base = (long long)blockIdx.x * BH_SEG;
    uint32_t s = 0;
    for (int k = 0; k < BH_ITEMS; ++k) {
        const long long i = base + k * BH_THREADS + threadIdx.x;
        if (i < total) s += hist[i];
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = BH_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) seg_sum[blockIdx.x] = red[0];
}

// one CTA of 1024 threads: exclusive scan of up to 1024 * k segment sums
__global__ void __launch_bounds__(1024) bh_segscan_kernel(uint32_t* seg_sum, int nseg) {
    __shared__ uint32_t warp_tot[33];
    uint32_t carry = 0;
    for (int base = 0; base < nseg; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < nseg ? seg_sum[i] : 0;
        uint32_t total;
        const uint32_t ex = bh_block_exclusive_scan(v, warp_tot, total);
        if (i < nseg) seg_sum[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(BH_THREADS) bh_segapply_kernel(uint32_t* hist, long long total, const uint32_t* seg_off) {
    __shared__ uint32_t warp_tot[33];
    const long long base = (long long)blockIdx.x * BH_SEG + (long long)threadIdx.x * BH_ITEMS;
    uint32_t v[BH_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < BH_ITEMS; ++k) {
        const long long i = base + k;
        v[k] = i < total ? hist[i] : 0;
        s += v[k];
    }
    uint32_t block_total;
    uint32_t run = seg_off[blockIdx.x] + bh_block_exclusive_scan(s, warp_tot, block_total);
#pragma unroll
    for (int k = 0; k < BH_ITEMS; ++k) {
        const long long i = base + k;
        if (i < total) hist[i] = run;
        run += v[k];
    }
}

// stable scatter of CTA b's chunk using the scanned histogram as per-(digit, CTA) base offsets
__global__ void __launch_bounds__(BH_THREADS) bh_scatter_kernel(long long n, const uint64_t* keys_in,
                                                                const uint32_t* idx_in, uint64_t* keys_out,
                                                                uint32_t* idx_out, int shift,
                                                                const uint32_t* hist, int nblocks) {
    __shared__ uint32_t digit_base[256];                    // next free output slot per digit
    __shared__ uint32_t warp_cnt[BH_THREADS / 32][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    digit_base[threadIdx.x] = hist[(size_t)threadIdx.x * nblocks + blockIdx.x];
    const long long base = (long long)blockIdx.x * BH_CHUNK;
    for (int it = 0; it < BH_ITEMS; ++it) {
        for (int w = 0; w < BH_THREADS / 32; ++w) warp_cnt[w][threadIdx.x] = 0;
        __syncthreads();
        const long long i = base + it * BH_THREADS + threadIdx.x;
        const bool live = i < n;
        uint64_t key = 0; uint32_t id = 0; uint32_t d = 0;
        if (live) { key = keys_in[i]; id = idx_in[i]; d = (uint32_t)((key >> shift) & 0xff); }
        // lanes with the same digit (dead lanes get a digit of their own: 256)
        const uint32_t peers = __match_any_sync(0xffffffffu, live ? d : 256u);
        const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (live && rank_in_warp == 0) warp_cnt[warp][d] = __popc(peers);
        __syncthreads();
        // per digit: exclusive scan over warps (thread t owns digit t), advance the digit base
        uint32_t mine_before = 0;
        {
            uint32_t run = 0;
            const uint32_t t = threadIdx.x;
            for (int w = 0; w < BH_THREADS / 32; ++w) { const uint32_t c = warp_cnt[w][t]; warp_cnt[w][t] = run; run += c; }
            mine_before = run;              // total of this round for digit t
        }
        __syncthreads();
        if (live) {
            const uint32_t pos = digit_base[d] + warp_cnt[warp][d] + rank_in_warp;
            keys_out[pos] = key;
            idx_out[pos] = id;
        }
        __syncthreads();
        digit_base[threadIdx.x] += mine_before;
    }
}

// v_r = (p_r * n) / (r + 1); per-chunk minimum
__global__ void __launch_bounds__(BH_THREADS) bh_chunkmin_kernel(long long n, const uint64_t* keys, double* chunk_min) {
    __shared__ double red[BH_THREADS];
    const long long base = (long long)blockIdx.x * BH_CHUNK + (long long)threadIdx.x * BH_ITEMS;
    double mn = INFINITY;
    for (int k = 0; k < BH_ITEMS; ++k) {
        const long long r = base + k;
        if (r < n) {
            const double p = bh_key_p(keys[r]);
            const double v = __ddiv_rn(__dmul_rn(p, (double)n), (double)(r + 1));
            mn = fmin(mn, v);
        }
    }
    red[threadIdx.x] = mn;
    __syncthreads();
    for (int s = BH_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = fmin(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_min[blockIdx.x] = red[0];
}

// carry[b] = min over chunks > b (inf for the last); thread b scans the later chunk minima
__global__ void bh_carry_kernel(int nchunks, const double* chunk_min, double* carry) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nchunks) return;
    double run = INFINITY;
    for (int i = b + 1; i < nchunks; ++i) run = fmin(run, chunk_min[i]);
    carry[b] = run;
}

// q_(r) = min(1, suffix-min), scattered back to gene order
__global__ void __launch_bounds__(BH_THREADS) bh_finish_kernel(long long n, const uint64_t* keys, const uint32_t* idx,
                                                               const double* carry, double* q) {
    __shared__ double tmin[BH_THREADS];
    const long long base = (long long)blockIdx.x * BH_CHUNK + (long long)threadIdx.x * BH_ITEMS;
    double v[BH_ITEMS];
    double mn = INFINITY;
    for (int k = BH_ITEMS - 1; k >= 0; --k) {
        const long long r = base + k;
        double x = INFINITY;
        if (r < n) {
            const double p = bh_key_p(keys[r]);
            x = __ddiv_rn(__dmul_rn(p, (double)n), (double)(r + 1));
        }
        mn = fmin(mn, x);
        v[k] = mn;                            // suffix min within the thread's run
    }
    tmin[threadIdx.x] = mn;
    __syncthreads();
    // min over later threads of the CTA and later chunks
    double later = carry[blockIdx.x];
    for (int t = threadIdx.x + 1; t < BH_THREADS; ++t) later = fmin(later, tmin[t]);
    for (int k = 0; k < BH_ITEMS; ++k) {
        const long long r = base + k;
        if (r < n) q[idx[r]] = fmin(1.0, fmin(v[k], later));
    }
}

}  // namespace fn
