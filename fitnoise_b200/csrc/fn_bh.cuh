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

// ------------------------------------------------------------------------------------------
// One-sweep form of the sort (up to BH_ONESWEEP_MAX p-values): ONE kernel per digit instead of five.
// Measured (tools/bench_bh.py, CUDA-graph replay): 84 us against 150 us at 20,000 p-values, 108 against 153
// at 200,000; from about 500,000 on the tiles' wait for their predecessors costs more than the four small
// launches it saves (269 against 251 us at 1M; publishing the counts before the ranking and looking back
// 8 or 16 tiles per round trip did not change that), so larger problems keep the five-kernel passes.
//
//   bh_hist_all_kernel   reads the keys once and counts all eight digits (8 x 256 global counters);
//   bh_bases_kernel      turns each digit's counts into exclusive offsets (one CTA);
//   bh_onesweep_kernel   per digit: a CTA takes the next tile of 2048 keys by ticket, ranks its keys by
//                        digit (stable: striped items in order, warp match + per-warp counts, then ONE scan
//                        per digit over the 8 x 8 (item, warp) counts), publishes its 256 digit counts and
//                        finds the counts of all earlier tiles by decoupled look-back (each (tile, digit)
//                        state word carries {flag, count}: "aggregate" = this tile alone, "inclusive" =
//                        this tile and everything before it; a reader walks back until it meets an
//                        inclusive word), then scatters.  A tile only ever waits for tiles with smaller
//                        tickets, which are already running: no deadlock whatever the grid.
// The state words are zeroed by one memset per adjustment (they cannot carry an epoch instead: the
// launch sequence is replayed from a CUDA graph with identical arguments).
// ------------------------------------------------------------------------------------------
#define BH_ONESWEEP_MAX 400000
#define BH_FLAG_AGG (1ull << 32)
#define BH_FLAG_INC (2ull << 32)

__global__ void __launch_bounds__(BH_THREADS) bh_hist_all_kernel(long long n, const uint64_t* keys, uint32_t* hist /* 8 x 256 */) {
    __shared__ uint32_t h[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += BH_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * BH_CHUNK;
    for (int it = 0; it < BH_ITEMS; ++it) {
        const long long i = base + it * BH_THREADS + threadIdx.x;
        const bool live = i < n;
        const uint64_t k = live ? keys[i] : 0ull;
        // low digits (mantissa bits of a p-value: spread out): plain shared-memory atomics
        if (live) {
#pragma unroll
            for (int d = 0; d < 5; ++d) atomicAdd(&h[d][(k >> (8 * d)) & 0xff], 1u);
        }
        // high digits (sign, exponent, leading mantissa bits: a handful of values): one atomic per distinct
        // digit of the warp, or the same-address atomics of 32 lanes serialise
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int d = 5; d < 8; ++d) {
            const uint32_t dig = live ? (uint32_t)((k >> (8 * d)) & 0xff) : 256u;
            const uint32_t peers = __match_any_sync(0xffffffffu, dig);
            if (live && (peers & ((1u << lane) - 1u)) == 0u) atomicAdd(&h[d][dig], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 256; i += BH_THREADS) {
        const uint32_t c = (&h[0][0])[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// hist[d][0..255] -> exclusive offsets in place, for the eight digits (grid 8, 256 threads)
__global__ void __launch_bounds__(256) bh_bases_kernel(uint32_t* hist) {
    __shared__ uint32_t warp_tot[33];
    uint32_t* row = hist + 256 * blockIdx.x;
    uint32_t total;
    const uint32_t ex = bh_block_exclusive_scan(row[threadIdx.x], warp_tot, total);
    row[threadIdx.x] = ex;
}

__device__ __forceinline__ unsigned long long bh_ld_state(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void bh_st_state(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(BH_THREADS) bh_onesweep_kernel(long long n, const uint64_t* keys_in, const uint32_t* idx_in,
                                                                 uint64_t* keys_out, uint32_t* idx_out, int shift,
                                                                 const uint32_t* digit_base /* 256 */,
                                                                 unsigned long long* state /* tiles x 256 */,
                                                                 unsigned int* ticket) {
    __shared__ uint16_t warp_cnt[BH_ITEMS][BH_THREADS / 32][256];      // 32 KB: a warp holds at most 32 keys of a digit
    __shared__ uint32_t base_sh[256];
    __shared__ unsigned int tile_sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) tile_sh = atomicAdd(ticket, 1u);
    {
        uint4* z = reinterpret_cast<uint4*>(&warp_cnt[0][0][0]);
        for (int i = threadIdx.x; i < (int)(sizeof(warp_cnt) / 16); i += BH_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const unsigned int tile = tile_sh;
    const long long base = (long long)tile * BH_CHUNK;
    uint64_t key[BH_ITEMS];
    uint32_t id[BH_ITEMS];
    uint32_t dr[BH_ITEMS];            // digit | rank within the warp's peers << 16 | live << 31
#pragma unroll
    for (int it = 0; it < BH_ITEMS; ++it) {
        const long long i = base + it * BH_THREADS + threadIdx.x;
        const bool live = i < n;
        key[it] = 0; id[it] = 0;
        uint32_t d = 256u;            // dead lanes get a digit of their own
        if (live) { key[it] = keys_in[i]; id[it] = idx_in[i]; d = (uint32_t)((key[it] >> shift) & 0xff); }
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (live && rank == 0) warp_cnt[it][warp][d] = (uint16_t)__popc(peers);
        dr[it] = d | (rank << 16) | (live ? 0x80000000u : 0u);
    }
    __syncthreads();
    // thread t owns digit t: exclusive offsets of the (item, warp) groups in order, the tile's count
    {
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int it = 0; it < BH_ITEMS; ++it)
#pragma unroll
            for (int w = 0; w < BH_THREADS / 32; ++w) { const uint32_t c = warp_cnt[it][w][d]; warp_cnt[it][w][d] = (uint16_t)run; run += c; }
        unsigned long long* mine = state + (size_t)tile * 256 + d;
        uint32_t before = 0;
        if (tile == 0) {
            bh_st_state(mine, BH_FLAG_INC | run);
        } else {
            bh_st_state(mine, BH_FLAG_AGG | run);
            for (long long t = (long long)tile - 1; t >= 0; --t) {
                unsigned long long sv;
                do { sv = bh_ld_state(state + (size_t)t * 256 + d); } while ((sv >> 32) == 0ull);
                before += (uint32_t)sv;
                if (sv & BH_FLAG_INC) break;
            }
            bh_st_state(mine, BH_FLAG_INC | (unsigned long long)(before + run));
        }
        base_sh[d] = digit_base[d] + before;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < BH_ITEMS; ++it) {
        if (dr[it] & 0x80000000u) {
            const uint32_t d = dr[it] & 0xffffu, rank = (dr[it] >> 16) & 0x7fffu;
            const uint32_t pos = base_sh[d] + warp_cnt[it][warp][d] + rank;
            keys_out[pos] = key[it];
            idx_out[pos] = id[it];
        }
    }
}

// carry[b] = min over chunks > b (inf for the last): one CTA, each thread a contiguous run of chunks
__global__ void __launch_bounds__(1024) bh_carry_scan_kernel(int nchunks, const double* chunk_min, double* carry) {
    __shared__ double seg[1024];
    const int per = (nchunks + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = (lo + per < nchunks) ? lo + per : nchunks;
    double mn = INFINITY;
    for (int i = hi - 1; i >= lo; --i) mn = fmin(mn, chunk_min[i]);
    seg[threadIdx.x] = mn;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {                    // suffix minimum over the threads' runs
        const double v = (threadIdx.x + off < 1024) ? seg[threadIdx.x + off] : INFINITY;
        __syncthreads();
        seg[threadIdx.x] = fmin(seg[threadIdx.x], v);
        __syncthreads();
    }
    double run = (threadIdx.x + 1 < 1024) ? seg[threadIdx.x + 1] : INFINITY;
    for (int i = hi - 1; i >= lo; --i) { carry[i] = run; run = fmin(run, chunk_min[i]); }
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
