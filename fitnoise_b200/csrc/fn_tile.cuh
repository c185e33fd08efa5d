// Gene-tile staging for the per-gene kernels (sm_100a).
//
// The genes x samples matrices are row-major and a tile of G consecutive genes is ONE contiguous
// block of G*m*8 bytes, so a whole tile moves HBM -> shared memory with a single 1-D bulk
// asynchronous copy (cp.async.bulk, the TMA engine; SASS: UBLKCP) that completes on an mbarrier.
// One elected thread per CTA issues the copies for up to `stages` tiles ahead; all threads then
// read their gene's row out of shared memory (thread-per-gene or lpg lanes per gene), which turns
// the row-per-thread access pattern into fully coalesced DRAM traffic without spending LSU issue
// slots or registers on the transfer.
//
// CTAs are persistent: CTA b handles tiles b, b+gridDim.x, ...
#pragma once
#include <stdint.h>

namespace fn {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// 1-D bulk copy global -> shared; dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// The same with an L2 eviction-priority hint (evict_last): a shard that fits the 126 MB L2 is re-read by
// every evaluation of the optimiser loop, and a cyclic sweep over a working set near the cache size
// evicts each line just before its reuse unless the lines are marked to stay.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
// order this thread's generic-proxy shared-memory writes before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

#define FN_MAX_STAGES 4

// What one tile stage holds.  All per-gene arrays are padded on the device to a multiple of
// FN_ROW_PAD genes, so a full-size copy of the last tile never runs off an allocation.
#define FN_ROW_PAD 1024
struct TileSrc {
    const double* y;         // n_pad x m
    const double* aux;       // n_pad x m or nullptr
    const double* gs;        // n_pad or nullptr
    const double* cst;       // n_pad or nullptr
    const uint8_t* status;   // n_pad (bit0 eligible, bit1 control) -- always present
    int m;
    int genes;               // G, a multiple of 16
    int l2_keep;             // 1: the y / aux tiles are fetched with the L2 evict_last hint (shard fits the L2)
    int row_stride;          // doubles between consecutive gene rows of y / aux INSIDE the stage: m (dense: the
                             // tile is one bulk copy) or m + 2 (padded rows, one bulk copy per gene row: with a
                             // stride of 2 mod 4 doubles the rows of 8 consecutive genes start in 8 distinct
                             // 16-byte bank groups, so thread-per-gene LDS.128 reads of a sample pair are
                             // conflict-free while every thread of the warp visits the SAME samples)
};

// Byte offsets of the arrays inside one stage buffer (uniform over the CTA, computed once).
struct TileLayout {
    uint32_t aux_off, gs_off, cst_off, status_off;   // y is at offset 0
    uint32_t stage_bytes, tx_bytes, row_bytes;
};

__host__ __device__ inline size_t tile_stage_bytes(int genes, int m, bool aux, bool gs, bool cst, int row_stride = 0) {
    size_t b = (size_t)genes * (row_stride > 0 ? row_stride : m) * 8 * (aux ? 2 : 1) + (size_t)genes * 8 * ((gs ? 1 : 0) + (cst ? 1 : 0)) + genes;
    return (b + 127) & ~(size_t)127;
}

__device__ __forceinline__ TileLayout tile_layout(const TileSrc& s) {
    TileLayout L;
    L.row_bytes = (uint32_t)s.genes * s.m * 8;                        // bytes one matrix contributes to a stage's transfer
    const uint32_t span = (uint32_t)s.genes * s.row_stride * 8;       // bytes it occupies in the stage
    uint32_t at = span;
    L.aux_off = at; if (s.aux) at += span;
    L.gs_off = at; if (s.gs) at += (uint32_t)s.genes * 8;
    L.cst_off = at; if (s.cst) at += (uint32_t)s.genes * 8;
    L.status_off = at; at += (uint32_t)s.genes;
    // bytes the bulk copies of a stage deliver: everything when the rows are dense; only the per-gene
    // scalars when they are padded (the rows then travel as cp.async, counted by arrivals, not bytes)
    L.tx_bytes = (s.row_stride == s.m) ? at : at - (s.aux ? 2u : 1u) * span;
    L.stage_bytes = (at + 127u) & ~127u;
    return L;
}

// Dense rows: called by ONE thread (threadIdx.x == 0): arm the stage's barrier and start the bulk copies
// of `tile` into it.
// Padded rows (row_stride != m): called by EVERY thread of the CTA.  A bulk copy per gene row would be
// the obvious form, but UBLKCP is a uniform-datapath instruction (one copy per warp instruction, lanes
// serialised) and the TMA unit sustains only about one such small operation per 100 ns per SM -- measured
// 0.57 TB/s over the GPU at 384-byte rows, slower than the kernel it was meant to feed.  So the rows move
// with 16-byte cp.async (LDGSTS) issued by all threads, coalesced in global memory, and each thread's
// copies arrive on the stage's mbarrier (cp.async.mbarrier.arrive.noinc: the barrier is initialised for
// blockDim.x + 1 arrivals); thread 0 adds the per-gene scalars as bulk copies with their byte count.
__device__ __forceinline__ bool tile_padded(const TileSrc& src) { return src.row_stride != src.m; }
__device__ __forceinline__ bool tile_issuer(const TileSrc& src) { return tile_padded(src) || threadIdx.x == 0; }
__device__ __forceinline__ uint32_t tile_barrier_count(const TileSrc& src) { return tile_padded(src) ? blockDim.x + 1u : 1u; }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_s(uint32_t dst_shared, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_shared), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tile_issue(const TileSrc& src, const TileLayout& L, unsigned char* stage,
                                           uint64_t* bar, long long tile) {
    const long long g0 = tile * src.genes;
    if (tile_padded(src)) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, L.tx_bytes);
            if (src.gs) bulk_g2s(stage + L.gs_off, src.gs + g0, (uint32_t)src.genes * 8, bar);
            if (src.cst) bulk_g2s(stage + L.cst_off, src.cst + g0, (uint32_t)src.genes * 8, bar);
            bulk_g2s(stage + L.status_off, src.status + g0, (uint32_t)src.genes, bar);
        }
        // chunk i (16 bytes) of the tile: row i / cpr, position i % cpr; thread t takes t, t + T, ...  The
        // shared-memory offset is carried along instead of being recomputed (one division per tile).
        const int cpr = src.m >> 1;                              // 16-byte chunks per row (m is even)
        const int total = src.genes * cpr, T = (int)blockDim.x;
        const uint32_t pitch = (uint32_t)src.row_stride * 8, gap = pitch - (uint32_t)cpr * 16u;
        const int row0 = (int)threadIdx.x / cpr, drow = T / cpr, dk = T - drow * cpr;
        int k = (int)threadIdx.x - row0 * cpr;
        const uint32_t step = (uint32_t)drow * pitch + (uint32_t)dk * 16u;
        uint32_t at = smem_u32(stage) + (uint32_t)row0 * pitch + (uint32_t)k * 16u;
        const char* gy = (const char*)(src.y + g0 * src.m) + (size_t)threadIdx.x * 16;
        const size_t gstep = (size_t)T * 16;
        if (src.aux) {
            const char* ga = (const char*)(src.aux + g0 * src.m) + (size_t)threadIdx.x * 16;
            for (int i = threadIdx.x; i < total; i += T, gy += gstep, ga += gstep) {
                cp_async16_s(at, gy);
                cp_async16_s(at + L.aux_off, ga);
                at += step; k += dk;
                if (k >= cpr) { k -= cpr; at += gap; }
            }
        } else {
#pragma unroll 4
            for (int i = threadIdx.x; i < total; i += T, gy += gstep) {
                cp_async16_s(at, gy);
                at += step; k += dk;
                if (k >= cpr) { k -= cpr; at += gap; }
            }
        }
        cp_async_arrive(bar);
        return;
    }
    mbar_expect_tx(bar, L.tx_bytes);
    if (src.l2_keep) {
        const uint64_t pol = l2_policy_evict_last();
        bulk_g2s_hint(stage, src.y + g0 * src.m, L.row_bytes, bar, pol);
        if (src.aux) bulk_g2s_hint(stage + L.aux_off, src.aux + g0 * src.m, L.row_bytes, bar, pol);
    } else {
        bulk_g2s(stage, src.y + g0 * src.m, L.row_bytes, bar);
        if (src.aux) bulk_g2s(stage + L.aux_off, src.aux + g0 * src.m, L.row_bytes, bar);
    }
    if (src.gs) bulk_g2s(stage + L.gs_off, src.gs + g0, (uint32_t)src.genes * 8, bar);
    if (src.cst) bulk_g2s(stage + L.cst_off, src.cst + g0, (uint32_t)src.genes * 8, bar);
    bulk_g2s(stage + L.status_off, src.status + g0, (uint32_t)src.genes, bar);
}

// Initialise the barriers and start the first stages-1 copies (call once, all threads; the
// caller's next __syncthreads() publishes the barrier initialisation).
__device__ __forceinline__ void tile_prologue(unsigned char* smem, uint64_t* bars, int stages, const TileSrc& src,
                                              const TileLayout& L, long long n_tiles) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) mbar_init(&bars[i], 1);      // dense rows only (K1, K3)
        mbar_fence_init();
    }
    if (tile_issuer(src)) {
        for (int s = 0; s < stages - 1; ++s) {
            const long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
            if (t < n_tiles) tile_issue(src, L, smem + (size_t)s * L.stage_bytes, &bars[s], t);
        }
    }
}

// Persistent tile loop.  `body(stage, tile)` is run by every thread of the CTA for each tile with
// `stage` pointing at the tile's buffer in shared memory; it may write that buffer (in-place
// gradient terms), in which case `wrote_smem` makes the ring fence those generic-proxy writes
// before the TMA engine refills the slot.  `smem` must be the CTA's dynamic shared memory base so
// that the compiler keeps the shared address space (LDS, not generic LD) for the body's loads.
template <class Body>
__device__ __forceinline__ void tile_loop(unsigned char* smem, uint64_t* bars, int stages, const TileSrc& src,
                                          const TileLayout& L, long long n_tiles, bool wrote_smem, Body body) {
    const long long step = gridDim.x;
    int slot = 0, pre = stages - 1;       // pre: slot consumed in the previous iteration
    uint32_t parity = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += step) {
        if (tile_issuer(src)) {
            const long long ahead = tile + (long long)(stages - 1) * step;
            if (ahead < n_tiles) tile_issue(src, L, smem + (size_t)pre * L.stage_bytes, &bars[pre], ahead);
        }
        mbar_wait(&bars[slot], parity);
        body(smem + (size_t)slot * L.stage_bytes, tile);
        if (wrote_smem) fence_proxy_async();
        __syncthreads();
        pre = slot;
        if (++slot == stages) { slot = 0; parity ^= 1u; }
    }
}


// ------------------------------------------------------------------------------------------
// Tile ring with state that outlives one pass over the CTA's tiles (the resident evaluation kernel).
//
// The CTA's tiles are b, b+G, ... (`cnt` of them).  `pass()` consumes them once.  With `wrap` the
// producer keeps running ahead ACROSS passes: while the tail of pass k is consumed, the first
// stages-1 tiles of pass k+1 are already being copied -- they are theta-independent input -- so
// a resident kernel that waits for its next command finds its first tiles in shared memory when the
// command arrives.  Without `wrap` a pass issues exactly its own tiles (the one-launch-per-
// evaluation form; `begin()` must then be called before every pass).
// ------------------------------------------------------------------------------------------
struct TileRing {
    int slot, pre;           // slot to consume next; slot consumed last (the one the producer refills)
    uint32_t parity;         // phase parity of `slot`
    long long cnt;           // tiles of this CTA per pass
    long long first, step;   // tile index of the CTA's first tile, distance between its tiles
    int stages;
    bool wrap;

    __device__ __forceinline__ long long tile_at(long long i) const {      // i < 2 cnt except when cnt < stages
        if (i >= cnt) { i -= cnt; if (i >= cnt) i %= cnt; }
        return first + i * step;
    }

    // all threads; barriers must have been initialised (mbar_init + fence) and published by a __syncthreads
    __device__ __forceinline__ void begin(unsigned char* smem, uint64_t* bars, const TileSrc& src, const TileLayout& L) {
        if (tile_issuer(src) && cnt > 0) {
            for (int s = 0; s < stages - 1; ++s) {
                if (!wrap && s >= cnt) break;
                const int sl = (slot + s) % stages;
                tile_issue(src, L, smem + (size_t)sl * L.stage_bytes, &bars[sl], tile_at(s));
            }
        }
    }

    template <class Body>
    __device__ __forceinline__ void pass(unsigned char* smem, uint64_t* bars, const TileSrc& src, const TileLayout& L,
                                         bool wrote_smem, Body body) {
        for (long long i = 0; i < cnt; ++i) {
            if (tile_issuer(src)) {
                const long long ahead = i + (long long)(stages - 1);
                if (wrap || ahead < cnt) tile_issue(src, L, smem + (size_t)pre * L.stage_bytes, &bars[pre], tile_at(ahead));
            }
            mbar_wait(&bars[slot], parity);
            body(smem + (size_t)slot * L.stage_bytes, first + i * step);
            if (wrote_smem) fence_proxy_async();
            __syncthreads();
            pre = slot;
            if (++slot == stages) { slot = 0; parity ^= 1u; }
        }
    }

    // wrap mode, before the CTA exits: wait for the copies that are still in flight (stages-1 tiles)
    __device__ __forceinline__ void drain(uint64_t* bars) {
        if (!wrap || cnt <= 0) return;
        int sl = slot;
        uint32_t par = parity;
        for (int s = 0; s < stages - 1; ++s) {
            mbar_wait(&bars[sl], par);
            if (++sl == stages) { sl = 0; par ^= 1u; }
        }
    }
};

__device__ __forceinline__ void ring_setup(TileRing& r, uint64_t* bars, int stages, long long n_tiles, bool wrap,
                                           uint32_t barrier_count = 1) {
    r.slot = 0; r.pre = stages - 1; r.parity = 0; r.stages = stages; r.wrap = wrap;
    r.first = blockIdx.x; r.step = gridDim.x;
    r.cnt = (long long)blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) mbar_init(&bars[i], barrier_count);
        mbar_fence_init();
    }
}

}  // namespace fn
