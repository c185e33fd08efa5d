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
};

struct TileStage {
    double* y;
    double* aux;
    double* gs;
    double* cst;
    uint8_t* status;
};

__host__ __device__ inline size_t tile_stage_bytes(int genes, int m, bool aux, bool gs, bool cst) {
    size_t b = (size_t)genes * m * 8 * (aux ? 2 : 1) + (size_t)genes * 8 * ((gs ? 1 : 0) + (cst ? 1 : 0)) + genes;
    return (b + 127) & ~(size_t)127;
}

struct TileRing {
    uint64_t* bars;          // FN_MAX_STAGES mbarriers (shared)
    unsigned char* base;     // stage buffers (shared, 128-byte aligned)
    size_t stage_bytes;
    int stages;
    TileSrc src;
    uint32_t tx_bytes;

    __device__ void init(uint64_t* bars_, unsigned char* base_, int stages_, const TileSrc& s) {
        bars = bars_; base = base_; stages = stages_; src = s;
        stage_bytes = tile_stage_bytes(s.genes, s.m, s.aux != nullptr, s.gs != nullptr, s.cst != nullptr);
        const uint32_t row = (uint32_t)s.genes * s.m * 8;
        tx_bytes = row * (s.aux ? 2 : 1) + (uint32_t)s.genes * 8 * ((s.gs ? 1 : 0) + (s.cst ? 1 : 0)) + s.genes;
        if (threadIdx.x == 0) {
            for (int i = 0; i < stages; ++i) mbar_init(&bars[i], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    __device__ TileStage stage(int slot) const {
        TileStage t;
        unsigned char* p = base + (size_t)slot * stage_bytes;
        const size_t row = (size_t)src.genes * src.m * 8;
        t.y = (double*)p; p += row;
        t.aux = nullptr; t.gs = nullptr; t.cst = nullptr;
        if (src.aux) { t.aux = (double*)p; p += row; }
        if (src.gs) { t.gs = (double*)p; p += (size_t)src.genes * 8; }
        if (src.cst) { t.cst = (double*)p; p += (size_t)src.genes * 8; }
        t.status = (uint8_t*)p;
        return t;
    }
    // called by ONE thread: start the copies of `tile` into `slot`
    __device__ void issue(int slot, long long tile) const {
        const TileStage t = stage(slot);
        uint64_t* bar = &bars[slot];
        const long long g0 = tile * src.genes;
        const uint32_t row = (uint32_t)src.genes * src.m * 8;
        mbar_expect_tx(bar, tx_bytes);
        bulk_g2s(t.y, src.y + g0 * src.m, row, bar);
        if (src.aux) bulk_g2s(t.aux, src.aux + g0 * src.m, row, bar);
        if (src.gs) bulk_g2s(t.gs, src.gs + g0, (uint32_t)src.genes * 8, bar);
        if (src.cst) bulk_g2s(t.cst, src.cst + g0, (uint32_t)src.genes * 8, bar);
        bulk_g2s(t.status, src.status + g0, (uint32_t)src.genes, bar);
    }
};

// Persistent tile loop.  `body(stage, tile)` is run by every thread of the CTA for each tile;
// it may write the stage's shared memory (in-place gradient terms).  `wrote_smem` tells the ring
// to fence those generic-proxy writes before the slot is refilled by the TMA engine.
template <class Body>
__device__ __forceinline__ void tile_loop(TileRing& ring, long long n_tiles, bool wrote_smem, Body body) {
    const long long first = blockIdx.x, step = gridDim.x;
    const int S = ring.stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S - 1; ++s) {
            const long long t = first + (long long)s * step;
            if (t < n_tiles) ring.issue(s, t);
        }
    }
    long long it = 0;
    for (long long tile = first; tile < n_tiles; tile += step, ++it) {
        const int slot = (int)(it % S);
        if (threadIdx.x == 0) {
            const long long ahead = tile + (long long)(S - 1) * step;
            if (ahead < n_tiles) ring.issue((int)((it + S - 1) % S), ahead);   // slot freed last iteration
        }
        mbar_wait(&ring.bars[slot], (uint32_t)((it / S) & 1));
        body(ring.stage(slot), tile);
        if (wrote_smem) fence_proxy_async();
        __syncthreads();
    }
}

}  // namespace fn
