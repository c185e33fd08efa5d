// sm_100a kernels of the Fitnoise hot path.  K1 prepare, K2 nll+grad, K3 coef, K3b test,
// noise p-values, weights, aux transform.  (K4, Benjamini-Hochberg, lives in fn_bh.cuh.)
//
// All per-gene kernels share one shape: persistent CTAs, gene tiles staged into shared memory by
// the TMA engine (fn_tile.cuh), a gene handled by `lpg` adjacent lanes (fn_gene.cuh), CTA sums
// reduced with shuffles, and a deterministic grid reduction finished by the last CTA to arrive.
#pragma once
#include <cuda_runtime.h>
#include "fn_gene.cuh"
#include "fn_tile.cuh"

namespace fn {

#ifdef FN_TRACE
// Development-only phase timestamps (globaltimer ns) of CTA 0 and of the last CTA; see tools/micro.
static __device__ unsigned long long g_trace[64];
#define FN_STAMP(slot) do { if (threadIdx.x == 0 && (blockIdx.x == 0)) { unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); g_trace[slot] = t__; } } while (0)
#define FN_STAMP_ANY(slot) do { if (threadIdx.x == 0) { unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); g_trace[slot] = t__; } } while (0)
#else
#define FN_STAMP(slot) do { } while (0)
#define FN_STAMP_ANY(slot) do { } while (0)
#endif

// The per-gene kernels are templates over the design width C.  Every width is instantiated in its
// own translation unit (fn_width.cu, compiled with -DFN_W=<width>) so that the widths build in
// parallel; the C-ABI layer launches them through this table (cudaLaunchKernel).
struct KernelSet {
    const void* prepare;          // prepare_kernel<C>
    const void* eval[4];          // eval_kernel<C, KIND_SCALE1 | KIND_SCALE_PS | KIND_PATSEQ | KIND_PATSEQ_SHARED (C <= 4, else null)>
    const void* eval_lean[2];     // eval_kernel<C, KIND_SCALE1 | KIND_SCALE_PS, true>: all genes complete, no weights (C <= 8, else null)
    const void* eval_factors;     // eval_factors_kernel<C>
    const void* coef;             // coef_kernel<C>
};

// accumulator layout of one CTA's partial sums.  ACC_PSUM: sum of the residual df of the genes that
// contributed -- checked against the preparation kernel's histogram (see StepVars::add_psum).
enum { ACC_VALUE = 0, ACC_DV = 1, ACC_GA = 2, ACC_GB = 3, ACC_USED = 4, ACC_BAD = 5, ACC_PSUM = 6, ACC_FIXED = 7 };

#define FN_MAX_RANKS 16

// Hand-over slots.  A result travels as {bits, tag} with tag = bits ^ seq_mix(seq), written with
// ONE aligned 16-byte store.  Whoever reads a slot (the host spinning on pinned memory, a peer GPU
// polling its mailbox) accepts it when (tag ^ bits) == seq_mix(expected seq).  No fence separates
// "data" from "flag" (each system-scope fence over PCIe / NVLink costs microseconds and a whole
// small-data evaluation takes ~10), and -- unlike a {value, seq} pair -- nothing depends on the
// 16-byte store being single-copy atomic, which the PTX memory model does not promise for a vector
// access: a torn read pairs the bits of one evaluation with the tag of another and fails the
// check unless both evaluations produced the very same bits, in which case the value is right
// anyway.  Only 8-byte single-copy atomicity is assumed (guaranteed for aligned accesses).
struct __align__(16) Slot {
    unsigned long long bits;
    unsigned long long tag;
};

FN_HD unsigned long long seq_mix(unsigned long long seq) { return seq * 0x9E3779B97F4A7C15ull; }   // odd: 0 only for seq 0

__device__ __forceinline__ void slot_store(Slot* s, double value, unsigned long long seq) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(value);
    asm volatile("st.global.v2.b64 [%0], {%1, %2};" ::"l"(s), "l"(bits), "l"(bits ^ seq_mix(seq)) : "memory");
}
// true (and the value) once the slot carries evaluation `seq`
__device__ __forceinline__ bool slot_try_load(const Slot* s, unsigned long long seq, double& value) {
    unsigned long long b, t;
    asm volatile("ld.volatile.global.v2.b64 {%0, %1}, [%2];" : "=l"(b), "=l"(t) : "l"(s) : "memory");
    value = __longlong_as_double((long long)b);
    return (b ^ t) == seq_mix(seq);
}

// Cross-GPU exchange fused into the kernel epilogue (multi-GPU, one process per GPU).
// Every rank owns a mailbox in its device memory: [parity 2][sender world][slot_count] Slots.  All
// ranks' mailboxes are mapped into every process with CUDA IPC, so the last CTA of rank r stores
// its sums straight into every peer's mailbox over NVLink (16-byte stores on the peer mapping),
// polls its OWN mailbox until every sender's slots carry this evaluation's seq, and adds them in
// rank order (so every rank computes bit-identical totals).  No NCCL call, no extra launch, no
// host round trip.  Two parities: a fast rank may already write evaluation k+1 while a slow one
// still reads k.
struct PeerComm {
    int world, rank;                       // world <= 1: no exchange
    int slot_count;                        // slots per (parity, sender)
    unsigned long long timeout_ns;         // give up waiting for a peer after this long
    Slot* mailbox[FN_MAX_RANKS];           // mailbox[r]: rank r's mailbox as mapped in this process
};

struct GridReduce {
    Slot* pslots;            // gridDim.x * FN_SLOT_ACC authenticated partials (nacc <= FN_SLOT_ACC)
    double* partials;        // gridDim.x * nacc
    unsigned int* ticket;
    double* out_dev;         // nacc doubles
    Slot* out_host;          // mapped pinned: nacc slots, or nullptr
    int nacc;
    PeerComm comm;
};

// What changes from one evaluation to the next (kernel parameters of a one-evaluation launch, the
// command block of the resident kernel).
struct StepVars {
    unsigned long long host_seq;   // seq of this evaluation in the out_host slots
    unsigned long long comm_seq;   // evaluation number of the exchange, identical on every rank, >= 1
    int add_on;                    // 1: add the three terms below to this rank's sums before the exchange
    double add_value;              // theta-independent part of the objective + sum_p N_p t0[p]
    double add_dv;                 // sum_p N_p t1[p]
    double add_psum;               // -sum_p p N_p: the exchanged ACC_PSUM total must come out as exactly 0
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// CTA-wide sum of NV thread-local doubles into shared `out[NV]` (deterministic order).
template <int NV>
__device__ __forceinline__ void block_sum(double* v, double* out, double* scratch /* 32*NV */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = v[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        if (lane == 0) scratch[warp * NV + i] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < nwarp; ++w) s += scratch[w * NV + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

#define FN_SLOT_ACC 8      // grids of up to this many accumulators reduce through authenticated slots

// Reducer side of the slot form of grid_reduce (cold code, kept out of line so that its registers do
// not weigh on the streaming loop): thread t adds the partials of CTAs t, t + T, ... into s[].
__device__ __forceinline__ void gather_slots(const Slot* pslots, int nacc, unsigned long long seq, double* s) {
    const unsigned int G = gridDim.x, T = blockDim.x;
#pragma unroll
    for (int i = 0; i < FN_SLOT_ACC; ++i) s[i] = 0.0;
    const unsigned long long t0 = global_ns();
    bool failed = false;
    for (unsigned int b = threadIdx.x; b < G; b += T) {
        const Slot* line = &pslots[(size_t)b * FN_SLOT_ACC];
        double v[FN_SLOT_ACC];
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int i = 0; i < FN_SLOT_ACC; ++i) {
                v[i] = 0.0;
                if (i < nacc) ok = slot_try_load(line + i, seq, v[i]) && ok;
            }
            if (ok) break;
            if (global_ns() - t0 > 10000000000ull) { failed = true; break; }     // a CTA died: NaN totals
        }
#pragma unroll
        for (int i = 0; i < FN_SLOT_ACC; ++i) s[i] += failed ? NAN : v[i];
    }
}


// Sum `nacc` per-CTA values over the grid in a fixed order, (multi-GPU) exchange the sums with the
// peer GPUs, and hand the totals to the host.  `vals` is this CTA's shared array of nacc doubles
// (complete and visible after a barrier).
//
// nacc <= FN_SLOT_ACC (every model without per-sample parameters): each CTA stores its partials as
// authenticated slots -- one 128-byte line per CTA, no fence, no atomic -- and is done.  CTA 0 is the
// reducer: thread t polls the lines of CTAs t, t + T, ... (all loads of a line in flight together,
// re-read until every slot carries this evaluation's seq), the thread sums are added by the fixed
// tree of block_sum: deterministic for a given grid, and the last partial to arrive is seen one L2
// round trip after it was stored.  (The slower alternative below -- partials, fence, ticket, and the
// last CTA to arrive re-reads everything -- costs the slowest CTA about 2.5 us more.)
// Larger nacc keeps the ticket form.  Either way the function is safe to call once per evaluation
// from a resident kernel: nothing of evaluation k can be mistaken for evaluation k+1.
__device__ __forceinline__ void grid_reduce(const GridReduce& gr, const StepVars& sv, const double* vals) {
    __shared__ int is_last;
    __shared__ double gsum[256];
    __shared__ double tot[FN_SLOT_ACC];
    static_assert(32 * FN_SLOT_ACC <= 256, "gsum doubles as block_sum scratch");
    const int T = blockDim.x, nacc = gr.nacc;
    const unsigned int G = gridDim.x;
    const PeerComm& pc = gr.comm;
    const bool exchange = pc.world > 1;
    const int parity = (int)(sv.comm_seq & 1ull);
    // total of component i over the grid -> peers' mailboxes, or device + host
    auto publish = [&](int i, double s) {
        if (sv.add_on) {
            if (i == ACC_VALUE) s += sv.add_value;
            else if (i == ACC_DV) s += sv.add_dv;
            else if (i == ACC_PSUM) s += sv.add_psum;
        }
        if (exchange) {
            const size_t mine = ((size_t)parity * pc.world + pc.rank) * pc.slot_count + i;
            for (int r = 0; r < pc.world; ++r) slot_store(&pc.mailbox[r][mine], s, sv.comm_seq);
        } else {
            gr.out_dev[i] = s;
            if (gr.out_host) slot_store(&gr.out_host[i], s, sv.host_seq);
        }
    };

    if (nacc <= FN_SLOT_ACC) {
        if (threadIdx.x < nacc)
            slot_store(&gr.pslots[(size_t)blockIdx.x * FN_SLOT_ACC + threadIdx.x], vals[threadIdx.x], sv.host_seq);
        if (blockIdx.x != 0) return;
        FN_STAMP_ANY(8);
        double s[FN_SLOT_ACC];
        gather_slots(gr.pslots, nacc, sv.host_seq, s);
        block_sum<FN_SLOT_ACC>(s, tot, gsum);
        if (threadIdx.x < nacc) publish(threadIdx.x, tot[threadIdx.x]);
        FN_STAMP_ANY(9);
    } else {
        for (int i = threadIdx.x; i < nacc; i += T)
            gr.partials[(size_t)blockIdx.x * nacc + i] = vals[i];
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int t = atomicAdd(gr.ticket, 1u);
            is_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (!is_last) return;
        FN_STAMP_ANY(8);
        if (threadIdx.x == 0) *gr.ticket = 0;
        __threadfence();
        __syncthreads();
        // Warp w takes components w, w + W, ... (eight at a time); its lanes add the partials of CTAs
        // lane, lane + 32, ... and a xor-shuffle tree combines them: a fixed order for a given grid.
        // All loads of a batch are independent, so the reduction costs a few L2 round trips -- the
        // earlier form (one thread per component walking the CTAs one dependent iteration at a time)
        // took 27 us at nacc = 103 on 148 CTAs, a quarter of a per-sample evaluation at 200k x 48.
        {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = T >> 5;
            constexpr int NB = 8;
            for (int i0 = warp; i0 < nacc; i0 += NB * W) {
                double a[NB];
#pragma unroll
                for (int k = 0; k < NB; ++k) a[k] = 0.0;
#pragma unroll 2
                for (unsigned int b = lane; b < G; b += 32) {
                    const double* row = gr.partials + (size_t)b * nacc;
#pragma unroll
                    for (int k = 0; k < NB; ++k) {
                        const int i = i0 + k * W;
                        if (i < nacc) a[k] += __ldcg(row + i);
                    }
                }
#pragma unroll
                for (int k = 0; k < NB; ++k) {
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], off);
                }
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                        if (i0 + k * W < nacc) publish(i0 + k * W, a[k]);
                }
            }
        }
        FN_STAMP_ANY(9);
    }
    if (!exchange) return;

    // Every component: wait for all senders' slots of this evaluation in my own mailbox, add in rank
    // order.  A peer that never shows up turns EVERY total into NaN, which the host reports as an error.
    // Thread (component, sender) polls ONE slot, so all slots are waited for at the same time (the
    // earlier form had one thread per component walk the senders one L2 round trip after the other);
    // the values meet in shared memory (256 / world components per round).
    const int W = pc.world, per = 256 / W;
    for (int c0 = 0; c0 < nacc; c0 += per) {
        const int cnt = (nacc - c0 < per) ? nacc - c0 : per;
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * W; e += T) {
            const int k = e / W, r = e - k * W;
            const Slot* slot = &pc.mailbox[pc.rank][((size_t)parity * W + r) * pc.slot_count + c0 + k];
            const unsigned long long t0 = global_ns();
            double got;
            while (!slot_try_load(slot, sv.comm_seq, got)) {
                if (global_ns() - t0 > pc.timeout_ns) { got = NAN; break; }
            }
            gsum[e] = got;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < cnt; k += T) {
            double s = 0.0;
            for (int r = 0; r < W; ++r) s += gsum[k * W + r];          // rank order: the same total on every rank
            gr.out_dev[c0 + k] = s;
            if (gr.out_host) slot_store(&gr.out_host[c0 + k], s, sv.host_seq);
        }
    }
}

// Per-evaluation tables t0[p], t1[p] (p = 0..m) of density_tables(), filled by the whole CTA: thread p
// evaluates ln Gamma and psi of (v + p)/2 with one lgamma_digamma() call (two logarithms, one
// division), the p-independent terms come from entry 0.  This sits on the critical path of every
// launch (~2-3 us of a ~10 us small-data kernel with the library lgamma and a ten-division psi).
__device__ __forceinline__ void fill_density_tables(int is_t, double v, int m, double* t0, double* t1) {
    __shared__ double common[3];
    if (!is_t) {
        for (int p = threadIdx.x; p <= m; p += blockDim.x) density_tables(0, v, p, t0[p], t1[p]);
        __syncthreads();
        return;
    }
    // entry p (thread p): ln Gamma and psi of (v + p)/2 in one call; entry 0 is also the common v/2 term
    for (int p = threadIdx.x; p <= m; p += blockDim.x) lgamma_digamma(0.5 * (v + p), t0[p], t1[p]);
    if (threadIdx.x == blockDim.x - 1) common[2] = log(3.14159265358979323846 * v);
    __syncthreads();
    if (threadIdx.x == 0) { common[0] = t0[0]; common[1] = t1[0]; }
    __syncthreads();
    for (int p = threadIdx.x; p <= m; p += blockDim.x) {
        t0[p] = -(t0[p] - common[0] - (0.5 * p) * common[2]);
        t1[p] = -0.5 * t1[p] + 0.5 * common[1] + 0.5 * p / v;
    }
    __syncthreads();
}

// Copy an m x C design table from global memory into shared memory with the padded row stride.
template <int C>
__device__ __forceinline__ void stage_design(double* dst, const double* src, int m) {
    constexpr int XS = XStride<C>::V;
    for (int i = threadIdx.x; i < m * C; i += blockDim.x) {
        const int j = i / C, k = i - j * C;
        dst[j * XS + k] = src[i];
    }
}

// ------------------------------------------------------------------------------------------
// shared-memory carve-up common to the per-gene kernels
// ------------------------------------------------------------------------------------------
struct GeomParams {
    long long n;             // genes
    long long n_tiles;
    int m;
    int genes;               // genes per tile  (= blockDim.x / lpg)
    int lpg;
    int stages;
    int width;               // template width C the designs are padded to
    int rows;                // 1: thread per gene on padded rows (lpg == 1, m even; ps_rows_* in fn_gene.cuh)
};

// dynamic shared memory of the per-gene kernels: [stage buffers][tables].  The stage area comes
// first so that every pointer into it is plain arithmetic on the __shared__ symbol (the compiler
// then emits LDS, not generic loads).

// ------------------------------------------------------------------------------------------
// K1 prepare
// ------------------------------------------------------------------------------------------
enum { PREP_VARSUM = 0, PREP_VARCNT = 1, PREP_BAD = 2, PREP_CSTSUM = 3, PREP_ELIG = 4, PREP_PARTIAL = 5, PREP_NACC = 6 };
// PREP_PARTIAL: eligible genes with fewer than m retained samples (0 => the lean evaluation kernels apply)

struct PrepParams {
    GeomParams g;
    TileSrc src;             // y, aux = RAW second array, status = control bits
    Family fam;
    const double* xn; const double* xc;     // m x width each
    int c_noise, c_ctrl;
    uint8_t* status_out;     // n_pad
    double* cst_out;         // n_pad
    double* avg_out;         // n_pad
    double* gs_out;          // n_pad or nullptr (PATSEQ: avgtail^2)
    unsigned long long* p_hist;   // m + 1 counters (zeroed by the caller): eligible genes by residual df k - c
    GridReduce gr;
    StepVars sv;
};

template <int C>
__global__ void __launch_bounds__(256) prepare_kernel(const PrepParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[FN_MAX_STAGES];
    __shared__ double red_out[PREP_NACC];
    __shared__ double red_scratch[32 * PREP_NACC];
    const int m = p.g.m;
    const TileLayout L = tile_layout(p.src);
    tile_prologue(smem_raw, bars, p.g.stages, p.src, L, p.g.n_tiles);
    double* xn = (double*)(smem_raw + (size_t)p.g.stages * L.stage_bytes);
    constexpr int XS = XStride<C>::V;
    double* xc = xn + m * XS;
    unsigned int* hist = (unsigned int*)(xc + m * XS);          // m + 1 counters of this CTA
    stage_design<C>(xn, p.xn, m);
    stage_design<C>(xc, p.xc, m);
    for (int i = threadIdx.x; i <= m; i += blockDim.x) hist[i] = 0u;
    __syncthreads();

    const int gl = threadIdx.x / p.g.lpg, lane = threadIdx.x % p.g.lpg;
    GeneView gv;
    gv.setup(m, lane, p.g.lpg, gl);
    gv.gs = 0.0;
    double acc[PREP_NACC];
#pragma unroll
    for (int i = 0; i < PREP_NACC; ++i) acc[i] = 0.0;

    tile_loop(smem_raw, bars, p.g.stages, p.src, L, p.g.n_tiles, false, [&](unsigned char* st, long long tile) {
        const long long gene = tile * p.g.genes + gl;
        gv.y = (double*)st + (size_t)gl * m;
        gv.aux = p.src.aux ? (double*)(st + L.aux_off) + (size_t)gl * m : nullptr;
        const bool control = (st[L.status_off + gl] & 2) != 0;
        const int c_real = control ? p.c_ctrl : p.c_noise;
        PrepOut po;
        prepare_gene<C>(p.fam, gv, control ? xc : xn, c_real, po);
        if (lane == 0 && gene < p.g.n) {
            p.status_out[gene] = (uint8_t)((po.eligible ? 1 : 0) | (control ? 2 : 0));
            p.cst_out[gene] = po.cst;
            p.avg_out[gene] = po.average;
            if (p.gs_out) p.gs_out[gene] = po.avg_tail * po.avg_tail;
            if (po.nfinite > 0) { acc[PREP_VARSUM] += po.yss / po.nfinite; acc[PREP_VARCNT] += 1.0; }
            acc[PREP_BAD] += po.bad_counts;
            if (po.eligible) {
                acc[PREP_CSTSUM] += po.cst; acc[PREP_ELIG] += 1.0;
                if (po.k < m) acc[PREP_PARTIAL] += 1.0;
                atomicAdd(&hist[po.k - c_real], 1u);            // 0 < k - c <= m
            }
        }
    });
    __syncthreads();
    for (int i = threadIdx.x; i <= m; i += blockDim.x)
        if (hist[i]) atomicAdd(&p.p_hist[i], (unsigned long long)hist[i]);
    block_sum<PREP_NACC>(acc, red_out, red_scratch);
    grid_reduce(p.gr, p.sv, red_out);
}

#ifndef FN_WIDTH_ONLY
// PATSEQ: raw counts -> rc = 1/max(count, 1)   (fitnoise.py:648,749)
__global__ void counts_to_rc_kernel(double* aux, long long total) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
        aux[i] = (aux[i] == aux[i]) ? 1.0 / fmax(aux[i], 1.0) : NAN;       // numpy.maximum keeps NaN: the sample is dropped
}
#endif

// ------------------------------------------------------------------------------------------
// K2 -log likelihood + gradient
// ------------------------------------------------------------------------------------------
// The kernel runs in one of two forms:
//   * one launch = one evaluation (rc.cmd_host == nullptr): theta and the step variables are kernel
//     parameters;
//   * RESIDENT (rc.cmd_host != nullptr): launched once per optimisation, it stays on the GPU and
//     executes one evaluation per COMMAND that the host writes into a block of mapped pinned
//     memory.  An evaluation then costs no launch call and no launch latency: CTA 0 polls the
//     command block over PCIe and re-publishes it in device memory, where every CTA polls it (L2);
//     every CTA already holds its first tiles in shared memory (TileRing, wrap mode).
//     The kernel leaves on a QUIT command, or by itself after `idle_timeout_ns` without a command
//     (the host notices and relaunches), so a host that dies or forgets it cannot pin the GPU.
// Command block (64-bit words; the host writes the sequence number last, the checksum covers every
// other word, so a half-written block is never accepted):
enum { CW_SEQ = 0, CW_OP = 1, CW_COMM_SEQ = 2, CW_V = 3, CW_THETA0 = 4, CW_ADD_VALUE = 5, CW_ADD_DV = 6,
       CW_ADD_PSUM = 7, CW_TA0 = 8, CW_TB0 = 9, CW_TAB = 10 };      // then ntab table words, then the checksum
enum { OP_EVAL = 1, OP_QUIT = 2, OP_QUIT_IDLE = 3, OP_FLUSH = 4 };   // FLUSH: sweep a buffer larger than the L2 (benchmarks)
FN_HD unsigned long long cmd_mix(unsigned long long w, int i) {
    return (w + 0x632BE59BD9B4E019ull) * (0x9E3779B97F4A7C15ull + 2ull * (unsigned long long)i);
}

struct ResidentCtl {
    const unsigned long long* cmd_host;   // device pointer of the mapped pinned command block, or nullptr
    Slot* cmd_dev;                        // device memory: the current command, one authenticated slot per word
    unsigned long long* exit_host;        // mapped pinned: set to the expected seq when the kernel leaves on idle
    unsigned long long next_seq;          // seq of the first command this launch handles
    unsigned long long idle_timeout_ns;
    int words;                            // command length in words, checksum included
    double* flush_buf;                    // OP_FLUSH: buffer larger than the L2, or nullptr
    unsigned long long flush_words;
};

// Read-modify-write a buffer larger than the L2 so that the next evaluation streams its input from
// DRAM (bench.py: "flush L2 between timed iterations" when a shard would otherwise stay L2-resident).
__device__ __forceinline__ void l2_flush_sweep(double* buf, unsigned long long words) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += stride)
        buf[i] = buf[i] + 1.0;
}
#ifndef FN_WIDTH_ONLY
__global__ void l2_flush_kernel(double* buf, unsigned long long words) { l2_flush_sweep(buf, words); }
#endif

__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void slot_store_bits(Slot* s, unsigned long long bits, unsigned long long seq) {
    asm volatile("st.global.v2.b64 [%0], {%1, %2};" ::"l"(s), "l"(bits), "l"(bits ^ seq_mix(seq)) : "memory");
}
__device__ __forceinline__ bool slot_try_load_bits(const Slot* s, unsigned long long seq, unsigned long long& bits) {
    unsigned long long t;
    asm volatile("ld.volatile.global.v2.b64 {%0, %1}, [%2];" : "=l"(bits), "=l"(t) : "l"(s) : "memory");
    return (bits ^ t) == seq_mix(seq);
}

// Warp 0 of CTA 0: wait for command `expect` in host memory and re-publish it in device memory as one
// authenticated slot per word (every CTA polls those: no flag, no fence -- a slot proves by itself
// that it belongs to `expect`).  The poll reads the whole block with one load per lane, so a command
// of up to 32 words (every model without per-sample parameters) is recognised, verified and
// forwarded after a single PCIe round trip.
__device__ __forceinline__ void resident_fetch(const ResidentCtl& rc, unsigned long long expect) {
    const int lane = threadIdx.x & 31;
    const int nw = rc.words;
    const unsigned long long t0 = global_ns();
    for (;;) {
        const unsigned long long w = lane < nw ? ld_sys_u64(rc.cmd_host + lane) : 0ull;
        const unsigned long long s = __shfl_sync(0xffffffffu, w, 0);
        if (s == expect) {
            unsigned long long sum = (lane < nw - 1) ? cmd_mix(w, lane) : 0ull;
            unsigned long long chk = (lane == nw - 1) ? w : 0ull;
            for (int base = 32; base < nw; base += 128) {          // long commands: four loads in flight per lane
                unsigned long long x[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { const int i = base + 32 * k + lane; x[k] = i < nw ? ld_sys_u64(rc.cmd_host + i) : 0ull; }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = base + 32 * k + lane;
                    if (i < nw - 1) sum += cmd_mix(x[k], i);
                    else if (i == nw - 1) chk = x[k];
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                sum += __shfl_xor_sync(0xffffffffu, sum, off);
                chk += __shfl_xor_sync(0xffffffffu, chk, off);
            }
            if (sum == chk) {
#ifdef FN_TRACE
                if (__shfl_sync(0xffffffffu, w, CW_OP) == (unsigned long long)OP_EVAL) FN_STAMP(10);
#endif
                // forward: the first 32 words from the registers, the rest re-read (only per-sample models)
                if (lane < nw - 1) slot_store_bits(rc.cmd_dev + lane, w, expect);
                for (int i = 32 + lane; i < nw - 1; i += 32) slot_store_bits(rc.cmd_dev + i, ld_sys_u64(rc.cmd_host + i), expect);
                return;
            }
            continue;                                  // the host is still writing: read it again
        }
        if (global_ns() - t0 > rc.idle_timeout_ns) break;
    }
    // idle: tell the host, then send every CTA home with a complete (fabricated) command
    if (lane == 0) *rc.exit_host = expect;
    for (int i = lane; i < nw - 1; i += 32) slot_store_bits(rc.cmd_dev + i, i == CW_OP ? (unsigned long long)OP_QUIT_IDLE : 0ull, expect);
}

// Every CTA: block until command `expect` is in cmd_dev; thread t takes words t, t + T, ...: the header
// words go to `hdr` (shared, CW_TAB words), the per-sample tables straight into ta / tb.  Returns
// the opcode (uniform over the CTA).
__device__ __forceinline__ int resident_wait(const ResidentCtl& rc, unsigned long long expect,
                                             unsigned long long* hdr, double* ta, double* tb, int m, bool has_ta) {
    if (blockIdx.x == 0 && threadIdx.x < 32) resident_fetch(rc, expect);
    int timed_out = 0;
    const unsigned long long t0 = global_ns();
    for (int i = threadIdx.x; i < rc.words - 1; i += blockDim.x) {
        unsigned long long bits = 0;
        while (!slot_try_load_bits(rc.cmd_dev + i, expect, bits)) {
            // CTA 0 always publishes within idle_timeout_ns; this bound only matters if CTA 0 is gone
            if (global_ns() - t0 > rc.idle_timeout_ns + 5000000000ull) { timed_out = 1; break; }
            __nanosleep(20);
        }
        if (i < CW_TAB) hdr[i] = bits;
        else if (has_ta && i - CW_TAB < m) ta[i - CW_TAB] = __longlong_as_double((long long)bits);
        else tb[i - CW_TAB - (has_ta ? m : 0)] = __longlong_as_double((long long)bits);
    }
    if (__syncthreads_or(timed_out)) return OP_QUIT_IDLE;
    return (int)hdr[CW_OP];
}

// Per-evaluation tables of the complete-row fast path of KIND_SCALE_PS (fn_gene.cuh, PsDesign), built by
// the whole CTA for one design: A = X'WX (thread per packed entry), its inverse (one thread: c <= 16),
// then one table row per thread.  `tri`: Tri<C>::N doubles of scratch, `sc`: 3 doubles.
template <int C>
__device__ __forceinline__ void build_ps_design(const double* X, int c_real, const double* w, const double* lt, int m,
                                                double* hx, int hx_stride, double* cj, double* tri, double* sc) {
    constexpr int N = Tri<C>::N;
    for (int e = threadIdx.x; e < N; e += blockDim.x) {
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        const int k = e - i * (i + 1) / 2;
        double a = 0.0;
        for (int j = 0; j < m; ++j) a += w[j] * X[j * C + i] * X[j * C + k];
        tri[e] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a[N], ainv[N];
        for (int e = 0; e < N; ++e) a[e] = tri[e];
        for (int i = c_real; i < C; ++i) a[FN_TRI(i, i)] = 1.0;
        double det;
        const bool ok = chol<C>(a, det, 0.0);
        chol_inverse<C>(a, ainv);
        for (int e = 0; e < N; ++e) tri[e] = ainv[e];
        double slt = 0.0;
        for (int j = 0; j < m; ++j) slt += lt[j];
        sc[0] = log(det); sc[1] = slt; sc[2] = ok ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        double x[C];
        load_x<C>(X, j, x);
        ps_table_row<C>(tri, x, w[j], hx + j * hx_stride, cj[j]);
    }
    __syncthreads();
}

struct EvalParams {
    GeomParams g;
    TileSrc src;             // y, aux (precision weights / rc), gs, cst (per-gene mode only), status
    const double* xn; const double* xc;
    int c_noise, c_ctrl;
    int has_ctrl;            // 0: no gene carries the control bit (the control design's per-evaluation tables are skipped)
    int kind, tail_own, v3, is_t;
    double v, theta0;        // one-launch form (resident: from the command)
    double ta0, tb0;         // shared theta_a / theta_b (general kinds); per-sample blocks come from ta / tb
    const double* ta; const double* tb;     // device tables of length m, or nullptr (shared: ta0 / tb0)
    int inplace_a, inplace_b;               // per-sample theta blocks
    double* pg_score; double* pg_d2; int* pg_p;     // per-gene outputs (nullptr in sum mode)
    GridReduce gr;           // nacc = ACC_FIXED (+ 2m when a block is per-sample)
    StepVars sv;             // one-launch form
    ResidentCtl rc;
};

template <int C, int KIND, bool LEAN = false>
__global__ void __launch_bounds__(256, ((C <= 4 || (LEAN && C <= 8)) ? 2 : 1)) eval_kernel(const EvalParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[FN_MAX_STAGES];
    __shared__ double red_scratch[32 * ACC_FIXED];
    __shared__ unsigned long long cmd_hdr[CW_TAB];     // resident form: header words of the current command
    const int m = p.g.m;
    const bool resident = p.rc.cmd_host != nullptr;
    const bool per_gene = p.pg_score != nullptr;
    FN_STAMP(0);
    const TileLayout L = tile_layout(p.src);
    TileRing ring;
    ring_setup(ring, bars, p.g.stages, p.g.n_tiles, resident, tile_barrier_count(p.src));
    // tables behind the stage area: xn, xc (m*C each), [hxn, hxc (m*C each): KIND_SCALE_PS], ta, tb (m each),
    // t0, t1 (m+1 each), vals (ACC_FIXED + 2m), colscr (2*blockDim when a theta block is per-sample).
    // The design-shaped tables come first: their rows must stay 16-byte aligned for LDS.128 (the stage
    // area is a multiple of 128 bytes, rows of an even width are multiples of 16).
    double* xn = (double*)(smem_raw + (size_t)p.g.stages * L.stage_bytes);
    constexpr int XS = XStride<C>::V;
    double* xc = xn + m * XS;
    double* hxn = xc + m * XS;
    double* hxc = hxn + (KIND == KIND_SCALE_PS ? m * XS : 0);
    double* ta = hxc + (KIND == KIND_SCALE_PS ? m * XS : 0);
    double* tb = ta + m;
    double* t0 = tb + m;
    double* t1 = t0 + (m + 1);
    double* vals = t1 + (m + 1);
    double* colscr = vals + (ACC_FIXED + 2 * m);
    const bool inplace = p.inplace_a || p.inplace_b;
    // KIND_SCALE_PS without weights: tables of the complete-row fast path (two designs), kappa per gene
    const bool fastps = KIND == KIND_SCALE_PS && p.src.aux == nullptr;
    double* cjn = colscr + (inplace ? 2 * blockDim.x : 0);
    double* cjc = cjn + m;
    double* pstri = cjc + m;
    double* pssc = pstri + Tri<C>::N;          // 2 x 3
    double* kap = pssc + 8;                    // genes per tile (<= 256)
    // rows layout: the two designs' tables interleaved row by row (fn_gene.cuh load_x2): hx2 re-uses the
    // hxn / hxc area (2 m C doubles), x2 is a second copy of the designs
    const bool rows_tables = fastps && p.g.rows != 0;
    double* x2 = kap + 256;
    x2 += (smem_u32(x2) >> 3) & 1u;           // 16-byte aligned (LDS.128 rows); the carve-up has the slack
    const int hx_stride = rows_tables ? 2 * C : XS;
    double* hxc_at = rows_tables ? hxn + C : hxc;
    __shared__ int ps_count[2];
    stage_design<C>(xn, p.xn, m);
    stage_design<C>(xc, p.xc, m);
    if (rows_tables)
        for (int i = threadIdx.x; i < m * C; i += blockDim.x) {
            const int j = i / C, k = i - j * C;
            x2[j * 2 * C + k] = p.xn[i];
            x2[j * 2 * C + C + k] = p.xc[i];
        }
    __syncthreads();                                   // barrier initialisation + designs visible
    ring.begin(smem_raw, bars, p.src, L);              // the first tiles are in flight
    FN_STAMP(1);

    const int gl = threadIdx.x / p.g.lpg, lane = threadIdx.x % p.g.lpg;
    GeneView gv;
    gv.setup(m, lane, p.g.lpg, gl);
    // per-sample gradient columns: thread (sub, col) sums column col over genes sub, sub+nsub, ...
    const int nsub = inplace ? blockDim.x / m : 0;
    const int col = inplace ? threadIdx.x % m : 0, sub = inplace ? threadIdx.x / m : 0;
    // rows layout (thread per gene, m even): thread (sub2, colp) sums the column PAIR colp over genes sub2, sub2+nsub2, ...
    const int nsub2 = inplace ? (int)blockDim.x / (m >> 1 ? m >> 1 : 1) : 0;
    const int colp = inplace ? (int)threadIdx.x % (m >> 1 ? m >> 1 : 1) : 0, sub2 = inplace ? (int)threadIdx.x / (m >> 1 ? m >> 1 : 1) : 0;

    unsigned long long expect = p.rc.next_seq;
    for (;;) {
        // ---- this evaluation's theta and step variables -----------------------------------
        EvalConst ec;
        StepVars sv;
        if (resident) {
            // (the command's per-sample tables land in ta / tb directly)
            const int op = resident_wait(p.rc, expect, cmd_hdr, ta, tb, m, p.ta != nullptr);
            if (op == OP_FLUSH) {
                if (p.rc.flush_buf) l2_flush_sweep(p.rc.flush_buf, p.rc.flush_words);
                __threadfence();
                __syncthreads();
                if (threadIdx.x == 0) {
                    if (atomicAdd(p.gr.ticket, 1u) == gridDim.x - 1) {      // last CTA: tell the host
                        *p.gr.ticket = 0;
                        __threadfence();
                        slot_store(&p.gr.out_host[0], 0.0, expect);
                    }
                }
                ++expect;
                continue;
            }
            if (op != OP_EVAL) break;
            FN_STAMP(11);
            ec.v = __longlong_as_double((long long)cmd_hdr[CW_V]);
            ec.theta0 = __longlong_as_double((long long)cmd_hdr[CW_THETA0]);
            sv.host_seq = expect;
            sv.comm_seq = cmd_hdr[CW_COMM_SEQ];
            sv.add_on = 1;
            sv.add_value = __longlong_as_double((long long)cmd_hdr[CW_ADD_VALUE]);
            sv.add_dv = __longlong_as_double((long long)cmd_hdr[CW_ADD_DV]);
            sv.add_psum = __longlong_as_double((long long)cmd_hdr[CW_ADD_PSUM]);
            if (KIND != KIND_SCALE1) {
                const double ta0 = __longlong_as_double((long long)cmd_hdr[CW_TA0]);
                const double tb0 = __longlong_as_double((long long)cmd_hdr[CW_TB0]);
                for (int i = threadIdx.x; i < m; i += blockDim.x) {
                    if (!p.ta) ta[i] = ta0;
                    if (!p.tb) tb[i] = tb0;
                }
            }
        } else {
            ec.v = p.v; ec.theta0 = p.theta0;
            sv = p.sv;
            if (KIND != KIND_SCALE1) {
                for (int i = threadIdx.x; i < m; i += blockDim.x) {
                    ta[i] = p.ta ? __ldcg(p.ta + i) : p.ta0;
                    tb[i] = p.tb ? __ldcg(p.tb + i) : p.tb0;
                }
            }
        }
        if (per_gene) fill_density_tables(p.is_t, ec.v, m, t0, t1);      // ends with a barrier
        else if (KIND != KIND_SCALE1) __syncthreads();
        PsDesign<C> pdn, pdc;
        if (fastps) {
            build_ps_design<C>(xn, p.c_noise, ta, tb, m, hxn, hx_stride, cjn, pstri, pssc);
            pdn.hx = hxn; pdn.cj = cjn; pdn.logdet_a = pssc[0]; pdn.slt = pssc[1]; pdn.ok = pssc[2] != 0.0;
            if (p.has_ctrl) {                  // control genes present: the second design's tables
                build_ps_design<C>(xc, p.c_ctrl, ta, tb, m, hxc_at, hx_stride, cjc, pstri, pssc + 3);
                pdc.hx = hxc_at; pdc.cj = cjc; pdc.logdet_a = pssc[3]; pdc.slt = pssc[4]; pdc.ok = pssc[5] != 0.0;
            } else {
                pdc = pdn;
                for (int j = threadIdx.x; j < m; j += blockDim.x) cjc[j] = 0.0;
            }
            if (threadIdx.x < 2) ps_count[threadIdx.x] = 0;
            __syncthreads();
        }
        FN_STAMP(2);
        ec.kind = (KIND == KIND_PATSEQ_SHARED) ? KIND_PATSEQ : KIND; ec.tail_own = p.tail_own; ec.v3 = p.v3; ec.is_t = p.is_t;
        ec.ta = ta; ec.tb = tb;
        ec.t0 = per_gene ? t0 : nullptr; ec.t1 = per_gene ? t1 : nullptr;
        ec.finish_setup();

        double acc[ACC_FIXED];
#pragma unroll
        for (int i = 0; i < ACC_FIXED; ++i) acc[i] = 0.0;
        double col_a = 0.0, col_b = 0.0;

        // accumulate one finished gene into the thread's partial sums / the per-gene outputs
        auto account = [&](const EvalOut& eo, long long gene) {
            if (gene >= p.g.n) return;
            if (eo.used) {
                if (eo.value == eo.value) {
                    acc[ACC_VALUE] += eo.value; acc[ACC_DV] += eo.dv;
                    acc[ACC_GA] += eo.ga; acc[ACC_GB] += eo.gb; acc[ACC_USED] += 1.0;
                    acc[ACC_PSUM] += (double)eo.p;
                } else {
                    acc[ACC_BAD] += 1.0;
                }
            }
            if (per_gene) {
                p.pg_score[gene] = eo.used ? eo.value : NAN;
                p.pg_d2[gene] = eo.d2;
                p.pg_p[gene] = eo.p;
            }
        };

        if constexpr (KIND == KIND_SCALE1) {
            // The lpg lanes of a gene all end the sweep holding the same (p, d2).  Instead of every
            // lane running the density tail (log1p, divisions) for the same gene, lane l keeps the
            // gene of every lpg-th tile and the tail runs once per lpg tiles with all 32 lanes of a
            // warp working on different genes.
            Scale1Mid pend;
            pend.state = 0; pend.d2 = 0.0; pend.det_a = 1.0; pend.p = 0;
            double pend_cst = 0.0;
            long long pend_gene = -1;
            int sel = 0;
            auto finalize = [&]() {
                if (pend_gene >= 0) {
                    EvalOut eo;
                    scale1_finish(ec, pend, pend_cst, eo);
                    account(eo, pend_gene);
                    pend_gene = -1;
                }
            };
            ring.pass(smem_raw, bars, p.src, L, false, [&](unsigned char* st, long long tile) {
                gv.y = (double*)st + gl * m;
                gv.aux = p.src.aux ? (double*)(st + L.aux_off) + gl * m : nullptr;
                const uint8_t status = st[L.status_off + gl];
                const bool control = (status & 2) != 0;
                Scale1Mid mid;
                scale1_sweep<C, LEAN>(ec, gv, control ? xc : xn, control ? p.c_ctrl : p.c_noise, (status & 1) != 0, mid);
                if (lane == sel) {
                    pend = mid;
                    pend_cst = p.src.cst ? ((const double*)(st + L.cst_off))[gl] : 0.0;
                    pend_gene = tile * p.g.genes + gl;
                }
                if (++sel == p.g.lpg) { sel = 0; finalize(); }
            });
            FN_STAMP(3);
            finalize();
            FN_STAMP(4);
        } else if constexpr (KIND == KIND_PATSEQ_SHARED) {
            // one sweep per gene, nothing written to the stage (eval_patseq_shared)
            ring.pass(smem_raw, bars, p.src, L, false, [&](unsigned char* st, long long tile) {
                const long long gene = tile * p.g.genes + gl;
                gv.y = (double*)st + (size_t)gl * m;
                gv.aux = (double*)(st + L.aux_off) + (size_t)gl * m;
                gv.gs = p.src.gs ? ((const double*)(st + L.gs_off))[gl] : 0.0;
                const uint8_t status = st[L.status_off + gl];
                const bool control = (status & 2) != 0;
                const double cst = p.src.cst ? ((const double*)(st + L.cst_off))[gl] : 0.0;
                EvalOut eo;
                eval_patseq_shared<C>(ec, gv, control ? xc : xn, control ? p.c_ctrl : p.c_noise, (status & 1) != 0, cst, eo);
                if (lane == 0) account(eo, gene);
            });
        } else {
            // PATSEQ keeps 1/sigma2 in the aux tile between its two sweeps: the stage is always written
            const int S = p.src.row_stride;                  // doubles between the rows of a stage
            const bool rows = KIND == KIND_SCALE_PS && fastps && p.g.rows != 0;
            ring.pass(smem_raw, bars, p.src, L, inplace || KIND == KIND_PATSEQ, [&](unsigned char* st, long long tile) {
                const long long gene = tile * p.g.genes + gl;
                double* ys = (double*)st;
                double* auxs = (double*)(st + L.aux_off);
                gv.y = ys + (size_t)gl * S;
                gv.aux = p.src.aux ? auxs + (size_t)gl * S : nullptr;
                gv.gs = p.src.gs ? ((const double*)(st + L.gs_off))[gl] : 0.0;
                const uint8_t status = st[L.status_off + gl];
                const bool control = (status & 2) != 0;
                const bool eligible = (status & 1) != 0;
                const double cst = p.src.cst ? ((const double*)(st + L.cst_off))[gl] : 0.0;
                const double* X = control ? xc : xn;
                const int c_real = control ? p.c_ctrl : p.c_noise;
                EvalOut eo;
                if (KIND == KIND_SCALE_PS && fastps) {
                    // complete genes: shared normal matrix (PsDesign); the others: the general two sweeps
                    const PsDesign<C>& pd = control ? pdc : pdn;
                    double beta[C];
                    const bool complete = rows ? ps_rows_beta<C>(pd, gv.y, m, eligible, beta)
                                                : ps_complete_beta<C>(pd, gv, eligible, beta);
                    eval_zero(eo);
                    if constexpr (!LEAN) {
                        if (warp_any(eligible && !complete))
                            eval_general<C, KIND>(ec, gv, X, c_real, eligible && !complete, cst, true, false, eo,
                                                  eligible && !complete);
                    } else if (eligible && !complete) {
                        eo.value = NAN; eo.used = 1;             // cannot happen (PREP_PARTIAL == 0): reported as bad
                    }
                    EvalOut ef;
                    double kappa;
                    if (rows) ps_rows_finish<C>(ec, pd, gv.y, m, control ? x2 + C : x2, c_real, complete, LEAN ? !complete : !eligible, beta, cst, ef, kappa);
                    else ps_complete_finish<C>(ec, pd, gv, X, c_real, complete, LEAN ? !complete : !eligible, beta, cst, ef, kappa);
                    if (complete) eo = ef;
                    // the column sum multiplies the gene's slots by this: -kappa for s_j = (w_j r_j)^2
                    if (lane == 0) kap[gl] = complete ? -kappa : 1.0;
                    {
                        // complete genes of this warp by design: one shared-memory atomic per warp, not per gene
                        const bool counted = lane == 0 && complete && kappa > 0.0 && gene < p.g.n;
                        const unsigned int mn = __ballot_sync(0xffffffffu, counted && !control);
                        const unsigned int mc = __ballot_sync(0xffffffffu, counted && control);
                        if ((threadIdx.x & 31) == 0) {
                            if (mn) atomicAdd(&ps_count[0], __popc(mn));
                            if (mc) atomicAdd(&ps_count[1], __popc(mc));
                        }
                    }
                } else if constexpr (!LEAN) {
                    eval_general<C, KIND>(ec, gv, X, c_real, eligible, cst, p.inplace_a != 0, p.inplace_b != 0, eo);
                }
                if (lane == 0) account(eo, gene);
                if (inplace && rows) {
                    // thread (sub2, column pair): one 16-byte load covers two columns of a gene's row
                    __syncthreads();
                    if (sub2 < nsub2) {
                        const double* at = ys + (size_t)sub2 * S + 2 * colp;
#pragma unroll 4
                        for (int g = sub2; g < p.g.genes; g += nsub2, at += (size_t)nsub2 * S) {
                            double s0, s1;
                            load_pair(at, s0, s1);
                            const double k = kap[g];
                            col_a = fma(k, s0, col_a);
                            col_b = fma(k, s1, col_b);          // (col_b doubles as the odd column's sum here)
                        }
                    }
                } else if (inplace) {
                    __syncthreads();
                    if (sub < nsub) {
                        for (int g = sub; g < p.g.genes; g += nsub) {
                            if (p.inplace_a) col_a += (KIND == KIND_SCALE_PS && fastps) ? kap[g] * ys[(size_t)g * S + col] : ys[(size_t)g * S + col];
                            if (p.inplace_b) col_b += auxs[(size_t)g * S + col];
                        }
                    }
                }
            });
        }

        block_sum<ACC_FIXED>(acc, vals, red_scratch);
        FN_STAMP(5);
        if (inplace && KIND == KIND_SCALE_PS && fastps && p.g.rows != 0) {
            // rows layout: colscr[sub2][m], thread (sub2, pair) holds columns 2 pair and 2 pair + 1
            if (sub2 < nsub2) { colscr[sub2 * m + 2 * colp] = col_a; colscr[sub2 * m + 2 * colp + 1] = col_b; }
            __syncthreads();
            if (threadIdx.x < m) {
                double sa = 0.0;
                for (int q = 0; q < nsub2; ++q) sa += colscr[q * m + threadIdx.x];
                sa += ps_count[0] * cjn[threadIdx.x] + ps_count[1] * cjc[threadIdx.x];
                vals[ACC_FIXED + threadIdx.x] = sa;
                vals[ACC_FIXED + m + threadIdx.x] = 0.0;
            }
            __syncthreads();
        } else if (inplace) {
            colscr[threadIdx.x] = col_a;
            colscr[blockDim.x + threadIdx.x] = col_b;
            __syncthreads();
            if (threadIdx.x < m) {
                double sa = 0.0, sb = 0.0;
                for (int q = 0; q < nsub; ++q) { sa += colscr[q * m + threadIdx.x]; sb += colscr[blockDim.x + q * m + threadIdx.x]; }
                // the gene-independent part of the complete genes' gradient, once per gene of this CTA
                if (fastps) sa += ps_count[0] * cjn[threadIdx.x] + ps_count[1] * cjc[threadIdx.x];
                vals[ACC_FIXED + threadIdx.x] = sa;
                vals[ACC_FIXED + m + threadIdx.x] = sb;
            }
            __syncthreads();
        }
        grid_reduce(p.gr, sv, vals);
        if (!resident) break;
        __syncthreads();             // vals / scratch are reused by the next evaluation
        ++expect;
    }
    ring.drain(bars);
}

// ------------------------------------------------------------------------------------------
// K2f -log likelihood + gradient of the factor models (covariance diag(sigma2) + F F')
// ------------------------------------------------------------------------------------------
struct EvalFactorsParams {
    GeomParams g;
    TileSrc src;             // y, aux (precision weights), cst (per-gene mode only), status
    const double* zn; const double* zc;     // [design | factors | 0], m x C, for ordinary / control genes
    int c_noise, c_ctrl, nf;
    int is_t;
    double v, theta0;
    double* pg_score; double* pg_d2; int* pg_p;
    GridReduce gr;           // nacc = ACC_FIXED + m * nf  (d/dF follows the fixed accumulators)
    StepVars sv;
};

template <int C>
__global__ void __launch_bounds__(256) eval_factors_kernel(const EvalFactorsParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[FN_MAX_STAGES];
    __shared__ double red_scratch[32 * ACC_FIXED];
    const int m = p.g.m, nf = p.nf, ncol = m * nf;
    const TileLayout L = tile_layout(p.src);
    tile_prologue(smem_raw, bars, p.g.stages, p.src, L, p.g.n_tiles);
    double* zn = (double*)(smem_raw + (size_t)p.g.stages * L.stage_bytes);
    constexpr int XS = XStride<C>::V;
    double* zc = zn + m * XS;
    double* t0 = zc + m * XS;
    double* t1 = t0 + (m + 1);
    double* vals = t1 + (m + 1);
    double* colscr = vals + (ACC_FIXED + ncol);
    double* gscr = colscr + blockDim.x;              // genes x m x nf: d/dF terms of the current tile
    stage_design<C>(zn, p.zn, m);
    stage_design<C>(zc, p.zc, m);
    const bool per_gene = p.pg_score != nullptr;
    if (per_gene) fill_density_tables(p.is_t, p.v, m, t0, t1);
    else __syncthreads();

    EvalConst ec;
    ec.kind = KIND_SCALE1; ec.tail_own = 0; ec.v3 = 0; ec.is_t = p.is_t; ec.v = p.v; ec.theta0 = p.theta0;
    ec.ta = nullptr; ec.tb = nullptr; ec.t0 = per_gene ? t0 : nullptr; ec.t1 = per_gene ? t1 : nullptr; ec.finish_setup();

    const int gl = threadIdx.x / p.g.lpg, lane = threadIdx.x % p.g.lpg;
    GeneView gv;
    gv.setup(m, lane, p.g.lpg, gl);
    gv.gs = 0.0;
    double acc[ACC_FIXED];
#pragma unroll
    for (int i = 0; i < ACC_FIXED; ++i) acc[i] = 0.0;
    const int nsub = blockDim.x / ncol;
    const int col = threadIdx.x % ncol, sub = threadIdx.x / ncol;
    double col_f = 0.0;

    tile_loop(smem_raw, bars, p.g.stages, p.src, L, p.g.n_tiles, false, [&](unsigned char* st, long long tile) {
        const long long gene = tile * p.g.genes + gl;
        gv.y = (double*)st + (size_t)gl * m;
        gv.aux = p.src.aux ? (double*)(st + L.aux_off) + (size_t)gl * m : nullptr;
        const uint8_t status = st[L.status_off + gl];
        const bool control = (status & 2) != 0;
        const double cst = p.src.cst ? ((const double*)(st + L.cst_off))[gl] : 0.0;
        EvalOut eo;
        eval_factors<C>(ec, gv, control ? zc : zn, control ? p.c_ctrl : p.c_noise, nf, (status & 1) != 0, cst,
                        gscr + (size_t)gl * ncol, eo);
        if (lane == 0 && gene < p.g.n) {
            if (eo.used) {
                if (eo.value == eo.value) {
                    acc[ACC_VALUE] += eo.value; acc[ACC_DV] += eo.dv; acc[ACC_GA] += eo.ga; acc[ACC_USED] += 1.0;
                    acc[ACC_PSUM] += (double)eo.p;
                } else {
                    acc[ACC_BAD] += 1.0;
                }
            }
            if (p.pg_score) {
                p.pg_score[gene] = eo.used ? eo.value : NAN;
                p.pg_d2[gene] = eo.d2;
                p.pg_p[gene] = eo.p;
            }
        }
        __syncthreads();
        if (sub < nsub)
            for (int g = sub; g < p.g.genes; g += nsub) col_f += gscr[(size_t)g * ncol + col];
    });

    block_sum<ACC_FIXED>(acc, vals, red_scratch);
    colscr[threadIdx.x] = col_f;
    __syncthreads();
    if (threadIdx.x < ncol) {
        double s = 0.0;
        for (int q = 0; q < nsub; ++q) s += colscr[q * ncol + threadIdx.x];
        vals[ACC_FIXED + threadIdx.x] = s;
    }
    __syncthreads();
    grid_reduce(p.gr, p.sv, vals);
}

#ifndef FN_WIDTH_ONLY
// noise p-values from (d2, p): Mvt.p_value F sf (distributions.py:150-155) / Mvnormal chi2 sf (:52-57)
__global__ void noise_p_kernel(long long n, const double* d2, const int* pdf, int is_t, double v, double* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d = d2[i];
    const int p = pdf[i];
    if (!(d == d) || p <= 0) { out[i] = NAN; return; }
    out[i] = is_t ? f_sf(d / p, (double)p, v) : chi2_sf(d, (double)p);
}
#endif

// ------------------------------------------------------------------------------------------
// K3 coefficients
// ------------------------------------------------------------------------------------------
struct CoefParams {
    GeomParams g;
    TileSrc src;
    const double* x;         // m x width, whitened coefficient design
    const double* rinv;      // c x c upper triangular
    int c;
    int nf;                  // factor columns following the design in x
    int kind, tail_own, v3, is_t;
    double v, theta0;
    const double* ta; const double* tb;
    double* coef;            // n x c
    double* covar;           // n x c x c
    double* dfpost;          // n
};

template <int C>
__global__ void __launch_bounds__(256) coef_kernel(const CoefParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[FN_MAX_STAGES];
    __shared__ double rinv[FN_MAXC * FN_MAXC];
    const int m = p.g.m, c = p.c;
    const TileLayout L = tile_layout(p.src);
    tile_prologue(smem_raw, bars, p.g.stages, p.src, L, p.g.n_tiles);
    double* x = (double*)(smem_raw + (size_t)p.g.stages * L.stage_bytes);
    double* ta = x + m * XStride<C>::V;
    double* tb = ta + m;
    stage_design<C>(x, p.x, m);
    for (int i = threadIdx.x; i < m; i += blockDim.x) { ta[i] = p.ta ? p.ta[i] : 0.0; tb[i] = p.tb ? p.tb[i] : 0.0; }
    for (int i = threadIdx.x; i < c * c; i += blockDim.x) rinv[i] = p.rinv[i];
    __syncthreads();

    EvalConst ec;
    ec.kind = p.kind; ec.tail_own = p.tail_own; ec.v3 = p.v3; ec.is_t = p.is_t; ec.v = p.v; ec.theta0 = p.theta0;
    ec.ta = ta; ec.tb = tb; ec.t0 = nullptr; ec.t1 = nullptr; ec.finish_setup();
    const int gl = threadIdx.x / p.g.lpg, lane = threadIdx.x % p.g.lpg;
    GeneView gv;
    gv.setup(m, lane, p.g.lpg, gl);

    tile_loop(smem_raw, bars, p.g.stages, p.src, L, p.g.n_tiles, false, [&](unsigned char* st, long long tile) {
        const long long gene = tile * p.g.genes + gl;
        gv.y = (double*)st + (size_t)gl * m;
        gv.aux = p.src.aux ? (double*)(st + L.aux_off) + (size_t)gl * m : nullptr;
        gv.gs = p.src.gs ? ((const double*)(st + L.gs_off))[gl] : 0.0;
        double beta[C], ainv[Tri<C>::N];
        CoefOut co;
        coef_gene<C>(ec, gv, x, c, p.nf, beta, ainv, co);
        if (lane != 0 || gene >= p.g.n) return;
        double* cf = p.coef + gene * c;
        double* cv = p.covar + gene * c * c;
        if (!co.ok) {
            for (int i = 0; i < c; ++i) cf[i] = NAN;
            for (int i = 0; i < c * c; ++i) cv[i] = NAN;
            p.dfpost[gene] = NAN;
            return;
        }
        // posterior scale (Mvt.conditional, distributions.py:210-214): (df + d2) / (df + p2)
        const double s2 = p.is_t ? (p.v + co.d2) / (p.v + co.p2) : 1.0;
        p.dfpost[gene] = p.is_t ? p.v + co.p2 : NAN;
        // back to the user's basis: coef = Rinv beta,  V = Rinv Ainv Rinv' s2
        // (wide designs keep the outer loops rolled: C^3 terms unrolled cost minutes of compile time)
        double tmp[C * C];
#pragma unroll (C > 8 ? 1 : 64)
        for (int i = 0; i < C; ++i) {
            if (i < c) {
                double a = 0.0;
#pragma unroll
                for (int t = 0; t < C; ++t) if (t >= i && t < c) a += rinv[i * c + t] * beta[t];
                cf[i] = a;
            }
#pragma unroll (C > 8 ? 1 : 64)
            for (int j = 0; j < C; ++j) {
                double a = 0.0;
#pragma unroll
                for (int t = 0; t < C; ++t)
                    if (t >= i && t < c && i < c && j < c) a += rinv[i * c + t] * ainv[t > j ? FN_TRI(t, j) : FN_TRI(j, t)];
                tmp[i * C + j] = a;
            }
        }
#pragma unroll (C > 8 ? 1 : 64)
        for (int i = 0; i < C; ++i)
#pragma unroll (C > 8 ? 1 : 64)
            for (int j = 0; j < C; ++j) {
                if (i < c && j < c) {
                    double a = 0.0;
#pragma unroll
                    for (int t = 0; t < C; ++t) if (t >= j && t < c) a += tmp[i * C + t] * rinv[j * c + t];
                    cv[i * c + j] = a * s2;
                }
            }
    });
}

// ------------------------------------------------------------------------------------------
// K3b contrasts + test: thread per gene over the compact per-gene results
// ------------------------------------------------------------------------------------------
struct TestParams {
    long long n;
    int c, kc, is_t;
    const double* coef; const double* covar; const double* dfpost;
    const double* contrasts;     // c x kc (device)
    double* contrasted;          // n x kc
    double* ccov;                // n x kc x kc or nullptr
    double* pval;                // n
};

#ifndef FN_WIDTH_ONLY
// KC = compile-time number of contrasts (1..4: everything in registers), 0 = any (local arrays)
template <int KC>
__global__ void test_kernel(const TestParams p) {
    __shared__ double cm[FN_MAXC * FN_MAXC];
    for (int i = threadIdx.x; i < p.c * p.kc; i += blockDim.x) cm[i] = p.contrasts[i];
    __syncthreads();
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.n) return;
    const double* V = p.covar + g * p.c * p.c;
    double* out = p.contrasted + g * p.kc;
    if (V[0] != V[0]) {                      // coef_dists[row] is None (fitnoise.py:385)
        for (int a = 0; a < p.kc; ++a) out[a] = NAN;
        if (p.ccov) for (int a = 0; a < p.kc * p.kc; ++a) p.ccov[g * p.kc * p.kc + a] = NAN;
        p.pval[g] = NAN;
        return;
    }
    double pv;
    const double* cf = p.coef + g * p.c;
    double* cc = p.ccov ? p.ccov + g * p.kc * p.kc : nullptr;
    if constexpr (KC > 0) test_gene_k<KC>(p.c, cf, V, p.dfpost[g], p.is_t, cm, out, cc, pv);
    else test_gene(p.c, p.kc, cf, V, p.dfpost[g], p.is_t, cm, out, cc, pv);
    p.pval[g] = pv;
}
#endif

// ------------------------------------------------------------------------------------------
// get_weights: 1/sigma2 per observation (fitnoise.py:399-404); inf/NaN exactly where the reference's are
// ------------------------------------------------------------------------------------------
struct WeightParams {
    long long n; int m;
    const double* y; const double* aux; const double* gs;
    int kind, tail_own, v3;
    double theta0;
    const double* ta; const double* tb;
    const double* fsq;       // factor models: sum_i F[j,i]^2 per sample (the diagonal of F F'), or nullptr
    double* out;
};

#ifndef FN_WIDTH_ONLY
__global__ void weights_kernel(const WeightParams p) {
    const long long total = p.n * p.m;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long g = i / p.m;
        const int j = (int)(i - g * p.m);
        double w;
        if (p.kind == KIND_PATSEQ) {
            const double rc = p.aux[i];
            const double gsv = p.gs ? p.gs[g] : 0.0;
            const double y = p.y[i];
            const double U = p.v3 ? rc * gsv : rc;
            const double T = p.tail_own ? (finite_d(y) ? y * y : 0.0) : gsv;
            w = 1.0 / (p.ta[j] * U + p.tb[j] * T);
        } else {
            const double wp = p.aux ? p.aux[i] : 1.0;
            if (p.fsq) w = 1.0 / (p.theta0 / wp + p.fsq[j]);
            else w = (p.kind == KIND_SCALE_PS) ? wp * p.ta[j] : wp / p.theta0;
        }
        p.out[i] = w;
    }
}
#endif

}  // namespace fn
