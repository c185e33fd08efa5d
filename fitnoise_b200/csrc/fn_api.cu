// C-ABI layer of libfitnoise_b200.so (declarations and reference citations: include/fitnoise_b200.h).
// Owns device memory, picks the tile geometry, launches the kernels of fn_kernels.cuh / fn_bh.cuh.
// There is no CPU path here: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <utility>
#include <string>
#include <thread>
#include <tuple>
#include <vector>
#include <nvtx3/nvToolsExt.h>      // header-only; a no-op unless a profiler attaches

#include "../../include/fitnoise_b200.h"
#include "fn_bh.cuh"
#include "fn_host.hpp"
#include "fn_kernels.cuh"
#include "fn_transform.cuh"

using namespace fn;

// per-width kernel tables (fn_width_inst.cuh)
#ifdef FN_SINGLE_TU
#define FN_W 1
#include "fn_width_inst.cuh"
#undef FN_W
#define FN_W 2
#include "fn_width_inst.cuh"
#undef FN_W
#define FN_W 3
#include "fn_width_inst.cuh"
#undef FN_W
#define FN_W 4
#include "fn_width_inst.cuh"
#undef FN_W
#define FN_W 6
#include "fn_width_inst.cuh"
#undef FN_W
#define FN_W 8
#include "fn_width_inst.cuh"
#undef FN_W
#define FN_W 12
#include "fn_width_inst.cuh"
#undef FN_W
#define FN_W 16
#include "fn_width_inst.cuh"
#undef FN_W
#else
namespace fn {
const KernelSet* kernel_set_w1(); const KernelSet* kernel_set_w2(); const KernelSet* kernel_set_w3();
const KernelSet* kernel_set_w4(); const KernelSet* kernel_set_w6(); const KernelSet* kernel_set_w8();
const KernelSet* kernel_set_w12(); const KernelSet* kernel_set_w16();
}
#endif

static const KernelSet* kernel_set(int width) {
    switch (width) {
        case 1: return kernel_set_w1();
        case 2: return kernel_set_w2();
        case 3: return kernel_set_w3();
        case 4: return kernel_set_w4();
        case 6: return kernel_set_w6();
        case 8: return kernel_set_w8();
        case 12: return kernel_set_w12();
        case 16: return kernel_set_w16();
        default: return nullptr;
    }
}

// NVTX range over one C-ABI call (K1 = fn_upload, K2 = fn_nll_grad / fn_noise_stats, K3 = fn_coef,
// K3b / K4 = fn_test*, fn_bh*): shows up as a named span in nsys / ncu timelines, costs a few ns otherwise.
// The per-evaluation call is only marked when FN_NVTX=1 (it runs every 10-25 us).
struct NvtxRange {
    bool on;
    explicit NvtxRange(const char* name, bool enabled = true) : on(enabled) { if (on) nvtxRangePushA(name); }
    ~NvtxRange() { if (on) nvtxRangePop(); }
};
static const bool g_nvtx_hot = getenv("FN_NVTX") != nullptr;

static thread_local std::string g_error;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define FN_ERR_CUDA -1
#define FN_ERR_ARG -2
#define FN_ERR_STATE -3
#define FN_ERR_UNSUPPORTED -4

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(FN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

static_assert(sizeof(fn_family) == sizeof(Family), "fn_family and fn::Family must match");

struct Geom {
    int threads, lpg, genes, stages, grid;
    size_t smem;
    long long n_tiles;
};

// grow-only device allocation: re-uploading data of the same shape costs no cudaMalloc/cudaFree
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct fn_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    size_t smem_optin = 0;      // largest dynamic + static shared memory of one CTA
    size_t smem_sm = 0;         // shared memory of one SM

    int64_t n = 0, n_pad = 0;
    int m = 0;
    bool has_aux = false, loaded = false, has_controls = false;
    Family fam{};
    Design dn, dc;
    int width = 0;
    double cst_sum = 0.0;

    // active arrays (point into the buffers below, or nullptr when the model does not use them)
    double *y = nullptr, *aux = nullptr, *gs = nullptr, *cst = nullptr, *avg = nullptr;
    uint8_t* status = nullptr;
    double *xn = nullptr, *xc = nullptr;
    DevBuf b_y, b_aux, b_gs, b_cst, b_avg, b_status, b_xn, b_xc, b_tab, b_pg_score, b_pg_d2, b_pg_p, b_noise_p,
        b_coef, b_covar, b_dfpost, b_xcoef, b_rinv;
    size_t tab_host_cap = 0;
    // factor models: current factor vectors (m x nf) and the augmented design tables built from them
    std::vector<double> factors;
    DevBuf b_zn, b_zc, b_fsq;
    double *zn = nullptr, *zc = nullptr, *fsq = nullptr;
    // pinned staging for uploads from pageable host memory
    static const int kStageWorkers = 8;
    void* stage_buf[2 * kStageWorkers] = {nullptr};
    cudaEvent_t stage_ev[2 * kStageWorkers] = {nullptr};
    size_t stage_bytes = 0;

    double* partials = nullptr;
    size_t partials_cap = 0;
    Slot* pslots = nullptr;            // per-CTA partial sums as authenticated slots (fn_kernels.cuh grid_reduce)
    size_t pslots_cap = 0;
    unsigned int* ticket = nullptr;
    double* out_dev = nullptr;
    Slot* out_host = nullptr;          // pinned, mapped: {value, seq} per accumulator
    Slot* out_host_devptr = nullptr;
    std::vector<double> out_vals;      // the values of the last completed launch
    size_t out_cap = 0;
    double* tab_dev = nullptr;         // 2m
    double* tab_host = nullptr;        // pinned 2m

    double *pg_score = nullptr, *pg_d2 = nullptr, *noise_p = nullptr;
    int* pg_p = nullptr;

    int coef_c = 0;
    bool coef_is_t = false;
    // device-resident results of the last fn_test / fn_test_bh (inside scratch_dev)
    double *test_out = nullptr, *test_p = nullptr, *test_q = nullptr;
    uint32_t* test_order = nullptr;
    int test_kc = 0;
    double *coef = nullptr, *covar = nullptr, *dfpost = nullptr, *x_coef = nullptr, *rinv_dev = nullptr;
    double* scratch_dev = nullptr;     // misc device scratch
    size_t scratch_cap = 0;

    int tune_lpg = 0, tune_threads = 0, tune_stages = 0, tune_ctas = 0;
    int64_t launches = 0;
    bool timing = false, timing_pending = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_ms = 0.0;

    // Benjamini-Hochberg as a captured CUDA graph (see bh_run)
    cudaGraphExec_t bh_exec = nullptr;
    int64_t bh_n = 0;
    const void* bh_p = nullptr;
    void* bh_q = nullptr;
    void* bh_work = nullptr;
    uint32_t* bh_order = nullptr;

    // fitnoise.transform (fn_transform_*): the count matrix and the design basis
    DevBuf b_tx, b_tq, b_ta, b_tmean;
    int64_t tr_n = 0;
    int tr_m = 0, tr_c = 0;

    unsigned long long host_seq = 0;   // completion-flag sequence of the pinned result buffer
    // histogram of residual df over the genes in the REML sum (prepare kernel): the (v, p)-only terms of
    // the objective are added on the host from it (table_terms)
    DevBuf b_phist, b_bhwork;
    std::vector<long long> p_hist;
    double psum_expected = 0.0;
    bool all_complete = false;         // no weights and every gene of the REML sum has all m samples: lean kernels
    // resident evaluation kernel (fn_resident_begin / _end): command block and state
    bool resident = false;
    int res_words = 0;
    unsigned long long* cmd_host = nullptr;        // pinned, mapped
    unsigned long long* cmd_host_devptr = nullptr;
    unsigned long long* exit_host = nullptr;       // pinned, mapped (one word)
    unsigned long long* exit_host_devptr = nullptr;
    Slot* cmd_dev = nullptr;                       // device copy of the command: one authenticated slot per word
    size_t cmd_cap = 0;                            // words the three buffers hold
    int64_t resident_launches = 0, resident_evals = 0;
    double* flush_buf = nullptr;                   // fn_l2_flush: 2 x L2 size
    size_t flush_words = 0;
    // fused multi-GPU exchange (fn_comm_*): mailboxes mapped with CUDA IPC
    int comm_world = 0, comm_rank = 0, comm_slot = 0;
    unsigned long long comm_seq = 0;
    Slot* comm_own = nullptr;
    Slot* comm_peer[FN_MAX_RANKS] = {nullptr};
};

static int leave_resident(fn_handle* h);     // stop the resident evaluation kernel, if one is running

#define LEAVE_RESIDENT(h)                          \
    do {                                           \
        int rc_res__ = leave_resident(h);          \
        if (rc_res__) return rc_res__;             \
    } while (0)

template <class T>
static void dfree(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

static cudaError_t ensure_buf(DevBuf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return cudaSuccess;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    cudaError_t e = cudaMalloc(&b.p, bytes ? bytes : 16);
    if (e == cudaSuccess) b.cap = bytes ? bytes : 16;
    return e;
}

// forget the uploaded data (buffers stay allocated for the next upload)
static void release_data(fn_handle* h) {
    h->y = h->aux = h->gs = h->cst = h->avg = nullptr;
    h->status = nullptr;
    h->xn = h->xc = nullptr;
    h->tab_dev = nullptr;
    h->pg_score = h->pg_d2 = h->noise_p = nullptr;
    h->pg_p = nullptr;
    h->coef = h->covar = h->dfpost = h->x_coef = h->rinv_dev = nullptr;
    h->test_out = h->test_p = h->test_q = nullptr; h->test_order = nullptr; h->test_kc = 0;
    h->zn = h->zc = h->fsq = nullptr;
    h->factors.clear();
    h->loaded = false;
    h->coef_c = 0;
}

static void free_buffers(fn_handle* h) {
    release_data(h);
    DevBuf* all[] = {&h->b_phist, &h->b_bhwork, &h->b_y, &h->b_aux, &h->b_gs, &h->b_cst, &h->b_avg, &h->b_status, &h->b_xn, &h->b_xc, &h->b_tab,
                     &h->b_pg_score, &h->b_pg_d2, &h->b_pg_p, &h->b_noise_p, &h->b_coef, &h->b_covar, &h->b_dfpost,
                     &h->b_xcoef, &h->b_rinv, &h->b_zn, &h->b_zc, &h->b_fsq, &h->b_tx, &h->b_tq, &h->b_ta, &h->b_tmean};
    for (DevBuf* b : all) { if (b->p) cudaFree(b->p); b->p = nullptr; b->cap = 0; }
    if (h->tab_host) { cudaFreeHost(h->tab_host); h->tab_host = nullptr; h->tab_host_cap = 0; }
    for (int i = 0; i < 2 * fn_handle::kStageWorkers; ++i) {
        if (h->stage_buf[i]) { cudaFreeHost(h->stage_buf[i]); h->stage_buf[i] = nullptr; }
        if (h->stage_ev[i]) { cudaEventDestroy(h->stage_ev[i]); h->stage_ev[i] = nullptr; }
    }
    h->stage_bytes = 0;
}

extern "C" const char* fn_last_error(void) { return g_error.c_str(); }
extern "C" const char* fn_version(void) { return "fitnoise_b200 0.2 (sm_100a)"; }
#ifndef FN_SOURCE_HASH
#define FN_SOURCE_HASH "unknown"
#endif
// hash of the sources this library was built from (fitnoise_b200/build.py); the Python binding
// compares it with the sources it sees and refuses a stale binary
extern "C" const char* fn_source_hash(void) { return FN_SOURCE_HASH; }

extern "C" int fn_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int fn_create(int device, void* stream, fn_handle** out) {
    if (!out) return fail(FN_ERR_ARG, "fn_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(FN_ERR_CUDA, "fn_create: no CUDA device (%s); fitnoise_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(FN_ERR_ARG, "fn_create: device %d out of range (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    fn_handle* h = new fn_handle();
    h->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    h->smem_sm = prop.sharedMemPerMultiprocessor;
    if (stream) {
        h->stream = (cudaStream_t)stream;
    } else {
        CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    CUDA_TRY(cudaMalloc(&h->ticket, sizeof(unsigned int)));
    CUDA_TRY(cudaMemset(h->ticket, 0, sizeof(unsigned int)));
    CUDA_TRY(cudaEventCreate(&h->ev0));
    CUDA_TRY(cudaEventCreate(&h->ev1));
    const char* env;
    if ((env = getenv("FN_LPG"))) h->tune_lpg = atoi(env);
    if ((env = getenv("FN_THREADS"))) h->tune_threads = atoi(env);
    if ((env = getenv("FN_STAGES"))) h->tune_stages = atoi(env);
    if ((env = getenv("FN_CTAS"))) h->tune_ctas = atoi(env);
    *out = h;
    return 0;
}

extern "C" int fn_comm_disconnect(fn_handle* h);

extern "C" int fn_destroy(fn_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    leave_resident(h);
    cudaStreamSynchronize(h->stream);
    fn_comm_disconnect(h);
    if (h->cmd_host) cudaFreeHost(h->cmd_host);
    if (h->exit_host) cudaFreeHost(h->exit_host);
    dfree(h->cmd_dev);
    dfree(h->flush_buf);
    if (h->bh_exec) { cudaGraphExecDestroy(h->bh_exec); h->bh_exec = nullptr; }
    free_buffers(h);
    dfree(h->partials); dfree(h->pslots); dfree(h->ticket); dfree(h->out_dev); dfree(h->scratch_dev);
    if (h->out_host) cudaFreeHost(h->out_host);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

extern "C" int fn_synchronize(fn_handle* h) {
    if (!h) return fail(FN_ERR_ARG, "NULL handle");
    LEAVE_RESIDENT(h);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int fn_set_tuning(fn_handle* h, int32_t lpg, int32_t threads, int32_t stages, int32_t ctas) {
    if (!h) return fail(FN_ERR_ARG, "NULL handle");
    LEAVE_RESIDENT(h);
    if (lpg && (lpg & (lpg - 1))) return fail(FN_ERR_ARG, "lpg must be a power of two");
    if (threads && ((threads & (threads - 1)) || threads < 32 || threads > 256))
        return fail(FN_ERR_ARG, "threads must be 32, 64, 128 or 256 (genes per tile must divide the %d-row padding)", FN_ROW_PAD);
    h->tune_lpg = lpg; h->tune_threads = threads; h->tune_stages = stages; h->tune_ctas = ctas;
    return 0;
}
extern "C" int64_t fn_launch_count(fn_handle* h) { return h ? h->launches : 0; }
extern "C" int fn_set_timing(fn_handle* h, int32_t on) { if (!h) return FN_ERR_ARG; h->timing = on != 0; return 0; }
extern "C" double fn_last_kernel_ms(fn_handle* h) {
    if (!h) return 0.0;
    if (h->timing_pending) {                    // events of the last timed evaluation kernel
        float ms = 0;
        if (cudaEventSynchronize(h->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess)
            h->last_ms = ms;
        h->timing_pending = false;
    }
    return h->last_ms;
}

// ------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------
// `one_big_stage`: the instruction-bound sweeps (weighted Model_normal/_t, unweighted per-sample
// models) run 10-18 % faster with half the lanes per gene and ONE ~100 KB stage per CTA than with two
// 50 KB stages: fewer lanes repeat the per-gene tail (reduction, factorisation, logarithm), and the
// second CTA of the SM covers the load (measured: profiles/r01_sweep_weighted.log).
// `rows_stride` > 0: the thread-per-gene layout on padded rows (fn_tile.cuh TileSrc::row_stride): one lane
// per gene, `rows_stride` doubles per row in the stage; 256 threads = 256 genes per tile and one CTA per SM
// measured best (tools/ps_rows_sweep.py).
static int choose_geom(fn_handle* h, bool aux, bool gs, bool cst, size_t table_bytes, Geom& g, int min_threads = 0,
                       size_t per_gene_extra = 0, bool one_big_stage = false, int rows_stride = 0) {
    const int m = h->m;
    int ctas = h->tune_ctas > 0 ? h->tune_ctas : (rows_stride > 0 ? 1 : 2);
    int threads = h->tune_threads > 0 ? h->tune_threads : 256;
    // small problems: more, smaller CTAs so that every SM gets work
    if (h->tune_threads == 0 && h->n < (int64_t)h->sm_count * 256) threads = 128;
    if (threads < min_threads) threads = 256;      // per-sample column sums need one thread per sample
    // The kernels hold up to 4.7 KB of static shared memory (reduction scratch, mbarriers) and the
    // system reserves 1 KB per CTA: leave room for both, or fewer CTAs than planned fit on an SM
    // (and a CTA at the very limit fails to launch).
    const size_t static_reserve = 6 * 1024;
    const size_t budget = h->smem_sm / ctas - static_reserve;   // dynamic bytes per CTA
    auto lanes_for = [&](size_t stage_target) {
        int l = 1;
        while (l < 32 && threads / (l * 2) >= 16 &&
               tile_stage_bytes(threads / l, m, aux, gs, cst) + (size_t)(threads / l) * per_gene_extra > stage_target) l *= 2;
        return l;
    };
    int lpg = rows_stride > 0 ? 1 : h->tune_lpg > 0 ? h->tune_lpg : lanes_for(50 * 1024);
    if (rows_stride == 0 && h->tune_lpg == 0 && one_big_stage && threads == 256 && budget > table_bytes + 256 + 64 * 1024) {
        // only when every CTA still gets enough tiles to balance the grid (large n)
        const int big = lanes_for(budget - table_bytes - 256);
        const long long tiles = (h->n + threads / big - 1) / (threads / big);
        if (tiles >= 8LL * h->sm_count * ctas) lpg = big;
    }
    // genes per tile must divide FN_ROW_PAD (the per-gene arrays are padded to that many rows, so a
    // full-size bulk copy of the last tile stays inside the allocations): a power of two >= 16
    if (lpg > 32 || threads % lpg || (threads / lpg) < 16 || ((threads / lpg) & (threads / lpg - 1)) || FN_ROW_PAD % (threads / lpg))
        return fail(FN_ERR_ARG, "invalid geometry: threads %d lpg %d (genes per tile must be a power of two, 16..%d)", threads, lpg, FN_ROW_PAD);
    const int genes = threads / lpg;
    const size_t stage = tile_stage_bytes(genes, m, aux, gs, cst, rows_stride);
    const size_t fixed = table_bytes + 256 + (size_t)genes * per_gene_extra;
    if (fixed + stage > h->smem_optin - static_reserve)
        return fail(FN_ERR_UNSUPPORTED, "m = %d is too large for one tile stage in shared memory (%zu bytes)", m, fixed + stage);
    int stages = h->tune_stages > 0 ? h->tune_stages : 3;      // as many as fit, see below
    if (stages > FN_MAX_STAGES) stages = FN_MAX_STAGES;
    while (stages > 1 && fixed + (size_t)stages * stage > budget) --stages;
    if (fixed + (size_t)stages * stage > budget) ctas = 1;      // a single CTA per SM then
    g.threads = threads; g.lpg = lpg; g.genes = genes; g.stages = stages;
    g.smem = fixed + (size_t)stages * stage;
    g.n_tiles = (h->n + genes - 1) / genes;
    long long grid = (long long)h->sm_count * ctas;
    if (grid > g.n_tiles) grid = g.n_tiles;
    if (grid < 1) grid = 1;
    g.grid = (int)grid;
    return 0;
}

static GeomParams geom_params(const fn_handle* h, const Geom& g) {
    GeomParams gp;
    gp.n = h->n; gp.n_tiles = g.n_tiles; gp.m = h->m; gp.genes = g.genes; gp.lpg = g.lpg;
    gp.stages = g.stages; gp.width = h->width; gp.rows = 0;
    return gp;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the function in the device's
// context, shared by every handle: remember the largest value set per (device, kernel) and only
// ever raise it.
static std::mutex g_smem_mutex;
static std::map<std::pair<int, const void*>, size_t> g_smem_set;

template <class K>
static int set_smem(fn_handle* h, K kernel, size_t bytes) {
    std::lock_guard<std::mutex> lock(g_smem_mutex);
    const std::pair<int, const void*> key(h->device, (const void*)kernel);
    auto it = g_smem_set.find(key);
    if (it != g_smem_set.end() && it->second >= bytes) return 0;
    CUDA_TRY(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    g_smem_set[key] = bytes;
    return 0;
}

// Every per-gene kernel is persistent (CTA b takes tiles b, b + grid, ...), and the ones that reduce
// over the grid let CTA 0 wait for the partial sums of all the others (fn_kernels.cuh grid_reduce):
// the grid must never exceed what the GPU holds at once, or the waiting CTA blocks the slot a queued
// CTA needs (measured: 40 % slower at c = 6, where registers allow one CTA per SM, not the two
// choose_geom plans for).  Occupancy per (device, kernel, threads, shared memory) is asked once.
static std::map<std::tuple<int, const void*, int, size_t>, int> g_occupancy;

static int bound_grid(fn_handle* h, const void* kernel, Geom& g) {
    if (!kernel) return fail(FN_ERR_UNSUPPORTED, "no kernel for this design width");
    int rc = set_smem(h, kernel, g.smem);
    if (rc) return rc;
    int per_sm = 0;
    {
        std::lock_guard<std::mutex> lock(g_smem_mutex);
        const auto key = std::make_tuple(h->device, kernel, g.threads, g.smem);
        auto it = g_occupancy.find(key);
        if (it != g_occupancy.end()) per_sm = it->second;
        else {
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, g.threads, g.smem));
            g_occupancy[key] = per_sm;
        }
    }
    if (per_sm < 1) return fail(FN_ERR_UNSUPPORTED, "the kernel does not fit on an SM (%d threads, %zu bytes of shared memory)", g.threads, g.smem);
    const long long cap = (long long)per_sm * h->sm_count;
    if (g.grid > cap) g.grid = (int)cap;
    return 0;
}

// launch one of the per-width kernels (all take their parameter block by value as the only argument)
static int launch_width_kernel(fn_handle* h, const void* kernel, int grid, int threads, size_t smem, void* params) {
    if (!kernel) return fail(FN_ERR_UNSUPPORTED, "no kernel for this design width");
    int rc = set_smem(h, kernel, smem);
    if (rc) return rc;
    void* args[] = {params};
    CUDA_TRY(cudaLaunchKernel(kernel, dim3(grid), dim3(threads), args, smem, h->stream));
    return 0;
}

static int ensure_reduce(fn_handle* h, int grid, int nacc) {
    const size_t need = (size_t)grid * nacc;
    if (need > h->partials_cap) {
        dfree(h->partials);
        CUDA_TRY(cudaMalloc(&h->partials, need * sizeof(double)));
        h->partials_cap = need;
    }
    if ((size_t)grid * FN_SLOT_ACC > h->pslots_cap) {
        dfree(h->pslots);
        const size_t cap = (size_t)grid * FN_SLOT_ACC + 1024;
        CUDA_TRY(cudaMalloc(&h->pslots, cap * sizeof(Slot)));
        CUDA_TRY(cudaMemset(h->pslots, 0, cap * sizeof(Slot)));
        h->pslots_cap = cap;
    }
    if ((size_t)nacc > h->out_cap) {
        dfree(h->out_dev);
        if (h->out_host) { cudaFreeHost(h->out_host); h->out_host = nullptr; }
        const size_t cap = (size_t)nacc + 64;
        CUDA_TRY(cudaMalloc(&h->out_dev, cap * sizeof(double)));
        CUDA_TRY(cudaHostAlloc(&h->out_host, cap * sizeof(Slot), cudaHostAllocMapped));
        memset(h->out_host, 0, cap * sizeof(Slot));
        CUDA_TRY(cudaHostGetDevicePointer(&h->out_host_devptr, h->out_host, 0));
        h->out_cap = cap;
    }
    return 0;
}

static int ensure_scratch(fn_handle* h, size_t bytes) {
    if (bytes > h->scratch_cap) {
        if (h->bh_exec) { cudaGraphExecDestroy(h->bh_exec); h->bh_exec = nullptr; }   // it points into the old scratch
        dfree(h->scratch_dev);
        CUDA_TRY(cudaMalloc(&h->scratch_dev, bytes));
        h->scratch_cap = bytes;
    }
    return 0;
}

#ifdef FN_TRACE
#include <chrono>
static double g_host_trace[4];
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
extern "C" void fn_host_trace(double* out) { for (int i = 0; i < 4; ++i) out[i] = g_host_trace[i]; }
#define FN_HOST_STAMP(i) g_host_trace[i] = now_us()
#else
#define FN_HOST_STAMP(i) do { } while (0)
#endif

static int resident_relaunch(fn_handle* h, unsigned long long next_seq);

static inline bool host_slot_ready(const volatile Slot* slot, unsigned long long seq, double* value) {
    // two aligned 8-byte loads; the pair is accepted only if it is consistent (fn_kernels.cuh, Slot)
    const unsigned long long bits = slot->bits;
    const unsigned long long tag = slot->tag;
    if ((bits ^ tag) != seq_mix(seq)) return false;
    memcpy(value, &bits, sizeof bits);
    return true;
}

// Wait for the kernel that carries `seq` to publish its totals in the pinned result slots and copy
// them to h->out_vals.  Every slot is one {bits, tag} pair of the last CTA that authenticates itself
// against `seq`; spinning on host memory sees the result one PCIe write after the last CTA finishes
// -- several microseconds sooner than cudaStreamSynchronize returns.  Falls back to the stream's
// status to notice a failed launch, and -- resident mode -- a kernel that left on its idle timeout
// before it saw the command, in which case it is launched again (the command is still in place).
static int wait_result(fn_handle* h, unsigned long long seq, int nacc) {
    h->out_vals.resize(nacc);
    unsigned long long spins = 0;
    int relaunches = 0;
    for (int i = 0; i < nacc; ++i) {
        const volatile Slot* slot = &h->out_host[i];
        double value;
        while (!host_slot_ready(slot, seq, &value)) {
            if ((++spins & 0x3fff) == 0) {
                cudaError_t e = cudaStreamQuery(h->stream);
                if (e == cudaSuccess) {
                    if (host_slot_ready(slot, seq, &value)) break;
                    if (h->resident && relaunches < 8) {
                        ++relaunches;
                        int rc = resident_relaunch(h, seq);
                        if (rc) return rc;
                        continue;
                    }
                    return fail(FN_ERR_CUDA, "kernel finished without publishing its result");
                }
                if (e != cudaErrorNotReady) return fail(FN_ERR_CUDA, "kernel failed: %s", cudaGetErrorString(e));
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        h->out_vals[i] = value;
    }
    return 0;
}

// after a plain stream synchronisation the slots of the last launch are complete
static void collect_result(fn_handle* h, int nacc) {
    h->out_vals.resize(nacc);
    for (int i = 0; i < nacc; ++i) memcpy(&h->out_vals[i], (const void*)&h->out_host[i].bits, sizeof(double));
}

static void fill_reduce(fn_handle* h, GridReduce& gr, StepVars& sv, int nacc, bool exchange) {
    gr.pslots = h->pslots; gr.partials = h->partials; gr.ticket = h->ticket; gr.out_dev = h->out_dev;
    gr.out_host = h->out_host_devptr; gr.nacc = nacc;
    sv.host_seq = ++h->host_seq;
    sv.comm_seq = 0; sv.add_on = 0; sv.add_value = 0.0; sv.add_dv = 0.0; sv.add_psum = 0.0;
    gr.comm.world = 0; gr.comm.rank = 0; gr.comm.slot_count = 0; gr.comm.timeout_ns = 0;
    for (int r = 0; r < FN_MAX_RANKS; ++r) gr.comm.mailbox[r] = nullptr;
    if (exchange && h->comm_world > 1) {
        gr.comm.world = h->comm_world; gr.comm.rank = h->comm_rank; gr.comm.slot_count = h->comm_slot;
        sv.comm_seq = ++h->comm_seq;
        gr.comm.timeout_ns = 20ull * 1000ull * 1000ull * 1000ull;
        for (int r = 0; r < h->comm_world; ++r) gr.comm.mailbox[r] = h->comm_peer[r];
    }
}

// ------------------------------------------------------------------------------------------
// upload + prepare
// ------------------------------------------------------------------------------------------
__global__ void controls_to_status_kernel(long long n, long long n_pad, const uint8_t* controls, uint8_t* status) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    status[i] = (i < n && controls && controls[i]) ? 2 : 0;
}

// Host -> device copy of a large array.  From pinned memory: one cudaMemcpyAsync.  From pageable
// memory (a plain numpy array) the runtime's own staging runs at the speed of ONE memcpy thread
// (~10 GB/s); here kStageWorkers threads each copy every kStageWorkers-th 8 MB chunk into their own
// pair of pinned buffers and queue the DMA, so the host copies and the PCIe transfers overlap.
static int copy_h2d(fn_handle* h, void* dst, const void* src, size_t bytes) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, src);
    if (e != cudaSuccess) { cudaGetLastError(); attr.type = cudaMemoryTypeUnregistered; }
    const size_t chunk = 8u << 20;
    if (attr.type != cudaMemoryTypeUnregistered || bytes < 4 * chunk) {
        CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
        return 0;
    }
    const int W = fn_handle::kStageWorkers;
    if (h->stage_bytes != chunk) {
        for (int i = 0; i < 2 * W; ++i) {
            CUDA_TRY(cudaHostAlloc(&h->stage_buf[i], chunk, cudaHostAllocDefault));
            CUDA_TRY(cudaEventCreateWithFlags(&h->stage_ev[i], cudaEventDisableTiming));
        }
        h->stage_bytes = chunk;
    }
    const size_t nchunks = (bytes + chunk - 1) / chunk;
    std::vector<std::thread> workers;
    std::vector<int> status(W, 0);
    for (int w = 0; w < W; ++w) {
        workers.emplace_back([=, &status]() {
            if (cudaSetDevice(h->device) != cudaSuccess) { status[w] = 1; return; }
            int turn = 0;
            for (size_t c = w; c < nchunks; c += W, turn ^= 1) {
                const int slot = 2 * w + turn;
                const size_t off = c * chunk, len = (off + chunk <= bytes) ? chunk : bytes - off;
                // the slot's previous DMA (of this call OR of an earlier copy_h2d whose transfers are
                // still queued on the stream) must have read the buffer before it is overwritten; an
                // event that was never recorded returns at once
                if (cudaEventSynchronize(h->stage_ev[slot]) != cudaSuccess) { status[w] = 1; return; }
                memcpy(h->stage_buf[slot], (const char*)src + off, len);
                if (cudaMemcpyAsync((char*)dst + off, h->stage_buf[slot], len, cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
                    cudaEventRecord(h->stage_ev[slot], h->stream) != cudaSuccess) { status[w] = 1; return; }
            }
        });
    }
    for (auto& t : workers) t.join();
    for (int w = 0; w < W; ++w) if (status[w]) return fail(FN_ERR_CUDA, "staged host-to-device copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// Device -> host copy of a large result array: the mirror image of copy_h2d.  Into pinned memory:
// one cudaMemcpyAsync (completes on the stream).  Into pageable memory (a plain numpy array):
// worker threads each pull every kStageWorkers-th 8 MB chunk through their own pinned buffers and
// copy it out, instead of the runtime's single-threaded staging.  Returns with the data in place
// in the staged case; the caller synchronises the stream afterwards in either case.
static int copy_d2h(fn_handle* h, void* dst, const void* src, size_t bytes) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, dst);
    if (e != cudaSuccess) { cudaGetLastError(); attr.type = cudaMemoryTypeUnregistered; }
    const size_t stage = 8u << 20;           // size of the pinned staging buffers (shared with copy_h2d)
    const size_t chunk = 1u << 20;           // results are tens of MB: small chunks keep all workers busy
    if (attr.type != cudaMemoryTypeUnregistered || bytes < 4 * chunk) {
        CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
        return 0;
    }
    const int W = fn_handle::kStageWorkers;
    if (h->stage_bytes != stage) {
        for (int i = 0; i < 2 * W; ++i) {
            CUDA_TRY(cudaHostAlloc(&h->stage_buf[i], stage, cudaHostAllocDefault));
            CUDA_TRY(cudaEventCreateWithFlags(&h->stage_ev[i], cudaEventDisableTiming));
        }
        h->stage_bytes = stage;
    }
    // the kernels that produce `src` were launched on h->stream: let the copy stream follow them
    cudaEvent_t ready = h->stage_ev[0];
    CUDA_TRY(cudaEventRecord(ready, h->stream));
    CUDA_TRY(cudaEventSynchronize(ready));
    const size_t nchunks = (bytes + chunk - 1) / chunk;
    std::vector<std::thread> workers;
    std::vector<int> status(W, 0);
    for (int w = 0; w < W; ++w) {
        workers.emplace_back([=, &status]() {
            if (cudaSetDevice(h->device) != cudaSuccess) { status[w] = 1; return; }
            cudaStream_t s;
            if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { status[w] = 1; return; }
            for (size_t c = w; c < nchunks; c += W) {
                const size_t off = c * chunk, len = (off + chunk <= bytes) ? chunk : bytes - off;
                void* buf = h->stage_buf[2 * w];
                if (cudaMemcpyAsync(buf, (const char*)src + off, len, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                    cudaStreamSynchronize(s) != cudaSuccess) { status[w] = 1; break; }
                memcpy((char*)dst + off, buf, len);
            }
            cudaStreamDestroy(s);
        });
    }
    for (auto& t : workers) t.join();
    for (int w = 0; w < W; ++w) if (status[w]) return fail(FN_ERR_CUDA, "staged device-to-host copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

#define D2H_TRY(dst, src, bytes)                                  \
    do {                                                          \
        int rc__ = copy_d2h(h, (dst), (src), (bytes));            \
        if (rc__) return rc__;                                    \
    } while (0)

static int upload_common(fn_handle* h, int64_t n, int32_t m, const double* y, const double* aux,
                         const fn_family* family, const uint8_t* controls, int32_t c_noise,
                         const double* design_noise, int32_t c_ctrl, const double* design_ctrl,
                         fn_prep_info* info, cudaMemcpyKind kind) {
    NvtxRange nvtx_range("fitnoise:K1 upload+prepare");
    if (!h) return fail(FN_ERR_ARG, "NULL handle");
    LEAVE_RESIDENT(h);
    // n == 0 is allowed: an empty shard (fewer genes than GPUs) still takes part in every exchange
    if ((!y && n > 0) || !family || !design_noise) return fail(FN_ERR_ARG, "fn_upload: y, family and design_noise are required");
    if (n < 0 || m < 1) return fail(FN_ERR_ARG, "fn_upload: bad shape (n=%lld, m=%d)", (long long)n, m);
    if (n >= (1LL << 31)) return fail(FN_ERR_UNSUPPORTED, "fn_upload: more than 2^31 genes per handle; shard across handles");
    Family fam;
    memcpy(&fam, family, sizeof fam);
    if (!family_valid(fam, m, aux != nullptr))
        return fail(FN_ERR_ARG, "fn_upload: inconsistent family (kind %d, na %d, nb %d, m %d, aux %s)", fam.kind, fam.na,
                    fam.nb, m, aux ? "given" : "missing");
    if ((fam.na == m && fam.kind != KIND_SCALE1) || fam.nb == m)
        if (m > 256) return fail(FN_ERR_UNSUPPORTED, "per-sample variance models support m <= 256 (m = %d)", m);
    CUDA_TRY(cudaSetDevice(h->device));
    Design dn, dc;
    if (c_noise < 1 || c_noise > FN_MAXC) return fail(FN_ERR_UNSUPPORTED, "design has %d columns; 1..%d supported", c_noise, FN_MAXC);
    if (!dn.build(m, c_noise, design_noise)) return fail(FN_ERR_ARG, "noise design (%d x %d) is rank deficient", m, c_noise);
    const double one = 1.0;
    std::vector<double> ones((size_t)m, one);
    if (!design_ctrl) { design_ctrl = ones.data(); c_ctrl = 1; }      // fitnoise.py:216
    if (c_ctrl < 1 || c_ctrl > FN_MAXC) return fail(FN_ERR_UNSUPPORTED, "control design has %d columns; 1..%d supported", c_ctrl, FN_MAXC);
    if (!dc.build(m, c_ctrl, design_ctrl)) return fail(FN_ERR_ARG, "control design (%d x %d) is rank deficient", m, c_ctrl);
    const int width = round_width((dn.c > dc.c ? dn.c : dc.c) + fam.nf);
    if (width < 0) return fail(FN_ERR_UNSUPPORTED, "design columns + factors = %d; at most %d supported", (dn.c > dc.c ? dn.c : dc.c) + fam.nf, FN_MAXC);
    if (fam.nf > 0 && m * fam.nf > 256) return fail(FN_ERR_UNSUPPORTED, "factor models support m * n_factors <= 256 (m = %d, n_factors = %d)", m, fam.nf);
    auto repad = [&](Design& d) {
        std::vector<double> q((size_t)m * width, 0.0);
        for (int j = 0; j < m; ++j)
            for (int i = 0; i < d.c; ++i) q[(size_t)j * width + i] = d.q[(size_t)j * d.width + i];
        d.q.swap(q);
        d.width = width;
    };
    repad(dn); repad(dc);

    CUDA_TRY(cudaStreamSynchronize(h->stream));
    release_data(h);
    h->has_controls = controls != nullptr && n > 0;
    h->n = n; h->m = m; h->fam = fam; h->has_aux = aux != nullptr; h->dn = dn; h->dc = dc; h->width = width;
    h->n_pad = (n + FN_ROW_PAD - 1) / FN_ROW_PAD * FN_ROW_PAD;
    if (h->n_pad == 0) h->n_pad = FN_ROW_PAD;
    const size_t rows = (size_t)h->n_pad * m * sizeof(double);
    const size_t used = (size_t)n * m * sizeof(double);
    CUDA_TRY(ensure_buf(h->b_y, rows));
    h->y = (double*)h->b_y.p;
    CUDA_TRY(cudaMemsetAsync((char*)h->y + used, 0xFF, rows - used, h->stream));       // NaN padding rows
    if (used == 0) { /* empty shard: padding only */ }
    else if (kind == cudaMemcpyHostToDevice) { int rcc = copy_h2d(h, h->y, y, used); if (rcc) return rcc; }
    else CUDA_TRY(cudaMemcpyAsync(h->y, y, used, kind, h->stream));
    if (aux) {
        CUDA_TRY(ensure_buf(h->b_aux, rows));
        h->aux = (double*)h->b_aux.p;
        CUDA_TRY(cudaMemsetAsync((char*)h->aux + used, 0xFF, rows - used, h->stream));
        if (used == 0) { /* empty shard */ }
        else if (kind == cudaMemcpyHostToDevice) { int rcc = copy_h2d(h, h->aux, aux, used); if (rcc) return rcc; }
        else CUDA_TRY(cudaMemcpyAsync(h->aux, aux, used, kind, h->stream));
    }
    CUDA_TRY(ensure_buf(h->b_status, h->n_pad));
    CUDA_TRY(ensure_buf(h->b_cst, h->n_pad * sizeof(double)));
    CUDA_TRY(ensure_buf(h->b_avg, h->n_pad * sizeof(double)));
    h->status = (uint8_t*)h->b_status.p; h->cst = (double*)h->b_cst.p; h->avg = (double*)h->b_avg.p;
    CUDA_TRY(cudaMemsetAsync(h->cst, 0, h->n_pad * sizeof(double), h->stream));
    const bool patseq = fam.kind == KIND_PATSEQ;
    const bool want_gs = patseq && (!fam.tail_own || fam.v3);
    if (want_gs) {
        CUDA_TRY(ensure_buf(h->b_gs, h->n_pad * sizeof(double)));
        h->gs = (double*)h->b_gs.p;
        CUDA_TRY(cudaMemsetAsync(h->gs, 0, h->n_pad * sizeof(double), h->stream));
    }
    {
        uint8_t* ctl_dev = nullptr;
        if (controls && n > 0) {
            if (kind == cudaMemcpyHostToDevice) {
                const int rcs = ensure_scratch(h, (size_t)n);
                if (rcs) return rcs;
                ctl_dev = (uint8_t*)h->scratch_dev;
                CUDA_TRY(cudaMemcpyAsync(ctl_dev, controls, (size_t)n, cudaMemcpyHostToDevice, h->stream));
            } else {
                ctl_dev = const_cast<uint8_t*>(controls);
            }
        }
        const int tb = 256;
        controls_to_status_kernel<<<(unsigned)((h->n_pad + tb - 1) / tb), tb, 0, h->stream>>>(n, h->n_pad, ctl_dev, h->status);
        ++h->launches;
    }
    CUDA_TRY(ensure_buf(h->b_xn, (size_t)m * width * sizeof(double)));
    CUDA_TRY(ensure_buf(h->b_xc, (size_t)m * width * sizeof(double)));
    h->xn = (double*)h->b_xn.p; h->xc = (double*)h->b_xc.p;
    CUDA_TRY(cudaMemcpyAsync(h->xn, dn.q.data(), (size_t)m * width * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->xc, dc.q.data(), (size_t)m * width * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(ensure_buf(h->b_tab, 2 * (size_t)m * sizeof(double)));
    h->tab_dev = (double*)h->b_tab.p;
    if (h->tab_host_cap < 2 * (size_t)m * sizeof(double)) {
        if (h->tab_host) cudaFreeHost(h->tab_host);
        CUDA_TRY(cudaHostAlloc(&h->tab_host, 2 * (size_t)m * sizeof(double), cudaHostAllocDefault));
        h->tab_host_cap = 2 * (size_t)m * sizeof(double);
    }

    // K1
    Geom g;
    const size_t table_bytes = 2 * (size_t)m * width * sizeof(double) + (((size_t)m + 1) * sizeof(unsigned int) + 15) / 16 * 16;
    int rc = choose_geom(h, h->has_aux, false, false, table_bytes, g, 0, 0, h->has_aux && fam.kind != KIND_PATSEQ);
    if (rc) return rc;
    {
        const KernelSet* ks0 = kernel_set(width);
        rc = bound_grid(h, ks0 ? ks0->prepare : nullptr, g);
        if (rc) return rc;
    }
    rc = ensure_reduce(h, g.grid, PREP_NACC);
    if (rc) return rc;
    CUDA_TRY(ensure_buf(h->b_phist, ((size_t)m + 1) * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemsetAsync(h->b_phist.p, 0, ((size_t)m + 1) * sizeof(unsigned long long), h->stream));
    PrepParams pp;
    pp.g = geom_params(h, g);
    pp.src.y = h->y; pp.src.aux = h->aux; pp.src.gs = nullptr; pp.src.cst = nullptr; pp.src.status = h->status;
    pp.src.m = m; pp.src.genes = g.genes; pp.src.row_stride = m; pp.src.l2_keep = 0;
    pp.fam = fam;
    pp.xn = h->xn; pp.xc = h->xc; pp.c_noise = c_noise; pp.c_ctrl = c_ctrl;
    pp.status_out = h->status; pp.cst_out = h->cst; pp.avg_out = h->avg; pp.gs_out = h->gs;
    pp.p_hist = (unsigned long long*)h->b_phist.p;
    fill_reduce(h, pp.gr, pp.sv, PREP_NACC, false);
    {
        const KernelSet* ks = kernel_set(width);
        rc = launch_width_kernel(h, ks ? ks->prepare : nullptr, g.grid, g.threads, g.smem, &pp);
        if (rc) return rc;
    }
    ++h->launches;
    CUDA_TRY(cudaGetLastError());
    if (patseq) {
        counts_to_rc_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(h->aux, (long long)n * m);
        ++h->launches;
        CUDA_TRY(cudaGetLastError());
    }
    h->p_hist.assign((size_t)m + 1, 0);
    static_assert(sizeof(long long) == sizeof(unsigned long long), "histogram word size");
    CUDA_TRY(cudaMemcpyAsync(h->p_hist.data(), h->b_phist.p, ((size_t)m + 1) * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    collect_result(h, PREP_NACC);
    h->psum_expected = 0.0;
    for (int pdf = 0; pdf <= m; ++pdf) h->psum_expected += (double)pdf * (double)h->p_hist[pdf];
    h->cst_sum = h->out_vals[PREP_CSTSUM];
    h->all_complete = !h->has_aux && h->out_vals[PREP_PARTIAL] == 0.0;
    if (info) {
        const double cnt = h->out_vals[PREP_VARCNT];
        info->row_variance_mean = cnt > 0 ? h->out_vals[PREP_VARSUM] / cnt : NAN;
        info->cst_sum = h->cst_sum;
        info->n_eligible = (int64_t)h->out_vals[PREP_ELIG];
        info->n_bad_counts = (int64_t)h->out_vals[PREP_BAD];
        info->n_variance_rows = (int64_t)cnt;
    }
    h->loaded = true;
    return 0;
}

extern "C" int fn_upload(fn_handle* h, int64_t n, int32_t m, const double* y, const double* aux,
                         const fn_family* family, const uint8_t* controls, int32_t c_noise,
                         const double* design_noise, int32_t c_ctrl, const double* design_ctrl,
                         fn_prep_info* info) {
    return upload_common(h, n, m, y, aux, family, controls, c_noise, design_noise, c_ctrl, design_ctrl, info,
                         cudaMemcpyHostToDevice);
}

extern "C" int fn_upload_dev(fn_handle* h, int64_t n, int32_t m, const double* y, const double* aux,
                             const fn_family* family, const uint8_t* controls, int32_t c_noise,
                             const double* design_noise, int32_t c_ctrl, const double* design_ctrl,
                             fn_prep_info* info) {
    return upload_common(h, n, m, y, aux, family, controls, c_noise, design_noise, c_ctrl, design_ctrl, info,
                         cudaMemcpyDeviceToDevice);
}

// ------------------------------------------------------------------------------------------
// evaluation
// ------------------------------------------------------------------------------------------
struct ThetaState {
    double v = 0, theta0 = 1;
    int is_t = 0;
    std::vector<double> ta, tb;
};

// theta -> the scalars and tables the kernels take.  `device_tables`: also copy the two per-sample
// tables (length m each) to h->tab_dev -- the coefficient / weight kernels read them from there, the
// evaluation kernel only when a theta block is per-sample (shared blocks travel as scalars).
static int stage_theta(fn_handle* h, const double* theta, int32_t P, ThetaState& ts, bool device_tables) {
    if (!h || !h->loaded) return fail(FN_ERR_STATE, "no data uploaded");
    if (P != theta_length(h->fam)) return fail(FN_ERR_ARG, "theta has %d entries, the family needs %d", P, theta_length(h->fam));
    if (P > 0 && !theta) return fail(FN_ERR_ARG, "theta is NULL");
    if (!unpack_theta(h->fam, h->m, theta, ts.v, ts.is_t, ts.theta0, ts.ta, ts.tb))
        return fail(FN_ERR_ARG, "theta must be finite and positive");
    if (device_tables && h->fam.kind != KIND_SCALE1) {
        CUDA_TRY(cudaSetDevice(h->device));
        // the previous evaluation was synchronised (or is ordered on the same stream before the copy)
        memcpy(h->tab_host, ts.ta.data(), h->m * sizeof(double));
        memcpy(h->tab_host + h->m, ts.tb.data(), h->m * sizeof(double));
        CUDA_TRY(cudaMemcpyAsync(h->tab_dev, h->tab_host, 2 * (size_t)h->m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    return 0;
}

static bool per_sample_a(const fn_handle* h) { return h->fam.na == h->m && h->fam.kind != KIND_SCALE1; }
static bool per_sample_b(const fn_handle* h) { return h->fam.nb == h->m && h->fam.kind != KIND_SCALE1; }
// which of the two per-sample tables (fn_host.hpp unpack_theta) the evaluation kernel needs in full:
// SCALE_PS uses both (1/theta_j and ln theta_j), PATSEQ one per per-sample block
static bool table_a(const fn_handle* h) { return per_sample_a(h); }
static bool table_b(const fn_handle* h) { return per_sample_b(h) || h->fam.kind == KIND_SCALE_PS; }

static int eval_nacc(const fn_handle* h) {
    return ACC_FIXED + ((per_sample_a(h) || per_sample_b(h)) ? 2 * h->m : 0);
}

// The terms of the objective that depend on (v, p) only, summed over this handle's genes with the
// preparation kernel's histogram N_p of residual df:  sum_p N_p t0[p]  and  d/dv = sum_p N_p t1[p]
// (density_tables: distributions.py:139-146 t / :44-46 normal).  A handful of distinct p occur in
// practice (one when nothing is missing), so this costs well under a microsecond on the host and
// the kernels need no per-evaluation table of ln Gamma / psi.
static void table_terms(const fn_handle* h, const ThetaState& ts, StepVars& sv) {
    double value = 0.5 * h->cst_sum, dv = 0.0;
    for (size_t pdf = 0; pdf < h->p_hist.size(); ++pdf) {
        if (!h->p_hist[pdf]) continue;
        double t0, t1;
        density_tables(ts.is_t, ts.v, (int)pdf, t0, t1);
        value += (double)h->p_hist[pdf] * t0;
        dv += (double)h->p_hist[pdf] * t1;
    }
    sv.add_on = 1; sv.add_value = value; sv.add_dv = dv; sv.add_psum = -h->psum_expected;
}

static int launch_eval(fn_handle* h, const ThetaState& ts, bool per_gene, bool exchange, bool resident = false) {
    const int m = h->m;
    const bool ia = per_sample_a(h), ib = per_sample_b(h);
    const bool inplace = ia || ib;
    Geom g;
    size_t table_bytes = (2 * (size_t)m * h->width + 2 * m + 2 * (m + 1) + ACC_FIXED + 2 * m + (inplace ? 2 * 256 : 0)) * sizeof(double);
    if (h->fam.kind == KIND_SCALE_PS)      // complete-row fast path: two designs' tables, scratch, kappa per gene
        table_bytes += (4 * (size_t)m * h->width + 2 * m + (size_t)h->width * (h->width + 1) / 2 + 8 + 256) * sizeof(double);
    const bool patseq_shared = h->fam.kind == KIND_PATSEQ && h->fam.na == 1 && h->fam.nb == 1 && h->width <= 4 &&
                               !getenv("FN_NO_PATSEQ_SHARED");
    const bool one_big_stage = (h->fam.kind == KIND_SCALE1 && h->has_aux) || (h->fam.kind == KIND_SCALE_PS && !h->has_aux) ||
                               patseq_shared;
    // unweighted per-sample models, m even: thread per gene on padded rows (broadcast table reads); falls
    // back to the lanes-per-gene layout when a stage of 128 such rows does not fit
    bool rows = h->fam.kind == KIND_SCALE_PS && !h->has_aux && m % 2 == 0 && h->tune_lpg <= 1 && !getenv("FN_NO_ROWS");
    const int rows_stride = (m % 4 == 0) ? m + 2 : m;
    int rc = rows ? choose_geom(h, false, h->gs != nullptr, per_gene, table_bytes, g, inplace ? m : 0, 0, false, rows_stride) : -1;
    if (rc) {
        rows = false;
        rc = choose_geom(h, h->has_aux, h->gs != nullptr, per_gene, table_bytes, g, inplace ? m : 0, 0, one_big_stage);
    }
    if (rc) return rc;
    if (inplace && m > g.threads) return fail(FN_ERR_UNSUPPORTED, "per-sample variance models need m <= threads per CTA (%d)", g.threads);
    const KernelSet* ks = kernel_set(h->width);
    const void* kernel = ks ? ks->eval[h->fam.kind] : nullptr;
    // PAT-Seq with one shared theta_a / theta_b: the single-sweep kernel (narrow designs)
    if (ks && h->fam.kind == KIND_PATSEQ && h->fam.na == 1 && h->fam.nb == 1 && ks->eval[KIND_PATSEQ_SHARED] &&
        !getenv("FN_NO_PATSEQ_SHARED"))
        kernel = ks->eval[KIND_PATSEQ_SHARED];
    // nothing missing, nothing weighted: the variants without the masked paths (fewer registers, two CTAs per SM)
    if (ks && h->all_complete && !per_gene && h->fam.kind <= KIND_SCALE_PS && ks->eval_lean[h->fam.kind] && !getenv("FN_NO_LEAN"))
        kernel = ks->eval_lean[h->fam.kind];
    if (!kernel) return fail(FN_ERR_UNSUPPORTED, "no kernel for this design width");
    rc = bound_grid(h, kernel, g);       // resident or not: all CTAs on the GPU at once
    if (rc) return rc;
    const int nacc = eval_nacc(h);
    rc = ensure_reduce(h, g.grid, nacc);
    if (rc) return rc;
    if (per_gene && !h->pg_score) {
        CUDA_TRY(ensure_buf(h->b_pg_score, h->n * sizeof(double)));
        CUDA_TRY(ensure_buf(h->b_pg_d2, h->n * sizeof(double)));
        CUDA_TRY(ensure_buf(h->b_pg_p, h->n * sizeof(int)));
        h->pg_score = (double*)h->b_pg_score.p; h->pg_d2 = (double*)h->b_pg_d2.p; h->pg_p = (int*)h->b_pg_p.p;
    }
    EvalParams ep;
    ep.g = geom_params(h, g);
    ep.src.y = h->y; ep.src.aux = h->aux; ep.src.gs = h->gs; ep.src.cst = per_gene ? h->cst : nullptr;
    ep.src.status = h->status; ep.src.m = m; ep.src.genes = g.genes; ep.src.row_stride = rows ? rows_stride : m;
    {
        // the resident matrices of this handle fit the L2: ask it to keep them between evaluations
        const char* env = getenv("FN_L2_KEEP");
        const size_t resident_bytes = (size_t)h->n_pad * m * sizeof(double) * (h->has_aux ? 2 : 1);
        const size_t limit = (size_t)(env ? atof(env) * 1048576.0 : 0.0);
        ep.src.l2_keep = (limit > 0 && resident_bytes <= limit) ? 1 : 0;
    }
    ep.g.rows = rows ? 1 : 0;
    ep.xn = h->xn; ep.xc = h->xc; ep.c_noise = h->dn.c; ep.c_ctrl = h->dc.c;
    ep.has_ctrl = h->has_controls ? 1 : 0;
    ep.kind = h->fam.kind; ep.tail_own = h->fam.tail_own; ep.v3 = h->fam.v3; ep.is_t = ts.is_t;
    ep.v = ts.v; ep.theta0 = ts.theta0;
    ep.ta0 = ts.ta.empty() ? 0.0 : ts.ta[0];
    ep.tb0 = ts.tb.empty() ? 0.0 : ts.tb[0];
    ep.ta = table_a(h) ? h->tab_dev : nullptr;
    ep.tb = table_b(h) ? h->tab_dev + m : nullptr;
    ep.inplace_a = ia; ep.inplace_b = ib;
    ep.pg_score = per_gene ? h->pg_score : nullptr;
    ep.pg_d2 = per_gene ? h->pg_d2 : nullptr;
    ep.pg_p = per_gene ? h->pg_p : nullptr;
    ep.rc.cmd_host = nullptr; ep.rc.cmd_dev = nullptr; ep.rc.exit_host = nullptr;
    ep.rc.next_seq = 0; ep.rc.idle_timeout_ns = 0; ep.rc.words = 0;
    ep.rc.flush_buf = nullptr; ep.rc.flush_words = 0;
    if (resident) {
        // host_seq / comm_seq advance per command, not per launch
        const unsigned long long hs = h->host_seq, cs = h->comm_seq;
        fill_reduce(h, ep.gr, ep.sv, nacc, exchange);
        h->host_seq = hs; h->comm_seq = cs;
        ep.rc.cmd_host = h->cmd_host_devptr; ep.rc.cmd_dev = h->cmd_dev; ep.rc.exit_host = h->exit_host_devptr;
        ep.rc.next_seq = h->host_seq + 1;
        const char* env = getenv("FN_RESIDENT_IDLE_MS");
        const double idle_ms = env ? atof(env) : 20.0;
        ep.rc.idle_timeout_ns = (unsigned long long)((idle_ms > 0.01 ? idle_ms : 0.01) * 1e6);
        ep.rc.words = h->res_words;
        ep.rc.flush_buf = h->flush_buf; ep.rc.flush_words = h->flush_words;
    } else {
        fill_reduce(h, ep.gr, ep.sv, nacc, exchange);
        if (!per_gene) table_terms(h, ts, ep.sv);
    }
    if (h->timing && !resident) CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    rc = launch_width_kernel(h, kernel, g.grid, g.threads, g.smem, &ep);
    if (rc) return rc;
    if (h->timing && !resident) { CUDA_TRY(cudaEventRecord(h->ev1, h->stream)); h->timing_pending = true; }
    ++h->launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// resident evaluation kernel
// ------------------------------------------------------------------------------------------
static int resident_words(const fn_handle* h) {
    return CW_TAB + (table_a(h) ? h->m : 0) + (table_b(h) ? h->m : 0) + 1;
}

static int resident_launch(fn_handle* h) {
    ThetaState ts;                       // theta arrives with each command
    ts.is_t = h->fam.dist != DIST_NORMAL;
    *(volatile unsigned long long*)h->exit_host = 0;
    // slots left behind by an earlier resident kernel (its QUIT, or the command it fabricated when it
    // gave up waiting) must not be taken for this kernel's first command
    CUDA_TRY(cudaMemsetAsync(h->cmd_dev, 0, h->cmd_cap * sizeof(Slot), h->stream));
    int rc = launch_eval(h, ts, false, true, true);
    if (rc) return rc;
    ++h->resident_launches;
    return 0;
}

static int resident_relaunch(fn_handle* h, unsigned long long next_seq) {
    // the kernel has left (idle timeout); the command `next_seq` is still in the block
    if (h->host_seq + 1 != next_seq && h->host_seq != next_seq) return fail(FN_ERR_STATE, "resident kernel: sequence mismatch");
    const unsigned long long keep = h->host_seq;
    h->host_seq = next_seq - 1;
    int rc = resident_launch(h);
    h->host_seq = keep;
    return rc;
}

extern "C" int fn_resident_begin(fn_handle* h) {
    if (!h || !h->loaded) return fail(FN_ERR_STATE, "no data uploaded");
    if (h->fam.nf > 0) return fail(FN_ERR_UNSUPPORTED, "factor models are evaluated one launch at a time");
    if (h->resident) return 0;
    CUDA_TRY(cudaSetDevice(h->device));
    const int words = resident_words(h);
    if ((size_t)words > h->cmd_cap) {
        if (h->cmd_host) { cudaFreeHost(h->cmd_host); h->cmd_host = nullptr; }
        if (h->cmd_dev) { cudaFree(h->cmd_dev); h->cmd_dev = nullptr; }
        const size_t cap = (size_t)words + 64;
        CUDA_TRY(cudaHostAlloc(&h->cmd_host, cap * sizeof(unsigned long long), cudaHostAllocMapped));
        memset(h->cmd_host, 0, cap * sizeof(unsigned long long));
        CUDA_TRY(cudaHostGetDevicePointer(&h->cmd_host_devptr, h->cmd_host, 0));
        CUDA_TRY(cudaMalloc(&h->cmd_dev, cap * sizeof(Slot)));
        h->cmd_cap = cap;
    }
    if (!h->exit_host) {
        CUDA_TRY(cudaHostAlloc(&h->exit_host, 64, cudaHostAllocMapped));
        CUDA_TRY(cudaHostGetDevicePointer(&h->exit_host_devptr, h->exit_host, 0));
    }
    h->res_words = words;
    int rc = resident_launch(h);
    if (rc) return rc;
    h->resident = true;
    return 0;
}

// Write one command; the sequence number goes last (the kernel accepts the block only with a
// matching checksum, so it can never act on a half-written one).
static void resident_command(fn_handle* h, int op, unsigned long long seq, unsigned long long comm_seq,
                             const ThetaState* ts, const StepVars* sv) {
    volatile unsigned long long* w = h->cmd_host;
    auto bits = [](double x) { unsigned long long b; memcpy(&b, &x, sizeof b); return b; };
    const int words = h->res_words;
    w[CW_OP] = (unsigned long long)op;
    w[CW_COMM_SEQ] = comm_seq;
    w[CW_V] = bits(ts ? ts->v : 0.0);
    w[CW_THETA0] = bits(ts ? ts->theta0 : 1.0);
    w[CW_ADD_VALUE] = bits(sv ? sv->add_value : 0.0);
    w[CW_ADD_DV] = bits(sv ? sv->add_dv : 0.0);
    w[CW_ADD_PSUM] = bits(sv ? sv->add_psum : 0.0);
    w[CW_TA0] = bits(ts && !ts->ta.empty() ? ts->ta[0] : 0.0);
    w[CW_TB0] = bits(ts && !ts->tb.empty() ? ts->tb[0] : 0.0);
    int at = CW_TAB;
    if (table_a(h)) for (int j = 0; j < h->m; ++j) w[at++] = bits(ts ? ts->ta[j] : 0.0);
    if (table_b(h)) for (int j = 0; j < h->m; ++j) w[at++] = bits(ts ? ts->tb[j] : 0.0);
    unsigned long long sum = cmd_mix(seq, CW_SEQ);
    for (int i = 1; i < words - 1; ++i) sum += cmd_mix(w[i], i);
    w[words - 1] = sum;
    std::atomic_thread_fence(std::memory_order_release);
    w[CW_SEQ] = seq;
    std::atomic_thread_fence(std::memory_order_seq_cst);
}

extern "C" int fn_resident_end(fn_handle* h) {
    if (!h) return fail(FN_ERR_ARG, "NULL handle");
    if (!h->resident) return 0;
    CUDA_TRY(cudaSetDevice(h->device));
    const unsigned long long seq = ++h->host_seq;
    resident_command(h, OP_QUIT, seq, 0, nullptr, nullptr);
    h->resident = false;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

static int leave_resident(fn_handle* h) { return (h && h->resident) ? fn_resident_end(h) : 0; }

extern "C" int64_t fn_resident_stats(fn_handle* h, int64_t* launches, int64_t* evaluations) {
    if (!h) return 0;
    if (launches) *launches = h->resident_launches;
    if (evaluations) *evaluations = h->resident_evals;
    return h->resident ? 1 : 0;
}

// Benchmarking aid: evict the data from the L2 (sweep a buffer twice its size) so that the next
// evaluation streams from DRAM.  Works in both forms; returns when the sweep is done.
extern "C" int fn_l2_flush(fn_handle* h) {
    if (!h) return fail(FN_ERR_ARG, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->device));
    if (!h->flush_buf) {
        const bool was = h->resident;
        LEAVE_RESIDENT(h);
        int l2 = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, h->device));
        h->flush_words = (size_t)(l2 > 0 ? l2 : (128 << 20)) * 2 / sizeof(double);
        CUDA_TRY(cudaMalloc(&h->flush_buf, h->flush_words * sizeof(double)));
        CUDA_TRY(cudaMemsetAsync(h->flush_buf, 0, h->flush_words * sizeof(double), h->stream));
        if (was) { int rc = fn_resident_begin(h); if (rc) return rc; }
    }
    if (h->resident) {
        const unsigned long long seq = ++h->host_seq;
        if (*(volatile unsigned long long*)h->exit_host == seq) {
            CUDA_TRY(cudaStreamSynchronize(h->stream));
            int rc = resident_relaunch(h, seq);
            if (rc) return rc;
        }
        resident_command(h, OP_FLUSH, seq, 0, nullptr, nullptr);
        return wait_result(h, seq, 1);
    }
    l2_flush_kernel<<<h->sm_count * 4, 256, 0, h->stream>>>(h->flush_buf, (unsigned long long)h->flush_words);
    ++h->launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

// one evaluation by the resident kernel: the totals are in h->out_vals afterwards
static int resident_eval(fn_handle* h, const ThetaState& ts) {
    StepVars sv;
    table_terms(h, ts, sv);
    const unsigned long long seq = ++h->host_seq;
    const unsigned long long comm_seq = h->comm_world > 1 ? ++h->comm_seq : 0;
    if (*(volatile unsigned long long*)h->exit_host == seq) {
        // the kernel left on its idle timeout while waiting for this very command
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        int rc = resident_relaunch(h, seq);
        if (rc) return rc;
    }
    resident_command(h, OP_EVAL, seq, comm_seq, &ts, &sv);
    FN_HOST_STAMP(2);
    ++h->resident_evals;
    return wait_result(h, seq, eval_nacc(h));
}

// ------------------------------------------------------------------------------------------
// factor models
// ------------------------------------------------------------------------------------------
static int set_factors(fn_handle* h, const double* F) {
    const int m = h->m, nf = h->fam.nf, width = h->width;
    if (nf <= 0) return fail(FN_ERR_STATE, "the uploaded family has no factors");
    if (!F) return fail(FN_ERR_ARG, "factors is NULL");
    { int rcl = leave_resident(h); if (rcl) return rcl; }
    for (int i = 0; i < m * nf; ++i) if (!(fabs(F[i]) < INFINITY)) return fail(FN_ERR_ARG, "factors must be finite");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));        // the previous kernel may still read the tables
    h->factors.assign(F, F + (size_t)m * nf);
    const std::vector<double> zn = augment_design(h->dn, width, nf, F), zc = augment_design(h->dc, width, nf, F);
    std::vector<double> fsq(m, 0.0);
    for (int j = 0; j < m; ++j) for (int f = 0; f < nf; ++f) fsq[j] += F[j * nf + f] * F[j * nf + f];
    const size_t zb = (size_t)m * width * sizeof(double);
    CUDA_TRY(ensure_buf(h->b_zn, zb));
    CUDA_TRY(ensure_buf(h->b_zc, zb));
    CUDA_TRY(ensure_buf(h->b_fsq, m * sizeof(double)));
    h->zn = (double*)h->b_zn.p; h->zc = (double*)h->b_zc.p; h->fsq = (double*)h->b_fsq.p;
    CUDA_TRY(cudaMemcpy(h->zn, zn.data(), zb, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->zc, zc.data(), zb, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->fsq, fsq.data(), m * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int fn_set_factors(fn_handle* h, const double* factors) {
    if (!h || !h->loaded) return fail(FN_ERR_STATE, "no data uploaded");
    return set_factors(h, factors);
}

#define FN_ERR_COMM -5
// A peer that did not deliver its sums within the exchange timeout makes the kernel publish NaN for
// every total: an error, not an objective value.
static int check_totals(fn_handle* h, const double* acc) {
    if (h->comm_world > 1 && acc[ACC_USED] != acc[ACC_USED])
        return fail(FN_ERR_COMM, "multi-GPU exchange failed: a peer rank did not deliver its sums in time "
                                 "(rank %d of %d); the ranks must call fn_nll_grad the same number of times", h->comm_rank, h->comm_world);
    if (acc[ACC_USED] != acc[ACC_USED])      // the grid reducer gave up on a CTA's partial sums (NaN counts are never cast to integers)
        return fail(FN_ERR_CUDA, "evaluation failed: the partial sums of a thread block never arrived");
    return 0;
}

static int launch_eval_factors(fn_handle* h, const ThetaState& ts, bool per_gene, bool exchange) {
    const int m = h->m, nf = h->fam.nf, ncol = m * nf;
    if (!h->zn) return fail(FN_ERR_STATE, "factors have not been set (fn_set_factors / fn_nll_grad_factors)");
    Geom g;
    const size_t table_bytes = (2 * (size_t)m * h->width + 2 * (m + 1) + ACC_FIXED + ncol + 256) * sizeof(double);
    int rc = choose_geom(h, h->has_aux, false, per_gene, table_bytes, g, ncol, (size_t)ncol * sizeof(double));
    if (rc) return rc;
    if (ncol > g.threads) return fail(FN_ERR_UNSUPPORTED, "factor models need m * n_factors <= threads per CTA (%d)", g.threads);
    {
        const KernelSet* ks0 = kernel_set(h->width);
        rc = bound_grid(h, ks0 ? ks0->eval_factors : nullptr, g);
        if (rc) return rc;
    }
    const int nacc = ACC_FIXED + ncol;
    rc = ensure_reduce(h, g.grid, nacc);
    if (rc) return rc;
    if (per_gene && !h->pg_score) {
        CUDA_TRY(ensure_buf(h->b_pg_score, h->n * sizeof(double)));
        CUDA_TRY(ensure_buf(h->b_pg_d2, h->n * sizeof(double)));
        CUDA_TRY(ensure_buf(h->b_pg_p, h->n * sizeof(int)));
        h->pg_score = (double*)h->b_pg_score.p; h->pg_d2 = (double*)h->b_pg_d2.p; h->pg_p = (int*)h->b_pg_p.p;
    }
    EvalFactorsParams ep;
    ep.g = geom_params(h, g);
    ep.src.y = h->y; ep.src.aux = h->aux; ep.src.gs = nullptr; ep.src.cst = per_gene ? h->cst : nullptr;
    ep.src.status = h->status; ep.src.m = m; ep.src.genes = g.genes; ep.src.row_stride = m; ep.src.l2_keep = 0;
    ep.zn = h->zn; ep.zc = h->zc; ep.c_noise = h->dn.c; ep.c_ctrl = h->dc.c; ep.nf = nf;
    ep.is_t = ts.is_t; ep.v = ts.v; ep.theta0 = ts.theta0;
    ep.pg_score = per_gene ? h->pg_score : nullptr;
    ep.pg_d2 = per_gene ? h->pg_d2 : nullptr;
    ep.pg_p = per_gene ? h->pg_p : nullptr;
    fill_reduce(h, ep.gr, ep.sv, nacc, exchange);
    if (!per_gene) table_terms(h, ts, ep.sv);
    if (h->timing) CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    {
        const KernelSet* ks = kernel_set(h->width);
        rc = launch_width_kernel(h, ks ? ks->eval_factors : nullptr, g.grid, g.threads, g.smem, &ep);
        if (rc) return rc;
    }
    if (h->timing) { CUDA_TRY(cudaEventRecord(h->ev1, h->stream)); h->timing_pending = true; }
    ++h->launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int fn_nll_grad_factors(fn_handle* h, const double* theta, int32_t P, const double* factors,
                                   double* value, double* grad, double* grad_factors, int64_t* n_used, int64_t* n_bad) {
    NvtxRange nvtx_range("fitnoise:K2f nll_grad_factors", g_nvtx_hot);
    ThetaState ts;
    int rc = stage_theta(h, theta, P, ts, false);
    if (rc) return rc;
    rc = set_factors(h, factors);
    if (rc) return rc;
    rc = launch_eval_factors(h, ts, false, true);
    if (rc) return rc;
    const int m = h->m, nf = h->fam.nf, nacc = ACC_FIXED + m * nf;
    rc = wait_result(h, h->host_seq, nacc);
    if (rc) return rc;
    const double* acc = h->out_vals.data();
    rc = check_totals(h, acc);
    if (rc) return rc;
    if (value) *value = acc[ACC_VALUE];
    if (grad) {
        int at = 0;
        if (h->fam.dist == DIST_T) grad[at++] = acc[ACC_DV];
        if (h->fam.na == 1) grad[at++] = acc[ACC_GA];
    }
    if (grad_factors) for (int i = 0; i < m * nf; ++i) grad_factors[i] = acc[ACC_FIXED + i];
    if (n_used) *n_used = (int64_t)acc[ACC_USED];
    if (n_bad) *n_bad = (int64_t)acc[ACC_BAD] + (acc[ACC_PSUM] != 0.0 ? 1 : 0);
    return 0;
}

// raw accumulators -> [value, grad[P], n_used, n_bad].  `ta` = theta_a by sample: the PATSEQ kernels
// return term a multiplied by theta_a(j).
static void assemble(const fn_handle* h, const double* acc, const double* ta, double* out) {
    const Family& f = h->fam;
    const int m = h->m;
    const bool scale_a = f.kind == KIND_PATSEQ;
    int at = 0;
    out[at++] = acc[ACC_VALUE];
    if (f.dist == DIST_T) out[at++] = acc[ACC_DV];
    if (f.na == 1) out[at++] = scale_a ? acc[ACC_GA] / ta[0] : acc[ACC_GA];
    else if (f.na == m) for (int j = 0; j < m; ++j) out[at++] = scale_a ? acc[ACC_FIXED + j] / ta[j] : acc[ACC_FIXED + j];
    if (f.nb == 1) out[at++] = acc[ACC_GB];
    else if (f.nb == m) for (int j = 0; j < m; ++j) out[at++] = acc[ACC_FIXED + m + j];
    out[at++] = acc[ACC_USED];
    // a residual-df total that differs from the preparation kernel's histogram means that a sample
    // retained at the start has a non-finite variance at this theta: the reference's fixed `retain`
    // (fitnoise.py:26-28) would feed that variance to the density and get a non-finite score
    out[at++] = acc[ACC_BAD] + (acc[ACC_PSUM] != 0.0 ? 1.0 : 0.0);
}


extern "C" int fn_nll_grad(fn_handle* h, const double* theta, int32_t P, double* value, double* grad,
                           int64_t* n_used, int64_t* n_bad) {
    NvtxRange nvtx_range("fitnoise:K2 nll_grad", g_nvtx_hot);
    FN_HOST_STAMP(0);
    ThetaState ts;
    int rc = stage_theta(h, theta, P, ts, !h->resident && (table_a(h) || table_b(h)));
    if (rc) return rc;
    if (h->fam.nf > 0) return fail(FN_ERR_STATE, "factor models are evaluated with fn_nll_grad_factors");
    FN_HOST_STAMP(1);
    if (h->resident) {
        rc = resident_eval(h, ts);
        if (rc) return rc;
    } else {
        rc = launch_eval(h, ts, false, true);
        if (rc) return rc;
        FN_HOST_STAMP(2);
        rc = wait_result(h, h->host_seq, eval_nacc(h));
        if (rc) return rc;
    }
    FN_HOST_STAMP(3);
    rc = check_totals(h, h->out_vals.data());
    if (rc) return rc;
    double out_small[64];
    std::vector<double> out_big;
    double* out = out_small;
    if (P + 3 > 64) { out_big.resize(P + 3); out = out_big.data(); }
    assemble(h, h->out_vals.data(), ts.ta.data(), out);
    if (value) *value = out[0];
    if (grad) for (int i = 0; i < P; ++i) grad[i] = out[1 + i];
    if (n_used) *n_used = (int64_t)out[P + 1];
    if (n_bad) *n_bad = (int64_t)out[P + 2];
    return 0;
}

// Kernel-only duration: `reps` evaluations queued back to back on the handle's stream between two
// CUDA events (no host wait and no idle gap in between, so launch latency is not counted).
extern "C" int fn_time_nll_grad(fn_handle* h, const double* theta, int32_t P, int32_t reps, double* ms_per_launch) {
    ThetaState ts;
    int rc = leave_resident(h);
    if (rc) return rc;
    rc = stage_theta(h, theta, P, ts, true);
    if (rc) return rc;
    if (h->fam.nf > 0) return fail(FN_ERR_STATE, "factor models are evaluated with fn_nll_grad_factors");
    if (reps < 1 || !ms_per_launch) return fail(FN_ERR_ARG, "fn_time_nll_grad: bad arguments");
    const bool was_timing = h->timing;
    h->timing = false;
    rc = launch_eval(h, ts, false, false);          // warm: attributes set, buffers sized
    if (rc) { h->timing = was_timing; return rc; }
    CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    for (int i = 0; i < reps && rc == 0; ++i) rc = launch_eval(h, ts, false, false);
    h->timing = was_timing;
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    CUDA_TRY(cudaEventSynchronize(h->ev1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    *ms_per_launch = (double)ms / reps;
    return 0;
}

// device-side assemble for the asynchronous (multi-GPU) form
__global__ void assemble_kernel(const double* acc, double* out, Family f, int m, const double* ta) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const bool scale_a = f.kind == KIND_PATSEQ;
    int at = 0;
    out[at++] = acc[ACC_VALUE];
    if (f.dist == DIST_T) out[at++] = acc[ACC_DV];
    if (f.na == 1) out[at++] = scale_a ? acc[ACC_GA] / ta[0] : acc[ACC_GA];
    else if (f.na == m) for (int j = 0; j < m; ++j) out[at++] = scale_a ? acc[ACC_FIXED + j] / ta[j] : acc[ACC_FIXED + j];
    if (f.nb == 1) out[at++] = acc[ACC_GB];
    else if (f.nb == m) for (int j = 0; j < m; ++j) out[at++] = acc[ACC_FIXED + m + j];
    out[at++] = acc[ACC_USED];
    out[at++] = acc[ACC_BAD] + (acc[ACC_PSUM] != 0.0 ? 1.0 : 0.0);
}

extern "C" int fn_nll_grad_dev(fn_handle* h, const double* theta, int32_t P, double* out_dev) {
    if (!out_dev) return fail(FN_ERR_ARG, "out_dev is NULL");
    ThetaState ts;
    int rc = leave_resident(h);
    if (rc) return rc;
    rc = stage_theta(h, theta, P, ts, true);
    if (rc) return rc;
    if (h->fam.nf > 0) return fail(FN_ERR_STATE, "factor models are evaluated with fn_nll_grad_factors");
    rc = launch_eval(h, ts, false, false);
    if (rc) return rc;
    assemble_kernel<<<1, 32, 0, h->stream>>>(h->out_dev, out_dev, h->fam, h->m, h->tab_dev);
    ++h->launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int fn_per_gene(fn_handle* h, const double* theta, int32_t P, double* score, double* d2, int32_t* p) {
    NvtxRange nvtx_range("fitnoise:K2 per_gene");
    ThetaState ts;
    int rc = leave_resident(h);
    if (rc) return rc;
    rc = stage_theta(h, theta, P, ts, true);
    if (rc) return rc;
    rc = h->fam.nf > 0 ? launch_eval_factors(h, ts, true, false) : launch_eval(h, ts, true, false);
    if (rc) return rc;
    if (score) D2H_TRY(score, h->pg_score, h->n * sizeof(double));
    if (d2) D2H_TRY(d2, h->pg_d2, h->n * sizeof(double));
    if (p) D2H_TRY(p, h->pg_p, h->n * sizeof(int));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int fn_noise_stats(fn_handle* h, const double* theta, int32_t P, double* noise_p, double* averages) {
    NvtxRange nvtx_range("fitnoise:K2 noise_stats");
    ThetaState ts;
    int rc = leave_resident(h);
    if (rc) return rc;
    rc = stage_theta(h, theta, P, ts, true);
    if (rc) return rc;
    rc = h->fam.nf > 0 ? launch_eval_factors(h, ts, true, false) : launch_eval(h, ts, true, false);
    if (rc) return rc;
    if (!h->noise_p) {
        CUDA_TRY(ensure_buf(h->b_noise_p, h->n * sizeof(double)));
        h->noise_p = (double*)h->b_noise_p.p;
    }
    const int tb = 128;
    if (h->n > 0) {
        noise_p_kernel<<<(unsigned)((h->n + tb - 1) / tb), tb, 0, h->stream>>>(h->n, h->pg_d2, h->pg_p, ts.is_t, ts.v, h->noise_p);
        ++h->launches;
    }
    CUDA_TRY(cudaGetLastError());
    if (noise_p) D2H_TRY(noise_p, h->noise_p, h->n * sizeof(double));
    if (averages) D2H_TRY(averages, h->avg, h->n * sizeof(double));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------
// coefficients, tests, weights
// ------------------------------------------------------------------------------------------
extern "C" int fn_coef(fn_handle* h, const double* theta, int32_t P, int32_t c, const double* design,
                       double* coef, double* covar, double* dfpost) {
    NvtxRange nvtx_range("fitnoise:K3 coef");
    ThetaState ts;
    int rc = leave_resident(h);
    if (rc) return rc;
    rc = stage_theta(h, theta, P, ts, true);
    if (rc) return rc;
    if (!design || c < 1 || c > FN_MAXC) return fail(FN_ERR_UNSUPPORTED, "design has %d columns; 1..%d supported", c, FN_MAXC);
    const int m = h->m;
    Design d;
    if (!d.build(m, c, design)) return fail(FN_ERR_ARG, "coefficient design (%d x %d) is rank deficient", m, c);
    CUDA_TRY(ensure_buf(h->b_coef, (size_t)h->n * c * sizeof(double)));
    CUDA_TRY(ensure_buf(h->b_covar, (size_t)h->n * c * c * sizeof(double)));
    CUDA_TRY(ensure_buf(h->b_dfpost, (size_t)h->n * sizeof(double)));
    const int nf = h->fam.nf;
    if (nf > 0) {
        if (h->factors.empty()) return fail(FN_ERR_STATE, "factors have not been set (fn_set_factors)");
        const int width = round_width(c + nf);
        if (width < 0) return fail(FN_ERR_UNSUPPORTED, "design columns + factors = %d; at most %d supported", c + nf, FN_MAXC);
        d.q = augment_design(d, width, nf, h->factors.data());
        d.width = width;
    }
    CUDA_TRY(ensure_buf(h->b_xcoef, (size_t)m * d.width * sizeof(double)));
    CUDA_TRY(ensure_buf(h->b_rinv, (size_t)c * c * sizeof(double)));
    h->coef = (double*)h->b_coef.p; h->covar = (double*)h->b_covar.p; h->dfpost = (double*)h->b_dfpost.p;
    h->x_coef = (double*)h->b_xcoef.p; h->rinv_dev = (double*)h->b_rinv.p;
    CUDA_TRY(cudaMemcpyAsync(h->x_coef, d.q.data(), (size_t)m * d.width * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->rinv_dev, d.rinv.data(), (size_t)c * c * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    h->coef_c = c;
    h->coef_is_t = ts.is_t != 0;

    Geom g;
    const size_t table_bytes = ((size_t)m * d.width + 2 * m) * sizeof(double);
    rc = choose_geom(h, h->has_aux, h->gs != nullptr, false, table_bytes, g, 0, 0, h->has_aux && h->fam.kind != KIND_PATSEQ);
    if (rc) return rc;
    CoefParams cp;
    cp.g = geom_params(h, g);
    cp.g.width = d.width;
    cp.src.y = h->y; cp.src.aux = h->aux; cp.src.gs = h->gs; cp.src.cst = nullptr; cp.src.status = h->status;
    cp.src.m = m; cp.src.genes = g.genes; cp.src.row_stride = m; cp.src.l2_keep = 0;
    cp.x = h->x_coef; cp.rinv = h->rinv_dev; cp.c = c; cp.nf = nf;
    cp.kind = h->fam.kind; cp.tail_own = h->fam.tail_own; cp.v3 = h->fam.v3; cp.is_t = ts.is_t;
    cp.v = ts.v; cp.theta0 = ts.theta0;
    cp.ta = h->fam.kind != KIND_SCALE1 ? h->tab_dev : nullptr;
    cp.tb = h->fam.kind != KIND_SCALE1 ? h->tab_dev + m : nullptr;
    cp.coef = h->coef; cp.covar = h->covar; cp.dfpost = h->dfpost;
    {
        const KernelSet* ks = kernel_set(d.width);
        rc = launch_width_kernel(h, ks ? ks->coef : nullptr, g.grid, g.threads, g.smem, &cp);
        if (rc) return rc;
    }
    ++h->launches;
    CUDA_TRY(cudaGetLastError());
    if (coef) D2H_TRY(coef, h->coef, (size_t)h->n * c * sizeof(double));
    if (covar) D2H_TRY(covar, h->covar, (size_t)h->n * c * c * sizeof(double));
    if (dfpost) D2H_TRY(dfpost, h->dfpost, (size_t)h->n * sizeof(double));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

static int bh_run(fn_handle* h, int64_t n, const double* p_dev, double* q_dev, char* work, uint32_t** order_dev = nullptr);
static size_t bh_work_bytes(int64_t n);

static int test_core(fn_handle* h, int32_t kc, const double* contrasts, double* contrasted, double* ccov,
                     double* pval, double* qval, uint32_t* order) {
    NvtxRange nvtx_range("fitnoise:K3b test");
    if (!h || !h->loaded || h->coef_c == 0) return fail(FN_ERR_STATE, "fn_test needs a preceding fn_coef");
    LEAVE_RESIDENT(h);
    const int c = h->coef_c;
    if (kc < 1 || kc > FN_MAXC || !contrasts) return fail(FN_ERR_UNSUPPORTED, "contrasts has %d columns; 1..%d supported", kc, FN_MAXC);
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t nb_con = (size_t)c * kc * sizeof(double);
    const size_t nb_out = (size_t)h->n * kc * sizeof(double);
    const size_t nb_cc = ccov ? (size_t)h->n * kc * kc * sizeof(double) : 0;
    const size_t nb_p = (size_t)h->n * sizeof(double);
    const size_t nb_bh = qval ? nb_p + 256 + bh_work_bytes(h->n) : 0;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    int rc = ensure_scratch(h, up(nb_con) + up(nb_out) + up(nb_cc) + up(nb_p) + up(nb_bh));
    if (rc) return rc;
    char* base = (char*)h->scratch_dev;
    double* d_con = (double*)base; base += up(nb_con);
    double* d_out = (double*)base; base += up(nb_out);
    double* d_cc = ccov ? (double*)base : nullptr; base += up(nb_cc);
    double* d_p = (double*)base; base += up(nb_p);
    double* d_q = (double*)base; base += up(nb_p);
    CUDA_TRY(cudaMemcpyAsync(d_con, contrasts, nb_con, cudaMemcpyHostToDevice, h->stream));
    TestParams tp;
    tp.n = h->n; tp.c = c; tp.kc = kc; tp.is_t = h->coef_is_t;
    tp.coef = h->coef; tp.covar = h->covar; tp.dfpost = h->dfpost; tp.contrasts = d_con;
    tp.contrasted = d_out; tp.ccov = d_cc; tp.pval = d_p;
    h->test_out = d_out; h->test_p = d_p; h->test_q = nullptr; h->test_order = nullptr; h->test_kc = kc;
    const int tb = 128;
    if (h->n > 0) {
        const unsigned blocks = (unsigned)((h->n + tb - 1) / tb);
        switch (kc) {
            case 1: test_kernel<1><<<blocks, tb, 0, h->stream>>>(tp); break;
            case 2: test_kernel<2><<<blocks, tb, 0, h->stream>>>(tp); break;
            case 3: test_kernel<3><<<blocks, tb, 0, h->stream>>>(tp); break;
            case 4: test_kernel<4><<<blocks, tb, 0, h->stream>>>(tp); break;
            default: test_kernel<0><<<blocks, tb, 0, h->stream>>>(tp); break;
        }
        ++h->launches;
    }
    CUDA_TRY(cudaGetLastError());
    if (contrasted) D2H_TRY(contrasted, d_out, nb_out);
    if (ccov) D2H_TRY(ccov, d_cc, nb_cc);
    if (pval) D2H_TRY(pval, d_p, nb_p);
    if (qval && h->n > 0) {
        uint32_t* d_order = nullptr;
        rc = bh_run(h, h->n, d_p, d_q, base, &d_order);   // p never leaves the device between test and BH
        if (rc) return rc;
        h->test_q = d_q; h->test_order = d_order;
        D2H_TRY(qval, d_q, nb_p);
        if (order) D2H_TRY(order, d_order, (size_t)h->n * sizeof(uint32_t));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int fn_test(fn_handle* h, int32_t kc, const double* contrasts, double* contrasted, double* ccov, double* pval) {
    return test_core(h, kc, contrasts, contrasted, ccov, pval, nullptr, nullptr);
}

extern "C" int fn_test_bh(fn_handle* h, int32_t kc, const double* contrasts, double* contrasted, double* pval,
                          double* qval, uint32_t* order) {
    if (!qval) return fail(FN_ERR_ARG, "fn_test_bh: qval is NULL");
    return test_core(h, kc, contrasts, contrasted, nullptr, pval, qval, order);
}

extern "C" int fn_coef_fetch(fn_handle* h, double* coef, double* covar, double* dfpost) {
    if (!h || !h->loaded || h->coef_c == 0) return fail(FN_ERR_STATE, "fn_coef_fetch needs a preceding fn_coef");
    LEAVE_RESIDENT(h);
    const int c = h->coef_c;
    CUDA_TRY(cudaSetDevice(h->device));
    if (coef) D2H_TRY(coef, h->coef, (size_t)h->n * c * sizeof(double));
    if (covar) D2H_TRY(covar, h->covar, (size_t)h->n * c * c * sizeof(double));
    if (dfpost) D2H_TRY(dfpost, h->dfpost, (size_t)h->n * sizeof(double));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int fn_weights(fn_handle* h, const double* theta, int32_t P, double* weights) {
    NvtxRange nvtx_range("fitnoise:weights");
    ThetaState ts;
    int rc = leave_resident(h);
    if (rc) return rc;
    rc = stage_theta(h, theta, P, ts, true);
    if (rc) return rc;
    if (!weights) return fail(FN_ERR_ARG, "weights is NULL");
    const size_t bytes = (size_t)h->n * h->m * sizeof(double);
    rc = ensure_scratch(h, bytes);
    if (rc) return rc;
    WeightParams wp;
    wp.n = h->n; wp.m = h->m; wp.y = h->y; wp.aux = h->aux; wp.gs = h->gs;
    wp.kind = h->fam.kind; wp.tail_own = h->fam.tail_own; wp.v3 = h->fam.v3; wp.theta0 = ts.theta0;
    wp.ta = h->tab_dev; wp.tb = h->tab_dev + h->m;
    wp.fsq = h->fam.nf > 0 ? h->fsq : nullptr;
    if (h->fam.nf > 0 && !h->fsq) return fail(FN_ERR_STATE, "factors have not been set (fn_set_factors)");
    wp.out = (double*)h->scratch_dev;
    weights_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(wp);
    ++h->launches;
    CUDA_TRY(cudaGetLastError());
    D2H_TRY(weights, wp.out, bytes);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------
// Benjamini-Hochberg
// ------------------------------------------------------------------------------------------
// Enqueue the kernels of one Benjamini-Hochberg adjustment on the handle's stream (15 launches in the
// one-sweep form, 44 in the five-kernels-per-digit form).
// number of kernel launches / graph nodes of one adjustment (what h->launches counts)
static bool bh_onesweep(int64_t n) {
    if (getenv("FN_BH_CLASSIC")) return false;
    return n <= BH_ONESWEEP_MAX || getenv("FN_BH_ONESWEEP") != nullptr;
}
static int bh_launch_count(int64_t n) { return bh_onesweep(n) ? 15 : 44; }

static void bh_enqueue(fn_handle* h, int64_t n, const double* p_dev, double* q_dev, char* work, uint32_t** order_dev) {
    const int nblocks = (int)((n + BH_CHUNK - 1) / BH_CHUNK);
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    uint64_t* keys0 = (uint64_t*)work; work += up(n * 8);
    uint64_t* keys1 = (uint64_t*)work; work += up(n * 8);
    uint32_t* idx0 = (uint32_t*)work; work += up(n * 4);
    uint32_t* idx1 = (uint32_t*)work; work += up(n * 4);
    uint32_t* hist = (uint32_t*)work; work += up((size_t)256 * nblocks * 4);
    double* chunk_min = (double*)work; work += up((size_t)nblocks * 8);
    double* carry = (double*)work; work += up((size_t)nblocks * 8);
    uint32_t* seg = (uint32_t*)work;
    if (bh_onesweep(n)) {
        // one-sweep sort (fn_bh.cuh): [digit counts 8 x 256][tickets 8 (padded)][state words 8 x tiles x 256], zeroed together
        const size_t nseg_bytes = up((((size_t)256 * nblocks + BH_SEG - 1) / BH_SEG) * 4);
        char* zone = (char*)seg + nseg_bytes;
        uint32_t* hist_all = (uint32_t*)zone;
        unsigned int* tickets = (unsigned int*)(zone + 8 * 256 * 4);
        unsigned long long* state = (unsigned long long*)(zone + 8 * 256 * 4 + 256);
        const size_t state_words = (size_t)nblocks * 256;
        cudaMemsetAsync(zone, 0, 8 * 256 * 4 + 256 + 8 * state_words * 8, h->stream);
        bh_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, p_dev, keys0, idx0);
        bh_hist_all_kernel<<<nblocks, BH_THREADS, 0, h->stream>>>(n, keys0, hist_all);
        bh_bases_kernel<<<8, 256, 0, h->stream>>>(hist_all);
        for (int pass = 0; pass < 8; ++pass) {
            bh_onesweep_kernel<<<nblocks, BH_THREADS, 0, h->stream>>>(n, keys0, idx0, keys1, idx1, 8 * pass, hist_all + 256 * pass,
                                                                      state + (size_t)pass * state_words, tickets + pass);
            std::swap(keys0, keys1);
            std::swap(idx0, idx1);
        }
        bh_chunkmin_kernel<<<nblocks, BH_THREADS, 0, h->stream>>>(n, keys0, chunk_min);
        bh_carry_scan_kernel<<<1, 1024, 0, h->stream>>>(nblocks, chunk_min, carry);
        bh_finish_kernel<<<nblocks, BH_THREADS, 0, h->stream>>>(n, keys0, idx0, carry, q_dev);
        if (order_dev) *order_dev = idx0;
        return;
    }
    const long long hist_total = (long long)256 * nblocks;
    const int nseg = (int)((hist_total + BH_SEG - 1) / BH_SEG);
    bh_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, p_dev, keys0, idx0);
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 8 * pass;
        bh_hist_kernel<<<nblocks, BH_THREADS, 0, h->stream>>>(n, keys0, shift, hist, nblocks);
        bh_segsum_kernel<<<nseg, BH_THREADS, 0, h->stream>>>(hist, hist_total, seg);
        bh_segscan_kernel<<<1, 1024, 0, h->stream>>>(seg, nseg);
        bh_segapply_kernel<<<nseg, BH_THREADS, 0, h->stream>>>(hist, hist_total, seg);
        bh_scatter_kernel<<<nblocks, BH_THREADS, 0, h->stream>>>(n, keys0, idx0, keys1, idx1, shift, hist, nblocks);
        std::swap(keys0, keys1);
        std::swap(idx0, idx1);
    }
    bh_chunkmin_kernel<<<nblocks, BH_THREADS, 0, h->stream>>>(n, keys0, chunk_min);
    bh_carry_scan_kernel<<<1, 1024, 0, h->stream>>>(nblocks, chunk_min, carry);
    bh_finish_kernel<<<nblocks, BH_THREADS, 0, h->stream>>>(n, keys0, idx0, carry, q_dev);
    if (order_dev) *order_dev = idx0;          // genes by ascending p (stable, NaN last)
}

// The adjustment is a fixed sequence of 15 or 44 small launches -- launch-bound at the usual 10^4..10^6
// p-values -- so it is captured once into a CUDA graph per (n, buffers) and replayed with a single
// cudaGraphLaunch afterwards.  Any failure of the capture machinery falls back to plain launches.
static int bh_run(fn_handle* h, int64_t n, const double* p_dev, double* q_dev, char* work, uint32_t** order_dev) {
    NvtxRange nvtx_range("fitnoise:K4 bh");
    const bool use_graph = getenv("FN_NO_GRAPH") == nullptr;
    if (use_graph && h->bh_exec && h->bh_n == n && h->bh_p == p_dev && h->bh_q == q_dev && h->bh_work == work) {
        if (cudaGraphLaunch(h->bh_exec, h->stream) == cudaSuccess) {
            h->launches += bh_launch_count(n);
            if (order_dev) *order_dev = h->bh_order;
            return 0;
        }
        cudaGetLastError();
    }
    if (h->bh_exec) { cudaGraphExecDestroy(h->bh_exec); h->bh_exec = nullptr; }
    if (use_graph && cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        uint32_t* order = nullptr;
        bh_enqueue(h, n, p_dev, q_dev, work, &order);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
        if (e == cudaSuccess && graph && cudaGraphInstantiate(&h->bh_exec, graph, 0) == cudaSuccess) {
            cudaGraphDestroy(graph);
            h->bh_n = n; h->bh_p = p_dev; h->bh_q = q_dev; h->bh_work = work; h->bh_order = order;
            if (cudaGraphLaunch(h->bh_exec, h->stream) == cudaSuccess) {
                h->launches += bh_launch_count(n);
                if (order_dev) *order_dev = order;
                return 0;
            }
        }
        if (graph) cudaGraphDestroy(graph);
        if (h->bh_exec) { cudaGraphExecDestroy(h->bh_exec); h->bh_exec = nullptr; }
        cudaGetLastError();
    } else {
        cudaGetLastError();
    }
    bh_enqueue(h, n, p_dev, q_dev, work, order_dev);
    h->launches += bh_launch_count(n);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static size_t bh_work_bytes(int64_t n) {
    const size_t nblocks = (n + BH_CHUNK - 1) / BH_CHUNK;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t nseg = (256 * nblocks + BH_SEG - 1) / BH_SEG;
    // + the one-sweep zone: digit counts, tickets, 8 x tiles x 256 state words
    return 2 * up(n * 8) + 2 * up(n * 4) + up(256 * nblocks * 4) + 2 * up(nblocks * 8) + up(nseg * 4) +
           (bh_onesweep(n) ? up(8 * 256 * 4 + 256 + 8 * nblocks * 256 * 8) : 0);
}

extern "C" int fn_bh_dev(fn_handle* h, int64_t n, const double* p_dev, double* q_dev) {
    if (!h) return fail(FN_ERR_ARG, "NULL handle");
    if (n < 0 || n >= (1LL << 32)) return fail(FN_ERR_UNSUPPORTED, "fn_bh: n out of range");
    if (n == 0) return 0;
    LEAVE_RESIDENT(h);
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = ensure_scratch(h, bh_work_bytes(n));
    if (rc) return rc;
    return bh_run(h, n, p_dev, q_dev, (char*)h->scratch_dev);
}

static int bh_host(fn_handle* h, int64_t n, const double* p, double* q, uint32_t* order);

extern "C" int fn_bh(fn_handle* h, int64_t n, const double* p, double* q) { return bh_host(h, n, p, q, nullptr); }

extern "C" int fn_bh_ranked(fn_handle* h, int64_t n, const double* p, double* q, uint32_t* order) {
    return bh_host(h, n, p, q, order);
}

static int bh_host(fn_handle* h, int64_t n, const double* p, double* q, uint32_t* order) {
    if (!h) return fail(FN_ERR_ARG, "NULL handle");
    if (n < 0 || n >= (1LL << 32)) return fail(FN_ERR_UNSUPPORTED, "fn_bh: n out of range");
    if (n == 0) return 0;
    if (!p || !q) return fail(FN_ERR_ARG, "fn_bh: NULL array");
    LEAVE_RESIDENT(h);
    CUDA_TRY(cudaSetDevice(h->device));
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    int rc = ensure_scratch(h, 2 * up(n * 8) + bh_work_bytes(n));
    if (rc) return rc;
    char* base = (char*)h->scratch_dev;
    double* p_dev = (double*)base; base += up(n * 8);
    double* q_dev = (double*)base; base += up(n * 8);
    CUDA_TRY(cudaMemcpyAsync(p_dev, p, n * 8, cudaMemcpyHostToDevice, h->stream));
    uint32_t* d_order = nullptr;
    rc = bh_run(h, n, p_dev, q_dev, base, &d_order);
    if (rc) return rc;
    D2H_TRY(q, q_dev, n * 8);
    if (order) D2H_TRY(order, d_order, n * sizeof(uint32_t));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------
// device-resident results, pinned host memory
// ------------------------------------------------------------------------------------------
// Where the last per-gene results live on the device (for device-to-device gathers across ranks).
extern "C" int fn_result_dev(fn_handle* h, int32_t which, void** ptr, int64_t* rows, int32_t* cols) {
    if (!h || !h->loaded || !ptr || !rows || !cols) return fail(FN_ERR_ARG, "fn_result_dev: bad arguments");
    void* p = nullptr;
    int c = 1;
    switch (which) {
        case FN_RES_SCORE: p = h->pg_score; break;
        case FN_RES_D2: p = h->pg_d2; break;
        case FN_RES_PDF: p = h->pg_p; break;
        case FN_RES_NOISE_P: p = h->noise_p; break;
        case FN_RES_AVERAGES: p = h->avg; break;
        case FN_RES_COEF: p = h->coef; c = h->coef_c; break;
        case FN_RES_COVAR: p = h->covar; c = h->coef_c * h->coef_c; break;
        case FN_RES_DFPOST: p = h->dfpost; break;
        case FN_RES_CONTRASTED: p = h->test_out; c = h->test_kc; break;
        case FN_RES_P: p = h->test_p; break;
        case FN_RES_Q: p = h->test_q; break;
        case FN_RES_ORDER: p = h->test_order; break;
        default: return fail(FN_ERR_ARG, "fn_result_dev: unknown result %d", which);
    }
    if (!p) return fail(FN_ERR_STATE, "fn_result_dev: result %d has not been computed", which);
    *ptr = p; *rows = h->n; *cols = c;
    return 0;
}

// Device -> host copy on the handle's stream (pinned destination: one DMA; pageable: staged), then wait.
extern "C" int fn_fetch(fn_handle* h, void* dst_host, const void* src_dev, int64_t bytes) {
    if (!h || (bytes > 0 && (!dst_host || !src_dev)) || bytes < 0) return fail(FN_ERR_ARG, "fn_fetch: bad arguments");
    LEAVE_RESIDENT(h);
    CUDA_TRY(cudaSetDevice(h->device));
    if (bytes) D2H_TRY(dst_host, src_dev, (size_t)bytes);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}

// Pinned host memory from a process-wide pool: result arrays are DMA targets (PCIe rate instead of
// the page-faulting, single-threaded path into fresh pageable memory), and pinning itself is slow
// (~1 ms per 4 MB), so freed blocks are kept and handed out again.
static std::mutex g_pin_mutex;
static std::multimap<size_t, void*> g_pin_free;
static std::map<void*, size_t> g_pin_size;
static size_t g_pin_held = 0;

extern "C" int fn_pinned_alloc(int64_t bytes, void** out) {
    if (!out || bytes < 0) return fail(FN_ERR_ARG, "fn_pinned_alloc: bad arguments");
    size_t cap = 4096;
    while (cap < (size_t)bytes) cap *= 2;
    if (cap > (64u << 20)) cap = ((size_t)bytes + (16u << 20) - 1) / (16u << 20) * (16u << 20);
    {
        std::lock_guard<std::mutex> lock(g_pin_mutex);
        auto it = g_pin_free.lower_bound(cap);
        if (it != g_pin_free.end() && it->first <= 2 * cap) {
            *out = it->second;
            g_pin_held -= it->first;
            g_pin_free.erase(it);
            return 0;
        }
    }
    void* p = nullptr;
    CUDA_TRY(cudaHostAlloc(&p, cap, cudaHostAllocPortable));
    std::lock_guard<std::mutex> lock(g_pin_mutex);
    g_pin_size[p] = cap;
    *out = p;
    return 0;
}

extern "C" int fn_pinned_free(void* p) {
    if (!p) return 0;
    std::lock_guard<std::mutex> lock(g_pin_mutex);
    auto it = g_pin_size.find(p);
    if (it == g_pin_size.end()) return fail(FN_ERR_ARG, "fn_pinned_free: not a block of fn_pinned_alloc");
    const size_t cap = it->second;
    if (g_pin_held + cap > ((size_t)4 << 30)) {          // keep at most 4 GB of idle pinned memory
        g_pin_size.erase(it);
        cudaFreeHost(p);
        return 0;
    }
    g_pin_free.insert(std::make_pair(cap, p));
    g_pin_held += cap;
    return 0;
}

// Benjamini-Hochberg on device-resident p (n values, e.g. gathered from all ranks): q_dev receives the
// q-values, *order_dev points at the genes by ascending p (inside the handle's scratch; valid until
// the next call that uses it).
extern "C" int fn_bh_ranked_dev(fn_handle* h, int64_t n, const double* p_dev, double* q_dev, uint32_t** order_dev) {
    if (!h) return fail(FN_ERR_ARG, "NULL handle");
    if (n < 0 || n >= (1LL << 32)) return fail(FN_ERR_UNSUPPORTED, "fn_bh: n out of range");
    if (n == 0) return 0;
    LEAVE_RESIDENT(h);
    CUDA_TRY(cudaSetDevice(h->device));
    // the scratch may hold the last test's results (p_dev may even point into it): use a dedicated buffer
    CUDA_TRY(ensure_buf(h->b_bhwork, bh_work_bytes(n)));
    return bh_run(h, n, p_dev, q_dev, (char*)h->b_bhwork.p, order_dev);
}

// ------------------------------------------------------------------------------------------
// fused multi-GPU exchange: mailboxes shared between the ranks' processes with CUDA IPC
// ------------------------------------------------------------------------------------------
extern "C" int fn_comm_mailbox(fn_handle* h, int32_t world, int32_t max_values, void* ipc_handle_out) {
    if (!h || !ipc_handle_out) return fail(FN_ERR_ARG, "fn_comm_mailbox: NULL argument");
    if (world < 2 || world > FN_MAX_RANKS) return fail(FN_ERR_ARG, "fn_comm_mailbox: world must be 2..%d", FN_MAX_RANKS);
    LEAVE_RESIDENT(h);
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->comm_own) { cudaFree(h->comm_own); h->comm_own = nullptr; }
    h->comm_slot = ((max_values + 15) / 16) * 16;
    const size_t bytes = (size_t)2 * world * h->comm_slot * sizeof(Slot);
    CUDA_TRY(cudaMalloc(&h->comm_own, bytes));
    CUDA_TRY(cudaMemset(h->comm_own, 0, bytes));
    CUDA_TRY(cudaDeviceSynchronize());
    cudaIpcMemHandle_t handle;
    CUDA_TRY(cudaIpcGetMemHandle(&handle, h->comm_own));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    memcpy(ipc_handle_out, &handle, sizeof handle);
    h->comm_world = 0;          // not connected yet
    return 0;
}

extern "C" int fn_comm_connect(fn_handle* h, int32_t rank, int32_t world, const void* ipc_handles) {
    if (!h || !ipc_handles || !h->comm_own) return fail(FN_ERR_STATE, "fn_comm_connect: call fn_comm_mailbox first");
    if (world < 2 || world > FN_MAX_RANKS || rank < 0 || rank >= world) return fail(FN_ERR_ARG, "fn_comm_connect: bad rank/world");
    LEAVE_RESIDENT(h);
    CUDA_TRY(cudaSetDevice(h->device));
    for (int r = 0; r < world; ++r) {
        if (r == rank) { h->comm_peer[r] = h->comm_own; continue; }
        cudaIpcMemHandle_t handle;
        memcpy(&handle, (const char*)ipc_handles + (size_t)r * sizeof handle, sizeof handle);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, handle, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            // close the mappings opened so far: comm_world stays 0, so fn_comm_disconnect would not
            for (int q = 0; q < r; ++q)
                if (q != rank && h->comm_peer[q]) cudaIpcCloseMemHandle(h->comm_peer[q]);
            for (int q = 0; q < FN_MAX_RANKS; ++q) h->comm_peer[q] = nullptr;
            return fail(FN_ERR_CUDA, "cudaIpcOpenMemHandle of rank %d's mailbox failed: %s", r, cudaGetErrorString(e));
        }
        h->comm_peer[r] = (Slot*)ptr;
    }
    h->comm_rank = rank; h->comm_world = world; h->comm_seq = 0;
    return 0;
}

extern "C" int fn_comm_disconnect(fn_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    leave_resident(h);
    cudaStreamSynchronize(h->stream);
    for (int r = 0; r < h->comm_world; ++r)
        if (r != h->comm_rank && h->comm_peer[r]) cudaIpcCloseMemHandle(h->comm_peer[r]);
    for (int r = 0; r < FN_MAX_RANKS; ++r) h->comm_peer[r] = nullptr;
    h->comm_world = 0;
    if (h->comm_own) { cudaFree(h->comm_own); h->comm_own = nullptr; }
    return 0;
}

#ifdef FN_TRACE
extern "C" int fn_trace_dump(fn_handle* h, unsigned long long* out64) {
    LEAVE_RESIDENT(h);
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpyFromSymbol(out64, fn::g_trace, 64 * sizeof(unsigned long long)));
    return 0;
}
#endif

// ------------------------------------------------------------------------------------------
// fitnoise.transform: moderated log transform of counts (fittransform.py)
// ------------------------------------------------------------------------------------------
extern "C" int fn_transform_upload(fn_handle* h, int64_t n, int32_t m, const double* x, int32_t c, const double* design) {
    if (!h) return fail(FN_ERR_ARG, "NULL handle");
    if (n < 0 || m < 1 || (!x && n > 0)) return fail(FN_ERR_ARG, "fn_transform_upload: bad arguments");
    if (m > 256) return fail(FN_ERR_UNSUPPORTED, "transform supports m <= 256 (m = %d)", m);
    if (c < 1 || c > TR_MAXC || !design) return fail(FN_ERR_UNSUPPORTED, "design has %d columns; 1..%d supported", c, TR_MAXC);
    Design d;
    if (!d.build(m, c, design)) return fail(FN_ERR_ARG, "design (%d x %d) is rank deficient", m, c);
    LEAVE_RESIDENT(h);
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    std::vector<double> q((size_t)m * c);
    for (int j = 0; j < m; ++j) for (int i = 0; i < c; ++i) q[(size_t)j * c + i] = d.q[(size_t)j * d.width + i];
    const size_t bytes = (size_t)n * m * sizeof(double);
    CUDA_TRY(ensure_buf(h->b_tx, bytes));
    CUDA_TRY(ensure_buf(h->b_tq, q.size() * sizeof(double)));
    if (bytes) { int rc = copy_h2d(h, h->b_tx.p, x, bytes); if (rc) return rc; }
    CUDA_TRY(cudaMemcpyAsync(h->b_tq.p, q.data(), q.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->tr_n = n; h->tr_m = m; h->tr_c = c;
    return 0;
}

#define FN_DISPATCH_TRANSFORM(K_, ORDER_, ...)                                                      \
    switch ((K_) * 10 + (ORDER_)) {                                                                 \
        case 11: { constexpr int K = 1, ORDER = 1; __VA_ARGS__; } break;                            \
        case 12: { constexpr int K = 1, ORDER = 2; __VA_ARGS__; } break;                            \
        case 13: { constexpr int K = 1, ORDER = 3; __VA_ARGS__; } break;                            \
        case 21: { constexpr int K = 2, ORDER = 1; __VA_ARGS__; } break;                            \
        case 22: { constexpr int K = 2, ORDER = 2; __VA_ARGS__; } break;                            \
        case 23: { constexpr int K = 2, ORDER = 3; __VA_ARGS__; } break;                            \
        case 31: { constexpr int K = 3, ORDER = 1; __VA_ARGS__; } break;                            \
        case 32: { constexpr int K = 3, ORDER = 2; __VA_ARGS__; } break;                            \
        case 33: { constexpr int K = 3, ORDER = 3; __VA_ARGS__; } break;                            \
        case 41: { constexpr int K = 4, ORDER = 1; __VA_ARGS__; } break;                            \
        case 42: { constexpr int K = 4, ORDER = 2; __VA_ARGS__; } break;                            \
        case 43: { constexpr int K = 4, ORDER = 3; __VA_ARGS__; } break;                            \
        case 81: { constexpr int K = 8, ORDER = 1; __VA_ARGS__; } break;                            \
        case 82: { constexpr int K = 8, ORDER = 2; __VA_ARGS__; } break;                            \
        case 83: { constexpr int K = 8, ORDER = 3; __VA_ARGS__; } break;                            \
        default: return fail(FN_ERR_UNSUPPORTED, "transform: unsupported m / order");               \
    }

// sums: [Syy, Stt, s[m], A[order][m], B[order][m], D[order][m]]  (fn_transform.cuh); a: order x m
extern "C" int fn_transform_sums(fn_handle* h, int32_t order, const double* a, double* sums) {
    if (!h || h->tr_m == 0) return fail(FN_ERR_STATE, "fn_transform_sums needs fn_transform_upload");
    if (order < 1 || order > TR_MAX_ORDER || !a || !sums) return fail(FN_ERR_UNSUPPORTED, "transform order must be 1..%d", TR_MAX_ORDER);
    const int m = h->tr_m;
    for (int i = 0; i < order * m; ++i) if (!(fabs(a[i]) < INFINITY)) return fail(FN_ERR_ARG, "transform parameters must be finite");
    LEAVE_RESIDENT(h);
    CUDA_TRY(cudaSetDevice(h->device));
    const int nacc = 2 + m * (1 + 3 * order);
    CUDA_TRY(ensure_buf(h->b_ta, (size_t)order * m * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(h->b_ta.p, a, (size_t)order * m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    const int threads = 256, nwarp = threads / 32;
    long long warps_needed = h->tr_n > 0 ? h->tr_n : 1;
    int grid = (int)std::min<long long>((warps_needed + nwarp - 1) / nwarp, (long long)h->sm_count * 2);
    int rc = ensure_reduce(h, grid, nacc);
    if (rc) return rc;
    TransformParams tp;
    tp.n = h->tr_n; tp.m = m; tp.c = h->tr_c; tp.order = order;
    tp.x = (const double*)h->b_tx.p; tp.q = (const double*)h->b_tq.p; tp.a = (const double*)h->b_ta.p;
    fill_reduce(h, tp.gr, tp.sv, nacc, false);
    int k = (m + 31) / 32;
    if (k > 4) k = 8;
    const size_t smem = ((size_t)nacc + 32 * nwarp + 32 * k * TR_MAXC) * sizeof(double);
    FN_DISPATCH_TRANSFORM(k, order, {
        rc = set_smem(h, transform_kernel<K, ORDER>, smem);
        if (rc) return rc;
        transform_kernel<K, ORDER><<<grid, threads, smem, h->stream>>>(tp);
    });
    ++h->launches;
    CUDA_TRY(cudaGetLastError());
    rc = wait_result(h, h->host_seq, nacc);
    if (rc) return rc;
    for (int i = 0; i < nacc; ++i) sums[i] = h->out_vals[i];
    return 0;
}

extern "C" int fn_transform_apply(fn_handle* h, int32_t order, const double* a, const double* col_mean, double* y) {
    if (!h || h->tr_m == 0) return fail(FN_ERR_STATE, "fn_transform_apply needs fn_transform_upload");
    if (order < 1 || order > TR_MAX_ORDER || !a || !col_mean || !y) return fail(FN_ERR_ARG, "fn_transform_apply: bad arguments");
    const int m = h->tr_m;
    LEAVE_RESIDENT(h);
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t bytes = (size_t)h->tr_n * m * sizeof(double);
    int rc = ensure_scratch(h, bytes ? bytes : 16);
    if (rc) return rc;
    CUDA_TRY(ensure_buf(h->b_ta, (size_t)order * m * sizeof(double)));
    CUDA_TRY(ensure_buf(h->b_tmean, (size_t)m * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(h->b_ta.p, a, (size_t)order * m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->b_tmean.p, col_mean, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    transform_apply_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(h->tr_n, m, order, (const double*)h->b_tx.p,
                                                                   (const double*)h->b_ta.p, (const double*)h->b_tmean.p,
                                                                   (double*)h->scratch_dev);
    ++h->launches;
    CUDA_TRY(cudaGetLastError());
    if (bytes) D2H_TRY(y, h->scratch_dev, bytes);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return 0;
}
