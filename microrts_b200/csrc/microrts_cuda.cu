// microrts_cuda.cu -- libmicrorts_cuda.so: the C ABI of include/microrts_cuda.h over the sm_100a engine (engine.cuh).
//
// Build (see __graft_entry__.build): nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared ...
// The same translation unit can be compiled for the host with -DMRTS_EMU by tests/emu/build.sh (test tooling that
// emulates warps with coroutines); that build is never shipped and never loaded by the microrts_b200 package.
#include "../../include/microrts_cuda.h"

#include <algorithm>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#ifdef MRTS_EMU
#include "cuda_shim.h"
#else
#include <cuda_runtime.h>
#include <dlfcn.h>
#endif

#include "engine.cuh"
#include "host_model.hpp"
#include "mcts.hpp"

using namespace mrts;

// ------------------------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }

// ------------------------------------------------------------------------------------------------------------------
// device backend
// ------------------------------------------------------------------------------------------------------------------
#ifdef MRTS_EMU
typedef void *stream_t;
static int dev_select(int) { return 0; }
static int dev_alloc(void **p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? 0 : -1; }
static void dev_free(void *p) { free(p); }
static int dev_h2d(void *d, const void *h, size_t n, stream_t) { memcpy(d, h, n); return 0; }
static int dev_d2h(void *h, const void *d, size_t n, stream_t) { memcpy(h, d, n); return 0; }
static int dev_d2d(void *d, const void *s, size_t n, stream_t) { memcpy(d, s, n); return 0; }
static int dev_zero(void *d, size_t n, stream_t) { memset(d, 0, n); return 0; }
static int dev_sync(stream_t) { return 0; }
static const char *dev_errstr() { return "emulator"; }
#else
typedef cudaStream_t stream_t;
static cudaError_t g_cuda_last = cudaSuccess;
static int ck(cudaError_t e) { if (e != cudaSuccess) { g_cuda_last = e; return -1; } return 0; }
static int dev_select(int d) { return ck(cudaSetDevice(d)); }
static int dev_alloc(void **p, size_t n) { return ck(cudaMalloc(p, n ? n : 1)); }
static void dev_free(void *p) { if (p) cudaFree(p); }
static int dev_h2d(void *d, const void *h, size_t n, stream_t s) { return ck(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); }
static int dev_d2h(void *h, const void *d, size_t n, stream_t s) { if (ck(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s))) return -1; return ck(cudaStreamSynchronize(s)); }
static int dev_d2d(void *d, const void *s_, size_t n, stream_t s) { return ck(cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s)); }
static int dev_zero(void *d, size_t n, stream_t s) { return ck(cudaMemsetAsync(d, 0, n, s)); }
static int dev_sync(stream_t s) { return ck(cudaStreamSynchronize(s)); }
static const char *dev_errstr() { return cudaGetErrorString(g_cuda_last); }
#endif

// ------------------------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------------------------
struct ResetParams {
    int32_t *hdr; uint32_t *units; const uint32_t *maps;
    const long long *seeds; const uint8_t *mask;
    long long n_games; int n_maps, map_words, cap, pcw, uw, keep_rng;
};

// (re)initialise games from their map's initial state; one warp per game
DEV void reset_kernel_body(const ResetParams &p, int tid, int nthreads, int bid, int nblocks) {
    int warp = tid >> 5, lane = tid & 31, wpc = nthreads >> 5;
    for (long long gi = (long long)bid * wpc + warp; gi < p.n_games; gi += (long long)nblocks * wpc) {
        if (p.mask && !p.mask[gi]) continue;
        const uint32_t *blob = p.maps + (size_t)(gi % p.n_maps) * p.map_words;
        const int32_t *ih = (const int32_t *)(blob + p.pcw);
        const uint32_t *iu = blob + p.pcw + MRTS_HDR_WORDS;
        int32_t *gh = p.hdr + gi * MRTS_HDR_WORDS;
        uint32_t *gu = p.units + gi * (long long)p.uw * p.cap;
        int n = ih[H_NUNITS];
        if (lane < MRTS_HDR_WORDS) {
            int32_t v = ih[lane];
            if (p.keep_rng) { // restart: the RNG streams keep running and the episode counter advances (JNIGridnetVecClient auto-reset)
                if (lane >= H_RNGP_LO && lane <= H_RNGD_HI) v = gh[lane];
                if (lane == H_SPARE) v = gh[lane] + 1;
                gh[lane] = v;
            } else {
            long long seed = p.seeds ? p.seeds[gi] : gi;
            unsigned long long sp = ((unsigned long long)seed ^ 0x5DEECE66DULL) & MASK48;
            unsigned long long sc = ((unsigned long long)(seed ^ 0x5851F42D4C957F2DLL) ^ 0x5DEECE66DULL) & MASK48;
            unsigned long long sd = ((unsigned long long)(seed ^ 0x14057B7EF767814FLL) ^ 0x5DEECE66DULL) & MASK48;
            if (lane == H_RNGP_LO) v = (int32_t)(uint32_t)sp;
            if (lane == H_RNGP_HI) v = (int32_t)(uint32_t)(sp >> 32);
            if (lane == H_RNGC_LO) v = (int32_t)(uint32_t)sc;
            if (lane == H_RNGC_HI) v = (int32_t)(uint32_t)(sc >> 32);
            if (lane == H_RNGD_LO) v = (int32_t)(uint32_t)sd;
            if (lane == H_RNGD_HI) v = (int32_t)(uint32_t)(sd >> 32);
            gh[lane] = v;
            }
        }
        for (int k = 0; k < p.uw; k++)
            for (int i = lane; i < n; i += 32) gu[k * p.cap + i] = iu[k * p.cap + i];
    }
}

struct ResultParams { const int32_t *hdr; const uint32_t *units; int32_t *out; long long n_games; int cap, uw; };
// out[g] = {time, winner, gameover, error bits}; one thread per game (PhysicalGameState.winner/gameover :334-387)
DEV void results_kernel_body(const ResultParams &p, long long gi) {
    if (gi >= p.n_games) return;
    const int32_t *h = p.hdr + gi * MRTS_HDR_WORDS;
    const uint32_t *w0 = p.units + gi * (long long)p.uw * p.cap;
    int n = h[H_NUNITS], c0 = 0, c1 = 0;
    for (int i = 0; i < n; i++) { int pl = (w0[i] >> 8) & 0xff; c0 += pl == 1; c1 += pl == 2; }
    int32_t *o = p.out + gi * 4;
    o[0] = h[H_TIME];
    o[1] = (c0 > 0 && c1 == 0) ? 0 : ((c1 > 0 && c0 == 0) ? 1 : -1);
    o[2] = (c0 == 0 || c1 == 0) ? 1 : 0;
    o[3] = h[H_ERR];
}

struct CopyParams {
    int32_t *dhdr; uint32_t *dunits; const int32_t *shdr; const uint32_t *sunits;
    const long long *src_index; const uint8_t *mask;
    long long n_dst, n_src; int cap, uw;
    const long long *dst_index; // scatter form: source game s goes to destination game dst_index[s] (< 0: skipped); src_index / mask unused
};
// GameState.clone() (GameState.java:582-604) for many games at once: game g of the destination batch becomes a copy of game
// src_index[g] (or g) of the source batch, where mask[g] != 0 (or everywhere); one warp per game
DEV void copy_kernel_body(const CopyParams &p, int tid, int nthreads, int bid, int nblocks) {
    int warp = tid >> 5, lane = tid & 31, wpc = nthreads >> 5;
    for (long long it = (long long)bid * wpc + warp; it < (p.dst_index ? p.n_src : p.n_dst); it += (long long)nblocks * wpc) {
        long long gi = it, si = it;
        if (p.dst_index) { gi = p.dst_index[it]; if (gi < 0 || gi >= p.n_dst) continue; }
        else {
            if (p.mask && !p.mask[gi]) continue;
            si = p.src_index ? p.src_index[gi] : gi;
        }
        if (si < 0 || si >= p.n_src) continue;
        const int32_t *sh = p.shdr + si * MRTS_HDR_WORDS;
        const uint32_t *su = p.sunits + si * (long long)p.uw * p.cap;
        uint32_t *du = p.dunits + gi * (long long)p.uw * p.cap;
        if (lane < MRTS_HDR_WORDS) p.dhdr[gi * MRTS_HDR_WORDS + lane] = sh[lane];
        int n = sh[H_NUNITS];
        for (int k = 0; k < p.uw; k++)
            for (int i = lane; i < n; i += 32) du[k * p.cap + i] = su[k * p.cap + i];
    }
}

struct ObsParams {
    const int32_t *hdr; const uint32_t *units; const uint32_t *maps; void *out;
    long long n_games; int n_maps, map_words, W, H, cap, uw, player, dtype;
};
// GameState.getVectorObservation for every game straight from the state in HBM (fully observable batches): one warp per
// game, no shared memory; reads 3 words per unit, writes the 6 planes (obs_emit)
DEV void observe_kernel_body(const ObsParams &p, int tid, int nthreads, int bid, int nblocks) {
    int warp = tid >> 5, lane = tid & 31, wpc = nthreads >> 5;
    for (long long gi = (long long)bid * wpc + warp; gi < p.n_games; gi += (long long)nblocks * wpc) {
        const uint32_t *blob = p.maps + (size_t)(gi % p.n_maps) * p.map_words;
        const uint32_t *un = p.units + gi * (long long)p.uw * p.cap;
        int n = p.hdr[gi * MRTS_HDR_WORDS + H_NUNITS];
        obs_emit(un + UW_W0 * p.cap, un + UW_W1 * p.cap, un + UW_A0 * p.cap, n, p.W, p.H, map_terrain(blob, p.W, p.H, p.cap), p.player, p.dtype,
                 (char *)p.out + (size_t)gi * obs_bytes_per_game(p.W, p.H, 6, p.dtype), lane);
    }
}

#ifndef MRTS_EMU
#ifndef MRTS_MIN_BLOCKS
#define MRTS_MIN_BLOCKS 7     // k_step_fast: 72 registers, 28 warps per SM on the 16x16 configuration (8 CTAs at 64 registers spill: measured 3% slower)
#endif
#ifndef MRTS_MIN_BLOCKS_OBS
#define MRTS_MIN_BLOCKS_OBS 4 // the observation variant runs on large maps, where shared memory bounds the occupancy anyway
#endif
#ifndef MRTS_MIN_BLOCKS_ROLLOUT
#define MRTS_MIN_BLOCKS_ROLLOUT 5
#endif
// k_step_fast: Game.start loop with RandomBiasedAI / PassiveAI under CANCEL_BOTH (the benchmark path); k_rollout:
// NaiveMCTS.simulate + evaluation; k_step: every other mode.  All three are persistent, one warp per game at a time.
__global__ void __launch_bounds__(MRTS_WARPS_PER_CTA * 32, MRTS_MIN_BLOCKS) k_step_fast(StepParams p) {
    step_kernel_body<KERNEL_FAST>(p, mrts_smem, threadIdx.x, blockDim.x, blockIdx.x, gridDim.x);
}
// the fixed-size copies (engine.cuh: MRTS_FIXED_VARIANTS)
template <int KERNEL, int W, int H, int CAP, int MINB>
__global__ void __launch_bounds__(MRTS_WARPS_PER_CTA * 32, MINB) k_fixed(StepParams p) {
    step_kernel_body<KERNEL, W, H, CAP>(p, mrts_smem, threadIdx.x, blockDim.x, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(MRTS_WARPS_PER_CTA * 32, MRTS_MIN_BLOCKS_OBS) k_step_fast_obs(StepParams p) {
    step_kernel_body<KERNEL_FAST_OBS>(p, mrts_smem, threadIdx.x, blockDim.x, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(MRTS_WARPS_PER_CTA * 32, MRTS_MIN_BLOCKS_ROLLOUT) k_rollout(StepParams p) {
    step_kernel_body<KERNEL_ROLLOUT>(p, mrts_smem, threadIdx.x, blockDim.x, blockIdx.x, gridDim.x);
}
#ifndef MRTS_MIN_BLOCKS_GENERIC
#define MRTS_MIN_BLOCKS_GENERIC 3
#endif
__global__ void __launch_bounds__(MRTS_WARPS_PER_CTA * 32, MRTS_MIN_BLOCKS_GENERIC) k_step(StepParams p) {
    step_kernel_body<KERNEL_GENERIC>(p, mrts_smem, threadIdx.x, blockDim.x, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(256) k_observe(ObsParams p) { observe_kernel_body(p, threadIdx.x, blockDim.x, blockIdx.x, gridDim.x); }
__global__ void __launch_bounds__(128) k_copy_games(CopyParams p) { copy_kernel_body(p, threadIdx.x, blockDim.x, blockIdx.x, gridDim.x); }
__global__ void __launch_bounds__(128) k_reset(ResetParams p) { reset_kernel_body(p, threadIdx.x, blockDim.x, blockIdx.x, gridDim.x); }
__global__ void __launch_bounds__(128) k_results(ResultParams p) { results_kernel_body(p, (long long)blockIdx.x * blockDim.x + threadIdx.x); }
#endif

// ------------------------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------------------------
struct mrts_utt { UttH h; };
struct mrts_map { MapH h; };

struct Staged { // external actions staged on the device
    int32_t *actions = nullptr, *counts = nullptr; // buffers this player owns
    const int32_t *rows = nullptr;                 // what the next step reads: `actions`, or the other player's buffer (interleaved staging)
    size_t cap_rows = 0;
    int max_k = 0, format = 0, fill = -1;
    long long stride = 0;                          // int32 elements between consecutive games' rows
    bool valid = false, has_counts = false;
};

// one fixed-size copy: which kernel it stands in for, for which map size / capacity, and how to launch it
struct FixedVariant {
    int kernel, W, H, cap;
#ifdef MRTS_EMU
    void (*body)(const StepParams &, unsigned char *, int, int, int, int);
#else
    const void *fn;
#endif
};
#ifdef MRTS_EMU
#define MRTS_FV_ENTRY(K, W, H, C, MINB) {K, W, H, C, &step_kernel_body<K, W, H, C>},
#else
#define MRTS_FV_ENTRY(K, W, H, C, MINB) {K, W, H, C, (const void *)k_fixed<K, W, H, C, MINB>},
#endif
static const FixedVariant g_fixed[] = {MRTS_FIXED_VARIANTS(MRTS_FV_ENTRY)};
#ifndef MRTS_EMU
// the generic kernel compiled for one layout each in translation units of their own (fixed_<W>x<H>.cu, fixed_generic.inc)
typedef const void *(*fixed_generic_getter)(int *W, int *H, int *cap, int *rush_only);
#define MRTS_FG_DECL(name) extern "C" __attribute__((visibility("hidden"))) const void *mrts_fixed_generic_##name(int *W, int *H, int *cap, int *rush_only);
MRTS_FG_DECL(24x24) MRTS_FG_DECL(16x16) MRTS_FG_DECL(8x8) MRTS_FG_DECL(24x24_rush) MRTS_FG_DECL(16x16_rush)
static const fixed_generic_getter g_fixed_generic[] = {mrts_fixed_generic_24x24, mrts_fixed_generic_16x16, mrts_fixed_generic_8x8,
                                                        mrts_fixed_generic_24x24_rush, mrts_fixed_generic_16x16_rush};
#endif
static const int N_FIXED = (int)(sizeof(g_fixed) / sizeof(g_fixed[0]));

struct mrts_batch {
    UttH utt;
    std::vector<MapH> maps_h; // the maps the batch was created from (a search clones the configuration for its node pool)
    int W = 0, H = 0, cap = 0, n_maps = 0, map_words = 0, device = 0;
    long long n = 0;
    uint32_t flags = 0;
    int policy[2] = {MRTS_POLICY_PASSIVE, MRTS_POLICY_PASSIVE}, pathfinder[2] = {0, 0};
    int32_t *d_hdr = nullptr; uint32_t *d_units = nullptr; uint32_t *d_maps = nullptr; uint32_t *d_cst = nullptr;
    unsigned long long *d_stats = nullptr;
    int32_t *d_results = nullptr; bool results_fresh = false; // [n][4] written by the last mrts_batch_step; stale after any other change of state
    void *d_tmp = nullptr; size_t tmp_bytes = 0; // staging for host arguments
    unsigned char *d_astar = nullptr; long long astar_stride = 0; // pathfinding scratch of large-map scripted batches
    unsigned char *d_ff = nullptr; long long ff_stride = 0;       // FloodFillPathFinding caches, allocated when a player asks for that pathfinder
    Staged staged[2];
    stream_t stream = nullptr;
    SmemLayout L, Lfast, Lslim; // generic kernel / specialised kernels (no pending lists, layout.h) / the fused step + observation kernel (no kind and claim maps either)
    struct Plan { int wpc = 2, grid = 3; size_t smem = 0; } plan[N_KERNELS]; // per kernel: warps (games in flight) per CTA, CTAs, shared memory
    Plan fixed_plan[N_KERNELS]; int fixed_of[N_KERNELS] = {-1, -1, -1, -1}; // the fixed-size copy that replaces kernel k for this batch, or -1
    const void *generic_fixed_fn = nullptr; // a fixed_<W>x<H>.cu kernel when the batch has its layout (plan in fixed_plan[KERNEL_GENERIC])
    const void *generic_rush_fn = nullptr; Plan rush_plan; // its rush-only copy (fixed_<W>x<H>_rush.cu), for steps whose policies allow it
    int max_range = 0, auto_reset = 0, scripted = 0, uw = MRTS_UNIT_WORDS_CORE;
    int sequential_issue = 0; int32_t *info_out = nullptr; uint32_t tm[6] = {0, 0, 0, 0, 0, 0};
    void *obs_out[2] = {nullptr, nullptr}; int obs_dtype = 0; // device buffers mrts_batch_step writes post-step observations to
    void *mask_out[2] = {nullptr, nullptr};                   // ... and the post-step bit-packed action masks
    int zero_bytes = 0;                                       // block of zeros per CTA of k_step_fast_obs (source of its bulk stores)
    int terr_bytes = 0, tmpl_bytes = 0;                       // ... and its staged terrain plane / grid template (single-map batches)
    int vec_reset = 0, vec_max_steps = 0;                     // in-kernel auto-reset of the JNIGridnetVecClient flow (mrts_batch_set_vec_autoreset)
    int out_stride = 1;                                       // game g's fused outputs go to game slot g * out_stride (mrts_batch_set_output_stride)
    long long launches = 0;
    long long host_io[2] = {0, 0}; // bytes moved by the kernels that carry no counters (k_observe), added to mrts_batch_io_bytes
    char last_kernel[64] = "";     // symbol of the most recent step-kernel launch (mrts_batch_last_kernel)
};

static int ensure_tmp(mrts_batch *b, size_t bytes) {
    if (bytes <= b->tmp_bytes) return 0;
    dev_free(b->d_tmp); b->d_tmp = nullptr; b->tmp_bytes = 0;
    if (dev_alloc(&b->d_tmp, bytes)) return -1;
    b->tmp_bytes = bytes;
    return 0;
}

static int launch_step(mrts_batch *b, StepParams &p) {
    p.hdr = b->d_hdr; p.units = b->d_units; p.maps = b->d_maps; p.cst = b->d_cst; p.stats = b->d_stats;
    p.work_counter = b->d_stats + 8; // games are handed out dynamically; every launch leaves the two counters at zero (step_kernel_body)
    p.n_games = b->n; p.n_maps = b->n_maps; p.map_words = b->map_words; p.W = b->W; p.H = b->H; p.cap = b->cap;
    p.conflict = b->utt.conflict; p.max_range = b->max_range; p.n_types = (int)b->utt.types.size();
    p.partial_obs = (b->flags & MRTS_FLAG_PARTIAL_OBS) ? 1 : 0; p.scripted = b->scripted; p.uw = b->uw;
    p.astar_scratch = b->d_astar; p.astar_stride = b->astar_stride; p.ff_cache = b->d_ff; p.ff_stride = b->ff_stride;
    p.po_policies = (b->flags & MRTS_FLAG_PO_POLICIES) ? 1 : 0;
    // kernel selection: the specialised kernels cover exactly the cases their loops implement
    int kernel = KERNEL_GENERIC;
    auto rb_or_passive = [](int pol) { return pol == MRTS_POLICY_RANDOM_BIASED || pol == MRTS_POLICY_PASSIVE; };
    if (p.mode == MODE_ROLLOUT) kernel = KERNEL_ROLLOUT;
    else if (p.mode == MODE_GAME && p.conflict == MRTS_CANCEL_BOTH && rb_or_passive(p.policy[0]) && rb_or_passive(p.policy[1]) && !p.info_out && !p.sequential_issue && !p.po_policies && !p.vec_reset)
        kernel = (p.obs_out[0] || p.obs_out[1] || p.mask_out[0] || p.mask_out[1]) ? KERNEL_FAST_OBS : KERNEL_FAST;
    p.zero_bytes = kernel == KERNEL_FAST_OBS ? b->zero_bytes : 0;
    p.terr_bytes = kernel == KERNEL_FAST_OBS ? b->terr_bytes : 0; p.tmpl_bytes = kernel == KERNEL_FAST_OBS ? b->tmpl_bytes : 0;
    const int fv = b->fixed_of[kernel];
    const bool gfx = kernel == KERNEL_GENERIC && b->generic_fixed_fn;
    // the rush-only copy: a game step in which neither player runs a defense, WorkerRushPlusPlus or a PO rush
    // the lean rush-only copy: a game step in which both players run a scripted rush with A* under CANCEL_BOTH, nothing else asked for
    auto is_rush = [](int pol) { return pol >= MRTS_POLICY_WORKER_RUSH && pol <= MRTS_POLICY_RANGED_RUSH; };
    const bool rush = gfx && b->generic_rush_fn && p.mode == MODE_GAME && is_rush(p.policy[0]) && is_rush(p.policy[1]) &&
                      p.pathfinder[0] == MRTS_PF_ASTAR && p.pathfinder[1] == MRTS_PF_ASTAR && p.conflict == MRTS_CANCEL_BOTH && !p.info_out &&
                      !p.sequential_issue && !p.vec_reset && !p.obs_out[0] && !p.obs_out[1] && !p.mask_out[0] && !p.mask_out[1];
    const mrts_batch::Plan &pl = rush ? b->rush_plan : ((fv >= 0 || gfx) ? b->fixed_plan[kernel] : b->plan[kernel]);
    p.L = kernel == KERNEL_GENERIC ? b->L : (kernel == KERNEL_FAST_OBS ? b->Lslim : b->Lfast);
    int threads = pl.wpc * 32;
    long long items = p.mode == MODE_ROLLOUT ? b->n * p.rollouts_per_game : b->n;
    long long need = (items + pl.wpc - 1) / pl.wpc;
    int grid = (int)std::min<long long>(pl.grid, std::max<long long>(need, 1));
    b->launches++;
    {
        static const char *names[N_KERNELS] = {"k_step_fast", "k_rollout", "k_step", "k_step_fast_obs"};
        if (rush) snprintf(b->last_kernel, sizeof b->last_kernel, "k_step_fixed_%dx%d_rush", b->W, b->H);
        else if (gfx) snprintf(b->last_kernel, sizeof b->last_kernel, "k_step_fixed_%dx%d", b->W, b->H);
        else if (fv >= 0) snprintf(b->last_kernel, sizeof b->last_kernel, "k_fixed<%s,%d,%d,%d>", names[kernel], b->W, b->H, b->cap);
        else snprintf(b->last_kernel, sizeof b->last_kernel, "%s", names[kernel]);
    }
#ifdef MRTS_EMU
    StepParams pc = p;
    emu::launch(grid, threads, pl.smem, [pc, threads, grid, kernel, fv](unsigned char *sm, int tid, int bid) {
        if (fv >= 0) g_fixed[fv].body(pc, sm, tid, threads, bid, grid);
        else if (kernel == KERNEL_FAST) step_kernel_body<KERNEL_FAST>(pc, sm, tid, threads, bid, grid);
        else if (kernel == KERNEL_FAST_OBS) step_kernel_body<KERNEL_FAST_OBS>(pc, sm, tid, threads, bid, grid);
        else if (kernel == KERNEL_ROLLOUT) step_kernel_body<KERNEL_ROLLOUT>(pc, sm, tid, threads, bid, grid);
        else step_kernel_body<KERNEL_GENERIC>(pc, sm, tid, threads, bid, grid);
    });
    return 0;
#else
    if (fv >= 0 || gfx) { void *args[] = {&p}; return ck(cudaLaunchKernel(rush ? b->generic_rush_fn : (gfx ? b->generic_fixed_fn : g_fixed[fv].fn), dim3(grid), dim3(threads), args, pl.smem, b->stream)); }
    if (kernel == KERNEL_FAST) k_step_fast<<<grid, threads, pl.smem, b->stream>>>(p);
    else if (kernel == KERNEL_FAST_OBS) {
        // With the action masks as well a game writes 128 KB (64x64): measured, ONE resident CTA (four games) per SM then moves more
        // bytes per second than two (5.9 against 4.9 TB/s) -- fewer write streams in flight at once.  The kernel's shared memory
        // request is padded so that no second CTA fits.  MRTS_DBG_OBS_CTAS overrides the cap (profiling experiments).
        size_t sm = pl.smem; int gr = grid, cap_ctas = (p.mask_out[0] || p.mask_out[1]) && b->W * b->H >= 1024 ? 1 : 0;
        if (const char *e = getenv("MRTS_DBG_OBS_CTAS")) cap_ctas = atoi(e);
        if (cap_ctas > 0) { sm = std::max(sm, (size_t)(233472 / (cap_ctas + 1) - 1024 + 16) & ~(size_t)15); gr = std::min(grid, cap_ctas * 148); }
        k_step_fast_obs<<<gr, threads, sm, b->stream>>>(p);
    }
    else if (kernel == KERNEL_ROLLOUT) k_rollout<<<grid, threads, pl.smem, b->stream>>>(p);
    else k_step<<<grid, threads, pl.smem, b->stream>>>(p);
    return ck(cudaGetLastError());
#endif
}

static int launch_observe(mrts_batch *b, int player, int dtype, void *d_out) {
    ObsParams p{b->d_hdr, b->d_units, b->d_maps, d_out, b->n, b->n_maps, b->map_words, b->W, b->H, b->cap, b->uw, player, dtype};
    b->launches++;
    b->host_io[1] += (long long)b->n * 6 * b->W * b->H * (dtype == MRTS_DTYPE_U8 ? 1 : 4); // (its reads, 12 bytes per live unit, are not counted)
#ifdef MRTS_EMU
    emu::launch(2, 64, 0, [p](unsigned char *, int tid, int bid) { observe_kernel_body(p, tid, 64, bid, 2); });
    return 0;
#else
    int grid = (int)std::min<long long>((b->n + 7) / 8, 148 * 8);
    k_observe<<<grid, 256, 0, b->stream>>>(p);
    return ck(cudaGetLastError());
#endif
}

// the searches are independent: their host work (tree descent, back-propagation, decoding of the device's lists) is spread over the
// host's cores
template <class F> static void parallel_for(int n, F fn) {
    int nt = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    if (n < 128 || nt < 2) { for (int i = 0; i < n; i++) fn(i); return; }
    std::vector<std::thread> th;
    int chunk = (n + nt - 1) / nt;
    for (int t = 0; t < nt; t++) {
        int lo = t * chunk, hi = std::min(n, lo + chunk);
        if (lo >= hi) break;
        th.emplace_back([lo, hi, &fn]() { for (int i = lo; i < hi; i++) fn(i); });
    }
    for (auto &x : th) x.join();
}

// ------------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------------
extern "C" {

int mrts_abi_version(void) { return MRTS_ABI_VERSION; }
const char *mrts_last_error(void) { return g_err.c_str(); }

int mrts_utt_create(int version, int conflict_policy, mrts_utt **out) {
    if (!out || version < 1 || version > 3 || conflict_policy < 1 || conflict_policy > 3) return fail(MRTS_E_ARG, "mrts_utt_create: version must be 1..3 and conflict policy 1..3");
    auto *u = new mrts_utt; u->h = make_utt(version, conflict_policy); *out = u; return MRTS_OK;
}
int mrts_utt_from_json(const char *json, mrts_utt **out) {
    if (!json || !out) return fail(MRTS_E_ARG, "mrts_utt_from_json: null argument");
    auto u = std::make_unique<mrts_utt>(); std::string err;
    if (!utt_from_json(json, u->h, err)) return fail(MRTS_E_PARSE, err);
    if (!utt_check_limits(u->h, err)) return fail(MRTS_E_LIMIT, err);
    *out = u.release(); return MRTS_OK;
}
int mrts_utt_num_types(const mrts_utt *u) { return u ? (int)u->h.types.size() : MRTS_E_ARG; }
int mrts_utt_get(const mrts_utt *u, int t, int f) {
    if (!u || t < 0 || t >= (int)u->h.types.size()) return fail(MRTS_E_ARG, "mrts_utt_get: bad type id");
    const UnitTypeH &x = u->h.types[t];
    switch (f) {
        case MRTS_UT_COST: return x.cost; case MRTS_UT_HP: return x.hp; case MRTS_UT_MIN_DAMAGE: return x.minDamage;
        case MRTS_UT_MAX_DAMAGE: return x.maxDamage; case MRTS_UT_ATTACK_RANGE: return x.attackRange;
        case MRTS_UT_PRODUCE_TIME: return x.produceTime; case MRTS_UT_MOVE_TIME: return x.moveTime;
        case MRTS_UT_ATTACK_TIME: return x.attackTime; case MRTS_UT_HARVEST_TIME: return x.harvestTime;
        case MRTS_UT_RETURN_TIME: return x.returnTime; case MRTS_UT_HARVEST_AMOUNT: return x.harvestAmount;
        case MRTS_UT_SIGHT_RADIUS: return x.sightRadius; case MRTS_UT_FLAGS: return x.flags();
        case MRTS_UT_N_PRODUCES: return (int)x.produces.size();
    }
    if (f >= MRTS_UT_PRODUCES0 && f < MRTS_UT_PRODUCES0 + (int)x.produces.size()) return x.produces[f - MRTS_UT_PRODUCES0];
    return fail(MRTS_E_ARG, "mrts_utt_get: bad field");
}
const char *mrts_utt_type_name(const mrts_utt *u, int t) { return (u && t >= 0 && t < (int)u->h.types.size()) ? u->h.types[t].name.c_str() : nullptr; }
int mrts_utt_conflict_policy(const mrts_utt *u) { return u ? u->h.conflict : MRTS_E_ARG; }
int mrts_utt_max_attack_range(const mrts_utt *u) { return u ? u->h.maxAttackRange() : MRTS_E_ARG; }
void mrts_utt_destroy(mrts_utt *u) { delete u; }

int mrts_map_from_xml(const char *xml, const mrts_utt *u, mrts_map **out) {
    if (!xml || !u || !out) return fail(MRTS_E_ARG, "mrts_map_from_xml: null argument");
    auto m = std::make_unique<mrts_map>(); std::string err;
    if (!map_from_xml(xml, u->h, m->h, err)) return fail(MRTS_E_PARSE, err);
    if (!map_check(m->h, u->h, err)) return fail(MRTS_E_LIMIT, err);
    *out = m.release(); return MRTS_OK;
}
int mrts_map_load_xml(const char *path, const mrts_utt *u, mrts_map **out) {
    if (!path || !u || !out) return fail(MRTS_E_ARG, "mrts_map_load_xml: null argument");
    std::ifstream f(path, std::ios::binary);
    if (!f) return fail(MRTS_E_IO, std::string("cannot open map file: ") + path);
    std::stringstream ss; ss << f.rdbuf();
    return mrts_map_from_xml(ss.str().c_str(), u, out);
}
int mrts_map_create(int w, int h, const uint8_t *terrain, int res0, int res1, int n_units, const int32_t *units, const mrts_utt *u, mrts_map **out) {
    if (!u || !out || w <= 0 || h <= 0 || n_units < 0 || (n_units && !units)) return fail(MRTS_E_ARG, "mrts_map_create: bad argument");
    auto m = std::make_unique<mrts_map>();
    m->h.w = w; m->h.h = h; m->h.res[0] = res0; m->h.res[1] = res1;
    m->h.terrain.assign((size_t)w * h, 0);
    if (terrain) for (size_t i = 0; i < (size_t)w * h; i++) m->h.terrain[i] = terrain[i] ? 1 : 0;
    for (int i = 0; i < n_units; i++) { const int32_t *r = units + i * 7; m->h.units.push_back(MapUnit{r[0], r[1], r[2], r[3], r[4], r[5], r[6]}); }
    std::string err;
    if (!map_check(m->h, u->h, err)) return fail(MRTS_E_LIMIT, err);
    *out = m.release(); return MRTS_OK;
}
int mrts_map_width(const mrts_map *m) { return m ? m->h.w : MRTS_E_ARG; }
int mrts_map_height(const mrts_map *m) { return m ? m->h.h : MRTS_E_ARG; }
int mrts_map_num_units(const mrts_map *m) { return m ? (int)m->h.units.size() : MRTS_E_ARG; }
int mrts_map_get_units(const mrts_map *m, int32_t *out) {
    if (!m || !out) return fail(MRTS_E_ARG, "mrts_map_get_units: null argument");
    for (size_t i = 0; i < m->h.units.size(); i++) { const MapUnit &u = m->h.units[i]; int32_t *r = out + i * 7; r[0] = u.type; r[1] = (int32_t)u.id; r[2] = u.player; r[3] = u.x; r[4] = u.y; r[5] = u.res; r[6] = u.hp; }
    return (int)m->h.units.size();
}
int mrts_map_get_terrain(const mrts_map *m, uint8_t *out) { if (!m || !out) return fail(MRTS_E_ARG, "mrts_map_get_terrain: null argument"); memcpy(out, m->h.terrain.data(), m->h.terrain.size()); return MRTS_OK; }
int mrts_map_resources(const mrts_map *m, int p) { return (m && p >= 0 && p < 2) ? m->h.res[p] : MRTS_E_ARG; }
void mrts_map_destroy(mrts_map *m) { delete m; }

void mrts_batch_destroy(mrts_batch *b) {
    if (!b) return;
    dev_select(b->device);
    dev_free(b->d_hdr); dev_free(b->d_units); dev_free(b->d_maps); dev_free(b->d_cst); dev_free(b->d_stats); dev_free(b->d_results); dev_free(b->d_tmp); dev_free(b->d_astar); dev_free(b->d_ff);
    for (auto &s : b->staged) { dev_free(s.actions); dev_free(s.counts); }
#ifndef MRTS_EMU
    if (b->stream) cudaStreamDestroy(b->stream);
#endif
    delete b;
}

int mrts_batch_create(const mrts_utt *u, const mrts_map *const *maps, int n_maps, int64_t n_games, int device, uint32_t flags, int unit_capacity, mrts_batch **out) {
    if (!u || !maps || n_maps < 1 || n_games < 1 || !out) return fail(MRTS_E_ARG, "mrts_batch_create: bad argument");
    std::string err;
    if (!utt_check_limits(u->h, err)) return fail(MRTS_E_LIMIT, err);
    if (!maps[0]) return fail(MRTS_E_ARG, "mrts_batch_create: null map handle");
    if (flags & MRTS_FLAG_SCRIPTED_AI) {
        // the scripted policies address the standard unit types by id (scripted.cuh: UT_BASE ...); the reference looks them up by
        // name (utt.getUnitType("Worker"), WorkerRush.java:52-55), so a table that orders or names them differently is refused
        static const char *std_names[] = {"Resource", "Base", "Barracks", "Worker", "Light", "Heavy", "Ranged"};
        for (size_t t = 0; t < 7; t++)
            if (t >= u->h.types.size() || u->h.types[t].name != std_names[t])
                return fail(MRTS_E_STATE, std::string("MRTS_FLAG_SCRIPTED_AI needs the standard unit types at ids 0..6 (Resource, Base, Barracks, Worker, Light, Heavy, Ranged); type ") + std::to_string(t) + " differs");
    }
    int W = maps[0]->h.w, H = maps[0]->h.h, bound = 0;
    for (int i = 0; i < n_maps; i++) {
        if (!maps[i] || maps[i]->h.w != W || maps[i]->h.h != H) return fail(MRTS_E_ARG, "mrts_batch_create: all maps of a batch must have the same size");
        if (!map_check(maps[i]->h, u->h, err)) return fail(MRTS_E_LIMIT, err);
        bound = std::max(bound, map_unit_bound(maps[i]->h));
        bound = std::max(bound, (int)maps[i]->h.units.size());
    }
    int cap = unit_capacity > 0 ? ((unit_capacity + 3) & ~3) : std::min(MRTS_MAX_CAP, (bound + 31) & ~31); // multiple of 4: 16-byte aligned unit word arrays
    if (cap > MRTS_MAX_CAP) return fail(MRTS_E_LIMIT, "unit capacity above 252");
    for (int i = 0; i < n_maps; i++) if ((int)maps[i]->h.units.size() > (unit_capacity > 0 ? unit_capacity : cap)) return fail(MRTS_E_LIMIT, "map has more initial units than the unit capacity");
    if (cap < 32) cap = 32;
    auto b = std::unique_ptr<mrts_batch, void (*)(mrts_batch *)>(new mrts_batch, mrts_batch_destroy);
    b->utt = u->h; b->W = W; b->H = H; b->cap = cap; b->n_maps = n_maps; b->n = n_games; b->flags = flags; b->device = device;
    b->max_range = u->h.maxAttackRange();
    for (int i = 0; i < n_maps; i++) b->maps_h.push_back(maps[i]->h);
    for (size_t t = 0; t < u->h.types.size() && t < 32; t++) { // the reward functions identify unit types by name (src/ai/reward/*.java)
        const std::string &nm = u->h.types[t].name; uint32_t bit = 1u << t;
        if (nm == "Worker") { b->tm[0] |= bit; b->tm[4] |= bit; }
        if (nm == "Barracks" || nm == "Base") b->tm[1] |= bit;
        if (nm == "Light" || nm == "Heavy" || nm == "Ranged") { b->tm[2] |= bit; b->tm[4] |= bit; }
        if (nm == "Base") b->tm[3] |= bit;
        if (nm == "Resource") b->tm[5] |= bit;
    }
    // scripted batches keep the pathfinding scratch in shared memory while a game's whole region stays small enough for
    // several games per SM; larger maps move it to a per-warp global scratch (L1/L2 resident)
    b->scripted = (flags & MRTS_FLAG_SCRIPTED_AI) ? 1 : 0;
    const int po_pol = (flags & MRTS_FLAG_PO_POLICIES) ? 1 : 0;
    if (po_pol && u->h.types.size() >= MRTS_MAX_TYPES) return fail(MRTS_E_LIMIT, "MRTS_FLAG_PO_POLICIES needs a unit type table with at most 7 types (type 7 marks hidden units)");
    if (b->scripted && mrts_smem_layout(W, H, cap, 1, po_pol).total > 48 * 1024) b->scripted = 2;
    b->uw = b->scripted ? MRTS_UNIT_WORDS : MRTS_UNIT_WORDS_CORE;
    b->L = mrts_smem_layout(W, H, cap, b->scripted, po_pol);
    b->Lfast = mrts_smem_layout(W, H, cap, b->scripted ? 2 : 0, 0, 0); // same unit words, no pathfinding scratch, no pending lists
    b->Lslim = mrts_smem_layout(W, H, cap, b->scripted ? 2 : 0, 0, 0, 1);
    b->map_words = mrts_map_blob_words(W, H, cap);
    if (dev_select(device)) return fail(MRTS_E_CUDA, std::string("cannot select CUDA device: ") + dev_errstr());
#ifndef MRTS_EMU
    cudaDeviceProp prop;
    if (ck(cudaGetDeviceProperties(&prop, device))) return fail(MRTS_E_CUDA, std::string("cudaGetDeviceProperties: ") + dev_errstr());
    // Per kernel: as many games in flight per SM as shared memory and registers allow (large maps need fewer, fatter
    // CTAs); grid = resident CTAs per SM x SMs (persistent kernel).
    const void *kernels[N_KERNELS] = {(const void *)k_step_fast, (const void *)k_rollout, (const void *)k_step, (const void *)k_step_fast_obs};
    auto make_plan = [&](const void *fn, int region, mrts_batch::Plan &out, int extra = 0) -> int {
        int best_wpc = 0, best_warps = 0, best_blocks = 0;
        if (ck(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin)) ||
            ck(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)))
            return fail(MRTS_E_CUDA, std::string("cudaFuncSetAttribute: ") + dev_errstr());
        for (int wpc = MRTS_WARPS_PER_CTA; wpc >= 1; wpc--) {
            size_t sm = MRTS_CONST_WORDS * 4 + (size_t)wpc * region + extra;
            if (sm > (size_t)prop.sharedMemPerBlockOptin) continue;
            int per_sm = 0;
            if (ck(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, wpc * 32, sm))) return fail(MRTS_E_CUDA, std::string("occupancy query: ") + dev_errstr());
            if (per_sm * wpc > best_warps) { best_warps = per_sm * wpc; best_wpc = wpc; best_blocks = per_sm; }
        }
        if (!best_wpc) return fail(MRTS_E_LIMIT, "map too large for the shared-memory resident engine");
        out.wpc = best_wpc;
        out.smem = MRTS_CONST_WORDS * 4 + (size_t)best_wpc * region + extra;
        out.grid = best_blocks * prop.multiProcessorCount;
        return 0;
    };
    // the fused step + observation kernel keeps one block of zeros per CTA: the source of the bulk stores that zero its outputs
    // (measured on config 5: two CTAs of four games per SM with 20 KB bulk stores beat three CTAs with 4 KB ones -- the write streams of
    // too many games in flight get in each other's way -- so the kernel spends shared memory on large sources instead of on occupancy)
    { int z5 = (5 * W * H) & ~15; b->zero_bytes = z5 >= 256 ? std::min(32768, z5) : 0; }
    if (n_maps == 1 && b->zero_bytes) {
        b->tmpl_bytes = b->Lslim.pcw * 4;
        if (((W * H) & 15) == 0 && W * H <= 16384) b->terr_bytes = W * H;
    }
    for (int kk = 0; kk < N_KERNELS; kk++) {
        int rc = make_plan(kernels[kk], kk == KERNEL_GENERIC ? b->L.total : (kk == KERNEL_FAST_OBS ? b->Lslim.total : b->Lfast.total), b->plan[kk], kk == KERNEL_FAST_OBS ? b->zero_bytes + b->terr_bytes + b->tmpl_bytes : 0);
        if (rc) return rc;
    }
#ifndef MRTS_NO_FIXED
    // a fixed-size copy stands in when the batch has exactly its map size and capacity and no scripted-policy unit words
    for (int v = 0; v < N_FIXED; v++)
        if (g_fixed[v].W == W && g_fixed[v].H == H && g_fixed[v].cap == cap && !b->scripted) {
            int rc = make_plan(g_fixed[v].fn, b->Lfast.total, b->fixed_plan[g_fixed[v].kernel]);
            if (rc) return rc;
            b->fixed_of[g_fixed[v].kernel] = v;
        }
    for (fixed_generic_getter get : g_fixed_generic) {
        int fw = 0, fh = 0, fcap = 0, rush_only = 0;
        const void *fn = get(&fw, &fh, &fcap, &rush_only);
        if (fw == W && fh == H && fcap == cap && b->scripted == 1 && !po_pol) {
            int rc = make_plan(fn, b->L.total, rush_only ? b->rush_plan : b->fixed_plan[KERNEL_GENERIC]);
            if (rc) return rc;
            (rush_only ? b->generic_rush_fn : b->generic_fixed_fn) = fn;
        }
    }
#endif
    if (ck(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking))) return fail(MRTS_E_CUDA, std::string("cudaStreamCreate: ") + dev_errstr());
#else
    for (int kk = 0; kk < N_KERNELS; kk++) { b->plan[kk].wpc = 2; b->plan[kk].smem = MRTS_CONST_WORDS * 4 + (size_t)2 * (kk == KERNEL_GENERIC ? b->L.total : (kk == KERNEL_FAST_OBS ? b->Lslim.total : b->Lfast.total)) + (kk == KERNEL_FAST_OBS ? b->zero_bytes : 0); b->plan[kk].grid = 3; }
    for (int v = 0; v < N_FIXED; v++)
        if (g_fixed[v].W == W && g_fixed[v].H == H && g_fixed[v].cap == cap && !b->scripted) { b->fixed_plan[g_fixed[v].kernel] = b->plan[g_fixed[v].kernel]; b->fixed_of[g_fixed[v].kernel] = v; }
#endif
    if (b->scripted == 2) {
        b->astar_stride = ((long long)MRTS_ASTAR_BYTES(W, H) + 255) & ~255LL;
        if (dev_alloc((void **)&b->d_astar, (size_t)b->plan[KERNEL_GENERIC].grid * b->plan[KERNEL_GENERIC].wpc * b->astar_stride)) return fail(MRTS_E_CUDA, std::string("device allocation failed: ") + dev_errstr());
    }
    size_t hdr_bytes = (size_t)n_games * MRTS_HDR_WORDS * 4, unit_bytes = (size_t)n_games * b->uw * cap * 4;
    if (dev_alloc((void **)&b->d_hdr, hdr_bytes) || dev_alloc((void **)&b->d_units, unit_bytes) ||
        dev_alloc((void **)&b->d_maps, (size_t)n_maps * b->map_words * 4) || dev_alloc((void **)&b->d_cst, MRTS_CONST_WORDS * 4) ||
        dev_alloc((void **)&b->d_stats, 20 * sizeof(unsigned long long)) || dev_alloc((void **)&b->d_results, (size_t)n_games * 16)) // 8 counters + the launches' work and exit counters + 8 words of all-reduce scratch + 2 global-memory byte counters
        return fail(MRTS_E_CUDA, std::string("device allocation failed: ") + dev_errstr());
    std::vector<uint32_t> blob, all;
    for (int i = 0; i < n_maps; i++) { build_map_blob(maps[i]->h, cap, blob); all.insert(all.end(), blob.begin(), blob.end()); }
    std::vector<uint32_t> cst; build_const_words(u->h, cst);
    if (dev_h2d(b->d_maps, all.data(), all.size() * 4, b->stream) || dev_h2d(b->d_cst, cst.data(), cst.size() * 4, b->stream) ||
        dev_zero(b->d_stats, 20 * sizeof(unsigned long long), b->stream) || dev_zero(b->d_units, unit_bytes, b->stream) || dev_sync(b->stream))
        return fail(MRTS_E_CUDA, std::string("device upload failed: ") + dev_errstr());
    mrts_batch *raw = b.release();
    int rc = mrts_batch_reset(raw, nullptr, 0);
    if (rc) { mrts_batch_destroy(raw); return rc; }
    *out = raw;
    return MRTS_OK;
}

int64_t mrts_batch_num_games(const mrts_batch *b) { return b ? b->n : MRTS_E_ARG; }
int mrts_batch_unit_capacity(const mrts_batch *b) { return b ? b->cap : MRTS_E_ARG; }
int mrts_batch_device(const mrts_batch *b) { return b ? b->device : MRTS_E_ARG; }
void *mrts_batch_stream(const mrts_batch *b) { return b ? (void *)b->stream : nullptr; }
int mrts_batch_sync(mrts_batch *b) { if (!b) return MRTS_E_ARG; if (dev_sync(b->stream)) return fail(MRTS_E_CUDA, std::string("stream sync: ") + dev_errstr()); return MRTS_OK; }
int64_t mrts_batch_launch_count(const mrts_batch *b) { return b ? b->launches : 0; }
const char *mrts_batch_last_kernel(const mrts_batch *b) { return b ? b->last_kernel : ""; }
int mrts_batch_io_bytes(mrts_batch *b, int64_t out[2]) {
    if (!b || !out) return fail(MRTS_E_ARG, "mrts_batch_io_bytes: null argument");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    if (dev_d2h(out, b->d_stats + MRTS_STATS_IO_SLOT, 2 * sizeof(int64_t), b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
    out[0] += b->host_io[0]; out[1] += b->host_io[1];
    return MRTS_OK;
}
int mrts_batch_num_planes(const mrts_batch *b) { return b ? ((b->flags & MRTS_FLAG_PARTIAL_OBS) ? 8 : 6) : MRTS_E_ARG; }
int mrts_batch_mask_width(const mrts_batch *b) { if (!b) return MRTS_E_ARG; int R = 2 * b->max_range + 1; return 1 + 6 + 16 + (int)b->utt.types.size() + R * R; }

static int do_reset(mrts_batch *b, const uint8_t *mask, const int64_t *seeds, int on_device, int keep_rng = 0) {
    if (b) b->results_fresh = false;
    if (!b) return fail(MRTS_E_ARG, "null batch");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    const long long *d_seeds = (const long long *)seeds; const uint8_t *d_mask = mask;
    if (!on_device && (seeds || mask)) {
        size_t sb = seeds ? (size_t)b->n * 8 : 0, mb = mask ? (size_t)b->n : 0;
        if (ensure_tmp(b, sb + mb)) return fail(MRTS_E_CUDA, std::string("staging allocation failed: ") + dev_errstr());
        if (seeds) { if (dev_h2d(b->d_tmp, seeds, sb, b->stream)) return fail(MRTS_E_CUDA, dev_errstr()); d_seeds = (const long long *)b->d_tmp; }
        if (mask) { if (dev_h2d((char *)b->d_tmp + sb, mask, mb, b->stream)) return fail(MRTS_E_CUDA, dev_errstr()); d_mask = (const uint8_t *)b->d_tmp + sb; }
    }
    ResetParams p{b->d_hdr, b->d_units, b->d_maps, d_seeds, d_mask, b->n, b->n_maps, b->map_words, b->cap, b->L.pcw, b->uw, keep_rng};
    b->launches++;
    if (!mask) { if (dev_zero(b->d_stats, 8 * sizeof(unsigned long long), b->stream) || dev_zero(b->d_stats + MRTS_STATS_IO_SLOT, 2 * sizeof(unsigned long long), b->stream)) return fail(MRTS_E_CUDA, dev_errstr()); }
    b->staged[0].valid = b->staged[1].valid = false;
#ifdef MRTS_EMU
    emu::launch(2, 128, 0, [p](unsigned char *, int tid, int bid) { reset_kernel_body(p, tid, 128, bid, 2); });
#else
    int grid = (int)std::min<long long>((b->n + 3) / 4, 148 * 16);
    k_reset<<<grid, 128, 0, b->stream>>>(p);
    if (ck(cudaGetLastError())) return fail(MRTS_E_CUDA, std::string("reset launch: ") + dev_errstr());
    if (!on_device && (seeds || mask)) if (dev_sync(b->stream)) return fail(MRTS_E_CUDA, dev_errstr()); // staging buffer reuse
#endif
    return MRTS_OK;
}
int mrts_batch_reset(mrts_batch *b, const int64_t *seeds, int on_device) { return do_reset(b, nullptr, seeds, on_device); }
int mrts_batch_reset_masked(mrts_batch *b, const uint8_t *mask, const int64_t *seeds, int on_device) {
    if (!mask) return fail(MRTS_E_ARG, "mrts_batch_reset_masked: mask is required");
    return do_reset(b, mask, seeds, on_device);
}

int mrts_batch_restart_masked(mrts_batch *b, const uint8_t *mask, int on_device) {
    if (!mask) return fail(MRTS_E_ARG, "mrts_batch_restart_masked: mask is required");
    return do_reset(b, mask, nullptr, on_device, 1);
}

int mrts_batch_copy_games(mrts_batch *dst, const mrts_batch *src, const int64_t *src_index, const uint8_t *mask, int on_device) {
    if (!dst || !src) return fail(MRTS_E_ARG, "mrts_batch_copy_games: null batch");
    if (dst->W != src->W || dst->H != src->H || dst->cap != src->cap || dst->uw != src->uw || dst->device != src->device)
        return fail(MRTS_E_ARG, "mrts_batch_copy_games: the batches must have the same map size, unit capacity, unit words and device");
    if (!src_index && dst->n > src->n) return fail(MRTS_E_ARG, "mrts_batch_copy_games: without src_index the source must hold at least as many games");
    if (dev_select(dst->device)) return fail(MRTS_E_CUDA, dev_errstr());
    dst->results_fresh = false;
    const long long *d_idx = (const long long *)src_index; const uint8_t *d_mask = mask;
    if (!on_device && (src_index || mask)) {
        size_t ib = src_index ? (size_t)dst->n * 8 : 0, mb = mask ? (size_t)dst->n : 0;
        if (ensure_tmp(dst, ib + mb)) return fail(MRTS_E_CUDA, std::string("staging allocation failed: ") + dev_errstr());
        if (src_index) { if (dev_h2d(dst->d_tmp, src_index, ib, dst->stream)) return fail(MRTS_E_CUDA, dev_errstr()); d_idx = (const long long *)dst->d_tmp; }
        if (mask) { if (dev_h2d((char *)dst->d_tmp + ib, mask, mb, dst->stream)) return fail(MRTS_E_CUDA, dev_errstr()); d_mask = (const uint8_t *)dst->d_tmp + ib; }
    }
    if (dev_sync(src->stream)) return fail(MRTS_E_CUDA, dev_errstr()); // the source's pending work is finished before its state is read on dst's stream
    CopyParams p{dst->d_hdr, dst->d_units, src->d_hdr, src->d_units, d_idx, d_mask, dst->n, src->n, dst->cap, dst->uw, nullptr};
    dst->launches++;
#ifdef MRTS_EMU
    emu::launch(2, 128, 0, [p](unsigned char *, int tid, int bid) { copy_kernel_body(p, tid, 128, bid, 2); });
#else
    int grid = (int)std::min<long long>((dst->n + 3) / 4, 148 * 16);
    k_copy_games<<<grid, 128, 0, dst->stream>>>(p);
    if (ck(cudaGetLastError())) return fail(MRTS_E_CUDA, std::string("copy launch: ") + dev_errstr());
    if (!on_device && (src_index || mask)) if (dev_sync(dst->stream)) return fail(MRTS_E_CUDA, dev_errstr()); // staging buffer reuse
#endif
    return MRTS_OK;
}

int mrts_batch_scatter_games(mrts_batch *dst, const mrts_batch *src, const int64_t *dst_index, int on_device) {
    if (!dst || !src || !dst_index) return fail(MRTS_E_ARG, "mrts_batch_scatter_games: null argument");
    if (dst->W != src->W || dst->H != src->H || dst->cap != src->cap || dst->uw != src->uw || dst->device != src->device)
        return fail(MRTS_E_ARG, "mrts_batch_scatter_games: the batches must have the same map size, unit capacity, unit words and device");
    if (dev_select(dst->device)) return fail(MRTS_E_CUDA, dev_errstr());
    dst->results_fresh = false;
    const long long *d_idx = (const long long *)dst_index;
    if (!on_device) {
        if (ensure_tmp(dst, (size_t)src->n * 8) || dev_h2d(dst->d_tmp, dst_index, (size_t)src->n * 8, dst->stream)) return fail(MRTS_E_CUDA, dev_errstr());
        d_idx = (const long long *)dst->d_tmp;
    }
    if (dev_sync(src->stream)) return fail(MRTS_E_CUDA, dev_errstr());
    CopyParams p{dst->d_hdr, dst->d_units, src->d_hdr, src->d_units, nullptr, nullptr, dst->n, src->n, dst->cap, dst->uw, d_idx};
    dst->launches++;
#ifdef MRTS_EMU
    emu::launch(2, 128, 0, [p](unsigned char *, int tid, int bid) { copy_kernel_body(p, tid, 128, bid, 2); });
#else
    int grid = (int)std::min<long long>((src->n + 3) / 4, 148 * 16);
    k_copy_games<<<grid, 128, 0, dst->stream>>>(p);
    if (ck(cudaGetLastError())) return fail(MRTS_E_CUDA, std::string("copy launch: ") + dev_errstr());
    if (!on_device && dev_sync(dst->stream)) return fail(MRTS_E_CUDA, dev_errstr());
#endif
    return MRTS_OK;
}

int mrts_batch_set_policy(mrts_batch *b, int player, int policy, int pathfinder) {
    if (!b || player < 0 || player > 1) return fail(MRTS_E_ARG, "mrts_batch_set_policy: bad argument");
    if (policy < MRTS_POLICY_EXTERNAL || policy > MRTS_POLICY_EMR_DETERMINISTICO) return fail(MRTS_E_ARG, "mrts_batch_set_policy: unknown policy");
    if (policy >= MRTS_POLICY_WORKER_RUSH && !b->scripted)
        return fail(MRTS_E_STATE, "scripted policies need a batch created with MRTS_FLAG_SCRIPTED_AI");
    if (pathfinder < MRTS_PF_ASTAR || pathfinder > MRTS_PF_FLOODFILL) return fail(MRTS_E_ARG, "unknown pathfinder");
    if (pathfinder == MRTS_PF_FLOODFILL && !b->d_ff) {
        // one cache per game and player: last frame + a valid bit and a W*H map of u16 distances per target position (scripted.cuh)
        long long cells = (long long)b->W * b->H, words = 1 + (cells + 31) / 32; words += words & 1;
        b->ff_stride = words * 4 + cells * cells * 2; b->ff_stride = (b->ff_stride + 15) & ~15LL;
        unsigned long long total = (unsigned long long)b->n * 2ULL * (unsigned long long)b->ff_stride;
        if (total > (64ULL << 30)) return fail(MRTS_E_LIMIT, "FloodFillPathFinding keeps one distance map per target cell, game and player: this batch would need more than 64 GB");
        if (dev_select(b->device) || dev_alloc((void **)&b->d_ff, (size_t)total) || dev_zero(b->d_ff, (size_t)total, b->stream)) { b->d_ff = nullptr; return fail(MRTS_E_CUDA, std::string("FloodFill cache allocation failed: ") + dev_errstr()); }
    }
    b->policy[player] = policy; b->pathfinder[player] = pathfinder;
    return MRTS_OK;
}

int mrts_batch_set_auto_reset(mrts_batch *b, int enable) { if (!b) return fail(MRTS_E_ARG, "null batch"); b->auto_reset = enable ? 1 : 0; return MRTS_OK; }

static int stage_actions(mrts_batch *b, int player, int format, const int32_t *actions, const int32_t *counts, int max_k, int fill, int on_device) {
    if (!b || player < 0 || player > 1 || max_k < 0 || (max_k > 0 && !actions)) return fail(MRTS_E_ARG, "bad action arguments");
    if (format != MRTS_ACTIONS_VECTOR && format != MRTS_ACTIONS_RAW) return fail(MRTS_E_ARG, "unknown action format");
    if (max_k > b->cap) return fail(MRTS_E_ARG, "max_k exceeds the unit capacity of the batch");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    Staged &s = b->staged[player];
    size_t rows = (size_t)b->n * (size_t)std::max(max_k, 1);
    if (rows > s.cap_rows || !s.counts) {
        dev_free(s.actions); dev_free(s.counts); s.actions = nullptr; s.counts = nullptr; s.cap_rows = 0;
        if (dev_alloc((void **)&s.actions, rows * 8 * 4) || dev_alloc((void **)&s.counts, (size_t)b->n * 4)) return fail(MRTS_E_CUDA, std::string("staging allocation failed: ") + dev_errstr());
        s.cap_rows = rows;
    }
    size_t ab = (size_t)b->n * max_k * 8 * 4;
    if (ab) { if (on_device ? dev_d2d(s.actions, actions, ab, b->stream) : dev_h2d(s.actions, actions, ab, b->stream)) return fail(MRTS_E_CUDA, dev_errstr()); }
    if (counts) { if (on_device ? dev_d2d(s.counts, counts, (size_t)b->n * 4, b->stream) : dev_h2d(s.counts, counts, (size_t)b->n * 4, b->stream)) return fail(MRTS_E_CUDA, dev_errstr()); }
    s.has_counts = counts != nullptr; // without counts every game has max_k rows (the kernels take the count from max_k)
    if (!on_device && dev_sync(b->stream)) return fail(MRTS_E_CUDA, dev_errstr()); // caller may reuse its host buffers
    s.max_k = max_k; s.format = format; s.fill = fill; s.valid = true; s.rows = s.actions; s.stride = (long long)max_k * 8;
    return MRTS_OK;
}

// Both players' PlayerActions of every game in ONE array laid out like JNIGridnetVecClient's self-play environments
// (src/tests/JNIGridnetVecClient.java:226-236): row block 2g is player 0 of game g, row block 2g + 1 is player 1.  One host -> device
// copy; with `async` the call returns without waiting for it (the host array must stay untouched until the batch is synchronised).
int mrts_batch_set_actions_interleaved(mrts_batch *b, int format, const int32_t *actions, int max_k, int fill_none_duration, int on_device, int async) {
    if (!b || max_k < 1 || !actions) return fail(MRTS_E_ARG, "bad action arguments");
    if (format != MRTS_ACTIONS_VECTOR && format != MRTS_ACTIONS_RAW) return fail(MRTS_E_ARG, "unknown action format");
    if (max_k > b->cap) return fail(MRTS_E_ARG, "max_k exceeds the unit capacity of the batch");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    Staged &s = b->staged[0];
    size_t rows = (size_t)b->n * 2 * (size_t)max_k;
    if (rows > s.cap_rows || !s.counts) {
        dev_free(s.actions); dev_free(s.counts); s.actions = nullptr; s.counts = nullptr; s.cap_rows = 0;
        if (dev_alloc((void **)&s.actions, rows * 8 * 4) || dev_alloc((void **)&s.counts, (size_t)b->n * 4)) return fail(MRTS_E_CUDA, std::string("staging allocation failed: ") + dev_errstr());
        s.cap_rows = rows;
    }
    if (on_device ? dev_d2d(s.actions, actions, rows * 32, b->stream) : dev_h2d(s.actions, actions, rows * 32, b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
    if (!on_device && !async && dev_sync(b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
    for (int pl = 0; pl < 2; pl++) {
        Staged &t = b->staged[pl];
        t.rows = s.actions + (size_t)pl * max_k * 8; t.stride = (long long)max_k * 16; t.max_k = max_k; t.format = format; t.fill = fill_none_duration;
        t.has_counts = false; t.valid = true;
    }
    return MRTS_OK;
}

int mrts_batch_set_actions(mrts_batch *b, int player, int format, const int32_t *actions, const int32_t *counts, int max_k, int fill_none_duration, int on_device) {
    return stage_actions(b, player, format, actions, counts, max_k, fill_none_duration, on_device);
}

static void fill_ext(mrts_batch *b, StepParams &p, int pl) {
    Staged &s = b->staged[pl];
    if (s.valid) { p.ext_actions[pl] = s.rows; p.ext_counts[pl] = s.has_counts ? s.counts : nullptr; p.ext_maxk[pl] = s.max_k; p.ext_format[pl] = s.format; p.ext_fill[pl] = s.fill; p.ext_stride[pl] = s.stride; }
}

int mrts_batch_issue(mrts_batch *b, int player, int format, const int32_t *actions, const int32_t *counts, int max_k, int fill_none_duration, int safe, int on_device) {
    if (b) b->results_fresh = false;
    int rc = stage_actions(b, player, format, actions, counts, max_k, fill_none_duration, on_device);
    if (rc) return rc;
    StepParams p; memset(&p, 0, sizeof p);
    p.mode = MODE_ISSUE_ONLY; p.issue_player = player; p.safe = safe ? 1 : 0;
    fill_ext(b, p, player);
    b->staged[player].valid = false;
    if (launch_step(b, p)) return fail(MRTS_E_CUDA, std::string("issue launch: ") + dev_errstr());
    return MRTS_OK;
}

int mrts_batch_step(mrts_batch *b, int n_cycles, int max_cycles) {
    if (!b || n_cycles < 0) return fail(MRTS_E_ARG, "mrts_batch_step: bad argument");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    StepParams p; memset(&p, 0, sizeof p);
    p.mode = MODE_GAME; p.n_cycles = n_cycles; p.max_cycles = max_cycles; p.safe = 1; p.auto_reset = b->auto_reset;
    for (int pl = 0; pl < 2; pl++) { p.policy[pl] = b->policy[pl]; p.pathfinder[pl] = b->pathfinder[pl]; if (b->policy[pl] == MRTS_POLICY_EXTERNAL) fill_ext(b, p, pl); b->staged[pl].valid = false; }
    p.sequential_issue = b->sequential_issue; p.info_out = b->info_out;
    p.tm_worker = b->tm[0]; p.tm_building = b->tm[1]; p.tm_combat = b->tm[2]; p.tm_base = b->tm[3]; p.tm_mobile = b->tm[4]; p.tm_resource = b->tm[5];
    bool fused = !(b->flags & MRTS_FLAG_PARTIAL_OBS); // partially observable batches observe in a second launch
    if (fused) { p.obs_out[0] = b->obs_out[0]; p.obs_out[1] = b->obs_out[1]; p.obs_dtype = b->obs_dtype; }
    p.out_stride = b->out_stride; p.vec_reset = b->vec_reset; p.vec_max_steps = b->vec_max_steps;
    p.mask_out[0] = b->mask_out[0]; p.mask_out[1] = b->mask_out[1]; // masks are of the full state (JNIGridnetClient.getMasks uses gs), fused in every batch
    p.results_out = b->d_results;
    if (launch_step(b, p)) return fail(MRTS_E_CUDA, std::string("step launch: ") + dev_errstr());
    b->results_fresh = true;
    if (!fused)
        for (int pl = 0; pl < 2; pl++)
            if (b->obs_out[pl]) { int rc = mrts_batch_observe(b, pl, b->obs_dtype, b->obs_out[pl], 1); if (rc) return rc; }
    return MRTS_OK;
}

// JNIGridnetVecClient.gameStep for the self-play environments in one call: both players' vector actions in environment order, then the
// one-cycle step (src/tests/JNIGridnetVecClient.java:226-236)
int mrts_batch_vec_step(mrts_batch *b, const int32_t *actions, int max_k, int on_device, int async) {
    int rc = mrts_batch_set_actions_interleaved(b, MRTS_ACTIONS_VECTOR, actions, max_k, 1, on_device, async);
    if (rc) return rc;
    return mrts_batch_step(b, 1, 0x3fffffff);
}

int mrts_batch_set_issue_order(mrts_batch *b, int sequential) { if (!b) return fail(MRTS_E_ARG, "null batch"); b->sequential_issue = sequential ? 1 : 0; return MRTS_OK; }
int mrts_batch_set_info_output(mrts_batch *b, int32_t *out) { if (!b) return fail(MRTS_E_ARG, "null batch"); b->info_out = out; return MRTS_OK; }

int mrts_batch_set_observation_outputs(mrts_batch *b, int dtype, void *out_player0, void *out_player1) {
    if (!b || (dtype != MRTS_DTYPE_U8 && dtype != MRTS_DTYPE_I32)) return fail(MRTS_E_ARG, "mrts_batch_set_observation_outputs: bad argument");
    if ((((uintptr_t)out_player0) | ((uintptr_t)out_player1)) & 15) return fail(MRTS_E_ARG, "observation buffers must be 16-byte aligned");
    b->obs_out[0] = out_player0; b->obs_out[1] = out_player1; b->obs_dtype = dtype;
    return MRTS_OK;
}

int mrts_batch_set_vec_autoreset(mrts_batch *b, int done_mode, int max_steps) {
    if (!b || done_mode < 0 || done_mode > 3 || (done_mode && max_steps < 1)) return fail(MRTS_E_ARG, "mrts_batch_set_vec_autoreset: bad argument");
    if (done_mode && (b->flags & MRTS_FLAG_PARTIAL_OBS)) return fail(MRTS_E_STATE, "in-kernel auto-reset needs fused observations (not a MRTS_FLAG_PARTIAL_OBS batch)");
    b->vec_reset = done_mode; b->vec_max_steps = max_steps;
    return MRTS_OK;
}

int mrts_batch_set_output_stride(mrts_batch *b, int game_stride) {
    if (!b || game_stride < 1) return fail(MRTS_E_ARG, "mrts_batch_set_output_stride: bad argument");
    b->out_stride = game_stride;
    return MRTS_OK;
}

int mrts_batch_set_mask_outputs(mrts_batch *b, void *out_player0, void *out_player1) {
    if (!b) return fail(MRTS_E_ARG, "null batch");
    if ((((uintptr_t)out_player0) | ((uintptr_t)out_player1)) & 15) return fail(MRTS_E_ARG, "mask buffers must be 16-byte aligned");
    if (mrts_batch_mask_width(b) > 128) return fail(MRTS_E_LIMIT, "fused masks hold at most 128 elements per cell (unit type table with a very long attack range): use mrts_batch_masks");
    b->mask_out[0] = out_player0; b->mask_out[1] = out_player1;
    return MRTS_OK;
}

int mrts_batch_cycle_to(mrts_batch *b, const int32_t *t_target, int n_cycles, int on_device) {
    if (b) b->results_fresh = false;
    if (!b) return fail(MRTS_E_ARG, "null batch");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    StepParams p; memset(&p, 0, sizeof p);
    p.mode = MODE_CYCLE_ONLY; p.n_cycles = n_cycles;
    if (t_target) {
        if (on_device) p.t_target = t_target;
        else {
            if (ensure_tmp(b, (size_t)b->n * 4) || dev_h2d(b->d_tmp, t_target, (size_t)b->n * 4, b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
            p.t_target = (const int32_t *)b->d_tmp;
        }
    }
    if (launch_step(b, p)) return fail(MRTS_E_CUDA, std::string("cycle launch: ") + dev_errstr());
    if (t_target && !on_device && dev_sync(b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
    return MRTS_OK;
}

int mrts_batch_cycle_to_decision(mrts_batch *b) {
    if (b) b->results_fresh = false;
    if (!b) return fail(MRTS_E_ARG, "null batch");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    StepParams p; memset(&p, 0, sizeof p);
    p.mode = MODE_CYCLE_DECISION;
    if (launch_step(b, p)) return fail(MRTS_E_CUDA, std::string("cycle launch: ") + dev_errstr());
    return MRTS_OK;
}

int mrts_batch_unit_actions(mrts_batch *b, int player, int none_duration, int max_choices, int max_actions, int32_t *out_hdr, int32_t *out_positions,
                            int32_t *out_choices, int32_t *out_lists, int on_device) {
    if (!b || player < 0 || player > 1 || max_choices < 1 || max_actions < 1 || !out_hdr || !out_positions || !out_choices || !out_lists)
        return fail(MRTS_E_ARG, "mrts_batch_unit_actions: bad argument");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    size_t nh = (size_t)b->n * 8, np = (size_t)b->n * b->cap, nc = (size_t)b->n * max_choices * 4, nl = (size_t)b->n * max_choices * max_actions;
    StepParams p; memset(&p, 0, sizeof p);
    p.mode = MODE_UNIT_ACTIONS; p.out_player = player; p.ua_none_duration = none_duration; p.ua_max_choices = max_choices; p.ua_max_actions = max_actions;
    if (on_device) { p.ua_hdr = out_hdr; p.ua_pos = out_positions; p.ua_choice = out_choices; p.ua_list = out_lists; }
    else {
        if (ensure_tmp(b, (nh + np + nc + nl) * 4)) return fail(MRTS_E_CUDA, std::string("staging allocation failed: ") + dev_errstr());
        int32_t *t = (int32_t *)b->d_tmp;
        p.ua_hdr = t; p.ua_pos = t + nh; p.ua_choice = t + nh + np; p.ua_list = t + nh + np + nc;
    }
    if (launch_step(b, p)) return fail(MRTS_E_CUDA, std::string("unit actions launch: ") + dev_errstr());
    if (!on_device) {
        if (dev_d2h(out_hdr, p.ua_hdr, nh * 4, b->stream) || dev_d2h(out_positions, p.ua_pos, np * 4, b->stream) ||
            dev_d2h(out_choices, p.ua_choice, nc * 4, b->stream) || dev_d2h(out_lists, p.ua_list, nl * 4, b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
    }
    return MRTS_OK;
}

int mrts_batch_evaluate(mrts_batch *b, int eval_fn, int maxplayer, int observer, float *out_eval, int on_device) {
    if (!b || !out_eval || maxplayer < 0 || maxplayer > 1 || observer > 1 || (eval_fn != 0 && eval_fn != 1)) return fail(MRTS_E_ARG, "mrts_batch_evaluate: bad argument");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    StepParams p; memset(&p, 0, sizeof p);
    p.mode = MODE_ROLLOUT; p.rollouts_per_game = 1; p.depth = -1; p.eval_fn = eval_fn; p.maxplayer = maxplayer; p.observer = observer < 0 ? -1 : observer;
    if (on_device) p.ro_eval = out_eval;
    else { if (ensure_tmp(b, (size_t)b->n * 4)) return fail(MRTS_E_CUDA, dev_errstr()); p.ro_eval = (float *)b->d_tmp; }
    if (launch_step(b, p)) return fail(MRTS_E_CUDA, std::string("evaluate launch: ") + dev_errstr());
    if (!on_device && dev_d2h(out_eval, p.ro_eval, (size_t)b->n * 4, b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
    return MRTS_OK;
}

int mrts_batch_pathfind(mrts_batch *b, int pathfinder, const int32_t *queries, int32_t *out_dir, int on_device) {
    if (!b || !queries || !out_dir) return fail(MRTS_E_ARG, "mrts_batch_pathfind: null argument");
    if (pathfinder < MRTS_PF_ASTAR || pathfinder > MRTS_PF_GREEDY) return fail(MRTS_E_ARG, "unknown pathfinder (FloodFillPathFinding is stateful: it only runs inside a scripted policy)");
    if (!b->scripted) return fail(MRTS_E_STATE, "pathfinding needs a batch created with MRTS_FLAG_SCRIPTED_AI (it owns the search scratch)");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    StepParams p; memset(&p, 0, sizeof p);
    p.mode = MODE_PATHFIND; p.pf_kind = pathfinder;
    size_t qb = (size_t)b->n * 12, ob = (size_t)b->n * 4;
    if (on_device) { p.pf_query = queries; p.pf_out = out_dir; }
    else {
        if (ensure_tmp(b, qb + ob) || dev_h2d(b->d_tmp, queries, qb, b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
        p.pf_query = (const int32_t *)b->d_tmp; p.pf_out = (int32_t *)((char *)b->d_tmp + qb);
    }
    if (launch_step(b, p)) return fail(MRTS_E_CUDA, std::string("pathfind launch: ") + dev_errstr());
    if (!on_device && dev_d2h(out_dir, p.pf_out, ob, b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
    return MRTS_OK;
}

int mrts_batch_rollout(mrts_batch *b, int rollouts_per_game, int depth, int eval_fn, int maxplayer, int observer, const int64_t *seeds,
                       float *out_eval, int32_t *out_time, int on_device) {
    if (!b || rollouts_per_game < 1 || depth < 0 || maxplayer < 0 || maxplayer > 1 || observer > 1 || (eval_fn != 0 && eval_fn != 1))
        return fail(MRTS_E_ARG, "mrts_batch_rollout: bad argument");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    size_t nr = (size_t)b->n * rollouts_per_game;
    StepParams p; memset(&p, 0, sizeof p);
    p.mode = MODE_ROLLOUT; p.rollouts_per_game = rollouts_per_game; p.depth = depth; p.eval_fn = eval_fn; p.maxplayer = maxplayer;
    p.observer = observer < 0 ? -1 : observer;
    if (on_device) { p.ro_seeds = (const long long *)seeds; p.ro_eval = out_eval; p.ro_time = out_time; }
    else {
        if (ensure_tmp(b, nr * 16)) return fail(MRTS_E_CUDA, std::string("staging allocation failed: ") + dev_errstr());
        char *t = (char *)b->d_tmp;
        if (seeds) { if (dev_h2d(t, seeds, nr * 8, b->stream)) return fail(MRTS_E_CUDA, dev_errstr()); p.ro_seeds = (const long long *)t; }
        p.ro_eval = (float *)(t + nr * 8); p.ro_time = (int32_t *)(t + nr * 12);
    }
    if (launch_step(b, p)) return fail(MRTS_E_CUDA, std::string("rollout launch: ") + dev_errstr());
    if (!on_device) {
        if (out_eval && dev_d2h(out_eval, p.ro_eval, nr * 4, b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
        if (out_time && dev_d2h(out_time, p.ro_time, nr * 4, b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
        if (dev_sync(b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
    }
    return MRTS_OK;
}

static int emit(mrts_batch *b, int mode, int player, int dtype, void *out, int on_device, size_t elems) {
    bool bits = mode == MODE_MASKS && dtype == MRTS_DTYPE_BITS;
    if (!b || !out || player < 0 || player > 1 || (dtype != MRTS_DTYPE_U8 && dtype != MRTS_DTYPE_I32 && !bits)) return fail(MRTS_E_ARG, "bad observation/mask argument");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    size_t bytes = bits ? (size_t)b->n * b->W * b->H * ((mrts_batch_mask_width(b) + 7) / 8) : elems * (dtype == MRTS_DTYPE_U8 ? 1 : 4);
    void *d_out = out;
    if (on_device && (((uintptr_t)out) & 15)) return fail(MRTS_E_ARG, "on-device observation/mask buffers must be 16-byte aligned (the planes are written with 16-byte stores)");
    if (!on_device) { if (ensure_tmp(b, bytes)) return fail(MRTS_E_CUDA, std::string("staging allocation failed: ") + dev_errstr()); d_out = b->d_tmp; }
    if (mode == MODE_OBSERVE && !(b->flags & MRTS_FLAG_PARTIAL_OBS)) {
        if (launch_observe(b, player, dtype, d_out)) return fail(MRTS_E_CUDA, std::string("launch: ") + dev_errstr());
    } else {
        StepParams p; memset(&p, 0, sizeof p);
        p.mode = mode; p.out = d_out; p.out_dtype = dtype; p.out_player = player;
        if (mode == MODE_MASKS && on_device) p.out_stride = b->out_stride; // on-device mask outputs follow mrts_batch_set_output_stride (environment order)
        if (launch_step(b, p)) return fail(MRTS_E_CUDA, std::string("launch: ") + dev_errstr());
    }
    if (!on_device && dev_d2h(out, d_out, bytes, b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
    return MRTS_OK;
}
int mrts_batch_observe(mrts_batch *b, int player, int dtype, void *out, int on_device) {
    if (!b) return fail(MRTS_E_ARG, "null batch");
    return emit(b, MODE_OBSERVE, player, dtype, out, on_device, (size_t)b->n * mrts_batch_num_planes(b) * b->W * b->H);
}
int mrts_batch_masks(mrts_batch *b, int player, int dtype, void *out, int on_device) {
    if (!b) return fail(MRTS_E_ARG, "null batch");
    return emit(b, MODE_MASKS, player, dtype, out, on_device, (size_t)b->n * b->W * b->H * mrts_batch_mask_width(b));
}

int mrts_batch_results(mrts_batch *b, int32_t *out, int on_device) {
    if (!b || !out) return fail(MRTS_E_ARG, "mrts_batch_results: null argument");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    if (b->results_fresh) { // the last step left them in d_results: a copy, no kernel
        if (on_device ? dev_d2d(out, b->d_results, (size_t)b->n * 16, b->stream) : dev_d2h(out, b->d_results, (size_t)b->n * 16, b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
        return MRTS_OK;
    }
    int32_t *d_out = out;
    if (!on_device) { if (ensure_tmp(b, (size_t)b->n * 16)) return fail(MRTS_E_CUDA, dev_errstr()); d_out = (int32_t *)b->d_tmp; }
    ResultParams p{b->d_hdr, b->d_units, d_out, b->n, b->cap, b->uw};
    b->launches++;
#ifdef MRTS_EMU
    for (long long g = 0; g < b->n; g++) results_kernel_body(p, g);
#else
    k_results<<<(unsigned)((b->n + 127) / 128), 128, 0, b->stream>>>(p);
    if (ck(cudaGetLastError())) return fail(MRTS_E_CUDA, std::string("results launch: ") + dev_errstr());
#endif
    if (!on_device && dev_d2h(out, d_out, (size_t)b->n * 16, b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
    return MRTS_OK;
}

int mrts_batch_copy_to_host(mrts_batch *b, void *host_dst, const void *device_src, size_t bytes) {
    if (!b || !host_dst || !device_src) return fail(MRTS_E_ARG, "mrts_batch_copy_to_host: null argument");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
#ifdef MRTS_EMU
    memcpy(host_dst, device_src, bytes);
#else
    if (ck(cudaMemcpyAsync(host_dst, device_src, bytes, cudaMemcpyDeviceToHost, b->stream))) return fail(MRTS_E_CUDA, dev_errstr());
#endif
    return MRTS_OK;
}

int mrts_batch_stats(mrts_batch *b, int64_t out[8]) {
    if (!b || !out) return fail(MRTS_E_ARG, "mrts_batch_stats: null argument");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    if (dev_d2h(out, b->d_stats, 8 * sizeof(int64_t), b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
    return MRTS_OK;
}

// ---- the run's single collective: all-reduce (sum) of the eight counters over NCCL --------------------------------
// NCCL is resolved at run time (dlopen) so that the library neither links a second copy next to the one a host process may
// already carry nor needs NCCL at all for single-GPU use.  Types per nccl.h (2.x ABI): ncclUniqueId = 128 bytes, ncclComm_t opaque,
// ncclInt64 = 4, ncclSum = 0, ncclSuccess = 0.
struct mrts_comm { void *comm = nullptr; int n_ranks = 1, rank = 0, device = 0; bool owned = true; };
#ifndef MRTS_EMU
namespace {
struct NcclApi {
    struct Uid { char b[128]; }; // ncclUniqueId, passed by value
    void *h = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, Uid, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
int nccl_load() {
    if (g_nccl.h) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); // the copy the host process already loaded, if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return -1;
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) return -1;
    g_nccl.h = h;
    return 0;
}
std::string nccl_err(int rc) { return g_nccl.GetErrorString ? std::string(g_nccl.GetErrorString(rc)) : ("NCCL error " + std::to_string(rc)); }
}
#endif

int mrts_nccl_unique_id(uint8_t out[MRTS_NCCL_UNIQUE_ID_BYTES]) {
    if (!out) return fail(MRTS_E_ARG, "mrts_nccl_unique_id: null argument");
#ifdef MRTS_EMU
    memset(out, 0, MRTS_NCCL_UNIQUE_ID_BYTES); return MRTS_OK;
#else
    if (nccl_load()) return fail(MRTS_E_STATE, "libnccl.so.2 cannot be loaded");
    int rc = g_nccl.GetUniqueId(out);
    if (rc) return fail(MRTS_E_CUDA, "ncclGetUniqueId: " + nccl_err(rc));
    return MRTS_OK;
#endif
}

int mrts_nccl_comm_create(const uint8_t id[MRTS_NCCL_UNIQUE_ID_BYTES], int n_ranks, int rank, int device, mrts_comm **out) {
    if (!id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(MRTS_E_ARG, "mrts_nccl_comm_create: bad argument");
    auto c = std::make_unique<mrts_comm>();
    c->n_ranks = n_ranks; c->rank = rank; c->device = device;
#ifndef MRTS_EMU
    if (nccl_load()) return fail(MRTS_E_STATE, "libnccl.so.2 cannot be loaded");
    if (dev_select(device)) return fail(MRTS_E_CUDA, std::string("cannot select CUDA device: ") + dev_errstr());
    NcclApi::Uid uid; memcpy(uid.b, id, sizeof uid.b);
    int rc = g_nccl.CommInitRank(&c->comm, n_ranks, uid, rank);
    if (rc) return fail(MRTS_E_CUDA, "ncclCommInitRank: " + nccl_err(rc));
#endif
    *out = c.release();
    return MRTS_OK;
}

int mrts_nccl_comm_wrap(void *nccl_comm, int device, mrts_comm **out) {
    if (!nccl_comm || !out) return fail(MRTS_E_ARG, "mrts_nccl_comm_wrap: null argument");
#ifndef MRTS_EMU
    if (nccl_load()) return fail(MRTS_E_STATE, "libnccl.so.2 cannot be loaded");
#endif
    auto *c = new mrts_comm; c->comm = nccl_comm; c->device = device; c->owned = false; c->n_ranks = 0;
    *out = c;
    return MRTS_OK;
}

void mrts_nccl_comm_destroy(mrts_comm *c) {
    if (!c) return;
#ifndef MRTS_EMU
    if (c->owned && c->comm && g_nccl.CommDestroy) { cudaSetDevice(c->device); g_nccl.CommDestroy(c->comm); }
#endif
    delete c;
}

int mrts_batch_stats_allreduce(mrts_batch *b, int64_t out[8], mrts_comm *comm) {
    if (!comm) return mrts_batch_stats(b, out);
    if (!b || !out) return fail(MRTS_E_ARG, "mrts_batch_stats_allreduce: null argument");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    unsigned long long *scratch = b->d_stats + 10;
    if (dev_d2d(scratch, b->d_stats, 8 * sizeof(int64_t), b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
#ifndef MRTS_EMU
    if (comm->comm) { // one ncclAllReduce of 8 int64 on the batch's stream: nothing to fuse with, the counters are final when it runs
        int rc = g_nccl.AllReduce(scratch, scratch, 8, /* ncclInt64 */ 4, /* ncclSum */ 0, comm->comm, b->stream);
        if (rc) return fail(MRTS_E_CUDA, "ncclAllReduce: " + nccl_err(rc));
    }
#endif
    if (dev_d2h(out, scratch, 8 * sizeof(int64_t), b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
    return MRTS_OK;
}

// ---- export / import --------------------------------------------------------------------------------------------------
int mrts_batch_export(mrts_batch *b, int64_t first, int64_t count, mrts_state_host *out) {
    if (!b || !out || first < 0 || count < 0 || first + count > b->n) return fail(MRTS_E_ARG, "mrts_batch_export: bad range");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    int cap = b->cap;
    std::vector<int32_t> hdr((size_t)count * MRTS_HDR_WORDS);
    std::vector<uint32_t> un((size_t)count * b->uw * cap);
    if (count && (dev_d2h(hdr.data(), b->d_hdr + first * MRTS_HDR_WORDS, hdr.size() * 4, b->stream) ||
                  dev_d2h(un.data(), b->d_units + first * (long long)b->uw * cap, un.size() * 4, b->stream)))
        return fail(MRTS_E_CUDA, dev_errstr());
    for (int64_t g = 0; g < count; g++) {
        const int32_t *h = &hdr[g * MRTS_HDR_WORDS];
        const uint32_t *u = &un[g * (size_t)b->uw * cap];
        int n = h[H_NUNITS], c0 = 0, c1 = 0;
        for (int i = 0; i < n; i++) { int pl = (u[UW_W0 * cap + i] >> 8) & 0xff; c0 += pl == 1; c1 += pl == 2; }
        if (out->header) {
            int32_t *o = out->header + g * 8;
            o[0] = h[H_TIME]; o[1] = h[H_RES0]; o[2] = h[H_RES1]; o[3] = n;
            o[4] = (c0 > 0 && c1 == 0) ? 0 : ((c1 > 0 && c0 == 0) ? 1 : -1);
            o[5] = (c0 == 0 || c1 == 0) ? 1 : 0; o[6] = h[H_ERR]; o[7] = h[H_NEXTID];
        }
        // rank of each assignment in insertion order
        std::vector<std::pair<uint32_t, int>> order;
        for (int i = 0; i < n; i++) if ((u[UW_A0 * cap + i] & 0xF) != AT_IDLE) order.emplace_back(u[UW_SEQ * cap + i], i);
        std::sort(order.begin(), order.end());
        std::vector<int> rank(n, 0);
        for (size_t r = 0; r < order.size(); r++) rank[order[r].second] = (int)r;
        for (int i = 0; i < cap; i++) {
            int32_t *ou = out->units ? out->units + (g * cap + i) * 8 : nullptr;
            int32_t *oa = out->actions ? out->actions + (g * cap + i) * 8 : nullptr;
            if (ou) memset(ou, 0, 32);
            if (oa) memset(oa, 0, 32);
            if (i >= n) continue;
            uint32_t w0 = u[UW_W0 * cap + i], w1 = u[UW_W1 * cap + i], a0 = u[UW_A0 * cap + i];
            bool has = (a0 & 0xF) != AT_IDLE;
            if (ou) { ou[0] = w0 & 0xff; ou[1] = (int)((w0 >> 8) & 0xff) - 1; ou[2] = (w0 >> 16) & 0xff; ou[3] = w0 >> 24; ou[4] = (int16_t)(w1 >> 16); ou[5] = (int16_t)(w1 & 0xffff); ou[6] = (int32_t)u[UW_ID * cap + i]; ou[7] = has; }
            if (oa && has) {
                int at = a0 & 0xF, ut = (a0 >> 8) & 0xff;
                oa[0] = at; oa[1] = (int32_t)u[UW_A1 * cap + i];
                oa[2] = at == MRTS_ATTACK ? (int)((a0 >> 16) & 0xff) : 0; oa[3] = at == MRTS_ATTACK ? (int)(a0 >> 24) : 0;
                oa[4] = ut == 0xFF ? -1 : ut; oa[5] = (int32_t)u[UW_TIS * cap + i]; oa[6] = rank[i];
            }
        }
        if (out->rng) for (int k = 0; k < 3; k++) out->rng[g * 3 + k] = (int64_t)((uint64_t)(uint32_t)h[H_RNGP_LO + 2 * k] | ((uint64_t)(uint32_t)h[H_RNGP_HI + 2 * k] << 32));
    }
    return MRTS_OK;
}

int mrts_batch_import(mrts_batch *b, int64_t first, int64_t count, const mrts_state_host *in) {
    if (b) b->results_fresh = false;
    if (!b || !in || !in->header || !in->units || first < 0 || count < 0 || first + count > b->n) return fail(MRTS_E_ARG, "mrts_batch_import: bad argument");
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    int cap = b->cap;
    std::vector<int32_t> hdr((size_t)count * MRTS_HDR_WORDS, 0);
    std::vector<uint32_t> un((size_t)count * b->uw * cap, 0);
    for (int64_t g = 0; g < count; g++) {
        int32_t *h = &hdr[g * MRTS_HDR_WORDS];
        uint32_t *u = &un[g * (size_t)b->uw * cap];
        const int32_t *ih = in->header + g * 8;
        int n = ih[3];
        if (n < 0 || n > cap) return fail(MRTS_E_LIMIT, "mrts_batch_import: more units than the batch capacity");
        h[H_TIME] = ih[0]; h[H_RES0] = ih[1]; h[H_RES1] = ih[2]; h[H_NUNITS] = n; h[H_ERR] = ih[6]; h[H_NEXTID] = ih[7];
        std::vector<std::pair<int, int>> order;
        long long maxid = -1;
        for (int i = 0; i < n; i++) {
            const int32_t *iu = in->units + (g * cap + i) * 8;
            u[UW_W0 * cap + i] = (uint32_t)(iu[0] & 0xff) | ((uint32_t)((iu[1] + 1) & 0xff) << 8) | ((uint32_t)(iu[2] & 0xff) << 16) | ((uint32_t)(iu[3] & 0xff) << 24);
            u[UW_W1 * cap + i] = ((uint32_t)iu[5] & 0xffffu) | ((uint32_t)iu[4] << 16);
            u[UW_ID * cap + i] = (uint32_t)iu[6];
            maxid = std::max<long long>(maxid, iu[6]);
            u[UW_A0 * cap + i] = AT_IDLE | (0xFFu << 8);
            if (iu[7] && in->actions) {
                const int32_t *ia = in->actions + (g * cap + i) * 8;
                uint32_t a0 = (uint32_t)(ia[0] & 0xF) | ((uint32_t)((ia[4] < 0 ? 0xFF : ia[4]) & 0xff) << 8);
                if (ia[0] == MRTS_ATTACK) a0 |= ((uint32_t)(ia[2] & 0xff) << 16) | ((uint32_t)(ia[3] & 0xff) << 24);
                u[UW_A0 * cap + i] = a0; u[UW_A1 * cap + i] = (uint32_t)ia[1]; u[UW_TIS * cap + i] = (uint32_t)ia[5];
                order.emplace_back(ia[6], i);
            }
        }
        std::sort(order.begin(), order.end());
        for (size_t r = 0; r < order.size(); r++) u[UW_SEQ * cap + order[r].second] = (uint32_t)r;
        h[H_NEXTSEQ] = (int32_t)order.size();
        if (h[H_NEXTID] <= maxid) h[H_NEXTID] = (int32_t)(maxid + 1);
        for (int k = 0; k < 3; k++) {
            uint64_t s = in->rng ? (uint64_t)in->rng[g * 3 + k] : jr_scramble((first + g) ^ (k == 1 ? SEED_XOR_CONFLICT : (k == 2 ? SEED_XOR_DAMAGE : 0)));
            h[H_RNGP_LO + 2 * k] = (int32_t)(uint32_t)s; h[H_RNGP_HI + 2 * k] = (int32_t)(uint32_t)(s >> 32);
        }
    }
    if (count && (dev_h2d(b->d_hdr + first * MRTS_HDR_WORDS, hdr.data(), hdr.size() * 4, b->stream) ||
                  dev_h2d(b->d_units + first * (long long)b->uw * cap, un.data(), un.size() * 4, b->stream) || dev_sync(b->stream)))
        return fail(MRTS_E_CUDA, dev_errstr());
    return MRTS_OK;
}

// ---- host half of getPlayerActions / PlayerActionGenerator (player_actions.hpp) ----------------------------------------------
static int fetch_views(mrts_batch *b, int player, int none_duration, int64_t first, int64_t count, std::vector<HView> &out) {
    // unit action lists of games [first, first + count) of the batch as host views.  The kernel runs over the whole batch into device
    // staging; the headers come back first, and of the (sparse) choice / list / position arrays only the leading part any game uses
    const int K = b->cap, MA = 64;
    const size_t nh = (size_t)b->n * 8, np = (size_t)b->n * b->cap, nc = (size_t)b->n * K * 4, nl = (size_t)b->n * K * MA;
    if (dev_select(b->device)) return fail(MRTS_E_CUDA, dev_errstr());
    if (ensure_tmp(b, (nh + np + nc + nl) * 4)) return fail(MRTS_E_CUDA, std::string("staging allocation failed: ") + dev_errstr());
    int32_t *t = (int32_t *)b->d_tmp, *d_hdr = t, *d_pos = t + nh, *d_ch = t + nh + np, *d_ls = t + nh + np + nc;
    int rc = mrts_batch_unit_actions(b, player, none_duration, K, MA, d_hdr, d_pos, d_ch, d_ls, 1);
    if (rc) return rc;
    std::vector<int32_t> hdr((size_t)count * 8);
    if (dev_d2h(hdr.data(), d_hdr + first * 8, hdr.size() * 4, b->stream)) return fail(MRTS_E_CUDA, dev_errstr());
    int kmax = 1, pmax = 1;
    for (int64_t g = 0; g < count; g++) { kmax = std::max(kmax, std::min(hdr[g * 8], K)); pmax = std::max(pmax, std::min(hdr[g * 8 + 5], b->cap)); }
    std::vector<int32_t> pos((size_t)count * pmax), ch((size_t)count * kmax * 4), ls((size_t)count * kmax * MA);
#ifdef MRTS_EMU
    for (int64_t g = 0; g < count; g++) {
        memcpy(&pos[g * pmax], d_pos + (first + g) * b->cap, (size_t)pmax * 4);
        memcpy(&ch[g * kmax * 4], d_ch + (first + g) * K * 4, (size_t)kmax * 16);
        memcpy(&ls[g * (size_t)kmax * MA], d_ls + (first + g) * (size_t)K * MA, (size_t)kmax * MA * 4);
    }
#else
    if (ck(cudaMemcpy2DAsync(pos.data(), (size_t)pmax * 4, d_pos + first * b->cap, (size_t)b->cap * 4, (size_t)pmax * 4, (size_t)count, cudaMemcpyDeviceToHost, b->stream)) ||
        ck(cudaMemcpy2DAsync(ch.data(), (size_t)kmax * 16, d_ch + first * K * 4, (size_t)K * 16, (size_t)kmax * 16, (size_t)count, cudaMemcpyDeviceToHost, b->stream)) ||
        ck(cudaMemcpy2DAsync(ls.data(), (size_t)kmax * MA * 4, d_ls + first * (size_t)K * MA, (size_t)K * MA * 4, (size_t)kmax * MA * 4, (size_t)count, cudaMemcpyDeviceToHost, b->stream)) ||
        dev_sync(b->stream))
        return fail(MRTS_E_CUDA, dev_errstr());
#endif
    out.resize((size_t)count);
    std::vector<uint8_t> bad((size_t)count, 0);
    parallel_for((int)count, [&](int g) {
        if (!decode_view(&hdr[(size_t)g * 8], &pos[(size_t)g * pmax], &ch[(size_t)g * kmax * 4], &ls[(size_t)g * kmax * MA], kmax, MA, b->W, none_duration, out[g])) bad[g] = 1;
    });
    for (int64_t g = 0; g < count; g++) if (bad[g]) return fail(MRTS_E_LIMIT, "a unit has more than 64 legal actions");
    return MRTS_OK;
}

struct mrts_pag { PlayerActionGenerator g; };

int mrts_pag_create(mrts_batch *b, int64_t game, int player, int none_duration, mrts_pag **out) {
    if (!b || !out || game < 0 || game >= b->n || player < 0 || player > 1) return fail(MRTS_E_ARG, "mrts_pag_create: bad argument");
    std::vector<HView> v;
    int rc = fetch_views(b, player, none_duration, game, 1, v);
    if (rc) return rc;
    auto p = std::make_unique<mrts_pag>();
    if (!p->g.init(v[0], b->utt)) return fail(MRTS_E_STATE, "Move generator created with no units that can execute actions");
    *out = p.release();
    return MRTS_OK;
}
void mrts_pag_destroy(mrts_pag *p) { delete p; }
int64_t mrts_pag_size(const mrts_pag *p) { return p ? p->g.size : MRTS_E_ARG; }
int64_t mrts_pag_generated(const mrts_pag *p) { return p ? p->g.generated : MRTS_E_ARG; }
int mrts_pag_num_choices(const mrts_pag *p) { return p ? (int)p->g.view.choices.size() : MRTS_E_ARG; }
static int rows_out(const HView &v, const HPlayerAction &pa, int32_t *rows, int max_k) {
    if ((int)pa.size() > max_k) return fail(MRTS_E_ARG, "max_k is smaller than the PlayerAction");
    for (size_t k = 0; k < pa.size(); k++) raw_row(v, pa[k], rows + k * 8);
    return (int)pa.size();
}
int mrts_pag_next(mrts_pag *p, int32_t *rows, int max_k) {
    if (!p || !rows) return fail(MRTS_E_ARG, "mrts_pag_next: null argument");
    HPlayerAction pa;
    if (!p->g.next(pa)) return MRTS_PAG_DONE;
    return rows_out(p->g.view, pa, rows, max_k);
}
int mrts_pag_random(mrts_pag *p, int64_t *rng_state, int32_t *rows, int max_k) {
    if (!p || !rows || !rng_state) return fail(MRTS_E_ARG, "mrts_pag_random: null argument");
    JavaRandom r; r.s = (uint64_t)*rng_state;
    HPlayerAction pa; p->g.random(r, pa);
    *rng_state = (int64_t)r.s;
    return rows_out(p->g.view, pa, rows, max_k);
}
int mrts_pag_randomize_order(mrts_pag *p, int64_t *rng_state) {
    if (!p || !rng_state) return fail(MRTS_E_ARG, "mrts_pag_randomize_order: null argument");
    JavaRandom r; r.s = (uint64_t)*rng_state;
    p->g.randomize_order(r);
    *rng_state = (int64_t)r.s;
    return MRTS_OK;
}
int64_t mrts_java_random_seed(int64_t seed) { return (int64_t)jr_scramble(seed); }

int mrts_batch_player_actions(mrts_batch *b, int64_t game, int player, int32_t *out_rows, int32_t *out_counts, int64_t max_player_actions, int max_k, int64_t *out_total) {
    if (!b || !out_total || game < 0 || game >= b->n || player < 0 || player > 1 || max_player_actions < 0 || (max_player_actions && (!out_rows || !out_counts)))
        return fail(MRTS_E_ARG, "mrts_batch_player_actions: bad argument");
    std::vector<HView> v;
    int rc = fetch_views(b, player, 10, game, 1, v);
    if (rc) return rc;
    std::vector<HPlayerAction> l;
    *out_total = player_actions(v[0], b->utt, l, max_player_actions);
    for (size_t i = 0; i < l.size(); i++) {
        int n = rows_out(v[0], l[i], out_rows + i * (size_t)max_k * 8, max_k);
        if (n < 0) return n;
        out_counts[i] = n;
    }
    return MRTS_OK;
}

// ---- NaiveMCTS over the batch (mcts.hpp) -----------------------------------------------------------------------------------
struct mrts_mcts {
    NaiveMctsHost H;
    mrts_batch *pool = nullptr, *work = nullptr;
    int T = 0, Tpad = 0, max_nodes = 0;
    std::vector<int64_t> idx; std::vector<uint8_t> mask; std::vector<int32_t> rows, counts; std::vector<float> ev; std::vector<int32_t> tm; std::vector<int64_t> seeds;
    int64_t slot(int t, int node) const { return (int64_t)node * Tpad + t; }
};
void mrts_mcts_destroy(mrts_mcts *m) { if (!m) return; mrts_batch_destroy(m->pool); mrts_batch_destroy(m->work); delete m; }

// node constructor for the work batch's games selected by `which`: the cycle loop, then the move generator's lists; attaches the nodes
static int mcts_finish_nodes(mrts_mcts *m, const std::vector<uint8_t> &which, bool roots) {
    int rc = mrts_batch_cycle_to_decision(m->work);
    if (rc) return rc;
    std::vector<HView> v[2];
    for (int pl = 0; pl < 2; pl++) { rc = fetch_views(m->work, pl, 10, 0, m->T, v[pl]); if (rc) return rc; }
    const int maxp = m->H.player, minp = 1 - maxp;
    parallel_for(m->T, [&](int t) {
        if (!which[t]) return;
        MTree &tr = m->H.trees[t];
        const HView &a = v[maxp][t];
        int type = -1; const HView *view = nullptr;
        if (a.winner != -1 || a.gameover) type = -1;
        else if (a.can[maxp]) { type = 0; view = &v[maxp][t]; }
        else if (a.can[minp]) { type = 1; view = &v[minp][t]; }
        if (roots) {
            MNode nn;
            m->H.init_node(tr, nn, type, a.time, view);
            tr.nodes.clear(); tr.nodes.push_back(std::move(nn)); tr.root_time = a.time;
            m->idx[t] = m->slot(t, 0);
        } else {
            int id = m->H.attach_new_node(tr, type, a.time, view);
            m->idx[t] = m->slot(t, id);
        }
    });
    for (int t = 0; t < m->T; t++) if (!which[t]) m->idx[t] = -1;
    return mrts_batch_scatter_games(m->pool, m->work, m->idx.data(), 0);
}

int mrts_mcts_create(mrts_batch *roots, int player, const mrts_mcts_params *prm, int max_nodes_per_tree, const int64_t *seeds, mrts_mcts **out) {
    if (!roots || !prm || !out || player < 0 || player > 1 || max_nodes_per_tree < 2) return fail(MRTS_E_ARG, "mrts_mcts_create: bad argument");
    if (roots->utt.conflict != MRTS_CANCEL_BOTH) return fail(MRTS_E_STATE, "the rollout kernel implements CANCEL_BOTH only");
    if (roots->scripted || (roots->flags & (MRTS_FLAG_PARTIAL_OBS | MRTS_FLAG_PO_POLICIES))) return fail(MRTS_E_STATE, "searches run on plain batches (no scripted-policy words, fully observable)");
    auto m = std::unique_ptr<mrts_mcts, void (*)(mrts_mcts *)>(new mrts_mcts, mrts_mcts_destroy);
    const int T = (int)roots->n, nm = roots->n_maps;
    m->T = T; m->Tpad = (T + nm - 1) / nm * nm; m->max_nodes = max_nodes_per_tree;
    m->H.P.lookahead = prm->lookahead; m->H.P.max_depth = prm->max_depth; m->H.P.e_l = prm->epsilon_l; m->H.P.e_g = prm->epsilon_g; m->H.P.e_0 = prm->epsilon_0;
    m->H.P.strategy = prm->global_strategy; m->H.P.fensa = prm->force_exploration; m->H.P.eval_fn = prm->eval_fn;
    m->H.algorithm = prm->algorithm ? 1 : 0;
    m->H.player = player; m->H.utt = roots->utt;
    mrts_utt u; u.h = roots->utt;
    std::vector<mrts_map> mh(nm); std::vector<const mrts_map *> mp(nm);
    for (int i = 0; i < nm; i++) { mh[i].h = roots->maps_h[i]; mp[i] = &mh[i]; }
    int rc = mrts_batch_create(&u, mp.data(), nm, (int64_t)m->Tpad * max_nodes_per_tree, roots->device, 0, roots->cap, &m->pool);
    if (rc) return rc;
    rc = mrts_batch_create(&u, mp.data(), nm, T, roots->device, 0, roots->cap, &m->work);
    if (rc) return rc;
    m->H.trees.resize(T);
    for (int t = 0; t < T; t++) {
        MTree &tr = m->H.trees[t];
        tr.seed = seeds ? seeds[t] : t; tr.r = JavaRandom(tr.seed); tr.sampler = JavaRandom(tr.seed ^ 0x2545F4914F6CDD1DLL);
    }
    m->idx.assign(T, -1); m->mask.assign(T, 0); m->counts.assign(T, 0); m->ev.assign(T, 0); m->tm.assign(T, 0); m->seeds.assign(T, 0);
    rc = mrts_batch_copy_games(m->work, roots, nullptr, nullptr, 0); // startNewComputation(player, gs.clone())
    if (rc) return rc;
    rc = mcts_finish_nodes(m.get(), std::vector<uint8_t>(T, 1), true);
    if (rc) return rc;
    *out = m.release();
    return MRTS_OK;
}

int mrts_mcts_iterate(mrts_mcts *m, int n_iterations) {
    if (!m || n_iterations < 0) return fail(MRTS_E_ARG, "mrts_mcts_iterate: bad argument");
    const int T = m->T;
    for (int it = 0; it < n_iterations; it++) {
        // 1. selectLeaf on the host for every search; a search either lands on an existing node or asks for a new one
        int max_k = 1, creating = 0;
        parallel_for(T, [m](int t) {
            MTree &tr = m->H.trees[t];
            tr.creating = false; tr.leaf = -1;
            if (m->H.algorithm == 0) m->H.select_leaf(tr, 0); else m->H.select_leaf_uct(tr, 0);
        });
        for (int t = 0; t < T; t++) {
            MTree &tr = m->H.trees[t];
            if (tr.creating) {
                if ((int)tr.nodes.size() >= m->max_nodes) return fail(MRTS_E_LIMIT, "a search tree outgrew max_nodes_per_tree");
                creating++; max_k = std::max(max_k, (int)tr.new_pa.size());
            }
        }
        // 2. the new nodes: clone the parent's state, issue the sampled PlayerAction (gs.cloneIssue), run the node's constructor
        if (creating) {
            std::vector<uint8_t> which(T, 0);
            for (int t = 0; t < T; t++) { MTree &tr = m->H.trees[t]; which[t] = tr.creating; m->idx[t] = tr.creating ? m->slot(t, tr.new_parent) : 0; }
            int rc = mrts_batch_copy_games(m->work, m->pool, m->idx.data(), which.data(), 0);
            if (rc) return rc;
            for (int pl = 0; pl < 2; pl++) { // the acting player of a node: the searching player at max nodes, the opponent at min nodes
                m->rows.assign((size_t)T * max_k * 8, 0); m->counts.assign(T, 0);
                bool any = false;
                for (int t = 0; t < T; t++) {
                    MTree &tr = m->H.trees[t];
                    if (!tr.creating) continue;
                    const MNode &par = tr.nodes[tr.new_parent];
                    if ((par.type == 0 ? m->H.player : 1 - m->H.player) != pl) continue;
                    any = true; m->counts[t] = (int32_t)tr.new_pa.size();
                    for (size_t k = 0; k < tr.new_pa.size(); k++) raw_row(par.view, tr.new_pa[k], &m->rows[((size_t)t * max_k + k) * 8]);
                }
                if (any) { rc = mrts_batch_issue(m->work, pl, MRTS_ACTIONS_RAW, m->rows.data(), m->counts.data(), max_k, -1, 0, 0); if (rc) return rc; }
            }
            rc = mcts_finish_nodes(m, which, false);
            if (rc) return rc;
        }
        // 3. one playout per search from its leaf (simulate + evaluate on the device)
        for (int t = 0; t < T; t++) { MTree &tr = m->H.trees[t]; m->idx[t] = m->slot(t, tr.leaf); m->seeds[t] = tr.seed * 1000003LL + tr.runs; }
        int rc = mrts_batch_copy_games(m->work, m->pool, m->idx.data(), nullptr, 0);
        if (rc) return rc;
        rc = mrts_batch_rollout(m->work, 1, m->H.P.lookahead, m->H.P.eval_fn, m->H.player, -1, m->seeds.data(), m->ev.data(), m->tm.data(), 0);
        if (rc) return rc;
        // 4. propagateEvaluation
        parallel_for(T, [m](int t) {
            MTree &tr = m->H.trees[t];
            int time = tr.nodes[tr.leaf].time + m->tm[t] - tr.root_time;
            double evaluation = (double)m->ev[t] * std::pow(0.99, time / 10.0);
            m->H.propagate(tr, tr.leaf, evaluation);
            tr.runs++;
        });
    }
    return MRTS_OK;
}

int mrts_mcts_num_nodes(const mrts_mcts *m, int64_t tree) { return (m && tree >= 0 && tree < m->T) ? (int)m->H.trees[tree].nodes.size() : MRTS_E_ARG; }
int mrts_mcts_root(const mrts_mcts *m, int64_t tree, int32_t *root_visits, double *root_accum, int32_t *child_visits, double *child_accum, int max_children) {
    if (!m || tree < 0 || tree >= m->T || !root_visits || !root_accum) return fail(MRTS_E_ARG, "mrts_mcts_root: bad argument");
    const MTree &tr = m->H.trees[tree]; const MNode &root = tr.nodes[0];
    *root_visits = root.visits; *root_accum = root.accum;
    for (int i = 0; i < (int)root.children.size() && i < max_children; i++) { child_visits[i] = tr.nodes[root.children[i]].visits; child_accum[i] = tr.nodes[root.children[i]].accum; }
    return (int)root.children.size();
}
int mrts_mcts_best_actions(const mrts_mcts *m, int32_t *out_rows, int32_t *out_counts, int max_k) {
    if (!m || !out_rows || !out_counts || max_k < 1) return fail(MRTS_E_ARG, "mrts_mcts_best_actions: bad argument");
    for (int t = 0; t < m->T; t++) {
        const MTree &tr = m->H.trees[t];
        int best = m->H.most_visited(tr);
        out_counts[t] = 0;
        if (best < 0) continue; // no children: the empty PlayerAction
        int n = rows_out(tr.nodes[0].view, tr.nodes[0].pas[best], out_rows + (size_t)t * max_k * 8, max_k);
        if (n < 0) return n;
        out_counts[t] = n;
    }
    return MRTS_OK;
}

} // extern "C"
