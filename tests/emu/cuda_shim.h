// cuda_shim.h -- TEST TOOLING ONLY.  Lets the device code of microrts_b200/csrc/engine.cuh run on the CPU so that
// warp-synchronous logic can be debugged without a GPU.  Every CUDA thread is a ucontext coroutine; warp collectives
// (__shfl_sync, __ballot_sync, __reduce_*_sync, __syncwarp) and __syncthreads are rendezvous points.  Between
// rendezvous points lanes run one at a time in a pseudo-random order that changes every round, so a missing
// __syncwarp() shows up as a parity failure instead of silently passing.  All collectives must be reached by all 32
// lanes of a warp from the same source line (checked).
//
// This is NOT a CPU backend of the product: libmicrorts_cuda.so never contains it, and nothing in microrts_b200/
// can load it.  It is built into tests/emu/libmicrorts_emu.so by tests/emu/build.sh.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace emu {

struct Thread {
    ucontext_t ctx;
    char *stack = nullptr;
    int tid = 0;
    bool done = false;
    int waiting = 0; // 0 runnable, 1 warp rendezvous, 2 block rendezvous
    int site = 0;
    uint64_t val = 0;
};

struct Block {
    std::vector<Thread> th;
    ucontext_t sched;
    int cur = -1;
    int nthreads = 0, bid = 0;
    std::vector<uint64_t> snap; // [nthreads] values of the last completed warp rendezvous
    unsigned char *smem = nullptr;
    std::function<void(unsigned char *, int, int)> body;
    uint32_t rng = 12345;
};

inline Block *&cur_block() { static thread_local Block *b = nullptr; return b; }

inline void thread_entry() {
    Block *b = cur_block();
    Thread &t = b->th[b->cur];
    b->body(b->smem, t.tid, b->bid);
    t.done = true;
    swapcontext(&t.ctx, &b->sched);
}

inline void yield_wait(int kind, int site, uint64_t v) {
    Block *b = cur_block();
    Thread &t = b->th[b->cur];
    t.waiting = kind; t.site = site; t.val = v;
    swapcontext(&t.ctx, &b->sched);
}

inline void run_block(Block &b) {
    cur_block() = &b;
    const size_t STK = 256 * 1024;
    for (int i = 0; i < b.nthreads; i++) {
        Thread &t = b.th[i];
        t.tid = i; t.done = false; t.waiting = 0;
        if (!t.stack) t.stack = (char *)malloc(STK);
        getcontext(&t.ctx);
        t.ctx.uc_stack.ss_sp = t.stack; t.ctx.uc_stack.ss_size = STK; t.ctx.uc_link = &b.sched;
        makecontext(&t.ctx, (void (*)())thread_entry, 0);
    }
    std::vector<int> order(b.nthreads);
    for (;;) {
        for (int i = 0; i < b.nthreads; i++) order[i] = i;
        for (int i = b.nthreads - 1; i > 0; i--) { b.rng = b.rng * 1664525u + 1013904223u; int j = (b.rng >> 8) % (i + 1); std::swap(order[i], order[j]); }
        bool progressed = false, all_done = true;
        for (int oi = 0; oi < b.nthreads; oi++) {
            Thread &t = b.th[order[oi]];
            if (t.done) continue;
            all_done = false;
            if (t.waiting) continue;
            b.cur = order[oi];
            swapcontext(&b.sched, &t.ctx);
            progressed = true;
        }
        if (all_done) break;
        // release completed rendezvous
        int nw = b.nthreads / 32;
        for (int w = 0; w < nw; w++) {
            int cnt = 0, site = -1; bool same = true, any_done = false;
            for (int l = 0; l < 32; l++) {
                Thread &t = b.th[w * 32 + l];
                if (t.done) any_done = true;
                if (t.waiting == 1) { cnt++; if (site < 0) site = t.site; else if (site != t.site) same = false; }
            }
            if (cnt == 32) {
                if (!same) { fprintf(stderr, "emu: divergent warp collective (block %d warp %d)\n", b.bid, w); for (int l = 0; l < 32; l++) fprintf(stderr, " lane %d line %d\n", l, b.th[w * 32 + l].site); abort(); }
                for (int l = 0; l < 32; l++) { b.snap[w * 32 + l] = b.th[w * 32 + l].val; b.th[w * 32 + l].waiting = 0; }
                progressed = true;
            } else if (cnt > 0 && any_done) { fprintf(stderr, "emu: lanes exited while others wait in a collective (line %d)\n", site); abort(); }
        }
        int cb = 0, alive = 0;
        for (auto &t : b.th) { if (!t.done) alive++; if (t.waiting == 2) cb++; }
        if (cb > 0 && cb == alive) { for (auto &t : b.th) if (t.waiting == 2) t.waiting = 0; progressed = true; }
        if (!progressed) {
            fprintf(stderr, "emu: deadlock in block %d\n", b.bid);
            for (auto &t : b.th) if (!t.done) fprintf(stderr, " tid %d waiting %d line %d\n", t.tid, t.waiting, t.site);
            abort();
        }
    }
}

// launch `body(smem, tid, bid)` over a grid; blocks run one after the other
inline void launch(int nblocks, int nthreads, size_t smem_bytes, std::function<void(unsigned char *, int, int)> body) {
    static thread_local Block b;
    if ((int)b.th.size() < nthreads) b.th.resize(nthreads);
    b.nthreads = nthreads; b.snap.assign(nthreads, 0); b.body = body;
    std::vector<unsigned char> smem(smem_bytes + 64);
    b.smem = (unsigned char *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    for (int i = 0; i < nblocks; i++) { b.bid = i; memset(b.smem, 0xCD, smem_bytes); run_block(b); }
}

inline int lane_id() { return cur_block()->cur & 31; }
inline const uint64_t *warp_snap() { Block *b = cur_block(); return &b->snap[(b->cur / 32) * 32]; }

inline unsigned ballot(int pred, int line) {
    yield_wait(1, line, pred ? 1 : 0);
    const uint64_t *s = warp_snap(); unsigned m = 0;
    for (int l = 0; l < 32; l++) if (s[l]) m |= 1u << l;
    return m;
}
template <typename T> inline T shfl(T v, int src, int line) {
    uint64_t bits = 0; memcpy(&bits, &v, sizeof(T));
    yield_wait(1, line, bits);
    T out; uint64_t r = warp_snap()[src & 31]; memcpy(&out, &r, sizeof(T));
    return out;
}
inline unsigned reduce_min_u(unsigned v, int line) {
    yield_wait(1, line, v);
    const uint64_t *s = warp_snap(); unsigned m = 0xFFFFFFFFu;
    for (int l = 0; l < 32; l++) if ((unsigned)s[l] < m) m = (unsigned)s[l];
    return m;
}
inline int reduce_add_i(int v, int line) {
    yield_wait(1, line, (uint64_t)(uint32_t)v);
    const uint64_t *s = warp_snap(); int a = 0;
    for (int l = 0; l < 32; l++) a += (int)(uint32_t)s[l];
    return a;
}
inline unsigned match_any(long long v, int line) {
    yield_wait(1, line, (uint64_t)v);
    const uint64_t *s = warp_snap(); unsigned m = 0;
    for (int l = 0; l < 32; l++) if (s[l] == (uint64_t)v) m |= 1u << l;
    return m;
}
inline void syncwarp(int line) { yield_wait(1, line, 0); }
inline void syncthreads(int line) { yield_wait(2, line, 0); }

} // namespace emu

#define __shfl_sync(m, v, src) emu::shfl((v), (src), __LINE__)
#define __ballot_sync(m, p) emu::ballot((p) ? 1 : 0, __LINE__)
#define __reduce_min_sync(m, v) emu::reduce_min_u((unsigned)(v), __LINE__)
#define __match_any_sync(m, v) emu::match_any((long long)(v), __LINE__)
#define __reduce_add_sync(m, v) emu::reduce_add_i((int)(v), __LINE__)
#define __syncwarp() emu::syncwarp(__LINE__)
#define __syncthreads() emu::syncthreads(__LINE__)
#define __popc(x) __builtin_popcount((unsigned)(x))
#define __ffs(x) __builtin_ffs((int)(x))
#define __clz(x) ((x) ? __builtin_clz((unsigned)(x)) : 32)
#define __dmul_rn(a, b) ((double)(a) * (double)(b))
struct uint4 { unsigned x, y, z, w; };
struct int4 { int x, y, z, w; };
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
static inline int atomicOr(int *p, int v) { int o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
// IEEE single operations without contraction (the emulator is built with -ffp-contract=off)
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
