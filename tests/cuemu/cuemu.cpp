// cuemu.cpp - fiber runtime of the CPU emulation harness (TEST INFRASTRUCTURE ONLY, see cuda_runtime.h).
// One OS thread; every CUDA thread of the running block is a ucontext fiber scheduled round-robin; a fiber runs
// until it reaches a rendezvous that is not complete yet (block barrier, warp collective) and yields.
#include "cuda_runtime.h"

#include <stdio.h>
#include <vector>

// Context switch: callee-saved registers only (x86-64 SysV), ~2 ns; ucontext's swapcontext makes a sigprocmask system
// call per switch, which dominated the ballot-heavy kernels.
#if defined(__x86_64__)
extern "C" void cuemu_switch(void** save_sp, void* new_sp);
asm(R"(
.text
.globl cuemu_switch
.type cuemu_switch,@function
cuemu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size cuemu_switch,.-cuemu_switch
)");
#else
#error "cuemu: context switch implemented for x86-64 only"
#endif

namespace cuemu {

ThreadState* g_cur = nullptr;
uint3 g_blockIdx = {0, 0, 0};
dim3  g_blockDim(1, 1, 1), g_gridDim(1, 1, 1);

static const size_t STACK_BYTES = 256 * 1024;
static const int    MAX_THREADS = 1024;

struct Fiber {
    void*       sp = nullptr;
    char*       stack = nullptr;
    ThreadState ts;
    bool        done = false;
};

struct Warp {
    unsigned live = 0;
    Coll     colls[4];
};

static std::vector<Fiber> g_fibers;
static std::vector<Warp>  g_warps;
static void*              g_main_sp = nullptr;
static Fiber*             g_fiber = nullptr;
static const std::function<void()>* g_body = nullptr;
static uint64_t g_progress = 0;
static int      g_live = 0, g_bar_arrived = 0;
static uint64_t g_bar_gen = 0;
static std::vector<char> g_dyn_smem;

static void yield() { cuemu_switch(&g_fiber->sp, g_main_sp); }

static void trampoline() {
    (*g_body)();
    Fiber* f = g_fiber;
    f->done = true;
    g_live--;
    g_warps[f->ts.lin >> 5].live &= ~(1u << (f->ts.lin & 31));
    g_progress++;
    cuemu_switch(&f->sp, g_main_sp);
    abort();     // a finished fiber is never resumed
}

int lane_id() { return g_cur->lin & 31; }
void* dyn_smem() { return g_dyn_smem.data(); }

void sync_block() {
    uint64_t gen = g_bar_gen;
    g_bar_arrived++;
    g_progress++;
    for (;;) {
        if (g_bar_gen != gen) return;
        if (g_bar_arrived >= g_live) {      // exited threads count as arrived
            g_bar_arrived = 0;
            g_bar_gen++;
            g_progress++;
            return;
        }
        yield();
    }
}

static Coll* find_coll(Warp& W, unsigned mask) {
    for (Coll& c : W.colls)
        if (c.mask == mask) return &c;
    return nullptr;
}

Coll* coll_enter(unsigned mask, uint64_t value, int pred) {
    Warp& W = g_warps[g_cur->lin >> 5];
    const int lane = g_cur->lin & 31;
    if (!((mask >> lane) & 1u)) {
        fprintf(stderr, "cuemu: lane %d calls a warp collective with mask %08x that does not name it\n", lane, mask);
        abort();
    }
    Coll* c;
    for (;;) {
        c = find_coll(W, mask);
        if (c && c->draining) { yield(); continue; }     // the previous collective on this mask is still being read
        break;
    }
    if (!c) {
        for (Coll& k : W.colls)
            if (k.mask == 0) { c = &k; break; }
        if (!c) { fprintf(stderr, "cuemu: too many concurrent warp collectives with different masks\n"); abort(); }
        c->mask = mask;
        c->arrived = 0;
        c->toread = 0;
        c->draining = false;
    }
    c->slot[lane] = value;
    c->pred[lane] = pred;
    c->arrived |= 1u << lane;
    g_progress++;
    for (;;) {
        if (c->draining) break;
        unsigned need = mask & W.live;
        if ((c->arrived & need) == need) {
            c->draining = true;
            c->toread = c->arrived;
            g_progress++;
            break;
        }
        yield();
    }
    return c;
}

void coll_leave(Coll* c) {
    const int lane = g_cur->lin & 31;
    c->toread &= ~(1u << lane);
    if (c->toread == 0) {
        c->arrived = 0;
        c->draining = false;
        c->mask = 0;
    }
    g_progress++;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const int nt = (int)(block.x * block.y * block.z);
    if (nt <= 0 || nt > MAX_THREADS) { fprintf(stderr, "cuemu: bad block size %d\n", nt); abort(); }
    if (g_fiber) { fprintf(stderr, "cuemu: nested kernel launch\n"); abort(); }
    if ((int)g_fibers.size() < nt) {
        size_t old = g_fibers.size();
        g_fibers.resize(nt);
        for (size_t i = old; i < g_fibers.size(); ++i) g_fibers[i].stack = (char*)aligned_alloc(64, STACK_BYTES);
    }
    g_dyn_smem.assign(smem + 16, 0);
    g_blockDim = block;
    g_gridDim = grid;
    g_body = &body;
    const int nw = (nt + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_blockIdx = {bx, by, bz};
                g_warps.assign(nw, Warp());
                for (Warp& W : g_warps)
                    for (Coll& c : W.colls) c.mask = 0;
                g_live = nt;
                g_bar_arrived = 0;
                for (int t = 0; t < nt; ++t) {
                    Fiber& f = g_fibers[t];
                    f.done = false;
                    f.ts.lin = t;
                    f.ts.tid.x = t % block.x;
                    f.ts.tid.y = (t / block.x) % block.y;
                    f.ts.tid.z = t / (block.x * block.y);
                    g_warps[t >> 5].live |= 1u << (t & 31);
                    // initial frame: six callee-saved register slots, then the entry point as the return address
                    void** top = (void**)(f.stack + STACK_BYTES);
                    top[-1] = nullptr;
                    top[-2] = (void*)trampoline;
                    for (int r = 3; r <= 8; ++r) top[-r] = nullptr;
                    f.sp = (void*)(top - 8);
                }
                int remaining = nt;
                while (remaining > 0) {
                    uint64_t p0 = g_progress;
                    remaining = 0;
                    for (int t = 0; t < nt; ++t) {
                        Fiber& f = g_fibers[t];
                        if (f.done) continue;
                        g_fiber = &f;
                        g_cur = &f.ts;
                        cuemu_switch(&g_main_sp, f.sp);
                        if (!f.done) remaining++;
                    }
                    if (remaining > 0 && g_progress == p0) {
                        fprintf(stderr, "cuemu: deadlock in block (%u,%u,%u): %d threads wait at a barrier/collective that cannot complete\n",
                                bx, by, bz, remaining);
                        abort();
                    }
                }
            }
    g_fiber = nullptr;
    g_cur = nullptr;
    g_body = nullptr;
}

}  // namespace cuemu
