// cuemu.cpp - fiber runtime of the CPU emulation harness (TEST INFRASTRUCTURE ONLY, see cuda_runtime.h).
// One OS thread; every CUDA thread of the running block is a ucontext fiber scheduled round-robin; a fiber runs
// until it reaches a rendezvous that is not complete yet (block barrier, warp collective) and yields.
#include "cuda_runtime.h"

#include <execinfo.h>
#include <signal.h>
#include <stdio.h>
#include <sys/mman.h>
#include <unistd.h>
#include <algorithm>
#include <deque>
#include <map>
#include <string>
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
// cooperative spin-wait (e.g. a modelled mbarrier wait): let the other threads of the block run.  Counts as progress so that
// the deadlock detector tolerates a few rounds of spinning, but a wait that never ends is still reported.
void fiber_yield() {
    static uint64_t spins = 0;
    if (++spins % 50000000ull == 0) { fprintf(stderr, "cuemu: a thread has been spinning in fiber_yield() for 5e7 rounds - modelled wait never satisfied?\n"); abort(); }
    g_progress++;
    yield();
}
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

// thread / block visiting order (CUEMU_ORDER = fwd | rev | rand[:seed]): a kernel whose result depends on which
// thread of a block (or which block) runs first between two synchronisation points has a race; running the suite
// under more than one order exposes it
static int      g_order = -1;     // 0 fwd, 1 rev, 2 rand
static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static uint64_t rng_next() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17; return g_rng; }
static void order_init() {
    if (g_order >= 0) return;
    const char* e = getenv("CUEMU_ORDER");
    g_order = 0;
    if (e && !strncmp(e, "rev", 3)) g_order = 1;
    if (e && !strncmp(e, "rand", 4)) {
        g_order = 2;
        if (e[4] == ':') g_rng ^= strtoull(e + 5, nullptr, 10) * 0xD1342543DE82EF95ull + 1;
    }
}

static void run_kernel(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const int nt = (int)(block.x * block.y * block.z);
    if (nt <= 0 || nt > MAX_THREADS) { fprintf(stderr, "cuemu: bad block size %d\n", nt); abort(); }
    if (g_fiber) { fprintf(stderr, "cuemu: nested kernel launch\n"); abort(); }
    order_init();
    std::vector<int> perm(nt);
    for (int t = 0; t < nt; ++t) perm[t] = g_order == 1 ? nt - 1 - t : t;
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
    const uint64_t nblocks = (uint64_t)grid.x * grid.y * grid.z;
    const uint64_t block_rot = g_order == 2 && nblocks ? rng_next() % nblocks : 0;
    for (uint64_t bi = 0; bi < nblocks; ++bi) {
            {
                // block order: forward; reverse under CUEMU_ORDER=rev; rotated by a random start under CUEMU_ORDER=rand
                uint64_t b = g_order == 1 ? nblocks - 1 - bi : bi;
                if (g_order == 2) b = (bi + block_rot) % nblocks;
                const unsigned bx = (unsigned)(b % grid.x), by = (unsigned)((b / grid.x) % grid.y), bz = (unsigned)(b / ((uint64_t)grid.x * grid.y));
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
                    if (g_order == 2)
                        for (int t = nt - 1; t > 0; --t) std::swap(perm[t], perm[(int)(rng_next() % (uint64_t)(t + 1))]);
                    for (int pt = 0; pt < nt; ++pt) {
                        const int t = perm[pt];
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
    }
    g_fiber = nullptr;
    g_cur = nullptr;
    g_body = nullptr;
}

// =============================================================================================
// runtime: streams, events, device memory (see the comment block in cuda_runtime.h)
// =============================================================================================
static int  g_strict = -1;
static long g_counters[4] = {0, 0, 0, 0};
static cudaError_t g_last_error = cudaSuccess;

static bool strict() {
    if (g_strict < 0) {
        const char* e = getenv("CUEMU_STRICT");
#if defined(__SANITIZE_ADDRESS__)
        g_strict = 0;                 // the sanitizer build keeps malloc memory (redzones) and synchronous execution
        (void)e;
#else
        g_strict = (e && e[0] == '0') ? 0 : 1;
#endif
    }
    return g_strict == 1;
}

// ---- device memory arena ---------------------------------------------------------------------
struct Alloc { char* map; size_t map_bytes; char* user; size_t bytes; bool live; };
static std::map<char*, Alloc> g_allocs;          // keyed by mapping base
static std::map<void*, size_t> g_pinned;         // cudaMallocHost ranges
static char*  g_arena = nullptr;
static size_t g_arena_cap = (size_t)1 << 40, g_arena_top = 0;
static int    g_dev_depth = 0;                   // > 0 while a device operation executes (arena accessible)
static const size_t PAGE = 4096;

static Alloc* find_alloc(const void* p) {
    auto it = g_allocs.upper_bound((char*)p);
    if (it == g_allocs.begin()) return nullptr;
    --it;
    Alloc& a = it->second;
    return ((char*)p >= a.map && (char*)p < a.map + a.map_bytes + PAGE) ? &a : nullptr;
}

static void segv_handler(int sig, siginfo_t* si, void*) {
    char* addr = (char*)si->si_addr;
    if (g_arena && addr >= g_arena && addr < g_arena + g_arena_cap) {
        Alloc* a = find_alloc(addr);
        char msg[512];
        const char* who = g_dev_depth > 0 ? "a KERNEL / device copy" : "HOST code (outside any kernel or copy)";
        const char* what = !a ? "unallocated device address space"
                         : !a->live ? "FREED device memory"
                         : (addr >= a->user && addr < a->user + a->bytes) ? "device memory"
                         : "out-of-bounds device memory (guard area)";
        int n = snprintf(msg, sizeof msg, "\ncuemu: %s accessed %s at %p", who, what, (void*)addr);
        if (a) n += snprintf(msg + n, sizeof msg - n, " [allocation %p + %zu bytes, offset %ld]", (void*)a->user, a->bytes, (long)(addr - a->user));
        n += snprintf(msg + n, sizeof msg - n, "\n");
        (void)!write(2, msg, n);
        void* bt[48];
        int k = backtrace(bt, 48);
        backtrace_symbols_fd(bt, k, 2);
    }
    signal(sig, SIG_DFL);
    raise(sig);
}

static void arena_init() {
    if (g_arena) return;
    g_arena = (char*)mmap(nullptr, g_arena_cap, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (g_arena == (char*)MAP_FAILED) { perror("cuemu: arena mmap"); abort(); }
    struct sigaction sa;
    memset(&sa, 0, sizeof sa);
    sa.sa_sigaction = segv_handler;
    sa.sa_flags = SA_SIGINFO | SA_NODEFER;
    sigaction(SIGSEGV, &sa, nullptr);
}

static void arena_access(bool on) {
    for (auto& kv : g_allocs)
        if (kv.second.live) mprotect(kv.second.map, kv.second.map_bytes, on ? PROT_READ | PROT_WRITE : PROT_NONE);
}
struct DeviceScope {
    DeviceScope() { if (strict() && g_dev_depth++ == 0) arena_access(true); }
    ~DeviceScope() { if (strict() && --g_dev_depth == 0) arena_access(false); }
};

// ---- streams / events ------------------------------------------------------------------------
struct Stream;
struct Dep { uint64_t stream_id, seq; };
struct Op {
    uint64_t seq;
    std::vector<Dep> deps;
    std::function<void()> fn;      // empty: a pure ordering point (stream-wait)
};
struct Stream {
    uint64_t id = 0;
    bool     nonblocking = false, legacy = false, running = false;
    uint64_t last_seq = 0;         // seq of the last op ever enqueued
    std::deque<Op> q;
    std::vector<std::function<void()>>* capture = nullptr;   // non-null while the stream is being captured into a graph
    bool capture_invalid = false;
};
struct Event { bool recorded = false; uint64_t stream_id = 0, seq = 0; };

static std::map<uint64_t, Stream*> g_streams;    // live streams by id
static uint64_t g_next_stream_id = 1, g_next_seq = 1;
static Stream*  g_legacy = nullptr;

static Stream* stream_of(cudaStream_t h) {
    if (!g_legacy) {
        g_legacy = new Stream();
        g_legacy->id = g_next_stream_id++;
        g_legacy->legacy = true;
        g_streams[g_legacy->id] = g_legacy;
    }
    if (h == nullptr || h == (cudaStream_t)1) return g_legacy;
    Stream* s = (Stream*)h;
    auto it = g_streams.find(s->id);
    if (it == g_streams.end() || it->second != s) { fprintf(stderr, "cuemu: use of an invalid / destroyed stream handle %p\n", (void*)h); abort(); }
    return s;
}

static void run_until(Stream* s, uint64_t seq);
static void run_dep(const Dep& d) {
    auto it = g_streams.find(d.stream_id);
    if (it != g_streams.end()) run_until(it->second, d.seq);
}
static void run_until(Stream* s, uint64_t seq) {
    while (!s->q.empty() && s->q.front().seq <= seq) {
        if (s->running) { fprintf(stderr, "cuemu: circular stream/event dependency (stream %lu)\n", (unsigned long)s->id); abort(); }
        s->running = true;
        Op op = std::move(s->q.front());
        for (const Dep& d : op.deps) run_dep(d);
        s->q.pop_front();
        if (op.fn) { DeviceScope ds; op.fn(); }
        s->running = false;
    }
}

static int g_capturing = 0;          // number of streams in capture mode
// a call that is illegal while a capture is in progress (global / thread-local capture modes): fail it and poison the capture
static bool capture_violation(const char* what) {
    if (!g_capturing) return false;
    fprintf(stderr, "cuemu: %s during stream capture: not permitted (the capture is invalidated)\n", what);
    for (auto& kv : g_streams)
        if (kv.second->capture) kv.second->capture_invalid = true;
    g_last_error = cudaErrorStreamCaptureUnsupported;
    return true;
}

static void enqueue(Stream* s, std::function<void()> fn, std::vector<Dep> deps = {}) {
    if (s->capture) {               // recorded into the graph, not executed
        if (fn) s->capture->push_back(std::move(fn));
        return;
    }
    Op op;
    op.seq = g_next_seq++;
    op.deps = std::move(deps);
    // legacy default stream: ordered after everything queued on blocking streams, and blocking streams after it
    if (s->legacy) {
        for (auto& kv : g_streams) {
            Stream* o = kv.second;
            if (o != s && !o->nonblocking && !o->q.empty()) op.deps.push_back({o->id, o->last_seq});
        }
    } else if (!s->nonblocking && g_legacy && !g_legacy->q.empty()) {
        op.deps.push_back({g_legacy->id, g_legacy->last_seq});
    }
    op.fn = std::move(fn);
    s->last_seq = op.seq;
    s->q.push_back(std::move(op));
    if (!strict()) run_until(s, s->last_seq);
    else g_counters[2]++;
}

static void flush_all() {
    for (;;) {
        Stream* pending = nullptr;
        for (auto& kv : g_streams)
            if (!kv.second->q.empty()) { pending = kv.second; break; }
        if (!pending) break;
        run_until(pending, pending->last_seq);
    }
}

void launch(const char* kernel_text, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, std::function<void()> body);

static std::map<std::string, int> g_func_smem;   // kernel text -> opted-in dynamic shared memory
static std::string squeeze(const char* t) {
    std::string r;
    for (; *t; ++t)
        if (!isspace((unsigned char)*t)) r += *t;
    return r;
}
cudaError_t func_set_attr(const char* call_text, int attr, int value) {
    // call_text = "kernel<targs>, attribute, value": strip the last two top-level arguments
    std::string t = squeeze(call_text);
    int depth = 0, commas = 0;
    size_t cut = t.size();
    for (size_t i = t.size(); i-- > 0;) {
        char ch = t[i];
        if (ch == ')' || ch == ']' || ch == '}') depth++;
        if (ch == '(' || ch == '[' || ch == '{') depth--;
        if (ch == ',' && depth == 0 && ++commas == 2) { cut = i; break; }
    }
    if (attr == cudaFuncAttributeMaxDynamicSharedMemorySize) {
        if (value > 227 * 1024) return g_last_error = cudaErrorInvalidValue;
        g_func_smem[t.substr(0, cut)] = value;
    }
    return cudaSuccess;
}

void launch(const char* kernel_text, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, std::function<void()> body) {
    const uint64_t nt = (uint64_t)block.x * block.y * block.z;
    bool bad = grid.x == 0 || grid.y == 0 || grid.z == 0 || nt == 0 || nt > 1024 || block.x > 1024 || block.y > 1024 || block.z > 64 ||
               grid.x > 2147483647u || grid.y > 65535u || grid.z > 65535u;
    cudaError_t err = bad ? cudaErrorInvalidConfiguration : cudaSuccess;
    if (!bad && smem > 48 * 1024) {
        auto it = g_func_smem.find(kernel_text);
        if (it == g_func_smem.end() || (size_t)it->second < smem) err = cudaErrorInvalidValue;
    }
    if (err != cudaSuccess) {
        g_counters[1]++;
        g_last_error = err;
        if (getenv("CUEMU_VERBOSE"))
            fprintf(stderr, "cuemu: launch of %s rejected: grid (%u,%u,%u) block (%u,%u,%u) smem %zu\n", kernel_text, grid.x, grid.y, grid.z,
                    block.x, block.y, block.z, smem);
        return;
    }
    Stream* s = stream_of(stream);
    enqueue(s, [grid, block, smem, body = std::move(body)]() { g_counters[0]++; run_kernel(grid, block, smem, body); });
}

}  // namespace cuemu

using namespace cuemu;

const char* cudaGetErrorString(cudaError_t e) {
    switch (e) {
        case cudaSuccess: return "no error";
        case cudaErrorInvalidValue: return "invalid argument (cuemu)";
        case cudaErrorMemoryAllocation: return "out of memory (cuemu)";
        case cudaErrorInvalidConfiguration: return "invalid configuration argument (cuemu)";
        default: return "cuemu error";
    }
}
cudaError_t cudaGetLastError() { cudaError_t e = g_last_error; g_last_error = cudaSuccess; return e; }
cudaError_t cudaPeekAtLastError() { return g_last_error; }

cudaError_t cudaMalloc(void** p, size_t bytes) {
    if (capture_violation("cudaMalloc")) return cudaErrorStreamCaptureUnsupported;
    if (!strict()) {
        size_t b = (bytes + 255) & ~(size_t)255;
        *p = aligned_alloc(256, b ? b : 256);
        if (*p) memset(*p, 0xCD, b ? b : 256);   // poison: uninitialised device memory must not look like zeros
        return *p ? cudaSuccess : cudaErrorMemoryAllocation;
    }
    arena_init();
    const size_t user = ((bytes ? bytes : 1) + 255) & ~(size_t)255;
    const size_t map_bytes = (user + PAGE - 1) & ~(PAGE - 1);
    if (g_arena_top + map_bytes + PAGE > g_arena_cap) return g_last_error = cudaErrorMemoryAllocation;
    Alloc a;
    a.map = g_arena + g_arena_top;
    a.map_bytes = map_bytes;
    a.user = a.map + (map_bytes - user);          // right-aligned: the allocation ends at the guard page
    a.bytes = bytes;
    a.live = true;
    g_arena_top += map_bytes + PAGE;              // the page after it is never made accessible
    mprotect(a.map, map_bytes, PROT_READ | PROT_WRITE);
    memset(a.map, 0xCD, map_bytes);
    if (g_dev_depth == 0) mprotect(a.map, map_bytes, PROT_NONE);
    g_allocs[a.map] = a;
    *p = a.user;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
    if (!p) return cudaSuccess;
    if (capture_violation("cudaFree")) return cudaErrorStreamCaptureUnsupported;
    if (!strict()) { free(p); return cudaSuccess; }
    g_counters[3]++;
    flush_all();                                   // cudaFree synchronises the device
    Alloc* a = find_alloc(p);
    if (!a || a->user != (char*)p || !a->live) { fprintf(stderr, "cuemu: cudaFree of %p which is not a live device allocation\n", p); abort(); }
    a->live = false;
    mprotect(a->map, a->map_bytes, PROT_NONE);
    madvise(a->map, a->map_bytes, MADV_DONTNEED);  // the address range is never reused: later accesses fault
    return cudaSuccess;
}
cudaError_t cudaMallocHost(void** p, size_t bytes) {
    size_t b = (bytes + 255) & ~(size_t)255;
    *p = aligned_alloc(256, b ? b : 256);
    if (!*p) return cudaErrorMemoryAllocation;
    memset(*p, 0xCD, b ? b : 256);
    g_pinned[*p] = b ? b : 256;
    return cudaSuccess;
}
cudaError_t cudaFreeHost(void* p) {
    if (!p) return cudaSuccess;
    flush_all();
    g_pinned.erase(p);
    free(p);
    return cudaSuccess;
}
static bool is_pinned(const void* p) {
    auto it = g_pinned.upper_bound((void*)p);
    if (it == g_pinned.begin()) return false;
    --it;
    return (const char*)p < (const char*)it->first + it->second;
}

// copies.  Async H2D from pinned memory reads its source when it RUNS; from pageable memory the source is staged at the
// call (as the driver does).  Async D2H into pageable memory is synchronous with the host; into pinned memory it is not.
static cudaError_t copy_async(void* d, const void* s, size_t n, cudaMemcpyKind k, Stream* st, bool blocking) {
    if (n == 0) return cudaSuccess;
    if (k == cudaMemcpyHostToHost) { memmove(d, s, n); return cudaSuccess; }
    const bool src_host = k == cudaMemcpyHostToDevice || (k == cudaMemcpyDefault && is_pinned(s));
    const bool dst_host = k == cudaMemcpyDeviceToHost || (k == cudaMemcpyDefault && is_pinned(d));
    if (src_host && !is_pinned(s) && st->capture) {
        capture_violation("cudaMemcpyAsync from pageable host memory");
        return cudaErrorStreamCaptureUnsupported;
    }
    if (src_host && !is_pinned(s)) {
        std::vector<char> stage((const char*)s, (const char*)s + n);
        enqueue(st, [d, n, stage = std::move(stage)]() { memcpy(d, stage.data(), n); });
    } else {
        enqueue(st, [d, s, n]() { memmove(d, s, n); });
    }
    if (blocking || (dst_host && !is_pinned(d))) {
        if (capture_violation("a copy that is synchronous with the host")) return cudaErrorStreamCaptureUnsupported;
        g_counters[3]++;
        run_until(st, st->last_seq);
    }
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind k) { return copy_async(d, s, n, k, stream_of(nullptr), true); }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t st) { return copy_async(d, s, n, k, stream_of(st), false); }
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind k, cudaStream_t st) {
    for (size_t r = 0; r < h; ++r) copy_async((char*)d + r * dp, (const char*)s + r * sp, w, k, stream_of(st), false);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st) {
    if (n) enqueue(stream_of(st), [d, v, n]() { memset(d, v, n); });
    return cudaSuccess;
}
cudaError_t cudaMemset(void* d, int v, size_t n) {
    Stream* s = stream_of(nullptr);
    if (n) enqueue(s, [d, v, n]() { memset(d, v, n); });
    run_until(s, s->last_seq);
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t st) { if (capture_violation("cudaStreamSynchronize")) return cudaErrorStreamCaptureUnsupported; Stream* s = stream_of(st); g_counters[3]++; run_until(s, s->last_seq); return cudaSuccess; }
cudaError_t cudaStreamQuery(cudaStream_t st) { return stream_of(st)->q.empty() ? cudaSuccess : cudaErrorNotReady; }
cudaError_t cudaDeviceSynchronize() { if (capture_violation("cudaDeviceSynchronize")) return cudaErrorStreamCaptureUnsupported; g_counters[3]++; flush_all(); return cudaSuccess; }
static cudaError_t stream_create(cudaStream_t* out, unsigned flags) {
    stream_of(nullptr);
    Stream* s = new Stream();
    s->id = g_next_stream_id++;
    s->nonblocking = (flags & cudaStreamNonBlocking) != 0;
    g_streams[s->id] = s;
    *out = (cudaStream_t)s;
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned flags, int) { return stream_create(s, flags); }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags) { return stream_create(s, flags); }
cudaError_t cudaStreamCreate(cudaStream_t* s) { return stream_create(s, 0); }
cudaError_t cudaStreamDestroy(cudaStream_t h) {
    Stream* s = stream_of(h);
    if (s->legacy) return cudaErrorInvalidValue;
    run_until(s, s->last_seq);        // the driver lets queued work finish; nothing may refer to the handle afterwards
    g_streams.erase(s->id);
    delete s;
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t) new Event(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t) new Event(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete (Event*)e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t eh, cudaStream_t st) {
    Event* e = (Event*)eh;
    Stream* s = stream_of(st);
    if (s->legacy) enqueue(s, nullptr);           // the legacy stream's record also covers the blocking streams
    e->recorded = true;
    e->stream_id = s->id;
    e->seq = s->last_seq;                         // everything queued on the stream so far
    return cudaSuccess;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t eh, unsigned) {
    Event* e = (Event*)eh;
    if (!e->recorded) return cudaSuccess;         // CUDA semantics: waiting on a never-recorded event is a no-op
    enqueue(stream_of(st), nullptr, {Dep{e->stream_id, e->seq}});
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t eh) {
    Event* e = (Event*)eh;
    g_counters[3]++;
    if (e->recorded) run_dep(Dep{e->stream_id, e->seq});
    return cudaSuccess;
}
cudaError_t cudaEventQuery(cudaEvent_t eh) {
    Event* e = (Event*)eh;
    if (!e->recorded) return cudaSuccess;
    auto it = g_streams.find(e->stream_id);
    if (it == g_streams.end() || it->second->q.empty() || it->second->q.front().seq > e->seq) return cudaSuccess;
    return cudaErrorNotReady;
}
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = 0.f;
    if (cudaEventQuery(a) != cudaSuccess || cudaEventQuery(b) != cudaSuccess) return g_last_error = cudaErrorNotReady;
    return cudaSuccess;
}

struct cuemu_graph_st { std::vector<std::function<void()>> ops; };
struct cuemu_graphexec_st { std::vector<std::function<void()>> ops; };

cudaError_t cudaStreamBeginCapture(cudaStream_t st, cudaStreamCaptureMode) {
    Stream* s = stream_of(st);
    if (s->legacy) {
        fprintf(stderr, "cuemu: cudaStreamBeginCapture on the legacy default stream is not supported by CUDA\n");
        return g_last_error = cudaErrorStreamCaptureUnsupported;
    }
    if (s->capture) return g_last_error = cudaErrorInvalidValue;
    s->capture = new std::vector<std::function<void()>>();
    s->capture_invalid = false;
    g_capturing++;
    return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t st, cudaGraph_t* graph) {
    Stream* s = stream_of(st);
    *graph = nullptr;
    if (!s->capture) return g_last_error = cudaErrorInvalidValue;
    std::vector<std::function<void()>>* ops = s->capture;
    s->capture = nullptr;
    g_capturing--;
    if (s->capture_invalid) { delete ops; return g_last_error = cudaErrorStreamCaptureInvalidated; }
    *graph = new cuemu_graph_st{std::move(*ops)};
    delete ops;
    return cudaSuccess;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* exec, cudaGraph_t graph, unsigned long long) {
    if (!graph) return g_last_error = cudaErrorInvalidValue;
    *exec = new cuemu_graphexec_st{graph->ops};
    return cudaSuccess;
}
cudaError_t cudaGraphLaunch(cudaGraphExec_t exec, cudaStream_t st) {
    if (!exec) return g_last_error = cudaErrorInvalidValue;
    Stream* s = stream_of(st);
    for (const auto& fn : exec->ops) enqueue(s, fn);
    return cudaSuccess;
}
cudaError_t cudaGraphDestroy(cudaGraph_t graph) { delete graph; return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t exec) {
    flush_all();                    // (queued copies of the ops hold their own state; be conservative)
    delete exec;
    return cudaSuccess;
}

extern "C" {
int cuemu_api_return(void* user_stream) {
    Stream* s = stream_of((cudaStream_t)user_stream);
    run_until(s, s->last_seq);
    int left = 0;
    for (auto& kv : g_streams)
        if (!kv.second->q.empty()) left++;
    return left;
}
void cuemu_flush_all() { flush_all(); }
// a host function as a stream operation (used by tests/cuemu/fake_nccl.cpp: communication calls are stream-ordered)
void cuemu_enqueue_host_fn(void* stream, void (*fn)(void*), void* arg) {
    enqueue(stream_of((cudaStream_t)stream), [fn, arg]() { fn(arg); });
}
int  cuemu_strict() { return strict() ? 1 : 0; }
long cuemu_counter(int which) { return which >= 0 && which < 4 ? g_counters[which] : -1; }
}
