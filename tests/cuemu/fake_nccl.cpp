// fake_nccl.cpp - TEST INFRASTRUCTURE: the handful of NCCL entry points the engine's dlopen shim resolves
// (mdgrad_b200/csrc/dist.cu), implemented over POSIX shared memory between PROCESSES that each run the CPU-emulated
// library (tests/cuemu).  Lets the slab-decomposed engine - halo send/recv in place, KE / flag / layer-count all-reduces on a
// side stream, local rebuild - run at world size 2..4 in the GPU-less container (tests/test_emu_dist.py).
//
// Semantics kept from NCCL: calls are STREAM operations (queued on the emulator's stream through cuemu_enqueue_host_fn and
// executed in stream order, i.e. only when the emulated host synchronises or an event dependency pulls them in); the sends
// and receives of one group progress together (all sends are posted before any receive blocks); every rank obtains the same
// bits from an all-reduce (contributions are combined in rank order).  A receive / all-reduce that waits longer than
// FAKE_NCCL_TIMEOUT seconds (default 120) aborts the process with a message instead of hanging the suite.
#include <dlfcn.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>
#include <atomic>
#include <vector>

typedef void* stream_t;
static void (*g_enqueue)(void* stream, void (*fn)(void*), void* arg) = nullptr;

static const int    MAX_RANKS = 4;
static const int    NSLOT = 4;                       // messages in flight per ordered pair
static const size_t MAXMSG = (size_t)4 << 20;        // bytes per message
static const size_t MAXCOLL = 64 * 1024;             // bytes per all-reduce contribution

struct Ring {
    std::atomic<uint64_t> head, tail;                // written / consumed message counts
    size_t size[NSLOT];
    char   data[NSLOT][MAXMSG];
};
struct Shared {
    std::atomic<int> attached;
    std::atomic<uint64_t> written[MAX_RANKS], consumed[MAX_RANKS];     // all-reduce sequence numbers per rank
    char coll[2][MAX_RANKS][MAXCOLL];
    Ring ring[MAX_RANKS][MAX_RANKS];                 // [src][dst]
};
struct Comm {
    Shared*  sh;
    int      rank, world;
    uint64_t coll_seq;                               // all-reduces ENQUEUED so far (stream order = the same on all ranks)
    char     name[64];
};
struct Uid { char b[128]; };

static double now() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static double timeout_s() { const char* e = getenv("FAKE_NCCL_TIMEOUT"); return e ? atof(e) : 120.0; }
template <typename F> static void wait_until(F cond, const char* what, int rank) {
    const double t0 = now();
    for (uint64_t it = 0; !cond(); ++it) {
        if ((it & 1023) == 1023) {
            usleep(50);
            if (now() - t0 > timeout_s()) { fprintf(stderr, "fake_nccl: rank %d waited > %.0f s in %s - peer missing or communication mismatch\n", rank, timeout_s(), what); abort(); }
        }
    }
}
static size_t dtype_size(int dt) { return dt == 8 ? 8 : (dt == 7 || dt == 2 || dt == 3) ? 4 : (dt == 0 || dt == 1) ? 1 : 8; }

struct P2P { Comm* c; int peer; void* buf; size_t bytes; bool send; };
struct GroupOp { std::vector<P2P> ops; };
static int g_group_depth = 0;
static std::vector<std::pair<stream_t, P2P>> g_group;

static bool try_send(const P2P& o) {
    Ring& r = o.c->sh->ring[o.c->rank][o.peer];
    if (o.bytes > MAXMSG) { fprintf(stderr, "fake_nccl: message of %zu bytes exceeds %zu\n", o.bytes, MAXMSG); abort(); }
    uint64_t h = r.head.load(std::memory_order_relaxed);
    if (h - r.tail.load(std::memory_order_acquire) >= (uint64_t)NSLOT) return false;      // ring full: the peer has to receive first
    r.size[h % NSLOT] = o.bytes;
    memcpy(r.data[h % NSLOT], o.buf, o.bytes);
    r.head.store(h + 1, std::memory_order_release);
    return true;
}
static bool try_recv(const P2P& o) {
    Ring& r = o.c->sh->ring[o.peer][o.c->rank];
    uint64_t t = r.tail.load(std::memory_order_relaxed);
    if (r.head.load(std::memory_order_acquire) <= t) return false;
    if (r.size[t % NSLOT] != o.bytes) {
        fprintf(stderr, "fake_nccl: rank %d expected %zu bytes from rank %d, the message has %zu\n", o.c->rank, o.bytes, o.peer, r.size[t % NSLOT]);
        abort();
    }
    memcpy(o.buf, r.data[t % NSLOT], o.bytes);
    r.tail.store(t + 1, std::memory_order_release);
    return true;
}
// All operations of a group progress together (as in NCCL): per (peer, direction) in call order, sends and receives
// interleaved, so that a group with more messages to one peer than the ring holds cannot deadlock against its mirror image.
static void run_group(void* arg) {
    GroupOp* g = (GroupOp*)arg;
    std::vector<char> done(g->ops.size(), 0);
    size_t left = g->ops.size();
    const int rank = left ? g->ops[0].c->rank : -1;
    wait_until([&] {
        for (size_t i = 0; i < g->ops.size(); ++i) {
            if (done[i]) continue;
            const P2P& o = g->ops[i];
            bool first = true;                      // FIFO per (peer, direction)
            for (size_t j = 0; j < i; ++j)
                if (!done[j] && g->ops[j].peer == o.peer && g->ops[j].send == o.send) { first = false; break; }
            if (!first) continue;
            if (o.send ? try_send(o) : try_recv(o)) { done[i] = 1; --left; }
        }
        return left == 0;
    }, "grouped send/recv", rank);
    delete g;
}

struct Coll { Comm* c; const void* send; void* recv; size_t count; int dtype, op; uint64_t seq; };
static void run_allreduce(void* arg) {
    Coll* k = (Coll*)arg;
    Comm* c = k->c;
    Shared* sh = c->sh;
    const size_t bytes = k->count * dtype_size(k->dtype);
    if (bytes > MAXCOLL) { fprintf(stderr, "fake_nccl: all-reduce of %zu bytes exceeds %zu\n", bytes, MAXCOLL); abort(); }
    const int b = (int)(k->seq & 1);
    // the buffer of this parity was last used by collective seq-2: everybody must have consumed it
    if (k->seq >= 2)
        for (int r = 0; r < c->world; ++r)
            wait_until([&] { return sh->consumed[r].load(std::memory_order_acquire) >= k->seq - 1; }, "all-reduce (buffer reuse)", c->rank);
    memcpy(sh->coll[b][c->rank], k->send, bytes);
    sh->written[c->rank].store(k->seq + 1, std::memory_order_release);
    for (int r = 0; r < c->world; ++r)
        wait_until([&] { return sh->written[r].load(std::memory_order_acquire) >= k->seq + 1; }, "all-reduce", c->rank);
    std::vector<char> acc(bytes);
    memcpy(acc.data(), sh->coll[b][0], bytes);
    for (int r = 1; r < c->world; ++r) {                       // rank order: the same bits on every rank
        for (size_t i = 0; i < k->count; ++i) {
            if (k->dtype == 8) { double* a = (double*)acc.data(); const double* x = (const double*)sh->coll[b][r]; a[i] = k->op == 0 ? a[i] + x[i] : (x[i] > a[i] ? x[i] : a[i]); }
            else if (k->dtype == 7) { float* a = (float*)acc.data(); const float* x = (const float*)sh->coll[b][r]; a[i] = k->op == 0 ? a[i] + x[i] : (x[i] > a[i] ? x[i] : a[i]); }
            else if (k->dtype == 2) { int* a = (int*)acc.data(); const int* x = (const int*)sh->coll[b][r]; a[i] = k->op == 0 ? a[i] + x[i] : (x[i] > a[i] ? x[i] : a[i]); }
            else { fprintf(stderr, "fake_nccl: all-reduce dtype %d not implemented\n", k->dtype); abort(); }
        }
    }
    memcpy(k->recv, acc.data(), bytes);
    sh->consumed[c->rank].store(k->seq + 1, std::memory_order_release);
    delete k;
}

static bool resolve() {
    if (!g_enqueue) *(void**)(&g_enqueue) = dlsym(RTLD_DEFAULT, "cuemu_enqueue_host_fn");
    if (!g_enqueue) fprintf(stderr, "fake_nccl: cuemu_enqueue_host_fn not found - load the emulated library with RTLD_GLOBAL first\n");
    return g_enqueue != nullptr;
}

extern "C" {
int ncclGetUniqueId(Uid* id) {
    memset(id->b, 0, sizeof id->b);
    snprintf(id->b, sizeof id->b, "/cuemu_nccl_%d_%lx", (int)getpid(), (unsigned long)(now() * 1e6));
    int fd = shm_open(id->b, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return 2;
    if (ftruncate(fd, sizeof(Shared)) != 0) { close(fd); return 2; }
    close(fd);                                                  // fresh pages are zero: all counters start at 0
    return 0;
}
int ncclCommInitRank(void** comm, int nranks, Uid id, int rank) {
    if (nranks > MAX_RANKS || !resolve()) return 4;
    int fd = shm_open(id.b, O_RDWR, 0600);
    if (fd < 0) return 2;
    void* p = mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return 2;
    Comm* c = new Comm();
    c->sh = (Shared*)p;
    c->rank = rank;
    c->world = nranks;
    c->coll_seq = 0;
    strncpy(c->name, id.b, sizeof c->name - 1);
    c->sh->attached.fetch_add(1);
    wait_until([&] { return c->sh->attached.load() >= nranks; }, "ncclCommInitRank (rendezvous)", rank);
    *comm = c;
    return 0;
}
int ncclCommDestroy(void* comm) {
    Comm* c = (Comm*)comm;
    if (!c) return 0;
    if (c->sh->attached.fetch_sub(1) == 1) shm_unlink(c->name);   // last one out removes the segment
    munmap(c->sh, sizeof(Shared));
    delete c;
    return 0;
}
int ncclGroupStart() { g_group_depth++; return 0; }
static void flush_group() {
    // one queued operation per stream, in call order
    while (!g_group.empty()) {
        stream_t st = g_group.front().first;
        GroupOp* g = new GroupOp();
        std::vector<std::pair<stream_t, P2P>> rest;
        for (auto& e : g_group) (e.first == st ? (void)g->ops.push_back(e.second) : (void)rest.push_back(e));
        g_group.swap(rest);
        g_enqueue(st, run_group, g);
    }
}
int ncclGroupEnd() {
    if (--g_group_depth == 0) flush_group();
    return 0;
}
static int p2p(bool send, void* buf, size_t count, int dtype, int peer, void* comm, stream_t st) {
    Comm* c = (Comm*)comm;
    if (!c || peer < 0 || peer >= c->world) return 4;
    g_group.push_back({st, P2P{c, peer, buf, count * dtype_size(dtype), send}});
    if (g_group_depth == 0) flush_group();
    return 0;
}
int ncclSend(const void* buf, size_t count, int dtype, int peer, void* comm, stream_t st) { return p2p(true, (void*)buf, count, dtype, peer, comm, st); }
int ncclRecv(void* buf, size_t count, int dtype, int peer, void* comm, stream_t st) { return p2p(false, buf, count, dtype, peer, comm, st); }
int ncclAllReduce(const void* send, void* recv, size_t count, int dtype, int op, void* comm, stream_t st) {
    Comm* c = (Comm*)comm;
    if (!c) return 4;
    g_enqueue(st, run_allreduce, new Coll{c, send, recv, count, dtype, op, c->coll_seq++});
    return 0;
}
int ncclBroadcast(const void*, void*, size_t, int, int, void*, stream_t) { return 3; }   // not used by the engine any more
const char* ncclGetErrorString(int r) { return r == 0 ? "ok" : r == 2 ? "system error (shared memory)" : r == 3 ? "not implemented in fake_nccl" : "invalid usage"; }
}
