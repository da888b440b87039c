// cuemu/cuda_runtime.h - TEST INFRASTRUCTURE ONLY.
//
// A host-side stand-in for <cuda_runtime.h> that lets the .cu sources of libmdgrad_b200.so be compiled by g++
// and executed on the CPU, thread by thread, for FUNCTIONAL checks of the kernels (indexing, shared-memory
// protocols, warp collectives, host-side launch logic) in the GPU-less development container.
// Every CUDA thread of a block is a fiber (ucontext); __syncthreads / __shfl_*_sync / __ballot_sync / ... are
// rendezvous points between fibers; blocks run one after the other; "device memory" is host memory.
//
// It is NOT a product path: nothing under mdgrad_b200/ loads the emulated library, it is built by
// tests/cuemu/build_emu.py into tests/cuemu/_build/ and used by tests/test_emu_*.py only.  It says nothing about
// performance and little about memory-model races - it checks that the kernels compute the right thing.
#pragma once
#define MDG_EMU 1
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <functional>
#include <tuple>

// ---------------------------------------------------------------------------------------------
// qualifiers
// ---------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

// ---------------------------------------------------------------------------------------------
// vector types
// ---------------------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct int3 { int x, y, z; };
struct float3 { float x, y, z; };
struct alignas(8) float2 { float x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline float3 make_float3(float x, float y, float z) { float3 r; r.x = x; r.y = y; r.z = z; return r; }
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }

// ---------------------------------------------------------------------------------------------
// the fiber runtime (cuemu.cpp)
// ---------------------------------------------------------------------------------------------
struct cuemu_stream_st;
namespace cuemu {
struct ThreadState {
    uint3 tid;
    int   lin;      // linear thread index in the block
};
extern ThreadState* g_cur;
extern uint3 g_blockIdx;
extern dim3  g_blockDim, g_gridDim;
void  launch(const char* kernel_text, dim3 grid, dim3 block, size_t smem, struct ::cuemu_stream_st* stream, std::function<void()> body);
void  sync_block();
// warp collective: deposit (value, pred) for this lane, wait for all live lanes of `mask`; the returned arrays stay
// valid until coll_leave()
struct Coll {
    unsigned mask, arrived, toread;
    bool     draining;
    uint64_t slot[32];
    int      pred[32];
};
Coll* coll_enter(unsigned mask, uint64_t value, int pred);
void  coll_leave(Coll* c);
void* dyn_smem();
int   lane_id();
void  fiber_yield();
}  // namespace cuemu

#define threadIdx (cuemu::g_cur->tid)
#define blockIdx (cuemu::g_blockIdx)
#define blockDim (cuemu::g_blockDim)
#define gridDim (cuemu::g_gridDim)
#define warpSize 32

static inline void __syncthreads() { cuemu::sync_block(); }
static inline void __threadfence_system() {}
static inline void __threadfence() {}

static inline void __syncwarp(unsigned mask = 0xffffffffu) { cuemu::coll_leave(cuemu::coll_enter(mask, 0, 0)); }

namespace cuemu {
template <typename T> inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle payload must be <= 8 bytes");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T> inline T from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}
template <typename T> inline T shfl_from(unsigned mask, T v, int src_lane) {
    Coll* c = coll_enter(mask, to_bits(v), 0);
    T r = from_bits<T>(c->slot[src_lane & 31]);
    coll_leave(c);
    return r;
}
}  // namespace cuemu

template <typename T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    int lane = cuemu::lane_id();
    int base = lane & ~(width - 1);
    return cuemu::shfl_from(mask, v, base + (src & (width - 1)));
}
template <typename T> inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    int lane = cuemu::lane_id();
    int src = lane ^ lanemask;
    if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
    return cuemu::shfl_from(mask, v, src);
}
template <typename T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    int lane = cuemu::lane_id();
    int src = lane - (int)delta;
    if (src < (lane & ~(width - 1))) src = lane;
    return cuemu::shfl_from(mask, v, src);
}
template <typename T> inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    int lane = cuemu::lane_id();
    int src = lane + (int)delta;
    if (src > (lane | (width - 1))) src = lane;
    return cuemu::shfl_from(mask, v, src);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    cuemu::Coll* c = cuemu::coll_enter(mask, 0, pred ? 1 : 0);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (((c->arrived >> l) & 1u) && c->pred[l]) r |= 1u << l;
    cuemu::coll_leave(c);
    return r;
}
static inline int __all_sync(unsigned mask, int pred) {
    cuemu::Coll* c = cuemu::coll_enter(mask, 0, pred ? 1 : 0);
    int r = 1;
    for (int l = 0; l < 32; ++l)
        if (((c->arrived >> l) & 1u) && !c->pred[l]) r = 0;
    cuemu::coll_leave(c);
    return r;
}
static inline int __any_sync(unsigned mask, int pred) {
    cuemu::Coll* c = cuemu::coll_enter(mask, 0, pred ? 1 : 0);
    int r = 0;
    for (int l = 0; l < 32; ++l)
        if (((c->arrived >> l) & 1u) && c->pred[l]) r = 1;
    cuemu::coll_leave(c);
    return r;
}
static inline unsigned __activemask() { return 0xffffffffu; }

// ---------------------------------------------------------------------------------------------
// intrinsics
// ---------------------------------------------------------------------------------------------
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
template <typename T> inline T __ldcg(const T* p) { return *p; }
template <typename T> inline void __stcs(T* p, T v) { *p = v; }
template <typename T> inline void __stcg(T* p, T v) { *p = v; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
#define __expf(x) expf(x)
#define __logf(x) logf(x)
static inline int   __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline int   __popc(unsigned x) { return __builtin_popcount(x); }
static inline int   __ffs(int x) { return __builtin_ffs(x); }
static inline int   __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int   __float2int_rn(float x) { return (int)rintf(x); }
static inline unsigned __vabsdiffu4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int k = 0; k < 4; ++k) {
        int x = (a >> (8 * k)) & 255, y = (b >> (8 * k)) & 255;
        r |= (unsigned)(x > y ? x - y : y - x) << (8 * k);
    }
    return r;
}
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
    for (int k = 0; k < 4; ++k) c += ((a >> (8 * k)) & 255) * ((b >> (8 * k)) & 255);
    return c;
}

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline long min(long a, long b) { return a < b ? a : b; }
static inline long max(long a, long b) { return a > b ? a : b; }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline double min(double a, double b) { return fmin(a, b); }
static inline double max(double a, double b) { return fmax(a, b); }

// one fiber runs at a time: plain read-modify-write is atomic
template <typename T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline float atomicAdd(float* p, double v) { float o = *p; *p = o + (float)v; return o; }
template <typename T> inline T atomicSub(T* p, T v) { T o = *p; *p = o - v; return o; }
template <typename T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <typename T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T> inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

// ---------------------------------------------------------------------------------------------
// runtime API (cuemu.cpp).  Two modes, chosen at run time by CUEMU_STRICT (default 1):
//   strict: GPU-like ASYNCHRONOUS semantics.  Launches / async copies / memsets are queued per stream and only run
//           when the host synchronises (stream / event / device sync, a blocking copy, cudaFree) - and then only
//           what that synchronisation guarantees (the stream, plus what its event waits and the legacy-default-stream
//           rules drag in).  A host read of a result before its synchronisation, a pinned staging buffer rewritten
//           before its queued copy ran, or a missing cudaStreamWaitEvent therefore FAIL here as they would on a GPU.
//           cudaMalloc memory comes from an arena that is PROT_NONE while no device operation is executing: host
//           code dereferencing a device pointer takes a SIGSEGV with a diagnostic; freed device memory stays
//           inaccessible for good (use after free), every allocation ends at a guard page.
//   relaxed (CUEMU_STRICT=0, and the sanitizer build): the old synchronous behaviour on plain malloc memory.
// Launch configurations are checked like the driver does (zero / oversized grid or block dimensions, dynamic shared
// memory above 48 KB without cudaFuncSetAttribute, above 227 KB at all): the launch is dropped and
// cudaGetLastError() returns cudaErrorInvalidConfiguration / cudaErrorInvalidValue.
// ---------------------------------------------------------------------------------------------
typedef int cudaError_t;
#define cudaSuccess 0
#define cudaErrorInvalidValue 1
#define cudaErrorMemoryAllocation 2
#define cudaErrorInvalidConfiguration 9
#define cudaErrorNotReady 600
typedef struct cuemu_stream_st* cudaStream_t;
typedef struct cuemu_event_st*  cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
#define cudaEventDisableTiming 2
#define cudaEventDefault 0
#define cudaStreamNonBlocking 1
#define cudaStreamDefault 0

struct cudaDeviceProp {
    char name[256];
    int  major, minor, multiProcessorCount;
    size_t totalGlobalMem, sharedMemPerBlockOptin;
};

const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaPeekAtLastError();
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof(*p));
    strcpy(p->name, "cuemu (CPU fibers)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
    p->totalGlobalMem = (size_t)1 << 34;
    p->sharedMemPerBlockOptin = 227 * 1024;
    return cudaSuccess;
}
cudaError_t cudaMalloc(void** p, size_t bytes);
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t bytes) { return cudaMalloc((void**)p, bytes); }
cudaError_t cudaFree(void* p);
cudaError_t cudaMallocHost(void** p, size_t bytes);
template <typename T> inline cudaError_t cudaMallocHost(T** p, size_t bytes) { return cudaMallocHost((void**)p, bytes); }
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t st = nullptr);
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind k, cudaStream_t st = nullptr);
cudaError_t cudaMemset(void* d, int v, size_t n);
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st = nullptr);
cudaError_t cudaStreamSynchronize(cudaStream_t st);
cudaError_t cudaStreamQuery(cudaStream_t st);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned flags, int prio);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamCreate(cudaStream_t* s);
cudaError_t cudaStreamDestroy(cudaStream_t s);
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -5; return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventQuery(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
// stream capture / graphs: ops enqueued on a capturing stream are recorded instead of queued; cudaGraphLaunch queues copies of
// them (kernel arguments frozen at capture time, as on a GPU).  Calls that are illegal during capture (synchronisations,
// cudaMalloc / cudaFree, copies that are synchronous with the host, capture on the legacy stream) fail as the driver does
// and invalidate the capture.
typedef struct cuemu_graph_st*     cudaGraph_t;
typedef struct cuemu_graphexec_st* cudaGraphExec_t;
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1, cudaStreamCaptureModeRelaxed = 2 };
#define cudaErrorStreamCaptureUnsupported 900
#define cudaErrorStreamCaptureInvalidated 901
cudaError_t cudaStreamBeginCapture(cudaStream_t st, cudaStreamCaptureMode mode);
cudaError_t cudaStreamEndCapture(cudaStream_t st, cudaGraph_t* graph);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* exec, cudaGraph_t graph, unsigned long long flags = 0);
cudaError_t cudaGraphLaunch(cudaGraphExec_t exec, cudaStream_t st);
cudaError_t cudaGraphDestroy(cudaGraph_t graph);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t exec);
namespace cuemu {
cudaError_t func_set_attr(const char* call_text, int attr, int value);
template <typename F> inline cudaError_t func_set_attr2(const char* call_text, F, int attr, int value) { return func_set_attr(call_text, attr, value); }
}
#define cudaFuncAttributeMaxDynamicSharedMemorySize 8
// the kernel is identified by its source text (the same text the launch rewriter records), template commas included
#define cudaFuncSetAttribute(...) cuemu::func_set_attr2(#__VA_ARGS__, __VA_ARGS__)

// test hooks (exported, C linkage): what a following operation on the caller's stream would observe / leftovers
extern "C" {
int  cuemu_api_return(void* user_stream);   // synchronise the caller's stream; returns the number of OTHER streams with pending work
void cuemu_flush_all();
void cuemu_enqueue_host_fn(void* stream, void (*fn)(void*), void* arg);
int  cuemu_strict();
long cuemu_counter(int which);              // 0 launches run, 1 launches rejected, 2 ops deferred past their enqueue call, 3 host-blocking synchronisations
}
