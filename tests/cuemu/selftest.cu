// cuemu self-test (TEST INFRASTRUCTURE): each case commits one GPU-only mistake that a synchronous host-memory emulation
// would hide, and reports whether the strict emulator (cuda_runtime.h) exposed it.  Driven by tests/test_emu_selftest.py:
//   selftest <case>   ->  prints "caught" / "missed" (cases that must die by SIGSEGV print the cuemu diagnostic to stderr)
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

__global__ void k_fill(int* p, int n, int v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_copy(const int* a, int* b, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = a[i];
}
__global__ void k_dyn(int* out) {
    extern __shared__ int buf[];
    buf[threadIdx.x] = threadIdx.x;
    __syncthreads();
    if (threadIdx.x == 0) out[0] = buf[blockDim.x - 1];
}
// result depends on which thread of the block runs first (a missing barrier): exposed by comparing thread orders
__global__ void k_racy(int* out) {
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    s = threadIdx.x;                 // no barrier, last writer wins
    __syncthreads();
    if (threadIdx.x == 0) out[0] = s;
}

int main(int argc, char** argv) {
    const char* c = argc > 1 ? argv[1] : "";
    int *d = nullptr, *d2 = nullptr, *pin = nullptr;
    cudaMalloc(&d, 256 * sizeof(int));
    cudaMalloc(&d2, 256 * sizeof(int));
    cudaMallocHost(&pin, 256 * sizeof(int));
    cudaStream_t s1, s2;
    cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    if (!strcmp(c, "ok")) {                         // a correct program passes
        k_fill<<<1, 256, 0, s1>>>(d, 256, 7);
        cudaMemcpyAsync(pin, d, 256 * sizeof(int), cudaMemcpyDeviceToHost, s1);
        cudaStreamSynchronize(s1);
        puts(pin[255] == 7 ? "caught" : "missed");  // "caught" = behaved as expected
    } else if (!strcmp(c, "host_deref")) {          // host code reads a device pointer: must SIGSEGV
        k_fill<<<1, 256, 0, s1>>>(d, 256, 7);
        cudaStreamSynchronize(s1);
        printf("%d\n", d[0]);
        puts("missed");
    } else if (!strcmp(c, "pinned_before_sync")) {  // D2H into pinned memory read before the stream was synchronised
        memset(pin, 0, 1024);
        k_fill<<<1, 256, 0, s1>>>(d, 256, 7);
        cudaMemcpyAsync(pin, d, 256 * sizeof(int), cudaMemcpyDeviceToHost, s1);
        puts(pin[0] == 7 ? "missed" : "caught");
    } else if (!strcmp(c, "pinned_reuse")) {        // a pinned staging buffer rewritten before its queued H2D copy ran
        pin[0] = 1;
        cudaMemcpyAsync(d, pin, sizeof(int), cudaMemcpyHostToDevice, s1);
        pin[0] = 2;
        cudaMemcpyAsync(d2, pin, sizeof(int), cudaMemcpyHostToDevice, s1);
        int h[2];
        cudaMemcpyAsync(&h[0], d, sizeof(int), cudaMemcpyDeviceToHost, s1);
        cudaMemcpyAsync(&h[1], d2, sizeof(int), cudaMemcpyDeviceToHost, s1);
        puts(h[0] == 1 ? "missed" : "caught");
    } else if (!strcmp(c, "missing_wait")) {        // consumer stream never waits for the producer stream
        k_fill<<<1, 256, 0, s1>>>(d, 256, 7);
        k_copy<<<1, 256, 0, s2>>>(d, d2, 256);
        cudaMemcpyAsync(pin, d2, 256 * sizeof(int), cudaMemcpyDeviceToHost, s2);
        cudaStreamSynchronize(s2);
        puts(pin[0] == 7 ? "missed" : "caught");
    } else if (!strcmp(c, "event_wait_ok")) {       // the same with the event dependency in place
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        k_fill<<<1, 256, 0, s1>>>(d, 256, 7);
        cudaEventRecord(e, s1);
        cudaStreamWaitEvent(s2, e, 0);
        k_copy<<<1, 256, 0, s2>>>(d, d2, 256);
        cudaMemcpyAsync(pin, d2, 256 * sizeof(int), cudaMemcpyDeviceToHost, s2);
        cudaStreamSynchronize(s2);
        puts(pin[0] == 7 ? "caught" : "missed");
    } else if (!strcmp(c, "zero_grid")) {           // empty input -> grid of 0 blocks: invalid configuration on a GPU
        k_fill<<<0, 256, 0, s1>>>(d, 0, 7);
        puts(cudaGetLastError() == cudaErrorInvalidConfiguration ? "caught" : "missed");
    } else if (!strcmp(c, "grid_y")) {
        k_fill<<<dim3(1, 70000, 1), 32, 0, s1>>>(d, 0, 7);
        puts(cudaGetLastError() == cudaErrorInvalidConfiguration ? "caught" : "missed");
    } else if (!strcmp(c, "smem_optin")) {          // > 48 KB of dynamic shared memory without the opt-in attribute
        k_dyn<<<1, 64, 100 * 1024, s1>>>(d);
        bool rejected = cudaGetLastError() != cudaSuccess;
        cudaFuncSetAttribute(k_dyn, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        k_dyn<<<1, 64, 100 * 1024, s1>>>(d);
        bool accepted = cudaGetLastError() == cudaSuccess;
        cudaStreamSynchronize(s1);                  // (s1 is non-blocking: the legacy-stream copy below would not wait for it)
        cudaMemcpy(pin, d, sizeof(int), cudaMemcpyDeviceToHost);
        puts(rejected && accepted && pin[0] == 63 ? "caught" : "missed");
    } else if (!strcmp(c, "use_after_free")) {      // kernel touches freed device memory: must SIGSEGV
        cudaFree(d);
        k_fill<<<1, 256, 0, s1>>>(d, 256, 7);
        cudaStreamSynchronize(s1);
        puts("missed");
    } else if (!strcmp(c, "oob")) {                 // kernel writes past the end of an allocation: must SIGSEGV
        k_fill<<<2, 256, 0, s1>>>(d, 512, 7);
        cudaStreamSynchronize(s1);
        puts("missed");
    } else if (!strcmp(c, "racy")) {                // prints the value; the driver compares CUEMU_ORDER=fwd with rev
        k_racy<<<1, 64, 0, s1>>>(d);
        cudaMemcpyAsync(pin, d, sizeof(int), cudaMemcpyDeviceToHost, s1);
        cudaStreamSynchronize(s1);
        printf("%d\n", pin[0]);
    } else if (!strcmp(c, "legacy_stream")) {       // legacy default stream orders against BLOCKING streams only
        cudaStream_t sb;
        cudaStreamCreate(&sb);
        k_fill<<<1, 256, 0, sb>>>(d, 256, 7);
        k_copy<<<1, 256>>>(d, d2, 256);             // legacy stream: implicitly after sb
        cudaMemcpy(pin, d2, 256 * sizeof(int), cudaMemcpyDeviceToHost);
        bool ok = pin[0] == 7;
        k_fill<<<1, 256, 0, s1>>>(d, 256, 9);       // non-blocking stream: NOT ordered with the legacy stream
        k_copy<<<1, 256>>>(d, d2, 256);
        cudaMemcpy(pin, d2, 256 * sizeof(int), cudaMemcpyDeviceToHost);
        puts(ok && pin[0] == 7 ? "caught" : "missed");
    } else if (!strcmp(c, "graph_ok")) {            // capture two kernels, replay twice: arguments frozen at capture time
        cudaGraph_t g;
        cudaGraphExec_t ge;
        int v = 3;
        cudaStreamBeginCapture(s1, cudaStreamCaptureModeThreadLocal);
        k_fill<<<1, 256, 0, s1>>>(d, 256, v);
        k_copy<<<1, 256, 0, s1>>>(d, d2, 256);
        bool ok = cudaStreamEndCapture(s1, &g) == cudaSuccess && cudaGraphInstantiate(&ge, g, 0) == cudaSuccess;
        v = 9;                                      // must not leak into the captured launch
        cudaMemsetAsync(d2, 0, 1024, s1);
        cudaStreamSynchronize(s1);
        cudaMemcpy(pin, d2, 4, cudaMemcpyDeviceToHost);
        ok = ok && pin[0] == 0;                     // capturing did not execute anything
        cudaGraphLaunch(ge, s1);
        cudaMemcpyAsync(pin, d2, 1024, cudaMemcpyDeviceToHost, s1);
        cudaStreamSynchronize(s1);
        puts(ok && pin[255] == 3 ? "caught" : "missed");
    } else if (!strcmp(c, "graph_legacy")) {        // capture on the legacy default stream (torch's default stream!) is an error
        puts(cudaStreamBeginCapture(nullptr, cudaStreamCaptureModeThreadLocal) == cudaErrorStreamCaptureUnsupported ? "caught" : "missed");
    } else if (!strcmp(c, "graph_sync_inside")) {   // a synchronisation / allocation inside the captured region invalidates it
        cudaGraph_t g;
        cudaStreamBeginCapture(s1, cudaStreamCaptureModeThreadLocal);
        k_fill<<<1, 256, 0, s1>>>(d, 256, 1);
        bool e1 = cudaStreamSynchronize(s1) != cudaSuccess;
        int* tmp = nullptr;
        bool e2 = cudaMalloc(&tmp, 64) != cudaSuccess;
        bool e3 = cudaStreamEndCapture(s1, &g) == cudaErrorStreamCaptureInvalidated && g == nullptr;
        puts(e1 && e2 && e3 ? "caught" : "missed");
    } else {
        puts("unknown case");
        return 2;
    }
    return 0;
}
