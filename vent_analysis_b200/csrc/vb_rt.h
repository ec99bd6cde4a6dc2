// Host-side runtime shim: the handful of CUDA runtime calls the library makes, with a
// host-memory stand-in for the emulation build (tests/emu, TEST INFRASTRUCTURE ONLY).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "vb_platform.h"

namespace vbrt {

inline thread_local std::string g_err;
inline void set_error(const std::string& s) { g_err = s; }

#if !defined(VB_EMULATE)
typedef cudaStream_t stream_t;

inline bool ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return false;
}
inline bool dev_malloc(void** p, size_t n) { return ok(cudaMalloc(p, n ? n : 1), "cudaMalloc"); }
inline void dev_free(void* p) { if (p) cudaFree(p); }
inline bool host_malloc(void** p, size_t n) { return ok(cudaMallocHost(p, n ? n : 1), "cudaMallocHost"); }
inline void host_free(void* p) { if (p) cudaFreeHost(p); }
inline bool h2d(void* d, const void* h, size_t n, stream_t s) {
    return ok(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync(H2D)");
}
inline bool d2h(void* h, const void* d, size_t n, stream_t s) {
    return ok(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync(D2H)");
}
inline bool dev_memset(void* d, int v, size_t n, stream_t s) { return ok(cudaMemsetAsync(d, v, n, s), "cudaMemsetAsync"); }
inline bool stream_sync(stream_t s) { return ok(cudaStreamSynchronize(s), "cudaStreamSynchronize"); }
inline bool device_sync() { return ok(cudaDeviceSynchronize(), "cudaDeviceSynchronize"); }
inline bool last_launch_ok(const char* what) { return ok(cudaGetLastError(), what); }
// Copies from pageable host memory can leave a benign `invalid argument` in the runtime's
// last-error slot although the call itself returned success; every API call is checked through
// its own return code, so the slot is cleared before a launch and read right after it.
inline void clear_stale_error() { cudaGetLastError(); }
inline bool set_device(int d) { return ok(cudaSetDevice(d), "cudaSetDevice"); }
inline int device_count() { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
inline int sm_count(int dev) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }
inline int max_smem_optin(int dev) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); return n; }
template <class K> inline bool allow_smem(K kernel, int bytes) {
    return ok(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), "cudaFuncSetAttribute(smem)");
}
template <class K> inline int static_smem(K kernel) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) { cudaGetLastError(); return 1024; }
    return (int)fa.sharedSizeBytes;
}
inline bool stream_create(stream_t* s) { return ok(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking), "cudaStreamCreate"); }
// a stream whose kernels the block scheduler prefers over those of the default-priority streams
inline bool stream_create_high(stream_t* s) {
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) { cudaGetLastError(); greatest = 0; }
    return ok(cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, greatest), "cudaStreamCreateWithPriority");
}
inline void stream_destroy(stream_t s) { if (s) cudaStreamDestroy(s); }
typedef cudaEvent_t event_t;
inline bool event_create(event_t* e) { return ok(cudaEventCreate(e), "cudaEventCreate"); }
inline void event_destroy(event_t e) { if (e) cudaEventDestroy(e); }
inline bool event_record(event_t e, stream_t s) { return ok(cudaEventRecord(e, s), "cudaEventRecord"); }
inline bool event_sync(event_t e) { return ok(cudaEventSynchronize(e), "cudaEventSynchronize"); }
inline bool event_elapsed(float* ms, event_t a, event_t b) { return ok(cudaEventElapsedTime(ms, a, b), "cudaEventElapsedTime"); }
inline bool stream_wait(stream_t s, event_t e) { return ok(cudaStreamWaitEvent(s, e, 0), "cudaStreamWaitEvent"); }
#else
typedef void* stream_t;
// "device memory" = shared anonymous mappings: the forked CTAs of an emulated cluster launch see one copy
// (AddressSanitizer builds: the instrumented heap, see vbemu::launch_cluster)
inline bool dev_malloc(void** p, size_t n) {
    if (!VB_EMU_SHARED_DEVMEM) { *p = malloc(n ? n : 1); if (*p) memset(*p, 0xAB, n); return *p != nullptr; }
    const size_t total = n + 256;
    void* m = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (m == MAP_FAILED) { *p = nullptr; return false; }
    *(size_t*)m = total;
    *p = (char*)m + 256;
    memset(*p, 0xAB, n);
    return true;
}
inline void dev_free(void* p) {
    if (!VB_EMU_SHARED_DEVMEM) { free(p); return; }
    if (p) { void* m = (char*)p - 256; munmap(m, *(size_t*)m); }
}
inline bool host_malloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p != nullptr; }
inline void host_free(void* p) { free(p); }
inline bool h2d(void* d, const void* h, size_t n, stream_t) { memcpy(d, h, n); return true; }
inline bool d2h(void* h, const void* d, size_t n, stream_t) { memcpy(h, d, n); return true; }
inline bool dev_memset(void* d, int v, size_t n, stream_t) { memset(d, v, n); return true; }
inline bool stream_sync(stream_t) { return true; }
inline bool device_sync() { return true; }
inline bool last_launch_ok(const char* what) {
    if (!vbemu::g_cluster_failed) return true;
    vbemu::g_cluster_failed = false;
    set_error(std::string(what) + ": a CTA of an emulated cluster launch died");
    return false;
}
inline void clear_stale_error() {}
inline bool set_device(int) { return true; }
inline int device_count() { return 1; }
inline int sm_count(int) { const char* e = getenv("VB_EMU_SMS"); return e ? atoi(e) : 2; }
inline int max_smem_optin(int) { return 232448; }
template <class K> inline bool allow_smem(K, int) { return true; }
template <class K> inline int static_smem(K) { return 0; }
inline bool stream_create(stream_t* s) { *s = nullptr; return true; }
inline bool stream_create_high(stream_t* s) { *s = nullptr; return true; }
inline void stream_destroy(stream_t) {}
typedef void* event_t;
inline bool event_create(event_t* e) { *e = nullptr; return true; }
inline void event_destroy(event_t) {}
inline bool event_record(event_t, stream_t) { return true; }
inline bool event_sync(event_t) { return true; }
inline bool event_elapsed(float* ms, event_t, event_t) { *ms = 0.0f; return true; }
inline bool stream_wait(stream_t, event_t) { return true; }
#endif

}  // namespace vbrt
