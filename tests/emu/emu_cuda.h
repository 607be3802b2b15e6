// CUDA-semantics shim used ONLY by the CPU test build under tests/emu/.
//
// It lets the library's real .cu sources (host planner AND kernel bodies) be
// compiled by g++ and executed on the host so that `-m "not gpu"` CI can check
// index math, packing layouts and error paths without a GPU.  Every CUDA thread
// of a block is an OS thread; __syncthreads() is a pthread barrier; blocks run
// one after another.  This is test infrastructure: the product package
// (jtransforms_b200) never builds, loads or falls back to it.
#pragma once
#include <pthread.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define JTB_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct double2 { double x, y; } __attribute__((aligned(16)));
struct float2 { float x, y; } __attribute__((aligned(8)));
struct float4 { float x, y, z, w; } __attribute__((aligned(16)));
static inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
template <class T> static inline T __ldg(const T* p) { return *p; }

namespace jtb_emu {
extern thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
extern pthread_barrier_t* g_bar;
extern unsigned char* g_smem;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
}  // namespace jtb_emu
#define threadIdx (jtb_emu::t_threadIdx)
#define blockIdx (jtb_emu::t_blockIdx)
#define blockDim (jtb_emu::t_blockDim)
#define gridDim (jtb_emu::t_gridDim)
static inline void __threadfence_system() {}
static inline void __syncthreads() { pthread_barrier_wait(jtb_emu::g_bar); }
#define JTB_DYN_SMEM(name) unsigned char* name = jtb_emu::g_smem
#define JTB_LAUNCH(kern, grid, block, smem, stream, ...) \
  jtb_emu::launch(dim3(grid), dim3(block), (smem), [=]() { kern(__VA_ARGS__); })

// ---- minimal runtime API -------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorUnknown = 999 };
typedef struct jtb_emu_stream* cudaStream_t;
typedef struct jtb_emu_event { double t; }* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaHostAllocDefault = 0, cudaStreamNonBlocking = 1,
       cudaEventDefault = 0, cudaEventDisableTiming = 2 };
static inline const char* cudaGetErrorString(cudaError_t e) { return e ? "emu error" : "no error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
enum { cudaHostRegisterDefault = 0 };
static inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = 0) {
  for (size_t r = 0; r < h; ++r) memmove((char*)d + r * dp, (const char*)s + r * sp, w);
  return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0);
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = (size_t)8 << 30; return cudaSuccess; }
