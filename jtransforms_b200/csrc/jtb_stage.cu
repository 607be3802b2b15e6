// Staged host <-> device copies for PAGEABLE caller memory (a Java heap double[] reached through a critical downcall,
// a plain numpy array).  cudaMemcpyAsync from pageable memory is staged by the driver on one thread (~15 GB/s
// measured); here T host threads copy 8 MiB chunks between the caller's array and page-locked bounce buffers while
// the DMA engine moves the previous chunks, so the copy approaches the PCIe rate without asking the caller to pin
// anything.  (The reference keeps everything in the Java heap -- fft/DoubleFFT_1D.java:243 -- so this is the path a
// drop-in caller hits.)
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "jtb_engine.h"

namespace jtb {

namespace {
constexpr size_t kChunk = (size_t)8 << 20;
constexpr int kMaxThreads = 8;

struct StagePool {
  int device = -1;
  int nthreads = 0;
  void* bounce[kMaxThreads] = {nullptr};
  cudaEvent_t ev[kMaxThreads] = {nullptr};
  cudaStream_t stream[kMaxThreads] = {nullptr};
};
StagePool g_pool;
std::mutex g_pool_mu;   // one staged copy at a time per process (the pool is rebuilt when the device changes)

int pool_init(int device) {
  if (g_pool.nthreads > 0 && g_pool.device == device) return ST_OK;
  if (g_pool.nthreads > 0) {            // another device: rebuild
    for (int t = 0; t < g_pool.nthreads; ++t) {
      cudaFreeHost(g_pool.bounce[t]);
      cudaEventDestroy(g_pool.ev[t]);
      cudaStreamDestroy(g_pool.stream[t]);
    }
    g_pool.nthreads = 0;
  }
  const char* et = getenv("JTB_STAGE_THREADS");
  int nt = et ? atoi(et) : (int)std::thread::hardware_concurrency() / 2;
  nt = std::max(1, std::min(nt, kMaxThreads));
  for (int t = 0; t < nt; ++t) {
    JTB_CUDA(cudaHostAlloc(&g_pool.bounce[t], kChunk, cudaHostAllocDefault));
    JTB_CUDA(cudaEventCreateWithFlags(&g_pool.ev[t], cudaEventDisableTiming));
    JTB_CUDA(cudaStreamCreateWithFlags(&g_pool.stream[t], cudaStreamNonBlocking));
  }
  g_pool.device = device;
  g_pool.nthreads = nt;
  return ST_OK;
}
}  // namespace

// true when `p` is ordinary pageable host memory (not cudaHostAlloc'ed / cudaHostRegister'ed)
bool host_is_pageable(const void* p) {
#ifdef JTB_EMU
  return getenv("JTB_EMU_PAGEABLE") != nullptr;
#else
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
#endif
}

// dev <- host (to_device) or host <- dev, `bytes` long; `after` (may be null) is an event the device side must wait for
// before the first chunk moves (device -> host: the kernels that produce the data).  Synchronous on return.
int staged_copy(int device, void* dev, void* host, size_t bytes, bool to_device, cudaEvent_t after) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  JTB_TRY(pool_init(device));
  const int nt = g_pool.nthreads;
  const size_t nchunks = (bytes + kChunk - 1) / kChunk;
  std::atomic<size_t> next(0);
  std::atomic<int> failed(0);
  auto worker = [&](int t) {
    if (cudaSetDevice(device) != cudaSuccess) { failed = 1; return; }
    cudaStream_t st = g_pool.stream[t];
    if (after && cudaStreamWaitEvent(st, after, 0) != cudaSuccess) { failed = 1; return; }
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= nchunks || failed.load()) break;
      const size_t off = i * kChunk, len = std::min(kChunk, bytes - off);
      if (to_device) {
        // the bounce buffer is free once its previous DMA has completed
        if (cudaEventSynchronize(g_pool.ev[t]) != cudaSuccess) { failed = 1; break; }
        memcpy(g_pool.bounce[t], (const char*)host + off, len);
        if (cudaMemcpyAsync((char*)dev + off, g_pool.bounce[t], len, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaEventRecord(g_pool.ev[t], st) != cudaSuccess) { failed = 1; break; }
      } else {
        if (cudaMemcpyAsync(g_pool.bounce[t], (const char*)dev + off, len, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaEventRecord(g_pool.ev[t], st) != cudaSuccess || cudaEventSynchronize(g_pool.ev[t]) != cudaSuccess) {
          failed = 1;
          break;
        }
        memcpy((char*)host + off, g_pool.bounce[t], len);
      }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) failed = 1;
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(worker, t);
  worker(0);
  for (auto& x : th) x.join();
  if (failed.load()) {
    cudaError_t e = cudaGetLastError();
    return cuda_fail(e == cudaSuccess ? cudaErrorUnknown : e, "staged host copy");
  }
  return ST_OK;
}

}  // namespace jtb
