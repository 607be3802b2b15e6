// Staged host <-> device copies for PAGEABLE caller memory (a Java heap double[] reached through a critical downcall,
// a plain numpy array).  cudaMemcpyAsync from pageable memory is staged by the driver on one thread (~15 GB/s
// measured); here T host threads copy 8 MiB chunks between the caller's array and page-locked bounce buffers while
// the DMA engine moves the previous chunks, so the copy approaches the PCIe rate without asking the caller to pin
// anything.  (The reference keeps everything in the Java heap -- fft/DoubleFFT_1D.java:243 -- so this is the path a
// drop-in caller hits.)  One pool per device (its own bounce buffers, streams and mutex), so the members of a
// multi-GPU plan stage their slabs concurrently over their own PCIe links.  The copy may be two-dimensional: `rows`
// rows of `width` bytes, dense on the device, `hpitch` bytes apart on the host -- the natural-order delivery of a
// k2-slabbed result into the caller's [S][R][C] array.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "jtb_engine.h"

namespace jtb {

namespace {
constexpr size_t kChunk = (size_t)8 << 20;
constexpr int kMaxThreads = 8;

struct StagePool {
  std::mutex mu;        // one staged copy at a time per device
  int nthreads = 0;
  void* bounce[kMaxThreads] = {nullptr};
  cudaEvent_t ev[kMaxThreads] = {nullptr};
  cudaStream_t stream[kMaxThreads] = {nullptr};
};
std::mutex g_pools_mu;
std::map<int, StagePool*> g_pools;

StagePool* pool_for(int device) {
  std::lock_guard<std::mutex> lk(g_pools_mu);
  auto it = g_pools.find(device);
  if (it != g_pools.end()) return it->second;
  StagePool* p = new StagePool();
  g_pools[device] = p;
  return p;
}

int pool_grow(StagePool* p, int want) {   // caller holds p->mu and has made the device current
  for (int t = p->nthreads; t < want; ++t) {
    JTB_CUDA(cudaHostAlloc(&p->bounce[t], kChunk, cudaHostAllocDefault));
    JTB_CUDA(cudaEventCreateWithFlags(&p->ev[t], cudaEventDisableTiming));
    JTB_CUDA(cudaStreamCreateWithFlags(&p->stream[t], cudaStreamNonBlocking));
    p->nthreads = t + 1;
  }
  return ST_OK;
}

// piecewise copy between a dense staging chunk and the (possibly pitched) host array
void host_piece(char* host, size_t hpitch, size_t width, size_t off, size_t len, char* dense, bool to_dense) {
  while (len > 0) {
    const size_t row = off / width, col = off - row * width;
    const size_t n = std::min(len, width - col);
    char* h = host + row * hpitch + col;
    if (to_dense) memcpy(dense, h, n); else memcpy(h, dense, n);
    dense += n; off += n; len -= n;
  }
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

int stage_threads_default() {
  const char* et = getenv("JTB_STAGE_THREADS");
  int nt = et ? atoi(et) : (int)std::thread::hardware_concurrency() / 2;
  return std::max(1, std::min(nt, kMaxThreads));
}

// dev (dense, rows*width bytes) <-> host (rows of `width` bytes, `hpitch` apart); `after` (may be null) is an event the
// device side must wait for before the first chunk moves (device -> host: the kernels that produce the data).
// Synchronous on return.  max_threads <= 0: JTB_STAGE_THREADS / half the cores.
int staged_copy_2d(int device, void* dev, void* host, size_t hpitch, size_t width, size_t rows, bool to_device,
                   cudaEvent_t after, int max_threads) {
  if (width == 0 || rows == 0) return ST_OK;
  StagePool* pool = pool_for(device);
  std::lock_guard<std::mutex> lk(pool->mu);
  JTB_CUDA(cudaSetDevice(device));
  int nt = max_threads > 0 ? std::min(max_threads, kMaxThreads) : stage_threads_default();
  const size_t bytes = width * rows;
  const size_t nchunks = (bytes + kChunk - 1) / kChunk;
  if ((size_t)nt > nchunks) nt = (int)nchunks;
  JTB_TRY(pool_grow(pool, nt));
  std::atomic<size_t> next(0);
  std::atomic<int> failed(0);
  auto worker = [&](int t) {
    if (cudaSetDevice(device) != cudaSuccess) { failed = 1; return; }
    cudaStream_t st = pool->stream[t];
    if (after && cudaStreamWaitEvent(st, after, 0) != cudaSuccess) { failed = 1; return; }
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= nchunks || failed.load()) break;
      const size_t off = i * kChunk, len = std::min(kChunk, bytes - off);
      if (to_device) {
        // the bounce buffer is free once its previous DMA has completed
        if (cudaEventSynchronize(pool->ev[t]) != cudaSuccess) { failed = 1; break; }
        host_piece((char*)host, hpitch, width, off, len, (char*)pool->bounce[t], true);
        if (cudaMemcpyAsync((char*)dev + off, pool->bounce[t], len, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaEventRecord(pool->ev[t], st) != cudaSuccess) { failed = 1; break; }
      } else {
        if (cudaMemcpyAsync(pool->bounce[t], (const char*)dev + off, len, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaEventRecord(pool->ev[t], st) != cudaSuccess || cudaEventSynchronize(pool->ev[t]) != cudaSuccess) {
          failed = 1;
          break;
        }
        host_piece((char*)host, hpitch, width, off, len, (char*)pool->bounce[t], false);
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

int staged_copy(int device, void* dev, void* host, size_t bytes, bool to_device, cudaEvent_t after) {
  return staged_copy_2d(device, dev, host, bytes, bytes, 1, to_device, after, 0);
}

}  // namespace jtb
