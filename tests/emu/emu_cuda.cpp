// Host-thread implementation of the tests/emu CUDA shim (test infrastructure only).
#include "emu_cuda.h"

#include <chrono>

namespace jtb_emu {
thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
pthread_barrier_t* g_bar = nullptr;
unsigned char* g_smem = nullptr;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const unsigned nt = block.x * block.y * block.z;
  if (nt == 0 || grid.x == 0) return;
  unsigned char* sm = (unsigned char*)aligned_alloc(128, (smem + 255) / 128 * 128 + 128);
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, nt);
  g_bar = &bar;
  g_smem = sm;
  auto worker = [&](unsigned tx) {
    t_blockDim = block;
    t_gridDim = grid;
    t_threadIdx = dim3(tx % block.x, (tx / block.x) % block.y, tx / (block.x * block.y));
    for (unsigned bz = 0; bz < grid.z; ++bz)
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
          t_blockIdx = dim3(bx, by, bz);
          body();
          pthread_barrier_wait(&bar);  // block boundary: shared memory is reused
        }
  };
  std::vector<std::thread> th;
  th.reserve(nt);
  for (unsigned i = 0; i < nt; ++i) th.emplace_back(worker, i);
  for (auto& t : th) t.join();
  pthread_barrier_destroy(&bar);
  g_bar = nullptr;
  g_smem = nullptr;
  free(sm);
}
}  // namespace jtb_emu

static double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new jtb_emu_event{0}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = now_ms(); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
