// Per-device context of libjtb200: library stream, growable workspaces, device table cache,
// thread-local error text.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "jtb_engine.h"

namespace jtb {

static thread_local char g_err[512] = "";
int g_limit_contig = 0, g_limit_strided = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error '%s' in %s", cudaGetErrorString(e), what);
  return e == cudaErrorMemoryAllocation ? ST_OOM : ST_CUDA;
}

void grid_for(i64 work_items, unsigned* grid, unsigned* block) {
  const unsigned b = 256;
  i64 g = (work_items + b - 1) / b;
  const i64 cap = 148LL * 8 * 4;     // grid-stride loops: a few waves of 148 SMs x 8 CTAs
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  *grid = (unsigned)g;
  *block = b;
}

int Ctx::ensure(DevBuf& b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return ST_OK;
  if (b.p) {
    JTB_CUDA(cudaDeviceSynchronize());   // kernels in flight may still use the old buffer
    JTB_CUDA(cudaFree(b.p));
    b.p = nullptr; b.bytes = 0;
  }
  size_t want = bytes < 256 ? 256 : bytes;
  JTB_CUDA(cudaMalloc(&b.p, want));
  b.bytes = want;
  return ST_OK;
}

int Ctx::ensure_pipeline() {
  if (s_in) return ST_OK;
  JTB_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
  JTB_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
  for (int i = 0; i < 3; ++i) {
    JTB_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
    JTB_CUDA(cudaEventCreateWithFlags(&ev_c[i], cudaEventDisableTiming));
    JTB_CUDA(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
  }
  return ST_OK;
}

void* Ctx::table(const std::string& key) {
  auto it = tables.find(key);
  if (it == tables.end()) return nullptr;
  if (recorder && recorder->insert(key).second) it->second.refs++;
  return it->second.p;
}

int Ctx::put_table(const std::string& key, const void* host, size_t bytes, void** dev_out) {
  void* d = nullptr;
  JTB_CUDA(cudaMalloc(&d, bytes < 16 ? 16 : bytes));
  // synchronous copy: the host vector dies when the caller returns
  JTB_CUDA(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
  // a pageable H2D cudaMemcpy may return before the DMA has landed; kernels on our non-blocking streams would
  // otherwise be able to read a half-written table
  JTB_CUDA(cudaDeviceSynchronize());
  *dev_out = d;
  return adopt_table(key, d, bytes);
}

int Ctx::adopt_table(const std::string& key, void* dev, size_t bytes) {
  TableEntry& t = tables[key];
  t.p = dev; t.bytes = bytes;
  if (recorder && recorder->insert(key).second) t.refs++;
  return ST_OK;
}

// Tables a destroyed plan was the last user of are freed (a caller sweeping over sizes -- the reference's benchmark
// loop -- would otherwise grow the cache monotonically; a Bluestein plan at n = 10^6 holds ~50 MB).  Small tables
// (stage twiddles, a few KiB, keyed by log2 n: a bounded set) stay cached.
void Ctx::release_tables(const std::set<std::string>& keys) {
  bool synced = false;
  for (const auto& k : keys) {
    auto it = tables.find(k);
    if (it == tables.end()) continue;
    if (--it->second.refs > 0 || it->second.bytes < ((size_t)64 << 10)) continue;
    if (!synced) { cudaSetDevice(device); cudaDeviceSynchronize(); synced = true; }   // kernels in flight may read it
    cudaFree(it->second.p);
    tables.erase(it);
    ++table_gen;
  }
}

// A stream that is being captured into a CUDA graph takes no part in the cross-stream ordering: an event recorded
// inside a capture belongs to the graph and can never be waited on from outside it (the caller replays the graph on
// one stream and owns the ordering of its replays).
static bool stream_capturing(cudaStream_t st) {
#ifdef JTB_EMU
  (void)st;
  return false;
#else
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
  return cs != cudaStreamCaptureStatusNone;
#endif
}

int Ctx::order_begin(cudaStream_t st) {
  if (stream_capturing(st)) return ST_OK;
  if (has_last && st != last_stream) {
    if (!ev_order) JTB_CUDA(cudaEventCreateWithFlags(&ev_order, cudaEventDisableTiming));
    JTB_CUDA(cudaStreamWaitEvent(st, ev_order, 0));
  }
  return ST_OK;
}
int Ctx::order_end(cudaStream_t st) {
  if (stream_capturing(st)) return ST_OK;
  if (!ev_order) JTB_CUDA(cudaEventCreateWithFlags(&ev_order, cudaEventDisableTiming));
  JTB_CUDA(cudaEventRecord(ev_order, st));
  last_stream = st; has_last = true;
  return ST_OK;
}

int Ctx::ensure_watchdog() {
  if (wd_dev) return ST_OK;
#ifdef JTB_EMU
  wd_host = (volatile int*)calloc(1, sizeof(int));
  wd_dev = (int*)wd_host;
#else
  void* h = nullptr;
  JTB_CUDA(cudaHostAlloc(&h, sizeof(int), cudaHostAllocMapped));
  *(volatile int*)h = 0;
  void* d = nullptr;
  JTB_CUDA(cudaHostGetDevicePointer(&d, h, 0));
  wd_host = (volatile int*)h; wd_dev = (int*)d;
#endif
  return ST_OK;
}
int Ctx::check_watchdog(const char* where) {
  if (!wd_host || *wd_host == 0) return ST_OK;
  const int code = *wd_host;
  *wd_host = 0;
  set_error("%s: a device-side wait timed out on device %d (code %d: %s); the result is invalid", where, device, code,
            code == 1 ? "slice team barrier" : "peer barrier -- a peer GPU did not publish its epoch");
  return ST_CUDA;
}

static std::mutex g_ctx_mu;
static std::map<int, Ctx*> g_ctx;

Ctx* get_ctx(int device) {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  auto it = g_ctx.find(device);
  if (it != g_ctx.end()) return it->second;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count < 1) {
    set_error("no CUDA device available (%s): libjtb200 has no CPU fallback", cudaGetErrorString(e));
    return nullptr;
  }
  if (device < 0 || device >= count) { set_error("device %d out of range (have %d)", device, count); return nullptr; }
  DeviceGuard dg(device);
  if (!dg.ok) { cuda_fail(cudaGetLastError(), "cudaSetDevice"); return nullptr; }
  Ctx* c = new Ctx();
  c->device = device;
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    cuda_fail(e, "cudaStreamCreate");
    delete c;
    return nullptr;
  }
  g_ctx[device] = c;
  return c;
}

__global__ void k_cast_c64_c32(const double2* in, float2* out, i64 count) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (i64)gridDim.x * blockDim.x) {
    const double2 z = in[i];
    float2 r; r.x = (float)z.x; r.y = (float)z.y;
    out[i] = r;
  }
}

}  // namespace jtb
