// Host-side engine of libjtb200: per-device context (stream, workspaces, twiddle/chirp
// table caches) and the typed transform drivers that turn one JTransforms API call into
// a short sequence of sm_100a kernel launches.
#pragma once
#include <functional>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "jtb_elem.cuh"
#include "jtb_tile_host.h"

namespace jtb {

enum { ST_OK = 0, ST_ARG = 1, ST_UNSUPPORTED = 2, ST_CUDA = 3, ST_OOM = 4, ST_NCCL = 5 };

void set_error(const char* fmt, ...);
const char* last_error();
int cuda_fail(cudaError_t e, const char* what);   // records message, returns ST_CUDA / ST_OOM

#define JTB_CUDA(expr)                                                   \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) return ::jtb::cuda_fail(_e, #expr);           \
  } while (0)
#define JTB_TRY(expr)                    \
  do {                                   \
    int _s = (expr);                     \
    if (_s != ::jtb::ST_OK) return _s;   \
  } while (0)

static inline bool is_pow2(i64 n) { return n > 0 && (n & (n - 1)) == 0; }
static inline int ilog2(i64 n) { int l = 0; while ((1LL << l) < n) ++l; return l; }
static inline i64 next_pow2(i64 n) { return 1LL << ilog2(n); }

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

enum { WK_FOURSTEP = 0, WK_BLUE = 1, WK_REAL = 2, WK_FULL = 3, WK_BIG = 4, WK_COUNT = 5 };

struct TableEntry {
  void* p = nullptr;
  size_t bytes = 0;
  int refs = 0;       // plans that looked this table up (jtb_plan_destroy releases them)
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;       // library stream for the host-pointer API
  std::mutex mu;
  DevBuf work[WK_COUNT];
  DevBuf io;                           // device copy of the caller's host array
  cudaStream_t s_in = nullptr, s_out = nullptr;   // copy streams of the pipelined batch path (created on first use)
  cudaEvent_t ev_in[3] = {nullptr, nullptr, nullptr}, ev_c[3] = {nullptr, nullptr, nullptr}, ev_out[3] = {nullptr, nullptr, nullptr};
  int ensure_pipeline();
  std::map<std::string, TableEntry> tables;   // device tables keyed by name
  std::set<std::string>* recorder = nullptr;  // keys touched by the plan whose call is running (under `mu`)
  unsigned long table_gen = 0;                // bumped whenever a table is freed (invalidates pointer memos)
  size_t work_cap = (size_t)8 << 30;   // chunk limit per workspace
  long launches = 0;                   // kernels launched (bench.py reports this)
  bool tile_init_done[2] = {false, false};
  // The workspaces above are shared by every caller stream of this device.  Calls are enqueued under `mu`; when a
  // call arrives on another stream than the previous one it first waits for the event recorded at the end of that
  // previous call, so two streams never run library kernels on the same workspace concurrently.
  cudaEvent_t ev_order = nullptr;
  cudaStream_t last_stream = nullptr;
  bool has_last = false;
  int order_begin(cudaStream_t st);
  int order_end(cudaStream_t st);
  // watchdog word of the spin-waiting kernels (team barrier of fft_slice2d_kernel, peer_barrier_kernel): mapped
  // pinned host memory, written by the device on a time-out, read and cleared by the host at the next sync point
  int* wd_dev = nullptr;
  volatile int* wd_host = nullptr;
  int ensure_watchdog();
  int check_watchdog(const char* where);   // ST_CUDA + message when a kernel gave up waiting
  int ensure(DevBuf& b, size_t bytes);
  void* table(const std::string& key);
  int put_table(const std::string& key, const void* host, size_t bytes, void** dev_out);
  int adopt_table(const std::string& key, void* dev, size_t bytes = 0);
  void release_tables(const std::set<std::string>& keys);   // jtb_plan_destroy
};

// RAII: every entry point makes its device current and restores the caller's on exit
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
    if (prev != device) ok = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

extern int g_limit_contig, g_limit_strided;   // test knobs: force the two-pass path at small sizes (0 = off)
Ctx* get_ctx(int device);   // creates on first use; nullptr + error on failure
bool host_is_pageable(const void* p);                                                                   // jtb_stage.cu
int staged_copy(int device, void* dev, void* host, size_t bytes, bool to_device, cudaEvent_t after);   // jtb_stage.cu
int staged_copy_2d(int device, void* dev, void* host, size_t hpitch, size_t width, size_t rows, bool to_device,
                   cudaEvent_t after, int max_threads);                                                   // jtb_stage.cu
void grid_for(i64 work_items, unsigned* grid, unsigned* block);

// Fusion options of one power-of-two c2c call (see TileParams)
template <typename T> struct Fuse {
  int swap_in = 0, swap_in2 = 0, swap_out1 = 0, swap_out = 0;
  const cx<T>* premul = nullptr;  int premul_conj = 0;
  const cx<T>* postmul = nullptr; int postmul_conj = 0;
  i64 valid_in = -1, valid_out = -1;
  int has_scale = 0; T scale = 1;
};

template <typename T> struct Engine {
  typedef cx<T> C;
  Ctx* ctx;
  cudaStream_t st;
  // > 0: strided lean passes launch at most this many (persistent) CTAs, each walking over several tiles -- lets a pass
  // share the SMs with a concurrent kernel on another stream instead of flooding every slot (pipelined slab exchange)
  int cta_limit = 0;
  Engine(Ctx* c, cudaStream_t s) : ctx(c), st(s) {}
  static const char* pname();
  static int max_logn_contig();    // longest line one CTA transforms (contiguous lines)
  static int max_logn_strided();   // ... when several adjacent strided lines must share the CTA

  int init_tiles();
  // tables ------------------------------------------------------------------
  int tile_tables(int logn, const C** tw /*[JTB_MAX_STAGES]*/, const C** rtw);
  int fs_tables(int logN, const C** A, const C** B, int* logL);
  int blue_tables(i64 n, const C** bk1, const C** bk2, i64* M);
  int dct_table(i64 n, const C** dtw);

  // power-of-two c2c on lines [l0, l1) (tile kernel or four-step).  `in` and `out` may alias when
  // gi == go (in place).
  int c2c_pow2(const C* in, const Geo& gi, C* out, const Geo& go, i64 l0, i64 l1, int logn, const Fuse<T>& f,
               int pro = PRO_DIRECT, int epi = EPI_DIRECT);
  int tile_call(const C* in, const Geo& gi, C* out, const Geo& go, i64 l0, i64 l1, int logn, TileParams<T>& p);
  // contiguous in-place lines longer than the two-pass limit: sub-transforms that are themselves two-pass
  int c2c_big_contig(C* a, i64 dist, i64 l0, i64 l1, int logn, bool inverse, bool has_scale, T scale);
  // any-length in-place c2c on a batch of lines
  int c2c_lines(C* a, const Geo& g, i64 nlines, i64 n, bool inverse, bool has_scale, T scale);
  // real <-> packed half spectrum (1-D rule of fft/DoubleFFT_1D.java), lines of n reals, geometry in REAL units
  int real_forward_lines(T* a, const Geo& g, i64 nlines, i64 n);
  int real_inverse_lines(T* a, const Geo& g, i64 nlines, i64 n, bool scale);
  // DCT/DST/DHT along lines (geometry in real units). kind: 1 DCT, 2 DST, 3 DHT
  int r2r_lines(T* a, const Geo& g, i64 nlines, i64 n, int kind, bool inverse, bool scale);
};

template <typename T>
int fast_scatter(Engine<T>& e, const cx<T>* a, i64 Ls, i64 R, i64 Cn, int nranks, int rank, void* const* peers,
                 bool inverse, i64 slice_base = -1, bool back = false, i64 col0 = 0, i64 ncols = -1);   // jtb_fast.cu
template <typename T> int fast_scatter_width(i64 R, i64 Cn);   // jtb_fast.cu
template <typename T> bool fast_pipe_has(i64 R, i64 S, i64 Cn, int nranks, int nb);   // jtb_fast.cu
template <typename T>
int fast_pipe_exchange(Engine<T>& e, const cx<T>* a, i64 Ls, i64 R, i64 Cn, int nranks, int rank, void* const* peers,
                       void* const* flag_ptrs, long long epoch, int* counters, int nb, bool inverse, bool has_scale, T scale,
                       bool* handled);   // jtb_fast.cu
template <typename T>
int fast_scatter_tma(Engine<T>& e, const cx<T>* a, i64 Ls, i64 R, i64 Cn, i64 S, int nranks, int rank, void* const* peers,
                     bool inverse, i64 col0, i64 ncols, bool* handled);   // jtb_tma.cu
template <typename T>
int fast_c2c_out(Engine<T>& e, cx<T>* a, const Geo& g, cx<T>* out, i64 out_dist, i64 out_stride, i64 nlines, int logn,
                 bool inverse, bool has_scale, T scale, bool* handled);   // jtb_fast.cu
template <typename T> bool fast_has_strided(int logn, i64 c0);   // jtb_fast.cu
template <typename T>
int fast_tma_c2c(Engine<T>& e, cx<T>* a, const Geo& g, i64 nlines, int logn, bool inverse, bool has_scale, T scale,
                 bool* handled);   // jtb_tma.cu
template <typename T> int fast_stage_table(Engine<T>& e, int logn, int loge, const cx<T>** out);   // jtb_fast.cu
template <typename T>
int fast_fourstep_contig(Engine<T>& e, const cx<T>* in, i64 in_dist, cx<T>* out, i64 out_dist, i64 l0, i64 l1, int logn,
                         bool swap_in, bool swap_out, bool has_scale, T scale, bool* handled);   // jtb_fast2.cu
template <typename T>
int fast_fourstep_strided(Engine<T>& e, cx<T>* a, const Geo& g, i64 nlines, int logn, bool inverse, bool has_scale,
                          T scale, bool* handled);
template <typename T>
int fast_threepass_contig(Engine<T>& e, cx<T>* a, i64 dist, i64 l0, i64 l1, int logn, bool inverse, bool has_scale,
                          T scale, bool* handled);   // jtb_fast2.cu
template <typename T> int fast_rfft_fwd(Engine<T>& e, cx<T>* a, i64 dist, i64 nlines, int logN, bool* handled);
template <typename T>
int fast_r2r_rows(Engine<T>& e, T* a, i64 dist, i64 nlines, i64 n, int kind, T f0, T f, bool* handled);
template <typename T> int fast_dht2d_rows(Engine<T>& e, T* a, i64 R, i64 n, T f, bool* handled);   // jtb_fast2.cu
template <typename T>
int fast_r2r_cols(Engine<T>& e, T* a, i64 n, i64 Cn, i64 batches, i64 bdist, int kind, T f0, T f, bool* handled);
template <typename T>
int fast_r2r_rows_inv(Engine<T>& e, T* a, i64 dist, i64 nlines, i64 n, int kind, T f0, T f, bool* handled);   // jtb_r2r_inv.cu
template <typename T>
int fast_rfft_inv(Engine<T>& e, cx<T>* a, i64 dist, i64 nlines, int logN, bool has_scale, T scale, bool* handled);   // jtb_r2r_inv.cu
template <typename T>
int fast_r2r_cols_single(Engine<T>& e, T* a, i64 n, i64 Cn, i64 batches, i64 bdist, int kind, bool inverse, T f0, T f,
                         bool* handled);
template <typename T>
int fast_r2r_cols_inv(Engine<T>& e, T* a, i64 n, i64 Cn, i64 batches, i64 bdist, int kind, T f0, T f, bool* handled);
template <typename T>
int fast_bluestein_contig(Engine<T>& e, cx<T>* a, i64 dist, i64 nlines, i64 n, bool inverse, bool has_scale, T scale,
                          bool* handled);
template <typename T>
int fast_slice2d(Engine<T>& e, cx<T>* a, i64 nslices, i64 N, bool inverse, bool has_scale, T scale, int nranks, int rank,
                 void* const* peers, bool* handled);   // jtb_fast.cu
template <typename T>
int mixed_c2c(Engine<T>& e, cx<T>* a, const Geo& g, i64 nlines, i64 n, bool inverse, bool has_scale, T scale,
              bool* handled);   // jtb_mixed.cu
template <typename T>
int mixed_twopass_contig(Engine<T>& e, cx<T>* a, i64 dist, i64 nlines, i64 n, bool inverse, bool has_scale, T scale,
                         bool* handled);   // jtb_mixed.cu
// what: 1 publish `epoch` to every peer, 2 wait until every peer has published it, 3 both (barrier)
int peer_barrier(Ctx* ctx, cudaStream_t st, void* const* flag_ptrs, int nranks, int rank, long long epoch, int what = 3);

extern template struct Engine<double>;
extern template struct Engine<float>;

}  // namespace jtb
