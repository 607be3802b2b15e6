// C ABI of libjtb200 (include/jtb200.h): plans, the N-D drivers that sequence the line kernels, and
// the host-pointer / device-pointer execution entry points.
//
// N-D drivers replace the reference's row/column/slice loops over the ConcurrencyUtils thread pool:
//   2-D complex  fft/DoubleFFT_2D.java:115-213 (+ cdft2d_subth :3352-3529)
//   2-D real     fft/DoubleFFT_2D.java:820-838 (xdft2d0_subth1 :3100, cdft2d_subth :3352, rdft2d_sub :2544)
//   3-D complex  fft/DoubleFFT_3D.java:145-325 (xdft3da_subth2 :5505, cdft3db_subth :6318)
//   3-D real     fft/DoubleFFT_3D.java:1339-1355 (rdft3d_sub :6909-7021)
//   DCT/DST/DHT  dct/DoubleDCT_2D.java:104-183, dht/DoubleDHT_2D.java:102-190 (+ yTransform :1288-1309)
#include <cstdlib>
#include <cstring>

#include <algorithm>
#include <thread>

#include "../../include/jtb200.h"
#include "jtb_engine.h"
#include "jtb_slab.h"

using namespace jtb;

struct jtb_plan {
  int kind, prec, rank, device;
  i64 dims[3];
  i64 total;
  Ctx* ctx;
  std::set<std::string> keys;          // device tables this plan looked up (released by jtb_plan_destroy)
  // multi-GPU (jtb_plan_set_devices)
  std::vector<int> devices;            // empty: single device
  std::vector<jtb_plan*> sub;          // one single-device plan per listed device (batch sharding)
  std::vector<jtb_slab*> slabs;        // slab members (rank-3 FFT plans whose slices and rows divide by P)
  std::vector<void*> slab_in;          // per member: device copy of its [S/P][R][C] slab
  std::vector<cudaStream_t> mstream;   // per member stream
  std::vector<cudaEvent_t> mev;
};

namespace {
// Scope of one plan call on one stream: makes the plan's device current (restored on exit), takes the device
// context's mutex, records which tables the plan uses, and orders the call after the previous one when that ran on
// another stream (the context's workspaces are shared by all streams).
struct PlanCall {
  Ctx* c;
  DeviceGuard dg;
  std::unique_lock<std::mutex> lk;
  cudaStream_t st;
  bool began = false;
  PlanCall(jtb_plan* p, cudaStream_t stream) : c(p->ctx), dg(p->device), lk(p->ctx->mu), st(stream) { c->recorder = &p->keys; }
  int begin() {
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    const int s = c->order_begin(st);
    began = s == ST_OK;
    return s;
  }
  ~PlanCall() {
    if (began) c->order_end(st);
    c->recorder = nullptr;
  }
};
// same for the plan-less device entry points
struct CtxCall {
  Ctx* c;
  DeviceGuard dg;
  std::unique_lock<std::mutex> lk;
  cudaStream_t st;
  bool began = false;
  CtxCall(Ctx* ctx, cudaStream_t stream) : c(ctx), dg(ctx->device), lk(ctx->mu), st(stream) {}
  int begin() {
    if (!dg.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    const int s = c->order_begin(st);
    began = s == ST_OK;
    return s;
  }
  ~CtxCall() { if (began) c->order_end(st); }
};
}  // namespace

namespace {

template <typename T> int nd_c2c(Engine<T>& e, cx<T>* a, int rank, const i64* d, bool inverse, bool scale) {
  typedef cx<T> C;
  (void)sizeof(C);
  if (rank == 1) return e.c2c_lines(a, geo_contig(d[0]), 1, d[0], inverse, scale, (T)(1.0 / (double)d[0]));
  if (rank == 2) {
    const i64 R = d[0], Cn = d[1];
    JTB_TRY(e.c2c_lines(a, geo_contig(Cn), R, Cn, inverse, false, (T)1));
    return e.c2c_lines(a, geo_make(Cn, 1, R * Cn, Cn), Cn, R, inverse, scale, (T)(1.0 / ((double)R * (double)Cn)));
  }
  const i64 S = d[0], R = d[1], Cn = d[2];
  bool fused = false;
  if (R == Cn) JTB_TRY(fast_slice2d<T>(e, a, S, R, inverse, false, (T)1, 1, 0, nullptr, &fused));   // rows + columns per slice, via L2
  // Axis-swapping variant (default for double; JTB_XPOSE=0/1 overrides): the k2 pass stores its rows into a work array
  // laid out [r][s][c], so that the k1 pass READS lines whose elements are one row (not one slice) apart and stores
  // them back in natural order; both strided passes then read with the small stride (L2 prefetch pays, one page per
  // 256 rows instead of one per row) and only their fire-and-forget stores use the slice stride.  512^3 double on
  // B200: 2.159 -> 1.973 ms.  Costs a second array; skipped when that does not fit (JTB_XPOSE_MAX_MB, 16 GiB).
  static const int xpose = getenv("JTB_XPOSE") ? atoi(getenv("JTB_XPOSE")) : (sizeof(T) == 8 ? 1 : 0);
  static const double xpose_max_mb = getenv("JTB_XPOSE_MAX_MB") ? atof(getenv("JTB_XPOSE_MAX_MB")) : 16384.0;
  const size_t xbytes = (size_t)(S * R * Cn) * sizeof(C);
  bool xp = !fused && xpose && is_pow2(R) && is_pow2(S) && fast_has_strided<T>(ilog2(R), Cn) &&
            fast_has_strided<T>(ilog2(S), Cn) && S * R * Cn < 0x7fffffffLL && (double)xbytes <= xpose_max_mb * 1048576.0;
  if (xp && e.ctx->work[WK_FOURSTEP].bytes < xbytes) {
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess || fr < xbytes + ((size_t)256 << 20)) { xp = false; cudaGetLastError(); }
  }
  if (xp) {
    JTB_TRY(e.ctx->ensure(e.ctx->work[WK_FOURSTEP], xbytes));
    C* wk = (C*)e.ctx->work[WK_FOURSTEP].p;
    JTB_TRY(e.c2c_lines(a, geo_contig(Cn), S * R, Cn, inverse, false, (T)1));
    bool h2 = false, h1 = false;
    JTB_TRY(fast_c2c_out<T>(e, a, geo_make(Cn, 1, R * Cn, Cn), wk, Cn, S * Cn, Cn * S, ilog2(R), inverse, false, (T)1, &h2));
    if (h2) {
      JTB_TRY(fast_c2c_out<T>(e, wk, geo_make(Cn, 1, S * Cn, Cn), a, Cn, R * Cn, Cn * R, ilog2(S), inverse, scale,
                              (T)(1.0 / ((double)S * (double)R * (double)Cn)), &h1));
      if (!h1) { set_error("internal: axis-swapping k1 pass unavailable"); return ST_UNSUPPORTED; }
      return ST_OK;
    }
    JTB_TRY(e.c2c_lines(a, geo_make(Cn, 1, R * Cn, Cn), Cn * S, R, inverse, false, (T)1));
    fused = true;   // rows and columns are done; fall through to the in-place k1 pass
  }
  if (!fused) {
    JTB_TRY(e.c2c_lines(a, geo_contig(Cn), S * R, Cn, inverse, false, (T)1));
    JTB_TRY(e.c2c_lines(a, geo_make(Cn, 1, R * Cn, Cn), Cn * S, R, inverse, false, (T)1));
  }
  return e.c2c_lines(a, geo_make(R * Cn, 1, S * R * Cn, R * Cn), R * Cn, S, inverse, scale,
                     (T)(1.0 / ((double)S * (double)R * (double)Cn)));
}

template <typename T> int untangle(Engine<T>& e, T* a, int rank, const i64* d, int dir) {
  unsigned g, b;
  if (rank == 2) {
    if (d[0] / 2 - 1 < 1) return ST_OK;
    grid_for(d[0] / 2 - 1, &g, &b);
    JTB_LAUNCH(k_untangle2d<T>, g, b, 0, e.st, a, d[0], d[1], dir);
  } else {
    grid_for(d[0] * (d[1] / 2 + 1), &g, &b);
    JTB_LAUNCH(k_untangle3d<T>, g, b, 0, e.st, a, d[0], d[1], d[2], dir);
  }
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  return ST_OK;
}

template <typename T> int real_packed(Engine<T>& e, T* a, int rank, const i64* d, bool inverse, bool scale) {
  typedef cx<T> C;
  if (rank == 1) {
    return inverse ? e.real_inverse_lines(a, geo_contig(d[0]), 1, d[0], scale)
                   : e.real_forward_lines(a, geo_contig(d[0]), 1, d[0]);
  }
  for (int k = 0; k < rank; ++k)
    if (!is_pow2(d[k])) {
      set_error(rank == 2 ? "rows and columns must be power of two numbers"
                          : "slices, rows and columns must be power of two numbers");
      return ST_ARG;
    }
  C* ac = (C*)a;
  if (rank == 2) {
    const i64 R = d[0], Cn = d[1], H = Cn / 2;
    if (!inverse) {
      JTB_TRY(e.real_forward_lines(a, geo_contig(Cn), R, Cn));
      JTB_TRY(e.c2c_lines(ac, geo_make(H, 1, R * H, H), H, R, false, false, (T)1));
      return untangle(e, a, 2, d, +1);
    }
    JTB_TRY(untangle(e, a, 2, d, -1));
    JTB_TRY(e.c2c_lines(ac, geo_make(H, 1, R * H, H), H, R, true, scale, (T)(1.0 / (double)R)));
    return e.real_inverse_lines(a, geo_contig(Cn), R, Cn, scale);
  }
  const i64 S = d[0], R = d[1], Cn = d[2], H = Cn / 2;
  if (!inverse) {
    JTB_TRY(e.real_forward_lines(a, geo_contig(Cn), S * R, Cn));
    JTB_TRY(e.c2c_lines(ac, geo_make(H, 1, R * H, H), H * S, R, false, false, (T)1));
    JTB_TRY(e.c2c_lines(ac, geo_make(R * H, 1, S * R * H, R * H), R * H, S, false, false, (T)1));
    return untangle(e, a, 3, d, +1);
  }
  JTB_TRY(untangle(e, a, 3, d, -1));
  JTB_TRY(e.c2c_lines(ac, geo_make(R * H, 1, S * R * H, R * H), R * H, S, true, scale, (T)(1.0 / (double)S)));
  JTB_TRY(e.c2c_lines(ac, geo_make(H, 1, R * H, H), H * S, R, true, scale, (T)(1.0 / (double)R)));
  return e.real_inverse_lines(a, geo_contig(Cn), S * R, Cn, scale);
}

// realForwardFull / realInverseFull: total reals in the first half of a 2*total array -> full complex result
template <typename T> int real_full(Engine<T>& e, T* a, int rank, const i64* d, i64 total, bool inverse, bool scale) {
  typedef cx<T> C;
  JTB_TRY(e.ctx->ensure(e.ctx->work[WK_FULL], (size_t)total * sizeof(C)));
  C* wk = (C*)e.ctx->work[WK_FULL].p;
  unsigned g, b;
  {
    // power-of-two sizes: packed real transform at half the work (realForward on a copy), then one expansion sweep
    // that fills the Hermitian half (the reference: realForward + fillSymmetric, fft/DoubleFFT_2D.java:956-975)
    bool p2 = total >= 4;
    for (int k = 0; k < rank; ++k) p2 = p2 && is_pow2(d[k]) && d[k] >= 2;
    static const bool off = getenv("JTB_NO_FASTFULL") != nullptr;
    if (p2 && !off) {
      T* pk = (T*)wk;
      JTB_CUDA(cudaMemcpyAsync(pk, a, (size_t)total * sizeof(T), cudaMemcpyDeviceToDevice, e.st));
      JTB_TRY(real_packed(e, pk, rank, d, false, false));
      const T f = (inverse && scale) ? (T)(1.0 / (double)total) : (T)1;
      grid_for(total, &g, &b);
      if (rank == 1) JTB_LAUNCH(k_expand_full_1d<T>, g, b, 0, e.st, pk, (C*)a, d[0], inverse ? 1 : 0, f);
      else if (rank == 2) JTB_LAUNCH(k_expand_full_2d<T>, g, b, 0, e.st, pk, (C*)a, d[0], d[1], inverse ? 1 : 0, f);
      else JTB_LAUNCH(k_expand_full_3d<T>, g, b, 0, e.st, pk, (C*)a, d[0], d[1], d[2], inverse ? 1 : 0, f);
      JTB_CUDA(cudaGetLastError());
      e.ctx->launches++;
      return ST_OK;
    }
  }
  R2RParams<T> p;
  p.a = a; p.work = wk; p.g = geo_contig(total); p.line_base = 0; p.nlines = 1; p.n = total;
  p.mode = PRE_R2C; p.dst = 0; p.f0 = p.f = (T)1; p.dtw = nullptr;
  grid_for(total, &g, &b);
  JTB_LAUNCH(k_r2r_pre<T>, g, b, 0, e.st, p);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  JTB_TRY(nd_c2c(e, wk, rank, d, inverse, scale));
  JTB_CUDA(cudaMemcpyAsync(a, wk, (size_t)total * sizeof(C), cudaMemcpyDeviceToDevice, e.st));
  return ST_OK;
}

template <typename T> int nd_r2r(Engine<T>& e, T* a, int kind, int rank, const i64* d, bool inverse, bool scale) {
  if (rank == 1) return e.r2r_lines(a, geo_contig(d[0]), 1, d[0], kind, inverse, scale);
  unsigned g, b;
  if (rank == 2) {
    const i64 R = d[0], Cn = d[1];
    JTB_TRY(e.r2r_lines(a, geo_make(Cn, 1, R * Cn, Cn), Cn, R, kind, inverse, scale));
    if (kind == JTB_DHT) {   // rows + yTransform in one pass (row pairs r, R-r per CTA)
      bool folded = false;
      JTB_TRY(fast_dht2d_rows<T>(e, a, R, Cn, (inverse && scale) ? (T)(1.0 / (double)Cn) : (T)1, &folded));
      if (folded) return ST_OK;
    }
    JTB_TRY(e.r2r_lines(a, geo_contig(Cn), R, Cn, kind, inverse, scale));
    if (kind == JTB_DHT) {
      grid_for((R / 2 + 1) * (Cn / 2 + 1), &g, &b);
      JTB_LAUNCH(k_ytransform2d<T>, g, b, 0, e.st, a, R, Cn);
      JTB_CUDA(cudaGetLastError());
      e.ctx->launches++;
    }
    return ST_OK;
  }
  const i64 S = d[0], R = d[1], Cn = d[2];
  JTB_TRY(e.r2r_lines(a, geo_contig(Cn), S * R, Cn, kind, inverse, scale));
  JTB_TRY(e.r2r_lines(a, geo_make(Cn, 1, R * Cn, Cn), Cn * S, R, kind, inverse, scale));
  JTB_TRY(e.r2r_lines(a, geo_make(R * Cn, 1, S * R * Cn, R * Cn), R * Cn, S, kind, inverse, scale));
  if (kind == JTB_DHT) {
    grid_for((S / 2 + 1) * (R / 2 + 1) * (Cn / 2 + 1), &g, &b);
    JTB_LAUNCH(k_ytransform3d<T>, g, b, 0, e.st, a, S, R, Cn);
    JTB_CUDA(cudaGetLastError());
    e.ctx->launches++;
  }
  return ST_OK;
}

bool op_valid(const jtb_plan* p, int op) {
  if (p->kind == JTB_FFT) return op >= JTB_C2C_FORWARD && op <= JTB_C2R_FULL;
  return op == JTB_R2R_FORWARD || op == JTB_R2R_INVERSE;
}

template <typename T>
int run_device(jtb_plan* p, int op, T* a, i64 howmany, i64 dist, bool scale, cudaStream_t st) {
  typedef cx<T> C;
  Engine<T> e(p->ctx, st);
  const i64* d = p->dims;
  if (p->rank == 1 && d[0] == 1) return ST_OK;   // n == 1 is a no-op (fft/DoubleFFT_1D.java:248-250)
  // batched 1-D fast paths: one launch sequence over all lines
  if (p->rank == 1 && howmany > 1) {
    const i64 n = d[0];
    switch (op) {
      case JTB_C2C_FORWARD:
      case JTB_C2C_INVERSE:
        if (dist % 2) { set_error("dist must be even for complex transforms"); return ST_ARG; }
        return e.c2c_lines((C*)a, geo_contig(dist / 2), howmany, n, op == JTB_C2C_INVERSE, scale && op == JTB_C2C_INVERSE,
                           (T)(1.0 / (double)n));
      case JTB_R2C_PACKED: return e.real_forward_lines(a, geo_contig(dist), howmany, n);
      case JTB_C2R_PACKED: return e.real_inverse_lines(a, geo_contig(dist), howmany, n, scale);
      case JTB_R2R_FORWARD:
      case JTB_R2R_INVERSE: return e.r2r_lines(a, geo_contig(dist), howmany, n, p->kind, op == JTB_R2R_INVERSE, scale);
      default: break;
    }
  }
  for (i64 b = 0; b < howmany; ++b) {
    T* ab = a + b * dist;
    int s = ST_OK;
    switch (op) {
      case JTB_C2C_FORWARD: s = nd_c2c(e, (C*)ab, p->rank, d, false, false); break;
      case JTB_C2C_INVERSE: s = nd_c2c(e, (C*)ab, p->rank, d, true, scale); break;
      case JTB_R2C_PACKED: s = real_packed(e, ab, p->rank, d, false, false); break;
      case JTB_C2R_PACKED: s = real_packed(e, ab, p->rank, d, true, scale); break;
      case JTB_R2C_FULL: s = real_full(e, ab, p->rank, d, p->total, false, false); break;
      case JTB_C2R_FULL: s = real_full(e, ab, p->rank, d, p->total, true, scale); break;
      case JTB_R2R_FORWARD: s = nd_r2r(e, ab, p->kind, p->rank, d, false, scale); break;
      case JTB_R2R_INVERSE: s = nd_r2r(e, ab, p->kind, p->rank, d, true, scale); break;
      default: set_error("unknown op %d", op); return ST_ARG;
    }
    if (s != ST_OK) return s;
  }
  return ST_OK;
}

int check_plan(const jtb_plan* p, int op) {
  if (!p) { set_error("null plan"); return ST_ARG; }
  if (!op_valid(p, op)) { set_error("op %d is not valid for this plan kind", op); return ST_ARG; }
  return ST_OK;
}

}  // namespace

extern "C" {

int jtb_plan_create(jtb_plan** out, int kind, int prec, int rank, const int64_t* dims, int device) {
  if (!out || !dims) { set_error("null argument"); return ST_ARG; }
  *out = nullptr;
  if (kind < JTB_FFT || kind > JTB_DHT || (prec != JTB_F64 && prec != JTB_F32) || rank < 1 || rank > 3) {
    set_error("bad kind/precision/rank");
    return ST_ARG;
  }
  if (rank == 1) {
    if (dims[0] < 1) { set_error("n must be greater than 0"); return ST_ARG; }
  } else {
    for (int k = 0; k < rank; ++k)
      if (dims[k] <= 1) {
        set_error(rank == 2 ? "rows and columns must be greater than 1" : "slices, rows and columns must be greater than 1");
        return ST_ARG;
      }
  }
  Ctx* ctx = get_ctx(device);
  if (!ctx) return ST_CUDA;
  jtb_plan* p = new jtb_plan();
  p->kind = kind; p->prec = prec; p->rank = rank; p->device = device; p->ctx = ctx;
  p->total = 1;
  for (int k = 0; k < 3; ++k) { p->dims[k] = k < rank ? dims[k] : 1; p->total *= p->dims[k]; }
  *out = p;
  return ST_OK;
}

static void plan_drop_multi(jtb_plan* p) {
  for (size_t g = 0; g < p->slabs.size(); ++g) {
    if (g < p->slab_in.size() && p->slab_in[g]) { DeviceGuard dg(p->devices[g]); cudaFree(p->slab_in[g]); }
    if (p->slabs[g]) jtb_slab_destroy(p->slabs[g]);
  }
  for (size_t g = 0; g < p->mstream.size(); ++g) {
    DeviceGuard dg(p->devices[g]);
    if (p->mstream[g]) cudaStreamDestroy(p->mstream[g]);
    if (g < p->mev.size() && p->mev[g]) cudaEventDestroy(p->mev[g]);
  }
  for (jtb_plan* q : p->sub) jtb_plan_destroy(q);
  p->slabs.clear(); p->slab_in.clear(); p->mstream.clear(); p->mev.clear(); p->sub.clear(); p->devices.clear();
}

int jtb_plan_destroy(jtb_plan* plan) {
  if (!plan) return ST_OK;
  plan_drop_multi(plan);
  {
    // tables this plan was the last user of are freed (fft/DoubleFFT_1D.java: the plan object owns w / bk1 / bk2)
    std::lock_guard<std::mutex> lk(plan->ctx->mu);
    DeviceGuard dg(plan->device);
    plan->ctx->release_tables(plan->keys);
  }
  delete plan;
  return ST_OK;
}

int jtb_plan_device_count(const jtb_plan* plan) {
  if (!plan) return 0;
  return plan->devices.empty() ? 1 : (int)plan->devices.size();
}

int jtb_plan_set_devices(jtb_plan* p, int ndev, const int* devices) {
  if (!p || !devices || ndev < 1 || ndev > 8) { set_error("1..8 devices"); return ST_ARG; }
  for (int g = 0; g < ndev; ++g)
    if (!get_ctx(devices[g])) return ST_CUDA;
  plan_drop_multi(p);
  if (ndev == 1) {
    if (devices[0] != p->device) {
      {
        std::lock_guard<std::mutex> lk(p->ctx->mu);
        DeviceGuard dg(p->device);
        p->ctx->release_tables(p->keys);
        p->keys.clear();
      }
      p->device = devices[0];
      p->ctx = get_ctx(devices[0]);
    }
    return ST_OK;
  }
  p->devices.assign(devices, devices + ndev);
  int rc = ST_OK;
  for (int g = 0; g < ndev && rc == ST_OK; ++g) {
    jtb_plan* q = nullptr;
    const int64_t dd[3] = {(int64_t)p->dims[0], (int64_t)p->dims[1], (int64_t)p->dims[2]};
    rc = jtb_plan_create(&q, p->kind, p->prec, p->rank, dd, devices[g]);
    if (rc == ST_OK) p->sub.push_back(q);
  }
  const bool slabbable = p->kind == JTB_FFT && p->rank == 3 && p->dims[0] % ndev == 0 && p->dims[1] % ndev == 0;
  if (rc == ST_OK && slabbable) {
    const size_t csz = p->prec == JTB_F64 ? 16 : 8;
    const size_t slab_bytes = (size_t)(p->dims[0] / ndev * p->dims[1] * p->dims[2]) * csz;
    p->slabs.assign(ndev, nullptr); p->slab_in.assign(ndev, nullptr); p->mstream.assign(ndev, nullptr); p->mev.assign(ndev, nullptr);
    for (int g = 0; g < ndev && rc == ST_OK; ++g) {
      rc = jtb_slab_create(&p->slabs[g], p->prec, p->dims[0], p->dims[1], p->dims[2], ndev, g, devices[g]);
      if (rc != ST_OK) break;
      DeviceGuard dg(devices[g]);
      cudaError_t e = cudaMalloc(&p->slab_in[g], slab_bytes);
      if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->mstream[g], cudaStreamNonBlocking);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->mev[g], cudaEventDisableTiming);
      if (e != cudaSuccess) rc = cuda_fail(e, "multi-GPU plan buffers");
    }
    if (rc == ST_OK) rc = jtb_slab_connect_local(p->slabs.data(), ndev);
    if (rc == ST_OK) {
      static const char* ex = getenv("JTB_EXCHANGE_NCCL");
      if (ex && atoi(ex)) {
        rc = jtb_slab_nccl_init_local(p->slabs.data(), ndev);
        for (int g = 0; g < ndev && rc == ST_OK; ++g) rc = jtb_slab_set_exchange(p->slabs[g], 1);
      }
    }
  }
  if (rc != ST_OK) {
    const std::string msg = last_error();
    plan_drop_multi(p);
    set_error("%s", msg.c_str());
  }
  return rc;
}

int64_t jtb_plan_elements(const jtb_plan* p, int op) {
  if (!p) return 0;
  switch (op) {
    case JTB_C2C_FORWARD: case JTB_C2C_INVERSE: case JTB_R2C_FULL: case JTB_C2R_FULL: return 2 * p->total;
    default: return p->total;
  }
}

int jtb_exec_device(jtb_plan* p, int op, void* dev_a, int64_t howmany, int64_t dist, int scale, void* stream) {
  JTB_TRY(check_plan(p, op));
  if (!dev_a) { set_error("null data pointer"); return ST_ARG; }
  if (howmany < 1) return ST_OK;
  cudaStream_t st = (cudaStream_t)stream;
  PlanCall call(p, st);
  JTB_TRY(call.begin());
  return p->prec == JTB_F64 ? run_device<double>(p, op, (double*)dev_a, howmany, dist, scale != 0, st)
                            : run_device<float>(p, op, (float*)dev_a, howmany, dist, scale != 0, st);
}

namespace {

// jtb_exec on a multi-GPU rank-3 FFT plan: ONE caller array [S][R][C], slab g over GPU g's PCIe link, the fused
// passes + exchange on every GPU, the k2-slabbed result delivered in natural order by pitched copies.
int exec_multi_fft3d(jtb_plan* p, bool inverse, bool scale, char* h) {
  const int P = (int)p->devices.size();
  const i64 S = p->dims[0], R = p->dims[1], Cn = p->dims[2];
  const size_t csz = p->prec == JTB_F64 ? 16 : 8;
  const size_t slab_bytes = (size_t)(S / P * R * Cn) * csz;
  const size_t width = (size_t)(R / P * Cn) * csz, hpitch = (size_t)(R * Cn) * csz;
  // the contexts involved, each locked once, in device order
  std::vector<Ctx*> ctxs;
  for (int g = 0; g < P; ++g) ctxs.push_back(get_ctx(p->devices[g]));
  std::vector<Ctx*> uniq(ctxs);
  std::sort(uniq.begin(), uniq.end(), [](Ctx* a, Ctx* b) { return a->device < b->device; });
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  std::vector<std::unique_lock<std::mutex>> locks;
  for (Ctx* c : uniq) locks.emplace_back(c->mu);
  for (Ctx* c : uniq) c->recorder = &p->keys;
  struct Unrec { std::vector<Ctx*>& u; ~Unrec() { for (Ctx* c : u) c->recorder = nullptr; } } unrec{uniq};
  DeviceGuard restore(p->devices[0]);

  const char* est = getenv("JTB_STAGE");
  const bool pageable = (!est || atoi(est) != 0) && host_is_pageable(h);
  int nt = 1;
  {
    const char* et = getenv("JTB_STAGE_THREADS");
    const int total = et ? atoi(et) : (int)std::thread::hardware_concurrency();
    nt = std::max(1, total / P);
  }
  auto fanout = [&](const std::function<int(int)>& fn) -> int {   // one host thread per member; first failure wins
    std::vector<int> rc(P, ST_OK);
    std::vector<std::string> msg(P);
    std::vector<std::thread> th;
    for (int g = 0; g < P; ++g)
      th.emplace_back([&, g]() { rc[g] = fn(g); if (rc[g] != ST_OK) msg[g] = last_error(); });
    for (auto& t : th) t.join();
    for (int g = 0; g < P; ++g)
      if (rc[g] != ST_OK) { set_error("%s", msg[g].c_str()); return rc[g]; }
    return ST_OK;
  };
  // host -> devices
  if (pageable) {
    JTB_TRY(fanout([&](int g) {
      return staged_copy_2d(p->devices[g], p->slab_in[g], h + (size_t)g * slab_bytes, slab_bytes, slab_bytes, 1, true, nullptr, nt);
    }));
  } else {
    for (int g = 0; g < P; ++g) {
      JTB_CUDA(cudaSetDevice(p->devices[g]));
      JTB_CUDA(cudaMemcpyAsync(p->slab_in[g], h + (size_t)g * slab_bytes, slab_bytes, cudaMemcpyHostToDevice, p->mstream[g]));
    }
  }
  // transform
  void* results[8];
  JTB_TRY(slab_group_run(p->slabs.data(), P, p->slab_in.data(), false, inverse, scale, results, p->mstream.data()));
  // devices -> host, natural order: rank g holds rows [g*R/P, (g+1)*R/P) of every slice
  int rc = ST_OK;
  if (pageable) {
    for (int g = 0; g < P; ++g) {
      JTB_CUDA(cudaSetDevice(p->devices[g]));
      JTB_CUDA(cudaEventRecord(p->mev[g], p->mstream[g]));
    }
    rc = fanout([&](int g) {
      return staged_copy_2d(p->devices[g], results[g], h + (size_t)g * width, hpitch, width, (size_t)S, false, p->mev[g], nt);
    });
  } else {
    for (int g = 0; g < P; ++g) {
      JTB_CUDA(cudaSetDevice(p->devices[g]));
      JTB_CUDA(cudaMemcpy2DAsync(h + (size_t)g * width, hpitch, results[g], width, width, (size_t)S, cudaMemcpyDeviceToHost,
                                 p->mstream[g]));
    }
  }
  for (int g = 0; g < P; ++g) {
    cudaSetDevice(p->devices[g]);
    const cudaError_t e = cudaStreamSynchronize(p->mstream[g]);
    if (e != cudaSuccess && rc == ST_OK) rc = cuda_fail(e, "multi-GPU transform");
  }
  for (Ctx* c : uniq)
    if (rc == ST_OK) rc = c->check_watchdog("multi-GPU 3-D transform");
  return rc;
}

// jtb_exec_batch on a multi-GPU plan: contiguous blocks of the batch, one per GPU, each through that GPU's own
// single-device pipeline on its own host thread (no collective)
int exec_multi_batch(jtb_plan* p, int op, void* host_a, i64 offa, i64 howmany, i64 dist, int scale) {
  const int P = (int)p->sub.size();
  std::vector<int> rc(P, ST_OK);
  std::vector<std::string> msg(P);
  std::vector<std::thread> th;
  const i64 per = (howmany + P - 1) / P;
  for (int g = 0; g < P; ++g) {
    const i64 b0 = (i64)g * per, cnt = std::min(per, howmany - b0);
    if (cnt <= 0) break;
    th.emplace_back([&, g, b0, cnt]() {
      rc[g] = jtb_exec_batch(p->sub[g], op, host_a, offa + b0 * dist, cnt, dist, scale);
      if (rc[g] != ST_OK) msg[g] = last_error();
    });
  }
  for (auto& t : th) t.join();
  for (int g = 0; g < P; ++g)
    if (rc[g] != ST_OK) { set_error("%s", msg[g].c_str()); return rc[g]; }
  return ST_OK;
}

}  // namespace

int jtb_exec_batch(jtb_plan* p, int op, void* host_a, int64_t offa, int64_t howmany, int64_t dist, int scale) {
  JTB_TRY(check_plan(p, op));
  if (!host_a) { set_error("null data pointer"); return ST_ARG; }
  if (howmany < 1) return ST_OK;
  if (offa < 0) { set_error("negative offset"); return ST_ARG; }
  const i64 elems = jtb_plan_elements(p, op);
  if (howmany > 1 && dist < elems) { set_error("dist smaller than one transform"); return ST_ARG; }
  const size_t esz = p->prec == JTB_F64 ? 8 : 4;
  const i64 span = (howmany - 1) * dist + elems;
  const bool full = op == JTB_R2C_FULL || op == JTB_C2R_FULL;
  const i64 in_span = (full && howmany == 1) ? p->total : span;
  char* h = (char*)host_a + (size_t)offa * esz;
  if (!p->devices.empty()) {
    if (howmany == 1 && !p->slabs.empty() && (op == JTB_C2C_FORWARD || op == JTB_C2C_INVERSE))
      return exec_multi_fft3d(p, op == JTB_C2C_INVERSE, scale != 0, h);
    if (howmany >= (i64)p->sub.size()) return exec_multi_batch(p, op, host_a, offa, howmany, dist, scale);
    return jtb_exec_batch(p->sub[0], op, host_a, offa, howmany, dist, scale);   // everything else: devices[0]
  }
  Ctx* c = p->ctx;
  PlanCall call(p, c->stream);
  JTB_TRY(call.begin());
  // Large batches run as a three-stage pipeline over chunks of whole transforms: H2D of chunk i+1, the kernels of
  // chunk i and D2H of chunk i-1 overlap (PCIe is full duplex), and the device copy is three chunks instead of the
  // whole span -- the chunked staging of SURVEY.md 8(f) rank 4.
  // Pageable caller memory (Java heap arrays, plain numpy arrays): multi-threaded staging through page-locked bounce
  // buffers instead of the driver's single-threaded pageable path (jtb_stage.cu); JTB_STAGE=0 disables it.
  {
    const char* est = getenv("JTB_STAGE");
    const char* emin = getenv("JTB_STAGE_MIN_MB");    // smaller arrays stay on the driver's own path (default 32 MiB)
    const size_t bytes = (size_t)span * esz;
    const size_t min_bytes = (size_t)((emin ? atof(emin) : 32.0) * 1048576.0);
    if ((!est || atoi(est) != 0) && bytes >= min_bytes && host_is_pageable(h)) {
      JTB_TRY(c->ensure_pipeline());
      JTB_TRY(c->ensure(c->io, bytes));
      JTB_TRY(staged_copy(p->device, c->io.p, h, (size_t)in_span * esz, true, nullptr));
      int s = p->prec == JTB_F64 ? run_device<double>(p, op, (double*)c->io.p, howmany, dist, scale != 0, c->stream)
                                 : run_device<float>(p, op, (float*)c->io.p, howmany, dist, scale != 0, c->stream);
      if (s != ST_OK) { cudaStreamSynchronize(c->stream); return s; }
      JTB_CUDA(cudaEventRecord(c->ev_c[0], c->stream));
      return staged_copy(p->device, c->io.p, h, bytes, false, c->ev_c[0]);
    }
  }
  {
    const char* emb = getenv("JTB_BATCH_MB");   // chunk size of the pipelined path in MiB; 0 disables it
    const double chunk_mb = emb ? atof(emb) : 64.0;   // measured: 2 GB batch e2e 83.9 ms unpipelined, 53.2 ms at 256 MiB, 46.1 ms at 64 MiB
    const size_t per = (size_t)dist * esz;
    i64 nb = per ? (i64)(chunk_mb * 1048576.0 / (double)per) : howmany;
    if (nb < 1) nb = 1;
    if (chunk_mb > 0 && howmany > 1 && howmany >= 3 * nb) {
      JTB_TRY(c->ensure_pipeline());
      const size_t chunk_elems = (size_t)((nb - 1) * dist + elems);
      const size_t chunk_bytes = (chunk_elems * esz + 255) / 256 * 256;
      JTB_TRY(c->ensure(c->io, 3 * chunk_bytes));
      int s = ST_OK;
      i64 idx = 0;
      for (i64 b0 = 0; b0 < howmany && s == ST_OK; b0 += nb, ++idx) {
        const i64 cnt = howmany - b0 < nb ? howmany - b0 : nb;
        const int slot = (int)(idx % 3);
        char* dev = (char*)c->io.p + (size_t)slot * chunk_bytes;
        char* hb = h + (size_t)b0 * per;
        const size_t bytes = (size_t)((cnt - 1) * dist + elems) * esz;
        if (idx >= 3) JTB_CUDA(cudaStreamWaitEvent(c->s_in, c->ev_out[slot], 0));   // slot's previous result is back
        JTB_CUDA(cudaMemcpyAsync(dev, hb, bytes, cudaMemcpyHostToDevice, c->s_in));
        JTB_CUDA(cudaEventRecord(c->ev_in[slot], c->s_in));
        JTB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_in[slot], 0));
        s = p->prec == JTB_F64 ? run_device<double>(p, op, (double*)dev, cnt, dist, scale != 0, c->stream)
                               : run_device<float>(p, op, (float*)dev, cnt, dist, scale != 0, c->stream);
        if (s != ST_OK) break;
        JTB_CUDA(cudaEventRecord(c->ev_c[slot], c->stream));
        JTB_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_c[slot], 0));
        JTB_CUDA(cudaMemcpyAsync(hb, dev, bytes, cudaMemcpyDeviceToHost, c->s_out));
        JTB_CUDA(cudaEventRecord(c->ev_out[slot], c->s_out));
      }
      cudaStreamSynchronize(c->s_in);
      cudaStreamSynchronize(c->stream);
      JTB_CUDA(cudaStreamSynchronize(c->s_out));
      return s;
    }
  }
  JTB_TRY(c->ensure(c->io, (size_t)span * esz));
  JTB_CUDA(cudaMemcpyAsync(c->io.p, h, (size_t)in_span * esz, cudaMemcpyHostToDevice, c->stream));
  int s = p->prec == JTB_F64 ? run_device<double>(p, op, (double*)c->io.p, howmany, dist, scale != 0, c->stream)
                             : run_device<float>(p, op, (float*)c->io.p, howmany, dist, scale != 0, c->stream);
  if (s != ST_OK) { cudaStreamSynchronize(c->stream); return s; }
  JTB_CUDA(cudaMemcpyAsync(h, c->io.p, (size_t)span * esz, cudaMemcpyDeviceToHost, c->stream));
  JTB_CUDA(cudaStreamSynchronize(c->stream));
  return ST_OK;
}

int jtb_exec(jtb_plan* p, int op, void* host_a, int64_t offa, int scale) {
  return jtb_exec_batch(p, op, host_a, offa, 1, 0, scale);
}

int jtb_exec_n(jtb_plan* p, int op, void* host_a, int64_t a_length, int64_t offa, int scale) {
  JTB_TRY(check_plan(p, op));
  if (offa < 0) { set_error("negative offset"); return ST_ARG; }
  const i64 need = offa + jtb_plan_elements(p, op);
  if (a_length < need) {
    set_error("array too short: length %lld, the transform touches elements [%lld, %lld)", (long long)a_length,
              (long long)offa, (long long)need);
    return ST_ARG;
  }
  return jtb_exec_batch(p, op, host_a, offa, 1, 0, scale);
}

int jtb_lines_c2c_device(int prec, int device, void* dev_a, int64_t n, int64_t nlines, int64_t c0, int64_t d0,
                         int64_t d3, int64_t stride, int inverse, double scale, void* stream) {
  if (!dev_a || n < 1 || nlines < 0 || c0 < 1 || stride < 1) { set_error("bad line geometry"); return ST_ARG; }
  Ctx* c = get_ctx(device);
  if (!c) return ST_CUDA;
  CtxCall call(c, (cudaStream_t)stream);
  JTB_TRY(call.begin());
  Geo g = geo_make(c0, d0, d3, stride);
  const bool has_scale = scale != 1.0;
  if (prec == JTB_F64) {
    Engine<double> e(c, (cudaStream_t)stream);
    return e.c2c_lines((double2*)dev_a, g, nlines, n, inverse != 0, has_scale, scale);
  }
  Engine<float> e(c, (cudaStream_t)stream);
  return e.c2c_lines((float2*)dev_a, g, nlines, n, inverse != 0, has_scale, (float)scale);
}

int jtb_lines_c2c_out_device(int prec, int device, void* dev_in, void* dev_out, int64_t n, int64_t nlines, int64_t c0,
                             int64_t d3, int64_t stride, int64_t out_d3, int64_t out_stride, int inverse, double scale,
                             void* stream) {
  if (!dev_in || !dev_out || n < 2 || nlines < 1 || c0 < 1 || stride < 1 || out_stride < 1) { set_error("bad line geometry"); return ST_ARG; }
  if (!is_pow2(n)) { set_error("out-of-place strided lines need a power-of-two length"); return ST_UNSUPPORTED; }
  Ctx* c = get_ctx(device);
  if (!c) return ST_CUDA;
  CtxCall call(c, (cudaStream_t)stream);
  JTB_TRY(call.begin());
  const Geo g = geo_make(c0, 1, d3, stride);
  bool handled = false;
  if (prec == JTB_F64) {
    Engine<double> e(c, (cudaStream_t)stream);
    JTB_TRY(fast_c2c_out<double>(e, (double2*)dev_in, g, (double2*)dev_out, out_d3, out_stride, nlines, ilog2(n), inverse != 0,
                                 scale != 1.0, scale, &handled));
  } else {
    Engine<float> e(c, (cudaStream_t)stream);
    JTB_TRY(fast_c2c_out<float>(e, (float2*)dev_in, g, (float2*)dev_out, out_d3, out_stride, nlines, ilog2(n), inverse != 0,
                                scale != 1.0, (float)scale, &handled));
  }
  if (!handled) { set_error("no lean strided kernel for this shape"); return ST_UNSUPPORTED; }
  return ST_OK;
}

int jtb_fft2d_slices_device(int prec, int device, void* dev_a, int64_t nslices, int64_t rows, int64_t cols, int nranks,
                            int rank, void* const* recv_ptrs, int inverse, void* stream) {
  if (!dev_a || nslices < 1 || rows < 2 || cols < 2) { set_error("bad argument"); return ST_ARG; }
  Ctx* c = get_ctx(device);
  if (!c) return ST_CUDA;
  CtxCall call(c, (cudaStream_t)stream);
  JTB_TRY(call.begin());
  if (prec != JTB_F64 && prec != JTB_F32) { set_error("bad precision"); return ST_ARG; }
  bool fused = false;
  if (prec == JTB_F64) {
    Engine<double> e(c, (cudaStream_t)stream);
    if (rows == cols) JTB_TRY(fast_slice2d<double>(e, (double2*)dev_a, nslices, rows, inverse != 0, false, 1.0, nranks, rank, recv_ptrs, &fused));
    if (fused) return ST_OK;
    JTB_TRY(e.c2c_lines((double2*)dev_a, geo_contig(cols), nslices * rows, cols, inverse != 0, false, 1.0));
    if (recv_ptrs) return fast_scatter<double>(e, (const double2*)dev_a, nslices, rows, cols, nranks, rank, recv_ptrs, inverse != 0);
    return e.c2c_lines((double2*)dev_a, geo_make(cols, 1, rows * cols, cols), cols * nslices, rows, inverse != 0, false, 1.0);
  }
  Engine<float> e(c, (cudaStream_t)stream);
  JTB_TRY(e.c2c_lines((float2*)dev_a, geo_contig(cols), nslices * rows, cols, inverse != 0, false, 1.0f));
  if (recv_ptrs) return fast_scatter<float>(e, (const float2*)dev_a, nslices, rows, cols, nranks, rank, recv_ptrs, inverse != 0);
  return e.c2c_lines((float2*)dev_a, geo_make(cols, 1, rows * cols, cols), cols * nslices, rows, inverse != 0, false, 1.0f);
}

int jtb_fft3d_k2_scatter(int prec, int device, const void* local_a, int64_t Ls, int64_t R, int64_t Cn, int nranks,
                         int rank, void* const* recv_ptrs, int inverse, void* stream) {
  if (!local_a || !recv_ptrs || Ls < 1 || R < 2 || Cn < 1 || rank < 0 || rank >= nranks) { set_error("bad argument"); return ST_ARG; }
  Ctx* c = get_ctx(device);
  if (!c) return ST_CUDA;
  CtxCall call(c, (cudaStream_t)stream);
  JTB_TRY(call.begin());
  if (prec == JTB_F64) {
    Engine<double> e(c, (cudaStream_t)stream);
    return fast_scatter<double>(e, (const double2*)local_a, Ls, R, Cn, nranks, rank, recv_ptrs, inverse != 0);
  }
  Engine<float> e(c, (cudaStream_t)stream);
  return fast_scatter<float>(e, (const float2*)local_a, Ls, R, Cn, nranks, rank, recv_ptrs, inverse != 0);
}

int jtb_fft3d_k2_scatter_chunk(int prec, int device, const void* local_a, int64_t Ls, int64_t slice_base, int64_t R,
                               int64_t Cn, int nranks, void* const* recv_ptrs, int inverse, void* stream) {
  if (!local_a || !recv_ptrs || Ls < 1 || R < 2 || Cn < 1 || slice_base < 0) { set_error("bad argument"); return ST_ARG; }
  Ctx* c = get_ctx(device);
  if (!c) return ST_CUDA;
  CtxCall call(c, (cudaStream_t)stream);
  JTB_TRY(call.begin());
  if (prec == JTB_F64) {
    Engine<double> e(c, (cudaStream_t)stream);
    return fast_scatter<double>(e, (const double2*)local_a, Ls, R, Cn, nranks, 0, recv_ptrs, inverse != 0, slice_base);
  }
  Engine<float> e(c, (cudaStream_t)stream);
  return fast_scatter<float>(e, (const float2*)local_a, Ls, R, Cn, nranks, 0, recv_ptrs, inverse != 0, slice_base);
}

int jtb_fft3d_k1_scatter(int prec, int device, const void* local_b, int64_t S, int64_t Rh, int64_t Cn, int nranks,
                         int rank, void* const* recv_ptrs, int inverse, void* stream) {
  if (!local_b || !recv_ptrs || S < 2 || Rh < 1 || Cn < 1 || rank < 0 || rank >= nranks) { set_error("bad argument"); return ST_ARG; }
  Ctx* c = get_ctx(device);
  if (!c) return ST_CUDA;
  CtxCall call(c, (cudaStream_t)stream);
  JTB_TRY(call.begin());
  if (prec == JTB_F64) {
    Engine<double> e(c, (cudaStream_t)stream);
    return fast_scatter<double>(e, (const double2*)local_b, 1, S, Rh * Cn, nranks, rank, recv_ptrs, inverse != 0, 0, true);
  }
  Engine<float> e(c, (cudaStream_t)stream);
  return fast_scatter<float>(e, (const float2*)local_b, 1, S, Rh * Cn, nranks, rank, recv_ptrs, inverse != 0, 0, true);
}

int jtb_peer_barrier(int device, void* const* flag_ptrs, int nranks, int rank, int64_t epoch, void* stream) {
  Ctx* c = get_ctx(device);
  if (!c) return ST_CUDA;
  if (!flag_ptrs || rank < 0 || rank >= nranks) { set_error("bad argument"); return ST_ARG; }
  DeviceGuard dg(device);
  std::lock_guard<std::mutex> lk(c->mu);
  return peer_barrier(c, (cudaStream_t)stream, flag_ptrs, nranks, rank, (long long)epoch);
}

int jtb_peer_alloc(int device, int64_t bytes, void** dev_ptr, unsigned char* handle64) {
  if (!dev_ptr || !handle64 || bytes < 1) { set_error("bad argument"); return ST_ARG; }
  if (!get_ctx(device)) return ST_CUDA;
  DeviceGuard dg(device);
  JTB_CUDA(cudaMalloc(dev_ptr, (size_t)bytes));
  JTB_CUDA(cudaMemset(*dev_ptr, 0, (size_t)bytes));
  memset(handle64, 0, 64);
#ifdef JTB_EMU
  memcpy(handle64, dev_ptr, sizeof(void*));
#else
  cudaIpcMemHandle_t h;
  JTB_CUDA(cudaIpcGetMemHandle(&h, *dev_ptr));
  static_assert(sizeof(h) == 64, "ipc handle size");
  memcpy(handle64, &h, 64);
#endif
  return ST_OK;
}
int jtb_peer_open(int device, const unsigned char* handle64, void** peer_ptr) {
  if (!handle64 || !peer_ptr) { set_error("bad argument"); return ST_ARG; }
  if (!get_ctx(device)) return ST_CUDA;
  DeviceGuard dg(device);
#ifdef JTB_EMU
  memcpy(peer_ptr, handle64, sizeof(void*));
#else
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  JTB_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
#endif
  return ST_OK;
}
int jtb_peer_close(int device, void* peer_ptr) {
  if (!get_ctx(device)) return ST_CUDA;
#ifndef JTB_EMU
  DeviceGuard dg(device);
  if (peer_ptr) JTB_CUDA(cudaIpcCloseMemHandle(peer_ptr));
#endif
  return ST_OK;
}
int jtb_peer_free(int device, void* dev_ptr) {
  if (!get_ctx(device)) return ST_CUDA;
  DeviceGuard dg(device);
  if (dev_ptr) JTB_CUDA(cudaFree(dev_ptr));
  return ST_OK;
}

int jtb_host_alloc(void** out, int64_t bytes) {
  if (!out || bytes < 0) { set_error("bad argument"); return ST_ARG; }
  if (!get_ctx(0)) return ST_CUDA;
  JTB_CUDA(cudaHostAlloc(out, (size_t)(bytes < 16 ? 16 : bytes), cudaHostAllocDefault));
  return ST_OK;
}
int jtb_host_free(void* p) {
  if (p) JTB_CUDA(cudaFreeHost(p));
  return ST_OK;
}

int jtb_host_register(void* p, int64_t bytes) {
  if (!p || bytes <= 0) { set_error("bad argument"); return ST_ARG; }
  if (!get_ctx(0)) return ST_CUDA;
  JTB_CUDA(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault));
  return ST_OK;
}
int jtb_host_unregister(void* p) {
  if (p) JTB_CUDA(cudaHostUnregister(p));
  return ST_OK;
}

int jtb_fill_uniform_device(int prec, int device, void* dev_a, int64_t count, uint64_t seed, double lo, double hi,
                            void* stream) {
  Ctx* c = get_ctx(device);
  if (!c) return ST_CUDA;
  if (!dev_a || count < 0) { set_error("bad argument"); return ST_ARG; }
  DeviceGuard dg(device);
  unsigned g, b;
  grid_for(count, &g, &b);
  if (prec == JTB_F64) JTB_LAUNCH(k_fill_uniform<double>, g, b, 0, (cudaStream_t)stream, (double*)dev_a, count, (unsigned long long)seed, lo, hi);
  else JTB_LAUNCH(k_fill_uniform<float>, g, b, 0, (cudaStream_t)stream, (float*)dev_a, count, (unsigned long long)seed, (float)lo, (float)hi);
  JTB_CUDA(cudaGetLastError());
  c->launches++;
  return ST_OK;
}

int jtb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
int64_t jtb_launch_count(int device) {
  Ctx* c = get_ctx(device);
  return c ? c->launches : -1;
}
int64_t jtb_debug_table_bytes(int device) {
  Ctx* c = get_ctx(device);
  if (!c) return -1;
  std::lock_guard<std::mutex> lk(c->mu);
  size_t b = 0;
  for (const auto& kv : c->tables) b += kv.second.bytes;
  return (int64_t)b;
}
int jtb_debug_set_limits(int logn_contig, int logn_strided) {
  g_limit_contig = logn_contig;
  g_limit_strided = logn_strided;
  return ST_OK;
}
const char* jtb_last_error(void) { return last_error(); }
const char* jtb_version(void) { return "jtb200 0.1 (sm_100a)"; }

}  // extern "C"
