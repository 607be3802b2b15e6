// Slab-decomposed 3-D complex transform over P GPUs inside libjtb200 (no torch, no Python on the data path).
//
// One jtb_slab is one member (rank) of the decomposition: rank g owns slices [g*Ls, (g+1)*Ls) as [Ls][R][C].
//   forward:  k3 pass + k2 pass on the local slab, the k2 stores being the re-slabbing all-to-all (each output row
//             goes straight into the receive buffer of the GPU that owns it -- NVLink peer stores), synchronisation,
//             k1 pass on the received [S][R/P][C] block.  The result stays k2-slabbed; the host path of jtb_exec
//             delivers it in natural order with a pitched device-to-host copy.
//   back:     k1 pass of a k2-slabbed block with its stores going to the owners' [Ls][R][C] slabs, then k2 and k3.
// Members connect either inside one process (cudaDeviceEnablePeerAccess; events order the exchange -- this is what
// jtb_plan_set_devices builds, so DoubleFFT_3D.complexForward(double[]) on one host array uses all GPUs) or across
// processes (CUDA IPC handles exchanged by the caller; a device-side flag barrier orders the exchange).
// Exchange variants: fused peer stores (kernels of jtb_fast.cuh) for the power-of-two shapes, a generic peer-store
// row copy for every other shape, or NCCL ncclSend/ncclRecv on the strided sub-blocks (JTB_EXCHANGE_NCCL; libnccl is
// dlopen'ed, never linked).
//
// Replaces the slice-axis gather of cdft3db_subth (fft/DoubleFFT_3D.java:6318-6520) and the thread-pool
// partitioning of fft/DoubleFFT_3D.java:145-325 when one transform is spread over several devices.
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "../../include/jtb200.h"
#include "jtb_engine.h"
#include "jtb_slab.h"

#ifndef JTB_EMU
#include <nccl.h>
#endif

using namespace jtb;

// ------------------------------------------------------------------------------------------------ NCCL (dlopen)
namespace {
#ifndef JTB_EMU
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (!tried) {
    tried = true;
    const char* names[] = {getenv("JTB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (api.h) {
      bool ok = true;
      auto sym = [&](const char* s) { void* p = dlsym(api.h, s); if (!p) ok = false; return p; };
      api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
      api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
      api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
      api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
      api.Send = (decltype(api.Send))sym("ncclSend");
      api.Recv = (decltype(api.Recv))sym("ncclRecv");
      api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
      if (!ok) { dlclose(api.h); api.h = nullptr; }
    }
  }
  return api.h ? &api : nullptr;
}
#define JTB_NCCL(expr)                                                                          \
  do {                                                                                          \
    ncclResult_t _r = (expr);                                                                   \
    if (_r != ncclSuccess) {                                                                    \
      set_error("NCCL error '%s' in %s", nccl_api()->GetErrorString(_r), #expr);                \
      return ST_NCCL;                                                                           \
    }                                                                                           \
  } while (0)
#endif
}  // namespace

// ------------------------------------------------------------------------------------------------ kernels
namespace jtb {

template <typename C> struct SlabPeers { C* p[8]; };

// Generic re-slabbing by peer stores: row (i1, i2) of the local block goes to the buffer of the GPU that owns it.
//   mode 0 (forward, block [Ls][R][C]):  peer = r / Rh,   row (slice0 + ls)*Rh + r % Rh   of the peer's [S][Rh][C]
//   mode 1 (back,    block [S][Rh][C]):  peer = k1 / Ls,  row (k1 % Ls)*R + rank*Rh + rl  of the peer's [Ls][R][C]
template <typename C>
__global__ void k_slab_rows_to_peers(const C* __restrict__ a, const SlabPeers<C> peers, i64 n1, i64 n2, i64 Cn, int mode,
                                     i64 Rh, i64 Ls, i64 R, i64 base) {
  const i64 nrows = n1 * n2;
  for (i64 row = blockIdx.x; row < nrows; row += gridDim.x) {
    const i64 i1 = row / n2, i2 = row - i1 * n2;
    int peer;
    i64 drow;
    if (mode == 0) { peer = (int)(i2 / Rh); drow = (base + i1) * Rh + (i2 - (i64)peer * Rh); }
    else { peer = (int)(i1 / Ls); drow = (i1 - (i64)peer * Ls) * R + base * Rh + i2; }
    const C* src = a + row * Cn;
    C* dst = peers.p[peer] + drow * Cn;
    for (i64 c = threadIdx.x; c < Cn; c += blockDim.x) dst[c] = src[c];
  }
}

}  // namespace jtb

// ------------------------------------------------------------------------------------------------ the member
struct jtb_slab {
  int prec = 0, P = 1, rank = 0, device = 0;
  i64 S = 0, R = 0, Cn = 0, Ls = 0, Rh = 0;
  Ctx* ctx = nullptr;
  size_t csz = 16, block_bytes = 0;
  void* recv[2] = {nullptr, nullptr};
  long long* flags = nullptr;            // int64[8], epoch published by every peer
  void* peer_recv[2][8];
  long long* peer_flags[8];
  unsigned char ipc_open[8];             // this peer's three mappings came from cudaIpcOpenMemHandle
  int mode = 0;                          // 0 unconnected, 1 IPC peers (flag barrier), 2 same-process group (events)
  jtb_slab* group[8];
  cudaEvent_t ev = nullptr;
  long long epoch = 0;
  int step = 0;
  int exchange = 0;                      // 0 peer stores, 1 NCCL
  void* comm = nullptr;
  // pipelined exchange (forward, fused peer stores): the slab is sent column block by column block and the slice-axis
  // pass of block j runs on `st2` as soon as block j has arrived from every peer, under the stores of blocks j+1..
  int nchunks = 1;                       // column blocks per step (JTB_SLAB_CHUNKS; 1 = one exchange, then the k1 pass)
  int fused_blocks = 0;                  // > 0: exchange + slice-axis pass in one persistent kernel (JTB_SLAB_FUSED)
  int* pipe_counters = nullptr;          // work queue + per-block tile counters of fft_pipe_kernel
  cudaStream_t st2 = nullptr;            // high-priority stream of the consumer side
  cudaEvent_t evc[16];                   // same-process groups: block j of this member has been stored
  cudaEvent_t ev_join = nullptr;
  long long chunk_epoch[16];             // IPC peers: flag value that announces block j of the current step
  // optional phase timing (jtb_slab_profile): events at the phase boundaries of the last step on its stream
  bool profile = false;
  cudaEvent_t pev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t pev_s[16], pev_k[16];      // pipelined step: block j stored (producer stream) / block j transformed (consumer)
  int last_nb = 0;
  int mark_chunk(cudaEvent_t* arr, int j, cudaStream_t st) {
    if (!profile) return ST_OK;
    if (!arr[j]) JTB_CUDA(cudaEventCreate(&arr[j]));
    JTB_CUDA(cudaEventRecord(arr[j], st));
    return ST_OK;
  }
  int mark(int i, cudaStream_t st) {
    if (!profile) return ST_OK;
    if (!pev[i]) JTB_CUDA(cudaEventCreate(&pev[i]));
    JTB_CUDA(cudaEventRecord(pev[i], st));
    return ST_OK;
  }
  jtb_slab() {
    memset(peer_recv, 0, sizeof peer_recv);
    memset(peer_flags, 0, sizeof peer_flags);
    memset(ipc_open, 0, sizeof ipc_open);
    memset(group, 0, sizeof group);
    memset(evc, 0, sizeof evc);
    memset(pev_s, 0, sizeof pev_s);
    memset(pev_k, 0, sizeof pev_k);
    memset(chunk_epoch, 0, sizeof chunk_epoch);
  }
};

namespace {

template <typename T> int slab_rows_to_peers(jtb_slab* m, Engine<T>& e, const cx<T>* a, int buf, bool back) {
  typedef cx<T> C;
  SlabPeers<C> pp;
  for (int h = 0; h < 8; ++h) pp.p[h] = h < m->P ? (C*)m->peer_recv[buf][h] : nullptr;
  const i64 n1 = back ? m->S : m->Ls, n2 = back ? m->Rh : m->R;
  i64 grid = n1 * n2;
  if (grid > 148 * 16) grid = 148 * 16;
  const unsigned block = m->Cn >= 256 ? 256u : (m->Cn >= 64 ? 64u : 32u);
  JTB_LAUNCH(k_slab_rows_to_peers<C>, (unsigned)grid, block, 0, e.st, a, pp, n1, n2, m->Cn, back ? 1 : 0, m->Rh, m->Ls,
             m->R, back ? (i64)m->rank : (i64)m->rank * m->Ls);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  return ST_OK;
}

// NCCL variant of the exchange for one member (the caller brackets the members of a same-process group with
// ncclGroupStart/End): Ls (forward) sub-blocks of Rh*C elements per peer, straight from / into the strided arrays.
template <typename T> int slab_nccl_enqueue(jtb_slab* m, const cx<T>* send, cx<T>* recvbuf, bool back, cudaStream_t st) {
#ifdef JTB_EMU
  (void)m; (void)send; (void)recvbuf; (void)back; (void)st;
  set_error("NCCL is not available in the emulated build");
  return ST_NCCL;
#else
  NcclApi* api = nccl_api();
  if (!api || !m->comm) { set_error("NCCL exchange requested but no communicator (jtb_slab_nccl_init)"); return ST_NCCL; }
  const ncclDataType_t dt = sizeof(T) == 8 ? ncclDouble : ncclFloat;
  const size_t cnt = (size_t)(2 * m->Rh * m->Cn);     // reals per sub-block
  const i64 sub = m->Rh * m->Cn;                      // complex elements per sub-block
  ncclComm_t comm = (ncclComm_t)m->comm;
  for (int h = 0; h < m->P; ++h)
    for (i64 ls = 0; ls < m->Ls; ++ls) {
      if (!back) {
        // send rows [h*Rh, (h+1)*Rh) of local slice ls; receive slice ls of rank h into [h*Ls + ls][Rh][C]
        JTB_NCCL(api->Send(send + (ls * m->R + (i64)h * m->Rh) * m->Cn, cnt, dt, h, comm, st));
        JTB_NCCL(api->Recv(recvbuf + ((i64)h * m->Ls + ls) * sub, cnt, dt, h, comm, st));
      } else {
        // send slice h*Ls + ls of the k2-slabbed block; receive rows [h*Rh, (h+1)*Rh) of local slice ls
        JTB_NCCL(api->Send(send + ((i64)h * m->Ls + ls) * sub, cnt, dt, h, comm, st));
        JTB_NCCL(api->Recv(recvbuf + (ls * m->R + (i64)h * m->Rh) * m->Cn, cnt, dt, h, comm, st));
      }
    }
  return ST_OK;
#endif
}

// phase A: the two in-slice passes of the local slab; with peer stores the second one IS the exchange
template <typename T> int slab_phase_a(jtb_slab* m, cx<T>* a, bool inverse, cudaStream_t st, int buf, bool* exchanged) {
  typedef cx<T> C;
  Engine<T> e(m->ctx, st);
  const i64 Ls = m->Ls, R = m->R, Cn = m->Cn;
  *exchanged = false;
  if (m->P > 1 && m->exchange == 0) {
    bool fused = false;
    if (R == Cn) JTB_TRY(fast_slice2d<T>(e, a, Ls, R, inverse, false, (T)1, m->P, m->rank, m->peer_recv[buf], &fused));
    if (fused) { *exchanged = true; return ST_OK; }
  }
  JTB_TRY(e.c2c_lines(a, geo_contig(Cn), Ls * R, Cn, inverse, false, (T)1));                   // k3: contiguous rows
  if (m->P > 1 && m->exchange == 0) {
    const int rc = fast_scatter<T>(e, a, Ls, R, Cn, m->P, m->rank, m->peer_recv[buf], inverse);  // k2 + exchange
    if (rc == ST_OK) { *exchanged = true; return ST_OK; }
    if (rc != ST_UNSUPPORTED) return rc;
  }
  JTB_TRY(e.c2c_lines(a, geo_make(Cn, 1, R * Cn, Cn), Cn * Ls, R, inverse, false, (T)1));      // k2: columns in place
  if (m->P > 1 && m->exchange == 0) {
    JTB_TRY(slab_rows_to_peers<T>(m, e, a, buf, false));
    *exchanged = true;
  }
  return ST_OK;
}

// phase B: the slice-axis pass on the re-slabbed block [S][Rh][C]
template <typename T> int slab_phase_b(jtb_slab* m, cx<T>* b, bool inverse, bool scale, cudaStream_t st) {
  Engine<T> e(m->ctx, st);
  const i64 S = m->S, Rh = m->Rh, Cn = m->Cn;
  return e.c2c_lines(b, geo_make(Rh * Cn, 1, S * Rh * Cn, Rh * Cn), Rh * Cn, S, inverse, scale,
                     (T)(1.0 / ((double)S * (double)m->R * (double)Cn)));
}

// the way back, phase A: k1 pass of the k2-slabbed block, stores into the owners' slabs
template <typename T> int slab_back_a(jtb_slab* m, cx<T>* b, cudaStream_t st, int buf, bool* exchanged) {
  Engine<T> e(m->ctx, st);
  const i64 S = m->S, Rh = m->Rh, Cn = m->Cn;
  *exchanged = false;
  if (m->exchange == 0) {
    const int rc = fast_scatter<T>(e, b, 1, S, Rh * Cn, m->P, m->rank, m->peer_recv[buf], true, 0, true);
    if (rc == ST_OK) { *exchanged = true; return ST_OK; }
    if (rc != ST_UNSUPPORTED) return rc;
  }
  JTB_TRY(e.c2c_lines(b, geo_make(Rh * Cn, 1, S * Rh * Cn, Rh * Cn), Rh * Cn, S, true, false, (T)1));
  if (m->exchange == 0) {
    JTB_TRY(slab_rows_to_peers<T>(m, e, b, buf, true));
    *exchanged = true;
  }
  return ST_OK;
}
template <typename T> int slab_back_b(jtb_slab* m, cx<T>* loc, bool scale, cudaStream_t st) {
  Engine<T> e(m->ctx, st);
  const i64 Ls = m->Ls, R = m->R, Cn = m->Cn;
  JTB_TRY(e.c2c_lines(loc, geo_make(Cn, 1, R * Cn, Cn), Cn * Ls, R, true, false, (T)1));
  return e.c2c_lines(loc, geo_contig(Cn), Ls * R, Cn, true, scale, (T)(1.0 / ((double)m->S * (double)R * (double)Cn)));
}

// ---- pipelined forward step: column blocks of the exchange overlap the slice-axis pass (see jtb_slab::nchunks)
// number of column blocks this member's shape supports (1: not pipelined)
template <typename T> int slab_pipe_blocks(const jtb_slab* m) {
  if (m->P < 2 || m->exchange != 0 || m->nchunks < 2 || !is_pow2(m->S)) return 1;
  const int w = fast_scatter_width<T>(m->R, m->Cn);
  if (w <= 0) return 1;
  int nb = m->nchunks > 16 ? 16 : m->nchunks;
  for (; nb > 1; --nb) {
    if (m->Cn % nb) continue;
    const i64 cc = m->Cn / nb;
    if (cc % w == 0 && fast_has_strided<T>(ilog2(m->S), cc)) break;
  }
  return nb;
}
int slab_pipe_ensure(jtb_slab* m, int nb) {
  if (!m->st2) {
    int lo = 0, hi = 0;
    JTB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    static const char* ep = getenv("JTB_PIPE_PRIO");   // 0: consumer stream at normal priority
    JTB_CUDA(cudaStreamCreateWithPriority(&m->st2, cudaStreamNonBlocking, (ep && atoi(ep) == 0) ? lo : hi));
    JTB_CUDA(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
  }
  for (int j = 0; j < nb; ++j)
    if (!m->evc[j]) JTB_CUDA(cudaEventCreateWithFlags(&m->evc[j], cudaEventDisableTiming));
  return ST_OK;
}
// producer side on `st`: rows in place, then the column (k2) pass block by block, its stores going to the owners;
// every block is announced (event for same-process groups, flag write for IPC peers)
template <typename T> int slab_pipe_produce(jtb_slab* m, cx<T>* a, bool inverse, cudaStream_t st, int buf, int nb) {
  Engine<T> e(m->ctx, st);
  const i64 Ls = m->Ls, R = m->R, Cn = m->Cn, cc = Cn / nb;
  JTB_TRY(slab_pipe_ensure(m, nb));
  JTB_TRY(e.c2c_lines(a, geo_contig(Cn), Ls * R, Cn, inverse, false, (T)1));
  for (int j = 0; j < nb; ++j) {
    JTB_TRY(fast_scatter<T>(e, a, Ls, R, Cn, m->P, m->rank, m->peer_recv[buf], inverse, -1, false, (i64)j * cc, cc));
    if (m->mode == 1) {
      m->chunk_epoch[j] = ++m->epoch;
      JTB_TRY(peer_barrier(m->ctx, st, (void* const*)m->peer_flags, m->P, m->rank, m->epoch, 1));
    } else {
      JTB_CUDA(cudaEventRecord(m->evc[j], st));
    }
    JTB_TRY(m->mark_chunk(m->pev_s, j, st));
  }
  m->last_nb = nb;
  return ST_OK;
}
// consumer side on `st2`: block j of the re-slabbed array [S][Rh][C] once every peer has announced it; `st` joins at the end
template <typename T> int slab_pipe_consume(jtb_slab* m, bool inverse, bool scale, cudaStream_t st, int buf, int nb) {
  Engine<T> e2(m->ctx, m->st2);
  {
    // the slice-axis pass shares the SMs with the exchange kernel of the next block: a bounded number of persistent CTAs
    // (JTB_PIPE_K1_CTAS, 0 = unbounded) so that it takes a fixed share of the slots instead of all of them
    static const char* ek = getenv("JTB_PIPE_K1_CTAS");
    e2.cta_limit = ek ? atoi(ek) : 148;
  }
  const i64 S = m->S, Rh = m->Rh, Cn = m->Cn, cc = Cn / nb;
  const T sc = (T)(1.0 / ((double)S * (double)m->R * (double)Cn));
  for (int j = 0; j < nb; ++j) {
    if (m->mode == 1) {
      JTB_TRY(peer_barrier(m->ctx, m->st2, (void* const*)m->peer_flags, m->P, m->rank, m->chunk_epoch[j], 2));
    } else {
      for (int h = 0; h < m->P; ++h) JTB_CUDA(cudaStreamWaitEvent(m->st2, m->group[h]->evc[j], 0));
    }
    cx<T>* b = (cx<T>*)m->recv[buf] + (i64)j * cc;
    JTB_TRY(e2.c2c_lines(b, geo_make(cc, 1, Cn, Rh * Cn), Rh * cc, S, inverse, scale, sc));
    JTB_TRY(m->mark_chunk(m->pev_k, j, m->st2));
  }
  JTB_CUDA(cudaEventRecord(m->ev_join, m->st2));
  JTB_CUDA(cudaStreamWaitEvent(st, m->ev_join, 0));
  return ST_OK;
}

// ---- exchange + slice-axis pass in one persistent kernel (fft_pipe_kernel): rows, then ONE launch
template <typename T> bool slab_fused_ok(const jtb_slab* m) {
  return m->P > 1 && m->exchange == 0 && m->fused_blocks > 0 && fast_pipe_has<T>(m->R, m->S, m->Cn, m->P, m->fused_blocks);
}
template <typename T> int slab_fused_step(jtb_slab* m, cx<T>* a, bool inverse, bool scale, cudaStream_t st, int buf) {
  Engine<T> e(m->ctx, st);
  const i64 Ls = m->Ls, R = m->R, Cn = m->Cn;
  JTB_TRY(e.c2c_lines(a, geo_contig(Cn), Ls * R, Cn, inverse, false, (T)1));
  JTB_CUDA(cudaMemsetAsync(m->pipe_counters, 0, 32 * sizeof(int), st));
  const long long epoch = ++m->epoch;
  bool handled = false;
  JTB_TRY(fast_pipe_exchange<T>(e, a, Ls, R, Cn, m->P, m->rank, m->peer_recv[buf], (void* const*)m->peer_flags, epoch,
                                m->pipe_counters, m->fused_blocks, inverse, scale,
                                (T)(1.0 / ((double)m->S * (double)R * (double)Cn)), &handled));
  if (!handled) { set_error("internal: fused exchange kernel unavailable"); return ST_UNSUPPORTED; }
  return ST_OK;
}

int slab_check_buffers(jtb_slab* m) {
  if (m->P > 1 && (m->mode == 0 || !m->recv[0])) { set_error("slab member is not connected to its peers"); return ST_ARG; }
  return ST_OK;
}

// single-process P == 1: all three passes in place
template <typename T> int slab_local_only(jtb_slab* m, cx<T>* a, bool inverse, bool scale, cudaStream_t st) {
  bool ex;
  JTB_TRY(slab_phase_a<T>(m, a, inverse, st, 0, &ex));
  return slab_phase_b<T>(m, a, inverse, scale, st);
}

}  // namespace

// ------------------------------------------------------------------------------------------------ internal API
namespace jtb {

// One step of a same-process group: `a[g]` = member g's [Ls][R][C] slab (device memory on its GPU), transformed on
// `st[g]`; results[g] = the k2-slabbed block [S][Rh][C] (member-owned receive buffer, valid until the step after
// next).  back = the distributed inverse of k2-slabbed blocks (results are the [Ls][R][C] slabs).
int slab_group_run(jtb_slab* const* ms, int n, void* const* a, bool back, bool inverse, bool scale, void** results,
                   cudaStream_t const* st) {
  if (n < 1 || !ms || !ms[0] || n != ms[0]->P) { set_error("group size does not match the decomposition"); return ST_ARG; }
  const bool f64 = ms[0]->prec == JTB_F64;
  if (n == 1) {
    jtb_slab* m = ms[0];
    DeviceGuard dg(m->device);
    JTB_TRY(m->ctx->order_begin(st[0]));
    if (back) {
      bool ex;
      JTB_TRY(f64 ? slab_back_a<double>(m, (double2*)a[0], st[0], 0, &ex) : slab_back_a<float>(m, (float2*)a[0], st[0], 0, &ex));
      JTB_TRY(f64 ? slab_back_b<double>(m, (double2*)a[0], scale, st[0]) : slab_back_b<float>(m, (float2*)a[0], scale, st[0]));
    } else {
      JTB_TRY(f64 ? slab_local_only<double>(m, (double2*)a[0], inverse, scale, st[0])
                  : slab_local_only<float>(m, (float2*)a[0], inverse, scale, st[0]));
    }
    JTB_TRY(m->ctx->order_end(st[0]));
    results[0] = a[0];
    return ST_OK;
  }
  for (int g = 0; g < n; ++g) {
    if (!ms[g] || ms[g]->mode != 2 || ms[g]->rank != g) { set_error("not a connected same-process group"); return ST_ARG; }
    JTB_TRY(slab_check_buffers(ms[g]));
  }
  const int buf = ms[0]->step & 1;
  {
    bool fused = !back && (f64 ? slab_fused_ok<double>(ms[0]) : slab_fused_ok<float>(ms[0]));
    for (int g = 0; g < n && fused; ++g)          // persistent spinning CTAs: every member needs a GPU of its own
      for (int h = 0; h < g; ++h)
        if (ms[g]->device == ms[h]->device) fused = false;
    if (fused) {
      for (int g = 0; g < n; ++g) {
        jtb_slab* m = ms[g];
        DeviceGuard dg(m->device);
        JTB_TRY(m->ctx->order_begin(st[g]));
        JTB_TRY(m->mark(0, st[g]));
        JTB_TRY(f64 ? slab_fused_step<double>(m, (double2*)a[g], inverse, scale, st[g], buf)
                    : slab_fused_step<float>(m, (float2*)a[g], inverse, scale, st[g], buf));
        JTB_TRY(m->mark(1, st[g]));
        JTB_TRY(m->mark(2, st[g]));
        JTB_TRY(m->mark(3, st[g]));
        JTB_TRY(m->ctx->order_end(st[g]));
        results[g] = m->recv[buf];
        m->step++;
      }
      return ST_OK;
    }
  }
  const int nb = back ? 1 : (f64 ? slab_pipe_blocks<double>(ms[0]) : slab_pipe_blocks<float>(ms[0]));
  if (nb > 1) {
    for (int g = 0; g < n; ++g) {
      jtb_slab* m = ms[g];
      DeviceGuard dg(m->device);
      JTB_TRY(m->ctx->order_begin(st[g]));
      JTB_TRY(m->mark(0, st[g]));
      JTB_TRY(f64 ? slab_pipe_produce<double>(m, (double2*)a[g], inverse, st[g], buf, nb)
                  : slab_pipe_produce<float>(m, (float2*)a[g], inverse, st[g], buf, nb));
      JTB_TRY(m->mark(1, st[g]));
      JTB_TRY(m->mark(2, st[g]));
    }
    for (int g = 0; g < n; ++g) {
      jtb_slab* m = ms[g];
      DeviceGuard dg(m->device);
      JTB_TRY(f64 ? slab_pipe_consume<double>(m, inverse, scale, st[g], buf, nb)
                  : slab_pipe_consume<float>(m, inverse, scale, st[g], buf, nb));
      JTB_TRY(m->mark(3, st[g]));
      JTB_TRY(m->ctx->order_end(st[g]));
      results[g] = m->recv[buf];
      m->step++;
    }
    return ST_OK;
  }
  bool exchanged[8] = {false};
  for (int g = 0; g < n; ++g) {
    jtb_slab* m = ms[g];
    DeviceGuard dg(m->device);
    JTB_TRY(m->ctx->order_begin(st[g]));
    JTB_TRY(m->mark(0, st[g]));
    if (back)
      JTB_TRY(f64 ? slab_back_a<double>(m, (double2*)a[g], st[g], buf, &exchanged[g])
                  : slab_back_a<float>(m, (float2*)a[g], st[g], buf, &exchanged[g]));
    else
      JTB_TRY(f64 ? slab_phase_a<double>(m, (double2*)a[g], inverse, st[g], buf, &exchanged[g])
                  : slab_phase_a<float>(m, (float2*)a[g], inverse, st[g], buf, &exchanged[g]));
    JTB_TRY(m->mark(1, st[g]));
    JTB_TRY(m->ctx->order_end(st[g]));
  }
  if (!exchanged[0]) {
#ifndef JTB_EMU
    NcclApi* api = nccl_api();
    if (!api) { set_error("libnccl.so.2 could not be loaded"); return ST_NCCL; }
    JTB_NCCL(api->GroupStart());
    int rc = ST_OK;
    for (int g = 0; g < n && rc == ST_OK; ++g) {
      jtb_slab* m = ms[g];
      DeviceGuard dg(m->device);
      rc = f64 ? slab_nccl_enqueue<double>(m, (const double2*)a[g], (double2*)m->recv[buf], back, st[g])
               : slab_nccl_enqueue<float>(m, (const float2*)a[g], (float2*)m->recv[buf], back, st[g]);
    }
    const ncclResult_t ge = api->GroupEnd();
    if (rc != ST_OK) return rc;
    JTB_NCCL(ge);
#else
    set_error("NCCL is not available in the emulated build");
    return ST_NCCL;
#endif
  } else {
    // every member's exchange stores are complete before any member starts the next pass
    for (int g = 0; g < n; ++g) {
      DeviceGuard dg(ms[g]->device);
      JTB_CUDA(cudaEventRecord(ms[g]->ev, st[g]));
    }
    for (int g = 0; g < n; ++g) {
      DeviceGuard dg(ms[g]->device);
      for (int h = 0; h < n; ++h)
        if (h != g) JTB_CUDA(cudaStreamWaitEvent(st[g], ms[h]->ev, 0));
    }
  }
  for (int g = 0; g < n; ++g) {
    jtb_slab* m = ms[g];
    DeviceGuard dg(m->device);
    JTB_TRY(m->ctx->order_begin(st[g]));
    JTB_TRY(m->mark(2, st[g]));
    if (back) JTB_TRY(f64 ? slab_back_b<double>(m, (double2*)m->recv[buf], scale, st[g]) : slab_back_b<float>(m, (float2*)m->recv[buf], scale, st[g]));
    else JTB_TRY(f64 ? slab_phase_b<double>(m, (double2*)m->recv[buf], inverse, scale, st[g])
                     : slab_phase_b<float>(m, (float2*)m->recv[buf], inverse, scale, st[g]));
    JTB_TRY(m->mark(3, st[g]));
    JTB_TRY(m->ctx->order_end(st[g]));
    results[g] = m->recv[buf];
    m->step++;
  }
  return ST_OK;
}

}  // namespace jtb

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

int jtb_slab_create(jtb_slab** out, int prec, int64_t S, int64_t R, int64_t Cn, int nranks, int rank, int device) {
  if (!out) { set_error("null argument"); return ST_ARG; }
  *out = nullptr;
  if ((prec != JTB_F64 && prec != JTB_F32) || S < 2 || R < 2 || Cn < 2 || nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks) {
    set_error("bad slab arguments (1..8 ranks; slices, rows and columns must be greater than 1)");
    return ST_ARG;
  }
  if (S % nranks || R % nranks) { set_error("slices and rows must be divisible by the number of ranks"); return ST_ARG; }
  Ctx* ctx = get_ctx(device);
  if (!ctx) return ST_CUDA;
  DeviceGuard dg(device);
  jtb_slab* m = new jtb_slab();
  m->prec = prec; m->P = nranks; m->rank = rank; m->device = device; m->ctx = ctx;
  m->S = S; m->R = R; m->Cn = Cn; m->Ls = S / nranks; m->Rh = R / nranks;
  m->csz = prec == JTB_F64 ? 16 : 8;
  m->block_bytes = (size_t)(S * m->Rh * Cn) * m->csz;
  {
    static const char* ex = getenv("JTB_EXCHANGE_NCCL");
    m->exchange = (ex && atoi(ex)) ? 1 : 0;
    static const char* fu = getenv("JTB_SLAB_FUSED");
    m->fused_blocks = fu ? atoi(fu) : 0;
    static const char* ch = getenv("JTB_SLAB_CHUNKS");
    // Measured on 2 x B200, 512^3 (profiles/r02_pipe_timeline_2gpu.log): the blocks do overlap, but the peer-store
    // exchange needs every SM's store slots -- a block's stores take 0.20 ms alone and 0.31 ms next to the slice-axis
    // pass of the previous block (0.10 ms alone), so the step is 1.56 ms against 1.28 ms for the fused in-slice kernel +
    // one barrier.  Off by default; JTB_SLAB_CHUNKS=2..16 switches the pipelined variant on.
    m->nchunks = ch ? atoi(ch) : 1;
    if (m->nchunks < 1) m->nchunks = 1;
  }
  if (nranks > 1) {
    cudaError_t e = cudaMalloc(&m->recv[0], m->block_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&m->recv[1], m->block_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->flags, 256 * sizeof(long long));   // 8 barrier slots + 8 x 16 block slots
    if (e == cudaSuccess) e = cudaMemset(m->flags, 0, 256 * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->pipe_counters, 32 * sizeof(int));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev, cudaEventDisableTiming);
    if (e != cudaSuccess) {
      const int rc = cuda_fail(e, "slab buffers");
      jtb_slab_destroy(m);
      return rc;
    }
    for (int b = 0; b < 2; ++b) m->peer_recv[b][rank] = m->recv[b];
    m->peer_flags[rank] = m->flags;
  }
  *out = m;
  return ST_OK;
}

int jtb_slab_destroy(jtb_slab* m) {
  if (!m) return ST_OK;
  DeviceGuard dg(m->device);
  cudaDeviceSynchronize();
#ifndef JTB_EMU
  for (int h = 0; h < m->P; ++h)
    if (m->ipc_open[h]) {
      for (int b = 0; b < 2; ++b) if (m->peer_recv[b][h]) cudaIpcCloseMemHandle(m->peer_recv[b][h]);
      if (m->peer_flags[h]) cudaIpcCloseMemHandle(m->peer_flags[h]);
    }
  if (m->comm && nccl_api()) nccl_api()->CommDestroy((ncclComm_t)m->comm);
#endif
  for (int b = 0; b < 2; ++b) if (m->recv[b]) cudaFree(m->recv[b]);
  if (m->flags) cudaFree(m->flags);
  if (m->pipe_counters) cudaFree(m->pipe_counters);
  if (m->ev) cudaEventDestroy(m->ev);
  if (m->st2) cudaStreamDestroy(m->st2);
  if (m->ev_join) cudaEventDestroy(m->ev_join);
  for (int j = 0; j < 16; ++j) {
    if (m->evc[j]) cudaEventDestroy(m->evc[j]);
    if (m->pev_s[j]) cudaEventDestroy(m->pev_s[j]);
    if (m->pev_k[j]) cudaEventDestroy(m->pev_k[j]);
  }
  for (int i = 0; i < 4; ++i) if (m->pev[i]) cudaEventDestroy(m->pev[i]);
  cudaGetLastError();
  delete m;
  return ST_OK;
}

int jtb_slab_export(jtb_slab* m, unsigned char* handles192) {
  if (!m || !handles192) { set_error("null argument"); return ST_ARG; }
  memset(handles192, 0, 192);
  if (m->P == 1) return ST_OK;
  DeviceGuard dg(m->device);
  void* ptrs[3] = {m->recv[0], m->recv[1], (void*)m->flags};
  for (int i = 0; i < 3; ++i) {
#ifdef JTB_EMU
    memcpy(handles192 + 64 * i, &ptrs[i], sizeof(void*));
#else
    cudaIpcMemHandle_t h;
    JTB_CUDA(cudaIpcGetMemHandle(&h, ptrs[i]));
    memcpy(handles192 + 64 * i, &h, 64);
#endif
  }
  return ST_OK;
}

int jtb_slab_connect_ipc(jtb_slab* m, const unsigned char* all_handles) {
  if (!m || !all_handles) { set_error("null argument"); return ST_ARG; }
  if (m->P == 1) { m->mode = 1; return ST_OK; }
  DeviceGuard dg(m->device);
  for (int h = 0; h < m->P; ++h) {
    if (h == m->rank) continue;
    void* ptrs[3] = {nullptr, nullptr, nullptr};
    for (int i = 0; i < 3; ++i) {
#ifdef JTB_EMU
      memcpy(&ptrs[i], all_handles + (size_t)h * 192 + 64 * i, sizeof(void*));
#else
      cudaIpcMemHandle_t ih;
      memcpy(&ih, all_handles + (size_t)h * 192 + 64 * i, 64);
      JTB_CUDA(cudaIpcOpenMemHandle(&ptrs[i], ih, cudaIpcMemLazyEnablePeerAccess));
#endif
    }
    m->peer_recv[0][h] = ptrs[0]; m->peer_recv[1][h] = ptrs[1]; m->peer_flags[h] = (long long*)ptrs[2];
    m->ipc_open[h] = 1;
  }
  m->mode = 1;
  return ST_OK;
}

int jtb_slab_connect_local(jtb_slab* const* ms, int n) {
  if (!ms || n < 1 || !ms[0] || n != ms[0]->P) { set_error("group size does not match the decomposition"); return ST_ARG; }
  for (int g = 0; g < n; ++g)
    if (!ms[g] || ms[g]->rank != g || ms[g]->P != n) { set_error("members must be passed in rank order"); return ST_ARG; }
#ifndef JTB_EMU
  for (int g = 0; g < n; ++g)
    for (int h = 0; h < n; ++h) {
      if (ms[g]->device == ms[h]->device) continue;
      DeviceGuard dg(ms[g]->device);
      int can = 0;
      JTB_CUDA(cudaDeviceCanAccessPeer(&can, ms[g]->device, ms[h]->device));
      if (!can) { set_error("device %d cannot access device %d (no peer path)", ms[g]->device, ms[h]->device); return ST_UNSUPPORTED; }
      const cudaError_t e = cudaDeviceEnablePeerAccess(ms[h]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
      cudaGetLastError();
    }
#endif
  for (int g = 0; g < n; ++g) {
    for (int h = 0; h < n; ++h) {
      ms[g]->peer_recv[0][h] = ms[h]->recv[0];
      ms[g]->peer_recv[1][h] = ms[h]->recv[1];
      ms[g]->peer_flags[h] = ms[h]->flags;
      ms[g]->group[h] = ms[h];
    }
    ms[g]->mode = 2;
  }
  return ST_OK;
}

int jtb_nccl_unique_id(unsigned char* id128) {
  if (!id128) { set_error("null argument"); return ST_ARG; }
#ifdef JTB_EMU
  set_error("NCCL is not available in the emulated build");
  return ST_NCCL;
#else
  NcclApi* api = nccl_api();
  if (!api) { set_error("libnccl.so.2 could not be loaded"); return ST_NCCL; }
  ncclUniqueId id;
  static_assert(sizeof(id) == 128, "ncclUniqueId size");
  JTB_NCCL(api->GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return ST_OK;
#endif
}

int jtb_slab_nccl_init(jtb_slab* m, const unsigned char* id128) {
  if (!m || !id128) { set_error("null argument"); return ST_ARG; }
#ifdef JTB_EMU
  set_error("NCCL is not available in the emulated build");
  return ST_NCCL;
#else
  NcclApi* api = nccl_api();
  if (!api) { set_error("libnccl.so.2 could not be loaded"); return ST_NCCL; }
  DeviceGuard dg(m->device);
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm;
  JTB_NCCL(api->CommInitRank(&comm, m->P, id, m->rank));
  m->comm = comm;
  return ST_OK;
#endif
}

int jtb_slab_nccl_init_local(jtb_slab* const* ms, int n) {
  if (!ms || n < 1) { set_error("bad argument"); return ST_ARG; }
#ifdef JTB_EMU
  set_error("NCCL is not available in the emulated build");
  return ST_NCCL;
#else
  NcclApi* api = nccl_api();
  if (!api) { set_error("libnccl.so.2 could not be loaded"); return ST_NCCL; }
  int devs[8];
  ncclComm_t comms[8];
  for (int g = 0; g < n; ++g) devs[g] = ms[g]->device;
  for (int g = 0; g < n; ++g)
    for (int h = 0; h < g; ++h)
      if (devs[g] == devs[h]) { set_error("NCCL needs distinct devices"); return ST_UNSUPPORTED; }
  JTB_NCCL(api->CommInitAll(comms, n, devs));
  for (int g = 0; g < n; ++g) ms[g]->comm = comms[g];
  return ST_OK;
#endif
}

int jtb_slab_set_exchange(jtb_slab* m, int use_nccl) {
  if (!m) { set_error("null argument"); return ST_ARG; }
  if (use_nccl && !m->comm && m->P > 1) { set_error("no NCCL communicator: call jtb_slab_nccl_init first"); return ST_NCCL; }
  m->exchange = use_nccl ? 1 : 0;
  return ST_OK;
}

int64_t jtb_slab_block_elements(const jtb_slab* m) { return m ? 2 * m->S * m->Rh * m->Cn : 0; }

// one member of a multi-PROCESS decomposition (peers connected through IPC): the exchange is ordered by the
// device-side flag barrier, or by NCCL's own stream semantics
static int slab_member_run(jtb_slab* m, void* a, bool back, bool inverse, bool scale, void** result, cudaStream_t st) {
  if (!m || !a || !result) { set_error("null argument"); return ST_ARG; }
  const bool f64 = m->prec == JTB_F64;
  DeviceGuard dg(m->device);
  std::lock_guard<std::mutex> lk(m->ctx->mu);
  if (m->P == 1) {
    jtb_slab* ms[1] = {m};
    void* aa[1] = {a};
    return slab_group_run(ms, 1, aa, back, inverse, scale, result, &st);
  }
  if (m->mode != 1) { set_error("member of a same-process group: use the group call"); return ST_ARG; }
  JTB_TRY(slab_check_buffers(m));
  JTB_TRY(m->ctx->order_begin(st));
  const int buf = m->step & 1;
  if (!back && (f64 ? slab_fused_ok<double>(m) : slab_fused_ok<float>(m))) {
    JTB_TRY(m->mark(0, st));
    JTB_TRY(f64 ? slab_fused_step<double>(m, (double2*)a, inverse, scale, st, buf) : slab_fused_step<float>(m, (float2*)a, inverse, scale, st, buf));
    JTB_TRY(m->mark(1, st));
    JTB_TRY(m->mark(2, st));
    JTB_TRY(m->mark(3, st));
    m->step++;
    JTB_TRY(m->ctx->order_end(st));
    *result = m->recv[buf];
    return ST_OK;
  }
  {
    const int nb = back ? 1 : (f64 ? slab_pipe_blocks<double>(m) : slab_pipe_blocks<float>(m));
    if (nb > 1) {
      JTB_TRY(m->mark(0, st));
      JTB_TRY(f64 ? slab_pipe_produce<double>(m, (double2*)a, inverse, st, buf, nb)
                  : slab_pipe_produce<float>(m, (float2*)a, inverse, st, buf, nb));
      JTB_TRY(m->mark(1, st));
      JTB_TRY(m->mark(2, st));
      m->step++;
      JTB_TRY(f64 ? slab_pipe_consume<double>(m, inverse, scale, st, buf, nb) : slab_pipe_consume<float>(m, inverse, scale, st, buf, nb));
      JTB_TRY(m->mark(3, st));
      JTB_TRY(m->ctx->order_end(st));
      *result = m->recv[buf];
      return ST_OK;
    }
  }
  bool exchanged = false;
  JTB_TRY(m->mark(0, st));
  if (back) JTB_TRY(f64 ? slab_back_a<double>(m, (double2*)a, st, buf, &exchanged) : slab_back_a<float>(m, (float2*)a, st, buf, &exchanged));
  else JTB_TRY(f64 ? slab_phase_a<double>(m, (double2*)a, inverse, st, buf, &exchanged)
                   : slab_phase_a<float>(m, (float2*)a, inverse, st, buf, &exchanged));
  JTB_TRY(m->mark(1, st));
  m->step++;
  if (exchanged) {
    m->epoch++;
    JTB_TRY(peer_barrier(m->ctx, st, (void* const*)m->peer_flags, m->P, m->rank, m->epoch));
  } else {
#ifndef JTB_EMU
    NcclApi* api = nccl_api();
    if (!api) { set_error("libnccl.so.2 could not be loaded"); return ST_NCCL; }
    JTB_NCCL(api->GroupStart());
    const int rc = f64 ? slab_nccl_enqueue<double>(m, (const double2*)a, (double2*)m->recv[buf], back, st)
                       : slab_nccl_enqueue<float>(m, (const float2*)a, (float2*)m->recv[buf], back, st);
    const ncclResult_t ge = api->GroupEnd();
    if (rc != ST_OK) return rc;
    JTB_NCCL(ge);
#else
    set_error("NCCL is not available in the emulated build");
    return ST_NCCL;
#endif
  }
  JTB_TRY(m->mark(2, st));
  if (back) JTB_TRY(f64 ? slab_back_b<double>(m, (double2*)m->recv[buf], scale, st) : slab_back_b<float>(m, (float2*)m->recv[buf], scale, st));
  else JTB_TRY(f64 ? slab_phase_b<double>(m, (double2*)m->recv[buf], inverse, scale, st)
                   : slab_phase_b<float>(m, (float2*)m->recv[buf], inverse, scale, st));
  JTB_TRY(m->mark(3, st));
  JTB_TRY(m->ctx->order_end(st));
  *result = m->recv[buf];
  return ST_OK;
}

int jtb_slab_forward(jtb_slab* m, void* dev_a, int inverse, int scale, void** result, void* stream) {
  return slab_member_run(m, dev_a, false, inverse != 0, scale != 0, result, (cudaStream_t)stream);
}
int jtb_slab_back(jtb_slab* m, void* dev_b, int scale, void** result, void* stream) {
  return slab_member_run(m, dev_b, true, true, scale != 0, result, (cudaStream_t)stream);
}

int jtb_slab_group_forward(jtb_slab* const* ms, int n, void* const* dev_a, int inverse, int scale, void** results,
                           void* const* streams) {
  if (!ms || !dev_a || !results || n < 1 || n > 8) { set_error("bad argument"); return ST_ARG; }
  cudaStream_t st[8];
  for (int g = 0; g < n; ++g) st[g] = streams ? (cudaStream_t)streams[g] : nullptr;
  return slab_group_run(ms, n, dev_a, false, inverse != 0, scale != 0, results, st);
}
int jtb_slab_group_back(jtb_slab* const* ms, int n, void* const* dev_b, int scale, void** results, void* const* streams) {
  if (!ms || !dev_b || !results || n < 1 || n > 8) { set_error("bad argument"); return ST_ARG; }
  cudaStream_t st[8];
  for (int g = 0; g < n; ++g) st[g] = streams ? (cudaStream_t)streams[g] : nullptr;
  return slab_group_run(ms, n, dev_b, true, true, scale != 0, results, st);
}

// phase timing of the last step (enable first): ms[0] = in-slice passes incl. the exchange stores, ms[1] = waiting for
// the peers (barrier / NCCL), ms[2] = slice-axis pass.  Synchronises with the step's stream.
int jtb_slab_profile(jtb_slab* m, int enable) {
  if (!m) { set_error("null argument"); return ST_ARG; }
  m->profile = enable != 0;
  return ST_OK;
}
int jtb_slab_last_times(jtb_slab* m, float* ms3) {
  if (!m || !ms3) { set_error("null argument"); return ST_ARG; }
  if (!m->profile || !m->pev[3]) { set_error("no profiled step (jtb_slab_profile)"); return ST_ARG; }
  DeviceGuard dg(m->device);
  JTB_CUDA(cudaEventSynchronize(m->pev[3]));
  for (int i = 0; i < 3; ++i) JTB_CUDA(cudaEventElapsedTime(&ms3[i], m->pev[i], m->pev[i + 1]));
  return ST_OK;
}

// timeline of the last profiled PIPELINED step, ms since its start: stored[j] = column block j has left (producer
// stream), done[j] = slice-axis pass of block j finished (consumer stream).  Returns the number of blocks (0: the last
// step was not pipelined) in *nblocks; arrays hold up to 16 entries.
int jtb_slab_chunk_times(jtb_slab* m, int* nblocks, float* stored, float* done) {
  if (!m || !nblocks || !stored || !done) { set_error("null argument"); return ST_ARG; }
  *nblocks = 0;
  if (!m->profile || !m->pev[0] || m->last_nb < 2) return ST_OK;
  DeviceGuard dg(m->device);
  JTB_CUDA(cudaEventSynchronize(m->pev[3]));
  for (int j = 0; j < m->last_nb; ++j) {
    if (!m->pev_s[j] || !m->pev_k[j]) return ST_OK;
    JTB_CUDA(cudaEventElapsedTime(&stored[j], m->pev[0], m->pev_s[j]));
    JTB_CUDA(cudaEventElapsedTime(&done[j], m->pev[0], m->pev_k[j]));
  }
  *nblocks = m->last_nb;
  return ST_OK;
}

// reads (and clears) the watchdog word of the member's device: JTB_ERR_CUDA when a spin-waiting kernel timed out
int jtb_slab_status(jtb_slab* m) {
  if (!m) { set_error("null argument"); return ST_ARG; }
  return m->ctx->check_watchdog("slab exchange");
}

}  // extern "C"
