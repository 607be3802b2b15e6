/* libjtb200 -- C ABI of the B200-native JTransforms transform path.
 *
 * JTransforms (wendykierp/JTransforms) has no FFI of its own: the boundary of the hot path is its public
 * Java API.  Each entry point below is what a Java shim (Panama FFM downcall or JNI) binds so that the
 * org.jtransforms classes keep their signatures while the arithmetic runs on the GPU.  Citations are
 * relative to src/main/java/org/jtransforms in the reference repository.
 *
 *   jtb_plan_create   <- constructors DoubleFFT_1D(long) fft/DoubleFFT_1D.java:117-188, DoubleFFT_2D(long,long)
 *                        fft/DoubleFFT_2D.java:75-98, DoubleFFT_3D(long,long,long) fft/DoubleFFT_3D.java:88-124,
 *                        DoubleDCT_1D dct/DoubleDCT_1D.java:88-138, DoubleDST_1D dst/DoubleDST_1D.java:59-65,
 *                        DoubleDHT_1D dht/DoubleDHT_1D.java:60-66 (+ 2-D/3-D and Float twins)
 *   jtb_exec          <- complexForward fft/DoubleFFT_1D.java:243-263, complexInverse :362-385, realForward
 *                        :524-561, realForwardFull :678-755, realInverse :946-989, realInverseFull :1112-1195;
 *                        2-D fft/DoubleFFT_2D.java:115,456,820,956,1077,1212; 3-D fft/DoubleFFT_3D.java:145,
 *                        :744,:1339,:1480,:1629,:1770; forward/inverse of dct/DoubleDCT_{1,2,3}D.java,
 *                        dst/DoubleDST_{1,2,3}D.java, dht/DoubleDHT_{1,2,3}D.java
 *   jtb_exec_batch    <- the caller loop over `offa` the reference needs for batches (no batched API there)
 *   jtb_exec_device   <- same transforms on device-resident data (benchmarks, multi-GPU composition)
 *   jtb_lines_c2c_device <- the strided line loops of the N-D drivers (fft/DoubleFFT_3D.java:5505-5713,
 *                        :6318-6520) exposed so a host layer can run slab-decomposed passes per GPU
 *   jtb_fft3d_k2_scatter / jtb_fft3d_k1_scatter / jtb_fft2d_slices_device <- the slice-axis gather of cdft3db_subth
 *                        (fft/DoubleFFT_3D.java:6318-6520) when the slices live on several GPUs: the re-slabbing
 *                        all-to-all (forward and back) fused into the pass's stores over NVLink peer mappings
 *   jtb_plan_set_devices / jtb_slab_* <- the thread-pool partitioning of one transform (ConcurrencyUtils.submit in
 *                        fft/DoubleFFT_3D.java:5523-5707, :6318-6520) mapped onto several GPUs behind the same call
 *   jtb_host_alloc / jtb_host_register <- the role of JLargeArrays' off-heap storage at the boundary
 *                        (DoubleLargeArray, fft/DoubleFFT_1D.java:280-304): page-locked caller memory
 *
 * All transforms are in place on interleaved (re, im) or real arrays exactly as the reference lays them out.
 * Functions return 0 on success or a JTB_ERR_* code; jtb_last_error() returns the thread-local message.
 * There is no CPU fallback: every compute entry point fails with JTB_ERR_CUDA when no device is present.
 */
#ifndef JTB200_H
#define JTB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jtb_plan jtb_plan;

enum { JTB_OK = 0, JTB_ERR_ARG = 1, JTB_ERR_UNSUPPORTED = 2, JTB_ERR_CUDA = 3, JTB_ERR_OOM = 4, JTB_ERR_NCCL = 5 };
enum { JTB_FFT = 0, JTB_DCT = 1, JTB_DST = 2, JTB_DHT = 3 };
enum { JTB_F64 = 0, JTB_F32 = 1 };
enum {
  JTB_C2C_FORWARD = 0, /* complexForward */
  JTB_C2C_INVERSE = 1, /* complexInverse(scale) */
  JTB_R2C_PACKED = 2,  /* realForward (packed half spectrum) */
  JTB_R2C_FULL = 3,    /* realForwardFull */
  JTB_C2R_PACKED = 4,  /* realInverse(scale) */
  JTB_C2R_FULL = 5,    /* realInverseFull(scale) */
  JTB_R2R_FORWARD = 6, /* DCT/DST/DHT forward(scale) */
  JTB_R2R_INVERSE = 7  /* DCT/DST/DHT inverse(scale) */
};

/* rank 1..3, dims[0..rank) = (n) | (rows, columns) | (slices, rows, columns); device = CUDA ordinal */
int jtb_plan_create(jtb_plan** out, int kind, int prec, int rank, const int64_t* dims, int device);
int jtb_plan_destroy(jtb_plan* plan);
/* number of array elements (doubles/floats) the op reads+writes for one transform */
int64_t jtb_plan_elements(const jtb_plan* plan, int op);

/* Multi-GPU plans (single process; replaces the ConcurrencyUtils thread-pool partitioning, fft/DoubleFFT_3D.java:145-325).
 * After jtb_plan_set_devices(plan, P, devices) -- P in 1..8, every pair peer-accessible --
 *   - jtb_exec of a rank-3 FFT plan with complexForward / complexInverse slab-decomposes the ONE caller array over the
 *     P GPUs: slab g travels host -> device g over that GPU's own PCIe link, the k3/k2 passes run per slab, the
 *     re-slabbing all-to-all is fused into the k2 stores over NVLink peer memory (NCCL ncclSend/ncclRecv with
 *     JTB_EXCHANGE_NCCL=1), the k1 pass runs on the re-slabbed blocks and a pitched device -> host copy delivers the
 *     result in natural [slices][rows][columns] order.  Slices and rows must be divisible by P.
 *   - jtb_exec_batch splits the batch into P contiguous blocks, one per GPU, each with its own copy/compute pipeline
 *     (no collective: BASELINE config 3, batched 1-D/2-D work).
 *   - everything else runs on devices[0].
 * The same device may be listed several times (virtual ranks; used by the single-GPU tests). */
int jtb_plan_set_devices(jtb_plan* plan, int ndev, const int* devices);
int jtb_plan_device_count(const jtb_plan* plan);

/* host array in, host array out (same addresses): H2D, kernels, D2H, synchronised on return */
int jtb_exec(jtb_plan* plan, int op, void* host_a, int64_t offa, int scale);
/* same, with the length of the caller's array (elements): JTB_ERR_ARG "array too short" instead of reading or writing
 * past the end -- what the Java shim calls so that a short double[] raises ArrayIndexOutOfBoundsException
 * (the reference: unchecked loops, fft/DoubleFFT_1D.java:243-263) */
int jtb_exec_n(jtb_plan* plan, int op, void* host_a, int64_t a_length, int64_t offa, int scale);
/* `howmany` transforms, transform b starting at host_a[offa + b*dist].  Spans of at least three chunks (64 MiB of
 * transforms each; JTB_BATCH_MB overrides, 0 disables) run as a three-slot ring -- H2D of chunk i+1, the kernels of
 * chunk i and D2H of chunk i-1 overlap on two copy streams -- so the device copy is three chunks, not the span.
 * Pinned host memory (jtb_host_alloc) is what makes the copies asynchronous. */
int jtb_exec_batch(jtb_plan* plan, int op, void* host_a, int64_t offa, int64_t howmany, int64_t dist, int scale);
/* device-resident: dev_a is a device pointer (16-byte aligned); asynchronous on `stream` (cudaStream_t, may be 0) */
int jtb_exec_device(jtb_plan* plan, int op, void* dev_a, int64_t howmany, int64_t dist, int scale, void* stream);

/* in-place complex FFT of length n along `nlines` strided lines of a device array (units: complex elements):
 * line l = i0 + c0*i3 starts at i0*d0 + i3*d3, its element j at + j*stride.  scale multiplies the output. */
int jtb_lines_c2c_device(int prec, int device, void* dev_a, int64_t n, int64_t nlines, int64_t c0, int64_t d0,
                         int64_t d3, int64_t stride, int inverse, double scale, void* stream);

/* out-of-place variant for power-of-two strided lines with a lean kernel: line group i3 (c0 adjacent lines, d0 == 1)
 * is read at i3*d3 with element stride `stride` and stored at dev_out + i3*out_d3 with element stride out_stride --
 * the axis-swapping k2 / k1 passes of the 3-D transform (bench: per-pass roofline).  JTB_ERR_UNSUPPORTED otherwise. */
int jtb_lines_c2c_out_device(int prec, int device, void* dev_in, void* dev_out, int64_t n, int64_t nlines, int64_t c0,
                             int64_t d3, int64_t stride, int64_t out_d3, int64_t out_stride, int inverse, double scale,
                             void* stream);

/* Slab-decomposed 3-D transform over P GPUs (one process per GPU).  Rank g holds slices [g*Ls, (g+1)*Ls) as
 * [Ls][R][C].  jtb_fft3d_k2_scatter runs the row-axis (k2) pass of every local slice and stores each output
 * row straight into the receive buffer of the GPU that owns it after re-slabbing over k2 (peer h = k2/(R/P)
 * receives [g*Ls+ls][k2 % (R/P)][c] of its [S][R/P][C] block): the all-to-all of the transpose is fused into
 * the kernel's stores over NVLink peer mappings.  recv_ptrs[h] = peer-mapped pointer of rank h's buffer.
 * Replaces cdft3db_subth's slice-axis gather (fft/DoubleFFT_3D.java:6318-6520) across devices. */
int jtb_fft3d_k2_scatter(int prec, int device, const void* local_a, int64_t Ls, int64_t R, int64_t C, int nranks,
                         int rank, void* const* recv_ptrs, int inverse, void* stream);
/* same for a chunk of the local slab: local_a points at `Ls` slices whose first one has GLOBAL slice index
 * slice_base (lets the caller pipeline the k3 pass of one chunk under the exchange of the previous one) */
int jtb_fft3d_k2_scatter_chunk(int prec, int device, const void* local_a, int64_t Ls, int64_t slice_base, int64_t R,
                               int64_t C, int nranks, void* const* recv_ptrs, int inverse, void* stream);
/* The way back (distributed complexInverse): rank h holds the k2-slabbed block [S][Rh][C].  Runs the slice-axis (k1)
 * pass over it and stores output row k1 straight into the buffer of the GPU that owns slice k1 after re-slabbing over
 * k1 (peer g = k1/(S/P) receives [k1 % (S/P)][h*Rh + r][c] of its [S/P][R][C] slab): the second all-to-all of a
 * round trip, fused into the kernel's stores like the first. */
int jtb_fft3d_k1_scatter(int prec, int device, const void* local_b, int64_t S, int64_t Rh, int64_t C, int nranks,
                         int rank, void* const* recv_ptrs, int inverse, void* stream);
/* Both in-slice passes (rows, then columns) of `nslices` rows x cols slices in place; recv_ptrs == NULL keeps the
 * result local, otherwise the column pass stores straight into the peers' receive buffers as
 * jtb_fft3d_k2_scatter does.  512 x 512 double slices run as ONE persistent cooperative kernel that keeps the
 * intermediate in L2 (xdft3da_subth2, fft/DoubleFFT_3D.java:5505-5713). */
int jtb_fft2d_slices_device(int prec, int device, void* dev_a, int64_t nslices, int64_t rows, int64_t cols, int nranks,
                            int rank, void* const* recv_ptrs, int inverse, void* stream);
/* Slab decomposition objects: one jtb_slab = one rank (GPU) of a P-way slab-decomposed 3-D complex transform.  Rank g
 * owns slices [g*S/P, (g+1)*S/P) as [S/P][R][C]; jtb_slab_forward leaves the result k2-slabbed ([S][R/P][C] in a
 * member-owned receive buffer, valid until the step after next); jtb_slab_back is the distributed complexInverse of
 * such a block (result: the [S/P][R][C] slab).  Members connect inside one process (jtb_slab_connect_local, peer
 * access + events; drive them with the group calls) or across processes (jtb_slab_export / jtb_slab_connect_ipc: 192
 * bytes of CUDA IPC handles per rank, gathered by the caller; a device-side flag barrier orders the exchange).
 * jtb_slab_set_exchange(m, 1) switches the exchange from peer stores to NCCL ncclSend/ncclRecv (communicator from
 * jtb_nccl_unique_id + jtb_slab_nccl_init, or jtb_slab_nccl_init_local); failures return JTB_ERR_NCCL.
 * jtb_slab_status reads the watchdog of the spin-waiting kernels (JTB_ERR_CUDA if a peer never arrived). */
typedef struct jtb_slab jtb_slab;
int jtb_slab_create(jtb_slab** out, int prec, int64_t slices, int64_t rows, int64_t columns, int nranks, int rank, int device);
int jtb_slab_destroy(jtb_slab* m);
int jtb_slab_export(jtb_slab* m, unsigned char* handles192);
int jtb_slab_connect_ipc(jtb_slab* m, const unsigned char* all_handles /* nranks x 192 */);
int jtb_slab_connect_local(jtb_slab* const* members, int n);
int jtb_nccl_unique_id(unsigned char* id128);
int jtb_slab_nccl_init(jtb_slab* m, const unsigned char* id128);
int jtb_slab_nccl_init_local(jtb_slab* const* members, int n);
int jtb_slab_set_exchange(jtb_slab* m, int use_nccl);
int64_t jtb_slab_block_elements(const jtb_slab* m); /* reals in one k2-slabbed block = reals in one slab */
int jtb_slab_forward(jtb_slab* m, void* dev_a, int inverse, int scale, void** result, void* stream);
int jtb_slab_back(jtb_slab* m, void* dev_b, int scale, void** result, void* stream);
int jtb_slab_group_forward(jtb_slab* const* members, int n, void* const* dev_a, int inverse, int scale, void** results,
                           void* const* streams);
int jtb_slab_group_back(jtb_slab* const* members, int n, void* const* dev_b, int scale, void** results,
                        void* const* streams);
int jtb_slab_status(jtb_slab* m);
/* phase timing of the last step (bench): ms3 = {in-slice passes + exchange stores, wait for peers, slice-axis pass} */
int jtb_slab_profile(jtb_slab* m, int enable);
int jtb_slab_last_times(jtb_slab* m, float* ms3);
/* timeline of the last profiled pipelined step (ms since its start, up to 16 column blocks): stored[j] = block j has left
 * on the producer stream, done[j] = slice-axis pass of block j finished on the consumer stream; *nblocks = 0 when the
 * step was not pipelined */
int jtb_slab_chunk_times(jtb_slab* m, int* nblocks, float* stored, float* done);

/* device-side barrier between the ranks on `stream`: publishes `epoch` into every peer's flag array and waits
 * for all peers to publish it (flag_ptrs[h] = peer-mapped int64[nranks] of rank h, zero-initialised). */
int jtb_peer_barrier(int device, void* const* flag_ptrs, int nranks, int rank, int64_t epoch, void* stream);
/* peer-shareable device memory: allocate + 64-byte IPC handle; map a peer's handle; unmap; free */
int jtb_peer_alloc(int device, int64_t bytes, void** dev_ptr, unsigned char* handle64);
int jtb_peer_open(int device, const unsigned char* handle64, void** peer_ptr);
int jtb_peer_close(int device, void* peer_ptr);
int jtb_peer_free(int device, void* dev_ptr);

/* pinned host memory for callers that want DMA-speed jtb_exec (Java: off-heap segments / LargeArray storage) */
int jtb_host_alloc(void** out, int64_t bytes);
int jtb_host_free(void* p);
/* page-lock memory the caller already owns (a Java off-heap segment, the storage of a DoubleLargeArray, a numpy
 * array) so that jtb_exec / jtb_exec_batch copy it at DMA speed; undo with jtb_host_unregister before freeing it */
int jtb_host_register(void* p, int64_t bytes);
int jtb_host_unregister(void* p);

/* synthetic input: a[i] = lo + (hi-lo) * u(seed + i), counter-based (bench and parity harness) */
int jtb_fill_uniform_device(int prec, int device, void* dev_a, int64_t count, uint64_t seed, double lo, double hi,
                            void* stream);

/* test knob: cap the single-pass line length (log2) so small sizes exercise the two-pass path; 0 = default */
int jtb_debug_set_limits(int logn_contig, int logn_strided);

/* test knob: bytes of twiddle / chirp tables currently cached on the device (plans release theirs on destroy) */
int64_t jtb_debug_table_bytes(int device);

int jtb_device_count(void);
int64_t jtb_launch_count(int device); /* kernels launched so far on that device's context */
const char* jtb_last_error(void);
const char* jtb_version(void);

#ifdef __cplusplus
}
#endif
#endif
