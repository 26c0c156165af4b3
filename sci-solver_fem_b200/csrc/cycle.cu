// cycle.cu — stage 3 kernels: SpMV family, fused partition-resident smoother, coarse solve and
// the fused PCG vector updates with device-resident scalars.
//
// Reference semantics:
//   pre :  x = w b/d ; nu1 x { x += w (b - Ain_off x - d x)/d } per partition   preRRSym_kernel1  gauss_seidel.cu:1278-1376
//          r = b - A x (in-partition part :1352-1375 + preAout_kernel :1977-2015) ; bc = R r (:2260)
//   post:  x += P xc (:4425-4427) ; b' = b - Aout x ; nu2 sweeps                postRelaxSym_kernel1 :3626-3736
//   PCG :  CG_Flex_Cycle, cgcycle.cu:6-69 (7 BLAS-1 launches + 3 blocking scalar reads per iteration upstream)
//
// B200 design: fp64, one CSR per level in the partition-contiguous numbering.  The smoother is
// atomic-free (the reference accumulates A_in x with shared-memory float atomics, order-
// nondeterministic): thread t owns row t of the partition, x lives in shared memory, both
// triangles of A_in are read from the row itself.  All reductions are fixed-order (warp
// shuffles, then a last-CTA pass over per-CTA partials), so a solve is bit-reproducible.
// Every kernel exits immediately once the device-side `done` flag is set, which lets the host
// enqueue iterations ahead of the convergence poll.
#include <cooperative_groups.h>

#include "kernels.h"

namespace fsb {

thread_local long long g_launch_counter = 0;

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// fixed-order CTA reduction (blockDim.x multiple of 32, <= 1024); result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* s_warp) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  if (lane == 0) s_warp[wid] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) for (int i = 0; i < nw; i++) t += s_warp[i];
  return t;
}

// partial -> partials[blockIdx]; returns true (to all threads) in the CTA that arrives last,
// which then owns the final fixed-order sum over all partials.
__device__ __forceinline__ bool publish_partial(double blocksum, double* partials, unsigned int* ticket, int* s_flag) {
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = blocksum;
    __threadfence();
    unsigned int tk = atomicInc(ticket, gridDim.x - 1);
    *s_flag = (tk == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  return *s_flag != 0;
}

__device__ __forceinline__ double final_sum(const double* partials, int n, double* s_warp) {
  __threadfence();
  // four independent strided chains per thread (fixed order): the loads of a thread overlap
  double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
  const int B = blockDim.x;
  int i = threadIdx.x;
  for (; i + 3 * B < n; i += 4 * B) {
    v0 += __ldcg(partials + i); v1 += __ldcg(partials + i + B); v2 += __ldcg(partials + i + 2 * B); v3 += __ldcg(partials + i + 3 * B);
  }
  for (; i < n; i += B) v0 += __ldcg(partials + i);
  __syncthreads();
  return block_sum((v0 + v1) + (v2 + v3), s_warp);
}

// ------------------------------------------------------------------ cross-GPU exchange (NVLink peer memory)
// Protocols: see DistDev / LLXchg (common.cuh).  A wait that exceeds ~10 s raises the error flag; once it is
// set every later wait returns immediately (the PCG update kernel turns it into `done`, the host reports it).
__device__ __forceinline__ bool wait_expired(const DistDev& d, long long t0, unsigned& spins) {
  if ((++spins & 1023u) != 0) return false;
  if (*(volatile int*)d.error) return true;
  if (clock64() - t0 > 20000000000ll) { *(volatile int*)d.error = 1; __threadfence_system(); return true; }  // ~10 s: never hang the GPU
  return false;
}
__device__ __forceinline__ void ll_store(uint4* p, double v, unsigned ep) {
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(ep), "r"(hi), "r"(ep) : "memory");
}
// polls a slot until both 8-byte halves carry the epoch; returns the value (0 when a wait timed out)
__device__ __forceinline__ double ll_load(const DistDev& d, const uint4* p, unsigned ep) {
  unsigned lo, f0, hi, f1, spins = 0;
  if (*(volatile int*)d.error) return 0.0;
  const long long t0 = clock64();
  while (true) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(p) : "memory");
    if (f0 == ep && f1 == ep) break;
    if (wait_expired(d, t0, spins)) return 0.0;
  }
  return __hiloint2double((int)hi, (int)lo);
}

// all-reduce(sum) of one double per GPU, executed by the last CTA of a reduction kernel: thread q sends this
// GPU's partial to GPU q and receives GPU q's; every GPU adds the contributions in rank order, so the result is
// bit-identical everywhere.  Slots are double-buffered by the parity of the all-reduce count: a fast GPU's next
// contribution cannot overwrite a slot a slow GPU has not read yet.
__device__ __forceinline__ double cross_sum(const DistDev& d, double local) {
  if (d.nranks == 1) return local;  // uniform
  __shared__ double s_part[kMaxRanks], s_val;
  __shared__ unsigned long long s_ep;
  if (threadIdx.x == 0) { s_val = local; s_ep = ++d.epoch[0]; }  // the partial is valid in thread 0 only
  __syncthreads();
  const unsigned ep = (unsigned)s_ep;
  const int par = (int)(s_ep & 1ull) * kMaxRanks;
  if (threadIdx.x < d.nranks) {
    ll_store(d.peer_red[threadIdx.x] + par + d.rank, s_val, ep);
    s_part[threadIdx.x] = ll_load(d, d.my_red + par + threadIdx.x, ep);
  }
  __syncthreads();
  double tot = 0.0;
  if (threadIdx.x == 0) for (int q = 0; q < d.nranks; q++) tot += s_part[q];
  return tot;
}

// Programmatic dependent launch: every kernel of the solve starts with griddepcontrol.wait
// (predecessor complete, its writes visible) and is launched through FSB_LAUNCH with the
// programmatic-serialization attribute (programmatic edges in the captured graph), which lets the
// launch processing of a kernel overlap the tail of its predecessor.  Measured: 643 -> 594 us per
// iteration with plain stream launches, 599 -> 585 inside the graph.  An early
// griddepcontrol.launch_dependents was tried and dropped: the successor's CTAs then fill whatever
// slots free up first, which packs the small coarse-level grids onto a few SMs (+15 % per iteration).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
inline cudaLaunchConfig_t pdl_config(dim3 grid, dim3 block, size_t smem, cudaStream_t stream) {
  static const bool on = !(getenv("FSB_PDL") && atoi(getenv("FSB_PDL")) == 0);  // tuning knob
  static cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cfg.attrs = attr; cfg.numAttrs = on ? 1 : 0;
  return cfg;
}
#define FSB_LAUNCH(kern, grid, block, smem, stream, ...)                       \
  do {                                                                         \
    cudaLaunchConfig_t cfg_ = pdl_config(dim3(grid), dim3(block), smem, stream); \
    FSB_CUDA(cudaLaunchKernelEx(&cfg_, kern, __VA_ARGS__));                    \
  } while (0)


// One exchange of the sharded solve (halo of a vector, restricted residual to the owners of the next level, coarse
// corrections back, all-gather onto the replicated levels): every GPU sends its list entries of `src` into its
// segment of each destination's receive buffer (coalesced 16-byte flag-in-data stores) and then drains its own
// receive buffer into `dst`.  No fence, no flag word, no barrier: a thread that has seen the epoch in both halves
// of a slot has the value.  The grid is small (<= one CTA per SM, all resident), every CTA sends before it polls.
__global__ void __launch_bounds__(256) ll_exchange_kernel(DistDev d, LLXchg x, const double* __restrict__ src, double* __restrict__ dst,
                                                          unsigned int* ticket, const int* __restrict__ done) {
  pdl_wait();
  __shared__ int s_last;
  if (done && *done) return;
  // the last CTA of the previous exchange kernel bumped the epoch; it is bumped again only after every CTA of this
  // grid has finished (ticket below), i.e. after every CTA has read it
  const unsigned ep = *(volatile unsigned*)d.xchg + 1u;
  const int S = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int k0 = t0; k0 < x.stotal; k0 += 4 * S) {  // four independent entries per trip
    int j[4], q[4];
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int k = k0 + u * S;
      j[u] = k < x.stotal ? __ldg(x.sidx + k) : -1;
      q[u] = 0;
      if (k < x.stotal) while (k >= __ldg(x.sptr + q[u] + 1)) q[u]++;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) v[u] = j[u] >= 0 ? src[j[u]] : 0.0;
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (j[u] >= 0) ll_store(x.peerbuf[q[u]] + (k0 + u * S - __ldg(x.sptr + q[u])), v[u], ep);
  }
  for (int k0 = t0; k0 < x.rtotal; k0 += 8 * S) {  // eight slots polled together: one L2 round trip when the data is already there
    unsigned lo[8], f0[8], hi[8], f1[8];
    int r[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int k = k0 + u * S;
      r[u] = -1;
      if (k < x.rtotal) {
        r[u] = __ldg(x.ridx + k);
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo[u]), "=r"(f0[u]), "=r"(hi[u]), "=r"(f1[u]) : "l"(x.mybuf + k) : "memory");
      }
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (r[u] < 0) continue;
      double v = __hiloint2double((int)hi[u], (int)lo[u]);
      if (f0[u] != ep || f1[u] != ep) v = ll_load(d, x.mybuf + (k0 + u * S), ep);  // not there yet: poll this slot
      dst[r[u]] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (s_last && threadIdx.x == 0) *d.xchg = ep;
}

// all-gather by peer stores with the fenced handshake (end of a solve: every GPU gets the full solution): the owned
// slice [begin, end) of a vector goes into every peer's copy, one system fence per CTA (thread 0 after the CTA
// barrier — fences are cumulative), the last CTA publishes an epoch flag to every peer and waits for theirs.
__global__ void __launch_bounds__(256) push_all_kernel(DistDev d, int begin, int end, const double* __restrict__ src, PeerPtrs dst,
                                                       unsigned int* ticket) {
  pdl_wait();
  __shared__ int s_last;
  __shared__ unsigned long long s_ep;
  for (int j = begin + blockIdx.x * blockDim.x + threadIdx.x; j < end; j += gridDim.x * blockDim.x) {
    const double v = src[j];
    for (int q = 0; q < d.nranks; q++) if (q != d.rank) dst.p[q][j] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) s_ep = ++d.epoch[1];
  __syncthreads();
  if (threadIdx.x < d.nranks && threadIdx.x != d.rank) {
    __threadfence_system();
    *(reinterpret_cast<volatile unsigned long long*>(d.peer_flags[threadIdx.x]) + d.rank) = s_ep;
    volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(d.my_flags) + threadIdx.x;
    unsigned spins = 0;
    const long long t0 = clock64();
    if (!*(volatile int*)d.error)
      while (*f < s_ep) if (wait_expired(d, t0, spins)) break;
    __threadfence_system();
  }
}

// ------------------------------------------------------------------ SpMV family (CSR-stream)
// A CTA owns SPMV_ROWS consecutive rows.  Their entries are contiguous in CSR, so the CTA streams
// them with fully coalesced, deeply unrolled loads (every thread keeps several independent
// val/col/x chains in flight), parks the products in shared memory and then thread t adds up the
// products of row t in column order: fixed summation order, no atomics, long rows handled by tiling.
constexpr int SPMV_ROWS = 256;
constexpr int SPMV_CAP = 4096;  // products per tile (32 KB)

template <int MODE, bool DOT>  // MODE 0: y = A x   1: y = b - A x   2: y += A x   3: y -= A x
__global__ void __launch_bounds__(SPMV_ROWS) csr_stream_kernel(int row_begin, int n, const int* __restrict__ ptr, const int* __restrict__ col,
                                                               const double* __restrict__ val, const double* __restrict__ x,
                                                               double* __restrict__ y, const double* __restrict__ b,
                                                               double* __restrict__ partials, PcgScalars* __restrict__ sc,
                                                               const int* __restrict__ done) {
  pdl_wait();
  __shared__ double prod[SPMV_CAP];
  __shared__ int sptr[SPMV_ROWS + 1];
  __shared__ double s_warp[32];
  __shared__ int s_flag;
  if (done && *done) return;
  const int t = threadIdx.x;
  const int r0 = row_begin + blockIdx.x * SPMV_ROWS;  // rows [row_begin, n): all rows, or this GPU's range of a sharded level
  const int nr = min(SPMV_ROWS, n - r0);
  // every thread learns the CTA's entry range directly (two broadcast loads): the streaming loads
  // below do not wait for a barrier
  const int e0 = __ldg(ptr + r0), e1 = __ldg(ptr + r0 + nr);
  if (t <= nr) sptr[t] = ptr[r0 + t];
  if (t == 0) sptr[nr] = e1;
  double acc = 0.0;
  bool first = true;
  for (int ts = e0; ts < e1; ts += SPMV_CAP) {
    const int te = min(ts + SPMV_CAP, e1);
    int e = ts + t;
    for (; e + 7 * SPMV_ROWS < te; e += 8 * SPMV_ROWS) {  // 8 independent (col, val) -> x chains per thread
      int c[8]; double v[8], xv[8];
#pragma unroll
      for (int u = 0; u < 8; u++) { c[u] = __ldg(col + e + u * SPMV_ROWS); v[u] = __ldg(val + e + u * SPMV_ROWS); }
#pragma unroll
      for (int u = 0; u < 8; u++) xv[u] = __ldg(x + c[u]);
#pragma unroll
      for (int u = 0; u < 8; u++) prod[e + u * SPMV_ROWS - ts] = v[u] * xv[u];
    }
    for (; e + 3 * SPMV_ROWS < te; e += 4 * SPMV_ROWS) {
      int c[4]; double v[4], xv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { c[u] = __ldg(col + e + u * SPMV_ROWS); v[u] = __ldg(val + e + u * SPMV_ROWS); }
#pragma unroll
      for (int u = 0; u < 4; u++) xv[u] = __ldg(x + c[u]);
#pragma unroll
      for (int u = 0; u < 4; u++) prod[e + u * SPMV_ROWS - ts] = v[u] * xv[u];
    }
    for (; e < te; e += SPMV_ROWS) prod[e - ts] = __ldg(val + e) * __ldg(x + __ldg(col + e));
    __syncthreads();
    const int ra = (t < nr) ? sptr[t] : 0, rb = (t < nr) ? sptr[t + 1] : 0;
    const int a = max(ra, ts), bnd = min(rb, te);
    for (int q = a; q < bnd; q++) acc += prod[q - ts];
    if (te < e1) __syncthreads();
    first = false;
  }
  if (first) __syncthreads();
  double contrib = 0.0;
  if (t < nr) {
    const int row = r0 + t;
    if (MODE == 0) y[row] = acc;
    else if (MODE == 1) y[row] = b[row] - acc;
    else if (MODE == 2) y[row] = y[row] + acc;
    else y[row] = y[row] - acc;
    if (DOT) contrib = x[row] * acc;
  }
  if (DOT) {
    double bs = block_sum(contrib, s_warp);
    if (publish_partial(bs, partials, &sc->ticket[0], &s_flag)) {
      double tot = final_sum(partials, gridDim.x, s_warp);
      if (threadIdx.x == 0) { sc->py = tot; sc->alpha = sc->rz_old / tot; }
    }
  }
}

template <int MODE>
__global__ void spmv_vector_kernel(int row_begin, int n, const int* __restrict__ ptr, const int* __restrict__ col,
                                   const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y,
                                   const double* __restrict__ b, const int* __restrict__ done);

template <int MODE, bool DOT>
void spmv_dispatch(const Ctx& c, const DCsr& A, const double* x, double* y, const double* b, double* partials, PcgScalars* sc,
                   const int* done, const char* name, RowRange rr = RowRange()) {
  int n = A.nrows;
  if (n == 0) return;
  g_launch_counter++;
  ProfScope ps(c, name);
  const bool ranged = rr.end >= 0;
  if (ranged && DOT) throw std::runtime_error("row-ranged CSR dot is not implemented (use the SELL copy)");
  const int r0 = ranged ? rr.begin : 0, r1 = ranged ? rr.end : n;
  if (r1 <= r0 && !DOT) return;
  // long rows or a small operator: one warp per row; otherwise (also a row range of short rows — A_out, P of a sharded
  // level) the coalesced stream kernel
  if (!DOT && ((double)A.nnz > 32.0 * n || (!ranged && n < 32768)))
    FSB_LAUNCH((spmv_vector_kernel<MODE>), cdiv((long long)(r1 - r0) * 32, 256), 256, 0, c.stream, r0, r1, A.ptr, A.col, A.val, x, y, b, done);
  else
    FSB_LAUNCH((csr_stream_kernel<MODE, DOT>), cdiv(r1 - r0, SPMV_ROWS), SPMV_ROWS, 0, c.stream, r0, r1, A.ptr, A.col, A.val, x, y, b, partials, sc, done);
  FSB_CHECK_LAUNCH();
}

// SELL-32 SpMV: thread per row, slice-major storage.  Every step of the k loop is one coalesced
// request per warp for col and one for val; four steps are kept in flight together with their x
// gathers.  The row sum runs in column order (same rounding sequence as a sequential CSR loop).
template <int MODE, bool DOT>
__global__ void __launch_bounds__(256) sell_spmv_kernel(DistDev dist, int row_begin, int row_end, int list_begin, int list_end,
                                                        const int* __restrict__ rowmap, int n, const long long* __restrict__ sptr,
                                                        const int* __restrict__ col, const double* __restrict__ val,
                                                        const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ b,
                                                        double* __restrict__ partials, PcgScalars* __restrict__ sc,
                                                        const int* __restrict__ done, int pprev, int finalize) {
  pdl_wait();
  __shared__ double s_warp[32];
  __shared__ int s_flag;
  if (done && *done) return;
  // list position i (rows may be length-sorted inside windows: rowmap); owned rows are [row_begin, row_end).
  // The grid may be smaller than the list (fused dot: fewer partial sums and tickets): grid-stride over 256-row chunks.
  const int lane = threadIdx.x & 31;
  const int nslices = (n + 31) >> 5;
  double contrib = 0.0;
  for (int i = list_begin + blockIdx.x * blockDim.x + threadIdx.x; (i & ~31) < list_end; i += gridDim.x * blockDim.x) {
    const int slice = i >> 5;
    // the row's own vector entries travel together with the matrix stream instead of after it
    const int row = (i < n) ? (rowmap ? __ldg(rowmap + i) : i) : -1;
    const bool mine = row >= row_begin && row < row_end;
    double own = 0.0;
    if (mine) {
      if (MODE == 1) own = b[row];
      else if (MODE >= 2) own = y[row];
      else if (DOT) own = x[row];
    }
    double acc = 0.0;
    if (slice < nslices) {
      const long long base = __ldg(sptr + slice);
      const int K = (int)((__ldg(sptr + slice + 1) - base) >> 5);
      const int* cp = col + base + lane;
      const double* vp = val + base + lane;
      int k = 0;
      for (; k + 8 <= K; k += 8) {  // eight (col, val) -> x chains in flight
        int c[8]; double v[8], xv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) { c[u] = __ldg(cp + (k + u) * 32); v[u] = __ldg(vp + (k + u) * 32); }
#pragma unroll
        for (int u = 0; u < 8; u++) xv[u] = __ldg(x + c[u]);
#pragma unroll
        for (int u = 0; u < 8; u++) acc += v[u] * xv[u];
      }
      for (; k + 4 <= K; k += 4) {
        const int c0 = __ldg(cp + k * 32), c1 = __ldg(cp + (k + 1) * 32), c2 = __ldg(cp + (k + 2) * 32), c3 = __ldg(cp + (k + 3) * 32);
        const double v0 = __ldg(vp + k * 32), v1 = __ldg(vp + (k + 1) * 32), v2 = __ldg(vp + (k + 2) * 32), v3 = __ldg(vp + (k + 3) * 32);
        const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
        acc += v0 * x0; acc += v1 * x1; acc += v2 * x2; acc += v3 * x3;
      }
      for (; k < K; k++) acc += __ldg(vp + k * 32) * __ldg(x + __ldg(cp + k * 32));
    }
    if (mine) {
      if (MODE == 0) y[row] = acc;
      else if (MODE == 1) y[row] = own - acc;
      else if (MODE == 2) y[row] = own + acc;
      else y[row] = own - acc;
      if (DOT) contrib += own * acc;
    }
  }
  if (DOT) {
    // The product may be split over several launches (interior rows while the halo of x is in flight, boundary rows
    // after it): all but the last one only park their CTA sums; the last one finds `pprev` parked sums right in front
    // of its own and finishes the fixed-order sum over all of them.
    double bs = block_sum(contrib, s_warp);
    if (!finalize) {
      if (threadIdx.x == 0) partials[blockIdx.x] = bs;
    } else if (publish_partial(bs, partials, &sc->ticket[0], &s_flag)) {
      double tot = final_sum(partials - pprev, gridDim.x + pprev, s_warp);
      tot = cross_sum(dist, tot);  // all-reduce over the GPUs of a sharded solve (no-op on one GPU)
      if (threadIdx.x == 0) { sc->py = tot; sc->alpha = sc->rz_old / tot; }
    }
  }
}

template <int MODE, bool DOT>
int sell_dispatch(const Ctx& c, const Sell& A, const double* x, double* y, const double* b, double* partials, PcgScalars* sc,
                  const int* done, const char* name, RowRange rr = RowRange(), int pprev = 0, int finalize = 1) {
  if (A.nrows == 0) return 0;
  const int r0 = rr.end >= 0 ? rr.begin : 0, r1 = rr.end >= 0 ? rr.end : A.nrows;
  g_launch_counter++;
  ProfScope ps(c, name);
  // list positions that can hold the owned rows: whole sort windows when the rows are length-sorted
  const int gran = A.window > 0 ? A.window : 32;
  const int l0 = r0 / gran * gran, l1 = std::min(A.nrows, (r1 + gran - 1) / gran * gran);
  if (l1 <= l0 && !DOT) return 0;  // (an empty owned range still takes part in the all-reduce)
  static const int dot_ctas_per_sm = getenv("FSB_DOT_CTAS") ? atoi(getenv("FSB_DOT_CTAS")) : 12;  // tuning knob
  int blocks = std::max(1, cdiv(std::max(0, l1 - l0), 256));
  if (DOT) blocks = std::min(blocks, c.num_sms * dot_ctas_per_sm);
  FSB_LAUNCH((sell_spmv_kernel<MODE, DOT>), blocks, 256, 0, c.stream, c.dist, r0, r1, l0, l1, A.rowmap.size() ? A.rowmap.get() : nullptr, A.nrows,
                                                                       A.sptr, A.col, A.val, x, y, b, partials, sc, done, pprev, finalize);
  FSB_CHECK_LAUNCH();
  return blocks;
}

// debug timestamps (tools only): one CTA stamps clock64 at phase boundaries when g_dbg_on != 0
// (g_dbg_on = (CTA index + 1) | kernel kind << 24; kind 0: cluster smoother, 1: ELL smoother (256-thread class), 3: shared-memory ELL smoother)
__device__ long long g_dbg[64];
__device__ int g_dbg_on = 0;
#ifdef FSB_DEBUG_STAMPS  // build with NVCC_EXTRA=-DFSB_DEBUG_STAMPS for tools/cluster_timing.py, tools/ell_timing.py
#define FSB_STAMPK(kind, k) do { if (g_dbg_on && (g_dbg_on >> 24) == (kind) && blockIdx.x == (g_dbg_on & 0xffffff) - 1 && threadIdx.x == 0 && (k) < 64) g_dbg[k] = clock64(); } while (0)
#else
#define FSB_STAMPK(kind, k) do { } while (0)
#endif
#define FSB_STAMP(k) FSB_STAMPK(0, k)

// ------------------------------------------------------------------ fused smoother
// (a) register-resident ELL kernel (fine levels).  One CTA per partition, one thread per row, rows in
// order of decreasing intra-partition length, so a warp's slab is 32 x Kw (entry k of lane l at
// slab[k*32 + l], 16-bit column = position of the column's row in that order): coalesced loads straight
// into registers, no partition-wide padding.  The nu sweeps + the residual pass then touch HBM no
// more: x lives double-buffered in shared memory in thread order, each thread keeps its own x in a
// register.  Vectors enter and leave through a small staging tile so that global accesses stay
// coalesced although the rows are handled in sorted order.  Threads beyond the partition's rows exit.
// Bank conflicts: the x gather of a half-warp (16 lanes x 8 bytes) is conflict-free only if the 16
// columns fall into 16 different 8-byte banks (or coincide).  The tile therefore holds TWO copies of x
// whose bank mapping differs by half a bank period (copy B starts at element BLOCK+8), and the setup
// (hierarchy.cu: ell_assign_slots) orders every row's entries and picks the copy (bit 15 of the stored
// column) so that the lanes of a half-warp collide as little as possible.
constexpr int ell_min_blocks(int maxk, int block) {  // register budget: 85 per thread up to 16 slots, 128 up to 24, then 255
  const int threads = maxk <= 16 ? 768 : maxk <= 24 ? 512 : 256;
  return threads / block > 0 ? threads / block : 1;
}
template <int MAXK, int BLOCK>
__global__ void __launch_bounds__(BLOCK, ell_min_blocks(MAXK, BLOCK))
smooth_ell_kernel(const EllDesc* __restrict__ desc, const double* __restrict__ ellval, const unsigned short* __restrict__ ellcol,
                  const unsigned short* __restrict__ ellrow, const double* __restrict__ diag, const double* __restrict__ b_src,
                  const int* __restrict__ gather, double* __restrict__ b_int, const double* __restrict__ x_in, double w, int nsweeps,
                  double* __restrict__ x_out, const int* __restrict__ scatter, double* __restrict__ x_ext, double* __restrict__ r_out,
                  const int* __restrict__ done) {
  pdl_wait();
  constexpr int COPYB = BLOCK + 8;             // element offset of the second copy
  __shared__ double sx[2][2 * BLOCK + 8];      // also the staging tile (local row order) on the way in and out
  // one dependent-load hop to everything: the descriptor (and the done flag) first, the slabs right after
  if (BLOCK == 256) FSB_STAMPK(1, 0);
  const uint4* dp = reinterpret_cast<const uint4*>(desc + blockIdx.x);
  const uint4 dh = dp[0];
  uint4 dk[2];
  dk[0] = dp[1];
  dk[1] = BLOCK > 512 ? dp[2] : make_uint4(0, 0, 0, 0);
  const int dn = done ? *done : 0;
  if (dn) return;
  const int r0 = (int)dh.x, np = (int)dh.y, t = threadIdx.x;
  if (BLOCK == 256) FSB_STAMPK(1, 1);
  if (t >= np) return;  // exited threads do not take part in the barriers below
  const int nbar = (np + 31) & ~31;  // barrier over the participating warps only
  // this thread's slab column: widths of the partition's warps are bytes of the descriptor
  int K = 0, skip = 0;
  {
    const unsigned kw[8] = {dk[0].x, dk[0].y, dk[0].z, dk[0].w, dk[1].x, dk[1].y, dk[1].z, dk[1].w};
    const int wq = t >> 5;
#pragma unroll
    for (int j = 0; j < BLOCK / 32; j++) {
      const int kj = (int)((kw[j >> 2] >> (8 * (j & 3))) & 0xffu);
      if (j < wq) skip += kj;
      if (j == wq) K = kj;  // warp-uniform
    }
  }
  const long long wbase = ((long long)dh.z | ((long long)dh.w << 32)) + 32LL * skip + (t & 31);
  const double* ev = ellval + wbase;
  const unsigned short* ec = ellcol + wbase;
  double v[MAXK];
  unsigned cpk[MAXK / 2];  // two 16-bit byte offsets into the x tile per register
#pragma unroll
  for (int k = 0; k < MAXK; k += 2) {
    unsigned c0 = 0, c1 = 0;
    v[k] = 0.0; v[k + 1] = 0.0;
    if (k < K) { v[k] = ev[32 * k]; c0 = ec[32 * k]; v[k + 1] = ev[32 * (k + 1)]; c1 = ec[32 * (k + 1)]; }  // K is even
    c0 = (c0 & 0x7fffu) + (c0 >> 15) * COPYB;  // element index inside the two-copy tile
    c1 = (c1 & 0x7fffu) + (c1 >> 15) * COPYB;
    cpk[k >> 1] = (c0 << 3) | (c1 << 19);
  }
  const int lr = ellrow[r0 + t];
  {
    const int row = r0 + t;
    const double bv = b_src[gather ? gather[row] : row], dv = diag[row];
    if (b_int) b_int[row] = bv;
    sx[1][t] = bv; sx[1][BLOCK + t] = dv;
    sx[0][t] = x_in ? x_in[row] : w * bv / dv;
  }
  if (BLOCK == 256) FSB_STAMPK(1, 2);
  asm volatile("bar.sync 1, %0;" ::"r"(nbar));
  if (BLOCK == 256) FSB_STAMPK(1, 3);
  const double b = sx[1][lr], d = sx[1][BLOCK + lr];
  double xv = sx[0][lr];
  const double wd = w / d;  // one division per stage instead of one per sweep (rounding-level deviation)
  asm volatile("bar.sync 1, %0;" ::"r"(nbar));
  sx[0][t] = xv; sx[0][COPYB + t] = xv;
  asm volatile("bar.sync 1, %0;" ::"r"(nbar));
  int cur = 0;
  const int npasses = nsweeps + (r_out ? 1 : 0);
  double res = 0.0;
  if (BLOCK == 256) FSB_STAMPK(1, 4);
  for (int it = 0; it < npasses; it++) {
    if (BLOCK == 256) FSB_STAMPK(1, 5 + it);
    const char* xs = reinterpret_cast<const char*>(sx[cur]);
    double s0 = 0.0, s1 = 0.0;  // two independent FMA chains (fixed order)
#pragma unroll
    for (int k = 0; k < MAXK; k += 2) {
      if (k < K) {
        s0 += v[k] * *reinterpret_cast<const double*>(xs + (cpk[k >> 1] & 0xffffu));
        s1 += v[k + 1] * *reinterpret_cast<const double*>(xs + (cpk[k >> 1] >> 16));
      }
    }
    const double s = s0 + s1;
    if (it < nsweeps) {
      xv = xv + wd * (b - s - d * xv);
      sx[cur ^ 1][t] = xv; sx[cur ^ 1][COPYB + t] = xv;
      asm volatile("bar.sync 1, %0;" ::"r"(nbar));
      cur ^= 1;
    } else {
      res = b - s - d * xv;  // in-partition residual b - A_in x - d x (gauss_seidel.cu:1352-1375)
    }
  }
  if (BLOCK == 256) FSB_STAMPK(1, 12);
  // back to row order through the tile that is not being read (its last readers passed a barrier)
  double* so = sx[cur ^ 1];
  so[lr] = xv;
  if (r_out) so[BLOCK + lr] = res;
  asm volatile("bar.sync 1, %0;" ::"r"(nbar));
  const int row = r0 + t;
  const double xo = so[t];
  if (x_out) x_out[row] = xo;
  if (scatter) x_ext[scatter[row]] = xo;
  if (r_out) r_out[row] = so[BLOCK + t];
  if (BLOCK == 256) FSB_STAMPK(1, 13);
}

// (a') shared-memory variant of (a) for the coarser levels (rows too long for registers, few
// partitions).  Same sorted slabs, but G lanes share a row (lane g owns the row's entries g, g+G, ...)
// and the slabs of the partition are brought into shared memory once by two TMA bulk copies
// (cp.async.bulk -> mbarrier); whatever exceeds the shared-memory budget streams from L2 in every pass.
// One CTA of 1024 threads per partition; up to SELLG_NVB virtual-row blocks per thread.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}


template <int SELLG_BLOCK>
__global__ void __launch_bounds__(SELLG_BLOCK, 1024 / SELLG_BLOCK)
smooth_sellg_kernel(const SellgDesc* __restrict__ desc, int G, int npad, int smemSlots, const long long* __restrict__ wptr,
                    const double* __restrict__ ellval, const unsigned short* __restrict__ ellcol, const unsigned short* __restrict__ ellrow,
                    const double* __restrict__ diag, const double* __restrict__ b_src, const int* __restrict__ gather,
                    double* __restrict__ b_int, const double* __restrict__ x_in, double w, int nsweeps, double* __restrict__ x_out,
                    const int* __restrict__ scatter, double* __restrict__ x_ext, double* __restrict__ r_out, const int* __restrict__ done) {
  pdl_wait();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: val[smemSlots] | x tiles 2 x (2*npad + 8) | b, d, w/d in sorted order 3 x npad | mbarrier | col[smemSlots]
  const int COPYB = npad + 8, TILE = 2 * npad + 8;  // npad: multiple of 16 >= rows of the largest partition
  double* sval = reinterpret_cast<double*>(smem_raw);
  double* sx0 = sval + smemSlots;
  double* sb = sx0 + 2 * TILE;
  double* sd = sb + npad;
  double* swd = sd + npad;
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(swd + npad);
  unsigned short* scol = reinterpret_cast<unsigned short*>(mbar + 2);
  constexpr int SELLG_NVB = 4;  // virtual-row blocks per thread: rows x lanes per row <= 4 * SELLG_BLOCK (split_partitions)
  const int t = threadIdx.x, lane = t & 31;
  const uint4* dp = reinterpret_cast<const uint4*>(desc + blockIdx.x);
  const uint4 d0 = dp[0], d1 = dp[1];
  const int dn = done ? *done : 0;
  if (dn) return;
  FSB_STAMPK(3, 0);
  const int r0 = (int)d0.x, np = (int)d0.y, w0 = (int)d0.z;
  const long long base = (long long)d1.x | ((long long)d1.y << 32);
  const int slots = (int)d1.z, inSmem = min(slots, smemSlots);
  const unsigned bar = smem_u32(mbar);
  if (t == 0) {
    mbar_init(bar, 1);
    mbar_expect_tx(bar, (unsigned)inSmem * 10u);
    if (inSmem > 0) {
      bulk_g2s(smem_u32(sval), ellval + base, (unsigned)inSmem * 8u, bar);
      bulk_g2s(smem_u32(scol), ellcol + base, (unsigned)inSmem * 2u, bar);
    }
  }
  // this thread's virtual rows: slab offset and width of their warps
  const int nvr = np * G, gsh = 31 - __clz(G);
  int woff[SELLG_NVB], K[SELLG_NVB];
#pragma unroll
  for (int vb = 0; vb < SELLG_NVB; vb++) {
    const int vt = vb * SELLG_BLOCK + t;
    woff[vb] = 0; K[vb] = 0;
    if ((vt & ~31) < nvr) {
      const long long a = wptr[w0 + (vt >> 5)], b = wptr[w0 + (vt >> 5) + 1];
      woff[vb] = (int)(a - base); K[vb] = (int)((b - a) >> 5);
    }
  }
  // vectors: in through the tiles in row order, then into sorted order (np <= 1024: at most 2 rows per thread)
  constexpr int NR = 1024 / SELLG_BLOCK;
  for (int i = t; i < np; i += SELLG_BLOCK) {
    const int row = r0 + i;
    const double bv = b_src[gather ? gather[row] : row], dv = diag[row];
    if (b_int) b_int[row] = bv;
    double* s1 = sx0 + TILE;
    s1[i] = bv; s1[npad + i] = dv;
    sx0[i] = x_in ? x_in[row] : w * bv / dv;
  }
  __syncthreads();
  double bq[NR], dq[NR], xq[NR];
#pragma unroll
  for (int j = 0; j < NR; j++) {
    const int i = t + j * SELLG_BLOCK;
    bq[j] = 0.0; dq[j] = 1.0; xq[j] = 0.0;
    if (i < np) { const int lr = ellrow[r0 + i]; const double* s1 = sx0 + TILE; bq[j] = s1[lr]; dq[j] = s1[npad + lr]; xq[j] = sx0[lr]; }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NR; j++) {
    const int i = t + j * SELLG_BLOCK;
    if (i < np) { sb[i] = bq[j]; sd[i] = dq[j]; swd[i] = w / dq[j]; sx0[i] = xq[j]; sx0[COPYB + i] = xq[j]; }
  }
  FSB_STAMPK(3, 1);
  mbar_wait(bar, 0);
  FSB_STAMPK(3, 2);
  __syncthreads();
  int cb = 0;
  const int npasses = nsweeps + (r_out ? 1 : 0);
  for (int it = 0; it < npasses; it++) {
    FSB_STAMPK(3, 3 + it);
    const double* xs = sx0 + cb * TILE;
    double* xn = sx0 + (cb ^ 1) * TILE;
    const bool sweep = it < nsweeps;
#pragma unroll
    for (int vb = 0; vb < SELLG_NVB; vb++) {
      const int vt = vb * SELLG_BLOCK + t;
      if ((vt & ~31) >= nvr) break;  // warp-uniform
      const int kk = K[vb], off = woff[vb] + lane;
      const bool in = woff[vb] + 32 * kk <= inSmem;  // the whole warp slab is shared-memory resident
      const double* pv = in ? sval + off : ellval + base + off;
      const unsigned short* pc = in ? scol + off : ellcol + base + off;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
      for (int k = 0; k < kk; k += 2) {
        const unsigned c0 = pc[32 * k], c1 = pc[32 * k + 32];
        const double v0 = pv[32 * k], v1 = pv[32 * k + 32];
        s0 += v0 * xs[(c0 & 0x7fffu) + (c0 >> 15) * COPYB];
        s1 += v1 * xs[(c1 & 0x7fffu) + (c1 >> 15) * COPYB];
      }
      double sum = s0 + s1;
      for (int o = G >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const int srow = vt >> gsh;
      if ((vt & (G - 1)) == 0 && srow < np) {
        const double xv = xs[srow];
        if (sweep) {
          const double xnew = xv + swd[srow] * (sb[srow] - sum - sd[srow] * xv);
          xn[srow] = xnew; xn[COPYB + srow] = xnew;
        } else {
          sb[srow] = sb[srow] - sum - sd[srow] * xv;  // in-partition residual, parked in sb (each row reads only its own b)
        }
      }
    }
    __syncthreads();
    if (sweep) cb ^= 1;
  }
  FSB_STAMPK(3, 10);
  // back to row order through the tile that is not being read
  double* so = sx0 + (cb ^ 1) * TILE;
  for (int i = t; i < np; i += SELLG_BLOCK) {
    const int lr = ellrow[r0 + i];
    so[lr] = sx0[cb * TILE + i];
    if (r_out) so[npad + lr] = sb[i];
  }
  __syncthreads();
  for (int i = t; i < np; i += SELLG_BLOCK) {
    const int row = r0 + i;
    const double xo = so[i];
    if (x_out) x_out[row] = xo;
    if (scatter) x_ext[scatter[row]] = xo;
    if (r_out) r_out[row] = so[npad + i];
  }
  FSB_STAMPK(3, 11);
}

// (b) cooperative CSR kernel (coarse levels: long rows, few partitions).  G lanes share a row,
// rows of the partition are processed BLOCK/G at a time, matrix entries stream from L1/L2.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) smooth_coop_kernel(int G, const int* __restrict__ pstart, const int* __restrict__ ptr,
                                                            const int* __restrict__ col, const double* __restrict__ val,
                                                            const double* __restrict__ diag, const double* __restrict__ b_src,
                                                            const int* __restrict__ gather, double* __restrict__ b_int,
                                                            const double* __restrict__ x_in, double w, int nsweeps,
                                                            double* __restrict__ x_out, const int* __restrict__ scatter,
                                                            double* __restrict__ x_ext, double* __restrict__ r_out,
                                                            const int* __restrict__ done) {
  pdl_wait();
  __shared__ double sx[2][1024], sb[1024], sd[1024], swd[1024];
  if (done && *done) return;
  const int r0 = pstart[blockIdx.x], np = pstart[blockIdx.x + 1] - r0, tid = threadIdx.x;
  for (int t = tid; t < np; t += BLOCK) {
    const int row = r0 + t;
    const double b = b_src[gather ? gather[row] : row], d = diag[row];
    sb[t] = b; sd[t] = d; swd[t] = w / d;
    if (b_int) b_int[row] = b;
    sx[0][t] = x_in ? x_in[row] : w * b / d;
  }
  __syncthreads();
  const int lane = tid & (G - 1), grp = tid / G, RP = BLOCK / G;
  const int npasses = nsweeps + (r_out ? 1 : 0);
  int cur = 0;
  for (int it = 0; it < npasses; it++) {
    const bool sweep = it < nsweeps;
    for (int rb = 0; rb < np; rb += RP) {
      const int t = rb + grp;
      double s = 0.0;
      if (t < np) {
        const int row = r0 + t, e1 = ptr[row + 1];
        for (int e = ptr[row] + lane; e < e1; e += G) {
          const int cc = col[e] - r0;
          if ((unsigned)cc < (unsigned)np && cc != t) s += val[e] * sx[cur][cc];
        }
      }
      for (int o = G >> 1; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, G);
      if (t < np && lane == 0) {
        const double xv = sx[cur][t];
        if (sweep) sx[cur ^ 1][t] = xv + swd[t] * (sb[t] - s - sd[t] * xv);
        else r_out[r0 + t] = sb[t] - s - sd[t] * xv;
      }
    }
    __syncthreads();
    if (sweep) cur ^= 1;
  }
  for (int t = tid; t < np; t += BLOCK) {
    const double xv = sx[cur][t];
    if (x_out) x_out[r0 + t] = xv;
    if (scatter) x_ext[scatter[r0 + t]] = xv;
  }
}

// (c) cluster kernel (coarse levels: long rows, few partitions).  A partition is owned by a
// thread-block CLUSTER of C CTAs: every CTA stages the CSR slice of its np/C rows in shared memory
// once (16-bit local columns; diagonal and inter-partition entries neutralised to 0 * x[t]) and
// keeps a full copy of the partition's x.  After each sweep the CTAs push their new x values into
// all C copies through distributed shared memory and meet at a cluster barrier, so the whole
// smoothing stage runs out of shared memory on C SMs per partition instead of one.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) smooth_cluster_kernel(int C, int G, int cap, int npmax, int chunkmax,
                                                               const int* __restrict__ pstart, const int* __restrict__ ptr,
                                                               const int* __restrict__ col, const double* __restrict__ val,
                                                               const double* __restrict__ diag, const double* __restrict__ b_src,
                                                               const int* __restrict__ gather, double* __restrict__ b_int,
                                                               const double* __restrict__ x_in, double w, int nsweeps,
                                                               double* __restrict__ x_out, const int* __restrict__ scatter,
                                                               double* __restrict__ x_ext, double* __restrict__ r_out,
                                                               const int* __restrict__ done) {
  pdl_wait();
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sval = reinterpret_cast<double*>(smem_raw);
  double* sx0 = sval + cap;
  double* sx1 = sx0 + npmax;
  double* sb = sx1 + npmax;
  double* sd = sb + chunkmax;
  double* swd = sd + chunkmax;
  int* srp = reinterpret_cast<int*>(swd + chunkmax);
  unsigned short* scol = reinterpret_cast<unsigned short*>(srp + chunkmax + 1);
  if (done && *done) return;  // uniform over the whole grid
  FSB_STAMP(0);
  const int crank = (int)cluster.block_rank();
  const int p = blockIdx.x / C;
  const int r0 = pstart[p], np = pstart[p + 1] - r0, tid = threadIdx.x;
  const int chunk = (np + C - 1) / C;
  const int m0 = min(crank * chunk, np);
  const int mn = min(chunk, np - m0);  // this CTA owns local rows [m0, m0 + mn)
  // full initial x (every CTA computes its own copy), own rows' b / d / slice
  for (int t = tid; t < np; t += BLOCK) {
    const int row = r0 + t;
    double x0;
    if (x_in) x0 = x_in[row];
    else x0 = w * b_src[gather ? gather[row] : row] / diag[row];
    sx0[t] = x0;
  }
  const int e0 = ptr[r0 + m0];
  for (int t = tid; t < mn; t += BLOCK) {
    const int row = r0 + m0 + t;
    const double b = b_src[gather ? gather[row] : row], d = diag[row];
    sb[t] = b; sd[t] = d; swd[t] = w / d;
    if (b_int) b_int[row] = b;
  }
  for (int t = tid; t <= mn; t += BLOCK) srp[t] = ptr[r0 + m0 + t] - e0;
  __syncthreads();
  FSB_STAMP(1);
  const int lane = tid & (G - 1), grp = tid / G, RP = BLOCK / G;
  {
    // stage the slice: flat, fully coalesced, every thread keeps several independent loads in flight;
    // columns become partition-local (0xFFFF = inter-partition entry) ...
    const int m = srp[mn];
    for (int q = tid; q < m; q += BLOCK) {
      const int cc = __ldg(col + e0 + q) - r0;
      sval[q] = __ldg(val + e0 + q);
      scol[q] = ((unsigned)cc < (unsigned)np) ? (unsigned short)cc : (unsigned short)0xFFFFu;
    }
    __syncthreads();
    // ... then the diagonal and the inter-partition entries are neutralised to 0 * x[own row]
    for (int t = grp; t < mn; t += RP) {
      const int qb = srp[t + 1], tl = m0 + t;
      for (int q = srp[t] + lane; q < qb; q += G) {
        const unsigned short cc = scol[q];
        if (cc == 0xFFFFu || cc == tl) { sval[q] = 0.0; scol[q] = (unsigned short)tl; }
      }
    }
  }
  FSB_STAMP(2);
  cluster.sync();  // all CTAs of the cluster are resident and initialised before remote writes start
  FSB_STAMP(3);
  double* xc = sx0;
  double* xn = sx1;
  for (int it = 0; it < nsweeps; it++) {
    for (int rb = 0; rb < mn; rb += RP) {
      const int t = rb + grp;
      double s = 0.0;
      if (t < mn) {
        // batches of 8 entries per lane: all column loads, then all x / value loads, then the FMAs —
        // the dependent chain is two shared-memory latencies per batch instead of two per entry
        const int qb = srp[t + 1];
        double s1 = 0.0;
        for (int q = srp[t] + lane; q < qb; q += 8 * G) {
          int cc[8]; double vv[8], xx[8];
#pragma unroll
          for (int u = 0; u < 8; u++) { const int qq = q + u * G; cc[u] = qq < qb ? scol[qq] : 0; vv[u] = qq < qb ? sval[qq] : 0.0; }
#pragma unroll
          for (int u = 0; u < 8; u++) xx[u] = xc[cc[u]];
#pragma unroll
          for (int u = 0; u < 8; u += 2) { s += vv[u] * xx[u]; s1 += vv[u + 1] * xx[u + 1]; }
        }
        s += s1;
      }
      for (int o = G >> 1; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, G);
      if (t < mn && lane == 0) {
        const double xv = xc[m0 + t];
        const double xnew = xv + swd[t] * (sb[t] - s - sd[t] * xv);
        for (int c = 0; c < C; c++) cluster.map_shared_rank(xn, c)[m0 + t] = xnew;  // DSMEM broadcast
      }
    }
    FSB_STAMP(4 + 2 * it);
    if (C == 1) __syncthreads(); else cluster.sync();
    FSB_STAMP(5 + 2 * it);
    double* tmp = xc; xc = xn; xn = tmp;
  }
  if (r_out) {
    for (int rb = 0; rb < mn; rb += RP) {
      const int t = rb + grp;
      double s = 0.0;
      if (t < mn) {
        // batches of 8 entries per lane: all column loads, then all x / value loads, then the FMAs —
        // the dependent chain is two shared-memory latencies per batch instead of two per entry
        const int qb = srp[t + 1];
        double s1 = 0.0;
        for (int q = srp[t] + lane; q < qb; q += 8 * G) {
          int cc[8]; double vv[8], xx[8];
#pragma unroll
          for (int u = 0; u < 8; u++) { const int qq = q + u * G; cc[u] = qq < qb ? scol[qq] : 0; vv[u] = qq < qb ? sval[qq] : 0.0; }
#pragma unroll
          for (int u = 0; u < 8; u++) xx[u] = xc[cc[u]];
#pragma unroll
          for (int u = 0; u < 8; u += 2) { s += vv[u] * xx[u]; s1 += vv[u + 1] * xx[u + 1]; }
        }
        s += s1;
      }
      for (int o = G >> 1; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, G);
      if (t < mn && lane == 0) r_out[r0 + m0 + t] = sb[t] - s - sd[t] * xc[m0 + t];
    }
  }
  FSB_STAMP(40);
  for (int t = tid; t < mn; t += BLOCK) {
    const double xv = xc[m0 + t];
    if (x_out) x_out[r0 + m0 + t] = xv;
    if (scatter) x_ext[scatter[r0 + m0 + t]] = xv;
  }
  FSB_STAMP(41);
}

// (d) dense partition blocks (dense_tail.cu: build_block_smoothers): a whole smoothing stage of a partition as small
// dense GEMVs.  One CTA per 32 rows of a partition: the partition's vector(s) in shared memory, a warp per row, lanes
// stride over the columns (coalesced), four independent accumulators per lane (fixed order).
//   pre  (x_in == null): x = S1 b            post: x = Gp x_in + S2 b
template <int ROWS>
__global__ void __launch_bounds__(256) smooth_blockdense_kernel(const int* __restrict__ work, const int* __restrict__ pstart,
                                                                const long long* __restrict__ off, const double* __restrict__ S1,
                                                                const double* __restrict__ Gp, const double* __restrict__ S2,
                                                                const double* __restrict__ b_src, const int* __restrict__ gather,
                                                                double* __restrict__ b_int, const double* __restrict__ x_in,
                                                                double* __restrict__ x_out, const int* __restrict__ scatter,
                                                                double* __restrict__ x_ext, const int* __restrict__ done) {
  pdl_wait();
  __shared__ double sb[1024], sx[1024];
  if (done && *done) return;
  const int p = work[2 * blockIdx.x], lr0 = work[2 * blockIdx.x + 1];
  const int r0 = pstart[p], m = pstart[p + 1] - r0;
  const long long base = off[p];
  for (int c = threadIdx.x; c < m; c += blockDim.x) {
    const double bv = b_src[gather ? gather[r0 + c] : r0 + c];
    sb[c] = bv;
    if (x_in) sx[c] = x_in[r0 + c];
    if (b_int && c >= lr0 && c < lr0 + ROWS) b_int[r0 + c] = bv;  // every row of the partition is saved by exactly one CTA
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a warp owns ROWS / 8 = 4 consecutive rows and streams them together: 4 (8 in the post stage) independent row
  // streams per lane, two columns each in flight; every row sum is reduced in a fixed order
  constexpr int RW = ROWS / 8;
  const int lrw = lr0 + warp * RW;
  double acc[RW][2];
#pragma unroll
  for (int j = 0; j < RW; j++) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
  const int nrow = max(0, min(RW, m - lrw));  // rows of this warp that exist
  {
    const double* A = (x_in ? S2 : S1) + base + (size_t)lrw * m;
    int c = lane;
    for (; c + 32 < m; c += 64) {
      const double v0 = sb[c], v1 = sb[c + 32];
#pragma unroll
      for (int j = 0; j < RW; j++)
        if (j < nrow) { acc[j][0] += A[(size_t)j * m + c] * v0; acc[j][1] += A[(size_t)j * m + c + 32] * v1; }
    }
    if (c < m) {
      const double v0 = sb[c];
#pragma unroll
      for (int j = 0; j < RW; j++) if (j < nrow) acc[j][0] += A[(size_t)j * m + c] * v0;
    }
  }
  if (x_in) {
    const double* G = Gp + base + (size_t)lrw * m;
    int c = lane;
    for (; c + 32 < m; c += 64) {
      const double v0 = sx[c], v1 = sx[c + 32];
#pragma unroll
      for (int j = 0; j < RW; j++)
        if (j < nrow) { acc[j][0] += G[(size_t)j * m + c] * v0; acc[j][1] += G[(size_t)j * m + c + 32] * v1; }
    }
    if (c < m) {
      const double v0 = sx[c];
#pragma unroll
      for (int j = 0; j < RW; j++) if (j < nrow) acc[j][0] += G[(size_t)j * m + c] * v0;
    }
  }
#pragma unroll
  for (int j = 0; j < RW; j++) {
    const double s = warp_sum(acc[j][0] + acc[j][1]);
    if (lane == 0 && j < nrow) {
      if (x_out) x_out[r0 + lrw + j] = s;
      if (scatter) x_ext[scatter[r0 + lrw + j]] = s;
    }
  }
}

// warp-per-row SpMV for long rows (restriction operators, coarse operators)
template <int MODE>
__global__ void __launch_bounds__(256) spmv_vector_kernel(int row_begin, int n, const int* __restrict__ ptr,
                                                          const int* __restrict__ col, const double* __restrict__ val,
                                                          const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ b,
                                                          const int* __restrict__ done) {
  pdl_wait();
  if (done && *done) return;
  const int row = row_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  double s = 0.0;
  const int e1 = ptr[row + 1];
  int e = ptr[row] + lane;
  for (; e + 96 < e1; e += 128) {  // four independent (col, val) -> x chains per lane: two latency hops per 128 entries
    const int c0 = __ldg(col + e), c1 = __ldg(col + e + 32), c2 = __ldg(col + e + 64), c3 = __ldg(col + e + 96);
    const double v0 = __ldg(val + e), v1 = __ldg(val + e + 32), v2 = __ldg(val + e + 64), v3 = __ldg(val + e + 96);
    const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
    s += v0 * x0; s += v1 * x1; s += v2 * x2; s += v3 * x3;
  }
  {  // tail of up to 4 x 32 entries, predicated so that the loads still issue together
    const bool p0 = e < e1, p1 = e + 32 < e1, p2 = e + 64 < e1, p3 = e + 96 < e1;
    const int c0 = p0 ? __ldg(col + e) : 0, c1 = p1 ? __ldg(col + e + 32) : 0, c2 = p2 ? __ldg(col + e + 64) : 0, c3 = p3 ? __ldg(col + e + 96) : 0;
    const double v0 = p0 ? __ldg(val + e) : 0.0, v1 = p1 ? __ldg(val + e + 32) : 0.0, v2 = p2 ? __ldg(val + e + 64) : 0.0, v3 = p3 ? __ldg(val + e + 96) : 0.0;
    const double x0 = p0 ? __ldg(x + c0) : 0.0, x1 = p1 ? __ldg(x + c1) : 0.0, x2 = p2 ? __ldg(x + c2) : 0.0, x3 = p3 ? __ldg(x + c3) : 0.0;
    s += v0 * x0; s += v1 * x1; s += v2 * x2; s += v3 * x3;
  }
  s = warp_sum(s);
  if (lane == 0) {
    if (MODE == 0) y[row] = s;
    else if (MODE == 1) y[row] = b[row] - s;
    else if (MODE == 2) y[row] = y[row] + s;
    else y[row] = y[row] - s;
  }
}

// y[r] += (A x)[r] for the rows r of a list (ghost copies of a sharded level apply the coarse correction themselves:
// x_ghost += P_ghost xc); rows are short (a prolongator row has <= 8 entries): one thread per listed row
__global__ void __launch_bounds__(256) spmv_list_add_kernel(int count, const int* __restrict__ rows, const int* __restrict__ ptr,
                                                            const int* __restrict__ col, const double* __restrict__ val,
                                                            const double* __restrict__ x, double* __restrict__ y, const int* __restrict__ done) {
  pdl_wait();
  if (done && *done) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const int r = rows[k];
  double s = 0.0;
  for (int e = ptr[r]; e < ptr[r + 1]; e++) s += val[e] * __ldg(x + col[e]);  // same order as the owner's row sum
  y[r] = y[r] + s;
}

// bc = R b - T x, T = R A: warp per coarse row, first the row of R against b, then the row of T against x;
// four independent (col, val) -> vector chains per lane in both parts
__device__ __forceinline__ double warp_row_dot(const int* __restrict__ ptr, const int* __restrict__ col, const double* __restrict__ val,
                                               const double* __restrict__ v, int row, int lane) {
  double s = 0.0;
  const int e1 = ptr[row + 1];
  int e = ptr[row] + lane;
  for (; e + 96 < e1; e += 128) {
    const int c0 = __ldg(col + e), c1 = __ldg(col + e + 32), c2 = __ldg(col + e + 64), c3 = __ldg(col + e + 96);
    const double v0 = __ldg(val + e), v1 = __ldg(val + e + 32), v2 = __ldg(val + e + 64), v3 = __ldg(val + e + 96);
    const double x0 = __ldg(v + c0), x1 = __ldg(v + c1), x2 = __ldg(v + c2), x3 = __ldg(v + c3);
    s += v0 * x0; s += v1 * x1; s += v2 * x2; s += v3 * x3;
  }
  {
    const bool p0 = e < e1, p1 = e + 32 < e1, p2 = e + 64 < e1, p3 = e + 96 < e1;
    const int c0 = p0 ? __ldg(col + e) : 0, c1 = p1 ? __ldg(col + e + 32) : 0, c2 = p2 ? __ldg(col + e + 64) : 0, c3 = p3 ? __ldg(col + e + 96) : 0;
    const double v0 = p0 ? __ldg(val + e) : 0.0, v1 = p1 ? __ldg(val + e + 32) : 0.0, v2 = p2 ? __ldg(val + e + 64) : 0.0, v3 = p3 ? __ldg(val + e + 96) : 0.0;
    const double x0 = p0 ? __ldg(v + c0) : 0.0, x1 = p1 ? __ldg(v + c1) : 0.0, x2 = p2 ? __ldg(v + c2) : 0.0, x3 = p3 ? __ldg(v + c3) : 0.0;
    s += v0 * x0; s += v1 * x1; s += v2 * x2; s += v3 * x3;
  }
  return s;
}
__global__ void __launch_bounds__(256) restrict_fused_kernel(int nc, const int* __restrict__ rptr, const int* __restrict__ rcol,
                                                             const double* __restrict__ rval, const double* __restrict__ b,
                                                             const int* __restrict__ tptr, const int* __restrict__ tcol,
                                                             const double* __restrict__ tval, const double* __restrict__ x,
                                                             double* __restrict__ bc, const int* __restrict__ done) {
  pdl_wait();
  if (done && *done) return;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= nc) return;
  const double s = warp_sum(warp_row_dot(rptr, rcol, rval, b, row, lane) - warp_row_dot(tptr, tcol, tval, x, row, lane));
  if (lane == 0) bc[row] = s;
}

// the same product with one 128-thread CTA per coarse row (few, long rows: R A has several hundred entries per row)
__device__ __forceinline__ double cta_row_dot(const int* __restrict__ ptr, const int* __restrict__ col, const double* __restrict__ val,
                                              const double* __restrict__ v, int row, int t) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  const int e1 = ptr[row + 1];
  int e = ptr[row] + t;
  for (; e + 384 < e1; e += 512) {
    const int c0 = __ldg(col + e), c1 = __ldg(col + e + 128), c2 = __ldg(col + e + 256), c3 = __ldg(col + e + 384);
    const double v0 = __ldg(val + e), v1 = __ldg(val + e + 128), v2 = __ldg(val + e + 256), v3 = __ldg(val + e + 384);
    s0 += v0 * __ldg(v + c0); s1 += v1 * __ldg(v + c1); s2 += v2 * __ldg(v + c2); s3 += v3 * __ldg(v + c3);
  }
  {
    const bool p0 = e < e1, p1 = e + 128 < e1, p2 = e + 256 < e1, p3 = e + 384 < e1;
    const int c0 = p0 ? __ldg(col + e) : 0, c1 = p1 ? __ldg(col + e + 128) : 0, c2 = p2 ? __ldg(col + e + 256) : 0, c3 = p3 ? __ldg(col + e + 384) : 0;
    const double v0 = p0 ? __ldg(val + e) : 0.0, v1 = p1 ? __ldg(val + e + 128) : 0.0, v2 = p2 ? __ldg(val + e + 256) : 0.0, v3 = p3 ? __ldg(val + e + 384) : 0.0;
    s0 += v0 * (p0 ? __ldg(v + c0) : 0.0); s1 += v1 * (p1 ? __ldg(v + c1) : 0.0); s2 += v2 * (p2 ? __ldg(v + c2) : 0.0); s3 += v3 * (p3 ? __ldg(v + c3) : 0.0);
  }
  return (s0 + s1) + (s2 + s3);
}
__global__ void __launch_bounds__(128) restrict_fused_row_kernel(int nc, const int* __restrict__ rptr, const int* __restrict__ rcol,
                                                                 const double* __restrict__ rval, const double* __restrict__ b,
                                                                 const int* __restrict__ tptr, const int* __restrict__ tcol,
                                                                 const double* __restrict__ tval, const double* __restrict__ x,
                                                                 double* __restrict__ bc, const int* __restrict__ done) {
  pdl_wait();
  __shared__ double s_warp[4];
  if (done && *done) return;
  const int row = blockIdx.x, t = threadIdx.x;
  const double w = warp_sum(cta_row_dot(rptr, rcol, rval, b, row, t) - cta_row_dot(tptr, tcol, tval, x, row, t));
  if ((t & 31) == 0) s_warp[t >> 5] = w;
  __syncthreads();
  if (t == 0) bc[row] = (s_warp[0] + s_warp[1]) + (s_warp[2] + s_warp[3]);
}

// x = M b, dense, warp per row: the coarsest level's inverse (replaces the host LU round trip of amg_level.cu:25-31)
// or the dense tail of the V-cycle (dense_tail.cu)
__global__ void __launch_bounds__(256) coarse_gemv_kernel(int n, const double* __restrict__ Ainv, const double* __restrict__ b,
                                                          double* __restrict__ x, const int* __restrict__ done) {
  pdl_wait();
  if (done && *done) return;
  int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  // four independent strided chains per lane (fixed order): the loads of a row overlap
  const double* a = Ainv + (size_t)row * n;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int j = lane;
  for (; j + 96 < n; j += 128) { s0 += a[j] * b[j]; s1 += a[j + 32] * b[j + 32]; s2 += a[j + 64] * b[j + 64]; s3 += a[j + 96] * b[j + 96]; }
  for (; j < n; j += 32) s0 += a[j] * b[j];
  const double s = warp_sum((s0 + s1) + (s2 + s3));
  if (lane == 0) x[row] = s;
}

// same product for the larger dense tail (n in the thousands): one 128-thread CTA per row, every load independent
__global__ void __launch_bounds__(128) coarse_gemv_row_kernel(int n, const double* __restrict__ M, const double* __restrict__ b,
                                                              double* __restrict__ x, const int* __restrict__ done) {
  pdl_wait();
  __shared__ double s_warp[4];
  if (done && *done) return;
  const int row = blockIdx.x, t = threadIdx.x;
  const double* a = M + (size_t)row * n;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int j = t;
  for (; j + 384 < n; j += 512) { s0 += a[j] * b[j]; s1 += a[j + 128] * b[j + 128]; s2 += a[j + 256] * b[j + 256]; s3 += a[j + 384] * b[j + 384]; }
  for (; j < n; j += 128) s0 += a[j] * b[j];
  const double w = warp_sum((s0 + s1) + (s2 + s3));
  if ((t & 31) == 0) s_warp[t >> 5] = w;
  __syncthreads();
  if (t == 0) x[row] = (s_warp[0] + s_warp[1]) + (s_warp[2] + s_warp[3]);
}

// ------------------------------------------------------------------ PCG vector kernels
__global__ void cg_init_kernel(PcgScalars* sc, double tol, int maxit, int hist_cap) {
  pdl_wait();
  sc->rz_old = sc->rz_new = sc->py = sc->alpha = sc->beta = sc->rr = sc->bnorm = 0.0;
  // maxit <= 0: the reference's while (niter < maxiters) loop applies no update at all (cgcycle.cu:34)
  sc->tol = tol; sc->done = maxit <= 0 ? 1 : 0; sc->niter = 0; sc->maxit = maxit; sc->hist_len = 0; sc->hist_cap = hist_cap; sc->err = 0;
  for (int i = 0; i < 4; i++) sc->ticket[i] = 0;
}

template <int WHICH>
__global__ void __launch_bounds__(256) dot_kernel(DistDev dist, int n, const double* __restrict__ a, const double* __restrict__ b,
                                                  double* __restrict__ partials, PcgScalars* __restrict__ sc) {
  pdl_wait();
  __shared__ double s_warp[32];
  __shared__ int s_flag;
  if (sc->done) return;
  double v = 0.0, vb = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (; i + stride < n; i += 2 * stride) { v += a[i] * b[i]; vb += a[i + stride] * b[i + stride]; }
  for (; i < n; i += stride) v += a[i] * b[i];
  v += vb;
  double bs = block_sum(v, s_warp);
  if (publish_partial(bs, partials, &sc->ticket[1], &s_flag)) {
    double tot = final_sum(partials, gridDim.x, s_warp);
    tot = cross_sum(dist, tot);  // all-reduce over the GPUs of a sharded solve (no-op on one GPU)
    if (threadIdx.x == 0) {
      if (WHICH == 0) sc->bnorm = sqrt(tot);
      else if (WHICH == 1) sc->rz_old = tot;
      else { sc->rz_new = tot; sc->beta = tot / sc->rz_old; sc->rz_old = tot; }
    }
  }
}

// x += alpha p ; r -= alpha y ; ||r||^2 ; convergence test and iteration count by the last CTA
__global__ void __launch_bounds__(256) cg_update_kernel(DistDev dist, int n, double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                                                        const double* __restrict__ y, double* __restrict__ partials,
                                                        PcgScalars* __restrict__ sc, double* __restrict__ hist) {
  pdl_wait();
  __shared__ double s_warp[32];
  __shared__ int s_flag;
  if (sc->done) return;
  const double alpha = sc->alpha;
  double v = 0.0, vb = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (; i + stride < n; i += 2 * stride) {  // two independent elements in flight per thread
    const long long j = i + stride;
    const double pi = p[i], pj = p[j], yi = y[i], yj = y[j], xi = x[i], xj = x[j], ri0 = r[i], rj0 = r[j];
    x[i] = xi + alpha * pi; x[j] = xj + alpha * pj;
    const double ri = ri0 + (-alpha) * yi, rj = rj0 + (-alpha) * yj;
    r[i] = ri; r[j] = rj;
    v += ri * ri; vb += rj * rj;
  }
  for (; i < n; i += stride) {
    x[i] += alpha * p[i];
    double ri = r[i] + (-alpha) * y[i];
    r[i] = ri;
    v += ri * ri;
  }
  v += vb;
  double bs = block_sum(v, s_warp);
  if (publish_partial(bs, partials, &sc->ticket[2], &s_flag)) {
    double tot = final_sum(partials, gridDim.x, s_warp);
    tot = cross_sum(dist, tot);
    if (threadIdx.x == 0) {
      sc->rr = tot;
      double rel = sqrt(tot) / sc->bnorm;
      if (sc->hist_len < sc->hist_cap) hist[sc->hist_len++] = rel;
      if (rel <= sc->tol) sc->done = 1;                 // cgcycle.cu:47
      else { sc->niter++; if (sc->niter >= sc->maxit) sc->done = 1; }  // :35, :50
      // a peer that never arrived at an exchange (sharded solve): stop iterating, the host reports it
      if (dist.nranks > 1 && *(volatile int*)dist.error) { sc->done = 1; sc->err = 1; }
    }
  }
}

__global__ void __launch_bounds__(256) cg_pdir_kernel(int n, double* __restrict__ p, const double* __restrict__ z,
                                                      const PcgScalars* __restrict__ sc, int first) {
  pdl_wait();
  if (sc->done) return;
  const double beta = first ? 0.0 : sc->beta;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = first ? z[i] : z[i] + beta * p[i];
}

__global__ void gather_kernel(int n, const int* __restrict__ idx, const double* __restrict__ src, double* __restrict__ dst) {
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
__global__ void scatter_kernel(int n, const int* __restrict__ idx, const double* __restrict__ src, double* __restrict__ dst) {
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = src[i];
}

// grid of the BLAS-1 kernels: 8 CTAs per SM for the pure streams, 4 for the ones that end in a reduction
// (half the partial sums and tickets; measured 21.0 -> 19.0 us for cg_update, 16.7 -> 14.7 us for the dot)
inline int vec_blocks(const Ctx& c, int n, bool reduces = false) {
  return std::max(1, std::min(cdiv(n, 256), c.num_sms * (reduces ? 4 : 8)));
}

}  // namespace

void launch_spmv(const Ctx& c, const DCsr& A, const double* x, double* y, int mode, const double* b, const int* done, const char* name, RowRange rr) {
  if (mode == 0) spmv_dispatch<0, false>(c, A, x, y, b, nullptr, nullptr, done, name, rr);
  else if (mode == 1) spmv_dispatch<1, false>(c, A, x, y, b, nullptr, nullptr, done, name, rr);
  else if (mode == 2) spmv_dispatch<2, false>(c, A, x, y, b, nullptr, nullptr, done, name, rr);
  else spmv_dispatch<3, false>(c, A, x, y, b, nullptr, nullptr, done, name, rr);
}

void launch_spmv_sell(const Ctx& c, const Sell& A, const double* x, double* y, int mode, const double* b, const int* done, const char* name, RowRange rr) {
  if (mode == 0) sell_dispatch<0, false>(c, A, x, y, b, nullptr, nullptr, done, name, rr);
  else if (mode == 1) sell_dispatch<1, false>(c, A, x, y, b, nullptr, nullptr, done, name, rr);
  else if (mode == 2) sell_dispatch<2, false>(c, A, x, y, b, nullptr, nullptr, done, name, rr);
  else sell_dispatch<3, false>(c, A, x, y, b, nullptr, nullptr, done, name, rr);
}
void launch_spmv_dot_sell(const Ctx& c, const Sell& A, const double* x, double* y, double* partials, PcgScalars* sc, RowRange rr) {
  sell_dispatch<0, true>(c, A, x, y, nullptr, partials, sc, &sc->done, "spmv_dot", rr);
}
int launch_spmv_dot_sell_part(const Ctx& c, const Sell& A, const double* x, double* y, double* partials, PcgScalars* sc, RowRange rr, int pprev, bool finalize) {
  return sell_dispatch<0, true>(c, A, x, y, nullptr, partials + pprev, sc, &sc->done, "spmv_dot", rr, pprev, finalize ? 1 : 0);
}

void launch_spmv_list_add(const Ctx& c, const DCsr& A, const int* rows, int count, const double* x, double* y, const int* done) {
  if (count <= 0) return;
  g_launch_counter++;
  ProfScope ps(c, "prolong_ghost");
  FSB_LAUNCH((spmv_list_add_kernel), cdiv(count, 256), 256, 0, c.stream, count, rows, A.ptr, A.col, A.val, x, y, done);
  FSB_CHECK_LAUNCH();
}
void launch_ll_exchange(const Ctx& c, const LLXchg& x, const double* src, double* dst, const int* done) {
  g_launch_counter++;
  ProfScope ps(c, "exchange");
  int blocks = std::max(1, std::min(cdiv(std::max(x.stotal, x.rtotal), 1024), c.num_sms));
  FSB_LAUNCH((ll_exchange_kernel), blocks, 256, 0, c.stream, c.dist, x, src, dst, c.dist_ticket, done);
  FSB_CHECK_LAUNCH();
}
void launch_push_all(const Ctx& c, int begin, int end, const double* src, const PeerPtrs& dst) {
  g_launch_counter++;
  ProfScope ps(c, "push_all");
  int blocks = std::max(1, std::min(cdiv(end - begin, 512), c.num_sms * 2));
  FSB_LAUNCH((push_all_kernel), blocks, 256, 0, c.stream, c.dist, begin, end, src, dst, c.dist_ticket);
  FSB_CHECK_LAUNCH();
}

void launch_spmv_dot(const Ctx& c, const DCsr& A, const double* x, double* y, double* partials, PcgScalars* sc) {
  spmv_dispatch<0, true>(c, A, x, y, nullptr, partials, sc, &sc->done, "spmv_dot");
}

void launch_smooth(const Ctx& c, const LevelData& L, const double* b_src, const int* gather, double* b_int, const double* x_in,
                   double w, int nsweeps, double* x_out, const int* scatter, double* x_ext, double* r_out, const int* done, bool owned_only) {
  g_launch_counter++;
  ProfScope ps(c, x_in ? "post_smooth" : "pre_smooth");
  cudaStream_t s = c.stream;
  // sharded level: this GPU smooths its own contiguous range of partitions only
  const int p0 = owned_only ? L.ownP0 : 0, pn = owned_only ? L.ownPn : L.nparts;
  if (pn <= 0) return;
  if (L.use_blockdense && !r_out && !owned_only) {  // few small partitions: the whole stage as dense GEMVs on precomputed blocks
    FSB_LAUNCH((smooth_blockdense_kernel<32>), L.bdCtas, 256, 0, s, L.bdWork, L.pstart, L.bdOff, L.bdS1, L.bdGp, L.bdS2, b_src, gather, b_int, x_in,
                                                x_out, scatter, x_ext, done);
    FSB_CHECK_LAUNCH();
    return;
  }
  const int* nl = owned_only ? L.nlistOwn : L.nlist;
  const DevBuf<EllDesc>* pl = owned_only ? L.plistOwn : L.plist;
  static const bool serial = getenv("FSB_ELL_SERIAL") && atoi(getenv("FSB_ELL_SERIAL")) != 0;  // tuning knob: size classes one after the other
#define FSB_ELL_TAIL L.ellval, L.ellcol, L.ellrow, L.diag, b_src, gather, b_int, x_in, w, nsweeps, x_out, scatter, x_ext, r_out, done
// the size classes run side by side (fork/join on two helper streams, also inside a captured graph):
// the CTAs of the small classes fill the tail of the big one
#define FSB_ELL_LAUNCH1(MK, BL, q)                                                                                   \
  do {                                                                                                             \
    cudaStream_t sq = s;                                                                                           \
    if (launched > 0 && c.side[launched - 1] && !serial) {                                                                  \
      sq = c.side[launched - 1];                                                                                   \
      FSB_CUDA(cudaStreamWaitEvent(sq, c.ev_fork, 0));                                                             \
    }                                                                                                              \
    FSB_LAUNCH((smooth_ell_kernel<MK, BL>), nl[q], (q == 2 && BL == 1024) ? ((L.maxPartRows + 31) & ~31) : BL, 0, sq, pl[q].get(), FSB_ELL_TAIL); \
    if (sq != s) { FSB_CUDA(cudaEventRecord(c.ev_join[launched - 1], sq)); FSB_CUDA(cudaStreamWaitEvent(s, c.ev_join[launched - 1], 0)); } \
    launched++;                                                                                                    \
  } while (0)
#define FSB_ELL_LAUNCH(MK)                                                                                          \
  do {                                                                                                             \
    int launched = 0;                                                                                              \
    if ((nl[0] > 0) + (nl[1] > 0) + (nl[2] > 0) > 1 && c.side[0] && !serial) FSB_CUDA(cudaEventRecord(c.ev_fork, s)); \
    if (nl[0] > 0) FSB_ELL_LAUNCH1(MK, 256, 0);                                                                     \
    if (nl[1] > 0) FSB_ELL_LAUNCH1(MK, 384, 1);                                                                     \
    if (nl[2] > 0 && L.maxPartRows <= 512) FSB_ELL_LAUNCH1(MK, 512, 2);                                             \
    else if (nl[2] > 0) FSB_ELL_LAUNCH1(MK, 1024, 2);                                                               \
    if (launched > 1) g_launch_counter += launched - 1;                                                            \
  } while (0)
  if (L.use_ell && L.ellMaxK <= 8) FSB_ELL_LAUNCH(8);
  else if (L.use_ell && L.ellMaxK <= 16) FSB_ELL_LAUNCH(16);
  else if (L.use_ell && L.ellMaxK <= 24 && L.maxPartRows <= 512) FSB_ELL_LAUNCH(24);
  else if (L.use_ell && L.ellMaxK <= 32 && L.maxPartRows <= 512) FSB_ELL_LAUNCH(32);
#undef FSB_ELL_LAUNCH
#undef FSB_ELL_LAUNCH1
#undef FSB_ELL_TAIL
  else if (L.use_sellg) {
    static PerDeviceOnce attr_once;
    if (attr_once.first(c.device)) {
      FSB_CUDA(cudaFuncSetAttribute(smooth_sellg_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      FSB_CUDA(cudaFuncSetAttribute(smooth_sellg_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    // more partitions than SMs: two 512-thread CTAs per SM, each with half of the shared memory; else one 1024-thread CTA
    static const char* env_two = getenv("FSB_SELLG_TWO");  // tuning knob
    const bool two = env_two ? atoi(env_two) != 0 : pn > c.num_sms;
    const int npad = (L.maxPartRows + 15) & ~15;
    const size_t fixed = (size_t)(2 * (2 * npad + 8) + 3 * npad) * 8 + 16;
    const size_t limit = (two ? 113 : 226) * 1024 - 1024;
    int smemSlots = (int)std::min<long long>(L.sellgMaxSlots, (long long)((limit - fixed) / 10) & ~31LL);
    if (smemSlots < 0) smemSlots = 0;
    const size_t smem = fixed + (size_t)smemSlots * 10;
#define FSB_SELLG_ARGS L.sellgDesc.get() + p0, L.ellG, npad, smemSlots, L.ellwptr, L.ellval, L.ellcol, L.ellrow, L.diag, b_src, gather, b_int, x_in, w, nsweeps, x_out, scatter, x_ext, r_out, done
    if (two) FSB_LAUNCH((smooth_sellg_kernel<512>), pn, 512, smem, s, FSB_SELLG_ARGS);
    else FSB_LAUNCH((smooth_sellg_kernel<1024>), pn, 1024, smem, s, FSB_SELLG_ARGS);
#undef FSB_SELLG_ARGS
  } else if (L.smemBytes > 0) {
    static PerDeviceOnce attr_once;
    if (attr_once.first(c.device))
      FSB_CUDA(cudaFuncSetAttribute(smooth_cluster_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pn * L.clusterC); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = L.smemBytes; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = L.clusterC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_config(dim3(1), dim3(1), 0, s).numAttrs ? 2 : 1;
    FSB_CUDA(cudaLaunchKernelEx(&cfg, smooth_cluster_kernel<512>, L.clusterC, L.coopG, L.maxChunkNnz, L.maxPartRows, L.maxChunkRows,
                                (const int*)L.pstart.get() + p0, (const int*)L.A.ptr.get(), (const int*)L.A.col.get(), (const double*)L.A.val.get(),
                                (const double*)L.diag.get(), b_src, gather, b_int, x_in, w, nsweeps, x_out, scatter, x_ext, r_out, done));
  } else
    FSB_LAUNCH((smooth_coop_kernel<512>), pn, 512, 0, s, L.coopG, L.pstart.get() + p0, L.A.ptr, L.A.col, L.A.val, L.diag, b_src, gather, b_int, x_in, w,
                                                    nsweeps, x_out, scatter, x_ext, r_out, done);
  FSB_CHECK_LAUNCH();
}

void debug_stamps(int cta_plus1, long long* out64) {
  if (out64) FSB_CUDA(cudaMemcpyFromSymbol(out64, g_dbg, sizeof(long long) * 64));
  FSB_CUDA(cudaMemcpyToSymbol(g_dbg_on, &cta_plus1, sizeof(int)));
}

void launch_restrict_fused(const Ctx& c, const DCsr& R, const double* b, const DCsr& T, const double* x, double* bc, const int* done) {
  g_launch_counter++;
  ProfScope ps(c, "restrict");
  if (R.nrows <= 16384 && (long long)T.nnz >= 128LL * R.nrows)  // few, long rows: a CTA per row
    FSB_LAUNCH((restrict_fused_row_kernel), R.nrows, 128, 0, c.stream, R.nrows, R.ptr, R.col, R.val, b, T.ptr, T.col, T.val, x, bc, done);
  else
    FSB_LAUNCH((restrict_fused_kernel), cdiv((long long)R.nrows * 32, 256), 256, 0, c.stream, R.nrows, R.ptr, R.col, R.val, b, T.ptr, T.col, T.val, x, bc, done);
  FSB_CHECK_LAUNCH();
}

void launch_coarse_solve(const Ctx& c, int n, const double* Ainv, const double* b, double* x, const int* done) {
  g_launch_counter++;
  ProfScope ps(c, "coarse_solve");
  if (n >= 512) FSB_LAUNCH((coarse_gemv_row_kernel), n, 128, 0, c.stream, n, Ainv, b, x, done);
  else FSB_LAUNCH((coarse_gemv_kernel), cdiv(n, 8), 256, 0, c.stream, n, Ainv, b, x, done);
  FSB_CHECK_LAUNCH();
}

void launch_cg_init(const Ctx& c, PcgScalars* sc, double tol, int maxit, int hist_cap) {
  FSB_LAUNCH((cg_init_kernel), 1, 1, 0, c.stream, sc, tol, maxit, hist_cap);
  FSB_CHECK_LAUNCH();
}

void launch_dot(const Ctx& c, int n, const double* a, const double* b, double* partials, PcgScalars* sc, int which) {
  g_launch_counter++;
  ProfScope ps(c, "dot");
  int blocks = vec_blocks(c, n, true);
  if (which == 0) FSB_LAUNCH((dot_kernel<0>), blocks, 256, 0, c.stream, c.dist, n, a, b, partials, sc);
  else if (which == 1) FSB_LAUNCH((dot_kernel<1>), blocks, 256, 0, c.stream, c.dist, n, a, b, partials, sc);
  else FSB_LAUNCH((dot_kernel<2>), blocks, 256, 0, c.stream, c.dist, n, a, b, partials, sc);
  FSB_CHECK_LAUNCH();
}

void launch_cg_update(const Ctx& c, int n, double* x, double* r, const double* p, const double* y, double* partials, PcgScalars* sc, double* hist) {
  g_launch_counter++;
  ProfScope ps(c, "cg_update");
  FSB_LAUNCH((cg_update_kernel), vec_blocks(c, n, true), 256, 0, c.stream, c.dist, n, x, r, p, y, partials, sc, hist);
  FSB_CHECK_LAUNCH();
}

void launch_cg_pdir(const Ctx& c, int n, double* p, const double* z, const PcgScalars* sc, int first) {
  g_launch_counter++;
  ProfScope ps(c, "cg_pdir");
  FSB_LAUNCH((cg_pdir_kernel), vec_blocks(c, n), 256, 0, c.stream, n, p, z, sc, first);
  FSB_CHECK_LAUNCH();
}

void launch_gather(const Ctx& c, int n, const int* idx, const double* src, double* dst) {
  g_launch_counter++;
  FSB_LAUNCH((gather_kernel), cdiv(n, 256), 256, 0, c.stream, n, idx, src, dst);
  FSB_CHECK_LAUNCH();
}
void launch_scatter(const Ctx& c, int n, const int* idx, const double* src, double* dst) {
  g_launch_counter++;
  FSB_LAUNCH((scatter_kernel), cdiv(n, 256), 256, 0, c.stream, n, idx, src, dst);
  FSB_CHECK_LAUNCH();
}

}  // namespace fsb
