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
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += __ldcg(partials + i);
  __syncthreads();
  return block_sum(v, s_warp);
}

// ------------------------------------------------------------------ SpMV family
// G lanes cooperate on one row: coalesced (col,val) loads, shuffle reduction.
template <int G, int MODE, bool DOT>
__global__ void __launch_bounds__(256) spmv_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ col,
                                                   const double* __restrict__ val, const double* __restrict__ x,
                                                   double* __restrict__ y, const double* __restrict__ b,
                                                   double* __restrict__ partials, PcgScalars* __restrict__ sc, const int* __restrict__ done) {
  __shared__ double s_warp[32];
  __shared__ int s_flag;
  if (done && *done) return;
  const int lane = threadIdx.x % G;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G;
  double s = 0.0;
  if (row < n) {
    int e1 = ptr[row + 1];
    for (int e = ptr[row] + lane; e < e1; e += G) s += val[e] * __ldg(x + col[e]);
  }
#pragma unroll
  for (int o = G >> 1; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, G);
  double contrib = 0.0;
  if (row < n && lane == 0) {
    if (MODE == 0) y[row] = s;
    else if (MODE == 1) y[row] = b[row] - s;
    else y[row] = y[row] + s;
    if (DOT) contrib = x[row] * s;
  }
  if (DOT) {
    double bs = block_sum(contrib, s_warp);
    if (publish_partial(bs, partials, &sc->ticket[0], &s_flag)) {
      double tot = final_sum(partials, gridDim.x, s_warp);
      if (threadIdx.x == 0) { sc->py = tot; sc->alpha = sc->rz_old / tot; }
    }
  }
}

template <int MODE, bool DOT>
void spmv_dispatch(const Ctx& c, const DCsr& A, const double* x, double* y, const double* b, double* partials, PcgScalars* sc, const int* done) {
  int n = A.nrows;
  if (n == 0) return;
  double avg = (double)A.nnz / n;
  int G = avg > 24 ? 32 : avg > 12 ? 16 : avg > 6 ? 8 : 4;
  int blocks = cdiv((long long)n * G, 256);
  g_launch_counter++;
  switch (G) {
    case 32: spmv_kernel<32, MODE, DOT><<<blocks, 256, 0, c.stream>>>(n, A.ptr, A.col, A.val, x, y, b, partials, sc, done); break;
    case 16: spmv_kernel<16, MODE, DOT><<<blocks, 256, 0, c.stream>>>(n, A.ptr, A.col, A.val, x, y, b, partials, sc, done); break;
    case 8: spmv_kernel<8, MODE, DOT><<<blocks, 256, 0, c.stream>>>(n, A.ptr, A.col, A.val, x, y, b, partials, sc, done); break;
    default: spmv_kernel<4, MODE, DOT><<<blocks, 256, 0, c.stream>>>(n, A.ptr, A.col, A.val, x, y, b, partials, sc, done); break;
  }
  FSB_CHECK_LAUNCH();
}

// ------------------------------------------------------------------ fused smoother
// One CTA per partition, thread t <-> row pstart+t, x double-buffered in shared memory.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) pre_smooth_kernel(const int* __restrict__ pstart, const int* __restrict__ ptr,
                                                           const int* __restrict__ col, const double* __restrict__ val,
                                                           const double* __restrict__ diag, const double* __restrict__ b_src,
                                                           const int* __restrict__ gather, double* __restrict__ b_int, double w,
                                                           int nsweeps, double* __restrict__ x, const int* __restrict__ done) {
  __shared__ double sx[2][BLOCK];
  if (done && *done) return;
  const int r0 = pstart[blockIdx.x], np = pstart[blockIdx.x + 1] - r0, t = threadIdx.x, row = r0 + t;
  const bool active = t < np;
  double b = 0.0, d = 1.0;
  int e0 = 0, e1 = 0;
  if (active) {
    b = b_src[gather ? gather[row] : row];
    d = diag[row];
    if (b_int) b_int[row] = b;
    e0 = ptr[row]; e1 = ptr[row + 1];
    sx[0][t] = w * b / d;
  }
  __syncthreads();
  int cur = 0;
  for (int it = 0; it < nsweeps; it++) {
    if (active) {
      double s = 0.0;
      for (int e = e0; e < e1; e++) {
        int cc = col[e] - r0;
        if ((unsigned)cc < (unsigned)np && cc != t) s += val[e] * sx[cur][cc];
      }
      double xv = sx[cur][t];
      sx[cur ^ 1][t] = xv + w * (b - s - d * xv) / d;
    }
    __syncthreads();
    cur ^= 1;
  }
  if (active) x[row] = sx[cur][t];
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) post_smooth_kernel(const int* __restrict__ pstart, const int* __restrict__ ptr,
                                                            const int* __restrict__ col, const double* __restrict__ val,
                                                            const double* __restrict__ diag, const double* __restrict__ b_int,
                                                            const double* __restrict__ x_in, double w, int nsweeps,
                                                            double* __restrict__ x_out, const int* __restrict__ scatter,
                                                            double* __restrict__ x_ext, const int* __restrict__ done) {
  __shared__ double sx[2][BLOCK];
  if (done && *done) return;
  const int r0 = pstart[blockIdx.x], np = pstart[blockIdx.x + 1] - r0, t = threadIdx.x, row = r0 + t;
  const bool active = t < np;
  double b = 0.0, d = 1.0;
  int e0 = 0, e1 = 0;
  if (active) {
    b = b_int[row];
    d = diag[row];
    e0 = ptr[row]; e1 = ptr[row + 1];
    // b' = b - A_out x : neighbours' x frozen at the value they had when the pass started
    for (int e = e0; e < e1; e++) {
      int cg = col[e];
      if ((unsigned)(cg - r0) >= (unsigned)np) b -= val[e] * __ldg(x_in + cg);
    }
    sx[0][t] = x_in[row];
  }
  __syncthreads();
  int cur = 0;
  for (int it = 0; it < nsweeps; it++) {
    if (active) {
      double s = 0.0;
      for (int e = e0; e < e1; e++) {
        int cc = col[e] - r0;
        if ((unsigned)cc < (unsigned)np && cc != t) s += val[e] * sx[cur][cc];
      }
      double xv = sx[cur][t];
      sx[cur ^ 1][t] = xv + w * (b - s - d * xv) / d;
    }
    __syncthreads();
    cur ^= 1;
  }
  if (active) {
    double xv = sx[cur][t];
    if (x_out) x_out[row] = xv;
    if (scatter) x_ext[scatter[row]] = xv;
  }
}

// x = Ainv b, dense, n < topSize_ (replaces the host LU round trip of amg_level.cu:25-31)
__global__ void __launch_bounds__(256) coarse_gemv_kernel(int n, const double* __restrict__ Ainv, const double* __restrict__ b,
                                                          double* __restrict__ x, const int* __restrict__ done) {
  if (done && *done) return;
  int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  double s = 0.0;
  for (int j = lane; j < n; j += 32) s += Ainv[(size_t)row * n + j] * b[j];
  s = warp_sum(s);
  if (lane == 0) x[row] = s;
}

// ------------------------------------------------------------------ PCG vector kernels
__global__ void cg_init_kernel(PcgScalars* sc, double tol, int maxit) {
  sc->rz_old = sc->rz_new = sc->py = sc->alpha = sc->beta = sc->rr = sc->bnorm = 0.0;
  sc->tol = tol; sc->done = 0; sc->niter = 0; sc->maxit = maxit; sc->hist_len = 0;
  for (int i = 0; i < 4; i++) sc->ticket[i] = 0;
}

template <int WHICH>
__global__ void __launch_bounds__(256) dot_kernel(int n, const double* __restrict__ a, const double* __restrict__ b,
                                                  double* __restrict__ partials, PcgScalars* __restrict__ sc) {
  __shared__ double s_warp[32];
  __shared__ int s_flag;
  if (sc->done) return;
  double v = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) v += a[i] * b[i];
  double bs = block_sum(v, s_warp);
  if (publish_partial(bs, partials, &sc->ticket[1], &s_flag)) {
    double tot = final_sum(partials, gridDim.x, s_warp);
    if (threadIdx.x == 0) {
      if (WHICH == 0) sc->bnorm = sqrt(tot);
      else if (WHICH == 1) sc->rz_old = tot;
      else { sc->rz_new = tot; sc->beta = tot / sc->rz_old; sc->rz_old = tot; }
    }
  }
}

// x += alpha p ; r -= alpha y ; ||r||^2 ; convergence test and iteration count by the last CTA
__global__ void __launch_bounds__(256) cg_update_kernel(int n, double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                                                        const double* __restrict__ y, double* __restrict__ partials,
                                                        PcgScalars* __restrict__ sc, double* __restrict__ hist) {
  __shared__ double s_warp[32];
  __shared__ int s_flag;
  if (sc->done) return;
  const double alpha = sc->alpha;
  double v = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    double ri = r[i] + (-alpha) * y[i];
    r[i] = ri;
    v += ri * ri;
  }
  double bs = block_sum(v, s_warp);
  if (publish_partial(bs, partials, &sc->ticket[2], &s_flag)) {
    double tot = final_sum(partials, gridDim.x, s_warp);
    if (threadIdx.x == 0) {
      sc->rr = tot;
      double rel = sqrt(tot) / sc->bnorm;
      hist[sc->hist_len++] = rel;
      if (rel <= sc->tol) sc->done = 1;                 // cgcycle.cu:47
      else { sc->niter++; if (sc->niter >= sc->maxit) sc->done = 1; }  // :35, :50
    }
  }
}

__global__ void __launch_bounds__(256) cg_pdir_kernel(int n, double* __restrict__ p, const double* __restrict__ z,
                                                      const PcgScalars* __restrict__ sc, int first) {
  if (sc->done) return;
  const double beta = first ? 0.0 : sc->beta;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = first ? z[i] : z[i] + beta * p[i];
}

__global__ void gather_kernel(int n, const int* __restrict__ idx, const double* __restrict__ src, double* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
__global__ void scatter_kernel(int n, const int* __restrict__ idx, const double* __restrict__ src, double* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = src[i];
}

inline int vec_blocks(const Ctx& c, int n) { return std::max(1, std::min(cdiv(n, 256), c.num_sms * 8)); }

}  // namespace

void launch_spmv(const Ctx& c, const DCsr& A, const double* x, double* y, int mode, const double* b, const int* done) {
  if (mode == 0) spmv_dispatch<0, false>(c, A, x, y, b, nullptr, nullptr, done);
  else if (mode == 1) spmv_dispatch<1, false>(c, A, x, y, b, nullptr, nullptr, done);
  else spmv_dispatch<2, false>(c, A, x, y, b, nullptr, nullptr, done);
}

void launch_spmv_dot(const Ctx& c, const DCsr& A, const double* x, double* y, double* partials, PcgScalars* sc) {
  spmv_dispatch<0, true>(c, A, x, y, nullptr, partials, sc, &sc->done);
}

void launch_pre_smooth(const Ctx& c, const LevelData& L, const double* b_src, const int* gather, double* b_int, double w,
                       int nsweeps, double* x, const int* done) {
  g_launch_counter++;
  if (L.maxPartRows <= 256)
    pre_smooth_kernel<256><<<L.nparts, 256, 0, c.stream>>>(L.pstart, L.A.ptr, L.A.col, L.A.val, L.diag, b_src, gather, b_int, w, nsweeps, x, done);
  else if (L.maxPartRows <= 512)
    pre_smooth_kernel<512><<<L.nparts, 512, 0, c.stream>>>(L.pstart, L.A.ptr, L.A.col, L.A.val, L.diag, b_src, gather, b_int, w, nsweeps, x, done);
  else
    pre_smooth_kernel<1024><<<L.nparts, 1024, 0, c.stream>>>(L.pstart, L.A.ptr, L.A.col, L.A.val, L.diag, b_src, gather, b_int, w, nsweeps, x, done);
  FSB_CHECK_LAUNCH();
}

void launch_post_smooth(const Ctx& c, const LevelData& L, const double* b_int, const double* x_in, double w, int nsweeps,
                        double* x_out, const int* scatter, double* x_ext, const int* done) {
  g_launch_counter++;
  if (L.maxPartRows <= 256)
    post_smooth_kernel<256><<<L.nparts, 256, 0, c.stream>>>(L.pstart, L.A.ptr, L.A.col, L.A.val, L.diag, b_int, x_in, w, nsweeps, x_out, scatter, x_ext, done);
  else if (L.maxPartRows <= 512)
    post_smooth_kernel<512><<<L.nparts, 512, 0, c.stream>>>(L.pstart, L.A.ptr, L.A.col, L.A.val, L.diag, b_int, x_in, w, nsweeps, x_out, scatter, x_ext, done);
  else
    post_smooth_kernel<1024><<<L.nparts, 1024, 0, c.stream>>>(L.pstart, L.A.ptr, L.A.col, L.A.val, L.diag, b_int, x_in, w, nsweeps, x_out, scatter, x_ext, done);
  FSB_CHECK_LAUNCH();
}

void launch_coarse_solve(const Ctx& c, int n, const double* Ainv, const double* b, double* x, const int* done) {
  g_launch_counter++;
  coarse_gemv_kernel<<<cdiv(n, 8), 256, 0, c.stream>>>(n, Ainv, b, x, done);
  FSB_CHECK_LAUNCH();
}

void launch_cg_init(const Ctx& c, PcgScalars* sc, double tol, int maxit) {
  cg_init_kernel<<<1, 1, 0, c.stream>>>(sc, tol, maxit);
  FSB_CHECK_LAUNCH();
}

void launch_dot(const Ctx& c, int n, const double* a, const double* b, double* partials, PcgScalars* sc, int which) {
  g_launch_counter++;
  int blocks = vec_blocks(c, n);
  if (which == 0) dot_kernel<0><<<blocks, 256, 0, c.stream>>>(n, a, b, partials, sc);
  else if (which == 1) dot_kernel<1><<<blocks, 256, 0, c.stream>>>(n, a, b, partials, sc);
  else dot_kernel<2><<<blocks, 256, 0, c.stream>>>(n, a, b, partials, sc);
  FSB_CHECK_LAUNCH();
}

void launch_cg_update(const Ctx& c, int n, double* x, double* r, const double* p, const double* y, double* partials, PcgScalars* sc, double* hist) {
  g_launch_counter++;
  cg_update_kernel<<<vec_blocks(c, n), 256, 0, c.stream>>>(n, x, r, p, y, partials, sc, hist);
  FSB_CHECK_LAUNCH();
}

void launch_cg_pdir(const Ctx& c, int n, double* p, const double* z, const PcgScalars* sc, int first) {
  g_launch_counter++;
  cg_pdir_kernel<<<vec_blocks(c, n), 256, 0, c.stream>>>(n, p, z, sc, first);
  FSB_CHECK_LAUNCH();
}

void launch_gather(const Ctx& c, int n, const int* idx, const double* src, double* dst) {
  g_launch_counter++;
  gather_kernel<<<cdiv(n, 256), 256, 0, c.stream>>>(n, idx, src, dst);
  FSB_CHECK_LAUNCH();
}
void launch_scatter(const Ctx& c, int n, const int* idx, const double* src, double* dst) {
  g_launch_counter++;
  scatter_kernel<<<cdiv(n, 256), 256, 0, c.stream>>>(n, idx, src, dst);
  FSB_CHECK_LAUNCH();
}

}  // namespace fsb
