// Dense tail of the V-cycle.
//
// The levels below a few thousand rows are pure latency: seven kernels per level for a few
// microseconds of work each.  One V-cycle on such a level is a fixed LINEAR map  x = M_l b  (Appendix B
// of SURVEY.md: Jacobi-type sweeps, restriction, recursion, prolongation, post-relaxation are all
// linear, and the initial guess of every cycle is zero), so for the levels l >= L_t with n_l <= kMaxRows the
// map is formed once at setup as a dense matrix and the whole sub-cycle becomes ONE dense GEMV per
// PCG iteration (launch_coarse_solve), exactly like the coarsest level's dense inverse already is.
//
// With  A = B + A_out  (B = intra-partition block incl. the diagonal),  W = w D^-1,  G = I - W B  (block
// diagonal),  S(k) = sum_{j<=k} G^j W,  nu1/nu2 inner sweeps and mu post-relaxation passes
// (gauss_seidel.cu:1312-1375, 3664-3735, 4425-4427):
//     pre        x = S(nu1) b                      r = (I - A S(nu1)) b
//     coarse     X = S(nu1) + P M_{l+1} R (I - A S(nu1))
//     post (mu x)  X <- (G^nu2 - S(nu2-1) A_out) X + S(nu2-1)
// in the level's permuted numbering; M_l is that X carried to the level's external numbering.  The
// block-diagonal factors are multiplied partition by partition (small cuBLAS DGEMMs on sub-blocks),
// the only full n^3 product is the post-relaxation update.  Same arithmetic as the kernels up to the
// order of the floating-point sums (differences ~1e-16 relative; iteration counts unchanged).
#include <cublas_v2.h>

#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "solver.h"

namespace fsb {

namespace {

#define FSB_CUBLAS(call)                                                                               \
  do {                                                                                                 \
    cublasStatus_t st_ = (call);                                                                       \
    if (st_ != CUBLAS_STATUS_SUCCESS) throw std::runtime_error("cuBLAS error " + std::to_string((int)st_) + " in " #call); \
  } while (0)

// C(m x n) = alpha A(m x k) B(k x n) + beta C, everything row-major with row strides lda/ldb/ldc
void gemm_rm(cublasHandle_t h, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
             double* C, int ldc) {
  if (m == 0 || n == 0) return;
  if (k == 0) { alpha = 0.0; k = 1; }
  FSB_CUBLAS(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, m, k, &alpha, B, ldb, A, lda, &beta, C, ldc));
}

__global__ void dt_row_partition(int nparts, const int* __restrict__ pstart, int* __restrict__ rowPart) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nparts) return;
  for (int r = pstart[p]; r < pstart[p + 1]; r++) rowPart[r] = p;
}
// A (CSR, permuted numbering) -> dense B (intra-partition entries incl. the diagonal) and dense A_out
__global__ void dt_split_dense(int n, const int* __restrict__ ptr, const int* __restrict__ col, const double* __restrict__ val,
                               const int* __restrict__ rowPart, double* __restrict__ Bd, double* __restrict__ Ao) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int p = rowPart[r];
  for (int e = ptr[r]; e < ptr[r + 1]; e++) {
    const int c = col[e];
    (rowPart[c] == p ? Bd : Ao)[(size_t)r * n + c] = val[e];
  }
}
__global__ void dt_densify(int nrows, int ncols, const int* __restrict__ ptr, const int* __restrict__ col, const double* __restrict__ val,
                           double* __restrict__ M) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  for (int e = ptr[r]; e < ptr[r + 1]; e++) M[(size_t)r * ncols + col[e]] = val[e];
}
// G = I - diag(W) B ;  S0 = diag(W)
__global__ void dt_make_G(int n, const double* __restrict__ diag, double w, const double* __restrict__ Bd, double* __restrict__ G,
                          double* __restrict__ S0) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * n) return;
  const int r = (int)(i / n), c = (int)(i % n);
  const double wr = w / diag[r];
  G[i] = (r == c ? 1.0 : 0.0) - wr * Bd[i];
  S0[i] = (r == c) ? wr : 0.0;
}
__global__ void dt_add_diag(int n, const double* __restrict__ diag, double w, double* __restrict__ M) {  // M += diag(w / d)
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) M[(size_t)r * n + r] += w / diag[r];
}
__global__ void dt_add_identity(int n, double* __restrict__ M) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) M[(size_t)r * n + r] += 1.0;
}
__global__ void dt_add(size_t count, const double* __restrict__ a, double* __restrict__ b) {  // b += a
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) b[i] += a[i];
}
// Mext[ip[i]][ip[j]] = Mint[i][j]
__global__ void dt_to_external(int n, const int* __restrict__ ip, const double* __restrict__ Mint, double* __restrict__ Mext) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * n) return;
  const int r = (int)(i / n), c = (int)(i % n);
  Mext[(size_t)ip[r] * n + ip[c]] = Mint[i];
}

}  // namespace

// operator of one V-cycle on level l (external numbering), given the operator of level l+1
static void level_operator(const Ctx& ctx, cublasHandle_t h, const LevelData& L, const Params& prm, const DBuf& Mnext, DBuf& Mext) {
  cudaStream_t s = ctx.stream;
  const int n = L.n, nc = L.nnout, np = L.nparts;
  const size_t nn = (size_t)n * n;
  const double w = prm.smootherWeight;
  const int nu1 = std::max(0, prm.preInnerIters), nu2 = std::max(0, prm.postInnerIters), mu = std::max(0, prm.postRelaxes);
  const int gb = (int)cdiv((long long)nn, 256);
  std::vector<int> ps = L.pstart.to_vector();
  IBuf rowPart(n, s);
  dt_row_partition<<<cdiv(np, 128), 128, 0, s>>>(np, L.pstart, rowPart);
  DBuf Bd(nn, s), Ao(nn, s), G(nn, s), Sa(nn, s), Sb(nn, s), S1(nn, s), S2(nn, s), Gp(nn, s), Gq(nn, s);
  Bd.zero(); Ao.zero(); Sb.zero(); S2.zero(); Gq.zero();
  dt_split_dense<<<cdiv(n, 128), 128, 0, s>>>(n, L.A.ptr, L.A.col, L.A.val, rowPart, Bd, Ao);
  dt_make_G<<<gb, 256, 0, s>>>(n, L.diag, w, Bd, G, Sa);
  // Horner on the diagonal blocks: S(k+1) = W + G S(k), S(0) = W.  S1 = S(nu1); S2 = S(nu2 - 1) (zero if nu2 == 0)
  auto block_mul = [&](const double* A, const double* B, double* C) {  // C_blk = A_blk B_blk for every partition
    for (int p = 0; p < np; p++) {
      const int r0 = ps[p], m = ps[p + 1] - r0;
      const size_t off = (size_t)r0 * n + r0;
      gemm_rm(h, m, m, m, 1.0, A + off, n, B + off, n, 0.0, C + off, n);
    }
  };
  double* cur = Sa;
  double* nxt = Sb;
  const int steps = std::max(nu1, nu2 - 1);
  for (int k = 0; k <= steps; k++) {
    if (k == nu1) S1.from_device(cur, nn);
    if (k == nu2 - 1) S2.from_device(cur, nn);
    if (k == steps) break;
    block_mul(G, cur, nxt);
    dt_add_diag<<<cdiv(n, 128), 128, 0, s>>>(n, L.diag, w, nxt);
    std::swap(cur, nxt);
  }
  // Gp = G^nu2 (identity for nu2 == 0)
  {
    Gp.zero();
    dt_add_identity<<<cdiv(n, 128), 128, 0, s>>>(n, Gp);
    double* a = Gp;
    double* b = Gq;
    for (int k = 0; k < nu2; k++) { block_mul(G, a, b); std::swap(a, b); }
    if (a != Gp.get()) Gp.from_device(a, nn);
  }
  // T = I - A S1 (column block by column block: S1 is block diagonal), A = B + A_out
  DBuf& Ad = Bd;  // B is not needed any more
  dt_add<<<gb, 256, 0, s>>>(nn, Ao, Ad);
  DBuf& T = G;    // nor is G
  T.zero();
  for (int p = 0; p < np; p++) {
    const int r0 = ps[p], m = ps[p + 1] - r0;
    gemm_rm(h, n, m, m, -1.0, Ad.get() + r0, n, S1.get() + (size_t)r0 * n + r0, n, 0.0, T.get() + r0, n);
  }
  dt_add_identity<<<cdiv(n, 128), 128, 0, s>>>(n, T);
  // X = S1 + P (M_{l+1} (R T))
  DBuf Pd((size_t)n * nc, s), Rd((size_t)nc * n, s), Q((size_t)nc * n, s), Q2((size_t)nc * n, s);
  Pd.zero(); Rd.zero();
  dt_densify<<<cdiv(n, 128), 128, 0, s>>>(n, nc, L.P.ptr, L.P.col, L.P.val, Pd);
  dt_densify<<<cdiv(nc, 128), 128, 0, s>>>(nc, n, L.R.ptr, L.R.col, L.R.val, Rd);
  gemm_rm(h, nc, n, n, 1.0, Rd, n, T, n, 0.0, Q, n);
  gemm_rm(h, nc, n, nc, 1.0, Mnext, nc, Q, n, 0.0, Q2, n);
  DBuf& X = Sa;
  X.from_device(S1, nn);
  gemm_rm(h, n, n, nc, 1.0, Pd, nc, Q2, n, 1.0, X, n);
  // post-relaxation passes: X <- (G^nu2 - S2 A_out) X + S2
  if (mu > 0) {
    DBuf& H = Sb;
    H.zero();
    for (int p = 0; p < np; p++) {  // row block by row block: S2 is block diagonal
      const int r0 = ps[p], m = ps[p + 1] - r0;
      gemm_rm(h, m, n, m, -1.0, S2.get() + (size_t)r0 * n + r0, n, Ao.get() + (size_t)r0 * n, n, 0.0, H.get() + (size_t)r0 * n, n);
    }
    dt_add<<<gb, 256, 0, s>>>(nn, Gp, H);
    double* xin = X;
    double* xout = Gq;
    for (int pass = 0; pass < mu; pass++) {
      FSB_CUDA(cudaMemcpyAsync(xout, S2.get(), sizeof(double) * nn, cudaMemcpyDeviceToDevice, s));
      gemm_rm(h, n, n, n, 1.0, H, n, xin, n, 1.0, xout, n);
      std::swap(xin, xout);
    }
    if (xin != X.get()) X.from_device(xin, nn);
  }
  Mext.alloc(nn, s);
  dt_to_external<<<gb, 256, 0, s>>>(n, L.agg.ipermutation, X, Mext);
  FSB_CHECK_LAUNCH();
  FSB_CUDA(cudaStreamSynchronize(s));  // the temporaries go out of scope
}

void Solver::build_dense_tail() {
  tail_level_ = -1;
  Mtail.release();
  tail_key_[0] = prm.preInnerIters; tail_key_[1] = prm.postInnerIters; tail_key_[2] = prm.postRelaxes; tail_key_[3] = prm.smootherWeight;
  const char* env = getenv("FSB_DENSE_TAIL");  // tuning / test knob, read at every setup: 0 disables, n > 1 sets the row limit
  const int limit = env ? atoi(env) : kDenseTailMaxRows;
  const int last = (int)levels.size() - 1;
  if (limit <= 1 || last < 2) return;
  int lt = -1;
  for (int l = 1; l < last; l++)
    if (levels[l].n <= limit) { lt = l; break; }
  if (lt < 0) return;
  // The tail is an optimisation of levels that the smoother / transfer kernels can also run: if cuBLAS is not
  // usable (library missing, out of memory for the n x n temporaries) the solve keeps those kernels.
  try {
    if (!cublas_) {
      cublasHandle_t h;
      FSB_CUBLAS(cublasCreate(&h));
      cublas_ = h;
    }
    cublasHandle_t h = static_cast<cublasHandle_t>(cublas_);
    FSB_CUBLAS(cublasSetStream(h, ctx.stream));
    FSB_CUBLAS(cublasSetPointerMode(h, CUBLAS_POINTER_MODE_HOST));
    DBuf Mcur;  // operator of the level below, external numbering (coarsest: the dense inverse)
    for (int l = last - 1; l >= lt; l--) {
      DBuf Mnew;
      level_operator(ctx, h, levels[l], prm, l == last - 1 ? Ainv : Mcur, Mnew);
      Mcur.swap(Mnew);
    }
    Mtail.swap(Mcur);
    tail_level_ = lt;
  } catch (const std::exception& e) {
    cudaGetLastError();
    if (prm.verbose) fprintf(stderr, "dense tail not built (%s): levels >= %d keep their kernels\n", e.what(), lt);
    tail_level_ = -1;
    Mtail.release();
  }
  tail_key_[0] = prm.preInnerIters; tail_key_[1] = prm.postInnerIters; tail_key_[2] = prm.postRelaxes; tail_key_[3] = prm.smootherWeight;
}

// ---------------------------------------------------------------------------------------------------------------
// Dense partition blocks: the same idea one step up.  A level with a few dozen partitions of a few hundred rows (the
// 21 202-row level 2 of the ~100 M-tet cube: 53 partitions) is too large for the dense tail (n^2) but its smoothing
// stages are pure latency — six barrier-separated passes over ~40 k entries per partition take 48 us, the bytes would
// take 4.  Per partition the stage is a fixed linear map of (b) resp. (x_in, b'), so S(nu1), G^nu2 and S(nu2-1) are
// formed once (Horner on the m x m diagonal block, the same recurrences as in level_operator above) and a stage becomes
// one pass of small dense GEMVs (cycle.cu: smooth_blockdense_kernel).  Used where the restricted residual comes from
// R b - (R A) x (no in-partition residual needed) and the three block sets stay below kBlockDenseMaxBytes.
// ---------------------------------------------------------------------------------------------------------------
namespace {
constexpr size_t kBlockDenseMaxBytes = 96u << 20;  // per block set: a stage then streams <= 96 / 192 MB (pre / post)
constexpr int kBlockDenseRows = 32;                 // rows of a block per CTA of the apply kernel

// in-partition entries (incl. the diagonal) of partition rows -> G = I - W B and S0 = W, both as packed m x m blocks
__global__ void bd_make_blocks(int n, const int* __restrict__ ptr, const int* __restrict__ col, const double* __restrict__ val,
                               const double* __restrict__ diag, const int* __restrict__ rowPart, const int* __restrict__ pstart,
                               const long long* __restrict__ off, double w, double* __restrict__ G, double* __restrict__ S0) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int p = rowPart[r], r0 = pstart[p], m = pstart[p + 1] - r0, lr = r - r0;
  const double wr = w / diag[r];
  double* g = G + off[p] + (size_t)lr * m;
  for (int e = ptr[r]; e < ptr[r + 1]; e++) {
    const int c = col[e] - r0;
    if ((unsigned)c < (unsigned)m) g[c] -= wr * val[e];
  }
  g[lr] += 1.0;
  S0[off[p] + (size_t)lr * m + lr] = wr;
}
__global__ void bd_add_w(int n, const double* __restrict__ diag, const int* __restrict__ rowPart, const int* __restrict__ pstart,
                         const long long* __restrict__ off, double w, double* __restrict__ M) {  // M_p += W_p
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int p = rowPart[r], r0 = pstart[p], m = pstart[p + 1] - r0, lr = r - r0;
  M[off[p] + (size_t)lr * m + lr] += w / diag[r];
}
__global__ void bd_identity(int n, const int* __restrict__ rowPart, const int* __restrict__ pstart, const long long* __restrict__ off,
                            double* __restrict__ M) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int p = rowPart[r], r0 = pstart[p], m = pstart[p + 1] - r0, lr = r - r0;
  M[off[p] + (size_t)lr * m + lr] = 1.0;
}
}  // namespace

void Solver::build_block_smoothers() {
  for (auto& L : levels) { L.use_blockdense = false; L.bdS1.release(); L.bdGp.release(); L.bdS2.release(); L.bdOff.release(); L.bdWork.release(); L.bdCtas = 0; }
  const char* env = getenv("FSB_BLOCK_DENSE");  // tuning / test knob, read at every setup: 0 disables
  if (env && atoi(env) == 0) return;
  const int last = (int)levels.size() - 1;
  const int end = tail_level_ >= 0 ? tail_level_ : last;
  cudaStream_t s = ctx.stream;
  const double w = prm.smootherWeight;
  const int nu1 = std::max(0, prm.preInnerIters), nu2 = std::max(0, prm.postInnerIters);
  if (prm.postRelaxes != 1) return;  // the blocks encode one post-relaxation pass
  for (int l = 1; l < end; l++) {
    LevelData& L = levels[l];
    if (L.RA.nrows == 0 || L.use_ell || L.nparts <= 0) continue;  // (the register-resident kernel is the better fit where it applies)
    std::vector<int> ps = L.pstart.to_vector();
    std::vector<long long> off(L.nparts + 1, 0);
    for (int p = 0; p < L.nparts; p++) { const long long m = ps[p + 1] - ps[p]; off[p + 1] = off[p] + m * m; }
    const size_t total = (size_t)off[L.nparts];
    if (total * sizeof(double) > kBlockDenseMaxBytes || L.nparts > 2 * ctx.num_sms) continue;
    try {
      if (!cublas_) {
        cublasHandle_t h;
        FSB_CUBLAS(cublasCreate(&h));
        cublas_ = h;
      }
      cublasHandle_t h = static_cast<cublasHandle_t>(cublas_);
      FSB_CUBLAS(cublasSetStream(h, s));
      FSB_CUBLAS(cublasSetPointerMode(h, CUBLAS_POINTER_MODE_HOST));
      const int n = L.n;
      IBuf rowPart(n, s);
      dt_row_partition<<<cdiv(L.nparts, 128), 128, 0, s>>>(L.nparts, L.pstart, rowPart);
      L.bdOff.alloc(L.nparts + 1, s);
      L.bdOff.from_host(off.data(), L.nparts + 1);
      DBuf G(total, s), Sa(total, s), Sb(total, s);
      G.zero(); Sa.zero();
      L.bdS1.alloc(total, s); L.bdGp.alloc(total, s); L.bdS2.alloc(total, s);
      L.bdS2.zero();
      bd_make_blocks<<<cdiv(n, 128), 128, 0, s>>>(n, L.A.ptr, L.A.col, L.A.val, L.diag, rowPart, L.pstart, L.bdOff, w, G, Sa);
      auto block_mul = [&](const double* A, const double* B, double* C) {  // C_p = A_p B_p for every partition
        for (int p = 0; p < L.nparts; p++) {
          const int m = ps[p + 1] - ps[p];
          gemm_rm(h, m, m, m, 1.0, A + off[p], m, B + off[p], m, 0.0, C + off[p], m);
        }
      };
      // Horner: S(k+1) = W + G S(k), S(0) = W;  S1 = S(nu1), S2 = S(nu2 - 1) (zero for nu2 == 0)
      double* cur = Sa;
      double* nxt = Sb;
      const int steps = std::max(nu1, nu2 - 1);
      for (int k = 0; k <= steps; k++) {
        if (k == nu1) L.bdS1.from_device(cur, total);
        if (k == nu2 - 1) L.bdS2.from_device(cur, total);
        if (k == steps) break;
        block_mul(G, cur, nxt);
        bd_add_w<<<cdiv(n, 128), 128, 0, s>>>(n, L.diag, rowPart, L.pstart, L.bdOff, w, nxt);
        std::swap(cur, nxt);
      }
      // Gp = G^nu2
      {
        DBuf& A = Sa;
        DBuf& B = Sb;
        A.zero();
        bd_identity<<<cdiv(n, 128), 128, 0, s>>>(n, rowPart, L.pstart, L.bdOff, A);
        double* a = A;
        double* b = B;
        for (int k = 0; k < nu2; k++) { block_mul(G, a, b); std::swap(a, b); }
        L.bdGp.from_device(a, total);
      }
      // CTA work list of the apply kernel: kBlockDenseRows rows of one partition per CTA
      std::vector<int> work;
      for (int p = 0; p < L.nparts; p++)
        for (int r = 0; r < ps[p + 1] - ps[p]; r += kBlockDenseRows) { work.push_back(p); work.push_back(r); }
      L.bdCtas = (int)work.size() / 2;
      L.bdWork.alloc(work.size(), s);
      L.bdWork.from_host(work.data(), work.size());
      FSB_CHECK_LAUNCH();
      FSB_CUDA(cudaStreamSynchronize(s));
      L.use_blockdense = true;
    } catch (const std::exception& e) {
      cudaGetLastError();
      if (prm.verbose) fprintf(stderr, "dense partition blocks not built on level %d (%s): the sweep kernels stay\n", l, e.what());
      L.use_blockdense = false;
      L.bdS1.release(); L.bdGp.release(); L.bdS2.release();
    }
  }
}

void Solver::destroy_cublas() {
  if (cublas_) { cublasDestroy(static_cast<cublasHandle_t>(cublas_)); cublas_ = nullptr; }
}

}  // namespace fsb
