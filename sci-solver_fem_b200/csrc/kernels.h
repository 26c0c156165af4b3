// kernels.h — launch wrappers of the solve-phase kernels (cycle.cu, pcg.cu).
#pragma once
#include "solver.h"

namespace fsb {

// y = A x (mode 0), y = b - A x (1), y += A x (2), y -= A x (3); CSR-stream kernel. `name` tags the profile.
void launch_spmv(const Ctx& c, const DCsr& A, const double* x, double* y, int mode, const double* b, const int* done, const char* name,
                 RowRange rr = RowRange());
// same operation on the SELL-32 copy of an operator (thread per row, coalesced, no staging)
void launch_spmv_sell(const Ctx& c, const Sell& A, const double* x, double* y, int mode, const double* b, const int* done, const char* name,
                      RowRange rr = RowRange());
void launch_spmv_dot_sell(const Ctx& c, const Sell& A, const double* x, double* y, double* partials, PcgScalars* sc, RowRange rr = RowRange());
// one part of a split y = A x, p.y: CTA sums parked behind the `pprev` sums of the earlier parts; the last part (finalize)
// finishes the dot product.  Returns the number of CTA sums this part parked.
int launch_spmv_dot_sell_part(const Ctx& c, const Sell& A, const double* x, double* y, double* partials, PcgScalars* sc, RowRange rr, int pprev, bool finalize);
// multi-GPU exchanges over NVLink peer memory (cycle.cu): flag-in-data exchange of list entries of `src` (mine) into `dst`
// (what the peers send me); fenced all-gather of a slice (end of a solve)
// y[r] += (A x)[r] for the listed rows (thread per row; ghost rows of a sharded level)
void launch_spmv_list_add(const Ctx& c, const DCsr& A, const int* rows, int count, const double* x, double* y, const int* done);
void launch_ll_exchange(const Ctx& c, const LLXchg& x, const double* src, double* dst, const int* done);
void launch_push_all(const Ctx& c, int begin, int end, const double* src, const PeerPtrs& dst);
// y = A x and the dot product x.y folded into the same pass; the last CTA finishes
// py and alpha = rz_old / py in device memory.
void launch_spmv_dot(const Ctx& c, const DCsr& A, const double* x, double* y, double* partials, PcgScalars* sc);

// fused smoothing stage of one level (partition-local Jacobi sweeps, x resident on chip):
//   pre  (x_in == null): x = w b / d, nsweeps sweeps, x -> x_out, in-partition residual -> r_out
//   post (x_in != null): nsweeps sweeps from x_in with b = b_src (already b - A_out x_in)
// b_src is read through `gather` (level > 0: external -> internal numbering) and saved to b_int when
// non-null; the result goes to x_out (internal) and/or is scattered to x_ext through `scatter`.
void launch_smooth(const Ctx& c, const LevelData& L, const double* b_src, const int* gather, double* b_int, const double* x_in,
                   double w, int nsweeps, double* x_out, const int* scatter, double* x_ext, double* r_out, const int* done,
                   bool owned_only = false);
// bc = R b - T x with T = R A: the restricted residual of a level without forming the residual (warp per coarse row)
void launch_restrict_fused(const Ctx& c, const DCsr& R, const double* b, const DCsr& T, const double* x, double* bc, const int* done);
void launch_coarse_solve(const Ctx& c, int n, const double* Ainv, const double* b, double* x, const int* done);

// PCG vector kernels (device-resident scalars)
void launch_cg_init(const Ctx& c, PcgScalars* sc, double tol, int maxit, int hist_cap);
void launch_dot(const Ctx& c, int n, const double* a, const double* b, double* partials, PcgScalars* sc, int which);  // which: 0 bnorm, 1 rz(first), 2 rz(new)+beta
void launch_cg_update(const Ctx& c, int n, double* x, double* r, const double* p, const double* y, double* partials, PcgScalars* sc, double* hist);
void launch_cg_pdir(const Ctx& c, int n, double* p, const double* z, const PcgScalars* sc, int first);
void launch_gather(const Ctx& c, int n, const int* idx, const double* src, double* dst);   // dst[i] = src[idx[i]]
void launch_scatter(const Ctx& c, int n, const int* idx, const double* src, double* dst);  // dst[idx[i]] = src[i]

extern thread_local long long g_launch_counter;

}  // namespace fsb
