// kernels.h — launch wrappers of the solve-phase kernels (cycle.cu, pcg.cu).
#pragma once
#include "solver.h"

namespace fsb {

// y = A x (mode 0), y = b - A x (mode 1), y = y + A x (mode 2); lanes-per-row chosen from nnz/row.
void launch_spmv(const Ctx& c, const DCsr& A, const double* x, double* y, int mode, const double* b, const int* done);
// y = A x and the dot product x.y folded into the same pass; the last CTA finishes
// py and alpha = rz_old / py in device memory.
void launch_spmv_dot(const Ctx& c, const DCsr& A, const double* x, double* y, double* partials, PcgScalars* sc);

// fused pre-smoothing of one level: x = w b / d, nsweeps partition-local Jacobi sweeps.
// b_src is read through `gather` (level > 0: external -> internal numbering) and, when b_int
// is non-null, saved in internal numbering for the residual and post-smoothing stages.
void launch_pre_smooth(const Ctx& c, const LevelData& L, const double* b_src, const int* gather, double* b_int,
                       double w, int nsweeps, double* x, const int* done);
// fused post-smoothing: b' = b - A_out x_in, nsweeps sweeps from x_in, result to x_out
// (internal) and/or scattered to x_ext through `scatter`.
void launch_post_smooth(const Ctx& c, const LevelData& L, const double* b_int, const double* x_in, double w, int nsweeps,
                        double* x_out, const int* scatter, double* x_ext, const int* done);
void launch_coarse_solve(const Ctx& c, int n, const double* Ainv, const double* b, double* x, const int* done);

// PCG vector kernels (device-resident scalars)
void launch_cg_init(const Ctx& c, PcgScalars* sc, double tol, int maxit);
void launch_dot(const Ctx& c, int n, const double* a, const double* b, double* partials, PcgScalars* sc, int which);  // which: 0 bnorm, 1 rz(first), 2 rz(new)+beta
void launch_cg_update(const Ctx& c, int n, double* x, double* r, const double* p, const double* y, double* partials, PcgScalars* sc, double* hist);
void launch_cg_pdir(const Ctx& c, int n, double* p, const double* z, const PcgScalars* sc, int first);
void launch_gather(const Ctx& c, int n, const int* idx, const double* src, double* dst);   // dst[i] = src[idx[i]]
void launch_scatter(const Ctx& c, int n, const int* idx, const double* src, double* dst);  // dst[idx[i]] = src[i]

extern thread_local long long g_launch_counter;

}  // namespace fsb
