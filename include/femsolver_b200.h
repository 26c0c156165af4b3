/* femsolver_b200.h — C ABI of the B200-native SCI-Solver_FEM solve path.
 *
 * The reference exposes this path as a C++ class compiled into static libraries
 * (class FEMSolver, /root/reference/src/FEMSolver.h:14-67; callers: src/examples/example1.cu:36-67,
 * example2.cu:32-64, src/test/sanity2D.cc:5-15, sanity3D.cc:5-15, tetVol.cc:5-15).  This header is
 * the plain-C boundary underneath our drop-in FEMSolver class (sci-solver_fem_b200/csrc/FEMSolver.h)
 * and the binding point for ctypes / cgo / JNI hosts: opaque handle, plain pointers and sizes,
 * no C++ or torch types.  Every function returns 0 on success and a negative code on failure;
 * fsb_last_error() gives the message.  All compute runs on the CUDA device chosen at creation;
 * there is no CPU fallback — without a device fsb_create() fails.
 *
 * One entry point per reference interface:
 *   fsb_create / fsb_destroy            FEMSolver::FEMSolver / ~FEMSolver        FEMSolver.cu:9-51
 *   fsb_set_param / fsb_get_param       the 20 public configuration fields       FEMSolver.h:41-61
 *   fsb_set_tet_mesh / fsb_set_tri_mesh TetMesh::read / TriMesh::read results    FEMSolver.cu:36-42
 *   fsb_assemble                        FEMSolver::getMatrixFromMesh             FEMSolver.cu:141-171
 *   fsb_matrix_rows                     FEMSolver::getMatrixRows                 FEMSolver.cu:95-97
 *   fsb_get_matrix_csr                  read access to FEMSolver::A_h_           FEMSolver.h:66
 *   fsb_set_matrix_values / _csr        FEMSolver::readMatlabSparseMatrix result FEMSolver.cu:177-356
 *   fsb_setup                           AMG::setup                               core/cuda/amg.cu:79-144
 *   fsb_solve                           AMG::solve (inside solveFEM)             amg.cu:173-200, FEMSolver.cu:77-87
 *   fsb_solve_fem                       FEMSolver::solveFEM (setup + solve)      FEMSolver.cu:58-93
 *   fsb_level_* getters                 AMG::printGridStatistics + test hooks    amg.cu:202-231
 */
#ifndef FEMSOLVER_B200_H
#define FEMSOLVER_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fsb_solver fsb_solver;

#define FSB_OK 0
#define FSB_ERR_INVALID (-1) /* std::invalid_argument upstream, e.g. "Error no matrix specified" */
#define FSB_ERR_RUNTIME (-2)
#define FSB_ERR_CUDA (-3)

/* library / device */
int fsb_version(void);
int fsb_device_count(void);
int fsb_create(fsb_solver** out, int device);
void fsb_destroy(fsb_solver* s);
const char* fsb_last_error(const fsb_solver* s); /* s may be NULL: error of the last failed fsb_create */

/* parameters: names are the reference field names without the trailing underscore
 * (maxLevels, maxIters, preInnerIters, postInnerIters, postRelaxes, cycleIters, dsType, topSize,
 *  randMisParameters, partitionMaxSize, aggregatorType, convergeType, tolerance, cycleType,
 *  solverType, smootherWeight, proOmega, device, blockSize, verbose) plus the additive ones
 * (seed, refLevel0NoPerm, useGraphs, checkEvery, profile). */
int fsb_set_param(fsb_solver* s, const char* name, double value);
int fsb_get_param(const fsb_solver* s, const char* name, double* value);

/* stage 1: mesh in, matrix on device.  xyz: nv*3 doubles (x,y,z per vertex; z ignored for triangles);
 * tets: ne*4 / tris: ne*3 zero-based vertex ids; labels: ne material labels or NULL. Host pointers. */
int fsb_set_tet_mesh(fsb_solver* s, int nv, const double* xyz, int ne, const int* tets, const int* labels);
int fsb_set_tri_mesh(fsb_solver* s, int nv, const double* xyz, int ne, const int* tris);
/* same, with DEVICE pointers (inputs already resident in HBM) */
int fsb_set_tet_mesh_device(fsb_solver* s, int nv, const double* xyz, int ne, const int* tets, const int* labels);
int fsb_set_tri_mesh_device(fsb_solver* s, int nv, const double* xyz, int ne, const int* tris);
int fsb_assemble(fsb_solver* s);
int fsb_matrix_rows(const fsb_solver* s);
long long fsb_matrix_nnz(const fsb_solver* s);
/* any of rowptr (rows+1), col (nnz), val (nnz) may be NULL */
int fsb_get_matrix_csr(fsb_solver* s, int* rowptr, int* col, double* val);
int fsb_set_matrix_values(fsb_solver* s, const double* val);
int fsb_set_matrix_csr(fsb_solver* s, int n, long long nnz, const int* rowptr, const int* col, const double* val);

/* stage 2 */
int fsb_setup(fsb_solver* s);
int fsb_num_levels(const fsb_solver* s);
int fsb_level_rows(const fsb_solver* s, int level);
long long fsb_level_nnz(const fsb_solver* s, int level);
/* integer arrays of a level: permutation, ipermutation, aggregateIdx, partitionIdx, partitionLabel,
 * xadjOut, adjOut, A_ptr, A_col, P_ptr, P_col, R_ptr, R_col, pstart.  buf == NULL: returns the length. */
/* scalar facts about a level: "nparts", "max_part_rows", "smoother" (0 register-resident ELL, 1 shared-memory ELL,
 * 2 cluster, 3 cooperative, 4 dense partition blocks), "dense_tail" (1: this level starts the dense tail), "sell" */
long long fsb_level_stat(const fsb_solver* s, int level, const char* name);
long long fsb_level_int(fsb_solver* s, int level, const char* name, int* buf, long long cap);
/* value arrays: A_val, P_val, R_val, diag; level == num_levels-1 also: Ainv */
long long fsb_level_val(fsb_solver* s, int level, const char* name, double* buf, long long cap);

/* stage 3: x holds the initial guess on entry and the solution on exit (FEMSolver.cu:81,86).
 * iters / relres may be NULL. */
int fsb_solve(fsb_solver* s, const double* b, double* x, int* iters, double* relres);
int fsb_solve_device(fsb_solver* s, const double* b_dev, double* x_dev, int* iters, double* relres);
int fsb_solve_fem(fsb_solver* s, const double* b, double* x, int* iters, double* relres); /* setup + solve */
/* relative residual ||r||/||b|| after every PCG iteration; buf == NULL returns the length */
int fsb_resid_history(const fsb_solver* s, double* buf, int cap);

/* kernel-level hooks (device pointers, level-0 partition-contiguous numbering) for parity tests and
 * micro-benchmarks: y = A x on the permuted fine operator; z = one V-cycle applied to r. */
int fsb_spmv_fine_device(fsb_solver* s, const double* x_dev, double* y_dev);
/* y = A x with the assembled (user-ordered) matrix, device pointers: builds right-hand sides without
 * moving the matrix to the host */
int fsb_apply_matrix_device(fsb_solver* s, const double* x_dev, double* y_dev);
int fsb_apply_matrix(fsb_solver* s, const double* x_host, double* y_host);   /* the same with host vectors */
int fsb_precondition_device(fsb_solver* s, const double* r_dev, double* z_dev);

/* stage 4 — sharded solve over the GPUs of one box (absent upstream; one process per GPU).
 * Every process builds the same mesh and calls fsb_setup (replicated, deterministic); then
 *   fsb_dist_prepare(rank, nranks)  deals the partitions of the fine level — and of every coarser level that still
 *                                   gives each GPU enough rows — out in contiguous nnz-balanced ranges, builds
 *                                   the push lists of every exchange and allocates the exchange arena;
 *   fsb_dist_blob                   returns this process's connection record (fsb_dist_blob_bytes() bytes: the CUDA
 *                                   IPC handle of the arena and the layout of its receive buffers); all-gather the
 *                                   records of all processes in rank order (e.g. with torch.distributed / MPI);
 *   fsb_dist_connect(blobs)         maps the peers' arenas and checks that every pair of GPUs agrees on what it
 *                                   exchanges (nranks x fsb_dist_blob_bytes() bytes, rank order).
 * After that fsb_solve / fsb_solve_device with solverType = 1 run sharded; all processes must call them
 * together.  fsb_solve_device reads b / x0 at this GPU's rows and leaves the FULL solution in x on every
 * GPU; fsb_solve (host buffers) moves only the slice [user_lo, user_hi) of b, x0 and x over PCIe (the
 * range of user-numbered rows that covers this GPU's rows, fsb_dist_info) — x outside it is untouched.
 * fsb_setup or fsb_dist_disconnect ends the sharded mode. */
int fsb_dist_prepare(fsb_solver* s, int rank, int nranks);
int fsb_dist_blob_bytes(void);
int fsb_dist_blob(fsb_solver* s, void* blob, long long* arena_bytes);
int fsb_dist_connect(fsb_solver* s, const void* blobs);
int fsb_dist_disconnect(fsb_solver* s);
/* first partition / first fine row / first coarse row of every rank (nranks+1 entries each); returns nranks */
int fsb_dist_ranges(const fsb_solver* s, int* part_begin, int* row_begin, int* coarse_begin);
/* the same for sharded level `level` (rows in that level's permuted numbering) */
int fsb_dist_level_ranges(const fsb_solver* s, int level, int* part_begin, int* row_begin, int* coarse_begin);
/* number of sharded levels, the host-copy slice, and per sharded level the number of values this GPU pushes per
 * exchange (4 entries per level: operator halo, residual halo, down, up; halo_values may be NULL); returns nranks */
int fsb_dist_info(const fsb_solver* s, int* sharded_levels, int* user_lo, int* user_hi, long long* halo_values);
/* interior ranges of sharded level `level` on this GPU (begin, end pairs; begin >= end: none): rows whose operator rows,
 * coarse rows whose restriction rows, rows whose prolongator rows reference nothing that arrives through an exchange —
 * their share of the consumer kernel runs while the exchange is in flight */
int fsb_dist_interior(const fsb_solver* s, int level, int* out6);
/* host-only helper: contiguous split of weighted partitions over nranks (out_begin has nranks+1 entries) */
void fsb_split_by_weight(int nparts, const long long* weights, int nranks, int* out_begin);

/* measurements of the last calls, milliseconds, CUDA events on the solver's stream:
 * "pattern", "assemble", "setup", "solve" */
double fsb_time_ms(const fsb_solver* s, const char* stage);
long long fsb_last_launches(const fsb_solver* s); /* kernels launched by the last fsb_solve* */
/* with parameter profile=1 the next solve times every kernel with CUDA events (graphs off); the
 * report is text, one line per (kernel, level): "name level launches total_ms".
 * buf == NULL returns the size needed. */
int fsb_profile_report(fsb_solver* s, char* buf, int cap);
void* fsb_stream(const fsb_solver* s);            /* cudaStream_t the kernels run on */

/* host-side tables the element kernels consume (no device needed) */
void fsb_tet_mass_integrals(double out10[10]);
void fsb_tri_quadrature(double zx[6], double zy[6], double wx[6], double wy[6]);

#ifdef __cplusplus
}
#endif
#endif /* FEMSOLVER_B200_H */
