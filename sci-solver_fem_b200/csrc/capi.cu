// capi.cu — extern "C" boundary (include/femsolver_b200.h) over fsb::Solver.
#include <cstring>
#include <string>

#include "../../include/femsolver_b200.h"
#include "solver.h"

namespace fsb { void split_by_weight(int nparts, const long long* w, int nranks, int* out); void debug_stamps(int cta_plus1, long long* out64); }

struct fsb_solver {
  fsb::Solver* impl = nullptr;
};

static thread_local std::string g_create_error;

namespace {

template <typename F>
int guarded(fsb_solver* s, F&& f) {
  if (!s || !s->impl) return FSB_ERR_INVALID;
  try {
    // every entry point runs on the solver's own device, whatever the caller's current device is
    if (cudaSetDevice(s->impl->ctx.device) != cudaSuccess) { cudaGetLastError(); throw fsb::CudaError("cudaSetDevice failed for the solver's device"); }
    f(*s->impl);
    return FSB_OK;
  } catch (const std::invalid_argument& e) {
    s->impl->last_error = e.what();
    return FSB_ERR_INVALID;
  } catch (const fsb::CudaError& e) {
    s->impl->last_error = e.what();
    cudaGetLastError();
    return FSB_ERR_CUDA;
  } catch (const std::exception& e) {
    s->impl->last_error = e.what();
    return FSB_ERR_RUNTIME;
  }
}

struct ParamRef { const char* name; int kind; size_t off; };  // kind 0 int, 1 double, 2 unsigned
#define PI_(n) {#n, 0, offsetof(fsb::Params, n)}
#define PD_(n) {#n, 1, offsetof(fsb::Params, n)}
const ParamRef kParams[] = {
  PI_(verbose), PI_(maxLevels), PI_(maxIters), PI_(preInnerIters), PI_(postInnerIters), PI_(postRelaxes), PI_(cycleIters),
  PI_(dsType), PI_(topSize), PI_(randMisParameters), PI_(partitionMaxSize), PI_(aggregatorType), PI_(convergeType),
  PI_(cycleType), PI_(solverType), PI_(device), PI_(blockSize), PD_(tolerance), PD_(smootherWeight), PD_(proOmega),
  {"seed", 2, offsetof(fsb::Params, seed)}, PI_(refLevel0NoPerm), PI_(useGraphs), PI_(checkEvery), PI_(profile)};

const ParamRef* find_param(const char* name) {
  for (const ParamRef& p : kParams) if (strcmp(p.name, name) == 0) return &p;
  return nullptr;
}

long long copy_ints(const fsb::IBuf& b, int* buf, long long cap) {
  long long n = (long long)b.size();
  if (!buf) return n;
  if (cap < n) return -1;
  if (n) b.to_host(buf, n);
  return n;
}
long long copy_vals(const fsb::DBuf& b, double* buf, long long cap) {
  long long n = (long long)b.size();
  if (!buf) return n;
  if (cap < n) return -1;
  if (n) b.to_host(buf, n);
  return n;
}

}  // namespace

extern "C" {

int fsb_version(void) { return 100; }

int fsb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int fsb_create(fsb_solver** out, int device) {
  if (!out) return FSB_ERR_INVALID;
  *out = nullptr;
  try {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      cudaGetLastError();
      g_create_error = "no CUDA device: this library has no CPU fallback";
      return FSB_ERR_CUDA;
    }
    fsb_solver* s = new fsb_solver();
    s->impl = new fsb::Solver(device);
    *out = s;
    return FSB_OK;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    return FSB_ERR_CUDA;
  }
}

void fsb_destroy(fsb_solver* s) {
  if (!s) return;
  delete s->impl;
  delete s;
}

const char* fsb_last_error(const fsb_solver* s) {
  if (!s || !s->impl) return g_create_error.c_str();
  return s->impl->last_error.c_str();
}

int fsb_set_param(fsb_solver* s, const char* name, double value) {
  if (!s || !s->impl || !name) return FSB_ERR_INVALID;
  const ParamRef* p = find_param(name);
  if (!p) { s->impl->last_error = std::string("unknown parameter: ") + name; return FSB_ERR_INVALID; }
  char* base = reinterpret_cast<char*>(&s->impl->prm);
  if (p->kind == 0) *reinterpret_cast<int*>(base + p->off) = (int)value;
  else if (p->kind == 1) *reinterpret_cast<double*>(base + p->off) = value;
  else *reinterpret_cast<unsigned*>(base + p->off) = (unsigned)value;
  return FSB_OK;
}

int fsb_get_param(const fsb_solver* s, const char* name, double* value) {
  if (!s || !s->impl || !name || !value) return FSB_ERR_INVALID;
  const ParamRef* p = find_param(name);
  if (!p) return FSB_ERR_INVALID;
  const char* base = reinterpret_cast<const char*>(&s->impl->prm);
  if (p->kind == 0) *value = *reinterpret_cast<const int*>(base + p->off);
  else if (p->kind == 1) *value = *reinterpret_cast<const double*>(base + p->off);
  else *value = *reinterpret_cast<const unsigned*>(base + p->off);
  return FSB_OK;
}

int fsb_set_tet_mesh(fsb_solver* s, int nv, const double* xyz, int ne, const int* tets, const int* labels) {
  return guarded(s, [&](fsb::Solver& S) { S.set_mesh(nv, xyz, ne, 4, tets, labels, false); });
}
int fsb_set_tri_mesh(fsb_solver* s, int nv, const double* xyz, int ne, const int* tris) {
  return guarded(s, [&](fsb::Solver& S) { S.set_mesh(nv, xyz, ne, 3, tris, nullptr, false); });
}
int fsb_set_tet_mesh_device(fsb_solver* s, int nv, const double* xyz, int ne, const int* tets, const int* labels) {
  return guarded(s, [&](fsb::Solver& S) { S.set_mesh(nv, xyz, ne, 4, tets, labels, true); });
}
int fsb_set_tri_mesh_device(fsb_solver* s, int nv, const double* xyz, int ne, const int* tris) {
  return guarded(s, [&](fsb::Solver& S) { S.set_mesh(nv, xyz, ne, 3, tris, nullptr, true); });
}
int fsb_assemble(fsb_solver* s) { return guarded(s, [&](fsb::Solver& S) { S.assemble(); }); }
int fsb_matrix_rows(const fsb_solver* s) { return (s && s->impl) ? s->impl->rows() : 0; }
long long fsb_matrix_nnz(const fsb_solver* s) { return (s && s->impl) ? s->impl->nnz() : 0; }
int fsb_get_matrix_csr(fsb_solver* s, int* rowptr, int* col, double* val) {
  return guarded(s, [&](fsb::Solver& S) { S.get_matrix(rowptr, col, val); });
}
int fsb_set_matrix_values(fsb_solver* s, const double* val) {
  return guarded(s, [&](fsb::Solver& S) { S.set_matrix_values(val, false); });
}
int fsb_set_matrix_csr(fsb_solver* s, int n, long long nnz, const int* rowptr, const int* col, const double* val) {
  return guarded(s, [&](fsb::Solver& S) { S.set_matrix_csr(n, (int)nnz, rowptr, col, val); });
}

int fsb_setup(fsb_solver* s) { return guarded(s, [&](fsb::Solver& S) { S.setup(); }); }
int fsb_num_levels(const fsb_solver* s) { return (s && s->impl) ? s->impl->num_levels() : 0; }
int fsb_level_rows(const fsb_solver* s, int level) {
  if (!s || !s->impl || level < 0 || level >= s->impl->num_levels()) return -1;
  return s->impl->level(level).A.nrows;
}
long long fsb_level_nnz(const fsb_solver* s, int level) {
  if (!s || !s->impl || level < 0 || level >= s->impl->num_levels()) return -1;
  return s->impl->level(level).A.nnz;
}

long long fsb_level_stat(const fsb_solver* s, int level, const char* name) {
  if (!s || !s->impl || !name || level < 0 || level >= s->impl->num_levels()) return -1;
  const fsb::LevelData& L = s->impl->level(level);
  const std::string n(name);
  if (n == "nparts") return L.nparts;
  if (n == "max_part_rows") return L.maxPartRows;
  if (n == "smoother") return L.use_blockdense ? 4 : L.use_ell ? 0 : L.use_sellg ? 1 : L.smemBytes > 0 ? 2 : 3;  // 0 register ELL, 1 shared-memory ELL, 2 cluster, 3 cooperative, 4 dense blocks
  if (n == "dense_tail") return level == s->impl->tail_level_ ? 1 : 0;
  if (n == "sell") return L.sA.ready() || L.sAout.ready() ? 1 : 0;
  return -1;
}
long long fsb_level_int(fsb_solver* s, int level, const char* name, int* buf, long long cap) {
  if (!s || !s->impl || level < 0 || level >= s->impl->num_levels()) return -1;
  long long rc = -1;
  int g = guarded(s, [&](fsb::Solver& S) {
    const fsb::LevelData& L = S.level(level);
    std::string n(name);
    const fsb::IBuf* b = nullptr;
    if (n == "permutation") b = &L.agg.permutation; else if (n == "ipermutation") b = &L.agg.ipermutation;
    else if (n == "aggregateIdx") b = &L.agg.aggregateIdx; else if (n == "partitionIdx") b = &L.agg.partitionIdx;
    else if (n == "partitionLabel") b = &L.agg.partitionLabel; else if (n == "xadjOut") b = &L.agg.xadjOut;
    else if (n == "adjOut") b = &L.agg.adjOut; else if (n == "A_ptr") b = &L.A.ptr; else if (n == "A_col") b = &L.A.col;
    else if (n == "P_ptr") b = &L.P.ptr; else if (n == "P_col") b = &L.P.col; else if (n == "R_ptr") b = &L.R.ptr;
    else if (n == "R_col") b = &L.R.col; else if (n == "pstart") b = &L.pstart;
    else if (n == "Aout_ptr") b = &L.Aout.ptr; else if (n == "Aout_col") b = &L.Aout.col;
    else throw std::invalid_argument("unknown level array: " + n);
    rc = copy_ints(*b, buf, cap);
  });
  return g == FSB_OK ? rc : g;
}

long long fsb_level_val(fsb_solver* s, int level, const char* name, double* buf, long long cap) {
  if (!s || !s->impl || level < 0 || level >= s->impl->num_levels()) return -1;
  long long rc = -1;
  int g = guarded(s, [&](fsb::Solver& S) {
    const fsb::LevelData& L = S.level(level);
    std::string n(name);
    const fsb::DBuf* b = nullptr;
    if (n == "A_val") b = &L.A.val; else if (n == "P_val") b = &L.P.val; else if (n == "R_val") b = &L.R.val;
    else if (n == "diag") b = &L.diag; else if (n == "Ainv" && level == S.num_levels() - 1) b = &S.Ainv;
    else throw std::invalid_argument("unknown level array: " + n);
    rc = copy_vals(*b, buf, cap);
  });
  return g == FSB_OK ? rc : g;
}

static void report(fsb::Solver& S, int* iters, double* relres) {
  if (iters) *iters = S.iterations;
  if (relres) *relres = S.final_relres;
}
int fsb_solve(fsb_solver* s, const double* b, double* x, int* iters, double* relres) {
  return guarded(s, [&](fsb::Solver& S) { S.solve(b, x, false); report(S, iters, relres); });
}
int fsb_solve_device(fsb_solver* s, const double* b, double* x, int* iters, double* relres) {
  return guarded(s, [&](fsb::Solver& S) { S.solve(b, x, true); report(S, iters, relres); });
}
int fsb_solve_fem(fsb_solver* s, const double* b, double* x, int* iters, double* relres) {
  return guarded(s, [&](fsb::Solver& S) { S.setup(); S.solve(b, x, false); report(S, iters, relres); });
}
int fsb_resid_history(const fsb_solver* s, double* buf, int cap) {
  if (!s || !s->impl) return -1;
  int n = (int)s->impl->resid_history.size();
  if (!buf) return n;
  if (cap < n) return -1;
  memcpy(buf, s->impl->resid_history.data(), sizeof(double) * n);
  return n;
}

int fsb_apply_matrix_device(fsb_solver* s, const double* x, double* y) {
  return guarded(s, [&](fsb::Solver& S) { S.apply_matrix(x, y); });
}
int fsb_apply_matrix(fsb_solver* s, const double* x, double* y) {
  return guarded(s, [&](fsb::Solver& S) {
    const int n = S.rows();
    if (n == 0) throw std::invalid_argument("Error no matrix specified");
    fsb::DBuf xd(n, S.ctx.stream), yd(n, S.ctx.stream);
    xd.from_host(x, n);
    S.apply_matrix(xd, yd);
    yd.to_host(y, n);
  });
}
int fsb_spmv_fine_device(fsb_solver* s, const double* x, double* y) {
  return guarded(s, [&](fsb::Solver& S) { if (!S.has_setup) throw std::runtime_error("setup first"); S.spmv_fine(x, y); });
}
int fsb_precondition_device(fsb_solver* s, const double* r, double* z) {
  return guarded(s, [&](fsb::Solver& S) { S.precondition(r, z); });
}

double fsb_time_ms(const fsb_solver* s, const char* stage) {
  if (!s || !s->impl) return -1;
  auto it = s->impl->times_ms.find(stage);
  return it == s->impl->times_ms.end() ? -1.0 : it->second;
}
long long fsb_last_launches(const fsb_solver* s) { return (s && s->impl) ? s->impl->launches : 0; }
int fsb_profile_report(fsb_solver* s, char* buf, int cap) {
  if (!s || !s->impl) return -1;
  std::string r;
  try { r = s->impl->profile_report(); } catch (const std::exception& e) { s->impl->last_error = e.what(); return -1; }
  if (!buf) return (int)r.size() + 1;
  if (cap < (int)r.size() + 1) return -1;
  memcpy(buf, r.c_str(), r.size() + 1);
  return (int)r.size() + 1;
}
int fsb_dist_prepare(fsb_solver* s, int rank, int nranks) { return guarded(s, [&](fsb::Solver& S) { S.dist_prepare(rank, nranks); }); }
int fsb_dist_blob_bytes(void) { return fsb::Solver::kDistBlobBytes; }
int fsb_dist_blob(fsb_solver* s, void* blob, long long* arena_bytes) {
  return guarded(s, [&](fsb::Solver& S) { S.dist_get_blob(blob, arena_bytes); });
}
int fsb_dist_connect(fsb_solver* s, const void* blobs) { return guarded(s, [&](fsb::Solver& S) { S.dist_connect(blobs); }); }
int fsb_dist_disconnect(fsb_solver* s) { return guarded(s, [&](fsb::Solver& S) { S.dist_disconnect(); }); }
int fsb_dist_ranges(const fsb_solver* s, int* part_begin, int* row_begin, int* coarse_begin) {
  return fsb_dist_level_ranges(s, 0, part_begin, row_begin, coarse_begin);
}
int fsb_dist_level_ranges(const fsb_solver* s, int level, int* part_begin, int* row_begin, int* coarse_begin) {
  if (!s || !s->impl) return FSB_ERR_INVALID;
  const auto& d = s->impl->dist;
  if (level < 0 || level >= (int)d.lev.size()) return FSB_ERR_INVALID;
  const auto& L = d.lev[level];
  for (int r = 0; r <= d.nranks; r++) {
    if (part_begin) part_begin[r] = L.pbeg[r];
    if (row_begin) row_begin[r] = L.rbeg[r];
    if (coarse_begin) coarse_begin[r] = L.abeg[r];
  }
  return d.nranks;
}
int fsb_dist_info(const fsb_solver* s, int* sharded_levels, int* user_lo, int* user_hi, long long* halo_values) {
  if (!s || !s->impl) return FSB_ERR_INVALID;
  const auto& d = s->impl->dist;
  if (sharded_levels) *sharded_levels = d.nshard;
  if (user_lo) *user_lo = d.user_lo;
  if (user_hi) *user_hi = d.user_hi;
  if (halo_values) {  // per sharded level: entries of the four push lists (operator columns, residual rows, down, up)
    for (int l = 0; l < d.nshard; l++) {
      halo_values[4 * l + 0] = d.lev[l].sendA.total; halo_values[4 * l + 1] = d.lev[l].sendR.total;
      halo_values[4 * l + 2] = d.lev[l].sendDown.total; halo_values[4 * l + 3] = d.lev[l].sendUp.total;
    }
  }
  return d.nranks;
}
int fsb_dist_interior(const fsb_solver* s, int level, int* out6) {
  if (!s || !s->impl || !out6) return FSB_ERR_INVALID;
  const auto& d = s->impl->dist;
  if (level < 0 || level >= (int)d.lev.size()) return FSB_ERR_INVALID;
  const auto& L = d.lev[level];
  out6[0] = L.intA.begin; out6[1] = L.intA.end; out6[2] = L.intR.begin; out6[3] = L.intR.end; out6[4] = L.intP.begin; out6[5] = L.intP.end;
  return FSB_OK;
}
// tools only (not part of the public header): `reps` back-to-back exchanges of channel `chan` (no compute in between, all
// ranks must call it together) timed with CUDA events -> microseconds per exchange; chan < 0: all-reduces instead
double fsb_dist_bench_exchange(fsb_solver* s, int chan, int reps) {
  double us = -1.0;
  guarded(s, [&](fsb::Solver& S) { us = S.dist_bench_exchange(chan, reps); });
  return us;
}
void fsb_split_by_weight(int nparts, const long long* weights, int nranks, int* out_begin) { fsb::split_by_weight(nparts, weights, nranks, out_begin); }
// tools only (not part of the public header): phase timestamps of one CTA of the cluster smoother
void fsb_debug_stamps(int cta_plus1, long long* out64) { fsb::debug_stamps(cta_plus1, out64); }
void* fsb_stream(const fsb_solver* s) { return (s && s->impl) ? (void*)s->impl->ctx.stream : nullptr; }

void fsb_tet_mass_integrals(double out10[10]) { fsb::tet_mass_integrals_host(out10); }
void fsb_tri_quadrature(double zx[6], double zy[6], double wx[6], double wy[6]) { fsb::tri_quadrature_host(zx, zy, wx, wy); }

}  // extern "C"
