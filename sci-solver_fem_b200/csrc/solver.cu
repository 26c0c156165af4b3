// solver.cu — host orchestration: FEMSolver::getMatrixFromMesh / solveFEM equivalents.
//
// Reference call stacks: FEMSolver::getMatrixFromMesh (src/FEMSolver.cu:141-171),
// FEMSolver::solveFEM (:58-93), AMG::setup (src/core/cuda/amg.cu:79-144), AMG::solve /
// solve_iteration (:149-200), AMG_Level::cycle / cycle_level0 (amg_level.cu:22-128),
// CG_Flex_Cycle (cgcycle.cu:6-69).
#include "solver.h"

#include <algorithm>
#include <cmath>
#include <chrono>
#include <climits>
#include <cstring>

#include "kernels.h"

namespace fsb {

Solver::Solver(int device) {
  ctx.device = device;
  FSB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  FSB_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx.num_sms = prop.multiProcessorCount;
  FSB_CUDA(cudaStreamCreateWithFlags(&ctx.stream, cudaStreamNonBlocking));
  FSB_CUDA(cudaEventCreateWithFlags(&ctx.ev_fork, cudaEventDisableTiming));
  for (int q = 0; q < 2; q++) {
    FSB_CUDA(cudaStreamCreateWithFlags(&ctx.side[q], cudaStreamNonBlocking));
    FSB_CUDA(cudaEventCreateWithFlags(&ctx.ev_join[q], cudaEventDisableTiming));
  }
  FSB_CUDA(cudaStreamCreateWithFlags(&ctx.xstream, cudaStreamNonBlocking));
  FSB_CUDA(cudaEventCreateWithFlags(&ctx.ev_xfork, cudaEventDisableTiming));
  FSB_CUDA(cudaEventCreateWithFlags(&ctx.ev_xjoin, cudaEventDisableTiming));
  // keep freed setup temporaries in the pool: setup allocates hundreds of short-lived buffers
  cudaMemPool_t pool;
  FSB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thr = UINT64_MAX;
  FSB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  FSB_CUDA(cudaEventCreate(&ev0_));
  FSB_CUDA(cudaEventCreate(&ev1_));
  FSB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&poll_host_), 2 * sizeof(PcgScalars)));
  for (int q = 0; q < 2; q++) FSB_CUDA(cudaEventCreateWithFlags(&poll_ev_[q], cudaEventDisableTiming));
  prm.device = device;
  ctx.prof = &profiler;
}

Solver::~Solver() {
  try { dist_disconnect(); } catch (...) {}
  destroy_graph();
  destroy_cublas();
  levels.clear();
  if (ev0_) cudaEventDestroy(ev0_);
  if (ev1_) cudaEventDestroy(ev1_);
  for (int q = 0; q < 2; q++) if (poll_ev_[q]) cudaEventDestroy(poll_ev_[q]);
  if (poll_host_) cudaFreeHost(poll_host_);
  // buffers are stream-ordered: drain before the stream goes away
  cudaStreamSynchronize(ctx.stream);
  for (int q = 0; q < 2; q++) {
    if (ctx.side[q]) { cudaStreamSynchronize(ctx.side[q]); cudaStreamDestroy(ctx.side[q]); }
    if (ctx.ev_join[q]) cudaEventDestroy(ctx.ev_join[q]);
  }
  if (ctx.ev_fork) cudaEventDestroy(ctx.ev_fork);
  if (ctx.xstream) { cudaStreamSynchronize(ctx.xstream); cudaStreamDestroy(ctx.xstream); }
  if (ctx.ev_xfork) cudaEventDestroy(ctx.ev_xfork);
  if (ctx.ev_xjoin) cudaEventDestroy(ctx.ev_xjoin);
}

void Solver::destroy_graph() {
  if (iter_graph_) { cudaGraphExecDestroy(iter_graph_); iter_graph_ = nullptr; }
}

void Solver::tic(const char*) { FSB_CUDA(cudaEventRecord(ev0_, ctx.stream)); }
void Solver::toc(const char* name) {
  FSB_CUDA(cudaEventRecord(ev1_, ctx.stream));
  FSB_CUDA(cudaEventSynchronize(ev1_));
  float ms = 0;
  FSB_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
  times_ms[name] = ms;
}

// ----------------------------------------------------------------------------- stage 1
// smallest / largest vertex index of the element list (a malformed .ele file must not reach the pattern kernels)
__global__ static void index_range_kernel(long long m, const int* __restrict__ e, int* __restrict__ lohi) {
  int lo = INT_MAX, hi = INT_MIN;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < m; k += (long long)gridDim.x * blockDim.x) {
    const int v = e[k];
    lo = min(lo, v); hi = max(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if ((threadIdx.x & 31) == 0 && lo <= hi) { atomicMin(&lohi[0], lo); atomicMax(&lohi[1], hi); }
}

void Solver::set_mesh(int nv, const double* xyz, int ne, int npe, const int* elems, const int* labels, bool on_device) {
  if (npe != 3 && npe != 4) throw std::invalid_argument("npe must be 3 (triangles) or 4 (tetrahedra)");
  if (nv <= 0 || ne < 0 || !xyz || (ne > 0 && !elems)) throw std::invalid_argument("empty mesh");
  // capacity of the 32-bit contribution ids / entry counts of the pattern stage (pattern.cu)
  if ((long long)ne * npe * npe + nv >= 0xffffffffLL) throw std::invalid_argument("mesh too large: elements x slots must stay below 2^32");
  FSB_CUDA(cudaSetDevice(ctx.device));
  cudaStream_t s = ctx.stream;
  mesh.nv = nv; mesh.ne = ne; mesh.npe = npe;
  mesh.xyz.alloc((size_t)nv * 3, s);
  mesh.elems.alloc((size_t)ne * npe, s);
  cudaMemcpyKind k = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  FSB_CUDA(cudaMemcpyAsync(mesh.xyz.get(), xyz, sizeof(double) * 3 * nv, k, s));
  FSB_CUDA(cudaMemcpyAsync(mesh.elems.get(), elems, sizeof(int) * (size_t)ne * npe, k, s));
  if (labels && npe == 4) {
    mesh.labels.alloc(ne, s);
    FSB_CUDA(cudaMemcpyAsync(mesh.labels.get(), labels, sizeof(int) * ne, k, s));
  } else mesh.labels.release();
  if (ne > 0) {
    IBuf lohi(2, s);
    const int init[2] = {INT_MAX, INT_MIN};
    lohi.from_host(init, 2);
    const long long m = (long long)ne * npe;
    index_range_kernel<<<std::min(cdiv(m, 1024), 4 * ctx.num_sms), 256, 0, s>>>(m, mesh.elems, lohi);
    FSB_CHECK_LAUNCH();
    std::vector<int> r = lohi.to_vector();
    if (r[0] < 0 || r[1] >= nv) {
      mesh.nv = mesh.ne = 0; mesh.xyz.release(); mesh.elems.release(); mesh.labels.release();
      throw std::invalid_argument("element vertex index out of range [0, " + std::to_string(nv) + "): " + std::to_string(r[0] < 0 ? r[0] : r[1]));
    }
  }
  FSB_CUDA(cudaStreamSynchronize(s));
  has_setup = false;
}

void Solver::assemble() {
  if (mesh.nv == 0) throw std::runtime_error("no mesh");
  FSB_CUDA(cudaSetDevice(ctx.device));
  cudaStream_t s = ctx.stream;
  tic("pattern");
  build_pattern(ctx, mesh, pat);
  toc("pattern");
  tic("assemble");
  A0.nrows = A0.ncols = pat.n; A0.nnz = pat.nnz;
  A0.ptr.alloc(pat.n + 1, s); A0.col.alloc(pat.nnz, s); A0.val.alloc(pat.nnz, s);
  A0.ptr.from_device(pat.ptr, pat.n + 1);
  A0.col.from_device(pat.col, pat.nnz);
  assemble_values(ctx, mesh, pat, A0.val);
  toc("assemble");
  // the gather lists are only needed while assembling
  pat.contrib.release(); pat.seg.release();
  custom_matrix = false;
  has_setup = false;
}

void Solver::get_matrix(int* ptr, int* col, double* val) {
  if (A0.nrows == 0) throw std::invalid_argument("Error no matrix specified");
  if (ptr) A0.ptr.to_host(ptr, A0.nrows + 1);
  if (col) A0.col.to_host(col, A0.nnz);
  if (val) A0.val.to_host(val, A0.nnz);
}

void Solver::set_matrix_values(const double* val, bool on_device) {
  if (A0.nrows == 0) throw std::invalid_argument("Error no matrix specified");
  FSB_CUDA(cudaMemcpyAsync(A0.val.get(), val, sizeof(double) * A0.nnz, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx.stream));
  FSB_CUDA(cudaStreamSynchronize(ctx.stream));
  has_setup = false;
}

void Solver::set_matrix_csr(int n, int nnz, const int* ptr, const int* col, const double* val) {
  if (pat.n != 0 && n != pat.n) throw std::invalid_argument("matrix size does not match the mesh");
  cudaStream_t s = ctx.stream;
  A0.nrows = A0.ncols = n; A0.nnz = nnz;
  A0.ptr.alloc(n + 1, s); A0.col.alloc(nnz, s); A0.val.alloc(nnz, s);
  A0.ptr.from_host(ptr, n + 1); A0.col.from_host(col, nnz); A0.val.from_host(val, nnz);
  FSB_CUDA(cudaStreamSynchronize(s));
  custom_matrix = true;
  has_setup = false;
}

// ----------------------------------------------------------------------------- stage 2
__global__ static void pstart_kernel(int nparts, const int* __restrict__ partitionIdx, const int* __restrict__ aggregateIdx, int* __restrict__ pstart) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p <= nparts) pstart[p] = aggregateIdx[partitionIdx[p]];
}
__global__ static void part_rows_kernel(int nparts, const int* __restrict__ pstart, int* __restrict__ rows) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nparts) rows[p] = pstart[p + 1] - pstart[p];
}

void Solver::setup() {
  if (A0.nrows == 0) throw std::invalid_argument("Error no matrix specified");  // FEMSolver::checkMatrixForValidContents
  if (pat.n == 0) throw std::runtime_error("setup needs the mesh graph: call assemble() first");
  if (prm.aggregatorType < 0 || prm.aggregatorType > 2)
    throw std::invalid_argument("aggregatorType_ 0 (OldMIS), 1 (METIS bottom-up) and 2 (METIS top-down = the MIS pipeline upstream) are implemented; "
                                "3-5 (AggMIS) are out of scope (SURVEY 8f-2)");
  if (prm.dsType != 0) throw std::invalid_argument("only dsType_ 0 is implemented (dsType_ 1 cannot run upstream either)");
  if (prm.aggregatorType != 1 && prm.partitionMaxSize > 1024) throw std::invalid_argument("partitionMaxSize_ must be <= 1024");
  FSB_CUDA(cudaSetDevice(ctx.device));
  cudaStream_t s = ctx.stream;
  dist_disconnect();
  destroy_graph();
  levels.clear();
  tic("setup");
  levels.emplace_back();
  {
    LevelData& L = levels[0];
    L.n = A0.nrows; L.level_id = 0;
    L.A.nrows = L.A.ncols = A0.nrows; L.A.nnz = A0.nnz;
    L.A.ptr.alloc(A0.nrows + 1, s); L.A.col.alloc(A0.nnz, s); L.A.val.alloc(A0.nnz, s);
    L.A.ptr.from_device(A0.ptr, A0.nrows + 1); L.A.col.from_device(A0.col, A0.nnz); L.A.val.from_device(A0.val, A0.nnz);
    graph_from_pattern(ctx, pat.n, pat.ptr, pat.col, L.xadj, L.adj);
  }
  int num_levels = 1;
  auto now_ms = [&]() { cudaStreamSynchronize(s); return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  for (const char* k : {"setup_aggregation", "setup_permute_split", "setup_prolongator", "setup_galerkin", "setup_coarse_inverse"}) times_ms[k] = 0;
  while (true) {
    LevelData& L = levels.back();
    int N = L.A.nrows;
    double t0 = now_ms();
    if (prm.verbose) printf("Rows: %d of max: %d\n", N, prm.topSize);
    if (N < prm.topSize || num_levels >= prm.maxLevels) {  // amg.cu:101
      if (N > 4096) throw std::runtime_error("coarsest level too large for the dense inverse");
      dense_inverse(ctx, L.A, Ainv);
      times_ms["setup_coarse_inverse"] += now_ms() - t0;
      break;
    }
    // createNextLevel (smoothedMG_amg_level.cu:402-494)
    compute_permutation(ctx, N, L.xadj, L.adj, prm.aggregatorType, prm.randMisParameters, prm.partitionMaxSize, prm.seed, L.agg);
    times_ms["setup_aggregation"] += now_ms() - t0; t0 = now_ms();
    L.nnout = L.agg.nAgg; L.nparts = L.agg.nParts;
    L.pstart.alloc(L.nparts + 1, s);
    pstart_kernel<<<cdiv(L.nparts + 1, 256), 256, 0, s>>>(L.nparts, L.agg.partitionIdx, L.agg.aggregateIdx, L.pstart);
    {
      IBuf rows(L.nparts, s);
      part_rows_kernel<<<cdiv(L.nparts, 256), 256, 0, s>>>(L.nparts, L.pstart, rows);
      L.maxPartRows = reduce_max_i32(rows, L.nparts, s);
    }
    if (L.maxPartRows > 1024) throw std::runtime_error("largest block size is larger than shared size");  // gauss_seidel.cu:2059-2064
    {
      DCsr B;
      permute_csr(ctx, L.A, L.agg.permutation, B);
      L.A.ptr.swap(B.ptr); L.A.col.swap(B.col); L.A.val.swap(B.val);
    }
    L.diag.alloc(N, s);
    extract_diag(ctx, L.A, L.diag);
    split_partitions(ctx, L);
    times_ms["setup_permute_split"] += now_ms() - t0; t0 = now_ms();
    build_prolongator(ctx, L.A, L.diag, L.agg.aggregateIdx, L.nnout, prm.proOmega, L.P);
    transpose_csr(ctx, L.P, L.R);
    static const int sell_min_rows = getenv("FSB_SELL_MINROWS") ? atoi(getenv("FSB_SELL_MINROWS")) : 32768;  // tuning knob
    // streaming copies for the levels where bandwidth (not latency) matters; levels with long rows keep the warp-per-row kernels
    if (N >= sell_min_rows && (double)L.A.nnz <= 24.0 * N) {
      if (L.level_id == 0) build_sell(ctx, L.A, L.sA);
      build_sell(ctx, L.Aout, L.sAout, 1024);  // rows have 0..8 inter-partition entries: sort by length inside 1024-row windows
      // prolongator rows have 1..8 entries (one per adjacent aggregate): length-sorted windows keep the slices tight
      static const int sp_window = getenv("FSB_SP_WINDOW") ? atoi(getenv("FSB_SP_WINDOW")) : 1024;  // tuning knob, 0: unsorted
      build_sell(ctx, L.P, L.sP, sp_window);
    }
    times_ms["setup_prolongator"] += now_ms() - t0; t0 = now_ms();
    DCsr AP, Ac;
    spgemm(ctx, L.A, L.P, AP);
    { double t1 = now_ms(); times_ms["setup_galerkin_AP_L" + std::to_string(L.level_id)] = t1 - t0; }
    { double t1 = now_ms(); spgemm(ctx, L.R, AP, Ac); times_ms["setup_galerkin_RAP_L" + std::to_string(L.level_id)] = now_ms() - t1; }
    times_ms["setup_galerkin"] += now_ms() - t0;
    // levels >= 1 above the dense tail (replicated in a sharded solve, latency-bound): the restricted residual
    // R (b - A x) is evaluated as R b - (R A) x, one kernel instead of the A_out update + restriction, and the
    // smoother skips its residual pass.  R A is formed explicitly (no symmetry assumption on a user-supplied A).
    static const bool fused_restrict = !(getenv("FSB_FUSED_RESTRICT") && atoi(getenv("FSB_FUSED_RESTRICT")) == 0);  // tuning knob
    if (fused_restrict && L.level_id >= 1 && N > kDenseTailMaxRows) {
      double t1 = now_ms();
      DCsr At, AtP;  // R A = (A^T P)^T: short rows on the left keep the product cheap (R has ~100 entries per row)
      transpose_csr(ctx, L.A, At);
      spgemm(ctx, At, L.P, AtP);
      transpose_csr(ctx, AtP, L.RA);
      times_ms["setup_RA_L" + std::to_string(L.level_id)] = now_ms() - t1;
    }
    L.b.alloc(N, s); L.x.alloc(N, s); L.x2.alloc(N, s); L.r.alloc(N, s);
    L.bc.alloc(L.nnout, s); L.xc.alloc(L.nnout, s);
    LevelData nx;
    nx.n = L.nnout; nx.level_id = num_levels;
    nx.A.nrows = Ac.nrows; nx.A.ncols = Ac.ncols; nx.A.nnz = Ac.nnz;
    nx.A.ptr.swap(Ac.ptr); nx.A.col.swap(Ac.col); nx.A.val.swap(Ac.val);
    nx.xadj.alloc(L.agg.xadjOut.size(), s); nx.adj.alloc(L.agg.adjOut.size(), s);
    nx.xadj.from_device(L.agg.xadjOut, L.agg.xadjOut.size());
    nx.adj.from_device(L.agg.adjOut, L.agg.adjOut.size());
    if (prm.verbose) printf("level %d: rows %d nnz %d aggregates %d partitions %d (largest %d rows)\n", L.level_id, N, L.A.nnz, L.nnout, L.nparts, L.maxPartRows);
    levels.push_back(std::move(nx));
    num_levels++;
  }
  {
    double t0 = now_ms();
    build_dense_tail();
    times_ms["setup_dense_tail"] = now_ms() - t0;
    t0 = now_ms();
    build_block_smoothers();
    times_ms["setup_block_smoothers"] = now_ms() - t0;
  }
  toc("setup");
  has_setup = true;
}

// ----------------------------------------------------------------------------- stage 3
// One V-cycle on level `lev`.  b_src is read through `gather` (external -> internal numbering,
// null when b_src is already internal); the result goes to x_dst in internal numbering
// (scatter == null) or is scattered to the external numbering through `scatter`.
// one exchange of a sharded solve: my list entries of `src` to the peers, theirs into `dst` (cycle.cu: ll_exchange_kernel)
void Solver::exchange_chan(int chan, const double* src, double* dst, const int* done, bool side_stream) {
  if (!side_stream) { launch_ll_exchange(ctx, dist.chan[chan].dev, src, dst, done); return; }
  // fork: the exchange runs on its own stream behind everything enqueued so far; the caller joins with ev_xjoin
  FSB_CUDA(cudaEventRecord(ctx.ev_xfork, ctx.stream));
  FSB_CUDA(cudaStreamWaitEvent(ctx.xstream, ctx.ev_xfork, 0));
  Ctx cx = ctx;
  cx.stream = ctx.xstream;
  launch_ll_exchange(cx, dist.chan[chan].dev, src, dst, done);
  FSB_CUDA(cudaEventRecord(ctx.ev_xjoin, ctx.xstream));
}

double Solver::dist_bench_exchange(int chan, int reps) {
  if (!dist.connected || dist.nranks < 2) throw std::runtime_error("not connected");
  cudaStream_t s = ctx.stream;
  LevelData& L0 = levels[0];
  if (scal.size() != 1) scal.alloc(1, s);
  if (partials.size() < 4096) partials.alloc(4096, s);
  launch_cg_init(ctx, scal.get(), 0.0, 1 << 30, 2);
  auto one = [&]() {
    if (chan >= 0) {
      if (chan >= dist.nchan || !dist.chan[chan].list) throw std::invalid_argument("no such channel");
      const int lev = chan < kChanLevel0 ? 0 : (chan - kChanLevel0) / kChanPerLevel;
      const int which = chan < kChanLevel0 ? kXPre : (chan - kChanLevel0) % kChanPerLevel;
      LevelData& L = levels[lev];
      double* v = (which == kDown) ? L.bc.get() : (which == kUp) ? L.xc.get() : (which == kRes) ? L.r.get() : L.x.get();
      exchange_chan(chan, v, v, nullptr);
    } else {
      launch_dot(ctx, 1024, L0.x, L0.x, partials, scal.get(), 1);
    }
  };
  for (int i = 0; i < 5; i++) one();
  FSB_CUDA(cudaStreamSynchronize(s));
  FSB_CUDA(cudaEventRecord(ev0_, s));
  for (int i = 0; i < reps; i++) one();
  FSB_CUDA(cudaEventRecord(ev1_, s));
  FSB_CUDA(cudaEventSynchronize(ev1_));
  float ms = 0;
  FSB_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
  return ms * 1e3 / reps;
}

// An exchange of vector v followed by its consumer kernel over `full` rows.  With an interior range (rows that reference
// nothing the exchange delivers) the exchange runs on its own stream while the consumer works on the interior rows; the
// boundary rows follow after the join.  Without one (or when `on` is false: no exchange at all) the consumer runs once.
template <typename F>
void Solver::with_exchange(bool on, int lev, int which, double* v, RowRange full, RowRange interior, const int* done, F&& consumer) {
  if (!on) { consumer(full); return; }
  if (interior.end <= interior.begin) { exchange(lev, which, v, v, done); consumer(full); return; }
  exchange_chan(kChanLevel0 + kChanPerLevel * lev + which, v, v, done, /*side_stream=*/true);
  consumer(interior);
  FSB_CUDA(cudaStreamWaitEvent(ctx.stream, ctx.ev_xjoin, 0));
  if (full.begin < interior.begin) { RowRange q; q.begin = full.begin; q.end = interior.begin; consumer(q); }
  if (interior.end < full.end) { RowRange q; q.begin = interior.end; q.end = full.end; consumer(q); }
}

void Solver::vcycle(int lev, const double* b_src, const int* gather, double* x_dst, const int* scatter, double* /*unused*/) {
  LevelData& L = levels[lev];
  const int* done = cg_active_ ? &scal.get()->done : nullptr;
  profiler.cur_level = lev;
  if (lev == (int)levels.size() - 1) {  // coarsest: direct solve (amg_level.cu:25-31)
    launch_coarse_solve(ctx, L.n, Ainv, b_src, x_dst, done);
    return;
  }
  if (lev == tail_level_) {  // this level and everything below it: one dense GEMV (dense_tail.cu)
    launch_coarse_solve(ctx, L.n, Mtail, b_src, x_dst, done);
    return;
  }
  const double w = prm.smootherWeight;
  const double* b_eff = gather ? L.b.get() : b_src;
  // sharded solve: this GPU owns a contiguous range of the level's partitions (dist.cu)
  const bool D = sharded(lev);
  const bool Dn = D && sharded(lev + 1);
  RowRange rr, rrc;
  const DistLevel* DL = D ? &dist.lev[lev] : nullptr;
  if (D) {
    if (prm.postRelaxes != 1) throw std::invalid_argument("the sharded solve supports postRelaxes_ == 1");
    rr.begin = DL->rbeg[dist.rank]; rr.end = DL->rbeg[dist.rank + 1];
    rrc.begin = DL->abeg[dist.rank]; rrc.end = DL->abeg[dist.rank + 1];
  }
  if (!D && L.RA.nrows > 0) {
    // pre: x = w b/d, nu1 sweeps; then bc = R (b - A x) = R b - (R A) x without forming the residual
    launch_smooth(ctx, L, b_src, gather, gather ? L.b.get() : nullptr, nullptr, w, prm.preInnerIters, L.x, nullptr, nullptr, nullptr, done, false);
    launch_restrict_fused(ctx, L.R, b_eff, L.RA, L.x, L.bc, done);
  } else {
    // pre: x = w b/d, nu1 sweeps, r = b - A_in x - d x  (one kernel, matrix slab read once)
    launch_smooth(ctx, L, b_src, gather, gather ? L.b.get() : nullptr, nullptr, w, prm.preInnerIters, L.x, nullptr, nullptr, L.r, done, D);
    // x across the cut, overlapped with the interior rows of its consumer: r -= A_out x   (preAout_kernel)
    with_exchange(D, lev, kXPre, L.x, rr, D ? DL->intA : RowRange(), done, [&](RowRange q) {
      if (L.sAout.ready()) launch_spmv_sell(ctx, L.sAout, L.x, L.r, 3, nullptr, done, "residual_out", q);
      else launch_spmv(ctx, L.Aout, L.x, L.r, 3, nullptr, done, "residual_out", q);
    });
    // r rows the peers restrict, overlapped with the interior coarse rows: bc = R r
    with_exchange(D, lev, kRes, L.r, rrc, D ? DL->intR : RowRange(), done, [&](RowRange q) {
      launch_spmv(ctx, L.R, L.r, L.bc, 0, nullptr, done, "restrict", q);
    });
  }
  if (D) {
    // restricted residual: to the owners of the next level's rows, or — next level replicated — all-gathered
    exchange(lev, kDown, L.bc, L.bc, done);
  }
  const bool next_is_coarsest = (lev + 1 == (int)levels.size() - 1);
  const int* ip = next_is_coarsest ? nullptr : levels[lev + 1].agg.ipermutation.get();
  vcycle(lev + 1, L.bc, ip, L.xc, ip, nullptr);
  profiler.cur_level = lev;
  // coarse corrections of the next level's rows I own that the peers' prolongator rows (and ghost rows) reference,
  // overlapped with the rows whose prolongator entries are all mine: x += P xc
  with_exchange(Dn, lev, kUp, L.xc, rr, Dn ? DL->intP : RowRange(), done, [&](RowRange q) {
    if (L.sP.ready()) launch_spmv_sell(ctx, L.sP, L.xc, L.x, 2, nullptr, done, "prolong_add", q);
    else launch_spmv(ctx, L.P, L.xc, L.x, 2, nullptr, done, "prolong_add", q);
  });
  // sharded level: the ghost copies of x (they hold the neighbours' x after pre-smoothing) get the same correction here,
  // bit-identical to what their owners compute — no halo exchange after the prolongation
  static const bool ghost_prolong = !(getenv("FSB_GHOST_PROLONG") && atoi(getenv("FSB_GHOST_PROLONG")) == 0);  // tuning knob
  if (D && ghost_prolong) launch_spmv_list_add(ctx, L.P, DL->sendA.ridx, DL->sendA.rtotal, L.xc, L.x, done);
  double* xin = L.x;
  double* xtmp = L.x2;
  for (int rel = 0; rel < prm.postRelaxes; rel++) {
    bool lastpass = (rel == prm.postRelaxes - 1);
    if (D && !ghost_prolong) exchange(lev, kXPost, xin, xin, done);
    if (L.sAout.ready()) launch_spmv_sell(ctx, L.sAout, xin, L.r, 1, b_eff, done, "bprime", rr);
    else launch_spmv(ctx, L.Aout, xin, L.r, 1, b_eff, done, "bprime", rr);         // b' = b - A_out x (x frozen for this pass)
    if (lastpass) launch_smooth(ctx, L, L.r, nullptr, nullptr, xin, w, prm.postInnerIters, scatter ? nullptr : x_dst, scatter, scatter ? x_dst : nullptr, nullptr, done, D);
    else { launch_smooth(ctx, L, L.r, nullptr, nullptr, xin, w, prm.postInnerIters, xtmp, nullptr, nullptr, nullptr, done, D); std::swap(xin, xtmp); }
  }
  if (prm.postRelaxes <= 0) {  // degenerate configuration: no post-relaxation pass
    if (scatter) launch_scatter(ctx, L.n, scatter, xin, x_dst);
    else FSB_CUDA(cudaMemcpyAsync(x_dst, xin, sizeof(double) * L.n, cudaMemcpyDeviceToDevice, ctx.stream));
  }
}

void Solver::apply_matrix(const double* x, double* y) {
  if (A0.nrows == 0) throw std::invalid_argument("Error no matrix specified");
  launch_spmv(ctx, A0, x, y, 0, nullptr, nullptr, "apply_matrix");
  FSB_CUDA(cudaStreamSynchronize(ctx.stream));
}

void Solver::spmv_fine(const double* x, double* y) { launch_spmv(ctx, levels.at(0).A, x, y, 0, nullptr, nullptr, "spmv"); FSB_CUDA(cudaStreamSynchronize(ctx.stream)); }

// the dense tail and the dense partition blocks encode the smoother parameters: rebuild them when they changed since setup()
void Solver::ensure_dense_tail() {
  bool dense = tail_level_ >= 0;
  for (const auto& L : levels) dense = dense || L.use_blockdense;
  // (also when the blocks were skipped because of postRelaxes_ != 1 and it is 1 now)
  if ((dense || tail_key_[2] != prm.postRelaxes) && (tail_key_[0] != prm.preInnerIters || tail_key_[1] != prm.postInnerIters || tail_key_[2] != prm.postRelaxes ||
                           tail_key_[3] != prm.smootherWeight)) {
    build_dense_tail();
    build_block_smoothers();
    destroy_graph();
  }
}

void Solver::precondition(const double* r, double* z) {
  if (!has_setup) throw std::runtime_error("precondition before setup");
  ensure_dense_tail();
  cg_active_ = false;
  vcycle(0, r, nullptr, z, nullptr, nullptr);
  FSB_CUDA(cudaStreamSynchronize(ctx.stream));
}

void Solver::enqueue_pcg_iteration() {
  const int n = levels[0].n;
  PcgScalars* sc = scal.get();
  const bool D = sharded(0);
  const int rb = D ? dist.lev[0].rbeg[dist.rank] : 0, re = D ? dist.lev[0].rbeg[dist.rank + 1] : n, nown = re - rb;
  RowRange rr;
  if (D) { rr.begin = rb; rr.end = re; }
  // y = A p, alpha = rz / (p.y).  Sharded: p crosses the cut first — on EVERY rank at this point of the sequence (an
  // exchange that one rank issues before and another after an all-reduce would be a circular wait) — and, where this GPU
  // has an interior range, WHILE the interior rows are multiplied.
  const bool overlap_p = D && dist.lev[0].intA.end > dist.lev[0].intA.begin;
  if (overlap_p) {
    const RowRange in = dist.lev[0].intA;
    exchange_chan(kChanP, cg_p, cg_p, &sc->done, /*side_stream=*/true);
    int parked = launch_spmv_dot_sell_part(ctx, levels[0].sA, cg_p, cg_y, partials, sc, in, 0, false);
    FSB_CUDA(cudaStreamWaitEvent(ctx.stream, ctx.ev_xjoin, 0));
    RowRange lo, hi;
    lo.begin = rb; lo.end = in.begin; hi.begin = in.end; hi.end = re;
    parked += launch_spmv_dot_sell_part(ctx, levels[0].sA, cg_p, cg_y, partials, sc, lo, parked, false);
    launch_spmv_dot_sell_part(ctx, levels[0].sA, cg_p, cg_y, partials, sc, hi, parked, true);
  } else {
    if (D) exchange_chan(kChanP, cg_p, cg_p, &sc->done);
    if (levels[0].sA.ready()) launch_spmv_dot_sell(ctx, levels[0].sA, cg_p, cg_y, partials, sc, rr);
    else launch_spmv_dot(ctx, levels[0].A, cg_p, cg_y, partials, sc);
  }
  launch_cg_update(ctx, nown, cg_x.get() + rb, cg_r.get() + rb, cg_p.get() + rb, cg_y.get() + rb, partials, sc, hist);  // x += alpha p, r -= alpha y, ||r||, test
  vcycle(0, cg_r, nullptr, cg_z, nullptr, nullptr);                          // z = M^-1 r
  launch_dot(ctx, nown, cg_r.get() + rb, cg_z.get() + rb, partials, sc, 2);  // rz_new, beta
  launch_cg_pdir(ctx, nown, cg_p.get() + rb, cg_z.get() + rb, sc, 0);        // p = z + beta p
}

void Solver::pcg(const double* b_user, double* x_user) {
  cudaStream_t s = ctx.stream;
  LevelData& L0 = levels[0];
  const int n = L0.n;
  const bool single = levels.size() == 1;
  const bool permute = !single && !prm.refLevel0NoPerm;
  // work vectors persist across solves so a captured iteration graph stays valid
  auto ensure = [&](DBuf& v, size_t m) { if (v.size() != m) v.alloc(m, s); };
  ensure(cg_b, n); ensure(cg_x, n); ensure(cg_r, n); ensure(cg_z, n); ensure(cg_p, n); ensure(cg_y, n);
  size_t npart = std::max<size_t>((size_t)cdiv((long long)n * 32, 256), (size_t)ctx.num_sms * 8) + 1;
  ensure(partials, npart);
  const int hist_cap = std::max(2, std::min(prm.maxIters, 1 << 20) + 2);  // maxIters_ used as 'infinite' must not size an allocation
  ensure(hist, (size_t)hist_cap);
  if (scal.size() != 1) scal.alloc(1, s);
  PcgScalars* sc = scal.get();
  cg_active_ = true;
  GraphKey key;
  memset(&key, 0, sizeof key);  // the struct has padding and is compared bytewise
  {
    void* ptrs[9] = {cg_b.get(), cg_x.get(), cg_r.get(), cg_z.get(), cg_p.get(), cg_y.get(), partials.get(), hist.get(), (void*)sc};
    memcpy(key.p, ptrs, sizeof ptrs);
    key.pre = prm.preInnerIters; key.post = prm.postInnerIters; key.relaxes = prm.postRelaxes; key.w = prm.smootherWeight;
  }
  if (iter_graph_ && memcmp(&key, &graph_key_, sizeof(GraphKey)) != 0) destroy_graph();
  graph_key_ = key;
  launch_cg_init(ctx, sc, prm.tolerance, prm.maxIters, hist_cap);
  const bool D = sharded(0);
  const int rb = D ? dist.lev[0].rbeg[dist.rank] : 0, re = D ? dist.lev[0].rbeg[dist.rank + 1] : n, nown = re - rb;
  // the whole iteration lives in the level-0 permuted numbering; a sharded solve touches its own rows only
  if (permute) {
    launch_gather(ctx, nown, L0.agg.ipermutation.get() + rb, b_user, cg_b.get() + rb);
    launch_gather(ctx, nown, L0.agg.ipermutation.get() + rb, x_user, cg_x.get() + rb);
  } else {
    FSB_CUDA(cudaMemcpyAsync(cg_b.get() + rb, b_user + rb, sizeof(double) * nown, cudaMemcpyDeviceToDevice, s));
    FSB_CUDA(cudaMemcpyAsync(cg_x.get() + rb, x_user + rb, sizeof(double) * nown, cudaMemcpyDeviceToDevice, s));
  }
  // bnorm first: its all-reduce is also the barrier that separates this solve's first peer stores from the
  // peers' last reads of the previous solve (final scatter of the all-gathered solution)
  launch_dot(ctx, nown, cg_b.get() + rb, cg_b.get() + rb, partials, sc, 0);
  RowRange rr;
  if (D) {
    rr.begin = rb; rr.end = re;
    exchange_chan(kChanX0, cg_x, cg_x, nullptr);  // initial guess across the cut
  }
  if (L0.sA.ready()) launch_spmv_sell(ctx, L0.sA, cg_x, cg_r, 1, cg_b, nullptr, "residual", rr);
  else launch_spmv(ctx, L0.A, cg_x, cg_r, 1, cg_b, nullptr, "residual"); // r = b - A x
  vcycle(0, cg_r, nullptr, cg_z, nullptr, nullptr);                // z = M^-1 r
  launch_cg_pdir(ctx, nown, cg_p.get() + rb, cg_z.get() + rb, sc, 1);  // p = z
  // (sharded: p crosses the cut as the first thing of every iteration)
  launch_dot(ctx, nown, cg_r.get() + rb, cg_z.get() + rb, partials, sc, 1);  // rz_old
  FSB_CUDA(cudaStreamSynchronize(s));

  profiler.cur_level = 0;
  const bool graphs = prm.useGraphs && !prm.profile;
  if (!graphs) destroy_graph();
  if (graphs && !iter_graph_) {
    cudaGraph_t g;
    long long before = g_launch_counter;
    FSB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    enqueue_pcg_iteration();
    FSB_CUDA(cudaStreamEndCapture(s, &g));
    FSB_CUDA(cudaGraphInstantiate(&iter_graph_, g, 0));
    FSB_CUDA(cudaGraphDestroy(g));
    iter_launches_ = g_launch_counter - before;  // kernels of one iteration (counted while capturing)
    g_launch_counter = before;
  }
  const long long per_iter = iter_launches_;
  // Convergence polling without pipeline bubbles: chunk i+1 is enqueued before the host waits for the scalars
  // of chunk i (pinned double buffer + events), so the GPU never idles for a host round trip; once the
  // device has set `done`, the kernels of an already enqueued chunk exit at their first instruction.
  PcgScalars h;
  int enq = 0;
  const int chunk = std::max(1, prm.checkEvery);
  auto enqueue_chunk = [&](int slot) {
    for (int i = 0; i < chunk; i++) {
      if (iter_graph_) { FSB_CUDA(cudaGraphLaunch(iter_graph_, s)); g_launch_counter += per_iter; }
      else enqueue_pcg_iteration();
      enq++;
    }
    FSB_CUDA(cudaMemcpyAsync(&poll_host_[slot], sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, s));
    FSB_CUDA(cudaEventRecord(poll_ev_[slot], s));
  };
  int slot = 0;
  enqueue_chunk(slot);
  while (true) {
    const bool more = enq <= prm.maxIters + chunk;
    if (more) enqueue_chunk(slot ^ 1);
    FSB_CUDA(cudaEventSynchronize(poll_ev_[slot]));
    h = poll_host_[slot];
    if (h.done || !more) break;
    slot ^= 1;
  }
  FSB_CUDA(cudaStreamSynchronize(s));
  iterations = h.niter;
  resid_history.resize(h.hist_len);
  if (h.hist_len) hist.to_host(resid_history.data(), h.hist_len);
  final_relres = h.hist_len ? resid_history.back() : -1;
  if (D) {
    if (h.err || dist.error.read(0)) throw std::runtime_error("sharded solve: a peer did not arrive at an exchange (timeout)");
    // every GPU ends up with the full solution
    launch_push_all(ctx, rb, re, cg_x, peers_at(dist.off_cgx));
    if (dist.error.read(0)) throw std::runtime_error("sharded solve: a peer did not arrive at the final exchange (timeout)");
  }
  if (permute) launch_scatter(ctx, n, L0.agg.ipermutation, cg_x, x_user);
  else FSB_CUDA(cudaMemcpyAsync(x_user, cg_x.get(), sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
  FSB_CUDA(cudaStreamSynchronize(s));
}

void Solver::solve(const double* b, double* x, bool on_device) {
  if (!has_setup) throw std::runtime_error("solve before setup");
  FSB_CUDA(cudaSetDevice(ctx.device));
  cudaStream_t s = ctx.stream;
  const int n = levels[0].n;
  ensure_dense_tail();
  g_launch_counter = 0;
  profiler.clear();
  profiler.on = prm.profile != 0;
  DBuf bd, xd;
  const double* bp = b;
  double* xp = x;
  // sharded PCG with host buffers: only the user-numbering slice that covers this GPU's rows crosses PCIe
  // (b and the initial guess in, the solution out); x outside [user_lo, user_hi) is left untouched on the host
  const bool slice = !on_device && dist.connected && dist.nranks > 1 && prm.solverType == 1;
  const int lo = slice ? dist.user_lo : 0, hi = slice ? dist.user_hi : n;
  if (!on_device) {
    bd.alloc(n, s); xd.alloc(n, s);
    FSB_CUDA(cudaMemcpyAsync(bd.get() + lo, b + lo, sizeof(double) * (hi - lo), cudaMemcpyHostToDevice, s));
    FSB_CUDA(cudaMemcpyAsync(xd.get() + lo, x + lo, sizeof(double) * (hi - lo), cudaMemcpyHostToDevice, s));
    bp = bd; xp = xd;
  }
  tic("solve");
  if (prm.solverType == 1) {
    pcg(bp, xp);
  } else {
    // AMG_SOLVER: exactly ONE V-cycle, whatever maxIters_/tolerance_ say (amg.cu:187-194);
    // the incoming x is overwritten (every cycle starts from a zero guess).
    cg_active_ = false;
    const bool single = levels.size() == 1;
    const int* ip = (single || prm.refLevel0NoPerm) ? nullptr : levels[0].agg.ipermutation.get();
    DBuf xo(n, s);
    vcycle(0, bp, ip, xo, ip, nullptr);
    FSB_CUDA(cudaMemcpyAsync(xp, xo.get(), sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
    iterations = 1; final_relres = -1; resid_history.clear();
  }
  toc("solve");
  launches = g_launch_counter;
  profiler.on = false;
  if (!on_device) {
    FSB_CUDA(cudaMemcpyAsync(x + lo, xd.get() + lo, sizeof(double) * (hi - lo), cudaMemcpyDeviceToHost, s));
    FSB_CUDA(cudaStreamSynchronize(s));
  }
}

std::string Solver::profile_report() {
  std::map<std::pair<std::string, int>, std::pair<long long, double>> agg;
  FSB_CUDA(cudaStreamSynchronize(ctx.stream));
  for (auto& r : profiler.recs) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) { cudaGetLastError(); continue; }
    auto& a = agg[{r.name, r.level}];
    a.first++; a.second += ms;
  }
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof line, "%s %d %lld %.6f\n", kv.first.first.c_str(), kv.first.second, kv.second.first, kv.second.second);
    out += line;
  }
  return out;
}

}  // namespace fsb
