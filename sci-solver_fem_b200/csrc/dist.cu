// dist.cu — stage 4: sharding the PCG solve over the GPUs of one box (absent upstream, F9).
//
// Design: one process per GPU.  Setup is REPLICATED (every GPU assembles the mesh and builds the
// identical, deterministic hierarchy — no setup communication at all); the solve is SHARDED on the
// fine level: the reference's partitions (<= 1024-row blocks whose rows are contiguous) are dealt out
// in contiguous, nnz-balanced ranges, i.e. a GPU is a super-partition in the reference's own
// A_in / A_out sense.  Every GPU keeps full-length vectors but computes only its rows; the values a
// peer needs (operator columns across the cut, restriction rows across the cut) are PUSHED into the
// peer's copy by plain stores through NVLink peer mappings (CUDA IPC), followed by a flag-based
// cross-GPU barrier executed by the last CTA of the push kernel.  Dot products are all-reduced the same
// way inside the reduction kernels (cycle.cu: cross_sum).  Coarse levels (>= 1) are small and are
// computed redundantly on every GPU after an all-gather (push_all) of the restricted residual.
#include <algorithm>
#include <numeric>

#include "kernels.h"
#include "solver.h"

namespace fsb {

// nnz-balanced contiguous split of `nparts` partitions with weights w[p] over nranks: out[r] = first
// partition of rank r (out[nranks] = nparts).  Pure host function (also exported for CPU tests).
void split_by_weight(int nparts, const long long* w, int nranks, int* out) {
  std::vector<long long> pre(nparts + 1, 0);
  for (int p = 0; p < nparts; p++) pre[p + 1] = pre[p] + w[p];
  out[0] = 0;
  for (int r = 1; r < nranks; r++) {
    long long target = pre[nparts] * r / nranks;
    int p = (int)(std::lower_bound(pre.begin(), pre.end(), target) - pre.begin());
    // closest boundary, monotone, and at least one partition per rank while partitions last
    if (p > 0 && target - pre[p - 1] < pre[p] - target) p--;
    p = std::max(p, out[r - 1] + (out[r - 1] < nparts ? 1 : 0));
    out[r] = std::min(p, nparts);
  }
  out[nranks] = nparts;
  for (int r = nranks - 1; r >= 1; r--) out[r] = std::min(out[r], out[r + 1]);
}

namespace {

struct Ranges { int b[kMaxRanks + 1]; int n; };
__device__ __forceinline__ int owner_of(const Ranges& rg, int i) {
  int q = 0;
  while (q + 1 < rg.n && i >= rg.b[q + 1]) q++;
  return q;
}

// flags[q * nown + (j - myb)] = 1 when a row owned by q != me references my row j
__global__ void mark_needed_kernel(int nrows, const int* __restrict__ ptr, const int* __restrict__ col, Ranges rowOwner, Ranges colOwner,
                                   int me, int myb, int mye, int* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  int q = owner_of(rowOwner, i);
  if (q == me) return;
  int nown = mye - myb;
  for (int e = ptr[i]; e < ptr[i + 1]; e++) {
    int j = col[e];
    if (j >= myb && j < mye) flags[(size_t)q * nown + (j - myb)] = 1;
  }
  (void)colOwner;
}
__global__ void fill_list_kernel(long long total, int nown, int myb, const int* __restrict__ flags, const int* __restrict__ pos, int* __restrict__ list) {
  long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k < total && flags[k]) list[pos[k] - 1] = myb + (int)(k % nown);
}

void build_send_list(const Ctx& c, int nranks, int nown, int myb, IBuf& flags, IBuf& list, IBuf& list_ptr, int& total) {
  cudaStream_t s = c.stream;
  size_t m = (size_t)nranks * nown;
  IBuf pos(m, s);
  inclusive_scan_i32(flags, pos, m, s);
  std::vector<int> ptr_h(nranks + 1, 0);
  for (int q = 0; q < nranks; q++) ptr_h[q + 1] = nown > 0 ? pos.read((size_t)(q + 1) * nown - 1) : 0;
  total = ptr_h[nranks];
  list.alloc(std::max(1, total), s);
  fill_list_kernel<<<cdiv((long long)m, 256), 256, 0, s>>>((long long)m, nown, myb, flags, pos, list);
  list_ptr.alloc(nranks + 1, s);
  list_ptr.from_host(ptr_h.data(), nranks + 1);
  FSB_CUDA(cudaStreamSynchronize(s));
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

void Solver::dist_disconnect() {
  FSB_CUDA(cudaStreamSynchronize(ctx.stream));
  destroy_graph();
  for (int q = 0; q < kMaxRanks; q++) {
    if (dist.peer[q] && q != dist.rank) cudaIpcCloseMemHandle(dist.peer[q]);
    dist.peer[q] = nullptr;
  }
  if (dist.arena) {
    // the vectors that were views into the arena become ordinary buffers again on the next solve
    cg_p.release(); cg_x.release();
    if (!levels.empty()) {
      LevelData& L0 = levels[0];
      cudaStream_t s = ctx.stream;
      L0.x.release(); L0.r.release(); L0.bc.release();
      L0.x.alloc(L0.n, s); L0.r.alloc(L0.n, s); L0.bc.alloc(std::max(1, L0.nnout), s);
    }
    cudaFree(dist.arena);
    dist.arena = nullptr;
  }
  dist.connected = false;
  dist.nranks = 1; dist.rank = 0;
  ctx.dist = DistDev();
  ctx.dist_ticket = nullptr;
}

void Solver::dist_prepare(int rank, int nranks) {
  if (!has_setup) throw std::runtime_error("dist_prepare before setup");
  if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) throw std::invalid_argument("bad rank / nranks (at most 8 GPUs of one box)");
  dist_disconnect();
  dist.rank = rank; dist.nranks = nranks;
  if (nranks == 1) return;
  if (levels.size() < 2 || !levels[0].use_ell || !levels[0].sA.ready() || !levels[0].sAout.ready() || !levels[0].sP.ready())
    throw std::runtime_error("sharded solve needs a fine level with >= 32768 rows on the ELL/SELL path");
  cudaStream_t s = ctx.stream;
  LevelData& L0 = levels[0];
  const int np = L0.nparts, n = L0.n, n1 = L0.nnout;
  // 1. nnz-balanced contiguous ranges of partitions
  std::vector<int> ps = L0.pstart.to_vector(), pidx = L0.agg.partitionIdx.to_vector();
  std::vector<int> rowptr = L0.A.ptr.to_vector();
  std::vector<long long> w(np);
  for (int p = 0; p < np; p++) w[p] = rowptr[ps[p + 1]] - rowptr[ps[p]];
  split_by_weight(np, w.data(), nranks, dist.pbeg);
  for (int r = 0; r <= nranks; r++) { dist.rbeg[r] = ps[dist.pbeg[r]]; dist.abeg[r] = pidx[dist.pbeg[r]]; }
  const int myb = dist.rbeg[rank], mye = dist.rbeg[rank + 1], nown = mye - myb;
  // 2. owned partition lists of the smoother's size classes
  {
    std::vector<EllDesc> lists[kEllClasses];
    for (int p = dist.pbeg[rank]; p < dist.pbeg[rank + 1]; p++) lists[ell_class(ps[p + 1] - ps[p])].push_back(L0.ellDescHost[p]);
    for (int q = 0; q < kEllClasses; q++) {
      L0.nlistOwn[q] = (int)lists[q].size();
      L0.plistOwn[q].alloc(std::max<size_t>(1, lists[q].size()), s);
      if (!lists[q].empty()) L0.plistOwn[q].from_host(lists[q].data(), lists[q].size());
    }
  }
  // 3. send lists: my rows that a peer's operator rows (A) / restriction rows (R) reference
  Ranges rowsR, aggsR;
  rowsR.n = aggsR.n = nranks;
  for (int r = 0; r <= nranks; r++) { rowsR.b[r] = dist.rbeg[r]; aggsR.b[r] = dist.abeg[r]; }
  {
    IBuf flags((size_t)nranks * std::max(nown, 1), s);
    flags.zero();
    mark_needed_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, L0.A.ptr, L0.A.col, rowsR, rowsR, rank, myb, mye, flags);
    build_send_list(ctx, nranks, nown, myb, flags, dist.sendA, dist.sendA_ptr, dist.nSendA);
    flags.zero();
    mark_needed_kernel<<<cdiv(n1, 256), 256, 0, s>>>(n1, L0.R.ptr, L0.R.col, aggsR, rowsR, rank, myb, mye, flags);
    build_send_list(ctx, nranks, nown, myb, flags, dist.sendR, dist.sendR_ptr, dist.nSendR);
  }
  // 4. arena (identical layout on every rank): flags | reduction slots | p | x | r | bc | cg_x
  size_t off = 0;
  dist.off_flags = off; off = align_up(off + kMaxRanks * sizeof(unsigned long long), 256);
  dist.off_red = off; off = align_up(off + 2 * kMaxRanks * sizeof(double), 256);
  dist.off_p = off; off = align_up(off + (size_t)n * 8, 256);
  dist.off_x = off; off = align_up(off + (size_t)n * 8, 256);
  dist.off_r = off; off = align_up(off + (size_t)n * 8, 256);
  dist.off_bc = off; off = align_up(off + (size_t)std::max(n1, 1) * 8, 256);
  dist.off_cgx = off; off = align_up(off + (size_t)n * 8, 256);
  dist.arena_bytes = off;
  FSB_CUDA(cudaMalloc((void**)&dist.arena, dist.arena_bytes));
  FSB_CUDA(cudaMemsetAsync(dist.arena, 0, dist.arena_bytes, s));
  cg_p.view(reinterpret_cast<double*>(dist.arena + dist.off_p), n, s);
  cg_x.view(reinterpret_cast<double*>(dist.arena + dist.off_cgx), n, s);
  L0.x.view(reinterpret_cast<double*>(dist.arena + dist.off_x), n, s);
  L0.r.view(reinterpret_cast<double*>(dist.arena + dist.off_r), n, s);
  L0.bc.view(reinterpret_cast<double*>(dist.arena + dist.off_bc), n1, s);
  dist.epoch.alloc(1, s); dist.epoch.zero();
  dist.error.alloc(1, s); dist.error.zero();
  dist.ticket.alloc(1, s); dist.ticket.zero();
  FSB_CUDA(cudaStreamSynchronize(s));
}

void Solver::dist_get_handle(void* handle64, long long* bytes) {
  if (!dist.arena) throw std::runtime_error("dist_get_handle before dist_prepare");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  FSB_CUDA(cudaIpcGetMemHandle(&h, dist.arena));
  memcpy(handle64, &h, 64);
  if (bytes) *bytes = (long long)dist.arena_bytes;
}

void Solver::dist_connect(const void* handles) {
  if (!dist.arena) throw std::runtime_error("dist_connect before dist_prepare");
  const char* hb = static_cast<const char*>(handles);
  for (int q = 0; q < dist.nranks; q++) {
    if (q == dist.rank) { dist.peer[q] = dist.arena; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, hb + 64 * q, 64);
    void* p = nullptr;
    FSB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    dist.peer[q] = static_cast<char*>(p);
  }
  DistDev d;
  d.rank = dist.rank; d.nranks = dist.nranks;
  d.my_flags = reinterpret_cast<unsigned long long*>(dist.arena + dist.off_flags);
  d.my_red = reinterpret_cast<double*>(dist.arena + dist.off_red);
  for (int q = 0; q < dist.nranks; q++) {
    d.peer_flags[q] = reinterpret_cast<unsigned long long*>(dist.peer[q] + dist.off_flags);
    d.peer_red[q] = reinterpret_cast<double*>(dist.peer[q] + dist.off_red);
  }
  d.epoch = dist.epoch.get();
  d.error = dist.error.get();
  ctx.dist = d;
  ctx.dist_ticket = dist.ticket.get();
  dist.connected = true;
  destroy_graph();
}

}  // namespace fsb
