// dist.cu — stage 4: sharding the PCG solve over the GPUs of one box (absent upstream, F9).
//
// Design: one process per GPU.  Setup is REPLICATED (every GPU assembles the mesh and builds the
// identical, deterministic hierarchy — no setup communication at all); the solve is SHARDED level by
// level: on every level that is large enough the reference's partitions (<= 1024-row blocks whose rows
// are contiguous in the level's permuted numbering) are dealt out in contiguous, nnz-balanced ranges, i.e.
// a GPU is a super-partition in the reference's own A_in / A_out sense (smoothedMG_amg_level.cu:199-304).
// Every GPU keeps full-length vectors but computes only its rows.  What a peer needs —
//   * operator columns across the cut (x after pre-smoothing, x after the coarse correction, the PCG
//     direction p),
//   * residual rows its restriction rows reference,
//   * restricted-residual entries of rows it owns on the next level ("down"), coarse corrections its
//     prolongator rows reference ("up"),
// is sent by ONE small kernel per exchange with a flag-in-data protocol (cycle.cu: ll_exchange_kernel; the idea of
// NCCL's LL protocol): every value travels as a 16-byte {lo, epoch, hi, epoch} store into the receiver's buffer over
// the NVLink peer mapping (CUDA IPC), the receiver polls its slots — no fence, no flag word, no barrier, and only
// the GPUs that actually share a cut talk to each other.  Dot products are all-reduced the same way inside the
// reduction kernels (cycle.cu: cross_sum).  The levels below the sharded ones are small and are computed
// redundantly on every GPU after an all-gather of the restricted residual — the coarse grid is agglomerated onto
// every GPU instead of onto one, so the way back up needs no communication.
#include <algorithm>
#include <climits>
#include <cstring>
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

// Generic "who needs which of my values" marking.  Row i (owner by rowOwner) references the entries
// col[ptr[i] .. ptr[i+1]) (ptr == null: the single entry col[i]); an entry j is first mapped through
// colmap when given (e.g. external -> internal numbering); it is MINE when it falls into [myb, mye).
// flags[q * nown + (j - myb)] = 1 when a row owned by q != me references my value j.
// rowmask (optional): bit q set <=> rank q consumes row i (its owner AND the ranks that hold it as a ghost row); without
// it the only consumer of a row is its owner.
__global__ void mark_needed_kernel(int nrows, const int* __restrict__ ptr, const int* __restrict__ col, const int* __restrict__ colmap,
                                   Ranges rowOwner, const unsigned* __restrict__ rowmask, int me, int myb, int mye,
                                   int* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  unsigned consumers = rowmask ? rowmask[i] : (1u << owner_of(rowOwner, i));
  consumers &= ~(1u << me);
  if (!consumers) return;
  int nown = mye - myb;
  const int e0 = ptr ? ptr[i] : i, e1 = ptr ? ptr[i + 1] : i + 1;
  for (int e = e0; e < e1; e++) {
    int j = col[e];
    if (colmap) j = colmap[j];
    if (j >= myb && j < mye)
      for (unsigned m = consumers; m; m &= m - 1) flags[(size_t)(__ffs(m) - 1) * nown + (j - myb)] = 1;
  }
}
// which ranks consume a row of a sharded level: its owner and the owners of the operator rows that reference it (exactly
// the relation the operator-halo lists are built from; no symmetry assumption).  mask must be zeroed.
__global__ void consumer_mask_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ col, Ranges rows, unsigned* __restrict__ mask) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned me = 1u << owner_of(rows, i);
  atomicOr(&mask[i], me);
  for (int e = ptr[i]; e < ptr[i + 1]; e++) {
    const int j = col[e];
    if (!(mask[j] & me)) atomicOr(&mask[j], me);
  }
}
__global__ void fill_list_kernel(long long total, int nown, int myb, const int* __restrict__ flags, const int* __restrict__ pos,
                                 const int* __restrict__ valmap, int* __restrict__ list) {
  long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k < total && flags[k]) {
    const int j = myb + (int)(k % nown);
    list[pos[k] - 1] = valmap ? valmap[j] : j;
  }
}

// what MY consumer rows [r0, r1) need from the others: flags over the (mapped) column space
__global__ void mark_needs_kernel(int r0, int r1, const int* __restrict__ ptr, const int* __restrict__ col, const int* __restrict__ colmap,
                                  const unsigned* __restrict__ rowmask, int me, int myb, int mye, int* __restrict__ flags) {
  int i = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r1) return;
  if (rowmask && !(rowmask[i] >> me & 1u)) return;  // consumer rows = all rows whose mask names me (own + ghost rows)
  const int e0 = ptr ? ptr[i] : i, e1 = ptr ? ptr[i + 1] : i + 1;
  for (int e = e0; e < e1; e++) {
    int j = col[e];
    if (colmap) j = colmap[j];
    if (j < myb || j >= mye) flags[j] = 1;
  }
}
__global__ void fill_needs_kernel(int ncols, const int* __restrict__ flags, const int* __restrict__ pos, const int* __restrict__ valmap,
                                  int* __restrict__ list) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < ncols && flags[j]) list[pos[j] - 1] = valmap ? valmap[j] : j;
}

// flag[i - r0] = 1 when consumer row i of [r0, r1) references a value that is not mine (it arrives through an exchange)
__global__ void ghost_ref_rows_kernel(int r0, int r1, const int* __restrict__ ptr, const int* __restrict__ col, const int* __restrict__ colmap,
                                      int myb, int mye, unsigned char* __restrict__ flag) {
  int i = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r1) return;
  unsigned char f = 0;
  for (int e = ptr[i]; e < ptr[i + 1]; e++) {
    int j = col[e];
    if (colmap) j = colmap[j];
    if (j < myb || j >= mye) { f = 1; break; }
  }
  flag[i - r0] = f;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// longest run of 1024-row blocks (globally aligned, fully inside [r0, r1)) without a ghost reference
RowRange interior_range(const Ctx& c, int r0, int r1, const int* ptr, const int* col, const int* colmap, int myb, int mye) {
  RowRange out;
  out.begin = out.end = 0;
  if (r1 - r0 < 4096) return out;
  cudaStream_t s = c.stream;
  DevBuf<unsigned char> flag(r1 - r0, s);
  ghost_ref_rows_kernel<<<cdiv(r1 - r0, 256), 256, 0, s>>>(r0, r1, ptr, col, colmap, myb, mye, flag);
  FSB_CHECK_LAUNCH();
  std::vector<unsigned char> f = flag.to_vector();
  const int B = 1024;
  int best0 = 0, best1 = 0, run0 = -1;
  for (int k = (r0 + B - 1) / B; (k + 1) * B <= r1; k++) {
    bool clean = true;
    for (int i = k * B; i < (k + 1) * B; i++) if (f[i - r0]) { clean = false; break; }
    if (clean) {
      if (run0 < 0) run0 = k * B;
      if ((k + 1) * B - run0 > best1 - best0) { best0 = run0; best1 = (k + 1) * B; }
    } else run0 = -1;
  }
  if (best1 - best0 >= (r1 - r0) / 4) { out.begin = best0; out.end = best1; }
  return out;
}

// One exchange: `nrows` consumer rows (owner by rowOwner) reference values (col / colmap) of which
// [myb, mye) are mine; the list stores valmap[j] (or j) = index into the exchanged vector.
void build_push_list(const Ctx& c, int nranks, int me, int nrows, const int* ptr, const int* col, const int* colmap, const Ranges& rowOwner,
                     const Ranges& colOwner, int myb, int mye, const int* valmap, Solver::PushList& out,
                     const unsigned* rowmask = nullptr) {
  cudaStream_t s = c.stream;
  const int nown = std::max(mye - myb, 0);
  out.total = 0;
  std::vector<int> ptr_h(nranks + 1, 0);
  if (nown > 0 && nrows > 0) {
    const size_t m = (size_t)nranks * nown;
    IBuf flags(m, s), pos(m, s);
    flags.zero();
    mark_needed_kernel<<<cdiv(nrows, 256), 256, 0, s>>>(nrows, ptr, col, colmap, rowOwner, rowmask, me, myb, mye, flags);
    inclusive_scan_i32(flags, pos, m, s);
    for (int q = 0; q < nranks; q++) ptr_h[q + 1] = pos.read((size_t)(q + 1) * nown - 1);
    out.total = ptr_h[nranks];
    out.idx.alloc(std::max(1, out.total), s);
    fill_list_kernel<<<cdiv((long long)m, 256), 256, 0, s>>>((long long)m, nown, myb, flags, pos, valmap, out.idx);
    FSB_CUDA(cudaStreamSynchronize(s));
  } else {
    out.idx.alloc(1, s);
  }
  out.ptr.alloc(nranks + 1, s);
  out.ptr.from_host(ptr_h.data(), nranks + 1);
  // what the peers send me: the values my own consumer rows reference, ascending in the owner's numbering (the
  // order in which the owner's list sends them), grouped by owner
  const int r0 = rowmask ? 0 : rowOwner.b[me], r1 = rowmask ? nrows : rowOwner.b[me + 1];
  const int ncols = colOwner.b[nranks];
  out.rtotal = 0;
  for (int q = 0; q <= nranks; q++) out.rptr[q] = 0;
  if (r1 > r0 && ncols > 0) {
    IBuf flags(ncols, s), pos(ncols, s);
    flags.zero();
    mark_needs_kernel<<<cdiv(r1 - r0, 256), 256, 0, s>>>(r0, r1, ptr, col, colmap, rowmask, me, myb, mye, flags);
    inclusive_scan_i32(flags, pos, ncols, s);
    for (int q = 1; q <= nranks; q++) out.rptr[q] = colOwner.b[q] > 0 ? pos.read((size_t)colOwner.b[q] - 1) : 0;
    out.rtotal = out.rptr[nranks];
    out.ridx.alloc(std::max(1, out.rtotal), s);
    fill_needs_kernel<<<cdiv(ncols, 256), 256, 0, s>>>(ncols, flags, pos, valmap, out.ridx);
    FSB_CHECK_LAUNCH();
    FSB_CUDA(cudaStreamSynchronize(s));
  } else {
    out.ridx.alloc(1, s);
  }
}

// all-gather as an exchange: my slice [b[me], b[me+1]) of a vector goes to every peer, theirs come to me
void build_allgather_list(const Ctx& c, int nranks, int me, const int* b, Solver::PushList& out) {
  cudaStream_t s = c.stream;
  const int mine = b[me + 1] - b[me];
  std::vector<int> idx, ptr(nranks + 1, 0), ridx;
  for (int q = 0; q < nranks; q++) {
    if (q != me) for (int j = b[me]; j < b[me + 1]; j++) idx.push_back(j);
    ptr[q + 1] = (int)idx.size();
  }
  out.rptr[0] = 0;
  for (int q = 0; q < nranks; q++) {
    if (q != me) for (int j = b[q]; j < b[q + 1]; j++) ridx.push_back(j);
    out.rptr[q + 1] = (int)ridx.size();
  }
  (void)mine;
  out.total = (int)idx.size(); out.rtotal = (int)ridx.size();
  out.idx.alloc(std::max<size_t>(1, idx.size()), s); out.ridx.alloc(std::max<size_t>(1, ridx.size()), s); out.ptr.alloc(nranks + 1, s);
  if (!idx.empty()) out.idx.from_host(idx.data(), idx.size());
  if (!ridx.empty()) out.ridx.from_host(ridx.data(), ridx.size());
  out.ptr.from_host(ptr.data(), nranks + 1);
  FSB_CUDA(cudaStreamSynchronize(s));
}

Ranges make_ranges(const int* b, int nranks) {
  Ranges r;
  r.n = nranks;
  for (int q = 0; q <= nranks; q++) r.b[q] = b[q];
  for (int q = nranks + 1; q <= kMaxRanks; q++) r.b[q] = b[nranks];
  return r;
}

__global__ void minmax_gather_kernel(int r0, int r1, const int* __restrict__ idx, int* __restrict__ lohi) {
  int lo = INT_MAX, hi = INT_MIN;
  for (int i = r0 + blockIdx.x * blockDim.x + threadIdx.x; i < r1; i += gridDim.x * blockDim.x) {
    const int v = idx[i];
    lo = min(lo, v); hi = max(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if ((threadIdx.x & 31) == 0 && lo <= hi) { atomicMin(&lohi[0], lo); atomicMax(&lohi[1], hi); }
}

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
    cudaStream_t s = ctx.stream;
    for (int l = 0; l < dist.nshard && l < (int)levels.size(); l++) {
      LevelData& L = levels[l];
      L.x.release(); L.r.release(); L.bc.release(); L.xc.release();
      L.x.alloc(L.n, s); L.r.alloc(L.n, s); L.bc.alloc(std::max(1, L.nnout), s); L.xc.alloc(std::max(1, L.nnout), s);
      L.ownP0 = 0; L.ownPn = 0;
    }
    FSB_CUDA(cudaStreamSynchronize(s));
    cudaFree(dist.arena);
    dist.arena = nullptr;
  }
  dist.lev.clear();
  dist.nshard = 0;
  dist.connected = false;
  dist.nranks = 1; dist.rank = 0;
  ctx.dist = DistDev();
  ctx.dist_ticket = nullptr;
}

void Solver::dist_prepare(int rank, int nranks) {
  if (!has_setup) throw std::runtime_error("dist_prepare before setup");
  if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) throw std::invalid_argument("bad rank / nranks (at most 8 GPUs of one box)");
  FSB_CUDA(cudaSetDevice(ctx.device));
  dist_disconnect();
  dist.rank = rank; dist.nranks = nranks;
  if (nranks == 1) return;
  if (levels.size() < 2 || !levels[0].use_ell || !levels[0].sA.ready() || !levels[0].sAout.ready() || !levels[0].sP.ready())
    throw std::runtime_error("sharded solve needs a fine level with >= 32768 rows on the ELL/SELL path");
  cudaStream_t s = ctx.stream;
  // which levels are sharded: the fine level always; a coarser one while every GPU still gets enough rows to fill it
  // and the level is a regular smoothing level (above the dense tail / the coarsest level)
  static const int min_rows_per_rank = getenv("FSB_SHARD_MINROWS") ? atoi(getenv("FSB_SHARD_MINROWS")) : 16384;  // tuning knob
  static const int max_shard_levels = getenv("FSB_SHARD_LEVELS") ? atoi(getenv("FSB_SHARD_LEVELS")) : 4;        // tuning knob
  const int last_smoothing = (tail_level_ >= 0 ? tail_level_ : (int)levels.size() - 1) - 1;
  int nshard = 1;
  while (nshard <= last_smoothing && nshard < max_shard_levels && levels[nshard].nparts >= nranks &&
         (long long)levels[nshard].n >= (long long)min_rows_per_rank * nranks && kChanLevel0 + kChanPerLevel * (nshard + 1) <= kMaxChan)
    nshard++;
  dist.nshard = nshard;
  dist.lev.clear();
  dist.lev.resize(nshard);
  // 1. nnz-balanced contiguous ranges of partitions on every sharded level
  std::vector<std::vector<int>> ps(nshard);
  for (int l = 0; l < nshard; l++) {
    LevelData& L = levels[l];
    DistLevel& D = dist.lev[l];
    const int np = L.nparts;
    ps[l] = L.pstart.to_vector();
    std::vector<int> pidx = L.agg.partitionIdx.to_vector(), rowptr = L.A.ptr.to_vector();
    std::vector<long long> w(np);
    for (int p = 0; p < np; p++) w[p] = rowptr[ps[l][p + 1]] - rowptr[ps[l][p]];
    split_by_weight(np, w.data(), nranks, D.pbeg);
    for (int r = 0; r <= nranks; r++) { D.rbeg[r] = ps[l][D.pbeg[r]]; D.abeg[r] = pidx[D.pbeg[r]]; }
    L.ownP0 = D.pbeg[rank]; L.ownPn = D.pbeg[rank + 1] - D.pbeg[rank];
    if (L.use_ell) {  // owned partition lists of the register-resident smoother's size classes
      std::vector<EllDesc> lists[kEllClasses];
      for (int p = D.pbeg[rank]; p < D.pbeg[rank + 1]; p++) lists[ell_class(ps[l][p + 1] - ps[l][p])].push_back(L.ellDescHost[p]);
      for (int q = 0; q < kEllClasses; q++) {
        L.nlistOwn[q] = (int)lists[q].size();
        L.plistOwn[q].alloc(std::max<size_t>(1, lists[q].size()), s);
        if (!lists[q].empty()) L.plistOwn[q].from_host(lists[q].data(), lists[q].size());
      }
    }
  }
  // 2. push lists
  for (int l = 0; l < nshard; l++) {
    LevelData& L = levels[l];
    DistLevel& D = dist.lev[l];
    const Ranges rows = make_ranges(D.rbeg, nranks), aggs = make_ranges(D.abeg, nranks);
    const int myb = D.rbeg[rank], mye = D.rbeg[rank + 1];
    // my rows that a peer's operator rows reference
    build_push_list(ctx, nranks, rank, L.n, L.A.ptr, L.A.col, nullptr, rows, rows, myb, mye, nullptr, D.sendA);
    // my rows that a peer's restriction rows reference (rows of R = next level's external numbering)
    build_push_list(ctx, nranks, rank, L.nnout, L.R.ptr, L.R.col, nullptr, aggs, rows, myb, mye, nullptr, D.sendR);
    if (l + 1 >= nshard) {
      // the next level is replicated: its right-hand side is all-gathered
      build_allgather_list(ctx, nranks, rank, D.abeg, D.sendDown);
    } else {
      LevelData& Ln = levels[l + 1];
      DistLevel& Dn = dist.lev[l + 1];
      const Ranges rowsN = make_ranges(Dn.rbeg, nranks);
      // down: internal row i of the next level reads bc[ipermutation[i]]; I computed the entries of my aggregates
      build_push_list(ctx, nranks, rank, Ln.n, nullptr, Ln.agg.ipermutation, nullptr, rowsN, aggs, D.abeg[rank], D.abeg[rank + 1], nullptr, D.sendDown);
      // up: a prolongator row references xc[e], e external on the next level; mine when permutation[e] is one of my
      // next-level rows; the pushed index is e = ipermutation[internal].  The consumers of a row are its owner AND the
      // ranks that hold it as a ghost row: those apply the coarse correction to their ghost copy themselves
      // (x_ghost += P_ghost xc, bit-identical to the owner's update), which saves the halo exchange after the prolongation.
      DevBuf<unsigned> cmask(L.n, s);
      cmask.zero();
      consumer_mask_kernel<<<cdiv(L.n, 256), 256, 0, s>>>(L.n, L.A.ptr, L.A.col, rows, cmask);
      build_push_list(ctx, nranks, rank, L.n, L.P.ptr, L.P.col, Ln.agg.permutation, rows, rowsN, Dn.rbeg[rank], Dn.rbeg[rank + 1], Ln.agg.ipermutation,
                      D.sendUp, cmask);
    }
  }
  // 2b. interior ranges of the big sharded levels: where the consumer of an exchange can start before the halo is in.
  // OFF by default (0): measured on 8 B200 with the ~100 M-tet cube (2.1 M rows per GPU, interior = 57 % of the rows) the
  // split costs more than it hides — 33.5 ms per solve with it, 32.3 ms without (profiles/r2_bench_cube255_8gpu*.json): an
  // exchange is ~14 us, while cutting a 35-100 us streaming kernel into three launches adds two ramps and two tails.
  // FSB_OVERLAP_MINROWS=<rows per GPU> switches it on for levels at least that large.
  static const int overlap_min_rows = getenv("FSB_OVERLAP_MINROWS") ? atoi(getenv("FSB_OVERLAP_MINROWS")) : 0;
  for (int l = 0; l < nshard; l++) {
    LevelData& L = levels[l];
    DistLevel& D = dist.lev[l];
    D.intA = D.intR = D.intP = RowRange();
    D.intA.end = D.intR.end = D.intP.end = 0;
    const int myb = D.rbeg[rank], mye = D.rbeg[rank + 1];
    if (overlap_min_rows <= 0 || mye - myb < overlap_min_rows) continue;
    // test knob: only the odd ranks split their consumers (GPUs decide on their interior ranges independently — mixed
    // decisions must work: every exchange sits at the same point of the sequence on every rank, whatever stream it runs on)
    if (getenv("FSB_OVERLAP_ODD_RANKS") && atoi(getenv("FSB_OVERLAP_ODD_RANKS")) != 0 && (rank & 1) == 0) continue;
    D.intA = interior_range(ctx, myb, mye, L.Aout.ptr, L.Aout.col, nullptr, myb, mye);
    D.intR = interior_range(ctx, D.abeg[rank], D.abeg[rank + 1], L.R.ptr, L.R.col, nullptr, myb, mye);
    if (l + 1 < nshard) {
      DistLevel& Dn = dist.lev[l + 1];
      D.intP = interior_range(ctx, myb, mye, L.P.ptr, L.P.col, levels[l + 1].agg.permutation, Dn.rbeg[rank], Dn.rbeg[rank + 1]);
    }
  }
  // 3. user-numbering range that covers my fine rows (host-buffer solves move only this slice over PCIe)
  {
    LevelData& L0 = levels[0];
    IBuf lohi(2, s);
    const int init[2] = {INT_MAX, INT_MIN};
    lohi.from_host(init, 2);
    const int r0 = dist.lev[0].rbeg[rank], r1 = dist.lev[0].rbeg[rank + 1];
    if (r1 > r0 && !prm.refLevel0NoPerm) {
      minmax_gather_kernel<<<std::min(cdiv(r1 - r0, 256), 4 * ctx.num_sms), 256, 0, s>>>(r0, r1, L0.agg.ipermutation, lohi);
      FSB_CHECK_LAUNCH();
      std::vector<int> r = lohi.to_vector();
      dist.user_lo = r[0]; dist.user_hi = r[1] + 1;
    } else { dist.user_lo = r0; dist.user_hi = r1; }
  }
  // 4. channels: one receive buffer per exchange site
  {
    for (int c = 0; c < kMaxChan; c++) dist.chan[c] = Channel();
    dist.chan[kChanP].list = &dist.lev[0].sendA;
    dist.chan[kChanX0].list = &dist.lev[0].sendA;
    for (int l = 0; l < nshard; l++) {
      DistLevel& D = dist.lev[l];
      const int c0 = kChanLevel0 + kChanPerLevel * l;
      dist.chan[c0 + kXPre].list = &D.sendA;
      dist.chan[c0 + kRes].list = &D.sendR;
      dist.chan[c0 + kDown].list = &D.sendDown;
      if (l + 1 < nshard) dist.chan[c0 + kUp].list = &D.sendUp;
      dist.chan[c0 + kXPost].list = &D.sendA;
    }
    dist.nchan = kChanLevel0 + kChanPerLevel * nshard;
  }
  // 5. arena: flags | all-reduce slots | p | cg_x | per sharded level x, r, bc, xc | receive buffers.  The offsets of the
  // receive buffers differ from rank to rank (they depend on what a rank receives) and travel in the blob.
  size_t off = 0;
  dist.off_flags = off; off = align_up(off + (size_t)kMaxRanks * sizeof(unsigned long long), 256);
  dist.off_red = off; off = align_up(off + 2 * kMaxRanks * sizeof(uint4), 256);
  const int n = levels[0].n;
  dist.off_p = off; off = align_up(off + (size_t)n * 8, 256);
  dist.off_cgx = off; off = align_up(off + (size_t)n * 8, 256);
  for (int l = 0; l < nshard; l++) {
    DistLevel& D = dist.lev[l];
    const size_t nl = (size_t)levels[l].n, nc = (size_t)std::max(levels[l].nnout, 1);
    D.off_x = off; off = align_up(off + nl * 8, 256);
    D.off_r = off; off = align_up(off + nl * 8, 256);
    D.off_bc = off; off = align_up(off + nc * 8, 256);
    D.off_xc = off; off = align_up(off + nc * 8, 256);
  }
  for (int c = 0; c < dist.nchan; c++) {
    Channel& ch = dist.chan[c];
    if (!ch.list) continue;
    ch.buf_off = off; off = align_up(off + (size_t)std::max(ch.list->rtotal, 1) * sizeof(uint4), 256);
  }
  dist.arena_bytes = off;
  FSB_CUDA(cudaMalloc((void**)&dist.arena, dist.arena_bytes));
  FSB_CUDA(cudaMemsetAsync(dist.arena, 0, dist.arena_bytes, s));
  cg_p.view(reinterpret_cast<double*>(dist.arena + dist.off_p), n, s);
  cg_x.view(reinterpret_cast<double*>(dist.arena + dist.off_cgx), n, s);
  for (int l = 0; l < nshard; l++) {
    LevelData& L = levels[l];
    DistLevel& D = dist.lev[l];
    L.x.view(reinterpret_cast<double*>(dist.arena + D.off_x), L.n, s);
    L.r.view(reinterpret_cast<double*>(dist.arena + D.off_r), L.n, s);
    L.bc.view(reinterpret_cast<double*>(dist.arena + D.off_bc), L.nnout, s);
    L.xc.view(reinterpret_cast<double*>(dist.arena + D.off_xc), L.nnout, s);
  }
  dist.epoch.alloc(2, s); dist.epoch.zero();
  dist.xchg.alloc(1, s); dist.xchg.zero();
  dist.error.alloc(1, s); dist.error.zero();
  dist.ticket.alloc(1, s); dist.ticket.zero();
  FSB_CUDA(cudaStreamSynchronize(s));
}

void Solver::dist_get_blob(void* blob, long long* bytes) {
  if (!dist.arena) throw std::runtime_error("dist_get_blob before dist_prepare");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  std::vector<unsigned char> raw(kDistBlobBytes, 0);
  DistBlob* b = reinterpret_cast<DistBlob*>(raw.data());
  cudaIpcMemHandle_t h;
  FSB_CUDA(cudaIpcGetMemHandle(&h, dist.arena));
  memcpy(b->handle, &h, 64);
  b->nchan = dist.nchan; b->nranks = dist.nranks;
  for (int c = 0; c < dist.nchan; c++) {
    const Channel& ch = dist.chan[c];
    b->buf_off[c] = ch.buf_off;
    for (int q = 0; q <= kMaxRanks; q++) b->rptr[c][q] = ch.list ? ch.list->rptr[std::min(q, dist.nranks)] : 0;
  }
  memcpy(blob, raw.data(), kDistBlobBytes);
  if (bytes) *bytes = (long long)dist.arena_bytes;
}

void Solver::dist_connect(const void* blobs) {
  if (!dist.arena) throw std::runtime_error("dist_connect before dist_prepare");
  FSB_CUDA(cudaSetDevice(ctx.device));
  const char* hb = static_cast<const char*>(blobs);
  std::vector<DistBlob> B(dist.nranks);
  for (int q = 0; q < dist.nranks; q++) memcpy(&B[q], hb + (size_t)kDistBlobBytes * q, sizeof(DistBlob));
  for (int q = 0; q < dist.nranks; q++)
    if (B[q].nchan != dist.nchan || B[q].nranks != dist.nranks) throw std::runtime_error("sharded solve: the ranks disagree on the exchange plan (setup must be replicated)");
  for (int q = 0; q < dist.nranks; q++) {
    if (q == dist.rank) { dist.peer[q] = dist.arena; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, B[q].handle, 64);
    void* p = nullptr;
    FSB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    dist.peer[q] = static_cast<char*>(p);
  }
  // exchange descriptors: my segment inside peer q's receive buffer of channel c starts at slot rptr_q[c][me]; what q
  // expects from me must be exactly what my list sends (both sides derived it from the replicated hierarchy)
  for (int c = 0; c < dist.nchan; c++) {
    Channel& ch = dist.chan[c];
    if (!ch.list) continue;
    const PushList& pl = *ch.list;
    std::vector<int> sptr = pl.ptr.to_vector();
    LLXchg x;
    x.sidx = pl.idx.get(); x.sptr = pl.ptr.get(); x.stotal = pl.total;
    x.ridx = pl.ridx.get(); x.rtotal = pl.rtotal;
    x.mybuf = reinterpret_cast<uint4*>(dist.arena + ch.buf_off);
    for (int q = 0; q < dist.nranks; q++) {
      const int expect = B[q].rptr[c][dist.rank + 1] - B[q].rptr[c][dist.rank], mine = sptr[q + 1] - sptr[q];
      if (q != dist.rank && expect != mine)
        throw std::runtime_error("sharded solve: channel " + std::to_string(c) + ": rank " + std::to_string(q) + " expects " + std::to_string(expect) +
                                 " values from rank " + std::to_string(dist.rank) + ", which sends " + std::to_string(mine));
      x.peerbuf[q] = reinterpret_cast<uint4*>(dist.peer[q] + B[q].buf_off[c]) + B[q].rptr[c][dist.rank];
    }
    ch.dev = x;
  }
  DistDev d;
  d.rank = dist.rank; d.nranks = dist.nranks;
  d.my_flags = reinterpret_cast<unsigned long long*>(dist.arena + dist.off_flags);
  d.my_red = reinterpret_cast<uint4*>(dist.arena + dist.off_red);
  for (int q = 0; q < dist.nranks; q++) {
    d.peer_flags[q] = reinterpret_cast<unsigned long long*>(dist.peer[q] + dist.off_flags);
    d.peer_red[q] = reinterpret_cast<uint4*>(dist.peer[q] + dist.off_red);
  }
  d.epoch = dist.epoch.get();
  d.xchg = dist.xchg.get();
  d.error = dist.error.get();
  ctx.dist = d;
  ctx.dist_ticket = dist.ticket.get();
  dist.connected = true;
  destroy_graph();
}

}  // namespace fsb
