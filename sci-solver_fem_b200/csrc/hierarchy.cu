// hierarchy.cu — stage 2b: level construction on device: symmetric permutation of A,
// smoothed prolongator, transpose, Galerkin product, coarsest-level inverse.
//
// Reference: SmoothedMG_AMG_Level::generateMatrixSymmetric_d (src/core/cuda/
// smoothedMG_amg_level.cu:199-304, matrixpermute_kernel :45-80), generateProlongatorFull_d
// (:320-387), generateNextLevelMatrixFull_d (:389-400, two cusp::multiply ESC SpGEMMs),
// coarse LU (amg.cu:101-108).
//
// B200 design: no global expand-sort-compress.  Every product is formed row-wise (Gustavson)
// by the thread that owns the output row, which keeps a small column-sorted accumulator list
// in a private scratch segment; summation order is therefore fixed (k ascending, then the
// B-row order), the output columns come out sorted, and the whole hierarchy is bit-reproducible.
// Compiled with -fmad=false so that products and sums round separately, as on the host.
#include <cub/cub.cuh>

#include "fsb_internal.h"
#include <ctime>
#include <cstdio>
#include "solver.h"

namespace fsb {
namespace {

__global__ void row_lengths_permuted(int n, const int* __restrict__ ptr, const int* __restrict__ perm, int* __restrict__ len) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) len[perm[i]] = ptr[i + 1] - ptr[i];
}

// matrixpermute_kernel: (row, col) -> (perm[row], perm[col]); rows re-sorted by column (insertion sort in the destination
// row, which stays in L1; a shared-memory list per thread was measured slower here: 3.2 vs 2.3 ms at N=118, occupancy-bound).
__global__ void permute_rows(int n, const int* __restrict__ ptrA, const int* __restrict__ colA, const double* __restrict__ valA,
                             const int* __restrict__ perm, const int* __restrict__ ptrB, int* __restrict__ colB, double* __restrict__ valB) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int base = ptrB[perm[i]];
  int cnt = 0;
  for (int e = ptrA[i]; e < ptrA[i + 1]; e++) {
    int cnew = perm[colA[e]];
    double v = valA[e];
    int q = base + cnt;
    while (q > base && colB[q - 1] > cnew) { colB[q] = colB[q - 1]; valB[q] = valB[q - 1]; q--; }
    colB[q] = cnew; valB[q] = v;
    cnt++;
  }
}

__global__ void diag_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ col, const double* __restrict__ val, double* __restrict__ diag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double d = 0.0;
  for (int e = ptr[i]; e < ptr[i + 1]; e++) if (col[e] == i) d = val[e];
  diag[i] = d;
}

__global__ void agg_of_row(int nAgg, const int* __restrict__ aggregateIdx, int* __restrict__ aggOf) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nAgg) return;
  for (int r = aggregateIdx[a]; r < aggregateIdx[a + 1]; r++) aggOf[r] = a;
}

// sorted-list accumulate: returns the new count
__device__ __forceinline__ int acc_insert(int* cols, double* vals, int cnt, int j, double v) {
  int lo = 0, hi = cnt;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (cols[mid] < j) lo = mid + 1; else hi = mid; }
  if (lo < cnt && cols[lo] == j) { vals[lo] += v; return cnt; }
  for (int q = cnt; q > lo; q--) { cols[q] = cols[q - 1]; vals[q] = vals[q - 1]; }
  cols[lo] = j; vals[lo] = 0.0 + v;
  return cnt + 1;
}

// The same sorted-list accumulate on a per-thread list in SHARED memory (entry i of thread t at [i * ROW_THREADS + t]):
// the global scratch rows of the thread-per-row kernels are ~100 B apart per thread and were read-modify-written once per
// product.  Returns the new count, or -1 when a new column does not fit (the caller redoes the row on the global scratch).
constexpr int ROW_THREADS = 64;   // threads per CTA of the thread-per-row setup kernels
constexpr int ROW_CAP = 32;       // list entries per thread in shared memory (24.5 KB per CTA)
__device__ __forceinline__ int acc_insert_sh(int* cols, double* vals, int cnt, int j, double v) {
  int lo = 0, hi = cnt;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (cols[mid * ROW_THREADS] < j) lo = mid + 1; else hi = mid; }
  if (lo < cnt && cols[lo * ROW_THREADS] == j) { vals[lo * ROW_THREADS] += v; return cnt; }
  if (cnt == ROW_CAP) return -1;
  for (int q = cnt; q > lo; q--) { cols[q * ROW_THREADS] = cols[(q - 1) * ROW_THREADS]; vals[q * ROW_THREADS] = vals[(q - 1) * ROW_THREADS]; }
  cols[lo * ROW_THREADS] = j; vals[lo * ROW_THREADS] = 0.0 + v;
  return cnt + 1;
}

// P = T - omega D^-1 A T, row-wise.  Terms of one (row, aggregate) are added in column order
// and the tentative 1 last — the order the reference's stable sort + reduce_by_key produces.
__global__ void prolongator_rows(int n, const int* __restrict__ ptr, const int* __restrict__ col, const double* __restrict__ val,
                                 const double* __restrict__ diag, const int* __restrict__ aggOf, double omega,
                                 int* __restrict__ scol, double* __restrict__ sval, int* __restrict__ count) {
  __shared__ int sh_c[ROW_CAP * ROW_THREADS];
  __shared__ double sh_v[ROW_CAP * ROW_THREADS];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long base = (long long)ptr[i] + i;  // room for row length + 1 entries
  int* cols = scol + base;
  double* vals = sval + base;
  double d = diag[i];
  int* lc = sh_c + threadIdx.x;
  double* lv = sh_v + threadIdx.x;
  int cnt = 0;
  for (int e = ptr[i]; e < ptr[i + 1] && cnt >= 0; e++) {
    double term = (-omega * val[e] * 1.0) / d;
    cnt = acc_insert_sh(lc, lv, cnt, aggOf[col[e]], term);
  }
  if (cnt >= 0) cnt = acc_insert_sh(lc, lv, cnt, aggOf[i], 1.0);
  if (cnt >= 0) {
    for (int q = 0; q < cnt; q++) { cols[q] = lc[q * ROW_THREADS]; vals[q] = lv[q * ROW_THREADS]; }
  } else {  // more distinct aggregates than the shared list holds: the same accumulation on the global scratch row
    cnt = 0;
    for (int e = ptr[i]; e < ptr[i + 1]; e++) {
      double term = (-omega * val[e] * 1.0) / d;
      cnt = acc_insert(cols, vals, cnt, aggOf[col[e]], term);
    }
    cnt = acc_insert(cols, vals, cnt, aggOf[i], 1.0);
  }
  count[i] = cnt;
}

__global__ void compact_rows(int n, const long long* __restrict__ sbase, const int* __restrict__ scol, const double* __restrict__ sval,
                             const int* __restrict__ ptrC, int* __restrict__ colC, double* __restrict__ valC) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long b = sbase[i];
  int o = ptrC[i], cnt = ptrC[i + 1] - o;
  for (int q = 0; q < cnt; q++) { colC[o + q] = scol[b + q]; valC[o + q] = sval[b + q]; }
}

__global__ void prolongator_bases(int n, const int* __restrict__ ptr, long long* __restrict__ sbase) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sbase[i] = (long long)ptr[i] + i;
}

__global__ void spgemm_bound(int n, const int* __restrict__ ptrA, const int* __restrict__ colA, const int* __restrict__ ptrB, long long* __restrict__ ub) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long s = 0;
  for (int e = ptrA[i]; e < ptrA[i + 1]; e++) { int k = colA[e]; s += ptrB[k + 1] - ptrB[k]; }
  ub[i] = s;
}

__global__ void spgemm_rows(int n, const int* __restrict__ ptrA, const int* __restrict__ colA, const double* __restrict__ valA,
                            const int* __restrict__ ptrB, const int* __restrict__ colB, const double* __restrict__ valB,
                            const long long* __restrict__ sbase, int* __restrict__ scol, double* __restrict__ sval, int* __restrict__ count,
                            long long min_products) {
  __shared__ int sh_c[ROW_CAP * ROW_THREADS];
  __shared__ double sh_v[ROW_CAP * ROW_THREADS];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // second pass after spgemm_warp_kernel: only the rows it flagged (count == -1) are redone here
  if (min_products > 0 && count[i] != -1) return;
  int* cols = scol + sbase[i];
  double* vals = sval + sbase[i];
  int* lc = sh_c + threadIdx.x;
  double* lv = sh_v + threadIdx.x;
  int cnt = 0;
  for (int e = ptrA[i]; e < ptrA[i + 1] && cnt >= 0; e++) {
    int k = colA[e];
    double a = valA[e];
    for (int f = ptrB[k]; f < ptrB[k + 1] && cnt >= 0; f++) cnt = acc_insert_sh(lc, lv, cnt, colB[f], a * valB[f]);
  }
  if (cnt >= 0) {
    for (int q = 0; q < cnt; q++) { cols[q] = lc[q * ROW_THREADS]; vals[q] = lv[q * ROW_THREADS]; }
  } else {  // more distinct columns than the shared list holds: the same accumulation on the global scratch row
    cnt = 0;
    for (int e = ptrA[i]; e < ptrA[i + 1]; e++) {
      int k = colA[e];
      double a = valA[e];
      for (int f = ptrB[k]; f < ptrB[k + 1]; f++) cnt = acc_insert(cols, vals, cnt, colB[f], a * valB[f]);
    }
  }
  count[i] = cnt;
}


// Warp-per-row Gustavson for rows with many products (R * (A P): ~850 products, ~40 distinct columns
// per coarse row).  Pass 1 collects the distinct columns in a shared-memory hash set and sorts them;
// pass 2 streams the products again in sequence order, groups them by output slot with a STABLE counting
// sort inside the warp (rank of a product among the earlier products of its slot: __match_any_sync over
// batches of 32 consecutive products + a running count per slot), and lane l then adds up the products of
// the slots it owns (slot mod 32 == l) as one contiguous run each — every slot is summed by one lane in
// product order, the same order as a sequential accumulation, so the result stays bit-identical to the
// host-order sum.  (The first version let every lane walk over ALL products of the row and pick its own:
// 91 % of the kernel's instructions, 10 of the 46 ms of the setup at N=118; profiles/r2_spgemm_before.txt.)
// Two instantiations: the small one (4 warps per CTA, two CTAs per SM) takes the rows of R (A P) on the big levels; rows
// it cannot hold (more than 512 entries in the A-row or more than 192 distinct output columns — the coarse rows of the
// last levels, whose operator rows span a good part of a few-hundred-column level) go to the large one (one warp per CTA),
// and only what exceeds that too falls back to the thread-per-row kernel.
template <int CHUNK_, int MAXROW_, int HASHBITS_, int MAXD_, int WARPS_>
struct WgCfg {
  static constexpr int CHUNK = CHUNK_;      // products staged per step
  static constexpr int MAXROW = MAXROW_;    // entries of the A-row (prefix table + staged A-row)
  static constexpr int HASHBITS = HASHBITS_;
  static constexpr int HASH = 1 << HASHBITS_;  // hash-set size; rows with more than MAXD distinct columns fall back
  static constexpr int MAXD = MAXD_;
  static constexpr int WARPS = WARPS_;
  // pval | pcol | pslot (u16) | hash | dist | pref | pb | va | sorted | prank (u16)
  static constexpr int SMEM_PER_WARP = CHUNK * 8 + CHUNK * 4 + CHUNK * 2 + HASH * 4 + HASH * 4 + (MAXROW + 1) * 4 + MAXROW * 4 + 8 + MAXROW * 8 +
                                       CHUNK * 8 + CHUNK * 2 + 16;
};
typedef WgCfg<512, 512, 8, 192, 4> WgSmall;
typedef WgCfg<512, 2048, 11, 1024, 1> WgLarge;

// products q0 .. q0+WG_CHUNK of the row: the A-row (B-row starts pb, values va) is staged in shared
// memory, so a product costs one global hop (colB / valB); four products per lane are in flight.
template <int WG_CHUNK>
__device__ __forceinline__ void wg_expand(int q0, int m, int na, int lane, const int* pref, const int* pb, const double* va,
                                          const int* __restrict__ colB, const double* __restrict__ valB, int* pcol, double* pval) {
  const int qe = min(m, q0 + WG_CHUNK);
  for (int qb = q0 + lane; qb < qe; qb += 128) {
    int f[4], e[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int q = qb + 32 * u;
      int lo = 0, hi = na;  // last A-entry with pref[lo] <= q
      if (q < qe) { while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pref[mid] <= q) lo = mid; else hi = mid; } }
      e[u] = lo; f[u] = q < qe ? pb[lo] + (q - pref[lo]) : -1;
    }
    int c[4]; double v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { c[u] = f[u] >= 0 ? __ldg(colB + f[u]) : 0; v[u] = (pval && f[u] >= 0) ? __ldg(valB + f[u]) : 0.0; }
#pragma unroll
    for (int u = 0; u < 4; u++) if (f[u] >= 0) { pcol[qb + 32 * u - q0] = c[u]; if (pval) pval[qb + 32 * u - q0] = va[e[u]] * v[u]; }
  }
}

template <typename Cfg, bool RETRY>
__global__ void __launch_bounds__(32 * Cfg::WARPS) spgemm_warp_kernel(int n, const int* __restrict__ ptrA, const int* __restrict__ colA,
                                                                      const double* __restrict__ valA, const int* __restrict__ ptrB,
                                                                      const int* __restrict__ colB, const double* __restrict__ valB,
                                                                      const long long* __restrict__ sbase, int* __restrict__ scol,
                                                                      double* __restrict__ sval, int* __restrict__ count) {
  constexpr int WG_CHUNK = Cfg::CHUNK, WG_MAXROW = Cfg::MAXROW, WG_HASH = Cfg::HASH, WG_MAXD = Cfg::MAXD, WG_WARPS = Cfg::WARPS;
  extern __shared__ __align__(16) unsigned char wg_smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* base = wg_smem + (size_t)w * ((Cfg::SMEM_PER_WARP + 15) & ~15);
  double* pval = reinterpret_cast<double*>(base);
  int* pcol = reinterpret_cast<int*>(base + WG_CHUNK * 8);
  unsigned short* pslot = reinterpret_cast<unsigned short*>(base + WG_CHUNK * 12);
  int* hash = reinterpret_cast<int*>(base + WG_CHUNK * 14);
  int* dist = hash + WG_HASH;
  int* pref = dist + WG_HASH;
  int* pb = pref + (WG_MAXROW + 1);
  double* va = reinterpret_cast<double*>(base + ((WG_CHUNK * 14 + WG_HASH * 8 + (WG_MAXROW + 1) * 4 + WG_MAXROW * 4 + 7) & ~7));
  double* sorted = va + WG_MAXROW;                                        // products of the chunk grouped by slot
  unsigned short* prank = reinterpret_cast<unsigned short*>(sorted + WG_CHUNK);
  int* scnt = hash;                                                       // the hash set is free after pass 1: per-slot counts / offsets
  const int i = blockIdx.x * WG_WARPS + w;
  if (i >= n) return;
  if (RETRY && count[i] != -1) return;  // second instantiation: only the rows the first one flagged
  const int a0 = ptrA[i], na = ptrA[i + 1] - a0;
  const int m = (int)(sbase[i + 1] - sbase[i]);
  if (na > WG_MAXROW) { if (lane == 0) count[i] = -1; return; }  // fallback row
  // prefix of B-row lengths over the entries of the A-row
  int run = 0;
  for (int e0 = 0; e0 < na; e0 += 32) {
    const int e = e0 + lane;
    int len = 0;
    if (e < na) { const int k = colA[a0 + e]; const int b0 = ptrB[k]; len = ptrB[k + 1] - b0; pb[e] = b0; va[e] = valA[a0 + e]; }
    int inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (e < na) pref[e] = run + inc - len;
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) pref[na] = run;
  for (int h = lane; h < WG_HASH; h += 32) hash[h] = -1;
  __syncwarp();
  // pass 1: distinct columns
  int overflow = 0;
  for (int q0 = 0; q0 < m; q0 += WG_CHUNK) {
    wg_expand<WG_CHUNK>(q0, m, na, lane, pref, pb, va, colB, valB, pcol, nullptr);
    __syncwarp();
    for (int q = q0 + lane; q < min(m, q0 + WG_CHUNK); q += 32) {
      const int c = pcol[q - q0];
      unsigned h = ((unsigned)c * 2654435761u) >> (32 - Cfg::HASHBITS);
      for (int probe = 0; probe < WG_HASH; probe++) {
        const int old = atomicCAS(&hash[h], -1, c);
        if (old == -1 || old == c) break;
        h = (h + 1) & (WG_HASH - 1);
        if (probe == WG_HASH - 1) overflow = 1;
      }
    }
    __syncwarp();
  }
  // compact + sort the distinct columns
  int nd = 0;
  for (int h0 = 0; h0 < WG_HASH; h0 += 32) {
    const int c = hash[h0 + lane];
    const unsigned ball = __ballot_sync(0xffffffffu, c != -1);
    if (c != -1) dist[nd + __popc(ball & ((1u << lane) - 1u))] = c;
    nd += __popc(ball);
  }
  overflow = __any_sync(0xffffffffu, overflow) || nd > WG_MAXD;
  if (overflow) { if (lane == 0) count[i] = -1; return; }
  for (int h = nd + lane; h < WG_HASH; h += 32) dist[h] = 0x7fffffff;
  __syncwarp();
  for (int k2 = 2; k2 <= WG_HASH; k2 <<= 1)
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int q = lane; q < WG_HASH; q += 32) {
        const int partner = q ^ j;
        if (partner > q) {
          const int x = dist[q], y = dist[partner];
          if ((x > y) == ((q & k2) == 0)) { dist[q] = y; dist[partner] = x; }
        }
      }
      __syncwarp();
    }
  // pass 2: numeric, every output slot summed by its owner lane in product order
  double acc[WG_MAXD / 32];
#pragma unroll
  for (int j = 0; j < WG_MAXD / 32; j++) acc[j] = 0.0;
  for (int q0 = 0; q0 < m; q0 += WG_CHUNK) {
    const int mc = min(m, q0 + WG_CHUNK) - q0;
    wg_expand<WG_CHUNK>(q0, m, na, lane, pref, pb, va, colB, valB, pcol, pval);
    __syncwarp();
    for (int h = lane; h < WG_HASH; h += 32) scnt[h] = 0;
    __syncwarp();
    // slot of every product (binary search in the sorted distinct list) and its rank among the earlier products of
    // that slot: batches of 32 consecutive products, lanes with equal slots found by __match_any_sync
    for (int q0b = 0; q0b < mc; q0b += 32) {
      const int q = q0b + lane;
      int sl = -1;
      if (q < mc) {
        const int c = pcol[q];
        int lo = 0, hi = nd - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (dist[mid] < c) lo = mid + 1; else hi = mid; }
        sl = lo;
        pslot[q] = (unsigned short)lo;
      }
      const unsigned peers = __match_any_sync(0xffffffffu, sl);
      if (q < mc) prank[q] = (unsigned short)(scnt[sl] + __popc(peers & ((1u << lane) - 1u)));
      __syncwarp();
      if (q < mc && (peers & ((1u << lane) - 1u)) == 0) scnt[sl] += __popc(peers);  // one lane per distinct slot
      __syncwarp();
    }
    // counts -> offsets (exclusive scan over the nd <= 192 slots; slot = j * 32 + lane)
    {
      int run = 0;
      for (int j0 = 0; j0 < nd; j0 += 32) {
        const int sl = j0 + lane;
        const int cnt = sl < nd ? scnt[sl] : 0;
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        __syncwarp();
        if (sl < nd) scnt[sl] = run + inc - cnt;
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (lane == 0) scnt[nd] = run;  // nd <= WG_MAXD < WG_HASH
    }
    __syncwarp();
    for (int q = lane; q < mc; q += 32) sorted[scnt[pslot[q]] + prank[q]] = pval[q];  // stable: sequence order inside a slot
    __syncwarp();
#pragma unroll
    for (int jj = 0; jj < WG_MAXD / 32; jj++) {
      const int sl = jj * 32 + lane;
      if (sl < nd) {
        const int e1 = scnt[sl + 1];
        double a = acc[jj];
        for (int k = scnt[sl]; k < e1; k++) a += sorted[k];
        acc[jj] = a;
      }
    }
    __syncwarp();
  }
  int* oc = scol + sbase[i];
  double* ov = sval + sbase[i];
#pragma unroll
  for (int jj = 0; jj < WG_MAXD / 32; jj++) {
    const int sl = jj * 32 + lane;
    if (sl < nd) { oc[sl] = dist[sl]; ov[sl] = 0.0 + acc[jj]; }
  }
  if (lane == 0) count[i] = nd;
}

__global__ void expand_rows(int n, const int* __restrict__ ptr, int* __restrict__ rows) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int e = ptr[i]; e < ptr[i + 1]; e++) rows[e] = i;
}
__global__ void count_cols(int nnz, const int* __restrict__ col, int* __restrict__ cnt) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nnz) atomicAdd(&cnt[col[e] + 1], 1);
}
__global__ void gather_transposed(int nnz, const int* __restrict__ order, const int* __restrict__ rows, const double* __restrict__ val,
                                  int* __restrict__ colT, double* __restrict__ valT) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nnz) { int src = order[e]; colT[e] = rows[src]; valT[e] = val[src]; }
}

void exclusive_scan_i64(const long long* in, long long* out, size_t n, cudaStream_t s) {
  void* tmp = nullptr; size_t bytes = 0;
  FSB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int64_t)n, s));
  FSB_CUDA(cudaMallocAsync(&tmp, bytes, s));
  FSB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, (int64_t)n, s));
  cudaFreeAsync(tmp, s);
}

// counts (n ints) -> ptr (n+1): exclusive scan with the total appended; returns the total
int counts_to_ptr(const Ctx& c, IBuf& count_plus1, int n, IBuf& ptr) {
  // count_plus1 has n+1 entries with the last one zeroed by the caller
  ptr.alloc(n + 1, c.stream);
  exclusive_scan_i32(count_plus1, ptr, n + 1, c.stream);
  return ptr.read(n);
}

// ---- dense inverse of the coarsest operator: Gauss-Jordan with partial pivoting, one CTA ----
__global__ void densify(int n, const int* __restrict__ ptr, const int* __restrict__ col, const double* __restrict__ val, double* __restrict__ M) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;  // M is n x 2n: [A | I]
  if (i >= n) return;
  for (int e = ptr[i]; e < ptr[i + 1]; e++) M[(size_t)i * 2 * n + col[e]] += val[e];
  M[(size_t)i * 2 * n + n + i] = 1.0;
}

__global__ void __launch_bounds__(1024) gauss_jordan(int n, double* __restrict__ M, double* __restrict__ Ainv) {
  __shared__ double s_val[1024];
  __shared__ int s_idx[1024];
  __shared__ int s_piv;
  const int t = threadIdx.x, T = blockDim.x, W = 2 * n;
  for (int k = 0; k < n; k++) {
    double best = -1.0; int bi = k;
    for (int i = k + t; i < n; i += T) { double a = fabs(M[(size_t)i * W + k]); if (a > best) { best = a; bi = i; } }
    s_val[t] = best; s_idx[t] = bi;
    __syncthreads();
    for (int off = T >> 1; off > 0; off >>= 1) {
      if (t < off) {
        // first maximal row wins (what a sequential partial-pivot search picks)
        if (s_val[t + off] > s_val[t] || (s_val[t + off] == s_val[t] && s_idx[t + off] < s_idx[t])) { s_val[t] = s_val[t + off]; s_idx[t] = s_idx[t + off]; }
      }
      __syncthreads();
    }
    if (t == 0) s_piv = s_idx[0];
    __syncthreads();
    int p = s_piv;
    if (p != k) for (int j = t; j < W; j += T) { double a = M[(size_t)k * W + j]; M[(size_t)k * W + j] = M[(size_t)p * W + j]; M[(size_t)p * W + j] = a; }
    __syncthreads();
    double d = M[(size_t)k * W + k];
    __syncthreads();
    for (int j = t; j < W; j += T) M[(size_t)k * W + j] /= d;
    __syncthreads();
    // eliminate column k from every other row: thread -> (row, column-chunk)
    for (long long q = t; q < (long long)n * W; q += T) {
      int i = (int)(q / W), j = (int)(q - (long long)i * W);
      if (i == k || j == k) continue;
      double f = M[(size_t)i * W + k];
      if (f != 0.0) M[(size_t)i * W + j] -= f * M[(size_t)k * W + j];
    }
    __syncthreads();
    for (int i = t; i < n; i += T) if (i != k) M[(size_t)i * W + k] = 0.0;
    __syncthreads();
  }
  for (long long q = t; q < (long long)n * n; q += T) {
    int i = (int)(q / n), j = (int)(q - (long long)i * n);
    Ainv[q] = M[(size_t)i * W + n + j];
  }
}

}  // namespace

void permute_csr(const Ctx& c, const DCsr& A, const int* perm, DCsr& B) {
  cudaStream_t s = c.stream;
  int n = A.nrows;
  IBuf len(n + 1, s);
  len.zero();
  row_lengths_permuted<<<cdiv(n, 256), 256, 0, s>>>(n, A.ptr, perm, len);
  B.nrows = A.nrows; B.ncols = A.ncols; B.nnz = A.nnz;
  B.ptr.alloc(n + 1, s);
  exclusive_scan_i32(len, B.ptr, n + 1, s);
  B.col.alloc(A.nnz, s); B.val.alloc(A.nnz, s);
  permute_rows<<<cdiv(n, 128), 128, 0, s>>>(n, A.ptr, A.col, A.val, perm, B.ptr, B.col, B.val);
  FSB_CHECK_LAUNCH();
}

void extract_diag(const Ctx& c, const DCsr& A, double* diag) {
  diag_kernel<<<cdiv(A.nrows, 256), 256, 0, c.stream>>>(A.nrows, A.ptr, A.col, A.val, diag);
  FSB_CHECK_LAUNCH();
}

void build_prolongator(const Ctx& c, const DCsr& A, const double* diag, const int* aggregateIdx, int nAgg, double omega, DCsr& P) {
  cudaStream_t s = c.stream;
  int n = A.nrows;
  IBuf aggOf(n, s), count(n + 1, s);
  count.zero();
  agg_of_row<<<cdiv(nAgg, 256), 256, 0, s>>>(nAgg, aggregateIdx, aggOf);
  size_t cap = (size_t)A.nnz + n;
  IBuf scol(cap, s); DBuf sval(cap, s);
  DevBuf<long long> sbase(n, s);
  prolongator_bases<<<cdiv(n, 256), 256, 0, s>>>(n, A.ptr, sbase);
  prolongator_rows<<<cdiv(n, ROW_THREADS), ROW_THREADS, 0, s>>>(n, A.ptr, A.col, A.val, diag, aggOf, omega, scol, sval, count);
  FSB_CHECK_LAUNCH();
  P.nrows = n; P.ncols = nAgg;
  P.nnz = counts_to_ptr(c, count, n, P.ptr);
  P.col.alloc(P.nnz, s); P.val.alloc(P.nnz, s);
  compact_rows<<<cdiv(n, 128), 128, 0, s>>>(n, sbase, scol, sval, P.ptr, P.col, P.val);
  FSB_CHECK_LAUNCH();
}

// At = A^T via a stable radix sort of the entries by column (cusp::transpose,
// smoothedMG_amg_level.cu:382): within a row of At the columns ascend.
void transpose_csr(const Ctx& c, const DCsr& A, DCsr& At) {
  cudaStream_t s = c.stream;
  IBuf rows(A.nnz, s), iotaE(A.nnz, s), order(A.nnz, s), sortedCols(A.nnz, s), cnt(A.ncols + 1, s);
  expand_rows<<<cdiv(A.nrows, 128), 128, 0, s>>>(A.nrows, A.ptr, rows);
  iota_i32(iotaE, A.nnz, s);
  sort_pairs_i32_i32(A.col, sortedCols, iotaE, order, A.nnz, bits_for(A.ncols), s);
  cnt.zero();
  count_cols<<<cdiv(A.nnz, 256), 256, 0, s>>>(A.nnz, A.col, cnt);
  At.nrows = A.ncols; At.ncols = A.nrows; At.nnz = A.nnz;
  At.ptr.alloc(A.ncols + 1, s);
  inclusive_scan_i32(cnt, At.ptr, A.ncols + 1, s);
  At.col.alloc(A.nnz, s); At.val.alloc(A.nnz, s);
  gather_transposed<<<cdiv(A.nnz, 256), 256, 0, s>>>(A.nnz, order, rows, A.val, At.col, At.val);
  FSB_CHECK_LAUNCH();
}

void spgemm(const Ctx& c, const DCsr& A, const DCsr& B, DCsr& C) {
  cudaStream_t s = c.stream;
  int n = A.nrows;
  DevBuf<long long> ub((size_t)n + 1, s), sbase((size_t)n + 1, s);
  ub.zero();
  spgemm_bound<<<cdiv(n, 256), 256, 0, s>>>(n, A.ptr, A.col, B.ptr, ub);
  exclusive_scan_i64(ub, sbase, (size_t)n + 1, s);
  long long cap = sbase.read(n);
  IBuf scol((size_t)cap, s), count(n + 1, s); DBuf sval((size_t)cap, s);
  count.zero();
  // rows with few products: one thread each; rows with many (long A-rows, e.g. R * (A P)): one warp each
  const bool long_rows = (double)A.nnz > 64.0 * n;
  if (long_rows) {
    static PerDeviceOnce attr_once;
    const size_t smemS = (size_t)((WgSmall::SMEM_PER_WARP + 15) & ~15) * WgSmall::WARPS, smemL = (size_t)((WgLarge::SMEM_PER_WARP + 15) & ~15) * WgLarge::WARPS;
    if (attr_once.first(c.device)) {
      FSB_CUDA(cudaFuncSetAttribute(spgemm_warp_kernel<WgSmall, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemS));
      FSB_CUDA(cudaFuncSetAttribute(spgemm_warp_kernel<WgLarge, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemL));
      FSB_CUDA(cudaFuncSetAttribute(spgemm_warp_kernel<WgLarge, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemL));
    }
    const char* env = getenv("FSB_SPGEMM_LARGE");  // test knob, read at every call: 1 = every row through the large instantiation
    if (env && atoi(env) == 1)
      spgemm_warp_kernel<WgLarge, false><<<cdiv(n, WgLarge::WARPS), 32 * WgLarge::WARPS, smemL, s>>>(n, A.ptr, A.col, A.val, B.ptr, B.col, B.val, sbase, scol, sval, count);
    else {
      spgemm_warp_kernel<WgSmall, false><<<cdiv(n, WgSmall::WARPS), 32 * WgSmall::WARPS, smemS, s>>>(n, A.ptr, A.col, A.val, B.ptr, B.col, B.val, sbase, scol, sval, count);
      // rows it flagged (long A-rows / many distinct columns): the large instantiation, restricted to them
      spgemm_warp_kernel<WgLarge, true><<<cdiv(n, WgLarge::WARPS), 32 * WgLarge::WARPS, smemL, s>>>(n, A.ptr, A.col, A.val, B.ptr, B.col, B.val, sbase, scol, sval, count);
    }
    // rows that exceed that too: thread-per-row kernel, restricted to them
    spgemm_rows<<<cdiv(n, ROW_THREADS), ROW_THREADS, 0, s>>>(n, A.ptr, A.col, A.val, B.ptr, B.col, B.val, sbase, scol, sval, count, 1);
  } else {
    spgemm_rows<<<cdiv(n, ROW_THREADS), ROW_THREADS, 0, s>>>(n, A.ptr, A.col, A.val, B.ptr, B.col, B.val, sbase, scol, sval, count, 0);
  }
  FSB_CHECK_LAUNCH();
  C.nrows = n; C.ncols = B.ncols;
  C.nnz = counts_to_ptr(c, count, n, C.ptr);
  C.col.alloc(C.nnz, s); C.val.alloc(C.nnz, s);
  compact_rows<<<cdiv(n, 128), 128, 0, s>>>(n, sbase, scol, sval, C.ptr, C.col, C.val);
  FSB_CHECK_LAUNCH();
}

void dense_inverse(const Ctx& c, const DCsr& A, DBuf& Ainv) {
  cudaStream_t s = c.stream;
  int n = A.nrows;
  DBuf M((size_t)n * 2 * n, s);
  M.zero();
  Ainv.alloc((size_t)n * n, s);
  densify<<<cdiv(n, 128), 128, 0, s>>>(n, A.ptr, A.col, A.val, M);
  gauss_jordan<<<1, 1024, 0, s>>>(n, M, Ainv);
  FSB_CHECK_LAUNCH();
}


// ---------------------------------------------------------------------------------------------
// Partition split.  The reference splits the permuted matrix into A_in (intra-partition strict
// upper triangle, packed 16|16 local coordinates, smoothedMG_amg_level.cu:24-43, 199-304) and
// A_out (inter-partition entries).  Here: A_out as CSR over all rows, and the intra-partition
// off-diagonal entries of BOTH triangles (atomic-free smoother) as length-sorted, warp-sliced ELL slabs
// with 16-bit columns: register-resident on the fine level (one lane per row), shared-memory resident
// with G lanes per row on coarser levels with many partitions; see "sorted, warp-sliced ELL slabs" below.
// ---------------------------------------------------------------------------------------------
namespace {

__global__ void row_partition_kernel(int nparts, const int* __restrict__ pstart, int* __restrict__ rowPart) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nparts) return;
  for (int r = pstart[p]; r < pstart[p + 1]; r++) rowPart[r] = p;
}

__global__ void count_in_out_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ col, const int* __restrict__ rowPart,
                                    const int* __restrict__ pstart, int* __restrict__ nin, int* __restrict__ nout, int* __restrict__ partK) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int p = rowPart[r], r0 = pstart[p], np = pstart[p + 1] - r0, ci = 0, co = 0;
  for (int e = ptr[r]; e < ptr[r + 1]; e++) {
    int c = col[e];
    if ((unsigned)(c - r0) < (unsigned)np) { if (c != r) ci++; } else co++;
  }
  nin[r] = ci; nout[r] = co;
  atomicMax(&partK[p], ci);
}

__global__ void fill_out_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ col, const double* __restrict__ val,
                                const int* __restrict__ rowPart, const int* __restrict__ pstart, const int* __restrict__ optr,
                                int* __restrict__ ocol, double* __restrict__ oval) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int p = rowPart[r], r0 = pstart[p], np = pstart[p + 1] - r0, o = optr[r];
  for (int e = ptr[r]; e < ptr[r + 1]; e++) {
    int c = col[e];
    if ((unsigned)(c - r0) >= (unsigned)np) { ocol[o] = c; oval[o] = val[e]; o++; }
  }
}

__global__ void chunk_nnz_kernel(int nparts, int C, const int* __restrict__ pstart, const int* __restrict__ ptr, int* __restrict__ cnnz) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nparts * C) return;
  int p = q / C, c = q % C;
  int r0 = pstart[p], np = pstart[p + 1] - r0, chunk = (np + C - 1) / C;
  int m0 = min(c * chunk, np), m1 = min(m0 + chunk, np);
  cnnz[q] = ptr[r0 + m1] - ptr[r0 + m0];
}

// ---- sorted, warp-sliced ELL slabs of the fine-level smoother -------------------------------------
// Inside a partition the rows are handled in order of decreasing intra-partition length (stable), so
// the 32 rows of a warp have (almost) the same length and the slab of a warp is 32 x Kw with
// Kw = length of its longest row rounded up to even: no partition-wide padding.  Columns are stored
// as positions in that sorted order (the x tile of the kernel lives in thread order).

__global__ void sort_key_kernel(int n, const int* __restrict__ rowPart, const int* __restrict__ nin, uint64_t* __restrict__ keys,
                                uint32_t* __restrict__ rows) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  keys[r] = ((uint64_t)rowPart[r] << 8) | (uint64_t)(255 - min(nin[r], 255));
  rows[r] = (uint32_t)r;
}

// sorted index i (partition-contiguous) -> local row (ellrow), local row -> sorted position (pos), sorted length
__global__ void sorted_maps_kernel(int n, const uint32_t* __restrict__ order, const int* __restrict__ rowPart, const int* __restrict__ pstart,
                                   const int* __restrict__ nin, unsigned short* __restrict__ ellrow, unsigned short* __restrict__ pos,
                                   int* __restrict__ lenSorted) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r = (int)order[i], r0 = pstart[rowPart[r]];
  ellrow[i] = (unsigned short)(r - r0);
  pos[r] = (unsigned short)(i - r0);
  lenSorted[i] = nin[r];
}

__global__ void warps_of_partition(int nparts, int G, const int* __restrict__ pstart, int* __restrict__ cnt) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nparts) cnt[p] = ((pstart[p + 1] - pstart[p]) * G + 31) >> 5;
  if (p == nparts) cnt[p] = 0;
}

__global__ void warp_slab_size(int nparts, int G, const int* __restrict__ pstart, const int* __restrict__ pwarp, const int* __restrict__ lenSorted,
                               int* __restrict__ warpPart, long long* __restrict__ sz) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nparts) return;
  int r0 = pstart[p], np = pstart[p + 1] - r0, w0 = pwarp[p];
  for (int j = 0; j * 32 < np * G; j++) {
    // rows are sorted by decreasing length: the first row of the warp is its longest; G lanes share a row
    int K = (lenSorted[r0 + 32 * j / G] + G - 1) / G;
    K += K & 1;
    warpPart[w0 + j] = p;
    sz[w0 + j] = 32LL * K;
  }
}

// one thread per (warp slab, lane): real entries first (column order), then zero padding that reads the row's own x
__global__ void fill_ell_kernel(long long nthreads, int G, const int* __restrict__ warpPart, const int* __restrict__ pstart, const int* __restrict__ pwarp,
                                const long long* __restrict__ wptr, const unsigned short* __restrict__ ellrow,
                                const unsigned short* __restrict__ pos, const int* __restrict__ ptr, const int* __restrict__ col,
                                const double* __restrict__ val, double* __restrict__ ellval, unsigned short* __restrict__ ellcol) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= nthreads) return;
  int w = (int)(gid >> 5), lane = (int)(gid & 31), p = warpPart[w];
  int r0 = pstart[p], np = pstart[p + 1] - r0, t = 32 * (w - pwarp[p]) + lane;  // t: virtual row = (sorted row, lane of the row)
  long long base = wptr[w] + lane;
  int K = (int)((wptr[w + 1] - wptr[w]) >> 5), k = 0;
  const int srow = t / G, g = t % G;
  if (srow < np) {
    int r = r0 + ellrow[r0 + srow], idx = 0;
    for (int e = ptr[r]; e < ptr[r + 1]; e++) {
      int c = col[e];
      if ((unsigned)(c - r0) < (unsigned)np && c != r) {
        if (idx % G == g) {  // lane g of the row takes its entries g, g + G, ...
          ellval[base + 32LL * k] = val[e];
          ellcol[base + 32LL * k] = pos[c];
          k++;
        }
        idx++;
      }
    }
  }
  for (; k < K; k++) { ellval[base + 32LL * k] = 0.0; ellcol[base + 32LL * k] = (unsigned short)min(srow, np - 1); }
}

// Slot and copy assignment.  The x tile of the smoother holds two copies of x whose bank mapping
// differs by half a bank period (copy A: bank = col mod 16, copy B: bank = (col + 8) mod 16; bit 15 of the
// stored column selects B).  A half-warp's gather of slot k is conflict-free when its 16 lanes hit 16
// different 8-byte banks or share addresses (broadcast).  The order of a row's entries inside the slab
// is free, so per half-warp the rows are placed one after the other: an entry first looks for a slot
// in which an earlier lane already reads the same address, then for a slot whose bank (either copy) is
// still unused, then for the least loaded one; padding slots point at an unused bank.
// One thread per (warp slab, half): setup-time work.
// Storage: the per-thread tables (2.2 KB per thread as local arrays) thrashed L1 at full occupancy — every access went to
// L2 and the kernel was one long latency chain (3.4 ms on level 0 of N=118).  They now live in shared memory, thread-fastest
// (index * AS_THREADS + thread), sized by the level's widest slab (MAXK = 16: 68 KB per 64-thread CTA, three CTAs per SM).
constexpr int AS_THREADS = 64;
template <int MAXK>
__global__ void __launch_bounds__(AS_THREADS) ell_assign_slots(int nwarps, int G, const int* __restrict__ warpPart, const int* __restrict__ pstart,
                                                               const int* __restrict__ pwarp, const long long* __restrict__ wptr,
                                                               const int* __restrict__ lenSorted, double* __restrict__ ellval,
                                                               unsigned short* __restrict__ ellcol) {
  extern __shared__ __align__(16) unsigned char as_smem[];
  double* s_ev = reinterpret_cast<double*>(as_smem);                        // [MAXK][T] entries of the current row, as stored
  double* s_ov = s_ev + MAXK * AS_THREADS;                                  // [MAXK][T] ... as placed
  short* s_elem = reinterpret_cast<short*>(s_ov + MAXK * AS_THREADS);       // [MAXK * 16][T] first address assigned to (slot, bank), -1 = unused
  unsigned short* s_ec = reinterpret_cast<unsigned short*>(s_elem + MAXK * 16 * AS_THREADS);  // [MAXK][T]
  unsigned short* s_oc = s_ec + MAXK * AS_THREADS;                          // [MAXK][T]
  unsigned char* s_cnt = reinterpret_cast<unsigned char*>(s_oc + MAXK * AS_THREADS);          // [MAXK * 16][T] distinct addresses in the bank
  const int tid = threadIdx.x;
#define bankElem(k, q) s_elem[((k) * 16 + (q)) * AS_THREADS + tid]
#define bankCnt(k, q) s_cnt[((k) * 16 + (q)) * AS_THREADS + tid]
#define EC(e) s_ec[(e) * AS_THREADS + tid]
#define OC(e) s_oc[(e) * AS_THREADS + tid]
#define EV(e) s_ev[(e) * AS_THREADS + tid]
#define OV(e) s_ov[(e) * AS_THREADS + tid]
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  int w = gid >> 1, h = gid & 1;
  if (w >= nwarps) return;
  int p = warpPart[w], r0 = pstart[p], np = pstart[p + 1] - r0, t0 = 32 * (w - pwarp[p]) + 16 * h;
  if (t0 >= np * G) return;
  const int K = (int)((wptr[w + 1] - wptr[w]) >> 5), nl = min(16, np * G - t0);
  if (K == 0 || K > MAXK) return;
  const long long base = wptr[w] + 16 * h;
  unsigned occ[16];                // per bank: slots in which it is already used (bit k <=> bankCnt[k][bank] > 0)
  for (int k = 0; k < K; k++)
    for (int q = 0; q < 16; q++) { bankElem(k, q) = -1; bankCnt(k, q) = 0; }
  for (int q = 0; q < 16; q++) occ[q] = 0u;
  for (int l = 0; l < nl; l++) {
    const int srow = (t0 + l) / G, g = (t0 + l) % G;
    const int len = min(max(lenSorted[r0 + srow] - g + G - 1, 0) / G, K);
    for (int e = 0; e < len; e++) { EC(e) = ellcol[base + 32LL * e + l]; EV(e) = ellval[base + 32LL * e + l]; }
    unsigned freeSlots = K == 32 ? 0xffffffffu : ((1u << K) - 1u), todo = len == 32 ? 0xffffffffu : ((1u << len) - 1u);
    auto place = [&](int e, int k, int copy) {
      const int cc = EC(e), b = (cc + 8 * copy) & 15;
      const short addr = (short)(cc | (copy << 15));
      OC(k) = (unsigned short)addr; OV(k) = EV(e);
      freeSlots &= ~(1u << k); todo &= ~(1u << e);
      if (bankElem(k, b) != addr) { bankCnt(k, b)++; occ[b] |= 1u << k; if (bankElem(k, b) == -1) bankElem(k, b) = addr; }
    };
    for (int e = 0; e < len; e++) {  // 1. share an address with an earlier lane (only slots whose bank is in use can match)
      const int cc = EC(e), bA = cc & 15, bB = (cc + 8) & 15;
      for (unsigned f = freeSlots & (occ[bA] | occ[bB]); f; f &= f - 1) {
        int k = __ffs(f) - 1;
        if (bankElem(k, bA) == (short)cc) { place(e, k, 0); break; }
        if (bankElem(k, bB) == (short)(cc | 0x8000)) { place(e, k, 1); break; }
      }
    }
    for (int e = 0; e < len; e++) {  // 2. an unused bank: the lowest free slot in which copy A's or copy B's bank is unused
      if (!(todo >> e & 1u)) continue;
      const int cc = EC(e), bA = cc & 15, bB = (cc + 8) & 15;
      const unsigned fa = freeSlots & ~occ[bA], fb = freeSlots & ~occ[bB];
      if (fa | fb) {
        const int k = __ffs(fa | fb) - 1;
        place(e, k, (fa >> k & 1u) ? 0 : 1);
      }
    }
    for (int e = 0; e < len; e++) {  // 3. the least loaded bank
      if (!(todo >> e & 1u)) continue;
      const int cc = EC(e);
      int best = 1 << 30, bk = 0, bc = 0;
      for (unsigned f = freeSlots; f; f &= f - 1) {
        int k = __ffs(f) - 1;
        for (int copy = 0; copy < 2; copy++) {
          int load = bankCnt(k, (cc + 8 * copy) & 15);
          if (load < best) { best = load; bk = k; bc = copy; }
        }
      }
      place(e, bk, bc);
    }
    for (unsigned f = freeSlots; f; f &= f - 1) {  // padding: 0 * x[some unused bank]
      int k = __ffs(f) - 1, cc = srow;
      for (int q = 0; q < 16; q++)
        if (!(occ[q] >> k & 1u) && q < np) { cc = q; break; }
      const int b = cc & 15;
      OC(k) = (unsigned short)cc; OV(k) = 0.0;
      if (bankElem(k, b) != (short)cc) { bankCnt(k, b)++; occ[b] |= 1u << k; if (bankElem(k, b) == -1) bankElem(k, b) = (short)cc; }
    }
    for (int k = 0; k < K; k++) { ellcol[base + 32LL * k + l] = OC(k); ellval[base + 32LL * k + l] = OV(k); }
  }
#undef bankElem
#undef bankCnt
#undef EC
#undef OC
#undef EV
#undef OV
}
template <int MAXK>
constexpr size_t assign_slots_smem() { return (size_t)AS_THREADS * (MAXK * 8 * 2 + MAXK * 16 * 2 + MAXK * 2 * 2 + MAXK * 16); }

}  // namespace

void split_partitions(const Ctx& c, LevelData& L) {
  cudaStream_t s = c.stream;
  const int n = L.n, np = L.nparts;
  static const bool trace = getenv("FSB_SETUP_TRACE") != nullptr;  // tools: wall-clock of the sub-steps on stderr
  double tlast = 0.0;
  auto lap = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(s);
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    const double now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    if (tlast > 0.0) fprintf(stderr, "[split L%d] %-12s %8.3f ms\n", L.level_id, what, now - tlast);
    tlast = now;
  };
  lap("start");
  IBuf rowPart(n, s), nin(n, s), nout(n + 1, s), partK(np, s);
  nout.zero(); partK.zero();
  row_partition_kernel<<<cdiv(np, 128), 128, 0, s>>>(np, L.pstart, rowPart);
  count_in_out_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, L.A.ptr, L.A.col, rowPart, L.pstart, nin, nout, partK);
  L.Aout.nrows = n; L.Aout.ncols = n;
  L.Aout.ptr.alloc(n + 1, s);
  exclusive_scan_i32(nout, L.Aout.ptr, n + 1, s);
  L.Aout.nnz = L.Aout.ptr.read(n);
  L.Aout.col.alloc(std::max(1, L.Aout.nnz), s); L.Aout.val.alloc(std::max(1, L.Aout.nnz), s);
  fill_out_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, L.A.ptr, L.A.col, L.A.val, rowPart, L.pstart, L.Aout.ptr, L.Aout.col, L.Aout.val);
  L.ellMaxK = reduce_max_i32(partK, np, s);
  lap("A_out");
  double avg = (double)L.A.nnz / std::max(1, n);
  L.coopG = avg > 192 ? 32 : avg > 96 ? 16 : avg > 48 ? 8 : avg > 16 ? 4 : 2;
  if (const char* g = getenv("FSB_COOP_G")) L.coopG = std::max(1, std::min(32, atoi(g)));  // tuning knob
  L.use_ell = (L.ellMaxK <= 16) || (L.ellMaxK <= 32 && L.maxPartRows <= 512);
  for (int q = 0; q < kEllClasses; q++) L.nlist[q] = 0;
  L.smemBytes = 0;
  if (!L.use_ell) {
    // smallest cluster size whose per-CTA slice fits comfortably (two CTAs per SM), else the
    // largest portable cluster if that fits at all; otherwise the L1/L2-streaming fallback
    for (int pass = 0; pass < 2 && L.smemBytes == 0; pass++) {
      for (int C = 1; C <= 8; C *= 2) {
        IBuf cnnz((size_t)np * C, s);
        chunk_nnz_kernel<<<cdiv((long long)np * C, 256), 256, 0, s>>>(np, C, L.pstart, L.A.ptr, cnnz);
        int mx = reduce_max_i32(cnnz, (size_t)np * C, s);
        int chunkRows = (L.maxPartRows + C - 1) / C;
        // layout of smooth_cluster_kernel: val[cap] | 2 x double[npmax] | 3 x double[chunkmax] | int[chunkmax+1] | u16[cap]
        size_t bytes = (size_t)mx * 8 + (size_t)L.maxPartRows * 16 + (size_t)chunkRows * 24 + ((size_t)chunkRows + 1) * 4 + (size_t)mx * 2 + 32;
        static const char* env_kb = getenv("FSB_CLUSTER_LIMIT_KB");  // tuning knob (default 110 KB: two CTAs per SM)
        size_t limit = pass == 0 ? (size_t)(env_kb ? atoi(env_kb) : 110) * 1024 : 220 * 1024;
        if (bytes <= limit) { L.clusterC = C; L.maxChunkNnz = mx; L.maxChunkRows = chunkRows; L.smemBytes = (int)bytes; break; }
      }
    }
  }
  // levels that do not take the register-resident kernel: the same sorted slabs with G lanes per row,
  // kept in shared memory by smooth_sellg_kernel (one CTA per partition)
  L.ellG = 1;
  L.use_sellg = false;
  {
    // tuning / test knob, read at every setup: 0 = cluster or cooperative kernels instead, 2 = also on levels with few partitions
    const char* env = getenv("FSB_SELLG");
    const int mode = env ? atoi(env) : 1;
    // (with few partitions the cluster kernel, which spreads a partition over several SMs, is the better fit)
    if (!L.use_ell && L.maxPartRows <= 1024 && (2 * np >= c.num_sms || mode == 2) && mode != 0) {
      const double meanRows = (double)n / std::max(1, np);
      int G = 1;
      const double target = np > c.num_sms ? 512.0 : 1024.0;          // threads of the CTA that will own a partition
      while (G < 32 && meanRows * (2 * G) <= target * 1.25) G *= 2;   // about one virtual row per thread
      while (G < 32 && (L.ellMaxK + G - 1) / G > 30) G *= 2;          // at most 32 slots per lane (even-rounded)
      if ((L.ellMaxK + G - 1) / G <= 30 && (long long)L.maxPartRows * G <= 4 * (long long)target) { L.use_sellg = true; L.ellG = G; }
    }
  }
  lap("cluster_cfg");
  if (L.use_sellg) L.smemBytes = 0;  // the cluster kernel is not used
  if (L.use_ell || L.use_sellg) {
    const int G = L.ellG;
    // rows of a partition by decreasing intra-partition length (stable radix sort of (partition, 255 - length))
    DevBuf<uint64_t> k0(n, s), k1(n, s);
    DevBuf<uint32_t> v0(n, s), order(n, s);
    sort_key_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, rowPart, nin, k0, v0);
    sort_pairs_u64_u32(k0, k1, v0, order, n, bits_for(np) + 8, s);
    DevBuf<unsigned short> pos(n, s);
    IBuf lenSorted(n, s);
    L.ellrow.alloc(n, s);
    sorted_maps_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, order, rowPart, L.pstart, nin, L.ellrow, pos, lenSorted);
    lap("sort_rows");
    // warp slabs
    IBuf wcnt((size_t)np + 1, s);
    warps_of_partition<<<cdiv(np + 1, 256), 256, 0, s>>>(np, G, L.pstart, wcnt);
    L.pwarp.alloc((size_t)np + 1, s);
    exclusive_scan_i32(wcnt, L.pwarp, (size_t)np + 1, s);
    const int nw = L.pwarp.read(np);
    L.ellWarps = nw;
    IBuf warpPart(std::max(1, nw), s);
    DevBuf<long long> sz((size_t)nw + 1, s);
    sz.zero();
    warp_slab_size<<<cdiv(np, 128), 128, 0, s>>>(np, G, L.pstart, L.pwarp, lenSorted, warpPart, sz);
    L.ellwptr.alloc((size_t)nw + 1, s);
    void* tmp = nullptr; size_t bytes = 0;
    FSB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, sz.get(), L.ellwptr.get(), nw + 1, s));
    FSB_CUDA(cudaMallocAsync(&tmp, bytes, s));
    FSB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, sz.get(), L.ellwptr.get(), nw + 1, s));
    cudaFreeAsync(tmp, s);
    const long long total = L.ellwptr.read(nw);
    L.ellSlots = total;
    L.ellval.alloc((size_t)std::max<long long>(total, 1), s);
    L.ellcol.alloc((size_t)std::max<long long>(total, 1), s);
    lap("slab_ptrs");
    fill_ell_kernel<<<cdiv(32LL * nw, 256), 256, 0, s>>>(32LL * nw, G, warpPart, L.pstart, L.pwarp, L.ellwptr, L.ellrow, pos, L.A.ptr, L.A.col, L.A.val,
                                                         L.ellval, L.ellcol);
    lap("fill");
    {
      // widest slab of the level: K = ceil(longest row / G) rounded up to even
      int kmax = (L.ellMaxK + G - 1) / G;
      kmax += kmax & 1;
      static PerDeviceOnce attr_once;
      if (attr_once.first(c.device)) {
        FSB_CUDA(cudaFuncSetAttribute(ell_assign_slots<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)assign_slots_smem<16>()));
        FSB_CUDA(cudaFuncSetAttribute(ell_assign_slots<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)assign_slots_smem<32>()));
      }
      if (kmax <= 16)
        ell_assign_slots<16><<<cdiv(2LL * nw, AS_THREADS), AS_THREADS, assign_slots_smem<16>(), s>>>(nw, G, warpPart, L.pstart, L.pwarp, L.ellwptr, lenSorted, L.ellval, L.ellcol);
      else
        ell_assign_slots<32><<<cdiv(2LL * nw, AS_THREADS), AS_THREADS, assign_slots_smem<32>(), s>>>(nw, G, warpPart, L.pstart, L.pwarp, L.ellwptr, lenSorted, L.ellval, L.ellcol);
    }
    lap("assign_slots");
    // per-partition descriptors and the lists by size class (host side: nparts is a few thousand)
    std::vector<int> ps = L.pstart.to_vector(), pw = L.pwarp.to_vector();
    std::vector<long long> wp = L.ellwptr.to_vector();
    if (L.use_ell) {
      L.ellDescHost.assign(np, EllDesc());
      std::vector<EllDesc> lists[kEllClasses];
      for (int p = 0; p < np; p++) {
        EllDesc& d = L.ellDescHost[p];
        d.r0 = ps[p]; d.np = ps[p + 1] - ps[p]; d.base = wp[pw[p]];
        for (int j = 0; j < 32; j++) d.K[j] = (pw[p] + j < pw[p + 1]) ? (unsigned char)((wp[pw[p] + j + 1] - wp[pw[p] + j]) >> 5) : 0;
        lists[ell_class(d.np)].push_back(d);
      }
      for (int q = 0; q < kEllClasses; q++) {
        L.nlist[q] = (int)lists[q].size();
        L.plist[q].alloc(std::max<size_t>(1, lists[q].size()), s);
        if (!lists[q].empty()) L.plist[q].from_host(lists[q].data(), lists[q].size());
      }
    } else {
      std::vector<SellgDesc> dl(np);
      long long maxSlots = 0;
      for (int p = 0; p < np; p++) {
        SellgDesc& d = dl[p];
        d.r0 = ps[p]; d.np = ps[p + 1] - ps[p]; d.w0 = pw[p]; d.nw = pw[p + 1] - pw[p];
        d.base = wp[pw[p]]; d.slots = (int)(wp[pw[p + 1]] - wp[pw[p]]); d.pad = 0;
        maxSlots = std::max<long long>(maxSlots, d.slots);
      }
      L.sellgDesc.alloc(std::max(1, np), s);
      if (np) L.sellgDesc.from_host(dl.data(), np);
      L.sellgMaxSlots = (int)maxSlots;
    }
    FSB_CUDA(cudaStreamSynchronize(s));
    lap("descriptors");
  }
  FSB_CHECK_LAUNCH();
}


// ---------------------------------------------------------------------------------------------
// CSR -> SELL-32 (the streaming format of the fine-level SpMV family)
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void sell_width_kernel(int n, int nslices, const int* __restrict__ ptr, const int* __restrict__ rowmap, long long* __restrict__ sz) {
  int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (s >= nslices) return;
  int i = s * 32 + lane, r = i < n ? (rowmap ? rowmap[i] : i) : -1;
  int len = r >= 0 ? ptr[r + 1] - ptr[r] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
  if (lane == 0) sz[s] = 32ll * len;
}
__global__ void sell_fill_kernel(int n, int ncols, int nslices, const int* __restrict__ ptr, const int* __restrict__ col, const double* __restrict__ val,
                                 const int* __restrict__ rowmap, const long long* __restrict__ sptr, int* __restrict__ scol, double* __restrict__ sval) {
  int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (s >= nslices) return;
  long long base = sptr[s];
  int K = (int)((sptr[s + 1] - base) >> 5), i = s * 32 + lane;
  int r = i < n ? (rowmap ? rowmap[i] : i) : -1;
  int e0 = r >= 0 ? ptr[r] : 0, len = r >= 0 ? ptr[r + 1] - e0 : 0;
  int padcol = min(max(r, 0), ncols - 1);  // any valid column: the padded value is 0
  for (int k = 0; k < K; k++) {
    bool has = k < len;
    scol[base + (long long)k * 32 + lane] = has ? col[e0 + k] : padcol;
    sval[base + (long long)k * 32 + lane] = has ? val[e0 + k] : 0.0;
  }
}
// sort key: window id in the high bits, row length (capped) in the low 8 bits
__global__ void sell_sort_keys_kernel(int n, int window, const int* __restrict__ ptr, int* __restrict__ keys) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) keys[r] = ((r / window) << 8) | min(ptr[r + 1] - ptr[r], 255);
}
}  // namespace

void build_sell(const Ctx& c, const DCsr& A, Sell& S, int sort_window) {
  cudaStream_t s = c.stream;
  S.nrows = A.nrows; S.ncols = A.ncols; S.nslices = (A.nrows + 31) / 32; S.window = sort_window;
  if (S.nslices == 0) return;
  const int* rowmap = nullptr;
  if (sort_window > 0) {  // SELL-32-sigma: rows ordered by length inside windows, so short rows do not pay for long ones
    IBuf keys(A.nrows, s), keys2(A.nrows, s), iota(A.nrows, s);
    sell_sort_keys_kernel<<<cdiv(A.nrows, 256), 256, 0, s>>>(A.nrows, sort_window, A.ptr, keys);
    iota_i32(iota, A.nrows, s);
    S.rowmap.alloc(A.nrows, s);
    sort_pairs_i32_i32(keys, keys2, iota, S.rowmap, A.nrows, 8 + bits_for(A.nrows / sort_window + 1), s);
    rowmap = S.rowmap;
  }
  DevBuf<long long> sz((size_t)S.nslices + 1, s);
  sz.zero();
  sell_width_kernel<<<cdiv(S.nslices, 8), 256, 0, s>>>(A.nrows, S.nslices, A.ptr, rowmap, sz);
  S.sptr.alloc((size_t)S.nslices + 1, s);
  exclusive_scan_i64(sz, S.sptr, (size_t)S.nslices + 1, s);
  S.nstored = S.sptr.read(S.nslices);
  S.col.alloc((size_t)std::max<long long>(S.nstored, 1), s);
  S.val.alloc((size_t)std::max<long long>(S.nstored, 1), s);
  sell_fill_kernel<<<cdiv(S.nslices, 8), 256, 0, s>>>(A.nrows, A.ncols, S.nslices, A.ptr, A.col, A.val, rowmap, S.sptr, S.col, S.val);
  FSB_CHECK_LAUNCH();
}

}  // namespace fsb
