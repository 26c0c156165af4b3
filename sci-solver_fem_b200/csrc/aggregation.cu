// aggregation.cu — stage 2a: two-level MIS aggregation (aggregatorType_ = 0, "OldMIS") on device.
//
// Reference: misHelpers::CP::OldMIS (src/core/cuda/ComputePermutationMethods.cu:22-150),
// randomizedMIS (randomizedMIS_GPU.cu:3-272), aggregateGraph / aggregateWeightedGraph /
// restrictPartitionSize / removeRunty* / getInducedGraph / remapInducedGraph
// (misHelpers.cu:13-75, 297-412, 513-820, 1078-1107, 1258-1280).
//
// All outputs (permutation, ipermutation, aggregateIdx, partitionIdx, partitionLabel, induced
// graph) are integer arrays that must be BIT-EXACT with the reference algorithm for a given
// seed; every kernel below keeps the reference's integer semantics (including its degenerate
// "first labelled neighbour wins" vote and the fp32 desirability arithmetic), while the
// implementation is restructured: histograms instead of sort+boundary-search for part sizes,
// a packed 64-bit atomicMax instead of a tuple sort for the per-partition best swap, packed
// 64-bit radix sorts for the induced graph, device-side convergence counters.
#include <cstdlib>
#include <ctime>
#include <mutex>

#include "fsb_internal.h"

namespace fsb {
namespace {

// tools: FSB_SETUP_TRACE=1 prints wall-clock laps and loop counts of the aggregation pipeline on stderr
struct AggTrace {
  bool on = getenv("FSB_SETUP_TRACE") != nullptr;
  double last = 0.0;
  cudaStream_t s;
  explicit AggTrace(cudaStream_t st) : s(st) { lap(nullptr, 0); }
  void lap(const char* what, long long count) {
    if (!on) return;
    cudaStreamSynchronize(s);
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    const double now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    if (what) fprintf(stderr, "[agg] %-22s %8.3f ms  (%lld)\n", what, now - last, count);
    last = now;
  }
};
long long g_trace_count[4] = {0, 0, 0, 0};  // MIS rounds, allocate sweeps, restrict iterations, runt removals

// ---------------------------------------------------------------- randomized distance-k MIS
__device__ __forceinline__ unsigned taus_step(unsigned z) {
  unsigned b = (((z << 13) ^ z) >> 19);
  return ((z << 12) ^ b);
}

__global__ void mis_gen_randoms(int n, int iterations, unsigned* __restrict__ randoms, const unsigned* __restrict__ seeds) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;  // exactly 32768 generator threads (randomizedMIS_GPU.cu:201)
  if (t >= 32768) return;
  unsigned z = seeds[t];
  int off = t;
  for (int i = 0; i < iterations; i++)
    if (off < n) { z = taus_step(z); randoms[off] = z; off += 32768; }
}

__global__ void mis_first_init(int n, const unsigned* __restrict__ randoms, int* __restrict__ best, int* __restrict__ origin) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) { origin[v] = v; best[v] = (int)(randoms[v] % 1000000u); }
}

__global__ void mis_reinit(int n, unsigned* __restrict__ randoms, int* __restrict__ best, int* __restrict__ origin,
                           const int* __restrict__ mis, int* __restrict__ incomplete) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) {
    unsigned z = taus_step(randoms[v]);
    origin[v] = v;
    int st = mis[v];
    best[v] = (st == -1) ? (int)(z % 1000000u) : (st == 1 ? 1000001 : 0);
    randoms[v] = z;
  }
  if (v == 0) incomplete[0] = 0;
}

// one propagation pass; FINAL also classifies (Iterate_Kernel / Final_Iterate_Kernel :65-152)
template <bool FINAL>
__global__ void mis_iterate(int n, const int* __restrict__ originIn, int* __restrict__ originOut, const int* __restrict__ bestIn,
                            int* __restrict__ bestOut, const int* __restrict__ xadj, const int* __restrict__ adj,
                            int* __restrict__ mis, int* __restrict__ incomplete) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  int b = bestIn[v], o = originIn[v];
  if (b < 1000001) {
    int e1 = xadj[v + 1];
    for (int e = xadj[v]; e < e1; e++) {
      int nb = adj[e];
      int ch = bestIn[nb], cho = originIn[nb];
      if (ch > 0 && ch == b && cho > o) o = cho;
      if (ch > b) { b = ch; o = cho; }
    }
  }
  bestOut[v] = b;
  originOut[v] = o;
  if (FINAL) {
    int st = -1;
    if (o == v) st = 1; else if (b == 1000001) st = 0;
    mis[v] = st;
    if (st == -1) incomplete[0] = 1;
  }
}

void randomized_mis(const Ctx& c, int n, const int* xadj, const int* adj, int k, unsigned seed, int* mis) {
  cudaStream_t s = c.stream;
  // upstream: srand(time(NULL)) and 32768 rand() calls per MIS; the seed is a solver parameter here, so the table is a pure
  // function of it: six MIS calls per setup share one table (libc's rand() takes a lock: 0.7 ms per table), and concurrent
  // solvers on other threads cannot interleave their rand() sequences.
  std::vector<unsigned> seeds_h;
  {
    static std::mutex mu;
    static std::vector<unsigned> table;
    static unsigned table_seed = 0;
    std::lock_guard<std::mutex> lock(mu);
    if (table.empty() || table_seed != seed) {
      table.resize(32768);
      srand(seed);
      for (int i = 0; i < 32768; i++) table[i] = (unsigned)rand();
      table_seed = seed;
    }
    seeds_h = table;
  }
  DevBuf<unsigned> seeds(32768, s), randoms(n, s);
  seeds.from_host(seeds_h.data(), 32768);
  IBuf bestA(n, s), bestB(n, s), orgA(n, s), orgB(n, s), incomplete(1, s);
  incomplete.zero();
  fill_i32(mis, n, -1, s);
  int nb = cdiv(n, 256);
  mis_gen_randoms<<<128, 256, 0, s>>>(n, (n + 32767) / 32768, randoms, seeds);
  mis_first_init<<<nb, 256, 0, s>>>(n, randoms, bestA, orgA);
  int *bi = bestA, *bo = bestB, *oi = orgA, *oo = orgB;
  bool first = true;
  for (int round = 0; round < 100000; round++) {
    if (!first) mis_reinit<<<nb, 256, 0, s>>>(n, randoms, bi, oi, mis, incomplete);
    first = false;
    for (int i = 0; i < k; i++) {
      if (i < k - 1) mis_iterate<false><<<nb, 256, 0, s>>>(n, oi, oo, bi, bo, xadj, adj, mis, incomplete);
      else mis_iterate<true><<<nb, 256, 0, s>>>(n, oi, oo, bi, bo, xadj, adj, mis, incomplete);
      std::swap(bi, bo); std::swap(oi, oo);
    }
    FSB_CHECK_LAUNCH();
    g_trace_count[0]++;
    if (incomplete.read(0) == 0) return;
  }
  throw std::runtime_error("randomizedMIS did not converge");
}

// ---------------------------------------------------------------- growing aggregates
// allocateNodesKernel (misHelpers.cu:13-75).  The 10-slot candidate vote never clears
// `candidate` on an empty slot, so the first labelled neighbour fills every slot and wins;
// the effective rule — first labelled neighbour in adjacency order — is what runs here.
__global__ void allocate_nodes(int n, const int* __restrict__ xadj, const int* __restrict__ adj, const int* __restrict__ partIn,
                               int* __restrict__ partOut, int* __restrict__ aggregated, int* __restrict__ unallocated) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  if (aggregated[v] != 0) return;
  int addTo = -1, e1 = xadj[v + 1];
  for (int e = xadj[v]; e < e1; e++) {
    int cand = partIn[adj[e]];
    if (cand != -1) { addTo = cand; break; }
  }
  partOut[v] = addTo;
  if (addTo != -1) aggregated[v] = 1; else atomicAdd(unallocated, 1);
}

__global__ void label_roots(int n, const int* __restrict__ mis, const int* __restrict__ scan, int* __restrict__ part) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) part[v] = (mis[v] == 0) ? -1 : scan[v] - 1;  // ifLabelOne (misHelpers.h:35-44)
}
__global__ void mark_aggregated(int n, const int* __restrict__ part, int* __restrict__ aggregated) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) aggregated[v] = (part[v] == -1) ? 0 : 1;  // findAggregated (misHelpers.h:64-73)
}
__global__ void hist_weighted(int n, const int* __restrict__ part, const int* __restrict__ w, int* __restrict__ sizes) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) atomicAdd(&sizes[part[v]], w ? w[v] : 1);
}
__global__ void make_stencil(int np, const int* __restrict__ sizes, int threshold, int* __restrict__ stencil) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < np) stencil[p] = sizes[p] < threshold ? 1 : 0;  // labelLessThan
}
__global__ void remove_runts(int n, int* __restrict__ part, const int* __restrict__ stencil, const int* __restrict__ sub) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;  // removeRuntyPartsKernel (misHelpers.cu:252-263)
  if (v >= n) return;
  int cpart = part[v];
  if (stencil[cpart] == 1) part[v] = -1; else part[v] = cpart - sub[cpart];
}

// removes every part whose (weighted) size is below `threshold`; returns the number removed
int remove_small_parts(const Ctx& c, int n, int* part, const int* w, int nparts, int threshold) {
  cudaStream_t s = c.stream;
  IBuf sizes(nparts, s), stencil(nparts, s), sub(nparts, s);
  sizes.zero();
  hist_weighted<<<cdiv(n, 256), 256, 0, s>>>(n, part, w, sizes);
  make_stencil<<<cdiv(nparts, 256), 256, 0, s>>>(nparts, sizes, threshold, stencil);
  inclusive_scan_i32(stencil, sub, nparts, s);
  int removed = sub.read(nparts - 1);
  if (removed == 0) return 0;
  remove_runts<<<cdiv(n, 256), 256, 0, s>>>(n, part, stencil, sub);
  FSB_CHECK_LAUNCH();
  return removed;
}

// returns the number of roots; part = root numbering (-1 elsewhere); aggregated = MIS stencil
int seed_parts_from_mis(const Ctx& c, int n, const int* xadj, const int* adj, int depth, unsigned seed, IBuf& part, IBuf& aggregated) {
  cudaStream_t s = c.stream;
  part.alloc(n, s); aggregated.alloc(n, s);
  randomized_mis(c, n, xadj, adj, depth, seed, aggregated);
  IBuf scan(n, s);
  inclusive_scan_i32(aggregated, scan, n, s);
  int misCount = scan.read(n - 1);
  label_roots<<<cdiv(n, 256), 256, 0, s>>>(n, aggregated, scan, part);
  FSB_CHECK_LAUNCH();
  return misCount;
}

// one sweep; returns the number of still-unallocated nodes
int allocate_sweep(const Ctx& c, int n, const int* xadj, const int* adj, IBuf& partIn, IBuf& partOut, IBuf& aggregated, IBuf& counter) {
  cudaStream_t s = c.stream;
  counter.zero();
  allocate_nodes<<<cdiv(n, 256), 256, 0, s>>>(n, xadj, adj, partIn, partOut, aggregated, counter);
  FSB_CHECK_LAUNCH();
  partIn.from_device(partOut, n);  // "partIn = partOut"
  g_trace_count[1]++;
  return counter.read(0);
}

// aggregateGraph (misHelpers.cu:513-601); returns the number of aggregates
int aggregate_graph(const Ctx& c, int n, const int* xadj, const int* adj, int minSize, int depth, unsigned seed, IBuf& part) {
  cudaStream_t s = c.stream;
  IBuf aggregated, partOut(n, s), counter(1, s);
  int nparts = seed_parts_from_mis(c, n, xadj, adj, depth, seed, part, aggregated);
  partOut.from_device(part, n);
  for (int guard = 0; guard < 1000000; guard++) {
    int unallocated = allocate_sweep(c, n, xadj, adj, part, partOut, aggregated, counter);
    if (unallocated == 0) {
      int removed = remove_small_parts(c, n, part, nullptr, nparts, minSize);
      if (removed == 0) return nparts;
      nparts -= removed;
      mark_aggregated<<<cdiv(n, 256), 256, 0, s>>>(n, part, aggregated);
      partOut.from_device(part, n);
    }
  }
  throw std::runtime_error("aggregateGraph did not converge");
}

// ---------------------------------------------------------------- partition size restriction
// findDesirabilityKernel (misHelpers.cu:297-378): fp32, rounded exactly where written; the
// `x = a*b; x += c` pair is one fused multiply-add, as nvcc's default contraction emits for
// the upstream kernel (the oracle uses fmaf at the same place).  Instead of sorting all nodes
// by (from-partition, desirability) the best node of each partition is kept with a packed
// 64-bit atomicMax: desirability >= 0 orders like its bit pattern, ties go to the highest
// index — exactly the last element of each group after upstream's stable sort.
__global__ void find_desirability(int n, int optimalSize, const int* __restrict__ xadj, const int* __restrict__ adj,
                                  const int* __restrict__ partition, const int* __restrict__ partSizes, const int* __restrict__ w,
                                  int* __restrict__ swap_to, unsigned long long* __restrict__ bestOfPart) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  int currentPart = partition[idx], cps = partSizes[currentPart], nodeSize = w[idx];
  int selfAdjacency = 0, addTo = -1;
  float best = 0.f;
  float cwf = __fdiv_rn((float)abs(cps - optimalSize), (float)optimalSize);
  float selfImp = __fmul_rn((float)(abs(cps - optimalSize) - abs((cps - nodeSize) - optimalSize)), cwf);
  if (selfImp > 0) {
    int cand[10], cnt[10];
#pragma unroll
    for (int i = 0; i < 10; i++) { cand[i] = -1; cnt[i] = 0; }
    int e1 = xadj[idx + 1];
    for (int e = xadj[idx]; e < e1; e++) {
      int cc = partition[adj[e]];
      if (cc == currentPart) selfAdjacency++;
      else {
#pragma unroll
        for (int j = 0; j < 10; j++) {
          if (cc != -1 && cand[j] == -1) { cand[j] = cc; cnt[j] = 1; cc = -1; }
          else if (cand[j] == cc) { cnt[j] += 1; cc = -1; }
        }
      }
    }
#pragma unroll
    for (int i = 1; i < 10; i++) {  // candidate 0 is skipped upstream (:351)
      if (cand[i] != -1) {
        int np = cand[i], nps = partSizes[np];
        float nwf = __fdiv_rn((float)abs(nps - optimalSize), (float)optimalSize);
        float ni = __fmaf_rn((float)(abs(nps - optimalSize) - abs((nps + nodeSize) - optimalSize)), nwf, selfImp);
        ni = __fmul_rn(ni, __fdiv_rn((float)cnt[i], (float)selfAdjacency));
        if (ni > best) { addTo = np; best = ni; }
      }
    }
  }
  swap_to[idx] = addTo;
  unsigned long long key = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned)idx;
  atomicMax(&bestOfPart[currentPart], key);
}

// makeSwapsKernel (misHelpers.cu:380-412): one candidate per source partition, all applied at once
__global__ void make_swaps(int nparts, int* __restrict__ partition, int* __restrict__ partSizes, const int* __restrict__ w,
                           const int* __restrict__ swap_to, const unsigned long long* __restrict__ bestOfPart) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nparts) return;
  unsigned long long key = bestOfPart[p];
  if (key == 0ull) return;  // partition without nodes, or only node 0 with zero desirability
  float des = __uint_as_float((unsigned)(key >> 32));
  int idx = (int)(key & 0xFFFFFFFFull);
  if ((double)des > .1) {
    int to = swap_to[idx], wt = w[idx];
    partition[idx] = to;
    atomicAdd(&partSizes[to], wt);
    atomicAdd(&partSizes[p], -wt);
  }
}

void restrict_partition_size(const Ctx& c, int n, const int* xadj, const int* adj, int* partition, const int* w, int nparts,
                             int maxSize, int fullSize) {
  cudaStream_t s = c.stream;
  IBuf partSizes(nparts, s), swap_to(n, s);
  DevBuf<unsigned long long> bestOfPart(nparts, s);
  partSizes.zero();
  hist_weighted<<<cdiv(n, 256), 256, 0, s>>>(n, partition, w, partSizes);
  int averageSize = fullSize / nparts;  // integer division, fixed for the whole loop (:689)
  int largest = reduce_max_i32(partSizes, nparts, s);
  for (int guard = 0; largest > maxSize; guard++) {
    if (guard > 100000) throw std::runtime_error("restrictPartitionSize does not terminate");
    bestOfPart.zero();
    find_desirability<<<cdiv(n, 256), 256, 0, s>>>(n, averageSize, xadj, adj, partition, partSizes, w, swap_to, bestOfPart);
    make_swaps<<<cdiv(nparts, 256), 256, 0, s>>>(nparts, partition, partSizes, w, swap_to, bestOfPart);
    FSB_CHECK_LAUNCH();
    g_trace_count[2]++;
    largest = reduce_max_i32(partSizes, nparts, s);
  }
}

// aggregateWeightedGraph (misHelpers.cu:603-677); returns the number of partitions
int aggregate_weighted_graph(const Ctx& c, int n, const int* xadj, const int* adj, const int* w, int maxSize, int fullSize,
                             int depth, unsigned seed, IBuf& part) {
  cudaStream_t s = c.stream;
  IBuf aggregated, partOut(n, s), counter(1, s);
  int misCount = seed_parts_from_mis(c, n, xadj, adj, depth, seed, part, aggregated);
  int nparts = misCount;
  partOut.from_device(part, n);
  bool firstTime = true;
  for (int guard = 0; guard < 1000000; guard++) {
    int unallocated = allocate_sweep(c, n, xadj, adj, part, partOut, aggregated, counter);
    if (unallocated != 0) continue;
    if (!firstTime || misCount < 10) {
      restrict_partition_size(c, n, xadj, adj, part, w, nparts, maxSize, fullSize);
      return nparts;
    }
    firstTime = false;
    // removeRuntyPartitions (misHelpers.cu:783-820): weighted size below 70 % of the average
    double averageSize = (double)fullSize / nparts;
    int threshold = (int)(averageSize * .7);
    nparts -= remove_small_parts(c, n, part, w, nparts, threshold);
    mark_aggregated<<<cdiv(n, 256), 256, 0, s>>>(n, part, aggregated);
    partOut.from_device(part, n);
  }
  throw std::runtime_error("aggregateWeightedGraph did not converge");
}

// ---------------------------------------------------------------- induced graph
__global__ void induced_pairs(int n, const int* __restrict__ xadj, const int* __restrict__ adj, const int* __restrict__ label,
                              uint64_t* __restrict__ keys) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  int tb = label[v], e1 = xadj[v + 1];
  for (int e = xadj[v]; e < e1; e++) {
    int nb = label[adj[e]];
    // (-1,-1) for intra-aggregate edges; +1 bias so the pair packs into an unsigned key
    keys[e] = (tb == nb) ? 0ull : (((uint64_t)(tb + 1) << 32) | (uint64_t)(unsigned)(nb + 1));
  }
}
__global__ void flag_unique64(long long n, const uint64_t* __restrict__ keys, int* __restrict__ flag) {
  long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k < n) flag[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1 : 0;
}
__global__ void compact_unique64(long long n, const uint64_t* __restrict__ keys, const int* __restrict__ pos, uint64_t* __restrict__ out) {
  long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k < n && (k == 0 || keys[k] != keys[k - 1])) out[pos[k] - 1] = keys[k];
}
// findPartIndicesNegStartKernel (misHelpers.cu:91-101) on the unique pair list + adjacency copy
__global__ void induced_offsets(int size, const uint64_t* __restrict__ u, int* __restrict__ xout, int* __restrict__ aout) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (idx >= size) return;
  int value = (int)(u[idx] >> 32) - 1;
  aout[idx - 1] = (int)(u[idx] & 0xFFFFFFFFull) - 1;
  bool last = (idx == size - 1) || ((int)(u[idx + 1] >> 32) - 1 != value);
  if (last) xout[value + 1] = idx;
}

__global__ void count_nonmonotone(int n, const int* __restrict__ off, int* __restrict__ bad) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < n && off[a + 1] < off[a]) atomicAdd(bad, 1);
}

// getInducedGraph (misHelpers.cu:1078-1107).  As upstream, the first unique pair is assumed to
// be the (-1,-1) marker and is dropped unconditionally.  Upstream silently builds non-monotone offsets when an
// aggregate has no edge to any other aggregate (a vertex no element connects to the rest of the mesh, or a small
// disconnected component: findPartIndicesNegStartKernel never writes its offset); here that is an error the caller sees.
void induced_graph(const Ctx& c, int n, const int* xadj, const int* adj, int nedges, const int* label, int nlabels, IBuf& xout, IBuf& aout) {
  cudaStream_t s = c.stream;
  if (nedges <= 0) throw std::invalid_argument("aggregation: the graph of this level has no edges (every vertex is isolated)");
  DevBuf<uint64_t> k0(nedges, s), k1(nedges, s);
  induced_pairs<<<cdiv(n, 256), 256, 0, s>>>(n, xadj, adj, label, k0);
  sort_keys_u64(k0, k1, nedges, 32 + bits_for(nlabels), s);
  IBuf flag(nedges, s), pos(nedges, s);
  flag_unique64<<<cdiv(nedges, 256), 256, 0, s>>>(nedges, k1, flag);
  inclusive_scan_i32(flag, pos, nedges, s);
  int size = pos.read(nedges - 1);
  compact_unique64<<<cdiv(nedges, 256), 256, 0, s>>>(nedges, k1, pos, k0);
  uint64_t lastKey;
  FSB_CUDA(cudaMemcpyAsync(&lastKey, k0.get() + (size - 1), sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  FSB_CUDA(cudaStreamSynchronize(s));
  int maxPart = (int)(lastKey >> 32) - 1;
  xout.alloc(maxPart + 2, s);
  xout.zero();
  aout.alloc(size - 1, s);
  if (size > 1) induced_offsets<<<cdiv(size - 1, 256), 256, 0, s>>>(size, k0, xout, aout);
  int lastv = size - 1;
  FSB_CUDA(cudaMemcpyAsync(xout.get() + (maxPart + 1), &lastv, sizeof(int), cudaMemcpyHostToDevice, s));
  IBuf bad(1, s);
  bad.zero();
  count_nonmonotone<<<cdiv(maxPart + 1, 256), 256, 0, s>>>(maxPart + 1, xout, bad);
  FSB_CHECK_LAUNCH();
  if (bad.read(0) != 0 || maxPart + 1 != nlabels)
    throw std::invalid_argument("aggregation: an aggregate has no edge to the rest of the graph (the mesh has a vertex that no element "
                                "connects to the others, or a small disconnected component); upstream builds a corrupt coarse graph here");
}

// ---------------------------------------------------------------- small index kernels
__global__ void part_indices_kernel(int size, const int* __restrict__ sorted, int* __restrict__ pi) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;  // findPartIndicesKernel (misHelpers.cu:77-89)
  if (idx >= size) return;
  int value = sorted[idx];
  int next = (idx != size - 1) ? sorted[idx + 1] : -1;
  if (value != next) pi[value + 1] = idx + 1;
}
// getPartIndices (misHelpers.cu:874-893): offsets of equal-key runs of an ascending array
void part_indices(const Ctx& c, const int* sorted, int size, int maxPart, IBuf& pi) {
  cudaStream_t s = c.stream;
  pi.alloc(maxPart + 2, s);
  pi.zero();
  part_indices_kernel<<<cdiv(size, 256), 256, 0, s>>>(size, sorted, pi);
  FSB_CUDA(cudaMemcpyAsync(pi.get() + (maxPart + 1), &size, sizeof(int), cudaMemcpyHostToDevice, s));
  FSB_CUDA(cudaStreamSynchronize(s));
}
__global__ void inverse_perm(int n, const int* __restrict__ p, int* __restrict__ inv) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) inv[p[i]] = i;
}
__global__ void gather_i32(int n, const int* __restrict__ idx, const int* __restrict__ src, int* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
__global__ void diff_i32(int n, const int* __restrict__ off, int* __restrict__ sizes) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sizes[i] = off[i + 1] - off[i];
}
__global__ void map_adjacency(int n, const int* __restrict__ xadj, const int* __restrict__ adj, const int* __restrict__ newId,
                              int* __restrict__ lab, int* __restrict__ mapped) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;  // mapAdjacencyToBlockKernel (misHelpers.cu:221-250) with a permutation
  if (v >= n) return;
  int tb = newId[v], e1 = xadj[v + 1];
  for (int e = xadj[v]; e < e1; e++) { lab[e] = tb; mapped[e] = newId[adj[e]]; }
}
__global__ void agg_start_indices(int n, const int* __restrict__ fineSort, int* __restrict__ startIdx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;  // getAggregateStartIndicesKernel (misHelpers.cu:199-209)
  if (i < n && (i == 0 || fineSort[i] != fineSort[i - 1])) startIdx[fineSort[i]] = i;
}

}  // namespace

int metis_aggregate_device_graph(const Ctx& c, int n, const int* xadj_d, const int* adj_d, int partSize, IBuf& label_d);  // metis_agg.cpp

// CP::OldMIS (ComputePermutationMethods.cu:22-150, agg_type 0) and CP::MetisBottomUp (:151-266, agg_type 1): the two
// pipelines differ only in how the fine labels (aggregates) and the coarse labels (partitions) are obtained.
// agg_type 2, CP::MetisTopDown (:267-351, dispatched at randMIS.cu:474-489), is — despite its name — the OldMIS
// pipeline statement for statement (no METIS call; only the timers and verbose prints are missing): same code path.
void compute_permutation(const Ctx& c, int n, const int* xadj, const int* adj, int agg_type, int parameters, int partMaxSize, unsigned seed,
                         Aggregation& out) {
  cudaStream_t s = c.stream;
  const int fineDepth = parameters % 100, coarseDepth = (parameters / 100) % 100, minAgg = (parameters / 10000) % 10;
  int nedges;
  FSB_CUDA(cudaMemcpyAsync(&nedges, xadj + n, sizeof(int), cudaMemcpyDeviceToHost, s));
  FSB_CUDA(cudaStreamSynchronize(s));

  AggTrace tr(s);
  for (auto& v : g_trace_count) v = 0;
  const int coarseSize = partMaxSize % 1000;  // MetisBottomUp :180-182
  int fineSize = (partMaxSize / 1000) % 1000;
  fineSize = fineSize <= 0 ? 1 : fineSize;
  IBuf fineAggregate;
  const bool mis_pipeline = agg_type == 0 || agg_type == 2;
  int nAgg = mis_pipeline ? aggregate_graph(c, n, xadj, adj, minAgg, fineDepth, seed, fineAggregate)
                           : metis_aggregate_device_graph(c, n, xadj, adj, fineSize, fineAggregate);

  tr.lap("fine aggregates", n);
  tr.lap("  MIS rounds", g_trace_count[0]); tr.lap("  allocate sweeps", g_trace_count[1]);
  for (auto& v : g_trace_count) v = 0;
  // rows ordered by (aggregate, vertex): stable sort of the vertex ids by aggregate id (:72-75)
  IBuf perm(n, s), fineSort(n, s), iotaN(n, s);
  iota_i32(iotaN, n, s);
  sort_pairs_i32_i32(fineAggregate, fineSort, iotaN, perm, n, bits_for(nAgg), s);
  // aggregate sizes = node weights of the induced graph (:85)
  IBuf aggIdx, weights(nAgg, s);
  part_indices(c, fineSort, n, nAgg - 1, aggIdx);
  diff_i32<<<cdiv(nAgg, 256), 256, 0, s>>>(nAgg, aggIdx, weights);

  induced_graph(c, n, xadj, adj, nedges, fineAggregate, nAgg, out.xadjOut, out.adjOut);
  if ((int)out.xadjOut.size() != nAgg + 1) throw std::invalid_argument("induced graph: an aggregate has no external edge");
  int nInducedEdges = (int)out.adjOut.size();
  tr.lap("sort + induced graph", nInducedEdges);

  IBuf coarse;
  int nParts = mis_pipeline ? aggregate_weighted_graph(c, nAgg, out.xadjOut, out.adjOut, weights, partMaxSize, n, coarseDepth, seed, coarse)
                             : metis_aggregate_device_graph(c, nAgg, out.xadjOut, out.adjOut, std::max(coarseSize, 1), coarse);

  tr.lap("coarse partitions", nAgg);
  tr.lap("  MIS rounds", g_trace_count[0]); tr.lap("  allocate sweeps", g_trace_count[1]); tr.lap("  restrict iterations", g_trace_count[2]);
  // remapInducedGraph (misHelpers.cu:1258-1280): aggregates renumbered by (partition, old id)
  {
    IBuf iotaA(nAgg, s), cperm(nAgg, s), csorted(nAgg, s), ciperm(nAgg, s);
    iota_i32(iotaA, nAgg, s);
    sort_pairs_i32_i32(coarse, csorted, iotaA, cperm, nAgg, bits_for(nParts), s);
    inverse_perm<<<cdiv(nAgg, 256), 256, 0, s>>>(nAgg, cperm, ciperm);
    IBuf lab(nInducedEdges, s), mapped(nInducedEdges, s), slab(nInducedEdges, s), smapped(nInducedEdges, s);
    map_adjacency<<<cdiv(nAgg, 256), 256, 0, s>>>(nAgg, out.xadjOut, out.adjOut, ciperm, lab, mapped);
    sort_pairs_i32_i32(lab, slab, mapped, smapped, nInducedEdges, bits_for(nAgg), s);
    out.adjOut.swap(smapped);
    part_indices(c, slab, nInducedEdges, nAgg - 1, out.xadjOut);
  }

  // partition label of every row (:119), then rows ordered by (partition, aggregate, vertex) (:124)
  IBuf plabel(n, s), plabelSorted(n, s), order(n, s), fineSort2(n, s), perm2(n, s);
  gather_i32<<<cdiv(n, 256), 256, 0, s>>>(n, fineSort, coarse, plabel);
  sort_pairs_i32_i32(plabel, plabelSorted, iotaN, order, n, bits_for(nParts), s);
  gather_i32<<<cdiv(n, 256), 256, 0, s>>>(n, order, fineSort, fineSort2);
  gather_i32<<<cdiv(n, 256), 256, 0, s>>>(n, order, perm, perm2);

  // aggregates renumbered by first position in that order (:130-138)
  {
    IBuf iotaA(nAgg, s), startIdx(nAgg, s), startSorted(nAgg, s), remapId(nAgg, s), iRemap(nAgg, s);
    iota_i32(iotaA, nAgg, s);
    startIdx.zero();
    agg_start_indices<<<cdiv(n, 256), 256, 0, s>>>(n, fineSort2, startIdx);
    sort_pairs_i32_i32(startIdx, startSorted, iotaA, remapId, nAgg, bits_for(n), s);
    inverse_perm<<<cdiv(nAgg, 256), 256, 0, s>>>(nAgg, remapId, iRemap);
    gather_i32<<<cdiv(n, 256), 256, 0, s>>>(n, fineSort2, iRemap, fineSort);  // fineSort := remapped ids per new row
  }

  // partitionIdx: aggregate offsets of partitions (:141-142); aggregateIdx: row offsets of aggregates (:145)
  {
    IBuf csorted(nAgg, s), dummyIn(nAgg, s), dummyOut(nAgg, s);
    iota_i32(dummyIn, nAgg, s);
    sort_pairs_i32_i32(coarse, csorted, dummyIn, dummyOut, nAgg, bits_for(nParts), s);
    part_indices(c, csorted, nAgg, nParts - 1, out.partitionIdx);
  }
  part_indices(c, fineSort, n, nAgg - 1, out.aggregateIdx);

  out.ipermutation.swap(perm2);  // new -> old
  out.permutation.alloc(n, s);
  inverse_perm<<<cdiv(n, 256), 256, 0, s>>>(n, out.ipermutation, out.permutation);
  out.partitionLabel.swap(plabelSorted);
  out.n = n; out.nAgg = nAgg; out.nParts = nParts;
  tr.lap("remap + final orders", nParts);
  FSB_CHECK_LAUNCH();
  FSB_CUDA(cudaStreamSynchronize(s));
}

}  // namespace fsb
