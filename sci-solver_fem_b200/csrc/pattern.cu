// pattern.cu — stage 1a: mesh -> CSR sparsity pattern + per-entry gather lists, on device.
//
// Replaces the reference's host path TetMesh::need_neighbors (src/core/cuda/tetmesh.cu:112-170)
// / TriMesh::need_neighbors (aggmis/cuda/TriMesh_connectivity.cu:96-131) + tetmesh2ell /
// trimesh2ell (src/core/cuda/cutil.cu:168-230, 297-348).  The result is the same matrix
// pattern — row i = {i} U N(i) — stored as CSR with ascending columns (what
// sort_by_row_and_column yields at smoothedMG_amg_level.cu:302), bit-exact.
//
// B200 design: instead of O(sum deg^2) std::find on one host thread, every element emits its
// npe*npe (row,col) pairs as packed 64-bit keys tagged with a contribution id
// (element*npe^2 + slot); ONE stable radix sort groups equal (row,col) and leaves the
// contributions of each matrix entry in element order.  Unique keys are the pattern; the
// sorted ids are the gather lists the assembly kernel sums in a fixed order
// (deterministic, atomic-free).
#include "fsb_internal.h"

namespace fsb {

static const uint32_t kNoContrib = 0xFFFFFFFFu;

// slot order of the reference scatter (perform_element_loop_3D.cuh:163-318): the 6 vertex
// pairs (0,1)(0,2)(0,3)(1,2)(1,3)(2,3), each in both directions, then the 4 diagonals.
__device__ __forceinline__ void slot_pair(int npe, int slot, int& a, int& b) {
  if (npe == 4) {
    const int pi[6] = {0, 0, 0, 1, 1, 2}, pj[6] = {1, 2, 3, 2, 3, 3};
    if (slot < 12) { int p = slot >> 1; a = (slot & 1) ? pj[p] : pi[p]; b = (slot & 1) ? pi[p] : pj[p]; }
    else { a = b = slot - 12; }
  } else {
    const int pi[3] = {0, 0, 1}, pj[3] = {1, 2, 2};
    if (slot < 6) { int p = slot >> 1; a = (slot & 1) ? pj[p] : pi[p]; b = (slot & 1) ? pi[p] : pj[p]; }
    else { a = b = slot - 6; }
  }
}

__global__ void emit_keys_kernel(int nv, long long ne, int npe, int cb, const int* __restrict__ elems,
                                 uint64_t* __restrict__ keys, uint32_t* __restrict__ cids) {
  const int ns = npe * npe;
  long long total = (long long)nv + ne * ns;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    if (k < nv) {  // every row owns its diagonal even if the vertex is in no element (cutil.cu:199)
      keys[k] = ((uint64_t)k << cb) | (uint64_t)k;
      cids[k] = kNoContrib;
      continue;
    }
    long long c = k - nv;
    long long e = c / ns;
    int slot = (int)(c - e * ns), a, b;
    slot_pair(npe, slot, a, b);
    int I = elems[e * npe + a], J = elems[e * npe + b];
    uint32_t cid = (uint32_t)c;
    // off-diagonal slot of a degenerate element (I == J): the reference's search over ELL
    // slots 1.. never finds the diagonal (perform_element_loop_3D.cuh:183), so it is dropped.
    if (a != b && I == J) cid = kNoContrib;
    keys[k] = ((uint64_t)I << cb) | (uint64_t)J;
    cids[k] = cid;
  }
}

__global__ void flag_unique_kernel(long long n, const uint64_t* __restrict__ keys, int* __restrict__ flag) {
  long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k < n) flag[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1 : 0;
}

__global__ void write_pattern_kernel(long long n, int cb, const uint64_t* __restrict__ keys, const int* __restrict__ pos,
                                     int* __restrict__ ptr, int* __restrict__ col, long long* __restrict__ seg) {
  long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint64_t key = keys[k];
  bool first = (k == 0) || (key != keys[k - 1]);
  if (!first) return;
  int slot = pos[k] - 1;
  uint64_t mask = (cb >= 64) ? ~0ull : ((1ull << cb) - 1ull);
  int row = (int)(key >> cb);
  col[slot] = (int)(key & mask);
  seg[slot] = k;
  if (k == 0 || (int)(keys[k - 1] >> cb) != row) ptr[row] = slot;
}

void build_pattern(const Ctx& c, const Mesh& m, Pattern& p) {
  cudaStream_t s = c.stream;
  const int ns = m.npe * m.npe;
  const long long total = (long long)m.nv + (long long)m.ne * ns;
  if ((long long)m.ne * ns >= 0xFFFFFFFFll) throw std::runtime_error("mesh too large for 32-bit contribution ids");
  const int cb = bits_for(m.nv > 0 ? m.nv - 1 : 0);
  DevBuf<uint64_t> k0(total, s), k1(total, s);
  DevBuf<uint32_t> c0(total, s);
  p.contrib.alloc(total, s);
  emit_keys_kernel<<<std::min<long long>(cdiv(total, 256), 148 * 64), 256, 0, s>>>(m.nv, m.ne, m.npe, cb, m.elems, k0, c0);
  FSB_CHECK_LAUNCH();
  sort_pairs_u64_u32(k0, k1, c0, p.contrib, total, 2 * cb, s);
  k0.release(); c0.release();
  IBuf flag(total, s), pos(total, s);
  flag_unique_kernel<<<cdiv(total, 256), 256, 0, s>>>(total, k1, flag);
  FSB_CHECK_LAUNCH();
  inclusive_scan_i32(flag, pos, total, s);
  flag.release();
  p.nnz = pos.read(total - 1);
  p.n = m.nv;
  p.ncontrib = total;
  p.ptr.alloc(m.nv + 1, s);
  p.col.alloc(p.nnz, s);
  p.seg.alloc((size_t)p.nnz + 1, s);
  write_pattern_kernel<<<cdiv(total, 256), 256, 0, s>>>(total, cb, k1, pos, p.ptr, p.col, p.seg);
  FSB_CHECK_LAUNCH();
  FSB_CUDA(cudaMemcpyAsync(p.ptr.get() + m.nv, &p.nnz, sizeof(int), cudaMemcpyHostToDevice, s));
  FSB_CUDA(cudaMemcpyAsync(p.seg.get() + p.nnz, &total, sizeof(long long), cudaMemcpyHostToDevice, s));
  FSB_CUDA(cudaStreamSynchronize(s));  // p.nnz / total are stack variables
}

// misHelpers::getAdjacency (misHelpers.cu:443-511): the level-0 graph is the (sorted)
// neighbour lists = the pattern without its diagonal.
__global__ void graph_ptr_kernel(int n, const int* __restrict__ ptr, int* __restrict__ xadj) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n) xadj[i] = ptr[i] - i;
}
__global__ void graph_adj_kernel(int n, const int* __restrict__ ptr, const int* __restrict__ col, int* __restrict__ adj) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int o = ptr[i] - i;
  for (int e = ptr[i]; e < ptr[i + 1]; e++) { int cc = col[e]; if (cc != i) adj[o++] = cc; }
}

void graph_from_pattern(const Ctx& c, int n, const int* ptr, const int* col, IBuf& xadj, IBuf& adj) {
  cudaStream_t s = c.stream;
  int nnz;
  FSB_CUDA(cudaMemcpyAsync(&nnz, ptr + n, sizeof(int), cudaMemcpyDeviceToHost, s));
  FSB_CUDA(cudaStreamSynchronize(s));
  xadj.alloc(n + 1, s);
  adj.alloc(nnz - n, s);
  graph_ptr_kernel<<<cdiv(n + 1, 256), 256, 0, s>>>(n, ptr, xadj);
  graph_adj_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, ptr, col, adj);
  FSB_CHECK_LAUNCH();
}

}  // namespace fsb
