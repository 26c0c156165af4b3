// metis_agg.cpp — aggregatorType_ = 1 ("METIS bottom-up"): k-way partitions as aggregates.
//
// Reference: Help::GetMetisAggregation (src/core/cuda/ComputePermutationMethods.cu:989-1047), the
// recursive large-graph variant (:1048-1094, GetSubGraphs :1095-1165) and EnsureConnectedAndNonEmpty
// (:1166-1218).  As upstream, the graph is brought to the host, partitioned by METIS and the labels
// go back to the device — the partitioner is third-party host code on both sides, not a fallback of
// ours.  Upstream links METIS 4.0.3 (fetched at configure time); here the METIS 5 static library
// shipped with the CUDA toolkit is called with default options (deterministic, 64-bit idx_t).
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "fsb_internal.h"

extern "C" int METIS_PartGraphKway(int64_t* nvtxs, int64_t* ncon, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* vsize,
                                   int64_t* adjwgt, int64_t* nparts, float* tpwgts, float* ubvec, int64_t* options, int64_t* objval,
                                   int64_t* part);

namespace fsb {
namespace {

typedef std::vector<int> ivec;

// Parts that METIS left disconnected are split into their connected components (label = largest
// vertex id reachable inside the part, by repeated sweeps), then labels are compacted in order.
int split_disconnected(const ivec& xadj, const ivec& adj, ivec& label) {
  const int n = (int)label.size();
  ivec top(n);
  for (int i = 0; i < n; i++) top[i] = i;
  for (bool changed = true; changed;) {
    changed = false;
    for (int v = 0; v < n; v++) {
      int best = top[v];
      for (int e = xadj[v]; e < xadj[v + 1]; e++) {
        int u = adj[e];
        if (label[u] == label[v] && top[u] > best) best = top[u];
      }
      if (best > top[v]) { top[v] = best; changed = true; }
    }
  }
  ivec ids(top);
  std::sort(ids.begin(), ids.end());
  ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
  for (int v = 0; v < n; v++) label[v] = (int)(std::lower_bound(ids.begin(), ids.end(), top[v]) - ids.begin());
  return (int)ids.size();
}

int metis_labels(const ivec& xadj, const ivec& adj, ivec& label, int partSize);

// nparts >= 8192: four-way split first, then every quarter is aggregated on its own (:1048-1094)
int metis_labels_large(const ivec& xadj, const ivec& adj, ivec& label, int partSize) {
  const int n = (int)xadj.size() - 1;
  metis_labels(xadj, adj, label, n / 4);
  std::vector<ivec> members;
  ivec local(n);
  for (int v = 0; v < n; v++) {
    int q = label[v];
    if (q + 1 > (int)members.size()) members.resize(q + 1);
    local[v] = (int)members[q].size();
    members[q].push_back(v);
  }
  const ivec quarter = label;
  int offset = 0;
  for (size_t q = 0; q < members.size(); q++) {
    const ivec& mem = members[q];
    ivec sx(mem.size() + 1, 0), sa, sub;
    for (size_t k = 0; k < mem.size(); k++) {
      for (int e = xadj[mem[k]]; e < xadj[mem[k] + 1]; e++)
        if (quarter[adj[e]] == (int)q) sa.push_back(local[adj[e]]);
      sx[k + 1] = (int)sa.size();
    }
    int cnt = metis_labels(sx, sa, sub, partSize);
    for (size_t k = 0; k < sub.size(); k++) label[mem[k]] = sub[k] + offset;
    offset += cnt;
  }
  return offset;
}

int metis_labels(const ivec& xadj, const ivec& adj, ivec& label, int partSize) {
  const int n = (int)xadj.size() - 1;
  label.assign(n, 0);
  int nparts = n / partSize;
  if (nparts >= 8192) return metis_labels_large(xadj, adj, label, partSize);
  nparts = std::max(nparts, 2);
  if (n > 0) {
    std::vector<int64_t> xa(xadj.begin(), xadj.end()), ad(adj.begin(), adj.end()), part(n, 0);
    int64_t nv = n, ncon = 1, np = nparts, cut = 0;
    if (METIS_PartGraphKway(&nv, &ncon, xa.data(), ad.data(), nullptr, nullptr, nullptr, &np, nullptr, nullptr, nullptr, &cut, part.data()) != 1)
      throw std::runtime_error("METIS_PartGraphKway failed");
    for (int v = 0; v < n; v++) label[v] = (int)part[v];
  }
  return split_disconnected(xadj, adj, label);
}

}  // namespace

// device graph -> host METIS -> device labels; returns the number of aggregates
int metis_aggregate_device_graph(const Ctx& c, int n, const int* xadj_d, const int* adj_d, int partSize, IBuf& label_d) {
  cudaStream_t s = c.stream;
  ivec xadj(n + 1);
  FSB_CUDA(cudaMemcpyAsync(xadj.data(), xadj_d, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost, s));
  FSB_CUDA(cudaStreamSynchronize(s));
  ivec adj(xadj[n]), label;
  if (xadj[n]) FSB_CUDA(cudaMemcpyAsync(adj.data(), adj_d, sizeof(int) * xadj[n], cudaMemcpyDeviceToHost, s));
  FSB_CUDA(cudaStreamSynchronize(s));
  int count = metis_labels(xadj, adj, label, partSize);
  label_d.alloc(n, s);
  label_d.from_host(label.data(), n);
  FSB_CUDA(cudaStreamSynchronize(s));
  return count;
}

}  // namespace fsb
