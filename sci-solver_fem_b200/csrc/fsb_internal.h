// fsb_internal.h — internal C++ interfaces between the stages of the path.
#pragma once
#include "common.cuh"

namespace fsb {

struct Mesh {
  int nv = 0, ne = 0, npe = 0;  // npe = 4 (tets) or 3 (tris)
  DBuf xyz;                     // nv*3, AoS: one vertex = 24 contiguous bytes per gather
  IBuf elems;                   // ne*npe, AoS (tets: one int4 per element)
  IBuf labels;                  // ne material labels (tets) or empty
};

// Stage 1 — pattern.cu / assembly.cu
struct Pattern {
  IBuf ptr, col;                // canonical CSR: ascending columns incl. the diagonal
  int n = 0, nnz = 0;
  DevBuf<uint32_t> contrib;     // element-contribution ids sorted by (row, col, element): the gather lists
  DevBuf<long long> seg;        // nnz+1 offsets into contrib
  size_t ncontrib = 0;
};
void build_pattern(const Ctx& c, const Mesh& m, Pattern& p);
void graph_from_pattern(const Ctx& c, int n, const int* ptr, const int* col, IBuf& xadj, IBuf& adj);
void tet_mass_integrals_host(double out[10]);
void tri_quadrature_host(double zx[6], double zy[6], double wx[6], double wy[6]);
void assemble_values(const Ctx& c, const Mesh& m, const Pattern& p, double* val);

// Stage 2 — aggregation.cu
struct Aggregation {
  IBuf permutation, ipermutation, aggregateIdx, partitionIdx, partitionLabel, xadjOut, adjOut;
  int n = 0, nAgg = 0, nParts = 0;
};
void compute_permutation(const Ctx& c, int n, const int* xadj, const int* adj, int agg_type, int parameters, int partMaxSize,
                         unsigned seed, Aggregation& out);  // agg_type 0: OldMIS, 1: METIS bottom-up

// Stage 2 — hierarchy.cu
void permute_csr(const Ctx& c, const DCsr& A, const int* perm, DCsr& B);
void extract_diag(const Ctx& c, const DCsr& A, double* diag);
void build_prolongator(const Ctx& c, const DCsr& A, const double* diag, const int* aggregateIdx, int nAgg, double omega, DCsr& P);
void transpose_csr(const Ctx& c, const DCsr& A, DCsr& At);
void spgemm(const Ctx& c, const DCsr& A, const DCsr& B, DCsr& C);
void dense_inverse(const Ctx& c, const DCsr& A, DBuf& Ainv);  // n x n row-major, n < 1024
void build_sell(const Ctx& c, const DCsr& A, Sell& S, int sort_window = 0);  // sort_window > 0: SELL-32-sigma
struct LevelData;
void split_partitions(const Ctx& c, LevelData& L);  // A -> A_out (CSR) + intra-partition sliced ELL

}  // namespace fsb
