// =============================================================================
// fem_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT A PRODUCT PATH)
//
// A CPU restatement of SCIInstitute/SCI-Solver_FEM's solve path
// (mesh -> sparsity pattern -> P1 assembly -> smoothed-aggregation AMG setup ->
// AMG V-cycle / AMG-preconditioned CG).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.
// The CUDA product (sci-solver_fem_b200/csrc) never links, imports or calls it.
//
// Parity status: the upstream CUDA build cannot be compiled here (CUSP and
// METIS 4.0.3 are fetched at configure time and are not vendored; see
// DESIGN.md), so this restatement is pinned against the reference's own test
// fixtures (tetVol / simple / simpleTri .mat files + meshes: pattern identity,
// known-answer solve) and against SciPy as independent ground truth.  Stages
// for which the reference holds no golden vector (assembled values,
// aggregates, P, RAP, smoother, PCG iteration counts) are "parity unpinned":
// they follow the cited reference lines statement by statement.
//
// Every function cites the reference file:line (relative to
// /root/reference/src/core unless it starts with FEMSolver).
// =============================================================================
#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

typedef std::vector<int> ivec;

template <typename T>
struct Csr {
  int nrows = 0, ncols = 0;
  ivec ptr, col;
  std::vector<T> val;
  size_t nnz() const { return col.size(); }
};

// -----------------------------------------------------------------------------
// a2: vertex adjacency.  cuda/tetmesh.cu:112-170 (TetMesh::need_neighbors) and
// aggmis/cuda/TriMesh_connectivity.cu:96-131 (TriMesh::need_neighbors):
// for each element, for each corner j, push the other corners (j+1..)%npe if
// not already present (std::find), discovery order.
// a3: cuda/cutil.cu:168-230 / 297-348 then sorts each list ascending in place.
// -----------------------------------------------------------------------------
static void need_neighbors(int nv, int ne, int npe, const int* elems,
                           std::vector<ivec>& nb, bool sort_lists) {
  nb.assign(nv, ivec());
  ivec cnt(nv, 0);
  for (int e = 0; e < ne; e++)
    for (int j = 0; j < npe; j++) cnt[elems[e * npe + j]]++;
  for (int i = 0; i < nv; i++) nb[i].reserve(cnt[i] + 2);
  for (int e = 0; e < ne; e++)
    for (int j = 0; j < npe; j++) {
      ivec& me = nb[elems[e * npe + j]];
      for (int d = 1; d < npe; d++) {
        int n = elems[e * npe + (j + d) % npe];
        if (std::find(me.begin(), me.end(), n) == me.end()) me.push_back(n);
      }
    }
  if (sort_lists)
    for (int i = 0; i < nv; i++) std::sort(nb[i].begin(), nb[i].end());
}

// Canonical CSR of the ELL pattern built at cutil.cu:197-226 (slot 0 = diagonal,
// then ascending neighbours): here stored with ascending columns INCLUDING the
// diagonal, which is what sort_by_row_and_column yields at
// smoothedMG_amg_level.cu:302.  Also returns the graph (pattern minus diagonal)
// that misHelpers::getAdjacency (misHelpers.cu:443-511) hands to the aggregator.
static void pattern_from_mesh(int nv, int ne, int npe, const int* elems,
                              ivec& ptr, ivec& col, ivec& xadj, ivec& adj) {
  std::vector<ivec> nb;
  need_neighbors(nv, ne, npe, elems, nb, true);
  ptr.assign(nv + 1, 0);
  xadj.assign(nv + 1, 0);
  for (int i = 0; i < nv; i++) {
    ptr[i + 1] = ptr[i] + (int)nb[i].size() + 1;
    xadj[i + 1] = xadj[i] + (int)nb[i].size();
  }
  col.resize(ptr[nv]);
  adj.resize(xadj[nv]);
  for (int i = 0; i < nv; i++) {
    int p = ptr[i];
    bool placed = false;
    for (size_t k = 0; k < nb[i].size(); k++) {
      if (!placed && nb[i][k] > i) { col[p++] = i; placed = true; }
      col[p++] = nb[i][k];
      adj[xadj[i] + k] = nb[i][k];
    }
    if (!placed) col[p++] = i;
  }
}

// -----------------------------------------------------------------------------
// a4: quadrature tables.  cuda/FEM3D.cu:58-136 (compute_gamma_3d), :138-187
// (JacobiPoly), :189-214 (JacobiPolyDerivative), :216-274 (JacobiGZeros, Newton
// with deflation, EPS 1e-6, PI = 3.1415927 from include/TriMesh.h:27),
// :276-322 (JacobiGLZW), :324-368 (JacobiGRZW).  FEM2D.cu holds identical twins.
// -----------------------------------------------------------------------------
static double gamma_ref(double x) {
  static const double g[] = {1.0, 0.5772156649015329, -0.6558780715202538, -0.420026350340952e-1,
    0.1665386113822915, -0.421977345555443e-1, -0.9621971527877e-2, 0.7218943246663e-2,
    -0.11651675918591e-2, -0.2152416741149e-3, 0.1280502823882e-3, -0.201348547807e-4,
    -0.12504934821e-5, 0.1133027232e-5, -0.2056338417e-6, 0.6116095e-8, 0.50020075e-8,
    -0.11812746e-8, 0.1043427e-9, 0.77823e-11, -0.36968e-11, 0.51e-12, -0.206e-13, -0.54e-14, 0.14e-14};
  double ga, gr, r = 1.0, z;
  if (x > 171.0) return 1e308;
  if (x == (int)x) {
    if (x > 0.0) { ga = 1.0; for (int i = 2; i < x; i++) ga *= i; }
    else ga = 1e308;
  } else {
    if (fabs(x) > 1.0) {
      z = fabs(x); int m = (int)z; r = 1.0;
      for (int k = 1; k <= m; k++) r *= (z - k);
      z -= m;
    } else z = x;
    gr = g[24];
    for (int k = 23; k >= 0; k--) gr = gr * z + g[k];
    ga = 1.0 / (gr * z);
    if (fabs(x) > 1.0) { ga *= r; if (x < 0.0) ga = -M_PI / (x * ga * sin(M_PI * x)); }
  }
  return ga;
}

typedef std::vector<double> dvec;

static void jacobi_poly(int degree, const dvec& x, int alpha, int beta, dvec& y) {
  size_t s = x.size();
  y.resize(s);
  if (degree == 0) { for (size_t i = 0; i < s; i++) y[i] = 1.0; }
  else if (degree == 1) { for (size_t i = 0; i < s; i++) y[i] = 0.5 * (alpha - beta + (alpha + beta + 2.0) * x[i]); }
  else {
    double degm1 = degree - 1.0;
    double tmp = 2.0 * degm1 + alpha + beta;
    double a1 = 2.0 * (degm1 + 1) * (degm1 + alpha + beta + 1) * tmp;
    double a2 = (tmp + 1) * (alpha * alpha - beta * beta);
    double a3 = tmp * (tmp + 1.0) * (tmp + 2.0);
    double a4 = 2.0 * (degm1 + alpha) * (degm1 + beta) * (tmp + 2.0);
    dvec p1, p2;
    jacobi_poly(degree - 1, x, alpha, beta, p1);
    jacobi_poly(degree - 2, x, alpha, beta, p2);
    for (size_t i = 0; i < s; i++) y[i] = ((a2 + a3 * x[i]) * p1[i] - a4 * p2[i]) / a1;
  }
}

static void jacobi_poly_deriv(int degree, const dvec& x, int alpha, int beta, dvec& y) {
  if (degree == 0) { y.assign(x.size(), 0.0); return; }
  dvec poly;
  jacobi_poly(degree - 1, x, alpha + 1, beta + 1, poly);
  y.resize(poly.size());
  for (size_t i = 0; i < poly.size(); i++) y[i] = 0.5 * (alpha + beta + degree + 1) * poly[i];
}

static void jacobi_gzeros(int degree, int alpha, int beta, dvec& z) {
  z.assign(degree, 0.0);
  if (degree == 0) return;
  const int maxit = 60;
  const double EPS = 1.0e-6;
  const double PI_REF = 3.1415927;
  double dth = PI_REF / (2.0 * degree);
  double rlast = 0.0;
  dvec r(1), poly(1), pder(1);
  for (int k = 0; k < degree; k++) {
    r[0] = -cos((2.0 * k + 1.0) * dth);
    if (k) r[0] = 0.5 * (r[0] + rlast);
    for (int j = 0; j < maxit; j++) {
      jacobi_poly(degree, r, alpha, beta, poly);
      jacobi_poly_deriv(degree, r, alpha, beta, pder);
      double sum = 0.0;
      for (int i = 0; i < k; i++) sum = sum + 1.0 / (r[0] - z[i]);
      double delr = -poly[0] / (pder[0] - sum * poly[0]);
      r[0] = r[0] + delr;
      if (fabs(delr) < EPS) break;
    }
    z[k] = r[0];
    rlast = r[0];
  }
}

static void jacobi_glzw(dvec& Z, dvec& w, int degree, int alpha, int beta) {
  Z.assign(degree, 0.0); w.assign(degree, 0.0);
  if (degree == 1) { Z[0] = 0.0; w[0] = 0.0; return; }
  int apb = alpha + beta;
  Z[0] = -1; Z[degree - 1] = 1;
  dvec t;
  jacobi_gzeros(degree - 2, alpha + 1, beta + 1, t);
  for (int i = 1; i < degree - 1; i++) Z[i] = t[i - 1];
  jacobi_poly(degree - 1, Z, alpha, beta, w);
  double tmp1 = pow(2.0, (double)(apb + 1));
  double tmp2 = gamma_ref(alpha + degree);
  double fac = tmp1 * tmp2 * gamma_ref(beta + degree);
  fac = fac / ((degree - 1) * gamma_ref(degree) * gamma_ref(alpha + beta + degree + 1));
  for (int j = 0; j < degree; j++) w[j] = fac / (w[j] * w[j]);
  w[0] = w[0] * (beta + 1);
  w[degree - 1] = w[degree - 1] * (alpha + 1);
}

static void jacobi_grzw(dvec& Z, dvec& w, int degree, int alpha, int beta) {
  Z.assign(degree, 0.0); w.assign(degree, 0.0);
  if (degree == 1) { Z[0] = 0.0; w[0] = 2.0; return; }
  int apb = alpha + beta;
  Z[0] = -1;
  dvec t;
  jacobi_gzeros(degree - 1, alpha, beta + 1, t);
  for (int i = 1; i < degree; i++) Z[i] = t[i - 1];
  jacobi_poly(degree - 1, Z, alpha, beta, w);
  double tmp = gamma_ref(alpha + degree);
  double fac = pow(2.0, (double)apb) * tmp * gamma_ref(beta + degree);
  fac = fac / (gamma_ref(degree) * (beta + degree) * gamma_ref(apb + degree + 1));
  for (int j = 0; j < degree; j++) w[j] = fac * (1 - Z[j]) / (w[j] * w[j]);
  w[0] = w[0] * (beta + 1);
}

// FEM3D.cu:501-557 (assemble), :370-393 (Transform2StdTetSpace), :395-421
// (EvalBasisTet), :452-475 (IntegrationInTet): the 10 reference mass integrals.
static void tet_mass_integrals(double out[10]) {
  const int D = 4;
  dvec zx, zy, zz, wx, wy, wz;
  jacobi_glzw(zx, wx, D, 0, 0);
  jacobi_grzw(zy, wy, D, 1, 0);
  jacobi_grzw(zz, wz, D, 2, 0);
  for (int i = 0; i < D; i++) { wy[i] /= 2; wz[i] /= 4; }
  double phi[4][4][4][4];  // [i][j][k][s]
  const double coef[4][4] = {{1, -1, -1, -1}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) for (int k = 0; k < D; k++) {
    double cx = zx[i], cy = zy[j], cz = zz[k];
    double X = (1 + cx) * 0.5 * (1 - cy) * 0.5 * (1 - cz) * 0.5;
    double Y = (1 + cy) * 0.5 * (1 - cz) * 0.5;
    double Z = (1 + cz) * 0.5;
    for (int s = 0; s < 4; s++) phi[i][j][k][s] = coef[s][0] + coef[s][1] * X + coef[s][2] * Y + coef[s][3] * Z;
  }
  int cnt = 0;
  for (int k = 0; k < 4; k++) for (int g = k; g < 4; g++) {
    double integral = 0;
    for (int p = 0; p < D; p++) {
      double ty = 0.0;
      for (int q = 0; q < D; q++) {
        double tz = 0.0;
        for (int r = 0; r < D; r++) tz += phi[p][q][r][k] * phi[p][q][r][g] * wz[r];
        ty += tz * wy[q];
      }
      integral += ty * wx[p];
    }
    out[cnt++] = integral;
  }
}

// unsigned binary search over a CSR row; the reference searches ELL slots
// 1..K-1 (perform_element_loop_3D.cuh:118-142); same hit/miss semantics.
static inline int find_slot(const ivec& ptr, const ivec& col, int row, int c) {
  int lo = ptr[row], hi = ptr[row + 1] - 1;
  while (hi >= lo) {
    int mid = lo + (hi - lo) / 2;
    if (col[mid] > c) hi = mid - 1;
    else if (col[mid] < c) lo = mid + 1;
    else return mid;
  }
  return -1;
}

// -----------------------------------------------------------------------------
// a5/a7: tet assembly.  perform_element_loop_3D.cuh:480-627 (device kernel; the
// material switch :585-610), host twin :629-756, compute_stiffness_matrix_3d
// :28-48, compute_massmatrix_vector_3d :74-116, sum_into_global :163-318.
// A += K + 1.0*M.  Elements are visited in index order (host twin order).
// `closed_form` replaces the quadrature integrals by the exact 2/15, 1/15.
// -----------------------------------------------------------------------------
static void assemble_tet(int nv, int ne, const int* tets, const double* vx, const double* vy, const double* vz,
                         const int* labels, const ivec& ptr, const ivec& col, double* val, bool closed_form) {
  double integrand[10];
  if (closed_form) {
    int c = 0;
    for (int k = 0; k < 4; k++) for (int g = k; g < 4; g++) integrand[c++] = (k == g) ? 2.0 / 15.0 : 1.0 / 15.0;
  } else tet_mass_integrals(integrand);
  std::fill(val, val + ptr[nv], 0.0);
  double co = 1.0;
  for (int e = 0; e < ne; e++) {
    int ids[4]; double x[4], y[4], z[4];
    for (int j = 0; j < 4; j++) { ids[j] = tets[4 * e + j]; x[j] = vx[ids[j]]; y[j] = vy[ids[j]]; z[j] = vz[ids[j]]; }
    double a1 = x[1] - x[3], a2 = y[1] - y[3], a3 = z[1] - z[3];
    double b1 = x[2] - x[3], b2 = y[2] - y[3], b3 = z[2] - z[3];
    double c1 = x[0] - x[3], c2 = y[0] - y[3], c3 = z[0] - z[3];
    double Tvol = fabs(c1 * (a2 * b3 - a3 * b2) + c2 * (a3 * b1 - a1 * b3) + c3 * (a1 * b2 - a2 * b1)) / 6.0;
    double a11 = x[0], a12 = y[0], a13 = z[0], a14 = 1.0, a21 = x[1], a22 = y[1], a23 = z[1], a24 = 1.0,
           a31 = x[2], a32 = y[2], a33 = z[2], a34 = 1.0, a41 = x[3], a42 = y[3], a43 = z[3], a44 = 1.0;
    double det = a11 * a22 * a33 * a44 + a11 * a23 * a34 * a42 + a11 * a24 * a32 * a43
      + a12 * a21 * a34 * a43 + a12 * a23 * a31 * a44 + a12 * a24 * a33 * a41
      + a13 * a21 * a32 * a44 + a13 * a22 * a34 * a41 + a13 * a24 * a31 * a42
      + a14 * a21 * a33 * a42 + a14 * a22 * a31 * a43 + a14 * a23 * a32 * a41
      - a11 * a22 * a34 * a43 - a11 * a23 * a32 * a44 - a11 * a24 * a33 * a42
      - a12 * a21 * a33 * a44 - a12 * a23 * a34 * a41 - a12 * a24 * a31 * a43
      - a13 * a21 * a34 * a42 - a13 * a22 * a31 * a44 - a13 * a24 * a32 * a41
      - a14 * a21 * a32 * a43 - a14 * a22 * a33 * a41 - a14 * a23 * a31 * a42;
    double b11 = a22 * a33 * a44 + a23 * a34 * a42 + a24 * a32 * a43 - a22 * a34 * a43 - a23 * a32 * a44 - a24 * a33 * a42;
    double b12 = a12 * a34 * a43 + a13 * a32 * a44 + a14 * a33 * a42 - a12 * a33 * a44 - a13 * a34 * a42 - a14 * a32 * a43;
    double b13 = a12 * a23 * a44 + a13 * a24 * a42 + a14 * a22 * a43 - a12 * a24 * a43 - a13 * a22 * a44 - a14 * a23 * a42;
    double b14 = a12 * a24 * a33 + a13 * a22 * a34 + a14 * a23 * a32 - a12 * a23 * a34 - a13 * a24 * a32 - a14 * a22 * a33;
    double b21 = a21 * a34 * a43 + a23 * a31 * a44 + a24 * a33 * a41 - a21 * a33 * a44 - a23 * a34 * a41 - a24 * a31 * a43;
    double b22 = a11 * a33 * a44 + a13 * a34 * a41 + a14 * a31 * a43 - a11 * a34 * a43 - a13 * a31 * a44 - a14 * a33 * a41;
    double b23 = a11 * a24 * a43 + a13 * a21 * a44 + a14 * a23 * a41 - a11 * a23 * a44 - a13 * a24 * a41 - a14 * a21 * a43;
    double b24 = a11 * a23 * a34 + a13 * a24 * a31 + a14 * a21 * a33 - a11 * a24 * a33 - a13 * a21 * a34 - a14 * a23 * a31;
    double b31 = a21 * a32 * a44 + a22 * a34 * a41 + a24 * a31 * a42 - a21 * a34 * a42 - a22 * a31 * a44 - a24 * a32 * a41;
    double b32 = a11 * a34 * a42 + a12 * a31 * a44 + a14 * a32 * a41 - a11 * a32 * a44 - a12 * a34 * a41 - a14 * a31 * a42;
    double b33 = a11 * a22 * a44 + a12 * a24 * a41 + a14 * a21 * a42 - a11 * a24 * a42 - a12 * a21 * a44 - a14 * a22 * a41;
    double b34 = a11 * a24 * a32 + a12 * a21 * a34 + a14 * a22 * a31 - a11 * a22 * a34 - a12 * a24 * a31 - a14 * a21 * a32;
    // b41..b44 (the constant terms) are computed upstream but never used by the stiffness.
    double coeffs[16];
    coeffs[0] = b11 / det; coeffs[1] = b21 / det; coeffs[2] = b31 / det; coeffs[3] = 0;
    coeffs[4] = b12 / det; coeffs[5] = b22 / det; coeffs[6] = b32 / det; coeffs[7] = 0;
    coeffs[8] = b13 / det; coeffs[9] = b23 / det; coeffs[10] = b33 / det; coeffs[11] = 0;
    coeffs[12] = b14 / det; coeffs[13] = b24 / det; coeffs[14] = b34 / det; coeffs[15] = 0;
    switch (labels ? labels[e] : 0) {
      case 0: co = 1.0; break; case 1: co = 1.0; break; case 2: co = 2; break; case 3: co = 3.0; break;
      case 4: co = 4.0; break; case 5: co = 5.0; break; case 6: co = 6.0; break;
      default: break;  // upstream leaves `co` at its previous value
    }
    double stiff[10], mass[10];
    int cnt = 0;
    for (int k = 0; k < 4; k++) for (int g = k; g < 4; g++)
      stiff[cnt++] = (coeffs[4 * k] * coeffs[4 * g] + coeffs[4 * k + 1] * coeffs[4 * g + 1] + coeffs[4 * k + 2] * coeffs[4 * g + 2]) * Tvol * co;
    double x1 = x[0], y1 = y[0], z1 = z[0], x2 = x[1], y2 = y[1], z2 = z[1], x3 = x[2], y3 = y[2], z3 = z[2], x4 = x[3], y4 = y[3], z4 = z[3];
    double dj = 0.125 * ((-x1 + x2) * (-y1 + y3) * (-z1 + z4) + (-y1 + y2) * (-z1 + z3) * (-x1 + x4) + (-z1 + z2) * (-x1 + x3) * (-y1 + y4)
      - (-x1 + x2) * (-z1 + z3) * (-y1 + y4) - (-z1 + z2) * (-y1 + y3) * (-x1 + x4) - (-y1 + y2) * (-x1 + x3) * (-z1 + z4));
    double jac = fabs(dj);
    for (int c = 0; c < 10; c++) mass[c] = integrand[c] * jac;
    const double lambda = 1.0;
    // pair order of sum_into_global_linear_system_cuda_3d: (0,1)->1 (0,2)->2 (0,3)->3 (1,2)->5 (1,3)->6 (2,3)->8
    static const int pi[6] = {0, 0, 0, 1, 1, 2}, pj[6] = {1, 2, 3, 2, 3, 3}, pk[6] = {1, 2, 3, 5, 6, 8};
    for (int p = 0; p < 6; p++) {
      double coef = stiff[pk[p]] + lambda * mass[pk[p]];
      int I = ids[pi[p]], J = ids[pj[p]];
      int loc = (I == J) ? -1 : find_slot(ptr, col, I, J);
      if (loc >= 0) val[loc] += coef;
      loc = (I == J) ? -1 : find_slot(ptr, col, J, I);
      if (loc >= 0) val[loc] += coef;
    }
    static const int dk[4] = {0, 4, 7, 9};
    for (int k = 0; k < 4; k++) {
      int loc = find_slot(ptr, col, ids[k], ids[k]);
      val[loc] += stiff[dk[k]] + lambda * mass[dk[k]];
    }
  }
}

// -----------------------------------------------------------------------------
// a6: triangle assembly.  perform_element_loop_2D.cuh:287-372 (kernel), :29-48
// (stiffness, fabs area), :71-147 (mass by 6x6 collapsed quadrature with the
// SIGNED jacobian /8), :190-285 (scatter).  FEM2D.cu:358-378 (weights: GLZW(6,0,0),
// GRZW(6,1,0), not rescaled).  forceFunction == 0 so b stays 0.
// -----------------------------------------------------------------------------
static void assemble_tri(int nv, int ne, const int* tris, const double* vx, const double* vy,
                         const ivec& ptr, const ivec& col, double* val, bool closed_form) {
  const int D = 6;
  dvec zx, zy, wx, wy;
  jacobi_glzw(zx, wx, D, 0, 0);
  jacobi_grzw(zy, wy, D, 1, 0);
  std::fill(val, val + ptr[nv], 0.0);
  for (int e = 0; e < ne; e++) {
    int ids[3]; double x[3], y[3];
    for (int j = 0; j < 3; j++) { ids[j] = tris[3 * e + j]; x[j] = vx[ids[j]]; y[j] = vy[ids[j]]; }
    double TArea = fabs(x[0] * y[2] - x[0] * y[1] + x[1] * y[0] - x[1] * y[2] + x[2] * y[1] - x[2] * y[0]) / 2.0;
    double a11 = x[0], a12 = y[0], a13 = 1.0, a21 = x[1], a22 = y[1], a23 = 1.0, a31 = x[2], a32 = y[2], a33 = 1.0;
    double det = a11 * a22 * a33 + a21 * a32 * a13 + a31 * a12 * a23 - a11 * a32 * a23 - a31 * a22 * a13 - a21 * a12 * a33;
    double b11 = a22 * a33 - a23 * a32, b12 = a13 * a32 - a12 * a33, b13 = a12 * a23 - a13 * a22;
    double b21 = a23 * a31 - a21 * a33, b22 = a11 * a33 - a13 * a31, b23 = a13 * a21 - a11 * a23;
    double b31 = a21 * a32 - a22 * a31, b32 = a12 * a31 - a11 * a32, b33 = a11 * a22 - a12 * a21;
    double cf[9] = {b11 / det, b21 / det, b31 / det, b12 / det, b22 / det, b32 / det, b13 / det, b23 / det, b33 / det};
    double stiff[6], mass[6];
    int cnt = 0;
    for (int k = 0; k < 3; k++) for (int g = k; g < 3; g++)
      stiff[cnt++] = (cf[3 * k] * cf[3 * g] + cf[3 * k + 1] * cf[3 * g + 1]) * TArea;
    double jac = (x[0] * y[1] - x[1] * y[0] - x[0] * y[2] + x[2] * y[0] + x[1] * y[2] - x[2] * y[1]) / 8;
    if (closed_form) {
      cnt = 0;  // exact P1 mass with the signed jacobian: jac*4 = signed area
      for (int k = 0; k < 3; k++) for (int g = k; g < 3; g++) mass[cnt++] = (k == g ? 1.0 / 6.0 : 1.0 / 12.0) * (jac * 4.0);
    } else {
      double qx[D][D], qy[D][D];
      for (int m = 0; m < D; m++) for (int j = 0; j < D; j++) {
        qx[m][j] = x[0] * (1 - zx[m]) * 0.5 * (1 - zy[j]) * 0.5 + x[1] * (1 + zx[m]) * 0.5 * (1 - zy[j]) * 0.5 + x[2] * (1 + zy[j]) * 0.5;
        qy[m][j] = y[0] * (1 - zx[m]) * 0.5 * (1 - zy[j]) * 0.5 + y[1] * (1 + zx[m]) * 0.5 * (1 - zy[j]) * 0.5 + y[2] * (1 + zy[j]) * 0.5;
      }
      cnt = 0;
      for (int k = 0; k < 3; k++) for (int g = k; g < 3; g++) {
        double A1 = cf[3 * k], B1 = cf[3 * k + 1], C1 = cf[3 * k + 2], A2 = cf[3 * g], B2 = cf[3 * g + 1], C2 = cf[3 * g + 2];
        double integral = 0;
        for (int p = 0; p < D; p++) {
          double ty = 0.0;
          for (int q = 0; q < D; q++) ty += ((A1 * qx[p][q] + B1 * qy[p][q] + C1) * (A2 * qx[p][q] + B2 * qy[p][q] + C2) * jac) * wy[q];
          integral += ty * wx[p];
        }
        mass[cnt++] = integral;
      }
    }
    static const int pi[3] = {0, 0, 1}, pj[3] = {1, 2, 2}, pk[3] = {1, 2, 4};
    for (int p = 0; p < 3; p++) {
      double coef = stiff[pk[p]] + 1.0 * mass[pk[p]];
      int I = ids[pi[p]], J = ids[pj[p]];
      int loc = (I == J) ? -1 : find_slot(ptr, col, I, J);
      if (loc >= 0) val[loc] += coef;
      loc = (I == J) ? -1 : find_slot(ptr, col, J, I);
      if (loc >= 0) val[loc] += coef;
    }
    static const int dk[3] = {0, 3, 5};
    for (int k = 0; k < 3; k++) val[find_slot(ptr, col, ids[k], ids[k])] += stiff[dk[k]] + 1.0 * mass[dk[k]];
  }
}

// =============================================================================
// Aggregation (aggregatorType_ = 0, "OldMIS")
// =============================================================================

// a11: cuda/randomizedMIS_GPU.cu:3-272.  Tausworthe step :15-16 / :41-42.
static inline unsigned taus(unsigned z) { unsigned b = (((z << 13) ^ z) >> 19); return (((z & UINT_MAX) << 12) ^ b); }

static void mis_round(int n, int k, const ivec& xadj, const ivec& adj, ivec& best, ivec& origin, ivec& mis, int& incomplete) {
  ivec best2(n), origin2(n);
  for (int it = 0; it < k; it++) {
#pragma omp parallel for schedule(static)
    for (int v = 0; v < n; v++) {  // Iterate_Kernel / Final_Iterate_Kernel :65-152
      int b = best[v], o = origin[v];
      if (b < 1000001) {
        for (int e = xadj[v]; e < xadj[v + 1]; e++) {
          int nb = adj[e], ch = best[nb], cho = origin[nb];
          if (ch > 0 && ch == b && cho > o) o = cho;
          if (ch > b) { b = ch; o = cho; }
        }
      }
      best2[v] = b; origin2[v] = o;
    }
    if (it == k - 1) {
      int inc = 0;
      for (int v = 0; v < n; v++) {
        int st = -1;
        if (origin2[v] == v) st = 1; else if (best2[v] == 1000001) st = 0;
        mis[v] = st;
        if (st == -1) inc = 1;
      }
      if (inc) incomplete = 1;
    }
    best.swap(best2); origin.swap(origin2);
  }
}

static void randomized_mis(const ivec& xadj, const ivec& adj, ivec& mis, int k, unsigned seed) {
  int n = (int)xadj.size() - 1;
  mis.assign(n, -1);
  std::vector<unsigned> randoms(n), seeds(32768);
  srand(seed);  // upstream: srand(time(NULL)) :195 — the seed is a parameter here
  for (int i = 0; i < 32768; i++) seeds[i] = (unsigned)rand();
  int iterations = (n + 32767) / 32768;
  for (int t = 0; t < 32768; t++) {  // Generate_Randoms_Kernel :3-20
    unsigned z = seeds[t]; int off = t;
    for (int i = 0; i < iterations; i++) if (off < n) { z = taus(z); randoms[off] = z; off += 32768; }
  }
  ivec best(n), origin(n);
  for (int v = 0; v < n; v++) { origin[v] = v; best[v] = (int)(randoms[v] % 1000000); }  // First_Initialize_Kernel :22-33
  int incomplete = 0;
  mis_round(n, k, xadj, adj, best, origin, mis, incomplete);
  while (incomplete == 1) {
    for (int v = 0; v < n; v++) {  // Initialize_Kernel :35-63
      unsigned z = taus(randoms[v]);
      origin[v] = v;
      int value = (mis[v] == 1) ? 1000001 : 0;
      best[v] = (mis[v] == -1) ? (int)(z % 1000000) : value;
      randoms[v] = z;
    }
    incomplete = 0;
    mis_round(n, k, xadj, adj, best, origin, mis, incomplete);
  }
}

// allocateNodesKernel, misHelpers.cu:13-75.  The 10-slot vote degenerates to
// "first labelled neighbour in adjacency order" (SURVEY Appendix C); restated
// literally so the degenerate behaviour is reproduced rather than assumed.
static void allocate_nodes(int n, const ivec& xadj, const ivec& adj, const ivec& partIn, ivec& partOut, ivec& aggregated) {
#pragma omp parallel for schedule(static)
  for (int idx = 0; idx < n; idx++) {
    if (aggregated[idx] != 0) continue;
    int cand[10], cnt[10];
    for (int i = 0; i < 10; i++) { cand[i] = -1; cnt[i] = 0; }
    for (int e = xadj[idx]; e < xadj[idx + 1]; e++) {
      int c = partIn[adj[e]];
      if (c != -1) {
        for (int j = 0; j < 10 && c != -1; j++) {
          if (cand[j] == -1) { cand[j] = c; cnt[j] = 1; }
          else if (cand[j] == c) { cnt[j] += 1; c = -1; }
        }
      }
    }
    int addTo = cand[0], count = cnt[0];
    for (int i = 1; i < 10; i++) if (cnt[i] > count) { count = cnt[i]; addTo = cand[i]; }
    partOut[idx] = addTo;
    if (addTo != -1) aggregated[idx] = 1;
  }
}

// getPartSizes, misHelpers.cu:822-872 (+findPartIndicesKernel :77-89, getSizes).
static void part_sizes(const ivec& partition, ivec& sizes, ivec* indices) {
  ivec temp = partition;
  std::sort(temp.begin(), temp.end());
  int maxPart = temp.back();
  ivec pi(maxPart + 2, 0);
  int size = (int)temp.size();
  for (int idx = 0; idx < size; idx++) {
    int value = temp[idx], next = (idx != size - 1) ? temp[idx + 1] : -1;
    if (value != next) pi[value + 1] = idx + 1;
  }
  sizes.resize(maxPart + 1);
  for (int i = 0; i <= maxPart; i++) sizes[i] = pi[i + 1] - pi[i];
  if (indices) *indices = pi;
}

// getPartIndices, misHelpers.cu:874-893.
static void part_indices(const ivec& sorted, ivec& pi) {
  int maxPart = sorted.back();
  pi.assign(maxPart + 2, 0);
  int size = (int)sorted.size();
  for (int idx = 0; idx < size; idx++) {
    int value = sorted[idx], next = (idx != size - 1) ? sorted[idx + 1] : -1;
    if (value != next) pi[value + 1] = idx + 1;
  }
  pi[pi.size() - 1] = size;
}

// getWeightedPartSizes, misHelpers.cu:1109-1141.
static void weighted_part_sizes(const ivec& partition, const ivec& w, ivec& sizes) {
  int maxPart = *std::max_element(partition.begin(), partition.end());
  sizes.assign(maxPart + 1, 0);
  for (size_t i = 0; i < partition.size(); i++) sizes[partition[i]] += w[i];
}

// removeRuntyPartsKernel :252-263 applied with a stencil + inclusive scan.
static bool remove_parts(ivec& partition, const ivec& sizes, int threshold, bool return_on_none) {
  int np = (int)sizes.size();
  ivec stencil(np), sub(np);
  int removed = 0, run = 0;
  for (int i = 0; i < np; i++) { stencil[i] = sizes[i] < threshold ? 1 : 0; removed += stencil[i]; run += stencil[i]; sub[i] = run; }
  if (removed == 0 && return_on_none) return true;
  for (size_t i = 0; i < partition.size(); i++) {
    int c = partition[i];
    if (stencil[c] == 1) partition[i] = -1; else partition[i] -= sub[c];
  }
  return false;
}

// a12: aggregateGraph, misHelpers.cu:513-601 (+removeRuntyParts :744-781).
static void aggregate_graph(int minSize, int depth, const ivec& xadj, const ivec& adj, ivec& partIn, unsigned seed) {
  int n = (int)xadj.size() - 1;
  randomized_mis(xadj, adj, partIn, depth, seed);
  ivec aggregated = partIn;
  int run = 0;
  for (int i = 0; i < n; i++) { run += partIn[i]; partIn[i] = run; }                 // inclusive_scan
  for (int i = 0; i < n; i++) partIn[i] = (aggregated[i] == 0) ? -1 : partIn[i] - 1; // ifLabelOne
  ivec partOut = partIn;
  bool complete = false;
  while (!complete) {
    allocate_nodes(n, xadj, adj, partIn, partOut, aggregated);
    partIn = partOut;
    int unallocated = (int)std::count(aggregated.begin(), aggregated.end(), 0);
    if (unallocated == 0) {
      ivec sizes; part_sizes(partIn, sizes, nullptr);
      complete = remove_parts(partIn, sizes, minSize, true);
      if (!complete) {
        for (int i = 0; i < n; i++) aggregated[i] = (partIn[i] == -1) ? 0 : 1;  // findAggregated
        partOut = partIn;
      }
    }
  }
}

// findDesirabilityKernel, misHelpers.cu:297-378.  fp32 arithmetic exactly where
// written; the `x = a*b; x += c` pair is evaluated as one fused multiply-add
// (what nvcc's default -fmad=true emits for the upstream kernel); the CUDA
// product uses fmaf at the same place (DESIGN.md "float desirability").
static void find_desirability(int size, int optimalSize, const ivec& xadj, const ivec& adj, const ivec& partition,
                              const ivec& partSizes, const ivec& w, ivec& swap_to, ivec& swap_from, ivec& swap_index,
                              std::vector<float>& des) {
  for (int idx = 0; idx < size; idx++) {
    int currentPart = partition[idx], currentPartSize = partSizes[currentPart], nodeSize = w[idx];
    int selfAdjacency = 0, addTo = -1;
    float best = 0;
    float cwf = (float)abs(currentPartSize - optimalSize) / (float)optimalSize;
    float selfImp = (float)(abs(currentPartSize - optimalSize) - abs((currentPartSize - nodeSize) - optimalSize)) * cwf;
    if (selfImp > 0) {
      int cand[10], cnt[10];
      for (int i = 0; i < 10; i++) { cand[i] = -1; cnt[i] = 0; }
      for (int e = xadj[idx]; e < xadj[idx + 1]; e++) {
        int c = partition[adj[e]];
        if (c == currentPart) selfAdjacency++;
        else for (int j = 0; j < 10; j++) {
          if (c != -1 && cand[j] == -1) { cand[j] = c; cnt[j] = 1; c = -1; }
          else if (cand[j] == c) { cnt[j] += 1; c = -1; }
        }
      }
      for (int i = 1; i < 10; i++) {  // candidate 0 is skipped upstream (:351)
        if (cand[i] != -1) {
          int np = cand[i], nps = partSizes[np];
          float nwf = (float)abs(nps - optimalSize) / (float)optimalSize;
          float ni = fmaf((float)(abs(nps - optimalSize) - abs((nps + nodeSize) - optimalSize)), nwf, selfImp);
          ni = ni * ((float)cnt[i] / (float)selfAdjacency);
          if (ni > best) { addTo = np; best = ni; }
        }
      }
    }
    swap_from[idx] = currentPart; swap_index[idx] = idx; swap_to[idx] = addTo; des[idx] = best;
  }
}

// restrictPartitionSize, misHelpers.cu:679-727 (+makeSwapsKernel :380-412).
static void restrict_partition_size(int maxSize, int fullSize, const ivec& xadj, const ivec& adj, ivec& partition, const ivec& w) {
  int size = (int)partition.size();
  ivec partSizes, swap_to(size), swap_from(size), swap_index(size);
  std::vector<float> des(size);
  weighted_part_sizes(partition, w, partSizes);
  int averageSize = fullSize / (int)partSizes.size();
  int largest = 0;
  for (int s : partSizes) largest = std::max(largest, s);
  int guard = 0;
  while (largest > maxSize) {
    find_desirability(size, averageSize, xadj, adj, partition, partSizes, w, swap_to, swap_from, swap_index, des);
    // thrust::sort_by_key on tuple<int,float> keys = stable merge sort
    ivec order(size);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      if (swap_from[a] != swap_from[b]) return swap_from[a] < swap_from[b];
      return des[a] < des[b];
    });
    ivec sf(size), st(size), si(size); std::vector<float> sd(size);
    for (int i = 0; i < size; i++) { sf[i] = swap_from[order[i]]; st[i] = swap_to[order[i]]; si[i] = swap_index[order[i]]; sd[i] = des[order[i]]; }
    for (int idx = 0; idx < size; idx++) {
      bool last = (idx == size - 1) || (sf[idx] != sf[idx + 1]);
      if (last && sd[idx] > .1) {
        int nodeWeight = w[si[idx]];
        partition[si[idx]] = st[idx];
        partSizes[st[idx]] += nodeWeight;
        partSizes[sf[idx]] -= nodeWeight;
      }
    }
    largest = 0;
    for (int s : partSizes) largest = std::max(largest, s);
    if (++guard > 100000) throw std::runtime_error("restrictPartitionSize does not terminate");
  }
}

// a14: aggregateWeightedGraph, misHelpers.cu:603-677 (+removeRuntyPartitions :783-820).
static void aggregate_weighted_graph(int maxSize, int fullSize, int depth, const ivec& xadj, const ivec& adj, ivec& partIn,
                                     const ivec& w, unsigned seed) {
  int n = (int)xadj.size() - 1;
  randomized_mis(xadj, adj, partIn, depth, seed);
  ivec aggregated = partIn;
  int run = 0;
  for (int i = 0; i < n; i++) { run += partIn[i]; partIn[i] = run; }
  int misCount = partIn.back();
  for (int i = 0; i < n; i++) partIn[i] = (aggregated[i] == 0) ? -1 : partIn[i] - 1;
  ivec partOut = partIn;
  bool complete = false, firstTime = true;
  while (!complete) {
    allocate_nodes(n, xadj, adj, partIn, partOut, aggregated);
    partIn = partOut;
    int unallocated = (int)std::count(aggregated.begin(), aggregated.end(), 0);
    if (unallocated == 0) {
      if (!firstTime || misCount < 10) {
        restrict_partition_size(maxSize, fullSize, xadj, adj, partIn, w);
        complete = true;
      } else {
        firstTime = false;
        ivec sizes; weighted_part_sizes(partIn, w, sizes);
        double averageSize = (double)fullSize / sizes.size();
        int threshold = (int)(averageSize * .7);
        remove_parts(partIn, sizes, threshold, false);
        for (int i = 0; i < n; i++) aggregated[i] = (partIn[i] == -1) ? 0 : 1;
        partOut = partIn;
      }
    }
  }
}

// getInducedGraph, misHelpers.cu:1078-1107 (+mapAdjacencyToBlockKernel :221-250,
// findPartIndicesNegStartKernel :91-101).  The leading (-1,-1) pair is dropped
// unconditionally, exactly as upstream does.
static void induced_graph(const ivec& xadj, const ivec& adj, const ivec& label, ivec& xout, ivec& aout) {
  int n = (int)xadj.size() - 1;
  std::vector<std::pair<int, int>> pr(adj.size());
  for (int i = 0; i < n; i++) {
    int tb = label[i];
    for (int e = xadj[i]; e < xadj[i + 1]; e++) {
      int nb = label[adj[e]];
      pr[e] = (tb == nb) ? std::make_pair(-1, -1) : std::make_pair(tb, nb);
    }
  }
  std::sort(pr.begin(), pr.end());
  pr.erase(std::unique(pr.begin(), pr.end()), pr.end());
  int size = (int)pr.size();
  int maxPart = pr.back().first;
  xout.assign(maxPart + 2, 0);
  for (int idx = 1; idx < size; idx++) {
    int value = pr[idx].first;
    bool diff = (idx == size - 1) ? true : (pr[idx + 1].first != value);
    if (diff) xout[value + 1] = idx;
  }
  xout[xout.size() - 1] = size - 1;
  aout.resize(size - 1);
  for (int i = 1; i < size; i++) aout[i - 1] = pr[i].second;
}

// remapInducedGraph, misHelpers.cu:1258-1280.
static void remap_induced_graph(ivec& xadj, ivec& adj, const ivec& partition) {
  int n = (int)partition.size();
  ivec cperm(n), ciperm(n);
  std::iota(cperm.begin(), cperm.end(), 0);
  std::stable_sort(cperm.begin(), cperm.end(), [&](int a, int b) { return partition[a] < partition[b]; });
  for (int i = 0; i < n; i++) ciperm[cperm[i]] = i;
  size_t m = adj.size();
  ivec lab(m), mapped(m);
  for (int i = 0; i < n; i++) {
    int tb = ciperm[i];
    for (int e = xadj[i]; e < xadj[i + 1]; e++) {
      int nb = ciperm[adj[e]];
      if (tb == nb) { lab[e] = -1; mapped[e] = -1; } else { lab[e] = tb; mapped[e] = nb; }
    }
  }
  ivec ord(m);
  std::iota(ord.begin(), ord.end(), 0);
  std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return lab[a] < lab[b]; });
  ivec slab(m);
  for (size_t i = 0; i < m; i++) { slab[i] = lab[ord[i]]; adj[i] = mapped[ord[i]]; }
  part_indices(slab, xadj);
}

// -----------------------------------------------------------------------------
// a15: METIS aggregation (aggregatorType_ = 1).  Help::GetMetisAggregation
// ComputePermutationMethods.cu:989-1047, _Large :1048-1094, GetSubGraphs :1095-1165,
// EnsureConnectedAndNonEmpty :1166-1218.  Upstream links METIS 4.0.3 (not vendored);
// this container only has METIS 5 (libmetis_static.a of the CUDA toolkit, 64-bit
// idx_t), so "same partitioner" means: oracle and CUDA product call the identical
// METIS 5 entry point with default options (deterministic).  Parity unpinned vs METIS 4.
// -----------------------------------------------------------------------------
extern "C" int METIS_PartGraphKway(int64_t* nvtxs, int64_t* ncon, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* vsize,
                                   int64_t* adjwgt, int64_t* nparts, float* tpwgts, float* ubvec, int64_t* options, int64_t* objval,
                                   int64_t* part);

static int ensure_connected_and_nonempty(const ivec& indices, const ivec& adjacency, ivec& aggregation) {
  int n = (int)aggregation.size();
  ivec temp(n);
  for (int i = 0; i < n; i++) temp[i] = i;
  bool changed = true;
  while (changed) {
    changed = false;
    for (int root = 0; root < n; root++) {
      int rootValue = temp[root], rootAggregate = aggregation[root];
      for (int e = indices[root]; e < indices[root + 1]; e++) {
        int nb = adjacency[e];
        if (rootAggregate == aggregation[nb] && temp[nb] > rootValue) rootValue = temp[nb];
      }
      if (rootValue > temp[root]) { temp[root] = rootValue; changed = true; }
    }
  }
  ivec mapping = temp;
  std::sort(mapping.begin(), mapping.end());
  mapping.erase(std::unique(mapping.begin(), mapping.end()), mapping.end());
  for (int i = 0; i < n; i++) aggregation[i] = (int)(std::lower_bound(mapping.begin(), mapping.end(), temp[i]) - mapping.begin());
  return (int)mapping.size();
}

static int metis_aggregation(const ivec& indices, const ivec& adjacency, ivec& result, int partSize);

static int metis_aggregation_large(const ivec& indices, const ivec& adjacency, ivec& result, int partSize) {
  int graphSize = (int)indices.size() - 1;
  metis_aggregation(indices, adjacency, result, graphSize / 4);
  // GetSubGraphs
  std::vector<ivec> nodeMaps;
  ivec mapToSub(graphSize);
  for (int i = 0; i < graphSize; i++) {
    int pid = result[i];
    while (pid + 1 > (int)nodeMaps.size()) nodeMaps.emplace_back();
    nodeMaps[pid].push_back(i);
    mapToSub[i] = (int)nodeMaps[pid].size() - 1;
  }
  ivec partition = result;
  int offset = 0;
  for (size_t g = 0; g < nodeMaps.size(); g++) {
    const ivec& nodes = nodeMaps[g];
    ivec ind(nodes.size() + 1, 0), adj;
    for (size_t k = 0; k < nodes.size(); k++) {
      int node = nodes[k];
      for (int e = indices[node]; e < indices[node + 1]; e++)
        if (partition[adjacency[e]] == (int)g) adj.push_back(mapToSub[adjacency[e]]);
      ind[k + 1] = (int)adj.size();
    }
    ivec agg;
    int aggCount = metis_aggregation(ind, adj, agg, partSize);
    for (size_t k = 0; k < agg.size(); k++) result[nodes[k]] = agg[k] + offset;
    offset += aggCount;
  }
  return offset;
}

static int metis_aggregation(const ivec& indices, const ivec& adjacency, ivec& result, int partSize) {
  int graphSize = (int)indices.size() - 1;
  result.assign(graphSize, 0);
  int nparts = graphSize / partSize;
  if (nparts < 8192) {
    if (nparts < 2) nparts = 2;
    std::vector<int64_t> xa(indices.begin(), indices.end()), ad(adjacency.begin(), adjacency.end()), part(graphSize, 0);
    int64_t nv = graphSize, ncon = 1, np = nparts, objval = 0;
    if (graphSize > 0) {
      int rc = METIS_PartGraphKway(&nv, &ncon, xa.data(), ad.data(), nullptr, nullptr, nullptr, &np, nullptr, nullptr, nullptr, &objval, part.data());
      if (rc != 1) throw std::runtime_error("METIS_PartGraphKway failed");
    }
    for (int i = 0; i < graphSize; i++) result[i] = (int)part[i];
    return ensure_connected_and_nonempty(indices, adjacency, result);
  }
  return metis_aggregation_large(indices, adjacency, result, partSize);
}

struct AggOut {
  ivec permutation, ipermutation, aggregateIdx, partitionIdx, partitionLabel, xadjOut, adjOut;
  ivec fineAggregate;  // per OLD vertex: final (renumbered) aggregate id
};

// a13: CP::OldMIS, ComputePermutationMethods.cu:22-150 (agg_type 0) and a15: CP::MetisBottomUp
// :151-266 (agg_type 1).  The two differ only in how the fine and the coarse labels are obtained.
// agg_type 2, CP::MetisTopDown (:267-351), calls no METIS routine at all: statement for statement it is the
// OldMIS pipeline without the timers and verbose prints (randMIS.cu:474-489 dispatches to it), so it shares the code.
static void compute_permutation(const ivec& xadj, const ivec& adj, int agg_type, int parameters, int part_max_size, unsigned seed, AggOut& o) {
  int n = (int)xadj.size() - 1;
  int fineDepth = parameters % 100, coarseDepth = (parameters / 100) % 100, minAgg = (parameters / 10000) % 10;
  int coarseSize = part_max_size % 1000, fineSize = (part_max_size / 1000) % 1000;  // MetisBottomUp :180-182
  fineSize = fineSize <= 0 ? 1 : fineSize;
  ivec fineAggregate;
  if (agg_type == 0 || agg_type == 2) aggregate_graph(minAgg, fineDepth, xadj, adj, fineAggregate, seed);
  else metis_aggregation(xadj, adj, fineAggregate, fineSize);
  ivec perm(n);
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return fineAggregate[a] < fineAggregate[b]; });
  ivec fineSort(n);
  for (int i = 0; i < n; i++) fineSort[i] = fineAggregate[perm[i]];
  ivec weights, aggIdx;
  part_sizes(fineSort, weights, &aggIdx);
  induced_graph(xadj, adj, fineAggregate, o.xadjOut, o.adjOut);
  ivec coarse;
  if (agg_type == 0 || agg_type == 2) aggregate_weighted_graph(part_max_size, n, coarseDepth, o.xadjOut, o.adjOut, coarse, weights, seed);
  else metis_aggregation(o.xadjOut, o.adjOut, coarse, coarseSize);
  remap_induced_graph(o.xadjOut, o.adjOut, coarse);
  ivec plabel(n);
  for (int i = 0; i < n; i++) plabel[i] = coarse[fineSort[i]];  // fillPartitionLabelKernel :190-197
  ivec ord(n);
  std::iota(ord.begin(), ord.end(), 0);
  std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return plabel[a] < plabel[b]; });
  ivec pl2(n), fs2(n), pm2(n);
  for (int i = 0; i < n; i++) { pl2[i] = plabel[ord[i]]; fs2[i] = fineSort[ord[i]]; pm2[i] = perm[ord[i]]; }
  plabel.swap(pl2); fineSort.swap(fs2); perm.swap(pm2);
  int nAgg = (int)aggIdx.size() - 1;
  ivec remapId(nAgg), remapIndex(nAgg, 0), iRemap(nAgg);
  std::iota(remapId.begin(), remapId.end(), 0);
  for (int idx = 0; idx < n; idx++)  // getAggregateStartIndicesKernel :199-209
    if (idx == 0 || fineSort[idx] != fineSort[idx - 1]) remapIndex[fineSort[idx]] = idx;
  std::stable_sort(remapId.begin(), remapId.end(), [&](int a, int b) { return remapIndex[a] < remapIndex[b]; });
  for (int i = 0; i < nAgg; i++) iRemap[remapId[i]] = i;
  for (int i = 0; i < n; i++) { fineSort[i] = iRemap[fineSort[i]]; fineAggregate[i] = iRemap[fineAggregate[i]]; }
  std::sort(coarse.begin(), coarse.end());
  part_indices(coarse, o.partitionIdx);
  part_indices(fineSort, o.aggregateIdx);
  o.ipermutation = perm;
  o.permutation.assign(n, -1);
  for (int i = 0; i < n; i++) o.permutation[perm[i]] = i;
  o.partitionLabel = plabel;
  o.fineAggregate = fineAggregate;
}

// =============================================================================
// Level construction
// =============================================================================
template <typename T>
struct Level {
  int n = 0, nnout = 0, level_id = 0;
  Csr<T> A;          // permuted, sorted by row/col (smoothedMG_amg_level.cu:302-303); unpermuted on the coarsest level
  AggOut agg;
  ivec xadj, adj;    // graph handed to the aggregator
  std::vector<T> diag;
  Csr<T> P, R;
  int nparts = 0, largestblocksize = 0;
  ivec pstart;       // first row of each partition (+ end)
  ivec AinBlockIdx, AoutBlockIdx;   // histogram version of smoothedMG_amg_level.cu:246-274
  // work vectors
  std::vector<T> bc, xc;
};

// generateMatrixSymmetric_d, smoothedMG_amg_level.cu:199-304 (+matrixpermute_kernel
// :45-80).  The in/out counts are taken by histogram (SURVEY Appendix C hazard note):
// a deliberate, documented deviation from the compacted reduce_by_key indexing.
template <typename T>
static void permute_and_split(Level<T>& L) {
  const Csr<T>& A0 = L.A;
  int n = A0.nrows;
  const ivec& perm = L.agg.permutation;
  const ivec& plabel = L.agg.partitionLabel;
  Csr<T> B; B.nrows = B.ncols = n; B.ptr.assign(n + 1, 0);
  for (int i = 0; i < n; i++) B.ptr[perm[i] + 1] = A0.ptr[i + 1] - A0.ptr[i];
  for (int i = 0; i < n; i++) B.ptr[i + 1] += B.ptr[i];
  B.col.resize(A0.nnz()); B.val.resize(A0.nnz());
  for (int i = 0; i < n; i++) {
    int r = perm[i], base = B.ptr[r], len = A0.ptr[i + 1] - A0.ptr[i];
    std::vector<std::pair<int, T>> row(len);
    for (int k = 0; k < len; k++) row[k] = std::make_pair(perm[A0.col[A0.ptr[i] + k]], A0.val[A0.ptr[i] + k]);
    std::sort(row.begin(), row.end(), [](const std::pair<int, T>& a, const std::pair<int, T>& b) { return a.first < b.first; });
    for (int k = 0; k < len; k++) { B.col[base + k] = row[k].first; B.val[base + k] = row[k].second; }
  }
  L.A = B;
  L.nparts = (int)L.agg.partitionIdx.size() - 1;
  L.pstart.resize(L.nparts + 1);
  for (int p = 0; p <= L.nparts; p++) L.pstart[p] = L.agg.aggregateIdx[L.agg.partitionIdx[p]];
  L.largestblocksize = 0;
  for (int p = 0; p < L.nparts; p++) L.largestblocksize = std::max(L.largestblocksize, L.pstart[p + 1] - L.pstart[p]);
  L.AinBlockIdx.assign(L.nparts + 1, 0); L.AoutBlockIdx.assign(L.nparts + 1, 0);
  for (int r = 0; r < n; r++) {
    int p = plabel[r];
    for (int e = B.ptr[r]; e < B.ptr[r + 1]; e++) {
      int c = B.col[e];
      if (plabel[c] == p) { if (c > r) L.AinBlockIdx[p + 1]++; } else L.AoutBlockIdx[p + 1]++;
    }
  }
  for (int p = 0; p < L.nparts; p++) { L.AinBlockIdx[p + 1] += L.AinBlockIdx[p]; L.AoutBlockIdx[p + 1] += L.AoutBlockIdx[p]; }
  L.diag.resize(n);  // gauss_seidel ctor: extract_diagonal, gauss_seidel.cu:61-70
  for (int r = 0; r < n; r++) {
    L.diag[r] = 0;
    for (int e = B.ptr[r]; e < B.ptr[r + 1]; e++) if (B.col[e] == r) L.diag[r] = B.val[e];
  }
}

// generateProlongatorFull_d, smoothedMG_amg_level.cu:320-387:
// P = T - omega * D^-1 * A * T, T[i, agg(i)] = 1, duplicates summed after a
// stable (row, col) sort (A-entries in column order first, then the T entry).
template <typename T>
static void build_prolongator(Level<T>& L, double proOmega) {
  const Csr<T>& A = L.A;
  int n = A.nrows, nagg = L.nnout;
  ivec aggOf(n);
  for (int a = 0; a < nagg; a++) for (int r = L.agg.aggregateIdx[a]; r < L.agg.aggregateIdx[a + 1]; r++) aggOf[r] = a;
  const T lambda = (T)proOmega;
  Csr<T>& P = L.P; P.nrows = n; P.ncols = nagg; P.ptr.assign(n + 1, 0); P.col.clear(); P.val.clear();
  std::vector<std::pair<int, T>> tmp;
  for (int i = 0; i < n; i++) {
    tmp.clear();
    T d = L.diag[i];
    for (int e = A.ptr[i]; e < A.ptr[i + 1]; e++) tmp.push_back(std::make_pair(aggOf[A.col[e]], (T)((-lambda * A.val[e] * (T)1) / d)));
    tmp.push_back(std::make_pair(aggOf[i], (T)1));
    std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<int, T>& a, const std::pair<int, T>& b) { return a.first < b.first; });
    for (size_t k = 0; k < tmp.size();) {
      size_t k2 = k; T s = 0;
      while (k2 < tmp.size() && tmp[k2].first == tmp[k].first) { s += tmp[k2].second; k2++; }
      P.col.push_back(tmp[k].first); P.val.push_back(s);
      k = k2;
    }
    P.ptr[i + 1] = (int)P.col.size();
  }
  // R = P^T (cusp::transpose :382)
  Csr<T>& R = L.R; R.nrows = nagg; R.ncols = n; R.ptr.assign(nagg + 1, 0);
  for (size_t e = 0; e < P.col.size(); e++) R.ptr[P.col[e] + 1]++;
  for (int a = 0; a < nagg; a++) R.ptr[a + 1] += R.ptr[a];
  R.col.resize(P.nnz()); R.val.resize(P.nnz());
  ivec fill(R.ptr.begin(), R.ptr.end() - 1);
  for (int i = 0; i < n; i++) for (int e = P.ptr[i]; e < P.ptr[i + 1]; e++) { int q = fill[P.col[e]]++; R.col[q] = i; R.val[q] = P.val[e]; }
}

// C = A*B, sorted columns, all structural entries kept (cusp::multiply ESC,
// called at smoothedMG_amg_level.cu:394,397).
template <typename T>
static void spgemm(const Csr<T>& A, const Csr<T>& B, Csr<T>& C) {
  C.nrows = A.nrows; C.ncols = B.ncols; C.ptr.assign(A.nrows + 1, 0);
  std::vector<ivec> cols(A.nrows); std::vector<std::vector<T>> vals(A.nrows);
#pragma omp parallel
  {
    std::vector<T> acc(B.ncols, 0); std::vector<char> mark(B.ncols, 0); ivec list;
#pragma omp for schedule(dynamic, 256)
    for (int i = 0; i < A.nrows; i++) {
      list.clear();
      for (int e = A.ptr[i]; e < A.ptr[i + 1]; e++) {
        int k = A.col[e]; T a = A.val[e];
        for (int f = B.ptr[k]; f < B.ptr[k + 1]; f++) {
          int j = B.col[f];
          if (!mark[j]) { mark[j] = 1; list.push_back(j); acc[j] = 0; }
          acc[j] += a * B.val[f];
        }
      }
      std::sort(list.begin(), list.end());
      cols[i] = list; vals[i].resize(list.size());
      for (size_t q = 0; q < list.size(); q++) { vals[i][q] = acc[list[q]]; mark[list[q]] = 0; }
    }
  }
  for (int i = 0; i < A.nrows; i++) C.ptr[i + 1] = C.ptr[i] + (int)cols[i].size();
  C.col.resize(C.ptr[A.nrows]); C.val.resize(C.ptr[A.nrows]);
  for (int i = 0; i < A.nrows; i++) { std::copy(cols[i].begin(), cols[i].end(), C.col.begin() + C.ptr[i]); std::copy(vals[i].begin(), vals[i].end(), C.val.begin() + C.ptr[i]); }
}

// dense LU with partial pivoting (cusp::detail::lu_solver, called at amg.cu:103-105
// and amg_level.cu:27-30).  CUSP is not vendored; standard Doolittle with row pivoting.
template <typename T>
struct DenseLU {
  int n = 0; std::vector<T> lu; ivec piv;
  void factor(const Csr<T>& A) {
    n = A.nrows; lu.assign((size_t)n * n, 0); piv.resize(n);
    for (int i = 0; i < n; i++) for (int e = A.ptr[i]; e < A.ptr[i + 1]; e++) lu[(size_t)i * n + A.col[e]] += A.val[e];
    for (int k = 0; k < n; k++) {
      int p = k; T mx = std::fabs(lu[(size_t)k * n + k]);
      for (int i = k + 1; i < n; i++) if (std::fabs(lu[(size_t)i * n + k]) > mx) { mx = std::fabs(lu[(size_t)i * n + k]); p = i; }
      piv[k] = p;
      if (p != k) for (int j = 0; j < n; j++) std::swap(lu[(size_t)k * n + j], lu[(size_t)p * n + j]);
      T d = lu[(size_t)k * n + k];
      for (int i = k + 1; i < n; i++) {
        T f = lu[(size_t)i * n + k] / d; lu[(size_t)i * n + k] = f;
        for (int j = k + 1; j < n; j++) lu[(size_t)i * n + j] -= f * lu[(size_t)k * n + j];
      }
    }
  }
  void solve(const std::vector<T>& b, std::vector<T>& x) const {
    x = b;
    // factor() swaps whole rows (LAPACK style: the multipliers already stored travel with their row), so ALL row
    // interchanges are applied to the right-hand side first, then the unit-lower solve.  (Interleaving the swap of step k
    // with the elimination of step k is only equivalent when no pivoting happens — true for the structured cubes, not
    // for the repo's TetGen cube, where the mistake cost the oracle 7 PCG iterations; found by an independent NumPy
    // V-cycle, tests/test_oracle_golden.py::test_coarse_lu_with_row_interchanges.)
    for (int k = 0; k < n; k++) if (piv[k] != k) std::swap(x[k], x[piv[k]]);
    for (int k = 0; k < n; k++) for (int i = k + 1; i < n; i++) x[i] -= lu[(size_t)i * n + k] * x[k];
    for (int i = n - 1; i >= 0; i--) { T s = x[i]; for (int j = i + 1; j < n; j++) s -= lu[(size_t)i * n + j] * x[j]; x[i] = s / lu[(size_t)i * n + i]; }
  }
};

struct Params {
  int maxLevels = 100, maxIters = 100, preInner = 5, postInner = 5, postRelaxes = 1, topSize = 256,
      randMisParameters = 90102, partitionMaxSize = 512, aggregatorType = 0, solverType = 0;
  double tolerance = 1e-6, smootherWeight = 1.0, proOmega = 0.67;
  unsigned seed = 0;
  int ref_level0_noperm = 0;  // 1 = reproduce SURVEY F3 (no permutation of b/x at level 0)
};

template <typename T>
struct Hierarchy {
  std::vector<Level<T>> levels;
  DenseLU<T> LU;
  Params prm;
  double t_setup = 0;

  // AMG::setup, amg.cu:79-144 + createNextLevel, smoothedMG_amg_level.cu:402-494.
  void setup(const Csr<T>& Afine, const ivec& xadj0, const ivec& adj0) {
    levels.clear();
    levels.emplace_back();
    levels[0].A = Afine; levels[0].n = Afine.nrows; levels[0].level_id = 0;
    levels[0].xadj = xadj0; levels[0].adj = adj0;
    int num_levels = 1;
    while (true) {
      Level<T>& L = levels.back();
      int N = L.A.nrows;
      if (N < prm.topSize || num_levels >= prm.maxLevels) { LU.factor(L.A); break; }
      if (prm.aggregatorType < 0 || prm.aggregatorType > 2) throw std::runtime_error("oracle: only aggregatorType_ 0 (OldMIS), 1 (METIS bottom-up) and 2 (\"METIS top-down\" = the MIS pipeline) are restated");
      compute_permutation(L.xadj, L.adj, prm.aggregatorType, prm.randMisParameters, prm.partitionMaxSize, prm.seed, L.agg);
      L.nnout = (int)L.agg.aggregateIdx.size() - 1;
      permute_and_split(L);
      if (L.largestblocksize > 1024) throw std::runtime_error("largest block size is larger than shared size");  // gauss_seidel.cu:2059-2064
      build_prolongator(L, prm.proOmega);
      Csr<T> AP, Ac;
      spgemm(L.A, L.P, AP);
      spgemm(L.R, AP, Ac);
      L.bc.assign(L.nnout, (T)-1); L.xc.assign(L.nnout, (T)-1);
      Level<T> nx;
      nx.A = Ac; nx.n = L.nnout; nx.level_id = num_levels; nx.xadj = L.agg.xadjOut; nx.adj = L.agg.adjOut;
      levels.push_back(std::move(nx));
      num_levels++;
    }
  }

  // One partition-local Jacobi sweep: x += w (b - Ain_off x - d x)/d.
  // preRRSym_kernel1, gauss_seidel.cu:1312-1375; postRelaxSym_kernel1 :3664-3735.
  void inner_sweeps(const Level<T>& L, int p, const T* brow, T* x, int nit, std::vector<T>& tmp) const {
    int r0 = L.pstart[p], r1 = L.pstart[p + 1];
    const T w = (T)prm.smootherWeight;
    for (int it = 0; it < nit; it++) {
      for (int r = r0; r < r1; r++) {
        T s = 0;
        for (int e = L.A.ptr[r]; e < L.A.ptr[r + 1]; e++) { int c = L.A.col[e]; if (c >= r0 && c < r1 && c != r) s += L.A.val[e] * x[c]; }
        tmp[r - r0] = s;
      }
      for (int r = r0; r < r1; r++) x[r] += w * (brow[r - r0] - tmp[r - r0] - L.diag[r] * x[r]) / L.diag[r];
    }
  }

  // AMG_Level::cycle, amg_level.cu:22-71; preRRRFullSymmetric gauss_seidel.cu:2017-2261;
  // postPCRFullSymmetric :4408-4662.  b and x are in the level's EXTERNAL numbering.
  void cycle(int lev, std::vector<T>& b, std::vector<T>& x) {
    Level<T>& L = levels[lev];
    if (lev == (int)levels.size() - 1) { LU.solve(b, x); return; }
    int n = L.n;
    const T w = (T)prm.smootherWeight;
    bool permute = (lev != 0) || !prm.ref_level0_noperm;
    if (permute) { std::vector<T> bo(n); for (int i = 0; i < n; i++) bo[i] = b[L.agg.ipermutation[i]]; b.swap(bo); }  // permutation_kernel1 :136-142
    std::vector<T> res(n, 0);
    x.assign(n, 0);
#pragma omp parallel
    {
      std::vector<T> tmp(L.largestblocksize), bl(L.largestblocksize);
#pragma omp for schedule(dynamic, 4)
      for (int p = 0; p < L.nparts; p++) {
        int r0 = L.pstart[p], r1 = L.pstart[p + 1];
        for (int r = r0; r < r1; r++) { bl[r - r0] = b[r]; x[r] = w * b[r] / L.diag[r]; }
        inner_sweeps(L, p, bl.data(), x.data(), prm.preInner, tmp);
        for (int r = r0; r < r1; r++) {
          T s = 0;
          for (int e = L.A.ptr[r]; e < L.A.ptr[r + 1]; e++) { int c = L.A.col[e]; if (c >= r0 && c < r1 && c != r) s += L.A.val[e] * x[c]; }
          res[r] = b[r] - s - L.diag[r] * x[r];
        }
      }
      // preAout_kernel :1977-2015
#pragma omp for schedule(dynamic, 4)
      for (int p = 0; p < L.nparts; p++) {
        int r0 = L.pstart[p], r1 = L.pstart[p + 1];
        for (int r = r0; r < r1; r++) {
          T s = 0;
          for (int e = L.A.ptr[r]; e < L.A.ptr[r + 1]; e++) { int c = L.A.col[e]; if (c < r0 || c >= r1) s += L.A.val[e] * x[c]; }
          res[r] -= s;
        }
      }
      // bc = R * residual :2260
#pragma omp for schedule(static)
      for (int a = 0; a < L.nnout; a++) {
        T s = 0;
        for (int e = L.R.ptr[a]; e < L.R.ptr[a + 1]; e++) s += L.R.val[e] * res[L.R.col[e]];
        L.bc[a] = s;
      }
    }
    cycle(lev + 1, L.bc, L.xc);
    // x += P xc :4425-4427
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
      T s = 0;
      for (int e = L.P.ptr[i]; e < L.P.ptr[i + 1]; e++) s += L.P.val[e] * L.xc[L.P.col[e]];
      x[i] = x[i] + s;
    }
    std::vector<T> xout(n);
    for (int rel = 0; rel < prm.postRelaxes; rel++) {
#pragma omp parallel
      {
        std::vector<T> tmp(L.largestblocksize), bl(L.largestblocksize);
#pragma omp for schedule(dynamic, 4)
        for (int p = 0; p < L.nparts; p++) {
          int r0 = L.pstart[p], r1 = L.pstart[p + 1];
          for (int r = r0; r < r1; r++) {
            T s = b[r];
            for (int e = L.A.ptr[r]; e < L.A.ptr[r + 1]; e++) { int c = L.A.col[e]; if (c < r0 || c >= r1) s += -L.A.val[e] * x[c]; }
            bl[r - r0] = s; xout[r] = x[r];
          }
          inner_sweeps(L, p, bl.data(), xout.data(), prm.postInner, tmp);
        }
      }
      x.swap(xout);
    }
    if (permute) { for (int i = 0; i < n; i++) xout[L.agg.ipermutation[i]] = x[i]; x.swap(xout); }  // permutation_kernel2 :144-150
  }
};

template <typename T>
static void spmv(const Csr<T>& A, const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < A.nrows; i++) {
    double s = 0;
    for (int e = A.ptr[i]; e < A.ptr[i + 1]; e++) s += (double)A.val[e] * x[A.col[e]];
    y[i] = s;
  }
}
static double dot(const dvec& a, const dvec& b) {
  double s = 0; size_t n = a.size();
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (size_t i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}

struct SolveStats { int iters = 0; double final_relres = 0; double t_solve = 0; std::vector<double> resid; };

// CG_Flex_Cycle, cgcycle.cu:6-69, with cycle_level0 (amg_level.cu:74-128) as the
// preconditioner, incl. the T-rounded write-back of r (:125-126).
template <typename T>
static void pcg(Hierarchy<T>& H, const double* b_user, double* x_user, SolveStats& st) {
  const Csr<T>& A = H.levels[0].A;  // Ahyb_d_CG = permuted A (amg.cu:115-117)
  int N = A.nrows;
  bool coarsest_only = H.levels.size() == 1;
  bool to_perm = !H.prm.ref_level0_noperm && !coarsest_only;
  const ivec* iperm = coarsest_only ? nullptr : &H.levels[0].agg.ipermutation;
  dvec b(N), x(N), y(N), z(N), r(N), p(N);
  for (int i = 0; i < N; i++) { int s = to_perm ? (*iperm)[i] : i; b[i] = b_user[s]; x[i] = x_user[s]; }
  // In "correct" mode the whole iteration lives in the permuted numbering, so the
  // level-0 cycle must not permute again: temporarily force noperm for level 0.
  int saved = H.prm.ref_level0_noperm;
  H.prm.ref_level0_noperm = 1;
  auto precond = [&](dvec& rr, dvec& zz) {
    std::vector<T> bt(N), xt(N, 0);
    for (int i = 0; i < N; i++) bt[i] = (T)rr[i];
    H.cycle(0, bt, xt);
    for (int i = 0; i < N; i++) { zz[i] = (double)xt[i]; rr[i] = (double)bt[i]; }
  };
  double bnorm = sqrt(dot(b, b));
  spmv(A, x.data(), y.data());
  for (int i = 0; i < N; i++) r[i] = b[i] - y[i];
  precond(r, z);
  p = z;
  double rzold = dot(r, z), rznew;
  int niter = 0;
  st.resid.clear();
  while (niter < H.prm.maxIters) {
    spmv(A, p.data(), y.data());
    double yp = dot(y, p);
    double alpha = rzold / yp;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) { x[i] += alpha * p[i]; r[i] += -alpha * y[i]; }
    double normr = sqrt(dot(r, r));
    st.resid.push_back(normr / bnorm);
    st.final_relres = normr / bnorm;
    if ((normr / bnorm) <= H.prm.tolerance) break;
    niter++;
    precond(r, z);
    rznew = dot(z, r);
    double beta = rznew / rzold;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) p[i] = z[i] + beta * p[i];
    rzold = rznew;
  }
  H.prm.ref_level0_noperm = saved;
  st.iters = niter;
  for (int i = 0; i < N; i++) { int s = to_perm ? (*iperm)[i] : i; x_user[s] = x[i]; }
}

}  // namespace orc

// =============================================================================
// C API (ctypes) — handle based.
// =============================================================================
using namespace orc;

struct OracleHandle {
  int precision = 64;  // 64: everything fp64 (north star); 32: reference mixed precision (AMGType=float hierarchy, fp64 PCG)
  Params prm;
  int nv = 0;
  ivec ptr, col, xadj, adj;
  dvec val;
  Hierarchy<double> Hd;
  Hierarchy<float> Hf;
  SolveStats st;
  double t_pattern = 0, t_assemble = 0, t_setup = 0, t_solve = 0;
  bool has_setup = false;
};

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }


extern "C" void* orc_create(int precision) { OracleHandle* h = new OracleHandle(); h->precision = precision; return h; }
extern "C" void orc_destroy(void* hh) { delete (OracleHandle*)hh; }

extern "C" void orc_set_threads(int t) {
#ifdef _OPENMP
  omp_set_num_threads(t);
#else
  (void)t;
#endif
}
extern "C" int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

extern "C" int orc_set_param(void* hh, const char* name, double v) {
  OracleHandle* h = (OracleHandle*)hh; Params& p = h->prm; std::string s(name);
  if (s == "maxLevels") p.maxLevels = (int)v; else if (s == "maxIters") p.maxIters = (int)v;
  else if (s == "preInnerIters") p.preInner = (int)v; else if (s == "postInnerIters") p.postInner = (int)v;
  else if (s == "postRelaxes") p.postRelaxes = (int)v; else if (s == "topSize") p.topSize = (int)v;
  else if (s == "randMisParameters") p.randMisParameters = (int)v; else if (s == "partitionMaxSize") p.partitionMaxSize = (int)v;
  else if (s == "aggregatorType") p.aggregatorType = (int)v; else if (s == "solverType") p.solverType = (int)v;
  else if (s == "tolerance") p.tolerance = v; else if (s == "smootherWeight") p.smootherWeight = v;
  else if (s == "proOmega") p.proOmega = v; else if (s == "seed") p.seed = (unsigned)v;
  else if (s == "refLevel0NoPerm") p.ref_level0_noperm = (int)v;
  else return -1;
  return 0;
}

// mesh -> pattern.  npe = 4 (tets) or 3 (tris).  Returns nnz.
extern "C" int orc_pattern(void* hh, int nv, int ne, int npe, const int* elems) {
  OracleHandle* h = (OracleHandle*)hh;
  double t0 = now_s();
  h->nv = nv;
  pattern_from_mesh(nv, ne, npe, elems, h->ptr, h->col, h->xadj, h->adj);
  h->val.assign(h->col.size(), 0.0);
  h->t_pattern = now_s() - t0;
  h->has_setup = false;
  return (int)h->col.size();
}
extern "C" void orc_get_pattern(void* hh, int* ptr, int* col) {
  OracleHandle* h = (OracleHandle*)hh;
  std::copy(h->ptr.begin(), h->ptr.end(), ptr); std::copy(h->col.begin(), h->col.end(), col);
}
extern "C" void orc_assemble_tet(void* hh, int ne, const int* tets, const double* vx, const double* vy, const double* vz, const int* labels, int closed_form) {
  OracleHandle* h = (OracleHandle*)hh; double t0 = now_s();
  assemble_tet(h->nv, ne, tets, vx, vy, vz, labels, h->ptr, h->col, h->val.data(), closed_form != 0);
  h->t_assemble = now_s() - t0; h->has_setup = false;
}
extern "C" void orc_assemble_tri(void* hh, int ne, const int* tris, const double* vx, const double* vy, int closed_form) {
  OracleHandle* h = (OracleHandle*)hh; double t0 = now_s();
  assemble_tri(h->nv, ne, tris, vx, vy, h->ptr, h->col, h->val.data(), closed_form != 0);
  h->t_assemble = now_s() - t0; h->has_setup = false;
}
extern "C" void orc_get_values(void* hh, double* v) { OracleHandle* h = (OracleHandle*)hh; std::copy(h->val.begin(), h->val.end(), v); }
extern "C" void orc_set_values(void* hh, const double* v) { OracleHandle* h = (OracleHandle*)hh; std::copy(v, v + h->val.size(), h->val.begin()); h->has_setup = false; }
extern "C" void orc_tet_mass_integrals(double* out10) { tet_mass_integrals(out10); }

template <typename T>
static void do_setup(OracleHandle* h, Hierarchy<T>& H) {
  Csr<T> A; A.nrows = A.ncols = h->nv; A.ptr = h->ptr; A.col = h->col; A.val.resize(h->val.size());
  for (size_t i = 0; i < h->val.size(); i++) A.val[i] = (T)h->val[i];
  H.prm = h->prm;
  H.setup(A, h->xadj, h->adj);
}
extern "C" int orc_setup(void* hh) {
  OracleHandle* h = (OracleHandle*)hh; double t0 = now_s();
  try { if (h->precision == 64) do_setup(h, h->Hd); else do_setup(h, h->Hf); }
  catch (std::exception& e) { fprintf(stderr, "oracle setup failed: %s\n", e.what()); return -1; }
  h->t_setup = now_s() - t0; h->has_setup = true;
  return h->precision == 64 ? (int)h->Hd.levels.size() : (int)h->Hf.levels.size();
}

template <typename T>
static int level_int_array(Hierarchy<T>& H, int lev, const std::string& s, const ivec** out) {
  if (lev < 0 || lev >= (int)H.levels.size()) return -1;
  Level<T>& L = H.levels[lev];
  if (s == "permutation") *out = &L.agg.permutation; else if (s == "ipermutation") *out = &L.agg.ipermutation;
  else if (s == "aggregateIdx") *out = &L.agg.aggregateIdx; else if (s == "partitionIdx") *out = &L.agg.partitionIdx;
  else if (s == "partitionLabel") *out = &L.agg.partitionLabel; else if (s == "xadjOut") *out = &L.agg.xadjOut;
  else if (s == "adjOut") *out = &L.agg.adjOut; else if (s == "A_ptr") *out = &L.A.ptr; else if (s == "A_col") *out = &L.A.col;
  else if (s == "P_ptr") *out = &L.P.ptr; else if (s == "P_col") *out = &L.P.col;
  else if (s == "R_ptr") *out = &L.R.ptr; else if (s == "R_col") *out = &L.R.col;
  else if (s == "AinBlockIdx") *out = &L.AinBlockIdx; else if (s == "AoutBlockIdx") *out = &L.AoutBlockIdx;
  else if (s == "fineAggregate") *out = &L.agg.fineAggregate;
  else return -1;
  return 0;
}
// size query (buf == NULL) or copy of a per-level integer array
extern "C" int orc_level_int(void* hh, int lev, const char* name, int* buf) {
  OracleHandle* h = (OracleHandle*)hh; const ivec* v = nullptr;
  int rc = h->precision == 64 ? level_int_array(h->Hd, lev, name, &v) : level_int_array(h->Hf, lev, name, &v);
  if (rc) return -1;
  if (buf) std::copy(v->begin(), v->end(), buf);
  return (int)v->size();
}
template <typename T>
static int level_val_array(Hierarchy<T>& H, int lev, const std::string& s, double* buf) {
  if (lev < 0 || lev >= (int)H.levels.size()) return -1;
  Level<T>& L = H.levels[lev];
  const std::vector<T>* v = nullptr;
  if (s == "A_val") v = &L.A.val; else if (s == "P_val") v = &L.P.val; else if (s == "R_val") v = &L.R.val; else if (s == "diag") v = &L.diag;
  else return -1;
  if (buf) for (size_t i = 0; i < v->size(); i++) buf[i] = (double)(*v)[i];
  return (int)v->size();
}
extern "C" int orc_level_val(void* hh, int lev, const char* name, double* buf) {
  OracleHandle* h = (OracleHandle*)hh;
  return h->precision == 64 ? level_val_array(h->Hd, lev, name, buf) : level_val_array(h->Hf, lev, name, buf);
}
extern "C" int orc_num_levels(void* hh) { OracleHandle* h = (OracleHandle*)hh; return h->precision == 64 ? (int)h->Hd.levels.size() : (int)h->Hf.levels.size(); }
extern "C" int orc_level_rows(void* hh, int lev) { OracleHandle* h = (OracleHandle*)hh; return h->precision == 64 ? h->Hd.levels[lev].A.nrows : h->Hf.levels[lev].A.nrows; }

// FEMSolver::solveFEM -> AMG::solve (amg.cu:149-200): solverType 0 = ONE V-cycle
// (SURVEY F1), solverType 1 = PCG.  x is the initial guess and the result.
template <typename T>
static void do_solve(OracleHandle* h, Hierarchy<T>& H, const double* b, double* x) {
  H.prm = h->prm;
  int N = h->nv;
  if (h->prm.solverType == 1) { pcg(H, b, x, h->st); return; }
  std::vector<T> bt(N), xt(N);
  for (int i = 0; i < N; i++) { bt[i] = (T)b[i]; xt[i] = (T)x[i]; }
  H.cycle(0, bt, xt);
  for (int i = 0; i < N; i++) x[i] = (double)xt[i];
  h->st.iters = 1; h->st.final_relres = -1; h->st.resid.clear();
}
extern "C" int orc_solve(void* hh, const double* b, double* x) {
  OracleHandle* h = (OracleHandle*)hh;
  if (!h->has_setup) return -1;
  double t0 = now_s();
  if (h->precision == 64) do_solve(h, h->Hd, b, x); else do_solve(h, h->Hf, b, x);
  h->t_solve = now_s() - t0;
  return h->st.iters;
}
// one application of the preconditioner z = M^-1 r in the level-0 PERMUTED numbering (no level-0 permutation)
template <typename T>
static void do_precond(OracleHandle* h, Hierarchy<T>& H, const double* r, double* z) {
  int N = h->nv; H.prm = h->prm; int saved = H.prm.ref_level0_noperm; H.prm.ref_level0_noperm = 1;
  std::vector<T> bt(N), xt(N, 0);
  for (int i = 0; i < N; i++) bt[i] = (T)r[i];
  H.cycle(0, bt, xt);
  for (int i = 0; i < N; i++) z[i] = (double)xt[i];
  H.prm.ref_level0_noperm = saved;
}
extern "C" int orc_precond_permuted(void* hh, const double* r, double* z) {
  OracleHandle* h = (OracleHandle*)hh;
  if (!h->has_setup) return -1;
  if (h->precision == 64) do_precond(h, h->Hd, r, z); else do_precond(h, h->Hf, r, z);
  return 0;
}
extern "C" double orc_final_relres(void* hh) { return ((OracleHandle*)hh)->st.final_relres; }
extern "C" int orc_resid_history(void* hh, double* buf) {
  OracleHandle* h = (OracleHandle*)hh;
  if (buf) std::copy(h->st.resid.begin(), h->st.resid.end(), buf);
  return (int)h->st.resid.size();
}
extern "C" double orc_time(void* hh, const char* what) {
  OracleHandle* h = (OracleHandle*)hh; std::string s(what);
  if (s == "pattern") return h->t_pattern; if (s == "assemble") return h->t_assemble;
  if (s == "setup") return h->t_setup; if (s == "solve") return h->t_solve;
  return -1;
}
// y = A x with the user-ordered assembled matrix (helper for tests)
extern "C" void orc_spmv(void* hh, const double* x, double* y) {
  OracleHandle* h = (OracleHandle*)hh;
  Csr<double> A; A.nrows = A.ncols = h->nv; A.ptr = h->ptr; A.col = h->col; A.val = h->val;
  spmv(A, x, y);
}
// standalone aggregation on a user graph (tests): returns nAgg; arrays sized by caller via queries through a temp handle
extern "C" int orc_randomized_mis(int n, const int* xadj, const int* adj, int k, unsigned seed, int* mis) {
  ivec xa(xadj, xadj + n + 1), ad(adj, adj + xadj[n]), m;
  randomized_mis(xa, ad, m, k, seed);
  std::copy(m.begin(), m.end(), mis);
  return 0;
}

