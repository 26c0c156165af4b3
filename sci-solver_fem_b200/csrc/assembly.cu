// assembly.cu — stage 1b: P1 element stiffness + mass, summed straight into the CSR pattern.
//
// Reference: element_loop_3d_kernel (src/core/cuda/perform_element_loop_3D.cuh:480-627) with
// compute_stiffness_matrix_3d (:28-48), compute_massmatrix_vector_3d (:74-116) and the scatter
// sum_into_global_linear_system_cuda_3d (:163-318); 2-D twin element_loop_kernel
// (perform_element_loop_2D.cuh:287-372, :29-48, :71-147, :190-285).  Operator: A = K + 1.0*M.
//
// B200 design: no binary searches and no atomics.  Kernel 1 computes the npe*npe contributions
// of every element (one thread per element, int4 element load, 24-byte vertex gathers) into a
// contribution buffer; kernel 2 walks each matrix entry's gather list (built once with the
// pattern, pattern.cu) and sums in element order.  The result is deterministic and — because
// this file is compiled with -fmad=false and evaluates the reference's expression trees —
// bit-identical to the host-order sum of the reference's own host twin
// (element_loop_3d_host, perform_element_loop_3D.cuh:629-756).
//
// NOTE: keep -fmad=false for this translation unit (see build.py): the reference's cofactor
// formulas cancel catastrophically (24 O(1) terms summing to O(h^3)), so a fused multiply-add
// anywhere changes the assembled values at the 1e-10 relative level on fine meshes.
#include "fsb_internal.h"

namespace fsb {

struct TetQuad { double integrand[10]; };
struct TriQuad { double zx[6], zy[6], wx[6], wy[6]; };

__global__ void __launch_bounds__(128) tet_contrib_kernel(long long ne, const int4* __restrict__ tets, const double* __restrict__ xyz,
                                                          const int* __restrict__ labels, TetQuad q, double* __restrict__ ce) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int4 t = tets[e];
  int ids[4] = {t.x, t.y, t.z, t.w};
  double x[4], y[4], z[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const double* v = xyz + 3ll * ids[j];
    x[j] = v[0]; y[j] = v[1]; z[j] = v[2];
  }
  double a1 = x[1] - x[3], a2 = y[1] - y[3], a3 = z[1] - z[3];
  double b1 = x[2] - x[3], b2 = y[2] - y[3], b3 = z[2] - z[3];
  double c1 = x[0] - x[3], c2 = y[0] - y[3], c3 = z[0] - z[3];
  double Tvol = fabs(c1 * (a2 * b3 - a3 * b2) + c2 * (a3 * b1 - a1 * b3) + c3 * (a1 * b2 - a2 * b1)) / 6.0;
  // [x y z 1]^-1 by cofactors; the first three entries of column k are grad(phi_k)
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
  double g[4][3];
  g[0][0] = (a22 * a33 * a44 + a23 * a34 * a42 + a24 * a32 * a43 - a22 * a34 * a43 - a23 * a32 * a44 - a24 * a33 * a42) / det;
  g[0][1] = (a21 * a34 * a43 + a23 * a31 * a44 + a24 * a33 * a41 - a21 * a33 * a44 - a23 * a34 * a41 - a24 * a31 * a43) / det;
  g[0][2] = (a21 * a32 * a44 + a22 * a34 * a41 + a24 * a31 * a42 - a21 * a34 * a42 - a22 * a31 * a44 - a24 * a32 * a41) / det;
  g[1][0] = (a12 * a34 * a43 + a13 * a32 * a44 + a14 * a33 * a42 - a12 * a33 * a44 - a13 * a34 * a42 - a14 * a32 * a43) / det;
  g[1][1] = (a11 * a33 * a44 + a13 * a34 * a41 + a14 * a31 * a43 - a11 * a34 * a43 - a13 * a31 * a44 - a14 * a33 * a41) / det;
  g[1][2] = (a11 * a34 * a42 + a12 * a31 * a44 + a14 * a32 * a41 - a11 * a32 * a44 - a12 * a34 * a41 - a14 * a31 * a42) / det;
  g[2][0] = (a12 * a23 * a44 + a13 * a24 * a42 + a14 * a22 * a43 - a12 * a24 * a43 - a13 * a22 * a44 - a14 * a23 * a42) / det;
  g[2][1] = (a11 * a24 * a43 + a13 * a21 * a44 + a14 * a23 * a41 - a11 * a23 * a44 - a13 * a24 * a41 - a14 * a21 * a43) / det;
  g[2][2] = (a11 * a22 * a44 + a12 * a24 * a41 + a14 * a21 * a42 - a11 * a24 * a42 - a12 * a21 * a44 - a14 * a22 * a41) / det;
  g[3][0] = (a12 * a24 * a33 + a13 * a22 * a34 + a14 * a23 * a32 - a12 * a23 * a34 - a13 * a24 * a32 - a14 * a22 * a33) / det;
  g[3][1] = (a11 * a23 * a34 + a13 * a24 * a31 + a14 * a21 * a33 - a11 * a24 * a33 - a13 * a21 * a34 - a14 * a23 * a31) / det;
  g[3][2] = (a11 * a24 * a32 + a12 * a21 * a34 + a14 * a22 * a31 - a11 * a22 * a34 - a12 * a24 * a31 - a14 * a21 * a32) / det;
  // material coefficient: labels 0..6 -> 1,1,2,3,4,5,6 (perform_element_loop_3D.cuh:585-610);
  // other labels are undefined upstream (stale register) — defined as 1.0 here.
  int lab = labels ? labels[e] : 0;
  double co = (lab >= 2 && lab <= 6) ? (double)lab : 1.0;
  // |det J|/8 of the map from [-1,1]^3 (compute_massmatrix_vector_3d)
  double x1 = x[0], y1 = y[0], z1 = z[0], x2 = x[1], y2 = y[1], z2 = z[1], x3 = x[2], y3 = y[2], z3 = z[2], x4 = x[3], y4 = y[3], z4 = z[3];
  double dj = 0.125 * ((-x1 + x2) * (-y1 + y3) * (-z1 + z4) + (-y1 + y2) * (-z1 + z3) * (-x1 + x4) + (-z1 + z2) * (-x1 + x3) * (-y1 + y4)
    - (-x1 + x2) * (-z1 + z3) * (-y1 + y4) - (-z1 + z2) * (-y1 + y3) * (-x1 + x4) - (-y1 + y2) * (-x1 + x3) * (-z1 + z4));
  double jac = fabs(dj);
  double A[4][4];
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int gg = k; gg < 4; gg++) {
      double st = (g[k][0] * g[gg][0] + g[k][1] * g[gg][1] + g[k][2] * g[gg][2]) * Tvol * co;
      double ms = q.integrand[cnt] * jac;
      A[k][gg] = st + 1.0 * ms;
      cnt++;
    }
  double* out = ce + e * 16;
  // slots 0..11: pairs (0,1)(0,2)(0,3)(1,2)(1,3)(2,3) in both directions; 12..15: diagonals
  out[0] = A[0][1]; out[1] = A[0][1]; out[2] = A[0][2]; out[3] = A[0][2]; out[4] = A[0][3]; out[5] = A[0][3];
  out[6] = A[1][2]; out[7] = A[1][2]; out[8] = A[1][3]; out[9] = A[1][3]; out[10] = A[2][3]; out[11] = A[2][3];
  out[12] = A[0][0]; out[13] = A[1][1]; out[14] = A[2][2]; out[15] = A[3][3];
}

__global__ void __launch_bounds__(128) tri_contrib_kernel(long long ne, const int* __restrict__ tris, const double* __restrict__ xyz,
                                                          TriQuad q, double* __restrict__ ce) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int ids[3] = {tris[3 * e], tris[3 * e + 1], tris[3 * e + 2]};
  double x[3], y[3];
#pragma unroll
  for (int j = 0; j < 3; j++) { x[j] = xyz[3ll * ids[j]]; y[j] = xyz[3ll * ids[j] + 1]; }
  double TArea = fabs(x[0] * y[2] - x[0] * y[1] + x[1] * y[0] - x[1] * y[2] + x[2] * y[1] - x[2] * y[0]) / 2.0;
  double a11 = x[0], a12 = y[0], a13 = 1.0, a21 = x[1], a22 = y[1], a23 = 1.0, a31 = x[2], a32 = y[2], a33 = 1.0;
  double det = a11 * a22 * a33 + a21 * a32 * a13 + a31 * a12 * a23 - a11 * a32 * a23 - a31 * a22 * a13 - a21 * a12 * a33;
  double cf[3][3];  // cf[k] = (a, b, c) of phi_k = a x + b y + c
  cf[0][0] = (a22 * a33 - a23 * a32) / det; cf[0][1] = (a23 * a31 - a21 * a33) / det; cf[0][2] = (a21 * a32 - a22 * a31) / det;
  cf[1][0] = (a13 * a32 - a12 * a33) / det; cf[1][1] = (a11 * a33 - a13 * a31) / det; cf[1][2] = (a12 * a31 - a11 * a32) / det;
  cf[2][0] = (a12 * a23 - a13 * a22) / det; cf[2][1] = (a13 * a21 - a11 * a23) / det; cf[2][2] = (a11 * a22 - a12 * a21) / det;
  // signed jacobian / 8 (the reference uses fabs for the area but not here: perform_element_loop_2D.cuh:92)
  double jac = (x[0] * y[1] - x[1] * y[0] - x[0] * y[2] + x[2] * y[0] + x[1] * y[2] - x[2] * y[1]) / 8;
  double A[3][3];
#pragma unroll
  for (int k = 0; k < 3; k++)
#pragma unroll
    for (int gg = k; gg < 3; gg++) {
      double st = (cf[k][0] * cf[gg][0] + cf[k][1] * cf[gg][1]) * TArea;
      double integral = 0;
      for (int p = 0; p < 6; p++) {
        double ty = 0.0;
        for (int r = 0; r < 6; r++) {
          double qx = x[0] * (1 - q.zx[p]) * 0.5 * (1 - q.zy[r]) * 0.5 + x[1] * (1 + q.zx[p]) * 0.5 * (1 - q.zy[r]) * 0.5 + x[2] * (1 + q.zy[r]) * 0.5;
          double qy = y[0] * (1 - q.zx[p]) * 0.5 * (1 - q.zy[r]) * 0.5 + y[1] * (1 + q.zx[p]) * 0.5 * (1 - q.zy[r]) * 0.5 + y[2] * (1 + q.zy[r]) * 0.5;
          ty += ((cf[k][0] * qx + cf[k][1] * qy + cf[k][2]) * (cf[gg][0] * qx + cf[gg][1] * qy + cf[gg][2]) * jac) * q.wy[r];
        }
        integral += ty * q.wx[p];
      }
      A[k][gg] = st + 1.0 * integral;
    }
  double* out = ce + e * 9;
  out[0] = A[0][1]; out[1] = A[0][1]; out[2] = A[0][2]; out[3] = A[0][2]; out[4] = A[1][2]; out[5] = A[1][2];
  out[6] = A[0][0]; out[7] = A[1][1]; out[8] = A[2][2];
}

// one thread per matrix entry: fixed-order sum of its gather list
__global__ void gather_sum_kernel(int nnz, const long long* __restrict__ seg, const uint32_t* __restrict__ contrib,
                                  const double* __restrict__ ce, double* __restrict__ val) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  double s = 0.0;
  for (long long q = seg[k]; q < seg[k + 1]; q++) {
    uint32_t c = contrib[q];
    if (c != 0xFFFFFFFFu) s += ce[c];
  }
  val[k] = s;
}

void assemble_values(const Ctx& c, const Mesh& m, const Pattern& p, double* val) {
  cudaStream_t s = c.stream;
  const int ns = m.npe * m.npe;
  DBuf ce((size_t)m.ne * ns, s);
  if (m.npe == 4) {
    TetQuad q;
    tet_mass_integrals_host(q.integrand);
    tet_contrib_kernel<<<cdiv(m.ne, 128), 128, 0, s>>>(m.ne, reinterpret_cast<const int4*>(m.elems.get()), m.xyz,
                                                      m.labels.size() ? m.labels.get() : nullptr, q, ce);
  } else {
    TriQuad q;
    tri_quadrature_host(q.zx, q.zy, q.wx, q.wy);
    tri_contrib_kernel<<<cdiv(m.ne, 128), 128, 0, s>>>(m.ne, m.elems, m.xyz, q, ce);
  }
  FSB_CHECK_LAUNCH();
  gather_sum_kernel<<<cdiv(p.nnz, 256), 256, 0, s>>>(p.nnz, p.seg, p.contrib, ce, val);
  FSB_CHECK_LAUNCH();
}

}  // namespace fsb
