// quadrature.cpp — host-side Gauss-Jacobi tables the element kernels consume.
//
// The reference computes these on the host as well and ships them to the device:
// FEM3D::assemble (src/core/cuda/FEM3D.cu:501-557) builds Gauss-Lobatto-Jacobi(0,0) x
// Gauss-Radau-Jacobi(1,0) x Gauss-Radau-Jacobi(2,0) rules of degree 4 and from them the 10
// reference mass integrals (IntegrationInTet :452-475); FEM2D::assemble (FEM2D.cu:358-378)
// builds GLJ(0,0) x GRJ(1,0) of degree 6.  Zeros come from a deflated Newton iteration
// (JacobiGZeros, FEM3D.cu:216-274: tolerance 1e-6 on the step, PI truncated to 3.1415927 for
// the initial guess only), so the tables are NOT exact Gauss rules to the last bit — the same
// iteration is carried out here so the assembled values match the reference's to the bit.
#include <cmath>
#include <vector>

#include "fsb_internal.h"

namespace fsb {
namespace {

typedef std::vector<double> Vec;

// Gamma at the (positive integer) arguments the weight formulas need: (x-1)!  (FEM3D.cu:92-101)
double gamma_int(int x) {
  double ga = 1.0;
  for (int i = 2; i < x; i++) ga *= i;
  return ga;
}

// P_n^{(alpha,beta)}(x) by the three-term recurrence (JacobiPoly, FEM3D.cu:138-187)
double jacobi_p(int n, double x, int alpha, int beta) {
  if (n == 0) return 1.0;
  if (n == 1) return 0.5 * (alpha - beta + (alpha + beta + 2.0) * x);
  double m = n - 1.0;
  double t = 2.0 * m + alpha + beta;
  double a1 = 2.0 * (m + 1) * (m + alpha + beta + 1) * t;
  double a2 = (t + 1) * (alpha * alpha - beta * beta);
  double a3 = t * (t + 1.0) * (t + 2.0);
  double a4 = 2.0 * (m + alpha) * (m + beta) * (t + 2.0);
  double p1 = jacobi_p(n - 1, x, alpha, beta), p2 = jacobi_p(n - 2, x, alpha, beta);
  return ((a2 + a3 * x) * p1 - a4 * p2) / a1;
}

double jacobi_dp(int n, double x, int alpha, int beta) {
  if (n == 0) return 0.0;
  return 0.5 * (alpha + beta + n + 1) * jacobi_p(n - 1, x, alpha + 1, beta + 1);
}

Vec gauss_zeros(int n, int alpha, int beta) {
  Vec z(n, 0.0);
  if (n == 0) return z;
  const double pi_ref = 3.1415927, eps = 1.0e-6;
  double dth = pi_ref / (2.0 * n), rlast = 0.0;
  for (int k = 0; k < n; k++) {
    double r = -cos((2.0 * k + 1.0) * dth);
    if (k) r = 0.5 * (r + rlast);
    for (int it = 0; it < 60; it++) {
      double poly = jacobi_p(n, r, alpha, beta), pder = jacobi_dp(n, r, alpha, beta);
      double sum = 0.0;
      for (int i = 0; i < k; i++) sum = sum + 1.0 / (r - z[i]);
      double delr = -poly / (pder - sum * poly);
      r = r + delr;
      if (fabs(delr) < eps) break;
    }
    z[k] = r;
    rlast = r;
  }
  return z;
}

// Gauss-Lobatto-Jacobi nodes/weights (JacobiGLZW, FEM3D.cu:276-322), degree >= 2
void lobatto(int n, int alpha, int beta, Vec& Z, Vec& W) {
  Z.assign(n, 0.0); W.assign(n, 0.0);
  Z[0] = -1; Z[n - 1] = 1;
  Vec in = gauss_zeros(n - 2, alpha + 1, beta + 1);
  for (int i = 1; i < n - 1; i++) Z[i] = in[i - 1];
  for (int i = 0; i < n; i++) W[i] = jacobi_p(n - 1, Z[i], alpha, beta);
  double fac = pow(2.0, (double)(alpha + beta + 1)) * gamma_int(alpha + n) * gamma_int(beta + n);
  fac = fac / ((n - 1) * gamma_int(n) * gamma_int(alpha + beta + n + 1));
  for (int i = 0; i < n; i++) W[i] = fac / (W[i] * W[i]);
  W[0] = W[0] * (beta + 1);
  W[n - 1] = W[n - 1] * (alpha + 1);
}

// Gauss-Radau-Jacobi nodes/weights, node at -1 (JacobiGRZW, FEM3D.cu:324-368), degree >= 2
void radau(int n, int alpha, int beta, Vec& Z, Vec& W) {
  Z.assign(n, 0.0); W.assign(n, 0.0);
  Z[0] = -1;
  Vec in = gauss_zeros(n - 1, alpha, beta + 1);
  for (int i = 1; i < n; i++) Z[i] = in[i - 1];
  for (int i = 0; i < n; i++) W[i] = jacobi_p(n - 1, Z[i], alpha, beta);
  double fac = pow(2.0, (double)(alpha + beta)) * gamma_int(alpha + n) * gamma_int(beta + n);
  fac = fac / (gamma_int(n) * (beta + n) * gamma_int(alpha + beta + n + 1));
  for (int i = 0; i < n; i++) W[i] = fac * (1 - Z[i]) / (W[i] * W[i]);
  W[0] = W[0] * (beta + 1);
}

}  // namespace

void tet_mass_integrals_host(double out[10]) {
  const int D = 4;
  Vec zx, zy, zz, wx, wy, wz;
  lobatto(D, 0, 0, zx, wx);
  radau(D, 1, 0, zy, wy);
  radau(D, 2, 0, zz, wz);
  for (int i = 0; i < D; i++) { wy[i] /= 2; wz[i] /= 4; }
  // phi_s at the collapsed-coordinate points: phi_0 = 1-X-Y-Z, phi_1 = X, phi_2 = Y, phi_3 = Z
  double phi[4][4][4][4];
  for (int i = 0; i < D; i++)
    for (int j = 0; j < D; j++)
      for (int k = 0; k < D; k++) {
        double X = (1 + zx[i]) * 0.5 * (1 - zy[j]) * 0.5 * (1 - zz[k]) * 0.5;
        double Y = (1 + zy[j]) * 0.5 * (1 - zz[k]) * 0.5;
        double Z = (1 + zz[k]) * 0.5;
        phi[0][i][j][k] = 1 + -1 * X + -1 * Y + -1 * Z;
        phi[1][i][j][k] = 0 + 1 * X + 0 * Y + 0 * Z;
        phi[2][i][j][k] = 0 + 0 * X + 1 * Y + 0 * Z;
        phi[3][i][j][k] = 0 + 0 * X + 0 * Y + 1 * Z;
      }
  int c = 0;
  for (int a = 0; a < 4; a++)
    for (int b = a; b < 4; b++) {
      double integral = 0;
      for (int p = 0; p < D; p++) {
        double ty = 0.0;
        for (int q = 0; q < D; q++) {
          double tz = 0.0;
          for (int r = 0; r < D; r++) tz += phi[a][p][q][r] * phi[b][p][q][r] * wz[r];
          ty += tz * wy[q];
        }
        integral += ty * wx[p];
      }
      out[c++] = integral;
    }
}

void tri_quadrature_host(double zx[6], double zy[6], double wx[6], double wy[6]) {
  Vec Zx, Zy, Wx, Wy;
  lobatto(6, 0, 0, Zx, Wx);
  radau(6, 1, 0, Zy, Wy);
  for (int i = 0; i < 6; i++) { zx[i] = Zx[i]; zy[i] = Zy[i]; wx[i] = Wx[i]; wy[i] = Wy[i]; }
}

}  // namespace fsb
