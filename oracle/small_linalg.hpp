// ORACLE — TEST INFRASTRUCTURE ONLY (see types.hpp header).  PARITY UNPINNED.
//
// small_linalg.hpp: the two Eigen 3.3 dense kernels laserMapping calls
// (third party, not vendored): SelfAdjointEigenSolver<Matrix3d>
// (laser_mapping.cpp:500: eigenvalues ascending, unit eigenvectors, sign
// arbitrary) and Matrix<double,5,3>::colPivHouseholderQr().solve()
// (laser_mapping.cpp:557: least-squares solution through a column-pivoted
// Householder QR; also visual_odometry.cpp:350-355 for 3x3).  Any backward
// stable method agrees with Eigen to O(1e-15) relative; the cyclic Jacobi
// method and a textbook pivoted Householder QR are used.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>

namespace oracle {

// A symmetric (row-major 3x3).  evals ascending, evecs[k] = k-th eigenvector (unit).
inline void sym_eig3(const double A[9], double evals[3], double evecs[3][3]) {
  double a[3][3] = {{A[0], A[1], A[2]}, {A[3], A[4], A[5]}, {A[6], A[7], A[8]}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-300 || off <= 1e-32 * diag) break;
    for (int p = 0; p < 2; ++p) {
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A * G
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- G^T * A
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
    }
  }
  int order[3] = {0, 1, 2};
  std::sort(order, order + 3, [&](int i, int j) { return a[i][i] < a[j][j]; });
  for (int k = 0; k < 3; ++k) {
    evals[k] = a[order[k]][order[k]];
    double n = 0.0;
    for (int r = 0; r < 3; ++r) n += v[r][order[k]] * v[r][order[k]];
    n = std::sqrt(n);
    for (int r = 0; r < 3; ++r) evecs[k][r] = v[r][order[k]] / n;
  }
}

// Least-squares solve of A (m x 3, row-major) x = b through a column-pivoted Householder QR.
// Negligible pivots (Eigen's threshold: column norm^2 <= (eps/m)^2 * max initial norm^2 ... ) give zero components.
template <int M>
inline void colpiv_qr_solve3(const double Ain[M * 3], const double bin[M], double x[3]) {
  double A[M][3], b[M];
  for (int i = 0; i < M; ++i) { b[i] = bin[i]; for (int c = 0; c < 3; ++c) A[i][c] = Ain[i * 3 + c]; }
  int perm[3] = {0, 1, 2};
  double maxnorm2 = 0.0;
  for (int c = 0; c < 3; ++c) { double s = 0; for (int i = 0; i < M; ++i) s += A[i][c] * A[i][c]; maxnorm2 = std::max(maxnorm2, s); }
  const double eps = std::numeric_limits<double>::epsilon();
  const double thresh = maxnorm2 * (eps / M) * (eps / M);
  int rank = 0;
  for (int k = 0; k < 3 && k < M; ++k) {
    int best = k; double bestn = -1.0;
    for (int c = k; c < 3; ++c) { double s = 0; for (int i = k; i < M; ++i) s += A[i][c] * A[i][c]; if (s > bestn) { bestn = s; best = c; } }
    if (bestn <= thresh) break;
    if (best != k) { for (int i = 0; i < M; ++i) std::swap(A[i][k], A[i][best]); std::swap(perm[k], perm[best]); }
    const double nrm = std::sqrt(bestn);
    const double alpha = A[k][k] > 0 ? -nrm : nrm;
    double v[M];
    for (int i = k; i < M; ++i) v[i] = A[i][k];
    v[k] -= alpha;
    double vn = 0; for (int i = k; i < M; ++i) vn += v[i] * v[i];
    if (vn > 0) {
      for (int c = k; c < 3; ++c) {
        double s = 0; for (int i = k; i < M; ++i) s += v[i] * A[i][c];
        s = 2.0 * s / vn;
        for (int i = k; i < M; ++i) A[i][c] -= s * v[i];
      }
      double s = 0; for (int i = k; i < M; ++i) s += v[i] * b[i];
      s = 2.0 * s / vn;
      for (int i = k; i < M; ++i) b[i] -= s * v[i];
    }
    ++rank;
  }
  double y[3] = {0, 0, 0};
  for (int k = rank - 1; k >= 0; --k) {
    double s = b[k];
    for (int c = k + 1; c < rank; ++c) s -= A[k][c] * y[c];
    y[k] = s / A[k][k];
  }
  for (int k = 0; k < 3; ++k) x[perm[k]] = y[k];
}

}  // namespace oracle
