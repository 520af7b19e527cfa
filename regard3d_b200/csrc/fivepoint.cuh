// fivepoint.cuh -- device side of the minimal solver behind GeometricFilter_EMatrix_AC(4.0, 2048)
// (src/R3DComputeMatches.cpp:2169-2171): OpenMVG 1.4 multiview/solver_essential_five_point.cpp
// (Nister / Stewenius five-point algorithm on bearing vectors).
//   1. 4-D nullspace {X,Y,Z,W} of the 5x9 epipolar system, E = xX + yY + zZ + W
//   2. det(E) = 0 and 2 E E^T E - tr(E E^T) E = 0 as a 10x20 matrix over
//      [xxx xxy xxz xyy xyz xzz yyy yyz yzz zzz | xx xy xz yy yz zz x y z 1]
//   3. Gauss-Jordan elimination of the cubic block
//   4. action matrix of "multiply by x" on [xx xy xz yy yz zz x y z 1]; each REAL eigenvalue
//      (elimination-Hessenberg + Francis double-shift QR) gives one model from the null vector of
//      (A - lambda I)
// One thread runs one hypothesis; the 10x20 / 11x11 work arrays live in local memory (L1-resident).
// Only + - * / sqrt fabs and comparisons: with --fmad=false (build.py) every operation rounds once, so the
// discrete decisions of the a-contrario RANSAC built on top are reproducible (see detmath.cuh).
#pragma once

namespace r3d {
namespace fp {

__device__ constexpr int kP2[4][4] = {{0, 1, 2, 6}, {1, 3, 4, 7}, {2, 4, 5, 8}, {6, 7, 8, 9}};
__device__ constexpr int kP3[10][4] = {{0, 1, 2, 10}, {1, 3, 4, 11}, {2, 4, 5, 12}, {3, 6, 7, 13}, {4, 7, 8, 14},
                            {5, 8, 9, 15}, {10, 11, 12, 16}, {11, 13, 14, 17}, {12, 14, 15, 18}, {16, 17, 18, 19}};

// degree-1 x degree-1 -> degree-2 (10 coefficients), accumulated into out with sign
__device__ inline void mul11(const double* a, const double* b, double sign, double* out) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) out[kP2[i][j]] = out[kP2[i][j]] + sign * (a[i] * b[j]);
}
// degree-2 x degree-1 -> degree-3 (20 coefficients), accumulated into out with sign
__device__ inline void mul21(const double* a, const double* b, double sign, double* out) {
  for (int i = 0; i < 10; ++i)
    for (int j = 0; j < 4; ++j) out[kP3[i][j]] = out[kP3[i][j]] + sign * (a[i] * b[j]);
}

// 4-D nullspace of the 5x9 system, orthonormalised (modified Gram-Schmidt in the order found)
__device__ inline bool nullspace_5x9(double A[5][9], double basis[4][9]) {
  int colperm[9];
  for (int j = 0; j < 9; ++j) colperm[j] = j;
  for (int r = 0; r < 5; ++r) {
    int pi = r, pj = r;
    double best = fabs(A[r][r]);
    for (int i = r; i < 5; ++i)
      for (int j = r; j < 9; ++j) {
        const double v = fabs(A[i][j]);
        if (v > best) { best = v; pi = i; pj = j; }
      }
    if (!(best > 0.0)) return false;
    if (pi != r)
      for (int j = 0; j < 9; ++j) { const double t = A[r][j]; A[r][j] = A[pi][j]; A[pi][j] = t; }
    if (pj != r) {
      for (int i = 0; i < 5; ++i) { const double t = A[i][r]; A[i][r] = A[i][pj]; A[i][pj] = t; }
      const int t = colperm[r]; colperm[r] = colperm[pj]; colperm[pj] = t;
    }
    for (int i = r + 1; i < 5; ++i) {
      const double f = A[i][r] / A[r][r];
      for (int j = r + 1; j < 9; ++j) A[i][j] = A[i][j] - f * A[r][j];
      A[i][r] = 0.0;
    }
  }
  for (int t = 0; t < 4; ++t) {
    double z[9];
    for (int k = 5; k < 9; ++k) z[k] = (k == 5 + t) ? 1.0 : 0.0;
    for (int r = 4; r >= 0; --r) {
      double s = 0.0;
      for (int j = r + 1; j < 9; ++j) s = s + A[r][j] * z[j];
      z[r] = -s / A[r][r];
    }
    double v[9];
    for (int k = 0; k < 9; ++k) v[colperm[k]] = z[k];
    for (int u = 0; u < t; ++u) {
      double dp = 0.0;
      for (int k = 0; k < 9; ++k) dp = dp + v[k] * basis[u][k];
      for (int k = 0; k < 9; ++k) v[k] = v[k] - dp * basis[u][k];
    }
    double nn = 0.0;
    for (int k = 0; k < 9; ++k) nn = nn + v[k] * v[k];
    nn = sqrt(nn);
    if (!(nn > 0.0)) return false;
    for (int k = 0; k < 9; ++k) basis[t][k] = v[k] / nn;
  }
  return true;
}

// the ten cubic constraints: M is 10 x 20
__device__ inline void constraints(const double basis[4][9], double M[10][20]) {
  double E[3][3][4];  // entry (r,c) as a degree-1 polynomial in (x, y, z, 1)
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      for (int t = 0; t < 4; ++t) E[r][c][t] = basis[t][3 * r + c];
  // row 0: det(E)
  for (int k = 0; k < 20; ++k) M[0][k] = 0.0;
  {
    double m2[10];
    for (int k = 0; k < 10; ++k) m2[k] = 0.0;
    mul11(E[1][1], E[2][2], 1.0, m2);
    mul11(E[1][2], E[2][1], -1.0, m2);
    mul21(m2, E[0][0], 1.0, M[0]);
    for (int k = 0; k < 10; ++k) m2[k] = 0.0;
    mul11(E[1][0], E[2][2], 1.0, m2);
    mul11(E[1][2], E[2][0], -1.0, m2);
    mul21(m2, E[0][1], -1.0, M[0]);
    for (int k = 0; k < 10; ++k) m2[k] = 0.0;
    mul11(E[1][0], E[2][1], 1.0, m2);
    mul11(E[1][1], E[2][0], -1.0, m2);
    mul21(m2, E[0][2], 1.0, M[0]);
  }
  // EEt, L = EEt - 0.5 tr(EEt) I, rows 1..9: L E
  double EEt[3][3][10];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      for (int k = 0; k < 10; ++k) EEt[i][j][k] = 0.0;
      for (int t = 0; t < 3; ++t) mul11(E[i][t], E[j][t], 1.0, EEt[i][j]);
    }
  double tr[10];
  for (int k = 0; k < 10; ++k) tr[k] = 0.5 * ((EEt[0][0][k] + EEt[1][1][k]) + EEt[2][2][k]);
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 10; ++k) EEt[i][i][k] = EEt[i][i][k] - tr[k];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double* row = M[1 + 3 * i + j];
      for (int k = 0; k < 20; ++k) row[k] = 0.0;
      for (int t = 0; t < 3; ++t) mul21(EEt[i][t], E[t][j], 1.0, row);
    }
}

// [A | C] (10 x 20) -> A = I by Gauss-Jordan with partial pivoting; C becomes B
__device__ inline bool gauss_jordan(double M[10][20]) {
  for (int c = 0; c < 10; ++c) {
    int pi = c;
    double best = fabs(M[c][c]);
    for (int i = c + 1; i < 10; ++i)
      if (fabs(M[i][c]) > best) { best = fabs(M[i][c]); pi = i; }
    if (!(best > 0.0)) return false;
    if (pi != c)
      for (int j = 0; j < 20; ++j) { const double t = M[c][j]; M[c][j] = M[pi][j]; M[pi][j] = t; }
    const double piv = M[c][c];
    for (int j = c; j < 20; ++j) M[c][j] = M[c][j] / piv;
    for (int i = 0; i < 10; ++i) {
      if (i == c) continue;
      const double f = M[i][c];
      if (f == 0.0) continue;
      for (int j = c; j < 20; ++j) M[i][j] = M[i][j] - f * M[c][j];
    }
  }
  return true;
}

// reduction to upper Hessenberg form by stabilised elementary similarity transformations
// (EISPACK elmhes); a is 1-indexed [11][11]
__device__ inline void elmhes(double a[11][11], int n) {
  for (int m = 2; m < n; ++m) {
    double x = 0.0;
    int i = m;
    for (int j = m; j <= n; ++j)
      if (fabs(a[j][m - 1]) > fabs(x)) { x = a[j][m - 1]; i = j; }
    if (i != m) {
      for (int j = m - 1; j <= n; ++j) { const double t = a[i][j]; a[i][j] = a[m][j]; a[m][j] = t; }
      for (int j = 1; j <= n; ++j) { const double t = a[j][i]; a[j][i] = a[j][m]; a[j][m] = t; }
    }
    if (x != 0.0) {
      for (i = m + 1; i <= n; ++i) {
        double y = a[i][m - 1];
        if (y != 0.0) {
          y = y / x;
          a[i][m - 1] = y;
          for (int j = m; j <= n; ++j) a[i][j] = a[i][j] - y * a[m][j];
          for (int j = 1; j <= n; ++j) a[j][m] = a[j][m] + y * a[j][i];
        }
      }
    }
  }
  for (int i = 3; i <= n; ++i)
    for (int j = 1; j <= i - 2; ++j) a[i][j] = 0.0;
}

__device__ inline double sign_of(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

// eigenvalues of an upper Hessenberg matrix (EISPACK hqr: Francis double-shift QR); 1-indexed.
// returns false when 30 iterations do not deflate an eigenvalue.
__device__ inline bool hqr(double a[11][11], int n, double* wr, double* wi) {
  int nn, m, l, k, j, its, i, mmin;
  double z, y, x, w, v, u, t, s, r = 0.0, q = 0.0, p = 0.0, anorm = 0.0;
  for (i = 1; i <= n; ++i)
    for (j = (i - 1 > 1 ? i - 1 : 1); j <= n; ++j) anorm = anorm + fabs(a[i][j]);
  nn = n;
  t = 0.0;
  while (nn >= 1) {
    its = 0;
    do {
      for (l = nn; l >= 2; --l) {
        s = fabs(a[l - 1][l - 1]) + fabs(a[l][l]);
        if (s == 0.0) s = anorm;
        if (fabs(a[l][l - 1]) + s == s) {
          a[l][l - 1] = 0.0;
          break;
        }
      }
      x = a[nn][nn];
      if (l == nn) {
        wr[nn] = x + t;
        wi[nn--] = 0.0;
      } else {
        y = a[nn - 1][nn - 1];
        w = a[nn][nn - 1] * a[nn - 1][nn];
        if (l == nn - 1) {
          p = 0.5 * (y - x);
          q = p * p + w;
          z = sqrt(fabs(q));
          x = x + t;
          if (q >= 0.0) {
            z = p + sign_of(z, p);
            wr[nn - 1] = wr[nn] = x + z;
            if (z != 0.0) wr[nn] = x - w / z;
            wi[nn - 1] = wi[nn] = 0.0;
          } else {
            wr[nn - 1] = wr[nn] = x + p;
            wi[nn] = z;
            wi[nn - 1] = -z;
          }
          nn -= 2;
        } else {
          if (its == 30) return false;
          if (its == 10 || its == 20) {
            t = t + x;
            for (i = 1; i <= nn; ++i) a[i][i] = a[i][i] - x;
            s = fabs(a[nn][nn - 1]) + fabs(a[nn - 1][nn - 2]);
            y = x = 0.75 * s;
            w = -0.4375 * s * s;
          }
          ++its;
          for (m = nn - 2; m >= l; --m) {
            z = a[m][m];
            r = x - z;
            s = y - z;
            p = (r * s - w) / a[m + 1][m] + a[m][m + 1];
            q = a[m + 1][m + 1] - z - r - s;
            r = a[m + 2][m + 1];
            s = fabs(p) + fabs(q) + fabs(r);
            p = p / s;
            q = q / s;
            r = r / s;
            if (m == l) break;
            u = fabs(a[m][m - 1]) * (fabs(q) + fabs(r));
            v = fabs(p) * (fabs(a[m - 1][m - 1]) + fabs(z) + fabs(a[m + 1][m + 1]));
            if (u + v == v) break;
          }
          for (i = m + 2; i <= nn; ++i) {
            a[i][i - 2] = 0.0;
            if (i != m + 2) a[i][i - 3] = 0.0;
          }
          for (k = m; k <= nn - 1; ++k) {
            if (k != m) {
              p = a[k][k - 1];
              q = a[k + 1][k - 1];
              r = 0.0;
              if (k != nn - 1) r = a[k + 2][k - 1];
              x = fabs(p) + fabs(q) + fabs(r);
              if (x != 0.0) {
                p = p / x;
                q = q / x;
                r = r / x;
              }
            }
            s = sign_of(sqrt(p * p + q * q + r * r), p);
            if (s != 0.0) {
              if (k == m) {
                if (l != m) a[k][k - 1] = -a[k][k - 1];
              } else {
                a[k][k - 1] = -s * x;
              }
              p = p + s;
              x = p / s;
              y = q / s;
              z = r / s;
              q = q / p;
              r = r / p;
              for (j = k; j <= nn; ++j) {
                p = a[k][j] + q * a[k + 1][j];
                if (k != nn - 1) {
                  p = p + r * a[k + 2][j];
                  a[k + 2][j] = a[k + 2][j] - p * z;
                }
                a[k + 1][j] = a[k + 1][j] - p * y;
                a[k][j] = a[k][j] - p * x;
              }
              mmin = nn < k + 3 ? nn : k + 3;
              for (i = l; i <= mmin; ++i) {
                p = x * a[i][k] + y * a[i][k + 1];
                if (k != nn - 1) {
                  p = p + z * a[i][k + 2];
                  a[i][k + 2] = a[i][k + 2] - p * r;
                }
                a[i][k + 1] = a[i][k + 1] - p * q;
                a[i][k] = a[i][k] - p;
              }
            }
          }
        }
      }
    } while (l < nn - 1);
  }
  return true;
}

// null vector of the (numerically) singular 10x10 matrix A: complete-pivoting elimination of 9
// columns, free variable = 1
__device__ inline bool null_vector_10(double A[10][10], double* v) {
  int colperm[10];
  for (int j = 0; j < 10; ++j) colperm[j] = j;
  for (int r = 0; r < 9; ++r) {
    int pi = r, pj = r;
    double best = fabs(A[r][r]);
    for (int i = r; i < 10; ++i)
      for (int j = r; j < 10; ++j) {
        const double a = fabs(A[i][j]);
        if (a > best) { best = a; pi = i; pj = j; }
      }
    if (!(best > 0.0)) return false;
    if (pi != r)
      for (int j = 0; j < 10; ++j) { const double t = A[r][j]; A[r][j] = A[pi][j]; A[pi][j] = t; }
    if (pj != r) {
      for (int i = 0; i < 10; ++i) { const double t = A[i][r]; A[i][r] = A[i][pj]; A[i][pj] = t; }
      const int t = colperm[r]; colperm[r] = colperm[pj]; colperm[pj] = t;
    }
    for (int i = r + 1; i < 10; ++i) {
      const double f = A[i][r] / A[r][r];
      for (int j = r + 1; j < 10; ++j) A[i][j] = A[i][j] - f * A[r][j];
      A[i][r] = 0.0;
    }
  }
  double z[10];
  z[9] = 1.0;
  for (int r = 8; r >= 0; --r) {
    double s = 0.0;
    for (int j = r + 1; j < 10; ++j) s = s + A[r][j] * z[j];
    z[r] = -s / A[r][r];
  }
  for (int k = 0; k < 10; ++k) v[colperm[k]] = z[k];
  return true;
}

// FivePointSolver::Solve.  b1, b2: 5 bearing vectors each (row k = (x,y,z) of point k); Eout: up to 10
// row-major 3x3 essential matrices with b2^T E b1 = 0.  Returns the number of models.
__device__ inline int five_point(const double* b1, const double* b2, double* Eout) {
  double A[5][9];
  for (int i = 0; i < 5; ++i) {  // EncodeEpipolarEquation on homogeneous (3-D) points
    const double* x1 = b1 + 3 * i;
    const double* x2 = b2 + 3 * i;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) A[i][3 * r + c] = x2[r] * x1[c];
  }
  double basis[4][9];
  if (!nullspace_5x9(A, basis)) return 0;
  double M[10][20];
  constraints(basis, M);
  if (!gauss_jordan(M)) return 0;
  // action matrix of multiplication by x on [xx xy xz yy yz zz x y z 1]
  double At[10][10];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 10; ++j) At[i][j] = -M[i][10 + j];
  for (int i = 6; i < 10; ++i)
    for (int j = 0; j < 10; ++j) At[i][j] = 0.0;
  At[6][0] = 1.0;
  At[7][1] = 1.0;
  At[8][2] = 1.0;
  At[9][6] = 1.0;
  double H[11][11], wr[11], wi[11];
  for (int i = 0; i < 10; ++i)
    for (int j = 0; j < 10; ++j) H[i + 1][j + 1] = At[i][j];
  elmhes(H, 10);
  if (!hqr(H, 10, wr, wi)) return 0;
  int nm = 0;
  for (int s = 1; s <= 10; ++s) {
    if (wi[s] != 0.0) continue;
    double S[10][10], v[10];
    for (int i = 0; i < 10; ++i)
      for (int j = 0; j < 10; ++j) S[i][j] = (i == j) ? At[i][j] - wr[s] : At[i][j];
    if (!null_vector_10(S, v)) continue;
    if (v[9] == 0.0) continue;
    const double x = v[6] / v[9], y = v[7] / v[9], z = v[8] / v[9];
    double* E = Eout + 9 * nm;
    for (int k = 0; k < 9; ++k) E[k] = ((x * basis[0][k] + y * basis[1][k]) + z * basis[2][k]) + basis[3][k];
    bool finite = true;
    for (int k = 0; k < 9; ++k) finite = finite && (E[k] == E[k]) && fabs(E[k]) <= 1.7e308;
    if (!finite) continue;
    ++nm;
  }
  return nm;
}

}  // namespace fp
}  // namespace r3d
