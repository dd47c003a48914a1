// Host-side pose stage: PoseEstimator::PnPSolver + PoseBA (pose_estimation.cpp:50-143) without OpenCV or
// Ceres.  Plain C++ (no CUDA): the work per marker is a 12x12 eigen problem and a 6-parameter least squares.
//
//   1. point selection                    pose_estimation.cpp:72-95   (which corners of which features take part)
//   2. undistortion, 5-coefficient model  pose_estimation.cpp:97-101  (cv::undistortPoints: 5 fixed-point iterations)
//   3. initial pose: EPnP                 pose_estimation.cpp:103     (cv::solvePnP(SOLVEPNP_EPNP): Lepetit, Moreno-Noguer
//                                                                      & Fua, IJCV 2009, restated from the paper)
//   4. refinement: Levenberg-Marquardt    pose_estimation.cpp:14-41,105-128 (Ceres, angle-axis + translation, pinhole
//                                                                      reprojection residual on the undistorted points)
//
// The refinement is run to convergence, so the result is the local minimum of the reprojection cost next to the EPnP
// estimate; the parity target against the oracle (cv2 EPnP + scipy LM) is 1e-4 rad / 1e-4 |t|.
#include <math.h>
#include <string.h>

#include <vector>

#include "../../include/ctag.h"

namespace {

// ---- small dense helpers ---------------------------------------------------------------------------------------------
// Cyclic Jacobi for a symmetric n x n matrix (row-major, n <= 12).  On return a holds the eigenvalues on its diagonal
// and the COLUMNS of v are the eigenvectors.
void jacobi_eig(double* a, double* v, int n) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) v[i * n + j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) (i == j ? diag : off) += a[i * n + j] * a[i * n + j];
    if (off <= 1e-30 * diag || off == 0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = a[p * n + q];
        if (apq == 0) continue;
        const double theta = (a[q * n + q] - a[p * n + p]) / (2 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        const double c = 1 / sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = a[k * n + p], akq = a[k * n + q];
          a[k * n + p] = c * akp - s * akq;
          a[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = a[p * n + k], aqk = a[q * n + k];
          a[p * n + k] = c * apk - s * aqk;
          a[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = v[k * n + p], vkq = v[k * n + q];
          v[k * n + p] = c * vkp - s * vkq;
          v[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
}

// Least squares min |A x - b| for a small m x n system (m >= n) through the normal equations with a pivoted
// Gauss-Jordan solve; a tiny ridge keeps rank-deficient cases (planar point sets) finite.
bool lstsq(const double* A, const double* b, int m, int n, double* x) {
  double N[6 * 7];
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int k = 0; k < m; ++k) s += A[k * n + i] * A[k * n + j];
      N[i * (n + 1) + j] = s;
    }
    double s = 0;
    for (int k = 0; k < m; ++k) s += A[k * n + i] * b[k];
    N[i * (n + 1) + n] = s;
  }
  double tr = 0;
  for (int i = 0; i < n; ++i) tr += N[i * (n + 1) + i];
  for (int i = 0; i < n; ++i) N[i * (n + 1) + i] += 1e-14 * tr;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (fabs(N[r * (n + 1) + c]) > fabs(N[piv * (n + 1) + c])) piv = r;
    if (N[piv * (n + 1) + c] == 0) return false;
    if (piv != c)
      for (int k = 0; k <= n; ++k) {
        const double t = N[c * (n + 1) + k];
        N[c * (n + 1) + k] = N[piv * (n + 1) + k];
        N[piv * (n + 1) + k] = t;
      }
    const double inv = 1 / N[c * (n + 1) + c];
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = N[r * (n + 1) + c] * inv;
      for (int k = c; k <= n; ++k) N[r * (n + 1) + k] -= f * N[c * (n + 1) + k];
    }
  }
  for (int i = 0; i < n; ++i) x[i] = N[i * (n + 1) + n] / N[i * (n + 1) + i];
  return true;
}

void rodrigues(const double* r, double* R) {
  const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (th < 1e-300) {
    R[0] = R[4] = R[8] = 1;
    R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = 0;
    return;
  }
  const double x = r[0] / th, y = r[1] / th, z = r[2] / th, c = cos(th), s = sin(th), c1 = 1 - c;
  R[0] = c + c1 * x * x, R[1] = c1 * x * y - s * z, R[2] = c1 * x * z + s * y;
  R[3] = c1 * x * y + s * z, R[4] = c + c1 * y * y, R[5] = c1 * y * z - s * x;
  R[6] = c1 * x * z - s * y, R[7] = c1 * y * z + s * x, R[8] = c + c1 * z * z;
}

// unit quaternion (w, x, y, z) -> rotation vector / matrix
void quat_to_rvec(const double* q, double* r) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  if (w < 0) w = -w, x = -x, y = -y, z = -z;
  const double vn = sqrt(x * x + y * y + z * z);
  if (vn < 1e-300) {
    r[0] = r[1] = r[2] = 0;
    return;
  }
  const double th = 2 * atan2(vn, w);
  r[0] = x / vn * th, r[1] = y / vn * th, r[2] = z / vn * th;
}
void quat_to_R(const double* q, double* R) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z), R[1] = 2 * (x * y - w * z), R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z), R[4] = 1 - 2 * (x * x + z * z), R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y), R[7] = 2 * (y * z + w * x), R[8] = 1 - 2 * (x * x + y * y);
}

// ---- 2. cv::undistortPoints(src, K, D, noArray(), K): normalise, 5 fixed-point iterations, back to pixels ------------
void undistort(const double* uv, int n, const double* K, const double* D, int nd, double* out_norm, double* out_px) {
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  const double k1 = nd > 0 ? D[0] : 0, k2 = nd > 1 ? D[1] : 0, p1 = nd > 2 ? D[2] : 0, p2 = nd > 3 ? D[3] : 0,
               k3 = nd > 4 ? D[4] : 0;
  for (int i = 0; i < n; ++i) {
    double x = (uv[2 * i] - cx) / fx, y = (uv[2 * i + 1] - cy) / fy;
    const double x0 = x, y0 = y;
    if (nd > 0) {
      for (int it = 0; it < 5; ++it) {
        const double r2 = x * x + y * y;
        const double icdist = 1 / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
        if (icdist < 0) {  // the model has folded over: keep the distorted point
          x = x0, y = y0;
          break;
        }
        const double dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x), dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
        x = (x0 - dx) * icdist;
        y = (y0 - dy) * icdist;
      }
    }
    out_norm[2 * i] = x, out_norm[2 * i + 1] = y;
    out_px[2 * i] = x * fx + cx, out_px[2 * i + 1] = y * fy + cy;
  }
}

// ---- 3. EPnP on normalised image points (focal 1, centre 0) -----------------------------------------------------------
struct Epnp {
  int n;
  const double* pw;  // [n][3] object points
  const double* xn;  // [n][2] normalised image points
  double cws[4][3];
  std::vector<double> alphas;  // [n][4]
  double V[4][12];             // eigenvectors of M^T M for the four smallest eigenvalues, V[0] the smallest
  double L[6][10], rho[6];

  void control_points() {
    double c0[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) c0[k] += pw[3 * i + k];
    for (int k = 0; k < 3; ++k) c0[k] /= n, cws[0][k] = c0[k];
    double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, E[9];
    for (int i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) C[3 * a + b] += (pw[3 * i + a] - c0[a]) * (pw[3 * i + b] - c0[b]);
    jacobi_eig(C, E, 3);
    double lmax = 0;
    for (int a = 0; a < 3; ++a) lmax = fmax(lmax, C[4 * a]);
    for (int a = 0; a < 3; ++a) {
      // a (nearly) planar point set would put a control point on the centroid: keep a small offset along the normal
      const double lam = fmax(C[4 * a], 1e-12 * lmax);
      const double k = sqrt(lam / n);
      for (int b = 0; b < 3; ++b) cws[a + 1][b] = c0[b] + k * E[3 * b + a];
    }
  }

  bool barycentric() {
    double CC[9], inv[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) CC[3 * a + b] = cws[b + 1][a] - cws[0][a];
    const double det = CC[0] * (CC[4] * CC[8] - CC[5] * CC[7]) - CC[1] * (CC[3] * CC[8] - CC[5] * CC[6]) +
                       CC[2] * (CC[3] * CC[7] - CC[4] * CC[6]);
    if (det == 0 || !isfinite(det)) return false;
    inv[0] = (CC[4] * CC[8] - CC[5] * CC[7]) / det, inv[1] = (CC[2] * CC[7] - CC[1] * CC[8]) / det;
    inv[2] = (CC[1] * CC[5] - CC[2] * CC[4]) / det, inv[3] = (CC[5] * CC[6] - CC[3] * CC[8]) / det;
    inv[4] = (CC[0] * CC[8] - CC[2] * CC[6]) / det, inv[5] = (CC[2] * CC[3] - CC[0] * CC[5]) / det;
    inv[6] = (CC[3] * CC[7] - CC[4] * CC[6]) / det, inv[7] = (CC[1] * CC[6] - CC[0] * CC[7]) / det;
    inv[8] = (CC[0] * CC[4] - CC[1] * CC[3]) / det;
    alphas.assign((size_t)4 * n, 0.0);
    for (int i = 0; i < n; ++i) {
      double d[3] = {pw[3 * i] - cws[0][0], pw[3 * i + 1] - cws[0][1], pw[3 * i + 2] - cws[0][2]};
      double* a = &alphas[4 * i];
      for (int k = 0; k < 3; ++k) a[k + 1] = inv[3 * k] * d[0] + inv[3 * k + 1] * d[1] + inv[3 * k + 2] * d[2];
      a[0] = 1 - a[1] - a[2] - a[3];
    }
    return true;
  }

  void null_space() {
    double MtM[144], E[144];
    memset(MtM, 0, sizeof(MtM));
    for (int i = 0; i < n; ++i) {
      const double* a = &alphas[4 * i];
      double r0[12], r1[12];
      for (int j = 0; j < 4; ++j) {
        r0[3 * j] = a[j], r0[3 * j + 1] = 0, r0[3 * j + 2] = -a[j] * xn[2 * i];
        r1[3 * j] = 0, r1[3 * j + 1] = a[j], r1[3 * j + 2] = -a[j] * xn[2 * i + 1];
      }
      for (int p = 0; p < 12; ++p)
        for (int q = 0; q < 12; ++q) MtM[12 * p + q] += r0[p] * r0[q] + r1[p] * r1[q];
    }
    jacobi_eig(MtM, E, 12);
    int order[12];
    for (int i = 0; i < 12; ++i) order[i] = i;
    for (int i = 0; i < 12; ++i)
      for (int j = i + 1; j < 12; ++j)
        if (MtM[13 * order[j]] < MtM[13 * order[i]]) {
          const int t = order[i];
          order[i] = order[j];
          order[j] = t;
        }
    for (int k = 0; k < 4; ++k)
      for (int p = 0; p < 12; ++p) V[k][p] = E[12 * p + order[k]];
  }

  void distance_system() {
    static const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
    for (int r = 0; r < 6; ++r) {
      double dv[4][3];
      for (int k = 0; k < 4; ++k)
        for (int c = 0; c < 3; ++c) dv[k][c] = V[k][3 * pa[r] + c] - V[k][3 * pb[r] + c];
      auto dot = [&](int a, int b) { return dv[a][0] * dv[b][0] + dv[a][1] * dv[b][1] + dv[a][2] * dv[b][2]; };
      // unknowns: b11 b12 b22 b13 b23 b33 b14 b24 b34 b44 (b_ij = beta_i beta_j)
      L[r][0] = dot(0, 0), L[r][1] = 2 * dot(0, 1), L[r][2] = dot(1, 1), L[r][3] = 2 * dot(0, 2), L[r][4] = 2 * dot(1, 2);
      L[r][5] = dot(2, 2), L[r][6] = 2 * dot(0, 3), L[r][7] = 2 * dot(1, 3), L[r][8] = 2 * dot(2, 3), L[r][9] = dot(3, 3);
      double d = 0;
      for (int c = 0; c < 3; ++c) d += (cws[pa[r]][c] - cws[pb[r]][c]) * (cws[pa[r]][c] - cws[pb[r]][c]);
      rho[r] = d;
    }
  }

  // linearised initial guesses for the betas with 4, 3 and 5 of the ten products kept
  void betas_n4(double* b) {
    double A[24], x[4];
    for (int r = 0; r < 6; ++r) A[4 * r] = L[r][0], A[4 * r + 1] = L[r][1], A[4 * r + 2] = L[r][3], A[4 * r + 3] = L[r][6];
    b[0] = b[1] = b[2] = b[3] = 0;
    if (!lstsq(A, rho, 6, 4, x) || x[0] == 0) return;
    const double s = x[0] < 0 ? -1.0 : 1.0;
    b[0] = sqrt(s * x[0]);
    b[1] = s * x[1] / b[0], b[2] = s * x[2] / b[0], b[3] = s * x[3] / b[0];
  }
  void betas_n2(double* b) {
    double A[18], x[3];
    for (int r = 0; r < 6; ++r) A[3 * r] = L[r][0], A[3 * r + 1] = L[r][1], A[3 * r + 2] = L[r][2];
    b[0] = b[1] = b[2] = b[3] = 0;
    if (!lstsq(A, rho, 6, 3, x)) return;
    if (x[0] < 0) {
      b[0] = sqrt(-x[0]);
      b[1] = x[2] < 0 ? sqrt(-x[2]) : 0;
    } else {
      b[0] = sqrt(x[0]);
      b[1] = x[2] > 0 ? sqrt(x[2]) : 0;
    }
    if (x[1] < 0) b[0] = -b[0];
  }
  void betas_n3(double* b) {
    double A[30], x[5];
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 5; ++c) A[5 * r + c] = L[r][c];
    b[0] = b[1] = b[2] = b[3] = 0;
    if (!lstsq(A, rho, 6, 5, x)) return;
    if (x[0] < 0) {
      b[0] = sqrt(-x[0]);
      b[1] = x[2] < 0 ? sqrt(-x[2]) : 0;
    } else {
      b[0] = sqrt(x[0]);
      b[1] = x[2] > 0 ? sqrt(x[2]) : 0;
    }
    if (x[1] < 0) b[0] = -b[0];
    b[2] = b[0] != 0 ? x[3] / b[0] : 0;
  }

  // Gauss-Newton on the six control-point distance equations
  void refine_betas(double* b) {
    for (int it = 0; it < 5; ++it) {
      double A[24], e[6], x[4];
      for (int r = 0; r < 6; ++r) {
        const double* l = L[r];
        A[4 * r] = 2 * l[0] * b[0] + l[1] * b[1] + l[3] * b[2] + l[6] * b[3];
        A[4 * r + 1] = l[1] * b[0] + 2 * l[2] * b[1] + l[4] * b[2] + l[7] * b[3];
        A[4 * r + 2] = l[3] * b[0] + l[4] * b[1] + 2 * l[5] * b[2] + l[8] * b[3];
        A[4 * r + 3] = l[6] * b[0] + l[7] * b[1] + l[8] * b[2] + 2 * l[9] * b[3];
        e[r] = rho[r] - (l[0] * b[0] * b[0] + l[1] * b[0] * b[1] + l[2] * b[1] * b[1] + l[3] * b[0] * b[2] + l[4] * b[1] * b[2] +
                         l[5] * b[2] * b[2] + l[6] * b[0] * b[3] + l[7] * b[1] * b[3] + l[8] * b[2] * b[3] + l[9] * b[3] * b[3]);
      }
      if (!lstsq(A, e, 6, 4, x)) return;
      for (int k = 0; k < 4; ++k) b[k] += x[k];
    }
  }

  // camera-frame control points from the betas, absolute orientation (Horn's quaternion), mean reprojection error
  double pose_from_betas(const double* b, double* quat, double* t) {
    double ccs[4][3];
    for (int a = 0; a < 4; ++a)
      for (int c = 0; c < 3; ++c) ccs[a][c] = b[0] * V[0][3 * a + c] + b[1] * V[1][3 * a + c] + b[2] * V[2][3 * a + c] + b[3] * V[3][3 * a + c];
    std::vector<double> pc((size_t)3 * n);
    for (int i = 0; i < n; ++i)
      for (int c = 0; c < 3; ++c)
        pc[3 * i + c] = alphas[4 * i] * ccs[0][c] + alphas[4 * i + 1] * ccs[1][c] + alphas[4 * i + 2] * ccs[2][c] + alphas[4 * i + 3] * ccs[3][c];
    if (pc[2] < 0)  // points have to be in front of the camera
      for (double& v : pc) v = -v;
    double mc[3] = {0, 0, 0}, mw[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
      for (int c = 0; c < 3; ++c) mc[c] += pc[3 * i + c], mw[c] += pw[3 * i + c];
    for (int c = 0; c < 3; ++c) mc[c] /= n, mw[c] /= n;
    double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // S[a][b] = sum w_a c_b
    for (int i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a)
        for (int c = 0; c < 3; ++c) S[3 * a + c] += (pw[3 * i + a] - mw[a]) * (pc[3 * i + c] - mc[c]);
    double N[16] = {S[0] + S[4] + S[8], S[5] - S[7],        S[6] - S[2],         S[1] - S[3],
                    S[5] - S[7],        S[0] - S[4] - S[8], S[1] + S[3],         S[6] + S[2],
                    S[6] - S[2],        S[1] + S[3],        -S[0] + S[4] - S[8], S[5] + S[7],
                    S[1] - S[3],        S[6] + S[2],        S[5] + S[7],         -S[0] - S[4] + S[8]};
    double E[16];
    jacobi_eig(N, E, 4);
    int best = 0;
    for (int k = 1; k < 4; ++k)
      if (N[5 * k] > N[5 * best]) best = k;
    for (int k = 0; k < 4; ++k) quat[k] = E[4 * k + best];
    double R[9];
    quat_to_R(quat, R);
    for (int a = 0; a < 3; ++a) t[a] = mc[a] - (R[3 * a] * mw[0] + R[3 * a + 1] * mw[1] + R[3 * a + 2] * mw[2]);
    double err = 0;
    for (int i = 0; i < n; ++i) {
      double p[3];
      for (int a = 0; a < 3; ++a) p[a] = R[3 * a] * pw[3 * i] + R[3 * a + 1] * pw[3 * i + 1] + R[3 * a + 2] * pw[3 * i + 2] + t[a];
      const double du = xn[2 * i] - p[0] / p[2], dv = xn[2 * i + 1] - p[1] / p[2];
      err += sqrt(du * du + dv * dv);
    }
    return isfinite(err) ? err / n : 1e300;
  }

  bool solve(double* rvec, double* tvec) {
    control_points();
    if (!barycentric()) return false;
    null_space();
    distance_system();
    double best = 1e300, bq[4] = {1, 0, 0, 0}, bt[3] = {0, 0, 1};
    for (int variant = 0; variant < 3; ++variant) {
      double b[4], q[4], t[3];
      if (variant == 0) betas_n4(b);
      else if (variant == 1) betas_n2(b);
      else betas_n3(b);
      refine_betas(b);
      const double e = pose_from_betas(b, q, t);
      if (e < best) {
        best = e;
        memcpy(bq, q, sizeof(bq));
        memcpy(bt, t, sizeof(bt));
      }
    }
    if (best >= 1e300) return false;
    quat_to_rvec(bq, rvec);
    memcpy(tvec, bt, sizeof(bt));
    return true;
  }
};

// ---- 4. Levenberg-Marquardt on (rvec, tvec), residual = K-projection - undistorted pixel ------------------------------
struct Reproj {
  int n;
  const double* pw;
  const double* px;  // undistorted pixels
  double fx, fy, cx, cy;
  void operator()(const double* p, double* r) const {
    double R[9];
    rodrigues(p, R);
    for (int i = 0; i < n; ++i) {
      const double* X = pw + 3 * i;
      const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + p[3], y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + p[4],
                   z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + p[5];
      r[2 * i] = fx * x / z + cx - px[2 * i];
      r[2 * i + 1] = fy * y / z + cy - px[2 * i + 1];
    }
  }
};

double sumsq(const std::vector<double>& r) {
  double s = 0;
  for (double v : r) s += v * v;
  return s;
}

void levenberg_marquardt(const Reproj& f, double* p) {
  const int m = 2 * f.n;
  std::vector<double> r(m), rp(m), rm(m), J((size_t)m * 6);
  f(p, r.data());
  double cost = sumsq(r), lambda = -1, nu = 2;
  for (int it = 0; it < 200; ++it) {
    for (int k = 0; k < 6; ++k) {  // central differences
      const double h = 1e-6 * fmax(1.0, fabs(p[k]));
      double q[6];
      memcpy(q, p, sizeof(q));
      q[k] = p[k] + h;
      f(q, rp.data());
      q[k] = p[k] - h;
      f(q, rm.data());
      for (int i = 0; i < m; ++i) J[(size_t)i * 6 + k] = (rp[i] - rm[i]) / (2 * h);
    }
    double A[36], g[6];
    for (int a = 0; a < 6; ++a) {
      for (int b = 0; b < 6; ++b) {
        double s = 0;
        for (int i = 0; i < m; ++i) s += J[(size_t)i * 6 + a] * J[(size_t)i * 6 + b];
        A[6 * a + b] = s;
      }
      double s = 0;
      for (int i = 0; i < m; ++i) s += J[(size_t)i * 6 + a] * r[i];
      g[a] = s;
    }
    if (lambda < 0) {
      double dmax = 0;
      for (int a = 0; a < 6; ++a) dmax = fmax(dmax, A[7 * a]);
      lambda = 1e-3 * dmax;
    }
    double gmax = 0;
    for (int a = 0; a < 6; ++a) gmax = fmax(gmax, fabs(g[a]));
    if (gmax < 1e-13) break;
    bool stepped = false, tiny = false;
    for (int tries = 0; tries < 40 && !stepped; ++tries) {
      double M[6 * 7], x[6];
      for (int a = 0; a < 6; ++a) {
        for (int b = 0; b < 6; ++b) M[7 * a + b] = A[6 * a + b] + (a == b ? lambda : 0.0);
        M[7 * a + 6] = -g[a];
      }
      bool ok = true;
      for (int c = 0; c < 6 && ok; ++c) {
        int piv = c;
        for (int rr = c + 1; rr < 6; ++rr)
          if (fabs(M[7 * rr + c]) > fabs(M[7 * piv + c])) piv = rr;
        if (M[7 * piv + c] == 0) {
          ok = false;
          break;
        }
        if (piv != c)
          for (int k = 0; k < 7; ++k) {
            const double t = M[7 * c + k];
            M[7 * c + k] = M[7 * piv + k];
            M[7 * piv + k] = t;
          }
        for (int rr = 0; rr < 6; ++rr) {
          if (rr == c) continue;
          const double fct = M[7 * rr + c] / M[7 * c + c];
          for (int k = c; k < 7; ++k) M[7 * rr + k] -= fct * M[7 * c + k];
        }
      }
      if (!ok) {
        lambda *= nu, nu *= 2;
        continue;
      }
      double q[6], xn = 0, pn = 0, pred = 0;
      for (int a = 0; a < 6; ++a) {
        x[a] = M[7 * a + 6] / M[7 * a + a];
        q[a] = p[a] + x[a];
        xn += x[a] * x[a], pn += p[a] * p[a];
        pred += x[a] * (lambda * x[a] - g[a]);
      }
      if (sqrt(xn) <= 1e-14 * (sqrt(pn) + 1e-14)) {
        tiny = true;
        break;
      }
      f(q, rp.data());
      const double c2 = sumsq(rp);
      if (isfinite(c2) && c2 < cost && pred > 0) {
        const double rho = (cost - c2) / pred;
        memcpy(p, q, sizeof(q));
        r.swap(rp);
        cost = c2;
        const double t = 2 * rho - 1;
        lambda *= fmax(1.0 / 3, 1 - t * t * t);
        nu = 2;
        stepped = true;
      } else {
        lambda *= nu, nu *= 2;
      }
    }
    if (tiny || !stepped) break;
  }
}

}  // namespace

extern "C" {

// pose_estimation.cpp:72-95: which corners of which features enter the PnP problem.  A feature whose two IDs disagree
// by more than one (or whose right ID is missing) is dropped at either end of a marker with more than three features;
// the inner corners 2,3,6,7 are only used when the IDs agree to within two.
int ctag_pose_select_points(const ctag_marker* mk, int* feature_of_point, int* corner_of_point, int cap) {
  if (!mk || !feature_of_point || !corner_of_point) return CTAG_ERR_ARG;
  const int n = mk->n_features;
  int np = 0;
  for (int j = 0; j < n && j < CTAG_MAX_FEATURES; ++j) {
    const int l = mk->id_left[j], r = mk->id_right[j];
    const int diff = l > r ? l - r : r - l;
    const bool bad = diff > 1 || r == -1;
    if (n > 3 && (j == 0 || j == n - 1) && bad) continue;
    static const int outer[4] = {0, 1, 4, 5}, inner[4] = {2, 3, 6, 7};
    for (int k = 0; k < 4; ++k)
      if (np < cap) feature_of_point[np] = j, corner_of_point[np] = outer[k], ++np;
    if (diff < 3 && r != -1)
      for (int k = 0; k < 4; ++k)
        if (np < cap) feature_of_point[np] = j, corner_of_point[np] = inner[k], ++np;
  }
  return np;
}

int ctag_estimate_pose(const ctag_marker* mk, const float* model_corners, int n_model_corners, const float* intrinsic,
                       const float* dist, int n_dist, double* rvec, double* tvec, double* rms_px) {
  if (!mk || !model_corners || !intrinsic || !rvec || !tvec || (n_dist > 0 && !dist)) return CTAG_ERR_ARG;
  int fidx[8 * CTAG_MAX_FEATURES], cidx[8 * CTAG_MAX_FEATURES];
  const int n = ctag_pose_select_points(mk, fidx, cidx, 8 * CTAG_MAX_FEATURES);
  if (n < 4) return CTAG_ERR_ARG;
  std::vector<double> uv((size_t)2 * n), pw((size_t)3 * n), xn((size_t)2 * n), px((size_t)2 * n);
  for (int i = 0; i < n; ++i) {
    const int j = fidx[i], k = cidx[i], id = mk->feature_pos[j] * 8 + k;
    if (id < 0 || id >= n_model_corners) return CTAG_ERR_ARG;
    uv[2 * i] = mk->corners[j][k][0], uv[2 * i + 1] = mk->corners[j][k][1];
    for (int c = 0; c < 3; ++c) pw[3 * i + c] = model_corners[3 * id + c];
  }
  double K[9], D[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < 9; ++i) K[i] = intrinsic[i];
  const int nd = n_dist < 5 ? (n_dist < 0 ? 0 : n_dist) : 5;
  for (int i = 0; i < nd; ++i) D[i] = dist[i];
  undistort(uv.data(), n, K, D, nd, xn.data(), px.data());
  Epnp ep;
  ep.n = n;
  ep.pw = pw.data();
  ep.xn = xn.data();
  double p[6];
  if (!ep.solve(p, p + 3)) return CTAG_ERR_ARG;
  Reproj f{n, pw.data(), px.data(), K[0], K[4], K[2], K[5]};
  levenberg_marquardt(f, p);
  for (int k = 0; k < 3; ++k) rvec[k] = p[k], tvec[k] = p[3 + k];
  if (rms_px) {
    std::vector<double> r((size_t)2 * n);
    f(p, r.data());
    double s = 0;
    for (double v : r) s += v * v;
    *rms_px = sqrt(s / n);
  }
  return CTAG_OK;
}

}  // extern "C"
