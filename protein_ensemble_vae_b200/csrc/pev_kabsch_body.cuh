// 3x3 Kabsch rotation in double precision (K4).  Host/device.
// Reference: generate_ensemble_pdbs.py:343-373 (kabsch_rmsd), scripts/validation_metrics.py:57-85.
#pragma once
#include "pev_hd.cuh"

namespace pev {

// cyclic Jacobi on a symmetric 3x3 matrix: K = V diag(w) V^T, eigenvalues sorted descending
PEV_HD void jacobi_eig3(double K[3][3], double V[3][3], double w[3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = fabs(K[0][1]) + fabs(K[0][2]) + fabs(K[1][2]);
    double diag = fabs(K[0][0]) + fabs(K[1][1]) + fabs(K[2][2]);
    if (off <= 1e-300 || off <= 1e-17 * diag) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (fabs(K[p][q]) <= 1e-300) continue;
        double theta = (K[q][q] - K[p][p]) / (2.0 * K[p][q]);
        double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {          // K <- K J
          double kp = K[k][p], kq = K[k][q];
          K[k][p] = c * kp - s * kq;
          K[k][q] = s * kp + c * kq;
        }
        for (int k = 0; k < 3; ++k) {          // K <- J^T K
          double pk = K[p][k], qk = K[q][k];
          K[p][k] = c * pk - s * qk;
          K[q][k] = s * pk + c * qk;
        }
        for (int k = 0; k < 3; ++k) {          // V <- V J
          double vp = V[k][p], vq = V[k][q];
          V[k][p] = c * vp - s * vq;
          V[k][q] = s * vp + c * vq;
        }
      }
  }
  for (int i = 0; i < 3; ++i) w[i] = K[i][i];
  for (int i = 0; i < 2; ++i)                  // sort descending, permuting columns of V
    for (int j = 0; j < 2 - i; ++j)
      if (w[j] < w[j + 1]) {
        double tw = w[j]; w[j] = w[j + 1]; w[j + 1] = tw;
        for (int k = 0; k < 3; ++k) { double tv = V[k][j]; V[k][j] = V[k][j + 1]; V[k][j + 1] = tv; }
      }
}

// H = sum_l a_l b_l^T (centred).  Returns the proper rotation R (row-major) maximising tr(R H),
// i.e. R a ~ b; equals V diag(1,1,sign det(V U^T)) U^T of the reference (:358-366).
PEV_HD void kabsch_rotation(const double H[3][3], double R[3][3]) {
  double K[3][3], V[3][3], w[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += H[k][i] * H[k][j];      // K = H^T H
      K[i][j] = s;
    }
  jacobi_eig3(K, V, w);
  double det = V[0][0] * (V[1][1] * V[2][2] - V[1][2] * V[2][1]) -
               V[0][1] * (V[1][0] * V[2][2] - V[1][2] * V[2][0]) +
               V[0][2] * (V[1][0] * V[2][1] - V[1][1] * V[2][0]);
  if (det < 0.0)
    for (int k = 0; k < 3; ++k) V[k][2] = -V[k][2];
  double U[3][3];                              // columns u1,u2,u3; u3 = u1 x u2 (signed sigma3)
  double n1 = 0.0, n2 = 0.0;
  for (int k = 0; k < 3; ++k) {
    U[k][0] = H[k][0] * V[0][0] + H[k][1] * V[1][0] + H[k][2] * V[2][0];
    n1 += U[k][0] * U[k][0];
  }
  n1 = sqrt(n1);
  bool ident = !(n1 > 1e-150);
  if (!ident) {
    for (int k = 0; k < 3; ++k) U[k][0] /= n1;
    double d = 0.0;
    for (int k = 0; k < 3; ++k) {
      U[k][1] = H[k][0] * V[0][1] + H[k][1] * V[1][1] + H[k][2] * V[2][1];
      d += U[k][1] * U[k][0];
    }
    for (int k = 0; k < 3; ++k) { U[k][1] -= d * U[k][0]; n2 += U[k][1] * U[k][1]; }
    n2 = sqrt(n2);
    if (n2 > 1e-12 * n1) {
      for (int k = 0; k < 3; ++k) U[k][1] /= n2;
    } else {                                   // rank-1 covariance: any unit vector orthogonal to u1
      int m = fabs(U[0][0]) < fabs(U[1][0]) ? (fabs(U[0][0]) < fabs(U[2][0]) ? 0 : 2)
                                            : (fabs(U[1][0]) < fabs(U[2][0]) ? 1 : 2);
      double e[3] = {0.0, 0.0, 0.0};
      e[m] = 1.0;
      double dd = U[m][0], nn = 0.0;
      for (int k = 0; k < 3; ++k) { U[k][1] = e[k] - dd * U[k][0]; nn += U[k][1] * U[k][1]; }
      nn = sqrt(nn);
      for (int k = 0; k < 3; ++k) U[k][1] /= nn;
    }
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      if (ident) { R[i][j] = (i == j) ? 1.0 : 0.0; continue; }
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += V[i][k] * U[j][k];      // R = V U^T
      R[i][j] = s;
    }
}

// serial whole-problem body (used by hostcheck and by lane 0 for tiny L); the kernel
// parallelises the three passes over a warp and calls kabsch_rotation between them.
PEV_HD float kabsch_rmsd_serial(const float* a, const float* b, const float* mask, int L, int mode) {
  double ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
  int n = 0;
  for (int l = 0; l < L; ++l) {
    if (mask && mask[l] == 0.f) continue;
    ++n;
    for (int k = 0; k < 3; ++k) { ca[k] += a[3 * l + k]; cb[k] += b[3 * l + k]; }
  }
  if (n == 0) return 0.f;
  for (int k = 0; k < 3; ++k) { ca[k] /= n; cb[k] /= n; }
  double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int l = 0; l < L; ++l) {
    if (mask && mask[l] == 0.f) continue;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) H[i][j] += (a[3 * l + i] - ca[i]) * (b[3 * l + j] - cb[j]);
  }
  double R[3][3];
  kabsch_rotation(H, R);
  double e = 0.0;
  for (int l = 0; l < L; ++l) {
    if (mask && mask[l] == 0.f) continue;
    double p[3] = {a[3 * l] - ca[0], a[3 * l + 1] - ca[1], a[3 * l + 2] - ca[2]};
    for (int i = 0; i < 3; ++i) {
      double q = 0.0;
      for (int k = 0; k < 3; ++k) q += (mode == 0 ? R[i][k] : R[k][i]) * p[k];
      double d = q - (b[3 * l + i] - cb[i]);
      e += d * d;
    }
  }
  return (float)sqrt(e / n);
}

// ------------------------------------------------------------------------------------------
// Geometry validity filter of the generation driver (generate_ensemble_pdbs.py:290-340), one conformer: CA-CA distances
// and CA-CA-CA angles over the COMPACTED valid residues (mask gaps are bridged, as there).  Status: 0 valid, 1 no valid
// residues, 2 extreme CA-CA distance (> 6.0), 3 abnormal average CA-CA distance (< 2.5 or > 5.0), 4 abnormal average
// CA-CA-CA angle (< 60 or > 180 degrees).  stats = {max distance, mean distance, mean angle in degrees}.
PEV_HD int validate_geometry_serial(const float* ca, const float* mask, int L, float* stats) {
  int n = 0, nd = 0, na = 0;
  float p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};          // previous and the one before it
  float dmax = 0.f, dsum = 0.f, asum = 0.f;
  for (int l = 0; l < L; ++l) {
    if (mask && mask[l] == 0.f) continue;
    const float p[3] = {ca[3 * l], ca[3 * l + 1], ca[3 * l + 2]};
    if (n >= 1) {
      const float v2[3] = {p[0] - p1[0], p[1] - p1[1], p[2] - p1[2]};
      const float d = sqrtf(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);                    // :306-308
      dmax = d > dmax ? d : dmax;
      dsum += d;
      ++nd;
      if (n >= 2) {
        const float v1[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
        const float n1 = sqrtf(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
        float c = (v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2]) / (n1 * d + 1e-8f);          // :328
        c = c < -1.f ? -1.f : (c > 1.f ? 1.f : c);
        asum += acosf(c) * 57.29577951308232f;                                                 // :330
        ++na;
      }
    }
    for (int k = 0; k < 3; ++k) { p2[k] = p1[k]; p1[k] = p[k]; }
    ++n;
  }
  const float davg = nd ? dsum / nd : 0.f, aavg = na ? asum / na : 0.f;
  if (stats) { stats[0] = dmax; stats[1] = davg; stats[2] = aavg; }
  if (n == 0) return 1;                                                                        // :302-303
  if (nd) {
    if (dmax > 6.0f) return 2;                                                                 // :314-315
    if (davg < 2.5f || davg > 5.0f) return 3;                                                  // :317-318
    if (na && (aavg < 60.f || aavg > 180.f)) return 4;                                         // :334-336
  }
  return 0;
}

}  // namespace pev
