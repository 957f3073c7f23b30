// Evaluation metrics of scripts/validation_metrics.py (TM-score :23-54, kabsch_align :57-85, lDDT :92-149, GDT :156-199,
// RMSF :206-241), per-item bodies shared by the CUDA kernels (csrc/metrics_kernels.cu) and tests/hostcheck.  Host/device.
#pragma once
#include <math.h>

#include "pev_kabsch_body.cuh"

namespace pev {

// kabsch_align (:57-85): superpose `a` onto `b` over ALL L residues (the reference passes no mask here); writes the
// aligned coordinates (may be null) and the per-residue distances |aligned_l - b_l| (may be null)
PEV_HD void superpose_serial(const float* a, const float* b, int L, float* aligned, float* dist) {
  double ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
  for (int l = 0; l < L; ++l)
    for (int k = 0; k < 3; ++k) { ca[k] += a[3 * l + k]; cb[k] += b[3 * l + k]; }
  for (int k = 0; k < 3; ++k) { ca[k] /= L; cb[k] /= L; }
  double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int l = 0; l < L; ++l)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) H[i][j] += (a[3 * l + i] - ca[i]) * (b[3 * l + j] - cb[j]);
  double R[3][3];
  kabsch_rotation(H, R);
  for (int l = 0; l < L; ++l) {
    const double p[3] = {a[3 * l] - ca[0], a[3 * l + 1] - ca[1], a[3 * l + 2] - ca[2]};
    double e = 0.0;
    for (int i = 0; i < 3; ++i) {
      const double q = R[i][0] * p[0] + R[i][1] * p[1] + R[i][2] * p[2] + cb[i];
      if (aligned) aligned[3 * l + i] = (float)q;
      const double d = q - b[3 * l + i];
      e += d * d;
    }
    if (dist) dist[l] = (float)sqrt(e);
  }
}

// TM-score (:42-52, over all L residues, d0 = 1.24 cbrt(L - 15) - 1.8) and GDT-TS / GDT-HA (:179-197, over the masked
// residues, in percent) from the per-residue distances after superposition
PEV_HD void superposition_scores(const float* dist, const float* mask, int L, float* tm, float* gdt_ts, float* gdt_ha) {
  const double d0 = 1.24 * cbrt((double)L - 15.0) - 1.8;
  double t = 0.0;
  int n = 0, c05 = 0, c1 = 0, c2 = 0, c4 = 0, c8 = 0;
  for (int l = 0; l < L; ++l) {
    const double d = dist[l];
    t += 1.0 / (1.0 + (d / d0) * (d / d0));
    if (mask && mask[l] == 0.f) continue;
    ++n;
    c05 += d < 0.5; c1 += d < 1.0; c2 += d < 2.0; c4 += d < 4.0; c8 += d < 8.0;
  }
  *tm = (float)(t / L);
  *gdt_ts = n ? (float)(25.0 * (c1 + c2 + c4 + c8) / n) : 0.f;
  *gdt_ha = n ? (float)(25.0 * (c05 + c1 + c2 + c4) / n) : 0.f;
}

// lDDT of residue i (:121-143): neighbours j with 0 < |t_i - t_j| < cutoff (and mask_j), fractions of preserved
// distances at 0.5 / 1 / 2 / 4 A; 0 for a masked residue or one without neighbours
PEV_HD float lddt_residue(const float* pred, const float* tru, const float* mask, int L, int i, float cutoff) {
  if (mask && mask[i] == 0.f) return 0.f;
  int nn = 0, cnt = 0;
  for (int j = 0; j < L; ++j) {
    if (mask && mask[j] == 0.f) continue;
    const float tx = tru[3 * i] - tru[3 * j], ty = tru[3 * i + 1] - tru[3 * j + 1], tz = tru[3 * i + 2] - tru[3 * j + 2];
    const float dt = sqrtf(tx * tx + ty * ty + tz * tz);
    if (!(dt < cutoff) || !(dt > 0.f)) continue;
    const float px = pred[3 * i] - pred[3 * j], py = pred[3 * i + 1] - pred[3 * j + 1], pz = pred[3 * i + 2] - pred[3 * j + 2];
    const float dd = fabsf(dt - sqrtf(px * px + py * py + pz * pz));
    ++nn;
    cnt += (dd < 0.5f) + (dd < 1.0f) + (dd < 2.0f) + (dd < 4.0f);
  }
  return nn ? (float)cnt / (4.0f * nn) : 0.f;
}

// RMSF of residue l (:235-239) over an ensemble already aligned to its first member: sqrt(mean_n |x_n - mean|^2)
PEV_HD float rmsf_residue(const float* aligned, int N, int L, int l) {
  double m[3] = {0, 0, 0};
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 3; ++k) m[k] += aligned[((int64_t)n * L + l) * 3 + k];
  for (int k = 0; k < 3; ++k) m[k] /= N;
  double s = 0.0;
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 3; ++k) {
      const double d = aligned[((int64_t)n * L + l) * 3 + k] - m[k];
      s += d * d;
    }
  return (float)sqrt(s / N);
}

}  // namespace pev
