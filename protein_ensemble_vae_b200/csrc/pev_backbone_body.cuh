// N / C placement and the 3-step peptide pull of EGNNDecoder.forward (models/en_gnn_decoder.py:260-310) for one residue of
// the packed batch, forward and backward -- the per-thread body of pev_backbone_fwd / pev_backbone_bwd (host/device).
//   x_n0 = x_ca + 1.46 normalize(n_dir),  x_c = x_ca + 1.52 normalize(c_dir)          (F.normalize: v / max(|v|, 1e-12))
//   3 x:  x_n[k] = a + (x_n[k] - a) clamp(1 + 0.15 (1.33 / (|x_n[k] - a| + 1e-8) - 1), 0.9, 1.1),  a = x_c[k-1]
//   (residues that start a conformer are not pulled).  A residue's result needs only x_c of its predecessor, so both
//   directions are one independent thread per residue; the backward thread of residue k also re-derives the pull of
//   residue k+1 to collect the gradient that flows into ITS x_c as the anchor.
#pragma once
#include "pev_hd.cuh"

namespace pev {

constexpr float kNCa = 1.46f, kCaC = 1.52f, kPep = 1.33f;

PEV_HD v3 unit_dir(v3 d) {
  const float n = norm(d);
  return d * (1.0f / fmaxf(n, 1e-12f));
}
// gradient of s * normalize(d) with respect to d, given g = dL/d(output)
PEV_HD v3 unit_dir_bwd(v3 d, v3 g, float s) {
  const float n = norm(d);
  if (n < 1e-12f) return g * (s / 1e-12f);                 // clamped branch: output = d / eps
  const v3 u = d * (1.0f / n);
  return (g - u * dot(g, u)) * (s / n);
}

PEV_HD float pull_scale(float dist) { return fminf(fmaxf(1.0f + 0.15f * (kPep / (dist + 1e-8f) - 1.0f), 0.90f), 1.10f); }

PEV_HD v3 pull3(v3 tail, v3 a) {
  for (int it = 0; it < 3; ++it) {
    const v3 vec = tail - a;
    tail = a + vec * pull_scale(norm(vec));
  }
  return tail;
}
// backward of pull3: given g = dL/d(tail_3), returns dL/d(tail_0) and adds dL/d(a) to *ga
PEV_HD v3 pull3_bwd(v3 tail0, v3 a, v3 g, v3* ga) {
  v3 t[3];
  t[0] = tail0;
  for (int it = 0; it < 2; ++it) {
    const v3 vec = t[it] - a;
    t[it + 1] = a + vec * pull_scale(norm(vec));
  }
  v3 gacc = zero3();
  for (int it = 2; it >= 0; --it) {
    const v3 vec = t[it] - a;
    const float d = norm(vec);
    const float raw = 1.0f + 0.15f * (kPep / (d + 1e-8f) - 1.0f);
    const float s = fminf(fmaxf(raw, 0.90f), 1.10f);
    // out = a + vec * s(d):  d out / d vec = s I + (ds/dd) vec vec^T / d  (ds/dd = 0 where the clamp is active)
    v3 gvec = g * s;
    if (raw > 0.90f && raw < 1.10f && d > 0.f) {
      const float dsdd = -0.15f * kPep / ((d + 1e-8f) * (d + 1e-8f));
      gvec += vec * (dsdd * dot(g, vec) / d);
    }
    gacc += g - gvec;                                       // a enters directly and through vec = t - a
    g = gvec;
  }
  *ga += gacc;
  return g;
}

// forward of residue k
PEV_HD void backbone_fwd_residue(const float* n_dir, int ldn, const float* c_dir, int ldc, const float* x_ca, const uint8_t* starts,
                                 int64_t N, int64_t k, float* x_n, float* x_c) {
  const v3 ca = ld3(x_ca + 3 * k);
  const v3 xc = ca + unit_dir(ld3(c_dir + k * ldc)) * kCaC;
  v3 xn = ca + unit_dir(ld3(n_dir + k * ldn)) * kNCa;
  if (k > 0 && !starts[k]) {
    const v3 a = ld3(x_ca + 3 * (k - 1)) + unit_dir(ld3(c_dir + (k - 1) * ldc)) * kCaC;      // x_c of the predecessor
    xn = pull3(xn, a);
  }
  st3(x_n + 3 * k, xn);
  st3(x_c + 3 * k, xc);
}

// backward of residue k: gradients with respect to n_dir[k], c_dir[k], x_ca[k]
PEV_HD void backbone_bwd_residue(const float* n_dir, int ldn, const float* c_dir, int ldc, const float* x_ca, const uint8_t* starts,
                                 int64_t N, int64_t k, const float* g_xn, const float* g_xc, float* g_ndir, float* g_cdir,
                                 float* g_xca) {
  const v3 ca = ld3(x_ca + 3 * k);
  const v3 nd = ld3(n_dir + k * ldn), cd = ld3(c_dir + k * ldc);
  const v3 xc = ca + unit_dir(cd) * kCaC;
  v3 gn = ld3(g_xn + 3 * k);
  v3 gc = ld3(g_xc + 3 * k);
  if (k > 0 && !starts[k]) {                                // own pull: gradient back to the un-pulled N position
    const v3 a = ld3(x_ca + 3 * (k - 1)) + unit_dir(ld3(c_dir + (k - 1) * ldc)) * kCaC;
    v3 unused = zero3();
    gn = pull3_bwd(ca + unit_dir(nd) * kNCa, a, gn, &unused);
  }
  if (k + 1 < N && !starts[k + 1]) {                        // the successor's pull uses this residue's C as its anchor
    const v3 ca1 = ld3(x_ca + 3 * (k + 1));
    pull3_bwd(ca1 + unit_dir(ld3(n_dir + (k + 1) * ldn)) * kNCa, xc, ld3(g_xn + 3 * (k + 1)), &gc);
  }
  st3(g_xca + 3 * k, gn + gc);
  st3(g_ndir + 3 * k, unit_dir_bwd(nd, gn, kNCa));
  st3(g_cdir + 3 * k, unit_dir_bwd(cd, gc, kCaC));
}

}  // namespace pev
