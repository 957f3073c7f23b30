// Per-thread bodies of the loss kernels (K3).  Host/device: loss_kernels.cu wraps them in
// __global__ launches, tests/hostcheck.cpp loops over them on the CPU.
// Reference: models/losses.py (line numbers cited per function).
#pragma once
#include "../../include/pev_b200.h"
#include "pev_hd.cuh"

namespace pev {

// per-thread forward accumulator slots
enum {
  RA_REC_CA = 0, RA_REC_N, RA_REC_C, RA_MSUM,            // per conformer
  RA_NCA, RA_CAC, RA_CN, RA_PM, RA_NCAC, RA_CNCA, RA_CACN,
  RA_CONS_NUM, RA_CONS_DEN, RA_RAMA, RA_OMEGA, RA_CE,
  RA_COUNT
};

PEV_HD v3 pt(const float* a, const pev_loss_args& A, int b, int i) {
  return ld3(a + ((int64_t)b * A.L + i) * 3);
}
PEV_HD float mk(const pev_loss_args& A, int b, int i) { return A.mask[(int64_t)b * A.L + i]; }

// (sin,cos) x (phi,psi,omega) slots of residue i plus the three torsions' saved state.
struct ResDih {
  float slot[6];
  Dihedral phi, psi, om;
  bool has_prev, has_next;      // pair masks m[i-1]&m[i], m[i]&m[i+1]  (models/losses.py:269,285,301)
};

PEV_HD void residue_dihedrals(const pev_loss_args& A, const float* N, const float* CA, const float* C,
                              int b, int i, ResDih& r) {
  for (int k = 0; k < 6; ++k) r.slot[k] = 0.f;
  bool mi = mk(A, b, i) != 0.f;
  r.has_prev = (i >= 1) && mi && (mk(A, b, i - 1) != 0.f);
  r.has_next = (i + 1 < A.L) && mi && (mk(A, b, i + 1) != 0.f);
  v3 n = pt(N, A, b, i), ca = pt(CA, A, b, i), c = pt(C, A, b, i);
  if (i >= 1) {
    v3 cm = pt(C, A, b, i - 1), cam = pt(CA, A, b, i - 1);
    r.phi = dihedral_fwd(cm, n, ca, c);           // C(i-1) N(i) CA(i) C(i)     :263-266
    r.om = dihedral_fwd(cam, cm, n, ca);          // CA(i-1) C(i-1) N(i) CA(i)  :295-298
    if (r.has_prev) {
      r.slot[0] = r.phi.s; r.slot[1] = r.phi.c;
      r.slot[4] = r.om.s;  r.slot[5] = r.om.c;
    }
  }
  if (i + 1 < A.L) {
    v3 np = pt(N, A, b, i + 1);
    r.psi = dihedral_fwd(n, ca, c, np);           // N(i) CA(i) C(i) N(i+1)     :279-282
    if (r.has_next) { r.slot[2] = r.psi.s; r.slot[3] = r.psi.c; }
  }
}

// ------------------------------------------------------------------------------------------ fwd
PEV_HD void residue_fwd(const pev_loss_args& A, int b, int i, float* acc) {
  const int64_t bi = (int64_t)b * A.L + i;
  const float m = A.mask[bi];
  acc[RA_MSUM] += m;
  if (A.pred_CA && A.target_CA) {
    v3 d = pt(A.pred_CA, A, b, i) - pt(A.target_CA, A, b, i);
    acc[RA_REC_CA] += dot(d, d) * m;                                   // :18-19
  }
  if (A.pred_N && A.target_N) {
    v3 d = pt(A.pred_N, A, b, i) - pt(A.target_N, A, b, i);
    acc[RA_REC_N] += dot(d, d) * m;
  }
  if (A.pred_C && A.target_C) {
    v3 d = pt(A.pred_C, A, b, i) - pt(A.target_C, A, b, i);
    acc[RA_REC_C] += dot(d, d) * m;
  }
  if (A.enable_geometry) {
    v3 n = pt(A.pred_N, A, b, i), ca = pt(A.pred_CA, A, b, i), c = pt(A.pred_C, A, b, i);
    acc[RA_NCA] += huberf_(norm(ca - n) - 1.46f, 0.02f) * m;           // :335-337
    acc[RA_CAC] += huberf_(norm(c - ca) - 1.52f, 0.02f) * m;           // :340-342
    const float D2R = 3.14159265358979323846f / 180.0f;
    acc[RA_NCAC] += huberf_(angle_fwd(n, ca, c).theta - 110.0f * D2R, 0.1f) * m;     // :382-385
    if (i + 1 < A.L) {
      float pm = m * A.mask[bi + 1];
      v3 np = pt(A.pred_N, A, b, i + 1), cap = pt(A.pred_CA, A, b, i + 1);
      acc[RA_PM] += pm;
      acc[RA_CN] += huberf_(norm(np - c) - 1.33f, 0.01f) * pm;         // :346-351
      acc[RA_CNCA] += huberf_(angle_fwd(c, np, cap).theta - 121.0f * D2R, 0.1f) * pm;   // :389-393
      acc[RA_CACN] += huberf_(angle_fwd(ca, c, np).theta - 116.0f * D2R, 0.1f) * pm;    // :399-403
    }
    ResDih r;
    residue_dihedrals(A, A.pred_N, A.pred_CA, A.pred_C, b, i, r);
    if (A.target_dih) {
      const float* t = A.target_dih + bi * 6;
      for (int k = 0; k < 6; ++k) {
        bool valid = (m != 0.f) && finitef_(r.slot[k]) && finitef_(t[k]);   // :64-66
        if (valid) {
          float d = r.slot[k] - t[k];
          acc[RA_CONS_NUM] += d * d;
          acc[RA_CONS_DEN] += 1.f;
        }
      }
    }
    acc[RA_RAMA] += rama_penalty(r.slot[0], r.slot[1], r.slot[2], r.slot[3], nullptr, nullptr) * m;
    acc[RA_OMEGA] += omega_penalty(r.slot[4], r.slot[5], nullptr) * m;
  }
  if (A.logits) {
    const float* lg = A.logits + bi * A.C;
    float mx = lg[0];
    for (int k = 1; k < A.C; ++k) mx = fmaxf(mx, lg[k]);
    float se = 0.f;
    for (int k = 0; k < A.C; ++k) se += expf(lg[k] - mx);
    // A label outside [0, C) is IGNORED: zero loss and zero gradient for that residue, as F.cross_entropy does for
    // its default ignore_index = -100 (which the reference's call at models/losses.py:431 inherits).  Other
    // out-of-range labels make PyTorch raise a device assert; here they are ignored too rather than read out of bounds.
    const int64_t lab = A.labels[bi];
    if (lab >= 0 && lab < A.C) {
      float ce = logf(se) + mx - lg[lab];
      acc[RA_CE] += ce * m;                                            // :431-434
    }
  }
}

// ------------------------------------------------------------------------------------------ bwd
// gradient pieces of the geometry terms "owned" by residue o (see DESIGN.md, K3a backward)
struct Grad7 {
  v3 CAm, Cm, N, CA, C, Np, CAp;
};

// cf[t] = coef[t] * inv_den[t] for the terms with a global denominator
PEV_HD void owner_geometry_grad(const pev_loss_args& A, const float* cf, int b, int o, Grad7& g) {
  g.CAm = g.Cm = g.N = g.CA = g.C = g.Np = g.CAp = zero3();
  const int64_t bo = (int64_t)b * A.L + o;
  const float m = A.mask[bo];
  const float D2R = 3.14159265358979323846f / 180.0f;
  v3 n = pt(A.pred_N, A, b, o), ca = pt(A.pred_CA, A, b, o), c = pt(A.pred_C, A, b, o);
  {  // N-CA, CA-C bonds and the N-CA-C angle
    v3 d = ca - n;
    float len = norm(d);
    if (len > 0.f) {
      v3 gd = d * (cf[PEV_T_BOND_NCA] * m * huber_gradf_(len - 1.46f, 0.02f) / len);
      g.CA += gd; g.N -= gd;
    }
    d = c - ca;
    len = norm(d);
    if (len > 0.f) {
      v3 gd = d * (cf[PEV_T_BOND_CAC] * m * huber_gradf_(len - 1.52f, 0.02f) / len);
      g.C += gd; g.CA -= gd;
    }
    Angle an = angle_fwd(n, ca, c);
    v3 ga, gb, gc;
    angle_bwd(an, cf[PEV_T_ANG_NCAC] * m * huber_gradf_(an.theta - 110.0f * D2R, 0.1f), ga, gb, gc);
    g.N += ga; g.CA += gb; g.C += gc;
  }
  if (o + 1 < A.L) {
    float pm = m * A.mask[bo + 1];
    v3 np = pt(A.pred_N, A, b, o + 1), cap = pt(A.pred_CA, A, b, o + 1);
    v3 d = np - c;
    float len = norm(d);
    if (len > 0.f) {
      v3 gd = d * (cf[PEV_T_BOND_CN] * pm * huber_gradf_(len - 1.33f, 0.01f) / len);
      g.Np += gd; g.C -= gd;
    }
    v3 ga, gb, gc;
    Angle a1 = angle_fwd(c, np, cap);
    angle_bwd(a1, cf[PEV_T_ANG_CNCA] * pm * huber_gradf_(a1.theta - 121.0f * D2R, 0.1f), ga, gb, gc);
    g.C += ga; g.Np += gb; g.CAp += gc;
    Angle a2 = angle_fwd(ca, c, np);
    angle_bwd(a2, cf[PEV_T_ANG_CACN] * pm * huber_gradf_(a2.theta - 116.0f * D2R, 0.1f), ga, gb, gc);
    g.CA += ga; g.C += gb; g.Np += gc;
  }
  // dihedral-derived terms of residue o
  ResDih r;
  residue_dihedrals(A, A.pred_N, A.pred_CA, A.pred_C, b, o, r);
  float gslot[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (A.target_dih) {
    const float* t = A.target_dih + bo * 6;
    for (int k = 0; k < 6; ++k)
      if ((m != 0.f) && finitef_(r.slot[k]) && finitef_(t[k]))
        gslot[k] += cf[PEV_T_DIH_CONS] * 2.0f * (r.slot[k] - t[k]);
  }
  {
    float dphi, dpsi, a, bq;
    rama_penalty(r.slot[0], r.slot[1], r.slot[2], r.slot[3], &dphi, &dpsi);
    atan2_grad(r.slot[0], r.slot[1], cf[PEV_T_RAMA] * m * dphi, &a, &bq);
    gslot[0] += a; gslot[1] += bq;
    atan2_grad(r.slot[2], r.slot[3], cf[PEV_T_RAMA] * m * dpsi, &a, &bq);
    gslot[2] += a; gslot[3] += bq;
    float dom;
    omega_penalty(r.slot[4], r.slot[5], &dom);
    atan2_grad(r.slot[4], r.slot[5], cf[PEV_T_OMEGA] * m * dom, &a, &bq);
    gslot[4] += a; gslot[5] += bq;
  }
  v3 q0, q1, q2, q3;
  if (r.has_prev) {
    dihedral_bwd(r.phi, gslot[0], gslot[1], q0, q1, q2, q3);   // C(o-1) N CA C
    g.Cm += q0; g.N += q1; g.CA += q2; g.C += q3;
    dihedral_bwd(r.om, gslot[4], gslot[5], q0, q1, q2, q3);    // CA(o-1) C(o-1) N CA
    g.CAm += q0; g.Cm += q1; g.N += q2; g.CA += q3;
  }
  if (r.has_next) {
    dihedral_bwd(r.psi, gslot[2], gslot[3], q0, q1, q2, q3);   // N CA C N(o+1)
    g.N += q0; g.CA += q1; g.C += q2; g.Np += q3;
  }
}

// gradients landing on residue i's three atoms.  cf as above; cf_rec = coef[REC_*] * inv_den_sample[b].
PEV_HD void residue_bwd(const pev_loss_args& A, const float* cf, const float* cf_rec, int b, int i,
                        v3& gN, v3& gCA, v3& gC) {
  gN = gCA = gC = zero3();
  const int64_t bi = (int64_t)b * A.L + i;
  const float m = A.mask[bi];
  if (A.pred_CA && A.target_CA)
    gCA += (pt(A.pred_CA, A, b, i) - pt(A.target_CA, A, b, i)) * (2.0f * m * cf_rec[0]);
  if (A.pred_N && A.target_N)
    gN += (pt(A.pred_N, A, b, i) - pt(A.target_N, A, b, i)) * (2.0f * m * cf_rec[1]);
  if (A.pred_C && A.target_C)
    gC += (pt(A.pred_C, A, b, i) - pt(A.target_C, A, b, i)) * (2.0f * m * cf_rec[2]);
  if (A.enable_geometry) {
    Grad7 g;
    if (i >= 1) {
      owner_geometry_grad(A, cf, b, i - 1, g);
      gN += g.Np; gCA += g.CAp;
    }
    owner_geometry_grad(A, cf, b, i, g);
    gN += g.N; gCA += g.CA; gC += g.C;
    if (i + 1 < A.L) {
      owner_geometry_grad(A, cf, b, i + 1, g);
      gCA += g.CAm; gC += g.Cm;
    }
  }
}

// cross-entropy gradient row; cfs = coef[SEQ] * inv_den[SEQ]
PEV_HD void ce_row_bwd(const pev_loss_args& A, float cfs, int64_t bi, float* g) {
  const float* lg = A.logits + bi * A.C;
  float mx = lg[0];
  for (int k = 1; k < A.C; ++k) mx = fmaxf(mx, lg[k]);
  float se = 0.f;
  for (int k = 0; k < A.C; ++k) se += expf(lg[k] - mx);
  int64_t lab = A.labels[bi];
  const bool used = lab >= 0 && lab < A.C;                             // ignored label: zero gradient row
  float s = used ? cfs * A.mask[bi] / se : 0.f;
  for (int k = 0; k < A.C; ++k) g[k] = s * expf(lg[k] - mx) - ((used && k == lab) ? cfs * A.mask[bi] : 0.f);
}

// ------------------------------------------------------------------------------------------
// pair_distance_loss row (models/losses.py:24-37): point i against all M selected points.
// P,T: [M,3] selected coordinates; pm: [M] their mask values.
PEV_HD void pair_row(const float* P, const float* T, const float* pm, int M, int i, float* num,
                     float* den, v3* grad, float cf) {
  v3 pi = ld3(P + 3 * i), ti = ld3(T + 3 * i);
  float mi = pm[i], n = 0.f, d = 0.f;
  v3 g = zero3();
  for (int j = 0; j < M; ++j) {
    v3 dp = pi - ld3(P + 3 * j);
    float w = mi * pm[j];
    float lp = norm(dp), lt = norm(ti - ld3(T + 3 * j));
    n += fabsf(lp - lt) * w;
    d += w;
    if (grad && lp > 0.f) g += dp * (signf_(lp - lt) * w / lp);
  }
  *num += n;
  *den += d;
  if (grad) *grad = g * (2.0f * cf);      // the (i,j) and (j,i) entries both carry point i
}

// clash_loss row (models/losses.py:439-517): atom a of a conformer against all other atoms.
// atoms: [3L,3] interleaved N,CA,C; am: [3L] atom masks.  num/den count each unordered pair twice.
PEV_HD void clash_row(const float* atoms, const float* am, int n_atoms, int a, float clash_dist,
                      float soft_margin, float* num, float* den, v3* grad, float cf) {
  v3 pa = ld3(atoms + 3 * a);
  float ma = am[a], n = 0.f, d = 0.f;
  int ra = a / 3;
  v3 g = zero3();
  for (int c = 0; c < n_atoms; ++c) {
    int sep = c / 3 - ra;
    if (sep < 2 && sep > -2) continue;                                  // :478-482
    float w = ma * am[c];
    v3 dv = pa - ld3(atoms + 3 * c);
    float dist = norm(dv);
    float v = fmaxf(clash_dist - dist, 0.f);                                  // :494-495
    n += (v < soft_margin ? 0.5f * v * v : v * v) * w;                         // :500-504
    d += w;
    if (grad && v > 0.f && dist > 0.f) g -= dv * ((v < soft_margin ? v : 2.0f * v) * w / dist);
  }
  *num += n;
  *den += d;
  if (grad) *grad = g * cf;
}

// ------------------------------------------------------------------------------------------
// dihedral-space terms on an explicit [B,L,6] tensor (public dihedral_consistency_loss,
// ramachandran_loss, omega_trans_loss; models/losses.py:60-155).
PEV_HD void dihterm_fwd(const float* dih, const float* tgt, float m, float* cons_num, float* cons_den,
                        float* rama, float* omega) {
  if (tgt)
    for (int k = 0; k < 6; ++k)
      if ((m != 0.f) && finitef_(dih[k]) && finitef_(tgt[k])) {
        float d = dih[k] - tgt[k];
        *cons_num += d * d;
        *cons_den += 1.f;
      }
  *rama += rama_penalty(dih[0], dih[1], dih[2], dih[3], nullptr, nullptr) * m;
  *omega += omega_penalty(dih[4], dih[5], nullptr) * m;
}
PEV_HD void dihterm_bwd(const float* dih, const float* tgt, float m, const float* cf3, float* g) {
  for (int k = 0; k < 6; ++k) g[k] = 0.f;
  if (tgt)
    for (int k = 0; k < 6; ++k)
      if ((m != 0.f) && finitef_(dih[k]) && finitef_(tgt[k])) g[k] += cf3[0] * 2.0f * (dih[k] - tgt[k]);
  float dphi, dpsi, dom, a, b;
  rama_penalty(dih[0], dih[1], dih[2], dih[3], &dphi, &dpsi);
  atan2_grad(dih[0], dih[1], cf3[1] * m * dphi, &a, &b); g[0] += a; g[1] += b;
  atan2_grad(dih[2], dih[3], cf3[1] * m * dpsi, &a, &b); g[2] += a; g[3] += b;
  omega_penalty(dih[4], dih[5], &dom);
  atan2_grad(dih[4], dih[5], cf3[2] * m * dom, &a, &b); g[4] += a; g[5] += b;
}

// compute_dihedrals_from_coords backward for residue i: collects the pieces of the torsions of
// residues i-1, i, i+1 that land on residue i's atoms.  gout: [B,L,6].
PEV_HD void dihedrals_bwd_residue(const pev_loss_args& A, const float* gout, int b, int i, v3& gN,
                                  v3& gCA, v3& gC) {
  gN = gCA = gC = zero3();
  v3 q0, q1, q2, q3;
  for (int o = i - 1; o <= i + 1; ++o) {
    if (o < 0 || o >= A.L) continue;
    ResDih r;
    residue_dihedrals(A, A.pred_N, A.pred_CA, A.pred_C, b, o, r);
    const float* g = gout + ((int64_t)b * A.L + o) * 6;
    if (r.has_prev) {
      dihedral_bwd(r.phi, g[0], g[1], q0, q1, q2, q3);          // C(o-1) N CA C
      if (o == i) { gN += q1; gCA += q2; gC += q3; }
      if (o == i + 1) gC += q0;
      dihedral_bwd(r.om, g[4], g[5], q0, q1, q2, q3);           // CA(o-1) C(o-1) N CA
      if (o == i) { gN += q2; gCA += q3; }
      if (o == i + 1) { gCA += q0; gC += q1; }
    }
    if (r.has_next) {
      dihedral_bwd(r.psi, g[2], g[3], q0, q1, q2, q3);          // N CA C N(o+1)
      if (o == i) { gN += q0; gCA += q1; gC += q2; }
      if (o == i - 1) gN += q3;
    }
  }
}

}  // namespace pev
