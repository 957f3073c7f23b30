// Accumulator layout, finalisation and the KL element of the loss kernels.  Host/device.
//   acc_global[2t], acc_global[2t+1]: numerator, denominator of base term t (doubles)
//   acc_sample[8b + {0,1,2}]: rmsd numerators CA,N,C; [3]: sum(mask); [4],[5]: clash numerator /
//   denominator, both counting every unordered atom pair twice.
//   inv_den[t] (t < PEV_NUM_TERMS): 1/denominator of term t; inv_den[NT + b] = 1/(B sum(mask_b));
//   inv_den[NT + B + b] = 1/(B (clash pairs_b + 1e-8)).
#pragma once
#include "../../include/pev_b200.h"
#include "pev_loss_body.cuh"

namespace pev {

enum { FIN_REC_CA = 0, FIN_REC_N, FIN_REC_C, FIN_MSUM, FIN_CLASH, FIN_PARTS };

// 0.5 (exp(lv) + mu^2 - 1 - lv), models/losses.py:43
PEV_HD float kl_elem(float mu, float lv) { return 0.5f * (expf(lv) + mu * mu - 1.0f - lv); }
PEV_HD void kl_elem_bwd(float mu, float lv, float g, float* gmu, float* glv) {
  *gmu = g * mu;
  *glv = g * 0.5f * (expf(lv) - 1.0f);
}

template <class AddF>
PEV_HD void scatter_residue_acc_t(const float* acc, double* ag, double* as_b, AddF add) {
  add(as_b + 0, (double)acc[RA_REC_CA]);
  add(as_b + 1, (double)acc[RA_REC_N]);
  add(as_b + 2, (double)acc[RA_REC_C]);
  add(as_b + 3, (double)acc[RA_MSUM]);
  add(ag + 2 * PEV_T_BOND_NCA, (double)acc[RA_NCA]);
  add(ag + 2 * PEV_T_BOND_CAC, (double)acc[RA_CAC]);
  add(ag + 2 * PEV_T_BOND_CN, (double)acc[RA_CN]);
  add(ag + 2 * PEV_T_BOND_CN + 1, (double)acc[RA_PM]);
  add(ag + 2 * PEV_T_ANG_NCAC, (double)acc[RA_NCAC]);
  add(ag + 2 * PEV_T_ANG_CNCA, (double)acc[RA_CNCA]);
  add(ag + 2 * PEV_T_ANG_CACN, (double)acc[RA_CACN]);
  add(ag + 2 * PEV_T_DIH_CONS, (double)acc[RA_CONS_NUM]);
  add(ag + 2 * PEV_T_DIH_CONS + 1, (double)acc[RA_CONS_DEN]);
  add(ag + 2 * PEV_T_RAMA, (double)acc[RA_RAMA]);
  add(ag + 2 * PEV_T_OMEGA, (double)acc[RA_OMEGA]);
  add(ag + 2 * PEV_T_SEQ, (double)acc[RA_CE]);
}

struct PlainAdd {
  PEV_HD void operator()(double* p, double v) const { *p += v; }
};
inline void scatter_residue_acc(const float* acc, double* ag, double* as_b) {
  scatter_residue_acc_t(acc, ag, as_b, PlainAdd());
}

// conformers start, start+step, ...: per-conformer ratios (models/losses.py:19-21, :510-515)
PEV_HD void finalize_partial(const double* as, int B, int start, int step, double* part, float* inv_den) {
  for (int b = start; b < B; b += step) {
    const double* s = as + 8 * (int64_t)b;
    double msum = s[3];
    part[FIN_REC_CA] += s[0] / msum;
    part[FIN_REC_N] += s[1] / msum;
    part[FIN_REC_C] += s[2] / msum;
    part[FIN_MSUM] += msum;
    double cden = 0.5 * s[5] + 1e-8;
    part[FIN_CLASH] += 0.5 * s[4] / cden;
    inv_den[PEV_NUM_TERMS + b] = (float)(1.0 / (msum * B));
    inv_den[PEV_NUM_TERMS + B + b] = (float)(1.0 / (cden * B));
  }
}

PEV_HD void finalize_combine(const double* ag, const double* part, int B, float* terms, float* inv_den) {
  double den[PEV_NUM_TERMS], num[PEV_NUM_TERMS];
  for (int t = 0; t < PEV_NUM_TERMS; ++t) { num[t] = ag[2 * t]; den[t] = ag[2 * t + 1]; }
  const double msum = part[FIN_MSUM], pm = ag[2 * PEV_T_BOND_CN + 1];
  num[PEV_T_REC_CA] = part[FIN_REC_CA]; den[PEV_T_REC_CA] = B;
  num[PEV_T_REC_N] = part[FIN_REC_N];   den[PEV_T_REC_N] = B;
  num[PEV_T_REC_C] = part[FIN_REC_C];   den[PEV_T_REC_C] = B;
  den[PEV_T_KL_G] = B;                                   // models/losses.py:51
  den[PEV_T_KL_L] = msum;                                // :57
  den[PEV_T_OMEGA] = msum;                               // :153
  den[PEV_T_RAMA] = msum;                                // :131
  den[PEV_T_BOND_NCA] = msum;                            // :337
  den[PEV_T_BOND_CAC] = msum;                            // :342
  den[PEV_T_BOND_CN] = pm;                               // :351
  den[PEV_T_ANG_NCAC] = msum;                            // :385
  den[PEV_T_ANG_CNCA] = pm;                              // :393
  den[PEV_T_ANG_CACN] = pm;                              // :403
  den[PEV_T_SEQ] = msum + 1e-8;                          // :435
  num[PEV_T_CLASH] = part[FIN_CLASH]; den[PEV_T_CLASH] = B;   // :514-515
  for (int t = 0; t < PEV_NUM_TERMS; ++t) {
    terms[t] = (float)(num[t] / den[t]);
    inv_den[t] = (float)(1.0 / den[t]);
  }
}

}  // namespace pev
