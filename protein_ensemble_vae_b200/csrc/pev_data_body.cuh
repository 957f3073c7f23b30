// Ragged batch -> padded batch: the per-row body of the device-side replacement for
// models/data.py's centring (:166-172) and zero-padding collate (:219-266).  Host/device (tests/hostcheck).
#pragma once
#include <stdint.h>

#include "pev_hd.cuh"

namespace pev {

struct UnpackArgs {
  const float* n;          // [T,3] packed rows of all conformers (T = cu[B])
  const float* ca;
  const float* c;
  const float* mask;       // [T]
  const float* dih;        // [T,6]
  const int64_t* labels;   // [T]
  const float* emb;        // [T,D] or null
  const int32_t* cu;       // [B+1]
  int32_t B, Lmax, D;
  float* o_n;              // [B,Lmax,3]
  float* o_ca;
  float* o_c;
  float* o_mask;           // [B,Lmax]
  float* o_dih;            // [B,Lmax,6]
  int64_t* o_labels;       // [B,Lmax]
  float* o_emb;            // [B,Lmax,D] or null
};

// row l of conformer b: centred coordinates (centroid cen of the valid CA atoms, :167-172), mask, dihedrals, label;
// rows past the conformer's length are zeros (:238-243)
PEV_HD void unpack_row(const UnpackArgs& a, int b, int l, const float* cen) {
  const int64_t o = (int64_t)b * a.Lmax + l;
  const int L = a.cu[b + 1] - a.cu[b];
  if (l < L) {
    const int64_t t = (int64_t)a.cu[b] + l;
    for (int k = 0; k < 3; ++k) {
      a.o_n[3 * o + k] = a.n[3 * t + k] - cen[k];
      a.o_ca[3 * o + k] = a.ca[3 * t + k] - cen[k];
      a.o_c[3 * o + k] = a.c[3 * t + k] - cen[k];
    }
    a.o_mask[o] = a.mask[t];
    for (int k = 0; k < 6; ++k) a.o_dih[6 * o + k] = a.dih[6 * t + k];
    a.o_labels[o] = a.labels[t];
  } else {
    for (int k = 0; k < 3; ++k) a.o_n[3 * o + k] = a.o_ca[3 * o + k] = a.o_c[3 * o + k] = 0.f;
    a.o_mask[o] = 0.f;
    for (int k = 0; k < 6; ++k) a.o_dih[6 * o + k] = 0.f;
    a.o_labels[o] = 0;
  }
}

}  // namespace pev
