// Per-thread bodies of the graph builder, the fp32 edge prologue (K1, fp32 form) and the
// exact-order segmented scatter / coordinate update (K2).  Host/device, see pev_hd.cuh.
// Reference: models/en_gnn_decoder.py:53-87 (EGNLayer.forward), :174-198 (graph).
#pragma once
#include "pev_hd.cuh"

namespace pev {

// ------------------------------------------------------------------------------------------ graph
// sum_{k<i} min(k, W)
PEV_HD int64_t band_prefix_lo(int64_t i, int64_t W) {
  return i <= W + 1 ? i * (i - 1) / 2 : W * (W + 1) / 2 + (i - W - 1) * W;
}
// number of band edges whose destination is one of the first i residues of an Lb-residue chain
PEV_HD int64_t band_row_offset(int64_t i, int64_t Lb, int64_t W) {
  return band_prefix_lo(i, W) + band_prefix_lo(Lb, W) - band_prefix_lo(Lb - i, W);
}
PEV_HD int band_lo(int i, int W) { return i - W > 0 ? i - W : 0; }
PEV_HD int band_hi(int i, int Lb, int W) { return i + W < Lb - 1 ? i + W : Lb - 1; }   // inclusive
PEV_HD int band_degree(int i, int Lb, int W) { return band_hi(i, Lb, W) - band_lo(i, W); }

PEV_HD int find_conformer(const int32_t* cu, int B, int64_t node) {
  int lo = 0, hi = B - 1;                       // last b with cu[b] <= node
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if ((int64_t)cu[mid] <= node) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// everything the graph builder writes for one node
PEV_HD void band_node_build(const int32_t* cu, const int64_t* edge_base, int B, int W, int64_t node,
                            int32_t* row_ptr, int32_t* row, int32_t* col, int32_t* csc_perm,
                            float* dinv, bool last_node) {
  int b = find_conformer(cu, B, node);
  int start = cu[b], Lb = cu[b + 1] - start, i = (int)(node - start);
  int64_t e0 = edge_base[b] + band_row_offset(i, Lb, W);
  int lo = band_lo(i, W), hi = band_hi(i, Lb, W), deg = hi - lo;
  if (row_ptr) {
    row_ptr[node] = (int32_t)e0;
    if (last_node) row_ptr[node + 1] = (int32_t)(e0 + deg);
  }
  if (dinv) dinv[node] = deg > 0 ? 1.0f / (float)deg : 0.f;
  int64_t e = e0;
  for (int j = lo; j <= hi; ++j) {
    if (j == i) continue;
    if (row) row[e] = (int32_t)node;
    if (col) col[e] = start + j;
    if (csc_perm) {
      // e-th slot of column `node`: the edge (j <- i), stored in row j's segment at i's rank
      int64_t ej = edge_base[b] + band_row_offset(j, Lb, W) + (i - band_lo(j, W)) - (i > j ? 1 : 0);
      csc_perm[e] = (int32_t)ej;
    }
    ++e;
  }
}

// ------------------------------------------------------------------------------------------ K1 fp32
// u[e,k] for one edge/feature; models/en_gnn_decoder.py:61-65 with the first Linear factored.
PEV_HD float edge_d2(const float* x, int r, int c) {
  v3 rel = ld3(x + 3 * (int64_t)r) - ld3(x + 3 * (int64_t)c);
  return dot(rel, rel);
}
PEV_HD float edge_prologue_elem(const float* AB, const float* wd, const float* b1, int H, int r, int c,
                                float d2, int k) {
  return AB[(int64_t)r * 2 * H + k] + AB[(int64_t)c * 2 * H + H + k] + wd[k] * d2 + b1[k];
}

// backward, feature part: for node i and feature k
PEV_HD void edge_prologue_bwd_feat(const float* gu, const float* x, const int32_t* row_ptr,
                                   const int32_t* col, const int32_t* col_ptr, const int32_t* csc_perm,
                                   int H, int64_t i, int k, float* gA, float* gB, float* gwd_part) {
  float a = 0.f, b = 0.f, p = 0.f;
  for (int64_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
    float g = gu[e * H + k];
    a += g;
    p += g * edge_d2(x, (int)i, col[e]);
  }
  for (int64_t q = col_ptr[i]; q < col_ptr[i + 1]; ++q) b += gu[(int64_t)csc_perm[q] * H + k];
  *gA = a; *gB = b; *gwd_part = p;
}
// backward, coordinate part for node i: d2 = |x_r - x_c|^2 -> gx_r += 2 gd2 rel, gx_c -= 2 gd2 rel
PEV_HD v3 edge_prologue_bwd_coord(const float* gd2, const float* x, const int32_t* row_ptr,
                                  const int32_t* row, const int32_t* col, const int32_t* col_ptr,
                                  const int32_t* csc_perm, int64_t i) {
  v3 g = zero3(), xi = ld3(x + 3 * i);
  for (int64_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e)
    g += (xi - ld3(x + 3 * (int64_t)col[e])) * (2.0f * gd2[e]);
  for (int64_t q = col_ptr[i]; q < col_ptr[i + 1]; ++q) {
    int64_t e = csc_perm[q];
    g -= (ld3(x + 3 * (int64_t)row[e]) - xi) * (2.0f * gd2[e]);
  }
  return g;
}

// ------------------------------------------------------------------------------------------ K2
#if defined(__CUDA_ARCH__)
#define PEV_FADD(a, b) __fadd_rn((a), (b))
#define PEV_FMUL(a, b) __fmul_rn((a), (b))
#define PEV_FSUB(a, b) __fsub_rn((a), (b))
#else
// host build: tests/hostcheck is compiled with -ffp-contract=off so these never fuse
#define PEV_FADD(a, b) ((a) + (b))
#define PEV_FMUL(a, b) ((a) * (b))
#define PEV_FSUB(a, b) ((a) - (b))
#endif

// agg[i,k]: plain fp32 adds in ascending edge order == CPU index_add_ (models/en_gnn_decoder.py:68-69)
PEV_HD float scatter_feature(const float* m, const int32_t* row_ptr, int H, int64_t i, int k) {
  float acc = 0.f;
  for (int64_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e) acc = PEV_FADD(acc, m[e * H + k]);
  return acc;
}
// x_out[i] following the reference's op order exactly (:61, :78-86): rel = x_i - x_j (fp32),
// prod = w * rel, sequential sum, * dinv, * 0.2, x + delta.  No FMA contraction.
PEV_HD v3 coord_update(const float* w, const float* x, const float* dinv, const int32_t* row_ptr,
                       const int32_t* col, int64_t i) {
  v3 xi = ld3(x + 3 * i), acc = zero3();
  for (int64_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
    v3 xj = ld3(x + 3 * (int64_t)col[e]);
    float we = w[e];
    acc.x = PEV_FADD(acc.x, PEV_FMUL(we, PEV_FSUB(xi.x, xj.x)));
    acc.y = PEV_FADD(acc.y, PEV_FMUL(we, PEV_FSUB(xi.y, xj.y)));
    acc.z = PEV_FADD(acc.z, PEV_FMUL(we, PEV_FSUB(xi.z, xj.z)));
  }
  if (dinv) {
    float d = dinv[i];
    acc.x = PEV_FMUL(acc.x, d); acc.y = PEV_FMUL(acc.y, d); acc.z = PEV_FMUL(acc.z, d);
  }
  acc.x = PEV_FMUL(acc.x, 0.2f); acc.y = PEV_FMUL(acc.y, 0.2f); acc.z = PEV_FMUL(acc.z, 0.2f);
  return mk3(PEV_FADD(xi.x, acc.x), PEV_FADD(xi.y, acc.y), PEV_FADD(xi.z, acc.z));
}

// backward of coord_update wrt w[e]
PEV_HD float coord_update_bwd_w(const float* gxo, const float* x, const float* dinv, int r, int c) {
  float s = 0.2f * (dinv ? dinv[r] : 1.0f);
  v3 rel = ld3(x + 3 * (int64_t)r) - ld3(x + 3 * (int64_t)c);
  return s * dot(ld3(gxo + 3 * (int64_t)r), rel);
}
// backward of coord_update wrt x[i]
PEV_HD v3 coord_update_bwd_x(const float* gxo, const float* w, const float* dinv, const int32_t* row_ptr,
                             const int32_t* row, const int32_t* col_ptr, const int32_t* csc_perm,
                             int64_t i) {
  v3 gi = ld3(gxo + 3 * i);
  float si = 0.2f * (dinv ? dinv[i] : 1.0f), wsum = 0.f;
  for (int64_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e) wsum += w[e];
  v3 g = gi + gi * (si * wsum);
  for (int64_t q = col_ptr[i]; q < col_ptr[i + 1]; ++q) {
    int64_t e = csc_perm[q];
    int r = row[e];
    g -= ld3(gxo + 3 * (int64_t)r) * (0.2f * (dinv ? dinv[r] : 1.0f) * w[e]);
  }
  return g;
}

}  // namespace pev
