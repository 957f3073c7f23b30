// TEST INFRASTRUCTURE ONLY.  Compiles the per-thread bodies of the CUDA kernels
// (protein_ensemble_vae_b200/csrc/pev_*_body.cuh) with g++ and loops over them on the CPU,
// exporting the same extern "C" names and signatures as include/pev_b200.h for the subset that
// has no tensor-core path.  It lets `pytest -m "not gpu"` check the kernels' arithmetic and the
// Python host logic against the oracle on a machine without a GPU.  The product never loads it.
#include <cstring>
#include <vector>

#include "../../include/pev_b200.h"
#include "../../protein_ensemble_vae_b200/csrc/pev_egnn_body.cuh"
#include "../../protein_ensemble_vae_b200/csrc/pev_kabsch_body.cuh"
#include "../../protein_ensemble_vae_b200/csrc/pev_backbone_body.cuh"
#include "../../protein_ensemble_vae_b200/csrc/pev_data_body.cuh"
#include "../../protein_ensemble_vae_b200/csrc/pev_metrics_body.cuh"
#include "../../protein_ensemble_vae_b200/csrc/pev_pdb_body.cuh"
#include "../../protein_ensemble_vae_b200/csrc/pev_loss_body.cuh"
#include "../../protein_ensemble_vae_b200/csrc/pev_loss_final.cuh"

using namespace pev;

extern "C" {

int pev_abi_version(void) { return PEV_ABI_VERSION; }
const char* pev_last_error(void) { return ""; }
int64_t pev_launch_count(void) { return 0; }

int pev_band_graph_build(const int32_t* cu, const int64_t* edge_base, int32_t B, int32_t W, int64_t N,
                         int32_t* row_ptr, int32_t* row, int32_t* col, int32_t* csc_perm, float* dinv,
                         void*) {
  if (N == 0 && row_ptr) row_ptr[0] = 0;
  for (int64_t n = 0; n < N; ++n)
    band_node_build(cu, edge_base, B, W, n, row_ptr, row, col, csc_perm, dinv, n == N - 1);
  return 0;
}

int pev_edge_prologue_fwd(const float* AB, const float* x, const float* wd, const float* b1,
                          const int32_t* row, const int32_t* col, int64_t E, int32_t H, float* u, void*) {
  for (int64_t e = 0; e < E; ++e) {
    float d2 = edge_d2(x, row[e], col[e]);
    for (int k = 0; k < H; ++k) u[e * H + k] = edge_prologue_elem(AB, wd, b1, H, row[e], col[e], d2, k);
  }
  return 0;
}

int pev_edge_prologue_bwd(const float* gu, const float* x, const float* wd, const int32_t* row_ptr,
                          const int32_t* row, const int32_t* col, const int32_t* col_ptr,
                          const int32_t* csc_perm, int64_t N, int64_t E, int32_t H, float* gAB, float* gx,
                          float* gwd_part, float* gd2, void*) {
  for (int64_t e = 0; e < E; ++e) {
    float s = 0.f;
    for (int k = 0; k < H; ++k) s += wd[k] * gu[e * H + k];
    gd2[e] = s;
  }
  for (int64_t i = 0; i < N; ++i) {
    for (int k = 0; k < H; ++k)
      edge_prologue_bwd_feat(gu, x, row_ptr, col, col_ptr, csc_perm, H, i, k, &gAB[i * 2 * H + k],
                             &gAB[i * 2 * H + H + k], &gwd_part[i * H + k]);
    st3(gx + 3 * i, edge_prologue_bwd_coord(gd2, x, row_ptr, row, col, col_ptr, csc_perm, i));
  }
  return 0;
}

int pev_scatter_coord_fwd(const float* m, const float* w, const float* x, const float* dinv,
                          const int32_t* row_ptr, const int32_t* col, int64_t N, int32_t H, float* agg,
                          float* x_out, void*) {
  for (int64_t i = 0; i < N; ++i) {
    if (agg)
      for (int k = 0; k < H; ++k) agg[i * H + k] = scatter_feature(m, row_ptr, H, i, k);
    if (x_out) st3(x_out + 3 * i, coord_update(w, x, dinv, row_ptr, col, i));
  }
  return 0;
}

int pev_scatter_coord_bwd(const float* gagg, const float* gxo, const float* w, const float* x,
                          const float* dinv, const int32_t* row_ptr, const int32_t* row,
                          const int32_t* col, const int32_t* col_ptr, const int32_t* csc_perm, int64_t N,
                          int64_t E, int32_t H, float* gm, float* gw, float* gx, void*) {
  for (int64_t e = 0; e < E; ++e) {
    if (gm) std::memcpy(gm + e * H, gagg + (int64_t)row[e] * H, sizeof(float) * H);
    if (gw) gw[e] = coord_update_bwd_w(gxo, x, dinv, row[e], col[e]);
  }
  if (gx)
    for (int64_t i = 0; i < N; ++i)
      st3(gx + 3 * i, coord_update_bwd_x(gxo, w, dinv, row_ptr, row, col_ptr, csc_perm, i));
  return 0;
}

// ------------------------------------------------------------------------------------------ losses
static void gather_sel(const pev_loss_args& A, int b, std::vector<float>& P, std::vector<float>& T,
                       std::vector<float>& pm, int M) {
  for (int s = 0; s < M; ++s) {
    int i = s * A.pair_stride;
    for (int k = 0; k < 3; ++k) {
      P[3 * s + k] = A.pred_CA[((int64_t)b * A.L + i) * 3 + k];
      T[3 * s + k] = A.target_CA[((int64_t)b * A.L + i) * 3 + k];
    }
    pm[s] = A.mask[(int64_t)b * A.L + i];
  }
}
static void gather_atoms(const pev_loss_args& A, int b, std::vector<float>& at, std::vector<float>& am) {
  for (int i = 0; i < A.L; ++i) {
    const float* src[3] = {A.pred_N, A.pred_CA, A.pred_C};
    for (int t = 0; t < 3; ++t) {
      for (int k = 0; k < 3; ++k) at[(3 * i + t) * 3 + k] = src[t][((int64_t)b * A.L + i) * 3 + k];
      am[3 * i + t] = A.mask[(int64_t)b * A.L + i];
    }
  }
}

int pev_loss_fwd(const pev_loss_args* Ap, double* ag, double* as, void*) {
  const pev_loss_args& A = *Ap;
  for (int b = 0; b < A.B; ++b) {
    float acc[RA_COUNT];
    for (int k = 0; k < RA_COUNT; ++k) acc[k] = 0.f;
    for (int i = 0; i < A.L; ++i) {
      float a1[RA_COUNT];
      for (int k = 0; k < RA_COUNT; ++k) a1[k] = 0.f;
      residue_fwd(A, b, i, a1);
      for (int k = 0; k < RA_COUNT; ++k) acc[k] += a1[k];
    }
    scatter_residue_acc(acc, ag, as + 8 * b);
    if (A.mu_l) {
      double s = 0.0;
      for (int i = 0; i < A.L; ++i) {
        int64_t bi = (int64_t)b * A.L + i;
        float r = 0.f;
        for (int d = 0; d < A.D; ++d) r += kl_elem(A.mu_l[bi * A.D + d], A.lv_l[bi * A.D + d]);
        s += r * A.mask[bi];
      }
      ag[2 * PEV_T_KL_L] += s;
    }
    if (A.mu_g) {
      float r = 0.f;
      for (int d = 0; d < A.G; ++d) r += kl_elem(A.mu_g[(int64_t)b * A.G + d], A.lv_g[(int64_t)b * A.G + d]);
      ag[2 * PEV_T_KL_G] += r;
    }
    if (A.pair_stride > 0) {
      int M = (A.L + A.pair_stride - 1) / A.pair_stride;
      std::vector<float> P(3 * M), T(3 * M), pm(M);
      gather_sel(A, b, P, T, pm, M);
      float n = 0.f, d = 0.f;
      for (int i = 0; i < M; ++i) pair_row(P.data(), T.data(), pm.data(), M, i, &n, &d, nullptr, 0.f);
      ag[2 * PEV_T_PAIR] += n;
      ag[2 * PEV_T_PAIR + 1] += d;
    }
    if (A.enable_clash) {
      std::vector<float> at(9 * A.L), am(3 * A.L);
      gather_atoms(A, b, at, am);
      float n = 0.f, d = 0.f;
      for (int a = 0; a < 3 * A.L; ++a) clash_row(at.data(), am.data(), 3 * A.L, a, A.clash_dist, A.soft_margin, &n, &d, nullptr, 0.f);
      as[8 * b + 4] += n;
      as[8 * b + 5] += d;
    }
  }
  return 0;
}

int pev_loss_finalize(const double* ag, const double* as, int32_t B, float* terms, float* inv_den, void*) {
  double part[FIN_PARTS];
  for (int k = 0; k < FIN_PARTS; ++k) part[k] = 0.0;
  finalize_partial(as, B, 0, 1, part, inv_den);
  finalize_combine(ag, part, B, terms, inv_den);
  return 0;
}

int pev_loss_bwd(const pev_loss_args* Ap, const float* coef, const float* inv_den, float* gN, float* gCA,
                 float* gC, float* glog, float* gmul, float* glvl, float* gmug, float* glvg, void*) {
  const pev_loss_args& A = *Ap;
  float cf[PEV_NUM_TERMS];
  for (int t = 0; t < PEV_NUM_TERMS; ++t) cf[t] = coef[t] * inv_den[t];
  for (int b = 0; b < A.B; ++b) {
    float cf_rec[3];
    for (int k = 0; k < 3; ++k) cf_rec[k] = coef[PEV_T_REC_CA + k] * inv_den[PEV_NUM_TERMS + b];
    float cf_clash = coef[PEV_T_CLASH] * inv_den[PEV_NUM_TERMS + A.B + b];
    std::vector<float> P, T, pm, at, am;
    int M = 0;
    if (A.pair_stride > 0) {
      M = (A.L + A.pair_stride - 1) / A.pair_stride;
      P.resize(3 * M), T.resize(3 * M), pm.resize(M);
      gather_sel(A, b, P, T, pm, M);
    }
    if (A.enable_clash) {
      at.resize(9 * A.L), am.resize(3 * A.L);
      gather_atoms(A, b, at, am);
    }
    for (int i = 0; i < A.L; ++i) {
      int64_t bi = (int64_t)b * A.L + i;
      v3 n, ca, c;
      residue_bwd(A, cf, cf_rec, b, i, n, ca, c);
      float dn = 0.f, dd = 0.f;
      v3 g;
      if (A.pair_stride > 0 && i % A.pair_stride == 0) {
        pair_row(P.data(), T.data(), pm.data(), M, i / A.pair_stride, &dn, &dd, &g, cf[PEV_T_PAIR]);
        ca += g;
      }
      if (A.enable_clash) {
        clash_row(at.data(), am.data(), 3 * A.L, 3 * i + 0, A.clash_dist, A.soft_margin, &dn, &dd, &g, cf_clash); n += g;
        clash_row(at.data(), am.data(), 3 * A.L, 3 * i + 1, A.clash_dist, A.soft_margin, &dn, &dd, &g, cf_clash); ca += g;
        clash_row(at.data(), am.data(), 3 * A.L, 3 * i + 2, A.clash_dist, A.soft_margin, &dn, &dd, &g, cf_clash); c += g;
      }
      if (gN) st3(gN + 3 * bi, n);
      if (gCA) st3(gCA + 3 * bi, ca);
      if (gC) st3(gC + 3 * bi, c);
      if (glog && A.logits) ce_row_bwd(A, cf[PEV_T_SEQ], bi, glog + bi * A.C);
      if (A.mu_l && gmul)
        for (int d = 0; d < A.D; ++d)
          kl_elem_bwd(A.mu_l[bi * A.D + d], A.lv_l[bi * A.D + d], cf[PEV_T_KL_L] * A.mask[bi],
                      &gmul[bi * A.D + d], &glvl[bi * A.D + d]);
    }
    if (A.mu_g && gmug)
      for (int d = 0; d < A.G; ++d)
        kl_elem_bwd(A.mu_g[(int64_t)b * A.G + d], A.lv_g[(int64_t)b * A.G + d], cf[PEV_T_KL_G],
                    &gmug[(int64_t)b * A.G + d], &glvg[(int64_t)b * A.G + d]);
  }
  return 0;
}

int pev_dihedrals_fwd(const float* N, const float* CA, const float* C, const float* mask, int32_t B,
                      int32_t L, float* out, void*) {
  pev_loss_args A;
  std::memset(&A, 0, sizeof(A));
  A.mask = mask; A.B = B; A.L = L;
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < L; ++i) {
      ResDih r;
      residue_dihedrals(A, N, CA, C, b, i, r);
      for (int k = 0; k < 6; ++k) out[((int64_t)b * L + i) * 6 + k] = r.slot[k];
    }
  return 0;
}

int pev_dihedrals_bwd(const float* N, const float* CA, const float* C, const float* mask,
                      const float* gout, int32_t B, int32_t L, float* gN, float* gCA, float* gC, void*) {
  pev_loss_args A;
  std::memset(&A, 0, sizeof(A));
  A.mask = mask; A.B = B; A.L = L; A.pred_N = N; A.pred_CA = CA; A.pred_C = C;
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < L; ++i) {
      v3 n, ca, c;
      dihedrals_bwd_residue(A, gout, b, i, n, ca, c);
      int64_t bi = (int64_t)b * L + i;
      st3(gN + 3 * bi, n); st3(gCA + 3 * bi, ca); st3(gC + 3 * bi, c);
    }
  return 0;
}

int pev_dihedral_terms_fwd(const float* dih, const float* target, const float* mask, int32_t B, int32_t L,
                           double* sums, void*) {
  for (int64_t bi = 0; bi < (int64_t)B * L; ++bi) {
    float cn = 0.f, cd = 0.f, ra = 0.f, om = 0.f;
    dihterm_fwd(dih + 6 * bi, target ? target + 6 * bi : nullptr, mask[bi], &cn, &cd, &ra, &om);
    sums[0] += cn; sums[1] += cd; sums[2] += ra; sums[3] += om; sums[4] += mask[bi];
  }
  return 0;
}

int pev_dihedral_terms_bwd(const float* dih, const float* target, const float* mask, const float* coef3,
                           int32_t B, int32_t L, float* gdih, void*) {
  for (int64_t bi = 0; bi < (int64_t)B * L; ++bi)
    dihterm_bwd(dih + 6 * bi, target ? target + 6 * bi : nullptr, mask[bi], coef3, gdih + 6 * bi);
  return 0;
}

int pev_kabsch_rmsd(const float* a, const float* b, const float* mask, int32_t S, int32_t L,
                    int32_t b_batch, int32_t mask_batch, int32_t mode, float* out, void*) {
  for (int s = 0; s < S; ++s)
    out[s] = kabsch_rmsd_serial(a + (int64_t)s * L * 3, b + (b_batch ? (int64_t)s * L * 3 : 0),
                                mask ? mask + (mask_batch ? (int64_t)s * L : 0) : nullptr, L, mode);
  return 0;
}

int pev_validate_geometry(const float* ca, const float* mask, int32_t S, int32_t L, int32_t mask_batch, int32_t* status,
                          float* stats, void*) {
  for (int s = 0; s < S; ++s)
    status[s] = validate_geometry_serial(ca + (int64_t)s * L * 3, mask ? mask + (mask_batch ? (int64_t)s * L : 0) : nullptr,
                                         L, stats ? stats + 3 * s : nullptr);
  return 0;
}

int pev_superpose_scores(const float* a, const float* b, const float* mask, int32_t S, int32_t L, int32_t b_batch,
                         int32_t mask_batch, float* aligned, float* dist, float* tm, float* gdt_ts, float* gdt_ha, void*) {
  for (int s = 0; s < S; ++s) {
    float* d = dist + (int64_t)s * L;
    superpose_serial(a + (int64_t)s * L * 3, b + (b_batch ? (int64_t)s * L * 3 : 0), L,
                     aligned ? aligned + (int64_t)s * L * 3 : nullptr, d);
    float t, g1, g2;
    superposition_scores(d, mask ? mask + (mask_batch ? (int64_t)s * L : 0) : nullptr, L, &t, &g1, &g2);
    if (tm) tm[s] = t;
    if (gdt_ts) gdt_ts[s] = g1;
    if (gdt_ha) gdt_ha[s] = g2;
  }
  return 0;
}

int pev_lddt(const float* pred, const float* tru, const float* mask, int32_t S, int32_t L, int32_t t_batch,
             int32_t mask_batch, float cutoff, float* per_residue, float* global, void*) {
  for (int s = 0; s < S; ++s) {
    const float* m = mask ? mask + (mask_batch ? (int64_t)s * L : 0) : nullptr;
    float sum = 0.f, cnt = 0.f;
    for (int i = 0; i < L; ++i) {
      const float v = lddt_residue(pred + (int64_t)s * L * 3, tru + (t_batch ? (int64_t)s * L * 3 : 0), m, L, i, cutoff);
      per_residue[(int64_t)s * L + i] = v;
      if (!m || m[i] != 0.f) { sum += v; cnt += 1.f; }
    }
    global[s] = cnt > 0.f ? sum / cnt : 0.f;
  }
  return 0;
}

int pev_rmsf(const float* aligned, int32_t N, int32_t L, float* out, void*) {
  for (int l = 0; l < L; ++l) out[l] = rmsf_residue(aligned, N, L, l);
  return 0;
}

int pev_unpack_center(const float* n, const float* ca, const float* c, const float* mask, const float* dih,
                      const int64_t* labels, const float* emb, const int32_t* cu, int32_t B, int32_t Lmax, int32_t D,
                      int32_t center, float* o_n, float* o_ca, float* o_c, float* o_mask, float* o_dih, int64_t* o_labels,
                      float* o_emb, void*) {
  UnpackArgs a = {n, ca, c, mask, dih, labels, emb, cu, B, Lmax, D, o_n, o_ca, o_c, o_mask, o_dih, o_labels, o_emb};
  for (int b = 0; b < B; ++b) {
    const int L = cu[b + 1] - cu[b];
    double s[3] = {0, 0, 0}, cnt = 0;
    if (center)
      for (int l = 0; l < L; ++l)
        if (mask[cu[b] + l] != 0.f) {
          for (int k = 0; k < 3; ++k) s[k] += ca[3 * (int64_t)(cu[b] + l) + k];
          cnt += 1;
        }
    const float cen[3] = {cnt > 0 ? (float)(s[0] / cnt) : 0.f, cnt > 0 ? (float)(s[1] / cnt) : 0.f, cnt > 0 ? (float)(s[2] / cnt) : 0.f};
    for (int l = 0; l < Lmax; ++l) {
      unpack_row(a, b, l, cen);
      if (emb)
        for (int d = 0; d < D; ++d)
          o_emb[((int64_t)b * Lmax + l) * D + d] = l < L ? emb[((int64_t)cu[b] + l) * D + d] : 0.f;
    }
  }
  return 0;
}

int pev_backbone_fwd(const float* n_dir, int32_t ldn, const float* c_dir, int32_t ldc, const float* x_ca, const uint8_t* starts,
                     int64_t N, float* x_n, float* x_c, void*) {
  for (int64_t k = 0; k < N; ++k) backbone_fwd_residue(n_dir, ldn, c_dir, ldc, x_ca, starts, N, k, x_n, x_c);
  return 0;
}

int pev_backbone_bwd(const float* n_dir, int32_t ldn, const float* c_dir, int32_t ldc, const float* x_ca, const uint8_t* starts,
                     int64_t N, const float* g_xn, const float* g_xc, float* g_ndir, float* g_cdir, float* g_xca, void*) {
  for (int64_t k = 0; k < N; ++k)
    backbone_bwd_residue(n_dir, ldn, c_dir, ldc, x_ca, starts, N, k, g_xn, g_xc, g_ndir, g_cdir, g_xca);
  return 0;
}

int64_t pev_pdb_models_bytes(int64_t model0, int32_t S, int32_t nv) { return pdb_block_offset(model0 + S, model0, nv); }

int pev_pdb_format_models(const float* n, const float* ca, const float* c, const int32_t* valid_idx, const uint8_t* prev_ok,
                          const char* resname, int32_t S, int32_t L, int32_t nv, int64_t model0, int32_t chain, char* out,
                          int32_t* overflow, void*) {
  PdbArgs a = {n, ca, c, valid_idx, prev_ok, resname, S, L, nv, model0, (char)chain, out};
  for (int s = 0; s < S; ++s)
    for (int k = 0; k < (nv > 0 ? nv : 1); ++k)
      if (!pdb_residue(a, s, k)) *overflow = 1;
  return 0;
}

}  // extern "C"
