/*
 * pev_b200.h -- C ABI of the B200 (sm_100a) hot path of Protein-Ensemble-VAE:
 * EGNN decoder message passing, geometric losses, batched Kabsch RMSD.
 *
 * The reference (mohit03031999/Protein-Ensemble-VAE) is pure Python on PyTorch and has
 * no FFI; the boundary a replacement must honour is its Python module / function
 * signatures.  This header is the layer under those signatures: every
 * entry point names the reference code it replaces (paths relative to the reference
 * root).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - row-major, contiguous tensors; float = IEEE binary32; indices int32 unless stated;
 *   - no allocation, no stream synchronisation, no host callback inside any call;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return 0 on success, non-zero on error (message via pev_last_error(), thread-local);
 *   - optional pointers may be NULL where stated.
 */
#ifndef PEV_B200_H
#define PEV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PEV_ABI_VERSION 1

/* number of base loss terms produced by pev_loss_*; see enum below */
#define PEV_NUM_TERMS 17

enum pev_term {
  PEV_T_REC_CA = 0,   /* rmsd_loss(pred_CA, target_CA)          models/losses.py:12-21   */
  PEV_T_REC_N = 1,    /* rmsd_loss(pred_N, target_N)                                     */
  PEV_T_REC_C = 2,    /* rmsd_loss(pred_C, target_C)                                     */
  PEV_T_PAIR = 3,     /* pair_distance_loss                     models/losses.py:24-37   */
  PEV_T_KL_G = 4,     /* kl_global                              models/losses.py:49-51   */
  PEV_T_KL_L = 5,     /* kl_local                               models/losses.py:54-57   */
  PEV_T_DIH_CONS = 6, /* dihedral_consistency_loss              models/losses.py:60-69   */
  PEV_T_OMEGA = 7,    /* omega_trans_loss                       models/losses.py:136-155 */
  PEV_T_RAMA = 8,     /* ramachandran_loss                      models/losses.py:72-131  */
  PEV_T_BOND_NCA = 9, /* bond_length_loss: N-CA                 models/losses.py:335-337 */
  PEV_T_BOND_CAC = 10,/*                  CA-C                  models/losses.py:340-342 */
  PEV_T_BOND_CN = 11, /*                  C(i)-N(i+1)           models/losses.py:345-351 */
  PEV_T_ANG_NCAC = 12,/* bond_angle_loss: N-CA-C                models/losses.py:382-385 */
  PEV_T_ANG_CNCA = 13,/*                  C(i)-N(i+1)-CA(i+1)   models/losses.py:388-393 */
  PEV_T_ANG_CACN = 14,/*                  CA(i)-C(i)-N(i+1)     models/losses.py:398-403 */
  PEV_T_SEQ = 15,     /* sequence_classification_loss           models/losses.py:411-437 */
  PEV_T_CLASH = 16    /* clash_loss                             models/losses.py:439-517 */
};

/* ---------------------------------------------------------------- runtime */
int pev_abi_version(void);
const char* pev_last_error(void);
/* number of kernels this library has launched since load (for bench.py's gpu_launches) */
int64_t pev_launch_count(void);

/* ---------------------------------------------------------------- graph (integer work)
 * Replaces EGNNDecoder.build_edge_index / degrees, models/en_gnn_decoder.py:174-198, for a
 * packed batch: conformer b owns nodes [cu_seqlens[b], cu_seqlens[b+1]) and the edges
 * [edge_base[b], edge_base[b+1]); within a conformer edges i<-j, 0<|i-j|<=W, sorted (i,j).
 * Outputs: row_ptr[N+1], row[E], col[E] (global node ids), csc_perm[E] = edge ids sorted by
 * (col,row) -- the band is symmetric so col_ptr == row_ptr --, dinv[N] = 1/deg as float
 * (0 where deg==0).  edge_base is int64 [B+1]; any output may be NULL. */
int pev_band_graph_build(const int32_t* cu_seqlens, const int64_t* edge_base, int32_t num_conformers,
                         int32_t max_neighbors, int64_t num_nodes, int32_t* row_ptr, int32_t* row,
                         int32_t* col, int32_t* csc_perm, float* dinv, void* stream);

/* ---------------------------------------------------------------- K1 (fp32 form): edge prologue
 * u[e,:] = A[row e,:] + B[col e,:] + wd * |x_row - x_col|^2 + b1, where AB[N,2H] holds
 * A = h Wa^T (columns 0..H-1) and B = h Wb^T (columns H..2H-1): the factored first edge
 * linear of EGNLayer.forward, models/en_gnn_decoder.py:60-65. */
int pev_edge_prologue_fwd(const float* AB, const float* x, const float* wd, const float* b1,
                          const int32_t* row, const int32_t* col, int64_t num_edges, int32_t H,
                          float* u, void* stream);
/* backward of the above: gAB[N,2H], gx[N,3] (overwritten), gwd_part[N,H] (per-node partial
 * sums of gu*d2; reduce over nodes for gwd).  Deterministic segmented sums over the CSR row
 * segments and the CSC column segments; no atomics.  scratch_gd2: float[E]. */
int pev_edge_prologue_bwd(const float* gu, const float* x, const float* wd, const int32_t* row_ptr,
                          const int32_t* row, const int32_t* col, const int32_t* col_ptr,
                          const int32_t* csc_perm, int64_t num_nodes, int64_t num_edges, int32_t H,
                          float* gAB, float* gx, float* gwd_part, float* scratch_gd2, void* stream);

/* ---------------------------------------------------------------- K2: segmented scatter-sum +
 * coordinate update.  agg[i,:] = sum_{e in row i} m[e,:] (ascending edge order, plain fp32
 * adds: bit-identical to the CPU index_add_ of models/en_gnn_decoder.py:68-69) and
 * x_out = x + (sum_e w_e (x_i - x_col e)) * dinv_i * 0.2 (models/en_gnn_decoder.py:78-86).
 * dinv may be NULL (degree_inv=None).  agg / x_out may be NULL to skip that half. */
int pev_scatter_coord_fwd(const float* m, const float* w, const float* x, const float* dinv,
                          const int32_t* row_ptr, const int32_t* col, int64_t num_nodes, int32_t H,
                          float* agg, float* x_out, void* stream);
/* backward: gm[e,:] = gagg[row e,:]; gw[e] = 0.2 dinv_i (gxo_i . rel_e);
 * gx = gxo + sum_{row=i} w_e c_i - sum_{col=i} w_e c_row(e), c_i = 0.2 dinv_i gxo_i. */
int pev_scatter_coord_bwd(const float* gagg, const float* gxo, const float* w, const float* x,
                          const float* dinv, const int32_t* row_ptr, const int32_t* row,
                          const int32_t* col, const int32_t* col_ptr, const int32_t* csc_perm,
                          int64_t num_nodes, int64_t num_edges, int32_t H, float* gm, float* gw,
                          float* gx, void* stream);

/* ---------------------------------------------------------------- K1 (bf16 tensor-core form)
 * Fused edge MLP of EGNLayer.forward (models/en_gnn_decoder.py:60-79), H = 256
 * (csrc/edge_tc2_kernels.cu).  Conventions:
 *  - half domain: pre-activations are carried as h = z/2; weight images are packed with scale 0.5
 *    (pev_pack_weight_bf16_scaled), biases are halved inside the kernels, and the node projection is
 *    ABh = 0.5 [h Wa^T + b1 | h Wb^T] staged as fp16 [N,512] (11-bit significand: its rounding adds to the
 *    first edge linear's output; the GEMM operands a, m and the weights are bf16);
 *  - tile image: a per-edge [E,256] bf16 tensor stored per 128-edge tile as the 64 KB SWIZZLE_128B
 *    shared-memory image [fq 4][eh 2][fg 8][r 8][128 B] (feature 64 fq + 8 fg + r, edge 64 eh + 8 c + i
 *    in 16-byte chunk c ^ r); pev_edge2_tile_image_bytes(E) is the buffer size. */
int pev_pack_weight_bf16_scaled(const float* W /*[256,256] row-major (out,in)*/, int32_t transpose,
                                float scale, void* packed /*131072 bytes*/, void* stream);
int64_t pev_edge2_tile_image_bytes(int64_t num_edges);
/* d2[e] = |x[row[e]] - x[col[e]]|^2 (models/en_gnn_decoder.py:61-62). */
int pev_edge_d2(const float* x, const int32_t* row, const int32_t* col, int64_t num_edges,
                float* d2 /*[E]*/, void* stream);
/* fwd1: hu = Ah_i + Bh_j + (wd/2) d2, a = silu(2 hu), hv = a (W2/2)^T + b2/2, m = silu(2 hv);
 * writes the mT tile images (operand of fwd2 / wgrad5) and, when hvT != NULL (a backward pass follows),
 * the hvT tile images; agg[N,256] = segment_sum(m) (zeroed inside; one fp32 atomic per segment and feature). */
int pev_edge2_fwd1(const void* ABh /*fp16 [N,512]*/, const float* d2 /*[E]*/, const float* wd,
                   const void* W2hp, const float* b2, const int32_t* row, const int32_t* col,
                   int64_t num_nodes, int64_t num_edges, void* hvT /*tile images or NULL*/,
                   void* mT /*tile images*/, float* agg /*[N,256]*/, void* stream);
/* fwd2: hs = m (W5/2)^T + b5/2, t = silu(2 hs), w = t . w6 + b6 -> w[E] (zeroed inside);
 * hs_out (bf16 [E,256]) is written when not NULL (kept for the backward pass). */
int pev_edge2_fwd2(const void* mT, const void* W5hp, const float* b5, const float* w6,
                   const float* b6 /*[1]*/, int64_t num_edges, float* w_out /*[E]*/,
                   void* hs_out /*bf16 [E,256] or NULL*/, void* stream);

/* bwd2 (backward of fwd2 and of the aggregation; SURVEY.md 8a, half domain: d silu(2h)/dh = 1 + r(h)):
 * ghs = gw w6 (1 + r(hs)), gm = ghs (W5/2) + gagg[row], ghv = gm (1 + r(hv)) -> ghvT tile images;
 * db2h[256] = sum_e ghv (db2 = db2h / 2; per-CTA partials in `workspace`, summed in a fixed order: no atomics).
 * W5thp = pack(W5, transpose=1, scale=0.5); `workspace`: pev_edge2_wgrad_workspace_bytes(). */
int pev_edge2_bwd2(const void* hs /*bf16 [E,256]*/, const float* gw /*[E]*/, const float* w6,
                   const void* W5thp, const float* gagg /*[N,256]*/, const int32_t* row,
                   const void* hvT, int64_t num_edges, float* workspace, void* ghvT /*tile images*/,
                   float* db2h, void* stream);
/* bwd1 (backward of fwd1): ga = ghv (W2/2), ghu = ga (1 + r(hu)) with hu rebuilt from ABh, d2, wd;
 * writes ghu (bf16 [E,256] = dL/dhu) and gd2[e] = ghu . (wd/2): the four column quarters of an edge are written to
 * gd2_parts[4][E] (scratch) and summed in a fixed order.  W2thp = pack(W2, transpose=1, scale=0.5). */
int pev_edge2_bwd1(const void* ghvT, const void* W2thp, const void* ABh /*fp16 [N,512]*/,
                   const float* d2, const int32_t* row, const int32_t* col, const float* wd,
                   int64_t num_edges, void* ghu /*bf16 [E,256]*/, float* gd2 /*[E]*/,
                   float* gd2_parts /*[4,E] scratch*/, void* stream);

/* Segment sums of ghu over CSR rows / CSC columns (gA | gB = dL/dABh, fp32 [N,512]) and
 * gwdh[256] = sum_e d2[e] ghu[e] (dL/dwd = gwdh / 2): one warp per node, deterministic order; per-block partials
 * of gwdh in `workspace` (pev_edge2_wgrad_workspace_bytes()), summed in a fixed order. */
int pev_edge2_sums(const void* ghu /*bf16 [E,256]*/, const float* d2, const int32_t* row_ptr,
                   const int32_t* col_ptr, const int32_t* csc_perm, int64_t num_nodes, int64_t num_edges,
                   float* workspace, float* gAB /*[N,512]*/, float* gwdh /*[256]*/, void* stream);
/* gx[i] += sum over the edges at node i of +-2 gd2[e] (x_row - x_col): the coordinate part of
 * pev_edge_prologue_bwd on its own (d d2 / d x). */
int pev_edge_coord_bwd_accum(const float* gd2, const float* x, const int32_t* row_ptr, const int32_t* row,
                             const int32_t* col, const int32_t* col_ptr, const int32_t* csc_perm,
                             int64_t num_nodes, int64_t num_edges, float* gx_accum, void* stream);
/* Weight gradients of the two 256x256 edge linears as split-K tcgen05 GEMMs over the edge dimension; both
 * operands come from hs / mT / ghvT and the node projection (a is rebuilt on the fly).  `workspace`
 * (pev_edge2_wgrad_workspace_bytes(): one 256x256 fp32 partial per CTA + the per-CTA column-sum partials of bwd2 /
 * wgrad5 / sums) is shared by the backward kernels of a layer; all partials are summed in a fixed order.
 *   wgrad5: dW5[k,f] = sum_e gs[e,k] m[e,f] (full-domain gradient of phi_x.0.weight), db5[256] = sum_e ghs
 *           (half domain: db5 = result / 2), dw6[256] = sum_e gw silu(s)  (no atomics)
 *   wgrad2: dW2[f,j] = sum_e gv[e,f] a[e,j] (full-domain gradient of phi_e.2.weight) */
int64_t pev_edge2_wgrad_workspace_bytes(void);
int pev_edge2_wgrad5(const void* hs, const float* gw, const float* w6, const void* mT, int64_t num_edges,
                     float* workspace, float* dW5 /*[256,256]*/, float* db5h, float* dw6, void* stream);
int pev_edge2_wgrad2(const void* ghvT, const void* ABh, const float* d2, const int32_t* row,
                     const int32_t* col, const float* wd, int64_t num_edges, float* workspace,
                     float* dW2 /*[256,256]*/, void* stream);

/* ---------------------------------------------------------------- node-level kernels (bf16 path)
 * y = LayerNorm(x + res) (models/en_gnn_decoder.py:72-73; res may be NULL), fp32, one warp per row,
 * D = 256 or 512; also writes r = x + res (if r_out != NULL), mean[N], rstd[N] for the backward pass. */
int pev_add_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta,
                          float eps, int64_t N, int32_t D, float* r_out, float* y, float* mean,
                          float* rstd, void* stream);
/* One pass over gy and r: gr = dL/dr and the column sums dgamma[D], dbeta[D].  Column sums go through per-block
 * partials in `workspace` (pev_node_workspace_bytes()) and a fixed-order reduction: no atomics, bit-reproducible. */
int64_t pev_node_workspace_bytes(void);
int pev_layernorm_bwd(const float* gy, const float* r, const float* gamma, const float* mean,
                      const float* rstd, int64_t N, int32_t D, float* workspace, float* gr, float* dgamma,
                      float* dbeta, void* stream);
/* out[D] = sum over rows of g[N,D] (bias gradients of the node-level linears), D <= 4096; per-block partials in
 * `workspace` (pev_node_workspace_bytes()), fixed-order reduction. */
int pev_column_sum(const float* g, int64_t N, int32_t D, float* workspace, float* out, void* stream);

/* ---------------------------------------------------------------- K3: losses
 * Forward accumulators: acc_global[2*PEV_NUM_TERMS] doubles (numerator, denominator per term;
 * pre-zeroed) and acc_sample[B*8] doubles (per conformer: rec_ca, rec_n, rec_c numerators,
 * sum(mask), clash numerator, clash denominator, 2 spare; pre-zeroed).
 * pev_loss_finalize turns them into terms[PEV_NUM_TERMS] (float) and
 * inv_den[PEV_NUM_TERMS + 2B] (float: per-term 1/denominator, then per conformer 1/(B sum(mask_b)),
 * then per conformer 1/(B (clash pairs_b + 1e-8))) used by the backward kernels.  Any input pointer may be NULL: its terms are skipped. */
typedef struct pev_loss_args {
  const float* pred_N;      /* [B,L,3] */
  const float* pred_CA;
  const float* pred_C;
  const float* target_N;
  const float* target_CA;
  const float* target_C;
  const float* mask;        /* [B,L] float 0/1 */
  const float* target_dih;  /* [B,L,6] or NULL */
  const float* logits;      /* [B,L,C] or NULL */
  const int64_t* labels;    /* [B,L] */
  const float* mu_l;        /* [B,L,D] or NULL */
  const float* lv_l;
  const float* mu_g;        /* [B,G] or NULL */
  const float* lv_g;
  int32_t B, L, C, D, G;
  int32_t pair_stride;      /* >0 enables PEV_T_PAIR */
  int32_t enable_clash;
  int32_t enable_geometry;  /* bond / angle / dihedral-derived terms from pred_N/CA/C */
  float clash_dist;         /* 3.2 in the reference (models/losses.py:439) */
  float soft_margin;        /* 0.5 */
} pev_loss_args;

int pev_loss_fwd(const pev_loss_args* args_host, double* acc_global, double* acc_sample,
                 void* stream);
int pev_loss_finalize(const double* acc_global, const double* acc_sample, int32_t B,
                      float* terms, float* inv_den, void* stream);
/* coef[PEV_NUM_TERMS]: upstream gradient per base term (device).  Output gradients
 * (overwritten; any may be NULL): same shapes as the corresponding inputs. */
int pev_loss_bwd(const pev_loss_args* args_host, const float* coef, const float* inv_den,
                 float* g_pred_N, float* g_pred_CA, float* g_pred_C, float* g_logits,
                 float* g_mu_l, float* g_lv_l, float* g_mu_g, float* g_lv_g, void* stream);

/* compute_dihedrals_from_coords, models/losses.py:235-308: out[B,L,6]. */
int pev_dihedrals_fwd(const float* N, const float* CA, const float* C, const float* mask, int32_t B,
                      int32_t L, float* out, void* stream);
int pev_dihedrals_bwd(const float* N, const float* CA, const float* C, const float* mask,
                      const float* gout, int32_t B, int32_t L, float* gN, float* gCA, float* gC,
                      void* stream);
/* dihedral-space terms on an explicit [B,L,6] tensor: dihedral_consistency_loss,
 * ramachandran_loss, omega_trans_loss (models/losses.py:60-155).  sums[6] doubles, pre-zeroed:
 * cons num, cons den, rama num, omega num, sum(mask), spare.  target may be NULL. */
int pev_dihedral_terms_fwd(const float* dih, const float* target, const float* mask, int32_t B,
                           int32_t L, double* sums, void* stream);
/* coef3 = device float[3]: d(loss)/d(term) already divided by the term's denominator. */
int pev_dihedral_terms_bwd(const float* dih, const float* target, const float* mask,
                           const float* coef3, int32_t B, int32_t L, float* gdih, void* stream);

/* ---------------------------------------------------------------- K4: batched Kabsch RMSD
 * Replaces kabsch_rmsd, generate_ensemble_pdbs.py:343-373, for S conformers at once.
 * a[S,L,3]; b[S or 1,L,3] (b_batch = 0 broadcasts one reference structure); mask[S or 1,L]
 * float (NULL = all valid).  mode 0: optimal-superposition RMSD (scripts/validation_metrics.py:57-85);
 * mode 1: bit-for-intent reproduction of the reference's c1 @ R convention.
 * out[S]; 0.0 for an empty selection as in the reference (:350-351). */
int pev_kabsch_rmsd(const float* a, const float* b, const float* mask, int32_t S, int32_t L,
                    int32_t b_batch, int32_t mask_batch, int32_t mode, float* out, void* stream);
/* All pairs of one ensemble: out[S,S] (symmetric, zero diagonal) = Kabsch RMSD of a[i] vs a[j] for i < j -- the
 * diversity loop of generate_ensemble_pdbs.py:591-595 (S(S-1)/2 host calls there) in one launch; mask[L] or NULL. */
int pev_kabsch_rmsd_pairs(const float* a, const float* mask, int32_t S, int32_t L, int32_t mode,
                          float* out /*[S,S]*/, void* stream);

/* ---------------------------------------------------------------- generation driver: geometry validity filter
 * Replaces validate_protein_geometry, generate_ensemble_pdbs.py:290-340 (a Python loop over residues with one
 * device->host .item() per distance / angle), for S conformers at once: CA-CA distances and CA-CA-CA angles over the
 * compacted valid residues.  ca[S,L,3]; mask[S or 1,L] float or NULL.  status[S]: 0 valid, 1 no valid residues,
 * 2 extreme CA-CA distance (> 6.0), 3 abnormal average CA-CA distance (< 2.5 or > 5.0), 4 abnormal average CA-CA-CA
 * angle (< 60 or > 180 degrees); stats[S,3] = {max distance, mean distance, mean angle} (may be NULL). */
int pev_validate_geometry(const float* ca, const float* mask, int32_t S, int32_t L, int32_t mask_batch,
                          int32_t* status, float* stats, void* stream);

/* ---------------------------------------------------------------- evaluation metrics
 * scripts/validation_metrics.py for S structures at once.  pev_superpose_scores: kabsch_align (:57-85, over all L
 * residues) of a[s] onto b[s or 0]; dist[S,L] = per-residue distance after superposition, aligned[S,L,3] (or NULL),
 * tm[S] = compute_tm_score_python (:23-54), gdt_ts / gdt_ha[S] = compute_gdt (:156-199, masked, percent) (each may be
 * NULL).  pev_lddt: compute_lddt (:92-149): per_residue[S,L] and global[S].  pev_rmsf: compute_rmsf (:206-241) of an
 * ensemble already aligned to its first member (pev_superpose_scores with b = a[0]): out[L]. */
int pev_superpose_scores(const float* a, const float* b, const float* mask, int32_t S, int32_t L, int32_t b_batch,
                         int32_t mask_batch, float* aligned, float* dist, float* tm, float* gdt_ts, float* gdt_ha,
                         void* stream);
int pev_lddt(const float* pred, const float* tru, const float* mask, int32_t S, int32_t L, int32_t t_batch,
             int32_t mask_batch, float cutoff, float* per_residue, float* global, void* stream);
int pev_rmsf(const float* aligned, int32_t N, int32_t L, float* out, void* stream);

/* ---------------------------------------------------------------- K1, fused CTA-pair form (csrc/edge_tc3_kernels.cu)
 * One kernel for the whole forward edge MLP of EGNLayer.forward (models/en_gnn_decoder.py:60-79): phi_e[1..3] ->
 * agg = index_add_(m) (:68-69) and phi_x (:76), with m kept on chip between the two 256 x 256 GEMMs.  Run by CTA
 * pairs (tcgen05 cta_group::2, each CTA holds one half of both weights).  Inputs as pev_edge2_fwd1 / fwd2 (ABh fp16
 * half-domain node projection, d2, packed 0.5 W2 / 0.5 W5 images from pev_pack_weight_bf16_scaled).  Outputs:
 * agg[N,256] (zeroed inside), w[E]; when hv_rows / hs_rows != NULL (a backward pass follows; both or neither) the
 * half-domain pre-activations hv = v/2 and hs = s/2 as plain bf16 rows [E,256]. */
int pev_edge3_fwd(const void* ABh /*fp16 [N,512]*/, const float* d2 /*[E]*/, const float* wd /*[256]*/,
                  const void* W2hp, const float* b2 /*[256]*/, const void* W5hp, const float* b5 /*[256]*/,
                  const float* w6 /*[256]*/, const float* b6 /*[1]*/, const int32_t* row, const int32_t* col,
                  int64_t num_nodes, int64_t num_edges, void* hv_rows /*bf16 [E,256] or NULL*/,
                  void* hs_rows /*bf16 [E,256] or NULL*/, float* agg /*[N,256]*/, float* w /*[E]*/, void* stream);

/* ---------------------------------------------------------------- node-level linears on tcgen05 (csrc/node_gemm_kernels.cu)
 * C[M,Nout] = epilogue(A W^T), A = [A1 | A2] (fp32 row-major [M,K1], [M,K2]; A2 may be NULL), W fp32 [Nout, K1+K2]
 * row-major, operands read as TF32 with fp32 accumulation.  Replaces the node-level nn.Linear calls of
 * EGNLayer.forward (models/en_gnn_decoder.py:65-73) and the element-wise kernels around them:
 *   epilogue 0 ABH     out = fp16(scale * (C + bias))                 h-halves of phi_e[0] (:65-66), scale = 0.5
 *            1 SILU    out = silu(C + bias), out2 = C + bias (or NULL)  phi_h[0..1] on [h | agg] (:70-71)
 *            2 RES_LN  r = C + bias + aux; out = LayerNorm(r) (gamma, beta, eps); out2 = r, mean, rstd (or NULL)  (:71-73)
 *            3 PLAIN   out = C + bias (+ aux)                          data gradients
 *            4 DSILU   out = (C + bias) * silu'(aux)                   backward through phi_h[1]
 * Nout in {256, 512}; K1, K2 multiples of 32; aux / out / out2 row-major with leading dimension Nout (RES_LN: 256). */
/* Weight gradients of the node-level linears: out[Mo, 256] (leading dimension ldc >= 256) = scale * G^T X with G fp32
 * [N, Mo] (Mo = 256 or 512), X fp32 [N, 256]: split-K over row slices on tcgen05 (TF32, MN-major operands), per-CTA
 * partials in `workspace` (pev_node_wgrad_workspace_bytes()) summed in a fixed order. */
int64_t pev_node_wgrad_workspace_bytes(void);
int pev_node_wgrad(const float* G, int32_t Mo, const float* X, int64_t N, float scale, float* workspace, float* out,
                   int32_t ldc, void* stream);
int pev_node_gemm(int32_t epilogue, const float* A1, int32_t K1, const float* A2, int32_t K2, const float* W,
                  const float* bias, int64_t M, int32_t Nout, float scale, const float* aux, const float* gamma,
                  const float* beta, float eps, void* out, float* out2, float* mean, float* rstd, void* stream);

/* fp32-accurate forms of the two calls above for the exact path (precision="fp32"): every product is the 3xTF32 sum
 * hi lo + lo hi + hi hi accumulated in fp32 (error ~2^-22 relative), operands split on chip.  Replace the fp32 nn.Linear
 * calls of EGNLayer.forward (models/en_gnn_decoder.py:61-79: phi_e[0] node halves, phi_e[2], phi_x[0], phi_h) and their
 * autograd GEMMs.  pev_split_tf32 writes the split image [hi ; lo] ([2 R, C]) of W [rows, cols] (transpose = 0) or of W^T
 * (transpose = 1); pev_node_gemm3: out[M,Nout] = A[M,K] W^T (+ bias) (+ res), W3 = split image of W [Nout,K];
 * pev_node_wgrad3: arguments as pev_node_wgrad (same workspace size). */
int pev_split_tf32(const float* W, int32_t rows, int32_t cols, int32_t transpose, float* out, void* stream);
int pev_node_gemm3(const float* A, int32_t K, const float* W3, const float* bias, int64_t M, int32_t Nout,
                   const float* res, float* out, void* stream);
int pev_node_wgrad3(const float* G, int32_t Mo, const float* X, int64_t N, float scale, float* workspace, float* out,
                    int32_t ldc, void* stream);

/* General linear layer on the same kernels (used by the encoder, models/encoder.py:30-262): out[M,Nout] (leading dimension
 * ldc) = dropout(act(A[M,K] W^T + bias)) + res, act = ReLU when relu != 0, dropout with probability p_drop on a counter
 * hash of (seed, row * Nout + col) (p_drop = 0: none); A / res / out may be column blocks of wider row-major
 * tensors (lda / ldres / ldc, multiples of 4).  precise = 0: TF32, W fp32 [Nout,K], Nout a multiple of 256; precise = 1:
 * 3xTF32, W = pev_split_tf32 image [2 Nout, K], Nout a multiple of 128.  K a multiple of 32.
 * pev_linear_wgrad: out[Mo,Kx] (leading dimension ldc) = scale * G^T X for G [N,Mo] (ldg) and X [N,Kx] (ldx), Mo a multiple
 * of 128, Kx of 256, with ceil(Mo / 256)(Kx / 256) <= the SM count, in one launch; workspace of pev_node_wgrad_workspace_bytes(). */
int pev_linear(int32_t precise, const float* A, int64_t lda, int32_t K, const float* W, const float* bias, int64_t M,
               int32_t Nout, int32_t relu, float p_drop, uint32_t seed, const float* res, int64_t ldres, float* out, int64_t ldc,
               void* stream);
/* Backward of those epilogues: out[i] = (keep(seed, i) && (y == NULL || y[i] != 0)) ? g[i] / (1 - p_drop) : 0 over n
 * contiguous elements (n a multiple of 4) -- nn.Dropout / nn.ReLU backward of models/encoder.py's blocks in one pass. */
int pev_mask_grad(const float* g, const float* y, int64_t n, float p_drop, uint32_t seed, float* out, void* stream);
/* out[cols, rows] = scale * in[rows, cols]^T (the pre-transposed weight of a data-gradient GEMM: g W = g (W^T)^T). */
int pev_transpose(const float* in, int32_t rows, int32_t cols, float scale, float* out, void* stream);
int pev_linear_wgrad(int32_t precise, const float* G, int64_t ldg, int32_t Mo, const float* X, int64_t ldx, int32_t Kx,
                     int64_t N, float scale, float* workspace, float* out, int64_t ldc, void* stream);

/* ---------------------------------------------------------------- encoder attention (csrc/attn_kernels.cu)
 * Per-conformer multi-head attention over packed rows (models/encoder.py:125-140, nn.MultiheadAttention with
 * key_padding_mask) as batched TF32 GEMMs on tcgen05.  The score buffer is [H, Np, Lpad] fp32: conformer b owns padded rows
 * cup[b] .. cup[b+1] (multiples of 128; m_tiles = Np / 128, tile_conf[t] = conformer of tile t), Lpad a multiple of 32
 * >= the longest conformer.  pev_attn_gemm forms: 0 out = buffer = scale * A B^T (A, B packed [N, lda / ldb], operand
 * columns a_col0 / b_col0 + h * hd); 1 out[N, ldo] (columns out_col0 + h * hd ..) = scale * P B with A = buffer;
 * 2 out = scale * P^T B.  pev_attn_softmax: backward = 0: rows of S -> probabilities in place (zeros at padding), G
 * (optional) = dropout-kept probabilities / (1 - p_drop); backward = 1: G (= dP) -> dS = P (dP - sum P dP) in place, the
 * dropout mask re-derived from (seed, element index). */
int pev_attn_gemm(int32_t form, const float* A, int64_t lda, int32_t a_col0, const float* B, int64_t ldb, int32_t b_col0,
                  const int32_t* tile_conf, const int32_t* cu, const int32_t* cup, int32_t m_tiles, int32_t H, int32_t hd,
                  int32_t Lpad, int64_t N, float scale, float* out, int64_t ldo, int32_t out_col0, void* stream);
int pev_attn_softmax(int32_t backward, float* S, float* G, const int32_t* tile_conf, const int32_t* cu, const int32_t* cup,
                     int32_t m_tiles, int32_t H, int32_t Lpad, float p_drop, uint32_t seed, void* stream);
/* Form 0 with the softmax fused into the GEMM epilogue (the scores never reach HBM): mode 1 (Lpad <= 256) out = probabilities,
 * out2 = dropout-kept probabilities / (1 - p_drop) or NULL; mode 2 (A = dO, B = V) out = dS = P (keep dP / (1 - p_drop) -
 * delta) with delta[h Np + padded row] = <dO, O> of that row and head (pev_attn_delta; row_pad[i] = padded row of packed row i). */
int pev_attn_scores(int32_t mode, const float* A, int64_t lda, int32_t a_col0, const float* B, int64_t ldb, int32_t b_col0,
                    const int32_t* tile_conf, const int32_t* cu, const int32_t* cup, int32_t m_tiles, int32_t H, int32_t hd,
                    int32_t Lpad, int64_t N, float scale, float* out, float* out2, const float* P, const float* delta,
                    float p_drop, uint32_t seed, void* stream);
int pev_attn_delta(const float* dO, const float* O, int64_t N, int32_t H, int32_t hd, const int32_t* row_pad, int64_t Np,
                   float* delta, void* stream);

/* ---------------------------------------------------------------- ragged packed batches (csrc/data_kernels.cu)
 * Device-side replacement for the host centring + zero-padding of models/data.py (:166-172, :219-266): the packed rows of
 * B conformers (n / ca / c [T,3], mask [T], dih [T,6], labels [T] int64, emb [T,D] or NULL; conformer b = rows
 * cu_seqlens[b] .. cu_seqlens[b+1]) -> padded [B,Lmax,...] tensors, coordinates centred on each conformer's valid-CA
 * centroid when center != 0, exact zeros past a conformer's length. */
int pev_unpack_center(const float* n, const float* ca, const float* c, const float* mask, const float* dih,
                      const int64_t* labels, const float* emb, const int32_t* cu_seqlens, int32_t B, int32_t Lmax, int32_t D,
                      int32_t center, float* o_n, float* o_ca, float* o_c, float* o_mask, float* o_dih, int64_t* o_labels,
                      float* o_emb, void* stream);

/* ---------------------------------------------------------------- backbone placement (csrc/backbone_kernels.cu)
 * N / C placement and the 3-step peptide pull of EGNNDecoder.forward (models/en_gnn_decoder.py:260-310) over the packed
 * residues: x_n = pull3(x_ca + 1.46 normalize(n_dir)), x_c = x_ca + 1.52 normalize(c_dir); n_dir / c_dir are the first three
 * channels of the offset heads' outputs (rows of ldn / ldc floats), starts[k] != 0 marks the first residue of a conformer (not
 * pulled).  pev_backbone_bwd: gradients with respect to the directions ([N,3] each) and x_ca from g_xn, g_xc. */
int pev_backbone_fwd(const float* n_dir, int32_t ldn, const float* c_dir, int32_t ldc, const float* x_ca, const uint8_t* starts,
                     int64_t N, float* x_n, float* x_c, void* stream);
int pev_backbone_bwd(const float* n_dir, int32_t ldn, const float* c_dir, int32_t ldc, const float* x_ca, const uint8_t* starts,
                     int64_t N, const float* g_xn, const float* g_xc, float* g_ndir, float* g_cdir, float* g_xca, void* stream);

/* ---------------------------------------------------------------- streaming ensemble PDB writer (csrc/pdb_kernels.cu)
 * The MODEL blocks that write_pdb (generate_ensemble_pdbs.py:148-288, with compute_backbone_oxygen :106-144) appends one
 * call per model: MODEL line, four ATOM records per valid residue (N, CA, C and the placed O), blank line, CONECT records,
 * TER, ENDMDL -- byte for byte, for models model0 .. model0 + S - 1 (1-based) of n / ca / c [S,L,3] into `out`
 * (pev_pdb_models_bytes(model0, S, nv) bytes).  valid_idx[nv]: indices of the residues with mask > 0.5; prev_ok[nv]: the
 * residue before it exists and is valid; resname[3 nv]: three-letter codes.  *overflow is set to 1 if a coordinate does not
 * fit %8.3f (|x| >= 9999.9995). */
int64_t pev_pdb_models_bytes(int64_t model0, int32_t S, int32_t nv);
int pev_pdb_format_models(const float* n, const float* ca, const float* c, const int32_t* valid_idx, const uint8_t* prev_ok,
                          const char* resname, int32_t S, int32_t L, int32_t nv, int64_t model0, int32_t chain, char* out,
                          int32_t* overflow, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PEV_B200_H */
