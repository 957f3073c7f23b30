// PDB text of one residue of one model -- the per-thread body of the streaming ensemble writer (SURVEY.md 8f, N2;
// generate_ensemble_pdbs.py:106-144 compute_backbone_oxygen, :148-288 write_pdb).  Host/device (tests/hostcheck).
// Every record has a fixed width, so the byte offset of any line of any model is a closed form of (model number, rank of
// the residue among the valid ones): all lines of an ensemble are formatted in parallel.
#pragma once
#include <math.h>
#include <stdint.h>

#include "pev_hd.cuh"

namespace pev {

constexpr int kPdbAtomLine = 81;     // "ATOM  %5d  N   RES C%4d    %8.3f%8.3f%8.3f  1.00  0.00           N  \n"
constexpr int kPdbConectLine = 17;   // "CONECT%5d%5d\n"

PEV_HD int pdb_digits(int64_t v) {
  int d = 1;
  while (v >= 10) { v /= 10; ++d; }
  return d;
}
// "MODEL     %4d\n": 15 bytes up to model 9999, one more per extra digit (Python widens the field)
PEV_HD int pdb_model_line_len(int64_t m) {
  const int d = pdb_digits(m);
  return 11 + (d > 4 ? d : 4);
}
// sum over models 1..m of the bytes by which their MODEL line exceeds 15
PEV_HD int64_t pdb_extra_bytes_upto(int64_t m) {
  int64_t extra = 0, lo = 10000;
  for (int d = 5; lo <= m; ++d, lo *= 10) {
    const int64_t hi = lo * 10 - 1 < m ? lo * 10 - 1 : m;
    extra += (int64_t)(d - 4) * (hi - lo + 1);
  }
  return extra;
}
// bytes of one MODEL block without the widening of its MODEL line
PEV_HD int64_t pdb_block_bytes(int nv) {
  return 15 + (int64_t)nv * 4 * kPdbAtomLine + 1 + (nv > 0 ? (int64_t)kPdbConectLine * (4 * nv - 1) : 0) + 4 + 7;
}
// first byte of the block of model number m (>= first), counted from the block of model `first`
PEV_HD int64_t pdb_block_offset(int64_t m, int64_t first, int nv) {
  return (m - first) * pdb_block_bytes(nv) + pdb_extra_bytes_upto(m - 1) - pdb_extra_bytes_upto(first - 1);
}

PEV_HD void pdb_put_int(char* p, int width, int64_t v) {           // right-aligned, blank-padded ("%{width}d", v >= 0)
  for (int k = width - 1; k >= 0; --k) {
    p[k] = (v > 0 || k == width - 1) ? (char)('0' + v % 10) : ' ';
    v /= 10;
  }
}
// "%8.3f" of a float32 exactly as Python formats it: x * 1000 is exact in double (24 + 10 significant bits), llrint rounds
// half to even like the correctly rounded decimal conversion; the sign of a negative value survives rounding to zero.
// Returns false when the value does not fit the field (|x| >= 9999.9995 or not finite).
PEV_HD bool pdb_put_f83(char* p, float x) {
  const double v = (double)x * 1000.0;
  if (!(fabs(v) < 9999999.5)) return false;
  int64_t q = llrint(fabs(v));
  const bool neg = signbit(x);
  if (neg && q >= 1000000) return false;                            // "-1000.000" is 9 characters
  char buf[8];
  buf[7] = (char)('0' + q % 10); q /= 10;
  buf[6] = (char)('0' + q % 10); q /= 10;
  buf[5] = (char)('0' + q % 10); q /= 10;
  buf[4] = '.';
  int k = 3;
  do { buf[k--] = (char)('0' + q % 10); q /= 10; } while (q > 0 && k >= 0);
  if (neg) buf[k--] = '-';
  while (k >= 0) buf[k--] = ' ';
  for (int j = 0; j < 8; ++j) p[j] = buf[j];
  return true;
}

// O position (:106-144) in the reference's own arithmetic: float32 throughout when the previous residue is valid (unit
// vector CA(i-1) -> C(i-1), 1.23 A from C), float64 sum rounded to float32 along +x otherwise
PEV_HD void pdb_oxygen(const float* ca_prev, const float* c_prev, const float* c, bool prev_ok, float* o) {
  if (!prev_ok) {
    o[0] = (float)((double)c[0] + 1.23);
    o[1] = c[1];
    o[2] = c[2];
    return;
  }
  const float r0 = c_prev[0] - ca_prev[0], r1 = c_prev[1] - ca_prev[1], r2 = c_prev[2] - ca_prev[2];
  float s = r0 * r0;
  s += r1 * r1;
  s += r2 * r2;
  const float den = sqrtf(s) + 1e-8f;
  o[0] = c[0] + (r0 / den) * 1.23f;
  o[1] = c[1] + (r1 / den) * 1.23f;
  o[2] = c[2] + (r2 / den) * 1.23f;
}

PEV_HD bool pdb_atom_line(char* p, int atom_num, const char* name6, const char* res3, char chain, int resnum, const float* x,
                          char element) {
  const char* a = "ATOM  ";
  for (int k = 0; k < 6; ++k) p[k] = a[k];
  pdb_put_int(p + 6, 5, atom_num);
  for (int k = 0; k < 6; ++k) p[11 + k] = name6[k];
  p[17] = res3[0]; p[18] = res3[1]; p[19] = res3[2];
  p[20] = ' ';
  p[21] = chain;
  pdb_put_int(p + 22, 4, resnum);
  p[26] = p[27] = p[28] = p[29] = ' ';
  bool ok = pdb_put_f83(p + 30, x[0]);
  ok = pdb_put_f83(p + 38, x[1]) && ok;
  ok = pdb_put_f83(p + 46, x[2]) && ok;
  const char* t = "  1.00  0.00           ";
  for (int k = 0; k < 23; ++k) p[54 + k] = t[k];
  p[77] = element; p[78] = ' '; p[79] = ' '; p[80] = '\n';
  return ok;
}

PEV_HD void pdb_conect_line(char* p, int a, int b) {
  const char* t = "CONECT";
  for (int k = 0; k < 6; ++k) p[k] = t[k];
  pdb_put_int(p + 6, 5, a);
  pdb_put_int(p + 11, 5, b);
  p[16] = '\n';
}

struct PdbArgs {
  const float* n;            // [S, L, 3]
  const float* ca;
  const float* c;
  const int32_t* valid_idx;  // [nv] residue index of the k-th valid residue
  const uint8_t* prev_ok;    // [nv] residue valid_idx[k] - 1 exists and is valid
  const char* resname;       // [nv * 3]
  int32_t S, L, nv;
  int64_t model0;            // number of the first model of this call (1-based); `first` of the file for the offsets
  char chain;
  char* out;                 // blocks of models model0 .. model0 + S - 1, contiguous
};

// everything residue k (k-th valid) of model s contributes; k == 0 also writes the MODEL / blank / TER / ENDMDL lines.
// Returns false if a coordinate did not fit its field.
PEV_HD bool pdb_residue(const PdbArgs& a, int s, int k) {
  const int64_t m = a.model0 + s;
  char* blk = a.out + pdb_block_offset(m, a.model0, a.nv);
  const int mlen = pdb_model_line_len(m);
  bool ok = true;
  if (k == 0) {
    const char* t = "MODEL     ";
    for (int j = 0; j < 10; ++j) blk[j] = t[j];
    pdb_put_int(blk + 10, mlen - 11, m);
    blk[mlen - 1] = '\n';
    char* tail = blk + mlen + (int64_t)a.nv * 4 * kPdbAtomLine;
    tail[0] = '\n';
    char* end = tail + 1 + (a.nv > 0 ? (int64_t)kPdbConectLine * (4 * a.nv - 1) : 0);
    const char* e = "TER\nENDMDL\n";
    for (int j = 0; j < 11; ++j) end[j] = e[j];
    if (a.nv == 0) return true;
  }
  const int i = a.valid_idx[k];
  const int64_t base = ((int64_t)s * a.L + i) * 3;
  float o[3];
  pdb_oxygen(a.ca + base - 3, a.c + base - 3, a.c + base, a.prev_ok[k] != 0, o);
  char* p = blk + mlen + (int64_t)k * 4 * kPdbAtomLine;
  const char* res = a.resname + 3 * k;
  const int an = 4 * k + 1;
  ok = pdb_atom_line(p, an, "  N   ", res, a.chain, i + 1, a.n + base, 'N') && ok;
  ok = pdb_atom_line(p + kPdbAtomLine, an + 1, "  CA  ", res, a.chain, i + 1, a.ca + base, 'C') && ok;
  ok = pdb_atom_line(p + 2 * kPdbAtomLine, an + 2, "  C   ", res, a.chain, i + 1, a.c + base, 'C') && ok;
  ok = pdb_atom_line(p + 3 * kPdbAtomLine, an + 3, "  O   ", res, a.chain, i + 1, o, 'O') && ok;
  char* q = blk + mlen + (int64_t)a.nv * 4 * kPdbAtomLine + 1 + (int64_t)kPdbConectLine * 4 * k;
  pdb_conect_line(q, an, an + 1);
  pdb_conect_line(q + kPdbConectLine, an + 1, an + 2);
  pdb_conect_line(q + 2 * kPdbConectLine, an + 2, an + 3);
  if (k < a.nv - 1) pdb_conect_line(q + 3 * kPdbConectLine, an + 2, an + 4);      // peptide bond C(k) - N(k+1)
  return ok;
}

}  // namespace pev
