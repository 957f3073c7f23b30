// Streaming ensemble PDB writer, device side: the MODEL blocks of generate_ensemble_pdbs.py:148-288
// (write_pdb, called once per model by the reference, :598-625) for a whole chunk of models at once.  All records are fixed
// width, so every (model, residue) thread knows its byte offsets; the text is produced at memory speed on the device and
// leaves as one contiguous buffer per chunk (protein_ensemble_vae_b200/generation.py streams the chunks to the file).
#include "../../include/pev_b200.h"
#include "pev_common.cuh"
#include "pev_pdb_body.cuh"

namespace pev {

__global__ void __launch_bounds__(128) pdb_format_kernel(const PdbArgs a, int32_t* __restrict__ overflow) {
  const int per = a.nv > 0 ? a.nv : 1;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)a.S * per) return;
  if (!pdb_residue(a, (int)(t / per), (int)(t % per))) atomicExch(overflow, 1);
}

}  // namespace pev

using namespace pev;

extern "C" int64_t pev_pdb_models_bytes(int64_t model0, int32_t S, int32_t nv) {
  return pdb_block_offset(model0 + S, model0, nv);
}

extern "C" int pev_pdb_format_models(const float* n, const float* ca, const float* c, const int32_t* valid_idx,
                                     const uint8_t* prev_ok, const char* resname, int32_t S, int32_t L, int32_t nv, int64_t model0,
                                     int32_t chain, char* out, int32_t* overflow, void* stream) {
  PEV_REQUIRE(n && ca && c && out && overflow && S >= 0 && L >= 0 && nv >= 0 && nv <= L && model0 >= 1, "bad argument");
  PEV_REQUIRE(nv == 0 || (valid_idx && prev_ok && resname), "null argument");
  PEV_REQUIRE(4 * (int64_t)nv <= 99999 && L <= 9999, "atom / residue numbers must fit their 5 / 4 character fields");
  if (S == 0) return 0;
  PdbArgs a = {n, ca, c, valid_idx, prev_ok, resname, S, L, nv, model0, (char)chain, out};
  const int64_t threads = (int64_t)S * (nv > 0 ? nv : 1);
  pdb_format_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, as_stream(stream)>>>(a, overflow);
  return after_launch("pdb_format_kernel");
}
