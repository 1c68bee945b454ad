// extern "C" surface of libsam3b.so (declared in include/sam3b.h).
#include "../../include/sam3b.h"

#include "common.h"
#include "gemm.cuh"

using namespace sam3b;

extern "C" {

const char* sam3b_last_error(void) { return last_error_message(); }
int sam3b_abi_version(void) { return SAM3B_ABI_VERSION; }

int sam3b_gemm(const sam3b_gemm_desc* d, void* stream) {
  if (!d) return fail(-1, "sam3b_gemm: null descriptor");
  GemmArgs a;
  a.M = d->M; a.N = d->N; a.K = d->K;
  a.A = d->A; a.lda = d->lda; a.a_mn = d->a_mn;
  a.B = d->B; a.ldb = d->ldb; a.b_mn = d->b_mn;
  a.dtype = d->dtype; a.epilogue = d->epilogue;
  a.C = d->C; a.ldc = d->ldc; a.C2 = d->C2; a.ldc2 = d->ldc2;
  a.bias = d->bias;
  a.residual = d->residual; a.ldres = d->ldres; a.res_row_mod = d->res_row_mod;
  a.aux = d->aux; a.ldaux = d->ldaux;
  a.rope = d->rope; a.rope_period = d->rope_period; a.rope_cols = d->rope_cols;
  a.alpha = d->alpha; a.splitk = d->splitk; a.c_trans = d->c_trans; a.bn = d->bn;
  a.dbg_lbo = d->dbg_lbo; a.dbg_sbo = d->dbg_sbo; a.max_ctas = d->max_ctas;
  return gemm_launch(a, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
