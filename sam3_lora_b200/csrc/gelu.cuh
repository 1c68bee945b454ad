// Exact-erf GELU and its derivative as device functions shared by the GEMM epilogues (gemm.cu) and the neck kernels
// (conv.cu).  Needs ex2_approx from ptx.cuh.
#pragma once
#include "ptx.cuh"

namespace sam3b {

// Exact-erf GELU (nn.GELU default, timm Mlp): gelu(h) = h * Phi(h), Phi = standard normal CDF.
// Forward, one MUFU + 10 FMA/ALU ops per element:
//   q = Phi(-|h|) = exp2(P(a)),  a = min(|h|, 6),  P = degree-7 fit of log2 Phi(-a)   (Phi(-6) = 1e-9)
//   gelu(h) = max(h, 0) - |h| * q
// Backward, one MUFU + 15 ops:
//   g = exp(-h^2/2),  S(a) = Phi(-a)/g - a/sqrt(2 pi)  (degree-8 fit, error weighted by g)
//   gelu'(h) = h < 0 ? g*S : 1 - g*S
// Constants and their float32 error check come from tools/fit_gelu.py: |gelu error| <= 3e-7, |gelu' error| <= 1.1e-6
// (the A&S 7.1.26 rational form used before had the same accuracy at 2 MUFU + 15 ops; libdevice erff costs 30-45).
// The epilogue, not the MMA, paces the fc1 / fc2-dgrad GEMMs (profiles/r01_ncu_v3_gemm2_gelu.txt), so ops count.
__device__ __forceinline__ float gelu_erf(float h) {
  const float a = fminf(fabsf(h), 6.f);
  float p = fmaf(-1.752960928e-06f, a, 6.019484742e-05f);
  p = fmaf(p, a, -9.220063879e-04f);
  p = fmaf(p, a, 8.487955181e-03f);
  p = fmaf(p, a, -5.395627800e-02f);
  p = fmaf(p, a, -4.584195313e-01f);
  p = fmaf(p, a, -1.151296121e+00f);
  p = fmaf(p, a, -9.999863347e-01f);
  return fmaf(-fabsf(h), ex2_approx(p), fmaxf(h, 0.f));
}
__device__ __forceinline__ float dgelu_erf(float h) {
  const float ah = fabsf(h);
  const float a = fminf(ah, 6.f);
  const float ap = ah * 0.84932180028801904f;   // sqrt(log2(e) / 2): g = exp2(-ap^2)
  const float g = ex2_approx(-(ap * ap));
  float s = fmaf(3.790287705e-05f, a, -6.153799106e-04f);
  s = fmaf(s, a, 4.399772528e-03f);
  s = fmaf(s, a, -1.882489903e-02f);
  s = fmaf(s, a, 5.614903928e-02f);
  s = fmaf(s, a, -1.299720504e-01f);
  s = fmaf(s, a, 2.492791102e-01f);
  s = fmaf(s, a, -7.978181042e-01f);
  s = fmaf(s, a, 4.999989938e-01f);
  const float gs = g * s;
  return h < 0.f ? gs : 1.f - gs;
}

}  // namespace sam3b
