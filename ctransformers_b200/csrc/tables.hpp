// Host-side scalar functions behind the fp16 lookup tables (reference: ggml.c:3556-3558 gelu, 3610-3612 silu, table
// construction 4319-4333).  Written with explicit fmaf()/separate operations so the values do not depend on what the
// host compiler decides to contract: the reference BINARY fuses (GELU_COEF_A*x)*x + 1.0f into one FMA (checked in the
// disassembly of ggml_init) and nothing else in these expressions can be fused.
#pragma once
#include <cmath>

namespace ctb {

inline float host_silu(float x) { return x / (1.0f + expf(-x)); }

inline float host_gelu(float x) {
  const float t = 0.044715f * x;
  const float inner = fmaf(t, x, 1.0f);
  const float a = 0.79788456080286535587989211986876f * x;
  const float arg = a * inner;
  const float h = 0.5f * x;
  const float u = 1.0f + tanhf(arg);
  return h * u;
}

}  // namespace ctb
