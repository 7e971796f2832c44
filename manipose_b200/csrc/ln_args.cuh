// Arguments of the residual Linear + LayerNorm epilogues (gemm2.cu: pair_linear_ln_kernel, pair_linear_ln64_kernel; mlp.cu).
#pragma once

namespace mp {

struct LnArgs {
  const float* bias;
  const float* post_g;   // NULL: no post-norm
  const float* post_b;
  const float* pos;      // NULL or [pos_mod, 512]
  const float* ln_g;     // NULL: no pre-norm / h output
  const float* ln_b;
  int has_xpre;             // with the post-norm: also store the value BEFORE it (tm_p; the training tape needs both)
  const float* row_scale;   // NULL or [M]: x = resid + row_scale[row] * (A W^T + bias) (per-sample DropPath factor)
  int no_x;                 // with post-norm AND pre-norm: do not store x_out (only h is wanted: the last block ahead of the heads)
  float post_eps, ln_eps;
  int pos_div, pos_mod;
};

}  // namespace mp
