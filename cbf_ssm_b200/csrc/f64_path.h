// Float64 batched path (f64_path.cu): host entry points used by api.cu.
#pragma once
#include <vector>

#include "common.cuh"

namespace cbf {

struct F64Args {
  Dims D;
  int dx, du, dy;
  const std::vector<Chain> *chains;          // live message chains (empty for CBFSSMHALF)
  const double *state_f, *state_b;           // float64 prologue states of the two GPs (cbf_gp_prologue)
  const float *var_x, *var_y, *u, *y, *eps_b, *z_b, *eps_f;
  Workspace ws;                              // X, H, Yb, stats, x0, x0b are used
  void *scratch;                             // f64_scratch_bytes() bytes
  cudaStream_t stream;
};

// Destinations inside the flat float64 kernel-level gradient (cbf_grad_layout)
struct F64Grad {
  double *base;
  long long total;
  double *f_P, *f_alpha, *f_S, *f_Z, *f_ell, *f_sig2;
  double *b_P, *b_alpha, *b_S, *b_Z, *b_ell, *b_sig2;
  double *var_x, *var_y;
};

size_t f64_scratch_bytes(int n_local, int T, int M, int dx, int dy, int din);
int f64_forward(const F64Args &a, double *terms);
int f64_backward(const F64Args &a, double w_ll, double w_kl, double w_en, const F64Grad &g);

}  // namespace cbf
