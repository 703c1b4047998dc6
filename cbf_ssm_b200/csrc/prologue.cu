// Parameter-only prologue of one sparse GP and its adjoint, float64, one CTA per GP.
//   forward : softplus constraints (tf_transform.py:19-21), K_zz (gp_tf.py:33-49),
//             Cholesky with 1e-8 jitter (gp_tf.py:52-65,129-130), P = (K_zz+1e-8 I)^-1,
//             alpha = P m, KL(q(u)||p(u)) (gp_tf.py:163-172).
//   backward: SURVEY 8a note 4 tail; checked in float64 by oracle/kernel_math.py
//             (gp_prologue / gp_prologue_adjoint).
// O(M^3) once per step; latency-bound, so a single 512-thread CTA with shared-memory
// staging of the M x M factors is the whole design.
#include "common.cuh"

namespace cbf {

__device__ __forceinline__ double softplus_d(double x) {
  return (x > 0.0 ? x + log1p(exp(-x)) : log1p(exp(x))) + 1e-10;
}
__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

__device__ double block_sum_d(double v, double *sh) {
  const int tid = threadIdx.x;
  __syncthreads();
  sh[tid] = v;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (tid < o) sh[tid] += sh[tid + o];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(512) gp_prologue_kernel(int M, int Din, int Dout, const double *__restrict__ Z,
                                                          const double *__restrict__ mean,
                                                          const double *__restrict__ Su,
                                                          const double *__restrict__ vu,
                                                          const double *__restrict__ lu, float *__restrict__ Z32,
                                                          float *__restrict__ ell32, float *__restrict__ sig232,
                                                          float *__restrict__ P32, float *__restrict__ alpha32,
                                                          float *__restrict__ S32, double *__restrict__ kl_out,
                                                          double *__restrict__ st, int stage) {
  __shared__ double sh[512];
  extern __shared__ double dyn[];
  const ProState o(M, Din, Dout);
  const int tid = threadIdx.x, nt = blockDim.x;
  double *ell = st + o.ell, *Zt = st + o.Zt, *K0 = st + o.K0, *P = st + o.P, *L = st + o.W1, *Li = st + o.W2;
  // The factorisation is a chain of M column steps with three barriers each: with the factor in global memory
  // every step pays L2 latency (0.40 ms at M = 100); staged in shared memory (stage = 1: L, 2: L and L^-1) the
  // chain runs at shared-memory latency.
  if (stage >= 1) L = dyn;
  if (stage >= 2) Li = dyn + (size_t)M * M;

  for (int j = tid; j < Din; j += nt) {
    const double e = softplus_d(lu[j]);
    ell[j] = e;
    st[o.sgl + j] = sigmoid_d(lu[j]);
    ell32[j] = (float)e;
  }
  if (tid == 0) {
    const double s2 = softplus_d(vu[0]);
    st[o.sig2] = s2;
    st[o.sgv] = sigmoid_d(vu[0]);
    sig232[0] = (float)s2;
  }
  double neg_half_log_s = 0.0;
  for (int i = tid; i < M * Dout; i += nt) {
    const double s = softplus_d(Su[i]);
    st[o.S + i] = s;
    st[o.sgS + i] = sigmoid_d(Su[i]);
    st[o.m + i] = mean[i];
    S32[i] = (float)s;
    neg_half_log_s -= 0.5 * log(s);
  }
  __syncthreads();
  for (int i = tid; i < M * Din; i += nt) {
    Zt[i] = Z[i] / ell[i % Din];
    Z32[i] = (float)Z[i];
  }
  __syncthreads();
  const double sig2 = st[o.sig2];
  for (int i = tid; i < M * M; i += nt) {
    const int r = i / M, c = i % M;
    double d2 = 0.0;
    for (int j = 0; j < Din; ++j) {
      const double e = Zt[r * Din + j] - Zt[c * Din + j];
      d2 += e * e;
    }
    const double k = sig2 * exp(-0.5 * d2);
    K0[i] = k;
    L[i] = k + (r == c ? 1e-8 : 0.0);
  }
  __syncthreads();
  // right-looking Cholesky, lower factor in L
  for (int j = 0; j < M; ++j) {
    if (tid == 0) L[j * M + j] = sqrt(L[j * M + j]);
    __syncthreads();
    const double dj = L[j * M + j];
    for (int i = j + 1 + tid; i < M; i += nt) L[i * M + j] /= dj;
    __syncthreads();
    const int rem = M - j - 1;
    for (int q = tid; q < rem * rem; q += nt) {
      const int i = j + 1 + q / rem, k = j + 1 + q % rem;
      if (k <= i) L[i * M + k] -= L[i * M + j] * L[k * M + j];
    }
    __syncthreads();
  }
  // Li = L^-1 (lower): thread per column
  for (int c = tid; c < M; c += nt) {
    for (int i = 0; i < M; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      if (i < c) { Li[i * M + c] = 0.0; continue; }
      for (int k = c; k < i; ++k) s -= L[i * M + k] * Li[k * M + c];
      Li[i * M + c] = s / L[i * M + i];
    }
  }
  __syncthreads();
  // P = Li^T Li
  for (int i = tid; i < M * M; i += nt) {
    const int r = i / M, c = i % M;
    if (c > r) continue;
    double s = 0.0;
    for (int k = r; k < M; ++k) s += Li[k * M + r] * Li[k * M + c];
    P[r * M + c] = s;
    P[c * M + r] = s;
  }
  __syncthreads();
  for (int i = tid; i < M * M; i += nt) P32[i] = (float)P[i];
  double quad = 0.0;
  for (int i = tid; i < M * Dout; i += nt) {
    const int r = i / Dout, d = i % Dout;
    double s = 0.0;
    for (int k = 0; k < M; ++k) s += P[r * M + k] * st[o.m + k * Dout + d];
    st[o.alpha + i] = s;
    alpha32[i] = (float)s;
    quad += s * st[o.m + i];
  }
  double tr = 0.0, logdet = 0.0;
  for (int r = tid; r < M; r += nt) {
    double ss = 0.0;
    for (int d = 0; d < Dout; ++d) ss += st[o.S + r * Dout + d];
    tr += P[r * M + r] * ss;
    logdet += log(L[r * M + r]);
  }
  const double total = block_sum_d(Dout * logdet + neg_half_log_s + 0.5 * (tr + quad), sh);
  if (tid == 0) kl_out[0] = total - 0.5 * (double)M * Dout;
  // cond_1(K_zz + 1e-8 I) = ||K||_1 ||P||_1 (largest absolute column sums; both matrices are at hand)
  double ck = 0.0, cp = 0.0;
  for (int c = tid; c < M; c += nt) {
    double sk = 1e-8, sp = 0.0;
    for (int r = 0; r < M; ++r) { sk += fabs(K0[r * M + c]); sp += fabs(P[r * M + c]); }
    ck = fmax(ck, sk); cp = fmax(cp, sp);
  }
  __syncthreads();
  sh[tid] = ck;
  __syncthreads();
  for (int q = nt / 2; q > 0; q >>= 1) { if (tid < q) sh[tid] = fmax(sh[tid], sh[tid + q]); __syncthreads(); }
  ck = sh[0];
  __syncthreads();
  sh[tid] = cp;
  __syncthreads();
  for (int q = nt / 2; q > 0; q >>= 1) { if (tid < q) sh[tid] = fmax(sh[tid], sh[tid + q]); __syncthreads(); }
  if (tid == 0) st[o.cond] = ck * sh[0];
}

__global__ void __launch_bounds__(512) gp_prologue_backward_kernel(
    int M, int Din, int Dout, const double *__restrict__ gP, const double *__restrict__ galpha,
    const double *__restrict__ gS, const double *__restrict__ gZ, const double *__restrict__ gell,
    const double *__restrict__ gsig2, double klw, double *__restrict__ st, double *__restrict__ oZ,
    double *__restrict__ om, double *__restrict__ oSu, double *__restrict__ ovu, double *__restrict__ olu) {
  __shared__ double sh[512];
  const ProState o(M, Din, Dout);
  const int tid = threadIdx.x, nt = blockDim.x;
  const double *P = st + o.P, *K0 = st + o.K0, *Zt = st + o.Zt, *mm = st + o.m, *S = st + o.S, *ell = st + o.ell;
  double *A = st + o.W1, *Bm = st + o.W2;
  const double sig2 = st[o.sig2];
  // A = P_bar total
  for (int i = tid; i < M * M; i += nt) {
    const int r = i / M, c = i % M;
    double s = gP[i];
    double mm2 = 0.0;
    for (int d = 0; d < Dout; ++d) {
      s += galpha[r * Dout + d] * mm[c * Dout + d];
      mm2 += mm[r * Dout + d] * mm[c * Dout + d];
    }
    if (r == c) {
      double ss = 0.0;
      for (int d = 0; d < Dout; ++d) ss += S[r * Dout + d];
      mm2 += ss;
    }
    A[i] = s + klw * 0.5 * mm2;
  }
  // zeta_mean and zeta_var_unc adjoints
  for (int i = tid; i < M * Dout; i += nt) {
    const int r = i / Dout, d = i % Dout;
    double s = 0.0;
    for (int k = 0; k < M; ++k) s += P[r * M + k] * galpha[k * Dout + d];
    om[i] = s + klw * st[o.alpha + i];
    oSu[i] = (gS[i] + klw * (0.5 * P[r * M + r] - 0.5 / S[i])) * st[o.sgS + i];
  }
  __syncthreads();
  // Bm = P A
  for (int i = tid; i < M * M; i += nt) {
    const int r = i / M, c = i % M;
    double s = 0.0;
    for (int k = 0; k < M; ++k) s += P[r * M + k] * A[k * M + c];
    Bm[i] = s;
  }
  __syncthreads();
  // A = E = (-(Bm P) + klw*0.5*Dout*P) .* K0
  double esum = 0.0;
  for (int i = tid; i < M * M; i += nt) {
    const int r = i / M, c = i % M;
    double s = 0.0;
    for (int k = 0; k < M; ++k) s += Bm[r * M + k] * P[k * M + c];
    const double e = (-s + klw * 0.5 * Dout * P[i]) * K0[i];
    A[i] = e;
    esum += e;
  }
  esum = block_sum_d(esum, sh);   // also orders the A writes before the reads below
  if (tid == 0) ovu[0] = (gsig2[0] + esum / sig2) * st[o.sgv];
  // Ztbar[i][k] = 2 sum_j W_ij (Zt_ik - Zt_jk), W = -0.5 (E + E^T)
  for (int q = tid; q < M * Din; q += nt) {
    const int i = q / Din, k = q % Din;
    double s = 0.0;
    for (int j = 0; j < M; ++j) s += -0.5 * (A[i * M + j] + A[j * M + i]) * (Zt[q] - Zt[j * Din + k]);
    const double ztb = 2.0 * s;
    oZ[q] = gZ[q] + ztb / ell[k];
    Bm[q] = ztb * Zt[q];           // Bm is free again
  }
  __syncthreads();
  for (int k = tid; k < Din; k += nt) {
    double s = 0.0;
    for (int i = 0; i < M; ++i) s += Bm[i * Din + k];
    olu[k] = (gell[k] - s / ell[k]) * st[o.sgl + k];
  }
}

__global__ void noise_forward_kernel(int dx, const double *__restrict__ vxu, const double *__restrict__ vyu,
                                     float *__restrict__ vx, float *__restrict__ vy) {
  const int j = threadIdx.x;
  if (j < dx) {
    vx[j] = (float)softplus_d(vxu[j]);
    vy[j] = (float)softplus_d(vyu[j]);
  }
}

__global__ void noise_backward_kernel(int dx, const double *__restrict__ vxu, const double *__restrict__ vyu,
                                      const double *__restrict__ gvx, const double *__restrict__ gvy,
                                      double *__restrict__ ovx, double *__restrict__ ovy) {
  const int j = threadIdx.x;
  if (j < dx) {
    ovx[j] = gvx[j] * sigmoid_d(vxu[j]);
    ovy[j] = gvy[j] * sigmoid_d(vyu[j]);
  }
}

}  // namespace cbf

using namespace cbf;

#define CBF_CUDA(expr)                                               \
  do {                                                               \
    cudaError_t _e = (expr);                                         \
    if (_e != cudaSuccess) {                                         \
      set_error("%s failed: %s", #expr, cudaGetErrorString(_e));     \
      return (int)_e;                                                \
    }                                                                \
  } while (0)

extern "C" {

CBF_API int64_t cbf_gp_prologue_state_doubles(int32_t M, int32_t Din, int32_t Dout) {
  if (M < 1 || Din < 1 || Dout < 1) return 0;
  return ProState(M, Din, Dout).total;
}

CBF_API int cbf_gp_prologue(int32_t M, int32_t Din, int32_t Dout, const double *zeta_pos, const double *zeta_mean,
                    const double *zeta_var_unc, const double *variance_unc, const double *lengthscales_unc,
                    float *Z32, float *ell32, float *sig232, float *P32, float *alpha32, float *S32,
                    double *kl_out, double *state, void *stream) {
  if (!zeta_pos || !zeta_mean || !zeta_var_unc || !variance_unc || !lengthscales_unc || !Z32 || !ell32 || !sig232 ||
      !P32 || !alpha32 || !S32 || !kl_out || !state) {
    set_error("cbf_gp_prologue: NULL argument");
    return CBF_ERR_NULL;
  }
  if (M < 1 || Din < 1 || Dout < 1) { set_error("cbf_gp_prologue: invalid shape"); return CBF_ERR_INVALID_SHAPE; }
  const size_t mm = sizeof(double) * (size_t)M * M, cap = 200 * 1024;
  const int stage = 2 * mm <= cap ? 2 : (mm <= cap ? 1 : 0);
  const size_t dyn = stage * mm;
  if (dyn > 0) CBF_CUDA(cudaFuncSetAttribute(gp_prologue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  gp_prologue_kernel<<<1, 512, dyn, static_cast<cudaStream_t>(stream)>>>(M, Din, Dout, zeta_pos, zeta_mean, zeta_var_unc,
                                                                        variance_unc, lengthscales_unc, Z32, ell32,
                                                                        sig232, P32, alpha32, S32, kl_out, state, stage); cbf_note_launch();
  CBF_CUDA(cudaGetLastError());
  return 0;
}

CBF_API int cbf_gp_prologue_backward(int32_t M, int32_t Din, int32_t Dout, const double *gP, const double *galpha,
                             const double *gS, const double *gZ, const double *gell, const double *gsig2,
                             double kl_weight, double *state, double *g_zeta_pos, double *g_zeta_mean,
                             double *g_zeta_var_unc, double *g_variance_unc, double *g_lengthscales_unc,
                             void *stream) {
  if (!gP || !galpha || !gS || !gZ || !gell || !gsig2 || !state || !g_zeta_pos || !g_zeta_mean || !g_zeta_var_unc ||
      !g_variance_unc || !g_lengthscales_unc) {
    set_error("cbf_gp_prologue_backward: NULL argument");
    return CBF_ERR_NULL;
  }
  if (M < 1 || Din < 1 || Dout < 1) { set_error("cbf_gp_prologue_backward: invalid shape"); return CBF_ERR_INVALID_SHAPE; }
  gp_prologue_backward_kernel<<<1, 512, 0, static_cast<cudaStream_t>(stream)>>>(
      M, Din, Dout, gP, galpha, gS, gZ, gell, gsig2, kl_weight, state, g_zeta_pos, g_zeta_mean,
      g_zeta_var_unc, g_variance_unc, g_lengthscales_unc); cbf_note_launch();
  CBF_CUDA(cudaGetLastError());
  return 0;
}

CBF_API int cbf_noise_forward(int32_t dx, const double *var_x_unc, const double *var_y_unc, float *var_x32, float *var_y32,
                      void *stream) {
  if (!var_x_unc || !var_y_unc || !var_x32 || !var_y32) { set_error("cbf_noise_forward: NULL argument"); return CBF_ERR_NULL; }
  if (dx < 1 || dx > 1024) { set_error("cbf_noise_forward: invalid dx"); return CBF_ERR_INVALID_SHAPE; }
  noise_forward_kernel<<<1, round_up(dx, 32), 0, static_cast<cudaStream_t>(stream)>>>(dx, var_x_unc, var_y_unc, var_x32,
                                                                                     var_y32); cbf_note_launch();
  CBF_CUDA(cudaGetLastError());
  return 0;
}

CBF_API int cbf_noise_backward(int32_t dx, const double *var_x_unc, const double *var_y_unc, const double *g_var_x,
                       const double *g_var_y, double *g_var_x_unc, double *g_var_y_unc, void *stream) {
  if (!var_x_unc || !var_y_unc || !g_var_x || !g_var_y || !g_var_x_unc || !g_var_y_unc) {
    set_error("cbf_noise_backward: NULL argument");
    return CBF_ERR_NULL;
  }
  if (dx < 1 || dx > 1024) { set_error("cbf_noise_backward: invalid dx"); return CBF_ERR_INVALID_SHAPE; }
  noise_backward_kernel<<<1, round_up(dx, 32), 0, static_cast<cudaStream_t>(stream)>>>(dx, var_x_unc, var_y_unc, g_var_x,
                                                                                      g_var_y, g_var_x_unc, g_var_y_unc); cbf_note_launch();
  CBF_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
