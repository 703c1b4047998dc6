// Float64 "batched" path: any M, any (dx, du, dy); the path for M > 128 (P no longer fits an SM: M = 500 is
// 2 MB in float64) and for inducing sets whose cond(K_zz) is beyond float32 (CBF_FLAG_FP64).
//
// Where the persistent kernels keep one particle tile on an SM for all T steps, this path walks time on the
// host and treats ALL particles of a step as one batch, like the reference graph does (cbfssm.py:107-111,
// 176-179: one GPModel.predict on [B*S, Din] per loop iteration):
//
//     x~ = [state, u_t, (y_t)] / ell          assemble kernel      [n, Din+1]   (last column = 1)
//     K  = sigma^2 exp(-|x~ - Z~|^2 / 2)      k kernel             [n, M]
//     A  = K P                                cuBLAS DGEMM         [n, M]       (the M x M contraction)
//     fmean = K alpha, fvar = sigma^2 - rowsum(K .* A) + (A .* A) S            moments kernel (warp / particle)
//     step arithmetic (sampling, KL, entropy), thread per particle
//
// and in reverse, per step: recompute K, A, the moments; step adjoint -> (g_mean, g_var); C = 2 A .* (g_var S^T)
// - 2 G K; PC = C P (DGEMM); k_bar = g_mean alpha^T + PC; w = k_bar .* K; x_bar, ell_bar sums (warp / particle);
// and the parameter adjoints as DGEMMs over the particle dimension accumulating in float64:
// P_bar += A_bar^T K, alpha_bar += K^T g_mean, S_bar += (A .* A)^T g_var, [U | r] += W^T [x~, 1].
// States, messages and message adjoints are kept in float64 arrays of this path's own (the float32 twins in the
// workspace are still written, for cbf_export_states / cbf_state_sums), so nothing between the float32 inputs
// (u, y, draws, var_x, var_y) and the gradient is rounded to float32.  cuBLAS is used for what it
// is: plain DGEMMs ([n x M] x [M x M]); every other kernel is in this file.  Particles are processed in slabs of
// at most kSlab so that the [n, M] matrices stay a few hundred MB.
//
// Mathematics: SURVEY.md 8a notes 1-5, oracle/kernel_math.py (gp_eval / gp_eval_reverse and the two rollouts).
#include <cublas_v2.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "f64_path.h"

namespace cbf {

namespace {

constexpr int kSlab = 1 << 16;          // particles per slab
constexpr int kMaxD = 32;               // upper bound on dx + du (+1) and on dx handled by the per-particle kernels
constexpr double kLog2PiE64 = 2.8378770664093454835606594728112;

#define F64_CUDA(expr)                                               \
  do {                                                               \
    cudaError_t _e = (expr);                                         \
    if (_e != cudaSuccess) {                                         \
      set_error("%s failed: %s", #expr, cudaGetErrorString(_e));     \
      return (int)_e;                                                \
    }                                                                \
  } while (0)
#define F64_BLAS(expr)                                               \
  do {                                                               \
    cublasStatus_t _s = (expr);                                      \
    if (_s != CUBLAS_STATUS_SUCCESS) {                               \
      set_error("%s failed: cuBLAS status %d", #expr, (int)_s);      \
      return 700 + (int)_s;                                          \
    }                                                                \
  } while (0)

struct Gp64 {                // device views of one GP's float64 prologue state
  const double *Zt, *ell, *P, *alpha, *S, *sig2;
  int M, Din, Dout;
};
Gp64 gp64(const double *st, int M, int Din, int Dout) {
  const ProState o(M, Din, Dout);
  return Gp64{st + o.Zt, st + o.ell, st + o.P, st + o.alpha, st + o.S, st + o.sig2, M, Din, Dout};
}

inline unsigned blocks_for(size_t n, int per = 256) {
  size_t b = (n + per - 1) / per;
  return (unsigned)(b > 148 * 32 ? 148 * 32 : (b < 1 ? 1 : b));
}

// ------------------------------------------------------------------------------------------------------
// GP evaluation pieces
// ------------------------------------------------------------------------------------------------------
// K[p][m] = sigma^2 exp(-0.5 |x~_p - Z~_m|^2)   (gp_tf.py:33-49 in difference form)
__global__ void k_kernel(int n, Gp64 g, const double *__restrict__ X1, double *__restrict__ K) {
  const int ld = g.Din + 1;
  const double sig2 = g.sig2[0];
  const size_t total = (size_t)n * g.M;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(i / g.M), m = (int)(i - (size_t)p * g.M);
    const double *x = X1 + (size_t)p * ld, *z = g.Zt + (size_t)m * g.Din;
    double d2 = 0.0;
    for (int j = 0; j < g.Din; ++j) {
      const double e = x[j] - z[j];
      d2 = fma(e, e, d2);
    }
    K[i] = sig2 * exp(-0.5 * d2);
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// fmean = K alpha ; fvar = max(sigma^2 - k.a + sum_m a_m^2 S_md, 0)   (gp_tf.py:140-159); one warp per particle
__global__ void moments_f64_kernel(int n, Gp64 g, const double *__restrict__ K, const double *__restrict__ A,
                                   double *__restrict__ FM, double *__restrict__ FV) {
  const int p = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (p >= n) return;
  const int M = g.M, Dout = g.Dout;
  double q = 0.0, fm[16], fv[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) { fm[d] = 0.0; fv[d] = 0.0; }
  const double *k = K + (size_t)p * M, *a = A + (size_t)p * M;
  for (int m = lane; m < M; m += 32) {
    const double km = k[m], am = a[m], a2 = am * am;
    q = fma(km, am, q);
#pragma unroll
    for (int d = 0; d < 16; ++d)
      if (d < Dout) {
        fm[d] = fma(km, g.alpha[(size_t)m * Dout + d], fm[d]);
        fv[d] = fma(a2, g.S[(size_t)m * Dout + d], fv[d]);
      }
  }
  q = warp_sum(q);
  const double sig2 = g.sig2[0];
#pragma unroll
  for (int d = 0; d < 16; ++d)
    if (d < Dout) {
      const double m1 = warp_sum(fm[d]), v1 = warp_sum(fv[d]);
      if (lane == 0) {
        FM[(size_t)p * Dout + d] = m1;
        FV[(size_t)p * Dout + d] = fmax(sig2 - q + v1, 0.0);     // exact lower bound, cf. gp_var_clamp
      }
    }
}

// C = 2 A .* (g_var S^T) - 2 G K ;  A_bar = C + G K ;  A2 = A .* A      (SURVEY 8a note 4)
__global__ void c_kernel(int n, Gp64 g, const double *__restrict__ K, const double *__restrict__ A,
                         const double *__restrict__ GV, double *__restrict__ Cm, double *__restrict__ AB,
                         double *__restrict__ A2) {
  const size_t total = (size_t)n * g.M;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(i / g.M), m = (int)(i - (size_t)p * g.M);
    const double *gv = GV + (size_t)p * g.Dout, *S = g.S + (size_t)m * g.Dout;
    double G = 0.0, cm = 0.0;
    for (int d = 0; d < g.Dout; ++d) { G += gv[d]; cm = fma(S[d], gv[d], cm); }
    const double a = A[i], k = K[i], b = a * cm;
    Cm[i] = 2.0 * b - 2.0 * G * k;
    AB[i] = 2.0 * b - G * k;
    A2[i] = a * a;
  }
}

// k_bar = g_mean alpha^T + PC ; w = k_bar .* K -> W ; per particle: x~_bar_j = -sum_m w delta_mj,
// L_j += sum_m w delta_mj^2, sw += sum_m w, sG += G.   One warp per particle.
__global__ void kbar_kernel(int n, Gp64 g, const double *__restrict__ K, const double *__restrict__ PC,
                            const double *__restrict__ GM, const double *__restrict__ GV,
                            const double *__restrict__ X1, double *__restrict__ W, double *__restrict__ XB,
                            double *__restrict__ Lacc, double *__restrict__ sacc) {
  const int p = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (p >= n) return;
  const int M = g.M, Din = g.Din, Dout = g.Dout, ld = Din + 1;
  const double *x = X1 + (size_t)p * ld, *gm = GM + (size_t)p * Dout;
  double px[kMaxD], L[kMaxD], sw = 0.0;
#pragma unroll
  for (int j = 0; j < kMaxD; ++j) { px[j] = 0.0; L[j] = 0.0; }
  for (int m = lane; m < M; m += 32) {
    double kb = PC[(size_t)p * M + m];
    for (int d = 0; d < Dout; ++d) kb = fma(g.alpha[(size_t)m * Dout + d], gm[d], kb);
    const double w = kb * K[(size_t)p * M + m];
    W[(size_t)p * M + m] = w;
    sw += w;
    const double *z = g.Zt + (size_t)m * Din;
#pragma unroll
    for (int j = 0; j < kMaxD; ++j)
      if (j < Din) {
        const double dl = x[j] - z[j], wd = w * dl;
        px[j] -= wd;
        L[j] = fma(wd, dl, L[j]);
      }
  }
  sw = warp_sum(sw);
#pragma unroll
  for (int j = 0; j < kMaxD; ++j)
    if (j < Din) {
      const double a = warp_sum(px[j]), b = warp_sum(L[j]);
      if (lane == 0) {
        XB[(size_t)p * Din + j] = a / g.ell[j];       // d/d(input_j) = d/d(x~_j) / ell_j
        Lacc[(size_t)p * Din + j] += b;
      }
    }
  if (lane == 0) {
    double G = 0.0;
    for (int d = 0; d < Dout; ++d) G += GV[(size_t)p * Dout + d];
    sacc[(size_t)p * 2] += sw;
    sacc[(size_t)p * 2 + 1] += G;
  }
}

// ------------------------------------------------------------------------------------------------------
// rollouts: thread per particle.  q = particle index within the slab, nl = q + p0 = local particle.
// ------------------------------------------------------------------------------------------------------
struct Roll {
  Dims D;
  int dx, du, dy, dh, din;
  int p0, ns;                        // slab
  const float *u, *y, *vx, *vy;      // u [B,T,du], y [B,T,dy]; var_x, var_y (float32, constrained)
  Workspace ws;
  // float64 copies of the stored states / messages / message adjoints ([t][j][npad], like their float32 twins in
  // the workspace, which are still written for cbf_export_states / cbf_state_sums): this path's own chain and
  // its reverse pass read these, so nothing between the inputs and the gradient is rounded to float32
  double *X64, *H64, *Yb64;
};

__device__ __forceinline__ int seq_of(const Roll &r, int nl) { return (r.D.n_offset + nl) / r.D.S; }

// message chain: hidden state entering step t_first (cbfssm.py:106,133-135)
__global__ void bm_init_kernel(Roll r, int run, int t, int init, const float *__restrict__ z_b, double *__restrict__ hid) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= r.ns) return;
  const int nl = r.p0 + q;
  const double z = init ? (double)z_b[((size_t)run * r.D.T + t) * r.D.n_local + nl] : 0.0;
  for (int j = 0; j < r.dh; ++j) hid[(size_t)q * r.dh + j] = z;
}
// hidden state entering step t of a chain in the reverse pass: stored output of step t+1 (or the chain's start)
__global__ void bm_load_hidden_kernel(Roll r, int run, int t, int t_top, int init, const float *__restrict__ z_b,
                                      double *__restrict__ hid) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= r.ns) return;
  const int nl = r.p0 + q;
  const size_t np = r.ws.npad;
  for (int j = 0; j < r.dh; ++j) {
    double v;
    if (t == t_top) v = init ? (double)z_b[((size_t)run * r.D.T + t) * r.D.n_local + nl] : 0.0;
    else v = r.H64[(((size_t)run * r.D.T + (t + 1)) * r.dh + j) * np + nl];
    hid[(size_t)q * r.dh + j] = v;
  }
}
// x~ = [hidden, u_t, y_t] / ell , 1     (cbfssm.py:137: hidden first, then u, then y)
__global__ void bm_assemble_kernel(Roll r, int t, const double *__restrict__ ell, const double *__restrict__ hid,
                                   double *__restrict__ X1) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= r.ns) return;
  const int b = seq_of(r, r.p0 + q);
  double *x = X1 + (size_t)q * (r.din + 1);
  for (int j = 0; j < r.dh; ++j) x[j] = hid[(size_t)q * r.dh + j] / ell[j];
  for (int j = 0; j < r.du; ++j) x[r.dh + j] = (double)r.u[((size_t)b * r.D.T + t) * r.du + j] / ell[r.dh + j];
  for (int j = 0; j < r.dy; ++j) x[r.dh + r.du + j] = (double)r.y[((size_t)b * r.D.T + t) * r.dy + j] / ell[r.dh + r.du + j];
  x[r.din] = 1.0;
}
// cbfssm.py:143-158
__global__ void bm_step_kernel(Roll r, int run, int t, const float *__restrict__ eps_b, const double *__restrict__ FM,
                               const double *__restrict__ FV, double *__restrict__ hid, double *__restrict__ ent) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= r.ns) return;
  const int nl = r.p0 + q;
  const size_t np = r.ws.npad;
  const double e = (double)eps_b[((size_t)run * r.D.T + t) * r.D.n_local + nl];
  const bool write = writer_run(t, r.D.R) == run;
  double en = 0.0;
  for (int j = 0; j < r.dh; ++j) {
    const double f = FV[(size_t)q * r.dh + j] + (double)r.vx[j];
    const double out = FM[(size_t)q * r.dh + j] + hid[(size_t)q * r.dh + j] + e * sqrt(f);
    r.ws.H[(((size_t)run * r.D.T + t) * r.dh + j) * np + nl] = (float)out;
    r.H64[(((size_t)run * r.D.T + t) * r.dh + j) * np + nl] = out;
    hid[(size_t)q * r.dh + j] = out;
    if (write) en += 0.5 * (kLog2PiE64 + log(f));
  }
  ent[q] += en;
}
// reverse of one message step: (g_mean, g_var) of the GP outputs from the incoming adjoints
__global__ void bm_adj_pre_kernel(Roll r, int run, int t, const float *__restrict__ eps_b, double w_en,
                                  const double *__restrict__ FV, const double *__restrict__ hb,
                                  double *__restrict__ GM, double *__restrict__ GV, double *__restrict__ vxacc) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= r.ns) return;
  const int nl = r.p0 + q;
  const size_t np = r.ws.npad;
  const double e = (double)eps_b[((size_t)run * r.D.T + t) * r.D.n_local + nl];
  const bool write = writer_run(t, r.D.R) == run;
  for (int j = 0; j < r.dh; ++j) {
    const double f = FV[(size_t)q * r.dh + j] + (double)r.vx[j];
    double ov = hb[(size_t)q * r.dh + j], fb = 0.0;
    if (write) {
      ov += r.Yb64[((size_t)t * r.dh + j) * np + nl];
      fb = w_en * 0.5 / f;
    }
    fb += ov * e * 0.5 / sqrt(f);
    GM[(size_t)q * r.dh + j] = ov;
    GV[(size_t)q * r.dh + j] = fb;
    vxacc[(size_t)q * r.dx + j] += fb;
  }
}
__global__ void bm_adj_post_kernel(Roll r, const double *__restrict__ XB, const double *__restrict__ GM,
                                   double *__restrict__ hb) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= r.ns) return;
  for (int j = 0; j < r.dh; ++j) hb[(size_t)q * r.dh + j] = XB[(size_t)q * r.din + j] + GM[(size_t)q * r.dh + j];
}

// y~_t = [y_t, y2_t]   (cbfssm.py:95-97); CBFSSMHALF has no y2
__device__ __forceinline__ double ytil_at(const Roll &r, int b, int nl, int t, int j) {
  if (j < r.dy) return (double)r.y[((size_t)b * r.D.T + t) * r.dy + j];
  if (r.D.half) return 0.0;
  return r.H64[(((size_t)writer_run(t, r.D.R) * r.D.T + t) * r.dh + (j - r.dy)) * r.ws.npad + nl];
}
// x_0 (cbfssm.py:168; cbfssmhalf.py:103) stored, its likelihood term
__global__ void fw_init_kernel(Roll r, double *__restrict__ xcur, double *__restrict__ sse) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= r.ns) return;
  const int nl = r.p0 + q, b = seq_of(r, nl);
  const size_t np = r.ws.npad;
  for (int j = 0; j < r.dx; ++j) {
    const double v = r.D.half ? (double)r.ws.x0[(size_t)b * r.dx + j] : ytil_at(r, b, nl, 0, j);
    r.ws.X[((size_t)0 * r.dx + j) * np + nl] = (float)v;
    r.X64[((size_t)0 * r.dx + j) * np + nl] = v;
    xcur[(size_t)q * r.dx + j] = v;
    if (j < r.dy) {
      const double d = (double)r.y[((size_t)b * r.D.T + 0) * r.dy + j] - v;
      sse[(size_t)q * r.dy + j] += d * d;
    }
  }
}
__global__ void fw_load_state_kernel(Roll r, int t, double *__restrict__ xcur) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= r.ns) return;
  const int nl = r.p0 + q;
  for (int j = 0; j < r.dx; ++j) xcur[(size_t)q * r.dx + j] = r.X64[((size_t)t * r.dx + j) * r.ws.npad + nl];
}
// x~ = [x_t, u_t] / ell , 1   (cbfssm.py:197)
__global__ void fw_assemble_kernel(Roll r, int t, const double *__restrict__ ell, const double *__restrict__ xcur,
                                   double *__restrict__ X1) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= r.ns) return;
  const int b = seq_of(r, r.p0 + q);
  double *x = X1 + (size_t)q * (r.din + 1);
  for (int j = 0; j < r.dx; ++j) x[j] = xcur[(size_t)q * r.dx + j] / ell[j];
  for (int j = 0; j < r.du; ++j) x[r.dx + j] = (double)r.u[((size_t)b * r.D.T + t) * r.du + j] / ell[r.dx + j];
  x[r.din] = 1.0;
}
// cbfssm.py:203-235 (cbfssmhalf.py:130-166 with ncond = dy)
__global__ void fw_step_kernel(Roll r, int t, const float *__restrict__ eps_f, const double *__restrict__ FM,
                               const double *__restrict__ FV, double *__restrict__ xcur, double *__restrict__ kl,
                               double *__restrict__ sse) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= r.ns) return;
  const int nl = r.p0 + q, b = seq_of(r, nl);
  const size_t np = r.ws.npad;
  const double e = (double)eps_f[(size_t)t * r.D.n_local + nl], kap = (double)r.D.kap;
  const bool do_cond = r.D.condition || (t < r.D.R - 1);
  double klp = 0.0;
  for (int j = 0; j < r.dx; ++j) {
    const double fm = FM[(size_t)q * r.dx + j] + xcur[(size_t)q * r.dx + j];
    const double fv = FV[(size_t)q * r.dx + j] + (double)r.vx[j];
    double xn;
    if (do_cond && j < r.D.ncond) {
      const double vy = (double)r.vy[j] + (kap - 1.0) * fv;
      const double s = vy + fv, kg = fv / s;
      const double mu = fm + kg * (ytil_at(r, b, nl, t + 1, j) - fm);
      const double omk = 1.0 - kg, sig = omk * omk * fv + kg * kg * vy;
      xn = mu + e * sqrt(sig);
      const double dm = mu - fm;
      klp += 0.5 * (log(fv) - log(sig) + (sig + dm * dm) / fv - 1.0);
    } else {
      xn = fm + e * sqrt(fv);
    }
    r.ws.X[((size_t)(t + 1) * r.dx + j) * np + nl] = (float)xn;
    r.X64[((size_t)(t + 1) * r.dx + j) * np + nl] = xn;
    xcur[(size_t)q * r.dx + j] = xn;
    if (j < r.dy) {
      const double d = (double)r.y[((size_t)b * r.D.T + (t + 1)) * r.dy + j] - xn;
      sse[(size_t)q * r.dy + j] += d * d;
    }
  }
  kl[q] += klp;
}
// adjoint of x_{T-1}: likelihood only
__global__ void fw_adj_init_kernel(Roll r, double w_ll, double *__restrict__ xb) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= r.ns) return;
  const int nl = r.p0 + q, b = seq_of(r, nl), t = r.D.T - 1;
  for (int j = 0; j < r.dx; ++j) {
    double v = 0.0;
    if (j < r.dy)
      v = w_ll * ((double)r.y[((size_t)b * r.D.T + t) * r.dy + j] - r.X64[((size_t)t * r.dx + j) * r.ws.npad + nl]) /
          (double)r.vy[j];
    xb[(size_t)q * r.dx + j] = v;
  }
}
// reverse of fw_step: xb = adjoint of x_{t+1} -> (g_mean, g_var), adjoint of y2[t+1], var_x / var_y sums
__global__ void fw_adj_pre_kernel(Roll r, int t, const float *__restrict__ eps_f, double w_kl,
                                  const double *__restrict__ xcur, const double *__restrict__ FM,
                                  const double *__restrict__ FV, const double *__restrict__ xb,
                                  double *__restrict__ GM, double *__restrict__ GV, double *__restrict__ vxacc,
                                  double *__restrict__ vyacc) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= r.ns) return;
  const int nl = r.p0 + q, b = seq_of(r, nl);
  const size_t np = r.ws.npad;
  const double e = (double)eps_f[(size_t)t * r.D.n_local + nl], kap = (double)r.D.kap;
  const bool do_cond = r.D.condition || (t < r.D.R - 1);
  for (int j = 0; j < r.dx; ++j) {
    const double fm = FM[(size_t)q * r.dx + j] + xcur[(size_t)q * r.dx + j];
    const double fv = FV[(size_t)q * r.dx + j] + (double)r.vx[j];
    const double xbj = xb[(size_t)q * r.dx + j];
    double fmb, fvb, ytb = 0.0;
    if (do_cond && j < r.D.ncond) {
      const double vy = (double)r.vy[j] + (kap - 1.0) * fv;
      const double s = vy + fv, rs = 1.0 / s, kg = fv * rs;
      const double yd = ytil_at(r, b, nl, t + 1, j) - fm;
      const double mu = fm + kg * yd, omk = 1.0 - kg;
      const double sig = omk * omk * fv + kg * kg * vy, dm = mu - fm, rfv = 1.0 / fv;
      const double mub = xbj + w_kl * dm * rfv;
      const double sigb = xbj * e * 0.5 / sqrt(sig) + w_kl * 0.5 * (rfv - 1.0 / sig);
      fvb = w_kl * 0.5 * (rfv - (sig + dm * dm) * rfv * rfv) + sigb * omk * omk;
      fmb = -w_kl * dm * rfv + mub * omk;
      const double kgb = sigb * (-2.0 * omk * fv + 2.0 * kg * vy) + mub * yd;
      double vyb = sigb * kg * kg;
      ytb = mub * kg;
      fvb += kgb * rs;
      const double sb = -kgb * fv * rs * rs;
      vyb += sb;
      fvb += sb;
      vyacc[(size_t)q * r.dx + j] += vyb;
      fvb += (kap - 1.0) * vyb;
    } else {
      fmb = xbj;
      fvb = xbj * e * 0.5 / sqrt(fv);
    }
    vxacc[(size_t)q * r.dx + j] += fvb;
    GM[(size_t)q * r.dx + j] = fmb;
    GV[(size_t)q * r.dx + j] = fvb;
    if (j >= r.dy && !r.D.half) r.Yb64[((size_t)(t + 1) * r.dh + (j - r.dy)) * np + nl] = ytb;
  }
}
// xb(x_t) = x_in_bar + g_mean (fm = fmean + x_t) + likelihood term of step t
__global__ void fw_adj_post_kernel(Roll r, int t, double w_ll, const double *__restrict__ xcur,
                                   const double *__restrict__ XB, const double *__restrict__ GM,
                                   double *__restrict__ xb) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= r.ns) return;
  const int nl = r.p0 + q, b = seq_of(r, nl);
  for (int j = 0; j < r.dx; ++j) {
    double lg = 0.0;
    if (j < r.dy)
      lg = w_ll * ((double)r.y[((size_t)b * r.D.T + t) * r.dy + j] - xcur[(size_t)q * r.dx + j]) / (double)r.vy[j];
    xb[(size_t)q * r.dx + j] = XB[(size_t)q * r.din + j] + GM[(size_t)q * r.dx + j] + lg;
  }
}
// adjoint of x_0: y2[0] part -> Yb[0] (cbfssm.py:168), or all of it -> x0b (CBFSSMHALF)
__global__ void fw_adj_final_kernel(Roll r, const double *__restrict__ xb) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= r.ns) return;
  const int nl = r.p0 + q;
  const size_t np = r.ws.npad;
  if (r.D.half) {
    for (int j = 0; j < r.dx; ++j) r.ws.x0b[(size_t)j * np + nl] = (float)xb[(size_t)q * r.dx + j];
  } else {
    for (int j = 0; j < r.dh; ++j) r.Yb64[((size_t)0 * r.dh + j) * np + nl] = xb[(size_t)q * r.dx + r.dy + j];
  }
}

// out[c] += sum_p in[p][c]
__global__ void colsum_kernel(int n, int cols, const double *__restrict__ in, double *__restrict__ out) {
  __shared__ double sh[256];
  const int c = blockIdx.y;
  double s = 0.0;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) s += in[(size_t)p * cols + c];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out + c, sh[0]);
}

// terms from the reduced statistics (cbfssm.py:245-251,183,99); stats = [sse_j (dy) | kl_x | entropy]
__global__ void terms_f64_kernel(int dy, const float *__restrict__ vy, double n_times_t, const double *__restrict__ stats,
                                 double *__restrict__ terms) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double ll = 0.0;
  for (int j = 0; j < dy; ++j) {
    const double v = (double)vy[j];
    ll += -0.5 * stats[j] / v - 0.5 * n_times_t * (log(v) + 1.8378770664093454836);
  }
  terms[0] = ll;
  terms[1] = stats[dy];
  terms[2] = stats[dy + 1];
}

// kernel-level gradient of one GP from the float64 accumulators (same meaning as finalize_gp_grad_kernel)
__global__ void finalize_f64_kernel(Gp64 g, const double *__restrict__ UR, const double *__restrict__ Lsum,
                                    const double *__restrict__ ssum, double *__restrict__ gZ,
                                    double *__restrict__ gell, double *__restrict__ gsig2) {
  const int M = g.M, Din = g.Din, tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  for (int i = tid; i < M * Din; i += nt) {
    const int m = i / Din, j = i % Din;
    gZ[i] = (UR[(size_t)m * (Din + 1) + j] - g.Zt[i] * UR[(size_t)m * (Din + 1) + Din]) / g.ell[j];
  }
  for (int j = tid; j < Din; j += nt) gell[j] = Lsum[j] / g.ell[j];
  if (tid == 0) gsig2[0] = ssum[0] / g.sig2[0] + ssum[1];
}
__global__ void finalize_noise_f64_kernel(int dx, int dy, const double *__restrict__ vxsum, const double *__restrict__ vysum,
                                          const double *__restrict__ stats, const float *__restrict__ vy, double w_ll,
                                          double n_times_t, double *__restrict__ gvx, double *__restrict__ gvy) {
  const int j = threadIdx.x;
  if (j < dx) gvx[j] = vxsum[j];
  if (j < dx) {
    double gg = vysum[j];
    if (j < dy) {
      const double v = (double)vy[j];
      gg += w_ll * (0.5 * stats[j] / (v * v) - 0.5 * n_times_t / v);
    }
    gvy[j] = gg;
  }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
struct Scratch {            // float64 buffers carved from the caller's workspace (per slab unless noted)
  double *X64, *H64, *Yb64; // all particles of the call: [T][dx][npad], [2][T][dh][npad], [T][dh][npad]
  double *X1, *K, *A, *C, *AB, *A2, *FM, *FV, *GM, *GV, *XB;
  double *state, *adj;      // xcur / hid  and  xb / hb
  double *ent, *kl, *sse;   // forward per-particle sums
  double *Lf, *sf, *Lb, *sb, *vxacc, *vyacc;   // reverse per-particle sums
  double *UR;               // 2 x [M, Din+1]: [U | r] of the forward GP, then of the message GP
  double *red;              // small reduction outputs
  void *blas_ws;
  size_t blas_ws_bytes;
};
constexpr size_t kBlasWs = (size_t)32 << 20;

size_t carve(Scratch *s, char *base, int ns, int M, int dx, int dy, int din, size_t npad, int T) {
  size_t o = 0;
  auto take = [&](double **p, size_t count) {
    if (s) *p = reinterpret_cast<double *>(base + o);
    o += (count * sizeof(double) + 255) / 256 * 256;
  };
  double *dummy;
  const size_t nm = (size_t)ns * M;
  take(s ? &s->X64 : &dummy, (size_t)T * dx * npad);
  take(s ? &s->H64 : &dummy, (size_t)2 * T * (dx - dy) * npad);
  take(s ? &s->Yb64 : &dummy, (size_t)T * (dx - dy) * npad);
  take(s ? &s->X1 : &dummy, (size_t)ns * (din + 1));
  take(s ? &s->K : &dummy, nm); take(s ? &s->A : &dummy, nm); take(s ? &s->C : &dummy, nm);
  take(s ? &s->AB : &dummy, nm); take(s ? &s->A2 : &dummy, nm);
  take(s ? &s->FM : &dummy, (size_t)ns * dx); take(s ? &s->FV : &dummy, (size_t)ns * dx);
  take(s ? &s->GM : &dummy, (size_t)ns * dx); take(s ? &s->GV : &dummy, (size_t)ns * dx);
  take(s ? &s->XB : &dummy, (size_t)ns * din);
  take(s ? &s->state : &dummy, (size_t)ns * dx); take(s ? &s->adj : &dummy, (size_t)ns * dx);
  take(s ? &s->ent : &dummy, ns); take(s ? &s->kl : &dummy, ns); take(s ? &s->sse : &dummy, (size_t)ns * dy);
  take(s ? &s->Lf : &dummy, (size_t)ns * din); take(s ? &s->sf : &dummy, (size_t)ns * 2);
  take(s ? &s->Lb : &dummy, (size_t)ns * din); take(s ? &s->sb : &dummy, (size_t)ns * 2);
  take(s ? &s->vxacc : &dummy, (size_t)ns * dx); take(s ? &s->vyacc : &dummy, (size_t)ns * dx);
  take(s ? &s->UR : &dummy, (size_t)2 * M * (din + 1));
  take(s ? &s->red : &dummy, 4 * (size_t)(din + dx + 8));
  if (s) { s->blas_ws = base + o; s->blas_ws_bytes = kBlasWs; }
  o += kBlasWs;
  return o;
}

thread_local cublasHandle_t g_blas = nullptr;
int blas_for(cudaStream_t st, const Scratch &s, cublasHandle_t *out) {
  if (!g_blas) F64_BLAS(cublasCreate(&g_blas));
  F64_BLAS(cublasSetStream(g_blas, st));
  F64_BLAS(cublasSetWorkspace(g_blas, s.blas_ws, s.blas_ws_bytes));
  F64_BLAS(cublasSetPointerMode(g_blas, CUBLAS_POINTER_MODE_HOST));
  *out = g_blas;
  return 0;
}

// A[n, M] (row-major) = K[n, M] P[M, M]   (P symmetric)
int gemm_kp(cublasHandle_t h, int n, int M, const double *K, const double *P, double *A) {
  const double one = 1.0, zero = 0.0;
  F64_BLAS(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, M, n, M, &one, P, M, K, M, &zero, A, M));
  cbf_note_launch();
  return 0;
}
// R[rows_r, cols_l] (row-major [cols_l... see below]) += sum_p L[p][i] R[p][j]:  out row-major [ni, nj], out[i][j] += sum_p Lm[p][i] Rm[p][j]
int gemm_acc(cublasHandle_t h, int n, int ni, int nj, const double *Lm, const double *Rm, double *out) {
  const double one = 1.0;
  // column-major view: out_cm (nj x ni)[j][i] = sum_p Rm_cm(nj x n)[j][p] * Lm_cm(ni x n)[i][p]
  F64_BLAS(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, nj, ni, n, &one, Rm, nj, Lm, ni, &one, out, nj));
  cbf_note_launch();
  return 0;
}

#define LAUNCH1(kernel, n, ...)                                                      \
  do {                                                                               \
    kernel<<<((n) + 255) / 256, 256, 0, st>>>(__VA_ARGS__);                          \
    cbf_note_launch();                                                               \
  } while (0)

// K, A, moments of the slab's current X1
int gp_forward(cublasHandle_t h, cudaStream_t st, const Gp64 &g, int ns, const Scratch &s) {
  k_kernel<<<blocks_for((size_t)ns * g.M), 256, 0, st>>>(ns, g, s.X1, s.K); cbf_note_launch();
  int rc = gemm_kp(h, ns, g.M, s.K, g.P, s.A);
  if (rc) return rc;
  moments_f64_kernel<<<((size_t)ns * 32 + 255) / 256, 256, 0, st>>>(ns, g, s.K, s.A, s.FM, s.FV); cbf_note_launch();
  return 0;
}
// reverse of the evaluation whose K, A are in the scratch; GM, GV given; XB (input adjoint) out; accumulators updated
int gp_reverse(cublasHandle_t h, cudaStream_t st, const Gp64 &g, int ns, const Scratch &s, double *Lacc, double *sacc,
               double *gP, double *galpha, double *gS, double *UR) {
  c_kernel<<<blocks_for((size_t)ns * g.M), 256, 0, st>>>(ns, g, s.K, s.A, s.GV, s.C, s.AB, s.A2); cbf_note_launch();
  int rc = gemm_kp(h, ns, g.M, s.C, g.P, s.A);                     // A <- PC (a is not needed any more)
  if (rc) return rc;
  kbar_kernel<<<((size_t)ns * 32 + 255) / 256, 256, 0, st>>>(ns, g, s.K, s.A, s.GM, s.GV, s.X1, s.C, s.XB, Lacc, sacc);   // C <- W
  cbf_note_launch();
  if ((rc = gemm_acc(h, ns, g.M, g.M, s.AB, s.K, gP))) return rc;              // P_bar[i][j]   += a_bar_i k_j
  if ((rc = gemm_acc(h, ns, g.M, g.Dout, s.K, s.GM, galpha))) return rc;       // alpha_bar[m][d] += k_m g_mean_d
  if ((rc = gemm_acc(h, ns, g.M, g.Dout, s.A2, s.GV, gS))) return rc;          // S_bar[m][d]   += a_m^2 g_var_d
  if ((rc = gemm_acc(h, ns, g.M, g.Din + 1, s.C, s.X1, UR))) return rc;        // [U | r][m][j] += w_m [x~, 1]_j
  return 0;
}

Roll make_roll(const F64Args &a, const Scratch &s, int p0, int ns) {
  Roll r;
  r.X64 = s.X64; r.H64 = s.H64; r.Yb64 = s.Yb64;
  r.D = a.D; r.dx = a.dx; r.du = a.du; r.dy = a.dy; r.dh = a.dx - a.dy; r.din = a.dx + a.du;
  r.p0 = p0; r.ns = ns; r.u = a.u; r.y = a.y; r.vx = a.var_x; r.vy = a.var_y; r.ws = a.ws;
  return r;
}

}  // namespace

size_t f64_scratch_bytes(int n_local, int T, int M, int dx, int dy, int din) {
  const int ns = n_local < kSlab ? n_local : kSlab;
  return carve(nullptr, nullptr, ns, M, dx, dy, din, (size_t)round_up(n_local, 32), T) + 256;
}

int f64_forward(const F64Args &a, double *terms) {
  if (a.dx + a.du + 1 > kMaxD || a.dx > 16) { set_error("float64 path: dx + du must be <= %d and dx <= 16", kMaxD - 1); return CBF_ERR_UNSUPPORTED_DIMS; }
  cudaStream_t st = a.stream;
  const int M = a.D.M, dx = a.dx, dy = a.dy, dh = dx - dy, din = dx + a.du, T = a.D.T;
  const Gp64 gf = gp64(a.state_f, M, din, dx);
  Gp64 gb = gf;
  if (!a.D.half) gb = gp64(a.state_b, M, din, dh);
  F64_CUDA(cudaMemsetAsync(a.ws.stats, 0, sizeof(double) * (dy + 2), st));
  for (int p0 = 0; p0 < a.D.n_local; p0 += kSlab) {
    const int ns = std::min(kSlab, a.D.n_local - p0);
    Scratch s;
    carve(&s, static_cast<char *>(a.scratch), std::min(kSlab, a.D.n_local), M, dx, dy, din, (size_t)a.D.npad, T);
    cublasHandle_t h;
    int rc = blas_for(st, s, &h);
    if (rc) return rc;
    const Roll r = make_roll(a, s, p0, ns);
    F64_CUDA(cudaMemsetAsync(s.ent, 0, sizeof(double) * ns, st));
    F64_CUDA(cudaMemsetAsync(s.kl, 0, sizeof(double) * ns, st));
    F64_CUDA(cudaMemsetAsync(s.sse, 0, sizeof(double) * (size_t)ns * dy, st));
    // ---- backward message: every live chain segment (cbfssm.py:101-158) ----
    for (const Chain &c : *a.chains) {
      LAUNCH1(bm_init_kernel, ns, r, c.run, c.t_hi, c.init, a.z_b, s.state);
      for (int t = c.t_hi; t >= c.t_lo; --t) {
        LAUNCH1(bm_assemble_kernel, ns, r, t, gb.ell, s.state, s.X1);
        if ((rc = gp_forward(h, st, gb, ns, s))) return rc;
        LAUNCH1(bm_step_kernel, ns, r, c.run, t, a.eps_b, s.FM, s.FV, s.state, s.ent);
      }
    }
    // ---- forward conditional rollout (cbfssm.py:160-237) ----
    LAUNCH1(fw_init_kernel, ns, r, s.state, s.sse);
    for (int t = 0; t + 1 < T; ++t) {
      LAUNCH1(fw_assemble_kernel, ns, r, t, gf.ell, s.state, s.X1);
      if ((rc = gp_forward(h, st, gf, ns, s))) return rc;
      LAUNCH1(fw_step_kernel, ns, r, t, a.eps_f, s.FM, s.FV, s.state, s.kl, s.sse);
    }
    colsum_kernel<<<dim3(blocks_for(ns) > 64 ? 64 : blocks_for(ns), dy), 256, 0, st>>>(ns, dy, s.sse, a.ws.stats); cbf_note_launch();
    colsum_kernel<<<dim3(blocks_for(ns) > 64 ? 64 : blocks_for(ns), 1), 256, 0, st>>>(ns, 1, s.kl, a.ws.stats + dy); cbf_note_launch();
    colsum_kernel<<<dim3(blocks_for(ns) > 64 ? 64 : blocks_for(ns), 1), 256, 0, st>>>(ns, 1, s.ent, a.ws.stats + dy + 1); cbf_note_launch();
    F64_CUDA(cudaGetLastError());
  }
  terms_f64_kernel<<<1, 32, 0, st>>>(dy, a.var_y, (double)a.D.n_local * T, a.ws.stats, terms); cbf_note_launch();
  F64_CUDA(cudaGetLastError());
  return 0;
}

int f64_backward(const F64Args &a, double w_ll, double w_kl, double w_en, const F64Grad &g) {
  if (a.dx + a.du + 1 > kMaxD || a.dx > 16) { set_error("float64 path: dx + du must be <= %d and dx <= 16", kMaxD - 1); return CBF_ERR_UNSUPPORTED_DIMS; }
  cudaStream_t st = a.stream;
  const int M = a.D.M, dx = a.dx, dy = a.dy, dh = dx - dy, din = dx + a.du, T = a.D.T;
  const Gp64 gf = gp64(a.state_f, M, din, dx);
  Gp64 gb = gf;
  if (!a.D.half) gb = gp64(a.state_b, M, din, dh);
  F64_CUDA(cudaMemsetAsync(g.base, 0, sizeof(double) * (size_t)g.total, st));
  const int nsmax = std::min(kSlab, a.D.n_local);
  Scratch s;
  carve(&s, static_cast<char *>(a.scratch), nsmax, M, dx, dy, din, (size_t)a.D.npad, T);
  // reduction targets: [Lf (din) | sf (2) | Lb (din) | sb (2) | vx (dx) | vy (dx)]
  double *rLf = s.red, *rsf = rLf + din, *rLb = rsf + 2, *rsb = rLb + din, *rvx = rsb + 2, *rvy = rvx + dx;
  F64_CUDA(cudaMemsetAsync(s.red, 0, sizeof(double) * (size_t)(2 * din + 4 + 2 * dx), st));
  F64_CUDA(cudaMemsetAsync(s.UR, 0, sizeof(double) * (size_t)2 * M * (din + 1), st));
  double *URf = s.UR, *URb = s.UR + (size_t)M * (din + 1);
  for (int p0 = 0; p0 < a.D.n_local; p0 += kSlab) {
    const int ns = std::min(kSlab, a.D.n_local - p0);
    cublasHandle_t h;
    int rc = blas_for(st, s, &h);
    if (rc) return rc;
    const Roll r = make_roll(a, s, p0, ns);
    F64_CUDA(cudaMemsetAsync(s.Lf, 0, sizeof(double) * (size_t)ns * din, st));
    F64_CUDA(cudaMemsetAsync(s.sf, 0, sizeof(double) * (size_t)ns * 2, st));
    F64_CUDA(cudaMemsetAsync(s.Lb, 0, sizeof(double) * (size_t)ns * din, st));
    F64_CUDA(cudaMemsetAsync(s.sb, 0, sizeof(double) * (size_t)ns * 2, st));
    F64_CUDA(cudaMemsetAsync(s.vxacc, 0, sizeof(double) * (size_t)ns * dx, st));
    F64_CUDA(cudaMemsetAsync(s.vyacc, 0, sizeof(double) * (size_t)ns * dx, st));
    // ---- reverse of the forward rollout ----
    LAUNCH1(fw_adj_init_kernel, ns, r, w_ll, s.adj);
    for (int t = T - 2; t >= 0; --t) {
      LAUNCH1(fw_load_state_kernel, ns, r, t, s.state);
      LAUNCH1(fw_assemble_kernel, ns, r, t, gf.ell, s.state, s.X1);
      if ((rc = gp_forward(h, st, gf, ns, s))) return rc;
      LAUNCH1(fw_adj_pre_kernel, ns, r, t, a.eps_f, w_kl, s.state, s.FM, s.FV, s.adj, s.GM, s.GV, s.vxacc, s.vyacc);
      if ((rc = gp_reverse(h, st, gf, ns, s, s.Lf, s.sf, g.f_P, g.f_alpha, g.f_S, URf))) return rc;
      LAUNCH1(fw_adj_post_kernel, ns, r, t, w_ll, s.state, s.XB, s.GM, s.adj);
    }
    LAUNCH1(fw_adj_final_kernel, ns, r, s.adj);
    // ---- reverse of the message chains ----
    for (const Chain &c : *a.chains) {
      F64_CUDA(cudaMemsetAsync(s.adj, 0, sizeof(double) * (size_t)ns * dh, st));
      for (int t = c.t_lo; t <= c.t_hi; ++t) {
        LAUNCH1(bm_load_hidden_kernel, ns, r, c.run, t, c.t_hi, c.init, a.z_b, s.state);
        LAUNCH1(bm_assemble_kernel, ns, r, t, gb.ell, s.state, s.X1);
        if ((rc = gp_forward(h, st, gb, ns, s))) return rc;
        LAUNCH1(bm_adj_pre_kernel, ns, r, c.run, t, a.eps_b, w_en, s.FV, s.adj, s.GM, s.GV, s.vxacc);
        if ((rc = gp_reverse(h, st, gb, ns, s, s.Lb, s.sb, g.b_P, g.b_alpha, g.b_S, URb))) return rc;
        LAUNCH1(bm_adj_post_kernel, ns, r, s.XB, s.GM, s.adj);
      }
    }
    // per-particle sums of this slab
    const unsigned gx = blocks_for(ns) > 64 ? 64 : blocks_for(ns);
    colsum_kernel<<<dim3(gx, din), 256, 0, st>>>(ns, din, s.Lf, rLf); cbf_note_launch();
    colsum_kernel<<<dim3(gx, 2), 256, 0, st>>>(ns, 2, s.sf, rsf); cbf_note_launch();
    colsum_kernel<<<dim3(gx, din), 256, 0, st>>>(ns, din, s.Lb, rLb); cbf_note_launch();
    colsum_kernel<<<dim3(gx, 2), 256, 0, st>>>(ns, 2, s.sb, rsb); cbf_note_launch();
    colsum_kernel<<<dim3(gx, dx), 256, 0, st>>>(ns, dx, s.vxacc, rvx); cbf_note_launch();
    colsum_kernel<<<dim3(gx, dx), 256, 0, st>>>(ns, dx, s.vyacc, rvy); cbf_note_launch();
    F64_CUDA(cudaGetLastError());
  }
  finalize_f64_kernel<<<8, 256, 0, st>>>(gf, URf, rLf, rsf, g.f_Z, g.f_ell, g.f_sig2); cbf_note_launch();
  if (!a.D.half) { finalize_f64_kernel<<<8, 256, 0, st>>>(gb, URb, rLb, rsb, g.b_Z, g.b_ell, g.b_sig2); cbf_note_launch(); }
  finalize_noise_f64_kernel<<<1, 32, 0, st>>>(dx, dy, rvx, rvy, a.ws.stats, a.var_y, w_ll, (double)a.D.n_local * T,
                                              g.var_x, g.var_y); cbf_note_launch();
  F64_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace cbf
