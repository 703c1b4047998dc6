// Per-particle step arithmetic shared by every kernel path (float32).
// Forward step:  cbfssm/model/cbfssm.py:203-235   (SURVEY 8a note 2)
// Message step:  cbfssm/model/cbfssm.py:143-156   (SURVEY 8a note 1)
#pragma once
#include "common.cuh"

namespace cbf {

// x_{t+1} and the KL_x summand from the GP prediction at x_t.
template <int DX>
__device__ __forceinline__ void fw_step(const float (&x)[DX], const float (&fm0)[DX], const float (&fv0)[DX],
                                        const float (&yt)[DX], float eps, const float *__restrict__ vx,
                                        const float *__restrict__ vyp, float kap, bool do_cond, int ncond,
                                        float (&xn)[DX], float &kl) {
  // ncond: leading state dims that are conditioned on y~ (dx for CBFSSM; dy for CBFSSMHALF,
  // cbfssmhalf.py:144-149, whose remaining dims have k = 0: mu = fmean, sig = fvar, KL summand 0)
#pragma unroll
  for (int j = 0; j < DX; ++j) {
    const float fm = fm0[j] + x[j];          // :205
    const float fv = fv0[j] + vx[j];         // :206
    if (do_cond && j < ncond) {
      const float vy = vyp[j] + (kap - 1.f) * fv;   // :214
      const float s = vy + fv;                      // :216
      const float kg = fv / s;                      // :217
      const float mu = fm + kg * (yt[j] - fm);      // :218
      const float omk = 1.f - kg;
      const float sig = omk * omk * fv + kg * kg * vy;   // :219-220
      xn[j] = mu + eps * sqrtf(sig);                     // :221
      const float dm = mu - fm;
      kl += 0.5f * (logf(fv / sig) + (sig + dm * dm) / fv - 1.f);   // :232-234
    } else {
      xn[j] = fm + eps * sqrtf(fv);                      // :224
    }
  }
}

// Reverse of fw_step. xb = adjoint of x_{t+1}. Outputs the adjoints of the raw GP
// outputs (fmb, fvb) and of y_tilde_{t+1} (ytb); accumulates var_x / var_y adjoints.
template <int DX>
__device__ __forceinline__ void fw_step_adjoint(const float (&x)[DX], const float (&fm0)[DX],
                                                const float (&fv0)[DX], const float (&yt)[DX], float eps,
                                                const float *__restrict__ vx, const float *__restrict__ vyp,
                                                float kap, bool do_cond, int ncond, float w_kl,
                                                const float (&xb)[DX], float (&fmb)[DX], float (&fvb)[DX],
                                                float (&ytb)[DX], float (&vxacc)[DX], float (&vyacc)[DX],
                                                bool accumulate) {
#pragma unroll
  for (int j = 0; j < DX; ++j) {
    const float fm = fm0[j] + x[j];
    const float fv = fv0[j] + vx[j];
    float fmb_j, fvb_j;
    if (do_cond && j < ncond) {
      const float vy = vyp[j] + (kap - 1.f) * fv;
      const float s = vy + fv;
      const float rs = 1.f / s;
      const float kg = fv * rs;
      const float yd = yt[j] - fm;
      const float mu = fm + kg * yd;
      const float omk = 1.f - kg;
      const float sig = omk * omk * fv + kg * kg * vy;
      const float dm = mu - fm;
      const float rfv = 1.f / fv;
      const float mub = xb[j] + w_kl * dm * rfv;
      const float sigb = xb[j] * eps * 0.5f * rsqrtf(sig) + w_kl * 0.5f * (rfv - 1.f / sig);
      fvb_j = w_kl * 0.5f * (rfv - (sig + dm * dm) * rfv * rfv) + sigb * omk * omk;
      fmb_j = -w_kl * dm * rfv + mub * omk;
      const float kgb = sigb * (-2.f * omk * fv + 2.f * kg * vy) + mub * yd;
      float vyb = sigb * kg * kg;
      ytb[j] = mub * kg;
      fvb_j += kgb * rs;
      const float sb = -kgb * fv * rs * rs;
      vyb += sb;
      fvb_j += sb;
      if (accumulate) vyacc[j] += vyb;
      fvb_j += (kap - 1.f) * vyb;
    } else {
      fmb_j = xb[j];
      fvb_j = xb[j] * eps * 0.5f * rsqrtf(fv);
      ytb[j] = 0.f;
    }
    if (accumulate) vxacc[j] += fvb_j;
    fmb[j] = fmb_j;
    fvb[j] = fvb_j;
  }
}

}  // namespace cbf
