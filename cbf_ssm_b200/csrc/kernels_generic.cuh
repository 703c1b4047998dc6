// Cooperative ("generic") kernel path: any M whose resident parameter set fits one SM.
//
// A CTA owns a tile of kNP = 32 particles and blockDim/32 "parts" (one warp each).
// All per-particle M-vectors (k, a, b, w) live in shared memory as [m][n] arrays
// (n contiguous, row stride kLD), the resident GP operands P = K_zz^-1 (padded),
// Z/ell, alpha, S stay in shared memory for all T steps, and each part handles a
// round-robin subset of 4-row groups of the M x M contraction.  The same [m][n]
// arrays are the K-major operands of the per-step 4x4-tiled outer-product
// accumulation of the parameter adjoints (P_bar += a_bar k^T, ...), whose
// accumulators also stay in shared memory until the CTA has finished all its work.
//
// Mathematics: SURVEY.md 8a notes 1-5; verified in float64 by oracle/kernel_math.py.
#pragma once
#include "common.cuh"
#include "step_math.cuh"

namespace cbf {

template <int DIN, int DOUT>
struct GpS {
  static constexpr int DINP = (DIN + 3) / 4 * 4;
  static constexpr int DOUTP = (DOUT + 3) / 4 * 4;
  float *P, *Zt, *al, *Sm, *il;
  float sig2, lsig;
  int M, MP, MG;

  __device__ static size_t floats(int M) {
    const int MP = round_up(M, 4);
    return (size_t)MP * MP + (size_t)MP * (DINP + 2 * DOUTP) + DINP + 4;
  }
  __host__ static size_t floats_host(int M) {
    const int MP = round_up(M, 4);
    return (size_t)MP * MP + (size_t)MP * (DINP + 2 * DOUTP) + DINP + 4;
  }
  // Carve from `base` and fill from global memory (all threads; caller syncs).
  __device__ float *init(float *base, const GpDev &g, int M_) {
    M = M_; MP = round_up(M, 4); MG = MP / 4;
    P = base; base += MP * MP;
    Zt = base; base += MP * DINP;
    al = base; base += MP * DOUTP;
    Sm = base; base += MP * DOUTP;
    il = base; base += DINP + 4;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < MP * MP; i += nt) {
      const int r = i / MP, c = i % MP;
      P[i] = (r < M && c < M) ? g.P[r * M + c] : 0.f;
    }
    for (int i = tid; i < MP * DINP; i += nt) {
      const int r = i / DINP, c = i % DINP;
      Zt[i] = (r < M && c < DIN) ? g.Z[r * DIN + c] / g.ell[c] : 0.f;
    }
    for (int i = tid; i < MP * DOUTP; i += nt) {
      const int r = i / DOUTP, c = i % DOUTP;
      const bool ok = (r < M && c < DOUT);
      al[i] = ok ? g.alpha[r * DOUT + c] : 0.f;
      Sm[i] = ok ? g.S[r * DOUT + c] : 0.f;
    }
    for (int i = tid; i < DINP; i += nt) il[i] = (i < DIN) ? 1.f / g.ell[i] : 0.f;
    sig2 = g.sig2[0];
    lsig = log2f(sig2);
    return base;
  }
};

// ---- one sparse-GP evaluation for the CTA's 32 particles (gp_tf.py:132-161) ----
// Every thread of particle n passes the same xin and receives the same (fm, fv).
// kk[m][n] = k_m; if SAVE_A also aa[m][n] = (P k)_m.  Contains two __syncthreads().
template <int DIN, int DOUT, bool SAVE_A>
__device__ __forceinline__ void gp_forward_coop(const GpS<DIN, DOUT> &g, float *kk, float *aa, float *red,
                                                int part, int parts, int n, const float (&xin)[DIN],
                                                float (&xt)[GpS<DIN, DOUT>::DINP], float (&fm)[DOUT],
                                                float (&fv)[DOUT]) {
  constexpr int DINP = GpS<DIN, DOUT>::DINP, DOUTP = GpS<DIN, DOUT>::DOUTP;
  constexpr int NRED = 1 + 2 * DOUT;
#pragma unroll
  for (int j = 0; j < DINP; ++j) xt[j] = (j < DIN) ? xin[j < DIN ? j : 0] * g.il[j] : 0.f;

  float pm[DOUT];
#pragma unroll
  for (int d = 0; d < DOUT; ++d) pm[d] = 0.f;
  for (int grp = part; grp < g.MG; grp += parts) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = 4 * grp + r;
      float d2 = 0.f;
#pragma unroll
      for (int j = 0; j < DINP; j += 4) {
        const float4 z = *reinterpret_cast<const float4 *>(g.Zt + m * DINP + j);
        float e;
        e = xt[j] - z.x; d2 = fmaf(e, e, d2);
        e = xt[j + 1] - z.y; d2 = fmaf(e, e, d2);
        e = xt[j + 2] - z.z; d2 = fmaf(e, e, d2);
        e = xt[j + 3] - z.w; d2 = fmaf(e, e, d2);
      }
      const float k = (m < g.M) ? exp2f(fmaf(kNegHalfLog2e, d2, g.lsig)) : 0.f;
      kk[m * kLD + n] = k;
#pragma unroll
      for (int d = 0; d < DOUT; ++d) pm[d] = fmaf(k, g.al[m * DOUTP + d], pm[d]);
    }
  }
  __syncthreads();

  float q = 0.f, pv[DOUT];
#pragma unroll
  for (int d = 0; d < DOUT; ++d) pv[d] = 0.f;
  for (int grp = part; grp < g.MG; grp += parts) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float *P0 = g.P + (4 * grp) * g.MP;
    for (int mp = 0; mp < g.MP; mp += 4) {
      const float4 p0 = *reinterpret_cast<const float4 *>(P0 + mp);
      const float4 p1 = *reinterpret_cast<const float4 *>(P0 + g.MP + mp);
      const float4 p2 = *reinterpret_cast<const float4 *>(P0 + 2 * g.MP + mp);
      const float4 p3 = *reinterpret_cast<const float4 *>(P0 + 3 * g.MP + mp);
      const float k0 = kk[(mp + 0) * kLD + n], k1 = kk[(mp + 1) * kLD + n];
      const float k2 = kk[(mp + 2) * kLD + n], k3 = kk[(mp + 3) * kLD + n];
      a0 = fmaf(p0.x, k0, a0); a0 = fmaf(p0.y, k1, a0); a0 = fmaf(p0.z, k2, a0); a0 = fmaf(p0.w, k3, a0);
      a1 = fmaf(p1.x, k0, a1); a1 = fmaf(p1.y, k1, a1); a1 = fmaf(p1.z, k2, a1); a1 = fmaf(p1.w, k3, a1);
      a2 = fmaf(p2.x, k0, a2); a2 = fmaf(p2.y, k1, a2); a2 = fmaf(p2.z, k2, a2); a2 = fmaf(p2.w, k3, a2);
      a3 = fmaf(p3.x, k0, a3); a3 = fmaf(p3.y, k1, a3); a3 = fmaf(p3.z, k2, a3); a3 = fmaf(p3.w, k3, a3);
    }
    const float av[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = 4 * grp + r;
      q = fmaf(kk[m * kLD + n], av[r], q);
      const float a2r = av[r] * av[r];
#pragma unroll
      for (int d = 0; d < DOUT; ++d) pv[d] = fmaf(a2r, g.Sm[m * DOUTP + d], pv[d]);
      if (SAVE_A) aa[m * kLD + n] = av[r];
    }
  }
  float *rp = red + (size_t)part * NRED * kNP + n;
  rp[0] = q;
#pragma unroll
  for (int d = 0; d < DOUT; ++d) {
    rp[(1 + d) * kNP] = pm[d];
    rp[(1 + DOUT + d) * kNP] = pv[d];
  }
  __syncthreads();
  q = 0.f;
#pragma unroll
  for (int d = 0; d < DOUT; ++d) { fm[d] = 0.f; fv[d] = 0.f; }
  for (int p = 0; p < parts; ++p) {
    const float *rq = red + (size_t)p * NRED * kNP + n;
    q += rq[0];
#pragma unroll
    for (int d = 0; d < DOUT; ++d) {
      fm[d] += rq[(1 + d) * kNP];
      fv[d] += rq[(1 + DOUT + d) * kNP];
    }
  }
#pragma unroll
  for (int d = 0; d < DOUT; ++d) fv[d] = gp_var_clamp(g.sig2 - q + fv[d]);
}

// ---- reverse of one GP evaluation (SURVEY 8a note 4) ----
// Needs kk, aa from gp_forward_coop<.., true>.  gm/gv = adjoints of (fmean, fvar).
// Leaves in shared memory the GEMM operands kk, bb(=a_bar), aa(=a^2), ww and the
// right-hand extras rx; returns xin_bar for the first NEED input dims; accumulates
// the per-thread scalar sums.  `live` = this particle exists (else contributes zeros).
template <int DIN, int DOUT, int NEED>
__device__ __forceinline__ void gp_reverse_coop(const GpS<DIN, DOUT> &g, float *kk, float *aa, float *bb,
                                                float *ww, float *rx, float *red, int part, int parts, int n,
                                                const float (&xt)[GpS<DIN, DOUT>::DINP],
                                                const float (&gm)[DOUT], const float (&gv)[DOUT], bool live,
                                                float (&xinb)[NEED], float (&Lacc)[DIN], float &sw, float &sG) {
  constexpr int DINP = GpS<DIN, DOUT>::DINP, DOUTP = GpS<DIN, DOUT>::DOUTP;
  constexpr int DG = (DOUT + 3) / 4;
  float G = 0.f;
#pragma unroll
  for (int d = 0; d < DOUT; ++d) G += gv[d];
  // b = a * (S gv)
  for (int grp = part; grp < g.MG; grp += parts) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = 4 * grp + r;
      float c = 0.f;
#pragma unroll
      for (int d = 0; d < DOUT; ++d) c = fmaf(g.Sm[m * DOUTP + d], gv[d], c);
      bb[m * kLD + n] = aa[m * kLD + n] * c;
    }
  }
  if (part == 0) {
#pragma unroll
    for (int d = 0; d < DOUT; ++d) {
      rx[d * kLD + n] = gm[d];
      rx[(4 * DG + d) * kLD + n] = gv[d];
    }
#pragma unroll
    for (int j = 0; j < DIN; ++j) rx[(8 * DG + j) * kLD + n] = live ? xt[j] : 0.f;
    rx[(8 * DG + DIN) * kLD + n] = live ? 1.f : 0.f;
    sG += G;
  }
  __syncthreads();
  // kbar = alpha gm + 2 P b - 2 G a ; w = kbar * k
  float px[NEED];
#pragma unroll
  for (int j = 0; j < NEED; ++j) px[j] = 0.f;
  for (int grp = part; grp < g.MG; grp += parts) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float *P0 = g.P + (4 * grp) * g.MP;
    for (int mp = 0; mp < g.MP; mp += 4) {
      const float4 p0 = *reinterpret_cast<const float4 *>(P0 + mp);
      const float4 p1 = *reinterpret_cast<const float4 *>(P0 + g.MP + mp);
      const float4 p2 = *reinterpret_cast<const float4 *>(P0 + 2 * g.MP + mp);
      const float4 p3 = *reinterpret_cast<const float4 *>(P0 + 3 * g.MP + mp);
      const float k0 = bb[(mp + 0) * kLD + n], k1 = bb[(mp + 1) * kLD + n];
      const float k2 = bb[(mp + 2) * kLD + n], k3 = bb[(mp + 3) * kLD + n];
      a0 = fmaf(p0.x, k0, a0); a0 = fmaf(p0.y, k1, a0); a0 = fmaf(p0.z, k2, a0); a0 = fmaf(p0.w, k3, a0);
      a1 = fmaf(p1.x, k0, a1); a1 = fmaf(p1.y, k1, a1); a1 = fmaf(p1.z, k2, a1); a1 = fmaf(p1.w, k3, a1);
      a2 = fmaf(p2.x, k0, a2); a2 = fmaf(p2.y, k1, a2); a2 = fmaf(p2.z, k2, a2); a2 = fmaf(p2.w, k3, a2);
      a3 = fmaf(p3.x, k0, a3); a3 = fmaf(p3.y, k1, a3); a3 = fmaf(p3.z, k2, a3); a3 = fmaf(p3.w, k3, a3);
    }
    const float pb[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = 4 * grp + r;
      const float a = aa[m * kLD + n];
      const float k = kk[m * kLD + n];
      float kb = 2.f * pb[r] - 2.f * G * a;
#pragma unroll
      for (int d = 0; d < DOUT; ++d) kb = fmaf(g.al[m * DOUTP + d], gm[d], kb);
      const float w = kb * k;
      ww[m * kLD + n] = w;
      aa[m * kLD + n] = a * a;
      sw += w;
#pragma unroll
      for (int j = 0; j < DIN; ++j) {
        const float dl = xt[j] - g.Zt[m * DINP + j];
        const float wd = w * dl;
        if (j < NEED) px[j < NEED ? j : 0] -= wd;
        Lacc[j] = fmaf(wd, dl, Lacc[j]);
      }
    }
  }
  float *rp = red + (size_t)part * NEED * kNP + n;
#pragma unroll
  for (int j = 0; j < NEED; ++j) rp[j * kNP] = px[j];
  __syncthreads();
  // a_bar = 2 b - G k   (all parts are done reading bb)
  for (int grp = part; grp < g.MG; grp += parts) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = 4 * grp + r;
      bb[m * kLD + n] = 2.f * bb[m * kLD + n] - G * kk[m * kLD + n];
    }
  }
#pragma unroll
  for (int j = 0; j < NEED; ++j) xinb[j] = 0.f;
  for (int p = 0; p < parts; ++p) {
    const float *rq = red + (size_t)p * NEED * kNP + n;
#pragma unroll
    for (int j = 0; j < NEED; ++j) xinb[j] += rq[j * kNP];
  }
#pragma unroll
  for (int j = 0; j < NEED; ++j) xinb[j] *= g.il[j];
  __syncthreads();
}

// ---- per-step accumulation of the parameter adjoints (K-dim = the 32 particles) ----
__device__ __forceinline__ void grad_gemm(const AccLayout &L, const float *kk, const float *ab,
                                          const float *asq, const float *ww, const float *rx, float *acc) {
  const int MG = L.MG, DG = L.DG, CG = L.CG;
  for (int tile = threadIdx.x; tile < L.ntiles; tile += blockDim.x) {
    const int rg = tile / CG, cg = tile - rg * CG;
    const float *left, *right;
    if (cg < MG) { left = ab; right = kk + 4 * cg * kLD; }
    else if (cg < MG + DG) { left = kk; right = rx + 4 * (cg - MG) * kLD; }
    else if (cg < MG + 2 * DG) { left = asq; right = rx + 4 * (cg - MG) * kLD; }
    else { left = ww; right = rx + 4 * (cg - MG) * kLD; }
    left += 4 * rg * kLD;
    float c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
#pragma unroll 2
    for (int n4 = 0; n4 < kNP; n4 += 4) {
      float4 l[4], r[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        l[i] = *reinterpret_cast<const float4 *>(left + i * kLD + n4);
        r[i] = *reinterpret_cast<const float4 *>(right + i * kLD + n4);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          c[i][j] = fmaf(l[i].x, r[j].x, c[i][j]);
          c[i][j] = fmaf(l[i].y, r[j].y, c[i][j]);
          c[i][j] = fmaf(l[i].z, r[j].z, c[i][j]);
          c[i][j] = fmaf(l[i].w, r[j].w, c[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 *p = reinterpret_cast<float4 *>(acc + ((size_t)i * L.ntiles + tile) * 4);
      float4 v = *p;
      v.x += c[i][0]; v.y += c[i][1]; v.z += c[i][2]; v.w += c[i][3];
      *p = v;
    }
  }
}

// Block-wide sum of `count` per-thread floats into out[0..count) (thread 0 writes).
__device__ __forceinline__ void block_sum_store(const float *vals, int count, float *scratch, float *out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  for (int i = 0; i < count; ++i) {
    float v = vals[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) scratch[i * nw + warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < count; ++i) {
      float s = 0.f;
      for (int w = 0; w < nw; ++w) s += scratch[i * nw + w];
      out[i] = s;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float *align16(float *p) {
  return reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
}

// =====================================================================================
// backward message, forward pass: one CTA = one particle tile x one chain segment
// =====================================================================================
template <int DX, int DU, int DY>
__global__ void __launch_bounds__(256) bm_forward_kernel(Dims D, ChainTable chains, GpDev gp, const float *__restrict__ vxg,
                                                         const float *__restrict__ u, const float *__restrict__ y,
                                                         const float *__restrict__ eps_b, const float *__restrict__ z_b,
                                                         Workspace ws, float *__restrict__ part_out) {
  constexpr int DH = DX - DY, DIN = DX + DU;
  extern __shared__ __align__(16) float smem[];
  GpS<DIN, DH> g;
  float *p = g.init(smem, gp, D.M);
  float *kk = p; p += g.MP * kLD;
  float *red = p;
  const int parts = blockDim.x >> 5, part = threadIdx.x >> 5, n = threadIdx.x & 31;
  p += parts * (1 + 2 * DH) * kNP;
  float *vx = p; p += 4 * ((DX + 3) / 4);
  if (threadIdx.x < DX) vx[threadIdx.x] = vxg[threadIdx.x];
  __syncthreads();

  const Chain ch = chains.c[blockIdx.y];
  const int nl = blockIdx.x * kNP + n;          // local particle
  const bool live = nl < D.n_local;
  const int b = live ? (D.n_offset + nl) / D.S : 0;
  const float *ub = u + (size_t)b * D.T * DU;
  const float *yb = y + (size_t)b * D.T * DY;
  const size_t np = ws.npad;

  float h[DH];
  float ent = 0.f;
#pragma unroll
  for (int j = 0; j < DH; ++j) h[j] = 0.f;
  if (ch.init == 1 && live) {
    const float z = z_b[((size_t)ch.run * D.T + ch.t_hi) * D.n_local + nl];
#pragma unroll
    for (int j = 0; j < DH; ++j) h[j] = z;
  }
#pragma unroll 1
  for (int t = ch.t_hi; t >= ch.t_lo; --t) {
    float xin[DIN], xt[GpS<DIN, DH>::DINP], fm[DH], fv[DH];
#pragma unroll
    for (int j = 0; j < DH; ++j) xin[j] = h[j];
#pragma unroll
    for (int j = 0; j < DU; ++j) xin[DH + j] = ub[t * DU + j];
#pragma unroll
    for (int j = 0; j < DY; ++j) xin[DH + DU + j] = yb[t * DY + j];
    const float e = live ? eps_b[((size_t)ch.run * D.T + t) * D.n_local + nl] : 0.f;
    gp_forward_coop<DIN, DH, false>(g, kk, nullptr, red, part, parts, n, xin, xt, fm, fv);
    const bool write = writer_run(t, D.R) == ch.run;
#pragma unroll
    for (int j = 0; j < DH; ++j) {
      const float f = fv[j] + vx[j];                 // cbfssm.py:146
      h[j] = fm[j] + h[j] + e * sqrtf(f);            // :145,150
      if (write) ent += 0.5f * (kLog2PiE + logf(f));   // :154-156
    }
    if (part == 0 && live) {
      float *Hp = ws.H + (((size_t)ch.run * D.T + t) * DH) * np + nl;
#pragma unroll
      for (int j = 0; j < DH; ++j) Hp[j * np] = h[j];
    }
  }
  float v[1] = {(part == 0 && live) ? ent : 0.f};
  block_sum_store(v, 1, red, part_out + ((size_t)blockIdx.y * gridDim.x + blockIdx.x));
}

// =====================================================================================
// forward conditional rollout, forward pass: one CTA = one particle tile, T-1 steps
// =====================================================================================
template <int DX, int DU, int DY>
__global__ void __launch_bounds__(256) fw_forward_kernel(Dims D, GpDev gp, const float *__restrict__ vxg, const float *__restrict__ vyg,
                                                         const float *__restrict__ u, const float *__restrict__ y,
                                                         const float *__restrict__ eps_f, Workspace ws,
                                                         float *__restrict__ part_out) {
  constexpr int DH = DX - DY, DIN = DX + DU;
  extern __shared__ __align__(16) float smem[];
  GpS<DIN, DX> g;
  float *p = g.init(smem, gp, D.M);
  float *kk = p; p += g.MP * kLD;
  float *red = p;
  const int parts = blockDim.x >> 5, part = threadIdx.x >> 5, n = threadIdx.x & 31;
  p += parts * (1 + 2 * DX) * kNP;
  float *vx = p; p += 4 * ((DX + 3) / 4);
  float *vy = p; p += 4 * ((DX + 3) / 4);
  if (threadIdx.x < DX) { vx[threadIdx.x] = vxg[threadIdx.x]; vy[threadIdx.x] = vyg[threadIdx.x]; }
  __syncthreads();

  const int nl = blockIdx.x * kNP + n;
  const bool live = nl < D.n_local;
  const int b = live ? (D.n_offset + nl) / D.S : 0;
  const float *ub = u + (size_t)b * D.T * DU;
  const float *yb = y + (size_t)b * D.T * DY;
  const size_t np = ws.npad;
  const int nr = live ? nl : 0;

  // y_tilde[t] = [y_t, y2[t]] with y2[t] = H[writer_run(t)][t]   (cbfssm.py:95-97)
  auto load_ytil = [&](int t, float(&yt)[DX]) {
#pragma unroll
    for (int j = 0; j < DY; ++j) yt[j] = yb[t * DY + j];
    const float *Hp = ws.H + (((size_t)writer_run(t, D.R) * D.T + t) * DH) * np + nr;
#pragma unroll
    for (int j = 0; j < DH; ++j) yt[DY + j] = (live && !D.half) ? Hp[j * np] : 0.f;
  };

  float x[DX], sse[DY + 1];
#pragma unroll
  for (int j = 0; j <= DY; ++j) sse[j] = 0.f;
  float kl = 0.f;
  load_ytil(0, x);                                       // x_0 = y_tilde[:, 0]  (cbfssm.py:168)
  if (D.half) {                                          // x_0 from the recognition model (cbfssmhalf.py:103)
#pragma unroll
    for (int j = 0; j < DX; ++j) x[j] = ws.x0[(size_t)b * DX + j];
  }
#pragma unroll 1
  for (int t = 0; t < D.T; ++t) {
    if (part == 0 && live) {
      float *Xp = ws.X + ((size_t)t * DX) * np + nl;
#pragma unroll
      for (int j = 0; j < DX; ++j) Xp[j * np] = x[j];
#pragma unroll
      for (int j = 0; j < DY; ++j) { const float d = yb[t * DY + j] - x[j]; sse[j] = fmaf(d, d, sse[j]); }
    }
    if (t == D.T - 1) break;
    float xin[DIN], xt[GpS<DIN, DX>::DINP], fm[DX], fv[DX], yt[DX], xn[DX];
#pragma unroll
    for (int j = 0; j < DX; ++j) xin[j] = x[j];
#pragma unroll
    for (int j = 0; j < DU; ++j) xin[DX + j] = ub[t * DU + j];
    load_ytil(t + 1, yt);
    const float e = live ? eps_f[(size_t)t * D.n_local + nl] : 0.f;
    gp_forward_coop<DIN, DX, false>(g, kk, nullptr, red, part, parts, n, xin, xt, fm, fv);
    const bool do_cond = D.condition || (t < D.R - 1);   // cbfssm.py:227
    fw_step<DX>(x, fm, fv, yt, e, vx, vy, D.kap, do_cond, D.ncond, xn, kl);
#pragma unroll
    for (int j = 0; j < DX; ++j) x[j] = xn[j];
  }
  sse[DY] = kl;
  if (!(part == 0 && live)) {
#pragma unroll
    for (int j = 0; j <= DY; ++j) sse[j] = 0.f;
  }
  block_sum_store(sse, DY + 1, red, part_out + (size_t)blockIdx.x * (DY + 1));
}

// =====================================================================================
// reverse of the forward rollout.  Persistent CTAs loop over particle tiles.
// =====================================================================================
template <int DX, int DU, int DY>
__global__ void __launch_bounds__(256) fw_reverse_kernel(Dims D, GpDev gp, const float *__restrict__ vxg, const float *__restrict__ vyg,
                                                         const float *__restrict__ u, const float *__restrict__ y,
                                                         const float *__restrict__ eps_f, float w_ll, float w_kl,
                                                         Workspace ws, float *__restrict__ part_out) {
  constexpr int DH = DX - DY, DIN = DX + DU;
  constexpr int DG = (DX + 3) / 4, XG = (DIN + 1 + 3) / 4;
  constexpr int NRED = (1 + 2 * DX) > DX ? (1 + 2 * DX) : DX;
  extern __shared__ __align__(16) float smem[];
  GpS<DIN, DX> g;
  float *p = g.init(smem, gp, D.M);
  const AccLayout L(D.M, DIN, DX, DX);
  float *kk = p; p += g.MP * kLD;
  float *aa = p; p += g.MP * kLD;
  float *bb = p; p += g.MP * kLD;
  float *ww = p; p += g.MP * kLD;
  float *rx = p; p += (8 * DG + 4 * XG) * kLD;
  float *acc = p; p += L.nacc;
  float *red = p;
  const int parts = blockDim.x >> 5, part = threadIdx.x >> 5, n = threadIdx.x & 31;
  p += parts * NRED * kNP;
  float *vx = p; p += 4 * ((DX + 3) / 4);
  float *vy = p; p += 4 * ((DX + 3) / 4);
  if (threadIdx.x < DX) { vx[threadIdx.x] = vxg[threadIdx.x]; vy[threadIdx.x] = vyg[threadIdx.x]; }
  for (int i = threadIdx.x; i < L.nacc; i += blockDim.x) acc[i] = 0.f;
  for (int i = threadIdx.x; i < (8 * DG + 4 * XG) * kLD; i += blockDim.x) rx[i] = 0.f;
  for (int i = threadIdx.x; i < 4 * g.MP * kLD; i += blockDim.x) kk[i] = 0.f;   // kk,aa,bb,ww contiguous
  __syncthreads();

  float Lacc[DIN], sw = 0.f, sG = 0.f, vxacc[DX], vyacc[DX];
#pragma unroll
  for (int j = 0; j < DIN; ++j) Lacc[j] = 0.f;
#pragma unroll
  for (int j = 0; j < DX; ++j) { vxacc[j] = 0.f; vyacc[j] = 0.f; }

  const int ntile = ceil_div(D.n_local, kNP);
  const size_t np = ws.npad;
  for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int nl = tile * kNP + n;
    const bool live = nl < D.n_local;
    const int nr = live ? nl : 0;
    const int b = (D.n_offset + nr) / D.S;
    const float *ub = u + (size_t)b * D.T * DU;
    const float *yb = y + (size_t)b * D.T * DY;
    const bool accum = (part == 0) && live;

    float xb[DX];   // adjoint of x_{t+1}, starts as the likelihood term at T-1
    {
      const float *Xp = ws.X + ((size_t)(D.T - 1) * DX) * np + nr;
#pragma unroll
      for (int j = 0; j < DX; ++j)
        xb[j] = (j < DY && live) ? w_ll * (yb[(D.T - 1) * DY + (j < DY ? j : 0)] - Xp[j * np]) / vy[j] : 0.f;
    }
#pragma unroll 1
    for (int t = D.T - 2; t >= 0; --t) {
      float x[DX], xin[DIN], xt[GpS<DIN, DX>::DINP], fm[DX], fv[DX], yt[DX];
      const float *Xp = ws.X + ((size_t)t * DX) * np + nr;
#pragma unroll
      for (int j = 0; j < DX; ++j) { x[j] = live ? Xp[j * np] : 0.f; xin[j] = x[j]; }
#pragma unroll
      for (int j = 0; j < DU; ++j) xin[DX + j] = ub[t * DU + j];
#pragma unroll
      for (int j = 0; j < DY; ++j) yt[j] = yb[(t + 1) * DY + j];
      {
        const float *Hp = ws.H + (((size_t)writer_run(t + 1, D.R) * D.T + (t + 1)) * DH) * np + nr;
#pragma unroll
        for (int j = 0; j < DH; ++j) yt[DY + j] = (live && !D.half) ? Hp[j * np] : 0.f;
      }
      const float e = live ? eps_f[(size_t)t * D.n_local + nl] : 0.f;
      gp_forward_coop<DIN, DX, true>(g, kk, aa, red, part, parts, n, xin, xt, fm, fv);
      const bool do_cond = D.condition || (t < D.R - 1);
      float fmb[DX], fvb[DX], ytb[DX];
      fw_step_adjoint<DX>(x, fm, fv, yt, e, vx, vy, D.kap, do_cond, D.ncond, w_kl, xb, fmb, fvb, ytb, vxacc, vyacc, accum);
      if (!live) {   // padding particles must not feed the parameter adjoints
#pragma unroll
        for (int j = 0; j < DX; ++j) { fmb[j] = 0.f; fvb[j] = 0.f; }
      }
      if (accum) {
        float *Yp = ws.Yb + ((size_t)(t + 1) * DH) * np + nl;
#pragma unroll
        for (int j = 0; j < DH; ++j) Yp[j * np] = ytb[DY + j];
      }
      float xinb[DX];
      gp_reverse_coop<DIN, DX, DX>(g, kk, aa, bb, ww, rx, red, part, parts, n, xt, fmb, fvb, live, xinb, Lacc, sw, sG);
      grad_gemm(L, kk, bb, aa, ww, rx, acc);
#pragma unroll
      for (int j = 0; j < DX; ++j) {
        float lg = 0.f;
        if (j < DY && live) lg = w_ll * (yb[t * DY + (j < DY ? j : 0)] - x[j]) / vy[j];
        xb[j] = xinb[j] + fmb[j] + lg;
      }
      __syncthreads();
    }
    if (accum && D.half) {   // CBFSSMHALF: adjoint of the recognition model's x_0 (all dims)
#pragma unroll
      for (int j = 0; j < DX; ++j) ws.x0b[j * np + nl] = xb[j];
    } else if (accum) {      // x_0 = y_tilde[:,0]: hidden part flows to y2[0]   (cbfssm.py:168)
      float *Yp = ws.Yb + nl;
#pragma unroll
      for (int j = 0; j < DH; ++j) Yp[j * np] = xb[DY + j];
    }
  }
  // ---- CTA partial: tile accumulators, then scalars [L_j | sw | sG | vx_bar | vy_bar] ----
  __syncthreads();
  float *out = part_out + (size_t)blockIdx.x * L.slot();
  for (int i = threadIdx.x; i < L.nacc; i += blockDim.x) out[i] = acc[i];
  float sc[DIN + 2 + 2 * DX];
#pragma unroll
  for (int j = 0; j < DIN; ++j) sc[j] = Lacc[j];
  sc[DIN] = sw; sc[DIN + 1] = sG;
#pragma unroll
  for (int j = 0; j < DX; ++j) { sc[DIN + 2 + j] = vxacc[j]; sc[DIN + 2 + DX + j] = vyacc[j]; }
  block_sum_store(sc, DIN + 2 + 2 * DX, red, out + L.scal_off());
}

// =====================================================================================
// reverse of the backward-message chains.  Work item = (particle tile, chain).
// =====================================================================================
template <int DX, int DU, int DY>
__global__ void __launch_bounds__(256) bm_reverse_kernel(Dims D, ChainTable chains, GpDev gp, const float *__restrict__ vxg,
                                                         const float *__restrict__ u, const float *__restrict__ y,
                                                         const float *__restrict__ eps_b, const float *__restrict__ z_b,
                                                         float w_en, Workspace ws, float *__restrict__ part_out) {
  constexpr int DH = DX - DY, DIN = DX + DU;
  constexpr int DG = (DH + 3) / 4, XG = (DIN + 1 + 3) / 4;
  constexpr int NRED = 1 + 2 * DH;
  extern __shared__ __align__(16) float smem[];
  GpS<DIN, DH> g;
  float *p = g.init(smem, gp, D.M);
  const AccLayout L(D.M, DIN, DH, DX);
  float *kk = p; p += g.MP * kLD;
  float *aa = p; p += g.MP * kLD;
  float *bb = p; p += g.MP * kLD;
  float *ww = p; p += g.MP * kLD;
  float *rx = p; p += (8 * DG + 4 * XG) * kLD;
  float *acc = p; p += L.nacc;
  float *red = p;
  const int parts = blockDim.x >> 5, part = threadIdx.x >> 5, n = threadIdx.x & 31;
  p += parts * NRED * kNP;
  float *vx = p; p += 4 * ((DX + 3) / 4);
  if (threadIdx.x < DX) vx[threadIdx.x] = vxg[threadIdx.x];
  for (int i = threadIdx.x; i < L.nacc; i += blockDim.x) acc[i] = 0.f;
  for (int i = threadIdx.x; i < (8 * DG + 4 * XG) * kLD; i += blockDim.x) rx[i] = 0.f;
  for (int i = threadIdx.x; i < 4 * g.MP * kLD; i += blockDim.x) kk[i] = 0.f;
  __syncthreads();

  float Lacc[DIN], sw = 0.f, sG = 0.f, vxacc[DX];
#pragma unroll
  for (int j = 0; j < DIN; ++j) Lacc[j] = 0.f;
#pragma unroll
  for (int j = 0; j < DX; ++j) vxacc[j] = 0.f;

  const int ntile = ceil_div(D.n_local, kNP);
  const int nitem = ntile * chains.count;
  const size_t np = ws.npad;
  for (int item = blockIdx.x; item < nitem; item += gridDim.x) {
    const int tile = item % ntile;
    const Chain ch = chains.c[item / ntile];
    const int nl = tile * kNP + n;
    const bool live = nl < D.n_local;
    const int nr = live ? nl : 0;
    const int b = (D.n_offset + nr) / D.S;
    const float *ub = u + (size_t)b * D.T * DU;
    const float *yb = y + (size_t)b * D.T * DY;
    const bool accum = (part == 0) && live;

    float hb[DH];
#pragma unroll
    for (int j = 0; j < DH; ++j) hb[j] = 0.f;
#pragma unroll 1
    for (int t = ch.t_lo; t <= ch.t_hi; ++t) {
      float hid[DH], xin[DIN], xt[GpS<DIN, DH>::DINP], fm[DH], fv[DH];
      if (t == ch.t_hi) {
        const float z = (ch.init == 1 && live) ? z_b[((size_t)ch.run * D.T + t) * D.n_local + nl] : 0.f;
#pragma unroll
        for (int j = 0; j < DH; ++j) hid[j] = z;
      } else {
        const float *Hp = ws.H + (((size_t)ch.run * D.T + (t + 1)) * DH) * np + nr;
#pragma unroll
        for (int j = 0; j < DH; ++j) hid[j] = live ? Hp[j * np] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < DH; ++j) xin[j] = hid[j];
#pragma unroll
      for (int j = 0; j < DU; ++j) xin[DH + j] = ub[t * DU + j];
#pragma unroll
      for (int j = 0; j < DY; ++j) xin[DH + DU + j] = yb[t * DY + j];
      const float e = live ? eps_b[((size_t)ch.run * D.T + t) * D.n_local + nl] : 0.f;
      gp_forward_coop<DIN, DH, true>(g, kk, aa, red, part, parts, n, xin, xt, fm, fv);
      const bool write = writer_run(t, D.R) == ch.run;
      float ob[DH], fvb[DH];
      const float *Yp = ws.Yb + ((size_t)t * DH) * np + nr;
#pragma unroll
      for (int j = 0; j < DH; ++j) {
        const float f = fv[j] + vx[j];
        float o = hb[j], fb = 0.f;
        if (write) {
          o += live ? Yp[j * np] : 0.f;
          fb = live ? w_en * 0.5f / f : 0.f;
        }
        fb += o * e * 0.5f * rsqrtf(f);
        ob[j] = o; fvb[j] = fb;
        if (accum) vxacc[j] += fb;
      }
      float xinb[DH];
      gp_reverse_coop<DIN, DH, DH>(g, kk, aa, bb, ww, rx, red, part, parts, n, xt, ob, fvb, live, xinb, Lacc, sw, sG);
      grad_gemm(L, kk, bb, aa, ww, rx, acc);
#pragma unroll
      for (int j = 0; j < DH; ++j) hb[j] = xinb[j] + ob[j];
      __syncthreads();
    }
  }
  __syncthreads();
  float *out = part_out + (size_t)blockIdx.x * L.slot();
  for (int i = threadIdx.x; i < L.nacc; i += blockDim.x) out[i] = acc[i];
  float sc[DIN + 2 + 2 * DX];
#pragma unroll
  for (int j = 0; j < DIN; ++j) sc[j] = Lacc[j];
  sc[DIN] = sw; sc[DIN + 1] = sG;
#pragma unroll
  for (int j = 0; j < DX; ++j) { sc[DIN + 2 + j] = vxacc[j]; sc[DIN + 2 + DX + j] = 0.f; }
  block_sum_store(sc, DIN + 2 + 2 * DX, red, out + L.scal_off());
}

}  // namespace cbf
