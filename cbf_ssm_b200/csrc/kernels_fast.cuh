// Register-resident kernel path for small M (compile-time M <= 32): one thread = one particle.
//
// The resident GP operands (P = K_zz^-1, Z/ell, alpha, S) sit in shared memory and are read
// with warp-uniform (broadcast) vector loads; the particle's M-vectors k, a = P k, b live in
// registers with fully unrolled, statically indexed loops; particle state streams through
// registers with coalesced [t][dim][n] global accesses.  No block-level synchronisation on
// the rollout itself.  The parameter adjoints (P_bar += a_bar k^T, ...) are accumulated per
// WARP: each lane stages its particle's vectors into a warp-private K-major shared tile, then
// the 32 lanes own TR x TC register tiles of the outer-product accumulation over the warp's
// 32 particles -- only __syncwarp() is needed and the accumulators stay in registers for the
// whole launch.  (TR, TC) is chosen at compile time per (M, Din, Dout).
//
// Mathematics: SURVEY.md 8a notes 1-5; verified in float64 by oracle/kernel_math.py.
#pragma once
#include "common.cuh"
#include "step_math.cuh"

namespace cbf {

#ifndef CBF_FAST_THREADS
#define CBF_FAST_THREADS 128
#endif
constexpr int kFastThreads = CBF_FAST_THREADS;
constexpr int kFastWarps = kFastThreads / 32;
// Resident CTAs per SM the reverse kernels are compiled for: 3 (168 registers) for small state dims; from
// dx = 12 the per-thread vectors (DOUT, DIN-sized) spill at 168 registers (13/6/7: 1 KB per thread), so those
// instantiations get 2 CTAs and 255 registers (Voliro shape: fw_reverse 13.8 -> 7.2 ms).
#ifdef CBF_REV_MINBLOCKS
template <int DX> constexpr int rev_minblocks() { return CBF_REV_MINBLOCKS; }
#else
template <int DX> constexpr int rev_minblocks() { return DX >= 12 ? 2 : 3; }
#endif
constexpr int kSLD = 36;   // staging row stride (floats): 32 lanes + 4, == 4 mod 32

struct TileCfg {
  int TR, TC, rounds;
};

constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }

// Compiler-only memory barrier.  The resident operands in shared memory are loop-invariant,
// so without it the compiler hoists / CSEs hundreds of shared loads out of the time loop
// (and from the first contraction into the second) and then spills them.
__device__ __forceinline__ void compiler_fence() { asm volatile("" ::: "memory"); }
// 2^x as one MUFU.EX2 (exp2f() adds range scaling: two FMULs and a predicate per call).
// The argument is -0.5*log2(e)*d^2 + log2(sigma^2) <= log2(sigma^2); results below the
// float32 normal range flush to zero, which is what the kernel value is there anyway.
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Tile shape of the per-warp accumulation (packed FP32: two columns per FFMA2).  Per 4
// particles a lane issues 2*TR*TC FFMA2 and TR+TC LDS.128; a non-uniform LDS.128 costs >= 4
// shared-memory wavefronts (measured) while the SM retires one wavefront per cycle against four
// issue slots, so a load is weighted 4.  TC must be 2 or 6: the 8 lanes of a quarter-warp read
// right-operand pair-rows TC/2 apart with a pair-row stride of 68 floats (17 16-byte granules),
// which is bank-conflict-free only if TC/2 is odd (TC = 2 with the old unpacked layout measured
// 8 wavefronts per load instead of 4).
constexpr TileCfg pick_tiles(int M, int Din, int Dout) {
  TileCfg best{5, 6, 1000};
  double best_cost = 1e30;
  for (int tr = 2; tr <= 8; ++tr)
    for (int tc = 2; tc <= 6; tc += 4) {
#ifdef CBF_FORCE_TC6
      if (tc != 6) continue;
#endif
      const int rg = cdiv(M, tr), cg = cdiv(M, tc) + 2 * cdiv(Dout, tc) + cdiv(Din + 1, tc);
      const int rounds = cdiv(rg * cg, 32);
      if (rounds * tr * tc > 60) continue;   // accumulator registers per lane
      const double cost = rounds * (2.0 * tr * tc + 4.0 * (tr + tc));
      if (cost < best_cost) { best_cost = cost; best = TileCfg{tr, tc, rounds}; }
    }
  return best;
}

// Packed FP32 FMA (sm_100 FFMA2): d.xy = a.xy * s + c.xy with the scalar s broadcast (ptxas folds
// the {s,s} pack into the .F32 operand form).  Same FLOP rate as FFMA, half the issue slots
// (tools/microbench/ffma2.cu: 71 vs 74 TFLOP/s on B200).
__device__ __forceinline__ unsigned long long pack2(float x, float y) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &x, float &y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float hsum2(unsigned long long v) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
  return x + y;
}
__device__ __forceinline__ unsigned long long ffma2_bcast(float ax, float ay, float s, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pack2(ax, ay)), "l"(pack2(s, s)), "l"(c));
  return d;
}

// ---- resident operands in the constant bank -------------------------------------------
// Measured on B200 (profiles/r01b_fast_bm_reverse_full.txt + source page): a warp-uniform
// LDS.128 costs 2 shared-memory wavefronts (8 useful bytes per wavefront), so streaming P, Z/ell,
// alpha and S through broadcast shared loads made the register path shared-memory bound
// (~77 % of shared-pipe cycles).  With CBF_CONST_OPERANDS the packed operands of the GP a
// kernel uses live in a __constant__ array instead and every use is an FFMA/FADD with a
// c[bank][imm] or uniform-register operand: no shared-memory traffic.
// The array is per translation unit; it is refreshed before each launch by a device-to-device
// cudaMemcpyToSymbolAsync on the caller's stream (see launch_fast.cuh), so calls of one
// instantiation must not overlap on different streams.
#ifndef CBF_CONST_OPERANDS
#define CBF_CONST_OPERANDS 1
#endif
constexpr bool kConstOps = CBF_CONST_OPERANDS != 0;
constexpr int kConstFloats = 2048;              // per GP slot (8 KB)
static __constant__ float c_ops[2][kConstFloats];      // [0] forward-rollout GP, [1] backward-message GP

template <int SLOT, int M, int DIN, int DOUT>
struct CO {
  static constexpr int MP = (M + 3) / 4 * 4;   // row stride of P: rows start 16-byte aligned (LDCU.128)
  static constexpr int DINE = (DIN + 1) / 2 * 2, DOUTE = (DOUT + 1) / 2 * 2;   // even row strides: pair operands
  static constexpr int oP = 0, oZ = M * MP, oA = oZ + M * DINE, oS = oA + M * DOUTE, oI = oS + M * DOUTE;
  static constexpr int oSig = oI + DIN, oLs = oSig + 1, TOTAL = oLs + 1;
  static_assert(TOTAL <= kConstFloats, "operands exceed the constant slot");
  static __device__ __forceinline__ float P(int m, int mp) { return c_ops[SLOT][oP + m * MP + mp]; }
  // pair j2 of row m of -Z/ell (zero padded), alpha, S: operands of FADD2 / FFMA2
  static __device__ __forceinline__ unsigned long long nZ2(int m, int j2) {
    return pack2(c_ops[SLOT][oZ + m * DINE + 2 * j2], c_ops[SLOT][oZ + m * DINE + 2 * j2 + 1]);
  }
  static __device__ __forceinline__ unsigned long long al2(int m, int d2) {
    return pack2(c_ops[SLOT][oA + m * DOUTE + 2 * d2], c_ops[SLOT][oA + m * DOUTE + 2 * d2 + 1]);
  }
  static __device__ __forceinline__ unsigned long long S2(int m, int d2) {
    return pack2(c_ops[SLOT][oS + m * DOUTE + 2 * d2], c_ops[SLOT][oS + m * DOUTE + 2 * d2 + 1]);
  }
  static __device__ __forceinline__ float nZ(int m, int j) { return c_ops[SLOT][oZ + m * DINE + j]; }
  static __device__ __forceinline__ float al(int m, int d) { return c_ops[SLOT][oA + m * DOUTE + d]; }
  static __device__ __forceinline__ float S(int m, int d) { return c_ops[SLOT][oS + m * DOUTE + d]; }
  static __device__ __forceinline__ float il(int j) { return c_ops[SLOT][oI + j]; }
  static __device__ __forceinline__ float sig2() { return c_ops[SLOT][oSig]; }
  static __device__ __forceinline__ float lsig() { return c_ops[SLOT][oLs]; }
};

// Packs one GP's operands in the CO<> layout (runtime sizes) into global scratch.  P is symmetrised
// and its rows are zero-padded to MP columns; Z is stored negated and scaled (-Z/ell).
static __global__ void pack_const_kernel(GpDev g, int M, int DIN, int DOUT, float *__restrict__ out) {
  const int MP = (M + 3) / 4 * 4, DINE = (DIN + 1) / 2 * 2, DOUTE = (DOUT + 1) / 2 * 2;
  const int oZ = M * MP, oA = oZ + M * DINE, oS = oA + M * DOUTE, oI = oS + M * DOUTE, oSig = oI + DIN;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  for (int i = tid; i < M * MP; i += nt) {
    const int a = i / MP, b = i % MP;
    out[i] = (b < M) ? 0.5f * (g.P[a * M + b] + g.P[b * M + a]) : 0.f;
  }
  for (int i = tid; i < M * DINE; i += nt) {
    const int m = i / DINE, j = i % DINE;
    out[oZ + i] = (j < DIN) ? -g.Z[m * DIN + j] / g.ell[j] : 0.f;
  }
  for (int i = tid; i < M * DOUTE; i += nt) {
    const int m = i / DOUTE, d = i % DOUTE;
    out[oA + i] = (d < DOUT) ? g.alpha[m * DOUT + d] : 0.f;
    out[oS + i] = (d < DOUT) ? g.S[m * DOUT + d] : 0.f;
  }
  for (int i = tid; i < DIN; i += nt) out[oI + i] = 1.f / g.ell[i];
  if (tid == 0) { out[oSig] = g.sig2[0]; out[oSig + 1] = log2f(g.sig2[0]); }
}

// a = P k with P from the constant bank.  Two output rows per FFMA2: the accumulator pair
// (a_2i, a_2i+1) takes the constant pair (P[mp][2i], P[mp][2i+1]) (= column pair of the symmetric P,
// fetched four at a time by LDCU.128 into uniform registers) times the broadcast k[mp]:
// M*MP/2 FFMA2 + M*MP/4 LDCU instead of M*M FFMA.
// With ADD_BASE the result starts from base[] (a += P k), which lets the caller fold an
// axpy into the contraction and end the live range of its operand early.
template <int SLOT, int M, int DIN, int DOUT, int MP, bool ADD_BASE = false>
__device__ __forceinline__ void matvec_const(const float (&k)[MP], float (&a)[MP]) {
  using C = CO<SLOT, M, DIN, DOUT>;
  static_assert(MP == C::MP, "padded sizes must agree");
  unsigned long long acc[MP / 2];
#pragma unroll
  for (int i = 0; i < MP / 2; ++i) acc[i] = ADD_BASE ? pack2(a[2 * i], a[2 * i + 1]) : 0ull;
#pragma unroll
  for (int mp = 0; mp < M; ++mp) {
#pragma unroll
    for (int i = 0; i < (M + 1) / 2; ++i)
      acc[i] = ffma2_bcast(C::P(mp, 2 * i), C::P(mp, 2 * i + 1), k[mp], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < MP / 2; ++i) unpack2(acc[i], a[2 * i], a[2 * i + 1]);
}

// Shared-memory image of one GP's operands, compile-time sizes.
template <int M, int DIN, int DOUT, int SLOT>
struct GpF {
  static constexpr int MP = (M + 3) / 4 * 4;
  static constexpr int DINP = (DIN + 3) / 4 * 4;
  static constexpr int DOUTP = (DOUT + 3) / 4 * 4;
  static constexpr int FLOATS = kConstOps ? 0 : MP * MP + MP * (DINP + 2 * DOUTP) + DINP + 4;
  using C = CO<SLOT, M, DIN, DOUT>;
  const float *P, *Zt, *al, *Sm, *il;
  float sig2, lsig;

  __device__ float *init(float *base, const GpDev &g) {
    if constexpr (kConstOps) {
      P = Zt = al = Sm = il = nullptr;
      sig2 = C::sig2();
      lsig = C::lsig();
      return base;
    }
    float *Pw = base; base += MP * MP;
    float *Zw = base; base += MP * DINP;
    float *aw = base; base += MP * DOUTP;
    float *Sw = base; base += MP * DOUTP;
    float *iw = base; base += DINP + 4;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < MP * MP; i += nt) {
      const int r = i / MP, c = i % MP;
      Pw[i] = (r < M && c < M) ? g.P[r * M + c] : 0.f;
    }
    for (int i = tid; i < MP * DINP; i += nt) {
      const int r = i / DINP, c = i % DINP;
      Zw[i] = (r < M && c < DIN) ? g.Z[r * DIN + c] / g.ell[c] : 0.f;
    }
    for (int i = tid; i < MP * DOUTP; i += nt) {
      const int r = i / DOUTP, c = i % DOUTP;
      const bool ok = (r < M && c < DOUT);
      aw[i] = ok ? g.alpha[r * DOUT + c] : 0.f;
      Sw[i] = ok ? g.S[r * DOUT + c] : 0.f;
    }
    for (int i = tid; i < DINP; i += nt) iw[i] = (i < DIN) ? 1.f / g.ell[i] : 0.f;
    P = Pw; Zt = Zw; al = aw; Sm = Sw; il = iw;
    sig2 = g.sig2[0];
    lsig = log2f(sig2);
    return base;
  }
};

template <int N>
__device__ __forceinline__ void ld_row(const float *__restrict__ p, float (&v)[N]) {
  static_assert(N % 4 == 0, "rows are padded to float4");
#pragma unroll
  for (int i = 0; i < N; i += 4) {
    const float4 t = *reinterpret_cast<const float4 *>(p + i);
    v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
  }
}

// a = P k for the thread's particle, 4 rows at a time (independent FMA chains).
template <int MP>
__device__ __forceinline__ void matvec_fast(const float *__restrict__ P, const float (&k)[MP], float (&a)[MP]) {
#pragma unroll
  for (int m0 = 0; m0 < MP; m0 += 4) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int mp = 0; mp < MP; mp += 4) {
      const float4 p0 = *reinterpret_cast<const float4 *>(P + (m0 + 0) * MP + mp);
      const float4 p1 = *reinterpret_cast<const float4 *>(P + (m0 + 1) * MP + mp);
      const float4 p2 = *reinterpret_cast<const float4 *>(P + (m0 + 2) * MP + mp);
      const float4 p3 = *reinterpret_cast<const float4 *>(P + (m0 + 3) * MP + mp);
      a0 = fmaf(p0.x, k[mp], a0); a0 = fmaf(p0.y, k[mp + 1], a0); a0 = fmaf(p0.z, k[mp + 2], a0); a0 = fmaf(p0.w, k[mp + 3], a0);
      a1 = fmaf(p1.x, k[mp], a1); a1 = fmaf(p1.y, k[mp + 1], a1); a1 = fmaf(p1.z, k[mp + 2], a1); a1 = fmaf(p1.w, k[mp + 3], a1);
      a2 = fmaf(p2.x, k[mp], a2); a2 = fmaf(p2.y, k[mp + 1], a2); a2 = fmaf(p2.z, k[mp + 2], a2); a2 = fmaf(p2.w, k[mp + 3], a2);
      a3 = fmaf(p3.x, k[mp], a3); a3 = fmaf(p3.y, k[mp + 1], a3); a3 = fmaf(p3.z, k[mp + 2], a3); a3 = fmaf(p3.w, k[mp + 3], a3);
    }
    a[m0] = a0; a[m0 + 1] = a1; a[m0 + 2] = a2; a[m0 + 3] = a3;
  }
}

// One sparse-GP evaluation (gp_tf.py:132-161) for the thread's particle.
template <int M, int DIN, int DOUT, int SLOT>
__device__ __forceinline__ void gp_forward_fast(const GpF<M, DIN, DOUT, SLOT> &g, const float (&xin)[DIN],
                                                float (&xt)[GpF<M, DIN, DOUT, SLOT>::DINP],
                                                float (&k)[GpF<M, DIN, DOUT, SLOT>::MP],
                                                float (&a)[GpF<M, DIN, DOUT, SLOT>::MP], float (&fm)[DOUT],
                                                float (&fv)[DOUT]) {
  using G = GpF<M, DIN, DOUT, SLOT>;
  constexpr int MP = G::MP, DINP = G::DINP, DOUTP = G::DOUTP;
  if constexpr (kConstOps) {
    using C = typename G::C;
    constexpr int J2 = C::DINE / 2, D2 = C::DOUTE / 2;
#pragma unroll
    for (int j = 0; j < DINP; ++j) xt[j] = (j < DIN) ? xin[j < DIN ? j : 0] * C::il(j < DIN ? j : 0) : 0.f;
    unsigned long long x2[J2], fm2[D2], fv2[D2];
#pragma unroll
    for (int j = 0; j < J2; ++j) x2[j] = pack2(xt[2 * j], xt[2 * j + 1]);
#pragma unroll
    for (int d = 0; d < D2; ++d) { fm2[d] = 0ull; fv2[d] = 0ull; }
#pragma unroll
    for (int m = 0; m < MP; ++m) {
      if (m < M) {
        unsigned long long acc = 0ull;   // (sum over even j, sum over odd j) of delta^2
#pragma unroll
        for (int j = 0; j < J2; ++j) {
          const unsigned long long dl = add2(x2[j], C::nZ2(m < M ? m : 0, j));
          acc = fma2(dl, dl, acc);
        }
        k[m] = fast_exp2(fmaf(kNegHalfLog2e, hsum2(acc), g.lsig));
        const unsigned long long kk = pack2(k[m], k[m]);
#pragma unroll
        for (int d = 0; d < D2; ++d) fm2[d] = fma2(C::al2(m < M ? m : 0, d), kk, fm2[d]);
      } else {
        k[m] = 0.f;
      }
    }
    matvec_const<SLOT, M, DIN, DOUT, MP>(k, a);
    float q = 0.f;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      q = fmaf(k[m], a[m], q);
      const float a2 = a[m] * a[m];
      const unsigned long long aa = pack2(a2, a2);
#pragma unroll
      for (int d = 0; d < D2; ++d) fv2[d] = fma2(C::S2(m, d), aa, fv2[d]);
    }
#pragma unroll
    for (int d = 0; d < D2; ++d) {
      float m0, m1, v0, v1;
      unpack2(fm2[d], m0, m1);
      unpack2(fv2[d], v0, v1);
      fm[2 * d] = m0; fv[2 * d] = gp_var_clamp(g.sig2 - q + v0);
      if (2 * d + 1 < DOUT) { fm[2 * d + 1 < DOUT ? 2 * d + 1 : 0] = m1; fv[2 * d + 1 < DOUT ? 2 * d + 1 : 0] = gp_var_clamp(g.sig2 - q + v1); }
    }
    return;
  }
  {
    float il[DINP];
    ld_row<DINP>(g.il, il);
#pragma unroll
    for (int j = 0; j < DINP; ++j) xt[j] = (j < DIN) ? xin[j < DIN ? j : 0] * il[j] : 0.f;
  }
#pragma unroll
  for (int d = 0; d < DOUT; ++d) fm[d] = 0.f;
#pragma unroll
  for (int m = 0; m < MP; ++m) {
    if (m < M) {
      float z[DINP];
      ld_row<DINP>(g.Zt + m * DINP, z);
      float d2 = 0.f;
#pragma unroll
      for (int j = 0; j < DIN; ++j) { const float e = xt[j] - z[j]; d2 = fmaf(e, e, d2); }
      k[m] = fast_exp2(fmaf(kNegHalfLog2e, d2, g.lsig));
      float al[DOUTP];
      ld_row<DOUTP>(g.al + m * DOUTP, al);
#pragma unroll
      for (int d = 0; d < DOUT; ++d) fm[d] = fmaf(k[m], al[d], fm[d]);
    } else {
      k[m] = 0.f;
    }
  }
  matvec_fast<MP>(g.P, k, a);
  float q = 0.f;
#pragma unroll
  for (int d = 0; d < DOUT; ++d) fv[d] = 0.f;
#pragma unroll
  for (int m = 0; m < M; ++m) {
    q = fmaf(k[m], a[m], q);
    const float a2 = a[m] * a[m];
    float S[DOUTP];
    ld_row<DOUTP>(g.Sm + m * DOUTP, S);
#pragma unroll
    for (int d = 0; d < DOUT; ++d) fv[d] = fmaf(a2, S[d], fv[d]);
  }
#pragma unroll
  for (int d = 0; d < DOUT; ++d) fv[d] = gp_var_clamp(g.sig2 - q + fv[d]);
}

// Saved GP evaluation of one particle-step: the forward kernels can leave (k, a = P k, fmean, fvar) in HBM as
// NPL float4 planes of npad particles (plane pl of evaluation slot e at (e * NPL + pl) * npad + n: a warp moves
// 512 contiguous bytes per plane), and the reverse kernels then load them instead of recomputing the
// evaluation -- about a quarter of their FMA-pipe cycles (the kernel vector and the M x M contraction) for
// 16 * NPL bytes per evaluation each way; the kernels sit at a few % of the HBM roof, so the trade is free
// until the extra workspace no longer fits (api.cu decides).
template <int MP, int DOUT>
struct SavedEval {
  static constexpr int NK = MP / 4, NF = (2 * DOUT + 3) / 4, NPL = 2 * NK + NF;
  static_assert(MP % 4 == 0, "MP is a multiple of 4");
  static __device__ __forceinline__ void store(float4 *__restrict__ base, size_t slot, size_t np, int nl,
                                               const float (&k)[MP], const float (&a)[MP], const float (&fm)[DOUT],
                                               const float (&fv)[DOUT]) {
    float4 *p = base + slot * NPL * np + nl;
#pragma unroll
    for (int i = 0; i < NK; ++i) p[(size_t)i * np] = make_float4(k[4 * i], k[4 * i + 1], k[4 * i + 2], k[4 * i + 3]);
#pragma unroll
    for (int i = 0; i < NK; ++i) p[(size_t)(NK + i) * np] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
    float f[4 * NF];
#pragma unroll
    for (int d = 0; d < 4 * NF; ++d) f[d] = 0.f;
#pragma unroll
    for (int d = 0; d < DOUT; ++d) { f[d] = fm[d]; f[DOUT + d] = fv[d]; }
#pragma unroll
    for (int i = 0; i < NF; ++i) p[(size_t)(2 * NK + i) * np] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
  }
  static __device__ __forceinline__ void load_ka(const float4 *__restrict__ base, size_t slot, size_t np, int nl,
                                                 float (&k)[MP], float (&a)[MP]) {
    const float4 *p = base + slot * NPL * np + nl;
#pragma unroll
    for (int i = 0; i < NK; ++i) {
      const float4 v = p[(size_t)i * np];
      k[4 * i] = v.x; k[4 * i + 1] = v.y; k[4 * i + 2] = v.z; k[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < NK; ++i) {
      const float4 v = p[(size_t)(NK + i) * np];
      a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
    }
  }
  static __device__ __forceinline__ void load_f(const float4 *__restrict__ base, size_t slot, size_t np, int nl,
                                                float4 (&q)[NF]) {
    const float4 *p = base + slot * NPL * np + nl;
#pragma unroll
    for (int i = 0; i < NF; ++i) q[i] = p[(size_t)(2 * NK + i) * np];
  }
  static __device__ __forceinline__ void unpack_f(const float4 (&q)[NF], float (&fm)[DOUT], float (&fv)[DOUT]) {
    float f[4 * NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) { f[4 * i] = q[i].x; f[4 * i + 1] = q[i].y; f[4 * i + 2] = q[i].z; f[4 * i + 3] = q[i].w; }
#pragma unroll
    for (int d = 0; d < DOUT; ++d) { fm[d] = f[d]; fv[d] = f[DOUT + d]; }
  }
};

// x~ = x / ell as gp_forward_fast forms it (for the reverse kernels when the evaluation itself is loaded)
template <int M, int DIN, int DOUT, int SLOT>
__device__ __forceinline__ void gp_scale_input(const GpF<M, DIN, DOUT, SLOT> &g, const float (&xin)[DIN],
                                               float (&xt)[GpF<M, DIN, DOUT, SLOT>::DINP]) {
  using G = GpF<M, DIN, DOUT, SLOT>;
  constexpr int DINP = G::DINP;
  if constexpr (kConstOps) {
    using C = typename G::C;
#pragma unroll
    for (int j = 0; j < DINP; ++j) xt[j] = (j < DIN) ? xin[j < DIN ? j : 0] * C::il(j < DIN ? j : 0) : 0.f;
  } else {
    float il[DINP];
    ld_row<DINP>(g.il, il);
#pragma unroll
    for (int j = 0; j < DINP; ++j) xt[j] = (j < DIN) ? xin[j < DIN ? j : 0] * il[j] : 0.f;
  }
}

// ---- per-warp staging tile + accumulation of the parameter adjoints ----
// Warp-private staging tile, rewritten every step by the 32 lanes (lane = particle n):
//   right matrix, column space of the accumulators  [k | g_mean | g_var | x~,1] (blocks padded to
//     TC): stored pair-interleaved, element (col, n) at (col/2)*kPS + 2n + (col&1), so one
//     LDS.128 yields two particles x two adjacent columns = two FFMA2 operand pairs;
//   left arrays [a_bar | k | a^2 | w], K-major: element (m, n) at which*LSTR + m*kSLD + n; LSTR = 8 mod 32
//     floats, so the four arrays start 8 banks apart: the lanes of a quarter warp (same row group, different
//     arrays) hit disjoint banks with their 128-bit loads (a_bar and a^2 collided before: 22 % of the
//     shared-memory wavefronts were conflicts).
constexpr int kPS = 68;   // pair-row stride (floats): 64 + 4
#ifndef CBF_ACC_UNROLL
#define CBF_ACC_UNROLL 2
#endif
constexpr int kAccUnroll = CBF_ACC_UNROLL;   // unroll of the 8-iteration particle loop of the accumulation
template <int M, int DIN, int DOUT>
struct WarpAcc {
  static constexpr TileCfg TCFG = pick_tiles(M, DIN, DOUT);
  static constexpr int TR = TCFG.TR, TC = TCFG.TC, ROUNDS = TCFG.rounds, TC2 = TC / 2;
  static_assert(TC % 2 == 0, "packed accumulation needs an even tile width");
  static constexpr int RG = cdiv(M, TR), CGk = cdiv(M, TC), CGd = cdiv(DOUT, TC), CGx = cdiv(DIN + 1, TC);
  static constexpr int CG = CGk + 2 * CGd + CGx, NTILES = RG * CG, NCOLS = CG * TC;
  static constexpr int colK = 0, colGm = CGk * TC, colGv = colGm + CGd * TC, colX = colGv + CGd * TC;
  static constexpr int LROWS = RG * TR;
  static constexpr int LSTR = LROWS * kSLD + (8 - (LROWS * kSLD) % 32 + 32) % 32;
  static constexpr int LEFT_OFF = (NCOLS / 2) * kPS;
  static constexpr int FLOATS = LEFT_OFF + 4 * LSTR;
  static constexpr int NACC = NTILES * TR * TC;
  enum { L_AB = 0, L_K = 1, L_ASQ = 2, L_W = 3 };

  unsigned long long acc[ROUNDS][TR][TC2];
  int loff[ROUNDS], roff[ROUNDS];   // staging offsets (floats) of this lane's tiles

  static __device__ __forceinline__ void put_left(float *stg, int lane, int which, int m, float v) {
    stg[LEFT_OFF + which * LSTR + m * kSLD + lane] = v;
  }
  // columns col0 .. col0+N-1 of the right matrix (col0 even), two per 64-bit store
  template <int N, int NMAX>
  static __device__ __forceinline__ void put_right(float *stg, int lane, int col0, const float (&v)[NMAX]) {
#pragma unroll
    for (int c = 0; c < N; c += 2) {
      const float2 pr = make_float2(v[c], (c + 1 < N) ? v[c + 1 < N ? c + 1 : 0] : 0.f);
      *reinterpret_cast<float2 *>(stg + ((col0 + c) / 2) * kPS + 2 * lane) = pr;
    }
  }

  __device__ __forceinline__ void init(int lane) {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
#pragma unroll
      for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < TC2; ++j) acc[r][i][j] = 0ull;
      int tile = lane + 32 * r;
      if (tile >= NTILES) tile = 0;   // idle slot: recompute tile 0, never written out
      const int rg = tile / CG, cg = tile - rg * CG;
      int which;
      if (cg < CGk) which = L_AB;
      else if (cg < CGk + CGd) which = L_K;
      else if (cg < CGk + 2 * CGd) which = L_ASQ;
      else which = L_W;
      loff[r] = LEFT_OFF + which * LSTR + rg * TR * kSLD;
      roff[r] = (cg * TC / 2) * kPS;
    }
  }

  // With CG dividing 32 a lane's tiles of all rounds (tile = lane + 32 r) lie in the same column group, so the
  // right-operand loads are shared by the rounds (M=20 message GP: 14 -> 12 LDS.128 per 4 particles).
  static constexpr bool kSharedRight = ROUNDS > 1 && (32 % CG == 0);

  // stg holds this step's vectors of the warp's 32 particles (all lanes have written + __syncwarp).
  __device__ __forceinline__ void accumulate(const float *__restrict__ stg) {
    if constexpr (kSharedRight) {
      const float *rp = stg + roff[0];
#pragma unroll(kAccUnroll)
      for (int n4 = 0; n4 < 32; n4 += 4) {
        float4 r0[TC2], r1[TC2];
#pragma unroll
        for (int j = 0; j < TC2; ++j) {
          r0[j] = *reinterpret_cast<const float4 *>(rp + j * kPS + 2 * n4);
          r1[j] = *reinterpret_cast<const float4 *>(rp + j * kPS + 2 * n4 + 4);
        }
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
          const float *lp = stg + loff[r];
          float4 l[TR];
#pragma unroll
          for (int i = 0; i < TR; ++i) l[i] = *reinterpret_cast<const float4 *>(lp + i * kSLD + n4);
#pragma unroll
          for (int i = 0; i < TR; ++i)
#pragma unroll
            for (int j = 0; j < TC2; ++j) {
              unsigned long long c = acc[r][i][j];
              c = ffma2_bcast(r0[j].x, r0[j].y, l[i].x, c);
              c = ffma2_bcast(r0[j].z, r0[j].w, l[i].y, c);
              c = ffma2_bcast(r1[j].x, r1[j].y, l[i].z, c);
              c = ffma2_bcast(r1[j].z, r1[j].w, l[i].w, c);
              acc[r][i][j] = c;
            }
        }
      }
      return;
    }
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const float *lp = stg + loff[r], *rp = stg + roff[r];
#pragma unroll(kAccUnroll)
      for (int n4 = 0; n4 < 32; n4 += 4) {
        float4 l[TR], r0[TC2], r1[TC2];
#pragma unroll
        for (int i = 0; i < TR; ++i) l[i] = *reinterpret_cast<const float4 *>(lp + i * kSLD + n4);
#pragma unroll
        for (int j = 0; j < TC2; ++j) {
          r0[j] = *reinterpret_cast<const float4 *>(rp + j * kPS + 2 * n4);
          r1[j] = *reinterpret_cast<const float4 *>(rp + j * kPS + 2 * n4 + 4);
        }
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
          for (int j = 0; j < TC2; ++j) {
            unsigned long long c = acc[r][i][j];
            c = ffma2_bcast(r0[j].x, r0[j].y, l[i].x, c);
            c = ffma2_bcast(r0[j].z, r0[j].w, l[i].y, c);
            c = ffma2_bcast(r1[j].x, r1[j].y, l[i].z, c);
            c = ffma2_bcast(r1[j].z, r1[j].w, l[i].w, c);
            acc[r][i][j] = c;
          }
      }
    }
  }

  __device__ __forceinline__ void store(float *__restrict__ out, int lane) const {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const int tile = lane + 32 * r;
      if (tile < NTILES) {
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
          for (int j = 0; j < TC2; ++j) {
            float x, y;
            unpack2(acc[r][i][j], x, y);
            out[((size_t)tile * TR + i) * TC + 2 * j] = x;
            out[((size_t)tile * TR + i) * TC + 2 * j + 1] = y;
          }
      }
    }
  }
};

// Reverse of one GP evaluation (SURVEY 8a note 4) for the thread's particle; stages the
// outer-product operands into the warp tile.  k, a from gp_forward_fast.
template <int M, int DIN, int DOUT, int NEED, int SLOT>
__device__ __forceinline__ void gp_reverse_fast(const GpF<M, DIN, DOUT, SLOT> &g, float *__restrict__ stg, int lane,
                                                const float (&xt)[GpF<M, DIN, DOUT, SLOT>::DINP],
                                                const float (&k)[GpF<M, DIN, DOUT, SLOT>::MP],
                                                const float (&a)[GpF<M, DIN, DOUT, SLOT>::MP],
                                                const float (&gm)[DOUT], const float (&gv)[DOUT], bool live,
                                                float (&xinb)[NEED], float (&Lacc)[DIN], float &sw, float &sG) {
  using G = GpF<M, DIN, DOUT, SLOT>;
  using W = WarpAcc<M, DIN, DOUT>;
  constexpr int MP = G::MP, DINP = G::DINP, DOUTP = G::DOUTP;
  if constexpr (kConstOps) {
    using C = typename G::C;
    float Gs = 0.f;
#pragma unroll
    for (int d = 0; d < DOUT; ++d) Gs += gv[d];
    sG += Gs;
    float b[MP];
#pragma unroll
    for (int m = 0; m < MP; ++m) {
      if (m < M) {
        float c = 0.f;
#pragma unroll
        for (int d = 0; d < DOUT; ++d) c = fmaf(C::S(m < M ? m : 0, d), gv[d], c);
        b[m] = a[m] * c;
      } else {
        b[m] = 0.f;
      }
    }
    // stage the operands that are final now, so that b dies at the end of the contraction
#pragma unroll
    for (int m = 0; m < M; ++m) {
      W::put_left(stg, lane, W::L_K, m, k[m]);
      W::put_left(stg, lane, W::L_AB, m, 2.f * b[m] - Gs * k[m]);
      W::put_left(stg, lane, W::L_ASQ, m, a[m] * a[m]);
    }
    W::template put_right<M, MP>(stg, lane, W::colK, k);
    float pb[MP];   // P b - G a, so that a dies here
#pragma unroll
    for (int m = 0; m < MP; ++m) pb[m] = -Gs * a[m];
    matvec_const<SLOT, M, DIN, DOUT, MP, true>(b, pb);
#ifdef CBF_PACKED_DELTA
    constexpr int J2 = C::DINE / 2, N2 = (NEED + 1) / 2;
    unsigned long long x2[J2], xs2[N2], L2[J2];   // packed over input-dim pairs (2j, 2j+1)
#pragma unroll
    for (int j = 0; j < J2; ++j) { x2[j] = pack2(xt[2 * j], xt[2 * j + 1]); L2[j] = 0ull; }
#pragma unroll
    for (int j = 0; j < N2; ++j) xs2[j] = 0ull;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      float kb = 2.f * pb[m];
#pragma unroll
      for (int d = 0; d < DOUT; ++d) kb = fmaf(C::al(m, d), gm[d], kb);
      const float w = kb * k[m];
      sw += w;
      const unsigned long long ww = pack2(w, w);
#pragma unroll
      for (int j = 0; j < J2; ++j) {
        const unsigned long long dl = add2(x2[j], C::nZ2(m, j));
        const unsigned long long wd = mul2(dl, ww);
        if (j < N2) xs2[j < N2 ? j : 0] = add2(xs2[j < N2 ? j : 0], wd);
        L2[j] = fma2(wd, dl, L2[j]);
      }
      W::put_left(stg, lane, W::L_W, m, w);
    }
#pragma unroll
    for (int j = 0; j < J2; ++j) {
      float l0, l1;
      unpack2(L2[j], l0, l1);
      Lacc[2 * j] += l0;
      if (2 * j + 1 < DIN) Lacc[2 * j + 1 < DIN ? 2 * j + 1 : 0] += l1;
    }
#pragma unroll
    for (int j = 0; j < N2; ++j) {
      float s0, s1;
      unpack2(xs2[j], s0, s1);
      xinb[2 * j] = -s0;
      if (2 * j + 1 < NEED) xinb[2 * j + 1 < NEED ? 2 * j + 1 : 0] = -s1;
    }
#else
#pragma unroll
    for (int j = 0; j < NEED; ++j) xinb[j] = 0.f;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      float kb = 2.f * pb[m];
#pragma unroll
      for (int d = 0; d < DOUT; ++d) kb = fmaf(C::al(m, d), gm[d], kb);
      const float w = kb * k[m];
      sw += w;
#pragma unroll
      for (int j = 0; j < DIN; ++j) {
        const float dl = xt[j] + C::nZ(m, j);
        const float wd = w * dl;
        if (j < NEED) xinb[j < NEED ? j : 0] -= wd;
        Lacc[j] = fmaf(wd, dl, Lacc[j]);
      }
      W::put_left(stg, lane, W::L_W, m, w);
    }
#endif
    W::template put_right<DOUT, DOUT>(stg, lane, W::colGm, gm);
    W::template put_right<DOUT, DOUT>(stg, lane, W::colGv, gv);
    {
      float xr[DIN + 1];
#pragma unroll
      for (int j = 0; j < DIN; ++j) xr[j] = live ? xt[j] : 0.f;
      xr[DIN] = live ? 1.f : 0.f;
      W::template put_right<DIN + 1, DIN + 1>(stg, lane, W::colX, xr);
    }
#pragma unroll
    for (int j = 0; j < NEED; ++j) xinb[j] *= C::il(j);
    return;
  }
  float Gs = 0.f;
#pragma unroll
  for (int d = 0; d < DOUT; ++d) Gs += gv[d];
  sG += Gs;
  float b[MP];
#pragma unroll
  for (int m = 0; m < MP; ++m) {
    if (m < M) {
      float S[DOUTP];
      ld_row<DOUTP>(g.Sm + m * DOUTP, S);
      float c = 0.f;
#pragma unroll
      for (int d = 0; d < DOUT; ++d) c = fmaf(S[d], gv[d], c);
      b[m] = a[m] * c;
    } else {
      b[m] = 0.f;
    }
  }
  float pb[MP];
  compiler_fence();
  matvec_fast<MP>(g.P, b, pb);
  compiler_fence();
#pragma unroll
  for (int j = 0; j < NEED; ++j) xinb[j] = 0.f;
#pragma unroll
  for (int m = 0; m < M; ++m) {
    float al[DOUTP];
    ld_row<DOUTP>(g.al + m * DOUTP, al);
    float kb = 2.f * pb[m] - 2.f * Gs * a[m];
#pragma unroll
    for (int d = 0; d < DOUT; ++d) kb = fmaf(al[d], gm[d], kb);
    const float w = kb * k[m];
    sw += w;
    float z[DINP];
    ld_row<DINP>(g.Zt + m * DINP, z);
#pragma unroll
    for (int j = 0; j < DIN; ++j) {
      const float dl = xt[j] - z[j];
      const float wd = w * dl;
      if (j < NEED) xinb[j < NEED ? j : 0] -= wd;
      Lacc[j] = fmaf(wd, dl, Lacc[j]);
    }
    W::put_left(stg, lane, W::L_K, m, k[m]);
    W::put_left(stg, lane, W::L_AB, m, 2.f * b[m] - Gs * k[m]);
    W::put_left(stg, lane, W::L_ASQ, m, a[m] * a[m]);
    W::put_left(stg, lane, W::L_W, m, w);
  }
  W::template put_right<M, MP>(stg, lane, W::colK, k);
  W::template put_right<DOUT, DOUT>(stg, lane, W::colGm, gm);
  W::template put_right<DOUT, DOUT>(stg, lane, W::colGv, gv);
  {
    float xr[DIN + 1];
#pragma unroll
    for (int j = 0; j < DIN; ++j) xr[j] = live ? xt[j] : 0.f;
    xr[DIN] = live ? 1.f : 0.f;
    W::template put_right<DIN + 1, DIN + 1>(stg, lane, W::colX, xr);
  }
  {
    float il[DINP];
    ld_row<DINP>(g.il, il);
#pragma unroll
    for (int j = 0; j < NEED; ++j) xinb[j] *= il[j];
  }
}

// Sum `count` per-thread floats over the CTA; thread 0 writes out[0..count).
template <int COUNT>
__device__ __forceinline__ void cta_sum_store(const float (&vals)[COUNT], float *scratch, float *out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < COUNT; ++i) {
    float v = vals[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) scratch[i * kFastWarps + warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < COUNT; ++i) {
      float s = 0.f;
      for (int w = 0; w < kFastWarps; ++w) s += scratch[i * kFastWarps + w];
      out[i] = s;
    }
  }
}

// =====================================================================================
template <int DX, int DU, int DY, int M, bool SAVE>
__global__ void __launch_bounds__(kFastThreads) bm_forward_fast_kernel(
    Dims D, ChainTable chains, GpDev gp, const float *__restrict__ vxg, const float *__restrict__ u,
    const float *__restrict__ y, const float *__restrict__ eps_b, const float *__restrict__ z_b, Workspace ws,
    float *__restrict__ part_out) {
  constexpr int DH = DX - DY, DIN = DX + DU;
  using G = GpF<M, DIN, DH, 1>;
  extern __shared__ __align__(16) float smem[];
  G g;
  float *p = g.init(smem, gp);
  float *scratch = p; p += 4 * kFastWarps;
  float *vx = p;
  if (threadIdx.x < DX) vx[threadIdx.x] = vxg[threadIdx.x];
  __syncthreads();

  const Chain ch = chains.c[blockIdx.y];
  const int nl = blockIdx.x * kFastThreads + threadIdx.x;
  const bool live = nl < D.n_local;
  const int nr = live ? nl : 0;
  const int b = (D.n_offset + nr) / D.S;
  const float *ub = u + (size_t)b * D.T * DU;
  const float *yb = y + (size_t)b * D.T * DY;
  const size_t np = ws.npad;

  float h[DH], ent = 0.f;
  {
    const float z = (ch.init == 1) ? z_b[((size_t)ch.run * D.T + ch.t_hi) * D.n_local + nr] : 0.f;
#pragma unroll
    for (int j = 0; j < DH; ++j) h[j] = z;
  }
#pragma unroll 1
  for (int t = ch.t_hi; t >= ch.t_lo; --t) {
    compiler_fence();
    float xin[DIN], xt[G::DINP], k[G::MP], a[G::MP], fm[DH], fv[DH];
#pragma unroll
    for (int j = 0; j < DH; ++j) xin[j] = h[j];
#pragma unroll
    for (int j = 0; j < DU; ++j) xin[DH + j] = ub[t * DU + j];
#pragma unroll
    for (int j = 0; j < DY; ++j) xin[DH + DU + j] = yb[t * DY + j];
    const float e = eps_b[((size_t)ch.run * D.T + t) * D.n_local + nr];
    gp_forward_fast<M, DIN, DH, 1>(g, xin, xt, k, a, fm, fv);
    if constexpr (SAVE) {
      if (live) SavedEval<G::MP, DH>::store(ws.KAb, (size_t)ch.run * D.T + t, np, nl, k, a, fm, fv);
    }
    const bool write = writer_run(t, D.R) == ch.run;
#pragma unroll
    for (int j = 0; j < DH; ++j) {
      const float f = fv[j] + vx[j];
      h[j] = fm[j] + h[j] + e * sqrtf(f);
      if (write) ent += 0.5f * (kLog2PiE + logf(f));
    }
    if (live) {
      float *Hp = ws.H + (((size_t)ch.run * D.T + t) * DH) * np + nl;
#pragma unroll
      for (int j = 0; j < DH; ++j) Hp[j * np] = h[j];
    }
  }
  const float v[1] = {live ? ent : 0.f};
  cta_sum_store<1>(v, scratch, part_out + ((size_t)blockIdx.y * gridDim.x + blockIdx.x));
}

// =====================================================================================
template <int DX, int DU, int DY, int M, bool SAVE>
__global__ void __launch_bounds__(kFastThreads, (SAVE && DX <= 4) ? 4 : 1) fw_forward_fast_kernel(
    Dims D, GpDev gp, const float *__restrict__ vxg, const float *__restrict__ vyg, const float *__restrict__ u,
    const float *__restrict__ y, const float *__restrict__ eps_f, Workspace ws, float *__restrict__ part_out) {
  constexpr int DH = DX - DY, DIN = DX + DU;
  using G = GpF<M, DIN, DX, 0>;
  extern __shared__ __align__(16) float smem[];
  G g;
  float *p = g.init(smem, gp);
  float *scratch = p; p += (DY + 1) * kFastWarps;
  float *vx = p; p += 4 * ((DX + 3) / 4);
  float *vy = p;
  if (threadIdx.x < DX) { vx[threadIdx.x] = vxg[threadIdx.x]; vy[threadIdx.x] = vyg[threadIdx.x]; }
  __syncthreads();

  const int nl = blockIdx.x * kFastThreads + threadIdx.x;
  const bool live = nl < D.n_local;
  const int nr = live ? nl : 0;
  const int b = (D.n_offset + nr) / D.S;
  const float *ub = u + (size_t)b * D.T * DU;
  const float *yb = y + (size_t)b * D.T * DY;
  const size_t np = ws.npad;

  auto load_ytil = [&](int t, float(&yt)[DX]) {
#pragma unroll
    for (int j = 0; j < DY; ++j) yt[j] = yb[t * DY + j];
    const float *Hp = ws.H + (((size_t)writer_run(t, D.R) * D.T + t) * DH) * np + nr;
#pragma unroll
    for (int j = 0; j < DH; ++j) yt[DY + j] = D.half ? 0.f : Hp[j * np];
  };

  float x[DX], sse[DY + 1], kl = 0.f;
#pragma unroll
  for (int j = 0; j <= DY; ++j) sse[j] = 0.f;
  load_ytil(0, x);
  if (D.half) {   // x_0 from the recognition model (cbfssmhalf.py:103)
#pragma unroll
    for (int j = 0; j < DX; ++j) x[j] = ws.x0[(size_t)b * DX + j];
  }
#pragma unroll 1
  for (int t = 0; t < D.T; ++t) {
    compiler_fence();
    if (live) {
      float *Xp = ws.X + ((size_t)t * DX) * np + nl;
#pragma unroll
      for (int j = 0; j < DX; ++j) Xp[j * np] = x[j];
#pragma unroll
      for (int j = 0; j < DY; ++j) { const float d = yb[t * DY + j] - x[j]; sse[j] = fmaf(d, d, sse[j]); }
    }
    if (t == D.T - 1) break;
    float xin[DIN], xt[G::DINP], k[G::MP], a[G::MP], fm[DX], fv[DX], yt[DX], xn[DX];
#pragma unroll
    for (int j = 0; j < DX; ++j) xin[j] = x[j];
#pragma unroll
    for (int j = 0; j < DU; ++j) xin[DX + j] = ub[t * DU + j];
    load_ytil(t + 1, yt);
    const float e = eps_f[(size_t)t * D.n_local + nr];
    gp_forward_fast<M, DIN, DX, 0>(g, xin, xt, k, a, fm, fv);
    if constexpr (SAVE) {
      if (live) SavedEval<G::MP, DX>::store(ws.KAf, (size_t)t, np, nl, k, a, fm, fv);
    }
    const bool do_cond = D.condition || (t < D.R - 1);
    fw_step<DX>(x, fm, fv, yt, e, vx, vy, D.kap, do_cond, D.ncond, xn, kl);
#pragma unroll
    for (int j = 0; j < DX; ++j) x[j] = xn[j];
  }
  sse[DY] = kl;
  if (!live) {
#pragma unroll
    for (int j = 0; j <= DY; ++j) sse[j] = 0.f;
  }
  cta_sum_store<DY + 1>(sse, scratch, part_out + (size_t)blockIdx.x * (DY + 1));
}

// =====================================================================================
// Reverse kernels: persistent CTAs; every WARP writes its own partial slot
// [tile accumulators | L_j, sum w, sum G, var_x_bar, var_y_bar].
// =====================================================================================
template <int DX, int DU, int DY, int M, bool SAVED>
__global__ void __launch_bounds__(kFastThreads, rev_minblocks<DX>()) fw_reverse_fast_kernel(
    Dims D, GpDev gp, const float *__restrict__ vxg, const float *__restrict__ vyg, const float *__restrict__ u,
    const float *__restrict__ y, const float *__restrict__ eps_f, float w_ll, float w_kl, Workspace ws,
    float *__restrict__ part_out, int slot) {
  constexpr int DH = DX - DY, DIN = DX + DU;
  using G = GpF<M, DIN, DX, 0>;
  using W = WarpAcc<M, DIN, DX>;
  extern __shared__ __align__(16) float smem[];
  G g;
  float *p = g.init(smem, gp);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *stg = p + warp * W::FLOATS; p += kFastWarps * W::FLOATS;
  float *vx = p; p += 4 * ((DX + 3) / 4);
  float *vy = p;
  if (threadIdx.x < DX) { vx[threadIdx.x] = vxg[threadIdx.x]; vy[threadIdx.x] = vyg[threadIdx.x]; }
  for (int i = lane; i < W::FLOATS; i += 32) stg[i] = 0.f;
  __syncthreads();

  W wacc;
  wacc.init(lane);
  float Lacc[DIN], sw = 0.f, sG = 0.f, vxacc[DX], vyacc[DX];
#pragma unroll
  for (int j = 0; j < DIN; ++j) Lacc[j] = 0.f;
#pragma unroll
  for (int j = 0; j < DX; ++j) { vxacc[j] = 0.f; vyacc[j] = 0.f; }

  const int ntile = ceil_div(D.n_local, kFastThreads);
  const size_t np = ws.npad;
  for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int nl = tile * kFastThreads + threadIdx.x;
    const bool live = nl < D.n_local;
    const int nr = live ? nl : 0;
    const int b = (D.n_offset + nr) / D.S;
    const float *ub = u + (size_t)b * D.T * DU;
    const float *yb = y + (size_t)b * D.T * DY;

    float xb[DX];
    {
      const float *Xp = ws.X + ((size_t)(D.T - 1) * DX) * np + nr;
#pragma unroll
      for (int j = 0; j < DX; ++j)
        xb[j] = (j < DY && live) ? w_ll * (yb[(D.T - 1) * DY + (j < DY ? j : 0)] - Xp[j * np]) / vy[j] : 0.f;
    }
    // Particle-major operands of a step (x_t, y2_{t+1}, eps_t) are fetched during the previous step's
    // accumulation phase, when few registers are live, so their HBM/L2 latency is off the critical path.
    using SE = SavedEval<G::MP, DX>;
    float xq[DX], hq[DH], eq;
    float4 fq[SE::NF];   // SAVED: this step's (fmean, fvar), fetched a step ahead like the other operands
    auto fetch = [&](int t) {
      const float *Xp = ws.X + ((size_t)t * DX) * np + nr;
#pragma unroll
      for (int j = 0; j < DX; ++j) xq[j] = Xp[j * np];
      const float *Hp = ws.H + (((size_t)writer_run(t + 1, D.R) * D.T + (t + 1)) * DH) * np + nr;
#pragma unroll
      for (int j = 0; j < DH; ++j) hq[j] = D.half ? 0.f : Hp[j * np];
      eq = eps_f[(size_t)t * D.n_local + nr];
      if constexpr (SAVED) SE::load_f(ws.KAf, (size_t)t, np, nr, fq);
    };
    if (D.T >= 2) fetch(D.T - 2);
#pragma unroll 1
    for (int t = D.T - 2; t >= 0; --t) {
    compiler_fence();
    // The loop body (~68 KB of SASS) exceeds the 32 KB L1.5 instruction cache; keeping the CTA's warps
    // within one step of each other lets them share fetched lines (measured -7.6 % kernel time).
    __syncthreads();
      float x[DX], xin[DIN], xt[G::DINP], k[G::MP], a[G::MP], fm[DX], fv[DX], yt[DX];
#pragma unroll
      for (int j = 0; j < DX; ++j) { x[j] = xq[j]; xin[j] = x[j]; }
#pragma unroll
      for (int j = 0; j < DU; ++j) xin[DX + j] = ub[t * DU + j];
#pragma unroll
      for (int j = 0; j < DY; ++j) yt[j] = yb[(t + 1) * DY + j];
#pragma unroll
      for (int j = 0; j < DH; ++j) yt[DY + j] = hq[j];
      const float e = eq;
      if constexpr (SAVED) {      // the evaluation the forward kernel left behind (k, a first: they are used last)
        SE::load_ka(ws.KAf, (size_t)t, np, nr, k, a);
        SE::unpack_f(fq, fm, fv);
        gp_scale_input<M, DIN, DX, 0>(g, xin, xt);
      } else {
        gp_forward_fast<M, DIN, DX, 0>(g, xin, xt, k, a, fm, fv);
      }
      const bool do_cond = D.condition || (t < D.R - 1);
      float fmb[DX], fvb[DX], ytb[DX];
      fw_step_adjoint<DX>(x, fm, fv, yt, e, vx, vy, D.kap, do_cond, D.ncond, w_kl, xb, fmb, fvb, ytb, vxacc, vyacc, live);
      if (live) {
        float *Yp = ws.Yb + ((size_t)(t + 1) * DH) * np + nl;
#pragma unroll
        for (int j = 0; j < DH; ++j) Yp[j * np] = ytb[DY + j];
      } else {
#pragma unroll
        for (int j = 0; j < DX; ++j) { fmb[j] = 0.f; fvb[j] = 0.f; }
      }
      float xinb[DX];
      __syncwarp();   // previous step's accumulate() has finished reading the staging tile
      gp_reverse_fast<M, DIN, DX, DX, 0>(g, stg, lane, xt, k, a, fmb, fvb, live, xinb, Lacc, sw, sG);
      __syncwarp();
      if (t > 0) fetch(t - 1);
      wacc.accumulate(stg);
#pragma unroll
      for (int j = 0; j < DX; ++j) {
        float lg = 0.f;
        if (j < DY && live) lg = w_ll * (yb[t * DY + (j < DY ? j : 0)] - x[j]) / vy[j];
        xb[j] = xinb[j] + fmb[j] + lg;
      }
    }
    if (live && D.half) {   // CBFSSMHALF: adjoint of the recognition model's x_0 (all dims)
#pragma unroll
      for (int j = 0; j < DX; ++j) ws.x0b[j * np + nl] = xb[j];
    } else if (live) {
      float *Yp = ws.Yb + nl;
#pragma unroll
      for (int j = 0; j < DH; ++j) Yp[j * np] = xb[DY + j];
    }
  }
  // ---- per-warp partial ----
  float *out = part_out + ((size_t)blockIdx.x * kFastWarps + warp) * slot;
  wacc.store(out, lane);
  float sc[DIN + 2 + 2 * DX];
#pragma unroll
  for (int j = 0; j < DIN; ++j) sc[j] = Lacc[j];
  sc[DIN] = sw; sc[DIN + 1] = sG;
#pragma unroll
  for (int j = 0; j < DX; ++j) { sc[DIN + 2 + j] = vxacc[j]; sc[DIN + 2 + DX + j] = vyacc[j]; }
  float *so = out + round_up(W::NACC, 4);
#pragma unroll
  for (int i = 0; i < DIN + 2 + 2 * DX; ++i) {
    float v = sc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) so[i] = v;
  }
}

template <int DX, int DU, int DY, int M, bool SAVED>
__global__ void __launch_bounds__(kFastThreads, rev_minblocks<DX>()) bm_reverse_fast_kernel(
    Dims D, ChainTable chains, GpDev gp, const float *__restrict__ vxg, const float *__restrict__ u,
    const float *__restrict__ y, const float *__restrict__ eps_b, const float *__restrict__ z_b, float w_en,
    Workspace ws, float *__restrict__ part_out, int slot) {
  constexpr int DH = DX - DY, DIN = DX + DU;
  using G = GpF<M, DIN, DH, 1>;
  using W = WarpAcc<M, DIN, DH>;
  extern __shared__ __align__(16) float smem[];
  G g;
  float *p = g.init(smem, gp);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *stg = p + warp * W::FLOATS; p += kFastWarps * W::FLOATS;
  float *vx = p;
  if (threadIdx.x < DX) vx[threadIdx.x] = vxg[threadIdx.x];
  for (int i = lane; i < W::FLOATS; i += 32) stg[i] = 0.f;
  __syncthreads();

  W wacc;
  wacc.init(lane);
  float Lacc[DIN], sw = 0.f, sG = 0.f, vxacc[DH];
#pragma unroll
  for (int j = 0; j < DIN; ++j) Lacc[j] = 0.f;
#pragma unroll
  for (int j = 0; j < DH; ++j) vxacc[j] = 0.f;

  const int ntile = ceil_div(D.n_local, kFastThreads);
  const int nitem = ntile * chains.count;
  const size_t np = ws.npad;
  for (int item = blockIdx.x; item < nitem; item += gridDim.x) {
    const int tile = item % ntile;
    const Chain ch = chains.c[item / ntile];
    const int nl = tile * kFastThreads + threadIdx.x;
    const bool live = nl < D.n_local;
    const int nr = live ? nl : 0;
    const int b = (D.n_offset + nr) / D.S;
    const float *ub = u + (size_t)b * D.T * DU;
    const float *yb = y + (size_t)b * D.T * DY;

    float hb[DH];
#pragma unroll
    for (int j = 0; j < DH; ++j) hb[j] = 0.f;
    // Particle-major operands of a step (message state, eps, adjoint of y2) are fetched during the previous
    // step's accumulation phase, when few registers are live.
    using SE = SavedEval<G::MP, DH>;
    float hq[DH], yq[DH], eq;
    float4 fq[SE::NF];
    auto fetch = [&](int t) {
      if constexpr (SAVED) SE::load_f(ws.KAb, (size_t)ch.run * D.T + t, np, nr, fq);
      if (t == ch.t_hi) {
        const float z = (ch.init == 1) ? z_b[((size_t)ch.run * D.T + t) * D.n_local + nr] : 0.f;
#pragma unroll
        for (int j = 0; j < DH; ++j) hq[j] = z;
      } else {
        const float *Hp = ws.H + (((size_t)ch.run * D.T + (t + 1)) * DH) * np + nr;
#pragma unroll
        for (int j = 0; j < DH; ++j) hq[j] = Hp[j * np];
      }
      eq = eps_b[((size_t)ch.run * D.T + t) * D.n_local + nr];
      const float *Yp = ws.Yb + ((size_t)t * DH) * np + nr;
#pragma unroll
      for (int j = 0; j < DH; ++j) yq[j] = Yp[j * np];
    };
    fetch(ch.t_lo);
#pragma unroll 1
    for (int t = ch.t_lo; t <= ch.t_hi; ++t) {
    compiler_fence();
#ifdef CBF_STEP_SYNC
    __syncthreads();
#endif
      float hid[DH], xin[DIN], xt[G::DINP], k[G::MP], a[G::MP], fm[DH], fv[DH];
#pragma unroll
      for (int j = 0; j < DH; ++j) { hid[j] = hq[j]; xin[j] = hid[j]; }
#pragma unroll
      for (int j = 0; j < DU; ++j) xin[DH + j] = ub[t * DU + j];
#pragma unroll
      for (int j = 0; j < DY; ++j) xin[DH + DU + j] = yb[t * DY + j];
      const float e = eq;
      if constexpr (SAVED) {
        SE::load_ka(ws.KAb, (size_t)ch.run * D.T + t, np, nr, k, a);
        SE::unpack_f(fq, fm, fv);
        gp_scale_input<M, DIN, DH, 1>(g, xin, xt);
      } else {
        gp_forward_fast<M, DIN, DH, 1>(g, xin, xt, k, a, fm, fv);
      }
      const bool write = writer_run(t, D.R) == ch.run;
      float ob[DH], fvb[DH];
#pragma unroll
      for (int j = 0; j < DH; ++j) {
        const float f = fv[j] + vx[j];
        float o = hb[j], fb = 0.f;
        if (write) {
          o += yq[j];
          fb = w_en * 0.5f / f;
        }
        fb += o * e * 0.5f * rsqrtf(f);
        if (!live) { o = 0.f; fb = 0.f; }
        ob[j] = o; fvb[j] = fb;
        vxacc[j] += fb;
      }
      float xinb[DH];
      __syncwarp();
      gp_reverse_fast<M, DIN, DH, DH, 1>(g, stg, lane, xt, k, a, ob, fvb, live, xinb, Lacc, sw, sG);
      __syncwarp();
      if (t < ch.t_hi) fetch(t + 1);
      wacc.accumulate(stg);
#pragma unroll
      for (int j = 0; j < DH; ++j) hb[j] = xinb[j] + ob[j];
    }
  }
  float *out = part_out + ((size_t)blockIdx.x * kFastWarps + warp) * slot;
  wacc.store(out, lane);
  float sc[DIN + 2 + 2 * DX];
#pragma unroll
  for (int j = 0; j < DIN; ++j) sc[j] = Lacc[j];
  sc[DIN] = sw; sc[DIN + 1] = sG;
#pragma unroll
  for (int j = 0; j < DX; ++j) { sc[DIN + 2 + j] = (j < DH) ? vxacc[j < DH ? j : 0] : 0.f; sc[DIN + 2 + DX + j] = 0.f; }
  float *so = out + round_up(W::NACC, 4);
#pragma unroll
  for (int i = 0; i < DIN + 2 + 2 * DX; ++i) {
    float v = sc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) so[i] = v;
  }
}

}  // namespace cbf
