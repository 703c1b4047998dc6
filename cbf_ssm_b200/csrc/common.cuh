// Shared declarations of the cbfssm_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/cbfssm_b200.h"

namespace cbf {

// Counts the kernel launches of this library on the calling host thread (cbf_launches_read; bench.py's
// gpu_launches).  Defined in api.cu.
void cbf_note_launch();

constexpr int kNP = 32;          // particles per CTA tile in the cooperative (generic) path
constexpr int kLD = kNP + 4;     // row stride of the [m][n] shared arrays (== 4 mod 32: conflict-free float4 rows)
constexpr int kMaxChains = 120;  // chain segments per launch (kernel parameter space)
constexpr float kLog2PiE = 2.8378770664093453f;   // log(2*pi*e)
constexpr float kNegHalfLog2e = -0.72134752044448170f;  // -0.5 * log2(e)

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// One live segment of a backward-message run (cbfssm.py:123-136): steps t = t_hi .. t_lo.
struct Chain {
  int run;
  int t_hi;
  int t_lo;
  int init;  // 0: h = 0 (cbfssm.py:106)   1: h = tile(z_b[run, t_top]) (cbfssm.py:133-135)
  int col0;  // number of live steps of all earlier chains (column block of the tensor-path operand matrices)
  // Tensor-path reverse pass in time windows: a chain may be cut into pieces [t_lo, t_hi] processed in
  // successive launches; t_top is the whole chain's first step (where `init` applies), `carry` says whether
  // the piece loads (bit 0) / stores (bit 1) the message adjoint in slot `id` of the carry buffer.
  int t_top, carry, id;
};
struct ChainTable {
  int count;
  Chain c[kMaxChains];
};

// Time window of the tensor-path reverse of the forward rollout: steps t_hi .. t_lo (descending) in one
// launch; `first` = the window starts at T-2 (adjoint initialised from the likelihood), `last` = it ends at
// t = 0; otherwise the state adjoint enters / leaves through Workspace::carry_f.
struct TimeWin {
  int t_hi, t_lo, first, last;
};

struct GpDev {
  const float *Z, *ell, *sig2, *P, *alpha, *S;
};

// Outer-product operands one tensor-path reverse kernel writes for one GP, already in the form the
// accumulation GEMM (kernels_outer.cuh) consumes: every value x is stored as two bfloat16 terms
// hi = rn(x), lo = rn(x - hi) (16 significant bits together), and the L = (live steps x particles)
// columns are cut into tiles of kOT.  A tile is a sequence of row-blocks; one row-block holds 8 rows x kOT
// columns as kOT 16-byte segments (segment c = the 8 rows' values of column c): the MN-major,
// no-swizzle core-matrix layout of tcgen05 (8 columns x 16 bytes = one 128-byte core matrix), so a tile
// is copied to shared memory by plain bulk copies and a writing thread (one column) stores 16 bytes per
// 8 rows.  Parts, hi copies first: a_bar | k' | a^2 | w (MB blocks each) | g_mean | g_var (DB) | [x~,1] (XB);
// the lo copy of a part sits RB blocks after its hi copy.  Rows beyond a part's size are written as zeros.
constexpr int kOT = 32;                 // columns per tile (two k-steps of 16)
constexpr int kOBlk = kOT * 16;         // bytes of one row-block
struct TcMats {
  unsigned char *blk;
  size_t L;        // valid columns
  int MB, DB, XB;  // row-blocks of an M-row part, of g_mean / g_var, of [x~,1]
  int RB;          // row-blocks per half tile = 4 MB + 2 DB + XB
  int bAb, bK, bA2, bW, bGm, bGv, bX1;   // first (hi) row-block of each part
  __host__ __device__ size_t tile_bytes() const { return (size_t)2 * RB * kOBlk; }
};

// Device views into the caller's workspace.
struct Workspace {
  float *X;        // [T][dx][npad]      forward states x_t                      (cbfssm.py:166-169,229)
  float *H;        // [2][T][dh][npad]   backward-message outputs of both runs   (cbfssm.py:150,158)
  float *Yb;       // [T][dh][npad]      adjoint of y2[t]
  float *fpart_bm; // [nchains][ptiles]  entropy partial per CTA
  float *fpart_fw; // [ptiles][dy+1]     sse_j, kl_x partial per CTA
  float *gpart_f;  // [grid][slot_f]     reverse partials, forward-rollout GP
  float *gpart_b;  // [grid][slot_b]     reverse partials, backward-message GP
  double *acc_f;   // [slot_f]           reduced
  double *acc_b;   // [slot_b]
  double *stats;   // sse[dy], kl_x, entropy of this shard (float64)
  float *carry_f;  // [dx][npad]            state adjoint between time windows (tensor path)
  float *carry_b;  // [chains][dh][npad]    message adjoint between chain pieces (tensor path)
  float *cpack;    // [2][2048] packed constant-bank images of the two GPs (register path)
  float *FVf;      // tensor path, optional: saved (fmean[dx], fvar[dx], amax) of every forward-rollout GP evaluation,
  float *FVb;      //   ... of every backward-message evaluation (slot = run * T + t): plane v of slot e at ((e * NV + v) * npad + n)
  float4 *KAf;     // register path, optional: saved (k, a, fmean, fvar) of every forward-rollout GP evaluation
  float4 *KAb;     //                          ... of every backward-message evaluation, slot = run * T + t
  const float *x0; // CBFSSMHALF: x_0 per sequence [B][dx] (output of the recognition model)
  float *x0b;      // CBFSSMHALF: adjoint of x_0 per particle [dx][npad]
  int npad;
};

// Per-problem constants derived from cbf_shape for the kernels.
struct Dims {
  int B, S, T, M, R, condition, n_offset, n_local, npad;
  float kap;
  int half;    // 1: CBFSSMHALF -- x_0 from ws.x0, no backward message, only the first dy dims conditioned
  int ncond;   // conditioned state dims: dx (CBFSSM) or dy (CBFSSMHALF)
};

// The sparse-GP predictive variance sigma^2 - k.P k + sum_m a_m^2 S_md (gp_tf.py:140,159) is >= 0 in exact
// arithmetic (k.(K_zz + jitter)^-1 k <= sigma^2), but the kernels form it in float32 from an explicit inverse, so
// for ill-conditioned K_zz the cancellation error can push it below zero; the reference (float64 Cholesky
// solves) never sees that.  Clamping at the exact lower bound keeps fvar + var_x >= var_x > 0, so the sqrt /
// log / reciprocal that follow stay finite instead of turning the loss and Adam's update into NaN.
__device__ __forceinline__ float gp_var_clamp(float v) { return fmaxf(v, 0.f); }

inline __host__ __device__ int writer_run(int t, int R) { return (t % (2 * R)) < R ? 0 : 1; }

// Layout of one reverse-pass accumulator block for a GP with (M, Din, Dout).
// The parameter adjoints are one outer-product accumulation  ACC[m][col] += left[m] * right[col]
// over (particle, step), with the "right" vector [k (M) | g_mean (Dout) | g_var (Dout) | x~ (Din), 1]
// and the left vector chosen per column block (a_bar | k | a^2 | w | w).  It is tiled TR x TC;
// each column block is padded to a multiple of TC.
//   interleaved = 0 (cooperative path): index = (i * ntiles + tile) * TC + j      (TR = TC = 4)
//   interleaved = 1 (register path)   : index = (tile * TR + i) * TC + j
struct AccLayout {
  int M, Din, Dout, TR, TC, interleaved;
  int RG, CGk, CGd, CGx, CG, ntiles, nacc;
  int colK, colGm, colGv, colX, ncols;   // first padded column of each block
  int nscal;                             // per-CTA scalar sums appended after the tiles
  // legacy names used by the cooperative kernels
  int MP, MG, DG, XG;
  __host__ __device__ AccLayout() {}
  __host__ __device__ AccLayout(int M_, int Din_, int Dout_, int dx, int TR_ = 4, int TC_ = 4, int inter = 0) {
    M = M_; Din = Din_; Dout = Dout_; TR = TR_; TC = TC_; interleaved = inter;
    RG = ceil_div(M, TR);
    CGk = ceil_div(M, TC); CGd = ceil_div(Dout, TC); CGx = ceil_div(Din + 1, TC);
    CG = CGk + 2 * CGd + CGx;
    ntiles = RG * CG; nacc = ntiles * TR * TC;
    colK = 0; colGm = CGk * TC; colGv = colGm + CGd * TC; colX = colGv + CGd * TC; ncols = CG * TC;
    nscal = Din + 2 + 2 * dx;   // L_j[Din], sum w, sum G, var_x_bar[dx], var_y_bar[dx]
    MP = round_up(M, 4); MG = MP / 4; DG = CGd; XG = CGx;
  }
  __host__ __device__ int slot() const { return round_up(nacc, 4) + round_up(nscal, 4); }
  __host__ __device__ int scal_off() const { return round_up(nacc, 4); }
  __host__ __device__ size_t index(int m, int col) const {
    const int rg = m / TR, i = m - rg * TR, cg = col / TC, j = col - cg * TC;
    const int tile = rg * CG + cg;
    return interleaved ? ((size_t)tile * TR + i) * TC + j : ((size_t)i * ntiles + tile) * TC + j;
  }
};

// float64 state of one GP's prologue (cbf_gp_prologue): constrained parameters, K_zz, P = (K_zz + 1e-8 I)^-1,
// alpha = P m, sigmoids for the adjoint, scratch.
struct ProState {   // offsets (doubles) into the caller's state buffer
  int64_t ell, sgl, sig2, sgv, S, sgS, m, Zt, K0, P, alpha, W1, W2, cond, total;
  __host__ __device__ ProState(int M, int Din, int Dout) {
    int64_t o = 0;
    ell = o; o += Din;
    sgl = o; o += Din;
    sig2 = o; o += 1;
    sgv = o; o += 1;
    S = o; o += (int64_t)M * Dout;
    sgS = o; o += (int64_t)M * Dout;
    m = o; o += (int64_t)M * Dout;
    Zt = o; o += (int64_t)M * Din;
    K0 = o; o += (int64_t)M * M;
    P = o; o += (int64_t)M * M;
    alpha = o; o += (int64_t)M * Dout;
    const int64_t w = (int64_t)M * (M > Din ? M : Din);   // scratch, also holds an [M, Din] temporary
    W1 = o; o += w;
    W2 = o; o += w;
    cond = o; o += 1;      // 1-norm condition number of K_zz + 1e-8 I (diagnostic: float32 reaches ~1e3)
    total = o;
  }
};

// Kernel entry points of one (dx,du,dy) instantiation, filled by CBF_INSTANTIATE.
struct DimOps {
  int dx, du, dy;
  cudaError_t (*bm_forward)(const Dims &, const ChainTable &, GpDev, const float *vx, const float *u,
                            const float *y, const float *eps_b, const float *z_b, Workspace, float *part_out,
                            cudaStream_t);
  cudaError_t (*fw_forward)(const Dims &, GpDev, const float *vx, const float *vy, const float *u,
                            const float *y, const float *eps_f, Workspace, float *part_out, cudaStream_t);
  cudaError_t (*fw_reverse)(const Dims &, GpDev, const float *vx, const float *vy, const float *u,
                            const float *y, const float *eps_f, float w_ll, float w_kl, Workspace,
                            float *part_out, int grid, cudaStream_t);
  cudaError_t (*bm_reverse)(const Dims &, const ChainTable &, GpDev, const float *vx, const float *u,
                            const float *y, const float *eps_b, const float *z_b, float w_en, Workspace,
                            float *part_out, int grid, cudaStream_t);
  // optional tensor-core (tcgen05) forward kernels for large M; nullptr when not compiled in
  cudaError_t (*bm_forward_tc)(const Dims &, const ChainTable &, GpDev, const float *vx, const float *u,
                               const float *y, const float *eps_b, const float *z_b, Workspace, float *part_out,
                               cudaStream_t);
  cudaError_t (*fw_forward_tc)(const Dims &, GpDev, const float *vx, const float *vy, const float *u,
                               const float *y, const float *eps_f, Workspace, float *part_out, cudaStream_t);
  cudaError_t (*fw_reverse_tc)(const Dims &, GpDev, const float *vx, const float *vy, const float *u,
                               const float *y, const float *eps_f, float w_ll, float w_kl, Workspace, TcMats,
                               TimeWin, float *spart, int nsc, cudaStream_t);
  cudaError_t (*bm_reverse_tc)(const Dims &, const ChainTable &, GpDev, const float *vx, const float *u,
                               const float *y, const float *eps_b, const float *z_b, float w_en, Workspace,
                               TcMats, float *spart, int nsc, cudaStream_t);
  size_t (*smem_tc)(int M, int which);     // which: 0 bm_fwd 1 fw_fwd 2 fw_rev 3 bm_rev
  size_t (*smem_bytes)(int M, int which);  // which: 0 bm_fwd 1 fw_fwd 2 fw_rev 3 bm_rev
  int (*occupancy)(int M, int which);      // resident CTAs/SM of the persistent reverse kernels
  void (*layouts)(int M, AccLayout *Lf, AccLayout *Lb);   // accumulator layouts of the two reverse kernels
  int (*saved_planes)(int which);          // register path: float4 planes of one saved evaluation (0 forward GP, 1 message GP); else nullptr
  int slots_per_cta;                       // partial-sum slots each reverse CTA writes
  int particles_per_cta;                   // particle tile of one CTA
  int fixed_M;                             // 0: any M (cooperative path); else the compiled-in M
};

// Register-resident ops for (dx,du,dy,M) if compiled in and allowed, else the cooperative ops.
const DimOps *find_ops(int dx, int du, int dy, int M, bool allow_fast);

void set_error(const char *fmt, ...);

}  // namespace cbf
