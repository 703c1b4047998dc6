// Parameter-adjoint accumulation of the tensor path as a tcgen05 / TMEM split-K reduction.
//
// The tensor-path reverse kernels leave, per GP, a tile-major float32 operand matrix (TcMats) over the
// L = (live steps x particles) columns: a_bar, k', a^2, w (M rows each), g_mean, g_var (Dout),
// [x~, 1] (Din+1).  The parameter adjoints are the four thin GEMMs over that huge K dimension
//     P_bar' = Ab K^T [M x M],  alpha_bar' = K Gm^T,  S_bar = A2 Gv^T,  [U | r] = W X1^T .
// They are memory-bound (13.5 FLOP/B is above the FP32-SIMT ridge but far below the tensor one),
// so each CTA streams a contiguous range of 16-column blocks once: 256 threads load the block (one
// contiguous ~26 KB region) with coalesced 16-byte loads, split every value x into hi = x truncated to TF32 and lo = x - hi
// (exact), and store both in the canonical K-major no-swizzle UMMA layout; one thread issues
// 4 blocks x 3 passes (hi.hi + lo.hi + hi.lo, the 3xTF32 scheme: ~2^-21 relative) x 2 k-steps of
// tcgen05.mma.kind::tf32 into fp32 accumulators in TMEM; the next tile's global loads overlap the
// MMAs.  Every kODrain tiles the accumulators are drained with tcgen05.ld and added (round-to-nearest)
// into per-thread float32 registers, because the tensor core's own fp32 accumulation degrades over long
// chains: measured against the SIMT path at L = 5e6, one chain per CTA gave 2-3e-4 relative error in
// P_bar, a drain every 32 tiles (192 accumulations) 5e-6 (tools/tc_accuracy.py).  Per-CTA partials are
// summed in float64 by a last kernel.
#pragma once
#include "kernels_tc.cuh"

namespace cbf {

constexpr int kOT = 16;            // columns per tile = 2 MMA k-steps of 8
constexpr int kOThreads = 256;
constexpr int kODrain = 32;        // tiles between accumulator drains (see below: longer chains lose accuracy)
constexpr int kOStages = 3;        // shared-memory ring depth (3 x ~72 KB)
constexpr int kOCols = 192;        // TMEM columns used: P_bar 0..127 | alpha 128..143 | S 144..159 | U 160..191

struct OuterArgs {
  TcMats m;
  int M, dout, din;
};

__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Rpart: [gridDim.x][128][kOCols] float64, fully written by every CTA.
__global__ void __launch_bounds__(kOThreads) tc_outer_kernel(OuterArgs a, double *__restrict__ Rpart) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar_storage[kOStages];
  __shared__ uint32_t tmem_storage;
  // the 24 MMAs of one tile (4 blocks x 3 passes x 2 k-steps): operand descriptors per stage, TMEM column and
  // instruction descriptor per MMA -- precomputed so that the single issuing thread's loop stays short
  __shared__ uint64_t dtab[kOStages][24][2];
  __shared__ uint32_t dmeta[24][2];
  const int tid = threadIdx.x;
  const int M = a.M, MP = round_up(M, 16);
  const int nx = a.din + 1, NX = round_up(nx, 16);          // rows of the X1 tile (16 or 32)
  // One stage (floats): 8 left tiles [128 x 16] (hi, lo of Ab, K, A2, W), then Gm, Gv (16 rows), X1 (NX rows),
  // hi+lo each.  kOStages stages form a ring: the producers (all threads) refill stage s while the MMAs
  // of the other stages are still running.
  const int nfl = 8 * 128 * kOT + 4 * 16 * kOT + 2 * NX * kOT;
  float *stage0 = reinterpret_cast<float *>(smem_raw);
  for (int i = tid; i < kOStages * nfl; i += kOThreads) stage0[i] = 0.f;   // rows >= M / unused rows stay zero
  uint32_t bar[kOStages];
#pragma unroll
  for (int q = 0; q < kOStages; ++q) bar[q] = smem_u32(&bar_storage[q]);
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < kOStages; ++q) mbar_init(bar[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_storage)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_storage;
  const uint32_t idP = umma_idesc_tf32(128, MP), id16 = umma_idesc_tf32(128, 16), idX = umma_idesc_tf32(128, NX);
  if (tid < kOStages * 24) {
    const int stg = tid / 24, i = tid % 24, b = i / 6, pass = (i % 6) / 2, k = i % 2;
    const uint32_t lbase = smem_u32(stage0 + (size_t)stg * nfl);
    const uint32_t lsz = 128 * kOT * 4;                      // bytes of one left tile
    const uint32_t sgm_a = lbase + 8 * lsz, sgv_a = sgm_a + 2 * 16 * kOT * 4, sx1_a = sgv_a + 2 * 16 * kOT * 4;
    // block b: left matrix b (a_bar, k', a^2, w) x right tile (k' | g_mean | g_var | [x~,1]) -> TMEM column block
    const uint32_t rb[4] = {lbase + 2 * lsz, sgm_a, sgv_a, sx1_a};
    const uint32_t rrows[4] = {128u, 16u, 16u, (uint32_t)NX};
    const uint32_t dcol[4] = {0u, 128u, 144u, 160u};
    const uint32_t ids[4] = {idP, id16, id16, idX};
    const uint32_t lh = lbase + (2 * b) * lsz, ll = lh + lsz;
    const uint32_t rh = rb[b], rl = rh + rrows[b] * kOT * 4;
    const uint32_t lboL = 128 * 16, lboR = rrows[b] * 16;
    const uint32_t la = (pass == 1) ? ll : lh, ra = (pass == 2) ? rl : rh;   // hi.hi + lo.hi + hi.lo
    dtab[stg][i][0] = umma_desc(la + k * 2 * lboL, lboL, 128);
    dtab[stg][i][1] = umma_desc(ra + k * 2 * lboR, lboR, 128);
    if (stg == 0) { dmeta[i][0] = dcol[b]; dmeta[i][1] = ids[b]; }
  }
  __syncthreads();

  const size_t L = a.m.L;
  const int R = a.m.R;
  const size_t ntile = (L + kOT - 1) / kOT;
  const size_t per = (ntile + gridDim.x - 1) / gridDim.x;
  const size_t t0 = (size_t)blockIdx.x * per, t1 = (t0 + per < ntile) ? t0 + per : ntile;

  // one block = R rows x 4 sixteen-byte chunks, contiguous in memory: item it -> (row = it / 4, chunk = it % 4)
  const int nitem = R * 4;
  constexpr int kMaxItems = 8;                                // ceil((4*128 + 64) * 4 / 256) = 9 at most
  float4 regs[kMaxItems + 1];

  auto fetch = [&](size_t tile) {
    const float4 *src = reinterpret_cast<const float4 *>(a.m.blk + tile * ((size_t)R * 16));
    const size_t col0 = tile * kOT;
    const bool full = col0 + kOT <= L;
#pragma unroll
    for (int q = 0; q <= kMaxItems; ++q) {
      const int it = tid + q * kOThreads;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (it < nitem) {
        v = src[it];
        if (!full) {                      // columns >= L of the last block were never written
          const size_t c = col0 + 4 * (it & 3);
          if (c >= L) v.x = 0.f;
          if (c + 1 >= L) v.y = 0.f;
          if (c + 2 >= L) v.z = 0.f;
          if (c + 3 >= L) v.w = 0.f;
        }
      }
      regs[q] = v;
    }
  };
  auto stash = [&](float *left) {
    float *sgm = left + 8 * 128 * kOT, *sgv = sgm + 2 * 16 * kOT, *sx1 = sgv + 2 * 16 * kOT;
#pragma unroll
    for (int q = 0; q <= kMaxItems; ++q) {
      const int it = tid + q * kOThreads;
      if (it < nitem) {
        const float4 v = regs[q];
        const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
        const int grow = it >> 2, ch = it & 3;
        float *th, *tl;
        int row, rows;
        if (grow < 4 * M) {               // rows a_bar | k' | a^2 | w  ->  left tiles 0..3
          const int mat = grow / M;
          row = grow - mat * M; rows = 128;
          th = left + (size_t)(2 * mat) * 128 * kOT;
          tl = th + 128 * kOT;
        } else {
          const int r2 = grow - 4 * M;
          if (r2 < a.dout) { th = sgm; rows = 16; row = r2; }
          else if (r2 < 2 * a.dout) { th = sgv; rows = 16; row = r2 - a.dout; }
          else { th = sx1; rows = NX; row = r2 - 2 * a.dout; }
          tl = th + rows * kOT;
        }
        // canonical K-major layout: 16-byte chunk ch of row `row` at ch*(rows*16 B) + row*16 B
        *reinterpret_cast<float4 *>(th + (size_t)ch * rows * 4 + row * 4) = hi;
        *reinterpret_cast<float4 *>(tl + (size_t)ch * rows * 4 + row * 4) = lo;
      }
    }
  };
  // this thread's share of the drained accumulators: TMEM lane m = (warp & 3) * 32 + lane, columns
  // [c_begin, c_begin + 96)
  constexpr int kHalf = kOCols / 2;
  float racc[kHalf];
#pragma unroll
  for (int e = 0; e < kHalf; ++e) racc[e] = 0.f;
  const int warp = tid >> 5, lane = tid & 31;
  const int c_begin = (warp < 4) ? 0 : kHalf;
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + c_begin;
  auto drain = [&]() {
    tc_fence_after();
#pragma unroll
    for (int cc = 0; cc < kHalf; cc += 16) {
      float v[16];
      tmem_ld16(trow + cc, v);
#pragma unroll
      for (int e = 0; e < 16; ++e) racc[cc + e] += v[e];
    }
    tc_fence_before();
  };

  uint32_t phase[kOStages];
  bool pending[kOStages];
#pragma unroll
  for (int q = 0; q < kOStages; ++q) { phase[q] = 0; pending[q] = false; }
  auto wait_stage = [&](int q) {
    if (pending[q]) {
      mbar_wait(bar[q], phase[q]);
      phase[q] ^= 1;
      pending[q] = false;
    }
  };
  int since = 0, st = 0;
  if (t0 < t1) fetch(t0);
  for (size_t tile = t0; tile < t1; ++tile) {
    // the MMAs that last read this stage must be done before it is overwritten
#pragma unroll
    for (int q = 0; q < kOStages; ++q)
      if (q == st) wait_stage(q);
    float *left = stage0 + (size_t)st * nfl;
    stash(left);
    async_proxy_fence();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t acc0 = since > 0 ? 1u : 0u;
#pragma unroll 6
      for (int i = 0; i < 24; ++i)
        umma_tf32(tmem + dmeta[i][0], dtab[st][i][0], dtab[st][i][1], dmeta[i][1], (i % 6 == 0) ? acc0 : 1u);
      umma_commit(bar[st]);
    }
#pragma unroll
    for (int q = 0; q < kOStages; ++q)
      if (q == st) pending[q] = true;
    if (tile + 1 < t1) fetch(tile + 1);        // overlaps the MMAs
    if (++since == kODrain || tile + 1 == t1) {
#pragma unroll
      for (int q = 0; q < kOStages; ++q) wait_stage(q);   // commits complete in order; drain needs all of them
      drain();
      since = 0;
    }
    st = (st + 1 == kOStages) ? 0 : st + 1;
  }
  {
    double *out = Rpart + ((size_t)blockIdx.x * 128 + (warp & 3) * 32 + lane) * kOCols + c_begin;
#pragma unroll
    for (int e = 0; e < kHalf; ++e) out[e] = (double)racc[e];
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}
// R[m][c] (float64, [M][Ctot]) = sum over CTAs of the partials, with the TMEM column blocks mapped to
// [P_bar' (M) | alpha_bar' (dout) | S_bar (dout) | U, r (din+1)].
__global__ void outer_reduce_kernel(const double *__restrict__ Rpart, int nparts, int M, int dout, int din,
                                    double *__restrict__ R) {
  const int Ctot = M + 2 * dout + din + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * Ctot) return;
  const int m = i / Ctot, c = i - m * Ctot;
  int src;
  if (c < M) src = c;
  else if (c < M + dout) src = 128 + (c - M);
  else if (c < M + 2 * dout) src = 144 + (c - M - dout);
  else src = 160 + (c - M - 2 * dout);
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += Rpart[((size_t)p * 128 + m) * kOCols + src];
  R[i] = s;
}

inline size_t outer_smem_bytes(int din) {
  const int NX = round_up(din + 1, 16);
  return sizeof(float) * kOStages * (8 * 128 * kOT + 4 * 16 * kOT + 2 * NX * kOT);
}

}  // namespace cbf
