// Tensor-core (tcgen05 / TMEM) forward rollout kernels for large M (16 <= M <= 128).
//
// For a CTA tile of 128 particles the M x M contraction a = P k of every time step is a real
// GEMM, D[128 x MP] = K[128 x MP] . P[MP x MP], and runs on the 5th-generation tensor cores:
//   * one thread = one particle = one TMEM lane; each step the thread evaluates its kernel vector
//     k'' = k / (sigma^2 max_m k) in (0,1] (SIMT: one packed-FMA chain + one MUFU.EX2 per inducing point; the squared
//     distances wait in the thread's own, still idle accumulator columns while their minimum is found), splits it
//     into two fp16 terms k'' = h1 + h2 (22 significant bits together) and writes both as its row of the A operand
//     pair -- in TMEM, next to the accumulator (tcgen05.st; lane = row, one 32-bit column = two K elements:
//     TcCtx::kATmem, tools/microbench/tmem_a.cu).  Only three-tile CTAs, whose 3 x 256 columns do not exist, keep
//     the operands in shared memory (K1, K2: canonical no-swizzle core-matrix layout, 16-byte chunk c of row r at
//     c*2048 + r*16: conflict-free 128-bit stores);
//   * P' = P / 2^e (|P'| <= 1024) is split the same way once per launch into the B operands P1, P2 (shared memory);
//   * one elected thread issues 3 x MP/16 tcgen05.mma.kind::f16 (K1 P1 + K2 P1 + K1 P2, A from TMEM, fp32
//     accumulation in TMEM; the dropped h2*h2 term is < 2^-22 relative) and commits to an mbarrier;
//   * every thread reads its accumulator row back with tcgen05.ld (16 columns at a time) and
//     forms k.a and sum_m a_m^2 S_md in fp32 (SIMT), re-reading its own k'' row from the operand columns.
// A CTA is 57 KB of shared memory (P1, P2, the small GP tables) and 256 TMEM columns: two CTAs per SM.
// The O(M.D) work stays on the SIMT pipes; only the 2 M^2 FLOPs per evaluation move to the tensor
// pipe.
//
// Reverse mode (fw_reverse_tc / bm_reverse_tc): the same tile does, per step, the recomputation
// a = P k and the second contraction P b on the tensor core (b is scaled per particle by a power
// of two into fp16 range and split the same way), the step adjoint and the k_bar / x_bar chain in
// SIMT, and writes the operands of the parameter-adjoint outer products (a_bar, k, a^2, w, g_mean,
// g_var, x~) as bfloat16 hi/lo pairs in the MMA-ready tile layout of common.cuh TcMats to the workspace
// (16 bytes per 8 rows per thread, fully coalesced); the accumulation over (particle, step) is the
// tcgen05 split-K reduction of kernels_outer.cuh.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kernels_fast.cuh"

namespace cbf {

constexpr int kTcThreads = 128;   // threads (= particles = TMEM lanes) of one particle tile
#ifndef CBF_TC_CHUNK_UNROLL
#define CBF_TC_CHUNK_UNROLL 1
#endif
constexpr int kTcChunkUnroll = CBF_TC_CHUNK_UNROLL;   // unroll of the 16-row chunk loops when M is a compile-time constant

// Barrier over the 128 threads of one particle tile (named barrier 1 + tile; 0 is __syncthreads).
__device__ __forceinline__ void tile_sync(int tile) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + tile), "r"(kTcThreads) : "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor):
// bits [0,14) start>>4, [16,30) leading-dim byte offset>>4 (between the two 16-byte K chunks of one
// instruction), [32,46) stride-dim byte offset>>4 (between 8-row groups), [46,48) version = 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): D = fp32, A = B = fp16, both K-major.
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// the same with the A operand in TMEM (lane = row, 8 columns = 16 packed K elements per instruction)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tMBAR_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra MBAR_DONE;\n\tbra MBAR_WAIT;\n\tMBAR_DONE:\n\t}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void async_proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16 consecutive accumulator columns of this thread's TMEM lane.  The issue and the wait are separate so
// that several loads (and independent global / shared loads) can be in flight before the first use.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Wait that carries the loaded registers as in/out operands, so no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16], uint32_t (&q)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(q[0]),
                 "+r"(q[1]), "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7]), "+r"(q[8]), "+r"(q[9]),
                 "+r"(q[10]), "+r"(q[11]), "+r"(q[12]), "+r"(q[13]), "+r"(q[14]), "+r"(q[15])
               :
               : "memory");
}
// 16 consecutive accumulator columns of this thread's TMEM lane written from registers.
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
      :
      : "r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),
        "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 4 consecutive 32-bit columns (8 packed 16-bit operand elements) of this thread's TMEM lane.
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&w)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3])
               : "memory");
}
__device__ __forceinline__ void tmem_ld4_issue(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait4(uint32_t (&a)[4], uint32_t (&b)[4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  tmem_ld16_issue(taddr, r);
  tmem_ld_wait(r);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
// x = h1 + h2 with two fp16 terms (round-to-nearest each)
__device__ __forceinline__ void split_h(float x, __half &h1, __half &h2) {
  h1 = __float2half_rn(x);
  h2 = __float2half_rn(x - __half2float(h1));
}

// Shared-memory / TMEM state of one CTA: the B operand pair P1/P2, resident small GP tables, barrier, TMEM base.
// NT particle tiles per CTA (NT x 128 threads): each tile has its own operand rows, TMEM columns, mbarrier
// and named barrier and runs independently; the tiles share P1/P2 and the small tables.  NT = 2 is used
// where one tile's CTA is too large for two CTAs per SM (large tables at M = 128), which restores 8 warps per SM.
//
// NG > 1 (latency variant, NT = 1 only): NG x 128 threads work on ONE particle tile, thread (g, lane) handling the
// 16-row chunks cc = g, g + NG, ... of particle `lane`'s M-vectors (warps w and w + 4 read the same TMEM lane
// quarter).  Per-particle partial sums cross the groups through the small exchange buffer `xch` in a fixed order, so
// all NG threads of a particle carry bit-identical per-particle state.  It shortens the serial chain of one time
// step by ~NG and is used when the launch has too few particle tiles to fill the SMs anyway (small minibatches,
// strong scaling, prediction of one long sequence).
template <int DIN, int DOUT, bool REV = false, int MC_ = 0, int NT_ = 1, int NG_ = 1>
struct TcCtx {
  static constexpr int NT = NT_;
  static constexpr int NG = NG_;
  static_assert(NG == 1 || NT == 1, "the M-split variant runs one particle tile per CTA");
  static constexpr int XV = 1 + (2 * DOUT + 2) + DIN;   // exchange values per (group, particle): d2min | q, amax, fm, fv | x_bar
  static constexpr int MC = MC_;     // compile-time M (0: runtime) -- lets the M loops unroll and drop their guards
  static constexpr int DINP = (DIN + 3) / 4 * 4, DOUTP = (DOUT + 3) / 4 * 4;
  // The A operands of the contractions (k'' and, in the reverse pass, b'': two fp16 terms each) live in TMEM next to
  // the accumulator when the tile's 256 columns fit: each thread writes its own particle's row with tcgen05.st (lane =
  // row; column c holds the K elements 2c, 2c+1, low half first -- tools/microbench/tmem_a.cu) and tcgen05.mma reads
  // them from there.  That takes the operand stores, their read-backs and the tensor core's A reads off the
  // shared-memory data path, the busiest unit of these kernels.  Three-tile CTAs (3 x 256 columns do not exist) keep
  // the operands in shared memory.
  static constexpr bool kATmem = NT <= 2;
  static constexpr uint32_t TMEM_COLS = kATmem ? 256u : 128u;   // accumulator (MP <= 128 columns) [+ 2 x MP/2 operand columns]
  static constexpr uint32_t kAOff = 128u;                        // first operand column of a tile
  static constexpr uint32_t kTmemAlloc = NT * TMEM_COLS <= 256u ? 256u : 512u;   // allocations are powers of two
  __half *P1, *P2, *K1, *K2, *B1, *B2;
  const float *Zt, *al, *Sm, *il;
  float *xch;           // (NG > 1) [NG][XV][128] partial sums between the groups of a particle
  uint32_t bar, tmem, tmem_base, idesc;
  uint32_t phase;
  int M, MP;
  int tile_, tl_;       // (NT > 1 or NG > 1) particle tile / group of this thread within the CTA, lane (= TMEM lane, operand row) in the tile
  __device__ __forceinline__ int tile_id() const { return NT == 1 ? 0 : tile_; }
  __device__ __forceinline__ int grp() const { return NG == 1 ? 0 : tile_; }
  __device__ __forceinline__ int lane_id() const { return (NT == 1 && NG == 1) ? (int)threadIdx.x : tl_; }
  __device__ __forceinline__ float *xslot(int g, int v) const { return xch + ((size_t)(g * XV + v)) * kTcThreads + tl_; }
  float sig2, pscale;   // P = pscale * P'
  float smax[DOUT];     // max_m S_md (bound used to scale b in the reverse pass)

  // shared-memory operand buffers (k'' / b'' rows and the parked squared distances): none when they live in TMEM
  __host__ __device__ static size_t kbuf_bytes(int MP) { return kATmem ? 0 : (size_t)NT * 2 * kTcThreads * MP * 2; }
  static size_t bytes(int M) {
    const int MP = round_up(M, 16);
    return (size_t)2 * MP * MP * 2 + kbuf_bytes(MP) +
           sizeof(float) * ((size_t)MP * (DINP + 2 * DOUTP) + DINP + 4) + 64 +
           (NG > 1 ? sizeof(float) * (size_t)NG * XV * kTcThreads : 0);
  }

  // Carve + fill (all threads).  Allocates TMEM (warp 0) and initialises the mbarrier.
  __device__ unsigned char *init(unsigned char *base, const GpDev &g, int M_, float *scratch) {
    M = M_; MP = round_up(M, 16);
    P1 = reinterpret_cast<__half *>(base); base += (size_t)MP * MP * 2;
    P2 = reinterpret_cast<__half *>(base); base += (size_t)MP * MP * 2;
    tile_ = threadIdx.x / kTcThreads; tl_ = threadIdx.x % kTcThreads;
    const int tile = tile_id();     // 0 in the M-split variant: all groups share one tile's buffers
    K1 = reinterpret_cast<__half *>(base) + (kATmem ? 0 : (size_t)(2 * tile) * kTcThreads * MP);   // (unused with kATmem)
    K2 = K1 + (kATmem ? 0 : (size_t)kTcThreads * MP);
    base += kbuf_bytes(MP);
    // The reverse pass reuses the K operand rows for b (k is re-read from the operand tile it was just written to,
    // an L2 hit), so one operand pair per tile suffices.
    B1 = K1; B2 = K2;
    {   // every row group the kernels visit is rewritten each evaluation; a padding group they skip (for_chunks) must
        // read as zeros in the contractions, so the buffers start zeroed
      uint4 *kz = reinterpret_cast<uint4 *>(base - kbuf_bytes(MP));
      const int n16 = (int)(kbuf_bytes(MP) / 16);
      for (int i = threadIdx.x; i < n16; i += blockDim.x) kz[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    // The tables have MP rows (Zt holds -Z/ell so that delta is one packed add): rows >= M hold Z/ell = 1e18 (squared distance ~1e37, so k'' underflows to an
    // exact 0 and never wins the minimum) and alpha = S = 0, which makes every padded row contribute exact
    // zeros everywhere -- the per-row loops need no m < M guards.
    float *Zw = reinterpret_cast<float *>(base); base += sizeof(float) * MP * DINP;
    float *aw = reinterpret_cast<float *>(base); base += sizeof(float) * MP * DOUTP;
    float *Sw = reinterpret_cast<float *>(base); base += sizeof(float) * MP * DOUTP;
    float *iw = reinterpret_cast<float *>(base); base += sizeof(float) * (DINP + 4);
    uint64_t *barp = reinterpret_cast<uint64_t *>(base); base += 32;   // one mbarrier per tile
    uint32_t *tmemp = reinterpret_cast<uint32_t *>(base); base += 16;
    base += 16;
    xch = reinterpret_cast<float *>(base);
    if (NG > 1) base += sizeof(float) * (size_t)NG * XV * kTcThreads;
    const int tid = threadIdx.x, nt = blockDim.x;
    // scale of P: |P / 2^e| <= 1024
    float mx = 0.f;
    for (int i = tid; i < M * M; i += nt) mx = fmaxf(mx, fabsf(g.P[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) scratch[tid >> 5] = mx;
    __syncthreads();
    mx = 0.f;
    for (int w = 0; w < (nt >> 5); ++w) mx = fmaxf(mx, scratch[w]);
    int e = 0;
    frexpf(fmaxf(mx, 1e-30f), &e);   // mx = f * 2^e, f in [0.5, 1)
    e -= 10;
    pscale = ldexpf(1.f, e);
    const float inv = ldexpf(1.f, -e);
    // B operand: element (n, kk) = P'[kk][n]; K-major chunks: c = kk/8 at c*(MP*8) + n*8 + kk%8
    for (int i = tid; i < MP * MP; i += nt) {
      const int c = i / (MP * 8), rem = i - c * (MP * 8), n = rem >> 3, kk = c * 8 + (rem & 7);
      const float v = (n < M && kk < M) ? g.P[kk * M + n] * inv : 0.f;
      __half h1, h2;
      split_h(v, h1, h2);
      P1[i] = h1;
      P2[i] = h2;
    }
    for (int i = tid; i < MP * DINP; i += nt) {
      const int r = i / DINP, c = i % DINP;
      Zw[i] = (c < DIN) ? (r < M ? -g.Z[r * DIN + c] / g.ell[c] : -1e18f) : 0.f;   // -Z/ell: delta = x~ + Zt
    }
    for (int i = tid; i < MP * DOUTP; i += nt) {
      const int r = i / DOUTP, c = i % DOUTP;
      const bool ok = (c < DOUT && r < M);
      aw[i] = ok ? g.alpha[r * DOUT + c] : 0.f;
      Sw[i] = ok ? g.S[r * DOUT + c] : 0.f;
    }
    for (int i = tid; i < DINP; i += nt) iw[i] = (i < DIN) ? 1.f / g.ell[i] : 0.f;
    Zt = Zw; al = aw; Sm = Sw; il = iw;
    sig2 = g.sig2[0];
#pragma unroll
    for (int d = 0; d < DOUT; ++d) smax[d] = 0.f;
    if (REV) {
      for (int m = 0; m < M; ++m)
#pragma unroll
        for (int d = 0; d < DOUT; ++d) smax[d] = fmaxf(smax[d], g.S[m * DOUT + d]);
    }
    bar = smem_u32(barp + tile);
    idesc = umma_idesc_f16(128, MP);
    phase = 0;
    if (tid == 0) {
#pragma unroll
      for (int q = 0; q < NT; ++q) mbar_init(smem_u32(barp + q), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (tid < 32) {   // one warp allocates the TMEM columns (fp32 accumulators, 128 lanes x MP <= 128 per tile)
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmemp)),
                   "r"(kTmemAlloc)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    async_proxy_fence();     // P1/P2 written through the generic proxy, read by the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tmem_base = *tmemp;
    tmem = tmem_base + (uint32_t)tile * TMEM_COLS;
    if constexpr (kATmem) {   // operand columns start as zeros (TMEM comes uninitialised; a padding group is never written)
      const uint32_t z[4] = {0u, 0u, 0u, 0u};
      const uint32_t trow = tmem + ((uint32_t)(tl_ & ~31) << 16) + kAOff;
      for (int c4 = 0; c4 < MP / 4; ++c4) tmem_st4(trow + c4 * 4, z);
      tmem_st_wait();
    }
    return base;
  }

  __device__ void release() {
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemAlloc) : "memory");
  }
};

// D[tmem_d] (+)= A1 P1 + A2 P1 + A1 P2 for the tile's 128 rows; all threads of the tile call (contains the tile barrier
// and the mbarrier wait).  a1/a2: fp16 split A operands written by the threads just before.
template <class Ctx>
__device__ __forceinline__ void tc_contract(Ctx &c, const __half *a1p, const __half *a2p, uint32_t tmem_d, uint32_t acc0 = 0) {
  if constexpr (Ctx::kATmem) tmem_st_wait();   // this thread's operand rows (and scaled accumulator row) are in TMEM
  async_proxy_fence();
  tc_fence_before();
  if (Ctx::NT == 1) __syncthreads(); else tile_sync(c.tile_id());
  if (c.lane_id() == 0 && c.grp() == 0) {
    tc_fence_after();
    const uint32_t lboA = kTcThreads * 16, lboB = c.MP * 16;
    const uint32_t a1 = smem_u32(a1p), a2 = smem_u32(a2p), b1 = smem_u32(c.P1), b2 = smem_u32(c.P2);
    const uint32_t ta1 = c.tmem + Ctx::kAOff, ta2 = ta1 + c.MP / 2;   // (kATmem) operand terms: MP / 2 columns each
    const int ks = c.MP / 16;
    uint32_t acc = acc0;   // 1: add to what the threads stored in the accumulator columns
    for (int pass = 0; pass < 3; ++pass) {
      const uint32_t ab = (pass == 1) ? a2 : a1, bb = (pass == 2) ? b2 : b1, tab = (pass == 1) ? ta2 : ta1;
      for (int k = 0; k < ks; ++k) {
        if constexpr (Ctx::kATmem)
          umma_f16_ts(tmem_d, tab + k * 8, umma_desc(bb + k * 2 * lboB, lboB, 128), c.idesc, acc);   // 8 columns = 16 K elements
        else
          umma_f16(tmem_d, umma_desc(ab + k * 2 * lboA, lboA, 128), umma_desc(bb + k * 2 * lboB, lboB, 128), c.idesc, acc);
        acc = 1;
      }
    }
    umma_commit(c.bar);
  }
  mbar_wait(c.bar, c.phase);
  c.phase ^= 1;
  tc_fence_after();
}

// The 16-row chunks cc = g0, g0 + NG, ... of an M-vector.  With a compile-time M whose last chunk has at most 8 live
// rows (M = 100: rows 96..99) that chunk is visited as ONE 8-row group (Rows<8>): its second group is all padding,
// the operand buffers keep the zeros they were initialised with there, and 8 of 112 rows of per-row work go away.
template <int N> struct Rows { static constexpr int value = N; };
template <int MC, int NG, class F>
__device__ __forceinline__ void for_chunks(int g0, int nch, F &&body) {
  constexpr bool kTail = MC > 0 && (MC % 16) >= 1 && (MC % 16) <= 8;
  if constexpr (kTail) {
#pragma unroll 1
    for (int cc = g0; cc < nch - 1; cc += NG) body(cc, Rows<16>{});
    if (NG == 1 || (nch - 1) % NG == g0) body(nch - 1, Rows<8>{});
  } else {
#pragma unroll 1
    for (int cc = g0; cc < nch; cc += NG) body(cc, Rows<16>{});
  }
}

// This thread's row (ROWS fp16-split values starting at column 16*cc) of an A-operand buffer pair.
template <int ROWS = 16>
__device__ __forceinline__ void tc_read_row16(const __half *b1, const __half *b2, int row, int cc, float (&v)[16]) {
#pragma unroll
  for (int hch = 0; hch < ROWS / 8; ++hch) {
    const size_t off = (size_t)(cc * 2 + hch) * (kTcThreads * 8) + row * 8;
    const uint4 v1 = *reinterpret_cast<const uint4 *>(b1 + off);
    const uint4 v2 = *reinterpret_cast<const uint4 *>(b2 + off);
    const uint32_t w1[4] = {v1.x, v1.y, v1.z, v1.w}, w2[4] = {v2.x, v2.y, v2.z, v2.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&w1[e]));
      const float2 f2 = __half22float2(*reinterpret_cast<const __half2 *>(&w2[e]));
      v[hch * 8 + 2 * e] = f1.x + f2.x;
      v[hch * 8 + 2 * e + 1] = f1.y + f2.y;
    }
  }
}
__device__ __forceinline__ void tc_write_row8(__half *b1, __half *b2, int row, int ch, const float (&v)[8]) {
  uint32_t w1[4], w2[4];   // two values per conversion (F2FP) instead of one (F2F)
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __half2 h1 = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
    const float2 f1 = __half22float2(h1);
    const __half2 h2 = __floats2half2_rn(v[2 * e] - f1.x, v[2 * e + 1] - f1.y);
    w1[e] = *reinterpret_cast<const uint32_t *>(&h1);
    w2[e] = *reinterpret_cast<const uint32_t *>(&h2);
  }
  const uint4 v1 = make_uint4(w1[0], w1[1], w1[2], w1[3]);
  const uint4 v2 = make_uint4(w2[0], w2[1], w2[2], w2[3]);
  const size_t off = (size_t)ch * (kTcThreads * 8) + row * 8;
  *reinterpret_cast<uint4 *>(b1 + off) = v1;
  *reinterpret_cast<uint4 *>(b2 + off) = v2;
}

// Operand rows of the tile's A operand pair, in TMEM (Ctx::kATmem) or in the shared-memory buffers b1 / b2.
template <class Ctx>
__device__ __forceinline__ void tc_put_row8(const Ctx &c, __half *b1, __half *b2, int row, int ch, const float (&v)[8]) {
  if constexpr (Ctx::kATmem) {
    uint32_t w1[4], w2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __half2 h1 = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
      const float2 f1 = __half22float2(h1);
      const __half2 h2 = __floats2half2_rn(v[2 * e] - f1.x, v[2 * e + 1] - f1.y);
      w1[e] = *reinterpret_cast<const uint32_t *>(&h1);
      w2[e] = *reinterpret_cast<const uint32_t *>(&h2);
    }
    const uint32_t trow = c.tmem + ((uint32_t)(row & ~31) << 16) + Ctx::kAOff + ch * 4;
    tmem_st4(trow, w1);
    tmem_st4(trow + c.MP / 2, w2);
  } else {
    tc_write_row8(b1, b2, row, ch, v);
  }
}
template <int ROWS, class Ctx>
__device__ __forceinline__ void tc_get_row16(const Ctx &c, const __half *b1, const __half *b2, int row, int cc, float (&v)[16]) {
  if constexpr (Ctx::kATmem) {
    const uint32_t trow = c.tmem + ((uint32_t)(row & ~31) << 16) + Ctx::kAOff + cc * 8;
    uint32_t w1[ROWS / 8][4], w2[ROWS / 8][4];
#pragma unroll
    for (int hch = 0; hch < ROWS / 8; ++hch) {
      tmem_ld4_issue(trow + hch * 4, w1[hch]);
      tmem_ld4_issue(trow + c.MP / 2 + hch * 4, w2[hch]);
    }
#pragma unroll
    for (int hch = 0; hch < ROWS / 8; ++hch) {
      tmem_ld_wait4(w1[hch], w2[hch]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&w1[hch][e]));
        const float2 f2 = __half22float2(*reinterpret_cast<const __half2 *>(&w2[hch][e]));
        v[hch * 8 + 2 * e] = f1.x + f2.x;
        v[hch * 8 + 2 * e + 1] = f1.y + f2.y;
      }
    }
  } else {
    tc_read_row16<ROWS>(b1, b2, row, cc, v);
  }
}

// x = hi + lo with two bfloat16 terms (round-to-nearest each): 8 values -> two 16-byte segments.
__device__ __forceinline__ void split_bf16x8(const float (&v)[8], uint4 &hi, uint4 &lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    h[e] = *reinterpret_cast<const uint32_t *>(&hh);
    const float h0 = __uint_as_float(h[e] << 16), h1 = __uint_as_float(h[e] & 0xFFFF0000u);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * e] - h0, v[2 * e + 1] - h1);
    l[e] = *reinterpret_cast<const uint32_t *>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// This particle-step's column of the outer-product operand tile (common.cuh TcMats): row-block r of the
// tile is at col + r * kOBlk, its lo copy RB blocks further.
struct TcOut {
  unsigned char *col;
  int RB, MB, DB, XB;
  int bAb, bK, bA2, bW, bGm, bGv, bX1;
  __device__ __forceinline__ void put8(int blk, const float (&v)[8]) const {
    uint4 hi, lo;
    split_bf16x8(v, hi, lo);
    *reinterpret_cast<uint4 *>(col + (size_t)blk * kOBlk) = hi;
    *reinterpret_cast<uint4 *>(col + (size_t)(blk + RB) * kOBlk) = lo;
  }
  __device__ __forceinline__ void get8(int blk, float (&v)[8]) const {
    const uint4 hi = *reinterpret_cast<const uint4 *>(col + (size_t)blk * kOBlk);
    const uint4 lo = *reinterpret_cast<const uint4 *>(col + (size_t)(blk + RB) * kOBlk);
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e] = __uint_as_float(h[e] << 16) + __uint_as_float(l[e] << 16);
      v[2 * e + 1] = __uint_as_float(h[e] & 0xFFFF0000u) + __uint_as_float(l[e] & 0xFFFF0000u);
    }
  }
  // raw hi/lo segments of row-blocks blk and blk+1 (zeros when absent): q = {hi0, lo0, hi1, lo1}
  __device__ __forceinline__ void get8_raw(int blk, bool live, uint4 (&q)[4]) const {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    q[0] = q[1] = q[2] = q[3] = z;
    if (live && blk - bK < MB) {
      q[0] = *reinterpret_cast<const uint4 *>(col + (size_t)blk * kOBlk);
      q[1] = *reinterpret_cast<const uint4 *>(col + (size_t)(blk + RB) * kOBlk);
    }
    if (live && blk + 1 - bK < MB) {
      q[2] = *reinterpret_cast<const uint4 *>(col + (size_t)(blk + 1) * kOBlk);
      q[3] = *reinterpret_cast<const uint4 *>(col + (size_t)(blk + 1 + RB) * kOBlk);
    }
  }
  static __device__ __forceinline__ void unpack8(const uint4 &hi, const uint4 &lo, float (&v)[8]) {
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e] = __uint_as_float(h[e] << 16) + __uint_as_float(l[e] << 16);
      v[2 * e + 1] = __uint_as_float(h[e] & 0xFFFF0000u) + __uint_as_float(l[e] & 0xFFFF0000u);
    }
  }
  // n values (n <= 8 * nblk) as nblk row-blocks starting at blk, zero padded
  template <int N>
  __device__ __forceinline__ void put_vec(int blk, const float (&v)[N]) const {
#pragma unroll
    for (int r = 0; r < (N + 7) / 8; ++r) {
      float w[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) w[e] = (8 * r + e < N) ? v[8 * r + e < N ? 8 * r + e : 0] : 0.f;
      put8(blk + r, w);
    }
  }
};

// One sparse-GP evaluation (gp_tf.py:132-161) for the CTA's 128 particles; every thread must call.
// kout (optional): operand-tile column receiving the normalised k'' for the outer-product GEMMs (gp_reverse_tc
// folds kscale into the partner operands a_bar and g_mean).
// amax: max_m |a''_m| of this particle's normalised accumulator row (scales b in the reverse pass);
// kscale: the particle's normalisation, k' = kscale * k''.
// With NG groups (Ctx::NG > 1) a thread handles the 16-row chunks cc = grp, grp + NG, ...; every result
// (fm, fv, amax, kscale) is the same in all NG threads of a particle.
// light (reverse kernels, when the forward kernels saved fm / fv / amax): only the kernel vector, its operand rows
// and D1 = K P' are produced; fm, fv, amax are left as the caller set them.
template <class Ctx, int DIN, int DOUT>
__device__ __forceinline__ void gp_forward_tc(Ctx &c, const float (&xin)[DIN], float (&xt)[(DIN + 3) / 4 * 4],
                                              float (&fm)[DOUT], float (&fv)[DOUT], const TcOut *kout,
                                              float &amax, float &kscale, const bool light = false) {
  constexpr int DINP = (DIN + 3) / 4 * 4, DOUTP = (DOUT + 3) / 4 * 4;
  constexpr int MC = Ctx::MC, NG = Ctx::NG;
  const int t = c.lane_id(), g0 = c.grp(), M = MC ? MC : c.M, MP = MC ? (MC + 15) / 16 * 16 : c.MP;
  (void)M;
  {
    float il[DINP];
    ld_row<DINP>(c.il, il);
#pragma unroll
    for (int j = 0; j < DINP; ++j) xt[j] = (j < DIN) ? xin[j < DIN ? j : 0] * il[j] : 0.f;
  }
  // packed FP32 (FADD2 / FFMA2 / FMUL2): pairs of input dims for delta, pairs of output dims for the sums
  constexpr int J2 = (DIN + 1) / 2, D2 = (DOUT + 1) / 2;
  // In the latency variant every sum over the inducing rows runs as two interleaved chains (even / odd rows): with
  // few warps a single 16-long dependent FMA chain per chunk leaves the issue slot idle for its latency.  The
  // throughput kernels keep one chain (measured: the second set of accumulators costs 1.4 % of the step at the bench
  // batch, and spills at DIN = 21 / DOUT = 14).
  constexpr bool kTwo = NG > 1 && DIN <= 8;
  constexpr int D2b = kTwo ? D2 : 1;
  unsigned long long x2[J2], fm2[D2], fv2[D2], fm2b[D2b], fv2b[D2b];
#pragma unroll
  for (int j = 0; j < J2; ++j) x2[j] = pack2(xt[2 * j], xt[2 * j + 1]);   // xt is zero-padded to DINP
#pragma unroll
  for (int d = 0; d < D2; ++d) { fm2[d] = 0ull; fv2[d] = 0ull; }
#pragma unroll
  for (int d = 0; d < D2b; ++d) { fm2b[d] = 0ull; fv2b[d] = 0ull; }
  // ---- kernel vector -> fp16 split operands ----
  // fp16 has a narrow exponent range and k' = exp(-d^2/2) can be 1e-12 for every inducing point (e.g.
  // 21 input dims), so each particle's vector is normalised by its own maximum: pass 1 parks the squared
  // distances (float32) in the thread's own K1/K2 slots and finds their minimum, pass 2 forms
  // k'' = exp(-(d^2 - d^2_min)/2) in (0,1] (max exactly 1), splits and overwrites.  k' = kscale * k''.
  float d2min = 3.0e38f, d2minb = 3.0e38f;
  for_chunks<MC, NG>(g0, MP / 16, [&](const int cc, auto rows_tag) {
#pragma unroll
  for (int hh = 0; hh < decltype(rows_tag)::value / 8; ++hh) {   // 8-row chunks 2cc, 2cc+1 of the own 16-row chunks cc
    const int ch = 2 * cc + hh;
    float dv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int m = ch * 8 + e;
      float z[DINP];
      ld_row<DINP>(c.Zt + m * DINP, z);
      unsigned long long acc = 0ull;
#pragma unroll
      for (int j = 0; j < J2; ++j) {
        const unsigned long long dl = add2(x2[j], pack2(z[2 * j], z[2 * j + 1]));
        acc = fma2(dl, dl, acc);
      }
      const float d2 = hsum2(acc);
      if (kTwo && (e & 1)) d2minb = fminf(d2minb, d2); else d2min = fminf(d2min, d2);
      dv[e] = d2;
    }
    if constexpr (Ctx::kATmem) {   // parked in the (idle until the contraction) accumulator columns of the own lane
      const uint32_t w0[4] = {__float_as_uint(dv[0]), __float_as_uint(dv[1]), __float_as_uint(dv[2]), __float_as_uint(dv[3])};
      const uint32_t w1[4] = {__float_as_uint(dv[4]), __float_as_uint(dv[5]), __float_as_uint(dv[6]), __float_as_uint(dv[7])};
      const uint32_t trow = c.tmem + ((uint32_t)(t & ~31) << 16) + ch * 8;
      tmem_st4(trow, w0);
      tmem_st4(trow + 4, w1);
    } else {
      const size_t off = (size_t)ch * (kTcThreads * 8) + t * 8;
      *reinterpret_cast<float4 *>(c.K1 + off) = make_float4(dv[0], dv[1], dv[2], dv[3]);
      *reinterpret_cast<float4 *>(c.K2 + off) = make_float4(dv[4], dv[5], dv[6], dv[7]);
    }
  }
  });
  if constexpr (Ctx::kATmem) tmem_st_wait();
  d2min = fminf(d2min, d2minb);
  if (NG > 1) {   // the particle's minimum over all groups
    *c.xslot(g0, 0) = d2min;
    __syncthreads();
#pragma unroll
    for (int g = 0; g < NG; ++g) d2min = fminf(d2min, *c.xslot(g, 0));
  }
  kscale = fast_exp2(kNegHalfLog2e * d2min);
  for_chunks<MC, NG>(g0, MP / 16, [&](const int cc, auto rows_tag) {
#pragma unroll
  for (int hh = 0; hh < decltype(rows_tag)::value / 8; ++hh) {
    const int ch = 2 * cc + hh;
    float dv[8];
    if constexpr (Ctx::kATmem) {
      uint32_t w0[4], w1[4];
      const uint32_t trow = c.tmem + ((uint32_t)(t & ~31) << 16) + ch * 8;
      tmem_ld4_issue(trow, w0);
      tmem_ld4_issue(trow + 4, w1);
      tmem_ld_wait4(w0, w1);
#pragma unroll
      for (int e = 0; e < 4; ++e) { dv[e] = __uint_as_float(w0[e]); dv[4 + e] = __uint_as_float(w1[e]); }
    } else {
      const size_t off = (size_t)ch * (kTcThreads * 8) + t * 8;
      const float4 da = *reinterpret_cast<const float4 *>(c.K1 + off), db = *reinterpret_cast<const float4 *>(c.K2 + off);
      dv[0] = da.x; dv[1] = da.y; dv[2] = da.z; dv[3] = da.w; dv[4] = db.x; dv[5] = db.y; dv[6] = db.z; dv[7] = db.w;
    }
    float kv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int m = ch * 8 + e;
      const float kp = fast_exp2(kNegHalfLog2e * (dv[e] - d2min));
      if (!light) {
        float al[DOUTP];
        ld_row<DOUTP>(c.al + m * DOUTP, al);
        const unsigned long long kk = pack2(kp, kp);
#pragma unroll
        for (int d = 0; d < D2; ++d) {
          if (kTwo && (e & 1)) fm2b[kTwo ? d : 0] = fma2(pack2(al[2 * d], al[2 * d + 1]), kk, fm2b[kTwo ? d : 0]);
          else fm2[d] = fma2(pack2(al[2 * d], al[2 * d + 1]), kk, fm2[d]);
        }
      }
      kv[e] = kp;
    }
    tc_put_row8(c, c.K1, c.K2, t, ch, kv);
    if (kout && ch < kout->MB) kout->put8(kout->bK + ch, kv);   // the normalised k'': kscale goes into its partners
  }
  });
  // ---- D1 = K P' on the tensor core ----
  tc_contract(c, c.K1, c.K2, c.tmem);
  if (light) {     // uniform over the CTA: the moments come from the forward pass
    tc_fence_before();
    return;
  }
  // ---- accumulator row back: q' = k'.a', v'_d = sum a'^2 S ----
  float q = 0.f, qb = 0.f, amaxb = 0.f;
  amax = 0.f;
  const uint32_t trow = c.tmem + ((uint32_t)(t & ~31) << 16);   // this warp's 32-lane quarter
  uint32_t ra[16];
  if (g0 < MP / 16) tmem_ld16_issue(trow + g0 * 16, ra);
  for_chunks<MC, NG>(g0, MP / 16, [&](const int cc, auto rows_tag) {
    constexpr int ROWS = decltype(rows_tag)::value;
    float a[16], kp[16];
    tc_get_row16<ROWS>(c, c.K1, c.K2, t, cc, kp);
    tmem_ld_wait(ra);
#pragma unroll
    for (int e = 0; e < 16; ++e) a[e] = __uint_as_float(ra[e]);
    if (cc + NG < MP / 16) tmem_ld16_issue(trow + (cc + NG) * 16, ra);   // next chunk in flight during this one's math
#pragma unroll
    for (int e = 0; e < ROWS; ++e) {
      const int m = cc * 16 + e;
      const float a2 = a[e] * a[e];
      float S[DOUTP];
      ld_row<DOUTP>(c.Sm + m * DOUTP, S);
      const unsigned long long aa = pack2(a2, a2);
      if (kTwo && (e & 1)) {
        qb = fmaf(kp[e], a[e], qb);
        amaxb = fmaxf(amaxb, fabsf(a[e]));
#pragma unroll
        for (int d = 0; d < D2b; ++d) fv2b[d] = fma2(pack2(S[2 * d], S[2 * d + 1]), aa, fv2b[d]);
      } else {
        q = fmaf(kp[e], a[e], q);
        amax = fmaxf(amax, fabsf(a[e]));
#pragma unroll
        for (int d = 0; d < D2; ++d) fv2[d] = fma2(pack2(S[2 * d], S[2 * d + 1]), aa, fv2[d]);
      }
    }
  });
  q += qb;
  amax = fmaxf(amax, amaxb);
#pragma unroll
  for (int d = 0; d < D2; ++d) {
    if (kTwo) {
      fm2[d] = add2(fm2[d], fm2b[kTwo ? d : 0]);
      fv2[d] = add2(fv2[d], fv2b[kTwo ? d : 0]);
    }
    float m0, m1, v0, v1;
    unpack2(fm2[d], m0, m1);
    unpack2(fv2[d], v0, v1);
    fm[2 * d] = m0; fv[2 * d] = v0;
    if (2 * d + 1 < DOUT) { fm[2 * d + 1 < DOUT ? 2 * d + 1 : 0] = m1; fv[2 * d + 1 < DOUT ? 2 * d + 1 : 0] = v1; }
  }
  if (NG > 1) {   // partial sums of the groups, added in group order by every thread of the particle
    *c.xslot(g0, 1) = q;
    *c.xslot(g0, 2) = amax;
#pragma unroll
    for (int d = 0; d < DOUT; ++d) { *c.xslot(g0, 3 + d) = fm[d]; *c.xslot(g0, 3 + DOUT + d) = fv[d]; }
    __syncthreads();
    q = 0.f; amax = 0.f;
#pragma unroll
    for (int d = 0; d < DOUT; ++d) { fm[d] = 0.f; fv[d] = 0.f; }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      q += *c.xslot(g, 1);
      amax = fmaxf(amax, *c.xslot(g, 2));
#pragma unroll
      for (int d = 0; d < DOUT; ++d) { fm[d] += *c.xslot(g, 3 + d); fv[d] += *c.xslot(g, 3 + DOUT + d); }
    }
  }
  // q, fv, amax were formed from the normalised k'' and a'' = P' k''; undo the per-particle scale
  const float s4 = c.sig2 * c.sig2, ps = c.pscale, k2 = kscale * kscale;
#pragma unroll
  for (int d = 0; d < DOUT; ++d) {
    fm[d] *= c.sig2 * kscale;
    fv[d] = gp_var_clamp(c.sig2 - ps * s4 * k2 * q + ps * ps * s4 * k2 * fv[d]);
  }
  tc_fence_before();   // order this step's tcgen05.ld before the next barrier / MMA
}

// Reverse of one GP evaluation (SURVEY 8a note 4) on the tile; follows gp_forward_tc of the same step
// (K1/K2 and the accumulator D1 still hold k' and a').  gm/gv: adjoints of (fmean, fvar).
// NG > 1: Lacc and sw receive this thread's chunks only (they are summed over all threads at the end of the
// kernel); sG is added by group 0; xinb is the full sum in every thread.
template <class Ctx, int DIN, int DOUT, int NEED>
__device__ __forceinline__ void gp_reverse_tc(Ctx &c, const float (&xt)[(DIN + 3) / 4 * 4], const float (&gm)[DOUT],
                                              const float (&gv)[DOUT], float amax, float kscale, bool live,
                                              const TcOut &o, float (&xinb)[NEED], float (&Lacc)[DIN], float &sw,
                                              float &sG) {
  constexpr int DINP = (DIN + 3) / 4 * 4, DOUTP = (DOUT + 3) / 4 * 4;
  constexpr int MC = Ctx::MC, NG = Ctx::NG;
  const int t = c.lane_id(), g0 = c.grp(), M = MC ? MC : c.M, MP = MC ? (MC + 15) / 16 * 16 : c.MP;
  (void)M;
  const float ps = c.pscale, sig2 = c.sig2;
  const float ascale = ps * sig2 * kscale;      // a = ascale * a'' (a'' = P' k'' is what D1 holds)
  float Gs = 0.f, cbound = 0.f;
#pragma unroll
  for (int d = 0; d < DOUT; ++d) { Gs += gv[d]; cbound += fabsf(gv[d]) * c.smax[d]; }
  if (g0 == 0) sG += Gs;
  // per-particle power-of-two scale so that |b''| = |a'' c| * 2^-e <= 1  (amax is of the normalised a'')
  int e2 = 0;
  frexpf(fmaxf(amax * cbound, 1e-30f), &e2);
  const float bsc = ldexpf(1.f, -e2), binv = ldexpf(1.f, e2);
  const uint32_t trow1 = c.tmem + ((uint32_t)(t & ~31) << 16);
  // k_bar needs 2 P b - 2 G a = 2 pbs (P' b'' + acoef a''): each thread overwrites its a'' row with acoef a'' once it
  // has formed b'' from it, and the second contraction accumulates onto that -- one accumulator serves both products
  const float acoef = -Gs * bsc / ps;
  // Per-particle factors are folded into the small operands and into loop constants instead of being applied per
  // inducing row: the K operand holds k'' (k' = kscale k''), the A2 operand a''^2 (a = ascale a'').
  // (pre-scaled copies of gm / gv only for few output dims: at DOUT = 14 the 28 extra live registers cost more than
  // the two multiplies per row they save -- measured +6 % on the Sarcos shape)
  constexpr bool kPre = DOUT <= 4;
  float gvs[kPre ? DOUT : 1], gms[kPre ? DOUT : 1];
  if constexpr (kPre) {
#pragma unroll
    for (int d = 0; d < DOUT; ++d) { gvs[d] = gv[d] * bsc; gms[d] = gm[d] * (sig2 * kscale); }
  }
  if (live && g0 == 0) {
    float gmk[DOUT], gva[DOUT];
#pragma unroll
    for (int d = 0; d < DOUT; ++d) { gmk[d] = gm[d] * kscale; gva[d] = gv[d] * (ascale * ascale); }
    o.template put_vec<DOUT>(o.bGm, gmk);
    o.template put_vec<DOUT>(o.bGv, gva);
    float x1[DIN + 1];
#pragma unroll
    for (int j = 0; j < DIN; ++j) x1[j] = xt[j];
    x1[DIN] = 1.f;
    o.template put_vec<DIN + 1>(o.bX1, x1);
  }
  // ---- b'' = a' (S gv) 2^-e -> fp16 split rows of B ----
  uint32_t ra1[16];
  if (g0 < MP / 16) tmem_ld16_issue(trow1 + g0 * 16, ra1);
  for_chunks<MC, NG>(g0, MP / 16, [&](const int cc, auto rows_tag) {
    constexpr int ROWS = decltype(rows_tag)::value;
    float a[16];
    tmem_ld_wait(ra1);
#pragma unroll
    for (int e = 0; e < 16; ++e) a[e] = __uint_as_float(ra1[e]);
    if (cc + NG < MP / 16) tmem_ld16_issue(trow1 + (cc + NG) * 16, ra1);
    float bv[16], a2[16];
#pragma unroll
    for (int e = 0; e < ROWS; ++e) {
      const int m = cc * 16 + e;
      float S[DOUTP];
      ld_row<DOUTP>(c.Sm + m * DOUTP, S);
      float cm = 0.f;
#pragma unroll
      for (int d = 0; d < DOUT; ++d) cm = fmaf(S[d], kPre ? gvs[kPre ? d : 0] : gv[d], cm);
      bv[e] = kPre ? a[e] * cm : a[e] * (cm * bsc);
      a2[e] = a[e] * a[e];
      a[e] *= acoef;
    }
    tmem_st16(trow1 + cc * 16, a);       // (columns of padded rows are never read back)
#pragma unroll
    for (int hh = 0; hh < ROWS / 8; ++hh) {
      float g8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) g8[e] = bv[8 * hh + e];
      tc_put_row8(c, c.B1, c.B2, t, 2 * cc + hh, g8);
      if (live && 2 * cc + hh < o.MB) {
#pragma unroll
        for (int e = 0; e < 8; ++e) g8[e] = a2[8 * hh + e];
        o.put8(o.bA2 + 2 * cc + hh, g8);
      }
    }
  });
  tmem_st_wait();
  tc_fence_before();
  uint4 kq[4];   // raw hi/lo segments of two row-blocks of k' (one 16-row chunk), fetched one chunk ahead
  o.get8_raw(o.bK + 2 * g0, live, kq);
  // ---- D <- acoef a'' + B P' ----
  tc_contract(c, c.B1, c.B2, c.tmem, 1);
  // ---- k_bar = alpha gm + 2 P b - 2 G a ; w = k_bar k ; a_bar = 2 b - G k ----
  const float pbs = ps * ascale * binv;         // (P b)_m = pbs * (P' b'')_m
  const float bs = ascale * binv;               // b_m = bs * b''_m
  const float kfac = sig2 * kscale;             // k_m = kfac * k''_m
  const float c_pb = 2.f * pbs * kfac, c_bb = 2.f * bs * kscale, c_k = -Gs * kfac * kscale;
  constexpr int J2 = (DIN + 1) / 2, N2 = (NEED + 1) / 2;
  constexpr bool kTwo = NG > 1 && DIN <= 8;   // two chains (even / odd rows): latency variant only, as in gp_forward_tc
  constexpr int N2b = kTwo ? N2 : 1, J2b = kTwo ? J2 : 1;
  unsigned long long x2[J2], xs2[N2], L2[J2], xs2b[N2b], L2b[J2b];   // packed over input-dim pairs (2j, 2j+1)
  float swb = 0.f;
#pragma unroll
  for (int j = 0; j < J2; ++j) { x2[j] = pack2(xt[2 * j], xt[2 * j + 1]); L2[j] = 0ull; }
#pragma unroll
  for (int j = 0; j < J2b; ++j) L2b[j] = 0ull;
#pragma unroll
  for (int j = 0; j < N2; ++j) xs2[j] = 0ull;
#pragma unroll
  for (int j = 0; j < N2b; ++j) xs2b[j] = 0ull;
  for_chunks<MC, NG>(g0, MP / 16, [&](const int cc, auto rows_tag) {
    constexpr int ROWS = decltype(rows_tag)::value;
    float pb[16], kp[16], bb[16];
    {
      uint32_t rp[16];
      tmem_ld16_issue(trow1 + cc * 16, rp);
      // k'' of this chunk was fetched one iteration ahead; fetch the next chunk's now
      const uint4 c0 = kq[0], c1 = kq[1], c2 = kq[2], c3 = kq[3];
      if (cc + NG < MP / 16) o.get8_raw(o.bK + 2 * (cc + NG), live, kq);
      tc_get_row16<ROWS>(c, c.B1, c.B2, t, cc, bb);
      {
        float k0[8], k1[8];
        TcOut::unpack8(c0, c1, k0);
        if (ROWS > 8) TcOut::unpack8(c2, c3, k1);
#pragma unroll
        for (int e = 0; e < 8; ++e) { kp[e] = k0[e]; if (ROWS > 8) kp[8 + e] = k1[e]; }
      }
      tmem_ld_wait(rp);
#pragma unroll
      for (int e = 0; e < 16; ++e) pb[e] = __uint_as_float(rp[e]);
    }
    float wv[16], abv[16];
#pragma unroll
    for (int e = 0; e < ROWS; ++e) {
      const int m = cc * 16 + e;
      float al[DOUTP];
      ld_row<DOUTP>(c.al + m * DOUTP, al);
      float kb;                                 // kfac * k_bar
      if constexpr (kPre) {
        kb = c_pb * pb[e];
#pragma unroll
        for (int d = 0; d < DOUT; ++d) kb = fmaf(al[d], gms[d], kb);
      } else {
        float dot = 0.f;
#pragma unroll
        for (int d = 0; d < DOUT; ++d) dot = fmaf(al[d], gm[d], dot);
        kb = fmaf(c_pb, pb[e], kfac * dot);
      }
      const float w = kb * kp[e];
      if (kTwo && (e & 1)) swb += w; else sw += w;
      float z[DINP];
      ld_row<DINP>(c.Zt + m * DINP, z);
      const unsigned long long ww = pack2(w, w);
#pragma unroll
      for (int j = 0; j < J2; ++j) {
        const unsigned long long dl = add2(x2[j], pack2(z[2 * j], z[2 * j + 1]));
        const unsigned long long wd = mul2(dl, ww);
        if (kTwo && (e & 1)) {
          if (j < N2) xs2b[(kTwo && j < N2) ? j : 0] = add2(xs2b[(kTwo && j < N2) ? j : 0], wd);
          L2b[kTwo ? j : 0] = fma2(wd, dl, L2b[kTwo ? j : 0]);
        } else {
          if (j < N2) xs2[j < N2 ? j : 0] = add2(xs2[j < N2 ? j : 0], wd);
          L2[j] = fma2(wd, dl, L2[j]);
        }
      }
      wv[e] = w;
      abv[e] = fmaf(c_k, kp[e], c_bb * bb[e]);  // kscale * a_bar (its partner operand holds k'')
    }
    if (live) {
#pragma unroll
      for (int hh = 0; hh < ROWS / 8; ++hh) {
        if (2 * cc + hh < o.MB) {
          float g8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) g8[e] = wv[8 * hh + e];
          o.put8(o.bW + 2 * cc + hh, g8);
#pragma unroll
          for (int e = 0; e < 8; ++e) g8[e] = abv[8 * hh + e];
          o.put8(o.bAb + 2 * cc + hh, g8);
        }
      }
    }
  });
  sw += swb;
#pragma unroll
  for (int j = 0; j < N2b; ++j) if (kTwo) xs2[j] = add2(xs2[j], xs2b[j]);
#pragma unroll
  for (int j = 0; j < J2; ++j) {
    if (kTwo) L2[j] = add2(L2[j], L2b[kTwo ? j : 0]);
    float l0, l1;
    unpack2(L2[j], l0, l1);
    Lacc[2 * j] += l0;
    if (2 * j + 1 < DIN) Lacc[2 * j + 1 < DIN ? 2 * j + 1 : 0] += l1;
  }
  {
    float xs[2 * N2];
#pragma unroll
    for (int j = 0; j < N2; ++j) unpack2(xs2[j], xs[2 * j], xs[2 * j + 1]);
    if (NG > 1) {   // x_bar heads the next step's dependency chain: every thread of the particle needs the full sum
      constexpr int X0 = 3 + 2 * DOUT;
#pragma unroll
      for (int j = 0; j < NEED; ++j) *c.xslot(g0, X0 + j) = xs[j];
      __syncthreads();
#pragma unroll
      for (int j = 0; j < NEED; ++j) {
        float v = 0.f;
#pragma unroll
        for (int g = 0; g < NG; ++g) v += *c.xslot(g, X0 + j);
        xs[j] = v;
      }
    }
    float il[DINP];
    ld_row<DINP>(c.il, il);
#pragma unroll
    for (int j = 0; j < NEED; ++j) xinb[j] = -xs[j] * il[j];
  }
  tc_fence_before();
}

// Sum COUNT per-thread floats over one particle tile (128 threads x NG groups); the tile's thread 0 writes
// out[0..COUNT) unless out is null.  scratch: 4 * NG * COUNT floats owned by the tile.
template <int COUNT, int NT, int NG = 1>
__device__ __forceinline__ void tile_sum_store(const float (&vals)[COUNT], float *scratch, float *out, int tile, int tl) {
  constexpr int NW = 4 * NG;
  const int lane = tl & 31, warp = (NG == 1) ? (tl >> 5) : (int)(threadIdx.x >> 5);
  if (NT == 1) __syncthreads(); else tile_sync(tile);
#pragma unroll
  for (int i = 0; i < COUNT; ++i) {
    float v = vals[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) scratch[i * NW + warp] = v;
  }
  if (NT == 1) __syncthreads(); else tile_sync(tile);
  if (tl == 0 && (NG == 1 || threadIdx.x == 0) && out != nullptr) {
    for (int i = 0; i < COUNT; ++i) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) sum += scratch[i * NW + w];
      out[i] = sum;
    }
  }
}

// =====================================================================================
template <int DX, int DU, int DY, int MC, int NT, int NG = 1>
__global__ void __launch_bounds__(NT * NG * kTcThreads, (NT == 1 && NG == 1) ? 2 : 1) bm_forward_tc_kernel(Dims D, ChainTable chains, GpDev gp,
                                                                   const float *__restrict__ vxg,
                                                                   const float *__restrict__ u,
                                                                   const float *__restrict__ y,
                                                                   const float *__restrict__ eps_b,
                                                                   const float *__restrict__ z_b, Workspace ws,
                                                                   float *__restrict__ part_out) {
  constexpr int DH = DX - DY, DIN = DX + DU;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ float scratch[4 * NT * NG + 8];
  __shared__ float vx[16];
  using Ctx = TcCtx<DIN, DH, false, MC, NT, NG>;
  Ctx c;
  c.init(smem_raw, gp, D.M, scratch);
  if (threadIdx.x < DX) vx[threadIdx.x] = vxg[threadIdx.x];
  __syncthreads();

  const Chain ch = chains.c[blockIdx.y];
  const int ntiles = ceil_div(D.n_local, kTcThreads), gtile = blockIdx.x * NT + c.tile_id();
  const int nl = gtile * kTcThreads + c.lane_id();
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
    float xin[DIN], fm[DH], fv[DH];
#pragma unroll
    for (int j = 0; j < DH; ++j) xin[j] = h[j];
#pragma unroll
    for (int j = 0; j < DU; ++j) xin[DH + j] = ub[t * DU + j];
#pragma unroll
    for (int j = 0; j < DY; ++j) xin[DH + DU + j] = yb[t * DY + j];
    const float e = eps_b[((size_t)ch.run * D.T + t) * D.n_local + nr];
    float xt[Ctx::DINP], amax, kscale;
    gp_forward_tc<Ctx, DIN, DH>(c, xin, xt, fm, fv, nullptr, amax, kscale);
    if (ws.FVb != nullptr && live && c.grp() == 0) {
      float *Fp = ws.FVb + (((size_t)ch.run * D.T + t) * (2 * DH + 1)) * np + nl;
#pragma unroll
      for (int j = 0; j < DH; ++j) { Fp[(size_t)j * np] = fm[j]; Fp[(size_t)(DH + j) * np] = fv[j]; }
      Fp[(size_t)(2 * DH) * np] = amax;
    }
    const bool write = writer_run(t, D.R) == ch.run;
#pragma unroll
    for (int j = 0; j < DH; ++j) {
      const float f = fv[j] + vx[j];
      h[j] = fm[j] + h[j] + e * sqrtf(f);
      if (write) ent += 0.5f * (kLog2PiE + logf(f));
    }
    if (live && c.grp() == 0) {
      float *Hp = ws.H + (((size_t)ch.run * D.T + t) * DH) * np + nl;
#pragma unroll
      for (int j = 0; j < DH; ++j) Hp[j * np] = h[j];
    }
  }
  c.release();
  const float v[1] = {(live && c.grp() == 0) ? ent : 0.f};
  tile_sum_store<1, NT, NG>(v, scratch + 4 * c.tile_id(), gtile < ntiles ? part_out + ((size_t)blockIdx.y * ntiles + gtile) : nullptr,
                        c.tile_id(), c.lane_id());
}

template <int DX, int DU, int DY, int MC, int NT, int NG = 1>
__global__ void __launch_bounds__(NT * NG * kTcThreads, (NT == 1 && NG == 1) ? 2 : 1) fw_forward_tc_kernel(Dims D, GpDev gp, const float *__restrict__ vxg,
                                                                   const float *__restrict__ vyg,
                                                                   const float *__restrict__ u,
                                                                   const float *__restrict__ y,
                                                                   const float *__restrict__ eps_f, Workspace ws,
                                                                   float *__restrict__ part_out) {
  constexpr int DH = DX - DY, DIN = DX + DU;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ float scratch[NT * NG * 4 * (DY + 1) + 8];
  __shared__ float vx[16], vy[16];
  using Ctx = TcCtx<DIN, DX, false, MC, NT, NG>;
  Ctx c;
  c.init(smem_raw, gp, D.M, scratch);
  if (threadIdx.x < DX) { vx[threadIdx.x] = vxg[threadIdx.x]; vy[threadIdx.x] = vyg[threadIdx.x]; }
  __syncthreads();

  const int ntiles = ceil_div(D.n_local, kTcThreads), gtile = blockIdx.x * NT + c.tile_id();
  const int nl = gtile * kTcThreads + c.lane_id();
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
    if (live && c.grp() == 0) {
      float *Xp = ws.X + ((size_t)t * DX) * np + nl;
#pragma unroll
      for (int j = 0; j < DX; ++j) Xp[j * np] = x[j];
#pragma unroll
      for (int j = 0; j < DY; ++j) { const float d = yb[t * DY + j] - x[j]; sse[j] = fmaf(d, d, sse[j]); }
    }
    if (t == D.T - 1) break;
    float xin[DIN], fm[DX], fv[DX], yt[DX], xn[DX];
#pragma unroll
    for (int j = 0; j < DX; ++j) xin[j] = x[j];
#pragma unroll
    for (int j = 0; j < DU; ++j) xin[DX + j] = ub[t * DU + j];
    load_ytil(t + 1, yt);
    const float e = eps_f[(size_t)t * D.n_local + nr];
    float xt[Ctx::DINP], amax, kscale;
    gp_forward_tc<Ctx, DIN, DX>(c, xin, xt, fm, fv, nullptr, amax, kscale);
    if (ws.FVf != nullptr && live && c.grp() == 0) {
      float *Fp = ws.FVf + ((size_t)t * (2 * DX + 1)) * np + nl;
#pragma unroll
      for (int j = 0; j < DX; ++j) { Fp[(size_t)j * np] = fm[j]; Fp[(size_t)(DX + j) * np] = fv[j]; }
      Fp[(size_t)(2 * DX) * np] = amax;
    }
    const bool do_cond = D.condition || (t < D.R - 1);
    fw_step<DX>(x, fm, fv, yt, e, vx, vy, D.kap, do_cond, D.ncond, xn, kl);
#pragma unroll
    for (int j = 0; j < DX; ++j) x[j] = xn[j];
  }
  c.release();
  sse[DY] = kl;
  if (!live || c.grp() != 0) {
#pragma unroll
    for (int j = 0; j <= DY; ++j) sse[j] = 0.f;
  }
  tile_sum_store<DY + 1, NT, NG>(sse, scratch + 4 * (DY + 1) * c.tile_id(), gtile < ntiles ? part_out + (size_t)gtile * (DY + 1) : nullptr,
                             c.tile_id(), c.lane_id());
}

__device__ __forceinline__ TcOut tc_out_at(const TcMats &m, size_t col) {
  TcOut o;
  o.col = m.blk + (col / kOT) * m.tile_bytes() + (size_t)(col % kOT) * 16;
  o.RB = m.RB; o.MB = m.MB; o.DB = m.DB; o.XB = m.XB;
  o.bAb = m.bAb; o.bK = m.bK; o.bA2 = m.bA2; o.bW = m.bW; o.bGm = m.bGm; o.bGv = m.bGv; o.bX1 = m.bX1;
  return o;
}

// =====================================================================================
// Reverse of the forward rollout on tcgen05: one CTA = 128 particles, T-1 steps.
// Per-CTA output: scalar sums [L_j | sum w | sum G | var_x_bar | var_y_bar] at spart[cta].
// =====================================================================================
template <int DX, int DU, int DY, int MC, int NT, int NG = 1>
__global__ void __launch_bounds__(NT * NG * kTcThreads, (NT == 1 && NG == 1) ? 2 : 1) fw_reverse_tc_kernel(Dims D, GpDev gp, const float *__restrict__ vxg,
                                                                   const float *__restrict__ vyg,
                                                                   const float *__restrict__ u,
                                                                   const float *__restrict__ y,
                                                                   const float *__restrict__ eps_f, float w_ll,
                                                                   float w_kl, Workspace ws, TcMats mats,
                                                                   TimeWin win, float *__restrict__ spart, int nsc) {
  constexpr int DH = DX - DY, DIN = DX + DU;
  using Ctx = TcCtx<DIN, DX, true, MC, NT, NG>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ float scratch[NT * NG * 4 * (DIN + 2 + 2 * DX)];
  __shared__ float vx[16], vy[16];
  Ctx c;
  c.init(smem_raw, gp, D.M, scratch);
  if (threadIdx.x < DX) { vx[threadIdx.x] = vxg[threadIdx.x]; vy[threadIdx.x] = vyg[threadIdx.x]; }
  __syncthreads();

  const int ntiles = ceil_div(D.n_local, kTcThreads), gtile = blockIdx.x * NT + c.tile_id();
  const int nl = gtile * kTcThreads + c.lane_id();
  const bool live = nl < D.n_local;
  const int nr = live ? nl : 0;
  const int b = (D.n_offset + nr) / D.S;
  const float *ub = u + (size_t)b * D.T * DU;
  const float *yb = y + (size_t)b * D.T * DY;
  const size_t np = ws.npad;

  float Lacc[DIN], sw = 0.f, sG = 0.f, vxacc[DX], vyacc[DX];
#pragma unroll
  for (int j = 0; j < DIN; ++j) Lacc[j] = 0.f;
#pragma unroll
  for (int j = 0; j < DX; ++j) { vxacc[j] = 0.f; vyacc[j] = 0.f; }

  float xb[DX];
  if (win.first) {
    const float *Xp = ws.X + ((size_t)(D.T - 1) * DX) * np + nr;
#pragma unroll
    for (int j = 0; j < DX; ++j)
      xb[j] = (j < DY && live) ? w_ll * (yb[(D.T - 1) * DY + (j < DY ? j : 0)] - Xp[j * np]) / vy[j] : 0.f;
  } else {   // adjoint of x_{t_hi+1} left by the previous (later-in-time) window
#pragma unroll
    for (int j = 0; j < DX; ++j) xb[j] = live ? ws.carry_f[j * np + nr] : 0.f;
  }
  // fetched one time step ahead for few state dims; for many (2 DX + 1 = 29 registers held across a whole evaluation
  // at DX = 14) the kernel spills instead, and the values are loaded where they are used
  constexpr bool kAhead = DX <= 8;
  float xnext[DX];   // x_t of the coming iteration, fetched one step ahead (it heads the step's dependency chain)
  if (kAhead) {
    const float *Xp = ws.X + ((size_t)win.t_hi * DX) * np + nr;
#pragma unroll
    for (int j = 0; j < DX; ++j) xnext[j] = Xp[j * np];
  }
  const bool saved = ws.FVf != nullptr;      // (fmean, fvar, amax) of every step left by fw_forward_tc
  float fnext[2 * DX + 1];
  auto load_saved = [&](int t) {
    const float *Fp = ws.FVf + ((size_t)t * (2 * DX + 1)) * np + nr;
#pragma unroll
    for (int j = 0; j < 2 * DX + 1; ++j) fnext[j] = Fp[(size_t)j * np];
  };
  if (saved && kAhead) load_saved(win.t_hi);
#pragma unroll 1
  for (int t = win.t_hi; t >= win.t_lo; --t) {
    float x[DX], xin[DIN], xt[Ctx::DINP], fm[DX], fv[DX], yt[DX], amax, kscale;
    if (!kAhead) {
      const float *Xp = ws.X + ((size_t)t * DX) * np + nr;
#pragma unroll
      for (int j = 0; j < DX; ++j) xnext[j] = Xp[j * np];
    }
#pragma unroll
    for (int j = 0; j < DX; ++j) { x[j] = xnext[j]; xin[j] = x[j]; }
    if (kAhead && t > win.t_lo) {
      const float *Xp = ws.X + ((size_t)(t - 1) * DX) * np + nr;
#pragma unroll
      for (int j = 0; j < DX; ++j) xnext[j] = Xp[j * np];
    }
#pragma unroll
    for (int j = 0; j < DU; ++j) xin[DX + j] = ub[t * DU + j];
#pragma unroll
    for (int j = 0; j < DY; ++j) yt[j] = yb[(t + 1) * DY + j];
    {
      const float *Hp = ws.H + (((size_t)writer_run(t + 1, D.R) * D.T + (t + 1)) * DH) * np + nr;
#pragma unroll
      for (int j = 0; j < DH; ++j) yt[DY + j] = D.half ? 0.f : Hp[j * np];
    }
    const float e = eps_f[(size_t)t * D.n_local + nr];
    const TcOut o = tc_out_at(mats, (size_t)(t - win.t_lo) * D.n_local + nr);
    if (saved) {
      if (!kAhead) load_saved(t);
#pragma unroll
      for (int j = 0; j < DX; ++j) { fm[j] = fnext[j]; fv[j] = fnext[DX + j]; }
      amax = fnext[2 * DX];
      if (kAhead && t > win.t_lo) load_saved(t - 1);
    }
    gp_forward_tc<Ctx, DIN, DX>(c, xin, xt, fm, fv, live ? &o : nullptr, amax, kscale, saved);
    const bool do_cond = D.condition || (t < D.R - 1);
    float fmb[DX], fvb[DX], ytb[DX];
    fw_step_adjoint<DX>(x, fm, fv, yt, e, vx, vy, D.kap, do_cond, D.ncond, w_kl, xb, fmb, fvb, ytb, vxacc, vyacc,
                        live && c.grp() == 0);
    if (live) {
      if (c.grp() == 0) {
        float *Yp = ws.Yb + ((size_t)(t + 1) * DH) * np + nl;
#pragma unroll
        for (int j = 0; j < DH; ++j) Yp[j * np] = ytb[DY + j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < DX; ++j) { fmb[j] = 0.f; fvb[j] = 0.f; }
    }
    float xinb[DX];
    gp_reverse_tc<Ctx, DIN, DX, DX>(c, xt, fmb, fvb, amax, kscale, live, o, xinb, Lacc, sw, sG);
#pragma unroll
    for (int j = 0; j < DX; ++j) {
      float lg = 0.f;
      if (j < DY && live) lg = w_ll * (yb[t * DY + (j < DY ? j : 0)] - x[j]) / vy[j];
      xb[j] = xinb[j] + fmb[j] + lg;
    }
  }
  if (c.grp() != 0) {
    // only group 0 stores per-particle results
  } else if (!win.last) {
    if (live) {
#pragma unroll
      for (int j = 0; j < DX; ++j) ws.carry_f[j * np + nl] = xb[j];
    }
  } else if (live && D.half) {   // CBFSSMHALF: adjoint of the recognition model's x_0 (all dims)
#pragma unroll
    for (int j = 0; j < DX; ++j) ws.x0b[j * np + nl] = xb[j];
  } else if (live) {
    float *Yp = ws.Yb + nl;
#pragma unroll
    for (int j = 0; j < DH; ++j) Yp[j * np] = xb[DY + j];
  }
  c.release();
  float sc[DIN + 2 + 2 * DX];
#pragma unroll
  for (int j = 0; j < DIN; ++j) sc[j] = Lacc[j];
  sc[DIN] = sw; sc[DIN + 1] = sG;
#pragma unroll
  for (int j = 0; j < DX; ++j) { sc[DIN + 2 + j] = vxacc[j]; sc[DIN + 2 + DX + j] = vyacc[j]; }
  tile_sum_store<DIN + 2 + 2 * DX, NT, NG>(sc, scratch + 4 * (DIN + 2 + 2 * DX) * c.tile_id(),
                                       gtile < ntiles ? spart + (size_t)gtile * nsc : nullptr, c.tile_id(), c.lane_id());
}

template <int DX, int DU, int DY, int MC, int NT, int NG = 1>
__global__ void __launch_bounds__(NT * NG * kTcThreads, (NT == 1 && NG == 1) ? 2 : 1) bm_reverse_tc_kernel(Dims D, ChainTable chains, GpDev gp,
                                                                   const float *__restrict__ vxg,
                                                                   const float *__restrict__ u,
                                                                   const float *__restrict__ y,
                                                                   const float *__restrict__ eps_b,
                                                                   const float *__restrict__ z_b, float w_en,
                                                                   Workspace ws, TcMats mats,
                                                                   float *__restrict__ spart, int nsc) {
  constexpr int DH = DX - DY, DIN = DX + DU;
  using Ctx = TcCtx<DIN, DH, true, MC, NT, NG>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ float scratch[NT * NG * 4 * (DIN + 2 + 2 * DX)];
  __shared__ float vx[16];
  Ctx c;
  c.init(smem_raw, gp, D.M, scratch);
  if (threadIdx.x < DX) vx[threadIdx.x] = vxg[threadIdx.x];
  __syncthreads();

  const Chain ch = chains.c[blockIdx.y];
  const int ntiles = ceil_div(D.n_local, kTcThreads), gtile = blockIdx.x * NT + c.tile_id();
  const int nl = gtile * kTcThreads + c.lane_id();
  const bool live = nl < D.n_local;
  const int nr = live ? nl : 0;
  const int b = (D.n_offset + nr) / D.S;
  const float *ub = u + (size_t)b * D.T * DU;
  const float *yb = y + (size_t)b * D.T * DY;
  const size_t np = ws.npad;

  float Lacc[DIN], sw = 0.f, sG = 0.f, vxacc[DH];
#pragma unroll
  for (int j = 0; j < DIN; ++j) Lacc[j] = 0.f;
#pragma unroll
  for (int j = 0; j < DH; ++j) vxacc[j] = 0.f;

  float hb[DH];
#pragma unroll
  for (int j = 0; j < DH; ++j) hb[j] = ((ch.carry & 1) && live) ? ws.carry_b[((size_t)ch.id * DH + j) * np + nr] : 0.f;
  // the message state entering step t: the chain's initial value at its first step, else the output of step t+1
  auto load_hidden = [&](int t, float(&hv)[DH]) {
    if (t == ch.t_top) {
      const float z = (ch.init == 1) ? z_b[((size_t)ch.run * D.T + t) * D.n_local + nr] : 0.f;
#pragma unroll
      for (int j = 0; j < DH; ++j) hv[j] = z;
    } else {
      const float *Hp = ws.H + (((size_t)ch.run * D.T + (t + 1)) * DH) * np + nr;
#pragma unroll
      for (int j = 0; j < DH; ++j) hv[j] = Hp[j * np];
    }
  };
  float hnext[DH];   // fetched one step ahead (it heads the step's dependency chain)
  float ynext[DH];   // adjoint of y2[t], likewise: its load sat directly in front of its use (4 % of the samples)
  auto load_ybar = [&](int t, float(&yv)[DH]) {
    const float *Yq = ws.Yb + ((size_t)t * DH) * np + nr;
#pragma unroll
    for (int j = 0; j < DH; ++j) yv[j] = Yq[j * np];
  };
  const bool saved = ws.FVb != nullptr;      // (fmean, fvar, amax) of every step left by bm_forward_tc
  constexpr bool kAhead = DX <= 8;           // as in fw_reverse_tc_kernel
  float fnext[2 * DH + 1];
  auto load_saved = [&](int t) {
    const float *Fp = ws.FVb + (((size_t)ch.run * D.T + t) * (2 * DH + 1)) * np + nr;
#pragma unroll
    for (int j = 0; j < 2 * DH + 1; ++j) fnext[j] = Fp[(size_t)j * np];
  };
  load_hidden(ch.t_lo, hnext);
  load_ybar(ch.t_lo, ynext);
  if (saved && kAhead) load_saved(ch.t_lo);
#pragma unroll 1
  for (int t = ch.t_lo; t <= ch.t_hi; ++t) {
    float hid[DH], ybar[DH], xin[DIN], xt[Ctx::DINP], fm[DH], fv[DH], amax, kscale;
#pragma unroll
    for (int j = 0; j < DH; ++j) { hid[j] = hnext[j]; ybar[j] = ynext[j]; }
    if (t < ch.t_hi) { load_hidden(t + 1, hnext); load_ybar(t + 1, ynext); }
#pragma unroll
    for (int j = 0; j < DH; ++j) xin[j] = hid[j];
#pragma unroll
    for (int j = 0; j < DU; ++j) xin[DH + j] = ub[t * DU + j];
#pragma unroll
    for (int j = 0; j < DY; ++j) xin[DH + DU + j] = yb[t * DY + j];
    const float e = eps_b[((size_t)ch.run * D.T + t) * D.n_local + nr];
    const TcOut o = tc_out_at(mats, ((size_t)ch.col0 + (t - ch.t_lo)) * D.n_local + nr);
    if (saved) {
      if (!kAhead) load_saved(t);
#pragma unroll
      for (int j = 0; j < DH; ++j) { fm[j] = fnext[j]; fv[j] = fnext[DH + j]; }
      amax = fnext[2 * DH];
      if (kAhead && t < ch.t_hi) load_saved(t + 1);
    }
    gp_forward_tc<Ctx, DIN, DH>(c, xin, xt, fm, fv, live ? &o : nullptr, amax, kscale, saved);
    const bool write = writer_run(t, D.R) == ch.run;
    float ob[DH], fvb[DH];
#pragma unroll
    for (int j = 0; j < DH; ++j) {
      const float f = fv[j] + vx[j];
      float ov = hb[j], fb = 0.f;
      if (write) {
        ov += ybar[j];
        fb = w_en * 0.5f / f;
      }
      fb += ov * e * 0.5f * rsqrtf(f);
      if (!live) { ov = 0.f; fb = 0.f; }
      ob[j] = ov; fvb[j] = fb;
      if (c.grp() == 0) vxacc[j] += fb;
    }
    float xinb[DH];
    gp_reverse_tc<Ctx, DIN, DH, DH>(c, xt, ob, fvb, amax, kscale, live, o, xinb, Lacc, sw, sG);
#pragma unroll
    for (int j = 0; j < DH; ++j) hb[j] = xinb[j] + ob[j];
  }
  if ((ch.carry & 2) && live && c.grp() == 0) {
#pragma unroll
    for (int j = 0; j < DH; ++j) ws.carry_b[((size_t)ch.id * DH + j) * np + nl] = hb[j];
  }
  c.release();
  float sc[DIN + 2 + 2 * DX];
#pragma unroll
  for (int j = 0; j < DIN; ++j) sc[j] = Lacc[j];
  sc[DIN] = sw; sc[DIN + 1] = sG;
#pragma unroll
  for (int j = 0; j < DX; ++j) { sc[DIN + 2 + j] = (j < DH) ? vxacc[j < DH ? j : 0] : 0.f; sc[DIN + 2 + DX + j] = 0.f; }
  tile_sum_store<DIN + 2 + 2 * DX, NT, NG>(sc, scratch + 4 * (DIN + 2 + 2 * DX) * c.tile_id(),
                                       gtile < ntiles ? spart + ((size_t)blockIdx.y * ntiles + gtile) * nsc : nullptr,
                                       c.tile_id(), c.lane_id());
}

template <int DX, int DU, int DY>
struct LaunchTc {
  static constexpr int DH = DX - DY, DIN = DX + DU;
  // dims with a compile-time M = 100 instantiation (run/template.py, SpringNonlinear, Sarcos shapes)
  static constexpr bool kHas100 = (DX == 4 && (DU == 1 || DU == 2)) || DX == 14;
  // dims whose tables can push a one-tile CTA past two CTAs per SM (dx >= 8 at M = 128: 64 KB of P terms + 29 KB of
  // tables): the two-tile kernels are compiled too, pick() takes them when the occupancy query says one CTA per SM
  static constexpr bool kDual = DX >= 8;
  // dims with the latency variant (kSplitG threads per particle, TcCtx NG): used when a launch has fewer CTAs than
  // the GPU has SMs, i.e. when the serial chain of a time step, not throughput, sets the kernel's duration
  static constexpr bool kSplit = DX <= 4;
  static constexpr int kSplitG = 2;
  // small dims at M <= 112: three tiles sharing P fit one SM's shared memory with shared-memory operands (12 warps per
  // SM instead of the 8 of two one-tile CTAs); taken on request only (want_tri)
  static constexpr bool kTri = !kDual && DX <= 4;
  static constexpr int kMulti = kDual ? 2 : (kTri ? 3 : 1);
  static constexpr size_t kMaxDyn = 227 * 1024;
  static size_t smem_b(int M) { return TcCtx<DIN, DH>::bytes(M); }
  static size_t smem_f(int M) { return TcCtx<DIN, DX>::bytes(M); }
  static size_t smem_rb(int M) { return TcCtx<DIN, DH, true>::bytes(M); }
  static size_t smem_rf(int M) { return TcCtx<DIN, DX, true>::bytes(M); }

  static int sm_count() {
    static int sms = 0;
    if (sms == 0) {
      int dev = 0;
      if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        sms = 148;
    }
    return sms;
  }
  static bool want_split(int ctas) {
    if (!kSplit) return false;
    const char *e = getenv("CBFSSM_B200_TC_SPLIT");      // 0: never, 1: always (tests), unset: by launch size
    if (e != nullptr && *e) return atoi(e) != 0;
    return ctas <= sm_count();
  }

  // Three-tile CTAs (12 warps per SM, operands in shared memory) were worth up to 4 % where they filled whole waves
  // better than one-tile CTAs; since the one-tile kernels keep their operands in TMEM they are as fast or faster at
  // every batch measured, so the variant is only taken on request (tests, measurements).
  static bool want_tri(int, int) {
    const char *e = getenv("CBFSSM_B200_TC_TILES");      // 3: three-tile CTAs where they fit; anything else: never
    return e != nullptr && *e && atoi(e) == 3;
  }

  // Launch `k1` (one particle tile per CTA), or `ks` (one tile, kSplitG threads per particle) when the launch is
  // latency-bound, or, when only one one-tile CTA fits an SM and the two-tile CTA fits at all, `k2` (two tiles
  // sharing P and the tables).  `launch(kernel, grid_x, threads, smem)` enqueues.
  template <class K1, class K2, class KS, class F>
  static cudaError_t pick(K1 k1, K2 k2, KS ks, size_t smem1, size_t smem2, size_t smems, int n_local, int nchain, F launch) {
    const int tiles = ceil_div(n_local, kTcThreads);
    cudaError_t e;
    if constexpr (kSplit) {
      if (want_split(tiles * nchain) && smems <= kMaxDyn) {
        e = cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smems);
        if (e != cudaSuccess) return e;
        launch(ks, tiles, kSplitG * kTcThreads, smems);
        return cudaGetLastError();
      }
    }
    if constexpr (kTri) {
      cudaFuncAttributes fa;
      if ((e = cudaFuncGetAttributes(&fa, k2)) != cudaSuccess) return e;
      if (smem2 + fa.sharedSizeBytes <= kMaxDyn && want_tri(tiles, nchain)) {
        e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return e;
        launch(k2, ceil_div(tiles, 3), 3 * kTcThreads, smem2);
        return cudaGetLastError();
      }
    }
    e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    if (e != cudaSuccess) return e;
    if constexpr (kDual) {
      int nb = 0;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k1, kTcThreads, smem1);
      if (e != cudaSuccess) return e;
      if (nb < 2 && smem2 <= kMaxDyn) {
        e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return e;
        launch(k2, ceil_div(tiles, 2), 2 * kTcThreads, smem2);
        return cudaGetLastError();
      }
    }
    launch(k1, tiles, kTcThreads, smem1);
    return cudaGetLastError();
  }
  static constexpr int G = kSplit ? kSplitG : 1;

  static cudaError_t fw_reverse(const Dims &D, GpDev gp, const float *vx, const float *vy, const float *u,
                                const float *y, const float *eps_f, float w_ll, float w_kl, Workspace ws,
                                TcMats mats, TimeWin win, float *spart, int nsc, cudaStream_t st) {
    auto launch = [&](auto kernel, int gx, int threads, size_t smem) {
      kernel<<<gx, threads, smem, st>>>(D, gp, vx, vy, u, y, eps_f, w_ll, w_kl, ws, mats, win, spart, nsc); cbf_note_launch();
    };
    const size_t s1 = TcCtx<DIN, DX, true, 0, 1>::bytes(D.M), s2 = TcCtx<DIN, DX, true, 0, kMulti>::bytes(D.M),
                 ss = TcCtx<DIN, DX, true, 0, 1, G>::bytes(D.M);
    if constexpr (kHas100) {
      if (D.M == 100)
        return pick(fw_reverse_tc_kernel<DX, DU, DY, 100, 1>, fw_reverse_tc_kernel<DX, DU, DY, 100, kMulti>,
                    fw_reverse_tc_kernel<DX, DU, DY, 100, 1, G>, s1, s2, ss, D.n_local, 1, launch);
    }
    return pick(fw_reverse_tc_kernel<DX, DU, DY, 0, 1>, fw_reverse_tc_kernel<DX, DU, DY, 0, kMulti>,
                fw_reverse_tc_kernel<DX, DU, DY, 0, 1, G>, s1, s2, ss, D.n_local, 1, launch);
  }
  static cudaError_t bm_reverse(const Dims &D, const ChainTable &ct, GpDev gp, const float *vx, const float *u,
                                const float *y, const float *eps_b, const float *z_b, float w_en, Workspace ws,
                                TcMats mats, float *spart, int nsc, cudaStream_t st) {
    if (ct.count == 0) return cudaSuccess;
    auto launch = [&](auto kernel, int gx, int threads, size_t smem) {
      kernel<<<dim3(gx, ct.count), threads, smem, st>>>(D, ct, gp, vx, u, y, eps_b, z_b, w_en, ws, mats, spart, nsc); cbf_note_launch();
    };
    const size_t s1 = TcCtx<DIN, DH, true, 0, 1>::bytes(D.M), s2 = TcCtx<DIN, DH, true, 0, kMulti>::bytes(D.M),
                 ss = TcCtx<DIN, DH, true, 0, 1, G>::bytes(D.M);
    if constexpr (kHas100) {
      if (D.M == 100)
        return pick(bm_reverse_tc_kernel<DX, DU, DY, 100, 1>, bm_reverse_tc_kernel<DX, DU, DY, 100, kMulti>,
                    bm_reverse_tc_kernel<DX, DU, DY, 100, 1, G>, s1, s2, ss, D.n_local, ct.count, launch);
    }
    return pick(bm_reverse_tc_kernel<DX, DU, DY, 0, 1>, bm_reverse_tc_kernel<DX, DU, DY, 0, kMulti>,
                bm_reverse_tc_kernel<DX, DU, DY, 0, 1, G>, s1, s2, ss, D.n_local, ct.count, launch);
  }

  static cudaError_t bm_forward(const Dims &D, const ChainTable &ct, GpDev gp, const float *vx, const float *u,
                                const float *y, const float *eps_b, const float *z_b, Workspace ws,
                                float *part_out, cudaStream_t st) {
    if (ct.count == 0) return cudaSuccess;
    auto launch = [&](auto kernel, int gx, int threads, size_t smem) {
      kernel<<<dim3(gx, ct.count), threads, smem, st>>>(D, ct, gp, vx, u, y, eps_b, z_b, ws, part_out); cbf_note_launch();
    };
    const size_t s1 = TcCtx<DIN, DH, false, 0, 1>::bytes(D.M), s2 = TcCtx<DIN, DH, false, 0, kMulti>::bytes(D.M),
                 ss = TcCtx<DIN, DH, false, 0, 1, G>::bytes(D.M);
    if constexpr (kHas100) {
      if (D.M == 100)
        return pick(bm_forward_tc_kernel<DX, DU, DY, 100, 1>, bm_forward_tc_kernel<DX, DU, DY, 100, kMulti>,
                    bm_forward_tc_kernel<DX, DU, DY, 100, 1, G>, s1, s2, ss, D.n_local, ct.count, launch);
    }
    return pick(bm_forward_tc_kernel<DX, DU, DY, 0, 1>, bm_forward_tc_kernel<DX, DU, DY, 0, kMulti>,
                bm_forward_tc_kernel<DX, DU, DY, 0, 1, G>, s1, s2, ss, D.n_local, ct.count, launch);
  }
  static cudaError_t fw_forward(const Dims &D, GpDev gp, const float *vx, const float *vy, const float *u,
                                const float *y, const float *eps_f, Workspace ws, float *part_out,
                                cudaStream_t st) {
    auto launch = [&](auto kernel, int gx, int threads, size_t smem) {
      kernel<<<gx, threads, smem, st>>>(D, gp, vx, vy, u, y, eps_f, ws, part_out); cbf_note_launch();
    };
    const size_t s1 = TcCtx<DIN, DX, false, 0, 1>::bytes(D.M), s2 = TcCtx<DIN, DX, false, 0, kMulti>::bytes(D.M),
                 ss = TcCtx<DIN, DX, false, 0, 1, G>::bytes(D.M);
    if constexpr (kHas100) {
      if (D.M == 100)
        return pick(fw_forward_tc_kernel<DX, DU, DY, 100, 1>, fw_forward_tc_kernel<DX, DU, DY, 100, kMulti>,
                    fw_forward_tc_kernel<DX, DU, DY, 100, 1, G>, s1, s2, ss, D.n_local, 1, launch);
    }
    return pick(fw_forward_tc_kernel<DX, DU, DY, 0, 1>, fw_forward_tc_kernel<DX, DU, DY, 0, kMulti>,
                fw_forward_tc_kernel<DX, DU, DY, 0, 1, G>, s1, s2, ss, D.n_local, 1, launch);
  }
};

}  // namespace cbf
