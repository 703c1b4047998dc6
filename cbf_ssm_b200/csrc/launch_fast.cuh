// Host launchers for one (dx,du,dy,M) instantiation of the register-resident kernels.
#pragma once
#include "kernels_fast.cuh"

namespace cbf {

template <int DX, int DU, int DY, int M>
struct LaunchFast {
  static constexpr int DH = DX - DY, DIN = DX + DU;
  static constexpr int VXP = 4 * ((DX + 3) / 4);
  using Gf = GpF<M, DIN, DX, 0>;
  using Gb = GpF<M, DIN, DH, 1>;
  using Wf = WarpAcc<M, DIN, DX>;
  using Wb = WarpAcc<M, DIN, DH>;

  static size_t smem_bytes(int, int which) {
    size_t f = 0;
    if (which == 0) f = Gb::FLOATS + 4 * kFastWarps + VXP;
    else if (which == 1) f = Gf::FLOATS + (DY + 1) * kFastWarps + 2 * VXP;
    else if (which == 2) f = Gf::FLOATS + (size_t)kFastWarps * Wf::FLOATS + 2 * VXP;
    else f = Gb::FLOATS + (size_t)kFastWarps * Wb::FLOATS + VXP;
    return f * sizeof(float);
  }

  static void layouts(int, AccLayout *Lf, AccLayout *Lb) {
    *Lf = AccLayout(M, DIN, DX, DX, Wf::TR, Wf::TC, 1);
    *Lb = AccLayout(M, DIN, DH, DX, Wb::TR, Wb::TC, 1);
  }

  template <typename K>
  static cudaError_t prep(K kernel, size_t smem) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }

  // Refresh the constant-bank image of one GP (slot 0: forward-rollout GP, 1: message GP).
  static cudaError_t load_const(int slot, GpDev gp, int dout, float *scratch, cudaStream_t st) {
    if (!kConstOps) return cudaSuccess;
    pack_const_kernel<<<1, 256, 0, st>>>(gp, M, DIN, dout, scratch); cbf_note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t bytes = sizeof(float) * (slot == 0 ? Gf::C::TOTAL : Gb::C::TOTAL);
    return cudaMemcpyToSymbolAsync(c_ops, scratch, bytes, sizeof(float) * (size_t)slot * kConstFloats,
                                   cudaMemcpyDeviceToDevice, st);
  }

  template <bool SAVE>
  static cudaError_t bm_forward_t(const Dims &D, const ChainTable &ct, GpDev gp, const float *vx, const float *u,
                                  const float *y, const float *eps_b, const float *z_b, Workspace ws,
                                  float *part_out, cudaStream_t st) {
    const size_t smem = smem_bytes(M, 0);
    cudaError_t e = prep(bm_forward_fast_kernel<DX, DU, DY, M, SAVE>, smem);
    if (e != cudaSuccess) return e;
    if ((e = load_const(1, gp, DH, ws.cpack + kConstFloats, st)) != cudaSuccess) return e;
    dim3 grid(ceil_div(D.n_local, kFastThreads), ct.count);
    bm_forward_fast_kernel<DX, DU, DY, M, SAVE><<<grid, kFastThreads, smem, st>>>(D, ct, gp, vx, u, y, eps_b, z_b, ws, part_out); cbf_note_launch();
    return cudaGetLastError();
  }
  static cudaError_t bm_forward(const Dims &D, const ChainTable &ct, GpDev gp, const float *vx, const float *u,
                                const float *y, const float *eps_b, const float *z_b, Workspace ws,
                                float *part_out, cudaStream_t st) {
    if (ct.count == 0) return cudaSuccess;
    return ws.KAb != nullptr ? bm_forward_t<true>(D, ct, gp, vx, u, y, eps_b, z_b, ws, part_out, st)
                             : bm_forward_t<false>(D, ct, gp, vx, u, y, eps_b, z_b, ws, part_out, st);
  }

  template <bool SAVE>
  static cudaError_t fw_forward_t(const Dims &D, GpDev gp, const float *vx, const float *vy, const float *u,
                                  const float *y, const float *eps_f, Workspace ws, float *part_out,
                                  cudaStream_t st) {
    const size_t smem = smem_bytes(M, 1);
    cudaError_t e = prep(fw_forward_fast_kernel<DX, DU, DY, M, SAVE>, smem);
    if (e != cudaSuccess) return e;
    if ((e = load_const(0, gp, DX, ws.cpack, st)) != cudaSuccess) return e;
    fw_forward_fast_kernel<DX, DU, DY, M, SAVE><<<ceil_div(D.n_local, kFastThreads), kFastThreads, smem, st>>>(
        D, gp, vx, vy, u, y, eps_f, ws, part_out); cbf_note_launch();
    return cudaGetLastError();
  }
  static cudaError_t fw_forward(const Dims &D, GpDev gp, const float *vx, const float *vy, const float *u,
                                const float *y, const float *eps_f, Workspace ws, float *part_out,
                                cudaStream_t st) {
    return ws.KAf != nullptr ? fw_forward_t<true>(D, gp, vx, vy, u, y, eps_f, ws, part_out, st)
                             : fw_forward_t<false>(D, gp, vx, vy, u, y, eps_f, ws, part_out, st);
  }

  // saved-evaluation planes per evaluation slot (common.cuh Workspace::KAf / KAb), float4 units per particle
  static int saved_planes(int which) { return which == 0 ? SavedEval<Gf::MP, DX>::NPL : SavedEval<Gb::MP, DH>::NPL; }

  template <bool SAVED>
  static cudaError_t fw_reverse_t(const Dims &D, GpDev gp, const float *vx, const float *vy, const float *u,
                                  const float *y, const float *eps_f, float w_ll, float w_kl, Workspace ws,
                                  float *part_out, int grid, cudaStream_t st) {
    const size_t smem = smem_bytes(M, 2);
    cudaError_t e = prep(fw_reverse_fast_kernel<DX, DU, DY, M, SAVED>, smem);
    if (e != cudaSuccess) return e;
    if ((e = load_const(0, gp, DX, ws.cpack, st)) != cudaSuccess) return e;
    AccLayout Lf, Lb;
    layouts(M, &Lf, &Lb);
    fw_reverse_fast_kernel<DX, DU, DY, M, SAVED><<<grid, kFastThreads, smem, st>>>(D, gp, vx, vy, u, y, eps_f, w_ll, w_kl, ws,
                                                                                  part_out, Lf.slot()); cbf_note_launch();
    return cudaGetLastError();
  }
  static cudaError_t fw_reverse(const Dims &D, GpDev gp, const float *vx, const float *vy, const float *u,
                                const float *y, const float *eps_f, float w_ll, float w_kl, Workspace ws,
                                float *part_out, int grid, cudaStream_t st) {
    return ws.KAf != nullptr ? fw_reverse_t<true>(D, gp, vx, vy, u, y, eps_f, w_ll, w_kl, ws, part_out, grid, st)
                             : fw_reverse_t<false>(D, gp, vx, vy, u, y, eps_f, w_ll, w_kl, ws, part_out, grid, st);
  }

  template <bool SAVED>
  static cudaError_t bm_reverse_t(const Dims &D, const ChainTable &ct, GpDev gp, const float *vx, const float *u,
                                  const float *y, const float *eps_b, const float *z_b, float w_en, Workspace ws,
                                  float *part_out, int grid, cudaStream_t st) {
    const size_t smem = smem_bytes(M, 3);
    cudaError_t e = prep(bm_reverse_fast_kernel<DX, DU, DY, M, SAVED>, smem);
    if (e != cudaSuccess) return e;
    if ((e = load_const(1, gp, DH, ws.cpack + kConstFloats, st)) != cudaSuccess) return e;
    AccLayout Lf, Lb;
    layouts(M, &Lf, &Lb);
    bm_reverse_fast_kernel<DX, DU, DY, M, SAVED><<<grid, kFastThreads, smem, st>>>(D, ct, gp, vx, u, y, eps_b, z_b, w_en, ws,
                                                                                  part_out, Lb.slot()); cbf_note_launch();
    return cudaGetLastError();
  }
  static cudaError_t bm_reverse(const Dims &D, const ChainTable &ct, GpDev gp, const float *vx, const float *u,
                                const float *y, const float *eps_b, const float *z_b, float w_en, Workspace ws,
                                float *part_out, int grid, cudaStream_t st) {
    return ws.KAb != nullptr ? bm_reverse_t<true>(D, ct, gp, vx, u, y, eps_b, z_b, w_en, ws, part_out, grid, st)
                             : bm_reverse_t<false>(D, ct, gp, vx, u, y, eps_b, z_b, w_en, ws, part_out, grid, st);
  }

  static int occupancy(int, int which) {
    const size_t smem = smem_bytes(M, which);
    int nb = 0;
    cudaError_t e;
    if (which == 2) {
      if (prep(fw_reverse_fast_kernel<DX, DU, DY, M, false>, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fw_reverse_fast_kernel<DX, DU, DY, M, false>, kFastThreads, smem);
    } else {
      if (prep(bm_reverse_fast_kernel<DX, DU, DY, M, false>, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, bm_reverse_fast_kernel<DX, DU, DY, M, false>, kFastThreads, smem);
    }
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return nb;
  }
};

template <int DX, int DU, int DY, int M>
DimOps make_fast_ops() {
  using L = LaunchFast<DX, DU, DY, M>;
  DimOps o;
  o.dx = DX; o.du = DU; o.dy = DY;
  o.bm_forward = &L::bm_forward;
  o.fw_forward = &L::fw_forward;
  o.fw_reverse = &L::fw_reverse;
  o.bm_reverse = &L::bm_reverse;
  o.bm_forward_tc = nullptr;
  o.fw_forward_tc = nullptr;
  o.fw_reverse_tc = nullptr;
  o.bm_reverse_tc = nullptr;
  o.smem_tc = nullptr;
  o.smem_bytes = &L::smem_bytes;
  o.occupancy = &L::occupancy;
  o.layouts = &L::layouts;
  o.saved_planes = &L::saved_planes;
  o.slots_per_cta = kFastWarps;
  o.particles_per_cta = kFastThreads;
  o.fixed_M = M;
  return o;
}

}  // namespace cbf

#define CBF_INSTANTIATE_FAST(DX, DU, DY, M)                           \
  namespace cbf {                                                     \
  const DimOps *fast_ops_##DX##_##DU##_##DY##_##M() {                 \
    static const DimOps o = make_fast_ops<DX, DU, DY, M>();           \
    return &o;                                                        \
  }                                                                   \
  }
