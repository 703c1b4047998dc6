// Host launchers for one (dx,du,dy) instantiation of the kernel templates.
#pragma once
#include "kernels_generic.cuh"
#include "kernels_tc.cuh"

namespace cbf {

inline int coop_parts(int M) {
  const int MG = round_up(M, 4) / 4;
  return MG < 1 ? 1 : (MG > 8 ? 8 : MG);
}

template <int DX, int DU, int DY>
struct Launch {
  static constexpr int DH = DX - DY, DIN = DX + DU;
  static constexpr int VXP = 4 * ((DX + 3) / 4);

  static size_t smem_bytes(int M, int which) {
    const int MP = round_up(M, 4), parts = coop_parts(M);
    size_t f = 0;
    if (which == 0) {
      f = GpS<DIN, DH>::floats_host(M) + (size_t)MP * kLD + (size_t)parts * (1 + 2 * DH) * kNP + VXP;
    } else if (which == 1) {
      f = GpS<DIN, DX>::floats_host(M) + (size_t)MP * kLD + (size_t)parts * (1 + 2 * DX) * kNP + 2 * VXP;
    } else if (which == 2) {
      const AccLayout L(M, DIN, DX, DX);
      const int DG = (DX + 3) / 4, XG = (DIN + 4) / 4;
      f = GpS<DIN, DX>::floats_host(M) + (size_t)4 * MP * kLD + (size_t)(8 * DG + 4 * XG) * kLD + L.nacc +
          (size_t)parts * (1 + 2 * DX) * kNP + 2 * VXP;
    } else {
      const AccLayout L(M, DIN, DH, DX);
      const int DG = (DH + 3) / 4, XG = (DIN + 4) / 4;
      f = GpS<DIN, DH>::floats_host(M) + (size_t)4 * MP * kLD + (size_t)(8 * DG + 4 * XG) * kLD + L.nacc +
          (size_t)parts * (1 + 2 * DH) * kNP + VXP;
    }
    return f * sizeof(float);
  }

  template <typename K>
  static cudaError_t prep(K kernel, size_t smem) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }

  static cudaError_t bm_forward(const Dims &D, const ChainTable &ct, GpDev gp, const float *vx, const float *u,
                                const float *y, const float *eps_b, const float *z_b, Workspace ws,
                                float *part_out, cudaStream_t st) {
    if (ct.count == 0) return cudaSuccess;
    const size_t smem = smem_bytes(D.M, 0);
    cudaError_t e = prep(bm_forward_kernel<DX, DU, DY>, smem);
    if (e != cudaSuccess) return e;
    dim3 grid(ceil_div(D.n_local, kNP), ct.count);
    bm_forward_kernel<DX, DU, DY><<<grid, 32 * coop_parts(D.M), smem, st>>>(D, ct, gp, vx, u, y, eps_b, z_b, ws, part_out); cbf_note_launch();
    return cudaGetLastError();
  }

  static cudaError_t fw_forward(const Dims &D, GpDev gp, const float *vx, const float *vy, const float *u,
                                const float *y, const float *eps_f, Workspace ws, float *part_out,
                                cudaStream_t st) {
    const size_t smem = smem_bytes(D.M, 1);
    cudaError_t e = prep(fw_forward_kernel<DX, DU, DY>, smem);
    if (e != cudaSuccess) return e;
    fw_forward_kernel<DX, DU, DY><<<ceil_div(D.n_local, kNP), 32 * coop_parts(D.M), smem, st>>>(
        D, gp, vx, vy, u, y, eps_f, ws, part_out); cbf_note_launch();
    return cudaGetLastError();
  }

  static cudaError_t fw_reverse(const Dims &D, GpDev gp, const float *vx, const float *vy, const float *u,
                                const float *y, const float *eps_f, float w_ll, float w_kl, Workspace ws,
                                float *part_out, int grid, cudaStream_t st) {
    const size_t smem = smem_bytes(D.M, 2);
    cudaError_t e = prep(fw_reverse_kernel<DX, DU, DY>, smem);
    if (e != cudaSuccess) return e;
    fw_reverse_kernel<DX, DU, DY><<<grid, 32 * coop_parts(D.M), smem, st>>>(D, gp, vx, vy, u, y, eps_f, w_ll, w_kl,
                                                                            ws, part_out); cbf_note_launch();
    return cudaGetLastError();
  }

  static cudaError_t bm_reverse(const Dims &D, const ChainTable &ct, GpDev gp, const float *vx, const float *u,
                                const float *y, const float *eps_b, const float *z_b, float w_en, Workspace ws,
                                float *part_out, int grid, cudaStream_t st) {
    const size_t smem = smem_bytes(D.M, 3);
    cudaError_t e = prep(bm_reverse_kernel<DX, DU, DY>, smem);
    if (e != cudaSuccess) return e;
    bm_reverse_kernel<DX, DU, DY><<<grid, 32 * coop_parts(D.M), smem, st>>>(D, ct, gp, vx, u, y, eps_b, z_b, w_en, ws,
                                                                            part_out); cbf_note_launch();
    return cudaGetLastError();
  }

  static void layouts(int M, AccLayout *Lf, AccLayout *Lb) {
    *Lf = AccLayout(M, DIN, DX, DX);
    *Lb = AccLayout(M, DIN, DH, DX);
  }

  // Resident CTAs per SM of the two persistent reverse kernels (0 if they do not fit).
  static int occupancy(int M, int which) {
    const size_t smem = smem_bytes(M, which);
    int nb = 0;
    cudaError_t e;
    if (which == 2) {
      if (prep(fw_reverse_kernel<DX, DU, DY>, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fw_reverse_kernel<DX, DU, DY>, 32 * coop_parts(M), smem);
    } else {
      if (prep(bm_reverse_kernel<DX, DU, DY>, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, bm_reverse_kernel<DX, DU, DY>, 32 * coop_parts(M), smem);
    }
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return nb;
  }
};

template <int DX, int DU, int DY>
DimOps make_ops() {
  DimOps o;
  o.dx = DX; o.du = DU; o.dy = DY;
  o.bm_forward = &Launch<DX, DU, DY>::bm_forward;
  o.fw_forward = &Launch<DX, DU, DY>::fw_forward;
  o.fw_reverse = &Launch<DX, DU, DY>::fw_reverse;
  o.bm_reverse = &Launch<DX, DU, DY>::bm_reverse;
  o.bm_forward_tc = &LaunchTc<DX, DU, DY>::bm_forward;
  o.fw_forward_tc = &LaunchTc<DX, DU, DY>::fw_forward;
  o.fw_reverse_tc = &LaunchTc<DX, DU, DY>::fw_reverse;
  o.bm_reverse_tc = &LaunchTc<DX, DU, DY>::bm_reverse;
  o.smem_tc = [](int M, int which) {
    using L = LaunchTc<DX, DU, DY>;
    return which == 0 ? L::smem_b(M) : which == 1 ? L::smem_f(M) : which == 2 ? L::smem_rf(M) : L::smem_rb(M);
  };
  o.smem_bytes = &Launch<DX, DU, DY>::smem_bytes;
  o.occupancy = &Launch<DX, DU, DY>::occupancy;
  o.layouts = &Launch<DX, DU, DY>::layouts;
  o.saved_planes = nullptr;
  o.slots_per_cta = 1;
  o.particles_per_cta = kNP;
  o.fixed_M = 0;
  return o;
}

}  // namespace cbf

#define CBF_INSTANTIATE(DX, DU, DY)                                   \
  namespace cbf {                                                     \
  const DimOps *ops_##DX##_##DU##_##DY() {                            \
    static const DimOps o = make_ops<DX, DU, DY>();                   \
    return &o;                                                        \
  }                                                                   \
  }
