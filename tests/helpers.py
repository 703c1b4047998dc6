"""Shared builders for oracle-vs-CUDA parity tests (seeded synthetic problems)."""
import numpy as np

from oracle import cbfssm_oracle as O


def make_problem(dx, du, dy, M, S, B, T, R, kap=1.0, lf=(10.0, 0.0), seed=0, strong=False, **cfg_over):
    """Config + raw params + data + injected draws.  ``strong`` uses larger GP signal /
    q(u) variance so that every gradient path carries weight (a harder parity case)."""
    kw = dict(dim_x=dx, dim_u=du, dim_y=dy, ind_pnt_num=M, samples=S, recog_len=R, k_factor=kap, loss_factors=lf)
    if strong:
        kw.update(zeta_mean=0.3, zeta_var=0.05, gp_var=0.5)
    kw.update(cfg_over)
    cfg = O.OracleConfig(**kw)
    params = O.init_params(cfg, seed)
    g = np.random.default_rng(seed + 1000)
    # smooth inputs/outputs (AR(1), rho=.9) of unit scale, like normalised trajectories
    def ar1(shape):
        e = g.standard_normal(shape)
        out = np.empty(shape)
        out[:, 0] = e[:, 0]
        for t in range(1, shape[1]):
            out[:, t] = 0.9 * out[:, t - 1] + np.sqrt(1 - 0.81) * e[:, t]
        return out
    u = ar1((B, T, du))
    y = ar1((B, T, dy))
    eps_b, z_b, eps_f = O.draw_noise(B, S, T, seed + 2000)
    return cfg, params, u, y, eps_b, z_b, eps_f


def rel_inf(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))
