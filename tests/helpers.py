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


def elementwise_err(a, b, floor=0.05):
    """max over entries of |a - b| / max(|b|, floor * max|b|).  ``rel_inf`` alone leaves entries far below a
    tensor's maximum unchecked; this bounds each entry relative to itself, down to ``floor`` of the maximum.
    (Every entry of these tensors is a float32 sum over all particle-steps whose rounding error scales with
    the summands, not with the entry, so an entry that nearly cancels carries no more significant digits than
    ``floor`` allows.)"""
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    scale = np.maximum(np.abs(b), floor * np.max(np.abs(b)) + 1e-300)
    return float(np.max(np.abs(a - b) / scale))


def elementwise_ok(a, b, tol, floor=0.05):
    return elementwise_err(a, b, floor) <= tol


# Named configurations of SURVEY.md 8(d) at the reference's own batch sizes
# (cfg4 / cfg5 with the particle count reduced so the oracle finishes in seconds).
NAMED_CASES = {
    # name: dims, M, S, B, T, R, kap, lf, seed, config overrides
    "cfg1_spring_template": dict(dx=4, du=1, dy=1, M=100, S=50, B=32, T=100, R=50, kap=1.0, lf=(10.0, 0.0), seed=0,
                                 over=dict(zeta_pos=2.0, zeta_mean=0.01, zeta_var=1e-4, gp_var=0.01, gp_len=1.0)),
    "cfg2_robomove_m20": dict(dx=4, du=2, dy=2, M=20, S=50, B=32, T=300, R=50, kap=1.0, lf=(20.0, 0.0), seed=1,
                              over=dict(zeta_pos=2.0, zeta_mean=0.01, zeta_var=1e-4, gp_var=0.01, gp_len=1.0)),
    "cfg3_sarcos": dict(dx=14, du=7, dy=7, M=100, S=20, B=5, T=250, R=16, kap=50.0, lf=(6.0, 0.0), seed=2,
                        over=dict(zeta_pos=2.0, zeta_mean=0.0025, zeta_var=1e-4, gp_var=0.25, gp_len=1.0,
                                  var_x=np.full(14, 4e-6), var_y=np.full(14, 0.0025))),
    "cfg4_voliro_shaped": dict(dx=13, du=6, dy=7, M=20, S=64, B=4, T=64, R=16, kap=1.0, lf=(20.0, 0.0), seed=3,
                               over=dict(zeta_pos=2.0, zeta_mean=0.01, zeta_var=1e-4, gp_var=0.25, gp_len=5.0,
                                         var_x=np.full(13, 0.02 ** 2), var_y=np.full(13, 0.05 ** 2))),
    "cfg5_sweep_d8_m100": dict(dx=8, du=1, dy=4, M=100, S=16, B=4, T=60, R=16, kap=1.0, lf=(10.0, 0.0), seed=4,
                               over=dict(zeta_pos=2.0, zeta_mean=0.01, zeta_var=1e-4, gp_var=0.01, gp_len=1.0)),
    "cfg2_predict_free_run": dict(dx=4, du=2, dy=2, M=20, S=50, B=2, T=120, R=50, kap=1.0, lf=(20.0, 0.0), seed=5,
                                  cond=False,
                                  over=dict(zeta_pos=2.0, zeta_mean=0.01, zeta_var=1e-4, gp_var=0.01, gp_len=1.0)),
}


def named_case(name):
    c = NAMED_CASES[name]
    cfg, params, u, y, eps_b, z_b, eps_f = make_problem(c["dx"], c["du"], c["dy"], c["M"], c["S"], c["B"], c["T"],
                                                        c["R"], c["kap"], c["lf"], seed=c["seed"], **c["over"])
    return cfg, params, u, y, eps_b, z_b, eps_f, c.get("cond", True)


# Harder cases ("strong": every gradient path carries weight; entropy factor on) that also get a
# fixture from the reference's own source (tests/golden/make_golden.py).
STRONG_REF_CASES = {
    # name: (dx, du, dy, M, S, B, T, R, kap, lf, seed, condition)
    "strong_m20": (4, 2, 2, 20, 16, 4, 24, 6, 1.0, (20.0, 0.1), 11, True),
    "strong_m100": (4, 2, 2, 100, 16, 8, 40, 10, 1.0, (10.0, 0.3), 12, True),
    "strong_m7_free_run": (4, 2, 2, 7, 5, 3, 17, 4, 3.0, (5.0, 0.2), 13, False),
    "strong_sarcos_dims_m33": (14, 7, 7, 33, 6, 2, 20, 4, 50.0, (6.0, 0.1), 14, True),
}


def strong_ref_case(name):
    dx, du, dy, M, S, B, T, R, kap, lf, seed, cond = STRONG_REF_CASES[name]
    return make_problem(dx, du, dy, M, S, B, T, R, kap, lf, seed=seed, strong=True) + (cond,)


HALF_REF_CASES = {
    # name: (dx, du, dy, M, S, B, T, R, kap, seed, condition, recog_model)
    "half_rnn_m20": (4, 2, 2, 20, 8, 3, 30, 8, 2.0, 5, True, "rnn"),
    "half_output_m20_free_run": (4, 1, 1, 20, 8, 3, 25, 5, 10.0, 6, False, "output"),
    "half_rnn_m100": (4, 1, 1, 100, 10, 2, 20, 6, 1.0, 7, True, "rnn"),
}


def half_ref_case(name):
    """CBFSSMHALF problem incl. GRU(16)+dense recognition weights (TF creation order)."""
    from oracle import cbfssmhalf_oracle as H
    dx, du, dy, M, S, B, T, R, kap, seed, cond, recog = HALF_REF_CASES[name]
    cfg, _, u, y, _, _, eps_f = make_problem(dx, du, dy, M, S, B, T, R, kap, (10.0, 0.0), seed=seed, strong=True)
    params = H.init_params_half(cfg, seed)
    g = np.random.default_rng(seed + 3000)
    nh, din = 16, du + dy

    def glorot(shape):
        lim = np.sqrt(6.0 / (shape[0] + shape[1]))
        return g.uniform(-lim, lim, shape)
    w = dict(gates_kernel=glorot((din + nh, 2 * nh)), gates_bias=np.ones(2 * nh) + 0.1 * g.standard_normal(2 * nh),
             candidate_kernel=glorot((din + nh, nh)), candidate_bias=0.1 * g.standard_normal(nh),
             dense_kernel=glorot((nh, dx)), dense_bias=0.1 * g.standard_normal(dx))
    return cfg, params, w, u, y, eps_f, cond, recog
