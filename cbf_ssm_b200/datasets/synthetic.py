"""Synthetic datasets with the shapes of the reference's named configurations
(SURVEY.md 8d).  The .mat files of the reference are not shipped and there is no
network, so each class regenerates data of the same shape and a similar distribution
from a seed; they feed both the CUDA path and the oracle."""
import numpy as np

from .base_ds import BaseDS


def _ar1(rng, shape, rho):
    e = rng.standard_normal(shape)
    out = np.empty(shape)
    out[..., 0, :] = e[..., 0, :]
    for t in range(1, shape[-2]):
        out[..., t, :] = rho * out[..., t - 1, :] + np.sqrt(1.0 - rho * rho) * e[..., t, :]
    return out


class _SyntheticDS(BaseDS):
    n_train_exp, n_test_exp, exp_len, rho = 1, 1, 1000, 0.95

    def __init__(self, seq_len, seq_stride, seed=0):
        super().__init__(seq_len, seq_stride)
        u, y = self.generate(np.random.default_rng(seed))
        ne = self.n_train_exp
        self.normalize_init(u[:ne].reshape(-1, self.dim_u), y[:ne].reshape(-1, self.dim_y))
        u, y = self.normalize(u, 'in'), self.normalize(y, 'out')
        self.train_in, self.train_out = u[:ne], y[:ne]
        self.test_in, self.test_out = u[ne:], y[ne:]
        self.create_batches()

    def generate(self, rng):
        n = self.n_train_exp + self.n_test_exp
        return (_ar1(rng, (n, self.exp_len, self.dim_u), self.rho),
                _ar1(rng, (n, self.exp_len, self.dim_y), self.rho))


class SpringNonlinearSynthetic(_SyntheticDS):
    """dim_u=1, dim_y=1; a 3-state linear spring driven through tanh(2u) with
    piecewise-constant U(-2,2) input, observed position with noise var 1e-4
    (the simulation of create_datasets/create_spring_nonlinear.py:36-84, restated)."""
    dim_u, dim_y = 1, 1
    n_train_exp, n_test_exp, exp_len = 1, 1, 5000

    def generate(self, rng):
        b, k, m, dt = 0.05, 1.0, 0.002, 0.01
        A = np.array([[1.0, dt, 0.0], [0.0, 1.0, dt], [-k / m, -b / m, 0.0]])
        Bv = np.array([0.0, 0.0, 1.0 / m])
        total = 2 * self.exp_len
        levels = rng.uniform(-2.0, 2.0, size=total // 100)
        x = np.array([1.0, 0.0, 0.0])
        us, ys = np.empty(total), np.empty(total)
        for _ in range(5):
            x = A @ x + Bv * np.tanh(2.0 * levels[0])
        for t in range(total):
            ut = levels[min(t // 100, levels.size - 1)]
            us[t] = ut
            ys[t] = x[0] + np.sqrt(1e-4) * rng.standard_normal()
            x = A @ x + Bv * np.tanh(2.0 * ut)
        return us.reshape(2, self.exp_len, 1), ys.reshape(2, self.exp_len, 1)


class RoboMoveSynthetic(_SyntheticDS):
    """dim_u=2 (speed, curvature), dim_y=2 (noisy position): a unicycle as in
    create_datasets/create_robomove.py:9-75, restated with smooth random controls."""
    dim_u, dim_y = 2, 2
    n_train_exp, n_test_exp, exp_len = 1, 1, 5000

    def generate(self, rng):
        total = 2 * self.exp_len
        ctl = _ar1(rng, (1, total, 2), 0.98)[0]
        speed = 0.05 * (1.0 + 0.5 * ctl[:, 0])
        curv = 0.8 * ctl[:, 1]
        pos, th = np.zeros(2), 0.0
        ys = np.empty((total, 2))
        for t in range(total):
            ys[t] = pos + np.sqrt(1e-4) * rng.standard_normal(2)
            th = (th + speed[t] * curv[t]) % (2 * np.pi)
            pos = pos + speed[t] * np.array([np.sin(th), np.cos(th)])
            pos = np.clip(pos, -3.0, 3.0)
        us = np.stack((speed, curv), axis=1)
        return us.reshape(2, self.exp_len, 2), ys.reshape(2, self.exp_len, 2)


class SarcosSynthetic(_SyntheticDS):
    """dim_u=7, dim_y=7: 60+6 experiments of 337 steps (cbfssm/datasets/prssm_ds.py:32-38;
    prssm/real_world_tasks.py: 66 x 674 rows, downsample 2, 60 train / 6 test)."""
    dim_u, dim_y = 7, 7
    n_train_exp, n_test_exp, exp_len, rho = 60, 6, 337, 0.95


class VoliroShapedSynthetic(_SyntheticDS):
    """dim_u=6, dim_y=7 (state 13): multi-experiment windows with the dims of
    cbfssm/model/voliro.py:13-18 for the CBFSSM class."""
    dim_u, dim_y = 6, 7
    n_train_exp, n_test_exp, exp_len, rho = 24, 4, 256, 0.9
