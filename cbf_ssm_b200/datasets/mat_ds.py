"""On-disk dataset format of the reference: MATLAB v5 ``.mat`` files with the keys ``ds_u``,
``ds_x``, ``ds_y`` ([time, dim] each, x_{i+1} = f(x_i, u_i), y_i = g(x_i)) and ``title``
(cbfssm/datasets/ds_manager.py:11-34), plus the three dataset classes built on it
(cbfssm/datasets/dsmanager_ds.py:6-63) and generators that write files of the same format
(create_datasets/create_spring_nonlinear.py, create_datasets/create_robomove.py: the systems are
restated here, vectorised where the dynamics allow; the noise stream differs from the reference's
legacy ``np.random`` global stream, so files are statistically, not bitwise, equivalent).

Files written by the reference's generators load unchanged:

    ds = RoboMove(seq_len=300, seq_stride=50)                 # <data_path>/robomove.mat
    ds = SpringNonlinear(100, 50, data_path="/data/cbfssm/")  # any directory
"""
import os

import numpy as np
import scipy.io

from .base_ds import BaseDS

DEFAULT_DATA_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data") + os.sep


class DSManager:
    """load / save / normalise in the reference's conventions (ds_manager.py:5-84)."""

    KEYS = ("ds_u", "ds_x", "ds_y")

    @staticmethod
    def load_ds(filename, normalize=False, print_title=True, dtype=np.float64):
        mat = scipy.io.loadmat(filename)
        if print_title:
            print("Loaded Dataset " + "".join(np.atleast_1d(mat["title"]).tolist()))
        u, x, y = (np.asarray(mat[k], dtype=dtype) for k in DSManager.KEYS)
        if normalize:
            u, x, y = (DSManager.normalize_ds(a) for a in (u, x, y))
        return u, x, y

    @staticmethod
    def save_ds(filename, u, x, y, title, dtype=np.float64):
        arrays = [np.asarray(a) for a in (u, x, y)]
        if any(a.ndim != 2 for a in arrays) or len({a.shape[0] for a in arrays}) != 1:
            raise AssertionError("u, x, y must be [ds_size, dim] with equal ds_size")
        payload = {k: a.astype(dtype) for k, a in zip(DSManager.KEYS, arrays)}
        payload["title"] = title
        scipy.io.savemat(filename, payload)

    @staticmethod
    def sample_ds(system, ds_size, u_fn):
        """Roll a system object (get_state / measure / propagate) for ds_size steps
        (ds_manager.py:61-79): record x_i, y_i = measure(), then u_i = u_fn(i, x_i) and propagate."""
        us, xs, ys = [], [], []
        for i in range(ds_size):
            x = system.get_state()
            xs.append(x)
            ys.append(system.measure())
            u = u_fn(i, x)
            us.append(u)
            system.propagate(u)
        return np.asarray(us), np.asarray(xs), np.asarray(ys)

    @staticmethod
    def normalize_ds(data):
        centred = data - data.mean(axis=0)
        return centred / centred.std(axis=0)


class DSManagerDS(BaseDS):
    """One long experiment split at ``split`` into train / test (dsmanager_ds.py:6-28)."""

    def __init__(self, seq_len, seq_stride, data_path=None):
        super().__init__(seq_len, seq_stride)
        self.data_path = DEFAULT_DATA_PATH if data_path is None else os.path.join(data_path, "")

    def prepare_data(self, path, split, y_crop=None):
        u, _, y = DSManager.load_ds(path)
        if y_crop is not None:
            y = y[:, :y_crop]
        self.normalize_init(u, y)                   # statistics over the WHOLE file, as the reference
        u, y = self.normalize(u, 'in'), self.normalize(y, 'out')
        self.train_in, self.test_in = u[None, :split], u[None, split:]
        self.train_out, self.test_out = y[None, :split], y[None, split:]
        self.create_batches()


class RoboMoveSimple(DSManagerDS):
    dim_u, dim_y = 2, 4

    def __init__(self, seq_len, seq_stride, data_path=None):
        super().__init__(seq_len, seq_stride, data_path)
        self.prepare_data(self.data_path + 'robomove_simple.mat', 25000)


class RoboMove(DSManagerDS):
    dim_u, dim_y = 2, 2

    def __init__(self, seq_len, seq_stride, data_path=None):
        super().__init__(seq_len, seq_stride, data_path)
        self.prepare_data(self.data_path + 'robomove.mat', 25000)


class SpringNonlinear(DSManagerDS):
    dim_u, dim_y = 1, 1

    def __init__(self, seq_len, seq_stride, data_path=None):
        super().__init__(seq_len, seq_stride, data_path)
        self.prepare_data(self.data_path + 'spring_nonlinear.mat', 5000, y_crop=1)


# --------------------------------------------------------------------------------------
# generators (same systems, same file format)
# --------------------------------------------------------------------------------------
def create_spring_nonlinear(filename, ds_size=10000, seed=0, hold=100):
    """3-state linear spring driven through tanh(2u); u piecewise constant U(-2,2) for ``hold``
    steps; y = x + N(0, 1e-4 I) (create_spring_nonlinear.py:36-84: b=.05, k=1, m=.002, dt=.01,
    no process noise).  ds_y has 3 columns; the dataset class keeps the first (y_crop=1)."""
    rng = np.random.default_rng(seed)
    b, k, m, dt = 0.05, 1.0, 0.002, 0.01
    A = np.array([[1.0, dt, 0.0], [0.0, 1.0, dt], [-k / m, -b / m, 0.0]])
    Bv = np.array([0.0, 0.0, 1.0 / m])
    levels = rng.uniform(-2.0, 2.0, size=(ds_size + hold - 1) // hold)
    u = np.repeat(levels, hold)[:ds_size, None]
    x = np.zeros((ds_size, 3))
    state = np.zeros(3)
    for i in range(ds_size):
        x[i] = state
        state = A @ state + Bv * np.tanh(2.0 * u[i, 0])
    y = x + np.sqrt(1e-4) * rng.standard_normal(x.shape)
    DSManager.save_ds(filename, u, x, y, 'Nonlinear Spring')
    return u, x, y


class _Unicycle:
    """Robot on a plane: u = [distance travelled this step, curvature]; heading measured from the
    +y axis (create_robomove.py:9-75).  ``simple`` keeps the heading as a (sin, cos) pair and
    observes the full state (create_robomove.py:78-153)."""

    def __init__(self, rng, sigma_x, sigma_y, simple):
        self.rng, self.sigma_x, self.sigma_y, self.simple = rng, sigma_x, sigma_y, simple
        self.pos = np.zeros(2)
        self.heading = np.array([0.0, 1.0])          # (sin, cos) of the orientation angle
        self.angle = 0.0

    def get_state(self):
        tail = self.heading if self.simple else [self.angle]
        return np.concatenate((self.pos, tail))

    def measure(self):
        x = self.get_state() if self.simple else self.pos
        return x + np.sqrt(self.sigma_y) * self.rng.standard_normal(x.shape)

    def propagate(self, u):
        dist, curv = float(u[0]), float(u[1])
        hx, hy = self.heading
        if abs(curv) < 1e-5:
            self.pos = self.pos + dist * self.heading
        else:
            side = np.sign(curv)
            radius = 1.0 / abs(curv)
            turn = side * dist / radius
            c, s = np.cos(turn), np.sin(turn)
            normal = side * np.array([hy, -hx])
            rotated = np.array([c * normal[0] + s * normal[1], -s * normal[0] + c * normal[1]])
            self.pos = self.pos + radius * (normal - rotated)
            self.heading = np.array([c * hx + s * hy, -s * hx + c * hy])
            self.angle = (self.angle + turn) % (2.0 * np.pi)
        self.pos = self.pos + np.sqrt(self.sigma_x) * self.rng.standard_normal(2)


def create_robomove(filename, ds_size=30000, seed=0, simple=False, sigma_x=1e-6, sigma_y=1e-4, hold=10):
    """Unicycle with piecewise-constant random speed / curvature commands that steer back towards the
    origin when the robot drifts away, so the trajectory stays bounded (the role of the reference's
    input function, create_robomove.py:156-204)."""
    rng = np.random.default_rng(seed)
    robot = _Unicycle(rng, sigma_x, sigma_y, simple)
    cmd = np.zeros(2)

    def u_fn(i, x):
        nonlocal cmd
        if i % hold == 0:
            speed = rng.uniform(0.05, 0.25)
            curv = rng.uniform(-2.0, 2.0)
            r = np.hypot(x[0], x[1])
            if r > 5.0:                                # turn towards the origin
                hx, hy = robot.heading
                cross = hx * (-x[1]) - hy * (-x[0])
                curv = -np.sign(cross) * rng.uniform(0.5, 2.0)
            cmd = np.array([speed, curv])
        return cmd.copy()

    u, x, y = DSManager.sample_ds(robot, ds_size, u_fn)
    DSManager.save_ds(filename, u, x, y, 'RoboMove Simple' if simple else 'RoboMove')
    return u, x, y
