"""Eager stand-in for the TensorFlow-1.8 symbols the CBF-SSM hot path uses.
TEST INFRASTRUCTURE ONLY (part of ``oracle/``; nothing under ``cbf_ssm_b200/`` imports it).

Purpose: TensorFlow 1.8 cannot be installed here (``setup.py:9`` pins it; Python 3.12, no
network), so the reference's model files could never be executed.  With this package first on
``sys.path`` (``oracle/run_reference.py`` arranges that) ``import tensorflow as tf`` resolves
here and the reference's **unmodified** ``cbfssm/model/{tf_transform,gp_tf,base_model,cbfssm,
cbfssmhalf}.py`` run as they are: every ``tf.*`` call evaluates immediately on float64
PyTorch-CPU tensors, ``tf.while_loop`` / ``tf.cond`` are Python control flow (only the taken
branch of a ``cond`` executes, as in TF), ``tf.gradients`` is autograd, and every
``tf.random_normal`` takes the next array of an injected queue, so the run is reproducible.
Building a model therefore *is* one execution of its graph on the minibatch given to
``configure()``.  What is restated here is TensorFlow's library behaviour (op semantics of
SURVEY.md Appendix A); what is executed unchanged is the reference's own algorithm.

Each function cites the TF-1.8 behaviour it mirrors and the reference call site that needs it.
"""
from __future__ import annotations

import builtins
import collections
import contextlib
import math
import sys
import types

import numpy as np
import torch

__version__ = "1.8.0-shim"


# --------------------------------------------------------------------------
# dtypes
# --------------------------------------------------------------------------
class DType:
    def __init__(self, name, torch_dtype, np_dtype):
        self.name, self.torch, self.np = name, torch_dtype, np_dtype

    def as_numpy_dtype(self):
        # TF exposes the NumPy scalar *class* as a property; gp_tf.py:126 calls it
        # (``as_numpy_dtype()``), producing a zero scalar whose ``.dtype`` old NumPy accepted as
        # a dtype spec.  NumPy 2 does not, so this is a method returning the class: same meaning.
        return self.np

    def __eq__(self, other):
        if isinstance(other, DType):
            return self.torch == other.torch
        return self.torch == other

    def __hash__(self):
        return hash(self.torch)

    def __repr__(self):
        return f"tf.{self.name}"


float64 = DType("float64", torch.float64, np.float64)
float32 = DType("float32", torch.float32, np.float32)
int64 = DType("int64", torch.int64, np.int64)
int32 = DType("int32", torch.int32, np.int32)
bool = DType("bool", torch.bool, np.bool_)   # noqa: A001  (tf.bool)
_pybool = builtins.bool


def _td(dtype):
    if dtype is None:
        return torch.float64
    if isinstance(dtype, DType):
        return dtype.torch
    return dtype


# --------------------------------------------------------------------------
# session state: what a feed_dict / the graph collections would hold
# --------------------------------------------------------------------------
class _State:
    def __init__(self):
        self.reset()

    def reset(self):
        self.sample_in = None
        self.sample_out = None
        self.condition = True
        self.draws = collections.deque()
        self.draws_taken = 0
        self.draw_provider = None      # callable(body, run, t, in_branch) -> array, see random_normal
        self.draw_log = []
        self.variables = []            # creation order == tf.global_variables()
        self.overrides = None          # list of arrays replacing initial values, creation order
        self.gradients = None          # filled by AdamOptimizer.minimize
        self.adam = None


shim = _State()


def configure(sample_in, sample_out, condition, draws=(), variable_values=None, draw_provider=None):
    """Provide what ``sess.run(fetches, feed_dict)`` would: the minibatch the dataset iterator
    yields (base_model.py:28), the ``condition`` placeholder, the normal draws in execution
    order, and optionally the values of the trainable variables in creation order."""
    shim.reset()
    shim.sample_in = torch.as_tensor(np.asarray(sample_in), dtype=torch.float64)
    shim.sample_out = torch.as_tensor(np.asarray(sample_out), dtype=torch.float64)
    shim.condition = _pybool(condition)
    shim.draws = collections.deque(np.asarray(d, dtype=np.float64) for d in draws)
    shim.draw_provider = draw_provider
    shim.overrides = None if variable_values is None else [np.asarray(v, dtype=np.float64) for v in variable_values]


# torch tensors need TF's static-shape accessor (gp_tf.py:151)
class _StaticShape:
    def __init__(self, shape):
        self._shape = tuple(shape)
        self.ndims = len(self._shape)

    def as_list(self):
        return list(self._shape)


torch.Tensor.get_shape = lambda self: _StaticShape(self.shape)


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(_td(dtype))
    return torch.as_tensor(np.asarray(x), dtype=_td(dtype) if dtype is not None else None)


# --------------------------------------------------------------------------
# graph plumbing
# --------------------------------------------------------------------------
class Graph:
    @contextlib.contextmanager
    def as_default(self):
        yield self


@contextlib.contextmanager
def name_scope(name):
    yield name


class _Placeholder:
    def __init__(self, dtype, shape):
        self.dtype, self.shape = dtype, shape


def placeholder(dtype, shape=None):
    # base_model.py:19-22.  The only placeholder the graph body reads directly is
    # ``condition`` (tf.bool); the data placeholders reach the graph through the iterator.
    if dtype == bool:
        return shim.condition
    return _Placeholder(dtype, shape)


class _Iterator:
    initializer = "iterator.initializer"

    def get_next(self):
        return shim.sample_in, shim.sample_out


class _Dataset:
    @staticmethod
    def from_tensor_slices(tensors):
        return _Dataset()

    def repeat(self, n):
        return self

    def shuffle(self, buf):
        return self

    def batch(self, n):
        return self

    def prefetch(self, buffer_size):
        return self

    def make_initializable_iterator(self):
        return _Iterator()


data = types.SimpleNamespace(Dataset=_Dataset)


class _OutOfRangeError(Exception):
    pass


errors = types.SimpleNamespace(OutOfRangeError=_OutOfRangeError)


def Variable(initial_value, dtype=None, name=None):
    """Trainable variable = float64 autograd leaf, registered in creation order."""
    init = np.asarray(initial_value, dtype=np.float64)
    idx = len(shim.variables)
    if shim.overrides is not None:
        val = shim.overrides[idx]
        if val.shape == () and init.shape == (1,):      # kern variance: backward() returns [1] (tf_transform.py:15)
            val = val.reshape(1)
        assert val.shape == init.shape, f"variable {idx}: override {val.shape} vs initial {init.shape}"
        init = val
    v = torch.tensor(init, dtype=_td(dtype)).requires_grad_(True)
    shim.variables.append(v)
    return v


def global_variables_initializer():
    return "init"


def convert_to_tensor(x, dtype=None):
    return _t(x, dtype)


def constant(value, dtype=None):
    return torch.tensor(value, dtype=_td(dtype))


def cast(x, dtype):
    return _t(x).to(_td(dtype))


def shape(x):
    return tuple(int(s) for s in x.shape)


def zeros(shape, dtype=None):
    return torch.zeros(tuple(int(s) for s in shape), dtype=_td(dtype))


def ones(shape, dtype=None):
    return torch.ones(tuple(int(s) for s in shape), dtype=_td(dtype))


def zeros_like(x):
    return torch.zeros_like(x)


def fill(dims, value):
    return _t(value).reshape(()).expand(*[int(d) for d in dims])


def stack(values, axis=0):
    if all(not isinstance(v, torch.Tensor) or v.dim() == 0 and not v.is_floating_point() for v in values):
        return [int(v) for v in values]          # a shape vector (gp_tf.py:46,141)
    return torch.stack([_t(v) for v in values], dim=axis)


# --------------------------------------------------------------------------
# element-wise / shape ops
# --------------------------------------------------------------------------
def add(a, b):
    return a + b


def multiply(a, b):
    return a * b


def negative(x):
    return -x


def square(x):
    return x * x


def sqrt(x):
    return torch.sqrt(x)


def exp(x):
    return torch.exp(x)


def log(x):
    return torch.log(x)


def abs(x):   # noqa: A001
    return torch.abs(x)


def pow(x, y):   # noqa: A001
    return torch.pow(x, y)


def reciprocal(x):
    return torch.reciprocal(x)


def sin(x):
    return torch.sin(x)


def cos(x):
    return torch.cos(x)


def reduce_sum(x, axis=None):
    return torch.sum(x) if axis is None else torch.sum(x, dim=axis)


def reshape(x, shape):
    return torch.reshape(x, tuple(int(s) for s in shape))


def squeeze(x, axis=None):
    return torch.squeeze(x) if axis is None else torch.squeeze(x, dim=axis)


def expand_dims(x, axis):
    return torch.unsqueeze(x, axis)


def transpose(x, perm=None):
    if perm is None:
        return x.permute(*reversed(range(x.dim())))
    return x.permute(*perm)


def tile(x, multiples):
    return x.repeat(*[int(m) for m in multiples])


def concat(values, axis):
    return torch.cat([_t(v) for v in values], dim=axis)


def reverse(x, axis):
    return torch.flip(x, dims=list(axis))


def matmul(a, b, transpose_a=False, transpose_b=False):
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return a @ b


def mod(a, b):
    return a % b


def equal(a, b):
    return a == b


def logical_or(a, b):
    return _pybool(a) or _pybool(b)


# --------------------------------------------------------------------------
# linear algebra (gp_tf.py:52-54,137,145)
# --------------------------------------------------------------------------
def diag_part(m):
    return torch.diagonal(m)


def matrix_set_diag(m, diag):
    # returns m with its main diagonal replaced by ``diag`` (gradient reaches both)
    eye = torch.eye(m.shape[-1], dtype=m.dtype)
    return m * (1.0 - eye) + torch.diag_embed(diag)


def cholesky(m):
    return torch.linalg.cholesky(m)


def matrix_triangular_solve(matrix, rhs, lower=True, adjoint=False):
    assert not adjoint
    return torch.linalg.solve_triangular(matrix, rhs, upper=not lower)


def norm(x, axis=None):
    return torch.linalg.norm(x) if axis is None else torch.linalg.norm(x, dim=axis)


# --------------------------------------------------------------------------
# control flow (cbfssm.py:107-111,133-136,151,156,176-179,228,233-235)
# --------------------------------------------------------------------------
def while_loop(cond, body, loop_vars, parallel_iterations=10):
    loop_vars = list(loop_vars)
    while _pybool(cond(*loop_vars)):
        loop_vars = list(body(*loop_vars))
    return loop_vars


def cond(pred, true_fn, false_fn):
    # only the taken branch runs; ops created in the other one (e.g. the resample draw,
    # cbfssm.py:134) do not execute and carry no gradient
    return true_fn() if _pybool(pred) else false_fn()


class TensorArray:
    """Write-once array of tensors (``clear_after_read=False``: reads never consume)."""

    def __init__(self, dtype, size, clear_after_read=True):
        self._items = [None] * int(size)

    def unstack(self, value):
        assert value.shape[0] == len(self._items)
        out = TensorArray(None, len(self._items))
        out._items = [value[i] for i in range(value.shape[0])]
        return out

    def read(self, index):
        item = self._items[int(index)]
        assert item is not None, f"TensorArray: read of unwritten index {index}"
        return item

    def write(self, index, value):
        assert self._items[int(index)] is None, f"TensorArray: index {index} written twice"
        out = TensorArray(None, len(self._items))
        out._items = list(self._items)
        out._items[int(index)] = value
        return out

    def stack(self):
        missing = [i for i, it in enumerate(self._items) if it is None]
        assert not missing, f"TensorArray.stack: unwritten indices {missing}"
        return torch.stack(self._items, dim=0)


# --------------------------------------------------------------------------
# randomness (cbfssm.py:134,149,209; cbfssmhalf.py:136)
# --------------------------------------------------------------------------
def _calling_body():
    """Which reference loop body asked for this draw: (function name, its locals ``run`` and ``t``,
    whether the call came from inside a ``tf.cond`` branch lambda)."""
    f = sys._getframe(2)
    in_branch = False
    while f is not None:
        name = f.f_code.co_name
        if name == "<lambda>":
            in_branch = True
        if name in ("_backward_body", "_forward_body"):
            return name, f.f_locals.get("run"), int(f.f_locals["t"]), in_branch
        f = f.f_back
    return None, None, None, in_branch


def random_normal(shape, dtype=None):
    """``tf.random_normal`` with injected values.  Either the next array of the queue given to
    ``configure`` (execution order), or -- when ``shim.draw_provider`` is set -- the array the
    provider returns for (loop body, run, t, inside-a-cond-branch), read from the *reference's own*
    frame: which resample draws are consumed is then decided by the reference's ``tf.cond``
    predicates alone, not by any restated schedule."""
    want = tuple(int(s) for s in shape)
    if shim.draw_provider is not None:
        where = _calling_body()
        d = np.asarray(shim.draw_provider(*where), dtype=np.float64).reshape(want)
        shim.draw_log.append(where)
    else:
        assert shim.draws, "random_normal: the injected queue of draws is exhausted"
        d = shim.draws.popleft()
    assert d.shape == want, f"random_normal: next injected draw has shape {d.shape}, graph asks for {want}"
    shim.draws_taken += 1
    return torch.as_tensor(d, dtype=_td(dtype))


# --------------------------------------------------------------------------
# tf.nn / tf.layers / tf.losses
# --------------------------------------------------------------------------
def _softplus(x):
    return torch.nn.functional.softplus(x, beta=1.0, threshold=1e9)


def _moments(x, axes):
    # TF-1.8 nn.moments: mean, then mean of squared difference to the (stop-gradient) mean
    # = the population variance (cbfssm.py:267,269)
    mean = torch.mean(x, dim=list(axes))
    var = torch.mean((x - torch.mean(x, dim=list(axes), keepdim=True).detach()) ** 2, dim=list(axes))
    return mean, var


def _glorot_uniform(shape):
    limit = math.sqrt(6.0 / (shape[0] + shape[1]))
    return np.random.uniform(-limit, limit, size=shape)


class _GRUCell:
    """TF-1.8 ``tf.nn.rnn_cell.GRUCell`` (cbfssmhalf.py:85): variables gates/kernel, gates/bias
    (constant 1), candidate/kernel, candidate/bias (0), created on first call in that order;
    [r, u] = sigmoid([x, h] Wg + bg); c = tanh([x, r*h] Wc + bc); h' = u*h + (1-u)*c."""

    def __init__(self, num_units):
        self.units = int(num_units)
        self.built = False

    def zero_state(self, batch_size, dtype):
        return torch.zeros((int(batch_size), self.units), dtype=_td(dtype))

    def _build(self, in_dim, dtype):
        n = self.units
        self.gate_kernel = Variable(_glorot_uniform((in_dim + n, 2 * n)), dtype=dtype)
        self.gate_bias = Variable(np.ones(2 * n), dtype=dtype)
        self.cand_kernel = Variable(_glorot_uniform((in_dim + n, n)), dtype=dtype)
        self.cand_bias = Variable(np.zeros(n), dtype=dtype)
        self.built = True

    def __call__(self, inputs, state):
        if not self.built:
            self._build(inputs.shape[1], inputs.dtype)
        n = self.units
        value = torch.sigmoid(torch.cat((inputs, state), dim=1) @ self.gate_kernel + self.gate_bias)
        r, u = value[:, :n], value[:, n:]
        c = torch.tanh(torch.cat((inputs, r * state), dim=1) @ self.cand_kernel + self.cand_bias)
        new_h = u * state + (1.0 - u) * c
        return new_h, new_h


def _dynamic_rnn(cell, inputs, initial_state=None, dtype=None, scope=None):
    # batch-major inputs [B, time, dim] (time_major=False default); returns (outputs, final state)
    state = initial_state
    outs = []
    for t in range(inputs.shape[1]):
        out, state = cell(inputs[:, t], state)
        outs.append(out)
    return torch.stack(outs, dim=1), state


def _dense(inputs, units, activation=None):
    # tf.layers.dense: kernel (glorot uniform) then bias (zeros); cbfssmhalf.py:91
    kernel = Variable(_glorot_uniform((inputs.shape[-1], int(units))), dtype=inputs.dtype)
    bias = Variable(np.zeros(int(units)), dtype=inputs.dtype)
    out = inputs @ kernel + bias
    return activation(out) if activation is not None else out


def _mean_squared_error(labels, predictions):
    # tf.losses.mean_squared_error casts both to float32 (cbfssm.py:270)
    d = predictions.to(torch.float32) - labels.to(torch.float32)
    return torch.mean(d * d)


nn = types.SimpleNamespace(
    softplus=_softplus, moments=_moments, relu=torch.relu,
    rnn_cell=types.SimpleNamespace(GRUCell=_GRUCell), dynamic_rnn=_dynamic_rnn)
layers = types.SimpleNamespace(dense=_dense)
losses = types.SimpleNamespace(mean_squared_error=_mean_squared_error)


# --------------------------------------------------------------------------
# tf.contrib.distributions (gp_tf.py:163-172, cbfssm.py:247-250)
# --------------------------------------------------------------------------
class _MVNDiag:
    def __init__(self, loc, scale_diag):
        self.loc, self.scale_diag = loc, scale_diag

    def log_prob(self, value):
        # TF-1.8 MultivariateNormalLinearOperator: Normal(0,1) log-density of the
        # standardised value summed over the event dim, minus log|det scale|
        z = (value - self.loc) / self.scale_diag
        k = value.shape[-1]
        return (-0.5 * torch.sum(z * z, dim=-1) - 0.5 * k * math.log(2.0 * math.pi)
                - torch.sum(torch.log(torch.abs(self.scale_diag)), dim=-1))


class _MVNTriL:
    def __init__(self, loc, scale_tril):
        self.loc, self.scale_tril = loc, scale_tril


def _kl_divergence(a, b):
    """TF-1.8 ``_kl_brute_force`` (mvn_linear_operator.py) for a = MVNDiag, b = MVNTriL:
    log|det b.scale| - log|det a.scale|
      + 0.5 * ( -n + ||b.scale^-1 a.scale||_F^2 + ||b.scale^-1 (b.mean - a.mean)||^2 ), per batch member."""
    assert isinstance(a, _MVNDiag) and isinstance(b, _MVNTriL)
    n = a.loc.shape[-1]
    a_dense = torch.diag_embed(a.scale_diag)
    b_inv_a = torch.linalg.solve_triangular(b.scale_tril, a_dense, upper=False)
    diff = (b.loc - a.loc).unsqueeze(-1)
    b_inv_d = torch.linalg.solve_triangular(b.scale_tril, diff, upper=False)
    logdet_b = torch.sum(torch.log(torch.abs(torch.diagonal(b.scale_tril, dim1=-2, dim2=-1))), dim=-1)
    logdet_a = torch.sum(torch.log(torch.abs(a.scale_diag)), dim=-1)
    return (logdet_b - logdet_a
            + 0.5 * (-float(n) + torch.sum(b_inv_a * b_inv_a, dim=(-2, -1)) + torch.sum(b_inv_d * b_inv_d, dim=(-2, -1))))


contrib = types.SimpleNamespace(distributions=types.SimpleNamespace(
    MultivariateNormalDiag=_MVNDiag, MultivariateNormalTriL=_MVNTriL, kl_divergence=_kl_divergence))
distributions = types.SimpleNamespace(Beta=None)


# --------------------------------------------------------------------------
# tf.train (cbfssm.py:273-277)
# --------------------------------------------------------------------------
class _AdamOptimizer:
    """``minimize`` = tf.gradients(loss, trainable variables) followed by one TF-1.8 Adam update
    (beta1 .9, beta2 .999, eps 1e-8; lr_t = lr sqrt(1-b2^t)/(1-b1^t); var -= lr_t m/(sqrt(v)+eps)).
    The gradients and the updated values are kept on ``tf.shim`` instead of being applied."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = float(learning_rate), beta1, beta2, epsilon

    def minimize(self, loss):
        vs = list(shim.variables)
        grads = torch.autograd.grad(loss, vs, allow_unused=True, retain_graph=True)
        grads = [g if g is not None else torch.zeros_like(v) for g, v in zip(grads, vs)]
        shim.gradients = grads
        lr_t = self.lr * math.sqrt(1.0 - self.b2) / (1.0 - self.b1)
        new = []
        for v, g in zip(vs, grads):
            m = (1.0 - self.b1) * g
            s = (1.0 - self.b2) * g * g
            new.append((v - lr_t * m / (torch.sqrt(s) + self.eps)).detach())
        shim.adam = new
        return "train_op"


class _Saver:
    pass


train = types.SimpleNamespace(AdamOptimizer=_AdamOptimizer, Saver=_Saver)


class ConfigProto:
    pass
