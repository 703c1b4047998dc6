"""The eager TensorFlow stand-in (oracle/tf_shim) restates TensorFlow-1.8 *library* behaviour; the reference's own
source then runs on top of it (oracle/run_reference.py).  TensorFlow cannot run here, so each restated behaviour is
checked against an independent implementation (NumPy / SciPy / torch.distributions / a hand computation) -- the
semantics listed in SURVEY.md Appendix A."""
import math
import os
import sys

import numpy as np
import pytest
import scipy.linalg
import scipy.special
import torch

SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "tf_shim")
if SHIM not in sys.path:
    sys.path.insert(0, SHIM)
import tensorflow as tf   # noqa: E402  (the shim)

assert tf.__version__.endswith("-shim")
T = lambda a: torch.tensor(np.asarray(a, dtype=np.float64))


def test_softplus_cholesky_triangular_solve_set_diag():
    g = np.random.default_rng(0)
    x = g.standard_normal(50) * 20
    assert np.allclose(tf.nn.softplus(T(x)).numpy(), np.logaddexp(0.0, x), rtol=1e-14)
    A = g.standard_normal((6, 6)); K = A @ A.T + 6 * np.eye(6)
    L = tf.cholesky(T(K)).numpy()
    assert np.allclose(L, np.linalg.cholesky(K), rtol=1e-12) and np.allclose(np.triu(L, 1), 0)
    B = g.standard_normal((6, 4))
    X1 = tf.matrix_triangular_solve(T(L), T(B), lower=True).numpy()
    assert np.allclose(X1, scipy.linalg.solve_triangular(L, B, lower=True), rtol=1e-12)
    X2 = tf.matrix_triangular_solve(tf.transpose(T(L)), T(X1), lower=False).numpy()
    assert np.allclose(X2, np.linalg.solve(K, B), rtol=1e-10)                      # the two solves of gp_tf.py:137,145
    M = tf.matrix_set_diag(T(K), tf.diag_part(T(K)) + 1e-8).numpy()                # gp_tf.py:53
    assert np.allclose(M, K + 1e-8 * np.eye(6), rtol=0, atol=0)


def test_mvn_diag_log_prob_and_kl_against_torch_distributions():
    g = np.random.default_rng(1)
    loc, scale, val = g.standard_normal((3, 5, 4)), g.uniform(0.2, 2.0, (3, 5, 4)), g.standard_normal((3, 5, 4))
    lp = tf.contrib.distributions.MultivariateNormalDiag(loc=T(loc), scale_diag=T(scale)).log_prob(T(val)).numpy()
    ref = torch.distributions.Independent(torch.distributions.Normal(T(loc), T(scale)), 1).log_prob(T(val)).numpy()
    assert lp.shape == (3, 5) and np.allclose(lp, ref, rtol=1e-12)
    # KL(N(m, diag s^2) || N(0, L L^T)), batched over output functions (gp_tf.py:163-172)
    n, d = 7, 3
    A = g.standard_normal((n, n)); L = np.linalg.cholesky(A @ A.T + n * np.eye(n))
    m, s = g.standard_normal((d, n)), g.uniform(0.1, 1.5, (d, n))
    a = tf.contrib.distributions.MultivariateNormalDiag(loc=T(m), scale_diag=T(s))
    b = tf.contrib.distributions.MultivariateNormalTriL(loc=tf.zeros((d, n), dtype=tf.float64),
                                                        scale_tril=tf.tile(tf.expand_dims(T(L), 0), [d, 1, 1]))
    kl = tf.contrib.distributions.kl_divergence(a, b).numpy()
    for i in range(d):
        q = torch.distributions.MultivariateNormal(T(m[i]), covariance_matrix=torch.diag(T(s[i] ** 2)))
        p = torch.distributions.MultivariateNormal(torch.zeros(n, dtype=torch.float64), scale_tril=T(L))
        assert kl[i] == pytest.approx(float(torch.distributions.kl_divergence(q, p)), rel=1e-10)


def test_moments_mse_tile_concat_transpose_shapes():
    g = np.random.default_rng(2)
    x = g.standard_normal((2, 3, 11, 4))
    mean, var = tf.nn.moments(T(x), axes=[2])
    assert np.allclose(mean.numpy(), x.mean(axis=2)) and np.allclose(var.numpy(), x.var(axis=2))      # population variance
    lab, pred = g.standard_normal((2, 3, 4)), g.standard_normal((2, 3, 4))
    mse = float(tf.losses.mean_squared_error(labels=T(lab), predictions=T(pred)))
    assert mse == pytest.approx(np.mean((lab.astype(np.float32) - pred.astype(np.float32)) ** 2), rel=1e-6)
    y = tf.tile(tf.expand_dims(T(x[:, :, 0, :]), axis=2), [1, 1, 5, 1])
    assert tuple(y.shape) == (2, 3, 5, 4) and np.array_equal(y.numpy()[:, :, 3], x[:, :, 0, :])
    z = tf.transpose(T(x), perm=[1, 0, 2, 3])
    assert tuple(z.shape) == (3, 2, 11, 4) and np.array_equal(z.numpy()[1, 0], x[0, 1])
    c = tf.concat((T(x), T(x[..., :2])), axis=3)
    assert tuple(c.shape) == (2, 3, 11, 6)
    assert tf.stack([tf.shape(T(x))[0]]) == [2] and tuple(tf.fill(tf.stack([4]), tf.squeeze(T([0.3]))).shape) == (4,)


def test_control_flow_tensor_array_and_random_normal_queue():
    tf.configure(np.zeros((1, 2, 1)), np.zeros((1, 2, 1)), True, draws=[np.full((2, 1), 1.5), np.full((2, 1), -2.0)])
    ta = tf.TensorArray(dtype=tf.float64, size=3, clear_after_read=False)
    ta = ta.write(0, T([1.0]))
    ran = []
    out = tf.cond(tf.equal(tf.mod(5 + 1, 6), 0), lambda: ran.append("t") or ta.write(1, T([2.0])), lambda: ran.append("f") or ta)
    assert ran == ["t"]                                                 # only the taken branch executes
    with pytest.raises(AssertionError):
        out.write(1, T([9.0]))                                          # write-once, like TF's TensorArray
    with pytest.raises(AssertionError):
        out.stack()                                                     # index 2 never written
    i, acc = tf.while_loop(lambda i, a: i >= 0, lambda i, a: (i - 1, a + i), [3, 0], parallel_iterations=1)
    assert (i, acc) == (-1, 6)
    a = tf.random_normal((2, 1), dtype=tf.float64)
    b = tf.random_normal((2, 1), dtype=tf.float64)
    assert float(a[0, 0]) == 1.5 and float(b[1, 0]) == -2.0
    with pytest.raises(AssertionError):
        tf.random_normal((2, 1), dtype=tf.float64)                      # queue exhausted: no silent fresh draws


def test_gradients_and_adam_first_step():
    tf.configure(np.zeros((1, 2, 1)), np.zeros((1, 2, 1)), True)
    v = tf.Variable(np.array([0.5, -1.0, 2.0]), dtype=tf.float64)
    w = tf.Variable(np.array([[1.0, 2.0]]), dtype=tf.float64)
    loss = tf.reduce_sum(tf.square(v) * 3.0) + tf.reduce_sum(tf.exp(w))
    tf.train.AdamOptimizer(learning_rate=0.1).minimize(loss)
    gv, gw = (x.numpy() for x in tf.shim.gradients)
    assert np.allclose(gv, 6.0 * np.array([0.5, -1.0, 2.0])) and np.allclose(gw, np.exp([[1.0, 2.0]]))
    lr_t = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)                       # TF: eps outside the bias-corrected root
    exp_v = np.array([0.5, -1.0, 2.0]) - lr_t * (0.1 * gv) / (np.sqrt(0.001 * gv * gv) + 1e-8)
    assert np.allclose(tf.shim.adam[0].numpy(), exp_v, rtol=1e-12)


def test_gru_cell_and_dense_follow_tf_1_8():
    """[r, u] = sigmoid([x, h] Wg + bg); c = tanh([x, r*h] Wc + bc); h' = u*h + (1-u)*c; variables created on first
    call in the order gates kernel, gates bias (ones), candidate kernel, candidate bias (zeros); then dense."""
    g = np.random.default_rng(3)
    B, din, nh = 2, 3, 16
    vals = [g.standard_normal((din + nh, 2 * nh)) * 0.3, np.ones(2 * nh), g.standard_normal((din + nh, nh)) * 0.3,
            np.zeros(nh), g.standard_normal((nh, 4)) * 0.3, g.standard_normal(4) * 0.1]
    tf.configure(np.zeros((1, 2, 1)), np.zeros((1, 2, 1)), True, variable_values=vals)
    cell = tf.nn.rnn_cell.GRUCell(nh)
    x = g.standard_normal((B, 5, din))
    _, state = tf.nn.dynamic_rnn(cell, T(x), initial_state=cell.zero_state(B, dtype=tf.float64), dtype=tf.float64)
    out = tf.layers.dense(state, 4).detach().numpy()
    assert [tuple(v.shape) for v in tf.shim.variables] == [(din + nh, 2 * nh), (2 * nh,), (din + nh, nh), (nh,), (nh, 4), (4,)]
    h = np.zeros((B, nh))
    for t in range(5):
        gates = scipy.special.expit(np.concatenate((x[:, t], h), 1) @ vals[0] + vals[1])
        r, u = gates[:, :nh], gates[:, nh:]
        c = np.tanh(np.concatenate((x[:, t], r * h), 1) @ vals[2] + vals[3])
        h = u * h + (1 - u) * c
    assert np.allclose(out, h @ vals[4] + vals[5], rtol=1e-12)
