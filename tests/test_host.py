"""CPU tests of the host side: C-ABI surface, dataset windows, minibatch iterator."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import kernel_math as KM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from cbf_ssm_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from cbf_ssm_b200 import _lib
    header = open(os.path.join(ROOT, "include", "cbfssm_b200.h")).read()
    declared = set(re.findall(r"CBF_API\s+[\w\s\*]+?\b(cbf_\w+)\s*\(", header))
    assert len(declared) >= 18
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.cbf_abi_version() == 2


def _shape(**kw):
    from cbf_ssm_b200._lib import cbf_shape
    d = dict(B=4, S=10, T=30, M=20, dx=4, du=2, dy=2, R=8, condition=1, n_offset=0, n_local=40, k_factor=1.0, flags=0)
    d.update(kw)
    return cbf_shape(*(d[f] for f, _ in cbf_shape._fields_))


def test_host_only_entry_points_without_gpu(lib):
    from cbf_ssm_b200._lib import cbf_grad_layout
    assert lib.cbf_supported(20, 4, 2, 2) == 2          # register-resident instantiation
    assert lib.cbf_supported(100, 4, 1, 1) == 1
    assert lib.cbf_supported(100, 14, 7, 7) == 1
    assert lib.cbf_supported(20, 5, 5, 5) == 0          # dy >= dx is not a model
    assert lib.cbf_supported(20, 5, 5, 2) == 3          # dims not compiled in: the float64 batched path takes them
    assert lib.cbf_supported(500, 4, 2, 2) == 3         # resident set exceeds one SM: float64 batched path
    assert lib.cbf_supported(500, 8, 1, 4) == 3 and lib.cbf_supported(129, 4, 2, 2) == 3
    assert lib.cbf_supported(128, 4, 2, 2) == 1         # too large for the cooperative kernels: tensor path only
    assert lib.cbf_supported(128, 14, 7, 7) == 1
    assert lib.cbf_supported(20, 17, 2, 5) == 0 and lib.cbf_supported(20, 16, 16, 5) == 0   # beyond the float64 path's dims
    n = C.c_size_t(0)
    s = _shape()
    assert lib.cbf_workspace_bytes(C.byref(s), C.byref(n)) == 0 and n.value > 0
    gl = cbf_grad_layout()
    assert lib.cbf_grad_layout_get(C.byref(s), C.byref(gl)) == 0
    M, din, dx, dh = 20, 6, 4, 2
    assert gl.total == 2 * (M * M + M * din + din + 1) + 2 * M * dx + 2 * M * dh + 2 * dx
    assert lib.cbf_gp_prologue_state_doubles(20, 6, 4) > 2 * 400


@pytest.mark.parametrize("bad", [dict(T=0), dict(dy=4), dict(n_local=41), dict(R=0), dict(M=0), dict(n_offset=-1)])
def test_invalid_shapes_are_rejected(lib, bad):
    n = C.c_size_t(0)
    s = _shape(**bad)
    assert lib.cbf_workspace_bytes(C.byref(s), C.byref(n)) == -1
    assert b"invalid shape" in lib.cbf_last_error_string()


def test_null_arguments_are_rejected(lib):
    assert lib.cbf_workspace_bytes(None, None) == -5
    assert lib.cbf_adam_step(4, None, None, None, None, 1, 0.01, 0.9, 0.999, 1e-8, None) == -5


def test_missing_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cbf_ssm_b200.engine import ElboEngine, ModelDims
    with pytest.raises(RuntimeError, match="no CPU path"):
        ElboEngine(ModelDims(4, 2, 2, 20, 5, 4))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cbf_ssm_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


@pytest.mark.parametrize("T,R", [(100, 50), (300, 50), (250, 16), (6, 50), (5000, 16), (33, 1)])
def test_chain_launch_count_matches_chain_table(T, R):
    from cbf_ssm_b200.engine import count_chain_batches
    n = len(KM.build_chains(T, R))
    assert count_chain_batches(T, R) == (0 if n == 0 else -(-n // 120))


def test_window_counts_of_the_named_datasets():
    """base_ds.py:54-77 semantics; counts cross-checked against the reference in SURVEY 8c."""
    from cbf_ssm_b200.datasets import BaseDS
    x = np.arange(5000 * 1, dtype=float).reshape(1, 5000, 1)
    assert BaseDS.rnn_batches(x, 100, 50).shape == (99, 100, 1)           # SpringNonlinear
    x = np.zeros((1, 25000, 2))
    assert BaseDS.rnn_batches(x, 300, 50).shape[0] == 495                 # RoboMove
    x = np.zeros((60, 337, 7))
    assert BaseDS.rnn_batches(x, 250, 10).shape[0] == 600                 # Sarcos
    # the remainder window holds the last samples
    x = np.arange(23, dtype=float).reshape(1, 23, 1)
    w = BaseDS.rnn_batches(x, 10, 4)
    assert w.shape[0] == 5 and w[-1, -1, 0] == 22 and w[-2, 0, 0] == 12
    with pytest.raises(AssertionError):
        BaseDS.rnn_batches(np.zeros((1, 5, 1)), 10, 1)


def test_normalisation_round_trip():
    from cbf_ssm_b200.datasets import SpringNonlinearSynthetic
    ds = SpringNonlinearSynthetic(100, 50, seed=0)
    assert ds.train_in_batch.shape == (99, 100, 1) and ds.test_out_batch.shape == (99, 100, 1)
    assert abs(ds.train_out.mean()) < 1e-9 and abs(ds.train_out.std() - 1) < 1e-9
    z = ds.normalize(np.array([[0.3]]), 'out')
    assert np.allclose(ds.denormalize(z, 'out'), 0.3)


class _FakeModel:
    """BaseModel with the device work replaced: returns the batch it was given."""

    def __new__(cls, config):
        from cbf_ssm_b200.model.base_model import BaseModel

        class M(BaseModel):
            def _build_graph(self):
                self.loss = self._handle("loss")
                self.idx = self._handle("idx")

            def _session_run(self, fetches, feed_dict):
                u, y = self._next_batch()
                return (np.float64(u.sum()), u[:, 0, 0].copy())
        return M(config)


def test_load_ds_and_run_follow_the_reference_iterator_semantics():
    from cbf_ssm_b200.model.base_model import Session

    class DS:
        dim_u, dim_y = 1, 1
    n = 10
    data = np.arange(n, dtype=float).reshape(n, 1, 1) * np.ones((1, 3, 1))
    model = _FakeModel({"ds": DS, "batch_size": 4, "shuffle": 10000, "shuffle_seed": 0})
    sess = Session(model)
    model.load_ds(sess, data, data)
    loss, idx = model.run(sess, (model.loss, model.idx), {model.condition: True})
    assert loss.shape == (3,)                       # ceil(10/4) batches, scalars -> vector (base_model.py:57)
    assert sorted(idx.tolist()) == list(range(n))   # every window exactly once, last batch short
    assert idx.tolist() != list(range(n))           # shuffled
    model.load_ds(sess, data, data, repeats=2)
    _, idx2 = model.run(sess, (model.loss, model.idx), {})
    assert sorted(idx2.tolist()) == sorted(list(range(n)) * 2)
    model2 = _FakeModel({"ds": DS, "batch_size": 4, "shuffle": 1})
    model2.load_ds(Session(model2), data, data)
    _, idx3 = model2.run(Session(model2), (model2.loss, model2.idx), {})
    assert idx3.tolist() == list(range(n))          # buffer of 1 = no shuffle
    model3 = _FakeModel({"ds": DS, "batch_size": 4, "shuffle": 3, "shuffle_seed": 1})
    model3.load_ds(Session(model3), data, data)
    _, idx4 = model3.run(Session(model3), (model3.loss, model3.idx), {})
    assert sorted(idx4.tolist()) == list(range(n))
    assert all(v <= i + 2 for i, v in enumerate(idx4.tolist()))   # bounded look-ahead of a size-3 buffer


# --------------------------------------------------------------------------------------
# on-disk dataset format: ds_manager.py:11-34, dsmanager_ds.py:6-63
# --------------------------------------------------------------------------------------
def test_mat_files_round_trip_and_feed_the_dataset_classes(tmp_path):
    import scipy.io
    from cbf_ssm_b200.datasets import (DSManager, RoboMove, RoboMoveSimple, SpringNonlinear, create_robomove,
                                       create_spring_nonlinear)
    d = str(tmp_path) + "/"
    u, x, y = create_spring_nonlinear(d + "spring_nonlinear.mat", seed=3)
    assert u.shape == (10000, 1) and x.shape == (10000, 3) and y.shape == (10000, 3)
    raw = scipy.io.loadmat(d + "spring_nonlinear.mat")           # the keys the reference reads
    assert {"ds_u", "ds_x", "ds_y", "title"} <= set(raw) and raw["ds_u"].dtype == np.float64
    u2, x2, y2 = DSManager.load_ds(d + "spring_nonlinear.mat", print_title=False)
    assert np.array_equal(u, u2) and np.array_equal(x, x2) and np.array_equal(y, y2)
    # x_{i+1} = f(x_i, u_i): the spring's linear recursion holds between consecutive rows
    A = np.array([[1.0, 0.01, 0.0], [0.0, 1.0, 0.01], [-500.0, -25.0, 0.0]])
    assert np.allclose(x[1:], x[:-1] @ A.T + np.outer(np.tanh(2 * u[:-1, 0]), [0, 0, 500.0]), atol=1e-9)
    un = DSManager.normalize_ds(u)
    assert np.allclose(un.mean(0), 0, atol=1e-12) and np.allclose(un.std(0), 1)

    ds = SpringNonlinear(100, 50, data_path=d)                  # y_crop=1, split 5000, stride 50 -> 99 windows
    assert ds.train_in.shape == (1, 5000, 1) and ds.train_out.shape == (1, 5000, 1)
    assert ds.train_in_batch.shape == (99, 100, 1) and ds.test_out_batch.shape == (99, 100, 1)
    whole = np.concatenate((ds.train_out[0], ds.test_out[0]))   # statistics are those of the whole file
    assert np.allclose(whole.mean(0), 0, atol=1e-9) and np.allclose(whole.std(0), 1)
    assert np.allclose(ds.denormalize(ds.train_out[0], 'out')[:, 0], y[:5000, 0])

    create_robomove(d + "robomove.mat", ds_size=30000, seed=1)
    rm = RoboMove(300, 50, data_path=d)
    assert (rm.dim_u, rm.dim_y) == (2, 2) and rm.train_in_batch.shape == (495, 300, 2)   # SURVEY 8d cfg 2
    assert rm.test_in_batch.shape == (95, 300, 2)
    create_robomove(d + "robomove_simple.mat", ds_size=26000, seed=2, simple=True)
    rs = RoboMoveSimple(300, 50, data_path=d)
    assert rs.train_out_batch.shape == (495, 300, 4) and rs.test_out.shape == (1, 1000, 4)
    with pytest.raises(AssertionError):
        DSManager.save_ds(d + "bad.mat", u[:5], x, y, "bad")


def test_unicycle_generator_obeys_its_kinematics():
    """create_robomove.py:22-52: a step of length s on a circle of curvature c turns the heading by s*c and
    moves the robot along the chord; straight motion for |c| < 1e-5."""
    from cbf_ssm_b200.datasets.mat_ds import _Unicycle
    rob = _Unicycle(np.random.default_rng(0), 0.0, 0.0, simple=False)
    rob.propagate([0.5, 0.0])
    assert np.allclose(rob.get_state(), [0.0, 0.5, 0.0])        # heading 0 points along +y
    start = rob.pos.copy()
    for _ in range(8):                                           # 8 steps of an eighth of a unit circle
        rob.propagate([2 * np.pi / 8, 1.0])
    assert np.allclose(rob.pos, start, atol=1e-12) and rob.angle == pytest.approx(0.0, abs=1e-12)
    rob.propagate([np.pi / 2, 1.0])                              # quarter turn to the right
    assert np.allclose(rob.pos - start, [1.0, 1.0]) and rob.angle == pytest.approx(np.pi / 2)


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU restatement on the host cores) prints one JSON line with the keys the
    driver reads; runs without a GPU."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--workload", "spring_template_b32"], capture_output=True, text=True,
                         timeout=600, check=True).stdout
    line = json.loads(out.strip().splitlines()[-1])
    # the CPU arm prints the same workload description the GPU arm prints (bench.config_dict), honours --steps /
    # --warmup as given, and adds the reference's own thread setting (trainer.py:24) and its rate-vs-batch sweep
    import bench
    w = bench.WORKLOADS["spring_template_b32"]
    assert line["config"] == bench.config_dict(w, w["batch"], 1, "weak") and line["steps"] == 2 and line["warmup"] == 1
    assert line["cpu_baseline"]["five_threads"]["threads"] == min(5, os.cpu_count()) and line["cpu_baseline"]["rate_vs_sample_batch"]
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "particle-steps/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["vs_baseline"] is None and line["gpu_launches"] == 0


def test_output_summary_writes_the_reference_layout(tmp_path):
    """summary.txt of repeated runs (cbfssm/outputs/output_summary.py:19-31): runs, mean, population std."""
    from cbf_ssm_b200.outputs import OutputSummary

    class FakeOutputs:
        def __init__(self, v):
            self.v = v

        def get_last_rmse(self):
            return self.v
    summ = OutputSummary(str(tmp_path / "s"), copy_main=False)
    for v in (0.5, 0.7, 0.9):
        summ.add_outputs(FakeOutputs(v))
    path = summ.write_summary()
    text = open(path).read().split("\n")
    assert text[:4] == ["RMSE", "====", "", "Runs:"] and text[4:7] == ["  0.500000", "  0.700000", "  0.900000"]
    assert text[7] == "Mean: 0.700000" and text[8] == "Std:  %f" % np.std([0.5, 0.7, 0.9])
    empty = OutputSummary(str(tmp_path / "e"), copy_main=False)
    empty.add_outputs(FakeOutputs(None))
    assert empty.write_summary() is None and not os.path.exists(tmp_path / "e" / "summary.txt")


def test_tf_checkpoint_bundle_round_trip_and_corruption(tmp_path):
    """tf.train.Saver's V2 bundle (.index SSTable + .data shard) written and read back without TensorFlow: many
    tensors (several table blocks, prefix-compressed keys), scalars, float32 / float64 / int64, checksums."""
    from cbf_ssm_b200.training import read_tf_checkpoint, reference_variable_names, write_tf_checkpoint
    from cbf_ssm_b200.training.tf_checkpoint import crc32c
    assert crc32c(b"123456789") == 0xE3069283                    # the CRC-32C check value
    g = np.random.default_rng(0)
    tensors = {"beta1_power": np.asarray(0.81), "global_step": np.asarray(7, dtype=np.int64)}
    for i in range(150):
        base = "Variable" if i == 0 else "Variable_%d" % i
        shape = [(20, 6), (20, 4), (1,), (6,), ()][i % 5]
        tensors[base] = g.standard_normal(shape)
        tensors[base + "/Adam"] = g.standard_normal(shape).astype(np.float32)
        tensors[base + "/Adam_1"] = g.standard_normal(shape)
    tensors["kern/Variable"] = g.standard_normal((1,))
    tensors["kern_1/Variable_1"] = g.standard_normal((6,))
    prefix = str(tmp_path / "model.ckpt")
    write_tf_checkpoint(prefix, tensors, block_size=512)
    back = read_tf_checkpoint(prefix)
    assert set(back) == set(tensors)
    for k, v in tensors.items():
        assert back[k].dtype == np.asarray(v).dtype and back[k].shape == np.asarray(v).shape and np.array_equal(back[k], v), k
    # one flipped byte in the data shard / in the index is detected
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[100] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError, match="checksum"):
        read_tf_checkpoint(prefix)
    data[100] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[40] ^= 1
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError):
        read_tf_checkpoint(prefix)
    # the reference graph's default variable names, in creation order (gp_tf.py:112-127, cbfssm.py:30-54)
    names = reference_variable_names()
    assert [names[k] for k in ("f.zeta_pos", "f.zeta_mean", "f.zeta_var_unc", "f.variance_unc", "f.lengthscales_unc")] == \
        ["Variable", "Variable_1", "Variable_2", "kern/Variable", "kern/Variable_1"]
    assert [names[k] for k in ("b.zeta_pos", "b.variance_unc", "var_x_unc", "var_y_unc")] == \
        ["Variable_3", "kern_1/Variable", "Variable_6", "Variable_7"]
    assert reference_variable_names(half=True)["var_y_unc"] == "Variable_4"
