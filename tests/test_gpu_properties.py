"""Size-independent properties, edge cases and the callers either side of the hot path,
on the GPU through the C ABI / the reference-shaped Python interface."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import cbfssm_oracle as O
from tests.helpers import make_problem, rel_inf
from tests.test_gpu_parity import run_engine

pytestmark = pytest.mark.gpu


def _engine(dx=4, du=2, dy=2, M=20, S=50, R=50, kap=1.0, lf=(20.0, 0.0), seed=1):
    from cbf_ssm_b200.engine import ElboEngine, ModelDims, init_param_arrays
    dims = ModelDims(dx, du, dy, M, S, R, kap, lf)
    eng = ElboEngine(dims)
    cfg = dict(zeta_pos=2.0, zeta_mean=0.01, zeta_var=1e-4, gp_var=0.01, gp_len=1.0,
               var_x=np.full(dx, 0.01), var_y=np.full(dx, 1.0))
    eng.set_params(init_param_arrays(dims, cfg, seed))
    return eng


def _inputs(eng, B, T, seed=0):
    d, dev = eng.dims, eng.device
    g = torch.Generator(device="cpu").manual_seed(seed)
    u = torch.randn(B, T, d.dim_u, generator=g).to(dev)
    y = torch.randn(B, T, d.dim_y, generator=g).to(dev)
    N = B * d.samples
    eb = torch.randn(2, T, N, generator=g).to(dev)
    zb = torch.randn(2, T, N, generator=g).to(dev)
    ef = torch.randn(max(T - 1, 1), N, generator=g).to(dev)
    return u, y, eb, zb, ef


@pytest.mark.parametrize("flags", [12, 1], ids=["register_or_tensor", "cooperative"])
def test_sum_of_particle_shards_equals_the_whole_at_bench_shape(flags):
    """Data-parallel contract at the BASELINE shape (M=20, T=300, S=50): the
    kernel-level gradient and the ELBO terms of contiguous particle shards add up to the
    unsharded result (fp32 summation-order tolerance); also bit-exact determinism."""
    eng = _engine()
    eng.flags = flags
    B, T = 64, 300
    u, y, eb, zb, ef = _inputs(eng, B, T)
    N = B * eng.dims.samples
    eng.forward(u, y, eb, zb, ef, True)
    eng.backward()
    whole_terms = eng.terms[:3].clone()
    whole = eng._gflat.clone()
    eng.forward(u, y, eb, zb, ef, True)
    eng.backward()
    assert torch.equal(whole, eng._gflat) and torch.equal(whole_terms, eng.terms[:3])   # deterministic reductions
    acc = torch.zeros_like(whole)
    tacc = torch.zeros_like(whole_terms)
    bounds = [0, 1000, 1031, 2400, N]          # ragged shards, not multiples of 32 / 128 / S
    for n0, n1 in zip(bounds[:-1], bounds[1:]):
        sl = slice(n0, n1)
        eng.forward(u, y, eb[:, :, sl].contiguous(), zb[:, :, sl].contiguous(), ef[:, sl].contiguous(), True,
                    n_offset=n0, n_local=n1 - n0)
        eng.backward()
        acc += eng._gflat
        tacc += eng.terms[:3]
    torch.cuda.synchronize()
    gl = eng._gl.total
    assert rel_inf(tacc.cpu().numpy(), whole_terms.cpu().numpy()) < 1e-6
    assert rel_inf(acc[:gl].cpu().numpy(), whole[:gl].cpu().numpy()) < 2e-5


@pytest.mark.parametrize("dims", [(4, 2, 2), (8, 1, 4)], ids=["dx4_one_tile_ctas", "dx8_two_tile_ctas"])
def test_tensor_path_shards_add_up(dims, monkeypatch):
    """The same contract on the tcgen05 path (M=100): ragged particle shards with n_offset != 0, the reverse
    pass in several time windows, one- and two-tile CTAs."""
    monkeypatch.setenv("CBFSSM_B200_TC_WINDOW_BYTES", str(40_000_000))
    dx, du, dy = dims
    eng = _engine(dx=dx, du=du, dy=dy, M=100, S=50, R=10, lf=(10.0, 0.5))
    eng.flags = 12
    B, T = 24, 40
    u, y, eb, zb, ef = _inputs(eng, B, T)
    N = B * eng.dims.samples
    eng.forward(u, y, eb, zb, ef, True)
    eng.backward()
    whole_terms, whole = eng.terms[:3].clone(), eng._gflat.clone()
    acc, tacc = torch.zeros_like(whole), torch.zeros_like(whole_terms)
    bounds = [0, 130, 517, 1000, N]            # shards that are not multiples of the 128-particle tile
    for n0, n1 in zip(bounds[:-1], bounds[1:]):
        sl = slice(n0, n1)
        eng.forward(u, y, eb[:, :, sl].contiguous(), zb[:, :, sl].contiguous(), ef[:, sl].contiguous(), True,
                    n_offset=n0, n_local=n1 - n0)
        eng.backward()
        acc += eng._gflat
        tacc += eng.terms[:3]
    torch.cuda.synchronize()
    gl = eng._gl.total
    assert rel_inf(tacc.cpu().numpy(), whole_terms.cpu().numpy()) < 1e-6
    assert rel_inf(acc[:gl].cpu().numpy(), whole[:gl].cpu().numpy()) < 5e-5


def test_loss_is_linear_in_the_loss_factors():
    """loss = -(l1/S)(loglik - kl_x) - (l2/S) entropy + kl_z  (cbfssm.py:257-262)."""
    e1 = _engine(lf=(20.0, 0.0))
    u, y, eb, zb, ef = _inputs(e1, 8, 60)
    o1 = {k: float(v) for k, v in e1.forward(u, y, eb, zb, ef, True).items()}
    e2 = _engine(lf=(5.0, 3.0))
    o2 = {k: float(v) for k, v in e2.forward(u, y, eb, zb, ef, True).items()}
    for k in ("loglik", "kl_x", "entropy", "kl_z_f", "kl_z_b"):
        assert o1[k] == o2[k]
    S = 50.0
    assert o2["loss"] == pytest.approx(-(5 / S) * (o2["loglik"] - o2["kl_x"]) - (3 / S) * o2["entropy"]
                                       + o2["kl_z_f"] + o2["kl_z_b"], rel=1e-12)


EDGE = [
    # dx du dy  M  S  B  T   R   cond    what
    (4, 2, 2, 20, 1, 1, 2, 50, True),      # single particle, two steps, T < R
    (4, 2, 2, 20, 3, 1, 1, 4, True),       # T = 1: no forward step, no eps_f
    (4, 2, 2, 20, 33, 1, 17, 1, True),     # R = 1: resample every second step
    (4, 2, 2, 20, 129, 1, 9, 4, False),    # 129 particles: one past a CTA tile
    (4, 1, 1, 100, 5, 3, 8, 8, True),      # T == R
    (3, 1, 1, 12, 31, 1, 16, 8, False),    # T == 2R, free-running prediction after R-1 steps
]


@pytest.mark.parametrize("flags", [12, 1, 0], ids=["register_or_tensor", "cooperative", "default"])
@pytest.mark.parametrize("case", EDGE, ids=lambda c: "M%d_S%d_B%d_T%d_R%d_c%d" % (c[3], c[4], c[5], c[6], c[7], c[8]))
def test_edge_shapes_match_oracle(case, flags):
    dx, du, dy, M, S, B, T, R, cond = case
    cfg, params, u, y, eb, zb, ef = make_problem(dx, du, dy, M, S, B, T, R, 1.0, (10.0, 0.5), seed=21, strong=True)
    res, gd = O.loss_and_grads(cfg, params, u, y, eb, zb, ef, cond)
    eng, out, _ = run_engine(cfg, params, u, y, eb, zb, ef, cond, flags)
    assert float(out["loss"]) == pytest.approx(float(res.loss.detach()), rel=1e-4)
    grads = eng.get_grads()
    bad = {k: rel_inf(grads[k], gd[k].numpy()) for k in O.PARAM_NAMES}
    bad = {k: v for k, v in bad.items() if not v < 1e-4 and np.max(np.abs(gd[k].numpy())) > 0}
    assert not bad, bad


def test_unsupported_shapes_fail_loudly():
    from cbf_ssm_b200._lib import CbfError
    from cbf_ssm_b200.engine import ElboEngine, ModelDims
    with pytest.raises(CbfError):
        ElboEngine(ModelDims(17, 2, 5, 20, 5, 4))       # beyond every path (float64 path: dim_x <= 16)
    with pytest.raises(CbfError):
        ElboEngine(ModelDims(4, 40, 2, 20, 5, 4))       # dim_x + dim_u > 31
    assert ElboEngine(ModelDims(4, 2, 2, 500, 5, 4)).kernel_path == 3      # M = 500: float64 batched path
    assert ElboEngine(ModelDims(5, 5, 2, 20, 5, 4)).kernel_path == 3       # dims without an instantiation


def test_adam_step_matches_tf_formula():
    eng = _engine(M=7)
    g = torch.Generator().manual_seed(3)
    theta0 = eng.theta.cpu().clone()
    m = torch.zeros_like(theta0)
    v = torch.zeros_like(theta0)
    ref = theta0.clone()
    for step in (1, 2, 3):
        grad = torch.randn(theta0.numel(), generator=g, dtype=torch.float64)
        eng.grad.copy_(grad)
        eng.adam_step(0.05)
        ref, m, v = O.adam_step_tf(ref, grad, m, v, step, 0.05)
    torch.cuda.synchronize()
    assert torch.allclose(eng.theta.cpu(), ref, rtol=1e-13, atol=1e-15)


def test_philox_normals_are_standard_and_reproducible():
    eng = _engine(M=7)
    a = torch.empty(1 << 20, device=eng.device)
    b = torch.empty_like(a)
    eng.fill_normal(a, 7, 0)
    eng.fill_normal(b, 7, 0)
    assert torch.equal(a, b)
    eng.fill_normal(b, 7, 1)
    assert not torch.equal(a, b)
    x = a.double().cpu().numpy()
    assert abs(x.mean()) < 5e-3 and abs(x.std() - 1) < 5e-3
    assert abs(np.mean(x ** 3)) < 2e-2 and abs(np.mean(x ** 4) - 3) < 5e-2
    assert abs(np.corrcoef(x, b.double().cpu().numpy())[0, 1]) < 5e-3
    c = torch.empty(1003, device=eng.device)            # ragged tail
    eng.fill_normal(c, 7, 0)
    assert torch.equal(c, a[:1003])


def test_model_and_trainer_follow_the_reference_interface(tmp_path):
    """run/template.py flow on a tiny problem: CBFSSM(config), Trainer.train(ds, epochs),
    retrain from model.ckpt, prediction handles with condition=False."""
    from cbf_ssm_b200.datasets import SpringNonlinearSynthetic
    from cbf_ssm_b200.model import CBFSSM, Session
    from cbf_ssm_b200.training import Trainer

    class SmallSpring(SpringNonlinearSynthetic):
        exp_len = 600
    ds = SmallSpring(40, 20, seed=0)
    dim_x = 4
    config = {'ds': SmallSpring, 'batch_size': 8, 'shuffle': 10000, 'dim_x': dim_x, 'ind_pnt_num': 20, 'samples': 16,
              'learning_rate': 0.01, 'loss_factors': np.asarray([10., 0.]), 'k_factor': 1., 'recog_len': 10,
              'zeta_pos': 2., 'zeta_mean': 0.1 ** 2, 'zeta_var': 0.01 ** 2, 'var_x': np.asarray([0.1 ** 2] * dim_x),
              'var_y': np.asarray([1. ** 2] * dim_x), 'gp_var': 0.1 ** 2, 'gp_len': 1., 'shuffle_seed': 0}
    model = CBFSSM(config, seed=0)
    trainer = Trainer(model, str(tmp_path))
    trainer.train(ds, 4, verbose=False)
    assert len(trainer.train_all) == 4 and len(trainer.test_all) == 4
    assert trainer.train_all[-1] < trainer.train_all[0]            # Adam on the ELBO makes progress
    assert os.path.exists(tmp_path / "best.ckpt") and os.path.exists(tmp_path / "model.ckpt")
    theta = model.engine.theta.clone()
    step = model.engine.adam_t
    trainer.train(ds, 1, retrain=True, verbose=False)              # curriculum restart (run_robomove.py:47)
    assert model.engine.adam_t > step and not torch.equal(theta, model.engine.theta)
    # prediction path of cbfssm/outputs/outputs.py:68-71
    sess = Session(model)
    model.load_ds(sess, ds.test_in_batch[:1], ds.test_out_batch[:1])
    pm, pv = sess.run((model.pred_mean, model.pred_var), {model.condition: False})
    assert pm.shape == (1, 40, 1) and pv.shape == (1, 40, 1) and np.all(pv > 0)
    noise = sess.run(model.var_dict['observation noise'])
    assert noise.shape == (dim_x,) and np.all(noise > 0)
    model.load_ds(sess, ds.test_in_batch, ds.test_out_batch)
    losses = model.run(sess, model.loss, {model.condition: True})
    assert losses[0].shape == (-(-ds.test_in_batch.shape[0] // 8),)   # list of per-fetch arrays (base_model.py:54)


def test_outputs_writes_the_reference_report_files(tmp_path):
    """cbfssm/outputs/outputs.py:36-164 on a tiny .mat-backed problem: predict_*.mat, mse.txt, var_dump.txt."""
    import scipy.io
    from cbf_ssm_b200.datasets import SpringNonlinear, create_spring_nonlinear
    from cbf_ssm_b200.model import CBFSSM
    from cbf_ssm_b200.outputs import Outputs
    from cbf_ssm_b200.training import Trainer

    data = str(tmp_path / "data") + "/"
    os.makedirs(data)
    create_spring_nonlinear(data + "spring_nonlinear.mat", ds_size=10000, seed=0)

    class DS(SpringNonlinear):
        def __init__(self, seq_len, seq_stride):
            super().__init__(seq_len, seq_stride, data_path=data)
    ds = DS(100, 500)
    dim_x = 4
    config = {'ds': DS, 'batch_size': 8, 'shuffle': 10000, 'dim_x': dim_x, 'ind_pnt_num': 20, 'samples': 16,
              'learning_rate': 0.01, 'loss_factors': np.asarray([10., 0.]), 'k_factor': 1., 'recog_len': 10,
              'zeta_pos': 2., 'zeta_mean': 0.1 ** 2, 'zeta_var': 0.01 ** 2, 'var_x': np.asarray([0.1 ** 2] * dim_x),
              'var_y': np.asarray([1. ** 2] * dim_x), 'gp_var': 0.1 ** 2, 'gp_len': 1., 'shuffle_seed': 0}
    model = CBFSSM(config, seed=0)
    model_dir, out_dir = str(tmp_path / "model"), str(tmp_path / "out")
    trainer = Trainer(model, model_dir)
    trainer.train(ds, 2, verbose=False)
    out = Outputs(out_dir)
    out.set_ds(ds)
    out.set_model(model, model_dir)
    out.set_trainer(trainer)
    out.create_all()
    for split in ("train", "test"):
        m = scipy.io.loadmat(os.path.join(out_dir, "predict_%s.mat" % split))
        assert m["mean"].shape == (300, 1) and m["std"].shape == (300, 1) and m["gt"].shape == (300, 1)
        assert np.all(m["std"] > 0) and np.all(np.isfinite(m["mean"]))
    lines = open(os.path.join(out_dir, "mse.txt")).read().split()
    assert lines[0] == "MSE:" and lines[2] == "RMSE:" and float(lines[3]) == pytest.approx(out.get_last_rmse(), abs=1e-6)
    assert float(lines[3]) == pytest.approx(np.sqrt(float(lines[1])), rel=1e-4)
    dump = open(os.path.join(out_dir, "var_dump.txt")).read()
    for name in model.var_dict:
        assert name + ":\n" in dump
    block = dump.split("IP pos f:\n")[1].split("\n\n")[0].strip().split("\n")
    assert len(block) == 20 and len(block[0].split()) == dim_x + 1      # [M, dx + du] rows
    assert np.loadtxt(os.path.join(out_dir, "training_loss.txt")).shape == (2, 3)


def test_bench_b200_arm_prints_the_contract_line():
    """bench.py on the GPU (small batch): one JSON line with the contract keys, roofline and end-to-end objects."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "2", "--warmup", "3", "--batch", "64",
                          "--no-cpu-baseline"], capture_output=True, text=True, timeout=600, check=True).stdout
    line = json.loads(out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["scaling"] == "weak" and line["value"] > 0 and line["gpu_launches"] > 0
    # headline = the target shape (M=100, D=4, T=100): u [64,100,2] + y [64,100,2] float32 per step
    assert line["e2e"]["h2d_bytes_per_step"] == 64 * 100 * 4 * 4 and line["e2e"]["d2h_bytes_per_step"] == 8
    assert 0 < line["e2e"]["value"] <= line["value"] * 1.05
    assert line["config"]["M"] == 100 and line["config"]["dx"] == 4
    r = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "simt", "hbm"):
        assert key in r, key
    # no fraction may exceed 1: tensor FLOPs against the tensor roof, O(M*D) FLOPs against the measured FP32 roof
    assert r["bound"] == "tensor" and 0 < r["frac"] < 1 and 0 < r["simt"]["frac"] < 1 and r["unit"] == "TFLOP/s"
    assert 60 < line["fp32_fma_peak_measured_tflops"] < 80
    assert "workload" in line["config"] and "model" not in line["config"]
    x = line["extra"]["robomove_m20"]                  # BASELINE.json configs[1] in the same line
    assert x["config"]["M"] == 20 and x["value"] > 0 and x["roofline"]["bound"] == "fp32_simt" and 0 < x["roofline"]["frac"] < 1


@pytest.mark.parametrize("M", [20, 100], ids=["register_m20", "tensor_m100"])
def test_predict_only_skips_dead_message_work_and_changes_nothing(M):
    """Free-running prediction (condition=False) reads y2[t] only for t < recog_len (cbfssm.py:227-228): with
    CBF_FLAG_PREDICT_ONLY the library runs one message chain of <= 2R steps instead of ~2T, and pred_mean / pred_var
    / x_final are bit-identical to the full evaluation."""
    eng = _engine(M=M, S=16, R=5)
    B, T = 3, 64
    u, y, eb, zb, ef = _inputs(eng, B, T, seed=4)
    outs = []
    for po in (False, True):
        n0 = eng.launches
        eng.forward(u, y, eb, zb, ef, False, predict_only=po)
        xf, _ = eng.export_states(y)
        pm, pv = eng.moments(xf, eng.dims.dim_y, eng.var_y)
        torch.cuda.synchronize()
        outs.append((xf.clone(), pm.clone(), pv.clone(), float(eng.terms[2])))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    assert outs[1][3] != outs[0][3]                     # the entropy term covers only the steps that were run
    with pytest.raises(Exception):
        eng.backward()                                  # a predict-only forward cannot be differentiated
    # the public API takes the short cut by itself when only prediction handles are fetched
    from cbf_ssm_b200.engine import ElboEngine
    eng.forward(u, y, eb, zb, ef, True, predict_only=True)   # no effect with condition=True: full message, backward allowed
    eng.backward()


@pytest.mark.parametrize("M,flags", [(20, 4), (100, 8)], ids=["register_m20", "tensor_m100"])
def test_saved_evaluations_equal_recomputation(M, flags, monkeypatch):
    """The register path keeps (k, a, fmean, fvar) of every GP evaluation for the reverse pass, the tensor path
    (fmean, fvar, amax), when the extra workspace fits its budget; otherwise the reverse kernels recompute them:
    both give bit-identical gradients."""
    outs = []
    for budget in ("0", None):
        if budget is None:
            monkeypatch.delenv("CBFSSM_B200_SAVE_EVAL_BYTES", raising=False)
        else:
            monkeypatch.setenv("CBFSSM_B200_SAVE_EVAL_BYTES", budget)
        eng = _engine(M=M, R=20, lf=(20.0, 0.3))
        eng.flags = flags
        B, T = 7, 120
        u, y, eb, zb, ef = _inputs(eng, B, T, seed=2)
        eng.forward(u, y, eb, zb, ef, True)
        eng.backward()
        torch.cuda.synchronize()
        outs.append((eng._gflat.clone(), eng.terms.clone(), int(eng._ws.numel())))
    assert outs[1][2] > outs[0][2]                       # the saved evaluations live in the workspace
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("M,B,T,flags", [(20, 4544, 300, 4), (100, 1515, 100, 8)], ids=["register_m20_bench_shape", "tensor_m100_bench_shape"])
def test_bench_scale_float32_accumulation_against_the_float64_path(M, B, T, flags):
    """Accuracy at the length the bench runs, not only at the <= 1 600 particles the oracle can afford: the float32
    paths' ELBO terms and all 12 gradients at the full bench shape (227 200 particles x 300 steps at M = 20,
    75 750 x 100 at M = 100: up to 7e7 particle-steps accumulated per tensor entry in float32 per warp / per CTA,
    float64 across them) against the float64 batched path on the same inputs, which accumulates everything in
    double and is itself checked against the oracle at 2e-5."""
    eng = _engine(M=M, S=50, R=50, lf=(20.0, 0.1))
    u, y, eb, zb, ef = _inputs(eng, B, T, seed=6)
    res = {}
    for name, fl in (("f32", flags), ("f64", 128)):
        eng.flags = fl
        out = eng.forward(u, y, eb, zb, ef, True)
        eng.backward()
        torch.cuda.synchronize()
        res[name] = ({k: float(v) for k, v in out.items()}, {k: v.copy() for k, v in eng.get_grads().items()})
    for k in ("loss", "loglik", "kl_x", "entropy"):
        a, b = res["f32"][0][k], res["f64"][0][k]
        assert abs(a - b) <= 2e-5 * abs(b), (k, a, b)
    worst = {k: rel_inf(res["f32"][1][k], res["f64"][1][k]) for k in O.PARAM_NAMES}
    print("bench-scale float32 vs float64:", {k: "%.1e" % v for k, v in worst.items()})
    bad = {k: v for k, v in worst.items() if not v < 1e-4}
    assert not bad, bad


def test_model_level_float64_precision_switch():
    """config['gpu_precision'] = 'float64' runs the model on the float64 batched path: same loss as float32 to
    float32 accuracy on a well-conditioned problem, and training still works through the reference-shaped API."""
    from cbf_ssm_b200.datasets import SpringNonlinearSynthetic
    from cbf_ssm_b200.model import CBFSSM

    class SmallSpring(SpringNonlinearSynthetic):
        exp_len = 300
    ds = SmallSpring(20, 10, seed=2)
    base = {'ds': SmallSpring, 'batch_size': 4, 'shuffle': 1, 'dim_x': 4, 'ind_pnt_num': 20, 'samples': 8,
            'learning_rate': 0.01, 'loss_factors': np.asarray([10., 0.2]), 'k_factor': 1., 'recog_len': 6,
            'zeta_pos': 2., 'zeta_mean': 0.1 ** 2, 'zeta_var': 0.01 ** 2, 'var_x': np.asarray([0.1 ** 2] * 4),
            'var_y': np.asarray([1. ** 2] * 4), 'gp_var': 0.1 ** 2, 'gp_len': 1.}
    u, y = ds.train_in_batch[:4], ds.train_out_batch[:4]
    T = u.shape[1]
    g = np.random.default_rng(1)
    draws = (g.standard_normal((2, T, 4, 8)), g.standard_normal((2, T, 4, 8)), g.standard_normal((T - 1, 4, 8)))
    losses = {}
    for prec in ("float32", "float64"):
        m = CBFSSM(dict(base, gpu_precision=prec), seed=3)
        assert m.engine.flags == (128 if prec == "float64" else 0)
        m.inject_draws(*draws)
        l0 = float(m.evaluate_batch(u, y, ["train", "loss"], True)[1])
        m.inject_draws(*draws)
        l1 = float(m.evaluate_batch(u, y, ["train", "loss"], True)[1])
        assert l1 < l0                                   # one Adam step on the same minibatch and draws
        losses[prec] = (l0, l1)
    assert losses["float32"][0] == pytest.approx(losses["float64"][0], rel=1e-5)
    with pytest.raises(ValueError):
        CBFSSM(dict(base, gpu_precision="float16"), seed=3)


def test_reference_checkpoint_export_import_round_trip(tmp_path):
    """A model written as a TensorFlow-V2 checkpoint under the reference graph's variable names and read back into a
    differently initialised model: parameters, Adam slots and step count identical (format: training/tf_checkpoint.py)."""
    from cbf_ssm_b200.datasets import SpringNonlinearSynthetic
    from cbf_ssm_b200.model import CBFSSM
    from cbf_ssm_b200.training import export_reference_checkpoint, import_reference_checkpoint, read_tf_checkpoint

    class SmallSpring(SpringNonlinearSynthetic):
        exp_len = 300
    ds = SmallSpring(20, 10, seed=2)
    cfg = {'ds': SmallSpring, 'batch_size': 4, 'shuffle': 1, 'dim_x': 4, 'ind_pnt_num': 20, 'samples': 8,
           'learning_rate': 0.01, 'loss_factors': np.asarray([10., 0.2]), 'k_factor': 1., 'recog_len': 6,
           'zeta_pos': 2., 'zeta_mean': 0.1 ** 2, 'zeta_var': 0.01 ** 2, 'var_x': np.asarray([0.1 ** 2] * 4),
           'var_y': np.asarray([1. ** 2] * 4), 'gp_var': 0.1 ** 2, 'gp_len': 1.}
    a = CBFSSM(cfg, seed=3)
    for _ in range(3):
        a.evaluate_batch(ds.train_in_batch[:4], ds.train_out_batch[:4], ["train", "loss"], True)
    prefix = str(tmp_path / "best.ckpt")
    export_reference_checkpoint(a, prefix)
    ck = read_tf_checkpoint(prefix)
    assert ck["Variable"].shape == (20, 5) and ck["kern/Variable"].shape == (1,) and ck["Variable_7"].shape == (4,)
    assert "Variable_3/Adam_1" in ck and float(ck["beta1_power"]) == pytest.approx(0.9 ** 4)
    b = CBFSSM(cfg, seed=99)
    loaded = import_reference_checkpoint(b, prefix)
    assert len(loaded) == 12 * 3 + 1
    assert torch.equal(a.engine.theta, b.engine.theta) and torch.equal(a.engine.adam_m, b.engine.adam_m)
    assert torch.equal(a.engine.adam_v, b.engine.adam_v) and b.engine.adam_t == a.engine.adam_t == 3


def test_condition_number_diagnostic_and_warning():
    """The prologue leaves cond_1(K_zz + 1e-8 I) in its state; CBFSSM warns at initialisation when the float32
    kernels cannot resolve the inducing set (M = 100 points in 3 input dims) and stays silent otherwise."""
    import warnings
    from cbf_ssm_b200.model import CBFSSM
    eng = _engine(M=20)
    eng.prologue()
    torch.cuda.synchronize()
    p = eng.get_params()
    kern = O.RBF(torch.tensor(p["f.variance_unc"]), torch.tensor(p["f.lengthscales_unc"]))
    K = kern.K(torch.tensor(p["f.zeta_pos"])).numpy() + 1e-8 * np.eye(20)
    ref = np.linalg.norm(K, 1) * np.linalg.norm(np.linalg.inv(K), 1)
    assert eng.cond_kzz()["f"] == pytest.approx(ref, rel=1e-6)

    def cfg(dx, du, dy, M):
        DS = type("DS", (), {"dim_u": du, "dim_y": dy})
        return {'ds': DS, 'batch_size': 4, 'shuffle': 1, 'dim_x': dx, 'ind_pnt_num': M, 'samples': 8, 'learning_rate': 0.01,
                'loss_factors': np.asarray([10., 0.]), 'k_factor': 1., 'recog_len': 6, 'zeta_pos': 2., 'zeta_mean': 0.01,
                'zeta_var': 1e-4, 'var_x': np.full(dx, 0.01), 'var_y': np.full(dx, 1.0), 'gp_var': 0.01, 'gp_len': 1.}
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        CBFSSM(cfg(4, 2, 2, 100), seed=1)                      # cond ~ 10: silent
        CBFSSM(dict(cfg(2, 1, 1, 100), gpu_precision="float64"), seed=1)
    with pytest.warns(RuntimeWarning, match="accuracy-limited"):
        CBFSSM(cfg(2, 1, 1, 100), seed=1)
