"""Large-N agreement of the tensor-core path with the cooperative float32 path (same inputs):
relative inf-norm difference of every kernel-level gradient block and of the ELBO terms."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cbf_ssm_b200.engine import ElboEngine, ModelDims, init_param_arrays

def run(flags, B=1024, T=100, M=100):
    dims = ModelDims(4, 2, 2, M, 50, 50, 1.0, (10.0, 0.0))
    eng = ElboEngine(dims)
    eng.flags = flags
    cfg = dict(zeta_pos=2.0, zeta_mean=0.01, zeta_var=1e-4, gp_var=0.01, gp_len=1.0, var_x=np.full(4, 0.01), var_y=np.full(4, 1.0))
    eng.set_params(init_param_arrays(dims, cfg, 1))
    g = torch.Generator().manual_seed(0)
    dev = eng.device
    u = torch.randn(B, T, 2, generator=g).to(dev); y = torch.randn(B, T, 2, generator=g).to(dev)
    N = B * 50
    eb = torch.randn(2, T, N, generator=g).to(dev); zb = torch.randn(2, T, N, generator=g).to(dev); ef = torch.randn(T - 1, N, generator=g).to(dev)
    eng.forward(u, y, eb, zb, ef, True); eng.backward(); torch.cuda.synchronize()
    return eng.terms[:3].cpu().numpy().copy(), eng.kernel_level_grads(), eng.get_grads()

t0, k0, g0 = run(1)
t1, k1, g1 = run(0)
rel = lambda a, b: float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))
print("terms", rel(t1, t0))
for k in k0: print("kernel-level %-10s %.2e" % (k, rel(k1[k], k0[k])))
print("raw grads max rel", max(rel(g1[k], g0[k]) for k in g0))
