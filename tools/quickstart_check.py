"""Runs the README quick-start (1 epoch) on cuda:0."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cbf_ssm_b200.datasets import RoboMoveSynthetic
from cbf_ssm_b200.model import CBFSSM
from cbf_ssm_b200.training import Trainer
from cbf_ssm_b200.outputs import Outputs

d = tempfile.mkdtemp()
ds = RoboMoveSynthetic(300, 50)
config = {'ds': RoboMoveSynthetic, 'batch_size': 32, 'shuffle': 10000, 'dim_x': 4, 'ind_pnt_num': 20, 'samples': 50,
          'learning_rate': 0.01, 'loss_factors': np.asarray([20., 0.]), 'k_factor': 1., 'recog_len': 50,
          'zeta_pos': 2., 'zeta_mean': 0.1 ** 2, 'zeta_var': 0.01 ** 2, 'var_x': np.asarray([0.1 ** 2] * 4),
          'var_y': np.asarray([1. ** 2] * 4), 'gp_var': 0.1 ** 2, 'gp_len': 1.}
model = CBFSSM(config)
trainer = Trainer(model, d + '/model')
trainer.train(ds, epochs=2)
out = Outputs(d + '/report'); out.set_ds(ds); out.set_model(model, d + '/model'); out.set_trainer(trainer)
out.create_all()
print("files:", sorted(os.listdir(d + '/report')), "rmse", out.get_last_rmse(), "throughput", trainer.throughput_all)
