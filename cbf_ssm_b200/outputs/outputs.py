"""Report writer with the reference's ``Outputs`` interface (cbfssm/outputs/outputs.py:11-164):
``Outputs(out_dir)``, ``set_ds / set_model / set_trainer``, ``create_all()``, ``get_last_rmse()``.

It is the caller that turns the hot path into the paper's numbers, so the numerical artefacts keep
the reference's file names and formats:

* ``predict_train.mat`` / ``predict_test.mat`` -- keys ``mean``, ``std``, ``gt`` of the free-running
  (``condition=False``) prediction over the first ``predict_size`` steps of experiment 0, denormalised;
* ``mse.txt``      -- ``MSE:  %f`` / ``RMSE: %f`` over whole test experiments (mean of per-experiment MSE);
* ``var_dump.txt`` -- every ``model.var_dict`` entry, ``% .4e`` per value, one row per matrix row;
* ``training_loss.txt`` (epoch, train, test) -- the reference plots these to ``training_loss.pdf``; the
  PDFs are written as well when matplotlib is importable (it is not part of this image).
"""
import math
import os

import numpy as np
import scipy.io

from ..model.base_model import Session

try:                                                   # plotting is optional
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
except ImportError:                                    # pragma: no cover
    plt = None


class Outputs:

    def __init__(self, out_dir):
        self.out_dir = out_dir
        self.ds = self.model = self.model_path = self.trainer = self.last_rmse = None
        os.makedirs(out_dir, exist_ok=True)

    def set_ds(self, ds):
        self.ds = ds

    def set_model(self, model, model_dir):
        self.model = model
        self.model_path = os.path.join(model_dir, 'best.ckpt')

    def set_trainer(self, trainer):
        self.trainer = trainer

    def get_last_rmse(self):
        return self.last_rmse

    def create_all(self):
        if self.model is None or self.ds is None:
            raise AssertionError("set_model and set_ds first")
        with self.model.graph.as_default(), Session(self.model) as sess:
            self.model.saver.restore(sess, self.model_path)
            print("Generating outputs...")
            self._create_all(sess)

    def _create_all(self, sess):
        self.training_stats()
        self.prediction(sess)
        self.test_mse(sess)
        self.var_dump(sess)

    def _file(self, name):
        return os.path.join(self.out_dir, name)

    # ---------------------------------------------------------------------------------
    def training_stats(self):
        if self.trainer is None:
            return
        print("  training stats")
        rows = np.column_stack((np.arange(len(self.trainer.train_all)), self.trainer.train_all,
                                self.trainer.test_all))
        np.savetxt(self._file('training_loss.txt'), rows, header="epoch train test")
        if plt is not None:                            # pragma: no cover
            plt.figure(1)
            plt.plot(self.trainer.train_all, label='train')
            plt.plot(self.trainer.test_all, label='test')
            plt.legend()
            plt.savefig(self._file('training_loss.pdf'))
            plt.close(1)

    def _predict_one(self, sess, data_in, data_out):
        model, ds = self.model, self.ds
        model.load_ds(sess, data_in, data_out)
        mean, var = sess.run((model.pred_mean, model.pred_var), feed_dict={model.condition: False})
        return (ds.denormalize(mean, 'out')[0], ds.denormalize(np.sqrt(var), 'out', shift=False)[0],
                ds.denormalize(data_out, 'out')[0])

    def prediction(self, sess, predict_size=300):
        print("  prediction")
        ds = self.ds
        predict_size = min(ds.train_in.shape[1], predict_size)
        for split, din, dout in (('train', ds.train_in, ds.train_out), ('test', ds.test_in, ds.test_out)):
            mean, std, gt = self._predict_one(sess, din[0:1, :predict_size], dout[0:1, :predict_size])
            scipy.io.savemat(self._file('predict_%s.mat' % split), {'mean': mean, 'std': std, 'gt': gt})
            if plt is not None:                        # pragma: no cover
                steps = np.arange(mean.shape[0])
                plt.figure(1, figsize=(6, 4))
                plt.plot(gt[:, 0], label='ground truth')
                plt.plot(mean[:, 0], label='prediction')
                plt.fill_between(steps, mean[:, 0] - 1.96 * std[:, 0], mean[:, 0] + 1.96 * std[:, 0], alpha=0.4)
                plt.legend(loc=2)
                plt.grid(True)
                plt.xlabel("time (steps)")
                plt.savefig(self._file('predict_%s.pdf' % split), bbox_inches='tight')
                plt.close(1)

    def test_mse(self, sess):
        print("  test mse")
        model, ds = self.model, self.ds
        per_experiment = []
        for i in range(ds.test_in.shape[0]):
            model.load_ds(sess, ds.test_in[i:i + 1], ds.test_out[i:i + 1])
            pred = model.run(sess, model.pred_mean, {model.condition: False})[0]
            pred = ds.denormalize(pred, 'out')[0]
            truth = ds.denormalize(ds.test_out[i:i + 1], 'out')[0]
            per_experiment.append(float(np.mean((truth - pred) ** 2)))   # sklearn mean_squared_error: uniform average
        mse = float(np.mean(per_experiment))
        rmse = math.sqrt(mse)
        with open(self._file('mse.txt'), 'w') as fh:
            fh.write("MSE:  %f\n" % mse)
            fh.write("RMSE: %f\n" % rmse)
        self.last_rmse = rmse

    def var_dump(self, sess):
        print("  var dump")
        model = self.model
        with open(self._file('var_dump.txt'), 'w') as fh:
            for name, handle in model.var_dict.items():
                value = np.asarray(sess.run(handle, feed_dict={model.condition: False}))
                fh.write(name + ":\n")
                if value.ndim == 1:
                    fh.write("".join("  % .4e" % v for v in value))
                elif value.ndim == 2:
                    for row in value:
                        fh.write("".join("  % .4e" % v for v in row) + "\n")
                fh.write("\n\n")
