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

    BEST = 'best.ckpt'

    def __init__(self, out_dir):
        os.makedirs(out_dir, exist_ok=True)
        self.out_dir = out_dir
        self.ds = self.model = self.model_path = self.trainer = self.last_rmse = None

    # -- wiring (outputs.py:23-34) ---------------------------------------------------------
    def set_ds(self, ds):
        self.ds = ds

    def set_model(self, model, model_dir):
        self.model, self.model_path = model, os.path.join(model_dir, self.BEST)

    def set_trainer(self, trainer):
        self.trainer = trainer

    def get_last_rmse(self):
        return self.last_rmse

    def create_all(self):
        """Restore the best checkpoint and write every report (outputs.py:36-49)."""
        if self.model is None or self.ds is None:
            raise AssertionError("set_model and set_ds first")
        with self.model.graph.as_default(), Session(self.model) as sess:
            self.model.saver.restore(sess, self.model_path)
            print("Generating outputs...")
            for step in (self.training_stats, self.prediction, self.test_mse, self.var_dump):
                step(sess)

    def _create_all(self, sess):                       # kept for callers that hold their own session
        for step in (self.training_stats, self.prediction, self.test_mse, self.var_dump):
            step(sess)

    def _path(self, name):
        return os.path.join(self.out_dir, name)

    def _free_run(self, sess, data_in, data_out):
        """condition=False prediction of one window, denormalised: mean, std, ground truth [T, dy]."""
        model, ds = self.model, self.ds
        model.load_ds(sess, data_in, data_out)
        mean, var = sess.run((model.pred_mean, model.pred_var), feed_dict={model.condition: False})
        to_units = lambda a, shift=True: ds.denormalize(a, 'out', shift=shift)[0]
        return to_units(mean), to_units(np.sqrt(var), False), to_units(data_out)

    # -- reports ---------------------------------------------------------------------------
    def training_stats(self, sess=None):
        tr = self.trainer
        if tr is None:
            return
        print("  training stats")
        table = np.stack((np.arange(len(tr.train_all), dtype=float), np.asarray(tr.train_all, dtype=float),
                          np.asarray(tr.test_all, dtype=float)), axis=1)
        np.savetxt(self._path('training_loss.txt'), table, header="epoch train test")
        if plt is not None:                            # pragma: no cover
            fig = plt.figure()
            plt.plot(table[:, 1], label='train')
            plt.plot(table[:, 2], label='test')
            plt.legend()
            fig.savefig(self._path('training_loss.pdf'))
            plt.close(fig)

    def prediction(self, sess, predict_size=300):
        print("  prediction")
        ds = self.ds
        n = min(ds.train_in.shape[1], predict_size)
        splits = {'train': (ds.train_in, ds.train_out), 'test': (ds.test_in, ds.test_out)}
        for name, (din, dout) in splits.items():
            mean, std, truth = self._free_run(sess, din[:1, :n], dout[:1, :n])
            scipy.io.savemat(self._path('predict_%s.mat' % name), {'mean': mean, 'std': std, 'gt': truth})
            if plt is not None:                        # pragma: no cover
                fig = plt.figure(figsize=(6, 4))
                band = 1.96 * std[:, 0]
                plt.plot(truth[:, 0], label='ground truth')
                plt.plot(mean[:, 0], label='prediction')
                plt.fill_between(np.arange(len(band)), mean[:, 0] - band, mean[:, 0] + band, alpha=0.4)
                plt.legend(loc=2)
                plt.grid(True)
                plt.xlabel("time (steps)")
                fig.savefig(self._path('predict_%s.pdf' % name), bbox_inches='tight')
                plt.close(fig)

    def test_mse(self, sess):
        """Whole test experiments, one at a time; MSE = mean over experiments of the mean squared error
        over (time, dim) in data units (sklearn's uniform average), RMSE = sqrt of that mean."""
        print("  test mse")
        model, ds = self.model, self.ds
        errs = []
        for exp_in, exp_out in zip(ds.test_in, ds.test_out):
            model.load_ds(sess, exp_in[None], exp_out[None])
            pred = model.run(sess, model.pred_mean, {model.condition: False})[0]
            diff = ds.denormalize(pred, 'out')[0] - ds.denormalize(exp_out[None], 'out')[0]
            errs.append(float(np.mean(diff * diff)))
        mse = float(np.mean(errs))
        self.last_rmse = math.sqrt(mse)
        with open(self._path('mse.txt'), 'w') as fh:
            fh.write("MSE:  %f\nRMSE: %f\n" % (mse, self.last_rmse))

    def var_dump(self, sess):
        """Every model.var_dict entry: name, then '% .4e' values (one line per matrix row), blank line."""
        print("  var dump")
        fmt = lambda row: "".join("  % .4e" % v for v in row)
        chunks = []
        for name, handle in self.model.var_dict.items():
            value = np.asarray(sess.run(handle, feed_dict={self.model.condition: False}))
            if value.ndim == 1:
                body = fmt(value)
            elif value.ndim == 2:
                body = "".join(fmt(row) + "\n" for row in value)
            else:
                body = ""
            chunks.append(name + ":\n" + body + "\n\n")
        with open(self._path('var_dump.txt'), 'w') as fh:
            fh.write("".join(chunks))
