"""Epoch driver with the reference ``Trainer`` interface.

Behaviour kept from cbfssm/training/trainer.py:10-63 (it is the caller of the hot path):

* ``Trainer(model, model_dir).train(ds, epochs, retrain=False)``;
* ``retrain`` restores ``<model_dir>/model.ckpt`` (run_robomove.py:47 curriculum), otherwise the
  model is re-initialised;
* one epoch = a pass over ``ds.train_*_batch`` fetching ``(train, loss)`` and a pass over
  ``ds.test_*_batch`` fetching ``loss`` only, both with ``condition=True``;
* the epoch losses are the *means of the per-minibatch losses* and are appended to
  ``train_all`` / ``test_all``;
* ``best.ckpt`` whenever the epoch's training loss is the lowest so far, ``model.ckpt`` at the end.

Added here: particle-steps/s per epoch in ``throughput_all`` (the hot path's own metric).
"""
import os
import time

import numpy as np

from ..model.base_model import Session

try:
    from tqdm import tqdm as _progress
except ImportError:  # pragma: no cover
    def _progress(it):
        return it


class Trainer:

    BEST = "best.ckpt"
    LAST = "model.ckpt"

    def __init__(self, model, model_dir):
        self.model = model
        self.model_dir = model_dir
        self.train_all = []
        self.test_all = []
        self.throughput_all = []

    # -- helpers -------------------------------------------------------------------------
    def _path(self, name):
        return os.path.join(self.model_dir, name)

    def _pass(self, sess, data_in, data_out, fetches):
        """One pass over a set of windows; returns the mean of the per-minibatch losses."""
        model = self.model
        model.load_ds(sess, data_in, data_out)
        results = model.run(sess, fetches, {model.condition: True})
        losses = np.asarray(results[-1], dtype=np.float64)          # loss is the last fetch
        if not np.all(np.isfinite(losses)):
            # The float32 kernels clamp the GP variance at its exact lower bound, so this is not the usual
            # cancellation NaN; a non-finite loss means the parameters themselves have left the representable range
            # (or cond(K_zz) is far beyond float32: try config['gpu_precision'] = 'float64').
            raise FloatingPointError("non-finite ELBO in minibatch %d of this pass; the update was applied -- restore "
                                     "the last checkpoint" % int(np.argmin(np.isfinite(losses))))
        return float(np.mean(losses))

    def _particle_steps(self, windows):
        n_seq, seq_len = windows.shape[0], windows.shape[1]
        return n_seq * seq_len * int(self.model.config['samples'])

    # -- public --------------------------------------------------------------------------
    def train(self, ds, epochs, retrain=False, verbose=True):
        model = self.model
        os.makedirs(self.model_dir, exist_ok=True)
        if verbose:
            print('\nTraining...\n')
        with model.graph.as_default(), Session(model) as sess:
            if retrain:
                model.saver.restore(sess, self._path(self.LAST))
            else:
                sess.run(model.init)

            best = float('inf')
            epoch_iter = _progress(range(epochs)) if verbose else range(epochs)
            for epoch in epoch_iter:
                tic = time.perf_counter()
                train_loss = self._pass(sess, ds.train_in_batch, ds.train_out_batch, (model.train, model.loss))
                elapsed = time.perf_counter() - tic
                test_loss = self._pass(sess, ds.test_in_batch, ds.test_out_batch, (model.loss,))

                self.train_all.append(train_loss)
                self.test_all.append(test_loss)
                self.throughput_all.append(self._particle_steps(ds.train_in_batch) / max(elapsed, 1e-9))
                if verbose:
                    print('[%04d]: Train %s, Test %s  (%.3g particle-steps/s)'
                          % (epoch, train_loss, test_loss, self.throughput_all[-1]))

                if train_loss < best:
                    best = train_loss
                    model.saver.save(sess, self._path(self.BEST))

            model.saver.save(sess, self._path(self.LAST))
        return self
