"""Host-side mirror of the reference's ``BaseModel`` (cbfssm/model/base_model.py:6-69).

The reference exposes TensorFlow placeholders / tensors as attributes and executes them
with ``sess.run``.  Here the same attribute names are lightweight ``Fetch`` handles and
``Session`` is a shim whose ``run`` evaluates the requested handles for the next
minibatch on the GPU (through ``ElboEngine`` and the C ABI).  ``load_ds`` / ``run``
keep the reference semantics: repeat -> shuffle(buffer) -> batch (last batch short),
``run`` drains the iterator and concatenates per-batch results along axis 0.
"""
from __future__ import annotations

import sys

import numpy as np


class OutOfRangeError(Exception):
    """Raised by ``Session.run`` when the minibatch iterator is exhausted
    (tf.errors.OutOfRangeError, base_model.py:64)."""


class Fetch:
    """Named handle standing in for a TF graph tensor / op / placeholder."""

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return f"<Fetch {self.name}>"


class _GraphShim:
    def as_default(self):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class Session:
    """``tf.Session`` look-alike: ``sess.run(fetches, feed_dict)`` on one model."""

    def __init__(self, model=None, config=None):
        self.model = model

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def run(self, fetches, feed_dict=None):
        model = self.model
        if model is None:
            f0 = fetches[0] if isinstance(fetches, (tuple, list)) else fetches
            model = getattr(f0, "model", None)
        if model is None:
            raise ValueError("Session.run: cannot infer the model; construct Session(model)")
        return model._session_run(fetches, feed_dict or {})


class BaseModel:

    def __init__(self, config, dtype=np.float64):
        self.config = config
        self.dtype = dtype
        self.graph = _GraphShim()
        self._build_ds_pipeline()
        self._build_graph()

    def _handle(self, name):
        f = Fetch(name)
        f.model = self
        return f

    def _build_ds_pipeline(self):
        # base_model.py:15-31
        self.data_in = self._handle("data_in")
        self.data_out = self._handle("data_out")
        self.repeats = self._handle("repeats")
        self.condition = self._handle("condition")
        self._ds = None
        self._order = None
        self._cursor = 0
        self._shuffle_rng = np.random.RandomState(self.config.get("shuffle_seed", None))

    def _build_graph(self):
        pass

    def load_ds(self, sess, data_in, data_out, repeats=1):
        """Initialise the minibatch iterator (base_model.py:36-40): repeat, then a
        shuffle buffer of size config['shuffle'] (tf.data semantics: uniform only when
        the buffer covers the data), then batches of config['batch_size']."""
        data_in = np.asarray(data_in)
        data_out = np.asarray(data_out)
        n = data_in.shape[0]
        stream = np.tile(np.arange(n), int(repeats))
        buf = int(self.config["shuffle"])
        rng = self._shuffle_rng
        if buf <= 1:
            order = stream
        elif buf >= stream.size:
            order = rng.permutation(stream)
        else:
            order = np.empty_like(stream)
            window = list(stream[:buf])
            nxt = buf
            for i in range(stream.size):
                j = rng.randint(len(window))
                order[i] = window[j]
                if nxt < stream.size:
                    window[j] = stream[nxt]
                    nxt += 1
                else:
                    window.pop(j)
        order = self._same_order_on_every_rank(order)
        self._ds = (data_in, data_out)
        self._order = order
        self._cursor = 0

    def _same_order_on_every_rank(self, order):
        """With a process group every rank shards the *same* minibatch, so the shuffled order is
        rank 0's, broadcast -- the default shuffle RNG is unseeded like the reference's."""
        group = getattr(self, "_group", None)
        if group is None or getattr(self, "world", 1) <= 1:
            return order
        import torch
        import torch.distributed as dist
        dev = self.engine.device if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.as_tensor(np.ascontiguousarray(order, dtype=np.int64)).to(dev)
        dist.broadcast(t, src=dist.get_global_rank(group, 0), group=group)
        return t.cpu().numpy()

    def _next_batch(self):
        if self._ds is None or self._cursor >= self._order.size:
            raise OutOfRangeError()
        bs = int(self.config["batch_size"])
        idx = self._order[self._cursor:self._cursor + bs]
        self._cursor += bs
        return self._ds[0][idx], self._ds[1][idx]

    @staticmethod
    def run(sess, tensors, feed_dict, show_progress=False):
        """Drain the iterator; concatenate per-batch results (base_model.py:42-69)."""
        res_all = None
        while True:
            try:
                res = sess.run(tensors, feed_dict=feed_dict)
            except OutOfRangeError:
                break
            if show_progress:
                sys.stdout.write('.')
                sys.stdout.flush()
            if not isinstance(res, tuple):
                res = (res,)
            if res_all is None:
                res_all = [None if r is None else np.atleast_1d(r) for r in res]
            else:
                for i, item in enumerate(res):
                    if item is not None:
                        res_all[i] = np.concatenate((res_all[i], np.atleast_1d(item)), axis=0)
        if show_progress:
            print()
        return res_all
