"""Dataset base with the reference's attributes (cbfssm/datasets/base_ds.py:5-85):
class attrs ``dim_u, dim_y``; arrays ``train_in/out, test_in/out`` as
[experiments, time, dim]; windows ``*_batch`` as [n_seq, seq_len, dim]."""
import numpy as np


class BaseDS:

    dim_u = None
    dim_y = None

    def __init__(self, seq_len, seq_stride):
        self.seq_len = seq_len
        self.seq_stride = seq_stride
        empty = np.empty(0)
        self.train_in = self.train_out = self.test_in = self.test_out = empty
        self.train_in_batch = self.train_out_batch = self.test_in_batch = self.test_out_batch = empty
        self.mean = {'in': np.empty(()), 'out': np.empty(())}
        self.std = {'in': np.empty(()), 'out': np.empty(())}

    def normalize_init(self, data_in, data_out):
        """Per-dimension mean / std of [time, dim] arrays (base_ds.py:25-31)."""
        if data_in.ndim != 2 or data_out.ndim != 2:
            raise AssertionError("normalize_init expects [time, dim] arrays")
        for key, arr in (('in', data_in), ('out', data_out)):
            self.mean[key] = arr.mean(axis=0)
            self.std[key] = (arr - self.mean[key]).std(axis=0)

    def normalize(self, data, key):
        return (data - self.mean[key]) / self.std[key]

    def denormalize(self, data, key, shift=True):
        scaled = data * self.std[key]
        return scaled + self.mean[key] if shift else scaled

    @staticmethod
    def rnn_batches(x, length, stride, _=0):
        """Sliding windows of ``length`` every ``stride`` over each experiment of
        x [experiments, time, dim]; a final window is added so the last samples are
        covered when (time - length) is not a multiple of stride (base_ds.py:54-77)."""
        if x.ndim != 3:
            raise AssertionError("data must be shaped as [experiments x time x dimension]")
        windows = []
        for ex in x:
            n = ex.shape[0]
            if n < length:
                raise AssertionError("Sequence length must be shorter than data.")
            starts = list(range(0, n - length + 1, stride))
            if (n - length) % stride > 0:
                starts.append(n - length)
            windows.extend(ex[s:s + length] for s in starts)
        return np.stack(windows, axis=0)

    def get_batches(self, seq_len, seq_stride):
        return tuple(self.rnn_batches(a, seq_len, seq_stride, 0)
                     for a in (self.train_in, self.train_out, self.test_in, self.test_out))

    def create_batches(self, verbose=False):
        (self.train_in_batch, self.train_out_batch,
         self.test_in_batch, self.test_out_batch) = self.get_batches(self.seq_len, self.seq_stride)
        if verbose:
            self.print_stats()

    def stats(self):
        """Sizes of the raw series and of the windowed sets."""
        def samples(a):
            return int(a.shape[0] * a.shape[1]) if a.ndim >= 2 else 0
        return {"sequence_length": self.seq_len,
                "train": {"samples": samples(self.train_in), "sequences": int(self.train_in_batch.shape[0])},
                "test": {"samples": samples(self.test_in), "sequences": int(self.test_in_batch.shape[0])}}

    def print_stats(self):
        st = self.stats()
        print('Dataset Stats:')
        print('  sequence length: %d' % st["sequence_length"])
        for part in ("train", "test"):
            print('  %s samples: %d' % (part, st[part]["samples"]))
            print('  %s sequences: %d' % (part, st[part]["sequences"]))
