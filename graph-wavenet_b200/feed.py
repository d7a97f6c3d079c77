"""Input feed for the hot path: the reference's batch iterator with the data resident on the
device.

``DataLoader`` keeps the reference's interface and semantics (Utils/util.py:14-54: pad with the last sample to a
multiple of the batch size, ``shuffle()`` draws ``np.random.permutation`` from the same host RNG stream, ``get_iterator()``
yields consecutive batches) but uploads ``xs`` / ``ys`` ONCE and yields device tensors in the loader layout
``[B, T, N, F]`` -- which is also the plan's physical BLNC layout, so ``x.transpose(1, 3)`` (train.py:245) is the strided
view the start conv reads directly.  The reference re-uploads every batch from pageable host memory, synchronously
(train.py:244-247); METR-LA's training split is 0.95 GB, 0.5 % of one B200's HBM.
"""
import numpy as np
import torch


class DataLoader(object):
    def __init__(self, xs, ys, batch_size, pad_with_last_sample=True, device=None):
        self.batch_size = batch_size
        self.current_ind = 0
        xs, ys = np.asarray(xs), np.asarray(ys)
        if pad_with_last_sample:
            num_padding = (batch_size - (len(xs) % batch_size)) % batch_size
            xs = np.concatenate([xs, np.repeat(xs[-1:], num_padding, axis=0)], axis=0)
            ys = np.concatenate([ys, np.repeat(ys[-1:], num_padding, axis=0)], axis=0)
        self.size = len(xs)
        self.num_batch = int(self.size // self.batch_size)
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self._x = torch.as_tensor(xs, dtype=torch.float32).to(self.device)
        self._y = torch.as_tensor(ys, dtype=torch.float32).to(self.device)
        self._order = np.arange(self.size)
        self._order_dev = torch.arange(self.size, device=self.device)

    # the reference exposes the (permuted) arrays; materialise them on demand
    @property
    def xs(self):
        return self._x.index_select(0, self._order_dev)

    @property
    def ys(self):
        return self._y.index_select(0, self._order_dev)

    def shuffle(self):
        permutation = np.random.permutation(self.size)      # same RNG draw as Utils/util.py:37
        self._order = self._order[permutation]               # the reference permutes its already-permuted arrays
        self._order_dev = torch.as_tensor(self._order, device=self.device)

    def get_iterator(self):
        self.current_ind = 0

        def _wrapper():
            while self.current_ind < self.num_batch:
                start_ind = self.batch_size * self.current_ind
                end_ind = min(self.size, self.batch_size * (self.current_ind + 1))
                idx = self._order_dev[start_ind:end_ind]
                yield (self._x.index_select(0, idx), self._y.index_select(0, idx))
                self.current_ind += 1

        return _wrapper()
