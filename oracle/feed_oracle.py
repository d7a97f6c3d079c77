"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's batch iterator (Utils/util.py:14-54), the oracle of the
device-resident feed in graph-wavenet_b200/feed.py.  Pinned: tests/golden/feed_order.json holds sample orders recorded from the REAL
reference loader (tests/tools/make_golden_feed.py); tests/test_feed.py checks this restatement against them.  Imported by tests/ only."""
import numpy as np


class DataLoaderOracle:
    def __init__(self, xs, ys, batch_size, pad_with_last_sample=True):      # Utils/util.py:15-35
        self.batch_size = batch_size
        if pad_with_last_sample:
            n = (batch_size - (len(xs) % batch_size)) % batch_size
            xs = np.concatenate([xs, np.repeat(xs[-1:], n, axis=0)], axis=0)
            ys = np.concatenate([ys, np.repeat(ys[-1:], n, axis=0)], axis=0)
        self.size = len(xs)
        self.num_batch = int(self.size // self.batch_size)
        self.xs, self.ys = xs, ys

    def shuffle(self):                                                       # Utils/util.py:36-40
        p = np.random.permutation(self.size)
        self.xs, self.ys = self.xs[p], self.ys[p]

    def batches(self):                                                       # Utils/util.py:42-54
        for i in range(self.num_batch):
            s, e = self.batch_size * i, min(self.size, self.batch_size * (i + 1))
            yield self.xs[s:e, ...], self.ys[s:e, ...]
