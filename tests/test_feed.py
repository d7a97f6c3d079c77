"""CPU: the device-resident batch iterator (graph-wavenet_b200/feed.py) against the numpy restatement of the reference's
``DataLoader`` (Utils/util.py:14-54): same padding, same batch count, same sample order for the same host RNG state over
several shuffled epochs; ragged and exact-multiple dataset sizes; pad_with_last_sample off."""
import numpy as np
import pytest
import torch

import __graft_entry__ as ge
from oracle.feed_oracle import DataLoaderOracle

ge.load_package()
from graph_wavenet_b200.feed import DataLoader  # noqa: E402


@pytest.mark.parametrize("n,bs,pad", [(23, 8, True), (24, 8, True), (5, 8, True), (23, 8, False), (1, 4, True)])
def test_feed_matches_reference_iterator(n, bs, pad):
    rng = np.random.RandomState(7)
    xs = rng.randn(n, 12, 9, 2).astype(np.float32)
    ys = rng.randn(n, 12, 9, 2).astype(np.float32)
    ours, ref = DataLoader(xs, ys, bs, pad), DataLoaderOracle(xs, ys, bs, pad)
    assert (ours.size, ours.num_batch) == (ref.size, ref.num_batch)
    for epoch in range(3):
        if epoch:
            np.random.seed(100 + epoch)
            ours.shuffle()
            np.random.seed(100 + epoch)
            ref.shuffle()
        got = list(ours.get_iterator())
        want = list(ref.batches())
        assert len(got) == len(want)
        for (gx, gy), (wx, wy) in zip(got, want):
            assert torch.equal(gx, torch.from_numpy(wx)) and torch.equal(gy, torch.from_numpy(wy))
            v = gx.transpose(1, 3)                      # train.py:245 -- the strided view the model consumes
            assert tuple(v.shape) == (wx.shape[0], 2, 9, 12)
        assert np.array_equal(ours.xs.numpy(), ref.xs)
