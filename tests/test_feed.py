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


def test_feed_oracle_is_pinned_to_the_real_reference_loader():
    """oracle/feed_oracle.py (and the product feed) against sample orders recorded from the REAL ``Utils/util.py``
    DataLoader by tests/tools/make_golden_feed.py: padding, batch count and three epochs of shuffled order."""
    import json, os
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "feed_order.json")))
    assert len(cases) == 5
    for c in cases:
        n, bs, pad = c["n"], c["batch_size"], c["pad"]
        xs = np.arange(n, dtype=np.float32).reshape(n, 1, 1, 1)
        for make in (DataLoaderOracle, DataLoader):
            dl = make(xs, -xs, bs, pad)
            assert (int(dl.size), int(dl.num_batch)) == (c["size"], c["num_batch"])
            for epoch, want in enumerate(c["epochs"]):
                if epoch:
                    np.random.seed(100 + epoch)
                    dl.shuffle()
                it = dl.batches() if make is DataLoaderOracle else dl.get_iterator()
                got = [[int(v) for v in np.asarray(bx).reshape(-1)] for bx, _ in it]
                assert got == want, (make.__name__, n, bs, pad, epoch)
