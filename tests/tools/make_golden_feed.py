"""Generate tests/golden/feed_order.json from the REAL reference ``DataLoader`` (Utils/util.py:14-54).  Build container only.
For every (n, batch_size, pad) case the sample ids seen by three epochs (no shuffle, then two shuffles seeded through the
numpy global RNG like train.py:46,242) are recorded: ``xs[i] = i`` so a batch's content is its sample order."""
import json, os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(__file__))
from make_golden import load_reference            # noqa: E402  (stubs ipdb / matplotlib / nibabel, puts the reference on sys.path)

load_reference()
import Utils.util as ref_util                      # noqa: E402

CASES = [(23, 8, True), (24, 8, True), (5, 8, True), (23, 8, False), (1, 4, True)]
out = []
for n, bs, pad in CASES:
    xs = np.arange(n, dtype=np.float32).reshape(n, 1, 1, 1)
    ys = -xs
    dl = ref_util.DataLoader(xs, ys, bs, pad_with_last_sample=pad)
    epochs = []
    for epoch in range(3):
        if epoch:
            np.random.seed(100 + epoch)
            dl.shuffle()
        batches = []
        for bx, by in dl.get_iterator():
            assert np.array_equal(bx, -by)
            batches.append([int(v) for v in bx.reshape(-1)])
        epochs.append(batches)
    out.append({"n": n, "batch_size": bs, "pad": pad, "size": int(dl.size), "num_batch": int(dl.num_batch), "epochs": epochs})
with open(os.path.join(ROOT, "tests", "golden", "feed_order.json"), "w") as f:
    json.dump(out, f)
print("wrote feed_order.json:", [(c["n"], c["batch_size"], c["pad"], c["num_batch"]) for c in out])
