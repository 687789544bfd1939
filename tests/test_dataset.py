""""Next" row N3 — segment indexing.  CPU tier: the oracle restatement against golden items produced by the reference's
own ``__getitem__`` / ``segment_to_task`` (oracle/make_golden_dataset.py).  GPU tier: megb200_gather_segments through
megb200.data.DeviceSegmentStore, BIT-EXACT (integer indices and the gathered f32 windows)."""
import os

import numpy as np
import pytest
import torch

from oracle import dataset_restatement as DR

GOLD = os.path.join(os.path.dirname(__file__), "golden", "dataset.npz")


def _world():
    g = np.load(GOLD)
    keys = [str(k) for k in g["keys"]]
    nseg = [int(n) for n in g["nseg"]]
    rec = {k: {f"task{t}": g[f"rec/{k}/task{t}"] for t in range(len(nseg))} for k in keys}
    ons = {k: {f"task{t}": g[f"onsets/{k}/task{t}"] for t in range(len(nseg))} for k in keys}
    return g, keys, nseg, rec, ons, g["valid_subjects"], int(g["T"])


def test_oracle_indexing_matches_reference_golden():
    g, keys, nseg, rec, ons, vs, T = _world()
    for n, (i, task, iit, onset) in enumerate(g["index"]):
        X, s, idx = DR.get_item(rec, ons, nseg, vs, T, int(i), keys[int(g["rec"][n])])
        assert idx == (int(task), int(iit), int(onset))
        assert np.array_equal(X, g["X"][n]) and s == int(g["subject_idx"][n])
    # boundaries of the cumulative table
    assert DR.segment_to_task(nseg, 0) == (0, 0)
    assert DR.segment_to_task(nseg, nseg[0] - 1) == (nseg[0] - 1, 0)
    assert DR.segment_to_task(nseg, nseg[0]) == (0, 1)
    assert DR.segment_to_task(nseg, sum(nseg) - 1) == (nseg[-1] - 1, len(nseg) - 1)


@pytest.mark.gpu
def test_device_gather_is_bit_exact_with_reference_golden():
    from megb200.data import DeviceSegmentStore
    g, keys, nseg, rec, ons, vs, T = _world()
    store = DeviceSegmentStore(rec, ons, nseg, vs, T, "cuda:0")
    X, subj, index = store.gather(g["index"][:, 0], rec=g["rec"])
    assert torch.equal(X.cpu(), torch.from_numpy(g["X"]))
    assert torch.equal(index.cpu(), torch.from_numpy(g["index"][:, 1:].astype(np.int64)))
    assert np.array_equal(subj.numpy(), g["subject_idx"].astype(np.int32))
    # the reference's random draw of the recording, reproduced from the same NumPy RNG state
    np.random.seed(1234)
    X2, _, _ = store.gather(g["index"][:, 0])
    assert torch.equal(X2.cpu(), torch.from_numpy(g["X"]))


@pytest.mark.gpu
def test_device_gather_large_and_errors():
    from megb200.data import DeviceSegmentStore
    rng = np.random.Generator(np.random.PCG64(3))
    C, T, nseg = 208, 360, [900, 650, 1200, 410]
    keys = [f"subject{s:02d}_sess{e}" for s in range(3) for e in range(2)]
    vs = np.array(["subject02", "subject00", "subject01"])
    rec, ons = {}, {}
    for k in keys:
        rec[k], ons[k] = {}, {}
        for t, n in enumerate(nseg):
            L = 4000 + 37 * t
            rec[k][f"task{t}"] = rng.standard_normal((C, L)).astype(np.float32)
            ons[k][f"task{t}"] = rng.integers(0, L - T + 1, size=n).astype(np.int64)
    store = DeviceSegmentStore(rec, ons, nseg, vs, T, "cuda:0")
    B = 256
    seg = rng.integers(0, sum(nseg), size=B)
    rc = rng.integers(0, len(keys), size=B).astype(np.int32)
    X, subj, index = store.gather(seg, rec=rc)
    for b in range(0, B, 17):
        Xr, sr, idx = DR.get_item(rec, ons, nseg, vs, T, int(seg[b]), keys[int(rc[b])])
        assert np.array_equal(X[b].cpu().numpy(), Xr) and int(subj[b]) == sr
        assert tuple(int(v) for v in index[b].cpu()) == idx
    with pytest.raises(IndexError):
        store.gather([sum(nseg)], rec=[0])
    ons[keys[0]]["task0"][0] = 4000 - T + 1                        # window one sample past the end
    bad = DeviceSegmentStore(rec, ons, nseg, vs, T, "cuda:0")
    with pytest.raises(ValueError):
        bad.gather([0], rec=[0])
