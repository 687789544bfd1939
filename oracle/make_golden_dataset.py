"""ORACLE — TEST INFRASTRUCTURE ONLY.  Golden vectors for "next" row N3 from the reference's OWN functions.

``Gwilliams2022DatasetBase.__getitem__`` and ``.segment_to_task`` are cut out of
/root/reference/meg_decoding/dataclass/gwilliams2022.py with ``ast`` and executed unchanged on a stand-in ``self``
(importing the module is not possible here: mne, torchaudio, natsort, a multiprocessing Manager at import time).
``np.random.choice`` is seeded so that the drawn recording keys can be stored.  Asserts that
oracle/dataset_restatement.py returns the same windows and indices, and writes tests/golden/dataset.npz.

    python oracle/make_golden_dataset.py
"""
from __future__ import annotations

import ast
import os
import sys
import types
from typing import Tuple  # noqa: F401  (used by the extracted source)

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dataset_restatement as DR  # noqa: E402

SRC = "/root/reference/meg_decoding/dataclass/gwilliams2022.py"


def reference_methods():
    tree = ast.parse(open(SRC).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Gwilliams2022DatasetBase")
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("__getitem__", "segment_to_task")]
    assert len(fns) == 2
    mod = ast.Module(body=fns, type_ignores=[])
    ns = {"np": np, "Tuple": Tuple}
    exec(compile(mod, SRC, "exec"), ns)
    return ns["__getitem__"], ns["segment_to_task"]


def make_world(seed=0):
    g = np.random.Generator(np.random.PCG64(seed))
    C, T = 6, 24
    nseg = [5, 1, 7, 3]                                     # ragged task sizes incl. a single-segment task
    keys = ["subject01_sess0", "subject01_sess1", "subject03_sess0"]
    valid_subjects = np.array(["subject03", "subject01"])  # the reference's set() order is arbitrary: keep it explicit
    recordings, onsets = {}, {}
    for k in keys:
        recordings[k], onsets[k] = {}, {}
        for t, n in enumerate(nseg):
            length = int(g.integers(T + 40, T + 120))
            recordings[k][f"task{t}"] = g.standard_normal((C, length)).astype(np.float32)
            o = np.sort(g.integers(0, length - T + 1, size=n))
            o[-1] = length - T                              # a window that ends exactly at the end of the recording
            onsets[k][f"task{t}"] = o.astype(np.int64)
    return C, T, nseg, keys, valid_subjects, recordings, onsets


def main():
    getitem, seg2task = reference_methods()
    C, T, nseg, keys, valid_subjects, recordings, onsets = make_world()
    total = sum(nseg)
    self = types.SimpleNamespace(X=recordings, meg_onsets=onsets, num_segments_foreach_task=nseg, seq_len_samp=T,
                                 valid_subjects=valid_subjects, Y=np.zeros((total, 1), dtype=np.float32))
    self.segment_to_task = lambda i: seg2task(self, i)
    out = {"nseg": np.array(nseg), "T": np.array(T), "keys": np.array(keys), "valid_subjects": valid_subjects}
    Xs, subj, drawn, idx = [], [], [], []
    np.random.seed(1234)
    order = list(range(total)) + [0, total - 1, 4, 5, 6, 12, 13]        # every segment + the task boundaries again
    for i in order:
        state = np.random.get_state()
        X, _, s = getitem(self, i)
        np.random.set_state(state)
        key = str(np.random.choice(list(recordings.keys())))            # the same draw the reference just made
        Xr, sr, (task, iit, onset) = DR.get_item(recordings, onsets, nseg, valid_subjects, T, i, key)
        assert X.shape == (C, T) and np.array_equal(X, Xr) and int(s) == sr, i
        assert (iit, f"task{task}") == seg2task(self, i), i
        Xs.append(X); subj.append(int(s)); drawn.append(keys.index(key)); idx.append((i, task, iit, onset))
    out.update(X=np.stack(Xs), subject_idx=np.array(subj), rec=np.array(drawn), index=np.array(idx))
    for k in keys:
        for t in range(len(nseg)):
            out[f"rec/{k}/task{t}"] = recordings[k][f"task{t}"]
            out[f"onsets/{k}/task{t}"] = onsets[k][f"task{t}"]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dataset.npz"), **out)
    print(f"dataset golden written: {len(order)} items, restatement == reference bit for bit")


if __name__ == "__main__":
    main()
