"""ORACLE — TEST INFRASTRUCTURE ONLY ("next" row N3: segment indexing of the Gwilliams datasets).

numpy restatement of ``Gwilliams2022DatasetBase.segment_to_task`` and the slicing part of ``__getitem__``
(``meg_decoding/dataclass/gwilliams2022.py:130-152``):

    nseg_task_accum = np.cumsum(num_segments_foreach_task)
    task            = np.searchsorted(nseg_task_accum, i + 1)            # side="left"
    i_in_task       = i - np.sum(num_segments_foreach_task[:task])
    onset           = meg_onsets[key][f"task{task}"][i_in_task]
    X               = recordings[key][f"task{task}"][:, onset : onset + seq_len_samp]
    subject_idx     = np.where(valid_subjects == key.split("_")[0])[0][0]

The recording key is drawn by the caller (the reference uses ``np.random.choice(list(self.X.keys()))``, ``:134``).
Pinned by oracle/make_golden_dataset.py, which executes the reference's own two functions (extracted from the source
file, because importing the module needs mne / torchaudio / natsort and starts a multiprocessing Manager).
Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np


def segment_to_task(num_segments_foreach_task, i: int):
    """gwilliams2022.py:145-152 -> (i_in_task, task index)."""
    n = np.asarray(num_segments_foreach_task, dtype=np.int64)
    accum = np.cumsum(n)
    task = int(np.searchsorted(accum, i + 1))
    return int(i - n[:task].sum()), task


def get_item(recordings, meg_onsets, num_segments_foreach_task, valid_subjects, seq_len_samp: int, i: int, key: str):
    """gwilliams2022.py:130-143 with the random key made explicit -> (X [C, seq_len_samp], subject_idx, (task, i_in_task, onset))."""
    i_in_task, task = segment_to_task(num_segments_foreach_task, i)
    onset = int(meg_onsets[key][f"task{task}"][i_in_task])
    X = recordings[key][f"task{task}"][:, onset:onset + seq_len_samp]
    subject_idx = int(np.where(np.asarray(valid_subjects) == key.split("_")[0])[0][0])
    return X, subject_idx, (task, i_in_task, onset)
