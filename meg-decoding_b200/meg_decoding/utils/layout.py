"""Sensor layout for the spatial attention (reference ``meg_decoding/utils/layout.py:6-47``).

The reference reads the layout from raw recordings through mne / mne_bids (not installed here, data absent).  The hot
path only needs the (C, 2) coordinates, min-max normalised and scaled into [0.1, 0.9]; they can be supplied directly
through ``args.sensor_locations`` (array-like (C, 2), raw coordinates).  Otherwise the reference's branches run: the GOD
branch (layout.py:36-38) reads ``args.montage_path`` / ``args.ch_region_path`` / ``args.region`` like
``matlab_utils/load_meg.py:105-143``; the Brennan / Gwilliams branches need mne (+ the raw data), as in the reference.
"""
import numpy as np
import torch


def normalise_locations(loc) -> torch.Tensor:
    """layout.py:41-47 — min-max per axis, then a 0.1 margin on every side."""
    loc = np.asarray(loc, dtype=np.float64)
    loc = (loc - loc.min(axis=0)) / (loc.max(axis=0) - loc.min(axis=0))
    loc = loc * 0.8 + 0.1
    return torch.from_numpy(loc.astype(np.float32))


def _get(args, name, default=None):
    try:
        return args.get(name, default) if hasattr(args, "get") else getattr(args, name, default)
    except Exception:
        return default


def roi_channels(args):
    """matlab_utils/load_meg.py:105-121 — 0-based channel indices of the requested "region/subregion" entries of the
    channel-region table (the JSON lists MATLAB's 1-based channel numbers)."""
    import json
    with open(args.ch_region_path, "r") as f:
        info = json.load(f)
    out = []
    for reg in args.region:
        parts = reg.split("/")
        assert len(parts) == 2, "got {}".format(parts)
        out += info[parts[0]][parts[1]]
    return [r - 1 for r in out]


def read_montage(args) -> np.ndarray:
    """matlab_utils/load_meg.py:133-143 — sensor positions (x, y, z) of the ROI channels from the montage CSV."""
    import csv
    rows = []
    with open(args.montage_path) as f:
        for row in csv.reader(f):
            rows.append([float(r) for r in row])
    return np.array(rows)[roi_channels(args), :]


def ch_locations_2d(args):
    given = None
    try:
        given = args.get("sensor_locations") if hasattr(args, "get") else getattr(args, "sensor_locations", None)
    except Exception:
        given = None
    if given is not None:
        return normalise_locations(np.asarray(given)[:, :2])
    dataset_name = args.dataset
    if dataset_name == "GOD":                                    # layout.py:36-38: montage file, no mne involved
        return normalise_locations(read_montage(args)[:, :2])
    try:
        import mne  # noqa: F401
    except ImportError as e:
        raise RuntimeError(
            "megb200: the sensor layout of {!r} needs mne (and the raw data); pass the (C, 2) coordinates as "
            "args.sensor_locations instead".format(dataset_name)) from e
    if dataset_name == "Brennan2018":
        montage = mne.channels.make_standard_montage("easycap-M10")
        info = mne.create_info(ch_names=montage.ch_names, sfreq=512.0, ch_types="eeg")
        info.set_montage(montage)
        loc = np.delete(mne.channels.find_layout(info, ch_type="eeg").pos[:, :2], 28, axis=0)
    elif dataset_name == "Gwilliams2022":
        import mne_bids
        bids_path = mne_bids.BIDSPath(subject="01", session="0", task="0", datatype="meg",
                                      root=f"{args.root_dir}/data/Gwilliams2022/")
        loc = mne.channels.find_layout(mne_bids.read_raw_bids(bids_path).info, ch_type="meg").pos[:, :2]
    else:
        raise ValueError(dataset_name)
    return normalise_locations(loc)
