"""ORACLE — TEST INFRASTRUCTURE ONLY.  Golden fixture for the GOD sensor-layout branch of ``ch_locations_2d``
(reference meg_decoding/utils/layout.py:36-45 -> matlab_utils/load_meg.py:105-143): runs the UNMODIFIED reference on its
own data/GOD/montage.csv + ch_region.json and stores the inputs (montage table, region table) and the (C, 2) output in
tests/golden/layout_god.npz.   python oracle/make_golden_layout.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "meg-decoding_b200"))
sys.path.insert(0, ROOT)
from megb200 import synth  # noqa: E402
from oracle import make_golden as MG  # noqa: E402


def main():
    M, _ = MG.import_reference()
    import meg_decoding.utils.layout as RL
    info = json.load(open(os.path.join(MG.REF, "data/GOD/ch_region.json")))
    regions = [f"{r}/{s}" for r in info for s in info[r]]
    out = {}
    for tag, region in (("all", regions), ("first2", regions[:2])):
        args = synth.Args(dataset="GOD", root_dir=".", montage_path=os.path.join(MG.REF, "data/GOD/montage.csv"),
                          ch_region_path=os.path.join(MG.REF, "data/GOD/ch_region.json"), region=region)
        loc = RL.ch_locations_2d(args).numpy()
        out["loc_" + tag] = loc
        out["region_" + tag] = np.array(region)
        print(tag, len(region), "regions ->", loc.shape)
    out["montage"] = np.loadtxt(os.path.join(MG.REF, "data/GOD/montage.csv"), delimiter=",")
    out["ch_region_json"] = np.array(json.dumps(info))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "layout_god.npz"), **out)


if __name__ == "__main__":
    main()
