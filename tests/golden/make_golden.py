"""Generates tests/golden/hotpath_golden.npz from the REFERENCE ITSELF (oracle/_ref = the unmodified sources of
/root/reference/object_tracking compiled by oracle/Makefile).  Run in the build container, where /root/reference
exists:   python tests/golden/make_golden.py
The fixture pins, per frame of a small seeded scene: per-point labels, the 80x120 grid flags, the 250x250 label grid,
the boxes (both ruleBasedFilter modes) and the full tracker table after every frame.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
synth = importlib.import_module("3d-lidar-multi-object-tracking_b200.synth")
from oracle.ref import RefOracle, labels_from_clouds  # noqa: E402

N_FRAMES = 6


def main():
    out = {}
    for mode in ("intended", "o2"):
        ref = RefOracle(mode)
        ref.tracker_reset()
        cfg = synth.SceneConfig(rings=24, azimuths=500, n_objects=40, lattice_pitch=4.5, ped_fraction=0.5, seed=21)
        for f, (ts, pts) in enumerate(synth.frames(cfg, N_FRAMES)):
            e, g = ref.ground_remove(pts)
            grid, k = ref.component_clustering(e)
            boxes, markers = ref.box_fitting(e, grid, k)
            tr = ref.tracker_step(boxes, ts)
            if mode == "intended":
                out[f"pts{f}"] = pts
                out[f"ts{f}"] = np.float64(ts)
                out[f"labels{f}"] = labels_from_clouds(pts, e, g)
                pg = ref.polar_grid(pts)
                out[f"isground{f}"] = pg["isground"]
                out[f"hground{f}"] = pg["hground"]
                out[f"grid{f}"] = grid.astype(np.int16)
                out[f"ncluster{f}"] = np.int32(k)
            out[f"boxes_{mode}{f}"] = boxes
            out[f"markers_{mode}{f}"] = markers
            out[f"manage_{mode}{f}"] = tr["track_manage"]
            out[f"targets_{mode}{f}"] = tr["targets"]
            out[f"vandyaw_{mode}{f}"] = tr["vandyaw"]
            out[f"dump_{mode}{f}"] = ref.tracker_dump()
    out["n_frames"] = np.int32(N_FRAMES)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hotpath_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
