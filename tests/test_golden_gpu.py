"""GPU parity against the committed golden vectors (generated from the reference itself by
tests/golden/make_golden.py): works on the GPU box even if oracle/_ref were missing."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_golden.npz")


@pytest.mark.parametrize("mode", ["intended", "o2"])
def test_cuda_path_reproduces_golden(pkg, mode):
    gold = np.load(GOLD)
    prm = pkg.default_params()
    prm.rule_filter = pkg.RULE_INTENDED if mode == "intended" else pkg.RULE_GCC13_O2_COMPAT
    ctx = pkg.Lmot(prm)
    try:
        for f in range(int(gold["n_frames"])):
            pts = gold[f"pts{f}"]
            if mode == "intended":
                out = ctx.ground_remove(pts)
                assert np.array_equal(out["labels"], gold[f"labels{f}"])
                pg = ctx.debug_polar_grid()
                assert np.array_equal(pg["isground"], gold[f"isground{f}"])
                grid, k = ctx.component_cluster(out["elevated"])
                assert k == int(gold[f"ncluster{f}"]) and np.array_equal(grid, gold[f"grid{f}"].astype(np.int32))
            r = ctx.frame(pts, float(gold[f"ts{f}"]))
            assert np.array_equal(r["boxes"].view(np.uint32), gold[f"boxes_{mode}{f}"].view(np.uint32))
            assert np.array_equal(r["track_manage"], gold[f"manage_{mode}{f}"])
            d, dg = ctx.tracker_dump(), gold[f"dump_{mode}{f}"]
            assert np.array_equal(d[:, :4], dg[:, :4])
            live = dg[:, 0] > 0
            if live.any():
                a, b = d[live][:, 4:175], dg[live][:, 4:175]
                assert (np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-3)).max() < 1e-4
    finally:
        ctx.close()
