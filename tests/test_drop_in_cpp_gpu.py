"""GPU: the C++ drop-in wrappers (include/lmot_drop_in.hpp -- the reference's own four signatures over the C ABI),
driven by a ROS-free replay of the three node callbacks (tests/cpp/drop_in_replay.cpp), against the reference."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "cpp", "drop_in_replay")


def test_cpp_drop_in_replay_matches_reference(tmp_path, ref_intended, synth):
    if not os.path.exists(EXE):
        pytest.skip("tests/cpp/drop_in_replay not built (python -c 'import __graft_entry__ as g; g.build()')")
    frames = [p for _, p in synth.frames(synth.SceneConfig(rings=32, azimuths=900, n_objects=40, seed=31), 6)]
    fin, fout = tmp_path / "frames.bin", tmp_path / "out.bin"
    np.stack(frames).astype(np.float32).tofile(fin)
    subprocess.run([EXE, str(fin), str(len(frames)), str(len(frames[0])), str(fout)], check=True, timeout=120)
    raw = np.fromfile(fout, np.int32)
    ref = ref_intended
    ref.tracker_reset()
    pos = 0
    for f, pts in enumerate(frames):
        e, g = ref.ground_remove(pts)
        grid, k = ref.component_clustering(e)
        boxes, _ = ref.box_fitting(e, grid, k)
        tr = ref.tracker_step(boxes, (f + 1) * 1e5)
        ne, ng, nc, nb, nt = raw[pos:pos + 5]; pos += 5
        assert (ne, ng, nc, nb, nt) == (len(e), len(g), k, len(boxes), len(tr["track_manage"]))
        got = raw[pos:pos + nb * 24].view(np.float32).reshape(nb, 8, 3); pos += nb * 24
        assert np.array_equal(got.view(np.uint32), boxes.view(np.uint32))
        assert np.array_equal(raw[pos:pos + nt], tr["track_manage"]); pos += nt
    assert pos == len(raw)
