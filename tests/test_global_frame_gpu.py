"""GPU parity of lmot_params.global_frame: the reference's `tracking` node moves every box into a dead-reckoned
"global" frame with tf before immUkfJpdaf and moves targetPoints / visBBs back afterwards
(/root/reference/object_tracking/tracking/main.cpp:76-83 broadcast, :142-158 boxes -> /global, :182-195 results -> /velodyne).
With the flag on the library does that on the device.  Oracle: the reference's own immUkfJpdaf fed boxes transformed by a NumPy
restatement of the same arithmetic (tf::Transform built from egoPoints[0] in double, narrowed to the float 4x4 that
pcl_ros::transformPointCloud multiplies with; tf / pcl_ros are not in /root/reference, so this arithmetic is the library's documented
contract, csrc/tracker.cu global_frame_xf / xf_apply).  Bars: identical trackManage / static / visible flags with a MOVING ego,
targets / visBBs / velocities within 1e-4.
"""
import numpy as np
import pytest

from test_tracker_gpu import _boxes_sequence

pytestmark = pytest.mark.gpu
f32 = np.float32


def xf_pair(ego):
    """(fwd: sensor -> global, back: global -> sensor), each 6 float32: rows (m0 m1 . m2), (m3 m4 . m5)"""
    x, y, yaw = float(ego[0]), float(ego[1]), float(ego[2])
    qz, qw = np.sin(yaw * 0.5), np.cos(yaw * 0.5)
    s = 2.0 / (qz * qz + qw * qw)
    zs = qz * s
    wz, zz = qw * zs, qz * zs
    r00, r01, r10, r11 = 1.0 - zz, -wz, wz, 1.0 - zz
    back = np.array([r00, r01, x, r10, r11, y], np.float64).astype(f32)
    tx, ty = -(r00 * x + r10 * y), -(r01 * x + r11 * y)
    fwd = np.array([r00, r10, tx, r01, r11, ty], np.float64).astype(f32)
    return fwd, back


def xf_apply(m, pts):
    p = np.asarray(pts, f32).reshape(-1, 3)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    zero, one = f32(0), f32(1)
    ox = ((m[0] * x + m[1] * y) + zero * z) + m[2]
    oy = ((m[3] * x + m[4] * y) + zero * z) + m[5]
    oz = ((zero * x + zero * y) + one * z) + zero
    return np.stack([ox, oy, oz], 1).astype(f32).reshape(np.shape(pts))


def _ego(f):
    return 5.0 + 0.3 * np.sin(0.4 * f), 0.015 * f + 0.002 * f * f      # v_gps 5 m/s, drifting yaw


def _ctx(pkg):
    p = pkg.default_params()
    p.global_frame = 1
    return pkg.Lmot(p)


def test_track_step_in_global_frame_matches_reference_on_transformed_boxes(pkg, ref_intended, synth):
    ref = ref_intended
    seq = _boxes_sequence(ref, synth, seed=13, n_frames=25, n_objects=50)
    ctx = _ctx(pkg)
    try:
        ref.tracker_reset()
        moved = 0.0
        for f, (ts, boxes) in enumerate(seq):
            v, yaw = _ego(f)
            ego = ctx.origin_points(ts, v, yaw)
            fwd, back = xf_pair(ego)
            a = ref.tracker_step(xf_apply(fwd, boxes), ts, v, yaw)
            b = ctx.track_step(boxes, ts, v, yaw)
            for k in ("track_manage", "is_static", "is_vis"):
                assert np.array_equal(a[k], b[k]), (f, k)
            np.testing.assert_allclose(b["targets"], xf_apply(back, a["targets"]), rtol=1e-4, atol=1e-4)
            np.testing.assert_allclose(b["vandyaw"], a["vandyaw"], rtol=1e-4, atol=1e-5)
            assert a["vis_bb"].shape == b["vis_bb"].shape
            np.testing.assert_allclose(b["vis_bb"], xf_apply(back, a["vis_bb"]), rtol=1e-4, atol=1e-4)
            moved = max(moved, float(np.hypot(ego[0], ego[1])))
        assert moved > 5.0 and (a["track_manage"] > 0).sum() > 10      # the ego really travelled, tracks really lived
        # ... and the flag changes the answer: sensor-frame tracking of the same boxes ends in a different table
        plain = pkg.Lmot()
        try:
            for f, (ts, boxes) in enumerate(seq):
                c = plain.track_step(boxes, ts, *_ego(f))
            assert not (len(c["track_manage"]) == len(b["track_manage"]) and np.array_equal(c["is_static"], b["is_static"])
                        and np.allclose(c["vandyaw"], b["vandyaw"], atol=1e-3))
        finally:
            plain.close()
    finally:
        ctx.close()


def test_frame_pipeline_in_global_frame(pkg, ref_intended, synth):
    """the same through lmot_frame_submit / collect (boxes transformed on the detection stream, semaphore posted by that kernel)
    and through a batched tick"""
    ref = ref_intended
    ctx = _ctx(pkg)
    try:
        ref.tracker_reset()
        frames = list(synth.frames(synth.SceneConfig(seed=17, n_objects=40), 10))
        want = []
        probe = _ctx(pkg)          # a second context only to peek the ego pose sequence (same fold)
        try:
            for f, (ts, pts) in enumerate(frames):
                v, yaw = _ego(f)
                e, g = ref.ground_remove(pts); grid, k = ref.component_clustering(e); boxes, _ = ref.box_fitting(e, grid, k)
                ego = probe.origin_points(ts, v, yaw)
                probe.track_step(np.zeros((0, 8, 3), f32), ts, v, yaw)       # advance its fold
                fwd, back = xf_pair(ego)
                want.append((boxes, ref.tracker_step(xf_apply(fwd, boxes), ts, v, yaw), back))
        finally:
            probe.close()
        for f, (ts, pts) in enumerate(frames):
            ctx.frame_submit(pts, ts, *_ego(f))
        for f, (boxes, a, back) in enumerate(want):
            r = ctx.frame_collect()
            assert np.array_equal(r["boxes"].view(np.uint32), boxes.view(np.uint32)), f      # published boxes stay in the sensor frame
            assert np.array_equal(r["track_manage"], a["track_manage"]), f
            np.testing.assert_allclose(r["targets"], xf_apply(back, a["targets"]), rtol=1e-4, atol=1e-4)
        # batched tick of one stream == the frame path
        ref.tracker_reset(); ctx.tracker_reset()
        for f, (ts, pts) in enumerate(frames[:5]):
            v, yaw = _ego(f)
            ego = ctx.origin_points(ts, v, yaw)
            fwd, back = xf_pair(ego)
            a = ref.tracker_step(xf_apply(fwd, want[f][0]), ts, v, yaw)
            r = ctx.batch([pts], ts, v, yaw)
            assert np.array_equal(r["track_manage"], a["track_manage"]), f
            np.testing.assert_allclose(r["targets"], xf_apply(back, a["targets"]), rtol=1e-4, atol=1e-4)
    finally:
        ctx.close()
