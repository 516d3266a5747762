"""GPU: the frame pipeline (several frames in flight inside one context) returns exactly what frame-at-a-time
calls return -- pipelining only changes when the host sees a result, never the result."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run_sync(pkg, frames, depth):
    prm = pkg.default_params()
    prm.pipeline_depth = depth
    ctx = pkg.Lmot(prm)
    try:
        return [ctx.frame(p, ts) for ts, p in frames]
    finally:
        ctx.close()


def _same(a, b):
    assert (a["n_elevated"], a["n_ground"], a["num_cluster"]) == (b["n_elevated"], b["n_ground"], b["num_cluster"])
    assert np.array_equal(a["boxes"].view(np.uint32), b["boxes"].view(np.uint32))
    for k in ("track_manage", "is_static", "is_vis"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("targets", "vandyaw", "vis_bb"):
        assert np.array_equal(a[k], b[k]), k      # same kernels, same inputs: bit-identical


def test_submit_collect_equals_frame_at_a_time(pkg, synth):
    frames = list(synth.frames(synth.SceneConfig(seed=13, n_objects=90, lattice_pitch=4.5), 14))
    base = _run_sync(pkg, frames, 1)
    for depth, ring in ((2, 2), (4, 3), (4, 32)):
        prm = pkg.default_params()
        prm.pipeline_depth = depth
        prm.result_ring = ring
        ctx = pkg.Lmot(prm)
        depth = ring
        try:
            got = []
            for ts, p in frames:
                if ctx.frames_in_flight() == depth:
                    got.append(ctx.frame_collect())
                ctx.frame_submit(p, ts)
            while ctx.frames_in_flight():
                got.append(ctx.frame_collect())
            assert len(got) == len(base)
            for a, b in zip(got, base):
                _same(a, b)
            with pytest.raises(pkg.LmotError):
                ctx.frame_collect()               # nothing in flight
        finally:
            ctx.close()


@pytest.mark.parametrize("zero_copy", ["0", "1"])
def test_pinned_host_frames(pkg, synth, monkeypatch, zero_copy):
    """lmot_frame_submit with page-locked frames (lmot_pinned_alloc): asynchronous copy (default), or LMOT_ZERO_COPY=1 -- no H2D
    copy at all, the ground kernel's bulk copies read the host buffer over PCIe.  Same results, bit for bit, as pageable
    numpy arrays and as frame-at-a-time calls."""
    monkeypatch.setenv("LMOT_ZERO_COPY", zero_copy)
    frames = list(synth.frames(synth.SceneConfig(seed=15, n_objects=80, lattice_pitch=4.2), 12))
    base = _run_sync(pkg, frames, 1)
    ctx = pkg.Lmot()
    try:
        pin = ctx.pinned_array((len(frames),) + frames[0][1].shape, np.float32)
        for i, (_, p) in enumerate(frames):
            pin[i] = p
        got = []
        for i, (ts, _) in enumerate(frames):
            ctx.frame_submit(pin[i], ts)
        while ctx.frames_in_flight():
            got.append(ctx.frame_collect())
        assert len(got) == len(base)
        for a, b in zip(got, base):
            _same(a, b)
        # ragged sizes and an unaligned view (falls back to the copy path) through the same entry point
        ctx.tracker_reset()
        ref = pkg.Lmot()
        try:
            for i, (ts, p) in enumerate(frames[:4]):
                m = len(p) - 37 * i
                flat = pin.reshape(-1)
                view = flat[1:1 + m * 4].reshape(m, 4)          # 4-byte offset: not 16-byte aligned
                view[:] = p[:m]
                ctx.frame_submit(view, ts)
                _same(ctx.frame_collect(), ref.frame(p[:m], ts))
        finally:
            ref.close()
    finally:
        ctx.close()


def test_frame_dev_async_matches(pkg, synth):
    import torch
    frames = list(synth.frames(synth.SceneConfig(seed=14, n_objects=60), 9))
    base = _run_sync(pkg, frames, 1)
    ctx = pkg.Lmot()
    try:
        st = torch.cuda.Stream()
        ctx.set_stream(st.cuda_stream)
        dev = [torch.from_numpy(p).cuda() for _, p in frames]
        torch.cuda.synchronize()
        for (ts, p), d in zip(frames, dev):
            ctx.frame_dev(d.data_ptr(), len(p), ts)
        last = ctx.frame_fetch()
        _same(last, base[-1])
        # detect-only path shares the slots
        ctx.detect_dev(dev[0].data_ptr(), len(frames[0][1]))
        r = ctx.frame_fetch()
        assert r["n_elevated"] == base[0]["n_elevated"] and r["num_cluster"] == base[0]["num_cluster"]
    finally:
        ctx.close()


def test_stage_calls_interleave_with_pipeline(pkg, synth, ref_intended):
    """Stage-by-stage entry points drain the pipeline first and leave the tracker state untouched."""
    frames = list(synth.frames(synth.SceneConfig(seed=15, n_objects=50), 6))
    base = _run_sync(pkg, frames, 1)
    ctx = pkg.Lmot()
    try:
        got = []
        for i, (ts, p) in enumerate(frames):
            ctx.frame_submit(p, ts)
            if i == 2:
                out = ctx.ground_remove(frames[0][1])          # drains; the in-flight results are dropped ...
                e, _ = ref_intended.ground_remove(frames[0][1])
                assert len(out["elevated"]) == len(e)
            else:
                got.append((i, ctx.frame_collect()))
        for i, r in got:                                        # ... but the tracker has consumed every frame in order
            assert np.array_equal(r["track_manage"], base[i]["track_manage"])
    finally:
        ctx.close()


def test_detect_only_submissions_between_tracked_frames(pkg, synth):
    """Detection-only submissions take result blocks from the same ring as tracked frames; the tracker's incremental outputs
    (each frame's block is seeded from the previous TRACKED frame's block) must not notice them.  Small ring -> wraps."""
    import torch
    frames = list(synth.frames(synth.SceneConfig(seed=16, n_objects=70, lattice_pitch=4.2), 16))
    base = _run_sync(pkg, frames, 1)
    prm = pkg.default_params()
    prm.pipeline_depth = 3
    prm.result_ring = 3
    ctx = pkg.Lmot(prm)
    try:
        st = torch.cuda.Stream()
        ctx.set_stream(st.cuda_stream)
        dev = [torch.from_numpy(p).cuda() for _, p in frames]
        torch.cuda.synchronize()
        for i, ((ts, p), d) in enumerate(zip(frames, dev)):
            ctx.frame_dev(d.data_ptr(), len(p), ts, 2.0 + 0.1 * i, 0.01 * i)
            if i % 3 == 1:
                ctx.detect_dev(dev[(i + 5) % len(dev)].data_ptr(), len(p))
            if i in (6, 15):
                ctx.frame_dev(d.data_ptr(), len(p), ts + 1, 2.0 + 0.1 * i, 0.01 * i)   # undo below: compare against a matching baseline
        last = ctx.frame_fetch()
    finally:
        ctx.close()
    # baseline with the same call sequence, frame at a time
    ctx = pkg.Lmot()
    try:
        for i, (ts, p) in enumerate(frames):
            want = ctx.frame(p, ts, 2.0 + 0.1 * i, 0.01 * i)
            if i in (6, 15):
                want = ctx.frame(p, ts + 1, 2.0 + 0.1 * i, 0.01 * i)
    finally:
        ctx.close()
    _same(last, want)
    assert len(base) == len(frames)
