"""Worker of tests/test_shared_tracker_nccl_gpu.py (one process per GPU, launched by torch.distributed.run): the C++ shared tracker
(host/shared_tracker.cpp -> liblmot_shared.so: NCCL all-gather of device box lists, tracker on the owner, NCCL broadcast of the
track table) against ONE reference tracker fed the same boxes on the owner rank."""
import hashlib
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d-lidar-multi-object-tracking_b200"


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lmot = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth")
    from oracle import ref as oracle
    ref = oracle.RefOracle("intended") if oracle.have_ref("intended") else oracle.PortOracle("intended")
    uid = [lmot.shared_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx = lmot.Lmot(device=local)
    st = lmot.SharedTracker(ctx, rank, world, 0, uid[0])

    def ref_boxes(pts):
        e, _ = ref.ground_remove(pts); grid, k = ref.component_clustering(e); b, _ = ref.box_fitting(e, grid, k)
        return b

    def tables_agree(tag):
        d = ctx.tracker_dump()
        h = hashlib.sha1(np.ascontiguousarray(d).tobytes()).hexdigest() + f":{len(d)}"
        hs = [None] * world
        dist.all_gather_object(hs, h)
        assert len(set(hs)) == 1, (tag, hs)
        return len(d)

    n_ticks = 6
    cfgs = [synth.SceneConfig(seed=50 + s, n_objects=30, rings=32, azimuths=900) for s in range(world)]
    streams = [list(synth.frames(c, n_ticks)) for c in cfgs]            # every rank renders every stream (the owner needs them for the oracle)
    # ---- mode "streams": world sensors per tick, ONE tracker step on the concatenation (BASELINE.json configs[3] across GPUs)
    ref.tracker_reset()
    for t in range(n_ticks):
        ts, pts = streams[rank][t]
        d = torch.from_numpy(pts).cuda()
        out = st.tick_dev(d.data_ptr(), len(pts), ts, mode=lmot.SHARED_STREAMS)
        if rank == 0:
            allb = np.concatenate([ref_boxes(streams[s][t][1]) for s in range(world)])
            a = ref.tracker_step(allb, ts)
            assert np.array_equal(out["track_manage"], a["track_manage"]), ("streams", t)
            assert np.array_equal(out["is_static"], a["is_static"]) and np.array_equal(out["is_vis"], a["is_vis"])
            np.testing.assert_allclose(out["targets"], a["targets"], rtol=1e-4, atol=1e-4)
        else:
            assert "n_tracks" in out
        nt = tables_agree(("streams", t))
    assert nt > 5
    us = st.last_us()
    # ---- mode "frames": ONE sensor, world consecutive frames per tick, folded in frame order (configs[4])
    ctx.tracker_reset()
    ref.tracker_reset()
    dist.barrier()
    one = list(synth.frames(synth.SceneConfig(seed=60, n_objects=30, rings=32, azimuths=900), n_ticks * world))
    for t in range(n_ticks):
        ts0 = one[t * world][0]
        ts, pts = one[t * world + rank]
        d = torch.from_numpy(pts).cuda()
        out = st.tick_dev(d.data_ptr(), len(pts), ts0, mode=lmot.SHARED_FRAMES, frame_dt_us=synth.DT_US)
        if rank == 0:
            for r in range(world):
                a = ref.tracker_step(ref_boxes(one[t * world + r][1]), one[t * world + r][0])
            assert np.array_equal(out["track_manage"], a["track_manage"]), ("frames", t)
            np.testing.assert_allclose(out["targets"], a["targets"], rtol=1e-4, atol=1e-4)
        tables_agree(("frames", t))
    st.close()
    ctx.close()
    dist.barrier()
    if rank == 0:
        print("NCCL_SHARED_OK", {k: round(v, 1) for k, v in us.items()})
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
