"""CPU, world_size 2 over gloo: the host logic of the multi-stream shared tracker (gather boxes in rank order ->
owner runs the tracker -> broadcast outputs).  The tracker backend here is the oracle port (this is a test);
on the GPU box the same class drives an Lmot context over NCCL (bench.py --shared-tracker)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _PortBackend:
    def __init__(self):
        from oracle import ref as oracle
        self.o = oracle.PortOracle("intended")
        self.o.tracker_reset()

    def track_step(self, boxes, ts, v, yaw):
        return self.o.tracker_step(boxes, ts, v, yaw)


def _boxes_for(rank, frame):
    rng = np.random.default_rng(100 * rank + frame)
    n = 3 + rank * 2
    c = np.stack([np.arange(n) * 6.0 - 10.0 + 0.1 * frame, np.full(n, 8.0 * (1 - 2 * rank))], 1) + rng.normal(0, 0.02, (n, 2))
    out = np.zeros((n, 8, 3), np.float32)
    half = np.array([[-1.0, -0.5], [-1.0, 0.5], [1.0, 0.5], [1.0, -0.5]])
    for k in range(4):
        out[:, k, :2] = c + half[k]; out[:, k, 2] = -2.0
        out[:, 4 + k, :2] = c + half[k]; out[:, 4 + k, 2] = 0.0
    return out


def _worker_sharded(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    st_mod = importlib.import_module("3d-lidar-multi-object-tracking_b200.shared_tracker")
    backend = _PortBackend() if rank == 0 else object()
    st = st_mod.SharedTracker(backend, owner=0, max_boxes=64, device="cpu")
    results = []
    for t in range(5):                      # tick t: rank r holds frame t*world + r of ONE sensor
        f = t * world + rank
        r = st.step(_boxes_for(0, f), (t * world + 1) * 1e5, frame_sharded=True, frame_dt_us=1e5)
        results.append((r["track_manage"].copy(), r["targets"].copy()))
    q.put((rank, results))
    dist.destroy_process_group()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    st_mod = importlib.import_module("3d-lidar-multi-object-tracking_b200.shared_tracker")
    backend = _PortBackend() if rank == 0 else object()
    st = st_mod.SharedTracker(backend, owner=0, max_boxes=64, device="cpu")
    results = []
    for f in range(8):
        r = st.step(_boxes_for(rank, f), (f + 1) * 1e5)
        results.append((r["track_manage"].copy(), r["targets"].copy(), r["box_counts"]))
    q.put((rank, results))
    dist.destroy_process_group()


def test_two_streams_one_tracker_gloo():
    from oracle import ref as oracle
    if not oracle.have_port():
        pytest.skip("oracle port not built")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference: one tracker fed with the concatenation in rank order
    o = oracle.PortOracle("intended")
    o.tracker_reset()
    for f in range(8):
        allb = np.concatenate([_boxes_for(0, f), _boxes_for(1, f)], 0)
        want = o.tracker_step(allb, (f + 1) * 1e5)
        for rank in (0, 1):
            tm, tg, counts = got[rank][f]
            assert counts == [3, 5]
            assert np.array_equal(tm, want["track_manage"])
            assert np.array_equal(tg, want["targets"])


def test_frame_sharded_one_sensor_gloo():
    """configs[4] shape: consecutive frames of one sensor live on different ranks; the owner folds them in frame order."""
    from oracle import ref as oracle
    if not oracle.have_port():
        pytest.skip("oracle port not built")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = oracle.PortOracle("intended")
    o.tracker_reset()
    for t in range(5):
        for r in range(2):
            f = t * 2 + r
            want = o.tracker_step(_boxes_for(0, f), (f + 1) * 1e5)
        for rank in (0, 1):
            tm, tg = got[rank][t]
            assert np.array_equal(tm, want["track_manage"]) and np.array_equal(tg, want["targets"])
