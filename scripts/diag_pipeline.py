"""Diagnostic (not part of the bench contract): where does the frame pipeline spend its time?  Run on a GPU box."""
import importlib, json, os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", os.environ.get("DIAG_CONN", "32"))
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lmot = importlib.import_module("3d-lidar-multi-object-tracking_b200")
synth = importlib.import_module("3d-lidar-multi-object-tracking_b200.synth")
import bench

K, W = 300, 20
ts, frames = bench.make_frames(synth, W + K)
n = frames.shape[1]
h = torch.from_numpy(frames).pin_memory(); d = h.cuda(); hn = h.numpy()
torch.cuda.synchronize()
res = {}
for depth in (1, 2, 4, 8):
    prm = lmot.default_params(); prm.pipeline_depth = depth; prm.result_ring = int(os.environ.get('DIAG_RING', '32'))
    ctx = lmot.Lmot(prm)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
    # A: device-resident, host never blocks
    ctx.tracker_reset()
    for i in range(W): ctx.frame_dev(d[i].data_ptr(), n, ts[i])
    torch.cuda.synchronize(); ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(st)
    for i in range(W, W + K): ctx.frame_dev(d[i].data_ptr(), n, ts[i])
    t_host = time.perf_counter() - t0
    ctx.flush(); e1.record(st); torch.cuda.synchronize()
    res[f"A_dev_depth{depth}"] = dict(fps=K / (e0.elapsed_time(e1) * 1e-3), host_submit_us=1e6 * t_host / K)
    ctx.frame_fetch()
    # B: host frames, submit/collect
    ctx.tracker_reset()
    for i in range(W): ctx.frame(hn[i], ts[i])
    t0 = time.perf_counter(); inflight = 0; tsub = tcol = 0.0
    for i in range(W, W + K):
        if inflight == prm.result_ring - 1:
            a = time.perf_counter(); ctx.frame_collect(); tcol += time.perf_counter() - a; inflight -= 1
        a = time.perf_counter(); ctx.frame_submit(hn[i], ts[i]); tsub += time.perf_counter() - a; inflight += 1
    while inflight: ctx.frame_collect(); inflight -= 1
    res[f"B_host_depth{depth}"] = dict(fps=K / (time.perf_counter() - t0), submit_us=1e6 * tsub / K, collect_us=1e6 * tcol / K)
    # C: device frames but collect every frame (same host pattern as B, no H2D)
    ctx.tracker_reset()
    for i in range(W): ctx.frame_dev(d[i].data_ptr(), n, ts[i])
    ctx.frame_fetch()
    ctx.debug_host_ns()
    t0 = time.perf_counter(); inflight = 0
    for i in range(W, W + K):
        if inflight == prm.result_ring - 1: ctx.frame_collect(); inflight -= 1
        ctx.frame_dev(d[i].data_ptr(), n, ts[i]); inflight += 1
    while inflight: ctx.frame_collect(); inflight -= 1
    hn_ = ctx.debug_host_ns()
    res[f"C_dev_collect_depth{depth}"] = dict(fps=K / (time.perf_counter() - t0), lib_submit_us=hn_[0] / 1e3 / K, lib_wait_us=hn_[1] / 1e3 / K, lib_copy_us=hn_[2] / 1e3 / K)
    # per-kernel warm times (synchronous frames)
    if depth == 1:
        ctx.tracker_reset(); ctx.enable_timing(True)
        acc = None
        for i in range(W + 100):
            ctx.frame_dev(d[i].data_ptr(), n, ts[i]); ctx.frame_fetch(want_boxes=False)
            km = np.array(ctx.last_kernel_ms())
            if i >= W and len(km) == len(bench.KERNEL_NAMES): acc = km if acc is None else acc + km
        ctx.enable_timing(False)
        res["kernel_us_warm"] = None if acc is None else {k: float(1e3 * v / 100) for k, v in zip(bench.KERNEL_NAMES, acc)}
        res["n_kernel_events_last"] = len(km)
    ctx.close()
# D: pure H2D of one frame per step
x = torch.empty_like(d[0]); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(K): x.copy_(h[i % (W + K)], non_blocking=True)
e1.record(); torch.cuda.synchronize()
res["D_h2d_only_fps"] = K / (e0.elapsed_time(e1) * 1e-3)
print(json.dumps(res, indent=1))
