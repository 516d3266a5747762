"""Diagnostic: where does ground_fused_kernel spend its time?  %globaltimer stamps of every CTA at the phase boundaries."""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d-lidar-multi-object-tracking_b200"
lmot = importlib.import_module(PKG)
synth = importlib.import_module(PKG + ".synth")
NAMES = ["start", "bin done", "barrier1 passed", "grid slice done", "barrier2 passed", "labels done", "counts summed", "end"]
for name, cfg, n in (("hdl64_120k", synth.SceneConfig(n_objects=150, lattice_pitch=3.8, ped_fraction=0.65, seed=1), 120000),
                     ("dense_1m", synth.dense_config(n_objects=150, lattice_pitch=3.8, ped_fraction=0.65, seed=7), 1000000)):
    fr = [p[:n] for _, p in synth.frames(cfg, 4)]
    d = torch.from_numpy(np.stack(fr)).cuda()
    prm = lmot.default_params(); prm.pipeline_depth = 1
    ctx = lmot.Lmot(prm)
    ctx.debug_phase_clock()
    for i in range(6):
        ctx.ground_remove_dev(d[i % 4].data_ptr(), n)
    ctx.sync()
    clk = ctx.debug_phase_clock().astype(np.int64)
    t0 = clk[:, 0].min()
    rel = (clk - t0) / 1e3
    print(f"{name}: {len(clk)} CTAs, kernel span {rel.max():.2f} us")
    for k in range(8):
        print(f"  {NAMES[k]:18s} min {rel[:, k].min():7.2f}  median {np.median(rel[:, k]):7.2f}  max {rel[:, k].max():7.2f} us")
    ctx.close()
