"""Diagnostic: where do the detection kernels spend their time?  %globaltimer stamps at the phase boundaries of
ground_fused_kernel (every CTA), ccl_bitmap_kernel (per frame) and box_fit_kernel (every CTA; the slowest one is the kernel)."""
import importlib, os, sys
os.environ.setdefault("LMOT_FUSE_CCL", "0")     # the CCL phase stamps need the stand-alone kernel (the product fuses it into the ground kernel's tail)
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d-lidar-multi-object-tracking_b200"
lmot = importlib.import_module(PKG)
synth = importlib.import_module(PKG + ".synth")
NAMES = ["start", "bin + exact pass + flush", "barrier 1 passed", "grid slice done", "barrier 2 passed", "labels done", "counts summed", "end"]
CCL = ["start", "planes loaded", "dilated + ids", "pieces linked", "flattened", "pair unions", "flattened again", "ranked", "labels written"]
SCENE = dict(n_objects=150, lattice_pitch=3.8, ped_fraction=0.65)


def ground_table(ctx, title):
    clk = ctx.debug_phase_clock().astype(np.int64)
    t0 = clk[:, 0].min()
    rel = (clk - t0) / 1e3
    print(f"{title}: {len(clk)} CTAs, kernel span {rel.max():.2f} us")
    for k in range(8):
        print(f"  {NAMES[k]:24s} min {rel[:, k].min():7.2f}  median {np.median(rel[:, k]):7.2f}  max {rel[:, k].max():7.2f} us")


for name, cfg, n in (("hdl64_120k", synth.SceneConfig(seed=1, **SCENE), 120000), ("dense_1m", synth.dense_config(seed=7, **SCENE), 1000000)):
    fr = [p[:n] for _, p in synth.frames(cfg, 4)]
    d = torch.from_numpy(np.stack(fr)).cuda()
    prm = lmot.default_params(); prm.pipeline_depth = 1
    ctx = lmot.Lmot(prm)
    ctx.debug_phase_clock()
    for i in range(6):
        ctx.ground_remove_dev(d[i % 4].data_ptr(), n)
    ctx.sync()
    ground_table(ctx, name + " (stand-alone launch)")
    ctx.close()

# batched tick: 8 streams x 120 k points, then the per-frame CCL and the box fitting of the union of clusters
F = 8
streams = [[p for _, p in synth.frames(synth.SceneConfig(seed=1 + s, **SCENE), 6)] for s in range(F)]
dev = [[torch.from_numpy(p).cuda() for p in st] for st in streams]
ctx = lmot.Lmot()
ctx.debug_phase_clock()
for t in range(6):
    ctx.batch_detect_dev([(dev[s][t].data_ptr(), len(streams[s][t])) for s in range(F)])
    ctx.batch_fetch()
ground_table(ctx, f"batched tick, {F} x 120 k")
ccl = ctx.debug_stage_clocks(1).astype(np.int64)[:F]
rel = (ccl[:, :9] - ccl[:, :1]) / 1e3
print("ccl_bitmap_kernel, one CTA per frame (us after the CTA's start; median / max over frames)")
for k in range(1, 9):
    print(f"  {CCL[k]:18s} {np.median(rel[:, k]):6.2f} / {rel[:, k].max():6.2f}")
fit = ctx.debug_stage_clocks(2).astype(np.int64)
fit = fit[fit[:, 0] > 0]
rel = (fit[:, :5] - fit[:, :1].min()) / 1e3
slow = int(np.argmax(rel[:, 3]))
print(f"box_fit_kernel, {len(fit)} CTAs: start min/max {rel[:, 0].min():.2f}/{rel[:, 0].max():.2f}; first cluster: point pass {np.median(rel[:, 1] - rel[:, 0]):.2f} med / "
      f"{(rel[:, 1] - rel[:, 0]).max():.2f} max, fit {np.median(rel[:, 2] - rel[:, 1]):.2f} med / {(rel[:, 2] - rel[:, 1]).max():.2f} max; "
      f"clusters done med {np.median(rel[:, 3]):.2f} max {rel[:, 3].max():.2f} (CTA {slow}: pass {rel[slow, 1] - rel[slow, 0]:.2f}, fit {rel[slow, 2] - rel[slow, 1]:.2f}); end max {rel[:, 4].max():.2f} us")
# the same for a single frame through the pipeline geometry
ctx.tracker_reset()
for t in range(4):
    ctx.frame_dev(dev[0][t].data_ptr(), len(streams[0][t]), 1e5 * (t + 1))
    ctx.frame_fetch()
ground_table(ctx, "single frame inside the pipeline (half of the SMs)")
ccl = ctx.debug_stage_clocks(1).astype(np.int64)[:1]
print("ccl single frame:", " ".join(f"{(ccl[0, k] - ccl[0, 0]) / 1e3:.2f}" for k in range(1, 9)))
fit = ctx.debug_stage_clocks(2).astype(np.int64)
fit = fit[fit[:, 0] > 0]
rel = (fit[:, :5] - fit[:, :1].min()) / 1e3
slow = int(np.argmax(rel[:, 3]))
print(f"box_fit single frame, {len(fit)} CTAs: clusters done med {np.median(rel[:, 3]):.2f} max {rel[:, 3].max():.2f} (CTA {slow}: pass {rel[slow, 1] - rel[slow, 0]:.2f}, fit {rel[slow, 2] - rel[slow, 1]:.2f}); end max {rel[:, 4].max():.2f} us")
ctx.close()
