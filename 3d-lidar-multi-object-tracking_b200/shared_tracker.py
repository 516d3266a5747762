"""Several sensor streams (one per rank / GPU) feeding ONE track table -- the only exchange step on this path
(BASELINE.json north_star).

Per tick every rank runs the detection stages on its own frame (nothing is shared there), then:
  1. all ranks all_gather their box counts and their (padded) box lists                      -- NCCL all_gather
  2. the owner rank concatenates the boxes in rank order and runs the tracker step on them   -- lmot_track_step
  3. the owner broadcasts the per-track outputs, and (GPU backend) the track table itself
     (T x sizeof(TrackState) bytes straight from device memory) so that every rank holds it  -- NCCL broadcast

The host logic is backend agnostic: the GPU backend is an `Lmot` context; tests/test_shared_tracker_gloo.py runs the same
code on CPU tensors over gloo with the oracle port as the tracker.  Box order = rank order, then each rank's own order,
which is what a single tracker receiving the concatenated `track_box` messages would see.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


class _DevView:
    """__cuda_array_interface__ wrapper so torch can view liblmot's device track table without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


class SharedTracker:
    def __init__(self, backend, owner: int = 0, max_boxes: int = 1024, device: str | torch.device = "cpu", group=None):
        self.backend, self.owner, self.max_boxes, self.group = backend, owner, max_boxes, group
        self.device = torch.device(device)
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._table = None
        if hasattr(backend, "tracker_table") and self.device.type == "cuda":
            ptr, bpt, cap = backend.tracker_table()
            self._bpt, self._cap = bpt, cap
            self._table = torch.as_tensor(_DevView(ptr, bpt * cap), device=self.device)

    def gather_boxes(self, boxes: np.ndarray):
        """-> (all boxes in rank order (M,8,3) float32, per-rank counts)"""
        b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 8, 3)
        m = b.shape[0]
        if m > self.max_boxes:
            raise ValueError("more boxes than max_boxes")
        counts = [torch.zeros(1, dtype=torch.int32, device=self.device) for _ in range(self.world)]
        dist.all_gather(counts, torch.tensor([m], dtype=torch.int32, device=self.device), group=self.group)
        counts = [int(c.item()) for c in counts]
        pad = torch.zeros((self.max_boxes, 8, 3), dtype=torch.float32, device=self.device)
        if m:
            pad[:m] = torch.from_numpy(b).to(self.device)
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(bufs, pad, group=self.group)
        allb = torch.cat([bufs[r][: counts[r]] for r in range(self.world)], 0) if sum(counts) else pad[:0]
        return allb.cpu().numpy(), counts

    def step(self, boxes: np.ndarray, timestamp_us: float, v_gps: float = 0.0, yaw_gps: float = 0.0, frame_sharded: bool = False,
             frame_dt_us: float = 1.0e5):
        """One tick of the shared tracker.  Every rank gets the same per-track outputs.

        frame_sharded=False: the ranks are different SENSORS observed at the same instant; their boxes are concatenated in
        rank order into ONE tracker step (BASELINE.json configs[3]).
        frame_sharded=True: the ranks hold CONSECUTIVE FRAMES of one sensor (rank r has frame t*world + r, configs[4]); the
        owner folds them into the table in frame order, one tracker step per rank, timestamps timestamp_us + r*frame_dt_us."""
        allb, counts = self.gather_boxes(boxes)
        hdr = torch.zeros(2, dtype=torch.int64, device=self.device)
        out = None
        if self.rank == self.owner:
            if frame_sharded:
                off = 0
                for r in range(self.world):
                    out = self.backend.track_step(allb[off: off + counts[r]], timestamp_us + r * frame_dt_us, v_gps, yaw_gps)
                    off += counts[r]
            else:
                out = self.backend.track_step(allb, timestamp_us, v_gps, yaw_gps)
            hdr[0] = len(out["track_manage"]); hdr[1] = len(out["vis_bb"])
        dist.broadcast(hdr, src=self.owner, group=self.group)
        T, V = int(hdr[0].item()), int(hdr[1].item())
        spec = (("targets", (T, 3), torch.float32), ("vandyaw", (T, 2), torch.float64), ("track_manage", (T,), torch.int32),
                ("is_static", (T,), torch.uint8), ("is_vis", (T,), torch.uint8), ("vis_bb", (V, 8, 3), torch.float32))
        res = {}
        for name, shape, dt in spec:
            t = torch.from_numpy(np.ascontiguousarray(out[name])).to(self.device) if self.rank == self.owner else torch.empty(shape, dtype=dt, device=self.device)
            if t.numel():
                dist.broadcast(t.reshape(-1), src=self.owner, group=self.group)
            res[name] = t.reshape(shape).cpu().numpy()
        if self._table is not None and T > 0:       # the table itself, device to device over NVLink
            dist.broadcast(self._table[: T * self._bpt], src=self.owner, group=self.group)
            if self.rank != self.owner:
                self.backend.tracker_set_num_tracks(T)
        res["box_counts"] = counts
        return res
