#!/usr/bin/env python
"""One RoIAlign launch per shape for ncu captures: python tools/prof_roi.py [K_gt] (default 1500 GT x 64)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_teacher_b200 import ops, synth  # noqa: E402
from point_teacher_b200.proposals import boxes_to_rois, fine_proposals_from_cfg  # noqa: E402

n_gt = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
dev = torch.device("cuda")
ds = synth.hbb_batch(seed=1, batch=1, gt_range=(n_gt, n_gt))
props, _ = fine_proposals_from_cfg([b.to(dev) for b in ds["pseudo_boxes"]], synth.stress_ext_cfg(8)[0], ds["img_metas"])
rs = boxes_to_rois(props)
feat = ops.nchw_to_nhwc(ds["feat"].to(dev), torch.bfloat16)
out = torch.empty((rs.shape[0], 12544), dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.roi_align_forward(feat, rs, 0, 0.125, out=out)
torch.cuda.synchronize()
print("ok", rs.shape)
