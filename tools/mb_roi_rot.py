"""Micro-benchmark: RoIAlignRotated (bf16 tensor-core path) on the config #3 RoI sets -- the 5000 second-level bags
(pass 1) and bags + 400 large negatives (pass 2) -- CUDA events, L2 flushed, median of 20.
Usage: [PTB200_ROT_CFG=43] python tools/mb_roi_rot.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_teacher_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda:0")
d = synth.obb_batch(seed=0)
cap = 100
pb = torch.cat([b[:cap] for b in d["pseudo_boxes"]]).to(dev)
idx = torch.tensor([i for i in range(2) for _ in range(cap)], dtype=torch.int32, device=dev)
wh = torch.tensor([[1024., 1024.]] * 2, device=dev)
base, _ = ops.bag_gen(ops.make_rois(pb, idx), wh, [1.0], None, 0, True)
cfg = synth.OBB_EXT_CFG[0]
bags, _ = ops.bag_gen(base, wh, cfg["base_ratios"], cfg["shake_ratio"], cfg["min_scale"], True)
negs = torch.cat(d["neg_boxes"][0]).to(dev)
nidx = torch.tensor([i for i in range(2) for _ in range(200)], dtype=torch.int32, device=dev)
nrois = ops.make_rois(negs, nidx)
both = torch.cat([bags, nrois]).contiguous()
fx = d["feat"].to(dev)
feat = ops.nchw_to_nhwc(fx, torch.float16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(rois, tag):
    out = torch.empty((rois.shape[0], 12544), dtype=torch.bfloat16, device=dev)
    ts = []
    for i in range(25):
        flush.zero_()
        feat.copy_(ops.nchw_to_nhwc(fx, torch.float16))          # as in the step: the map was just written, i.e. L2-warm
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200000)
        e0.record()
        ops.roi_align_forward(feat, rois, ops.OUT_BF16_BINMAJOR, 0.125, 2, True, rotated=True, clockwise=True, out=out)
        e1.record()
        torch.cuda.synchronize()
        if i >= 5:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    nbytes = rois.shape[0] * 12544 * 2 + feat.numel() * 2 + rois.numel() * 4
    med = ts[len(ts) // 2]
    print(f"cfg={os.environ.get('PTB200_ROT_CFG', 'default')} {tag}: K={rois.shape[0]} median {med:.1f} us  "
          f"{nbytes / med / 1e3:.0f} GB/s  frac {nbytes / med / 1e3 / 6554.6:.3f}")


run(bags, "bags")
run(nrois, "negatives")
run(both, "bags+negatives")
