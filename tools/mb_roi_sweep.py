"""RoIAlign (bf16 tensor-core path, horizontal) over the number of RoIs K: time(K) = fixed + K * per_roi.  Separates
the launch / pipeline-fill / drain cost of a launch from the steady per-RoI rate (HBM-store-bound), i.e. what bounds
the roofline fraction at the bench's K = 5000.  CUDA events, L2 flushed then the feature map re-written (L2-warm, as
in the step), median of 15.  Usage: python tools/mb_roi_sweep.py [rot]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_teacher_b200 import ops, synth  # noqa: E402

rot = len(sys.argv) > 1 and sys.argv[1] == "rot"
dev = torch.device("cuda:0")
d = synth.obb_batch(seed=1, gt_range=(1500, 1500)) if rot else synth.hbb_batch(seed=1, gt_range=(1500, 1500))
hw = d["img_metas"][0]["img_shape"][0]
pb = torch.cat(d["pseudo_boxes"]).to(dev)
idx = torch.tensor([i for i in range(2) for _ in range(1500)], dtype=torch.int32, device=dev)
wh = torch.tensor([[float(hw), float(hw)]] * 2, device=dev)
cfg = synth.stress_ext_cfg(8)[0]
bags, _ = ops.bag_gen(ops.make_rois(pb, idx), wh, cfg["base_ratios"], cfg["shake_ratio"], cfg["min_scale"], rot)   # 192 000 RoIs
perm = torch.randperm(bags.shape[0], generator=torch.Generator().manual_seed(0)).to(dev)
bags = bags[perm].contiguous()
fx = d["feat"].to(dev)
feat = ops.nchw_to_nhwc(fx, torch.float16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []
for K in (148, 296, 1184, 2500, 5000, 10000, 20000, 48000, 96000):
    rois = bags[:K].contiguous()
    out = torch.empty((K, 12544), dtype=torch.bfloat16, device=dev)
    ts = []
    for i in range(20):
        flush.zero_()
        feat.copy_(ops.nchw_to_nhwc(fx, torch.float16))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200000)
        e0.record()
        ops.roi_align_forward(feat, rois, ops.OUT_BF16_BINMAJOR, 0.125, 2 if rot else 0, True, rotated=rot, clockwise=True, out=out)
        e1.record()
        torch.cuda.synchronize()
        if i >= 5:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    med = ts[len(ts) // 2]
    nbytes = K * 12544 * 2 + feat.numel() * 2 + rois.numel() * 4
    rows.append((K, med))
    print(f"K={K:6d}  {med:8.1f} us   {nbytes / med / 1e3:6.0f} GB/s   frac {nbytes / med / 1e3 / 6554.6:.3f}")
(k0, t0), (k1, t1) = rows[-3], rows[-1]
per = (t1 - t0) / (k1 - k0)
print(f"steady per-RoI {per * 1e3:.2f} ns  ->  fixed cost at K=5000: {dict(rows)[5000] - per * 5000:.1f} us of {dict(rows)[5000]:.1f} us")
