#!/usr/bin/env python
"""RoIAlign backward micro-benchmark: the bench workload's geometry (2 images, 100x100 stride-8 map, 256 channels,
K tiny-object RoIs).  PTB200_RA_BWD_MMA=0 selects the register kernel, default the tensor-core one; the result is
checked against torchvision's autograd on the first 400 RoIs."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_teacher_b200 import ops, synth

dev = torch.device("cuda")
K = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
g = torch.Generator().manual_seed(0)
boxes = synth.make_boxes(g, K, (800, 800), median=14, hi=120)
rois = torch.cat([torch.randint(0, 2, (K, 1), generator=g).float(), boxes], 1).to(dev)
dA = (torch.randn(K, 49 * 256, generator=g) * 1e-3).to(torch.bfloat16).to(dev)
shape = (2, 100, 100, 256)

import torchvision
n = min(K, 400)
x = torch.zeros(2, 256, 100, 100, device=dev, requires_grad=True)
torchvision.ops.roi_align(x, rois[:n], (7, 7), 0.125, 0, True).backward(
    dA[:n].float().view(n, 7, 7, 256).permute(0, 3, 1, 2))
got = ops.nhwc_to_nchw_f32(ops.roi_align_backward(dA[:n], rois[:n], shape, 0.125))
rel = ((got - x.grad).abs().max() / x.grad.abs().max()).item()
print("max rel err vs torchvision autograd (400 RoIs): %.3e" % rel)

dfeat = torch.zeros(shape, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(25):
    dfeat.zero_()
    flush.fill_(i & 1)
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    ops.roi_align_backward(dA, rois, shape, 0.125, dfeat=dfeat)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
ts = sorted(ts[5:])
print("K=%d path=%s: median %.1f us  min %.1f us  checksum %.6e" %
      (K, "regs" if os.environ.get("PTB200_RA_BWD_MMA", "1")[0] == "0" else "mma", ts[len(ts) // 2], ts[0],
       dfeat.double().sum().item()))
