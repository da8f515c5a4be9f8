#!/usr/bin/env python
"""Microbenchmark of the G x A IoU / NWD matrix kernels against a plain fill of the same 60 MB (device time)."""
import torch, sys
sys.path.insert(0, '/root/repo')
from point_teacher_b200 import ops, synth
dev = torch.device('cuda')
g = torch.Generator().manual_seed(0)
gts = synth.make_boxes(g, 1500, (800, 800)).to(dev)
an = synth.make_boxes(g, 10000, (800, 800)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, fl, n=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        if fl: flush.zero_()
        torch.cuda._sleep(400000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts)//2] * 1e3
out = torch.empty((1500, 10000), device=dev)
for fl in (True, False):
    print('flush', fl, 'iou', t(lambda: ops.bbox_overlaps(gts, an, 'iou'), fl), 'wd', t(lambda: ops.bbox_metric(gts, an, 'wd', calc=1), fl),
          'fill', t(lambda: out.fill_(1.0), fl), 'empty-only', t(lambda: torch.empty((1500, 10000), device=dev), fl))
