#!/usr/bin/env python
"""Eager OBB (config #3) phase-2 steps for ncu launch lists: python tools/prof_obb.py [steps]."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_teacher_b200 import synth
from point_teacher_b200.mil_head import RotatedMILHead
from point_teacher_b200.refine import phase2_refine
dev = torch.device("cuda")
d = synth.obb_batch(seed=0)
torch.manual_seed(0)
head = RotatedMILHead(num_classes=9, num_stages=1, top_k=3, precision="bf16").to(dev)
to = lambda l: [t.to(dev) for t in l]
x = d["feat"].to(dev)
args = (d["img_metas"], to(d["pseudo_boxes"]), to(d["pseudo_points"]), to(d["pseudo_labels"]), to(d["gt_boxes"]))
negs = [to(d["neg_boxes"][0])]
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
        phase2_refine(head, (x,), *args, synth.OBB_FINE_CFG, synth.OBB_EXT_CFG, num_stages=1, neg_boxes=negs)
torch.cuda.synchronize()
print("ok")
