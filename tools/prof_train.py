#!/usr/bin/env python
"""Two eager training steps of the bench workload (for ncu launch lists)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_teacher_b200 import synth
from point_teacher_b200.mil_head import MILHead
from point_teacher_b200.train import Phase2Trainer
dev = torch.device("cuda")
d = synth.hbb_batch(seed=0)
torch.manual_seed(0)
head = MILHead(num_classes=8, num_stages=1, top_k=1, precision="bf16").to(dev)
to = lambda l: [t.to(dev) for t in l]
x = d["feat"].to(dev).requires_grad_(True)
tr = Phase2Trainer(head, synth.HBB_FINE_CFG, synth.HBB_EXT_CFG, num_stages=1, cap=100)
args = (d["img_metas"], to(d["pseudo_boxes"]), to(d["pseudo_points"]), to(d["pseudo_labels"]), to(d["gt_boxes"]))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    x.grad = None
    tr.step((x,), *args, neg_boxes=[to(d["neg_boxes"][0])], reduce_logs=False)
torch.cuda.synchronize()
print("ok")
