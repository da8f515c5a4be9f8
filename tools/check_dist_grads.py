#!/usr/bin/env python
"""N-rank check of the in-place bucket reduction (``MILGradBucket.reduce_group_`` / ``finish_``): every rank runs the
manual training step on ITS OWN images, once with the overlapped NCCL reductions and once with them switched off
(local gradients); the reduced gradients must equal the plain all-reduce mean of the local ones on every rank.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dist_grads.py"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_teacher_b200 import dist as pdist, synth
from point_teacher_b200.mil_head import MILHead
from point_teacher_b200.train import Phase2Trainer

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
fd = os.dup(1); os.dup2(2, 1)
dist.init_process_group("nccl", device_id=dev)
dist.barrier(); torch.cuda.synchronize()
os.dup2(fd, 1); os.close(fd)
d = synth.hbb_batch(seed=rank, batch=2, img_hw=(256, 256), gt_range=(6, 10), n_neg=20)
torch.manual_seed(0)
head = MILHead(num_classes=8, num_stages=1, top_k=1, precision="bf16").to(dev)
to = lambda l: [t.to(dev) for t in l]
args = (d["img_metas"], to(d["pseudo_boxes"]), to(d["pseudo_points"]), to(d["pseudo_labels"]), to(d["gt_boxes"]))


def run(reduce):
    real = pdist.world
    if not reduce:
        pdist.world = lambda: 1
    try:
        tr = Phase2Trainer(head, synth.HBB_FINE_CFG, synth.HBB_EXT_CFG, num_stages=1, cap=100)
        x = d["feat"].to(dev).requires_grad_(True)
        tr.step((x,), *args, neg_boxes=[to(d["neg_boxes"][0])], reduce_logs=False)
        torch.cuda.synchronize()
        return {n: p.grad.detach().clone() for n, p in tr.bucket.named}
    finally:
        pdist.world = real


red = run(True)
loc = run(False)
worst = 0.0
for n, g in loc.items():
    m = g.clone()
    dist.all_reduce(m)
    m /= world
    if m.abs().max() < 1e-8:          # fc_ins.bias: mathematically zero (softmax over the bag is shift-invariant)
        assert (red[n] - m).abs().max() < 1e-8, n
        continue
    err = ((red[n] - m).abs().max() / m.abs().max()).item()
    same = torch.tensor([red[n].double().sum().item()], device=dev)
    lo, hi = same.clone(), same.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert lo.item() == hi.item(), (n, "ranks disagree after the reduction")
    worst = max(worst, err)
    assert err < 1e-5, (n, err)
if rank == 0:
    print("world %d: reduced gradients == all-reduce mean of the local gradients on every rank "
          "(worst relative difference %.2e over %d tensors)" % (world, worst, len(loc)), flush=True)
torch.cuda.synchronize()
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
