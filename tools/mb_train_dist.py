#!/usr/bin/env python
"""Captured training step under torchrun for several values of the GEMM's SM reserve (SMs the persistent GEMMs leave
to NCCL while a gradient all-reduce is in flight).  Usage:
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/mb_train_dist.py 0 16 24 32"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_teacher_b200 import synth
from point_teacher_b200.mil_head import MILHead
from point_teacher_b200.train import CapturedTrainStep

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    sys.stdout.flush()
    fd = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier(); torch.cuda.synchronize()
    os.dup2(fd, 1); os.close(fd)
d = synth.hbb_batch(seed=rank)
torch.manual_seed(0)
head = MILHead(num_classes=8, num_stages=1, top_k=1, precision="bf16").to(dev)
to = lambda l: [t.to(dev) for t in l]
tin = dict(feat=d["feat"].to(dev), pseudo_boxes=to(d["pseudo_boxes"]), pseudo_points=to(d["pseudo_points"]),
           pseudo_labels=to(d["pseudo_labels"]), gt_boxes=to(d["gt_boxes"]), neg_boxes=[to(d["neg_boxes"][0])])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
keep = []
# arguments: <sm_reserve>[:<reduce scheme split|group>[:<early roi bwd 0|1>]] ...
for arg in sys.argv[1:] or ["0"]:
    parts = arg.split(":")
    r = int(parts[0])
    os.environ["PTB200_NCCL_SM_RESERVE"] = str(r)
    if len(parts) > 1:
        os.environ["PTB200_REDUCE_SCHEME"] = parts[1]
    if len(parts) > 2:
        os.environ["PTB200_EARLY_ROI_BWD"] = parts[2]
    c = CapturedTrainStep(head, tin, d["img_metas"], synth.HBB_FINE_CFG, synth.HBB_EXT_CFG, num_stages=1, cap=100)
    keep.append(c)
    for _ in range(3):
        c.replay()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = []
    for _ in range(40):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); c.replay(); b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("world %d  cfg %-14s  train step %.4f ms (max over ranks)" % (world, arg, t.item()), flush=True)
if world > 1:
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
