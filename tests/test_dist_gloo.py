"""N > 1 host logic on CPU: world_size-2 gloo processes: image sharding, the loss-scalar mean and
the flat-bucket all-reduce of the MIL-head gradients (which parameters are in the bucket, averaging, scatter-back)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from point_teacher_b200 import dist as pd
        from point_teacher_b200.mil_head import MILHead
        res = {}
        res["shard5"] = pd.image_shard(5)
        res["shard4"] = pd.image_shard(4)
        losses = {"stage0_loss_mil_bags": torch.tensor(1.0 + rank), "stage0_loss_mil_bbox": torch.tensor(0.5 * (rank + 1)),
                  "coarse_bboxes_iou": torch.tensor(0.25)}
        red = pd.reduce_mean_losses(losses)
        res["losses"] = {k: float(v) for k, v in red.items()}
        torch.manual_seed(0)                          # identical replicas, like DDP after broadcast
        head = MILHead(num_classes=8, num_stages=1, top_k=1, in_channels=8)
        bucket = pd.MILGradBucket(head)
        res["names"] = bucket.names()
        res["numel"] = bucket.numel
        g = torch.Generator().manual_seed(100 + rank)
        for n, p in head.named_parameters():
            if n.startswith(pd.USED_PREFIXES):
                p.grad = torch.randn(p.shape, generator=g)
        mine = {n: p.grad.clone() for n, p in bucket.named}
        bucket.all_reduce_()
        gathered = [None] * world
        dist.all_gather_object(gathered, {n: v for n, v in mine.items()})
        err = 0.0
        for n, p in bucket.named:
            mean = sum(gathered[r][n] for r in range(world)) / world
            err = max(err, float((p.grad - mean).abs().max()))
        res["grad_err"] = err
        res["untouched"] = all(p.grad is None for n, p in head.named_parameters() if not n.startswith(pd.USED_PREFIXES))
        # direct mode: the backward kernels write into the flat buffer and each (stage, branch) group is reduced on its
        # own, asynchronously, the moment its branch is done (the multi-rank 'group' scheme; the un-permute of
        # finish_() is a CUDA kernel and is covered by tools/check_dist_grads.py on GPUs)
        b2 = pd.MILGradBucket(head)
        b2.flat.copy_(torch.randn(b2.flat.shape, generator=torch.Generator().manual_seed(7 + rank)))
        local = b2.flat.clone()
        b2.reduce_group_(0, "reg")
        b2.reduce_group_(0, "bag")
        for w in b2._works:
            w.wait()
        both = [None] * world
        dist.all_gather_object(both, local)
        res["direct_err"] = float((b2.flat - sum(both) / world).abs().max())
        # layout invariants the kernels rely on: groups tile the buffer, 16-byte aligned weight slots, the FC1 weight
        # (reduced / un-permuted separately on one rank) is the last entry of its group
        gs = b2.groups
        res["layout_ok"] = (gs[0]["lo"] == 0 and all(a["hi"] == b["lo"] for a, b in zip(gs, gs[1:])) and
                            gs[-1]["hi"] == b2.flat.numel() and
                            all(b2.offsets[n] % 4 == 0 for n in b2.names() if n.endswith((".0.weight", ".1.weight"))) and
                            all(g["names"][-1].endswith(".0.weight") and
                                b2.offsets[g["names"][-1]] + dict(b2.named)[g["names"][-1]].numel() <= g["hi"] and
                                g["lo"] < g["small_hi"] <= b2.offsets[g["names"][-1]] for g in gs))
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert r0["shard5"] == [0, 1, 2] and r1["shard5"] == [3, 4]
    assert r0["shard4"] == [0, 1] and r1["shard4"] == [2, 3]
    for r in (r0, r1):
        assert abs(r["losses"]["stage0_loss_mil_bags"] - 1.5) < 1e-6
        assert abs(r["losses"]["stage0_loss_mil_bbox"] - 0.75) < 1e-6
        assert abs(r["losses"]["coarse_bboxes_iou"] - 0.25) < 1e-6
        assert r["grad_err"] < 1e-6 and r["untouched"]
        assert r["direct_err"] < 1e-6 and r["layout_ok"]
        assert all(n.split(".")[0] in ("shared_fcs_reg", "shared_fcs_bag", "fc_cls", "fc_ins", "fc_reg") for n in r["names"])
        # per stage: 2 x (Linear(8*49 -> 1024) + Linear(1024 -> 1024)) + fc_cls + fc_ins + fc_reg
        assert r["numel"] == 2 * (8 * 49 * 1024 + 1024 + 1024 * 1024 + 1024) + 2 * (1024 * 8 + 8) + 1024 * 4 + 4


def test_single_process_is_identity():
    from point_teacher_b200 import dist as pd
    assert pd.world() == 1 and pd.rank() == 0
    assert pd.image_shard(3) == [0, 1, 2]
    assert pd.image_shard(7, 2, 3) == [5, 6]
    out = pd.reduce_mean_losses({"a": torch.tensor(2.0)})
    assert float(out["a"]) == 2.0
