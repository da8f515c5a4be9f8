"""Multi-GPU plumbing of the phase-2 path: one process per GPU, images sharded across ranks,
NO data-path collective (bags, RoIs, scores and selection are independent per image; an image's RoIs only read that
image's feature map).  The only exchange steps are
  * the mean of the logged loss scalars across ranks, as mmdet's ``BaseDetector._parse_losses`` does
    (HBB_TOD/mmdet/models/detectors/base.py: ``dist.all_reduce(loss_value.div_(dist.get_world_size()))``), and
  * one flat-bucket all-reduce (average) of the MIL-head parameter gradients per step (DDP semantics of
    HBB_TOD/tools/train.py -> mmdet.apis.train_detector -> MMDistributedDataParallel), restricted to the parameters
    the path actually uses (``shared_fcs_reg``, ``shared_fcs_bag``, ``fc_cls``, ``fc_ins``, ``fc_reg``): the
    constructed-but-unused ``shared_fcs``, ``shared_fcs_refine`` and ``fc_iou`` never receive a gradient.
Per-rank loss denominators (``avg_factor = K_local``, ``num_sample_local``) are kept per process exactly like the
reference under DDP.  Backend: NCCL over NVLink on the GPUs, gloo on CPU (tests)."""
import torch
import torch.distributed as dist

USED_PREFIXES = ("shared_fcs_reg.", "shared_fcs_bag.", "fc_cls.", "fc_ins.", "fc_reg.")


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def image_shard(num_images, rank_=None, world_=None):
    """Contiguous, balanced partition of image indices: the first ``num_images % world`` ranks own one more."""
    r = rank() if rank_ is None else rank_
    w = world() if world_ is None else world_
    base, extra = divmod(num_images, w)
    start = r * base + min(r, extra)
    return list(range(start, start + base + (1 if r < extra else 0)))


def reduce_mean_losses(losses, group=None):
    """Mean over ranks of every logged scalar (one collective on a stacked vector); identity for one process."""
    if world() == 1:
        return dict(losses)
    keys = sorted(losses)
    vec = torch.stack([losses[k].detach().reshape(()).float() for k in keys])
    dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    vec /= world()
    return {k: vec[i] for i, k in enumerate(keys)}


class MILGradBucket:
    """One flat fp32 bucket over the gradients of the MIL-head parameters the path uses (27.8 M parameters = 111 MB
    for one stage); ``all_reduce_()`` averages it across ranks with a single collective and scatters the result
    back into ``param.grad``."""

    def __init__(self, head, prefixes=USED_PREFIXES):
        self.named = [(n, p) for n, p in head.named_parameters() if n.startswith(prefixes) and p.requires_grad]
        if not self.named:
            raise ValueError("no MIL-head parameters found")
        self.numel = sum(p.numel() for _, p in self.named)
        dev = self.named[0][1].device
        self.flat = torch.zeros((self.numel,), dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for _, p in self.named:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def names(self):
        return [n for n, _ in self.named]

    def pack_(self):
        for (_, p), v in zip(self.named, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        return self.flat

    def all_reduce_(self, group=None, async_op=False):
        self.pack_()
        if world() == 1:
            return self.unpack_()
        self.flat /= world()
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            return work
        return self.unpack_()

    def unpack_(self):
        for (_, p), v in zip(self.named, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        return self
