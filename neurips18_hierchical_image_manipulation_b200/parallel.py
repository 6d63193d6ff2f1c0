"""Data parallelism for the mask2image hot path: one process per GPU (torchrun), replicated weights, per-rank
minibatch shards, ONE sum-allreduce of the flat [G | D] gradient buffer per step (NCCL over NVLink on GPUs; any
torch.distributed backend works -- the CPU tests use gloo).  InstanceNorm statistics are per sample and every loss
is a mean over equally sized shards, so mean-of-shard-gradients == full-batch gradient:
this replaces the reference's single-process nn.DataParallel (models/models.py:21-22)."""
import torch
import torch.distributed as dist


def configure_nccl_env():
    """Call BEFORE init_process_group.  The fused step overlaps its bucketed gradient allreduce with the backward pass
    and reserves HM_COMM_SMS (default 8) SMs for it (hm_set_sm_limit); NCCL is told to use at most that many CTAs so it
    does not take SMs the persistent engines count on.  An explicit NCCL_MAX_CTAS in the environment wins."""
    import os
    os.environ.setdefault("NCCL_MAX_CTAS", os.environ.get("HM_COMM_SMS", "8"))


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_sum_(flat, group=None):
    """In-place sum-allreduce of a flat gradient buffer; returns the scale (1/world_size) that turns the sum into
    the full-batch mean gradient (folded into the fused Adam kernel's grad_scale)."""
    _, n = world()
    if n > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / n


def allreduce_sum_async_(flat, group=None):
    """Start an in-place sum-allreduce of `flat` on the communicator's own stream and return a handle whose .wait()
    makes the CURRENT stream wait for it (None when there is a single rank).  The fused step starts the generator
    segment right after the generator's backward pass so that it overlaps the discriminator's backward pass."""
    _, n = world()
    if n > 1:
        return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
    return None


def allreduce_losses_(losses, group=None):
    """Mean of the per-rank loss scalars (what train_mask2image.py:68 computes over DataParallel replicas)."""
    _, n = world()
    if n > 1:
        dist.all_reduce(losses, op=dist.ReduceOp.SUM, group=group)
        losses.div_(n)
    return losses


def shard_seed(base_seed, rank=None):
    r = world()[0] if rank is None else rank
    return base_seed + r
