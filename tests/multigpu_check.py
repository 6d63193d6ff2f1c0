"""Multi-rank correctness of the PRODUCT path (not the oracle): run under torchrun on >= 2 GPUs of one box,

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py [--out gpurun_out/multigpu_check.json]

Every rank holds (a) a data-parallel replica that trains on ITS shard of each batch (B images per rank, gradients
allreduced inside the CUDA-graphed fused step) and (b) a stand-alone model (opt.data_parallel=False) that trains on the
WHOLE batch (world x B images).  Because InstanceNorm is per sample and every loss is a batch mean, (a) must equal (b)
(train_mask2image.py:68 averages the replica losses):

  * step 0 (identical weights): mean over ranks of the shard losses == full-batch losses to 1e-5;
  * later steps: still equal to 5e-3 (fp32 atomics order in the weight gradients differs, Adam's first steps turn
    sign flips of ~zero gradients into +-lr moves);
  * after K graph-replayed steps all replicas hold BIT-IDENTICAL weights (integer checksum, MAX - MIN == 0), and the
    replica weights agree with the stand-alone model's to the same bound the graph-vs-eager test uses.

tests/test_multigpu_gpu.py launches this script when the box has >= 2 GPUs.
"""
import argparse
import contextlib
import io
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def box2mask_check(rank, world, local):
    """box2mask (TwoStreamAE_mask with --use_gan) under data parallelism, the way nn.DataParallel trains it
    (models/models.py:21-22): every replica normalises with the BatchNorm statistics of ITS shard, the shard gradients are
    averaged.  (a) data-parallel replicas on their shards vs (b) a stand-alone model that runs forward + backward on every
    shard in turn (accumulating gradients) and steps with their mean: losses at step 0 to 1e-5, weights after that step,
    then bit-identical replicas after graph-replayed iterations."""
    from neurips18_hierchical_image_manipulation_b200 import parallel
    from neurips18_hierchical_image_manipulation_b200.models import Options, create_model
    from neurips18_hierchical_image_manipulation_b200.synthetic import box2mask_batch

    def mk(dp):
        opt = Options(model="AE_maskgen_twostream", isTrain=False, gpu_ids=[local], precision="bf16x3", name="mgpu_b2m",
                      label_nc=6, output_nc=6, conv_dim=32, n_blocks=2, num_layers=3, conv_size=4, which_stream="obj_context",
                      cond_in="ctx_obj", use_output_gate=True, num_resnetblocks=1, norm_layer="batch", lr=2e-4, beta1=0.5,
                      beta2=0.999, use_gan=True, which_gan="patch_multiscale", gan_weight=0.1, num_layers_D=3, ndf=16,
                      use_ganFeat_loss=True, lambda_feat=1.0, checkpoints_dir="/tmp/hm_mgpu", data_parallel=dp)
        with contextlib.redirect_stdout(io.StringIO()):
            return create_model(opt)
    a, b = mk(True), mk(False)
    assert torch.equal(a.fpG.flat, b.fpG.flat) and torch.equal(a.fpD.flat, b.fpD.flat)
    B = 2
    full = box2mask_batch(world * B, 64, 6, seed=300)
    shards = [{k: v[r * B:(r + 1) * B] for k, v in full.items()} for r in range(world)]
    args = lambda d: (d["label_map"], None, d["mask_ctx_in"], None, d["mask_out"], d["mask_obj_inst"], d["cls"], d["mask_in"])  # noqa: E731
    # (b) stand-alone emulation of one data-parallel iteration
    b.optimizer.zero_grad(); b.optimizer_D.zero_grad()
    lb = torch.zeros(5, device=b.device)
    for sh in shards:
        ls, _ = b.forward(*args(sh), train=False)
        lb += torch.stack([v.float() for v in ls])
        b.backward_losses()
        b.netD.backward(b._last["d_real"], 1.0, 0.5, True)
        b.netD.backward(b._last["d_fake"], 0.0, 0.5, True)
    lb /= world
    b.optimizer.step(grad_scale=1.0 / world)
    b.optimizer_D.step(grad_scale=1.0 / world)
    # (a) the data-parallel iteration on this rank's shard
    ls, _ = a.forward(*args(shards[rank]))
    la = torch.stack([ls[0], ls[1], ls[3], ls[4], ls[5]]).float()
    dist.all_reduce(la)
    la /= world
    torch.cuda.synchronize()
    out = dict(loss_err=float(((la - lb).abs() / lb.abs().clamp_min(1e-12)).max()))
    assert out["loss_err"] < 1e-5, (la.tolist(), lb.tolist())
    dG, dD = (a.fpG.flat - b.fpG.flat).abs(), (a.fpD.flat - b.fpD.flat).abs()
    out.update(step1_G_mean=float(dG.mean()), step1_D_mean=float(dD.mean()), step1_G_max=float(dG.max()))
    assert out["step1_G_mean"] < 0.02 * 2e-4 and out["step1_D_mean"] < 0.02 * 2e-4, out      # Adam's first step is +-lr
    for i in range(4):
        sh = box2mask_batch(B, 64, 6, seed=400 + 10 * i + rank)
        a.forward(*args(sh))
    torch.cuda.synchronize()
    a.ctx.check_pipeline(); b.ctx.check_pipeline()
    out["graph"] = isinstance(a._graph, dict)
    assert out["graph"], "the data-parallel box2mask iteration was not captured into a CUDA graph"
    ident = True
    for flat in (a.fpG.flat, a.fpD.flat):
        chk = flat.view(torch.int32).to(torch.int64).sum().reshape(1)
        hi, lo = chk.clone(), chk.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        ident = ident and bool((hi - lo).item() == 0)
    out["replicas_identical"] = ident
    assert ident, "box2mask replicas diverged"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--per-rank", type=int, default=2)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    from neurips18_hierchical_image_manipulation_b200 import parallel
    parallel.configure_nccl_env()
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from neurips18_hierchical_image_manipulation_b200.models import Options, create_model
    from neurips18_hierchical_image_manipulation_b200.synthetic import synthetic_batch

    def mk(dp):
        opt = Options(label_nc=9, ngf=16, n_downsample_global=2, n_blocks_global=2, ndf=16, num_D=2, no_instance=False,
                      use_output_gate=True, gpu_ids=[local], precision="bf16x3", name="mgpu", vgg_weights="random",
                      checkpoints_dir="/tmp/hm_mgpu", data_parallel=dp)
        with contextlib.redirect_stdout(io.StringIO()):
            return create_model(opt).module
    a, b = mk(True), mk(False)
    assert torch.equal(a.flat, b.flat)          # same init seed on every rank and model
    B = args.per_rank
    res = dict(world=world, per_rank=B, steps=args.steps, loss_err=[], graph=None)
    for i in range(args.steps):
        full = synthetic_batch(world * B, 96, 128, 9, seed=100 + i)       # the same full batch on every rank
        shard = {k: v[rank * B:(rank + 1) * B] for k, v in full.items()}
        kw = lambda d: dict(label=d["label"], inst=d["inst"], image=d["image"], feat=None, mask_in=d["mask_in"],  # noqa: E731
                            mask_out=d["mask_out"])
        la = parallel.allreduce_losses_(a.optimize_parameters(**kw(shard)).clone())
        lb = b.optimize_parameters(**kw(full)).clone()
        torch.cuda.synchronize()
        err = float(((la - lb).abs() / lb.abs().clamp_min(1e-12)).max())
        res["loss_err"].append(err)
        assert err < (1e-5 if i == 0 else 5e-3), (i, err, la.tolist(), lb.tolist())
    a.ctx.check_pipeline(); b.ctx.check_pipeline()
    res["graph"] = isinstance(a._graph, dict)
    assert res["graph"], "the data-parallel fused step (with its allreduce) was not captured into a CUDA graph"
    chk = a.flat.view(torch.int32).to(torch.int64).sum().reshape(1)
    hi, lo = chk.clone(), chk.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    res["replicas_identical"] = bool((hi - lo).item() == 0)
    assert res["replicas_identical"], "replicas diverged"
    d = (a.flat - b.flat).abs()
    res["dp_vs_single_max"], res["dp_vs_single_mean"] = float(d.max()), float(d.mean())
    assert res["dp_vs_single_max"] < 2.5e-3 and res["dp_vs_single_mean"] < 2e-5, res
    res["box2mask"] = box2mask_check(rank, world, local)
    dist.barrier()
    if rank == 0:
        print("MULTIGPU_CHECK " + json.dumps(res))
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as fh:
                json.dump(res, fh)
    # tearing the communicator down while CUDA graphs that captured its kernels are alive can block: drop the models
    # first and leave without the (optional) destroy
    del a, b
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
