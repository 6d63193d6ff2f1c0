#!/usr/bin/env python
"""bench.py -- mask2image training throughput on B200 (BASELINE.json metric: train images/sec @512x1024).

    python bench.py --gpus N --steps K --warmup W            # this implementation (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle restatement), rank 0

One "step" = one full training iteration of train_mask2image.py:58-86 on a synthetic Cityscapes-shaped batch
(BASELINE config #2: 512x1024, 35 classes, --no_instance, GlobalGenerator 4 down / 9 res, 3-scale D, VGG19 feature
matching, 4 images per GPU): encode + G fwd + D fwd (fake, real) + VGG fwd x2 + both backward passes + grad
allreduce + Adam x2.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, LABEL_NC, PER_GPU_BATCH = 512, 1024, 35, 4
# conv MACs only, 2 FLOP/MAC, BASELINE.md section 3 (fwd 1000.0 + bwd 1312.0 GMAC per image)
TFLOP_PER_IMAGE = 4.624


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return d, "measured"
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0), "fallback"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def workload_desc(n_gpus):
    return dict(workload="BASELINE config #2/#3: mask2image 512x1024, 35-class synthetic labels, --no_instance, "
                         "GlobalGenerator(ngf64, 4 down, 9 res) + 3-scale MultiscaleDiscriminator + VGG19 feat-match, "
                         "full train step, %d images/GPU" % PER_GPU_BATCH,
                global_batch=PER_GPU_BATCH * n_gpus, per_gpu_batch=PER_GPU_BATCH, parallelism="dp%d" % n_gpus,
                l2="per-step working set (tens of GB of activations) >> 126 MB L2: no flush needed",
                vgg_weights="seeded random (no network for ImageNet weights)")


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU algorithm (oracle restatement of its PyTorch modules)
# ------------------------------------------------------------------------------------------------------------
def cpu_step_time(h, w, reps=1):
    from oracle import model as O
    from tests.util_weights import random_d_sd, random_g_sd
    opt = O.Opt(num_D=3)
    g_sd, d_sd = random_g_sd(LABEL_NC + 3, 3, 64, 4, 9), random_d_sd(LABEL_NC + 6, 64, 3, 3)
    vgg = O.vgg19_random_state_dict()
    batch = O.synthetic_batch(1, h, w, LABEL_NC, seed=1234)
    ts = []
    for _ in range(reps):
        t0 = time.time()
        O.train_step(opt, g_sd, d_sd, vgg, batch)
        ts.append(time.time() - t0)
    return min(ts)


def run_cpu(budget_s, steps, warmup):
    """Time `steps` CPU training steps (after `warmup`) on a sample of the workload that fits the budget."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t_probe = cpu_step_time(64, 128)  # 1/64 of the pixels
    per_step = budget_s / max(1, steps + warmup)
    size = (64, 128)
    for hh, ww in ((512, 1024), (256, 512), (128, 256)):
        est = t_probe * (hh * ww) / (64 * 128)
        if est <= per_step:
            size = (hh, ww)
            break
    from oracle import model as O
    from tests.util_weights import random_d_sd, random_g_sd
    opt = O.Opt(num_D=3)
    g_sd, d_sd = random_g_sd(LABEL_NC + 3, 3, 64, 4, 9), random_d_sd(LABEL_NC + 6, 64, 3, 3)
    vgg = O.vgg19_random_state_dict()
    batch = O.synthetic_batch(1, size[0], size[1], LABEL_NC, seed=1234)
    state = None
    for _ in range(warmup):
        _, _, _, _, state = O.train_step(opt, g_sd, d_sd, vgg, batch, state)
    t0 = time.time()
    for _ in range(steps):
        _, _, _, _, state = O.train_step(opt, g_sd, d_sd, vgg, batch, state)
    dt = (time.time() - t0) / max(1, steps)
    frac = size[0] * size[1] / float(H * W)
    ips = frac / dt  # 512x1024-image equivalents per second
    sample = "1 image of %dx%d per step (%.4g of a 512x1024 image; images/sec in 512x1024-equivalents), %d timed + %d " \
             "warm-up steps, torch CPU fp32 oracle of the reference modules" % (size[0], size[1], frac, steps, warmup)
    return ips, dt * 1e3, cores, sample


def main_torch_gpu(args, out_fd):
    """Opt-in extra baseline (`--impl torch_gpu`, never run by default): the oracle restatement of the reference's
    modules executed by stock PyTorch / cuDNN on the same B200 (fp32 tensors, torch's default TF32 convolutions), one
    full training step of config #2 per iteration -- the practical bar SURVEY section 8(d) asks to report, since the
    reference publishes no GPU numbers.  Baseline measurement only; the product never touches this path."""
    from oracle import model as O
    from tests.util_weights import random_d_sd, random_g_sd
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = os.environ.get("HM_CUDNN_BENCHMARK", "0") == "1"   # the reference turns it on
    opt = O.Opt(num_D=3)
    cu = lambda sd: {k: v.to(dev) for k, v in sd.items()}  # noqa: E731
    g_sd, d_sd = cu(random_g_sd(LABEL_NC + 3, 3, 64, 4, 9)), cu(random_d_sd(LABEL_NC + 6, 64, 3, 3))
    vgg = cu(O.vgg19_random_state_dict())
    batch = {k: v.to(dev) for k, v in O.synthetic_batch(PER_GPU_BATCH, H, W, LABEL_NC, seed=1234).items()}
    state = None
    for _ in range(max(1, args.warmup)):
        _, _, _, _, state = O.train_step(opt, g_sd, d_sd, vgg, batch, state)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        _, _, _, _, state = O.train_step(opt, g_sd, d_sd, vgg, batch, state)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    _emit(dict(impl="torch_gpu", metric="mask2image train images/sec @512x1024", value=PER_GPU_BATCH / (ms / 1e3),
               unit="images/sec", n_gpus=1, steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True,
               dtype="f32 tensors, cuDNN TF32 convolutions (torch defaults)", data="synthetic", config=workload_desc(1),
               note="stock PyTorch/cuDNN autograd + per-tensor Adam on the oracle restatement of the reference modules; "
                    "includes the host syncs of float(loss); cudnn.benchmark=%s" % torch.backends.cudnn.benchmark), out_fd)
    return 0


def main_reference(args, out_fd):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ips, ms, cores, sample = run_cpu(150.0, args.steps, args.warmup)
    line = dict(impl="reference", metric="mask2image train images/sec @512x1024", value=ips, unit="images/sec",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", config=workload_desc(args.gpus),
                cpu_baseline=dict(value=ips, unit="images/sec", cores=cores, kind="port", sample=sample),
                e2e=dict(value=ips, unit="images/sec", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    _emit(line, out_fd)
    return 0


# ------------------------------------------------------------------------------------------------------------
def k1_traffic(split):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE K1 launch from the committed `ncu --set full` capture
    (profiles/k1_traffic.json, written by tools/k1_traffic.py from the .ncu-rep); None if there is no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as fh:
            d = json.load(fh)["bf16x3" if split else "bf16"]
        return float(d["dram_bytes_read"] + d["dram_bytes_write"]), d.get("source")
    except Exception:
        return None, None


def time_k1(model, pk, split):
    """Dominant kernel: the residual-block 3x3 conv (1024->1024 @32x64 x B) = hm_kgemm_kernel<256>, timed alone
    with CUDA events on the launching stream (burst peak applies)."""
    from neurips18_hierchical_image_manipulation_b200 import ops
    ctx = model.ctx
    conv = [c for k, c in model.netG.stages if k == "resA"][0]
    B = PER_GPU_BATCH
    x = ops.Operand(ctx, B, 32, 64, 1024, border=1, zero=True)
    x.hi.normal_(0, 0.5)
    if x.lo is not None:
        x.lo.normal_(0, 0.002)
    y = torch.empty(B, 32, 64, 1024, device=ctx.device)
    for _ in range(3):
        conv.forward(x, 0, out32=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        conv.forward(x, 0, out32=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * B * 32 * 64 * 1024 * 1024 * 9          # algorithmic: 19.33 GMAC per image per conv
    executed = flops * (3 if split else 1)
    traffic, traffic_src = k1_traffic(split)
    return dict(bound="tensor", achieved=flops / ms / 1e9, peak=pk["bf16_tflops"], unit="TFLOP/s",
                frac=flops / ms / 1e9 / pk["bf16_tflops"], traffic=traffic, traffic_unit="bytes per launch (DRAM read+write)",
                traffic_source=traffic_src, kernel="hm_kgemm2_kernel<256> CTA-pair (res-block conv3x3 1024->1024, M=%d N=1024 K=9216)" % (B * 2048),
                ms_per_launch=ms, executed_tflops=executed / ms / 1e9, executed_frac=executed / ms / 1e9 / pk["bf16_tflops"],
                note="achieved counts ALGORITHMIC conv flops; in bf16x3 (fp32-parity) mode every product is issued 3x "
                     "(hi*hi + lo*hi + hi*lo), executed_* counts those tensor-core flops")


def run_mode(model, precision_name, batch_dev, batch_pinned, steps, warmup, world, rank, sampler=None):
    import torch.distributed as dist
    m = model
    kw_dev = dict(label=batch_dev["label"], inst=batch_dev["inst"], image=batch_dev["image"], feat=None,
                  mask_in=batch_dev["mask_in"], mask_out=batch_dev["mask_out"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=m.device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / n

    for _ in range(warmup):
        m.optimize_parameters(**kw_dev)
    barrier()
    if sampler:
        sampler.start()
    l0 = m.ctx.launches
    ms_dev = timed(lambda: m.optimize_parameters(**kw_dev), steps)
    launches = (m.ctx.launches - l0) // steps
    clocks = sampler.stop() if sampler else None
    m.ctx.check_pipeline()

    # end to end: host (pinned) inputs -> H2D inside the step, D2H read of the 5 losses every step
    host_losses = torch.empty(5, dtype=torch.float32, pin_memory=True)
    kw_host = dict(label=batch_pinned["label"], inst=batch_pinned["inst"], image=batch_pinned["image"], feat=None,
                   mask_in=batch_pinned["mask_in"], mask_out=batch_pinned["mask_out"])

    def e2e_step():
        ls = m.optimize_parameters(**kw_host)
        host_losses.copy_(ls, non_blocking=False)

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, steps)
    h2d = sum(batch_pinned[k].numel() * 4 for k in ("label", "image", "mask_in"))
    return dict(ms=ms_dev, ms_e2e=ms_e2e, launches=launches, clocks=clocks, h2d=h2d, d2h=20,
                losses=[float(x) for x in host_losses])


def _emit(line, fd):
    os.write(fd, (json.dumps(line) + "\n").encode())


def main():
    # stdout must carry exactly ONE JSON line: park the real stdout and send everything else (NCCL banners, library
    # chatter) to stderr
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200")
    ap.add_argument("--precision", type=str, default="bf16x3", help="primary precision mode (bf16x3 = fp32 parity)")
    ap.add_argument("--no-alt", action="store_true", help="skip the secondary (plain bf16) measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args, out_fd)
    if args.impl == "torch_gpu":
        return main_torch_gpu(args, out_fd)
    if args.warmup < 3:
        args.warmup = 3

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from neurips18_hierchical_image_manipulation_b200.models import Options, create_model
    from neurips18_hierchical_image_manipulation_b200.synthetic import synthetic_batch

    pk, pk_kind = peaks()
    batch = synthetic_batch(PER_GPU_BATCH, H, W, LABEL_NC, seed=1234 + rank)
    batch_pinned = {k: v.pin_memory() for k, v in batch.items()}
    dev = torch.device("cuda", local)
    batch_dev = {k: v.to(dev) for k, v in batch.items()}

    results = {}
    # the alternate precision modes are a single-GPU side measurement; multi-GPU runs time the primary mode only
    skip_alt = args.no_alt or world > 1
    modes = [args.precision] + ([] if skip_alt else [p for p in ("bf16x3", "mixed", "bf16") if p != args.precision])
    roof = None
    for i, prec in enumerate(modes):
        opt = Options(label_nc=LABEL_NC, no_instance=True, netG="global", ngf=64, n_downsample_global=4, n_blocks_global=9,
                      num_D=3, n_layers_D=3, ndf=64, gpu_ids=[local], precision=prec, name="bench")
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            model = create_model(opt).module
        sampler = ClockSampler(local) if (rank == 0 and i == 0) else None
        results[prec] = run_mode(model, prec, batch_dev, batch_pinned, args.steps, args.warmup, world, rank, sampler)
        if rank == 0:
            r = time_k1(model, pk, prec != "bf16")
            results[prec]["k1"] = r
            if i == 0:
                roof = r
        del model
        torch.cuda.empty_cache()

    if rank == 0:
        prim = results[args.precision]
        gb = PER_GPU_BATCH * world
        value = gb / (prim["ms"] / 1e3)
        cpu = None
        if not args.no_cpu and world == 1:   # the CPU baseline leg runs on rank 0 at N=1 only
            ips, ms, cores, sample = run_cpu(25.0, 1, 0)
            cpu = dict(value=ips, unit="images/sec", cores=cores, kind="port", sample=sample)
        roof["peak_source"] = "%s (MEASURED_PEAKS.json bf16_tflops, burst: kernel timed alone)" % pk_kind
        # whole-step tensor roofline against the sustained peak
        step_flops = TFLOP_PER_IMAGE * PER_GPU_BATCH
        line = dict(metric="mask2image train images/sec @512x1024", value=value, unit="images/sec", n_gpus=world,
                    steps=args.steps, warmup=args.warmup, ms_per_step=prim["ms"], higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="bf16x3 (bf16 hi/lo split operands, 3 tcgen05 products, fp32 accumulate; "
                    "fp32-parity mode)" if args.precision == "bf16x3" else "bf16 (fp32 accumulate)", data="synthetic",
                    config=workload_desc(world), clocks=prim["clocks"],
                    e2e=dict(value=gb / (prim["ms_e2e"] / 1e3), unit="images/sec", h2d_bytes_per_step=prim["h2d"],
                             d2h_bytes_per_step=prim["d2h"], ms_per_step=prim["ms_e2e"]),
                    gpu_launches=prim["launches"], roofline=roof, cpu_baseline=cpu,
                    step_tensor_roofline=dict(algorithmic_tflop_per_step=step_flops,
                                              achieved_tflops=step_flops / (prim["ms"] / 1e3),
                                              peak_sustained=pk.get("bf16_tflops_sustained"),
                                              frac=step_flops / (prim["ms"] / 1e3) / pk.get("bf16_tflops_sustained", 1400.0)),
                    losses_last_step=prim["losses"])
        for prec in modes[1:]:
            r = results[prec]
            line["alt_precision_" + prec] = dict(value=gb / (r["ms"] / 1e3), unit="images/sec", ms_per_step=r["ms"],
                                                 e2e=gb / (r["ms_e2e"] / 1e3), gpu_launches=r["launches"],
                                                 k1_tflops=r["k1"]["achieved"], k1_frac=r["k1"]["frac"],
                                                 note=("forward bf16x3 (outputs / losses within the fp32 tolerance), gradient GEMMs "
                                                       "single bf16 products" if prec == "mixed" else
                                                       "plain bf16 products: NOT within the 1e-3 fp32 tolerance "
                                                       "(generator output ~1e-2 rel); reported for reference"))
        sys.stdout.flush()
        _emit(line, out_fd)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
