"""The reference's training loop (train_mask2image.py:43-131) on synthetic Cityscapes-shaped batches, driving this
implementation through the reference's own call sequence -- the only changed line is the import of create_model.

    python examples/train_mask2image_synthetic.py                      # reference sequence: two backward / step pairs
    python examples/train_mask2image_synthetic.py --fused              # optimize_parameters(): fused, CUDA-graphed
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_mask2image_synthetic.py --fused

Uses the flag set of scripts/train_mask2image_city.sh by default (two-stream generator, 256x256 crops, batch 8).
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from neurips18_hierchical_image_manipulation_b200.models import Options, create_model   # was: from models.models import create_model
from neurips18_hierchical_image_manipulation_b200.synthetic import synthetic_batch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batchSize", type=int, default=8)
    ap.add_argument("--fineSize", type=int, default=256)
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--fused", action="store_true")
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from neurips18_hierchical_image_manipulation_b200 import parallel
        parallel.configure_nccl_env()
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    opt = Options(vgg_weights="random", name="synthetic_city", model="pix2pixHD_condImg", label_nc=35, output_nc=3, no_instance=True,
                  netG="global_twostream", which_encoder="ctx_label", use_skip=True, use_output_gate=True, no_imgCond=True,
                  mask_gan_input=True, n_downsample_global=4, n_layers_D=3, batchSize=args.batchSize, gpu_ids=[local],
                  precision=args.precision, checkpoints_dir="./checkpoints")
    model = create_model(opt)                                               # train_mask2image.py:39
    s = args.fineSize
    t0 = time.time()
    for i in range(args.iters):
        data = synthetic_batch(args.batchSize, s, s, opt.label_nc, seed=1234 + i + 1000 * local)
        if args.fused:
            losses = model.module.optimize_parameters(label=data["label"], inst=data["inst"], image=data["image"], feat=None,
                                                      mask_in=data["mask_in"], mask_out=data["mask_out"])
            loss_dict = dict(zip(model.module.loss_names, losses))
        else:
            losses, generated = model(label=data["label"], inst=data["inst"], image=data["image"], feat=None,  # :58-65
                                      mask_in=data["mask_in"], mask_out=data["mask_out"], infer=False)
            losses = [torch.mean(x) for x in losses]                                                              # :68
            loss_dict = dict(zip(model.module.loss_names, losses))
            loss_D = (loss_dict["D_fake"] + loss_dict["D_real"]) * 0.5                                            # :72
            loss_G = loss_dict["G_GAN"] + loss_dict["G_GAN_Feat"] + loss_dict["G_VGG"]                            # :73
            model.module.optimizer_G.zero_grad(); loss_G.backward(); model.module.optimizer_G.step()             # :78-80
            model.module.optimizer_D.zero_grad(); loss_D.backward(); model.module.optimizer_D.step()             # :83-86
        if i % 5 == 0 and local == 0:
            print("iter %d  " % i + "  ".join("%s %.4f" % (k, float(v)) for k, v in loss_dict.items()))         # :92-96
    torch.cuda.synchronize()
    if local == 0:
        print("%.1f images/sec (host-driven loop incl. synthetic data generation)" %
              (args.iters * args.batchSize / (time.time() - t0)))
        model.module.save("latest")                                                                               # :107


if __name__ == "__main__":
    main()
