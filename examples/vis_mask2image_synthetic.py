"""The reference's test loop (vis_mask2image.py:14-45) on synthetic Cityscapes-shaped samples, driving this
implementation through the reference's own call sequence -- the only changed lines are the import of create_model, the
synthetic data source and writing the visuals as .npy arrays instead of an HTML page (the visualiser is outside the
hot path).

    python examples/train_mask2image_synthetic.py --iters 4      # writes ./checkpoints/synthetic_city/latest_net_G.pth
    python examples/vis_mask2image_synthetic.py --how_many 3     # loads it and synthesises

Like the reference: batchSize 1, isTrain=False (create_model returns the bare model, models/models.py:21), the generator
checkpoint must exist ("Generator must exist!"), inference() takes label / inst / image / mask_in / mask_out.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from neurips18_hierchical_image_manipulation_b200.models import Options, create_model   # was: from models.models import create_model
from neurips18_hierchical_image_manipulation_b200.synthetic import synthetic_batch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--how_many", type=int, default=3)
    ap.add_argument("--fineSize", type=int, default=256)
    ap.add_argument("--which_epoch", default="latest")
    ap.add_argument("--results_dir", default="./results")
    ap.add_argument("--name", default="synthetic_city")
    ap.add_argument("--checkpoints_dir", default="./checkpoints")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    opt = Options(name=args.name, model="pix2pixHD_condImg", label_nc=35, output_nc=3, no_instance=True,
                  netG="global_twostream", which_encoder="ctx_label", use_skip=True, use_output_gate=True, no_imgCond=True,
                  mask_gan_input=True, n_downsample_global=4, batchSize=1, gpu_ids=[0], isTrain=False,
                  which_epoch=args.which_epoch, checkpoints_dir=args.checkpoints_dir)
    model = create_model(opt)                                                                   # vis_mask2image.py:22
    web_dir = os.path.join(args.results_dir, opt.name, "test_%s" % opt.which_epoch)             # :25
    os.makedirs(web_dir, exist_ok=True)
    for i in range(args.how_many):                                                              # :28-30
        data = synthetic_batch(1, args.fineSize, args.fineSize, opt.label_nc, seed=4321 + i)
        generated = model.inference(label=data["label"], inst=data["inst"], image=data["image"],   # :32-38
                                    mask_in=data["mask_in"], mask_out=data["mask_out"])
        visuals = model.get_current_visuals()                                                   # :40
        print("process image... %s" % ("%05d" % i))                                             # :42
        for label, im in visuals.items():                                                       # :43 (save_images)
            np.save(os.path.join(web_dir, "%05d_%s.npy" % (i, label)), im)
        assert generated.shape == (1, 3, args.fineSize, args.fineSize)


if __name__ == "__main__":
    main()
