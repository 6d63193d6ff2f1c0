"""The reference's box2mask loop (train_box2mask.py:45-146, vis_box2mask.py:36-60) on synthetic box / context / instance
masks, driving this implementation through the reference's own call sequence with the flag set of
scripts/train_box2mask_city.sh (--use_gan --which_gan patch_multiscale --gan_weight 0.1 --use_ganFeat_loss --no_comb ...).
The only changed line is the import of create_model.

    python examples/train_box2mask_synthetic.py [--iters 20] [--batchSize 16] [--comb]     (--comb: omit --no_comb)
    python examples/train_box2mask_synthetic.py --ade      (scripts/train_box2mask_ade.sh: 49 classes, --norm_layer instance,
                                                            --add_dilated_layers, --lr_control, batch 8)
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from neurips18_hierchical_image_manipulation_b200.models import Options, create_model   # was: from models.models import create_model
from neurips18_hierchical_image_manipulation_b200.synthetic import box2mask_batch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batchSize", type=int, default=16)
    ap.add_argument("--fineSize", type=int, default=256)
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--comb", action="store_true")
    ap.add_argument("--ade", action="store_true")
    args = ap.parse_args()
    extra = dict(label_nc=35, output_nc=35)
    if args.ade:
        args.batchSize = min(args.batchSize, 8)
        extra = dict(label_nc=49, output_nc=49, norm_layer="instance", add_dilated_layers=True, lr_control=True)
    torch.cuda.set_device(0)
    opt = Options(model="AE_maskgen_twostream", name="synthetic_box2mask", isTrain=True, gpu_ids=[0],
                  use_gan=True, which_gan="patch_multiscale", gan_weight=0.1, num_layers_D=3, ndf=64, use_ganFeat_loss=True,
                  lambda_feat=1.0, which_stream="obj_context", cond_in="ctx_obj", conv_dim=64, conv_size=4, num_layers=3,
                  num_resnetblocks=1, n_blocks=6, use_output_gate=True, no_comb=not args.comb,
                  objReconLoss="bce", beta1=0.5, beta2=0.999, lr=0.0002, niter=400, niter_decay=100,
                  batchSize=args.batchSize, precision=args.precision, checkpoints_dir="./checkpoints",
                  **dict(dict(norm_layer="batch"), **extra))
    model = create_model(opt)                                                       # train_box2mask.py:45
    t0 = time.time()
    for i in range(args.iters):
        data = box2mask_batch(args.batchSize, args.fineSize, opt.label_nc, seed=100 + i)
        losses, reconstructed = model.module.forward(                              # :62-69 (optimizer steps happen inside)
            data["label_map"], None, data["mask_ctx_in"], None, data["mask_out"], data["mask_obj_inst"], data["cls"],
            data["mask_in"], eval_mode=False)
        loss_dict = dict(zip(model.module.loss_names, losses))                     # :74
        if i % 5 == 0:
            generated = model.module.generate(dict(label_map=data["label_map"], mask_obj_in=None,   # :90-100
                                                   mask_ctx_in=data["mask_ctx_in"], mask_obj_out=None,
                                                   mask_out=data["mask_out"], mask_obj_inst=data["mask_obj_inst"],
                                                   cls=data["cls"], mask_in=data["mask_in"]))
            inside = data["mask_out"].bool()
            acc = float((generated["comb_pred_label"].cpu() == data["label_map"].long())[inside].float().mean())
            print("iter %d  " % i + "  ".join("%s %.4f" % (k, float(v)) for k, v in loss_dict.items())
                  + "  eval-mode layout accuracy inside the box %.3f" % acc)
    torch.cuda.synchronize()
    print("%.1f images/sec (host-driven loop incl. synthetic data generation and the eval passes)" %
          (args.iters * args.batchSize / (time.time() - t0)))
    model.module.save("latest")                                                    # :126
    model.module.update_learning_rate(opt.niter + 1, 0)                            # :146


if __name__ == "__main__":
    main()
