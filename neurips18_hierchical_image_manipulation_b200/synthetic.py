"""Synthetic Cityscapes-shaped batches: the tensor contract of the reference's data layer
(data/segmentation_dataset.py:78-131, train_mask2image.py:58-65) without its file IO.  CPU fp32 NCHW tensors:
label (B,1,H,W) integer-valued in [0,label_nc), inst (B,1,H,W), image (B,3,H,W) in [-1,1], mask_in / mask_out
(B,1,H,W) in {0,1} -- one axis-aligned box per sample, mask_out a dilated copy (segmentation_dataset.py:113-119)."""
import torch
import torch.nn.functional as F


def synthetic_batch(B, H, W, label_nc=35, seed=1234):
    g = torch.Generator().manual_seed(seed)
    bh, bw = max(H // 16, 1), max(W // 16, 1)
    lab = torch.randint(0, label_nc, (B, 1, bh, bw), generator=g).float()
    label = F.interpolate(lab, size=(H, W), mode="nearest")
    ins = torch.randint(0, 20, (B, 1, bh, bw), generator=g).float()
    inst = F.interpolate(ins, size=(H, W), mode="nearest")
    image = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    mask_in = torch.zeros(B, 1, H, W)
    mask_out = torch.zeros(B, 1, H, W)
    for b in range(B):
        side_h = int(torch.randint(max(H // 8, 1), max(H // 2, 2), (1,), generator=g))
        side_w = int(torch.randint(max(H // 8, 1), max(H // 2, 2), (1,), generator=g))
        y0 = int(torch.randint(0, H - side_h + 1, (1,), generator=g))
        x0 = int(torch.randint(0, W - side_w + 1, (1,), generator=g))
        mask_in[b, :, y0:y0 + side_h, x0:x0 + side_w] = 1
        dh, dw = int(side_h * 0.15), int(side_w * 0.15)
        mask_out[b, :, max(0, y0 - dh):min(H, y0 + side_h + dh), max(0, x0 - dw):min(W, x0 + side_w + dw)] = 1
    return dict(label=label, inst=inst, image=image, mask_in=mask_in, mask_out=mask_out)


def box2mask_batch(B, S, label_nc, seed):
    """Synthetic box2mask sample set (what data/ + TwoStreamAE_mask.encode_input :127-151 consume): a blocky label map, a
    box (mask_in), its margin-expanded region (mask_out), an elliptical instance mask of class `cls` inside the box and
    the context map with the region wiped."""
    g = torch.Generator().manual_seed(seed)
    lab = torch.randint(1, label_nc, (B, 1, S // 8, S // 8), generator=g).float()
    label_map = torch.nn.functional.interpolate(lab, size=(S, S), mode="nearest")
    cls = torch.randint(1, label_nc - 1, (B, 1), generator=g)
    mask_out, mask_in, inst = torch.zeros(B, 1, S, S), torch.zeros(B, 1, S, S), torch.zeros(B, 1, S, S)
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    for b in range(B):
        y0, x0 = int(torch.randint(4, S // 4, (1,), generator=g)), int(torch.randint(4, S // 4, (1,), generator=g))
        h, w = int(torch.randint(S // 4, S // 2, (1,), generator=g)), int(torch.randint(S // 4, S // 2, (1,), generator=g))
        mask_in[b, :, y0:y0 + h, x0:x0 + w] = 1
        mask_out[b, :, max(0, y0 - 4):y0 + h + 4, max(0, x0 - 4):x0 + w + 4] = 1
        ell = (((yy - (y0 + h / 2.0)) / (h / 2.0)) ** 2 + ((xx - (x0 + w / 2.0)) / (w / 2.0)) ** 2) <= 1.0
        inst[b, 0] = ell.float()
        label_map[b, 0][ell] = float(cls[b, 0])
    return dict(label_map=label_map, mask_ctx_in=label_map * (1 - mask_out), mask_out=mask_out, mask_in=mask_in,
                mask_obj_inst=inst, cls=cls.float())
