"""ORACLE (test infrastructure only; never imported by the product path).

CPU restatement of the reference's per-sample input transforms for the part-based pipeline:
`torchreid/data/transforms.py:24-163` (build_transforms: Resize -> [flip] -> [pad 10 + crop] -> Normalize ->
[CoarseDropout] -> ToTensorV2, masks: PermuteMasksDim -> grouping -> AddBackgroundMask -> ResizeMasks) and
`torchreid/data/masks_transforms/mask_transform.py:21-85`.

Third-party arithmetic outside /root/reference, restated from the published algorithms:
  * `cv2.resize` (opencv-python, unpinned in requirements.txt:10; 4.13.0 in this image) as used by albumentations.Resize:
    INTER_LINEAR on uint8 images (11-bit fixed-point coefficients, `resize.cpp` HResizeLinear / VResizeLinear 8u) and
    INTER_NEAREST on the float mask stack -- pinned bit-for-bit against cv2 itself by tests/test_oracle_vs_golden.py;
  * `albumentations.Normalize` (requirements.txt:21, unpinned, ABSENT from this image -> parity for this one step is
    pinned to the published 1.x algorithm only): img.astype(float32); img -= mean*255; img *= 1/(std*255);
  * `torch.nn.functional.interpolate(mode='nearest')` and `softmax` (mask_transform.py:52,81) -- torch is present.
The mask classes of the reference itself are imported (with a stub for the albumentations base class) by
`oracle/make_golden.py inputs` to mint tests/golden/inputs.npz.

Random augmentations: albumentations draws from Python's `random`; the stream cannot be reproduced without the package,
so flip / crop offset / erase rectangle are PARAMETERS here and in the CUDA path (the host draws them).
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def _linear_coeffs(dst, src, clamp_weights):
    """resize.cpp (INTER_LINEAR): source index and 11-bit fixed-point weights for every destination coordinate.
    Horizontally (clamp_weights) an out-of-range tap moves the whole weight onto the border pixel; vertically the
    weights stay as computed and the two ROW INDICES are clipped separately (so a border row is added to itself with
    two truncations -- visible as a round-down of exact .5 ties)."""
    scale = float(src) / float(dst)
    idx = np.zeros(dst, np.int64)
    w = np.zeros((dst, 2), np.int64)
    for d in range(dst):
        fx = np.float32((d + 0.5) * scale - 0.5)
        sx = int(np.floor(fx))
        fx = np.float32(fx - np.float32(sx))
        if clamp_weights:
            if sx < 0:
                sx, fx = 0, np.float32(0)
            if sx >= src - 1:
                sx, fx = src - 1, np.float32(0)
        idx[d] = sx
        # saturate_cast<short>(float): round half to even
        w[d, 0] = int(np.rint(np.float32((np.float32(1) - fx) * np.float32(COEF_SCALE))))
        w[d, 1] = int(np.rint(np.float32(fx * np.float32(COEF_SCALE))))
    return idx, w


def resize_linear_u8(img, height, width):
    """cv2.resize(img, (width, height), interpolation=cv2.INTER_LINEAR) for uint8 HWC images."""
    h, w, _ = img.shape
    if (h, w) == (height, width):
        return img.copy()
    xi, xw = _linear_coeffs(width, w, True)
    yi, yw = _linear_coeffs(height, h, False)
    s = img.astype(np.int64)
    x1 = np.minimum(xi + 1, w - 1)
    rows = s[:, xi, :] * xw[:, 0][None, :, None] + s[:, x1, :] * xw[:, 1][None, :, None]       # [h, width, c], scale 2^11
    r0, r1 = rows[np.clip(yi, 0, h - 1)], rows[np.clip(yi + 1, 0, h - 1)]
    b0, b1 = yw[:, 0][:, None, None], yw[:, 1][:, None, None]
    out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def nearest_index_cv2(dst, src):
    """cv2 INTER_NEAREST: min(floor(x * (1 / (dst/src))), src - 1) in double precision."""
    ifx = 1.0 / (float(dst) / float(src))
    return np.minimum(np.floor(np.arange(dst) * ifx).astype(np.int64), src - 1)


def nearest_index_torch(dst, src):
    """F.interpolate(mode='nearest'): floor(x * (src / dst)) with a float32 scale."""
    scale = np.float32(src) / np.float32(dst)
    return np.minimum(np.floor(np.arange(dst, dtype=np.float32) * scale).astype(np.int64), src - 1)


def normalize(img_u8, mean, std):
    """albumentations.Normalize (1.x functional.normalize), float32 throughout."""
    mean = np.array(mean, dtype=np.float32) * np.float32(255.0)
    std = np.array(std, dtype=np.float32) * np.float32(255.0)
    den = np.reciprocal(std, dtype=np.float32)
    img = img_u8.astype(np.float32)
    img -= mean
    img *= den
    return img


def preprocess_image(img, height, width, mean, std, flip=False, crop=None, erase=None, pad=10,
                     erase_fill=(0.485, 0.456, 0.406)):
    """transforms.py:73-124 for one RGB uint8 HWC image -> float32 [3, height, width].
    crop = (x0, y0) inside the image zero-padded by `pad` on every side (PadIfNeeded + RandomCrop, :92-95);
    erase = (x0, y0, x1, y1) filled with `erase_fill` AFTER normalisation (CoarseDropout sits behind Normalize, :112-117)."""
    x = resize_linear_u8(img, height, width)
    if flip:
        x = x[:, ::-1]
    if crop is not None:
        p = np.zeros((height + 2 * pad, width + 2 * pad, 3), np.uint8)
        p[pad:pad + height, pad:pad + width] = x
        x = p[crop[1]:crop[1] + height, crop[0]:crop[0] + width]
    f = normalize(np.ascontiguousarray(x), mean, std)
    if erase is not None:
        f[erase[1]:erase[3], erase[0]:erase[2]] = np.array(erase_fill, np.float32)
    return np.ascontiguousarray(f.transpose(2, 0, 1))


def preprocess_masks(fields, height, width, groups, mask_scale=4, strategy="threshold", threshold=0.5,
                     softmax_weight=15, flip=False, crop=None, erase=None, pad=10, combine="max"):
    """One sample's part-confidence stack [36, hm, wm] float32 -> [K+1, height/scale, width/scale] float32.
    groups: K lists of channel indices (MaskGroupingTransform, mask_transform.py:21-38)."""
    import torch
    c, hm, wm = fields.shape
    yi, xi = nearest_index_cv2(height, hm), nearest_index_cv2(width, wm)
    m = fields[:, yi][:, :, xi]                                         # albumentations.Resize on masks: INTER_NEAREST
    if flip:
        m = m[:, :, ::-1]
    if crop is not None:
        p = np.zeros((c, height + 2 * pad, width + 2 * pad), np.float32)   # mask_value=0
        p[:, pad:pad + height, pad:pad + width] = m
        m = p[:, crop[1]:crop[1] + height, crop[0]:crop[0] + width]
    m = np.ascontiguousarray(m)
    if erase is not None:
        m[:, erase[1]:erase[3], erase[0]:erase[2]] = 0                  # mask_fill_value=0
    t = torch.from_numpy(m)
    parts = []
    for g in groups:
        sel = t[list(g)]
        parts.append((sel.sum(0) if combine == "sum" else sel.max(0)[0]).clamp(0, 1))
    t = torch.stack(parts)
    if strategy == "sum":                                               # mask_transform.py:67-79
        bg = (1 - t.sum(0)).clamp(0, 1)
    elif strategy == "threshold":
        bg = (t.max(0)[0] < threshold).to(t.dtype)
    elif strategy == "diff_from_max":
        bg = (1 - t.max(0)[0]).clamp(0, 1)
    else:
        raise ValueError("Background mask combine strategy {} not supported".format(strategy))
    t = torch.cat([bg.unsqueeze(0), t])
    t = torch.softmax(t * softmax_weight, 0) if softmax_weight > 0 else t / t.sum(0)
    ho, wo = int(height / mask_scale), int(width / mask_scale)
    oy, ox = nearest_index_torch(ho, height), nearest_index_torch(wo, width)
    return t.numpy()[:, oy][:, :, ox].astype(np.float32)
