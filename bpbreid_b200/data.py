"""Input pipeline on the device: host-side mirror of the per-sample transform list that
`torchreid.data.transforms.build_transforms` composes (transforms.py:24-163) and of the mask transforms it ends with
(`data/masks_transforms/mask_transform.py:21-85`, grouping tables `pifpaf_mask_transform.py`).

The reference runs this list per sample in albumentations / cv2 worker processes.  Here a batch of decoded images (RGB
uint8, any sizes) and their 36-channel pifpaf confidence stacks goes to the GPU in one pinned copy each and two kernel
launches produce the model inputs: images [N,3,H,W] float32 and external part masks [N,K+1,H/4,W/4] float32
(csrc/input_pipeline.cu through `bpb_preprocess_images` / `bpb_preprocess_masks`).  cv2's fixed-point bilinear resize is
reproduced bit for bit, so features do not shift against a model trained with the reference's loader.

Random augmentations ('rf', 'rc', 're'): albumentations draws from Python's `random`; the draws are made here on the
host (`draw_params`) with the reference's distributions and handed to the kernels as per-sample parameters.
Not built: 'random_occlusion' (needs the occluder image bank) and 'color_jitter'.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

# pifpaf_mask_transform.py:6-20: the 17 keypoint + 19 joint confidence fields, in file order
PIFPAF_KEYPOINTS = ["nose", "left_eye", "right_eye", "left_ear", "right_ear", "left_shoulder", "right_shoulder",
                    "left_elbow", "right_elbow", "left_wrist", "right_wrist", "left_hip", "right_hip", "left_knee",
                    "right_knee", "left_ankle", "right_ankle"]
PIFPAF_JOINTS = ["left_ankle_to_left_knee", "left_knee_to_left_hip", "right_ankle_to_right_knee",
                 "right_knee_to_right_hip", "left_hip_to_right_hip", "left_shoulder_to_left_hip",
                 "right_shoulder_to_right_hip", "left_shoulder_to_right_shoulder", "left_shoulder_to_left_elbow",
                 "right_shoulder_to_right_elbow", "left_elbow_to_left_wrist", "right_elbow_to_right_wrist",
                 "left_eye_to_right_eye", "nose_to_left_eye", "nose_to_right_eye", "left_eye_to_left_ear",
                 "right_eye_to_right_ear", "left_ear_to_left_shoulder", "right_ear_to_right_shoulder"]
PIFPAF_PARTS = PIFPAF_KEYPOINTS + PIFPAF_JOINTS
_IDX = {k: i for i, k in enumerate(PIFPAF_PARTS)}

_HEAD = ["nose", "left_eye", "right_eye", "left_ear", "right_ear", "left_eye_to_right_eye", "nose_to_left_eye",
         "nose_to_right_eye", "left_eye_to_left_ear", "right_eye_to_right_ear", "left_ear_to_left_shoulder",
         "right_ear_to_right_shoulder"]
# the two groupings the shipped configs use (configs/bpbreid/*.yaml: masks.preprocess 'five_v' / 'eight';
# pifpaf_mask_transform.py:211-232, 352-374); checked against the reference's tables by tests/test_oracle_vs_golden.py
MASK_GROUPS = {
    "five_v": [
        ("head_mask", _HEAD),
        ("upper_arms_torso_mask", ["left_elbow", "right_elbow", "left_shoulder_to_left_elbow",
                                   "right_shoulder_to_right_elbow", "left_shoulder", "right_shoulder",
                                   "left_shoulder_to_right_shoulder"]),
        ("lower_arms_torso_mask", ["left_wrist", "right_wrist", "left_elbow_to_left_wrist", "right_elbow_to_right_wrist",
                                   "left_hip", "right_hip", "right_shoulder_to_right_hip"]),
        ("legs_mask", ["left_hip", "right_hip", "left_knee", "right_knee", "left_ankle_to_left_knee",
                       "left_knee_to_left_hip", "right_ankle_to_right_knee", "right_knee_to_right_hip"]),
        ("feet_mask", ["left_ankle", "right_ankle"])],
    "eight": [
        ("head_mask", _HEAD),
        ("left_arm_mask", ["left_shoulder", "left_elbow", "left_wrist", "left_shoulder_to_left_elbow",
                           "left_elbow_to_left_wrist"]),
        ("right_arm_mask", ["right_shoulder", "right_elbow", "right_wrist", "right_shoulder_to_right_elbow",
                            "right_elbow_to_right_wrist"]),
        ("torso_mask", ["left_hip", "right_hip", "left_hip_to_right_hip", "left_shoulder_to_left_hip",
                        "right_shoulder_to_right_hip", "left_shoulder_to_right_shoulder"]),
        ("left_leg_mask", ["left_knee", "left_ankle_to_left_knee", "left_knee_to_left_hip", "left_hip_to_right_hip"]),
        ("right_leg_mask", ["right_knee", "right_ankle_to_right_knee", "right_knee_to_right_hip"]),
        ("left_feet_mask", ["left_ankle"]),
        ("right_feet_mask", ["right_ankle"])],
}
_STRATEGIES = {"sum": 0, "threshold": 1, "diff_from_max": 2}


def group_table(preprocess):
    """[K, 36] uint8 membership matrix of a grouping (MaskGroupingTransform's parts_grouping over parts_map)."""
    if preprocess not in MASK_GROUPS:
        raise KeyError(f"masks preprocess '{preprocess}' is not built here (available: {sorted(MASK_GROUPS)})")
    tab = np.zeros((len(MASK_GROUPS[preprocess]), len(PIFPAF_PARTS)), np.uint8)
    for i, (_, members) in enumerate(MASK_GROUPS[preprocess]):
        for k in members:
            tab[i, _IDX[k]] = 1
    return tab


class DeviceInputPipeline:
    """build_transforms(height, width, config, mask_scale=4, transforms, norm_mean, norm_std, masks_preprocess,
    softmax_weight, mask_filtering_threshold, background_computation_strategy) for the pifpaf-mask configuration
    (transforms.py:24-163), applied to a whole batch on the GPU."""

    PAD = 10                                                            # transforms.py:93

    def __init__(self, height, width, mask_scale=4, transforms=("rc", "re"), norm_mean=(0.485, 0.456, 0.406),
                 norm_std=(0.229, 0.224, 0.225), masks_preprocess="eight", softmax_weight=15,
                 mask_filtering_threshold=0.5, background_computation_strategy="threshold", combine_mode="max",
                 device=None):
        if background_computation_strategy not in _STRATEGIES:
            raise ValueError("Background mask combine strategy {} not supported".format(background_computation_strategy))
        if isinstance(transforms, str):
            transforms = [transforms]
        self.transforms = [t.lower() for t in (transforms or [])]
        unknown = [t for t in self.transforms if t in ("random_occlusion", "ro", "color_jitter", "cj")]
        if unknown:
            raise NotImplementedError(f"transforms {unknown} are not built in the device pipeline")
        self.height, self.width, self.mask_scale = int(height), int(width), mask_scale
        self.out_h, self.out_w = int(height / mask_scale), int(width / mask_scale)          # mask_transform.py:49
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        mean = np.array(norm_mean, dtype=np.float32) * np.float32(255.0)                    # albumentations.Normalize
        std = np.array(norm_std, dtype=np.float32) * np.float32(255.0)
        self._mean255 = np.ascontiguousarray(mean)
        self._inv_std255 = np.ascontiguousarray(np.reciprocal(std, dtype=np.float32))
        self._fill = np.array([0.485, 0.456, 0.406], np.float32)                            # transforms.py:116
        tab = group_table(masks_preprocess)
        self.parts_num = tab.shape[0]
        self.parts_names = [n for n, _ in MASK_GROUPS[masks_preprocess]]
        self._bits = np.array([sum(1 << k for k in np.nonzero(r)[0]) for r in tab], dtype=np.uint64)
        self._strategy = _STRATEGIES[background_computation_strategy]
        self._combine_sum = 1 if combine_mode == "sum" else 0
        self.softmax_weight, self.threshold = float(softmax_weight), float(mask_filtering_threshold)

    # ---------------------------------------------------------------- augmentation parameters (host RNG)
    def draw_params(self, n, rng):
        """Per-sample flip / crop offset / erase rectangle with the reference's distributions (HorizontalFlip p=0.5,
        PadIfNeeded(10) + RandomCrop, CoarseDropout one hole 15-65 % of each side with p=0.5; transforms.py:87-117)."""
        h, w = self.height, self.width
        out = []
        for _ in range(n):
            flip = ("random_flip" in self.transforms or "rf" in self.transforms) and rng.rand() < 0.5
            crop = None
            if "random_crop" in self.transforms or "rc" in self.transforms:
                crop = (int(rng.randint(0, 2 * self.PAD + 1)), int(rng.randint(0, 2 * self.PAD + 1)))
            erase = None
            if ("random_erase" in self.transforms or "re" in self.transforms) and rng.rand() < 0.5:
                eh = int(rng.randint(int(h * 0.15), int(h * 0.65) + 1))
                ew = int(rng.randint(int(w * 0.15), int(w * 0.65) + 1))
                y0, x0 = int(rng.randint(0, h - eh + 1)), int(rng.randint(0, w - ew + 1))
                erase = (x0, y0, x0 + ew, y0 + eh)
            out.append(dict(flip=bool(flip), crop=crop, erase=erase))
        return out

    # ---------------------------------------------------------------- batch packing
    def _pack(self, arrays, shapes_hw, params, dtype):
        n = len(arrays)
        sizes = [a.size for a in arrays]
        offsets = np.zeros(n, np.int64)
        offsets[1:] = np.cumsum(sizes[:-1])
        flat = torch.empty(int(sum(sizes)), dtype=dtype).pin_memory()
        fv = flat.numpy()
        for a, o, s in zip(arrays, offsets, sizes):
            fv[o:o + s] = np.ascontiguousarray(a).reshape(-1)
        geom = np.zeros((n, 10), np.int32)
        for i, ((sh, sw), p) in enumerate(zip(shapes_hw, params)):
            crop = p.get("crop") or (-1, -1)
            erase = p.get("erase") or (0, 0, 0, 0)
            geom[i] = [sh, sw, int(bool(p.get("flip"))), crop[0], crop[1], erase[0], erase[1], erase[2], erase[3], 0]
        dev = self.device
        return (flat.to(dev, non_blocking=True), torch.from_numpy(offsets).to(dev), torch.from_numpy(geom).to(dev))

    def __call__(self, images, fields=None, params=None):
        """images: list of RGB uint8 HWC arrays (read_image, utils/tools.py:113-114); fields: list of float32
        [36, hm, wm] confidence stacks as stored on disk (read_masks before its transpose, utils/tools.py:139);
        params: per-sample dicts from `draw_params` (None: the test transform).  Returns CUDA tensors
        (images [N,3,H,W], masks [N,K+1,H/s,W/s] or None)."""
        n = len(images)
        if params is None:
            params = [dict(flip=False, crop=None, erase=None)] * n
        if len(params) != n or (fields is not None and len(fields) != n):
            raise ValueError("images, fields and params must have the same length")
        for im in images:
            if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
                raise ValueError("images must be RGB uint8 HWC arrays")
        s = _lib.stream_ptr()
        with torch.cuda.device(self.device):
            pix, off, geom = self._pack(images, [im.shape[:2] for im in images], params, torch.uint8)
            out = torch.empty(n, 3, self.height, self.width, dtype=torch.float32, device=self.device)
            _lib.call("bpb_preprocess_images", pix.data_ptr(), off.data_ptr(), geom.data_ptr(), n, self.height, self.width,
                      self.PAD, self._mean255.ctypes.data, self._inv_std255.ctypes.data, self._fill.ctypes.data,
                      out.data_ptr(), s)
            masks = None
            if fields is not None:
                for f in fields:
                    if f.dtype != np.float32 or f.ndim != 3 or f.shape[0] != len(PIFPAF_PARTS):
                        raise ValueError("fields must be float32 [36, h, w] pifpaf confidence stacks")
                fl, foff, fgeom = self._pack(fields, [f.shape[1:] for f in fields], params, torch.float32)
                masks = torch.empty(n, self.parts_num + 1, self.out_h, self.out_w, dtype=torch.float32, device=self.device)
                _lib.call("bpb_preprocess_masks", fl.data_ptr(), foff.data_ptr(), fgeom.data_ptr(), n, self.height,
                          self.width, self.out_h, self.out_w, self.PAD, self._bits.ctypes.data, self.parts_num,
                          self._combine_sum, self._strategy, self.threshold, self.softmax_weight, masks.data_ptr(), s)
                masks._keep = (fl, foff, fgeom)                          # inputs stay alive until the stream has read them
            out._keep = (pix, off, geom)
        return out, masks
