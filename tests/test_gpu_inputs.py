"""-m gpu: the device input pipeline (csrc/input_pipeline.cu through the C-ABI) against tests/golden/inputs.npz -- the
reference's own mask transforms + cv2.resize composed in build_transforms' order (oracle/make_golden.py inputs) -- and
against the oracle restatement.  Images: the uint8 resize is bit-exact, the float32 normalisation uses the same two IEEE
operations -> exact equality asserted.  Masks: gathers, max and compares are exact; softmax goes through expf -> 1e-6."""
import numpy as np
import pytest
import torch

from oracle import input_oracle as IO
from tests import golden_inputs as GI

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


@pytest.mark.parametrize("name", list(GI.INPUT_CASES))
def test_input_pipeline_matches_reference_golden(golden, name):
    from bpbreid_b200 import data
    c = GI.input_case(name)
    g = golden("inputs")
    pipe = data.DeviceInputPipeline(c["height"], c["width"], masks_preprocess=c["preprocess"],
                                    softmax_weight=c["softmax_weight"], mask_filtering_threshold=c["threshold"],
                                    background_computation_strategy=c["strategy"], device=DEV)
    imgs, masks = pipe(c["images"], c["fields"], c["params"])
    torch.cuda.synchronize()
    np.testing.assert_array_equal(imgs.cpu().numpy(), g[f"{name}/images"])
    err = np.abs(masks.cpu().numpy() - g[f"{name}/masks"]).max()
    print(f"input pipeline {name}: images bit-identical, masks max abs err {err:.2e}")
    assert masks.shape == g[f"{name}/masks"].shape and err < 1e-6
    assert abs(float(masks.sum(1).mean()) - 1.0) < 1e-5                 # softmax rows


@pytest.mark.parametrize("strategy,softmax_weight,combine", [("threshold", 15, "max"), ("sum", 0, "sum"),
                                                             ("diff_from_max", 4, "max")])
def test_input_pipeline_random_sizes_vs_oracle(strategy, softmax_weight, combine):
    """Odd source sizes (up- and down-scaling on both axes, 1-pixel-wide sources), every background strategy, both
    combine modes and the softmax-free normalisation (mask_transform.py:80-83, 0/0 -> NaN like the reference) against the
    oracle, which is pinned bit-for-bit to cv2.resize and to the reference's mask classes by the CPU suite."""
    from bpbreid_b200 import data
    rs = np.random.RandomState(9)
    sizes = [(17, 300), (301, 9), (1, 64), (64, 1), (255, 129), (256, 128), (700, 333), (96, 48)]
    images = [rs.randint(0, 256, (h, w, 3)).astype(np.uint8) for h, w in sizes]
    fields = [(rs.rand(36, max(h // 7, 1), max(w // 7, 1)) * (rs.rand(36, max(h // 7, 1), max(w // 7, 1)) > 0.5)).astype(np.float32)
              for h, w in sizes]
    pipe = data.DeviceInputPipeline(256, 128, transforms=["rf", "rc", "re"], masks_preprocess="five_v", device=DEV,
                                    background_computation_strategy=strategy, softmax_weight=softmax_weight,
                                    combine_mode=combine)
    params = pipe.draw_params(len(images), np.random.RandomState(4))
    imgs, masks = pipe(images, fields, params)
    torch.cuda.synchronize()
    tab = data.group_table("five_v")
    groups = [np.nonzero(r)[0].tolist() for r in tab]
    for i, (im, f, p) in enumerate(zip(images, fields, params)):
        ref_i = IO.preprocess_image(im, 256, 128, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225), p["flip"], p["crop"], p["erase"])
        ref_m = IO.preprocess_masks(f, 256, 128, groups, 4, strategy, 0.5, softmax_weight, p["flip"], p["crop"], p["erase"],
                                    combine=combine)
        np.testing.assert_array_equal(imgs[i].cpu().numpy(), ref_i)
        np.testing.assert_allclose(masks[i].cpu().numpy(), ref_m, rtol=0, atol=1e-6, equal_nan=True)


def test_input_pipeline_feeds_the_model():
    """The pipeline's outputs are what BPBreID.forward takes (images + external_parts_masks, bpbreid.py:222)."""
    from bpbreid_b200 import data
    c = GI.model_case("resnet50_k5", n=4)
    m = c["model"]
    m.load_state_dict(c["state_dict"])
    m = m.to(DEV).eval()
    ic = GI.input_case("test_256_diffmax")
    pipe = data.DeviceInputPipeline(256, 128, masks_preprocess="five_v", device=DEV)
    imgs, masks = pipe(ic["images"][:4], ic["fields"][:4])
    with torch.no_grad():
        emb, vis, ids, pix, _, out_masks = m(imgs, external_parts_masks=masks)
    torch.cuda.synchronize()
    assert emb["parts"].shape == (4, 5, 512) and torch.isfinite(emb["parts"]).all()
