"""CPU: host-side logic and the C ABI surface (no compute calls: there is no GPU here)."""
import ctypes
import os

import pytest
import torch

from bpbreid_b200 import _lib, losses, metrics, models
from tests import golden_inputs as GI


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -m bpbreid_b200.build"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _lib.header_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert set(_lib._SIGS) | {"bpb_last_error", "bpb_abi_version"} == set(syms)
    loaded = _lib.load()
    assert loaded.bpb_abi_version() == 2


def test_model_builder_api_and_errors():
    with pytest.raises(KeyError):
        models.build_model("resnet9000", 10)
    cfg = GI.make_cfg("hrnet32", 5, pooling="nope")
    with pytest.raises(ValueError):
        models.build_model("bpbreid", 10, loss="part_based", pretrained=False, config=cfg)
    with pytest.raises(KeyError):
        models.build_model("bpbreid", 10, loss="part_based", pretrained=False, config=GI.make_cfg("vgg", 5))
    m = GI.build_product_model("resnet50", 5, 17)
    # extra kwargs are swallowed like the reference (models/__init__.py:128-134 always passes use_gpu)
    assert m.parts_num == 5 and m.num_classes == 17
    assert m.concat_parts_identity_classifier.classifier.weight.shape == (17, 5 * 512)
    assert m.global_identity_classifier.bn.bias.requires_grad is False      # bpbreid.py:407
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 3, 256, 128))


@pytest.mark.parametrize("backbone,K,nkeys", [("hrnet32", 5, 2022), ("resnet50", 5, 409), ("fastreid_resnet_ibn", 8, 451)])
def test_state_dict_key_counts(backbone, K, nkeys):
    # key *names* are checked against the reference in the build container by oracle/make_golden.py
    # (load_state_dict(strict=True) into the reference model); counts are from SURVEY.md section 8f-1
    m = GI.build_product_model(backbone, K)
    assert len(m.state_dict()) == nkeys


def test_loss_and_metric_api_errors():
    with pytest.raises(ValueError):
        losses.init_part_based_triplet_loss("no_such_loss")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        losses.CrossEntropyLoss()(torch.zeros(4, 10), torch.zeros(4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        losses.PartAveragedTripletLoss()(torch.zeros(4, 2, 8), torch.zeros(4, dtype=torch.long))
    with pytest.raises(ValueError):
        losses.BodyPartAttentionLoss(loss_type="zz")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            metrics.evaluate_rank(torch.zeros(2, 3).numpy(), [0, 1], [0, 1, 2], [0, 0], [1, 1, 1])


def test_planner_tiling_rule():
    from bpbreid_b200.plan import _tc_tileable
    assert _tc_tileable(32, 32, 64, 32) and _tc_tileable(256, 256, 8, 4) and _tc_tileable(64, 64, 16, 8)
    assert not _tc_tileable(3, 64, 128, 64)
    # 384x128 inputs: the 24x8 / 12x4 maps tile as (images x rows x width) = 2x8x8 / 8x4x4 (tc_gemm.cu conv_tiling)
    assert _tc_tileable(128, 128, 24, 8) and _tc_tileable(256, 256, 12, 4) and _tc_tileable(32, 32, 96, 32)
    assert not _tc_tileable(64, 64, 10, 12)          # width must divide 128
    assert _lib.query("bpb_conv_num_mtiles", 4, 24, 8) == 2 * 3 and _lib.query("bpb_conv_num_mtiles", 11, 12, 4) == 2 * 3
    assert _lib.query("bpb_conv_num_mtiles", 256, 64, 32) == 256 * 16


def test_flat_parameter_store_keeps_state_dict_contract():
    """bpbreid_b200.train.FlatParams re-points every nn.Parameter into one fp32 buffer: names, shapes and values of the
    state_dict are unchanged (reference checkpoints keep loading), load_state_dict writes through into the flat buffer,
    and the gradient buckets of the data-parallel all-reduce tile the buffer in backward order."""
    from bpbreid_b200 import train
    m = GI.build_product_model("hrnet32", 5, 751)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    fl = train.FlatParams(m, "cpu")
    after = m.state_dict()
    assert list(before) == list(after)
    for k in before:
        assert torch.equal(before[k], after[k]), k
    lo, hi = fl.p.data_ptr(), fl.p.data_ptr() + 4 * fl.total
    for name, p, off, n in fl.entries:
        assert lo <= p.data_ptr() < hi and p.data_ptr() == lo + 4 * off and off % train.ALIGN == 0
    sd = {k: (v + 1 if v.is_floating_point() else v) for k, v in before.items()}
    m.load_state_dict(sd)
    name, p, off, n = fl.entries[3]
    assert torch.equal(fl.p[off:off + n].view(p.shape), sd[name])         # written through
    ts = train.TrainStep(m)
    ts.flat = fl

    class FakePlan:
        marks = [(100, "stage4"), (300, "stage3"), (500, "stage2")]
        bops = [None] * 700
    b = ts._buckets(FakePlan())
    assert [x[0] for x in b] == [100, 300, 500, 700]
    assert b[0][1][1] == fl.total and b[-1][1][0] == 0
    assert all(b[i][1][0] == b[i + 1][1][1] for i in range(3))             # contiguous, descending
    assert b[0][1][1] - b[0][1][0] > 0.7 * fl.total                         # head + stage 4 + incre = most parameters


def test_extract_test_embeddings_matches_reference_semantics():
    from bpbreid_b200 import engine
    n, K, D = 3, 5, 8
    g = torch.Generator().manual_seed(0)
    emb = {"bn_foreg": torch.randn(n, D, generator=g), "parts": torch.randn(n, K, D, generator=g)}
    vis = {"foreg": torch.rand(n, generator=g) > 0.5, "parts": torch.rand(n, K, generator=g) > 0.5}
    masks = {"foreg": torch.rand(n, 4, 2, generator=g), "parts": torch.rand(n, K, 4, 2, generator=g)}
    f, v, pm, pix = engine.extract_test_embeddings((emb, vis, None, "pix", None, masks), ["bn_foreg", "parts"])
    assert f.shape == (n, K + 1, D) and v.shape == (n, K + 1) and pm.shape == (n, K + 1, 4, 2) and pix == "pix"
    assert torch.equal(f[:, 0], emb["bn_foreg"]) and torch.equal(f[:, 1:], emb["parts"])
    assert torch.equal(v[:, 0], vis["foreg"])                              # bn_foreg takes the visibility of foreg


def test_input_pipeline_host_logic():
    """bpbreid_b200.data: group tables, parameter draws and error behaviour (no device needed)."""
    import numpy as np
    from bpbreid_b200 import data
    t5, t8 = data.group_table("five_v"), data.group_table("eight")
    assert t5.shape == (5, 36) and t8.shape == (8, 36) and t5.sum() == 36 and t8.sum() == 36 + 1  # 'eight' shares one joint
    with pytest.raises(KeyError):
        data.group_table("eleven")
    pipe = data.DeviceInputPipeline(384, 128, transforms=["rf", "rc", "re"], masks_preprocess="five_v", device="cuda:0")
    assert (pipe.out_h, pipe.out_w, pipe.parts_num) == (96, 32, 5) and len(pipe.parts_names) == 5
    ps = pipe.draw_params(400, np.random.RandomState(0))
    assert 0.35 < np.mean([p["flip"] for p in ps]) < 0.65 and 0.35 < np.mean([p["erase"] is not None for p in ps]) < 0.65
    for p in ps:
        assert 0 <= p["crop"][0] <= 20 and 0 <= p["crop"][1] <= 20
        if p["erase"] is not None:
            x0, y0, x1, y1 = p["erase"]
            assert 0 <= x0 < x1 <= 128 and 0 <= y0 < y1 <= 384
            assert int(128 * 0.15) <= x1 - x0 <= int(128 * 0.65) and int(384 * 0.15) <= y1 - y0 <= int(384 * 0.65)
    test_only = data.DeviceInputPipeline(256, 128, transforms=None, device="cuda:0")
    assert all(p == dict(flip=False, crop=None, erase=None) for p in test_only.draw_params(3, np.random.RandomState(1)))
    with pytest.raises(ValueError):
        data.DeviceInputPipeline(256, 128, background_computation_strategy="median", device="cuda:0")
    with pytest.raises(NotImplementedError):
        data.DeviceInputPipeline(256, 128, transforms=["random_occlusion"], device="cuda:0")


def test_rerank_rejects_cpu_inputs_without_a_device():
    """re_ranking has no CPU path: on a box without CUDA it must raise, never fall back (prompt, section 3)."""
    import numpy as np
    import torch
    from bpbreid_b200 import rerank
    if torch.cuda.is_available():
        pytest.skip("CUDA present: covered by tests/test_gpu_rerank.py")
    d = np.zeros((2, 3), np.float32)
    with pytest.raises(RuntimeError):
        rerank.re_ranking(d, np.zeros((2, 2), np.float32), np.zeros((3, 3), np.float32))


def test_ctypes_signatures_agree_with_the_header():
    """Every prototype of include/bpbreid_sm100.h parsed and compared, argument by argument, with the ctypes table of
    bpbreid_b200/_lib.py: pointers -> c_void_p, int -> c_int, int64_t -> c_int64, float -> c_float.  A drifted binding
    would otherwise pass garbage across the ABI silently."""
    import re
    hdr = open(os.path.join(os.path.dirname(_lib.__file__), "..", "include", "bpbreid_sm100.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", " ", hdr)
    protos = re.findall(r"\b(?:int|int64_t|const char\*)\s+(bpb_\w+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)
    assert len(protos) >= 70
    kind = {"c_void_p": "ptr", "c_int": "int", "c_long": "i64", "c_longlong": "i64", "c_float": "float"}
    checked = 0
    for name, args in protos:
        if name not in _lib._SIGS:
            continue
        args = " ".join(args.split())
        parts = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        want = []
        for a in parts:
            if "*" in a:
                want.append("ptr")
            elif re.match(r"(const\s+)?int64_t\b", a):
                want.append("i64")
            elif re.match(r"(const\s+)?float\b", a):
                want.append("float")
            elif re.match(r"(const\s+)?int\b", a):
                want.append("int")
            else:
                raise AssertionError(f"{name}: unparsed parameter '{a}'")
        got = [kind[t.__name__] for t in _lib._SIGS[name]]
        assert got == want, f"{name}: header {want} vs ctypes {got}"
        checked += 1
    assert checked == len(_lib._SIGS)
