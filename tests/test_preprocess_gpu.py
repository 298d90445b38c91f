"""CLIP preprocessing kernel (lhrs_clip_preprocess) and the input stager, through the C ABI.

Bars: the uint8 resize + crop is BIT-EXACT with the oracle (= Pillow, tests/test_oracle.py) on every geometry, incl. the golden
fixture produced by Pillow / transformers themselves; float32 pixel_values are exactly equal; bf16 pixel_values equal the
float32 ones rounded to nearest even (what ``.to(torch.bfloat16)`` does to the reference's tensor)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _mk():
    spec = importlib.util.spec_from_file_location("mkpre", os.path.join(GOLDEN, "make_golden_preprocess.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    return mk


def test_preprocess_matches_golden_and_oracle():
    from oracle import preprocess as P
    from lhrs_bot_b200.preprocess import ClipPreprocessor
    mk = _mk()
    ref = np.load(os.path.join(GOLDEN, "preprocess_ref.npz"))
    pre32, pre16 = ClipPreprocessor(dtype=torch.float32), ClipPreprocessor()
    for i, (h, w) in enumerate(mk.CASES):
        img = mk.synthetic_image(h, w, i)
        dev = torch.from_numpy(img)[None].to(DEV)
        out32, crop = pre32.preprocess_device(dev, return_resized=True)
        assert np.array_equal(crop[0].cpu().numpy(), ref[f"crop_{h}x{w}"]), f"{h}x{w}: uint8 crop differs from Pillow's"
        want = P.clip_preprocess(img)
        assert np.array_equal(out32[0].cpu().numpy(), want), f"{h}x{w}: float32 pixel_values differ from the oracle"
        if (h, w) in mk.FLOAT_CASES:
            assert np.array_equal(out32[0].cpu().numpy(), ref[f"pixel_values_{h}x{w}"]), f"{h}x{w}: differs from the HF processor"
        out16 = pre16.preprocess_device(dev)
        assert out16.dtype == torch.bfloat16 and torch.equal(out16.cpu(), torch.from_numpy(want)[None].to(torch.bfloat16))


@pytest.mark.parametrize("shape", [(4, 512, 512), (3, 333, 97), (2, 100, 1000), (5, 224, 224), (1, 2048, 1536), (2, 31, 47)])
def test_preprocess_batches_random_tiles(shape):
    """Random noise tiles (worst case for the clamp), batches, extreme aspect ratios, rows wider than 48 KB of shared memory."""
    from oracle import preprocess as P
    from lhrs_bot_b200.preprocess import ClipPreprocessor
    B, h, w = shape
    rng = np.random.default_rng(h * 7 + w)
    imgs = rng.integers(0, 256, (B, h, w, 3), dtype=np.uint8)
    out, crop = ClipPreprocessor(dtype=torch.float32).preprocess_device(torch.from_numpy(imgs).to(DEV), return_resized=True)
    for b in range(B):
        assert np.array_equal(crop[b].cpu().numpy(), P.resized_crop_u8(imgs[b])), f"image {b}"
        assert np.array_equal(out[b].cpu().numpy(), P.clip_preprocess(imgs[b]))


def test_processor_call_shape_mixed_geometries_and_errors():
    from oracle import preprocess as P
    from lhrs_bot_b200.preprocess import ClipPreprocessor
    rng = np.random.default_rng(3)
    imgs = [rng.integers(0, 256, s, dtype=np.uint8) for s in [(300, 200, 3), (256, 256, 3), (300, 200, 3)]]
    pv = ClipPreprocessor(dtype=torch.float32)(imgs, return_tensors="pt")["pixel_values"]
    assert pv.shape == (3, 3, 224, 224) and pv.is_cuda
    for i, im in enumerate(imgs):
        assert np.array_equal(pv[i].cpu().numpy(), P.clip_preprocess(im))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ClipPreprocessor().preprocess_device(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))
    with pytest.raises(ValueError):
        ClipPreprocessor().preprocess_device(torch.zeros(1, 8, 8, 4, dtype=torch.uint8, device=DEV))


def test_input_stager_overlapped_batches():
    """The stager yields every batch once, in order, on the device, with uint8 rgb tiles preprocessed on its side stream."""
    from oracle import preprocess as P
    from lhrs_bot_b200.preprocess import ClipPreprocessor, InputStager
    rng = np.random.default_rng(0)
    host = [dict(rgb=torch.from_numpy(rng.integers(0, 256, (2, 256, 256, 3), dtype=np.uint8)),
                 input_ids=torch.arange(i, i + 12).view(2, 6), tag=i) for i in range(5)]
    seen = []
    for i, b in enumerate(InputStager(host, DEV, ClipPreprocessor(dtype=torch.float32))):
        assert b["rgb"].is_cuda and b["rgb"].shape == (2, 3, 224, 224) and b["input_ids"].is_cuda and b["tag"] == i
        assert torch.equal(b["input_ids"].cpu(), host[i]["input_ids"])
        assert np.array_equal(b["rgb"][1].cpu().numpy(), P.clip_preprocess(host[i]["rgb"][1].numpy()))
        seen.append(i)
    assert seen == list(range(5))
