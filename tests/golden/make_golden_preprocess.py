"""Generates tests/golden/preprocess_ref.npz: outputs of Pillow's BICUBIC resize and of transformers' PIL-backed CLIP image
processor (the code path of the reference's pinned transformers==4.36.1 CLIPImageProcessor) on seeded synthetic RGB tiles.

    python tests/golden/make_golden_preprocess.py        # run in the build container (needs Pillow + transformers)

Inputs are not stored: they are regenerated from the seed by `synthetic_image` below (imported by the tests), so the fixture
only holds the expected uint8 crops and, for two cases, the float32 pixel_values.
"""
import os

import numpy as np

CASES = [(300, 260), (512, 512), (128, 160), (224, 224), (600, 800), (97, 333), (224, 500), (256, 256), (1024, 768)]
FLOAT_CASES = [(300, 260), (128, 160)]


def synthetic_image(h, w, seed):
    """Smooth gradients + texture + noise, uint8 (h, w, 3): exercises the clamp (overshoot of the bicubic lobes) and every tap."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    base = np.stack([127 + 120 * np.sin(x / 7.0 + seed), 127 + 120 * np.cos(y / 5.0), 255.0 * ((x // 8 + y // 8) % 2)], -1)
    noise = rng.integers(-40, 41, (h, w, 3))
    return np.clip(base + noise, 0, 255).astype(np.uint8)


def main():
    import PIL
    import PIL.Image as Image
    import transformers
    from transformers.models.clip import CLIPImageProcessorPil
    proc = CLIPImageProcessorPil()
    out = {"pillow_version": np.array(PIL.__version__), "transformers_version": np.array(transformers.__version__)}
    for i, (h, w) in enumerate(CASES):
        img = synthetic_image(h, w, i)
        short, long_ = (w, h) if w <= h else (h, w)
        ns, nl = 224, int(224 * long_ / short)
        nw, nh = (ns, nl) if w <= h else (nl, ns)
        r = np.asarray(Image.fromarray(img).resize((nw, nh), resample=Image.BICUBIC))
        top, left = (nh - 224) // 2, (nw - 224) // 2
        out[f"crop_{h}x{w}"] = r[top:top + 224, left:left + 224].copy()
        pv = proc(images=Image.fromarray(img), return_tensors="np")["pixel_values"][0]
        # the processor's own crop must equal Pillow's resize + crop pushed through the documented float formula
        x = (out[f"crop_{h}x{w}"] * (1 / 255)).astype(np.float32)
        y = ((x - np.array(proc.image_mean, np.float32)) / np.array(proc.image_std, np.float32)).transpose(2, 0, 1)
        assert np.array_equal(y, pv), (h, w)
        if (h, w) in FLOAT_CASES:
            out[f"pixel_values_{h}x{w}"] = pv.astype(np.float32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "preprocess_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
