"""Throughput of lhrs_clip_preprocess (HBM-bound) next to Pillow + numpy on the host: python tools/preprocess_bench.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lhrs_bot_b200.preprocess import ClipPreprocessor  # noqa: E402


def main():
    pre = ClipPreprocessor()
    rows = []
    for (B, H, W) in [(32, 256, 256), (32, 512, 512), (32, 800, 800), (8, 1024, 1024)]:
        rng = np.random.default_rng(0)
        host = torch.from_numpy(rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)).pin_memory()
        dev = host.cuda()
        out = torch.empty((B, 3, 224, 224), device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            pre.preprocess_device(dev, out=out)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        ts = []
        for _ in range(10):
            flush.zero_()                                   # evict the tiles from L2 between iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); pre.preprocess_device(dev, out=out); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            pre.preprocess_device(host.cuda(non_blocking=True), out=out)
        e1.record(); torch.cuda.synchronize()
        ms_e2e = e0.elapsed_time(e1) / 5
        bytes_alg = B * H * W * 3 + out.numel() * 2
        cpu = None
        try:
            import PIL.Image as Image
            from lhrs_bot_b200.preprocess import OPENAI_CLIP_MEAN, OPENAI_CLIP_STD
            t0 = time.perf_counter()
            for b in range(min(B, 8)):
                short, long_ = (W, H) if W <= H else (H, W)
                nl = int(224 * long_ / short)
                nh, nw = (nl, 224) if W <= H else (224, nl)
                r = np.asarray(Image.fromarray(host[b].numpy()).resize((nw, nh), resample=Image.BICUBIC))
                c = r[(nh - 224) // 2:(nh - 224) // 2 + 224, (nw - 224) // 2:(nw - 224) // 2 + 224]
                x = ((c * (1 / 255)).astype(np.float32) - np.array(OPENAI_CLIP_MEAN, np.float32)) / np.array(OPENAI_CLIP_STD, np.float32)
                x.transpose(2, 0, 1).copy()
            cpu = (time.perf_counter() - t0) / min(B, 8) * 1e3
        except ImportError:
            pass
        rows.append(dict(batch=B, H=H, W=W, ms=ms, images_per_s=B / ms * 1e3, algorithmic_GBps=bytes_alg / ms / 1e6,
                         ms_with_h2d=ms_e2e, images_per_s_with_h2d=B / ms_e2e * 1e3, pillow_ms_per_image_1core=cpu))
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
