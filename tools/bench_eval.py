#!/usr/bin/env python
"""Times the device evaluation metrics (lgs_b200.eval_metrics) on BASELINE's 64 x 2048 range image: point-cloud
conversion, nearest-neighbour search in range-image order (pruning active) and on shuffled clouds (tiled brute force),
and the whole PointsMeter.update; prints one JSON line.  Reference numbers: profiles/r01_ref_chamfer_cuda_timing.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lidar-gs_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from lgs_b200 import eval_metrics as M  # noqa: E402
from make_goldens_eval import beams_of, range_image  # noqa: E402  (input generator only)


def timed(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    dev = "cuda:0"
    H, W = 64, 2048
    gt = torch.from_numpy(range_image(H, W, 99, drop=0.0)).to(dev)
    pred = torch.from_numpy(range_image(H, W, 99, drop=0.0, noise=0.05).astype(np.float32)).to(dev)
    beams = beams_of(H)
    a = M.pano_to_lidar(pred, beam_inclinations=beams)[None].contiguous()
    b = M.pano_to_lidar(gt, beam_inclinations=beams)[None].contiguous()
    out = dict(n=int(a.shape[1]), m=int(b.shape[1]))
    out["pano_to_lidar_ms"] = timed(lambda: M.pano_to_lidar(pred, beam_inclinations=beams))
    stats = torch.zeros(3, dtype=torch.int64, device=dev)
    M.nn_distance(a, b, stats=stats)
    ev, tot, ld = (int(v) for v in stats.cpu())
    out["range_images"] = dict(nn_ms=timed(lambda: M.nn_distance(a, b)), warp_tiles_evaluated=ev, warp_tiles_total=tot,
                               cta_tiles_loaded=ld)
    sa = a[:, torch.randperm(a.shape[1], device=dev)].contiguous()
    sb = b[:, torch.randperm(b.shape[1], device=dev)].contiguous()
    stats.zero_()
    M.nn_distance(sa, sb, stats=stats)
    ev, tot, ld = (int(v) for v in stats.cpu())
    ms = timed(lambda: M.nn_distance(sa, sb), n=5)
    out["shuffled"] = dict(nn_ms=ms, warp_tiles_evaluated=ev, warp_tiles_total=tot,
                           pairs_per_s=2.0 * a.shape[1] * b.shape[1] / (ms * 1e-3))
    meter = M.PointsMeter(scale=1, intrinsics=None, beam_inclinations=beams)
    out["points_meter_update_ms"] = timed(lambda: meter.update(pred[None], gt[None]))
    out["cd_fscore"] = [float(v) for v in meter.measure()]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
