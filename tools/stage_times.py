#!/usr/bin/env python
"""Per-kernel times of the rasterizer forward + backward on the C2 scene (CUDA events around each launch, serialised on
one stream) -- the quick loop used while tuning kernels:  python tools/stage_times.py [--reps 50] [--seeds 1314 1315]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gps_gaussian_b200 import _lib, synth  # noqa: E402
from gps_gaussian_b200.introspect import RasterCall, to_device  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=50)
ap.add_argument("--res", type=int, default=1024)
ap.add_argument("--seeds", type=int, nargs="*", default=[1314, 1315, 1316, 1317])
ap.add_argument("--no-backward", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
calls = [RasterCall(sc, to_device(sc, dev), dev) for sc in (synth.stereo_pair_scene(a.res, seed=s) for s in a.seeds)]
g = torch.randn(3, a.res, a.res, device=dev)
for c in calls:
    c.forward()
    if not a.no_backward:
        c.backward(g)
torch.cuda.synchronize()
_lib.profile_enable(True)
_lib.profile_read()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    for c in calls:
        c.forward()
        if not a.no_backward:
            c.backward(g)
e1.record()
torch.cuda.synchronize()
prof = _lib.profile_read()
_lib.profile_enable(False)
out = {k: round(v["ms"] / max(v["calls"], 1) * 1e3, 1) for k, v in prof.items() if v["calls"]}
out["total_us_per_view"] = round(e0.elapsed_time(e1) * 1e3 / (a.reps * len(calls)), 1)
out["N_dup_mean"] = sum(c.num_rendered for c in calls) / len(calls)
print(json.dumps(out))
