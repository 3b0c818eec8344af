"""Experiment: tile_sort_gather variants (GPSG_SORT_RBITS = radix digit width, GPSG_SORT_GB = gather batch) on the C2 workload."""
import os
import sys
import torch
sys.path.insert(0, "/root/repo")
from gps_gaussian_b200 import synth, _lib
from gps_gaussian_b200.introspect import RasterCall, to_device, make_settings
from gps_gaussian_b200.planned import PlannedRasterizer

dev = torch.device("cuda", 0)
V, RES = 8, 1024
scenes = [synth.stereo_pair_scene(RES, seed=1314 + k) for k in range(V)]
calls = [RasterCall(sc, to_device(sc, dev), dev) for sc in scenes]
for c in calls:
    c.forward()
torch.cuda.synchronize()
planned = [PlannedRasterizer(c.P, RES, RES, int(c.num_rendered * 1.25) + 1024, dev) for c in calls]
pargs = [(make_settings(c.sc), c.inp["means3D"], c.inp["colors"], c.inp["opacity"], c.inp["scales"], c.inp["rots"]) for c in calls]
side = [torch.cuda.Stream(dev) for _ in range(8)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rb, gb in (("4", "1"), ("4", "2"), ("4", "4"), ("5", "4"), ("6", "4")):
    os.environ["GPSG_SORT_RBITS"], os.environ["GPSG_SORT_GB"] = rb, gb
    for pr, a in zip(planned, pargs):
        pr.forward(*a)
    torch.cuda.synchronize()
    same = all(torch.equal(pr.color, c.color) for pr, c in zip(planned, calls))
    _lib.profile_enable(True)
    for _ in range(10):
        for pr, a in zip(planned, pargs):
            pr.forward(*a)
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    kms = prof["tile_sort_gather"]["ms"] / prof["tile_sort_gather"]["calls"]

    def step():
        main = torch.cuda.current_stream(dev)
        for st in side:
            st.wait_stream(main)
        for k, (pr, a) in enumerate(zip(planned, pargs)):
            with torch.cuda.stream(side[k % len(side)]):
                pr.forward(*a)
        for st in side:
            main.wait_stream(st)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(60):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"rbits {rb} gb {gb}: tile_sort_gather {kms * 1e3:.1f} us, {V * 60 / (e0.elapsed_time(e1) * 1e-3):.0f} views/s, identical={same}", flush=True)
