"""Experiment: how the device-resident views/s of bench.py's planned loop depends on stream count and stream priorities
(alternating high/low priority lets one view's latency-bound binning kernels pre-empt another view's pending compositing CTAs)."""
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def run(n_streams, prio, steps=60):
    lo, hi = 0, -1
    if prio == "none":
        side = [torch.cuda.Stream(dev) for _ in range(n_streams)]
    elif prio == "alt":
        side = [torch.cuda.Stream(dev, priority=hi if i % 2 else lo) for i in range(n_streams)]
    elif prio == "ladder":                       # B200 exposes several priority levels: spread them
        side = [torch.cuda.Stream(dev, priority=-(i % 4)) for i in range(n_streams)]

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
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return V * steps / (e0.elapsed_time(e1) * 1e-3)


for n in (2, 4, 8):
    for prio in ("none", "alt", "ladder"):
        print(f"streams={n} prio={prio}: {run(n, prio):.0f} views/s", flush=True)
print("profile events on:")
_lib.profile_enable(True)
print(f"streams=4 prio=none events on: {run(4, 'none'):.0f}")
_lib.profile_read(); _lib.profile_enable(False)
