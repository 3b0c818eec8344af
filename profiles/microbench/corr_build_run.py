"""Runs the fp16 correlation build (tcgen05 kernel) a few times at the C2 input size -- target for `ncu -k regex:corr_build`."""
import sys
import torch
sys.path.insert(0, "/root/repo")
from gps_gaussian_b200.corr import CorrBlockFast1D
g = torch.Generator("cuda").manual_seed(0)
f1 = torch.randn(2, 192, 128, 128, device="cuda", generator=g).half()
f2 = torch.randn(2, 192, 128, 128, device="cuda", generator=g).half()
for _ in range(3):
    blk = CorrBlockFast1D(f1, f2)
torch.cuda.synchronize()
print("ok", float(blk.corr_pyramid[0].float().abs().max()))
