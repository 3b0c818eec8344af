"""Debug aid: run the C5-size stage-2 step until the rasterizer fails, keeping the inputs of the last render call;
on failure they are written to the given path for tools/repro_c5_state.py replay."""
import os, sys, math, tempfile
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
path = sys.argv[1]
res = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
from gps_gaussian_b200 import harness, synth_dataset
root = tempfile.mkdtemp(prefix="gpsg_dbg_")
synth_dataset.write_dataset(root, n_train=2, n_val=1, res=res, hr=True)
cfg = harness.load_cfg(root, src_res=res, num_steps=1000, batch_size=2)
st = harness.C3State(cfg)
import diff_gaussian_rasterization as dgr
last = {}
orig = dgr._RasterizeGaussians.forward
def spy(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
    last.clear()
    last.update(xyz=means3D.detach().float().cpu(), rgb=colors_precomp.detach().float().cpu() * 2 - 1, op=opacities.detach().float().cpu(),
                scale=scales.detach().float().cpu(), rot=rotations.detach().float().cpu(), H=rs.image_height, W=rs.image_width,
                FovX=2 * math.atan(rs.tanfovx), FovY=2 * math.atan(rs.tanfovy), view=rs.viewmatrix.cpu(), proj=rs.projmatrix.cpu(),
                campos=rs.campos.cpu())
    r = orig(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs)
    print("   render: P", int(means3D.shape[0]), "N", ctx.num_rendered, "max radius", int(r[1].max()), "mem GiB %.1f" % (torch.cuda.memory_allocated() / 2**30), flush=True)
    return r
dgr._RasterizeGaussians.forward = staticmethod(spy)
batches = [st.batch(k) for k in (0, 2)]
copy = lambda d: {k: (dict(v) if isinstance(v, dict) else v) for k, v in d.items()}
try:
    for k in range(40):
        out = harness.c3_step(st, copy(batches[k % 2]))
        torch.cuda.synchronize()
        o = last
        import gc
        a0 = torch.cuda.memory_allocated() / 2**30
        if os.environ.get("GPSG_DBG_GC"):
            gc.collect()
        print(k, f"mem {a0:.1f} GiB -> {torch.cuda.memory_allocated() / 2**30:.1f} GiB (peak {torch.cuda.max_memory_allocated() / 2**30:.1f})", "loss", float(out["loss"]), "scale", out["scale_after"], "P", o["xyz"].shape[0], "finite", bool(torch.isfinite(o["xyz"]).all()),
              "|xyz|max", float(o["xyz"].abs().max()), "scale max", float(o["scale"].max()), flush=True)
except Exception as e:
    print("FAILED at step", k, repr(e)[:300], flush=True)
    torch.save([dict(last)], path)
    print("saved", path, "P", last["xyz"].shape[0], "finite", bool(torch.isfinite(last["xyz"]).all()), "|xyz|max", float(last["xyz"].abs().max()))
    sys.exit(3)
