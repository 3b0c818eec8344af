"""CPU tests of the oracle itself (no GPU): analytic closed forms, autograd cross-check of the
hand-written backward, binning invariants.  These are what pin the rasterizer oracle, since the
reference ships no golden vectors for this path."""
import math

import numpy as np
import pytest
import torch

from gps_gaussian_b200 import synth
from helpers import oracle_forward, rel_err
from oracle.raster_oracle import RasterOracle, taichi_splat


def _single_gaussian_scene(res=64, sigma_px=2.0, opacity=0.8, bg=(0.1, 0.2, 0.3)):
    K0, E0 = synth.ring_camera(0.0, res)
    cam = synth.novel_camera(K0, E0, K0, E0, res, res, 0.5)
    R, t = E0[:, :3], E0[:, 3]
    z = 2.0
    pc = np.array([0.0, 0.0, z])                       # ON the optical axis => J02 = J12 = 0, Sigma2D exactly isotropic
    xyz = (R.T @ (pc - t)).astype(np.float32)[None]
    s = np.float32(sigma_px * z / K0[0, 0])
    return dict(means3D=xyz, colors=np.array([[1.0, 0.5, 0.25]], np.float32), opacity=np.array([[opacity]], np.float32),
                scales=np.full((1, 3), s, np.float32), rots=np.array([[1, 0, 0, 0]], np.float32),
                view=cam["world_view_transform"], proj=cam["full_proj_transform"], campos=cam["camera_center"],
                tanfovx=math.tan(cam["FovX"] * 0.5), tanfovy=math.tan(cam["FovY"] * 0.5), W=res, H=res,
                bg=np.asarray(bg, np.float32)), sigma_px


def test_single_isotropic_gaussian_closed_form():
    """alpha(r) = min(0.99, o*exp(-r^2/(2(sigma^2+0.3)))), C = c*alpha + (1-alpha)*bg  (SURVEY 8c-iv)."""
    sc, sig = _single_gaussian_scene()
    _, st = oracle_forward(sc, "f64")
    assert st["n_visible"] == 1
    cx, cy = st["means2D"][0]
    # principal point (cx, cy) in continuous coords == pixel INDEX (cx-0.5, cy-0.5): ndc2pix folds the -0.5
    assert abs(cx - (32.0 - 0.5)) < 1e-3 and abs(cy - (32.0 + 25.0 * 64 / 1024 - 0.5)) < 1e-3
    var = sig * sig + 0.3
    assert st["radii"][0] == math.ceil(3 * math.sqrt(var))
    ys, xs = np.mgrid[0:64, 0:64]
    r2 = (xs - cx) ** 2 + (ys - cy) ** 2
    alpha = np.minimum(0.99, 0.8 * np.exp(-r2 / (2 * var)))
    alpha[alpha < 1 / 255] = 0
    # only tiles in the splat's rect see it
    rect = st["rects"][0]
    mask = np.zeros((64, 64), bool)
    mask[rect[1] * 16:rect[3] * 16, rect[0] * 16:rect[2] * 16] = True
    alpha = np.where(mask, alpha, 0)
    for ch, c in enumerate((1.0, 0.5, 0.25)):
        want = c * alpha + (1 - alpha) * sc["bg"][ch]
        assert np.abs(st["color"][ch] - want).max() < 2e-4       # det-normalised conic vs exact var: ~1e-5


def test_front_to_back_order_and_bg():
    sc, _ = _single_gaussian_scene(opacity=0.99)
    # second, farther Gaussian of another colour exactly behind the first
    far = sc["means3D"][0] + (sc["means3D"][0] - sc["campos"]) * 0.25
    sc["means3D"] = np.stack([far, sc["means3D"][0]]).astype(np.float32)        # far one FIRST in index order
    sc["colors"] = np.array([[0, 1, 0], [1, 0, 0]], np.float32)
    sc["opacity"] = np.array([[0.99], [0.99]], np.float32)
    sc["scales"] = np.repeat(sc["scales"], 2, 0); sc["rots"] = np.repeat(sc["rots"], 2, 0)
    _, st = oracle_forward(sc, "f64")
    tile = (33 // 16) * 4 + 31 // 16
    s, e = st["ranges"][tile]
    assert list(st["vals"][s:e]) == [1, 0]                        # near Gaussian first although it has the larger index
    c = st["color"][:, 33, 31]
    assert c[0] > 0.9 and c[1] < 0.08                            # near (red) dominates


@pytest.mark.parametrize("res,P,spread,mul", [(64, 600, 0.45, 3.0), (48, 300, 0.3, 6.0)])
def test_backward_matches_fp64_autograd(res, P, spread, mul):
    from oracle.raster_torch64 import render_autograd
    sc = synth.random_cube_scene(P, res, spread=spread, scale_mul=mul, bg=(0.2, 0.5, 0.7), seed=7)
    o, st = oracle_forward(sc, "f64")
    T = lambda a: torch.tensor(np.asarray(a, np.float64), requires_grad=True)
    m, c, op, s, r = T(sc["means3D"]), T(sc["colors"]), T(sc["opacity"]), T(sc["scales"]), T(sc["rots"])
    img = render_autograd(st, m, c, op, s, r)
    assert np.abs(img.detach().numpy() - st["color"]).max() < 1e-12
    g = np.random.default_rng(0).standard_normal(st["color"].shape)
    (img * torch.tensor(g)).sum().backward()
    gr = o.backward(st, g)
    for name, a, b in (("means3D", m.grad, gr["dL_dmeans3D"]), ("colors", c.grad, gr["dL_dcolors"]),
                       ("opacity", op.grad.reshape(-1), gr["dL_dopacity"]), ("scales", s.grad, gr["dL_dscales"]),
                       ("rots", r.grad, gr["dL_drots"])):
        assert rel_err(b, a.numpy()) < 1e-6, name                 # 1e-7 eps in 1/(denom^2+1e-7) is the floor


def test_backward_matches_fp64_autograd_with_independent_binning():
    """Same cross-check, but the autograd forward takes its discrete state (tile lists, ranges) from the independent numpy
    restatement instead of gpsg_oracle.c: the truth the hand-written A.6-A.8 backward is held to then shares nothing with the
    C oracle (VERDICT r1 missing #6)."""
    from oracle import raster_independent as ri
    from oracle.raster_torch64 import render_autograd
    sc = synth.random_cube_scene(500, 56, spread=0.45, scale_mul=3.0, bg=(0.2, 0.5, 0.7), seed=9)
    ind = ri.forward_scene(sc)
    st_ind = dict(W=sc["W"], H=sc["H"], ranges=ind["ranges"], vals=ind["point_list"],
                  inputs=dict(view=sc["view"], proj=sc["proj"], tanfovx=sc["tanfovx"], tanfovy=sc["tanfovy"], bg=sc["bg"]))
    T = lambda a: torch.tensor(np.asarray(a, np.float64), requires_grad=True)
    m, c, op, s, r = T(sc["means3D"]), T(sc["colors"]), T(sc["opacity"]), T(sc["scales"]), T(sc["rots"])
    img = render_autograd(st_ind, m, c, op, s, r)
    assert np.abs(img.detach().numpy() - ind["color"]).max() < 1e-12          # torch forward == numpy forward
    g = np.random.default_rng(2).standard_normal(ind["color"].shape)
    (img * torch.tensor(g)).sum().backward()
    o, st = oracle_forward(sc, "f64")
    assert np.array_equal(st["vals"], ind["point_list"]) and np.abs(st["color"] - ind["color"]).max() < 1e-12
    gr = o.backward(st, g)
    for name, a, b in (("means3D", m.grad, gr["dL_dmeans3D"]), ("colors", c.grad, gr["dL_dcolors"]),
                       ("opacity", op.grad.reshape(-1), gr["dL_dopacity"]), ("scales", s.grad, gr["dL_dscales"]),
                       ("rots", r.grad, gr["dL_drots"])):
        assert rel_err(b, a.numpy()) < 1e-6, name


def test_f32_and_f64_oracles_agree():
    sc = synth.random_cube_scene(3000, 128, seed=3)
    _, a = oracle_forward(sc, "f32")
    _, b = oracle_forward(sc, "f64")
    # 1e-4 everywhere except at pixels where a hard threshold (alpha<1/255, T<1e-4, power>0) flips between
    # precisions; such a flip moves a pixel by at most ~1/255 and is inherent to the algorithm
    d = np.abs(a["color"] - b["color"]).max(0)
    assert (d > 1e-4).mean() < 5e-4 and d.max() < 1e-2
    assert (a["radii"] != b["radii"]).mean() < 1e-3


def test_threshold_margins_explain_every_f32_f64_compositing_difference():
    """The flip allowance as a theorem, exercised on the CPU with the fp64 oracle standing in for the
    device: compositing the SAME per-Gaussian state in fp32 and in fp64, every pixel that is NOT within `RasterOracle.EPS`
    of a hard threshold (alpha vs 1/255, test_T vs 1e-4, power vs 0) agrees to 2e-6 with identical n_contrib, so any pixel
    over 1e-4 must be a near-threshold one; likewise every gradient of a Gaussian no near-threshold pixel evaluates."""
    from helpers import assert_grad_parity, assert_image_parity
    sc = synth.random_cube_scene(40_000, 256, spread=0.5, scale_mul=2.0, seed=21)
    o32, a = oracle_forward(sc, "f32")
    b = RasterOracle("f64").render_state(a, nthreads=8)
    rec = assert_image_parity("cpu_selfcheck", b["color"], b["final_T"], b["n_contrib"], o32, a)
    assert 0 < rec["near"] < 0.02 * rec["pixels"] and rec["max_err_not_near"] <= 2e-6
    # gradients: "device" = the fp32 oracle's own backward; checked against fp32/fp64 backward forced onto its decisions
    g = np.random.default_rng(1).standard_normal((3, 256, 256)).astype(np.float32)
    got = o32.backward(a, g)
    got = {"dL_dmeans3D": got["dL_dmeans3D"], "dL_dcolors": got["dL_dcolors"], "dL_dopacity": got["dL_dopacity"],
           "dL_dscales": got["dL_dscales"], "dL_drots": got["dL_drots"], "dL_dmeans2D": got["dL_dmean2D"]}
    rep = assert_grad_parity("cpu_selfcheck", sc, got, a, a["final_T"], a["n_contrib"], g)
    r32 = rep[("f32", "dL_dmeans3D")]
    assert r32["max_err_clean"] == r32["max_err_shared"] == r32["max_err_own"] == 0.0     # same code, same decisions
    assert rep[("f64", "dL_dmeans3D")]["own"] < 0.1 * 40_000


@pytest.mark.parametrize("P,res,kw", [(3000, 128, dict(seed=3)), (4000, 250, dict(spread=0.6, scale_mul=4.0, bg=(0.3, 0.6, 0.9), seed=11)),
                                      (2000, 130, dict(spread=3.0, seed=11)), (10_000, 256, dict())])
def test_independent_numpy_restatement_agrees_with_the_c_oracle(P, res, kw):
    """oracle/raster_independent.py takes only the raw call arguments (no state of gpsg_oracle.c): culling, radii, tile
    counts, the sorted 64-bit keys, point list and tile ranges must be IDENTICAL to the C oracle's fp64 build, the image,
    final_T to 1e-12 and n_contrib identical (VERDICT r1 missing #6)."""
    from oracle import raster_independent as ri
    sc = synth.random_cube_scene(P, res, **kw)
    _, a = oracle_forward(sc, "f64")
    b = ri.forward_scene(sc)
    assert np.array_equal(a["radii"], b["radii"]) and np.array_equal(a["tiles_touched"], b["tiles_touched"])
    assert a["num_rendered"] == b["num_rendered"] > 0
    assert np.array_equal(a["keys"], b["keys"]) and np.array_equal(a["vals"], b["point_list"])
    assert np.array_equal(a["ranges"], b["ranges"])
    assert np.abs(a["means2D"] - b["means2D"]).max() < 1e-9 and np.abs(a["conic_opacity"][:, :3] - b["conic"]).max() < 1e-9
    assert np.abs(a["color"] - b["color"]).max() < 1e-12 and np.abs(a["final_T"] - b["final_T"]).max() < 1e-12
    assert np.array_equal(a["n_contrib"], b["n_contrib"])


def test_non_finite_inputs_are_culled_in_both_restatements():
    """NaN / inf scales, rotations, positions: culled (radius 0 is not a splat), so sum(tiles_touched) == number of emitted
    pairs -- the invariant whose violation crashed the tile sort on the device (tests/test_raster_gpu.py has the GPU half)."""
    from oracle import raster_independent as ri
    sc = synth.random_cube_scene(3000, 128, seed=3)
    sc["scales"][::7] = np.nan; sc["rots"][3::11] = np.nan; sc["means3D"][1::13] = np.nan
    sc["means3D"][5::17, 0] = np.inf; sc["scales"][2::23, 1] = np.inf
    for dt in ("f32", "f64"):
        _, a = oracle_forward(sc, dt)
        assert a["num_rendered"] == int(a["tiles_touched"].sum()) == int((a["ranges"][:, 1].astype(np.int64) - a["ranges"][:, 0]).sum())
        assert np.all((a["radii"] > 0) == (a["tiles_touched"] > 0)) and np.isfinite(a["color"]).all()
    with np.errstate(all="ignore"):
        b = ri.forward_scene(sc)
    assert np.array_equal(a["radii"], b["radii"]) and np.array_equal(a["vals"], b["point_list"])
    assert np.abs(a["color"] - b["color"]).max() < 1e-12


def test_independent_restatement_randomised_sweep():
    """Seeded sweep over image size, point count, spread, splat size, background: the two restatements (C, scalar chains;
    numpy, matrix form + global argsort) must agree on every integer output and on the image to 1e-11."""
    from oracle import raster_independent as ri
    rng = np.random.default_rng(77)
    for k in range(10):
        res = int(rng.integers(24, 220))
        P = int(rng.integers(1, 2500))
        sc = synth.random_cube_scene(P, res, spread=float(rng.uniform(0.2, 2.5)), scale_mul=float(rng.uniform(0.5, 8.0)),
                                     bg=tuple(rng.uniform(0, 1, 3)), seed=int(rng.integers(1 << 30)))
        _, a = oracle_forward(sc, "f64")
        b = ri.forward_scene(sc)
        assert np.array_equal(a["radii"], b["radii"]) and np.array_equal(a["tiles_touched"], b["tiles_touched"]), k
        assert np.array_equal(a["keys"], b["keys"]) and np.array_equal(a["vals"], b["point_list"]), k
        assert np.array_equal(a["ranges"], b["ranges"]) and np.array_equal(a["n_contrib"], b["n_contrib"]), k
        assert np.abs(a["color"] - b["color"]).max() < 1e-11 and np.abs(a["final_T"] - b["final_T"]).max() < 1e-11, k


def test_binning_invariants():
    sc = synth.random_cube_scene(5000, 200, seed=5)               # 200 is not a multiple of 16
    _, st = oracle_forward(sc, "f32")
    keys = st["keys"]
    assert np.all(keys[1:] >= keys[:-1])                          # sortedness
    assert st["num_rendered"] == int(st["tiles_touched"].sum())
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    gx = (200 + 15) // 16
    for t in np.unique(tiles):
        s, e = st["ranges"][t]
        assert np.all(tiles[s:e] == t) and (s == 0 or tiles[s - 1] != t) and (e == len(tiles) or tiles[e] != t)
    # every pair's tile lies inside its Gaussian's rect
    r = st["rects"][st["vals"]]
    assert np.all((tiles % gx >= r[:, 0]) & (tiles % gx < r[:, 2]) & (tiles // gx >= r[:, 1]) & (tiles // gx < r[:, 3]))
    # equal (tile, depth) keys keep Gaussian-index order (stable sort)
    same = keys[1:] == keys[:-1]
    assert np.all(st["vals"][1:][same] > st["vals"][:-1][same])


def test_empty_and_culled():
    sc = synth.random_cube_scene(50, 64, seed=1)
    sc["means3D"] = (sc["means3D"] + (sc["campos"] - np.array([0, 0.85, 0], np.float32)) * 3).astype(np.float32)  # behind cam
    _, st = oracle_forward(sc, "f32")
    assert st["n_visible"] == 0 and st["num_rendered"] == 0
    assert np.allclose(st["color"], sc["bg"][:, None, None]) and np.all(st["final_T"] == 1)
    o = RasterOracle("f32")
    assert not o.mark_visible(sc["means3D"], sc["view"]).any()


def test_taichi_splat_restatement():
    """Sequential semantics of reference lib/TaichiRender.py:12-23 (z-buffer on inverse depth)."""
    pts = np.array([[[3.2, 4.9, 0.5, 1, 0, 0], [3.7, 4.1, 0.8, 0, 1, 0], [3.0, 4.0, 0.6, 0, 0, 1],
                     [-5.0, 99.0, 0.1, 1, 1, 1], [1.0, 1.0, 9.0, 9, 9, 9]]], np.float32)
    mask = np.array([[1, 1, 1, 1, 0]], np.float32)
    depth, color = taichi_splat(pts, mask, 8)
    assert depth[0, 0, 4, 3] == np.float32(0.8) and list(color[0, :, 4, 3]) == [0, 1, 0]      # nearest (largest 1/z) wins
    assert depth[0, 0, 7, 0] == np.float32(0.1) and list(color[0, :, 7, 0]) == [1, 1, 1]      # clamped to the border
    assert color[0, 0, 1, 1] == -1                                                           # masked point ignored
