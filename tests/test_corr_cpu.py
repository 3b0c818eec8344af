"""CPU: the correlation oracle against the golden vectors produced by the reference's own CorrBlock1D
(tests/golden/make_corr_golden.py) -- this is what PINS the corr oracle."""
import os

import numpy as np

from oracle.corr_oracle import CorrOracle

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "corr_golden.npz"))


def test_volume_and_pyramid_match_reference():
    o = CorrOracle("f64")
    pyr = o.pyramid(G["fmap1"], G["fmap2"], 4)
    for i in range(4):
        assert pyr[i].shape == G[f"level{i}"].shape
        assert np.abs(pyr[i] - G[f"level{i}"]).max() < 1e-12


def test_lookup_matches_reference():
    o = CorrOracle("f64")
    pyr = [G[f"level{i}"] for i in range(4)]
    out = o.lookup(pyr, G["coords"], 4)
    assert out.shape == G["out"].shape
    assert np.abs(out - G["out"]).max() < 2e-6          # reference returns .float()


def test_sampler_backward_matches_reference_autograd():
    o = CorrOracle("f64")
    cx = G["coords"][:, 0]
    for i in range(4):
        go = G["grad_out"][:, 9 * i:9 * i + 9]
        gv = o.sample_bwd(G[f"level{i}"].shape, cx / 2 ** i, go, 4)
        assert np.abs(gv - G[f"grad_level{i}"]).max() < 1e-6    # grad passes through the reference's .float() cast


def test_f32_oracle_close():
    o = CorrOracle("f32")
    out = o.lookup(o.pyramid(G["fmap1"], G["fmap2"], 4), G["coords"].astype(np.float32), 4)
    assert np.abs(out - G["out"]).max() < 1e-4          # coords/16 in fp32 -> interpolation weights differ by ~1e-6 * |corr|
