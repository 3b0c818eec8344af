"""Direct C-ABI driver that also exposes the saved rasterizer state (depths, means2D, conics,
tiles_touched, sorted keys / point list, tile ranges, final_T, n_contrib) as torch tensors.
Used by the parity tests ("tile indices bit-exact") and by bench.py; not on the training path."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _sub(buf, ptr, nbytes, dtype):
    off = int(ptr) - buf.data_ptr()
    assert 0 <= off and off + nbytes <= buf.numel(), (off, nbytes, buf.numel())
    return buf[off:off + nbytes].view(dtype)


def make_settings(sc, debug=False):
    s = _lib.RasterSettings()
    s.image_height, s.image_width = int(sc["H"]), int(sc["W"])
    s.tanfovx, s.tanfovy = float(sc["tanfovx"]), float(sc["tanfovy"])
    s.bg[:] = [float(v) for v in np.asarray(sc["bg"]).reshape(3)]
    s.scale_modifier = float(sc.get("scale_modifier", 1.0))
    s.viewmatrix[:] = [float(v) for v in np.asarray(sc["view"], np.float32).reshape(16)]
    s.projmatrix[:] = [float(v) for v in np.asarray(sc["proj"], np.float32).reshape(16)]
    s.sh_degree = 3
    s.campos[:] = [float(v) for v in np.asarray(sc["campos"], np.float32).reshape(3)]
    s.prefiltered, s.debug = 0, int(debug)
    return s


def to_device(sc, device="cuda"):
    """numpy scene dict (synth.py) -> dict of contiguous fp32 CUDA tensors for the five per-Gaussian inputs."""
    t = lambda k: torch.from_numpy(np.ascontiguousarray(sc[k], np.float32)).to(device) if sc.get(k) is not None else None
    return dict(means3D=t("means3D"), colors=t("colors"), opacity=t("opacity"), scales=t("scales"), rots=t("rots"),
                cov3D_precomp=t("cov3D_precomp"))


class RasterCall:
    """One forward (and optionally backward) through gpsg_rasterize_* with raw tensors."""

    def __init__(self, sc, dev_inputs=None, device="cuda"):
        self.sc = sc
        self.device = torch.device(device)
        self.inp = dev_inputs if dev_inputs is not None else to_device(sc, device)
        self.settings = make_settings(sc)
        self.P = int(self.inp["means3D"].shape[0])
        self.H, self.W = int(sc["H"]), int(sc["W"])
        self.color = torch.empty((3, self.H, self.W), dtype=torch.float32, device=self.device)
        self.radii = torch.empty((self.P,), dtype=torch.int32, device=self.device)
        self.num_rendered = 0
        self.bufs = None

    def _p(self, t):
        return C.c_void_p(t.data_ptr()) if (t is not None and t.numel() > 0) else None

    def forward(self):
        i = self.inp
        n = C.c_int32(0)
        dev = self.device
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        _lib.begin_alloc(dev)
        try:
            rc = _lib.lib.gpsg_rasterize_forward(
                C.byref(self.settings), idx, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), self.P, 0,
                self._p(i["means3D"]), self._p(i["colors"]), None, self._p(i["opacity"]), self._p(i["scales"]),
                self._p(i["rots"]), self._p(i.get("cov3D_precomp")), self._p(self.color), self._p(self.radii),
                _lib.ALLOC_CB, C.c_void_p(1), _lib.ALLOC_CB, C.c_void_p(2), _lib.ALLOC_CB, C.c_void_p(3), C.byref(n))
        finally:
            bufs = _lib.end_alloc()
        _lib.check(rc, "gpsg_rasterize_forward")
        self.bufs = (bufs.get(1), bufs.get(2), bufs.get(3))
        self.num_rendered = int(n.value)
        return self.color

    def backward(self, grad_color, want_cov3D=False):
        i = self.inp
        dev = self.device
        P = self.P
        new = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        out = dict(dL_dmeans2D=new(P, 3), dL_dcolors=new(P, 3), dL_dopacity=new(P, 1), dL_dmeans3D=new(P, 3),
                   dL_dscales=new(P, 3), dL_drots=new(P, 4), dL_dcov3D=new(P, 6) if want_cov3D else None)
        ws = torch.empty(int(_lib.lib.gpsg_rasterize_backward_workspace_bytes(P)), dtype=torch.uint8, device=dev)
        g = grad_color.to(torch.float32).contiguous()
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        rc = _lib.lib.gpsg_rasterize_backward(
            C.byref(self.settings), idx, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), P, 0,
            self.num_rendered, self._p(i["means3D"]), self._p(i["colors"]), None, self._p(i["opacity"]),
            self._p(i["scales"]), self._p(i["rots"]), self._p(i.get("cov3D_precomp")), self._p(self.radii),
            self._p(self.bufs[0]), self._p(self.bufs[1]), self._p(self.bufs[2]), self._p(g),
            self._p(out["dL_dmeans2D"]), self._p(out["dL_dcolors"]), self._p(out["dL_dopacity"]),
            self._p(out["dL_dmeans3D"]), self._p(out["dL_dcov3D"]), None, self._p(out["dL_dscales"]),
            self._p(out["dL_drots"]), self._p(ws))
        _lib.check(rc, "gpsg_rasterize_backward")
        self._ws = ws
        return out

    def state(self):
        """Saved buffers as torch tensors (views into the scratch buffers)."""
        geom, binning, image = self.bufs
        P, N, H, W = self.P, self.num_rendered, self.H, self.W
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        st = dict(radii=self.radii, num_rendered=N)
        gv, bv, iv = _lib.GeomView(), _lib.BinningView(), _lib.ImageView()
        _lib.check(_lib.lib.gpsg_geom_view(C.c_void_p(geom.data_ptr()), P, C.byref(gv)), "gpsg_geom_view")
        _lib.check(_lib.lib.gpsg_image_view(C.c_void_p(image.data_ptr()), W, H, C.byref(iv)), "gpsg_image_view")
        if P > 0:
            st["depths"] = _sub(geom, gv.depths, 4 * P, torch.float32)
            st["means2D"] = _sub(geom, gv.means2D, 8 * P, torch.float32).view(P, 2)
            st["conic_opacity"] = _sub(geom, gv.conic_opacity, 16 * P, torch.float32).view(P, 4)
            st["tiles_touched"] = _sub(geom, gv.tiles_touched, 4 * P, torch.int32)
            st["point_offsets"] = _sub(geom, gv.point_offsets, 4 * P, torch.int32)
        if N > 0:
            _lib.check(_lib.lib.gpsg_binning_view(C.c_void_p(binning.data_ptr()), N, C.byref(bv)), "gpsg_binning_view")
            st["keys"] = _sub(binning, bv.point_list_keys, 8 * N, torch.int64)
            st["point_list"] = _sub(binning, bv.point_list, 4 * N, torch.int32)
        st["final_T"] = _sub(image, iv.final_T, 4 * H * W, torch.float32).view(H, W)
        st["n_contrib"] = _sub(image, iv.n_contrib, 4 * H * W, torch.int32).view(H, W)
        st["ranges"] = _sub(image, iv.ranges, 8 * tiles, torch.int32).view(tiles, 2)
        return st
