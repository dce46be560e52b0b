"""Normals-coded infill mask of `stereo_rerender.py --infill_mask` (:583-606,727-819,838-907): the mask video the
infill engines consume (movie_2_3D passes --infill_mask by default, movie_2_3D.py:440).

Per frame, on the GPU: the stereo render itself (row kernel or generic path, unchanged), the edge test of the mesh
builder (E1), and per eye the edge-point splat (E2) and the mask / image painting (E3).  On the host, exactly as the
reference does it with OpenCV: TELEA inpainting of everything that is not a coded normal, copied into the hole pixels,
then the black-ignoring Gaussian -- fanned over a thread pool (cv2 releases the GIL), one eye of one frame per task.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Optional

import numpy as np
import torch

from . import geometry as geo
from . import ops
from .stereo import StereoRerenderer

GREEN = (0, 255, 0)


def masked_blur(img: np.ndarray, ksize=(6, 6), sigma=0) -> np.ndarray:
    """stereo_rerender.masked_blur (:114-153): Gaussian blur in which pure-black pixels carry no weight."""
    import cv2

    g = cv2.getGaussianKernel(ksize[0], sigma)
    kernel = g @ g.T
    black = np.all(img == 0, axis=2)
    weights = cv2.filter2D((~black).astype(np.float32), -1, kernel, borderType=cv2.BORDER_ISOLATED)
    total = cv2.filter2D(img.astype(np.float32), -1, kernel, borderType=cv2.BORDER_ISOLATED)
    out = total / np.where(weights == 0, 1.0, weights)[..., None]
    out[weights == 0] = 0
    out[black] = 0
    return np.clip(out, 0, 255).astype(np.uint8)


def infill_using_normals(color_img, hole_mask, normal_map, max_steps=400):
    """stereo_rerender.infill_using_normals (:155-240; imported by basic_nomal_infill.py:10 and
    stereo_dissoclusion_net_infill.py:10): every hole pixel marches along the x, y direction its normal codes until it
    leaves the hole and copies the colour it finds there.  color_img (H, W, 3) u8, hole_mask (H, W) bool / u8,
    normal_map (H, W, 3) float in [-1, 1].  NumPy in -> NumPy out, CUDA tensors in -> CUDA tensor out; the input
    image is not modified (the reference works on a copy)."""
    from .depth_frames_helper import _down, _up

    img, as_np = _up(color_img, torch.uint8)
    if not as_np:
        img = img.clone()
    hole = hole_mask.to(torch.uint8) if isinstance(hole_mask, torch.Tensor) else np.asarray(hole_mask).astype(np.uint8)
    hole, _ = _up(hole, torch.uint8)
    normals, _ = _up(normal_map, torch.float32)
    return _down(ops.normal_march_infill_f32(img, hole, normals.contiguous(), max_steps), as_np)


def make_infill_mask(boolean_mask, normals):
    """stereo_rerender.make_infill_mask (:89-91): a placeholder that returns None in the reference, kept importable."""
    return None


def telea_fill_holes(mask_u8: np.ndarray, green: np.ndarray, area: np.ndarray) -> np.ndarray:
    """`cv2.inpaint(mask, area, 3, INPAINT_TELEA)` as far as the GREEN (hole) pixels are concerned -- the only pixels of
    the result the reference keeps (stereo_rerender.py:805-807).  The reference's inpaint area is everything that is not
    a coded normal, i.e. ~99.9 % of the image (2.4 s per 1080p eye), although the fast-marching front reaches the hole
    pixels first: they sit next to the normals painted into them.  Pixels the front would reach only after every hole
    pixel is done cannot influence a hole pixel, so they are left out of the area.  Leaving them out turns them into
    known (black) pixels that start a front of their own; with D = the largest distance of a hole pixel from a coded
    normal, that front needs more than D + 3 marching time to come within the 3-pixel inpaint radius of a hole pixel as
    long as the area reaches 2 D + 8 pixels beyond the holes (a pixel influenced by both fronts before time D would be
    within D of a normal AND within D of the cut).  With R = ceil(2.2 D) + 10 the hole pixels come out bit-identical
    (tests: random masks against the full area, and the golden images produced by the reference's own lines; margins
    below ~1.5 D do differ).  MDVT_TELEA_FULL=1 selects the reference's area."""
    import os

    import cv2

    if not green.any():
        return mask_u8          # nothing of the inpainted image would be kept
    known = ~area
    if os.environ.get("MDVT_TELEA_FULL") == "1" or not known.any():
        sub = area
    else:
        dist_known = cv2.distanceTransform(area.astype(np.uint8), cv2.DIST_L2, cv2.DIST_MASK_PRECISE)   # distance to a coded normal
        reach = int(np.ceil(2.2 * float(dist_known[green].max()))) + 10
        near_holes = cv2.distanceTransform((~green).astype(np.uint8), cv2.DIST_L2, cv2.DIST_MASK_PRECISE) <= reach
        sub = area & near_holes
    return cv2.inpaint(mask_u8, sub.astype(np.uint8) * 255, inpaintRadius=3, flags=cv2.INPAINT_TELEA)


def finish_mask(mask_u8: np.ndarray) -> np.ndarray:
    """stereo_rerender.py:803-808,817 on the u8 image the GPU produced: inpaint (TELEA, radius 3) all background-green
    and black pixels from the coded normals, keep the result in the green (hole) pixels only, masked blur."""
    green = np.all(mask_u8 == np.asarray(GREEN, dtype=np.uint8), axis=-1)
    area = green | np.all(mask_u8 == 0, axis=-1)
    filled = telea_fill_holes(mask_u8, green, area)
    mask = mask_u8.astype(np.float64) / 255.0
    mask[green] = filled[green].astype(np.float32) / 255.0
    blurred = masked_blur((mask * 255).astype(np.uint8)).astype(np.float32) / 255.0
    return (blurred * 255).astype(np.uint8)


class InfillMaskRenderer:
    """StereoRerenderer plus the normals-coded mask.  render_device returns (sbs u8 (n, H, 2W, 3) with edge colours
    painted into the holes, mask image u8 (n, H, 2W, 3) BEFORE inpainting); finish() runs the host part."""

    def __init__(self, renderer: StereoRerenderer, workers: Optional[int] = None):
        import os

        self.r = renderer
        self.pool = ThreadPoolExecutor(max_workers=max(1, workers or (os.cpu_count() or 8)))
        self._zbuf = None
        self._flags = self._normals = self._holes = None

    def render_device(self, depth_rgb: torch.Tensor, colour: torch.Tensor, start_frame: int = 0, out_sbs: Optional[torch.Tensor] = None,
                      out_mask_img: Optional[torch.Tensor] = None, code_normals: bool = True, paint_edge_colours: bool = True,
                      out_depth: Optional[torch.Tensor] = None):
        r, p = self.r, self.r.p
        n, h, w, _ = depth_rgb.shape
        dev = depth_rgb.device
        if out_sbs is None:
            out_sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device=dev)
        if out_mask_img is None:
            out_mask_img = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device=dev)
        if self._holes is None or self._holes.shape != (n, h, 2 * w):
            self._holes = torch.empty((n, h, 2 * w), dtype=torch.uint8, device=dev)
        if self._zbuf is None or tuple(self._zbuf.shape) != (h, w):
            self._zbuf = ops.new_zbuf(1, w, h, dev)[0]
            self._flags = torch.empty((h, w), dtype=torch.uint8, device=dev)
            self._normals = torch.empty((h, w, 3), dtype=torch.float64, device=dev)
        # 1. the stereo render with a plain u8 hole mask (what bg_mask is in the reference, :740,854)
        r.render_device(depth_rgb, colour, start_frame, out_sbs, self._holes, out_depth, mask_rgb=False)
        # 2. edge vertices once per frame, edge points once per eye
        for k in range(n):
            f = start_frame + k
            xf = p.xfov_of(f)
            K = geo.compute_camera_matrix(xf, None if p.xfovs is not None else p.yfov, w, h)
            src = ops.make_source(w, h, K, p.max_depth, "D1", True, geo.master_fov_depth_scale(p.master_xfov, xf), True)  # mesh grid: of_by_one (:575)
            ops.edge_vertices(depth_rgb[k], src, K, True, flags=self._flags, normals=self._normals)
            for e, view in enumerate(r.views_of(f)):
                half = slice(e * w, (e + 1) * w)
                ops.edge_splat(depth_rgb[k], src, K, self._flags, view.M, K, w, h, self._zbuf)
                ops.edge_resolve(self._zbuf, depth_rgb[k], src, K, self._normals if code_normals else None, view.M, colour[k],
                                 self._holes[k, :, half], out_mask_img[k, :, half], out_sbs[k, :, half] if paint_edge_colours else None,
                                 p.bg_rgb, code_normals)
        return out_sbs, out_mask_img

    def basic_infill(self, sbs: torch.Tensor, final_mask_img: torch.Tensor, max_steps: int = 400) -> torch.Tensor:
        """--do_basic_infill (stereo_rerender.py:810-812,898-900): fill the holes of both eyes of the frames rendered by the
        last render_device call by marching along the normals coded in the FINAL mask images (n, H, 2W, 3) u8 (device)."""
        n, h, w2, _ = sbs.shape
        w = w2 // 2
        for k in range(n):
            for e in range(2):
                half = slice(e * w, (e + 1) * w)
                ops.normal_march_infill(sbs[k, :, half], self._holes[k, :, half], final_mask_img[k, :, half], max_steps)
        return sbs

    def finish_async(self, mask_img_host: np.ndarray) -> "DeferredFrames":
        """(n, H, 2W, 3) u8 pre-inpaint mask images (host) -> final mask frames, finished on the worker pool; each eye is
        finished on its own, as the reference does (left_img_mask / right_img_mask, :805-808,893-896).  The inputs are
        copied before this returns (the caller may reuse its buffer); `.result()` waits for the frames."""
        n, h, w2, _ = mask_img_host.shape
        w = w2 // 2
        out = np.empty_like(mask_img_host)

        def task(src, k, e):
            out[k, :, e * w:(e + 1) * w] = finish_mask(src)

        futures = [self.pool.submit(task, np.ascontiguousarray(mask_img_host[k, :, e * w:(e + 1) * w]), k, e)
                   for k in range(n) for e in range(2)]
        return DeferredFrames(out, futures)

    def finish(self, mask_img_host: np.ndarray) -> np.ndarray:
        return self.finish_async(mask_img_host).result()


class DeferredFrames:
    """Frames that host workers are still producing (the TELEA + blur tail of the infill mask): the frame loop keeps
    reading, rendering and encoding the other outputs meanwhile and asks for `.result()` one chunk later."""

    def __init__(self, frames: np.ndarray, futures):
        self._frames, self._futures = frames, futures

    def result(self) -> np.ndarray:
        for f in self._futures:
            f.result()
        return self._frames
