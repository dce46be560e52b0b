"""VR180 output: rectilinear render -> 180-degree equirectangular frame of the same size
(stereo_rerender.convert_to_equirectangular, stereo_rerender.py:25-86).

The coordinate maps depend on (H, W, input_fov) only: they are built once on the host in float64 exactly as the
reference builds them (so the float32 maps handed to the remap are the same bits), cached on the device, and the
per-pixel bilinear remap runs in `mdvt_remap_bilinear_u8x3`, bit-exact with cv2.remap."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

from . import ops
from .depth_frames_helper import _device, _down, _up

_maps: Dict[Tuple[int, int, float, str], Tuple[torch.Tensor, torch.Tensor]] = {}


def equirect_maps(height: int, width: int, input_fov: float):
    """(map_x, map_y) float32 (H, W): output pixel -> source pixel, -1 outside the source's field of view."""
    cx, cy = (width - 1) / 2.0, (height - 1) / 2.0
    gx, gy = np.meshgrid(np.linspace(0, width - 1, width), np.linspace(0, height - 1, height))
    theta = (gx - cx) / cx * (np.pi / 2)          # longitude in [-90, 90] degrees
    phi = (gy - cy) / cy * (np.pi / 2)            # latitude
    half = np.radians(input_fov / 2.0)
    fx, fy = cx / np.tan(half), cy / np.tan(half)  # pinhole focal lengths of the rendered view
    inside = (np.abs(theta) <= half) & (np.abs(phi) <= half)
    map_x = fx * np.tan(theta) + cx
    map_y = fy * np.tan(phi) + cy
    map_x[~inside] = -1
    map_y[~inside] = -1
    return map_x.astype(np.float32), map_y.astype(np.float32)


def device_maps(height: int, width: int, input_fov: float, device):
    key = (height, width, float(input_fov), str(device))
    if key not in _maps:
        if len(_maps) > 16:
            _maps.clear()
        mx, my = equirect_maps(height, width, input_fov)
        _maps[key] = (torch.from_numpy(mx).to(device), torch.from_numpy(my).to(device))
    return _maps[key]


def convert_to_equirectangular(image, input_fov=100):
    """Drop-in for stereo_rerender.convert_to_equirectangular: (H, W, 3) u8 -> (H, W, 3) u8."""
    img, as_np = _up(image, torch.uint8)
    h, w = img.shape[:2]
    mx, my = device_maps(h, w, input_fov, img.device)
    return _down(ops.remap_bilinear(img, mx, my), as_np)
