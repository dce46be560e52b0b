"""Generate tests/golden/dropin_helpers.npz by RUNNING the reference's own helper functions (read-only checkout):
calculate_normals, get_cam_view, gl_look_at, open_cv_w2c_to_gl_view, reject_outliers, apply_side_view_to_paralax_mask
(depth_map_tools.py) and infill_using_normals (stereo_rerender.py) -- the names other reference scripts import from
the two modules, which the drop-in modules of this repo now export too.

    python oracle/make_dropin_golden.py            # needs /root/reference (or MDVT_REFERENCE_ROOT)

TEST INFRASTRUCTURE ONLY.  The fixture holds reference *outputs* plus the small seeded inputs that produced them.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_bridge  # noqa: E402
from metric_depth_video_toolbox_b200.synth import SyntheticClip  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    dmt = ref_bridge.load("depth_map_tools")
    sr = ref_bridge.load("stereo_rerender")
    dfh = ref_bridge.load("depth_frames_helper")
    out = {}
    # ---- calculate_normals on a decoded synthetic frame (step edges + a ramp) and on noise --------------------
    w, h = 96, 64
    depth_rgb, _ = SyntheticClip(w, h, 3, zero_fraction=0.01).frame(1)
    depth = dfh.decode_rgb_depth_frame(depth_rgb, 100, True)
    K = dmt.compute_camera_matrix(60.0, 47.0, w, h)
    out["cn_depth"], out["cn_K"] = depth, K
    out["cn_normals"] = dmt.calculate_normals(depth, K)
    rng = np.random.default_rng(5)
    noise = rng.uniform(0.3, 30.0, size=(33, 41)).astype(np.float32)
    K2 = dmt.compute_camera_matrix(90.0, None, 41, 33)
    out["cn_depth2"], out["cn_K2"] = noise, K2
    out["cn_normals2"] = dmt.calculate_normals(noise, K2)
    # ---- small matrix helpers -----------------------------------------------------------------------------------
    out["cam_view_fwd"] = dmt.get_cam_view(0.0315, 0.0063)
    out["cam_view_rev"] = dmt.get_cam_view(-0.0315, 0.02, reverse=True)
    out["gl_look_at"] = dmt.gl_look_at(np.array([1.0, 2.0, 3.0], dtype=np.float32), np.array([0.5, -1.0, -4.0], dtype=np.float32),
                                       np.array([0.0, 1.0, 0.0], dtype=np.float32))
    T = np.eye(4)
    c, s = np.cos(0.3), np.sin(0.3)
    T[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]) @ np.array([[1, 0, 0], [0, np.cos(0.1), -np.sin(0.1)], [0, np.sin(0.1), np.cos(0.1)]])
    T[:3, 3] = (0.4, -0.2, 1.5)
    out["w2c_in"], out["w2c_gl"] = T, dmt.open_cv_w2c_to_gl_view(T)
    data = rng.normal(size=40)
    out["outlier_in"], out["outlier_mask"] = data, dmt.reject_outliers(data, 1.5)
    pm = rng.random((12, 16)) > 0.5
    nm = rng.normal(size=(12, 16, 3)).astype(np.float32)
    out["side_pm"], out["side_n"] = pm, nm
    out["side_right"], out["side_left"] = dmt.apply_side_view_to_paralax_mask(pm, nm, True), dmt.apply_side_view_to_paralax_mask(pm, nm, False)
    # ---- infill_using_normals with its own signature (float normal map) ----------------------------------------
    hh, ww = 48, 64
    img = rng.integers(1, 256, size=(hh, ww, 3), dtype=np.uint8)
    hole = np.zeros((hh, ww), dtype=bool)
    hole[10:22, 12:30] = True
    hole[30:44, 40:60] = True
    hole[5, 50:58] = True
    hole[rng.random((hh, ww)) < 0.02] = True
    normal_map = rng.uniform(-1, 1, size=(hh, ww, 3)).astype(np.float32)
    normal_map[12:14, 14:20] = (0.0, 1.0, 0.0)      # "green": no normal (:176)
    normal_map[16, 14:20, :2] = 0.0                 # zero direction: invalid (:172)
    normal_map[32:36, 42:50] = (1.0, 0.0, 0.3)      # axis-aligned marches (ties of rint)
    normal_map[36:40, 42:50] = (-0.5, 0.5, 0.0)
    img[hole] = 0
    out["iun_img"], out["iun_hole"], out["iun_normals"] = img, hole, normal_map
    out["iun_out"] = sr.infill_using_normals(img, hole, normal_map)
    out["iun_out_12"] = sr.infill_using_normals(img, hole, normal_map, max_steps=12)
    # ---- encode_depth_as_uint32 on a float64 depth array (clip and scale stay float64 in the reference) ----------
    d64 = rng.uniform(0, 100, size=(32, 48))
    d64[0, :6] = [0.0, 100.0, 100.0000001, -1e-12, 99.99999999, 50.0]
    d64[1, :4] = [100.0 * k / 65535.0 for k in (1, 2, 65534, 65535)]
    d64[2, :8] = np.nextafter(np.float32(37.5), np.float32(40)).astype(np.float64) - 1e-9   # values float32 would round across a code
    out["enc64_depth"] = d64
    for md in (100, 20):
        codes = dfh.encode_depth_as_uint32(d64, md)
        out[f"enc64_codes_md{md}"] = codes
        out[f"enc64_bgr16_md{md}"] = dfh.encode_data_as_BGR(codes, 48, 32, bit16=True)
    np.savez_compressed(os.path.join(OUT, "dropin_helpers.npz"), **out)
    print("wrote", os.path.join(OUT, "dropin_helpers.npz"), {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__" and "--convert-helpers" not in sys.argv:
    main()


def convert_helpers_golden():
    """tests/golden/convert_helpers.npz: the two NumPy helpers of convert_metric_depth_video_to_other_format.py on seeded inputs."""
    import numpy as np

    from oracle import ref_bridge

    ref = ref_bridge.load("convert_metric_depth_video_to_other_format")
    rng = np.random.default_rng(77)
    img = np.concatenate([rng.uniform(-1.0, 14.0, size=(6, 9)), np.array([[0.0, 1e-5, 1e-4, 10.0, 9.99999, 1e9, 2.5, 0.3, 7.0]])]).astype(np.float32)
    depth = rng.uniform(0.0, 6.0, size=400)
    depth[::17] = 0.0
    target = 1.0 / (0.83 * (1.0 / np.where(depth > 0, depth, 1.0)) + 0.021) + rng.normal(0, 1e-3, size=400)
    target[5::23] = -1.0
    s, t = ref.estimate_scale_shift(depth, target)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "convert_helpers.npz"), img=img, bytes_default=ref.float_image_to_byte_image(img),
                        bytes_custom=ref.float_image_to_byte_image(img, max_value=5.0, scale=200, log_scale=2), depth=depth, target=target,
                        scale_shift=np.array([s, t]))


if __name__ == "__main__" and "--convert-helpers" in sys.argv:
    convert_helpers_golden()
