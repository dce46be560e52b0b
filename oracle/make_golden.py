"""Generate tests/golden/*.npz by RUNNING the reference (read-only checkout) on seeded inputs.

    python oracle/make_golden.py            # needs /root/reference (or MDVT_REFERENCE_ROOT)

TEST INFRASTRUCTURE ONLY.  The fixtures hold reference *outputs* (plus the small inputs that
produced them); no reference source is stored.  Environment the committed vectors were made
with: Python 3.12.3, NumPy 2.3.5, OpenCV 4.13.0, SciPy 1.18.1 (reference @ 1b621e96).
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


def all_code_frame(seed=7):
    """(256, 256, 3) RGB frame covering every (R, B) pair once, G random (G must be ignored by
    D1/D3 and averaged by D2)."""
    rng = np.random.default_rng(seed)
    r, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    g = rng.integers(0, 256, size=r.shape, dtype=np.uint8)
    return np.stack((r, g, b), axis=-1)


def golden_decode(dfh):
    rgb = all_code_frame()
    out = {"rgb": rgb}
    h, w = rgb.shape[:2]
    for md in (100, 20):
        out[f"d1_depth_md{md}"] = dfh.decode_rgb_depth_frame(rgb, md, True)
        out[f"d1_depth24_md{md}"] = dfh.decode_rgb_depth_frame(rgb, md, False)
        # D2: convert_metric_depth_video_to_other_format.py:646-652, D3: find_convergence_depth.py:56-60
        ns = {"np": np, "rgb": rgb, "frame_height": h, "frame_width": w, "MODEL_maxOUTPUT_depth": md}
        ref_bridge.exec_lines("convert_metric_depth_video_to_other_format.py", 646, 652, ns)
        out[f"d2_depth_md{md}"] = ns["depth"]
        ns = {"np": np, "rgb": rgb, "frame_height": h, "frame_width": w, "MODEL_maxOUTPUT_depth": md}
        ref_bridge.exec_lines("find_convergence_depth.py", 56, 60, ns)
        out[f"d3_depth_md{md}"] = ns["depth"]
    out["d1_codes"] = dfh.decode_rgb_as_data(rgb, w, h, True)
    out["d1_codes24"] = dfh.decode_rgb_as_data(rgb, w, h, False)
    np.savez_compressed(os.path.join(OUT, "decode_all_codes.npz"), **out)


def golden_encode(dfh):
    rng = np.random.default_rng(11)
    d = rng.uniform(0, 100, size=(48, 64)).astype(np.float32)
    d[0, :8] = [0.0, 100.0, 101.0, -3.0, 1e-9, 99.99999, 0.0015, 50.0]
    d[1, :4] = [np.float32(100.0 * k / 65535.0) for k in (1, 2, 65534, 65535)]
    out = {"depth": d}
    for md in (100, 20):
        codes = dfh.encode_depth_as_uint32(d, md)
        out[f"codes_md{md}"] = codes
        out[f"bgr16_md{md}"] = dfh.encode_data_as_BGR(codes, 64, 48, True)
        out[f"bgr24_md{md}"] = dfh.encode_data_as_BGR(codes, 64, 48, False)
    np.savez_compressed(os.path.join(OUT, "encode.npz"), **out)


def golden_camera(dmt):
    cases = [(60.0, None, 640, 480), (None, 45.0, 640, 480), (60.0, 40.0, 1920, 1080), (90.0, None, 3840, 2160),
             (75, 75, 1920, 1920), (33.3, None, 64, 48)]
    Ks, fovs = [], []
    for fx, fy, w, h in cases:
        K = dmt.compute_camera_matrix(fx, fy, w, h)
        Ks.append(K)
        fovs.append(dmt.fov_from_camera_matrix(K))
    np.savez_compressed(os.path.join(OUT, "camera.npz"),
                        cases=np.array([[np.nan if v is None else v for v in c] for c in cases], dtype=np.float64),
                        K=np.array(Ks), fov=np.array(fovs, dtype=np.float64))


def random_pose(rng, max_angle=0.05, max_shift=0.2):
    a, b, c = rng.uniform(-max_angle, max_angle, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = rng.uniform(-max_shift, max_shift, 3)
    return T


def golden_geometry(dfh, dmt, sr):
    """One 64x48 synthetic frame through every importable stage, plus a strided sample of a
    640x480 one (config 1 size) to keep the fixture small."""
    rng = np.random.default_rng(5)
    for tag, (w, h), stride in (("64x48", (64, 48), 1), ("640x480", (640, 480), 101)):
        clip = SyntheticClip(w, h, 3, seed=1234, zero_fraction=0.005)
        depth_rgb, colour = clip.frame(1)
        out = {"depth_rgb": depth_rgb if stride == 1 else np.zeros(0, np.uint8),
               "colour": colour if stride == 1 else np.zeros(0, np.uint8),
               "size": np.array([w, h]), "stride": np.array(stride)}
        depth = dfh.decode_rgb_depth_frame(depth_rgb, 100, True)
        K = dmt.compute_camera_matrix(60.0, None, w, h)
        out["K"] = K
        out["depth_sample"] = depth.reshape(-1)[::stride]
        for obo in (False, True):
            pts, hh, ww = dmt.create_point_cloud_from_depth(depth, K, obo)
            assert pts.dtype == np.float64 and (hh, ww) == (h, w)
            out[f"xyz_obo{int(obo)}"] = pts[::stride]
        pts, _, _ = dmt.create_point_cloud_from_depth(depth, K, False)
        T = random_pose(rng)
        out["T"] = T
        moved = dmt.transform_points(pts, T)
        out["xyz_T"] = moved[::stride]
        # cv2.projectPoints twin; z == 0 points are excluded by the caller in the oracle
        uv = dmt.project_3d_points_to_2d(moved, K)
        out["uv_T"] = uv[::stride]
        if stride == 1:
            # reference point painter, stereo_rerender.py:746-755 + the write at :814, run on the
            # left-eye points (translate +ipd/2, :725) of the frame
            eye = moved + np.array([0.0315, 0.0, 0.0])
            front = eye[:, 2] > 1e-4
            ns = {"np": np, "frame_width": w, "frame_height": h,
                  "points_2d": dmt.project_3d_points_to_2d(eye[front], K), "points_3d": eye[front],
                  "edge_colors": colour.reshape(-1, 3)[front].astype(np.float64) / 255.0,
                  "unprojected_normals": np.zeros((int(front.sum()), 3))}
            ref_bridge.exec_lines("stereo_rerender.py", 746, 755, ns)
            img = np.zeros((h, w, 3), dtype=np.float64)
            vp, vc = ns["valid_points"], ns["valid_colors"]
            img[vp[:, 1], vp[:, 0]] = vc  # far -> near, last write wins (:814)
            out["painter_img"] = (img * 255).astype(np.uint8)  # :819
            drawn = np.zeros((h, w), dtype=bool)
            drawn[vp[:, 1], vp[:, 0]] = True
            out["painter_drawn"] = drawn
        np.savez_compressed(os.path.join(OUT, f"geometry_{tag}.npz"), **out)


def golden_misc(dmt, sr):
    rng = np.random.default_rng(21)
    out = {}
    for k, (pos, tgt) in enumerate([((2.0, 2.0, -4.0), (0.1, -0.2, 7.5)), ((0.0, 0.0, 0.0), (0.0, 0.0, 1.0)),
                                    ((-1.5, 0.3, 2.0), (4.0, 1.0, 9.0))]):
        out[f"lookat_in{k}"] = np.array([pos, tgt], dtype=np.float64)
        out[f"lookat_out{k}"] = dmt.cam_look_at(np.array(pos).astype(np.float32), np.array(tgt, dtype=np.float64))
    out["conv_angle_in"] = np.array([[0.5, 0.063], [3.0, 0.063], [12.5, 0.07], [100.0, 0.063]])
    out["conv_angle_out"] = np.array([sr.convergence_angle(d, p) for d, p in out["conv_angle_in"]])
    for k, n in enumerate((2, 5, 60, 300)):  # n == 1 raises ValueError inside savgol_filter (reference quirk)
        vals = rng.uniform(1.0, 9.0, n)
        if n > 3:
            vals[rng.random(n) < 0.15] = np.nan
            vals[1] = 3.0  # keep at least one number
        out[f"conv_in{k}"] = vals.copy()
        filled = sr.fill_nan_with_closest(list(vals))
        out[f"conv_filled{k}"] = np.array(filled)
        out[f"conv_smooth{k}"] = np.asarray(sr.curve_fit(filled))
    # VR180 remap (stereo_rerender.py:25-86) on small frames: the reference's own maps + cv2.remap
    eq = np.random.default_rng(33)
    for k, (hw, fov) in enumerate((((48, 64), 75.0), ((60, 60), 100.0), ((33, 70), 120.0))):
        img = eq.integers(0, 256, hw + (3,), dtype=np.uint8)
        out[f"equirect_in{k}"], out[f"equirect_fov{k}"] = img, np.array(fov)
        out[f"equirect_out{k}"] = sr.convert_to_equirectangular(img, input_fov=fov)
    np.savez_compressed(os.path.join(OUT, "misc.npz"), **out)


def main():
    if not ref_bridge.available():
        raise SystemExit(f"reference checkout not found at {ref_bridge.REFERENCE_ROOT}")
    os.makedirs(OUT, exist_ok=True)
    dfh = ref_bridge.load("depth_frames_helper")
    dmt = ref_bridge.load("depth_map_tools")
    sr = ref_bridge.load("stereo_rerender")
    golden_decode(dfh)
    golden_encode(dfh)
    golden_camera(dmt)
    golden_geometry(dfh, dmt, sr)
    golden_misc(dmt, sr)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
