"""Checks shared by the tests and by __graft_entry__.smoke(): which target pixels a float32 evaluation of a projection may
legitimately paint differently from the float64 oracle.  Test infrastructure, like the rest of oracle/: the
product never imports it."""
import numpy as np

from . import mdvt_oracle as orc


def explained_map(u64, v64, z64, w, h, near_plane=None):
    """(h, w) bool: the target pixels a float32 evaluation of the same projection may legitimately paint differently from the
    float64 oracle -- a source whose float64 (u', v') lies within 1e-3 px of a .5 rounding boundary rounds into the pixel or
    one of its 8 neighbours, or the pixel's two nearest candidates differ by less than 1e-5 relative in z' (the
    same criteria as boundary_explained, for comparisons of final images where no winner ids are at hand)."""
    near_plane = orc.NEAR_PLANE if near_plane is None else near_plane
    u64, v64, z64 = (np.asarray(a, dtype=np.float64).reshape(-1) for a in (u64, v64, z64))
    valid = np.isfinite(u64) & np.isfinite(v64) & (z64 > near_plane)
    with np.errstate(invalid="ignore"):
        near_half = valid & ((np.abs((u64 - np.floor(u64)) - 0.5) < 1e-3) | (np.abs((v64 - np.floor(v64)) - 0.5) < 1e-3))
    pad = np.zeros((h + 2, w + 2), dtype=bool)
    ur, vr = np.rint(u64[near_half]), np.rint(v64[near_half])
    keep = (ur >= -1) & (ur <= w) & (vr >= -1) & (vr <= h)
    ur, vr = ur[keep].astype(np.int64), vr[keep].astype(np.int64)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            rr, cc = vr + dy, ur + dx
            ok = (rr >= 0) & (rr < h) & (cc >= 0) & (cc < w)
            pad[rr[ok] + 1, cc[ok] + 1] = True
    out = pad[1:-1, 1:-1].copy()
    with np.errstate(invalid="ignore"):
        ui, vi = np.rint(u64), np.rint(v64)
        inside = valid & (ui >= 0) & (ui < w) & (vi >= 0) & (vi < h)
    t = (vi[inside] * w + ui[inside]).astype(np.int64)
    z = z64[inside]
    order = np.lexsort((z, t))
    t, z = t[order], z[order]
    if len(t) > 1:
        first = np.flatnonzero(np.r_[True, t[1:] != t[:-1]])
        second = first + 1
        has2 = second < len(t)
        has2[has2] &= t[second[has2]] == t[first[has2]]
        za, zb = z[first[has2]], z[second[has2]]
        tie = np.abs(za - zb) <= 1e-5 * np.maximum(np.abs(za), np.abs(zb))
        out.reshape(-1)[t[first[has2]][tie]] = True
    return out


def assert_differs_only_where_explained(got, want, maps, max_fraction=2e-3):
    """Final images (h, W[, 3]) of a float32 path and of the float64 oracle: every differing pixel lies in the explained map
    (one map per view, side by side like the images), and they are few."""
    diff = (np.asarray(got) != np.asarray(want))
    if diff.ndim == 3:
        diff = diff.any(axis=-1)
    explained = np.concatenate(list(maps), axis=1) if isinstance(maps, (list, tuple)) else maps
    assert explained.shape == diff.shape, (explained.shape, diff.shape)
    bad = diff & ~explained
    assert not bad.any(), f"{int(bad.sum())} differing pixels are not next to a rounding boundary or a z tie, e.g. {np.argwhere(bad)[:5].tolist()}"
    assert diff.mean() <= max_fraction, f"{diff.mean():.2e} of the pixels differ"
