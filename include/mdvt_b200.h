/*
 * mdvt_b200.h -- C ABI of libmdvt_b200.so: the dense per-frame path of
 * calledit/metric_depth_video_toolbox (RGB-encoded depth decode -> pinhole unproject -> pose ->
 * perspective re-divide -> z-buffered forward splat + hole mask) as sm_100a CUDA kernels.
 *
 * The reference has no FFI of its own: its boundary for this path is the Python module surface
 * of depth_frames_helper.py / depth_map_tools.py and the per-frame loops of stereo_rerender.py,
 * 3d_view_depthfile.py and convert_metric_depth_video_to_other_format.py.  Each
 * entry point below names the reference lines it replaces; INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; nothing is allocated,
 *     freed or retained; all work is enqueued asynchronously on `stream` (a cudaStream_t passed
 *     as void*, NULL = legacy default stream); entry points are re-entrant.
 *   - frames are dense row-major u8x3 in RGB order, exactly what the reference holds after
 *     cv2.cvtColor(BGR2RGB) (stereo_rerender.py:493); pixel index = row * W + col.
 *   - return value: MDVT_OK or a negative mdvt_status; mdvt_last_error() gives the text for the
 *     calling thread.
 *   - float32 arithmetic is IEEE round-to-nearest, one rounding per written operation, no FMA
 *     contraction (oracle/kernel_model.py restates the op order); the depth decode is bit-exact
 *     with the reference (integer code, one float32 multiply or divide).
 */
#ifndef MDVT_B200_H_
#define MDVT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDVT_ABI_VERSION 11

#if defined(__GNUC__)
#define MDVT_API __attribute__((visibility("default")))
#else
#define MDVT_API
#endif

typedef enum mdvt_status {
    MDVT_OK = 0,
    MDVT_ERR_INVALID_ARGUMENT = -1,
    MDVT_ERR_UNSUPPORTED = -2,   /* shape / alignment the requested kernel cannot take */
    MDVT_ERR_CUDA = -3,          /* a CUDA runtime call failed; see mdvt_last_error() */
    MDVT_ERR_NO_DEVICE = -4      /* no sm_100 device visible */
} mdvt_status;

/* Which of the reference's three (non-identical) depth decoders to reproduce. */
typedef enum mdvt_decoder {
    MDVT_DECODE_D1 = 0, /* depth_frames_helper.py:13-24,63-75,99-103  hi=R lo=B, code * fl32(max/255^4) */
    MDVT_DECODE_D2 = 1, /* convert_metric_depth_video_to_other_format.py:646-652  hi=(R+G)>>1, code / fl32(255^4/max) */
    MDVT_DECODE_D3 = 2, /* find_convergence_depth.py:56-60  hi=R lo=B, code / fl32(255^4/max) */
    MDVT_SOURCE_F32 = 3 /* not a decoder: the source buffer is a plane of float32 depths the caller already holds
                           (the `depth_map` argument of depth_map_tools.get_mesh_from_depth_map /
                           create_point_cloud_from_depth, depth_map_tools.py:1104-1133); dec_const / bit16 unused */
} mdvt_decoder;

/* Empty z-buffer slot: all ones.  A filled slot is (float_bits(z') << 32) | source_pixel_index. */
#define MDVT_ZBUF_EMPTY 0xFFFFFFFFFFFFFFFFull

/* Source-frame description shared by the unproject / splat entry points. */
typedef struct mdvt_source {
    int32_t width, height;
    int32_t decoder;      /* mdvt_decoder */
    int32_t bit16;        /* 1: 16-bit wire format (the only one the scripts use); 0: 24-bit (D1 only) */
    float dec_const;      /* D1: fl32(max_depth / 255^4) (multiplier); D2/D3: fl32(255^4 / max_depth) (divisor) */
    float depth_scale;    /* stereo_rerender.py:537-541 master-FOV scale as fl32; 1.0f = none */
    float fx, fy, cx, cy; /* depth_map_tools.py:902-934, rounded to float32 */
    float grid_sx, grid_sy; /* of_by_one stretch fl32((W+1)/W), fl32((H+1)/H); 1.0f = exact pixel grid
                               (depth_map_tools.py:1118-1123) */
} mdvt_source;

/* One virtual camera: p_view = M * [p_src; 1] (3x4 row-major), then u = fx X/Z + cx, v = fy Y/Z + cy. */
typedef struct mdvt_view {
    float M[12];
    float fx, fy, cx, cy;
} mdvt_view;

/* Per-frame constants of the row-local stereo fast path (no pose file, no convergence:
 * v' == v and u' = u +- fx*(ipd/2)/z, stereo_rerender.py:458-459,704-725,831-836). */
typedef struct mdvt_stereo_frame {
    float dec_const;    /* as mdvt_source.dec_const (decoder D1) */
    float depth_scale;  /* as mdvt_source.depth_scale */
    float fx_half_ipd;  /* fl32(fx * ipd / 2), ipd in metres */
    float near_plane;   /* cull z' <= near (depth_map_tools.py:1520: 1e-4) */
} mdvt_stereo_frame;

/* resolve / stereo flags */
#define MDVT_FLAG_BG_COLLIDE   0x1u /* a rendered colour equal to `bg_rgb` counts as a hole (stereo_rerender.py:740,854) */
#define MDVT_FLAG_RESET_ZBUF   0x2u /* resolve leaves the z-buffer empty for the next frame */
#define MDVT_FLAG_ANYWIDTH      0x8u /* mdvt_stereo_rows: force the any-width kernel even when W % 32 == 0 (test aid; same results) */
#define MDVT_FLAG_MASK_RGB     0x4u /* hole mask written as u8x3 (bg_rgb at holes, black elsewhere, :787-793) instead of u8 {0,255} */

/* ---- library -------------------------------------------------------------------------------- */
MDVT_API int mdvt_abi_version(void);
MDVT_API const char *mdvt_version(void);
MDVT_API const char *mdvt_last_error(void);
/* SM count, L2 bytes, shared memory per block (opt-in) of the current device; any pointer may be NULL. */
MDVT_API int mdvt_device_info(int *sm_count, int *l2_bytes, int *smem_optin_bytes, int *cc_major, int *cc_minor);

/* ---- wire-format codec ---------------------------------------------------------------------- */
/* depth_frames_helper.decode_rgb_depth_frame / decode_rgb_as_data / decode_uint32_as_depth
 * (depth_frames_helper.py:13-24,63-75,99-103) and the inline D2/D3 variants.
 * rgb: n_pixels*3 u8.  out_codes (u32) and out_depth (f32) may each be NULL. */
MDVT_API int mdvt_decode_depth(const uint8_t *rgb, int64_t n_pixels, int decoder, int bit16, float dec_const,
                      uint32_t *out_codes, float *out_depth, void *stream);

/* depth_frames_helper.encode_depth_as_uint32 + encode_data_as_BGR (depth_frames_helper.py:5-11,48-61):
 * clip to [0,max_depth], (255^4/max_depth * double(depth)) truncated to u32, bytes -> u8x3.
 * `bgr_order` = 1 writes B,G,R (what the reference returns for cv2), 0 writes R,G,B.
 * out_codes may be NULL; out_pix may be NULL. */
MDVT_API int mdvt_encode_depth(const float *depth, int64_t n_pixels, double max_depth, int bit16, int bgr_order,
                      uint32_t *out_codes, uint8_t *out_pix, void *stream);

/* The same for a float64 depth array: the reference clips in the array's own dtype and multiplies float64(depth), so a
 * float64 input must not be rounded to float32 first (its codes differ in the low bits and, at truncation boundaries, in
 * the 16-bit wire bytes). */
MDVT_API int mdvt_encode_depth_f64(const double *depth, int64_t n_pixels, double max_depth, int bit16, int bgr_order,
                          uint32_t *out_codes, uint8_t *out_pix, void *stream);

/* decode_uint32_as_depth (depth_frames_helper.py:13-24) on a plane of codes the caller holds: D1 multiplies by
 * dec_const, D2/D3 divide by it. */
MDVT_API int mdvt_codes_to_depth(const uint32_t *codes, int64_t n_pixels, int decoder, float dec_const, float *out_depth,
                        void *stream);
/* encode_data_as_BGR (depth_frames_helper.py:48-61): bytes of a u32 plane -> u8x3 (16-bit: R=G=byte3, B=byte2;
 * 24-bit: R=byte2, G=byte1, B=byte0); bgr_order as in mdvt_encode_depth. */
MDVT_API int mdvt_codes_to_pixels(const uint32_t *codes, int64_t n_pixels, int bit16, int bgr_order, uint8_t *out_pix,
                         void *stream);

/* ---- other depth export formats ---------------------------------------------------------------- */
/* `depth_src`: n*3 u8 wire-format pixels, or n float32 depths with decoder == MDVT_SOURCE_F32.
 * Grey depth video frames (convert_metric_depth_video_to_other_format.py:752-760):
 * out = astype(rint(depth * factor)) with factor = fl32(255^2 / max_depth) -> u16 x1 (--bit16) or
 * fl32(255 / max_depth) -> u8 replicated to 3 channels (--bit8).  out_bits 8|16, channels 1|3 (16-bit: 1). */
MDVT_API int mdvt_depth_to_grey(const void *depth_src, int64_t n_pixels, int decoder, int bit16, float dec_const,
                       float factor, int out_bits, int channels, void *out, void *stream);
/* Touchly reverse-depth plane (stereo_rerender.py:548-552; rendered-depth variant :687-690,:826-829 with
 * zero_is_far = 1): 255 - rint(max(0, min(depth*depth_scale, tmax) - tmin) * gain), gain = fl32(255/(tmax-tmin))
 * evaluated by the caller in double like the script does; u8 x3, row r at out_rgb + r*out_pitch so it can be
 * written into its slot of the stacked output frame. */
MDVT_API int mdvt_touchly_depth(const void *depth_src, int width, int height, int decoder, int bit16, float dec_const,
                       float depth_scale, float touchly_min, float touchly_max, float gain, int zero_is_far,
                       uint8_t *out_rgb, int64_t out_pitch, void *stream);

/* Hole mask plane u8 {0, non-zero} -> one bit per pixel, most significant bit first (numpy.packbits order), (n_pixels + 7) / 8
 * bytes: the form in which the host API ships masks over PCIe when asked to (StereoRerenderer.render_host(mask_format="bits");
 * the reference's mask is the u8 / green-black image of stereo_rerender.py:787-793, recovered with numpy.unpackbits).
 * mask: 16-byte aligned, bits: 2-byte aligned. */
MDVT_API int mdvt_pack_mask_bits(const uint8_t *mask, int64_t n_pixels, uint8_t *bits, void *stream);

/* cv2.remap(src, map_x, map_y, INTER_LINEAR, BORDER_CONSTANT, border_rgb) for u8x3 images and float32 maps
 * (dst_w*dst_h each), bit-exact with OpenCV's fixed-point bilinear (1/32-pixel coordinates, 15-bit weights): the
 * per-pixel part of stereo_rerender.convert_to_equirectangular (stereo_rerender.py:25-86), used for --vr180 /
 * --touchly0 (:914-916).  Row r of src / dst at + r*pitch bytes. */
MDVT_API int mdvt_remap_bilinear_u8x3(const uint8_t *src, int src_w, int src_h, int64_t src_pitch, const float *map_x,
                             const float *map_y, int dst_w, int dst_h, uint32_t border_rgb, uint8_t *dst,
                             int64_t dst_pitch, void *stream);

/* ---- decode + unproject (+ pose) -> point cloud --------------------------------------------- */
/* `depth_src` below is the u8x3 wire-format frame (n*3 bytes), or n float32 depths when
 * src->decoder == MDVT_SOURCE_F32.
 * decode -> depth_scale -> depth_map_tools.create_point_cloud_from_depth (depth_map_tools.py:1112-1133)
 * -> optional 3x4 pose (transform_points, depth_map_tools.py:977-1004).  pose_host: 12 floats on
 * the HOST (row-major 3x4) or NULL.  out_xyz: n*3 f32. */
MDVT_API int mdvt_unproject_f32(const void *depth_src, const mdvt_source *src_host, const float *pose_host,
                       float *out_xyz, void *stream);
/* Same in float64 with the reference's own operation order (K and pose given as doubles): xyz is
 * bit-identical to NumPy for the unprojection.  K_host: fx, fy, cx, cy; pose_host: 12 doubles or NULL.
 * out_xyz: n*3 f64 -- the .ply export path (convert_metric_depth_video_to_other_format.py:692-749). */
MDVT_API int mdvt_unproject_f64(const void *depth_src, const mdvt_source *src_host, const double *K_host,
                       const double *pose_host, double *out_xyz, void *stream);

/* depth_map_tools.transform_points (depth_map_tools.py:977-1004): out = [xyz 1] @ T.T without the w divide.
 * pose_host: upper 3x4 of T, 12 doubles row-major, on the HOST.  In place (out_xyz == xyz) is allowed. */
MDVT_API int mdvt_transform_points_f64(const double *xyz, int64_t n_points, const double *pose_host, double *out_xyz,
                              void *stream);
/* depth_map_tools.project_3d_points_to_2d (depth_map_tools.py:1057-1060; cv2.projectPoints with zero
 * rvec/tvec/distortion): out_uv = (fx X/Z + cx, fy Y/Z + cy), n*2 f64.  K_host: fx, fy, cx, cy. */
MDVT_API int mdvt_project_points_f64(const double *xyz, int64_t n_points, const double *K_host, double *out_uv,
                            void *stream);

/* ---- generic novel-view path: fused decode/unproject/pose/project + z-buffered splat, then resolve */
MDVT_API int mdvt_zbuf_clear(uint64_t *zbuf, int64_t n_slots, void *stream);

/* K1+K2.  For each of n_views cameras (views_host: HOST array, n_views <= 4) every source pixel is
 * decoded, unprojected, moved by M, culled if z' <= near, projected, rounded half-to-even, bounds
 * checked against out_w x out_h and merged into zbuf[view] with one 64-bit atomicMin
 * (nearest z' wins, ties -> lowest source index; stereo_rerender.py:746-755,814 and the GL depth
 * test of depth_map_tools.py:1563-1572).  zbuf: n_views * out_w*out_h u64, pre-cleared.
 * The payload written is id_offset + source pixel index, so several objects can share one z-buffer and one
 * concatenated colour table (render() takes a list of objects, depth_map_tools.py:1422).
 * out_uvz (optional, n_views * n_pixels * 3 f32) receives (u', v', z') per source pixel for parity checks. */
MDVT_API int mdvt_project_splat(const void *depth_src, const mdvt_source *src_host, const mdvt_view *views_host,
                       int n_views, float near_plane, int out_w, int out_h, uint32_t id_offset, uint64_t *zbuf,
                       float *out_uvz, void *stream);

/* Same visibility rule for explicit points (n_points x 3 float32, in the space the views' M maps from): the
 * reference's point painter (stereo_rerender.py:746-755,814) and render() of point clouds. */
MDVT_API int mdvt_splat_points(const float *xyz, int64_t n_points, const mdvt_view *views_host, int n_views,
                      float near_plane, int out_w, int out_h, uint32_t id_offset, uint64_t *zbuf, void *stream);

/* K3.  zbuf (one view, out_w*out_h) + source colours -> image / hole mask / depth plane.
 * out_rgb: row r starts at out_rgb + r*rgb_pitch (bytes), so a view can be written straight into its
 * half of a side-by-side frame (cv2.hconcat, stereo_rerender.py:918); same for out_mask / mask_pitch.
 * out_depth (f32, row r at out_depth + r*depth_pitch floats, depth_pitch 0 = out_w; 0 where nothing was drawn --
 * the `left_depth` / `right_depth` planes of stereo_rerender.py:738,852), out_ids (int32, dense, -1 = hole) optional.
 * bg_rgb / fill_rgb: 0x00BBGGRR packed (R in the low byte). */
MDVT_API int mdvt_resolve(uint64_t *zbuf, const uint8_t *colour_rgb, int out_w, int out_h, uint32_t bg_rgb,
                 uint32_t fill_rgb, uint32_t flags, uint8_t *out_rgb, int64_t rgb_pitch, uint8_t *out_mask,
                 int64_t mask_pitch, float *out_depth, int64_t depth_pitch, int32_t *out_ids, void *stream);

/* Where the per-frame, per-view planes of mdvt_render_views go: plane (frame f, view v) starts at
 * base + f*frame_stride + v*view_stride, its row r at + r*row_pitch (all in BYTES).  A side-by-side stereo
 * frame (cv2.hconcat, stereo_rerender.py:918) is {frame_stride = H*2W*3, view_stride = W*3, row_pitch = 2W*3}. */
typedef struct mdvt_plane_layout {
    void *base;
    int64_t frame_stride, view_stride, row_pitch;
} mdvt_plane_layout;

/* The generic frame loop in one call (stereo_rerender.py:471-941 with a pose file / convergence rotation;
 * 3d_view_depthfile.py:133-255): for each of n_frames frames, K1+K2 of all n_views cameras
 * (views_host[f*n_views + v]) into `zbuf`, then K3 of all views into the planes described by rgb_out / mask_out
 * (optional) / depth_out (optional, f32).  sources_host: one mdvt_source per frame (per_frame_source = 1) or one for
 * all.  depth_src / colour_rgb: frame f at + f*frame_stride bytes.
 * zbuf: zbuf_sets * n_views planes of out_w*out_h slots, all ones on entry (mdvt_zbuf_clear) and on return.  Inside the
 * call a slot holds epoch << 56 | float_bits(z') << 25 | 0x00BBGGRR of the nearest point: the colour rides in the key (no
 * gather in K3; candidates with bit-identical z' are ordered by colour, where the index-keyed mdvt_project_splat takes the
 * lowest source index and the reference an unstable argsort), and the epoch byte counts down from frame to frame, so
 * planes are not re-armed between frames.  zbuf_sets = 2 lets the call run odd frames on an internal second stream
 * (forked from / joined to `stream` with events) next to the even ones. */
MDVT_API int mdvt_render_views(const void *depth_src, int64_t depth_frame_stride, const uint8_t *colour_rgb,
                      int64_t colour_frame_stride, int n_frames, const mdvt_source *sources_host, int per_frame_source,
                      const mdvt_view *views_host, int n_views, float near_plane, int out_w, int out_h, uint64_t *zbuf,
                      int zbuf_sets, uint32_t bg_rgb, uint32_t fill_rgb, uint32_t flags, const mdvt_plane_layout *rgb_out,
                      const mdvt_plane_layout *mask_out, const mdvt_plane_layout *depth_out, void *stream);

/* The camera of `3d_view_depthfile.py --render` (3d_view_depthfile.py:231-241, depth_map_tools.py:1618-1638,1528-1552):
 * look-at point = vertex centroid of the frame (Open3D get_center()) with per-axis --tx/--ty/--tz overrides. */
typedef struct mdvt_lookat {
    double cam_pos[3];     /* --x --y --z after .astype(np.float32) (:240), widened */
    double target[3];      /* --tx --ty --tz */
    int32_t target_set[3]; /* 0: that component is the centroid's */
    int32_t reserved;
    double y_scale;        /* fy / fx: render() scales geometry Y and passes fx in both focal slots (:1528-1552) */
    float fx, fy, cx, cy;  /* intrinsics of the rendering camera (fy == fx for the reference's path) */
} mdvt_lookat;

/* The frame loop of `3d_view_depthfile.py --render` (:133-255) for n_frames frames in ONE call with no host round
 * trip: per frame the vertex centroid (as mdvt_centroid, vertex grid `centroid_src_host`, usually the of_by_one
 * stretched grid of mesh mode), the look-at camera of that frame computed ON THE DEVICE into views_dev[f], K1+K2 of
 * the frame (pixel grid `src_host`) through that camera, K3 into the planes of rgb_out / mask_out (optional).
 * poses_host: n_frames x 16 doubles (row-major 4x4, the frame's transformation) or NULL.  sums_dev: n_frames x
 * (4 + MDVT_REDUCE_SCRATCH_DOUBLES) doubles (results {sum X, sum Y, sum Z, n} first); views_dev: n_frames
 * mdvt_view, 16-byte aligned; both stay valid for the caller to read back.  zbuf: zbuf_sets (1 or 2) planes of out_w x
 * out_h slots, all ones on entry and on return (colour-keyed and epoch-tagged inside the call, see mdvt_render_views).
 * touched: zbuf_sets * mdvt_touched_bytes(out_w, out_h) bytes of scratch (per-segment "something was drawn here" flags that
 * let K3 skip the z-buffer traffic of empty regions).  With zbuf_sets = 2 the odd frames run on an internal second stream
 * (forked from / joined to `stream` with events) next to the even ones. */
MDVT_API int64_t mdvt_touched_bytes(int out_w, int out_h);
MDVT_API int mdvt_novel_view_frames(const void *depth_src, int64_t depth_frame_stride, const uint8_t *colour_rgb,
                           int64_t colour_frame_stride, int n_frames, const mdvt_source *centroid_src_host,
                           const mdvt_source *src_host, const double *K_host, const double *poses_host,
                           const mdvt_lookat *look_host, float near_plane, int out_w, int out_h, uint64_t *zbuf, int zbuf_sets,
                           double *sums_dev, mdvt_view *views_dev, uint8_t *touched, uint32_t bg_rgb, uint32_t fill_rgb,
                           uint32_t flags, const mdvt_plane_layout *rgb_out, const mdvt_plane_layout *mask_out,
                           void *stream);

/* ---- per-frame reductions (fixed summation order: reproducible) -------------------------------- */
/* Result buffers hold 4 doubles of result followed by MDVT_REDUCE_SCRATCH_DOUBLES doubles of scratch. */
#define MDVT_REDUCE_SCRATCH_DOUBLES 4096

/* Sum of the unprojected (+posed) vertices in float64 with the reference's evaluation order
 * (create_point_cloud_from_depth, then Open3D get_center() = mean of vertices, 3d_view_depthfile.py:231):
 * out_sums = {sum X, sum Y, sum Z, n}.  K_host: fx, fy, cx, cy doubles; pose_host: 12 doubles or NULL. */
MDVT_API int mdvt_centroid(const void *depth_src, const mdvt_source *src_host, const double *K_host,
                  const double *pose_host, double *out_sums, void *stream);

/* find_convergence_depth.py:56-80: sum / count (/ sum of squares) of the decoded depth over the pixels whose
 * mask byte is > mask_gt (mask NULL: all pixels).  out_sums = {sum, count, sum of squares, 0}. */
MDVT_API int mdvt_depth_sum(const void *depth_src, int64_t n_pixels, int decoder, int bit16, float dec_const,
                   const uint8_t *mask, int mask_gt, double *out_sums, void *stream);

/* ---- normals-coded infill mask: the per-pixel parts (stereo_rerender.py --infill_mask) ------------ */
/* E1.  The edge test of the reference's mesh builder on the depth grid (depth_map_tools.py:1243-1376, called with
 * remove_edges=True, return_normals_of_removed=True at stereo_rerender.py:583), float64 like the reference:
 * out_flags[p] = 1 where vertex p belongs to a triangle seen at more than angle_threshold_deg (89) degrees,
 * out_normals[3p..] = the unit normal the reference attaches to it (written for flagged vertices only; may be NULL).
 * cell_flags_scratch: (W-1)*(H-1) bytes.  The grid is the source's (grid_sx/sy: of_by_one stretch). */
MDVT_API int mdvt_edge_vertices(const void *depth_src, const mdvt_source *src_host, const double *K_host,
                       double angle_threshold_deg, uint8_t *cell_flags_scratch, uint8_t *out_flags, double *out_normals,
                       void *stream);
/* E1 on explicit vertices: depth_map_tools.create_mesh_from_point_cloud(points, height, width, remove_edges=True, ...)
 * (depth_map_tools.py:1186-1416) called directly with a grid-organised (H*W, 3) float64 point array. */
MDVT_API int mdvt_edge_vertices_xyz(const double *xyz, int width, int height, double angle_threshold_deg,
                           uint8_t *cell_flags_scratch, uint8_t *out_flags, double *out_normals, void *stream);
/* E2.  Flagged vertices -> "edge points" (stereo_rerender.py:589-606) -> pose (12 doubles, frame -> eye camera,
 * :615-619,723-732) -> projection with the render camera (K_render_host: fx, fy, cx, cy as the float32-rounded values
 * the reference hands to cv2.projectPoints, :735) -> np.round -> z-buffered into zbuf (out_w*out_h u64, pre-cleared;
 * nearest wins, :745-755). */
MDVT_API int mdvt_edge_splat(const void *depth_src, const mdvt_source *src_host, const double *K_host, const uint8_t *flags,
                    const double *pose_host, const double *K_render_host, int out_w, int out_h, uint64_t *zbuf,
                    void *stream);
/* E3.  Per target pixel of one eye: mask image (u8x3) = bg_rgb at holes, black elsewhere; with code_normals the
 * border holes get the fixed inward normals of :796-799 and a hole pixel hit by an edge point gets that point's
 * normal, rotated into the eye camera, as (n+1)/2*255 (:733,778-780,802); `image` (optional, in/out) receives the edge
 * point's colour there (:813-814).  hole_mask: u8, non-zero = hole (what mdvt_resolve / mdvt_stereo_rows wrote).
 * Leaves zbuf empty.  The TELEA inpainting + masked blur that follow (:805-808) are OpenCV calls on the host. */
MDVT_API int mdvt_edge_resolve(uint64_t *zbuf, const void *depth_src, const mdvt_source *src_host, const double *K_host,
                      const double *normals, const double *pose_host, const uint8_t *colour_rgb, const uint8_t *hole_mask,
                      int64_t hole_pitch, int out_w, int out_h, uint32_t bg_rgb, int code_normals, uint8_t *image,
                      int64_t image_pitch, uint8_t *mask_img, int64_t mask_pitch, void *stream);

/* --do_basic_infill: stereo_rerender.infill_using_normals (stereo_rerender.py:155-240,810-812) for one eye, in place.
 * image: u8x3 eye image with black holes; hole_mask: u8, non-zero = hole; mask_img: the FINAL mask image (u8x3, after
 * inpainting and blur) whose R, G channels code the march direction as (m/255)*2-1.  max_steps: 400 in the reference. */
MDVT_API int mdvt_normal_march_infill(uint8_t *image, int64_t image_pitch, const uint8_t *hole_mask, int64_t hole_pitch,
                             const uint8_t *mask_img, int64_t mask_pitch, int width, int height, int max_steps,
                             void *stream);

/* The same march with the function's own signature (stereo_rerender.infill_using_normals(color_img, hole_mask, normal_map,
 * max_steps), stereo_rerender.py:155-240, imported by basic_nomal_infill.py:10,101): normal_map is a dense (H, W, 3)
 * float32 plane whose x, y components give the direction; a normal of exactly (0, 1, 0) marks "no normal" (:176). */
MDVT_API int mdvt_normal_march_infill_f32(uint8_t *image, int64_t image_pitch, const uint8_t *hole_mask, int64_t hole_pitch,
                                 const float *normal_map, int width, int height, int max_steps, void *stream);

/* depth_map_tools.calculate_normals(depth, K) (depth_map_tools.py:20-60): per-pixel unit normals of a float32 depth plane
 * from the forward differences of the unprojected points, y and z negated; float32 with NumPy's operation order.
 * K_host: fx, fy, cx, cy doubles (rounded to float32 like NumPy rounds Python scalars).  out_normals: H*W*3 f32. */
MDVT_API int mdvt_calculate_normals(const float *depth, int width, int height, const double *K_host, float *out_normals,
                           void *stream);

/* ---- row-local stereo fast path: ONE fused kernel, frames batched ---------------------------- */
/* Whole stereo_rerender.py frame loop body (:512-541 decode+scale, :583 unproject, :723-738,:831-852 eye
 * poses + render, :740,:787-793,:854 hole mask, :918 hconcat) for a batch of frames when there is no
 * pose file and no convergence rotation.  depth_rgb / colour_rgb: n_frames*H*W*3 u8;
 * frames_dev: DEVICE array of mdvt_stereo_frame, n_frames entries (or 1 entry if per_frame == 0);
 * out_sbs: n_frames * H * 2W * 3 u8 (left | right); out_mask: n_frames * H * 2W u8 {0,255}
 * (or u8x3 with MDVT_FLAG_MASK_RGB), may be NULL; out_depth: n_frames * H * 2W float32 rendered depth
 * (left | right, 0 where nothing was drawn -- render(depth=-2), :738,:852), may be NULL.  z-buffers live in
 * shared memory; nothing else touches HBM.  Requires W <= 65535. */
MDVT_API int mdvt_stereo_rows(const uint8_t *depth_rgb, const uint8_t *colour_rgb, int n_frames, int width, int height,
                     const mdvt_stereo_frame *frames_dev, int per_frame, uint32_t bg_rgb, uint32_t fill_rgb,
                     uint32_t flags, uint8_t *out_sbs, uint8_t *out_mask, float *out_depth, void *stream);

/* ---- stereo with a convergence rotation: ONE fused kernel, frames batched ------------------------ */
/* Per-frame constants (DEVICE array, one entry per frame).  view[0] = left eye, view[1] = right eye; each M must be a
 * rotation about the y axis plus a translation along x (M[1] = M[4] = M[6] = M[7] = M[9] = M[11] = 0, M[5] = 1), which
 * is what stereo_rerender.py builds without a pose file (:704-725,831-836): then a pixel's target row does not depend
 * on its depth and one CTA can own a target row.  The output has the size of the source frames. */
typedef struct mdvt_conv_frame {
    float dec_const, depth_scale, near_plane, reserved;
    float fx, fy, cx, cy;   /* source camera, exact pixel grid */
    mdvt_view view[2];
} mdvt_conv_frame;

/* Same inputs / outputs / flags as mdvt_stereo_rows; results are bit-identical to mdvt_render_views with the same cameras
 * (nearest Zv wins, candidates with bit-identical Zv are ordered by packed colour: the rule of the colour-keyed frame loops),
 * without the global z-buffer.  Round 1's target-row kernel (per-column source-row prediction); any width up to 4096. */
MDVT_API int mdvt_stereo_conv_rows(const uint8_t *depth_rgb, const uint8_t *colour_rgb, int n_frames, int width, int height,
                          const mdvt_conv_frame *frames_dev, uint32_t bg_rgb, uint32_t fill_rgb, uint32_t flags,
                          uint8_t *out_sbs, uint8_t *out_mask, float *out_depth, void *stream);

/* The same job in "virtual source row" form (csrc/mdvt_stereo_vrows.cu): along a target row the source row index is a
 * staircase in the source column, so the TMA engine assembles one virtual source row per eye in shared memory and the
 * row is processed like a row of mdvt_stereo_rows, with two 32-bit shared-memory atomic passes (Zv, then colour) standing
 * in for the 64-bit key.  Bit-identical to mdvt_render_views / mdvt_stereo_conv_rows (nearest Zv wins, candidates with
 * bit-identical Zv are ordered by packed colour).  Requires W % 32 == 0, W <= 3840, 16-byte aligned buffers and poses inside
 * the limits mdvt_stereo_conv_vrows_supported checks (else MDVT_ERR_UNSUPPORTED / undefined rows flagged in status_dev).
 * status_dev: optional n_frames int32 (DEVICE), set to 1 for a frame whose geometry left the limits (never for frames the
 * host check accepted); zero it before the call. */
MDVT_API int mdvt_stereo_conv_vrows(const uint8_t *depth_rgb, const uint8_t *colour_rgb, int n_frames, int width, int height,
                           const mdvt_conv_frame *frames_dev, uint32_t bg_rgb, uint32_t fill_rgb, uint32_t flags,
                           uint8_t *out_sbs, uint8_t *out_mask, float *out_depth, int32_t *status_dev, void *stream);

/* Host-side check (no CUDA call): 1 when every frame of frames_host (HOST array) is a pose mdvt_stereo_conv_vrows handles
 * at this size -- y-rotation + x-shift, row step within [0.9, 1.12], staircase slope below 0.4 rows per 15 columns
 * (convergence distances down to ~0.6 m at 1080p) -- else 0 (use mdvt_render_views). */
MDVT_API int mdvt_stereo_conv_vrows_supported(const mdvt_conv_frame *frames_host, int n_frames, int width, int height);

/* ---- FFV1 result-video encoder -----------------------------------------------------------------------------------------
 * Replaces the entropy coder behind the reference's cv2.VideoWriter(fourcc "FFV1") result writers
 * (stereo_rerender.py:420-442,941; depth_frames_helper.py:125-161; 3d_view_depthfile.py:118-127), i.e. libavcodec's FFV1
 * version 3 encoder with the parameters OpenCV selects (Golomb-Rice coder, RGB colourspace with the JPEG2000 RCT, 8 bit,
 * per-slice CRC), with two differences that keep the stream standard and its decoded frames bit-identical: every frame
 * is a key frame, and a frame is cut into nh x nv (<= 1024) slices instead of 2 x 2 -- one device thread codes one slice.
 * alpha = 1 adds the constant-255 alpha plane OpenCV's BGRA input produces; alpha = 0 writes the 3-plane stream.
 * context_model = 0 uses libavcodec's quant tables for 8-bit content (666 contexts per plane context, what OpenCV's files
 * carry: at OpenCV's 2 x 2 slices + alpha the key-frame packets are byte-identical to libavcodec's); context_model = 1
 * writes a 5-level table instead (63 contexts: a tenth of the coder state per slice; the tables travel in the
 * configuration record, so any FFV1 decoder follows); context_model = 2 writes a 3-level table (14 contexts, 224 bytes of
 * coder state per slice, which the device coder then keeps in shared memory: ~1.8 x the encoder throughput of model 1 for ~1 %
 * larger files on film-like content; the `states` buffer is not touched in that case unless alpha = 1). */

/* HOST function, no device needed.  Writes the codec configuration record (Matroska CodecPrivate; <= 64 bytes) and the
 * range-coded header of each of the nh * nv slices (16 bytes reserved per slice; header_len_host[s] bytes used). */
MDVT_API int mdvt_ffv1_stream_setup(int width, int height, int nh, int nv, int alpha, int context_model, uint8_t *config_host,
                           int config_capacity, int *config_len, uint8_t *headers_host, int32_t *header_len_host);

/* Bytes to reserve per slice (worst case of the coder, a multiple of 16), and bytes of coder state for a batch; -1 on
 * bad arguments. */
MDVT_API int64_t mdvt_ffv1_slice_capacity(int width, int height, int nh, int nv, int alpha);
MDVT_API int64_t mdvt_ffv1_state_bytes(int n_frames, int nh, int nv, int alpha, int context_model);

/* Encodes n_frames u8x3 frames (RGB order, or BGR with bgr_order = 1).  headers / header_len: DEVICE copies of what
 * mdvt_ffv1_stream_setup wrote.  states: mdvt_ffv1_state_bytes scratch.  slices: n_frames * nh * nv * capacity bytes of
 * scratch.  Results: sizes[f * S + s] = bytes of slice s of frame f (S = nh * nv), offsets[f * S + s] = its position in
 * `packed`, offsets[n_frames * S] = total bytes; packed[offsets[f * S] .. offsets[(f + 1) * S]) is the packet of frame f,
 * ready for a Matroska SimpleBlock with the key flag.  `packed` must hold n_frames * S * capacity bytes. */
MDVT_API int mdvt_ffv1_encode_frames(const uint8_t *frames, int64_t frame_stride, int64_t row_pitch, int n_frames, int width,
                            int height, int nh, int nv, int alpha, int context_model, int bgr_order, const uint8_t *headers,
                            const int32_t *header_len, void *states, uint8_t *slices, int64_t capacity, int32_t *sizes,
                            int64_t *offsets, uint8_t *packed, void *stream);

/* ---- FFV1 reader for the streams written above ------------------------------------------------------------------------
 * The mirror image of mdvt_ffv1_encode_frames: one device thread decodes one slice.  Only streams with this library's
 * parameters are accepted (FFV1 v3 written by cv2.VideoWriter has 2 x 2 slices and carries coder state across the 12
 * frames of a GOP -- 4 serial threads per GOP -- and stays with cv2.VideoCapture on the host, as in the reference:
 * stereo_rerender.py:471-503, depth_frames_helper.py load_video_frames_from_path). */

/* HOST function.  Reads nh / nv / alpha / context_model from a configuration record (Matroska CodecPrivate) and returns
 * MDVT_OK only if the record is byte for byte what mdvt_ffv1_stream_setup writes for them; MDVT_ERR_UNSUPPORTED otherwise. */
MDVT_API int mdvt_ffv1_parse_config(const uint8_t *config_host, int config_len, int width, int height, int *nh, int *nv, int *alpha,
                           int *context_model);

/* packets: the n_frames packets back to back (DEVICE), packet_offsets[n_frames + 1] their bounds.  states:
 * mdvt_ffv1_state_bytes scratch; slice_offsets: n_frames * nh * nv int64 of scratch.  frames: u8x3 output, RGB order (BGR
 * with bgr_order = 1).  status[f] (DEVICE, one per frame): 0 = decoded; -2 the slice sizes of the packet do not add up,
 * -3 a slice header is not the expected one (e.g. a non-key frame), -4 a slice size is inconsistent, -5 the bit stream
 * of a slice overran, -6 a slice's CRC-32 is wrong (every slice is checked before it is decoded, as ffv1dec.c does). */
MDVT_API int mdvt_ffv1_decode_frames(const uint8_t *packets, const int64_t *packet_offsets, int n_frames, int width, int height, int nh,
                            int nv, int alpha, int context_model, int bgr_order, const uint8_t *headers, const int32_t *header_len, void *states,
                            int64_t *slice_offsets, uint8_t *frames, int64_t frame_stride, int64_t row_pitch, int32_t *status,
                            void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MDVT_B200_H_ */
