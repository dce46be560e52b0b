// Stereo with a convergence rotation, "virtual source rows" form (stereo_rerender.py:704-725,831-836 with
// --convergence_file, the way movie_2_3D drives it): ONE fused kernel per batch of frames, no global z-buffer, the
// structure of the row-local kernel (mdvt_stereo_rows.cu) kept as far as the geometry allows.
//
// Each eye pose is a rotation about the camera's y axis followed by a shift along x.  In ray form (mdvt_common.cuh)
// that leaves B_u = B_z = T_v = T_z = 0: Zv = z * r_z(j), u' = (z r_u(j) + T_u) / Zv, and v' = r_v(i, j) / r_z(j) does
// not depend on the depth.  The source row that lands in target row r is, per source column j,
//     i*(j) = ((r A_z - A_v) j + (r C_z - C_v)) / B_v            -- LINEAR in j, |slope| ~ sin(theta) |r - cy| / fx,
// so along a target row the source row index is a staircase with a handful of steps.  A CTA owns one target row of both
// eyes at a time.  Its producer warp evaluates the staircase per 16-pixel sub-block and eye in float64 and has the TMA
// engine assemble, per eye, a VIRTUAL SOURCE ROW in shared memory: for each run of sub-blocks with the same source row
// one cp.async.bulk of that row's depth bytes and one of its colour bytes, at the columns' natural offsets.  The compute
// warps then treat the virtual row like the row-local kernel treats a real one: lane-strided columns, word loads, decode,
// the exact float32 arithmetic of the generic path for Zv and u' (mdvt_splat.cu: FFMA chain, refined reciprocal, correctly
// rounded quotient, magic-number rounding), conflict-free shared-memory atomics.
//   * Where the float64 prediction says |v' - r| < 0.495 for every column of a sub-block, rint(v') == r is certain (the
//     float32 chain is within 1e-3 px of the prediction) and v' is not even computed.  Otherwise (sub-blocks next to a
//     step of the staircase, ~1 % of the pixels) the candidate goes through the exact v' arithmetic and must round to r,
//     and the neighbouring source row that can also round into r is fetched as an ALTERNATE sub-block (48 + 48 bytes)
//     and evaluated the same way.
//   * Visibility: nearest Zv wins, candidates with bit-identical Zv are ordered by packed colour -- the colour-keyed order
//     of the generic frame loop (mdvt_render_views), so the results are bit-identical to it (asserted by the tests on every
//     byte, mask and depth plane).  Shared memory has no 64-bit atomic min, so the 55-bit key is split over two 32-bit
//     planes and two passes: (1) every candidate: ATOMS.MIN of float_bits(Zv) into the z plane; barrier; (2) every
//     candidate whose Zv is the slot's minimum: ATOMS.MIN of its colour into the colour plane.  A thread keeps its
//     candidates (slot, Zv bits, colour) in registers between the passes.
//   * Phase B reads the colour plane four pixels at a time, re-arms both planes, packs RGB / mask bytes into the dead raw
//     buffer of the row; the producer warp hands that to a TMA bulk store and meanwhile has loaded the next row.
// HBM traffic is the algorithmic 14 B/px (each source row is read by ~2-3 neighbouring target rows of the same CTA and by
// both eyes: L2 hits).  Geometry outside the limits below (checked on the host by mdvt_stereo_conv_vrows_supported) goes
// through the generic frame loop instead.
#include <cmath>
#include <cstdlib>

#include "mdvt_common.cuh"

namespace mdvt {

constexpr int kSub = 16;            // pixels per sub-block: 48 bytes of u8x3, the TMA granularity of the assembly
constexpr int kAltSlots = 48;       // alternate sub-blocks with their data staged in shared memory (more are read from global memory)
constexpr uint32_t kEmpty32 = 0xFFFFFFFFu;
constexpr float kVMagic = 12582912.0f;  // 1.5 * 2^23
constexpr int kVMagicBits = 0x4B400000;
constexpr float kVMagicInt = 8388608.0f;
// (prediction windows, in pixels, sit next to their use in the producer: the float32 chain and the prediction of v' differ by
// < 2e-3 px for coordinates < 8192)
constexpr double kStepMin = 0.9, kStepMax = 1.12, kMaxSubSpan = 0.4;

struct VrowSmem {
    int alt_off, alt_stride;      // [2 buffers][1 + 4 nsub] u32: count, then eye << 31 | sub-block << 16 | source row
    int altdata_off, altdata_stride;  // [2 buffers][kAltSlots][96]: depth 48 | colour 48
    int raw_off, raw_stride;      // [2 buffers][2 eyes][depth 3W | colour 3W]
    int zp_off, cp_off;           // [2 eyes][W + 4] u32 each
    int out_off, mask_off, total; // staging of the left | right output row and of its mask row
};

__host__ __device__ inline VrowSmem vrow_smem_layout(int width, int mask_bpp) {
    VrowSmem L;
    const int nsub = width / kSub;
    int off = 64 + 2 * 96;  // [0,16): two mbarriers; [64, 256): two StairFrame slots
    L.alt_off = off;  L.alt_stride = (1 + 4 * nsub) * 4;       off += 2 * L.alt_stride;
    off = (off + 15) & ~15;
    L.altdata_off = off; L.altdata_stride = kAltSlots * 96;    off += 2 * L.altdata_stride;
    L.raw_off = off;  L.raw_stride = 12 * width;               off += 2 * L.raw_stride;
    L.zp_off = off;   off += 2 * (width + 4) * 4;
    L.cp_off = off;   off += 2 * (width + 4) * 4;
    L.out_off = off;  off += 6 * width;
    L.mask_off = off; off += 2 * width * mask_bpp;
    L.total = (off + 15) & ~15;
    return L;
}

// The staircase of one target row and eye (float64; the same text runs on the host in the limits check).
struct Staircase {
    double alpha, beta;   // i*(j) = alpha j + beta
    double smin, smax;    // range of |d v' / d i| = |B_v / r_z(j)| over the row
};
__host__ __device__ inline Staircase staircase_of(const RayView &rv, int r, int width) {
    Staircase s;
    const double Az = rv.A[2], Av = rv.A[1], Cz = rv.C[2], Cv = rv.C[1], Bv = rv.B[1];
    s.alpha = ((double)r * Az - Av) / Bv;
    s.beta = ((double)r * Cz - Cv) / Bv;
    const double s0 = fabs(Bv / Cz), s1 = fabs(Bv / (Az * (double)(width - 1) + Cz));
    s.smin = s0 < s1 ? s0 : s1;
    s.smax = s0 < s1 ? s1 : s0;
    return s;
}
__host__ __device__ inline bool staircase_ok(const Staircase &s) {
    return s.smin >= kStepMin && s.smax <= kStepMax && fabs(s.alpha) * (kSub - 1) < kMaxSubSpan;
}

struct EyeConsts {  // float32 ray-form coefficients the compute warps use (B_u = B_z = T_v = T_z = 0)
    float Au, Av, Az, Cu, Cv, Cz, Bv, Tu;
};

// ---- shared memory through 32-bit addresses: loads / reductions with immediate offsets, no generic-pointer arithmetic ----
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t atoms_add(uint32_t addr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void reds_min(uint32_t addr, uint32_t v) { asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// One candidate through the exact float32 arithmetic of splat_pixel() (mdvt_splat.cu), with the additions that are exact
// zeros left out: target column (>= width when culled or out of range) and the key bits of Zv; Zv and its refined
// reciprocal are handed on for the row test of the alternates.
__device__ __forceinline__ uint32_t vrow_project(float z, float fj, const EyeConsts &k, float near_plane, uint32_t width, uint32_t &zbits,
                                                 float &Zv, float &rz) {
    const float cju = __fmaf_rn(k.Au, fj, k.Cu), cjz = __fmaf_rn(k.Az, fj, k.Cz);
    const float nu = __fmaf_rn(z, cju, k.Tu);
    Zv = __fmaf_rn(z, cjz, 0.0f);
    rz = rcp_refined(Zv);
    const float u = div_rn_by(nu, Zv, rz);
    const uint32_t ui = (uint32_t)(__float_as_int(__fadd_rn(u, kVMagic)) - kVMagicBits);
    zbits = __float_as_uint(Zv);
    return min(Zv > near_plane ? ui : 0xFFFFFFFFu, width);
}
// rint(v') of the same candidate when it comes from source row `frow_src` (only where the prediction is not certain)
__device__ __forceinline__ int vrow_target_row(float z, float fj, const EyeConsts &k, float frow_src, float Zv, float rz) {
    const float cjv = __fmaf_rn(k.Av, fj, k.Cv);
    const float nv = __fmaf_rn(z, __fmaf_rn(k.Bv, frow_src, cjv), 0.0f);
    const float v = div_rn_by(nv, Zv, rz);
    return __float_as_int(__fadd_rn(v, kVMagic)) - kVMagicBits;
}

__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar_addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void named_barrier(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Per-frame constants of the staircase (shared memory, two slots keyed by frame parity): i*(j) of target row r is
// alpha j + beta with alpha = r k1 - k0, beta = r c1 - c0 (k1 = A_z / B_v, k0 = A_v / B_v, c1 = C_z / B_v, c0 = C_v / B_v, float64),
// and the row step |B_v / r_z(j)| lies in [smin, smax] for every row.
struct StairFrame {
    double k1[2], k0[2], c1[2], c0[2];
    float smin[2], smax[2];
    int frame, pad;
};

// MASK_MODE: 0 none, 1 u8 {0,255}, 2 u8x3 (bg colour / black).  T threads; every thread owns the columns tid + n T, n < CPT.
// GUARD: W < T * CPT (columns past the row are culled).
template <int MASK_MODE, int T, int CPT, bool GUARD>
__global__ void __launch_bounds__(T, T <= 384 ? 2 : 1)
    stereo_conv_vrows_kernel(const uint8_t *__restrict__ depth_rgb, const uint8_t *__restrict__ colour_rgb, int n_units, int width, int height,
                             const mdvt_conv_frame *__restrict__ frames, uint32_t bg_rgb, uint32_t fill_rgb, int collide,
                             uint8_t *__restrict__ out_sbs, uint8_t *__restrict__ out_mask, float *__restrict__ out_depth,
                             int32_t *__restrict__ status) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int mask_bpp = MASK_MODE == 2 ? 3 : 1;
    constexpr int kWarps = T / 32;
    const VrowSmem L = vrow_smem_layout(width, mask_bpp);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);  // bar[0], bar[1]: the two sets of row buffers
    StairFrame *s_stair = reinterpret_cast<StairFrame *>(smem + 64);  // [2]
    const int tid = threadIdx.x, lane = tid & 31;
    const int nsub = width / kSub;
    const int n_items = 2 * ((nsub + 31) / 32);  // (eye, chunk of 32 sub-blocks): the units of the row preparation, one warp each
    const uint32_t row_bytes = 3u * width;
    const int plane = width + 4;  // slots per eye (W + the dummy slot, padded to 16 bytes)

    if (tid == 0) {
        mbar_init(&bar[0], (uint32_t)n_items);
        mbar_init(&bar[1], (uint32_t)n_items);
        mbar_fence_init();
        s_stair[0].frame = s_stair[1].frame = -1;
        *reinterpret_cast<uint32_t *>(smem + L.alt_off) = 0u;
        *reinterpret_cast<uint32_t *>(smem + L.alt_off + L.alt_stride) = 0u;
    }
    {   // z planes: all ones; colour planes: the flagged fill colour (any real colour, < 2^24, beats it; a hole reads as the fill)
        uint32_t *zp = reinterpret_cast<uint32_t *>(smem + L.zp_off), *cp = reinterpret_cast<uint32_t *>(smem + L.cp_off);
        for (int k = tid; k < 2 * plane; k += T) {
            zp[k] = kEmpty32;
            cp[k] = fill_rgb | 0xFF000000u;
        }
    }
    __syncthreads();

    // contiguous block of target rows per CTA: successive rows need almost the same source rows (L2 / TMA locality) and the
    // frame constants change once or twice per CTA
    const int per_cta = (n_units + gridDim.x - 1) / gridDim.x;
    const int unit_begin = blockIdx.x * per_cta, unit_end = min(n_units, unit_begin + per_cta);
    if (unit_begin >= unit_end) return;
    int frame = unit_begin / height, r = unit_begin - frame * height;  // the row being processed
    int frame2 = frame, r2 = r;                                        // the row being prepared (two ahead in the steady state)
    const int warp = tid >> 5;

    // ---- row preparation: one (eye, chunk) item per warp --------------------------------------------------------
    // Sub-blocks whose source row is CERTAIN for every column go into the virtual row: runs of sub-blocks with one source row ->
    // one bulk copy per array (a lane finds the end of its run in the ballot of the run starts).  Every other sub-block gets
    // zero depth bytes (code 0 never passes Zv > near >= 0) and its candidates -- the predicted row where it is not certain,
    // the neighbouring row where that one can round into r too -- go on the list of alternates, which are evaluated with the
    // exact row test.  Per row the warp forms alpha and beta - r in float64 and rounds them to float32: the sub-block
    // arithmetic runs on delta(j) = i*(j) - r, a number of a few tens at most (absolute error < 1e-5 rows against the 5e-3
    // margin of the windows).  Every item arrives once on the row's mbarrier with its own transaction bytes.
    auto prepare_items = [&](int buf) {  // row (frame2, r2) -> buffers `buf`
        constexpr float kCertainF = 0.494f, kPossibleF = 0.506f;
        for (int item = warp; item < n_items; item += kWarps) {
            const int e = item & 1, s = (item >> 1) * 32 + lane;
            StairFrame *sf = &s_stair[frame2 & 1];
            double k1, k0, c1, c0;
            float smin, smax;
            if (sf->frame == frame2) {
                k1 = sf->k1[e]; k0 = sf->k0[e]; c1 = sf->c1[e]; c0 = sf->c0[e]; smin = sf->smin[e]; smax = sf->smax[e];
            } else {  // first row of a frame in this CTA: every preparing warp evaluates the constants itself, the eye-0 warp of chunk 0 publishes them
                const mdvt_conv_frame *fc = frames + frame2;
                float M[12];
#pragma unroll
                for (int k = 0; k < 12; ++k) M[k] = __ldg(&fc->view[e].M[k]);
                const RayView rv = make_ray_view(__ldg(&fc->fx), __ldg(&fc->fy), __ldg(&fc->cx), __ldg(&fc->cy), 1.0f, 1.0f, M, __ldg(&fc->view[e].fx),
                                                 __ldg(&fc->view[e].fy), __ldg(&fc->view[e].cx), __ldg(&fc->view[e].cy));
                const double Bv = rv.B[1];
                k1 = (double)rv.A[2] / Bv; k0 = (double)rv.A[1] / Bv; c1 = (double)rv.C[2] / Bv; c0 = (double)rv.C[1] / Bv;
                const Staircase top = staircase_of(rv, 0, width), bottom = staircase_of(rv, height - 1, width);
                smin = (float)top.smin * 0.999999f; smax = (float)top.smax * 1.000001f;  // the step does not depend on the row
                const bool bad = !staircase_ok(top) || !staircase_ok(bottom) || rv.B[0] != 0.0f || rv.B[2] != 0.0f || rv.T[1] != 0.0f || rv.T[2] != 0.0f;
                if (bad && status && lane == 0) status[frame2] = 1;
                if (item < 2 && lane == 0) {  // items 0 / 1: eye 0 / 1 of the first chunk (always present)
                    sf->k1[e] = k1; sf->k0[e] = k0; sf->c1[e] = c1; sf->c0[e] = c0; sf->smin[e] = smin; sf->smax[e] = smax;
                }
            }
            const double rd = (double)r2;
            const float alpha = (float)(rd * k1 - k0), beta = (float)((rd * c1 - c0) - rd);
            const bool in = s < nsub;
            const float da = __fmaf_rn(alpha, (float)(s * kSub), beta), db = __fmaf_rn(alpha, (float)(s * kSub + kSub - 1), beta);
            const float dlo = fminf(da, db), dhi = fmaxf(da, db);
            const float ipd = rintf(0.5f * (da + db));
            const bool certain = fmaxf(fabsf(ipd - da), fabsf(ipd - db)) * smax < kCertainF;
            const bool up = (ipd + 1.0f - dhi) * smin <= kPossibleF;    // row ip + 1 can round into r somewhere in the sub-block
            const bool down = (dlo - (ipd - 1.0f)) * smin <= kPossibleF;
            const int ip = r2 + (int)ipd;
            const bool ip_ok = ip >= 0 && ip < height;
            const int prim = !in ? -2 : ((ip_ok && certain) ? ip : -1);   // source row of the virtual row, -1: none
            if (in && up && down && status) status[frame2] = 1;  // cannot happen inside the limits (kMaxSubSpan)
            uint8_t *raw = smem + L.raw_off + buf * L.raw_stride;
            uint8_t *vd = raw + e * 2 * row_bytes, *vc = vd + row_bytes;
            const uint8_t *dframe = depth_rgb + (int64_t)frame2 * height * row_bytes;
            const uint8_t *cframe = colour_rgb + (int64_t)frame2 * height * row_bytes;
            // alternates: the uncertain predicted row, then the neighbour; slots through one shared-memory atomic per item
            uint32_t *alt = reinterpret_cast<uint32_t *>(smem + L.alt_off + buf * L.alt_stride);
            uint8_t *altdata = smem + L.altdata_off + buf * L.altdata_stride;
            const int ia = ip + (up ? 1 : -1);
            const bool alt_p = in && ip_ok && !certain, alt_n = in && (up || down) && ia >= 0 && ia < height;
            const uint32_t mp = __ballot_sync(0xFFFFFFFFu, alt_p), mn = __ballot_sync(0xFFFFFFFFu, alt_n);
            int n_staged = 0;
            if (mp | mn) {
                uint32_t base = 0;
                if (lane == 0) base = atoms_add(smem_addr(alt), (uint32_t)(__popc(mp) + __popc(mn)));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                const uint32_t below = (1u << lane) - 1u;
                const uint32_t slot_p = base + __popc(mp & below), slot_n = base + __popc(mp) + __popc(mn & below);
                auto add = [&](uint32_t slot, int row) {
                    alt[1 + slot] = ((uint32_t)e << 31) | ((uint32_t)s << 16) | (uint32_t)row;
                    if (slot < (uint32_t)kAltSlots) {
                        const int64_t goff = ((int64_t)row * width + (int64_t)s * kSub) * 3;
                        bulk_load(altdata + 96 * slot, dframe + goff, 48u, &bar[buf]);
                        bulk_load(altdata + 96 * slot + 48, cframe + goff, 48u, &bar[buf]);
                    }
                };
                if (alt_p) add(slot_p, ip);
                if (alt_n) add(slot_n, ia);
                const int total = __popc(mp) + __popc(mn);
                n_staged = max(0, min((int)base + total, kAltSlots) - min((int)base, kAltSlots));
            }
            // runs
            int prev = __shfl_up_sync(0xFFFFFFFFu, prim, 1);
            if (lane == 0) prev = -3;
            const uint32_t starts = __ballot_sync(0xFFFFFFFFu, in && prim != prev);
            const int n_loaded = __popc(__ballot_sync(0xFFFFFFFFu, prim >= 0));
            if (prim == -1) {
                uint4 *z4 = reinterpret_cast<uint4 *>(vd + 48 * s);
                z4[0] = z4[1] = z4[2] = make_uint4(0u, 0u, 0u, 0u);
            } else if (prim >= 0 && prim != prev) {
                const uint32_t above = lane == 31 ? 0u : (starts >> (lane + 1));
                int len = above ? __ffs(above) : 32 - lane;
                len = min(len, nsub - s);
                const int64_t goff = ((int64_t)prim * width + (int64_t)s * kSub) * 3;
                bulk_load(vd + 48 * s, dframe + goff, 48u * (uint32_t)len, &bar[buf]);
                bulk_load(vc + 48 * s, cframe + goff, 48u * (uint32_t)len, &bar[buf]);
            }
            __syncwarp();
            if (lane == 0) mbar_expect_tx(&bar[buf], 96u * (uint32_t)(n_loaded + n_staged));
        }
    };
    auto publish_frame = [&]() {  // after the items of a row: the constants of its frame are in place for the following rows
        if (tid == 0) s_stair[frame2 & 1].frame = frame2;
    };

    // prologue: the first two rows
    prepare_items(0);
    __syncthreads();
    publish_frame();
    __syncthreads();
    if (unit_begin + 1 < unit_end) {
        if (++r2 == height) { r2 = 0; ++frame2; }
        prepare_items(1);
        __syncthreads();
        publish_frame();
    }
    __syncthreads();

    const uint32_t sm = smem_addr(smem);
    const uint32_t byte0 = 3u * tid;
    const uint32_t shift = (byte0 & 3u) * 8u;  // loop-invariant: the column step T moves 3T bytes, a multiple of 4
    const uint32_t dp_off = byte0 & ~3u;
    const uint32_t flagged_fill = fill_rgb | 0xFF000000u;
    const uint32_t bg_match = (collide & 1) ? bg_rgb : kEmpty32;
    const int dbg = collide >> 8;  // development switch (MDVT_VROWS_SKIP): 1 = no pass 2, 2 = no phase B, 4 = no pass 1 reductions
    const uint4 empty4 = make_uint4(kEmpty32, kEmpty32, kEmpty32, kEmpty32), fill4 = make_uint4(flagged_fill, flagged_fill, flagged_fill, flagged_fill);
    const uint32_t zp_a = sm + L.zp_off;                       // z plane of eye 0; eye 1 at + 4 plane
    const uint32_t cp_delta = (uint32_t)(L.cp_off - L.zp_off);  // colour plane of the same slot
    const uint32_t eye_stride = 4u * (uint32_t)plane;
    const uint32_t w32 = (uint32_t)width;
    const uint32_t bar_a = sm;  // the two mbarriers
    const uint32_t out_a = sm + L.out_off, mask_a = sm + L.mask_off;
    const float fj0 = __int2float_rn(tid);
    int ray_frame = -1;
    EyeConsts ec[2];
    float dec16 = 0.0f, neg_bias = 0.0f, depth_scale = 0.0f, near_plane = 0.0f;

    int it = 0;
    for (int unit = unit_begin; unit < unit_end; ++unit, ++it) {
        const int buf = it & 1;
        if (frame != ray_frame) {  // once or twice per CTA: every thread evaluates both eyes' coefficients itself (float64, ~200 instructions)
            const mdvt_conv_frame *fc = frames + frame;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float M[12];
#pragma unroll
                for (int k = 0; k < 12; ++k) M[k] = __ldg(&fc->view[e].M[k]);
                const RayView rv = make_ray_view(__ldg(&fc->fx), __ldg(&fc->fy), __ldg(&fc->cx), __ldg(&fc->cy), 1.0f, 1.0f, M,
                                                 __ldg(&fc->view[e].fx), __ldg(&fc->view[e].fy), __ldg(&fc->view[e].cx), __ldg(&fc->view[e].cy));
                ec[e].Au = rv.A[0]; ec[e].Av = rv.A[1]; ec[e].Az = rv.A[2];
                ec[e].Cu = rv.C[0]; ec[e].Cv = rv.C[1]; ec[e].Cz = rv.C[2];
                ec[e].Bv = rv.B[1]; ec[e].Tu = rv.T[0];
            }
            dec16 = __fmul_rn(__ldg(&fc->dec_const), 65536.0f);  // exact: fl32(c16 << 16) * dec == fl32(c16) * dec16
            neg_bias = -__fmul_rn(kVMagicInt, dec16);
            depth_scale = __ldg(&fc->depth_scale);
            near_plane = __ldg(&fc->near_plane);
            ray_frame = frame;
        }
        const uint32_t raw_a = sm + L.raw_off + buf * L.raw_stride;
        const uint32_t bd0 = raw_a + dp_off, bd1 = bd0 + 2u * row_bytes;   // this thread's first pixel in the depth rows of eye 0 / 1
        const uint32_t alt_a = sm + L.alt_off + buf * L.alt_stride;
        const uint8_t *altdata = smem + L.altdata_off + buf * L.altdata_stride;
        mbar_wait_a(bar_a + 8u * (uint32_t)buf, (uint32_t)((it >> 1) & 1));

        // ---- pass 1: every candidate -> z plane ------------------------------------------------------------
        uint32_t c_addr[CPT][2], c_z[CPT][2];  // byte address of the candidate's z-plane slot, key bits of its Zv
        auto decode_z = [&](uint32_t lo, uint32_t hi) {
            const uint32_t px = __funnelshift_r(lo, hi, shift);           // [R, G, B, next]
            const uint32_t t = __byte_perm(px, 0x4B000000u, 0x7402);      // 0x4B00RRBB: float value 2^23 + code16
            return __fmul_rn(__fmaf_rn(__uint_as_float(t), dec16, neg_bias), depth_scale);
        };
        constexpr int NB = CPT % 3 == 0 ? 3 : (CPT % 2 == 0 ? 2 : CPT);  // columns per batch: all shared-memory loads first, reductions last
#pragma unroll
        for (int n0 = 0; n0 < CPT; n0 += NB) {
            uint32_t dlo[NB][2], dhi[NB][2];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const uint32_t o = (uint32_t)((n0 + b) * 3 * T);
                dlo[b][0] = lds32(bd0 + o); dhi[b][0] = lds32(bd0 + o + 4u);
                dlo[b][1] = lds32(bd1 + o); dhi[b][1] = lds32(bd1 + o + 4u);
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int n = n0 + b;
                const float fj = __fadd_rn(fj0, (float)(n * T));  // exact
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float z = decode_z(dlo[b][e], dhi[b][e]);
                    float Zv, rz;
                    uint32_t slot = vrow_project(z, fj, ec[e], near_plane, w32, c_z[n][e], Zv, rz);
                    if (GUARD) slot = (tid + n * T < width) ? slot : w32;
                    c_addr[n][e] = zp_a + (uint32_t)e * eye_stride + 4u * slot;
                }
            }
            if (!(dbg & 4)) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                reds_min(c_addr[n0 + b][0], c_z[n0 + b][0]);
                reds_min(c_addr[n0 + b][1], c_z[n0 + b][1]);
            }
            }
        }
        // alternates: half a warp per sub-block and source row, with the exact row test; the first kAltSlots from shared
        // memory, the rest (strong rotations only) from global memory
        const int n_alt = (int)lds32(alt_a);
        auto alternate = [&](int a, uint32_t &addr, uint32_t &zb, uint32_t &col) {
            const uint32_t ent = lds32(alt_a + 4u + 4u * (uint32_t)a);
            const int e = (int)(ent >> 31), s = (int)((ent >> 16) & 0x7FFFu), ia = (int)(ent & 0xFFFFu);
            const int j = s * kSub + (lane & 15);
            uint32_t red, blue;
            if (a < kAltSlots) {
                const uint8_t *d = altdata + 96 * a + 3 * (lane & 15);
                red = d[0]; blue = d[2];
                col = (uint32_t)d[48] | ((uint32_t)d[49] << 8) | ((uint32_t)d[50] << 16);
            } else {
                const int64_t off = ((int64_t)frame * height * width + (int64_t)ia * width + j) * 3;
                red = __ldg(depth_rgb + off); blue = __ldg(depth_rgb + off + 2);
                col = (uint32_t)__ldg(colour_rgb + off) | ((uint32_t)__ldg(colour_rgb + off + 1) << 8) | ((uint32_t)__ldg(colour_rgb + off + 2) << 16);
            }
            const uint32_t t = 0x4B000000u | (red << 8) | blue;
            const float z = __fmul_rn(__fmaf_rn(__uint_as_float(t), dec16, neg_bias), depth_scale);
            const EyeConsts &k = e ? ec[1] : ec[0];
            const float fj = __int2float_rn(j);
            float Zv, rz;
            uint32_t slot = vrow_project(z, fj, k, near_plane, w32, zb, Zv, rz);
            if (vrow_target_row(z, fj, k, (float)ia, Zv, rz) != r) slot = w32;
            addr = zp_a + (uint32_t)e * eye_stride + 4u * slot;
        };
        const int a0 = 2 * warp + (lane >> 4);  // this half-warp's first alternate stays in registers for pass 2, later ones are re-evaluated
        uint32_t a0_addr = zp_a + 4u * w32, a0_z = kEmpty32, a0_col = kEmpty32;
        if (a0 < n_alt) {
            alternate(a0, a0_addr, a0_z, a0_col);
            reds_min(a0_addr, a0_z);
        }
        for (int a = a0 + 2 * kWarps; a < n_alt; a += 2 * kWarps) {
            uint32_t addr, zb, col;
            alternate(a, addr, zb, col);
            reds_min(addr, zb);
        }
        named_barrier(1, T);  // (A) the z planes hold the nearest Zv of every slot
        if (tid == 0) {
            sts32(alt_a, 0u);      // everybody has read this row's alternate count: the list is free for the row after next
            sts32(sm + 48u, 0u);   // phase B's work counter
        }

        // ---- pass 2: the candidates that hold a slot's minimum -> colour plane ------------------------------------
        // (the colours are read here, not in pass 1: fewer registers live across the barrier; a candidate parked on the dummy
        //  slot may write its colour there, nobody reads it)
        if (!(dbg & 1))
#pragma unroll
        for (int n0 = 0; n0 < CPT; n0 += NB) {
            uint32_t zwin[NB][2], clo[NB][2], chi[NB][2];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const uint32_t o = (uint32_t)((n0 + b) * 3 * T) + row_bytes;
                zwin[b][0] = lds32(c_addr[n0 + b][0]); zwin[b][1] = lds32(c_addr[n0 + b][1]);
                clo[b][0] = lds32(bd0 + o); chi[b][0] = lds32(bd0 + o + 4u);
                clo[b][1] = lds32(bd1 + o); chi[b][1] = lds32(bd1 + o + 4u);
            }
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const uint32_t col = __funnelshift_r(clo[b][e], chi[b][e], shift) & 0xFFFFFFu;
                    reds_min(c_addr[n0 + b][e] + cp_delta, zwin[b][e] == c_z[n0 + b][e] ? col : kEmpty32);
                }
        }
        if (a0 < n_alt && lds32(a0_addr) == a0_z) reds_min(a0_addr + cp_delta, a0_col);
        for (int a = a0 + 2 * kWarps; a < n_alt; a += 2 * kWarps) {
            uint32_t addr, zb, col;
            alternate(a, addr, zb, col);
            if (lds32(addr) == zb) reds_min(addr + cp_delta, col);
        }
        if (tid == 0) bulk_wait_read<0>();  // the previous row's staged output has left shared memory (its store was issued a row ago)
        named_barrier(1, T);  // (B) both planes final; this row's virtual rows and alternates are dead
        if (unit + 2 < unit_end) {  // ... so the row after next is prepared into the same buffers: its loads have a full row time
            if (++r2 == height) { r2 = 0; ++frame2; }
            prepare_items(buf);
        }

        // ---- phase B: planes -> colours, hole mask, depth; planes re-armed ---------------------------------------------
        {
            const int groups = width / 4;  // per eye; item k < 2 groups: eye e = k / groups, 4 consecutive target pixels
            constexpr int mwpg = MASK_MODE == 2 ? 3 : 1;
            // 32 groups at a time, handed out through a shared-memory counter: the warps that prepared a row item above join later
            for (; !(dbg & 2);) {
                int k = 0;
                if (lane == 0) k = (int)atoms_add(sm + 48u, 32u);
                k = __shfl_sync(0xFFFFFFFFu, k, 0) + lane;
                if (k - lane >= 2 * groups) break;
                if (k >= 2 * groups) continue;
                const int e = k >= groups;
                const uint32_t za = zp_a + 16u * (uint32_t)k + (e ? 16u : 0u);  // eye 1's plane starts 4 slots (the dummy tail) later
                const uint4 c4 = lds128(za + cp_delta);
                if (out_depth) {
                    const uint4 z4 = lds128(za);
                    float4 d;
                    d.x = z4.x == kEmpty32 ? 0.0f : __uint_as_float(z4.x);
                    d.y = z4.y == kEmpty32 ? 0.0f : __uint_as_float(z4.y);
                    d.z = z4.z == kEmpty32 ? 0.0f : __uint_as_float(z4.z);
                    d.w = z4.w == kEmpty32 ? 0.0f : __uint_as_float(z4.w);
                    reinterpret_cast<float4 *>(out_depth + (int64_t)unit * 2 * width)[k] = d;
                }
                sts128(za + cp_delta, fill4);
                sts128(za, empty4);
                const uint32_t p0 = c4.x == bg_match ? flagged_fill : c4.x;  // an empty slot already holds the flagged fill colour
                const uint32_t p1 = c4.y == bg_match ? flagged_fill : c4.y;
                const uint32_t p2 = c4.z == bg_match ? flagged_fill : c4.z;
                const uint32_t p3 = c4.w == bg_match ? flagged_fill : c4.w;
                const uint32_t oa = out_a + 12u * (uint32_t)k;  // left | right output row
                sts32(oa, __byte_perm(p0, p1, 0x4210));
                sts32(oa + 4u, __byte_perm(p1, p2, 0x5421));
                sts32(oa + 8u, __byte_perm(p2, p3, 0x6542));
                if (MASK_MODE == 1) {
                    sts32(mask_a + 4u * (uint32_t)k, __byte_perm(__byte_perm(p0, p1, 0x0073), __byte_perm(p2, p3, 0x0073), 0x5410));
                } else if (MASK_MODE == 2) {
                    const uint32_t m0 = (p0 >> 24) ? bg_rgb : 0u, m1 = (p1 >> 24) ? bg_rgb : 0u, m2 = (p2 >> 24) ? bg_rgb : 0u, m3 = (p3 >> 24) ? bg_rgb : 0u;
                    const uint32_t ma = mask_a + 4u * (uint32_t)(mwpg * k);
                    sts32(ma, __byte_perm(m0, m1, 0x4210));
                    sts32(ma + 4u, __byte_perm(m1, m2, 0x5421));
                    sts32(ma + 8u, __byte_perm(m2, m3, 0x6542));
                }
            }
        }
        fence_async_smem();
        named_barrier(1, T);  // (C) staged row complete
        if (unit + 2 < unit_end) publish_frame();
        if (tid == 0) {
            bulk_store(out_sbs + (int64_t)unit * 2 * row_bytes, smem + L.out_off, 2 * row_bytes);
            if (MASK_MODE != 0) bulk_store(out_mask + (int64_t)unit * 2 * width * mask_bpp, smem + L.mask_off, 2 * width * mask_bpp);
            bulk_commit();
        }
        if (++r == height) { r = 0; ++frame; }
    }
    if (tid == 0) bulk_wait_all<0>();
}

}  // namespace mdvt

using namespace mdvt;

// Host check of the geometric limits of the kernel above (per frame: both eyes, the extreme target rows; the staircase
// slope is linear in the row and the step monotone in the column, so the extremes bound every row).
extern "C" int mdvt_stereo_conv_vrows_supported(const mdvt_conv_frame *frames_host, int n_frames, int width, int height) {
    if (!frames_host || n_frames < 0 || width <= 0 || height <= 0) return 0;
    if (width % 32 != 0 || width > 3840 || height > 0xFFFE) return 0;
    for (int f = 0; f < n_frames; ++f) {
        const mdvt_conv_frame &fc = frames_host[f];
        if (!(fc.near_plane >= 0.0f)) return 0;
        for (int e = 0; e < 2; ++e) {
            const mdvt_view &vw = fc.view[e];
            const RayView rv = make_ray_view(fc.fx, fc.fy, fc.cx, fc.cy, 1.0f, 1.0f, vw.M, vw.fx, vw.fy, vw.cx, vw.cy);
            if (rv.B[0] != 0.0f || rv.B[2] != 0.0f || rv.T[1] != 0.0f || rv.T[2] != 0.0f || !(rv.B[1] > 0.0f)) return 0;
            if (!staircase_ok(staircase_of(rv, 0, width)) || !staircase_ok(staircase_of(rv, height - 1, width))) return 0;
        }
    }
    return 1;
}

static int vrows_dbg() {
    static const int v = getenv("MDVT_VROWS_SKIP") ? atoi(getenv("MDVT_VROWS_SKIP")) : 0;
    return v;
}
static bool vrows_t384() {  // development switch: MDVT_VROWS_T=384 -> 384 threads x 5 columns instead of 320 x 6 at widths up to 1920
    static const bool v = getenv("MDVT_VROWS_T") && atoi(getenv("MDVT_VROWS_T")) == 384;
    return v;
}

extern "C" int mdvt_stereo_conv_vrows(const uint8_t *depth_rgb, const uint8_t *colour_rgb, int n_frames, int width, int height,
                                      const mdvt_conv_frame *frames_dev, uint32_t bg_rgb, uint32_t fill_rgb, uint32_t flags, uint8_t *out_sbs,
                                      uint8_t *out_mask, float *out_depth, int32_t *status_dev, void *stream) {
    MDVT_REQUIRE(n_frames >= 0, "negative frame count");
    MDVT_REQUIRE(width > 0 && height > 0, "bad frame size %dx%d", width, height);
    if (width % 32 != 0 || width > 3840 || height > 0xFFFE) {
        set_error("mdvt_stereo_conv_vrows takes widths that are multiples of 32 up to 3840 and at most 65534 rows (got %dx%d)", width, height);
        return MDVT_ERR_UNSUPPORTED;
    }
    if (n_frames == 0) return MDVT_OK;
    MDVT_REQUIRE(depth_rgb && colour_rgb && frames_dev && out_sbs, "NULL buffer");
    MDVT_REQUIRE((int64_t)n_frames * height <= 0x7FFFFFFFll, "too many rows in one batch");
    auto aligned16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (!aligned16(depth_rgb) || !aligned16(colour_rgb) || !aligned16(out_sbs) || (out_mask && !aligned16(out_mask)) ||
        (out_depth && !aligned16(out_depth))) {
        set_error("mdvt_stereo_conv_vrows moves rows with cp.async.bulk: every buffer must be 16-byte aligned");
        return MDVT_ERR_UNSUPPORTED;
    }
    const int mode = !out_mask ? 0 : ((flags & MDVT_FLAG_MASK_RGB) ? 2 : 1);
    const VrowSmem L = vrow_smem_layout(width, mode == 2 ? 3 : 1);
    int dev = 0, smem_optin = 0;
    MDVT_CUDA_TRY(cudaGetDevice(&dev));
    MDVT_CUDA_TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (L.total > smem_optin) {
        set_error("row of width %d needs %d bytes of shared memory, device offers %d", width, L.total, smem_optin);
        return MDVT_ERR_UNSUPPORTED;
    }
    const int n_units = n_frames * height;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LAUNCH_TC(M, TT, CC, GG)                                                                                                          \
    do {                                                                                                                              \
        auto kernel = stereo_conv_vrows_kernel<M, TT, CC, GG>;                                                                            \
        MDVT_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));                            \
        int ctas = 0;                                                                                                                 \
        MDVT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kernel, TT, L.total));                                \
        if (ctas < 1) ctas = 1;                                                                                                       \
        int grid = sm_count() * ctas;                                                                                                 \
        if (grid > n_units) grid = n_units;                                                                                           \
        const int per = (n_units + grid - 1) / grid;                                                                                  \
        grid = (n_units + per - 1) / per; /* no empty CTAs */                                                                         \
        kernel<<<grid, TT, L.total, st>>>(depth_rgb, colour_rgb, n_units, width, height, frames_dev, bg_rgb & 0xFFFFFF,          \
                                               fill_rgb & 0xFFFFFF, ((flags & MDVT_FLAG_BG_COLLIDE) ? 1 : 0) | (vrows_dbg() << 8), out_sbs, out_mask, out_depth, \
                                               status_dev);                                                                           \
    } while (0)
#define LAUNCH_G(M, TT, CC)                                           \
    do {                                                              \
        if (width == TT * CC) LAUNCH_TC(M, TT, CC, false);            \
        else LAUNCH_TC(M, TT, CC, true);                              \
    } while (0)
#define LAUNCH_M(M)                                                   \
    do {                                                              \
        if (width <= 640) LAUNCH_G(M, 320, 2);                        \
        else if (width <= 1280) LAUNCH_G(M, 320, 4);                  \
        else if (width <= 1920 && vrows_t384()) LAUNCH_G(M, 384, 5);  \
        else if (width <= 1920) LAUNCH_G(M, 320, 6);                  \
        else LAUNCH_G(M, 640, 6);                                     \
    } while (0)
    if (mode == 0) LAUNCH_M(0);
    else if (mode == 1) LAUNCH_M(1);
    else LAUNCH_M(2);
#undef LAUNCH_M
#undef LAUNCH_G
#undef LAUNCH_TC
    MDVT_CUDA_TRY(cudaGetLastError());
    return MDVT_OK;
}
