// Stereo with a convergence rotation, "virtual source rows" form (stereo_rerender.py:704-725,831-836 with
// --convergence_file, the way movie_2_3D drives it): ONE fused kernel per batch of frames, no global z-buffer, the
// structure of the row-local kernel (mdvt_stereo_rows.cu) kept as far as the geometry allows.
//
// Each eye pose is a rotation about the camera's y axis followed by a shift along x.  In ray form (mdvt_common.cuh)
// that leaves B_u = B_z = T_v = T_z = 0: Zv = z * r_z(j), u' = (z r_u(j) + T_u) / Zv, and v' = r_v(i, j) / r_z(j) does
// not depend on the depth.  The source row that lands in target row r is, per source column j,
//     i*(j) = ((r A_z - A_v) j + (r C_z - C_v)) / B_v            -- LINEAR in j, |slope| ~ sin(theta) |r - cy| / fx,
// so along a target row the source row index is a staircase with a handful of steps.
//   * A unit of work is one target row of ONE eye.  CTAs of 5 compute warps + 2 producer warps, 4 per SM at 1080p, each
//     walking a contiguous block of units.
//   * A PRODUCER warp (one lane per staircase step, float64; one warp per eye, because assembling a row of many steps takes
//     a single warp longer than the compute warps need for it) has the TMA engine assemble the unit's VIRTUAL SOURCE ROW in
//     shared memory: for every stretch of columns whose source row is certain, one cp.async.bulk of that row's depth bytes and
//     one of its colour bytes at the columns' natural offsets (16-pixel sub-blocks of 48 bytes are the granularity).  "Certain"
//     means the prediction |v' - r| < 0.494 holds with a margin the float32 chain cannot eat (it is within 2e-3 px of the
//     prediction).  The narrow zones around the steps of the staircase -- about 1 % of the pixels at a convergence distance of
//     5 m -- get zero depth bytes in the virtual row instead, and both source rows that can round into r there go on a list
//     of ALTERNATE sub-blocks (48 + 48 bytes each, also fetched by TMA).  The producer runs two units ahead of the compute
//     warps, so a unit's loads are issued a full unit time before their first use.
//   * The COMPUTE warps treat the virtual row like the row-local kernel treats a real one: lane-strided columns, word loads,
//     decode, the exact float32 arithmetic of the generic path for Zv and u' (mdvt_splat.cu: FFMA chain, refined reciprocal,
//     correctly rounded quotient, magic-number rounding; v' is not computed at all for certain columns), conflict-free
//     shared-memory reductions.  Alternates additionally go through the exact v' arithmetic and must round to r.
//   * Visibility: nearest Zv wins, candidates with bit-identical Zv are ordered by packed colour -- the colour-keyed order
//     of the generic frame loop (mdvt_render_views), so the results are bit-identical to it (asserted by the tests on every
//     byte, mask and depth plane).  Shared memory has no 64-bit atomic min, so the 55-bit key is split over two 32-bit
//     planes and two passes: (1) every candidate: ATOMS.MIN of float_bits(Zv) into the z plane; barrier; (2) every
//     candidate whose Zv is the slot's minimum: ATOMS.MIN of its colour into the colour plane.  A thread keeps its
//     candidates (slot address, Zv bits) in registers between the passes and reads the colour in pass 2.
//   * Phase B reads the colour plane four pixels at a time (an empty slot already holds the flagged fill colour), re-arms both
//     planes and packs RGB / mask bytes into a staging row that leaves through a TMA bulk store.
// HBM traffic is the algorithmic 14 B/px (each source row is read by ~2-3 neighbouring target rows of the same CTA and by
// both eyes: L2 hits).  Geometry outside the limits below (checked on the host by mdvt_stereo_conv_vrows_supported) goes
// through the generic frame loop instead.
// Measured on the B200 (32 frames per launch): 1080p 12.5 us per frame at a convergence distance of 5 m, 14.2 at 2 m, 17.6 at
// 1 m (generic two-lane loop: 17.4; round 1's target-row kernel: 22.6); 4K 51.6 us (generic loop 89.3); profiles/r02_vrows_*.
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "mdvt_common.cuh"

#ifndef MDVT_VROWS_NB
#define MDVT_VROWS_NB 4
#endif
#ifndef MDVT_VROWS_NP
#define MDVT_VROWS_NP 2     // producer warps per CTA (1 or 2)
#endif
#ifndef MDVT_VROWS_MINB
#define MDVT_VROWS_MINB 4   // resident CTAs per SM the register allocation aims at (tuning aid)
#endif

namespace mdvt {

constexpr int kSub = 16;            // pixels per sub-block: 48 bytes of u8x3, the TMA granularity of the assembly
#ifndef MDVT_VROWS_ALTSLOTS
#define MDVT_VROWS_ALTSLOTS 24
#endif
constexpr int kAltSlots = MDVT_VROWS_ALTSLOTS;       // alternate sub-blocks with their data staged in shared memory (more are read from global memory)
constexpr uint32_t kEmpty32 = 0xFFFFFFFFu;
constexpr float kVMagic = 12582912.0f;  // 1.5 * 2^23
constexpr int kVMagicBits = 0x4B400000;
constexpr float kVMagicInt = 8388608.0f;
constexpr double kStepMin = 0.9, kStepMax = 1.12, kMaxSubSpan = 0.4;

struct VrowSmem {
    int alt_off, alt_stride;      // [2 buffers][1 + 4 nsub] u32: count, then sub-block << 16 | source row
    int altdata_off, altdata_stride;  // [2 buffers][kAltSlots][96]: depth 48 | colour 48
    int raw_off, raw_stride;      // [2 buffers][depth 3W | colour 3W]: the virtual source row of one eye
    int zp_off, cp_off;           // [W + 4] u32 each
    int out_off, mask_off, total; // staging of the eye's half of the output row and of its mask row
};

__host__ __device__ inline VrowSmem vrow_smem_layout(int width, int mask_bpp) {
    VrowSmem L;
    const int nsub = width / kSub;
    int off = 64 + 4 * 96 + 64;  // [0,16): two mbarriers; [64, 448): two StairFrame slots per producer warp; [448, 512): both eyes' EyeConsts
    L.alt_off = off;  L.alt_stride = (1 + 4 * nsub) * 4;       off += 2 * L.alt_stride;
    off = (off + 15) & ~15;
    L.altdata_off = off; L.altdata_stride = kAltSlots * 96;    off += 2 * L.altdata_stride;
    L.raw_off = off;  L.raw_stride = 6 * width;                off += 2 * L.raw_stride;
    L.zp_off = off;   off += (width + 4) * 4;
    L.cp_off = off;   off += (width + 4) * 4;
    L.out_off = off;  off += 3 * width;
    L.mask_off = off; off += width * mask_bpp;
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
template <bool CULL = true>
__device__ __forceinline__ uint32_t vrow_project(float z, float fj, const EyeConsts &k, float near_plane, uint32_t width, uint32_t &zbits,
                                                 float &Zv, float &rz) {
    const float cju = __fmaf_rn(k.Au, fj, k.Cu), cjz = __fmaf_rn(k.Az, fj, k.Cz);
    const float nu = __fmaf_rn(z, cju, k.Tu);
    Zv = __fmaf_rn(z, cjz, 0.0f);
    rz = rcp_refined(Zv);
    const float u = div_rn_by(nu, Zv, rz);
    const uint32_t ui = (uint32_t)(__float_as_int(__fadd_rn(u, kVMagic)) - kVMagicBits);
    zbits = __float_as_uint(Zv);
    // CULL = false (rows whose near plane lies below the depth of code 1 for every column): only code 0 could fail Zv > near,
    // and that one already leaves the row -- z = 0 makes the reciprocal infinite and u a NaN, whose "integer" is >= width
    return CULL ? min(Zv > near_plane ? ui : 0xFFFFFFFFu, width) : min(ui, width);
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
// bar.sync is the ALIGNED barrier: every thread of a warp must arrive together, so the warp reconverges first (the loops over
// the alternates leave half-warps on different paths)
__device__ __forceinline__ void named_barrier(int id, int count) {
    __syncwarp();
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// Per-frame constants of the staircase (shared memory, two slots keyed by frame parity): i*(j) of target row r is
// alpha j + beta with alpha = r k1 - k0, beta = r c1 - c0 (k1 = A_z / B_v, k0 = A_v / B_v, c1 = C_z / B_v, c0 = C_v / B_v, float64),
// and the row step |B_v / r_z(j)| lies in [smin, smax] for every row.
struct StairFrame {
    double k1[2], k0[2], c1[2], c0[2], m[2];
    int frame, pad;
};

// The virtual source row of one unit (target row r2 of eye e), assembled by ONE warp: see "row preparation" in the kernel below.
// vd: the unit's row buffer (depth bytes, then colour bytes); alt / alt_count_a: its list of alternates (generic pointer /
// shared address of the count, which must be 0 on entry); bar: the unit's mbarrier (one arrival, with the transaction bytes).
template <int SLOTS>
__device__ __forceinline__ void vrow_assemble(const StairFrame *sf, int e, int r2, int width, int height, const uint8_t *dframe, const uint8_t *cframe,
                                              uint8_t *vd, uint32_t *alt, uint32_t alt_count_a, uint8_t *altdata, uint64_t *bar, int lane) {
    const int nsub = width / kSub;
    const uint32_t row_bytes = 3u * width;
    const double rd = (double)r2, m = sf->m[e];
    const double alpha = rd * sf->k1[e] - sf->k0[e], beta = (rd * sf->c1[e] - sf->c0[e]) - rd;
    const double d0 = beta, d1 = alpha * (double)(width - 1) + beta;
    const double dmin = fmin(d0, d1), dmax = fmax(d0, d1);
    const int kA = (int)ceil(dmin - 1.0 + m), kB = (int)floor(dmax - m);   // zones kA .. kB meet the row
    const int K = max(0, kB - kA + 1);
    const bool rising = alpha >= 0.0;
    const bool flat = fabs(alpha) < 1e-12;
    const double inv_alpha = flat ? 0.0 : 1.0 / alpha;
    uint8_t *vc = vd + row_bytes;
    const int k_clean = (int)rint(0.5 * (d0 + d1));  // the one clean row when no zone meets the row
    int tx = 0;            // sub-blocks (48 + 48 bytes) this lane has bulk copies in flight for
    int carry_sb = -1;     // last sub-block of the previous zone (in column order)
    for (int c0 = 0; c0 <= K; c0 += 32) {
        const int c = c0 + lane;
        // zone c (c < K): its index k, its sub-block range [sa, sb]; lane K carries the sentinel sa = nsub
        const int k = rising ? kA + c : kB - c;
        int sa = nsub, sb = nsub;
        if (c < K) {
            if (flat) {
                sa = 0; sb = nsub - 1;
            } else {
                const double ja = ((double)k + m - beta) * inv_alpha, jb = ((double)k + 1.0 - m - beta) * inv_alpha;
                const double jlo = fmin(ja, jb) - 0.25, jhi = fmax(ja, jb) + 0.25;
                // (a zone that numerically just misses the row is clamped onto the edge sub-block: testing that one exactly is harmless)
                sa = (int)fmin(fmax(floor(jlo * (1.0 / kSub)), 0.0), (double)(nsub - 1));
                sb = (int)fmin(fmax(floor(jhi * (1.0 / kSub)), (double)sa), (double)(nsub - 1));
            }
        }
        int prev_sb = __shfl_up_sync(0xFFFFFFFFu, sb, 1);
        if (lane == 0) prev_sb = carry_sb;
        carry_sb = __shfl_sync(0xFFFFFFFFu, sb, 31);
        if (c <= K) {
            // clean stretch before zone c (after the last zone for c == K)
            const int s_first = prev_sb + 1, s_last = sa - 1;
            if (s_last >= s_first) {
                const int krow = K == 0 ? k_clean : (c < K ? (rising ? k : k + 1) : (rising ? kB + 1 : kA));
                const int row = r2 + krow;
                if (row >= 0 && row < height) {
                    const int64_t goff = ((int64_t)row * width + (int64_t)s_first * kSub) * 3;
                    const uint32_t bytes = 48u * (uint32_t)(s_last - s_first + 1);
                    bulk_load(vd + 48 * s_first, dframe + goff, bytes, bar);
                    bulk_load(vc + 48 * s_first, cframe + goff, bytes, bar);
                    tx += s_last - s_first + 1;
                } else {
                    for (int s = s_first; s <= s_last; ++s) {
                        uint4 *z4 = reinterpret_cast<uint4 *>(vd + 48 * s);
                        z4[0] = z4[1] = z4[2] = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            }
            // the zone itself
            if (c < K) {
                for (int s = sa; s <= sb; ++s) {
                    uint4 *z4 = reinterpret_cast<uint4 *>(vd + 48 * s);
                    z4[0] = z4[1] = z4[2] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                    for (int up = 0; up < 2; ++up) {
                        const int row = r2 + k + up;
                        if (row < 0 || row >= height) continue;
                        const uint32_t slot = atoms_add(alt_count_a, 1u);
                        alt[1 + slot] = ((uint32_t)s << 16) | (uint32_t)row;
                        if (slot < (uint32_t)SLOTS) {
                            const int64_t goff = ((int64_t)row * width + (int64_t)s * kSub) * 3;
                            bulk_load(altdata + 96 * slot, dframe + goff, 48u, bar);
                            bulk_load(altdata + 96 * slot + 48, cframe + goff, 48u, bar);
                            ++tx;
                        }
                    }
                }
            }
        }
    }
    tx = __reduce_add_sync(0xFFFFFFFFu, tx);
    __syncwarp();  // every lane's zero bytes and list entries are in place before the arrival that publishes them
    if (lane == 0) mbar_expect_tx(bar, 96u * (uint32_t)tx);
}

// MASK_MODE: 0 none, 1 u8 {0,255}, 2 u8x3 (bg colour / black).  A unit of work is one target row of ONE eye (both eyes of a
// row are consecutive units of the same CTA); T threads, every thread owns the columns tid + n T, n < CPT.  GUARD: W < T * CPT
// (columns past the row are culled).
template <int MASK_MODE, int T, int CPT, bool GUARD>
__global__ void __launch_bounds__(T + 32 * MDVT_VROWS_NP, T <= 192 ? MDVT_VROWS_MINB : 2)
    stereo_conv_vrows_kernel(const uint8_t *__restrict__ depth_rgb, const uint8_t *__restrict__ colour_rgb, int n_units, int width, int height,
                             const mdvt_conv_frame *__restrict__ frames, uint32_t bg_rgb, uint32_t fill_rgb, int collide,
                             uint8_t *__restrict__ out_sbs, uint8_t *__restrict__ out_mask, float *__restrict__ out_depth,
                             int32_t *__restrict__ status) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int mask_bpp = MASK_MODE == 2 ? 3 : 1;
    constexpr int kWarps = T / 32;
    const VrowSmem L = vrow_smem_layout(width, mask_bpp);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);  // bar[0], bar[1]: the two sets of row buffers
    StairFrame *s_stair = reinterpret_cast<StairFrame *>(smem + 64);  // [producer warp][frame parity]
    EyeConsts *s_eye = reinterpret_cast<EyeConsts *>(smem + 448);     // [2]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NP = MDVT_VROWS_NP;
    const bool producer = tid >= T;  // the last NP warps prepare rows (staircase, TMA loads), the others compute
    const uint32_t row_bytes = 3u * width;
    const uint32_t sm = smem_addr(smem);

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
        s_stair[0].frame = s_stair[1].frame = s_stair[2].frame = s_stair[3].frame = -1;
        *reinterpret_cast<uint32_t *>(smem + L.alt_off) = 0u;
        *reinterpret_cast<uint32_t *>(smem + L.alt_off + L.alt_stride) = 0u;
    }
    {   // z plane: all ones; colour plane: the flagged fill colour (any real colour, < 2^24, beats it; a hole reads as the fill)
        uint32_t *zp = reinterpret_cast<uint32_t *>(smem + L.zp_off), *cp = reinterpret_cast<uint32_t *>(smem + L.cp_off);
        for (int k = tid; k < width + 4; k += T + 32 * NP) {
            zp[k] = kEmpty32;
            cp[k] = fill_rgb | 0xFF000000u;
        }
    }
    __syncthreads();

    // contiguous block of units per CTA (an even number, so both eyes of a row stay together): successive rows need almost the
    // same source rows (L2 / TMA locality) and the frame constants change once or twice per CTA
    const int per_cta = (((n_units + gridDim.x - 1) / gridDim.x) + 1) & ~1;
    const int unit_begin = blockIdx.x * per_cta, unit_end = min(n_units, unit_begin + per_cta);
    if (unit_begin >= unit_end) return;
    // unit -> (frame, row, eye); tracked incrementally for the unit being processed and the unit being prepared (two ahead)
    int frame = (unit_begin >> 1) / height, r = (unit_begin >> 1) - frame * height, eye = unit_begin & 1;
    int frame2 = frame, r2 = r, eye2 = eye;
    auto advance = [&](int &f, int &row, int &e) {
        if (e == 0) { e = 1; return; }
        e = 0;
        if (++row == height) { row = 0; ++f; }
    };

    // ---- row preparation (producer warp): one lane per step of the staircase --------------------------------------------
    // delta(j) = i*(j) - r = alpha j + beta' is linear in the column.  Around an integer k, for delta in (k - m, k + m) with
    //     m = min(0.494 / smax, 1 - 0.506 / smin),
    // source row r + k is CERTAIN to round into target row r and no neighbour can (the float32 chain is within 2e-3 px of this
    // prediction; [smin, smax] bounds the row step |B_v / r_z(j)|).  Between those clean stretches lie the ZONES
    // [k + m, k + 1 - m]: there rows r + k and r + k + 1 are both candidates and both get the exact row test.  So the producer
    // enumerates the zones that meet the row -- lane c takes zone c in column order -- turns each into a range of sub-blocks
    // (conservatively: +- 0.25 px), and
    //   * the clean stretch before zone c (lane c; lane K the one after the last zone) becomes ONE bulk copy per array of
    //     source row r + k into the virtual row (zero depth bytes where that row is outside the frame: code 0 never passes
    //     Zv > near >= 0);
    //   * every sub-block of a zone gets zero depth bytes in the virtual row and two entries on the list of alternates.
    // alpha, beta' and the zone borders are evaluated in float64 (a handful of operations per lane).  The warp arrives once on
    // the unit's mbarrier with the sum of its transaction bytes (posted after the copies are issued: the phase cannot complete
    // before that arrival, and tx-counts may run negative meanwhile).
    auto prepare_items = [&](int buf) {  // unit (frame2, r2, eye2) -> buffers `buf`
        const int e = eye2;
        StairFrame *sf = &s_stair[2 * (warp - kWarps) + (frame2 & 1)];
        if (sf->frame != frame2) {  // first unit of a frame in this CTA (always eye 0 or the CTA's very first unit): both eyes at once
            __syncwarp();
            if (lane < 2) {
                const mdvt_conv_frame *fc = frames + frame2;
                float M[12];
#pragma unroll
                for (int k = 0; k < 12; ++k) M[k] = __ldg(&fc->view[lane].M[k]);
                const RayView rv = make_ray_view(__ldg(&fc->fx), __ldg(&fc->fy), __ldg(&fc->cx), __ldg(&fc->cy), 1.0f, 1.0f, M,
                                                 __ldg(&fc->view[lane].fx), __ldg(&fc->view[lane].fy), __ldg(&fc->view[lane].cx), __ldg(&fc->view[lane].cy));
                const double Bv = rv.B[1];
                sf->k1[lane] = (double)rv.A[2] / Bv; sf->k0[lane] = (double)rv.A[1] / Bv;
                sf->c1[lane] = (double)rv.C[2] / Bv; sf->c0[lane] = (double)rv.C[1] / Bv;
                const Staircase top = staircase_of(rv, 0, width), bottom = staircase_of(rv, height - 1, width);  // the step does not depend on the row
                const double m = fmin(0.494 / (top.smax * 1.000001), 1.0 - 0.506 / (top.smin * 0.999999));
                sf->m[lane] = m;
                const bool bad = !staircase_ok(top) || !staircase_ok(bottom) || rv.B[0] != 0.0f || rv.B[2] != 0.0f || rv.T[1] != 0.0f || rv.T[2] != 0.0f ||
                                 !(m > 0.4);
                if (bad && status) status[frame2] = 1;
            }
            __syncwarp();
            if (lane == 0) sf->frame = frame2;
            __syncwarp();
        }
        vrow_assemble<kAltSlots>(sf, e, r2, width, height, depth_rgb + (int64_t)frame2 * height * row_bytes, colour_rgb + (int64_t)frame2 * height * row_bytes,
                      smem + L.raw_off + buf * L.raw_stride, reinterpret_cast<uint32_t *>(smem + L.alt_off + buf * L.alt_stride),
                      sm + L.alt_off + buf * L.alt_stride, smem + L.altdata_off + buf * L.altdata_stride, &bar[buf], lane);
    };
    if (producer) {
        // Two units ahead of the compute warps: the buffers of unit u (virtual row, alternates) are free once every compute warp
        // has passed barrier (B) of unit u, and the unit's loads are issued a full unit time before their first use.  With two
        // producer warps, warp pw owns buffer set pw = the units of eye pw (unit_begin is even) and joins only their barrier (B)
        // -- the compute warps alternate between two barrier ids -- so it has two unit times per row; assembling a row of many
        // steps takes one warp longer than the compute warps need for it.
        const int pw = warp - kWarps;
        if (NP == 2) {
            if (pw == 1) {
                if (unit_begin + 1 >= unit_end) return;
                advance(frame2, r2, eye2);
            }
            prepare_items(pw);
            for (int unit = unit_begin + pw; unit < unit_end; unit += 2) {
                if (pw) named_barrier(3, T + 32);  // (B) of this unit
                else named_barrier(2, T + 32);
                if (unit + 2 < unit_end) {
                    advance(frame2, r2, eye2);
                    advance(frame2, r2, eye2);
                    prepare_items(pw);
                }
            }
            return;
        }
        prepare_items(0);
        if (unit_begin + 1 < unit_end) {
            advance(frame2, r2, eye2);
            prepare_items(1);
        }
        int it = 0;
        for (int unit = unit_begin; unit < unit_end; ++unit, ++it) {
            named_barrier(2, T + 32);  // (B) of this unit
            if (unit + 2 < unit_end) {
                advance(frame2, r2, eye2);
                prepare_items(it & 1);
            }
        }
        return;
    }

    const uint32_t byte0 = 3u * tid;
    const uint32_t shift = (byte0 & 3u) * 8u;  // loop-invariant: the column step T moves 3T bytes, a multiple of 4
    const uint32_t dp_off = byte0 & ~3u;
    const uint32_t flagged_fill = fill_rgb | 0xFF000000u;
    const uint32_t bg_match = (collide & 1) ? bg_rgb : kEmpty32;
    const uint4 empty4 = make_uint4(kEmpty32, kEmpty32, kEmpty32, kEmpty32), fill4 = make_uint4(flagged_fill, flagged_fill, flagged_fill, flagged_fill);
    const uint32_t zp_a = sm + L.zp_off;
    const uint32_t cp_delta = (uint32_t)(L.cp_off - L.zp_off);  // colour plane of the same slot
    const uint32_t w32 = (uint32_t)width;
    const uint32_t bar_a = sm;  // the two mbarriers
    const uint32_t out_a = sm + L.out_off, mask_a = sm + L.mask_off;
    const float fj0 = __int2float_rn(tid);
    int ray_frame = -1;
    float dec16 = 0.0f, neg_bias = 0.0f, depth_scale = 0.0f, near_plane = 0.0f;

    int it = 0;
    for (int unit = unit_begin; unit < unit_end; ++unit, ++it) {
        const int buf = it & 1;
        if (frame != ray_frame) {  // once or twice per CTA: both eyes' float32 coefficients (float64 inside make_ray_view) -> shared memory
            const mdvt_conv_frame *fc = frames + frame;
            dec16 = __fmul_rn(__ldg(&fc->dec_const), 65536.0f);  // exact: fl32(c16 << 16) * dec == fl32(c16) * dec16
            neg_bias = -__fmul_rn(kVMagicInt, dec16);
            depth_scale = __ldg(&fc->depth_scale);
            near_plane = __ldg(&fc->near_plane);
            named_barrier(1, T);  // nobody still reads the previous frame's coefficients
            if (tid < 2) {
                float M[12];
#pragma unroll
                for (int k = 0; k < 12; ++k) M[k] = __ldg(&fc->view[tid].M[k]);
                const RayView rv = make_ray_view(__ldg(&fc->fx), __ldg(&fc->fy), __ldg(&fc->cx), __ldg(&fc->cy), 1.0f, 1.0f, M, __ldg(&fc->view[tid].fx),
                                                 __ldg(&fc->view[tid].fy), __ldg(&fc->view[tid].cx), __ldg(&fc->view[tid].cy));
                EyeConsts k;
                k.Au = rv.A[0]; k.Av = rv.A[1]; k.Az = rv.A[2];
                k.Cu = rv.C[0]; k.Cv = rv.C[1]; k.Cz = rv.C[2];
                k.Bv = rv.B[1]; k.Tu = rv.T[0];
                s_eye[tid] = k;
            }
            named_barrier(1, T);
            ray_frame = frame;
        }
        const EyeConsts ec = s_eye[eye];
        const uint32_t raw_a = sm + L.raw_off + buf * L.raw_stride;
        const uint32_t bd = raw_a + dp_off;   // this thread's first pixel in the depth row
        const uint32_t alt_a = sm + L.alt_off + buf * L.alt_stride;
        const uint8_t *altdata = smem + L.altdata_off + buf * L.altdata_stride;
        mbar_wait_a(bar_a + 8u * (uint32_t)buf, (uint32_t)((it >> 1) & 1));
        // The unit's alternates (sub-block << 16 | source row): the first kAltSlots have their bytes staged in shared memory, the
        // rest come from global memory (L2: the rows were just copied for the neighbouring units).  Each half-warp fetches its
        // first alternate HERE, ahead of pass 1, so that the latency hides behind the main columns, and keeps it in registers for
        // pass 2; later ones (strong rotations only) are re-evaluated there.
        const int n_alt = (int)lds32(alt_a);
        auto alt_fetch = [&](int a, uint32_t &ent, uint32_t &red, uint32_t &blue, uint32_t &col) {
            ent = lds32(alt_a + 4u + 4u * (uint32_t)a);
            if (a < kAltSlots) {
                const uint8_t *d = altdata + 96 * a + 3 * (lane & 15);
                red = d[0]; blue = d[2];
                col = (uint32_t)d[48] | ((uint32_t)d[49] << 8) | ((uint32_t)d[50] << 16);
            } else {
                const int s = (int)(ent >> 16), ia = (int)(ent & 0xFFFFu);
                const int64_t off = ((int64_t)frame * height * width + (int64_t)ia * width + s * kSub + (lane & 15)) * 3;
                red = __ldg(depth_rgb + off); blue = __ldg(depth_rgb + off + 2);
                col = (uint32_t)__ldg(colour_rgb + off) | ((uint32_t)__ldg(colour_rgb + off + 1) << 8) | ((uint32_t)__ldg(colour_rgb + off + 2) << 16);
            }
        };
        const int a0 = 2 * warp + (lane >> 4);
        uint32_t a0_ent = 0u, a0_red = 0u, a0_blue = 0u, a0_col = kEmpty32;
        if (a0 < n_alt) alt_fetch(a0, a0_ent, a0_red, a0_blue, a0_col);

        // ---- pass 1: every candidate -> z plane ------------------------------------------------------------
        uint32_t c_addr[CPT], c_z[CPT];  // byte address of the candidate's z-plane slot, key bits of its Zv
        constexpr int NB = CPT % MDVT_VROWS_NB == 0 ? MDVT_VROWS_NB : (CPT % 4 == 0 ? 4 : (CPT % 5 == 0 ? 5 : (CPT % 3 == 0 ? 3 : (CPT % 2 == 0 ? 2 : 1))));  // columns per batch: loads first, reductions last
        auto pass1 = [&](auto cull_tag) {
            constexpr bool CULL = decltype(cull_tag)::value;
#pragma unroll
            for (int n0 = 0; n0 < CPT; n0 += NB) {
                uint32_t dlo[NB], dhi[NB];
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const uint32_t o = (uint32_t)((n0 + b) * 3 * T);
                    dlo[b] = dhi[b] = 0u;
                    if (!GUARD || tid + (n0 + b) * T < width) {  // columns past the row: nothing to read (their bytes would lie outside the row buffers)
                        dlo[b] = lds32(bd + o); dhi[b] = lds32(bd + o + 4u);
                    }
                }
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const int n = n0 + b;
                    const float fj = __fadd_rn(fj0, (float)(n * T));  // exact
                    const uint32_t px = __funnelshift_r(dlo[b], dhi[b], shift);           // [R, G, B, next]
                    const uint32_t t = __byte_perm(px, 0x4B000000u, 0x7402);              // 0x4B00RRBB: float value 2^23 + code16
                    const float z = __fmul_rn(__fmaf_rn(__uint_as_float(t), dec16, neg_bias), depth_scale);
                    float Zv, rz;
                    uint32_t slot = vrow_project<CULL>(z, fj, ec, near_plane, w32, c_z[n], Zv, rz);
                    if (GUARD) slot = (tid + n * T < width) ? slot : w32;
                    c_addr[n] = zp_a + 4u * slot;
                }
#pragma unroll
                for (int b = 0; b < NB; ++b) reds_min(c_addr[n0 + b], c_z[n0 + b]);
            }
        };
        // the cull test is needed only when some non-zero code can lie at or below the near plane: r_z(j) is linear in j, so
        // its minimum over the row sits at an end column; the depth of code 1 is dec16 * depth_scale
        const float rz_min = fminf(ec.Cz, __fmaf_rn(ec.Az, (float)(width - 1), ec.Cz));
        if (near_plane < __fmul_rn(__fmul_rn(dec16, depth_scale), rz_min) * 0.999f) pass1(std::false_type{});
        else pass1(std::true_type{});
        // alternates: half a warp per sub-block and source row, with the exact row test
        auto alt_project = [&](uint32_t ent, uint32_t red, uint32_t blue, uint32_t &addr, uint32_t &zb) {
            const int s = (int)(ent >> 16), ia = (int)(ent & 0xFFFFu);
            const int j = s * kSub + (lane & 15);
            const uint32_t t = 0x4B000000u | (red << 8) | blue;
            const float z = __fmul_rn(__fmaf_rn(__uint_as_float(t), dec16, neg_bias), depth_scale);
            const float fj = __int2float_rn(j);
            float Zv, rz;
            uint32_t slot = vrow_project(z, fj, ec, near_plane, w32, zb, Zv, rz);
            if (vrow_target_row(z, fj, ec, (float)ia, Zv, rz) != r) slot = w32;
            addr = zp_a + 4u * slot;
        };
        uint32_t a0_addr = zp_a + 4u * w32, a0_z = kEmpty32;
        if (a0 < n_alt) {
            alt_project(a0_ent, a0_red, a0_blue, a0_addr, a0_z);
            reds_min(a0_addr, a0_z);
        }
        for (int a = a0 + 2 * kWarps; a < n_alt; a += 2 * kWarps) {
            uint32_t ent, red, blue, col, addr, zb;
            alt_fetch(a, ent, red, blue, col);
            alt_project(ent, red, blue, addr, zb);
            reds_min(addr, zb);
        }
        named_barrier(1, T);  // (A) the z plane holds the nearest Zv of every slot
        if (tid == 0) {
            sts32(alt_a, 0u);      // everybody has read this unit's alternate count: the list is free for the unit after next
        }

        // ---- pass 2: the candidates that hold a slot's minimum -> colour plane ------------------------------------
        // (the colours are read here, not in pass 1: fewer registers live across the barrier; a candidate parked on the dummy
        //  slot may write its colour there, nobody reads it)
#pragma unroll
        for (int n0 = 0; n0 < CPT; n0 += NB) {
            uint32_t zwin[NB], clo[NB], chi[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const uint32_t o = (uint32_t)((n0 + b) * 3 * T) + row_bytes;
                zwin[b] = lds32(c_addr[n0 + b]);
                clo[b] = chi[b] = 0u;
                if (!GUARD || tid + (n0 + b) * T < width) {
                    clo[b] = lds32(bd + o); chi[b] = lds32(bd + o + 4u);
                }
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const uint32_t col = __funnelshift_r(clo[b], chi[b], shift) & 0xFFFFFFu;
                reds_min(c_addr[n0 + b] + cp_delta, zwin[b] == c_z[n0 + b] ? col : kEmpty32);
            }
        }
        if (a0 < n_alt && lds32(a0_addr) == a0_z) reds_min(a0_addr + cp_delta, a0_col);
        for (int a = a0 + 2 * kWarps; a < n_alt; a += 2 * kWarps) {
            uint32_t ent, red, blue, col, addr, zb;
            alt_fetch(a, ent, red, blue, col);
            alt_project(ent, red, blue, addr, zb);
            if (lds32(addr) == zb) reds_min(addr + cp_delta, col);
        }
        if (tid == 0) bulk_wait_read<0>();  // the previous unit's staged output has left shared memory (its store was issued a unit ago)
        // (B) both planes final; this unit's virtual row and alternates are dead: its producer reloads them
        if (NP == 2 && buf) named_barrier(3, T + 32);
        else named_barrier(2, T + 32);

        // ---- phase B: planes -> colours, hole mask, depth; planes re-armed ---------------------------------------------
        const int row_unit = unit >> 1;  // frame * height + r
        {
            const int groups = width / 4;  // 4 consecutive target pixels each
            constexpr int mwpg = MASK_MODE == 2 ? 3 : 1;
            float4 *drow = out_depth ? reinterpret_cast<float4 *>(out_depth + ((int64_t)row_unit * 2 + eye) * width) : nullptr;
#pragma unroll 3
            for (int k = tid; k < groups; k += T) {
                const uint32_t za = zp_a + 16u * (uint32_t)k;
                const uint4 c4 = lds128(za + cp_delta);
                if (drow) {
                    const uint4 z4 = lds128(za);
                    float4 d;
                    d.x = z4.x == kEmpty32 ? 0.0f : __uint_as_float(z4.x);
                    d.y = z4.y == kEmpty32 ? 0.0f : __uint_as_float(z4.y);
                    d.z = z4.z == kEmpty32 ? 0.0f : __uint_as_float(z4.z);
                    d.w = z4.w == kEmpty32 ? 0.0f : __uint_as_float(z4.w);
                    drow[k] = d;
                }
                sts128(za + cp_delta, fill4);
                sts128(za, empty4);
                const uint32_t p0 = c4.x == bg_match ? flagged_fill : c4.x;  // an empty slot already holds the flagged fill colour
                const uint32_t p1 = c4.y == bg_match ? flagged_fill : c4.y;
                const uint32_t p2 = c4.z == bg_match ? flagged_fill : c4.z;
                const uint32_t p3 = c4.w == bg_match ? flagged_fill : c4.w;
                const uint32_t oa = out_a + 12u * (uint32_t)k;
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
        named_barrier(1, T);  // (C) staged half row complete
        if (tid == 0) {  // this eye's half of the side-by-side row and of its mask row
            bulk_store(out_sbs + ((int64_t)row_unit * 2 + eye) * row_bytes, smem + L.out_off, row_bytes);
            if (MASK_MODE != 0) bulk_store(out_mask + ((int64_t)row_unit * 2 + eye) * width * mask_bpp, smem + L.mask_off, width * mask_bpp);
            bulk_commit();
        }
        advance(frame, r, eye);
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

static int vrows_t() {  // development switch: MDVT_VROWS_T=192 | 128 -> 6 compute warps x 10 columns / 4 x 15 instead of 5 x 12 at 1920
    static const int v = getenv("MDVT_VROWS_T") ? atoi(getenv("MDVT_VROWS_T")) : 0;
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
    MDVT_REQUIRE((int64_t)n_frames * height * 2 <= 0x7FFFFFFFll, "too many rows in one batch");
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
    const int n_units = n_frames * height * 2;  // (row, eye)
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LAUNCH_TC(M, TT, CC, GG)                                                                                                          \
    do {                                                                                                                              \
        auto kernel = stereo_conv_vrows_kernel<M, TT, CC, GG>;                                                                            \
        MDVT_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));                            \
        int ctas = 0;                                                                                                                 \
        MDVT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kernel, TT + 32 * MDVT_VROWS_NP, L.total));                \
        if (ctas < 1) ctas = 1;                                                                                                       \
        int grid = sm_count() * ctas;                                                                                                 \
        if (grid > n_units) grid = n_units;                                                                                           \
        const int per = (((n_units + grid - 1) / grid) + 1) & ~1; /* both eyes of a row in one CTA */                                 \
        grid = (n_units + per - 1) / per; /* no empty CTAs */                                                                         \
        kernel<<<grid, TT + 32 * MDVT_VROWS_NP, L.total, st>>>(depth_rgb, colour_rgb, n_units, width, height, frames_dev, bg_rgb & 0xFFFFFF,          \
                                               fill_rgb & 0xFFFFFF, (flags & MDVT_FLAG_BG_COLLIDE) ? 1 : 0, out_sbs, out_mask, out_depth, \
                                               status_dev);                                                                           \
    } while (0)
#define LAUNCH_G(M, TT, CC)                                           \
    do {                                                              \
        if (width == TT * CC) LAUNCH_TC(M, TT, CC, false);            \
        else LAUNCH_TC(M, TT, CC, true);                              \
    } while (0)
#define LAUNCH_M(M)                                                   \
    do {                                                              \
        if (width <= 320) LAUNCH_G(M, 160, 2);                        \
        else if (width <= 640) LAUNCH_G(M, 160, 4);                   \
        else if (width <= 1280) LAUNCH_G(M, 160, 8);                  \
        else if (width == 1920 && vrows_t() == 192) LAUNCH_TC(M, 192, 10, false); \
        else if (width == 1920 && vrows_t() == 128) LAUNCH_TC(M, 128, 15, false); \
        else if (width <= 1920) LAUNCH_G(M, 160, 12);                 \
        else LAUNCH_G(M, 320, 12);                                    \
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
