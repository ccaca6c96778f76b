// K4, production kernel: detector planes on an AFFINE pixel grid, voxel
// coordinates in 32-bit fixed point (detector.py:194-244, 278-300).
//
// The reference rotates the P x P coordinate grids three times per
// orientation in fp64 (fma chains) and floor-divides the result.  Doing that
// per pixel costs ~40 fp64 operations and bounds the exact kernel
// (gx_detector.cu) by the fp64 pipe.  But the detector grid is affine in
// (row, col) -- make_detector is a meshgrid of two linspaces and rotations are
// linear -- so the voxel coordinate of pixel (r, c) under orientation o is
//
//        t_a(r, c) = o_a + c u_a + r v_a   (+- E_a),        a = x, y, z
//
// where (o, u, v) interpolate the EXACT corner values of the reference's
// chain (the host runs the chain on the corners) and E_a is a rigorous bound
// on everything the model ignores: the measured deviation of the base grid
// from affinity, the rounding of the three fma chains propagated step by step,
// the rounding of p - qmin, and the fixed-point quantisation.  The kernel
// evaluates t in unsigned fixed point with F fractional bits (adds and shifts
// only), takes floor(t) when t is farther than E from an integer, and sends
// the remaining pixels (a few 1e-5) through the exact fp64 chain.  A
// component that provably has the same p - qmin for every pixel (plane lying
// in a grid-aligned plane, e.g. phi = theta = 0) gets its index from the host.
// Voxel indices are therefore bit-identical to the reference's for every pixel
// and orientation, at ~18 integer instructions per pixel and orientation.
//
// Work decomposition: one CTA of 128 threads per 32 x 16 pixel tile, 2 x 2
// pixels per thread, warps cover 16 x 8 pixel patches (25 pixels share a voxel
// at config 5, so a warp gather touches a handful of sectors and lives on L1
// hits).  Orientations are looped inside the CTA (image touched once);
// per-orientation fixed-point records are staged in shared memory in chunks.
// Small images are additionally split over orientation ranges (grid.y) and
// combined with fp64 atomics so that the grid fills 148 SMs.
#include <math.h>
#include <string.h>
#include "gx_common.cuh"
#include "gx_tma.cuh"

#define GA_TW 32
#define GA_TH 16
#define GA_THREADS 128
#define GA_CHUNK 64

struct AffRecord {              // per orientation, host -> device
    double o[3], u[3], v[3];    // voxel coordinate model, offset included
    int32_t U[3], V[3];         // round(u 2^F), round(v 2^F)
    float w;
    int32_t n_const;            // components whose index the host proved constant
};
static_assert(sizeof(AffRecord) == 104, "AffRecord layout");

struct AffSmem {                // per orientation, per tile
    uint32_t tx, ty, tz;
    float w;
    int32_t ux, uy, uz, pad0;
    int32_t vx, vy, vz, pad1;
};

struct AffLaunch {
    const float *iq_shifted;    // iq - off * (Vx Vz + Vz + 1)
    const float *iq;
    uint32_t Vx, Vy, Vz;
    uint32_t lo, hix, hiy, hiz; // clamp bounds in offset coordinates
    uint32_t shift;             // off * (Vx Vz + Vz + 1)
    int F;
    uint32_t HM, half;
    double scale;               // 2^F
    double b_o[3], b_u[3], b_v[3];   // base grid (q units) for the in-box test
    double rin2;
    double qmin[3], dq, inv_dq;
    const double *px, *py, *pz;
    const double *R27;
    const AffRecord *rec;
    int rows, cols, n_orient, per_split, n_split;
    double *image;
    int probe;
    int64_t *index_out;
    unsigned long long *slow_count;
};

__device__ __forceinline__ void aff_chain(const double *R, double &x, double &y, double &z)
{
    const double a = __fma_rn(R[2], z, __fma_rn(R[1], y, __dmul_rn(R[0], x)));
    const double b = __fma_rn(R[5], z, __fma_rn(R[4], y, __dmul_rn(R[3], x)));
    const double c = __fma_rn(R[8], z, __fma_rn(R[7], y, __dmul_rn(R[6], x)));
    x = a; y = b; z = c;
}

__device__ __forceinline__ uint32_t aff_exact_index(double p, double qmin, double dq, double inv_dq, uint32_t n)
{
    const double q = gx_floordiv(__dsub_rn(p, qmin), dq, inv_dq);
    if (!(q > 0.0)) return 0u;
    if (q >= (double)n) return n - 1u;
    return (uint32_t)q;
}

// the reference's own arithmetic for one pixel and orientation (rare path)
__device__ __noinline__ uint32_t aff_exact_voxel(const AffLaunch &L, int64_t i, int o)
{
    double x = L.px[i], y = L.py[i], z = L.pz[i];
    const double *R = L.R27 + (size_t)o * 27;
    aff_chain(R, x, y, z);
    aff_chain(R + 9, x, y, z);
    aff_chain(R + 18, x, y, z);
    const uint32_t ix = aff_exact_index(x, L.qmin[0], L.dq, L.inv_dq, L.Vx);
    const uint32_t iy = aff_exact_index(y, L.qmin[1], L.dq, L.inv_dq, L.Vy);
    const uint32_t iz = aff_exact_index(z, L.qmin[2], L.dq, L.inv_dq, L.Vz);
    return (iy * L.Vx + ix) * L.Vz + iz + L.shift;
}

__device__ __forceinline__ uint32_t umin3(uint32_t a, uint32_t b, uint32_t c) { return min(min(a, b), c); }

template <bool CLAMP, bool PROBE>
__device__ __forceinline__ void aff_tile(const AffLaunch &L, AffSmem *s_rec, int r0, int c0, int o_begin, int o_end)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cx = 2 * ((warp & 1) * 8 + (lane & 7));
    const int ry = 2 * ((warp >> 1) * 4 + (lane >> 3));
    const int r = r0 + ry, c = c0 + cx;
    const bool live0 = r < L.rows && c < L.cols, live1 = r < L.rows && c + 1 < L.cols;
    const bool live2 = r + 1 < L.rows && c < L.cols, live3 = r + 1 < L.rows && c + 1 < L.cols;
    const int64_t i0 = (int64_t)r * L.cols + c;
    const int F = L.F;
    const uint32_t HM = L.HM, Vx = L.Vx, Vz = L.Vz;
    const float *iqs = L.iq_shifted;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
    unsigned slow = 0;

    for (int o0 = o_begin; o0 < o_end; o0 += GA_CHUNK) {
        const int nc = min(GA_CHUNK, o_end - o0);
        __syncthreads();
        if (tid < nc) {
            const AffRecord *q = L.rec + (o0 + tid);
            AffSmem e;
            const double fr = (double)r0, fc = (double)c0;
            const double t0 = __fma_rn(fr, q->v[0], __fma_rn(fc, q->u[0], q->o[0]));
            const double t1 = __fma_rn(fr, q->v[1], __fma_rn(fc, q->u[1], q->o[1]));
            const double t2 = __fma_rn(fr, q->v[2], __fma_rn(fc, q->u[2], q->o[2]));
            e.tx = (uint32_t)__double2ll_rn(t0 * L.scale) + L.half;
            e.ty = (uint32_t)__double2ll_rn(t1 * L.scale) + L.half;
            e.tz = (uint32_t)__double2ll_rn(t2 * L.scale) + L.half;
            e.w = q->w;
            e.ux = q->U[0]; e.uy = q->U[1]; e.uz = q->U[2]; e.pad0 = 0;
            e.vx = q->V[0]; e.vy = q->V[1]; e.vz = q->V[2]; e.pad1 = 0;
            s_rec[tid] = e;
        }
        __syncthreads();
#pragma unroll 2
        for (int o = 0; o < nc; ++o) {
            const uint4 A = *reinterpret_cast<const uint4 *>(&s_rec[o].tx);
            const uint4 B = *reinterpret_cast<const uint4 *>(&s_rec[o].ux);
            const uint4 C = *reinterpret_cast<const uint4 *>(&s_rec[o].vx);
            const float w = __uint_as_float(A.w);
            const uint32_t x0 = A.x + (uint32_t)cx * B.x + (uint32_t)ry * C.x;
            const uint32_t y0 = A.y + (uint32_t)cx * B.y + (uint32_t)ry * C.y;
            const uint32_t z0 = A.z + (uint32_t)cx * B.z + (uint32_t)ry * C.z;
            const uint32_t x1 = x0 + B.x, y1 = y0 + B.y, z1 = z0 + B.z;
            const uint32_t x2 = x0 + C.x, y2 = y0 + C.y, z2 = z0 + C.z;
            const uint32_t x3 = x2 + B.x, y3 = y2 + B.y, z3 = z2 + B.z;
            uint32_t v0, v1, v2, v3;
#define GA_VOXEL(X, Y, Z, OUT)                                               \
            do {                                                             \
                uint32_t jx = (X) >> F, jy = (Y) >> F, jz = (Z) >> F;        \
                if (CLAMP) {                                                 \
                    jx = min(max(jx, L.lo), L.hix);                          \
                    jy = min(max(jy, L.lo), L.hiy);                          \
                    jz = min(max(jz, L.lo), L.hiz);                          \
                }                                                            \
                OUT = (jy * Vx + jx) * Vz + jz;                              \
            } while (0)
            GA_VOXEL(x0, y0, z0, v0);
            GA_VOXEL(x1, y1, z1, v1);
            GA_VOXEL(x2, y2, z2, v2);
            GA_VOXEL(x3, y3, z3, v3);
#undef GA_VOXEL
            const uint32_t m0 = umin3(x0 & HM, y0 & HM, z0 & HM);
            const uint32_t m1 = umin3(x1 & HM, y1 & HM, z1 & HM);
            const uint32_t m2 = umin3(x2 & HM, y2 & HM, z2 & HM);
            const uint32_t m3 = umin3(x3 & HM, y3 & HM, z3 & HM);
            if (min(min(m0, m1), min(m2, m3)) == 0u) {
                // within the error bound of a voxel edge: the reference's own arithmetic decides
                if (m0 == 0u && live0) { v0 = aff_exact_voxel(L, i0, o0 + o); ++slow; }
                if (m1 == 0u && live1) { v1 = aff_exact_voxel(L, i0 + 1, o0 + o); ++slow; }
                if (m2 == 0u && live2) { v2 = aff_exact_voxel(L, i0 + L.cols, o0 + o); ++slow; }
                if (m3 == 0u && live3) { v3 = aff_exact_voxel(L, i0 + L.cols + 1, o0 + o); ++slow; }
            }
            a0 = fmaf(w, __ldg(iqs + v0), a0);
            a1 = fmaf(w, __ldg(iqs + v1), a1);
            a2 = fmaf(w, __ldg(iqs + v2), a2);
            a3 = fmaf(w, __ldg(iqs + v3), a3);
            if (PROBE && o0 + o == L.probe) {
                if (live0) L.index_out[i0] = (int64_t)(v0 - L.shift);
                if (live1) L.index_out[i0 + 1] = (int64_t)(v1 - L.shift);
                if (live2) L.index_out[i0 + L.cols] = (int64_t)(v2 - L.shift);
                if (live3) L.index_out[i0 + L.cols + 1] = (int64_t)(v3 - L.shift);
            }
        }
        // fp32 partial sums cover at most GA_CHUNK orientations, then widen
        d0 += (double)a0; d1 += (double)a1; d2 += (double)a2; d3 += (double)a3;
        a0 = a1 = a2 = a3 = 0.f;
    }
    if (L.n_split == 1) {
        if (live0) L.image[i0] += d0;
        if (live1) L.image[i0 + 1] += d1;
        if (live2) L.image[i0 + L.cols] += d2;
        if (live3) L.image[i0 + L.cols + 1] += d3;
    } else {
        if (live0) atomicAdd(L.image + i0, d0);
        if (live1) atomicAdd(L.image + i0 + 1, d1);
        if (live2) atomicAdd(L.image + i0 + L.cols, d2);
        if (live3) atomicAdd(L.image + i0 + L.cols + 1, d3);
    }
    if (L.slow_count && slow) atomicAdd(L.slow_count, (unsigned long long)slow);
}

template <bool PROBE>
__global__ void __launch_bounds__(GA_THREADS)
detector_affine_kernel(const __grid_constant__ AffLaunch L)
{
    __shared__ __align__(16) AffSmem s_rec[GA_CHUNK];
    const int tiles_x = (L.cols + GA_TW - 1) / GA_TW;
    const int tile = blockIdx.x;
    const int r0 = (tile / tiles_x) * GA_TH, c0 = (tile % tiles_x) * GA_TW;
    const int o_begin = blockIdx.y * L.per_split;
    const int o_end = min(L.n_orient, o_begin + L.per_split);
    // a tile whose (unclipped) corners all lie within the sphere inscribed in the voxel box
    // can never leave the box under rotation: no clamping needed for any of its pixels
    double rmax2 = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double fr = (double)(r0 + ((k & 2) ? GA_TH - 1 : 0)), fc = (double)(c0 + ((k & 1) ? GA_TW - 1 : 0));
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double p = L.b_o[a] + fc * L.b_u[a] + fr * L.b_v[a];
            s += p * p;
        }
        rmax2 = fmax(rmax2, s);
    }
    if (rmax2 <= L.rin2) aff_tile<false, PROBE>(L, s_rec, r0, c0, o_begin, o_end);
    else aff_tile<true, PROBE>(L, s_rec, r0, c0, o_begin, o_end);
}

// ------------------------------------------------ TMA-brick variant (A/B) ----
// north_star's kernel (4) names "TMA-staged voxel tiles".  For a 32 x 16-pixel tile and one orientation the
// pixels' voxel coordinates span at most 31 |u_a| + 15 |v_a| voxels along axis a (6.7 at config 5), so an
// 8 (y) x 8 (x) x 12 (z) brick of the voxel grid holds every voxel the tile can touch: the z origin is rounded
// down to a multiple of 4 voxels because the innermost start coordinate of a tensor-map box must be 16-byte
// aligned (an unaligned start raises "illegal instruction": found the hard way, tilted orientations only).
// This variant streams those bricks (cp.async.bulk.tensor.3d -> UTMALDG.3D, 3 KB each, mbarrier ring of
// GA_RING slots, thread 0 as producer) and gathers from shared memory; pixels inside the error band still take
// the reference's fp64 chain and read the grid itself, tiles that touch the box boundary run the clamped LDG
// path.  Result identical to the LDG kernel; measured 45 % SLOWER (1.91 vs 1.26 - 1.36 ms per 360 orientations at
// config 5, profiles/r04_summary.md): the LDG gather lives on L1 hits and is bound by instruction issue, and
// the ring adds waits and arrivals to exactly that resource.  Kept as an A/B switch: GIWAXS_B200_DETECTOR_TMA=1.
#define GA_B 8
#define GA_BZ 12                 // innermost (z) extent: the box starts at a multiple of 4 voxels (16 bytes)
#define GA_RING 8
#define GA_BRICK (GA_B * GA_B * GA_BZ)

struct AffBrick { int32_t bz, bx, by; uint32_t oc; };   // TMA coordinates of the brick and its flat origin

__global__ void __launch_bounds__(GA_THREADS)
detector_affine_brick_kernel(const __grid_constant__ AffLaunch L, const __grid_constant__ CUtensorMap map, uint32_t off)
{
    __shared__ __align__(16) AffSmem s_rec[GA_CHUNK];
    __shared__ __align__(16) AffBrick s_brick[GA_CHUNK];
    __shared__ __align__(128) float s_ring[GA_RING][GA_BRICK];
    __shared__ __align__(8) uint64_t s_full[GA_RING], s_empty[GA_RING];
    const int tiles_x = (L.cols + GA_TW - 1) / GA_TW;
    const int tile = blockIdx.x;
    const int r0 = (tile / tiles_x) * GA_TH, c0 = (tile % tiles_x) * GA_TW;
    const int o_begin = blockIdx.y * L.per_split;
    const int o_end = min(L.n_orient, o_begin + L.per_split);
    double rmax2 = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double fr = (double)(r0 + ((k & 2) ? GA_TH - 1 : 0)), fc = (double)(c0 + ((k & 1) ? GA_TW - 1 : 0));
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double p = L.b_o[a] + fc * L.b_u[a] + fr * L.b_v[a];
            s += p * p;
        }
        rmax2 = fmax(rmax2, s);
    }
    if (rmax2 > L.rin2) {                       // tile can leave the voxel box: clamped gather from the grid
        aff_tile<true, false>(L, s_rec, r0, c0, o_begin, o_end);
        return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < GA_RING; ++i) { mbar_init(s_full + i, 1); mbar_init(s_empty + i, GA_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int cx = 2 * ((warp & 1) * 8 + (lane & 7));
    const int ry = 2 * ((warp >> 1) * 4 + (lane >> 3));
    const int r = r0 + ry, c = c0 + cx;
    const bool live0 = r < L.rows && c < L.cols, live1 = r < L.rows && c + 1 < L.cols;
    const bool live2 = r + 1 < L.rows && c < L.cols, live3 = r + 1 < L.rows && c + 1 < L.cols;
    const int64_t i0 = (int64_t)r * L.cols + c;
    const int F = L.F;
    const uint32_t HM = L.HM;
    const float *iqs = L.iq_shifted;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
    unsigned g = 0;                              // running orientation number of this CTA: ring slot g % GA_RING

    for (int o0 = o_begin; o0 < o_end; o0 += GA_CHUNK) {
        const int nc = min(GA_CHUNK, o_end - o0);
        __syncthreads();
        if (tid < nc) {
            const AffRecord *q = L.rec + (o0 + tid);
            AffSmem e;
            const double fr = (double)r0, fc = (double)c0;
            const double t0 = __fma_rn(fr, q->v[0], __fma_rn(fc, q->u[0], q->o[0]));
            const double t1 = __fma_rn(fr, q->v[1], __fma_rn(fc, q->u[1], q->o[1]));
            const double t2 = __fma_rn(fr, q->v[2], __fma_rn(fc, q->u[2], q->o[2]));
            e.tx = (uint32_t)__double2ll_rn(t0 * L.scale) + L.half;
            e.ty = (uint32_t)__double2ll_rn(t1 * L.scale) + L.half;
            e.tz = (uint32_t)__double2ll_rn(t2 * L.scale) + L.half;
            e.w = q->w;
            e.ux = q->U[0]; e.uy = q->U[1]; e.uz = q->U[2]; e.pad0 = 0;
            e.vx = q->V[0]; e.vy = q->V[1]; e.vz = q->V[2]; e.pad1 = 0;
            s_rec[tid] = e;
            // lowest voxel index the tile's pixels can take along each axis: the coordinate is affine in
            // (row, col), so its minimum sits at a corner of the tile
            const uint32_t ex = (uint32_t)(GA_TW - 1) * (uint32_t)e.ux, fx = (uint32_t)(GA_TH - 1) * (uint32_t)e.vx;
            const uint32_t ey = (uint32_t)(GA_TW - 1) * (uint32_t)e.uy, fy = (uint32_t)(GA_TH - 1) * (uint32_t)e.vy;
            const uint32_t ez = (uint32_t)(GA_TW - 1) * (uint32_t)e.uz, fz = (uint32_t)(GA_TH - 1) * (uint32_t)e.vz;
            const uint32_t mx = min(min(e.tx, e.tx + ex), min(e.tx + fx, e.tx + ex + fx)) >> F;
            const uint32_t my = min(min(e.ty, e.ty + ey), min(e.ty + fy, e.ty + ey + fy)) >> F;
            const uint32_t mz = min(min(e.tz, e.tz + ez), min(e.tz + fz, e.tz + ez + fz)) >> F;
            AffBrick b;
            const uint32_t mz4 = mz - ((mz - off) & 3u);       // z origin rounded down to 4 voxels past the grid start
            b.bx = (int32_t)(mx - off); b.by = (int32_t)(my - off); b.bz = (int32_t)(mz4 - off);
            b.oc = (my * GA_B + mx) * GA_BZ + mz4;
            s_brick[tid] = b;
        }
        __syncthreads();
        if (tid == 0) {
            // every slot is free here (all warps finished the previous chunk): start the ring
            for (int o = 0; o < min(GA_RING - 1, nc); ++o) {
                const unsigned gi = g + o, slot = gi % GA_RING;
                mbar_wait(s_empty + slot, ((gi / GA_RING) & 1) ^ 1);
                mbar_expect_tx(s_full + slot, GA_BRICK * 4);
                tma_load_3d(s_ring[slot], &map, s_brick[o].bz, s_brick[o].bx, s_brick[o].by, s_full + slot);
            }
        }
#pragma unroll 1
        for (int o = 0; o < nc; ++o, ++g) {
            if (tid == 0 && o + GA_RING - 1 < nc) {
                const unsigned gi = g + GA_RING - 1, slot = gi % GA_RING;
                mbar_wait(s_empty + slot, ((gi / GA_RING) & 1) ^ 1);
                mbar_expect_tx(s_full + slot, GA_BRICK * 4);
                const AffBrick b = s_brick[o + GA_RING - 1];
                tma_load_3d(s_ring[slot], &map, b.bz, b.bx, b.by, s_full + slot);
            }
            const uint4 A = *reinterpret_cast<const uint4 *>(&s_rec[o].tx);
            const uint4 B = *reinterpret_cast<const uint4 *>(&s_rec[o].ux);
            const uint4 C = *reinterpret_cast<const uint4 *>(&s_rec[o].vx);
            const uint32_t oc = s_brick[o].oc;
            const float w = __uint_as_float(A.w);
            const uint32_t x0 = A.x + (uint32_t)cx * B.x + (uint32_t)ry * C.x;
            const uint32_t y0 = A.y + (uint32_t)cx * B.y + (uint32_t)ry * C.y;
            const uint32_t z0 = A.z + (uint32_t)cx * B.z + (uint32_t)ry * C.z;
            const uint32_t x1 = x0 + B.x, y1 = y0 + B.y, z1 = z0 + B.z;
            const uint32_t x2 = x0 + C.x, y2 = y0 + C.y, z2 = z0 + C.z;
            const uint32_t x3 = x2 + B.x, y3 = y2 + B.y, z3 = z2 + B.z;
#define GA_LOCAL(X, Y, Z) ((((Y) >> F) * GA_B + ((X) >> F)) * GA_BZ + ((Z) >> F) - oc)
            const uint32_t l0 = GA_LOCAL(x0, y0, z0), l1 = GA_LOCAL(x1, y1, z1);
            const uint32_t l2 = GA_LOCAL(x2, y2, z2), l3 = GA_LOCAL(x3, y3, z3);
#undef GA_LOCAL
            const uint32_t m0 = umin3(x0 & HM, y0 & HM, z0 & HM);
            const uint32_t m1 = umin3(x1 & HM, y1 & HM, z1 & HM);
            const uint32_t m2 = umin3(x2 & HM, y2 & HM, z2 & HM);
            const uint32_t m3 = umin3(x3 & HM, y3 & HM, z3 & HM);
            const unsigned slot = g % GA_RING;
            mbar_wait(s_full + slot, (g / GA_RING) & 1);
            const float *brick = s_ring[slot];
            // (a pixel of a partial tile lies outside the image: its brick index is clamped, its value unused)
            float v0 = brick[min(l0, (uint32_t)GA_BRICK - 1)], v1 = brick[min(l1, (uint32_t)GA_BRICK - 1)];
            float v2 = brick[min(l2, (uint32_t)GA_BRICK - 1)], v3 = brick[min(l3, (uint32_t)GA_BRICK - 1)];
            if (min(min(m0, m1), min(m2, m3)) == 0u) {
                // within the error bound of a voxel edge: the reference's own arithmetic decides, from the grid
                if (m0 == 0u && live0) v0 = __ldg(iqs + aff_exact_voxel(L, i0, o0 + o));
                if (m1 == 0u && live1) v1 = __ldg(iqs + aff_exact_voxel(L, i0 + 1, o0 + o));
                if (m2 == 0u && live2) v2 = __ldg(iqs + aff_exact_voxel(L, i0 + L.cols, o0 + o));
                if (m3 == 0u && live3) v3 = __ldg(iqs + aff_exact_voxel(L, i0 + L.cols + 1, o0 + o));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty + slot);
            a0 = fmaf(w, v0, a0);
            a1 = fmaf(w, v1, a1);
            a2 = fmaf(w, v2, a2);
            a3 = fmaf(w, v3, a3);
        }
        d0 += (double)a0; d1 += (double)a1; d2 += (double)a2; d3 += (double)a3;
        a0 = a1 = a2 = a3 = 0.f;
    }
    if (L.n_split == 1) {
        if (live0) L.image[i0] += d0;
        if (live1) L.image[i0 + 1] += d1;
        if (live2) L.image[i0 + L.cols] += d2;
        if (live3) L.image[i0 + L.cols + 1] += d3;
    } else {
        if (live0) atomicAdd(L.image + i0, d0);
        if (live1) atomicAdd(L.image + i0 + 1, d1);
        if (live2) atomicAdd(L.image + i0 + L.cols, d2);
        if (live3) atomicAdd(L.image + i0 + L.cols + 1, d3);
    }
}

// --------------------------------------------------------- affine fit ----
// max |p(r,c) - L(r,c)| per component, L = interpolation of the corners
// p[0,0], p[0,-1], p[-1,0] evaluated in fp64.
__global__ void __launch_bounds__(256)
affine_deviation_kernel(const double *__restrict__ px, const double *__restrict__ py, const double *__restrict__ pz,
                        int rows, int cols, unsigned long long *out3)
{
    const double *p[3] = {px, py, pz};
    double o[3], u[3], v[3], dev[3] = {0.0, 0.0, 0.0};
    const int C = max(cols - 1, 1), R = max(rows - 1, 1);
    for (int a = 0; a < 3; ++a) {
        o[a] = p[a][0];
        u[a] = (p[a][cols - 1] - o[a]) / (double)C;
        v[a] = (p[a][(int64_t)(rows - 1) * cols] - o[a]) / (double)R;
    }
    const int64_t n = (int64_t)rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols), c = (int)(i % cols);
        for (int a = 0; a < 3; ++a) {
            const double m = o[a] + (double)c * u[a] + (double)r * v[a];
            dev[a] = fmax(dev[a], fabs(p[a][i] - m));
        }
    }
    for (int a = 0; a < 3; ++a) {
        double d = dev[a];
        for (int s = 16; s > 0; s >>= 1) d = fmax(d, __shfl_xor_sync(0xffffffffu, d, s));
        if ((threadIdx.x & 31) == 0 && d > 0.0) atomicMax(out3 + a, (unsigned long long)__double_as_longlong(d));
    }
}

extern "C" int gx_grid_affine_fit(const double *d_px, const double *d_py, const double *d_pz, int rows, int cols,
                                  double *d_scratch3, double *h_corners9, double *h_dev3, void *stream)
{
    GX_REQUIRE(d_px && d_py && d_pz && d_scratch3 && h_corners9 && h_dev3, "NULL pointer");
    GX_REQUIRE(rows > 1 && cols > 1, "grid must be at least 2 x 2");
    cudaStream_t st = gx_stream(stream);
    GX_CUDA(cudaMemsetAsync(d_scratch3, 0, 3 * sizeof(double), st));
    int64_t blocks = ((int64_t)rows * cols + 255) / 256;
    if (blocks > GX_SM_COUNT * 8) blocks = GX_SM_COUNT * 8;
    affine_deviation_kernel<<<(int)blocks, 256, 0, st>>>(d_px, d_py, d_pz, rows, cols,
                                                          reinterpret_cast<unsigned long long *>(d_scratch3));
    int rc = gx_check_launch("gx_grid_affine_fit");
    if (rc != GX_OK) return rc;
    const double *p[3] = {d_px, d_py, d_pz};
    const int64_t at[3] = {0, cols - 1, (int64_t)(rows - 1) * cols};
    for (int j = 0; j < 3; ++j)
        for (int a = 0; a < 3; ++a)
            GX_CUDA(cudaMemcpyAsync(h_corners9 + 3 * j + a, p[a] + at[j], sizeof(double), cudaMemcpyDeviceToHost, st));
    GX_CUDA(cudaMemcpyAsync(h_dev3, d_scratch3, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    GX_CUDA(cudaStreamSynchronize(st));
    return GX_OK;
}

// ------------------------------------------------- host: model + bounds ----
static void aff_host_matvec(const double *R, const double *p, double *o)
{
    for (int r = 0; r < 3; ++r) {
        double t = R[3 * r] * p[0];
        t = fma(R[3 * r + 1], p[1], t);
        t = fma(R[3 * r + 2], p[2], t);
        o[r] = t;
    }
}

static double aff_host_floordiv(double a, double b)
{
    const double inv_b = 1.0 / b;
    double q = floor(a * inv_b);
    const double r = fma(-q, b, a);
    if (r < 0.0) q -= 1.0;
    else if (r >= b) q += 1.0;
    return q;
}

// clamped voxel index of a_ = p - qmin, as the reference computes it
static double aff_host_index(double a, double dq, int n)
{
    const double q = aff_host_floordiv(a, dq);
    return !(q > 0.0) ? 0.0 : (q >= (double)n ? (double)(n - 1) : q);
}

// order-preserving map double <-> int64 (bisection over neighbouring doubles)
static int64_t aff_ord(double v)
{
    int64_t b;
    memcpy(&b, &v, 8);
    return b < 0 ? (int64_t)0x8000000000000000ull - b : b;
}
static double aff_unord(int64_t o)
{
    const int64_t b = o < 0 ? (int64_t)0x8000000000000000ull - o : o;
    double v;
    memcpy(&v, &b, 8);
    return v;
}

extern "C" int gx_affine_record_bytes(void) { return (int)sizeof(AffRecord); }
extern "C" int gx_affine_plan_doubles(void) { return 8; }

// h_corners9 = p[0,0], p[0,-1], p[-1,0] of the base grid (rows), h_dev3 = measured deviation
// from their interpolation (gx_grid_affine_fit); h_R [n][3][9], h_w [n].
// Fills n records and h_plan = {F, half, off, rin2, E_max in voxels, n_const, n_edge_locked, n_step}.
// Returns GX_ERR_UNSUPPORTED when the grid is too far from affine or the coordinates do not fit
// the fixed-point format; the caller then uses gx_detector_accumulate.
extern "C" int gx_host_affine_orientations(const double *h_corners9, const double *h_dev3, int rows, int cols,
                                           const double *h_R, const double *h_w, int n, double qx_min,
                                           double qy_min, double qz_min, double dq, int Vy, int Vx, int Vz,
                                           void *h_records, double *h_plan)
{
    GX_REQUIRE(h_corners9 && h_dev3 && h_R && h_w && h_records && h_plan, "NULL pointer");
    GX_REQUIRE(rows > 1 && cols > 1 && n > 0 && dq > 0.0 && Vx > 0 && Vy > 0 && Vz > 0, "bad arguments");
    AffRecord *rec = reinterpret_cast<AffRecord *>(h_records);
    const double qmin[3] = {qx_min, qy_min, qz_min};
    const int V[3] = {Vx, Vy, Vz};
    const double ulp = ldexp(1.0, -53);
    const double gamma = 4.0 * ulp;            // 3 roundings of an fma chain, rounded up
    const double Cn = (double)(cols - 1), Rn = (double)(rows - 1);
    // pixels of a partial tile extrapolate beyond the image but, in an unclamped tile, stay in the box
    double tmin = 0.0, tmax = (double)(Vx > Vy ? (Vx > Vz ? Vx : Vz) : (Vy > Vz ? Vy : Vz)), e_max = 0.0;
    int n_const = 0, n_locked = 0, n_step = 0;
    // in-box radius: |p| <= rin keeps every component inside [qmin, qmin + V dq) under any rotation
    double rin = 1e300;
    for (int a = 0; a < 3; ++a) {
        rin = fmin(rin, -qmin[a]);
        rin = fmin(rin, qmin[a] + (double)V[a] * dq);
    }
    for (int o = 0; o < n; ++o) {
        double c[4][3], D[3];
        for (int j = 0; j < 3; ++j)
            for (int a = 0; a < 3; ++a) c[j][a] = h_corners9[3 * j + a];
        for (int a = 0; a < 3; ++a) {
            c[3][a] = c[1][a] + c[2][a] - c[0][a];
            // measured deviation + rounding of the device's own model evaluation
            D[a] = h_dev3[a] + 16.0 * ulp * (fabs(c[0][a]) + fabs(c[1][a]) + fabs(c[2][a]));
        }
        for (int s = 0; s < 3; ++s) {
            const double *R = h_R + ((size_t)o * 3 + s) * 9;
            double pmax[3], nd[3];
            for (int a = 0; a < 3; ++a) {
                pmax[a] = 0.0;
                for (int j = 0; j < 4; ++j) pmax[a] = fmax(pmax[a], fabs(c[j][a]));
                pmax[a] = (pmax[a] + D[a]) * (1.0 + 8.0 * ulp);
            }
            for (int i = 0; i < 3; ++i) {
                double lin = 0.0, rho = 0.0;
                for (int k = 0; k < 3; ++k) {
                    lin += fabs(R[3 * i + k]) * D[k];
                    rho += fabs(R[3 * i + k]) * pmax[k];
                }
                nd[i] = (lin + 4.0 * gamma * rho) * (1.0 + 16.0 * ulp);
            }
            for (int j = 0; j < 3; ++j) {
                double q[3];
                aff_host_matvec(R, c[j], q);
                c[j][0] = q[0]; c[j][1] = q[1]; c[j][2] = q[2];
            }
            for (int a = 0; a < 3; ++a) {
                c[3][a] = c[1][a] + c[2][a] - c[0][a];
                D[a] = nd[a] + 8.0 * ulp * (fabs(c[0][a]) + fabs(c[1][a]) + fabs(c[2][a]));
            }
        }
        AffRecord &r = rec[o];
        memset(&r, 0, sizeof(r));
        r.w = (float)h_w[o];
        bool locked = false;
        for (int a = 0; a < 3; ++a) {
            double lo = c[0][a], hi = c[0][a], pm = 0.0;
            for (int j = 0; j < 4; ++j) {
                lo = fmin(lo, c[j][a]); hi = fmax(hi, c[j][a]);
                pm = fmax(pm, fabs(c[j][a]));
            }
            lo = nextafter(lo - D[a], -INFINITY);
            hi = nextafter(hi + D[a], INFINITY);
            const double alo = lo - qmin[a], ahi = hi - qmin[a];     // rounding is monotone
            const double ilo = aff_host_index(alo, dq, V[a]), ihi = aff_host_index(ahi, dq, V[a]);
            if (ilo == ihi) {
                // every pixel lands in the same voxel along this axis (p - qmin itself may still vary)
                r.o[a] = ilo + 0.5; r.u[a] = 0.0; r.v[a] = 0.0;
                ++r.n_const;
                ++n_const;
                continue;
            }
            const double spread = (hi - lo) / dq;
            if (spread < 1e-2 && ihi == ilo + 1.0) {
                // The coordinate barely moves over the detector but straddles a voxel edge (e.g. a plane
                // lying in q_z = 0 up to 1e-16 rounding noise, with q_z = 0 on an edge): which side a pixel
                // falls on is decided by how p - qmin ROUNDS.  Find the two neighbouring doubles A1 < A2 of
                // p - qmin where the index steps; the step is at p = qmin + (A1 + A2)/2.  Model
                // t = ilo + 1 + (p - p_step) / big, so floor(t) is the index and the kernel's edge band
                // guards the rounding boundary.
                int64_t o1 = aff_ord(alo), o2 = aff_ord(ahi);
                while (o2 - o1 > 1) {
                    const int64_t mid = o1 + (o2 - o1) / 2;
                    if (aff_host_index(aff_unord(mid), dq, V[a]) == ilo) o1 = mid; else o2 = mid;
                }
                const double A1 = aff_unord(o1), A2 = aff_unord(o2);
                // s + err = A1 + qmin exactly (TwoSum)
                const double sum = A1 + qmin[a], bb = sum - A1;
                const double err = (A1 - (sum - bb)) + (qmin[a] - bb);
                const double gap = (A2 - A1) * 0.5;
                const double ext = 2.0 + 128.0 / fmin(Cn, Rn);         // partial tiles extrapolate
                const double big = (hi - lo) * ext * 1.01;
                const double c0s = ((c[0][a] - sum) - err) - gap;       // p[0,0] - p_step
                const double d_step = D[a] + 8.0 * ulp * (fabs(c[0][a] - sum) + fabs(err) + gap);
                if (d_step / big < 1e-7) {
                    r.o[a] = (ilo + 1.0) + c0s / big;
                    r.u[a] = (c[1][a] - c[0][a]) / (Cn * big);
                    r.v[a] = (c[2][a] - c[0][a]) / (Rn * big);
                    const double mag = fabs(r.o[a]) + Cn * fabs(r.u[a]) + Rn * fabs(r.v[a]);
                    e_max = fmax(e_max, d_step / big + 64.0 * ulp * (mag + 1024.0));
                    ++n_step;
                    continue;
                }
            }
            r.o[a] = (c[0][a] - qmin[a]) / dq;
            r.u[a] = (c[1][a] - c[0][a]) / (Cn * dq);
            r.v[a] = (c[2][a] - c[0][a]) / (Rn * dq);
            const double e_ref = (D[a] + 2.0 * ulp * (pm + D[a] + fabs(qmin[a]))) / dq;
            const double mag = fabs(r.o[a]) + Cn * fabs(r.u[a]) + Rn * fabs(r.v[a]);
            const double e = e_ref + 64.0 * ulp * (mag + 1024.0);
            e_max = fmax(e_max, e);
            tmin = fmin(tmin, (lo - qmin[a]) / dq);
            tmax = fmax(tmax, (hi - qmin[a]) / dq);
            // nearly constant and on a voxel edge: every pixel will take the exact path
            if (spread < 1e-3) locked = true;
        }
        if (locked) ++n_locked;
    }
    // partial tiles extrapolate up to a tile beyond the image; offsets keep coordinates positive
    const double off = ceil(-tmin) + 4.0;
    const double top = tmax + off + 4.0;
    GX_REQUIRE(top < 1048576.0 && off < 1048576.0, "voxel coordinates out of range");
    int F = 31 - ilogb(top);
    if (F > 24) F = 24;
    const double scale = ldexp(1.0, F);
    // device: fp64 tile base (rounded to nearest) + c'' U + r'' V with U, V rounded to nearest
    const double e_quant = (0.5 * (1.0 + (GA_TW - 1) + (GA_TH - 1)) + 1.5) / scale;
    const double e_tot = e_max + e_quant;
    const double e_fix = ceil(e_tot * scale) + 1.0;
    int kb = 0;
    while (ldexp(1.0, kb) < e_fix) ++kb;             // half = 2^kb >= e_fix
    if (kb + 1 > F - 2) {
        gx_set_error("gx_host_affine_orientations: error bound %.3g voxels too large for the fixed-point filter", e_tot);
        return GX_ERR_UNSUPPORTED;
    }
    for (int o = 0; o < n; ++o) {
        AffRecord &r = rec[o];
        for (int a = 0; a < 3; ++a) {
            r.o[a] += off;
            r.U[a] = (int32_t)llrint(r.u[a] * scale);
            r.V[a] = (int32_t)llrint(r.v[a] * scale);
        }
    }
    h_plan[0] = (double)F;
    h_plan[1] = ldexp(1.0, kb);
    h_plan[2] = off;
    h_plan[3] = rin > 0.0 ? rin * rin * (1.0 - 1e-9) : -1.0;
    h_plan[4] = e_tot;
    h_plan[5] = (double)n_const;
    h_plan[6] = (double)n_locked;
    h_plan[7] = (double)n_step;
    return GX_OK;
}

static int affine_launch(const float *d_iq, int Vy, int Vx, int Vz, double qx_min,
                         double qy_min, double qz_min, double dq, const double *d_px,
                         const double *d_py, const double *d_pz, int rows, int cols,
                         const double *h_corners9, const void *d_records, const double *d_R,
                         int n_orient, const double *h_plan, double *d_image, int probe,
                         int64_t *d_index_out, unsigned long long *d_slow_count, void *stream,
                         const float *d_iq_padded, int Vz_padded)
{
    GX_REQUIRE(d_iq && d_px && d_py && d_pz && h_corners9 && d_records && d_R && h_plan && d_image, "NULL pointer");
    GX_REQUIRE(Vy > 0 && Vx > 0 && Vz > 0 && dq > 0.0, "bad voxel grid");
    GX_REQUIRE(rows > 1 && cols > 1 && n_orient > 0, "empty input");
    AffLaunch L;
    memset(&L, 0, sizeof(L));
    const uint32_t off = (uint32_t)h_plan[2];
    GX_REQUIRE(((double)off + Vy) * ((double)off + Vx) * ((double)off + Vz) < 4294967296.0,
               "voxel grid too large for 32-bit offsets");
    L.shift = off * ((uint32_t)Vx * (uint32_t)Vz + (uint32_t)Vz + 1u);
    L.iq = d_iq;
    L.iq_shifted = d_iq - (ptrdiff_t)L.shift;
    L.Vx = (uint32_t)Vx; L.Vy = (uint32_t)Vy; L.Vz = (uint32_t)Vz;
    L.lo = off; L.hix = off + Vx - 1; L.hiy = off + Vy - 1; L.hiz = off + Vz - 1;
    L.F = (int)h_plan[0];
    L.half = (uint32_t)h_plan[1];
    L.HM = ((1u << L.F) - 1u) & ~(2u * L.half - 1u);
    L.scale = ldexp(1.0, L.F);
    for (int a = 0; a < 3; ++a) {
        L.b_o[a] = h_corners9[a];
        L.b_u[a] = (h_corners9[3 + a] - h_corners9[a]) / (double)(cols - 1);
        L.b_v[a] = (h_corners9[6 + a] - h_corners9[a]) / (double)(rows - 1);
    }
    L.rin2 = h_plan[3];
    L.qmin[0] = qx_min; L.qmin[1] = qy_min; L.qmin[2] = qz_min;
    L.dq = dq; L.inv_dq = 1.0 / dq;
    L.px = d_px; L.py = d_py; L.pz = d_pz;
    L.R27 = d_R;
    L.rec = reinterpret_cast<const AffRecord *>(d_records);
    L.rows = rows; L.cols = cols; L.n_orient = n_orient;
    L.image = d_image;
    L.probe = probe;
    L.index_out = d_index_out;
    L.slow_count = d_slow_count;
    const int tiles = ((cols + GA_TW - 1) / GA_TW) * ((rows + GA_TH - 1) / GA_TH);
    // enough CTAs for ~8 per SM: split the orientation range when the image is small
    int n_split = 1;
    const int chunks = (n_orient + GA_CHUNK - 1) / GA_CHUNK;
    while (tiles * n_split < GX_SM_COUNT * 8 && n_split < chunks) ++n_split;
    const int per = ((chunks + n_split - 1) / n_split) * GA_CHUNK;
    n_split = (n_orient + per - 1) / per;
    L.per_split = per;
    L.n_split = n_split;
    dim3 grid((unsigned)tiles, (unsigned)n_split);
    if (d_iq_padded) {
        // TMA-brick variant: 3-D tensor map of the grid padded to a 16-byte row pitch, box 8 x 8 x 8
        GX_REQUIRE(Vz_padded >= Vz && Vz_padded % 4 == 0 && !(d_index_out && probe >= 0) && !d_slow_count,
                   "brick variant: padded pitch must be a multiple of 4 floats; no probe / slow count");
        gx_encode_tiled_fn encode = gx_tensor_map_encoder();
        if (!encode) return GX_ERR_UNSUPPORTED;
        CUtensorMap map;
        const cuuint64_t dims[3] = {(cuuint64_t)Vz, (cuuint64_t)Vx, (cuuint64_t)Vy};
        const cuuint64_t strides[2] = {(cuuint64_t)Vz_padded * 4, (cuuint64_t)Vz_padded * 4 * (cuuint64_t)Vx};
        const cuuint32_t box[3] = {GA_BZ, GA_B, GA_B};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(d_iq_padded), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            gx_set_error("gx_detector_accumulate_affine_brick: cuTensorMapEncodeTiled failed (%d)", (int)r);
            return GX_ERR_CUDA;
        }
        detector_affine_brick_kernel<<<grid, GA_THREADS, 0, gx_stream(stream)>>>(L, map, off);
        return gx_check_launch("gx_detector_accumulate_affine_brick");
    }
    if (d_index_out && probe >= 0)
        detector_affine_kernel<true><<<grid, GA_THREADS, 0, gx_stream(stream)>>>(L);
    else
        detector_affine_kernel<false><<<grid, GA_THREADS, 0, gx_stream(stream)>>>(L);
    return gx_check_launch("gx_detector_accumulate_affine");
}

extern "C" int gx_detector_accumulate_affine(const float *d_iq, int Vy, int Vx, int Vz, double qx_min,
                                             double qy_min, double qz_min, double dq, const double *d_px,
                                             const double *d_py, const double *d_pz, int rows, int cols,
                                             const double *h_corners9, const void *d_records, const double *d_R,
                                             int n_orient, const double *h_plan, double *d_image, int probe,
                                             int64_t *d_index_out, unsigned long long *d_slow_count, void *stream)
{
    return affine_launch(d_iq, Vy, Vx, Vz, qx_min, qy_min, qz_min, dq, d_px, d_py, d_pz, rows, cols, h_corners9,
                         d_records, d_R, n_orient, h_plan, d_image, probe, d_index_out, d_slow_count, stream, NULL, 0);
}

// Same result through the TMA-brick variant of the gather (see detector_affine_brick_kernel): d_iq_padded is a
// copy of the grid with rows of Vz_padded floats (a multiple of 4: tensor-map strides are multiples of 16
// bytes).  The caller must have checked that every tile's voxel span fits the 8 x 8 x 8 brick:
// (31 |U_a| + 15 |V_a|) / 2^F < 7 for every orientation record and axis.
extern "C" int gx_detector_accumulate_affine_brick(const float *d_iq, const float *d_iq_padded, int Vz_padded, int Vy,
                                                   int Vx, int Vz, double qx_min, double qy_min, double qz_min, double dq,
                                                   const double *d_px, const double *d_py, const double *d_pz, int rows,
                                                   int cols, const double *h_corners9, const void *d_records,
                                                   const double *d_R, int n_orient, const double *h_plan, double *d_image,
                                                   void *stream)
{
    GX_REQUIRE(d_iq_padded != NULL, "padded grid missing");
    return affine_launch(d_iq, Vy, Vx, Vz, qx_min, qy_min, qz_min, dq, d_px, d_py, d_pz, rows, cols, h_corners9,
                         d_records, d_R, n_orient, h_plan, d_image, -1, NULL, NULL, stream, d_iq_padded, Vz_padded);
}
