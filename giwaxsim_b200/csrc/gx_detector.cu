// K4 / K4b: rotated detector planes intersected with the voxel grid, summed
// over orientations, then mirrored (detector.py:33-300, comparison.py:790-870).
//
// Exactness contract: the reference rotates the full P x P coordinate grids
// three times per orientation with `R @ X` (OpenBLAS dgemm = fma chain
// k=0,1,2) and floor-divides the result.  The kernel applies the same three
// chains per pixel in fp64 and the exact floor-divide, so the voxel index of
// every pixel of every orientation is bit-identical.  The three matrices per
// orientation come from the host (gx_host_orientation_matrices), which tracks
// the three corner pixels through the same chains.
#include <math.h>
#include <string.h>
#include "gx_common.cuh"

__device__ __forceinline__ void matvec_chain(const double *R, double x, double y, double z,
                                             double &ox, double &oy, double &oz)
{
    ox = __fma_rn(R[2], z, __fma_rn(R[1], y, __dmul_rn(R[0], x)));
    oy = __fma_rn(R[5], z, __fma_rn(R[4], y, __dmul_rn(R[3], x)));
    oz = __fma_rn(R[8], z, __fma_rn(R[7], y, __dmul_rn(R[6], x)));
}

struct Mat3 { double m[9]; };

__global__ void __launch_bounds__(256)
rotate_points_kernel(Mat3 R, const double *__restrict__ x, const double *__restrict__ y,
                     const double *__restrict__ z, int64_t n, double *ox, double *oy, double *oz)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double a, b, c;
        matvec_chain(R.m, x[i], y[i], z[i], a, b, c);
        ox[i] = a; oy[i] = b; oz[i] = c;
    }
}

extern "C" int gx_rotate_points(const double *h_R9, const double *d_x, const double *d_y, const double *d_z,
                                int64_t n, double *d_ox, double *d_oy, double *d_oz, void *stream)
{
    GX_REQUIRE(h_R9 && d_x && d_y && d_z && d_ox && d_oy && d_oz, "NULL pointer");
    GX_REQUIRE(n > 0, "no points");
    Mat3 R;
    for (int i = 0; i < 9; ++i) R.m[i] = h_R9[i];
    int64_t blocks = (n + 255) / 256;
    if (blocks > GX_SM_COUNT * 16) blocks = GX_SM_COUNT * 16;
    rotate_points_kernel<<<(int)blocks, 256, 0, gx_stream(stream)>>>(R, d_x, d_y, d_z, n, d_ox, d_oy, d_oz);
    return gx_check_launch("gx_rotate_points");
}

// ------------------------------------------------------------ accumulate ----
// One thread per pixel, orientations looped inside so the image is touched
// once per launch; the 27 matrix entries + weight of the current orientation
// chunk are staged in shared memory and read as warp-wide broadcasts.
#define DET_THREADS 128
#define DET_CHUNK 64          // orientations staged per shared-memory refill

struct DetGrid {
    const float *iq;
    int Vy, Vx, Vz;
    double qx_min, qy_min, qz_min, dq, inv_dq;
};

__device__ __forceinline__ int clamp_index(double p, double qmin, double dq, double inv_dq, int n)
{
    double q = gx_floordiv(__dsub_rn(p, qmin), dq, inv_dq);
    // np.clip(idx, 0, n-1) after astype(int)
    if (!(q > 0.0)) return 0;
    if (q >= (double)n) return n - 1;
    return (int)q;
}

__global__ void __launch_bounds__(DET_THREADS)
detector_accumulate_kernel(DetGrid g, const double *__restrict__ px, const double *__restrict__ py,
                           const double *__restrict__ pz, int64_t n_pix,
                           const double *__restrict__ R, const double *__restrict__ w, int n_orient,
                           double *image, int probe, int64_t *index_out)
{
    __shared__ double s_R[DET_CHUNK][28];
    __shared__ int s_skip[DET_CHUNK];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n_pix;
    const double x0 = live ? px[i] : 0.0, y0 = live ? py[i] : 0.0, z0 = live ? pz[i] : 0.0;
    double acc = 0.0;
    for (int o0 = 0; o0 < n_orient; o0 += DET_CHUNK) {
        const int nc = min(DET_CHUNK, n_orient - o0);
        __syncthreads();
        for (int t = threadIdx.x; t < nc * 28; t += blockDim.x) {
            const int o = t / 28, k = t - o * 28;
            s_R[o][k] = (k < 27) ? R[(size_t)(o0 + o) * 27 + k] : w[o0 + o];
        }
        __syncthreads();
        // a step that is exactly the identity (rotation by 0 degrees) returns its input
        // unchanged under the fma chain, so it can be skipped without changing a bit
        for (int o = threadIdx.x; o < nc; o += blockDim.x) {
            int skip = 0;
            for (int k = 0; k < 3; ++k) {
                bool ident = true;
                for (int e = 0; e < 9; ++e) ident = ident && (s_R[o][9 * k + e] == ((e % 4 == 0) ? 1.0 : 0.0));
                if (ident) skip |= 1 << k;
            }
            s_skip[o] = skip;
        }
        __syncthreads();
        if (!live) continue;
        for (int o = 0; o < nc; ++o) {
            const double *m = s_R[o];
            const int skip = s_skip[o];
            double x3 = x0, y3 = y0, z3 = z0, a, b, c;
            if (!(skip & 1)) { matvec_chain(m, x3, y3, z3, a, b, c); x3 = a; y3 = b; z3 = c; }
            if (!(skip & 2)) { matvec_chain(m + 9, x3, y3, z3, a, b, c); x3 = a; y3 = b; z3 = c; }
            if (!(skip & 4)) { matvec_chain(m + 18, x3, y3, z3, a, b, c); x3 = a; y3 = b; z3 = c; }
            const int ix = clamp_index(x3, g.qx_min, g.dq, g.inv_dq, g.Vx);
            const int iy = clamp_index(y3, g.qy_min, g.dq, g.inv_dq, g.Vy);
            const int iz = clamp_index(z3, g.qz_min, g.dq, g.inv_dq, g.Vz);
            const size_t v = ((size_t)iy * g.Vx + ix) * g.Vz + iz;
            acc += (double)__ldg(&g.iq[v]) * m[27];
            if (index_out && o0 + o == probe) index_out[i] = (int64_t)v;
        }
    }
    if (live) image[i] += acc;
}

extern "C" int gx_detector_accumulate(const float *d_iq, int Vy, int Vx, int Vz,
                                      double qx_min, double qy_min, double qz_min, double dq,
                                      const double *d_px, const double *d_py, const double *d_pz, int64_t n_pix,
                                      const double *d_R, const double *d_w, int n_orient,
                                      double *d_image, int probe, int64_t *d_index_out, void *stream)
{
    GX_REQUIRE(d_iq && d_px && d_py && d_pz && d_R && d_w && d_image, "NULL pointer");
    GX_REQUIRE(Vy > 0 && Vx > 0 && Vz > 0 && dq > 0.0, "bad voxel grid");
    GX_REQUIRE(n_pix > 0 && n_orient > 0, "empty input");
    DetGrid g;
    g.iq = d_iq; g.Vy = Vy; g.Vx = Vx; g.Vz = Vz;
    g.qx_min = qx_min; g.qy_min = qy_min; g.qz_min = qz_min; g.dq = dq; g.inv_dq = 1.0 / dq;
    int64_t blocks = (n_pix + DET_THREADS - 1) / DET_THREADS;
    GX_REQUIRE(blocks < 2147483647LL, "too many pixels");
    detector_accumulate_kernel<<<(int)blocks, DET_THREADS, 0, gx_stream(stream)>>>(
        g, d_px, d_py, d_pz, n_pix, d_R, d_w, n_orient, d_image, probe, d_index_out);
    return gx_check_launch("gx_detector_accumulate");
}

// ---------------------------------------------------- accumulate, filtered ----
// The exact kernel above spends ~60 fp64 operations per pixel and orientation
// and is bound by the fp64 pipe, not by memory.  This variant evaluates the
// voxel coordinate t_c = (p_c - qmin_c)/dq first in fp32 from one collapsed
// 3x3 matrix per orientation (host: M = R3 R2 R1, scaled by 1/dq): 9 FFMA.
// The host also supplies a rigorous bound delta_c on |t_fp32 - t_exact|
// (rounding of the inputs, of M and of the three FMAs).  floor(t_fp32) equals
// the reference's index whenever t_fp32 is farther than delta from an integer;
// only the remaining lanes (a few 1e-4 of the pixels, or whole orientations
// whose pixels sit exactly on voxel edges) redo the exact fp64 chain.  The
// result is therefore bit-identical to the exact kernel's.
struct FastOrient {            // 160 bytes per orientation, staged in shared memory
    float m[9];
    float off[3];
    float slack[3];            // 0.5 - delta_c: |frac - 0.5| must not exceed it
    int32_t skip;              // bit k set: rotation step k is exactly the identity
    double w;
    double R1[9];              // first exact rotation step (the only one when skip == 6)
    double pad2[2];
};
static_assert(sizeof(FastOrient) == 160, "FastOrient layout");

#define DETF_THREADS 256
#define DETF_CHUNK 64
#define DETF_UNROLL 4

// floor(t) for t clamped to [0.5, n - 0.5] (so the result is already the clamped
// voxel index) via the 1.5*2^23 rounding trick; returns |frac - 0.5|.
__device__ __forceinline__ float fast_floor_bits(float t, float hi, int &bits)
{
    t = fminf(fmaxf(t, 0.5f), hi);
    const float MAGIC = 12582912.0f;
    const float s = (t - 0.5f) + MAGIC;          // MAGIC + nearest integer to t - 0.5
    bits = __float_as_int(s);                     // index = bits - 0x4B400000
    return fabsf((t - (s - MAGIC)) - 0.5f);
}

// exact clamped index of coordinate p with a one-entry memo: grid-aligned
// orientations give the same p - qmin for every pixel and orientation, and the
// floor-divide (7 fp64 operations) is then replaced by one comparison.
struct IndexMemo { double a; int idx; };
__device__ __forceinline__ int exact_index_memo(double p, double qmin, double dq, double inv_dq, int n,
                                                IndexMemo &memo)
{
    const double a = __dsub_rn(p, qmin);
    if (a == memo.a) return memo.idx;
    double q = gx_floordiv(a, dq, inv_dq);
    int idx = !(q > 0.0) ? 0 : (q >= (double)n ? n - 1 : (int)q);
    memo.a = a; memo.idx = idx;
    return idx;
}

__device__ __forceinline__ double chain_row(const double *R, int row, double x, double y, double z)
{
    return __fma_rn(R[3 * row + 2], z, __fma_rn(R[3 * row + 1], y, __dmul_rn(R[3 * row], x)));
}

// all three steps (tilted orientations: needed for a few 1e-3 of the pixels; out of line,
// matrices read from global memory)
__device__ __noinline__ void exact_chain_full(const double *R27, int skip, double &x, double &y, double &z)
{
    for (int k = 0; k < 3; ++k) {
        if ((skip >> k) & 1) continue;
        double a, b, c;
        matvec_chain(R27 + 9 * k, x, y, z, a, b, c);
        x = a; y = b; z = c;
    }
}

// exact index of one component with a scalar one-entry memo (kept in registers)
#define GX_EXACT_INDEX(P, QMIN, N, MEMO_A, MEMO_I, OUT)                          \
    do {                                                                          \
        const double a_ = __dsub_rn(P, QMIN);                                     \
        if (a_ != MEMO_A) {                                                       \
            const double q_ = gx_floordiv(a_, g.dq, g.inv_dq);                    \
            MEMO_I = !(q_ > 0.0) ? 0 : (q_ >= (double)(N) ? (N) - 1 : (int)q_);   \
            MEMO_A = a_;                                                          \
        }                                                                         \
        OUT = MEMO_I;                                                             \
    } while (0)

// One thread per pixel, all orientations looped inside (the image is touched once).
// The per-orientation records are read straight from global memory with
// warp-uniform 128-bit loads: the table is small (160 B x orientations), stays in
// L1 and needs no shared-memory staging -- hence no block barrier, which in the
// staged variants cost a third of the run time because warps that take the exact
// path for a few lanes made the whole CTA wait.  Shared memory is left unused on
// purpose: the voxel gather lives on L1 hits and wants all 228 KB as cache.
template <bool PROBE>
__global__ void __launch_bounds__(DETF_THREADS, 4)
detector_accumulate_fast_kernel(DetGrid g, const double *__restrict__ px, const double *__restrict__ py,
                                const double *__restrict__ pz, int64_t n_pix,
                                const FastOrient *__restrict__ fo, const double *__restrict__ R27,
                                int n_orient,
                                double *image, int probe, int64_t *index_out, unsigned long long *slow_count)
{
    const int64_t i = (int64_t)blockIdx.x * DETF_THREADS + threadIdx.x;
    if (i >= n_pix) return;
    const float hx = (float)g.Vx - 0.5f, hy = (float)g.Vy - 0.5f, hz = (float)g.Vz - 0.5f;
    const unsigned MB = 0x4B400000u;
    const unsigned uVx = (unsigned)g.Vx, uVz = (unsigned)g.Vz;
    const float *iq = g.iq;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    unsigned slow = 0;
    const double x0 = px[i], y0 = py[i], z0 = pz[i];
    const float fx = (float)x0, fy = (float)y0, fz = (float)z0;
    double max_ = nan, may_ = nan, maz_ = nan;      // memo keys (p - qmin) per component
    int mix_ = 0, miy_ = 0, miz_ = 0;               // memo values
    double acc = 0.0;
#pragma unroll 4
    for (int o = 0; o < n_orient; ++o) {
        const float4 *rec = reinterpret_cast<const float4 *>(fo + o);
        const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2), q3 = __ldg(rec + 3);
        const double w = __ldg(&fo[o].w);
        const float tx = fmaf(q0.z, fz, fmaf(q0.y, fy, fmaf(q0.x, fx, q2.y)));
        const float ty = fmaf(q1.y, fz, fmaf(q1.x, fy, fmaf(q0.w, fx, q2.z)));
        const float tz = fmaf(q2.x, fz, fmaf(q1.w, fy, fmaf(q1.z, fx, q2.w)));
        int bx, by, bz;
        const bool okx = fast_floor_bits(tx, hx, bx) <= q3.x;
        const bool oky = fast_floor_bits(ty, hy, by) <= q3.y;
        const bool okz = fast_floor_bits(tz, hz, bz) <= q3.z;
        unsigned ix = (unsigned)bx - MB, iy = (unsigned)by - MB, iz = (unsigned)bz - MB;
        if (!(okx && oky && okz)) {
            const int skip = __float_as_int(q3.w);
            double x3 = x0, y3 = y0, z3 = z0;
            if (skip == 6) {
                // only the first step rotates: each failing component needs one row of it
                const double *R1 = fo[o].R1;
                if (!okx) x3 = chain_row(R1, 0, x0, y0, z0);
                if (!oky) y3 = chain_row(R1, 1, x0, y0, z0);
                if (!okz) z3 = chain_row(R1, 2, x0, y0, z0);
            } else {
                exact_chain_full(R27 + (size_t)o * 27, skip, x3, y3, z3);
            }
            int e_;
            if (!okx) { GX_EXACT_INDEX(x3, g.qx_min, g.Vx, max_, mix_, e_); ix = (unsigned)e_; }
            if (!oky) { GX_EXACT_INDEX(y3, g.qy_min, g.Vy, may_, miy_, e_); iy = (unsigned)e_; }
            if (!okz) { GX_EXACT_INDEX(z3, g.qz_min, g.Vz, maz_, miz_, e_); iz = (unsigned)e_; }
            ++slow;
        }
        const unsigned v = (iy * uVx + ix) * uVz + iz;
        acc += (double)__ldg(iq + v) * w;
        if (PROBE && o == probe) index_out[i] = (int64_t)v;
    }
    image[i] += acc;
    if (slow_count && slow) atomicAdd(slow_count, (unsigned long long)slow);
}
#undef GX_EXACT_INDEX

extern "C" int gx_detector_accumulate_fast(const float *d_iq, int Vy, int Vx, int Vz,
                                           double qx_min, double qy_min, double qz_min, double dq,
                                           const double *d_px, const double *d_py, const double *d_pz,
                                           int64_t n_pix, const void *d_fast, const double *d_R, int n_orient,
                                           double *d_image, int probe, int64_t *d_index_out,
                                           unsigned long long *d_slow_count, void *stream)
{
    GX_REQUIRE(d_iq && d_px && d_py && d_pz && d_fast && d_R && d_image, "NULL pointer");
    GX_REQUIRE(Vy > 0 && Vx > 0 && Vz > 0 && dq > 0.0, "bad voxel grid");
    GX_REQUIRE(Vy < (1 << 21) && Vx < (1 << 21) && Vz < (1 << 21) && (int64_t)Vy * Vx * Vz < (1LL << 32),
               "voxel grid too large for the fp32 filter");
    GX_REQUIRE(n_pix > 0 && n_orient > 0, "empty input");
    DetGrid g;
    g.iq = d_iq; g.Vy = Vy; g.Vx = Vx; g.Vz = Vz;
    g.qx_min = qx_min; g.qy_min = qy_min; g.qz_min = qz_min; g.dq = dq; g.inv_dq = 1.0 / dq;
    const int64_t n_tiles = (n_pix + DETF_THREADS - 1) / DETF_THREADS;
    GX_REQUIRE(n_tiles < 2147483647LL, "too many pixels");
    const int blocks = (int)n_tiles;
    const FastOrient *fo = reinterpret_cast<const FastOrient *>(d_fast);
    const bool probe_on = d_index_out && probe >= 0;
    if (probe_on)
        detector_accumulate_fast_kernel<true><<<blocks, DETF_THREADS, 0, gx_stream(stream)>>>(
            g, d_px, d_py, d_pz, n_pix, fo, d_R, n_orient, d_image, probe, d_index_out, d_slow_count);
    else
        detector_accumulate_fast_kernel<false><<<blocks, DETF_THREADS, 0, gx_stream(stream)>>>(
            g, d_px, d_py, d_pz, n_pix, fo, d_R, n_orient, d_image, probe, d_index_out, d_slow_count);
    return gx_check_launch("gx_detector_accumulate_fast");
}

// Host helper: the per-orientation records of the filtered kernel.
// h_R [n][3][9] (gx_host_orientation_matrices), h_w [n]; pmax[3] = max |p0| per
// component over the detector grid; h_fast receives n records of gx_fast_record_bytes().
extern "C" int gx_fast_record_bytes(void) { return (int)sizeof(FastOrient); }

extern "C" int gx_host_fast_orientations(const double *h_R, const double *h_w, int n, double qx_min,
                                         double qy_min, double qz_min, double dq, const double *pmax,
                                         void *h_fast)
{
    GX_REQUIRE(h_R && h_w && pmax && h_fast && n > 0 && dq > 0.0, "bad arguments");
    FastOrient *out = reinterpret_cast<FastOrient *>(h_fast);
    const double qmin[3] = {qx_min, qy_min, qz_min};
    const double eps = ldexp(1.0, -24);
    for (int o = 0; o < n; ++o) {
        const double *R1 = h_R + (size_t)o * 27, *R2 = R1 + 9, *R3 = R1 + 18;
        double A[9], M[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                double s = 0.0;
                for (int k = 0; k < 3; ++k) s += R2[3 * r + k] * R1[3 * k + c];
                A[3 * r + c] = s;
            }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                double s = 0.0;
                for (int k = 0; k < 3; ++k) s += R3[3 * r + k] * A[3 * k + c];
                M[3 * r + c] = s / dq;
            }
        FastOrient f;
        memset(&f, 0, sizeof(f));
        for (int r = 0; r < 3; ++r) {
            double mag = 0.0;
            for (int c = 0; c < 3; ++c) {
                f.m[3 * r + c] = (float)M[3 * r + c];
                mag += fabs(M[3 * r + c]) * pmax[c];
            }
            const double off = -qmin[r] / dq;
            f.off[r] = (float)off;
            // fp32 rounding of p0, of M, of off and of three FMAs, with a 2x margin, plus the
            // (tiny) difference between the collapsed product and the reference's rounded chain
            const double delta = 12.0 * eps * (mag + fabs(off) + 1.0) + 1e-6;
            f.slack[r] = (float)(0.5 - delta) * (1.0f - 1e-6f);
        }
        for (int k = 0; k < 3; ++k) {
            const double *Rk = R1 + 9 * k;
            bool ident = true;
            for (int e = 0; e < 9; ++e) ident = ident && (Rk[e] == ((e % 4 == 0) ? 1.0 : 0.0));
            if (ident) f.skip |= 1 << k;
        }
        for (int e = 0; e < 9; ++e) f.R1[e] = R1[e];
        f.w = h_w[o];
        out[o] = f;
    }
    return GX_OK;
}

// ------------------------------------------- host: orientation matrices ----
// Pure host arithmetic (compiled without fp contraction; explicit fma where
// OpenBLAS fuses) mirroring, operation by operation:
//   np.cross / np.linalg.norm (ddot = fma chain) / in-place divide
//   rotation_matrix (utilities.py:222-245)
//   R @ corners (fma chain), so the next axis is derived from rotated corners.
static void host_matvec(const double *R, const double *p, double *o)
{
    for (int r = 0; r < 3; ++r) {
        double t = R[3 * r] * p[0];
        t = fma(R[3 * r + 1], p[1], t);
        t = fma(R[3 * r + 2], p[2], t);
        o[r] = t;
    }
}

static void host_normalize(double *u)
{
    double t = u[0] * u[0];
    t = fma(u[1], u[1], t);
    t = fma(u[2], u[2], t);
    double n = sqrt(t);
    u[0] /= n; u[1] /= n; u[2] /= n;
}

static void host_rotation_matrix(const double *u, double c, double s, double *R)
{
    const double ux = u[0], uy = u[1], uz = u[2];
    const double k = 1 - c;
    // each line keeps the reference's association: ((ux*uy)*k) -/+ (uz*s), c + ((ux*ux)*k)
    R[0] = c + (ux * ux) * k;         R[1] = (ux * uy) * k - uz * s;   R[2] = (ux * uz) * k + uy * s;
    R[3] = (uy * ux) * k + uz * s;    R[4] = c + (uy * uy) * k;        R[5] = (uy * uz) * k - ux * s;
    R[6] = (uz * ux) * k - uy * s;    R[7] = (uz * uy) * k + ux * s;   R[8] = c + (uz * uz) * k;
}

extern "C" int gx_host_orientation_matrices(const double *h_corners, const double *h_cs, int n, double *h_R)
{
    GX_REQUIRE(h_corners && h_cs && h_R && n > 0, "bad arguments");
    for (int o = 0; o < n; ++o) {
        double p[3][3];
        for (int i = 0; i < 9; ++i) p[i / 3][i % 3] = h_corners[i];
        for (int step = 0; step < 3; ++step) {
            double across[3], down[3], u[3];
            for (int k = 0; k < 3; ++k) { across[k] = p[1][k] - p[0][k]; down[k] = p[2][k] - p[0][k]; }
            if (step == 0) {          // psi: detector normal = cross(across, down)
                u[0] = across[1] * down[2] - across[2] * down[1];
                u[1] = across[2] * down[0] - across[0] * down[2];
                u[2] = across[0] * down[1] - across[1] * down[0];
            } else if (step == 1) {   // phi: vertical axis
                for (int k = 0; k < 3; ++k) u[k] = down[k];
            } else {                  // theta: horizontal axis
                for (int k = 0; k < 3; ++k) u[k] = across[k];
            }
            host_normalize(u);
            double *R = h_R + ((size_t)o * 3 + step) * 9;
            host_rotation_matrix(u, h_cs[6 * o + 2 * step], h_cs[6 * o + 2 * step + 1], R);
            for (int c = 0; c < 3; ++c) {
                double q[3];
                host_matvec(R, p[c], q);
                p[c][0] = q[0]; p[c][1] = q[1]; p[c][2] = q[2];
            }
        }
    }
    return GX_OK;
}

// -------------------------------------------------------------- epilogue ----
__global__ void __launch_bounds__(256)
detector_epilogue_kernel(const double *__restrict__ img, int rows, int cols, int mirror, int finish, double *out)
{
    const int64_t n = (int64_t)rows * cols;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(o / cols), c = (int)(o % cols);
        double v;
        if (mirror) {
            const int rr = rows - 1 - r, cc = cols - 1 - c;
            const bool mid_r = (rows & 1) && r == rows / 2, mid_c = (cols & 1) && c == cols / 2;
            if (mid_r && mid_c) v = img[o] * 4;
            else if (mid_c) v = (img[o] + img[(int64_t)rr * cols + c]) * 2;       // column rule runs last
            else if (mid_r) v = (img[o] + img[(int64_t)r * cols + cc]) * 2;
            else v = ((img[o] + img[(int64_t)r * cols + cc]) + img[(int64_t)rr * cols + c]) + img[(int64_t)rr * cols + cc];
        } else {
            v = img[o];
        }
        if (finish) {
            if (v != v) v = 1e-6;
            if (v <= 0.0) v = 1e-6;
            if (finish == 1) v = v * 1e-6;
        }
        out[o] = v;
    }
}

extern "C" int gx_detector_epilogue(const double *d_image, int rows, int cols, int mirror, int finish,
                                    double *d_out, void *stream)
{
    GX_REQUIRE(d_image && d_out && rows > 0 && cols > 0, "bad arguments");
    GX_REQUIRE(d_image != d_out, "in-place epilogue is not supported");
    int64_t n = (int64_t)rows * cols;
    int64_t blocks = (n + 255) / 256;
    if (blocks > GX_SM_COUNT * 16) blocks = GX_SM_COUNT * 16;
    detector_epilogue_kernel<<<(int)blocks, 256, 0, gx_stream(stream)>>>(d_image, rows, cols, mirror, finish, d_out);
    return gx_check_launch("gx_detector_epilogue");
}
