// K0: per-atom prologue -- coordinate extents, phi-invariant z rows and the
// row sort, per-rotation y' range and index bounding box, parity probe.
// Everything here is fp64 and reproduces NumPy's rounding so that the integer
// pixel indices equal the reference's bit for bit (voxelgrids.py:319-349).
#include <limits.h>
#include "gx_common.cuh"

#define ATOM_THREADS 256

// ---------------------------------------------------------------- min/max ----
__global__ void minmax_init_kernel(unsigned long long *o)
{
    int i = threadIdx.x;
    if (i < 6) o[i] = (i & 1) ? 0ull : ~0ull;   // even: running min, odd: running max
}

__global__ void __launch_bounds__(ATOM_THREADS)
minmax_kernel(const double *__restrict__ c, int64_t A, unsigned long long *o)
{
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double v = c[3 * i + k];
            lo[k] = fmin(lo[k], v);
            hi[k] = fmax(hi[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int off = 16; off; off >>= 1) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], off));
            hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], off));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&o[2 * k], gx_ord(lo[k]));
            atomicMax(&o[2 * k + 1], gx_ord(hi[k]));
        }
    }
}

__global__ void minmax_decode_kernel(unsigned long long *o)
{
    int i = threadIdx.x;
    if (i < 6) reinterpret_cast<double *>(o)[i] = gx_unord(o[i]);
}

extern "C" int gx_coords_minmax(const double *d_coords, int64_t A, double *d_out6, void *stream)
{
    GX_REQUIRE(d_coords && d_out6, "NULL pointer");
    GX_REQUIRE(A > 0, "no atoms");
    cudaStream_t st = gx_stream(stream);
    unsigned long long *o = reinterpret_cast<unsigned long long *>(d_out6);
    minmax_init_kernel<<<1, 32, 0, st>>>(o);
    int64_t blocks = (A + ATOM_THREADS - 1) / ATOM_THREADS;
    if (blocks > GX_SM_COUNT * 8) blocks = GX_SM_COUNT * 8;
    minmax_kernel<<<(int)blocks, ATOM_THREADS, 0, st>>>(d_coords, A, o);
    minmax_decode_kernel<<<1, 32, 0, st>>>(o);
    return gx_check_launch("gx_coords_minmax");
}

// ------------------------------------------------------------- row sort ----
__device__ __forceinline__ int atom_row(double z, double z_min, double r, double inv_r, int N)
{
    // (z - min) // r, as int; rows >= N are the "invalid" bin N
    double q = gx_floordiv(__dsub_rn(z, z_min), r, inv_r);
    return (q >= (double)N) ? N : (int)q;
}

__global__ void __launch_bounds__(ATOM_THREADS)
row_hist_kernel(const double *__restrict__ c, int64_t A, double z_min, double r, int N, int32_t *hist)
{
    const double inv_r = 1.0 / r;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&hist[atom_row(c[3 * i + 2], z_min, r, inv_r, N)], 1);
}

// exclusive scan of n = N+1 bins by one block; also resets the cursors
__global__ void __launch_bounds__(1024)
row_scan_kernel(int32_t *hist_cursor, int32_t *row_start, int n)
{
    __shared__ int32_t warp_sums[32];
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        int i = base + threadIdx.x;
        int32_t v = (i < n) ? hist_cursor[i] : 0;
        int32_t x = v;
        for (int off = 1; off < 32; off <<= 1) {
            int32_t y = __shfl_up_sync(0xffffffffu, x, off);
            if ((threadIdx.x & 31) >= off) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int32_t w = warp_sums[threadIdx.x];
            for (int off = 1; off < 32; off <<= 1) {
                int32_t y = __shfl_up_sync(0xffffffffu, w, off);
                if (threadIdx.x >= off) w += y;
            }
            warp_sums[threadIdx.x] = w;
        }
        __syncthreads();
        int32_t prefix = carry + ((threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0) + x - v;
        if (i < n) { row_start[i] = prefix; hist_cursor[i] = prefix; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) row_start[n] = carry;
}

__global__ void __launch_bounds__(ATOM_THREADS)
row_scatter_kernel(const double *__restrict__ c, int64_t A, double z_min, double r, int N,
                   const uint8_t *__restrict__ species, const float2 *__restrict__ f,
                   int32_t *cursor, double *xs, double *ys, int32_t *perm,
                   uint8_t *species_out, float2 *f_out)
{
    const double inv_r = 1.0 / r;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A; i += (int64_t)gridDim.x * blockDim.x) {
        int row = atom_row(c[3 * i + 2], z_min, r, inv_r, N);
        int pos = atomicAdd(&cursor[row], 1);
        xs[pos] = c[3 * i];
        ys[pos] = c[3 * i + 1];
        perm[pos] = (int32_t)i;
        if (species) species_out[pos] = species[i];
        if (f) f_out[pos] = f[i];
    }
}

extern "C" int gx_atoms_sort_rows(const double *d_coords, int64_t A, double z_min, double r, int N,
                                  const uint8_t *d_species, const gx_float2 *d_f,
                                  double *d_xs, double *d_ys, int32_t *d_perm,
                                  uint8_t *d_species_out, gx_float2 *d_f_out,
                                  int32_t *d_row_start, int32_t *d_cursor, void *stream)
{
    GX_REQUIRE(d_coords && d_xs && d_ys && d_perm && d_row_start && d_cursor, "NULL pointer");
    GX_REQUIRE(A > 0 && A < (int64_t)INT_MAX, "atom count must be in (0, 2^31)");
    GX_REQUIRE(N >= 16 && r > 0.0, "bad grid size or voxel size");
    GX_REQUIRE(!d_species || d_species_out, "d_species_out missing");
    GX_REQUIRE(!d_f || d_f_out, "d_f_out missing");
    cudaStream_t st = gx_stream(stream);
    GX_CUDA(cudaMemsetAsync(d_cursor, 0, (size_t)(N + 2) * sizeof(int32_t), st));
    int64_t blocks = (A + ATOM_THREADS - 1) / ATOM_THREADS;
    if (blocks > GX_SM_COUNT * 8) blocks = GX_SM_COUNT * 8;
    row_hist_kernel<<<(int)blocks, ATOM_THREADS, 0, st>>>(d_coords, A, z_min, r, N, d_cursor);
    row_scan_kernel<<<1, 1024, 0, st>>>(d_cursor, d_row_start, N + 1);
    row_scatter_kernel<<<(int)blocks, ATOM_THREADS, 0, st>>>(
        d_coords, A, z_min, r, N, d_species, reinterpret_cast<const float2 *>(d_f), d_cursor, d_xs, d_ys,
        d_perm, d_species_out, reinterpret_cast<float2 *>(d_f_out));
    return gx_check_launch("gx_atoms_sort_rows");
}

// --------------------------------------------------------------- y range ----
// min / max over all atoms of y' for a chunk of rotations.  Each thread keeps
// YR_K atoms in registers and sweeps the rotations (sin/cos broadcast from
// shared memory); warps merge into per-warp shared slots, one pair of global
// atomics per (block, rotation) at the end.
#define YR_K 8
#define YR_CHUNK 256

__global__ void yrange_init_kernel(unsigned long long *o, int n_phi)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * n_phi) o[i] = (i & 1) ? 0ull : ~0ull;
}

__global__ void __launch_bounds__(ATOM_THREADS)
yrange_kernel(const double *__restrict__ xs, const double *__restrict__ ys, int64_t A,
              const double *__restrict__ d_sin, const double *__restrict__ d_cos, int n_phi,
              unsigned long long *o)
{
    __shared__ double s_sin[YR_CHUNK], s_cos[YR_CHUNK];
    __shared__ double s_lo[YR_CHUNK][ATOM_THREADS / 32], s_hi[YR_CHUNK][ATOM_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int p = threadIdx.x; p < n_phi; p += blockDim.x) {
        s_sin[p] = d_sin[p];
        s_cos[p] = d_cos[p];
    }
    for (int p = threadIdx.x; p < n_phi * (ATOM_THREADS / 32); p += blockDim.x) {
        (&s_lo[0][0])[p] = INFINITY;
        (&s_hi[0][0])[p] = -INFINITY;
    }
    __syncthreads();
    const int64_t tile = (int64_t)ATOM_THREADS * YR_K;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < A; base += (int64_t)gridDim.x * tile) {
        double x[YR_K], y[YR_K];
        bool ok[YR_K];
#pragma unroll
        for (int k = 0; k < YR_K; ++k) {
            int64_t i = base + (int64_t)k * ATOM_THREADS + threadIdx.x;
            ok[k] = i < A;
            x[k] = ok[k] ? xs[i] : 0.0;
            y[k] = ok[k] ? ys[i] : 0.0;
        }
        for (int p = 0; p < n_phi; ++p) {
            const double s = s_sin[p], c = s_cos[p];
            double lo = INFINITY, hi = -INFINITY;
#pragma unroll
            for (int k = 0; k < YR_K; ++k) {
                double v = gx_rot_y(x[k], y[k], s, c);
                if (ok[k]) { lo = fmin(lo, v); hi = fmax(hi, v); }
            }
            for (int off = 16; off; off >>= 1) {
                lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
                hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
            }
            if (lane == 0) {
                s_lo[p][warp] = fmin(s_lo[p][warp], lo);
                s_hi[p][warp] = fmax(s_hi[p][warp], hi);
            }
        }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < n_phi; p += blockDim.x) {
        double lo = INFINITY, hi = -INFINITY;
#pragma unroll
        for (int w = 0; w < ATOM_THREADS / 32; ++w) { lo = fmin(lo, s_lo[p][w]); hi = fmax(hi, s_hi[p][w]); }
        atomicMin(&o[2 * p], gx_ord(lo));
        atomicMax(&o[2 * p + 1], gx_ord(hi));
    }
}

// Few atoms (the convex-hull candidates: ~1e3 .. 1e4), many rotations: one warp per rotation, lanes over the
// atoms.  (yrange_kernel loops over the rotations inside every block: with five blocks' worth of atoms that was
// 137 us per 256 rotations, latency-bound; min / max are exact whatever the order, so the result is the same.)
#define YR_SMALL_MAX 65536
__global__ void __launch_bounds__(ATOM_THREADS)
yrange_small_kernel(const double *__restrict__ xs, const double *__restrict__ ys, int A,
                    const double *__restrict__ d_sin, const double *__restrict__ d_cos, int n_phi, double *out)
{
    const int lane = threadIdx.x & 31;
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= n_phi) return;
    const double s = d_sin[p], c = d_cos[p];
    double lo = INFINITY, hi = -INFINITY;
    for (int i = lane; i < A; i += 32) {
        const double v = gx_rot_y(xs[i], ys[i], s, c);
        lo = fmin(lo, v); hi = fmax(hi, v);
    }
    for (int off = 16; off; off >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if (lane == 0) { out[2 * p] = lo; out[2 * p + 1] = hi; }
}

__global__ void yrange_decode_kernel(unsigned long long *o, int n_phi)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * n_phi) reinterpret_cast<double *>(o)[i] = gx_unord(o[i]);
}

extern "C" int gx_slice_yrange(const double *d_xs, const double *d_ys, int64_t A,
                               const double *d_sin, const double *d_cos, int n_phi,
                               double *d_yrange, void *stream)
{
    GX_REQUIRE(d_xs && d_ys && d_sin && d_cos && d_yrange, "NULL pointer");
    GX_REQUIRE(A > 0 && n_phi > 0, "empty input");
    cudaStream_t st = gx_stream(stream);
    if (A <= YR_SMALL_MAX) {
        const int warps_per_block = ATOM_THREADS / 32;
        yrange_small_kernel<<<(n_phi + warps_per_block - 1) / warps_per_block, ATOM_THREADS, 0, st>>>(
            d_xs, d_ys, (int)A, d_sin, d_cos, n_phi, d_yrange);
        return gx_check_launch("gx_slice_yrange");
    }
    unsigned long long *o = reinterpret_cast<unsigned long long *>(d_yrange);
    yrange_init_kernel<<<(2 * n_phi + 255) / 256, 256, 0, st>>>(o, n_phi);
    const int64_t tile = (int64_t)ATOM_THREADS * YR_K;
    int64_t blocks = (A + tile - 1) / tile;
    if (blocks > GX_SM_COUNT * 2) blocks = GX_SM_COUNT * 2;
    for (int p0 = 0; p0 < n_phi; p0 += YR_CHUNK) {
        int n = n_phi - p0 < YR_CHUNK ? n_phi - p0 : YR_CHUNK;
        yrange_kernel<<<(int)blocks, ATOM_THREADS, 0, st>>>(d_xs, d_ys, A, d_sin + p0, d_cos + p0, n, o + 2 * p0);
    }
    yrange_decode_kernel<<<(2 * n_phi + 255) / 256, 256, 0, st>>>(o, n_phi);
    return gx_check_launch("gx_slice_yrange");
}

// ------------------------------------------------------------------ bbox ----
// Fast path: when no atom can be clipped (all z rows < N and the largest y
// index < N) the bbox is {0, floor-div of the y' span, first row, last row},
// because NumPy's floor-divide is monotone.  Otherwise a full pass over the
// atoms of that rotation evaluates the reference's valid mask.
__global__ void bbox_fast_kernel(const double *__restrict__ yrange, const int32_t *__restrict__ row_start,
                                 int N, double r, int n_phi, int32_t *bbox, int32_t *need_full)
{
    __shared__ int s_first, s_last;
    if (threadIdx.x == 0) { s_first = INT_MAX; s_last = -1; }
    __syncthreads();
    int first = INT_MAX, last = -1;
    for (int z = threadIdx.x; z < N; z += blockDim.x)
        if (row_start[z + 1] > row_start[z]) { first = min(first, z); last = max(last, z); }
    atomicMin(&s_first, first);
    atomicMax(&s_last, last);
    __syncthreads();
    const bool z_clipped = row_start[N + 1] > row_start[N];
    const double inv_r = 1.0 / r;
    int any = 0;
    for (int p = threadIdx.x; p < n_phi; p += blockDim.x) {
        double span = __dsub_rn(yrange[2 * p + 1], yrange[2 * p]);
        double top = gx_floordiv(span, r, inv_r);
        bool full = z_clipped || !(top < (double)N) || s_last < 0;
        any |= full ? 1 : 0;
        need_full[p] = full ? 1 : 0;
        if (full) {
            bbox[4 * p + 0] = INT_MAX; bbox[4 * p + 1] = -1; bbox[4 * p + 2] = INT_MAX; bbox[4 * p + 3] = -1;
        } else {
            bbox[4 * p + 0] = 0; bbox[4 * p + 1] = (int)top; bbox[4 * p + 2] = s_first; bbox[4 * p + 3] = s_last;
        }
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) need_full[n_phi] = any;          // lets the full pass return at once when nothing is clipped
}

#define BBOX_ROWS 16
__global__ void __launch_bounds__(ATOM_THREADS)
bbox_full_kernel(const double *__restrict__ xs, const double *__restrict__ ys,
                 const int32_t *__restrict__ row_start, int N, double r,
                 const double *__restrict__ d_sin, const double *__restrict__ d_cos,
                 const double *__restrict__ yrange, const int32_t *__restrict__ need_full, int n_phi, int32_t *bbox)
{
    // grid.y is bounded (one launch used to carry n_phi x N/16 CTAs that exit at once when no rotation is
    // clipped - the usual case: 241 us of empty blocks per 1800 rotations); a CTA walks its share of the rotations
    if (!need_full[n_phi]) return;
    for (int p = blockIdx.y; p < n_phi; p += gridDim.y) {
    if (!need_full[p]) continue;
    const double s = d_sin[p], c = d_cos[p], shift = yrange[2 * p], inv_r = 1.0 / r;
    int ylo = INT_MAX, yhi = -1, zlo = INT_MAX, zhi = -1;
    const int z0 = blockIdx.x * BBOX_ROWS;
    for (int z = z0; z < min(z0 + BBOX_ROWS, N); ++z) {
        for (int i = row_start[z] + threadIdx.x; i < row_start[z + 1]; i += blockDim.x) {
            double a = __dsub_rn(gx_rot_y(xs[i], ys[i], s, c), shift);
            double q = gx_floordiv(a, r, inv_r);
            if (q < (double)N) {
                int yi = (int)q;
                ylo = min(ylo, yi); yhi = max(yhi, yi); zlo = min(zlo, z); zhi = max(zhi, z);
            }
        }
    }
    for (int off = 16; off; off >>= 1) {
        ylo = min(ylo, __shfl_xor_sync(0xffffffffu, ylo, off));
        yhi = max(yhi, __shfl_xor_sync(0xffffffffu, yhi, off));
        zlo = min(zlo, __shfl_xor_sync(0xffffffffu, zlo, off));
        zhi = max(zhi, __shfl_xor_sync(0xffffffffu, zhi, off));
    }
    if ((threadIdx.x & 31) == 0 && yhi >= 0) {
        atomicMin(&bbox[4 * p + 0], ylo); atomicMax(&bbox[4 * p + 1], yhi);
        atomicMin(&bbox[4 * p + 2], zlo); atomicMax(&bbox[4 * p + 3], zhi);
    }
    }
}

__global__ void bbox_empty_kernel(int32_t *bbox, int n_phi)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_phi && bbox[4 * p + 1] < 0) { bbox[4 * p] = -1; bbox[4 * p + 2] = -1; bbox[4 * p + 3] = -1; }
}

extern "C" int gx_slice_bbox(const double *d_xs, const double *d_ys, const int32_t *d_row_start, int N,
                             double r, const double *d_sin, const double *d_cos, const double *d_yrange,
                             int n_phi, int32_t *d_bbox, int32_t *d_scratch, void *stream)
{
    GX_REQUIRE(d_xs && d_ys && d_row_start && d_sin && d_cos && d_yrange && d_bbox && d_scratch, "NULL pointer");
    GX_REQUIRE(n_phi > 0 && N >= 16, "bad sizes");
    cudaStream_t st = gx_stream(stream);
    int32_t *need_full = d_scratch;
    bbox_fast_kernel<<<1, 1024, 0, st>>>(d_yrange, d_row_start, N, r, n_phi, d_bbox, need_full);
    bbox_full_kernel<<<dim3((N + BBOX_ROWS - 1) / BBOX_ROWS, n_phi < 8 ? n_phi : 8), ATOM_THREADS, 0, st>>>(
        d_xs, d_ys, d_row_start, N, r, d_sin, d_cos, d_yrange, need_full, n_phi, d_bbox);
    bbox_empty_kernel<<<(n_phi + 255) / 256, 256, 0, st>>>(d_bbox, n_phi);
    return gx_check_launch("gx_slice_bbox");
}

// ---------------------------------------------------------- parity probe ----
__global__ void __launch_bounds__(ATOM_THREADS)
pixel_index_kernel(const double *__restrict__ xs, const double *__restrict__ ys,
                   const int32_t *__restrict__ perm, const int32_t *__restrict__ row_start,
                   int64_t A, int N, double r, double s, double c, double shift,
                   int64_t *y_idx, int64_t *z_idx)
{
    const double inv_r = 1.0 / r;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A; i += (int64_t)gridDim.x * blockDim.x) {
        // row of sorted atom i: last z with row_start[z] <= i
        int lo = 0, hi = N + 1;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (row_start[mid] <= (int32_t)i) lo = mid; else hi = mid;
        }
        double a = __dsub_rn(gx_rot_y(xs[i], ys[i], s, c), shift);
        double q = gx_floordiv(a, r, inv_r);
        int32_t o = perm[i];
        y_idx[o] = (int64_t)q;
        z_idx[o] = lo;
    }
}

extern "C" int gx_atom_pixel_indices(const double *d_xs, const double *d_ys, const int32_t *d_perm,
                                     const int32_t *d_row_start, int64_t A, int N, double r,
                                     double sin_phi, double cos_phi, double y_shift,
                                     int64_t *d_y_idx, int64_t *d_z_idx, void *stream)
{
    GX_REQUIRE(d_xs && d_ys && d_perm && d_row_start && d_y_idx && d_z_idx, "NULL pointer");
    GX_REQUIRE(A > 0, "no atoms");
    int64_t blocks = (A + ATOM_THREADS - 1) / ATOM_THREADS;
    if (blocks > GX_SM_COUNT * 8) blocks = GX_SM_COUNT * 8;
    pixel_index_kernel<<<(int)blocks, ATOM_THREADS, 0, gx_stream(stream)>>>(
        d_xs, d_ys, d_perm, d_row_start, A, N, r, sin_phi, cos_phi, y_shift, d_y_idx, d_z_idx);
    return gx_check_launch("gx_atom_pixel_indices");
}

// ------------------------------------------------- extreme-atom candidates ----
// min/max of y' over all atoms is needed for every rotation (voxelgrids.py:323)
// but only atoms on (or within eps of) the boundary of the 2-D convex hull of
// the (x, y) positions can attain it.  The host builds an inner polygon from the
// atoms that are extreme along a few dozen directions (gx_extreme_atoms), and
// gx_hull_filter keeps the atoms that are not strictly inside it by more than
// eps: if an atom is at distance d inside a polygon whose vertices are atoms,
// some vertex beats it by at least d in every direction, so dropping it cannot
// change any rounded min or max as long as d >> 1 ulp of the coordinates.
__global__ void __launch_bounds__(ATOM_THREADS)
extreme_atoms_kernel(const double *__restrict__ xs, const double *__restrict__ ys, int64_t A,
                     const double *__restrict__ d_sin, const double *__restrict__ d_cos,
                     const double *__restrict__ yrange, int n, int32_t *index)
{
    __shared__ double s_sin[64], s_cos[64], s_lo[64], s_hi[64];
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        s_sin[p] = d_sin[p]; s_cos[p] = d_cos[p]; s_lo[p] = yrange[2 * p]; s_hi[p] = yrange[2 * p + 1];
    }
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = xs[i], y = ys[i];
        for (int p = 0; p < n; ++p) {
            const double v = gx_rot_y(x, y, s_sin[p], s_cos[p]);
            if (v == s_lo[p]) atomicMin(&index[2 * p], (int32_t)i);
            if (v == s_hi[p]) atomicMin(&index[2 * p + 1], (int32_t)i);
        }
    }
}

extern "C" int gx_extreme_atoms(const double *d_xs, const double *d_ys, int64_t A,
                                const double *d_sin, const double *d_cos, const double *d_yrange, int n,
                                int32_t *d_index, void *stream)
{
    GX_REQUIRE(d_xs && d_ys && d_sin && d_cos && d_yrange && d_index, "NULL pointer");
    GX_REQUIRE(A > 0 && n > 0 && n <= 64, "need 1..64 directions");
    cudaStream_t st = gx_stream(stream);
    GX_CUDA(cudaMemsetAsync(d_index, 0x7f, (size_t)2 * n * sizeof(int32_t), st));
    int64_t blocks = (A + ATOM_THREADS - 1) / ATOM_THREADS;
    if (blocks > GX_SM_COUNT * 8) blocks = GX_SM_COUNT * 8;
    extreme_atoms_kernel<<<(int)blocks, ATOM_THREADS, 0, st>>>(d_xs, d_ys, A, d_sin, d_cos, d_yrange, n, d_index);
    return gx_check_launch("gx_extreme_atoms");
}

__global__ void __launch_bounds__(ATOM_THREADS)
hull_filter_kernel(const double *__restrict__ xs, const double *__restrict__ ys, int64_t A,
                   const double *__restrict__ edges, int m, double eps, int32_t *count,
                   double *xs_out, double *ys_out, int capacity)
{
    __shared__ double s_e[64 * 3];
    for (int k = threadIdx.x; k < 3 * m; k += blockDim.x) s_e[k] = edges[k];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = xs[i], y = ys[i];
        double depth = INFINITY;                 // signed distance to the nearest edge, inside > 0
        for (int k = 0; k < m; ++k) depth = fmin(depth, s_e[3 * k] * x + s_e[3 * k + 1] * y + s_e[3 * k + 2]);
        if (!(depth > eps)) {
            const int slot = atomicAdd(count, 1);
            if (slot < capacity) { xs_out[slot] = x; ys_out[slot] = y; }
        }
    }
}

extern "C" int gx_hull_filter(const double *d_xs, const double *d_ys, int64_t A, const double *d_edges, int m,
                              double eps, int32_t *d_count, double *d_xs_out, double *d_ys_out, int capacity,
                              void *stream)
{
    GX_REQUIRE(d_xs && d_ys && d_edges && d_count && d_xs_out && d_ys_out, "NULL pointer");
    GX_REQUIRE(A > 0 && m >= 3 && m <= 64 && capacity > 0 && eps > 0.0, "bad polygon or capacity");
    cudaStream_t st = gx_stream(stream);
    GX_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int32_t), st));
    int64_t blocks = (A + ATOM_THREADS - 1) / ATOM_THREADS;
    if (blocks > GX_SM_COUNT * 8) blocks = GX_SM_COUNT * 8;
    hull_filter_kernel<<<(int)blocks, ATOM_THREADS, 0, st>>>(d_xs, d_ys, A, d_edges, m, eps, d_count, d_xs_out,
                                                             d_ys_out, capacity);
    return gx_check_launch("gx_hull_filter");
}

// ------------------------------------------------------ species coding ----
// Element symbols arrive as NumPy '<U1' / '<U2' arrays, i.e. `width` UCS-4 code
// points per atom.  key = cp0 + (cp1 << 7) (14 bits for ASCII symbols).  Pass 1
// histograms the keys (warp-aggregated atomics: a slab has a handful of distinct
// elements, so all lanes hit the same few counters); the host turns the occupied
// bins into a key -> species-code table; pass 2 writes the uint8 codes.
// d_hist: 16385 counters, the last one counts atoms with a non-ASCII code point.
#define SPECIES_KEYS 16384

__device__ __forceinline__ uint32_t species_key(const uint32_t *cp, int width, int64_t i, bool &bad)
{
    const uint32_t c0 = cp[i * width];
    const uint32_t c1 = width > 1 ? cp[i * width + 1] : 0u;
    bad = (c0 | c1) >= 128u;
    return (c0 & 127u) | ((c1 & 127u) << 7);
}

__global__ void __launch_bounds__(ATOM_THREADS)
species_hist_kernel(const uint32_t *__restrict__ cp, int width, int64_t A, uint32_t *hist)
{
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < A; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const bool live = i < A;
        bool bad = false;
        uint32_t key = live ? species_key(cp, width, i, bad) : 0xffffffffu;
        if (live && bad) key = SPECIES_KEYS;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (live && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(hist + key, (uint32_t)__popc(peers));
    }
}

__global__ void __launch_bounds__(ATOM_THREADS)
species_code_kernel(const uint32_t *__restrict__ cp, int width, int64_t A, const uint8_t *__restrict__ lut,
                    uint8_t *codes)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A; i += (int64_t)gridDim.x * blockDim.x) {
        bool bad;
        codes[i] = lut[species_key(cp, width, i, bad)];
    }
}

extern "C" int gx_species_histogram(const uint32_t *d_codepoints, int width, int64_t A, uint32_t *d_hist,
                                    void *stream)
{
    GX_REQUIRE(d_codepoints && d_hist, "NULL pointer");
    GX_REQUIRE((width == 1 || width == 2) && A > 0, "width must be 1 or 2 code points");
    GX_CUDA(cudaMemsetAsync(d_hist, 0, (SPECIES_KEYS + 1) * sizeof(uint32_t), gx_stream(stream)));
    int64_t blocks = (A + ATOM_THREADS - 1) / ATOM_THREADS;
    if (blocks > GX_SM_COUNT * 8) blocks = GX_SM_COUNT * 8;
    species_hist_kernel<<<(int)blocks, ATOM_THREADS, 0, gx_stream(stream)>>>(d_codepoints, width, A, d_hist);
    return gx_check_launch("gx_species_histogram");
}

extern "C" int gx_species_codes(const uint32_t *d_codepoints, int width, int64_t A, const uint8_t *d_lut,
                                uint8_t *d_codes, void *stream)
{
    GX_REQUIRE(d_codepoints && d_lut && d_codes, "NULL pointer");
    GX_REQUIRE((width == 1 || width == 2) && A > 0, "width must be 1 or 2 code points");
    int64_t blocks = (A + ATOM_THREADS - 1) / ATOM_THREADS;
    if (blocks > GX_SM_COUNT * 8) blocks = GX_SM_COUNT * 8;
    species_code_kernel<<<(int)blocks, ATOM_THREADS, 0, gx_stream(stream)>>>(d_codepoints, width, A, d_lut, d_codes);
    return gx_check_launch("gx_species_codes");
}


// ---------------------------------------------------------------- content checksum ----
// Wrap-around sum of the 64-bit words of a device array; fp32 input is widened to fp64 first, so the
// value equals the same sum taken on the host over the float64 copy a driver hands to its caller.
// Used to decide whether a host array still equals the device copy it was made from.
__global__ void __launch_bounds__(ATOM_THREADS)
checksum64_kernel(const void *__restrict__ data, int64_t n, int widen, unsigned long long *out)
{
    unsigned long long acc = 0ull;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (widen) {
        const float *p = static_cast<const float *>(data);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            acc += (unsigned long long)__double_as_longlong((double)p[i]);
    } else {
        const unsigned long long *p = static_cast<const unsigned long long *>(data);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += p[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

extern "C" int gx_checksum64(const void *d_data, int64_t n, int widen_f32, uint64_t *d_out, void *stream)
{
    GX_REQUIRE(d_data && d_out && n >= 0, "bad arguments");
    GX_CUDA(cudaMemsetAsync(d_out, 0, sizeof(uint64_t), gx_stream(stream)));
    if (n == 0) return GX_OK;
    int64_t blocks = (n + ATOM_THREADS - 1) / ATOM_THREADS;
    if (blocks > GX_SM_COUNT * 8) blocks = GX_SM_COUNT * 8;
    checksum64_kernel<<<(int)blocks, ATOM_THREADS, 0, gx_stream(stream)>>>(d_data, n, widen_f32,
                                                                          reinterpret_cast<unsigned long long *>(d_out));
    return gx_check_launch("gx_checksum64");
}


// ------------------------------------------------- per-row sums of |Re f|, |Im f| ----
struct AbsTable { double v[2 * GX_MAX_SPECIES]; };      // kernel-parameter copy of the host table

// Largest sum over the atoms of one z pixel row of |Re f| and of |Im f|: what a single pixel of a row
// can receive at most, which sizes the fixed-point scale of the fused row kernel's integer accumulators
// (gx_fused_args.max_row_abs_re / _im).  One warp per row; atoms are sorted by row.
__global__ void __launch_bounds__(ATOM_THREADS)
row_abs_f_max_kernel(const uint8_t *__restrict__ species, const float2 *__restrict__ f,
                     const int32_t *__restrict__ row_start, int N, AbsTable table_abs,
                     int n_species, unsigned long long *out2)
{
    __shared__ double s_tab[2 * GX_MAX_SPECIES];
    if (threadIdx.x < 2 * GX_MAX_SPECIES) s_tab[threadIdx.x] = (int)threadIdx.x < 2 * n_species ? table_abs.v[threadIdx.x] : 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    double best_re = 0.0, best_im = 0.0;
    for (int z = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; z < N; z += warps) {
        const int beg = row_start[z], end = row_start[z + 1];
        double re = 0.0, im = 0.0;
        for (int i = beg + lane; i < end; i += 32) {
            if (species) { const int sp = species[i]; re += s_tab[2 * sp]; im += s_tab[2 * sp + 1]; }
            else { const float2 v = f[i]; re += fabs((double)v.x); im += fabs((double)v.y); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            re += __shfl_xor_sync(0xffffffffu, re, o);
            im += __shfl_xor_sync(0xffffffffu, im, o);
        }
        best_re = fmax(best_re, re);
        best_im = fmax(best_im, im);
    }
    if (lane == 0) {
        atomicMax(out2, gx_ord(best_re));
        atomicMax(out2 + 1, gx_ord(best_im));
    }
}

__global__ void row_abs_f_init_kernel(unsigned long long *out2)
{
    if (threadIdx.x < 2) out2[threadIdx.x] = gx_ord(0.0);
}

__global__ void row_abs_f_decode_kernel(unsigned long long *out2)
{
    if (threadIdx.x < 2) {
        const double v = gx_unord(out2[threadIdx.x]);
        reinterpret_cast<double *>(out2)[threadIdx.x] = v;
    }
}

extern "C" int gx_row_abs_f_max(const uint8_t *d_species, const gx_float2 *d_f, const int32_t *d_row_start, int N,
                                const double *h_table_abs, int n_species, double *d_out2, void *stream)
{
    GX_REQUIRE((d_species || d_f) && d_row_start && d_out2 && N > 0, "bad arguments");
    GX_REQUIRE(n_species >= 0 && n_species <= GX_MAX_SPECIES && (!d_species || (h_table_abs && n_species > 0)),
               "species table missing");
    cudaStream_t st = gx_stream(stream);
    AbsTable tab;
    memset(&tab, 0, sizeof(tab));
    if (d_species)
        for (int k = 0; k < 2 * n_species; ++k) tab.v[k] = h_table_abs[k];
    row_abs_f_init_kernel<<<1, 32, 0, st>>>(reinterpret_cast<unsigned long long *>(d_out2));
    int blocks = (N * 32 + ATOM_THREADS - 1) / ATOM_THREADS;
    if (blocks > GX_SM_COUNT * 8) blocks = GX_SM_COUNT * 8;
    row_abs_f_max_kernel<<<blocks, ATOM_THREADS, 0, st>>>(d_species, reinterpret_cast<const float2 *>(d_f), d_row_start, N,
                                                           tab, n_species, reinterpret_cast<unsigned long long *>(d_out2));
    row_abs_f_decode_kernel<<<1, 32, 0, st>>>(reinterpret_cast<unsigned long long *>(d_out2));
    return gx_check_launch("gx_row_abs_f_max");
}
