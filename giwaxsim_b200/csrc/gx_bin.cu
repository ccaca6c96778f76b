// K3 / K3b: binning of slice intensities into the 3-D voxel grid and the
// finalisation (average, crop, f0 weighting).
//   voxelgrids.py:396-401,464-506 ; comparison.py:765-786 ; voxelgrids.py:16-48,828-857
//
// Index layout follows the reference: grids are [qy][qx][qz] with qz fastest.
// A slice column fixes (iy,ix), a slice row fixes iz, so consecutive rows of
// one column hit consecutive iz addresses: warps are laid out along rows and
// the fp32 RED.ADDs coalesce.  Counts are rank-1 per slice (H[iy,ix] x m[iz]);
// the driver keeps only H (u32, q_num^2) and m (q_num) unless the caller bins
// with varying row tables, in which case the full 3-D u32 count grid is used.
#include "gx_common.cuh"

__device__ __forceinline__ int bin_index(double v, double qmin, double qmax, double dq, double inv_dq,
                                         int q_num, bool &ok)
{
    ok = (v <= qmax) && (v >= qmin);
    double q = gx_floordiv(__dsub_rn(v, qmin), dq, inv_dq);
    int i = (int)q;
    if (i < 0 || i >= q_num) ok = false;   // the reference would raise IndexError here
    return i;
}

__global__ void slice_col_index_kernel(const double *__restrict__ xl, const double *__restrict__ xr,
                                       const double *__restrict__ yl, const double *__restrict__ yr,
                                       int N, double qmin, double qmax, double dq, int q_num, int32_t *col)
{
    const int p = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const double inv_dq = 1.0 / dq;
    // np.linspace: step = (stop - start) / (N - 1), value = j*step + start, last = stop
    const double x0 = xl[p], x1 = xr[p], y0 = yl[p], y1 = yr[p];
    const double div = (double)(N - 1);
    const double sx = __ddiv_rn(__dsub_rn(x1, x0), div), sy = __ddiv_rn(__dsub_rn(y1, y0), div);
    const double qx = gx_linspace(j, N, x0, sx, x1), qy = gx_linspace(j, N, y0, sy, y1);
    bool okx, oky;
    const int ix = bin_index(qx, qmin, qmax, dq, inv_dq, q_num, okx);
    const int iy = bin_index(qy, qmin, qmax, dq, inv_dq, q_num, oky);
    col[(size_t)p * N + j] = (okx && oky) ? iy * q_num + ix : -1;
}

extern "C" int gx_slice_col_index(const double *d_qx_left, const double *d_qx_right,
                                  const double *d_qy_left, const double *d_qy_right, int n_phi, int N,
                                  double qmin, double qmax, double dq, int q_num, int32_t *d_col, void *stream)
{
    GX_REQUIRE(d_qx_left && d_qx_right && d_qy_left && d_qy_right && d_col, "NULL pointer");
    GX_REQUIRE(n_phi > 0 && N > 1 && q_num > 0 && dq > 0.0, "bad sizes");
    GX_REQUIRE((int64_t)q_num * q_num < 2147483647LL, "q_num too large");
    slice_col_index_kernel<<<dim3((N + 255) / 256, n_phi), 256, 0, gx_stream(stream)>>>(
        d_qx_left, d_qx_right, d_qy_left, d_qy_right, N, qmin, qmax, dq, q_num, d_col);
    return gx_check_launch("gx_slice_col_index");
}

__global__ void axis_col_index_kernel(const double *__restrict__ qx, const double *__restrict__ qy, int n,
                                      double qmin, double qmax, double dq, int q_num, int32_t *col)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double inv_dq = 1.0 / dq;
    bool okx, oky;
    const int ix = bin_index(qx[j], qmin, qmax, dq, inv_dq, q_num, okx);
    const int iy = bin_index(qy[j], qmin, qmax, dq, inv_dq, q_num, oky);
    col[j] = (okx && oky) ? iy * q_num + ix : -1;
}

extern "C" int gx_axis_col_index(const double *d_qx, const double *d_qy, int n, double qmin, double qmax,
                                 double dq, int q_num, int32_t *d_col, void *stream)
{
    GX_REQUIRE(d_qx && d_qy && d_col, "NULL pointer");
    GX_REQUIRE(n > 0 && q_num > 0 && dq > 0.0, "bad sizes");
    axis_col_index_kernel<<<(n + 255) / 256, 256, 0, gx_stream(stream)>>>(d_qx, d_qy, n, qmin, qmax, dq, q_num, d_col);
    return gx_check_launch("gx_axis_col_index");
}

__global__ void axis_row_index_kernel(const double *__restrict__ qz, int n, double qmin, double qmax,
                                      double dq, int q_num, int32_t *row)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    bool ok;
    const int iz = bin_index(qz[j], qmin, qmax, dq, 1.0 / dq, q_num, ok);
    row[j] = ok ? iz : -1;
}

extern "C" int gx_axis_row_index(const double *d_qz, int n, double qmin, double qmax, double dq, int q_num,
                                 int32_t *d_row, void *stream)
{
    GX_REQUIRE(d_qz && d_row, "NULL pointer");
    GX_REQUIRE(n > 0 && q_num > 0 && dq > 0.0, "bad sizes");
    axis_row_index_kernel<<<(n + 255) / 256, 256, 0, gx_stream(stream)>>>(d_qz, n, qmin, qmax, dq, q_num, d_row);
    return gx_check_launch("gx_axis_row_index");
}

__global__ void row_histogram_kernel(const int32_t *__restrict__ row, int n, uint32_t *m)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n && row[j] >= 0) atomicAdd(&m[row[j]], 1u);
}

extern "C" int gx_row_histogram(const int32_t *d_row, int n, int q_num, uint32_t *d_m, void *stream)
{
    GX_REQUIRE(d_row && d_m && n > 0 && q_num > 0, "bad arguments");
    cudaStream_t st = gx_stream(stream);
    GX_CUDA(cudaMemsetAsync(d_m, 0, (size_t)q_num * sizeof(uint32_t), st));
    row_histogram_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_row, n, d_m);
    return gx_check_launch("gx_row_histogram");
}

// ------------------------------------------------------------ accumulate ----
// block = 32 rows x 8 columns: lanes run along rows (contiguous iz) so the
// REDs coalesce; the 8 columns of a block re-use the row indices.
#define BIN_ROWS 32
#define BIN_COLS 8
__global__ void __launch_bounds__(BIN_ROWS * BIN_COLS)
bin_slices_kernel(const float *__restrict__ iq, int rows, int cols, const int32_t *__restrict__ col,
                  int col_stride, const int32_t *__restrict__ row, int q_num,
                  float *sum, uint32_t *count3, uint32_t *count2)
{
    const int b = blockIdx.z;
    const int r = blockIdx.y * BIN_ROWS + threadIdx.x;
    const int c = blockIdx.x * BIN_COLS + threadIdx.y;
    if (c >= cols) return;
    const int yx = col[(size_t)b * col_stride + c];
    if (yx < 0) return;
    if (count2 && blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(&count2[yx], 1u);
    if (r >= rows) return;
    const int iz = row[r];
    if (iz < 0) return;
    const size_t v = (size_t)yx * q_num + iz;
    atomicAdd(&sum[v], iq[((size_t)b * rows + r) * cols + c]);
    if (count3) atomicAdd(&count3[v], 1u);
}

extern "C" int gx_bin_slices(const float *d_iq2d, int batch, int rows, int cols,
                             const int32_t *d_col, int col_stride, const int32_t *d_row, int q_num,
                             float *d_sum, uint32_t *d_count3, uint32_t *d_count2, void *stream)
{
    GX_REQUIRE(d_iq2d && d_col && d_row && d_sum, "NULL pointer");
    GX_REQUIRE(d_count3 || d_count2, "one of d_count3 / d_count2 is required");
    GX_REQUIRE(batch > 0 && rows > 0 && cols > 0 && q_num > 0, "bad sizes");
    GX_REQUIRE(batch <= 65535, "batch too large for one launch");
    dim3 grid((cols + BIN_COLS - 1) / BIN_COLS, (rows + BIN_ROWS - 1) / BIN_ROWS, batch);
    bin_slices_kernel<<<grid, dim3(BIN_ROWS, BIN_COLS), 0, gx_stream(stream)>>>(
        d_iq2d, rows, cols, d_col, col_stride, d_row, q_num, d_sum, d_count3, d_count2);
    return gx_check_launch("gx_bin_slices");
}

// -------------------------------------------------------------- finalise ----
struct Aff { double a[9]; double Z; };

// exp(-b k (qx^2+qy^2+qz^2)) = e_b(qx) e_b(qy) e_b(qz): the four Gaussians of the Cromer-Mann sum are
// tabulated per axis value (4 x V exps evaluated in fp64, once per block) instead of per voxel.
// The kernel is a stream: 4 B read + 4 B written per voxel, so the per-voxel arithmetic is fp32 on
// fp64-prepared factors (table entries, 1/count per column and per row): the weight
// ((sum a_t e_t + c)/Z)^2 and the quotient carry a few fp32 ulps (~3e-7 relative; the result is
// fp32 and the bar is 1e-4 of the maximum).  The first version did two fp64 divisions and eight
// fp64 multiply-adds per voxel and ran at 0.96 TB/s (522 us for 403^3); this one is bound by HBM.
// One warp per (iy, ix) column, lanes along iz (contiguous addresses).  [col_begin, col_end) selects
// a slab of columns, so that N ranks can finalise 1/N of the grid each (reduce-scatter design).
#define FIN_THREADS 256
__global__ void __launch_bounds__(FIN_THREADS)
voxel_finalize_kernel(const float *__restrict__ sum, const uint32_t *__restrict__ count3,
                      const uint32_t *__restrict__ count2, const uint32_t *__restrict__ m,
                      int q_num, int lo, int V, const double *__restrict__ axis, Aff aff, int weighted,
                      int col_begin, int col_end, float *iq)
{
    extern __shared__ float4 s_e[];                 // [V] (e_0..e_3)(q_i) then [V] floats 1/m
    float *s_rm = reinterpret_cast<float *>(s_e + V);
    const double k = 1.0 / (16.0 * 3.14159265358979323846 * 3.14159265358979323846);
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
        const double q = axis[i + lo];
        const double q2k = (q * q) * k;
        s_e[i] = weighted ? make_float4((float)exp(-aff.a[1] * q2k), (float)exp(-aff.a[3] * q2k),
                                        (float)exp(-aff.a[5] * q2k), (float)exp(-aff.a[7] * q2k))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t mi = m ? m[i + lo] : 1u;
        s_rm[i] = mi ? (float)(1.0 / (double)mi) : 0.f;
    }
    __syncthreads();
    const float a0 = (float)(aff.a[0] / aff.Z), a1 = (float)(aff.a[2] / aff.Z), a2 = (float)(aff.a[4] / aff.Z),
                a3 = (float)(aff.a[6] / aff.Z), c = (float)(aff.a[8] / aff.Z);
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int col = col_begin + blockIdx.x * wpb + (threadIdx.x >> 5); col < col_end; col += gridDim.x * wpb) {
        const int iy = col / V, ix = col - iy * V;
        const size_t yx = (size_t)(iy + lo) * q_num + (ix + lo);
        const float *src = sum + yx * q_num + lo;
        float *dst = iq + (size_t)col * V;
        if (count3) {
            // explicit 3-D counts (worker API / small grids): exact quotient, weight in fp32
            const uint32_t *c3 = count3 + yx * q_num + lo;
            const float4 ex = s_e[ix], ey = s_e[iy];
            for (int iz = lane; iz < V; iz += 32) {
                const uint32_t cnt = c3[iz];
                float out = 0.f;
                if (cnt) {
                    const double qv = (double)src[iz] / (double)cnt;
                    if (weighted) {
                        const float4 ez = s_e[iz];
                        const float f = fmaf(a0, ex.x * ey.x * ez.x, fmaf(a1, ex.y * ey.y * ez.y,
                                        fmaf(a2, ex.z * ey.z * ez.z, fmaf(a3, ex.w * ey.w * ez.w, c))));
                        out = (float)(qv * (double)f * (double)f);
                    } else {
                        out = (float)qv;
                    }
                }
                dst[iz] = out;
            }
            continue;
        }
        const uint32_t cyx = count2[yx];
        if (cyx == 0u) {
            for (int iz = lane; iz < V; iz += 32) dst[iz] = 0.f;
            continue;
        }
        if (!weighted) {
            // generate_voxel_grid_low_mem: plain sum/count, correctly rounded
            for (int iz = lane; iz < V; iz += 32) {
                const uint32_t mi = m[iz + lo];
                dst[iz] = mi ? (float)((double)src[iz] / ((double)cyx * (double)mi)) : 0.f;
            }
            continue;
        }
        const float rc = (float)(1.0 / (double)cyx);
        const float4 ex = s_e[ix], ey = s_e[iy];
        const float e0 = a0 * (ex.x * ey.x), e1 = a1 * (ex.y * ey.y), e2 = a2 * (ex.z * ey.z), e3 = a3 * (ex.w * ey.w);
        for (int iz = lane; iz < V; iz += 32) {
            const float4 ez = s_e[iz];
            const float f = fmaf(e0, ez.x, fmaf(e1, ez.y, fmaf(e2, ez.z, fmaf(e3, ez.w, c))));
            dst[iz] = (src[iz] * rc) * s_rm[iz] * (f * f);
        }
    }
}

extern "C" int gx_voxel_finalize(const float *d_sum, const uint32_t *d_count3, const uint32_t *d_count2,
                                 const uint32_t *d_m, int q_num, int lo, int hi, const double *d_axis,
                                 const double *h_aff9, double Z, int64_t col_begin, int64_t col_end,
                                 float *d_iq, void *stream)
{
    GX_REQUIRE(d_sum && d_axis && d_iq, "NULL pointer");
    GX_REQUIRE(d_count3 || (d_count2 && d_m), "count grids missing");
    GX_REQUIRE(q_num > 0 && lo >= 0 && hi > lo && hi <= q_num && (Z != 0.0 || !h_aff9), "bad crop range");
    Aff aff;
    for (int i = 0; i < 9; ++i) aff.a[i] = h_aff9 ? h_aff9[i] : 0.0;
    aff.Z = h_aff9 ? Z : 1.0;
    const int V = hi - lo;
    const int64_t columns = (int64_t)V * V;
    if (col_end < 0) col_end = columns;                       // whole grid
    GX_REQUIRE(col_begin >= 0 && col_begin <= col_end && col_end <= columns && columns < (int64_t)1 << 31,
               "bad column range");
    if (col_begin == col_end) return GX_OK;
    const size_t smem = (size_t)V * (sizeof(float4) + sizeof(float));
    if (smem > 200 * 1024) {
        gx_set_error("gx_voxel_finalize: cropped grid side %d too large", V);
        return GX_ERR_UNSUPPORTED;
    }
    GX_CUDA(cudaFuncSetAttribute(voxel_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int wpb = FIN_THREADS / 32;
    int64_t blocks = (col_end - col_begin + wpb - 1) / wpb;
    if (blocks > (int64_t)GX_SM_COUNT * 8) blocks = (int64_t)GX_SM_COUNT * 8;
    voxel_finalize_kernel<<<(int)blocks, FIN_THREADS, smem, gx_stream(stream)>>>(
        d_sum, d_count3, d_count2, d_m, q_num, lo, V, d_axis, aff, h_aff9 != nullptr, (int)col_begin, (int)col_end,
        d_iq);
    return gx_check_launch("gx_voxel_finalize");
}

// ---------------------------------------------------- shell scaling (N4) ----
// aff_num_qs > 1 branch of generate_voxel_grid_low_mem (voxelgrids.py:650-707): voxels with
// lower < |q| <= upper are multiplied by `factor`.  |q| is the reference's
// sqrt(qx_mesh**2 + qy_mesh**2 + qz_mesh**2) with every operation rounded separately; the grid is
// indexed [qy, qx, qz] (np.meshgrid default 'xy').
__global__ void __launch_bounds__(256)
voxel_shell_scale_kernel(float *iq, int V, const double *__restrict__ axis, double lower, double upper, float factor)
{
    const int columns = V * V;
    for (int col = blockIdx.x; col < columns; col += gridDim.x) {
        const int iy = col / V, ix = col - iy * V;
        const double x = axis[ix], y = axis[iy];
        const double xy = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        for (int iz = threadIdx.x; iz < V; iz += blockDim.x) {
            const double z = axis[iz];
            const double qr = __dsqrt_rn(__dadd_rn(xy, __dmul_rn(z, z)));
            if (qr <= upper && qr > lower) iq[(size_t)col * V + iz] *= factor;
        }
    }
}

extern "C" int gx_voxel_shell_scale(float *d_iq, int V, const double *d_axis, double lower, double upper,
                                    double factor, void *stream)
{
    GX_REQUIRE(d_iq && d_axis && V > 0, "bad arguments");
    int64_t blocks = (int64_t)V * V;
    if (blocks > GX_SM_COUNT * 8) blocks = GX_SM_COUNT * 8;
    voxel_shell_scale_kernel<<<(int)blocks, 256, 0, gx_stream(stream)>>>(d_iq, V, d_axis, lower, upper, (float)factor);
    return gx_check_launch("gx_voxel_shell_scale");
}

// ---------------------------------------------------------- crop window ----
// downselect_voxelgrid (voxelgrids.py:16-48) keeps the voxels lo <= i < hi on
// every axis and the crop commutes with the accumulation, so the production
// driver accumulates only that window: bin indices outside it become -1 (the
// column / row is then neither transformed nor binned) and the others are
// re-based to a [V]^3 grid, V = hi - lo.  packed != 0: entries are iy*q_num+ix.
__global__ void __launch_bounds__(256)
window_indices_kernel(int32_t *idx, int64_t n, int q_num, int lo, int hi, int packed)
{
    const int V = hi - lo;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = idx[i];
        if (v < 0) continue;
        int out;
        if (packed) {
            const int iy = v / q_num, ix = v - iy * q_num;
            out = (iy >= lo && iy < hi && ix >= lo && ix < hi) ? (iy - lo) * V + (ix - lo) : -1;
        } else {
            out = (v >= lo && v < hi) ? v - lo : -1;
        }
        idx[i] = out;
    }
}

extern "C" int gx_window_indices(int32_t *d_index, int64_t n, int q_num, int lo, int hi, int packed, void *stream)
{
    GX_REQUIRE(d_index && n > 0, "bad arguments");
    GX_REQUIRE(q_num > 0 && lo >= 0 && hi > lo && hi <= q_num, "bad window");
    int64_t blocks = (n + 255) / 256;
    if (blocks > GX_SM_COUNT * 16) blocks = GX_SM_COUNT * 16;
    window_indices_kernel<<<(int)blocks, 256, 0, gx_stream(stream)>>>(d_index, n, q_num, lo, hi, packed);
    return gx_check_launch("gx_window_indices");
}
