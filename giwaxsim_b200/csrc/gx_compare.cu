// Post-hoc transforms of the finished detector image used by the fit loop (second half):
// the polar warp / unwarp pair behind shift_peak (tools/comparison.py:469-592) and the masked linear fit
// of optimize_scale_offset (:873-882).  They act on the trimmed 2-D image (a few hundred pixels a side),
// fp64 like the reference; keeping them on the device lets evaluate_fit / the Nelder-Mead loop
// (fit_slabsize.py:19-47) go from the slab to the residual without a host round trip per transform.
#include <math.h>
#include "gx_common.cuh"

// scipy.ndimage.map_coordinates(order=1, mode='constant', cval): a coordinate outside [0, n-1] on either
// axis yields cval; inside, the bilinear blend of the four neighbours (the upper neighbour of an exact
// edge coordinate has weight 0).  Accumulation order as in ni_interpolation.c: rows outer, columns inner,
// value * w_row * w_col.
__device__ __forceinline__ double sample_linear(const double *__restrict__ img, int rows, int cols, double y, double x,
                                                double cval)
{
    if (!(y >= 0.0 && y <= (double)(rows - 1) && x >= 0.0 && x <= (double)(cols - 1))) return cval;
    const double fy = floor(y), fx = floor(x);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = min(y0 + 1, rows - 1), x1 = min(x0 + 1, cols - 1);
    const double ty = y - fy, tx = x - fx;
    double t = 0.0;
    t += img[(size_t)y0 * cols + x0] * (1.0 - ty) * (1.0 - tx);
    t += img[(size_t)y0 * cols + x1] * (1.0 - ty) * tx;
    t += img[(size_t)y1 * cols + x0] * ty * (1.0 - tx);
    t += img[(size_t)y1 * cols + x1] * ty * tx;
    return t;
}

// linear_polar (comparison.py:469-499): out[i, j] = img(ys, xs), rs = linspace(0, r, out_w)[j],
// ts = linspace(0, 2 pi, out_h)[i], xs = rs cos(ts) + o_col, ys = rs sin(ts) + o_row.
__global__ void __launch_bounds__(256)
polar_warp_kernel(const double *__restrict__ img, int rows, int cols, double o_row, double o_col, double r,
                  int out_h, int out_w, double step_r, double step_t, double cval, double *out)
{
    const size_t n = (size_t)out_h * out_w;
    const double two_pi = 2.0 * 3.14159265358979323846;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / out_w), j = (int)(e - (size_t)i * out_w);
        const double rs = gx_linspace(j, out_w, 0.0, step_r, r);
        const double ts = gx_linspace(i, out_h, 0.0, step_t, two_pi);
        double sn, cs;
        sincos(ts, &sn, &cs);
        const double xs = __dadd_rn(__dmul_rn(rs, cs), o_col), ys = __dadd_rn(__dmul_rn(rs, sn), o_row);
        out[e] = sample_linear(img, rows, cols, ys, xs, cval);
    }
}

// polar_linear (comparison.py:501-539): out[y, x] = polar(ts, rs) with the radius / angle of (y - o_row,
// x - o_col) scaled to the polar image's index space.
__global__ void __launch_bounds__(256)
polar_unwarp_kernel(const double *__restrict__ polar, int ph, int pw, double r, double o_row, double o_col,
                    int out_h, int out_w, double cval, double *out)
{
    const size_t n = (size_t)out_h * out_w;
    const double two_pi = 2.0 * 3.14159265358979323846;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(e / out_w), x = (int)(e - (size_t)y * out_w);
        const double dy = (double)y - o_row, dx = (double)x - o_col;
        double rs = sqrt(__dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx)));
        double ts = atan2(dy, dx);
        if (ts < 0.0) ts += two_pi;
        rs = __dmul_rn(__ddiv_rn(rs, r), (double)(pw - 1));
        ts = __dmul_rn(__ddiv_rn(ts, two_pi), (double)(ph - 1));
        out[e] = sample_linear(polar, ph, pw, ts, rs, cval);
    }
}

// add_pad + the zero mask of shift_peak (comparison.py:541-585): out[i, j] = src[i, map[j]], forced to 0
// where zero_ref[i, j] == 0.
__global__ void __launch_bounds__(256)
gather_columns_kernel(const double *__restrict__ src, int rows, int src_cols, const int32_t *__restrict__ map,
                      int out_cols, const double *__restrict__ zero_ref, double *out)
{
    const size_t n = (size_t)rows * out_cols;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / out_cols), j = (int)(e - (size_t)i * out_cols);
        double v = src[(size_t)i * src_cols + map[j]];
        if (zero_ref && zero_ref[e] == 0.0) v = 0.0;
        out[e] = v;
    }
}

// sums of the normal equations of  min |scale x + offset - y|^2  over the pixels with mask == 0
// (optimize_scale_offset, comparison.py:873-882): out5 = {n, sum x, sum y, sum x x, sum x y}
__global__ void __launch_bounds__(256)
masked_fit_sums_kernel(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ mask,
                       size_t n, double *out5)
{
    double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        if (mask[e] == 0.0) {
            const double a = x[e], b = y[e];
            s[0] += 1.0; s[1] += a; s[2] += b; s[3] += a * a; s[4] += a * b;
        }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    }
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 5; ++k) atomicAdd(out5 + k, s[k]);
}

static int blocks_for(size_t n)
{
    size_t b = (n + 255) / 256;
    if (b > (size_t)GX_SM_COUNT * 8) b = (size_t)GX_SM_COUNT * 8;
    return b < 1 ? 1 : (int)b;
}

extern "C" int gx_polar_warp(const double *d_img, int rows, int cols, double o_row, double o_col, double r,
                             int out_h, int out_w, double cval, double *d_out, void *stream)
{
    GX_REQUIRE(d_img && d_out && rows > 0 && cols > 0 && out_h > 0 && out_w > 0, "bad arguments");
    const double two_pi = 2.0 * 3.14159265358979323846;
    const double step_r = out_w > 1 ? r / (double)(out_w - 1) : 0.0;          // np.linspace step
    const double step_t = out_h > 1 ? two_pi / (double)(out_h - 1) : 0.0;
    polar_warp_kernel<<<blocks_for((size_t)out_h * out_w), 256, 0, gx_stream(stream)>>>(
        d_img, rows, cols, o_row, o_col, r, out_h, out_w, step_r, step_t, cval, d_out);
    return gx_check_launch("gx_polar_warp");
}

extern "C" int gx_polar_unwarp(const double *d_polar, int ph, int pw, double r, double o_row, double o_col,
                               int out_h, int out_w, double cval, double *d_out, void *stream)
{
    GX_REQUIRE(d_polar && d_out && ph > 0 && pw > 0 && out_h > 0 && out_w > 0 && r > 0.0, "bad arguments");
    polar_unwarp_kernel<<<blocks_for((size_t)out_h * out_w), 256, 0, gx_stream(stream)>>>(
        d_polar, ph, pw, r, o_row, o_col, out_h, out_w, cval, d_out);
    return gx_check_launch("gx_polar_unwarp");
}

extern "C" int gx_gather_columns(const double *d_src, int rows, int src_cols, const int32_t *d_map, int out_cols,
                                 const double *d_zero_ref, double *d_out, void *stream)
{
    GX_REQUIRE(d_src && d_map && d_out && rows > 0 && src_cols > 0 && out_cols > 0, "bad arguments");
    gather_columns_kernel<<<blocks_for((size_t)rows * out_cols), 256, 0, gx_stream(stream)>>>(
        d_src, rows, src_cols, d_map, out_cols, d_zero_ref, d_out);
    return gx_check_launch("gx_gather_columns");
}

extern "C" int gx_masked_fit_sums(const double *d_x, const double *d_y, const double *d_mask, int64_t n,
                                  double *d_out5, void *stream)
{
    GX_REQUIRE(d_x && d_y && d_mask && d_out5 && n > 0, "bad arguments");
    GX_CUDA(cudaMemsetAsync(d_out5, 0, 5 * sizeof(double), gx_stream(stream)));
    masked_fit_sums_kernel<<<blocks_for((size_t)n), 256, 0, gx_stream(stream)>>>(d_x, d_y, d_mask, (size_t)n, d_out5);
    return gx_check_launch("gx_masked_fit_sums");
}
