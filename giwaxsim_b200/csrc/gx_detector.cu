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
        if (!live) continue;
        for (int o = 0; o < nc; ++o) {
            const double *m = s_R[o];
            double x1, y1, z1, x2, y2, z2, x3, y3, z3;
            matvec_chain(m, x0, y0, z0, x1, y1, z1);
            matvec_chain(m + 9, x1, y1, z1, x2, y2, z2);
            matvec_chain(m + 18, x2, y2, z2, x3, y3, z3);
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
            v = v * 1e-6;
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
