// K2: batched 2-D FFT with fused |.|^2 and fftshift  (voxelgrids.py:388-392).
//
// Two launches per batch, both staging whole 1-D transforms in shared memory:
//   fft_rows   one CTA per grid row      : c64 row -> FFT along y -> work[z][ky]
//   fft_cols   one CTA per TC columns    : work[:, ky..ky+TC) -> FFT along z ->
//              |X|^2 written at the fftshift-ed position as fp32
// Transform lengths that are not a power of two (the reference's
// ceil(2 pi / (q r)) = 1048, 2095, ...) go through Bluestein's chirp-z on the
// same engine.  cuFFT is not used anywhere (it is the comparison bar only).
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include "gx_common.cuh"
#include "gx_fft_engine.cuh"

// ------------------------------------------------------------------ plan ----
extern "C" int64_t gx_fft_plan_bytes(int N)
{
    GxFftLayout g = gx_fft_layout(N);
    if (g.M == 0) {
        gx_set_error("gx_fft_plan_bytes: unsupported transform length %d "
                     "(need 16 <= N; powers of two up to 16384, any other N up to 8192)", N);
        return GX_ERR_UNSUPPORTED;
    }
    return (int64_t)g.total * (int64_t)sizeof(float2);
}

static void unit_root(long long num, long long den, double sign, float2 *out)
{
    // exp(sign * 2 pi i * num/den) with num reduced exactly first
    num %= den;
    if (num < 0) num += den;
    double a = 2.0 * M_PI * (double)num / (double)den;
    out->x = (float)cos(a);
    out->y = (float)(sign * sin(a));
}

// host reference forward DFT of length M (fp64, O(M log M) radix-2) used only
// to build the Bluestein filter spectrum
static void host_fft(double *re, double *im, int M)
{
    for (int i = 1, j = 0; i < M; ++i) {
        int bit = M >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
    }
    for (int len = 2; len <= M; len <<= 1) {
        for (int i = 0; i < M; i += len) {
            for (int k = 0; k < len / 2; ++k) {
                double a = -2.0 * M_PI * (double)k / (double)len;
                double wr = cos(a), wi = sin(a);
                double ur = re[i + k], ui = im[i + k];
                double vr = re[i + k + len / 2] * wr - im[i + k + len / 2] * wi;
                double vi = re[i + k + len / 2] * wi + im[i + k + len / 2] * wr;
                re[i + k] = ur + vr; im[i + k] = ui + vi;
                re[i + k + len / 2] = ur - vr; im[i + k + len / 2] = ui - vi;
            }
        }
    }
}

static int host_pos(int L, int k)
{
    int r[4]; int np = gx_sched_radices(L, r);
    int S = 1 << L, p = 0;
    for (int i = 0; i < np; ++i) { S /= r[i]; p += (k % r[i]) * S; k /= r[i]; }
    return p;
}

extern "C" int gx_fft_plan_fill(int N, void *h_plan)
{
    GX_REQUIRE(h_plan != NULL, "h_plan is NULL");
    GxFftLayout g = gx_fft_layout(N);
    if (g.M == 0) { gx_set_error("gx_fft_plan_fill: unsupported transform length %d", N); return GX_ERR_UNSUPPORTED; }
    float2 *tab = (float2 *)h_plan;
    int r[4]; int np = gx_sched_radices(g.L, r);
    int S = g.M;
    for (int p = 0; p < np; ++p) {
        S /= r[p];
        if (S == 1) continue;
        const long long bsz = (long long)S * r[p];
        for (int k = 1; k < r[p]; ++k)
            for (int t = 0; t < S; ++t)
                unit_root((long long)t * k, bsz, -1.0, &tab[g.tw_off[p] + (k - 1) * S + t]);
    }
    if (g.bluestein) {
        const int M = g.M;
        // chirp[n] = exp(-i pi n^2 / N) = exp(-2 pi i (n^2 mod 2N) / 2N)
        for (int n = 0; n < N; ++n)
            unit_root(((long long)n * n) % (2LL * N), 2LL * N, -1.0, &tab[g.chirp_off + n]);
        double *re = (double *)calloc((size_t)M, sizeof(double));
        double *im = (double *)calloc((size_t)M, sizeof(double));
        if (!re || !im) { free(re); free(im); gx_set_error("gx_fft_plan_fill: out of host memory"); return GX_ERR_INVALID; }
        for (int m = 0; m < N; ++m) {
            long long q = ((long long)m * m) % (2LL * N);
            double a = M_PI * (double)q / (double)N;
            re[m] = cos(a); im[m] = sin(a);
            if (m) { re[M - m] = re[m]; im[M - m] = im[m]; }
        }
        host_fft(re, im, M);
        for (int k = 0; k < M; ++k) {
            float2 v; v.x = (float)(re[k] / M); v.y = (float)(im[k] / M);
            tab[g.bhat_off + host_pos(g.L, k)] = v;
        }
        free(re); free(im);
    }
    return GX_OK;
}

// --------------------------------------------------------------- kernels ----
// One CTA per grid row: load, FFT along y in shared memory, store natural order.
template <int L>
__global__ void __launch_bounds__(256)
fft_rows_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, GxFftLayout g,
                const float2 *__restrict__ plan)
{
    extern __shared__ float2 smem[];
    constexpr int M = 1 << L;
    const int N = g.N;
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t row = (size_t)blockIdx.y * N + blockIdx.x;
    const float2 *src = in + row * N;
    for (int n = tid; n < M; n += nt) {
        float2 v = make_float2(0.f, 0.f);
        if (n < N) {
            v = src[n];
            if (g.bluestein) v = gx_cmul(v, plan[g.chirp_off + n]);
        }
        smem[gx_phys(n)] = v;
    }
    __syncthreads();
    gx_dft_block<L, 1, 0>(smem, g, plan, tid, nt);
    float2 *dst = out + row * N;
    for (int k = tid; k < N; k += nt) dst[k] = gx_dft_result<L>(smem, g, plan, k);
}

// One CTA per TC adjacent columns: FFT along z, |X|^2, fftshift on both axes.
template <int L, int TC>
__global__ void __launch_bounds__(512)
fft_cols_abs2_kernel(const float2 *__restrict__ in, float *__restrict__ out, GxFftLayout g,
                     const float2 *__restrict__ plan, float dc_re, float dc_im)
{
    extern __shared__ float2 smem[];
    constexpr int M = 1 << L;
    constexpr int BS = M + (M >> 4) + (M >> 8) + 1;
    const int N = g.N;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int col0 = blockIdx.x * TC;
    const float2 *src = in + (size_t)blockIdx.y * N * N;
    for (int w = tid; w < TC * M; w += nt) {
        const int c = w % TC, n = w / TC;
        float2 v = make_float2(0.f, 0.f);
        if (n < N && col0 + c < N) {
            v = src[(size_t)n * N + col0 + c];
            if (g.bluestein) v = gx_cmul(v, plan[g.chirp_off + n]);
        }
        smem[c * BS + gx_phys(n)] = v;
    }
    __syncthreads();
    gx_dft_block<L, TC, BS>(smem, g, plan, tid, nt);
    float *dst = out + (size_t)blockIdx.y * N * N;
    const int half = N / 2;
    for (int w = tid; w < TC * N; w += nt) {
        const int c = w % TC, k = w / TC;
        const int col = col0 + c;
        if (col >= N) continue;
        float2 v = gx_dft_result<L>(smem + c * BS, g, plan, k);
        if (k == 0 && col == 0) { v.x += dc_re; v.y += dc_im; }
        int kr = k + half; if (kr >= N) kr -= N;
        int kc = col + half; if (kc >= N) kc -= N;
        dst[(size_t)kr * N + kc] = v.x * v.x + v.y * v.y;
    }
}

template <int L, int TC>
static int launch_fft2(const float2 *grid, float2 *work, float *iq, int batch, const GxFftLayout &g,
                       const float2 *plan, float dc_re, float dc_im, cudaStream_t st)
{
    constexpr int M = 1 << L;
    const size_t buf = (size_t)gx_phys_len(M) * sizeof(float2);
    const int N = g.N;
    GX_CUDA(cudaFuncSetAttribute(fft_rows_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)buf));
    GX_CUDA(cudaFuncSetAttribute(fft_cols_abs2_kernel<L, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(buf * TC)));
    int nt_rows = M / 16 < 32 ? 32 : (M / 16 > 256 ? 256 : M / 16);
    fft_rows_kernel<L><<<dim3(N, batch), nt_rows, buf, st>>>(grid, work, g, plan);
    if (int e = gx_check_launch("fft_rows_kernel")) return e;
    int nt_cols = TC * M / 16 < 64 ? 64 : (TC * M / 16 > 512 ? 512 : TC * M / 16);
    fft_cols_abs2_kernel<L, TC><<<dim3((N + TC - 1) / TC, batch), nt_cols, buf * TC, st>>>(work, iq, g, plan, dc_re, dc_im);
    return gx_check_launch("fft_cols_abs2_kernel");
}

extern "C" int gx_fft2_abs2_shift(const gx_float2 *d_grid, gx_float2 *d_work, float *d_iq2d,
                                  int batch, int N, const void *d_plan, double dc_re, double dc_im,
                                  void *stream)
{
    GX_REQUIRE(d_grid && d_work && d_iq2d && d_plan, "NULL pointer");
    GX_REQUIRE(batch > 0, "batch must be positive");
    GxFftLayout g = gx_fft_layout(N);
    if (g.M == 0) { gx_set_error("gx_fft2_abs2_shift: unsupported grid size %d", N); return GX_ERR_UNSUPPORTED; }
    const float2 *grid = reinterpret_cast<const float2 *>(d_grid);
    float2 *work = reinterpret_cast<float2 *>(d_work);
    const float2 *plan = reinterpret_cast<const float2 *>(d_plan);
    cudaStream_t st = gx_stream(stream);
    const float dr = (float)dc_re, di = (float)dc_im;
    switch (g.L) {
    case 4: return launch_fft2<4, 8>(grid, work, d_iq2d, batch, g, plan, dr, di, st);
    case 5: return launch_fft2<5, 8>(grid, work, d_iq2d, batch, g, plan, dr, di, st);
    case 6: return launch_fft2<6, 8>(grid, work, d_iq2d, batch, g, plan, dr, di, st);
    case 7: return launch_fft2<7, 8>(grid, work, d_iq2d, batch, g, plan, dr, di, st);
    case 8: return launch_fft2<8, 8>(grid, work, d_iq2d, batch, g, plan, dr, di, st);
    case 9: return launch_fft2<9, 8>(grid, work, d_iq2d, batch, g, plan, dr, di, st);
    case 10: return launch_fft2<10, 8>(grid, work, d_iq2d, batch, g, plan, dr, di, st);
    case 11: return launch_fft2<11, 8>(grid, work, d_iq2d, batch, g, plan, dr, di, st);
    case 12: return launch_fft2<12, 4>(grid, work, d_iq2d, batch, g, plan, dr, di, st);
    case 13: return launch_fft2<13, 2>(grid, work, d_iq2d, batch, g, plan, dr, di, st);
    case 14: return launch_fft2<14, 1>(grid, work, d_iq2d, batch, g, plan, dr, di, st);
    }
    gx_set_error("gx_fft2_abs2_shift: unsupported log2 size %d", g.L);
    return GX_ERR_UNSUPPORTED;
}
