// Device code shared by the staged projection kernel (gx_project.cu) and the
// fused slice kernels (gx_fused.cu): per-row atom scatter with species counting
// and the background / edge-blend completion of a pixel (voxelgrids.py:338-379).
#pragma once
#include "gx_common.cuh"

#define PROJ_THREADS 256

struct ProjArgs {
    const double *xs, *ys;
    const uint8_t *species;
    const float2 *f;
    const int32_t *row_start;
    const float2 *table;
    int n_species;
    const double *sn, *cs, *yrange;
    const int32_t *bbox;
    const float2 *base;      // [n_phi][N]  num_missing*avg_f - P   (or -P_eff)
    const float *my, *mz;    // [n_phi][N]  blend mask x "inside the atom box" (always present)
    int N;
    double r;
    float ped_re, ped_im;
    int fill_bkg, sigma;
};

// streaming loads that do not allocate in L1 (atoms are read once per CTA; L1 is
// kept for the twiddle table and the per-rotation vectors)
__device__ __forceinline__ double ld_stream_f64(const double *p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned ld_stream_u8(const uint8_t *p)
{
    unsigned v;
    asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// y pixel of one atom of rotation (s, c, shift): (y' - min y') // r as double
__device__ __forceinline__ double atom_y_pixel(double x, double y, double s, double c, double shift,
                                               double r, double inv_r)
{
    return gx_floordiv(__dsub_rn(gx_rot_y(x, y, s, c), shift), r, inv_r);
}

// y pixel as an integer: floor-divide with the +-1 fix-up applied to the converted integer
// (same value as atom_y_pixel; saturates for absurdly distant atoms, which then fail `< N`)
__device__ __forceinline__ int atom_y_pixel_int(double x, double y, double s, double c, double shift,
                                                double r, double inv_r)
{
    const double a = __dsub_rn(gx_rot_y(x, y, s, c), shift);
    const double q = floor(__dmul_rn(a, inv_r));
    const double rem = __fma_rn(-q, r, a);
    int qi = __double2int_rz(q);
    qi += (rem >= r) ? 1 : 0;
    qi -= (rem < 0.0) ? 1 : 0;
    return qi;
}

// U atoms per thread, loads first, NO branch around an atom: an index past the row end re-reads the
// last atom (clamped) and only the final ATOMS is neutralised (adds 0 to word 0), so the U ~10-deep
// fp64 dependency chains interleave instead of running one after the other.
template <int U, bool TAIL>
__device__ __forceinline__ void scatter_batch(const ProjArgs &a, int i0, int end, int nt, double s, double c,
                                              double shift, double r, double inv_r, uint32_t *words, int NP)
{
    const int N = a.N, last = end - 1;
    double x[U], y[U];
    unsigned sp[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int i = TAIL ? min(i0 + u * nt, last) : i0 + u * nt;
        x[u] = ld_stream_f64(a.xs + i);
        y[u] = ld_stream_f64(a.ys + i);
        sp[u] = ld_stream_u8(a.species + i);
    }
    int q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = atom_y_pixel_int(x[u], y[u], s, c, shift, r, inv_r);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const bool ok = (!TAIL || i0 + u * nt < end) && ((unsigned)q[u] < (unsigned)N);
        const unsigned word = ok ? (sp[u] >> 1) * NP + q[u] : 0u;
        const unsigned inc = ok ? 1u << ((sp[u] & 1u) * 16) : 0u;
        atomicAdd(&words[word], inc);
    }
}

// Count atoms [beg,end) of one z-row into the species counters: word plane sp>>1 (stride NP
// words), 16-bit field sp&1.  Whole batches of 4 x blockDim atoms first, then the remainder in
// batches of 2 and 1 (a row of 2441 atoms costs 2 x 4 + 1 + a partial single instead of 3 x 4:
// the clamped tail of a 4-batch did the full arithmetic for atoms that do not exist).
__device__ __forceinline__ void scatter_species(const ProjArgs &a, int beg, int end, double s, double c,
                                                double shift, uint32_t *words, int NP)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const double r = a.r, inv_r = 1.0 / a.r;
    int i0 = beg + tid;
    int left = end - beg;                                   // atoms not yet claimed by a batch (CTA-uniform)
    for (; left >= 4 * nt; left -= 4 * nt, i0 += 4 * nt)
        scatter_batch<4, false>(a, i0, end, nt, s, c, shift, r, inv_r, words, NP);
    if (left >= 2 * nt) {
        scatter_batch<2, false>(a, i0, end, nt, s, c, shift, r, inv_r, words, NP);
        left -= 2 * nt; i0 += 2 * nt;
    }
    if (left >= nt) {
        scatter_batch<1, false>(a, i0, end, nt, s, c, shift, r, inv_r, words, NP);
        left -= nt; i0 += nt;
    }
    if (i0 < end) scatter_batch<1, false>(a, i0, end, nt, s, c, shift, r, inv_r, words, NP);   // per-thread tail
}

// value of a pixel relative to the pedestal P:
//   (atoms + base[y]) * mask,   mask = mz[z] * my[y]
// base = num_missing*avg_f - P (fill_bkg) or -P_eff; the masks carry both the
// Gaussian edge blend and the "outside the atom box -> exactly P" overwrite of
// voxelgrids.py:358-361 (mask 0 there), see gx_slice_vectors.
__device__ __forceinline__ float2 finish_pixel(float2 atoms, float2 b, float m)
{
    return make_float2((atoms.x + b.x) * m, (atoms.y + b.y) * m);
}
