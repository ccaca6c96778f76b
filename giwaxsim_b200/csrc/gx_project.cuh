// Device code shared by the staged projection kernel (gx_project.cu) and the
// fused slice kernels (gx_fused.cu): per-row atom scatter with species counting
// and the background / edge-blend completion of a pixel (voxelgrids.py:338-379).
#pragma once
#include "gx_common.cuh"

#define PROJ_THREADS 256

// -------------------------------------------------------------- row build ----
struct ProjArgs {
    const double *xs, *ys;
    const uint8_t *species;
    const float2 *f;
    const int32_t *row_start;
    const float2 *table;
    int n_species;
    const double *sn, *cs, *yrange;
    const int32_t *bbox;
    const float2 *base;
    const float *my, *mz;
    int N;
    double r;
    float ped_re, ped_im;
    int fill_bkg, sigma;
};

// Accumulate the atoms of row z of rotation p into acc[0..N) (complex64,
// shared).  words: ceil(n_species/2)*N u32 counters (species path) or unused.
// On return acc holds sum of f over the atoms of each pixel; caller syncs.
template <bool SPECIES>
__device__ __forceinline__ void scatter_row(const ProjArgs &a, int p, int z, float2 *acc, uint32_t *words,
                                            const float2 *s_table)
{
    const int N = a.N, tid = threadIdx.x, nt = blockDim.x;
    const double s = a.sn[p], c = a.cs[p], shift = a.yrange[2 * p], r = a.r, inv_r = 1.0 / a.r;
    const int nwords = SPECIES ? ((a.n_species + 1) >> 1) * N : 0;
    for (int y = tid; y < N; y += nt) acc[y] = make_float2(0.f, 0.f);
    if (SPECIES) for (int y = tid; y < nwords; y += nt) words[y] = 0u;
    __syncthreads();
    const int beg = a.row_start[z], end = a.row_start[z + 1];
    for (int c0 = beg; c0 < end; c0 += 65535) {
        const int c1 = min(c0 + 65535, end);
        for (int i = c0 + tid; i < c1; i += nt) {
            double q = gx_floordiv(__dsub_rn(gx_rot_y(a.xs[i], a.ys[i], s, c), shift), r, inv_r);
            if (q < (double)N) {
                const int yi = (int)q;
                if (SPECIES) {
                    const int sp = a.species[i];
                    atomicAdd(&words[(sp >> 1) * N + yi], 1u << ((sp & 1) * 16));
                } else {
                    const float2 f = a.f[i];
                    atomicAdd(&acc[yi].x, f.x);
                    atomicAdd(&acc[yi].y, f.y);
                }
            }
        }
        if (SPECIES) {
            __syncthreads();
            const int npair = (a.n_species + 1) >> 1;
            for (int y = tid; y < N; y += nt) {
                float2 v = acc[y];
                for (int w = 0; w < npair; ++w) {
                    const uint32_t cnt = words[w * N + y];
                    if (cnt) {
                        const float n0 = (float)(cnt & 0xffffu), n1 = (float)(cnt >> 16);
                        const float2 f0 = s_table[2 * w], f1 = s_table[2 * w + 1];
                        v.x += n0 * f0.x + n1 * f1.x;
                        v.y += n0 * f0.y + n1 * f1.y;
                        words[w * N + y] = 0u;
                    }
                }
                acc[y] = v;
            }
        }
        __syncthreads();
    }
}

// Complete pixel (z,y): value relative to the pedestal:
//   inside the atom bbox interior : atoms + num_missing*avg_f - P   (fill_bkg)
//   outside                       : 0                               (fill_bkg)
//   no fill_bkg                   : atoms - P_eff
// times the blend mask when smooth > 0.
__device__ __forceinline__ float2 finish_pixel(const ProjArgs &a, int p, int z, int y, float2 atoms,
                                               const int4 &bb, float mzv)
{
    float2 v;
    const float2 b = a.base[(size_t)p * a.N + y];
    if (a.fill_bkg) {
        const bool inside = (z >= bb.z) && (z < bb.w) && (y > bb.x) && (y < bb.y);
        v = inside ? make_float2(atoms.x + b.x, atoms.y + b.y) : make_float2(0.f, 0.f);
    } else {
        v = make_float2(atoms.x + b.x, atoms.y + b.y);
    }
    if (a.sigma > 0) {
        const float m = mzv * a.my[(size_t)p * a.N + y];
        v.x *= m; v.y *= m;
    }
    return v;
}

