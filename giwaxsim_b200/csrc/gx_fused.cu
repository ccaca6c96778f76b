// Fused stage-A slice pipeline: the production path of voxelgridmaker_fitting.
//
//   F1  slice_rows_fused   one CTA per (rotation, z-row):
//         atoms of the row -> species counters in shared memory (ATOMS.ADD) ->
//         complex row (background, edge blend, pedestal removed) -> 1-D FFT along
//         y in the same shared memory -> only the q-columns the voxel box keeps
//         are written (N x Kc complex64 instead of N x N).
//   F2  slice_cols_fused   one CTA per (rotation, TC kept columns):
//         column tile (rows of the atom band only; the rest is zero once the
//         pedestal is removed) -> 1-D FFT along z -> |.|^2 of the kept q-rows ->
//         fp32 RED.ADD straight into the 3-D voxel sum, u32 per-column count.
//
// What never touches HBM any more (reference: voxelgrids.py:338-392,464-503):
// the zero-filled N x N grid, the scatter read-modify-writes, the pre-FFT grid,
// the row-FFT output outside the kept columns, and the N x N intensity image.
// The constant pedestal P of the background model is subtracted before the
// transform (its spectrum is a delta) and P*N^2 is added back to the DC
// coefficient in F2, which also keeps fp32 round-off away from the Bragg signal.
//
// Grid order: the rotation index is the fastest-varying block index, so the
// CTAs resident at any moment work on the same few z-rows for all rotations of
// the batch and the row's atoms are fetched from HBM once and re-read from L2.
#include <stdlib.h>
#include "gx_project.cuh"
#include "gx_fft_engine.cuh"

#ifndef GX_F2_TC12
#define GX_F2_TC12 4        // columns per F2 CTA at N = 4096 (2 -> 3 CTAs/SM was measured: 1 % slower)
#endif
#ifndef GX_F2_TILE_FAST
#define GX_F2_TILE_FAST 0   // F2 grid order: 0 = rotation fastest, 1 = column tile fastest
#endif
#ifndef GX_PASS1_TWP
#define GX_PASS1_TWP GX_TWP  // twiddles of the middle pass (16 distinct sets, 1.9 KB): 1 = four loads + products, 0 = fifteen loads
#endif
#ifndef GX_F1_MINBLOCKS
#define GX_F1_MINBLOCKS 4   // CTAs/SM the row kernel is compiled for (3 would leave ~84 KB of L1: measured no faster)
#endif

// Per-rotation / per-row scalars of the row kernel, staged through constant memory by gx_slices_fused
// (device-to-device cudaMemcpyToSymbolAsync on the launch stream): a CTA reads them with uniform
// constant-bank loads instead of opening with an L2 round trip before it can even decide whether its
// row is active.  One launch at a time per device may use them (the path runs on one stream per GPU).
#ifndef GX_CONST_BATCH
#define GX_CONST_BATCH 128      // rotations per launch the constant tables hold
#endif
#define GX_CONST_ROWS 8192      // largest grid side
__constant__ double c_sn[GX_CONST_BATCH], c_cs[GX_CONST_BATCH], c_yrange[2 * GX_CONST_BATCH];
__constant__ int32_t c_bbox[4 * GX_CONST_BATCH], c_colrange[2 * GX_CONST_BATCH];
__constant__ int32_t c_row_start[GX_CONST_ROWS + 2];

struct FusedArgs {
    ProjArgs proj;
    const float2 *dmy;         // [n_phi][N] (num_missing - max_voxels, my): base = dmy.x * avg_f
    float af_re, af_im;
    float2 ftab[GX_MAX_SPECIES];   // species f-values in the parameter (constant) bank: FFMA operands
                                   // straight from c[][] cost neither registers nor shared-memory loads
    GxFftLayout lay;
    const float2 *plan;
    const int32_t *col;        // [n_phi][N] packed iy*q_num+ix or -1
    const int32_t *colrange;   // [n_phi][2] first kept column, one past last
    const int32_t *row_index;  // [N] iz or -1
    int row_lo, row_hi;        // kept shifted rows [row_lo, row_hi)
    float2 *work;              // [n_phi][N][KC]
    int KC;
    int q_num;
    float *vsum;
    uint32_t *count2;
    float dc_re, dc_im;        // pedestal * N^2
    int n_phi;
    int use_const;             // row-kernel scalars are in the constant tables
};

__device__ __forceinline__ void active_band(const ProjArgs &a, int p, int &za, int &zb)
{
    // rows whose pedestal-free content can be non-zero
    const int z_min = a.bbox[4 * p + 2], z_max = a.bbox[4 * p + 3];
    if (a.fill_bkg) { za = z_min; zb = z_max - 1; }
    else if (a.sigma > 0) { za = 0; zb = a.N - 1; }
    else { za = z_min; zb = z_max; }
}

// ------------------------------------------------------------------ F1 ----
// Single-chunk rows (<= 65535 atoms, i.e. always except for huge crystalline rows): the pixels of
// this thread's first butterfly are produced in one go from the species counters,
//     v = (sum_w n0 f0 + n1 f1 + d * avg_f) * mz * my,
// pixel-outer / plane-inner so that every counter load is independent of every other (the
// plane-outer accumulation into px serialised on the LDS latency and spilled px under the
// 64-register cap).  px arrives holding the prefetched (d, my) of each pixel and leaves holding v.
// NSP > 0: number of species known at compile time (planes = (NSP + 1) / 2, f-values constant-bank
//          operands, and the unused upper half of the last plane of an odd count costs nothing);
// NSP == 0: any number of planes, f-values re-read from shared memory.
// EXACT: N == S0 * R0 (power-of-two grid, no Bluestein padding) and NB0 * NT == S0: every pixel index is
// in range, so clamps, range selects and the butterfly-count test disappear and all counter loads
// become one base register plus immediates.
template <int NSP, int NB0, int R0, int S0, int NT, bool EXACT>
__device__ __forceinline__ void flush_finish(float2 (&px)[NB0][R0], const uint32_t *words, int NP,
                                             const float2 *s_table, const float2 (&f)[GX_MAX_SPECIES], int npair,
                                             int tid, int N, float af_re, float af_im, float mzv)
{
#pragma unroll
    for (int i = 0; i < NB0; ++i) {
        const int t = tid + i * NT;
        if (!EXACT && t >= S0) {
#pragma unroll
            for (int n = 0; n < R0; ++n) px[i][n] = make_float2(0.f, 0.f);
            continue;
        }
        // no branch per pixel (a pixel beyond N reads a clamped address and is zeroed by a select):
        // the R0 x planes counter loads are then free to be issued back to back
#pragma unroll
        for (int n = 0; n < R0; ++n) {
            const int y = t + S0 * n;
            const int yy = EXACT ? y : min(y, N - 1);
            float sx = 0.f, sy = 0.f;
            if (NSP > 0) {
#pragma unroll
                for (int w = 0; w < (NSP + 1) / 2; ++w) {
                    // counts -> fp32 through the 2^23 mantissa trick (PRMT + FADD, no I2F)
                    const uint32_t cnt = words[w * NP + yy];
                    const float n0 = __uint_as_float(__byte_perm(cnt, 0x4B000000u, 0x7610)) - 8388608.f;
                    sx = fmaf(n0, f[2 * w].x, sx);
                    sy = fmaf(n0, f[2 * w].y, sy);
                    if (2 * w + 1 < NSP) {
                        const float n1 = __uint_as_float(__byte_perm(cnt, 0x4B000000u, 0x7632)) - 8388608.f;
                        sx = fmaf(n1, f[2 * w + 1].x, sx);
                        sy = fmaf(n1, f[2 * w + 1].y, sy);
                    }
                }
            } else {
                for (int w = 0; w < npair; ++w) {
                    const uint32_t cnt = words[w * NP + yy];
                    const float2 f0 = s_table[2 * w], f1 = s_table[2 * w + 1];
                    const float n0 = __uint_as_float(0x4B000000u | (cnt & 0xffffu)) - 8388608.f;
                    const float n1 = __uint_as_float(__byte_perm(cnt, 0x4B000000u, 0x7632)) - 8388608.f;
                    sx = fmaf(n1, f1.x, fmaf(n0, f0.x, sx));
                    sy = fmaf(n1, f1.y, fmaf(n0, f0.y, sy));
                }
            }
            const float2 dm = px[i][n];
            const float m = (EXACT || y < N) ? mzv * dm.y : 0.f;
            px[i][n] = make_float2(fmaf(dm.x, af_re, sx) * m, fmaf(dm.x, af_im, sy) * m);
        }
    }
}

// Pixel ownership follows the first FFT pass: butterfly t of pass 0 combines the
// pixels t + S0*n (n < R0), so the thread that runs butterfly t also gathers the
// species counts of exactly those pixels, completes them in registers and feeds
// them straight into its radix-R0 butterfly: the finished row is never written
// to shared memory in natural order and never read back by pass 0.
// NSP: number of species fixed at compile time (1..6; the flush is then fully
// unrolled with the f-values as constant-bank operands), 0 = any number (generic loop).
template <int L, bool SPECIES, bool BLUE, int NSP>
__global__ void __launch_bounds__(PROJ_THREADS, (L >= 13) ? 2 : GX_F1_MINBLOCKS)
slice_rows_fused(FusedArgs fa)
{
    typedef GxSched<L> Sc;
    constexpr int M = 1 << L;
    constexpr int NT = PROJ_THREADS;
    constexpr int R0 = Sc::R0, S0 = M / R0;
    constexpr int NB0 = (S0 + NT - 1) / NT;                // pass-0 butterflies per thread
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *buf = reinterpret_cast<float2 *>(smem_raw);
    uint32_t *words = reinterpret_cast<uint32_t *>(smem_raw);   // aliases buf (used strictly before it)
    __shared__ float2 s_table[GX_MAX_SPECIES];
    const ProjArgs &a = fa.proj;
    const int p = blockIdx.x, z = blockIdx.y;
    const int N = BLUE ? a.N : M;                          // power-of-two grids are transformed at their own length
    const int tid = threadIdx.x;
    constexpr bool EXACT = !BLUE && NB0 * NT == S0;        // every pixel index t + S0 n is a pixel of the row
    // every per-rotation / per-row scalar is requested before the first branch, so the CTA
    // pays one L2 round trip for all of them instead of one per dependent use
    int z_min, z_max, jlo, jhi, beg, end;
    double s, c, shift;
    if (fa.use_const) {
        z_min = c_bbox[4 * p + 2]; z_max = c_bbox[4 * p + 3];
        jlo = c_colrange[2 * p]; jhi = c_colrange[2 * p + 1];
        s = c_sn[p]; c = c_cs[p]; shift = c_yrange[2 * p];
        beg = c_row_start[z]; end = c_row_start[z + 1];
    } else {
        z_min = __ldg(a.bbox + 4 * p + 2); z_max = __ldg(a.bbox + 4 * p + 3);
        jlo = __ldg(fa.colrange + 2 * p); jhi = __ldg(fa.colrange + 2 * p + 1);
        s = __ldg(a.sn + p); c = __ldg(a.cs + p); shift = __ldg(a.yrange + 2 * p);
        beg = __ldg(a.row_start + z); end = __ldg(a.row_start + z + 1);
    }
    const float mzv = __ldg(a.mz + (size_t)p * N + z);
    int za, zb;
    if (a.fill_bkg) { za = z_min; zb = z_max - 1; }
    else if (a.sigma > 0) { za = 0; zb = N - 1; }
    else { za = z_min; zb = z_max; }
    if (z < za || z > zb) return;
    if (jhi <= jlo) return;

    const int NP = (N + 3) & ~3;                           // counter plane stride (words)
    const float2 *dmy = fa.dmy + (size_t)p * N;
    float2 px[NB0][R0];
    bool finished = false;                                 // px already holds the completed pixels

    if (SPECIES) {
        const bool single = end - beg <= 65535;            // 16-bit counters cannot wrap: one scatter, one flush
        // the f-value table in shared memory is only read by the generic flush and by chunked rows
        if ((NSP == 0 || !single) && tid < GX_MAX_SPECIES)
            s_table[tid] = tid < a.n_species ? a.table[tid] : make_float2(0.f, 0.f);
        const int npair = (a.n_species + 1) >> 1;
        uint4 *words4 = reinterpret_cast<uint4 *>(smem_raw);
        if constexpr (EXACT && NSP > 0 && (M / 4) % NT == 0) {
            // plane count and length known at compile time: straight-line 16-byte stores, immediate offsets
#pragma unroll
            for (int k = 0; k < ((NSP + 1) / 2) * (M / 4) / NT; ++k) words4[tid + k * NT] = make_uint4(0u, 0u, 0u, 0u);
        } else {
            for (int y = tid; y < npair * (NP / 4); y += NT) words4[y] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        if (single) {
            // (hand-pipelining the atom loads across the zeroing barrier and across batches was
            // measured: no gain, the kernel is bound by instruction issue, not by load latency)
            scatter_species(a, beg, end, s, c, shift, words, NP);
            // (d, my) of this thread's pixels, requested before the barrier so that the L2 round
            // trip overlaps the wait for the slowest warp
#pragma unroll
            for (int i = 0; i < NB0; ++i)
#pragma unroll
                for (int n = 0; n < R0; ++n) {
                    const int t = tid + i * NT, y = t + S0 * n;
                    px[i][n] = __ldg(dmy + (EXACT ? y : min(y, N - 1)));   // clamped: pixels beyond N are zeroed later
                }
            __syncthreads();
            flush_finish<NSP, NB0, R0, S0, NT, EXACT>(px, words, NP, s_table, fa.ftab, npair, tid, N, fa.af_re, fa.af_im, mzv);
            __syncthreads();   // every counter read is done before buf is written
            finished = true;
        } else {
#pragma unroll
            for (int i = 0; i < NB0; ++i)
#pragma unroll
                for (int n = 0; n < R0; ++n) px[i][n] = make_float2(0.f, 0.f);
            // rows with more than 65535 atoms are counted in chunks
            for (int c0 = beg; c0 < end; c0 += 65535) {
                const int c1 = min(c0 + 65535, end);
                scatter_species(a, c0, c1, s, c, shift, words, NP);
                __syncthreads();
                const bool more = c1 < end;
                for (int w = 0; w < npair; ++w) {
                    const float2 f0 = s_table[2 * w], f1 = s_table[2 * w + 1];
                    uint32_t *plane = words + w * NP;
#pragma unroll
                    for (int i = 0; i < NB0; ++i) {
                        const int t = tid + i * NT;
#pragma unroll
                        for (int n = 0; n < R0; ++n) {
                            const int y = t + S0 * n;
                            if (t < S0 && y < N) {
                                const uint32_t cnt = plane[y];
                                const float n0 = __uint_as_float(0x4B000000u | (cnt & 0xffffu)) - 8388608.f;
                                const float n1 = __uint_as_float(__byte_perm(cnt, 0x4B000000u, 0x7632)) - 8388608.f;
                                px[i][n].x = fmaf(n1, f1.x, fmaf(n0, f0.x, px[i][n].x));
                                px[i][n].y = fmaf(n1, f1.y, fmaf(n0, f0.y, px[i][n].y));
                            }
                        }
                    }
                }
                __syncthreads();   // every counter read is done before the next chunk / before buf is written
                if (more) {
                    for (int y = tid; y < npair * (NP / 4); y += NT) words4[y] = make_uint4(0u, 0u, 0u, 0u);
                    __syncthreads();
                }
            }
        }
    } else {
        // generic per-atom f: accumulate straight into the (padded) row buffer
        const double r = a.r, inv_r = 1.0 / a.r;
        for (int y = tid; y < M; y += NT) buf[gx_phys(y)] = make_float2(0.f, 0.f);
        __syncthreads();
        for (int i = beg + tid; i < end; i += NT) {
            const double q = atom_y_pixel(a.xs[i], a.ys[i], s, c, shift, r, inv_r);
            if (q < (double)N) {
                const float2 f = a.f[i];
                float2 *dst = &buf[gx_phys((int)q)];
                atomicAdd(&dst->x, f.x);
                atomicAdd(&dst->y, f.y);
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NB0; ++i)
#pragma unroll
            for (int n = 0; n < R0; ++n) {
                const int t = tid + i * NT, y = t + S0 * n;
                if (t < S0 && y < N) px[i][n] = buf[gx_phys(y)];
            }
        __syncthreads();
    }

    // complete the pixels (pedestal-free) in registers and run the first pass on them:
    // (atoms + d * avg_f) * mz * my with (d, my) packed in one 8-byte load per pixel
    const float2 *tw0 = fa.plan + fa.lay.tw_off[0];
#pragma unroll
    for (int i = 0; i < NB0; ++i) {
        const int t = tid + i * NT;
        if (t < S0) {
            if (!finished) {
                float2 dm[R0];
#pragma unroll
                for (int n = 0; n < R0; ++n) {
                    const int y = t + S0 * n;
                    dm[n] = (y < N) ? __ldg(dmy + y) : make_float2(0.f, 0.f);
                }
#pragma unroll
                for (int n = 0; n < R0; ++n) {
                    const int y = t + S0 * n;
                    float2 v = make_float2(0.f, 0.f);
                    if (y < N) {
                        const float m = mzv * dm[n].y;
                        v = make_float2(fmaf(dm[n].x, fa.af_re, px[i][n].x) * m, fmaf(dm[n].x, fa.af_im, px[i][n].y) * m);
                    }
                    px[i][n] = v;
                }
            }
            if (BLUE) {
#pragma unroll
                for (int n = 0; n < R0; ++n) {
                    const int y = t + S0 * n;
                    if (y < N) px[i][n] = gx_cmul(px[i][n], fa.plan[fa.lay.chirp_off + y]);
                }
            }
            gx_fft_pass0_from_regs<L>(px[i], buf, tw0, t);
        }
    }
    __syncthreads();
    // only the coefficients jlo - N/2 <= k < jhi - N/2 are read below: when they lie within +-512 of DC
    // (always for the 4096 grids of the path: ~ +-290) the last pass computes 2 (a few butterflies: 4)
    // of its 16 outputs per butterfly
    bool full = true;
    if constexpr (L == 12 && !BLUE) {
        const int klo = jlo - M / 2, khi = jhi - M / 2;
        if (klo >= -512 && khi <= 512) {
            gx_fft_pass<16, 16, M, 1, 0, false, GX_PASS1_TWP>(buf, fa.plan + fa.lay.tw_off[1], tid, NT);
            __syncthreads();
            gx_fft_lastpass16_lowband<M>(buf, klo, khi, tid, NT);
            __syncthreads();
            full = false;
        }
    }
    if (full) gx_dft_block<L, 1, 0, BLUE ? 1 : 0, true>(buf, fa.lay, fa.plan, tid, NT);

    // kept q-columns only; shifted column j holds unshifted coefficient j - N/2 (mod N)
    float2 *dst = fa.work + ((size_t)p * N + z) * fa.KC;
    const int half = N / 2;
    for (int jj = tid; jj < jhi - jlo; jj += NT) {
        int k = jlo + jj - half;
        if (k < 0) k += N;
        dst[jj] = gx_dft_result<L, BLUE ? 1 : 0>(buf, fa.lay, fa.plan, k);
    }
}

// ------------------------------------------------------------------ F2 ----
template <int L, int TC, bool BLUE>
__global__ void __launch_bounds__(512)
slice_cols_fused(FusedArgs fa)
{
    constexpr int M = 1 << L;
    constexpr int BS0 = M + (M >> 4) + (M >> 8) + 1;
    // buffer stride == 16/TC (mod 16) float2: the TC interleaved columns of a
    // half-warp land on disjoint banks when the tile is loaded and read out
    constexpr int WANT = (16 / TC) % 16;
    constexpr int BS = BS0 + ((WANT - (BS0 % 16)) + 16) % 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *smem = reinterpret_cast<float2 *>(smem_raw);
    const ProjArgs &a = fa.proj;
#if GX_F2_TILE_FAST
    const int p = blockIdx.y, tile = blockIdx.x;          // neighbouring CTAs read neighbouring columns of one rotation
#else
    const int p = blockIdx.x, tile = blockIdx.y;
#endif
    const int N = a.N, tid = threadIdx.x, nt = blockDim.x;
    const int jlo = fa.colrange[2 * p], jhi = fa.colrange[2 * p + 1];
    const int kc = jhi - jlo;
    const int jj0 = tile * TC;
    if (jj0 >= kc) return;
    int za, zb;
    active_band(a, p, za, zb);

    const float2 *src = fa.work + (size_t)p * N * fa.KC + jj0;
    for (int w = tid; w < TC * M; w += nt) {
        const int cc = w % TC, n = w / TC;
        float2 v = make_float2(0.f, 0.f);
        if (n >= za && n <= zb && jj0 + cc < kc) {
            v = src[(size_t)n * fa.KC + cc];
            if (BLUE) v = gx_cmul(v, fa.plan[fa.lay.chirp_off + n]);
        }
        smem[cc * BS + gx_phys(n)] = v;
    }
    __syncthreads();
    // only the q-rows row_lo - N/2 <= kz < row_hi - N/2 are binned below: band-limited last pass as in F1
    bool full = true;
    if constexpr (L == 12 && !BLUE) {
        const int klo = fa.row_lo - M / 2, khi = fa.row_hi - M / 2;
        if (klo >= -512 && khi <= 512) {
            gx_fft_pass<16, M / 16, M, TC, BS, false>(smem, fa.plan + fa.lay.tw_off[0], tid, nt);
            __syncthreads();
            gx_fft_pass<16, 16, M, TC, BS, false, GX_PASS1_TWP>(smem, fa.plan + fa.lay.tw_off[1], tid, nt);
            __syncthreads();
            gx_fft_lastpass16_lowband<M, TC, BS>(smem, klo, khi, tid, nt);
            __syncthreads();
            full = false;
        }
    }
    if (full) gx_dft_block<L, TC, BS, BLUE ? 1 : 0>(smem, fa.lay, fa.plan, tid, nt);

    const int half = N / 2;
    const int kr = fa.row_hi - fa.row_lo;
    for (int w = tid; w < TC * kr; w += nt) {
        const int cc = w / kr, ii = w - cc * kr;
        if (jj0 + cc >= kc) break;
        const int j = jlo + jj0 + cc;
        const int yx = fa.col[(size_t)p * N + j];
        if (yx < 0) continue;
        const int i = fa.row_lo + ii;
        const int iz = fa.row_index[i];
        if (ii == 0) atomicAdd(&fa.count2[yx], 1u);
        if (iz < 0) continue;
        int kz = i - half;
        if (kz < 0) kz += N;
        float2 v = gx_dft_result<L, BLUE ? 1 : 0>(smem + cc * BS, fa.lay, fa.plan, kz);
        if (kz == 0 && j == half) { v.x += fa.dc_re; v.y += fa.dc_im; }
        atomicAdd(&fa.vsum[(size_t)yx * fa.q_num + iz], v.x * v.x + v.y * v.y);
    }
}

// -------------------------------------------------------------- col range ----
__global__ void __launch_bounds__(256)
col_range_kernel(const int32_t *__restrict__ col, int N, int32_t *range)
{
    __shared__ int s_lo, s_hi;
    const int p = blockIdx.x;
    if (threadIdx.x == 0) { s_lo = N; s_hi = 0; }
    __syncthreads();
    int lo = N, hi = 0;
    for (int j = threadIdx.x; j < N; j += blockDim.x)
        if (col[(size_t)p * N + j] >= 0) { lo = min(lo, j); hi = max(hi, j + 1); }
    atomicMin(&s_lo, lo);
    atomicMax(&s_hi, hi);
    __syncthreads();
    if (threadIdx.x == 0) { range[2 * p] = s_lo; range[2 * p + 1] = s_hi; }
}

extern "C" int gx_slice_col_range(const int32_t *d_col, int n_phi, int N, int32_t *d_range, void *stream)
{
    GX_REQUIRE(d_col && d_range && n_phi > 0 && N > 0, "bad arguments");
    col_range_kernel<<<n_phi, 256, 0, gx_stream(stream)>>>(d_col, N, d_range);
    return gx_check_launch("gx_slice_col_range");
}

// ---------------------------------------------------------------- launch ----
template <int L, int TC>
static int launch_fused(const FusedArgs &fa, bool species, int phases, cudaStream_t st)
{
    constexpr int M = 1 << L;
    constexpr int BS0 = M + (M >> 4) + (M >> 8) + 1;
    constexpr int WANT = (16 / TC) % 16;
    constexpr int BS = BS0 + ((WANT - (BS0 % 16)) + 16) % 16;
    const int N = fa.proj.N;
    const bool blue = fa.lay.bluestein != 0;
    size_t smem1 = (size_t)gx_phys_len(M) * sizeof(float2);
    if (species) {
        size_t w = (size_t)((fa.proj.n_species + 1) / 2) * ((N + 3) & ~3) * sizeof(uint32_t);
        if (w > smem1) smem1 = w;
    }
    const size_t smem2 = (size_t)BS * TC * sizeof(float2);
    if (smem1 > 227 * 1024 || smem2 > 227 * 1024) {
        gx_set_error("gx_slices_fused: shared memory need (%zu / %zu B) exceeds 227 KB", smem1, smem2);
        return GX_ERR_UNSUPPORTED;
    }
    const dim3 grid1(fa.n_phi, N);
#define GX_LAUNCH_ROWS(SP, BL, NPR)                                                                         \
    do {                                                                                                    \
        GX_CUDA(cudaFuncSetAttribute(slice_rows_fused<L, SP, BL, NPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)smem1));                                                          \
        slice_rows_fused<L, SP, BL, NPR><<<grid1, PROJ_THREADS, smem1, st>>>(fa);                          \
    } while (0)
    // compile-time species counts only for the large transforms (L >= 10), where F1 dominates the run
    const int nsp = (species && L >= 10 && fa.proj.n_species >= 1 && fa.proj.n_species <= 6) ? fa.proj.n_species : 0;
#define GX_ROWS_NSP(BL)                                                                  \
    do {                                                                                 \
        switch (nsp) {                                                                   \
        case 1: GX_LAUNCH_ROWS(true, BL, (L >= 10 ? 1 : 0)); break;                      \
        case 2: GX_LAUNCH_ROWS(true, BL, (L >= 10 ? 2 : 0)); break;                      \
        case 3: GX_LAUNCH_ROWS(true, BL, (L >= 10 ? 3 : 0)); break;                      \
        case 4: GX_LAUNCH_ROWS(true, BL, (L >= 10 ? 4 : 0)); break;                      \
        case 5: GX_LAUNCH_ROWS(true, BL, (L >= 10 ? 5 : 0)); break;                      \
        case 6: GX_LAUNCH_ROWS(true, BL, (L >= 10 ? 6 : 0)); break;                      \
        default: GX_LAUNCH_ROWS(true, BL, 0); break;                                     \
        }                                                                                \
    } while (0)
    if (phases & 1) {
        if (species && blue) GX_ROWS_NSP(true);
        else if (species) GX_ROWS_NSP(false);
        else if (blue) GX_LAUNCH_ROWS(false, true, 0);
        else GX_LAUNCH_ROWS(false, false, 0);
        if (int e = gx_check_launch("slice_rows_fused")) return e;
    }
#undef GX_ROWS_NSP
#undef GX_LAUNCH_ROWS
    if (!(phases & 2)) return GX_OK;
    int nt = TC * M / 16;
    nt = nt < 64 ? 64 : (nt > 512 ? 512 : nt);
#if GX_F2_TILE_FAST
    const dim3 grid2((fa.KC + TC - 1) / TC, fa.n_phi);
#else
    const dim3 grid2(fa.n_phi, (fa.KC + TC - 1) / TC);
#endif
    if (blue) {
        GX_CUDA(cudaFuncSetAttribute(slice_cols_fused<L, TC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        slice_cols_fused<L, TC, true><<<grid2, nt, smem2, st>>>(fa);
    } else {
        GX_CUDA(cudaFuncSetAttribute(slice_cols_fused<L, TC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        slice_cols_fused<L, TC, false><<<grid2, nt, smem2, st>>>(fa);
    }
    return gx_check_launch("slice_cols_fused");
}

extern "C" int gx_slices_fused(const gx_fused_args *h, void *stream)
{
    GX_REQUIRE(h != NULL, "NULL argument block");
    GX_REQUIRE(h->d_xs && h->d_ys && h->d_row_start && h->d_sin && h->d_cos && h->d_yrange && h->d_bbox &&
               h->d_dmy && h->d_plan && h->d_col && h->d_colrange && h->d_row_index && h->d_work &&
               h->d_sum && h->d_count2, "NULL pointer");
    GX_REQUIRE(h->n_species >= 0 && h->n_species <= GX_MAX_SPECIES, "n_species out of range");
    GX_REQUIRE(h->n_species == 0 ? h->d_f != NULL : (h->d_species != NULL && h->d_table != NULL),
               "species/f inputs missing");
    GX_REQUIRE(h->d_mz, "mask vector missing");
    GX_REQUIRE(h->n_phi > 0 && h->n_phi <= 65535 && h->N >= 16 && h->KC > 0 && h->q_num > 0, "bad sizes");
    GX_REQUIRE(h->row_lo >= 0 && h->row_hi <= h->N && h->row_lo <= h->row_hi, "bad kept-row range");
    FusedArgs fa;
    ProjArgs &a = fa.proj;
    a.xs = h->d_xs; a.ys = h->d_ys; a.species = h->d_species; a.f = reinterpret_cast<const float2 *>(h->d_f);
    a.row_start = h->d_row_start; a.table = reinterpret_cast<const float2 *>(h->d_table);
    a.n_species = h->n_species; a.sn = h->d_sin; a.cs = h->d_cos; a.yrange = h->d_yrange; a.bbox = h->d_bbox;
    a.base = NULL; a.my = NULL; a.mz = h->d_mz;
    fa.dmy = reinterpret_cast<const float2 *>(h->d_dmy);
    fa.af_re = (float)h->avg_f_re; fa.af_im = (float)h->avg_f_im;
    for (int k = 0; k < GX_MAX_SPECIES; ++k)
        fa.ftab[k] = k < h->n_species ? make_float2(h->table[k].x, h->table[k].y) : make_float2(0.f, 0.f);
    a.N = h->N; a.r = h->r; a.ped_re = (float)h->pedestal_re; a.ped_im = (float)h->pedestal_im;
    a.fill_bkg = h->fill_bkg; a.sigma = h->smooth_sigma;
    fa.lay = gx_fft_layout(h->N);
    if (fa.lay.M == 0) { gx_set_error("gx_slices_fused: unsupported grid size %d", h->N); return GX_ERR_UNSUPPORTED; }
    fa.plan = reinterpret_cast<const float2 *>(h->d_plan);
    fa.col = h->d_col; fa.colrange = h->d_colrange; fa.row_index = h->d_row_index;
    fa.row_lo = h->row_lo; fa.row_hi = h->row_hi;
    fa.work = reinterpret_cast<float2 *>(h->d_work); fa.KC = h->KC; fa.q_num = h->q_num;
    fa.vsum = h->d_sum; fa.count2 = h->d_count2;
    const double n2 = (double)h->N * (double)h->N;
    const bool has_ped = h->fill_bkg || h->smooth_sigma > 0;
    fa.dc_re = has_ped ? (float)(h->pedestal_re * n2) : 0.f;
    fa.dc_im = has_ped ? (float)(h->pedestal_im * n2) : 0.f;
    fa.n_phi = h->n_phi;
    const bool species = h->n_species > 0;
    GX_REQUIRE(h->phases >= 0 && h->phases <= 3, "phases must be 0 (both), 1 (rows), 2 (columns) or 3");
    const int phases = h->phases == 0 ? 3 : h->phases;
    cudaStream_t st = gx_stream(stream);
    fa.use_const = (h->n_phi <= GX_CONST_BATCH && h->N <= GX_CONST_ROWS && !getenv("GIWAXS_B200_NO_CONST")) ? 1 : 0;
    if (fa.use_const && (phases & 1)) {
        const size_t n = (size_t)h->n_phi;
        const cudaMemcpyKind dd = cudaMemcpyDeviceToDevice;
        GX_CUDA(cudaMemcpyToSymbolAsync(c_sn, h->d_sin, n * sizeof(double), 0, dd, st));
        GX_CUDA(cudaMemcpyToSymbolAsync(c_cs, h->d_cos, n * sizeof(double), 0, dd, st));
        GX_CUDA(cudaMemcpyToSymbolAsync(c_yrange, h->d_yrange, 2 * n * sizeof(double), 0, dd, st));
        GX_CUDA(cudaMemcpyToSymbolAsync(c_bbox, h->d_bbox, 4 * n * sizeof(int32_t), 0, dd, st));
        GX_CUDA(cudaMemcpyToSymbolAsync(c_colrange, h->d_colrange, 2 * n * sizeof(int32_t), 0, dd, st));
        GX_CUDA(cudaMemcpyToSymbolAsync(c_row_start, h->d_row_start, ((size_t)h->N + 2) * sizeof(int32_t), 0, dd, st));
    }
    switch (fa.lay.L) {
    case 4: return launch_fused<4, 8>(fa, species, phases, st);
    case 5: return launch_fused<5, 8>(fa, species, phases, st);
    case 6: return launch_fused<6, 8>(fa, species, phases, st);
    case 7: return launch_fused<7, 8>(fa, species, phases, st);
    case 8: return launch_fused<8, 8>(fa, species, phases, st);
    case 9: return launch_fused<9, 8>(fa, species, phases, st);
    case 10: return launch_fused<10, 8>(fa, species, phases, st);
    case 11: return launch_fused<11, 8>(fa, species, phases, st);
    case 12: return launch_fused<12, GX_F2_TC12>(fa, species, phases, st);
    case 13: return launch_fused<13, 2>(fa, species, phases, st);
    }
    gx_set_error("gx_slices_fused: unsupported log2 size %d", fa.lay.L);
    return GX_ERR_UNSUPPORTED;
}
