// Fused stage-A slice pipeline: the production path of voxelgridmaker_fitting.
//
//   F1  slice_rows_fused   one CTA per (rotation, z-row):
//         atoms of the row -> fixed-point integer accumulators in shared memory
//         (ATOMS.ADD) -> complex row (background, edge blend, pedestal removed) ->
//         1-D FFT along y in the same shared memory -> only the q-columns the voxel
//         box keeps are written (N x Kc complex64 instead of N x N).
//   F2  slice_cols_tma     N = 1024, 2048, 4096: persistent CTAs, column tiles streamed in by TMA
//         ((N/16)-row x 4-column boxes through an mbarrier ring), split 16 x N/16
//         transform, only the kept outputs formed and binned at once;
//       slice_cols_fused   other sizes: one CTA per (rotation, TC kept columns):
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
#include <cuda.h>
#include "gx_project.cuh"
#include "gx_fft_engine.cuh"
#include "gx_tma.cuh"

#ifndef GX_F2_TC12
#define GX_F2_TC12 4        // columns per F2 CTA at N = 4096 (2 -> 3 CTAs/SM was measured: 1 % slower)
#endif
#ifndef GX_F2_TILE_FAST
#define GX_F2_TILE_FAST 0   // F2 grid order: 0 = rotation fastest, 1 = column tile fastest
#endif
#ifndef GX_PASS1_TWP
#define GX_PASS1_TWP GX_TWP  // twiddles of the middle pass (16 distinct sets, 1.9 KB): 1 = four loads + products, 0 = fifteen loads
#endif
#ifndef GX_SCATTER_U
#define GX_SCATTER_U 2      // atoms per thread in flight in the row kernel's scatter (1 .. 4 measured: profiles/r04_summary.md)
#endif
#ifndef GX_F1_MINBLOCKS
#define GX_F1_MINBLOCKS 5   // CTAs/SM the 4096-point row kernel is compiled for (others: 4): 48 registers, 4 bytes spilled with
                            // two atoms in flight; 61.7 (4 CTAs, 64 registers) -> 59.8 us per slice, 6 CTAs: 68.3, 3: 66.3
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

// The constant tables belong to one launch at a time.  Launches of one device that come from DIFFERENT
// streams are ordered through an event: the stream that staged the tables last records it after its row
// kernel, and a launch from another stream waits for it before overwriting them (round 1 only documented
// "one stream per device"; two engines on two streams silently corrupted each other).
#include <mutex>
static std::mutex g_const_mutex;
static cudaStream_t g_const_stream[64];
static cudaEvent_t g_const_event[64];
static bool g_const_used[64];

static int const_tables_acquire(cudaStream_t st)
{
    int dev = 0;
    GX_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return GX_OK;
    std::lock_guard<std::mutex> lock(g_const_mutex);
    if (g_const_used[dev] && g_const_stream[dev] != st) GX_CUDA(cudaStreamWaitEvent(st, g_const_event[dev], 0));
    return GX_OK;
}

static int const_tables_release(cudaStream_t st)
{
    int dev = 0;
    GX_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return GX_OK;
    std::lock_guard<std::mutex> lock(g_const_mutex);
    if (!g_const_event[dev]) GX_CUDA(cudaEventCreateWithFlags(&g_const_event[dev], cudaEventDisableTiming));
    GX_CUDA(cudaEventRecord(g_const_event[dev], st));
    g_const_stream[dev] = st;
    g_const_used[dev] = true;
    return GX_OK;
}

struct FusedArgs {
    ProjArgs proj;
    const float2 *dmy;         // [n_phi][N] (num_missing - max_voxels, my): base = dmy.x * avg_f
    float af_re, af_im;
    int2 ffx[GX_MAX_SPECIES];      // species f-values in fixed point for the integer accumulators (see scatter_fixed)
    float fx_scale_re, fx_scale_im;    // 2^k_re, 2^k_im (generic per-atom f path converts on the fly)
    float fx_inv_re, fx_inv_im;        // 2^-k_re, 2^-k_im
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
    double *dc_acc;            // [2 GX_DC_SLOTS] fp64 sums of the DC samples + their voxel keys (NULL: straight into vsum)
    int n_phi;
    int use_const;             // row-kernel scalars are in the constant tables
    int chunk_atoms;           // atoms one pass of the integer accumulators may take (31-bit headroom)
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
// Integer row accumulators.  Each atom adds its scattering factor f = Z + f' + i f'' to its pixel with
// two native shared-memory integer atomics (ATOMS.ADD; fp32 shared atomics are CAS loops on sm_100a
// and would make the sum depend on atom order):
//     plane 0  += round(Re f 2^k_re)          plane 1  += round(Im f 2^k_im)
// k_re, k_im are the largest exponents for which the atoms of one chunk cannot overflow 31 bits even if
// all of them fall into one pixel (gx_slices_fused: from the largest sum of |Re f| over a z row; 2740
// atoms per row at the headline size give k_re = 17, i.e. f resolved to 3.8e-6 absolute, 1e-6 relative).
// Integer sums commute, so the row is independent of atom order (the reference's np.add.at is a
// sequential fp64 sum; its threaded accumulators race).  The per-pixel completion is then
//     v = (w0 2^-k_re + d avg_f.re,  w1 2^-k_im + d avg_f.im) * mz * my
// - two loads and two conversions per pixel whatever the number of species.  (Round 1 counted atoms
// per species in 16-bit fields and formed sum_s n_s f_s per pixel: 27 instructions per pixel with five
// species, 24 % of the kernel's instructions.  A three-plane variant with an exact integer part - Z, f', f''
// separately - was measured first: 12.5 % fewer instructions but 3 ATOMS per atom saturate the shared-memory
// pipe, 62.3 vs 58.0 us per slice; profiles/r04_summary.md.)
template <int U, bool TAIL, bool SPECIES>
__device__ __forceinline__ void scatter_fixed_batch(const ProjArgs &a, const int2 *s_ffx, float sc_re, float sc_im,
                                                    int i0, int end, int nt, double s, double c, double shift,
                                                    double r, double inv_r, int32_t *acc, int NP)
{
    const int N = a.N, last = end - 1;
    double x[U], y[U];
    unsigned sp[U];
    float2 fv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int i = TAIL ? min(i0 + u * nt, last) : i0 + u * nt;
        x[u] = ld_stream_f64(a.xs + i);
        y[u] = ld_stream_f64(a.ys + i);
        if (SPECIES) sp[u] = ld_stream_u8(a.species + i);
        else fv[u] = __ldg(a.f + i);
    }
    int q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = atom_y_pixel_int(x[u], y[u], s, c, shift, r, inv_r);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const bool ok = (!TAIL || i0 + u * nt < end) && ((unsigned)q[u] < (unsigned)N);
        int2 F;
        if (SPECIES) F = s_ffx[sp[u]];
        else F = make_int2(__float2int_rn(fv[u].x * sc_re), __float2int_rn(fv[u].y * sc_im));
        // no branch around an atom: the atomics are predicated (an atom outside the grid, or a species
        // with Im f == 0 such as hydrogen, issues none)
        if (ok) atomicAdd(acc + q[u], F.x);
        if (ok && F.y != 0) atomicAdd(acc + NP + q[u], F.y);
    }
}

// atoms [beg, end) of one z-row: whole batches of 4 x blockDim atoms, then 2, 1 and a per-thread tail
// (no arithmetic for atoms that do not exist)
template <bool SPECIES>
__device__ __forceinline__ void scatter_fixed(const ProjArgs &a, const int2 *s_ffx, float sc_re, float sc_im,
                                              int beg, int end, double s, double c, double shift, int32_t *acc, int NP)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const double r = a.r, inv_r = 1.0 / a.r;
    int i0 = beg + tid;
    int left = end - beg;
    for (; left >= GX_SCATTER_U * nt; left -= GX_SCATTER_U * nt, i0 += GX_SCATTER_U * nt)
        scatter_fixed_batch<GX_SCATTER_U, false, SPECIES>(a, s_ffx, sc_re, sc_im, i0, end, nt, s, c, shift, r, inv_r, acc, NP);
    if (GX_SCATTER_U > 2 && left >= 2 * nt) {
        scatter_fixed_batch<2, false, SPECIES>(a, s_ffx, sc_re, sc_im, i0, end, nt, s, c, shift, r, inv_r, acc, NP);
        left -= 2 * nt; i0 += 2 * nt;
    }
    if (left >= nt) {
        scatter_fixed_batch<1, false, SPECIES>(a, s_ffx, sc_re, sc_im, i0, end, nt, s, c, shift, r, inv_r, acc, NP);
        left -= nt; i0 += nt;
    }
    if (i0 < end) scatter_fixed_batch<1, false, SPECIES>(a, s_ffx, sc_re, sc_im, i0, end, nt, s, c, shift, r, inv_r, acc, NP);
}

// Pixels of this thread's first butterflies from the accumulators.  FINISH: px arrives holding the
// prefetched (d, my) of each pixel and leaves holding the completed value; else the atom sums are added to px
// (rows counted in several chunks).  EXACT: N == S0 * R0 and NB0 * NT == S0 (power-of-two grid, no Bluestein
// padding): every pixel index is in range, all loads are one base register plus immediates.
template <int NB0, int R0, int S0, int NT, bool EXACT, bool FINISH>
__device__ __forceinline__ void flush_fixed(float2 (&px)[NB0][R0], const int32_t *acc, int NP, int tid, int N,
                                            float inv_re, float inv_im, float af_re, float af_im, float mzv)
{
    // FINISH: v = ((w0 + d af_re 2^k_re) m 2^-k_re, (w1 + d af_im 2^k_im) m 2^-k_im) - the scales ride on the
    // background coefficient and on the mask, so a pixel costs two conversions, two FFMA and three FMUL
    const float bg_re = af_re / inv_re, bg_im = af_im / inv_im;       // exact: the scales are powers of two
    const float mz_re = mzv * inv_re, mz_im = mzv * inv_im;
#pragma unroll
    for (int i = 0; i < NB0; ++i) {
        const int t = tid + i * NT;
        if (!EXACT && t >= S0) {
            if (FINISH) {
#pragma unroll
                for (int n = 0; n < R0; ++n) px[i][n] = make_float2(0.f, 0.f);
            }
            continue;
        }
#pragma unroll
        for (int n = 0; n < R0; ++n) {
            const int y = t + S0 * n;
            const int yy = EXACT ? y : min(y, N - 1);
            const float w0 = (float)acc[yy], w1 = (float)acc[NP + yy];
            if (FINISH) {
                const float2 dm = px[i][n];
                const float my = (EXACT || y < N) ? dm.y : 0.f;
                px[i][n] = make_float2(fmaf(dm.x, bg_re, w0) * (mz_re * my), fmaf(dm.x, bg_im, w1) * (mz_im * my));
            } else if (EXACT || y < N) {
                px[i][n].x = fmaf(w0, inv_re, px[i][n].x);
                px[i][n].y = fmaf(w1, inv_im, px[i][n].y);
            }
        }
    }
}

// Pixel ownership follows the first FFT pass: butterfly t of pass 0 combines the
// pixels t + S0*n (n < R0), so the thread that runs butterfly t also gathers the
// accumulators of exactly those pixels, completes them in registers and feeds
// them straight into its radix-R0 butterfly: the finished row is never written
// to shared memory in natural order and never read back by pass 0.
// SPECIES: atoms carry a species code (f from the 16-entry table) / a per-atom complex64 f.
// ROWPERM: row z is written to slot 256 (z mod 16) + z / 16 of the work buffer, the order the
// TMA-fed column kernel consumes (16 chunks of 256 rows, each a 256-point sub-transform).
template <int L, bool SPECIES, bool BLUE, bool ROWPERM>
__global__ void __launch_bounds__(PROJ_THREADS, (L >= 14) ? 1 : (L >= 13) ? 2 : ((L == 12 && !BLUE) ? GX_F1_MINBLOCKS : 4))
slice_rows_fused(FusedArgs fa)
{
    typedef GxSched<L> Sc;
    constexpr int M = 1 << L;
    constexpr int NT = PROJ_THREADS;
    constexpr int R0 = Sc::R0, S0 = M / R0;
    constexpr int NB0 = (S0 + NT - 1) / NT;                // pass-0 butterflies per thread
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *buf = reinterpret_cast<float2 *>(smem_raw);
    int32_t *acc = reinterpret_cast<int32_t *>(smem_raw);     // aliases buf (used strictly before it)
    __shared__ int2 s_ffx[GX_MAX_SPECIES];
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
    jhi = min(jhi, jlo + fa.KC);                           // the work buffer holds KC columns per row
    if (jhi <= jlo) return;

    const int NP = (N + 3) & ~3;                           // accumulator plane stride (words)
    const float2 *dmy = fa.dmy + (size_t)p * N;
    float2 px[NB0][R0];
    bool finished = false;                                 // px already holds the completed pixels

    if (SPECIES && tid < GX_MAX_SPECIES) s_ffx[tid] = fa.ffx[tid];
    int4 *acc4 = reinterpret_cast<int4 *>(smem_raw);
    if constexpr (EXACT && (2 * M / 4) % NT == 0) {
        // plane length known at compile time: straight-line 16-byte stores, immediate offsets
#pragma unroll
        for (int k = 0; k < (2 * M / 4) / NT; ++k) acc4[tid + k * NT] = make_int4(0, 0, 0, 0);
    } else {
        for (int y = tid; y < 2 * (NP / 4); y += NT) acc4[y] = make_int4(0, 0, 0, 0);
    }
    __syncthreads();
    if (end - beg <= fa.chunk_atoms) {
        // one scatter, one flush (every row of a non-crystalline slab)
        scatter_fixed<SPECIES>(a, s_ffx, fa.fx_scale_re, fa.fx_scale_im, beg, end, s, c, shift, acc, NP);
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
        flush_fixed<NB0, R0, S0, NT, EXACT, true>(px, acc, NP, tid, N, fa.fx_inv_re, fa.fx_inv_im, fa.af_re, fa.af_im, mzv);
        __syncthreads();   // every accumulator read is done before buf is written
        finished = true;
    } else {
        // rows with more atoms than one chunk may hold without overflow are summed chunk by chunk in fp32
#pragma unroll
        for (int i = 0; i < NB0; ++i)
#pragma unroll
            for (int n = 0; n < R0; ++n) px[i][n] = make_float2(0.f, 0.f);
        for (int c0 = beg; c0 < end; c0 += fa.chunk_atoms) {
            const int c1 = min(c0 + fa.chunk_atoms, end);
            scatter_fixed<SPECIES>(a, s_ffx, fa.fx_scale_re, fa.fx_scale_im, c0, c1, s, c, shift, acc, NP);
            __syncthreads();
            flush_fixed<NB0, R0, S0, NT, EXACT, false>(px, acc, NP, tid, N, fa.fx_inv_re, fa.fx_inv_im, 0.f, 0.f, 0.f);
            __syncthreads();   // every accumulator read is done before the next chunk / before buf is written
            if (c1 < end) {
                for (int y = tid; y < 2 * (NP / 4); y += NT) acc4[y] = make_int4(0, 0, 0, 0);
                __syncthreads();
            }
        }
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
    int slot = z;
    if constexpr (ROWPERM) slot = GxSplit<L>::slot(z);
    float2 *dst = fa.work + ((size_t)p * N + slot) * fa.KC;
    const int half = N / 2;
    for (int jj = tid; jj < jhi - jlo; jj += NT) {
        int k = jlo + jj - half;
        if (k < 0) k += N;
        dst[jj] = gx_dft_result<L, BLUE ? 1 : 0>(buf, fa.lay, fa.plan, k);
    }
}

// ------------------------------------------------------------------ F2 ----
// The DC sample of a slice (the coefficient k = 0 of column N/2: |sum of all f + pedestal N^2|^2, the same
// huge number for every rotation, ~1e10 x its neighbours) is accumulated in fp64 on the side and folded into
// the voxel grid once per run (gx_fold_dc): 1800 fp32 additions of it random-walked to 1.4e-6 of the maximum
// and made the result depend on how the rotations were split over ranks.  Column N/2 of the symmetric
// linspace axis sits half a step off q = 0 on a side that depends on the rotation, so the samples fall into
// up to four neighbouring voxels: the side table has GX_DC_SLOTS (key = voxel + 1, fp64 sum) entries,
// claimed with a compare-and-swap; a full table falls back to the fp32 grid.
#define GX_DC_SLOTS 8
__device__ __forceinline__ void add_dc_sample(const FusedArgs &fa, size_t voxel, float2 v)
{
    const double val = (double)v.x * (double)v.x + (double)v.y * (double)v.y;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(fa.dc_acc + GX_DC_SLOTS);
    const unsigned long long key = (unsigned long long)voxel + 1ull;
    for (int s = 0; s < GX_DC_SLOTS; ++s) {
        const unsigned long long prev = atomicCAS(keys + s, 0ull, key);
        if (prev == 0ull || prev == key) { atomicAdd(fa.dc_acc + s, val); return; }
    }
    atomicAdd(fa.vsum + voxel, (float)val);
}

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
        if (kz == 0 && j == half) {
            v.x += fa.dc_re; v.y += fa.dc_im;
            if (fa.dc_acc) { add_dc_sample(fa, (size_t)yx * fa.q_num + iz, v); continue; }
        }
        atomicAdd(&fa.vsum[(size_t)yx * fa.q_num + iz], v.x * v.x + v.y * v.y);
    }
}

// ------------------------------------------- F2, TMA-fed (N = 1024, 2048, 4096) ----
// Same result as slice_cols_fused for the power-of-two grids with a 16-16-R2 schedule, restructured around
// asynchronous tile movement: the [n_phi N, KC] complex64 work buffer is a 2-D tensor map and a dedicated
// producer thread streams (N/16)-row x 4-column boxes (cp.async.bulk.tensor.2d -> UTMALDG, completion on
// mbarriers) through an 8-slot ring while 16 consumer warps transform.  Persistent CTAs (one per SM) walk the
// (rotation, column tile) list, so the boxes of the next tile are already landing while the current tile finishes.
//
// The column transform is split so that two of its three passes need only ONE box (gx_fft_engine.cuh, GxSplit):
//   z = 16 u + c  (F1 writes row z to slot RB c + u, RB = N / 16: ROWPERM)
//   Y_c[k'] = sum_u x[16 u + c] W_RB^{u k'}                        RB-point DIF inside box c (passes alpha, beta)
//   X[k' + RB m] = sum_c W_N^{c k'} Y_c[k'] W_16^{c m}             radix-16 DIT across the boxes   (pass gamma)
// alpha reads the dense box ([row][4 columns], 32 B per row: lanes = 4 rows x 4 columns of a half-warp are 128
// contiguous bytes) and writes the padded per-column layout, beta is in place, gamma reads 16 values one box
// apart, forms only the outputs the voxel window keeps (m = 0, 15, sometimes 1, 14: gx_dft16_lowband) and adds
// |X|^2 straight into the voxel sum: the transformed column is never stored, not even in shared memory.
#define F2T_TC 4
#define F2T_NBOX 16
#define F2T_RING 8
#define F2T_CONSUMERS 512
#define F2T_THREADS (F2T_CONSUMERS + 32)

template <int L> struct F2T {
    enum {
        N = 1 << L, RB = GxSplit<L>::RB, R2 = GxSplit<L>::R2,
        BOX_BYTES = RB * F2T_TC * 8,
        BS0 = N + (N >> 4) + (N >> 8) + 1,
        BS = BS0 + ((4 - (BS0 % 16)) + 16) % 16,       // float2 per column buffer: == 4 (mod 16), see alpha's stores
        SMEM = F2T_RING * BOX_BYTES + F2T_TC * BS * 8 + 2 * F2T_RING * 8
    };
};

template <int L>
__global__ void __launch_bounds__(F2T_THREADS, 1)
slice_cols_tma(FusedArgs fa, const __grid_constant__ CUtensorMap tmap)
{
    typedef F2T<L> G;
    constexpr int N = G::N, RB = G::RB, R2 = G::R2, BS = G::BS, BOX_BYTES = G::BOX_BYTES;
    extern __shared__ __align__(128) unsigned char smem_tma[];
    float2 *ring = reinterpret_cast<float2 *>(smem_tma);
    float2 *work = reinterpret_cast<float2 *>(smem_tma + F2T_RING * BOX_BYTES);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_tma + F2T_RING * BOX_BYTES + F2T_TC * BS * 8);
    uint64_t *empty = full + F2T_RING;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < F2T_RING; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int tiles = fa.KC / F2T_TC;
    const int total = fa.n_phi * tiles;

    if (tid >= F2T_CONSUMERS) {
        // ---- producer: one thread keeps the ring full, across tile boundaries
        if (tid == F2T_CONSUMERS) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            for (int f = blockIdx.x; f < total; f += gridDim.x) {
                const int p = f % fa.n_phi, tile = f / fa.n_phi;
                const int kc = min(fa.colrange[2 * p + 1] - fa.colrange[2 * p], fa.KC);
                if (tile * F2T_TC >= kc) continue;
                for (int c = 0; c < F2T_NBOX; ++c) {
                    const int slot = c & (F2T_RING - 1);
                    mbar_wait(empty + slot, ((c >> 3) & 1) ^ 1);      // n-th refill of a slot: n = 2 tile_iter + c / 8
                    mbar_expect_tx(full + slot, BOX_BYTES);
                    tma_load_2d(reinterpret_cast<unsigned char *>(ring) + slot * BOX_BYTES, &tmap, tile * F2T_TC,
                                p * N + c * RB, full + slot);
                }
            }
        }
        return;
    }

    // ---- consumers
    const int warp = tid >> 5, lane = tid & 31;
    const int pair = warp >> 1;                       // boxes pair and pair + 8 of every tile; ring slot = pair
    const int e = (warp & 1) * 32 + lane;             // 0..63 inside the pair
    const float2 *tw0 = fa.plan + fa.lay.tw_off[0], *tw1 = fa.plan + fa.lay.tw_off[1];
    const int half = N / 2;
    const int klo = fa.row_lo - half, khi = fa.row_hi - half;
    const bool lowband = klo >= -2 * RB && khi <= 2 * RB;
    for (int f = blockIdx.x; f < total; f += gridDim.x) {
        const int p = f % fa.n_phi, tile = f / fa.n_phi;
        const int jlo = fa.colrange[2 * p];
        const int kc = min(fa.colrange[2 * p + 1] - jlo, fa.KC);
        const int jj0 = tile * F2T_TC;
        if (jj0 >= kc) continue;
#pragma unroll 1
        for (int r = 0; r < 2; ++r) {
            const int c = pair + 8 * r;
            // alpha: radix-16 DIF, stride R2, of the RB-point sub-transform; dense box -> padded column buffers
            mbar_wait(full + pair, r);
            if (e < R2 * F2T_TC) {
                const int col = e & 3, t = e >> 2;
                const float2 *st = ring + pair * (RB * F2T_TC) + t * F2T_TC + col;
                float2 v[16];
#pragma unroll
                for (int n = 0; n < 16; ++n) v[n] = st[n * R2 * F2T_TC];
                gx_split_alpha<L, GX_PASS1_TWP>(v, 1, work + col * BS + gx_phys(c * RB + t), tw1, t);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + pair);              // this warp is done with the box
            named_barrier(1 + pair, 64);
            // beta: radix-R2, stride 1, in place
            {
                const int blk = e & 15, col = e >> 4;
                gx_split_beta<L>(work + col * BS + gx_phys(c * RB + R2 * blk));
            }
        }
        named_barrier(9, F2T_CONSUMERS);
        // gamma: radix-16 DIT across the boxes, only the kept outputs, binned at once
#pragma unroll 1
        for (int g = tid; g < F2T_TC * RB; g += F2T_CONSUMERS) {
            const int col = g / RB, kp = g - col * RB;
            if (jj0 + col >= kc) continue;
            const int j = jlo + jj0 + col;
            const int yx = fa.col[(size_t)p * N + j];
            if (yx < 0) continue;
            if (kp == 0) atomicAdd(&fa.count2[yx], 1u);
            float2 v[16];
            gx_split_gamma_inputs<L, GX_TWP>(work + col * BS, tw0, kp, v);
            float *dst = fa.vsum + (size_t)yx * fa.q_num;
            if (lowband) {
                const bool w1 = kp + RB < khi, w14 = kp - 2 * RB >= klo;
                float2 x0, x15, x1, x14;
                gx_dft16_lowband_vals(v, w1 || w14, x0, x15, x1, x14);
                int iz;
                if (kp < khi && (iz = fa.row_index[kp + half]) >= 0) {
                    if (kp == 0 && j == half) {
                        x0.x += fa.dc_re; x0.y += fa.dc_im;
                        if (fa.dc_acc) add_dc_sample(fa, (size_t)yx * fa.q_num + iz, x0);
                        else atomicAdd(dst + iz, x0.x * x0.x + x0.y * x0.y);
                    } else {
                        atomicAdd(dst + iz, x0.x * x0.x + x0.y * x0.y);
                    }
                }
                if (kp - RB >= klo && (iz = fa.row_index[kp - RB + half]) >= 0) atomicAdd(dst + iz, x15.x * x15.x + x15.y * x15.y);
                if (w1 && (iz = fa.row_index[kp + RB + half]) >= 0) atomicAdd(dst + iz, x1.x * x1.x + x1.y * x1.y);
                if (w14 && (iz = fa.row_index[kp - 2 * RB + half]) >= 0) atomicAdd(dst + iz, x14.x * x14.x + x14.y * x14.y);
            } else {
                GxDft<16>::run(v);
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int kz = kp + RB * m;                        // unshifted coefficient index
                    int i = kz + half;
                    if (i >= N) i -= N;
                    if (i < fa.row_lo || i >= fa.row_hi) continue;
                    const int iz = fa.row_index[i];
                    if (iz < 0) continue;
                    float2 x = v[m];
                    if (kz == 0 && j == half) {
                        x.x += fa.dc_re; x.y += fa.dc_im;
                        if (fa.dc_acc) { add_dc_sample(fa, (size_t)yx * fa.q_num + iz, x); continue; }
                    }
                    atomicAdd(dst + iz, x.x * x.x + x.y * x.y);
                }
            }
        }
        named_barrier(9, F2T_CONSUMERS);      // every read of the column buffers is done before the next tile's alpha
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
// 2-D tensor map of the [rows, KC] complex64 work buffer, box = 256 rows x 4 columns (one ring slot)
static int work_tensor_map(CUtensorMap *map, void *work, size_t rows, int KC, int box_rows)
{
    gx_encode_tiled_fn encode = gx_tensor_map_encoder();
    if (!encode) return GX_ERR_UNSUPPORTED;
    const cuuint64_t dims[2] = {(cuuint64_t)KC, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)KC * 8};
    const cuuint32_t box[2] = {F2T_TC, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, work, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        gx_set_error("gx_slices_fused: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return GX_ERR_CUDA;
    }
    return GX_OK;
}

// The TMA-fed column kernel exists for the power-of-two grids with a 16-16-R2 schedule (N = 1024, 2048, 4096) and
// is the default for 2048 and 4096: measured 18.3 vs 19.9 and 60.9 vs 65.2 us per slice against the LDG-fed
// kernel (whole fused path, scripts/time_fused.py), but 10.2 vs 9.4 at N = 1024, where a tile is too small to fill
// a persistent 544-thread CTA (GIWAXS_B200_TMA_MIN_L=10 selects it there too; GIWAXS_B200_NO_TMA=1 disables it).
static bool cols_tma_ok(const FusedArgs &fa, int L, bool blue)
{
    const char *e = getenv("GIWAXS_B200_TMA_MIN_L");
    int min_l = e ? atoi(e) : 11;
    if (min_l < 10) min_l = 10;
    return L >= min_l && L <= 12 && !blue && fa.KC % 8 == 0 && (reinterpret_cast<uintptr_t>(fa.work) & 15) == 0 &&
           !getenv("GIWAXS_B200_NO_TMA");
}

template <int L, int TC>
static int launch_fused(const FusedArgs &fa, bool species, int phases, cudaStream_t st)
{
    constexpr int M = 1 << L;
    constexpr int BS0 = M + (M >> 4) + (M >> 8) + 1;
    constexpr int WANT = (16 / TC) % 16;
    constexpr int BS = BS0 + ((WANT - (BS0 % 16)) + 16) % 16;
    const int N = fa.proj.N;
    const bool blue = fa.lay.bluestein != 0;
    const bool tma = cols_tma_ok(fa, L, blue);
    size_t smem1 = (size_t)gx_phys_len(M) * sizeof(float2);
    const size_t acc_bytes = (size_t)2 * ((N + 3) & ~3) * sizeof(int32_t);
    if (acc_bytes > smem1) smem1 = acc_bytes;
    const size_t smem2 = (size_t)BS * TC * sizeof(float2);
    if (smem1 > 227 * 1024 || smem2 > 227 * 1024) {
        gx_set_error("gx_slices_fused: shared memory need (%zu / %zu B) exceeds 227 KB", smem1, smem2);
        return GX_ERR_UNSUPPORTED;
    }
    const dim3 grid1(fa.n_phi, N);
#define GX_LAUNCH_ROWS(SP, BL, RP)                                                                          \
    do {                                                                                                    \
        GX_CUDA(cudaFuncSetAttribute(slice_rows_fused<L, SP, BL, RP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)smem1));                                                          \
        slice_rows_fused<L, SP, BL, RP><<<grid1, PROJ_THREADS, smem1, st>>>(fa);                           \
    } while (0)
    if (phases & 1) {
        if constexpr (L >= 10 && L <= 12) {
            if (tma) { if (species) GX_LAUNCH_ROWS(true, false, true); else GX_LAUNCH_ROWS(false, false, true); }
        }
        if (!tma) {
            if (species && blue) GX_LAUNCH_ROWS(true, true, false);
            else if (species) GX_LAUNCH_ROWS(true, false, false);
            else if (blue) GX_LAUNCH_ROWS(false, true, false);
            else GX_LAUNCH_ROWS(false, false, false);
        }
        if (int e = gx_check_launch("slice_rows_fused")) return e;
    }
#undef GX_LAUNCH_ROWS
    if (!(phases & 2)) return GX_OK;
    if constexpr (L >= 10 && L <= 12) if (tma) {
        CUtensorMap map;
        if (int e = work_tensor_map(&map, fa.work, (size_t)fa.n_phi * N, fa.KC, F2T<L>::RB)) return e;
        static int sms = 0;
        if (!sms) {
            int dev = 0;
            GX_CUDA(cudaGetDevice(&dev));
            GX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        }
        const int total = fa.n_phi * (fa.KC / F2T_TC);
        GX_CUDA(cudaFuncSetAttribute(slice_cols_tma<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2T<L>::SMEM));
        slice_cols_tma<L><<<total < sms ? total : sms, F2T_THREADS, F2T<L>::SMEM, st>>>(fa, map);
        return gx_check_launch("slice_cols_tma");
    }
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

// Fixed-point plan of the integer accumulators (see scatter_fixed): exponents as large as 31 bits allow.
// max_row_abs_re / max_row_abs_im: the largest sum of |Re f| / |Im f| over the atoms of one z row (a bound on
// what a single pixel can receive).  A row whose bound would push an exponent below 14 bits is summed in
// chunks of atoms instead.
static void fixed_point_plan(const gx_fused_args *h, FusedArgs &fa)
{
    double f_re = 1e-30, f_im = 1e-30;                          // largest |Re f|, |Im f| of one atom
    for (int k = 0; k < h->n_species; ++k) {
        f_re = fmax(f_re, fabs(h->table_f64[2 * k]));
        f_im = fmax(f_im, fabs(h->table_f64[2 * k + 1]));
    }
    if (h->n_species == 0) { f_re = h->max_abs_f_re > 0.0 ? h->max_abs_f_re : 128.0; f_im = h->max_abs_f_im > 0.0 ? h->max_abs_f_im : 128.0; }
    const double rows = h->max_row_atoms > 0 ? (double)h->max_row_atoms : 65535.0;
    // bound on one pixel's sums: the caller's per-row sums if given, else atoms x largest f
    double b_re = h->max_row_abs_re > 0.0 ? h->max_row_abs_re : rows * f_re;
    double b_im = h->max_row_abs_im > 0.0 ? h->max_row_abs_im : rows * f_im;
    const double lim = 2147483647.0 * 0.999;                    // (rounding of each addend: at most 0.5 each)
    int chunk = 0x7fffffff;
    const double floor_scale = 16384.0;                         // never coarser than 2^-14
    if (b_re * floor_scale + rows > lim || b_im * floor_scale + rows > lim) {
        // too populous a row for one pass: chunk by atoms
        double c = fmin(lim / (floor_scale * f_re + 1.0), lim / (floor_scale * f_im + 1.0));
        chunk = c < 256.0 ? 256 : (c > 1e9 ? 1000000000 : (int)c);
        b_re = chunk * f_re; b_im = chunk * f_im;
    }
    const double n_add = fmin(rows, (double)chunk);
    int k_re = (int)floor(log2((lim - n_add) / b_re)), k_im = (int)floor(log2((lim - n_add) / b_im));
    if (k_re > 24) k_re = 24;
    if (k_im > 24) k_im = 24;
    fa.chunk_atoms = chunk;
    fa.fx_scale_re = (float)ldexp(1.0, k_re); fa.fx_scale_im = (float)ldexp(1.0, k_im);
    fa.fx_inv_re = (float)ldexp(1.0, -k_re); fa.fx_inv_im = (float)ldexp(1.0, -k_im);
    for (int k = 0; k < GX_MAX_SPECIES; ++k) {
        int2 v = make_int2(0, 0);
        if (k < h->n_species)
            v = make_int2((int)nearbyint(ldexp(h->table_f64[2 * k], k_re)), (int)nearbyint(ldexp(h->table_f64[2 * k + 1], k_im)));
        fa.ffx[k] = v;
    }
}

extern "C" int gx_slices_fused(const gx_fused_args *h, void *stream)
{
    GX_REQUIRE(h != NULL, "NULL argument block");
    GX_REQUIRE(h->d_xs && h->d_ys && h->d_row_start && h->d_sin && h->d_cos && h->d_yrange && h->d_bbox &&
               h->d_dmy && h->d_plan && h->d_col && h->d_colrange && h->d_row_index && h->d_work &&
               h->d_sum && h->d_count2, "NULL pointer");
    GX_REQUIRE(h->n_species >= 0 && h->n_species <= GX_MAX_SPECIES, "n_species out of range");
    GX_REQUIRE(h->n_species == 0 ? h->d_f != NULL : h->d_species != NULL, "species/f inputs missing");
    GX_REQUIRE(h->d_mz, "mask vector missing");
    GX_REQUIRE(h->n_phi > 0 && h->n_phi <= 65535 && h->N >= 16 && h->KC > 0 && h->q_num > 0, "bad sizes");
    GX_REQUIRE(h->row_lo >= 0 && h->row_hi <= h->N && h->row_lo <= h->row_hi, "bad kept-row range");
    for (int k = 0; k < h->n_species; ++k)
        GX_REQUIRE(fabs(h->table_f64[2 * k]) < 1e6 && fabs(h->table_f64[2 * k + 1]) < 1e6, "scattering factor out of range");
    FusedArgs fa;
    ProjArgs &a = fa.proj;
    a.xs = h->d_xs; a.ys = h->d_ys; a.species = h->d_species; a.f = reinterpret_cast<const float2 *>(h->d_f);
    a.row_start = h->d_row_start; a.table = reinterpret_cast<const float2 *>(h->d_table);
    a.n_species = h->n_species; a.sn = h->d_sin; a.cs = h->d_cos; a.yrange = h->d_yrange; a.bbox = h->d_bbox;
    a.base = NULL; a.my = NULL; a.mz = h->d_mz;
    fa.dmy = reinterpret_cast<const float2 *>(h->d_dmy);
    fa.af_re = (float)h->avg_f_re; fa.af_im = (float)h->avg_f_im;
    fixed_point_plan(h, fa);
    a.N = h->N; a.r = h->r; a.ped_re = (float)h->pedestal_re; a.ped_im = (float)h->pedestal_im;
    a.fill_bkg = h->fill_bkg; a.sigma = h->smooth_sigma;
    fa.lay = gx_fft_layout(h->N);
    if (fa.lay.M == 0) { gx_set_error("gx_slices_fused: unsupported grid size %d", h->N); return GX_ERR_UNSUPPORTED; }
    fa.plan = reinterpret_cast<const float2 *>(h->d_plan);
    fa.col = h->d_col; fa.colrange = h->d_colrange; fa.row_index = h->d_row_index;
    fa.row_lo = h->row_lo; fa.row_hi = h->row_hi;
    fa.work = reinterpret_cast<float2 *>(h->d_work); fa.KC = h->KC; fa.q_num = h->q_num;
    fa.vsum = h->d_sum; fa.count2 = h->d_count2; fa.dc_acc = h->d_dc;
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
        if (int e = const_tables_acquire(st)) return e;
        const size_t n = (size_t)h->n_phi;
        const cudaMemcpyKind dd = cudaMemcpyDeviceToDevice;
        GX_CUDA(cudaMemcpyToSymbolAsync(c_sn, h->d_sin, n * sizeof(double), 0, dd, st));
        GX_CUDA(cudaMemcpyToSymbolAsync(c_cs, h->d_cos, n * sizeof(double), 0, dd, st));
        GX_CUDA(cudaMemcpyToSymbolAsync(c_yrange, h->d_yrange, 2 * n * sizeof(double), 0, dd, st));
        GX_CUDA(cudaMemcpyToSymbolAsync(c_bbox, h->d_bbox, 4 * n * sizeof(int32_t), 0, dd, st));
        GX_CUDA(cudaMemcpyToSymbolAsync(c_colrange, h->d_colrange, 2 * n * sizeof(int32_t), 0, dd, st));
        GX_CUDA(cudaMemcpyToSymbolAsync(c_row_start, h->d_row_start, ((size_t)h->N + 2) * sizeof(int32_t), 0, dd, st));
    }
    int rc;
    switch (fa.lay.L) {
    case 4: rc = launch_fused<4, 8>(fa, species, phases, st); break;
    case 5: rc = launch_fused<5, 8>(fa, species, phases, st); break;
    case 6: rc = launch_fused<6, 8>(fa, species, phases, st); break;
    case 7: rc = launch_fused<7, 8>(fa, species, phases, st); break;
    case 8: rc = launch_fused<8, 8>(fa, species, phases, st); break;
    case 9: rc = launch_fused<9, 8>(fa, species, phases, st); break;
    case 10: rc = launch_fused<10, 8>(fa, species, phases, st); break;
    case 11: rc = launch_fused<11, 8>(fa, species, phases, st); break;
    case 12: rc = launch_fused<12, GX_F2_TC12>(fa, species, phases, st); break;
    case 13: rc = launch_fused<13, 2>(fa, species, phases, st); break;
    case 14: rc = launch_fused<14, 1>(fa, species, phases, st); break;
    default:
        gx_set_error("gx_slices_fused: unsupported log2 size %d", fa.lay.L);
        rc = GX_ERR_UNSUPPORTED;
    }
    if (fa.use_const && (phases & 1) && rc == GX_OK) rc = const_tables_release(st);
    return rc;
}

__global__ void fold_dc_kernel(double *dc, float *vsum)
{
    const int s = threadIdx.x;
    if (blockIdx.x == 0 && s < GX_DC_SLOTS) {
        unsigned long long *keys = reinterpret_cast<unsigned long long *>(dc + GX_DC_SLOTS);
        const unsigned long long key = keys[s];
        if (key != 0ull) {
            atomicAdd(vsum + (size_t)(key - 1ull), (float)dc[s]);
            dc[s] = 0.0;
            keys[s] = 0ull;
        }
    }
}

// d_sum[voxel] += the fp64 sums of the DC samples gathered in d_dc (16 x 8 bytes, zero-initialised) since the
// last fold; d_dc is reset.
extern "C" int gx_fold_dc(double *d_dc, float *d_sum, void *stream)
{
    GX_REQUIRE(d_dc && d_sum, "NULL pointer");
    fold_dc_kernel<<<1, 32, 0, gx_stream(stream)>>>(d_dc, d_sum);
    return gx_check_launch("gx_fold_dc");
}

// 1 when gx_slices_fused will consume the work buffer in the permuted row order of the TMA-fed column
// kernel for this grid size / column count (the caller must then hand it a buffer whose never-written
// rows are zero: the TMA boxes cover every row slot, also those of rows outside the atom band).
extern "C" int gx_fused_wants_zeroed_work(int N, int KC)
{
    FusedArgs fa;
    fa.KC = KC; fa.work = nullptr;
    const GxFftLayout lay = gx_fft_layout(N);
    return (lay.M != 0 && cols_tma_ok(fa, lay.L, lay.bluestein != 0)) ? 1 : 0;
}
