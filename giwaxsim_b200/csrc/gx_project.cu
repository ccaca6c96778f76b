// K1 / K1b: projection of the rotated atoms onto the y-z pixel grid, amorphous
// background fill and Gaussian edge blend (voxelgrids.py:338-379).
//
// One CTA owns one (rotation, z-row).  Atoms were sorted by z-row once
// (gx_atoms_sort_rows), so the CTA streams only its own row's atoms, computes
// their y pixel in fp64 (bit-exact with NumPy), and counts them per species in
// shared memory with native u32 ATOMS.ADD: two 16-bit counters per word, flushed
// into a complex64 row accumulator every <= 65535 atoms so no counter can wrap.
// (sm_100a has no native shared-memory fp32 atomic add -- it compiles to a CAS
// loop -- so integer species counts are both faster and order-independent.)
// The row is then completed in place (background, outside-box overwrite, edge
// blend) and written once, coalesced: the grid is never zero-filled, never hit
// by a global atomic and never read back.
#include "gx_project.cuh"

// --------------------------------------------------------- slice vectors ----
__device__ __forceinline__ double chord_length(const gx_chord &k, double x)
{
    // rectangular_collapse_lengths, per abscissa (voxelgrids.py:260-307); all
    // products/sums individually rounded like NumPy scalar arithmetic
    if (k.mode == 0) return (x < k.hor) ? k.ver : 0.0;
    if (k.mode == 1) return (x < k.ver) ? k.hor : 0.0;
    if (x == 0.0) return 0.0;
    if (x < k.stop1) {
        double tot = __dadd_rn(k.rise, __dmul_rn(x, k.tan_phi));
        double miss = __dmul_rn(__dsub_rn(k.vcos, x), k.tan_theta);
        return __dsub_rn(tot, miss);
    }
    if (x <= k.stop2) return k.mid;
    if (x < k.stop12) {
        double rem = __dsub_rn(k.hor, __ddiv_rn(__dsub_rn(x, k.vcos), k.cos_phi));
        return __ddiv_rn(rem, k.cos_theta);
    }
    return 0.0;
}

// Python slice [a:b] on a length-n axis -> half-open index range
__device__ __forceinline__ void py_slice(int a, int b, int n, int &lo, int &hi)
{
    if (a < 0) a = max(a + n, 0); else a = min(a, n);
    if (b < 0) b = max(b + n, 0); else b = min(b, n);
    lo = a; hi = max(a, b);
}

// gaussian_filter1d(mask, mode='wrap')[i] for mask = 1 on [lo,hi): the taps k in [-radius, radius]
// with (i + k) mod n in [lo, hi).  cw[m] = w[0] + ... + w[m-1] (prefix sums, cw[0] = 0), so each of the
// (at most three) images of the box under the wrap contributes one difference of prefix sums.
__device__ __forceinline__ double box_taps(const double *cw, int radius, int a, int b)
{
    // taps k with a <= k < b, clipped to [-radius, radius]
    a = max(a, -radius);
    b = min(b, radius + 1);
    return b > a ? cw[b + radius] - cw[a + radius] : 0.0;
}

__device__ __forceinline__ float smoothed_box(int i, int lo, int hi, int n, const double *w, const double *cw,
                                              int radius)
{
    if (hi <= lo) return 0.f;
    if (cw != NULL && 2 * radius + 1 <= n) {
        double acc = box_taps(cw, radius, lo - i, hi - i);
        acc += box_taps(cw, radius, lo - n - i, hi - n - i);
        acc += box_taps(cw, radius, lo + n - i, hi + n - i);
        return (float)acc;
    }
    double acc = 0.0;
    for (int k = -radius; k <= radius; ++k) {
        int j = (i + k) % n;
        if (j < 0) j += n;
        if (j >= lo && j < hi) acc += w[k + radius];
    }
    return (float)acc;
}

__global__ void __launch_bounds__(256)
slice_vectors_kernel(const gx_chord *__restrict__ chord, const int32_t *__restrict__ bbox, int N, double r,
                     double max_voxels, double af_re, double af_im, double ped_re, double ped_im,
                     int fill_bkg, int sigma, const double *__restrict__ gauss, int radius,
                     float2 *base, float *my, float *mz, float2 *dmy)
{
    // prefix sums of the Gaussian taps (<= 2 * 4 * sigma + 2 entries), once per CTA
    extern __shared__ double s_cw[];
    const double *cw = NULL;
    if (sigma > 0 && 2 * radius + 2 <= 4096) {
        if (threadIdx.x == 0) {
            double c = 0.0;
            s_cw[0] = 0.0;
            for (int t = 0; t <= 2 * radius; ++t) { c += gauss[t]; s_cw[t + 1] = c; }
        }
        __syncthreads();
        cw = s_cw;
    }
    const int p = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const size_t o = (size_t)p * N + i;
    const int y_min = bbox[4 * p + 0], y_max = bbox[4 * p + 1], z_min = bbox[4 * p + 2], z_max = bbox[4 * p + 3];
    // base relative to the pedestal P = avg_f * max_voxels, in units of avg_f: base = d_rel * avg_f
    float d_rel;
    if (fill_bkg) {
        const gx_chord k = chord[p];
        const double inv_r = 1.0 / r;
        double x = __dmul_rn((double)i, r);
        double len = chord_length(k, x);
        double q = gx_floordiv(len, r, inv_r);
        double nm = (double)(long long)__dsub_rn(max_voxels, q);   // astype(int) truncates
        d_rel = (float)__dsub_rn(nm, max_voxels);                  // small integer: exact in fp32
        base[o] = make_float2((float)__dsub_rn(__dmul_rn(nm, af_re), ped_re),
                              (float)__dsub_rn(__dmul_rn(nm, af_im), ped_im));
    } else {
        base[o] = make_float2((float)-ped_re, (float)-ped_im);
        d_rel = (ped_re != 0.0 || ped_im != 0.0) ? (float)-max_voxels : 0.f;
    }
    float vy = 1.f, vz = 1.f;
    if (sigma > 0) {
        int lo, hi;
        py_slice(y_min + sigma, y_max - sigma, N, lo, hi);
        vy = smoothed_box(i, lo, hi, N, gauss, cw, radius);
        py_slice(z_min + sigma, z_max - sigma, N, lo, hi);
        vz = smoothed_box(i, lo, hi, N, gauss, cw, radius);
    }
    if (fill_bkg) {
        // voxelgrids.py:358-361: rows < z_min or >= z_max and columns <= y_min or
        // >= y_max are overwritten with the pedestal, i.e. are 0 relative to it
        if (!(i > y_min && i < y_max)) vy = 0.f;
        if (!(i >= z_min && i < z_max)) vz = 0.f;
    }
    my[o] = vy;
    mz[o] = vz;
    if (dmy) dmy[o] = make_float2(d_rel, vy);
}

extern "C" int gx_slice_vectors(const gx_chord *d_chord, const int32_t *d_bbox, int n_phi, int N, double r,
                                double max_voxels, double avg_f_re, double avg_f_im,
                                double pedestal_re, double pedestal_im,
                                int fill_bkg, int smooth_sigma, const double *d_gauss, int gauss_radius,
                                gx_float2 *d_base, float *d_my, float *d_mz, gx_float2 *d_dmy, void *stream)
{
    GX_REQUIRE(d_bbox && d_base && d_my && d_mz, "NULL pointer");
    GX_REQUIRE(!fill_bkg || d_chord, "fill_bkg needs chord constants");
    GX_REQUIRE(smooth_sigma <= 0 || d_gauss, "smooth needs Gaussian weights");
    GX_REQUIRE(n_phi > 0 && N >= 16, "bad sizes");
    const size_t smem = smooth_sigma > 0 ? (size_t)(2 * gauss_radius + 2) * sizeof(double) : 0;
    slice_vectors_kernel<<<dim3((N + 255) / 256, n_phi), 256, smem <= 32768 ? smem : 0, gx_stream(stream)>>>(
        d_chord, d_bbox, N, r, max_voxels, avg_f_re, avg_f_im, pedestal_re, pedestal_im, fill_bkg,
        smooth_sigma, d_gauss, gauss_radius, reinterpret_cast<float2 *>(d_base), d_my, d_mz,
        reinterpret_cast<float2 *>(d_dmy));
    return gx_check_launch("gx_slice_vectors");
}

// ------------------------------------------------------- staged row kernel ----
// Writes the reference's pre-FFT grid (parity probe T2 and the unfused path).
template <bool SPECIES>
__global__ void __launch_bounds__(PROJ_THREADS)
project_rows_kernel(ProjArgs a, float2 *grid)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *acc = reinterpret_cast<float2 *>(smem_raw);
    uint32_t *words = reinterpret_cast<uint32_t *>(acc + a.N);
    __shared__ float2 s_table[GX_MAX_SPECIES];
    const int z = blockIdx.x, p = blockIdx.y;
    const int N = a.N, tid = threadIdx.x, nt = blockDim.x;
    const int npair = (a.n_species + 1) >> 1;
    const double s = a.sn[p], c = a.cs[p], shift = a.yrange[2 * p];
    if (SPECIES && tid < GX_MAX_SPECIES) s_table[tid] = tid < a.n_species ? a.table[tid] : make_float2(0.f, 0.f);
    for (int y = tid; y < N; y += nt) acc[y] = make_float2(0.f, 0.f);
    if (SPECIES) for (int y = tid; y < npair * N; y += nt) words[y] = 0u;
    __syncthreads();
    const int beg = a.row_start[z], end = a.row_start[z + 1];
    if (SPECIES) {
        for (int c0 = beg; c0 < end; c0 += 65535) {      // 16-bit counters cannot wrap
            scatter_species(a, c0, min(c0 + 65535, end), s, c, shift, words, N);
            __syncthreads();
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
            __syncthreads();
        }
    } else {
        const double r = a.r, inv_r = 1.0 / a.r;
        for (int i = beg + tid; i < end; i += nt) {
            const double q = atom_y_pixel(a.xs[i], a.ys[i], s, c, shift, r, inv_r);
            if (q < (double)N) {
                const float2 f = a.f[i];
                atomicAdd(&acc[(int)q].x, f.x);
                atomicAdd(&acc[(int)q].y, f.y);
            }
        }
        __syncthreads();
    }
    const float mzv = a.mz[(size_t)p * N + z];
    const bool has_ped = a.fill_bkg || a.sigma > 0;
    float2 *dst = grid + ((size_t)p * N + z) * N;
    for (int y = tid; y < N; y += nt) {
        float2 v = finish_pixel(acc[y], a.base[(size_t)p * N + y], mzv * a.my[(size_t)p * N + y]);
        if (has_ped) { v.x += a.ped_re; v.y += a.ped_im; }
        dst[y] = v;
    }
}

extern "C" int gx_project_slices(const double *d_xs, const double *d_ys, const uint8_t *d_species,
                                 const gx_float2 *d_f, const int32_t *d_row_start,
                                 const gx_float2 *d_table, int n_species,
                                 const double *d_sin, const double *d_cos, const double *d_yrange,
                                 const int32_t *d_bbox, const gx_float2 *d_base, const float *d_my,
                                 const float *d_mz, int n_phi, int N, double r,
                                 double pedestal_re, double pedestal_im, int fill_bkg, int smooth_sigma,
                                 gx_float2 *d_grid, void *stream)
{
    GX_REQUIRE(d_xs && d_ys && d_row_start && d_sin && d_cos && d_yrange && d_bbox && d_base && d_my && d_mz &&
               d_grid, "NULL pointer");
    GX_REQUIRE(n_species >= 0 && n_species <= GX_MAX_SPECIES, "n_species out of range");
    GX_REQUIRE(n_species == 0 ? d_f != NULL : (d_species != NULL && d_table != NULL), "species/f inputs missing");
    GX_REQUIRE(n_phi > 0 && N >= 16, "bad sizes");
    ProjArgs a;
    a.xs = d_xs; a.ys = d_ys; a.species = d_species; a.f = reinterpret_cast<const float2 *>(d_f);
    a.row_start = d_row_start; a.table = reinterpret_cast<const float2 *>(d_table); a.n_species = n_species;
    a.sn = d_sin; a.cs = d_cos; a.yrange = d_yrange; a.bbox = d_bbox;
    a.base = reinterpret_cast<const float2 *>(d_base); a.my = d_my; a.mz = d_mz;
    a.N = N; a.r = r; a.ped_re = (float)pedestal_re; a.ped_im = (float)pedestal_im;
    a.fill_bkg = fill_bkg; a.sigma = smooth_sigma;
    cudaStream_t st = gx_stream(stream);
    float2 *grid = reinterpret_cast<float2 *>(d_grid);
    size_t smem = (size_t)N * sizeof(float2) + (size_t)((n_species + 1) / 2) * N * sizeof(uint32_t);
    if (smem > 227 * 1024) { gx_set_error("gx_project_slices: row of %d pixels x %d species needs %zu B shared memory", N, n_species, smem); return GX_ERR_UNSUPPORTED; }
    if (n_species > 0) {
        GX_CUDA(cudaFuncSetAttribute(project_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        project_rows_kernel<true><<<dim3(N, n_phi), PROJ_THREADS, smem, st>>>(a, grid);
    } else {
        GX_CUDA(cudaFuncSetAttribute(project_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        project_rows_kernel<false><<<dim3(N, n_phi), PROJ_THREADS, smem, st>>>(a, grid);
    }
    return gx_check_launch("gx_project_slices");
}
