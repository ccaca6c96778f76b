/*
 * giwaxs_b200 -- C ABI of the B200 (sm_100a) implementation of GIWAXSim's
 * reciprocal-space hot path.
 *
 * The reference (tchaney97/giwaxsim) is pure Python; the interface this
 * library replaces is the set of NumPy/SciPy calls inside
 *   tools/voxelgrids.py:311-416  rotate_project_fft_coords   (one phi slice)
 *   tools/voxelgrids.py:464-506  process_file2               (3-D binning)
 *   tools/comparison.py:765-786  sum/count, crop, f0 weight  (finalise)
 *   tools/detector.py:33-162     rotate_about_*              (plane rotation)
 *   tools/detector.py:194-232    intersect_detector          (gather)
 *   tools/detector.py:246-275    mirror_vertical_horizontal  (epilogue)
 * and is bound from Python through ctypes (giwaxsim_b200/_lib.py; the stub a
 * reference maintainer would add is shown in INTEGRATION.md).
 *
 * Conventions
 *  - plain C: pointers, sizes and doubles only.  Pointers named d_* are
 *    DEVICE pointers (the Python host passes torch tensor.data_ptr()); h_*
 *    are host pointers.  `stream` is a cudaStream_t passed as void*
 *    (NULL = legacy default stream).  Nothing here allocates device memory.
 *  - every function returns GX_OK (0) or a negative GX_ERR_* code;
 *    gx_last_error() returns a thread-local message for the last failure.
 *  - angle-dependent scalars (sin, cos, linspace end points, chord-length
 *    constants) are evaluated on the host with the same NumPy expressions the
 *    reference uses and passed in as fp64, so that every integer index the
 *    device derives is bit-identical to the reference's.
 *  - there is no CPU fallback: without an sm_100 device every compute entry
 *    point fails with GX_ERR_NO_DEVICE.
 */
#ifndef GIWAXS_B200_H
#define GIWAXS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GX_ABI_VERSION 5   /* 4: gx_fused_args.phases, gx_comm_*, column range of gx_voxel_finalize, host-boundary calls; 5: gx_host_widen_f32_f64 */

#define GX_OK 0
#define GX_ERR_INVALID (-1)     /* bad argument                              */
#define GX_ERR_CUDA (-2)        /* CUDA runtime error (see gx_last_error)    */
#define GX_ERR_UNSUPPORTED (-3) /* size outside what the kernels implement   */
#define GX_ERR_NO_DEVICE (-4)   /* no sm_100 GPU visible                     */

#define GX_MAX_SPECIES 16       /* distinct f-values on the counting path    */
#define GX_FFT_MAX_LOG2 13      /* largest in-smem pow2 transform (8192)     */

typedef struct gx_float2 { float x, y; } gx_float2;

int gx_abi_version(void);
const char *gx_last_error(void);
/* GX_OK iff `device` exists, is compute capability 10.x and can run the
 * embedded sm_100a image. */
int gx_device_check(int device);

/* ------------------------------------------------------------------ atoms */
/* min/max of each coordinate column of an [A,3] fp64 array.
 * d_out6 = {xmin,xmax,ymin,ymax,zmin,zmax}.      (comparison.py:712-714)   */
int gx_coords_minmax(const double *d_coords, int64_t A, double *d_out6, void *stream);

/* z pixel row of every atom (phi-invariant) and a counting sort by row.
 *   z_idx = (z - z_min) // r ; rows >= N are "invalid" and sort last.
 * Outputs: d_xs,d_ys [A] coordinates in row order; d_perm [A] original index
 * of each sorted atom; d_species_out/d_f_out the permuted per-atom species
 * code / complex64 f (either input may be NULL); d_row_start [N+2] with
 * row_start[z]..row_start[z+1] the atoms of row z and row_start[N+1] == A.
 * d_cursor is scratch of N+2 int32.          (voxelgrids.py:324,329,332)   */
int gx_atoms_sort_rows(const double *d_coords, int64_t A, double z_min, double r, int N,
                       const uint8_t *d_species, const gx_float2 *d_f,
                       double *d_xs, double *d_ys, int32_t *d_perm,
                       uint8_t *d_species_out, gx_float2 *d_f_out,
                       int32_t *d_row_start, int32_t *d_cursor, void *stream);

/* Species coding on the device (comparison.py:735-739 looks f up per atom from
 * the element symbol).  d_codepoints: the NumPy '<U1' / '<U2' element array
 * viewed as uint32 [A][width].  gx_species_histogram fills d_hist[16385]:
 * bin k < 16384 counts atoms with key cp0 + (cp1 << 7), the last bin counts
 * atoms with a non-ASCII code point.  gx_species_codes writes
 * d_codes[i] = d_lut[key_i] (d_lut: 16384 uint8, built by the host from the
 * occupied bins).                                                          */
int gx_species_histogram(const uint32_t *d_codepoints, int width, int64_t A, uint32_t *d_hist, void *stream);
int gx_species_codes(const uint32_t *d_codepoints, int width, int64_t A, const uint8_t *d_lut,
                     uint8_t *d_codes, void *stream);

/* Wrap-around sum of the 64-bit words of d_data [n]; widen_f32 != 0: d_data is
 * fp32 and each value is widened to fp64 first (the checksum of the float64
 * host copy a driver returns).  Lets voxelgridmaker_fitting -> detectormaker_fitting
 * and slabmaker_fitting -> voxelgridmaker_fitting reuse the device copy of an
 * array only while the caller's host array still has the same content
 * (the reference always reads the host array: comparison.py:673,790).         */
int gx_checksum64(const void *d_data, int64_t n, int widen_f32, uint64_t *d_out, void *stream);

/* d_out2 = {max over z rows of sum |Re f|, max over z rows of sum |Im f|} of the atoms
 * sorted by gx_atoms_sort_rows (species codes + h_table_abs [n_species][2] = |Re f_s|,
 * |Im f_s|, or per-atom d_f): the most a single pixel of a row can receive, which sizes
 * the fixed-point scale of gx_slices_fused (gx_fused_args.max_row_abs_re / _im).      */
int gx_row_abs_f_max(const uint8_t *d_species, const gx_float2 *d_f, const int32_t *d_row_start, int N,
                     const double *h_table_abs, int n_species, double *d_out2, void *stream);

/* min and max over all atoms of y' = fma(y, cos, x*sin) for n_phi rotations.
 * d_yrange [n_phi][2].                   (utilities.py:303-317, vg.py:323) */
int gx_slice_yrange(const double *d_xs, const double *d_ys, int64_t A,
                    const double *d_sin, const double *d_cos, int n_phi,
                    double *d_yrange, void *stream);

/* Candidate reduction for gx_slice_yrange.  Only atoms on (or within eps of)
 * the boundary of the convex hull of the (x,y) positions can attain the min or
 * max of y' for any rotation.
 *  gx_extreme_atoms: for n <= 64 rotations whose d_yrange is known, the lowest
 *    atom index attaining the min and the max -> d_index [n][2].
 *  gx_hull_filter: d_edges [m][3] = (nx, ny, d) with nx*x + ny*y + d the signed
 *    distance to edge k of a convex polygon whose vertices are atoms (inside
 *    positive); atoms whose distance to the nearest edge is <= eps are appended
 *    to d_xs_out/d_ys_out; *d_count receives how many qualified (may exceed
 *    `capacity`, in which case the caller must fall back to all atoms).      */
int gx_extreme_atoms(const double *d_xs, const double *d_ys, int64_t A,
                     const double *d_sin, const double *d_cos, const double *d_yrange, int n,
                     int32_t *d_index, void *stream);
int gx_hull_filter(const double *d_xs, const double *d_ys, int64_t A, const double *d_edges, int m,
                   double eps, int32_t *d_count, double *d_xs_out, double *d_ys_out, int capacity,
                   void *stream);

/* index bounding box {y_min,y_max,z_min,z_max} of the valid atoms of every
 * rotation; -1 entries when no atom is valid.  d_scratch: n_phi + 1 int32.
 *                                                        (vg.py:332-349)   */
int gx_slice_bbox(const double *d_xs, const double *d_ys, const int32_t *d_row_start, int N,
                  double r, const double *d_sin, const double *d_cos, const double *d_yrange,
                  int n_phi, int32_t *d_bbox, int32_t *d_scratch, void *stream);

/* per-atom pixel indices of ONE rotation in the ORIGINAL atom order (parity
 * probe T1): y_idx, z_idx as the reference's astype(int) values.           */
int gx_atom_pixel_indices(const double *d_xs, const double *d_ys, const int32_t *d_perm,
                          const int32_t *d_row_start, int64_t A, int N, double r,
                          double sin_phi, double cos_phi, double y_shift,
                          int64_t *d_y_idx, int64_t *d_z_idx, void *stream);

/* ------------------------------------------------ per-slice row vectors  */
/* chord-length constants of rectangular_collapse_lengths (vg.py:253-285),
 * evaluated on the host per rotation. */
typedef struct gx_chord {
    double hor, ver;          /* after the phi>90 swap                     */
    double stop1, stop2, stop12, mid;
    double vcos, rise, tan_phi, tan_theta, cos_phi, cos_theta;
    int32_t mode;             /* 0: phi==0, 1: phi==90, 2: general         */
    int32_t pad;
} gx_chord;

/* Builds, for n_phi rotations, the three length-N vectors the row kernel
 * consumes:
 *   d_base [n_phi][N] complex64 = num_missing[y]*avg_f - pedestal
 *                                 (or -pedestal when !fill_bkg)
 *   d_my, d_mz [n_phi][N] fp32  = Gaussian-smoothed box masks (smooth>0)
 *   d_dmy [n_phi][N] float2     = (num_missing - max_voxels, my): the packed
 *                                 form the fused row kernel reads (base =
 *                                 dmy.x * avg_f); optional (NULL to skip)
 * h_gauss: host array of 2*radius+1 fp64 weights (scipy _gaussian_kernel1d).
 *                                                   (vg.py:344-379)        */
int gx_slice_vectors(const gx_chord *d_chord, const int32_t *d_bbox, int n_phi, int N, double r,
                     double max_voxels, double avg_f_re, double avg_f_im,
                     double pedestal_re, double pedestal_im,
                     int fill_bkg, int smooth_sigma, const double *d_gauss, int gauss_radius,
                     gx_float2 *d_base, float *d_my, float *d_mz, gx_float2 *d_dmy, void *stream);

/* ------------------------------------------------------ projection (K1)  */
/* Scatter + background + edge blend for n_phi rotations; one CTA per
 * (rotation,row), atoms counted per species in shared memory.
 * d_table: species f-values (complex64, n_species <= GX_MAX_SPECIES), or
 * n_species == 0 to use per-atom d_f (generic path).
 * d_grid [n_phi][N][N] complex64 receives the reference's pre-FFT grid
 * (rows = z, cols = y).                               (vg.py:338-379)      */
int gx_project_slices(const double *d_xs, const double *d_ys, const uint8_t *d_species,
                      const gx_float2 *d_f, const int32_t *d_row_start,
                      const gx_float2 *d_table, int n_species,
                      const double *d_sin, const double *d_cos, const double *d_yrange,
                      const int32_t *d_bbox, const gx_float2 *d_base, const float *d_my,
                      const float *d_mz, int n_phi, int N, double r,
                      double pedestal_re, double pedestal_im, int fill_bkg, int smooth_sigma,
                      gx_float2 *d_grid, void *stream);

/* ------------------------------------------------------------- FFT (K2)  */
/* Size in bytes of the plan table for transform length N (any N with
 * 16 <= N, pow2 up to 8192 or Bluestein with 2N-1 <= 8192); <0 on error.   */
int64_t gx_fft_plan_bytes(int N);
/* Fill a HOST buffer of gx_fft_plan_bytes(N) bytes (twiddles, chirp, filter
 * spectrum; all computed in fp64, stored fp32).  Copy it to the device and
 * pass that pointer as d_plan below. */
int gx_fft_plan_fill(int N, void *h_plan);

/* |fftshift(fft2(grid))|^2 for `batch` N x N complex64 grids -> fp32.
 * d_work: batch*N*N complex64 scratch.  dc_re/dc_im is added to the
 * unshifted (0,0) coefficient before squaring (pedestal * N^2), or 0.
 *                                                     (vg.py:388-392)      */
int gx_fft2_abs2_shift(const gx_float2 *d_grid, gx_float2 *d_work, float *d_iq2d,
                       int batch, int N, const void *d_plan, double dc_re, double dc_im,
                       void *stream);

/* ---------------------------------------------------------- binning (K3) */
/* Column / row voxel indices of a slice.  Column j has
 *   qx = linspace(qx_left, qx_right, N)[j], qy likewise (NumPy rounding);
 * kept iff qmin <= q <= qmax on both; index = (q - qmin) // dq.
 * d_col [n_phi][N] packs iy*q_num+ix, or -1 if masked.  (vg.py:396-401,475) */
int gx_slice_col_index(const double *d_qx_left, const double *d_qx_right,
                       const double *d_qy_left, const double *d_qy_right, int n_phi, int N,
                       double qmin, double qmax, double dq, int q_num, int32_t *d_col, void *stream);
/* Same from explicit axis values (process_file2's det_h_qx / det_h_qy).    */
int gx_axis_col_index(const double *d_qx, const double *d_qy, int n, double qmin, double qmax,
                      double dq, int q_num, int32_t *d_col, void *stream);
/* Row index iz (or -1) from explicit qz values (det_v_qz).  (vg.py:476,499) */
int gx_axis_row_index(const double *d_qz, int n, double qmin, double qmax, double dq, int q_num,
                      int32_t *d_row, void *stream);

/* Accumulate `batch` slices: sum[iy,ix,iz] += I ; and either the full 3-D
 * count (d_count3 != NULL) or the rank-1 factor H[iy,ix] += 1 per kept
 * column (d_count2 != NULL; the row factor is gx_row_histogram).
 *                                                     (vg.py:484-503)      */
int gx_bin_slices(const float *d_iq2d, int batch, int rows, int cols,
                  const int32_t *d_col, int col_stride, const int32_t *d_row, int q_num,
                  float *d_sum, uint32_t *d_count3, uint32_t *d_count2, void *stream);
int gx_row_histogram(const int32_t *d_row, int n, int q_num, uint32_t *d_m, void *stream);

/* iq = sum/count (0 where count == 0), cropped to [lo,hi)^3, times
 * ((sum_i a_i exp(-b_i q^2/16pi^2) + c)/Z)^2.  count = count3, or
 * count2[iy,ix]*m[iz] when d_count3 is NULL.  d_axis [q_num] fp64 voxel axis.
 * d_iq [(hi-lo)^3] fp32.     (comparison.py:769-786, vg.py:16-48,828-857)
 * h_aff9 == NULL: plain sum/count without the f0 weight, as returned by
 * generate_voxel_grid_low_mem (vg.py:633-641).
 * Only the (iy, ix) columns col_begin <= iy*V+ix < col_end of the cropped grid
 * are written (col_end < 0: all) - N ranks finalise 1/N of the grid each after
 * a reduce-scatter of the partial sums.                                     */
int gx_voxel_finalize(const float *d_sum, const uint32_t *d_count3, const uint32_t *d_count2,
                      const uint32_t *d_m, int q_num, int lo, int hi, const double *d_axis,
                      const double *h_aff9, double Z, int64_t col_begin, int64_t col_end,
                      float *d_iq, void *stream);

/* d_iq[iy,ix,iz] *= factor where lower < sqrt(qx^2+qy^2+qz^2) <= upper, the
 * radius formed like NumPy forms it (each product, sum and the root rounded
 * once): the shell mask of the aff_num_qs > 1 branch (vg.py:650-707).
 * d_iq [V^3] fp32 in place, d_axis [V] fp64.                                 */
int gx_voxel_shell_scale(float *d_iq, int V, const double *d_axis, double lower, double upper,
                         double factor, void *stream);

/* ------------------------------------------- fused slice pipeline (A) -- */
/* [first, one-past-last) kept column of every rotation from the table that
 * gx_slice_col_index wrote (the kept set is one interval).  d_range [n_phi][2] */
int gx_slice_col_range(const int32_t *d_col, int n_phi, int N, int32_t *d_range, void *stream);

/* Everything rotate_project_fft_coords + process_file2 do for a batch of
 * rotations (voxelgrids.py:311-416,464-506), in two launches: per-row scatter +
 * background + blend + FFT along y with only the kept q-columns written to
 * d_work [n_phi][N][KC] complex64; per-column-tile FFT along z + |.|^2 +
 * accumulation into d_sum [q_num^3] fp32 and d_count2 [q_num^2] u32 (the count
 * of voxel (iy,ix,iz) is d_count2[iy,ix] * m[iz], m from gx_row_histogram).
 * The per-rotation tables are those of gx_slice_yrange / gx_slice_bbox /
 * gx_slice_vectors / gx_slice_col_index / gx_slice_col_range; row_lo/row_hi
 * bound the kept shifted rows of d_row_index.  For n_phi <= 128 the small
 * per-rotation tables and d_row_start are copied device-to-device into
 * __constant__ memory on `stream` first: issue the fused launches of one
 * device on one stream. */
typedef struct gx_fused_args {
    const double *d_xs, *d_ys;
    const uint8_t *d_species;
    const gx_float2 *d_f;
    const int32_t *d_row_start;
    const gx_float2 *d_table;
    const double *d_sin, *d_cos, *d_yrange;
    const int32_t *d_bbox;
    const gx_float2 *d_dmy;     /* gx_slice_vectors: (num_missing - max_voxels, my) */
    const float *d_mz;
    const void *d_plan;
    const int32_t *d_col, *d_colrange, *d_row_index;
    gx_float2 *d_work;
    float *d_sum;
    uint32_t *d_count2;
    double *d_dc;               /* [16] zero-initialised, or NULL: fp64 side sums of the k = 0 sample of the centre column of
                                   every slice (8 sums + 8 voxel keys); fold into d_sum with gx_fold_dc after the last batch */
    double r, pedestal_re, pedestal_im, avg_f_re, avg_f_im;
    int32_t n_species, n_phi, N, KC, q_num, row_lo, row_hi, fill_bkg, smooth_sigma;
    int32_t phases;              /* 0 or 3: both launches; 1: row kernel only; 2: column kernel only (per-kernel timing) */
    int32_t max_row_atoms;       /* most atoms in one z pixel row (0: unknown -> 65535) */
    int32_t pad;
    double max_row_abs_re, max_row_abs_im;   /* largest sum over one z row of |Re f| / |Im f| (0: atoms x largest f):
                                                sizes the fixed-point scale of the row accumulators */
    double max_abs_f_re, max_abs_f_im;       /* n_species == 0: bounds on |Re f|, |Im f| of one atom (0: 128) */
    double table_f64[2 * GX_MAX_SPECIES];   /* host copy of the species f-values, (re, im) pairs, full precision */
} gx_fused_args;
int gx_slices_fused(const gx_fused_args *h_args, void *stream);
int gx_fold_dc(double *d_dc, float *d_sum, void *stream);
/* 1 when, for this grid side and kept-column bound, gx_slices_fused feeds the column
 * transform by TMA from a work buffer in permuted row order: d_work must then have
 * been zero-filled once before the first call of a run (rows outside the atom band
 * are never written, and the TMA boxes read every row slot).                    */
int gx_fused_wants_zeroed_work(int N, int KC);

/* Restrict bin indices to the crop window lo <= i < hi of downselect_voxelgrid
 * (voxelgrids.py:16-48; the crop commutes with the accumulation): entries
 * outside become -1, the others are re-based to a [hi-lo]^3 grid.  packed != 0:
 * entries are iy*q_num+ix (gx_slice_col_index), else plain iz
 * (gx_axis_row_index).  In place.                                           */
int gx_window_indices(int32_t *d_index, int64_t n, int q_num, int lo, int hi, int packed, void *stream);

/* ------------------------------------------------- slab builder (next row N3) */
/* slabmaker_fitting on the device (comparison.py:605-671): replicas of the unit
 * cell  x = ((x0 + ax*i) + bx*j) + cx*k,  y = (y0 + by*j) + cy*k,  z = z0 + cz*k
 * for i < nx, j < ny, k < nz (nx = num_x + 1, ...), every product and sum rounded
 * in fp64 like NumPy; atom order k, j, i, atom.  The replicated array is never
 * stored; the three passes recompute it.
 *  gx_slab_minmax : d_out6 = {xmin,xmax,ymin,ymax,zmin,zmax} over all replicas
 *  gx_slab_count  : with p = x - min (h_min3) keep h_lo3 <= p <= h_hi3 on all
 *                   axes; d_tile_count [gx_slab_tiles()] receives the EXCLUSIVE
 *                   prefix of the kept atoms per tile of 1024 replica atoms,
 *                   *d_total their number, d_kept_min3 the minimum p of the kept
 *  gx_slab_write  : kept atoms in reference order: d_xyz_out [M][3] = p -
 *                   h_kept_min3, d_species_out [M] = species of the source atom  */
typedef struct gx_slab_args {
    const double *d_cell_xyz;      /* [n_cell][3] unit-cell coordinates      */
    const uint8_t *d_cell_species; /* [n_cell] or NULL                        */
    int64_t n_cell;
    int32_t nx, ny, nz, pad;
    double ax, bx, by, cx, cy, cz;
} gx_slab_args;
int64_t gx_slab_tiles(const gx_slab_args *h_args);
int gx_slab_minmax(const gx_slab_args *h_args, double *d_out6, void *stream);
int gx_slab_count(const gx_slab_args *h_args, const double *h_min3, const double *h_lo3, const double *h_hi3,
                  int64_t *d_tile_count, int64_t *d_total, double *d_kept_min3, void *stream);
int gx_slab_write(const gx_slab_args *h_args, const double *h_min3, const double *h_lo3, const double *h_hi3,
                  const double *h_kept_min3, const int64_t *d_tile_offset, double *d_xyz_out,
                  uint8_t *d_species_out, void *stream);

/* ------------------------------- post-hoc image transforms (next row N4) */
/* The polar warp pair of shift_peak and the masked linear fit of optimize_scale_offset
 * (tools/comparison.py:469-592, 873-882), fp64, on the trimmed detector image.
 * Sampling = scipy.ndimage.map_coordinates(order=1, mode='constant', cval).
 *  gx_polar_warp   : linear_polar - d_out [out_h][out_w], radius linspace(0, r, out_w),
 *                    angle linspace(0, 2 pi, out_h) about (o_row, o_col)
 *  gx_polar_unwarp : polar_linear - d_out [out_h][out_w] from d_polar [ph][pw], radius r
 *  gx_gather_columns : d_out[i][j] = d_src[i][d_map[j]], zeroed where d_zero_ref[i][j] == 0
 *                    (add_pad's column duplication + the zero mask; d_zero_ref may be NULL)
 *  gx_masked_fit_sums: d_out5 = {n, sum x, sum y, sum x^2, sum x y} over mask == 0           */
int gx_polar_warp(const double *d_img, int rows, int cols, double o_row, double o_col, double r,
                  int out_h, int out_w, double cval, double *d_out, void *stream);
int gx_polar_unwarp(const double *d_polar, int ph, int pw, double r, double o_row, double o_col,
                    int out_h, int out_w, double cval, double *d_out, void *stream);
int gx_gather_columns(const double *d_src, int rows, int src_cols, const int32_t *d_map, int out_cols,
                      const double *d_zero_ref, double *d_out, void *stream);
int gx_masked_fit_sums(const double *d_x, const double *d_y, const double *d_mask, int64_t n,
                       double *d_out5, void *stream);

/* The same accumulation as gx_detector_accumulate_affine with the gather fed by TMA: for every
 * (32 x 16-pixel tile, orientation) the 8 x 8 x 12 brick of voxels the tile can touch is
 * streamed into shared memory (cp.async.bulk.tensor.3d, mbarrier ring) and the pixels read
 * it there.  d_iq_padded: copy of d_iq with rows of Vz_padded floats (multiple of 4: tensor
 * map strides are multiples of 16 bytes).  Precondition (caller): (31 |U_a| + 15 |V_a|) / 2^F
 * < 7 for every record and axis, so that a tile's voxel span fits the brick.  A/B variant of
 * north_star's kernel (4): identical result, measured 45 % slower than the L1-fed gather
 * (DESIGN.md section 4.4).                                                               */
int gx_detector_accumulate_affine_brick(const float *d_iq, const float *d_iq_padded, int Vz_padded, int Vy,
                                        int Vx, int Vz, double qx_min, double qy_min, double qz_min, double dq,
                                        const double *d_px, const double *d_py, const double *d_pz, int rows,
                                        int cols, const double *h_corners9, const void *d_records,
                                        const double *d_R, int n_orient, const double *h_plan, double *d_image,
                                        void *stream);

/* ----------------------------------------------------------- host boundary */
/* Results are returned as float64 host arrays (the reference's types).  With N ranks on a
 * node the array lives in a pooled shared-memory segment mapped by every rank; each rank
 * page-locks ITS slab once (gx_host_register) and from then on DMAs the widened slab
 * straight into place (gx_copy_to_host_async) - no host-side copy or conversion.        */
int gx_host_register(void *h_ptr, int64_t nbytes);
int gx_host_unregister(void *h_ptr);
int gx_copy_to_host_async(void *h_dst, const void *d_src, int64_t nbytes, void *stream);
/* One rank: the fp32 grid crosses PCIe chunk by chunk into a page-locked staging buffer and is
 * widened there -> the float64 array the caller receives (comparison.py:769-786 returns float64):
 * h_dst[i] = (double)h_src[i], i < n, on `threads` host threads (a persistent pool inside the
 * library; the calling thread is one of them) with non-temporal stores.  Host-only: no CUDA call. */
int gx_host_widen_f32_f64(const float *h_src, double *h_dst, int64_t n, int threads);

/* ----------------------------------------------------- multi-GPU exchange */
/* One process per GPU; the partial voxel sums / counts of stage A and the
 * partial detector images of stage B meet in a sum (the reference's shared
 * accumulators: voxelgrids.py:502-503, detector.py:298).  NCCL (the libnccl.so.2
 * already loaded in the process) is driven through these entry points, on the
 * caller's stream, so a collective follows the kernel that produced its input
 * without a host synchronisation.
 *  gx_comm_unique_id : rank 0 fills 128 bytes; ship them to the other ranks
 *                      (any out-of-band channel) and pass them to gx_comm_init
 *  gx_comm_init      : collective over `world` ranks; the calling thread's
 *                      current CUDA device is this rank's GPU
 *  gx_comm_all_reduce: in-place sum; dtype GX_DTYPE_F32 / _U32 / _F64
 *  gx_comm_reduce_scatter_f32 : d_buf [world * count_per_rank]; on return slab
 *                      `rank` of this rank's buffer holds the sum over ranks
 *  gx_comm_all_gather_f32     : slab `rank` of every rank -> all of d_buf on all  */
#define GX_DTYPE_F32 0
#define GX_DTYPE_U32 1
#define GX_DTYPE_F64 2
int gx_comm_unique_id(void *h_id128);
int gx_comm_init(const void *h_id128, int rank, int world, void **out_comm);
int gx_comm_destroy(void *comm);
int gx_comm_all_reduce(void *comm, void *d_buf, int64_t count, int dtype, void *stream);
int gx_comm_reduce_scatter_f32(void *comm, float *d_buf, int64_t count_per_rank, void *stream);
int gx_comm_all_gather_f32(void *comm, float *d_buf, int64_t count_per_rank, void *stream);

/* --------------------------------------------------------- detector (K4) */
/* p <- R p for n points, R row-major 3x3, fma chain k=0,1,2.
 *                                                  (detector.py:71,113,155) */
int gx_rotate_points(const double *h_R9, const double *d_x, const double *d_y, const double *d_z,
                     int64_t n, double *d_ox, double *d_oy, double *d_oz, void *stream);

/* For every pixel and each of n_orient orientations: three sequential
 * rotations (d_R [n_orient][3][9]), floor-bin to the voxel grid, clamp,
 * gather, weight, accumulate into d_image (fp64, += semantics).
 * d_iq [Vy][Vx][Vz] fp32.  d_index_out (optional) receives the clamped flat
 * voxel index of every pixel for orientation `probe` (parity probe T7).
 *                                       (detector.py:194-244, 289-298)     */
int gx_detector_accumulate(const float *d_iq, int Vy, int Vx, int Vz,
                           double qx_min, double qy_min, double qz_min, double dq,
                           const double *d_px, const double *d_py, const double *d_pz, int64_t n_pix,
                           const double *d_R, const double *d_w, int n_orient,
                           double *d_image, int probe, int64_t *d_index_out, void *stream);

/* Same result as gx_detector_accumulate, bit for bit, at a fraction of the
 * fp64 work: the voxel coordinate of every pixel is first evaluated in fp32
 * from one collapsed matrix per orientation; only pixels whose fp32 coordinate
 * lies within a host-computed rigorous error bound of a voxel edge fall back to
 * the exact fp64 chain.  d_fast: n_orient records from
 * gx_host_fast_orientations; d_slow_count (optional) counts fall-backs.     */
int gx_detector_accumulate_fast(const float *d_iq, int Vy, int Vx, int Vz,
                                double qx_min, double qy_min, double qz_min, double dq,
                                const double *d_px, const double *d_py, const double *d_pz, int64_t n_pix,
                                const void *d_fast, const double *d_R, int n_orient,
                                double *d_image, int probe, int64_t *d_index_out,
                                unsigned long long *d_slow_count, void *stream);
/* Host helper: records for the kernel above.  h_R [n][3][9], h_w [n]; h_pmax[3]
 * = max |x|, |y|, |z| over the detector pixels; h_fast: n records of
 * gx_fast_record_bytes() bytes each.                                        */
int gx_fast_record_bytes(void);
int gx_host_fast_orientations(const double *h_R, const double *h_w, int n, double qx_min, double qy_min,
                              double qz_min, double dq, const double *h_pmax, void *h_fast);

/* Production detector kernel: same result as gx_detector_accumulate, bit for
 * bit in the voxel indices, for a pixel grid that is affine in (row, col)
 * (make_detector + rotations, detector.py:5-162).  Voxel coordinates are
 * evaluated in 32-bit fixed point from a per-orientation affine model that
 * interpolates the reference's exact corner values; pixels within a rigorous
 * error bound of a voxel edge fall back to the exact fp64 chain.
 *   1. gx_grid_affine_fit: corners p[0,0], p[0,-1], p[-1,0] (h_corners9) and
 *      max deviation of the grid from their interpolation per component
 *      (h_dev3); d_scratch3 = 3 doubles of device scratch.  Synchronises.
 *   2. gx_host_affine_orientations: n records of gx_affine_record_bytes()
 *      and a plan of gx_affine_plan_doubles() doubles {frac bits, half-width
 *      of the edge band in fixed-point units, coordinate offset, in-box
 *      radius^2, error bound in voxels, #components the host proved constant,
 *      #edge-locked orientations (a nearly constant coordinate on a voxel
 *      edge that could not be modelled: every pixel takes the exact path),
 *      #components modelled as a single rounding step}; GX_ERR_UNSUPPORTED
 *      if the grid is not affine enough -> use gx_detector_accumulate.
 *   3. gx_detector_accumulate_affine: the gather; d_records = device copy of
 *      the records.  d_image fp64, += semantics (atomic when the orientation
 *      range is split).                     (detector.py:194-244, 289-298)  */
int gx_grid_affine_fit(const double *d_px, const double *d_py, const double *d_pz, int rows, int cols,
                       double *d_scratch3, double *h_corners9, double *h_dev3, void *stream);
int gx_affine_record_bytes(void);
int gx_affine_plan_doubles(void);
int gx_host_affine_orientations(const double *h_corners9, const double *h_dev3, int rows, int cols,
                                const double *h_R, const double *h_w, int n, double qx_min, double qy_min,
                                double qz_min, double dq, int Vy, int Vx, int Vz, void *h_records,
                                double *h_plan);
int gx_detector_accumulate_affine(const float *d_iq, int Vy, int Vx, int Vz, double qx_min, double qy_min,
                                  double qz_min, double dq, const double *d_px, const double *d_py,
                                  const double *d_pz, int rows, int cols, const double *h_corners9,
                                  const void *d_records, const double *d_R, int n_orient,
                                  const double *h_plan, double *d_image, int probe, int64_t *d_index_out,
                                  unsigned long long *d_slow_count, void *stream);

/* Host helper: the three per-orientation rotation matrices, derived exactly
 * as rotate_psi_phi_theta does from the current corner pixels.
 * h_corners [3][3] = p[0,0], p[0,-1], p[-1,0] of the base detector;
 * h_cs [n][6] = cos,sin of radians(psi), radians(phi), radians(theta)
 * evaluated with NumPy; h_R [n][3][9].   (detector.py:58-68,104-110,146-152,
 * utilities.py:222-245)                                                    */
int gx_host_orientation_matrices(const double *h_corners, const double *h_cs, int n, double *h_R);

/* mirror != 0: four-fold mirror with the odd-size centre rules
 * (detector.py:246-275).  finish == 1: NaN/<=0 -> 1e-6, then *1e-6
 * (comparison.py:859-868); finish == 2: the floor only, as the two-step
 * driver does (old_modules/detectormaker.py:152-153).  Out of place.        */
int gx_detector_epilogue(const double *d_image, int rows, int cols, int mirror, int finish,
                         double *d_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GIWAXS_B200_H */
