// Slab builder on the device (tools/comparison.py:605-671):
// tile the unit cell (num_x+1) x (num_y+1) x (num_z+1) times and cut the centred
// orthorhombic slab, with the reference's arithmetic reproduced operation by
// operation so that every coordinate is bit-identical:
//     x = ((x0 + ax*i) + bx*j) + cx*k      each product and sum rounded (fp64)
//     y = ( y0         + by*j) + cy*k
//     z =   z0                 + cz*k
// (i, j, k = 0 adds an exact zero, as the reference's untouched first copy),
// then  p = x - min(x),  keep  buf <= p <= ext - buf  on all three axes,
// then  out = p - min(p over the kept atoms).
// The replicated array (8x the slab) is never materialised: three passes
// recompute the coordinates (min/max; count + kept minimum; ordered write).
// Output order is the reference's (k outermost, then j, then i, then the atom),
// produced by a block-level exclusive scan of per-block counts.
#include "gx_common.cuh"

#define SLAB_THREADS 256
#define SLAB_ITEMS 4
#define SLAB_TILE (SLAB_THREADS * SLAB_ITEMS)

struct SlabGeom {
    const double *cell;          // [n0][3]
    const uint8_t *species;      // [n0] or NULL
    int64_t n0, total;
    int nx, ny, nz;
    double ax, bx, by, cx, cy, cz;
};

__device__ __forceinline__ void slab_point(const SlabGeom &g, int64_t idx, double &x, double &y, double &z, int64_t &t)
{
    t = idx % g.n0;
    int64_t rep = idx / g.n0;
    const int i = (int)(rep % g.nx);
    rep /= g.nx;
    const int j = (int)(rep % g.ny);
    const int k = (int)(rep / g.ny);
    const double fi = (double)i, fj = (double)j, fk = (double)k;
    x = g.cell[3 * t]; y = g.cell[3 * t + 1]; z = g.cell[3 * t + 2];
    x = __dadd_rn(__dadd_rn(__dadd_rn(x, __dmul_rn(g.ax, fi)), __dmul_rn(g.bx, fj)), __dmul_rn(g.cx, fk));
    y = __dadd_rn(__dadd_rn(y, __dmul_rn(g.by, fj)), __dmul_rn(g.cy, fk));
    z = __dadd_rn(z, __dmul_rn(g.cz, fk));
}

__device__ __forceinline__ void slab_minmax_commit(double lo, double hi, unsigned long long *mn, unsigned long long *mx)
{
    for (int s = 16; s > 0; s >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, s));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    }
    if ((threadIdx.x & 31) == 0) {
        if (mn) atomicMin(mn, gx_ord(lo));
        if (mx) atomicMax(mx, gx_ord(hi));
    }
}

__global__ void slab_init_kernel(unsigned long long *o6)
{
    // {xmin, xmax, ymin, ymax, zmin, zmax} in the order-preserving encoding
    if (threadIdx.x < 6) o6[threadIdx.x] = (threadIdx.x & 1) ? 0ull : ~0ull;
}

__global__ void slab_decode_kernel(unsigned long long *o, int n)
{
    if (threadIdx.x < n) reinterpret_cast<double *>(o)[threadIdx.x] = gx_unord(o[threadIdx.x]);
}

__global__ void __launch_bounds__(SLAB_THREADS)
slab_minmax_kernel(SlabGeom g, unsigned long long *o6)
{
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < g.total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        double p[3];
        int64_t t;
        slab_point(g, idx, p[0], p[1], p[2], t);
        for (int a = 0; a < 3; ++a) { lo[a] = fmin(lo[a], p[a]); hi[a] = fmax(hi[a], p[a]); }
    }
    for (int a = 0; a < 3; ++a) slab_minmax_commit(lo[a], hi[a], o6 + 2 * a, o6 + 2 * a + 1);
}

struct SlabCut { double mn[3], lo[3], hi[3], kept_min[3]; };

__device__ __forceinline__ bool slab_keep(const SlabGeom &g, const SlabCut &c, int64_t idx, double p[3], int64_t &t)
{
    slab_point(g, idx, p[0], p[1], p[2], t);
    bool keep = true;
    for (int a = 0; a < 3; ++a) {
        p[a] = __dsub_rn(p[a], c.mn[a]);
        keep = keep && p[a] >= c.lo[a] && p[a] <= c.hi[a];
    }
    return keep;
}

// per-tile kept count (tile = SLAB_TILE consecutive replica atoms) and the minimum of the kept coordinates
__global__ void __launch_bounds__(SLAB_THREADS)
slab_count_kernel(SlabGeom g, SlabCut c, int64_t *tile_count, unsigned long long *kept_min3)
{
    __shared__ int s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double lo[3] = {inf, inf, inf};
    const int64_t base = (int64_t)blockIdx.x * SLAB_TILE + (int64_t)threadIdx.x * SLAB_ITEMS;
    int n = 0;
    for (int u = 0; u < SLAB_ITEMS; ++u) {
        const int64_t idx = base + u;
        if (idx >= g.total) break;
        double p[3];
        int64_t t;
        if (slab_keep(g, c, idx, p, t)) {
            ++n;
            for (int a = 0; a < 3; ++a) lo[a] = fmin(lo[a], p[a]);
        }
    }
    for (int s = 16; s > 0; s >>= 1) n += __shfl_xor_sync(0xffffffffu, n, s);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&s_count, n);
    for (int a = 0; a < 3; ++a) slab_minmax_commit(lo[a], 0.0, kept_min3 + a, NULL);
    __syncthreads();
    if (threadIdx.x == 0) tile_count[blockIdx.x] = s_count;
}

// in-place exclusive scan of n int64 counts by one CTA; total -> *d_total
__global__ void __launch_bounds__(1024)
slab_scan_kernel(int64_t *v, int64_t n, int64_t *d_total)
{
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t x = i < n ? v[i] : 0;
        int64_t incl = x;
        for (int s = 1; s < 32; s <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= s) incl += y;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int64_t w = s_warp[lane], wi = w;
            for (int s = 1; s < 32; s <<= 1) {
                const int64_t y = __shfl_up_sync(0xffffffffu, wi, s);
                if (lane >= s) wi += y;
            }
            s_warp[lane] = wi - w;                 // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int64_t carry = s_carry;
        if (i < n) v[i] = carry + s_warp[warp] + incl - x;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[warp] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *d_total = s_carry;
}

__global__ void __launch_bounds__(SLAB_THREADS)
slab_write_kernel(SlabGeom g, SlabCut c, const int64_t *__restrict__ tile_offset, double *out_xyz, uint8_t *out_species)
{
    __shared__ int s_warp[SLAB_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SLAB_TILE + (int64_t)threadIdx.x * SLAB_ITEMS;
    double p[SLAB_ITEMS][3];
    int64_t t[SLAB_ITEMS];
    bool keep[SLAB_ITEMS];
    int n = 0;
    for (int u = 0; u < SLAB_ITEMS; ++u) {
        const int64_t idx = base + u;
        keep[u] = idx < g.total && slab_keep(g, c, idx, p[u], t[u]);
        n += keep[u] ? 1 : 0;
    }
    // exclusive scan of the per-thread counts over the CTA (thread order == atom order)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = n;
    for (int s = 1; s < 32; s <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    int64_t o = tile_offset[blockIdx.x] + before + incl - n;
    for (int u = 0; u < SLAB_ITEMS; ++u) {
        if (!keep[u]) continue;
        for (int a = 0; a < 3; ++a) out_xyz[3 * o + a] = __dsub_rn(p[u][a], c.kept_min[a]);
        if (out_species) out_species[o] = g.species[t[u]];
        ++o;
    }
}

static int slab_geom(const gx_slab_args *h, SlabGeom &g)
{
    GX_REQUIRE(h && h->d_cell_xyz, "NULL pointer");
    GX_REQUIRE(h->n_cell > 0 && h->nx > 0 && h->ny > 0 && h->nz > 0, "bad replica counts");
    g.cell = h->d_cell_xyz; g.species = h->d_cell_species;
    g.n0 = h->n_cell; g.nx = h->nx; g.ny = h->ny; g.nz = h->nz;
    g.total = (int64_t)h->n_cell * h->nx * h->ny * h->nz;
    GX_REQUIRE(g.total / SLAB_TILE < 2147483647LL, "too many replica atoms");
    g.ax = h->ax; g.bx = h->bx; g.by = h->by; g.cx = h->cx; g.cy = h->cy; g.cz = h->cz;
    return GX_OK;
}

static void slab_cut(const double *h_min3, const double *h_lo3, const double *h_hi3, const double *h_kept3, SlabCut &c)
{
    for (int a = 0; a < 3; ++a) {
        c.mn[a] = h_min3[a]; c.lo[a] = h_lo3[a]; c.hi[a] = h_hi3[a];
        c.kept_min[a] = h_kept3 ? h_kept3[a] : 0.0;
    }
}

extern "C" int64_t gx_slab_tiles(const gx_slab_args *h)
{
    if (!h) return 0;
    const int64_t total = (int64_t)h->n_cell * h->nx * h->ny * h->nz;
    return (total + SLAB_TILE - 1) / SLAB_TILE;
}

extern "C" int gx_slab_minmax(const gx_slab_args *h, double *d_out6, void *stream)
{
    SlabGeom g;
    if (int e = slab_geom(h, g)) return e;
    GX_REQUIRE(d_out6, "NULL pointer");
    cudaStream_t st = gx_stream(stream);
    unsigned long long *o = reinterpret_cast<unsigned long long *>(d_out6);
    slab_init_kernel<<<1, 32, 0, st>>>(o);
    int64_t blocks = (g.total + SLAB_THREADS - 1) / SLAB_THREADS;
    if (blocks > GX_SM_COUNT * 16) blocks = GX_SM_COUNT * 16;
    slab_minmax_kernel<<<(int)blocks, SLAB_THREADS, 0, st>>>(g, o);
    slab_decode_kernel<<<1, 32, 0, st>>>(o, 6);
    return gx_check_launch("gx_slab_minmax");
}

extern "C" int gx_slab_count(const gx_slab_args *h, const double *h_min3, const double *h_lo3, const double *h_hi3,
                             int64_t *d_tile_count, int64_t *d_total, double *d_kept_min3, void *stream)
{
    SlabGeom g;
    if (int e = slab_geom(h, g)) return e;
    GX_REQUIRE(h_min3 && h_lo3 && h_hi3 && d_tile_count && d_total && d_kept_min3, "NULL pointer");
    SlabCut c;
    slab_cut(h_min3, h_lo3, h_hi3, NULL, c);
    cudaStream_t st = gx_stream(stream);
    unsigned long long *km = reinterpret_cast<unsigned long long *>(d_kept_min3);
    GX_CUDA(cudaMemsetAsync(km, 0xff, 3 * sizeof(unsigned long long), st));
    const int64_t tiles = (g.total + SLAB_TILE - 1) / SLAB_TILE;
    slab_count_kernel<<<(int)tiles, SLAB_THREADS, 0, st>>>(g, c, d_tile_count, km);
    slab_scan_kernel<<<1, 1024, 0, st>>>(d_tile_count, tiles, d_total);
    slab_decode_kernel<<<1, 32, 0, st>>>(km, 3);
    return gx_check_launch("gx_slab_count");
}

extern "C" int gx_slab_write(const gx_slab_args *h, const double *h_min3, const double *h_lo3, const double *h_hi3,
                             const double *h_kept_min3, const int64_t *d_tile_offset, double *d_xyz_out,
                             uint8_t *d_species_out, void *stream)
{
    SlabGeom g;
    if (int e = slab_geom(h, g)) return e;
    GX_REQUIRE(h_min3 && h_lo3 && h_hi3 && h_kept_min3 && d_tile_offset && d_xyz_out, "NULL pointer");
    GX_REQUIRE(!d_species_out || h->d_cell_species, "species requested but the unit cell has none");
    SlabCut c;
    slab_cut(h_min3, h_lo3, h_hi3, h_kept_min3, c);
    const int64_t tiles = (g.total + SLAB_TILE - 1) / SLAB_TILE;
    slab_write_kernel<<<(int)tiles, SLAB_THREADS, 0, gx_stream(stream)>>>(g, c, d_tile_offset, d_xyz_out, d_species_out);
    return gx_check_launch("gx_slab_write");
}
