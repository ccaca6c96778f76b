/* TEST INFRASTRUCTURE ONLY -- rounding-exact restatements of the three BLAS
 * calls the GIWAXSim reference makes on the hot path, so that the oracle does
 * not depend on which OpenBLAS micro-kernel the host CPU selects.
 *
 *   rotz_y      np.dot(coords, Rz.T)[:,1]        tools/utilities.py:315
 *   matvec3     R(3x3) @ vstack(x,y,z)(3xn)      tools/detector.py:71,113,155
 *   dot3        x.dot(x) inside np.linalg.norm   tools/detector.py:65,107,149
 *
 * Each is the fused-multiply-add chain k = 0,1,2 that OpenBLAS' dgemm/ddot
 * kernels evaluate (first product rounded, then two fma).  That this equals
 * the reference's NumPy result bit for bit is pinned by
 * tests/test_oracle_vs_reference.py (live reference) and by the golden
 * fixtures under tests/golden/.  Compiled with -ffp-contract=off so that only
 * the explicit fma() calls fuse.
 */
#include <math.h>
#include <stddef.h>

void ox_rotz_y(const double *x, const double *y, size_t n, double s, double c, double *out)
{
    for (size_t i = 0; i < n; ++i) {
        double t = x[i] * s;
        out[i] = fma(y[i], c, t);
    }
}

/* out[r][i] = fma(R[r][2], z[i], fma(R[r][1], y[i], R[r][0]*x[i])) */
void ox_matvec3(const double *R, const double *x, const double *y, const double *z,
                size_t n, double *ox, double *oy, double *oz)
{
    for (size_t i = 0; i < n; ++i) {
        double a = x[i], b = y[i], c = z[i];
        double t0 = R[0] * a; t0 = fma(R[1], b, t0); t0 = fma(R[2], c, t0);
        double t1 = R[3] * a; t1 = fma(R[4], b, t1); t1 = fma(R[5], c, t1);
        double t2 = R[6] * a; t2 = fma(R[7], b, t2); t2 = fma(R[8], c, t2);
        ox[i] = t0; oy[i] = t1; oz[i] = t2;
    }
}

double ox_dot3_fma(const double *v)
{
    double t = v[0] * v[0];
    t = fma(v[1], v[1], t);
    t = fma(v[2], v[2], t);
    return t;
}

double ox_dot3_plain(const double *v)
{
    double t = v[0] * v[0];
    double u = v[1] * v[1];
    t = t + u;
    u = v[2] * v[2];
    return t + u;
}
