// Shared device/host helpers for the giwaxs_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/giwaxs_b200.h"

#ifndef GX_SM_COUNT
#define GX_SM_COUNT 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this
#endif

void gx_set_error(const char *fmt, ...);
int gx_check_launch(const char *what);

#define GX_REQUIRE(cond, msg)                                   \
    do {                                                        \
        if (!(cond)) {                                          \
            gx_set_error("%s: %s", __func__, msg);              \
            return GX_ERR_INVALID;                              \
        }                                                       \
    } while (0)

#define GX_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) {                                                   \
            gx_set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e_)); \
            return (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver)  \
                       ? GX_ERR_NO_DEVICE : GX_ERR_CUDA;                           \
        }                                                                          \
    } while (0)

static inline cudaStream_t gx_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------
// NumPy's float `a // b` for b > 0: the exact floor of the true quotient
// (npy_divmod uses an exact fmod).  q0 is within 1 of it; the fma residual is
// exact for the right q and has the right sign for its neighbours.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double gx_floordiv(double a, double b, double inv_b)
{
    double q = floor(__dmul_rn(a, inv_b));
    double r = __fma_rn(-q, b, a);
    if (r < 0.0) q -= 1.0;
    else if (r >= b) q += 1.0;
    return q;
}

// y' of np.dot(coords, Rz.T): first product rounded, second fused.
__device__ __forceinline__ double gx_rot_y(double x, double y, double s, double c)
{
    return __fma_rn(y, c, __dmul_rn(x, s));
}

// np.linspace(start, stop, n)[j] given step = (stop-start)/(n-1) from the host.
__device__ __forceinline__ double gx_linspace(int j, int n, double start, double step, double stop)
{
    if (j == n - 1 && n > 1) return stop;
    return __dadd_rn(__dmul_rn((double)j, step), start);
}

// order-preserving map double <-> uint64 for atomicMin / atomicMax
__device__ __forceinline__ unsigned long long gx_ord(double v)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double gx_unord(unsigned long long o)
{
    unsigned long long b = (o & 0x8000000000000000ull) ? (o & 0x7fffffffffffffffull) : ~o;
    return __longlong_as_double((long long)b);
}

static inline unsigned long long gx_ord_host(double v)
{
    unsigned long long b;
    memcpy(&b, &v, 8);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
