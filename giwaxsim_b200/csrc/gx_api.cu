// Library-level entry points: version, error reporting, device check.
#include <stdarg.h>
#include <string.h>
#include "gx_common.cuh"

static thread_local char g_err[512] = "";

void gx_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int gx_check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        gx_set_error("%s: launch failed -> %s", what, cudaGetErrorString(e));
        return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? GX_ERR_NO_DEVICE : GX_ERR_CUDA;
    }
    return GX_OK;
}

extern "C" int gx_abi_version(void) { return GX_ABI_VERSION; }

extern "C" const char *gx_last_error(void) { return g_err; }

__global__ void gx_probe_kernel(int *out) { if (threadIdx.x == 0) *out = 100; }

extern "C" int gx_device_check(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        gx_set_error("gx_device_check: no CUDA device visible (%s); this library has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return GX_ERR_NO_DEVICE;
    }
    GX_REQUIRE(device >= 0 && device < n, "device index out of range");
    cudaDeviceProp p;
    GX_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) {
        gx_set_error("gx_device_check: device %d is sm_%d%d; this library carries an sm_100a image only",
                     device, p.major, p.minor);
        return GX_ERR_NO_DEVICE;
    }
    int cur = 0;
    GX_CUDA(cudaGetDevice(&cur));
    GX_CUDA(cudaSetDevice(device));
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, gx_probe_kernel);
    cudaSetDevice(cur);
    if (e != cudaSuccess) {
        gx_set_error("gx_device_check: sm_100a kernel image not loadable on device %d (%s)", device,
                     cudaGetErrorString(e));
        cudaGetLastError();
        return GX_ERR_NO_DEVICE;
    }
    return GX_OK;
}

// ------------------------------------------------------- host boundary ----
// Page-lock a range of host memory the caller owns (a slab of a pooled POSIX shared-memory result
// segment) so that results can be DMA'd straight into their final place, and the asynchronous copy
// itself.  Registration is expensive (~0.2 ms per MB): callers register pooled buffers once.
extern "C" int gx_host_register(void *h_ptr, int64_t nbytes)
{
    GX_REQUIRE(h_ptr && nbytes > 0, "bad arguments");
    cudaError_t e = cudaHostRegister(h_ptr, (size_t)nbytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return GX_OK; }
    if (e != cudaSuccess) {
        gx_set_error("gx_host_register: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return GX_ERR_CUDA;
    }
    return GX_OK;
}

extern "C" int gx_host_unregister(void *h_ptr)
{
    if (!h_ptr) return GX_OK;
    cudaError_t e = cudaHostUnregister(h_ptr);
    if (e != cudaSuccess) cudaGetLastError();
    return GX_OK;
}

extern "C" int gx_copy_to_host_async(void *h_dst, const void *d_src, int64_t nbytes, void *stream)
{
    GX_REQUIRE(h_dst && d_src && nbytes >= 0, "bad arguments");
    GX_CUDA(cudaMemcpyAsync(h_dst, d_src, (size_t)nbytes, cudaMemcpyDeviceToHost, gx_stream(stream)));
    return GX_OK;
}
