// cos_model_hook.cu — device-gradient hook for the user model of the reference's own SVGD test
// (tests/test_svgd.cpp:77-92):  p(x) = a cos(x0) + b cos(x1) + c x0 x1 + d,  grad log p = grad p / p.
// Built into tests/cuda/_build/libcos_hook.so by tests/helpers.py:build_cos_hook(); registered through
// svgdb_set_model_device_hook (include/svgd_b200.h), the replacement for CppAD-taped Model lambdas.
#include <cuda_runtime.h>
#include <stdint.h>

struct CosParams { double a, b, c, d; };

__global__ void cos_model_grad_kernel(const double *X, double *G, int64_t row0, int64_t n_rows, CosParams p)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const double x0 = X[(row0 + r) * 2], x1 = X[(row0 + r) * 2 + 1];
    const double den = p.a * cos(x0) + p.b * cos(x1) + p.c * x0 * x1 + p.d;
    G[r * 2] = (-p.a * sin(x0) + p.c * x1) / den;
    G[r * 2 + 1] = (-p.b * sin(x1) + p.c * x0) / den;
}

extern "C" int cos_model_grad(const double *X_dev, double *G_dev, int64_t n_total, int32_t d, int64_t row0, int64_t n_rows,
                              void *cuda_stream, void *user)
{
    (void)n_total;
    if (d != 2 || !user) return 1;
    if (n_rows <= 0) return 0;
    const CosParams p = *static_cast<const CosParams *>(user); // host memory, copied by value into the launch
    cos_model_grad_kernel<<<(unsigned)((n_rows + 127) / 128), 128, 0, static_cast<cudaStream_t>(cuda_stream)>>>(X_dev, G_dev, row0, n_rows, p);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}
