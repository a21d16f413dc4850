// Accuracy of the fp64 MUFU seeds (rcp.approx.ftz.f64 / rsqrt.approx.ftz.f64) and of Newton refinements.
#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>
#include <vector>
#include <random>
__global__ void k(const double* x, double* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = x[i], y, z;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(v));
    out[8 * i + 0] = y;
    double e = fma(-v, y, 1.0); double y1 = fma(y, e, y);
    out[8 * i + 1] = y1;
    e = fma(-v, y1, 1.0); double y2 = fma(y1, e, y1);
    out[8 * i + 2] = y2;
    // cubic then quadratic
    e = fma(-v, y, 1.0); double yc = fma(y, fma(e, e, e), y);
    e = fma(-v, yc, 1.0); yc = fma(yc, e, yc);
    out[8 * i + 3] = yc;
    out[8 * i + 4] = z;
    double hx = 0.5 * v;
    e = fma(-hx * z, z, 0.5); double z1 = fma(z, e, z);
    out[8 * i + 5] = z1;
    e = fma(-hx * z1, z1, 0.5); double z2 = fma(z1, e, z1);
    out[8 * i + 6] = z2;
    // cubic then quadratic
    e = fma(-v * z, z, 1.0); double zc = fma(z * e, fma(0.375, e, 0.5), z);
    e = fma(-v * zc, zc, 1.0); zc = fma(0.5 * zc, e, zc);
    out[8 * i + 7] = zc;
}
int main() {
    const int n = 1 << 20;
    std::vector<double> h(n);
    std::mt19937_64 g(1);
    std::uniform_real_distribution<double> u(-12, 12);
    for (auto& v : h) v = pow(10.0, u(g)) * (1.0 + 1e-3 * u(g));
    double *dx, *dout;
    cudaMalloc(&dx, n * 8); cudaMalloc(&dout, n * 64);
    cudaMemcpy(dx, h.data(), n * 8, cudaMemcpyHostToDevice);
    k<<<n / 256, 256>>>(dx, dout, n);
    std::vector<double> o(8 * (size_t)n);
    cudaMemcpy(o.data(), dout, n * 64, cudaMemcpyDeviceToHost);
    double worst[8] = {0};
    for (int i = 0; i < n; i++)
        for (int c = 0; c < 8; c++) {
            long double ref = c < 4 ? 1.0L / (long double)h[i] : 1.0L / sqrtl((long double)h[i]);
            double err = (double)fabsl(((long double)o[8 * (size_t)i + c] - ref) / ref);
            if (err > worst[c]) worst[c] = err;
        }
    const char* names[8] = {"rcp seed", "rcp +1 newton", "rcp +2 newton", "rcp cubic+quad", "rsqrt seed", "rsqrt +1 newton", "rsqrt +2 newton", "rsqrt cubic+quad"};
    for (int c = 0; c < 8; c++) printf("%-18s max rel err %.3e (2^%.1f)\n", names[c], worst[c], log2(worst[c]));
    return 0;
}
