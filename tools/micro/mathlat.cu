// Micro-benchmark: dependent-chain latency (cycles per call) of the double-precision libm calls on the K1a critical loop.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void chain(int n, long long *out, double *sink, double seed)
{
    double v = seed + threadIdx.x * 1e-6, w = 0.37 + threadIdx.x * 1e-7;
    long long t0 = clock64();
    for (int j = 0; j < n; j++) {
        if (OP == 0) v = atan2(v, w) + 0.3;
        if (OP == 1) v = tanh(v) + 0.2;
        if (OP == 2) v = hypot(v, w);
        if (OP == 3) v = sqrt(v) + 0.5;
        if (OP == 4) v = 1.414213562 / v + 0.4;
        if (OP == 5) v = cos(v) + sin(v) * 1e-3;
        if (OP == 6) v = log10(v + 2.0);
        if (OP == 7) v = fmod(v * 400.0, 360.0) * 0.01 + 0.1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = v;
}
int main()
{
    long long *d; double *s; long long h; cudaMalloc(&d, 8); cudaMalloc(&s, 32 * 8);
    const int n = 20000; const char *names[] = {"atan2", "tanh", "hypot", "sqrt", "div", "sincos", "log10", "fmod"};
#define RUN(k) chain<k><<<1, 32>>>(n, d, s, 0.7); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("%-7s %.0f cycles per dependent call\n", names[k], (double)h / n);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
