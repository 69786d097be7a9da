// dependent-call latency of the short transcendental functions of jaero_b200/csrc/demod_device.cuh next to the library calls
#include <cstdio>
#include <cmath>
#include "../../jaero_b200/csrc/demod_device.cuh"
template <int W> __global__ void k(double *io, int n, long long *cyc)
{
    double y = io[threadIdx.x], x = io[32 + threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        if (W == 0) y = atan2(y + 0.37, x);
        if (W == 1) y = jb::atan2_fast(y + 0.37, x);
        if (W == 2) y = y / (x + y * 1e-9) * 1.0000001 + 0.25;
        if (W == 3) y = jb::div_fast(y, x + y * 1e-9) * 1.0000001 + 0.25;
        if (W == 4) y = jb::div_exact(y, 360.0, 1.0 / 360.0) + x;
        if (W == 5) y = (double)((int)(y * 1.7)) + x * 0.5;           // F2I + I2F round trip (osc_index)
    }
    long long t1 = clock64();
    io[threadIdx.x] = y;
    if (threadIdx.x == 0) *cyc = (t1 - t0) / n;
}
int main()
{
    double *io; long long *c; cudaMallocManaged(&io, 64 * 8); cudaMallocManaged(&c, 8);
    const char *names[] = {"atan2 (library)", "atan2_fast", "division (IEEE)", "div_fast", "div_exact", "f2i+i2f"};
    for (int w = 0; w < 6; w++) {
        for (int i = 0; i < 64; i++) io[i] = 0.3 + 0.01 * i;
        if (w == 0) k<0><<<1, 32>>>(io, 20000, c); if (w == 1) k<1><<<1, 32>>>(io, 20000, c); if (w == 2) k<2><<<1, 32>>>(io, 20000, c);
        if (w == 3) k<3><<<1, 32>>>(io, 20000, c); if (w == 4) k<4><<<1, 32>>>(io, 20000, c); if (w == 5) k<5><<<1, 32>>>(io, 20000, c);
        cudaDeviceSynchronize();
        printf("%-18s %lld cycles per dependent call\n", names[w], *c);
    }
    return 0;
}
