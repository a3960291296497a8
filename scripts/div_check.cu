// exhaustive-ish check of SharedDivisor against IEEE division on the GPU (run: nvcc ... && ./a.out)
#include <cstdio>
#include <cstdint>
#include "../instantavatar_b200/csrc/ia_device.cuh"
__global__ void k(unsigned long long* bad, unsigned long long* tot, uint32_t seed) {
    uint32_t x = seed + blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long nb = 0;
    for (int it = 0; it < 20000; it++) {
        x = x * 1664525u + 1013904223u; uint32_t ua = x; x = x * 1664525u + 1013904223u; uint32_t ub = x;
        float a = __uint_as_float(ua), b = __uint_as_float(ub);
        if (!isfinite(a) || !isfinite(b)) continue;
        // concentrate on the magnitudes the solver sees, but keep random mantissas/exponents too
        if (it & 1) { a = ldexpf(a, -(int)((ua >> 23) & 0xff) + 127 - (int)(ub % 40)); b = ldexpf(b, -(int)((ub >> 23) & 0xff) + 127 - (int)(ua % 40)); }
        ia::SharedDivisor d(b);
        float q1 = d.div(a), q2 = a / b;
        if (__float_as_uint(q1) != __float_as_uint(q2) && !(q1 != q1 && q2 != q2)) nb++;
    }
    atomicAdd(bad, nb); atomicAdd(tot, 20000ull);
}
int main() {
    unsigned long long *bad, *tot; cudaMallocManaged(&bad, 8); cudaMallocManaged(&tot, 8); *bad = 0; *tot = 0;
    k<<<2048, 256>>>(bad, tot, 12345u); cudaDeviceSynchronize();
    printf("divisions checked %llu mismatches %llu\n", *tot, *bad);
    return *bad != 0;
}
