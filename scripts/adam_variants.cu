// adam_variants.cu -- why does the dense Adam pass (13 M parameters, 443 MB of HBM traffic) run at 2.2 TB/s, and which
// kernel shape reaches the copy bandwidth?  Variants of adam_dev_kernel (csrc/ia_train.cu) timed with CUDA events after an
// L2 flush; prints one JSON line per variant.   Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o
// scripts/adam_variants scripts/adam_variants.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ void upd(float& pp, float gg, float& mm, float& vv, float inv, float beta1, float beta2, float step_size,
                                    float inv_bc2, float eps) {
    const float gi = gg * inv;
    mm = __fmaf_rn(beta1, mm, (1.f - beta1) * gi);
    vv = __fmaf_rn(beta2, vv, (1.f - beta2) * gi * gi);
    float sq;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(vv));
    pp = __fmaf_rn(-step_size, __fdividef(mm, __fmaf_rn(sq, inv_bc2, eps)), pp);
}

// V0: the product kernel's shape (m, v loaded after the found_inf / state loads resolve)
__global__ void __launch_bounds__(256) adam_v0(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                                               long n4, const float* __restrict__ state, const float* __restrict__ found_inf, __half2* __restrict__ half_out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 gr = g[i];
    g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool skip = found_inf && *found_inf != 0.f;
    float4 pi = p[i];
    if (!skip) {
        const float lr = state[0], beta1 = state[1], beta2 = state[2], eps = state[3], bc1 = state[5], bc2_sqrt = state[6], inv = state[7];
        float4 mi = m[i], vi = v[i];
        const float step_size = lr / bc1, inv_bc2 = 1.0f / bc2_sqrt;
        upd(pi.x, gr.x, mi.x, vi.x, inv, beta1, beta2, step_size, inv_bc2, eps); upd(pi.y, gr.y, mi.y, vi.y, inv, beta1, beta2, step_size, inv_bc2, eps);
        upd(pi.z, gr.z, mi.z, vi.z, inv, beta1, beta2, step_size, inv_bc2, eps); upd(pi.w, gr.w, mi.w, vi.w, inv, beta1, beta2, step_size, inv_bc2, eps);
        m[i] = mi; v[i] = vi; p[i] = pi;
    }
    half_out[2 * i] = __floats2half2_rn(pi.x, pi.y);
    half_out[2 * i + 1] = __floats2half2_rn(pi.z, pi.w);
}

// V1: all four loads up front (no dependence on the flag), state in shared memory, streaming (evict-first) loads / stores,
//     U float4 per thread, grid-stride over a persistent grid
template <int U, bool kStream>
__global__ void __launch_bounds__(256) adam_v1(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                                               long n4, const float* __restrict__ state, const float* __restrict__ found_inf, uint2* __restrict__ half_out) {
    __shared__ float st[8];
    __shared__ float fi;
    if (threadIdx.x < 8) st[threadIdx.x] = state[threadIdx.x];
    if (threadIdx.x == 8) fi = found_inf ? *found_inf : 0.f;
    __syncthreads();
    const bool skip = fi != 0.f;
    const float lr = st[0], beta1 = st[1], beta2 = st[2], eps = st[3], bc1 = st[5], bc2_sqrt = st[6], inv = st[7];
    const float step_size = lr / bc1, inv_bc2 = 1.0f / bc2_sqrt;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long base = (long)blockIdx.x * blockDim.x + threadIdx.x; base < n4; base += stride * U) {
        float4 gr[U], pi[U], mi[U], vi[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long i = base + u * stride;
            if (i < n4) {
                if (kStream) { gr[u] = __ldcs(g + i); pi[u] = __ldcs(p + i); mi[u] = __ldcs(m + i); vi[u] = __ldcs(v + i); }
                else { gr[u] = g[i]; pi[u] = p[i]; mi[u] = m[i]; vi[u] = v[i]; }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long i = base + u * stride;
            if (i >= n4) continue;
            if (!skip) {
                upd(pi[u].x, gr[u].x, mi[u].x, vi[u].x, inv, beta1, beta2, step_size, inv_bc2, eps); upd(pi[u].y, gr[u].y, mi[u].y, vi[u].y, inv, beta1, beta2, step_size, inv_bc2, eps);
                upd(pi[u].z, gr[u].z, mi[u].z, vi[u].z, inv, beta1, beta2, step_size, inv_bc2, eps); upd(pi[u].w, gr[u].w, mi[u].w, vi[u].w, inv, beta1, beta2, step_size, inv_bc2, eps);
            }
            const __half2 h0 = __floats2half2_rn(pi[u].x, pi[u].y), h1 = __floats2half2_rn(pi[u].z, pi[u].w);
            const uint2 hh = make_uint2(*reinterpret_cast<const unsigned*>(&h0), *reinterpret_cast<const unsigned*>(&h1));
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kStream) {
                __stcs(g + i, z);
                if (!skip) { __stcs(m + i, mi[u]); __stcs(v + i, vi[u]); __stcs(p + i, pi[u]); }
                __stcs(half_out + i, hh);
            } else {
                g[i] = z;
                if (!skip) { m[i] = mi[u]; v[i] = vi[u]; p[i] = pi[u]; }
                half_out[i] = hh;
            }
        }
    }
}

__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, long n4) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) b[i] = a[i];
}

int main() {
    const long n = 13036208, n4 = n / 4;
    float *p, *g, *m, *v, *state, *found, *flush;
    __half2* h;
    CK(cudaMalloc(&p, n * 4)); CK(cudaMalloc(&g, n * 4)); CK(cudaMalloc(&m, n * 4)); CK(cudaMalloc(&v, n * 4)); CK(cudaMalloc(&h, n * 2));
    CK(cudaMalloc(&state, 32)); CK(cudaMalloc(&found, 4)); CK(cudaMalloc(&flush, 256 << 20));
    CK(cudaMemset(p, 0, n * 4)); CK(cudaMemset(g, 0, n * 4)); CK(cudaMemset(m, 0, n * 4)); CK(cudaMemset(v, 0, n * 4)); CK(cudaMemset(found, 0, 4));
    const float hs[8] = {1e-2f, 0.9f, 0.99f, 1e-15f, 3.f, 0.271f, 0.1726f, 1.f / 1024.f};
    CK(cudaMemcpy(state, hs, 32, cudaMemcpyHostToDevice));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double bytes = (double)n * 34.0;
    auto time = [&](const char* name, auto launch, double nbytes) {
        float best = 1e9f;
        for (int r = 0; r < 5; r++) {
            CK(cudaMemsetAsync(flush, r, 256 << 20));
            CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("{\"variant\": \"%s\", \"ms\": %.4f, \"GBps\": %.1f}\n", name, best, nbytes / (best * 1e-3) / 1e9); fflush(stdout);
    };
    time("copy_52MB_x4 (reference: 8 streams)", [&] { copy_kernel<<<sms * 8, 256>>>((float4*)p, (float4*)g, n4); copy_kernel<<<sms * 8, 256>>>((float4*)m, (float4*)v, n4); }, (double)n * 16.0);
    time("v0_product_shape", [&] { adam_v0<<<(unsigned)((n4 + 255) / 256), 256>>>((float4*)p, (float4*)g, (float4*)m, (float4*)v, n4, state, found, h); }, bytes);
    time("v1_U1_plain_flatgrid", [&] { adam_v1<1, false><<<(unsigned)((n4 + 255) / 256), 256>>>((float4*)p, (float4*)g, (float4*)m, (float4*)v, n4, state, found, (uint2*)h); }, bytes);
    time("v1_U1_stream_flatgrid", [&] { adam_v1<1, true><<<(unsigned)((n4 + 255) / 256), 256>>>((float4*)p, (float4*)g, (float4*)m, (float4*)v, n4, state, found, (uint2*)h); }, bytes);
    time("v1_U2_stream_persist8", [&] { adam_v1<2, true><<<sms * 8, 256>>>((float4*)p, (float4*)g, (float4*)m, (float4*)v, n4, state, found, (uint2*)h); }, bytes);
    time("v1_U4_stream_persist8", [&] { adam_v1<4, true><<<sms * 8, 256>>>((float4*)p, (float4*)g, (float4*)m, (float4*)v, n4, state, found, (uint2*)h); }, bytes);
    time("v1_U2_plain_persist8", [&] { adam_v1<2, false><<<sms * 8, 256>>>((float4*)p, (float4*)g, (float4*)m, (float4*)v, n4, state, found, (uint2*)h); }, bytes);
    time("v1_U4_stream_persist4", [&] { adam_v1<4, true><<<sms * 4, 256>>>((float4*)p, (float4*)g, (float4*)m, (float4*)v, n4, state, found, (uint2*)h); }, bytes);
    time("v1_U2_stream_persist16", [&] { adam_v1<2, true><<<sms * 16, 256>>>((float4*)p, (float4*)g, (float4*)m, (float4*)v, n4, state, found, (uint2*)h); }, bytes);
    return 0;
}
