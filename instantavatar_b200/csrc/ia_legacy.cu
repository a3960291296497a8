// ia_legacy.cu -- kernel-for-kernel replacements of the reference's raymarcher extension
// (instant_avatar/renderers/cuda/raymarcher.cpp:16-81): raymarch_test, raymarch_train, composite_test.
// They serve Raymarcher's legacy path (any `model(pts)` callable / foreign deformers); the fused kernels do not use
// them.  Same numerics contract as the rest of the library (-fmad=false, explicit fma): bit-exact with the oracle and
// with the reference's own kernels (tests/golden/ref_cuda_golden.npz).
#include <math.h>
#include <stdint.h>

#include "ia_host.h"

namespace {

__device__ __forceinline__ float clampf_(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }

// raymarcher.cu:116-161 ; depths [N][N_steps] must be zero-initialised by the caller (the reference uses at::zeros)
__global__ void __launch_bounds__(256) raymarch_train_kernel(int n_rays, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                             const float* __restrict__ nears, const float* __restrict__ fars,
                                                             const uint8_t* __restrict__ grid, int gs, const float* __restrict__ scale,
                                                             const float* __restrict__ offset, const float* __restrict__ step_size,
                                                             int N_steps, float* __restrict__ depths) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_rays) return;
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
    const float cx = offset[0], cy = offset[1], cz = offset[2];
    const float sx = (float)gs / scale[0], sy = (float)gs / scale[1], sz = (float)gs / scale[2];
    const float far = fars[n], dt = step_size[n];
    int s = 0;
    float t = nears[n];
    while (t < far && s < N_steps) {
        const float x = __fmaf_rn(t, dx, ox), y = __fmaf_rn(t, dy, oy), z = __fmaf_rn(t, dz, oz);
        const int nx = (int)clampf_((x - cx) * sx, 0.0f, (float)gs - 1.0f);
        const int ny = (int)clampf_((y - cy) * sy, 0.0f, (float)gs - 1.0f);
        const int nz = (int)clampf_((z - cz) * sz, 0.0f, (float)gs - 1.0f);
        if (grid[((long)nx * gs + ny) * gs + nz]) { depths[(long)n * N_steps + s] = t; s++; }
        t += dt;
    }
}

// raymarcher.cu:13-73 ; pts/deltas/depths zero-initialised by the caller; nears is updated in place (:72)
__global__ void __launch_bounds__(256) raymarch_test_kernel(int n_alive, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                            float* __restrict__ nears, const float* __restrict__ fars,
                                                            const int64_t* __restrict__ alive, const uint8_t* __restrict__ grid, int gs,
                                                            const float* __restrict__ scale, const float* __restrict__ offset,
                                                            const float* __restrict__ step_size, int N_steps, float* __restrict__ pts,
                                                            float* __restrict__ deltas, float* __restrict__ depths) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_alive) return;
    const long n = alive[i];
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
    const float cx = offset[0], cy = offset[1], cz = offset[2];
    const float sx = (float)gs / scale[0], sy = (float)gs / scale[1], sz = (float)gs / scale[2];
    const float far = fars[n], dt = step_size[n];
    int s = 0;
    float t = nears[n];
    while (t < far && s < N_steps) {
        const float x = __fmaf_rn(t, dx, ox), y = __fmaf_rn(t, dy, oy), z = __fmaf_rn(t, dz, oz);
        const int nx = (int)clampf_((x - cx) * sx, 0.0f, (float)gs - 1.0f);
        const int ny = (int)clampf_((y - cy) * sy, 0.0f, (float)gs - 1.0f);
        const int nz = (int)clampf_((z - cz) * sz, 0.0f, (float)gs - 1.0f);
        if (grid[((long)nx * gs + ny) * gs + nz]) {
            const long o = (long)i * N_steps + s;
            pts[o * 3] = x; pts[o * 3 + 1] = y; pts[o * 3 + 2] = z;
            deltas[o] = dt; depths[o] = t;
            s++;
        }
        t += dt;
    }
    nears[n] = t;
}

// raymarcher.cu:200-235 ; color/depth/nohit accumulated in place
__global__ void __launch_bounds__(256) composite_test_kernel(int n_alive, int N_steps, const float* __restrict__ rgb_vals,
                                                             const float* __restrict__ sigma_vals, const float* __restrict__ delta_vals,
                                                             const float* __restrict__ depth_vals, const int64_t* __restrict__ alive,
                                                             float* __restrict__ color, float* __restrict__ depth, float* __restrict__ nohit,
                                                             float thresh) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_alive) return;
    const long n = alive[i];
    float T = nohit[n];
    float c0 = color[n * 3], c1 = color[n * 3 + 1], c2 = color[n * 3 + 2], d = depth[n];
    int s = 0;
    while (s < N_steps && T > 1e-4f && delta_vals[(long)i * N_steps + s] > 0.f) {
        const long o = (long)i * N_steps + s;
        const float tau = expf(-sigma_vals[o] * delta_vals[o]);
        const float al = 1.0f - tau;
        if (al < thresh) { s++; continue; }
        const float w = al * T;
        c0 = __fmaf_rn(w, rgb_vals[o * 3], c0); c1 = __fmaf_rn(w, rgb_vals[o * 3 + 1], c1); c2 = __fmaf_rn(w, rgb_vals[o * 3 + 2], c2);
        d = __fmaf_rn(w, depth_vals[o], d);
        T *= tau;
        s++;
    }
    color[n * 3] = c0; color[n * 3 + 1] = c1; color[n * 3 + 2] = c2; depth[n] = d; nohit[n] = T;
}

}  // namespace

extern "C" {

int ia_raymarch_train(const float* rays_o, const float* rays_d, const float* nears, const float* fars, int n_rays,
                      const uint8_t* density_grid, int grid_size, const float* scale, const float* offset,
                      const float* step_size, int N_steps, float* depths, ia_stream_t stream) {
    IA_REQUIRE(n_rays >= 0 && N_steps > 0 && grid_size > 0);
    if (n_rays == 0) return IA_OK;
    IA_REQUIRE(rays_o && rays_d && nears && fars && density_grid && scale && offset && step_size && depths);
    raymarch_train_kernel<<<(n_rays + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n_rays, rays_o, rays_d, nears, fars, density_grid,
                                                                                   grid_size, scale, offset, step_size, N_steps, depths);
    IA_CHECK_CUDA(cudaPeekAtLastError());
    return IA_OK;
}

int ia_raymarch_test(const float* rays_o, const float* rays_d, float* nears, const float* fars, const int64_t* alive,
                     int n_alive, const uint8_t* density_grid, int grid_size, const float* scale, const float* offset,
                     const float* step_size, int N_steps, float* pts, float* deltas, float* depths, ia_stream_t stream) {
    IA_REQUIRE(n_alive >= 0 && N_steps > 0 && grid_size > 0);
    if (n_alive == 0) return IA_OK;
    IA_REQUIRE(rays_o && rays_d && nears && fars && alive && density_grid && scale && offset && step_size && pts && deltas && depths);
    raymarch_test_kernel<<<(n_alive + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n_alive, rays_o, rays_d, nears, fars, alive, density_grid,
                                                                                   grid_size, scale, offset, step_size, N_steps, pts, deltas, depths);
    IA_CHECK_CUDA(cudaPeekAtLastError());
    return IA_OK;
}

int ia_composite_test(const float* rgb_vals, const float* sigma_vals, const float* delta_vals, const float* depth_vals,
                      const int64_t* alive, int n_alive, int N_steps, float* color, float* depth, float* no_hit, float thresh,
                      ia_stream_t stream) {
    IA_REQUIRE(n_alive >= 0 && N_steps >= 0);
    if (n_alive == 0 || N_steps == 0) return IA_OK;
    IA_REQUIRE(rgb_vals && sigma_vals && delta_vals && depth_vals && alive && color && depth && no_hit);
    composite_test_kernel<<<(n_alive + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n_alive, N_steps, rgb_vals, sigma_vals, delta_vals,
                                                                                    depth_vals, alive, color, depth, no_hit, thresh);
    IA_CHECK_CUDA(cudaPeekAtLastError());
    return IA_OK;
}

}  // extern "C"
