// gather_ceiling.cu -- what the memory system of one B200 can deliver for the access shape of the fused kernels:
// every lane gathers its own trilinear footprint (4 x-pair records of 96 B = 12 sectors) from an L2-resident 50 MB
// table, the next address depending on the loaded data (a Broyden iterate).  Variants measure other fetch shapes
// (quad-cooperative lines, per-lane bulk copies and tensor-map boxes through the TMA engine, shared-memory reads) so
// that the roofline of deform_query_kernel / render_fwd_kernel is a measured ceiling, not a guess.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o scripts/gather_ceiling scripts/gather_ceiling.cu -lcuda
// Run:    scripts/gather_ceiling            (prints one JSON line per variant)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int D = 32, H = 128, W = 128;          // voxel field of the avatar (deformer_torch.py:134-135)
constexpr int kRecFloats = 24;                    // x-pair record: 96 B
constexpr unsigned kFull = 0xffffffffu;

struct __align__(32) F8 { float v[8]; };
__device__ __forceinline__ F8 ldg256(const float* p) {
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_box4(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra.uni WD;\nbra.uni WL;\nWD:\n}\n"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ uint32_t lcg(uint32_t s) { return s * 1664525u + 1013904223u; }

enum Mode {
    LANE_LDG256 = 0,   // the fused kernels' shape: lane = sample, 4 records x 3 LDG.256
    LANE_LDG128 = 1,   // same with 6 LDG.128 per record
    QUAD_LINE = 2,     // 4 lanes read the 4 sectors of one 128-byte padded record (8 records per instruction)
    LANE_BULK96 = 3,   // lane = sample, 4 x cp.async.bulk of 96 B into shared memory, LDS.128 read-back
    LANE_TMABOX = 4,   // lane = sample, one cp.async.bulk.tensor box (12 floats x 2 x 2 x 2 = 384 B), LDS read-back
    SMEM_ONLY = 5,     // only the shared-memory read-back of 384 B per lane (conflict-free layout)
    LANE_SECTOR = 6,   // lane = sample, ONE LDG.256 per iteration (pure sector rate, 1 sector per lane)
    OCT_LINE = 7,      // 8 lanes read one 128-byte record with LDG.128 (4 records per instruction)
};

struct Args {
    const float* table;      // records of rec_stride floats
    int rec_stride;          // floats per record (24 or 32)
    int iters;               // footprints per lane
    int coherent;            // 1: lanes of a warp stay within a small voxel neighbourhood (as the kernels' batches do)
    float* sink;
    const CUtensorMap* map;
    unsigned long long* sectors;  // requested sectors (counter)
};

// footprint -> 4 record indices (y, z neighbours), as sample_field12 forms them
__device__ __forceinline__ void footprint(uint32_t rnd, uint32_t warp_rnd, int lane, int coherent, unsigned rec[4], int& x, int& y, int& z) {
    if (coherent) {
        // warp-level centre + a lane offset inside a 10 x 3 x 3 voxel neighbourhood (6 grid cells of 3.3 cm x 5 jitters)
        const int cx = 8 + (warp_rnd % (W - 24)), cy = 4 + ((warp_rnd >> 8) % (H - 12)), cz = 2 + ((warp_rnd >> 16) % (D - 8));
        x = cx + (rnd % 10); y = cy + ((rnd >> 8) % 3); z = cz + ((rnd >> 16) % 3);
    } else {
        x = rnd % (W - 1); y = (rnd >> 8) % (H - 1); z = (rnd >> 16) % (D - 1);
    }
    rec[0] = (z * H + y) * W + x; rec[1] = (z * H + y + 1) * W + x; rec[2] = ((z + 1) * H + y) * W + x; rec[3] = ((z + 1) * H + y + 1) * W + x;
}

template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) gather_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t s = (blockIdx.x * WARPS + warp) * 32 + lane + 12345u;
    uint32_t ws = (blockIdx.x * WARPS + warp) * 7919u + 17u;
    float acc[12];
#pragma unroll
    for (int c = 0; c < 12; c++) acc[c] = 0.f;
    unsigned long long nsect = 0;
    // per-warp staging for the TMA-engine variants: lane regions of 400 B (bulk) / 384 B (tensor box, 128-B aligned)
    constexpr int kLaneBytes = MODE == LANE_TMABOX ? 384 : 400;
    unsigned char* wbase = smem + 128 + (size_t)warp * 32 * kLaneBytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem) + warp;
    if constexpr (MODE == LANE_BULK96 || MODE == LANE_TMABOX) {
        if (lane == 0) mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    if constexpr (MODE == SMEM_ONLY) {
        for (int i = lane; i < 32 * 100; i += 32) reinterpret_cast<float*>(wbase)[i] = (float)i * 1e-9f;
        __syncwarp();
    }
    uint32_t phase = 0;
    for (int it = 0; it < a.iters; it++) {
        s = lcg(s); ws = lcg(ws);
        const uint32_t wr = __shfl_sync(kFull, ws, 0);
        unsigned rec[4]; int x, y, z;
        footprint(s >> 4, wr >> 4, lane, a.coherent, rec, x, y, z);
        const float w0 = 0.25f + (float)(s & 15) * 1e-3f;
        if constexpr (MODE == LANE_LDG256) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float* p = a.table + (size_t)rec[k] * a.rec_stride;
                const F8 A = ldg256(p), B = ldg256(p + 8), C = ldg256(p + 16);
#pragma unroll
                for (int c = 0; c < 8; c++) acc[c] = fmaf(A.v[c], w0, acc[c]);
#pragma unroll
                for (int c = 0; c < 4; c++) acc[8 + c] = fmaf(B.v[c], w0, acc[8 + c]);
#pragma unroll
                for (int c = 0; c < 4; c++) acc[c] = fmaf(B.v[4 + c], w0, acc[c]);
#pragma unroll
                for (int c = 0; c < 8; c++) acc[4 + c] = fmaf(C.v[c], w0, acc[4 + c]);
            }
            nsect += 12;
        } else if constexpr (MODE == LANE_LDG128) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float4* p = reinterpret_cast<const float4*>(a.table + (size_t)rec[k] * a.rec_stride);
#pragma unroll
                for (int j = 0; j < 6; j++) {
                    const float4 v = __ldg(p + j);
                    acc[(4 * j) % 12] = fmaf(v.x, w0, acc[(4 * j) % 12]); acc[(4 * j + 1) % 12] = fmaf(v.y, w0, acc[(4 * j + 1) % 12]);
                    acc[(4 * j + 2) % 12] = fmaf(v.z, w0, acc[(4 * j + 2) % 12]); acc[(4 * j + 3) % 12] = fmaf(v.w, w0, acc[(4 * j + 3) % 12]);
                }
            }
            nsect += 12;
        } else if constexpr (MODE == LANE_SECTOR) {
            const F8 A = ldg256(a.table + (size_t)rec[0] * a.rec_stride);
#pragma unroll
            for (int c = 0; c < 8; c++) acc[c] = fmaf(A.v[c], w0, acc[c]);
            nsect += 1;
        } else if constexpr (MODE == QUAD_LINE) {
            // the 32 footprints of the warp = 128 records; instruction i serves records of samples 8*(i/4) .. +7
#pragma unroll 4
            for (int i = 0; i < 16; i++) {
                const int src = 8 * (i >> 2) + (lane >> 2);
                const unsigned r0 = __shfl_sync(kFull, rec[0], src), r1 = __shfl_sync(kFull, rec[1], src);
                const unsigned r2 = __shfl_sync(kFull, rec[2], src), r3 = __shfl_sync(kFull, rec[3], src);
                const unsigned r = (i & 3) == 0 ? r0 : ((i & 3) == 1 ? r1 : ((i & 3) == 2 ? r2 : r3));
                const F8 A = ldg256(a.table + (size_t)r * a.rec_stride + 8 * (lane & 3));
#pragma unroll
                for (int c = 0; c < 8; c++) acc[c] = fmaf(A.v[c], w0, acc[c]);
            }
            nsect += 16;  // per lane: 16 sectors requested (4 per record incl. the pad sector), 12 useful per sample
        } else if constexpr (MODE == OCT_LINE) {
#pragma unroll 4
            for (int i = 0; i < 32; i++) {
                const int src = 4 * (i >> 2) + (lane >> 3);
                const unsigned r0 = __shfl_sync(kFull, rec[0], src), r1 = __shfl_sync(kFull, rec[1], src);
                const unsigned r2 = __shfl_sync(kFull, rec[2], src), r3 = __shfl_sync(kFull, rec[3], src);
                const unsigned r = (i & 3) == 0 ? r0 : ((i & 3) == 1 ? r1 : ((i & 3) == 2 ? r2 : r3));
                const float4 v = __ldg(reinterpret_cast<const float4*>(a.table + (size_t)r * a.rec_stride) + (lane & 7));
                acc[0] = fmaf(v.x, w0, acc[0]); acc[1] = fmaf(v.y, w0, acc[1]); acc[2] = fmaf(v.z, w0, acc[2]); acc[3] = fmaf(v.w, w0, acc[3]);
            }
            nsect += 16;
        } else if constexpr (MODE == LANE_BULK96 || MODE == LANE_TMABOX) {
            if (lane == 0) mbar_expect_tx(bar, 32 * 384);
            __syncwarp();
            unsigned char* mine = wbase + lane * kLaneBytes;
            if constexpr (MODE == LANE_BULK96) {
#pragma unroll
                for (int k = 0; k < 4; k++) bulk_g2s(mine + 96 * k, a.table + (size_t)rec[k] * a.rec_stride, 96, bar);
            } else {
                tma_box4(mine, a.map, 0, x, y, z, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            const float4* m4 = reinterpret_cast<const float4*>(mine);
#pragma unroll
            for (int j = 0; j < 24; j++) {
                const float4 v = m4[j];
                acc[(4 * j) % 12] = fmaf(v.x, w0, acc[(4 * j) % 12]); acc[(4 * j + 1) % 12] = fmaf(v.y, w0, acc[(4 * j + 1) % 12]);
                acc[(4 * j + 2) % 12] = fmaf(v.z, w0, acc[(4 * j + 2) % 12]); acc[(4 * j + 3) % 12] = fmaf(v.w, w0, acc[(4 * j + 3) % 12]);
            }
            __syncwarp();
            nsect += 12;
        } else if constexpr (MODE == SMEM_ONLY) {
            const float4* m4 = reinterpret_cast<const float4*>(wbase + lane * kLaneBytes);
#pragma unroll
            for (int j = 0; j < 24; j++) {
                const float4 v = m4[j];
                acc[(4 * j) % 12] = fmaf(v.x, w0, acc[(4 * j) % 12]); acc[(4 * j + 1) % 12] = fmaf(v.y, w0, acc[(4 * j + 1) % 12]);
                acc[(4 * j + 2) % 12] = fmaf(v.z, w0, acc[(4 * j + 2) % 12]); acc[(4 * j + 3) % 12] = fmaf(v.w, w0, acc[(4 * j + 3) % 12]);
            }
            nsect += 12;
        }
        // the next footprint depends on the loaded data (values are tiny: the contribution is 0 at run time)
        float t = 0.f;
#pragma unroll
        for (int c = 0; c < 12; c++) t += acc[c];
        s ^= (uint32_t)(fabsf(t) * 1e-30f);
    }
    float t = 0.f;
#pragma unroll
    for (int c = 0; c < 12; c++) t += acc[c];
    if (t == 123.456f) a.sink[0] = t;
#pragma unroll
    for (int o = 16; o; o >>= 1) nsect += __shfl_xor_sync(kFull, nsect, o);
    if (lane == 0) atomicAdd(a.sectors, nsect);
}

template <int MODE, int WARPS>
static void run(const char* name, Args a, int sms, double clock_ghz) {
    size_t smem = 128;
    if (MODE == LANE_BULK96 || MODE == SMEM_ONLY) smem += (size_t)WARPS * 32 * 400;
    if (MODE == LANE_TMABOX) smem += (size_t)WARPS * 32 * 384;
    if (smem > 227 * 1024) return;
    CK(cudaFuncSetAttribute(gather_kernel<MODE, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; rep++) {  // first launch warms L2
        CK(cudaMemset(a.sectors, 0, 8));
        CK(cudaEventRecord(e0));
        gather_kernel<MODE, WARPS><<<sms, WARPS * 32, smem>>>(a);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
    }
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    unsigned long long sect;
    CK(cudaMemcpy(&sect, a.sectors, 8, cudaMemcpyDeviceToHost));
    const double samples = (double)sms * WARPS * 32 * a.iters;
    const double sps = sect / (ms * 1e-3);
    printf("{\"variant\": \"%s\", \"warps_per_sm\": %d, \"coherent\": %d, \"ms\": %.4f, \"footprints_per_s\": %.4g, \"sectors_per_s\": %.4g, "
           "\"GBps\": %.1f, \"sectors_per_clk_per_sm\": %.3f, \"cycles_per_footprint_per_sm\": %.2f}\n",
           name, WARPS, a.coherent, ms, samples / (ms * 1e-3), sps, sps * 32 / 1e9, sps / (sms * clock_ghz * 1e9),
           (ms * 1e-3 * clock_ghz * 1e9) / (samples / sms));
    fflush(stdout);
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    int dev = 0, sms = 0, khz = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    const double ghz = khz * 1e-6;
    const int iters = argc > 1 ? atoi(argv[1]) : 400;
    const size_t V = (size_t)D * H * W;
    float *t24, *t32, *t12, *sink;
    unsigned long long* sectors;
    CK(cudaMalloc(&t24, V * 24 * 4 + 4096)); CK(cudaMalloc(&t32, V * 32 * 4)); CK(cudaMalloc(&t12, V * 12 * 4));
    CK(cudaMalloc(&sink, 64)); CK(cudaMalloc(&sectors, 8));
    CK(cudaMemset(t24, 0, V * 24 * 4 + 4096)); CK(cudaMemset(t32, 0, V * 32 * 4)); CK(cudaMemset(t12, 0, V * 12 * 4));
    // tensor map over the plain voxel-major field [D][H][W][12 floats], box = 12 x 2 x 2 x 2, zero fill outside
    CUtensorMap hmap, *dmap = nullptr;
    bool have_map = false;
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && fn) {
            const cuuint64_t dims[4] = {12, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D};
            const cuuint64_t strides[3] = {48, 48ull * W, 48ull * W * H};
            const cuuint32_t box[4] = {12, 2, 2, 2}, es[4] = {1, 1, 1, 1};
            CUresult r = ((EncodeTiled)fn)(&hmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, t12, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS) {
                CK(cudaMalloc(&dmap, sizeof(CUtensorMap)));
                CK(cudaMemcpy(dmap, &hmap, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
                have_map = true;
            } else fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
        }
    }
    printf("{\"device_sms\": %d, \"clock_ghz\": %.3f, \"table_MB\": %.1f, \"iters\": %d}\n", sms, ghz, V * 96 / 1e6, iters);
    for (int coherent = 0; coherent < 2; coherent++) {
        Args a{t24, 24, iters, coherent, sink, dmap, sectors};
        Args a32{t32, 32, iters, coherent, sink, dmap, sectors};
        run<LANE_LDG256, 8>("lane_ldg256", a, sms, ghz);
        run<LANE_LDG256, 12>("lane_ldg256", a, sms, ghz);
        run<LANE_LDG256, 16>("lane_ldg256", a, sms, ghz);
        run<LANE_LDG256, 24>("lane_ldg256", a, sms, ghz);
        run<LANE_LDG256, 32>("lane_ldg256", a, sms, ghz);
        run<LANE_LDG128, 12>("lane_ldg128", a, sms, ghz);
        run<LANE_LDG128, 32>("lane_ldg128", a, sms, ghz);
        run<LANE_SECTOR, 12>("lane_one_sector", a, sms, ghz);
        run<LANE_SECTOR, 32>("lane_one_sector", a, sms, ghz);
        run<QUAD_LINE, 12>("quad_line128", a32, sms, ghz);
        run<QUAD_LINE, 32>("quad_line128", a32, sms, ghz);
        run<OCT_LINE, 12>("oct_line128", a32, sms, ghz);
        run<OCT_LINE, 32>("oct_line128", a32, sms, ghz);
        run<LANE_BULK96, 8>("lane_bulk96_to_smem", a, sms, ghz);
        run<LANE_BULK96, 16>("lane_bulk96_to_smem", a, sms, ghz);
        if (have_map) {
            Args am{t12, 12, iters, coherent, sink, dmap, sectors};
            run<LANE_TMABOX, 8>("lane_tma_box_to_smem", am, sms, ghz);
            run<LANE_TMABOX, 16>("lane_tma_box_to_smem", am, sms, ghz);
        }
    }
    Args a{t24, 24, iters, 0, sink, dmap, sectors};
    run<SMEM_ONLY, 8>("smem_readback_only", a, sms, ghz);
    run<SMEM_ONLY, 16>("smem_readback_only", a, sms, ghz);
    return 0;
}
