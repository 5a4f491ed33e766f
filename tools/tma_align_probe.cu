// Probe: does cp.async.bulk.tensor.3d (tiled, f32, no swizzle) accept an innermost start coordinate that is not a
// multiple of 4 elements (16 bytes)?  Each case runs in its own process-level CUDA context reset so a fault is
// attributed to the case.   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma_align_probe tools/tma_align_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("  %s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_probe(const __grid_constant__ CUtensorMap tmap, float* out, int x0, int y0, int bw, int bh)
{
    extern __shared__ __align__(128) unsigned char sm[];
    float* tile = (float*)sm;
    uint64_t* bar = (uint64_t*)(sm + (size_t)bw * bh * 4);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bw * bh * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(s32(tile)), "l"(&tmap), "r"(s32(bar)), "r"(x0), "r"(y0), "r"(0) : "memory");
    }
    __syncthreads();
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0, 0x2000;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                 ::"r"(s32(bar)) : "memory");
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = tile[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int run(int x0, int y0, int W, int H, int pitch, int bw, int bh)
{
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    float* g; CK(cudaMalloc(&g, (size_t)pitch * H * 4));
    std::vector<float> h((size_t)pitch * H);
    for (int y = 0; y < H; ++y) for (int x = 0; x < pitch; ++x) h[(size_t)y * pitch + x] = (float)(y * 10000 + x);
    CK(cudaMemcpy(g, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, 1};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("  encode failed %d\n", (int)r); return 1; }
    float* out; CK(cudaMalloc(&out, (size_t)bw * bh * 4));
    const size_t smem = (size_t)bw * bh * 4 + 64;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_probe<<<1, 128, smem>>>(tm, out, x0, y0, bw, bh);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> o((size_t)bw * bh);
    CK(cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int y = 0; y < bh; ++y)
        for (int x = 0; x < bw; ++x) {
            const int gx = x0 + x, gy = y0 + y;
            const float want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? (float)(gy * 10000 + gx) : 0.f;
            if (o[(size_t)y * bw + x] != want) ++bad;
        }
    printf("  x0=%d y0=%d box=%dx%d W=%d pitch=%d: %s (%ld mismatches)\n", x0, y0, bw, bh, W, pitch, bad ? "WRONG DATA" : "ok", bad);
    return bad ? 2 : 0;
}

int main(int argc, char** argv)
{
    if (argc >= 8) return run(atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]));
    printf("usage: x0 y0 W H pitch bw bh\n");
    return 0;
}
