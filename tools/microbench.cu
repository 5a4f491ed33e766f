// Micro-benchmark (measurement tool, not product): FP32 FMA issue ceiling on sm_100a with the operand
// pattern of k_ncc_tiled's inner loop, scalar FFMA vs packed FFMA2 (fma.rn.f32x2), registers only and
// with the loop's shared-memory loads.  Prints achieved TFLOP/s and the fraction of SMs*128*2*clk.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long pack(float a, float b) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void unpack(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}

// MODE 0: scalar FFMA, registers only.  MODE 1: FFMA2 registers only.
// MODE 2: scalar FFMA + smem loads like the real loop.  MODE 3: FFMA2 + smem loads.
template <int MODE, int CY>
__global__ void __launch_bounds__(256, 1) k_fma(float* out, int iters, int pitch)
{
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 64 * pitch; i += blockDim.x) sm[i] = 1.0f + 1e-6f * (float)i;
    __syncthreads();
    const float* base = sm + (threadIdx.x & 31) * pitch + (threadIdx.x >> 5) * 8;
    float sum = 0.f;
    if (MODE == 0 || MODE == 2) {
        float acc[CY][8], wa[CY][8], wb[CY][8], t[8];
#pragma unroll
        for (int cy = 0; cy < CY; ++cy)
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc[cy][i] = 0.f; wa[cy][i] = base[i + cy]; wb[cy][i] = base[8 + i + cy]; }
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = sm[i] * 1e-3f;
        for (int it = 0; it < iters; ++it) {
            if (MODE == 2) {
                const float* p = base + (it & 7) * 8;
#pragma unroll
                for (int cy = 0; cy < CY; ++cy) {
                    float4 a = *reinterpret_cast<const float4*>(p + cy * 8 * pitch + 8);
                    float4 b = *reinterpret_cast<const float4*>(p + cy * 8 * pitch + 12);
                    wb[cy][0] = a.x; wb[cy][1] = a.y; wb[cy][2] = a.z; wb[cy][3] = a.w; wb[cy][4] = b.x; wb[cy][5] = b.y; wb[cy][6] = b.z; wb[cy][7] = b.w;
                }
                float4 a = *reinterpret_cast<const float4*>(sm + (it & 15) * 8);
                float4 b = *reinterpret_cast<const float4*>(sm + (it & 15) * 8 + 4);
                t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = a.w; t[4] = b.x; t[5] = b.y; t[6] = b.z; t[7] = b.w;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int cy = 0; cy < CY; ++cy)
#pragma unroll
                    for (int cx = 0; cx < 8; ++cx) {
                        const int i = k + cx;
                        const float v = i < 8 ? wa[cy][i] : wb[cy][i - 8];
                        acc[cy][cx] = fmaf(v, t[k], acc[cy][cx]);
                    }
#pragma unroll
            for (int cy = 0; cy < CY; ++cy)
#pragma unroll
                for (int i = 0; i < 8; ++i) { float x = wa[cy][i]; wa[cy][i] = wb[cy][i]; wb[cy][i] = x; }
        }
#pragma unroll
        for (int cy = 0; cy < CY; ++cy)
#pragma unroll
            for (int i = 0; i < 8; ++i) sum += acc[cy][i];
    } else {
        // candidate cx owns an accumulator PAIR: .x gathers even-dx products, .y odd-dx products (cx even),
        // shifted by one for odd cx (second, shifted template copy): every operand is an aligned register pair.
        unsigned long long acc[CY][8], wa[CY][4], wb[CY][4], te[4], to[4];
#pragma unroll
        for (int cy = 0; cy < CY; ++cy) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[cy][i] = 0ull;
#pragma unroll
            for (int i = 0; i < 4; ++i) { wa[cy][i] = pack(base[2 * i + cy], base[2 * i + 1 + cy]); wb[cy][i] = pack(base[8 + 2 * i + cy], base[9 + 2 * i + cy]); }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { te[i] = pack(sm[2 * i] * 1e-3f, sm[2 * i + 1] * 1e-3f); to[i] = pack(sm[2 * i + 1] * 1e-3f, sm[2 * i + 2] * 1e-3f); }
        for (int it = 0; it < iters; ++it) {
            if (MODE == 3) {
                const float* p = base + (it & 7) * 8;
#pragma unroll
                for (int cy = 0; cy < CY; ++cy) {
                    ulonglong2 a = *reinterpret_cast<const ulonglong2*>(p + cy * 8 * pitch + 8);
                    ulonglong2 b = *reinterpret_cast<const ulonglong2*>(p + cy * 8 * pitch + 12);
                    wb[cy][0] = a.x; wb[cy][1] = a.y; wb[cy][2] = b.x; wb[cy][3] = b.y;
                }
                ulonglong2 a = *reinterpret_cast<const ulonglong2*>(sm + (it & 15) * 8);
                ulonglong2 b = *reinterpret_cast<const ulonglong2*>(sm + (it & 15) * 8 + 4);
                ulonglong2 c = *reinterpret_cast<const ulonglong2*>(sm + 256 + (it & 15) * 8);
                ulonglong2 d = *reinterpret_cast<const ulonglong2*>(sm + 256 + (it & 15) * 8 + 4);
                te[0] = a.x; te[1] = a.y; te[2] = b.x; te[3] = b.y; to[0] = c.x; to[1] = c.y; to[2] = d.x; to[3] = d.y;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)            // pair index along dx
#pragma unroll
                for (int cy = 0; cy < CY; ++cy)
#pragma unroll
                    for (int cx = 0; cx < 8; ++cx) {
                        // window pair index (cx even: (cx + 2k)/2 ; cx odd: (cx + 1 + 2k)/2)
                        const int wi = (cx + (cx & 1) + 2 * k) >> 1;
                        const unsigned long long v = wi < 4 ? wa[cy][wi] : wb[cy][wi - 4];
                        acc[cy][cx] = ffma2(v, (cx & 1) ? to[k] : te[k], acc[cy][cx]);
                    }
#pragma unroll
            for (int cy = 0; cy < CY; ++cy)
#pragma unroll
                for (int i = 0; i < 4; ++i) { unsigned long long x = wa[cy][i]; wa[cy][i] = wb[cy][i]; wb[cy][i] = x; }
        }
#pragma unroll
        for (int cy = 0; cy < CY; ++cy)
#pragma unroll
            for (int i = 0; i < 8; ++i) { float a, b; unpack(acc[cy][i], a, b); sum += a + b; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}


// PAT 0: acc[i] += a*b (a, b loop invariant)   PAT 1: acc[i] += w[i]*b (fixed pairing, b invariant)
// PAT 2: acc[i] += w[i]*t[i] (fixed pairing, three fresh operands)
template <int PAT>
__global__ void __launch_bounds__(256, 1) k_pat(float* out, int iters, float a0, float b0, long long* clk)
{
    float acc[32], w[32], t[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { acc[i] = 0.f; w[i] = a0 + i * 1e-3f + threadIdx.x * 1e-5f; t[i] = b0 + i * 1e-4f + threadIdx.x * 1e-7f; }
    long long c0 = clock64();
    unsigned long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 32; ++i)
                acc[i] = PAT == 0 ? fmaf(a0, b0, acc[i]) : PAT == 1 ? fmaf(w[i], b0, acc[i]) : PAT == 2 ? fmaf(w[i], t[i], acc[i])
                       : PAT == 3 ? fmaf(w[i], t[r], acc[i])            /* vector t, same for the 32 FMAs of a group (.reuse) */
                       : PAT == 4 ? fmaf(w[i], w[i], acc[i])            /* two reads of ONE register + accumulator */
                                  : fmaf(w[(i + r) & 31], t[r], acc[i]); /* sliding pairing + reused vector t */
    }
    long long c1 = clock64();
    unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) sum += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    if (blockIdx.x == 0 && threadIdx.x == 0) { clk[0] = c1 - c0; clk[1] = (long long)(g1 - g0); }
}

template <int PAT>
void run_pat(const char* name, int sms, double clk_ghz, int threads)
{
    const int iters = 8192;
    float* out; CK(cudaMalloc(&out, (size_t)sms * threads * 4));
    long long* clk; CK(cudaMalloc(&clk, 16));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int w = 0; w < 2; ++w) k_pat<PAT><<<sms, threads>>>(out, iters, 1.0f, 1e-3f, clk);
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(a));
        k_pat<PAT><<<sms, threads>>>(out, iters, 1.0f, 1e-3f, clk);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    long long h[2]; CK(cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost));
    const double fma = (double)sms * threads * iters * 256.0;
    const double tf = 2.0 * fma / (best * 1e-3) / 1e12;
    const double warps_per_smsp = threads / 32.0 / 4.0;
    printf("%-34s thr=%4d  %8.3f ms  %7.2f TFLOP/s  %5.1f%% of nominal | in-kernel %.0f MHz, %.3f cyc per FFMA per SMSP\n", name, threads, best, tf,
           100.0 * tf / (sms * 128 * 2 * clk_ghz * 1e-3), 1e3 * (double)h[0] / (double)h[1], (double)h[0] / (iters * 256.0 * warps_per_smsp));
    CK(cudaFree(out)); CK(cudaFree(clk));
}

template <int MODE, int CY>
void run(const char* name, int sms, double clk_ghz, int threads, int ctas_per_sm)
{
    const int pitch = 100, iters = 4096;
    const size_t smem = 64 * pitch * 4;
    float* out; CK(cudaMalloc(&out, (size_t)sms * ctas_per_sm * threads * 4));
    CK(cudaFuncSetAttribute(k_fma<MODE, CY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int w = 0; w < 2; ++w) k_fma<MODE, CY><<<sms * ctas_per_sm, threads, smem>>>(out, iters, pitch);
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(a));
        k_fma<MODE, CY><<<sms * ctas_per_sm, threads, smem>>>(out, iters, pitch);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    const double fma = (double)sms * ctas_per_sm * threads * iters * 64.0 * CY;
    const double tf = 2.0 * fma / (best * 1e-3) / 1e12;
    printf("%-34s CY=%d thr=%4d cta/sm=%d  %8.3f ms  %7.2f TFLOP/s  %5.1f%% of %.1f (SMs*128*2*%.3f GHz)\n", name, CY, threads, ctas_per_sm, best, tf,
           100.0 * tf / (sms * 128 * 2 * clk_ghz * 1e-3), sms * 128 * 2 * clk_ghz * 1e-3, clk_ghz);
    CK(cudaFree(out));
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    const double ghz = clk * 1e-6;
    printf("device %s, %d SMs, max clock %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    const int sms = p.multiProcessorCount;
    run_pat<0>("FFMA acc+=a*b (invariant a,b)", sms, ghz, 256);
    run_pat<1>("FFMA acc+=w[i]*b (fixed pairing)", sms, ghz, 256);
    run_pat<2>("FFMA acc+=w[i]*t[i] (3 fresh)", sms, ghz, 256);
    run_pat<3>("FFMA acc+=w[i]*t[r] (vector t reused x32)", sms, ghz, 256);
    run_pat<3>("FFMA acc+=w[i]*t[r] (vector t reused x32)", sms, ghz, 128);
    run_pat<4>("FFMA acc+=w[i]*w[i]", sms, ghz, 256);
    run_pat<5>("FFMA acc+=w[(i+r)%32]*t[r] (sliding, t reused)", sms, ghz, 256);
    run<0, 4>("FFMA  regs only", sms, ghz, 256, 1);
    run<0, 4>("FFMA  regs only", sms, ghz, 128, 1);
    run<0, 4>("FFMA  regs only", sms, ghz, 128, 3);
    run<1, 4>("FFMA2 regs only", sms, ghz, 256, 1);
    run<1, 4>("FFMA2 regs only", sms, ghz, 128, 1);
    run<2, 4>("FFMA  + LDS.128 (loop shape)", sms, ghz, 256, 1);
    run<2, 4>("FFMA  + LDS.128 (loop shape)", sms, ghz, 128, 2);
    run<2, 2>("FFMA  + LDS.128 (loop shape)", sms, ghz, 256, 2);
    run<3, 4>("FFMA2 + LDS.128 (loop shape)", sms, ghz, 256, 1);
    run<3, 4>("FFMA2 + LDS.128 (loop shape)", sms, ghz, 128, 2);
    run<3, 2>("FFMA2 + LDS.128 (loop shape)", sms, ghz, 256, 2);
    return 0;
}
