// Micro-benchmark 2 (measurement tool, not product): inner-loop formulations of the NCC cross term.
// Every variant computes the same thing as k_ncc_tiled's dy/dx loops for one thread tile
// (8 x CY candidates, template row of TP floats in shared memory, sliding 16-float window) and differs
// only in instruction ORDER / operand sourcing, to find what the sm_100a operand collector sustains.
//   V0  for k / for cy / for cx                      (the shipped source order)
//   V1  V0 + __syncwarp() after every k group         (scheduling fence)
//   V2  for cy / for k / for cx                      (8-FMA runs, one candidate row at a time)
//   V3  for each window element i: all (cx,k=i-cx)   (window value is the reused operand)
//   V4  template from __constant__ memory            (uniform-register operand; single-template upper bound)
//   V5  V0 with the loads for the NEXT 8-dx chunk issued before the FMAs of the current one (explicit prefetch)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int TP = 64, TH = 64, P = 92, SB = 41, ROWS = 4 * SB + TH + 8;  // tile like the 64x64 / R80 case
__constant__ float c_templ[TP * TH];
__device__ volatile int g_never_src;


template <int CY>
__device__ __forceinline__ void load8(float (&w)[CY][8], const float* p, int rstride)
{
#pragma unroll
    for (int cy = 0; cy < CY; ++cy) {
        const float4 a = *reinterpret_cast<const float4*>(p + cy * rstride);
        const float4 b = *reinterpret_cast<const float4*>(p + cy * rstride + 4);
        w[cy][0] = a.x; w[cy][1] = a.y; w[cy][2] = a.z; w[cy][3] = a.w; w[cy][4] = b.x; w[cy][5] = b.y; w[cy][6] = b.z; w[cy][7] = b.w;
    }
}
__device__ __forceinline__ void loadt(float (&t)[8], const float* p)
{
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = a.w; t[4] = b.x; t[5] = b.y; t[6] = b.z; t[7] = b.w;
}

template <int V, int CY>
__device__ __forceinline__ void sweep(float (&r)[CY][8], const float (&lo)[CY][8], const float (&hi)[CY][8], const float (&t)[8], int g_never = 0)
{
    if (V == 2) {
#pragma unroll
        for (int cy = 0; cy < CY; ++cy)
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int cx = 0; cx < 8; ++cx) { const int i = k + cx; r[cy][cx] = fmaf(i < 8 ? lo[cy][i] : hi[cy][i - 8], t[k], r[cy][cx]); }
    } else if (V == 3) {
#pragma unroll
        for (int cy = 0; cy < CY; ++cy)
#pragma unroll
            for (int i = 0; i < 15; ++i)
#pragma unroll
                for (int cx = 0; cx < 8; ++cx) { const int k = i - cx; if (k >= 0 && k < 8) r[cy][cx] = fmaf(i < 8 ? lo[cy][i] : hi[cy][i - 8], t[k], r[cy][cx]); }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int cy = 0; cy < CY; ++cy)
#pragma unroll
                for (int cx = 0; cx < 8; ++cx) { const int i = k + cx; r[cy][cx] = fmaf(i < 8 ? lo[cy][i] : hi[cy][i - 8], t[k], r[cy][cx]); }
            if (V == 1) __syncwarp();
            if (V == 6 || (V == 7 && (k & 1))) { if (g_never) asm volatile("trap;"); }
        }
    }
}

template <int V, int CY>
__global__ void __launch_bounds__(128, 2) k_loop(float* out, const float* __restrict__ gt, int reps, int never)
{
    extern __shared__ __align__(16) float sm[];
    float* s_tile = sm;
    float* s_templ = sm + ROWS * P;
    for (int i = threadIdx.x; i < ROWS * P; i += blockDim.x) s_tile[i] = 0.5f + 1e-4f * (float)(i % 977);
    for (int i = threadIdx.x; i < TP * TH; i += blockDim.x) s_templ[i] = gt[i];
    __syncthreads();
    const int q = threadIdx.x, col = q / SB, slot = q - col * SB;
    if (col >= 3) return;
    const float* base = s_tile + slot * P + col * 8;
    const int rstride = SB * P;
    float acc[CY][8];
#pragma unroll
    for (int cy = 0; cy < CY; ++cy)
#pragma unroll
        for (int cx = 0; cx < 8; ++cx) acc[cy][cx] = 0.f;
    for (int rep = 0; rep < reps; ++rep)
        for (int dy = 0; dy < TH; ++dy) {
            const float* frow = base + dy * P;
            const float* trow = (V == 4 ? c_templ : s_templ) + dy * TP;
            float racc[CY][8];
#pragma unroll
            for (int cy = 0; cy < CY; ++cy)
#pragma unroll
                for (int cx = 0; cx < 8; ++cx) racc[cy][cx] = 0.f;
            float wa[CY][8], wb[CY][8], ta[8], tb[8];
            load8<CY>(wa, frow, rstride);
            if (V == 5) {
                loadt(ta, trow);
                load8<CY>(wb, frow + 8, rstride);
#pragma unroll 1
                for (int j = 0; j < TP; j += 16) {
                    loadt(tb, trow + j + 8);
                    sweep<0, CY>(racc, wa, wb, ta);
                    load8<CY>(wa, frow + j + 16, rstride);        // window for chunk j+16.. (harmless over-read at the end)
                    if (j + 16 < TP) loadt(ta, trow + j + 16);
                    sweep<0, CY>(racc, wb, wa, tb);
                    load8<CY>(wb, frow + j + 24, rstride);
                }
            } else if (V == 8 || V == 9) {
                // V8: template values loaded once per row (no t loads in the chunk loop); V9: as V0 but every window load issued twice
                loadt(ta, trow);
#pragma unroll 1
                for (int j = 0; j < TP; j += 16) {
                    load8<CY>(wb, frow + j + 8, rstride);
                    if (V == 9) { load8<CY>(wb, frow + j + 8 + 4 * P, rstride); loadt(ta, trow + j); }
                    sweep<0, CY>(racc, wa, wb, ta);
                    load8<CY>(wa, frow + j + 16, rstride);
                    if (V == 9) { load8<CY>(wa, frow + j + 16 + 4 * P, rstride); loadt(ta, trow + j + 8); }
                    sweep<0, CY>(racc, wb, wa, ta);
                }
            } else {
#pragma unroll 1
                for (int j = 0; j < TP; j += 16) {
                    load8<CY>(wb, frow + j + 8, rstride);
                    loadt(ta, trow + j);
                    sweep<V, CY>(racc, wa, wb, ta, never);
                    load8<CY>(wa, frow + j + 16, rstride);
                    loadt(ta, trow + j + 8);
                    sweep<V, CY>(racc, wb, wa, ta, never);
                }
            }
#pragma unroll
            for (int cy = 0; cy < CY; ++cy)
#pragma unroll
                for (int cx = 0; cx < 8; ++cx) acc[cy][cx] += racc[cy][cx];
        }
    float s = 0.f;
#pragma unroll
    for (int cy = 0; cy < CY; ++cy)
#pragma unroll
        for (int cx = 0; cx < 8; ++cx) s += acc[cy][cx];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// ---- y-sliding formulation: thread = 8 x-candidates x CY ADJACENT candidate rows; loop order: 8-dx chunk (outer),
// template row dy (inner).  Per dy: ONE new 16-float row window (4 LDS.128) + the 8 template values (2 LDS.128)
// feed 64*CY FFMA; the CY live row windows rotate through statically named registers (dy unrolled by CY).
template <int CY>
__global__ void __launch_bounds__(128, 2) k_loop_y(float* out, const float* __restrict__ gt, int reps, int G /* row groups per column */)
{
    extern __shared__ __align__(16) float sm[];
    float* s_tile = sm;
    float* s_templ = sm + ROWS * P;
    for (int i = threadIdx.x; i < ROWS * P; i += blockDim.x) s_tile[i] = 0.5f + 1e-4f * (float)(i % 977);
    for (int i = threadIdx.x; i < TP * TH; i += blockDim.x) s_templ[i] = gt[i];
    __syncthreads();
    // all 128 threads active for every CY: 4 columns x 32 row groups, rows wrapped into the tile (perf test only)
    const int q = threadIdx.x, col = q >> 5, g = q & 31;
    (void)G;
    const float* base = s_tile + ((CY * g) % (ROWS - TH - CY)) * P + (col % 3) * 8;
    float acc[CY][8];
#pragma unroll
    for (int i = 0; i < CY; ++i)
#pragma unroll
        for (int cx = 0; cx < 8; ++cx) acc[i][cx] = 0.f;
    for (int rep = 0; rep < reps; ++rep)
        for (int j = 0; j < TP; j += 8) {
            float racc[CY][8], w[CY][16], t[8];
#pragma unroll
            for (int i = 0; i < CY; ++i)
#pragma unroll
                for (int cx = 0; cx < 8; ++cx) racc[i][cx] = 0.f;
#pragma unroll
            for (int r = 0; r < CY - 1; ++r) {
                const float* p = base + r * P + j;
#pragma unroll
                for (int v = 0; v < 4; ++v) { const float4 a = *reinterpret_cast<const float4*>(p + 4 * v); w[r][4 * v] = a.x; w[r][4 * v + 1] = a.y; w[r][4 * v + 2] = a.z; w[r][4 * v + 3] = a.w; }
            }
#pragma unroll 1
            for (int dy0 = 0; dy0 < TH; dy0 += CY) {
#pragma unroll
                for (int u = 0; u < CY; ++u) {
                    const int dy = dy0 + u;
                    if (dy < TH) {
                        const float* p = base + (dy + CY - 1) * P + j;
                        float(&wn)[16] = w[(u + CY - 1) % CY];
#pragma unroll
                        for (int v = 0; v < 4; ++v) { const float4 a = *reinterpret_cast<const float4*>(p + 4 * v); wn[4 * v] = a.x; wn[4 * v + 1] = a.y; wn[4 * v + 2] = a.z; wn[4 * v + 3] = a.w; }
                        loadt(t, s_templ + dy * TP + j);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
#pragma unroll
                            for (int i = 0; i < CY; ++i)
#pragma unroll
                                for (int cx = 0; cx < 8; ++cx) racc[i][cx] = fmaf(w[(u + i) % CY][k + cx], t[k], racc[i][cx]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < CY; ++i)
#pragma unroll
                for (int cx = 0; cx < 8; ++cx) acc[i][cx] += racc[i][cx];
        }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CY; ++i)
#pragma unroll
        for (int cx = 0; cx < 8; ++cx) s += acc[i][cx];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CY>
void run_y(const char* name, int sms, double ghz, const float* gt)
{
    const int reps = 8, ctas = sms * 2;
    const int G = (ROWS - TH - 8) / CY < 42 ? (ROWS - TH - 8) / CY : 42;   // row groups per column that fit the tile
    const int active = 128;
    const size_t smem = (size_t)(ROWS * P + TP * TH + 64) * 4;
    float* out; CK(cudaMalloc(&out, (size_t)ctas * 128 * 4));
    CK(cudaFuncSetAttribute(k_loop_y<CY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k_loop_y<CY><<<ctas, 128, smem>>>(out, gt, 1, G);
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(a));
        k_loop_y<CY><<<ctas, 128, smem>>>(out, gt, reps, G);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    const double fma = (double)ctas * active * reps * TH * TP * 8.0 * CY;
    const double tf = 2.0 * fma / (best * 1e-3) / 1e12;
    printf("Y  CY=%d %-46s %8.3f ms  %6.2f TFLOP/s  %5.1f%% of nominal (%.1f%% of the %d/128 lane ceiling)\n", CY, name, best, tf,
           100.0 * tf / (sms * 128 * 2 * ghz * 1e-3), 100.0 * tf / (sms * 128 * 2 * ghz * 1e-3) * 128.0 / active, active);
    CK(cudaFree(out));
}


// ---- y-sliding, explicitly software-pipelined: the window row and template values of step dy+1 are loaded into a spare
// register set before the FMAs of step dy; the chunk totals live in shared memory (frees 8*CY registers for the spare set).
template <int CY>
__global__ void __launch_bounds__(128, 2) k_loop_yp(float* out, const float* __restrict__ gt, int reps)
{
    extern __shared__ __align__(16) float sm[];
    float* s_tile = sm;
    float* s_templ = sm + ROWS * P;
    float* s_acc = s_templ + TP * TH;                       // [8*CY][128] chunk totals
    for (int i = threadIdx.x; i < ROWS * P; i += blockDim.x) s_tile[i] = 0.5f + 1e-4f * (float)(i % 977);
    for (int i = threadIdx.x; i < TP * TH; i += blockDim.x) s_templ[i] = gt[i];
    for (int i = threadIdx.x; i < 8 * CY * 128; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const int q = threadIdx.x, col = q >> 5, g = q & 31;
    const float* base = s_tile + ((CY * g) % (ROWS - TH - CY)) * P + (col % 3) * 8;
    for (int rep = 0; rep < reps; ++rep)
        for (int j = 0; j < TP; j += 8) {
            float racc[CY][8], w[CY + 1][16], t[2][8];
#pragma unroll
            for (int i = 0; i < CY; ++i)
#pragma unroll
                for (int cx = 0; cx < 8; ++cx) racc[i][cx] = 0.f;
#pragma unroll
            for (int r = 0; r < CY; ++r) {                  // rows 0 .. CY-1 (the window of step 0 complete)
                const float* p = base + r * P + j;
#pragma unroll
                for (int v = 0; v < 4; ++v) { const float4 a = *reinterpret_cast<const float4*>(p + 4 * v); w[r][4 * v] = a.x; w[r][4 * v + 1] = a.y; w[r][4 * v + 2] = a.z; w[r][4 * v + 3] = a.w; }
            }
            loadt(t[0], s_templ + j);
            // steps run in groups of CY+1 so that the (CY+1)-slot window ring and the 2-slot template ring rotate through
            // static register names
#pragma unroll 1
            for (int dy0 = 0; dy0 < TH; dy0 += (CY + 1)) {
#pragma unroll
                for (int u = 0; u < (CY + 1); ++u) {
                    const int dy = dy0 + u;
                    if (dy < TH) {
                        // prefetch for step dy+1: window row dy+CY into the spare slot, template row dy+1
                        const float* p = base + (dy + CY) * P + j;
                        float(&wn)[16] = w[(u + CY) % (CY + 1)];
#pragma unroll
                        for (int v = 0; v < 4; ++v) { const float4 a = *reinterpret_cast<const float4*>(p + 4 * v); wn[4 * v] = a.x; wn[4 * v + 1] = a.y; wn[4 * v + 2] = a.z; wn[4 * v + 3] = a.w; }
                        loadt(t[(u + 1) & 1], s_templ + (dy + 1 < TH ? dy + 1 : dy) * TP + j);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
#pragma unroll
                            for (int i = 0; i < CY; ++i)
#pragma unroll
                                for (int cx = 0; cx < 8; ++cx) racc[i][cx] = fmaf(w[(u + i) % (CY + 1)][k + cx], t[u & 1][k], racc[i][cx]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < CY; ++i)
#pragma unroll
                for (int cx = 0; cx < 8; ++cx) s_acc[(i * 8 + cx) * 128 + threadIdx.x] += racc[i][cx];
        }
    float s = 0.f;
    for (int i = 0; i < 8 * CY; ++i) s += s_acc[i * 128 + threadIdx.x];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CY, bool PIPE>
void run_y2(const char* name, int sms, double ghz, const float* gt, int ctas_per_sm)
{
    const int reps = 8, ctas = sms * ctas_per_sm;
    const size_t smem = (size_t)(ROWS * P + TP * TH + 64 + (PIPE ? 8 * CY * 128 : 0)) * 4;
    float* out; CK(cudaMalloc(&out, (size_t)ctas * 128 * 4));
    if (PIPE) CK(cudaFuncSetAttribute(k_loop_yp<CY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CK(cudaFuncSetAttribute(k_loop_y<CY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    auto go = [&](int r) { if (PIPE) k_loop_yp<CY><<<ctas, 128, smem>>>(out, gt, r); else k_loop_y<CY><<<ctas, 128, smem>>>(out, gt, r, 32); };
    go(1);
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(a)); go(reps); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    const double fma = (double)ctas * 128 * reps * TH * TP * 8.0 * CY;
    const double tf = 2.0 * fma / (best * 1e-3) / 1e12;
    printf("Y%s CY=%d %-40s ctas/SM=%d %8.3f ms  %6.2f TFLOP/s  %5.1f%% of nominal (x%d warps per sub-partition)\n", PIPE ? "P" : " ", CY, name, ctas_per_sm, best, tf,
           100.0 * tf / (sms * 128 * 2 * ghz * 1e-3), ctas_per_sm);
    CK(cudaFree(out));
}


// ---- y-sliding CY=5, 12 warps per SM in ONE 384-thread CTA: chunk totals in shared memory so the loop fits 168 registers
template <int CY, int NT>
__global__ void __launch_bounds__(NT, 1) k_loop_y12(float* out, const float* __restrict__ gt, int reps, int rows)
{
    extern __shared__ __align__(16) float sm[];
    float* s_tile = sm;
    float* s_templ = sm + rows * P;
    float* s_acc = s_templ + TP * TH;                       // [8*CY][NT] chunk totals
    for (int i = threadIdx.x; i < rows * P; i += blockDim.x) s_tile[i] = 0.5f + 1e-4f * (float)(i % 977);
    for (int i = threadIdx.x; i < TP * TH; i += blockDim.x) s_templ[i] = gt[i];
    for (int i = threadIdx.x; i < 8 * CY * NT; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const int q = threadIdx.x, col = q >> 5, g = q & 31;
    const float* base = s_tile + ((CY * g) % (rows - TH - CY)) * P + (col % 3) * 8;
    for (int rep = 0; rep < reps; ++rep)
        for (int j = 0; j < TP; j += 8) {
            float racc[CY][8], w[CY][16], t[8];
#pragma unroll
            for (int i = 0; i < CY; ++i)
#pragma unroll
                for (int cx = 0; cx < 8; ++cx) racc[i][cx] = 0.f;
#pragma unroll
            for (int r = 0; r < CY - 1; ++r) {
                const float* p = base + r * P + j;
#pragma unroll
                for (int v = 0; v < 4; ++v) { const float4 a = *reinterpret_cast<const float4*>(p + 4 * v); w[r][4 * v] = a.x; w[r][4 * v + 1] = a.y; w[r][4 * v + 2] = a.z; w[r][4 * v + 3] = a.w; }
            }
#pragma unroll 1
            for (int dy0 = 0; dy0 < TH; dy0 += CY) {
#pragma unroll
                for (int u = 0; u < CY; ++u) {
                    const int dy = dy0 + u;
                    if (dy < TH) {
                        const float* p = base + (dy + CY - 1) * P + j;
                        float(&wn)[16] = w[(u + CY - 1) % CY];
#pragma unroll
                        for (int v = 0; v < 4; ++v) { const float4 a = *reinterpret_cast<const float4*>(p + 4 * v); wn[4 * v] = a.x; wn[4 * v + 1] = a.y; wn[4 * v + 2] = a.z; wn[4 * v + 3] = a.w; }
                        loadt(t, s_templ + dy * TP + j);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
#pragma unroll
                            for (int i = 0; i < CY; ++i)
#pragma unroll
                                for (int cx = 0; cx < 8; ++cx) racc[i][cx] = fmaf(w[(u + i) % CY][k + cx], t[k], racc[i][cx]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < CY; ++i)
#pragma unroll
                for (int cx = 0; cx < 8; ++cx) s_acc[(i * 8 + cx) * NT + threadIdx.x] += racc[i][cx];
        }
    float s = 0.f;
    for (int i = 0; i < 8 * CY; ++i) s += s_acc[i * NT + threadIdx.x];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CY, int NT>
void run_y12(const char* name, int sms, double ghz, const float* gt)
{
    const int reps = 8, ctas = sms, rows = 228;
    const size_t smem = (size_t)(rows * P + TP * TH + 64 + 8 * CY * NT) * 4;
    float* out; CK(cudaMalloc(&out, (size_t)ctas * NT * 4));
    CK(cudaFuncSetAttribute(k_loop_y12<CY, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k_loop_y12<CY, NT><<<ctas, NT, smem>>>(out, gt, 1, rows);
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(a)); k_loop_y12<CY, NT><<<ctas, NT, smem>>>(out, gt, reps, rows); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    const double fma = (double)ctas * NT * reps * TH * TP * 8.0 * CY;
    const double tf = 2.0 * fma / (best * 1e-3) / 1e12;
    printf("Y12 CY=%d %-40s NT=%d %8.3f ms  %6.2f TFLOP/s  %5.1f%% of nominal (%d warps per sub-partition)\n", CY, name, NT, best, tf,
           100.0 * tf / (sms * 128 * 2 * ghz * 1e-3), NT / 128);
    CK(cudaFree(out));
}

template <int V, int CY>
void run(const char* name, int sms, double ghz, const float* gt)
{
    const int reps = 8, ctas = sms * 2;
    const size_t smem = (size_t)(ROWS * P + TP * TH + 64) * 4;
    float* out; CK(cudaMalloc(&out, (size_t)ctas * 128 * 4));
    CK(cudaFuncSetAttribute(k_loop<V, CY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k_loop<V, CY><<<ctas, 128, smem>>>(out, gt, 1, 0);
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(a));
        k_loop<V, CY><<<ctas, 128, smem>>>(out, gt, reps, 0);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    const double fma = (double)ctas * 123 * reps * TH * TP * 8.0 * CY;   // 123 active threads per CTA
    const double tf = 2.0 * fma / (best * 1e-3) / 1e12;
    // lane-slot ceiling: 128 threads launched per CTA, 123 working
    printf("V%d CY=%d %-46s %8.3f ms  %6.2f TFLOP/s  %5.1f%% of nominal (%.1f%% of the 123/128 lane ceiling)\n", V, CY, name, best, tf,
           100.0 * tf / (sms * 128 * 2 * ghz * 1e-3), 100.0 * tf / (sms * 128 * 2 * ghz * 1e-3) * 128.0 / 123.0);
    CK(cudaFree(out));
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    const double ghz = clk * 1e-6;
    const int sms = p.multiProcessorCount;
    float* gt; CK(cudaMalloc(&gt, TP * TH * 4));
    float* h = (float*)malloc(TP * TH * 4);
    for (int i = 0; i < TP * TH; ++i) h[i] = 1e-3f * (float)((i * 37) % 101 - 50);
    CK(cudaMemcpy(gt, h, TP * TH * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpyToSymbol(c_templ, h, TP * TH * 4));
    printf("device %s, %d SMs, %.3f GHz; 2 CTAs/SM x 128 threads, tile pitch %d, SB %d\n", p.name, sms, ghz, P, SB);
    run<0, 4>("for k/cy/cx (shipped order)", sms, ghz, gt);
    run<1, 4>("V0 + __syncwarp per k group", sms, ghz, gt);
    run<2, 4>("for cy/k/cx", sms, ghz, gt);
    run<3, 4>("per window element (w reused)", sms, ghz, gt);
    run<4, 4>("template in __constant__ (UR operand)", sms, ghz, gt);
    run<5, 4>("V0 + explicit next-chunk prefetch", sms, ghz, gt);
    run<8, 4>("V0 without template loads in the chunk loop", sms, ghz, gt);
    run<9, 4>("V0 with every window load issued twice", sms, ghz, gt);
    run<6, 4>("V0 + opaque branch after every k group", sms, ghz, gt);
    run<7, 4>("V0 + opaque branch after every 2 k groups", sms, ghz, gt);
    run_y2<5, false>("y-sliding", sms, ghz, gt, 1);
    run_y2<5, false>("y-sliding", sms, ghz, gt, 2);
    run_y2<5, true>("y-sliding, software-pipelined", sms, ghz, gt, 1);
    run_y2<5, true>("y-sliding, software-pipelined", sms, ghz, gt, 2);
    run_y12<5, 256>("chunk totals in smem, 1 CTA/SM", sms, ghz, gt);
    run_y12<5, 384>("chunk totals in smem, 1 CTA/SM", sms, ghz, gt);
    run_y12<5, 512>("chunk totals in smem, 1 CTA/SM", sms, ghz, gt);
    run_y<3>("y-sliding", sms, ghz, gt);
    run_y<4>("y-sliding (rows 4 apart: 2-way bank conflicts expected)", sms, ghz, gt);
    run_y<5>("y-sliding", sms, ghz, gt);
    run_y<7>("y-sliding", sms, ghz, gt);
    run<0, 2>("for k/cy/cx", sms, ghz, gt);
    run<6, 2>("V0 + opaque branch after every k group", sms, ghz, gt);
    run<2, 2>("for cy/k/cx", sms, ghz, gt);
    run<4, 2>("template in __constant__ (UR operand)", sms, ghz, gt);
    run<5, 2>("V0 + explicit next-chunk prefetch", sms, ghz, gt);
    return 0;
}
