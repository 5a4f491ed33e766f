// tools/tc_probe.cu -- tcgen05 probe for the tensor-core NCC cross term (measurement tool, not product code).
//
// Formulation under test (DESIGN.md "Tensor-core variant"): per template row dy the search is a GEMM
//     D[y][j] += sum_k  G[y + dy][k] * T_dy[k][j],      T_dy[k][j] = digit(tc[dy][k - j])  (banded Toeplitz, 0 outside 0 <= k-j < tw)
// with A = the 8-bit gray image itself (unsigned 8 bit, exact), B = one signed 8-bit DIGIT of the fixed-point centred
// template, D = exact int32 sums in TMEM (tcgen05.mma kind::i8).  Three digits -> three accumulators, combined in FP64.
//
// What the probe establishes on real hardware:
//  (A) descriptor semantics.  Both operands use the NO-SWIZZLE K-major canonical layout ((8,n),2):((16 B, SBO),LBO):
//      - the image tile is stored 16-pixel-chunk-major, addr(row, chunk) = chunk * CH + row * 16: an 8-row core matrix is 128
//        contiguous bytes for ANY start row, so the row shift dy is a 16-byte bump of the descriptor's start address;
//      - the Toeplitz operand is never materialised: core matrix (8 candidates a, 16 positions m) of T_dy depends only on
//        2m - a, so with the candidates of a block enumerated in REVERSE groups (a' = A-1-a) the matrix is
//        addr(a', m) = base + (2m + a') * 128, i.e. SBO = 128 B, LBO = 256 B over ~22 blocks of 128 B that ALIAS each other;
//      - band-aware issue: a K-step only touches the candidates whose taps it holds, via a column offset into D and a block
//        offset into B.
//      Checked against a CPU integer reference, bit for bit.
//  (B) cycles per tcgen05.mma as a function of (M, N) in SS mode, and for the issue patterns the kernel would use.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x2000;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// no-swizzle K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start >> 4 at [0,14),
// leading byte offset >> 4 at [16,30), stride byte offset >> 4 at [32,46), version 1 at [46,48), layout type 0 at [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// kind::i8 instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2 at [4,6)), A = unsigned 8 bit (0 at [7,10)),
// B = signed 8 bit (1 at [10,13)), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ inline uint32_t make_idesc_i8(int M, int N) { return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__host__ __device__ inline uint32_t make_idesc_f16(int M, int N) { return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mma_f16_nomask(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_i8_nomask(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------------------
// (A) correctness: one CTA, M = 128 candidate rows, one block of NB = 64 candidates, tw = 64 taps, th template rows.
constexpr int kRows = 192, kChunks = 8, kCH = kRows * 16;   // image tile: 192 rows x 128 pixels, chunk-major
constexpr int kA = 8;                                       // candidate groups of 8 in the block (NB = 64)
constexpr int kBlocks = 2 * kChunks + kA;                   // Toeplitz blocks d' = 2m + a' in [0, 22)

__global__ void __launch_bounds__(128) k_probe_correct(const uint8_t* __restrict__ img /* [191][128] */, const int8_t* __restrict__ dig /* [th][64] */,
                                                       int th, int band_aware, int* __restrict__ out /* [128][64], column = candidate j */)
{
    extern __shared__ __align__(128) unsigned char sm[];
    uint8_t* sA = sm;                                   // kChunks * kCH
    uint8_t* sB = sm + kChunks * kCH;                   // kBlocks * 128
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + kBlocks * 128);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // image tile, chunk-major: addr(row, chunk) = chunk * CH + row * 16
    for (int i = tid; i < kRows * 128; i += 128) {
        const int r = i >> 7, x = i & 127;
        sA[(x >> 4) * kCH + r * 16 + (x & 15)] = r < 191 ? img[r * 128 + x] : 0;
    }
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc64 = make_idesc_i8(128, 64), idesc32 = make_idesc_i8(128, 32);
    uint32_t phase = 0;
    for (int dy = 0; dy < th; ++dy) {
        // Toeplitz blocks of this template row: block d', row s (candidate within its group of 8), byte e (position within
        // its 16): tap index 8 (d' - (A-1)) + e - s
        for (int i = tid; i < kBlocks * 128; i += 128) {
            const int d = i >> 7, s = (i >> 4) & 7, e = i & 15;
            const int tap = 8 * (d - (kA - 1)) + e - s;
            sB[i] = (tap >= 0 && tap < 64) ? (uint8_t)dig[dy * 64 + tap] : 0;
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            for (int q = 0; q < 4; ++q) {   // a K-step = 32 image columns = two 16-byte chunks
                // band-aware: a full-width step goes first so that every column of D is initialised by an overwriting MMA
                const int kc = band_aware ? (q == 0 ? 1 : q == 1 ? 0 : q) : q;
                const uint64_t ad = make_desc(smem_u32(sA) + (2 * kc) * kCH + dy * 16, kCH, 128);
                const uint32_t acc = (dy > 0 || q > 0) ? 1u : 0u;
                // candidates with taps in this K-step: j in [32 kc - 63, 32 kc + 31] -> kc 0: j 0..31 (a' 4..7, D columns 32..63),
                // kc 1, 2: all; kc 3: j 32..63 (a' 0..3, D columns 0..31)
                if (band_aware && kc == 0) mma_i8(tmem + 32, ad, make_desc(smem_u32(sB) + (4 * kc + 4) * 128, 256, 128), idesc32, acc);
                else if (band_aware && kc == 3) mma_i8(tmem, ad, make_desc(smem_u32(sB) + (4 * kc) * 128, 256, 128), idesc32, acc);
                else mma_i8(tmem, ad, make_desc(smem_u32(sB) + (4 * kc) * 128, 256, 128), idesc64, acc);
            }
            mma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        __syncthreads();
    }
    // epilogue: warp w owns TMEM lanes 32w .. 32w+31 (= candidate rows), 64 columns
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 16; ++i) {
            const int col = c0 + i, ap = col >> 3, s = col & 7;
            const int j = 8 * (kA - 1 - ap) + s;          // D column (a', s) holds candidate j
            out[(warp * 32 + lane) * 64 + j] = (int)v[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// (B) throughput: every SM issues `reps` x the MMA list (N values), operands wherever they fall in a 64 KB smem area
struct Pattern {
    int n;
    int nn[12];
    int distinct;   // != 0: consecutive MMAs alternate between two accumulators (no back-to-back dependency on one D)
};
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// The issue loop as the product kernel runs it: ONE WARP walks the MMA list, every operand is derived from warp-uniform
// values (kernel parameters, loop counters, one shuffled TMEM base) so that ptxas keeps descriptors in uniform registers and
// emits a bare UTCIMMA per step; a divergent `if (tid == 0)` issue loop costs ~160-240 cycles per MMA in R2UR waterfalls
// (first version of this probe), far above the N/2-cycle floor of the instruction.
__global__ void __launch_bounds__(128) k_probe_rate(int M, Pattern pat, int reps, int f16, long long* __restrict__ cycles, int variant)
{
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 96 * 1024);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = f16 ? 0x3C003C00u : 0x01010101u;
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    long long t0 = 0, t1 = 0;
    if (warp == 1) {
        const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
        const uint32_t a0 = __shfl_sync(0xffffffffu, smem_u32(sm), 0), b0 = a0 + 48 * 1024;
        const uint32_t ibase = f16 ? make_idesc_f16(M, 0) : make_idesc_i8(M, 0);
        t0 = clock64();
        if (variant & 4) {
            // straight-line issue: 8 MMAs per trip, descriptors advanced with uniform adds only, one N for all
            const uint32_t id = ibase | ((uint32_t)(pat.nn[0] >> 3) << 17);
            const uint32_t lay = (variant & 8) ? 2u : 0u;   // 8: SWIZZLE_128B layout type (timing only)
            uint64_t ad = make_desc(a0, 3072, 128) | ((uint64_t)lay << 61), bd = make_desc(b0, 256, 128) | ((uint64_t)lay << 61);
            for (int r = 0; r < reps * pat.n; r += 8) {
                if (elect_one()) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const uint32_t dcol = pat.distinct ? (uint32_t)((u & 1) * 256) : 0u;
                        if (f16 && (variant & 1)) mma_f16_nomask(tmem + dcol, ad + (uint64_t)(u & 3), bd + (uint64_t)(8 * (u & 3)), id, 1u);
                        else if (f16) mma_f16(tmem + dcol, ad + (uint64_t)(u & 3), bd + (uint64_t)(8 * (u & 3)), id, 1u);
                        else if (variant & 1) mma_i8_nomask(tmem + dcol, ad + (uint64_t)(u & 3), bd + (uint64_t)(8 * (u & 3)), id, 1u);
                        else mma_i8(tmem + dcol, ad + (uint64_t)(u & 3), bd + (uint64_t)(8 * (u & 3)), id, 1u);
                    }
                }
                __syncwarp();
            }
        } else
        for (int r = 0; r < reps; ++r) {
#pragma unroll 1
            for (int i = 0; i < pat.n; ++i) {
                const int N = pat.nn[i];                        // kernel parameter bank: uniform
                // A: rows at 16-byte stride, K chunks 3 KB apart (the image layout); B: aliased Toeplitz blocks
                const uint64_t ad = make_desc(a0 + (uint32_t)(((r + i) & 7) * 16) + (uint32_t)(2 * i) * 3072u, 3072, 128);
                const uint64_t bd = make_desc(b0 + (uint32_t)(4 * i) * 128u, 256, 128);
                const uint32_t id = ibase | ((uint32_t)(N >> 3) << 17);
                const uint32_t dcol = pat.distinct ? (uint32_t)((i & 1) * 256) : 0u;
                if (elect_one()) {
                    if (f16) mma_f16(tmem + dcol, ad, bd, id, 1u);
                    else mma_i8(tmem + dcol, ad, bd, id, 1u);
                }
            }
        }
        if (elect_one()) mma_commit(bar);
        __syncwarp();
    }
    if (!(variant & 2) || warp == 1) mbar_wait(bar, 0);
    if (tid == 32) {
        t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(512u) : "memory");
}

static double run_rate(int sms, int M, const Pattern& p, int reps, int f16, double* us_out, int variant = 0)
{
    long long* d_cyc;
    CK(cudaMalloc(&d_cyc, sizeof(long long) * sms));
    const size_t smem = 96 * 1024 + 64;
    CK(cudaFuncSetAttribute(k_probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k_probe_rate<<<sms, 128, smem>>>(M, p, reps, f16, d_cyc, variant);   // warm
    CK(cudaEventRecord(a));
    k_probe_rate<<<sms, 128, smem>>>(M, p, reps, f16, d_cyc, variant);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    std::vector<long long> h(sms);
    CK(cudaMemcpy(h.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double mx = 0;
    for (long long v : h) mx = v > mx ? v : mx;
    CK(cudaFree(d_cyc));
    if (us_out) *us_out = ms * 1e3;
    return mx / ((double)reps * p.n);
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s, %d SMs\n", prop.name, prop.multiProcessorCount);
    // ---- (A) correctness
    const int th = 64;
    std::vector<uint8_t> img(191 * 128);
    std::vector<int8_t> dig(th * 64);
    uint32_t s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
    for (auto& v : img) v = (uint8_t)(rnd() & 255);
    for (auto& v : dig) v = (int8_t)((int)(rnd() % 255) - 127);
    std::vector<int> want(128 * 64, 0);
    for (int y = 0; y < 128; ++y)
        for (int j = 0; j < 64; ++j) {
            long long a = 0;
            for (int dy = 0; dy < th; ++dy)
                for (int dx = 0; dx < 64; ++dx) a += (long long)img[(y + dy) * 128 + j + dx] * dig[dy * 64 + dx];
            want[y * 64 + j] = (int)a;
        }
    uint8_t* d_img; int8_t* d_dig; int* d_out;
    CK(cudaMalloc(&d_img, img.size())); CK(cudaMalloc(&d_dig, dig.size())); CK(cudaMalloc(&d_out, want.size() * 4));
    CK(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_dig, dig.data(), dig.size(), cudaMemcpyHostToDevice));
    const size_t smemA = kChunks * kCH + kBlocks * 128 + 64;
    CK(cudaFuncSetAttribute(k_probe_correct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
    for (int band = 0; band < 2; ++band) {
        CK(cudaMemset(d_out, 0xFF, want.size() * 4));
        k_probe_correct<<<1, 128, smemA>>>(d_img, d_dig, th, band, d_out);
        CK(cudaDeviceSynchronize());
        std::vector<int> got(want.size());
        CK(cudaMemcpy(got.data(), d_out, got.size() * 4, cudaMemcpyDeviceToHost));
        long long bad = 0;
        int first = -1;
        for (size_t i = 0; i < got.size(); ++i)
            if (got[i] != want[i]) { if (first < 0) first = (int)i; ++bad; }
        printf("[A] %s issue: %lld of %zu candidates differ from the CPU integer reference%s\n", band ? "band-aware" : "full-rectangle", bad,
               got.size(), bad ? "" : "  -> EXACT");
        if (bad) printf("    first mismatch at row %d candidate %d: got %d want %d\n", first / 64, first % 64, got[first], want[first]);
    }
    // ---- (B) cycles per MMA (max over SMs), all SMs busy
    const int sms = prop.multiProcessorCount, reps = 2000;
    printf("[B] cycles per tcgen05.mma (max over SMs, all SMs issuing); floor per B300_MICROARCH = N/2 for M = 128\n");
    struct { const char* name; int f16, variant, distinct, M; } vs[] = {
        {"i8  loop issue (elect, shuffled base)", 0, 0, 0, 128},
        {"i8  loop issue, waiters parked at bar.sync", 0, 2, 0, 128},
        {"i8  straight-line x8", 0, 4, 0, 128},
        {"i8  straight-line x8, waiters parked", 0, 6, 0, 128},
        {"i8  straight-line x8, parked, 2 accumulators", 0, 6, 1, 128},
        {"i8  straight-line x8, parked, NO-MASK form", 0, 7, 0, 128},
        {"i8  straight-line x8, parked, no-mask, 2 accum.", 0, 7, 1, 128},
        {"i8  straight-line x8, parked, no-mask, M=64", 0, 7, 0, 64},
        {"i8  straight-line x8, parked, M=64", 0, 6, 0, 64},
        {"i8  straight-line x8, parked, SW128 layout type", 0, 14, 0, 128},
        {"f16 straight-line x8, parked, mask form", 1, 6, 0, 128},
        {"f16 straight-line x8, parked, no-mask form", 1, 7, 0, 128},
    };
    for (auto& v : vs) {
        printf("  %-48s:", v.name);
        for (int N : {16, 32, 64, 96, 128, 256}) {
            Pattern p{8, {N, N, N, N, N, N, N, N}, v.distinct};
            printf("  N=%d: %.1f", N, run_rate(sms, v.M, p, reps / 4, v.f16, nullptr, v.variant));
        }
        printf("\n");
    }
    // the issue patterns of one (template row, digit): candidate block of 80 (5 K-steps) and of 176 (8 K-steps)
    struct { const char* name; Pattern p; double useful_cols; } pats[] = {
        {"block 80, band-aware", {5, {32, 64, 80, 48, 16}, 0}, 80 * 2.0},
        {"block 80, full rectangles", {5, {80, 80, 80, 80, 80}, 0}, 80 * 2.0},
        {"block 176, band-aware", {8, {32, 64, 96, 96, 96, 80, 48, 16}, 0}, 161 * 2.0},
        {"block 176, full rectangles", {8, {176, 176, 176, 176, 176, 176, 176, 176}, 0}, 161 * 2.0},
        {"block 64, band-aware", {4, {32, 64, 64, 32}, 0}, 64 * 2.0},
    };
    for (auto& q : pats) {
        double us = 0;
        const double c = run_rate(sms, 128, q.p, reps, 0, &us) * q.p.n;
        // useful MACs of the pattern: 128 rows x (useful candidates) x 64 taps; q.useful_cols = candidates * 64 / 32 K-steps' worth
        const double useful = 128.0 * q.useful_cols * 32.0;
        printf("  pattern %-28s: %.0f cycles per (row, digit) step, %.0f useful MAC/clk/SM per digit -> %.1f useful TMAC/s per digit on %d SMs @1.965 GHz (kernel %.0f us)\n",
               q.name, c, useful / c, useful / c * sms * 1.965e9 / 1e12, sms, us);
    }
    return 0;
}
