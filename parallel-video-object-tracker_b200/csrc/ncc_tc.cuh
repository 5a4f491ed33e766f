// k_ncc_tc -- the NCC cross term on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
// PVT_KERNEL_TC; BASELINE.json north_star item 3 ("a tcgen05 implicit-GEMM variant only if ncu shows the FP32 path is
// compute-bound and the tolerance gate holds": profiles/ncc_search_C4_r1.txt has the FMA pipe at 80.6 %, tools/tc_emulate.py
// and tests/ hold the gate).  Reference semantics unchanged: tracker/src/ncc_cpu.cpp:12 (TM_CCOEFF_NORMED), main.cpp:135-151.
//
// Formulation.  For one template row dy the cross term of all candidates (y, j) of a window is a GEMM
//       D[y][j] += sum_k G[y + dy][k] * T_dy[k][j],      T_dy[k][j] = q[dy][k - j - o]  for 0 <= k - j - o < tw, else 0
//   A = the 8-bit gray levels themselves (toGrayF32's input, utils.hpp:8: f = fl32(g * fl32(1/255))), UNSIGNED 8 bit: exact;
//   B = the centred template in 16-bit fixed point, q = rint(tc * 2^k), as TWO signed 8-bit digits q = 256 d1 + d0: two GEMMs
//       into two int32 accumulators; D is EXACT integer arithmetic, independent of summation order -- equal windows score
//       equal bits wherever they sit (exact ties stay exact), which no floating-point tensor-core mode could promise;
//   cross = c255 * 2^-k * (256 D1 + D0) - (wsum / N) * dc, with dc = sum(q) 2^-k - sum(tc): the DC part of the quantisation
//       error is removed through the window sum the statistics kernels already hold in FP64; what is left is
//       sum (g - mean_w) * delta, which scales with the window's contrast like the score's denominator does.
//       Measured against the cv2 goldens (tools/tc_emulate.py i8x2, and the GPU tests): <= 4e-5 on scores, same peaks.
// Layouts (tools/tc_probe.cu established them on B200, bit-exact against a CPU integer reference):
//   * image tile in shared memory, 16-pixel-chunk-major: addr(row, chunk) = chunk * (rows * 16) + row * 16.  An 8-row core
//     matrix of the no-swizzle K-major canonical layout is then 128 contiguous bytes for ANY first row: the row shift dy is
//     +1 in the descriptor's 16-byte start-address field.  ONE 4-D TMA copy (16 B, row, chunk, stream) of the u8 gray plane
//     produces exactly this layout; the window origin is rounded down to 16 pixels and the remainder o goes into T_dy.
//   * the Toeplitz operand is never materialised: its core matrix (8 candidates, 16 image columns) only depends on
//     2 m - a (m = 16-column index, a = candidate group), so with the groups of a block enumerated in reverse it is
//     addr(a', m) = base + (2 m + a') * 128: stride-byte-offset 128, leading-byte-offset 256 over ~12 non-zero 128-byte
//     blocks per (template row, digit) that alias each other, inside a zero-filled run of blocks.
//   * band-aware issue: a 32-column K-step only meets the candidates whose taps it holds; the MMA is issued for that
//     sub-range (N' = 16..96 of 176 columns) through a column offset into D and a block offset into B.
//   * column tiles: a window wider than one accumulator (TcCfg.XW candidate columns, <= 256) -- 4K windows, the whole-frame pass of
//     the lost-object mode (tracker_ghc/src/main.cpp:186-193) -- is cut into TcCfg.xtiles tiles; a CTA = (track, 128 rows, XW
//     columns) treats its tile as a window of its own (origin win[0] + xt * XW: own alignment o, own band, partial last tile) and
//     only the epilogue's indices into the window's maps use the full row length.
// Roles in the CTA (9 warps): warp 0 = TMA + MMA issue (one elected lane; every operand warp-uniform, see tc_probe.cu for
// what a divergent issue loop costs), warps 1-8 build the Toeplitz blocks of the coming template rows into a ring of stages
// (full / empty mbarriers; empty is signalled by tcgen05.commit) and afterwards run the epilogue: TMEM -> registers (a warp
// reads its lane quadrant, two warps per quadrant take alternate 16-column chunks), the window statistics of the chunk
// staged through shared memory with coalesced cp.async (a lane owns a candidate ROW: read directly, every load would touch
// 32 lines), OpenCV's normalisation in FP64 as 16 independent branch-free chains per thread, (score, index) key, warp max,
// atomicMax.  (Round 2 timeline of one CTA, C5: prologue 3 us | 64 template rows of MMAs 48 us | epilogue 79 us with the
// first version -- one warp per scheduler, a branchy FP64 chain and two uncoalesced loads per candidate.)
#pragma once
#include "pvt_device.cuh"

namespace pvt {

constexpr int kTcKMax = 10;        // K-steps (of 32 image columns) the issue code is unrolled for: Wmax + tw + 14 <= 320
constexpr int kTcThreads = 288;       // warp 0: TMA + MMA issue; warps 1-8: Toeplitz producers, then the epilogue (two warps per TMEM lane quadrant)
constexpr int kTcEpiBytes = 8 * 2 * 32 * 17 * 8;   // epilogue staging: per warp [denom | wsum][32 rows][16 candidates + 1 pad] doubles, over the dead main-loop buffers
constexpr int kTcPadL = 16, kTcPadR = 48;   // zero bytes left / right of a digit row in shared memory

struct TcCfg {
    int AG;          // candidate groups of 8 in an accumulator: XW / 8
    int KS;          // K-steps: ceil((XW + mtw - 1 + 15) / 32)
    int rows;        // image-tile rows = 128 + mth - 1 (TMA box)
    int nblk;        // Toeplitz blocks per (stage, digit): 4 KS + AG (band blocks inside a run of zero blocks)
    int stages;      // depth of the block ring
    int tpp;         // digit-row pitch in global memory (mtw rounded up to 16)
    int mtiles;      // ceil(Hmax / 128)
    int tmem_cols;   // power of two >= 2 * 8 * AG
    int spin;        // 1: the ring's waits poll (mbarrier.test_wait) instead of suspending in try_wait
    int XW;          // candidate columns per accumulator (multiple of 16, <= 256): Wmax rounded up to 16 when one accumulator covers the
                     // window, else the window is cut into xtiles column tiles of XW candidates (4K windows, the whole-frame pass)
    int xtiles;      // ceil(Wmax / XW); a CTA = (track, 128-row tile, column tile)
};

// ring waits: polling variant (the suspending try_wait of pvt_device.cuh wakes late when the producer / consumer chain is short)
__device__ __forceinline__ void mbar_wait_poll(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITP_%=:\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONEP_%=;\n\t"
        "bra WAITP_%=;\n\t"
        "DONEP_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity, int spin)
{
    if (spin) mbar_wait_poll(bar, parity);
    else mbar_wait(bar, parity);
}

// ---- tcgen05 / TMA wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// no-swizzle K-major shared-memory matrix descriptor: start >> 4 at [0,14), LBO >> 4 at [16,30), SBO >> 4 at [32,46), version 1 at [46,48)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// kind::i8 instruction descriptor: D = S32, A = unsigned 8 bit, B = signed 8 bit, K-major both, N >> 3 at [17,23), M = 128 (>> 4 at [24,29))
__device__ __forceinline__ uint32_t tc_idesc(int N) { return (2u << 4) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24); }
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void cp_async8(void* dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// num / t for normal operands without the library routine's slow-path branch (which keeps the compiler from interleaving the
// 16 independent chains of an epilogue chunk): reciprocal seed, two Newton steps, quotient with one residual correction
__device__ __forceinline__ double div_nr(double num, double t)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(t));
    double e = __fma_rn(-t, y, 1.0);
    y = __fma_rn(y, e, y);
    e = __fma_rn(-t, y, 1.0);
    y = __fma_rn(y, e, y);
    const double q = __dmul_rn(num, y);
    return __fma_rn(__fma_rn(-t, q, num), y, q);
}
// ncc_finalize's rule (OpenCV common_matchTemplate, TM_CCOEFF_NORMED) as selects: |num| < t -> num / t; < 1.125 t -> +-1; else 0
__device__ __forceinline__ float ncc_finalize_sel(double num, double t)
{
    const double an = fabs(num);
    const double q = div_nr(num, t > 0.0 ? t : 1.0);
    const double one = num > 0.0 ? 1.0 : -1.0;
    const double r = an < t ? q : (an < __dmul_rn(t, 1.125) ? one : 0.0);
    return (float)r;
}

// candidates whose taps meet K-step kc (image-tile columns [32 kc, 32 kc + 32)), as a range of REVERSED 8-candidate groups
// aligned to 16 columns: first group a0 (even) and group count n8 (even; 0 = the step holds no tap of any candidate)
__device__ __forceinline__ void tc_band(int kc, int o, int tw, int ww, int AG, int* a0, int* n8)
{
    int jlo = 32 * kc - tw + 1 - o, jhi = 32 * kc + 31 - o;
    if (jlo < 0) jlo = 0;
    if (jhi > ww - 1) jhi = ww - 1;
    if (jhi < jlo) { *a0 = 0; *n8 = 0; return; }
    int lo = AG - 1 - (jhi >> 3), hi = AG - 1 - (jlo >> 3);   // reversed group indices, lo <= hi
    lo &= ~1;
    hi |= 1;
    *a0 = lo;
    *n8 = hi - lo + 1;
}

__global__ void __launch_bounds__(kTcThreads, 1) k_ncc_tc(Ctx c, TcCfg g, const __grid_constant__ CUtensorMap tmap8)
{
    extern __shared__ __align__(1024) unsigned char sm_tc[];
    const int per_track = g.mtiles * g.xtiles;
    const int track = blockIdx.x / per_track, bt = blockIdx.x - track * per_track;
    const int mt = bt / g.xtiles, xt = bt - mt * g.xtiles;
    TrackState& t = c.tracks[track];
    const unsigned long long step = *c.step;
    if (!track_stepped(c, t, step)) return;
    const int wfull = t.win[2], wh = t.win[3], row0 = mt * 128;
    const int jx0 = xt * g.XW;                                // first candidate column of this column tile
    if (row0 >= wh || jx0 >= wfull) return;                   // clamped window: no such row / column tile (CTA-uniform)
    // from here on the CTA sees its column tile as a window of its own: origin win[0] + jx0, ww candidates wide; only the
    // epilogue's indices into the window's maps (statistics, scores, peak key) use the full row length wfull
    const int ww = min(g.XW, wfull - jx0);
    trace_begin(c, step, TR_NCC);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int th = t.h, tw = t.w, ox = t.win[0] + jx0, o = ox & 15;
    const int CH = g.rows * 16;                               // bytes per 16-pixel chunk of the image tile
    const int blk_bytes = g.nblk * 128;                       // one (stage, digit) run of Toeplitz blocks
    const int drow = kTcPadL + g.tpp + kTcPadR;               // digit row pitch in shared memory

    unsigned char* sA = sm_tc;
    unsigned char* sB = sA + (size_t)CH * 2 * g.KS;
    unsigned char* sD = sB + (size_t)g.stages * 2 * blk_bytes;               // digits [2][mth][drow]
    size_t main_bytes = (size_t)(sD - sm_tc) + (((size_t)2 * c.mth * drow + 15) & ~(size_t)15);
    if (main_bytes < (size_t)kTcEpiBytes) main_bytes = kTcEpiBytes;          // the epilogue staging reuses [0, kTcEpiBytes)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm_tc + main_bytes);
    uint64_t* bar_tile = bars;                                 // [0] image tile landed
    uint64_t* bar_done = bars + 1;                             // [1] all MMAs complete
    uint64_t* full = bars + 2;                                 // [stages] Toeplitz blocks of a template row written
    uint64_t* empty = full + g.stages;                         // [stages] ... and consumed (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty + g.stages);

    if (warp == 0) {
        if (lane == 0) {
            mbar_init(bar_tile, 1);
            mbar_init(bar_done, 1);
            for (int s = 0; s < g.stages; ++s) { mbar_init(&full[s], (kTcThreads - 32) / 32); mbar_init(&empty[s], 1); }
            fence_mbar_init();
            mbar_arrive_expect_tx(bar_tile, (uint32_t)(CH * 2 * g.KS));
            tma_load_4d(sA, &tmap8, bar_tile, 0, t.win[1] + row0, ox >> 4, t.stream);
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)g.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        // zero the block ring (the zero blocks around the band are never written again) and stage the two digit planes with
        // zero margins: a Toeplitz block row is then one unaligned 16-byte window of a digit row
        const int p = tid - 32, np = kTcThreads - 32;
        uint4* zb = reinterpret_cast<uint4*>(sB);
        for (int i = p; i < g.stages * 2 * blk_bytes / 16; i += np) zb[i] = make_uint4(0u, 0u, 0u, 0u);
        uint32_t* zd = reinterpret_cast<uint32_t*>(sD);
        for (int i = p; i < 2 * th * drow / 4; i += np) zd[i] = 0u;
    }
    __syncthreads();
    if (warp != 0) {
        const int p = tid - 32, np = kTcThreads - 32;
        const unsigned char* gd = reinterpret_cast<const unsigned char*>(c.tdig) + (size_t)track * 2 * c.mth * g.tpp;
        const int w4 = g.tpp >> 2;
        for (int i = p; i < 2 * th * w4; i += np) {
            const int dg = i / (th * w4), r = (i - dg * th * w4) / w4, x4 = i - (dg * th + r) * w4;
            *reinterpret_cast<uint32_t*>(sD + (size_t)(dg * th + r) * drow + kTcPadL + 4 * x4) =
                __ldg(reinterpret_cast<const uint32_t*>(gd + ((size_t)dg * c.mth + r) * g.tpp) + x4);
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // phase stamps of CTA 0 (pvt_trace_enable; tools/tc_timeline.py): FINALIZE slot = prologue done | MMAs all issued,
    // FRINGE slot = image tile landed | , TAIL slot = all MMAs complete | epilogue done
    unsigned long long* trc = (c.trace && blockIdx.x == 0) ? c.trace + (step % kRing) * 16 : nullptr;
    if (trc && tid == 0) trc[TR_FINALIZE * 2] = gtime();

    // band of non-zero Toeplitz blocks: block d holds taps 8 (d - AG + 1) - s - o + [0, 16), s = 0..7
    // (a block is non-zero iff its s = 0 row reaches tap 0 and its s = 7 row starts at or before tap tw - 1)
    const int dlo_c = g.AG - 2 + (o >> 3);
    int dhi = g.AG - 1 + ((tw + 6 + o) >> 3);
    if (dhi > g.nblk - 1) dhi = g.nblk - 1;
    const int nb = dhi - dlo_c + 1;

    if (warp == 0) {
        // ===== MMA issue =====
        const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
        const uint32_t sa = __shfl_sync(0xffffffffu, smem_u32(sA), 0), sb = __shfl_sync(0xffffffffu, smem_u32(sB), 0);
        const int u_o = __shfl_sync(0xffffffffu, o, 0), u_tw = __shfl_sync(0xffffffffu, tw, 0), u_ww = __shfl_sync(0xffffffffu, ww, 0);
        const int u_th = __shfl_sync(0xffffffffu, th, 0);
        const int NW = g.AG * 8;
        // per K-step: A descriptor (template row 0), B descriptor offset (16-byte units from a stage's start), D column, idesc.
        // BOTH DIGITS IN ONE MMA: the blocks of the two digits are interleaved in the ring (block d of digit dg at 2 d + dg), so
        // the N index a'' = 2 a' + dg walks candidate group a', digit dg at the same 128-byte stride (SBO 128) and a 16-column
        // K block advances 4 blocks (LBO 512); the accumulator holds (group, digit, 8 candidates) = 16 a' + 8 dg + s.  One MMA of
        // N = 16 n8 instead of two of N = 8 n8: the instruction's fixed cost (~94 - 120 cycles for any N <= 128, tools/tc_probe.cu) is
        // paid half as often.
        // (a band of more than 16 candidate groups -- templates wider than ~96 pixels -- exceeds N = 256 and is issued in two segments)
        uint64_t ad[kTcKMax];
        uint32_t boff[kTcKMax][2], dcol[kTcKMax][2], idn[kTcKMax][2];
        bool on[kTcKMax][2];
#pragma unroll
        for (int kc = 0; kc < kTcKMax; ++kc) {
            int a0 = 0, n8 = 0;
            if (kc < g.KS) tc_band(kc, u_o, u_tw, u_ww, g.AG, &a0, &n8);
            ad[kc] = tc_desc(sa + (uint32_t)(2 * kc) * (uint32_t)CH, (uint32_t)CH, 128);
#pragma unroll
            for (int sg = 0; sg < 2; ++sg) {
                const int b0 = a0 + 16 * sg, nn = min(16, n8 - 16 * sg);
                on[kc][sg] = nn > 0;
                boff[kc][sg] = (uint32_t)(8 * kc + 2 * b0) * 8u;
                dcol[kc][sg] = (uint32_t)b0 * 16u;
                idn[kc][sg] = tc_idesc(max(nn, 2) * 16);
            }
        }
        const uint64_t bd0 = tc_desc(sb, 512, 128);
        const uint32_t id_half = tc_idesc(NW);                  // first template row: AG / 2 groups x 2 digits x 8 = NW columns per half
        mbar_wait(bar_tile, 0);
        if (trc && lane == 0) trc[TR_FRINGE * 2] = gtime();
        for (int dy = 0; dy < u_th; ++dy) {
            const int st = dy % g.stages;
            tc_wait(&full[st], (uint32_t)(dy / g.stages) & 1u, g.spin);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t bst = bd0 + (uint64_t)((uint32_t)(st * 2 * blk_bytes) >> 4);
                if (dy == 0) {
                    // first template row: full-width rectangles (in two halves of AG / 2 candidate groups: N <= 256), the first
                    // one of each half overwrites -> every accumulator column is initialised (zero blocks outside the band
                    // contribute nothing)
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                        for (int kc = 0; kc < kTcKMax; ++kc)
                            if (kc < g.KS)
                                tc_mma(tmem + (uint32_t)(hf * NW), ad[kc], bst + (uint64_t)((8 * kc + hf * g.AG) * 8), id_half, kc > 0 ? 1u : 0u);
                } else {
#pragma unroll
                    for (int kc = 0; kc < kTcKMax; ++kc)
#pragma unroll
                        for (int sg = 0; sg < 2; ++sg)
                            if (on[kc][sg]) tc_mma(tmem + dcol[kc][sg], ad[kc] + (uint64_t)dy, bst + (uint64_t)boff[kc][sg], idn[kc][sg], 1u);
                }
                tc_commit(&empty[st]);                         // arrives when the MMAs that read this stage have completed
                if (dy == u_th - 1) tc_commit(bar_done);
            }
            __syncwarp();
        }
        if (trc && lane == 0) trc[TR_FINALIZE * 2 + 1] = gtime();
    } else {
        // ===== Toeplitz block producers (128 threads) =====
        const int p = tid - 32;
        for (int dy = 0; dy < th; ++dy) {
            const int st = dy % g.stages;
            if (dy >= g.stages) tc_wait(&empty[st], (uint32_t)(dy / g.stages - 1) & 1u, g.spin);
            for (int i = p; i < 2 * nb * 8; i += kTcThreads - 32) {
                const int dg = i / (nb * 8), r = i - dg * nb * 8;
                const int d = dlo_c + (r >> 3), s = r & 7;
                const int off = kTcPadL + 8 * (d - g.AG + 1) - s - o;            // first tap of this block row, in the padded digit row
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (off >= 0 && off + 16 <= drow) {
                    const uint32_t* w = reinterpret_cast<const uint32_t*>(sD + (size_t)(dg * th + dy) * drow) + (off >> 2);
                    const uint32_t sh = (uint32_t)(off & 3) * 8u;
                    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
                    v = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
                }
                *reinterpret_cast<uint4*>(sB + (size_t)(st * 2) * blk_bytes + (size_t)(2 * d + dg) * 128 + s * 16) = v;   // digits interleaved
            }
            fence_async_smem();                                // generic-proxy writes -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[st]);
        }
        // ===== epilogue: TMEM lane = candidate row; warp w may read lanes 32 (w % 4) .. + 31 =====
        mbar_wait(bar_done, 0);
        tc_fence_after();
        if (trc && tid == 32) trc[TR_TAIL * 2] = gtime();
        const uint32_t tmem = *tmem_slot;
        const int e = warp - 1, q4 = warp & 3, half = e >> 2;       // (quadrant, half) pairs are covered once by warps 1 .. 8
        const int y = row0 + q4 * 32 + lane;
        const bool rowok = y < wh;
        double* stg = reinterpret_cast<double*>(sm_tc) + (size_t)e * (2 * 32 * 17);   // [denom | wsum][row of the warp][17]
        const size_t woff = (size_t)track * c.Hmax * c.Wmax + (size_t)jx0;
        const double* dnb = c.denom + woff;
        const double* wsb = c.wsum + woff;
        float* mp = (c.params->keep_maps && rowok) ? c.maps + woff + (size_t)y * wfull : nullptr;
        const int flat = t.flat;
        const double sc = (double)(1.0f / 255.0f) * t.tc_inv, dc = t.tc_dc;     // fl32(1/255): the ingest's own constant (utils.hpp:12)
        const int NW = g.AG * 8;
        unsigned long long key = 0ull;
        for (int c0 = half * 16; c0 < NW; c0 += 32) {
            // accumulator columns [c0, c0 + 16) = reversed candidate groups ga, ga - 1  ->  candidates [jb, jb + 16), jb = 8 (ga - 1)
            const int jb = 8 * (g.AG - 2 - (c0 >> 3));
            {
                const int jc = min(jb + (lane & 15), ww - 1);
#pragma unroll
                for (int i = 0; i < 16; ++i) {                      // a warp instruction covers two rows x 16 candidates (128 B each)
                    const int r = 2 * i + (lane >> 4);
                    const size_t o = (size_t)min(row0 + q4 * 32 + r, wh - 1) * wfull + jc;
                    cp_async8(stg + r * 17 + (lane & 15), dnb + o);
                    cp_async8(stg + (32 + r) * 17 + (lane & 15), wsb + o);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            uint32_t hi[16], lo[16];
            {
                // accumulator columns: group a' = c0 / 8 at [16 a', 16 a' + 16) = 8 x digit 0, then 8 x digit 1; the next group follows
                uint32_t v1[16], v2[16];
                const uint32_t ta = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(2 * c0);
                tc_ld16(ta, v1);
                tc_ld16(ta + 16u, v2);
                tc_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) { lo[i] = v1[i]; hi[i] = v1[8 + i]; lo[8 + i] = v2[i]; hi[8 + i] = v2[8 + i]; }
            }
            tc_ld_wait();
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            double num[16], dn[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int jl = i < 8 ? 8 + i : i - 8;               // accumulator column -> position in the staged candidate range
                const long long acc = (long long)(int)hi[i] * 256 + (long long)(int)lo[i];
                const double cross = __dsub_rn(__dmul_rn((double)acc, sc), __dmul_rn(stg[(32 + lane) * 17 + jl], dc));
                num[i] = (double)(float)cross;                      // the cross term is a float32 in cv::matchTemplate
                dn[i] = stg[lane * 17 + jl];
            }
            float v[16];
            if (flat == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = ncc_finalize_sel(num[i], dn[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = flat == 1 ? 1.0f : (float)(num[i] / dn[i]);   // 2: PVT_FORMULA_EPS, t > 0 always
            }
            if (rowok) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int j = jb + (i < 8 ? 8 + i : i - 8);
                    if (j < ww) {
                        if (mp) mp[j] = v[i];
                        const unsigned long long k2 = peak_key(v[i], (unsigned int)(y * wfull + jx0 + j));
                        key = k2 > key ? k2 : key;
                    }
                }
            }
            __syncwarp();                                          // the next chunk overwrites the staging rows
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            unsigned long long ok = shfl_xor_u64(key, m);
            key = ok > key ? ok : key;
        }
        if (lane == 0 && key) atomicMax(&t.peak, key);
        if (trc && tid == 32) trc[TR_TAIL * 2 + 1] = gtime();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"((uint32_t)g.tmem_cols) : "memory");
    trace_end(c, step, TR_NCC);
}

}  // namespace pvt
