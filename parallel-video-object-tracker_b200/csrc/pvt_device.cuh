// Device-side state layout and helpers shared by all kernels of libpvt (sm_100a only).
//
// HBM layout per context (allocated once in pvt_create, no per-frame cudaMalloc -- the reference
// mallocs/frees three buffers per call, tracker/src/baseline_kernel.cu:340-358):
//   gray     [max_streams][H][pitch]  f32   toGrayF32 image of each stream's current frame (TMA source)
//   templ    [max_tracks][mth*mtw]    f32   the tracker's template, exactly the reference's templ_gray_f32
//   templc   [max_tracks][mth*mtp]    f32   fl32(templ - mean_t), chunk-major [x/8][y][x%8], columns zero-padded to 8
//   vsum/vsq [max_tracks][Hmax+mth][VW] f64 column prefix sums of f and f^2 over the search tile (scratch)
//   denom    [max_tracks][Hmax*Wmax]  f64   OpenCV's normaliser t = sqrt(diff2)*sigma_t*sqrt(N), 0 when flat
//   maps     [max_tracks][Hmax*Wmax]  f32   optional (keep_maps)
//   tracks   [max_tracks]             TrackState
//   table    [RING][max_streams]      FrameDesc   what each stream receives at each time step
//   results  [RING][max_tracks]       pvt_result
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pvt.h"

namespace pvt {

constexpr int kRing = 64;      // time steps that may be in flight
constexpr int kCX = 8;         // candidates per thread along x (contiguous)
constexpr int kCY = 5;         // candidates per thread along y (adjacent rows; odd => conflict-free LDS.128)

struct FrameDesc {
    const void* data;
    unsigned long long step;
    int format;
    int valid;
};

// which frame-table row a time step reads: row0 + (phase + step - step0) % ring_len.  Default {0, kRing, 0, 0, 0} = step % kRing
// (one row uploaded per step); a resident frame ring uploads its rows once and every later step needs no H2D at all.
struct SeqDesc {
    unsigned long long step0;
    int ring_len, row0;
    int prefetch;              // != 0: the ring's frames are pinned host memory -> k_prefetch_roi stages the next step's pixels
    int phase;                 // ring position of step0: row = row0 + (phase + step - step0) % ring_len (a caller that passes the
};                             // same ring from a later position re-uses the uploaded rows)

// What k_prefetch_roi staged for a track: gray f32 pixels of region [x0, x1) x [y0, y1) of the frame `data`, valid for `step`
struct StageHdr {
    int x0, y0, x1, y1;
    unsigned long long step;
    const void* data;
    int win[4];                // the track's search window at the START of the current step (stored by k_ingest_roi): what
                               // k_prefetch_roi grows -- it must not read the box itself, which the step's update moves
    unsigned long long cur_step;   // ... and the step it belongs to: k_prefetch_roi must not read *Ctx.step either (the update
};                                 // advances it while the low-priority prefetch branch may still be waiting for an SM)

struct TrackState {
    int active, stream;
    int x, y, w, h;            // bbox; (w, h) is also the template size (main.cpp never changes it)
    int tp;                    // centred-template row pitch in floats (w rounded up to 8)
    int flat;                  // 1: sigma_t^2 < DBL_EPSILON -> whole map is 1 (OpenCV common_matchTemplate); 2: Ctx.formula == EPS (plain quotient)
    double mean, templ_norm;   // mean_t, sigma_t * sqrt(N)   (EPS formula: (fl32(sigma_t + 1e-6) + 1e-6) * N)
    unsigned long long peak;   // packed (ordered score << 32 | ~index); 0 = empty
    int win[4];                // minTx, minTy, width, height of the current search window
    unsigned int ticket;       // CTAs of k_ncc_finalize / k_step_fused that are done with this track (last one runs the update)
    unsigned int arrive;       // k_step_fused: search CTAs of this track whose partial cross terms are stored (in-kernel barrier)
    unsigned int stats_done;   // k_step_fused: k_winstats CTAs of this track that have stored their normalisers in this step
    // lost-object re-acquisition (tracker_ghc/src/main.cpp:143-144, 183-239); only used when Ctx.lost_mode != 0
    int lost_count;            // consecutive frames below the confidence threshold (lost_frame_count)
    int use_global;            // use_global_search: the track is searched over the whole frame ...
    unsigned long long global_since;   // ... from this time step on (the flag is set by the PREVIOUS step's update)
    // tensor-core path (PVT_KERNEL_TC): the centred template in 16-bit fixed point q = rint(tc * 2^k) lives in Ctx.tdig
    double tc_inv;             // 2^-k
    double tc_dc;              // (sum(q) * 2^-k - sum(tc)) / N: the quantisation's DC part, taken out through the window sum
};

struct DevParams {
    int rx, ry;                // the global pass has its own copy with rx = W, ry = H (window == whole map) ...
    double min_conf, strong_conf, lr;   // ... and min_conf = NCC_GLOBAL_CONFIDENCE
    int keep_maps;
    int lost_threshold;        // LOST_FRAME_THRESHOLD (tracker_ghc/src/main.cpp:23)
};

// geometry + pointers every kernel needs; passed by value (fits in the parameter bank)
struct Ctx {
    int lost_mode;             // != 0: tracker_ghc semantics; a time step = local pass (global_pass 0) + global pass (1)
    int formula;               // pvt_formula, fixed at creation
    int global_pass;           // which pass this launch belongs to: a track is handled by exactly one pass per step
    int* stream_need;          // global pass: streams with at least one whole-frame track this step (whole-frame ingest)
    int W, H, pitch;           // frame geometry, gray-plane pitch in floats
    size_t plane;              // floats per gray plane
    int max_streams, max_tracks;
    int mtw, mth, mtp;         // template maxima: width, height, padded pitch
    int Wmax, Hmax, VW;        // window maxima (2*rx+1, 2*ry+1) and vsum row pitch
    float* gray;
    unsigned char* gray8;      // PVT_KERNEL_TC: the 8-bit gray plane itself [streams][H][pitch8] (operand A of the tensor-core search); else NULL
    int pitch8;                // bytes per row, multiple of 16
    size_t plane8;             // bytes per plane, multiple of 16
    signed char* tdig;         // PVT_KERNEL_TC: template digits [tracks][2][mth][tpp] (digit 0 = low byte), tpp = mtw rounded up to 16
    int tpp;
    double* wsum;              // PVT_KERNEL_TC: window sums [tracks][Hmax*Wmax] next to denom (k_rowsum)
    float* templ;
    float* templc;
    double* vsum;
    double* vsq;
    double* denom;
    float* maps;
    float* partial;            // [parts][max_tracks][tiles][8*kCY] K-split partial cross terms, tile-major
    float* fringe_acc;         // [max_tracks][parts][Hmax + Wmax] partial cross terms of the fringe column (by y) and row (by x), K-split mode
    float* stage;              // [max_tracks][stage_h][stage_w] next step's search-tile superset, staged by k_prefetch_roi (NULL: off)
    StageHdr* stage_hdr;       // [max_tracks]
    int stage_w, stage_h;
    TrackState* tracks;
    FrameDesc* table;
    SeqDesc* seq;
    pvt_result* results;
    DevParams* params;
    unsigned long long* step;  // device-side time-step counter (advanced by the update kernel)
    unsigned int* ticket;      // tracks whose update is done in this step (the last one advances the step counter)
    unsigned long long* macs;  // algorithmic MACs searched so far (n_cand * tw * th per track per step)
    unsigned long long* macs_grid;  // the share of them inside k_ncc_search's thread-tile grid (gridW x gridH candidates)
    int gridW, gridH;          // 8 * TileCfg.C, kCY * TileCfg.G
    unsigned int* fault;       // mapped host word: a bounded device-side wait gave up (k_step_fused) -> the next host sync returns PVT_ERR_CUDA
    unsigned long long* trace; // optional [kRing][8 kernels][2] globaltimer stamps (first CTA start, last CTA end); NULL = off
};

// tracker/src/main.cpp:135-146, same int arithmetic (all operands >= 0, so / truncates like C)
__host__ __device__ inline void search_window(int x, int y, int w, int h, int outW, int outH, int rx, int ry, int* win)
{
    int cx = x + w / 2, cy = y + h / 2;
    int minTx = cx - rx - w / 2; if (minTx < 0) minTx = 0;
    int maxTx = cx + rx - w / 2; if (maxTx > outW - 1) maxTx = outW - 1;
    int minTy = cy - ry - h / 2; if (minTy < 0) minTy = 0;
    int maxTy = cy + ry - h / 2; if (maxTy > outH - 1) maxTy = outH - 1;
    win[0] = minTx; win[1] = minTy; win[2] = maxTx - minTx + 1; win[3] = maxTy - minTy + 1;
}

__device__ __forceinline__ size_t table_row(const Ctx& c, unsigned long long step)
{
    const SeqDesc q = *c.seq;
    return (size_t)(q.row0 + (int)((step - q.step0 + (unsigned long long)q.phase) % (unsigned long long)q.ring_len)) * c.max_streams;
}
// whole-frame (global) search applies to this track at this step (decided by the previous step's update)
__device__ __forceinline__ bool track_global(const TrackState& t, unsigned long long step)
{
    return t.active && t.use_global && t.global_since <= step;
}
// does this pass own the track at this step?  (inactive tracks belong to the local pass, which reports them)
__device__ __forceinline__ bool track_owned(const Ctx& c, const TrackState& t, unsigned long long step)
{
    return !c.lost_mode || (track_global(t, step) == (c.global_pass != 0));
}
__device__ __forceinline__ bool track_stepped(const Ctx& c, const TrackState& t, unsigned long long step)
{
    if (!t.active || !track_owned(c, t, step)) return false;
    return c.table[table_row(c, step) + t.stream].valid != 0;
}

// track_stepped with every independent load issued before the first branch: *step, the sequence descriptor and the track's
// fields go out together (one L2 round trip), the frame-table entry follows (a second one).  The branchy form above walks
// t.active -> *step -> lost-mode fields -> SeqDesc -> table[..].valid as a chain of up to five dependent round trips -- on the
// single-stream step that is 1 - 2 us in front of every kernel.  (An inactive slot has stream 0: the speculative table read is in range.)
__device__ __forceinline__ bool track_stepped_ld(const Ctx& c, const TrackState& t, unsigned long long& step_out, FrameDesc* fd = nullptr)
{
    const unsigned long long s = *c.step;
    const SeqDesc q = *c.seq;
    const int active = t.active, stream = t.stream, ug = t.use_global;
    const unsigned long long gs = t.global_since;
    step_out = s;
    const size_t row = (size_t)(q.row0 + (int)((s - q.step0 + (unsigned long long)q.phase) % (unsigned long long)q.ring_len)) * c.max_streams;
    int valid;
    if (fd) { *fd = c.table[row + stream]; valid = fd->valid; }
    else valid = c.table[row + stream].valid;
    const bool glob = active && ug && gs <= s;
    const bool owned = !c.lost_mode || (glob == (c.global_pass != 0));
    return active && owned && valid != 0;
}

// monotone float -> uint map (larger float <=> larger uint); -0 is folded into +0 first so that it
// ties with +0 exactly like a float comparison in cv::minMaxLoc
__device__ __forceinline__ unsigned int ord_f32(float v)
{
    unsigned int u = __float_as_uint(v + 0.0f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord_f32(unsigned int o)
{
    unsigned int u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}
// (score desc, row-major index asc): max over keys == first maximum in row-major order
__device__ __forceinline__ unsigned long long peak_key(float v, unsigned int idx)
{
    return ((unsigned long long)ord_f32(v) << 32) | (unsigned long long)(0xffffffffu - idx);
}

// ---- optional device-side timeline (pvt_trace_enable): first-CTA start and last-CTA end of every kernel, per step
enum { TR_INGEST = 0, TR_COLPREFIX = 1, TR_ROWSUM = 2, TR_NCC = 3, TR_FINALIZE = 4, TR_UPDATE = 5, TR_FRINGE = 6, TR_TAIL = 7 };
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_begin(const Ctx& c, unsigned long long step, int k)
{
    if (c.trace && threadIdx.x == 0 && threadIdx.y == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
        c.trace[((step % kRing) * 8 + k) * 2] = gtime();
}
__device__ __forceinline__ void trace_end(const Ctx& c, unsigned long long step, int k)
{
    if (c.trace && threadIdx.x == 0 && threadIdx.y == 0) atomicMax(&c.trace[((step % kRing) * 8 + k) * 2 + 1], gtime());
}

// ---- programmatic dependent launch ---------------------------------------------------------
// pdl_trigger: kernels launched behind this one with the programmatic attribute may start now (they still block in
// pdl_wait until this grid has completed and its writes are visible).  Both are no-ops for plain launches.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- PTX helpers: mbarrier, TMA tensor load, bulk copy -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x2000;\n\t"   // suspend-time hint (ns): sleep in HW, no hot spin
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 3-D tile (x, y, plane) of the gray pool -> dense [boxH][boxW] in shared memory; out-of-range
// elements are zero-filled by the TMA unit
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// contiguous global -> shared, bytes % 16 == 0, both 16-byte aligned
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m)
{
    unsigned int lo = (unsigned int)v, hi = (unsigned int)(v >> 32);
    lo = __shfl_xor_sync(0xffffffffu, lo, m);
    hi = __shfl_xor_sync(0xffffffffu, hi, m);
    return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ double shfl_up_f64(double v, int d)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(0xffffffffu, lo, d);
    hi = __shfl_up_sync(0xffffffffu, hi, d);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_f64(double v, int src)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return __hiloint2double(hi, lo);
}

}  // namespace pvt
