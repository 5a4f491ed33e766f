// Hand-written sm_100a kernels of the NCC tracking hot path.  One time step = the launch sequence
//   many tracks:   k_ingest[_roi] -> k_winstats -> k_ncc_search ~> k_ncc_fringe [-> k_ncc_tail_finalize] -> k_update      (or k_ncc_tc: ncc_tc.cuh)
//   one stream:    k_ingest_roi ~> { k_ncc_local || k_winstats } ~> k_update
//   in between:    k_ingest_roi ~> k_ncc_search (K-split) || k_winstats -> k_ncc_finalize (its last CTA runs the update)
// (~> = programmatic dependent launch) captured once as a CUDA graph; every kernel finds "which frame / which step" through the
// device-side step counter and frame table, so the graph is launched unchanged for every frame and
// nothing returns to the host between frames.
#pragma once
#include <float.h>

#include "pvt_device.cuh"

namespace pvt {

// defined in section (5), used earlier
__device__ void write_digits(const Ctx& c, int track, TrackState& t, const float* s_t, double mean, int tw, int th);
__device__ void finish_template(const Ctx& c, int track, TrackState& t, const float* s_t, double* red, int tw, int th, bool have_sums = false,
                                double ps = 0.0, double pq = 0.0);
__device__ void track_update(const Ctx& c, int track, unsigned long long step, bool stepped, float* s_t, double* red);

// =============================================================================================
// (1) ingest: BGR u8 -> gray -> f32/255          reference: tracker/include/utils.hpp:5-14
//     cvtColor(BGR2GRAY) 8u: (3735 B + 19235 G + 9798 R + 16384) >> 15   (bit-exact, tests G6)
//     convertTo(CV_32F, 1.0f/255.0f): one rounding of g * 0x1.010102p-8
// HBM-bound: 3 B read + 4 B written per pixel.  Each thread converts 4 pixels: three aligned 32-bit
// loads (a warp reads 384 contiguous bytes) and one 16-byte store (a warp writes 512 contiguous bytes).
// =============================================================================================
__device__ __forceinline__ float gray_to_f32(unsigned int g) { return __fmul_rn((float)g, 1.0f / 255.0f); }
__device__ __forceinline__ unsigned int bgr_to_gray(unsigned int b, unsigned int g, unsigned int r)
{
    return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

// 4 consecutive pixels of one source row -> 4 toGrayF32 values (n < 4 at the right frame edge)
__device__ __forceinline__ float4 ingest_group(const FrameDesc& d, const unsigned char* row, int x, int n)
{
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d.format == PVT_FMT_BGR8) {
        const unsigned char* p = row + 3 * x;
        if (n == 4 && ((((size_t)p) & 3) == 0)) {
            const unsigned int* p32 = (const unsigned int*)p;
            unsigned int a = __ldg(p32), b = __ldg(p32 + 1), e = __ldg(p32 + 2);
            // bytes: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
            o.x = gray_to_f32(bgr_to_gray(a & 255u, (a >> 8) & 255u, (a >> 16) & 255u));
            o.y = gray_to_f32(bgr_to_gray(a >> 24, b & 255u, (b >> 8) & 255u));
            o.z = gray_to_f32(bgr_to_gray((b >> 16) & 255u, b >> 24, e & 255u));
            o.w = gray_to_f32(bgr_to_gray((e >> 8) & 255u, (e >> 16) & 255u, e >> 24));
        } else {
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            for (int i = 0; i < n; ++i) v[i] = gray_to_f32(bgr_to_gray(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
            o = make_float4(v[0], v[1], v[2], v[3]);
        }
    } else if (d.format == PVT_FMT_GRAY8) {
        const unsigned char* p = row + x;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < n; ++i) v[i] = gray_to_f32(p[i]);
        o = make_float4(v[0], v[1], v[2], v[3]);
    } else {  // PVT_FMT_GRAYF32: already toGrayF32 output, re-pitch into the pool
        const float* p = (const float*)row + x;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < n; ++i) v[i] = p[i];
        o = make_float4(v[0], v[1], v[2], v[3]);
    }
    return o;
}

// Full-frame ingest.  HBM-bound: 3 B read + 4 B written per pixel = 14.5 MB per 1080p frame.
// Grid = (column blocks, rows, streams): no index division anywhere (round 1 spent more instructions on a 64-bit
// gid / groups-per-row than on the conversion: ncu had the kernel issue-bound at 80 % of the slots, 78 % of the copy rate).
// Fast path (BGR8, W % 16 == 0, 16-byte aligned rows -- every common video geometry): a thread converts 16 consecutive
// pixels: three 16-byte loads (48 contiguous bytes), per pixel one PRMT (gather B, G, R into a word) + two DP2A
// (3735 B + 19235 G + 16384, then + 9798 R: the 15-bit fixed point of cv::cvtColor, weights fit 16 bits) + shift + I2F +
// FMUL, and two 32-byte stores (st.global.v8.f32: one full sector each).  Anything else takes the 4-pixel generic path.
constexpr int kIngestThreads = 128;
constexpr int kIngestRows = 4;       // rows per CTA of k_ingest
// the gray level behind a toGrayF32 value of a u8-sourced frame: f = fl32(g * fl32(1/255)) -> rint(f * 255) == g for g = 0..255
__device__ __forceinline__ unsigned int gray8_of(float f) { return (unsigned int)__float2int_rn(f * 255.0f) & 255u; }
__device__ __forceinline__ void store_gray8(const Ctx& c, int stream, int x, int y, const float4& o)
{
    *reinterpret_cast<unsigned int*>(c.gray8 + (size_t)stream * c.plane8 + (size_t)y * c.pitch8 + x) =
        gray8_of(o.x) | (gray8_of(o.y) << 8) | (gray8_of(o.z) << 16) | (gray8_of(o.w) << 24);
}
__device__ __forceinline__ float gray_px(unsigned int w /* bytes: B G R x */)
{
    unsigned int acc = __dp2a_lo((19235u << 16) | 3735u, w, 16384u);   // 3735 * B + 19235 * G + 16384
    acc = __dp2a_hi(9798u, w, acc);                                     // + 9798 * R + 0 * x
    return gray_to_f32(acc >> 15);
}
__device__ __forceinline__ void st_v8(float* p, const float (&v)[8])
{
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__global__ void __launch_bounds__(kIngestThreads) k_ingest(Ctx c)
{
    pdl_trigger();
    const unsigned long long step = *c.step;
    const int stream = blockIdx.z, y0 = blockIdx.y * kIngestRows;
    if (c.global_pass && !c.stream_need[stream]) return;   // whole-frame pass: only streams with a lost track
    const FrameDesc d = c.table[table_row(c, step) + stream];
    if (!d.valid) return;
    trace_begin(c, step, TR_INGEST);
    const bool fast = d.format == PVT_FMT_BGR8 && (c.W & 15) == 0 && ((((size_t)d.data) | d.step) & 15) == 0;
    if (fast) {
        // a CTA converts kIngestRows consecutive rows: the preamble above (three dependent loads) is paid once per 4 x 2048 pixels,
        // and a thread has its 12 loads (192 bytes) in flight before the first conversion
        const int x = (blockIdx.x * kIngestThreads + threadIdx.x) * 16;
        if (x < c.W) {
            uint4 q[kIngestRows][3];
#pragma unroll
            for (int r = 0; r < kIngestRows; ++r) {
                const int y = min(y0 + r, c.H - 1);              // (clamped duplicates are not stored)
                const uint4* p = reinterpret_cast<const uint4*>((const unsigned char*)d.data + (size_t)y * d.step + 3 * x);
                q[r][0] = __ldg(p); q[r][1] = __ldg(p + 1); q[r][2] = __ldg(p + 2);
            }
#pragma unroll
            for (int r = 0; r < kIngestRows; ++r) {
                const int y = y0 + r;
                if (y < c.H) {
                    float* orow = c.gray + (size_t)stream * c.plane + (size_t)y * c.pitch;
                    const unsigned int w[12] = {q[r][0].x, q[r][0].y, q[r][0].z, q[r][0].w, q[r][1].x, q[r][1].y, q[r][1].z, q[r][1].w,
                                                q[r][2].x, q[r][2].y, q[r][2].z, q[r][2].w};
                    float o[16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {   // 4 pixels = 3 words: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
                        const unsigned int a = w[3 * g], b = w[3 * g + 1], e = w[3 * g + 2];
                        o[4 * g] = gray_px(a);                                   // bytes 0 1 2 of a
                        o[4 * g + 1] = gray_px(__byte_perm(a, b, 0x0543));       // a3 b0 b1
                        o[4 * g + 2] = gray_px(__byte_perm(b, e, 0x0432));       // b2 b3 e0
                        o[4 * g + 3] = gray_px(e >> 8);                          // e1 e2 e3
                    }
                    if (c.gray8) {   // tensor-core search: keep the gray levels themselves (exactly rint(f * 255))
                        unsigned int b[4];
#pragma unroll
                        for (int g = 0; g < 4; ++g)
                            b[g] = gray8_of(o[4 * g]) | (gray8_of(o[4 * g + 1]) << 8) | (gray8_of(o[4 * g + 2]) << 16) | (gray8_of(o[4 * g + 3]) << 24);
                        *reinterpret_cast<uint4*>(c.gray8 + (size_t)stream * c.plane8 + (size_t)y * c.pitch8 + x) = make_uint4(b[0], b[1], b[2], b[3]);
                    }
                    if ((c.pitch & 7) == 0) {
                        st_v8(orow + x, *reinterpret_cast<const float(*)[8]>(o));
                        st_v8(orow + x + 8, *reinterpret_cast<const float(*)[8]>(o + 8));
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; ++g) *reinterpret_cast<float4*>(orow + x + 4 * g) = make_float4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
                    }
                }
            }
        }
    } else {
        // generic: 4-pixel groups, 4 per thread (strided by the block so that a warp's accesses stay contiguous)
        const int gpr = (c.W + 3) >> 2;
        for (int y = y0; y < min(y0 + kIngestRows, c.H); ++y) {
            float* orow = c.gray + (size_t)stream * c.plane + (size_t)y * c.pitch;
            const unsigned char* irow = (const unsigned char*)d.data + (size_t)y * d.step;
#pragma unroll 1
            for (int k = 0; k < 4; ++k) {
                const int g = (blockIdx.x * 4 + k) * kIngestThreads + threadIdx.x;
                if (g >= gpr) break;
                const int x = g << 2;
                const float4 o4 = ingest_group(d, irow, x, min(4, c.W - x));
                *reinterpret_cast<float4*>(orow + x) = o4;  // pitch % 4 == 0
                if (c.gray8) store_gray8(c, stream, x, y, o4);
            }
        }
    }
    trace_end(c, step, TR_INGEST);
}

// ROI ingest: one track's search tile only (window + template extent, origin from device state).  The source may be a
// device buffer or PINNED HOST memory read directly over PCIe (zero-copy): ~150 KB per track and step for a 1080p /
// 64x64 / R80 track instead of a 6.2 MB frame.  Pixels outside the tiles keep older (finite) values; nothing on the
// path reads them: every consumer (k_colprefix, the TMA tile's useful part, the EMA patch) stays inside the tile.
__global__ void __launch_bounds__(256) k_ingest_roi(Ctx c)
{
    pdl_trigger();
    const int track = blockIdx.y;
    const TrackState& t = c.tracks[track];
    unsigned long long step;
    FrameDesc d;
    const bool stepped = track_stepped_ld(c, t, step, &d);        // all of the preamble's loads go out before the first branch
    const DevParams P = *c.params;
    const int bx_ = t.x, by_ = t.y, btw = t.w, bth = t.h, bstream = t.stream;
    if (!stepped) return;
    trace_begin(c, step, TR_INGEST);
    int win[4];
    search_window(bx_, by_, btw, bth, c.W - btw + 1, c.H - bth + 1, P.rx, P.ry, win);
    const int x0 = win[0] & ~3, x1 = min(c.W, win[0] + win[2] + btw - 1), rows = win[3] + bth - 1;
    const int gpr = (x1 - x0 + 3) >> 2;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (c.stage && gid < 4) c.stage_hdr[track].win[gid] = win[gid];   // for k_prefetch_roi, which runs behind this kernel
    if (c.stage && gid == 0) c.stage_hdr[track].cur_step = step;
    if (gid >= gpr * rows) return;
    const int r = gid / gpr, x = x0 + ((gid - r * gpr) << 2), y = win[1] + r;
    float* out = c.gray + (size_t)bstream * c.plane + (size_t)y * c.pitch + x;
    // Did k_prefetch_roi stage this tile during the previous step (pinned host rings)?  Then it is a device-to-device copy
    // of already converted pixels; the zero-copy read over PCIe happened off the critical path.
    // (Only inside a sequence over a pinned ring, whose frames the caller keeps unchanged: SeqDesc.prefetch.)
    if (c.stage && c.seq->prefetch) {
        const StageHdr h = c.stage_hdr[track];
        if (h.step == step && h.data == d.data && x0 >= h.x0 && x1 <= h.x1 && win[1] >= h.y0 && win[1] + rows <= h.y1) {
            const float* sp = c.stage + (size_t)track * c.stage_w * c.stage_h + (size_t)(y - h.y0) * c.stage_w + (x - h.x0);
            const float4 o4 = *reinterpret_cast<const float4*>(sp);
            *reinterpret_cast<float4*>(out) = o4;
            if (c.gray8) store_gray8(c, bstream, x, y, o4);
            trace_end(c, step, TR_INGEST);
            return;
        }
    }
    const unsigned char* row = (const unsigned char*)d.data + (size_t)y * d.step;
    const float4 o4 = ingest_group(d, row, x, min(4, c.W - x));
    *reinterpret_cast<float4*>(out) = o4;
    if (c.gray8) store_gray8(c, bstream, x, y, o4);
    trace_end(c, step, TR_INGEST);
}

// Pinned host rings (SeqDesc.prefetch): while time step k computes, convert the pixels step k+1 will most likely need into
// a per-track staging buffer: the current search tile grown by HALF the search radius on every side (the next tile lies
// inside it whenever the box moves by at most R/2 in this step; otherwise the next ingest simply reads zero-copy as
// before -- the staging is an accelerator, never a correctness dependency).  ~1.9x the tile's bytes cross PCIe, but
// beside the step instead of in front of its search.  Runs on a parallel graph branch that joins at the end of the step;
// it works from the window k_ingest_roi recorded, not from the box, which the step's update moves.
__global__ void __launch_bounds__(256) k_prefetch_roi(Ctx c, int debug_delay_ns)
{
    // k_ncc_local shape: launched behind k_winstats with a programmatic dependency and NO griddepcontrol.wait: it only has to start
    // after the search CTAs have their SMs (k_winstats' CTAs start with them), not after the statistics (behind their end it
    // finished after the update and held up the next step's ingest: 21.7 instead of 19.9 us per step on pinned host rings).
    // Starting before the ingest has stored this step's header is harmless: the tile is then staged under the previous step's
    // tag, which no ingest accepts (the staging is an accelerator, never a correctness dependency).
    const SeqDesc q = *c.seq;
    if (!q.prefetch) return;
    if (debug_delay_ns > 0) {   // test hook (PVT_DEBUG_PREFETCH_DELAY_US): start late, as if the SMs had been busy
        const unsigned long long t0 = gtime();
        while (gtime() - t0 < (unsigned long long)debug_delay_ns) { }
    }
    const int track = blockIdx.y;
    const TrackState& t = c.tracks[track];
    if (!t.active) return;
    StageHdr* hdr = c.stage_hdr + track;
    // The step this branch belongs to comes from the header k_ingest_roi filled, NOT from *c.step: the branch only joins
    // at the end of the step, after the update has advanced the counter, and a CTA that starts late would otherwise stage
    // frame step+2 under the tag of step+1.
    const unsigned long long step = hdr->cur_step;
    if (step == ~0ull) return;                           // no ingest has run for this track in this sequence yet
    if (c.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) c.trace[((step % kRing) * 8 + TR_ROWSUM) * 2] = gtime();   // (spare slot)
    const size_t nrow = (size_t)(q.row0 + (int)((step + 1ull - q.step0 + (unsigned long long)q.phase) % (unsigned long long)q.ring_len)) * c.max_streams;
    const FrameDesc d = c.table[nrow + t.stream];
    if (!d.valid) {
        if (blockIdx.x == 0 && threadIdx.x == 0) hdr->step = ~0ull;
        return;
    }
    const DevParams P = *c.params;
    const int win[4] = {hdr->win[0], hdr->win[1], hdr->win[2], hdr->win[3]};
    if (win[2] <= 0 || win[3] <= 0) return;
    const int gx = P.rx / 2, gy = P.ry / 2;
    const int x0 = max(0, win[0] - gx) & ~3, x1 = min(c.W, win[0] + win[2] + t.w - 1 + gx);
    const int y0 = max(0, win[1] - gy), y1 = min(c.H, win[1] + win[3] + t.h - 1 + gy);
    const int gpr = (x1 - x0 + 3) >> 2, rows = y1 - y0;
    if (gpr * 4 > c.stage_w || rows > c.stage_h) {      // cannot happen with the sizes pvt_create derives; never overrun
        if (blockIdx.x == 0 && threadIdx.x == 0) hdr->step = ~0ull;
        return;
    }
    const int gid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    if (gid == 0) { hdr->x0 = x0; hdr->y0 = y0; hdr->x1 = x1; hdr->y1 = y1; hdr->step = step + 1ull; hdr->data = d.data; }
    // A SMALL grid (it shares the GPU with the step's own kernels), every thread converts several 4-pixel groups and has all
    // their loads in flight before the first conversion: the reads are PCIe round trips
    float* stage = c.stage + (size_t)track * c.stage_w * c.stage_h;
    const int total = gpr * rows;
    const bool fast = d.format == PVT_FMT_BGR8 && ((((size_t)d.data) | d.step) & 3) == 0 && (c.W & 3) == 0;
    for (int g0 = gid; g0 < total; g0 += 4 * nthr) {
        if (fast) {
            unsigned int a[4], b[4], e[4];
            int off[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int g = g0 + k * nthr;
                off[k] = -1;
                if (g < total) {
                    const int r = g / gpr, x = x0 + ((g - r * gpr) << 2);
                    const unsigned int* p32 = (const unsigned int*)((const unsigned char*)d.data + (size_t)(y0 + r) * d.step + 3 * x);
                    a[k] = __ldg(p32); b[k] = __ldg(p32 + 1); e[k] = __ldg(p32 + 2);
                    off[k] = r * c.stage_w + (x - x0);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (off[k] >= 0) {
                    float4 o;  // bytes: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
                    o.x = gray_to_f32(bgr_to_gray(a[k] & 255u, (a[k] >> 8) & 255u, (a[k] >> 16) & 255u));
                    o.y = gray_to_f32(bgr_to_gray(a[k] >> 24, b[k] & 255u, (b[k] >> 8) & 255u));
                    o.z = gray_to_f32(bgr_to_gray((b[k] >> 16) & 255u, b[k] >> 24, e[k] & 255u));
                    o.w = gray_to_f32(bgr_to_gray((e[k] >> 8) & 255u, (e[k] >> 16) & 255u, e[k] >> 24));
                    *reinterpret_cast<float4*>(stage + off[k]) = o;
                }
            }
        } else {
            for (int k = 0; k < 4; ++k) {
                const int g = g0 + k * nthr;
                if (g >= total) break;
                const int r = g / gpr, x = x0 + ((g - r * gpr) << 2);
                const unsigned char* row = (const unsigned char*)d.data + (size_t)(y0 + r) * d.step;
                *reinterpret_cast<float4*>(stage + (size_t)r * c.stage_w + (x - x0)) = ingest_group(d, row, x, min(4, c.W - x));
            }
        }
    }
    if (c.trace && threadIdx.x == 0) atomicMax(&c.trace[((step % kRing) * 8 + TR_ROWSUM) * 2 + 1], gtime());
}

// =============================================================================================
// (2) window statistics in FP64 -- what cv::matchTemplate gets from cv::integral(..., CV_64F):
//     wsum = sum of f, wsq = sum of f^2 over every candidate's tw x th window, then OpenCV's
//     normaliser (common_matchTemplate, TM_CCOEFF_NORMED):
//       diff2 = max(wsq - wsum^2 * invArea, 0)
//       t     = diff2 <= min(0.5, 10*FLT_EPSILON*wsq) ? 0 : sqrt(diff2) * sigma_t / sqrt(invArea)
//     Two scan kernels with short dependency chains and many independent loads (the tile is L2-resident; what
//     costs is serialised L2 latency, not bandwidth):
//       k_colprefix: C[r][X] = sum_{r' < r} f[r'][X] over the search tile, per 32-column strip; each thread owns a
//                    chunk of rows (independent loads), chunk offsets through shared memory.
//       k_rowsum:    one warp per candidate row y: D[X] = C[y+th][X] - C[y][X] (vertical box sums), inclusive prefix
//                    of D along x (8 values per lane + warp scan), box = P[x+tw] - P[x], normaliser.
//     For u8-sourced frames every partial sum of f is exactly representable in double, so wsum is
//     bit-identical to OpenCV's whole-frame integral; wsq agrees to ~1e-15 relative.
// =============================================================================================
__global__ void __launch_bounds__(1024) k_colprefix(Ctx c)
{
    __shared__ double tot[2][32][33];
    pdl_trigger();
    const int track = blockIdx.y;
    TrackState& t = c.tracks[track];
    const unsigned long long step = *c.step;
    if (!track_stepped(c, t, step)) return;
    trace_begin(c, step, TR_COLPREFIX);
    const DevParams P = *c.params;
    int win[4];
    search_window(t.x, t.y, t.w, t.h, c.W - t.w + 1, c.H - t.h + 1, P.rx, P.ry, win);
    if (blockIdx.x == 0 && threadIdx.y == 0 && threadIdx.x < 4) t.win[threadIdx.x] = win[threadIdx.x];
    const int tileW = win[2] + t.w - 1, tileH = win[3] + t.h - 1;
    const int lx = threadIdx.x, ch = threadIdx.y, nch = blockDim.y;
    const int X = blockIdx.x * 32 + lx;
    if (blockIdx.x * 32 >= tileW) return;
    const int rc = (tileH + nch - 1) / nch;                    // rows per chunk
    const int r0 = min(ch * rc, tileH), r1 = min(r0 + rc, tileH);
    const bool okx = X < tileW;
    const float* col = c.gray + (size_t)t.stream * c.plane + (size_t)win[1] * c.pitch + win[0] + X;
    pdl_wait();   // K-split shape: launched behind the ingest with a programmatic dependency; the gray plane is complete now
    double s = 0.0, q = 0.0;
    if (okx) {
        // loads are issued in batches of 16 before any is consumed: one L2 round trip per batch instead of one per row
        for (int r = r0; r < r1; r += 16) {
            float v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = (r + k < r1) ? __ldg(col + (size_t)(r + k) * c.pitch) : 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const double d = (double)v[k];
                s += d;
                q += d * d;
            }
        }
    }
    tot[0][ch][lx] = s;
    tot[1][ch][lx] = q;
    __syncthreads();
    if (!okx) return;
    s = 0.0; q = 0.0;
    for (int k = 0; k < ch; ++k) { s += tot[0][k][lx]; q += tot[1][k][lx]; }
    const int rows = c.Hmax + c.mth;                             // C has tileH + 1 <= Hmax + mth rows
    double* cs = c.vsum + ((size_t)track * rows) * c.VW + X;
    double* cq = c.vsq + ((size_t)track * rows) * c.VW + X;
    if (ch == 0) { cs[0] = 0.0; cq[0] = 0.0; }
    for (int r = r0; r < r1; r += 16) {                         // second pass over the chunk: L1 hits
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = (r + k < r1) ? __ldg(col + (size_t)(r + k) * c.pitch) : 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (r + k < r1) {
                const double d = (double)v[k];
                s += d;
                q += d * d;
                cs[(size_t)(r + k + 1) * c.VW] = s;
                cq[(size_t)(r + k + 1) * c.VW] = q;
            }
        }
    }
    if (lx == 0) { if (c.trace) atomicMax(&c.trace[((step % kRing) * 8 + TR_COLPREFIX) * 2 + 1], gtime()); }
}

// OpenCV's normaliser of one candidate from its window sums (common_matchTemplate, TM_CCOEFF_NORMED), exactly in its operation
// order, no contraction: wndMean2 = t*t; wndMean2 *= invArea; ...   (formula != 0: the reference CUDA kernels' eps variant)
__device__ __forceinline__ double window_denom(double wsum, double wsq, double invArea, double tn, int formula)
{
    const double wm2 = __dmul_rn(__dmul_rn(wsum, wsum), invArea);
    double diff2 = __dsub_rn(wsq, wm2);
    if (diff2 < 0.0) diff2 = 0.0;
    double lim = __dmul_rn(10.0 * (double)FLT_EPSILON, wsq);
    if (lim > 0.5) lim = 0.5;
    double dnv = (diff2 <= lim) ? 0.0 : __dmul_rn(sqrt(diff2), tn);
    if (formula) {
        // baseline_kernel.cu:44-48: std = sqrtf(fmaxf(var, 1e-6f)); the score divides by (std + 1e-6f) * (templStd + 1e-6f) * N
        const float sd = sqrtf(fmaxf((float)(diff2 * invArea), 1e-6f));
        dnv = (double)(sd + 1e-6f) * tn;
    }
    return dnv;
}

__global__ void __launch_bounds__(256) k_rowsum(Ctx c, int pw /* doubles per prefix row in smem */)
{
    extern __shared__ double sm_d[];
    const int track = blockIdx.y;
    const TrackState& t = c.tracks[track];
    const unsigned long long step = *c.step;
    pdl_trigger();
    if (!track_stepped(c, t, step)) return;
    pdl_wait();   // k_colprefix (the window it stored in t.win, the column prefix sums) has completed
    trace_begin(c, step, TR_ROWSUM);
    const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int y = blockIdx.x * warps + w;
    const int ww = t.win[2], wh = t.win[3], tw = t.w, th = t.h;
    if (y >= wh) return;  // warp-uniform
    const int tileW = ww + tw - 1;
    double* Ps = sm_d + (size_t)w * 2 * pw;
    double* Pq = Ps + pw;
    const int rows = c.Hmax + c.mth;
    const double* cs0 = c.vsum + ((size_t)track * rows + y) * c.VW;
    const double* cq0 = c.vsq + ((size_t)track * rows + y) * c.VW;
    const double* cs1 = cs0 + (size_t)th * c.VW;
    const double* cq1 = cq0 + (size_t)th * c.VW;
    // Prefix row P[0 .. tileW] of this candidate row in shared memory, element i at PH(i) = i + (i >> 3): one double of
    // padding per 8, so that lanes owning 8 consecutive elements (stride 9 doubles) and lanes reading consecutive
    // elements are both bank-conflict-free.
#define PH(i) ((i) + ((i) >> 3))
    double carry_s = 0.0, carry_q = 0.0;
    for (int base = 0; base < tileW; base += 256) {
        // (1) COALESCED global loads: lane reads elements base + lane + 32 m of the four prefix rows (a warp instruction
        // touches 256 contiguous bytes; the lane-owns-8-consecutive pattern cost 8x the L1 tag lookups and made the kernel
        // L1TEX-bound at 80 %), vertical box sums D = C[y+th] - C[y] staged through shared memory
        {
            double a0[8], a1[8], b0[8], b1[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) {                       // 32 independent loads in flight per lane
                const int i = base + lane + 32 * m;
                const bool ok = i < tileW;
                a0[m] = ok ? __ldg(cs0 + i) : 0.0;
                a1[m] = ok ? __ldg(cs1 + i) : 0.0;
                b0[m] = ok ? __ldg(cq0 + i) : 0.0;
                b1[m] = ok ? __ldg(cq1 + i) : 0.0;
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int i = base + lane + 32 * m;
                if (i < tileW) {
                    Ps[PH(i + 1)] = a1[m] - a0[m];              // vertical box sum of column i
                    Pq[PH(i + 1)] = b1[m] - b0[m];
                }
            }
        }
        __syncwarp();
        // (2) inclusive prefix along x: every lane scans its 8 consecutive elements, then a warp scan of the lane totals
        const int i0 = base + lane * 8;
        double ls[8], lq[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool ok = i0 + k < tileW;
            ls[k] = ok ? Ps[PH(i0 + k + 1)] : 0.0;
            lq[k] = ok ? Pq[PH(i0 + k + 1)] : 0.0;
        }
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            ls[k] += ls[k - 1];
            lq[k] += lq[k - 1];
        }
        double is = ls[7], iq = lq[7];  // inclusive warp scan of the lane totals
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            double ns = shfl_up_f64(is, d), nq = shfl_up_f64(iq, d);
            if (lane >= d) { is += ns; iq += nq; }
        }
        const double off_s = carry_s + (is - ls[7]), off_q = carry_q + (iq - lq[7]);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (i0 + k < tileW) {
                Ps[PH(i0 + k + 1)] = off_s + ls[k];
                Pq[PH(i0 + k + 1)] = off_q + lq[k];
            }
        carry_s += shfl_f64(is, 31);
        carry_q += shfl_f64(iq, 31);
    }
    if (lane == 0) { Ps[0] = 0.0; Pq[0] = 0.0; }
    __syncwarp();
    const double invArea = 1.0 / ((double)th * (double)tw);
    const double tn = t.templ_norm;
    double* dn = c.denom + (size_t)track * c.Hmax * c.Wmax + (size_t)y * ww;
    for (int x = lane; x < ww; x += 32) {
        const double wsum = Ps[PH(x + tw)] - Ps[PH(x)];
        const double wsq = Pq[PH(x + tw)] - Pq[PH(x)];
        dn[x] = window_denom(wsum, wsq, invArea, tn, c.formula);
        if (c.wsum) c.wsum[(size_t)track * c.Hmax * c.Wmax + (size_t)y * ww + x] = wsum;   // the tensor-core search removes its DC error with it
    }
    if (lane == 0 && c.trace) atomicMax(&c.trace[((step % kRing) * 8 + TR_ROWSUM) * 2 + 1], gtime());
}
#undef PH

// k_winstats: the same statistics in ONE kernel and without the FP64 prefix arrays (k_colprefix + k_rowsum write and re-read
// 2 x 8 B per tile pixel through L2/HBM, and are two dependent launches on the single-stream step).
// A CTA (8 warps) owns NX x NY candidates of one track; its tile has NX + tw - 1 <= 256 columns.
//   (A) thread = tile column: vertical box sums of the band's first candidate row, S = sum f, Q = sum f^2 over th rows (FP64,
//       32 loads in flight per thread);
//   (V) per group of 8 candidate rows: the column slides down, S += f[y + th] - f[y] (both rows fetched one group ahead), and
//       leaves its 8 x (S, Q) in shared memory, one padded row per candidate row;
//   (H) warp = candidate row (k_rowsum's scheme): a lane scans its 8 consecutive columns, one warp scan of the lane totals,
//       the prefix row goes back to shared memory, box = P[x + tw] - P[x], OpenCV's normaliser (window_denom) -> denom.
// Numerics: every partial sum of f (u8-sourced: multiples of 2^-31 below 2^21) is exact in double, so wsum is the same number
// whatever the order -- bit-identical to OpenCV's whole-frame integral and to k_rowsum; wsq agrees to ~1e-15 relative.
struct StatCfg {
    int NX, NY;          // candidates per CTA along x (256 - (mtw - 1)) and along y
    int xtiles, ybands;  // CTAs per track = xtiles * ybands
    int signal;          // != 0: every CTA that stored normalisers counts itself in TrackState.stats_done (k_step_fused waits for them)
};
constexpr int kStatThreads = 256;
constexpr int kStatRowD = kStatThreads + kStatThreads / 8 + 2;   // doubles per padded row: element i at i + (i >> 3), i = 0 .. 256
__global__ void __launch_bounds__(kStatThreads, 3) k_winstats(Ctx c, StatCfg sc)
{
    __shared__ double sD[8][2][kStatRowD];          // [candidate row of the group][sum | sum of squares][padded prefix row]
    pdl_trigger();
    const int track = blockIdx.y;
    TrackState& t = c.tracks[track];
    unsigned long long step;
    const bool stepped = track_stepped_ld(c, t, step);            // all of the preamble's loads go out before the first branch
    const DevParams P = *c.params;
    const int bx_ = t.x, by_ = t.y, tw = t.w, th = t.h, tstream = t.stream;
    const double tn = t.templ_norm;
    if (!stepped) return;
    trace_begin(c, step, TR_COLPREFIX);
    int win[4];
    search_window(bx_, by_, tw, th, c.W - tw + 1, c.H - th + 1, P.rx, P.ry, win);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (blockIdx.x == 0 && tid < 4) t.win[tid] = win[tid];
    const int ww = win[2], wh = win[3];
    const int yb = blockIdx.x / sc.xtiles, xt = blockIdx.x - yb * sc.xtiles;
    const int x0 = xt * sc.NX, y0 = yb * sc.NY;
    if (x0 >= ww || y0 >= wh) return;
    const int nx = min(sc.NX, ww - x0), ny = min(sc.NY, wh - y0);
    const bool colok = tid < nx + tw - 1;
    const size_t pitch = (size_t)c.pitch;
    const float* col = c.gray + (size_t)tstream * c.plane + (size_t)(win[1] + y0) * pitch + win[0] + x0 + (colok ? tid : 0);
    const double invArea = 1.0 / ((double)th * (double)tw);
    const int formula = c.formula;
#define PH(i) ((i) + ((i) >> 3))
    pdl_wait();   // latency shape: launched behind the ingest with a programmatic dependency; the gray plane is complete now
    // the slide rows of the first group are requested together with the first column sums: one L2 round trip less
    float vin[8], vout[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const bool need = k + 1 < ny;                           // the slide behind candidate row k
        vin[k] = need ? __ldg(col + (size_t)(k + th) * pitch) : 0.f;
        vout[k] = need ? __ldg(col + (size_t)k * pitch) : 0.f;
    }
    double S = 0.0, Q = 0.0;
    {
        const float* p = col;
        for (int r = 0; r < th; r += 32, p += 32 * pitch) {      // 32 independent loads in flight per thread
            float v[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = (r + k < th) ? __ldg(p + (size_t)k * pitch) : 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const double d = (double)v[k];
                S += d;
                Q += d * d;
            }
        }
    }
    if (!colok) { S = 0.0; Q = 0.0; }
    const size_t woff = (size_t)track * c.Hmax * c.Wmax;
    for (int yy0 = 0; yy0 < ny; yy0 += 8) {
        // (V) this column's vertical box sums of candidate rows yy0 .. yy0 + 7
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            sD[k][0][PH(tid + 1)] = S;
            sD[k][1][PH(tid + 1)] = Q;
            const double di = (double)vin[k], dq = (double)vout[k];
            S += di - dq;                                       // exact (see above); zero where no slide is needed
            Q += di * di - dq * dq;
        }
        if (!colok) { S = 0.0; Q = 0.0; }
        // request the next group's slide rows now: they arrive while the warps scan this group
        if (yy0 + 8 < ny) {
            const float* pin = col + (size_t)(yy0 + 8 + th) * pitch;
            const float* pout = col + (size_t)(yy0 + 8) * pitch;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const bool need = yy0 + 8 + k + 1 < ny;
                vin[k] = need ? __ldg(pin + (size_t)k * pitch) : 0.f;
                vout[k] = need ? __ldg(pout + (size_t)k * pitch) : 0.f;
            }
        }
        __syncthreads();
        // (H) warp wid: candidate row yy0 + wid
        const int yy = yy0 + wid;
        if (yy < ny) {                                          // warp-uniform
            double* Ps = sD[wid][0];
            double* Pq = sD[wid][1];
            const int i0 = lane * 8;
            double ls[8], lq[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { ls[k] = Ps[PH(i0 + k + 1)]; lq[k] = Pq[PH(i0 + k + 1)]; }
#pragma unroll
            for (int k = 1; k < 8; ++k) { ls[k] += ls[k - 1]; lq[k] += lq[k - 1]; }
            double is = ls[7], iq = lq[7];                      // inclusive warp scan of the lane totals
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double ns = shfl_up_f64(is, d), nq = shfl_up_f64(iq, d);
                if (lane >= d) { is += ns; iq += nq; }
            }
            const double off_s = is - ls[7], off_q = iq - lq[7];
#pragma unroll
            for (int k = 0; k < 8; ++k) { Ps[PH(i0 + k + 1)] = off_s + ls[k]; Pq[PH(i0 + k + 1)] = off_q + lq[k]; }
            if (lane == 0) { Ps[0] = 0.0; Pq[0] = 0.0; }
            __syncwarp();
            double* dn = c.denom + woff + (size_t)(y0 + yy) * ww + x0;
            double* wsp = c.wsum ? c.wsum + woff + (size_t)(y0 + yy) * ww + x0 : nullptr;
            for (int x = lane; x < nx; x += 32) {
                const double wsum = Ps[PH(x + tw)] - Ps[PH(x)];
                const double wsq = Pq[PH(x + tw)] - Pq[PH(x)];
                dn[x] = window_denom(wsum, wsq, invArea, tn, formula);
                if (wsp) wsp[x] = wsum;                         // the tensor-core search removes its DC error with it
            }
        }
        __syncthreads();                                        // the next group overwrites the rows
    }
#undef PH
    if (sc.signal) {
        __threadfence();                                        // this thread's normalisers are visible device-wide ...
        __syncthreads();
        if (tid == 0) atomicAdd(&t.stats_done, 1u);             // ... before the CTA counts as done
    }
    if (tid == 0 && c.trace) atomicMax(&c.trace[((step % kRing) * 8 + TR_COLPREFIX) * 2 + 1], gtime());   // one kernel: one slot
}

// OpenCV's final rule for TM_CCOEFF_NORMED (common_matchTemplate): never NaN, always in [-1, 1]
__device__ __forceinline__ float ncc_finalize(float num_f32, double t, int flat_templ)
{
    if (flat_templ) return flat_templ == 1 ? 1.0f : (float)((double)num_f32 / t);   // 2: PVT_FORMULA_EPS, t > 0 always
    double num = (double)num_f32;
    double r;
    if (fabs(num) < t) r = num / t;
    else if (fabs(num) < t * 1.125) r = num > 0 ? 1.0 : -1.0;
    else r = 0.0;
    return (float)r;
}

// =============================================================================================
// Cross-term accumulation order (shared by both NCC kernels, so their results are bit-identical):
//   for every 8-column chunk j of the (zero-padded) centred template:
//       r = 0;  for dy = 0..th-1: for k = 0..7:  r = fma(f[y+dy][x+j+k], tc[dy][j+k], r)      (512 products at 64 rows)
//       acc += r
//   i.e. one FP32 partial per template chunk, chunks then added in FP32 (blocked summation in the sense
//   of SURVEY.md §7 scheme (B); products use the centred template fl32(t - mean_t)).
//   The centred template is stored CHUNK-MAJOR: templc[chunk][dy][8], so a chunk is one contiguous slice.
// =============================================================================================

// (3a) k_ncc_direct: verification twin of the reference's naive kernel (baseline_kernel.cu:21-64):
//      one thread per candidate, operands from global/L2.  PVT_KERNEL_DIRECT.
__global__ void __launch_bounds__(256) k_ncc_direct(Ctx c)
{
    const int track = blockIdx.y;
    TrackState& t = c.tracks[track];
    const unsigned long long step = *c.step;
    if (!track_stepped(c, t, step)) return;
    const int ww = t.win[2], wh = t.win[3];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long key = 0ull;
    if (idx < ww * wh) {
        const int y = idx / ww, x = idx - y * ww;
        const float* f = c.gray + (size_t)t.stream * c.plane + (size_t)(t.win[1] + y) * c.pitch + t.win[0] + x;
        const float* tc = c.templc + (size_t)track * c.mth * c.mtp;
        const int th = t.h, tw = t.w;
        float acc = 0.f;
        for (int j = 0; j < tw; j += 8) {
            const float* tj = tc + (size_t)(j >> 3) * th * 8;
            const int kn = min(8, tw - j);
            float r = 0.f;
            for (int dy = 0; dy < th; ++dy)
                for (int k = 0; k < kn; ++k) r = fmaf(f[(size_t)dy * c.pitch + j + k], tj[dy * 8 + k], r);
            acc += r;
        }
        const double dn = c.denom[(size_t)track * c.Hmax * c.Wmax + idx];
        const float v = ncc_finalize(acc, dn, t.flat);
        if (c.params->keep_maps) c.maps[(size_t)track * c.Hmax * c.Wmax + idx] = v;
        key = peak_key(v, (unsigned int)idx);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        unsigned long long o = shfl_xor_u64(key, m);
        key = o > key ? o : key;
    }
    if ((threadIdx.x & 31) == 0 && key) atomicMax(&t.peak, key);
}

// =============================================================================================
// (3b) k_ncc_search: the production cross-term kernel ("y-sliding" register blocking).
//   Thread tile = 8 consecutive candidates in x  x  CY = 5 ADJACENT candidate rows (40 FP32 accumulators).
//   Loop order: template chunk j (8 columns) outermost, template row dy innermost.  At step dy the thread holds the
//   CY frame-row windows (16 floats each) its candidate rows need; going to dy+1 drops the oldest window and loads
//   ONE new one (4 LDS.128) plus the 8 template values of that row (2 LDS.128, warp-uniform broadcast): 6 shared
//   loads feed 64*CY = 320 FFMA (1.9 % of the instruction stream; the x-sliding predecessor needed 3.9 % and ran at
//   75 % of the FP32 peak where this loop sustains 87 % -- tools/microbench2.cu, profiles/).  The window registers
//   rotate through static names by unrolling dy CY times.
//   Thread tiles of a track are numbered column-major, q = col * G + g (G = row groups per column); a CTA takes
//   128 consecutive tiles.  Consecutive lanes are consecutive row groups, i.e. frame rows CY apart, and the tile
//   pitch is == 4 (mod 8) floats, so with CY odd every quarter-warp LDS.128 hits 8 distinct 16-byte bank groups.
//   Staging: ONE cp.async.bulk.tensor.3d (TMA) brings the CTA's sub-tile (<= span columns x all rows) of the
//   stream's gray plane; the centred template streams through a 4-deep ring of 8-column slices (cp.async.bulk with
//   full/empty mbarriers, no CTA-wide sync in the loop), so a CTA needs ~107 KB and two CTAs share an SM.  Window origin comes from device state.
//   K-split (single-stream latency mode): grid.z parts, part p handles template rows [d0, d1) of chunk range
//   [j0, j1); partial sums go to a scratch plane per part and k_ncc_finalize adds them in a fixed order.
//   Epilogue (no K-split): OpenCV normalisation in FP64, (score desc, index asc) key, warp-shuffle max, atomicMax.
// =============================================================================================
constexpr int kTilesPerCta = 128;

// Shift every row of the shared tile [rows][P] left by S (1..3) floats, in place: row[i] = row[i + S].  Lanes take
// consecutive rows; P / 4 is odd (P == 4 mod 8), so the 8 lanes of a quarter warp hit 8 distinct 16-byte bank groups.
// The last float4 of a row keeps stale (finite) data: the box is 4 floats wider than anything the loop reads.
template <int S>
__device__ __forceinline__ void shift_rows_left(float* tile, int P, int rows)
{
    const int nv = P >> 2;
    for (int r = threadIdx.x; r < rows; r += blockDim.x) {
        float4* row = reinterpret_cast<float4*>(tile + (size_t)r * P);
        // eight vectors are read before the seven they produce are written: the shared-memory latency is paid once per
        // batch, not once per vector (the loop is latency-bound otherwise)
        for (int m0 = 0; m0 + 1 < nv; m0 += 7) {
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (m0 + k < nv) ? row[m0 + k] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                if (m0 + k + 1 < nv) {
                    float4 o;
                    if (S == 1) o = make_float4(v[k].y, v[k].z, v[k].w, v[k + 1].x);
                    else if (S == 2) o = make_float4(v[k].z, v[k].w, v[k + 1].x, v[k + 1].y);
                    else o = make_float4(v[k].w, v[k + 1].x, v[k + 1].y, v[k + 1].z);
                    row[m0 + k] = o;
                }
            }
        }
    }
}


struct TileCfg {
    // Thread-tile grid, relative to the WINDOW ORIGIN: column c = candidates x in [8c, 8c+8), row group g = candidate
    // rows [CY*g, CY*g + CY).  G = ceil(Hmax / CY) and C = ceil(Wmax / 8) -- except that a remainder of exactly ONE row /
    // column (2R+1 = 161 = 32*5 + 1 = 20*8 + 1, the reference's default radius) is left out of the grid: a 33rd row
    // group / 21st column would spend full 8 x CY tiles on 1/5 resp. 1/8 useful candidates.  Those "fringe" candidates
    // (x >= 8C or y >= CY*G) are computed by k_ncc_fringe in the same accumulation order, concurrently.
    int G, C;
    int GB, bands;       // row groups per band (a CTA's tiles live in one band), bands = ceil(G / GB)
    int ctas_band;       // CTAs per band (ceil(GB * C / 128))
    int span;            // columns one CTA's 128 tiles can touch
    int boxW, boxH;      // TMA box == shared tile [boxH][boxW]
    int pj, pd;          // K-split: parts along template chunks and along template rows (pj * pd = grid.z)
    // 1-D item grid: item = track * cpt + CTA-in-track.  Tail splitting (only when pj * pd == 1): the items of the last,
    // partial round (item >= n_full) are cut into tail_ps parts along the template chunks, so that round costs
    // 1/tail_ps of a full one; their partial sums are reduced per item by k_ncc_tail_finalize.
    int cpt, n_full, n_tail, tail_ps;
};

// The body of k_ncc_search (one CTA = 128 thread tiles of one track, or one K-split / tail part of them).  FUSED (k_step_fused):
// a `return` only leaves this function -- the calling kernel goes on to the in-kernel second stage and the update.
template <int CY, bool FUSED>
__device__ __forceinline__ void ncc_search_body(const Ctx& c, const TileCfg& g, const CUtensorMap* tmap_p, unsigned char* sm_raw, TrackState& t,
                                                unsigned long long step, int track, int blk, int part, int pj, int pd, int tail_k)
{
    trace_begin(c, step, TR_NCC);
    // the window is derived here (not read from t.win): in K-split mode this kernel runs concurrently with the
    // statistics kernels, which are the ones that store it
    int win[4];
    {
        const DevParams P = *c.params;
        search_window(t.x, t.y, t.w, t.h, c.W - t.w + 1, c.H - t.h + 1, P.rx, P.ry, win);
    }
    const int ww = win[2], wh = win[3];
    const int th = t.h, nchunk = t.tp >> 3, tstream = t.stream;   // read BEFORE griddepcontrol.wait (a memory clobber)
    // A TMA tile must start on a 16-byte boundary in x (an unaligned innermost start coordinate raises "illegal
    // instruction" on sm_100a: tools/tma_align_probe.cu), but the window origin is arbitrary.  The tile is therefore
    // fetched from the origin rounded DOWN to 4 pixels and, when xs = origin & 3 is not 0, every row is shifted left by
    // xs floats in shared memory once (~0.7 us of a ~190 us CTA), so that thread-tile columns start at the window origin.
    const int xs = win[0] & 3;
    const int band = blk / g.ctas_band, q0 = (blk - band * g.ctas_band) * kTilesPerCta;
    const int c_lo = q0 / g.GB;
    const int row0 = band * g.GB * CY;  // first candidate row of this band
    if (c_lo * 8 >= ww || row0 >= wh) return;

    // K-split part -> template chunk range [j0, j1) and row range [d0, d1)
    const int pjx = part % pj, pdx = part / pj;
    const int j0 = (nchunk * pjx) / pj;
    const int d0 = (th * pdx) / pd, d1 = (th * (pdx + 1)) / pd;
    const int nd = d1 - d0;
    const int j1 = nd > 0 ? (nchunk * (pjx + 1)) / pj : j0;
    const bool split = pj * pd > 1;

    constexpr int kRingT = 4;  // template slices in flight (full/empty mbarrier ring; no CTA-wide sync in the loop)
    float* s_tile = reinterpret_cast<float*>(sm_raw);
    const size_t sstride = (size_t)c.mth * 8;                      // kRingT slices of [mth][8]
    float* s_templ = s_tile + (size_t)g.boxW * g.boxH;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_templ + kRingT * sstride);  // [0] tile, [1..4] full, [5..8] empty
    uint64_t* full = bars + 1;
    uint64_t* empty = bars + 1 + kRingT;
    const float* gtempl = c.templc + (size_t)track * c.mth * c.mtp;
    const uint32_t slice_bytes = (uint32_t)nd * 32u;
    const int nj = j1 - j0;

    const int q = q0 + threadIdx.x;
    const int col = q / g.GB, gl = q - col * g.GB;
    const int grp = band * g.GB + gl;
    const bool active = col < g.C && col * 8 < ww && grp < g.G && grp * CY < wh;
    // warps without a single active tile (e.g. half of a track's last CTA) take no part in the template-slice ring and
    // leave right after setup: left in the loop they would spin on the slice barriers for the CTA's whole lifetime and
    // take issue slots from the co-resident CTA.  The `empty` barriers count the participating warps only.
    const bool warp_on = __any_sync(0xffffffffu, active) || threadIdx.x < 32;   // warp 0 hosts the producer thread: always in
    const int n_part = __syncthreads_count(warp_on && (threadIdx.x & 31) == 0);

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
#pragma unroll
        for (int s = 0; s < kRingT; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], n_part);
        }
        fence_mbar_init();
        // K-split shape: launched behind the ingest with a programmatic dependency -> the gray plane must be complete before
        // the tile is fetched.  (Unsplit shape: launched behind k_rowsum, whose output only the epilogue needs: see below.)
        if (split) pdl_wait();
        mbar_arrive_expect_tx(&bars[0], (uint32_t)(g.boxW * g.boxH) * 4u);
        tma_load_3d(s_tile, tmap_p, &bars[0], win[0] - xs + (c_lo + j0) * 8, win[1] + row0 + d0, tstream);
        if (split && c.trace && blockIdx.x == 0 && part == 0) c.trace[((step % kRing) * 8 + TR_FRINGE) * 2] = gtime();      // K-split phase stamps of CTA 0:
        for (int s = 0; s < 2 && s < nj; ++s) {
            mbar_arrive_expect_tx(&full[s], slice_bytes);
            bulk_load(s_templ + (size_t)s * sstride, gtempl + ((size_t)(j0 + s) * th + d0) * 8, slice_bytes, &full[s]);
        }
    }
    __syncthreads();  // barriers initialised before anybody polls them
    const int P = g.boxW;
    if (xs) {   // CTA-uniform.  Every warp (also those about to leave) takes its share of rows.
        mbar_wait(&bars[0], 0);
        if (xs == 1) shift_rows_left<1>(s_tile, P, g.boxH);
        else if (xs == 2) shift_rows_left<2>(s_tile, P, g.boxH);
        else shift_rows_left<3>(s_tile, P, g.boxH);
        __syncthreads();
        if (!warp_on) return;
    } else {
        if (!warp_on) return;
        mbar_wait(&bars[0], 0);
    }

    const float* base = s_tile + (size_t)(gl * CY) * P + (col - c_lo) * 8;
    const bool stamp = split && c.trace && blockIdx.x == 0 && part == 0 && threadIdx.x == 0;   // tile requested | landed, loop done | partial sums stored
    if (stamp) c.trace[((step % kRing) * 8 + TR_FRINGE) * 2 + 1] = gtime();

    float acc[CY][8];
#pragma unroll
    for (int i = 0; i < CY; ++i)
#pragma unroll
        for (int cx = 0; cx < 8; ++cx) acc[i][cx] = 0.f;

    for (int it = 0; it < nj; ++it) {
        const int j = j0 + it, slot = it & (kRingT - 1);
        if (threadIdx.x == 0 && it + 2 < nj) {
            // producer: slice it+2 goes to the slot slice it-2 used; wait until every warp released that slot
            const int c2 = it + 2, s2 = c2 & (kRingT - 1);
            if (c2 >= kRingT) mbar_wait(&empty[s2], (uint32_t)((c2 / kRingT) - 1) & 1u);
            mbar_arrive_expect_tx(&full[s2], slice_bytes);
            bulk_load(s_templ + (size_t)s2 * sstride, gtempl + ((size_t)(j0 + c2) * th + d0) * 8, slice_bytes, &full[s2]);
        }
        mbar_wait(&full[slot], (uint32_t)(it / kRingT) & 1u);
        if (active) {
            const float* st = s_templ + (size_t)slot * sstride;
            const float* fb = base + (j - j0) * 8;
            float racc[CY][8], w[CY][16];
#pragma unroll
            for (int i = 0; i < CY; ++i)
#pragma unroll
                for (int cx = 0; cx < 8; ++cx) racc[i][cx] = 0.f;
#pragma unroll
            for (int r = 0; r < CY - 1; ++r) {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float4 a = *reinterpret_cast<const float4*>(fb + r * P + 4 * v);
                    w[r][4 * v] = a.x; w[r][4 * v + 1] = a.y; w[r][4 * v + 2] = a.z; w[r][4 * v + 3] = a.w;
                }
            }
#pragma unroll 1
            for (int e0 = 0; e0 < nd; e0 += CY) {
#pragma unroll
                for (int u = 0; u < CY; ++u) {
                    const int e = e0 + u;  // template row d0 + e; tile rows are relative to d0
                    if (e < nd) {
                        float(&wn)[16] = w[(u + CY - 1) % CY];
                        const float* pr = fb + (size_t)(e + CY - 1) * P;
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const float4 a = *reinterpret_cast<const float4*>(pr + 4 * v);
                            wn[4 * v] = a.x; wn[4 * v + 1] = a.y; wn[4 * v + 2] = a.z; wn[4 * v + 3] = a.w;
                        }
                        float tt[8];
                        {
                            const float4 a = *reinterpret_cast<const float4*>(st + e * 8);
                            const float4 b = *reinterpret_cast<const float4*>(st + e * 8 + 4);
                            tt[0] = a.x; tt[1] = a.y; tt[2] = a.z; tt[3] = a.w; tt[4] = b.x; tt[5] = b.y; tt[6] = b.z; tt[7] = b.w;
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k)
#pragma unroll
                            for (int i = 0; i < CY; ++i)
#pragma unroll
                                for (int cx = 0; cx < 8; ++cx) racc[i][cx] = fmaf(w[(u + i) % CY][k + cx], tt[k], racc[i][cx]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < CY; ++i)
#pragma unroll
                for (int cx = 0; cx < 8; ++cx) acc[i][cx] += racc[i][cx];
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[slot]);  // this warp is done with the slice
    }

    unsigned long long key = 0ull;
    if (stamp) c.trace[((step % kRing) * 8 + TR_TAIL) * 2] = gtime();
    if (active) {
        const size_t woff = (size_t)track * c.Hmax * c.Wmax;
        if (split) {
            // K-split: store the partial sums TILE-MAJOR (this thread's 8 x CY values contiguous: coalesced float4
            // stores, no write amplification); k_ncc_finalize adds the parts in order and normalises
            const size_t tiles_track = (size_t)g.cpt * kTilesPerCta;
            float4* po = reinterpret_cast<float4*>(
                c.partial + (tail_k >= 0 ? ((size_t)tail_k * kTilesPerCta + threadIdx.x)
                                         : (((size_t)part * c.max_tracks + track) * tiles_track + (size_t)blk * kTilesPerCta + threadIdx.x)) * (8 * CY));
#pragma unroll
            for (int i = 0; i < CY; ++i) {
                po[2 * i] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                po[2 * i + 1] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
            }
        } else {
            pdl_wait();   // unsplit shape: the window statistics (k_rowsum, the programmatic predecessor) are complete
            const double* dn = c.denom + woff;
            float* mp = c.params->keep_maps ? c.maps + woff : nullptr;
            const int flat = t.flat;
#pragma unroll
            for (int i = 0; i < CY; ++i) {
                const int y = grp * CY + i;
                if (y < wh) {
                    // issue the row's 8 normaliser loads back to back (the window registers are dead by now), then finalise
                    double d8[8];
#pragma unroll
                    for (int cx = 0; cx < 8; ++cx) {
                        const int x = col * 8 + cx;
                        d8[cx] = x < ww ? __ldg(dn + y * ww + x) : 0.0;
                    }
#pragma unroll
                    for (int cx = 0; cx < 8; ++cx) {
                        const int x = col * 8 + cx;
                        if (x < ww) {
                            const unsigned int idx = (unsigned int)(y * ww + x);
                            const float v = ncc_finalize(acc[i][cx], d8[cx], flat);
                            if (mp) mp[idx] = v;
                            const unsigned long long k = peak_key(v, idx);
                            key = k > key ? k : key;
                        }
                    }
                }
            }
        }
    }
    if (!split) {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            unsigned long long o = shfl_xor_u64(key, m);
            key = o > key ? o : key;
        }
        if ((threadIdx.x & 31) == 0 && key) atomicMax(&t.peak, key);
    }
    if (stamp) c.trace[((step % kRing) * 8 + TR_TAIL) * 2 + 1] = gtime();
    trace_end(c, step, TR_NCC);
}

template <int CY>
__global__ void __launch_bounds__(kTilesPerCta, 2) k_ncc_search(Ctx c, TileCfg g, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(128) unsigned char sm_raw[];
    // programmatic dependent launch: k_ncc_fringe (launched right behind this kernel in the throughput shape, and
    // independent of its results) may start as soon as every CTA of this grid has been dispatched, i.e. it fills the
    // SM slots that free up while the last round of search CTAs is still running
    pdl_trigger();
    int item = blockIdx.x, part = blockIdx.z, pj = g.pj, pd = g.pd, tail_k = -1;
    if (g.tail_ps > 1 && item >= g.n_full) {           // a part of a tail item
        tail_k = item - g.n_full;
        item = g.n_full + tail_k / g.tail_ps;
        part = tail_k - (tail_k / g.tail_ps) * g.tail_ps;
        pj = g.tail_ps;
        pd = 1;
    }
    const int track = item / g.cpt, blk = item - track * g.cpt;
    TrackState& t = c.tracks[track];
    const unsigned long long step = *c.step;
    if (!track_stepped(c, t, step)) return;
    ncc_search_body<CY, false>(c, g, &tmap, sm_raw, t, step, track, blk, part, pj, pd, tail_k);
}

// (3b') k_step_fused: search + second stage + update of the K-split (single-stream latency) shape in ONE launch.
//   The K-split step used to be k_ncc_search -> k_ncc_finalize (whose last CTA runs the update): a kernel boundary and a second
//   preamble on the critical path of every frame.  Here the search CTAs of a track (all co-resident: the planner only picks
//   this kernel when they fit one wave with room to spare) store their partial cross terms, meet at an arrival counter, wait for
//   the statistics kernel's completion count (k_winstats runs beside this kernel on a parallel graph branch), then EVERY CTA
//   reduces a slice of the track's thread tiles in part order (coalesced float4 reads of the tile-major partial sums),
//   normalises, and feeds the peak; the last CTA through the ticket runs track_update.  Spins are bounded (a device that
//   cannot co-schedule the grid raises Ctx.fault -> PVT_ERR_CUDA at the next host sync instead of hanging).
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool spin_until(const unsigned int* p, unsigned int target)
{
    if (ld_acquire_u32(p) >= target) return true;
    const unsigned long long t0 = gtime();
    while (ld_acquire_u32(p) < target)
        if (gtime() - t0 > 200000000ull) return false;      // 0.2 s: six orders of magnitude above a step
    return true;
}

template <int CY>
__global__ void __launch_bounds__(kTilesPerCta, 2) k_step_fused(Ctx c, TileCfg g, const __grid_constant__ CUtensorMap tmap, StatCfg sc)
{
    extern __shared__ __align__(128) unsigned char sm_raw[];
    __shared__ double red[64];
    __shared__ int s_last;
    pdl_trigger();
    const int parts = g.pj * g.pd, nbx = g.n_full;                  // K-split plans have no tail items
    const int item = blockIdx.x % nbx, part = blockIdx.x / nbx;
    const int track = item / g.cpt, blk = item - track * g.cpt;
    TrackState& t = c.tracks[track];
    const unsigned long long step = *c.step;
    const bool stepped = track_stepped(c, t, step);
    const int n_ctas = g.cpt * parts;                               // CTAs of this track
    const int tid = threadIdx.x;
    if (stepped) {
        ncc_search_body<CY, true>(c, g, &tmap, sm_raw, t, step, track, blk, part, g.pj, g.pd, -1);
        pdl_wait();                                                  // every thread: the ingest's gray plane (the EMA reads it) is visible
        int win[4];
        {
            const DevParams P = *c.params;
            search_window(t.x, t.y, t.w, t.h, c.W - t.w + 1, c.H - t.h + 1, P.rx, P.ry, win);
        }
        const int ww = win[2], wh = win[3];
        __threadfence();                                             // this thread's partial sums are visible device-wide ...
        __syncthreads();
        if (tid == 0) {
            atomicAdd(&t.arrive, 1u);                                // ... before the CTA counts as arrived
            const unsigned int n_stat = (unsigned int)(((ww + sc.NX - 1) / sc.NX) * ((wh + sc.NY - 1) / sc.NY));
            if (!spin_until(&t.arrive, (unsigned int)n_ctas) || !spin_until(&t.stats_done, n_stat)) { *c.fault = 1u; __threadfence_system(); }
        }
        __syncthreads();
        trace_begin(c, step, TR_FINALIZE);
        // second stage: units of one float4 (4 candidates of one row of a thread tile); CTA k takes units [k U, (k + 1) U)
        const int tiles_track = g.cpt * kTilesPerCta, U_total = tiles_track * 2 * CY;
        const int k = part * g.cpt + blk, U = (U_total + n_ctas - 1) / n_ctas;
        const size_t plane4 = (size_t)c.max_tracks * tiles_track * (2 * CY);      // float4 per part
        const float4* pbase = reinterpret_cast<const float4*>(c.partial) + (size_t)track * tiles_track * (2 * CY);
        const size_t woff = (size_t)track * c.Hmax * c.Wmax;
        const double* dn = c.denom + woff;
        float* mp = c.params->keep_maps ? c.maps + woff : nullptr;
        const int flat = t.flat;
        unsigned long long key = 0ull;
        for (int u = k * U + tid; u < min((k + 1) * U, U_total); u += kTilesPerCta) {
            const int tile = u / (2 * CY), r = u - tile * (2 * CY), i = r >> 1, half = r & 1;
            const int tb = tile / kTilesPerCta, band = tb / g.ctas_band, q = (tb - band * g.ctas_band) * kTilesPerCta + (tile - tb * kTilesPerCta);
            const int col = q / g.GB, grp = band * g.GB + (q - col * g.GB);
            const int y = grp * CY + i, x0 = col * 8 + half * 4;
            if (col < g.C && x0 < ww && grp < g.G && y < wh) {
                const float4* src = pbase + (size_t)tile * (2 * CY) + r;
                double d4[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) d4[e] = x0 + e < ww ? __ldcg(dn + (size_t)y * ww + x0 + e) : 0.0;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int p0 = 0; p0 < parts; p0 += 16) {             // 16 independent 16-byte loads in flight, added in part order
                    float4 pv[16];
#pragma unroll
                    for (int b = 0; b < 16; ++b) pv[b] = p0 + b < parts ? __ldcg(src + (size_t)(p0 + b) * plane4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int b = 0; b < 16; ++b)
                        if (p0 + b < parts) { acc.x += pv[b].x; acc.y += pv[b].y; acc.z += pv[b].z; acc.w += pv[b].w; }
                }
                const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (x0 + e < ww) {
                        const unsigned int idx = (unsigned int)(y * ww + x0 + e);
                        const float v = ncc_finalize(a4[e], d4[e], flat);
                        if (mp) mp[idx] = v;
                        const unsigned long long k2 = peak_key(v, idx);
                        key = k2 > key ? k2 : key;
                    }
                }
            }
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            unsigned long long o = shfl_xor_u64(key, m);
            key = o > key ? o : key;
        }
        if ((tid & 31) == 0 && key) atomicMax(&t.peak, key);
    }
    // the last CTA of this track (all peaks are in) performs the gate / EMA / state update
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int k = atomicAdd(&t.ticket, 1u);
        s_last = (k == (unsigned int)n_ctas - 1u);
        if (s_last) { t.ticket = 0u; t.arrive = 0u; t.stats_done = 0u; }   // everybody is past the waits
    }
    __syncthreads();
    if (stepped) trace_end(c, step, TR_FINALIZE);
    if (!s_last) return;
    __threadfence();
    if (c.trace && tid == 0) c.trace[((step % kRing) * 8 + TR_UPDATE) * 2] = gtime();
    track_update(c, track, step, stepped, reinterpret_cast<float*>(sm_raw), red);
    trace_end(c, step, TR_UPDATE);
}

__device__ __forceinline__ void cp_async4(float* dst, const float* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// (3b'') k_ncc_local: the single-stream (latency) search with the K-split INSIDE the CTA -- no partial sums in global memory, no
//   second-stage kernel, no statistics kernel.  A CTA owns a patch of 8 x (CY * TR) candidates of one track (TR thread tiles
//   stacked in y) and holds everything the patch needs in shared memory: its (CY*TR + th - 1) x (8 + tp) corner of the gray
//   plane and the whole centred template.  Thread = (thread tile tr, row part pdx, chunk part pjx): the same 8 x CY register
//   tile, loop and accumulation order as a K-split part of k_ncc_search (part = pdx * PJ + pjx covers template rows
//   [th pdx / PD, th (pdx+1) / PD) of chunks [nchunk pjx / PJ, ...)), so the scores are BIT-identical to the K-split path with
//   pj = PJ, pd = PD.  The parts meet in shared memory and are added in part order; before the loop all warps compute the
//   patch's window statistics from the same tile (FP64 column sums sliding down, then added along x: wsum exact, wsq to
//   ~1e-15 like k_winstats) -- the normalisers never leave the SM.  Per step: k_ingest_roi ~> k_ncc_local -> k_update.
//   Why: measured on B200 (C2), a K-split search CTA spends 8.9 of its 12 us in the FMA loop because ONE warp per scheduler
//   issues an FFMA only every ~2.4 cycles (two co-resident warps reach 87 % of the pipe); here every SM runs 8 warps, and the
//   3 us second stage plus a kernel boundary disappear.
struct LocalCfg {
    int TR;              // thread tiles (of CY candidate rows) per CTA; the patch is 8 x CY*TR candidates
    int PJ, PD;          // parts inside the CTA along the template's 8-column chunks and along its rows
    int bx, by;          // CTAs per track along x (ceil(Wmax / 8)) and y (ceil(ceil(Hmax / CY) / TR))
    int P;               // tile pitch in floats (8 + mtp rounded up to == 4 mod 8)
    int tileH;           // CY * TR + mth - 1
    int TS;              // template chunk stride in floats (mth * 8 + 4)
    int nfma;            // TR * PJ * PD threads run the FMA loop (the CTA is that rounded up to a warp)
    int gstats;          // 1: the normalisers come from k_winstats, which runs beside this kernel on SMs the plan leaves free
                         //    (TrackState.stats_done counts its CTAs); 0: computed here from the tile, before the loop
    int sNX, sNY;        // k_winstats' tile (StatCfg) -> how many of its CTAs store normalisers for a (clamped) window
    int update;          // 1: the track's last CTA (ticket) runs track_update; 0: k_update follows as its own launch
    int late_trigger;    // 1: griddepcontrol.launch_dependents behind the FMA loop (k_update launched with a programmatic dependency)
};
constexpr int kLocalRed = 8 * kCY + 4;   // floats per thread in the reduction buffer (44: conflict-free float4 stores)

template <int CY>
__global__ void __launch_bounds__(256, 1) k_ncc_local(Ctx c, LocalCfg g)
{
    extern __shared__ __align__(16) unsigned char sm_loc[];
    if (!g.late_trigger) pdl_trigger();
    const int per_track = g.bx * g.by;
    const int track = blockIdx.x / per_track, b = blockIdx.x - track * per_track;
    const int byi = b / g.bx, bxi = b - byi * g.bx;
    TrackState& t = c.tracks[track];
    unsigned long long step;
    const bool stepped = track_stepped_ld(c, t, step);            // all of the preamble's loads go out before the first branch
    const DevParams PP = *c.params;
    const int bx_ = t.x, by_ = t.y, tw = t.w, th = t.h, tp = t.tp, tstream = t.stream;
    __shared__ double red_u[64];
    __shared__ int s_last;
    if (!stepped) {
        // no search for this track in this step: its first CTA still reports it (and takes part in advancing the step)
        if (g.update && b == 0) track_update(c, track, step, false, reinterpret_cast<float*>(sm_loc), red_u);
        return;
    }
    trace_begin(c, step, TR_NCC);
    int win[4];
    search_window(bx_, by_, tw, th, c.W - tw + 1, c.H - th + 1, PP.rx, PP.ry, win);
    const int tid = threadIdx.x, lane = tid & 31;
    if (b == 0 && tid < 4) t.win[tid] = win[tid];               // k_update reads it (k_colprefix / k_winstats store it in the other shapes)
    const int ww = win[2], wh = win[3], nchunk = tp >> 3;
    const int px0 = 8 * bxi, py0 = CY * g.TR * byi;               // patch origin inside the window
    if (px0 >= ww || py0 >= wh) return;                           // clamped window: CTA-uniform
    const int nrow = min(CY * g.TR, wh - py0);                    // candidate rows of this patch
    const int P = g.P;
    float* s_tile = reinterpret_cast<float*>(sm_loc);             // [tileH][P]
    float* s_t = s_tile + (size_t)g.tileH * P;                    // [chunks][TS]
    float* s_red = s_t + (size_t)(c.mtp >> 3) * g.TS;             // [nfma][44] partial sums; before that: the statistics warp's scratch
    double* s_dn = reinterpret_cast<double*>(s_red + (size_t)g.nfma * kLocalRed);   // [CY * TR * 8] normalisers
    {
        // the centred template does not depend on this step's frame: on its way before the wait for the ingest
        const float4* tsrc = reinterpret_cast<const float4*>(c.templc + (size_t)track * c.mth * c.mtp);
        const int nq = th * 2;                                     // float4 per chunk
        for (int i = tid; i < nchunk * nq; i += blockDim.x) {
            const int j = i / nq, q = i - j * nq;
            cp_async16(reinterpret_cast<float4*>(s_t + (size_t)j * g.TS) + q, tsrc + (size_t)j * nq + q);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    pdl_wait();                                                    // the ingest's gray plane is complete
    {
        // tile: rows [py0, py0 + nrow + th - 1), columns [px0, px0 + 8 + tp) of the window's corner of the gray plane.  The window
        // origin has no alignment: aligned 16-byte loads from the origin rounded down to 4 pixels, scattered as scalars (the
        // shift costs nothing); zeros outside the frame (they only meet masked candidates / zero template columns)
        const int ax0 = win[0] + px0, ay0 = win[1] + py0, tw_used = 8 + tp, th_used = nrow + th - 1;
        const int gx0 = ax0 & ~3, nvec = (ax0 + tw_used - gx0 + 3) >> 2;
        const float* src = c.gray + (size_t)tstream * c.plane + (size_t)ay0 * c.pitch + gx0;
        const int total = th_used * nvec;
        for (int i0 = tid; i0 < total; i0 += 8 * (int)blockDim.x) {
            float4 v[8];
            int rr[8], xx[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {                          // eight independent loads in flight per thread: one round trip for a C2 patch
                const int i = i0 + k * (int)blockDim.x;
                rr[k] = -1;
                v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < total) {
                    const int r = i / nvec, q = i - r * nvec;
                    rr[k] = r; xx[k] = gx0 + 4 * q;
                    if (ay0 + r < c.H && xx[k] + 4 <= c.pitch) v[k] = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * c.pitch) + q);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (rr[k] >= 0) {
                    const float e4[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int x = xx[k] + e - ax0;
                        if (x >= 0 && x < tw_used) s_tile[rr[k] * P + x] = (xx[k] + e < c.W) ? e4[e] : 0.f;
                    }
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    // phase stamps of CTA 0 (pvt_trace_enable; tools/timeline.py): FRINGE slot = staged | FMA loop done, TAIL slot = statistics done | reduced
    unsigned long long* trc = (c.trace && blockIdx.x == 0) ? c.trace + (step % kRing) * 16 : nullptr;
    if (trc && tid == 0) trc[TR_FRINGE * 2] = gtime();

    if (!g.gstats) {
        // ---- the patch's window statistics from the tile in shared memory, by all warps (scratch: the reduction buffer, which is
        // unused until barrier A).  FP64 adds have ~40 cycles of latency here and one warp issues them slowly (a dedicated
        // statistics warp beside the FMA loop took 10 - 13 us), so the work is cut into short chains over all 256 threads:
        // (1) three partial column sums per tile column, (2) the column's box sum sliding down the patch's candidate rows,
        // (3) one thread per candidate adds its tw column sums with four accumulators and applies OpenCV's normaliser.
        const int ncol = 8 + tw - 1, seg = (th + 2) / 3;
        double* cS = reinterpret_cast<double*>(s_red);            // [nrow][ncol] vertical box sums of f
        double* cQ = cS + (size_t)nrow * ncol;                     // ... of f^2
        double* part = cQ + (size_t)nrow * ncol;                   // [3][ncol][2]
        for (int w = tid; w < 3 * ncol; w += blockDim.x) {
            const int sgi = w / ncol, col = w - sgi * ncol, r0 = sgi * seg, r1 = min(th, r0 + seg);
            double S0 = 0.0, S1 = 0.0, Q0 = 0.0, Q1 = 0.0;
            int r = r0;
            for (; r + 1 < r1; r += 2) {
                const double a = (double)s_tile[r * P + col], b2 = (double)s_tile[(r + 1) * P + col];
                S0 += a; Q0 += a * a;
                S1 += b2; Q1 += b2 * b2;
            }
            if (r < r1) { const double a = (double)s_tile[r * P + col]; S0 += a; Q0 += a * a; }
            part[(size_t)w * 2] = S0 + S1;
            part[(size_t)w * 2 + 1] = Q0 + Q1;
        }
        __syncthreads();
        if (trc && tid == 0) trc[TR_COLPREFIX * 2] = gtime();
        for (int col = tid; col < ncol; col += blockDim.x) {
            double S = part[(size_t)col * 2] + part[(size_t)(ncol + col) * 2] + part[(size_t)(2 * ncol + col) * 2];
            double Q = part[(size_t)col * 2 + 1] + part[(size_t)(ncol + col) * 2 + 1] + part[(size_t)(2 * ncol + col) * 2 + 1];
            for (int y = 0; y < nrow; ++y) {
                cS[(size_t)y * ncol + col] = S;
                cQ[(size_t)y * ncol + col] = Q;
                if (y + 1 < nrow) {
                    const double di = (double)s_tile[(y + th) * P + col], dq = (double)s_tile[y * P + col];
                    S += di - dq;                                 // exact: sums of u8-sourced f fit a double (k_winstats)
                    Q += di * di - dq * dq;
                }
            }
        }
        __syncthreads();
        if (trc && tid == 0) trc[TR_COLPREFIX * 2 + 1] = gtime();
        const double invArea = 1.0 / ((double)th * (double)tw), tn = t.templ_norm;
        for (int idx = tid; idx < nrow * 8; idx += blockDim.x) {
            const double* rs = cS + (size_t)(idx >> 3) * ncol + (idx & 7);
            const double* rq = cQ + (size_t)(idx >> 3) * ncol + (idx & 7);
            double S[4] = {0.0, 0.0, 0.0, 0.0}, Q[4] = {0.0, 0.0, 0.0, 0.0};
            int k = 0;
            for (; k + 3 < tw; k += 4) {
#pragma unroll
                for (int m = 0; m < 4; ++m) { S[m] += rs[k + m]; Q[m] += rq[k + m]; }
            }
            for (; k < tw; ++k) { S[0] += rs[k]; Q[0] += rq[k]; }
            s_dn[idx] = window_denom((S[0] + S[1]) + (S[2] + S[3]), (Q[0] + Q[1]) + (Q[2] + Q[3]), invArea, tn, c.formula);
        }
        // (no barrier needed here: the scratch is next written behind barrier A, s_dn next read behind barrier B)
    }
    if (trc && tid == 0) trc[TR_TAIL * 2] = gtime();

    float acc[CY][8];
#pragma unroll
    for (int i = 0; i < CY; ++i)
#pragma unroll
        for (int cx = 0; cx < 8; ++cx) acc[i][cx] = 0.f;
    if (tid < g.nfma) {
        const int pjx = tid % g.PJ, pdx = (tid / g.PJ) % g.PD, tr = tid / (g.PJ * g.PD);
        const int d0 = (th * pdx) / g.PD, d1 = (th * (pdx + 1)) / g.PD, nd = d1 - d0;
        const int j0 = (nchunk * pjx) / g.PJ, j1 = nd > 0 ? (nchunk * (pjx + 1)) / g.PJ : j0;
        if (tr * CY < nrow) {
            const float* base = s_tile + (size_t)(tr * CY + d0) * P;
            for (int j = j0; j < j1; ++j) {
                const float* st = s_t + (size_t)j * g.TS + d0 * 8;
                const float* fb = base + j * 8;
                float racc[CY][8], w[CY][16];
#pragma unroll
                for (int i = 0; i < CY; ++i)
#pragma unroll
                    for (int cx = 0; cx < 8; ++cx) racc[i][cx] = 0.f;
#pragma unroll
                for (int r = 0; r < CY - 1; ++r) {
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float4 a = *reinterpret_cast<const float4*>(fb + r * P + 4 * v);
                        w[r][4 * v] = a.x; w[r][4 * v + 1] = a.y; w[r][4 * v + 2] = a.z; w[r][4 * v + 3] = a.w;
                    }
                }
#pragma unroll 1
                for (int e0 = 0; e0 < nd; e0 += CY) {
#pragma unroll
                    for (int u = 0; u < CY; ++u) {
                        const int e = e0 + u;                      // template row d0 + e
                        if (e < nd) {
                            float(&wn)[16] = w[(u + CY - 1) % CY];
                            const float* pr = fb + (size_t)(e + CY - 1) * P;
#pragma unroll
                            for (int v = 0; v < 4; ++v) {
                                const float4 a = *reinterpret_cast<const float4*>(pr + 4 * v);
                                wn[4 * v] = a.x; wn[4 * v + 1] = a.y; wn[4 * v + 2] = a.z; wn[4 * v + 3] = a.w;
                            }
                            float tt[8];
                            {
                                const float4 a = *reinterpret_cast<const float4*>(st + e * 8);
                                const float4 bq = *reinterpret_cast<const float4*>(st + e * 8 + 4);
                                tt[0] = a.x; tt[1] = a.y; tt[2] = a.z; tt[3] = a.w; tt[4] = bq.x; tt[5] = bq.y; tt[6] = bq.z; tt[7] = bq.w;
                            }
#pragma unroll
                            for (int k = 0; k < 8; ++k)
#pragma unroll
                                for (int i = 0; i < CY; ++i)
#pragma unroll
                                    for (int cx = 0; cx < 8; ++cx) racc[i][cx] = fmaf(w[(u + i) % CY][k + cx], tt[k], racc[i][cx]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < CY; ++i)
#pragma unroll
                    for (int cx = 0; cx < 8; ++cx) acc[i][cx] += racc[i][cx];
            }
        }
    }
    if (trc && tid == 0) trc[TR_FRINGE * 2 + 1] = gtime();
    // k_update is launched behind this kernel with a programmatic dependency: released HERE (when every CTA is past its loop), its CTA
    // becomes resident during the reduction -- not during the loop, where an early CTA was measured to slow the search -- reads the
    // step, the track and the old template, and blocks in griddepcontrol.wait until this grid has completed
    if (g.late_trigger) pdl_trigger();
    __syncthreads();                                               // (A) statistics done: the scratch becomes the reduction buffer
    if (tid < g.nfma) {
        float4* po = reinterpret_cast<float4*>(s_red + (size_t)tid * kLocalRed);
#pragma unroll
        for (int i = 0; i < CY; ++i) {
            po[2 * i] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            po[2 * i + 1] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
        }
    }
    // (measured and dropped: the spare lanes behind the FMA threads of the last warp polling the statistics count and staging the
    //  normalisers during the loop -- the divergent halves of that warp run one after the other: peak +15.1 instead of +10.8 us)
    if (g.gstats && tid == 0) {                                    // k_winstats (the other graph branch) has stored this window's normalisers?
        const unsigned int n_stat = (unsigned int)(((ww + g.sNX - 1) / g.sNX) * ((wh + g.sNY - 1) / g.sNY));
        if (!spin_until(&t.stats_done, n_stat)) { *c.fault = 1u; __threadfence_system(); }
    }
    __syncthreads();                                               // (B)
    unsigned long long key = 0ull;
    if (tid < nrow * 8) {
        const int y = tid >> 3, x = tid & 7, tr = y / CY, i = y - tr * CY;
        if (px0 + x < ww) {
            const unsigned int idx = (unsigned int)((py0 + y) * ww + px0 + x);
            const double dnv = g.gstats ? __ldcg(c.denom + (size_t)track * c.Hmax * c.Wmax + idx) : s_dn[tid];   // on its way during the sums
            float a = 0.f;
            const int parts = g.PJ * g.PD;
            const float* src = s_red + (size_t)(tr * parts) * kLocalRed + i * 8 + x;
#pragma unroll 8
            for (int p = 0; p < parts; ++p) a += src[(size_t)p * kLocalRed];    // part order: p = pdx * PJ + pjx, as k_ncc_finalize adds them
            const float v = ncc_finalize(a, dnv, t.flat);
            if (c.params->keep_maps) c.maps[(size_t)track * c.Hmax * c.Wmax + idx] = v;
            key = peak_key(v, idx);
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        unsigned long long o = shfl_xor_u64(key, m);
        key = o > key ? o : key;
    }
    if (lane == 0 && key) atomicMax(&t.peak, key);
    if (trc && tid == 0) trc[TR_TAIL * 2 + 1] = gtime();
    trace_end(c, step, TR_NCC);
    if (!g.update) return;
    // the last CTA of this track through the ticket (all peaks are in) runs the gate / EMA / state update: no k_update launch,
    // no kernel boundary between the peak and the update (~2 us on B200).  CTAs outside a clamped window left early and do not count.
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int n_cta = (unsigned int)(((ww + 7) >> 3) * ((wh + CY * g.TR - 1) / (CY * g.TR)));
        const unsigned int k = atomicAdd(&t.ticket, 1u);
        s_last = (k == n_cta - 1u);
        if (s_last) t.ticket = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (c.trace && tid == 0) c.trace[((step % kRing) * 8 + TR_UPDATE) * 2] = gtime();
    track_update(c, track, step, true, reinterpret_cast<float*>(sm_loc), red_u);
    trace_end(c, step, TR_UPDATE);
}

// (3c) k_ncc_fringe: the candidates the thread-tile grid leaves out (TileCfg: the single column x = 8C and/or the
//      single row y = CY*G -- both exist for the default 161 x 161 window), register-blocked like the search itself.
//      A thread owns 8 consecutive candidates of the fringe row (8 x 1 tile) or of the fringe column (1 x 8 tile) and ONE
//      template chunk: per template row it fetches the chunk's 8 template values (2 LDS.128) and
//        row tile:    the 16-float window of that frame row (4 LDS.128)                 -> acc[cx] += w[k + cx] * t[k]
//        column tile: ONE new 8-float frame row; the other seven stay in registers (2 LDS.128) -> acc[yy] += blk[yy][k] * t[k]
//      i.e. 64 FMAs for 4 - 6 shared loads, eight independent chains per thread.  Every chain runs dy outer, k inner,
//      exactly as in k_ncc_search; one thread per candidate finally adds the chunk results in k_ncc_search's order
//      INCLUDING the K-split part structure (part p = pdx*pj + pjx covers template rows [d0, d1) of chunks [j0, j1);
//      parts added in part order as k_ncc_finalize does): equal windows get equal scores wherever they are computed.
//      A CTA takes up to FringeCfg.tpc tiles of one kind and one track; thread = (tile, chunk).  The strip of the gray
//      plane the tiles cover and the centred template are staged in shared memory with cp.async (4-byte copies: the
//      window origin has no alignment).  The strip pitch has an odd number of 16-byte groups, so lanes on consecutive
//      strip rows (column tiles) read vectors without bank conflicts.
//      Scheduling: in the throughput shape the kernel is launched right behind k_ncc_search with a programmatic
//      dependency (it needs none of the search's results), so its CTAs start when the last search CTA has been
//      dispatched and fill the SM slots that free up during the search's ragged last round.  In the latency shape it is
//      a third graph branch beside the statistics and the search (FringeCfg.defer).
//      Shared memory: strip | template [nchunk][th * 8 + 4] | chain results [nchunk][8 * tpc].
__host__ __device__ inline int fringe_pitch(int sw)
{
    const int p = (sw + 3) & ~3;
    return ((p >> 2) & 1) ? p : p + 4;
}

struct FringeCfg {
    int tpc;             // tiles (of 8 candidates) per CTA
    int colg, rowg;      // CTAs per track for the fringe column (ceil(ceil(Hmax / 8) / tpc)) and the fringe row; 0 = none
    int strip_floats;    // shared floats reserved for the strip
    int threads;         // tpc * (mtp / 8) rounded up to a warp
    int defer;           // 1 (K-split / latency shape): grid.z = pd, a CTA computes ONE row part pdx = blockIdx.z and stores the
                         //    raw partial cross terms of its parts to Ctx.fringe_acc; k_ncc_finalize adds them in part order and
                         //    normalises, so the kernel needs no window statistics and runs beside them and the search
};

__device__ __forceinline__ void load8(float (&d)[8], const float* p)
{
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}

constexpr int kFringeThreads = 256;
__global__ void __launch_bounds__(kFringeThreads) k_ncc_fringe(Ctx c, TileCfg g, FringeCfg fc)
{
    extern __shared__ __align__(16) float sm_fr[];
    __shared__ int s_d[34];                                         // K-split row bounds: part pdx covers rows [s_d[pdx], s_d[pdx+1])
    const int track = blockIdx.y;
    TrackState& t = c.tracks[track];
    const unsigned long long step = *c.step;
    if (!track_stepped(c, t, step)) return;
    // the window is derived here (not read from t.win): see k_ncc_search
    int win[4];
    {
        const DevParams P = *c.params;
        search_window(t.x, t.y, t.w, t.h, c.W - t.w + 1, c.H - t.h + 1, P.rx, P.ry, win);
    }
    trace_begin(c, step, TR_FRINGE);
    const int ww = win[2], wh = win[3];
    const int fx = 8 * g.C, fy = kCY * g.G;                        // the grid covers [0, fx) x [0, fy)
    const bool is_col = (int)blockIdx.x < fc.colg;
    const int grp = is_col ? blockIdx.x : blockIdx.x - fc.colg;
    int n;                                                          // candidates of this kind
    if (is_col) { if (ww <= fx) return; n = wh; }                   // (fx, y), y = 0 .. wh-1
    else { if (wh <= fy) return; n = min(ww, fx); }                 // (x, fy), x = 0 .. min(ww, fx)-1
    const int c0 = grp * fc.tpc * 8;
    if (c0 >= n) return;
    const int nc = min(fc.tpc * 8, n - c0), ntile = (nc + 7) >> 3;
    const int th = t.h, tp = t.tp, nchunk = tp >> 3, pj = g.pj, pd = g.pd;
    if ((int)threadIdx.x <= pd) s_d[threadIdx.x] = (th * (int)threadIdx.x) / pd;
    const int pz = fc.defer ? (int)blockIdx.z : 0;                  // the row part this CTA computes (non-deferred: pd == 1)
    // strip: window-relative origin (sx0, sy0), sw x sh floats; it spans the PADDED template width tp and whole tiles.
    // Elements outside the frame are stored as 0 (they only meet masked candidates or zero template columns).
    const int sx0 = is_col ? fx : c0, sy0 = is_col ? c0 : fy;
    const int sw = is_col ? tp : ntile * 8 + tp, sh = is_col ? ntile * 8 + th - 1 : th;
    const int SP = fringe_pitch(sw);
    float* s_strip = sm_fr;
    float* s_t = sm_fr + fc.strip_floats;
    const int TS = th * 8 + 4;                                      // template chunk stride: lanes on consecutive chunks read
                                                                    // consecutive 16-byte bank groups
    float* s_r = s_t + (size_t)(c.mtp >> 3) * (c.mth * 8 + 4);
    const int ncand_pad = fc.tpc * 8;
    {
        const int ax0 = win[0] + sx0, ay0 = win[1] + sy0;
        const float* src = c.gray + (size_t)t.stream * c.plane + (size_t)ay0 * c.pitch + ax0;
        const int nwarp = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
        // only what this CTA's row part [d0, d1) touches (everything when there is a single part)
        const int d0 = (th * pz) / pd, d1 = fc.defer ? (th * (pz + 1)) / pd : th;
        const int r_lo = d0, r_hi = is_col ? min(sh, ntile * 8 + d1 - 1) : d1;
        for (int r = r_lo + wid; r < r_hi; r += nwarp)
            for (int x = lane; x < sw; x += 32) {
                if (ay0 + r < c.H && ax0 + x < c.W) cp_async4(s_strip + r * SP + x, src + (size_t)r * c.pitch + x);
                else s_strip[r * SP + x] = 0.f;
            }
        const float4* tsrc = reinterpret_cast<const float4*>(c.templc + (size_t)track * c.mth * c.mtp);
        float4* tdst = reinterpret_cast<float4*>(s_t);
        const int nq = (d1 - d0) * 2;                               // float4 per chunk
        for (int i = threadIdx.x; i < nchunk * nq; i += blockDim.x) {
            const int j = i / nq, q = d0 * 2 + (i - j * nq);
            cp_async16(tdst + j * (TS >> 2) + q, tsrc + j * th * 2 + q);
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    {
        // thread = (chunk, tile), chunk fastest: neighbouring lanes read neighbouring 32-byte pieces of a strip row
        const int nchm = c.mtp >> 3, j = threadIdx.x % nchm, tile = threadIdx.x / nchm;
        if (tile < ntile && j < nchunk) {
            const float* tj = s_t + (size_t)j * TS;
            float* out = s_r + (size_t)j * ncand_pad + tile * 8;
            if (!is_col) {
                const float* base = s_strip + tile * 8 + j * 8;                  // 16-byte aligned: c0, tile*8, j*8 are multiples of 8
                {
                    const int pdx = pz;
                    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
                    for (int dy = s_d[pdx]; dy < s_d[pdx + 1]; ++dy) {
                        float w[16], tt[8];
                        load8(*reinterpret_cast<float(*)[8]>(w), base + dy * SP);
                        load8(*reinterpret_cast<float(*)[8]>(w + 8), base + dy * SP + 8);
                        load8(tt, tj + dy * 8);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
#pragma unroll
                            for (int cx = 0; cx < 8; ++cx) acc[cx] = fmaf(w[k + cx], tt[k], acc[cx]);
                    }
                    *reinterpret_cast<float4*>(out) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    *reinterpret_cast<float4*>(out + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
                }
            } else {
                const float* base = s_strip + (size_t)(tile * 8) * SP + j * 8;   // rows tile*8 + yy + dy
                {
                    const int pdx = pz;
                    const int d0 = s_d[pdx], nd = s_d[pdx + 1] - d0;
                    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    float blk[8][8];                                             // ring of 8 frame rows, 8 floats each
#pragma unroll
                    for (int r = 0; r < 7; ++r) load8(blk[r], base + (size_t)(d0 + r) * SP);
#pragma unroll 1
                    for (int e0 = 0; e0 < nd; e0 += 8) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int e = e0 + u;
                            if (e < nd) {
                                load8(blk[(u + 7) & 7], base + (size_t)(d0 + e + 7) * SP);
                                float tt[8];
                                load8(tt, tj + (d0 + e) * 8);
#pragma unroll
                                for (int k = 0; k < 8; ++k)
#pragma unroll
                                    for (int yy = 0; yy < 8; ++yy) acc[yy] = fmaf(blk[(u + yy) & 7][k], tt[k], acc[yy]);
                            }
                        }
                    }
                    *reinterpret_cast<float4*>(out) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    *reinterpret_cast<float4*>(out + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
                }
            }
        }
    }
    __syncthreads();
    unsigned long long key = 0ull;
    for (int l = threadIdx.x; l < nc; l += blockDim.x) {
        const int x = is_col ? fx : c0 + l, y = is_col ? c0 + l : fy;
        if (fc.defer) {
            // the parts (pjx, pz) of this candidate: chunk results added in chunk order, as a K-split search CTA does
            const bool empty = s_d[pz + 1] <= s_d[pz];
            for (int pjx = 0; pjx < pj; ++pjx) {
                const int j0 = (nchunk * pjx) / pj, j1 = empty ? j0 : (nchunk * (pjx + 1)) / pj;
                float pacc = 0.f;
                for (int j = j0; j < j1; ++j) pacc += s_r[(size_t)j * ncand_pad + l];
                c.fringe_acc[((size_t)track * pj * pd + (size_t)pz * pj + pjx) * (c.Hmax + c.Wmax) + (is_col ? y : c.Hmax + x)] = pacc;
            }
        } else {
            float acc = 0.f;
            for (int j = 0; j < nchunk; ++j) acc += s_r[(size_t)j * ncand_pad + l];
            const unsigned int idx = (unsigned int)(y * ww + x);
            const double dn = c.denom[(size_t)track * c.Hmax * c.Wmax + idx];
            const float v = ncc_finalize(acc, dn, t.flat);
            if (c.params->keep_maps) c.maps[(size_t)track * c.Hmax * c.Wmax + idx] = v;
            const unsigned long long k2 = peak_key(v, idx);
            key = k2 > key ? k2 : key;
        }
    }
    if (!fc.defer) {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            unsigned long long o = shfl_xor_u64(key, m);
            key = o > key ? o : key;
        }
        if ((threadIdx.x & 31) == 0 && key) atomicMax(&t.peak, key);
    }
    trace_end(c, step, TR_FRINGE);
}

// Tail items' second stage (throughput mode): one CTA per tail item, thread = thread tile of k_ncc_search; adds the
// tail_ps partial sums of its 8 x CY candidates in part order, then the same FP64 normalisation / peak epilogue.
template <int CY>
__global__ void __launch_bounds__(kTilesPerCta) k_ncc_tail_finalize(Ctx c, TileCfg g)
{
    const int item = g.n_full + blockIdx.x;
    const int track = item / g.cpt, blk = item - track * g.cpt;
    TrackState& t = c.tracks[track];
    const unsigned long long step = *c.step;
    if (!track_stepped(c, t, step)) return;
    trace_begin(c, step, TR_TAIL);
    const int ww = t.win[2], wh = t.win[3];
    const int band = blk / g.ctas_band, q = (blk - band * g.ctas_band) * kTilesPerCta + threadIdx.x;
    const int col = q / g.GB, grp = band * g.GB + (q - col * g.GB);
    unsigned long long key = 0ull;
    if (col < g.C && col * 8 < ww && grp < g.G && grp * CY < wh) {
        float acc[CY][8];
#pragma unroll
        for (int i = 0; i < CY; ++i)
#pragma unroll
            for (int cx = 0; cx < 8; ++cx) acc[i][cx] = 0.f;
        // four parts' partial sums are fetched per batch (40 independent 16-byte loads in flight: one L2 round trip per
        // batch instead of one per part), then added in part order
        for (int p0 = 0; p0 < g.tail_ps; p0 += 4) {
            float4 v[4][2 * CY];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const float4* pi = reinterpret_cast<const float4*>(c.partial + (((size_t)blockIdx.x * g.tail_ps + min(p0 + b, g.tail_ps - 1)) * kTilesPerCta + threadIdx.x) * (8 * CY));
#pragma unroll
                for (int i = 0; i < 2 * CY; ++i) v[b][i] = __ldg(pi + i);
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (p0 + b < g.tail_ps) {
#pragma unroll
                    for (int i = 0; i < CY; ++i) {
                        acc[i][0] += v[b][2 * i].x; acc[i][1] += v[b][2 * i].y; acc[i][2] += v[b][2 * i].z; acc[i][3] += v[b][2 * i].w;
                        acc[i][4] += v[b][2 * i + 1].x; acc[i][5] += v[b][2 * i + 1].y; acc[i][6] += v[b][2 * i + 1].z; acc[i][7] += v[b][2 * i + 1].w;
                    }
                }
            }
        }
        const size_t woff = (size_t)track * c.Hmax * c.Wmax;
        const double* dn = c.denom + woff;
        float* mp = c.params->keep_maps ? c.maps + woff : nullptr;
        const int flat = t.flat;
#pragma unroll
        for (int i = 0; i < CY; ++i) {
            const int y = grp * CY + i;
            if (y < wh) {
                double d8[8];
#pragma unroll
                for (int cx = 0; cx < 8; ++cx) {
                    const int x = col * 8 + cx;
                    d8[cx] = x < ww ? __ldg(dn + y * ww + x) : 0.0;
                }
#pragma unroll
                for (int cx = 0; cx < 8; ++cx) {
                    const int x = col * 8 + cx;
                    if (x < ww) {
                        const unsigned int idx = (unsigned int)(y * ww + x);
                        const float v = ncc_finalize(acc[i][cx], d8[cx], flat);
                        if (mp) mp[idx] = v;
                        const unsigned long long k = peak_key(v, idx);
                        key = k > key ? k : key;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        unsigned long long o = shfl_xor_u64(key, m);
        key = o > key ? o : key;
    }
    if ((threadIdx.x & 31) == 0 && key) atomicMax(&t.peak, key);
    trace_end(c, step, TR_TAIL);
}

// K-split second stage: add the parts' partial sums in part order, normalise, pick the peak.
// NOTE the summation order differs from the unsplit kernel (parts regroup template rows), so scores may differ
// from it in the last bits; within one configuration every candidate is summed identically (exact ties stay exact).
constexpr int kFinalizeThreads = 512;   // upper bound; also the width of the fused update: launched 512 wide for templates above
                                        // 8192 pixels (a 128 x 128 template: two EMA batches instead of four, C3 -5 %), else 256
__global__ void __launch_bounds__(kFinalizeThreads) k_ncc_finalize(Ctx c, TileCfg g)
{
    extern __shared__ float sm_f[];
    __shared__ double red[64];
    __shared__ int s_last;
    const int track = blockIdx.y, parts = g.pj * g.pd;
    TrackState& t = c.tracks[track];
    const unsigned long long step = *c.step;
    const bool stepped = track_stepped(c, t, step);
    pdl_wait();   // launched behind the search with a programmatic dependency (K-split shape): the partial sums are complete
    trace_begin(c, step, TR_FINALIZE);
    if (stepped) {
        const int ww = t.win[2], n = ww * t.win[3];
        const int idx = blockIdx.x * blockDim.x + threadIdx.x;
        unsigned long long key = 0ull;
        if (idx < n) {
            const int y = idx / ww, x = idx - y * ww;
            // where this candidate's partial cross terms are: (a) fringe candidate -> Ctx.fringe_acc, one float per part
            // (k_ncc_fringe); (b) grid candidate -> (thread tile, slot) of k_ncc_search, column-major tiles inside row
            // bands.  One common load loop, so that warps holding both kinds do not pay two L2 round trips.
            const float* src;
            size_t stride;
            if (x >= 8 * g.C || y >= kCY * g.G) {
                stride = (size_t)(c.Hmax + c.Wmax);
                src = c.fringe_acc + (size_t)track * parts * stride + (x >= 8 * g.C ? y : c.Hmax + x);
            } else {
                const int col = x >> 3, cx = x & 7;
                const int grp = y / kCY, i = y - grp * kCY;
                const int band = grp / g.GB, gl = grp - band * g.GB;
                const size_t tile = (size_t)band * g.ctas_band * kTilesPerCta + (size_t)col * g.GB + gl;
                const size_t tiles_track = (size_t)g.bands * g.ctas_band * kTilesPerCta;
                stride = (size_t)c.max_tracks * tiles_track * (8 * kCY);
                src = c.partial + ((size_t)track * tiles_track + tile) * (8 * kCY) + i * 8 + cx;
            }
            const size_t woff = (size_t)track * c.Hmax * c.Wmax;
            const double dnv = __ldg(c.denom + woff + idx);
            float acc = 0.f;
            for (int p0 = 0; p0 < parts; p0 += 32) {           // 32 independent loads in flight, added in part order
                float pv[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) pv[k] = (p0 + k < parts) ? __ldg(src + (size_t)(p0 + k) * stride) : 0.f;
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (p0 + k < parts) acc += pv[k];
            }
            const float v = ncc_finalize(acc, dnv, t.flat);
            if (c.params->keep_maps) c.maps[woff + idx] = v;
            key = peak_key(v, (unsigned int)idx);
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            unsigned long long o = shfl_xor_u64(key, m);
            key = o > key ? o : key;
        }
        if ((threadIdx.x & 31) == 0 && key) atomicMax(&t.peak, key);
    }
    // the last CTA of this track (all peaks are in) performs the gate / EMA / state update: no extra launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int k = atomicAdd(&t.ticket, 1u);
        s_last = (k == gridDim.x - 1);
        if (s_last) t.ticket = 0u;
    }
    __syncthreads();
    trace_end(c, step, TR_FINALIZE);
    if (!s_last) return;
    __threadfence();
    if (c.trace && threadIdx.x == 0) c.trace[((step % kRing) * 8 + TR_UPDATE) * 2] = gtime();
    track_update(c, track, step, stepped, sm_f, red);
    trace_end(c, step, TR_UPDATE);
}

// =============================================================================================
// (4) track initialisation / state injection (single CTA): main.cpp:70-71  templ = frame_gray_f32(bbox).clone()
//     template statistics follow cv::meanStdDev in double (mean = s/N, sigma^2 = max(sq/N - mean^2, 0));
//     matchTemplate returns an all-ones map if sigma^2 < DBL_EPSILON; templNorm = sqrt(sigma^2)/sqrt(1/N).
// =============================================================================================
__global__ void __launch_bounds__(256) k_track_init(Ctx c, int track, int stream, int x, int y, int w, int h)
{
    extern __shared__ float sm_f[];
    __shared__ double red[64];
    TrackState& t = c.tracks[track];
    const float* g = c.gray + (size_t)stream * c.plane;
    float* tp_ = c.templ + (size_t)track * c.mth * c.mtw;
    for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
        const int r = i / w, col = i - r * w;
        const float v = g[(size_t)(y + r) * c.pitch + x + col];
        tp_[i] = v;
        sm_f[i] = v;
    }
    if (threadIdx.x == 0) {
        t.active = 1; t.stream = stream; t.x = x; t.y = y; t.w = w; t.h = h; t.peak = 0ull; t.ticket = 0u; t.arrive = 0u; t.stats_done = 0u;
        t.lost_count = 0; t.use_global = 0; t.global_since = 0ull;
        t.win[0] = t.win[1] = t.win[2] = t.win[3] = 0;
    }
    __syncthreads();
    finish_template(c, track, t, sm_f, red, w, h);
}

// re-derive statistics after pvt_set_state wrote bbox/template from the host
__global__ void __launch_bounds__(256) k_track_refresh(Ctx c, int track)
{
    extern __shared__ float sm_f[];
    __shared__ double red[64];
    TrackState& t = c.tracks[track];
    const float* tp_ = c.templ + (size_t)track * c.mth * c.mtw;
    const int w = t.w, h = t.h;
    for (int i = threadIdx.x; i < w * h; i += blockDim.x) sm_f[i] = tp_[i];
    __syncthreads();
    finish_template(c, track, t, sm_f, red, w, h);
}

// PVT_KERNEL_TC_GLOBAL: the local pass runs the FP32 kernels and its update does not maintain the template's 8-bit digits (4 us of
// a 38 us step); the whole-frame pass derives them here, right before k_ncc_tc, for the tracks it owns in this step.
__global__ void __launch_bounds__(256) k_track_digits(Ctx c)
{
    extern __shared__ float sm_f[];
    const int track = blockIdx.x;
    TrackState& t = c.tracks[track];
    unsigned long long step;
    if (!track_stepped_ld(c, t, step)) return;
    const float* tp_ = c.templ + (size_t)track * c.mth * c.mtw;
    const int w = t.w, h = t.h;
    const double mean = t.mean;
    for (int i = threadIdx.x; i < w * h; i += blockDim.x) sm_f[i] = tp_[i];
    __syncthreads();
    write_digits(c, track, t, sm_f, mean, w, h);
}

// =============================================================================================
// (5) track_update: peak -> gates -> bbox -> EMA -> next frame's template statistics, by ONE CTA per track.
//     main.cpp:150-161.  bestVal is the float peak widened to double and compared in double
//     (0.7f < 0.7, so a float compare would flip decisions).  cv::addWeighted on CV_32F:
//     templ' = (float) fma((double)templ, 1-lr, (double)patch * lr)   -- bit-exact (tests G5).
//     The new template goes through shared memory once: EMA, its FP64 sum / sum of squares (cv::meanStdDev),
//     then the centred chunk-major copy -- no second trip to HBM.
//     The CTA that finishes the last track of the step advances the device step counter, so the next graph
//     launch picks the next frame-table entry without any host action.
//     Called by k_update (its own launch) or by the last k_ncc_finalize CTA of the track (K-split mode).
// =============================================================================================
__device__ void block_sum2(double& s, double& q, double* red /* 2 * 32 doubles */)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        s += __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(s), m), __shfl_xor_sync(0xffffffffu, __double2loint(s), m));
        q += __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(q), m), __shfl_xor_sync(0xffffffffu, __double2loint(q), m));
    }
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[w] = s; red[32 + w] = q; }
    __syncthreads();
    s = 0.0; q = 0.0;
    for (int i = 0; i < nw; ++i) { s += red[i]; q += red[32 + i]; }   // same order in every thread: identical result
}

// PVT_KERNEL_TC: the centred template tc = fl32(t - mean) in 16-bit fixed point, q = rint(tc * 2^k) with k the largest
// power of two that keeps |q| <= 127 * 256, as two signed 8-bit digits q = 256 d1 + d0 (operand B of k_ncc_tc), plus what
// the epilogue needs to undo the scale and the DC part of the rounding: 2^-k and (sum(q) 2^-k - sum(tc)) / N.
__device__ void write_digits(const Ctx& c, int track, TrackState& t, const float* s_t, double mean, int tw, int th)
{
    __shared__ double r3[3][32];
    const int n = tw * th, tid = threadIdx.x, nw = (blockDim.x + 31) >> 5, w = tid >> 5;
    double mx = 0.0, st = 0.0;
    for (int i = tid; i < n; i += blockDim.x) {
        const double v = (double)(float)((double)s_t[i] - mean);
        mx = fmax(mx, fabs(v));
        st += v;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        mx = fmax(mx, __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(mx), m), __shfl_xor_sync(0xffffffffu, __double2loint(mx), m)));
        st += __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(st), m), __shfl_xor_sync(0xffffffffu, __double2loint(st), m));
    }
    __syncthreads();
    if ((tid & 31) == 0) { r3[0][w] = mx; r3[1][w] = st; }
    __syncthreads();
    mx = 0.0; st = 0.0;
    for (int i = 0; i < nw; ++i) { mx = fmax(mx, r3[0][i]); st += r3[1][i]; }
    int k = 0;
    if (mx > 0.0) {
        int e;
        const double fr = frexp(32512.0 / mx, &e);    // 32512 / mx = fr * 2^e, fr in [0.5, 1)  ->  floor(log2) = e - 1
        (void)fr;
        k = e - 1;
        if (k > 60) k = 60;
    }
    const double scale = ldexp(1.0, k), inv = ldexp(1.0, -k);
    signed char* d0 = c.tdig + (size_t)track * 2 * c.mth * c.tpp;
    signed char* d1 = d0 + (size_t)c.mth * c.tpp;
    double sq = 0.0;
    const int tpp = c.tpp;
    for (int i = tid; i < th * tpp; i += blockDim.x) {
        const int y = i / tpp, x = i - y * tpp;
        int q = 0;
        if (x < tw) q = __double2int_rn((double)(float)((double)s_t[y * tw + x] - mean) * scale);
        const int hi = (q + 128) >> 8, lo = q - 256 * hi;
        d0[(size_t)y * tpp + x] = (signed char)lo;
        d1[(size_t)y * tpp + x] = (signed char)hi;
        sq += (double)q;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1)
        sq += __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(sq), m), __shfl_xor_sync(0xffffffffu, __double2loint(sq), m));
    if ((tid & 31) == 0) r3[2][w] = sq;
    __syncthreads();
    if (tid == 0) {
        sq = 0.0;
        for (int i = 0; i < nw; ++i) sq += r3[2][i];          // integers below 2^53: exact in any order
        t.tc_inv = inv;
        t.tc_dc = (sq * inv - st) / (double)n;
    }
}

// statistics + centred chunk-major template from the template held in shared memory (s_t, th*tw floats)
// (tw, th) = (t.w, t.h), passed in registers: callers sit behind a barrier, where re-reading them costs an L2 round trip
// have_sums: the caller accumulated this thread's partial sums (ps, pq) over its elements i = tid, tid + blockDim, ... in that order
// while it produced them (the EMA loop): the same numbers as the loop below, without the pass over shared memory
__device__ void finish_template(const Ctx& c, int track, TrackState& t, const float* s_t, double* red, int tw, int th, bool have_sums, double ps, double pq)
{
    const int n = tw * th, tid = threadIdx.x;
    double s = ps, q = pq;
    if (!have_sums) {
        s = 0.0; q = 0.0;
        for (int i = tid; i < n; i += blockDim.x) {
            const double v = (double)s_t[i];
            s += v;
            q += v * v;
        }
    }
    block_sum2(s, q, red);
    const double scale = 1.0 / (double)n;
    const double mean = __dmul_rn(s, scale);
    double var = __dsub_rn(__dmul_rn(q, scale), __dmul_rn(mean, mean));
    if (var < 0.0) var = 0.0;
    const double sdv = sqrt(var);
    const double norm2 = __dmul_rn(sdv, sdv);
    const int tpad = (tw + 7) & ~7;
    if (tid == 0) {
        t.mean = mean;
        t.flat = norm2 < DBL_EPSILON;
        t.templ_norm = sqrt(norm2) / sqrt(scale);
        if (c.formula) {
            // baseline_kernel.cu:331-332 templStd = (float)(stddev + 1e-6f); :49 1 / (templStd + 1e-6f); :62 ... / N
            t.flat = 2;
            t.templ_norm = (double)((float)(sdv + (double)1e-6f) + 1e-6f) * (double)n;
        }
        t.tp = tpad;
    }
    if (c.tdig) write_digits(c, track, t, s_t, mean, tw, th);
    // centred template, CHUNK-MAJOR: tc[(x / 8) * th * 8 + y * 8 + (x % 8)], columns >= tw are zero.
    // Element i = y * tpad + x per thread, consecutive lanes = consecutive x: conflict-free shared-memory reads (a thread per
    // 8-float chunk row read with a stride of tw floats between lanes: 32-way bank conflicts, 2.2 of the update's 6.7 us on
    // a 64 x 64 template) and 32-byte store segments.  (y, x) advance incrementally: one division per thread.
    float* tc = c.templc + (size_t)track * c.mth * c.mtp;
    const int stride = blockDim.x, dq = stride / tpad, dr = stride - dq * tpad;
    int y = tid / tpad, x = tid - y * tpad;
    for (int i = tid; i < th * tpad; i += stride) {
        const float v = x < tw ? (float)((double)s_t[y * tw + x] - mean) : 0.f;
        tc[(size_t)(x >> 3) * th * 8 + y * 8 + (x & 7)] = v;
        x += dr; y += dq;
        if (x >= tpad) { x -= tpad; y += 1; }
    }
}

__device__ void track_update(const Ctx& c, int track, unsigned long long step, bool stepped, float* s_t, double* red)
{
    TrackState& t = c.tracks[track];
    pvt_result* res = c.results + (step % kRing) * c.max_tracks + track;
    // tracker_ghc semantics (Ctx.lost_mode): a step is a local pass followed by a whole-frame pass; a track is reported
    // by the pass that owns it (read before this call changes the track's mode)
    const bool owned = track_owned(c, t, step);
    // Everything this function needs from global memory that does not depend on the peak goes out HERE, before the first
    // branch (loads behind `if (stepped)` are a further dependent L2 round trip on the step's critical path): the parameters,
    // the peak, the track's fields -- and the first batch of the OLD template, whose addresses are known (used only if the EMA runs).
    const DevParams P = *c.params;
    const int tw = t.w, th = t.h, tstream = t.stream, ox = t.x, oy = t.y;
    float* const tp_ = c.templ + (size_t)track * c.mth * c.mtw;
    float tv0[16];
    {
        const int n0 = tw * th, stride0 = blockDim.x;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int i = threadIdx.x + k * stride0;
            tv0[k] = (stepped && i < n0) ? tp_[i] : 0.f;
        }
    }
    // k_update launched behind the search with a programmatic dependency (k_ncc_local shape): everything above ran while the
    // search was still computing; the peak and the window are the search's results.  (A no-op in every other launch.)
    pdl_wait();
    const unsigned long long key = *((volatile unsigned long long*)&t.peak);
    const int ww = t.win[2], wh = t.win[3], w0x = t.win[0], w0y = t.win[1];
    if (stepped) {
        const float val = unord_f32((unsigned int)(key >> 32));
        const unsigned int idx = 0xffffffffu - (unsigned int)(key & 0xffffffffull);
        const int bx = w0x + (int)(idx % (unsigned int)ww), by = w0y + (int)(idx / (unsigned int)ww);
        const double best = (double)val;
        const bool moved = best >= P.min_conf;
        const bool updated = moved && best >= P.strong_conf;
        const int nx = moved ? bx : ox, ny = moved ? by : oy;
        __syncthreads();  // everyone has read t.peak / t.x / t.y
        if (updated) {
            const int n = tw * th;
            const double alpha = 1.0 - P.lr, beta = P.lr;
            const float* g = c.gray + (size_t)tstream * c.plane + (size_t)ny * c.pitch + nx;
            // one batch of 32 loads per thread covers a 64 x 64 template with 256 threads: one L2 round trip, then the EMA.
            // (row, column) of pixel i advance incrementally: no division per element
            const int stride = blockDim.x, dq = stride / tw, dr = stride - dq * tw;
            int r = threadIdx.x / tw, col = threadIdx.x - r * tw;
            double es = 0.0, eq = 0.0;
            for (int i0 = threadIdx.x; i0 < n; i0 += 16 * stride) {
                float pv[16], tv[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int i = i0 + k * stride;
                    pv[k] = i < n ? g[(size_t)r * c.pitch + col] : 0.f;
                    tv[k] = i0 == (int)threadIdx.x ? tv0[k] : (i < n ? tp_[i] : 0.f);   // first batch: prefetched above
                    col += dr; r += dq;
                    if (col >= tw) { col -= tw; r += 1; }
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int i = i0 + k * stride;
                    if (i < n) {
                        const double pb = __dmul_rn((double)pv[k], beta);
                        const float v = (float)fma((double)tv[k], alpha, pb);
                        tp_[i] = v;
                        s_t[i] = v;
                        es += (double)v;                         // the new template's sums, in finish_template's own order
                        eq += (double)v * (double)v;
                    }
                }
            }
            // (no barrier here: finish_template's block reduction has two before anybody reads another thread's s_t)
            finish_template(c, track, t, s_t, red, tw, th, true, es, eq);
        }
        if (threadIdx.x == 0) {
            t.x = nx; t.y = ny; t.peak = 0ull;
            t.arrive = 0u; t.stats_done = 0u;                   // k_winstats may count (StatCfg.signal) although this step's kernels did not wait
            if (c.lost_mode) {
                // tracker_ghc/src/main.cpp:213-239: found -> reset the counter and go back to the local search (the box
                // comes from a map position, so it is never outside the frame); else count the frame as lost.
                // :183-185: from LOST_FRAME_THRESHOLD consecutive lost frames on, the NEXT frames search the whole map
                // with NCC_GLOBAL_CONFIDENCE as the acceptance threshold (P.min_conf of the global pass)
                if (moved) { t.lost_count = 0; t.use_global = 0; }
                else t.lost_count += 1;
                if (!t.use_global && t.lost_count >= P.lost_threshold) { t.use_global = 1; t.global_since = step + 1ull; }
            }
            atomicAdd(c.macs, (unsigned long long)ww * wh * tw * th);
            atomicAdd(c.macs_grid, (unsigned long long)min(ww, c.gridW) * min(wh, c.gridH) * tw * th);
            res->x = nx; res->y = ny; res->w = tw; res->h = th;
            res->conf = val; res->moved = moved; res->updated = updated; res->searched = c.global_pass ? 2 : 1; res->valid = 1;
            res->track = track; res->step = (int32_t)step;
        }
    } else if (owned && threadIdx.x == 0) {
        res->x = t.x; res->y = t.y; res->w = t.w; res->h = t.h;
        res->conf = __int_as_float(0x7fc00000);
        res->moved = 0; res->updated = 0; res->searched = 0; res->valid = (uint8_t)(t.active != 0);
        res->track = track; res->step = (int32_t)step;
    }
    // last track done -> advance the time step (every kernel of this step has read *c.step already).
    // With two passes per step (lost-object mode) k_step_advance does it after both.
    if (c.lost_mode) return;
    // No fences: the counter is only read by the kernels of the NEXT step, which start after this kernel has completed
    // (nobody in this launch reads it again), and a context with a single track needs no arrival count either.
    if (threadIdx.x == 0) {
        if (c.max_tracks == 1) {
            *c.step = step + 1ull;
        } else {
            const unsigned int n = atomicAdd(c.ticket, 1u);
            if (n == (unsigned int)c.max_tracks - 1u) {
                *c.ticket = 0u;
                *c.step = step + 1ull;
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_update(Ctx c)
{
    extern __shared__ float sm_f[];
    __shared__ double red[64];
    const int track = blockIdx.x;
    unsigned long long step;
    const bool stepped = track_stepped_ld(c, c.tracks[track], step);
    trace_begin(c, step, TR_UPDATE);
    track_update(c, track, step, stepped, sm_f, red);
    trace_end(c, step, TR_UPDATE);
}

// Lost-object mode, after the local pass: which streams carry a track that is searched over the whole frame this step --
// and whether there is any.  The whole-frame pass is the body of a CUDA-graph CONDITIONAL node: this kernel sets the
// condition on the device (no host round trip), so a step in which no track is lost pays for two tiny kernels only.
__global__ void k_global_mark(Ctx c, cudaGraphConditionalHandle cond, int use_cond)
{
    __shared__ int any;
    const unsigned long long step = *c.step;
    if (threadIdx.x == 0) any = 0;
    for (int s = threadIdx.x; s < c.max_streams; s += blockDim.x) c.stream_need[s] = 0;
    __syncthreads();
    for (int track = threadIdx.x; track < c.max_tracks; track += blockDim.x) {
        const TrackState& t = c.tracks[track];
        if (track_global(t, step)) { c.stream_need[t.stream] = 1; any = 1; }
    }
    __syncthreads();
    if (threadIdx.x == 0 && use_cond) cudaGraphSetConditional(cond, any ? 1u : 0u);
}

// Lost-object mode, last kernel of a step: both passes have reported their tracks; advance the device time step.
__global__ void k_step_advance(Ctx c)
{
    if (threadIdx.x == 0) { *c.ticket = 0u; *c.step = *c.step + 1ull; }
}

// Start of a frame sequence (pvt_submit_sequence) or return to per-step submission: set the sequence descriptor, passed
// by value, and (invalidate) forget what k_prefetch_roi staged before -- the caller may have refilled its buffers.
__global__ void k_seq_begin(Ctx c, SeqDesc q, int invalidate)
{
    if (threadIdx.x == 0) *c.seq = q;
    if (invalidate && c.stage_hdr)
        for (int t = threadIdx.x; t < c.max_tracks; t += blockDim.x) { c.stage_hdr[t].step = ~0ull; c.stage_hdr[t].cur_step = ~0ull; }
}

// hold step (batch mode, main.cpp:118-123): no NCC, no update; emit the stale box and advance
__global__ void k_hold(Ctx c)
{
    const unsigned long long step = *c.step;
    for (int track = threadIdx.x; track < c.max_tracks; track += blockDim.x) {
        const TrackState& t = c.tracks[track];
        pvt_result* res = c.results + (step % kRing) * c.max_tracks + track;
        res->x = t.x; res->y = t.y; res->w = t.w; res->h = t.h;
        res->conf = __int_as_float(0x7fc00000);
        res->moved = 0; res->updated = 0; res->searched = 0;
        res->valid = (uint8_t)(t.active != 0);
        res->track = track; res->step = (int32_t)step;
    }
    __syncthreads();
    if (threadIdx.x == 0) *c.step = step + 1ull;
}

// =============================================================================================
// (6) overlay: the sink side of the loop, tracker/src/main.cpp:166  cv::rectangle(frame, bbox, {0,255,0}, 2)
//     cv::rectangle(Rect) with thickness 2 on an 8-bit image paints the 3-pixel band around the lines (x0,y0)-(x1,y1),
//     x1 = x + w - 1, y1 = y + h - 1, minus the four outer corner pixels (round line caps): the outer box [x0-1, x1+1] x
//     [y0-1, y1+1] without its hole [x0+2, x1-2] x [y0+2, y1-2] and its four corners, clipped to the image
//     (pinned to cv2 4.13.0 by tests/test_overlay.py).  One thread per pixel of the outer box; grid.y = box.
// =============================================================================================
__global__ void __launch_bounds__(256) k_overlay(unsigned char* img, size_t step, int W, int H, const int* boxes, int nbox, int b, int g, int r)
{
    const int k = blockIdx.y;
    if (k >= nbox) return;
    const int x = boxes[4 * k], y = boxes[4 * k + 1], w = boxes[4 * k + 2], h = boxes[4 * k + 3];
    if (w <= 0 || h <= 0) return;
    const int x0 = x, y0 = y, x1 = x + w - 1, y1 = y + h - 1;
    const int ow = x1 - x0 + 3, oh = y1 - y0 + 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ow * oh; i += gridDim.x * blockDim.x) {
        const int px = x0 - 1 + i % ow, py = y0 - 1 + i / ow;
        if (px < 0 || py < 0 || px >= W || py >= H) continue;
        const bool hole = px >= x0 + 2 && px <= x1 - 2 && py >= y0 + 2 && py <= y1 - 2;
        const bool corner = (px == x0 - 1 || px == x1 + 1) && (py == y0 - 1 || py == y1 + 1);
        if (hole || corner) continue;
        unsigned char* q = img + (size_t)py * step + 3 * px;
        q[0] = (unsigned char)b; q[1] = (unsigned char)g; q[2] = (unsigned char)r;
    }
}

}  // namespace pvt

#include "ncc_tc.cuh"
