// libpvt.so -- C-ABI entry points (include/pvt.h) over the sm_100a kernels in kernels.cuh.
// Host code is C++; no torch, no CPU compute path: without a CUDA device every call fails loudly.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "kernels.cuh"

using namespace pvt;

namespace {

thread_local std::string g_err = "";

int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return fail(PVT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + \
                                          ":" + std::to_string(__LINE__) + ")");                            \
    } while (0)

constexpr int kMultiStep = 4;   // steps per launch of graph_multi: a graph-to-graph boundary costs ~4 us on the device, an edge inside a graph ~1.5 us
constexpr int kMaxGraphSteps = 64; // largest exact-length step graph (latency shape): the results ring holds 64 steps
constexpr int kLongStep = 16;   // ... and of graph_long (latency shape only: there 4 us per four 27 us steps is still 4 %)
constexpr int kStageDepth = 3;  // host frames in flight per stream (H2D overlaps the previous step's kernels)
constexpr size_t kSmemBudget = 227u * 1024u;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct EventPair {
    cudaEvent_t a, b;
    int cls;
};

// Everything the kernels of one pass over the tracks need.  A time step is ONE pass (the local search window of
// tracker/src/main.cpp) or, in lost-object mode (tracker_ghc/src/main.cpp), that pass followed by a whole-frame pass
// over the tracks that are currently lost: same kernels, window == the whole NCC map, its own plan and scratch.
struct Pass {
    Ctx d{};                   // by-value kernel argument
    TileCfg tile{};
    CUtensorMap tmap{};
    size_t ncc_smem = 0;
    int rowsum_warps = 8, rowsum_pw = 0;
    int colprefix_chunks = 8;  // row chunks per 32-column strip in k_colprefix (blockDim.y)
    StatCfg stat{};            // k_winstats (one statistics kernel); NX == 0: the two-kernel statistics (k_colprefix + k_rowsum)
    bool fused = false;        // K-split shape: k_step_fused (search + second stage + update in one launch) instead of k_ncc_search -> k_ncc_finalize
    size_t fused_smem = 0;
    LocalCfg local{};          // latency shape: k_ncc_local (K-split and window statistics inside the CTA); TR == 0: not used
    size_t local_smem = 0;
    FringeCfg fringe{};        // CTAs of k_ncc_fringe per track (candidates outside the thread-tile grid); all 0 = none
    size_t fringe_smem = 0;
    bool roi_ingest = false;   // k_ingest_roi instead of k_ingest
    bool pdl = true;           // k_ncc_fringe behind the search with a programmatic dependency (plain stream order otherwise)
    bool prefetch = false;     // k_prefetch_roi beside the step (graphs for pinned host rings)
    TcCfg tc{};                // PVT_KERNEL_TC: geometry of k_ncc_tc for this pass's windows (XW == 0: the pass keeps the FP32 search)
    size_t tc_smem = 0;
    CUtensorMap tmap8{};
};

}  // namespace

struct pvt_ctx {
    pvt_params params{};
    pvt_config cfg{};
    Ctx d{};  // by-value kernel argument
    TileCfg tile{};
    CUtensorMap tmap{};
    size_t ncc_smem = 0;
    int rowsum_warps = 8, rowsum_pw = 0;
    int kps = 5;               // kernels per searched time step (for pvt_launch_count)
    bool lost_mode = false;    // params.lost_frame_threshold > 0 at creation: every searched step also runs the whole-frame pass
    Pass G{};                  // the whole-frame pass (lost-object re-acquisition)
    int kps_global = 0;
    cudaGraphExec_t graph_global = nullptr;
    FringeCfg fringe{};        // CTAs of k_ncc_fringe per track (candidates outside the thread-tile grid); all 0 = none
    size_t fringe_smem = 0;
    bool roi_ingest = false;   // k_ingest_roi instead of k_ingest (pvt_params.ingest)
    int colprefix_chunks = 8;  // row chunks per 32-column strip in k_colprefix (blockDim.y)
    StatCfg stat{};            // k_winstats geometry (NX == 0: two-kernel statistics)
    bool fused = false;        // k_step_fused (see Pass)
    size_t fused_smem = 0;
    LocalCfg local{};          // k_ncc_local (see Pass)
    size_t local_smem = 0;
    unsigned int* h_fault = nullptr;   // mapped pinned word behind Ctx.fault
    size_t templ_smem = 0;     // th*tw floats of dynamic shared memory for the update / init kernels
    cudaStream_t compute = nullptr, copy = nullptr, aux = nullptr, aux2 = nullptr, aux3 = nullptr;   // aux*: further branches inside the captured graph
    cudaEvent_t ev_join3 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr;
    cudaGraphExec_t graph = nullptr, graph_hold = nullptr, graph_prof = nullptr;
    cudaGraphExec_t graph_multi = nullptr;   // kMultiStep consecutive time steps in one launch (resident frame rings)
    cudaGraphExec_t graph_pf = nullptr, graph_multi_pf = nullptr;   // the same with k_prefetch_roi (pinned host rings)
    cudaGraphExec_t graph_long = nullptr, graph_long_pf = nullptr;  // kLongStep steps per launch (K-split shape)
    // latency shape: graphs of exactly n consecutive steps, built on first use ([0]: device-resident ring, [1]: pinned host ring with
    // the prefetch branch).  A 20-step sequence is then ONE launch: the host is out of the loop for the whole sequence (two launches
    // left a window in which a busy host -- eight ranks, their NCCL threads -- stalls the device between them).
    cudaGraphExec_t graph_n[2][kMaxGraphSteps + 1] = {};
    cudaEvent_t pev[5][2]{};           // profiling: event-record NODES inside graph_prof, one pair per kernel class (+ k_ncc_search alone)
    bool graph_valid = false;
    std::vector<void*> allocs;
    FrameDesc* h_table = nullptr;  // pinned, [kRing][max_streams]
    cudaEvent_t table_ev[kRing]{};
    bool table_ev_used[kRing]{};
    std::vector<void*> stage;  // [max_streams * kStageDepth], lazily allocated
    size_t stage_bytes = 0;
    cudaEvent_t ev_done[kStageDepth]{}, ev_copied[kStageDepth]{};
    cudaEvent_t rb_ready[2]{}, rb_done[2]{};   // pipelined result read-back (pvt_submit_sequence)
    bool ev_done_used[kStageDepth]{};
    unsigned long long submitted = 0;  // time steps enqueued
    int hold_pending = 0;              // batch mode: frames since the last searched one
    bool seq_default = true;           // device SeqDesc is {0, kRing, 0}
    int64_t launches = 0;
    bool profiling = false;
    std::vector<EventPair> ev_pool;
    size_t ev_next = 0;
    pvt_profile prof{};
    unsigned long long* d_macs = nullptr;
    unsigned long long* d_trace = nullptr;
    cudaEvent_t timer_a = nullptr, timer_b = nullptr;
    pvt_result* h_results = nullptr;  // pinned, [kRing][max_tracks]
    std::vector<int> track_stream;    // host mirror: -1 = inactive
    std::vector<char> plane_full;     // per stream: the gray plane holds ONE complete frame (false once a ROI-ingest step refreshed only tiles)
    int prefetch_delay_ns = 0;        // test hook PVT_DEBUG_PREFETCH_DELAY_US: k_prefetch_roi starts late
    // pvt_submit_sequence: what the device-side frame table / SeqDesc currently hold, so that an unchanged ring costs no upload
    std::vector<FrameDesc> seq_rows;  // the rows uploaded by the last resident sequence ([ring_len][max_streams])
    int seq_ring_len = 0;
    bool seq_rows_valid = false;      // d.table rows [0, seq_ring_len) == seq_rows (any per-step submit invalidates it)
    std::vector<FrameDesc> seq_scratch;
    TcCfg tc{};                       // PVT_KERNEL_TC: geometry of k_ncc_tc, its shared memory and the 4-D tensor map of the u8 gray plane
    size_t tc_smem = 0;
    CUtensorMap tmap8{};
    bool tc_ready = false;            // created with PVT_KERNEL_TC (gray8 / digits / window sums allocated)
    bool host_ptr_is_dev = false;     // cudaDevAttrCanUseHostPointerForRegisteredMem: a pinned host pointer is its own device alias
};

namespace {

// PVT_KERNEL_TC_GLOBAL: the planner's FP32 kernels for the local windows, the tensor cores for the whole-frame pass only
inline int local_kernel(const pvt_ctx* c) { return c->params.kernel == PVT_KERNEL_TC_GLOBAL ? (int)PVT_KERNEL_AUTO : (int)c->params.kernel; }
inline int pass_kernel(const pvt_ctx* c, const Ctx& d)
{
    if (c->params.kernel != PVT_KERNEL_TC_GLOBAL) return (int)c->params.kernel;
    return d.global_pass ? (int)PVT_KERNEL_TC : (int)PVT_KERNEL_AUTO;
}
inline bool wants_tc(int kernel) { return kernel == PVT_KERNEL_TC || kernel == PVT_KERNEL_TC_GLOBAL; }

template <typename T>
int dev_alloc(pvt_ctx* c, T** p, size_t n, bool zero = true)
{
    void* q = nullptr;
    CK(cudaMalloc(&q, std::max<size_t>(n * sizeof(T), 256)));
    if (zero) CK(cudaMemset(q, 0, std::max<size_t>(n * sizeof(T), 256)));
    c->allocs.push_back(q);
    *p = (T*)q;
    return PVT_OK;
}

// Plan k_ncc_search for a context: row-band height (GB groups of kCY rows), and the K-split (pj parts along the
// template's 8-column chunks x pd parts along its rows).  K-split serves two purposes: (1) large templates / windows
// (4K, 128x128, R=160) only fit the 256-row TMA box and the shared memory when a CTA sees part of the template;
// (2) with few tracks (a single 1080p stream is 106 MMAC = 2.9 us of FP32 peak) the work must be cut finely enough
// to occupy all SMs.  Cost model, calibrated on B200 with the device timeline (tools/timeline.py, profiles/):
// CTAs run in rounds of 2 per SM; a round costs ~3 us of fixed latency (launch, TMA tile, partial-sum stores) plus the
// per-thread FMA count over ~870 FMA/us (two warps per sub-partition) or ~1250 FMA/us (one); K-split adds the
// second-stage reduction.  Smallest estimate wins; ties go to fewer parts, then taller bands.
// k_ncc_fringe geometry: tiles (of 8 candidates) per CTA and the largest K-split row-part count whose chain results
// still fit its shared memory: strip | template | results [pd][chunks][8 * tpc].  Returns false when nothing fits
// (then the thread-tile grid keeps the remainder and there is no fringe kernel).
// the two-kernel statistics (k_colprefix + k_rowsum and their FP64 prefix arrays) are only kept for templates too wide for
// k_winstats' one-column-per-thread tile, and behind PVT_STATS_LEGACY=1 for the before/after comparison in the tests
bool stats_legacy(int mtw)
{
    const char* legacy = getenv("PVT_STATS_LEGACY");
    return kStatThreads - (mtw - 1) < 32 || (legacy && *legacy == '1');
}

int fringe_strip_floats(int tpc, int mtp, int mth)
{
    const int col = (tpc * 8 + mth - 1) * fringe_pitch(mtp), row = mth * fringe_pitch(tpc * 8 + mtp);
    return std::max(col, row);
}
bool fringe_plan(int mtp, int mth, int* tpc_out, int* pd_cap_out)
{
    const int nch = mtp / 8;
    if (nch > kFringeThreads) return false;
    const long long budget = 216 * 1024 / 4;
    const int want[4] = {16, 8, 4, 1};   // prefer many tiles per CTA
    for (int wi = 0; wi < 4; ++wi)
        for (int tpc = std::min(kFringeThreads / nch, 24); tpc >= 1; --tpc) {
            const long long left = budget - fringe_strip_floats(tpc, mtp, mth) - (long long)nch * (mth * 8 + 4);
            // chain results: [chunks][8 * tpc] (a CTA computes one K-split row part)
            if (left >= (long long)nch * tpc * 8 && tpc >= std::min(want[wi], 24)) {
                *tpc_out = tpc;
                *pd_cap_out = 32;
                return true;
            }
        }
    return false;
}

bool choose_plan(int sm_count, int n_tracks, int mtw, int mtp, int mth, int Wmax, int Hmax, TileCfg* out, size_t* smem_out)
{
    const int CY = kCY;
    const int nch = mtp / 8;
    // a remainder of exactly one row / column stays out of the thread-tile grid (k_ncc_fringe takes it): see TileCfg
    int pd_cap = 0, tpc = 0;
    (void)mtw;
    const bool fringe_ok = fringe_plan(mtp, mth, &tpc, &pd_cap);
    // ... in UNSPLIT plans (the throughput shape, where every SM slot counts).  K-split plans (few tracks: the GPU is not
    // full, what counts is the critical path of the step) keep the remainder in the grid as masked lanes: a sixth CTA per
    // part runs beside the other five on an idle SM, and the step has one kernel and one graph branch less.
    const int Gf = (fringe_ok && Hmax > CY && Hmax % CY == 1) ? Hmax / CY : (Hmax + CY - 1) / CY;
    const int Cf = (fringe_ok && Wmax > 8 && Wmax % 8 == 1) ? Wmax / 8 : (Wmax + 7) / 8;
    const int Gc = (Hmax + CY - 1) / CY, Cc = (Wmax + 7) / 8;
    auto span_of = [](int C, int GB) { return std::min(C, kTilesPerCta % GB == 0 ? kTilesPerCta / GB : (kTilesPerCta - 1) / GB + 2); };
    const long long slots = (long long)sm_count * 2;
    double best = 1e300;
    // does the unsplit plan exist and fill at least one whole round?  then never K-split globally: the partial round at
    // the end is handled by tail splitting (pvt_create), which has none of the K-split's traffic
    bool unsplit_fills = false;
    {
        const int G = Gf, C = Cf;
        const int boxH = G * CY + mth - 1, span = span_of(C, G), boxW = 8 * span + mtp + 4;
        const size_t smem = (size_t)boxW * boxH * 4 + (size_t)4 * mth * 32 + 128;
        const long long ctas = (long long)n_tracks * ((G * C + kTilesPerCta - 1) / kTilesPerCta);
        unsplit_fills = boxH <= 256 && boxW <= 256 && 2 * (smem + 1024) <= 228u * 1024u && ctas >= slots;
    }
    for (int pj = 1; pj <= (unsplit_fills ? 1 : nch); ++pj)
        for (int pd = 1; pd <= (unsplit_fills ? 1 : std::min(mth, 32)); ++pd) {
            const int nchp = (nch + pj - 1) / pj, ndp = (mth + pd - 1) / pd;
            if ((pj > 1 && (pj - 1) * nchp >= nch) || (pd > 1 && (pd - 1) * ndp >= mth)) continue;  // a part would be empty
            const int G = pj * pd > 1 ? Gc : Gf, C = pj * pd > 1 ? Cc : Cf;
            for (int GB = G; GB >= 1; --GB) {
                const int boxH = GB * CY + ndp - 1;
                if (boxH > 256) continue;
                const int span = span_of(C, GB);
                const int boxW = 8 * span + 8 * nchp + 4;
                if (boxW > 256) continue;
                const size_t smem = (size_t)boxW * boxH * 4 + (size_t)4 * mth * 32 + 128;
                if (2 * (smem + 1024) > 228u * 1024u) continue;   // two CTAs per SM
                const int bands = (G + GB - 1) / GB, ctas_band = (GB * C + kTilesPerCta - 1) / kTilesPerCta;
                const long long ctas = (long long)n_tracks * bands * ctas_band * pj * pd;
                // FMAs per thread, plus the CY-1 window rows every template chunk preloads before its first row (they weigh
                // in when a K-split part holds only a few template rows)
                const double work = 8.0 * CY * 8.0 * nchp * (ndp + (pj * pd > 1 ? CY - 1 : 0));
                double rounds = (double)((ctas + slots - 1) / slots);
                const int parts = pj * pd;
                if (parts == 1 && ctas > slots) {
                    // the partial last round of an unsplit plan is cut along the template chunks (tail splitting, build_plan)
                    const long long rem = ctas % slots;
                    const int ps = rem > 0 ? (int)std::min<long long>(nch, slots / rem) : 0;
                    if (rem > 0 && rem * 5 <= slots * 4 && ps >= 2) rounds = (double)(ctas / slots) + 1.0 / ps + 0.03;
                }
                const double rate = ctas > sm_count ? 870.0 : 1250.0;
                // the second stage reads parts*4+8 bytes and the first writes parts*4 bytes per candidate and track
                const double split_us = parts > 1 ? 3.0 + 0.12 * parts + 2.5e-6 * (double)n_tracks * Wmax * Hmax * parts : 0.0;
                // K-split shape: the window statistics run beside the search (k_colprefix CTAs are 1024 threads wide and
                // cannot share an SM with two search CTAs).  A search grid that fills every SM slot starves them and the
                // statistics become the critical path (measured on C3: rowsum 58 us instead of 20): leave 8 SMs free.
                const double starve_us = (parts > 1 && ctas > slots - 16 && ctas <= slots) ? 0.5 * (work / rate) : 0.0;
                const double us = rounds * (3.0 + work / rate) + split_us + starve_us;
                const double key = us * (1.0 + 1e-4 * parts) - 1e-6 * GB;
                if (key < best) {
                    best = key;
                    *out = TileCfg{G, C, GB, bands, ctas_band, span, boxW, boxH, pj, pd, bands * ctas_band, 0, 0, 0};
                    *smem_out = smem;
                }
            }
        }
    return best < 1e300;
}

int encode_tmap(const Ctx& d, const TileCfg& tile, CUtensorMap* out)
{
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(PVT_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
    cuuint64_t dims[3] = {(cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.max_streams};
    cuuint64_t strides[2] = {(cuuint64_t)d.pitch * 4, (cuuint64_t)d.plane * 4};
    cuuint32_t box[3] = {(cuuint32_t)tile.boxW, (cuuint32_t)tile.boxH, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)fn)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d.gray, dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PVT_ERR_CUDA, "cuTensorMapEncodeTiled failed, CUresult " + std::to_string((int)r));
    return PVT_OK;
}

// process-wide: the dynamic shared-memory limit of a kernel is only ever raised (several contexts may coexist)
int raise_smem(const void* fn, size_t bytes)
{
    static std::mutex mu;
    static std::map<const void*, size_t> cur;
    std::lock_guard<std::mutex> lk(mu);
    size_t& v = cur[fn];
    if (bytes > v) {
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        v = bytes;
    }
    return PVT_OK;
}

int encode_tmap(const Ctx& d, const TileCfg& tile, CUtensorMap* out);
int tc_geometry(pvt_ctx* c, const Ctx& d, int sm_count, TcCfg& g, size_t* smem, CUtensorMap* tmap8);

// kernels one pass launches per searched step: ingest, 2 statistics, search, [fringe], [tail reduction] + update | finalize
int pass_kernels(const TileCfg& t, const FringeCfg& f, const StatCfg& st, bool fused = false, bool local = false, int extra_local = 0)
{
    if (local) return 2 + extra_local;   // ingest, [k_winstats,] k_ncc_local [, k_update]: extra_local counts the optional ones
    if (fused) return 3;   // ingest, k_winstats, k_step_fused
    return (st.NX > 0 ? 3 : 4) + ((t.pj * t.pd > 1) ? 1 : (t.tail_ps > 1 ? 2 : 1)) + (f.colg + f.rowg > 0 ? 1 : 0);
}

// ... of the local pass under the context's kernel choice (k_ncc_direct: ingest, 2 statistics, search, update)
int kernels_per_step(const pvt_ctx* c)
{
    // k_ncc_direct / k_ncc_tc: ingest, statistics (1 or 2 kernels), search, update
    return local_kernel(c) != PVT_KERNEL_AUTO ? (c->stat.NX > 0 ? 4 : 5) : pass_kernels(c->tile, c->fringe, c->stat, c->fused, c->local.TR > 0, c->local.gstats + (c->local.update ? 0 : 1));
}

// item grid + tail splitting (see TileCfg)
void plan_items(TileCfg& g, int max_tracks, int mtp, int sm_count)
{
    g.cpt = g.bands * g.ctas_band;
    const long long items = (long long)max_tracks * g.cpt, slots = (long long)sm_count * 2;
    g.n_full = (int)items; g.n_tail = 0; g.tail_ps = 0;
    const char* no_tail = getenv("PVT_NO_TAIL_SPLIT");
    if (g.pj * g.pd == 1 && items > slots && !(no_tail && *no_tail == '1')) {
        const long long rem = items % slots;
        const int nch = mtp / 8;
        const int ps = rem > 0 ? (int)std::min<long long>(nch, slots / rem) : 0;
        if (rem > 0 && rem * 5 <= slots * 4 && ps >= 2) {
            g.n_tail = (int)rem; g.n_full = (int)(items - rem); g.tail_ps = ps;
        }
    }
}

// Plan one pass: p.d holds the geometry (W, H, templates, Wmax, Hmax, VW) and the shared pointers on entry; on return the
// plan (tile grid, K-split / tail split, fringe, shared-memory sizes, TMA descriptor) and the pass's own scratch.
int build_plan(pvt_ctx* c, Pass& p, int sm_count, int ingest, bool allow_env)
{
    Ctx& d = p.d;
    if (!choose_plan(sm_count, d.max_tracks, d.mtw, d.mtp, d.mth, d.Wmax, d.Hmax, &p.tile, &p.ncc_smem)) { return fail(PVT_ERR_UNSUPPORTED, "no k_ncc_search plan fits this template / window size");
    }
    if (const char* e = allow_env ? getenv("PVT_PLAN") : getenv("PVT_PLAN_GLOBAL")) {  // experiments / tests: "GB,pj,pd[,f]" overrides the planner
        int GB = 0, pj = 0, pd = 0, f = -1;                              // f = 1 / 0: remainder row / column out of / in the grid
        if (sscanf(e, "%d,%d,%d,%d", &GB, &pj, &pd, &f) >= 3 && GB > 0 && pj > 0 && pd > 0 && pd <= 32) {
            TileCfg& g = p.tile;
            if (f >= 0) {
                g.G = (f && d.Hmax > kCY && d.Hmax % kCY == 1) ? d.Hmax / kCY : (d.Hmax + kCY - 1) / kCY;
                g.C = (f && d.Wmax > 8 && d.Wmax % 8 == 1) ? d.Wmax / 8 : (d.Wmax + 7) / 8;
            }
            const int nch = d.mtp / 8, nchp = (nch + pj - 1) / pj, ndp = (d.mth + pd - 1) / pd;
            g.GB = std::min(GB, g.G); g.pj = pj; g.pd = pd;
            g.bands = (g.G + g.GB - 1) / g.GB; g.ctas_band = (g.GB * g.C + kTilesPerCta - 1) / kTilesPerCta;
            g.span = std::min(g.C, kTilesPerCta % g.GB == 0 ? kTilesPerCta / g.GB : (kTilesPerCta - 1) / g.GB + 2);
            g.boxW = 8 * g.span + 8 * nchp + 4; g.boxH = g.GB * kCY + ndp - 1;
            p.ncc_smem = (size_t)g.boxW * g.boxH * 4 + (size_t)4 * d.mth * 32 + 128;
            if (g.boxW > 256 || g.boxH > 256 || p.ncc_smem + 1024 > kSmemBudget) { return fail(PVT_ERR_INVALID, "PVT_PLAN does not fit the TMA box / shared memory");
            }
        }
    }
    plan_items(p.tile, d.max_tracks, d.mtp, sm_count);
    {
        const double tiles = (double)d.max_tracks * (d.Wmax + d.mtw) * (d.Hmax + d.mth), frames_px = (double)d.max_streams * d.W * d.H;
        p.roi_ingest = !d.global_pass && (ingest == PVT_INGEST_ROI || (ingest == PVT_INGEST_AUTO && tiles <= 0.5 * frames_px));
    }
    d.gridW = 8 * p.tile.C;
    d.gridH = kCY * p.tile.G;
    {
        int tpc = 1, pd_cap = 0;
        const bool ok = fringe_plan(d.mtp, d.mth, &tpc, &pd_cap);
        const int col_tiles = (d.Hmax + 7) / 8, row_tiles = (std::min(d.Wmax, 8 * p.tile.C) + 7) / 8;
        tpc = std::max(1, std::min(tpc, std::max(col_tiles, row_tiles)));
        p.fringe.tpc = tpc;
        p.fringe.colg = d.Wmax > 8 * p.tile.C ? (col_tiles + tpc - 1) / tpc : 0;
        p.fringe.rowg = d.Hmax > kCY * p.tile.G ? (row_tiles + tpc - 1) / tpc : 0;
        p.fringe.strip_floats = fringe_strip_floats(tpc, d.mtp, d.mth);
        p.fringe.threads = std::min(kFringeThreads, (tpc * (d.mtp / 8) + 31) & ~31);
        if (p.fringe.colg + p.fringe.rowg > 0 && (!ok || p.tile.pd > pd_cap)) { return fail(PVT_ERR_UNSUPPORTED, "internal: k_ncc_fringe plan does not fit shared memory");
        }
    }
    p.fringe.defer = p.tile.pj * p.tile.pd > 1 ? 1 : 0;
    if (p.fringe.defer && p.fringe.colg + p.fringe.rowg > 0)
        { int r_ = dev_alloc(c, &d.fringe_acc, (size_t)d.max_tracks * p.tile.pj * p.tile.pd * (d.Hmax + d.Wmax)); if (r_) return r_; }
    p.fringe_smem = ((size_t)p.fringe.strip_floats + (size_t)(d.mtp / 8) * (d.mth * 8 + 4) + (size_t)(d.mtp / 8) * p.fringe.tpc * 8) * sizeof(float);
    if (p.fringe.colg + p.fringe.rowg > 0) {
        if (p.fringe_smem > 220u * 1024u) { return fail(PVT_ERR_UNSUPPORTED, "internal: k_ncc_fringe does not fit shared memory"); }
        { int r_ = raise_smem((const void*)k_ncc_fringe, p.fringe_smem); if (r_) return r_; }
    }
    if (getenv("PVT_DEBUG_PLAN"))
        fprintf(stderr, "[pvt] plan: tracks=%d G=%d C=%d GB=%d bands=%d ctas/band=%d span=%d box=%dx%d pj=%d pd=%d smem=%zu | items full=%d tail=%d x%d | fringe ctas %d+%d x %d thr, %d tiles/cta, smem %zu\n",
                d.max_tracks, p.tile.G, p.tile.C, p.tile.GB, p.tile.bands, p.tile.ctas_band, p.tile.span, p.tile.boxW, p.tile.boxH,
                p.tile.pj, p.tile.pd, p.ncc_smem, p.tile.n_full, p.tile.n_tail, p.tile.tail_ps, p.fringe.colg, p.fringe.rowg, p.fringe.threads, p.fringe.tpc, p.fringe_smem);
    if (p.tile.tail_ps > 1)           // tail items' partial cross terms: [tail part][128 tiles][8 * kCY]
        { int r_ = dev_alloc(c, &d.partial, (size_t)p.tile.n_tail * p.tile.tail_ps * kTilesPerCta * 8 * kCY, false); if (r_) return r_; }
    if (p.tile.pj * p.tile.pd > 1)   // tile-major partial cross terms: [parts][tracks][CTAs per track * 128 tiles][8 * kCY]
        { int r_ = dev_alloc(c, &d.partial, (size_t)p.tile.pj * p.tile.pd * d.max_tracks * p.tile.bands * p.tile.ctas_band * kTilesPerCta * 8 * kCY, false); if (r_) return r_; }
    { int r_ = raise_smem((const void*)k_ncc_search<kCY>, p.ncc_smem); if (r_) return r_; }
    p.rowsum_pw = (d.VW + 8) + ((d.VW + 8) >> 3) + 2;   // prefix row incl. one padding double per 8 (k_rowsum PH())
    p.rowsum_warps = (int)std::max<size_t>(1, std::min<size_t>(8, (200u * 1024u) / ((size_t)2 * p.rowsum_pw * sizeof(double))));
    { int r_ = raise_smem((const void*)k_rowsum, (size_t)p.rowsum_warps * 2 * p.rowsum_pw * sizeof(double)); if (r_) return r_; }
    // k_colprefix: ~28 rows per thread (8 chunks for a 224-row tracker tile, up to 32 for full-frame maps)
    p.colprefix_chunks = std::max(1, std::min(32, (d.Hmax + d.mth - 1 + 27) / 28));
    // k_winstats: a CTA of 256 threads owns one tile column per thread -> NX = 256 - (mtw - 1) candidates along x; NY candidate
    // rows per CTA: tall bands when there are CTAs to spare (every band re-reads mth - 1 rows), short ones when a few tracks
    // must spread over the GPU (the single-stream step waits for the slowest CTA)
    p.stat = StatCfg{};
    {
        const int NX = kStatThreads - (d.mtw - 1);
        if (!stats_legacy(d.mtw)) {
            const int xt = (d.Wmax + NX - 1) / NX;
            // bands: every band re-reads mth - 1 rows (the first column sums), so few tall bands do the least work -- but the CTAs
            // run in waves of 3 per SM and a single-stream step waits for the slowest CTA.  Cost model in row units: a band
            // costs mth row loads + ~3 units per candidate row (the horizontal scan and the normaliser) + a fixed part.
            const long long slots = 3LL * sm_count;
            int NY = d.Hmax;
            double best = 1e300;
            for (int nb = 1; nb <= std::max(1, (d.Hmax + 7) / 8); ++nb) {
                const int ny = (d.Hmax + nb - 1) / nb, nbe = (d.Hmax + ny - 1) / ny;
                const long long ctas = (long long)d.max_tracks * xt * nbe;
                const double cost = (double)((ctas + slots - 1) / slots) * (d.mth + 3.0 * ny + 20.0);
                if (cost < best) { best = cost; NY = ny; }
            }
            p.stat = StatCfg{NX, NY, xt, (d.Hmax + NY - 1) / NY};
        }
    }
    // k_step_fused: the K-split step in one launch.  Its CTAs wait for each other inside the kernel, so every search CTA of
    // the context must be resident at once, with room left for the statistics kernel beside them: at most one search CTA per
    // SM on average (two fit), no fringe kernel, no lost-object pass, and the update's template must fit the CTA's shared memory.
    p.fused = false;
    p.fused_smem = std::max(p.ncc_smem, (size_t)d.mth * d.mtw * sizeof(float));
    {
        // Measured on B200 (C2, round 2): 34.0 us per step against 26.3 us for k_ncc_search -> k_ncc_finalize -- the kernel boundary
        // it removes costs less than what it adds: a 128-thread update (10.2 us instead of 5.6 with 256 threads) and a second stage
        // in which only ~54 threads per CTA have work (5.4 us instead of 3.1).  Kept behind PVT_FUSED=1, with its parity tests.
        const char* nf = getenv("PVT_FUSED");
        const long long search_ctas = (long long)d.max_tracks * p.tile.cpt * p.tile.pj * p.tile.pd;
        if (allow_env && (nf && *nf == '1') && local_kernel(c) == PVT_KERNEL_AUTO && !c->lost_mode && p.tile.pj * p.tile.pd > 1 &&
            p.fringe.colg + p.fringe.rowg == 0 && p.stat.NX > 0 && search_ctas <= sm_count && p.fused_smem + 1024 <= kSmemBudget) {
            p.fused = true;
            p.stat.signal = 1;
            { int r_ = raise_smem((const void*)k_step_fused<kCY>, p.fused_smem); if (r_) return r_; }
        }
    }
    // k_ncc_local: when one wave of CTAs (<= one per SM) covers every track's window with patches of 8 x 5 TR candidates, the
    // K-split and the window statistics move inside the CTA (kernels.cuh).  Plan: the TR with the fewest template rows per thread.
    p.local = LocalCfg{};
    {
        const char* nl = getenv("PVT_NO_LOCAL");
        // lost-object mode: the local pass takes the latency shape too (round 2, late: 43.8 -> 37.9 us per 1080p step with no track
        // lost; PVT_LOCAL_LOST=0 restores the K-split shape there)
        const char* ll = getenv("PVT_LOCAL_LOST");
        const bool lost_ok = !c->lost_mode || !(ll && *ll == '0');
        const bool want = allow_env && !(nl && *nl == '1') && local_kernel(c) == PVT_KERNEL_AUTO && lost_ok && !d.global_pass &&
                          p.tile.pj * p.tile.pd > 1 && !getenv("PVT_PLAN");
        const int nch = d.mtp / 8, Gall = (d.Hmax + kCY - 1) / kCY, bx = (d.Wmax + 7) / 8;
        // statistics: from k_winstats running beside the search on SMs the plan leaves free (one per statistics CTA in the worst
        // placement: nothing can wait for a CTA that has no room), else inside the search CTAs before their loop (+4.6 us on C2)
        const char* ls = getenv("PVT_LOCAL_STATS");
        const long long stat_ctas = p.stat.NX > 0 ? (long long)d.max_tracks * p.stat.xtiles * p.stat.ybands : (1LL << 40);
        double best = 1e300;
        for (int pass_g = 1; want && pass_g >= 0 && p.local.TR == 0; --pass_g)
        for (int TR = 1; TR * kCY * 8 <= 256; ++TR) {
            if (pass_g && ls && *ls == '1') break;
            const int by = (Gall + TR - 1) / TR;
            if ((long long)d.max_tracks * bx * by + (pass_g ? stat_ctas : 0) > sm_count) continue;
            for (int PJ = nch; PJ >= 1; --PJ) {
                if (nch % PJ) continue;
                const int PD = std::min(std::min(d.mth, 32), 256 / (TR * PJ));   // 256 threads keep 255 registers (288 would be allocated as 12 warps: 168)
                if (PD < 1) continue;
                const int nfma = TR * PJ * PD, threads = 32 * ((nfma + 31) / 32);
                if (threads > 256 || TR * kCY * 8 > threads) continue;
                LocalCfg g{};
                g.TR = TR; g.PJ = PJ; g.PD = PD; g.bx = bx; g.by = by; g.nfma = nfma;
                g.gstats = pass_g; g.sNX = std::max(p.stat.NX, 1); g.sNY = std::max(p.stat.NY, 1);
                // update inside the kernel (last CTA of the track): measured 22.2 us per C2 step against 21.6 us with k_update as
                // its own launch (one-shot code behind a ticket runs slower than the boundary it saves): opt-in, PVT_LOCAL_UPDATE=1
                { const char* lu = getenv("PVT_LOCAL_UPDATE"); g.update = (lu && *lu == '1') ? 1 : 0; }
                // k_update behind the search with a programmatic dependency released late (behind the FMA loop): 21.9 -> 20.2 us per C2 step
                { const char* lt = getenv("PVT_LATE_TRIGGER"); g.late_trigger = (lt && *lt == '0') ? 0 : 1; }
                g.P = 8 + d.mtp;
                while (g.P % 8 != 4) ++g.P;
                g.tileH = kCY * TR + d.mth - 1;
                g.TS = d.mth * 8 + 4;
                const size_t red_bytes = (size_t)nfma * kLocalRed * 4;
                if ((size_t)(2 * kCY * TR + 6) * (8 + d.mtw - 1) * 8 > red_bytes) continue;   // the statistics scratch lives in the reduction buffer
                const size_t smem = std::max((size_t)g.tileH * g.P * 4 + (size_t)nch * g.TS * 4 + red_bytes + (size_t)kCY * TR * 8 * 8 + 64, (size_t)d.mth * d.mtw * 4);
                if (smem + 1024 > kSmemBudget) continue;
                const double rows = (double)((d.mth + PD - 1) / PD + kCY - 1) * (nch / PJ);   // per-thread template rows incl. the window preload
                if (rows < best) { best = rows; p.local = g; p.local_smem = smem; }
            }
        }
        if (p.local.TR > 0) {
            { int r_ = raise_smem((const void*)k_ncc_local<kCY>, p.local_smem); if (r_) return r_; }
            d.gridW = 8 * p.local.bx;
            d.gridH = kCY * p.local.TR * p.local.by;
            if (p.local.gstats) p.stat.signal = 1;
            if (getenv("PVT_DEBUG_PLAN"))
                fprintf(stderr, "[pvt] local plan: gstats=%d TR=%d PJ=%d PD=%d ctas/track=%dx%d threads=%d P=%d smem=%zu\n", p.local.gstats, p.local.TR, p.local.PJ, p.local.PD,
                        p.local.bx, p.local.by, 32 * ((p.local.nfma + 31) / 32), p.local.P, p.local_smem);
        }
    }
    return encode_tmap(p.d, p.tile, &p.tmap);
}

// Lost-object mode: the whole-frame pass.  Same kernels, window == the whole NCC map (its DevParams copy carries
// rx = W, ry = H and NCC_GLOBAL_CONFIDENCE as the acceptance threshold), own plan and scratch sized for the full frame.
int build_global_pass(pvt_ctx* c, int sm_count)
{
    Pass& g = c->G;
    g.d = c->d;
    Ctx& d = g.d;
    d.global_pass = 1;
    d.Wmax = d.W; d.Hmax = d.H;                         // upper bounds of the map (W - tw + 1) x (H - th + 1)
    d.VW = (d.Wmax + d.mtw - 1 + 7) & ~7;
    d.maps = nullptr; d.partial = nullptr; d.fringe_acc = nullptr; d.trace = nullptr;
    d.wsum = nullptr;                                   // (its window sums do not fit the local buffer: own array below)
    const size_t win = (size_t)d.Wmax * d.Hmax;
    d.vsum = nullptr; d.vsq = nullptr;
    if (stats_legacy(d.mtw)) {
        { int r = dev_alloc(c, &d.vsum, (size_t)d.max_tracks * (d.Hmax + d.mth) * d.VW, false); if (r) return r; }
        { int r = dev_alloc(c, &d.vsq, (size_t)d.max_tracks * (d.Hmax + d.mth) * d.VW, false); if (r) return r; }
    }
    { int r = dev_alloc(c, &d.denom, (size_t)d.max_tracks * win, false); if (r) return r; }
    { int r = dev_alloc(c, &d.params, 1); if (r) return r; }
    { int r = dev_alloc(c, &d.stream_need, (size_t)d.max_streams); if (r) return r; }
    int r = build_plan(c, g, sm_count, PVT_INGEST_FULL, false);
    if (r) return r;
    if (c->tc_ready) {
        // PVT_KERNEL_TC: the whole-frame search on the tensor cores too, the map cut into column tiles (TcCfg.xtiles).
        // PVT_TC_GLOBAL=0 keeps the FP32 K-split search for this pass (the round-2 state; A/B measurements)
        const char* e = getenv("PVT_TC_GLOBAL");
        if (!(e && *e == '0')) {
            { int r2 = dev_alloc(c, &d.wsum, (size_t)d.max_tracks * win, false); if (r2) return r2; }
            { int r2 = tc_geometry(c, d, sm_count, g.tc, &g.tc_smem, &g.tmap8); if (r2) return r2; }
        }
    }
    c->kps_global = 2;   // k_global_mark + k_step_advance; the conditional body's kernels (they only run while a track is lost) are not counted
    return PVT_OK;
}

// PVT_KERNEL_TC: candidate columns per accumulator.  One accumulator of Wmax columns where that fits (<= 256 columns, <= kTcKMax
// K-steps: the 161-wide windows of C2/C4/C5); else the window is cut into column tiles: a CTA = (track, 128 rows, XW columns).
// The width is chosen on a small cost model of one CTA (tools/tc_timeline.py: ~130 clk per MMA = per (template row, K-step),
// ~900 clk of epilogue per 16 columns, ~6k clk of prologue) times the number of waves the pass's CTAs need on the GPU.
int tc_pick_xw(const Ctx& d, int sm_count)
{
    auto ks_of = [&](int xw) { return (xw + d.mtw - 1 + 15 + 31) / 32; };
    const int NW = (d.Wmax + 15) & ~15;
    if (const char* e = getenv(d.global_pass ? "PVT_TC_XW_GLOBAL" : "PVT_TC_XW")) {   // experiments
        const int xw = (atoi(e) + 15) & ~15;
        if (xw >= 16 && xw <= 256 && ks_of(xw) <= kTcKMax) return std::min(xw, NW);
    }
    if (!d.global_pass && NW <= 256 && ks_of(NW) <= kTcKMax) return NW;
    // the whole-frame pass is sized for ONE lost track and the map it really has, (W - tw + 1) x (H - th + 1)
    const int wt = d.global_pass ? std::max(1, d.Wmax - d.mtw + 1) : d.Wmax, ht = d.global_pass ? std::max(1, d.Hmax - d.mth + 1) : d.Hmax;
    const long long tracks = d.global_pass ? 1 : d.max_tracks;
    int best_xw = 0;
    double best = 1e300;
    for (int xw = 16; xw <= 256 && xw <= NW; xw += 16) {
        const int ks = ks_of(xw);
        if (ks > kTcKMax) break;
        const long long ctas = tracks * ((ht + 127) / 128) * ((wt + xw - 1) / xw);
        const long long waves = (ctas + sm_count - 1) / sm_count;
        const double cost = (double)waves * ((double)d.mth * ks * 130.0 + (xw / 16) * 900.0 + 6000.0);
        if (cost < best) { best = cost; best_xw = xw; }
    }
    return best_xw;
}

// PVT_KERNEL_TC: geometry of k_ncc_tc for one pass (window bounds d.Wmax x d.Hmax), its shared memory and the 4-D tensor map
// (16 B, row, 16-pixel chunk, stream) of the u8 gray plane whose box lands in shared memory chunk-major (ncc_tc.cuh)
// host arithmetic only (pvt_tc_plan_query calls it without a device)
int tc_shape(const Ctx& d, int sm_count, TcCfg& g, size_t* smem)
{
    g = TcCfg{};
    g.XW = tc_pick_xw(d, sm_count);
    g.rows = 128 + d.mth - 1;
    if (g.XW <= 0 || g.rows > 256)
        return fail(PVT_ERR_UNSUPPORTED, "PVT_KERNEL_TC: template height <= 129 and template width <= 260 required");
    g.xtiles = (d.Wmax + g.XW - 1) / g.XW;
    g.AG = g.XW / 8;
    g.KS = (g.XW + d.mtw - 1 + 15 + 31) / 32;
    g.nblk = 4 * g.KS + g.AG;
    g.tpp = (d.mtw + 15) & ~15;
    g.mtiles = (d.Hmax + 127) / 128;
    g.tmem_cols = 32;
    while (g.tmem_cols < 2 * g.XW) g.tmem_cols <<= 1;
    auto smem_of = [&](int stages) {
        const size_t main_bytes = (size_t)16 * g.rows * 2 * g.KS + (size_t)stages * 2 * g.nblk * 128 + ((((size_t)2 * d.mth * (kTcPadL + g.tpp + kTcPadR)) + 15) & ~(size_t)15);
        return std::max(main_bytes, (size_t)kTcEpiBytes) + sizeof(uint64_t) * (2 + 2 * stages) + 64;
    };
    g.stages = 4;
    if (const char* e = getenv("PVT_TC_STAGES")) g.stages = std::max(2, std::min(16, atoi(e)));   // experiments
    if (const char* e = getenv("PVT_TC_SPIN")) g.spin = atoi(e);
    while (g.stages > 2 && smem_of(g.stages) > kSmemBudget - 1024) --g.stages;
    *smem = smem_of(g.stages);
    if (*smem > kSmemBudget - 1024) return fail(PVT_ERR_UNSUPPORTED, "PVT_KERNEL_TC: tile does not fit shared memory");
    return PVT_OK;
}

int tc_geometry(pvt_ctx* c, const Ctx& d, int sm_count, TcCfg& g, size_t* smem, CUtensorMap* tmap8)
{
    (void)c;
    { int r = tc_shape(d, sm_count, g, smem); if (r) return r; }
    { int r = raise_smem((const void*)k_ncc_tc, *smem); if (r) return r; }
    if (getenv("PVT_DEBUG_PLAN"))
        fprintf(stderr, "[pvt] tc plan (%s pass): XW=%d xtiles=%d mtiles=%d KS=%d stages=%d tmem=%d smem=%zu\n", d.global_pass ? "whole-frame" : "local", g.XW, g.xtiles,
                g.mtiles, g.KS, g.stages, g.tmem_cols, *smem);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(PVT_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
    cuuint64_t dims[4] = {16, (cuuint64_t)d.H, (cuuint64_t)(d.pitch8 / 16), (cuuint64_t)d.max_streams};
    cuuint64_t strides[3] = {(cuuint64_t)d.pitch8, 16, (cuuint64_t)d.plane8};
    cuuint32_t box[4] = {16, (cuuint32_t)g.rows, (cuuint32_t)(2 * g.KS), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = ((EncodeTiledFn)fn)(tmap8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, d.gray8, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PVT_ERR_CUDA, "cuTensorMapEncodeTiled (u8 gray plane, 4-D) failed, CUresult " + std::to_string((int)r));
    return PVT_OK;
}

// PVT_KERNEL_TC: the context's buffers (u8 gray plane, template digits, window sums) and the local pass's geometry
int setup_tc(pvt_ctx* c, int sm_count)
{
    Ctx& d = c->d;
    d.pitch8 = (d.W + 15) & ~15;
    d.plane8 = ((size_t)d.pitch8 * d.H + 255) & ~(size_t)255;
    d.tpp = (d.mtw + 15) & ~15;
    { int r = dev_alloc(c, &d.gray8, d.plane8 * d.max_streams); if (r) return r; }
    { int r = dev_alloc(c, &d.tdig, (size_t)d.max_tracks * 2 * d.mth * d.tpp); if (r) return r; }
    { int r = dev_alloc(c, &d.wsum, (size_t)d.max_tracks * d.Wmax * d.Hmax, false); if (r) return r; }
    { int r = tc_geometry(c, d, sm_count, c->tc, &c->tc_smem, &c->tmap8); if (r) return r; }
    c->tc_ready = true;
    return PVT_OK;
}

int upload_params(pvt_ctx* c)
{
    DevParams p{};
    p.rx = c->params.search_radius_x;
    p.ry = c->params.search_radius_y;
    p.min_conf = c->params.ncc_min_confidence;
    p.strong_conf = c->params.ncc_strong_confidence;
    p.lr = c->params.template_update_lr;
    p.keep_maps = c->params.keep_maps;
    p.lost_threshold = c->params.lost_frame_threshold;
    CK(cudaMemcpyAsync(c->d.params, &p, sizeof(p), cudaMemcpyHostToDevice, c->compute));
    CK(cudaStreamSynchronize(c->compute));
    if (c->lost_mode && c->G.d.params) {
        p.rx = c->d.W; p.ry = c->d.H;                       // window == the whole map (tracker_ghc/src/main.cpp:188-193)
        p.min_conf = c->params.ncc_global_confidence;       // :217
        p.keep_maps = 0;
        CK(cudaMemcpyAsync(c->G.d.params, &p, sizeof(p), cudaMemcpyHostToDevice, c->compute));
        CK(cudaStreamSynchronize(c->compute));
    }
    return PVT_OK;
}

int upload_seq(pvt_ctx* c, unsigned long long step0, int ring_len, int row0, int prefetch = 0)
{
    k_seq_begin<<<1, 64, 0, c->compute>>>(c->d, SeqDesc{step0, ring_len, row0, prefetch, 0}, 0);
    CK(cudaGetLastError());
    c->seq_default = (step0 == 0 && ring_len == kRing && row0 == 0 && !prefetch);
    return PVT_OK;
}

int validate_params(const pvt_params* p)
{
    if (!p) return fail(PVT_ERR_INVALID, "params is NULL");
    if (p->mode == PVT_MODE_CPU)
        return fail(PVT_ERR_UNSUPPORTED, "PVT_MODE_CPU: libpvt has no CPU path (the CPU oracle lives in oracle/, test-only)");
    if (p->mode < PVT_MODE_NAIVE || p->mode > PVT_MODE_BATCH) return fail(PVT_ERR_INVALID, "unknown mode");
    if (p->kernel < PVT_KERNEL_AUTO || p->kernel > PVT_KERNEL_TC_GLOBAL) return fail(PVT_ERR_INVALID, "unknown kernel variant");
    if (p->ingest < PVT_INGEST_AUTO || p->ingest > PVT_INGEST_ROI) return fail(PVT_ERR_INVALID, "unknown ingest mode");
    if (p->search_radius_x < 0 || p->search_radius_y < 0) return fail(PVT_ERR_INVALID, "negative search radius");
    if (p->mode == PVT_MODE_BATCH && p->batch_size < 1) return fail(PVT_ERR_INVALID, "batch_size < 1");
    if (!(p->template_update_lr >= 0.0 && p->template_update_lr <= 1.0)) return fail(PVT_ERR_INVALID, "template_update_lr outside [0,1]");
    if (p->lost_frame_threshold < 0) return fail(PVT_ERR_INVALID, "negative lost_frame_threshold");
    if (p->formula != PVT_FORMULA_CCOEFF_NORMED && p->formula != PVT_FORMULA_EPS) return fail(PVT_ERR_INVALID, "unknown formula");
    return PVT_OK;
}

enum { CLS_INGEST = 0, CLS_STATS = 1, CLS_NCC = 2, CLS_UPDATE = 3, CLS_SEARCH_KERNEL = 4 };

// PVT_DEBUG_SYNC=1: launch kernels directly (no graph), synchronise after each and name the one that faulted
bool debug_sync()
{
    static const bool on = [] { const char* e = getenv("PVT_DEBUG_SYNC"); return e && *e && *e != '0'; }();
    return on;
}
int dbg(pvt_ctx* c, const char* what)
{
    if (!debug_sync()) return PVT_OK;
    cudaError_t e = cudaStreamSynchronize(c->compute);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PVT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return PVT_OK;
}

// profiling inside a captured graph: external event-record nodes (GPU-side timestamps, no host launch gaps)
int pnode(pvt_ctx* c, int cls, int which, cudaStream_t st)
{
    CK(cudaEventRecordWithFlags(c->pev[cls][which], st, cudaEventRecordExternal));
    return PVT_OK;
}

// the local pass as the kernels see it (the context's own fields are the master copy)
Pass local_pass(const pvt_ctx* c)
{
    Pass p;
    p.d = c->d; p.tile = c->tile; p.tmap = c->tmap; p.ncc_smem = c->ncc_smem; p.rowsum_warps = c->rowsum_warps; p.rowsum_pw = c->rowsum_pw;
    p.tc = c->tc; p.tc_smem = c->tc_smem; p.tmap8 = c->tmap8;
    if (c->params.kernel == PVT_KERNEL_TC_GLOBAL) p.d.tdig = nullptr;   // the FP32 local pass keeps no template digits: k_track_digits (whole-frame pass)
    p.colprefix_chunks = c->colprefix_chunks; p.stat = c->stat; p.fused = c->fused; p.fused_smem = c->fused_smem; p.local = c->local; p.local_smem = c->local_smem; p.fringe = c->fringe; p.fringe_smem = c->fringe_smem; p.roi_ingest = c->roi_ingest;
    return p;
}
Pass global_pass(const pvt_ctx* c)
{
    Pass p = c->G;
    p.d.trace = nullptr;   // the device timeline describes the local pass
    p.pdl = false;         // body of a conditional graph node: plain stream order
    return p;
}

// Launch with a programmatic dependency on the previous kernel of the stream (PDL): the kernel's CTAs may start while
// that kernel is still draining (it has issued griddepcontrol.launch_dependents), run their preamble, and block in
// griddepcontrol.wait until it has completed -- the launch latency of every such edge (~1.5 us on B200) leaves the
// critical path of a time step.  pdl == false: plain stream order.
template <typename... KArgs, typename... Args>
int launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, kern, args...));
    return PVT_OK;
}

// k_ingest: (column blocks, rows, streams); a block covers 16 * 128 = 2048 pixels of a row on either path
dim3 ingest_grid(const Ctx& d)
{
    return dim3((unsigned)((d.W + 16 * kIngestThreads - 1) / (16 * kIngestThreads)), (unsigned)((d.H + kIngestRows - 1) / kIngestRows), (unsigned)d.max_streams);
}

// The kernels of one searched time step.  capturing: being recorded into a CUDA graph on c->compute (fork/join allowed).
// profile (only while capturing): external event-record NODES around each kernel class, so the measured durations are
// GPU-side and contain no host launch gaps.  Classes: ingest | statistics | search (k_ncc_search [+ tail reduction]) |
// update (k_ncc_finalize incl. the fused update in K-split mode, else k_update).
int launch_step_kernels(pvt_ctx* c, const Pass& p, bool profile, bool capturing = false)
{
    const Ctx& d = p.d;
    const int kern = pass_kernel(c, d);
    profile = profile && capturing;
    if (profile) { int r = pnode(c, CLS_INGEST, 0, c->compute); if (r) return r; }
    if (p.roi_ingest) {
        const int roi_groups = ((d.VW + 4 + 3) / 4 + 1) * (d.Hmax + d.mth);
        // (a programmatic edge update(k) ~> ingest(k+1) was measured again in round 2, with the trigger late in the update: the gap
        //  stays 1.8 us -- it is the update grid's completion + flush, not the ingest's launch)
        k_ingest_roi<<<dim3((roi_groups + 255) / 256, d.max_tracks), 256, 0, c->compute>>>(d);
    } else {
        k_ingest<<<ingest_grid(d), kIngestThreads, 0, c->compute>>>(d);
    }
    if (profile) { int r = pnode(c, CLS_INGEST, 1, c->compute); if (r) return r; }
    { int r = dbg(c, "k_ingest"); if (r) return r; }
    // pinned host rings: stage the next step's pixels beside this step (a branch that joins before the box moves)
    bool join3 = false;
    // (k_ncc_local with its statistics on a parallel branch fills all but a few SMs with one CTA each: a prefetch CTA that lands on
    //  an SM first keeps a search CTA out until it has finished its PCIe reads -- measured: e2e 37.3k -> 31.3k frames/s.  There the
    //  prefetch is launched BEHIND k_winstats on the statistics branch instead: by then every search CTA has its SM.)
    const bool local_gs = p.local.TR > 0 && p.local.gstats && kern == PVT_KERNEL_AUTO && !d.global_pass;
    const bool pf_on_stats_branch = local_gs && capturing && d.stage && p.roi_ingest && p.prefetch;
    if (d.stage && p.roi_ingest && !d.global_pass && p.prefetch && !pf_on_stats_branch) {
        const dim3 pgrid(24, (unsigned)d.max_tracks);   // small on purpose: see k_prefetch_roi
        if (capturing) {
            CK(cudaEventRecord(c->ev_fork, c->compute));
            CK(cudaStreamWaitEvent(c->aux3, c->ev_fork, 0));
            k_prefetch_roi<<<pgrid, 256, 0, c->aux3>>>(d, c->prefetch_delay_ns);
            CK(cudaEventRecord(c->ev_join3, c->aux3));
            join3 = true;
        } else {
            k_prefetch_roi<<<pgrid, 256, 0, c->compute>>>(d, 0);
            { int r = dbg(c, "k_prefetch_roi"); if (r) return r; }
        }
    }

    if (p.local.TR > 0 && kern == PVT_KERNEL_AUTO && !d.global_pass) {
        // latency shape, K-split and statistics inside the CTA: ingest ~> k_ncc_local -> k_update
        const bool pdl_l = capturing && !profile && p.pdl;
        const unsigned threads = 32u * (unsigned)((p.local.nfma + 31) / 32);
        const bool forkl = capturing && p.local.gstats;
        cudaStream_t sst = forkl ? c->aux : c->compute;
        if (forkl) {
            CK(cudaEventRecord(c->ev_fork, c->compute));
            CK(cudaStreamWaitEvent(c->aux, c->ev_fork, 0));
        }
        // (profiling graph: no event nodes on the statistics branch -- they delay k_winstats, and the search would be timed waiting for it)
        if (profile) { int r = pnode(c, CLS_STATS, 0, c->compute); if (r) return r; r = pnode(c, CLS_STATS, 1, c->compute); if (r) return r; }
        if (p.local.gstats) { int r = launch_pdl(k_winstats, dim3((unsigned)(p.stat.xtiles * p.stat.ybands), d.max_tracks), dim3(kStatThreads), 0, sst, pdl_l, d, p.stat); if (r) return r; }
        if (pf_on_stats_branch) { int r = launch_pdl(k_prefetch_roi, dim3(24, (unsigned)d.max_tracks), dim3(256), 0, c->aux, pdl_l, d, c->prefetch_delay_ns); if (r) return r; }
        if (forkl) CK(cudaEventRecord(c->ev_join, c->aux));
        if (profile) { int r = pnode(c, CLS_NCC, 0, c->compute); if (r) return r; r = pnode(c, CLS_SEARCH_KERNEL, 0, c->compute); if (r) return r; }
        { int r = launch_pdl(k_ncc_local<kCY>, dim3((unsigned)(d.max_tracks * p.local.bx * p.local.by)), dim3(threads), p.local_smem, c->compute, pdl_l, d, p.local); if (r) return r; }
        if (profile) { int r = pnode(c, CLS_SEARCH_KERNEL, 1, c->compute); if (r) return r; r = pnode(c, CLS_NCC, 1, c->compute); if (r) return r; }
        { int r = dbg(c, "k_ncc_local"); if (r) return r; }
        if (p.local.update) {   // the update ran inside k_ncc_local (last CTA of each track)
            if (profile) { int r = pnode(c, CLS_UPDATE, 0, c->compute); if (r) return r; r = pnode(c, CLS_UPDATE, 1, c->compute); if (r) return r; }
            if (forkl) CK(cudaStreamWaitEvent(c->compute, c->ev_join, 0));
            if (join3) CK(cudaStreamWaitEvent(c->compute, c->ev_join3, 0));
            CK(cudaGetLastError());
            return PVT_OK;
        }
        if (profile) { int r = pnode(c, CLS_UPDATE, 0, c->compute); if (r) return r; }
        // Programmatic edge search ~> update: the update's CTA becomes resident when every search CTA is past its FMA loop
        // (LocalCfg.late_trigger), reads the step, the track and the old template, and blocks in griddepcontrol.wait until the peak
        // exists: 21.9 -> 20.2 us per step.  (Released at the START of the search the early CTA slowed the search: 23.4 us.)
        if (p.local.late_trigger) { int r = launch_pdl(k_update, dim3((unsigned)d.max_tracks), dim3(256), c->templ_smem, c->compute, pdl_l, d); if (r) return r; }
        else k_update<<<d.max_tracks, 256, c->templ_smem, c->compute>>>(d);
        if (profile) { int r = pnode(c, CLS_UPDATE, 1, c->compute); if (r) return r; }
        { int r = dbg(c, "k_update"); if (r) return r; }
        if (forkl) CK(cudaStreamWaitEvent(c->compute, c->ev_join, 0));   // the statistics branch ends inside this step
        if (join3) CK(cudaStreamWaitEvent(c->compute, c->ev_join3, 0));
        CK(cudaGetLastError());
        return PVT_OK;
    }
    // K-split mode inside a captured graph: the statistics kernels and the search only meet in k_ncc_finalize, so they
    // run as two concurrent branches (fork after ingest, join before finalize)
    const bool tc = kern == PVT_KERNEL_TC && p.tc.XW > 0;   // tensor-core search: ingest -> statistics -> k_ncc_tc -> update
    const bool ksplit = kern == PVT_KERNEL_AUTO ? p.tile.pj * p.tile.pd > 1
                                                             : (kern == PVT_KERNEL_TC && !tc && p.tile.pj * p.tile.pd > 1);
    const bool fork = capturing && ksplit;
    // PVT_KERNEL_TC_GLOBAL: the template digits of the tracks this whole-frame pass owns (k_track_digits), beside the statistics
    bool join_dig = false;
    if (tc && d.global_pass && c->params.kernel == PVT_KERNEL_TC_GLOBAL) {
        if (capturing) {
            CK(cudaEventRecord(c->ev_fork, c->compute));
            CK(cudaStreamWaitEvent(c->aux2, c->ev_fork, 0));
            k_track_digits<<<d.max_tracks, 256, c->templ_smem, c->aux2>>>(d);
            CK(cudaEventRecord(c->ev_join2, c->aux2));
            join_dig = true;
        } else {
            k_track_digits<<<d.max_tracks, 256, c->templ_smem, c->compute>>>(d);
            { int r = dbg(c, "k_track_digits"); if (r) return r; }
        }
    }
    cudaStream_t sstats = c->compute;
    if (fork) {
        CK(cudaEventRecord(c->ev_fork, c->compute));
        CK(cudaStreamWaitEvent(c->aux, c->ev_fork, 0));
        sstats = c->aux;
    }
    if (profile) { int r = pnode(c, CLS_STATS, 0, sstats); if (r) return r; }
    const bool pdl = capturing && !profile && p.pdl;   // programmatic edges: ingest ~> search / colprefix, colprefix ~> rowsum, search ~> finalize
    if (p.stat.NX > 0) {
        { int r = launch_pdl(k_winstats, dim3((unsigned)(p.stat.xtiles * p.stat.ybands), d.max_tracks), dim3(kStatThreads), 0, sstats, pdl, d, p.stat); if (r) return r; }
    } else {
        { int r = launch_pdl(k_colprefix, dim3((d.VW + 31) / 32, d.max_tracks), dim3(32, p.colprefix_chunks), 0, sstats, pdl, d); if (r) return r; }
        { int r = dbg(c, "k_colprefix"); if (r) return r; }
        { int r = launch_pdl(k_rowsum, dim3((d.Hmax + p.rowsum_warps - 1) / p.rowsum_warps, d.max_tracks), dim3(p.rowsum_warps * 32),
                             (size_t)p.rowsum_warps * 2 * p.rowsum_pw * sizeof(double), sstats, pdl, d, p.rowsum_pw); if (r) return r; }
    }
    if (profile) { int r = pnode(c, CLS_STATS, 1, sstats); if (r) return r; }
    { int r = dbg(c, "k_rowsum"); if (r) return r; }
    // the candidates outside the thread-tile grid: after the statistics (they need the normaliser), beside the search
    const bool fringe = !tc && kern != PVT_KERNEL_DIRECT && p.fringe.colg + p.fringe.rowg > 0;
    const dim3 fgrid((unsigned)(p.fringe.colg + p.fringe.rowg), (unsigned)d.max_tracks, (unsigned)(p.fringe.defer ? p.tile.pd : 1));
    if (fork) CK(cudaEventRecord(c->ev_join, c->aux));

    if (profile) { int r = pnode(c, CLS_NCC, 0, c->compute); if (r) return r; }
    bool join2 = false, fringe_after = false;
    if (fringe) {
        if (fork) {
            // latency shape: a third branch from the same fork point; the cross terms are left in fringe_acc
            // (FringeCfg.defer) and normalised by k_ncc_finalize, so this branch does not wait for the statistics
            CK(cudaStreamWaitEvent(c->aux2, c->ev_fork, 0));
            k_ncc_fringe<<<fgrid, p.fringe.threads, p.fringe_smem, c->aux2>>>(d, p.tile, p.fringe);
            CK(cudaEventRecord(c->ev_join2, c->aux2));
            join2 = true;
        } else if (capturing && p.pdl) {
            fringe_after = true;   // throughput shape: right behind the search kernel with a programmatic dependency (below)
        } else {
            k_ncc_fringe<<<fgrid, p.fringe.threads, p.fringe_smem, c->compute>>>(d, p.tile, p.fringe);
            { int r = dbg(c, "k_ncc_fringe"); if (r) return r; }
        }
    }
    if (tc) {
        if (profile) { int r = pnode(c, CLS_SEARCH_KERNEL, 0, c->compute); if (r) return r; }
        if (join_dig) CK(cudaStreamWaitEvent(c->compute, c->ev_join2, 0));
        k_ncc_tc<<<(unsigned)(d.max_tracks * p.tc.mtiles * p.tc.xtiles), kTcThreads, p.tc_smem, c->compute>>>(d, p.tc, p.tmap8);
        if (profile) { int r = pnode(c, CLS_SEARCH_KERNEL, 1, c->compute); if (r) return r; }
        if (profile) { int r = pnode(c, CLS_NCC, 1, c->compute); if (r) return r; }
    } else if (kern == PVT_KERNEL_DIRECT) {
        k_ncc_direct<<<dim3((d.Wmax * d.Hmax + 255) / 256, d.max_tracks), 256, 0, c->compute>>>(d);
        if (profile) { int r = pnode(c, CLS_NCC, 1, c->compute); if (r) return r; }
    } else {
        const int parts = p.tile.pj * p.tile.pd;
        const unsigned nbx = (unsigned)(p.tile.n_full + p.tile.n_tail * std::max(p.tile.tail_ps, 1));
        if (profile) { int r = pnode(c, CLS_SEARCH_KERNEL, 0, c->compute); if (r) return r; }
        // (K-split shape: the search directly follows the ingest on this stream and waits for it before fetching its tile.
        //  Unsplit shape: a plain dependency on k_rowsum -- starting the search during the statistics' tail was measured and
        //  lost 4 % on C4: k_ncc_fringe then has to wait for the whole search grid before it may read the normalisers)
        const bool fused = p.fused && kern == PVT_KERNEL_AUTO && !d.global_pass;
        if (fused) {
            // search + second stage + update in one launch; it waits for k_winstats (the other branch) through TrackState.stats_done
            { int r = launch_pdl(k_step_fused<kCY>, dim3(nbx * (unsigned)parts), dim3(kTilesPerCta), p.fused_smem, c->compute, pdl && fork, d, p.tile, p.tmap, p.stat); if (r) return r; }
            if (profile) {
                for (int w = 0; w < 2; ++w) { int r = pnode(c, CLS_SEARCH_KERNEL, 1, c->compute); if (r) return r; r = pnode(c, CLS_NCC, 1, c->compute); if (r) return r; }
                { int r = pnode(c, CLS_UPDATE, 0, c->compute); if (r) return r; }
                { int r = pnode(c, CLS_UPDATE, 1, c->compute); if (r) return r; }
            }
            if (fork) CK(cudaStreamWaitEvent(c->compute, c->ev_join, 0));   // the statistics branch ends inside this step
            { int r = dbg(c, "k_step_fused"); if (r) return r; }
            if (join3) CK(cudaStreamWaitEvent(c->compute, c->ev_join3, 0));
            CK(cudaGetLastError());
            return PVT_OK;
        }
        { int r = launch_pdl(k_ncc_search<kCY>, dim3(nbx, 1, parts), dim3(kTilesPerCta), p.ncc_smem, c->compute, pdl && fork, d, p.tile, p.tmap); if (r) return r; }
        if (profile) { int r = pnode(c, CLS_SEARCH_KERNEL, 1, c->compute); if (r) return r; }
        if (fringe_after) {
            // programmatic dependent launch: the fringe kernel may begin once all search CTAs have been dispatched
            // (k_ncc_search issues griddepcontrol.launch_dependents first thing); it needs none of the search's results
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = fgrid; cfg.blockDim = dim3(p.fringe.threads); cfg.dynamicSmemBytes = p.fringe_smem; cfg.stream = c->compute;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = profile ? 0 : 1;   // the profiling graph times k_ncc_search alone
            cfg.attrs = at; cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, k_ncc_fringe, d, p.tile, p.fringe));
        }
        if (p.tile.tail_ps > 1) k_ncc_tail_finalize<kCY><<<p.tile.n_tail, kTilesPerCta, 0, c->compute>>>(d, p.tile);
        if (join2) CK(cudaStreamWaitEvent(c->compute, c->ev_join2, 0));   // the search class ends when the fringe has ended too
        if (profile) { int r = pnode(c, CLS_NCC, 1, c->compute); if (r) return r; }
        if (fork) CK(cudaStreamWaitEvent(c->compute, c->ev_join, 0));
        if (parts > 1) {
            const int fthr = d.mth * d.mtw > 8192 ? kFinalizeThreads : 256;
            if (profile) { int r = pnode(c, CLS_UPDATE, 0, c->compute); if (r) return r; }
            { int r = launch_pdl(k_ncc_finalize, dim3((d.Wmax * d.Hmax + fthr - 1) / fthr, d.max_tracks), dim3(fthr), c->templ_smem, c->compute, pdl && fork, d, p.tile); if (r) return r; }
            if (profile) { int r = pnode(c, CLS_UPDATE, 1, c->compute); if (r) return r; }
        }
    }
    { int r = dbg(c, "k_ncc"); if (r) return r; }

    if (!ksplit) {   // otherwise the update ran inside k_ncc_finalize
        if (profile) { int r = pnode(c, CLS_UPDATE, 0, c->compute); if (r) return r; }
        k_update<<<d.max_tracks, 256, c->templ_smem, c->compute>>>(d);
        if (profile) { int r = pnode(c, CLS_UPDATE, 1, c->compute); if (r) return r; }
        { int r = dbg(c, "k_update"); if (r) return r; }
    }
    if (join3) CK(cudaStreamWaitEvent(c->compute, c->ev_join3, 0));   // the prefetch branch ends with the step
    CK(cudaGetLastError());
    return PVT_OK;
}

// Lost-object mode: append  k_global_mark -> IF (any track lost) { the whole-frame pass } -> k_step_advance  to graph g behind
// `deps`.  The condition is set on the device by k_global_mark, so nothing returns to the host between the passes.
int add_global_tail(pvt_ctx* c, cudaGraph_t g, const cudaGraphNode_t* deps, size_t n_deps, cudaGraphNode_t* last = nullptr)
{
    const Pass gp = global_pass(c);
    Ctx gd = gp.d;
    cudaGraphConditionalHandle cond;
    CK(cudaGraphConditionalHandleCreate(&cond, g, 0, cudaGraphCondAssignDefault));
    int use_cond = 1;
    void* margs[3] = {&gd, &cond, &use_cond};
    cudaKernelNodeParams kp{};
    kp.func = (void*)k_global_mark; kp.gridDim = dim3(1); kp.blockDim = dim3(256); kp.sharedMemBytes = 0; kp.kernelParams = margs;
    cudaGraphNode_t n_mark, n_if, n_adv;
    CK(cudaGraphAddKernelNode(&n_mark, g, deps, n_deps, &kp));
    cudaGraphNodeParams cp{};
    cp.type = cudaGraphNodeTypeConditional;
    cp.conditional.handle = cond; cp.conditional.type = cudaGraphCondTypeIf; cp.conditional.size = 1;
    CK(cudaGraphAddNode(&n_if, g, &n_mark, 1, &cp));
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    CK(cudaStreamBeginCaptureToGraph(c->compute, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    int r = launch_step_kernels(c, gp, false, true);
    cudaError_t e = cudaStreamEndCapture(c->compute, nullptr);
    if (r) return r;
    CK(e);
    void* aargs[1] = {&gd};
    cudaKernelNodeParams ka{};
    ka.func = (void*)k_step_advance; ka.gridDim = dim3(1); ka.blockDim = dim3(32); ka.sharedMemBytes = 0; ka.kernelParams = aargs;
    CK(cudaGraphAddKernelNode(&n_adv, g, &n_if, 1, &ka));
    if (last) *last = n_adv;
    return PVT_OK;
}

// Lost-object mode: n consecutive time steps in one graph.  Per step: the local pass captured behind the previous step's
// k_step_advance, then (behind the local pass's leaves) k_global_mark -> IF { whole-frame pass } -> k_step_advance with a
// conditional handle of its own.
int capture_lost_steps_graph(pvt_ctx* c, const Pass& lp, int n, cudaGraphExec_t* out)
{
    cudaGraph_t g = nullptr;
    CK(cudaGraphCreate(&g, 0));
    cudaGraphNode_t prev{};
    for (int k = 0; k < n; ++k) {
        CK(cudaStreamBeginCaptureToGraph(c->compute, g, k ? &prev : nullptr, nullptr, k ? 1 : 0, cudaStreamCaptureModeThreadLocal));
        int r = launch_step_kernels(c, lp, false, true);
        cudaError_t e = cudaStreamEndCapture(c->compute, nullptr);
        if (r) { cudaGraphDestroy(g); return r; }
        CK(e);
        size_t nn = 0, ne = 0;
        CK(cudaGraphGetNodes(g, nullptr, &nn));
        std::vector<cudaGraphNode_t> nodes(nn);
        CK(cudaGraphGetNodes(g, nodes.data(), &nn));
        CK(cudaGraphGetEdges_v2(g, nullptr, nullptr, nullptr, &ne));   // _v2: the graph has programmatic edges
        std::vector<cudaGraphNode_t> from(ne), to(ne);
        std::vector<cudaGraphEdgeData> ed(ne);
        if (ne) CK(cudaGraphGetEdges_v2(g, from.data(), to.data(), ed.data(), &ne));
        std::vector<cudaGraphNode_t> leaves;   // earlier steps end in their k_step_advance, which this step's roots depend on
        for (cudaGraphNode_t nd : nodes)
            if (std::find(from.begin(), from.end(), nd) == from.end()) leaves.push_back(nd);
        r = add_global_tail(c, g, leaves.data(), leaves.size(), &prev);
        if (r) { cudaGraphDestroy(g); return r; }
    }
    CK(cudaGraphInstantiate(out, g, 0));
    CK(cudaGraphDestroy(g));
    CK(cudaGraphUpload(*out, c->compute));
    return PVT_OK;
}

// every kernel finds its frame through the device-side step counter, so consecutive steps can share one launch
int capture_steps_graph(pvt_ctx* c, const Pass& p, int n, cudaGraphExec_t* out)
{
    cudaGraph_t gs = nullptr;
    CK(cudaStreamBeginCapture(c->compute, cudaStreamCaptureModeThreadLocal));
    int rr = PVT_OK;
    for (int k = 0; k < n && !rr; ++k) rr = launch_step_kernels(c, p, false, true);
    cudaError_t ee = cudaStreamEndCapture(c->compute, &gs);
    if (rr) return rr;
    CK(ee);
    CK(cudaGraphInstantiate(out, gs, 0));
    CK(cudaGraphDestroy(gs));
    CK(cudaGraphUpload(*out, c->compute));   // the first launch must not pay the upload inside a caller's timed loop
    return PVT_OK;
}

// latency shape: the graph of exactly n steps (built on first use)
int steps_graph(pvt_ctx* c, int n, bool pf, cudaGraphExec_t* out)
{
    cudaGraphExec_t& g = c->graph_n[pf ? 1 : 0][n];
    if (!g) {
        Pass p = local_pass(c);
        p.prefetch = pf;
        int r = c->lost_mode ? capture_lost_steps_graph(c, p, n, &g) : capture_steps_graph(c, p, n, &g);
        if (r) return r;
    }
    *out = g;
    return PVT_OK;
}

int build_graphs(pvt_ctx* c)
{
    if (c->graph) { cudaGraphExecDestroy(c->graph); c->graph = nullptr; }
    if (c->graph_hold) { cudaGraphExecDestroy(c->graph_hold); c->graph_hold = nullptr; }
    cudaGraph_t g = nullptr;
    const Pass lp = local_pass(c);
    int r;
    cudaError_t e;
    if (!c->lost_mode) {
        CK(cudaStreamBeginCapture(c->compute, cudaStreamCaptureModeThreadLocal));
        r = launch_step_kernels(c, lp, false, true);
        e = cudaStreamEndCapture(c->compute, &g);
        if (r) return r;
        CK(e);
        CK(cudaGraphInstantiate(&c->graph, g, 0));
        CK(cudaGraphDestroy(g));
        CK(cudaGraphUpload(c->graph, c->compute));
    } else {
        // the local pass, then (behind its leaves) the conditional whole-frame pass
        if ((r = capture_lost_steps_graph(c, lp, 1, &c->graph))) return r;
    }
    for (cudaGraphExec_t* ge : {&c->graph_multi, &c->graph_long, &c->graph_pf, &c->graph_multi_pf, &c->graph_long_pf})
        if (*ge) { cudaGraphExecDestroy(*ge); *ge = nullptr; }
    for (int a = 0; a < 2; ++a)
        for (int n = 0; n <= kMaxGraphSteps; ++n)
            if (c->graph_n[a][n]) { cudaGraphExecDestroy(c->graph_n[a][n]); c->graph_n[a][n] = nullptr; }
    auto capture_steps = [&](const Pass& p, int n, cudaGraphExec_t* out) -> int { return capture_steps_graph(c, p, n, out); };
    const bool latency_shape = c->params.kernel != PVT_KERNEL_DIRECT && lp.tile.pj * lp.tile.pd > 1;
    if (!c->lost_mode) {
        if ((r = capture_steps(lp, kMultiStep, &c->graph_multi))) return r;
        if (latency_shape && (r = capture_steps(lp, kLongStep, &c->graph_long))) return r;
    }
    if (!c->lost_mode && c->d.stage) {
        Pass pp = lp;
        pp.prefetch = true;
        if ((r = capture_steps(pp, 1, &c->graph_pf))) return r;
        if ((r = capture_steps(pp, kMultiStep, &c->graph_multi_pf))) return r;
        if (latency_shape && (r = capture_steps(pp, kLongStep, &c->graph_long_pf))) return r;
    }
    if (c->graph_prof) { cudaGraphExecDestroy(c->graph_prof); c->graph_prof = nullptr; }
    CK(cudaStreamBeginCapture(c->compute, cudaStreamCaptureModeThreadLocal));
    r = launch_step_kernels(c, lp, true, true);
    e = cudaStreamEndCapture(c->compute, &g);
    if (r) return r;
    CK(e);
    CK(cudaGraphInstantiate(&c->graph_prof, g, 0));
    CK(cudaGraphDestroy(g));
    if (c->graph_global) { cudaGraphExecDestroy(c->graph_global); c->graph_global = nullptr; }
    if (c->lost_mode) {
        // stand-alone copy of the whole-frame tail, launched behind the profiling graph
        CK(cudaGraphCreate(&g, 0));
        r = add_global_tail(c, g, nullptr, 0);
        if (r) return r;
        CK(cudaGraphInstantiate(&c->graph_global, g, 0));
        CK(cudaGraphDestroy(g));
    }
    CK(cudaStreamBeginCapture(c->compute, cudaStreamCaptureModeThreadLocal));
    k_hold<<<1, 256, 0, c->compute>>>(c->d);
    CK(cudaStreamEndCapture(c->compute, &g));
    CK(cudaGraphInstantiate(&c->graph_hold, g, 0));
    CK(cudaGraphDestroy(g));
    c->graph_valid = true;
    return PVT_OK;
}

size_t frame_row_bytes(const pvt_ctx* c, int format)
{
    return (size_t)c->cfg.frame_w * (format == PVT_FMT_BGR8 ? 3 : format == PVT_FMT_GRAY8 ? 1 : 4);
}

int check_frame(const pvt_ctx* c, const pvt_frame* f)
{
    if (!f->data) return fail(PVT_ERR_INVALID, "frame.data is NULL");
    if (f->stream < 0 || f->stream >= c->cfg.max_streams) return fail(PVT_ERR_INVALID, "frame.stream out of range");
    if (f->format < PVT_FMT_BGR8 || f->format > PVT_FMT_GRAYF32) return fail(PVT_ERR_INVALID, "unknown frame format");
    if (f->memory != PVT_MEM_HOST && f->memory != PVT_MEM_DEVICE && f->memory != PVT_MEM_HOST_PINNED) return fail(PVT_ERR_INVALID, "unknown frame memory kind");
    if (f->step < frame_row_bytes(c, f->format)) return fail(PVT_ERR_INVALID, "frame.step smaller than one row");
    if (f->format == PVT_FMT_GRAYF32 && (f->step % 4 || ((size_t)f->data) % 4)) return fail(PVT_ERR_INVALID, "f32 frame not 4-byte aligned");
    if (f->format == PVT_FMT_GRAYF32 && wants_tc(c->params.kernel))
        return fail(PVT_ERR_INVALID, "PVT_KERNEL_TC searches the 8-bit gray levels: GRAYF32 frames are not accepted by this context");
    return PVT_OK;
}

// Enqueue one time step.  hold: batch-mode frame that is not searched (main.cpp:118-123).
int enqueue_step(pvt_ctx* c, int n_frames, const pvt_frame* frames, bool hold)
{
    const int slot = (int)(c->submitted % kRing);
    const int ms = c->cfg.max_streams;
    if (!c->seq_default) { int r = upload_seq(c, 0, kRing, 0); if (r) return r; }
    FrameDesc* row = c->h_table + (size_t)slot * ms;
    c->seq_rows_valid = false;   // this step's row overwrites part of what a resident sequence uploaded
    if (c->table_ev_used[slot]) CK(cudaEventSynchronize(c->table_ev[slot]));  // previous upload of this pinned row is done
    for (int s = 0; s < ms; ++s) row[s] = FrameDesc{nullptr, 0, 0, 0};
    const int sd = (int)(c->submitted % kStageDepth);
    bool copied = false;
    if (!hold) {
        for (int i = 0; i < n_frames; ++i) {
            const pvt_frame* f = frames + i;
            int r = check_frame(c, f);
            if (r) return r;
            if (row[f->stream].valid) return fail(PVT_ERR_INVALID, "two frames for one stream in one step");
            FrameDesc fd{f->data, (unsigned long long)f->step, f->format, 1};
            bool zero_copy = false;
            const bool host = f->memory != PVT_MEM_DEVICE;
            if (host && c->roi_ingest) {
                // pinned (cudaHostAlloc / cudaHostRegister) memory is readable from the device under UVA: let k_ingest_roi
                // pull just the search tiles over PCIe; pageable memory falls back to the staged full-frame copy
                cudaPointerAttributes pa{};
                if (cudaPointerGetAttributes(&pa, f->data) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer) {
                    fd.data = pa.devicePointer;
                    zero_copy = true;
                } else {
                    cudaGetLastError();
                }
            }
            if (host && !zero_copy) {
                void*& st = c->stage[(size_t)f->stream * kStageDepth + sd];
                if (!st) {
                    CK(cudaMalloc(&st, c->stage_bytes));
                    c->allocs.push_back(st);
                }
                if (!copied && c->ev_done_used[sd]) CK(cudaStreamWaitEvent(c->copy, c->ev_done[sd], 0));  // staging slot free again
                const size_t rb = frame_row_bytes(c, f->format);
                CK(cudaMemcpy2DAsync(st, rb, f->data, f->step, rb, c->cfg.frame_h, cudaMemcpyHostToDevice, c->copy));
                copied = true;
                fd.data = st;
                fd.step = rb;
            }
            row[f->stream] = fd;
            c->plane_full[f->stream] = !c->roi_ingest;   // the ROI ingest refreshes the tracks' search tiles only
            c->prof.ingest_bytes += (c->profiling && !c->roi_ingest) ? (double)c->cfg.frame_w * c->cfg.frame_h * ((f->format == PVT_FMT_BGR8 ? 3 : f->format == PVT_FMT_GRAY8 ? 1 : 4) + 4) : 0.0;
        }
    }
    if (copied) {
        CK(cudaEventRecord(c->ev_copied[sd], c->copy));
        CK(cudaStreamWaitEvent(c->compute, c->ev_copied[sd], 0));
    }
    CK(cudaMemcpyAsync(c->d.table + (size_t)slot * ms, row, sizeof(FrameDesc) * ms, cudaMemcpyHostToDevice, c->compute));
    CK(cudaEventRecord(c->table_ev[slot], c->compute));
    c->table_ev_used[slot] = true;
    if (hold) {
        if (!c->graph_valid) { int r = build_graphs(c); if (r) return r; }
        CK(cudaGraphLaunch(c->graph_hold, c->compute));
        c->launches += 1;
    } else if (debug_sync()) {
        int r = launch_step_kernels(c, local_pass(c), false);
        if (r) return r;
        c->launches += c->kps;
        if (c->lost_mode) {
            const Pass gp = global_pass(c);
            k_global_mark<<<1, 256, 0, c->compute>>>(gp.d, 0, 0);
            r = launch_step_kernels(c, gp, false);
            if (r) return r;
            k_step_advance<<<1, 32, 0, c->compute>>>(gp.d);
            { int r2 = dbg(c, "k_step_advance"); if (r2) return r2; }
            const bool gtc = wants_tc(c->params.kernel) && gp.tc.XW > 0;   // ingest, statistics (1 or 2 kernels), k_ncc_tc, update
            c->launches += c->kps_global + (gtc ? (gp.stat.NX > 0 ? 4 : 5) + (c->params.kernel == PVT_KERNEL_TC_GLOBAL ? 1 : 0) : pass_kernels(gp.tile, gp.fringe, gp.stat));
        }
    } else if (c->profiling) {
        // measurement pass: the same graph with event-record nodes around every kernel class, one step at a time
        if (!c->graph_valid) { int r = build_graphs(c); if (r) return r; }
        CK(cudaGraphLaunch(c->graph_prof, c->compute));
        if (c->lost_mode) { CK(cudaGraphLaunch(c->graph_global, c->compute)); c->launches += c->kps_global; }
        CK(cudaStreamSynchronize(c->compute));
        c->launches += c->kps;
        c->prof.steps += 1;
        double* acc[4] = {&c->prof.ingest_ms, &c->prof.stats_ms, &c->prof.ncc_ms, &c->prof.update_ms};
        int64_t* cnt[4] = {&c->prof.ingest_launches, &c->prof.stats_launches, &c->prof.ncc_launches, &c->prof.update_launches};
        for (int k = 0; k < 4; ++k) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, c->pev[k][0], c->pev[k][1]));
            *acc[k] += ms;
            *cnt[k] += (k == CLS_STATS) ? 2 : 1;
        }
        if (c->params.kernel != PVT_KERNEL_DIRECT) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, c->pev[CLS_SEARCH_KERNEL][0], c->pev[CLS_SEARCH_KERNEL][1]));
            c->prof.search_kernel_ms += ms;
        }
    } else {
        if (!c->graph_valid) { int r = build_graphs(c); if (r) return r; }
        CK(cudaGraphLaunch(c->graph, c->compute));   // lost-object mode: includes the conditional whole-frame pass
        c->launches += c->kps + (c->lost_mode ? c->kps_global : 0);
    }
    if (copied) {
        CK(cudaEventRecord(c->ev_done[sd], c->compute));
        c->ev_done_used[sd] = true;
    }
    c->submitted += 1;
    return PVT_OK;
}

// a batch of result rows on its way to the host (pvt_submit_sequence, resident rings)
struct Readback {
    unsigned long long first;   // first time step of the batch
    int n, out_row, ev;         // steps, first row in results_out, event pair
};
int finish_readback(pvt_ctx* c, Readback& rb, pvt_result* results_out, int mt)
{
    CK(cudaEventSynchronize(c->rb_done[rb.ev]));
    if (results_out)
        for (int k = 0; k < rb.n; ++k)
            std::memcpy(results_out + (size_t)(rb.out_row + k) * mt, c->h_results + (size_t)((rb.first + k) % kRing) * mt, sizeof(pvt_result) * mt);
    rb.n = 0;
    return PVT_OK;
}

int resolve_profile(pvt_ctx* c)
{
    CK(cudaStreamSynchronize(c->compute));
    for (size_t i = 0; i < c->ev_next; ++i) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, c->ev_pool[i].a, c->ev_pool[i].b));
        switch (c->ev_pool[i].cls) {
            case CLS_INGEST: c->prof.ingest_ms += ms; c->prof.ingest_launches += 1; break;
            case CLS_STATS: c->prof.stats_ms += ms; c->prof.stats_launches += 2; break;
            case CLS_NCC: c->prof.ncc_ms += ms; c->prof.ncc_launches += 1; break;
            default: c->prof.update_ms += ms; c->prof.update_launches += 1; break;
        }
    }
    c->ev_next = 0;
    return PVT_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

int pvt_version(void) { return PVT_VERSION; }

// The k_ncc_search plan pvt_create would derive, without touching a device (host logic only; used by the CPU tests).
int pvt_plan_query(int sm_count, int n_tracks, int templ_w, int templ_h, int frame_w, int frame_h, int radius_x, int radius_y, int32_t out[16])
{
    if (!out || sm_count <= 0 || n_tracks <= 0 || templ_w <= 0 || templ_h <= 0 || templ_w > frame_w || templ_h > frame_h || radius_x < 0 || radius_y < 0)
        return fail(PVT_ERR_INVALID, "bad plan query");
    const int mtp = (templ_w + 7) & ~7;
    const int Wmax = std::min(2 * radius_x + 1, frame_w), Hmax = std::min(2 * radius_y + 1, frame_h);
    TileCfg g{};
    size_t smem = 0;
    if (!choose_plan(sm_count, n_tracks, templ_w, mtp, templ_h, Wmax, Hmax, &g, &smem)) return fail(PVT_ERR_UNSUPPORTED, "no k_ncc_search plan fits this template / window size");
    plan_items(g, n_tracks, mtp, sm_count);
    const int32_t v[16] = {g.G, g.C, g.GB, g.bands, g.ctas_band, g.span, g.boxW, g.boxH, g.pj, g.pd, g.cpt, g.n_full, g.n_tail, g.tail_ps,
                           (int32_t)smem, (int32_t)((Wmax > 8 * g.C ? 1 : 0) | (Hmax > kCY * g.G ? 2 : 0))};
    std::memcpy(out, v, sizeof(v));
    return PVT_OK;
}
int pvt_tc_plan_query(int sm_count, int n_tracks, int templ_w, int templ_h, int frame_w, int frame_h, int radius_x, int radius_y, int whole_frame_pass,
                      int32_t out[8])
{
    if (!out || sm_count <= 0 || n_tracks <= 0 || templ_w <= 0 || templ_h <= 0 || templ_w > frame_w || templ_h > frame_h || radius_x < 0 || radius_y < 0)
        return fail(PVT_ERR_INVALID, "bad plan query");
    Ctx d{};
    d.W = frame_w; d.H = frame_h; d.mtw = templ_w; d.mth = templ_h; d.max_tracks = n_tracks;
    d.global_pass = whole_frame_pass ? 1 : 0;
    d.Wmax = whole_frame_pass ? frame_w : std::min(2 * radius_x + 1, frame_w);     // as pvt_create / build_global_pass size the passes
    d.Hmax = whole_frame_pass ? frame_h : std::min(2 * radius_y + 1, frame_h);
    TcCfg g{};
    size_t smem = 0;
    { int r = tc_shape(d, sm_count, g, &smem); if (r) return r; }
    const int32_t v[8] = {g.XW, g.xtiles, g.mtiles, g.KS, g.AG, g.tmem_cols, g.stages, (int32_t)smem};
    std::memcpy(out, v, sizeof(v));
    return PVT_OK;
}
const char* pvt_last_error(void) { return g_err.c_str(); }

int pvt_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(PVT_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return n;
}

int pvt_device_info(int device, int* sm_count, int* sm_clock_khz, int* mem_clock_khz, size_t* mem_bytes, int* cc)
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, device));
    int clk = 0, mclk = 0;
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, device));
    CK(cudaDeviceGetAttribute(&mclk, cudaDevAttrMemoryClockRate, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (sm_clock_khz) *sm_clock_khz = clk;
    if (mem_clock_khz) *mem_clock_khz = mclk;
    if (mem_bytes) *mem_bytes = p.totalGlobalMem;
    if (cc) *cc = p.major * 10 + p.minor;
    return PVT_OK;
}

void pvt_default_params(pvt_params* p)
{
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->search_radius_x = 80;           // main.cpp:13
    p->search_radius_y = 80;           // main.cpp:14
    p->ncc_min_confidence = 0.40;      // main.cpp:17
    p->ncc_strong_confidence = 0.70;   // main.cpp:18
    p->template_update_lr = 0.10;      // main.cpp:19
    p->batch_size = 4;                 // main.cpp:11
    p->mode = PVT_MODE_NAIVE;          // main.cpp:8
    p->kernel = PVT_KERNEL_AUTO;
    p->lost_frame_threshold = 0;       // tracker/src/main.cpp has no lost-object logic
    p->ncc_global_confidence = 0.60;   // tracker_ghc/src/main.cpp:17
}

void pvt_default_params_ghc(pvt_params* p)
{
    if (!p) return;
    pvt_default_params(p);
    p->search_radius_x = 60;           // tracker_ghc/src/main.cpp:9
    p->search_radius_y = 60;           // :10
    p->ncc_min_confidence = 0.40;      // :15
    p->ncc_global_confidence = 0.60;   // :17
    p->ncc_strong_confidence = 0.70;   // :19
    p->template_update_lr = 0.10;      // :21
    p->lost_frame_threshold = 50;      // :23
}

int pvt_alloc_pinned(void** out, size_t bytes)
{
    if (!out) return fail(PVT_ERR_INVALID, "out is NULL");
    CK(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return PVT_OK;
}
int pvt_free_pinned(void* p)
{
    CK(cudaFreeHost(p));
    return PVT_OK;
}

int pvt_destroy(pvt_ctx* c)
{
    if (!c) return PVT_OK;
    cudaSetDevice(c->cfg.device);
    if (c->compute) cudaStreamSynchronize(c->compute);
    if (c->copy) cudaStreamSynchronize(c->copy);
    if (c->graph) cudaGraphExecDestroy(c->graph);
    if (c->graph_hold) cudaGraphExecDestroy(c->graph_hold);
    if (c->graph_prof) cudaGraphExecDestroy(c->graph_prof);
    if (c->graph_multi) cudaGraphExecDestroy(c->graph_multi);
    if (c->graph_long) cudaGraphExecDestroy(c->graph_long);
    if (c->graph_long_pf) cudaGraphExecDestroy(c->graph_long_pf);
    if (c->graph_pf) cudaGraphExecDestroy(c->graph_pf);
    if (c->graph_multi_pf) cudaGraphExecDestroy(c->graph_multi_pf);
    if (c->graph_global) cudaGraphExecDestroy(c->graph_global);
    for (int k = 0; k < 5; ++k) { if (c->pev[k][0]) cudaEventDestroy(c->pev[k][0]); if (c->pev[k][1]) cudaEventDestroy(c->pev[k][1]); }
    for (void* p : c->allocs) cudaFree(p);
    if (c->h_table) cudaFreeHost(c->h_table);
    for (int a = 0; a < 2; ++a)
        for (int n = 0; n <= kMaxGraphSteps; ++n)
            if (c->graph_n[a][n]) cudaGraphExecDestroy(c->graph_n[a][n]);
    if (c->h_results) cudaFreeHost(c->h_results);
    if (c->h_fault) cudaFreeHost(c->h_fault);
    for (int i = 0; i < kRing; ++i) if (c->table_ev[i]) cudaEventDestroy(c->table_ev[i]);
    for (int i = 0; i < kStageDepth; ++i) {
        if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
        if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
    }
    for (auto& p : c->ev_pool) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (int k = 0; k < 2; ++k) { if (c->rb_ready[k]) cudaEventDestroy(c->rb_ready[k]); if (c->rb_done[k]) cudaEventDestroy(c->rb_done[k]); }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_join2) cudaEventDestroy(c->ev_join2);
    if (c->aux) cudaStreamDestroy(c->aux);
    if (c->aux2) cudaStreamDestroy(c->aux2);
    if (c->aux3) cudaStreamDestroy(c->aux3);
    if (c->ev_join3) cudaEventDestroy(c->ev_join3);
    if (c->timer_a) cudaEventDestroy(c->timer_a);
    if (c->timer_b) cudaEventDestroy(c->timer_b);
    if (c->compute) cudaStreamDestroy(c->compute);
    if (c->copy) cudaStreamDestroy(c->copy);
    delete c;
    return PVT_OK;
}

int pvt_create(pvt_ctx** out, const pvt_params* params, const pvt_config* cfg)
{
    if (!out || !cfg) return fail(PVT_ERR_INVALID, "out/config is NULL");
    *out = nullptr;
    int r = validate_params(params);
    if (r) return r;
    if (cfg->frame_w <= 0 || cfg->frame_h <= 0 || cfg->max_streams <= 0 || cfg->max_tracks <= 0)
        return fail(PVT_ERR_INVALID, "frame geometry / stream / track counts must be positive");
    if (cfg->max_templ_w <= 0 || cfg->max_templ_h <= 0 || cfg->max_templ_w > cfg->frame_w || cfg->max_templ_h > cfg->frame_h)
        return fail(PVT_ERR_INVALID, "template larger than the frame (ncc_cpu.cpp:9-10)");
    if (cfg->max_streams > 65535 || cfg->max_tracks > 65535) return fail(PVT_ERR_INVALID, "at most 65535 streams / tracks per context");
    if (cfg->frame_h > 65535 || cfg->frame_w > (1 << 20)) return fail(PVT_ERR_INVALID, "frame larger than 1048576 x 65535");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(PVT_ERR_CUDA, std::string("no usable CUDA device (libpvt has no CPU fallback): ") + cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(PVT_ERR_INVALID, "device ordinal out of range");
    CK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) return fail(PVT_ERR_CUDA, "libpvt is built for sm_100a (B200) only; device is sm_" + std::to_string(prop.major * 10 + prop.minor));

    pvt_ctx* c = new pvt_ctx();
    c->params = *params;
    c->cfg = *cfg;
    if (c->cfg.max_radius_x <= 0) c->cfg.max_radius_x = params->search_radius_x;
    if (c->cfg.max_radius_y <= 0) c->cfg.max_radius_y = params->search_radius_y;
    if (params->search_radius_x > c->cfg.max_radius_x || params->search_radius_y > c->cfg.max_radius_y) {
        delete c;
        return fail(PVT_ERR_INVALID, "search radius exceeds config.max_radius");
    }
    c->lost_mode = params->lost_frame_threshold > 0;
    Ctx& d = c->d;
    d.lost_mode = c->lost_mode ? 1 : 0;
    d.formula = params->formula;
    d.global_pass = 0;
    d.W = cfg->frame_w;
    d.H = cfg->frame_h;
    d.pitch = (d.W + 3) & ~3;
    d.plane = ((size_t)d.pitch * d.H + 63) & ~(size_t)63;
    d.max_streams = cfg->max_streams;
    d.max_tracks = cfg->max_tracks;
    d.mtw = cfg->max_templ_w;
    d.mth = cfg->max_templ_h;
    d.mtp = (d.mtw + 7) & ~7;
    // window maxima: main.cpp:143-146 gives at most 2R+1 positions per axis, and never more than the map
    d.Wmax = std::min(2 * c->cfg.max_radius_x + 1, d.W);
    d.Hmax = std::min(2 * c->cfg.max_radius_y + 1, d.H);
    d.VW = (d.Wmax + d.mtw - 1 + 7) & ~7;

#define CR(x)                  \
    do {                       \
        int r_ = (x);          \
        if (r_) {              \
            pvt_destroy(c);    \
            return r_;         \
        }                      \
    } while (0)
#define CKD(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            pvt_destroy(c);                                                                           \
            return fail(PVT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
        }                                                                                             \
    } while (0)

    // kernel nodes inherit the priority of the stream they were captured on: the branches that carry the fringe kernel
    // (and, in the latency shape, the statistics) yield to the search kernel on the main stream
    int prio_lo = 0, prio_hi = 0;
    CKD(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CKD(cudaStreamCreateWithPriority(&c->compute, cudaStreamNonBlocking, prio_hi));
    CKD(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
    CKD(cudaStreamCreateWithPriority(&c->aux, cudaStreamNonBlocking, prio_hi));
    CKD(cudaStreamCreateWithPriority(&c->aux2, cudaStreamNonBlocking, prio_lo));
    CKD(cudaStreamCreateWithPriority(&c->aux3, cudaStreamNonBlocking, prio_lo));
    CKD(cudaEventCreateWithFlags(&c->ev_join3, cudaEventDisableTiming));
    CKD(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CKD(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CKD(cudaEventCreateWithFlags(&c->ev_join2, cudaEventDisableTiming));
    const size_t win = (size_t)d.Wmax * d.Hmax;
    CR(dev_alloc(c, &d.gray, d.plane * d.max_streams));
    CR(dev_alloc(c, &d.templ, (size_t)d.max_tracks * d.mth * d.mtw));
    CR(dev_alloc(c, &d.templc, (size_t)d.max_tracks * d.mth * d.mtp));
    if (stats_legacy(d.mtw)) {
        CR(dev_alloc(c, &d.vsum, (size_t)d.max_tracks * (d.Hmax + d.mth) * d.VW, false));   // column prefix sums, tileH + 1 rows
        CR(dev_alloc(c, &d.vsq, (size_t)d.max_tracks * (d.Hmax + d.mth) * d.VW, false));
    }
    CR(dev_alloc(c, &d.denom, (size_t)d.max_tracks * win, false));
    if (params->keep_maps) CR(dev_alloc(c, &d.maps, (size_t)d.max_tracks * win));
    CR(dev_alloc(c, &d.tracks, (size_t)d.max_tracks));
    CR(dev_alloc(c, &d.table, (size_t)kRing * d.max_streams));
    CR(dev_alloc(c, &d.seq, 1));
    CR(dev_alloc(c, &d.results, (size_t)kRing * d.max_tracks));
    CR(dev_alloc(c, &d.params, 1));
    CR(dev_alloc(c, &d.step, 1));
    CR(dev_alloc(c, &d.ticket, 1));
    CR(dev_alloc(c, &d.macs, 2));
    d.macs_grid = d.macs + 1;
    c->d_macs = d.macs;
    CR(dev_alloc(c, &c->d_trace, (size_t)kRing * 16));
    CKD(cudaHostAlloc((void**)&c->h_fault, 64, cudaHostAllocMapped));
    *c->h_fault = 0u;
    CKD(cudaHostGetDevicePointer((void**)&d.fault, c->h_fault, 0));
    CKD(cudaHostAlloc((void**)&c->h_table, sizeof(FrameDesc) * kRing * d.max_streams, cudaHostAllocDefault));
    CKD(cudaHostAlloc((void**)&c->h_results, sizeof(pvt_result) * kRing * d.max_tracks, cudaHostAllocDefault));
    for (int i = 0; i < kRing; ++i) CKD(cudaEventCreateWithFlags(&c->table_ev[i], cudaEventDisableTiming));
    for (int i = 0; i < kStageDepth; ++i) {
        CKD(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
        CKD(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
    }
    for (int k = 0; k < 5; ++k) { CKD(cudaEventCreate(&c->pev[k][0])); CKD(cudaEventCreate(&c->pev[k][1])); }
    for (int k = 0; k < 2; ++k) {
        CKD(cudaEventCreateWithFlags(&c->rb_ready[k], cudaEventDisableTiming));
        CKD(cudaEventCreateWithFlags(&c->rb_done[k], cudaEventDisableTiming));
    }
    CKD(cudaEventCreate(&c->timer_a));
    CKD(cudaEventCreate(&c->timer_b));
    c->stage.assign((size_t)d.max_streams * kStageDepth, nullptr);
    c->stage_bytes = (size_t)d.W * d.H * 4;
    c->track_stream.assign(d.max_tracks, -1);
    c->plane_full.assign(d.max_streams, 0);
    {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrCanUseHostPointerForRegisteredMem, cfg->device) == cudaSuccess) c->host_ptr_is_dev = v != 0;
        else cudaGetLastError();
        if (const char* e = getenv("PVT_DEBUG_PREFETCH_DELAY_US")) c->prefetch_delay_ns = std::max(0, atoi(e)) * 1000;
    }

    {
        Pass lp;
        lp.d = c->d;
        CR(build_plan(c, lp, prop.multiProcessorCount, params->ingest, true));
        c->d = lp.d; c->tile = lp.tile; c->tmap = lp.tmap; c->ncc_smem = lp.ncc_smem; c->rowsum_warps = lp.rowsum_warps;
        c->rowsum_pw = lp.rowsum_pw; c->colprefix_chunks = lp.colprefix_chunks; c->stat = lp.stat; c->fused = lp.fused; c->fused_smem = lp.fused_smem; c->local = lp.local; c->local_smem = lp.local_smem; c->fringe = lp.fringe; c->fringe_smem = lp.fringe_smem;
        c->roi_ingest = lp.roi_ingest;
    }
    c->kps = kernels_per_step(c);
    if (c->roi_ingest && !c->lost_mode && d.max_tracks <= 8) {
        // staging buffers of k_prefetch_roi: the search tile grown by the search radius on every side
        d.stage_w = (d.Wmax + c->cfg.max_radius_x + d.mtw + 8 + 3) & ~3;
        d.stage_h = d.Hmax + c->cfg.max_radius_y + d.mth;
        CR(dev_alloc(c, &d.stage, (size_t)d.max_tracks * d.stage_w * d.stage_h));
        CR(dev_alloc(c, &d.stage_hdr, (size_t)d.max_tracks));
    }
    if (wants_tc(params->kernel)) CR(setup_tc(c, prop.multiProcessorCount));
    c->templ_smem = (size_t)d.mth * d.mtw * sizeof(float);
    if (c->templ_smem > 48u * 1024u) {
        CR(raise_smem((const void*)k_update, c->templ_smem));
        CR(raise_smem((const void*)k_ncc_finalize, c->templ_smem));
        CR(raise_smem((const void*)k_track_init, c->templ_smem));
        CR(raise_smem((const void*)k_track_refresh, c->templ_smem));
        CR(raise_smem((const void*)k_track_digits, c->templ_smem));
    }
    if (c->lost_mode) CR(build_global_pass(c, prop.multiProcessorCount));
    CR(upload_params(c));
    CR(upload_seq(c, 0, kRing, 0));
    CKD(cudaDeviceSynchronize());  // the zero-fills above ran on the default stream; kernels use non-blocking streams
#undef CR
#undef CKD
    *out = c;
    return PVT_OK;
}

int pvt_set_params(pvt_ctx* c, const pvt_params* p)
{
    if (!c) return fail(PVT_ERR_INVALID, "ctx is NULL");
    int r = validate_params(p);
    if (r) return r;
    if (p->search_radius_x > c->cfg.max_radius_x || p->search_radius_y > c->cfg.max_radius_y)
        return fail(PVT_ERR_INVALID, "search radius exceeds the maxima the context was created with");
    if (p->keep_maps && !c->d.maps) return fail(PVT_ERR_INVALID, "keep_maps must be set at pvt_create");
    if ((p->lost_frame_threshold > 0) != c->lost_mode) return fail(PVT_ERR_INVALID, "lost-object mode (lost_frame_threshold > 0) must be chosen at pvt_create");
    if (p->formula != c->params.formula) return fail(PVT_ERR_INVALID, "the score formula must be chosen at pvt_create (templates carry its statistics)");
    if (wants_tc(p->kernel) && !c->tc_ready) return fail(PVT_ERR_INVALID, "PVT_KERNEL_TC must be chosen at pvt_create (it allocates the 8-bit gray plane and the template digits)");
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->compute));
    const bool regraph = p->kernel != c->params.kernel || p->ingest != c->params.ingest;
    {
        const Ctx& d = c->d;
        const double tiles = (double)d.max_tracks * (d.Wmax + d.mtw) * (d.Hmax + d.mth), frames_px = (double)d.max_streams * d.W * d.H;
        c->roi_ingest = p->ingest == PVT_INGEST_ROI || (p->ingest == PVT_INGEST_AUTO && tiles <= 0.5 * frames_px);
    }
    if (p->mode != c->params.mode || p->batch_size != c->params.batch_size) c->hold_pending = 0;   // a new cadence starts from a full batch
    if (c->params.kernel == PVT_KERNEL_TC_GLOBAL && p->kernel != PVT_KERNEL_TC_GLOBAL) {
        // the FP32 local pass of PVT_KERNEL_TC_GLOBAL kept no template digits (k_track_digits derives them per whole-frame pass):
        // a local tensor-core search needs them for every track from its first step -- re-derived whenever the context leaves this
        // kernel choice (also towards AUTO: a later switch to PVT_KERNEL_TC must not meet digits of an older template)
        for (int t = 0; t < c->d.max_tracks; ++t)
            if (c->track_stream[t] >= 0) k_track_refresh<<<1, 256, c->templ_smem, c->compute>>>(c->d, t);
        CK(cudaGetLastError());
    }
    c->params = *p;
    c->kps = kernels_per_step(c);
    if (regraph) c->graph_valid = false;
    return upload_params(c);
}

// a bounded device-side wait gave up (k_step_fused could not see all its CTAs / the statistics arrive): report it, once
#define FAULT_CHECK(c)                                                                                                        \
    do {                                                                                                                      \
        if ((c)->h_fault && *(volatile unsigned int*)(c)->h_fault) {                                                           \
            *(c)->h_fault = 0u;                                                                                               \
            return fail(PVT_ERR_CUDA, "k_step_fused: an in-kernel wait timed out (grid not co-scheduled); results of this step are invalid"); \
        }                                                                                                                     \
    } while (0)

int pvt_sync(pvt_ctx* c)
{
    if (!c) return fail(PVT_ERR_INVALID, "ctx is NULL");
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->copy));
    CK(cudaStreamSynchronize(c->compute));
    FAULT_CHECK(c);
    return PVT_OK;
}

// ingest one frame into its stream's gray plane right now (used by track_init / to_gray / map API)
static int ingest_now(pvt_ctx* c, const pvt_frame* f)
{
    int r = pvt_sync(c);
    if (r) return r;
    r = check_frame(c, f);
    if (r) return r;
    const Ctx& d = c->d;
    // a private table row at the CURRENT device step, then k_ingest alone; the step counter is untouched
    if (!c->seq_default) { int r2 = upload_seq(c, 0, kRing, 0); if (r2) return r2; }
    const int slot = (int)(c->submitted % kRing);
    std::vector<FrameDesc> row(d.max_streams, FrameDesc{nullptr, 0, 0, 0});
    FrameDesc fd{f->data, (unsigned long long)f->step, f->format, 1};
    c->seq_rows_valid = false;
    if (f->memory != PVT_MEM_DEVICE) {
        const size_t rb = frame_row_bytes(c, f->format);
        // the context's own staging slot of this stream (allocated once; nothing is in flight after the pvt_sync above) --
        // the reference mallocs and frees per call (baseline_kernel.cu:340-358), this path does not
        void*& st = c->stage[(size_t)f->stream * kStageDepth];
        if (!st) {
            CK(cudaMalloc(&st, c->stage_bytes));
            c->allocs.push_back(st);
        }
        // stream-ordered on the compute stream: a synchronous cudaMemcpy from pageable memory may return
        // while its DMA is still in flight, and the non-blocking compute stream would not wait for it
        CK(cudaMemcpy2DAsync(st, rb, f->data, f->step, rb, d.H, cudaMemcpyHostToDevice, c->compute));
        fd.data = st;
        fd.step = rb;
    }
    row[f->stream] = fd;
    cudaError_t e = cudaMemcpyAsync(d.table + (size_t)slot * d.max_streams, row.data(), sizeof(FrameDesc) * d.max_streams,
                                    cudaMemcpyHostToDevice, c->compute);
    if (e == cudaSuccess) {
        k_ingest<<<ingest_grid(d), kIngestThreads, 0, c->compute>>>(d);
        c->launches += 1;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->compute);
    }
    if (e != cudaSuccess) return fail(PVT_ERR_CUDA, std::string("ingest: ") + cudaGetErrorString(e));
    c->plane_full[f->stream] = 1;
    return PVT_OK;
}

int pvt_track_init(pvt_ctx* c, int track, int stream, const pvt_frame* frame0, int x, int y, int w, int h)
{
    if (!c) return fail(PVT_ERR_INVALID, "ctx is NULL");
    if (track < 0 || track >= c->cfg.max_tracks) return fail(PVT_ERR_INVALID, "track out of range");
    if (stream < 0 || stream >= c->cfg.max_streams) return fail(PVT_ERR_INVALID, "stream out of range");
    if (w <= 0 || h <= 0) return fail(PVT_ERR_INVALID, "empty ROI (main.cpp:66-69)");
    if (w > c->cfg.max_templ_w || h > c->cfg.max_templ_h) return fail(PVT_ERR_INVALID, "ROI larger than config.max_templ");
    if (x < 0 || y < 0 || x + w > c->cfg.frame_w || y + h > c->cfg.frame_h) return fail(PVT_ERR_INVALID, "ROI outside the frame");
    CK(cudaSetDevice(c->cfg.device));
    if (frame0) {
        if (frame0->stream != stream) return fail(PVT_ERR_INVALID, "frame0.stream != stream");
        int r = ingest_now(c, frame0);
        if (r) return r;
    } else {
        // utils.hpp:5-14 / main.cpp:70-71 cut the template from a COMPLETELY converted frame.  A context on the ROI ingest
        // only refreshes its tracks' search tiles, so after its first step the plane is a patchwork of frames.
        if (!c->plane_full[stream])
            return fail(PVT_ERR_STATE, "pvt_track_init(frame0 = NULL): this stream's plane holds no complete frame (the context uses the "
                                       "ROI ingest, which refreshes search tiles only); pass the frame, or create the context with PVT_INGEST_FULL");
        int r = pvt_sync(c);
        if (r) return r;
    }
    k_track_init<<<1, 256, c->templ_smem, c->compute>>>(c->d, track, stream, x, y, w, h);
    c->launches += 1;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->compute));
    c->track_stream[track] = stream;
    return PVT_OK;
}

int pvt_track_remove(pvt_ctx* c, int track)
{
    if (!c) return fail(PVT_ERR_INVALID, "ctx is NULL");
    if (track < 0 || track >= c->cfg.max_tracks) return fail(PVT_ERR_INVALID, "track out of range");
    CK(cudaSetDevice(c->cfg.device));
    int r = pvt_sync(c);
    if (r) return r;
    CK(cudaMemsetAsync(&c->d.tracks[track], 0, sizeof(TrackState), c->compute));
    CK(cudaStreamSynchronize(c->compute));
    c->track_stream[track] = -1;
    return PVT_OK;
}

int pvt_submit(pvt_ctx* c, int n_frames, const pvt_frame* frames)
{
    if (!c) return fail(PVT_ERR_INVALID, "ctx is NULL");
    if (n_frames < 0 || (n_frames > 0 && !frames)) return fail(PVT_ERR_INVALID, "frames is NULL");
    CK(cudaSetDevice(c->cfg.device));
    bool hold = false;
    const int hold_before = c->hold_pending;
    if (c->params.mode == PVT_MODE_BATCH && c->params.batch_size > 1) {
        // main.cpp:115-130: frames are collected until the batch is full; only then is one searched
        if (++c->hold_pending < c->params.batch_size) hold = true;
        else c->hold_pending = 0;
    }
    const int r = enqueue_step(c, n_frames, frames, hold);
    if (r) c->hold_pending = hold_before;   // a rejected call (bad frame, ...) does not shift the --batch=N cadence
    return r;
}

int pvt_collect(pvt_ctx* c, pvt_result* results, int max_steps)
{
    if (!c) return fail(PVT_ERR_INVALID, "ctx is NULL");
    CK(cudaSetDevice(c->cfg.device));
    int r = pvt_sync(c);
    if (r) return r;
    int n = (int)std::min<unsigned long long>(std::min<unsigned long long>(c->submitted, (unsigned long long)std::max(max_steps, 0)), kRing);
    if (!results || n == 0) return n;
    const int mt = c->cfg.max_tracks;
    CK(cudaMemcpy(c->h_results, c->d.results, sizeof(pvt_result) * kRing * mt, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
        const unsigned long long s = c->submitted - n + i;
        std::memcpy(results + (size_t)i * mt, c->h_results + (size_t)(s % kRing) * mt, sizeof(pvt_result) * mt);
    }
    return n;
}

int pvt_submit_sequence(pvt_ctx* c, int n_steps, int n_frames, const pvt_frame* frames, int ring_len, int collect_every,
                        pvt_result* results_out)
{
    if (!c) return fail(PVT_ERR_INVALID, "ctx is NULL");
    if (n_steps < 0 || n_frames <= 0 || ring_len <= 0 || !frames) return fail(PVT_ERR_INVALID, "bad sequence arguments");
    if (collect_every > kRing) return fail(PVT_ERR_INVALID, "collect_every exceeds the 64-step results ring");
    const int mt = c->cfg.max_tracks;
    // Resident ring fast path: every frame is device memory -- or PINNED host memory that the ROI ingest reads in place
    // over PCIe (zero-copy: each step's tiles cross the bus inside k_ingest_roi) -- and the ring fits the table -> upload
    // the ring's rows and a SeqDesc ONCE; after that a time step is a bare graph launch (no host-side bookkeeping per frame).
    bool resident = ring_len <= kRing && !c->profiling && !debug_sync() && n_steps > 0;
    const int ms = c->cfg.max_streams;
    // The ring's rows as the device will see them, built in a scratch vector the context keeps (no allocation per call).
    // Pointer kinds: DEVICE as is; HOST_PINNED is the caller's promise that the buffer is page-locked -- under unified
    // addressing its device alias is the pointer itself, no driver query; plain HOST is probed (cudaPointerGetAttributes).
    std::vector<FrameDesc>& rows = c->seq_scratch;
    bool any_host = false;
    if (resident) {
        rows.assign((size_t)ring_len * ms, FrameDesc{nullptr, 0, 0, 0});
        for (int k = 0; resident && k < ring_len; ++k)
            for (int i = 0; i < n_frames; ++i) {
                const pvt_frame* f = frames + (size_t)k * n_frames + i;
                int r = check_frame(c, f);
                if (r) return r;
                FrameDesc& slot = rows[(size_t)k * ms + f->stream];
                if (slot.valid) return fail(PVT_ERR_INVALID, "two frames for one stream in one step");
                const void* dp = f->data;
                if (f->memory != PVT_MEM_DEVICE) {
                    any_host = true;
                    if (!c->roi_ingest) { resident = false; break; }
                    if (!(f->memory == PVT_MEM_HOST_PINNED && c->host_ptr_is_dev)) {
                        cudaPointerAttributes pa{};
                        if (cudaPointerGetAttributes(&pa, f->data) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer) dp = pa.devicePointer;
                        else {
                            cudaGetLastError();
                            if (f->memory == PVT_MEM_HOST_PINNED) return fail(PVT_ERR_INVALID, "frame declared PVT_MEM_HOST_PINNED is not page-locked memory");
                            resident = false;
                            break;
                        }
                    }
                }
                slot = FrameDesc{dp, (unsigned long long)f->step, f->format, 1};
                c->plane_full[f->stream] = !c->roi_ingest;
            }
    }
    if (resident) {
        CK(cudaSetDevice(c->cfg.device));
        // Is this the ring the device table already holds, possibly rotated (callers walk a ring and pass it from its
        // current position)?  Then nothing is uploaded: the rotation goes into SeqDesc.phase.
        int phase = -1;
        const size_t row_bytes = sizeof(FrameDesc) * (size_t)ms;
        if (c->seq_rows_valid && c->seq_ring_len == ring_len) {
            for (int r0 = 0; r0 < ring_len && phase < 0; ++r0) {
                if (std::memcmp(rows.data(), c->seq_rows.data() + (size_t)r0 * ms, row_bytes) != 0) continue;
                bool same = true;
                for (int k = 1; k < ring_len && same; ++k)
                    same = std::memcmp(rows.data() + (size_t)k * ms, c->seq_rows.data() + (size_t)((k + r0) % ring_len) * ms, row_bytes) == 0;
                if (same) phase = r0;
            }
        }
        Readback pend[2] = {};
        bool had_prev = false;
        if (phase < 0) {
            for (int k = 0; k < kRing; ++k)   // earlier uploads from the pinned table must have left it
                if (c->table_ev_used[k]) { CK(cudaEventSynchronize(c->table_ev[k])); c->table_ev_used[k] = false; }
            std::memcpy(c->h_table, rows.data(), row_bytes * ring_len);
            CK(cudaMemcpyAsync(c->d.table, c->h_table, row_bytes * ring_len, cudaMemcpyHostToDevice, c->compute));
            CK(cudaEventRecord(c->table_ev[0], c->compute));
            c->table_ev_used[0] = true;
            c->seq_rows.swap(rows);
            c->seq_ring_len = ring_len;
            c->seq_rows_valid = true;
            phase = 0;
        }
        const bool pf = any_host && c->d.stage && !c->lost_mode;
        // one tiny kernel sets the sequence descriptor (passed by value: no staging copy, no pinned buffer to guard) and, for
        // pinned rings, drops whatever an earlier sequence staged (the caller may have refilled the buffers since)
        k_seq_begin<<<1, 64, 0, c->compute>>>(c->d, SeqDesc{c->submitted, ring_len, 0, pf ? 1 : 0, phase}, pf ? 1 : 0);
        c->seq_default = false;
        int r = PVT_OK;
        if (!c->graph_valid) { r = build_graphs(c); if (r) return r; }
        cudaGraphExec_t g_one = pf ? c->graph_pf : c->graph, g_multi = pf ? c->graph_multi_pf : c->graph_multi;
        cudaGraphExec_t g_long = pf ? c->graph_long_pf : c->graph_long;
        const bool batch = c->params.mode == PVT_MODE_BATCH && c->params.batch_size > 1;
        for (int s = 0; s < n_steps; ++s) {
            // several steps per launch while no result read-back falls inside the group
            auto fits = [&](cudaGraphExec_t ge, int n) {
                return !batch && ge && s + n <= n_steps && (collect_every <= 0 || (s % collect_every) + n <= collect_every);
            };
            int group = fits(g_long, kLongStep) ? kLongStep : fits(g_multi, kMultiStep) ? kMultiStep : 1;
            cudaGraphExec_t ge = group == kLongStep ? g_long : g_multi;
            static const bool lost_multi = [] { const char* e = getenv("PVT_LOST_MULTI"); return !(e && *e == '0'); }();
            if ((g_long || (c->lost_mode && lost_multi)) && !batch) {
                // latency shape: everything up to the next result read-back (at most 64 steps) in ONE launch
                int n = n_steps - s;
                if (collect_every > 0) n = std::min(n, collect_every - (s % collect_every));
                n = std::min(n, c->lost_mode ? 16 : kMaxGraphSteps);   // (every step of a lost-mode graph carries a whole-frame pass as the body of its IF node)
                if (n > 1) {
                    int r3 = steps_graph(c, n, pf, &ge);
                    if (r3) return r3;
                    group = n;
                }
            }
            if (group > 1) {
                CK(cudaGraphLaunch(ge, c->compute));
                c->launches += (int64_t)group * (c->kps + (pf ? 1 : 0) + (c->lost_mode ? c->kps_global : 0));
                c->submitted += group;
                s += group - 1;
                if (collect_every <= 0 || (s + 1) % collect_every != 0) continue;
            } else {
            bool hold = false;
            if (batch) { if (++c->hold_pending < c->params.batch_size) hold = true; else c->hold_pending = 0; }
            CK(cudaGraphLaunch(hold ? c->graph_hold : g_one, c->compute));
            c->launches += hold ? 1 : c->kps + (pf ? 1 : 0) + (c->lost_mode ? c->kps_global : 0);
            c->submitted += 1;
            }
            if (collect_every > 0 && (s + 1) % collect_every == 0) {
                // Device -> host read of these steps' results WITHOUT draining the pipeline: the copy runs on the copy stream
                // behind an event, the host goes on enqueuing the next batch and picks the rows up one batch later.
                // (Batches of at most half the 64-step results ring; larger ones are read back synchronously.)
                const unsigned long long first = c->submitted - collect_every;
                const bool lag = 2 * collect_every <= kRing;
                const int b = (int)(((s + 1) / collect_every - 1) & 1);
                if (lag && pend[b ^ 1].n) { int r2 = finish_readback(c, pend[b ^ 1], results_out, mt); if (r2) return r2; }
                CK(cudaEventRecord(c->rb_ready[b], c->compute));
                CK(cudaStreamWaitEvent(c->copy, c->rb_ready[b], 0));
                {   // the batch's slots are contiguous in the results ring up to one wrap-around: at most two copies
                    const int s0 = (int)(first % kRing), n0 = std::min(collect_every, kRing - s0);
                    CK(cudaMemcpyAsync(c->h_results + (size_t)s0 * mt, c->d.results + (size_t)s0 * mt, sizeof(pvt_result) * mt * n0,
                                       cudaMemcpyDeviceToHost, c->copy));
                    if (n0 < collect_every)
                        CK(cudaMemcpyAsync(c->h_results, c->d.results, sizeof(pvt_result) * mt * (collect_every - n0), cudaMemcpyDeviceToHost, c->copy));
                }
                CK(cudaEventRecord(c->rb_done[b], c->copy));
                // the steps that overwrite result slots (64 steps after the ones that filled them) must not start before the
                // slots have been copied: the NEXT batch waits for the PREVIOUS batch's copy (this batch's own when not lagging)
                if (!lag) CK(cudaStreamWaitEvent(c->compute, c->rb_done[b], 0));
                else if (had_prev) CK(cudaStreamWaitEvent(c->compute, c->rb_done[b ^ 1], 0));
                had_prev = true;
                pend[b] = Readback{first, collect_every, s + 1 - collect_every, b};
                if (!lag) { int r2 = finish_readback(c, pend[b], results_out, mt); if (r2) return r2; }
            }
        }
        for (int b = 0; b < 2; ++b)
            if (pend[b].n) { int r2 = finish_readback(c, pend[b], results_out, mt); if (r2) return r2; }
        return PVT_OK;
    }
    for (int s = 0; s < n_steps; ++s) {
        int r = pvt_submit(c, n_frames, frames + (size_t)(s % ring_len) * n_frames);
        if (r) return r;
        if (collect_every > 0 && (s + 1) % collect_every == 0) {
            // device -> host read of these steps' results (async copy, then wait for it: the e2e contract)
            const unsigned long long first = c->submitted - collect_every;
            for (int k = 0; k < collect_every; ++k) {
                const int slot = (int)((first + k) % kRing);
                CK(cudaMemcpyAsync(c->h_results + (size_t)slot * mt, c->d.results + (size_t)slot * mt, sizeof(pvt_result) * mt,
                                   cudaMemcpyDeviceToHost, c->compute));
            }
            CK(cudaStreamSynchronize(c->compute));
            if (results_out)
                for (int k = 0; k < collect_every; ++k)
                    std::memcpy(results_out + (size_t)(s + 1 - collect_every + k) * mt, c->h_results + (size_t)((first + k) % kRing) * mt,
                                sizeof(pvt_result) * mt);
        }
    }
    return PVT_OK;
}

int pvt_step(pvt_ctx* c, int n_frames, const pvt_frame* frames, pvt_result* results)
{
    int r = pvt_submit(c, n_frames, frames);
    if (r) return r;
    const int mt = c->cfg.max_tracks;
    const int slot = (int)((c->submitted - 1) % kRing);
    if (results) {
        CK(cudaMemcpyAsync(c->h_results + (size_t)slot * mt, c->d.results + (size_t)slot * mt, sizeof(pvt_result) * mt,
                           cudaMemcpyDeviceToHost, c->compute));
        CK(cudaStreamSynchronize(c->compute));
        std::memcpy(results, c->h_results + (size_t)slot * mt, sizeof(pvt_result) * mt);
    } else {
        CK(cudaStreamSynchronize(c->compute));
    }
    FAULT_CHECK(c);
    return PVT_OK;
}

int pvt_get_state(pvt_ctx* c, int track, int32_t bbox[4], float* templ, size_t templ_step_bytes)
{
    if (!c) return fail(PVT_ERR_INVALID, "ctx is NULL");
    if (track < 0 || track >= c->cfg.max_tracks) return fail(PVT_ERR_INVALID, "track out of range");
    CK(cudaSetDevice(c->cfg.device));
    int r = pvt_sync(c);
    if (r) return r;
    TrackState t;
    CK(cudaMemcpy(&t, &c->d.tracks[track], sizeof(t), cudaMemcpyDeviceToHost));
    if (!t.active) return fail(PVT_ERR_STATE, "track is not initialised");
    if (bbox) { bbox[0] = t.x; bbox[1] = t.y; bbox[2] = t.w; bbox[3] = t.h; }
    if (templ) {
        if (templ_step_bytes < (size_t)t.w * 4) return fail(PVT_ERR_INVALID, "templ_step_bytes too small");
        CK(cudaMemcpy2D(templ, templ_step_bytes, c->d.templ + (size_t)track * c->d.mth * c->d.mtw, (size_t)t.w * 4, (size_t)t.w * 4, t.h,
                        cudaMemcpyDeviceToHost));
    }
    return PVT_OK;
}

int pvt_set_state(pvt_ctx* c, int track, const int32_t bbox[4], const float* templ, size_t templ_step_bytes)
{
    if (!c) return fail(PVT_ERR_INVALID, "ctx is NULL");
    if (track < 0 || track >= c->cfg.max_tracks) return fail(PVT_ERR_INVALID, "track out of range");
    if (!bbox) return fail(PVT_ERR_INVALID, "bbox is NULL");
    CK(cudaSetDevice(c->cfg.device));
    int r = pvt_sync(c);
    if (r) return r;
    TrackState t;
    CK(cudaMemcpy(&t, &c->d.tracks[track], sizeof(t), cudaMemcpyDeviceToHost));
    if (!t.active && !templ) return fail(PVT_ERR_STATE, "track is not initialised and no template given");
    const int w = bbox[2], h = bbox[3];
    if (w <= 0 || h <= 0 || w > c->cfg.max_templ_w || h > c->cfg.max_templ_h) return fail(PVT_ERR_INVALID, "bad template size");
    if (bbox[0] < 0 || bbox[1] < 0 || bbox[0] + w > c->cfg.frame_w || bbox[1] + h > c->cfg.frame_h) return fail(PVT_ERR_INVALID, "bbox outside the frame");
    if (!templ && (w != t.w || h != t.h)) return fail(PVT_ERR_INVALID, "template size change needs a template");
    if (!t.active) { t.stream = std::max(c->track_stream[track], 0); c->track_stream[track] = t.stream; }
    t.active = 1; t.x = bbox[0]; t.y = bbox[1]; t.w = w; t.h = h; t.peak = 0ull;
    CK(cudaMemcpyAsync(&c->d.tracks[track], &t, sizeof(t), cudaMemcpyHostToDevice, c->compute));
    if (templ) {
        if (templ_step_bytes < (size_t)w * 4) return fail(PVT_ERR_INVALID, "templ_step_bytes too small");
        CK(cudaMemcpy2DAsync(c->d.templ + (size_t)track * c->d.mth * c->d.mtw, (size_t)w * 4, templ, templ_step_bytes, (size_t)w * 4, h,
                             cudaMemcpyHostToDevice, c->compute));
    }
    k_track_refresh<<<1, 256, c->templ_smem, c->compute>>>(c->d, track);
    c->launches += 1;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->compute));
    return PVT_OK;
}

int pvt_get_lost_state(pvt_ctx* c, int track, int* lost_frame_count, int* use_global_search)
{
    if (!c) return fail(PVT_ERR_INVALID, "ctx is NULL");
    if (track < 0 || track >= c->cfg.max_tracks) return fail(PVT_ERR_INVALID, "track out of range");
    CK(cudaSetDevice(c->cfg.device));
    int r = pvt_sync(c);
    if (r) return r;
    TrackState t;
    CK(cudaMemcpy(&t, &c->d.tracks[track], sizeof(t), cudaMemcpyDeviceToHost));
    if (!t.active) return fail(PVT_ERR_STATE, "track is not initialised");
    if (lost_frame_count) *lost_frame_count = t.lost_count;
    if (use_global_search) *use_global_search = t.use_global;
    return PVT_OK;
}

int pvt_set_lost_state(pvt_ctx* c, int track, int lost_frame_count, int use_global_search)
{
    if (!c) return fail(PVT_ERR_INVALID, "ctx is NULL");
    if (track < 0 || track >= c->cfg.max_tracks) return fail(PVT_ERR_INVALID, "track out of range");
    if (lost_frame_count < 0) return fail(PVT_ERR_INVALID, "negative lost_frame_count");
    if (use_global_search && !c->lost_mode) return fail(PVT_ERR_STATE, "context was created without lost-object mode");
    CK(cudaSetDevice(c->cfg.device));
    int r = pvt_sync(c);
    if (r) return r;
    TrackState t;
    CK(cudaMemcpy(&t, &c->d.tracks[track], sizeof(t), cudaMemcpyDeviceToHost));
    if (!t.active) return fail(PVT_ERR_STATE, "track is not initialised");
    t.lost_count = lost_frame_count;
    t.use_global = use_global_search ? 1 : 0;
    t.global_since = 0ull;   // applies from the next submitted step on
    CK(cudaMemcpy(&c->d.tracks[track], &t, sizeof(t), cudaMemcpyHostToDevice));
    return PVT_OK;
}

int pvt_get_window_map(pvt_ctx* c, int track, float* out, size_t out_step_bytes, int32_t win[4])
{
    if (!c) return fail(PVT_ERR_INVALID, "ctx is NULL");
    if (track < 0 || track >= c->cfg.max_tracks) return fail(PVT_ERR_INVALID, "track out of range");
    if (!c->d.maps) return fail(PVT_ERR_STATE, "context was created without params.keep_maps");
    CK(cudaSetDevice(c->cfg.device));
    int r = pvt_sync(c);
    if (r) return r;
    TrackState t;
    CK(cudaMemcpy(&t, &c->d.tracks[track], sizeof(t), cudaMemcpyDeviceToHost));
    if (!t.active || t.win[2] <= 0) return fail(PVT_ERR_STATE, "track has not been stepped yet");
    if (win) for (int i = 0; i < 4; ++i) win[i] = t.win[i];
    if (out) {
        if (out_step_bytes < (size_t)t.win[2] * 4) return fail(PVT_ERR_INVALID, "out_step_bytes too small");
        CK(cudaMemcpy2D(out, out_step_bytes, c->d.maps + (size_t)track * c->d.Hmax * c->d.Wmax, (size_t)t.win[2] * 4, (size_t)t.win[2] * 4,
                        t.win[3], cudaMemcpyDeviceToHost));
    }
    return PVT_OK;
}

int pvt_to_gray_f32(pvt_ctx* c, const pvt_frame* frame, float* out, size_t out_step_bytes, int out_memory)
{
    if (!c || !frame || !out) return fail(PVT_ERR_INVALID, "NULL argument");
    if (out_step_bytes < (size_t)c->cfg.frame_w * 4) return fail(PVT_ERR_INVALID, "out_step_bytes too small");
    CK(cudaSetDevice(c->cfg.device));
    int r = ingest_now(c, frame);
    if (r) return r;
    CK(cudaMemcpy2D(out, out_step_bytes, c->d.gray + (size_t)frame->stream * c->d.plane, (size_t)c->d.pitch * 4, (size_t)c->d.W * 4, c->d.H,
                    out_memory == PVT_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
    return PVT_OK;
}

// ---- map-level operators (baseline_kernel.hpp:8-17) ---------------------------------------------
namespace {
std::mutex g_map_mu;
std::map<std::tuple<int, int, int, int, int, int>, pvt_ctx*> g_map_ctx;

int map_ctx(int device, int formula, int fw, int fh, int tw, int th, pvt_ctx** out)
{
    auto key = std::make_tuple(device, formula, fw, fh, tw, th);
    auto it = g_map_ctx.find(key);
    if (it != g_map_ctx.end()) { *out = it->second; return PVT_OK; }
    pvt_params p;
    pvt_default_params(&p);
    p.search_radius_x = fw;  // window == the whole map
    p.search_radius_y = fh;
    p.keep_maps = 1;
    p.formula = formula;
    p.ncc_min_confidence = 2.0;      // a map operator has no tracker state: no score reaches 2, so the box never moves and the
    p.ncc_strong_confidence = 2.0;   // template is never blended -- one state upload serves every frame of a batch
    pvt_config cfg{};
    cfg.device = device; cfg.frame_w = fw; cfg.frame_h = fh; cfg.max_streams = 1; cfg.max_tracks = 1;
    cfg.max_templ_w = tw; cfg.max_templ_h = th;
    int r = pvt_create(out, &p, &cfg);
    if (r) return r;
    if (g_map_ctx.size() >= 4) {  // small cache: the reference allocates per call (baseline_kernel.cu:340-358)
        pvt_destroy(g_map_ctx.begin()->second);
        g_map_ctx.erase(g_map_ctx.begin());
    }
    g_map_ctx[key] = *out;
    return PVT_OK;
}
}  // namespace

int pvt_ncc_match_batched_f(int device, int formula, int n, const float* const* frames, int fw, int fh, size_t fstep_bytes, const float* templ,
                            int tw, int th, size_t tstep_bytes, float* const* outs, size_t ostep_bytes)
{
    if (formula != PVT_FORMULA_CCOEFF_NORMED && formula != PVT_FORMULA_EPS) return fail(PVT_ERR_INVALID, "unknown formula");
    if (n <= 0 || !frames || !templ || !outs) return fail(PVT_ERR_INVALID, "empty batch / NULL argument (baseline_kernel.cu:412)");
    if (tw <= 0 || th <= 0 || fw < tw || fh < th) return fail(PVT_ERR_INVALID, "frame smaller than template (ncc_cpu.cpp:9-10)");
    const int outW = fw - tw + 1, outH = fh - th + 1;
    if (fstep_bytes < (size_t)fw * 4 || tstep_bytes < (size_t)tw * 4 || ostep_bytes < (size_t)outW * 4) return fail(PVT_ERR_INVALID, "step too small");
    std::lock_guard<std::mutex> lk(g_map_mu);
    pvt_ctx* c = nullptr;
    int r = map_ctx(device, formula, fw, fh, tw, th, &c);
    if (r) return r;
    for (int i = 0; i < n; ++i)
        if (!frames[i] || !outs[i]) return fail(PVT_ERR_INVALID, "NULL frame / output in batch");
    // every map is taken against the SAME template (baseline_kernel.cu:462-466), and the map context never moves its box
    // or blends its template (map_ctx), so the state goes up once; per frame: staged H2D, one graph launch, one D2H of the map
    const int32_t bbox[4] = {0, 0, tw, th};
    c->track_stream[0] = 0;
    r = pvt_set_state(c, 0, bbox, templ, tstep_bytes);
    if (r) return r;
    for (int i = 0; i < n; ++i) {
        pvt_frame f{0, PVT_FMT_GRAYF32, PVT_MEM_HOST, 0, frames[i], fstep_bytes};
        r = pvt_submit(c, 1, &f);
        if (r) return r;
        // stream-ordered behind the step on the compute stream; the map of a 1-track context starts at c->d.maps
        CK(cudaMemcpy2DAsync(outs[i], ostep_bytes, c->d.maps, (size_t)outW * 4, (size_t)outW * 4, outH, cudaMemcpyDeviceToHost, c->compute));
        CK(cudaStreamSynchronize(c->compute));   // pageable destination: the copy has landed when this returns
    }
    return PVT_OK;
}

int pvt_ncc_match_batched(int device, int n, const float* const* frames, int fw, int fh, size_t fstep_bytes, const float* templ, int tw,
                          int th, size_t tstep_bytes, float* const* outs, size_t ostep_bytes)
{
    return pvt_ncc_match_batched_f(device, PVT_FORMULA_CCOEFF_NORMED, n, frames, fw, fh, fstep_bytes, templ, tw, th, tstep_bytes, outs, ostep_bytes);
}

int pvt_ncc_match(int device, int mode, const float* frame, int fw, int fh, size_t fstep_bytes, const float* templ, int tw, int th,
                  size_t tstep_bytes, float* out, size_t ostep_bytes)
{
    const int formula = (mode & PVT_MODE_FLAG_EPS) ? PVT_FORMULA_EPS : PVT_FORMULA_CCOEFF_NORMED;
    mode &= ~PVT_MODE_FLAG_EPS;
    if (mode == PVT_MODE_CPU) return fail(PVT_ERR_UNSUPPORTED, "PVT_MODE_CPU: libpvt has no CPU path");
    if (mode < PVT_MODE_NAIVE || mode > PVT_MODE_BATCH) return fail(PVT_ERR_INVALID, "unknown mode");
    return pvt_ncc_match_batched_f(device, formula, 1, &frame, fw, fh, fstep_bytes, templ, tw, th, tstep_bytes, &out, ostep_bytes);
}

// ---- measurement hooks ---------------------------------------------------------------------------
int pvt_draw_boxes(pvt_ctx* c, const pvt_frame* f, int n, const int32_t* boxes, const uint8_t* bgr)
{
    if (!c || !f) return fail(PVT_ERR_INVALID, "ctx / frame is NULL");
    if (n < 0 || (n > 0 && !boxes)) return fail(PVT_ERR_INVALID, "boxes is NULL");
    if (f->format != PVT_FMT_BGR8) return fail(PVT_ERR_INVALID, "pvt_draw_boxes paints BGR8 frames (main.cpp:166 draws on the decoded frame)");
    if (!f->data || f->step < (size_t)c->cfg.frame_w * 3) return fail(PVT_ERR_INVALID, "frame.data is NULL or frame.step smaller than one row");
    if (n == 0) return PVT_OK;
    const int W = c->cfg.frame_w, H = c->cfg.frame_h;
    // the tracker's boxes always lie inside the frame (they come from a map position); OpenCV clips lines that leave the image
    // before it thickens them, which changes their caps -- that case is not restated, so it is rejected rather than painted differently
    for (int i = 0; i < n; ++i) {
        const int32_t* b = boxes + 4 * i;
        if (b[2] <= 0 || b[3] <= 0 || b[0] < 0 || b[1] < 0 || b[0] + b[2] > W || b[1] + b[3] > H)
            return fail(PVT_ERR_INVALID, "pvt_draw_boxes: box " + std::to_string(i) + " is empty or leaves the frame");
    }
    CK(cudaSetDevice(c->cfg.device));
    { int r = pvt_sync(c); if (r) return r; }
    int* d_boxes = nullptr;
    CK(cudaMalloc(&d_boxes, sizeof(int) * 4 * (size_t)n));
    cudaError_t e = cudaMemcpyAsync(d_boxes, boxes, sizeof(int) * 4 * (size_t)n, cudaMemcpyHostToDevice, c->compute);
    unsigned char* img = (unsigned char*)f->data;
    size_t step = f->step;
    void* stage = nullptr;
    if (e == cudaSuccess && f->memory != PVT_MEM_DEVICE) {   // host frame: round trip through a staging buffer
        e = cudaMalloc(&stage, (size_t)W * 3 * H);
        if (e == cudaSuccess) e = cudaMemcpy2DAsync(stage, (size_t)W * 3, f->data, f->step, (size_t)W * 3, H, cudaMemcpyHostToDevice, c->compute);
        img = (unsigned char*)stage; step = (size_t)W * 3;
    }
    if (e == cudaSuccess) {
        const int b = bgr ? bgr[0] : 0, g = bgr ? bgr[1] : 255, r = bgr ? bgr[2] : 0;
        k_overlay<<<dim3(8, (unsigned)n), 256, 0, c->compute>>>(img, step, W, H, d_boxes, n, b, g, r);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && stage) e = cudaMemcpy2DAsync((void*)f->data, f->step, stage, (size_t)W * 3, (size_t)W * 3, H, cudaMemcpyDeviceToHost, c->compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->compute);
    cudaFree(d_boxes);
    if (stage) cudaFree(stage);
    if (e != cudaSuccess) return fail(PVT_ERR_CUDA, std::string("pvt_draw_boxes: ") + cudaGetErrorString(e));
    return PVT_OK;
}

int pvt_profile_enable(pvt_ctx* c, int on)
{
    if (!c) return fail(PVT_ERR_INVALID, "ctx is NULL");
    CK(cudaSetDevice(c->cfg.device));
    int r = resolve_profile(c);
    if (r) return r;
    c->profiling = on != 0;
    return PVT_OK;
}

int pvt_profile_get(pvt_ctx* c, pvt_profile* out, int reset)
{
    if (!c || !out) return fail(PVT_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->cfg.device));
    int r = resolve_profile(c);
    if (r) return r;
    unsigned long long macs[2] = {0, 0};
    CK(cudaMemcpy(macs, c->d_macs, sizeof(macs), cudaMemcpyDeviceToHost));
    c->prof.ncc_macs = (double)macs[0];
    c->prof.search_kernel_macs = (double)macs[1];
    *out = c->prof;
    if (reset) {
        c->prof = pvt_profile{};
        CK(cudaMemsetAsync(c->d_macs, 0, 2 * sizeof(unsigned long long), c->compute));
        CK(cudaStreamSynchronize(c->compute));
    }
    return PVT_OK;
}

int pvt_trace_enable(pvt_ctx* c, int on)
{
    if (!c) return fail(PVT_ERR_INVALID, "ctx is NULL");
    CK(cudaSetDevice(c->cfg.device));
    int r = pvt_sync(c);
    if (r) return r;
    c->d.trace = on ? c->d_trace : nullptr;
    c->graph_valid = false;  // kernel arguments are baked into the graph
    CK(cudaMemsetAsync(c->d_trace, 0, sizeof(unsigned long long) * kRing * 16, c->compute));
    CK(cudaStreamSynchronize(c->compute));
    return PVT_OK;
}

int pvt_trace_get(pvt_ctx* c, uint64_t* out, int max_steps)
{
    if (!c || !out) return fail(PVT_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->cfg.device));
    int r = pvt_sync(c);
    if (r) return r;
    const int n = (int)std::min<unsigned long long>(std::min<unsigned long long>(c->submitted, (unsigned long long)std::max(max_steps, 0)), kRing);
    std::vector<unsigned long long> h((size_t)kRing * 16);
    CK(cudaMemcpy(h.data(), c->d_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
        const unsigned long long s = c->submitted - n + i;
        std::memcpy(out + (size_t)i * 16, h.data() + (size_t)(s % kRing) * 16, 16 * sizeof(unsigned long long));
    }
    return n;
}

int64_t pvt_launch_count(pvt_ctx* c) { return c ? c->launches : 0; }

int pvt_search_kind(pvt_ctx* c, char* name, int name_bytes)
{
    if (!c) return fail(PVT_ERR_INVALID, "ctx is NULL");
    const char* k = "k_ncc_search";
    if (c->params.kernel == PVT_KERNEL_TC) k = "k_ncc_tc";
    else if (c->params.kernel == PVT_KERNEL_DIRECT) k = "k_ncc_direct";
    else if (c->local.TR > 0 && local_kernel(c) == PVT_KERNEL_AUTO) k = "k_ncc_local";
    else if (c->fused) k = "k_step_fused";
    else if (c->tile.pj * c->tile.pd > 1) k = "k_ncc_search+k_ncc_finalize";
    if (name && name_bytes > 0) { std::strncpy(name, k, (size_t)name_bytes - 1); name[name_bytes - 1] = 0; }
    return c->kps;
}

int pvt_timer_start(pvt_ctx* c)
{
    if (!c) return fail(PVT_ERR_INVALID, "ctx is NULL");
    CK(cudaSetDevice(c->cfg.device));
    int r = pvt_sync(c);
    if (r) return r;
    CK(cudaEventRecord(c->timer_a, c->compute));
    return PVT_OK;
}

int pvt_timer_stop(pvt_ctx* c, double* ms)
{
    if (!c || !ms) return fail(PVT_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->copy));
    CK(cudaEventRecord(c->timer_b, c->compute));
    CK(cudaEventSynchronize(c->timer_b));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, c->timer_a, c->timer_b));
    *ms = (double)f;
    return PVT_OK;
}

}  // extern "C"
