/*
 * pvt.h -- C ABI of the B200-native NCC tracking hot path (libpvt.so).
 *
 * Drop-in boundary for ONE path of askEric0/Parallel-Video-Object-Tracker: per-frame normalized
 * cross-correlation search of the object template over the +-R window around the previous box,
 * peak pick, confidence-gated EMA template update.  Plain pointers and sizes only; no C++ / torch
 * types; nothing here throws or calls exit().  All compute runs in hand-written sm_100a CUDA
 * kernels; there is NO CPU fallback: every entry point fails with PVT_ERR_CUDA when no CUDA device
 * is usable.
 *
 * Reference interfaces replaced (paths relative to /root/reference):
 *   tracker/include/baseline_kernel.hpp:8-17   six baseline::ncc_match_* operators     -> pvt_ncc_match, pvt_ncc_match_batched
 *   tracker/include/utils.hpp:5-14             toGrayF32                                -> pvt_to_gray_f32 (and the ingest stage of pvt_step)
 *   tracker/src/main.cpp:6-20                  SEARCH_RADIUS_X/Y, NCC_*_CONFIDENCE, LR  -> pvt_params
 *   tracker/src/main.cpp:70-71                 template cut from frame 0                -> pvt_track_init
 *   tracker/src/main.cpp:103-161               NCC -> window clamp -> minMaxLoc -> gates -> addWeighted -> pvt_step / pvt_submit
 *   tracker/src/main.cpp:115-130               --batch=N hold semantics                 -> pvt_params.mode == PVT_MODE_BATCH
 *   tracker/src/baseline_kernel.cu:12-18       checkCuda -> exit(1)                     -> negative return code + pvt_last_error()
 *   tracker_ghc/src/main.cpp:17-23,183-239     lost-object logic: whole-frame re-acquisition -> pvt_params.lost_frame_threshold / ncc_global_confidence
 *   tracker/src/baseline_kernel.cu:44-49,62,329-332  score formula of the five CUDA kernels (differs from --cpu) -> pvt_params.formula = PVT_FORMULA_EPS,
 *                                                    PVT_MODE_FLAG_EPS, pvt_ncc_match_batched_f (opt-in; the default is the --cpu path's)
 *
 * Results follow the reference's `--cpu` path (cv::matchTemplate TM_CCOEFF_NORMED, OpenCV 4.13.0):
 * identical peak per frame (ties -> lowest row-major index), scores within 1e-4, identical bbox
 * trajectory (tests/ check this against oracle/ and the cv2 golden vectors).
 *
 * Threading: one pvt_ctx belongs to one device and is used from one host thread at a time;
 * separate contexts are independent (that is the multi-GPU model: shard tracks over contexts).
 */
#ifndef PVT_H_
#define PVT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PVT_API __attribute__((visibility("default")))
#else
#define PVT_API
#endif

#define PVT_VERSION 200

/* return codes */
#define PVT_OK               0
#define PVT_ERR_INVALID     (-1) /* bad argument; where the reference CV_Asserts (ncc_cpu.cpp:7-10, baseline_kernel.cu:315-325) */
#define PVT_ERR_CUDA        (-2) /* CUDA failure or no device; where the reference calls std::exit (baseline_kernel.cu:12-18)   */
#define PVT_ERR_UNSUPPORTED (-3) /* e.g. PVT_MODE_CPU: the library has no CPU path                                             */
#define PVT_ERR_STATE       (-4) /* call out of order (track not initialised, results pending, ...)                            */
#define PVT_ERR_NOMEM       (-5)

/* tracker/src/main.cpp:29-41 mode flags.  All GPU modes produce the same (OpenCV-semantics) numbers;
 * the mode selects hold semantics (BATCH) and, with PVT_KERNEL_AUTO, nothing else. */
typedef enum pvt_mode {
    PVT_MODE_NAIVE = 0,       /* default in the reference (NCC_MODE = "naive", main.cpp:8) */
    PVT_MODE_CPU = 1,         /* --cpu : rejected with PVT_ERR_UNSUPPORTED (oracle/ is the CPU path, test-only) */
    PVT_MODE_SHARED = 2,      /* --shared */
    PVT_MODE_CONST = 3,       /* --const */
    PVT_MODE_CONST_TILED = 4, /* --const_tiled */
    PVT_MODE_BATCH = 5        /* --batch=N : search every N-th frame, hold the box in between (main.cpp:115-130) */
} pvt_mode;

typedef enum pvt_kernel {
    PVT_KERNEL_AUTO = 0,   /* production kernel (TMA-staged tile, register-blocked FP32) */
    PVT_KERNEL_DIRECT = 1, /* one thread per candidate, global loads: verification twin of the reference's naive kernel */
    PVT_KERNEL_TC = 2,     /* tensor-core search (tcgen05.mma kind::i8, exact integer cross term of the 16-bit fixed-point centred
                            * template; scores within 1e-4 of the CPU path like the others, peaks identical).  For the throughput
                            * shapes (many tracks / ROIs per GPU) and for wide maps: windows wider than one accumulator (256
                            * candidate columns: 4K windows, the whole-frame pass of the lost-object mode) are cut into column
                            * tiles.  Must be chosen at pvt_create (it allocates an 8-bit gray plane); needs 8-bit frames (BGR8 /
                            * GRAY8 -- a GRAYF32 frame is rejected), template height <= 129, template width <= 260. */
    PVT_KERNEL_TC_GLOBAL = 3 /* lost-object mode: the planner's FP32 kernels for the local windows (the single-stream latency
                            * shape is faster there), the tensor-core search for the whole-frame pass only
                            * (tracker_ghc/src/main.cpp:186-193).  Same creation-time and frame-format rules as PVT_KERNEL_TC;
                            * pvt_set_params may switch such a context between AUTO, TC and TC_GLOBAL at any time. */
} pvt_kernel;

/* frame ingest (utils.hpp:5-14 toGrayF32).  FULL converts whole frames, as the reference does.  ROI converts only each
 * track's search tile (window + template extent; the only pixels the path ever reads) with the window origin taken from
 * device state; pinned host frames are then read zero-copy over PCIe instead of being copied whole.  Results are
 * identical.  AUTO picks ROI when the tiles of a context cover less than half of its frames' area. */
typedef enum pvt_ingest { PVT_INGEST_AUTO = 0, PVT_INGEST_FULL = 1, PVT_INGEST_ROI = 2 } pvt_ingest;

/* Score formula.  CCOEFF_NORMED is what the reference's CPU path computes (cv::matchTemplate, ncc_cpu.cpp:12) and what
 * the parity gates are stated against.  EPS is the formula of the reference's five CUDA kernels
 * (baseline_kernel.cu:44-49,62 with the host-side template statistics of :329-332):
 *     ncc = cov / ((sqrt(max(var_w, 1e-6)) + 1e-6) * (sigma_t + 1e-6 + 1e-6) * N)
 * for callers that depend on the reference's GPU-mode numbers: no clamp rules, a flat template scores 0 (not 1), a flat
 * window scores ~0 through the 1e-3 floor.  Window sums come from the same FP64 integrals and the cross term from the same
 * FP32 kernel as the default formula, so the values agree with exact arithmetic to ~1e-6; the reference's own kernels
 * accumulate everything sequentially in FP32 and sit ~4e-5 away on textured windows.  Must be chosen at pvt_create. */
typedef enum pvt_formula { PVT_FORMULA_CCOEFF_NORMED = 0, PVT_FORMULA_EPS = 1 } pvt_formula;

typedef enum pvt_format { PVT_FMT_BGR8 = 0, PVT_FMT_GRAY8 = 1, PVT_FMT_GRAYF32 = 2 } pvt_format;
/* HOST: any host pointer; the library probes whether it is page-locked (one driver query per frame descriptor) and stages
 * pageable buffers.  HOST_PINNED: the caller states the buffer is page-locked (pvt_alloc_pinned, cudaHostAlloc,
 * cudaHostRegister) -- no query on the submission path; a buffer that is not ends in PVT_ERR_INVALID or a CUDA fault. */
typedef enum pvt_memory { PVT_MEM_HOST = 0, PVT_MEM_DEVICE = 1, PVT_MEM_HOST_PINNED = 2 } pvt_memory;

/* tracker/src/main.cpp:6-20 */
typedef struct pvt_params {
    int search_radius_x;          /* SEARCH_RADIUS_X = 80 */
    int search_radius_y;          /* SEARCH_RADIUS_Y = 80 */
    double ncc_min_confidence;    /* NCC_MIN_CONFIDENCE = 0.40   (compared in double, main.cpp:153) */
    double ncc_strong_confidence; /* NCC_STRONG_CONFIDENCE = 0.70 (main.cpp:157) */
    double template_update_lr;    /* TEMPLATE_UPDATE_LR = 0.10   (main.cpp:159) */
    int batch_size;               /* BATCH_SIZE = 4, used when mode == PVT_MODE_BATCH */
    int mode;                     /* pvt_mode */
    int kernel;                   /* pvt_kernel */
    int keep_maps;                /* != 0: keep every track's last window map for pvt_get_window_map (tests) */
    int ingest;                   /* pvt_ingest */
    /* lost-object re-acquisition, tracker_ghc/src/main.cpp:17-23,143-144,183-239 (the second tracker of the reference).
     * 0 = off (tracker/src/main.cpp semantics).  > 0: after that many consecutive frames below the confidence
     * threshold the track is searched over the WHOLE frame (global arg-max of the full NCC map) and only accepted
     * at ncc_global_confidence; once found it returns to the local window.  Must be chosen at pvt_create (it sizes
     * the whole-frame scratch: ~53 MB per track at 1080p); the value may be changed later while it stays > 0. */
    int lost_frame_threshold;     /* LOST_FRAME_THRESHOLD = 50 in tracker_ghc */
    int formula;                  /* pvt_formula */
    int reserved;
    double ncc_global_confidence; /* NCC_GLOBAL_CONFIDENCE = 0.60 in tracker_ghc */
} pvt_params;

typedef struct pvt_config {
    int device;       /* CUDA device ordinal */
    int frame_w;      /* all streams of a context share one frame geometry */
    int frame_h;
    int max_streams;  /* independent frame sources (videos) */
    int max_tracks;   /* tracked objects; each is bound to one stream */
    int max_templ_w;  /* largest template (= bbox) a track may use */
    int max_templ_h;
    int max_radius_x; /* 0: take pvt_params.search_radius_x at creation */
    int max_radius_y;
    int reserved[7];
} pvt_config;

/* one input frame for one stream at one time step */
typedef struct pvt_frame {
    int stream;       /* 0 .. max_streams-1 */
    int format;       /* pvt_format; BGR8 is what cv::VideoCapture hands the reference (main.cpp:95) */
    int memory;       /* pvt_memory.  Host buffers: pageable ones are staged with cudaMemcpyAsync; pinned ones (pvt_alloc_pinned,
                       * cudaHostAlloc, cudaHostRegister) are read in place by the ROI ingest, tiles only.  Either way the
                       * buffer must stay unchanged until the step has run (pvt_step returns / pvt_sync / pvt_collect). */
    int reserved;
    const void* data; /* first row */
    size_t step;      /* bytes per row (cv::Mat::step) */
} pvt_frame;

/* per-frame output the reference only draws (main.cpp:166); the new API emits it */
typedef struct pvt_result {
    int32_t x, y, w, h; /* bbox after this frame */
    float conf;         /* bestVal of main.cpp:150 (NaN when the frame was held in batch mode) */
    uint8_t moved;      /* conf >= ncc_min_confidence    (main.cpp:153) */
    uint8_t updated;    /* conf >= ncc_strong_confidence (main.cpp:157): template EMA applied */
    uint8_t searched;   /* 0 for frames held by batch mode, 1 local window search, 2 whole-frame search (lost-object mode) */
    uint8_t valid;      /* track was active and its stream had a frame */
    int32_t track;
    int32_t step;       /* time-step index since creation */
} pvt_result;

/* device-time accounting, filled when profiling is enabled (bench.py roofline) */
typedef struct pvt_profile {
    double ingest_ms, stats_ms, ncc_ms, update_ms; /* summed CUDA-event time per kernel class */
    int64_t ingest_launches, stats_launches, ncc_launches, update_launches;
    int64_t steps;
    double ncc_macs;       /* algorithmic MACs (n_cand * tw * th) summed over the profiled steps */
    double ingest_bytes;   /* algorithmic bytes read + written by the ingest kernel */
    double search_kernel_ms;   /* k_ncc_search alone (ncc_ms also covers the tail reduction and k_ncc_fringe) */
    double search_kernel_macs; /* MACs of the candidates k_ncc_search's thread-tile grid computes (the rest: k_ncc_fringe) */
} pvt_profile;

typedef struct pvt_ctx pvt_ctx;

PVT_API int pvt_version(void);
PVT_API const char* pvt_last_error(void); /* thread-local, never NULL */
PVT_API int pvt_device_count(void);       /* >= 0, or PVT_ERR_CUDA */
/* sm count, clock (kHz), memory clock (kHz), total memory bytes, cc major*10+minor */
PVT_API int pvt_device_info(int device, int* sm_count, int* sm_clock_khz, int* mem_clock_khz, size_t* mem_bytes, int* cc);

PVT_API void pvt_default_params(pvt_params* p); /* the constants of main.cpp:6-20 */
PVT_API void pvt_default_params_ghc(pvt_params* p); /* the constants of tracker_ghc/src/main.cpp:9-23 (radius 60, lost-object mode on) */
PVT_API int pvt_create(pvt_ctx** out, const pvt_params* params, const pvt_config* config);
PVT_API int pvt_destroy(pvt_ctx* ctx);
PVT_API int pvt_set_params(pvt_ctx* ctx, const pvt_params* params); /* radii must stay <= the creation maxima */

/* pinned host memory for frames that are streamed from the host */
PVT_API int pvt_alloc_pinned(void** out, size_t bytes);
PVT_API int pvt_free_pinned(void* p);

/* main.cpp:70-71: ingest `frame0` into its stream (NULL: keep the stream's current image) and cut the
 * template of `track` from it at (x, y, w, h).  Synchronous.  frame0 == NULL needs the stream's plane to hold one complete
 * frame: that is the case after a pvt_track_init with a frame and after steps of a context that ingests whole frames
 * (PVT_INGEST_FULL, or AUTO with many tracks); a context on the ROI ingest only refreshes the search tiles of its
 * tracks, so there a later pvt_track_init(frame0 = NULL) fails with PVT_ERR_STATE instead of cutting stale pixels. */
PVT_API int pvt_track_init(pvt_ctx* ctx, int track, int stream, const pvt_frame* frame0, int x, int y, int w, int h);
PVT_API int pvt_track_remove(pvt_ctx* ctx, int track);

/* One time step (main.cpp:98-161 for every active track whose stream got a frame): ingest, window
 * statistics, NCC search, peak, gates, EMA -- one CUDA-graph launch, no host round trip inside.
 * pvt_step waits and writes one pvt_result per track slot 0..max_tracks-1 (results may be NULL).
 * pvt_submit only enqueues; pvt_collect waits for everything submitted and returns the results of
 * the last `max_steps` steps, oldest first, max_tracks entries per step; returns the step count. */
PVT_API int pvt_step(pvt_ctx* ctx, int n_frames, const pvt_frame* frames, pvt_result* results);
PVT_API int pvt_submit(pvt_ctx* ctx, int n_frames, const pvt_frame* frames);
PVT_API int pvt_collect(pvt_ctx* ctx, pvt_result* results, int max_steps);
/* The frame loop of main.cpp:93-169 in one call: n_steps time steps, step s uses the n_frames descriptors
 * frames[(s % ring_len) * n_frames ...] (a ring of resident or pinned-host frames).  Enqueues only; at most
 * the last 64 steps' results are kept (pvt_collect).  collect_every > 0: every that many steps the results
 * ring is read back to the host inside the loop (results_out may be NULL to discard after reading). */
PVT_API int pvt_submit_sequence(pvt_ctx* ctx, int n_steps, int n_frames, const pvt_frame* frames, int ring_len,
                                int collect_every, pvt_result* results_out);
PVT_API int pvt_sync(pvt_ctx* ctx);

/* tracker state = {bbox, template} (the reference keeps it in host variables, main.cpp:63-71) */
PVT_API int pvt_get_state(pvt_ctx* ctx, int track, int32_t bbox[4], float* templ, size_t templ_step_bytes);
PVT_API int pvt_set_state(pvt_ctx* ctx, int track, const int32_t bbox[4], const float* templ, size_t templ_step_bytes);
/* lost-object state of a track (tracker_ghc/src/main.cpp:143-144 lost_frame_count, use_global_search), for checkpoint / resume.
 * use_global_search is reported as the NEXT frame will evaluate it (:183-185 raises it at the start of a frame). */
PVT_API int pvt_get_lost_state(pvt_ctx* ctx, int track, int* lost_frame_count, int* use_global_search);
PVT_API int pvt_set_lost_state(pvt_ctx* ctx, int track, int lost_frame_count, int use_global_search);
/* last window map of a track (needs params.keep_maps): win = {minTx, minTy, width, height} (main.cpp:143-147) */
PVT_API int pvt_get_window_map(pvt_ctx* ctx, int track, float* out, size_t out_step_bytes, int32_t win[4]);

/* utils.hpp:5-14 toGrayF32 on its own: BGR8/GRAY8 -> f32/255 image (host or device destination) */
PVT_API int pvt_to_gray_f32(pvt_ctx* ctx, const pvt_frame* frame, float* out, size_t out_step_bytes, int out_memory);

/* baseline_kernel.hpp:8-17 map-level operators: full (fh-th+1) x (fw-tw+1) map, host buffers in, host
 * buffer out, synchronous -- the contract of ncc_match_naive_cuda / _shared_cuda / _const / _const_tiled
 * (the low byte of mode picks nothing but is validated; PVT_MODE_CPU -> PVT_ERR_UNSUPPORTED).  Values are those of the
 * reference's CPU operator ncc_match_cpu unless mode carries PVT_MODE_FLAG_EPS, which selects PVT_FORMULA_EPS, the
 * formula of those four GPU operators.  Unlike them there is no 4096-pixel template limit (baseline_kernel.cu:500)
 * and no 48 KB limit.  pvt_ncc_match_batched takes the same flag in `formula` through pvt_ncc_match_batched_f. */
#define PVT_MODE_FLAG_EPS 0x100
PVT_API int pvt_ncc_match(int device, int mode, const float* frame, int fw, int fh, size_t fstep_bytes,
                          const float* templ, int tw, int th, size_t tstep_bytes, float* out, size_t ostep_bytes);
/* baseline_kernel.hpp:14 ncc_match_naive_cuda_batched: n frames of one geometry against one template */
PVT_API int pvt_ncc_match_batched(int device, int n, const float* const* frames, int fw, int fh, size_t fstep_bytes,
                                  const float* templ, int tw, int th, size_t tstep_bytes, float* const* outs, size_t ostep_bytes);
PVT_API int pvt_ncc_match_batched_f(int device, int formula, int n, const float* const* frames, int fw, int fh, size_t fstep_bytes,
                                    const float* templ, int tw, int th, size_t tstep_bytes, float* const* outs, size_t ostep_bytes);

/* The search-kernel plan pvt_create would derive for this geometry on a device with `sm_count` SMs; pure host logic (no
 * device needed).  out = {G, C, GB, bands, ctas_per_band, span, boxW, boxH, pj, pd, ctas_per_track, n_full, n_tail, tail_parts,
 * shared bytes per CTA, fringe bits (1: remainder column, 2: remainder row computed by k_ncc_fringe)}. */
PVT_API int pvt_plan_query(int sm_count, int n_tracks, int templ_w, int templ_h, int frame_w, int frame_h, int radius_x, int radius_y,
                           int32_t out[16]);

/* The geometry of the tensor-core search (PVT_KERNEL_TC, k_ncc_tc) for the local windows (whole_frame_pass = 0) or for the
 * whole-frame pass of the lost-object mode (1: tracker_ghc/src/main.cpp:186-193, the radii are ignored); pure host logic.
 * out = {candidate columns per accumulator XW, column tiles per window, 128-row tiles per window, K-steps of 32 image columns,
 * candidate groups of 8 per accumulator, TMEM columns, depth of the Toeplitz block ring, shared bytes per CTA}.
 * PVT_ERR_UNSUPPORTED: template higher than 129 rows / wider than 260 columns. */
PVT_API int pvt_tc_plan_query(int sm_count, int n_tracks, int templ_w, int templ_h, int frame_w, int frame_h, int radius_x, int radius_y,
                              int whole_frame_pass, int32_t out[8]);

/* The sink side of the reference loop (tracker/src/main.cpp:166  cv::rectangle(frame, bbox, {0,255,0}, 2)): paint the boxes
 * (n x {x, y, w, h}) onto a BGR8 frame, in place, with cv::rectangle's thickness-2 pixel coverage (bit-identical to OpenCV 4.13).
 * frame->memory says where the pixels live: device frames are painted where they are; host frames make the round trip through
 * a staging buffer.  Boxes must lie inside the frame (the tracker's always do); bgr = {b, g, r}; NULL = the reference's green {0, 255, 0}. */
PVT_API int pvt_draw_boxes(pvt_ctx* ctx, const pvt_frame* frame, int n, const int32_t* boxes_xywh, const uint8_t* bgr);

/* measurement hooks */
PVT_API int pvt_profile_enable(pvt_ctx* ctx, int on); /* on: per-kernel CUDA events, plain stream launches */
PVT_API int pvt_profile_get(pvt_ctx* ctx, pvt_profile* out, int reset);
/* device-side timeline: with tracing on, every kernel stamps %globaltimer (ns) at its first CTA's start and last
 * CTA's end; pvt_trace_get copies the stamps of the last <= 64 steps: out[step][8 kernel slots][2], slots =
 * ingest, statistics (k_winstats; k_colprefix of the two-kernel statistics), rowsum, search (k_ncc_search / k_ncc_local / k_ncc_tc),
 * ncc_finalize, update, ncc_fringe, ncc_tail_finalize.  Slots a plan does not launch carry phase stamps of CTA 0 instead:
 * k_ncc_local and the K-split k_ncc_search use the last two (staged | loop done, statistics / loop done | reduced), k_ncc_tc
 * slots 4, 6, 7, k_prefetch_roi slot 2 in the k_ncc_local shape (tools/timeline.py, tools/tc_timeline.py decode them).
 * Returns the number of steps written. */
PVT_API int pvt_trace_enable(pvt_ctx* ctx, int on);
PVT_API int pvt_trace_get(pvt_ctx* ctx, uint64_t* out, int max_steps);
PVT_API int64_t pvt_launch_count(pvt_ctx* ctx); /* kernels launched by this context so far (graph nodes counted per launch) */
/* which search kernel the context's plan runs per step (pvt_create decides from the geometry): writes its name into `name`
 * ("k_ncc_search", "k_ncc_search+k_ncc_finalize" (K-split), "k_step_fused", "k_ncc_local", "k_ncc_tc", "k_ncc_direct") and
 * returns the number of kernels one searched step launches. */
PVT_API int pvt_search_kind(pvt_ctx* ctx, char* name, int name_bytes);
/* device time of everything submitted between pvt_timer_start and pvt_timer_stop, CUDA events on the context's stream */
PVT_API int pvt_timer_start(pvt_ctx* ctx);
PVT_API int pvt_timer_stop(pvt_ctx* ctx, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* PVT_H_ */
