/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the NCC tracking hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product library (libpvt.so) never links, loads or calls
 * it and has no CPU fallback.
 *
 * What it restates.  The reference's `--cpu` path is
 *     toGrayF32                         /root/reference/tracker/include/utils.hpp:5-14
 *     baseline::ncc_match_cpu           /root/reference/tracker/src/ncc_cpu.cpp:5-13
 *     window clamp / peak / gate / EMA  /root/reference/tracker/src/main.cpp:135-161
 *     batch hold semantics              /root/reference/tracker/src/main.cpp:115-130
 *     lost-object re-acquisition        /root/reference/tracker_ghc/src/main.cpp:17-23,143-144,183-239 (orc_track_clip_ghc)
 *     the CUDA kernels' eps formula     /root/reference/tracker/src/baseline_kernel.cu:21-64,329-332 (orc_ncc_window_eps;
 *                                       PARITY UNPINNED: those kernels cannot be run in the build container, see there)
 * and every arithmetic call on it lands in a THIRD-PARTY dependency that is not vendored under
 * /root/reference: OpenCV (imgproc/core), pinned at 4.13.0 by tracker/Makefile:19-27
 * (opencv_core4130.lib ...).  The functions below restate OpenCV 4.13.0's published algorithms
 * for those calls (cvtColor BGR2GRAY 8u fixed point, convertTo scale, integral 32f->64f,
 * meanStdDev, matchTemplate TM_CCOEFF_NORMED normalisation, minMaxLoc, addWeighted 32f).
 *
 * Parity pin.  The reference ships no tests, golden vectors or fixtures, so the
 * pin is the real library: tests/golden/make_golden.py runs cv2 4.13.0 (same version) in the
 * build container through oracle/cv2_harness.py and commits its outputs; tests/test_oracle_*.py
 * checks every function here against those fixtures (bit-exact for the integer/byte/EMA parts,
 * <=1.5e-7 against the IPP-off matchTemplate, <=1e-4 against the default IPP-on one, identical
 * peak per frame and identical trajectories).
 *
 * One deliberate difference from a literal transcription: the cross term sum(f*t) is accumulated
 * here in double (error ~1e-13) and then rounded to float32 exactly where OpenCV stores it
 * (crossCorr writes a CV_32F result), instead of reproducing IPP's / the DFT's internal float
 * noise, which is build- and CPU-dependent (SURVEY.md §8(c) "oracle noise model").
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define ORC_API __attribute__((visibility("default")))

/* worker threads used for the candidate rows of orc_ncc_window (rows are independent, so the
 * result does not depend on the thread count); override with orc_set_threads(). */
static int g_threads = 0;
ORC_API int orc_num_threads(void)
{
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    return (int)n;
}
ORC_API void orc_set_threads(int n) { g_threads = n; }

/* ---- ingest: utils.hpp:8 cvtColor(BGR2GRAY), 8u -------------------------------------------
 * OpenCV 4.x 8u path: 15-bit fixed point, B2Y=3735 G2Y=19235 R2Y=9798, rounding add 1<<14. */
ORC_API void orc_bgr2gray(const uint8_t* bgr, int w, int h, size_t step, uint8_t* gray, size_t gstep)
{
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = bgr + (size_t)y * step;
        uint8_t* d = gray + (size_t)y * gstep;
        for (int x = 0; x < w; ++x, s += 3)
            d[x] = (uint8_t)((3735 * s[0] + 19235 * s[1] + 9798 * s[2] + 16384) >> 15);
    }
}

/* ---- ingest: utils.hpp:12 gray.convertTo(gray_f32, CV_32F, 1.0f/255.0f) -------------------
 * one rounding of the exact product g * (float)(1/255)  ->  a 256-entry table. */
ORC_API void orc_gray_to_f32(const uint8_t* gray, int w, int h, size_t gstep, float* out, size_t ostep_bytes)
{
    const float a = 1.0f / 255.0f;
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = gray + (size_t)y * gstep;
        float* d = (float*)((char*)out + (size_t)y * ostep_bytes);
        for (int x = 0; x < w; ++x) d[x] = (float)((double)s[x] * (double)a);
    }
}

ORC_API void orc_to_gray_f32(const uint8_t* bgr, int w, int h, size_t step, float* out, size_t ostep_bytes)
{
    uint8_t* g = (uint8_t*)malloc((size_t)w * h);
    orc_bgr2gray(bgr, w, h, step, g, (size_t)w);
    orc_gray_to_f32(g, w, h, (size_t)w, out, ostep_bytes);
    free(g);
}

/* ---- template statistics: cv::meanStdDev on CV_32FC1 (inside matchTemplate; also
 * baseline_kernel.cu:329-332).  Population mean / sigma in double. */
ORC_API void orc_mean_stddev(const float* t, int tw, int th, size_t tstep_bytes, double* mean, double* sdv)
{
    double s = 0.0, sq = 0.0;
    for (int y = 0; y < th; ++y) {
        const float* r = (const float*)((const char*)t + (size_t)y * tstep_bytes);
        for (int x = 0; x < tw; ++x) {
            double v = r[x];
            s += v;
            sq += v * v;
        }
    }
    double n = (double)tw * th;
    double m = s / n;
    double var = sq / n - m * m;
    if (var < 0) var = 0;
    *mean = m;
    *sdv = sqrt(var);
}

/* ---- cv::integral(img, sum, sqsum, CV_64F) on CV_32FC1: (h+1) x (w+1) doubles, row 0 / col 0
 * zero, each row: running row sums added to the row above (OpenCV's integral_ loop order). */
static void integral_f32(const float* f, int w, int h, size_t fstep_bytes, double* sum, double* sq)
{
    int sw = w + 1;
    memset(sum, 0, sizeof(double) * sw);
    memset(sq, 0, sizeof(double) * sw);
    for (int y = 0; y < h; ++y) {
        const float* r = (const float*)((const char*)f + (size_t)y * fstep_bytes);
        double* s1 = sum + (size_t)(y + 1) * sw;
        double* q1 = sq + (size_t)(y + 1) * sw;
        const double* s0 = s1 - sw;
        const double* q0 = q1 - sw;
        double s = 0.0, q = 0.0;
        s1[0] = 0.0;
        q1[0] = 0.0;
        for (int x = 0; x < w; ++x) {
            double it = r[x];
            s += it;
            q += it * it;
            s1[x + 1] = s0[x + 1] + s;
            q1[x + 1] = q0[x + 1] + q;
        }
    }
}

typedef struct ncc_job {
    const float* frame; size_t fstep_bytes; const float* tc; int tw, th, x0, y0, ww, wh;
    float* out; size_t ostep_bytes; const double* sum; const double* sq; int sw;
    double mean_t, templNorm, invArea; int first, stride;
} ncc_job;

/* candidate rows first, first+stride, ...: crossCorr + common_matchTemplate for one row each */
static void ncc_rows_run(const ncc_job* j, int first, int stride)
{
    const int tw = j->tw, th = j->th, sw = j->sw;
    for (int y = first; y < j->wh; y += stride) {
        float* o = (float*)((char*)j->out + (size_t)y * j->ostep_bytes);
        int Y = j->y0 + y;
        for (int x = 0; x < j->ww; ++x) {
            int X = j->x0 + x;
            double acc = 0.0;
            for (int dy = 0; dy < th; ++dy) {
                const float* fr = (const float*)((const char*)j->frame + (size_t)(Y + dy) * j->fstep_bytes) + X;
                const float* tr = j->tc + (size_t)dy * tw;
                double a = 0.0;
                for (int dx = 0; dx < tw; ++dx) a += (double)fr[dx] * (double)tr[dx];
                acc += a;
            }
            float cc = (float)acc; /* crossCorr result is CV_32F */
            double num = cc, t;
            const double* p0 = j->sum + (size_t)Y * sw + X;
            const double* p2 = j->sum + (size_t)(Y + th) * sw + X;
            const double* q0 = j->sq + (size_t)Y * sw + X;
            const double* q2 = j->sq + (size_t)(Y + th) * sw + X;
            t = p0[0] - p0[tw] - p2[0] + p2[tw];
            double wndMean2 = t * t;
            num -= t * j->mean_t;
            wndMean2 *= j->invArea;
            double wndSum2 = q0[0] - q0[tw] - q2[0] + q2[tw];
            double diff2 = wndSum2 - wndMean2;
            if (diff2 < 0) diff2 = 0;
            double lim = 10 * (double)FLT_EPSILON * wndSum2;
            if (lim > 0.5) lim = 0.5;
            if (diff2 <= lim) t = 0;
            else t = sqrt(diff2) * j->templNorm;
            if (fabs(num) < t) num /= t;
            else if (fabs(num) < t * 1.125) num = num > 0 ? 1 : -1;
            else num = 0;
            o[x] = (float)num;
        }
    }
}
static void* ncc_rows(void* p)
{
    const ncc_job* j = (const ncc_job*)p;
    ncc_rows_run(j, j->first, j->stride);
    return NULL;
}

/* ---- ncc_cpu.cpp:12  cv::matchTemplate(frame, templ, map, TM_CCOEFF_NORMED), restricted to the
 * candidate rectangle [x0,x0+ww) x [y0,y0+wh) of the full map (the only part main.cpp:147-151
 * ever reads).  out is ww x wh, row stride ostep_bytes.  Integrals are taken over the WHOLE
 * frame, as OpenCV does, so the rounding of the box sums is OpenCV's.
 *
 * crossCorr:            cc   = (float) sum_{dy,dx} f[y+dy][x+dx] * t[dy][dx]
 * common_matchTemplate: num  = (double)cc - wsum * mean_t
 *                       d2   = max(wsq - wsum^2 * invArea, 0)
 *                       t    = d2 <= min(0.5, 10*FLT_EPSILON*wsq) ? 0 : sqrt(d2) * sigma_t/sqrt(invArea)
 *                       out  = |num| < t ? num/t : |num| < 1.125 t ? +-1 : 0
 * and the whole map is 1 when sigma_t^2 < DBL_EPSILON. */
ORC_API int orc_ncc_window(const float* frame, int fw, int fh, size_t fstep_bytes,
                           const float* templ, int tw, int th, size_t tstep_bytes,
                           int x0, int y0, int ww, int wh, float* out, size_t ostep_bytes)
{
    if (tw <= 0 || th <= 0 || fw < tw || fh < th) return -1;
    int outW = fw - tw + 1, outH = fh - th + 1;
    if (x0 < 0 || y0 < 0 || ww <= 0 || wh <= 0 || x0 + ww > outW || y0 + wh > outH) return -2;

    double mean_t, sdv_t;
    orc_mean_stddev(templ, tw, th, tstep_bytes, &mean_t, &sdv_t);
    double templNorm = sdv_t * sdv_t;
    if (templNorm < DBL_EPSILON) {
        for (int y = 0; y < wh; ++y) {
            float* o = (float*)((char*)out + (size_t)y * ostep_bytes);
            for (int x = 0; x < ww; ++x) o[x] = 1.0f;
        }
        return 0;
    }
    double invArea = 1.0 / ((double)th * tw);
    templNorm = sqrt(templNorm);
    templNorm /= sqrt(invArea);

    int sw = fw + 1;
    double* sum = (double*)malloc(sizeof(double) * (size_t)sw * (fh + 1));
    double* sq = (double*)malloc(sizeof(double) * (size_t)sw * (fh + 1));
    if (!sum || !sq) { free(sum); free(sq); return -3; }
    integral_f32(frame, fw, fh, fstep_bytes, sum, sq);

    /* contiguous copy of the template so the inner loop is a plain dot product */
    float* tc = (float*)malloc(sizeof(float) * (size_t)tw * th);
    for (int y = 0; y < th; ++y)
        memcpy(tc + (size_t)y * tw, (const char*)templ + (size_t)y * tstep_bytes, sizeof(float) * tw);

    ncc_job job = { frame, fstep_bytes, tc, tw, th, x0, y0, ww, wh, out, ostep_bytes, sum, sq, sw,
                    mean_t, templNorm, invArea, 0, 1 };
    int nt = orc_num_threads();
    if (nt > wh) nt = wh;
    if (nt <= 1) {
        ncc_rows(&job);
    } else {
        pthread_t th_[64];
        ncc_job jobs[64];
        for (int i = 0; i < nt; ++i) {
            jobs[i] = job; jobs[i].first = i; jobs[i].stride = nt;
            if (pthread_create(&th_[i], NULL, ncc_rows, &jobs[i])) { jobs[i].stride = -1; ncc_rows_run(&jobs[i], i, nt); }
        }
        for (int i = 0; i < nt; ++i) if (jobs[i].stride > 0) pthread_join(th_[i], NULL);
    }
    free(tc);
    free(sum);
    free(sq);
    return 0;
}

/* full map: the literal ncc_match_cpu contract (ncc_cpu.cpp:5-13), out is (fh-th+1) x (fw-tw+1) */
ORC_API int orc_ncc_match_cpu(const float* frame, int fw, int fh, size_t fstep_bytes,
                              const float* templ, int tw, int th, size_t tstep_bytes,
                              float* out, size_t ostep_bytes)
{
    return orc_ncc_window(frame, fw, fh, fstep_bytes, templ, tw, th, tstep_bytes,
                          0, 0, fw - tw + 1, fh - th + 1, out, ostep_bytes);
}

/* ---- baseline_kernel.cu:21-64 nccKernelNaive (+ host statistics :329-332): the formula all five CUDA kernels of the
 * reference share, restricted to a candidate rectangle like orc_ncc_window.  FP32 throughout, two
 * sequential passes per candidate in the kernel's loop order (dy outer, dx inner):
 *     pass 1  sum += v; sumSq += v*v                       mean = sum/N; var = sumSq/N - mean*mean; std = sqrtf(fmaxf(var, 1e-6f))
 *     pass 2  cov += (v - mean) * (t - templMean)          ncc = cov * (1/(std + 1e-6f)) * (1/(templStd + 1e-6f)) / N
 * with templMean = (float)mean_t, templStd = (float)(sigma_t + 1e-6f) from cv::meanStdDev (double, population).
 * nvcc contracts a*b+c into one FMA by default (-fmad=true), so the two accumulations are written with fmaf here.
 * PARITY UNPINNED: the reference ships no vectors for its GPU modes and its kernels cannot run in the build container
 * (no GPU), so this restatement is checked only against exact (float64) evaluation of the same formula
 * (tests/test_oracle_golden.py), from which sequential FP32 summation drifts by ~4e-5 on textured 64x64 windows. */
ORC_API int orc_ncc_window_eps(const float* frame, int fw, int fh, size_t fstep_bytes,
                               const float* templ, int tw, int th, size_t tstep_bytes,
                               int x0, int y0, int ww, int wh, float* out, size_t ostep_bytes)
{
    if (tw <= 0 || th <= 0 || fw < tw || fh < th) return -1;
    int outW = fw - tw + 1, outH = fh - th + 1;
    if (x0 < 0 || y0 < 0 || ww <= 0 || wh <= 0 || x0 + ww > outW || y0 + wh > outH) return -2;
    double mean_t, sdv_t;
    orc_mean_stddev(templ, tw, th, tstep_bytes, &mean_t, &sdv_t);
    const float templMean = (float)mean_t;
    const float templStd = (float)(sdv_t + 1e-6f);
    const int N = tw * th;
    for (int oy = 0; oy < wh; ++oy) {
        float* o = (float*)((char*)out + (size_t)oy * ostep_bytes);
        for (int ox = 0; ox < ww; ++ox) {
            float sum = 0.0f, sumSq = 0.0f;
            for (int dy = 0; dy < th; ++dy) {
                const float* fr = (const float*)((const char*)frame + (size_t)(y0 + oy + dy) * fstep_bytes) + x0 + ox;
                for (int dx = 0; dx < tw; ++dx) {
                    const float v = fr[dx];
                    sum += v;
                    sumSq = fmaf(v, v, sumSq);
                }
            }
            const float mean = sum / N;
            const float var = fmaf(-mean, mean, sumSq / N);
            const float sd = sqrtf(fmaxf(var, 1e-6f));
            const float inv_std = 1.0f / (sd + 1e-6f);
            const float inv_t_std = 1.0f / (templStd + 1e-6f);
            float cov = 0.0f;
            for (int dy = 0; dy < th; ++dy) {
                const float* fr = (const float*)((const char*)frame + (size_t)(y0 + oy + dy) * fstep_bytes) + x0 + ox;
                const float* tr = (const float*)((const char*)templ + (size_t)dy * tstep_bytes);
                for (int dx = 0; dx < tw; ++dx) cov = fmaf(fr[dx] - mean, tr[dx] - templMean, cov);
            }
            o[ox] = cov * inv_std * inv_t_std / (float)N;
        }
    }
    return 0;
}

/* which map the tracker loops below search: 0 = orc_ncc_window (the --cpu path, the parity target), 1 = orc_ncc_window_eps
 * (what main.cpp:103-133 gets from any of the GPU modes).  Test-only global, not thread-safe. */
static int g_formula = 0;
ORC_API void orc_set_formula(int f) { g_formula = f; }

/* ---- main.cpp:135-146 search window (C int arithmetic) ------------------------------------ */
ORC_API void orc_search_window(int x, int y, int w, int h, int outW, int outH, int rx, int ry, int* win /*x0,y0,ww,wh*/)
{
    int cx = x + w / 2, cy = y + h / 2;
    int minTx = cx - rx - w / 2; if (minTx < 0) minTx = 0;
    int maxTx = cx + rx - w / 2; if (maxTx > outW - 1) maxTx = outW - 1;
    int minTy = cy - ry - h / 2; if (minTy < 0) minTy = 0;
    int maxTy = cy + ry - h / 2; if (maxTy > outH - 1) maxTy = outH - 1;
    win[0] = minTx; win[1] = minTy; win[2] = maxTx - minTx + 1; win[3] = maxTy - minTy + 1;
}

/* ---- main.cpp:150 cv::minMaxLoc(view, 0, &bestVal, 0, &bestLoc): maximum and its FIRST
 * occurrence in row-major order of the view; NaN never wins (comparisons with NaN are false). */
ORC_API void orc_max_loc(const float* map, int w, int h, size_t step_bytes, double* best, int* bx, int* by)
{
    float b = -INFINITY;
    int ix = 0, iy = 0, found = 0;
    for (int y = 0; y < h; ++y) {
        const float* r = (const float*)((const char*)map + (size_t)y * step_bytes);
        for (int x = 0; x < w; ++x)
            if (!found ? (r[x] == r[x]) : (r[x] > b)) { b = r[x]; ix = x; iy = y; found = 1; }
    }
    *best = (double)b; *bx = ix; *by = iy;
}

/* ---- main.cpp:159 cv::addWeighted(templ, 1-lr, patch, lr, 0.0, templ) on CV_32FC1 ----------
 * OpenCV 4.13 widens both inputs to double, computes fma(a, alpha, b*beta) with double scalars
 * and rounds once to float (validated bit-for-bit against cv2 in tests/golden). */
ORC_API void orc_add_weighted(float* templ, const float* patch, int n, double alpha, double beta)
{
    for (int i = 0; i < n; ++i) {
        volatile double pb = (double)patch[i] * beta; /* product rounded to double first */
        templ[i] = (float)fma((double)templ[i], alpha, pb);
    }
}

typedef struct orc_record {
    int x, y, w, h;
    float conf;
    int moved, updated, searched;
} orc_record;

/* ---- one tracked frame: main.cpp:103-161 with mode == "cpu" --------------------------------
 * gray: the frame's toGrayF32 image; templ: tw*th contiguous, updated in place; bbox in/out.
 * map_out (optional, ww*wh contiguous) receives the window map; win_out (optional) the window. */
ORC_API int orc_track_step(const float* gray, int fw, int fh, float* templ, int tw, int th,
                           int* bx, int* by, int rx, int ry, double min_conf, double strong_conf, double lr,
                           orc_record* rec, float* map_out, int* win_out)
{
    int outW = fw - tw + 1, outH = fh - th + 1, win[4];
    if (outW <= 0 || outH <= 0) return -1;
    orc_search_window(*bx, *by, tw, th, outW, outH, rx, ry, win);
    if (win[2] <= 0 || win[3] <= 0) return -2;
    float* map = map_out ? map_out : (float*)malloc(sizeof(float) * (size_t)win[2] * win[3]);
    int rc = (g_formula ? orc_ncc_window_eps : orc_ncc_window)(gray, fw, fh, sizeof(float) * (size_t)fw, templ, tw, th, sizeof(float) * (size_t)tw,
                            win[0], win[1], win[2], win[3], map, sizeof(float) * (size_t)win[2]);
    if (rc) { if (!map_out) free(map); return rc; }
    double best; int lx, ly;
    orc_max_loc(map, win[2], win[3], sizeof(float) * (size_t)win[2], &best, &lx, &ly);
    if (!map_out) free(map);
    if (win_out) memcpy(win_out, win, sizeof(win));
    int moved = 0, updated = 0;
    if (best >= min_conf) {
        *bx = lx + win[0];
        *by = ly + win[1];
        moved = 1;
        if (best >= strong_conf) {
            float* patch = (float*)malloc(sizeof(float) * (size_t)tw * th);
            for (int y = 0; y < th; ++y)
                memcpy(patch + (size_t)y * tw, gray + (size_t)(*by + y) * fw + *bx, sizeof(float) * tw);
            orc_add_weighted(templ, patch, tw * th, 1 - lr, lr);
            free(patch);
            updated = 1;
        }
    }
    if (rec) {
        rec->x = *bx; rec->y = *by; rec->w = tw; rec->h = th;
        rec->conf = (float)best; rec->moved = moved; rec->updated = updated; rec->searched = 1;
    }
    return 0;
}

/* ---- whole clip: main.cpp:70-71 (template cut) + :93-169 loop; batch>1 follows :115-130 -----
 * frames: n BGR u8 frames, each fh*step bytes apart... (contiguous, step = 3*fw). recs: n-1. */
ORC_API int orc_track_clip(const uint8_t* frames, int n, int fw, int fh,
                           int x, int y, int tw, int th, int rx, int ry,
                           double min_conf, double strong_conf, double lr, int batch,
                           orc_record* recs, float* templ_out)
{
    size_t fbytes = (size_t)fw * fh * 3;
    float* gray = (float*)malloc(sizeof(float) * (size_t)fw * fh);
    float* templ = (float*)malloc(sizeof(float) * (size_t)tw * th);
    orc_to_gray_f32(frames, fw, fh, (size_t)fw * 3, gray, sizeof(float) * (size_t)fw);
    for (int r = 0; r < th; ++r) memcpy(templ + (size_t)r * tw, gray + (size_t)(y + r) * fw + x, sizeof(float) * tw);
    int pending = 0, rc = 0;
    for (int k = 1; k < n && !rc; ++k) {
        orc_record* rec = recs + (k - 1);
        orc_to_gray_f32(frames + fbytes * k, fw, fh, (size_t)fw * 3, gray, sizeof(float) * (size_t)fw);
        if (batch > 1 && ++pending < batch) {
            rec->x = x; rec->y = y; rec->w = tw; rec->h = th;
            rec->conf = NAN; rec->moved = 0; rec->updated = 0; rec->searched = 0;
            continue;
        }
        pending = 0;
        rc = orc_track_step(gray, fw, fh, templ, tw, th, &x, &y, rx, ry, min_conf, strong_conf, lr, rec, NULL, NULL);
    }
    if (templ_out) memcpy(templ_out, templ, sizeof(float) * (size_t)tw * th);
    free(gray);
    free(templ);
    return rc;
}

/* ---- tracker_ghc: the second tracker of the reference, with lost-object re-acquisition ---------
 * tracker_ghc/src/main.cpp:145-239 (demo loop; the record loop :330-410 is the same logic), mode "cpu".
 * State: lost_frame_count, use_global_search (:143-144).  Per frame:
 *   :181   bbox_outside = isBboxOutsideFrame(curr_bbox) (:49-55) -- the box always comes from a position of the NCC map,
 *          so its centre is inside the frame; evaluated anyway, as the reference does
 *   :183   if (bbox_outside || lost_frame_count >= LOST_FRAME_THRESHOLD) use_global_search = true
 *   :186   global: minMaxLoc over the WHOLE map; else the local window (:195-203) -- same clamp as tracker/
 *   :217   threshold = use_global_search ? NCC_GLOBAL_CONFIDENCE : NCC_MIN_CONFIDENCE
 *   :218   found: move, lost_frame_count = 0, use_global_search = false (box inside), EMA if >= NCC_STRONG_CONFIDENCE
 *   :236   else lost_frame_count++
 * searched = 1 local window, 2 whole map. */
typedef struct orc_record_ghc {
    int x, y, w, h;
    float conf;
    int moved, updated, searched, lost_count, use_global;
} orc_record_ghc;

static int bbox_outside_frame(int x, int y, int w, int h, int fw, int fh)
{
    int cx = x + w / 2, cy = y + h / 2;
    return (cx < 0 || cx >= fw || cy < 0 || cy >= fh) || (x + w < 0 || x >= fw || y + h < 0 || y >= fh);
}

ORC_API int orc_track_clip_ghc(const uint8_t* frames, int n, int fw, int fh,
                               int x, int y, int tw, int th, int rx, int ry,
                               double min_conf, double global_conf, double strong_conf, double lr, int lost_threshold,
                               orc_record_ghc* recs, float* templ_out)
{
    size_t fbytes = (size_t)fw * fh * 3;
    int outW = fw - tw + 1, outH = fh - th + 1;
    if (outW <= 0 || outH <= 0) return -1;
    float* gray = (float*)malloc(sizeof(float) * (size_t)fw * fh);
    float* templ = (float*)malloc(sizeof(float) * (size_t)tw * th);
    float* map = (float*)malloc(sizeof(float) * (size_t)outW * outH);
    orc_to_gray_f32(frames, fw, fh, (size_t)fw * 3, gray, sizeof(float) * (size_t)fw);
    for (int r = 0; r < th; ++r) memcpy(templ + (size_t)r * tw, gray + (size_t)(y + r) * fw + x, sizeof(float) * tw);
    int lost = 0, use_global = 0, rc = 0;
    for (int k = 1; k < n && !rc; ++k) {
        orc_record_ghc* rec = recs + (k - 1);
        orc_to_gray_f32(frames + fbytes * k, fw, fh, (size_t)fw * 3, gray, sizeof(float) * (size_t)fw);
        if (bbox_outside_frame(x, y, tw, th, fw, fh) || lost >= lost_threshold) use_global = 1;
        int win[4] = {0, 0, outW, outH};
        if (!use_global) {
            orc_search_window(x, y, tw, th, outW, outH, rx, ry, win);
            if (win[2] <= 0 || win[3] <= 0) { win[0] = 0; win[1] = 0; win[2] = outW; win[3] = outH; }   /* :204-210 fallback */
        }
        /* only the searched part of the map is needed: every cell of the map is independent of the others */
        rc = orc_ncc_window(gray, fw, fh, sizeof(float) * (size_t)fw, templ, tw, th, sizeof(float) * (size_t)tw,
                            win[0], win[1], win[2], win[3], map, sizeof(float) * (size_t)win[2]);
        if (rc) break;
        double best; int lx, ly;
        orc_max_loc(map, win[2], win[3], sizeof(float) * (size_t)win[2], &best, &lx, &ly);
        const int searched = use_global ? 2 : 1;
        const double thr = use_global ? global_conf : min_conf;
        int moved = 0, updated = 0;
        if (best >= thr) {
            x = lx + win[0]; y = ly + win[1];
            moved = 1;
            lost = 0;
            if (!bbox_outside_frame(x, y, tw, th, fw, fh)) use_global = 0;
            if (best >= strong_conf) {
                float* patch = (float*)malloc(sizeof(float) * (size_t)tw * th);
                for (int r = 0; r < th; ++r) memcpy(patch + (size_t)r * tw, gray + (size_t)(y + r) * fw + x, sizeof(float) * tw);
                orc_add_weighted(templ, patch, tw * th, 1 - lr, lr);
                free(patch);
                updated = 1;
            }
        } else {
            lost++;
        }
        rec->x = x; rec->y = y; rec->w = tw; rec->h = th; rec->conf = (float)best;
        rec->moved = moved; rec->updated = updated; rec->searched = searched; rec->lost_count = lost; rec->use_global = use_global;
    }
    if (templ_out) memcpy(templ_out, templ, sizeof(float) * (size_t)tw * th);
    free(map);
    free(gray);
    free(templ);
    return rc;
}
