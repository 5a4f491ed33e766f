// Host-side mirror of the reference's operator interface for the NCC path, on top of the C ABI (include/pvt.h).
//
//   reference: /root/reference/tracker/include/baseline_kernel.hpp:6-18  (namespace baseline, six ncc_match_* operators)
//              /root/reference/tracker/include/utils.hpp:5-14            (toGrayF32)
//
// Same names, argument order, argument meaning and failure behaviour: inputs must be single-channel float32, the
// frame must be at least as large as the template (the reference CV_Asserts: ncc_cpu.cpp:7-10,
// baseline_kernel.cu:315-325), the output is (re)allocated by the callee as (fh-th+1) x (fw-tw+1) float32
// (baseline_kernel.cu:327) and the call is synchronous.  Violations throw pvt::Error (the reference throws
// cv::Exception); CUDA failures throw too instead of calling std::exit (baseline_kernel.cu:12-18).
//
// With -DPVT_WITH_OPENCV the operators take cv::Mat exactly like the reference (drop-in for main.cpp:103-133);
// without OpenCV (this image has no OpenCV C++ headers) they take pvt::Mat, a minimal row-major matrix view with
// cv::Mat's fields (rows, cols, step, data).  ncc_match_cpu is declared for source compatibility and throws: the
// library has no CPU path (the CPU implementation lives in oracle/, test-only).
//
// Values: by default every operator returns the map of the reference's CPU operator (cv::matchTemplate TM_CCOEFF_NORMED,
// the parity target).  With -DPVT_BASELINE_GPU_FORMULA the GPU-named operators return the eps formula of the reference's own
// CUDA kernels instead (baseline_kernel.cu:44-49,62; pvt_formula in pvt.h) for callers that depend on those numbers.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pvt.h"

#ifdef PVT_WITH_OPENCV
#include <opencv2/core.hpp>
#endif

namespace pvt {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("pvt error " + std::to_string(c) + ": " + m), code(c) {}
};
inline void check(int rc)
{
    if (rc < 0) throw Error(rc, pvt_last_error());
}

// cv::Rect stand-in (main.cpp:63)
struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
};

// Minimal cv::Mat stand-in: owns (or views) a row-major 2-D array of `elem` bytes per pixel.
struct Mat {
    int rows = 0, cols = 0, elem = 4;  // elem: 4 = CV_32FC1, 1 = CV_8UC1, 3 = CV_8UC3
    size_t step = 0;                   // bytes per row
    uint8_t* data = nullptr;
    std::shared_ptr<std::vector<uint8_t>> own;

    Mat() = default;
    Mat(int r, int c, int e = 4) { create(r, c, e); }
    Mat(int r, int c, int e, void* p, size_t s) : rows(r), cols(c), elem(e), step(s), data((uint8_t*)p) {}
    void create(int r, int c, int e = 4)
    {
        if (own && rows == r && cols == c && elem == e) return;
        rows = r; cols = c; elem = e; step = (size_t)c * e;
        own = std::make_shared<std::vector<uint8_t>>(step * (size_t)r);
        data = own->data();
    }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    Mat operator()(const Rect& r) const { Mat m(r.height, r.width, elem, data + (size_t)r.y * step + (size_t)r.x * elem, step); m.own = own; return m; }
    Mat clone() const
    {
        Mat m(rows, cols, elem);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elem);
        return m;
    }
};

#ifdef PVT_WITH_OPENCV
using MatArg = cv::Mat;
inline bool is_f32c1(const cv::Mat& m) { return m.type() == CV_32FC1; }
inline void create_f32(cv::Mat& m, int r, int c) { m.create(r, c, CV_32FC1); }
inline const float* fptr(const cv::Mat& m) { return m.ptr<float>(); }
inline float* fptr(cv::Mat& m) { return m.ptr<float>(); }
inline size_t step_of(const cv::Mat& m) { return m.step; }
#else
using MatArg = Mat;
inline bool is_f32c1(const Mat& m) { return m.elem == 4; }
inline void create_f32(Mat& m, int r, int c) { m.create(r, c, 4); }
inline const float* fptr(const Mat& m) { return m.ptr<float>(); }
inline float* fptr(Mat& m) { return m.ptr<float>(); }
inline size_t step_of(const Mat& m) { return m.step; }
#endif

inline int& default_device()
{
    static int d = 0;
    return d;
}

#ifdef PVT_BASELINE_GPU_FORMULA
constexpr int kFormula = PVT_FORMULA_EPS;
#else
constexpr int kFormula = PVT_FORMULA_CCOEFF_NORMED;
#endif

inline void ncc_match_mode(int mode, const MatArg& frame, const MatArg& templ, MatArg& ncc_map)
{
    if (kFormula == PVT_FORMULA_EPS) mode |= PVT_MODE_FLAG_EPS;
    if (!is_f32c1(frame) || !is_f32c1(templ)) throw Error(PVT_ERR_INVALID, "CV_32FC1 inputs required (ncc_cpu.cpp:7-8)");
    if (frame.cols < templ.cols || frame.rows < templ.rows) throw Error(PVT_ERR_INVALID, "frame smaller than template (ncc_cpu.cpp:9-10)");
    create_f32(ncc_map, frame.rows - templ.rows + 1, frame.cols - templ.cols + 1);
    check(pvt_ncc_match(default_device(), mode, fptr(frame), frame.cols, frame.rows, step_of(frame), fptr(templ), templ.cols, templ.rows,
                        step_of(templ), fptr(ncc_map), step_of(ncc_map)));
}

}  // namespace pvt

namespace baseline {

using Mat = pvt::MatArg;

// baseline_kernel.hpp:8
inline void ncc_match_naive_cuda(const Mat& frame_gray_f32, const Mat& templ_gray_f32, Mat& ncc_map) { pvt::ncc_match_mode(PVT_MODE_NAIVE, frame_gray_f32, templ_gray_f32, ncc_map); }
// baseline_kernel.hpp:10
inline void ncc_match_shared_cuda(const Mat& frame_gray_f32, const Mat& templ_gray_f32, Mat& ncc_map) { pvt::ncc_match_mode(PVT_MODE_SHARED, frame_gray_f32, templ_gray_f32, ncc_map); }
// baseline_kernel.hpp:12 -- CPU operator: not part of this library (throws PVT_ERR_UNSUPPORTED)
inline void ncc_match_cpu(const Mat& frame_gray_f32, const Mat& templ_gray_f32, Mat& ncc_map) { pvt::ncc_match_mode(PVT_MODE_CPU, frame_gray_f32, templ_gray_f32, ncc_map); }
// baseline_kernel.hpp:16
inline void ncc_match_const(const Mat& frame_gray_f32, const Mat& templ_gray_f32, Mat& ncc_map) { pvt::ncc_match_mode(PVT_MODE_CONST, frame_gray_f32, templ_gray_f32, ncc_map); }
// baseline_kernel.hpp:17
inline void ncc_match_const_tiled(const Mat& frame_gray_f32, const Mat& templ_gray_f32, Mat& ncc_map) { pvt::ncc_match_mode(PVT_MODE_CONST_TILED, frame_gray_f32, templ_gray_f32, ncc_map); }

// baseline_kernel.hpp:14: all frames share one geometry (baseline_kernel.cu:419-422), one map per frame
inline void ncc_match_naive_cuda_batched(const std::vector<Mat>& frames_gray_f32, const Mat& templ_gray_f32, std::vector<Mat>& ncc_maps)
{
    if (frames_gray_f32.empty()) throw pvt::Error(PVT_ERR_INVALID, "empty batch (baseline_kernel.cu:412)");
    const int fw = frames_gray_f32[0].cols, fh = frames_gray_f32[0].rows;
    if (!pvt::is_f32c1(templ_gray_f32)) throw pvt::Error(PVT_ERR_INVALID, "CV_32FC1 template required");
    if (fw < templ_gray_f32.cols || fh < templ_gray_f32.rows) throw pvt::Error(PVT_ERR_INVALID, "frame smaller than template");
    ncc_maps.resize(frames_gray_f32.size());
    std::vector<const float*> in;
    std::vector<float*> out;
    size_t fstep = pvt::step_of(frames_gray_f32[0]);
    for (size_t i = 0; i < frames_gray_f32.size(); ++i) {
        const Mat& f = frames_gray_f32[i];
        if (!pvt::is_f32c1(f) || f.cols != fw || f.rows != fh || pvt::step_of(f) != fstep)
            throw pvt::Error(PVT_ERR_INVALID, "all frames must be CV_32FC1 of one geometry (baseline_kernel.cu:419-422)");
        pvt::create_f32(ncc_maps[i], fh - templ_gray_f32.rows + 1, fw - templ_gray_f32.cols + 1);
        in.push_back(pvt::fptr(f));
        out.push_back(pvt::fptr(ncc_maps[i]));
    }
    pvt::check(pvt_ncc_match_batched_f(pvt::default_device(), pvt::kFormula, (int)in.size(), in.data(), fw, fh, fstep, pvt::fptr(templ_gray_f32), templ_gray_f32.cols,
                                     templ_gray_f32.rows, pvt::step_of(templ_gray_f32), out.data(), pvt::step_of(ncc_maps[0])));
}

}  // namespace baseline
