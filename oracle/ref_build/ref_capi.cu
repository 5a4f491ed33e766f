// TEST / BENCH INFRASTRUCTURE ONLY -- C ABI over the reference's OWN GPU operators, compiled from
// /root/reference/tracker/src/baseline_kernel.cu as it lies there (see Makefile).  Used by tests/ to pin
// PVT_FORMULA_EPS against the real kernels and by bench.py's `ref_gpu_baseline` leg.  Never linked into libpvt.
#include "baseline_kernel.hpp"   // /root/reference/tracker/include, resolved against the shim <opencv2/opencv.hpp>

#include <cstdio>
#include <vector>

extern "C" {

// mode: 0 naive, 2 shared, 3 const, 4 const_tiled (pvt_mode numbering).  Dense row-major buffers.  0 on success, -1 when the
// reference asserts (e.g. template > 4096 px in the const modes, baseline_kernel.cu:500).
__attribute__((visibility("default"))) int ref_ncc_match(int mode, const float* frame, int fw, int fh, const float* templ, int tw, int th, float* out)
{
    try {
        cv::Mat f(fh, fw, CV_32FC1, (void*)frame), t(th, tw, CV_32FC1, (void*)templ), o;
        switch (mode) {
            case 0: baseline::ncc_match_naive_cuda(f, t, o); break;
            case 2: baseline::ncc_match_shared_cuda(f, t, o); break;
            case 3: baseline::ncc_match_const(f, t, o); break;
            case 4: baseline::ncc_match_const_tiled(f, t, o); break;
            default: return -2;
        }
        std::memcpy(out, o.ptr<float>(), (size_t)o.rows * o.cols * sizeof(float));
        return 0;
    } catch (const cv::Exception& e) {
        std::fprintf(stderr, "[ref] %s\n", e.what());
        return -1;
    }
}

__attribute__((visibility("default"))) int ref_ncc_match_batched(int n, const float* const* frames, int fw, int fh, const float* templ, int tw, int th, float* const* outs)
{
    try {
        std::vector<cv::Mat> fs, os;
        for (int i = 0; i < n; ++i) fs.emplace_back(fh, fw, CV_32FC1, (void*)frames[i]);
        cv::Mat t(th, tw, CV_32FC1, (void*)templ);
        baseline::ncc_match_naive_cuda_batched(fs, t, os);
        for (int i = 0; i < n; ++i) std::memcpy(outs[i], os[i].ptr<float>(), (size_t)os[i].rows * os[i].cols * sizeof(float));
        return 0;
    } catch (const cv::Exception& e) {
        std::fprintf(stderr, "[ref] %s\n", e.what());
        return -1;
    }
}

}  // extern "C"
