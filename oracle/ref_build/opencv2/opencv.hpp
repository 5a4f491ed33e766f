// TEST INFRASTRUCTURE ONLY.  The smallest stand-in for <opencv2/opencv.hpp> that lets the reference's own CUDA source
//   /root/reference/tracker/src/baseline_kernel.cu   (five NCC kernels :21-304, host wrappers :311-596)
// compile UNMODIFIED in an image without OpenCV C++ headers: the file only uses cv::Mat::{rows, cols, type, ptr<T>, create},
// cv::Scalar::operator[], cv::meanStdDev and CV_Assert / CV_32FC1 (grep "cv::\|CV_" in that file).  Nothing here is copied
// from OpenCV or from the reference; cv::meanStdDev restates OpenCV's published semantics for a CV_32FC1 array (double
// accumulation of sum and sum of squares, mean = s/N, stddev = sqrt(max(sq/N - mean^2, 0))), the same statement
// oracle/ncc_oracle.c:orc_mean_stddev makes and tests/test_oracle_golden.py pins bit-exactly against cv2 4.13.0.
// Also serves `-DPVT_WITH_OPENCV` compile checks of parallel-video-object-tracker_b200/host/baseline_kernel.hpp.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_32F 5
#define CV_32FC1 5
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_Assert(expr)                                                                                   \
    do {                                                                                                  \
        if (!(expr)) throw cv::Exception(std::string("CV_Assert failed: ") + #expr);                      \
    } while (0)

namespace cv {

class Exception : public std::runtime_error {
public:
    explicit Exception(const std::string& m) : std::runtime_error(m) {}
};

struct Scalar {
    double val[4] = {0, 0, 0, 0};
    double& operator[](int i) { return val[i]; }
    const double& operator[](int i) const { return val[i]; }
};

// dense row-major matrix, owning or borrowing (a header over caller memory, like cv::Mat(rows, cols, type, data, step))
class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;           // bytes per row
    unsigned char* data = nullptr;

    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* p, size_t step_bytes = 0) : rows(r), cols(c), data((unsigned char*)p), type_(type)
    {
        step = step_bytes ? step_bytes : (size_t)c * elem(type);
    }
    Mat(const Mat& o) { *this = o; }
    Mat& operator=(const Mat& o)
    {
        if (this == &o) return *this;
        rows = o.rows; cols = o.cols; step = o.step; type_ = o.type_; own_ = o.own_;
        data = own_.empty() ? o.data : own_.data();
        return *this;
    }
    void create(int r, int c, int type)
    {
        if (r == rows && c == cols && type == type_ && data) return;
        rows = r; cols = c; type_ = type; step = (size_t)c * elem(type);
        own_.assign((size_t)r * step, 0);
        data = own_.data();
    }
    int type() const { return type_; }
    int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == (size_t)cols * elem(type_); }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }

private:
    static size_t elem(int type) { return type == CV_32FC1 ? 4 : type == CV_8UC3 ? 3 : 1; }
    int type_ = CV_32FC1;
    std::vector<unsigned char> own_;
};

inline void meanStdDev(const Mat& m, Scalar& mean, Scalar& stddev)
{
    CV_Assert(m.type() == CV_32FC1);
    double s = 0.0, q = 0.0;
    for (int r = 0; r < m.rows; ++r) {
        const float* p = m.ptr<float>(r);
        for (int c = 0; c < m.cols; ++c) {
            const double v = (double)p[c];
            s += v;
            q += v * v;
        }
    }
    const double scale = 1.0 / ((double)m.rows * (double)m.cols);
    const double mu = s * scale;
    double var = q * scale - mu * mu;
    if (var < 0.0) var = 0.0;
    mean[0] = mu;
    stddev[0] = std::sqrt(var);
}

}  // namespace cv
