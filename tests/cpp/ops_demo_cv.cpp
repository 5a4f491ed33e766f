// TEST ONLY.  The `-DPVT_WITH_OPENCV` branch of host/baseline_kernel.hpp: the six baseline::ncc_match_* operators on cv::Mat,
// signature-for-signature what /root/reference/tracker/include/baseline_kernel.hpp:8-17 declares and main.cpp:103-133 calls.
// This image has no OpenCV C++; the build uses the minimal <opencv2/core.hpp> stand-in under oracle/ref_build/ (the same one
// the reference's own baseline_kernel.cu compiles against), so the branch is at least compiled and run in CI.
//   ops_demo_cv frame.f32 fw fh templ.f32 tw th out_prefix
#include <cstdio>
#include <fstream>
#include <iostream>

#include "baseline_kernel.hpp"

static cv::Mat load(const char* path, int w, int h)
{
    cv::Mat m(h, w, CV_32FC1);
    std::ifstream f(path, std::ios::binary);
    if (!f.read((char*)m.data, (size_t)w * h * 4)) throw std::runtime_error(std::string("cannot read ") + path);
    return m;
}
static void save(const std::string& path, const cv::Mat& m)
{
    std::ofstream f(path, std::ios::binary);
    for (int r = 0; r < m.rows; ++r) f.write((const char*)m.ptr<float>(r), (size_t)m.cols * 4);
}

// the exact call shapes of the reference (compile-time check of the signatures)
static void (*const k_ops[])(const cv::Mat&, const cv::Mat&, cv::Mat&) = {baseline::ncc_match_naive_cuda, baseline::ncc_match_shared_cuda,
                                                                           baseline::ncc_match_cpu, baseline::ncc_match_const,
                                                                           baseline::ncc_match_const_tiled};
static void (*const k_batched)(const std::vector<cv::Mat>&, const cv::Mat&, std::vector<cv::Mat>&) = baseline::ncc_match_naive_cuda_batched;

int main(int argc, char** argv)
{
    if (argc == 2 && std::string(argv[1]) == "--signatures") { std::cout << sizeof(k_ops) / sizeof(k_ops[0]) + (k_batched ? 1 : 0) << "\n"; return 0; }
    if (argc != 8) { std::cerr << "usage: ops_demo_cv frame.f32 fw fh templ.f32 tw th out_prefix\n"; return 2; }
    try {
        cv::Mat frame = load(argv[1], atoi(argv[2]), atoi(argv[3])), templ = load(argv[4], atoi(argv[5]), atoi(argv[6]));
        const std::string out = argv[7];
        cv::Mat m;
        baseline::ncc_match_naive_cuda(frame, templ, m);  save(out + ".naive.f32", m);
        baseline::ncc_match_const_tiled(frame, templ, m); save(out + ".const_tiled.f32", m);
        std::vector<cv::Mat> frames{frame, frame}, maps;
        baseline::ncc_match_naive_cuda_batched(frames, templ, maps);
        save(out + ".batched1.f32", maps[1]);
        bool threw = false;
        try { baseline::ncc_match_cpu(frame, templ, m); } catch (const pvt::Error& e) { threw = e.code == PVT_ERR_UNSUPPORTED; }
        if (!threw) return 3;
        std::cout << "ok " << m.rows << "x" << m.cols << "\n";
    } catch (const std::exception& e) { std::cerr << e.what() << "\n"; return 1; }
    return 0;
}
