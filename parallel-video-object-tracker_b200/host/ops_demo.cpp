// Exercises the reference-named map-level operators (baseline_kernel.hpp mirror) from C++:
//   ops_demo frame.f32 fw fh templ.f32 tw th out_prefix
// writes out_prefix.{naive,shared,const,const_tiled,batched0,batched1}.f32 ; exit code 3 if ncc_match_cpu did NOT throw.
#include <cstdio>
#include <fstream>
#include <iostream>

#include "baseline_kernel.hpp"

static pvt::Mat load(const char* path, int w, int h)
{
    pvt::Mat m(h, w, 4);
    std::ifstream f(path, std::ios::binary);
    if (!f.read((char*)m.data, (size_t)w * h * 4)) throw std::runtime_error(std::string("cannot read ") + path);
    return m;
}
static void save(const std::string& path, const pvt::Mat& m)
{
    std::ofstream f(path, std::ios::binary);
    for (int r = 0; r < m.rows; ++r) f.write((const char*)m.ptr<float>(r), (size_t)m.cols * 4);
}

int main(int argc, char** argv)
{
    if (argc != 8) { std::cerr << "usage: ops_demo frame.f32 fw fh templ.f32 tw th out_prefix\n"; return 2; }
    try {
        pvt::Mat frame = load(argv[1], atoi(argv[2]), atoi(argv[3])), templ = load(argv[4], atoi(argv[5]), atoi(argv[6]));
        const std::string out = argv[7];
        pvt::Mat m;
        baseline::ncc_match_naive_cuda(frame, templ, m);  save(out + ".naive.f32", m);
        baseline::ncc_match_shared_cuda(frame, templ, m); save(out + ".shared.f32", m);
        baseline::ncc_match_const(frame, templ, m);       save(out + ".const.f32", m);
        baseline::ncc_match_const_tiled(frame, templ, m); save(out + ".const_tiled.f32", m);
        // a strided view (cv::Mat ROI semantics): template cut out of a wider matrix
        pvt::Mat wide(templ.rows, templ.cols + 5, 4);
        for (int r = 0; r < templ.rows; ++r) std::memcpy(wide.ptr<float>(r), templ.ptr<float>(r), (size_t)templ.cols * 4);
        baseline::ncc_match_naive_cuda(frame, wide(pvt::Rect{0, 0, templ.cols, templ.rows}), m); save(out + ".view.f32", m);
        std::vector<pvt::Mat> frames{frame, frame.clone()}, maps;
        baseline::ncc_match_naive_cuda_batched(frames, templ, maps);
        save(out + ".batched0.f32", maps[0]); save(out + ".batched1.f32", maps[1]);
        bool threw = false;
        try { baseline::ncc_match_cpu(frame, templ, m); } catch (const pvt::Error& e) { threw = e.code == PVT_ERR_UNSUPPORTED; }
        if (!threw) return 3;
        try { baseline::ncc_match_naive_cuda(templ, frame, m); return 4; } catch (const pvt::Error& e) { if (e.code != PVT_ERR_INVALID) return 4; }
        std::cout << "ok " << m.rows << "x" << m.cols << "\n";
    } catch (const std::exception& e) { std::cerr << e.what() << "\n"; return 1; }
    return 0;
}
