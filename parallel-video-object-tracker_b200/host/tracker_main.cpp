// CLI twin of the reference driver for the NCC path: /root/reference/tracker/src/main.cpp:23-185.
//
// Same flags (--shared --const --const_tiled --batch=N; default naive), same banner and summary lines, same
// per-frame logic -- but the per-frame loop body (main.cpp:98-161: toGrayF32 -> NCC -> window clamp -> minMaxLoc ->
// gates -> addWeighted) runs on the GPU inside pvt_step, with the tracker state resident on the device.
// Differences forced by this image (no OpenCV C++, no video codecs):
//   * input is a raw BGR clip instead of ../data/car.mp4:   magic "PVTBGR1\n", int32 W, H, N, then N*H*W*3 bytes
//   * the ROI comes from --roi x,y,w,h instead of cv::selectROI (main.cpp:63; the reference has no default ROI)
//   * instead of an annotated .mp4 (no encoder here) the per-frame bbox/confidence goes to --out FILE as CSV, and --video-out FILE
//     writes the annotated frames -- cv::rectangle(frame, bbox, {0,255,0}, 2) of main.cpp:166, painted by pvt_draw_boxes with
//     OpenCV's pixel coverage -- as a raw clip in the input's own container ("PVTBGR1\n", W, H, N, frames)
//   * --gpu-formula (new) scores with the eps formula of the reference's CUDA kernels (pvt_formula) instead of the --cpu path's
//   * --cpu is rejected: the library has no CPU path (the reference's CPU mode is restated in oracle/, test-only)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "baseline_kernel.hpp"

// main.cpp:6-20
static const std::string NCC_MODE = "naive";
static const int BATCH_SIZE = 4;
static const int SEARCH_RADIUS_X = 80;
static const int SEARCH_RADIUS_Y = 80;
static const double NCC_MIN_CONFIDENCE = 0.40;
static const double NCC_STRONG_CONFIDENCE = 0.70;
static const double TEMPLATE_UPDATE_LR = 0.10;

int main(int argc, char** argv)
{
    std::string mode = NCC_MODE, input, out_csv, out_video;
    int batch = BATCH_SIZE;
    pvt::Rect bbox;
    bool have_roi = false, gpu_formula = false;
    for (int i = 1; i < argc; ++i) {
        std::string arg = argv[i];
        if (arg == "--cpu") mode = "cpu";
        else if (arg == "--shared") mode = "shared";
        else if (arg == "--const") mode = "const";
        else if (arg == "--const_tiled") mode = "const_tiled";
        else if (arg.rfind("--batch=", 0) == 0) { mode = "batch"; batch = std::max(1, std::atoi(arg.substr(8).c_str())); }
        else if (arg == "--roi" && i + 1 < argc) { have_roi = std::sscanf(argv[++i], "%d,%d,%d,%d", &bbox.x, &bbox.y, &bbox.width, &bbox.height) == 4; }
        else if (arg == "--out" && i + 1 < argc) out_csv = argv[++i];
        else if (arg == "--video-out" && i + 1 < argc) out_video = argv[++i];
        else if (arg == "--gpu-formula") gpu_formula = true;
        else if (arg[0] != '-') input = arg;
    }
    std::cout << "--------\nNCC Tracker Starting\nInput video : " << input << "\nMode        : " << mode << "\n";
    if (mode == "batch") std::cout << "Batch size  : " << batch << "\n";
    std::cout << "--------\n\n";
    if (mode == "cpu") { std::cerr << "--cpu is not available: libpvt has no CPU path (see oracle/ for the CPU restatement)\n"; return -1; }

    std::ifstream f(input, std::ios::binary);
    char magic[8];
    int32_t W = 0, H = 0, N = 0;
    if (!f || !f.read(magic, 8) || std::string(magic, 8) != "PVTBGR1\n" || !f.read((char*)&W, 4) || !f.read((char*)&H, 4) || !f.read((char*)&N, 4)) {
        std::cerr << "Cannot open video.\n";  // main.cpp:54
        return -1;
    }
    if (!have_roi || bbox.width == 0 || bbox.height == 0) { std::cerr << " No ROI selected.\n"; return -1; }  // main.cpp:66-69
    // page-locked frame buffer: the ROI ingest then reads just the search tile from it over PCIe (zero-copy) instead of
    // staging the whole frame through a pageable copy; falls back to a plain buffer if pinning fails
    struct FrameBuf {
        uint8_t* p = nullptr; bool pinned = false; size_t n = 0;
        explicit FrameBuf(size_t bytes) : n(bytes) { void* q = nullptr; if (pvt_alloc_pinned(&q, bytes) == PVT_OK) { p = (uint8_t*)q; pinned = true; } else p = new uint8_t[bytes]; }
        ~FrameBuf() { if (pinned) pvt_free_pinned(p); else delete[] p; }
        uint8_t* data() { return p; }
        size_t size() const { return n; }
    } frame((size_t)W * H * 3);
    if (N < 1 || !f.read((char*)frame.data(), frame.size())) return -1;

    pvt_params p;
    pvt_default_params(&p);
    p.search_radius_x = SEARCH_RADIUS_X; p.search_radius_y = SEARCH_RADIUS_Y;
    p.ncc_min_confidence = NCC_MIN_CONFIDENCE; p.ncc_strong_confidence = NCC_STRONG_CONFIDENCE; p.template_update_lr = TEMPLATE_UPDATE_LR;
    p.batch_size = batch;
    p.formula = gpu_formula ? PVT_FORMULA_EPS : PVT_FORMULA_CCOEFF_NORMED;
    p.mode = mode == "shared" ? PVT_MODE_SHARED : mode == "const" ? PVT_MODE_CONST : mode == "const_tiled" ? PVT_MODE_CONST_TILED
             : mode == "batch" ? PVT_MODE_BATCH : PVT_MODE_NAIVE;
    pvt_config cfg{};
    cfg.device = 0; cfg.frame_w = W; cfg.frame_h = H; cfg.max_streams = 1; cfg.max_tracks = 1;
    cfg.max_templ_w = bbox.width; cfg.max_templ_h = bbox.height;
    pvt_ctx* ctx = nullptr;
    try {
        pvt::check(pvt_create(&ctx, &p, &cfg));
        pvt_frame fr{0, PVT_FMT_BGR8, PVT_MEM_HOST, 0, frame.data(), (size_t)W * 3};
        pvt::check(pvt_track_init(ctx, 0, 0, &fr, bbox.x, bbox.y, bbox.width, bbox.height));  // main.cpp:70-71

        std::ofstream csv;
        if (!out_csv.empty()) { csv.open(out_csv); csv << "frame,x,y,w,h,conf,moved,updated,searched\n"; }
        std::ofstream vout;                                   // cv::VideoWriter writer(...)  main.cpp:76-82
        if (!out_video.empty()) {
            vout.open(out_video, std::ios::binary);
            if (!vout) { std::cerr << " Cannot open video writer.\n"; pvt_destroy(ctx); return -1; }
            const int32_t hdr[3] = {W, H, N - 1};
            vout.write("PVTBGR1\n", 8); vout.write((const char*)hdr, 12);
        }
        int frame_count = 0;
        double t_tot = 0.0;
        auto t_start = std::chrono::steady_clock::now();
        for (int k = 1; k < N; ++k) {  // while (cap >> frame)  main.cpp:93-96
            if (!f.read((char*)frame.data(), frame.size())) break;
            auto t1 = std::chrono::steady_clock::now();
            pvt_result r;
            pvt::check(pvt_step(ctx, 1, &fr, &r));  // main.cpp:98-161 on the GPU
            t_tot += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
            bbox.x = r.x; bbox.y = r.y;
            if (vout.is_open()) {                             // cv::rectangle(frame, bbox, {0,255,0}, 2); writer.write(frame);  main.cpp:166-167
                const int32_t box[4] = {r.x, r.y, r.w, r.h};
                pvt::check(pvt_draw_boxes(ctx, &fr, 1, box, nullptr));
                vout.write((const char*)frame.data(), (std::streamsize)frame.size());
            }
            if (csv.is_open()) csv << k << ',' << r.x << ',' << r.y << ',' << r.w << ',' << r.h << ',' << r.conf << ',' << (int)r.moved << ',' << (int)r.updated << ',' << (int)r.searched << "\n";
            frame_count++;
        }
        double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
        std::cout << "\n--------\n Tracking Complete\n Mode       : " << mode << "\n Frames     : " << frame_count << "\n Time (sec) : " << elapsed
                  << "\n Computation Time (sec)  : " << t_tot << "\n FPS        : " << frame_count / elapsed << "\n--------\n";  // main.cpp:175-182
        pvt_destroy(ctx);
    } catch (const pvt::Error& e) {
        std::cerr << e.what() << "\n";
        if (ctx) pvt_destroy(ctx);
        return -1;
    }
    return 0;
}
