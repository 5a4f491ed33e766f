// CLI twin of the reference's SECOND driver, the one with lost-object re-acquisition:
// /root/reference/tracker_ghc/src/main.cpp:57-285 (demo_tracker with --first: the template comes from the first frame).
//
// Same positional video argument and mode flags (--shared --const --const_tiled --batch=N; default "cuda"), same
// "Tracking mode:" and "Interactive tracking summary:" lines, same constants (:9-23) and the same per-frame logic
// (:145-239: NCC map -> local window OR whole map when lost -> confidence threshold by search kind -> move /
// lost_frame_count / use_global_search -> addWeighted) -- but that loop body runs on the GPU inside pvt_step with
// lost_frame_count and use_global_search resident on the device (pvt_params.lost_frame_threshold > 0).
// Differences forced by this image (no OpenCV C++, no video codecs, no display):
//   * input is a raw BGR clip:   magic "PVTBGR1\n", int32 W, H, N, then N*H*W*3 bytes (first frame = template frame)
//   * the ROI comes from --roi x,y,w,h instead of cv::selectROI (:116); constants may be overridden for tests with
//     --radius R, --lost N, --global C (the reference hard-codes them); --tc-global selects PVT_KERNEL_TC_GLOBAL
//   * the per-frame box goes to --out FILE as CSV instead of being drawn (:241)
//   * --cpu is rejected: the library has no CPU path (the CPU mode is restated in oracle/, test-only)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "baseline_kernel.hpp"

int main(int argc, char** argv)
{
    std::string video_path = (argc > 1) ? argv[1] : "data/car.mp4";   // :58
    std::string mode = "cuda", out_csv;
    int batch_size = 0;
    pvt::Rect roi;
    bool have_roi = false;
    pvt_params p;
    pvt_default_params_ghc(&p);   // :9-23
    for (int i = 2; i < argc; ++i) {   // :63-74
        std::string arg = argv[i];
        if (arg == "--cpu") mode = "cpu";
        else if (arg == "--shared") mode = "shared";
        else if (arg == "--const") mode = "const";
        else if (arg == "--const_tiled") mode = "const_tiled";
        else if (arg == "--record" || arg == "--first") {}   // no display / codec here: always the first frame, never a video file
        else if (arg.rfind("--batch=", 0) == 0) { mode = "batch"; batch_size = std::max(1, std::atoi(arg.substr(8).c_str())); }
        else if (arg == "--roi" && i + 1 < argc) { have_roi = std::sscanf(argv[++i], "%d,%d,%d,%d", &roi.x, &roi.y, &roi.width, &roi.height) == 4; }
        else if (arg == "--out" && i + 1 < argc) out_csv = argv[++i];
        else if (arg == "--radius" && i + 1 < argc) p.search_radius_x = p.search_radius_y = std::atoi(argv[++i]);
        else if (arg == "--lost" && i + 1 < argc) p.lost_frame_threshold = std::atoi(argv[++i]);
        else if (arg == "--global" && i + 1 < argc) p.ncc_global_confidence = std::atof(argv[++i]);
        else if (arg == "--gpu-formula") p.formula = PVT_FORMULA_EPS;   // score like the reference's CUDA kernels, not like --cpu
        else if (arg == "--tc-global") p.kernel = PVT_KERNEL_TC_GLOBAL; // whole-frame re-acquisition search (:186-193) on the tensor cores (8-bit sources)
    }
    if (mode == "cpu") { std::cerr << "--cpu is not available: libpvt has no CPU path (see oracle/ for the CPU restatement)\n"; return -1; }
    std::ifstream f(video_path, std::ios::binary);
    char magic[8];
    int32_t W = 0, H = 0, N = 0;
    if (!f || !f.read(magic, 8) || std::string(magic, 8) != "PVTBGR1\n" || !f.read((char*)&W, 4) || !f.read((char*)&H, 4) || !f.read((char*)&N, 4)) {
        std::cerr << "Cannot open video: " << video_path << std::endl;   // :84
        return -1;
    }
    // page-locked frame buffer: the ROI ingest then reads just the search tile from it over PCIe (zero-copy) instead of
    // staging the whole frame through a pageable copy; falls back to a plain buffer if pinning fails
    struct FrameBuf {
        uint8_t* p = nullptr; bool pinned = false; size_t n = 0;
        explicit FrameBuf(size_t bytes) : n(bytes) { void* q = nullptr; if (pvt_alloc_pinned(&q, bytes) == PVT_OK) { p = (uint8_t*)q; pinned = true; } else p = new uint8_t[bytes]; }
        ~FrameBuf() { if (pinned) pvt_free_pinned(p); else delete[] p; }
        uint8_t* data() { return p; }
        size_t size() const { return n; }
    } frame((size_t)W * H * 3);
    if (N < 1 || !f.read((char*)frame.data(), frame.size())) { std::cerr << "Cannot read first frame from video." << std::endl; return -1; }   // :91
    std::cout << "Select template from the first frame.\n";   // :94
    if (!have_roi || roi.width == 0 || roi.height == 0) { std::cerr << "No template selected" << std::endl; return -1; }   // :117-120
    if (p.lost_frame_threshold < 1) { std::cerr << "--lost must be >= 1\n"; return -1; }

    // the reference's batch mode (:155 falls through to the naive kernel per frame in this driver) searches every frame
    p.mode = mode == "shared" ? PVT_MODE_SHARED : mode == "const" ? PVT_MODE_CONST : mode == "const_tiled" ? PVT_MODE_CONST_TILED : PVT_MODE_NAIVE;
    (void)batch_size;
    pvt_config cfg{};
    cfg.device = 0; cfg.frame_w = W; cfg.frame_h = H; cfg.max_streams = 1; cfg.max_tracks = 1;
    cfg.max_templ_w = roi.width; cfg.max_templ_h = roi.height;
    pvt_ctx* ctx = nullptr;
    try {
        pvt::check(pvt_create(&ctx, &p, &cfg));
        pvt_frame fr{0, PVT_FMT_BGR8, PVT_MEM_HOST, 0, frame.data(), (size_t)W * 3};
        pvt::check(pvt_track_init(ctx, 0, 0, &fr, roi.x, roi.y, roi.width, roi.height));   // :122-123
        std::cout << "Tracking mode: " << mode << std::endl;   // :131
        std::ofstream csv;
        if (!out_csv.empty()) { csv.open(out_csv); csv << "frame,x,y,w,h,conf,moved,updated,searched,lost_frame_count,use_global_search\n"; }
        int total_frames = 0;
        auto t_start = std::chrono::steady_clock::now();
        for (int k = 1; k < N; ++k) {   // while (true) { if (!cap.read(frame)) break;   :145-147
            if (!f.read((char*)frame.data(), frame.size())) break;
            pvt_result r;
            pvt::check(pvt_step(ctx, 1, &fr, &r));   // :149-239 on the GPU
            if (csv.is_open()) {
                int lost = 0, glob = 0;
                pvt::check(pvt_get_lost_state(ctx, 0, &lost, &glob));
                csv << k << ',' << r.x << ',' << r.y << ',' << r.w << ',' << r.h << ',' << r.conf << ',' << (int)r.moved << ',' << (int)r.updated << ','
                    << (int)r.searched << ',' << lost << ',' << glob << "\n";
            }
            total_frames++;   // :249
        }
        double time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
        double avg_fps = (time > 0.0) ? (double(total_frames) / time) : 0.0;
        std::cout << "Interactive tracking summary: " << "frames=" << total_frames << ", " << "time=" << time << " s, " << "FPS=" << avg_fps << std::endl;   // :277-281
        pvt_destroy(ctx);
    } catch (const pvt::Error& e) {
        std::cerr << e.what() << "\n";
        if (ctx) pvt_destroy(ctx);
        return -1;
    }
    return 0;
}
