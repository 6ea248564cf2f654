// venusaur_headless.cpp -- headless replacement for the frame loop of the reference's Core/Core.cpp:239-430
// (SURVEY 8f rank 1): same objects, same call sequence (Scene, Camera, Renderer::Init, CUDAOutputBuffer,
// Renderer::Draw per frame), but no window / GL / ImGui: the frame is read back with
// CUDAOutputBuffer::getHostPointer() (CUDAOutputBuffer.h:348-372) and written as a PPM, stats go to stdout as JSON.
//
//   g++ -std=c++17 -Iinclude tools/venusaur_headless.cpp -o venusaur_headless -Lvenusaur_b200 -lvenusaur_b200 -Wl,-rpath,$PWD/venusaur_b200
//   ./venusaur_headless --width 1200 --height 800 --frames 64 --max-depth 50 --out frame.ppm
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "venusaur/Renderer.h"

int main(int argc, char** argv) {
    int width = 1200, height = 800, frames = 16, max_depth = 4, spp = 16, device = 0;   // Core.cpp:27-28, Renderer.h:53, RayTracer.cu:172
    std::string out = "frame.ppm";
    bool exact = false, wavefront = false;
    std::vector<std::pair<std::string, double>> options;
    for (int i = 1; i < argc; i++) {
        auto next = [&](const char* name) { if (i + 1 >= argc) { fprintf(stderr, "%s needs a value\n", name); exit(2); } return argv[++i]; };
        if (!strcmp(argv[i], "--width")) width = atoi(next("--width"));
        else if (!strcmp(argv[i], "--height")) height = atoi(next("--height"));
        else if (!strcmp(argv[i], "--frames")) frames = atoi(next("--frames"));
        else if (!strcmp(argv[i], "--max-depth")) max_depth = atoi(next("--max-depth"));
        else if (!strcmp(argv[i], "--spp")) spp = atoi(next("--spp"));
        else if (!strcmp(argv[i], "--device")) device = atoi(next("--device"));
        else if (!strcmp(argv[i], "--out")) out = next("--out");
        else if (!strcmp(argv[i], "--exact")) exact = true;
        else if (!strcmp(argv[i], "--wavefront")) wavefront = true;
        else if (!strcmp(argv[i], "--opt")) {                      // library tuning knob name=value (include/venusaur_b200.h)
            const std::string kv = next("--opt");
            const size_t eq = kv.find('=');
            if (eq == std::string::npos) { fprintf(stderr, "--opt needs name=value\n"); return 2; }
            options.emplace_back(kv.substr(0, eq), atof(kv.c_str() + eq + 1));
        }
        else { fprintf(stderr, "usage: %s [--width W] [--height H] [--frames N] [--spp S] [--max-depth D] [--device I] [--exact] [--wavefront] [--opt name=value]... [--out file.ppm]\n", argv[0]); return 2; }
    }
    try {
        // Core.cpp:21-31
        const venusaur::vec3 lookfrom(13, 2, 3), lookat(0, 0, 0);
        Camera camera(lookfrom, 20.0f, static_cast<float>(width) / static_cast<float>(height), 0.1f, 10.0f);
        Scene scene;
        Renderer renderer;
        renderer.SetDevice(device);
        renderer.SetMaxDepth(static_cast<uint32_t>(max_depth));
        for (const auto& kv : options) renderer.SetOption(kv.first, kv.second);
        renderer.SetSamplesPerPixel(static_cast<uint32_t>(spp));
        uint32_t flags = 0;
        if (exact) flags |= VN_EXACT;
        if (wavefront) flags |= VN_WAVEFRONT;
        renderer.SetFlags(flags);
        renderer.Init(scene, "");                                                   // Core.cpp:246
        CUDAOutputBuffer<uchar4> output_buffer(CUDAOutputBufferType::CUDA_DEVICE, width, height, device);   // Core.cpp:252 (+ the device)
        camera.SetForward(lookat - lookfrom);                                       // Core.cpp:355

        unsigned long long segments = 0;
        double kernel_ms = 0.0;
        const auto t0 = std::chrono::steady_clock::now();
        for (int f = 0; f < frames; f++) {                                          // Core.cpp:358-430 without the window
            renderer.Draw(camera, output_buffer);                                   // Core.cpp:394
            const vn_stats st = renderer.Stats();
            segments += st.segments;
            kernel_ms += st.ms_render;
        }
        const uchar4* px = output_buffer.getHostPointer();                          // replaces getPBO + GL upload, Core.cpp:396
        const double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

        if (!out.empty()) {
            FILE* fp = fopen(out.c_str(), "wb");
            if (!fp) { fprintf(stderr, "cannot write %s\n", out.c_str()); return 1; }
            fprintf(fp, "P6\n%d %d\n255\n", width, height);
            for (int y = height - 1; y >= 0; y--)                                   // row 0 is the bottom of the picture (Core.cpp:141)
                for (int x = 0; x < width; x++) fwrite(&px[static_cast<size_t>(y) * width + x], 1, 3, fp);
            fclose(fp);
        }
        printf("{\"width\": %d, \"height\": %d, \"frames\": %d, \"spp_per_frame\": %d, \"max_depth\": %d, \"spheres\": %zu, "
               "\"segments\": %llu, \"kernel_ms\": %.3f, \"wall_ms\": %.3f, \"mrays_per_s_kernel\": %.1f, \"mrays_per_s_wall\": %.1f, \"image\": \"%s\"}\n",
               width, height, frames, spp, max_depth, scene.m_spheres.size(), segments, kernel_ms, wall_ms,
               segments / (kernel_ms * 1e3), segments / (wall_ms * 1e3), out.c_str());
        renderer.Cleanup();                                                         // Core.cpp end of main
    } catch (const Exception& e) {
        fprintf(stderr, "venusaur_headless: %s\n", e.what());
        return 1;
    }
    return 0;
}
