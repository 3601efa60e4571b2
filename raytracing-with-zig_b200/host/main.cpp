// rtz_main — the executable of reference src/main.zig:14-36 / build.zig:16-25 on the B200 path.
//   rtz_main [-DimgWidth=N] [-DsamplesPerPixel=N] [-DfileName=NAME] [-Dseed=N] [-DnumGpus=N]
// (the reference takes these as `zig build -D...` options; defaults 3840 / 500 / chapter14.ppm / none).
// Writes images/<fileName> relative to the working directory, which must exist (Q18).
#include <cstdlib>
#include <iostream>

#include "rtz_host.hpp"

int main(int argc, char** argv) {
    auto& cfg = rtz::config();
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&](const char* k) -> const char* {
            const std::string p = std::string("-D") + k + "=";
            return a.compare(0, p.size(), p) == 0 ? argv[i] + p.size() : nullptr;
        };
        if (const char* v = val("imgWidth")) cfg.imgWidth = std::strtoull(v, nullptr, 0);
        else if (const char* v = val("samplesPerPixel")) cfg.samplesPerPixel = std::strtoull(v, nullptr, 0);
        else if (const char* v = val("fileName")) cfg.fileName = v;
        else if (const char* v = val("seed")) cfg.seed = std::strtoull(v, nullptr, 0);
        else if (const char* v = val("numGpus")) cfg.numGpus = (int32_t)std::strtol(v, nullptr, 0);
        else {
            std::cerr << "unknown option " << a << "\n";
            return 2;
        }
    }
    try {
        rtz_stats st;
        rtz::mainRender(&st);
        std::cerr << "Done. " << st.samples << " samples, " << st.segments << " segments, " << st.sphere_tests
                  << " ray-sphere tests, trace " << st.trace_ms << " ms on " << st.gpus << " GPU(s)\n";
    } catch (const rtz::RenderFailed& e) {
        std::cerr << "error.RenderFailed: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
