// shim_demo.cpp -- drives the reference's call surface (FeatureExtractor::process, matchFeatures) through
// include/orbx_shim.hpp, exactly as src/main.cpp:41-51 + src/CameraPoseEstimator.cpp:409 would, and dumps the results
// so that tests/test_gpu_cpp_shim.py can compare them with the oracle.
//
//   shim_demo <w> <h> <nframes> <frames.raw> <out.bin> <nfeatures> <ratio>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "orbx_shim.hpp"

using namespace orbx_shim;

int main(int argc, char** argv)
{
    if (argc != 8) { std::fprintf(stderr, "usage: %s w h nframes frames.raw out.bin nfeatures ratio\n", argv[0]); return 2; }
    const int w = std::atoi(argv[1]), h = std::atoi(argv[2]), nframes = std::atoi(argv[3]), nfeatures = std::atoi(argv[6]);
    const float ratio = (float)std::atof(argv[7]);
    std::vector<uint8_t> raw((size_t)w * h * nframes);
    FILE* f = std::fopen(argv[4], "rb");
    if (!f || std::fread(raw.data(), 1, raw.size(), f) != raw.size()) { std::fprintf(stderr, "cannot read %s\n", argv[4]); return 2; }
    std::fclose(f);
    try {
        DataManager dm;
        dm.frames.resize((size_t)nframes);
        for (int i = 0; i < nframes; i++) dm.frames[(size_t)i].frameBuffer = Mat(h, w, raw.data() + (size_t)i * w * h);
        FeatureExtractor node(nfeatures);            // ORBSlam.addStage(new FeatureExtractor())
        node.init();
        FILE* o = std::fopen(argv[5], "wb");
        if (!o) return 2;
        for (int i = 0; i < nframes; i++) {          // ORBSlam.process(dm, i)
            if (!node.validationCheck(dm, i)) return 3;
            node.process(dm, i);
            const Features& ft = dm.frames[(size_t)i].features;
            int32_t n = (int32_t)ft.positions.size();
            std::fwrite(&n, 4, 1, o);
            std::fwrite(ft.positions.data(), sizeof(Point2d), (size_t)n, o);
            std::fwrite(ft.scales.data(), sizeof(double), (size_t)n, o);
            std::fwrite(ft.mapPointsIndices.data(), sizeof(int), (size_t)n, o);
            std::fwrite(ft.descriptors.data, 32, (size_t)n, o);
            if (i > 0) {                             // matchFeatures(desc_cur, desc_prev, matches, ratio)
                std::vector<DMatch> matches;
                matchFeatures(ft.descriptors, dm.frames[(size_t)i - 1].features.descriptors, matches, ratio);
                int32_t m = (int32_t)matches.size();
                std::fwrite(&m, 4, 1, o);
                std::fwrite(matches.data(), sizeof(DMatch), (size_t)m, o);
            }
        }
        std::fclose(o);
        node.destroy();
    } catch (const Error& e) {
        std::fprintf(stderr, "orbx error %d: %s\n", e.status, e.what());
        return 1;
    }
    return 0;
}
