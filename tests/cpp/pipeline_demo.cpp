// pipeline_demo.cpp -- the fast path in the reference's host language: a loaded sequence goes through
// orbx_shim::SequenceFrontEnd (orbx_submit_batch / orbx_wait_batch, pinned staging via orbx_host_alloc) and the results are
// dumped in the same format as shim_demo.cpp, so the test can hold both against the oracle.
//
//   pipeline_demo <w> <h> <nframes> <frames.raw> <out.bin> <nfeatures> <ratio> <batch>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "orbx_shim.hpp"

using namespace orbx_shim;

int main(int argc, char** argv)
{
    if (argc != 9 && argc != 10) { std::fprintf(stderr, "usage: %s w h nframes frames.raw out.bin nfeatures ratio batch [filter]\n", argv[0]); return 2; }
    const bool filter = argc == 10;        // also run computeFundamentalMatrix on every (frame, predecessor) pair
    const int w = std::atoi(argv[1]), h = std::atoi(argv[2]), nframes = std::atoi(argv[3]), nfeatures = std::atoi(argv[6]);
    const float ratio = (float)std::atof(argv[7]);
    const int batch = std::atoi(argv[8]);
    std::vector<uint8_t> raw((size_t)w * h * nframes);
    FILE* f = std::fopen(argv[4], "rb");
    if (!f || std::fread(raw.data(), 1, raw.size(), f) != raw.size()) { std::fprintf(stderr, "cannot read %s\n", argv[4]); return 2; }
    std::fclose(f);
    try {
        DataManager dm;
        dm.frames.resize((size_t)nframes);
        for (int i = 0; i < nframes; i++) dm.frames[(size_t)i].frameBuffer = Mat(h, w, raw.data() + (size_t)i * w * h);
        SequenceFrontEnd fe(nfeatures, ratio, batch, w, h);
        if (filter) fe.enableFilter(3., 0.85);
        std::vector<std::vector<unsigned char> > status, st1, st2;
        std::vector<std::vector<double> > F, F1, F2;
        std::vector<std::vector<DMatch> > matches;
        // two calls: the second one's first frame must be matched against the first one's last frame
        const int half = nframes / 2;
        std::vector<std::vector<DMatch> > m1, m2;
        fe.process(dm, 0, half, m1, st1, F1);
        fe.process(dm, half, nframes - half, m2, st2, F2);
        matches = m1;
        matches.insert(matches.end(), m2.begin(), m2.end());
        status = st1; status.insert(status.end(), st2.begin(), st2.end());
        F = F1; F.insert(F.end(), F2.begin(), F2.end());
        FILE* o = std::fopen(argv[5], "wb");
        if (!o) return 2;
        for (int i = 0; i < nframes; i++) {
            const Features& ft = dm.frames[(size_t)i].features;
            int32_t n = (int32_t)ft.positions.size();
            std::fwrite(&n, 4, 1, o);
            std::fwrite(ft.positions.data(), sizeof(Point2d), (size_t)n, o);
            std::fwrite(ft.scales.data(), sizeof(double), (size_t)n, o);
            std::fwrite(ft.mapPointsIndices.data(), sizeof(int), (size_t)n, o);
            std::fwrite(ft.descriptors.data, 32, (size_t)n, o);
            if (i > 0) {
                int32_t m = (int32_t)matches[(size_t)i].size();
                std::fwrite(&m, 4, 1, o);
                std::fwrite(matches[(size_t)i].data(), sizeof(DMatch), (size_t)m, o);
                if (filter) {
                    std::fwrite(status[(size_t)i].data(), 1, (size_t)m, o);
                    std::fwrite(F[(size_t)i].data(), sizeof(double), 9, o);
                }
            }
        }
        std::fclose(o);
    } catch (const Error& e) {
        std::fprintf(stderr, "orbx error %d: %s\n", e.status, e.what());
        return 1;
    }
    return 0;
}
