// fmat_demo -- drives orbx_shim::computeFundamentalMatrix the way CameraPoseEstimator does (src/CameraPoseEstimator.cpp:291,419):
// positions of two frames and a DMatch list in, status + inlier positions + F out.
// usage: fmat_demo n points.bin out.bin      points.bin: n x (x1, y1, x2, y2) doubles; match i = (queryIdx n-1-i, trainIdx i)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "orbx_shim.hpp"

using namespace orbx_shim;

int main(int argc, char** argv)
{
    if (argc != 4) { std::fprintf(stderr, "usage: %s n points.bin out.bin\n", argv[0]); return 2; }
    const int n = std::atoi(argv[1]);
    std::vector<double> raw((size_t)n * 4);
    FILE* f = std::fopen(argv[2], "rb");
    if (!f || std::fread(raw.data(), sizeof(double), raw.size(), f) != raw.size()) { std::fprintf(stderr, "cannot read %s\n", argv[2]); return 2; }
    std::fclose(f);
    std::vector<Point2d> positions1((size_t)n), positions2((size_t)n);
    std::vector<DMatch> matches((size_t)n);
    for (int i = 0; i < n; i++) {
        const Point2d a = {raw[(size_t)4 * i], raw[(size_t)4 * i + 1]}, b = {raw[(size_t)4 * i + 2], raw[(size_t)4 * i + 3]};
        positions1[(size_t)(n - 1 - i)] = a;            // the keypoint tables are not in match order
        positions2[(size_t)i] = b;
        DMatch m = {n - 1 - i, i, 0, 0.f};
        matches[(size_t)i] = m;
    }
    try {
        std::vector<Point2d> in1, in2;
        std::vector<unsigned char> status;
        double F[9];
        computeFundamentalMatrix(positions1, positions2, matches, in1, in2, F, status);
        FILE* o = std::fopen(argv[3], "wb");
        if (!o) return 2;
        const int32_t ni = (int32_t)in1.size();
        std::fwrite(&ni, 4, 1, o);
        std::fwrite(status.data(), 1, status.size(), o);
        std::fwrite(in1.data(), sizeof(Point2d), in1.size(), o);
        std::fwrite(in2.data(), sizeof(Point2d), in2.size(), o);
        std::fwrite(F, sizeof(double), 9, o);
        std::fclose(o);
    } catch (const Error& e) {
        std::fprintf(stderr, "fmat_demo: %s\n", e.what());
        return 1;
    }
    return 0;
}
