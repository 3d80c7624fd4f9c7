// shim_opencv_demo -- compiles include/orbx_shim.hpp in its ORBX_SHIM_USE_OPENCV form, the one a maintainer of the reference
// builds against the real OpenCV headers, here against tests/cpp/stub_opencv (this image has no OpenCV C++
// headers), and drives every free function of the shim through cv:: types: detect / compute, matchFeatures (twice: the
// per-thread matcher is reused), computeFundamentalMatrix with a cv::Mat F, TriangulateMultiplePointsFromTwoView with
// cv::Mat cameras.  With a GPU it checks the calls against each other and prints "ok"; without one it must fail loudly.
// usage: shim_opencv_demo
#define ORBX_SHIM_USE_OPENCV
#include <cmath>
#include <cstdio>
#include <vector>

#include "orbx_shim.hpp"

using namespace orbx_shim;

static unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

int main()
{
    try {
        // a textured frame: blocks of random brightness
        const int w = 320, h = 240;
        cv::Mat img;
        img.create(h, w, CV_8UC1);
        unsigned s = 7;
        std::vector<unsigned char> blocks((size_t)(w / 8) * (h / 8));
        for (size_t i = 0; i < blocks.size(); i++) blocks[i] = (unsigned char)(lcg(s) & 255);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) img.at<unsigned char>(y, x) = blocks[(size_t)(y / 8) * (w / 8) + x / 8];
        OrbFeatureDetector detector(300);
        OrbDescriptorExtractor extractor(300);
        std::vector<cv::KeyPoint> keypoints;
        detector.detect(img, keypoints);
        cv::Mat descriptor;
        extractor.compute(img, keypoints, descriptor);
        if (keypoints.empty() || descriptor.rows != (int)keypoints.size() || descriptor.cols != 32) { std::fprintf(stderr, "bad extraction\n"); return 3; }

        std::vector<cv::DMatch> m1, m2, m3;
        matchFeatures(descriptor, descriptor, m1);
        matchFeatures(descriptor, descriptor, m2, 0.8f);
        threadMatcher().matchRatio(descriptor, descriptor, m3, 0.8f);
        if (m1.size() != m2.size() || m1.size() != m3.size()) { std::fprintf(stderr, "matchFeatures is not repeatable\n"); return 3; }
        for (size_t i = 0; i < m1.size(); i++)
            if (m1[i].queryIdx != m3[i].queryIdx || m1[i].trainIdx != m3[i].trainIdx || m1[i].queryIdx != m1[i].trainIdx) { std::fprintf(stderr, "self-match %zu is wrong\n", i); return 3; }

        // two views of random points: x2 = x1 + parallax(depth)
        const int n = 200;
        std::vector<cv::Point2d> p1((size_t)n), p2((size_t)n);
        std::vector<cv::DMatch> matches((size_t)n);
        std::vector<double> depth((size_t)n);
        for (int i = 0; i < n; i++) {
            const double X = (lcg(s) % 8000) / 1000. - 4., Y = (lcg(s) % 6000) / 1000. - 3., Z = 4. + (lcg(s) % 8000) / 1000.;
            depth[(size_t)i] = Z;
            p1[(size_t)i] = cv::Point2d((float)(500. * X / Z + 160.), (float)(500. * Y / Z + 120.));
            p2[(size_t)i] = cv::Point2d((float)(500. * (X + 0.5) / Z + 160.), (float)(500. * (Y + 0.1) / Z + 120.));
            matches[(size_t)i].queryIdx = i; matches[(size_t)i].trainIdx = i;
        }
        std::vector<cv::Point2d> in1, in2;
        std::vector<unsigned char> status;
        cv::Mat F;
        computeFundamentalMatrix(p1, p2, matches, in1, in2, F, status);
        if (F.rows != 3 || F.cols != 3 || in1.size() < (size_t)(n * 9 / 10)) { std::fprintf(stderr, "fundamental filter: %zu inliers of %d\n", in1.size(), n); return 3; }

        double k[9] = {500, 0, 160, 0, 500, 120, 0, 0, 1}, rt1[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}, rt2[12] = {1, 0, 0, 0.5, 0, 1, 0, 0.1, 0, 0, 1, 0};
        cv::Mat K(3, 3, CV_64F, k), Rt1(3, 4, CV_64F, rt1), Rt2(3, 4, CV_64F, rt2);
        std::vector<cv::Point3d> pts3d;
        const int front = TriangulateMultiplePointsFromTwoView(p1, p2, Rt1, Rt2, K, K, pts3d, true);
        if (front != n || pts3d.size() != (size_t)n) { std::fprintf(stderr, "triangulation: %d of %d in front\n", front, n); return 3; }
        for (int i = 0; i < n; i++)
            if (std::fabs(pts3d[(size_t)i].z - depth[(size_t)i]) > 0.05 * depth[(size_t)i]) { std::fprintf(stderr, "depth %d: %g vs %g\n", i, pts3d[(size_t)i].z, depth[(size_t)i]); return 3; }
        if (TriangulateMultiplePointsFromTwoView(p1, p2, Rt1, Rt2, K, K, pts3d) != 0) { std::fprintf(stderr, "countFront = false must return 0\n"); return 3; }
        std::printf("ok: %zu keypoints, %zu self-matches, %zu/%d inliers, %d points in front\n", keypoints.size(), m1.size(), in1.size(), n, front);
    } catch (const Error& e) {
        std::fprintf(stderr, "shim_opencv_demo: %s\n", e.what());
        return 1;
    }
    return 0;
}
