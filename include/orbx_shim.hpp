// orbx_shim.hpp -- header-only C++ host side over the C ABI of liborbx.so (include/orbx.h).
//
// It re-creates exactly the objects the reference's hot path uses, with the same names, argument meaning and error
// behaviour, so that the reference's two call sites compile unchanged against it:
//
//   src/FeatureExtractor.h:23-24      OrbFeatureDetector detector;  OrbDescriptorExtractor extractor;
//   src/FeatureExtractor.cpp:17,19    detector.detect(frameBuffer, keypoints);  extractor.compute(frameBuffer, keypoints, descriptor);
//   src/CameraPoseEstimator.cpp:200-213   static void matchFeatures(const Mat&, const Mat&, vector<DMatch>&, float ratio = 0.8)
//                                          { BFMatcher matcher(NORM_HAMMING, false); matcher.knnMatch(d1, d2, raw, 2); ... }
//   src/CameraPoseEstimator.cpp:545-586   static void computeFundamentalMatrix(positions1, positions2, matches, ..., F, status)
//   src/CameraPoseEstimator.cpp:134-152   static int TriangulateMultiplePointsFromTwoView(pts1, pts2, Rt1, Rt2, K1, K2, result, countFront)
//
// Without OpenCV headers (this image has none) the POD mirrors below stand in for cv::KeyPoint / cv::DMatch / cv::Mat;
// define ORBX_SHIM_USE_OPENCV before including to bind to the real cv types instead (same memory layouts:
// orbx_keypoint == cv::KeyPoint, orbx_dmatch == cv::DMatch).  Bad input throws (the reference never catches
// cv::Exception either).  No algorithm runs on the CPU here: every call goes to the GPU through liborbx.so.
#pragma once
#include <cstdint>
#include <algorithm>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "orbx.h"

#ifdef ORBX_SHIM_USE_OPENCV
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#endif

namespace orbx_shim {

#ifdef ORBX_SHIM_USE_OPENCV
using cv::DMatch;
using cv::KeyPoint;
using cv::Mat;
using cv::NORM_HAMMING;
using cv::Point2d;
using cv::Point3d;
static inline const uint8_t* mat_ptr(const Mat& m) { return m.data; }
static inline void mat_create_u8(Mat& m, int rows, int cols) { m.create(rows, cols, CV_8UC1); }
static inline size_t mat_step(const Mat& m) { return m.step; }
static inline int mat_channels(const Mat& m) { return m.channels(); }
#else
enum { NORM_HAMMING = 6 };
struct Point2f { float x, y; };
struct Point2d { double x, y; };
struct Point3d { double x, y, z; };
struct KeyPoint {               // field order and size of cv::KeyPoint (28 bytes)
    Point2f pt; float size, angle, response; int octave, class_id;
};
struct DMatch {                 // cv::DMatch (16 bytes)
    int queryIdx, trainIdx, imgIdx; float distance;
};
struct Mat {                    // the part of cv::Mat the hot path touches: an 8-bit matrix, 1 channel (gray, descriptors) or 3 (BGR)
    int rows = 0, cols = 0;
    size_t step = 0;
    uint8_t* data = nullptr;
    int nchannels = 1;
    std::vector<uint8_t> storage;
    Mat() {}
    Mat(int r, int c, uint8_t* borrowed, size_t stp = 0, int ch = 1)
        : rows(r), cols(c), step(stp ? stp : (size_t)c * ch), data(borrowed), nchannels(ch) {}
    int channels() const { return nchannels; }
    void create(int r, int c) { rows = r; cols = c; step = (size_t)c; storage.assign((size_t)r * c, 0); data = storage.data(); }
    bool empty() const { return rows == 0 || cols == 0; }
    uint8_t* ptr(int r) { return data + (size_t)r * step; }
    const uint8_t* ptr(int r) const { return data + (size_t)r * step; }
};
static inline const uint8_t* mat_ptr(const Mat& m) { return m.data; }
static inline void mat_create_u8(Mat& m, int rows, int cols) { m.create(rows, cols); }
static inline size_t mat_step(const Mat& m) { return m.step; }
static inline int mat_channels(const Mat& m) { return m.nchannels; }
#endif

static_assert(sizeof(KeyPoint) == sizeof(orbx_keypoint), "KeyPoint must have cv::KeyPoint's layout");
static_assert(sizeof(DMatch) == sizeof(orbx_dmatch), "DMatch must have cv::DMatch's layout");
static_assert(sizeof(Point2d) == 16 && sizeof(Point3d) == 24, "Point2d / Point3d must be packed doubles");

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& what) : std::runtime_error(what), status(s) {}
};
static inline void check(int status, const char* where)
{
    if (status != ORBX_OK) throw Error(status, std::string(where) + ": " + orbx_last_error());
}

// cv::ORB with the reference's constructor defaults (src/FeatureExtractor.h:23-24 default-constructs it).
class ORB {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };
    explicit ORB(int nfeatures = 500, float scaleFactor = 1.2f, int nlevels = 8, int edgeThreshold = 31, int firstLevel = 0,
                 int WTA_K = 2, int scoreType = HARRIS_SCORE, int patchSize = 31, int device = 0, int maxWidth = 1920, int maxHeight = 1080)
        : h_(nullptr), device_(device), maxw_(maxWidth), maxh_(maxHeight)
    {
        orbx_default_params(&p_);
        p_.nfeatures = nfeatures; p_.scale_factor = scaleFactor; p_.nlevels = nlevels; p_.edge_threshold = edgeThreshold;
        p_.first_level = firstLevel; p_.wta_k = WTA_K; p_.score_type = scoreType; p_.patch_size = patchSize;
    }
    ~ORB() { if (h_) orbx_destroy(h_); }
    ORB(const ORB&) = delete;
    ORB& operator=(const ORB&) = delete;

    // FeatureDetector::detect(image, keypoints): keypoints is cleared and filled (canonical order: octave, y, x)
    void detect(const Mat& image, std::vector<KeyPoint>& keypoints)
    {
        ensure(image);
        int cap = default_cap(), n = 0;
        keypoints.resize((size_t)cap);
        int rc = orbx_detect(h_, mat_ptr(image), image.cols, image.rows, mat_step(image), reinterpret_cast<orbx_keypoint*>(keypoints.data()), cap, &n);
        if (rc == ORBX_E_CAPACITY) {   // ties at a retention cut can return more than nfeatures (OpenCV does the same)
            cap = orbx_max_keypoints(h_);
            keypoints.resize((size_t)cap);
            rc = orbx_detect(h_, mat_ptr(image), image.cols, image.rows, mat_step(image), reinterpret_cast<orbx_keypoint*>(keypoints.data()), cap, &n);
        }
        check(rc, "ORB::detect");
        keypoints.resize((size_t)n);
    }

    // DescriptorExtractor::compute(image, keypoints, descriptors): keypoints may be filtered / regrouped by octave;
    // descriptors becomes N x 32 CV_8U with row i describing keypoints[i]
    void compute(const Mat& image, std::vector<KeyPoint>& keypoints, Mat& descriptors)
    {
        ensure(image);
        int n = (int)keypoints.size();
        mat_create_u8(descriptors, n > 0 ? n : 1, 32);
        check(orbx_compute(h_, mat_ptr(image), image.cols, image.rows, mat_step(image), reinterpret_cast<orbx_keypoint*>(keypoints.data()), &n,
                           descriptors.data),
              "ORB::compute");
        keypoints.resize((size_t)n);
        descriptors.rows = n;
    }

    // cv::ORB::operator()(image, mask, keypoints, descriptors): detect + compute with one pyramid
    void operator()(const Mat& image, std::vector<KeyPoint>& keypoints, Mat& descriptors)
    {
        ensure(image);
        int cap = default_cap(), n = 0;
        for (int attempt = 0; attempt < 2; attempt++) {
            keypoints.resize((size_t)cap);
            mat_create_u8(descriptors, cap, 32);
            int rc = orbx_detect_and_compute(h_, mat_ptr(image), image.cols, image.rows, mat_step(image),
                                             reinterpret_cast<orbx_keypoint*>(keypoints.data()), descriptors.data, cap, &n);
            if (rc == ORBX_E_CAPACITY && attempt == 0) { cap = orbx_max_keypoints(h_); continue; }
            check(rc, "ORB::operator()");
            break;
        }
        keypoints.resize((size_t)n);
        descriptors.rows = n;
    }

    orbx_handle handle() { return h_; }

private:
    int default_cap() const { return p_.nfeatures + (p_.nfeatures / 4 > 512 ? p_.nfeatures / 4 : 512); }
    void ensure(const Mat& image)
    {
        if (image.empty() || !mat_ptr(image)) throw Error(ORBX_E_INVALID, "ORB: empty image");
        if (h_ && (image.cols > maxw_ || image.rows > maxh_)) { orbx_destroy(h_); h_ = nullptr; }
        if (!h_) {
            if (image.cols > maxw_) maxw_ = image.cols;
            if (image.rows > maxh_) maxh_ = image.rows;
            check(orbx_create(&h_, &p_, device_, maxw_, maxh_, 1), "ORB: orbx_create");   // lazily, like ProcessingNode::init()
            channels_ = 1;
        }
        // cv::ORB converts non-gray frames itself (the reference loads them CV_LOAD_IMAGE_UNCHANGED, src/FrameLoader.cpp:62)
        const int ch = mat_channels(image);
        if (ch != channels_) { check(orbx_set_input_channels(h_, ch), "ORB: orbx_set_input_channels"); channels_ = ch; }
    }
    int channels_ = 1;
    orbx_handle h_;
    orbx_params p_;
    int device_, maxw_, maxh_;
};

typedef ORB OrbFeatureDetector;        // OpenCV 2.4 typedefs, the names used at src/FeatureExtractor.h:23-24
typedef ORB OrbDescriptorExtractor;

// cv::BFMatcher(NORM_HAMMING, crossCheck = false)
class BFMatcher {
public:
    explicit BFMatcher(int normType = NORM_HAMMING, bool crossCheck = false, int device = 0) : h_(nullptr)
    {
        if (normType != NORM_HAMMING || crossCheck) throw Error(ORBX_E_INVALID, "BFMatcher: only (NORM_HAMMING, false) is provided");
        check(hamx_create(&h_, device), "BFMatcher");
    }
    ~BFMatcher() { if (h_) hamx_destroy(h_); }
    BFMatcher(const BFMatcher&) = delete;
    BFMatcher& operator=(const BFMatcher&) = delete;

    // knnMatch(queryDescriptors, trainDescriptors, matches, k = 2): matches.size() == query rows; each row sorted by
    // (distance, trainIdx) and shorter than 2 only if the train set has fewer than 2 rows
    void knnMatch(const Mat& query, const Mat& train, std::vector<std::vector<DMatch> >& matches, int k)
    {
        if (k != 2) throw Error(ORBX_E_INVALID, "BFMatcher::knnMatch: the reference calls it with k = 2 only");
        if ((query.rows && query.cols != 32) || (train.rows && train.cols != 32) || (query.rows && mat_step(query) != 32) ||
            (train.rows && mat_step(train) != 32))
            throw Error(ORBX_E_INVALID, "BFMatcher::knnMatch: descriptors must be continuous N x 32 CV_8U");
        std::vector<orbx_dmatch> flat((size_t)query.rows * 2);
        std::vector<int32_t> counts((size_t)query.rows);
        check(hamx_knn2(h_, mat_ptr(query), query.rows, mat_ptr(train), train.rows, flat.data(), counts.data()), "BFMatcher::knnMatch");
        matches.assign((size_t)query.rows, std::vector<DMatch>());
        for (int i = 0; i < query.rows; i++)
            for (int j = 0; j < counts[(size_t)i]; j++) {
                DMatch m;
                std::memcpy(static_cast<void*>(&m), &flat[(size_t)2 * i + j], sizeof(m));   // same layout (static_assert above)
                matches[(size_t)i].push_back(m);
            }
    }

    // the whole of matchFeatures() in one submission (knnMatch k=2 + ratio test on the device)
    void matchRatio(const Mat& d1, const Mat& d2, std::vector<DMatch>& matches, float ratio)
    {
        std::vector<orbx_dmatch> good((size_t)(d1.rows > 0 ? d1.rows : 1));
        int64_t n = 0;
        check(hamx_match_ratio(h_, mat_ptr(d1), d1.rows, mat_ptr(d2), d2.rows, ratio, good.data(), &n), "BFMatcher::matchRatio");
        matches.resize((size_t)n);
        if (n) std::memcpy(static_cast<void*>(matches.data()), good.data(), (size_t)n * sizeof(DMatch));
    }

    hamx_handle handle() { return h_; }

private:
    hamx_handle h_;
};

// The reference's routine, line for line in behaviour (src/CameraPoseEstimator.cpp:200-213).  The reference constructs a
// BFMatcher per call, which costs nothing on the CPU; here a matcher owns a CUDA stream and device workspaces, and the
// reference calls matchFeatures 5 times per frame (:405-409), so one matcher per host thread is created on first use and
// kept (the reference's pipeline is single-threaded, src/main.cpp:49-51).
static inline BFMatcher& threadMatcher()
{
    static thread_local BFMatcher matcher(NORM_HAMMING, false);
    return matcher;
}
static inline void matchFeatures(const Mat& descriptors1, const Mat& descriptors2, std::vector<DMatch>& matches, float ratio = 0.8f)
{
    BFMatcher& matcher = threadMatcher();
    std::vector<std::vector<DMatch> > raw_matches;
    matcher.knnMatch(descriptors1, descriptors2, raw_matches, 2);
    matches.clear();
    for (size_t i = 0; i < raw_matches.size(); i++) {
        if (raw_matches[i].size() < 2) continue;   // the reference indexes [1] unguarded; such rows cannot pass a ratio test
        if (raw_matches[i][0].distance < raw_matches[i][1].distance * ratio) matches.push_back(raw_matches[i][0]);
    }
}

// computeFundamentalMatrix (src/CameraPoseEstimator.cpp:545-586): RANSAC status of the matches + the 8-point matrix of the
// inliers, on the GPU (fmx_fundamental_batch).  F is row-major 3x3 (all zeros when there is no model).  MAX_DISTANCE and
// CONFIDENCE are the reference's constants (src/ParamConfig.h:24-25).
class FundamentalFilter {
public:
    explicit FundamentalFilter(int device = 0) : h_(nullptr) { check(fmx_create(&h_, device), "FundamentalFilter"); }
    ~FundamentalFilter() { if (h_) fmx_destroy(h_); }
    FundamentalFilter(const FundamentalFilter&) = delete;
    FundamentalFilter& operator=(const FundamentalFilter&) = delete;

    void compute(const std::vector<Point2d>& positions1, const std::vector<Point2d>& positions2, const std::vector<DMatch>& matches,
                 std::vector<Point2d>& inlierPositions1, std::vector<Point2d>& inlierPositions2, double F[9],
                 std::vector<unsigned char>& status, double maxDistance = 3., double confidence = 0.85)
    {
        const size_t n = matches.size();
        std::vector<float> a(2 * n + 2), b(2 * n + 2);
        for (size_t i = 0; i < n; i++) {
            const Point2d& p = positions1.at((size_t)matches[i].queryIdx);
            const Point2d& q = positions2.at((size_t)matches[i].trainIdx);
            a[2 * i] = (float)p.x; a[2 * i + 1] = (float)p.y;          // findFundamentalMat converts its inputs to CV_32F
            b[2 * i] = (float)q.x; b[2 * i + 1] = (float)q.y;
        }
        status.assign(n ? n : 1, 0);
        const int32_t count = (int32_t)n;
        int32_t ninl = 0;
        check(fmx_fundamental_batch(h_, a.data(), b.data(), &count, 1, n ? (int)n : 1, maxDistance, confidence, status.data(), F, &ninl),
              "computeFundamentalMatrix");
        status.resize(n);
        inlierPositions1.clear();
        inlierPositions2.clear();
        for (size_t i = 0; i < n; i++)
            if (status[i]) {
                inlierPositions1.push_back(positions1[(size_t)matches[i].queryIdx]);
                inlierPositions2.push_back(positions2[(size_t)matches[i].trainIdx]);
            }
    }
    fmx_handle handle() { return h_; }

private:
    fmx_handle h_;
};

static inline void computeFundamentalMatrix(const std::vector<Point2d>& positions1, const std::vector<Point2d>& positions2,
                                            const std::vector<DMatch>& matches, std::vector<Point2d>& inlierPositions1,
                                            std::vector<Point2d>& inlierPositions2, double F[9], std::vector<unsigned char>& status)
{
    static thread_local FundamentalFilter f;     // one per host thread, kept: see threadMatcher()
    f.compute(positions1, positions2, matches, inlierPositions1, inlierPositions2, F, status);
}
#ifdef ORBX_SHIM_USE_OPENCV
static inline void computeFundamentalMatrix(const std::vector<Point2d>& positions1, const std::vector<Point2d>& positions2,
                                            const std::vector<DMatch>& matches, std::vector<Point2d>& inlierPositions1,
                                            std::vector<Point2d>& inlierPositions2, Mat& F, std::vector<unsigned char>& status)
{
    double f[9];
    computeFundamentalMatrix(positions1, positions2, matches, inlierPositions1, inlierPositions2, f, status);
    F.create(3, 3, CV_64F);
    std::memcpy(F.data, f, sizeof(f));
}
#endif

// TriangulateMultiplePointsFromTwoView (src/CameraPoseEstimator.cpp:134-152) and the four-candidate test built on it
// (:334-349), on the GPU (trx_triangulate / trx_triangulate_hypotheses).  Rt are 3x4, K 3x3, row-major doubles.
class Triangulator {
public:
    explicit Triangulator(int device = 0) : h_(nullptr) { check(trx_create(&h_, device), "Triangulator"); }
    ~Triangulator() { if (h_) trx_destroy(h_); }
    Triangulator(const Triangulator&) = delete;
    Triangulator& operator=(const Triangulator&) = delete;

    int triangulate(const std::vector<Point2d>& pts1, const std::vector<Point2d>& pts2, const double* Rt1, const double* Rt2,
                    const double* K1, const double* K2, std::vector<Point3d>& result, bool countFront)
    {
        if (pts1.size() != pts2.size()) throw Error(ORBX_E_INVALID, "TriangulateMultiplePointsFromTwoView: pts1 and pts2 differ in length");
        result.assign(pts1.size(), Point3d());
        int32_t nfront = 0;
        check(trx_triangulate(h_, reinterpret_cast<const double*>(pts1.data()), reinterpret_cast<const double*>(pts2.data()), (int)pts1.size(),
                              Rt1, Rt2, K1, K2, reinterpret_cast<double*>(result.data()), nullptr, &nfront),
              "TriangulateMultiplePointsFromTwoView");
        return countFront ? (int)nfront : 0;       // the reference returns 0 unless asked to count (:150-151)
    }
    // index of the candidate with the most points in front (first maximum), its points in `result`, all counts in `counts`
    int bestHypothesis(const std::vector<Point2d>& pts1, const std::vector<Point2d>& pts2, const double* Rt1, const double* Rts, int nhyp,
                       const double* K1, const double* K2, std::vector<Point3d>& result, std::vector<int>& counts)
    {
        if (pts1.size() != pts2.size()) throw Error(ORBX_E_INVALID, "bestHypothesis: pts1 and pts2 differ in length");
        result.assign(pts1.size(), Point3d());
        std::vector<int32_t> c((size_t)nhyp);
        int32_t best = -1;
        check(trx_triangulate_hypotheses(h_, reinterpret_cast<const double*>(pts1.data()), reinterpret_cast<const double*>(pts2.data()),
                                         (int)pts1.size(), Rt1, Rts, nhyp, K1, K2, reinterpret_cast<double*>(result.data()), c.data(), &best),
              "bestHypothesis");
        counts.assign(c.begin(), c.end());
        return (int)best;
    }
    trx_handle handle() { return h_; }

private:
    trx_handle h_;
};

static inline int TriangulateMultiplePointsFromTwoView(const std::vector<Point2d>& pts1, const std::vector<Point2d>& pts2, const double* Rt1,
                                                       const double* Rt2, const double* K1, const double* K2, std::vector<Point3d>& result,
                                                       bool countFront = false)
{
    static thread_local Triangulator t;          // one per host thread, kept: see threadMatcher()
    return t.triangulate(pts1, pts2, Rt1, Rt2, K1, K2, result, countFront);
}
#ifdef ORBX_SHIM_USE_OPENCV
// the reference's own argument list (const Mat& Rt1, Rt2, K1, K2: CV_64F, 3x4 and 3x3)
static inline const double* mat_f64(const Mat& m, int rows, int cols, const char* what)
{
    if (m.rows != rows || m.cols != cols || m.channels() != 1 || m.elemSize() != 8 || (size_t)m.step != (size_t)cols * 8)
        throw Error(ORBX_E_INVALID, std::string(what) + ": expected a continuous CV_64F matrix");
    return reinterpret_cast<const double*>(m.data);
}
static inline int TriangulateMultiplePointsFromTwoView(const std::vector<Point2d>& pts1, const std::vector<Point2d>& pts2, const Mat& Rt1,
                                                       const Mat& Rt2, const Mat& K1, const Mat& K2, std::vector<Point3d>& result,
                                                       bool countFront = false)
{
    return TriangulateMultiplePointsFromTwoView(pts1, pts2, mat_f64(Rt1, 3, 4, "Rt1"), mat_f64(Rt2, 3, 4, "Rt2"), mat_f64(K1, 3, 3, "K1"),
                                                mat_f64(K2, 3, 3, "K2"), result, countFront);
}
#endif

// ---- frame ingest: what FrameLoader's imread(path, CV_LOAD_IMAGE_UNCHANGED) (src/FrameLoader.cpp:62) returns for baseline JPEG
// files, decoded on the GPU a batch at a time (jpgx_*): grey files as 1-channel matrices, YCbCr files as 3-channel BGR.
class JpegDecoder {
public:
    explicit JpegDecoder(int device = 0) : h_(nullptr) { check(jpgx_create(&h_, device), "JpegDecoder"); }
    ~JpegDecoder() { if (h_) jpgx_destroy(h_); }
    JpegDecoder(const JpegDecoder&) = delete;
    JpegDecoder& operator=(const JpegDecoder&) = delete;

    // false for files the GPU path does not take (progressive, 12-bit, CMYK ...): keep imread for those
    static bool supported(const std::vector<uint8_t>& file, int* width = nullptr, int* height = nullptr, int* channels = nullptr)
    {
        int32_t info[6];
        if (jpgx_probe(file.data(), file.size(), info) != ORBX_OK) return false;
        if (width) *width = info[0];
        if (height) *height = info[1];
        if (channels) *channels = info[4];
        return true;
    }
    // all files of one size and kind; frames[i] is rows x cols x channels, 8 bits, as imread would have returned it
    void decode(const std::vector<std::vector<uint8_t> >& files, std::vector<Mat>& frames)
    {
        frames.clear();
        if (files.empty()) return;
        int32_t info[6];
        check(jpgx_probe(files[0].data(), files[0].size(), info), "JpegDecoder::decode");
        const int w = info[0], h = info[1], ch = info[4];
        std::vector<const uint8_t*> ptrs(files.size());
        std::vector<size_t> sizes(files.size());
        for (size_t i = 0; i < files.size(); i++) { ptrs[i] = files[i].data(); sizes[i] = files[i].size(); }
        std::vector<uint8_t> raw(files.size() * (size_t)w * h * ch);
        check((ch == 1 ? jpgx_decode_gray_batch : jpgx_decode_bgr_batch)(h_, ptrs.data(), sizes.data(), (int)files.size(), w, h, raw.data(),
                                                                          (size_t)w * h * ch, (size_t)w * ch), "JpegDecoder::decode");
        frames.resize(files.size());
        for (size_t i = 0; i < files.size(); i++) {
#ifdef ORBX_SHIM_USE_OPENCV
            frames[i].create(h, w, ch == 1 ? CV_8UC1 : CV_8UC3);
            for (int y = 0; y < h; y++) std::memcpy(frames[i].ptr(y), raw.data() + (i * h + y) * (size_t)w * ch, (size_t)w * ch);
#else
            frames[i].create(h, w * ch);
            frames[i].cols = w; frames[i].nchannels = ch; frames[i].step = (size_t)w * ch;
            std::memcpy(frames[i].data, raw.data() + i * (size_t)w * h * ch, (size_t)w * h * ch);
#endif
        }
    }
    jpgx_handle handle() { return h_; }

private:
    jpgx_handle h_;
};

// ---- the vendored DBoW2 (ThirdParty/DBoW2/DBoW2), the part a loop closer calls: BowVector / FeatureVector keep the
// reference's types (std::map), OrbVocabulary stands where TemplatedVocabulary<FORB::TDescriptor, FORB> stood; transform()
// and score() run on the GPU (bowx_*).
namespace DBoW2 {
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
enum WeightingType { TF_IDF, TF, IDF, BINARY };                                               // BowVector.h:36-42
enum ScoringType { L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT };            // BowVector.h:45-53
typedef std::map<WordId, WordValue> BowVector;                                                // BowVector.h:56-57
typedef std::map<NodeId, std::vector<unsigned int> > FeatureVector;                           // FeatureVector.h:21-22

class OrbVocabulary {
public:
    explicit OrbVocabulary(int device = 0) : h_(nullptr) { check(bowx_create(&h_, device), "OrbVocabulary"); }
    ~OrbVocabulary() { if (h_) bowx_destroy(h_); }
    OrbVocabulary(const OrbVocabulary&) = delete;
    OrbVocabulary& operator=(const OrbVocabulary&) = delete;

    // TemplatedVocabulary.h:1333-1416.  Empty lines are skipped: the reference's own loader turns the newline that ends a
    // file written by saveToTextFile into one more child of the root with an uninitialised descriptor.
    bool loadFromTextFile(const std::string& filename)
    {
        std::ifstream f(filename.c_str());
        if (!f) return false;
        std::string line;
        if (!std::getline(f, line)) return false;
        int k = 0, L = 0, n1 = 0, n2 = 0;
        { std::stringstream ss(line); ss >> k >> L >> n1 >> n2; }
        if (k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) return false;
        std::vector<int32_t> parent(1, 0);
        std::vector<uint8_t> leaf(1, 0), desc(32, 0);
        std::vector<double> weight(1, 0.);
        while (std::getline(f, line)) {
            std::stringstream ss(line);
            int pid, is_leaf;
            if (!(ss >> pid >> is_leaf)) continue;
            parent.push_back(pid);
            leaf.push_back(is_leaf > 0);
            for (int i = 0; i < 32; i++) { int b = 0; ss >> b; desc.push_back((uint8_t)b); }
            double w = 0;
            ss >> w;
            weight.push_back(w);
        }
        check(bowx_set_vocabulary(h_, k, L, n1, n2, (int)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data()), "loadFromTextFile");
        return true;
    }
    unsigned int size() const { return (unsigned)info(5); }
    bool empty() const { return size() == 0; }
    int getBranchingFactor() const { return info(0); }
    int getDepthLevels() const { return info(1); }
    ScoringType getScoringType() const { return (ScoringType)info(2); }
    WeightingType getWeightingType() const { return (WeightingType)info(3); }

    // transform(features, v, fv, levelsup) with the features as the rows of one n x 32 descriptor matrix (what
    // OrbDescriptorExtractor::compute returns) ...
    void transform(const Mat& descriptors, BowVector& v, FeatureVector& fv, int levelsup) const { run(descriptors, v, &fv, levelsup); }
    void transform(const Mat& descriptors, BowVector& v) const { run(descriptors, v, nullptr, 0); }
    // ... or, as in the reference, as a vector of 1 x 32 rows
    void transform(const std::vector<Mat>& features, BowVector& v, FeatureVector& fv, int levelsup) const { run(gather(features), v, &fv, levelsup); }
    void transform(const std::vector<Mat>& features, BowVector& v) const { run(gather(features), v, nullptr, 0); }
    WordId transform(const Mat& feature) const
    {
        if (empty()) return 0;
        uint32_t word = 0, node = 0;
        double weight = 0;
        check(bowx_transform_features(h_, mat_ptr(feature), 1, 0, &word, &weight, &node), "transform");
        return word;
    }
    double score(const BowVector& a, const BowVector& b) const
    {
        std::vector<uint32_t> w1, w2;
        std::vector<double> v1, v2;
        flatten(a, w1, v1); flatten(b, w2, v2);
        double s = 0;
        check(bowx_score(h_, w1.data(), v1.data(), (int)w1.size(), w2.data(), v2.data(), (int)w2.size(), &s), "score");
        return s;
    }
    // score(query, entry) for every entry of a database of vectors, one launch
    std::vector<double> score(const BowVector& query, const std::vector<BowVector>& database) const
    {
        std::vector<uint32_t> qw, words;
        std::vector<double> qv, vals;
        flatten(query, qw, qv);
        std::vector<int64_t> start(database.size());
        std::vector<int32_t> count(database.size());
        for (size_t e = 0; e < database.size(); e++) {
            start[e] = (int64_t)words.size();
            count[e] = (int32_t)database[e].size();
            for (BowVector::const_iterator it = database[e].begin(); it != database[e].end(); ++it) { words.push_back(it->first); vals.push_back(it->second); }
        }
        std::vector<double> scores(database.size());
        check(bowx_score_batch(h_, qw.data(), qv.data(), (int)qw.size(), start.data(), count.data(), words.data(), vals.data(), (int64_t)words.size(),
                               (int)database.size(), scores.data()), "score");
        return scores;
    }
    NodeId getParentNode(WordId wid, int levelsup) const { uint32_t n = 0; check(bowx_parent_node(h_, wid, levelsup, &n), "getParentNode"); return n; }
    WordValue getWordWeight(WordId wid) const { double w = 0; check(bowx_word_weight(h_, wid, &w), "getWordWeight"); return w; }
    int stopWords(double minWeight) { int32_t c = 0; check(bowx_stop_words(h_, minWeight, &c), "stopWords"); return c; }
    bowx_handle handle() { return h_; }

private:
    int info(int i) const { int32_t v[6]; check(bowx_vocabulary_info(h_, v), "OrbVocabulary"); return v[i]; }
    static void flatten(const BowVector& m, std::vector<uint32_t>& w, std::vector<double>& v)
    {
        w.reserve(m.size()); v.reserve(m.size());
        for (BowVector::const_iterator it = m.begin(); it != m.end(); ++it) { w.push_back(it->first); v.push_back(it->second); }
    }
    static Mat gather(const std::vector<Mat>& features)
    {
        Mat all;
        mat_create_u8(all, (int)features.size(), 32);
        for (size_t i = 0; i < features.size(); i++) {
            if (features[i].rows != 1 || features[i].cols != 32) throw Error(ORBX_E_INVALID, "transform: a feature is not a 1 x 32 row");
            std::memcpy(all.ptr((int)i), mat_ptr(features[i]), 32);
        }
        return all;
    }
    void run(const Mat& d, BowVector& v, FeatureVector* fv, int levelsup) const
    {
        v.clear();
        if (fv) fv->clear();
        const int n = d.rows;
        if (n == 0 || empty()) return;
        if (d.cols != 32 || mat_channels(d) != 1 || mat_step(d) != 32) throw Error(ORBX_E_INVALID, "transform: descriptors must be a continuous n x 32 8-bit matrix");
        const int32_t count = n;
        std::vector<uint32_t> words((size_t)n), nodes, feats;
        std::vector<double> vals((size_t)n);
        std::vector<int32_t> offs;
        int32_t nbow = 0, nfv = 0;
        if (fv) { nodes.resize((size_t)n); feats.resize((size_t)n); offs.resize((size_t)n + 1); }
        check(bowx_transform_batch(h_, mat_ptr(d), &count, 1, n, levelsup, words.data(), vals.data(), &nbow, fv ? nodes.data() : nullptr,
                                   fv ? offs.data() : nullptr, fv ? feats.data() : nullptr, fv ? &nfv : nullptr), "transform");
        for (int i = 0; i < nbow; i++) v.insert(v.end(), BowVector::value_type(words[(size_t)i], vals[(size_t)i]));
        for (int g = 0; g < nfv; g++)
            fv->insert(fv->end(), FeatureVector::value_type(nodes[(size_t)g], std::vector<unsigned int>(feats.begin() + offs[(size_t)g], feats.begin() + offs[(size_t)g + 1])));
    }
    bowx_handle h_;
};
}  // namespace DBoW2

#ifndef ORBX_SHIM_USE_OPENCV
// ---- the slice of the reference's data model and node API that the front-end touches
struct Features {                           // src/Frame.h:22-34
    std::vector<Point2d> positions;
    Mat descriptors;
    std::vector<int> mapPointsIndices;
    std::vector<double> scales;
};
struct Frame { Mat frameBuffer; Features features; };          // src/Frame.h:36-72 (front-end fields)
struct DataManager { std::vector<Frame> frames; };             // src/DataManager.h:23-36

class ProcessingNode {                      // src/ProcessingNode.h:16-32
public:
    explicit ProcessingNode(const std::string& n) : name(n) {}
    virtual void init() {}
    virtual void finish() {}
    virtual void destroy() {}
    virtual ~ProcessingNode() {}
    virtual void process(DataManager&, int) {}
    virtual bool validationCheck(DataManager&, int) { return true; }
    std::string name;
};

class FeatureExtractor : public ProcessingNode {   // src/FeatureExtractor.{h,cpp}
public:
    FeatureExtractor() : ProcessingNode("FeatureExtractor") {}
    explicit FeatureExtractor(int nfeatures) : ProcessingNode("FeatureExtractor"), detector(nfeatures), extractor(nfeatures) {}
    virtual void process(DataManager& data, int frameIdx)
    {
        Frame& frame = data.frames[(size_t)frameIdx];
        std::vector<KeyPoint> keypoints;
        detector.detect(frame.frameBuffer, keypoints);
        Mat descriptor;
        extractor.compute(frame.frameBuffer, keypoints, descriptor);
        frame.features.descriptors = descriptor;
        if (frame.features.descriptors.storage.size()) frame.features.descriptors.data = frame.features.descriptors.storage.data();
        for (size_t i = 0; i < keypoints.size(); i++) {
            Point2d p = { (double)keypoints[i].pt.x, (double)keypoints[i].pt.y };
            frame.features.positions.push_back(p);
            frame.features.scales.push_back((double)keypoints[i].size);
        }
        frame.features.mapPointsIndices.resize(keypoints.size(), -1);
    }
    virtual bool validationCheck(DataManager&, int) { return true; }

private:
    OrbFeatureDetector detector;
    OrbDescriptorExtractor extractor;
};

// The fast path for a loaded sequence (the reference loads every frame before the loop, src/main.cpp:35-37): batches of
// frames go through orbx_submit_batch / orbx_wait_batch, so uploads, kernels and downloads of consecutive batches overlap,
// and every frame is matched against its predecessor on the device (matchFeatures(cur, prev), CameraPoseEstimator.cpp:409).
// Results land where FeatureExtractor::process and matchFeatures would have put them.
class SequenceFrontEnd {
public:
    SequenceFrontEnd(int nfeatures, float ratio, int batch, int width, int height, int device = 0)
        : orb_(nullptr), bf_(nullptr), fm_(nullptr), ratio_(ratio), batch_(batch), w_(width), h_(height), cap_(0), depth_(0),
          device_(device), max_distance_(3.), confidence_(0.85), status_out_(nullptr), F_out_(nullptr)
    {
        orbx_params p;
        orbx_default_params(&p);
        p.nfeatures = nfeatures;
        check(orbx_create(&orb_, &p, device, width, height, batch), "SequenceFrontEnd: orbx_create");
        check(hamx_create(&bf_, device), "SequenceFrontEnd: hamx_create");
        cap_ = std::min(orbx_max_keypoints(orb_), nfeatures + std::max(nfeatures / 4, 512));
        depth_ = orbx_pipeline_depth(orb_);
        slots_.resize((size_t)depth_);
        for (size_t i = 0; i < slots_.size(); i++) {
            Slot& s = slots_[i];
            check(orbx_host_alloc((size_t)batch * width * height, (void**)&s.frames), "SequenceFrontEnd: pinned frames");
            check(orbx_host_alloc((size_t)batch * cap_ * sizeof(orbx_keypoint), (void**)&s.kps), "SequenceFrontEnd: pinned keypoints");
            check(orbx_host_alloc((size_t)batch * cap_ * 32, (void**)&s.desc), "SequenceFrontEnd: pinned descriptors");
            check(orbx_host_alloc((size_t)batch * cap_ * sizeof(orbx_dmatch), (void**)&s.good), "SequenceFrontEnd: pinned matches");
            s.counts.resize((size_t)batch);
            s.ngood.resize((size_t)batch);
        }
    }
    ~SequenceFrontEnd()
    {
        while (orb_ && orbx_batches_in_flight(orb_) > 0) orbx_wait_batch(orb_);
        for (size_t i = 0; i < slots_.size(); i++) {
            orbx_host_free(slots_[i].frames); orbx_host_free(slots_[i].kps); orbx_host_free(slots_[i].desc); orbx_host_free(slots_[i].good);
            if (slots_[i].status) orbx_host_free(slots_[i].status);
            if (slots_[i].F) orbx_host_free(slots_[i].F);
        }
        if (fm_) fmx_destroy(fm_);
        if (bf_) hamx_destroy(bf_);
        if (orb_) orbx_destroy(orb_);
    }
    SequenceFrontEnd(const SequenceFrontEnd&) = delete;
    SequenceFrontEnd& operator=(const SequenceFrontEnd&) = delete;

    // Append computeFundamentalMatrix (src/CameraPoseEstimator.cpp:545-586) of every (frame, predecessor) pair to each batch:
    // process() then also fills status[i - first] (one byte per match, OpenCV's RANSAC mask) and F[i - first] (9 doubles,
    // the 8-point matrix of the inliers; all zeros when there is no model).
    void enableFilter(double maxDistance = 3., double confidence = 0.85)
    {
        if (!fm_) check(fmx_create(&fm_, device_), "SequenceFrontEnd: fmx_create");
        max_distance_ = maxDistance;
        confidence_ = confidence;
        for (size_t i = 0; i < slots_.size(); i++) {
            Slot& s = slots_[i];
            if (!s.status) check(orbx_host_alloc((size_t)batch_ * cap_, (void**)&s.status), "SequenceFrontEnd: pinned status");
            if (!s.F) check(orbx_host_alloc((size_t)batch_ * 9 * sizeof(double), (void**)&s.F), "SequenceFrontEnd: pinned F");
            s.ninl.resize((size_t)batch_);
        }
    }

    // frames [first, first + count) of dm (all width x height, CV_8UC1): fills frames[i].features and matches[i - first]
    // (frame i against frame i-1; empty for the first frame of a sequence -- call reset() to start a new one)
    void process(DataManager& dm, int first, int count, std::vector<std::vector<DMatch> >& matches)
    {
        std::vector<std::vector<unsigned char> > status;
        std::vector<std::vector<double> > F;
        process(dm, first, count, matches, status, F);
    }
    void process(DataManager& dm, int first, int count, std::vector<std::vector<DMatch> >& matches,
                 std::vector<std::vector<unsigned char> >& status, std::vector<std::vector<double> >& F)
    {
        status_out_ = fm_ ? &status : nullptr;
        F_out_ = fm_ ? &F : nullptr;
        status.assign(fm_ ? (size_t)count : 0, std::vector<unsigned char>());
        F.assign(fm_ ? (size_t)count : 0, std::vector<double>(9, 0.));
        matches.assign((size_t)count, std::vector<DMatch>());
        std::vector<std::pair<int, int> > pending;     // (first frame, count) of the batches in flight, oldest first
        size_t next_slot = 0, oldest_slot = 0;
        for (int b = first; b < first + count; b += batch_) {
            const int n = std::min(batch_, first + count - b);
            if ((int)pending.size() == depth_) { collect(dm, pending.front(), slots_[oldest_slot], matches, first); pending.erase(pending.begin()); oldest_slot = (oldest_slot + 1) % slots_.size(); }
            Slot& s = slots_[next_slot];
            std::vector<const uint8_t*> ptrs((size_t)n);
            for (int i = 0; i < n; i++) {              // gather into pinned memory (frames of a cv::Mat sequence are pageable)
                const Mat& fb = dm.frames[(size_t)(b + i)].frameBuffer;
                if (fb.rows != h_ || fb.cols != w_ || mat_channels(fb) != 1) throw Error(ORBX_E_INVALID, "SequenceFrontEnd: frame size / type mismatch");
                uint8_t* dst = s.frames + (size_t)i * w_ * h_;
                for (int y = 0; y < h_; y++) std::memcpy(dst + (size_t)y * w_, fb.ptr(y), (size_t)w_);
                ptrs[(size_t)i] = dst;
            }
            check(orbx_submit_batch_filtered(orb_, bf_, fm_, ptrs.data(), n, w_, h_, (size_t)w_, ratio_, s.kps, s.desc, cap_, s.counts.data(),
                                             s.good, s.ngood.data(), max_distance_, confidence_, s.status, s.F,
                                             fm_ ? s.ninl.data() : nullptr), "SequenceFrontEnd: orbx_submit_batch");
            pending.push_back(std::make_pair(b, n));
            next_slot = (next_slot + 1) % slots_.size();
        }
        while (!pending.empty()) { collect(dm, pending.front(), slots_[oldest_slot], matches, first); pending.erase(pending.begin()); oldest_slot = (oldest_slot + 1) % slots_.size(); }
    }
    void reset() { check(orbx_reset_sequence(orb_), "SequenceFrontEnd: reset"); }

private:
    struct Slot {
        uint8_t* frames; orbx_keypoint* kps; uint8_t* desc; orbx_dmatch* good;
        std::vector<int32_t> counts; std::vector<int64_t> ngood;
        uint8_t* status; double* F; std::vector<int32_t> ninl;
        Slot() : frames(nullptr), kps(nullptr), desc(nullptr), good(nullptr), status(nullptr), F(nullptr) {}
    };
    void collect(DataManager& dm, std::pair<int, int> batch, Slot& s, std::vector<std::vector<DMatch> >& matches, int first)
    {
        check(orbx_wait_batch(orb_), "SequenceFrontEnd: orbx_wait_batch");
        for (int i = 0; i < batch.second; i++) {
            Features& ft = dm.frames[(size_t)(batch.first + i)].features;
            const int n = s.counts[(size_t)i];
            const orbx_keypoint* k = s.kps + (size_t)i * cap_;
            ft.positions.resize((size_t)n); ft.scales.resize((size_t)n); ft.mapPointsIndices.assign((size_t)n, -1);
            for (int j = 0; j < n; j++) {               // FeatureExtractor.cpp:20-25
                ft.positions[(size_t)j].x = (double)k[j].x; ft.positions[(size_t)j].y = (double)k[j].y;
                ft.scales[(size_t)j] = (double)k[j].size;
            }
            mat_create_u8(ft.descriptors, n, 32);
            if (n) std::memcpy(ft.descriptors.data, s.desc + (size_t)i * cap_ * 32, (size_t)n * 32);
            const DMatch* g = reinterpret_cast<const DMatch*>(s.good + (size_t)i * cap_);
            matches[(size_t)(batch.first + i - first)].assign(g, g + s.ngood[(size_t)i]);
            if (status_out_) {
                const uint8_t* st = s.status + (size_t)i * cap_;
                (*status_out_)[(size_t)(batch.first + i - first)].assign(st, st + s.ngood[(size_t)i]);
                (*F_out_)[(size_t)(batch.first + i - first)].assign(s.F + (size_t)i * 9, s.F + (size_t)i * 9 + 9);
            }
        }
    }
    orbx_handle orb_;
    hamx_handle bf_;
    fmx_handle fm_;
    float ratio_;
    int batch_, w_, h_, cap_, depth_, device_;
    double max_distance_, confidence_;
    std::vector<std::vector<unsigned char> >* status_out_;
    std::vector<std::vector<double> >* F_out_;
    std::vector<Slot> slots_;
};
#endif

}  // namespace orbx_shim
