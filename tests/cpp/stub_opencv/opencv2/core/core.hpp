// Minimal stand-in for <opencv2/core/core.hpp> (OpenCV 2.4 layout), TEST INFRASTRUCTURE ONLY: it exists so that the
// ORBX_SHIM_USE_OPENCV branch of include/orbx_shim.hpp -- the one a maintainer of the reference compiles, against the real
// headers (src/CommonCV.h:9-21) -- is compiled by the test suite of an image that has no OpenCV C++ headers.  Only the
// members the shim touches exist; layouts follow cv::KeyPoint / cv::DMatch / cv::Mat's public fields.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <cstring>

#define CV_8U 0
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)

namespace cv {
typedef unsigned char uchar;
enum { NORM_HAMMING = 6 };
template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<double> Point3d;

class Mat {
public:
    int flags, rows, cols;
    uchar* data;
    struct Step { size_t v; operator size_t() const { return v; } } step;
    Mat() : flags(0), rows(0), cols(0), data(0), owned_(0) { step.v = 0; }
    Mat(int r, int c, int type, void* borrowed, size_t stp = 0) : flags(type), rows(r), cols(c), data((uchar*)borrowed), owned_(0)
    { step.v = stp ? stp : (size_t)c * elemSize(); }
    Mat(const Mat& o) : flags(0), rows(0), cols(0), data(0), owned_(0) { step.v = 0; *this = o; }
    Mat& operator=(const Mat& o)
    {
        if (this == &o) return *this;
        release();
        flags = o.flags; rows = o.rows; cols = o.cols; step = o.step;
        if (o.owned_) { owned_ = (uchar*)std::malloc(step.v * (size_t)rows + 1); std::memcpy(owned_, o.data, step.v * (size_t)rows); data = owned_; }
        else data = o.data;
        return *this;
    }
    ~Mat() { release(); }
    void create(int r, int c, int type)
    {
        release();
        flags = type; rows = r; cols = c; step.v = (size_t)c * elemSize();
        owned_ = (uchar*)std::calloc(step.v * (size_t)r + 1, 1);
        data = owned_;
    }
    int channels() const { return (flags >> CV_CN_SHIFT) + 1; }
    size_t elemSize() const { return (size_t)channels() * ((flags & 7) == CV_64F ? 8 : 1); }
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    uchar* ptr(int r) { return data + (size_t)r * step.v; }
    const uchar* ptr(int r) const { return data + (size_t)r * step.v; }
    template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step.v))[c]; }
    template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step.v))[c]; }

private:
    void release() { if (owned_) std::free(owned_); owned_ = 0; data = 0; }
    uchar* owned_;
};
}  // namespace cv
