// Stand-in for <opencv2/core/core.hpp>, only for building the reference's vendored DBoW2 (ThirdParty/DBoW2) into
// oracle/_ref/libdbow_ref.so: this image has no OpenCV C++ headers.  TEST INFRASTRUCTURE, not part of the product.
// cv::Mat here is a plain owning byte matrix with the few members FORB.cpp and TemplatedVocabulary.h touch; cv::FileStorage /
// cv::FileNode exist so that the YAML save()/load() members compile (they are virtual, hence instantiated) and report "not
// opened" -- the harness loads vocabularies through the reference's text loader instead.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <cstdint>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5

namespace cv {

class Mat {
public:
    int rows = 0, cols = 0;
    Mat() {}
    void create(int r, int c, int type) { rows = r; cols = c; type_ = type; buf_.assign((size_t)r * c * (type == CV_32F ? 4 : 1), 0); }
    void create(size_t r, int c, int type) { create((int)r, c, type); }
    void release() { rows = cols = 0; buf_.clear(); }
    bool empty() const { return buf_.empty(); }
    Mat clone() const { return *this; }
    static Mat zeros(int r, int c, int type) { Mat m; m.create(r, c, type); return m; }
    template <typename T> T* ptr(int row = 0) { return reinterpret_cast<T*>(buf_.data() + (size_t)row * cols * (type_ == CV_32F ? 4 : 1)); }
    template <typename T> const T* ptr(int row = 0) const { return reinterpret_cast<const T*>(buf_.data() + (size_t)row * cols * (type_ == CV_32F ? 4 : 1)); }
private:
    int type_ = CV_8U;
    std::vector<unsigned char> buf_;
};

class FileNode {
public:
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator float() const { return 0.f; }
    operator double() const { return 0.; }
    operator std::string() const { return std::string(); }
};

class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage(const char*, int) {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
};
template <typename T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

}  // namespace cv
