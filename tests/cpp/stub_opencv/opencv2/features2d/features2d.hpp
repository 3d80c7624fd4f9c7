// Minimal stand-in for <opencv2/features2d/features2d.hpp>: see ../core/core.hpp (test infrastructure only).
#pragma once
#include "../core/core.hpp"

namespace cv {
class KeyPoint {
public:
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};
struct DMatch {
    int queryIdx, trainIdx, imgIdx; float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(0) {}
};
}  // namespace cv
