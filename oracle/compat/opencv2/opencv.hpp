// opencv2/opencv.hpp -- TEST INFRASTRUCTURE: a declaration-only stand-in for the OpenCV names that
// homo/fhe_resize.h:23-141 and homo/fhe_decode.h use in their CLIENT-side display / comparison
// helpers (resize_image_opencv, compare_resize_opencv, show_image*, save_image_rgb).  The image has no
// C++ OpenCV; the server-side graphs of those headers (Cubic ... ResizeImage, homomorphic_sin ...)
// never call any of this, so the stand-ins only have to let the unchanged headers compile
// (`make -C oracle ref`); calling one throws.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_8UC3 16
#define CV_IMWRITE_PNG_COMPRESSION 16

namespace cv {
using std::vector;
typedef std::string String;
enum { IMREAD_COLOR = 1, INTER_LINEAR = 1, INTER_CUBIC = 2, WINDOW_AUTOSIZE = 1 };
[[noreturn]] inline void no_opencv() { throw std::logic_error("OpenCV is a client-side dependency of the reference; not available here"); }
struct Size {
    Size(int, int) {}
};
struct Vec3b {
    std::uint8_t val[3];
};
struct Mat {
    Mat() {}
    Mat(int, int, int, void *) {}
    std::uint8_t *data = nullptr;
    template <class T>
    T &at(int, int) { no_opencv(); }
};
inline Mat imread(const std::string &, int = 1) { no_opencv(); }
inline void resize(const Mat &, Mat &, Size, double, double, int) { no_opencv(); }
inline void namedWindow(const std::string &, int) { no_opencv(); }
inline void imshow(const std::string &, const Mat &) { no_opencv(); }
inline int waitKey(int) { no_opencv(); }
inline bool imwrite(const std::string &, const Mat &, const std::vector<int> & = std::vector<int>()) { no_opencv(); }
}  // namespace cv
