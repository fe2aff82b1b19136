// Stand-in for the subset of OpenCV used by dense_mapping/test_monocular_mapping.cpp.
// TEST INFRASTRUCTURE ONLY; written from scratch.  The hot path touches only Mat::data,
// Mat::step, Mat::ptr<T>(row), rows, cols; the GUI / file functions used by main() and
// the viz helpers are inert stubs so that the unmodified translation unit links.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;

// OpenCV's type codes are global macros; global enumerators here, re-exported into cv:: for qualified use
enum { CV_8U = 0, CV_8UC1 = 0, CV_64F = 6, CV_8UC3 = 16 };

namespace cv {

using ::CV_8U; using ::CV_8UC1; using ::CV_64F; using ::CV_8UC3;
enum { IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };
enum { THRESH_BINARY_INV = 1 };
enum { COLOR_GRAY2BGR = 8 };

struct Rect { int x, y, width, height; Rect(int a, int b, int c, int d) : x(a), y(b), width(c), height(d) {} };
struct Point2f { float x, y; Point2f(float a, float b) : x(a), y(b) {} };
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; } };
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };

class Mat {
public:
    uchar *data;
    size_t step;
    int rows, cols;
    Mat() : data(nullptr), step(0), rows(0), cols(0), type_(0) {}
    Mat(int r, int c, int type) { alloc(r, c, type); }
    Mat(int r, int c, int type, double init) {
        alloc(r, c, type);
        if (type == CV_64F) { double *p = reinterpret_cast<double *>(data); for (size_t i = 0; i < size_t(r) * c; i++) p[i] = init; }
        else std::memset(data, int(init), size_t(r) * c);
    }
    Mat(Size sz, int type, const Scalar &init) {
        alloc(sz.height, sz.width, type);
        if (type == CV_64F) { double *p = reinterpret_cast<double *>(data); for (size_t i = 0; i < size_t(rows) * cols; i++) p[i] = init.v[0]; }
        else std::memset(data, int(init.v[0]), size_t(rows) * step);
    }
    Mat(int r, int c, int type, void *ext, size_t ext_step) : data(static_cast<uchar *>(ext)), step(ext_step), rows(r), cols(c), type_(type) {}
    Mat(const Mat &m, const Rect &roi) : data(m.data + size_t(roi.y) * m.step + size_t(roi.x) * m.elem()), step(m.step), rows(roi.height), cols(roi.width), type_(m.type_), own_(m.own_) {}
    template <typename T> T *ptr(int r) { return reinterpret_cast<T *>(data + size_t(r) * step); }
    template <typename T> const T *ptr(int r) const { return reinterpret_cast<const T *>(data + size_t(r) * step); }
    int type() const { return type_; }
    int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
    size_t elem() const { return type_ == CV_64F ? 8 : (type_ == CV_8UC3 ? 3 : 1); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    // Mat::convertTo(dst, CV_8U) of a CV_64F matrix: saturate_cast<uchar>(v) = clamp(round-half-even(v), 0, 255)
    // (cvRound -> lrint); dst may alias *this, as at ref:202.  Other conversions are not used by the reference file.
    void convertTo(Mat &dst, int rtype) const {
        if (type_ != CV_64F || rtype != CV_8U) { if (&dst != this) dst = *this; return; }
        Mat out(rows, cols, CV_8U);
        for (int y = 0; y < rows; y++) {
            const double *s = ptr<double>(y);
            uchar *d = out.ptr<uchar>(y);
            for (int x = 0; x < cols; x++) {
                const long v = std::lrint(s[x]);
                d[x] = uchar(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
        }
        dst = out;
    }
private:
    void alloc(int r, int c, int type) {
        type_ = type; rows = r; cols = c; step = size_t(c) * elem();
        own_.reset(new std::vector<uchar>(size_t(r) * step));
        data = own_->data();
    }
    int type_;
    std::shared_ptr<std::vector<uchar>> own_;
};

template <typename T> class Mat_ : public Mat {  // typed view; Mat converts implicitly as in OpenCV
public:
    Mat_() {}
    Mat_(const Mat &m) : Mat(m) {}
};

inline Mat operator*(const Mat &m, double) { return m; }
inline Mat operator-(const Mat &a, const Mat &) { return a; }
inline Mat imread(const std::string &, int) { return Mat(); }
inline bool imwrite(const std::string &, const Mat &) { return true; }
inline void imshow(const std::string &, const Mat &) {}
inline int waitKey(int) { return 0; }
// cv::threshold on a CV_64F matrix, THRESH_BINARY_INV: dst = src > thresh ? 0 : maxval (NaN compares false -> maxval)
inline double threshold(const Mat &src, Mat &dst, double thresh, double maxval, int type) {
    if (src.type() != CV_64F || type != THRESH_BINARY_INV) { dst = src; return thresh; }
    Mat out(src.rows, src.cols, CV_64F);
    for (int y = 0; y < src.rows; y++) {
        const double *s = src.ptr<double>(y);
        double *d = out.ptr<double>(y);
        for (int x = 0; x < src.cols; x++) d[x] = s[x] > thresh ? 0.0 : maxval;
    }
    dst = out;
    return thresh;
}
inline void cvtColor(const Mat &, Mat &, int) {}
inline void circle(Mat &, Point2f, int, const Scalar &, int) {}
inline void line(Mat &, Point2f, Point2f, const Scalar &, int) {}

}  // namespace cv
