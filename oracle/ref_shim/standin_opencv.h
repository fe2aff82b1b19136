// Stand-in for the subset of OpenCV used by dense_mapping/test_monocular_mapping.cpp.
// TEST INFRASTRUCTURE ONLY; written from scratch.  The hot path touches only Mat::data,
// Mat::step, Mat::ptr<T>(row), rows, cols; the GUI / file functions used by main() and
// the viz helpers are inert stubs so that the unmodified translation unit links.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;

namespace cv {

enum { CV_8U = 0, CV_8UC1 = 0, CV_64F = 6 };
enum { IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };
enum { THRESH_BINARY_INV = 1 };
enum { COLOR_GRAY2BGR = 8 };

struct Rect { int x, y, width, height; Rect(int a, int b, int c, int d) : x(a), y(b), width(c), height(d) {} };
struct Point2f { float x, y; Point2f(float a, float b) : x(a), y(b) {} };
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; } };

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
    Mat(int r, int c, int type, void *ext, size_t ext_step) : data(static_cast<uchar *>(ext)), step(ext_step), rows(r), cols(c), type_(type) {}
    Mat(const Mat &m, const Rect &roi) : data(m.data + size_t(roi.y) * m.step + size_t(roi.x) * m.elem()), step(m.step), rows(roi.height), cols(roi.width), type_(m.type_), own_(m.own_) {}
    template <typename T> T *ptr(int r) { return reinterpret_cast<T *>(data + size_t(r) * step); }
    template <typename T> const T *ptr(int r) const { return reinterpret_cast<const T *>(data + size_t(r) * step); }
    int type() const { return type_; }
    size_t elem() const { return type_ == CV_64F ? 8 : 1; }
    void convertTo(Mat &, int) const {}
private:
    void alloc(int r, int c, int type) {
        type_ = type; rows = r; cols = c; step = size_t(c) * elem();
        own_.reset(new std::vector<uchar>(size_t(r) * step));
        data = own_->data();
    }
    int type_;
    std::shared_ptr<std::vector<uchar>> own_;
};

inline Mat operator*(const Mat &m, double) { return m; }
inline Mat operator-(const Mat &a, const Mat &) { return a; }
inline Mat imread(const std::string &, int) { return Mat(); }
inline bool imwrite(const std::string &, const Mat &) { return true; }
inline void imshow(const std::string &, const Mat &) {}
inline int waitKey(int) { return 0; }
inline double threshold(const Mat &, Mat &, double, double, int) { return 0; }
inline void cvtColor(const Mat &, Mat &, int) {}
inline void circle(Mat &, Point2f, int, const Scalar &, int) {}
inline void line(Mat &, Point2f, Point2f, const Scalar &, int) {}

}  // namespace cv
