// =============================================================================
// oracle/shim/opencv2/core/core.hpp — TEST INFRASTRUCTURE ONLY.
//
// A minimal stand-in for the OpenCV 2.4 C++ *types and calls* that the reference's
// inference path uses, so that the UNMODIFIED reference sources under
// /root/reference/{include,src} compile in an image without OpenCV
// (oracle/Makefile target _ref/libcrf_ref.so).  Nothing under
// face_alignment_cvpr_2012_b200/ includes or links this.
//
// What is real and what is a compile-only stub:
//   * REAL (executed on the inference path): Mat (ref-counted, ROI views, at<T>, clone, convertTo, setTo),
//     Point_/Rect_/Size_/Scalar_, norm(Point_), cvtColor(BGR2GRAY), resize(INTER_LINEAR), integral(CV_32F),
//     filter2D(CV_32F), pow, add, normalize(NORM_MINMAX), Sobel(CV_8U), erode/dilate 3x3, equalizeHist, Canny,
//     sum.  Their ARITHMETIC is the oracle's (oracle/crf_oracle.cc exports orc_cv_*), i.e. cv2 4.13 semantics
//     pinned bit-exactly by tests/test_oracle_golden.py — this header only adapts types.
//   * COMPILE-ONLY (training / UI code that is never run here; they abort when called): imshow, waitKey,
//     rectangle, circle, imread, CascadeClassifier, getTickCount is real (std::chrono).
// =============================================================================
#ifndef CRF_SHIM_OPENCV_CORE_HPP
#define CRF_SHIM_OPENCV_CORE_HPP

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_PI 3.1415926535897932384626433832795
#define CV_Assert(expr) do { if (!(expr)) throw cv::Exception(std::string("CV_Assert failed: ") + #expr); } while (0)

namespace cv {

struct Exception : std::runtime_error { explicit Exception(const std::string& s) : std::runtime_error(s) {} };

[[noreturn]] inline void shim_unsupported(const char* what) {
  std::fprintf(stderr, "opencv shim: %s is a compile-only stub (not on the inference path)\n", what);
  std::abort();
}

template <typename T> inline T saturate_cast(float v);
template <> inline int saturate_cast<int>(float v) { return (int)std::lrintf(v); }   // cvRound: half to even
template <> inline float saturate_cast<float>(float v) { return v; }
template <typename T> inline T saturate_cast(double v);
template <> inline int saturate_cast<int>(double v) { return (int)std::lrint(v); }
template <> inline float saturate_cast<float>(double v) { return (float)v; }
template <> inline double saturate_cast<double>(double v) { return v; }
template <typename T> inline T saturate_cast(int v) { return (T)v; }

template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  // Point_<int> = Point_<float> rounds through saturate_cast (OpenCV 2.4 core/operations.hpp); include/MeanShift.hpp:75
  template <typename U> Point_(const Point_<U>& p) : x(saturate_cast<T>(p.x)), y(saturate_cast<T>(p.y)) {}
  Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
  Point_& operator-=(const Point_& o) { x -= o.x; y -= o.y; return *this; }
};
template <typename T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
// Point_<int> *= float: saturate_cast<int>(x * b) (src/FaceForest.cpp:257)
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, float b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> inline std::ostream& operator<<(std::ostream& o, const Point_<T>& p) { return o << "[" << p.x << ", " << p.y << "]"; }
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
// cv::norm(Point_<T>): sqrt((double)x*x + (double)y*y)  (include/MeanShift.hpp:71,122)
template <typename T> inline double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }

template <typename T> struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T> struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w_, T h_) : x(x_), y(y_), width(w_), height(h_) {}
};
template <typename T> inline std::ostream& operator<<(std::ostream& o, const Rect_<T>& r) { return o << "[" << r.width << " x " << r.height << " from (" << r.x << ", " << r.y << ")]"; }
typedef Rect_<int> Rect;

template <typename T> struct Scalar_ {
  T val[4];
  Scalar_() { val[0] = val[1] = val[2] = val[3] = 0; }
  Scalar_(T a, T b = 0, T c = 0, T d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  static Scalar_ all(T v) { return Scalar_(v, v, v, v); }
  T operator[](int i) const { return val[i]; }
  T& operator[](int i) { return val[i]; }
};
typedef Scalar_<double> Scalar;

template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8UC1 }; };
template <> struct DataType<char> { enum { type = 1 }; };
template <> struct DataType<unsigned short> { enum { type = 2 }; };
template <> struct DataType<short> { enum { type = 3 }; };
template <> struct DataType<int> { enum { type = CV_32SC1 }; };
template <> struct DataType<unsigned int> { enum { type = CV_32SC1 }; };
template <> struct DataType<float> { enum { type = CV_32FC1 }; };
template <> struct DataType<double> { enum { type = CV_64F }; };

class Mat {
 public:
  int rows, cols;
  size_t step;   // bytes per row
  uchar* data;
  Mat() : rows(0), cols(0), step(0), data(nullptr), type_(CV_8UC1) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
  Mat(int r, int c, int type, const Scalar& v) : Mat() { create(r, c, type); setTo(v); }
  // view of caller-owned memory (no copy)
  Mat(int r, int c, int type, void* ext, size_t step_ = 0) : rows(r), cols(c), step(step_ ? step_ : (size_t)c * esz(type)), data((uchar*)ext), type_(type) {}
  template <typename T> explicit Mat(const std::vector<T>& v) : Mat() {
    create((int)v.size(), 1, DataType<T>::type);
    for (size_t i = 0; i < v.size(); i++) at<T>((int)i, 0) = v[i];
  }
  void create(int r, int c, int type) {
    if (data && rows == r && cols == c && type_ == type && buf_) return;
    rows = r; cols = c; type_ = type; step = (size_t)c * esz(type);
    buf_ = std::make_shared<std::vector<uchar>>((size_t)r * step + 16, 0);
    data = buf_->data();
  }
  void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> 3) + 1; }
  size_t elemSize() const { return esz(type_); }
  bool isContinuous() const { return step == (size_t)cols * elemSize(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  Size size() const { return Size(cols, rows); }
  template <typename T> T& at(int y, int x) { return *reinterpret_cast<T*>(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const { return *reinterpret_cast<const T*>(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> T* ptr(int y) { return reinterpret_cast<T*>(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y) const { return reinterpret_cast<const T*>(data + (size_t)y * step); }
  Mat operator()(const Rect& r) const {   // ROI view sharing the buffer (src/FaceForest.cpp:199)
    CV_Assert(r.x >= 0 && r.y >= 0 && r.width >= 0 && r.height >= 0 && r.x + r.width <= cols && r.y + r.height <= rows);
    Mat m; m.rows = r.height; m.cols = r.width; m.step = step; m.type_ = type_; m.buf_ = buf_;
    m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
    return m;
  }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int y = 0; y < rows; y++) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
    return m;
  }
  Mat& setTo(const Scalar& s) {
    for (int y = 0; y < rows; y++)
      for (int x = 0; x < cols; x++)
        for (int c = 0; c < channels(); c++) {
          const size_t o = (size_t)y * step + ((size_t)x * channels() + c) * (elemSize() / channels());
          switch (depth()) {
            case CV_8U: data[o] = (uchar)s[c]; break;
            case CV_32S: *reinterpret_cast<int*>(data + o) = (int)s[c]; break;
            case CV_32F: *reinterpret_cast<float*>(data + o) = (float)s[c]; break;
            case CV_64F: *reinterpret_cast<double*>(data + o) = s[c]; break;
            default: shim_unsupported("Mat::setTo depth");
          }
        }
    return *this;
  }
  // convertTo(CV_8UC1, alpha): saturate_cast<uchar>(v * alpha) with round-half-even (FeatureChannelFactory.hpp:276,283)
  void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const;
  Mat t() const { Mat m(cols, rows, type_); for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) std::memcpy(m.data + (size_t)x * m.step + (size_t)y * elemSize(), data + (size_t)y * step + (size_t)x * elemSize(), elemSize()); return m; }
  Mat& operator/=(double d) {
    CV_Assert(depth() == CV_32F);
    for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) at<float>(y, x) = (float)(at<float>(y, x) / d);
    return *this;
  }
  static size_t esz(int type) {
    static const size_t d[8] = {1, 1, 2, 2, 4, 4, 8, 0};
    return d[type & 7] * (size_t)((type >> 3) + 1);
  }
 private:
  int type_;
  std::shared_ptr<std::vector<uchar>> buf_;
};
template <typename T> class Mat_ : public Mat {};

inline std::ostream& operator<<(std::ostream& o, const Mat& m) {
  o << "[";
  for (int y = 0; y < m.rows; y++)
    for (int x = 0; x < m.cols; x++) {
      if (m.depth() == CV_32S) o << m.at<int>(y, x); else if (m.depth() == CV_32F) o << m.at<float>(y, x); else if (m.depth() == CV_8U) o << (int)m.at<uchar>(y, x);
      o << (x + 1 < m.cols ? ", " : (y + 1 < m.rows ? "; " : ""));
    }
  return o << "]";
}

inline int64_t getTickCount() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline double getTickFrequency() { return 1e9; }

Scalar sum(const Mat& m);                       // src/ImageSample.cpp:41,44 (non-integral branch)
void add(const Mat& a, const Mat& b, Mat& dst);  // CV_32F element-wise (FeatureChannelFactory.hpp:269)
void pow(const Mat& src, double power, Mat& dst);  // power 2 -> x*x, power 0.5 -> sqrt (FeatureChannelFactory.hpp:267-270)
enum { NORM_MINMAX = 32 };
void normalize(const Mat& src, Mat& dst, double alpha = 1, double beta = 0, int norm_type = 4, int dtype = -1);

}  // namespace cv
#endif
