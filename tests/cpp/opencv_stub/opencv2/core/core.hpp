// tests/cpp/opencv_stub: the handful of OpenCV 2.4 core declarations include/crf_b200_compat.hpp touches when it is built with
// CRF_B200_WITH_OPENCV (namespace cvlite = cv).  TEST INFRASTRUCTURE: this image has no OpenCV C++; the stub exists so that the
// OpenCV branch of the header is compiled by the test-suite and cannot rot.  Signatures follow opencv2/core/core.hpp 2.4.9.
#pragma once
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32F 5
#define CV_32FC1 5
namespace cv {
template <typename _Tp> class Point_ { public: _Tp x, y; Point_() : x(0), y(0) {} Point_(_Tp _x, _Tp _y) : x(_x), y(_y) {} };
template <typename _Tp> class Rect_ { public: _Tp x, y, width, height; Rect_() : x(0), y(0), width(0), height(0) {} Rect_(_Tp _x, _Tp _y, _Tp _w, _Tp _h) : x(_x), y(_y), width(_w), height(_h) {} };
typedef Point_<int> Point;
typedef Rect_<int> Rect;
class Mat {
 public:
  Mat() : flags(0), rows(0), cols(0), data(0), step(0) {}
  Mat(int _rows, int _cols, int _type) : flags(0), rows(0), cols(0), data(0), step(0) { create(_rows, _cols, _type); }
  Mat(int _rows, int _cols, int _type, void* _data, size_t _step = 0) : flags(_type), rows(_rows), cols(_cols), data((unsigned char*)_data), step(_step ? _step : (size_t)_cols * esz(_type)) {}
  void create(int _rows, int _cols, int _type) { flags = _type; rows = _rows; cols = _cols; step = (size_t)_cols * esz(_type); buf.reset(new std::vector<unsigned char>((size_t)_rows * step + 16)); data = buf->data(); }
  int type() const { return flags; }
  int channels() const { return (flags >> 3) + 1; }
  bool empty() const { return data == 0 || rows * cols == 0; }
  template <typename _Tp> _Tp& at(int i0, int i1) { return *(_Tp*)(data + (size_t)i0 * step + (size_t)i1 * sizeof(_Tp)); }
  template <typename _Tp> const _Tp& at(int i0, int i1) const { return *(const _Tp*)(data + (size_t)i0 * step + (size_t)i1 * sizeof(_Tp)); }
  int flags, rows, cols;
  unsigned char* data;
  size_t step;
 private:
  static size_t esz(int t) { return ((t & 7) == 5 ? 4 : 1) * (size_t)((t >> 3) + 1); }
  std::shared_ptr<std::vector<unsigned char> > buf;
};
}  // namespace cv
