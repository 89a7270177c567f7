// =============================================================================
// oracle/shim/cvshim.cc — TEST INFRASTRUCTURE ONLY.
//
// Out-of-line part of the OpenCV / Boost stand-ins (oracle/shim/opencv2, oracle/shim/boost) that let the UNMODIFIED
// reference sources compile here.  The image-processing ARITHMETIC is not defined in this file: every call forwards to
// the oracle's cv2-4.13-pinned stage functions (oracle/crf_oracle.cc, orc_cv_* / orc_bgr2gray / orc_resize); this file
// only adapts cv::Mat to dense planes.
//
// cv::filter2D: OpenCV evaluates kernels up to 7x7 directly (raster order, separate rounding: bit-exact here) and larger
// ones through a DFT, which no direct sum reproduces bit for bit.  Two stand-ins for the large kernels:
//   gabor mode 0 (default, "canonical"): a kernel that is a member of the Gabor bank of FeatureChannelFactory::initGaborKernels
//     is evaluated in the canonical separable arithmetic the oracle and the CUDA path share;
//   gabor mode 1 ("direct"): raster sum accumulated in double and rounded once — the closest f32 to the exact response.
// crf_ref_set_gabor_mode() switches; tests use mode 1 to measure how far the canonical choice is from a neutral one.
// =============================================================================
#include <opencv2/core/core.hpp>
#include <opencv2/imgproc/imgproc.hpp>
#include <boost/crf_boost_shim.hpp>

#include <atomic>
#include <mutex>

extern "C" {
void orc_bgr2gray(const uint8_t* bgr, int rows, int cols, size_t step, uint8_t* gray);
void orc_resize(const uint8_t* src, int sh, int sw, size_t sstep, uint8_t* dst, int dh, int dw);
int orc_gabor_bank(int* widths, float* re, float* im, int cap);
void orc_cv_integral(const uint8_t* src, int H, int W, float* dst);
void orc_cv_sobel(const uint8_t* src, int H, int W, int dx, int dy, uint8_t* dst);
void orc_cv_minmax(const uint8_t* src, int H, int W, uint8_t* mn, uint8_t* mx);
void orc_cv_equalize(const uint8_t* src, int H, int W, uint8_t* dst);
void orc_cv_canny(const uint8_t* src, int H, int W, int low, int high, uint8_t* dst);
void orc_cv_filter2d(const uint8_t* src, int H, int W, const float* kern, int kw, int mode, float* dst);
void orc_cv_gabor_canonical(const uint8_t* src, int H, int W, int index, float* re, float* im);
}

namespace {
std::atomic<int> g_gabor_mode{0};
std::atomic<int> g_threads{0};

// dense copy of an 8-bit single-channel Mat (ROI views have a larger step)
std::vector<uint8_t> dense_u8(const cv::Mat& m) {
  CV_Assert(m.type() == CV_8UC1);
  std::vector<uint8_t> d((size_t)m.rows * m.cols);
  for (int y = 0; y < m.rows; y++) std::memcpy(d.data() + (size_t)y * m.cols, m.data + (size_t)y * m.step, (size_t)m.cols);
  return d;
}
void store_u8(const std::vector<uint8_t>& d, int rows, int cols, cv::Mat& dst) {
  cv::Mat out(rows, cols, CV_8UC1);
  for (int y = 0; y < rows; y++) std::memcpy(out.data + (size_t)y * out.step, d.data() + (size_t)y * cols, (size_t)cols);
  dst = out;
}
void store_f32(const std::vector<float>& d, int rows, int cols, cv::Mat& dst) {
  cv::Mat out(rows, cols, CV_32FC1);
  for (int y = 0; y < rows; y++) std::memcpy(out.data + (size_t)y * out.step, d.data() + (size_t)y * cols, sizeof(float) * (size_t)cols);
  dst = out;
}

struct Bank {
  int widths[35];
  std::vector<float> re, im;
  std::vector<size_t> off;
  Bank() {
    const int n = orc_gabor_bank(widths, nullptr, nullptr, 0);
    re.resize((size_t)n); im.resize((size_t)n);
    orc_gabor_bank(widths, re.data(), im.data(), n);
    size_t o = 0;
    for (int i = 0; i < 35; i++) { off.push_back(o); o += (size_t)widths[i] * widths[i]; }
  }
  // index of the bank kernel equal to `k` (bit for bit), part = 0 real / 1 imaginary; -1 if none
  int find(const std::vector<float>& k, int kw, int& part) const {
    for (int i = 0; i < 35; i++) {
      if (widths[i] != kw) continue;
      if (std::memcmp(k.data(), re.data() + off[i], sizeof(float) * k.size()) == 0) { part = 0; return i; }
      if (std::memcmp(k.data(), im.data() + off[i], sizeof(float) * k.size()) == 0) { part = 1; return i; }
    }
    return -1;
  }
};
const Bank& bank() { static Bank b; return b; }
}  // namespace

extern "C" void crf_ref_set_gabor_mode(int mode) { g_gabor_mode = mode; }
extern "C" void crf_ref_set_threads(int n) { g_threads = n; }

unsigned boost::thread::hardware_concurrency() {
  const int n = g_threads.load();
  return n > 0 ? (unsigned)n : std::max(1u, std::thread::hardware_concurrency());
}

namespace cv {

void Mat::convertTo(Mat& dst, int rtype, double alpha, double beta) const {
  if (rtype == type() && alpha == 1 && beta == 0) { dst = clone(); return; }
  CV_Assert(type() == CV_32FC1 && (rtype & 7) == CV_8U);
  // cvtScale f32 -> u8: saturate_cast<uchar>(v * alpha + beta) in f32, round half to even (FeatureChannelFactory.hpp:276,283)
  Mat out(rows, cols, CV_8UC1);
  const float a = (float)alpha, b = (float)beta;
  for (int y = 0; y < rows; y++)
    for (int x = 0; x < cols; x++) {
      float q = at<float>(y, x) * a;
      if (b != 0.f) q = q + b;
      const long iv = std::lrintf(q);
      out.at<uchar>(y, x) = (uchar)std::min<long>(std::max<long>(iv, 0), 255);
    }
  dst = out;
}

Scalar sum(const Mat& m) {
  CV_Assert(m.type() == CV_8UC1);
  double s = 0;
  for (int y = 0; y < m.rows; y++) for (int x = 0; x < m.cols; x++) s += m.at<uchar>(y, x);
  return Scalar(s);
}

void add(const Mat& a, const Mat& b, Mat& dst) {
  CV_Assert(a.type() == CV_32FC1 && b.type() == CV_32FC1 && a.rows == b.rows && a.cols == b.cols);
  Mat out(a.rows, a.cols, CV_32FC1);
  for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) out.at<float>(y, x) = a.at<float>(y, x) + b.at<float>(y, x);
  dst = out;
}

void pow(const Mat& src, double power, Mat& dst) {
  CV_Assert(src.type() == CV_32FC1 && (power == 2 || power == 0.5));
  Mat out(src.rows, src.cols, CV_32FC1);
  for (int y = 0; y < src.rows; y++)
    for (int x = 0; x < src.cols; x++) {
      const float v = src.at<float>(y, x);
      out.at<float>(y, x) = power == 2 ? v * v : std::sqrt(v);
    }
  dst = out;
}

void normalize(const Mat& src, Mat& dst, double alpha, double beta, int norm_type, int) {
  CV_Assert(src.type() == CV_32FC1 && norm_type == NORM_MINMAX && src.rows * src.cols > 0);
  // cv::normalize(NORM_MINMAX): scale / shift in double, applied by convertTo as one single-rounded FMA in f32 (cv2 4.13)
  double smin = src.at<float>(0, 0), smax = smin;
  for (int y = 0; y < src.rows; y++)
    for (int x = 0; x < src.cols; x++) { const double v = src.at<float>(y, x); smin = std::min(smin, v); smax = std::max(smax, v); }
  const double dmin = std::min(alpha, beta), dmax = std::max(alpha, beta);
  const double scale = (dmax - dmin) * ((smax - smin) > 2.220446049250313e-16 ? 1. / (smax - smin) : 0.);
  const double shift = dmin - smin * scale;
  const float a = (float)scale, b = (float)shift;
  Mat out(src.rows, src.cols, CV_32FC1);
  for (int y = 0; y < src.rows; y++) for (int x = 0; x < src.cols; x++) out.at<float>(y, x) = std::fmaf(src.at<float>(y, x), a, b);
  dst = out;
}

void cvtColor(const Mat& src, Mat& dst, int code) {
  CV_Assert(code == COLOR_BGR2GRAY && src.type() == CV_8UC3);
  Mat out(src.rows, src.cols, CV_8UC1);
  orc_bgr2gray(src.data, src.rows, src.cols, src.step, out.data);
  dst = out;
}

void resize(const Mat& src, Mat& dst, Size dsize, double, double, int interpolation) {
  CV_Assert(src.type() == CV_8UC1 && interpolation == INTER_LINEAR && dsize.width > 0 && dsize.height > 0);
  Mat out(dsize.height, dsize.width, CV_8UC1);
  orc_resize(src.data, src.rows, src.cols, src.step, out.data, dsize.height, dsize.width);
  dst = out;
}

void integral(const Mat& src, Mat& sum_, int sdepth) {
  CV_Assert(sdepth == CV_32F);
  const std::vector<uint8_t> d = dense_u8(src);
  std::vector<float> o((size_t)(src.rows + 1) * (src.cols + 1));
  orc_cv_integral(d.data(), src.rows, src.cols, o.data());
  store_f32(o, src.rows + 1, src.cols + 1, sum_);
}

void equalizeHist(const Mat& src, Mat& dst) {
  const std::vector<uint8_t> d = dense_u8(src);
  std::vector<uint8_t> o(d.size());
  orc_cv_equalize(d.data(), src.rows, src.cols, o.data());
  store_u8(o, src.rows, src.cols, dst);
}

void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy) {
  CV_Assert(ddepth == CV_8U && dx + dy == 1);
  const std::vector<uint8_t> d = dense_u8(src);
  std::vector<uint8_t> o(d.size());
  orc_cv_sobel(d.data(), src.rows, src.cols, dx, dy, o.data());
  store_u8(o, src.rows, src.cols, dst);
}

static void check_ones3(const Mat& k) {
  CV_Assert(k.rows == 3 && k.cols == 3 && k.type() == CV_8UC1);
  for (int y = 0; y < 3; y++) for (int x = 0; x < 3; x++) CV_Assert(k.at<uchar>(y, x) == 1);
}
void erode(const Mat& src, Mat& dst, const Mat& kernel) {
  check_ones3(kernel);
  const std::vector<uint8_t> d = dense_u8(src);
  std::vector<uint8_t> o(d.size());
  orc_cv_minmax(d.data(), src.rows, src.cols, o.data(), nullptr);
  store_u8(o, src.rows, src.cols, dst);
}
void dilate(const Mat& src, Mat& dst, const Mat& kernel) {
  check_ones3(kernel);
  const std::vector<uint8_t> d = dense_u8(src);
  std::vector<uint8_t> o(d.size());
  orc_cv_minmax(d.data(), src.rows, src.cols, nullptr, o.data());
  store_u8(o, src.rows, src.cols, dst);
}

void Canny(const Mat& image, Mat& edges, double threshold1, double threshold2) {
  const std::vector<uint8_t> d = dense_u8(image);
  std::vector<uint8_t> o(d.size());
  // OpenCV floors the thresholds for the integer L1 magnitude: cvFloor(-1) = -1, cvFloor(5) = 5
  orc_cv_canny(d.data(), image.rows, image.cols, (int)std::floor(threshold1), (int)std::floor(threshold2), o.data());
  store_u8(o, image.rows, image.cols, edges);
}

void filter2D(const Mat& src, Mat& dst, int ddepth, const Mat& kernel) {
  CV_Assert(ddepth == CV_32F && kernel.type() == CV_32FC1 && kernel.rows == kernel.cols && (kernel.rows & 1));
  const int kw = kernel.rows, H = src.rows, W = src.cols;
  const std::vector<uint8_t> d = dense_u8(src);
  std::vector<float> k((size_t)kw * kw);
  for (int y = 0; y < kw; y++) for (int x = 0; x < kw; x++) k[(size_t)y * kw + x] = kernel.at<float>(y, x);
  std::vector<float> o((size_t)H * W);
  int part = 0, idx = -1;
  if (kw * kw >= 50 && g_gabor_mode.load() == 0) idx = bank().find(k, kw, part);   // OpenCV's direct / DFT switch is at 50 taps
  if (idx >= 0) {
    // gaborTransform (FeatureChannelFactory.hpp:265-266) asks for the real then the imaginary part of the same kernel on the
    // same image from the same thread: keep the pair of the last canonical evaluation
    static thread_local struct { const uchar* src = nullptr; int idx = -1, H = 0, W = 0; std::vector<float> re, im; } last;
    if (!(last.src == src.data && last.idx == idx && last.H == H && last.W == W) || part == 0) {
      last.re.resize(o.size()); last.im.resize(o.size());
      orc_cv_gabor_canonical(d.data(), H, W, idx, last.re.data(), last.im.data());
      last.src = src.data; last.idx = idx; last.H = H; last.W = W;
    }
    o = part == 0 ? last.re : last.im;
  } else {
    orc_cv_filter2d(d.data(), H, W, k.data(), kw, kw * kw >= 50 ? 1 : 0, o.data());
  }
  store_f32(o, H, W, dst);
}

}  // namespace cv
