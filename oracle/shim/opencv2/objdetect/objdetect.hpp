// oracle/shim/opencv2/objdetect/objdetect.hpp — TEST INFRASTRUCTURE ONLY (see core/core.hpp).
// cv::CascadeClassifier: load() only records the path (FaceForest's ctor requires it to succeed, src/FaceForest.cpp:23);
// detectMultiScale is a compile-only stub — the oracle build calls analyzeFace with given boxes, never analyzeImage.
#ifndef CRF_SHIM_OPENCV_OBJDETECT_HPP
#define CRF_SHIM_OPENCV_OBJDETECT_HPP
#include <opencv2/core/core.hpp>
namespace cv {
class CascadeClassifier {
 public:
  bool load(const std::string& path) { path_ = path; return true; }
  void detectMultiScale(const Mat&, std::vector<Rect>&, double = 1.1, int = 3, int = 0, Size = Size(), Size = Size()) { shim_unsupported("cv::CascadeClassifier::detectMultiScale"); }
 private:
  std::string path_;
};
}  // namespace cv
#endif
