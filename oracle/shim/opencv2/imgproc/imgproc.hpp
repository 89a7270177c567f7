// oracle/shim/opencv2/imgproc/imgproc.hpp — TEST INFRASTRUCTURE ONLY (see core/core.hpp).
// The imgproc calls of FaceForest::analyzeFace and FeatureChannelFactory::extractChannel; arithmetic = the oracle's
// cv2-4.13-pinned restatements (oracle/crf_oracle.cc, orc_cv_*), adapted to cv::Mat in oracle/shim/cvshim.cc.
#ifndef CRF_SHIM_OPENCV_IMGPROC_HPP
#define CRF_SHIM_OPENCV_IMGPROC_HPP
#include <opencv2/core/core.hpp>
namespace cv {
enum { COLOR_BGR2GRAY = 6, INTER_NEAREST = 0, INTER_LINEAR = 1 };
void cvtColor(const Mat& src, Mat& dst, int code);                                         // src/FaceForest.cpp:196
void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);  // :204
void integral(const Mat& src, Mat& sum, int sdepth = -1);                                  // FeatureChannelFactory.hpp:51 ...
void equalizeHist(const Mat& src, Mat& dst);                                               // :62
void filter2D(const Mat& src, Mat& dst, int ddepth, const Mat& kernel);                   // :265-266
void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy);                          // :124-125
void erode(const Mat& src, Mat& dst, const Mat& kernel);                                   // :149
void dilate(const Mat& src, Mat& dst, const Mat& kernel);                                  // :150
void Canny(const Mat& image, Mat& edges, double threshold1, double threshold2);           // :169
}  // namespace cv
#endif
