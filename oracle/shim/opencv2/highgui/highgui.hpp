// oracle/shim/opencv2/highgui/highgui.hpp — TEST INFRASTRUCTURE ONLY (see core/core.hpp).
// UI / file calls of the reference's debug and training code: compile-only stubs, never on the inference path.
#ifndef CRF_SHIM_OPENCV_HIGHGUI_HPP
#define CRF_SHIM_OPENCV_HIGHGUI_HPP
#include <opencv2/core/core.hpp>
#include <opencv2/imgproc/imgproc.hpp>
namespace cv {
enum { IMREAD_COLOR = 1, WINDOW_AUTOSIZE = 1, FONT_HERSHEY_SIMPLEX = 0 };
inline void imshow(const std::string&, const Mat&) { shim_unsupported("cv::imshow"); }
inline int waitKey(int = 0) { shim_unsupported("cv::waitKey"); }
inline Mat imread(const std::string&, int = 1) { shim_unsupported("cv::imread"); }
inline void rectangle(Mat&, Rect, const Scalar&, int = 1) { shim_unsupported("cv::rectangle"); }
inline void circle(Mat&, Point, int, const Scalar&, int = 1) { shim_unsupported("cv::circle"); }
}  // namespace cv
#endif
