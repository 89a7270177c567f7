// crf_b200_compat.hpp — the reference's C++ interface for the inference path, re-exposed on top of the C ABI
// (include/crf_b200.h).  Header-only; link with -lcrf_b200.
//
// Same class and member names, argument meaning and error behaviour as MatrixPlayer/face_alignment_cvpr_2012
// (citations are file:line under the reference tree).  OpenCV types are replaced by layout-compatible PODs in
// namespace cvlite unless CRF_B200_WITH_OPENCV is defined, in which case cv::Mat / cv::Rect / cv::Point are used
// directly (the image this was built in has no OpenCV C++, so that branch is compiled only by the maintainer).
#ifndef CRF_B200_COMPAT_HPP
#define CRF_B200_COMPAT_HPP

#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "crf_b200.h"

#ifdef CRF_B200_WITH_OPENCV
#include <opencv2/core/core.hpp>
namespace cvlite = cv;
#else
namespace cvlite {
struct Rect { int x = 0, y = 0, width = 0, height = 0; Rect() {} Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {} };
struct Point { int x = 0, y = 0; Point() {} Point(int x_, int y_) : x(x_), y(y_) {} };
// 8-bit, 3-channel BGR image view (cv::Mat of type CV_8UC3)
struct Mat {
  unsigned char* data = nullptr; int rows = 0, cols = 0; size_t step = 0;
  Mat() {}
  Mat(int r, int c, unsigned char* d, size_t s = 0) : data(d), rows(r), cols(c), step(s ? s : (size_t)c * 3) {}
};
}  // namespace cvlite
#endif

namespace crf_b200 {

// include/Constants.hpp:24-60 (the fields inference reads)
struct ForestParam {
  int max_depth = 0, ntrees = 0, face_size = 125;
  float patch_size_ratio = 0.25f;
  std::string tree_path;
  std::vector<int> features;
};

// include/FaceForest.hpp:33-58, include/MeanShift.hpp:16-25
struct HeadPoseEstimatorOption { int num_head_pose_labels = 5; int step_size = 4; float min_foreground_probability = 0.5f; };
struct MultiPartEstimatorOption { int num_parts = 10; int step_size = 3; int min_samples = 2; float min_forground = 0.5f; float min_pf = 0.25f; float max_variance = 25.f; };
struct MeanShiftOption { int kernel_size = 10; int max_iterations = 7; float stopping_criteria = 0.05f; };

// include/FaceForest.hpp:60-68
struct FaceForestOptions {
  ForestParam head_pose_forest_param, mp_forest_param;
  HeadPoseEstimatorOption pose_option;
  MultiPartEstimatorOption multi_part_option;
  MeanShiftOption mean_shift_option;
  int device = 0;
};

// include/FaceForest.hpp:70-75
struct Face {
  float headpose = 0.f;
  cvlite::Rect bbox;
  std::vector<cvlite::Point> ffd_cordinates;
};

// include/MeanShift.hpp:27-50.  The GPU context is explicit (the reference's is a pure static).
struct Vote { cvlite::Point pos; float weight = 0.f; bool check = true; };  // include/face_utils.hpp:34-42

class FaceForest {
 public:
  FaceForest() {}
  explicit FaceForest(FaceForestOptions option) { load(option); }   // include/FaceForest.hpp:88-91
  ~FaceForest() { if (ctx_) crf_ctx_destroy(ctx_); if (model_) crf_model_free(model_); }
  FaceForest(const FaceForest&) = delete;
  FaceForest& operator=(const FaceForest&) = delete;

  // src/FaceForest.cpp:15-58: on failure prints the error and leaves the object un-initialised.
  bool load(const FaceForestOptions& o) {
    option_ = o;
    // a tree path that names a pre-packed image (*.crfb200, written by crf_model_save_packed) holds both forests
    const std::string& tp = o.mp_forest_param.tree_path;
    const bool packed = tp.size() > 8 && tp.compare(tp.size() - 8, 8, ".crfb200") == 0;
    int rc = packed ? crf_model_load_packed(tp.c_str(), &model_)
                    : crf_model_load(o.head_pose_forest_param.tree_path.c_str(), o.head_pose_forest_param.ntrees, o.mp_forest_param.tree_path.c_str(),
                                     o.mp_forest_param.ntrees, &model_);
    if (rc != CRF_OK) { std::fprintf(stderr, "(!) Error loading forest: %s\n", crf_last_error()); return false; }
    crf_options_t co;
    crf_options_default(&co);
    co.hp_stride = o.pose_option.step_size; co.hp_min_foreground = o.pose_option.min_foreground_probability;
    co.ffd_stride = o.multi_part_option.step_size; co.ffd_min_samples = o.multi_part_option.min_samples;
    co.ffd_min_foreground = o.multi_part_option.min_forground; co.ffd_min_pf = o.multi_part_option.min_pf; co.ffd_max_variance = o.multi_part_option.max_variance;
    co.ms_kernel_size = o.mean_shift_option.kernel_size; co.ms_max_iterations = o.mean_shift_option.max_iterations;
    co.ms_stopping_criteria = o.mean_shift_option.stopping_criteria;
    rc = crf_ctx_create(model_, o.device, &co, &ctx_);
    if (rc != CRF_OK) { std::fprintf(stderr, "(!) Error creating the GPU context: %s\n", crf_last_error()); return false; }
    is_inizialized = true;
    return true;
  }

  // src/FaceForest.cpp:183-258
  void analyzeFace(const cvlite::Mat img, cvlite::Rect face_bbox, Face& face) {
    require_init();
    std::vector<Face> faces;
    std::vector<cvlite::Rect> boxes(1, face_bbox);
    analyze(img, boxes, faces);
    face = faces[0];
  }

  // src/FaceForest.cpp:161-181 with detectFace()'s boxes supplied by the caller (one launch for all faces of the frame)
  void analyzeImage(const cvlite::Mat img, const std::vector<cvlite::Rect>& faces_bboxes, std::vector<Face>& faces) {
    require_init();
    analyze(img, faces_bboxes, faces);
  }

  bool is_inizialized = false;   // sic (include/FaceForest.hpp:153)
  crf_ctx* context() const { return ctx_; }

 private:
  void require_init() const {
    if (!is_inizialized) throw std::logic_error("CV_Assert(is_inizialized) failed (src/FaceForest.cpp:167,191)");
  }
  void analyze(const cvlite::Mat& img, const std::vector<cvlite::Rect>& boxes, std::vector<Face>& faces) {
    std::vector<crf_rect_t> r(boxes.size());
    for (size_t i = 0; i < boxes.size(); i++) r[i] = crf_rect_t{boxes[i].x, boxes[i].y, boxes[i].width, boxes[i].height};
    std::vector<crf_face_t> out(boxes.size());
    const int rc = crf_analyze_faces(ctx_, img.data, img.rows, img.cols, img.step, r.data(), (int)r.size(), out.data());
    if (rc != CRF_OK) throw std::runtime_error(crf_last_error());
    faces.resize(boxes.size());
    for (size_t i = 0; i < boxes.size(); i++) {
      faces[i].headpose = out[i].headpose;
      faces[i].bbox = boxes[i];
      faces[i].ffd_cordinates.resize(CRF_NUM_PARTS);
      for (int p = 0; p < CRF_NUM_PARTS; p++) faces[i].ffd_cordinates[p] = cvlite::Point(out[i].ffd[p][0], out[i].ffd[p][1]);
    }
  }
  FaceForestOptions option_;
  crf_model* model_ = nullptr;
  crf_ctx* ctx_ = nullptr;
};

// include/MeanShift.hpp:41-50
struct MeanShift {
  static void shift(crf_ctx* ctx, const std::vector<Vote>& votes, cvlite::Point& result) {
    std::vector<float> v;
    for (const Vote& q : votes)
      if (q.check) { v.push_back((float)q.pos.x); v.push_back((float)q.pos.y); v.push_back(q.weight); }
    int r[2] = {0, 0};
    if (crf_stage_meanshift(ctx, v.data(), (int)(v.size() / 3), nullptr, r, nullptr) != CRF_OK) throw std::runtime_error(crf_last_error());
    result = cvlite::Point(r[0], r[1]);
  }
};

}  // namespace crf_b200
#endif  // CRF_B200_COMPAT_HPP
