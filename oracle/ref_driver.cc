// =============================================================================
// oracle/ref_driver.cc — TEST INFRASTRUCTURE ONLY.
//
// C entry points around the REAL reference code.  oracle/Makefile compiles the reference's own sources where they lie
// (/root/reference/src/{FaceForest,face_utils,ImageSample,HeadPoseSample,MPSample}.cpp with /root/reference/include),
// unmodified, against the type stand-ins of oracle/shim/ (this image has no OpenCV / Boost headers) into
// oracle/_ref/libcrf_ref.so.  Everything the reference computes itself on the path therefore runs as the reference
// wrote it: Forest::load / Tree::load through its own serialize() methods, Forest::evaluateMT, Tree::evaluateMT,
// TreeNode::eval, *Sample::eval / evalTest, ImageSample::evalTest, getHeadPoseVotesMT, getFacialFeaturesVotesMT,
// areaUnderCurve, MeanShift::shift, FaceForest::estimateHeadPose / estimateFacialFeatures / analyzeFace (composition,
// rescale), ThreadPool.  Only the third-party OpenCV arithmetic (cvtColor, resize, integral, filter2D, Sobel, ...) comes
// from the shim, which forwards to the oracle's cv2-4.13-pinned stage functions.
//
// Used by tests/ (oracle-vs-reference pins, golden generation) and bench.py's cpu_baseline; never by the product.
// `#define private public` below only opens the reference's classes to this driver (composed forest, jungle); the
// reference translation units themselves are compiled untouched.
// =============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include <opencv2/core/core.hpp>
#include <boost/crf_boost_shim.hpp>

#define private public
#include <FaceForest.hpp>
#include <face_utils.hpp>
#include <MeanShift.hpp>
#undef private

namespace {

struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };
struct Quiet {   // the reference PRINTs one line per tree while loading
  NullBuf nb; std::streambuf* old = nullptr;
  Quiet() { if (!std::getenv("CRF_REF_VERBOSE")) old = std::cout.rdbuf(&nb); }
  ~Quiet() { if (old) std::cout.rdbuf(old); }
};

std::string g_err;

// Leaf* -> Boost object id (= pre-order node index, the id the GPU path and the oracle report)
template <class S> void index_leaves(TreeNode<S>* root, std::unordered_map<const void*, int>& ids) {
  int next = 0;
  std::vector<TreeNode<S>*> st{root};
  while (!st.empty()) {
    TreeNode<S>* n = st.back(); st.pop_back();
    const int id = next++;
    if (n->isLeaf()) ids[n->getLeaf()] = id;
    else { st.push_back(n->right); st.push_back(n->left); }
  }
}

struct Ref {
  FaceForest* ff = nullptr;
  std::unordered_map<const void*, int> hp_ids, mp_ids;
  std::unordered_map<const void*, std::pair<int, int>> mp_tree_of;   // Tree* -> (pose forest, index)
  ~Ref() { delete ff; }
};

ForestParam make_param(const char* path, int ntrees, int max_depth) {
  // the fields loadConfigFile fills from data/config_{headpose,ffd}.txt (src/face_utils.cpp:50-140)
  ForestParam p;
  p.max_depth = max_depth; p.min_patches = 20; p.ntests = 2000; p.ntrees = ntrees; p.nimages = 400; p.npatches = 200;
  p.face_size = 125; p.patch_size_ratio = 0.25f; p.tree_path = path; p.image_path = "";
  p.features.push_back(0); p.features.push_back(1); p.features.push_back(2);
  return p;
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// FaceForest::FaceForest(FaceForestOptions) (src/FaceForest.cpp:15-58)
void* ref_create(const char* hp_dir, int hp_ntrees, const char* ffd_dir, int ffd_ntrees) {
  Quiet q;
  try {
    FaceForestOptions o;
    o.hp_forest_param = make_param(hp_dir, hp_ntrees, 15);
    o.mp_forest_param = make_param(ffd_dir, ffd_ntrees, 20);
    o.fd_option.path_face_cascade = "unused";
    Ref* r = new Ref();
    r->ff = new FaceForest(o);
    if (!r->ff->is_inizialized) { g_err = "FaceForest failed to initialise (see stderr)"; delete r; return nullptr; }
    for (int t = 0; t < r->ff->m_hp_forest.numberOfTrees(); t++) index_leaves<HeadPoseSample>(r->ff->m_hp_forest.getTree(t)->root, r->hp_ids);
    for (size_t f = 0; f < r->ff->m_mp_jungle.size(); f++)
      for (int t = 0; t < r->ff->m_mp_jungle[f].numberOfTrees(); t++) {
        Tree<MPSample>* tr = r->ff->m_mp_jungle[f].getTree(t);
        index_leaves<MPSample>(tr->root, r->mp_ids);
        r->mp_tree_of[tr] = std::make_pair((int)f, t);
      }
    return r;
  } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}
void ref_free(void* h) { delete (Ref*)h; }

void ref_set_strides(void* h, int hp_stride, int ffd_stride) {
  Ref* r = (Ref*)h;
  r->ff->m_options.hp_option.step_size = hp_stride;
  r->ff->m_options.mp_option.step_size = ffd_stride;
}

int ref_num_trees(void* h, int which) {
  Ref* r = (Ref*)h;
  if (which < 0) return r->ff->m_hp_forest.numberOfTrees();
  return which < (int)r->ff->m_mp_jungle.size() ? r->ff->m_mp_jungle[which].numberOfTrees() : -1;
}

// FaceForest::analyzeFace (src/FaceForest.cpp:183-258).  out_ffd: 10 x 2 ints (bbox-relative, original pixels);
// list_forest / list_tree (cap entries): the composed m_mp_forest after the call; returns its size.
int ref_analyze_face(void* h, const uint8_t* bgr, int rows, int cols, size_t step, int bx, int by, int bw, int bh, float* headpose, int* out_ffd,
                     int* list_forest, int* list_tree, int cap) {
  Ref* r = (Ref*)h;
  try {
    cv::Mat img(rows, cols, CV_8UC3, const_cast<uint8_t*>(bgr), step);
    Face face;
    r->ff->analyzeFace(img, cv::Rect(bx, by, bw, bh), face);
    if (headpose) *headpose = face.headpose;
    for (size_t i = 0; i < face.ffd_cordinates.size() && i < 10; i++) { out_ffd[2 * i] = face.ffd_cordinates[i].x; out_ffd[2 * i + 1] = face.ffd_cordinates[i].y; }
    const int n = r->ff->m_mp_forest.numberOfTrees();
    for (int i = 0; i < n && i < cap; i++) {
      const auto it = r->mp_tree_of.find(r->ff->m_mp_forest.getTree(i));
      if (list_forest) list_forest[i] = it->second.first;
      if (list_tree) list_tree[i] = it->second.second;
    }
    return n;
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// An ImageSample whose channels are the cv::integral of caller-supplied 8-bit planes [C][H][W] (m_feature_channels is a
// public member, include/ImageSample.hpp:177): lets the tests drive the reference's forest code with any channel data.
void* ref_sample_from_planes(const uint8_t* planes, int C, int H, int W) {
  cv::Mat dummy(H, W, CV_8UC1);
  ImageSample* s = new ImageSample(dummy, std::vector<int>(), true);
  for (int c = 0; c < C; c++) {
    cv::Mat p(H, W, CV_8UC1, const_cast<uint8_t*>(planes + (size_t)c * H * W));
    cv::Mat integral_img;
    cv::integral(p, integral_img, CV_32F);
    s->m_feature_channels.push_back(integral_img);
  }
  return s;
}
// ImageSample::ImageSample(img, features, use_integral = true) on a scaled gray face (src/ImageSample.cpp:11-20)
void* ref_sample_create(const uint8_t* gray, int H, int W, const int* features, int nfeatures) {
  cv::Mat img(H, W, CV_8UC1, const_cast<uint8_t*>(gray));
  return new ImageSample(img.clone(), std::vector<int>(features, features + nfeatures), true);
}
int ref_sample_channels(void* s) { return (int)((ImageSample*)s)->m_feature_channels.size(); }
// integral plane c as f32 [(H+1)][(W+1)]
void ref_sample_plane(void* s, int c, float* out) {
  const cv::Mat& m = ((ImageSample*)s)->m_feature_channels[c];
  for (int y = 0; y < m.rows; y++) std::memcpy(out + (size_t)y * m.cols, m.ptr<float>(y), sizeof(float) * (size_t)m.cols);
}
void ref_sample_free(void* s) { delete (ImageSample*)s; }

// ImageSample::evalTest(SimplePatchFeature, Rect) (src/ImageSample.cpp:30-64)
int ref_eval_test(void* s, int channel, const int* r1, const int* r2, int px, int py, int patch) {
  SimplePatchFeature f;
  f.feature_channel = channel;
  f.rect1 = cv::Rect(r1[0], r1[1], r1[2], r1[3]);
  f.rect2 = cv::Rect(r2[0], r2[1], r2[2], r2[3]);
  return ((ImageSample*)s)->evalTest(f, cv::Rect(px, py, patch, patch));
}

// Head pose: Forest<HeadPoseSample>::evaluateMT on the grid of getHeadPoseVotesMT (leaf ids, [patch][tree], x outer / y inner)
// and FaceForest::estimateHeadPose -> getHeadPoseVotesMT (src/face_utils.cpp:183-242) for mean and variance.
int ref_eval_hp(void* h, void* sample, int W, int H, int stride, int32_t* leaf_ids, float* headpose, float* variance) {
  Ref* r = (Ref*)h;
  ImageSample* s = (ImageSample*)sample;
  try {
    const Forest<HeadPoseSample>& forest = r->ff->m_hp_forest;
    const int patch = forest.getParam().getPatchSize(), nt = forest.numberOfTrees();
    int n = 0;
    if (leaf_ids) {
      std::vector<HeadPoseLeaf*> leafs((size_t)nt);
      for (int x = 0; x < W - patch; x += stride)
        for (int y = 0; y < H - patch; y += stride) {
          HeadPoseSample hs(s, cv::Rect(x, y, patch, patch));
          forest.evaluateMT(&hs, leafs.data());
          for (int t = 0; t < nt; t++) leaf_ids[(size_t)n * nt + t] = r->hp_ids.at(leafs[t]);
          n++;
        }
    }
    HeadPoseEstimatorOption o = r->ff->m_options.hp_option;
    o.step_size = stride;
    if (headpose && variance) FaceForest::estimateHeadPose(*s, cv::Rect(0, 0, W, H), forest, o, headpose, variance);
    return n;
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

float ref_area_under_curve(float x1, float x2, double mean, double std_) { return areaUnderCurve(x1, x2, mean, std_); }

// Facial features for an explicit composed forest: leaf ids as above, the vote lists of getFacialFeaturesVotesMT
// (src/face_utils.cpp:244-302) and MeanShift::shift per part (FaceForest::estimateFacialFeatures, src/FaceForest.cpp:74-95).
// votes_xyw: [10][vote_cap][3] (optional); n_votes[10]; rounded[10][2].
int ref_eval_ffd(void* h, void* sample, int W, int H, int stride, const int* forest_idx, const int* tree_idx, int ntrees, int32_t* leaf_ids,
                 int* n_votes, float* votes_xyw, int vote_cap, int* rounded) {
  Ref* r = (Ref*)h;
  ImageSample* s = (ImageSample*)sample;
  try {
    Forest<MPSample> forest;
    forest.setParam(r->ff->m_options.mp_forest_param);
    for (int i = 0; i < ntrees; i++) forest.addTree(r->ff->m_mp_jungle[forest_idx[i]].getTree(tree_idx[i]));
    const int patch = forest.getParam().getPatchSize();
    int n = 0;
    if (leaf_ids) {
      std::vector<MPLeaf*> leafs((size_t)ntrees);
      for (int x = 0; x < W - patch; x += stride)
        for (int y = 0; y < H - patch; y += stride) {
          MPSample ms(s, cv::Rect(x, y, patch, patch));
          forest.evaluateMT(&ms, leafs.data());
          for (int t = 0; t < ntrees; t++) leaf_ids[(size_t)n * ntrees + t] = r->mp_ids.at(leafs[t]);
          n++;
        }
    }
    MultiPartEstimatorOption o = r->ff->m_options.mp_option;
    o.step_size = stride;
    if (n_votes) {
      std::vector<std::vector<Vote> > votes(o.num_parts);
      getFacialFeaturesVotesMT(*s, forest, cv::Rect(0, 0, W, H), votes, o);
      for (int p = 0; p < o.num_parts && p < 10; p++) {
        n_votes[p] = (int)votes[p].size();
        for (int k = 0; votes_xyw && k < n_votes[p] && k < vote_cap; k++) {
          float* v = votes_xyw + ((size_t)p * vote_cap + k) * 3;
          v[0] = (float)votes[p][k].pos.x; v[1] = (float)votes[p][k].pos.y; v[2] = votes[p][k].weight;
        }
      }
    }
    if (rounded) {
      std::vector<cv::Point> ffd;
      FaceForest::estimateFacialFeatures(*s, cv::Rect(0, 0, W, H), forest, o, ffd);
      for (size_t p = 0; p < ffd.size() && p < 10; p++) { rounded[2 * p] = ffd[p].x; rounded[2 * p + 1] = ffd[p].y; }
    }
    return n;
  } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// MeanShift::shift (include/MeanShift.hpp:41-76) on a caller-supplied vote list.  rounded = the reference's result.
// mean_f / iters (optional) trace the same loop through the class's own getMean / getWeightedMean statics.
void ref_meanshift(const float* votes_xyw, int n, int* rounded, float* mean_f, int* iters) {
  std::vector<Vote> votes((size_t)n);
  for (int i = 0; i < n; i++) { votes[i].pos.x = (int)votes_xyw[3 * i]; votes[i].pos.y = (int)votes_xyw[3 * i + 1]; votes[i].weight = votes_xyw[3 * i + 2]; votes[i].check = true; }
  MeanShiftOption o;
  cv::Point_<int> result;
  MeanShift::shift(votes, result, o);
  rounded[0] = result.x; rounded[1] = result.y;
  if (mean_f || iters) {
    cv::Point_<float> mean;
    MeanShift::getMean(votes, mean);
    int it = 0;
    bool coverg = false;
    for (int i = 0; (i < o.max_iterations) && (coverg == false); i++) {
      cv::Point_<float> shifted_mean;
      MeanShift::getWeightedMean(votes, mean, o.kernel_size, shifted_mean);
      if (cv::norm(shifted_mean - mean) < o.stopping_criteria) coverg = true;
      mean = shifted_mean;
      it++;
    }
    if (mean_f) { mean_f[0] = mean.x; mean_f[1] = mean.y; }
    if (iters) *iters = it;
  }
}

}  // extern "C"
