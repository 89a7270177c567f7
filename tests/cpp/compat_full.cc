// Exercises the WHOLE reference-shaped interface of include/crf_b200_compat.hpp on a GPU, following the flow of the reference's
// own FaceForest::analyzeFace (src/FaceForest.cpp:183-258) step by step with the public classes: ImageSample, Forest::load,
// Forest::evaluateMT, Tree::evaluateMT / Tree::load, TreeNode, HeadPoseSample / MPSample, estimateHeadPose, getHeadPoseVotesMT,
// areaUnderCurve, forest composition through Forest::addTree / getTree, estimateFacialFeatures, getFacialFeaturesVotesMT,
// MeanShift::shift.  Prints `key value...` lines; tests/test_gpu_compat.py compares every one with the oracle.
// usage: compat_full <hp_dir> <ffd_dir>
#include <cmath>
#include <cstdio>
#include <vector>

#include "crf_b200_compat.hpp"

using namespace crf_b200;

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const std::string hp_dir = argv[1], ffd_dir = argv[2];
  ForestParam hp_param, mp_param;
  hp_param.tree_path = hp_dir; hp_param.ntrees = 15; hp_param.face_size = 125; hp_param.patch_size_ratio = 0.25f; hp_param.features = {0, 1, 2};
  mp_param = hp_param; mp_param.tree_path = ffd_dir; mp_param.ntrees = 20;

  // ---- FaceForest: the batch path
  FaceForestOptions ff_options;
  ff_options.hp_forest_param = hp_param;
  ff_options.mp_forest_param = mp_param;
  FaceForest ff(ff_options);
  if (!ff.is_inizialized) { std::puts("FAIL init"); return 1; }
  const int rows = 120, cols = 160;
  std::vector<unsigned char> px((size_t)rows * cols * 3);
  for (size_t i = 0; i < px.size(); i++) px[i] = (unsigned char)((i * 2654435761u) >> 24);
  cvlite::Mat img(rows, cols, CV_8UC3, px.data());
  const cvlite::Rect bbox(10, 5, 100, 100);
  Face face;
  ff.analyzeFace(img, bbox, face);
  std::printf("analyzeFace.headpose %.9g\n", face.headpose);
  std::printf("analyzeFace.ffd");
  for (const cvlite::Point& p : face.ffd_cordinates) std::printf(" %d %d", p.x, p.y);
  std::printf("\nanalyzeFace.composed");
  for (int i = 0; i < ff.composedForest().numberOfTrees(); i++) std::printf(" %d:%d", ff.composedForest().getTree(i)->which(), ff.composedForest().getTree(i)->index());
  std::printf("\n");

  // ---- the same face, step by step as src/FaceForest.cpp:196-257 does it
  // cvtColor + ROI + resize are OpenCV calls in the reference; here the library's stage does them
  std::vector<unsigned char> g(521 * 125);
  int W = 0, H = 0;
  check(crf_stage_gray_resize(ff.context(), px.data(), rows, cols, (size_t)cols * 3, crf_rect_t{bbox.x, bbox.y, bbox.width, bbox.height}, g.data(), &W, &H));
  cvlite::Mat img_scaled(H, W, CV_8UC1, g.data());
  const float scale = static_cast<float>(hp_param.face_size) / static_cast<float>(bbox.width);
  ImageSample sample(img_scaled, hp_param.features, true);
  std::printf("sample.channels %d\n", (int)sample.m_feature_channels.size());
  for (int c : {0, 5, 20, 37}) std::printf("sample.corner %d %.1f\n", c, sample.m_feature_channels[c].at<float>(H, W));
  SimplePatchFeature test;
  test.feature_channel = 9; test.rect1 = cvlite::Rect(3, 4, 10, 7); test.rect2 = cvlite::Rect(12, 15, 5, 13);
  std::printf("sample.evalTest %d\n", sample.evalTest(test, cvlite::Rect(20, 30, 31, 31)));
  {
    // the !m_use_integral branch (src/ImageSample.cpp:40-47): 8-bit feature channels, cv::sum over the rectangles
    ImageSample plain(img_scaled, hp_param.features, false);
    std::printf("plain.type8u %d\n", (int)(plain.m_feature_channels[9].type() == CV_8UC1));
    std::printf("plain.evalTest %d\n", plain.evalTest(test, cvlite::Rect(20, 30, 31, 31)));
  }

  Forest<HeadPoseSample> hp_forest;
  if (!hp_forest.load(hp_dir, hp_param)) { std::puts("FAIL hp load"); return 1; }
  std::printf("hp_forest.trees %d patch %d\n", hp_forest.numberOfTrees(), hp_forest.getParam().getPatchSize());
  HeadPoseSample hs(&sample, cvlite::Rect(12, 20, 31, 31));
  std::vector<HeadPoseLeaf*> leafs((size_t)hp_forest.numberOfTrees());
  hp_forest.evaluateMT(&hs, leafs.data());
  std::printf("hp_forest.evaluateMT");
  for (HeadPoseLeaf* l : leafs) std::printf(" %d:%.6g:%d", l->hp_nsamples, l->hp_foreground, l->hp_labels[2]);
  std::printf("\n");
  HeadPoseLeaf* one = nullptr;
  Tree<HeadPoseSample>::evaluateMT(&hs, hp_forest.getTree(3)->root, &one);
  std::printf("tree.evaluateMT.root %d\n", one == leafs[3]);
  TreeNode<HeadPoseSample>* r = hp_forest.getTree(3)->root;
  HeadPoseLeaf* via = nullptr;   // one level by hand (TreeNode::eval), the rest from the inner node
  Tree<HeadPoseSample>::evaluateMT(&hs, r->eval(&hs) ? r->left : r->right, &via);
  std::printf("tree.evaluateMT.inner %d\n", via == leafs[3]);
  std::printf("tree.root.split %d %d %d %d %d %d\n", r->split.feature.feature_channel, r->split.feature.rect1.x, r->split.feature.rect1.width, r->split.feature.rect2.y,
              r->split.feature.rect2.height, r->split.threshold);

  Tree<MPSample>* mpt = nullptr;
  if (!Tree<MPSample>::load(&mpt, ffd_dir + "/forest_1/tree_004.txt") || !mpt->isFinished()) { std::puts("FAIL tree load"); return 1; }
  MPSample ms(&sample, cvlite::Rect(40, 50, 31, 31));
  MPLeaf* ml = nullptr;
  Tree<MPSample>::evaluateMT(&ms, mpt->root, &ml);
  std::printf("mp_tree.evaluateMT %d %.6g %d %d %.6g\n", ml->mp_samples, ml->mp_foreground, ml->mp_parts_offset[3].x, ml->mp_parts_offset[3].y, ml->mp_prob_foreground[7]);
  Tree<MPSample>* missing = nullptr;
  std::printf("mp_tree.load.missing %d\n", (int)Tree<MPSample>::load(&missing, ffd_dir + "/forest_1/tree_999.txt"));

  float headpose = 0, variance = 0;
  FaceForest::estimateHeadPose(sample, cvlite::Rect(0, 0, W, H), hp_forest, HeadPoseEstimatorOption(), &headpose, &variance);
  std::printf("estimateHeadPose %.9g %.9g\n", headpose, variance);
  float hp2 = 0, var2 = 0;
  HeadPoseEstimatorOption dense; dense.step_size = 2;
  getHeadPoseVotesMT(sample, ff.headPoseForest(), cvlite::Rect(0, 0, W, H), &hp2, &var2, dense);
  std::printf("getHeadPoseVotesMT.step2 %.9g %.9g\n", hp2, var2);

  // src/FaceForest.cpp:214-250, verbatim but for the member names
  std::vector<Forest<MPSample> >& m_mp_jungle = ff.jungle();
  int hist_size = static_cast<int>(m_mp_jungle.size());
  std::vector<float> poseT(hist_size + 1);
  poseT[0] = -2.5; poseT[1] = -0.35; poseT[2] = -0.20; poseT[3] = -poseT[2]; poseT[4] = -poseT[1]; poseT[5] = -poseT[0];
  std::vector<float> pose_freq(hist_size);
  float max_area = 0;
  int dominant_headpose = 0;
  std::printf("areaUnderCurve");
  for (int j = 0; j < hist_size; j++) {
    float area = areaUnderCurve(poseT[j], poseT[j + 1], headpose, std::sqrt((double)variance));   // the double overload
    pose_freq[j] = area;
    std::printf(" %.9g", area);
    if (max_area < area) { max_area = area; dominant_headpose = j; }
  }
  std::printf("\n");
  Forest<MPSample> m_mp_forest;
  m_mp_forest.setParam(mp_param);
  m_mp_forest.cleanForest();
  for (unsigned i = 0; i < m_mp_jungle.size(); i++) {
    int ntrees = static_cast<int>(floor(pose_freq[i] * mp_param.ntrees));
    for (int j = 0; j < ntrees; j++) m_mp_forest.addTree(m_mp_jungle[i].getTree(j));
  }
  for (int i = m_mp_forest.numberOfTrees(); i < mp_param.ntrees; i++) m_mp_forest.addTree(m_mp_jungle[dominant_headpose].getTree(i));
  std::printf("composed");
  for (int i = 0; i < m_mp_forest.numberOfTrees(); i++) std::printf(" %d:%d", m_mp_forest.getTree(i)->which(), m_mp_forest.getTree(i)->index());
  std::printf("\n");

  std::vector<cvlite::Point> ffd;
  FaceForest::estimateFacialFeatures(sample, cvlite::Rect(0, 0, W, H), m_mp_forest, MultiPartEstimatorOption(), ffd);
  std::printf("estimateFacialFeatures");
  for (const cvlite::Point& p : ffd) std::printf(" %d %d", p.x, p.y);
  std::printf("\nrescaled");
  for (const cvlite::Point& p : ffd) std::printf(" %d %d", (int)std::lrintf(p.x * (1.0f / scale)), (int)std::lrintf(p.y * (1.0f / scale)));   // Point_<int> *= float
  std::printf("\n");

  std::vector<std::vector<Vote> > votes(10);
  getFacialFeaturesVotesMT(sample, m_mp_forest, cvlite::Rect(0, 0, W, H), votes, MultiPartEstimatorOption());
  std::printf("votes");
  for (const std::vector<Vote>& v : votes) {
    long long h = 0;
    for (const Vote& q : v) h = (h * 31 + (q.pos.x + 1000) * 7 + (q.pos.y + 1000) + (long long)(q.weight * 1024)) % 1000000007LL;
    std::printf(" %d:%lld", (int)v.size(), h);
  }
  std::printf("\nMeanShift.shift");
  MeanShiftOption ms_option;
  for (int i = 0; i < 10; i++) {
    cvlite::Point_<int> res;
    MeanShift::shift(votes[i], res, ms_option);
    std::printf(" %d %d", res.x, res.y);
  }
  std::printf("\n");
  MPSample one_patch(&sample, cvlite::Rect(33, 41, 31, 31));
  std::vector<MPLeaf*> mleafs((size_t)m_mp_forest.numberOfTrees());
  m_mp_forest.evaluateMT(&one_patch, mleafs.data());
  std::printf("mp_forest.evaluateMT");
  for (MPLeaf* l : mleafs) std::printf(" %d", l->mp_samples);
  // device = -1: the same analyzeImage with every visible GPU behind the one caller
  FaceForestOptions all = ff_options;
  all.device = -1;
  FaceForest ff_all(all);
  std::vector<Face> many;
  std::vector<cvlite::Rect> boxes{bbox, cvlite::Rect(40, 10, 90, 105), cvlite::Rect(0, 0, 120, 120), bbox};
  ff_all.analyzeImage(img, boxes, many);
  bool same = many.size() == 4 && many[0].headpose == face.headpose && many[3].headpose == face.headpose;
  for (size_t i = 0; same && i < 10; i++) same = many[0].ffd_cordinates[i].x == face.ffd_cordinates[i].x && many[3].ffd_cordinates[i].y == face.ffd_cordinates[i].y;
  std::printf("\nmulti.same %d\n", (int)same);
  if (argc >= 5) {   // <cascade.xml> <image.ppm>: FaceForest::analyzeImage(img, faces) as the reference's demo calls it (detectFace + analyzeFace)
    FILE* fp = std::fopen(argv[4], "rb");
    int pw = 0, ph = 0, maxv = 0;
    if (!fp || std::fscanf(fp, "P6 %d %d %d", &pw, &ph, &maxv) != 3) { std::puts("FAIL ppm"); return 1; }
    std::fgetc(fp);
    std::vector<unsigned char> rgb((size_t)pw * ph * 3), bgr((size_t)pw * ph * 3);
    if (std::fread(rgb.data(), 1, rgb.size(), fp) != rgb.size()) { std::puts("FAIL ppm data"); return 1; }
    std::fclose(fp);
    for (size_t i = 0; i < rgb.size(); i += 3) { bgr[i] = rgb[i + 2]; bgr[i + 1] = rgb[i + 1]; bgr[i + 2] = rgb[i]; }
    FaceForestOptions with_fd = ff_options;
    with_fd.fd_option.path_face_cascade = argv[3];
    FaceForest ffd(with_fd);
    if (!ffd.is_inizialized) { std::puts("FAIL cascade init"); return 1; }
    std::vector<Face> found;
    ffd.analyzeImage(cvlite::Mat(ph, pw, CV_8UC3, bgr.data()), found);
    std::printf("detect.faces %d", (int)found.size());
    for (const Face& f : found) std::printf(" %d %d %d %d %.9g %d %d", f.bbox.x, f.bbox.y, f.bbox.width, f.bbox.height, f.headpose, f.ffd_cordinates[0].x, f.ffd_cordinates[0].y);
    std::printf("\n");
    FaceForestOptions bad_fd = ff_options;
    bad_fd.fd_option.path_face_cascade = "/nonexistent/cascade.xml";
    FaceForest ffb(bad_fd);
    std::printf("detect.badcascade %d\n", (int)ffb.is_inizialized);
  }
  std::printf("compat_full ok\n");
  return 0;
}
