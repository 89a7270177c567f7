// =============================================================================
// crf_oracle.cc — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain C++17 restatement (no OpenCV, no Boost) of the reference's Conditional
// Regression Forest inference path.  It exists to CHECK the CUDA path; it is not
// the product and nothing under face_alignment_cvpr_2012_b200/ may link, import
// or call it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference leg use it.
//
// PARITY PINNED TO THE REFERENCE'S OWN CODE (round 2): oracle/_ref/libcrf_ref.so is the reference's
// src/{FaceForest,face_utils,ImageSample,HeadPoseSample,MPSample}.cpp + include/*.hpp compiled UNMODIFIED where
// they lie (oracle/Makefile `ref`) against type stand-ins for OpenCV / Boost (oracle/shim/; this image has neither).
// tests/test_reference_pin.py holds this restatement to it bit for bit: the 115 shipped archives read through the
// reference's serialize() methods, leaf ids, head-pose mean / variance, areaUnderCurve, forest composition, vote
// lists, MeanShift::shift, analyzeFace end to end.  The one thing the reference does NOT define itself is the
// OpenCV arithmetic underneath (cvtColor, resize, integral, Sobel, erode/dilate, equalizeHist, Canny, filter2D,
// normalize, convertTo): those stages are pinned bit-exactly against cv2 4.13 by tests/golden/make_golden.py +
// tests/test_oracle_golden.py, except filter2D with kernels >= 9x9, where cv2 (2.4.9 and 4.13) takes a DFT path no
// direct sum can match bit for bit — there the canonical arithmetic is DEFINED here (separable form) and its
// distance to cv2 and to a double-accumulated direct sum is pinned statistically (+-1 LSB of the u8 plane, rate
// < 2e-4).
//
// All file:line citations are relative to /root/reference/.
// =============================================================================
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <dirent.h>
#include <functional>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;
void set_err(const std::string& s) { g_err = s; }

// ----------------------------------------------------------------------------
// POD stand-ins for cv::Rect / cv::Point (include/opencv_serialization.hpp:65-79)
// ----------------------------------------------------------------------------
struct Rect { int x = 0, y = 0, width = 0, height = 0; };
struct Point { int x = 0, y = 0; };

// include/Constants.hpp:24-60
struct ForestParam {
  int max_depth = 0, min_patches = 0, ntests = 0, ntrees = 0, nimages = 0, npatches = 0, face_size = 0;
  float patch_size_ratio = 0.f;
  std::string tree_path, image_path;
  std::vector<int> features;
  int getPatchSize() const { return static_cast<int>(std::round(face_size * patch_size_ratio)); }
};

// include/ImageSample.hpp:22-91 (SimplePatchFeature), include/ThresholdSplit.hpp:51-67
struct Split {
  int feature_channel = 0;
  Rect rect1, rect2;
  double info = 0;
  int threshold = 0;
};

// include/HeadPoseSample.hpp:144-162 and include/MPSample.hpp:137-159 folded in one record.
struct Leaf {
  // head pose
  int hp_nsamples = 0;
  float hp_foreground = 0;
  int hp_labels[5] = {0, 0, 0, 0, 0};
  int hp_nlabels = 0;
  // multi part
  int mp_samples = 0;
  Point mp_parts_offset[10];
  float mp_parts_variance[10] = {0};
  float mp_prob_foreground[10] = {0};
  float mp_foreground = 0;
  int mp_nparts = 0;
  // bookkeeping for parity: Boost object id == pre-order index inside its tree
  int object_id = -1;
};

// include/TreeNode.hpp:138-146 — pointer-linked node with embedded leaf + split.
struct TreeNode {
  int depth = -1;
  bool is_leaf = false, has_split = false;
  Leaf leaf;
  Split split;
  TreeNode* left = nullptr;
  TreeNode* right = nullptr;
  int object_id = -1;
  ~TreeNode() { delete left; delete right; }
};

struct Tree {
  int m_num_nodes = 0, i_node = 0;
  ForestParam m_param;
  std::string m_save_path;
  TreeNode* root = nullptr;
  int n_nodes_parsed = 0, n_leaves_parsed = 0, max_depth_seen = 0;
  ~Tree() { delete root; }
  bool isFinished() const {  // include/Tree.hpp:72-79
    if (m_num_nodes == 0) return false;
    return i_node == m_num_nodes;
  }
};

// ----------------------------------------------------------------------------
// Boost text archive v10 reader (grammar: SURVEY Appendix B; field orders from
// Tree.hpp:334-343, Constants.hpp:44-59, TreeNode.hpp:148-164,
// ThresholdSplit.hpp:60-67, ImageSample.hpp:83-90, opencv_serialization.hpp:65-79,
// HeadPoseSample.hpp:154-161, MPSample.hpp:149-158).
// ----------------------------------------------------------------------------
struct Tokens {
  const char* p;
  const char* end;
  bool ok = true;
  void skip_ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p; }
  bool at_end() { skip_ws(); return p >= end; }
  long long i64() {
    skip_ws();
    if (p >= end) { ok = false; return 0; }
    char* q = nullptr;
    long long v = std::strtoll(p, &q, 10);
    if (q == p) { ok = false; return 0; }
    p = q;
    return v;
  }
  int i32() { return static_cast<int>(i64()); }
  double f64() {
    skip_ws();
    if (p >= end) { ok = false; return 0; }
    char* q = nullptr;
    double v = std::strtod(p, &q);
    if (q == p) { ok = false; return 0; }
    p = q;
    return v;
  }
  float f32() {
    skip_ws();
    if (p >= end) { ok = false; return 0; }
    char* q = nullptr;
    float v = std::strtof(p, &q);
    if (q == p) { ok = false; return 0; }
    p = q;
    return v;
  }
  std::string word() {
    skip_ws();
    const char* s = p;
    while (p < end && !(*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p;
    return std::string(s, p);
  }
  std::string str() {  // length-prefixed string
    long long n = i64();
    if (!ok || n < 0 || p >= end) { ok = false; return ""; }
    ++p;  // single separator
    if (p + n > end) { ok = false; return ""; }
    std::string s(p, p + n);
    p += n;
    return s;
  }
};

struct ClassSeen {
  bool node = false, split = false, feature = false, rect = false, hp_leaf = false, mp_leaf = false,
       vec_point = false, point = false;
};

enum Kind { KIND_HP = 0, KIND_MP = 1 };

static void class_info(Tokens& t, bool& seen) {
  if (!seen) { t.i32(); t.i32(); seen = true; }
}

static TreeNode* parse_nodeptr(Tokens& t, ClassSeen& cs, Kind kind, Tree* tree, int depth_guard) {
  if (depth_guard > 64) { t.ok = false; return nullptr; }
  int class_id = t.i32();
  if (!t.ok) return nullptr;
  if (class_id == -1) return nullptr;  // null pointer (never present in shipped files)
  class_info(t, cs.node);
  int object_id = t.i32();
  TreeNode* n = new TreeNode();
  n->object_id = object_id;
  n->depth = t.i32();
  n->is_leaf = t.i32() != 0;
  n->has_split = t.i32() != 0;
  tree->n_nodes_parsed++;
  tree->max_depth_seen = std::max(tree->max_depth_seen, n->depth);
  if (n->is_leaf) {
    tree->n_leaves_parsed++;
    Leaf& L = n->leaf;
    L.object_id = object_id;
    if (kind == KIND_HP) {
      class_info(t, cs.hp_leaf);
      L.hp_nsamples = t.i32();
      L.hp_foreground = t.f32();
      int cnt = t.i32();
      t.i32();  // item_version
      L.hp_nlabels = cnt;
      if (cnt < 0 || cnt > 5) { t.ok = false; return n; }
      for (int i = 0; i < cnt; i++) L.hp_labels[i] = t.i32();
    } else {
      class_info(t, cs.mp_leaf);
      L.mp_samples = t.i32();
      class_info(t, cs.vec_point);
      int cnt = t.i32();
      t.i32();
      if (cnt < 0 || cnt > 10) { t.ok = false; return n; }
      L.mp_nparts = cnt;
      for (int i = 0; i < cnt; i++) {
        class_info(t, cs.point);
        L.mp_parts_offset[i].x = t.i32();
        L.mp_parts_offset[i].y = t.i32();
      }
      int c2 = t.i32(); t.i32();
      if (c2 < 0 || c2 > 10) { t.ok = false; return n; }
      for (int i = 0; i < c2; i++) L.mp_parts_variance[i] = t.f32();
      int c3 = t.i32(); t.i32();
      if (c3 < 0 || c3 > 10) { t.ok = false; return n; }
      for (int i = 0; i < c3; i++) L.mp_prob_foreground[i] = t.f32();
      L.mp_foreground = t.f32();
    }
  }
  if (n->has_split) {
    class_info(t, cs.split);
    class_info(t, cs.feature);
    Split& s = n->split;
    s.feature_channel = t.i32();
    class_info(t, cs.rect);
    s.rect1.x = t.i32(); s.rect1.y = t.i32(); s.rect1.width = t.i32(); s.rect1.height = t.i32();
    s.rect2.x = t.i32(); s.rect2.y = t.i32(); s.rect2.width = t.i32(); s.rect2.height = t.i32();
    s.info = t.f64();
    s.threshold = t.i32();
  }
  if (!n->is_leaf) {
    n->left = parse_nodeptr(t, cs, kind, tree, depth_guard + 1);
    if (!t.ok) return n;
    n->right = parse_nodeptr(t, cs, kind, tree, depth_guard + 1);
  }
  return n;
}

// include/Tree.hpp:193-237 (Tree::load)
static Tree* tree_load(const std::string& path, Kind kind) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) { set_err("File not found: " + path); return nullptr; }
  std::fseek(f, 0, SEEK_END);
  long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::string buf(static_cast<size_t>(sz), '\0');
  if (sz > 0 && std::fread(&buf[0], 1, static_cast<size_t>(sz), f) != static_cast<size_t>(sz)) {
    std::fclose(f); set_err("short read: " + path); return nullptr;
  }
  std::fclose(f);
  Tokens t{buf.data(), buf.data() + buf.size()};
  if (t.word() != "22" || t.word() != "serialization::archive") { set_err("bad archive header: " + path); return nullptr; }
  t.i32();  // archive version (10)
  Tree* tree = new Tree();
  t.i32(); t.i32();  // class info Tree
  tree->m_num_nodes = t.i32();
  tree->i_node = t.i32();
  t.i32(); t.i32();  // class info ForestParam
  ForestParam& fp = tree->m_param;
  fp.max_depth = t.i32(); fp.min_patches = t.i32(); fp.ntests = t.i32(); fp.ntrees = t.i32();
  fp.nimages = t.i32(); fp.npatches = t.i32(); fp.face_size = t.i32(); fp.patch_size_ratio = t.f32();
  fp.tree_path = t.str(); fp.image_path = t.str();
  int nf = t.i32(); t.i32();
  if (!t.ok || nf < 0 || nf > 64) { delete tree; set_err("bad ForestParam: " + path); return nullptr; }
  for (int i = 0; i < nf; i++) fp.features.push_back(t.i32());
  tree->m_save_path = t.str();
  ClassSeen cs;
  tree->root = parse_nodeptr(t, cs, kind, tree, 0);
  if (!t.ok || !tree->root) { delete tree; set_err("Exception during tree serialization: " + path); return nullptr; }
  if (!t.at_end()) { delete tree; set_err("trailing tokens: " + path); return nullptr; }
  return tree;
}

// include/Forest.hpp:81-153
struct Forest {
  std::vector<Tree*> m_trees;  // non-owning, never freed — as in the reference (Forest.hpp:183)
  ForestParam m_forest_param;
  int numberOfTrees() const { return static_cast<int>(m_trees.size()); }
  bool load(const std::string& path, const ForestParam& fp, Kind kind, int max_trees = -1) {
    m_forest_param = fp;
    if (max_trees == -1) max_trees = fp.ntrees;
    for (int i = 0; i < fp.ntrees; i++) {
      if (numberOfTrees() > max_trees) continue;
      char buffer[1024];
      std::snprintf(buffer, sizeof buffer, "%s/tree_%03d.txt", path.c_str(), i);
      Tree* tree = tree_load(buffer, kind);
      if (!tree) return false;
      if (!tree->isFinished()) { delete tree; set_err(std::string("Tree is not finished successfully: ") + buffer); return false; }
      m_trees.push_back(tree);
    }
    return true;
  }
};

// ----------------------------------------------------------------------------
// Image stages (OpenCV semantics = cv2 4.13)
// ----------------------------------------------------------------------------
struct Plane8 { int rows = 0, cols = 0; std::vector<uint8_t> d; uint8_t& at(int r, int c) { return d[(size_t)r * cols + c]; } uint8_t at(int r, int c) const { return d[(size_t)r * cols + c]; } };
struct PlaneF { int rows = 0, cols = 0; std::vector<float> d; float& at(int r, int c) { return d[(size_t)r * cols + c]; } float at(int r, int c) const { return d[(size_t)r * cols + c]; } };

// src/FaceForest.cpp:196  cv::cvtColor(BGR2GRAY): 15-bit fixed point (A.1)
static inline uint8_t bgr2gray_px(int b, int g, int r) {
  return static_cast<uint8_t>((r * 9798 + g * 19235 + b * 3735 + 16384) >> 15);
}

static inline int border101(int p, int len) {  // BORDER_REFLECT_101
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * len - 2 - p;
  }
  return p;
}

// src/FaceForest.cpp:199-204  ROI + cv::resize(INTER_LINEAR), 8-bit fixed-point path (A.2)
static void resize_linear_u8(const uint8_t* src, int sh, int sw, size_t sstep, Plane8& dst, int dh, int dw) {
  dst.rows = dh; dst.cols = dw; dst.d.assign((size_t)dh * dw, 0);
  if (dh <= 0 || dw <= 0) return;
  const int SCALE = 2048;
  double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
  double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<short> alpha((size_t)dw * 2), beta((size_t)dh * 2);
  for (int dx = 0; dx < dw; dx++) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)std::floor(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx;
    alpha[dx * 2] = (short)std::lrint((1.f - fx) * SCALE);   // saturate_cast<short>(float): cvRound
    alpha[dx * 2 + 1] = (short)std::lrint(fx * SCALE);
  }
  for (int dy = 0; dy < dh; dy++) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)std::floor(fy);
    fy -= sy;
    yofs[dy] = sy;
    beta[dy * 2] = (short)std::lrint((1.f - fy) * SCALE);
    beta[dy * 2 + 1] = (short)std::lrint(fy * SCALE);
  }
  std::vector<int> row0(dw), row1(dw);
  for (int dy = 0; dy < dh; dy++) {
    int sy0 = std::min(std::max(yofs[dy], 0), sh - 1);
    int sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
    const uint8_t* S0 = src + (size_t)sy0 * sstep;
    const uint8_t* S1 = src + (size_t)sy1 * sstep;
    for (int dx = 0; dx < dw; dx++) {
      int sx = xofs[dx];
      int sx1 = std::min(sx + 1, sw - 1);
      int a0 = alpha[dx * 2], a1 = alpha[dx * 2 + 1];
      row0[dx] = S0[sx] * a0 + S0[sx1] * a1;
      row1[dx] = S1[sx] * a0 + S1[sx1] * a1;
    }
    int b0 = beta[dy * 2], b1 = beta[dy * 2 + 1];
    for (int dx = 0; dx < dw; dx++) {
      int v = (((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2;
      dst.at(dy, dx) = (uint8_t)std::min(std::max(v, 0), 255);
    }
  }
}

// cv::integral(src, dst, CV_32F) (FeatureChannelFactory.hpp:51 etc.), sums held in f32 (A.3)
static void integral_f32(const Plane8& src, PlaneF& dst) {
  dst.rows = src.rows + 1; dst.cols = src.cols + 1;
  dst.d.assign((size_t)dst.rows * dst.cols, 0.f);
  for (int y = 0; y < src.rows; y++) {
    float s = 0.f;
    for (int x = 0; x < src.cols; x++) {
      s += (float)src.at(y, x);
      dst.at(y + 1, x + 1) = dst.at(y, x + 1) + s;
    }
  }
}

// FeatureChannelFactory.hpp:120-125  cv::Sobel(img, dst, CV_8U, dx, dy), 3x3, REFLECT_101, saturate (A.5)
static void sobel_u8(const Plane8& src, Plane8& dst, int dx, int dy) {
  dst.rows = src.rows; dst.cols = src.cols; dst.d.assign(src.d.size(), 0);
  for (int y = 0; y < src.rows; y++) {
    int ym = border101(y - 1, src.rows), yp = border101(y + 1, src.rows);
    for (int x = 0; x < src.cols; x++) {
      int xm = border101(x - 1, src.cols), xp = border101(x + 1, src.cols);
      int v;
      if (dx == 1 && dy == 0)
        v = (src.at(ym, xp) + 2 * src.at(y, xp) + src.at(yp, xp)) - (src.at(ym, xm) + 2 * src.at(y, xm) + src.at(yp, xm));
      else
        v = (src.at(yp, xm) + 2 * src.at(yp, x) + src.at(yp, xp)) - (src.at(ym, xm) + 2 * src.at(ym, x) + src.at(ym, xp));
      dst.at(y, x) = (uint8_t)std::min(std::max(v, 0), 255);
    }
  }
}

// FeatureChannelFactory.hpp:142-150  erode / dilate with 3x3 ones; border never wins (A.5b)
static void minmax3x3_u8(const Plane8& src, Plane8& mn, Plane8& mx) {
  mn.rows = mx.rows = src.rows; mn.cols = mx.cols = src.cols;
  mn.d.assign(src.d.size(), 0); mx.d.assign(src.d.size(), 0);
  for (int y = 0; y < src.rows; y++)
    for (int x = 0; x < src.cols; x++) {
      int lo = 255, hi = 0;
      for (int j = -1; j <= 1; j++)
        for (int i = -1; i <= 1; i++) {
          int yy = y + j, xx = x + i;
          if (yy < 0 || yy >= src.rows || xx < 0 || xx >= src.cols) continue;
          int v = src.at(yy, xx);
          lo = std::min(lo, v); hi = std::max(hi, v);
        }
      mn.at(y, x) = (uint8_t)lo; mx.at(y, x) = (uint8_t)hi;
    }
}

// FeatureChannelFactory.hpp:58-70  cv::equalizeHist (FC_NORM): histogram, first non-empty bin i0, scale = 255 / (N - hist[i0]) in f32,
// lut[i] = cvRound(cumulative(i0 < j <= i) * scale), lut[i0] = 0; a constant image maps to itself.
static void equalize_hist_u8(const Plane8& src, Plane8& dst) {
  dst.rows = src.rows; dst.cols = src.cols; dst.d.assign(src.d.size(), 0);
  int hist[256] = {0};
  for (uint8_t v : src.d) hist[v]++;
  int i = 0;
  while (i < 256 && !hist[i]) ++i;
  const int total = (int)src.d.size();
  if (i == 256) return;
  if (hist[i] == total) { std::fill(dst.d.begin(), dst.d.end(), (uint8_t)i); return; }
  uint8_t lut[256] = {0};
  const float scale = (256 - 1.f) / (total - hist[i]);
  int sum = 0;
  for (lut[i++] = 0; i < 256; ++i) {
    sum += hist[i];
    long v = std::lrintf(sum * scale);
    lut[i] = (uint8_t)std::min<long>(std::max<long>(v, 0), 255);
  }
  for (size_t p = 0; p < src.d.size(); p++) dst.d[p] = lut[src.d[p]];
}

// FeatureChannelFactory.hpp:166-179  cv::Canny(img, canny_img, -1, 5) (FC_CANNY): aperture 3, L1 gradient magnitude,
// thresholds low = -1 / high = 5 (every local maximum of the gradient is a candidate; those above 5 seed the hysteresis).
// Restates OpenCV's canny.cpp: 16-bit Sobel with BORDER_REPLICATE, magnitude |dx| + |dy| with a zero border, non-maximum
// suppression in the three direction classes (TG22 = tan(22.5 deg) in 15-bit fixed point) with its asymmetric > / >=
// comparisons, then 8-connected hysteresis from the strong pixels.  Pinned bit-exactly against cv2 4.13
// (tests/test_oracle_golden.py::test_canny_matches_cv2).
static void canny_u8(const Plane8& src, Plane8& dst, int low = -1, int high = 5) {
  const int H = src.rows, W = src.cols, P = W + 2;
  dst.rows = H; dst.cols = W; dst.d.assign(src.d.size(), 0);
  auto px = [&](int y, int x) { return (int)src.at(std::min(std::max(y, 0), H - 1), std::min(std::max(x, 0), W - 1)); };
  std::vector<int> dx((size_t)H * W), dy((size_t)H * W), mag((size_t)(H + 2) * P, 0);
  std::vector<uint8_t> mp((size_t)(H + 2) * P, 1);   // 1 = not an edge, 0 = candidate, 2 = edge
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const int gx = (px(y - 1, x + 1) + 2 * px(y, x + 1) + px(y + 1, x + 1)) - (px(y - 1, x - 1) + 2 * px(y, x - 1) + px(y + 1, x - 1));
      const int gy = (px(y + 1, x - 1) + 2 * px(y + 1, x) + px(y + 1, x + 1)) - (px(y - 1, x - 1) + 2 * px(y - 1, x) + px(y - 1, x + 1));
      dx[(size_t)y * W + x] = gx; dy[(size_t)y * W + x] = gy;
      mag[(size_t)(y + 1) * P + x + 1] = std::abs(gx) + std::abs(gy);
    }
  const int TG22 = (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5);
  std::vector<int> stack;
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const size_t c = (size_t)(y + 1) * P + x + 1;
      const int m = mag[c];
      bool ok = false;
      if (m > low) {
        const int xs = dx[(size_t)y * W + x], ys = dy[(size_t)y * W + x];
        const int ax = std::abs(xs), ay = std::abs(ys) << 15;
        const int tg22x = ax * TG22;
        if (ay < tg22x) ok = m > mag[c - 1] && m >= mag[c + 1];
        else {
          const int tg67x = tg22x + (ax << 16);
          if (ay > tg67x) ok = m > mag[c - P] && m >= mag[c + P];
          else { const int sg = (xs ^ ys) < 0 ? -1 : 1; ok = m > mag[c - P - sg] && m > mag[c + P + sg]; }
        }
      }
      if (ok) { mp[c] = m > high ? 2 : 0; if (m > high) stack.push_back((int)c); }
    }
  while (!stack.empty()) {
    const int c = stack.back(); stack.pop_back();
    for (int j = -1; j <= 1; j++)
      for (int i = -1; i <= 1; i++) {
        const int q = c + j * P + i;
        if (mp[q] == 0) { mp[q] = 2; stack.push_back(q); }
      }
  }
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) dst.at(y, x) = mp[(size_t)(y + 1) * P + x + 1] == 2 ? 255 : 0;
}

// FeatureChannelFactory.hpp:186-251  createKernel / initGaborKernels (A.4)
struct GaborKernel {
  int width = 0;
  std::vector<float> re, im;   // 2-D tables as createKernel builds them, row-major [row j][col i]
  // separable factors of the same kernel (used for width >= 9, see gabor_response_separable):
  //   t1(x,y) e^{i(ax x + ay y)} = [g(x) e^{i ax x}] [g(y) e^{i ay y}],  g(x) = sqrt(k^2/sigma^2) exp(-x^2 k^2 / (2 sigma^2))
  std::vector<float> hx_re, hx_im, hy_re, hy_im, g1;
  float dc = 0.f;              // exp(-sigma^2 / 2): the real part subtracts dc * t1(x,y)
};
static void create_gabor_kernel(int iMu, int iNu, double sigma, double dF, GaborKernel& out) {
  double F = dF;
  double k = (M_PI / 2) / std::pow(F, (double)iNu);
  double phi = M_PI * iMu / 8;
  double width = std::round((sigma / k) * 6 + 1);
  if (std::fmod(width, 2.0) == 0.0) width++;
  int w = (int)width;
  out.width = w;
  out.re.assign((size_t)w * w, 0.f);
  out.im.assign((size_t)w * w, 0.f);
  int off_set = (int)((width - 1) / 2);
  for (int i = 0; i < w; i++)
    for (int j = 0; j < w; j++) {
      int x = i - off_set, y = j - off_set;
      double dTemp1 = (std::pow(k, 2) / std::pow(sigma, 2)) *
                      std::exp(-(std::pow((double)x, 2) + std::pow((double)y, 2)) * std::pow(k, 2) / (2 * std::pow(sigma, 2)));
      double dTemp2 = std::cos(k * std::cos(phi) * x + k * std::sin(phi) * y) - std::exp(-(std::pow(sigma, 2) / 2));
      double dTemp3 = std::sin(k * std::cos(phi) * x + k * std::sin(phi) * y);
      out.re[(size_t)j * w + i] = (float)(dTemp1 * dTemp2);
      out.im[(size_t)j * w + i] = (float)(dTemp1 * dTemp3);
    }
  out.hx_re.resize(w); out.hx_im.resize(w); out.hy_re.resize(w); out.hy_im.resize(w); out.g1.resize(w);
  const double ax = k * std::cos(phi), ay = k * std::sin(phi);
  for (int i = 0; i < w; i++) {
    const double x = (double)(i - off_set);
    const double env = std::sqrt(k * k / (sigma * sigma)) * std::exp(-(x * x) * k * k / (2 * sigma * sigma));
    out.g1[i] = (float)env;
    out.hx_re[i] = (float)(env * std::cos(ax * x)); out.hx_im[i] = (float)(env * std::sin(ax * x));
    out.hy_re[i] = (float)(env * std::cos(ay * x)); out.hy_im[i] = (float)(env * std::sin(ay * x));
  }
  out.dc = (float)std::exp(-(sigma * sigma) / 2);
}
static const std::vector<GaborKernel>& gabor_bank() {
  static std::vector<GaborKernel> bank = [] {
    std::vector<GaborKernel> b;
    double sigma = 1.0 / 2.0 * M_PI;
    double dF = std::sqrt(2.0);
    for (int iNu = 0; iNu <= 4; iNu++)
      for (int iMu = 0; iMu < 7; iMu++) {
        b.emplace_back();
        create_gabor_kernel(iMu, iNu, sigma, dF, b.back());
      }
    return b;
  }();
  return bank;
}

// cv::filter2D(src u8 -> CV_32F, kernel), correlation, anchor centre, REFLECT_101.
// CANONICAL accumulation: raster order over the kernel, acc starts at 0.
// 7x7 kernels: acc = acc + (float)px * k with separately rounded product and sum — bit-identical to
// cv2 4.13.  9x9 and larger: cv2 (2.4.9 and 4.13 alike) switches to a DFT path that no direct sum can
// match bit for bit, so the canonical form there is acc = fmaf(px, k, acc) (one rounding per tap, the
// more accurate direct sum); its distance to cv2 is pinned statistically (+-1 LSB of the u8 plane).
// Skipping exact-zero coefficients, as cv2 does, cannot change a value (x + 0*px == x).
static void filter2d_f32(const Plane8& src, const std::vector<float>& kern, int kw, PlaneF& dst) {
  const int H = src.rows, W = src.cols, r = kw / 2;
  dst.rows = H; dst.cols = W; dst.d.resize((size_t)H * W);
  // padded float copy so the inner loop is branch-free (thread-local scratch: no malloc churn)
  const int PW = W + 2 * r, PH = H + 2 * r;
  static thread_local std::vector<float> pad;
  static thread_local std::vector<float> acc;
  pad.resize((size_t)PW * PH);
  for (int y = 0; y < PH; y++) {
    int sy = border101(y - r, H);
    for (int x = 0; x < PW; x++) pad[(size_t)y * PW + x] = (float)src.at(sy, border101(x - r, W));
  }
  acc.resize((size_t)W);
  std::vector<int> tap_off; std::vector<float> tap_k;
  for (int j = 0; j < kw; j++)
    for (int i = 0; i < kw; i++) {
      float k = kern[(size_t)j * kw + i];
      if (k == 0.f) continue;
      tap_off.push_back(j * PW + i); tap_k.push_back(k);
    }
  const size_t nt = tap_k.size();
  // Row-at-a-time so the compiler can vectorise ACROSS pixels; every pixel still sees its taps in
  // raster order with separately rounded product and sum (build uses -ffp-contract=off).
  for (int y = 0; y < H; y++) {
    std::fill(acc.begin(), acc.end(), 0.f);
    float* __restrict a = acc.data();
    for (size_t t = 0; t < nt; t++) {
      const float* __restrict row = &pad[(size_t)y * PW + tap_off[t]];
      const float k = tap_k[t];
      for (int x = 0; x < W; x++) a[x] = a[x] + row[x] * k;   // == cv2's filter2D bit for bit for the 7x7 kernels
    }
    std::memcpy(&dst.d[(size_t)y * W], a, sizeof(float) * (size_t)W);
  }
}

// Complex Gabor response for width >= 9 in separable form.  cv2 (2.4.9 and 4.13 alike) evaluates these sizes through a DFT,
// so no direct sum is "the" reference arithmetic; the kernel is exactly (complex Gaussian-windowed exponential in x) x (the same
// in y) minus dc x (separable Gaussian), which costs 6K instead of K^2 multiply-adds per pixel.  Canonical order, all f32 fmaf:
//   row pass over the REFLECT_101-padded rows:  Rre = sum_i P(r, x+i) hx_re[i],  Rim likewise,  Gr = sum_i P(r, x+i) g1[i]
//   column pass, j ascending:  re = fmaf(Rre, hy_re[j], re); re = fmaf(-Rim, hy_im[j], re);
//                              im = fmaf(Rre, hy_im[j], im); im = fmaf( Rim, hy_re[j], im);  G = fmaf(Gr, g1[j], G)
//   re = fmaf(-dc, G, re)
// Its distance to cv2 is pinned statistically like every other form (+-1 LSB of the u8 plane, rate < 1e-3).
static void gabor_response_separable(const Plane8& src, const GaborKernel& gk, PlaneF& re_out, PlaneF& im_out) {
  const int H = src.rows, W = src.cols, K = gk.width, r = K / 2;
  const int PW = W + 2 * r, PH = H + 2 * r;
  static thread_local std::vector<float> pad, rre, rim, gr;
  pad.resize((size_t)PW * PH); rre.resize((size_t)PH * W); rim.resize((size_t)PH * W); gr.resize((size_t)PH * W);
  for (int y = 0; y < PH; y++) {
    const int sy = border101(y - r, H);
    for (int x = 0; x < PW; x++) pad[(size_t)y * PW + x] = (float)src.at(sy, border101(x - r, W));
  }
  for (int y = 0; y < PH; y++) {
    float* __restrict a = &rre[(size_t)y * W]; float* __restrict b = &rim[(size_t)y * W]; float* __restrict g = &gr[(size_t)y * W];
    for (int x = 0; x < W; x++) { a[x] = 0.f; b[x] = 0.f; g[x] = 0.f; }
    for (int i = 0; i < K; i++) {
      const float* __restrict p = &pad[(size_t)y * PW + i];
      const float cr = gk.hx_re[i], ci = gk.hx_im[i], cg = gk.g1[i];
      for (int x = 0; x < W; x++) { a[x] = std::fmaf(p[x], cr, a[x]); b[x] = std::fmaf(p[x], ci, b[x]); g[x] = std::fmaf(p[x], cg, g[x]); }
    }
  }
  re_out.rows = im_out.rows = H; re_out.cols = im_out.cols = W;
  re_out.d.resize((size_t)H * W); im_out.d.resize((size_t)H * W);
  std::vector<float> G((size_t)W);
  for (int y = 0; y < H; y++) {
    float* __restrict re = &re_out.d[(size_t)y * W]; float* __restrict im = &im_out.d[(size_t)y * W];
    for (int x = 0; x < W; x++) { re[x] = 0.f; im[x] = 0.f; G[x] = 0.f; }
    for (int j = 0; j < K; j++) {
      const float* __restrict a = &rre[(size_t)(y + j) * W]; const float* __restrict b = &rim[(size_t)(y + j) * W]; const float* __restrict g = &gr[(size_t)(y + j) * W];
      const float cr = gk.hy_re[j], ci = gk.hy_im[j], cg = gk.g1[j];
      for (int x = 0; x < W; x++) {
        re[x] = std::fmaf(a[x], cr, re[x]); re[x] = std::fmaf(-b[x], ci, re[x]);
        im[x] = std::fmaf(a[x], ci, im[x]); im[x] = std::fmaf(b[x], cr, im[x]);
        G[x] = std::fmaf(g[x], cg, G[x]);
      }
    }
    for (int x = 0; x < W; x++) re[x] = std::fmaf(-gk.dc, G[x], re[x]);
  }
}

// FeatureChannelFactory.hpp:253-286  gaborTransform (everything after filter2D is exactly reproducible)
static void gabor_transform(const Plane8& src, const GaborKernel& gk, Plane8& out8) {
  static thread_local PlaneF r_mat, i_mat;
  static thread_local std::vector<float> mag;
  if (gk.width >= 9) {
    gabor_response_separable(src, gk, r_mat, i_mat);
  } else {
    filter2d_f32(src, gk.re, gk.width, r_mat);
    filter2d_f32(src, gk.im, gk.width, i_mat);
  }
  const size_t n = r_mat.d.size();
  mag.resize(n);
  double smin = 0, smax = 0;
  for (size_t p = 0; p < n; p++) {
    float rr = r_mat.d[p] * r_mat.d[p];            // cv::pow(x,2)
    float ii = i_mat.d[p] * i_mat.d[p];
    float s = ii + rr;                             // cv::add(i_mat, r_mat); no contraction (-ffp-contract=off)
    float m = std::sqrt(s);                        // cv::pow(x,0.5) == sqrt
    mag[p] = m;
    if (p == 0) { smin = smax = m; } else { smin = std::min(smin, (double)m); smax = std::max(smax, (double)m); }
  }
  // cv::normalize(NORM_MINMAX, 0..1): scale/shift in double, applied as single-rounded FMA in f32
  double dscale = (1.0 - 0.0) * ((smax - smin) > 2.220446049250313e-16 ? 1. / (smax - smin) : 0.);
  double dshift = 0.0 - smin * dscale;
  float a = (float)dscale, b = (float)dshift;
  out8.rows = src.rows; out8.cols = src.cols; out8.d.assign(n, 0);
  for (size_t p = 0; p < n; p++) {
    float v = std::fmaf(mag[p], a, b);
    float q = v * 255.f;                             // convertTo(CV_8UC1, 255)
    long iv = std::lrintf(q);                        // cvRound: half to even
    out8.d[p] = (uint8_t)std::min<long>(std::max<long>(iv, 0), 255);
  }
}

// ----------------------------------------------------------------------------
// Thread pool with the reference's shape (include/ThreadPool.hpp:24-64):
// N workers created per pool, submit = post a task, join_all = drain + join.
// ----------------------------------------------------------------------------
class ThreadPool {
 public:
  explicit ThreadPool(int n) {
    if (n < 1) n = 1;
    for (int i = 0; i < n; i++) workers_.emplace_back([this] { run(); });
  }
  void submit(std::function<void()> f) {
    { std::lock_guard<std::mutex> lk(m_); q_.push_back(std::move(f)); }
    cv_.notify_one();
  }
  void join_all() {
    { std::lock_guard<std::mutex> lk(m_); done_ = true; }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
    workers_.clear();
  }
  ~ThreadPool() { if (!workers_.empty()) join_all(); }
 private:
  void run() {
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [this] { return done_ || !q_.empty(); });
        if (q_.empty()) return;
        f = std::move(q_.front());
        q_.pop_front();
      }
      f();
    }
  }
  std::vector<std::thread> workers_;
  std::deque<std::function<void()>> q_;
  std::mutex m_;
  std::condition_variable cv_;
  bool done_ = false;
};

// ----------------------------------------------------------------------------
// ImageSample (include/ImageSample.hpp:146-199, src/ImageSample.cpp:11-90)
// ----------------------------------------------------------------------------
struct ImageSample {
  std::vector<PlaneF> m_feature_channels;   // integral planes (use_integral = true)
  std::vector<Plane8> planes8;              // the 8-bit planes before cv::integral (parity output)
  // threads <= 1: serial; > 1: one pool task per Gabor filter (FeatureChannelFactory.hpp:79-87)
  ImageSample(const Plane8& img, std::vector<int> features, int threads) {
    std::sort(features.begin(), features.end());          // ImageSample.cpp:86
    for (int f : features) extract(f, img, threads);
    m_feature_channels.resize(planes8.size());
    for (size_t i = 0; i < planes8.size(); i++) integral_f32(planes8[i], m_feature_channels[i]);
  }
  void extract(int feature, const Plane8& img, int threads) {
    switch (feature) {
      case 0: planes8.push_back(img); break;                                   // FC_GRAY
      case 1: {                                                                // FC_GABOR
        const auto& bank = gabor_bank();
        size_t old = planes8.size();
        planes8.resize(old + bank.size());
        if (threads > 1) {
          ThreadPool e(threads);
          for (size_t i = 0; i < bank.size(); i++)
            e.submit([&, i] { gabor_transform(img, bank[i], planes8[old + i]); });
          e.join_all();
        } else {
          for (size_t i = 0; i < bank.size(); i++) gabor_transform(img, bank[i], planes8[old + i]);
        }
        break;
      }
      case 2: {                                                                // FC_SOBEL
        Plane8 a, b;
        sobel_u8(img, a, 0, 1);   // "sob_x" = d/dy (FeatureChannelFactory.hpp:124)
        sobel_u8(img, b, 1, 0);   // "sob_y" = d/dx (:125)
        planes8.push_back(a); planes8.push_back(b);
        break;
      }
      case 3: {                                                                // FC_MIN_MAX
        Plane8 a, b;
        minmax3x3_u8(img, a, b);
        planes8.push_back(a); planes8.push_back(b);
        break;
      }
      case 5: {                                                                // FC_NORM
        Plane8 a;
        equalize_hist_u8(img, a);
        planes8.push_back(a);
        break;
      }
      case 4: {                                                                // FC_CANNY
        Plane8 a;
        canny_u8(img, a);
        planes8.push_back(a);
        break;
      }
      default: break;
    }
  }
  // src/ImageSample.cpp:30-64, integral branch (A.6)
  int evalTest(const Split& test, const Rect& rect, bool* in_bounds = nullptr) const {
    const PlaneF& img = m_feature_channels[test.feature_channel];
    (void)in_bounds;
    int R1_a = (int)img.at(rect.y + test.rect1.y, rect.x + test.rect1.x);
    int R1_b = (int)img.at(rect.y + test.rect1.y, rect.x + test.rect1.x + test.rect1.width);
    int R1_c = (int)img.at(rect.y + test.rect1.y + test.rect1.height, rect.x + test.rect1.x);
    int R1_d = (int)img.at(rect.y + test.rect1.y + test.rect1.height, rect.x + test.rect1.x + test.rect1.width);
    int p1 = (int)((R1_d - R1_b - R1_c + R1_a) / static_cast<float>(test.rect1.width * test.rect1.height));
    int R2_a = (int)img.at(rect.y + test.rect2.y, rect.x + test.rect2.x);
    int R2_b = (int)img.at(rect.y + test.rect2.y, rect.x + test.rect2.x + test.rect2.width);
    int R2_c = (int)img.at(rect.y + test.rect2.y + test.rect2.height, rect.x + test.rect2.x);
    int R2_d = (int)img.at(rect.y + test.rect2.y + test.rect2.height, rect.x + test.rect2.x + test.rect2.width);
    int p2 = (int)((R2_d - R2_b - R2_c + R2_a) / static_cast<float>(test.rect2.width * test.rect2.height));
    return p1 - p2;
  }
};

struct PatchSample { const ImageSample* m_image; Rect m_patch_bbox; };

static std::atomic<long long> g_visits{0};   // node visits (for ALG_BYTES bookkeeping)
static bool g_count_visits = false;

// include/Tree.hpp:174-191 — recursive, pointer chasing, as in the reference
static void tree_evaluateMT(const PatchSample* s, TreeNode* node, Leaf** leaf, long long* visits) {
  if (visits) ++*visits;
  if (node->is_leaf) *leaf = &node->leaf;
  else {
    // *Sample::eval (HeadPoseSample.cpp:39-46, MPSample.cpp:77-84)
    bool go_left = s->m_image->evalTest(node->split, s->m_patch_bbox) <= node->split.threshold;
    if (go_left) tree_evaluateMT(s, node->left, leaf, visits);
    else tree_evaluateMT(s, node->right, leaf, visits);
  }
}
// include/Forest.hpp:81-90.  The Forest is taken BY VALUE: boost::bind copies it per task in the
// reference (face_utils.cpp:215, :273); the copy cost is part of the reference's CPU path.
static void forest_evaluateMT(Forest forest, const PatchSample* s, Leaf** leafs) {
  long long v = 0;
  for (int i = 0; i < forest.numberOfTrees(); i++, leafs++)
    tree_evaluateMT(s, forest.m_trees[i]->root, leafs, g_count_visits ? &v : nullptr);
  if (g_count_visits) g_visits += v;
}

struct HeadPoseEstimatorOption { int num_head_pose_labels = 5; int step_size = 4; float min_foreground_probability = 0.5f; };
struct MultiPartEstimatorOption { int num_parts = 10; int step_size = 3; int min_samples = 2; float min_forground = 0.5f; float min_pf = 0.25f; float max_variance = 25.f; };
struct MeanShiftOption { int kernel_size = 10; int max_iterations = 7; float stopping_criteria = 0.05f; };
struct Vote { Point pos; float weight = 0.f; bool check = false; };

template <class F>
static void run_patches(const std::vector<PatchSample>& samples, int threads, F&& per_sample) {
  if (threads > 1) {
    ThreadPool e(threads);
    for (size_t i = 0; i < samples.size(); i++) e.submit([&, i] { per_sample(i); });
    e.join_all();
  } else {
    for (size_t i = 0; i < samples.size(); i++) per_sample(i);
  }
}

static void make_grid(const Rect& face_bbox, int patch_size, int step, const ImageSample* img, std::vector<PatchSample>& samples) {
  // face_utils.cpp:198-207 / :256-265 — x outer, y inner (A.7)
  for (int x = face_bbox.x; x < face_bbox.x + face_bbox.width - patch_size; x += step)
    for (int y = face_bbox.y; y < face_bbox.y + face_bbox.height - patch_size; y += step) {
      PatchSample s; s.m_image = img; s.m_patch_bbox = Rect{x, y, patch_size, patch_size};
      samples.push_back(s);
    }
}

// src/face_utils.cpp:183-242
static void getHeadPoseVotesMT(const ImageSample& sample, const Forest& forest, Rect face_bbox, float* headpose,
                               float* variance, HeadPoseEstimatorOption options, int threads,
                               std::vector<Leaf*>* leafs_out) {
  int patch_size = forest.m_forest_param.getPatchSize();
  int num_trees = forest.numberOfTrees();
  std::vector<PatchSample> samples;
  make_grid(face_bbox, patch_size, options.step_size, &sample, samples);
  std::vector<Leaf*> leafs(samples.size() * (size_t)num_trees);
  run_patches(samples, threads, [&](size_t i) { forest_evaluateMT(forest, &samples[i], &leafs[i * num_trees]); });
  float n = 0, sum = 0, sum_sq = 0;
  for (size_t i = 0; i < leafs.size(); ++i) {
    if (leafs[i]->hp_foreground > options.min_foreground_probability) {
      float m = 0;
      for (int j = 0; j < options.num_head_pose_labels; j++) m += leafs[i]->hp_labels[j] * j;
      m /= (leafs[i]->hp_nsamples * leafs[i]->hp_foreground);
      sum += m;
      sum_sq += m * m;
      n++;
    }
  }
  float mean = sum / n;
  float var = (sum_sq / n) - (mean * mean);
  mean -= 2;
  var *= 0.05f;  // NORM_HEADPOSE_VARIANCE_FACTOR (Constants.hpp:67)
  *headpose = mean;
  *variance = var;
  if (leafs_out) *leafs_out = leafs;
}

// src/face_utils.cpp:244-302
static void getFacialFeaturesVotesMT(const ImageSample& sample, const Forest& forest, Rect face_bbox,
                                     std::vector<std::vector<Vote>>& votes, MultiPartEstimatorOption options,
                                     int threads, std::vector<Leaf*>* leafs_out) {
  int patch_size = forest.m_forest_param.getPatchSize();
  std::vector<PatchSample> samples;
  make_grid(face_bbox, patch_size, options.step_size, &sample, samples);
  int num_trees = forest.numberOfTrees();
  std::vector<Leaf*> leafs(samples.size() * (size_t)num_trees);
  run_patches(samples, threads, [&](size_t i) { forest_evaluateMT(forest, &samples[i], &leafs[i * num_trees]); });
  int i_sample = 0;
  for (size_t k = 0; k < leafs.size(); k++) {
    Leaf* L = leafs[k];
    int offset_x = samples[i_sample / num_trees].m_patch_bbox.x + patch_size / 2;
    int offset_y = samples[i_sample / num_trees].m_patch_bbox.y + patch_size / 2;
    for (unsigned int i = 0; i < votes.size(); i++) {
      float min_pf = options.min_pf;
      if (i == 0 || i == 7) min_pf *= 1.5;
      if (L->mp_foreground > options.min_forground && L->mp_prob_foreground[i] > min_pf &&
          L->mp_parts_variance[i] < options.max_variance && L->mp_samples > options.min_samples) {
        Vote v;
        v.pos.x = L->mp_parts_offset[i].x + offset_x;
        v.pos.y = L->mp_parts_offset[i].y + offset_y;
        v.weight = L->mp_foreground;
        v.check = true;
        votes[i].push_back(v);
      }
    }
    i_sample++;
  }
  if (leafs_out) *leafs_out = leafs;
}

// src/face_utils.cpp:304-323
static float areaUnderCurve(float x1, float x2, double mean, double std_) {
  double sum = 0;
  double step = 0.01;
  double t;
  for (double x = x1; x < x2; x += step) {
    t = (x - mean) / std_;
    sum += std::exp(-0.5 * (t * t)) * step;
  }
  return (float)(sum * 1.0 / (std_ * std::sqrt(2 * M_PI)));
}

// include/MeanShift.hpp:52-135
static inline long cv_round(float v) { return std::lrintf(v); }
static void meanshift_shift(const std::vector<Vote>& votes, Point& result, float* mean_f, int* iters,
                            int num_iterations, int kernel, float stopping_criteria) {
  bool coverg = false;
  float mx = 0.f, my = 0.f;
  {
    float sum_w = 0;
    for (size_t i = 0; i < votes.size(); i++) {
      if (!votes[i].check) continue;
      float w = votes[i].weight;
      mx += votes[i].pos.x * w;
      my += votes[i].pos.y * w;
      sum_w += w;
    }
    if (sum_w > 0) { mx /= sum_w; my /= sum_w; }
  }
  int it = 0;
  float lamda = (float)kernel;
  for (int i = 0; (i < num_iterations) && (coverg == false); i++) {
    float sx = 0.f, sy = 0.f, sum_w = 0;
    for (size_t k = 0; k < votes.size(); k++) {
      if (!votes[k].check) continue;
      float dx = mx - (float)votes[k].pos.x, dy = my - (float)votes[k].pos.y;
      float d = (float)std::sqrt((double)dx * dx + (double)dy * dy);   // cv::norm(Point2f)
      d = expf(-d / lamda);
      float w = votes[k].weight * d;
      sx += votes[k].pos.x * w;
      sy += votes[k].pos.y * w;
      sum_w += w;
    }
    if (sum_w > 0) { sx /= sum_w; sy /= sum_w; }
    float ex = sx - mx, ey = sy - my;
    if (std::sqrt((double)ex * ex + (double)ey * ey) < stopping_criteria) coverg = true;
    mx = sx; my = sy;
    it++;
  }
  result.x = (int)cv_round(mx);   // Point_<int> = Point_<float>: saturate_cast -> cvRound
  result.y = (int)cv_round(my);
  if (mean_f) { mean_f[0] = mx; mean_f[1] = my; }
  if (iters) *iters = it;
}

// ----------------------------------------------------------------------------
// Model = what FaceForest's constructor loads (src/FaceForest.cpp:15-58)
// ----------------------------------------------------------------------------
struct Model {
  Forest hp_forest;
  std::vector<Forest> mp_jungle;
  ForestParam hp_param, mp_param;
  std::vector<std::string> mp_forest_paths;
};

static bool is_dir(const std::string& p) { struct stat st; return ::stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }

static ForestParam default_param(int ntrees, int max_depth) {
  // values of data/config_headpose.txt / data/config_ffd.txt that inference reads
  ForestParam p;
  p.ntrees = ntrees; p.max_depth = max_depth; p.face_size = 125; p.patch_size_ratio = 0.25f;
  p.features = {0, 1, 2};
  return p;
}

}  // namespace

// =============================================================================
// C interface (ctypes)
// =============================================================================
extern "C" {

typedef struct {
  float headpose, variance;
  int tree_counts[5];
  int dominant;
  int scaled_w, scaled_h;
  float scale;
  float ffd_f[10][2];
  int ffd_scaled[10][2];
  int ffd[10][2];
  int ms_iters[10];
  int n_votes[10];
  int flags;
} orc_face_t;

typedef struct {
  int hp_stride, ffd_stride;
  int threads;          // <=1 serial, >1 ThreadPool with that many workers (reference: hardware_concurrency)
  int features_mask;    // bit f set => feature id f enabled (default 0b111 = gray, gabor, sobel)
  int headpose_only;
} orc_options_t;

const char* orc_last_error() { return g_err.c_str(); }

void* orc_model_load(const char* hp_dir, int hp_ntrees, const char* ffd_dir, int ffd_ntrees) {
  Model* m = new Model();
  m->hp_param = default_param(hp_ntrees, 15);
  m->mp_param = default_param(ffd_ntrees, 20);
  if (hp_dir && *hp_dir) {
    m->hp_param.tree_path = hp_dir;
    if (!m->hp_forest.load(hp_dir, m->hp_param, KIND_HP)) { delete m; return nullptr; }
    if (m->hp_forest.numberOfTrees() > 0) {
      const ForestParam& sp = m->hp_forest.m_trees[0]->m_param;
      m->hp_param.face_size = sp.face_size; m->hp_param.patch_size_ratio = sp.patch_size_ratio;
      m->hp_forest.m_forest_param = m->hp_param;
    }
  }
  if (ffd_dir && *ffd_dir) {
    m->mp_param.tree_path = ffd_dir;
    DIR* d = ::opendir(ffd_dir);
    if (!d) { set_err(std::string("cannot open ") + ffd_dir); delete m; return nullptr; }
    while (dirent* e = ::readdir(d)) {
      std::string name = e->d_name;
      if (name == "." || name == "..") continue;
      std::string p = std::string(ffd_dir) + "/" + name;
      if (is_dir(p)) m->mp_forest_paths.push_back(p);
    }
    ::closedir(d);
    std::sort(m->mp_forest_paths.begin(), m->mp_forest_paths.end());   // FaceForest.cpp:44
    for (auto& p : m->mp_forest_paths) {
      Forest f;
      if (!f.load(p, m->mp_param, KIND_MP)) { delete m; return nullptr; }
      m->mp_jungle.push_back(f);
    }
    if (!m->mp_jungle.empty() && m->mp_jungle[0].numberOfTrees() > 0) {
      const ForestParam& sp = m->mp_jungle[0].m_trees[0]->m_param;
      m->mp_param.face_size = sp.face_size; m->mp_param.patch_size_ratio = sp.patch_size_ratio;
      for (auto& f : m->mp_jungle) f.m_forest_param = m->mp_param;
    }
  }
  return m;
}

void orc_model_free(void* h) {
  Model* m = (Model*)h;
  if (!m) return;
  for (Tree* t : m->hp_forest.m_trees) delete t;
  for (auto& f : m->mp_jungle) for (Tree* t : f.m_trees) delete t;
  delete m;
}

// Reads the packed binary forest image written by the product's crf_model_save_packed
// (face_alignment_cvpr_2012_b200/csrc/model.cc: magic "CRFB200M", version 2, FNV-1a 64 trailer) into
// the oracle's own pointer-linked trees.  The text archives (282 MB) cannot travel to the GPU box;
// tests/test_model_loader.py proves here, where /root/reference exists, that this image holds exactly
// what the oracle's own Boost-archive parser reads from the shipped files.
}  // extern "C"
namespace {
#pragma pack(push, 1)
struct PkNode { int32_t right; int32_t thr_raw; uint8_t ch, depth; uint8_t r1[4], r2[4]; uint8_t is_leaf; uint8_t pad; };
struct PkParam { int32_t max_depth, min_patches, ntests, ntrees, nimages, npatches, face_size; float patch_size_ratio; int32_t n_features; int32_t features[8]; };
struct PkHpLeaf { int32_t nsamples; float foreground; int32_t labels[5]; int32_t object_id; };
struct PkMpLeaf { int32_t samples; int16_t off[10][2]; float var[10]; float pf[10]; float fg; int32_t oid; };
#pragma pack(pop)
struct PkReader {
  const uint8_t* p; const uint8_t* e; bool ok = true;
  template <class T> T pod() { T v{}; if (p + sizeof(T) > e) { ok = false; return v; } std::memcpy(&v, p, sizeof(T)); p += sizeof(T); return v; }
};
static bool read_packed_forest(PkReader& r, Forest& f, Kind kind) {
  r.pod<int32_t>();  // kind
  int32_t nt = r.pod<int32_t>();
  if (!r.ok || nt < 0 || nt > 4096) return false;
  for (int t = 0; t < nt; t++) {
    Tree* tree = new Tree();
    f.m_trees.push_back(tree);
    tree->m_num_nodes = r.pod<int32_t>(); tree->i_node = r.pod<int32_t>(); tree->max_depth_seen = r.pod<int32_t>();
    PkParam pp = r.pod<PkParam>();
    tree->m_param.max_depth = pp.max_depth; tree->m_param.ntrees = pp.ntrees; tree->m_param.face_size = pp.face_size;
    tree->m_param.patch_size_ratio = pp.patch_size_ratio;
    for (int i = 0; i < pp.n_features && i < 8; i++) tree->m_param.features.push_back(pp.features[i]);
    int32_t nn = r.pod<int32_t>();
    if (!r.ok || nn < 1) return false;
    std::vector<PkNode> recs((size_t)nn);
    for (auto& q : recs) q = r.pod<PkNode>();
    std::vector<TreeNode*> nodes((size_t)nn);
    std::vector<TreeNode*> leaves;
    for (int i = 0; i < nn; i++) {
      TreeNode* n = new TreeNode();
      nodes[i] = n;
      n->object_id = i; n->depth = recs[i].depth; n->is_leaf = recs[i].is_leaf != 0; n->has_split = !n->is_leaf;
      n->leaf.object_id = i;
      if (n->is_leaf) leaves.push_back(n);
      else {
        Split& s = n->split;
        s.feature_channel = recs[i].ch; s.threshold = recs[i].thr_raw;
        s.rect1 = Rect{recs[i].r1[0], recs[i].r1[1], recs[i].r1[2], recs[i].r1[3]};
        s.rect2 = Rect{recs[i].r2[0], recs[i].r2[1], recs[i].r2[2], recs[i].r2[3]};
      }
    }
    tree->root = nodes[0];
    tree->n_nodes_parsed = nn; tree->n_leaves_parsed = (int)leaves.size();
    for (int i = 0; i < nn; i++)
      if (!nodes[i]->is_leaf) {
        if (recs[i].right <= i + 1 || recs[i].right >= nn) return false;
        nodes[i]->left = nodes[i + 1]; nodes[i]->right = nodes[recs[i].right];
      }
    int32_t nh = r.pod<int32_t>();
    if (!r.ok || nh < 0 || nh > nn) return false;
    for (int i = 0; i < nh; i++) {
      PkHpLeaf q = r.pod<PkHpLeaf>();
      if (kind != KIND_HP || i >= (int)leaves.size()) return false;
      Leaf& L = leaves[i]->leaf;
      L.hp_nsamples = q.nsamples; L.hp_foreground = q.foreground; L.hp_nlabels = 5;
      for (int j = 0; j < 5; j++) L.hp_labels[j] = q.labels[j];
    }
    int32_t nm = r.pod<int32_t>();
    if (!r.ok || nm < 0 || nm > nn) return false;
    for (int i = 0; i < nm; i++) {
      PkMpLeaf q = r.pod<PkMpLeaf>();
      if (kind != KIND_MP || i >= (int)leaves.size()) return false;
      Leaf& L = leaves[i]->leaf;
      L.mp_samples = q.samples; L.mp_nparts = 10; L.mp_foreground = q.fg;
      for (int j = 0; j < 10; j++) { L.mp_parts_offset[j].x = q.off[j][0]; L.mp_parts_offset[j].y = q.off[j][1]; L.mp_parts_variance[j] = q.var[j]; L.mp_prob_foreground[j] = q.pf[j]; }
    }
    if (!r.ok || (int)leaves.size() != nh + nm) return false;
  }
  return r.ok;
}
}  // namespace
extern "C" {

void* orc_model_load_packed(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { set_err(std::string("File not found: ") + path); return nullptr; }
  std::fseek(f, 0, SEEK_END);
  long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf((size_t)std::max(sz, 0L));
  bool ok = sz > 0 && std::fread(buf.data(), 1, buf.size(), f) == buf.size();
  std::fclose(f);
  if (!ok || buf.size() < 20 || std::memcmp(buf.data(), "CRFB200M", 8) != 0) { set_err("not a packed CRF model"); return nullptr; }
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i + 8 < buf.size(); i++) { h ^= buf[i]; h *= 1099511628211ull; }
  uint64_t want; std::memcpy(&want, buf.data() + buf.size() - 8, 8);
  if (h != want) { set_err("checksum mismatch"); return nullptr; }
  PkReader r{buf.data() + 8, buf.data() + buf.size() - 8};
  if (r.pod<uint32_t>() != 2) { set_err("packed model version mismatch"); return nullptr; }
  Model* m = new Model();
  int32_t hp_cfg = r.pod<int32_t>(), mp_cfg = r.pod<int32_t>(), face_size = r.pod<int32_t>();
  r.pod<int32_t>(); r.pod<int32_t>();  // patch size, channel count (derived)
  m->hp_param = default_param(hp_cfg, 15); m->mp_param = default_param(mp_cfg, 20);
  m->hp_param.face_size = m->mp_param.face_size = face_size;
  bool good = read_packed_forest(r, m->hp_forest, KIND_HP);
  int32_t nj = good ? r.pod<int32_t>() : 0;
  for (int i = 0; good && i < nj; i++) { m->mp_jungle.emplace_back(); good = read_packed_forest(r, m->mp_jungle.back(), KIND_MP); }
  if (!good || !r.ok) { set_err("corrupt packed model"); orc_model_free(m); return nullptr; }
  if (m->hp_forest.numberOfTrees() > 0) m->hp_param.patch_size_ratio = m->hp_forest.m_trees[0]->m_param.patch_size_ratio;
  m->mp_param.patch_size_ratio = m->hp_param.patch_size_ratio;
  m->hp_forest.m_forest_param = m->hp_param;
  for (auto& jf : m->mp_jungle) jf.m_forest_param = m->mp_param;
  return m;
}

// Leaf payload dump for cross-checking loaders: per leaf in pre-order, 36 floats:
// HP: [oid, nsamples, fg, labels x5, 0...]; MP: [oid, samples, fg, off x20, var x10, pf x10 -> 43] (cap 44)
int orc_leaf_dump(void* h, int which, int tree, float* out, int cap_leaves) {
  Model* m = (Model*)h;
  const Forest& f = which < 0 ? m->hp_forest : m->mp_jungle[which];
  if (tree < 0 || tree >= f.numberOfTrees()) return -1;
  std::vector<TreeNode*> st{f.m_trees[tree]->root};
  int n = 0;
  while (!st.empty()) {
    TreeNode* nd = st.back(); st.pop_back();
    if (nd->is_leaf) {
      if (n < cap_leaves && out) {
        float* o = out + (size_t)n * 44;
        std::memset(o, 0, sizeof(float) * 44);
        const Leaf& L = nd->leaf;
        o[0] = (float)nd->object_id;
        if (which < 0) { o[1] = (float)L.hp_nsamples; o[2] = L.hp_foreground; for (int j = 0; j < 5; j++) o[3 + j] = (float)L.hp_labels[j]; }
        else {
          o[1] = (float)L.mp_samples; o[2] = L.mp_foreground;
          for (int j = 0; j < 10; j++) { o[3 + 2 * j] = (float)L.mp_parts_offset[j].x; o[4 + 2 * j] = (float)L.mp_parts_offset[j].y; o[23 + j] = L.mp_parts_variance[j]; o[33 + j] = L.mp_prob_foreground[j]; }
        }
      }
      n++;
    } else { st.push_back(nd->right); st.push_back(nd->left); }
  }
  return n;
}


// counts: [0]=hp trees [1]=hp nodes [2]=hp leaves [3]=n mp forests [4]=mp trees [5]=mp nodes [6]=mp leaves
// [7]=hp max depth [8]=mp max depth [9]=patch size [10]=face size
int orc_model_info(void* h, int* counts) {
  Model* m = (Model*)h;
  std::memset(counts, 0, sizeof(int) * 11);
  counts[0] = m->hp_forest.numberOfTrees();
  for (Tree* t : m->hp_forest.m_trees) { counts[1] += t->n_nodes_parsed; counts[2] += t->n_leaves_parsed; counts[7] = std::max(counts[7], t->max_depth_seen); }
  counts[3] = (int)m->mp_jungle.size();
  for (auto& f : m->mp_jungle)
    for (Tree* t : f.m_trees) { counts[4]++; counts[5] += t->n_nodes_parsed; counts[6] += t->n_leaves_parsed; counts[8] = std::max(counts[8], t->max_depth_seen); }
  counts[9] = m->hp_param.getPatchSize();
  counts[10] = m->hp_param.face_size;
  return 0;
}

// Flatten one tree in pre-order (== Boost object id order) for cross-checking the product's packer.
// which: -1 = head-pose forest, 0..4 = jungle forest.  Each node -> 16 ints:
// [is_leaf, depth, ch, r1x,r1y,r1w,r1h, r2x,r2y,r2w,r2h, thr, left_oid, right_oid, nsamples, object_id]
int orc_tree_dump(void* h, int which, int tree, int* out, int cap_nodes) {
  Model* m = (Model*)h;
  const Forest& f = which < 0 ? m->hp_forest : m->mp_jungle[which];
  if (tree < 0 || tree >= f.numberOfTrees()) return -1;
  std::vector<TreeNode*> st{f.m_trees[tree]->root};
  int n = 0;
  while (!st.empty()) {
    TreeNode* nd = st.back(); st.pop_back();
    if (n < cap_nodes) {
      int* o = out + (size_t)n * 16;
      o[0] = nd->is_leaf; o[1] = nd->depth; o[2] = nd->split.feature_channel;
      o[3] = nd->split.rect1.x; o[4] = nd->split.rect1.y; o[5] = nd->split.rect1.width; o[6] = nd->split.rect1.height;
      o[7] = nd->split.rect2.x; o[8] = nd->split.rect2.y; o[9] = nd->split.rect2.width; o[10] = nd->split.rect2.height;
      o[11] = nd->split.threshold;
      o[12] = nd->left ? nd->left->object_id : -1; o[13] = nd->right ? nd->right->object_id : -1;
      o[14] = which < 0 ? nd->leaf.hp_nsamples : nd->leaf.mp_samples;
      o[15] = nd->object_id;
    }
    n++;
    if (!nd->is_leaf) { st.push_back(nd->right); st.push_back(nd->left); }
  }
  return n;
}

void orc_bgr2gray(const uint8_t* bgr, int rows, int cols, size_t step, uint8_t* gray) {
  for (int y = 0; y < rows; y++) {
    const uint8_t* p = bgr + (size_t)y * step;
    for (int x = 0; x < cols; x++) gray[(size_t)y * cols + x] = bgr2gray_px(p[3 * x], p[3 * x + 1], p[3 * x + 2]);
  }
}

// Scaled size as src/FaceForest.cpp:202-204 computes it.
void orc_scaled_size(int roi_cols, int roi_rows, int face_size, int* w, int* h, float* scale_out) {
  float scale = static_cast<float>(face_size) / static_cast<float>(roi_cols);
  *w = (int)(roi_cols * scale);
  *h = (int)(roi_rows * scale);
  if (scale_out) *scale_out = scale;
}

void orc_resize(const uint8_t* src, int sh, int sw, size_t sstep, uint8_t* dst, int dh, int dw) {
  Plane8 d;
  resize_linear_u8(src, sh, sw, sstep, d, dh, dw);
  std::memcpy(dst, d.d.data(), d.d.size());
}

int orc_num_planes(int features_mask) {
  int n = 0;
  if (features_mask & 1) n += 1;
  if (features_mask & 2) n += 35;
  if (features_mask & 4) n += 2;
  if (features_mask & 8) n += 2;
  if (features_mask & 16) n += 1;
  if (features_mask & 32) n += 1;
  return n;
}

static std::vector<int> features_from_mask(int mask) {
  std::vector<int> f;
  for (int i = 0; i < 6; i++) if (mask & (1 << i)) f.push_back(i);
  return f;
}

// planes8: C x H x W u8, integrals: C x (H+1) x (W+1) f32 (either may be NULL)
int orc_channels(const uint8_t* gray, int H, int W, int features_mask, int threads, uint8_t* planes8, float* integrals) {
  Plane8 img; img.rows = H; img.cols = W; img.d.assign(gray, gray + (size_t)H * W);
  ImageSample s(img, features_from_mask(features_mask), threads);
  for (size_t c = 0; c < s.planes8.size(); c++) {
    if (planes8) std::memcpy(planes8 + c * (size_t)H * W, s.planes8[c].d.data(), (size_t)H * W);
    if (integrals) std::memcpy(integrals + c * (size_t)(H + 1) * (W + 1), s.m_feature_channels[c].d.data(), sizeof(float) * (size_t)(H + 1) * (W + 1));
  }
  return (int)s.planes8.size();
}

// Gabor bank export: widths[35]; coefficient arrays re/im concatenated in bank order.
int orc_gabor_bank(int* widths, float* re, float* im, int cap) {
  const auto& bank = gabor_bank();
  int n = 0;
  for (size_t i = 0; i < bank.size(); i++) {
    widths[i] = bank[i].width;
    for (size_t k = 0; k < bank[i].re.size(); k++) {
      if (n < cap) { if (re) re[n] = bank[i].re[k]; if (im) im[n] = bank[i].im[k]; }
      n++;
    }
  }
  return n;
}

// raw f32 filter2D response of one Gabor kernel (for pinning against cv2.filter2D)
void orc_gabor_response(const uint8_t* gray, int H, int W, int index, float* re, float* im) {
  Plane8 img; img.rows = H; img.cols = W; img.d.assign(gray, gray + (size_t)H * W);
  const auto& gk = gabor_bank()[index];
  PlaneF r, i;
  filter2d_f32(img, gk.re, gk.width, r);
  filter2d_f32(img, gk.im, gk.width, i);
  std::memcpy(re, r.d.data(), sizeof(float) * r.d.size());
  std::memcpy(im, i.d.data(), sizeof(float) * i.d.size());
}

// ---- single-stage exports used by oracle/shim/cvshim.cc (the OpenCV stand-in of the real-reference build, _ref/):
// each is one of the cv2-pinned stages above on dense H x W planes.
void orc_cv_integral(const uint8_t* src, int H, int W, float* dst /* (H+1) x (W+1) */) {
  Plane8 s; s.rows = H; s.cols = W; s.d.assign(src, src + (size_t)H * W);
  PlaneF d; integral_f32(s, d);
  std::memcpy(dst, d.d.data(), sizeof(float) * d.d.size());
}
void orc_cv_sobel(const uint8_t* src, int H, int W, int dx, int dy, uint8_t* dst) {
  Plane8 s; s.rows = H; s.cols = W; s.d.assign(src, src + (size_t)H * W);
  Plane8 d; sobel_u8(s, d, dx, dy);
  std::memcpy(dst, d.d.data(), d.d.size());
}
void orc_cv_minmax(const uint8_t* src, int H, int W, uint8_t* mn, uint8_t* mx) {
  Plane8 s; s.rows = H; s.cols = W; s.d.assign(src, src + (size_t)H * W);
  Plane8 a, b; minmax3x3_u8(s, a, b);
  if (mn) std::memcpy(mn, a.d.data(), a.d.size());
  if (mx) std::memcpy(mx, b.d.data(), b.d.size());
}
void orc_cv_equalize(const uint8_t* src, int H, int W, uint8_t* dst) {
  Plane8 s; s.rows = H; s.cols = W; s.d.assign(src, src + (size_t)H * W);
  Plane8 d; equalize_hist_u8(s, d);
  std::memcpy(dst, d.d.data(), d.d.size());
}
void orc_cv_canny(const uint8_t* src, int H, int W, int low, int high, uint8_t* dst) {
  Plane8 s; s.rows = H; s.cols = W; s.d.assign(src, src + (size_t)H * W);
  Plane8 d; canny_u8(s, d, low, high);
  std::memcpy(dst, d.d.data(), d.d.size());
}
// cv::filter2D(u8 -> f32) with a caller-supplied kw x kw kernel.  mode 0: raster order, product and sum rounded separately
// (== cv2 for the sizes cv2 evaluates directly); mode 1: raster order accumulated in double, rounded once (the closest
// f32 to the exact response — the neutral stand-in for cv2's DFT path, which no direct sum matches bit for bit).
void orc_cv_filter2d(const uint8_t* src, int H, int W, const float* kern, int kw, int mode, float* dst) {
  Plane8 s; s.rows = H; s.cols = W; s.d.assign(src, src + (size_t)H * W);
  if (mode == 0) {
    PlaneF d; filter2d_f32(s, std::vector<float>(kern, kern + (size_t)kw * kw), kw, d);
    std::memcpy(dst, d.d.data(), sizeof(float) * d.d.size());
    return;
  }
  const int r = kw / 2;
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      double acc = 0;
      for (int j = 0; j < kw; j++) {
        const int sy = border101(y + j - r, H);
        for (int i = 0; i < kw; i++) acc += (double)s.at(sy, border101(x + i - r, W)) * (double)kern[(size_t)j * kw + i];
      }
      dst[(size_t)y * W + x] = (float)acc;
    }
}
// Canonical complex response of bank kernel `index` (separable form for widths >= 9, raster for 7): what gabor_transform uses.
void orc_cv_gabor_canonical(const uint8_t* src, int H, int W, int index, float* re, float* im) {
  Plane8 s; s.rows = H; s.cols = W; s.d.assign(src, src + (size_t)H * W);
  const auto& gk = gabor_bank()[index];
  PlaneF r, i;
  if (gk.width >= 9) gabor_response_separable(s, gk, r, i);
  else { filter2d_f32(s, gk.re, gk.width, r); filter2d_f32(s, gk.im, gk.width, i); }
  if (re) std::memcpy(re, r.d.data(), sizeof(float) * r.d.size());
  if (im) std::memcpy(im, i.d.data(), sizeof(float) * i.d.size());
}

int orc_num_patches(int W, int H, int patch, int step) {
  int nx = 0, ny = 0;
  for (int x = 0; x < W - patch; x += step) nx++;
  for (int y = 0; y < H - patch; y += step) ny++;
  return nx * ny;
}

struct Sample38 {
  ImageSample* s;
};

// Build an ImageSample from a scaled gray face (H x W).
void* orc_sample_create(const uint8_t* gray, int H, int W, int features_mask, int threads) {
  Plane8 img; img.rows = H; img.cols = W; img.d.assign(gray, gray + (size_t)H * W);
  return new ImageSample(img, features_from_mask(features_mask), threads);
}
void orc_sample_free(void* s) { delete (ImageSample*)s; }

// Build an ImageSample directly from caller-supplied u8 planes (C x H x W): lets tests drive the
// forest stages with synthetic channel data.
void* orc_sample_from_planes(const uint8_t* planes, int C, int H, int W) {
  Plane8 dummy; dummy.rows = H; dummy.cols = W; dummy.d.assign((size_t)H * W, 0);
  ImageSample* s = new ImageSample(dummy, {}, 1);
  s->planes8.resize(C);
  s->m_feature_channels.resize(C);
  for (int c = 0; c < C; c++) {
    s->planes8[c].rows = H; s->planes8[c].cols = W;
    s->planes8[c].d.assign(planes + (size_t)c * H * W, planes + (size_t)(c + 1) * H * W);
    integral_f32(s->planes8[c], s->m_feature_channels[c]);
  }
  return s;
}

static bool compose_forest(const Model* m, const int* forest_idx, const int* tree_idx, int n, Forest& out) {
  out.m_forest_param = m->mp_param;
  for (int i = 0; i < n; i++) {
    if (forest_idx[i] < 0 || forest_idx[i] >= (int)m->mp_jungle.size()) return false;
    const Forest& f = m->mp_jungle[forest_idx[i]];
    if (tree_idx[i] < 0 || tree_idx[i] >= f.numberOfTrees()) return false;
    out.m_trees.push_back(f.m_trees[tree_idx[i]]);
  }
  return true;
}

// Leaf ids [patch][tree] (object id inside the tree) for the head-pose forest.
int orc_eval_hp(void* model, void* sample, int H, int W, int stride, int threads, int32_t* leaf_ids, float* headpose, float* variance, long long* visits) {
  Model* m = (Model*)model;
  ImageSample* s = (ImageSample*)sample;
  HeadPoseEstimatorOption o; o.step_size = stride;
  std::vector<Leaf*> leafs;
  float hp = 0, var = 0;
  g_count_visits = visits != nullptr; g_visits = 0;
  getHeadPoseVotesMT(*s, m->hp_forest, Rect{0, 0, W, H}, &hp, &var, o, threads, &leafs);
  g_count_visits = false;
  if (visits) *visits = g_visits;
  if (leaf_ids) for (size_t i = 0; i < leafs.size(); i++) leaf_ids[i] = leafs[i]->object_id;
  if (headpose) *headpose = hp;
  if (variance) *variance = var;
  return (int)leafs.size();
}

// Composition (src/FaceForest.cpp:215-250).  Returns number of trees in the composed forest and
// fills forest_idx/tree_idx (capacity cap).  Counts are clamped to the trees a forest holds; the
// reference indexes out of bounds there (undefined behaviour) — flagged in *flags bit 0.
int orc_compose(void* model, float headpose, float variance, int* tree_counts, int* dominant, int* forest_idx, int* tree_idx, int cap, int* flags) {
  Model* m = (Model*)model;
  int hist_size = (int)m->mp_jungle.size();
  if (flags) *flags = 0;
  if (hist_size != 5) { set_err("jungle must hold 5 forests (FaceForest.cpp:216-222)"); return -1; }
  std::vector<float> poseT(hist_size + 1);
  poseT[0] = -2.5; poseT[1] = -0.35; poseT[2] = -0.20; poseT[3] = -poseT[2]; poseT[4] = -poseT[1]; poseT[5] = -poseT[0];
  std::vector<float> pose_freq(hist_size);
  float max_area = 0;
  int dominant_headpose = 0;
  for (int j = 0; j < hist_size; j++) {
    float area = areaUnderCurve(poseT[j], poseT[j + 1], headpose, std::sqrt((double)variance));
    pose_freq[j] = area;
    if (max_area < area) { max_area = area; dominant_headpose = j; }
  }
  int ntrees_cfg = m->mp_param.ntrees;
  int n = 0;
  for (int i = 0; i < hist_size; i++) {
    float prod = pose_freq[i] * ntrees_cfg;
    double fl = std::floor(prod);
    int ntrees = (fl != fl || fl < -2147483648.0 || fl > 2147483647.0) ? INT32_MIN : (int)fl;   // x86 cvttsd2si semantics
    int avail = m->mp_jungle[i].numberOfTrees();
    if (ntrees > avail) { ntrees = avail; if (flags) *flags |= 1; }
    int cnt = 0;
    for (int j = 0; j < ntrees; j++) {
      if (n < cap) { forest_idx[n] = i; tree_idx[n] = j; n++; cnt++; }
      else if (flags) *flags |= 1;
    }
    if (tree_counts) tree_counts[i] = cnt;
  }
  for (int i = n; i < ntrees_cfg; i++) {
    if (i >= m->mp_jungle[dominant_headpose].numberOfTrees()) { if (flags) *flags |= 1; break; }
    if (n < cap) { forest_idx[n] = dominant_headpose; tree_idx[n] = i; n++; }
  }
  if (dominant) *dominant = dominant_headpose;
  return n;
}

float orc_area_under_curve(float x1, float x2, double mean, double std_) { return areaUnderCurve(x1, x2, mean, std_); }

// FFD stage on an explicit composed forest.  votes_xyw (optional): per part, up to vote_cap
// votes as (x, y, weight) float triples; n_votes[10] always filled.
int orc_eval_ffd(void* model, void* sample, int H, int W, int stride, int threads, const int* forest_idx, const int* tree_idx, int ntrees,
                 int32_t* leaf_ids, int* n_votes, float* votes_xyw, int vote_cap, float* mean_f, int* rounded, int* iters, long long* visits) {
  Model* m = (Model*)model;
  ImageSample* s = (ImageSample*)sample;
  Forest f;
  if (!compose_forest(m, forest_idx, tree_idx, ntrees, f)) { set_err("bad composed forest"); return -1; }
  MultiPartEstimatorOption o; o.step_size = stride;
  std::vector<std::vector<Vote>> votes(o.num_parts);
  std::vector<Leaf*> leafs;
  g_count_visits = visits != nullptr; g_visits = 0;
  getFacialFeaturesVotesMT(*s, f, Rect{0, 0, W, H}, votes, o, threads, &leafs);
  g_count_visits = false;
  if (visits) *visits = g_visits;
  if (leaf_ids) for (size_t i = 0; i < leafs.size(); i++) leaf_ids[i] = leafs[i]->object_id;
  MeanShiftOption ms;
  for (int p = 0; p < o.num_parts; p++) {
    if (n_votes) n_votes[p] = (int)votes[p].size();
    if (votes_xyw)
      for (size_t k = 0; k < votes[p].size() && (int)k < vote_cap; k++) {
        float* v = votes_xyw + ((size_t)p * vote_cap + k) * 3;
        v[0] = (float)votes[p][k].pos.x; v[1] = (float)votes[p][k].pos.y; v[2] = votes[p][k].weight;
      }
    Point r; float mf[2]; int it = 0;
    meanshift_shift(votes[p], r, mf, &it, ms.max_iterations, ms.kernel_size, ms.stopping_criteria);
    if (mean_f) { mean_f[2 * p] = mf[0]; mean_f[2 * p + 1] = mf[1]; }
    if (rounded) { rounded[2 * p] = r.x; rounded[2 * p + 1] = r.y; }
    if (iters) iters[p] = it;
  }
  return (int)leafs.size();
}

// MeanShift on a caller-supplied vote list (x, y, w triples).
void orc_meanshift(const float* votes_xyw, int n, float* mean_f, int* rounded, int* iters) {
  std::vector<Vote> v((size_t)n);
  for (int i = 0; i < n; i++) { v[i].pos.x = (int)votes_xyw[3 * i]; v[i].pos.y = (int)votes_xyw[3 * i + 1]; v[i].weight = votes_xyw[3 * i + 2]; v[i].check = true; }
  Point r; MeanShiftOption ms;
  meanshift_shift(v, r, mean_f, iters, ms.max_iterations, ms.kernel_size, ms.stopping_criteria);
  if (rounded) { rounded[0] = r.x; rounded[1] = r.y; }
}

// FaceForest::analyzeFace (src/FaceForest.cpp:183-258) for one face of a BGR image.
// stats (optional, 4 x int64): [0]=hp visits [1]=ffd visits [2]=total votes [3]=total meanshift iterations
int orc_analyze_face(void* model, const uint8_t* bgr, int rows, int cols, size_t step, int bx, int by, int bw, int bh,
                     const orc_options_t* opt, orc_face_t* out, long long* stats) {
  Model* m = (Model*)model;
  std::memset(out, 0, sizeof *out);
  if (bx < 0 || by < 0 || bw <= 0 || bh <= 0 || bx + bw > cols || by + bh > rows) { set_err("bbox outside image"); return -1; }
  const int threads = opt ? opt->threads : 1;
  // cvtColor on the ROI only: per-pixel op, bit-identical to converting the whole frame (Appendix E.9)
  std::vector<uint8_t> roi((size_t)bw * bh);
  for (int y = 0; y < bh; y++) {
    const uint8_t* p = bgr + (size_t)(by + y) * step + (size_t)bx * 3;
    for (int x = 0; x < bw; x++) roi[(size_t)y * bw + x] = bgr2gray_px(p[3 * x], p[3 * x + 1], p[3 * x + 2]);
  }
  float scale = static_cast<float>(m->hp_param.face_size) / static_cast<float>(bw);
  int sw = (int)(bw * scale), sh = (int)(bh * scale);
  Plane8 img_scaled;
  resize_linear_u8(roi.data(), bh, bw, (size_t)bw, img_scaled, sh, sw);
  out->scaled_w = sw; out->scaled_h = sh; out->scale = scale;
  int patch = m->hp_param.getPatchSize();
  if (sw <= patch || sh <= patch) { set_err("scaled face smaller than a patch"); return -2; }

  int fmask = opt && opt->features_mask ? opt->features_mask : 7;
  ImageSample sample(img_scaled, features_from_mask(fmask), threads);

  HeadPoseEstimatorOption hpo; if (opt && opt->hp_stride > 0) hpo.step_size = opt->hp_stride;
  float headpose = 0, variance = 0;
  g_count_visits = stats != nullptr; g_visits = 0;
  getHeadPoseVotesMT(sample, m->hp_forest, Rect{0, 0, sw, sh}, &headpose, &variance, hpo, threads, nullptr);
  if (stats) { stats[0] = g_visits; stats[1] = stats[2] = stats[3] = 0; }
  out->headpose = headpose; out->variance = variance;
  if (opt && opt->headpose_only) { g_count_visits = false; return 0; }

  int fidx[128], tidx[128], flags = 0;
  int nt = orc_compose(model, headpose, variance, out->tree_counts, &out->dominant, fidx, tidx, 128, &flags);
  if (nt < 0) { g_count_visits = false; return -3; }
  out->flags = flags;
  Forest f;
  compose_forest(m, fidx, tidx, nt, f);

  MultiPartEstimatorOption mpo; if (opt && opt->ffd_stride > 0) mpo.step_size = opt->ffd_stride;
  std::vector<std::vector<Vote>> votes(mpo.num_parts);
  g_visits = 0;
  getFacialFeaturesVotesMT(sample, f, Rect{0, 0, sw, sh}, votes, mpo, threads, nullptr);
  if (stats) stats[1] = g_visits;
  g_count_visits = false;
  MeanShiftOption ms;
  for (int p = 0; p < mpo.num_parts; p++) {
    Point r; float mf[2]; int it = 0;
    meanshift_shift(votes[p], r, mf, &it, ms.max_iterations, ms.kernel_size, ms.stopping_criteria);
    out->ffd_f[p][0] = mf[0]; out->ffd_f[p][1] = mf[1];
    out->ffd_scaled[p][0] = r.x; out->ffd_scaled[p][1] = r.y;
    out->ms_iters[p] = it; out->n_votes[p] = (int)votes[p].size();
    float inv = 1.0f / scale;                       // FaceForest.cpp:256-257: Point_<int> *= float
    out->ffd[p][0] = (int)cv_round(r.x * inv);
    out->ffd[p][1] = (int)cv_round(r.y * inv);
    if (stats) { stats[2] += (long long)votes[p].size(); stats[3] += it; }
  }
  return 0;
}

// Timed batch driver for the CPU baseline: n equal-size BGR crops (box = whole crop), processed
// one after another the way the reference's mains do (eval_ffd.cpp:76-110), each face fanning out
// over `threads` pool workers.  Returns seconds; per-face milliseconds optional.
double orc_analyze_crops_timed(void* model, const uint8_t* bgr_batch, int n, int rows, int cols, const orc_options_t* opt, orc_face_t* out, double* ms_per_face) {
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n; i++) {
    auto a = std::chrono::steady_clock::now();
    orc_analyze_face(model, bgr_batch + (size_t)i * rows * cols * 3, rows, cols, (size_t)cols * 3, 0, 0, cols, rows, opt, &out[i], nullptr);
    auto b = std::chrono::steady_clock::now();
    if (ms_per_face) ms_per_face[i] = std::chrono::duration<double, std::milli>(b - a).count();
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

int orc_hardware_concurrency() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
