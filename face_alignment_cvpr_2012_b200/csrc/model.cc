// Model loading: Boost text-archive v10 reader (no Boost needed) + packed binary image.
// Grammar: SURVEY.md Appendix B; field orders from the reference's serialize() methods
// (include/Tree.hpp:334-343, include/Constants.hpp:44-59, include/TreeNode.hpp:148-164,
//  include/ThresholdSplit.hpp:60-67, include/ImageSample.hpp:83-90,
//  include/opencv_serialization.hpp:65-79, include/HeadPoseSample.hpp:154-161,
//  include/MPSample.hpp:149-158).
#include "model.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <sys/stat.h>
#include <thread>

#include "../../include/crf_b200.h"

namespace crf {
namespace {

// ---- whitespace-separated token cursor over a whole file held in memory
class Cursor {
 public:
  Cursor(const char* b, const char* e) : p_(b), e_(e) {}
  bool good() const { return good_; }
  bool exhausted() {
    ws();
    return p_ >= e_;
  }
  long long integer() {
    ws();
    if (p_ >= e_) return fail();
    bool neg = false;
    if (*p_ == '-') { neg = true; ++p_; } else if (*p_ == '+') ++p_;
    if (p_ >= e_ || *p_ < '0' || *p_ > '9') return fail();
    long long v = 0;
    while (p_ < e_ && *p_ >= '0' && *p_ <= '9') v = v * 10 + (*p_++ - '0');
    if (p_ < e_ && !space(*p_)) return fail();  // e.g. "3.5" where an int is expected
    return neg ? -v : v;
  }
  double real64() {
    ws();
    if (p_ >= e_) return (double)fail();
    char* q = nullptr;
    double v = std::strtod(p_, &q);
    if (q == p_) return (double)fail();
    p_ = q;
    return v;
  }
  float real32() {
    ws();
    if (p_ >= e_) return (float)fail();
    char* q = nullptr;
    float v = std::strtof(p_, &q);
    if (q == p_) return (float)fail();
    p_ = q;
    return v;
  }
  // Boost string: "<len> <bytes>"
  std::string text() {
    long long n = integer();
    if (!good_ || n < 0 || p_ >= e_) { fail(); return std::string(); }
    ++p_;
    if (p_ + n > e_) { fail(); return std::string(); }
    std::string s(p_, p_ + n);
    p_ += n;
    return s;
  }
  std::string bare() {
    ws();
    const char* s = p_;
    while (p_ < e_ && !space(*p_)) ++p_;
    return std::string(s, p_);
  }

 private:
  static bool space(char c) { return c == ' ' || c == '\n' || c == '\r' || c == '\t'; }
  void ws() { while (p_ < e_ && space(*p_)) ++p_; }
  long long fail() { good_ = false; return 0; }
  const char* p_;
  const char* e_;
  bool good_ = true;
};

// Boost writes "tracking version" once per class, the first time the class appears in a file.
struct FirstUse {
  bool node = true, split = true, feature = true, rect = true, hp_leaf = true, mp_leaf = true, points = true, point = true;
};
inline void class_header(Cursor& c, bool& first) {
  if (first) { c.integer(); c.integer(); first = false; }
}

bool read_whole_file(const std::string& path, std::string& buf) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  buf.resize(n > 0 ? (size_t)n : 0);
  size_t got = n > 0 ? std::fread(&buf[0], 1, (size_t)n, f) : 0;
  std::fclose(f);
  return got == buf.size();
}

// Reads one "TreeNode*" record (without recursing into the children). Returns node index or -1.
int read_node(Cursor& c, FirstUse& fu, ForestKind kind, FlatTree& t, std::string& err) {
  long long class_id = c.integer();
  if (!c.good()) { err = "truncated node pointer"; return -1; }
  if (class_id < 0) { err = "null child pointer"; return -1; }
  class_header(c, fu.node);
  long long oid = c.integer();
  const int idx = (int)t.nodes.size();
  if (oid != idx) { err = "object id is not the pre-order index"; return -1; }
  FlatNode n;
  long long depth = c.integer();
  long long is_leaf = c.integer();
  long long has_split = c.integer();
  if (!c.good() || depth < 0 || depth > 62) { err = "bad node header"; return -1; }
  if ((is_leaf != 0) == (has_split != 0)) { err = "node is neither pure leaf nor pure split (unfinished tree?)"; return -1; }
  n.depth = (uint8_t)depth;
  t.max_depth = std::max<int32_t>(t.max_depth, (int32_t)depth);
  if (is_leaf) {
    if (kind == KIND_HEADPOSE) {
      class_header(c, fu.hp_leaf);
      HpLeaf L;
      L.nsamples = (int32_t)c.integer();
      L.foreground = c.real32();
      long long cnt = c.integer();
      c.integer();  // item_version
      if (!c.good() || cnt != CRF_NUM_HEADPOSE_CLASSES) { err = "head-pose leaf must hold 5 labels"; return -1; }
      for (int i = 0; i < 5; i++) L.labels[i] = (int32_t)c.integer();
      L.object_id = idx;
      n.leaf = (int32_t)t.hp_leaves.size();
      t.hp_leaves.push_back(L);
    } else {
      class_header(c, fu.mp_leaf);
      MpLeaf L;
      L.samples = (int32_t)c.integer();
      class_header(c, fu.points);
      long long cnt = c.integer();
      c.integer();
      if (!c.good() || cnt != CRF_NUM_PARTS) { err = "multi-part leaf must hold 10 offsets"; return -1; }
      for (int i = 0; i < CRF_NUM_PARTS; i++) {
        class_header(c, fu.point);
        L.offset[i][0] = (int32_t)c.integer();
        L.offset[i][1] = (int32_t)c.integer();
      }
      long long c2 = c.integer(); c.integer();
      if (!c.good() || c2 != CRF_NUM_PARTS) { err = "multi-part leaf must hold 10 variances"; return -1; }
      for (int i = 0; i < CRF_NUM_PARTS; i++) L.variance[i] = c.real32();
      long long c3 = c.integer(); c.integer();
      if (!c.good() || c3 != CRF_NUM_PARTS) { err = "multi-part leaf must hold 10 probabilities"; return -1; }
      for (int i = 0; i < CRF_NUM_PARTS; i++) L.prob_foreground[i] = c.real32();
      L.foreground = c.real32();
      L.object_id = idx;
      n.leaf = (int32_t)t.mp_leaves.size();
      t.mp_leaves.push_back(L);
    }
  } else {
    class_header(c, fu.split);
    class_header(c, fu.feature);
    long long ch = c.integer();
    class_header(c, fu.rect);
    long long r[8];
    for (int i = 0; i < 8; i++) r[i] = c.integer();
    c.real64();  // info (double) — stored, unused by inference
    long long thr = c.integer();
    if (!c.good()) { err = "truncated split"; return -1; }
    if (ch < 0 || ch > 63) { err = "feature channel outside 0..63"; return -2; }
    for (int i = 0; i < 8; i++)
      if (r[i] < 0 || r[i] > 255) { err = "rectangle field outside 0..255"; return -2; }
    if (r[2] < 1 || r[3] < 1 || r[6] < 1 || r[7] < 1) { err = "empty rectangle"; return -2; }
    n.channel = (uint8_t)ch;
    for (int i = 0; i < 4; i++) { n.r1[i] = (uint8_t)r[i]; n.r2[i] = (uint8_t)r[4 + i]; }
    n.threshold_raw = (int32_t)std::max<long long>(std::min<long long>(thr, INT32_MAX), INT32_MIN);
    n.threshold = (int16_t)std::max<long long>(std::min<long long>(thr, 255), -256);
  }
  if (!c.good()) { err = "truncated node"; return -1; }
  t.nodes.push_back(n);
  return idx;
}

}  // namespace

int parse_tree_file(const std::string& path, ForestKind kind, FlatTree& t, std::string& err) {
  std::string buf;
  if (!read_whole_file(path, buf)) { err = "File not found: " + path; return CRF_ERR_IO; }
  Cursor c(buf.data(), buf.data() + buf.size());
  if (c.bare() != "22" || c.bare() != "serialization::archive") { err = "not a Boost text archive: " + path; return CRF_ERR_FORMAT; }
  c.integer();               // archive library version (10 in the shipped files)
  c.integer(); c.integer();  // class header of Tree
  t = FlatTree();
  t.num_nodes_hdr = (int32_t)c.integer();
  t.i_node = (int32_t)c.integer();
  c.integer(); c.integer();  // class header of ForestParam
  ForestParamLite& fp = t.param;
  fp.max_depth = (int32_t)c.integer(); fp.min_patches = (int32_t)c.integer(); fp.ntests = (int32_t)c.integer();
  fp.ntrees = (int32_t)c.integer(); fp.nimages = (int32_t)c.integer(); fp.npatches = (int32_t)c.integer();
  fp.face_size = (int32_t)c.integer(); fp.patch_size_ratio = c.real32();
  c.text(); c.text();  // tree_path, image_path
  long long nf = c.integer(); c.integer();
  if (!c.good() || nf < 0 || nf > 8) { err = "bad ForestParam in " + path; return CRF_ERR_FORMAT; }
  fp.n_features = (int32_t)nf;
  for (int i = 0; i < nf; i++) fp.features[i] = (int32_t)c.integer();
  c.text();  // m_save_path
  if (!c.good()) { err = "truncated header in " + path; return CRF_ERR_FORMAT; }
  t.nodes.reserve(1 << 15);

  // Pre-order walk with an explicit stack: (parent index, children read so far).
  FirstUse fu;
  std::string why;
  struct Frame { int node; int done; };
  std::vector<Frame> stack;
  int root = read_node(c, fu, kind, t, why);
  if (root < 0) { err = "Exception during tree serialization (" + why + "): " + path; return root == -2 ? CRF_ERR_UNSUPPORTED : CRF_ERR_FORMAT; }
  if (t.nodes[root].leaf < 0) stack.push_back({root, 0});
  while (!stack.empty()) {
    Frame& f = stack.back();
    if (f.done == 2) { stack.pop_back(); continue; }
    const int parent = f.node, which = f.done++;
    int child = read_node(c, fu, kind, t, why);
    if (child < 0) { err = "Exception during tree serialization (" + why + "): " + path; return child == -2 ? CRF_ERR_UNSUPPORTED : CRF_ERR_FORMAT; }
    if (which == 0) t.nodes[parent].left = child; else t.nodes[parent].right = child;
    if (t.nodes[child].leaf < 0) stack.push_back({child, 0});
    if (stack.size() > 64) { err = "tree deeper than 64 levels: " + path; return CRF_ERR_FORMAT; }
  }
  if (!c.exhausted()) { err = "trailing tokens after the root subtree: " + path; return CRF_ERR_FORMAT; }
  return CRF_OK;
}

// Forest<S>::load (include/Forest.hpp:103-129): tree_000 .. tree_{ntrees-1}; any missing or unfinished
// tree aborts the whole load (:125-126, :142-152).
int load_forest_dir(const std::string& dir, int ntrees, ForestKind kind, FlatForest& out, std::string& err) {
  out.kind = kind;
  out.trees.assign((size_t)std::max(ntrees, 0), FlatTree());
  std::vector<int> rc((size_t)std::max(ntrees, 0), 0);
  std::vector<std::string> errs((size_t)std::max(ntrees, 0));
  std::atomic<int> next{0};
  unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  int nthreads = (int)std::min<unsigned>(hw, (unsigned)std::max(ntrees, 1));
  auto worker = [&] {
    for (;;) {
      int i = next.fetch_add(1);
      if (i >= ntrees) return;
      char name[64];
      std::snprintf(name, sizeof name, "/tree_%03d.txt", i);
      rc[i] = parse_tree_file(dir + name, kind, out.trees[i], errs[i]);
      if (rc[i] == CRF_OK && !out.trees[i].isFinished()) { rc[i] = CRF_ERR_FORMAT; errs[i] = "Tree is not finished successfully: " + dir + name; }
    }
  };
  std::vector<std::thread> pool;
  for (int k = 1; k < nthreads; k++) pool.emplace_back(worker);
  worker();
  for (auto& th : pool) th.join();
  for (int i = 0; i < ntrees; i++)
    if (rc[i] != CRF_OK) { err = errs[i]; out.trees.clear(); return rc[i]; }
  return CRF_OK;
}

static bool is_directory(const std::string& p) {
  struct stat st;
  return ::stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

int load_model_dirs(const std::string& hp_dir, int hp_ntrees, const std::string& ffd_dir, int ffd_ntrees, Model& m, std::string& err) {
  m = Model();
  m.hp_ntrees_cfg = hp_ntrees;
  m.mp_ntrees_cfg = ffd_ntrees;
  if (!is_directory(hp_dir)) { err = "(!) Error loading head-pose forest: no directory " + hp_dir; return CRF_ERR_IO; }
  int rc = load_forest_dir(hp_dir, hp_ntrees, KIND_HEADPOSE, m.hp, err);
  if (rc != CRF_OK) return rc;
  if (!is_directory(ffd_dir)) { err = "(!) Error loading facial-feature-detect forest: no directory " + ffd_dir; return CRF_ERR_IO; }
  std::vector<std::string> subdirs;
  if (DIR* d = ::opendir(ffd_dir.c_str())) {
    while (dirent* e = ::readdir(d)) {
      std::string name = e->d_name;
      if (name == "." || name == "..") continue;
      if (is_directory(ffd_dir + "/" + name)) subdirs.push_back(ffd_dir + "/" + name);
    }
    ::closedir(d);
  }
  std::sort(subdirs.begin(), subdirs.end());  // src/FaceForest.cpp:44
  m.jungle.resize(subdirs.size());
  for (size_t i = 0; i < subdirs.size(); i++) {
    rc = load_forest_dir(subdirs[i], ffd_ntrees, KIND_MULTIPART, m.jungle[i], err);
    if (rc != CRF_OK) return rc;
  }
  return validate_model(m, err);
}

static int check_against_features(Model& m, std::string& err) {
  int planes = 0;
  for (size_t i = 0; i < m.features.size(); i++) {
    const int np = planes_of_feature(m.features[i]);
    if (np == 0) { err = "unknown feature channel id " + std::to_string(m.features[i]) + " (include/FeatureChannelFactory.hpp:18-23 defines 0..5)"; return CRF_ERR_UNSUPPORTED; }
    if (i > 0 && m.features[i] == m.features[i - 1]) { err = "feature channel id listed twice"; return CRF_ERR_UNSUPPORTED; }
    planes += np;
  }
  if (planes < 1 || planes > 64) { err = "feature list gives " + std::to_string(planes) + " planes (1..64 supported)"; return CRF_ERR_UNSUPPORTED; }
  m.num_channels = planes;
  auto check_forest = [&](const FlatForest& f, const char* what) -> bool {
    for (const FlatTree& t : f.trees) {
      if (t.param.face_size != m.face_size || (int32_t)std::round(t.param.face_size * t.param.patch_size_ratio) != m.patch_size) {
        err = std::string(what) + ": trees disagree on face/patch size"; return false;
      }
      if ((f.kind == KIND_HEADPOSE && !t.mp_leaves.empty()) || (f.kind == KIND_MULTIPART && !t.hp_leaves.empty())) {
        err = std::string(what) + ": leaf records of the wrong kind"; return false;
      }
      for (const FlatNode& n : t.nodes) {
        if (n.leaf >= 0) continue;
        if (n.r1[0] + n.r1[2] >= m.patch_size || n.r1[1] + n.r1[3] >= m.patch_size || n.r2[0] + n.r2[2] >= m.patch_size ||
            n.r2[1] + n.r2[3] >= m.patch_size) { err = std::string(what) + ": split rectangle leaves the patch"; return false; }
        // the reference would index m_feature_channels out of bounds (src/ImageSample.cpp:38); here it would read another face's planes
        if ((int)n.channel >= planes) {
          err = std::string(what) + ": a split reads feature channel " + std::to_string((int)n.channel) + " but the feature list provides only " + std::to_string(planes) + " planes";
          return false;
        }
      }
    }
    return true;
  };
  if (!check_forest(m.hp, "head-pose forest")) return CRF_ERR_UNSUPPORTED;
  for (auto& f : m.jungle)
    if (!check_forest(f, "facial-feature forest")) return CRF_ERR_UNSUPPORTED;
  return CRF_OK;
}

int validate_model(const Model& mc, std::string& err) {
  Model& m = const_cast<Model&>(mc);
  // a model may hold the head-pose forest only, or facial-feature forests only (Forest<S>::load on its own); what a call needs is
  // checked where it is made
  const FlatTree* first = !m.hp.trees.empty() ? &m.hp.trees[0] : nullptr;
  for (auto& f : m.jungle) if (!first && !f.trees.empty()) first = &f.trees[0];
  if (!first) { err = "model holds no tree"; return CRF_ERR_FORMAT; }
  const ForestParamLite& p = first->param;
  m.face_size = p.face_size;
  m.patch_size = (int32_t)std::round(p.face_size * p.patch_size_ratio);  // ForestParam::getPatchSize (Constants.hpp:26-30)
  if (m.face_size != 125 || m.patch_size != 31) { err = "only face_size 125 / patch 31 models are supported by the device layout"; return CRF_ERR_UNSUPPORTED; }
  // features -> planes in sorted id order (src/ImageSample.cpp:77-90, FeatureChannelFactory.hpp:35-183)
  m.features.assign(p.features, p.features + std::max(0, std::min(p.n_features, 8)));
  std::sort(m.features.begin(), m.features.end());
  return check_against_features(m, err);
}

int set_model_features(Model& m, const int* features, int n, std::string& err) {
  if (n < 1 || n > 6 || !features) { err = "feature list must hold 1..6 ids"; return CRF_ERR_ARG; }
  const std::vector<int32_t> saved = m.features;
  const int saved_planes = m.num_channels;
  m.features.assign(features, features + n);
  std::sort(m.features.begin(), m.features.end());
  const int rc = check_against_features(m, err);
  if (rc != CRF_OK) { m.features = saved; m.num_channels = saved_planes; }
  return rc;
}

// ---------------------------------------------------------------------------------------------
// Packed binary image ("next" row f1).  Little-endian PODs; FNV-1a 64 trailer over everything before it.
// ---------------------------------------------------------------------------------------------
namespace {
const char kMagic[8] = {'C', 'R', 'F', 'B', '2', '0', '0', 'M'};
const uint32_t kVersion = 2;

struct Writer {
  std::vector<uint8_t> b;
  template <class T> void pod(const T& v) { const uint8_t* p = (const uint8_t*)&v; b.insert(b.end(), p, p + sizeof(T)); }
  void raw(const void* p, size_t n) { const uint8_t* q = (const uint8_t*)p; b.insert(b.end(), q, q + n); }
};
struct Reader {
  const uint8_t* p; const uint8_t* e; bool ok = true;
  template <class T> T pod() { T v{}; if (p + sizeof(T) > e) { ok = false; return v; } std::memcpy(&v, p, sizeof(T)); p += sizeof(T); return v; }
  void raw(void* dst, size_t n) { if (p + n > e) { ok = false; return; } std::memcpy(dst, p, n); p += n; }
};
uint64_t fnv1a(const uint8_t* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}
#pragma pack(push, 1)
struct NodeRec { int32_t right; int32_t thr_raw; uint8_t ch, depth; uint8_t r1[4], r2[4]; uint8_t is_leaf; uint8_t pad; };
struct MpLeafRec { int32_t samples; int16_t off[10][2]; float var[10]; float pf[10]; float fg; int32_t oid; };
#pragma pack(pop)

void put_forest(Writer& w, const FlatForest& f) {
  w.pod<int32_t>((int32_t)f.kind);
  w.pod<int32_t>((int32_t)f.trees.size());
  for (const FlatTree& t : f.trees) {
    w.pod(t.num_nodes_hdr); w.pod(t.i_node); w.pod(t.max_depth); w.pod(t.param);
    w.pod<int32_t>((int32_t)t.nodes.size());
    for (const FlatNode& n : t.nodes) {
      NodeRec r{};
      r.right = n.right; r.thr_raw = n.threshold_raw; r.ch = n.channel; r.depth = n.depth;
      std::memcpy(r.r1, n.r1, 4); std::memcpy(r.r2, n.r2, 4); r.is_leaf = n.leaf >= 0;
      w.pod(r);
    }
    w.pod<int32_t>((int32_t)t.hp_leaves.size());
    if (!t.hp_leaves.empty()) w.raw(t.hp_leaves.data(), sizeof(HpLeaf) * t.hp_leaves.size());
    w.pod<int32_t>((int32_t)t.mp_leaves.size());
    for (const MpLeaf& L : t.mp_leaves) {
      MpLeafRec r{};
      r.samples = L.samples;
      for (int i = 0; i < 10; i++) { r.off[i][0] = (int16_t)L.offset[i][0]; r.off[i][1] = (int16_t)L.offset[i][1]; r.var[i] = L.variance[i]; r.pf[i] = L.prob_foreground[i]; }
      r.fg = L.foreground; r.oid = L.object_id;
      w.pod(r);
    }
  }
}
bool get_forest(Reader& r, FlatForest& f) {
  f.kind = (ForestKind)r.pod<int32_t>();
  int32_t nt = r.pod<int32_t>();
  if (!r.ok || nt < 0 || nt > 4096) return false;
  f.trees.assign((size_t)nt, FlatTree());
  for (FlatTree& t : f.trees) {
    t.num_nodes_hdr = r.pod<int32_t>(); t.i_node = r.pod<int32_t>(); t.max_depth = r.pod<int32_t>(); t.param = r.pod<ForestParamLite>();
    int32_t nn = r.pod<int32_t>();
    if (!r.ok || nn < 1 || nn > (1 << 26)) return false;
    t.nodes.resize((size_t)nn);
    int32_t leaf_counter = 0;
    for (int32_t i = 0; i < nn; i++) {
      NodeRec q = r.pod<NodeRec>();
      FlatNode& n = t.nodes[i];
      n.channel = q.ch; n.depth = q.depth; std::memcpy(n.r1, q.r1, 4); std::memcpy(n.r2, q.r2, 4);
      n.threshold_raw = q.thr_raw;
      n.threshold = (int16_t)std::max(std::min(q.thr_raw, 255), -256);
      if (q.is_leaf) { n.leaf = leaf_counter++; n.left = n.right = -1; }
      else { n.leaf = -1; n.left = i + 1; n.right = q.right; if (n.right <= i + 1 || n.right >= nn) return false; }
    }
    int32_t nh = r.pod<int32_t>();
    if (!r.ok || nh < 0 || nh > nn) return false;
    t.hp_leaves.resize((size_t)nh);
    if (nh) r.raw(t.hp_leaves.data(), sizeof(HpLeaf) * (size_t)nh);
    int32_t nm = r.pod<int32_t>();
    if (!r.ok || nm < 0 || nm > nn) return false;
    t.mp_leaves.resize((size_t)nm);
    for (MpLeaf& L : t.mp_leaves) {
      MpLeafRec q = r.pod<MpLeafRec>();
      L.samples = q.samples;
      for (int i = 0; i < 10; i++) { L.offset[i][0] = q.off[i][0]; L.offset[i][1] = q.off[i][1]; L.variance[i] = q.var[i]; L.prob_foreground[i] = q.pf[i]; }
      L.foreground = q.fg; L.object_id = q.oid;
    }
    if (!r.ok || leaf_counter != nh + nm) return false;
  }
  return r.ok;
}
}  // namespace

int save_model_packed(const Model& m, const std::string& path, std::string& err) {
  for (const auto& f : m.jungle)
    for (const auto& t : f.trees)
      for (const auto& L : t.mp_leaves)
        for (int i = 0; i < 10; i++)
          if (L.offset[i][0] < -32768 || L.offset[i][0] > 32767 || L.offset[i][1] < -32768 || L.offset[i][1] > 32767) { err = "leaf offset does not fit int16"; return CRF_ERR_UNSUPPORTED; }
  Writer w;
  w.raw(kMagic, 8);
  w.pod(kVersion);
  w.pod(m.hp_ntrees_cfg); w.pod(m.mp_ntrees_cfg); w.pod(m.face_size); w.pod(m.patch_size); w.pod(m.num_channels);
  put_forest(w, m.hp);
  w.pod<int32_t>((int32_t)m.jungle.size());
  for (const auto& f : m.jungle) put_forest(w, f);
  uint64_t h = fnv1a(w.b.data(), w.b.size());
  w.pod(h);
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) { err = "cannot write " + path; return CRF_ERR_IO; }
  bool ok = std::fwrite(w.b.data(), 1, w.b.size(), f) == w.b.size();
  ok = (std::fclose(f) == 0) && ok;
  if (!ok) { err = "short write " + path; return CRF_ERR_IO; }
  return CRF_OK;
}

int load_model_packed(const std::string& path, Model& m, std::string& err) {
  std::string buf;
  if (!read_whole_file(path, buf)) { err = "File not found: " + path; return CRF_ERR_IO; }
  if (buf.size() < 8 + 4 + 8 || std::memcmp(buf.data(), kMagic, 8) != 0) { err = "not a packed CRF model: " + path; return CRF_ERR_FORMAT; }
  const uint8_t* b = (const uint8_t*)buf.data();
  uint64_t want;
  std::memcpy(&want, b + buf.size() - 8, 8);
  if (fnv1a(b, buf.size() - 8) != want) { err = "checksum mismatch: " + path; return CRF_ERR_FORMAT; }
  Reader r{b + 8, b + buf.size() - 8};
  if (r.pod<uint32_t>() != kVersion) { err = "packed model version mismatch: " + path; return CRF_ERR_FORMAT; }
  m = Model();
  m.hp_ntrees_cfg = r.pod<int32_t>(); m.mp_ntrees_cfg = r.pod<int32_t>(); m.face_size = r.pod<int32_t>(); m.patch_size = r.pod<int32_t>(); m.num_channels = r.pod<int32_t>();
  if (!get_forest(r, m.hp)) { err = "corrupt packed model: " + path; return CRF_ERR_FORMAT; }
  int32_t nj = r.pod<int32_t>();
  if (!r.ok || nj < 0 || nj > 64) { err = "corrupt packed model: " + path; return CRF_ERR_FORMAT; }
  m.jungle.resize((size_t)nj);
  for (auto& f : m.jungle)
    if (!get_forest(r, f)) { err = "corrupt packed model: " + path; return CRF_ERR_FORMAT; }
  return validate_model(m, err);
}

}  // namespace crf
