// Host-side forests of the CRF path: flat, pointer-free trees in pre-order (index == Boost object id).
// Replaces Forest<S>::load / Tree<S>::load (reference include/Forest.hpp:103-153, include/Tree.hpp:193-237)
// and the jungle enumeration of FaceForest::FaceForest (src/FaceForest.cpp:39-55).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace crf {

enum ForestKind : int { KIND_HEADPOSE = 0, KIND_MULTIPART = 1 };

// ForestParam (include/Constants.hpp:24-60), the fields inference reads plus the stored training ones.
struct ForestParamLite {
  int32_t max_depth = 0, min_patches = 0, ntests = 0, ntrees = 0, nimages = 0, npatches = 0, face_size = 0;
  float patch_size_ratio = 0.f;
  int32_t n_features = 0;
  int32_t features[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

// One node in pre-order.  Internal nodes carry the ThresholdSplit<SimplePatchFeature>
// (include/ThresholdSplit.hpp:51-67, include/ImageSample.hpp:79-90); leaves point into the leaf arrays.
struct FlatNode {
  int32_t left = -1, right = -1;  // pre-order indices of the children (-1 for leaves)
  int32_t leaf = -1;              // index into the tree's leaf arrays (-1 for internal nodes)
  int16_t threshold = 0;          // clamped to [-256, 255]: |mean1 - mean2| <= 255, so the clamp is behaviour-preserving
  uint8_t channel = 0;
  uint8_t depth = 0;
  uint8_t r1[4] = {0, 0, 0, 0};   // x, y, w, h
  uint8_t r2[4] = {0, 0, 0, 0};
  int32_t threshold_raw = 0;      // value as stored in the archive
};

// HeadPoseLeaf (include/HeadPoseSample.hpp:144-162)
struct HpLeaf {
  int32_t nsamples = 0;
  float foreground = 0.f;
  int32_t labels[5] = {0, 0, 0, 0, 0};
  int32_t object_id = -1;
};

// MPLeaf (include/MPSample.hpp:137-159)
struct MpLeaf {
  int32_t samples = 0;
  int32_t offset[10][2] = {};
  float variance[10] = {};
  float prob_foreground[10] = {};
  float foreground = 0.f;
  int32_t object_id = -1;
};

struct FlatTree {
  int32_t num_nodes_hdr = 0, i_node = 0;  // Tree::m_num_nodes / i_node (include/Tree.hpp:72-79, :334-343)
  int32_t max_depth = 0;
  ForestParamLite param;
  std::vector<FlatNode> nodes;
  std::vector<HpLeaf> hp_leaves;
  std::vector<MpLeaf> mp_leaves;
  bool isFinished() const { return num_nodes_hdr != 0 && i_node == num_nodes_hdr; }
};

struct FlatForest {
  ForestKind kind = KIND_HEADPOSE;
  std::vector<FlatTree> trees;
};

struct Model {
  FlatForest hp;
  std::vector<FlatForest> jungle;  // pose forests in lexicographic directory order
  int32_t hp_ntrees_cfg = 0, mp_ntrees_cfg = 0;
  int32_t face_size = 125;
  int32_t patch_size = 31;
  int32_t num_channels = 38;  // features {0,1,2}: gray + 35 Gabor + 2 Sobel
  // Feature channels in the order ImageSample::extractFeatureChannels builds them (sorted ids, src/ImageSample.cpp:77-90);
  // taken from the head-pose forest's stored ForestParam (src/FaceForest.cpp:207 uses hp_forest_param.features for BOTH forests)
  std::vector<int32_t> features;
};

// Planes FeatureChannelFactory::extractChannel appends for one feature id (include/FeatureChannelFactory.hpp:35-183); 0 = unknown id.
inline int planes_of_feature(int f) { return f == 0 ? 1 : f == 1 ? 35 : f == 2 ? 2 : f == 3 ? 2 : f == 4 ? 1 : f == 5 ? 1 : 0; }

// All return 0 on success, negative crf_status otherwise, and fill err.
int parse_tree_file(const std::string& path, ForestKind kind, FlatTree& out, std::string& err);
int load_forest_dir(const std::string& dir, int ntrees, ForestKind kind, FlatForest& out, std::string& err);
int load_model_dirs(const std::string& hp_dir, int hp_ntrees, const std::string& ffd_dir, int ffd_ntrees, Model& out, std::string& err);
int save_model_packed(const Model& m, const std::string& path, std::string& err);
int load_model_packed(const std::string& path, Model& out, std::string& err);
int validate_model(const Model& m, std::string& err);
// Replaces the feature list (sorted on entry) and re-validates (every split must read a plane the list provides).
int set_model_features(Model& m, const int* features, int n, std::string& err);

}  // namespace crf
