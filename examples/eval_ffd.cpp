// eval_ffd on the GPU library, in the reference's own language.
//
// Mirrors MatrixPlayer/face_alignment_cvpr_2012's evaluation main: src/eval_ffd.cpp:57-179 (evalForest, getInterOccularDist
// :33-46, the 90 % / 10 % split per head-pose class, output/errors.txt), loadConfigFile (src/face_utils.cpp:50-140) and
// loadAnnotations (:142-181), written against include/crf_b200_compat.hpp, i.e. the reference's FaceForest interface on
// top of the C ABI.  Host code only; every stage of analyzeFace runs in libcrf_b200.so.
//
// Images: the reference decodes with cv::imread.  This image has no OpenCV C++ and no JPEG decoder, so the driver reads
// binary PPM (P6) files: for an annotation "name.jpg" it opens "name.ppm" next to the annotation file (tools/jpg2ppm.py
// converts a directory).  Built with -DCRF_B200_WITH_OPENCV a maintainer swaps load_image for cv::imread.
//
//   g++ -std=c++17 -O2 -Iinclude examples/eval_ffd.cpp -Lface_alignment_cvpr_2012_b200/_lib -lcrf_b200 -o eval_ffd
//   ./eval_ffd [--all] [--headpose] [--annotations FILE] [--out output/errors.txt] [config_ffd.txt [config_headpose.txt]]
// --headpose runs the reference's other evaluation main instead (src/eval_headpose.cpp:57-139): annotations from the head-pose
// config, "Real:<pose> Predict:<headpose>" per image, no error file.
// A "Path to trees" that names a *.crfb200 file (the pre-packed forest image) instead of a directory loads that
// image; both config files then name the same file.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "crf_b200_compat.hpp"

namespace {

const int NUM_HEADPOSE_CLASSES = 5;          // include/Constants.hpp:66
const float TRAIN_IMAGES_PERCENTAGE = 0.9f;  // include/Constants.hpp:68

typedef crf_b200::ForestParam Param;          // include/Constants.hpp:24-60

struct FaceAnnotation {                       // include/face_utils.hpp:44-52
  std::string url;
  cvlite::Rect bbox;
  int pose = 0;
  std::vector<cvlite::Point> parts;
};

// src/face_utils.cpp:50-140: a label line followed by a value line, eleven times.
bool loadConfigFile(const std::string& path, Param& param) {
  std::ifstream file(path.c_str());
  if (file.is_open()) {
    std::cout << "Open configuration file: " << path << std::endl;
    std::string line;
    auto value = [&]() { std::getline(file, line); std::getline(file, line); return line; };
    param.image_path = value();
    std::cout << "> Path to image annotations: " << line << std::endl;
    param.tree_path = value();
    std::cout << "> Path to trees: " << line << std::endl;
    param.ntrees = std::atoi(value().c_str());
    param.ntests = std::atoi(value().c_str());
    param.max_depth = std::atoi(value().c_str());
    param.min_patches = std::atoi(value().c_str());
    param.nimages = std::atoi(value().c_str());
    param.npatches = std::atoi(value().c_str());
    param.face_size = std::atoi(value().c_str());
    param.patch_size_ratio = (float)std::atof(value().c_str());
    std::istringstream fs(value());
    for (float f; fs >> f;) param.features.push_back((int)f);
    return true;
  }
  std::cout << "Default ForestParam initialization ..." << std::endl;
  param.max_depth = 15; param.min_patches = 20; param.ntests = 250; param.ntrees = 10; param.nimages = 500; param.npatches = 200;
  param.face_size = 100; param.patch_size_ratio = 0.25f;
  return false;
}

// src/face_utils.cpp:142-181: url x y w h pose n (x y)*n, '#' lines skipped.
bool loadAnnotations(const std::string& path, std::vector<FaceAnnotation>& annotations) {
  std::ifstream file(path.c_str());
  if (!file.is_open()) { std::cerr << "(!) Error: Could not open annotations file: " << path << std::endl; return false; }
  std::cout << "Open annotations file: " << path << std::endl;
  std::string line;
  while (std::getline(file, line)) {
    std::vector<std::string> strs;
    std::istringstream ls(line);
    for (std::string t; std::getline(ls, t, ' ');) strs.push_back(t);
    if (strs.size() < 7 || strs[0] == "#") continue;
    FaceAnnotation a;
    a.url = strs[0];
    a.bbox = cvlite::Rect(std::atoi(strs[1].c_str()), std::atoi(strs[2].c_str()), std::atoi(strs[3].c_str()), std::atoi(strs[4].c_str()));
    a.pose = std::atoi(strs[5].c_str());
    const int n = std::atoi(strs[6].c_str());
    if ((int)strs.size() < 7 + 2 * n) continue;
    for (int i = 0; i < n; i++) a.parts.push_back(cvlite::Point(std::atoi(strs[7 + 2 * i].c_str()), std::atoi(strs[8 + 2 * i].c_str())));
    annotations.push_back(a);
  }
  return true;
}

// src/face_utils.cpp:18-28 (the image lives next to the annotation file); P6 instead of cv::imread, stored BGR like cv::Mat.
bool load_image(const std::string& ann_path, const std::string& name, std::vector<unsigned char>& bgr, int& rows, int& cols) {
  const size_t pos = ann_path.rfind('/') + 1, dot = name.rfind('.');
  const std::string file = ann_path.substr(0, pos) + (dot == std::string::npos ? name : name.substr(0, dot)) + ".ppm";
  std::ifstream f(file.c_str(), std::ios::binary);
  if (!f.is_open()) return false;
  std::string magic; int maxv = 0;
  auto token = [&](std::string& t) {   // header tokens, '#' comments skipped
    for (;;) { if (!(f >> t)) return false; if (t[0] != '#') return true; std::getline(f, t); }
  };
  std::string t;
  if (!token(magic) || magic != "P6" || !token(t)) return false;
  cols = std::atoi(t.c_str());
  if (!token(t)) return false;
  rows = std::atoi(t.c_str());
  if (!token(t)) return false;
  maxv = std::atoi(t.c_str());
  if (maxv != 255 || rows <= 0 || cols <= 0) return false;
  f.get();   // the single whitespace after maxval
  bgr.resize((size_t)rows * cols * 3);
  f.read(reinterpret_cast<char*>(bgr.data()), (std::streamsize)bgr.size());
  if (!f) return false;
  for (size_t i = 0; i < bgr.size(); i += 3) std::swap(bgr[i], bgr[i + 2]);   // RGB -> BGR
  return true;
}

// src/eval_ffd.cpp:33-46: distance between the centres of the two eyes (parts 0,1 and 6,7)
float getInterOccularDist(const FaceAnnotation& a) {
  const float lx = (a.parts[0].x + a.parts[1].x) / 2.f, ly = (a.parts[0].y + a.parts[1].y) / 2.f;
  const float rx = (a.parts[6].x + a.parts[7].x) / 2.f, ry = (a.parts[6].y + a.parts[7].y) / 2.f;
  return (float)std::sqrt((double)(lx - rx) * (lx - rx) + (double)(ly - ry) * (ly - ry));
}

}  // namespace

int main(int argc, char** argv) {
  std::string ffd_config_file = "data/config_ffd.txt", headpose_config_file = "data/config_headpose.txt", out_path = "output/errors.txt", ann_override;
  bool all = false, headpose = false;
  int npos = 0;
  for (int i = 1; i < argc; i++) {
    if (!std::strcmp(argv[i], "--all")) all = true;
    else if (!std::strcmp(argv[i], "--headpose")) headpose = true;
    else if (!std::strcmp(argv[i], "--annotations") && i + 1 < argc) ann_override = argv[++i];
    else if (!std::strcmp(argv[i], "--out") && i + 1 < argc) out_path = argv[++i];
    else if (npos++ == 0) ffd_config_file = argv[i];
    else headpose_config_file = argv[i];
  }
  Param hp_param, mp_param;
  if (!loadConfigFile(headpose_config_file, hp_param)) return EXIT_FAILURE;
  if (!loadConfigFile(ffd_config_file, mp_param)) return EXIT_FAILURE;
  if (!ann_override.empty()) mp_param.image_path = hp_param.image_path = ann_override;
  const std::string ann_path = headpose ? hp_param.image_path : mp_param.image_path;   // src/eval_headpose.cpp:118 / src/eval_ffd.cpp:146

  std::vector<FaceAnnotation> annotations;
  if (!loadAnnotations(ann_path, annotations)) return EXIT_FAILURE;

  // src/eval_ffd.cpp:150-165: by head-pose class, the last 10 % of each class is the test set
  std::vector<std::vector<FaceAnnotation>> ann(NUM_HEADPOSE_CLASSES), test_ann(NUM_HEADPOSE_CLASSES);
  for (const FaceAnnotation& a : annotations)
    if (a.pose + 2 >= 0 && a.pose + 2 < NUM_HEADPOSE_CLASSES) ann[a.pose + 2].push_back(a);
  for (size_t i = 0; i < ann.size(); i++) {
    const int num_train_imgs = all ? 0 : static_cast<int>(ann[i].size() * TRAIN_IMAGES_PERCENTAGE);
    test_ann[i].insert(test_ann[i].begin(), ann[i].begin() + num_train_imgs, ann[i].end());
  }

  crf_b200::FaceForestOptions ff_options;
  ff_options.hp_forest_param = hp_param;
  ff_options.mp_forest_param = mp_param;
  crf_b200::FaceForest ff(ff_options);
  if (!ff.is_inizialized) return EXIT_FAILURE;

  std::vector<std::vector<float>> errors;
  for (const auto& cls : test_ann)
    for (const FaceAnnotation& a : cls) {
      std::vector<unsigned char> bgr; int rows = 0, cols = 0;
      if ((!headpose && a.parts.size() < 8) || !load_image(ann_path, a.url, bgr, rows, cols)) { std::cerr << "(!) Error: Could not load: " << a.url << std::endl; continue; }
      crf_b200::Face face;
      ff.analyzeFace(cvlite::Mat(rows, cols, CV_8UC3, bgr.data()), a.bbox, face);
      if (headpose) { std::cout << "Real:" << a.pose << " Predict:" << face.headpose << std::endl; continue; }   // src/eval_headpose.cpp:88
      std::vector<float> err;
      const float iod = getInterOccularDist(a);
      for (size_t j = 0; j < face.ffd_cordinates.size() && j < a.parts.size(); j++) {
        const double dx = a.parts[j].x - face.ffd_cordinates[j].x, dy = a.parts[j].y - face.ffd_cordinates[j].y;
        err.push_back((float)std::sqrt(dx * dx + dy * dy) / iod);
      }
      errors.push_back(err);
    }

  if (headpose) return EXIT_SUCCESS;
  std::ofstream ofs(out_path.c_str(), std::ios::out);
  if (!ofs.is_open()) { std::cerr << "(!) Error: Could not write: " << out_path << std::endl; return EXIT_FAILURE; }
  double sum = 0; size_t n = 0;
  for (const auto& e : errors) {
    for (float v : e) { ofs << v << " "; sum += v; n++; }
    ofs << std::endl;
  }
  std::cout << errors.size() << " faces, mean normalised error " << (n ? sum / n : 0.0) << ", written to " << out_path << std::endl;
  return EXIT_SUCCESS;
}
