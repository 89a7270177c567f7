// Compiles against include/crf_b200_compat.hpp (the binding a reference maintainer adds) and
// exercises the reference-shaped error behaviour: a FaceForest whose forests cannot be loaded stays
// un-initialised and asserts on use (src/FaceForest.cpp:31-36,167,191).
// usage: compat_smoke <hp_dir> <ffd_dir> [run]
#include <cstdio>
#include <cstring>
#include <vector>

#include "crf_b200_compat.hpp"

int main(int argc, char** argv) {
  using namespace crf_b200;
  FaceForestOptions bad;
  bad.hp_forest_param.tree_path = "/nonexistent/trees_headpose";
  bad.hp_forest_param.ntrees = 15;
  bad.mp_forest_param.tree_path = "/nonexistent/trees_ffd";
  bad.mp_forest_param.ntrees = 20;
  FaceForest ff(bad);
  if (ff.is_inizialized) { std::puts("FAIL: initialised from missing directories"); return 1; }
  bool threw = false;
  try {
    Face f;
    std::vector<unsigned char> px(100 * 100 * 3, 0);
    ff.analyzeFace(cvlite::Mat(100, 100, CV_8UC3, px.data()), cvlite::Rect(0, 0, 100, 100), f);
  } catch (const std::logic_error&) { threw = true; }
  if (!threw) { std::puts("FAIL: use before init did not assert"); return 1; }
  if (argc >= 3) {
    FaceForestOptions o;
    o.hp_forest_param.tree_path = argv[1]; o.hp_forest_param.ntrees = 15;
    o.mp_forest_param.tree_path = argv[2]; o.mp_forest_param.ntrees = 20;
    FaceForest g(o);
    if (argc >= 4 && std::strcmp(argv[3], "run") == 0) {
      if (!g.is_inizialized) { std::puts("FAIL: could not initialise on the GPU"); return 1; }
      std::vector<unsigned char> px(120 * 160 * 3);
      for (size_t i = 0; i < px.size(); i++) px[i] = (unsigned char)((i * 2654435761u) >> 24);
      std::vector<Face> faces;
      std::vector<cvlite::Rect> boxes{cvlite::Rect(10, 5, 100, 100), cvlite::Rect(40, 10, 90, 105)};
      g.analyzeImage(cvlite::Mat(120, 160, CV_8UC3, px.data()), boxes, faces);
      if (faces.size() != 2 || faces[0].ffd_cordinates.size() != 10) { std::puts("FAIL: bad result shape"); return 1; }
      std::printf("headpose %.6f %.6f first point (%d,%d)\n", faces[0].headpose, faces[1].headpose, faces[0].ffd_cordinates[0].x, faces[0].ffd_cordinates[0].y);
    } else if (g.is_inizialized) {
      std::puts("note: a GPU is present, context created");
    } else {
      std::printf("no GPU: %s\n", crf_last_error());
    }
  }
  std::puts("compat_smoke ok");
  return 0;
}
