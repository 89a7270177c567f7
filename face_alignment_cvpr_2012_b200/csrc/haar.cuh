// Face-box source of FaceForest::detectFace (reference src/FaceForest.cpp:136-159):
// cv::CascadeClassifier::detectMultiScale for the stump-based HAAR cascade the reference ships
// (data/haarcascade_frontalface_alt.xml, new-format XML), evaluated on the GPU.
//
// The algorithm is OpenCV's (objdetect/cascadedetect.cpp; not under /root/reference): per scale factor = 1.3^k a bilinear
// pyramid level of the gray frame, integral images of values and squares, then every window position of the level's
// ystep grid walks the 22 stages / 2135 stumps until a stage sum falls below its threshold; surviving windows are grouped
// by cv::groupRectangles.  Here: one thread per window position (adjacent threads = adjacent windows, so the corner
// loads of a stump coalesce; the cascade tables are broadcast loads), early exit per thread, results as two byte maps
// (passed all stages / rejected by stage 0 — the latter drives OpenCV's "skip one more step" scan rule, which is a
// sequential dependence along x and is applied on the host together with the grouping, as OpenCV does it serially too).
// Arithmetic follows oracle/haar.py, which is pinned against cv2 4.13 on the shipped images (IoU of the boxes >= 0.9:
// cv2 builds its pyramid with INTER_LINEAR_EXACT, this path with the INTER_LINEAR arithmetic of k_gray_resize).
#pragma once

namespace crf {

struct HaarStage { int first, count; float threshold; };
struct HaarWeak { int feature; float threshold, left, right; };
struct HaarRect { int x, y, w, h; float weight; };
struct HaarFeature { HaarRect r[3]; };

struct Cascade {
  int win_w = 0, win_h = 0;
  std::vector<HaarStage> stages;
  std::vector<HaarWeak> weak;
  std::vector<HaarFeature> features;
};

// ---- minimal XML reader for OpenCV's FileStorage cascade files: elements with text, no attributes needed ----------------
struct XmlNode {
  std::string name, text;
  std::vector<XmlNode> kids;
  const XmlNode* child(const char* n) const { for (auto& k : kids) if (k.name == n) return &k; return nullptr; }
};
static bool xml_parse(const std::string& s, size_t& i, XmlNode& out) {
  // precondition: s[i] == '<' of an opening tag
  size_t e = s.find('>', i);
  if (e == std::string::npos) return false;
  std::string tag = s.substr(i + 1, e - i - 1);
  const bool self_closing = !tag.empty() && tag.back() == '/';
  if (self_closing) tag.pop_back();
  out.name = tag.substr(0, tag.find_first_of(" \t\r\n"));
  i = e + 1;
  if (self_closing) return true;
  for (;;) {
    const size_t lt = s.find('<', i);
    if (lt == std::string::npos) return false;
    out.text.append(s, i, lt - i);
    if (s.compare(lt, 4, "<!--") == 0) { const size_t c = s.find("-->", lt); if (c == std::string::npos) return false; i = c + 3; continue; }
    if (s[lt + 1] == '/') { const size_t c = s.find('>', lt); if (c == std::string::npos) return false; i = c + 1; return true; }
    out.kids.emplace_back();
    i = lt;
    if (!xml_parse(s, i, out.kids.back())) return false;
  }
}

static int load_cascade(const std::string& path, Cascade& c, std::string& err) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) { err = "File not found: " + path; return CRF_ERR_IO; }
  std::string s;
  char buf[1 << 16];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
  std::fclose(f);
  size_t i = s.find("<opencv_storage");
  XmlNode root;
  if (i == std::string::npos || !xml_parse(s, i, root)) { err = "not an OpenCV cascade file: " + path; return CRF_ERR_FORMAT; }
  const XmlNode* cas = root.child("cascade");
  if (!cas || !cas->child("stages") || !cas->child("features") || !cas->child("width") || !cas->child("height")) {
    err = "unsupported cascade layout (expected the new-format opencv-cascade-classifier): " + path; return CRF_ERR_FORMAT;
  }
  if (cas->child("featureType") && cas->child("featureType")->text.find("HAAR") == std::string::npos) { err = "only HAAR cascades are supported"; return CRF_ERR_UNSUPPORTED; }
  c.win_w = std::atoi(cas->child("width")->text.c_str());
  c.win_h = std::atoi(cas->child("height")->text.c_str());
  for (const XmlNode& ft : cas->child("features")->kids) {
    HaarFeature hf{};
    if (ft.child("tilted") && std::atoi(ft.child("tilted")->text.c_str()) != 0) { err = "tilted features are not supported"; return CRF_ERR_UNSUPPORTED; }
    const XmlNode* rects = ft.child("rects");
    if (!rects || rects->kids.empty() || rects->kids.size() > 3) { err = "feature with an unsupported rectangle count"; return CRF_ERR_FORMAT; }
    for (size_t k = 0; k < rects->kids.size(); k++) {
      double v[5] = {0, 0, 0, 0, 0};
      if (std::sscanf(rects->kids[k].text.c_str(), "%lf %lf %lf %lf %lf", &v[0], &v[1], &v[2], &v[3], &v[4]) != 5) { err = "bad feature rectangle"; return CRF_ERR_FORMAT; }
      hf.r[k] = HaarRect{(int)v[0], (int)v[1], (int)v[2], (int)v[3], (float)v[4]};
      if (hf.r[k].x < 0 || hf.r[k].y < 0 || hf.r[k].w < 1 || hf.r[k].h < 1 || hf.r[k].x + hf.r[k].w > c.win_w || hf.r[k].y + hf.r[k].h > c.win_h) { err = "feature rectangle outside the window"; return CRF_ERR_FORMAT; }
    }
    c.features.push_back(hf);
  }
  for (const XmlNode& st : cas->child("stages")->kids) {
    const XmlNode* thr = st.child("stageThreshold");
    const XmlNode* wcs = st.child("weakClassifiers");
    if (!thr || !wcs) { err = "stage without threshold / classifiers"; return CRF_ERR_FORMAT; }
    HaarStage hs{(int)c.weak.size(), 0, (float)std::atof(thr->text.c_str())};
    for (const XmlNode& wc : wcs->kids) {
      const XmlNode* in = wc.child("internalNodes");
      const XmlNode* lv = wc.child("leafValues");
      int a = 0, b = 0, fi = 0; double t = 0, l = 0, r = 0;
      if (!in || !lv || std::sscanf(in->text.c_str(), "%d %d %d %lf", &a, &b, &fi, &t) != 4 || std::sscanf(lv->text.c_str(), "%lf %lf", &l, &r) != 2) { err = "bad weak classifier"; return CRF_ERR_FORMAT; }
      if (a != 0 || b != -1) { err = "only stump-based cascades are supported"; return CRF_ERR_UNSUPPORTED; }
      if (fi < 0 || fi >= (int)c.features.size()) { err = "weak classifier names a missing feature"; return CRF_ERR_FORMAT; }
      c.weak.push_back(HaarWeak{fi, (float)t, (float)l, (float)r});
      hs.count++;
    }
    c.stages.push_back(hs);
  }
  if (c.stages.empty() || c.win_w < 3 || c.win_h < 3) { err = "empty cascade"; return CRF_ERR_FORMAT; }
  return CRF_OK;
}

// ---- kernels ---------------------------------------------------------------------------------------------------------------
// pyramid level: BGR2GRAY of the source pixels, then INTER_LINEAR (k_gray_resize's arithmetic) to sw x sh; u8 [sh][sw]
__global__ void k_haar_level(const uint8_t* __restrict__ bgr, int rows, int cols, size_t step, uint8_t* __restrict__ out, int sh, int sw, double scale_x, double scale_y) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
  if (dx >= sw || dy >= sh) return;
  int v;
  if (sw == cols && sh == rows) {
    v = gray_at(bgr, step, dy, dx);
  } else {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= cols - 1) { fx = 0; sx = cols - 1; }
    const int a0 = __float2int_rn((1.f - fx) * 2048.f), a1 = __float2int_rn(fx * 2048.f);
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)floorf(fy);
    fy -= sy;
    const int b0 = __float2int_rn((1.f - fy) * 2048.f), b1 = __float2int_rn(fy * 2048.f);
    const int sy0 = min(max(sy, 0), rows - 1), sy1 = min(max(sy + 1, 0), rows - 1), sx1 = min(sx + 1, cols - 1);
    const int row0 = gray_at(bgr, step, sy0, sx) * a0 + gray_at(bgr, step, sy0, sx1) * a1;
    const int row1 = gray_at(bgr, step, sy1, sx) * a0 + gray_at(bgr, step, sy1, sx1) * a1;
    v = (((b0 * (row0 >> 4)) >> 16) + ((b1 * (row1 >> 4)) >> 16) + 2) >> 2;
    v = min(max(v, 0), 255);
  }
  out[(size_t)dy * sw + dx] = (uint8_t)v;
}

// integral images of values (S) and squares (Q, modulo 2^32 as OpenCV keeps them), (sh + 1) x (sw + 1), pitch P = sw + 1.
// Pass 1: one warp per row, inclusive scan along x.  Pass 2: one thread per column, running sum down the rows.
__global__ void k_haar_rowscan(const uint8_t* __restrict__ img, int sh, int sw, uint32_t* __restrict__ S, uint32_t* __restrict__ Q) {
  const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (y >= sh) return;
  const int P = sw + 1;
  uint32_t cs = 0, cq = 0;
  if (lane == 0) { S[(size_t)(y + 1) * P] = 0; Q[(size_t)(y + 1) * P] = 0; }
  for (int x0 = 0; x0 < sw; x0 += 32) {
    const int x = x0 + lane;
    uint32_t v = x < sw ? img[(size_t)y * sw + x] : 0u, q = v * v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t nv = __shfl_up_sync(0xffffffffu, v, o), nq = __shfl_up_sync(0xffffffffu, q, o);
      if (lane >= o) { v += nv; q += nq; }
    }
    if (x < sw) { S[(size_t)(y + 1) * P + x + 1] = cs + v; Q[(size_t)(y + 1) * P + x + 1] = cq + q; }
    cs += __shfl_sync(0xffffffffu, v, 31);
    cq += __shfl_sync(0xffffffffu, q, 31);
  }
}
__global__ void k_haar_colscan(int sh, int sw, uint32_t* __restrict__ S, uint32_t* __restrict__ Q) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, P = sw + 1;
  if (x > sw) return;
  uint32_t s = 0, q = 0;
  S[x] = 0; Q[x] = 0;
  for (int y = 1; y <= sh; y++) {
    s += S[(size_t)y * P + x]; q += Q[(size_t)y * P + x];
    S[(size_t)y * P + x] = s; Q[(size_t)y * P + x] = q;
  }
}

// One thread per window of the level's ystep grid.  flags[iy][ix]: bit 0 = passed every stage, bit 1 = rejected by stage 0.
__global__ void __launch_bounds__(128) k_haar_eval(const uint32_t* __restrict__ S, const uint32_t* __restrict__ Q, int sw, int sh, int win_w, int win_h, int ystep,
                                                   int nx, int ny, const HaarStage* __restrict__ stages, int nstages, const HaarWeak* __restrict__ weak,
                                                   const HaarFeature* __restrict__ feats, uint8_t* __restrict__ flags) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y;
  if (ix >= nx || iy >= ny) return;
  const int x = ix * ystep, y = iy * ystep, P = sw + 1;
  const uint32_t* __restrict__ s0 = S + (size_t)y * P + x;
  const uint32_t* __restrict__ q0 = Q + (size_t)y * P + x;
  auto rs = [&](const uint32_t* __restrict__ p, int rx, int ry, int rw, int rh) -> uint32_t {
    return p[(size_t)(ry + rh) * P + rx + rw] - p[(size_t)ry * P + rx + rw] - p[(size_t)(ry + rh) * P + rx] + p[(size_t)ry * P + rx];
  };
  // HaarEvaluator::setWindow: variance normalisation over the inner (w - 2) x (h - 2) rect
  const double area = (double)((win_w - 2) * (win_h - 2));
  const double vs = (double)rs(s0, 1, 1, win_w - 2, win_h - 2), vq = (double)rs(q0, 1, 1, win_w - 2, win_h - 2);
  const double nf = area * vq - vs * vs;
  uint8_t out = 0;
  if (nf > 0.) {
    const float inv = (float)(1.0 / sqrt(nf));
    if ((float)area * inv < 0.1f) {
      out = 1;
      for (int si = 0; si < nstages; si++) {
        const HaarStage st = stages[si];
        float sum = 0.f;
        for (int k = st.first; k < st.first + st.count; k++) {
          const HaarWeak wk = weak[k];
          const HaarFeature& f = feats[wk.feature];
          float v = 0.f;
#pragma unroll
          for (int r = 0; r < 3; r++) {
            const HaarRect hr = f.r[r];
            if (hr.weight != 0.f) v = v + hr.weight * (float)rs(s0, hr.x, hr.y, hr.w, hr.h);
          }
          sum = sum + ((v * inv < wk.threshold) ? wk.left : wk.right);
        }
        if (!(sum >= st.threshold)) { out = si == 0 ? 2 : 0; break; }
      }
    }
  }
  flags[(size_t)iy * nx + ix] = out;
}

// ---- host: scales, scan rule, cv::groupRectangles ----------------------------------------------------------------------------
static void group_rectangles(std::vector<crf_rect_t>& rects, int group_threshold, double eps) {
  const int n = (int)rects.size();
  if (group_threshold <= 0 || n == 0) return;
  std::vector<int> parent((size_t)n);
  for (int i = 0; i < n; i++) parent[(size_t)i] = i;
  auto find = [&](int i) { while (parent[(size_t)i] != i) { parent[(size_t)i] = parent[(size_t)parent[(size_t)i]]; i = parent[(size_t)i]; } return i; };
  auto similar = [&](const crf_rect_t& a, const crf_rect_t& b) {
    const double d = eps * (std::min(a.width, b.width) + std::min(a.height, b.height)) * 0.5;
    return std::abs(a.x - b.x) <= d && std::abs(a.y - b.y) <= d && std::abs(a.x + a.width - b.x - b.width) <= d && std::abs(a.y + a.height - b.y - b.height) <= d;
  };
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++)
      if (similar(rects[(size_t)i], rects[(size_t)j])) { const int a = find(i), b = find(j); if (a != b) parent[(size_t)b] = a; }
  std::vector<int> label((size_t)n, -1), root_label((size_t)n, -1);
  int nc = 0;
  for (int i = 0; i < n; i++) { const int r = find(i); if (root_label[(size_t)r] < 0) root_label[(size_t)r] = nc++; label[(size_t)i] = root_label[(size_t)r]; }
  std::vector<long long> acc((size_t)nc * 4, 0);
  std::vector<int> cnt((size_t)nc, 0);
  for (int i = 0; i < n; i++) {
    long long* a = &acc[(size_t)label[(size_t)i] * 4];
    a[0] += rects[(size_t)i].x; a[1] += rects[(size_t)i].y; a[2] += rects[(size_t)i].width; a[3] += rects[(size_t)i].height;
    cnt[(size_t)label[(size_t)i]]++;
  }
  std::vector<crf_rect_t> rr((size_t)nc);
  for (int l = 0; l < nc; l++) {
    const float s = 1.f / (float)cnt[(size_t)l];
    rr[(size_t)l] = crf_rect_t{(int)std::lrintf((float)acc[(size_t)l * 4] * s), (int)std::lrintf((float)acc[(size_t)l * 4 + 1] * s),
                              (int)std::lrintf((float)acc[(size_t)l * 4 + 2] * s), (int)std::lrintf((float)acc[(size_t)l * 4 + 3] * s)};
  }
  std::vector<crf_rect_t> keep;
  for (int i = 0; i < nc; i++) {
    if (cnt[(size_t)i] <= group_threshold) continue;
    const crf_rect_t r1 = rr[(size_t)i];
    const int n1 = cnt[(size_t)i];
    int j = 0;
    for (; j < nc; j++) {
      const int n2 = cnt[(size_t)j];
      if (j == i || n2 <= group_threshold) continue;
      const crf_rect_t r2 = rr[(size_t)j];
      const int dx = (int)std::lrint(r2.width * eps), dy = (int)std::lrint(r2.height * eps);
      if (r1.x >= r2.x - dx && r1.y >= r2.y - dy && r1.x + r1.width <= r2.x + r2.width + dx && r1.y + r1.height <= r2.y + r2.height + dy && (n2 > std::max(3, n1) || n1 < 3)) break;
    }
    if (j == nc) keep.push_back(r1);
  }
  rects.swap(keep);
}

}  // namespace crf

struct crf_cascade { crf::Cascade c; };
