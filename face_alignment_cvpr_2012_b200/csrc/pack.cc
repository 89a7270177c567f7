// Packs host-side FlatForests into the device image (device_forest.h).
#include <algorithm>
#include <deque>

#include "../../include/crf_b200.h"
#include "device_forest.h"

namespace crf {

int pack_forests(const std::vector<const FlatForest*>& forests, ForestKind kind, const PackOptions& opt, PackedForest& out, std::string& err) {
  out = PackedForest();
  for (const FlatForest* f : forests) {
    out.forest_base.push_back((int32_t)out.roots.size());
    out.forest_ntrees.push_back((int32_t)f->trees.size());
    for (const FlatTree& t : f->trees) {
      if (t.nodes.empty()) { err = "empty tree"; return CRF_ERR_FORMAT; }
      const int32_t leaf_base = (int32_t)out.leaf_oid.size();
      out.leaf_base.push_back(leaf_base);
      out.max_depth = t.max_depth > out.max_depth ? t.max_depth : out.max_depth;
      // leaves keep the tree's pre-order leaf numbering
      const size_t nleaves = kind == KIND_HEADPOSE ? t.hp_leaves.size() : t.mp_leaves.size();
      out.leaf_oid.resize(leaf_base + nleaves);
      if (kind == KIND_HEADPOSE) {
        out.hp_m.resize(leaf_base + nleaves);
        for (size_t l = 0; l < nleaves; l++) {
          const HpLeaf& L = t.hp_leaves[l];
          out.leaf_oid[leaf_base + l] = L.object_id;
          // src/face_utils.cpp:222-228, same float operation order
          float m = -1.f;
          if (L.foreground > opt.hp_min_foreground) {
            m = 0;
            for (int j = 0; j < 5; j++) m += L.labels[j] * j;
            m /= (L.nsamples * L.foreground);
            if (!(m >= 0.f)) m = __builtin_nanf("");  // cannot happen for finished trees; keeps the NaN path visible
          }
          out.hp_m[leaf_base + l] = m;
        }
      } else {
        out.mp_leaf.resize(leaf_base + nleaves);
        out.mp_mask.resize(leaf_base + nleaves);
        for (size_t l = 0; l < nleaves; l++) {
          const MpLeaf& L = t.mp_leaves[l];
          out.leaf_oid[leaf_base + l] = L.object_id;
          DevMpLeaf d{};
          uint16_t mask = 0;
          for (int i = 0; i < kParts; i++) {
            if (L.offset[i][0] < -32768 || L.offset[i][0] > 32767 || L.offset[i][1] < -32768 || L.offset[i][1] > 32767) {
              err = "leaf offset does not fit int16"; return CRF_ERR_UNSUPPORTED;
            }
            d.off[i][0] = (int16_t)L.offset[i][0];
            d.off[i][1] = (int16_t)L.offset[i][1];
            // src/face_utils.cpp:284-290
            float min_pf = opt.ffd_min_pf;
            if (i == 0 || i == 7) min_pf *= 1.5;
            if (L.foreground > opt.ffd_min_foreground && L.prob_foreground[i] > min_pf && L.variance[i] < opt.ffd_max_variance &&
                L.samples > opt.ffd_min_samples)
              mask |= (uint16_t)(1u << i);
          }
          d.weight = L.foreground;
          out.mp_leaf[leaf_base + l] = d;
          out.mp_mask[leaf_base + l] = mask;
        }
      }
      // breadth-first slot assignment
      const int32_t base = (int32_t)out.slots.size();
      out.roots.push_back(base);
      std::vector<int32_t> slot_of(t.nodes.size(), -1);
      // DevSlotN index of every internal node, ~(global leaf) of every leaf
      std::vector<int32_t> n_of(t.nodes.size(), 0);
      auto tag = [&](int32_t ni) {
        if (t.nodes[ni].leaf >= 0) n_of[ni] = ~(leaf_base + t.nodes[ni].leaf);
        else { n_of[ni] = (int32_t)out.slotsn.size(); out.slotsn.emplace_back(); }
      };
      tag(0);
      out.nroot.push_back(n_of[0]);
      std::deque<int32_t> q;
      out.slots.emplace_back();
      slot_of[0] = base;
      q.push_back(0);
      while (!q.empty()) {
        const int32_t ni = q.front();
        q.pop_front();
        const FlatNode& n = t.nodes[ni];
        DevSlot s{};
        if (n.leaf >= 0) {
          s.is_leaf = 1;
          s.child = leaf_base + n.leaf;
        } else {
          if (n.left < 0 || n.right < 0 || n.left >= (int32_t)t.nodes.size() || n.right >= (int32_t)t.nodes.size()) { err = "dangling child"; return CRF_ERR_FORMAT; }
          const int32_t pair = (int32_t)out.slots.size();
          out.slots.emplace_back();
          out.slots.emplace_back();
          slot_of[n.left] = pair;
          slot_of[n.right] = pair + 1;
          q.push_back(n.left);
          q.push_back(n.right);
          tag(n.left);
          tag(n.right);
          const uint32_t area1 = (uint32_t)n.r1[2] * n.r1[3], area2 = (uint32_t)n.r2[2] * n.r2[3];
          if (area1 == 0 || area2 == 0) { err = "empty rectangle"; return CRF_ERR_UNSUPPORTED; }
          if (n.r1[0] + n.r1[2] > kPatch || n.r1[1] + n.r1[3] > kPatch || n.r2[0] + n.r2[2] > kPatch || n.r2[1] + n.r2[3] > kPatch) { err = "rectangle leaves the patch"; return CRF_ERR_UNSUPPORTED; }
          auto strips = [](int w, int h, uint8_t& ns, uint16_t& hs, uint16_t& hl) {
            const int rows = std::max(1, kStripArea / w);          // rows per strip so that w * rows <= 257
            const int n_ = (h + rows - 1) / rows;
            ns = (uint8_t)n_;
            hs = (uint16_t)(rows * kRowStride);
            hl = (uint16_t)((h - (n_ - 1) * rows) * kRowStride);
          };
          s.a1 = (uint16_t)(n.r1[1] * kRowStride + n.r1[0]);
          s.w1 = n.r1[2];
          strips(n.r1[2], n.r1[3], s.ns1, s.hs1, s.hl1);
          s.a2 = (uint16_t)(n.r2[1] * kRowStride + n.r2[0]);
          s.w2 = n.r2[2];
          strips(n.r2[2], n.r2[3], s.ns2, s.hs2, s.hl2);
          s.ch = n.channel;
          s.thr = n.threshold;
          s.m1 = magic_for_area(area1);
          s.m2 = magic_for_area(area2);
          s.child = pair;
        }
        out.slots[slot_of[ni]] = s;
        DevSlot16 c{};
        c.child = s.child;
        if (n.leaf >= 0) c.r1 = 1u << 26;
        else {
          c.r1 = (uint32_t)n.r1[0] | (uint32_t)n.r1[1] << 5 | (uint32_t)n.r1[2] << 10 | (uint32_t)n.r1[3] << 15 | (uint32_t)n.channel << 20;
          c.r2 = (uint32_t)n.r2[0] | (uint32_t)n.r2[1] << 5 | (uint32_t)n.r2[2] << 10 | (uint32_t)n.r2[3] << 15 | (uint32_t)(n.threshold + 256) << 20;
          c.areas = (uint32_t)n.r1[2] * n.r1[3] | ((uint32_t)n.r2[2] * n.r2[3]) << 16;
        }
        if (out.slots16.size() < out.slots.size()) out.slots16.resize(out.slots.size());
        out.slots16[slot_of[ni]] = c;
        DevSlotW w{};
        if (n.leaf >= 0) {
          w.m2 = (uint32_t)s.child;
          w.child = slot_of[ni];
          w.tw = 0x7fffu | 1u << 31;
        } else {
          w.px1 = (uint32_t)n.channel * kWinPlaneBytes + n.r1[0] * 4u;
          w.px2 = (uint32_t)n.channel * kWinPlaneBytes + n.r2[0] * 4u;
          w.yh1 = (uint32_t)n.r1[1] * kWinRowBytes | ((uint32_t)n.r1[3] * kWinRowBytes) << 16;
          w.yh2 = (uint32_t)n.r2[1] * kWinRowBytes | ((uint32_t)n.r2[3] * kWinRowBytes) << 16;
          w.m1 = s.m1; w.m2 = s.m2;
          w.child = s.child;
          w.tw = (uint32_t)(uint16_t)n.threshold | (n.r1[2] * 4u) << 16 | (n.r2[2] * 4u) << 24;
          for (const uint8_t* r : {n.r1, n.r2}) out.max_extent = std::max(out.max_extent, std::max(r[0] + r[2], r[1] + r[3]));
        }
        if (out.slotsw.size() < out.slots.size()) out.slotsw.resize(out.slots.size());
        out.slotsw[slot_of[ni]] = w;
        if (n.leaf < 0) {
          DevSlotN r{};
          r.pw1 = w.px1 | (n.r1[2] * 4u) << 18;
          r.pw2 = w.px2 | (n.r2[2] * 4u) << 18;
          r.yh1 = w.yh1; r.yh2 = w.yh2; r.m1 = s.m1; r.m2 = s.m2;
          const int32_t thr = std::min(255, std::max(-256, (int)n.threshold));   // |mean1 - mean2| <= 255: the clamp preserves the test
          r.left_thr = (int32_t)(((uint32_t)n_of[n.left] & ((1u << kSlotNChildBits) - 1)) | (uint32_t)thr << kSlotNChildBits);
          r.right = n_of[n.right];
          out.slotsn[n_of[ni]] = r;
        }
      }
    }
  }
  // the child tag of DevSlotN holds 21 bits + sign; larger forests simply do not get the form (k_traverse_win stays in use)
  if (out.slotsn.size() >= (1u << (kSlotNChildBits - 1)) || out.leaf_oid.size() >= (1u << (kSlotNChildBits - 1))) { out.slotsn.clear(); out.nroot.clear(); }
  return CRF_OK;
}

}  // namespace crf
