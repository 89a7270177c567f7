// Device image of the forests: flat 32-byte slot records (one L2 sector each), leaf tables with
// the reference's per-leaf arithmetic folded in at load time.
//
// Replaces the pointer-linked TreeNode<S> (reference include/TreeNode.hpp:138-146: embedded Leaf +
// Split + two heap pointers) for Tree<S>::evaluateMT (include/Tree.hpp:174-191).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "model.h"

namespace crf {

constexpr int kRowStride = 128;  // elements per integral row (CRF_ROW_STRIDE)
// Element type of the integral planes.  uint32_t: exact sums, 4 corner loads per rectangle.  uint16_t: sums modulo 2^16
// with rectangles cut into strips of area <= 257 (see DevSlot) — built, parity-green and MEASURED 1.7x SLOWER on B200
// (the traversal is bound by L1 wavefronts, i.e. by the number of load instructions x lines, not by bytes: the 16 % of
// rectangles that need a second strip add loads and a divergent path, and half-width rows save no wavefronts).
typedef uint32_t stack_t;
constexpr bool kStack16 = sizeof(stack_t) == 2;
constexpr int kStripArea = kStack16 ? 257 : (1 << 20);   // 255 * 257 < 2^16; one strip always with 32-bit sums
constexpr uint32_t kSumMask = kStack16 ? 0xffffu : 0xffffffffu;
constexpr int kPatch = 31;       // ForestParam::getPatchSize() for face_size 125 (include/Constants.hpp:26-30)
constexpr int kHalfPatch = 15;   // patch_size / 2 (src/face_utils.cpp:281-282)
constexpr int kParts = 10;
constexpr int kMaxList = 128;    // composed-forest capacity per face (pathological compositions, see engine)

// One slot = one tree node.  Children of an internal node are adjacent: left = child, right = child + 1
// (left is serialised first: include/TreeNode.hpp:161-162).  Slots of a tree are laid out breadth-first
// so the hot top levels share cache lines.
// With 16-bit planes (kStack16) a rectangle sum D - B - C + A is exact modulo 2^16, hence exact outright for areas
// <= 257 (u8 pixels); larger rectangles are cut into `ns` horizontal strips of at most floor(257 / w) rows, each
// summed exactly.  With the default 32-bit planes ns is always 1.
struct alignas(32) DevSlot {
  uint16_t a1;       // rect1: y*kRowStride + x (element offset of the top-left corner)
  uint8_t w1, ns1;   // width; number of strips (1..4)
  uint16_t hs1;      // rows of a full strip * kRowStride
  uint16_t hl1;      // rows of the last strip * kRowStride
  uint16_t a2;       // rect2
  uint8_t w2, ns2;
  uint16_t hs2, hl2;
  uint8_t ch;        // feature channel == plane index
  uint8_t is_leaf;
  int16_t thr;       // ThresholdSplit::threshold clamped to [-256, 255] (|mean1 - mean2| <= 255)
  uint32_t m1, m2;   // floor(2^31 / area) + 1: mean = umulhi(sum << 1, m) == sum / area for sum <= 255*area
  int32_t child;     // internal: slot of the left child; leaf: forest-global leaf index
};
static_assert(sizeof(DevSlot) == 32, "slot must be one 32-byte sector");

// Compact 16-byte form of the same node (two per 32-byte sector, so both children of a node share a sector): the
// reciprocals are dropped and the two integer means are recomputed with a float reciprocal + exact fix-up.
//   r1 = x1 | y1<<5 | w1<<10 | h1<<15 | ch<<20 | leaf<<26      r2 = x2 | y2<<5 | w2<<10 | h2<<15 | (thr+256)<<20
//   child = slot of the left child (leaf: forest-global leaf index)     areas = area1 | area2<<16
struct alignas(16) DevSlot16 { uint32_t r1, r2; int32_t child; uint32_t areas; };
static_assert(sizeof(DevSlot16) == 16, "compact slot is 16 bytes");

// Window form of the same node for the dense (stride-1) traversal that gathers from SHARED memory (k_traverse_win):
// a CTA keeps the 38 x 38-sample window of all planes that an 8 x 8 tile of patches can touch (rectangles of the model
// end at offset <= kWinExtent inside the patch) as [plane][ring row][kWinCols] u32, rows addressed modulo kWinRows so that
// stepping the tile down by 8 patches re-loads only 8 rows.  Everything a node test needs is pre-multiplied into byte
// offsets of that layout; leaves loop onto themselves (child = own slot, zero rectangles, threshold 32767) so a finished
// walk can keep executing the same instruction stream as its neighbour.
constexpr int kWinTile = 8;                    // patches per tile side
constexpr int kWinExtent = 30;                 // largest x + w / y + h the window covers
constexpr int kWinRows = kWinTile + kWinExtent;     // 38 ring rows
constexpr int kWinCols = 40;                   // 38 columns used; 40 = 160-byte rows, pitch == 8 (mod 32 banks)
constexpr int kWinRowBytes = kWinCols * 4;
constexpr int kWinPlaneBytes = kWinRows * kWinRowBytes;   // 6080
struct alignas(32) DevSlotW {
  uint32_t px1, px2;   // ch * kWinPlaneBytes + x * 4
  uint32_t yh1, yh2;   // y * kWinRowBytes | (h * kWinRowBytes) << 16
  uint32_t m1, m2;     // magic_for_area; leaf: m2 = forest-global leaf index
  int32_t child;       // internal: slot of the left child; leaf: own slot
  uint32_t tw;         // (uint16)thr | (w1 * 4) << 16 | (w2 * 4) << 24 | leaf << 31
};
static_assert(sizeof(DevSlotW) == 32, "window slot is one 32-byte sector");

// Internal-nodes-only window form (k_traverse_win2).  Leaves have no record at all: a child is either the index of another
// DevSlotN (>= 0) or ~(forest-global leaf index) (< 0), so a walk knows that it has arrived — and where — from its PARENT's
// record.  That (a) drops the last record fetch of every walk and halves the array (every second node of a full binary tree is a
// leaf), and (b) takes the freshly fetched record out of the loop condition, which lets the two walks of a lane run half an
// iteration apart: the record fetch of one is in flight while the other does its node test (software pipelining inside the warp).
// Records of a tree are breadth-first, `nroot_of_slot` maps a DevSlot root index to the DevSlotN index (or ~leaf for a one-leaf tree).
struct alignas(32) DevSlotN {
  uint32_t pw1, pw2;   // ch * kWinPlaneBytes + x * 4 (18 bits) | (w * 4) << 18
  uint32_t yh1, yh2;   // y * kWinRowBytes | (h * kWinRowBytes) << 16
  uint32_t m1, m2;     // magic_for_area
  int32_t left_thr;    // left child in bits 0..21 (two's complement), threshold in bits 22..31
  int32_t right;       // right child
};
static_assert(sizeof(DevSlotN) == 32, "internal-node slot is one 32-byte sector");
constexpr int kSlotNChildBits = 22;   // |child| < 2^21: forests of up to 2 M internal nodes / leaves

// MPLeaf (include/MPSample.hpp:137-159) with the vote predicate of src/face_utils.cpp:285-290 folded
// into `mask` (bit i: part i votes) for the options of the context.
struct alignas(16) DevMpLeaf {   // 48 bytes: three 128-bit loads
  int16_t off[kParts][2];
  float weight;  // forground
  uint32_t pad;
};
static_assert(sizeof(DevMpLeaf) == 48, "leaf record is three 16-byte words");

struct PackedForest {
  std::vector<DevSlot> slots;
  std::vector<DevSlot16> slots16;    // same slots, compact form
  std::vector<DevSlotW> slotsw;      // same slots, window form (valid when max_extent <= kWinExtent)
  std::vector<DevSlotN> slotsn;      // internal nodes only, children tagged (k_traverse_win2); empty when the forest is too large for the tag
  std::vector<int32_t> nroot;        // per tree (parallel to `roots`): DevSlotN index of the root, or ~leaf for a one-leaf tree
  int max_extent = 0;                // largest x + w or y + h over all rectangles
  std::vector<int32_t> roots;        // slot of the root of tree t (forest-major for the jungle)
  std::vector<int32_t> forest_base;  // jungle: first tree of pose forest f in `roots`
  std::vector<int32_t> forest_ntrees;
  std::vector<int32_t> leaf_base;    // first forest-global leaf index of tree t
  std::vector<int32_t> leaf_oid;     // Boost object id (pre-order node index) of leaf l inside its tree
  // head pose
  std::vector<float> hp_m;           // expected label of the leaf (src/face_utils.cpp:224-228), -1 when fg <= min_fg
  // multi part
  std::vector<DevMpLeaf> mp_leaf;
  std::vector<uint16_t> mp_mask;
  int max_depth = 0;
};

struct PackOptions {
  float hp_min_foreground = 0.5f;
  int ffd_min_samples = 2;
  float ffd_min_foreground = 0.5f;
  float ffd_min_pf = 0.25f;
  float ffd_max_variance = 25.f;
};

int pack_forests(const std::vector<const FlatForest*>& forests, ForestKind kind, const PackOptions& opt, PackedForest& out, std::string& err);

// Exact division helper shared by host checks and the device code.
inline uint32_t magic_for_area(uint32_t area) { return (uint32_t)((1ull << 31) / area) + 1u; }

}  // namespace crf
