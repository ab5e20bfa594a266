// Host harness for how the attention kernels deal their work items to CTAs (csrc/attn_fwd.cu, csrc/attn_bwd.cu).
// tests/test_attn_deal_host.py cuts the planning functions (fa_fwd_plan, fa_bwd_tail_plan, fa_bwd_splits) and the index
// decode at the top of fa_fwd_db_kernel / fa_bwd_kernel out of the sources VERBATIM and pastes them at the markers.
// Checked without a GPU: every (batch, head, tile) item is computed, whole items by one CTA over the whole walk, the
// items of a split last wave (or of a split query walk) by `parts` CTAs over disjoint, non-empty ranges that tile the
// walk, each with its own workspace slot.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <tuple>
#include <vector>

// what the planning functions ask the runtime: no device here, 148 SMs assumed (their own fallback)
enum { cudaDevAttrMultiProcessorCount = 16 };
static int cudaGetDevice(int* d) { *d = 0; return 1; }
static int cudaDeviceGetAttribute(int* v, int, int) { *v = 148; return 0; }

struct Idx { int x, y, z; };
struct FwdP { int kv_tiles, n_whole, parts, q_tiles, H; };
struct BwdP { int k_tiles, n_whole, tail_parts, q_splits, q_tiles, H, B; };

/*@@FA_FWD_PLAN@@*/
/*@@FA_BWD_PLANS@@*/

struct FwdItem { int item, part, j_begin, j_end, qt, h, b; };
static FwdItem fwd_decode(const FwdP& p, Idx blockIdx) {
/*@@FA_FWD_DECODE@@*/
  return {item, part, j_begin, j_end, qt, h, b};
}

struct BwdItem { int kt, h, b, split, splits, tail_slot, t0, T; };
static BwdItem bwd_decode(const BwdP& p, Idx blockIdx) {
/*@@FA_BWD_DECODE@@*/
  return {kt, h, b, split, splits, tail_slot, t0, T};
}

#define REQUIRE(c, ...) do { if (!(c)) { std::printf("FAIL " __VA_ARGS__); std::printf("\n"); return 1; } } while (0)

static int check_fwd(int B, int H, int Nq, int Nk) {
  FwdP p;
  p.q_tiles = (Nq + 127) / 128; p.kv_tiles = (Nk + 63) / 64; p.H = H;
  const int items = p.q_tiles * H * B;
  fa_fwd_plan(items, p.kv_tiles, &p.n_whole, &p.parts);
  const int n_split = items - p.n_whole, ctas = p.n_whole + n_split * p.parts;
  REQUIRE(p.n_whole >= 0 && p.n_whole <= items && p.parts >= 1 && (n_split == 0 || p.parts >= 2), "fwd plan B=%d H=%d Nq=%d Nk=%d", B, H, Nq, Nk);
  std::map<std::tuple<int, int, int>, std::vector<std::pair<int, int>>> cover;
  std::set<int> slots;
  for (int c = 0; c < ctas; ++c) {
    const FwdItem it = fwd_decode(p, {c, 0, 0});
    REQUIRE(it.qt >= 0 && it.qt < p.q_tiles && it.h >= 0 && it.h < H && it.b >= 0 && it.b < B, "fwd decode range cta %d", c);
    REQUIRE(it.j_begin < it.j_end && it.j_begin >= 0 && it.j_end <= p.kv_tiles, "fwd empty key range cta %d", c);
    if (it.part >= 0) {
      const int slot = (it.item - p.n_whole) * p.parts + it.part;
      REQUIRE(slot >= 0 && slot < n_split * p.parts && slots.insert(slot).second, "fwd workspace slot cta %d", c);
    }
    cover[{it.b, it.h, it.qt}].push_back({it.j_begin, it.j_end});
  }
  REQUIRE((int)cover.size() == items, "fwd items covered %d of %d (B=%d H=%d Nq=%d Nk=%d)", (int)cover.size(), items, B, H, Nq, Nk);
  for (auto& kv : cover) {
    auto r = kv.second;
    std::sort(r.begin(), r.end());
    int at = 0;
    for (auto& ab : r) { REQUIRE(ab.first == at, "fwd key ranges do not tile the walk"); at = ab.second; }
    REQUIRE(at == p.kv_tiles, "fwd key walk ends at %d of %d", at, p.kv_tiles);
  }
  return 0;
}

static int check_bwd(int B, int H, int Nq, int Nk, bool masked) {
  BwdP p;
  p.H = H; p.B = B; p.q_tiles = (Nq + 127) / 128;
  const int k_tiles = (Nk + 127) / 128;
  p.q_splits = fa_bwd_splits(B, H, Nq, Nk, masked);
  REQUIRE(p.q_splits >= 1 && (p.q_splits == 1 || p.q_tiles / p.q_splits >= 1), "bwd splits");
  p.k_tiles = 0; p.n_whole = 0; p.tail_parts = 1;
  Idx grid = {k_tiles, H, B * p.q_splits};
  int n_split = 0;
  if (p.q_splits == 1) {      // as in b200_fa_bwd: 1-D grid, the sparse last wave split along the query walk
    const int items = grid.x * grid.y * grid.z;
    p.k_tiles = grid.x;
    p.n_whole = items;
    fa_bwd_tail_plan(items, p.q_tiles, &p.n_whole, &p.tail_parts);
    n_split = items - p.n_whole;
    REQUIRE(p.tail_parts >= 1 && (n_split == 0 || p.tail_parts >= 2), "bwd tail plan");
    grid = {p.n_whole + n_split * p.tail_parts, 1, 1};
  }
  std::map<std::tuple<int, int, int>, std::vector<std::pair<int, int>>> cover;
  std::set<int> slots;
  for (int z = 0; z < grid.z; ++z)
    for (int y = 0; y < grid.y; ++y)
      for (int x = 0; x < grid.x; ++x) {
        const BwdItem it = bwd_decode(p, {x, y, z});
        REQUIRE(it.kt >= 0 && it.kt < k_tiles && it.h >= 0 && it.h < H && it.b >= 0 && it.b < B, "bwd decode range");
        REQUIRE(it.T >= 1 && it.t0 >= 0 && it.t0 + it.T <= p.q_tiles, "bwd empty query range (T=%d) B=%d H=%d Nq=%d Nk=%d", it.T, B, H, Nq, Nk);
        if (it.tail_slot >= 0) REQUIRE(it.tail_slot < n_split * p.tail_parts && slots.insert(it.tail_slot).second, "bwd tail slot");
        cover[{it.b, it.h, it.kt}].push_back({it.t0, it.t0 + it.T});
      }
  REQUIRE((int)cover.size() == k_tiles * H * B, "bwd items covered");
  for (auto& kv : cover) {
    auto r = kv.second;
    std::sort(r.begin(), r.end());
    int at = 0;
    for (auto& ab : r) { REQUIRE(ab.first == at, "bwd query ranges do not tile the walk"); at = ab.second; }
    REQUIRE(at == p.q_tiles, "bwd query walk ends at %d of %d", at, p.q_tiles);
  }
  return 0;
}

int main() {
  long cases = 0;
  const int Ns[] = {1, 64, 127, 128, 129, 256, 1000, 1584, 3168, 3328, 5280, 6144, 6336, 12672, 20000};
  const int Nks[] = {1, 15, 64, 255, 256, 257, 1024, 1584, 3328, 5280, 6144, 12672, 20000};
  for (int B : {1, 2, 3, 4, 8})
    for (int H : {1, 2, 4, 8, 32})
      for (int Nq : Ns)
        for (int Nk : Nks) {
          if (Nk >= 512 && check_fwd(B, H, Nq, Nk)) return 1;       // fa_fwd_db_kernel serves 512 keys and more
          if (check_bwd(B, H, Nq, Nk, false) || check_bwd(B, H, Nq, Nk, true)) return 1;
          cases += 3;
        }
  std::printf("OK %ld cases\n", cases);
  return 0;
}
