// Host harness for the persistent tile walk of csrc/gemm.cu (GemmParams / tile_coords / GemmUnit / GemmWalk).
// tests/test_gemm_walk_host.py cuts those definitions out of gemm.cu VERBATIM, pastes them where the marker below
// stands, compiles with g++ and runs: the scheduling logic that decides which CTA computes which k blocks of which tile
// is plain integer code, so it can be checked exhaustively without a GPU.

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <tuple>
#include <vector>
#define __device__
#define __forceinline__ inline
typedef uint16_t bf16;
using std::min;
/*@@GEMM_WALK_SOURCE@@*/
static int fail(const char* what, const GemmParams& p, int P) {
  std::printf("FAIL %s: m_tiles=%d n_tiles=%d group_m=%d kb1=%d kb2=%d splits=%d groups=%d sk_tiles=%d P=%d\n", what,
              p.m_tiles, p.n_tiles, p.group_m, p.kb1, p.kb2, p.splits, p.groups, p.sk_tiles, P);
  return 1;
}

// every k block of every (tile, split / group) is visited exactly once over all CTAs; stream-K tiles have exactly one
// owner (role 0 or 2) and at most four helpers (role 1) whose ranges follow the owner's; tile_coords is a bijection
static int check(const GemmParams& p, int P) {
  const int tiles = p.m_tiles * p.n_tiles, kb_all = p.kb1 + p.kb2;
  std::vector<int> seen(tiles, 0);
  for (int t = 0; t < tiles; ++t) {
    int tm, tn;
    tile_coords(p, t, tm, tn);
    if (tm < 0 || tm >= p.m_tiles || tn < 0 || tn >= p.n_tiles) return fail("tile_coords range", p, P);
    seen[tm * p.n_tiles + tn]++;
  }
  for (int t = 0; t < tiles; ++t)
    if (seen[t] != 1) return fail("tile_coords bijection", p, P);
  const int zs = p.splits * p.groups;   // a work item's z is the split-K slice OR the batch group (they exclude each other)
  const int gs = p.groups;              // coverage is counted per (tile, group, k block): the splits of a tile share it
  std::vector<int> cover((size_t)tiles * gs * kb_all, 0);
  std::vector<int> owners(tiles, 0), helpers(tiles, 0);
  for (int w = 0; w < P; ++w) {
    GemmWalk walk(p, w, P);
    GemmUnit u;
    int guard = 0;
    bool first_unit = true;
    while (walk.next(u)) {
      if (++guard > 1 << 20) return fail("walk does not terminate", p, P);
      if (u.tile < 0 || u.tile >= tiles || u.z < 0 || u.z >= zs) return fail("unit range", p, P);
      if (u.kb_begin < 0 || u.kb_end > kb_all || u.kb_begin >= u.kb_end) return fail("empty or out-of-range k range", p, P);
      if (u.role != 0) {
        if (u.tile >= p.sk_tiles) return fail("partial unit outside the stream-K region", p, P);
        if (u.role == 1) {
          helpers[u.tile]++;
          if (!first_unit) return fail("a dumped partial is not the CTA's first unit", p, P);   // the kernel relies on it
        }
        if (u.role == 2) owners[u.tile]++;
        if ((u.role == 1) != (u.kb_begin > 0)) return fail("role / first-k-block rule", p, P);
      } else if (u.tile < p.sk_tiles) {
        owners[u.tile]++;
      }
      const int g = p.groups > 1 ? u.z : 0;
      for (int kb = u.kb_begin; kb < u.kb_end; ++kb) cover[((size_t)u.tile * gs + g) * kb_all + kb]++;
      first_unit = false;
    }
  }
  for (size_t i = 0; i < cover.size(); ++i)
    if (cover[i] != 1) return fail(cover[i] == 0 ? "k block never visited" : "k block visited twice", p, P);
  for (int t = 0; t < p.sk_tiles; ++t) {
    if (owners[t] != 1) return fail("stream-K tile without exactly one owner", p, P);
    if (helpers[t] > 4) return fail("stream-K tile shared by more than five CTAs", p, P);
  }
  return 0;
}

int main() {
  long cases = 0;
  GemmParams p = {};
  // whole-tile walks, grouped rasterisation, split-K, batched groups
  for (int m = 1; m <= 13; ++m)
    for (int n = 1; n <= 9; ++n)
      for (int gm : {1, 2, 3, 5, m})
        for (int kb : {1, 3, 32, 97})
          for (int sp : {1, 2, 3, 7})
            for (int groups : {1, 4})
              for (int P : {1, 2, 7, 74, 148}) {
                if (sp > 1 && groups > 1) continue;
                int s = std::min(sp, kb);
                while (s > 1 && (s - 1) * ((kb + s - 1) / s) >= kb) --s;     // the host's rule: every split owns a k block
                p.m_tiles = m; p.n_tiles = n; p.group_m = std::min(gm, m); p.kb1 = kb - kb / 4; p.kb2 = kb / 4;
                p.splits = s; p.groups = groups; p.sk_tiles = 0;
                if (check(p, P)) return 1;
                ++cases;
              }
  // stream-K over the last partial wave, with the host's admission rule (gemm_impl): T > P, R*3 >= P, R*10 <= 9*P, kb >= 8
  for (int P : {4, 7, 37, 74, 148})
    for (int T = P + 1; T <= 4 * P + 3; ++T)
      for (int kb : {8, 9, 32, 33, 96, 128, 129}) {
        const int R = T % P;
        if (!(R * 3 >= P && R * 10 <= 9 * P)) continue;
        for (int n : {1, 2, 8}) {
          if (T % n) continue;
          p.m_tiles = T / n; p.n_tiles = n; p.group_m = p.m_tiles; p.kb1 = kb; p.kb2 = 0; p.splits = 1; p.groups = 1;
          p.sk_tiles = R;
          if (check(p, P)) return 1;
          ++cases;
        }
      }
  std::printf("OK %ld cases\n", cases);
  return 0;
}
