// plum_b200 — moving a std::mt19937 between the host and the device-resident chain (pg_chain_set_rng / pg_chain_get_rng).
//
// The state of record is what operator<< of std::mt19937 prints: the 624 state words and the position.  The portable
// way in and out of the engine is that text form (slow_*: ~60 us per generator).  libstdc++ stores exactly those numbers
// as `uint_fast32_t _M_x[624]; size_t _M_p;`, so when a one-off self-test confirms that layout against the text form the
// bytes are copied instead (~1 us) — it matters when one host thread hands hundreds of generators back and forth per batch.
#ifndef PLUM_B200_HOST_MT_STATE_H_
#define PLUM_B200_HOST_MT_STATE_H_

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <random>
#include <sstream>

namespace plum_mt {

inline void slow_export(const std::mt19937& g, uint32_t* state, int* pos) {
  std::stringstream ss;
  ss << g;
  for (int i = 0; i < 624; i++) { unsigned long v; ss >> v; state[i] = (uint32_t)v; }
  unsigned long p; ss >> p;
  *pos = (int)p;
}
inline void slow_import(std::mt19937& g, const uint32_t* state, int pos) {
  std::stringstream ss;
  for (int i = 0; i < 624; i++) ss << state[i] << ' ';
  ss << pos;
  ss >> g;
}

struct Layout { std::uint_fast32_t x[624]; std::size_t p; };

inline bool fast_ok() {
  static const bool ok = [] {
    if (sizeof(std::mt19937) != sizeof(Layout)) return false;
    for (unsigned seed = 1; seed <= 3; seed++) {
      std::mt19937 g(seed * 2654435761u);
      for (unsigned i = 0; i < 300 * seed; i++) (void)g();
      uint32_t a[624]; int pa = 0;
      slow_export(g, a, &pa);
      Layout L;
      std::memcpy(&L, &g, sizeof(L));
      if ((int)L.p != pa) return false;
      for (int i = 0; i < 624; i++) if ((uint32_t)L.x[i] != a[i] || (L.x[i] >> 32) != 0) return false;
      // and back: a generator rebuilt from the bytes continues the same stream
      std::mt19937 h;
      std::memcpy(&h, &L, sizeof(L));
      std::mt19937 g2 = g;
      for (int i = 0; i < 700; i++) if (h() != g2()) return false;
    }
    return true;
  }();
  return ok;
}

inline void export_state(const std::mt19937& g, uint32_t* state, int* pos) {
  if (!fast_ok()) { slow_export(g, state, pos); return; }
  Layout L;
  std::memcpy(&L, &g, sizeof(L));
  for (int i = 0; i < 624; i++) state[i] = (uint32_t)L.x[i];
  *pos = (int)L.p;
}
inline void import_state(std::mt19937& g, const uint32_t* state, int pos) {
  if (!fast_ok()) { slow_import(g, state, pos); return; }
  Layout L;
  for (int i = 0; i < 624; i++) L.x[i] = state[i];
  L.p = (std::size_t)pos;
  std::memcpy(&g, &L, sizeof(L));
}

}  // namespace plum_mt

#endif
