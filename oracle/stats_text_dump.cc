// stats_text_dump.cc -- TEST INFRASTRUCTURE: prints the bytes of the reference's own HtStatistics::print / toCsvString /
// toCsvStringHeader (ht_statistics.cc:16-79) for deterministic random tables, through oracle/_ref/libhj3d_ref.so
// (the unmodified reference templates).  Used by oracle/gen_stats_text_golden.py to write tests/golden/ht_statistics_text.json.
// A stand-alone program because the harness library carries its own libstdc++ locale state, which must not meet Python's.
//   g++ -O1 -o /tmp/stats_text_dump oracle/stats_text_dump.cc -Loracle/_ref -lhj3d_ref -Wl,-rpath,$PWD/oracle/_ref
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
extern "C" {
struct ks_t { uint32_t tb, ko, kb, hid, ro; };
struct st_t { uint64_t v[19]; };
void* ref_build(int kind, const void* tuples, uint64_t n, ks_t ks, uint64_t D);
void ref_stats(void* h, st_t* s);
uint64_t ref_stats_text(void* h, char* buf, uint64_t cap);
}
int main(int argc, char** argv) {
  if (argc < 5) return 2;
  const int kind = atoi(argv[1]); const uint64_t n = strtoull(argv[2], 0, 10), kmax = strtoull(argv[3], 0, 10), D = strtoull(argv[4], 0, 10);
  std::vector<uint32_t> B(2 * n + 2);
  uint64_t x = 88172645463325252ull;
  for (uint64_t i = 0; i < n; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; B[2 * i] = (uint32_t)i; B[2 * i + 1] = (uint32_t)(x % kmax); }
  void* t = ref_build(kind, B.data(), n, ks_t{8, 4, 4, 0, 0xFFFFFFFFu}, D);
  st_t s; ref_stats(t, &s);
  for (int i = 0; i < 19; ++i) printf("%llu%c", (unsigned long long)s.v[i], i == 18 ? '\n' : ' ');
  static char buf[8192];
  ref_stats_text(t, buf, sizeof buf);
  fputs(buf, stdout);
  return 0;
}
