// prints hj3d's HtStatistics (hostcpp/hj3d/ht_statistics.hh) text output for the 14 numbers given on the command line
#include <cstdlib>
#include <iostream>
#include "hj3d/ht_statistics.hh"
int main(int argc, char** argv) {
  if (argc < 15) return 2;
  hj3d_stats s{};
  uint64_t* f = &s.num_buckets;
  for (int i = 0; i < 14; ++i) f[i] = strtoull(argv[1 + i], nullptr, 10);
  const HtStatistics h = HtStatistics::from(s);
  h.print(std::cout);
  std::cout << '\x1e' << h.toCsvString() << '\x1e' << HtStatistics::toCsvStringHeader();
  return 0;
}
