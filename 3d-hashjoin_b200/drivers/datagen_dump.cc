// prints the generated relations of the experiment drivers (used by tests/test_datagen.py)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "hj3d/datagen.hh"
int main(int argc, char** argv) {
  if (argc >= 6 && !strcmp(argv[1], "exp1")) {
    auto d = hj3d::gen::experiment1(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]) != 0, atoi(argv[5]));
    printf("%zu\n", d.numDvSa);
    for (auto v : d.Rk) printf("%u ", v);
    printf("\n");
    for (auto v : d.Sa) printf("%u ", v);
    printf("\n");
    return 0;
  }
  if (argc >= 7 && !strcmp(argv[1], "exp4")) {
    auto d = hj3d::gen::experiment4(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]));
    for (auto v : d.Sa) printf("%u ", v);
    printf("\n");
    for (auto v : d.Ta) printf("%u ", v);
    printf("\n");
    return 0;
  }
  return 2;
}
