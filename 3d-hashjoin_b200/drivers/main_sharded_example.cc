// main_sharded_example -- the key/foreign-key join of main_experiment1 (main_experiment1.cc:624-1285, plans Csr and Nsr)
// sharded over G ranks held by ONE process, written against the C ABI only (include/hj3d.h): what a C++ driver that owns
// all GPUs of a box does.  The relations live in HOST memory, as in the reference (RelationRS = std::vector<tuple_t>,
// algebra.hh:98-106); every rank streams its slice through the exchange (hj3d_exchange_begin_host: chunked upload,
// partition level 1 / peer stores per chunk), builds the shard of the table it owns from what it received and probes it.
// With --skew (Zipf foreign keys) the slices are uploaded first and the probe side is exchanged with hot-key probe
// replication (hj3d_exchange_hot_sample / HJ3D_XCHG_HOT / hj3d_parts_hot_answers).
// The merged counters and HtStatistics must equal those of the unsharded join (hj3d_join_host on one GPU).
//
//   main_sharded_example.out [-R log2] [-S log2] [-g ranks] [--skew] [-p Csr|Nsr]
// Ranks are spread round-robin over the visible GPUs (several ranks may share one: the data plane is the same).
// Prints one CSV line (header first) and exits 0 only if sharded == unsharded.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hj3d.h"
#include "hj3d/datagen.hh"

namespace {

struct tuple3 { uint32_t k, a, b; };   // tuple_uint32_3_t of main_experiment1.cc:86

#define CHECK(call)                                                                        \
  do {                                                                                     \
    const int _rc = (call);                                                                \
    if (_rc < 0) { fprintf(stderr, "hj3d: %s failed: %s\n", #call, hj3d_last_error()); exit(3); } \
  } while (0)

bool same(const hj3d_stats& a, const hj3d_stats& b) { return memcmp(&a, &b, sizeof a) == 0; }

}  // namespace

int main(int argc, char** argv) {
  uint32_t log2R = 16, log2S = 19; int G = 2; bool skew = false; std::string plan = "Csr";
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "-R" && i + 1 < argc) log2R = (uint32_t)atoi(argv[++i]);
    else if (a == "-S" && i + 1 < argc) log2S = (uint32_t)atoi(argv[++i]);
    else if (a == "-g" && i + 1 < argc) G = atoi(argv[++i]);
    else if (a == "-p" && i + 1 < argc) plan = argv[++i];
    else if (a == "--skew") skew = true;
    else if (a == "--no-skew") skew = false;
    else { fprintf(stderr, "usage: %s [-R log2] [-S log2] [-g ranks] [--skew] [-p Csr|Nsr]\n", argv[0]); return 2; }
  }
  if (G < 1 || G > 16 || (plan != "Csr" && plan != "Nsr")) { fprintf(stderr, "bad -g / -p\n"); return 2; }
  const int mode = plan == "Csr" ? 1 : 3;                      // chaining + IsBuildKeyUnique | nested + unnest
  const int kind = plan == "Csr" ? HJ3D_CHAINING : HJ3D_NESTED;

  // the reference's generator (Experiment1::init), relations as row stores in host memory
  const auto d = hj3d::gen::experiment1(log2R, log2S, skew, 0);
  const uint64_t nR = d.Rk.size(), nS = d.Sa.size(), D = nR;   // buckets = |R| / b, b = 1 (main_experiment1.cc:651)
  std::vector<tuple3> R(nR), S(nS);
  for (uint64_t i = 0; i < nR; ++i) R[i] = {d.Rk[i], 0, 0};
  for (uint64_t i = 0; i < nS; ++i) S[i] = {(uint32_t)i, d.Sa[i], 0};
  const hj3d_keyspec ksR{12, 0, 4, HJ3D_HASH_MURMUR32, HJ3D_NO_ROWID}, ksS{12, 4, 4, HJ3D_HASH_MURMUR32, HJ3D_NO_ROWID};

  // contexts: rank r on GPU r % (number of GPUs)
  std::vector<hj3d_ctx*> ctx(G, nullptr);
  int n_dev = 0;
  for (; n_dev < G; ++n_dev) { hj3d_ctx* probe = nullptr; if (hj3d_ctx_create(n_dev, &probe) < 0) break; ctx[n_dev] = probe; }
  if (n_dev == 0) { fprintf(stderr, "hj3d: %s\n", hj3d_last_error()); return 3; }   // no CPU fallback
  for (int r = n_dev; r < G; ++r) CHECK(hj3d_ctx_create(r % n_dev, &ctx[r]));

  // ---- unsharded: one GPU, host buffers in, counters out
  hj3d_counters c1{}, u1{}; hj3d_stats st1{};
  CHECK(hj3d_join_host(ctx[0], mode, R.data(), nR, ksR, D, S.data(), nS, ksS, HJ3D_F_CHECKSUM, nullptr, 0, &c1, &u1, &st1));
  const hj3d_counters& res1 = mode == 3 ? u1 : c1;

  // ---- sharded: G ranks in this process
  std::vector<hj3d_comm*> cm(G, nullptr);
  CHECK(hj3d_comm_create_local(ctx.data(), G, cm.data()));
  std::vector<hj3d_table*> tab(G, nullptr);
  for (int r = 0; r < G; ++r) {
    CHECK(hj3d_comm_set_option(cm[r], HJ3D_XOPT_MIN_RANGE_WIDTH, 256));
    CHECK(hj3d_comm_reserve(cm[r], 0, nR * 2 / G + 65536, 4));
    CHECK(hj3d_comm_reserve(cm[r], 1, nS * (skew ? 3 : 2) / G + 65536, 4));
  }
  for (int r = 0; r < G; ++r) {
    uint64_t lo = 0, hi = 0;
    CHECK(hj3d_comm_shard(cm[r], D, &lo, &hi));
    CHECK(hj3d_table_create_shard(ctx[r], kind, D, lo, hi, &tab[r]));
  }
  auto first = [&](uint64_t n, int r) { return n * (uint64_t)r / (uint64_t)G; };
  std::vector<hj3d_parts*> pR(G, nullptr), pS(G, nullptr);
  std::vector<void*> dR(G, nullptr), dS(G, nullptr);
  // begin for every rank, then end for every rank (single process: events instead of a collective)
  for (int r = 0; r < G; ++r) {
    const uint64_t b0 = first(nR, r), b1 = first(nR, r + 1), p0 = first(nS, r), p1 = first(nS, r + 1);
    if (!skew) {   // host slices streamed through the exchange
      CHECK(hj3d_exchange_begin_host(cm[r], 0, R.data() + b0, b1 - b0, ksR, D, (uint32_t)b0, 0, nullptr));
      CHECK(hj3d_exchange_begin_host(cm[r], 1, S.data() + p0, p1 - p0, ksS, D, (uint32_t)p0, 0, nullptr));
    } else {       // --skew: device-resident slices; the probe side's hottest keys stay local (HJ3D_XCHG_HOT)
      CHECK(hj3d_mem_alloc(ctx[r], (b1 - b0) * sizeof(tuple3) + 16, &dR[r]));
      CHECK(hj3d_mem_alloc(ctx[r], (p1 - p0) * sizeof(tuple3) + 16, &dS[r]));
      CHECK(hj3d_memcpy_h2d(ctx[r], dR[r], R.data() + b0, (b1 - b0) * sizeof(tuple3)));
      CHECK(hj3d_memcpy_h2d(ctx[r], dS[r], S.data() + p0, (p1 - p0) * sizeof(tuple3)));
      CHECK(hj3d_exchange_hot_sample(cm[r], 1, dS[r], p1 - p0, ksS));
    }
  }
  if (skew)
    for (int r = 0; r < G; ++r) {
      const uint64_t b0 = first(nR, r), b1 = first(nR, r + 1), p0 = first(nS, r), p1 = first(nS, r + 1);
      CHECK(hj3d_exchange_begin(cm[r], 0, dR[r], b1 - b0, ksR, D, (uint32_t)b0, 0));
      CHECK(hj3d_exchange_begin(cm[r], 1, dS[r], p1 - p0, ksS, D, (uint32_t)p0, HJ3D_XCHG_HOT));
    }
  uint64_t n_recv = 0, n_hot = 0;
  for (int r = 0; r < G; ++r) {
    int rc = hj3d_exchange_end(cm[r], 0, dR[r], (uint32_t)first(nR, r), nR, &pR[r]);
    if (rc < 0) { fprintf(stderr, "hj3d: %s\n", hj3d_last_error()); return 3; }
    int rc2 = hj3d_exchange_end(cm[r], 1, dS[r], (uint32_t)first(nS, r), nS, &pS[r]);
    if (rc2 < 0) { fprintf(stderr, "hj3d: %s\n", hj3d_last_error()); return 3; }
    if (rc == HJ3D_OVERFLOW || rc2 == HJ3D_OVERFLOW) { fprintf(stderr, "exchange regions overflowed: reserve more\n"); return 4; }
    uint64_t got = 0, hot = 0;
    CHECK(hj3d_parts_info(pR[r], &got, nullptr, nullptr, nullptr, nullptr));
    CHECK(hj3d_parts_hot(pS[r], &hot));
    n_recv += got; n_hot += hot;
    CHECK(hj3d_table_build_parts(ctx[r], tab[r], pR[r]));
  }
  if (skew)          // every rank answers the hot keys from its shard; the sum is what the owners answered
    for (int r = 0; r < G; ++r) CHECK(hj3d_parts_hot_answers(ctx[r], tab[r], pS[r], mode));
  hj3d_counters c{}, u{};
  std::vector<hj3d_stats> st(G);
  for (int r = 0; r < G; ++r) {
    hj3d_counters cr{}, ur{};
    CHECK(hj3d_probe_parts(ctx[r], tab[r], pS[r], mode, HJ3D_F_CHECKSUM, nullptr, 0, &cr, &ur));
    const hj3d_counters& rr = mode == 3 ? ur : cr;
    c.matches += cr.matches; c.num_cmps += cr.num_cmps;
    u.out_tuples += rr.out_tuples; u.checksum_sum += rr.checksum_sum; u.checksum_xor ^= rr.checksum_xor;
    CHECK(hj3d_table_stats(ctx[r], tab[r], &st[r]));
    hj3d_parts_destroy(pR[r]); hj3d_parts_destroy(pS[r]);
    if (dR[r]) hj3d_mem_free(ctx[r], dR[r]);
    if (dS[r]) hj3d_mem_free(ctx[r], dS[r]);
  }
  hj3d_stats merged{};
  CHECK(hj3d_stats_merge(st.data(), (uint32_t)G, &merged));

  const bool ok = n_recv == nR && c.matches == c1.matches && c.num_cmps == c1.num_cmps && u.out_tuples == res1.out_tuples &&
                  u.checksum_sum == res1.checksum_sum && u.checksum_xor == res1.checksum_xor && same(merged, st1);
  printf("log2CardR;log2CardS;skew;plan;ranks;gpus;c_htProbe;c_htProbeCmp;c_top;checksum_sum;checksum_xor;ht_buckets;ht_empty;hot_tuples_kept_local;sharded_equals_unsharded\n");
  printf("%u;%u;%d;%s;%d;%d;%llu;%llu;%llu;%llu;%llu;%llu;%llu;%llu;%s\n", log2R, log2S, (int)skew, plan.c_str(), G, n_dev,
         (unsigned long long)c.matches, (unsigned long long)c.num_cmps, (unsigned long long)u.out_tuples,
         (unsigned long long)u.checksum_sum, (unsigned long long)u.checksum_xor, (unsigned long long)merged.num_buckets,
         (unsigned long long)merged.num_empty, (unsigned long long)n_hot, ok ? "yes" : "NO");
  for (int r = 0; r < G; ++r) { hj3d_table_destroy(ctx[r], tab[r]); hj3d_comm_destroy(cm[r]); }
  for (int r = 0; r < G; ++r) hj3d_ctx_destroy(ctx[r]);
  return ok ? 0 : 1;
}
