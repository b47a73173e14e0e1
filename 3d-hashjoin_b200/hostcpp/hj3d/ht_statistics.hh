// hj3d/ht_statistics.hh -- HtStatistics with the reference's public fields and output formats
// (ht_statistics.hh:18-54, ht_statistics.cc:16-103), filled from hj3d_table_stats.
#pragma once

#include <cstddef>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/hj3d.h"

// the subset of util/aggregate.hh:27-68 the statistics expose
template <class Num>
class Aggregate {
  public:
    Aggregate() { init(); }
    void init() { _min = std::numeric_limits<Num>::max(); _max = std::numeric_limits<Num>::min(); _sum = _sumsq = _count = 0; }
    void set(Num mn, Num mx, Num sum, Num sumsq, Num count) { _min = mn; _max = mx; _sum = sum; _sumsq = sumsq; _count = count; }
    Num    count() const { return _count; }
    Num    min()   const { return _min; }
    Num    max()   const { return _max; }
    Num    sum()   const { return _sum; }
    Num    sumsq() const { return _sumsq; }
    double avg()   const { return ((double)sum() / (double)count()); }
  private:
    Num _min, _max, _sum, _sumsq, _count;
};

struct HtBucketStatistics { size_t _bucketIndex, _numEntries, _chainLen; };

struct HtStatistics {
  size_t _numBuckets = 0, _numEmptyBuckets = 0, _numEntries = 0, _numDistinctKeys = 0;
  Aggregate<size_t> _collisionChainLen, _collisionChainLenNonempty;
  Aggregate<size_t> _numDistinctKeysPerBucket, _numDistinctKeysPerNonemptyBucket;   // never filled by the reference either
  std::vector<HtBucketStatistics> _bucketStats;

  double numEntriesPerKey() const { return (_numEntries + 0.0) / _numDistinctKeys; }
  double fracEmptyBuckets() const { return (_numEmptyBuckets + 0.0) / _numBuckets; }

  static HtStatistics from(const hj3d_stats& s) {
    HtStatistics h;
    h._numBuckets = s.num_buckets; h._numEmptyBuckets = s.num_empty; h._numEntries = s.num_entries;
    h._numDistinctKeys = s.num_distinct_keys;
    h._collisionChainLen.set(s.cc_min, s.cc_max, s.cc_sum, s.cc_sumsq, s.cc_count);
    h._collisionChainLenNonempty.set(s.ccne_min, s.ccne_max, s.ccne_sum, s.ccne_sumsq, s.ccne_count);
    return h;
  }

  void print(std::ostream& os = std::cout) const {
    os << "#buckets                  = " << _numBuckets << "\n";
    os << "#empty buckets            = " << _numEmptyBuckets << "\n";
    os << "#entries                  = " << _numEntries << "\n";
    os << "#distinct keys            = " << _numDistinctKeys << "\n";
    os << "cc length:                  " << _collisionChainLen.avg() << " | " << _collisionChainLen.min() << " | "
       << _collisionChainLen.max() << "\n";
    os << "cc length nonempty:         " << _collisionChainLenNonempty.avg() << " | " << _collisionChainLenNonempty.min()
       << " | " << _collisionChainLenNonempty.max() << "\n";
  }
  std::string toCsvString() const {
    std::stringstream res;
    res << _numBuckets << ";" << _numEmptyBuckets << ";" << _numEntries << ";" << _numDistinctKeys << ";";
    res << _collisionChainLen.avg() << ";" << _collisionChainLen.min() << ";" << _collisionChainLen.max() << ";";
    res << _collisionChainLenNonempty.avg() << ";" << _collisionChainLenNonempty.min() << ";"
        << _collisionChainLenNonempty.max() << ";";
    return res.str();
  }
  static std::string toCsvStringHeader() {
    return "#buckets;#empty_buckets;#entries;#distinct_keys;#e/b_avg;#e/b_min;#e/b_max;#e/neb_avg;#e/neb_min;#e/neb_max;";
  }
  void printCsv(std::ostream& os = std::cout) const { os << toCsvString(); }
  static void printCsvHeader(std::ostream& os = std::cout) { os << toCsvStringHeader(); }
};
