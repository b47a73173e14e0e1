// hj3d/ht.hh -- host facades of the two hash tables with the reference's public surface
// (HtChaining1: ht_chaining.hh:38-158, HtNested1: ht_nested.hh:71-251); the table itself lives on the device
// behind hj3d_table (include/hj3d.h).
//
// insert() only records the tuple: the device table is (re)built from everything inserted since the last
// clear() when the build operator finishes its strand (seal()), which is when the reference's table is
// complete too.  Row id = insertion rank, so every order-dependent counter of the reference is reproduced.
#pragma once

#include <cstdint>
#include <vector>

#include "concepts.hh"
#include "ht_statistics.hh"
#include "runtime.hh"

namespace hj3d::detail {

// Tuples pushed into an operator: a contiguous slab of the relation's vector (the common case, no copy of
// pointers) or an arbitrary sequence of tuple pointers (behind a selection).
template <class T>
class TupleSeq {
  public:
    void clear() { _first = nullptr; _n = 0; _ptrs.clear(); _contig = true; }
    void push(T* t) {
      if (_contig) {
        if (_n == 0) { _first = t; _n = 1; return; }
        if (t == _first + _n) { ++_n; return; }
        _ptrs.reserve(_n + 1);                      // first gap: switch to explicit pointers
        for (size_t i = 0; i < _n; ++i) _ptrs.push_back(_first + i);
        _contig = false;
      }
      _ptrs.push_back(t); ++_n;
    }
    void push_bulk(T* first, size_t n) {
      if (n == 0) return;
      if (_contig && (_n == 0 || first == _first + _n)) { if (_n == 0) _first = first; _n += n; return; }
      for (size_t i = 0; i < n; ++i) push(first + i);
    }
    size_t size() const { return _n; }
    T* at(size_t i) const { return _contig ? _first + i : _ptrs[i]; }
    // copy the tuples to the device (gathering them into a staging vector first if they are not contiguous)
    void* upload(DevBuf& buf, std::vector<std::remove_const_t<T>>& staging) const {
      using V = std::remove_const_t<T>;
      // Runtime::cache_uploads: the caller promises that a relation does not change between runs (the drivers' repeat
      // loop, util/measure_helpers.hh:15-41, re-scans the same vectors): the device copy of a contiguous slab is kept
      if (_contig && Runtime::instance().cache_uploads() && buf.get() && buf.tag_ptr == (const void*)_first && buf.tag_n == _n * sizeof(V))
        return buf.get();
      void* d = buf.ensure(_n * sizeof(V));
      buf.tag_ptr = _contig ? (const void*)_first : nullptr; buf.tag_n = _n * sizeof(V);
      const V* src = _first;
      if (!_contig) {
        staging.resize(_n);
        for (size_t i = 0; i < _n; ++i) staging[i] = *_ptrs[i];
        src = staging.data();
      }
      check(hj3d_memcpy_h2d(Runtime::instance().ctx(), d, src, _n * sizeof(V)));
      return d;
    }
  private:
    T*              _first = nullptr;
    size_t          _n = 0;
    std::vector<T*> _ptrs;
    bool            _contig = true;
};

template <typename Tdata, alg_hashfun_c Thashfun>
class DeviceTable {
  public:
    using data_t      = Tdata;
    using hashfun_t   = Thashfun;
    using hashvalue_t = typename hashfun_t::output_t;
    using stats_t     = HtStatistics;

    DeviceTable(int kind, size_t nbuckets) : _kind(kind), _numBuckets(nbuckets), _ks(keyspec_of<hashfun_t>()) {
      check(hj3d_table_create(Runtime::instance().ctx(), kind, nbuckets, &_t));
    }
    DeviceTable(const DeviceTable&) = delete;
    DeviceTable& operator=(const DeviceTable&) = delete;
    ~DeviceTable() { if (_t) hj3d_table_destroy(Runtime::instance().ctx(), _t); }

    size_t      numBuckets()              const { return _numBuckets; }
    hashvalue_t hash(const data_t* aData) const { return hashfun_t::eval(aData); }
    size_t      size()                    const { return _rows.size(); }

    void insert(data_t* aData) { _rows.push(aData); _sealed = false; _head.clear(); }
    void insert_bulk(data_t* first, size_t n) { _rows.push_bulk(first, n); _sealed = false; _head.clear(); }

    void clear() {                                               // ht_chaining.hh:250-258 / ht_nested.hh:438-447
      _rows.clear(); _head.clear();
      check(hj3d_table_clear(Runtime::instance().ctx(), _t));
      _sealed = false;
    }

    // build the device table from everything inserted so far (called by the build operator's fin())
    bool seal() {                                                // true: a device build ran
      if (_sealed) return false;
      hj3d_ctx* c = Runtime::instance().ctx();
      check(hj3d_table_clear(c, _t));
      void* d = _rows.upload(_dbuild, _staging);
      check(hj3d_table_build(c, _t, d, _rows.size(), _ks));
      _sealed = true;
      return true;
    }

    stats_t makeStatistics() const { return HtStatistics::from(raw_stats()); }
    hj3d_stats raw_stats() const {
      const_cast<DeviceTable*>(this)->seal();
      hj3d_stats s;
      check(hj3d_table_stats(Runtime::instance().ctx(), _t, &s));
      return s;
    }

    // used by the probe / unnest operators
    hj3d_table*         handle()  const { const_cast<DeviceTable*>(this)->seal(); return _t; }
    const hj3d_keyspec& keyspec() const { return _ks; }
    data_t*             row(uint32_t rowid) const { return _rows.at(rowid); }
    // predicate "row id is the first tuple inserted into its bucket" (= the directory entry, the head of the chain walk,
    // ht_chaining.hh:185-194); only the tuple-materialising host path of AlgHashJoinProbe needs it (emission order)
    auto bucket_min_rows() const {
      if (_head.size() != _numBuckets) {
        _head.assign(_numBuckets, 0xFFFFFFFFu);
        for (size_t i = 0; i < _rows.size(); ++i) {
          const size_t b = (size_t)hashfun_t::eval(_rows.at(i)) % _numBuckets;
          if (_head[b] == 0xFFFFFFFFu) _head[b] = (uint32_t)i;
        }
      }
      return [this](uint32_t rowid) { return _head[(size_t)hashfun_t::eval(_rows.at(rowid)) % _numBuckets] == rowid; };
    }

  protected:
    int          _kind;
    size_t       _numBuckets;
    hj3d_keyspec _ks;
    hj3d_table*  _t = nullptr;
    TupleSeq<data_t> _rows;
    DevBuf       _dbuild;
    std::vector<std::remove_const_t<data_t>> _staging;
    bool         _sealed = false;
    mutable std::vector<uint32_t> _head;   // bucket -> first inserted row (lazily, bucket_min_rows)
};

}  // namespace hj3d::detail

// ---------------------------------------------------------------------------------------------------
template <typename Tdata, alg_hashfun_c Thashfun, alg_binary_predicate_c Tcontenteqfun>
class HtChaining1 : public hj3d::detail::DeviceTable<Tdata, Thashfun> {
  static_assert(std::same_as<Tdata, typename Thashfun::input_t>, "Thashfun::input_t does not match Tdata");
  using base_t = hj3d::detail::DeviceTable<Tdata, Thashfun>;
  public:
    using data_t = Tdata; using hashfun_t = Thashfun; using hashvalue_t = typename Thashfun::output_t;
    using eqfun_t = Tcontenteqfun; using stats_t = HtStatistics;
    // layout of the reference's node (ht_chaining.hh:69-72): the drivers print sizeof(Node) (main_experiment1.cc:707)
    struct Node { Node* _next; data_t* _data; hashvalue_t _hashvalue; };

    HtChaining1(const size_t aNumBuckets, [[maybe_unused]] const uint32_t aReservoirLog2ChunkSize)
      : base_t(HJ3D_CHAINING, aNumBuckets) {
      hj3d::check_key_equality_predicate<eqfun_t, hashfun_t, hashfun_t>("HtChaining1: Tcontenteqfun");
    }

    size_t getRsvSize()              const { return this->raw_stats().rsv_main; }          // ht_chaining.hh:113
    size_t memoryConsupmtion()       const { auto s = this->raw_stats(); return s.mem_dir + s.mem_main; }
    size_t memoryConsupmtionDir()    const { return this->raw_stats().mem_dir; }
    size_t memoryConsupmtionChains() const { return this->raw_stats().mem_main; }
};

template <typename Tdata, alg_hashfun_c Thashfun, alg_binary_predicate_c Tcontenteqfun>
class HtNested1 : public hj3d::detail::DeviceTable<Tdata, Thashfun> {
  static_assert(std::same_as<Tdata, typename Thashfun::input_t>, "Thashfun::input_t does not match Tdata");
  using base_t = hj3d::detail::DeviceTable<Tdata, Thashfun>;
  public:
    using data_t = Tdata; using hashfun_t = Thashfun; using hashvalue_t = typename Thashfun::output_t;
    using eqfun_t = Tcontenteqfun; using stats_t = HtStatistics;

    struct SubNode { SubNode* _next; data_t* _data; };                                      // ht_nested.hh:163-166
    // Same layout as the reference's MainNode (ht_nested.hh:127-130).  On the host a MainNode is a handle to a
    // key group of the device table: _data is the group's first tuple, the group index and the owning table
    // ride in the two link fields (the sub chain itself is only walked on the device, by AlgUnnestHt).
    struct MainNode {
      MainNode*   _next;            // owner table (opaque)
      SubNode*    _subchain_head;   // group_ref (opaque)
      data_t*     _data;
      hashvalue_t _hashvalue;
      const data_t*     data()      const { return _data; }
      hashvalue_t       hashvalue() const { return _hashvalue; }
      bool              isEmpty()   const { return _data == nullptr; }
      uint32_t          group_ref() const { return (uint32_t)reinterpret_cast<uintptr_t>(_subchain_head); }
      const HtNested1*  owner()     const { return reinterpret_cast<const HtNested1*>(_next); }
    };

    HtNested1(const size_t aNumBuckets, [[maybe_unused]] const uint32_t aMainRsvLog2ChunkSize,
              [[maybe_unused]] const uint32_t aSubRsvLog2ChunkSize)
      : base_t(HJ3D_NESTED, aNumBuckets) {
      hj3d::check_key_equality_predicate<eqfun_t, hashfun_t, hashfun_t>("HtNested1: Tcontenteqfun");
    }

    size_t getRsvMainSize()              const { return this->raw_stats().rsv_main; }      // ht_nested.hh:192-193
    size_t getRsvSubSize()               const { return this->raw_stats().rsv_sub; }
    size_t memoryConsupmtion()           const { auto s = this->raw_stats(); return s.mem_dir + s.mem_main + s.mem_sub; }
    size_t memoryConsupmtionDir()        const { return this->raw_stats().mem_dir; }
    size_t memoryConsupmtionMainChains() const { return this->raw_stats().mem_main; }
    size_t memoryConsupmtionSubChains()  const { return this->raw_stats().mem_sub; }

    MainNode make_node(uint32_t group_ref, uint32_t first_row) const {
      MainNode m;
      m._next = reinterpret_cast<MainNode*>(const_cast<HtNested1*>(this));
      m._subchain_head = reinterpret_cast<SubNode*>(static_cast<uintptr_t>(group_ref));
      m._data = this->row(first_row);
      m._hashvalue = Thashfun::eval(m._data);
      return m;
    }
};
