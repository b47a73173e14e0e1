/*
 * oracle_join.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle_join.h).
 *
 * Plain-C restatement of the reference's single-threaded chaining and nested
 * ("3D") hash tables and of the probe / unnest operators that use them.  Nodes
 * are linked exactly as in the reference (directory entry holds the first
 * tuple, later tuples go to reservoir nodes), only that "pointers" are indices
 * and the stored data_t* is a row id.
 */
#include "oracle_join.h"

#include <stdlib.h>
#include <string.h>

#define NIL   ((int64_t)-1)  /* nullptr                                                  */
#define EMPTY ((int64_t)-2)  /* EMPTY_ENTRY sentinel 0x1 (ht_chaining.hh:65, ht_nested.hh:106) */

/* ---- util/hasht.hh:52-61 ---- */
uint32_t orc_murmur32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85ebca6bu;
  x ^= x >> 13;
  x *= 0xc2b2ae35u;
  x ^= x >> 16;
  return x;
}

/* ---- util/hasht.hh:63-72 ---- */
uint64_t orc_murmur64(uint64_t x) {
  x ^= (x >> 33);
  x *= 0xFF51AFD7ED558CCDull;
  x ^= (x >> 33);
  x *= 0xC4CEB9FE1A95EC63ull;
  x ^= (x >> 33);
  return x;
}

uint64_t orc_pair_mix(uint32_t left, uint32_t right) {
  uint64_t x = ((uint64_t)left << 32) | (uint64_t)right;
  x *= 0x9E3779B97F4A7C15ull;
  x ^= x >> 32;
  return x;
}

static inline void count_pair(orc_counters* c, uint32_t l, uint32_t r) {
  uint64_t m = orc_pair_mix(l, r);
  c->checksum_sum += m;
  c->checksum_xor ^= m;
}

/* raw key bits of tuple i (zero extended) */
static inline uint64_t load_key(const void* tuples, uint64_t i, const orc_keyspec* ks) {
  const unsigned char* p = (const unsigned char*)tuples + i * (uint64_t)ks->tuple_bytes + ks->key_offset;
  if (ks->key_bytes == 8) { uint64_t v; memcpy(&v, p, 8); return v; }
  uint32_t v; memcpy(&v, p, 4); return v;
}

static inline uint32_t load_rowid(const void* tuples, uint64_t i, const orc_keyspec* ks) {
  if (ks->rowid_offset == 0xFFFFFFFFu) return (uint32_t)i;
  const unsigned char* p = (const unsigned char*)tuples + i * (uint64_t)ks->tuple_bytes + ks->rowid_offset;
  uint32_t v; memcpy(&v, p, 4); return v;
}

/* the Hashfun* functors of the drivers: main_experiment1.cc:288-301, main_experiment4.cc:348-368,
 * main_algebra_example.cc:48-66 */
static inline uint64_t hash_key(uint64_t key, uint32_t hash_id) {
  switch (hash_id) {
    case 0:  return orc_murmur32((uint32_t)key);
    case 1:  return orc_murmur64(key);
    default: return orc_murmur64((uint64_t)(int64_t)(int32_t)(uint32_t)key);
  }
}

/* ------------------------------------------------------------------ tables */

typedef struct {          /* HtChaining1::Node, ht_chaining.hh:69-103 (24 B in the reference) */
  int64_t  next;
  uint64_t hash;
  uint64_t key;           /* the reference dereferences _data for the key; we keep a copy */
  uint32_t data;          /* row id instead of data_t*                                    */
} cnode;

typedef struct {          /* HtNested1::MainNode, ht_nested.hh:111-160 (32 B in the reference) */
  int64_t  next;
  int64_t  sub_head;
  uint64_t hash;
  uint64_t key;
  uint32_t data;
  uint32_t gid;           /* creation order index (our group_ref)                         */
} mnode;

typedef struct {          /* HtNested1::SubNode, ht_nested.hh:163-183 (16 B in the reference) */
  int64_t  next;
  uint32_t data;
} snode;

struct orc_table {
  int        kind;
  uint64_t   D, size;
  uint32_t   hash_id, key_bytes;
  /* chaining */
  cnode*     cdir;  cnode* cpool;  uint64_t cpool_n, cpool_cap;
  /* nested: main nodes >= 0 live in mpool, dir entries are addressed as -(idx)-3 */
  mnode*     mdir;  mnode* mpool;  uint64_t mpool_n, mpool_cap;
  snode*     spool; uint64_t spool_n, spool_cap;
  int64_t*   gid2node; uint64_t ngroups, gid_cap;
};

#define DIRREF(i)   (-(int64_t)(i) - 3)
#define IS_DIRREF(r) ((r) <= -3)
#define DIRIDX(r)   ((uint64_t)(-((r) + 3)))

static inline mnode* mref(const orc_table* t, int64_t r) {
  return IS_DIRREF(r) ? &t->mdir[DIRIDX(r)] : &t->mpool[r];
}

static void* grow(void* p, uint64_t* cap, uint64_t need, size_t elem) {
  if (need <= *cap) return p;
  uint64_t nc = *cap ? *cap * 2 : 1024;  /* Reservoir chunking (util/reservoir.hh:155-172) is unobservable */
  if (nc < need) nc = need;
  p = realloc(p, nc * elem);
  *cap = nc;
  return p;
}

/* HtChaining1::insert, ht_chaining.hh:181-196 */
static void chaining_insert(orc_table* t, uint64_t key, uint32_t rowid) {
  uint64_t h = hash_key(key, t->hash_id);
  cnode* d = &t->cdir[h % t->D];                       /* getDirIndex, ht_chaining.hh:139 */
  if (d->next == EMPTY) {                              /* :185-188 */
    d->data = rowid; d->key = key; d->hash = h; d->next = NIL;
  } else {                                             /* :189-194 new node right after the dir entry */
    t->cpool = (cnode*)grow(t->cpool, &t->cpool_cap, t->cpool_n + 1, sizeof(cnode));
    d = &t->cdir[h % t->D];
    cnode* nn = &t->cpool[t->cpool_n];
    nn->data = rowid; nn->key = key; nn->hash = h; nn->next = d->next;
    d->next = (int64_t)t->cpool_n++;
  }
  ++t->size;
}

static void register_group(orc_table* t, int64_t ref) {
  t->gid2node = (int64_t*)grow(t->gid2node, &t->gid_cap, t->ngroups + 1, sizeof(int64_t));
  mref(t, ref)->gid = (uint32_t)t->ngroups;
  t->gid2node[t->ngroups++] = ref;
}

/* HtNested1::insertAtMainNode + insertIntoSubchain, ht_nested.hh:386-412 */
static void nested_insert_at(orc_table* t, int64_t ref, uint64_t key, uint64_t h, uint32_t rowid) {
  mnode* m = mref(t, ref);
  if (m->next == EMPTY) {                              /* :389-390  MainNode::init :147-152 */
    m->next = NIL; m->sub_head = NIL; m->data = rowid; m->key = key; m->hash = h;
    register_group(t, ref);
  } else {                                             /* :399-412 new sub node becomes the head */
    t->spool = (snode*)grow(t->spool, &t->spool_cap, t->spool_n + 1, sizeof(snode));
    snode* s = &t->spool[t->spool_n];
    s->data = rowid;
    s->next = m->sub_head;                             /* NIL if the chain was empty */
    m->sub_head = (int64_t)t->spool_n++;
  }
}

/* HtNested1::insert, ht_nested.hh:287-311 (+ findMainNode :414-436, isMainNodeMatch :241-243) */
static void nested_insert(orc_table* t, uint64_t key, uint32_t rowid) {
  uint64_t h = hash_key(key, t->hash_id);
  uint64_t di = h % t->D;
  int64_t  ref = DIRREF(di);
  mnode*   d = &t->mdir[di];
  if (d->next == EMPTY || (d->hash == h && d->key == key)) {      /* :294 */
    nested_insert_at(t, ref, key, h, rowid);
  } else {
    /* findMainNode(aData, dirEntry): walk while hasNext, testing the *next* node (:428-434) */
    int64_t cur = ref;
    int     found = 0;
    for (;;) {
      mnode* c = mref(t, cur);
      if (c->next == NIL) break;                                  /* !hasNext */
      mnode* nx = mref(t, c->next);
      if (nx->hash == h && nx->key == key) { cur = c->next; found = 1; break; }
      cur = c->next;
    }
    if (found) {
      nested_insert_at(t, cur, key, h, rowid);                    /* :300-302 */
    } else {                                                      /* :303-308 append at the TAIL */
      t->mpool = (mnode*)grow(t->mpool, &t->mpool_cap, t->mpool_n + 1, sizeof(mnode));
      int64_t nn = (int64_t)t->mpool_n++;
      t->mpool[nn].next = EMPTY;                                  /* Reservoir entry is default constructed */
      t->mpool[nn].sub_head = NIL;
      mref(t, cur)->next = nn;
      nested_insert_at(t, nn, key, h, rowid);
    }
  }
  ++t->size;
}

orc_table* orc_build(int kind, const void* tuples, uint64_t n, orc_keyspec ks, uint64_t num_buckets) {
  if (num_buckets == 0) return NULL;
  orc_table* t = (orc_table*)calloc(1, sizeof(orc_table));
  t->kind = kind; t->D = num_buckets; t->hash_id = ks.hash_id; t->key_bytes = ks.key_bytes;
  if (kind == 0) {
    t->cdir = (cnode*)malloc(num_buckets * sizeof(cnode));        /* ctor ht_chaining.hh:106-107 */
    for (uint64_t i = 0; i < num_buckets; ++i) { t->cdir[i].next = EMPTY; t->cdir[i].hash = 0; t->cdir[i].key = 0; t->cdir[i].data = 0; }
    for (uint64_t i = 0; i < n; ++i)                              /* AlgScan::run -> AlgHashJoinBuild::step, algebra.hh:263-266,574-577 */
      chaining_insert(t, load_key(tuples, i, &ks), load_rowid(tuples, i, &ks));
  } else {
    t->mdir = (mnode*)malloc(num_buckets * sizeof(mnode));        /* ctor ht_nested.hh:255-259 */
    for (uint64_t i = 0; i < num_buckets; ++i) { t->mdir[i].next = EMPTY; t->mdir[i].sub_head = NIL; t->mdir[i].hash = 0; t->mdir[i].key = 0; t->mdir[i].data = 0; t->mdir[i].gid = 0; }
    for (uint64_t i = 0; i < n; ++i)                              /* AlgNestJoinBuild::step, algebra.hh:386-389 */
      nested_insert(t, load_key(tuples, i, &ks), load_rowid(tuples, i, &ks));
  }
  return t;
}

void orc_table_free(orc_table* t) {
  if (!t) return;
  free(t->cdir); free(t->cpool); free(t->mdir); free(t->mpool); free(t->spool); free(t->gid2node);
  free(t);
}

/* ---- util/aggregate.hh:27-52 ---- */
typedef struct { uint64_t mn, mx, sum, sumsq, cnt; } agg;
static void agg_init(agg* a) { a->mn = UINT64_MAX; a->mx = 0; a->sum = a->sumsq = a->cnt = 0; }
static void agg_step(agg* a, uint64_t x) {
  if (x < a->mn) a->mn = x;
  if (x > a->mx) a->mx = x;
  a->sum += x; a->sumsq += x * x; a->cnt += 1;
}

static int cmp_u32(const void* a, const void* b) {
  uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
  return (x > y) - (x < y);
}

/* makeStatistics: ht_chaining.hh:260-292, ht_nested.hh:450-482 */
void orc_table_stats(const orc_table* t, orc_stats* s) {
  memset(s, 0, sizeof(*s));
  agg all, ne; agg_init(&all); agg_init(&ne);
  s->num_buckets = t->D;
  s->num_entries = t->size;
  if (t->kind == 0) {
    /* _numDistinctKeys = |unordered_set<key_t>| over node hash values, key_t == POSIX int
     * (ht_chaining.hh:267,282): distinct low 32 bits of the hash values. */
    uint32_t* hv = (uint32_t*)malloc((t->size ? t->size : 1) * sizeof(uint32_t));
    uint64_t  nh = 0;
    for (uint64_t b = 0; b < t->D; ++b) {
      const cnode* d = &t->cdir[b];
      if (d->next == EMPTY) { ++s->num_empty; agg_step(&all, 0); continue; }
      uint64_t len = 0;
      for (const cnode* c = d; c; c = (c->next == NIL ? NULL : &t->cpool[c->next])) { ++len; hv[nh++] = (uint32_t)c->hash; }
      agg_step(&all, len); agg_step(&ne, len);
    }
    qsort(hv, nh, sizeof(uint32_t), cmp_u32);
    uint64_t dk = 0;
    for (uint64_t i = 0; i < nh; ++i) if (i == 0 || hv[i] != hv[i - 1]) ++dk;
    free(hv);
    s->num_distinct_keys = dk;
    s->rsv_main = t->cpool_n;                       /* getRsvSize, ht_chaining.hh:113 */
    s->rsv_sub  = 0;
    s->mem_dir  = t->D * 24;                        /* memoryConsupmtionDir, :169-171, sizeof(Node) == 24 */
    s->mem_main = t->cpool_n * 24;                  /* memoryConsupmtionChains, :173-177 */
    s->mem_sub  = 0;
  } else {
    for (uint64_t b = 0; b < t->D; ++b) {
      const mnode* d = &t->mdir[b];
      if (d->next == EMPTY) { ++s->num_empty; agg_step(&all, 0); continue; }
      uint64_t len = 0;
      for (const mnode* c = d; c; c = (c->next == NIL ? NULL : &t->mpool[c->next])) { ++len; ++s->num_distinct_keys; }
      agg_step(&all, len); agg_step(&ne, len);
    }
    s->rsv_main = t->mpool_n;                       /* getRsvMainSize, ht_nested.hh:192 */
    s->rsv_sub  = t->spool_n;                       /* getRsvSubSize,  ht_nested.hh:193 */
    s->mem_dir  = t->D * 32;                        /* ht_nested.hh:270-272, sizeof(MainNode) == 32 */
    s->mem_main = t->mpool_n * 32;                  /* :276-278 */
    s->mem_sub  = t->spool_n * 16;                  /* :282-284, sizeof(SubNode) == 16 */
  }
  s->cc_min = all.mn; s->cc_max = all.mx; s->cc_sum = all.sum; s->cc_sumsq = all.sumsq; s->cc_count = all.cnt;
  s->ccne_min = ne.mn; s->ccne_max = ne.mx; s->ccne_sum = ne.sum; s->ccne_sumsq = ne.sumsq; s->ccne_count = ne.cnt;
}

/* ------------------------------------------------------------------ operators */

static inline void emit(uint32_t* out, uint64_t cap, orc_counters* c, uint32_t l, uint32_t r) {
  if (out) {
    if (c->out_tuples < cap) { out[2 * c->out_tuples] = l; out[2 * c->out_tuples + 1] = r; ++c->out_written; }
    else c->overflow = 1;
  }
  ++c->out_tuples;
}

/* AlgHashJoinProbe::step, algebra.hh:625-659 */
void orc_probe_chaining(const orc_table* t, const void* tuples, uint64_t n, orc_keyspec ks,
                        const uint32_t* gather, int build_key_unique,
                        uint32_t* out, uint64_t cap, orc_counters* c) {
  memset(c, 0, sizeof(*c));
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t src = gather ? gather[i] : i;
    uint64_t key = load_key(tuples, src, &ks);
    uint64_t h   = hash_key(key, ks.hash_id);                     /* :631 */
    const cnode* it = &t->cdir[h % t->D];                         /* findDirEntryByOther, ht_chaining.hh:236-248 */
    if (it->next == EMPTY) continue;                              /* :640-643 (no comparisons counted) */
    uint64_t cmps = 0;
    for (; it; it = (it->next == NIL ? NULL : &t->cpool[it->next])) {
      ++cmps;                                                     /* :646 */
      if (it->hash == h && it->key == key) {                      /* :647-648 hash equality AND join predicate */
        ++c->matches;                                             /* inc(), :651 */
        count_pair(c, (uint32_t)i, it->data);
        emit(out, cap, c, (uint32_t)i, it->data);
        if (build_key_unique) break;                              /* :653-655 */
      }
    }
    c->num_cmps += cmps;                                          /* :658 */
  }
}

/* AlgNestJoinProbe::step (algebra.hh:435-459) with findMainNodeByOther (ht_nested.hh:354-382) */
void orc_probe_nested(const orc_table* t, const void* tuples, uint64_t n, orc_keyspec ks,
                      const uint32_t* gather, uint32_t* out, uint64_t cap, orc_counters* c) {
  memset(c, 0, sizeof(*c));
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t src = gather ? gather[i] : i;
    uint64_t key = load_key(tuples, src, &ks);
    uint64_t h   = hash_key(key, ks.hash_id);
    const mnode* m = &t->mdir[h % t->D];
    uint64_t cmps = 0;
    const mnode* hit = NULL;
    do {                                                          /* ht_nested.hh:371-379 */
      if (m->next == EMPTY) break;                                /* :372 */
      ++cmps;
      if (m->hash == h && m->key == key) { hit = m; break; }      /* :374-377 */
      m = (m->next == NIL) ? NULL : &t->mpool[m->next];
    } while (m);
    c->num_cmps += cmps;                                          /* algebra.hh:449 */
    if (hit) {                                                    /* algebra.hh:453-458 */
      ++c->matches;
      count_pair(c, (uint32_t)i, hit->data);   /* checksum over (left, first row of the group): group ids are layout specific */
      emit(out, cap, c, (uint32_t)i, hit->gid);
    }
  }
}

/* AlgUnnestHt::step, algebra.hh:510-541 */
void orc_unnest(const orc_table* t, const uint32_t* left, const uint32_t* gref, uint64_t n,
                uint32_t* out, uint64_t cap, orc_counters* c) {
  memset(c, 0, sizeof(*c));
  for (uint64_t i = 0; i < n; ++i) {
    const mnode* m = mref(t, t->gid2node[gref[i]]);               /* getMainNode */
    ++c->matches;
    count_pair(c, left[i], m->data);                              /* :526-530 first element lives in the MainNode */
    emit(out, cap, c, left[i], m->data);
    for (int64_t s = m->sub_head; s != NIL; s = t->spool[s].next) {  /* :532-539 */
      count_pair(c, left[i], t->spool[s].data);
      emit(out, cap, c, left[i], t->spool[s].data);
    }
  }
  c->matches = c->out_tuples;                                     /* AlgUnnestHt::_count counts outputs (:486-487) */
}

uint64_t orc_num_groups(const orc_table* t) { return t->ngroups; }

uint64_t orc_group_len(const orc_table* t, uint32_t gref) {
  const mnode* m = mref(t, t->gid2node[gref]);
  uint64_t len = 1;
  for (int64_t s = m->sub_head; s != NIL; s = t->spool[s].next) ++len;
  return len;
}
