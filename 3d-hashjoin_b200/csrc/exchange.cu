// exchange.cu -- the multi-GPU data plane: bucket-range sharding of a relation over the GPUs of one node.
//
// The join shards by bucket range: equal keys -> equal bucket, so a partition that is a function of the bucket index keeps
// every chain / key group on one GPU.  The exchange IS partition level 1 of the engine: a rank partitions its slice of a
// relation into the coarse bucket ranges the local join works with anyway, and the partition kernel stores every range
// straight into the receive buffer of the GPU that owns it -- peer-mapped memory, so the sorted runs of a tile travel
// over NVLink as the kernel writes them (k_part_scatter<.., PEER>, partition.cuh).  There is no separate all-to-all and
// no host round trip between partitioning and the local join: what a rank receives is already the coarse-partitioned
// input of its build / probe pipeline (hj3d_table_build_parts / hj3d_probe_parts continue at partition level 2).
//
// Receive buffer of a rank: one region of cap_seg records per (owned range, source rank); a source reserves space in
// "its" regions with its own cursors, so no remote atomics are needed.  The per-(source, range) counts travel by ONE
// small all-gather, which is also the barrier that orders every peer's stores before the owner's reads.
//   * one process per GPU  : NCCL for the count all-gather, CUDA IPC to map the peers' receive buffers
//   * one process, N GPUs  : a local group (hj3d_comm_create_local): plain peer access, events and copies
// Skewed keys (Zipf) overflow fixed-capacity regions; HJ3D_XCHG_EXACT first exchanges a histogram and then packs the
// regions tightly at exact offsets (two passes over the local slice).
#include <dlfcn.h>
#include <nccl.h>

#include <memory>

#include "engine_internal.hh"
#include "hot.cuh"
#include "partition.cuh"

namespace {

// ---- NCCL, loaded at run time (the library is only needed for multi-process sharding) -------------------------------
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
  bool load() {
    if (h) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) { h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
    GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
    AllGather = (decltype(AllGather))dlsym(h, "ncclAllGather");
    AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
    GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllGather || !AllReduce || !GetErrorString) { err = "libnccl lacks a required symbol"; h = nullptr; return false; }
    return true;
  }
};
NcclApi& nccl() { static NcclApi a; return a; }

#define NCCL_TRY(expr)                                                                                       \
  do {                                                                                                       \
    ncclResult_t _r = (expr);                                                                                \
    if (_r != ncclSuccess) return fail(HJ3D_ERR_CUDA, std::string(#expr) + ": " + nccl().GetErrorString(_r)); \
  } while (0)

constexpr int kSlots = 2;              // relations in flight (build side, probe side)
constexpr uint32_t kMaxRanges = 1024;

struct ExchangePlan {
  uint64_t D = 0;
  uint32_t width = 0, n_ranges = 0, rpo = 1, rpo_shift = 0;   // bucket ranges, ranges per owner (power of two)
  uint64_t lo(uint32_t r) const { const uint64_t v = (uint64_t)r * rpo * width; return v < D ? v : D; }
  uint64_t hi(uint32_t r) const { const uint64_t v = (uint64_t)(r + 1) * rpo * width; return v < D ? v : D; }
  uint32_t owned(uint32_t r) const { const uint64_t a = (uint64_t)r * rpo, b = a + rpo; return a >= n_ranges ? 0u : (uint32_t)((b < n_ranges ? b : n_ranges) - a); }
};

struct LocalGroup;   // single-process mode: what the ranks share

}  // namespace

struct hj3d_comm {
  hj3d_ctx* ctx = nullptr;
  int world = 1, rank = 0;
  ncclComm_t nc = nullptr;                       // multi-process mode
  std::shared_ptr<LocalGroup> group;             // single-process mode
  int64_t target_ranges = 256, min_width = 16384, max_width = 1 << 21, xchg_threads = 0;
  // receive buffers of this rank and the peers' mapped views of them
  void*    recv[kSlots] = {nullptr, nullptr};
  uint64_t recv_records[kSlots] = {0, 0};
  uint32_t rec_bytes[kSlots] = {8, 8};
  void*    peer_recv[kSlots][kMaxPeers] = {};
  bool     peer_is_ipc[kSlots][kMaxPeers] = {};
  // per-slot exchange state
  unsigned long long* d_cursor[kSlots] = {nullptr, nullptr};   // [kMaxRanges] my counts per range
  unsigned long long* d_all[kSlots] = {nullptr, nullptr};      // [world][kMaxRanges] everybody's counts
  unsigned long long* d_pstart[kSlots] = {nullptr, nullptr};   // [kMaxRanges + 1] my region offsets inside the owners' buffers
  unsigned long long* d_bar = nullptr;                         // [1 + kMaxPeers] barrier scratch (exact mode)
  cudaEvent_t ev_scatter[kSlots] = {nullptr, nullptr};
  ExchangePlan plan[kSlots];
  uint64_t cap_seg[kSlots] = {0, 0};
  hj3d_keyspec ks[kSlots] = {};
  uint64_t n_local[kSlots] = {0, 0};
  bool exact[kSlots] = {false, false};
  bool pending[kSlots] = {false, false};
  bool streaming[kSlots] = {false, false};       // HJ3D_XCHG_MORE: further chunks of the local slice follow (hj3d_exchange_append)
  // host-resident slices (hj3d_exchange_begin_host): chunked upload on the ctx's copy stream into a comm-owned staging buffer,
  // one event per chunk, level 1 per chunk on the comm's own exchange stream (the caller's stream stays free until _end)
  cudaStream_t xstream = nullptr;
  cudaEvent_t ev_fence = nullptr;
  cudaEvent_t ev_ready[kSlots] = {nullptr, nullptr};             // counts gathered (NCCL) / scatter finished, on xstream
  std::vector<cudaEvent_t> chunk_ev[kSlots];
  size_t n_chunks[kSlots] = {0, 0};
  void*  stage[kSlots] = {nullptr, nullptr};
  size_t stage_bytes[kSlots] = {0, 0};
  bool   host_streamed[kSlots] = {false, false};
  bool   async_on[kSlots] = {false, false};       // HJ3D_XCHG_ASYNC: the exchange runs on xstream, hj3d_exchange_end joins it
  hj3d_selection sel[kSlots] = {};
  void* h_pinned = nullptr;                      // world * kMaxRanges * 8 bytes
  // hot-key probe replication (hot.cuh)
  unsigned long long* d_sample[kSlots] = {nullptr, nullptr};      // [kHotSample] this rank's sample of the slot's relation
  unsigned long long* d_sample_all[kSlots] = {nullptr, nullptr};  // [kHotSample] everybody's
  void* d_hot_table[kSlots] = {nullptr, nullptr};                 // HotTable<KeyT>
  void* hot_buf[kSlots] = {nullptr, nullptr};                     // local segment of the hot tuples' records
  uint64_t hot_buf_records[kSlots] = {0, 0};
  bool  hot_sampled[kSlots] = {false, false}, hot_selected[kSlots] = {false, false}, hot_on[kSlots] = {false, false};
  uint32_t sample_m[kSlots] = {0, 0};
  HotAnswers* d_ans[kSlots] = {nullptr, nullptr};                 // this rank's answers / the sum over all ranks
  HotAnswers* d_ans_sum[kSlots] = {nullptr, nullptr};
  cudaEvent_t ev_sample[kSlots] = {nullptr, nullptr}, ev_ans[kSlots] = {nullptr, nullptr};
};

namespace {

struct LocalGroup { std::vector<hj3d_comm*> ranks; };

ExchangePlan make_plan(const hj3d_comm* cm, uint64_t D) {
  ExchangePlan p; p.D = D;
  uint64_t w = 1;
  const uint64_t want = (D + (uint64_t)cm->target_ranges - 1) / (uint64_t)cm->target_ranges;
  while (w < want) w <<= 1;
  uint64_t mw = 1; while ((int64_t)mw < cm->min_width) mw <<= 1;
  if (w < mw) w = mw;
  // the local join continues at partition level 2 only if a range holds at most 1024 fine partitions (2048 buckets each for
  // 8-byte slots of unique keys): wider ranges would be compacted and partitioned from scratch (engine.cu, plan_probe)
  uint64_t xw = 1; while ((int64_t)(xw << 1) <= cm->max_width) xw <<= 1;
  if (w > xw && xw >= mw) w = xw;
  while ((D + w - 1) / w > kMaxRanges) w <<= 1;
  p.width = (uint32_t)(w > 0x80000000ull ? 0x80000000ull : w);
  p.n_ranges = (uint32_t)((D + p.width - 1) / p.width);
  uint32_t per = (p.n_ranges + cm->world - 1) / cm->world;
  p.rpo = 1; p.rpo_shift = 0;
  while (p.rpo < per) { p.rpo <<= 1; ++p.rpo_shift; }
  return p;
}

// region offsets (uniform mode): range q of owner o = q >> shift, segment of source `me`
__global__ void k_xchg_starts(uint32_t n_ranges, uint32_t rpo_shift, uint32_t world, uint32_t me, unsigned long long cap_seg,
                              unsigned long long* __restrict__ pstart) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q > n_ranges) return;
  const uint32_t local = q & ((1u << rpo_shift) - 1u);
  pstart[q] = q == n_ranges ? 0ull : ((unsigned long long)local * world + me) * cap_seg;   // [n_ranges]: the hot segment, a buffer of its own
}

__global__ void k_add_u32(uint32_t* __restrict__ acc, const uint32_t* __restrict__ x, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) acc[i] += x[i];
}

// owner side: segment table of my ranges from the gathered counts.  uniform: fixed regions; exact: tightly packed
__global__ void k_xchg_segments(const unsigned long long* __restrict__ all, uint32_t stride, uint32_t world, uint32_t first_range,
                                uint32_t n_owned, unsigned long long cap_seg, int exact,
                                unsigned long long* __restrict__ seg_start, unsigned long long* __restrict__ seg_count) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  unsigned long long run = 0;
  for (uint32_t p = 0; p < n_owned; ++p)
    for (uint32_t s = 0; s < world; ++s) {
      const unsigned long long cnt = all[(size_t)s * stride + first_range + p];
      const uint32_t q = p * world + s;
      seg_start[q] = exact ? run : (unsigned long long)q * cap_seg;
      seg_count[q] = (!exact && cnt > cap_seg) ? cap_seg : cnt;      // what was actually stored (overflow is reported on the host)
      run += cnt;
    }
}

// TH = block size: 512 (tiles of 8192 8-byte records, two blocks per SM) or 1024 (tiles of 16384: the sorted runs a tile
// stores into one peer region are twice as long, which is what the NVLink store rate depends on; the default across GPUs)
template <int HASH, int TH>
int launch_scatter_t(hj3d_comm* cm, int slot, Src src, uint32_t rowid_base, unsigned long long cap, cudaStream_t st) {
  using KeyT = typename HashT<HASH>::key_t;
  hj3d_ctx* c = cm->ctx;
  const ExchangePlan& pl = cm->plan[slot];
  const int kTile = TH * PartCfg<KeyT>::kItems;
  const uint32_t nb = blocks_for(src.n, kTile);
  if (!nb) return HJ3D_OK;
  PeerOut peer{};
  for (int r = 0; r < cm->world; ++r) peer.base[r] = cm->peer_recv[slot][r];
  peer.owner_shift = pl.rpo_shift;
  const Dir d = make_dir(pl.D, 0, pl.D);
  const PartFn pf = make_partfn(pl.width, 0, (uint32_t)pl.D);
  if (cm->hot_on[slot]) {      // one more partition: the hot keys' tuples stay here
    peer.hot_q = pl.n_ranges; peer.hot_base = cm->hot_buf[slot]; peer.hot_cap = cm->hot_buf_records[slot]; peer.hot_table = cm->d_hot_table[slot];
    const uint32_t fan = pl.n_ranges + 1;
    const size_t sm = part_smem_bytes<KeyT>(fan, TH, false);
    auto kfn = k_part_scatter<HASH, false, false, TH, false, true, true>;
    CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kfn<<<nb, TH, sm, st>>>(src, nullptr, d, pf, fan, fan, rowid_base, cap, cm->d_pstart[slot], cm->d_cursor[slot], (Slot<KeyT>*)nullptr, peer);
  } else {
    const size_t sm = part_smem_bytes<KeyT>(pl.n_ranges, TH, false);
    auto kfn = k_part_scatter<HASH, false, false, TH, false, true>;
    CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kfn<<<nb, TH, sm, st>>>(src, nullptr, d, pf, pl.n_ranges, pl.n_ranges, rowid_base, cap, cm->d_pstart[slot], cm->d_cursor[slot],
                            (Slot<KeyT>*)nullptr, peer);
  }
  ++c->launches;
  CUDA_TRY(cudaGetLastError());
  return HJ3D_OK;
}

template <int HASH>
int launch_scatter(hj3d_comm* cm, int slot, Src src, uint32_t rowid_base, unsigned long long cap, cudaStream_t st = nullptr) {
  if (!st) st = cm->ctx->stream;
  const int64_t th = cm->xchg_threads ? cm->xchg_threads : (cm->world > 1 ? 1024 : 512);   // measured at 2 GPUs: 4.55 vs 4.75 ms (128 ranges), 11.8 vs 12.8 ms (512)
  if (th == 1024) return launch_scatter_t<HASH, 1024>(cm, slot, src, rowid_base, cap, st);
  return launch_scatter_t<HASH, 512>(cm, slot, src, rowid_base, cap, st);
}

template <int HASH>
int launch_hist(hj3d_comm* cm, int slot, Src src) {
  hj3d_ctx* c = cm->ctx;
  const ExchangePlan& pl = cm->plan[slot];
  const uint32_t nb = blocks_for(src.n, kPartTile);
  if (!nb) return HJ3D_OK;
  k_part_hist<HASH><<<nb, kPartThreads, 0, c->stream>>>(src, make_dir(pl.D, 0, pl.D), make_partfn(pl.width, 0, (uint32_t)pl.D), pl.n_ranges, cm->d_cursor[slot]);
  ++c->launches;
  CUDA_TRY(cudaGetLastError());
  return HJ3D_OK;
}

// everybody's counts of this slot -> d_all (also the barrier: a rank's contribution leaves after its scatter kernel)
int gather_counts(hj3d_comm* cm, int slot, cudaStream_t st = nullptr) {
  hj3d_ctx* c = cm->ctx;
  if (cm->nc) {
    NCCL_TRY(nccl().AllGather(cm->d_cursor[slot], cm->d_all[slot], kMaxRanges, ncclUint64, cm->nc, st ? st : c->stream));
    return HJ3D_OK;
  }
  for (hj3d_comm* o : cm->group->ranks) {       // single process: wait for every rank's kernel, then copy its counts over
    if (!o) continue;
    CUDA_TRY(cudaStreamWaitEvent(c->stream, o->ev_scatter[slot], 0));       // (a rank's own scatter may have run on its exchange stream)
    CUDA_TRY(cudaMemcpyAsync(cm->d_all[slot] + (size_t)o->rank * kMaxRanges, o->d_cursor[slot], kMaxRanges * 8, cudaMemcpyDefault, c->stream));
  }
  return HJ3D_OK;
}

int free_slot(hj3d_comm* cm, int slot) {
  for (int r = 0; r < cm->world; ++r) {
    if (cm->peer_is_ipc[slot][r] && cm->peer_recv[slot][r]) cudaIpcCloseMemHandle(cm->peer_recv[slot][r]);
    cm->peer_recv[slot][r] = nullptr; cm->peer_is_ipc[slot][r] = false;
  }
  if (cm->recv[slot]) { cudaStreamSynchronize(cm->ctx->stream); cudaFree(cm->recv[slot]); }
  cm->recv[slot] = nullptr; cm->recv_records[slot] = 0;
  return HJ3D_OK;
}

int comm_alloc_state(hj3d_comm* cm) {
  CUDA_TRY(cudaSetDevice(cm->ctx->device));
  for (int s = 0; s < kSlots; ++s) {
    CUDA_TRY(cudaMalloc((void**)&cm->d_cursor[s], (kMaxRanges + 8) * 8));
    CUDA_TRY(cudaMalloc((void**)&cm->d_sample[s], kHotSample * 8));
    CUDA_TRY(cudaMalloc((void**)&cm->d_sample_all[s], kHotSample * 8));
    CUDA_TRY(cudaMalloc(&cm->d_hot_table[s], sizeof(HotTable<uint64_t>)));
    CUDA_TRY(cudaMalloc((void**)&cm->d_ans[s], sizeof(HotAnswers)));
    CUDA_TRY(cudaMalloc((void**)&cm->d_ans_sum[s], sizeof(HotAnswers)));
    CUDA_TRY(cudaEventCreateWithFlags(&cm->ev_sample[s], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&cm->ev_ans[s], cudaEventDisableTiming));
    CUDA_TRY(cudaMalloc((void**)&cm->d_all[s], (size_t)cm->world * kMaxRanges * 8));
    CUDA_TRY(cudaMalloc((void**)&cm->d_pstart[s], (kMaxRanges + 1) * 8));
    CUDA_TRY(cudaEventCreateWithFlags(&cm->ev_scatter[s], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&cm->ev_ready[s], cudaEventDisableTiming));
  }
  CUDA_TRY(cudaEventCreateWithFlags(&cm->ev_fence, cudaEventDisableTiming));
  CUDA_TRY(cudaMalloc((void**)&cm->d_bar, (1 + kMaxPeers) * 8));
  CUDA_TRY(cudaMemset(cm->d_bar, 0, (1 + kMaxPeers) * 8));
  CUDA_TRY(cudaMallocHost(&cm->h_pinned, (size_t)(cm->world + 1) * kMaxRanges * 8));
  return HJ3D_OK;
}

template <int HASH> void launch_hot_select(hj3d_comm* cm, int slot, uint32_t m, cudaStream_t st) {
  using KeyT = typename HashT<HASH>::key_t;
  const int sm = kHotSample * (8 + 2);
  cudaFuncSetAttribute(k_hot_select<HASH>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
  k_hot_select<HASH><<<1, 1024, sm, st>>>(cm->d_sample_all[slot], m, (HotTable<KeyT>*)cm->d_hot_table[slot]);
}

}  // namespace

extern "C" {

int hj3d_comm_unique_id(void* id128) {
  if (!id128) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (!nccl().load()) return fail(HJ3D_ERR_UNSUPPORTED, nccl().err);
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  NCCL_TRY(nccl().GetUniqueId((ncclUniqueId*)id128));
  return HJ3D_OK;
}

int hj3d_comm_create(hj3d_ctx* c, int world, int rank, const void* id128, hj3d_comm** out) {
  if (!c || !out || world < 1 || rank < 0 || rank >= world || world > kMaxPeers) return fail(HJ3D_ERR_INVALID, "bad comm arguments");
  *out = nullptr;
  if (world > 1 && !id128) return fail(HJ3D_ERR_INVALID, "id128 == NULL");
  CUDA_TRY(cudaSetDevice(c->device));
  auto cm = std::make_unique<hj3d_comm>();
  cm->ctx = c; cm->world = world; cm->rank = rank;
  if (world > 1) cm->target_ranges = 128;   // longer runs per peer store: measured at 2 GPUs, 2^27 x 2^30: 11.7 ms against 12.3 (256) / 12.2 (64)
  if (world > 1) {
    if (!nccl().load()) return fail(HJ3D_ERR_UNSUPPORTED, nccl().err);
    ncclUniqueId id; memcpy(&id, id128, sizeof id);
    NCCL_TRY(nccl().CommInitRank(&cm->nc, world, id, rank));
  } else {
    cm->group = std::make_shared<LocalGroup>();
    cm->group->ranks.push_back(cm.get());
  }
  HJ_TRY(comm_alloc_state(cm.get()));
  *out = cm.release();
  return HJ3D_OK;
}

int hj3d_comm_create_local(hj3d_ctx** ctxs, int world, hj3d_comm** out) {
  if (!ctxs || !out || world < 1 || world > kMaxPeers) return fail(HJ3D_ERR_INVALID, "bad comm arguments");
  auto group = std::make_shared<LocalGroup>();
  std::vector<std::unique_ptr<hj3d_comm>> v;
  for (int r = 0; r < world; ++r) {
    if (!ctxs[r]) return fail(HJ3D_ERR_INVALID, "ctx == NULL");
    auto cm = std::make_unique<hj3d_comm>();
    cm->ctx = ctxs[r]; cm->world = world; cm->rank = r; cm->group = group;
    HJ_TRY(comm_alloc_state(cm.get()));
    v.push_back(std::move(cm));
  }
  for (int a = 0; a < world; ++a)                      // peer access between distinct devices of the group
    for (int b = 0; b < world; ++b) {
      const int da = ctxs[a]->device, db = ctxs[b]->device;
      if (da == db) continue;
      int can = 0;
      CUDA_TRY(cudaDeviceCanAccessPeer(&can, da, db));
      if (!can) return fail(HJ3D_ERR_UNSUPPORTED, "devices of the local group cannot access each other's memory");
      CUDA_TRY(cudaSetDevice(da));
      cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(HJ3D_ERR_CUDA, cudaGetErrorString(e));
      cudaGetLastError();
    }
  for (int r = 0; r < world; ++r) { group->ranks.push_back(v[r].get()); out[r] = v[r].release(); }
  return HJ3D_OK;
}

int hj3d_comm_set_option(hj3d_comm* cm, int opt, int64_t v) {
  if (!cm) return fail(HJ3D_ERR_INVALID, "NULL argument");
  switch (opt) {
    case HJ3D_XOPT_TARGET_RANGES: if (v >= 1 && v <= (int64_t)kMaxRanges) cm->target_ranges = v; break;
    case HJ3D_XOPT_MIN_RANGE_WIDTH: if (v >= 1) cm->min_width = v; break;
    case HJ3D_XOPT_MAX_RANGE_WIDTH: if (v >= 1) cm->max_width = v; break;
    case HJ3D_XOPT_THREADS: if (v == 0 || v == 512 || v == 1024) cm->xchg_threads = v; break;
    default: return fail(HJ3D_ERR_INVALID, "unknown comm option");
  }
  return HJ3D_OK;
}

int hj3d_comm_destroy(hj3d_comm* cm) {
  if (!cm) return HJ3D_OK;
  cudaSetDevice(cm->ctx->device);
  cudaStreamSynchronize(cm->ctx->stream);
  if (cm->ctx->copy_stream) cudaStreamSynchronize(cm->ctx->copy_stream);
  if (cm->xstream) cudaStreamSynchronize(cm->xstream);
  for (int s = 0; s < kSlots; ++s) {
    free_slot(cm, s);
    cudaFree(cm->d_cursor[s]); cudaFree(cm->d_all[s]); cudaFree(cm->d_pstart[s]);
    if (cm->ev_scatter[s]) cudaEventDestroy(cm->ev_scatter[s]);
    if (cm->ev_ready[s]) cudaEventDestroy(cm->ev_ready[s]);
    for (cudaEvent_t e : cm->chunk_ev[s]) cudaEventDestroy(e);
    cudaFree(cm->stage[s]);
    cudaFree(cm->d_sample[s]); cudaFree(cm->d_sample_all[s]); cudaFree(cm->d_hot_table[s]); cudaFree(cm->hot_buf[s]);
    cudaFree(cm->d_ans[s]); cudaFree(cm->d_ans_sum[s]);
    if (cm->ev_sample[s]) cudaEventDestroy(cm->ev_sample[s]);
    if (cm->ev_ans[s]) cudaEventDestroy(cm->ev_ans[s]);
  }
  if (cm->ev_fence) cudaEventDestroy(cm->ev_fence);
  if (cm->xstream) { cudaStreamSynchronize(cm->xstream); cudaStreamDestroy(cm->xstream); }
  cudaFree(cm->d_bar);
  if (cm->h_pinned) cudaFreeHost(cm->h_pinned);
  if (cm->nc) nccl().CommDestroy(cm->nc);
  if (cm->group) { auto& r = cm->group->ranks; for (auto& p : r) if (p == cm) p = nullptr; }
  delete cm;
  return HJ3D_OK;
}

// Collective: (re)allocate this rank's receive buffer of `slot` and map every peer's.
int hj3d_comm_reserve(hj3d_comm* cm, int slot, uint64_t records, uint32_t key_bytes) {
  if (!cm || slot < 0 || slot >= kSlots || (key_bytes != 4 && key_bytes != 8)) return fail(HJ3D_ERR_INVALID, "bad reserve arguments");
  hj3d_ctx* c = cm->ctx;
  CUDA_TRY(cudaSetDevice(c->device));
  const uint32_t rb = key_bytes == 8 ? 16 : 8;
  HJ_TRY(free_slot(cm, slot));
  if (records < 1024) records = 1024;
  HJ_TRY(raw_alloc(&cm->recv[slot], records * rb));
  cm->recv_records[slot] = records; cm->rec_bytes[slot] = rb;
  if (cm->nc) {
    // exchange CUDA IPC handles through the communicator itself: 64 bytes per rank
    cudaIpcMemHandle_t mine;
    CUDA_TRY(cudaIpcGetMemHandle(&mine, cm->recv[slot]));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    unsigned char* d_h = (unsigned char*)cm->d_all[slot];                 // staging: world * 64 bytes fit easily
    unsigned char* d_m = (unsigned char*)cm->d_cursor[slot];
    CUDA_TRY(cudaMemcpyAsync(d_m, &mine, 64, cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(nccl().AllGather(d_m, d_h, 64, ncclUint8, cm->nc, c->stream));
    std::vector<cudaIpcMemHandle_t> all(cm->world);
    CUDA_TRY(cudaMemcpyAsync(all.data(), d_h, (size_t)cm->world * 64, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < cm->world; ++r) {
      if (r == cm->rank) { cm->peer_recv[slot][r] = cm->recv[slot]; continue; }
      void* p = nullptr;
      CUDA_TRY(cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess));
      cm->peer_recv[slot][r] = p; cm->peer_is_ipc[slot][r] = true;
    }
  } else {
    // single process: every rank of the group sees every buffer; (re)publish to all members that exist
    for (hj3d_comm* o : cm->group->ranks) {
      if (!o) continue;
      o->peer_recv[slot][cm->rank] = cm->recv[slot];
      cm->peer_recv[slot][o->rank] = o->recv[slot];
    }
  }
  return HJ3D_OK;
}

int hj3d_comm_shard(hj3d_comm* cm, uint64_t D, uint64_t* lo, uint64_t* hi) {
  if (!cm || !D || D > 0xFFFFFFFFull || !lo || !hi) return fail(HJ3D_ERR_INVALID, "bad shard arguments");
  const ExchangePlan p = make_plan(cm, D);
  *lo = p.lo(cm->rank); *hi = p.hi(cm->rank);
  return HJ3D_OK;
}

int hj3d_exchange_begin_select(hj3d_comm* cm, int slot, const void* d_tuples, uint64_t n, hj3d_keyspec ks, uint64_t D, uint32_t rowid_base,
                               uint32_t flags, const hj3d_selection* sel);

int hj3d_exchange_begin(hj3d_comm* cm, int slot, const void* d_tuples, uint64_t n, hj3d_keyspec ks, uint64_t D, uint32_t rowid_base,
                        uint32_t flags) {
  return hj3d_exchange_begin_select(cm, slot, d_tuples, n, ks, D, rowid_base, flags, nullptr);
}

// what _begin, _begin_select and _begin_host share: argument checks, the plan and the uniform regions of this slot
static int prepare_slot(hj3d_comm* cm, int slot, uint64_t n, hj3d_keyspec ks, uint64_t D, uint32_t flags, const hj3d_selection* sel) {
  if (!cm || slot < 0 || slot >= kSlots) return fail(HJ3D_ERR_INVALID, "bad exchange arguments");
  if (!D || D > 0xFFFFFFFFull) return fail(HJ3D_ERR_INVALID, "bad num_buckets");
  if (n > 0xFFFFFFF0ull) return fail(HJ3D_ERR_UNSUPPORTED, "more than 2^32-16 tuples per rank");
  HJ_TRY(check_keyspec(ks));
  const uint32_t rb = ks.key_bytes == 8 ? 16 : 8;
  if (!cm->recv[slot] || cm->rec_bytes[slot] != rb) return fail(HJ3D_ERR_INVALID, "hj3d_comm_reserve has not been called for this slot / key width");
  for (int r = 0; r < cm->world; ++r) if (!cm->peer_recv[slot][r]) return fail(HJ3D_ERR_INVALID, "a peer has not reserved its receive buffer yet");
  if ((flags & HJ3D_XCHG_EXACT) && (flags & HJ3D_XCHG_MORE)) return fail(HJ3D_ERR_UNSUPPORTED, "HJ3D_XCHG_EXACT reads the slice twice: it cannot be streamed");
  if (sel && sel->op && (sel->op > 6 || sel->attr_offset % 4 || sel->attr_offset + 4 > ks.tuple_bytes)) return fail(HJ3D_ERR_INVALID, "bad selection");
  ExchangePlan& pl = cm->plan[slot];
  pl = make_plan(cm, D);
  cm->ks[slot] = ks; cm->n_local[slot] = n; cm->exact[slot] = (flags & HJ3D_XCHG_EXACT) != 0;
  // uniform regions: every (owned range, source) pair gets the same share of the owner's buffer
  uint64_t min_recv = cm->recv_records[slot];
  if (cm->group) for (hj3d_comm* o : cm->group->ranks) if (o && o->recv_records[slot] < min_recv) min_recv = o->recv_records[slot];
  cm->cap_seg[slot] = (min_recv / ((uint64_t)pl.rpo * cm->world)) & ~1ull;
  if (!cm->exact[slot] && cm->cap_seg[slot] < 2) return fail(HJ3D_ERR_INVALID, "receive buffer too small for the range x source regions");
  cm->sel[slot] = (sel && sel->op) ? *sel : hj3d_selection{0, 0, 0};
  cm->hot_on[slot] = false;
  if (flags & HJ3D_XCHG_HOT) {
    if (flags & HJ3D_XCHG_EXACT) return fail(HJ3D_ERR_UNSUPPORTED, "HJ3D_XCHG_HOT and HJ3D_XCHG_EXACT exclude each other");
    if (!cm->hot_sampled[slot]) return fail(HJ3D_ERR_INVALID, "HJ3D_XCHG_HOT needs hj3d_exchange_hot_sample on this slot first");
    cm->hot_on[slot] = true;
  }
  return HJ3D_OK;
}

// hot set of a slot: every rank runs k_hot_select on the same gathered sample (single process: gathered here)
static int hot_select(hj3d_comm* cm, int slot) {
  hj3d_ctx* c = cm->ctx;
  const uint32_t m_r = cm->sample_m[slot];
  uint32_t m = m_r * (uint32_t)cm->world;
  if (!cm->nc) {
    m = 0;
    for (hj3d_comm* o : cm->group->ranks) {
      if (!o) continue;
      if (!o->hot_sampled[slot] || o->sample_m[slot] != m_r) return fail(HJ3D_ERR_INVALID, "every rank of the group must call hj3d_exchange_hot_sample before the first hj3d_exchange_begin");
      CUDA_TRY(cudaStreamWaitEvent(c->stream, o->ev_sample[slot], 0));
      CUDA_TRY(cudaMemcpyAsync(cm->d_sample_all[slot] + (size_t)o->rank * m_r, o->d_sample[slot], (size_t)m_r * 8, cudaMemcpyDefault, c->stream));
      m += m_r;
    }
  }
  switch (cm->ks[slot].hash_id) {
    case HJ3D_HASH_MURMUR32: launch_hot_select<HJ3D_HASH_MURMUR32>(cm, slot, m, c->stream); break;
    case HJ3D_HASH_MURMUR64: launch_hot_select<HJ3D_HASH_MURMUR64>(cm, slot, m, c->stream); break;
    default:                 launch_hot_select<HJ3D_HASH_MURMUR64_SEXT32>(cm, slot, m, c->stream); break;
  }
  ++c->launches;
  CUDA_TRY(cudaGetLastError());
  cm->hot_selected[slot] = true;
  return HJ3D_OK;
}

// Sample the slot's relation for hot keys (collective; before hj3d_exchange_begin(.., HJ3D_XCHG_HOT)).
int hj3d_exchange_hot_sample(hj3d_comm* cm, int slot, const void* d_tuples, uint64_t n, hj3d_keyspec ks) {
  if (!cm || slot < 0 || slot >= kSlots) return fail(HJ3D_ERR_INVALID, "bad exchange arguments");
  if (n && !d_tuples) return fail(HJ3D_ERR_INVALID, "d_tuples == NULL");
  HJ_TRY(check_keyspec(ks));
  hj3d_ctx* c = cm->ctx;
  CUDA_TRY(cudaSetDevice(c->device));
  const uint32_t m_r = (uint32_t)kHotSample / (uint32_t)cm->world;
  const Src src = make_src(d_tuples, n, ks, nullptr);
  if (ks.key_bytes == 8) k_hot_sample<uint64_t><<<blocks_for(m_r, 256), 256, 0, c->stream>>>(src, m_r, cm->d_sample[slot]);
  else                   k_hot_sample<uint32_t><<<blocks_for(m_r, 256), 256, 0, c->stream>>>(src, m_r, cm->d_sample[slot]);
  ++c->launches;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(cm->ev_sample[slot], c->stream));
  cm->sample_m[slot] = m_r; cm->ks[slot] = ks;
  cm->hot_sampled[slot] = true; cm->hot_selected[slot] = false;
  if (cm->nc) {
    NCCL_TRY(nccl().AllGather(cm->d_sample[slot], cm->d_sample_all[slot], m_r, ncclUint64, cm->nc, c->stream));
    HJ_TRY(hot_select(cm, slot));
  }
  return HJ3D_OK;
}

// this rank's answers buffer (engine.cu fills it), and the all-reduce of the answers
void* hj3d_comm_hot_ans_buffer(hj3d_comm* cm, int slot) { return cm->d_ans[slot]; }
int hj3d_comm_hot_reduce_begin(hj3d_comm* cm, int slot) {
  hj3d_ctx* c = cm->ctx;
  constexpr size_t n32 = sizeof(HotAnswers) / 4;
  if (cm->nc) NCCL_TRY(nccl().AllReduce(cm->d_ans[slot], cm->d_ans_sum[slot], n32, ncclUint32, ncclSum, cm->nc, c->stream));
  else CUDA_TRY(cudaEventRecord(cm->ev_ans[slot], c->stream));
  return HJ3D_OK;
}
int hj3d_comm_hot_reduce_end(hj3d_comm* cm, int slot, const void** d_sum) {
  hj3d_ctx* c = cm->ctx;
  constexpr uint32_t n32 = (uint32_t)(sizeof(HotAnswers) / 4);
  if (!cm->nc) {
    CUDA_TRY(cudaMemsetAsync(cm->d_ans_sum[slot], 0, sizeof(HotAnswers), c->stream));
    for (hj3d_comm* o : cm->group->ranks) {
      if (!o) continue;
      CUDA_TRY(cudaStreamWaitEvent(c->stream, o->ev_ans[slot], 0));
      k_add_u32<<<blocks_for(n32, 256), 256, 0, c->stream>>>((uint32_t*)cm->d_ans_sum[slot], (const uint32_t*)o->d_ans[slot], n32);
      ++c->launches;
    }
    CUDA_TRY(cudaGetLastError());
  }
  *d_sum = cm->d_ans_sum[slot];
  return HJ3D_OK;
}

static Src slot_src(const hj3d_comm* cm, int slot, const void* d_tuples, uint64_t n) {
  Src src = make_src(d_tuples, n, cm->ks[slot], nullptr);
  src.sel_off = cm->sel[slot].attr_offset; src.sel_op = cm->sel[slot].op; src.sel_cst = cm->sel[slot].constant;
  return src;
}

static int scatter_by_hash(hj3d_comm* cm, int slot, Src src, uint32_t rowid_base, unsigned long long cap, cudaStream_t st = nullptr) {
  switch (cm->ks[slot].hash_id) {
    case HJ3D_HASH_MURMUR32: return launch_scatter<HJ3D_HASH_MURMUR32>(cm, slot, src, rowid_base, cap, st);
    case HJ3D_HASH_MURMUR64: return launch_scatter<HJ3D_HASH_MURMUR64>(cm, slot, src, rowid_base, cap, st);
    default:                 return launch_scatter<HJ3D_HASH_MURMUR64_SEXT32>(cm, slot, src, rowid_base, cap, st);
  }
}

int hj3d_exchange_begin_select(hj3d_comm* cm, int slot, const void* d_tuples, uint64_t n, hj3d_keyspec ks, uint64_t D, uint32_t rowid_base,
                               uint32_t flags, const hj3d_selection* sel) {
  if (n && !d_tuples) return fail(HJ3D_ERR_INVALID, "d_tuples == NULL");
  HJ_TRY(prepare_slot(cm, slot, n, ks, D, flags, sel));
  hj3d_ctx* c = cm->ctx;
  CUDA_TRY(cudaSetDevice(c->device));
  const ExchangePlan& pl = cm->plan[slot];
  const Src src = slot_src(cm, slot, d_tuples, n);
  cm->host_streamed[slot] = false;
  cm->async_on[slot] = (flags & HJ3D_XCHG_ASYNC) != 0;
  if (cm->async_on[slot] && (flags & (HJ3D_XCHG_EXACT | HJ3D_XCHG_MORE))) return fail(HJ3D_ERR_UNSUPPORTED, "HJ3D_XCHG_ASYNC goes with the single-pass exchange only");
  if (cm->hot_on[slot]) {
    if (flags & HJ3D_XCHG_MORE) return fail(HJ3D_ERR_UNSUPPORTED, "HJ3D_XCHG_HOT cannot be streamed");
    if (cm->hot_buf_records[slot] < n) {            // worst case: every local tuple is hot
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      if (cm->xstream) CUDA_TRY(cudaStreamSynchronize(cm->xstream));
      cudaFree(cm->hot_buf[slot]); cm->hot_buf[slot] = nullptr; cm->hot_buf_records[slot] = 0;
      HJ_TRY(raw_alloc(&cm->hot_buf[slot], (n ? n : 1) * (size_t)(ks.key_bytes == 8 ? 16 : 8)));
      cm->hot_buf_records[slot] = n ? n : 1;
    }
    if (!cm->hot_selected[slot]) HJ_TRY(hot_select(cm, slot));
  }
  // HJ3D_XCHG_ASYNC: everything below runs on the communicator's own stream, ordered after what the ctx stream holds now;
  // the ctx stream is free for other work (building the table of the other relation) until hj3d_exchange_end joins it
  cudaStream_t st = c->stream;
  if (cm->async_on[slot]) {
    if (!cm->xstream) CUDA_TRY(cudaStreamCreateWithFlags(&cm->xstream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventRecord(cm->ev_fence, c->stream));
    CUDA_TRY(cudaStreamWaitEvent(cm->xstream, cm->ev_fence, 0));
    st = cm->xstream;
  }
  std::unique_ptr<PhaseTimer> pt;                    // (the phase events live on the ctx stream)
  if (!cm->async_on[slot]) pt.reset(new PhaseTimer(c, PH_PARTITION));
  CUDA_TRY(cudaMemsetAsync(cm->d_cursor[slot], 0, (kMaxRanges + 8) * 8, st));
  int rc = HJ3D_OK;
  if (!cm->exact[slot]) {
    k_xchg_starts<<<blocks_for(pl.n_ranges + 1, 256), 256, 0, st>>>(pl.n_ranges, pl.rpo_shift, cm->world, cm->rank, cm->cap_seg[slot], cm->d_pstart[slot]);
    ++c->launches;
    rc = scatter_by_hash(cm, slot, src, rowid_base, cm->cap_seg[slot], st);
  } else {
    switch (ks.hash_id) {
      case HJ3D_HASH_MURMUR32: rc = launch_hist<HJ3D_HASH_MURMUR32>(cm, slot, src); break;
      case HJ3D_HASH_MURMUR64: rc = launch_hist<HJ3D_HASH_MURMUR64>(cm, slot, src); break;
      default:                 rc = launch_hist<HJ3D_HASH_MURMUR64_SEXT32>(cm, slot, src); break;
    }
  }
  if (rc < 0) return rc;
  cm->pending[slot] = true;
  cm->streaming[slot] = (flags & HJ3D_XCHG_MORE) != 0;
  if (cm->streaming[slot]) return HJ3D_OK;           // the counts leave with the last chunk
  CUDA_TRY(cudaEventRecord(cm->ev_scatter[slot], st));
  if (cm->nc) HJ_TRY(gather_counts(cm, slot, st));   // multi-process: enqueue the all-gather right behind the kernel
  if (cm->async_on[slot]) CUDA_TRY(cudaEventRecord(cm->ev_ready[slot], st));
  return HJ3D_OK;
}

// The local slice lives in HOST memory.  It is uploaded in chunks on the ctx's copy stream and every chunk goes through
// level 1 on the comm's exchange stream as soon as it has landed, under the upload of the chunks behind it.  The caller's
// stream is not touched before hj3d_exchange_end, so what it queues meanwhile (the build of the other relation) overlaps.
int hj3d_exchange_begin_host(hj3d_comm* cm, int slot, const void* h_tuples, uint64_t n, hj3d_keyspec ks, uint64_t D, uint32_t rowid_base,
                             uint32_t flags, const hj3d_selection* sel) {
  if (n && !h_tuples) return fail(HJ3D_ERR_INVALID, "h_tuples == NULL");
  if (flags & (HJ3D_XCHG_EXACT | HJ3D_XCHG_MORE | HJ3D_XCHG_HOT)) return fail(HJ3D_ERR_UNSUPPORTED, "hj3d_exchange_begin_host streams the slice once: no HJ3D_XCHG_EXACT / _MORE / _HOT");
  HJ_TRY(prepare_slot(cm, slot, n, ks, D, flags, sel));
  if ((uint64_t)rowid_base + n > 0xFFFFFFFFull) return fail(HJ3D_ERR_UNSUPPORTED, "row ids past 2^32");
  hj3d_ctx* c = cm->ctx;
  CUDA_TRY(cudaSetDevice(c->device));
  if (!c->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  if (!cm->xstream) CUDA_TRY(cudaStreamCreateWithFlags(&cm->xstream, cudaStreamNonBlocking));
  const ExchangePlan& pl = cm->plan[slot];
  const size_t bytes = (size_t)n * ks.tuple_bytes;
  if (cm->stage_bytes[slot] < bytes) {
    CUDA_TRY(cudaStreamSynchronize(c->stream)); CUDA_TRY(cudaStreamSynchronize(cm->xstream)); CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
    cudaFree(cm->stage[slot]); cm->stage[slot] = nullptr; cm->stage_bytes[slot] = 0;
    HJ_TRY(raw_alloc(&cm->stage[slot], bytes));
    cm->stage_bytes[slot] = bytes;
  }
  uint64_t chunk_rows = c->host_chunk_bytes > 0 ? std::max<uint64_t>(1, (uint64_t)c->host_chunk_bytes / ks.tuple_bytes) : n;
  chunk_rows = std::max<uint64_t>(chunk_rows, (n + 4095) / 4096);                       // at most 4096 chunks
  const size_t n_chunks = n ? (size_t)((n + chunk_rows - 1) / chunk_rows) : 0;
  while (cm->chunk_ev[slot].size() < n_chunks) {
    cudaEvent_t e; CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); cm->chunk_ev[slot].push_back(e);
  }
  // both side streams start where the caller's stream is now (the buffers of an earlier exchange may still be read there)
  CUDA_TRY(cudaEventRecord(cm->ev_fence, c->stream));
  CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, cm->ev_fence, 0));
  CUDA_TRY(cudaStreamWaitEvent(cm->xstream, cm->ev_fence, 0));
  CUDA_TRY(cudaMemsetAsync(cm->d_cursor[slot], 0, (kMaxRanges + 8) * 8, cm->xstream));
  k_xchg_starts<<<blocks_for(pl.n_ranges + 1, 256), 256, 0, cm->xstream>>>(pl.n_ranges, pl.rpo_shift, cm->world, cm->rank, cm->cap_seg[slot], cm->d_pstart[slot]);
  ++c->launches;
  for (size_t i = 0; i < n_chunks; ++i) {
    const uint64_t r0 = i * chunk_rows, rn = std::min<uint64_t>(chunk_rows, n - r0);
    uint8_t* d_chunk = (uint8_t*)cm->stage[slot] + r0 * ks.tuple_bytes;
    CUDA_TRY(cudaMemcpyAsync(d_chunk, (const uint8_t*)h_tuples + r0 * ks.tuple_bytes, rn * ks.tuple_bytes, cudaMemcpyHostToDevice, c->copy_stream));
    CUDA_TRY(cudaEventRecord(cm->chunk_ev[slot][i], c->copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(cm->xstream, cm->chunk_ev[slot][i], 0));
    HJ_TRY(scatter_by_hash(cm, slot, slot_src(cm, slot, d_chunk, rn), (uint32_t)(rowid_base + r0), cm->cap_seg[slot], cm->xstream));
  }
  CUDA_TRY(cudaEventRecord(cm->ev_scatter[slot], cm->xstream));
  if (cm->nc) HJ_TRY(gather_counts(cm, slot, cm->xstream));
  CUDA_TRY(cudaEventRecord(cm->ev_ready[slot], cm->xstream));
  cm->n_chunks[slot] = n_chunks;
  cm->pending[slot] = true; cm->streaming[slot] = false; cm->host_streamed[slot] = true;
  return HJ3D_OK;
}

// A further chunk of the local slice (streamed upload: a chunk is partitioned while the next one is still on its way
// from the host).  The per-range cursors simply keep counting, so the chunks of one source share its regions.
int hj3d_exchange_append(hj3d_comm* cm, int slot, const void* d_tuples, uint64_t n, uint32_t rowid_base, uint32_t flags) {
  if (!cm || slot < 0 || slot >= kSlots) return fail(HJ3D_ERR_INVALID, "bad exchange arguments");
  if (!cm->pending[slot] || !cm->streaming[slot]) return fail(HJ3D_ERR_INVALID, "hj3d_exchange_append needs a hj3d_exchange_begin with HJ3D_XCHG_MORE");
  if (n && !d_tuples) return fail(HJ3D_ERR_INVALID, "d_tuples == NULL");
  if (cm->n_local[slot] + n > 0xFFFFFFF0ull) return fail(HJ3D_ERR_UNSUPPORTED, "more than 2^32-16 tuples per rank");
  hj3d_ctx* c = cm->ctx;
  CUDA_TRY(cudaSetDevice(c->device));
  const int rc = scatter_by_hash(cm, slot, slot_src(cm, slot, d_tuples, n), rowid_base, cm->cap_seg[slot]);
  if (rc < 0) return rc;
  cm->n_local[slot] += n;
  if (flags & HJ3D_XCHG_MORE) return HJ3D_OK;
  cm->streaming[slot] = false;
  CUDA_TRY(cudaEventRecord(cm->ev_scatter[slot], c->stream));
  if (cm->nc) HJ_TRY(gather_counts(cm, slot));
  return HJ3D_OK;
}

// second pass of the exact mode: offsets from the gathered histogram, scatter, barrier
static int exact_second_pass(hj3d_comm* cm, int slot, const void* d_tuples, uint32_t rowid_base, unsigned long long* h_all) {
  hj3d_ctx* c = cm->ctx;
  const ExchangePlan& pl = cm->plan[slot];
  CUDA_TRY(cudaMemcpyAsync(h_all, cm->d_all[slot], (size_t)cm->world * kMaxRanges * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  unsigned long long* h_ps = h_all + (size_t)cm->world * kMaxRanges;
  uint64_t worst = 0;
  for (int o = 0; o < cm->world; ++o) {                // owner o's buffer: ranges in order, sources in order, tightly packed
    unsigned long long run = 0;
    for (uint32_t p = 0; p < pl.owned(o); ++p) {
      const uint32_t q = o * pl.rpo + p;
      for (int s = 0; s < cm->world; ++s) { if (s == cm->rank) h_ps[q] = run; run += h_all[(size_t)s * kMaxRanges + q]; }
    }
    if (run > worst) worst = run;
  }
  uint64_t min_recv = cm->recv_records[slot];
  if (cm->group) for (hj3d_comm* o : cm->group->ranks) if (o && o->recv_records[slot] < min_recv) min_recv = o->recv_records[slot];
  if (worst > min_recv) return fail(HJ3D_ERR_NOMEM, "receive buffer too small for the exact exchange (a rank would receive " + std::to_string(worst) + " records)");
  h_ps[pl.n_ranges] = 0;
  CUDA_TRY(cudaMemcpyAsync(cm->d_pstart[slot], h_ps, ((size_t)pl.n_ranges + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemsetAsync(cm->d_cursor[slot], 0, (kMaxRanges + 8) * 8, c->stream));
  const int rc = scatter_by_hash(cm, slot, slot_src(cm, slot, d_tuples, cm->n_local[slot]), rowid_base, ~0ull);
  if (rc < 0) return rc;
  CUDA_TRY(cudaEventRecord(cm->ev_scatter[slot], c->stream));
  return HJ3D_OK;
}

int hj3d_exchange_end(hj3d_comm* cm, int slot, const void* d_tuples, uint32_t rowid_base, uint64_t rowid_bound, hj3d_parts** out) {
  if (!cm || slot < 0 || slot >= kSlots || !out) return fail(HJ3D_ERR_INVALID, "bad exchange arguments");
  if (!cm->pending[slot]) return fail(HJ3D_ERR_INVALID, "hj3d_exchange_begin has not been called for this slot");
  *out = nullptr;
  hj3d_ctx* c = cm->ctx;
  CUDA_TRY(cudaSetDevice(c->device));
  const ExchangePlan& pl = cm->plan[slot];
  unsigned long long* h_all = (unsigned long long*)cm->h_pinned;
  if (cm->host_streamed[slot]) {
    // Wait for the upload on the HOST before anything is queued on the caller's stream: the small copies below would sit
    // behind the chunk events and share the copy engine's queue with the chunks in flight -- measured, that time-slices
    // the upload down to 17..47 GB/s instead of 55.  Nothing is lost: this call synchronises anyway.
    if (cm->n_chunks[slot]) CUDA_TRY(cudaEventSynchronize(cm->chunk_ev[slot][cm->n_chunks[slot] - 1]));
    CUDA_TRY(cudaStreamWaitEvent(c->stream, cm->ev_ready[slot], 0));
    cm->host_streamed[slot] = false;
  }
  if (cm->async_on[slot]) {
    CUDA_TRY(cudaStreamWaitEvent(c->stream, cm->ev_ready[slot], 0));
    cm->async_on[slot] = false;
  }
  if (cm->streaming[slot]) {                         // ended without a final append: the chunks so far are the slice
    cm->streaming[slot] = false;
    CUDA_TRY(cudaEventRecord(cm->ev_scatter[slot], c->stream));
    if (cm->nc) HJ_TRY(gather_counts(cm, slot));
  }
  if (!cm->nc) HJ_TRY(gather_counts(cm, slot));
  if (cm->exact[slot]) {
    if (cm->group && cm->group->ranks.size() > 1)
      return fail(HJ3D_ERR_UNSUPPORTED, "HJ3D_XCHG_EXACT needs one process per rank (or a group of one): the second pass is a collective");
    HJ_TRY(exact_second_pass(cm, slot, d_tuples, rowid_base, h_all));
    // barrier: everybody's second pass is done before anybody reads (the gathered cursors are not needed)
    if (cm->nc) NCCL_TRY(nccl().AllGather(cm->d_bar, cm->d_bar + 1, 1, ncclUint64, cm->nc, c->stream));
  }
  const uint32_t n_owned = pl.owned(cm->rank), n_seg = n_owned * cm->world;
  auto parts = std::make_unique<hj3d_parts>();
  parts->recs = cm->recv[slot]; parts->key_bytes = cm->ks[slot].key_bytes; parts->hash_id = cm->ks[slot].hash_id;
  parts->D = pl.D; parts->bucket_lo = pl.lo(cm->rank); parts->bucket_hi = pl.hi(cm->rank);
  parts->n_ranges = n_owned; parts->n_src = cm->world; parts->range_width = pl.width; parts->cap_seg = cm->cap_seg[slot];
  parts->rowid_bound = rowid_bound;
  CUDA_TRY(cudaMalloc((void**)&parts->d_start, ((size_t)n_seg + 1) * 8));
  CUDA_TRY(cudaMalloc((void**)&parts->d_count, ((size_t)n_seg + 1) * 8));
  k_xchg_segments<<<1, 32, 0, c->stream>>>(cm->d_all[slot], kMaxRanges, cm->world, cm->rank * pl.rpo, n_owned, cm->cap_seg[slot], cm->exact[slot] ? 1 : 0,
                                           parts->d_start, parts->d_count);
  ++c->launches;
  if (!cm->exact[slot]) CUDA_TRY(cudaMemcpyAsync(h_all, cm->d_all[slot], (size_t)cm->world * kMaxRanges * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  // host view: what I received, what I sent to others, overflow of any region anybody wrote (every rank sees all counts)
  uint64_t recv = 0, sent = 0, mine = 0; int overflow = 0;
  for (int s = 0; s < cm->world; ++s)
    for (uint32_t q = 0; q < pl.n_ranges; ++q) {
      const uint64_t cnt = h_all[(size_t)s * kMaxRanges + q];
      const int owner = (int)(q >> pl.rpo_shift);
      if (!cm->exact[slot] && cnt > cm->cap_seg[slot]) overflow = 1;
      const uint64_t stored = (!cm->exact[slot] && cnt > cm->cap_seg[slot]) ? cm->cap_seg[slot] : cnt;
      if (owner == cm->rank) recv += stored;
      if (s == cm->rank && owner != cm->rank) sent += stored;
      if (s == cm->rank) mine += cnt;
    }
  parts->n_total = recv; parts->n_sent_remote = sent; parts->n_local_selected = mine; parts->overflow = overflow;
  if (cm->hot_on[slot]) {
    unsigned long long* h_hot = h_all + (size_t)cm->world * kMaxRanges;             // my own count of the extra partition
    CUDA_TRY(cudaMemcpyAsync(h_hot, cm->d_cursor[slot] + pl.n_ranges, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const uint64_t hot = *h_hot;
    parts->hot_recs = cm->hot_buf[slot]; parts->hot_count = hot; parts->hot_table = cm->d_hot_table[slot];
    parts->comm = cm; parts->slot = slot; parts->hot_mode = -1;
    parts->n_local_selected += hot;
    cm->hot_sampled[slot] = false;                   // a sample serves one exchange
  }
  cm->pending[slot] = false;
  *out = parts.release();
  return overflow ? HJ3D_OVERFLOW : HJ3D_OK;
}

int hj3d_parts_selected(hj3d_parts* p, uint64_t* n_local_selected) {
  if (!p || !n_local_selected) return fail(HJ3D_ERR_INVALID, "NULL argument");
  *n_local_selected = p->n_local_selected;
  return HJ3D_OK;
}

int hj3d_parts_info(hj3d_parts* p, uint64_t* n_records, uint64_t* n_sent_remote, uint64_t* bucket_lo, uint64_t* bucket_hi, int* overflow) {
  if (!p) return fail(HJ3D_ERR_INVALID, "NULL argument");
  if (n_records) *n_records = p->n_total;
  if (n_sent_remote) *n_sent_remote = p->n_sent_remote;
  if (bucket_lo) *bucket_lo = p->bucket_lo;
  if (bucket_hi) *bucket_hi = p->bucket_hi;
  if (overflow) *overflow = p->overflow;
  return HJ3D_OK;
}

int hj3d_parts_destroy(hj3d_parts* p) {
  if (!p) return HJ3D_OK;
  delete p;
  return HJ3D_OK;
}

}  // extern "C"
