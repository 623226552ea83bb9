// peer.cu -- one-sided voxel exchange over NVLink peer memory (the reference is single-process).
//
// Every rank owns an INBOX, a cudaMalloc block that the other ranks of the node map through CUDA IPC.  After its
// submaps are fused, a rank PUSHES: one kernel groups the local voxels by owner (mix64(key) % world), reserves
// slots in the owners' inboxes with one remote atomic per (CTA, owner) and stores the records (key, count, fp32
// sums) straight into peer memory with 16-byte stores over NVLink -- pack and transfer are the same pass, no
// send buffer, no collective.  Contributor-log entries follow the same way; a last kernel raises every peer's
// `arrived` counter.  The owner DRAINS: a one-thread kernel waits for `arrived == world`, then the records are
// inserted into the owner's map (hash insert + vector REDs), all on the owner's stream without the host.
//
// Inbox block:  2 headers (256 B each) | 2 row regions (cap_rows x (16 + 4d) B) | 2 contributor regions
// (cap_contrib x 32 B).  Exchanges alternate between the two halves (parity = epoch & 1): the owner resets a half
// at the end of its drain, and a peer can only reach that half again after it has drained the exchange in between,
// which needs the owner's own push of that exchange -- queued behind the reset on the owner's stream.
#include <algorithm>
#include <cstring>

#include "hash.cuh"

namespace vsm {

constexpr int kMaxWorld = 64;
constexpr uint32_t kInboxRowOverflow = 1u, kInboxContribOverflow = 2u, kInboxTimeout = 4u, kInboxHashErr = 8u;

struct InboxHeader {
  uint32_t n_rows;     // row slots reserved by the senders (can exceed the capacity: overflow)
  uint32_t n_contrib;  // contributor slots reserved
  uint32_t arrived;    // senders that have finished pushing
  uint32_t flags;
  uint32_t pad[60];
};
static_assert(sizeof(InboxHeader) == 256, "inbox header is 256 bytes");

struct InboxLayout {
  size_t row_stride, rows_off[2], contrib_off[2], total;
};
static InboxLayout inbox_layout(int d, int64_t cap_rows, int64_t cap_contrib) {
  InboxLayout L;
  L.row_stride = 16 + (size_t)d * 4;
  size_t off = 2 * sizeof(InboxHeader);
  for (int h = 0; h < 2; ++h) {
    L.rows_off[h] = off;
    off += (size_t)cap_rows * L.row_stride;
    off = (off + 255) & ~(size_t)255;
  }
  for (int h = 0; h < 2; ++h) {
    L.contrib_off[h] = off;
    off += (size_t)cap_contrib * 32;
    off = (off + 255) & ~(size_t)255;
  }
  L.total = off;
  return L;
}

__host__ __device__ __forceinline__ uint32_t owner_of_key(unsigned long long key, uint32_t world) {
  return (uint32_t)((mix64(key ^ 0x9E3779B97F4A7C15ull) >> 32) % world);  // same owner as exchange.cu
}

struct PeerPtrs {
  InboxHeader* hdr[kMaxWorld];
  uint8_t* rows[kMaxWorld];
  uint8_t* contrib[kMaxWorld];
};

__device__ __forceinline__ void st_v4(void* p, uint4 v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// One CTA takes tiles of 256 voxels: per-owner counts in shared memory, one remote atomic per owner reserves the
// tile's slots, then every warp copies whole records (129 x 16 B at d = 512) into the owners' inboxes.
__global__ void __launch_bounds__(256) push_rows_kernel(const unsigned long long* __restrict__ vkey,
                                                        const uint32_t* __restrict__ vcount, const float* __restrict__ vsum,
                                                        uint32_t n, int d, uint32_t world, PeerPtrs peers,
                                                        uint32_t cap_rows, size_t row_stride) {
  __shared__ uint32_t s_cnt[kMaxWorld], s_base[kMaxWorld];
  __shared__ uint32_t s_slot[256];
  __shared__ uint8_t s_owner[256];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int nvec = d / 4;
  for (uint32_t tile = blockIdx.x * 256u; tile < n; tile += gridDim.x * 256u) {
    if (threadIdx.x < kMaxWorld) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t i = tile + threadIdx.x;
    uint32_t o = 0, r = 0;
    if (i < n) {
      o = owner_of_key(vkey[i], world);
      r = atomicAdd(&s_cnt[o], 1u);
    }
    __syncthreads();
    if (threadIdx.x < world && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd_system(&peers.hdr[threadIdx.x]->n_rows, s_cnt[threadIdx.x]);
    __syncthreads();
    s_slot[threadIdx.x] = (i < n) ? s_base[o] + r : 0xFFFFFFFFu;
    s_owner[threadIdx.x] = (uint8_t)o;
    __syncthreads();
    const uint32_t in_tile = min(256u, n - tile);
    for (uint32_t j = warp; j < in_tile; j += 8) {
      const uint32_t slot = s_slot[j], ow = s_owner[j], src = tile + j;
      if (slot >= cap_rows) {
        if (lane == 0) atomicOr_system(&peers.hdr[ow]->flags, kInboxRowOverflow);
        continue;
      }
      uint8_t* dst = peers.rows[ow] + (size_t)slot * row_stride;
      const uint4* srow = reinterpret_cast<const uint4*>(vsum + (size_t)src * d);
      if (lane == 0) {
        const unsigned long long k = vkey[src];
        st_v4(dst, make_uint4((uint32_t)k, (uint32_t)(k >> 32), vcount[src], 0u));
      }
      for (int c = lane; c < nvec; c += 32) st_v4(dst + 16 + (size_t)c * 16, srow[c]);
    }
    __syncthreads();
  }
  __threadfence_system();  // this thread's peer stores are performed before the kernel (and the signal behind it) ends
}

__global__ void __launch_bounds__(256) push_contrib_kernel(const int32_t* __restrict__ log_gid, const int32_t* __restrict__ log_sub,
                                                           const unsigned long long* __restrict__ log_mask,
                                                           const unsigned long long* __restrict__ vkey, uint32_t n,
                                                           uint32_t world, PeerPtrs peers, uint32_t cap_contrib) {
  __shared__ uint32_t s_cnt[kMaxWorld], s_base[kMaxWorld];
  for (uint32_t tile = blockIdx.x * 256u; tile < n; tile += gridDim.x * 256u) {
    if (threadIdx.x < kMaxWorld) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t e = tile + threadIdx.x;
    uint32_t o = 0, r = 0;
    unsigned long long key = 0ull;
    bool act = false;
    if (e < n) {
      const int g = log_gid[e];
      if (g >= 0) {
        act = true;
        key = vkey[g];
        o = owner_of_key(key, world);
        r = atomicAdd(&s_cnt[o], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x < world && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd_system(&peers.hdr[threadIdx.x]->n_contrib, s_cnt[threadIdx.x]);
    __syncthreads();
    if (act) {
      const uint32_t slot = s_base[o] + r;
      if (slot >= cap_contrib) {
        atomicOr_system(&peers.hdr[o]->flags, kInboxContribOverflow);
      } else {
        uint8_t* dst = peers.contrib[o] + (size_t)slot * 32;
        const unsigned long long m0 = log_mask[2 * (size_t)e], m1 = log_mask[2 * (size_t)e + 1];
        st_v4(dst, make_uint4((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)log_sub[e], 0u));
        st_v4(dst + 16, make_uint4((uint32_t)m0, (uint32_t)(m0 >> 32), (uint32_t)m1, (uint32_t)(m1 >> 32)));
      }
    }
    __syncthreads();
  }
  __threadfence_system();
}

// runs behind the two push kernels on the sender's stream: their stores are complete, tell every owner
__global__ void push_signal_kernel(PeerPtrs peers, uint32_t world) {
  if (threadIdx.x < world) {
    __threadfence_system();
    atomicAdd_system(&peers.hdr[threadIdx.x]->arrived, 1u);
  }
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// owner side: wait until every sender has signalled (bounded: a dead peer must not hang the GPU)
__global__ void inbox_wait_kernel(InboxHeader* hdr, uint32_t world, unsigned long long timeout_ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (ld_acquire_sys(&hdr->arrived) < world) {
    __nanosleep(200);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) {
      atomicOr(&hdr->flags, kInboxTimeout);
      break;
    }
  }
  __threadfence_system();
}

__global__ void __launch_bounds__(256) drain_keys_kernel(GlobalStore g, const InboxHeader* hdr, const uint8_t* __restrict__ rows,
                                                         size_t row_stride, uint32_t cap_rows, int32_t* __restrict__ gid_out,
                                                         uint32_t* err) {
  const uint32_t n = min(hdr->n_rows, cap_rows);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint4 h = *reinterpret_cast<const uint4*>(rows + (size_t)i * row_stride);
    const unsigned long long key = (unsigned long long)h.x | ((unsigned long long)h.y << 32);
    const int gid = global_find_or_insert(g, key, err);
    gid_out[i] = gid;
    if (gid >= 0) atomicAdd(&g.vcount[gid], h.z);
  }
}

__global__ void __launch_bounds__(256) drain_rows_kernel(const InboxHeader* hdr, const uint8_t* __restrict__ rows,
                                                         size_t row_stride, uint32_t cap_rows, const int32_t* __restrict__ gid,
                                                         int d, float* __restrict__ vsum) {
  const uint32_t n = min(hdr->n_rows, cap_rows);
  const int lane = lane_id();
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = warp; i < n; i += n_warps) {
    const int g = gid[i];
    if (g < 0) continue;
    const uint8_t* src = rows + (size_t)i * row_stride + 16;
    float* dst = vsum + (size_t)g * d;
    for (int c = lane; c < d / 4; c += 32) {
      const uint4 v = ld_stream_v4(src + (size_t)c * 16);
      red_add_v4(dst + 4 * c, __uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
    }
  }
}

__global__ void __launch_bounds__(256) drain_contrib_kernel(GlobalStore g, const InboxHeader* hdr,
                                                            const uint8_t* __restrict__ contrib, uint32_t cap_contrib,
                                                            const uint32_t* __restrict__ map_state, uint32_t log_cap,
                                                            int32_t* __restrict__ log_gid, int32_t* __restrict__ log_sub,
                                                            unsigned long long* __restrict__ log_mask) {
  const uint32_t n = min(hdr->n_contrib, cap_contrib);
  const uint32_t log_base = map_state[1];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (log_base + i >= log_cap) break;
    const uint4 a = *reinterpret_cast<const uint4*>(contrib + (size_t)i * 32);
    const uint4 b = *reinterpret_cast<const uint4*>(contrib + (size_t)i * 32 + 16);
    const unsigned long long key = (unsigned long long)a.x | ((unsigned long long)a.y << 32);
    int gid = -1;
    uint64_t h = mix64(key) & g.gmask;
    for (uint64_t probes = 0; probes <= g.gmask; ++probes) {
      const unsigned long long cur = g.gkeys[h];
      if (cur == kEmptyKey) break;
      if (cur == key) {
        gid = g.gids[h];
        break;
      }
      h = (h + 1) & g.gmask;
    }
    const size_t e = (size_t)log_base + i;
    log_gid[e] = gid;  // -1 entries are skipped by vsm_finalize
    log_sub[e] = (int32_t)a.z;
    log_mask[2 * e] = (unsigned long long)b.x | ((unsigned long long)b.y << 32);
    log_mask[2 * e + 1] = (unsigned long long)b.z | ((unsigned long long)b.w << 32);
  }
}

// last kernel of a drain: report, advance the map's log length, and reset this half of the inbox for its next use
__global__ void drain_finish_kernel(InboxHeader* hdr, uint32_t cap_contrib, uint32_t* map_state, uint32_t log_cap,
                                    const uint32_t* hash_err, uint32_t* report /* [n_rows, n_contrib, flags, n_vox, hash err, log n] */) {
  uint32_t flags = hdr->flags;
  if (*hash_err) flags |= kInboxHashErr;
  const uint32_t n_contrib = min(hdr->n_contrib, cap_contrib);
  if ((unsigned long long)map_state[1] + n_contrib > log_cap) flags |= kInboxContribOverflow;
  report[0] += hdr->n_rows;  // reports accumulate over the drains queued since the slot was last collected
  report[1] += hdr->n_contrib;
  report[2] |= flags;
  map_state[1] = min(map_state[1] + n_contrib, log_cap);
  report[3] = map_state[0];
  report[5] = map_state[1];  // (report[4] is the hash-error counter itself)
  hdr->n_rows = 0u;
  hdr->n_contrib = 0u;
  hdr->flags = 0u;
  __threadfence_system();
  hdr->arrived = 0u;
}

static PeerPtrs peer_ptrs(void* const* inbox_ptrs, int world, const InboxLayout& L, int half) {
  PeerPtrs p{};
  for (int o = 0; o < world; ++o) {
    uint8_t* base = reinterpret_cast<uint8_t*>(inbox_ptrs[o]);
    p.hdr[o] = reinterpret_cast<InboxHeader*>(base) + half;
    p.rows[o] = base + L.rows_off[half];
    p.contrib[o] = base + L.contrib_off[half];
  }
  return p;
}

}  // namespace vsm

using namespace vsm;

extern "C" int vsm_inbox_bytes(int32_t dim, int64_t cap_rows, int64_t cap_contrib, int64_t* bytes_host) {
  if (dim <= 0 || dim % 4 || cap_rows < 1 || cap_contrib < 1 || cap_rows >= ((int64_t)1 << 31) ||
      cap_contrib >= ((int64_t)1 << 31) || !bytes_host) {
    set_error("vsm_inbox_bytes: bad arguments");
    return VSM_E_INVALID;
  }
  *bytes_host = (int64_t)inbox_layout(dim, cap_rows, cap_contrib).total;
  return VSM_OK;
}

extern "C" int vsm_peer_alloc(int32_t device, int64_t bytes, void** ptr_out, void* ipc_handle_out /* 64 bytes */) {
  if (bytes <= 0 || !ptr_out) {
    set_error("vsm_peer_alloc: bad arguments");
    return VSM_E_INVALID;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  VSM_CUDA(cudaSetDevice(device));
  void* p = nullptr;
  VSM_CUDA(cudaMalloc(&p, (size_t)bytes));  // plain cudaMalloc: memory from the stream-ordered pool cannot be shared
  VSM_CUDA(cudaMemset(p, 0, (size_t)bytes));
  if (ipc_handle_out) {
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
      cudaFree(p);
      set_error("cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
      return VSM_E_CUDA;
    }
    memcpy(ipc_handle_out, &h, 64);
  }
  *ptr_out = p;
  return VSM_OK;
}

extern "C" int vsm_peer_open(int32_t device, const void* ipc_handle, void** ptr_out) {
  if (!ipc_handle || !ptr_out) {
    set_error("vsm_peer_open: bad arguments");
    return VSM_E_INVALID;
  }
  VSM_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, 64);
  VSM_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return VSM_OK;
}

extern "C" int vsm_peer_close(int32_t device, void* ptr) {
  if (!ptr) return VSM_OK;
  VSM_CUDA(cudaSetDevice(device));
  VSM_CUDA(cudaIpcCloseMemHandle(ptr));
  return VSM_OK;
}

extern "C" int vsm_peer_free(int32_t device, void* ptr) {
  if (!ptr) return VSM_OK;
  VSM_CUDA(cudaSetDevice(device));
  VSM_CUDA(cudaDeviceSynchronize());
  VSM_CUDA(cudaFree(ptr));
  return VSM_OK;
}

extern "C" int vsm_partials_push(vsm_map* m, int32_t world, void* const* inbox_ptrs_host, int64_t cap_rows,
                                 int64_t cap_contrib, int64_t epoch, void* stream) {
  if (!m || world < 1 || world > kMaxWorld || !inbox_ptrs_host || cap_rows < 1 || cap_contrib < 1 || epoch < 0) {
    set_error("vsm_partials_push: bad arguments (world <= %d)", kMaxWorld);
    return VSM_E_INVALID;
  }
  if (m->dense_loaded) {
    set_error("vsm_partials_push: dense-loaded maps have no keys");
    return VSM_E_STATE;
  }
  for (int o = 0; o < world; ++o)
    if (!inbox_ptrs_host[o]) {
      set_error("vsm_partials_push: null inbox pointer for rank %d", o);
      return VSM_E_INVALID;
    }
  VSM_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  VSM_TRY(fuse_collect_pending(m, s));
  const InboxLayout L = inbox_layout(m->d, cap_rows, cap_contrib);
  const PeerPtrs peers = peer_ptrs(inbox_ptrs_host, world, L, (int)(epoch & 1));
  const uint32_t V = (uint32_t)m->n_vox;
  if (V) {
    push_rows_kernel<<<grid_for(V, 256, sm_count() * 4), 256, 0, s>>>(m->vkey.as<unsigned long long>(), m->vcount.as<uint32_t>(),
                                                               m->vsum.as<float>(), V, m->d, (uint32_t)world, peers,
                                                               (uint32_t)cap_rows, L.row_stride);
    VSM_LAUNCHED();
  }
  const uint32_t M = (uint32_t)m->log_n;
  if (M) {
    push_contrib_kernel<<<grid_for(M, 256, sm_count() * 4), 256, 0, s>>>(m->log_gid.as<int32_t>(), m->log_fuse.as<int32_t>(),
                                                                  m->log_mask.as<unsigned long long>(),
                                                                  m->vkey.as<unsigned long long>(), M, (uint32_t)world, peers,
                                                                  (uint32_t)cap_contrib);
    VSM_LAUNCHED();
  }
  push_signal_kernel<<<1, kMaxWorld, 0, s>>>(peers, (uint32_t)world);
  VSM_LAUNCHED();
  return VSM_OK;
}

// Queue one drain on `stream` without waiting for it: the device waits for `world` signals, inserts the received
// records into the map and writes a 4-word report into report slot `slot` (0..3).  The map must already have room
// (vsm_map_reserve / vsm_map_reserve_log): nothing about the inbox contents is known on the host, and a drain that
// runs out of room is reported by the collect, not grown for.
static int drain_enqueue(vsm_map* m, void* inbox, int32_t world, int64_t cap_rows, int64_t cap_contrib, int64_t epoch,
                         double timeout_s, int slot, bool fresh_report, cudaStream_t s) {
  m->finalized = false;
  m->ck_built = false;
  m->norms_valid = false;
  if (!m->drain_report.p) {
    VSM_TRY(m->drain_report.ensure(4 * 32, s));
    VSM_CUDA(cudaMemsetAsync(m->drain_report.p, 0, m->drain_report.bytes, s));
  }
  VSM_TRY(m->xch_tmp.ensure((size_t)cap_rows * 4, s));
  uint32_t* report = m->drain_report.as<uint32_t>() + 8 * slot;  // [n_rows, n_contrib, flags, n_vox | hash err, log n, -, -]
  if (fresh_report) VSM_CUDA(cudaMemsetAsync(report, 0, 32, s));
  const InboxLayout L = inbox_layout(m->d, cap_rows, cap_contrib);
  const int half = (int)(epoch & 1);
  uint8_t* base = reinterpret_cast<uint8_t*>(inbox);
  InboxHeader* hdr = reinterpret_cast<InboxHeader*>(base) + half;
  const uint8_t* rows = base + L.rows_off[half];
  const uint8_t* contrib = base + L.contrib_off[half];
  const unsigned long long timeout_ns = (unsigned long long)(std::max(timeout_s, 0.001) * 1e9);
  uint32_t* hash_err = report + 4;
  inbox_wait_kernel<<<1, 1, 0, s>>>(hdr, (uint32_t)world, timeout_ns);
  VSM_LAUNCHED();
  drain_keys_kernel<<<grid_for(cap_rows, 256, sm_count() * 8), 256, 0, s>>>(global_store(m), hdr, rows, L.row_stride,
                                                                            (uint32_t)cap_rows, m->xch_tmp.as<int32_t>(), hash_err);
  VSM_LAUNCHED();
  drain_rows_kernel<<<sm_count() * 8, 256, 0, s>>>(hdr, rows, L.row_stride, (uint32_t)cap_rows, m->xch_tmp.as<int32_t>(), m->d,
                                                   m->vsum.as<float>());
  VSM_LAUNCHED();
  drain_contrib_kernel<<<grid_for(cap_contrib, 256, sm_count() * 8), 256, 0, s>>>(
      global_store(m), hdr, contrib, (uint32_t)cap_contrib, m->d_n_vox.as<uint32_t>(),
      (uint32_t)std::min<int64_t>(m->log_cap, 0xFFFFFFFFll), m->log_gid.as<int32_t>(), m->log_fuse.as<int32_t>(),
      m->log_mask.as<unsigned long long>());
  VSM_LAUNCHED();
  drain_finish_kernel<<<1, 1, 0, s>>>(hdr, (uint32_t)cap_contrib, m->d_n_vox.as<uint32_t>(),
                                      (uint32_t)std::min<int64_t>(m->log_cap, 0xFFFFFFFFll), hash_err, report);
  VSM_LAUNCHED();
  return VSM_OK;
}

static int drain_collect(vsm_map* m, int slot, int32_t world, double timeout_s, int64_t cap_rows, int64_t cap_contrib,
                         int64_t* n_rows_host, int64_t* n_contrib_host, uint32_t* flags_host, cudaStream_t s) {
  uint32_t rep[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  VSM_TRY(read_back(m, rep, m->drain_report.as<uint32_t>() + 8 * slot, sizeof(rep), s));  // synchronises the stream
  m->n_vox = std::min<int64_t>(rep[3], m->vcap);
  m->log_n = rep[5];
  if (n_rows_host) *n_rows_host = rep[0];
  if (n_contrib_host) *n_contrib_host = rep[1];
  if (flags_host) *flags_host = rep[2];
  if (rep[2] & kInboxTimeout) {
    if (world > 0)
      set_error("vsm_partials_drain: timed out after %.1f s waiting for %d senders", timeout_s, world);
    else
      set_error("vsm_partials_drain: a queued drain timed out waiting for its senders");
    return VSM_E_STATE;
  }
  if (rep[2] & kInboxHashErr) {
    set_error("vsm_partials_drain: the map ran out of voxel capacity (%lld) while draining: reserve more", (long long)m->vcap);
    return VSM_E_NOMEM;
  }
  if (rep[2] & (kInboxRowOverflow | kInboxContribOverflow)) {
    set_error("vsm_partials_drain: inbox too small, or contributor log full (%u rows for %lld slots, %u contributor entries "
              "for %lld slots, log capacity %lld)", rep[0], (long long)cap_rows, rep[1], (long long)cap_contrib,
              (long long)m->log_cap);
    return VSM_E_NOMEM;
  }
  return VSM_OK;
}

static int drain_check_args(vsm_map* m, void* inbox, int32_t world, int64_t cap_rows, int64_t cap_contrib, int64_t epoch) {
  if (!m || !inbox || world < 1 || world > kMaxWorld || cap_rows < 1 || cap_contrib < 1 || epoch < 0) {
    set_error("vsm_partials_drain: bad arguments");
    return VSM_E_INVALID;
  }
  if (m->dense_loaded) {
    set_error("vsm_partials_drain: dense-loaded maps cannot be merged into");
    return VSM_E_STATE;
  }
  return VSM_OK;
}

extern "C" int vsm_partials_drain(vsm_map* m, void* inbox, int32_t world, int64_t cap_rows, int64_t cap_contrib,
                                  int64_t epoch, double timeout_s, int64_t* n_rows_host, int64_t* n_contrib_host,
                                  uint32_t* flags_host, void* stream) {
  VSM_TRY(drain_check_args(m, inbox, world, cap_rows, cap_contrib, epoch));
  VSM_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  VSM_TRY(fuse_collect_pending(m, s));
  // room for everything the inbox can hold: nothing about its contents is known on the host yet
  VSM_TRY(map_grow(m, m->n_vox + cap_rows, s));
  VSM_TRY(log_grow(m, m->log_n + cap_contrib, s));
  VSM_TRY(drain_enqueue(m, inbox, world, cap_rows, cap_contrib, epoch, timeout_s, 0, true, s));
  return drain_collect(m, 0, world, timeout_s, cap_rows, cap_contrib, n_rows_host, n_contrib_host, flags_host, s);
}

extern "C" int vsm_partials_drain_async(vsm_map* m, void* inbox, int32_t world, int64_t cap_rows, int64_t cap_contrib,
                                        int64_t epoch, double timeout_s, int32_t report_slot, void* stream) {
  VSM_TRY(drain_check_args(m, inbox, world, cap_rows, cap_contrib, epoch));
  if (report_slot < 0 || report_slot > 3) {
    set_error("vsm_partials_drain_async: report slot must be 0..3");
    return VSM_E_INVALID;
  }
  if (!m->pending.empty()) {
    set_error("vsm_partials_drain_async: the map has uncollected fuse calls");
    return VSM_E_STATE;
  }
  VSM_CUDA(cudaSetDevice(m->device));
  return drain_enqueue(m, inbox, world, cap_rows, cap_contrib, epoch, timeout_s, report_slot, false, (cudaStream_t)stream);
}

extern "C" int vsm_partials_drain_collect(vsm_map* m, int32_t report_slot, int64_t* n_rows_host, int64_t* n_contrib_host,
                                          uint32_t* flags_host, void* stream) {
  if (!m || report_slot < 0 || report_slot > 3 || !m->drain_report.p) {
    set_error("vsm_partials_drain_collect: no drain was queued with this report slot");
    return VSM_E_INVALID;
  }
  VSM_CUDA(cudaSetDevice(m->device));
  const int st = drain_collect(m, report_slot, 0, 0.0, 0, 0, n_rows_host, n_contrib_host, flags_host, (cudaStream_t)stream);
  cudaMemsetAsync(m->drain_report.as<uint32_t>() + 8 * report_slot, 0, 32, (cudaStream_t)stream);  // the slot starts over
  return st;
}
