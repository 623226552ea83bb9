// api.cu -- library plumbing: errors, device buffers, map life cycle.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <mutex>
#include <vector>

#include "state.cuh"

namespace vsm {

static thread_local char g_err[1024] = "";
int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// All device memory of the library comes from the device's stream-ordered pool (cudaMallocAsync) with the
// release threshold lifted, so that re-creating maps and growing scratch buffers costs microseconds after
// warm-up instead of a cudaMalloc/cudaFree pair (milliseconds, with a device-wide synchronisation).
// Allocation happens on a private stream that is synchronised right away: the block is then valid on any stream.
static cudaStream_t alloc_stream_for(int dev) {
  static cudaStream_t streams[64] = {nullptr};
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 0 || dev >= 64) return nullptr;
  if (!streams[dev]) {
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(dev);
    cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaSetDevice(cur);
  }
  return streams[dev];
}

static double host_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static bool host_trace() {
  static const bool on = getenv("VSM_TRACE") && getenv("VSM_TRACE")[0] == '1';
  return on;
}

int DevBuf::ensure(size_t need, cudaStream_t s, size_t keep_bytes, double slack) {
  if (need <= bytes && p != nullptr) return VSM_OK;
  const double t_begin = host_trace() ? host_ms() : 0.0;
  struct Report {
    double t0;
    size_t need, keep;
    ~Report() {
      if (t0 > 0.0 && host_ms() - t0 > 1.0)
        fprintf(stderr, "[vsm trace] allocation of %.1f MB (keeping %.1f MB) took %.2f ms\n", need * 1e-6, keep * 1e-6, host_ms() - t0);
    }
  } report{t_begin, need, keep_bytes};
  if (need == 0) need = 16;
  size_t want = (size_t)((double)need * slack);
  want = (want + 255) & ~(size_t)255;
  int cur = 0;
  VSM_CUDA(cudaGetDevice(&cur));
  cudaStream_t as = alloc_stream_for(cur);
  void* np = nullptr;
  cudaError_t e = cudaMallocAsync(&np, want, as);
  if (e != cudaSuccess && want > need) {
    cudaGetLastError();
    want = (need + 255) & ~(size_t)255;
    e = cudaMallocAsync(&np, want, as);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(as);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("device allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
    return VSM_E_NOMEM;
  }
  if (p != nullptr) {
    if (keep_bytes) VSM_CUDA(cudaMemcpyAsync(np, p, keep_bytes, cudaMemcpyDeviceToDevice, s));
    VSM_CUDA(cudaStreamSynchronize(s));  // nothing on the caller's stream still uses the old block
    VSM_CUDA(cudaFreeAsync(p, alloc_stream_for(dev)));
  }
  p = np;
  bytes = want;
  dev = cur;
  return VSM_OK;
}

void DevBuf::release() {
  if (p) {
    cudaFreeAsync(p, alloc_stream_for(dev));
    cudaGetLastError();
  }
  p = nullptr;
  bytes = 0;
}

Workspace* workspace_for_device(int device) {
  static Workspace* table[64] = {nullptr};
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (device < 0 || device >= 64) return nullptr;
  if (!table[device]) table[device] = new Workspace();
  return table[device];
}

WsLease::WsLease(Workspace* ws, cudaStream_t s) : ws_(ws), s_(s), status_(VSM_OK) {
  ws_->mu.lock();
  if (!ws_->ev_last_use && cudaEventCreateWithFlags(&ws_->ev_last_use, cudaEventDisableTiming) != cudaSuccess) {
    set_error("workspace: cannot create the stream-order event: %s", cudaGetErrorString(cudaGetLastError()));
    status_ = VSM_E_CUDA;
    return;
  }
  if (ws_->used && ws_->last_stream != s_) {
    // the previous borrower's kernels (another stream) still own the buffers: order this stream behind them
    if (cudaStreamWaitEvent(s_, ws_->ev_last_use, 0) != cudaSuccess || join_accumulates(ws_, s_) != VSM_OK) {
      set_error("workspace: cannot order stream behind the previous borrower: %s", cudaGetErrorString(cudaGetLastError()));
      status_ = VSM_E_CUDA;
    }
  }
}

WsLease::~WsLease() {
  if (status_ == VSM_OK && ws_->ev_last_use) {
    if (cudaEventRecord(ws_->ev_last_use, s_) == cudaSuccess) {
      ws_->last_stream = s_;
      ws_->used = true;
    } else {
      cudaGetLastError();
    }
  }
  ws_->mu.unlock();
}

int read_back(vsm_map* m, void* dst_host, const void* src_dev, size_t bytes, cudaStream_t s) {
  if (m->pinned_bytes < bytes) {
    if (m->pinned) cudaFreeHost(m->pinned);
    m->pinned = nullptr;
    m->pinned_bytes = 0;
    VSM_CUDA(cudaMallocHost(&m->pinned, std::max<size_t>(bytes, 4096)));
    m->pinned_bytes = std::max<size_t>(bytes, 4096);
  }
  VSM_CUDA(cudaMemcpyAsync(m->pinned, src_dev, bytes, cudaMemcpyDeviceToHost, s));
  VSM_CUDA(cudaStreamSynchronize(s));
  memcpy(dst_host, m->pinned, bytes);
  return VSM_OK;
}

}  // namespace vsm

using namespace vsm;

extern "C" int vsm_abi_version(void) { return VSM_ABI_VERSION; }
extern "C" const char* vsm_last_error(void) { return g_err; }
extern "C" int64_t vsm_launch_count(void) { return g_launches; }

// Retired maps are kept (cleared) per device and handed out again by vsm_map_create: building map after map -- the
// benchmark's step, an evaluation loop -- then allocates and frees nothing.  Only maps below a size bound are kept.
static std::mutex g_cache_mu;
static std::vector<vsm_map*> g_cache;
constexpr size_t kCacheMaxBytes = (size_t)8 << 30;  // per cached map
constexpr size_t kCacheMaxMaps = 4;

static size_t map_bytes(const vsm_map* m) {
  return m->gkeys.bytes + m->gids.bytes + m->vkey.bytes + m->vcount.bytes + m->vsum.bytes + m->log_gid.bytes +
         m->log_fuse.bytes + m->log_mask.bytes;
}
static int map_destroy_now(vsm_map* m);

extern "C" int vsm_map_create(const vsm_config* cfg, vsm_map** out) {
  if (!cfg || !out) {
    set_error("null config or output");
    return VSM_E_INVALID;
  }
  *out = nullptr;
  if (!(cfg->voxel_size > 0.0)) {
    set_error("voxel_size must be > 0");  // submap.py:236, map.py:185
    return VSM_E_INVALID;
  }
  if (cfg->emb_dtype != VSM_F32 && cfg->emb_dtype != VSM_BF16) {
    set_error("emb_dtype must be VSM_F32 or VSM_BF16");
    return VSM_E_INVALID;
  }
  const int esize = cfg->emb_dtype == VSM_BF16 ? 2 : 4;
  if (cfg->dim <= 0 || cfg->dim % 8 != 0 || (int64_t)cfg->dim * esize > 32 * 8 * 16) {
    set_error("dim must be a positive multiple of 8 with at most %d bytes per row", 32 * 8 * 16);
    return VSM_E_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("no CUDA device: libvsm has no CPU fallback");
    return VSM_E_CUDA;
  }
  int dev = cfg->device;
  if (dev < 0) VSM_CUDA(cudaGetDevice(&dev));
  if (dev >= ndev) {
    set_error("device %d out of range (%d devices)", dev, ndev);
    return VSM_E_INVALID;
  }
  VSM_CUDA(cudaSetDevice(dev));
  {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    for (size_t i = 0; i < g_cache.size(); ++i) {
      vsm_map* c = g_cache[i];
      if (c->device == dev && c->d == cfg->dim && c->cfg.emb_dtype == cfg->emb_dtype) {
        g_cache.erase(g_cache.begin() + i);
        c->cfg = *cfg;
        c->vs_f = (float)cfg->voxel_size;
        const int st = map_grow(c, std::max<int64_t>(cfg->voxel_capacity, 1024), nullptr);
        if (st != VSM_OK) {
          map_destroy_now(c);
          return st;
        }
        *out = c;
        return VSM_OK;
      }
    }
  }
  vsm_map* m = new vsm_map();
  m->cfg = *cfg;
  m->device = dev;
  m->d = cfg->dim;
  m->esize = esize;
  m->vs_f = (float)cfg->voxel_size;
  m->ws = workspace_for_device(dev);
  if (!m->ws) {
    delete m;
    set_error("device ordinal %d not supported", dev);
    return VSM_E_INVALID;
  }
  int st = m->d_n_vox.ensure(2 * sizeof(uint32_t), nullptr);
  if (st == VSM_OK) st = cudaMemset(m->d_n_vox.p, 0, 2 * sizeof(uint32_t)) == cudaSuccess ? VSM_OK : VSM_E_CUDA;
  if (st == VSM_OK) st = map_grow(m, std::max<int64_t>(cfg->voxel_capacity, 1024), nullptr);
  if (st != VSM_OK) {
    vsm_map_destroy(m);
    return st;
  }
  VSM_CUDA(cudaDeviceSynchronize());
  *out = m;
  return VSM_OK;
}

extern "C" int vsm_map_destroy(vsm_map* m) {
  if (!m) return VSM_OK;
  cudaSetDevice(m->device);
  if (map_bytes(m) <= kCacheMaxBytes) {
    // keep it: cleared now (on the default stream), reused by the next vsm_map_create of the same shape
    cudaDeviceSynchronize();
    if (vsm_map_clear(m, nullptr) == VSM_OK) {
      m->profiling = false;
      m->prof = vsm_profile{};
      m->tc_fallbacks = 0;
      std::lock_guard<std::mutex> lock(g_cache_mu);
      if (g_cache.size() < kCacheMaxMaps) {
        g_cache.push_back(m);
        return VSM_OK;
      }
    }
  }
  return map_destroy_now(m);
}

extern "C" int vsm_map_cache_release(void) {
  std::vector<vsm_map*> parked;
  {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    parked.swap(g_cache);
  }
  for (vsm_map* c : parked) map_destroy_now(c);
  return VSM_OK;
}

extern "C" int vsm_pool_trim(void) {
  int dev = 0;
  cudaMemPool_t pool;
  VSM_CUDA(cudaGetDevice(&dev));
  VSM_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
  VSM_CUDA(cudaDeviceSynchronize());
  VSM_CUDA(cudaMemPoolTrimTo(pool, 0));
  return VSM_OK;
}

static int map_destroy_now(vsm_map* m) {
  cudaSetDevice(m->device);
  cudaDeviceSynchronize();
  vsm::DevBuf* bufs[] = {&m->gkeys,     &m->gids,      &m->vkey,       &m->vcount,    &m->vsum,       &m->d_n_vox,
                         &m->log_gid,   &m->log_fuse,  &m->log_mask,   &m->ctr, &m->ctr_ring,      &m->sel,        &m->sel_hist,
                         &m->cub_tmp,
                         &m->stage_pts, &m->stage_conf, &m->stage_emb[0], &m->stage_emb[1], &m->sorted_keys,
                         &m->id_of_rank, &m->rank_of_id, &m->csr_off,  &m->csr_sub,   &m->csr_mask,   &m->dense_centers,
                         &m->ck_keys,   &m->ck_val,    &m->q_cand,     &m->q_tmp,     &m->q_norm,
                         &m->q_tc,      &m->q_tc_cand,  &m->q_shadow,  &m->xch_tmp,   &m->drain_report};
  for (auto* b : bufs) b->release();
  for (auto& f : m->fuses) f.point_gid.release();
  if (m->pinned) cudaFreeHost(m->pinned);
  for (int b = 0; b < 2; ++b) {
    if (m->pinned_stage[b]) cudaFreeHost(m->pinned_stage[b]);
    if (m->ev_stage[b]) cudaEventDestroy(m->ev_stage[b]);
    if (m->ev_copy[b]) cudaEventDestroy(m->ev_copy[b]);
  }
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  for (int i = 0; i < 3; ++i)
    if (m->ev_prof[i]) cudaEventDestroy(m->ev_prof[i]);
  for (auto& c : m->pending) c.precheck_mask.release();
  for (int r = 0; r < vsm::kCallRing; ++r)
    for (int i = 0; i < 3; ++i)
      if (m->ev_ring[r][i]) cudaEventDestroy(m->ev_ring[r][i]);
  cudaGetLastError();
  delete m;
  return VSM_OK;
}

extern "C" int vsm_map_clear(vsm_map* m, void* stream) {
  if (!m) {
    set_error("null map");
    return VSM_E_INVALID;
  }
  VSM_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (m->ws) VSM_TRY(vsm::join_accumulates(m->ws, s));
  VSM_CUDA(cudaStreamSynchronize(s));  // queued fuse calls are dropped with the contents
  VSM_CUDA(cudaMemsetAsync(m->gkeys.p, 0xFF, m->gcap * 8, s));
  VSM_CUDA(cudaMemsetAsync(m->gids.p, 0xFF, m->gcap * 4, s));
  // rows at or beyond the voxel count have never been written (sums start at zero and only ids < n_vox are added to)
  {
    uint32_t state[2] = {0, 0};
    VSM_CUDA(cudaMemcpy(state, m->d_n_vox.p, sizeof(state), cudaMemcpyDeviceToHost));
    const size_t used = std::min<size_t>((size_t)m->vcap, std::max<size_t>((size_t)m->n_vox, (size_t)state[0]));
    if (used) {
      VSM_CUDA(cudaMemsetAsync(m->vcount.p, 0, used * 4, s));
      VSM_CUDA(cudaMemsetAsync(m->vsum.p, 0, used * (size_t)m->d * 4, s));
    }
  }
  VSM_CUDA(cudaMemsetAsync(m->d_n_vox.p, 0, 2 * sizeof(uint32_t), s));
  if (m->drain_report.p) VSM_CUDA(cudaMemsetAsync(m->drain_report.p, 0, m->drain_report.bytes, s));
  for (auto& c : m->pending) c.precheck_mask.release();
  m->pending.clear();
  m->stats_backlog.clear();
  VSM_CUDA(cudaStreamSynchronize(s));
  m->n_vox = 0;
  m->log_n = 0;
  m->last_n_occ = 0;
  for (auto& f : m->fuses) f.point_gid.release();
  m->fuses.clear();
  m->finalized = false;
  m->norms_valid = false;
  m->dense_loaded = false;
  m->ck_built = false;
  m->csr_entries = 0;
  return VSM_OK;
}

// vsm_map_clear without the host waiting: the resets are queued on `stream` behind whatever still reads the map there
// (an exchange push, an export).  Only valid when no fuse call is pending; the caller orders other streams itself.
extern "C" int vsm_map_clear_async(vsm_map* m, void* stream) {
  if (!m) {
    set_error("null map");
    return VSM_E_INVALID;
  }
  if (!m->pending.empty()) {
    set_error("vsm_map_clear_async: the map has uncollected fuse calls");
    return VSM_E_STATE;
  }
  VSM_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  VSM_CUDA(cudaMemsetAsync(m->gkeys.p, 0xFF, m->gcap * 8, s));
  VSM_CUDA(cudaMemsetAsync(m->gids.p, 0xFF, m->gcap * 4, s));
  const size_t used = std::min<size_t>((size_t)m->vcap, (size_t)m->n_vox);
  if (used) {
    VSM_CUDA(cudaMemsetAsync(m->vcount.p, 0, used * 4, s));
    VSM_CUDA(cudaMemsetAsync(m->vsum.p, 0, used * (size_t)m->d * 4, s));
  }
  VSM_CUDA(cudaMemsetAsync(m->d_n_vox.p, 0, 2 * sizeof(uint32_t), s));
  if (m->drain_report.p) VSM_CUDA(cudaMemsetAsync(m->drain_report.p, 0, m->drain_report.bytes, s));
  m->stats_backlog.clear();
  m->n_vox = 0;
  m->log_n = 0;
  m->last_n_occ = 0;
  for (auto& f : m->fuses) f.point_gid.release();
  m->fuses.clear();
  m->finalized = false;
  m->norms_valid = false;
  m->dense_loaded = false;
  m->ck_built = false;
  m->csr_entries = 0;
  return VSM_OK;
}

extern "C" int vsm_map_reserve_log(vsm_map* m, int64_t entries, void* stream) {
  if (!m || entries < 0) {
    set_error("vsm_map_reserve_log: bad arguments");
    return VSM_E_INVALID;
  }
  VSM_CUDA(cudaSetDevice(m->device));
  return log_grow(m, entries, (cudaStream_t)stream);
}

extern "C" int vsm_map_reserve(vsm_map* m, int64_t voxel_capacity, void* stream) {
  if (!m) {
    set_error("null map");
    return VSM_E_INVALID;
  }
  VSM_CUDA(cudaSetDevice(m->device));
  return map_grow(m, voxel_capacity, (cudaStream_t)stream);
}

extern "C" int vsm_num_log_entries(const vsm_map* m, int64_t* out_host) {
  if (!m || !out_host) {
    set_error("null argument");
    return VSM_E_INVALID;
  }
  *out_host = m->log_n;
  return VSM_OK;
}

extern "C" int vsm_num_voxels(const vsm_map* m, int64_t* out_host) {
  if (!m || !out_host) {
    set_error("null argument");
    return VSM_E_INVALID;
  }
  *out_host = m->n_vox;
  return VSM_OK;
}
