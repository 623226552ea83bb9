// state.cuh -- host-side state of a vsm_map and its device workspaces.
//
// Data layout in HBM:
//   global hash     gkeys[GCAP] u64 (packed key | kEmptyKey), gids[GCAP] i32 -> dense voxel id
//   dense store     vkey[VCAP] u64, vcount[VCAP] u32, vsum[VCAP*d] f32 (fp32 sums, never normalised in place)
//   contributor log one entry per (fuse call, voxel): voxel id, fuse index, 2 x u64 frame mask
//   per-call scratch world points float4[N_px], point->slot i32[N_px], two submap-local hash tables,
//                   voxel-sorted (pixel, voxel id) lists for the accumulate kernel
#pragma once
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace vsm {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int dev = 0;  // device the block lives on
  // grow to at least `need` bytes; old contents are dropped unless keep_bytes > 0
  int ensure(size_t need, cudaStream_t s, size_t keep_bytes = 0, double slack = 1.0);
  void release();
  template <class T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

// submap-local open-addressing table (device view).  One slot is ONE 32-byte sector: a claim reads and updates a
// single sector instead of one per array (the tables are sized for the pixel count, ~0.5 GB, and probed at random).
struct alignas(32) Slot {
  unsigned long long key;      // kEmptyKey when free
  uint32_t count;              // points of the key in this call
  uint32_t lid;                // dense local id (index into slot_list), set by the compact step
  unsigned long long mask[2];  // frames of the submap that hit the voxel (fine table only)
};
static_assert(sizeof(Slot) == 32, "a table slot is one 32-byte sector");
struct LocalTable {
  Slot* slots;
  uint32_t* slot_list;  // occupied slots in claim order
  uint32_t* n_occ;      // number of occupied slots (device counter)
  uint32_t cap_mask;    // capacity - 1
  uint32_t limit;       // claims beyond this many flag an overflow (3/4 of the capacity)
  uint32_t* overflow;   // the call's tbl_overflow flag
};

// device-resident counters of one fuse call (zeroed at the start of every call)
struct FuseCounters {
  unsigned long long n_conf, n_finite, n_bbox, n_fused, n_bad_emb;
  uint32_t n_occ_a, n_occ_b;
  uint32_t n_new;     // this call's distinct voxels that are new to the map
  uint32_t sel_miss;  // the one-pass percentile select could not answer: repeat the call with the radix select
  uint32_t range_err, internal_err;
  uint32_t abort;      // set on the device when the call must stop before touching the global map
  uint32_t seg_total;  // running allocation of sorted-list positions
  uint32_t n_check;    // check-only entries appended behind the fused ones
  uint32_t log_base;   // first contributor-log entry of this call (allocated on the device)
  uint32_t bad_index;  // selected pixels whose embedding index lies outside the table
  uint32_t ticket;     // blocks of the capacity-check kernel that have finished (the last one decides)
  float bounds[6];     // bbox filter bounds laid out [axis][lo,hi]
  uint32_t vox_base;   // voxels in the map before this call: ids >= vox_base are new, their sum rows are still zero
  uint32_t range_dropped;  // points dropped because a finite voxel coordinate cannot be packed ("coord_range_policy" 1)
  // one-table preparation (fuse.cu, "prep v7"): every bbox survivor is inserted into the fine table; coarse cells are
  // counted per distinct voxel afterwards
  uint32_t n_kept;        // local voxels that survive the coarse-cell filter (+ pseudo voxels of irregular points)
  uint32_t n_distinct;    // distinct fine voxels among them (n_submap_voxels)
  uint32_t n_irr;         // irregular points: float32 coarse cell != the cell derived from the voxel's integer coordinates
  uint32_t irr_overflow;  // the irregular-point list overflowed: repeat the call with the two-table preparation
  uint32_t ticket2;       // last-block ticket of the coarse-count kernel
  uint32_t tbl_overflow;  // a submap-local table sized from earlier calls filled up: repeat the call with full-size tables
};

// radix-select state (device)
constexpr int kSelMaxTargets = 12;
struct SelectState {
  unsigned long long n;                    // valid elements per column
  unsigned long long rank[kSelMaxTargets]; // 0-based rank wanted
  unsigned long long rem[kSelMaxTargets];  // rank inside the current prefix bucket
  uint32_t prefix[kSelMaxTargets];
  float value[kSelMaxTargets];
  float gamma[kSelMaxTargets];             // lerp weight of the (lo,hi) pair the target belongs to
  uint32_t nan_count[4];
  int n_targets;
  int targets_per_col;
  int alias[kSelMaxTargets];               // lowest target of the same column with the same prefix (fast select)
};

// a fuse call that has been queued on the stream and not yet collected
struct PendingCall {
  vsm_fuse_params p;
  const float* pts;
  const float* conf;
  const uint8_t* emb;
  const uint8_t* emb_ok;
  int slot;             // index into the counter / event rings
  int fuse_index;       // index into vsm_map::fuses
  bool profiled;
  bool v7;              // queued with the one-table preparation (which counters hold the call's voxel count)
  DevBuf precheck_mask; // row mask computed for VSM_FUSE_EMB_PRECHECK, alive until the call is collected
};

constexpr int kCallRing = 64;  // fuse calls in flight per map

struct FuseRecord {
  int32_t submap_id;
  int32_t S, H, W, end_idx, stride;
  int64_t n_fused;
  int64_t log_begin, log_end;  // entries of the contributor log written by this call
  DevBuf point_gid;            // int32[S*H*W] voxel id per pixel (-1: not fused) if KEEP_POINT_INDEX
};

}  // namespace vsm

namespace vsm {
// Per-device scratch shared by all maps of the process (fuse calls borrow it under the mutex): world points,
// point -> slot, the two submap-local hash tables (left clean by every call), per-local-voxel arrays and the
// voxel-sorted point lists.  Kept across maps so that building a new map does not re-allocate ~1 GB.
struct Workspace {
  std::mutex mu;
  DevBuf pw;        // float4[N_px]
  DevBuf pt_slot;   // int32[N_px]
  DevBuf pt2;       // uint2[N_px]: (fine-table slot | -1, ordinal of the point inside its voxel) -- one-table preparation
  DevBuf kept;      // uint32[N_sel]: slots of the local voxels that survive the coarse-cell filter
  DevBuf irr;       // irregular-point list (IrrEntry)
  DevBuf ta_slots, ta_list;  // coarse table
  DevBuf tb_slots, tb_list;  // fine table
  uint64_t ta_cap = 0, tb_cap = 0;
  // Stream order of the borrowers: every use of the workspace ends with an event on the borrower's stream; a borrower
  // on ANOTHER stream waits for it first (two maps of one device fused on different streams share these buffers).
  cudaEvent_t ev_last_use = nullptr;
  cudaStream_t last_stream = nullptr;
  bool used = false;
  DevBuf lv_cnt, lv_off, lv_cursor, lv_gid;  // per local voxel
  DevBuf sorted_pix[2], sorted_gid;
  DevBuf sel_bracket;  // scratch of the one-pass percentile select
  int64_t hint_n_occ = 0;  // most distinct voxels any collected fuse call on this device had (growth heuristic, table prefix)
  double hint_vs = 0.0;    // voxel size of the calls the hint comes from (0: unknown)
  // The accumulate kernel of a (voxel-sorted, device-resident) fuse call runs on this side stream, so that the
  // preparation kernels of the NEXT call -- queued on the caller's stream -- overlap it.  The sorted entry lists
  // are double-buffered; ev_acc_done[b] marks the end of the last accumulate that read sorted_pix[b].
  cudaStream_t acc_stream = nullptr;
  cudaEvent_t ev_prep_done[2] = {nullptr, nullptr};
  cudaEvent_t ev_acc_done[2] = {nullptr, nullptr};
  bool acc_used[2] = {false, false};
  int acc_parity = 0;
  // Off by default: measured on B200, the HBM-saturating accumulate kernel and the latency-bound preparation
  // kernels slow each other down by as much as the overlap hides.  VSM_OVERLAP=1 / vsm_set_option.
  bool overlap = false;
  // SM partitions (CUDA green contexts, "green_prep_sms" option): with the overlap on, the preparation kernels of a call
  // run on a stream confined to `green_sms[0]` SMs and the accumulate kernel on a stream confined to the other
  // `green_sms[1]`, so that the HBM-bound accumulate of call i and the latency / issue-bound preparation of call
  // i+1 run side by side without competing for registers and warp slots on the same SM.
  bool green_on = false;
  void* green_ctx[2] = {nullptr, nullptr};  // CUgreenCtx of the active partition
  cudaStream_t green_stream[2] = {nullptr, nullptr};  // [0] preparation, [1] accumulate
  int green_sms[2] = {0, 0};
  // every partition ever asked for stays alive (a handful of contexts): switching between splits, or off and on
  // again, creates and destroys nothing
  struct GreenPart {
    int requested;
    void* ctx[2];
    cudaStream_t stream[2];
    int sms[2];
  };
  std::vector<GreenPart> green_parts;
  cudaEvent_t ev_fork = nullptr, ev_prep_join = nullptr;
};
Workspace* workspace_for_device(int device);

// Borrows the device workspace for work queued on stream `s`: takes the mutex, makes `s` wait for the previous
// borrower's kernels if they ran on a different stream, and on release records the end of this use on `s`.
class WsLease {
 public:
  WsLease(Workspace* ws, cudaStream_t s);
  ~WsLease();
  int status() const { return status_; }
  WsLease(const WsLease&) = delete;
  WsLease& operator=(const WsLease&) = delete;

 private:
  Workspace* ws_;
  cudaStream_t s_;
  int status_;
};
}  // namespace vsm

struct vsm_map {
  vsm_config cfg{};
  int device = 0;
  int d = 0;
  int esize = 4;      // bytes per embedding element
  float vs_f = 0.f;   // (float)voxel_size
  vsm::Workspace* ws = nullptr;

  // global hash + dense store
  uint64_t gcap = 0;
  vsm::DevBuf gkeys, gids;
  int64_t vcap = 0;
  int64_t n_vox = 0;  // host mirror of *d_n_vox
  int64_t last_n_occ = 0;  // distinct voxels of the previous fuse call (growth heuristic)
  vsm::DevBuf vkey, vcount, vsum;
  vsm::DevBuf d_n_vox;  // device-side map state: uint32 [0] voxel count, [1] contributor-log entries
  vsm::DevBuf ctr_ring;  // kCallRing x FuseCounters, one slot per queued fuse call
  std::vector<vsm::PendingCall> pending;
  std::vector<vsm_fuse_stats> stats_backlog;  // stats of queued calls collected internally, handed out by vsm_fuse_collect
  cudaEvent_t ev_ring[vsm::kCallRing][3] = {};
  vsm_fuse_stats last_stats{};

  // contributor log
  int64_t log_n = 0, log_cap = 0;
  vsm::DevBuf log_gid, log_fuse, log_mask;
  std::vector<vsm::FuseRecord> fuses;

  // per-call scratch
  vsm::DevBuf ctr;       // FuseCounters
  vsm::DevBuf sel;       // SelectState
  vsm::DevBuf sel_hist;  // histograms
  vsm::DevBuf cub_tmp;
  vsm::DevBuf stage_pts, stage_conf, stage_emb[2];  // host-entry staging
  void* pinned = nullptr;                           // pinned host scratch (counters read-back)
  size_t pinned_bytes = 0;
  void* pinned_stage[2] = {nullptr, nullptr};
  size_t pinned_stage_bytes = 0;
  cudaEvent_t ev_stage[2] = {nullptr, nullptr};
  cudaEvent_t ev_copy[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;

  // profiling (vsm_profile_enable)
  bool profiling = false;
  cudaEvent_t ev_prof[3] = {nullptr, nullptr, nullptr};
  vsm_profile prof{};

  // finalisation products
  bool finalized = false;
  bool dense_loaded = false;  // rows came from vsm_map_load_dense: rank == id, keys unknown
  vsm::DevBuf sorted_keys, id_of_rank, rank_of_id;
  vsm::DevBuf csr_off, csr_sub, csr_mask;
  int64_t csr_entries = 0;
  vsm::DevBuf dense_centers;  // float[V*3] for loaded maps
  // compat lookup table keyed by reconstructed coords (3 x int64 hashed)
  vsm::DevBuf ck_keys, ck_val;
  uint64_t ck_cap = 0;
  bool ck_built = false;
  vsm::DevBuf drain_report;  // 4 report slots of queued drains (vsm_partials_drain_async)
  vsm::DevBuf xch_tmp;  // exchange scratch (owner offsets of a pack, voxel ids of a merge / drain): per map, not shared
  // query scratch
  vsm::DevBuf q_cand, q_tmp, q_norm;
  vsm::DevBuf q_tc, q_tc_cand;     // tensor-core engine: thresholds / padded prompts; candidate ids + keys
  bool norms_valid = false;        // q_norm holds ||sum_v|| for the current contents
  vsm::DevBuf q_shadow;            // bf16 copy of the sums for the tensor-core passes of engine 3 (built on demand)
  bool shadow_valid = false;
  uint32_t tc_last_candidates = 0; // longest candidate list of the last engine-2 query
  int64_t tc_fallbacks = 0;        // engine-2 queries answered by engine 1 (candidate overflow)
};

namespace vsm {
// strided float source with an optional flag word per element
struct SelSrc {
  const float* base;  // element i, column c at base[i*stride + c]
  int stride;
  int ncol;
  int flag_off;        // offset of the flag word inside the element, -1: no flags
  uint32_t flag_need;  // (flags & flag_need) == flag_need -> element takes part
  int64_t n_items;
};

// out_dev receives ncol*npct floats laid out [col][pct]; valid elements are counted in pass 0.
int run_percentiles(SelectState* st, uint32_t* hist, const SelSrc& src, int npct, float q0, float q1, float* out_dev,
                    cudaStream_t s);
// the same, for a caller that has already filled pass 0 (hist[0..3*2048), after select_reset) and knows the
// element count on the device (n_dev)
int select_reset(SelectState* st, uint32_t* hist, cudaStream_t s);
int run_percentiles_after_hist0(SelectState* st, uint32_t* hist, const SelSrc& src, int npct, float q0, float q1,
                                float* out_dev, const unsigned long long* n_dev, cudaStream_t s);
// the same for the world-point layout, with the one-block steps merged (plan + pick, pick, pick + lerp) and targets
// that share a prefix histogrammed once: 5 launches instead of 8
int run_percentiles_world_fast(SelectState* st, uint32_t* hist, const float4* pw, int64_t n_items, uint32_t flag_need,
                               float q0, float q1, float* out_dev, const unsigned long long* n_dev, cudaStream_t s);
// process-wide scratch per device: select state, histograms, 16 result floats
int select_scratch(SelectState** st, uint32_t** hist, float** out);
int map_grow(vsm_map* m, int64_t need_voxels, cudaStream_t s);
int log_grow(vsm_map* m, int64_t need_entries, cudaStream_t s);
int fuse_collect_pending(vsm_map* m, cudaStream_t s);  // collect queued fuse calls (no-op if none)
int join_accumulates(Workspace* ws, cudaStream_t s);    // make s wait for the accumulate kernels queued on the side stream
int read_back(vsm_map* m, void* dst_host, const void* src_dev, size_t bytes, cudaStream_t s);
}  // namespace vsm
