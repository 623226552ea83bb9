/*
 * vsm.h -- C ABI of libvsm.so: semantic voxel mapping + text query on B200 (sm_100a).
 *
 * The reference (juexZZ/VGGT-SLAM) has no FFI for this path: its callers use
 * Python objects (vggt_slam/submap.py, map.py, semantic_voxel.py).  This header
 * is the boundary a maintainer binds under those classes (ctypes stub in
 * INTEGRATION.md).  Each entry point cites the reference code it replaces
 * (paths relative to the upstream repository root).
 *
 * Conventions
 *   - every function returns an int status, 0 = VSM_OK; vsm_last_error() returns
 *     a thread-local message for the last failing call on this thread;
 *   - pointers named *_dev are device pointers on the map's CUDA device, *_host
 *     are host pointers; plain C types only, no torch / C++ types;
 *   - `stream` is a cudaStream_t (CUstream) passed as void*; NULL = default stream;
 *   - calls are asynchronous on `stream` unless the documentation says they
 *     return a host value, in which case they synchronise that stream;
 *   - handles are opaque, owned by the caller until *_destroy; one map may be
 *     used from one thread at a time; one CUDA context per process (spawned
 *     worker processes are fine, fork after first use is not);
 *   - no exception crosses this boundary and there is NO CPU fallback: without a
 *     CUDA device every compute entry point returns VSM_E_CUDA.
 */
#ifndef VSM_H_
#define VSM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSM_ABI_VERSION 3

#if defined(__GNUC__)
#define VSM_API __attribute__((visibility("default")))
#else
#define VSM_API
#endif

/* status codes */
#define VSM_OK 0
#define VSM_E_INVALID 1        /* bad argument (reference: ValueError / TypeError) */
#define VSM_E_CUDA 2           /* CUDA runtime error, or no device */
#define VSM_E_NOMEM 3          /* device allocation failed */
#define VSM_E_COORD_RANGE 4    /* a finite voxel coordinate outside +-(2^20-1) cells (option "coord_range_policy" 0) */
#define VSM_E_NONFINITE_EMB 5  /* optimistic filter pass met a non-finite embedding row: map poisoned, redo with VSM_FUSE_EMB_PRECHECK */
#define VSM_E_STATE 6          /* call not valid in the map's current state */
#define VSM_E_TOO_MANY_FRAMES 7/* more than VSM_MAX_FRAMES frames in one submap */
#define VSM_E_INTERNAL 8       /* hash table probe limit hit or similar: a bug */

#define VSM_MAX_FRAMES 128     /* frames per submap (contributor masks are 2 x u64) */
#define VSM_MAX_PROMPTS 256
#define VSM_MAX_TOPK 1024

/* embedding element types */
#define VSM_F32 0
#define VSM_BF16 1

typedef struct vsm_map vsm_map;

typedef struct vsm_config {
  double voxel_size;       /* python float of the reference; kernels divide by (float)voxel_size */
  int32_t dim;             /* embedding channels d (multiple of 8) */
  int32_t emb_dtype;       /* VSM_F32 | VSM_BF16: element type of the embeddings passed to fuse calls */
  int64_t voxel_capacity;  /* initial voxel capacity; grows on demand */
  int32_t device;          /* CUDA ordinal, -1 = current device */
  int32_t reserved;
} vsm_config;

/* flags of vsm_fuse_params.flags */
#define VSM_FUSE_FILTERS 1u          /* the three per-submap outlier filters of GraphMap.build_semantic_voxel_map (map.py:247-280) */
#define VSM_FUSE_KEEP_POINT_INDEX 2u /* remember point -> voxel (np.unique's `inverse`) for export */
#define VSM_FUSE_EMB_PRECHECK 4u     /* read embeddings twice: exact finite-row filter before the percentiles (map.py:247) */
#define VSM_FUSE_PIXEL_ORDER 8u      /* accumulate in pixel order (streaming kernel) instead of voxel-sorted order */

typedef struct vsm_fuse_params {
  int32_t S, H, W;          /* pointmap array dims: points (S,H,W,3), conf (S,H,W), emb (S,H,W,d) */
  int32_t end_idx;          /* frames [0,end_idx) are fused (min(S, last_non_loop_frame_index+1), map.py:205-207) */
  int32_t stride;           /* pixel stride (map.py:213-216) */
  float conf_threshold;     /* Submap.conf_threshold, float32 (submap.py:38) */
  double H_world_map[16];   /* row-major 4x4 float64 (map.py:73-76) */
  int32_t submap_id;
  uint32_t flags;
  double bbox_lo_pct;       /* 0.5  (map.py:257) */
  double bbox_hi_pct;       /* 99.5 (map.py:258) */
  double coarse_factor;     /* 3.0  (map.py:271) */
  int32_t coarse_min_points;/* 10   (map.py:272) */
  int32_t frame_base;       /* index, inside its submap, of this call's frame 0 (per-frame streaming: S = 1 calls) */
  /* Indexed embeddings.  The dense (S,H,W,d) array the reference fuses is piecewise constant: the embedder
   * paints one CLIP vector per SAM mask into the pixels (semantic_embedder.py:324-349, zeros where no mask).  With
   * emb_index_dev != NULL the `emb` argument of a fuse call is that TABLE, (emb_rows, d), and pixel p carries
   * table[emb_index_dev[p]]: the same map as fusing the expanded array, from 4 bytes per pixel instead of 2d or 4d.
   * Indices outside [0, emb_rows) fail the call with VSM_E_INVALID before the map is touched.  Device fuse calls only. */
  const int32_t* emb_index_dev; /* int32 (S,H,W) on the device, or NULL for dense embeddings */
  int32_t emb_rows;
  int32_t reserved;
} vsm_fuse_params;

typedef struct vsm_fuse_stats {
  int64_t n_conf;          /* points passing conf >= threshold (and stride / end_idx) */
  int64_t n_finite;        /* ... and finite (filter 1) */
  int64_t n_bbox;          /* ... and inside the percentile box (filter 2) */
  int64_t n_fused;         /* ... and in a coarse cell with enough points (filter 3) = points accumulated */
  int64_t n_submap_voxels; /* distinct voxels this call touched */
  int64_t n_map_voxels;    /* voxels in the map after the call */
  int64_t n_bad_emb_rows;  /* non-finite embedding rows met while accumulating */
  float bbox_lo[3], bbox_hi[3];
  int64_t n_range_dropped; /* points dropped because a finite voxel coordinate lies outside +-(2^20-1) cells
                              (only with vsm_set_option("coord_range_policy", 1); 0 otherwise) */
} vsm_fuse_stats;

/* ---- library ---------------------------------------------------------- */
VSM_API int vsm_abi_version(void);
VSM_API const char* vsm_last_error(void);
/* number of kernels launched by this library in this process so far (bench.py's gpu_launches) */
VSM_API int64_t vsm_launch_count(void);

/* All maps of one device share a scratch workspace (world points, submap-local hash tables).  Calls may come from
 * any stream: a call queued on a stream other than the previous borrower's first waits (on the device) for that
 * borrower's kernels, so two maps fused on two streams serialise instead of corrupting each other.
 *
 * process-wide options: "prep_variant" bit mask of the preparation kernels (1 = a warp takes a 4x8 pixel patch, 2 = probe
 * the frame mask before the atomic OR, 4 = select with merged one-block steps; default 5; results identical for every
 * value); "coord_range_policy" 0 (default) = a finite point whose voxel coordinate cannot be packed fails the call with
 * VSM_E_COORD_RANGE, 1 = such points are dropped and counted in vsm_fuse_stats.n_range_dropped (the reference accepts
 * any int64 coordinate; +-(2^20-1) cells is +-21 km at 2 cm); "select_mode" 0 default (= 1) / 1 three-pass radix select / 2 bracket select (sampled brackets,
 * collect fused into the world-point kernel) for the bbox percentiles (identical results; tests force each); "overlap" 0/1 runs the accumulate kernel on a side stream beside the
 * next call's preparation kernels (current device).  Counters: "select_misses" = fuse calls repeated with the radix
 * select because the bracket select could not answer; "capacity_retries" = calls repeated after the map or the
 * contributor log had to grow; "early_collects" = times a submit had to collect the queued calls itself. */
VSM_API int vsm_set_option(const char* key, int64_t value);
VSM_API int vsm_get_counter(const char* key, int64_t* out_host);

/* ---- map life cycle ---------------------------------------------------- */
/* replaces: the python containers built at the end of map.py:375-381 / submap.py:306-311 */
VSM_API int vsm_map_create(const vsm_config* cfg, vsm_map** out);
/* A destroyed map of up to 8 GB is cleared and parked (at most 4 per process); the next vsm_map_create with the same
 * device, dim and embedding type takes it over, so that building map after map allocates nothing.
 * vsm_map_cache_release frees the parked maps of every device. */
VSM_API int vsm_map_destroy(vsm_map* m);
VSM_API int vsm_map_cache_release(void);
/* libvsm keeps every block it has freed in the current device's stream-ordered pool (re-creating maps and scratch costs
 * microseconds); this hands the unused ones back to the device, e.g. before another allocator needs the room */
VSM_API int vsm_pool_trim(void);
VSM_API int vsm_map_clear(vsm_map* m, void* stream);
VSM_API int vsm_map_reserve(vsm_map* m, int64_t voxel_capacity, void* stream);
/* room for `entries` contributor-log entries (one per fuse call and voxel, or per received contributor record) */
VSM_API int vsm_map_reserve_log(vsm_map* m, int64_t entries, void* stream);
/* vsm_map_clear without the host waiting: the resets are queued on `stream` behind whatever still reads the map there
 * (an exchange push).  Only valid when no fuse call is pending; ordering against other streams is the caller's. */
VSM_API int vsm_map_clear_async(vsm_map* m, void* stream);

/* ---- a1: Submap.add_all_points' np.percentile(conf, pct) (submap.py:38) --
 * numpy-2 'linear' method on float32 data (index arithmetic in float32).  Synchronises. */
VSM_API int vsm_conf_threshold(const float* conf_dev, int64_t n, double percentile, float* out_host, void* stream);

/* ---- a5: (H @ [p;1]) / w (submap.py:171-174, 185-188; map.py:232-234) ----
 * out_f64 != 0: out_dev is double[n*3] (the reference's float64 result);
 * out_f64 == 0: out_dev is float[n*3] (rounded like .astype(float32)). */
VSM_API int vsm_transform_points(const float* pts_dev, int64_t n, const double* H_host16, void* out_dev, int out_f64,
                         void* stream);

/* ---- a4+a5: Submap.get_points_in_world_frame / get_points_colors ---------
 * (submap.py:155-164, 182-188, 217-219): boolean gather conf >= thr on the
 * [::stride, ::stride] grid in (s,h,w) order, then the transform.  Outputs may
 * be NULL.  out_world_dev: double[n_sel*3]; out_colors_dev: uint8[n_sel*3];
 * capacity S*ceil(H/stride)*ceil(W/stride) rows.  Synchronises (returns n). */
VSM_API int vsm_select_points(const float* pts_dev, const float* conf_dev, const uint8_t* colors_dev, int32_t S, int32_t H,
                      int32_t W, int32_t stride, float conf_threshold, const double* H_host16, double* out_world_dev,
                      uint8_t* out_colors_dev, int64_t* n_selected_host, void* stream);

/* ---- a6/a7: fusion ------------------------------------------------------ *
 * One iteration of GraphMap.build_semantic_voxel_map's per-submap loop plus
 * its share of the global voxelisation (map.py:196-362), or, without
 * VSM_FUSE_FILTERS, Submap.get_semantic_voxel_in_world_frame (submap.py:246-293).
 * Inputs are device arrays; emb_dev has the map's emb_dtype.  emb_ok_dev is an
 * optional uint8[S*H*W] row mask (1 = embedding row finite) produced by
 * vsm_embedding_row_mask; pass NULL otherwise.  Synchronises; fills stats_host. */
VSM_API int vsm_fuse_submap(vsm_map* m, const float* pts_dev, const float* conf_dev, const void* emb_dev,
                    const uint8_t* emb_ok_dev, const vsm_fuse_params* p, vsm_fuse_stats* stats_host, void* stream);

/* The two halves of vsm_fuse_submap.  _async queues the call's kernels on `stream` and returns without waiting
 * (the inputs must stay valid until the call is collected); _collect synchronises once for all queued calls,
 * repeats the ones the device stopped for lack of room (after growing the map), writes up to max_stats stats
 * in call order, the number of collected calls to n_stats_host, and returns the first failing call's status.
 * vsm_finalize and the exchange calls collect implicitly. */
VSM_API int vsm_fuse_submap_async(vsm_map* m, const float* pts_dev, const float* conf_dev, const void* emb_dev,
                                  const uint8_t* emb_ok_dev, const vsm_fuse_params* p, void* stream);
VSM_API int vsm_fuse_collect(vsm_map* m, vsm_fuse_stats* stats_host, int32_t max_stats, int32_t* n_stats_host,
                             void* stream);

/* Same call with HOST arrays (pinned or pageable): stages geometry first, then
 * streams the embeddings frame by frame through pinned double buffers while
 * the pixel-order accumulate kernel consumes them.  This is the end-to-end
 * entry bench.py's `e2e` times.  emb_host has the map's emb_dtype. */
VSM_API int vsm_fuse_submap_host(vsm_map* m, const float* pts_host, const float* conf_host, const void* emb_host,
                         const vsm_fuse_params* p, vsm_fuse_stats* stats_host, void* stream);

/* Device-time profile of the fuse calls since vsm_profile_enable(m, 1): CUDA events recorded on the caller's
 * stream around the whole call and around the accumulate kernel (bench.py's roofline numbers). */
typedef struct vsm_profile {
  double fuse_ms;             /* device time of the fuse calls, first kernel to last */
  double accumulate_ms;       /* device time of the accumulate kernel launches */
  int64_t fuse_calls;
  int64_t accumulate_launches;
  int64_t accumulate_bytes;   /* algorithmic bytes of those launches: n_fused*d*elem_size + n_submap_voxels*d*4 */
  int64_t points_fused;
} vsm_profile;
VSM_API int vsm_profile_enable(vsm_map* m, int on);
VSM_API int vsm_profile_get(const vsm_map* m, vsm_profile* out_host);

/* uint8[S*H*W] : 1 where conf>=thr (on the stride grid, frame<end_idx) and all d channels are finite (map.py:247) */
VSM_API int vsm_embedding_row_mask(const vsm_map* m, const float* conf_dev, const void* emb_dev, const vsm_fuse_params* p,
                           uint8_t* out_mask_dev, void* stream);

/* ---- finalisation and export (map.py:340-348, 360-362; semantic_voxel.py:31-41) ---- */
/* sorts the voxel keys lexicographically (np.unique(axis=0) order), builds rank maps and the contributor CSR. Synchronises. */
VSM_API int vsm_finalize(vsm_map* m, void* stream);
VSM_API int vsm_num_voxels(const vsm_map* m, int64_t* out_host);
/* contributor-log entries held (one per collected fuse call and voxel it touched, plus merged / drained records) */
VSM_API int vsm_num_log_entries(const vsm_map* m, int64_t* out_host);
/* sorted order; any pointer may be NULL.  coords int64[V*3] true keys, centers float[V*3] = (coords+0.5)*vs,
 * counts int64[V], recon_coords int64[V*3] = floor(centers/vs - 0.5) (semantic_voxel.py:62-66, lossy) */
VSM_API int vsm_export_geometry(const vsm_map* m, int64_t* coords_dev, float* centers_dev, int64_t* counts_dev,
                        int64_t* recon_coords_dev, void* stream);
/* rows [r0,r1) of the sorted feature matrix: float[(r1-r0)*d] = sum / count (map.py:340, 360) */
VSM_API int vsm_export_features(const vsm_map* m, int64_t r0, int64_t r1, float* out_dev, void* stream);
/* contributor CSR in sorted voxel order: offsets int64[V+1]; entries: submap id int32[M], frame mask uint64[M*2]
 * (bit f = frame f of that submap contributed).  n_entries via vsm_num_contributor_entries. */
VSM_API int vsm_num_contributor_entries(const vsm_map* m, int64_t* out_host);
VSM_API int vsm_export_contributors(const vsm_map* m, int64_t* offsets_dev, int32_t* submap_ids_dev, uint64_t* masks_dev,
                            void* stream);
/* np.unique's inverse for the fuse call number `fuse_index` (needs VSM_FUSE_KEEP_POINT_INDEX): int32[S*H*W],
 * sorted voxel index of each pixel's point or -1 if the pixel was not fused */
VSM_API int vsm_export_point_index(const vsm_map* m, int32_t fuse_index, int32_t* out_dev, int64_t n_pixels, void* stream);

/* a loaded map (SemanticVoxelMap.load_from_directory, semantic_voxel.py:150-165): rows are taken in file order */
VSM_API int vsm_map_load_dense(vsm_map* m, const float* centers_dev, const float* features_dev, int64_t V, void* stream);

/* The same load in row blocks, for files larger than the device can hold twice: begin sizes the map for V rows, rows
 * copies rows [r0, r0+n_rows) (centres float[n_rows*3], features float[n_rows*d], device pointers), vsm_finalize ends it. */
VSM_API int vsm_map_load_begin(vsm_map* m, int64_t V, void* stream);
VSM_API int vsm_map_load_rows(vsm_map* m, int64_t r0, int64_t n_rows, const float* centers_dev, const float* features_dev,
                              void* stream);

/* ---- a12: position -> voxel index (semantic_voxel.py:68-80) -------------- *
 * compat != 0 reproduces the reference's table keyed by the lossy reconstructed coordinates
 * (later index wins on collisions); compat == 0 uses the true keys.  idx = -1 if absent. */
VSM_API int vsm_lookup(vsm_map* m, const float* pos_dev, int64_t M, int64_t* idx_dev, int compat, void* stream);

/* ---- a13: query_with_embedding (semantic_voxel.py:97-116), batched over P prompts ---- *
 * scores[p, v] = features[v] . q[p]  (features = sum / count; normalize != 0 divides by ||features[v]||).
 * Returns per prompt the top-k sorted voxel indices (ties: lower index first) and float32 scores.
 * q_dev float[P*d]; idx_dev int64[P*k]; score_dev float[P*k].
 * engine: 0 = auto, 1 = exact fp32 CUDA-core kernel, 2 = tcgen05 tensor-core kernel (TF32 on the fp32 sums) + exact
 * fp32 rescoring, 3 = the same on a bf16 shadow of the sums (tcgen05 kind::f16; half the bytes per query, +2 bytes per
 * value of memory, built on the first query after a finalisation).  Every engine returns the same answer.
 * vsm_set_option("query_shadow", 1) makes engine 0 choose 3 on large maps. */
VSM_API int vsm_query(vsm_map* m, const float* q_dev, int32_t P, int32_t k, int normalize, int engine, int64_t* idx_dev,
              float* score_dev, void* stream);

/* engine-2 diagnostics: longest candidate list of the last tensor-core query, and how many engine-2 queries were
 * answered by engine 1 because a candidate list overflowed */
VSM_API int vsm_query_stats(const vsm_map* m, int64_t* last_candidates_host, int64_t* fallbacks_host);

/* frees the bf16 shadow of engine 3 (it is rebuilt by the next engine-3 query) */
VSM_API int vsm_query_shadow_release(vsm_map* m);

/* ---- multi-GPU exchange (no reference counterpart) ------------- *
 * Voxels are owned by mix64(key) % world.  pack: groups this map's voxels by owner and writes
 * keys uint64[V], counts uint32[V], sums float[V*d] in owner-major order, per-owner counts to counts_host[world].
 * merge: adds received partial voxels into this map (insert key, add count and sums). */
VSM_API int vsm_partials_pack(vsm_map* m, int32_t world, uint64_t* keys_dev, uint32_t* counts_dev, float* sums_dev,
                      int64_t* owner_counts_host, void* stream);
VSM_API int vsm_partials_merge(vsm_map* m, const uint64_t* keys_dev, const uint32_t* counts_dev, const float* sums_dev,
                       int64_t n, void* stream);
/* contributor log in pack order: for each log entry its voxel key, submap id and frame mask */
VSM_API int vsm_contrib_pack(vsm_map* m, int32_t world, uint64_t* keys_dev, int32_t* submap_ids_dev, uint64_t* masks_dev,
                     int64_t* owner_counts_host, void* stream);
VSM_API int vsm_contrib_merge(vsm_map* m, const uint64_t* keys_dev, const int32_t* submap_ids_dev, const uint64_t* masks_dev,
                      int64_t n, void* stream);

/* ---- one-sided exchange over NVLink peer memory (csrc/peer.cu) ------------- *
 * Every rank owns an inbox (vsm_peer_alloc: cudaMalloc + CUDA IPC handle) that its peers map (vsm_peer_open).
 * vsm_partials_push: ONE pass groups this map's voxels and contributor entries by owner and stores the records
 * straight into the owners' inboxes (remote slot reservation + 16-byte peer stores), then signals every owner.
 * inbox_ptrs_host[world]: device pointers of all inboxes as seen from this process (own inbox included).
 * vsm_partials_drain: waits on the device for `world` signals (bounded by timeout_s), inserts the received records
 * into this map and resets the inbox half.  Exchanges alternate halves: pass epoch = 0, 1, 2, ... on every rank.
 * Returns VSM_E_NOMEM if the inbox was too small (n_rows_host / n_contrib_host tell the sizes needed), VSM_E_STATE on
 * timeout.  flags_host: 1 row overflow, 2 contributor overflow, 4 timeout, 8 hash error. */
VSM_API int vsm_inbox_bytes(int32_t dim, int64_t cap_rows, int64_t cap_contrib, int64_t* bytes_host);
VSM_API int vsm_peer_alloc(int32_t device, int64_t bytes, void** ptr_out, void* ipc_handle_out /* 64 bytes or NULL */);
VSM_API int vsm_peer_open(int32_t device, const void* ipc_handle /* 64 bytes */, void** ptr_out);
VSM_API int vsm_peer_close(int32_t device, void* ptr);
VSM_API int vsm_peer_free(int32_t device, void* ptr);
VSM_API int vsm_partials_push(vsm_map* m, int32_t world, void* const* inbox_ptrs_host, int64_t cap_rows, int64_t cap_contrib,
                      int64_t epoch, void* stream);
VSM_API int vsm_partials_drain(vsm_map* m, void* inbox, int32_t world, int64_t cap_rows, int64_t cap_contrib, int64_t epoch,
                       double timeout_s, int64_t* n_rows_host, int64_t* n_contrib_host, uint32_t* flags_host, void* stream);
/* The two halves of vsm_partials_drain for exchanges that run BESIDE fusion (rounds of a long build: the exchange of
 * round r on one stream while round r+1 is fused on another).  _async queues the wait + merge on `stream` and returns;
 * the map must already have room for what can arrive (vsm_map_reserve, vsm_map_reserve_log): it is not grown.
 * _collect synchronises `stream`, reads the report of slot `report_slot` (0..3, one per drain in flight), updates the
 * map's voxel count and maps overflow / timeout to the same status codes as vsm_partials_drain. */
VSM_API int vsm_partials_drain_async(vsm_map* m, void* inbox, int32_t world, int64_t cap_rows, int64_t cap_contrib,
                                     int64_t epoch, double timeout_s, int32_t report_slot, void* stream);
VSM_API int vsm_partials_drain_collect(vsm_map* m, int32_t report_slot, int64_t* n_rows_host, int64_t* n_contrib_host,
                                       uint32_t* flags_host, void* stream);

/* sorted packed keys of this map's voxels (uint64[V]); pack/unpack helpers for global ranking across owners */
VSM_API int vsm_export_packed_keys(const vsm_map* m, uint64_t* keys_dev, void* stream);

/* global index of this shard's keys among the keys of all shards (disjoint, each sorted): all_keys_dev is the
 * all-gather buffer, shard r = all_keys_dev[r*shard_stride .. r*shard_stride + shard_sizes_host[r]); ranks_dev int64[n_mine] */
VSM_API int vsm_global_ranks(const uint64_t* my_keys_dev, int64_t n_mine, const uint64_t* all_keys_dev, int64_t shard_stride,
                             const int64_t* shard_sizes_host, int32_t world, int64_t* ranks_dev, void* stream);

/* ---- f3: scoring of the projective RANSAC hypotheses (h_solve.py:16-41, 150-160) -------------------------------- *
 * H_dev float[B*16] row-major 4x4 hypotheses, X1_dev / X2_dev float[N*3]: counts_dev[b] = #{n : ||(H_b [X1_n;1])_xyz / w
 * - X2_n||_2 < threshold} in float32 arithmetic; best_dev int32[2] = (argmax, its count), first maximum wins like
 * torch.argmax.  best_idx_host / best_count_host may be NULL (then the call does not synchronise). */
VSM_API int vsm_ransac_score(const float* H_dev, const float* X1_dev, const float* X2_dev, int64_t N, int32_t B,
                             float threshold, int32_t* counts_dev, int32_t* best_dev, int32_t* best_idx_host,
                             int32_t* best_count_host, void* stream);

/* ---- f4: occupancy grid of a point cloud (get_occupancy.py:130-179 build_occupancy_from_pointcloud) ------------- *
 * pts_dev float[n*3].  Cells in np.unique(axis=0) order: centers_dev float[cells*3], blocked_dev uint8[cells],
 * keys_dev int64[cells*2], minz_dev float[cells]; outputs may be NULL.  n_cells_host receives the number of cells
 * (call with cap_cells = 0 and NULL outputs to size the arrays; VSM_E_NOMEM if cap_cells is too small), n_kept_host
 * the points that survived the finite / ceiling filters.  Synchronises. */
VSM_API int vsm_occupancy_build(const float* pts_dev, int64_t n, double voxel_size, double ceiling_z, double height_thresh,
                                int64_t cap_cells, float* centers_dev, uint8_t* blocked_dev, int64_t* keys_dev,
                                float* minz_dev, int64_t* n_cells_host, int64_t* n_kept_host, void* stream);

/* ---- f2: producer hand-off (solver.py:249-263, 301, 337-340, 478-480) on the device ------------------------------ *
 * vsm_unproject_depth: depth (S,H,W) float32, cam_to_world double[S*12] (row-major 3x4 [R|t], the closed-form inverse
 * of the extrinsic), intrinsic float[S*9] -> world points (S,H,W,3), float32 or float64 (vggt.utils.geometry's
 * unproject_depth_map_to_point_map restated; that dependency is not vendored upstream: parity unpinned).
 * vsm_images_to_colors: images (S,3,H,W) float32 in [0,1] -> colors (S,H,W,3) uint8 = (img * 255).astype(uint8).
 * vsm_scale_points: p *= scale over n_floats floats (product in float64, stored float32, as numpy's in-place op). */
VSM_API int vsm_unproject_depth(const float* depth_dev, const double* cam_to_world_dev, const float* intrinsic_dev, int32_t S,
                                int32_t H, int32_t W, void* out_dev, int out_f64, void* stream);
VSM_API int vsm_images_to_colors(const float* images_dev, int32_t S, int32_t H, int32_t W, uint8_t* colors_dev, void* stream);
VSM_API int vsm_scale_points(float* pts_dev, int64_t n_floats, double scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VSM_H_ */
