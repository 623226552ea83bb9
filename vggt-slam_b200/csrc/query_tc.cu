// query_tc.cu -- tensor-core engine of vsm_query (engine 2): the (voxels x 512) x (512 x prompts) contraction on
// Blackwell's 5th-generation tensor cores, hand-written for sm_100a:
//   * TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) stages 128-voxel x 32-channel fp32 tiles of the voxel sums into
//     shared memory through a 4-stage mbarrier pipeline.  Up to 64 prompts per pass: the prompt block (64 x d) is
//     loaded once per CTA and stays resident.  128 or 256 prompts per pass: the matching 32-channel slice of the
//     prompt block (16 / 32 KB, L2-resident) travels with every voxel stage, so that 256 prompts cost ONE pass over
//     the voxels instead of four;
//   * one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=64/128/256, K=8 per instruction) with the fp32
//     accumulators in tensor memory (two N-column accumulators -- all 512 columns at N=256: the epilogue of tile t
//     overlaps the MMAs of t+1);
//   * four epilogue warps read the accumulators back with tcgen05.ld, scale by 1/count (or the cosine
//     normaliser) and compare with a per-prompt threshold; survivors are appended to per-prompt candidate lists.
// TF32 drops mantissa bits, so the tensor-core scores only SELECT candidates, conservatively:
//   1. thresholds: the exact fp32 engine scores a strided sample of the voxels; the k-th best sample score T_p is
//      a lower bound of the k-th best score overall;
//   2. the tensor-core pass keeps voxel v for prompt p if  a(v,p) + 1.5 * 2^-9 * ||f_v|| * ||q_p|| >= T_p  (2^-9 ||f|| ||q||
//      bounds the TF32 product error summed over d channels, Cauchy-Schwarz; see kTf32Margin);
//   3. the candidates are re-scored exactly in fp32 and the top-k taken with the same keys as engine 1.
// The result is therefore identical to engine 1 (vggt_slam/semantic_voxel.py:97-116 semantics).  If a candidate
// list overflows, the call falls back to engine 1.
#include <cuda.h>
#include <cuda_bf16.h>

#include "state.cuh"

namespace vsm {

int query_exact_rows(vsm_map* m, const float* q_dev, int P, int k, int normalize, uint32_t row_stride, uint32_t n_rows,
                     int64_t* idx_dev, float* score_dev, cudaStream_t s);  // query.cu
int query_exact(vsm_map* m, const float* q_dev, int P, int k, int normalize, int64_t* idx_dev, float* score_dev,
                cudaStream_t s);
int query_select_from_keys(vsm_map* m, const unsigned long long* keys, const uint32_t* counts, int P, int cap, int k,
                           int64_t* idx_dev, float* score_dev, cudaStream_t s);

namespace tc {

constexpr int kTileM = 128;      // voxels per tile (MMA M)
constexpr int kMaxTileN = 256;   // prompts per pass (MMA N): 64 (resident prompt block), 128 or 256 (streamed)
constexpr int kChunkK = 32;      // fp32 channels per pipeline stage: 128 bytes = one swizzle row
constexpr uint32_t kHalfBytes = kTileM * kChunkK * 4;  // 128 voxels x 32 channels: 16 KB
// |a_tf32 . b_tf32 - a . b| <= sum |a_i b_i| (2 eps + eps^2) <= (2^-9 + 2^-20) ||a|| ||b||  for eps = 2^-10 (the tensor core
// reads 10 explicit mantissa bits of each fp32 operand) -- Cauchy-Schwarz; the fp32 accumulation of d <= 1024 terms adds
// less than 1e-4 ||a|| ||b||.  1.5 x 2^-9 leaves half of the bound as slack.
constexpr float kTf32Margin = 1.5f * 0.001953125f;
// bf16 shadow (engine 3): both operands are ROUNDED to 8 significant bits (relative error <= 2^-9 each), so
// |a_bf . b_bf - a . b| <= sum |a_i b_i| (2 * 2^-9 + 2^-18) <= (2^-8 + 2^-18) ||a|| ||b||; 1.5 x 2^-8 again leaves half of the
// bound as slack (and covers the fp32 accumulation, < 1e-4 ||a|| ||b||).
constexpr float kBf16Margin = 1.5f * 0.00390625f;
constexpr int kWarpStage = 128;  // candidates an epilogue warp stages in shared memory before one global append

// TN prompts per pass.  Streamed prompt slices are re-read from L2 for every voxel tile, so a streamed tile takes
// 256 voxels (two M=128 MMAs per K step share one slice): half the L2 traffic per voxel.
template <int TN>
struct Cfg {
  static constexpr bool kStream = TN > 64;                      // prompt slices travel with the voxel stages
  // (build with -DVSM_TC256_MH=1 for 128-voxel, double-buffered tiles at 256 prompts: measured the same 6.6 ms per call)
#ifndef VSM_TC256_MH
#define VSM_TC256_MH 2
#endif
  static constexpr int kMH = !kStream ? 1 : (TN == 256 ? VSM_TC256_MH : 2);  // M=128 halves per tile
  static constexpr int kTileRows = kMH * kTileM;
  static constexpr uint32_t kABytes = kMH * kHalfBytes;
  static constexpr uint32_t kBChunkBytes = TN * kChunkK * 4;    // one 32-channel slice of the prompt block
  static constexpr uint32_t kStageBytes = kABytes + (kStream ? kBChunkBytes : 0u);
  static constexpr int kStages = !kStream ? 4 : (TN == 128 ? 4 : (kMH == 2 ? 3 : 4));
  static constexpr int kAccBufs = (2 * kMH * TN <= 512) ? 2 : 1;  // accumulator sets in the 512 TMEM columns
  static constexpr int kTmemCols = kAccBufs * kMH * TN;
  static constexpr int kThreads = 128 + 128 * kMH;              // warp 0 TMA, 1 MMA, 2 TMEM alloc, 4.. epilogue
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  uint32_t spins = 0;
  do {
    if (++spins == 0x20000000u) __trap();  // minutes of polling: a protocol error must not hang the GPU
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// ---- CTA pairs (cta_group::2): two SMs execute one MMA of M = 256; every barrier the MMA thread waits on lives in the
// leader CTA (rank 0).  Shared-memory addresses of the two CTAs of a pair differ in bit 24: clearing it names the
// leader's copy of a variable from either CTA.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// both CTAs of the pair issue their own loads; the bytes are counted on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T with M = 256: each CTA supplies its 128 rows of A and its half of B's rows
__device__ __forceinline__ void mma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0u;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
// arrives, when the pair's MMAs issued so far are done, on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, TF32 inputs, fp32 accumulate; issued by one thread for the CTA
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same with BF16 inputs (kind::f16, K = 16 per instruction)
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0u;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
// arrives on the mbarrier when all MMAs issued so far by this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the same, naming the registers an earlier tcgen05.ld fills: the compiler must not read them before this point
__device__ __forceinline__ void tmem_ld_wait_for(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B (8-row x 128-byte atoms, 1024 bytes apart along M/N)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                        // leading byte offset (unused: one swizzle atom along K)
  d |= (uint64_t)(1024u >> 4) << 32;             // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                        // layout type SWIZZLE_128B
  return d;
}
// instruction descriptor: D=F32, A=B=TF32 (kind::tf32 format 2) or BF16 (kind::f16 format 1), both K-major, N=tile_n, M=128
__host__ __device__ constexpr uint32_t make_idesc(int tile_n, int tile_m = kTileM, bool bf16 = false) {
  return (1u << 4) | ((bf16 ? 1u : 2u) << 7) | ((bf16 ? 1u : 2u) << 10) | ((uint32_t)(tile_n >> 3) << 17) |
         ((uint32_t)(tile_m >> 4) << 24);
}

struct TcArgs {
  const uint32_t* vcount;
  const float* vnorm;    // ||sum_v||_2 per voxel id
  const float* thr;      // per prompt of this pass: sample threshold T_p (already in the scored units)
  const float* qnorm;    // per prompt of this pass: ||q_p||_2
  uint32_t n_rows;       // voxel rows scored by this launch: ids 0, row_stride, 2*row_stride, ...
  uint32_t row_stride;
  int pb;                // prompts in this pass (<= TN)
  int p0;                // first prompt of the pass
  int normalize;
  unsigned long long* pairs;  // flat candidate list: (prompt << 32 | voxel id), in arrival order
  uint32_t* pair_cnt;         // [1] entries appended (can exceed pair_cap: overflow)
  uint32_t pair_cap;
  int n_kchunks;         // 128-byte slices of a row: d / 32 (fp32 sums) or d / 64 (bf16 shadow)
  float margin;          // kTf32Margin or kBf16Margin
};

template <int TN, bool BF>
__global__ void __launch_bounds__(Cfg<TN>::kThreads, 1)
query_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, TcArgs a) {
  using C = Cfg<TN>;
  constexpr int kChunkE = BF ? 2 * kChunkK : kChunkK;  // elements of a 128-byte slice
  constexpr int kStages = C::kStages;
  constexpr int kThreads = C::kThreads;
  constexpr uint32_t kABytes = C::kABytes;
  extern __shared__ __align__(1024) uint8_t smem[];
  // layout, resident prompts: [B: n_kchunks x 8 KB][A: kStages x 16 KB][barriers]
  //         streamed prompts: [kStages x (A 32 KB | B slice)][barriers]
  uint8_t* smem_b = smem;
  uint8_t* smem_a = C::kStream ? smem : smem + (size_t)a.n_kchunks * C::kBChunkBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + (size_t)kStages * C::kStageBytes);
  uint64_t* full = bars;                   // [kStages] TMA -> MMA
  uint64_t* empty = bars + kStages;        // [kStages] MMA -> TMA
  uint64_t* tfull = bars + 2 * kStages;    // [2] MMA -> epilogue
  uint64_t* tempty = tfull + 2;            // [2] epilogue -> MMA
  uint64_t* bfull = tempty + 2;            // [1] prompt block resident
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bfull + 1);
  __shared__ float2 s_col[TN];  // per prompt column: (||q_p||, threshold T_p); padding columns can never pass
  __shared__ unsigned long long s_stage[4 * C::kMH][kWarpStage];  // candidates staged per epilogue warp
  __shared__ uint32_t s_wcnt[4 * C::kMH];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_tiles = (a.n_rows + C::kTileRows - 1) / C::kTileRows;

  for (int i = threadIdx.x; i < TN; i += kThreads)
    s_col[i] = i < a.pb ? make_float2(a.qnorm[i], a.thr[i]) : make_float2(0.f, __int_as_float(0x7f800000));
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4 * C::kMH);  // one arrive per epilogue warp
    }
    mbar_init(bfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_smem, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    if (!C::kStream) {
      mbar_expect_tx(bfull, (uint32_t)a.n_kchunks * C::kBChunkBytes);
      for (int kc = 0; kc < a.n_kchunks; ++kc)
        tma_load_2d(smem_b + (size_t)kc * C::kBChunkBytes, &map_b, bfull, kc * kChunkE, 0);
    }
    uint32_t stage = 0, phase = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int kc = 0; kc < a.n_kchunks; ++kc) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], C::kStageBytes);
        uint8_t* st = smem_a + (size_t)stage * C::kStageBytes;
        tma_load_2d(st, &map_a, &full[stage], kc * kChunkE, (int)(tile * C::kTileRows));
        if (C::kStream) tma_load_2d(st + kABytes, &map_b, &full[stage], kc * kChunkE, 0);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc(TN, kTileM, BF);
    if (!C::kStream) mbar_wait(bfull, 0);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);  // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * (C::kMH * TN);
      for (int kc = 0; kc < a.n_kchunks; ++kc) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        uint8_t* st = smem_a + (size_t)stage * C::kStageBytes;
        const uint64_t bdesc = make_smem_desc(smem_u32(C::kStream ? st + kABytes : smem_b + (size_t)kc * C::kBChunkBytes));
#pragma unroll
        for (int h = 0; h < C::kMH; ++h) {
          const uint64_t adesc = make_smem_desc(smem_u32(st + (size_t)h * kHalfBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // K = 8 tf32 / 16 bf16 (32 bytes) per instruction: advance 32 bytes inside the atom
            if (BF)
              mma_bf16(tmem_d + h * TN, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kc | k) ? 1u : 0u);
            else
              mma_tf32(tmem_d + h * TN, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kc | k) ? 1u : 0u);
          }
        }
        mma_commit(&empty[stage]);  // frees the smem stage when these MMAs are done
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      mma_commit(&tfull[acc]);  // accumulator complete
      if (++acc == C::kAccBufs) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> threshold test -> candidate lists =====
    const int ew = warp & 3;        // TMEM lane quarter this warp may read
    const int mh = (warp - 4) >> 2;  // which M=128 half of the tile
    // A candidate costs a slot in the global list.  Reserving it with a global atomic per hit stalls the warp for a
    // full round trip per candidate (measured: the epilogue of a 256-prompt tile took longer than its MMAs), so
    // each warp stages its hits in shared memory (a shared-memory counter) and appends ~64 at a time.
    unsigned long long* wbuf = s_stage[warp - 4];
    uint32_t* wcnt = &s_wcnt[warp - 4];  // staged entries of this warp
    if (lane == 0) *wcnt = 0u;
    __syncwarp();
    auto flush = [&]() {
      __syncwarp();
      const uint32_t wn = min(*wcnt, (uint32_t)kWarpStage);
      if (wn) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.pair_cnt, wn);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (uint32_t i = lane; i < wn; i += 32)
          if (base + i < a.pair_cap) a.pairs[base + i] = wbuf[i];
      }
      __syncwarp();
      if (lane == 0) *wcnt = 0u;
      __syncwarp();
    };
    uint32_t acc = 0, acc_phase = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const uint32_t row = tile * C::kTileRows + mh * kTileM + ew * 32 + lane;
      const uint32_t id = row * a.row_stride;
      float inv = 0.f, fn = 0.f;
      const bool valid = row < a.n_rows;
      if (valid) {
        const float cnt = (float)a.vcount[id];
        const float nrm = a.vnorm[id];
        fn = __fdiv_rn(nrm, cnt);  // ||f_v||
        inv = a.normalize ? __fdiv_rn(1.0f, fmaxf(nrm, 1e-12f * cnt)) : __fdiv_rn(1.0f, cnt);
        if (a.normalize) fn = (nrm == nrm) ? 1.0f : nrm;  // scored quantity is f/||f||: unit norm (NaN sums stay NaN)
      }
      const float margin = a.margin * fn;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (acc * C::kMH + mh) * TN;
      const int n_chunks = min(TN, (a.pb + 31) & ~31) / 32;  // columns beyond the pass's prompts are padding
      // keep (v, p) iff  x = a(v,p)/count + margin_v * ||q_p|| - T_p  is not < 0 (a NaN passes, like torch.topk ranks it
      // first).  Per score: one shared-memory read, two FMAs, one compare, one bit into the lane's hit mask.  Hits are
      // rare (< 1e-3): the set bits are walked afterwards, each taking a slot of the warp's staging buffer.  The next
      // chunk's accumulators are already on their way from tensor memory while this one is tested.
      uint32_t r[2][32];
      tmem_ld32(taddr, r[0]);
#pragma unroll 1
      for (int c = 0; c < n_chunks; c += 2) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (c + h >= n_chunks) break;  // warp-uniform
          tmem_ld_wait_for(r[h]);
          if (c + h + 1 < n_chunks) tmem_ld32(taddr + (c + h + 1) * 32, r[h ^ 1]);
          const int c0 = (c + h) * 32;
          uint32_t hits = 0u;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float2 col = s_col[c0 + j];
            const float x = fmaf(__uint_as_float(r[h][j]), inv, fmaf(margin, col.x, -col.y));
            hits |= (!(x < 0.f)) ? (1u << j) : 0u;
          }
          if (!valid) hits = 0u;
          while (hits) {
            const int j = __ffs(hits) - 1;
            hits &= hits - 1;
            const unsigned long long e = ((unsigned long long)(uint32_t)(a.p0 + c0 + j) << 32) | id;
            const uint32_t pos = atomicAdd(wcnt, 1u);
            if (pos < (uint32_t)kWarpStage) {
              wbuf[pos] = e;
            } else {  // staging buffer full (dense ties): straight to the list
              const uint32_t g = atomicAdd(a.pair_cnt, 1u);
              if (g < a.pair_cap) a.pairs[g] = e;
            }
          }
          __syncwarp();
          if (*wcnt > (uint32_t)(kWarpStage / 2)) flush();  // warp-uniform: every lane reads the same counter
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == C::kAccBufs) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
}

// ---- 128 / 256 prompts on CTA pairs ------------------------------------------------------------------------------------
// One tile = 256 voxels x TN prompts on TWO SMs: each CTA stages its 128 voxels and TN/2 prompts (its half of
// the prompt slice: half the L2 reads, half the shared-memory operand traffic per SM), the leader's elected thread
// issues tcgen05.mma.cta_group::2 (M = 256), each CTA's tensor memory receives its 128 rows x 256 columns -- which
// leaves room for TWO accumulator sets, so the epilogue of tile t overlaps the MMAs of tile t+1 again.
template <int TN>
struct PairCfg {
  static constexpr int kTN = TN;                                     // prompts per pass: 128 or 256
  static constexpr uint32_t kABytes = kHalfBytes;                    // my 128 voxels x 32 channels
  static constexpr uint32_t kBBytes = (TN / 2) * kChunkK * 4;        // my half of the prompts x 32 channels
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;         // 32 / 24 KB per CTA and stage
  static constexpr int kStages = TN == 256 ? 6 : 8;
  static constexpr int kThreads = 256;  // warp 0 TMA, 1 MMA (leader), 2 TMEM alloc, 4..7 epilogue
  static constexpr int kTmemCols = 2 * TN;
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 256;
};

template <int TN, bool BF>
__global__ void __launch_bounds__(256, 1)
query_tc_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, TcArgs a) {
  using PC = PairCfg<TN>;
  constexpr int kChunkE = BF ? 2 * kChunkK : kChunkK;
  constexpr int kTN = PC::kTN, kStages = PC::kStages, kThreads = PC::kThreads, kTmemCols = PC::kTmemCols;
  constexpr uint32_t kABytes = PC::kABytes, kStageBytes = PC::kStageBytes;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
  uint64_t* full = bars;                 // [kStages] both producers -> leader's MMA thread (leader's copy is used)
  uint64_t* empty = bars + kStages;      // [kStages] MMA -> the producer of this CTA
  uint64_t* tfull = bars + 2 * kStages;  // [2] MMA -> the epilogue of this CTA
  uint64_t* tempty = tfull + 2;          // [2] both epilogues -> leader's MMA thread (leader's copy is used)
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tempty + 2);
  __shared__ float2 s_col[kTN];
  __shared__ unsigned long long s_stage[4][kWarpStage];
  __shared__ uint32_t s_wcnt[4];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t n_tiles = (a.n_rows + 2 * kTileM - 1) / (2 * kTileM);
  const uint32_t pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;

  for (int i = threadIdx.x; i < kTN; i += kThreads)
    s_col[i] = i < a.pb ? make_float2(a.qnorm[i], a.thr[i]) : make_float2(0.f, __int_as_float(0x7f800000));
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 2);   // one arrive.expect_tx per producer of the pair
      mbar_init(&empty[i], 1);  // the pair's commit, multicast to both CTAs
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);  // four epilogue warps in each CTA
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_base_smem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers and tensor memory exist before anything is signalled across
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer (both CTAs) =====
    uint32_t stage = 0, phase = 0;
    for (uint32_t tile = pair0; tile < n_tiles; tile += pair_step) {
      for (int kc = 0; kc < a.n_kchunks; ++kc) {
        mbar_wait(&empty[stage], phase ^ 1);
        const uint32_t lfull = smem_u32(&full[stage]) & kPeerBitMask;
        mbar_expect_tx_cluster(lfull, kStageBytes);
        uint8_t* st = smem + (size_t)stage * kStageBytes;
        tma_load_2d_pair(st, &map_a, lfull, kc * kChunkE, (int)(tile * 2 * kTileM + crank * kTileM));
        tma_load_2d_pair(st + kABytes, &map_b, lfull, kc * kChunkE, (int)(crank * (kTN / 2)));
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && lane == 0 && crank == 0) {
    // ===== MMA issuer (leader CTA only) =====
    constexpr uint32_t idesc = make_idesc(kTN, 2 * kTileM, BF);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (uint32_t tile = pair0; tile < n_tiles; tile += pair_step) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);  // both epilogues have drained this accumulator set
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * kTN;
      for (int kc = 0; kc < a.n_kchunks; ++kc) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        uint8_t* st = smem + (size_t)stage * kStageBytes;
        const uint64_t adesc = make_smem_desc(smem_u32(st));
        const uint64_t bdesc = make_smem_desc(smem_u32(st + kABytes));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (BF)
            mma_bf16_pair(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kc | k) ? 1u : 0u);
          else
            mma_tf32_pair(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kc | k) ? 1u : 0u);
        }
        mma_commit_pair(&empty[stage], 3);  // both producers may refill this stage
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      mma_commit_pair(&tfull[acc], 3);  // both epilogues may read this accumulator set
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): as query_tc_kernel's, on this CTA's 128 rows =====
    const int ew = warp & 3;
    unsigned long long* wbuf = s_stage[warp - 4];
    uint32_t* wcnt = &s_wcnt[warp - 4];
    if (lane == 0) *wcnt = 0u;
    __syncwarp();
    auto flush = [&]() {
      __syncwarp();
      const uint32_t wn = min(*wcnt, (uint32_t)kWarpStage);
      if (wn) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.pair_cnt, wn);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (uint32_t i = lane; i < wn; i += 32)
          if (base + i < a.pair_cap) a.pairs[base + i] = wbuf[i];
      }
      __syncwarp();
      if (lane == 0) *wcnt = 0u;
      __syncwarp();
    };
    uint32_t acc = 0, acc_phase = 0;
    for (uint32_t tile = pair0; tile < n_tiles; tile += pair_step) {
      const uint32_t row = tile * 2 * kTileM + crank * kTileM + ew * 32 + lane;
      const uint32_t id = row * a.row_stride;
      float inv = 0.f, fn = 0.f;
      const bool valid = row < a.n_rows;
      if (valid) {
        const float cnt = (float)a.vcount[id];
        const float nrm = a.vnorm[id];
        fn = __fdiv_rn(nrm, cnt);
        inv = a.normalize ? __fdiv_rn(1.0f, fmaxf(nrm, 1e-12f * cnt)) : __fdiv_rn(1.0f, cnt);
        if (a.normalize) fn = (nrm == nrm) ? 1.0f : nrm;
      }
      const float margin = a.margin * fn;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * kTN;
      const int n_chunks = min(kTN, (a.pb + 31) & ~31) / 32;
      uint32_t r[2][32];
      tmem_ld32(taddr, r[0]);
#pragma unroll 1
      for (int c = 0; c < n_chunks; c += 2) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (c + h >= n_chunks) break;
          tmem_ld_wait_for(r[h]);
          if (c + h + 1 < n_chunks) tmem_ld32(taddr + (c + h + 1) * 32, r[h ^ 1]);
          const int c0 = (c + h) * 32;
          uint32_t hits = 0u;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float2 col = s_col[c0 + j];
            const float x = fmaf(__uint_as_float(r[h][j]), inv, fmaf(margin, col.x, -col.y));
            hits |= (!(x < 0.f)) ? (1u << j) : 0u;
          }
          if (!valid) hits = 0u;
          while (hits) {
            const int j = __ffs(hits) - 1;
            hits &= hits - 1;
            const unsigned long long e = ((unsigned long long)(uint32_t)(a.p0 + c0 + j) << 32) | id;
            const uint32_t pos = atomicAdd(wcnt, 1u);
            if (pos < (uint32_t)kWarpStage) {
              wbuf[pos] = e;
            } else {
              const uint32_t g = atomicAdd(a.pair_cnt, 1u);
              if (g < a.pair_cap) a.pairs[g] = e;
            }
          }
          __syncwarp();
          if (*wcnt > (uint32_t)(kWarpStage / 2)) flush();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(smem_u32(&tempty[acc]) & kPeerBitMask);  // the leader's copy
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    flush();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves (or frees tensor memory) while the peer may still use the pair's resources
  if (warp == 2) tmem_dealloc_pair(tmem_base, kTmemCols);
}

// ---- helpers ---------------------------------------------------------------------------------------------------
// ||sum_v||_2 per voxel id (one warp per row)
__global__ void __launch_bounds__(256) row_norm_kernel(const float* __restrict__ vsum, uint32_t V, int d, float* __restrict__ out) {
  const int lane = lane_id();
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t id = warp; id < V; id += n_warps) {
    float ss = 0.f;
    for (int c = lane; c < d / 4; c += 32) {
      const uint4 u = ld_stream_v4(vsum + (size_t)id * d + 4 * c);
      const float x = __uint_as_float(u.x), y = __uint_as_float(u.y), z = __uint_as_float(u.z), w = __uint_as_float(u.w);
      ss = fmaf(x, x, fmaf(y, y, fmaf(z, z, fmaf(w, w, ss))));
    }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) out[id] = sqrtf(ss);
  }
}

// per prompt: ||q_p|| and the threshold T_p = k-th best sample score (or -inf if the sample has fewer than k rows)
__global__ void prompt_prep_kernel(const float* __restrict__ q, int P, int d, const float* __restrict__ sample_scores, int k,
                                   int have_sample, float* __restrict__ qnorm, float* __restrict__ thr) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float ss = 0.f;
  for (int c = 0; c < d; ++c) ss = fmaf(q[(size_t)p * d + c], q[(size_t)p * d + c], ss);
  qnorm[p] = sqrtf(ss);
  float t = __int_as_float(0xff800000);  // -inf
  if (have_sample) {
    const float s = sample_scores[(size_t)p * k + (k - 1)];
    if (s == s) t = s;  // NaN (fewer than k sample rows, or NaN scores) -> keep everything
  }
  thr[p] = t;
}

// zero-padded copy of the prompts of one pass: [tile_n][d]
__global__ void pad_prompts_kernel(const float* __restrict__ q, int p0, int pb, int d, int tile_n, float* __restrict__ out) {
  const int total = tile_n * d;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int p = i / d;
    out[i] = p < pb ? q[(size_t)(p0 + p) * d + (i % d)] : 0.f;
  }
}

// bf16 shadow of the voxel sums (engine 3): round-to-nearest-even, 8 values per thread
__global__ void __launch_bounds__(256) shadow_kernel(const float* __restrict__ vsum, size_t n8, __nv_bfloat16* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(vsum)[2 * i], b = reinterpret_cast<const float4*>(vsum)[2 * i + 1];
    __nv_bfloat162 o[4] = {__floats2bfloat162_rn(a.x, a.y), __floats2bfloat162_rn(a.z, a.w), __floats2bfloat162_rn(b.x, b.y),
                           __floats2bfloat162_rn(b.z, b.w)};
    reinterpret_cast<uint4*>(out)[i] = *reinterpret_cast<uint4*>(o);
  }
}
// zero-padded bf16 copy of the prompts of one pass: [tile_n][d]
__global__ void pad_prompts_bf16_kernel(const float* __restrict__ q, int p0, int pb, int d, int tile_n, __nv_bfloat16* __restrict__ out) {
  const int n = tile_n * d;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int p = i / d;
    out[i] = __float2bfloat16_rn(p < pb ? q[(size_t)(p0 + p) * d + (i - p * d)] : 0.f);
  }
}

// exact fp32 re-scoring of the candidates: one warp per (prompt, voxel) entry of the flat list; the key
// (ordered(score) << 32 | ~rank) goes to the prompt's own list
__global__ void __launch_bounds__(256) rescore_kernel(const float* __restrict__ vsum, const uint32_t* __restrict__ vcount,
                                                      const uint32_t* __restrict__ rank_of_id, const float* __restrict__ q,
                                                      int d, int normalize, const unsigned long long* __restrict__ pairs,
                                                      const uint32_t* __restrict__ pair_cnt, uint32_t pair_cap,
                                                      uint32_t* __restrict__ cand_cnt, uint32_t cap,
                                                      unsigned long long* __restrict__ keys) {
  const int lane = lane_id();
  const uint32_t n = min(*pair_cnt, pair_cap);
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += n_warps) {
    const unsigned long long e = pairs[i];
    const uint32_t id = (uint32_t)e, p = (uint32_t)(e >> 32);
    const float* qp = q + (size_t)p * d;
    float dot = 0.f, ss = 0.f;
    for (int c = lane; c < d / 4; c += 32) {
      const float4 v = *reinterpret_cast<const float4*>(vsum + (size_t)id * d + 4 * c);
      const float4 w = *reinterpret_cast<const float4*>(qp + 4 * c);
      ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
      dot = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, dot))));
    }
    for (int o = 16; o > 0; o >>= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
      dot += __shfl_xor_sync(0xffffffffu, dot, o);
    }
    if (lane == 0) {
      const float cnt = (float)vcount[id];
      float inv = 1.0f;
      if (normalize) inv = __fdiv_rn(1.0f, fmaxf(__fdiv_rn(sqrtf(ss), cnt), 1e-12f));
      const float sc = __fmul_rn(__fdiv_rn(dot, cnt), inv);
      const uint32_t pos = atomicAdd(&cand_cnt[p], 1u);
      if (pos < cap)
        keys[(size_t)p * cap + pos] =
            ((unsigned long long)float_to_ordered(sc) << 32) | (unsigned long long)(0xFFFFFFFFu - rank_of_id[id]);
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// rows x cols fp32 matrix whose rows start row_stride_elems floats apart (a strided sample of the voxel rows when
// row_stride_elems > cols)
static int make_map_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                       uint64_t row_stride_elems = 0, bool bf16 = false) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || !p) {
      cudaGetLastError();
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return VSM_E_CUDA;
    }
    fn = (EncodeTiledFn)p;
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {(row_stride_elems ? row_stride_elems : cols) * (bf16 ? 2 : 4)};
  const cuuint32_t box[2] = {(cuuint32_t)(bf16 ? 2 * kChunkK : kChunkK), box_rows};  // 128 bytes wide either way
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with %d", (int)r);
    return VSM_E_CUDA;
  }
  return VSM_OK;
}

}  // namespace tc

namespace tc {

struct Scratch {
  bool bf;                    // tensor-core passes read the bf16 shadow (engine 3) instead of the fp32 sums
  const __nv_bfloat16* shadow;
  float *qnorm, *thr, *qpad, *ssc;
  int64_t* sidx;
  uint32_t* cand_cnt;         // [P] keys per prompt, then [1] entries of the flat list
  unsigned long long* pairs;  // [P * cap] flat (prompt, voxel) candidates of the tensor-core pass
  unsigned long long* keys;   // [P][cap] exact keys per prompt
  uint32_t cap;
};

// One level: tensor-core pass over the voxel rows 0, stride, 2*stride, ... (n_rows of them) with the thresholds in
// sc.thr, exact re-scoring of the candidates, top-k.  fell_back != null (the level that answers the query): the
// candidate counts are read back (the level's one synchronisation) and *fell_back is set, nothing written, when a
// list overflowed or came up short.  fell_back == null (a level that only produces thresholds): nothing is read back
// -- whatever subset of candidates survived an overflow, the k-th best of their EXACT scores is still a valid lower
// bound, and a list shorter than k yields NaN = "keep everything" downstream.
static int tc_level(vsm_map* m, const float* q_dev, int P, int k, int normalize, uint32_t stride, uint32_t n_rows,
                    const Scratch& sc, int tile_n, size_t smem, int64_t* idx_dev, float* score_dev, bool* fell_back,
                    cudaStream_t s) {
  const int d = m->d;
  const bool bf = sc.bf;
  const int n_kchunks = bf ? d / (2 * kChunkK) : d / kChunkK;
  if (fell_back) *fell_back = false;
  VSM_CUDA(cudaMemsetAsync(sc.cand_cnt, 0, (size_t)(P + 1) * 4, s));
  uint32_t* pair_cnt = sc.cand_cnt + P;
  const uint32_t pair_cap = (uint32_t)std::min<uint64_t>((uint64_t)P * sc.cap, 0xFFFFFFFFull);
  CUtensorMap map_a, map_b;
  static const bool use_pair = !(getenv("VSM_TC_PAIR") && getenv("VSM_TC_PAIR")[0] == '0');
  // 256 prompts: CTA pairs (cta_group::2).  At 128 prompts the pair kernel measured slower than the single-CTA one
  // (4.71 vs 4.61 ms per call at 10 M voxels): that pass is closer to the HBM roof than to the tensor roof.
  const bool pair = tile_n == 256 && use_pair;
  const uint32_t tile_rows = pair ? 2 * kTileM
                                  : (tile_n == 64 ? Cfg<64>::kTileRows : (tile_n == 128 ? Cfg<128>::kTileRows : Cfg<256>::kTileRows));
  VSM_TRY(make_map_2d(&map_a, bf ? (const void*)sc.shadow : (const void*)m->vsum.as<float>(), (uint64_t)n_rows, (uint64_t)d,
                      pair ? kTileM : tile_rows, (uint64_t)stride * d, bf));
  VSM_TRY(make_map_2d(&map_b, sc.qpad, (uint64_t)tile_n, (uint64_t)d, (uint32_t)(pair ? tile_n / 2 : tile_n), 0, bf));
  const int n_sm = sm_count();
  const uint32_t n_tiles = (n_rows + tile_rows - 1) / tile_rows;
  const int grid = pair ? 2 * (int)std::min<uint32_t>(n_tiles, (uint32_t)n_sm / 2) : (int)std::min<uint32_t>(n_tiles, (uint32_t)n_sm);
  const size_t pair_smem = PairCfg<256>::kSmemBytes;
  if (pair) {
    VSM_CUDA(cudaFuncSetAttribute(query_tc_pair_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair_smem));
    VSM_CUDA(cudaFuncSetAttribute(query_tc_pair_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair_smem));
  }
  for (int p0 = 0; p0 < P; p0 += tile_n) {
    const int pb = std::min(tile_n, P - p0);
    if (bf)
      pad_prompts_bf16_kernel<<<64, 256, 0, s>>>(q_dev, p0, pb, d, tile_n, reinterpret_cast<__nv_bfloat16*>(sc.qpad));
    else
      pad_prompts_kernel<<<64, 256, 0, s>>>(q_dev, p0, pb, d, tile_n, sc.qpad);
    VSM_LAUNCHED();
    TcArgs a;
    a.vcount = m->vcount.as<uint32_t>();
    a.vnorm = m->q_norm.as<float>();
    a.thr = sc.thr + p0;
    a.qnorm = sc.qnorm + p0;
    a.n_rows = n_rows;
    a.row_stride = stride;
    a.pb = pb;
    a.p0 = p0;
    a.normalize = normalize;
    a.pairs = sc.pairs;
    a.pair_cnt = pair_cnt;
    a.pair_cap = pair_cap;
    a.n_kchunks = n_kchunks;
    a.margin = bf ? kBf16Margin : kTf32Margin;
    if (tile_n == 64 && bf)
      query_tc_kernel<64, true><<<grid, Cfg<64>::kThreads, smem, s>>>(map_a, map_b, a);
    else if (tile_n == 64)
      query_tc_kernel<64, false><<<grid, Cfg<64>::kThreads, smem, s>>>(map_a, map_b, a);
    else if (!pair && tile_n == 128 && bf)
      query_tc_kernel<128, true><<<grid, Cfg<128>::kThreads, smem, s>>>(map_a, map_b, a);
    else if (!pair && tile_n == 128)
      query_tc_kernel<128, false><<<grid, Cfg<128>::kThreads, smem, s>>>(map_a, map_b, a);
    else if (!pair && bf)
      query_tc_kernel<256, true><<<grid, Cfg<256>::kThreads, smem, s>>>(map_a, map_b, a);
    else if (!pair)
      query_tc_kernel<256, false><<<grid, Cfg<256>::kThreads, smem, s>>>(map_a, map_b, a);
    else {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)grid);
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = pair_smem;
      cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (bf)
        VSM_CUDA(cudaLaunchKernelEx(&cfg, query_tc_pair_kernel<256, true>, map_a, map_b, a));
      else
        VSM_CUDA(cudaLaunchKernelEx(&cfg, query_tc_pair_kernel<256, false>, map_a, map_b, a));
    }
    VSM_LAUNCHED();
  }
  // the list's length lives on the device: a fixed grid strides over it
  rescore_kernel<<<sm_count() * 8, 256, 0, s>>>(m->vsum.as<float>(), m->vcount.as<uint32_t>(), m->rank_of_id.as<uint32_t>(), q_dev, d,
                                         normalize, sc.pairs, pair_cnt, pair_cap, sc.cand_cnt, sc.cap, sc.keys);
  VSM_LAUNCHED();
  if (fell_back) {
    std::vector<uint32_t> h_cnt(P + 1);
    VSM_TRY(read_back(m, h_cnt.data(), sc.cand_cnt, (size_t)(P + 1) * 4, s));  // the one synchronisation of a query
    uint32_t mx = 0, mn = 0xFFFFFFFFu;
    for (int p = 0; p < P; ++p) {
      mx = std::max(mx, h_cnt[p]);
      mn = std::min(mn, h_cnt[p]);
    }
    m->tc_last_candidates = mx;
    if (h_cnt[P] > pair_cap || mx > sc.cap || mn < (uint32_t)k) {
      *fell_back = true;  // a list overflowed (dense ties) -- or, impossibly, lost candidates
      return VSM_OK;
    }
  }
  return query_select_from_keys(m, sc.keys, sc.cand_cnt, P, (int)sc.cap, k, idx_dev, score_dev, s);
}

}  // namespace tc

int query_tc(vsm_map* m, const float* q_dev, int P, int k, int normalize, int64_t* idx_dev, float* score_dev,
             cudaStream_t s, bool bf) {
  using namespace tc;
  const uint32_t V = (uint32_t)m->n_vox;
  const int d = m->d;
  if (d % (bf ? 2 * kChunkK : kChunkK) != 0 || d > 1024) {
    set_error("vsm_query engine %d needs d to be a multiple of %d and <= 1024 (d=%d)", bf ? 3 : 2, bf ? 2 * kChunkK : kChunkK, d);
    return VSM_E_INVALID;
  }
  const int n_kchunks = bf ? d / (2 * kChunkK) : d / kChunkK;
  // prompts per pass: 64 with the prompt block resident in shared memory, 128 / 256 with its slices streamed
  const int tile_n = P <= 64 ? 64 : (P <= 128 ? 128 : 256);
  const size_t smem = tile_n == 64    ? (size_t)n_kchunks * Cfg<64>::kBChunkBytes + (size_t)Cfg<64>::kStages * Cfg<64>::kStageBytes + 256
                      : tile_n == 128 ? (size_t)Cfg<128>::kStages * Cfg<128>::kStageBytes + 256
                                      : (size_t)Cfg<256>::kStages * Cfg<256>::kStageBytes + 256;
  if (smem > 227 * 1024) {
    set_error("vsm_query engine 2: d=%d needs %zu bytes of shared memory", d, smem);
    return VSM_E_INVALID;
  }
  if (tile_n == 64 && bf)
    VSM_CUDA(cudaFuncSetAttribute(query_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (tile_n == 64)
    VSM_CUDA(cudaFuncSetAttribute(query_tc_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (tile_n == 128 && bf)
    VSM_CUDA(cudaFuncSetAttribute(query_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (tile_n == 128)
    VSM_CUDA(cudaFuncSetAttribute(query_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (bf)
    VSM_CUDA(cudaFuncSetAttribute(query_tc_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else
    VSM_CUDA(cudaFuncSetAttribute(query_tc_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // ---- scratch: norms (cached per finalisation), thresholds, candidates ------------------------------------
  Scratch sc;
  sc.cap = 1u << 16;
  sc.bf = bf;
  sc.shadow = nullptr;
  VSM_TRY(m->q_norm.ensure((size_t)std::max<uint32_t>(V, 1) * 4, s));
  if (!m->norms_valid) {
    row_norm_kernel<<<sm_count() * 8, 256, 0, s>>>(m->vsum.as<float>(), V, d, m->q_norm.as<float>());
    VSM_LAUNCHED();
    m->norms_valid = true;
    m->shadow_valid = false;
  }
  if (bf) {
    // the bf16 shadow of the sums: built once per finalisation (norms_valid is reset whenever the sums change), 2 bytes
    // per value beside the exact fp32 store -- the tensor-core passes then read half the bytes
    VSM_TRY(m->q_shadow.ensure((size_t)std::max<uint32_t>(V, 1) * d * 2, s));
    if (!m->shadow_valid) {
      shadow_kernel<<<sm_count() * 16, 256, 0, s>>>(m->vsum.as<float>(), (size_t)V * d / 8, m->q_shadow.as<__nv_bfloat16>());
      VSM_LAUNCHED();
      m->shadow_valid = true;
    }
    sc.shadow = m->q_shadow.as<__nv_bfloat16>();
  }
  // layout of q_tc: [qnorm P][thr P][padded prompts kMaxTileN*d][sample idx P*k (i64)][sample scores P*k][cand_cnt P + 1]
  const size_t off_qn = 0, off_thr = off_qn + (size_t)P * 4, off_pad = (off_thr + (size_t)P * 4 + 255) & ~(size_t)255;
  const size_t off_sidx = (off_pad + (size_t)kMaxTileN * d * 4 + 255) & ~(size_t)255;
  const size_t off_ssc = off_sidx + (size_t)P * k * 8, off_cnt = (off_ssc + (size_t)P * k * 4 + 255) & ~(size_t)255;
  const size_t small_bytes = off_cnt + (size_t)(P + 1) * 4;
  VSM_TRY(m->q_tc.ensure(small_bytes, s));
  VSM_TRY(m->q_tc_cand.ensure((size_t)P * sc.cap * 16, s));  // flat (prompt, voxel) list + keys per prompt
  uint8_t* base = m->q_tc.as<uint8_t>();
  sc.qnorm = (float*)(base + off_qn);
  sc.thr = (float*)(base + off_thr);
  sc.qpad = (float*)(base + off_pad);
  sc.sidx = (int64_t*)(base + off_sidx);
  sc.ssc = (float*)(base + off_ssc);
  sc.cand_cnt = (uint32_t*)(base + off_cnt);
  sc.pairs = m->q_tc_cand.as<unsigned long long>();
  sc.keys = sc.pairs + (size_t)P * sc.cap;

  // ---- 1. thresholds: the k-th best EXACT score of a subset of the voxels bounds the k-th best overall from below.
  // Nested strided subsets, each scored by a tensor-core level of its own (the exact engine would need P/8 passes
  // per subset: 5.6 ms of a 16 ms call at P = 256, measured):  C (ids 0, 32*stride, ...; a few thousand rows) with
  // threshold -inf, i.e. all of C re-scored exactly;  B (ids 0, stride, ...; >= 65536 rows) with C's thresholds;
  // then the pass over everything with B's.
  const uint32_t target_b = std::max<uint32_t>(std::min<uint32_t>(65536u, V / 4), V / 256);
  const uint32_t stride_b = std::max<uint32_t>(1, V / std::max<uint32_t>(target_b, 1u));
  const uint32_t n_b = (V + stride_b - 1) / stride_b;
  const uint32_t stride_c = stride_b * 32;
  const uint32_t n_c = (V + stride_c - 1) / stride_c;
  int have_sample = 0;
  if (stride_b > 1 && n_c >= (uint32_t)std::max(4 * k, 1024) && n_c <= sc.cap) {
    // level C: every row of C is a candidate (threshold -inf), scored exactly by the re-scoring kernel -- one
    // tensor-core launch instead of P/8 passes of the exact engine
    prompt_prep_kernel<<<(P + 63) / 64, 64, 0, s>>>(q_dev, P, d, sc.ssc, k, 0, sc.qnorm, sc.thr);
    VSM_LAUNCHED();
    VSM_TRY(tc_level(m, q_dev, P, k, normalize, stride_c, n_c, sc, tile_n, smem, sc.sidx, sc.ssc, nullptr, s));
    prompt_prep_kernel<<<(P + 63) / 64, 64, 0, s>>>(q_dev, P, d, sc.ssc, k, 1, sc.qnorm, sc.thr);
    VSM_LAUNCHED();
    VSM_TRY(tc_level(m, q_dev, P, k, normalize, stride_b, n_b, sc, tile_n, smem, sc.sidx, sc.ssc, nullptr, s));
    have_sample = 1;
  } else if (n_b >= (uint32_t)k && stride_b > 1) {
    VSM_TRY(query_exact_rows(m, q_dev, P, k, normalize, stride_b, n_b, sc.sidx, sc.ssc, s));
    have_sample = 1;
  }
  prompt_prep_kernel<<<(P + 63) / 64, 64, 0, s>>>(q_dev, P, d, sc.ssc, k, have_sample, sc.qnorm, sc.thr);
  VSM_LAUNCHED();

  // ---- 2. + 3. tensor-core pass over all voxels, exact re-scoring + top-k ---------------------------------------
  bool fell_back = false;
  VSM_TRY(tc_level(m, q_dev, P, k, normalize, 1u, V, sc, tile_n, smem, idx_dev, score_dev, &fell_back, s));
  if (fell_back) {
    m->tc_fallbacks += 1;
    return query_exact(m, q_dev, P, k, normalize, idx_dev, score_dev, s);
  }
  return VSM_OK;
}

}  // namespace vsm

extern "C" int vsm_query_shadow_release(vsm_map* m) {
  if (!m) {
    vsm::set_error("null map");
    return VSM_E_INVALID;
  }
  cudaSetDevice(m->device);
  cudaDeviceSynchronize();
  m->q_shadow.release();
  m->shadow_valid = false;
  return VSM_OK;
}
