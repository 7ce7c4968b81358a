// K4 — per-partition feature-histogram build (sm_100a).
//
// Replaces [UPSTREAM lightgbmlib 3.2.110] Dataset::ConstructHistograms /
// DenseBin<uint8>::ConstructHistogram, reached from the reference at
// lightgbm/src/main/scala/com/microsoft/ml/spark/lightgbm/booster/LightGBMBooster.scala:351-361
// (LGBM_BoosterUpdateOneIter).  SURVEY.md §8(a) row a4.
//
// Layout in HBM
//   bins : uint8 [num_tiles][rows_stride][32]   ("tile-major": a feature tile = 32 features,
//          one row of a tile = one 32-byte sector, so both the streamed root pass and the
//          index-list gather of a leaf move whole sectors)
//   qgh  : int4  [N]  per-row fixed-point gradient/hessian words {g_hi, g_lo, h_hi, h_lo}
//          (see quantize.h).  Written once per tree by the gradient kernel.
//   idx  : int32 row-index list of the data partition (leaves are contiguous ranges)
//   hist : int64 [num_feat_padded][256][2]  (g, h) fixed-point sums per leaf slot.
//
// Why fixed point: shared memory on sm_100a has exactly one native atomic add,
// ATOMS.ADD (32-bit integer).  atomicAdd on float/double/u64 in shared memory compiles to an
// ATOMS.CAST.SPIN compare-and-swap loop (checked with cuobjdump).  The reference accumulates
// fp32 gradients into fp64 bins, so fp32 accumulation is not good enough to reproduce its
// tree structure.  We therefore split a 36-bit fixed-point value into two 18-bit fields and
// accumulate each with a native 32-bit atomic; 2^14 rows can be added before a field can
// overflow, then the CTA flushes its sub-histogram into the int64 leaf histogram in L2 with
// RED.ADD.64.  Integer sums are exact and order-independent, so the result is bit-reproducible
// run to run and across ranks (the NCCL reduction is an int64 sum).
//
// Bank mapping: the sub-histogram planes are laid out [bin][feature-of-tile], so feature f lives in
// bank f.  Every ATOMS instruction of a warp addresses 32 DIFFERENT features, so it is conflict-free by
// construction regardless of the bin distribution (ncu: 1.0 wavefront per ATOMS).
//   k4_hist_build    (v2): warp step = 4 rows x 32 features, lane = (row selector, bin word), bytes rotated by the row selector.
//   k4_hist_build_ws (v3, the engine's kernel): warp step = 32 rows x 4 features, lane = row for 8 steps; the lane's (g,h)
//     quadruple stays in registers (one 16-byte LDS per 32 cells instead of one per 4), bin word and byte rotated by the lane id.
// The LSU data pipe is the binding unit (ncu: ~94 % busy): an ATOMS wavefront costs ~0.93 cycles, so the v3 loop spends
// 16 x 0.93 (atomics) + 1 (bin word) + 0.5 (q) cycles per 128 cells.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200gbm {

constexpr int kTileFeat = 32;                      // features per tile == lanes per warp
constexpr int kBins = 256;                         // uint8 bin ids
constexpr int kLoBits = 18;                        // low fixed-point field
constexpr int kFlushRows = 1 << (32 - kLoBits);    // rows a sub-histogram may absorb (16384)
#ifndef B200GBM_K4_EXPERIMENT
#define B200GBM_K4_EXPERIMENT 0
#endif
#ifndef B200GBM_STAGE_ROWS
#define B200GBM_STAGE_ROWS 512
#endif
#ifndef B200GBM_STAGES
#define B200GBM_STAGES 3
#endif
#ifndef B200GBM_HIST_THREADS
#define B200GBM_HIST_THREADS 512
#endif
constexpr int kStageRows = B200GBM_STAGE_ROWS;     // rows per staged sub-chunk
constexpr int kStages = B200GBM_STAGES;            // cp.async ring depth
constexpr int kHistThreads = B200GBM_HIST_THREADS;
constexpr int kStageSlots = (kStageRows * 2 + kHistThreads - 1) / kHistThreads;   // half-rows a thread stages per stage
constexpr int kHistWarps = kHistThreads / 32;
constexpr int kPlaneWords = kBins * kTileFeat;     // 8192 words per plane
constexpr int kHistSmemBytes =
    4 * kPlaneWords * 4 + kStages * (kStageRows * 32 + kStageRows * 16);

// Device-resident work descriptor: the controller kernels write it, so the host never has to
// know leaf sizes (no host sync inside a tree).
struct HistWork {
  int begin;      // first position in idx (or first row when use_idx == 0)
  int count;      // rows of the leaf on this rank
  int use_idx;    // 0: rows are begin..begin+count-1 directly (root of a full pass)
  int buf;        // which of the two index buffers holds the leaf's row list
};

// (g,h) words of a leaf in partition order: qord[begin + i] = qgh[idx[begin + i]] (no-op for the identity root)
__global__ void __launch_bounds__(256)
k_gather_q(const HistWork* __restrict__ work, const int* __restrict__ idx0, const int* __restrict__ idx1, const int4* __restrict__ qgh,
           int4* __restrict__ qord) {
  const HistWork w = *work;
  if (!w.use_idx) return;
  const int* __restrict__ idx = w.buf ? idx1 : idx0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w.count; i += gridDim.x * blockDim.x) {
    const int p = w.begin + i;
    qord[p] = qgh[idx[p]];
  }
}


__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  int sz = valid ? 16 : 0;   // src-size 0 => zero fill, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// NATOM selects how many planes are accumulated: 4 = (g_hi,g_lo,h_hi,h_lo) general case,
// 3 = constant-hessian objectives (g_hi,g_lo,count) [UPSTREAM is_constant_hessian path].
template <int NATOM>
__global__ void __launch_bounds__(kHistThreads, 1)
k4_hist_build(const uint8_t* __restrict__ bins, size_t rows_stride, int num_tiles,
              const int4* __restrict__ qgh, const int* __restrict__ idx0, const int* __restrict__ idx1,
              const HistWork* __restrict__ work, unsigned long long* __restrict__ hist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned* plane = reinterpret_cast<unsigned*>(smem_raw);              // [4][kPlaneWords]
  unsigned char* stage_bins = smem_raw + 4 * kPlaneWords * 4;           // [kStages][256][32]
  int4* stage_q = reinterpret_cast<int4*>(stage_bins + kStages * kStageRows * 32);  // [kStages][256]

  const HistWork w = *work;
  const int n = w.count;
  if (n <= 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int* __restrict__ idx = w.buf ? idx1 : idx0;

  // rows per work item: large enough to amortise the flush, small enough to fill the grid
  long long cells_rows = static_cast<long long>(n) * num_tiles;
  int rpi = static_cast<int>((cells_rows + gridDim.x - 1) / gridDim.x);
  rpi = max(rpi, 2048);
  rpi = min(rpi, kFlushRows);
  rpi = (rpi + kStageRows - 1) / kStageRows * kStageRows;
  const int chunks = (n + rpi - 1) / rpi;
  const int items = chunks * num_tiles;

  for (int e = tid; e < 4 * kPlaneWords; e += kHistThreads) plane[e] = 0u;
  __syncthreads();

  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int tile = item % num_tiles;
    const int chunk = item / num_tiles;
    const int row0 = chunk * rpi;                       // position inside the leaf
    const int nrows = min(rpi, n - row0);
    const int nst = (nrows + kStageRows - 1) / kStageRows;
    const uint8_t* tbins = bins + static_cast<size_t>(tile) * rows_stride * 32;

    // a "slot" = half a row of bins (16 B); the thread that stages half 0 also stages the row's qgh word
    auto row_of = [&](int st, int slot) -> int {
      int hrow = slot * kHistThreads + tid;
      if (hrow >= kStageRows * 2) return -1;
      int p = row0 + st * kStageRows + (hrow >> 1);
      if (p >= row0 + nrows) return -1;
      return w.use_idx ? idx[w.begin + p] : (w.begin + p);
    };
    auto issue = [&](int st, int slot, int r) {
      int hrow = slot * kHistThreads + tid;
      if (hrow >= kStageRows * 2) return;
      int srow = hrow >> 1, shalf = hrow & 1;
      int buf = st % kStages;
      bool ok = r >= 0;
      size_t rr = ok ? static_cast<size_t>(r) : 0;
      cp_async16(stage_bins + (buf * kStageRows + srow) * 32 + shalf * 16,
                 tbins + rr * 32 + shalf * 16, ok);
      if (shalf == 0) cp_async16(stage_q + buf * kStageRows + srow, qgh + rr, ok);
    };

    int rnext[kStageSlots];
#pragma unroll
    for (int sl = 0; sl < kStageSlots; ++sl) rnext[sl] = row_of(0, sl);
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
      if (s < nst) {
#pragma unroll
        for (int sl = 0; sl < kStageSlots; ++sl) {
          int r = rnext[sl];
          rnext[sl] = (s + 1 < nst) ? row_of(s + 1, sl) : -1;
          issue(s, sl, r);
        }
      }
      cp_async_commit();
    }
    for (int s = 0; s < nst; ++s) {
      cp_async_wait<kStages - 2>();
      __syncthreads();
      int sn = s + kStages - 1;
      if (sn < nst) {
#pragma unroll
        for (int sl = 0; sl < kStageSlots; ++sl) {
          int r = rnext[sl];
          rnext[sl] = (sn + 1 < nst) ? row_of(sn + 1, sl) : -1;
          issue(sn, sl, r);
        }
      }
      cp_async_commit();

      const int buf = s % kStages;
      const unsigned* sw = reinterpret_cast<const unsigned*>(stage_bins + buf * kStageRows * 32);   // 8 words per row
      const int4* sq = stage_q + buf * kStageRows;
      // One warp step = 4 rows x 32 features.  Lane = (rsel = lane>>3, word = lane&7) reads the 4 bins of features
      // 4*word..4*word+3 of row r0+rsel with ONE LDS.32 (the warp reads 128 contiguous bytes) and that row's 4
      // fixed-point words with one LDS.128.  In sub-step k it handles feature 4*word + ((k+rsel)&3): for fixed k the map
      // (rsel, word) -> feature is a bijection onto 0..31, and plane[.][bin*32+feature] puts feature f in bank f, so all
      // 32 lanes hit distinct banks for ANY bin values: 16 conflict-free ATOMS per 128 cells, loads amortised 4x.
      const int rsel = lane >> 3, wsel = lane & 7;
#pragma unroll 2
      for (int k4 = 0; k4 < kStageRows / (4 * kHistWarps); ++k4) {
        const int r = (warp + k4 * kHistWarps) * 4 + rsel;
        const unsigned word = sw[r * 8 + wsel];
        const int4 q = sq[r];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int kk = (k + rsel) & 3;
          const unsigned b = (word >> (8 * kk)) & 0xFFu;
          const unsigned a = b * 32u + static_cast<unsigned>(wsel * 4 + kk);
          atomicAdd(&plane[a], static_cast<unsigned>(q.x));
          atomicAdd(&plane[kPlaneWords + a], static_cast<unsigned>(q.y));
          atomicAdd(&plane[2 * kPlaneWords + a], static_cast<unsigned>(q.z));
          if (NATOM == 4) atomicAdd(&plane[3 * kPlaneWords + a], static_cast<unsigned>(q.w));
        }
      }
    }
    cp_async_wait<0>();
    __syncthreads();
    // flush the sub-histogram into the leaf histogram (int64, L2-resident) and re-zero it
    for (int e = tid; e < kPlaneWords; e += kHistThreads) {
      unsigned ghi = plane[e], glo = plane[kPlaneWords + e];
      unsigned hhi = plane[2 * kPlaneWords + e];
      unsigned hlo = (NATOM == 4) ? plane[3 * kPlaneWords + e] : 0u;
      if (ghi | glo | hhi | hlo) {
        int f = tile * 32 + (e & 31);
        int b = e >> 5;
        long long g = (static_cast<long long>(static_cast<int>(ghi)) << kLoBits) +
                      static_cast<long long>(glo);
        long long h;
        if (NATOM == 4)
          h = (static_cast<long long>(static_cast<int>(hhi)) << kLoBits) +
              static_cast<long long>(hlo);
        else
          h = static_cast<long long>(hhi);   // plain row count
        size_t o = (static_cast<size_t>(f) * kBins + b) * 2;
        if (g) atomicAdd(&hist[o], static_cast<unsigned long long>(g));
        if (h) atomicAdd(&hist[o + 1], static_cast<unsigned long long>(h));
        plane[e] = 0u;
        plane[kPlaneWords + e] = 0u;
        plane[2 * kPlaneWords + e] = 0u;
        if (NATOM == 4) plane[3 * kPlaneWords + e] = 0u;
      }
    }
    __syncthreads();
  }
}


// ===================================================================================================
// K4 v3 — warp-specialised variant: one PRODUCER warp stages rows into a ring of shared-memory stages,
// 16 CONSUMER warps do nothing but the conflict-free scatter.  Stages are handed over with mbarriers
// (full / empty), so there is no block-wide barrier per stage and the producer runs ahead across work
// items (it prefetches the next item while the consumers flush the current sub-histogram).
//   * contiguous row ranges (root pass, use_idx == 0): TMA bulk copies — one elected lane issues
//     cp.async.bulk (UBLKCP) for the stage's bins (rows*32 B) and qgh (rows*16 B), completion is signalled
//     by the mbarrier's transaction count;
//   * index-list leaves: the 32 producer lanes issue 16-byte cp.async gathers (LDGSTS) and arrive on the
//     mbarrier with cp.async.mbarrier.arrive.noinc when their copies have landed.
constexpr int kWsConsumerWarps = 16;
#ifndef B200GBM_WS_PRODUCERS
#define B200GBM_WS_PRODUCERS 2
#endif
constexpr int kWsProducerWarps = B200GBM_WS_PRODUCERS;            // 1 lane issues the TMA bulk copies; all lanes share the gathers
constexpr int kWsThreads = (kWsConsumerWarps + kWsProducerWarps) * 32;
constexpr int kWsStageRows = 512;
constexpr int kWsStages = 3;
constexpr int kWsSmemBytes = 4 * kPlaneWords * 4 + kWsStages * (kWsStageRows * 32 + kWsStageRows * 16) + 64;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

template <int NATOM>
__global__ void __launch_bounds__(kWsThreads, 1)
k4_hist_build_ws(const uint8_t* __restrict__ bins, size_t rows_stride, int num_tiles, const int4* __restrict__ qgh,
                 const int4* __restrict__ qord, const int* __restrict__ idx0, const int* __restrict__ idx1,
                 const HistWork* __restrict__ work, unsigned long long* __restrict__ hist) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned* plane = reinterpret_cast<unsigned*>(smem_raw);
  unsigned char* stage_bins = smem_raw + 4 * kPlaneWords * 4;
  int4* stage_q = reinterpret_cast<int4*>(stage_bins + kWsStages * kWsStageRows * 32);
  unsigned long long* full_bar = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(stage_q) + kWsStages * kWsStageRows * 16);
  unsigned long long* empty_bar = full_bar + kWsStages;

  const HistWork w = *work;
  const int n = w.count;
  if (n <= 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int* __restrict__ idx = w.buf ? idx1 : idx0;

  // Work items are (tile, row chunk) pairs in TILE-MAJOR order; every CTA takes one contiguous range of them, so
  // consecutive items of a CTA mostly belong to the same feature tile and the sub-histogram is flushed only when the
  // tile changes or the 2^14-row field headroom is used up (small and medium leaves: one flush per CTA).
  long long cells_rows = static_cast<long long>(n) * num_tiles;
  int rpi = static_cast<int>((cells_rows + 4LL * gridDim.x - 1) / (4LL * gridDim.x));
  rpi = (rpi + kWsStageRows - 1) / kWsStageRows * kWsStageRows;
  rpi = max(rpi, kWsStageRows);
  rpi = min(rpi, kFlushRows);
  const int chunks = (n + rpi - 1) / rpi;
  const long long items = static_cast<long long>(chunks) * num_tiles;
  const int i0 = static_cast<int>(items * blockIdx.x / gridDim.x);
  const int i1 = static_cast<int>(items * (blockIdx.x + 1) / gridDim.x);
  if (i0 >= i1) return;

  if (tid == 0) {
    for (int i = 0; i < kWsStages; ++i) { mbar_init(&full_bar[i], kWsProducerWarps * 32); mbar_init(&empty_bar[i], kWsConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int e = tid; e < 4 * kPlaneWords; e += kWsThreads) plane[e] = 0u;
  __syncthreads();

  unsigned gs = 0;   // global stage counter (same sequence on both sides)
  if (warp >= kWsConsumerWarps) {
    // ------------------------------------------------------------------ producer warps
    const int ptid = tid - kWsConsumerWarps * 32;                 // 0 .. kWsProducerWarps*32-1
    constexpr int kPT = kWsProducerWarps * 32;
    constexpr int kPerLane = kWsStageRows / kPT;
    for (int item = i0; item < i1; ++item) {
      const int tile = item / chunks, chunk = item - tile * chunks;
      const int row0 = chunk * rpi;
      const int nrows = min(rpi, n - row0);
      const int nst = (nrows + kWsStageRows - 1) / kWsStageRows;
      const uint8_t* tbins = bins + static_cast<size_t>(tile) * rows_stride * 32;
      for (int s = 0; s < nst; ++s, ++gs) {
        const int slot = gs % kWsStages;
        mbar_wait(&empty_bar[slot], ((gs / kWsStages) & 1u) ^ 1u);
        const int p0 = row0 + s * kWsStageRows;
        const int rows = min(kWsStageRows, row0 + nrows - p0);
        unsigned char* db = stage_bins + slot * kWsStageRows * 32;
        int4* dq = stage_q + slot * kWsStageRows;
        if (!w.use_idx) {
          if (ptid == 0) {
            mbar_arrive_expect_tx(&full_bar[slot], static_cast<unsigned>(rows) * 48u);
            const size_t r = static_cast<size_t>(w.begin + p0);
            tma_bulk_g2s(db, tbins + r * 32, static_cast<unsigned>(rows) * 32u, &full_bar[slot]);
            tma_bulk_g2s(dq, qgh + r, static_cast<unsigned>(rows) * 16u, &full_bar[slot]);
          } else {
            mbar_arrive(&full_bar[slot]);
          }
        } else {
          // leaf = index list: the bin rows are gathered sector by sector with cp.async; the (g,h) words were put into LEAF ORDER
          // once per leaf by k_gather_q (qord[position]), so every feature tile streams them with one bulk copy instead of
          // re-gathering 16 B out of a 64 B DRAM atom 16 times (ncu: the per-tile q gather doubled the DRAM traffic of a leaf pass)
          const int* ip = idx + w.begin + p0;
          if (ptid == 0) {
            mbar_expect_tx(&full_bar[slot], static_cast<unsigned>(rows) * 16u);
            tma_bulk_g2s(dq, qord + static_cast<size_t>(w.begin + p0), static_cast<unsigned>(rows) * 16u, &full_bar[slot]);
          }
          // (one 32-byte cp.async.bulk per row was tried instead: the TMA unit retires ~1 such copy per 32 cycles per SM, 4.5x slower)
          // a lane pair fetches the two 16-byte halves of ONE row sector, so an LDGSTS instruction touches 16 sectors instead of
          // 32 (ncu: the data return costs one LSU wavefront per distinct sector: 28 per instruction with one half-row per lane)
          int rr[2 * kPerLane];
#pragma unroll
          for (int k = 0; k < 2 * kPerLane; ++k) { const int j = (ptid + k * kPT) >> 1; rr[k] = j < rows ? ip[j] : -1; }
#pragma unroll
          for (int k = 0; k < 2 * kPerLane; ++k) {
            const int hh = ptid + k * kPT, j = hh >> 1, half = (hh & 1) * 16;
            if (rr[k] >= 0) cp_async16(db + j * 32 + half, tbins + static_cast<size_t>(rr[k]) * 32 + half, true);
          }
          cp_async_mbar_arrive_noinc(&full_bar[slot]);
        }
      }
    }
    cp_async_wait<0>();
  } else {
    // ------------------------------------------------------------------ consumer warps
    // Warp step = 32 rows x 4 features: lane owns ROW g*32+lane for 8 steps and keeps its (g,h) quadruple in registers, so the
    // 16-byte q load (4 LSU wavefronts for a warp) is paid once per 1024 cells instead of once per 128.  At step k the lane reads
    // bin word w = ((lane>>2)+k)&7 of its row: bank = 8*(lane&3)+w, all 32 distinct.  Inside the word the byte order is rotated
    // by lane&3, so the 32 atomics of one instruction hit features 4w+kk = 32 distinct banks (plane index = bin*32 + feature).
#if B200GBM_K4_EXPERIMENT != 0
    unsigned exp_sink = 0;
#endif
    const int sub = lane >> 2, rot = lane & 3;
    int acc_rows = 0;                                             // rows absorbed by the sub-histogram since the last flush
    for (int item = i0; item < i1; ++item) {
      const int tile = item / chunks, chunk = item - tile * chunks;
      const int row0 = chunk * rpi;
      const int nrows = min(rpi, n - row0);
      const int nst = (nrows + kWsStageRows - 1) / kWsStageRows;
      for (int s = 0; s < nst; ++s, ++gs) {
        const int slot = gs % kWsStages;
        mbar_wait(&full_bar[slot], (gs / kWsStages) & 1u);
        const int rows = min(kWsStageRows, nrows - s * kWsStageRows);
        const unsigned* sw = reinterpret_cast<const unsigned*>(stage_bins + slot * kWsStageRows * 32);
        const int4* sq = stage_q + slot * kWsStageRows;
        const int groups = (rows + 31) >> 5;
        for (int g = warp; g < groups; g += kWsConsumerWarps) {
          const int r = g * 32 + lane;
          if (r < rows) {
            const int4 q = sq[r];
            const unsigned* rw = sw + r * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int w8 = (sub + k) & 7;
              const unsigned word = rw[w8];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int kk = (j + rot) & 3;
                const unsigned b = (word >> (8 * kk)) & 0xFFu;
                const unsigned a = b * 32u + static_cast<unsigned>(w8 * 4 + kk);
#if B200GBM_K4_EXPERIMENT == 0
                atomicAdd(&plane[a], static_cast<unsigned>(q.x));
                atomicAdd(&plane[kPlaneWords + a], static_cast<unsigned>(q.y));
                atomicAdd(&plane[2 * kPlaneWords + a], static_cast<unsigned>(q.z));
                if (NATOM == 4) atomicAdd(&plane[3 * kPlaneWords + a], static_cast<unsigned>(q.w));
#else   // tools/ubench_hist.cu only: cost model of the loop with 0 / 1 / 2 atomics per cell
                if (B200GBM_K4_EXPERIMENT >= 2) atomicAdd(&plane[a], static_cast<unsigned>(q.x));
                if (B200GBM_K4_EXPERIMENT >= 3) atomicAdd(&plane[kPlaneWords + a], static_cast<unsigned>(q.y));
                if (B200GBM_K4_EXPERIMENT == 1) exp_sink ^= a + static_cast<unsigned>(q.x ^ q.y ^ q.z ^ q.w);
#endif
              }
            }
          }
        }
#if B200GBM_K4_EXPERIMENT != 0
        if (exp_sink == 0x12345679u) plane[lane] = exp_sink;
#endif
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[slot]);
      }
      acc_rows += nrows;
      bool flush = (item + 1 == i1);
      if (!flush) {
        const int ntile = (item + 1) / chunks, nchunk = (item + 1) - ntile * chunks;
        const int nnext = min(rpi, n - nchunk * rpi);
        flush = (ntile != tile) || (acc_rows + nnext > kFlushRows);
      }
      if (!flush) continue;
      acc_rows = 0;
      // all consumers finished: flush the sub-histogram (consumer-only named barrier; the producers keep staging)
      asm volatile("bar.sync 1, %0;\n" ::"n"(kWsConsumerWarps * 32) : "memory");
      for (int e = tid; e < kPlaneWords; e += kWsConsumerWarps * 32) {
        unsigned ghi = plane[e], glo = plane[kPlaneWords + e];
        unsigned hhi = plane[2 * kPlaneWords + e];
        unsigned hlo = (NATOM == 4) ? plane[3 * kPlaneWords + e] : 0u;
        if (ghi | glo | hhi | hlo) {
          const int f = tile * 32 + (e & 31);
          const int b = e >> 5;
          long long g = (static_cast<long long>(static_cast<int>(ghi)) << kLoBits) + static_cast<long long>(glo);
          long long h = (NATOM == 4) ? (static_cast<long long>(static_cast<int>(hhi)) << kLoBits) + static_cast<long long>(hlo)
                                     : static_cast<long long>(hhi);
          const size_t o = (static_cast<size_t>(f) * kBins + b) * 2;
          if (g) atomicAdd(&hist[o], static_cast<unsigned long long>(g));
          if (h) atomicAdd(&hist[o + 1], static_cast<unsigned long long>(h));
          plane[e] = 0u;
          plane[kPlaneWords + e] = 0u;
          plane[2 * kPlaneWords + e] = 0u;
          if (NATOM == 4) plane[3 * kPlaneWords + e] = 0u;
        }
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(kWsConsumerWarps * 32) : "memory");
    }
  }
}


}  // namespace b200gbm
