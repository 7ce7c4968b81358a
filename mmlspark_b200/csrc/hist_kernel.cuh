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
// Bank mapping: a warp owns one row at a time, lane l owns feature l of the tile and the
// sub-histogram planes are laid out [bin][lane], so lane l only ever touches bank l: every
// ATOMS instruction is conflict-free by construction regardless of the bin distribution.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200gbm {

constexpr int kTileFeat = 32;                      // features per tile == lanes per warp
constexpr int kBins = 256;                         // uint8 bin ids
constexpr int kLoBits = 18;                        // low fixed-point field
constexpr int kFlushRows = 1 << (32 - kLoBits);    // rows a sub-histogram may absorb (16384)
constexpr int kStageRows = 256;                    // rows per staged sub-chunk
constexpr int kStages = 3;                         // cp.async ring depth
constexpr int kHistThreads = 512;
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

    // thread t stages half a row of bins (16 B); threads < 256 also stage one qgh word
    const int srow = tid >> 1, shalf = tid & 1;
    auto row_of = [&](int st) -> int {
      int p = row0 + st * kStageRows + srow;
      if (p >= row0 + nrows) return -1;
      return w.use_idx ? idx[w.begin + p] : (w.begin + p);
    };
    auto issue = [&](int st, int r) {
      int buf = st % kStages;
      bool ok = r >= 0;
      size_t rr = ok ? static_cast<size_t>(r) : 0;
      cp_async16(stage_bins + (buf * kStageRows + srow) * 32 + shalf * 16,
                 tbins + rr * 32 + shalf * 16, ok);
      if (shalf == 0) cp_async16(stage_q + buf * kStageRows + srow, qgh + rr, ok);
    };

    int rnext = row_of(0);
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
      if (s < nst) {
        int r = rnext;
        rnext = (s + 1 < nst) ? row_of(s + 1) : -1;
        issue(s, r);
      }
      cp_async_commit();
    }
    for (int s = 0; s < nst; ++s) {
      cp_async_wait<kStages - 2>();
      __syncthreads();
      int sn = s + kStages - 1;
      if (sn < nst) {
        int r = rnext;
        rnext = (sn + 1 < nst) ? row_of(sn + 1) : -1;
        issue(sn, r);
      }
      cp_async_commit();

      const int buf = s % kStages;
      const unsigned char* sb = stage_bins + buf * kStageRows * 32;
      const int4* sq = stage_q + buf * kStageRows;
#pragma unroll 4
      for (int k = 0; k < kStageRows / kHistWarps; ++k) {
        int r = warp + k * kHistWarps;
        unsigned b = sb[r * 32 + lane];
        int4 q = sq[r];
        unsigned a = b * 32u + lane;
        atomicAdd(&plane[a], static_cast<unsigned>(q.x));
        atomicAdd(&plane[kPlaneWords + a], static_cast<unsigned>(q.y));
        if (NATOM == 4) {
          atomicAdd(&plane[2 * kPlaneWords + a], static_cast<unsigned>(q.z));
          atomicAdd(&plane[3 * kPlaneWords + a], static_cast<unsigned>(q.w));
        } else {
          atomicAdd(&plane[2 * kPlaneWords + a], static_cast<unsigned>(q.z));
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

}  // namespace b200gbm
