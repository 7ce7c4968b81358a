// REJECTED EXPERIMENT (round 2), kept for the record and for tools/ubench_hist.cu — not part of the product library.
// Result on B200 (profiles/r02_k4v5_direct_load.md): 1.14e12 cells/s against 1.44e12 for the shipped warp-specialised TMA kernel.
#pragma once
#include "../mmlspark_b200/csrc/hist_kernel.cuh"

namespace b200gbm {

// ===================================================================================================
// K4 v5 — direct-load variant: no shared-memory staging at all.  The v3/v4 kernel is bound by the LSU data pipe (ncu: 94 % busy),
// and a quarter of its wavefronts are not atomics but the staging itself: 3.05 wavefronts per 128 cells of TMA smem fill plus 1.8 of
// LDS to read the staged rows back.  Here every warp loads its rows straight into registers with 128-bit global loads:
//   * a LANE PAIR owns a row: lane 2p / 2p+1 load the two 16-byte halves of row p's 32-byte tile sector (contiguous passes: one fully
//     coalesced 512-byte request per instruction; index-list leaves: 16 sectors per instruction, the same count the cp.async gathers
//     needed), two instructions cover the warp's 32 rows (row p and row 16 + p), plus two 16-byte loads of the rows' (g,h) words;
//   * a lane then adds 2 rows x 16 features.  At step (k, j) it takes byte (b + j) & 3 of word (a + k) & 3, with p = 4a + b: the 16 lanes
//     of one half cover the 16 (word, byte) pairs, so the 32 lanes of an ATOMS still address 32 different features = 32 banks;
//   * the next row group's loads are issued before the current group's atomics (register double buffering) — with 20 warps per SM a
//     warp comes back to its loads every ~2500 cycles, several DRAM latencies.
// Expected LSU wavefronts per 128 cells: 16.8 atomics + 1.5 loads (contiguous) / 4.75 (gathered) instead of 16.8 + 4.85 / 6.8.
#ifndef B200GBM_DL_WARPS
#define B200GBM_DL_WARPS 20
#endif
constexpr int kDlWarps = B200GBM_DL_WARPS;
constexpr int kDlThreads = kDlWarps * 32;
constexpr int kDlSmemBytes = 4 * kPlaneWords * 4;

struct DlGroup { uint4 ba, bb; int4 qa, qb; };

template <int NATOM>
__global__ void __launch_bounds__(kDlThreads, 1)
k4_hist_build_dl(const uint8_t* __restrict__ bins, size_t rows_stride, int num_tiles, const int4* __restrict__ qgh,
                 const int4* __restrict__ qord, const int* __restrict__ idx0, const int* __restrict__ idx1,
                 const HistWork* __restrict__ work, unsigned long long* __restrict__ hist) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned* plane = reinterpret_cast<unsigned*>(smem_raw);
  const HistWork w = *work;
  const int n = w.count;
  if (n <= 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int* __restrict__ idx = w.buf ? idx1 : idx0;
  // same work decomposition as v3: (tile, row chunk) items in tile-major order, one contiguous item range per CTA
  long long cells_rows = static_cast<long long>(n) * num_tiles;
  int rpi = static_cast<int>((cells_rows + 4LL * gridDim.x - 1) / (4LL * gridDim.x));
  rpi = (rpi + 511) / 512 * 512;
  rpi = max(rpi, 512);
  rpi = min(rpi, kFlushRows);
  const int chunks = (n + rpi - 1) / rpi;
  const long long items = static_cast<long long>(chunks) * num_tiles;
  const int i0 = static_cast<int>(items * blockIdx.x / gridDim.x);
  const int i1 = static_cast<int>(items * (blockIdx.x + 1) / gridDim.x);
  if (i0 >= i1) return;
  {
    uint4* p4 = reinterpret_cast<uint4*>(plane);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int e = tid; e < kPlaneWords; e += kDlThreads) p4[e] = z;      // 4 planes x 8192 words = 8192 uint4
  }
  __syncthreads();

  const int pr = lane >> 1, half = lane & 1;         // row pair index, which 16-byte half of the sector
  const int a = pr >> 2, b = pr & 3;
  const unsigned fbase = static_cast<unsigned>(half * 16);
  int acc_rows = 0;
  for (int item = i0; item < i1; ++item) {
    const int tile = item / chunks, chunk = item - tile * chunks;
    const int row0 = chunk * rpi;
    const int nrows = min(rpi, n - row0);
    const uint8_t* tbins = bins + static_cast<size_t>(tile) * rows_stride * 32 + half * 16;
    const int groups = (nrows + 31) >> 5;
    auto load = [&](int g, DlGroup* d, bool* va, bool* vb) {
      const int pa = row0 + g * 32 + pr, pb = pa + 16;
      *va = pa < row0 + nrows; *vb = pb < row0 + nrows;
      d->ba = d->bb = make_uint4(0u, 0u, 0u, 0u);
      d->qa = d->qb = make_int4(0, 0, 0, 0);
      if (w.use_idx) {
        if (*va) { const int r = idx[w.begin + pa]; d->ba = *reinterpret_cast<const uint4*>(tbins + static_cast<size_t>(r) * 32); d->qa = qord[w.begin + pa]; }
        if (*vb) { const int r = idx[w.begin + pb]; d->bb = *reinterpret_cast<const uint4*>(tbins + static_cast<size_t>(r) * 32); d->qb = qord[w.begin + pb]; }
      } else {
        if (*va) { const size_t r = static_cast<size_t>(w.begin + pa); d->ba = *reinterpret_cast<const uint4*>(tbins + r * 32); d->qa = qgh[r]; }
        if (*vb) { const size_t r = static_cast<size_t>(w.begin + pb); d->bb = *reinterpret_cast<const uint4*>(tbins + r * 32); d->qb = qgh[r]; }
      }
    };
    auto scatter = [&](const uint4& bw, const int4& q) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int ws = (a + k) & 3;
        const unsigned word = ws == 0 ? bw.x : ws == 1 ? bw.y : ws == 2 ? bw.z : bw.w;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kk = (b + j) & 3;
          const unsigned bin = (word >> (8 * kk)) & 0xFFu;
          const unsigned addr = bin * 32u + fbase + static_cast<unsigned>(ws * 4 + kk);
          atomicAdd(&plane[addr], static_cast<unsigned>(q.x));
          atomicAdd(&plane[kPlaneWords + addr], static_cast<unsigned>(q.y));
          atomicAdd(&plane[2 * kPlaneWords + addr], static_cast<unsigned>(q.z));
          if (NATOM == 4) atomicAdd(&plane[3 * kPlaneWords + addr], static_cast<unsigned>(q.w));
        }
      }
    };
    DlGroup cur, nxt;
    bool cva = false, cvb = false, nva = false, nvb = false;
    int g = warp;
    if (g < groups) load(g, &cur, &cva, &cvb);
    for (; g < groups; g += kDlWarps) {
      const int gn = g + kDlWarps;
      if (gn < groups) load(gn, &nxt, &nva, &nvb);
      if (cva) scatter(cur.ba, cur.qa);
      if (cvb) scatter(cur.bb, cur.qb);
      cur = nxt; cva = nva; cvb = nvb;
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
    __syncthreads();
    for (int e = tid; e < kPlaneWords; e += kDlThreads) {
      unsigned ghi = plane[e], glo = plane[kPlaneWords + e];
      unsigned hhi = plane[2 * kPlaneWords + e];
      unsigned hlo = (NATOM == 4) ? plane[3 * kPlaneWords + e] : 0u;
      if (ghi | glo | hhi | hlo) {
        const int f = tile * 32 + (e & 31);
        const int bb2 = e >> 5;
        long long gq = (static_cast<long long>(static_cast<int>(ghi)) << kLoBits) + static_cast<long long>(glo);
        long long hq = (NATOM == 4) ? (static_cast<long long>(static_cast<int>(hhi)) << kLoBits) + static_cast<long long>(hlo)
                                    : static_cast<long long>(hhi);
        const size_t o = (static_cast<size_t>(f) * kBins + bb2) * 2;
        if (gq) atomicAdd(&hist[o], static_cast<unsigned long long>(gq));
        if (hq) atomicAdd(&hist[o + 1], static_cast<unsigned long long>(hq));
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
