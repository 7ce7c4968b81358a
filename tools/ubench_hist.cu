// Microbenchmarks that decide the K4 histogram design (run on the B200 via gpurun).
//  Part A: shared-memory scatter-add throughput per SM for the candidate accumulator schemes
//          (native ATOMS.ADD.32 owner-bank / random-bank, CAS float, non-atomic owner RMW ...).
//  Part B: the production k4_hist_build kernel on synthetic tile-major bins, checked against a
//          host computation, timed with CUDA events, reported as cells/s and algorithmic GB/s.
// Output: one JSON document on stdout.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include "../mmlspark_b200/csrc/hist_kernel.cuh"
#include "k4_direct_load_experiment.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

using namespace b200gbm;

__device__ __host__ inline unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// ---------------------------------------------------------------- Part A
template <int MODE>
__global__ void __launch_bounds__(1024, 1)
ubench(int iters, unsigned long long* cyc_out, unsigned* sink) {
  extern __shared__ __align__(16) unsigned char sm[];
  unsigned* plane = reinterpret_cast<unsigned*>(sm);                  // 4 planes of 8192 words
  unsigned char* sb = sm + 4 * kPlaneWords * 4;                       // 256 rows x 32 B
  int4* sq = reinterpret_cast<int4*>(sb + 256 * 32);                  // 256 rows
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  for (int e = tid; e < 4 * kPlaneWords; e += blockDim.x) plane[e] = 0;
  for (int e = tid; e < 256 * 32; e += blockDim.x) sb[e] = hash32(e * 2654435761u + blockIdx.x) % 255;
  for (int e = tid; e < 256; e += blockDim.x) sq[e] = make_int4(e - 100, e * 7 + 1, 3, e + 5);
  __syncthreads();
  long long t0 = clock64();
  if (MODE == 9 || MODE == 10) {   // 4 rows x 32 features per warp step (the production mapping); 10 = 3 planes
    const unsigned* sw = reinterpret_cast<const unsigned*>(sb);
    const int rsel = lane >> 3, wsel = lane & 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
      for (int g4 = warp; g4 < 64; g4 += nwarp) {
        const int r = g4 * 4 + rsel;
        const unsigned word = sw[r * 8 + wsel];
        const int4 q = sq[r];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int kk = (k + rsel) & 3;
          const unsigned b = (word >> (8 * kk)) & 0xFFu;
          const unsigned a = b * 32u + (unsigned)(wsel * 4 + kk);
          atomicAdd(&plane[a], (unsigned)q.x); atomicAdd(&plane[kPlaneWords + a], (unsigned)q.y);
          atomicAdd(&plane[2 * kPlaneWords + a], (unsigned)q.z);
          if (MODE == 9) atomicAdd(&plane[3 * kPlaneWords + a], (unsigned)q.w);
        }
      }
      __syncthreads();
    }
  } else
  for (int it = 0; it < iters; ++it) {
#pragma unroll 4
    for (int r = warp; r < 256; r += nwarp) {
      unsigned b = sb[r * 32 + lane];
      int4 q = sq[r];
      unsigned a = (MODE == 3) ? (lane * 256u + b) : (b * 32u + lane);
      if (MODE == 0 || MODE == 3) {
        atomicAdd(&plane[a], (unsigned)q.x); atomicAdd(&plane[kPlaneWords + a], (unsigned)q.y);
        atomicAdd(&plane[2 * kPlaneWords + a], (unsigned)q.z); atomicAdd(&plane[3 * kPlaneWords + a], (unsigned)q.w);
      } else if (MODE == 1) {
        atomicAdd(&plane[a], (unsigned)q.x); atomicAdd(&plane[kPlaneWords + a], (unsigned)q.y);
      } else if (MODE == 2) {
        atomicAdd(&plane[a], (unsigned)q.x);
      } else if (MODE == 6) {
        atomicAdd(&plane[a], (unsigned)q.x); atomicAdd(&plane[kPlaneWords + a], (unsigned)q.y);
        atomicAdd(&plane[2 * kPlaneWords + a], (unsigned)q.z);
      } else if (MODE == 4) {   // float CAS-loop atomics
        float* fp = reinterpret_cast<float*>(plane);
        atomicAdd(&fp[a], __int_as_float(q.x) * 1e-30f); atomicAdd(&fp[kPlaneWords + a], 1.0f);
      } else if (MODE == 5) {   // non-atomic owner RMW, 64-bit g and h (throughput probe only)
        unsigned long long* p64 = reinterpret_cast<unsigned long long*>(plane);
        unsigned long long g = p64[a], h = p64[kPlaneWords + a];
        p64[a] = g + (unsigned)q.x; p64[kPlaneWords + a] = h + (unsigned)q.y;
      } else if (MODE == 7) {   // loads only
        if (b == 255u && q.x == 123456) plane[a] = 1;
      } else if (MODE == 8) {   // one 64-bit CAS-loop atomic
        unsigned long long* p64 = reinterpret_cast<unsigned long long*>(plane);
        atomicAdd(&p64[a], (unsigned long long)(unsigned)q.x);
      }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0) cyc_out[blockIdx.x] = (unsigned long long)(t1 - t0);
  unsigned acc = 0;
  for (int e = tid; e < 4 * kPlaneWords; e += blockDim.x) acc += plane[e];
  if (acc == 0xdeadbeef) sink[0] = acc;
}

template <int MODE>
static double run_mode(int threads, int iters, int nsm) {
  unsigned long long* d_cyc; unsigned* d_sink;
  CK(cudaMalloc(&d_cyc, nsm * sizeof(unsigned long long))); CK(cudaMalloc(&d_sink, 4));
  int smem = 4 * kPlaneWords * 4 + 256 * 32 + 256 * 16;
  CK(cudaFuncSetAttribute(ubench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  ubench<MODE><<<nsm, threads, smem>>>(8, d_cyc, d_sink);
  ubench<MODE><<<nsm, threads, smem>>>(iters, d_cyc, d_sink);
  CK(cudaDeviceSynchronize());
  std::vector<unsigned long long> c(nsm);
  CK(cudaMemcpy(c.data(), d_cyc, nsm * 8, cudaMemcpyDeviceToHost));
  double s = 0; for (auto v : c) s += (double)v;
  s /= nsm;
  CK(cudaFree(d_cyc)); CK(cudaFree(d_sink));
  return (double)iters * 256.0 * 32.0 / s;   // cells per cycle per SM
}

// ---------------------------------------------------------------- Part B
__global__ void gen_bins(uint8_t* bins, size_t rows_stride, int num_tiles, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;   // one 4-byte word each
  size_t total = (size_t)num_tiles * rows_stride * 8;
  for (; i < total; i += (size_t)gridDim.x * blockDim.x) {
    unsigned w = 0;
    for (int k = 0; k < 4; ++k) w |= (hash32((unsigned)(i * 4 + k) ^ seed) % 255u) << (8 * k);
    reinterpret_cast<unsigned*>(bins)[i] = w;
  }
}
__global__ void gen_q(int4* q, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
    long long g = (long long)(int)hash32((unsigned)i ^ seed) * 11LL;                 // signed ~36 bit
    long long h = (long long)(hash32((unsigned)i * 3u + seed) >> 1) * 9LL;           // unsigned ~35 bit
    q[i] = make_int4((int)(g >> kLoBits), (int)(g & ((1 << kLoBits) - 1)), (int)(h >> kLoBits), (int)(h & ((1 << kLoBits) - 1)));
  }
}
__global__ void gen_idx(int* idx, int n, int stride) {   // every `stride`-th row, ascending
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n; i += gridDim.x * blockDim.x) idx[i] = i * stride;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  int nsm = prop.multiProcessorCount;
  printf("{\n \"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", prop.name, nsm, prop.clockRate);

  // ---- Part A
  printf(" \"smem_scatter_cells_per_cycle_per_sm\": {\n");
  const int it = 400;
  for (int threads : {512, 1024}) {
    if (argc > 3 && atoi(argv[3]) == 1) { printf("  \"t%d\": {}%s\n", threads, threads == 512 ? "," : ""); continue; }
    printf("  \"t%d\": {", threads);
    printf("\"u32x4_rows4x32_rotated\": %.3f, ", run_mode<9>(threads, it, nsm));
    printf("\"u32x3_rows4x32_rotated\": %.3f, ", run_mode<10>(threads, it, nsm));
    printf("\"u32x4_ownerbank\": %.3f, ", run_mode<0>(threads, it, nsm));
    printf("\"u32x3_ownerbank\": %.3f, ", run_mode<6>(threads, it, nsm));
    printf("\"u32x2_ownerbank\": %.3f, ", run_mode<1>(threads, it, nsm));
    printf("\"u32x1_ownerbank\": %.3f, ", run_mode<2>(threads, it, nsm));
    printf("\"u32x4_randombank\": %.3f, ", run_mode<3>(threads, it, nsm));
    printf("\"f32x2_cas_ownerbank\": %.3f, ", run_mode<4>(threads, it / 4, nsm));
    printf("\"u64x2_nonatomic_rmw\": %.3f, ", run_mode<5>(threads, it, nsm));
    printf("\"u64x1_cas\": %.3f, ", run_mode<8>(threads, it / 4, nsm));
    printf("\"loads_only\": %.3f}%s\n", run_mode<7>(threads, it, nsm), threads == 512 ? "," : "");
  }
  printf(" },\n");

  if (const char* e = getenv("B200GBM_L2_FETCH")) {
    CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e)));
  }
  { size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf(" \"l2_fetch_granularity\": %zu,\n", g); }
  // ---- Part B
  const int F = 256, num_tiles = F / 32;
  size_t N = (argc > 1) ? (size_t)atoll(argv[1]) : 10000000;
  size_t rows_stride = (N + 255) / 256 * 256;
  uint8_t* d_bins; int4* d_q; int4* d_qord; int* d_idx; unsigned long long* d_hist; HistWork* d_work;
  size_t slot_elems = (size_t)F * 256 * 2;
  CK(cudaMalloc(&d_bins, (size_t)num_tiles * rows_stride * 32));
  CK(cudaMalloc(&d_q, N * sizeof(int4)));
  CK(cudaMalloc(&d_qord, N * sizeof(int4)));
  CK(cudaMalloc(&d_idx, N * sizeof(int)));
  CK(cudaMalloc(&d_hist, slot_elems * 8 * 2));
  CK(cudaMalloc(&d_work, sizeof(HistWork) * 4));
  gen_bins<<<nsm * 8, 256>>>(d_bins, rows_stride, num_tiles, N, 12345u);
  gen_q<<<nsm * 8, 256>>>(d_q, N, 777u);
  CK(cudaDeviceSynchronize());
  CK(cudaFuncSetAttribute(k4_hist_build<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmemBytes));
  CK(cudaFuncSetAttribute(k4_hist_build<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmemBytes));
  CK(cudaFuncSetAttribute(k4_hist_build_ws<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmemBytes));
  CK(cudaFuncSetAttribute(k4_hist_build_ws<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmemBytes));
  const int variant = argc > 2 ? atoi(argv[2]) : 0;      // 0 = v2, 1 = v3/v4 warp-specialised (production in round 1), 2 = v5 direct-load
  const bool use_ws = variant == 1;
  const bool use_dl = variant == 2;
  CK(cudaFuncSetAttribute(k4_hist_build_dl<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDlSmemBytes));
  CK(cudaFuncSetAttribute(k4_hist_build_dl<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDlSmemBytes));

  // correctness on the first 1M rows (contiguous) and on a strided index list
  {
    int n_chk = (int)std::min<size_t>(N, 1000000);
    HistWork hw[2] = {{0, n_chk, 0, 0}, {0, n_chk / 3, 1, 0}};
    CK(cudaMemcpy(d_work, hw, sizeof(hw), cudaMemcpyHostToDevice));
    gen_idx<<<nsm, 256>>>(d_idx, n_chk / 3, 3);
    CK(cudaMemset(d_hist, 0, slot_elems * 8 * 2));
    if (use_dl) {
      k4_hist_build_dl<4><<<nsm, kDlThreads, kDlSmemBytes>>>(d_bins, rows_stride, num_tiles, d_q, d_qord, d_idx, d_idx, d_work, d_hist);
      k_gather_q<<<nsm * 8, 256>>>(d_work + 1, d_idx, d_idx, d_q, d_qord);
      k4_hist_build_dl<4><<<nsm, kDlThreads, kDlSmemBytes>>>(d_bins, rows_stride, num_tiles, d_q, d_qord, d_idx, d_idx, d_work + 1, d_hist + slot_elems);
    } else if (use_ws) {
      k4_hist_build_ws<4><<<nsm, kWsThreads, kWsSmemBytes>>>(d_bins, rows_stride, num_tiles, d_q, d_qord, d_idx, d_idx, d_work, d_hist);
      k_gather_q<<<nsm * 8, 256>>>(d_work + 1, d_idx, d_idx, d_q, d_qord);
      k4_hist_build_ws<4><<<nsm, kWsThreads, kWsSmemBytes>>>(d_bins, rows_stride, num_tiles, d_q, d_qord, d_idx, d_idx, d_work + 1, d_hist + slot_elems);
    } else {
    k4_hist_build<4><<<nsm, kHistThreads, kHistSmemBytes>>>(d_bins, rows_stride, num_tiles, d_q, d_idx, d_idx, d_work, d_hist);
    k4_hist_build<4><<<nsm, kHistThreads, kHistSmemBytes>>>(d_bins, rows_stride, num_tiles, d_q, d_idx, d_idx, d_work + 1, d_hist + slot_elems);
    }
    CK(cudaDeviceSynchronize());
    std::vector<long long> got(slot_elems * 2), want(slot_elems * 2, 0);
    CK(cudaMemcpy(got.data(), d_hist, slot_elems * 16, cudaMemcpyDeviceToHost));
    std::vector<uint8_t> hb((size_t)num_tiles * rows_stride * 32);
    CK(cudaMemcpy(hb.data(), d_bins, hb.size(), cudaMemcpyDeviceToHost));
    std::vector<int4> hq(n_chk);
    CK(cudaMemcpy(hq.data(), d_q, (size_t)n_chk * 16, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n_chk; ++i) {
      long long g = ((long long)hq[i].x << kLoBits) + hq[i].y, h = ((long long)hq[i].z << kLoBits) + hq[i].w;
      bool in2 = (i % 3 == 0) && (i / 3 < n_chk / 3);
      for (int t = 0; t < num_tiles; ++t) {
        const uint8_t* row = &hb[((size_t)t * rows_stride + i) * 32];
        for (int l = 0; l < 32; ++l) {
          size_t o = ((size_t)(t * 32 + l) * 256 + row[l]) * 2;
          want[o] += g; want[o + 1] += h;
          if (in2) { want[slot_elems + o] += g; want[slot_elems + o + 1] += h; }
        }
      }
    }
    size_t bad = 0; for (size_t i = 0; i < want.size(); ++i) bad += (got[i] != want[i]);
    printf(" \"k4_check\": {\"rows\": %d, \"mismatches\": %zu},\n", n_chk, bad);
  }

  // timing: full pass (contiguous) and gathered pass (every 2nd row), NATOM 4 and 3
  auto time_it = [&](int natom, int n, int use_idx, int reps) -> float {
    HistWork hw = {0, n, use_idx, 0};
    CK(cudaMemcpy(d_work, &hw, sizeof(hw), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f, tot = 0;
    for (int r = 0; r < reps + 2; ++r) {
      CK(cudaMemsetAsync(d_hist, 0, slot_elems * 8));
      CK(cudaEventRecord(e0));
      if ((use_ws || use_dl) && use_idx) k_gather_q<<<nsm * 8, 256>>>(d_work, d_idx, d_idx, d_q, d_qord);     // part of a leaf pass: timed
      if (use_dl) {
        if (natom == 4) k4_hist_build_dl<4><<<nsm, kDlThreads, kDlSmemBytes>>>(d_bins, rows_stride, num_tiles, d_q, d_qord, d_idx, d_idx, d_work, d_hist);
        else k4_hist_build_dl<3><<<nsm, kDlThreads, kDlSmemBytes>>>(d_bins, rows_stride, num_tiles, d_q, d_qord, d_idx, d_idx, d_work, d_hist);
      } else if (use_ws) {
        if (natom == 4) k4_hist_build_ws<4><<<nsm, kWsThreads, kWsSmemBytes>>>(d_bins, rows_stride, num_tiles, d_q, d_qord, d_idx, d_idx, d_work, d_hist);
        else k4_hist_build_ws<3><<<nsm, kWsThreads, kWsSmemBytes>>>(d_bins, rows_stride, num_tiles, d_q, d_qord, d_idx, d_idx, d_work, d_hist);
      } else if (natom == 4) k4_hist_build<4><<<nsm, kHistThreads, kHistSmemBytes>>>(d_bins, rows_stride, num_tiles, d_q, d_idx, d_idx, d_work, d_hist);
      else k4_hist_build<3><<<nsm, kHistThreads, kHistSmemBytes>>>(d_bins, rows_stride, num_tiles, d_q, d_idx, d_idx, d_work, d_hist);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (r >= 2) { best = std::min(best, ms); tot += ms; }
    }
    (void)best;
    return tot / reps;
  };
  printf(" \"k4_timing\": [\n");
  struct Cfg { int natom; double frac; int use_idx; };
  Cfg cfgs[] = {{4, 1.0, 0}, {3, 1.0, 0}, {4, 0.5, 1}, {4, 0.1, 1}, {4, 0.01, 1}, {4, 0.001, 1}};
  for (size_t c = 0; c < sizeof(cfgs) / sizeof(cfgs[0]); ++c) {
    int n = (int)(N * cfgs[c].frac);
    if (cfgs[c].use_idx) { gen_idx<<<nsm, 256>>>(d_idx, n, (int)(1.0 / cfgs[c].frac)); CK(cudaDeviceSynchronize()); }
    float ms = time_it(cfgs[c].natom, n, cfgs[c].use_idx, 5);
    double cells = (double)n * F;
    double bytes = (double)n * F + (double)n * 16 * num_tiles / num_tiles /*qgh once per tile below*/;
    bytes = (double)n * (F + 16.0 * num_tiles + (cfgs[c].use_idx ? 4.0 * num_tiles : 0.0)) + (double)F * 256 * 16;
    printf("  {\"natom\": %d, \"rows\": %d, \"gather\": %d, \"ms\": %.4f, \"gcells_per_s\": %.2f, \"algo_GBps\": %.1f}%s\n",
           cfgs[c].natom, n, cfgs[c].use_idx, ms, cells / ms * 1e-6, bytes / ms * 1e-6, c + 1 < sizeof(cfgs) / sizeof(cfgs[0]) ? "," : "");
  }
  printf(" ]\n}\n");
  return 0;
}
