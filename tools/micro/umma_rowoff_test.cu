// Micro-test: can a K-major SWIZZLE_128B A operand start at an arbitrary ROW offset inside a TMA-loaded
// box, using the descriptor's "matrix base offset" field?  (Would let one 128+2d-row load serve all three
// taps of a dilated conv.)  D[128 x 128] = A[off : off+128, 0:64] * B[128, 64]^T for off = 0..8.
#include <cstdio>
#include <vector>
#include <cmath>
#include <cuda_fp16.h>
#include "../../speech_diarization_b200/csrc/gemm_host.cuh"
using namespace sd;

__global__ void __launch_bounds__(128) k_test(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB,
                                              int off, int mode, float* D) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* sA = sm;            // 144 rows x 128 B = 18 KB
  uint8_t* sB = sm + 18432;    // 128 rows x 128 B = 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 18432 + 16384);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
  if (warp == 0) { __syncwarp(); tmem_alloc(slot, 128); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 18432 + 16384);
    tma_load_2d(sA, &mA, bar, 0, 0);
    tma_load_2d(sB, &mB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sA) + off * 128;
    uint64_t da = make_smem_desc_sw128(a_addr);
    if (mode == 1) da |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;  // matrix base offset
    const uint64_t db = make_smem_desc_sw128(smem_u32(sB));
    const uint32_t idesc = make_idesc_f16(128, 0);
    for (int kk = 0; kk < 4; ++kk) umma_f16(tm, da + 2 * kk, db + 2 * kk, idesc, kk ? 1u : 0u);
    umma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 128; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tm + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 128); }
}

int main() {
  const int RA = 144, K = 64, N = 128;
  std::vector<__half> hA(RA * K), hB(N * K);
  srand(1);
  for (auto& x : hA) x = __float2half((rand() % 200 - 100) / 100.f);
  for (auto& x : hB) x = __float2half((rand() % 200 - 100) / 100.f);
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 128 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap mA, mB;
  if (make_tmap_f16(&mA, dA, RA, K, K, RA) || make_tmap_f16(&mB, dB, N, K, K, N)) { printf("encode failed\n"); return 1; }
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  for (int mode = 0; mode < 2; ++mode)
    for (int off = 0; off <= 9; ++off) {
      cudaMemset(dD, 0, 128 * 128 * 4);
      k_test<<<1, 128, 40960>>>(mA, mB, off, mode, dD);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d off %d: %s\n", mode, off, cudaGetErrorString(e)); return 2; }
      std::vector<float> hD(128 * 128);
      cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)__half2float(hA[(m + off) * K + k]) * __half2float(hB[n * K + k]);
          maxerr = fmax(maxerr, fabs(ref - hD[m * 128 + n]));
        }
      printf("mode %d (base_offset %s) off %d: max err %.4g %s\n", mode, mode ? "set" : "0", off, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    }
  return 0;
}
