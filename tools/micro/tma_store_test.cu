// Micro-test of 2-D / 3-D TMA stores with clipping (development aid for the GEMM epilogue).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../../speech_diarization_b200/csrc/gemm_host.cuh"
using namespace sd;

__global__ void k_store(const __grid_constant__ CUtensorMap m, int rank, int c0, int c1, int c2) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __half* t = reinterpret_cast<__half*>(sm);
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int row = i / 64, col = i % 64;
    const int piece = col / 8;
    const int off = row * 128 + ((piece ^ (row & 7)) << 4) + (col % 8) * 2;
    *reinterpret_cast<__half*>(sm + off) = __float2half(row + col * 0.001f);
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (rank == 3) tma_store_3d(&m, sm, c0, c1, c2);
    else tma_store_2d(&m, sm, c0, c1);
    tma_store_commit();
    tma_store_wait_all();
  }
}

int run(const char* name, int rank, int T, int Tp, int B, int ld, int c0, int c1, int c2) {
  const int H = 4;
  const long rows = (long)B * Tp;
  __half* d;
  cudaMalloc(&d, rows * ld * 2);
  cudaMemset(d, 0, rows * ld * 2);
  CUtensorMap m;
  int st = rank == 3 ? make_tmap_f16_interior(&m, d, ld, Tp, T, H, B) : make_tmap_f16(&m, d, rows, ld, ld, 128);
  if (st) { printf("%s: encode failed %d\n", name, st); return 1; }
  cudaFuncSetAttribute(k_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  k_store<<<1, 128, 16384>>>(m, rank, c0, c1, c2);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: KERNEL ERROR %s\n", name, cudaGetErrorString(e)); return 2; }
  std::vector<__half> h(rows * ld);
  cudaMemcpy(h.data(), d, rows * ld * 2, cudaMemcpyDeviceToHost);
  long nz = 0, bad = 0;
  for (long r = 0; r < rows; ++r)
    for (int c = 0; c < ld; ++c) {
      float v = __half2float(h[r * ld + c]);
      if (v != 0.f) {
        ++nz;
        int srow, scol = c - c0;
        if (rank == 3) { int b = r / Tp, t = r % Tp - H; srow = t - c1; if (b != c2 || t < 0 || t >= T) ++bad; }
        else srow = r - c1;
        float want = __half2float(__float2half(srow + scol * 0.001f));
        if (scol < 0 || scol >= 64 || srow < 0 || srow >= 128 || v != want) ++bad;
      }
    }
  printf("%s: ok nonzero=%ld bad=%ld\n", name, nz, bad);
  cudaFree(d);
  return bad != 0;
}

int main() {
  int f = 0;
  f |= run("2d in-bounds", 2, 0, 160, 3, 1024, 64, 128, 0);
  f |= run("2d clipped rows", 2, 0, 160, 3, 1024, 0, 400, 0);
  f |= run("3d t0=0", 3, 151, 160, 3, 1024, 0, 0, 1);
  f |= run("3d t0=-4", 3, 151, 160, 3, 1024, 64, -4, 0);
  f |= run("3d t0=124", 3, 151, 160, 3, 1024, 0, 124, 0);
  f |= run("3d t0=-36 b=1", 3, 151, 160, 3, 1024, 0, -36, 1);
  f |= run("3d short T=26", 3, 26, 48, 2, 1024, 0, -4, 0);
  f |= run("3d short T=26 t0=-52", 3, 26, 48, 2, 1024, 0, -52, 1);
  f |= run("3d ld=128", 3, 151, 160, 3, 128, 64, -4, 2);
  return f;
}
