// gemm_host.cuh — host side of gemm_tc.cuh: TMA tensor-map encoding (driver entry
// point fetched at run time, so the library loads on a box without libcuda) and
// the launch helper.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include "gemm_tc.cuh"
#include "sd_status.h"

namespace sd {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled get_tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  return fn;
}

// 2-D f16 tensor [rows][cols] with row pitch `ld` elements; box = 64 columns x box_rows,
// 128-byte swizzle, out-of-bounds elements read as zero (negative coordinates allowed).
inline int make_tmap_f16(CUtensorMap* m, const void* base, long rows, long cols, long ld,
                         int box_rows) {
  PFN_tmapEncodeTiled enc = get_tmap_encoder();
  if (!enc) return SD_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16 || box_rows < 1 || box_rows > 256)
    return SD_ERR_ARG;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SD_OK : SD_ERR_DRIVER;
}

// 3-D f16 view for REFLECT-layer stores: {channels, T interior frames, utterances}.  `base` is the
// activation tensor [B*Tp, ld]; element (c, t, b) lives at base + ((b*Tp + H + t) * ld + c).
// Boxes are 64 channels x 128 frames x 1 utterance; frames outside [0, T) are clipped on store.
inline int make_tmap_f16_interior(CUtensorMap* m, const void* base, long ld, int Tp, int T, int H, int B) {
  PFN_tmapEncodeTiled enc = get_tmap_encoder();
  if (!enc) return SD_ERR_DRIVER;
  const char* p = static_cast<const char*>(base) + static_cast<size_t>(H) * ld * 2;
  if ((reinterpret_cast<uintptr_t>(p) & 15) || (ld * 2) % 16) return SD_ERR_ARG;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(ld), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(Tp) * ld * 2};
  cuuint32_t box[3] = {BK, BM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<char*>(p), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SD_OK : SD_ERR_DRIVER;
}

// The same 3-D view {channels, T interior frames, windows} with a small UNSWIZZLED box (box_cols channels x box_rows
// frames x 1 window; box_cols * 2 bytes a multiple of 16): the target of per-warp TMA stores from a row-major staging
// tile (res2net_pipe.cuh).  Frames >= T are clipped on store, so a partial last block cannot touch the padding rows.
inline int make_tmap_f16_interior_plain(CUtensorMap* m, const void* base, long ld, int Tp, int T, int H, int B,
                                        int box_cols, int box_rows) {
  PFN_tmapEncodeTiled enc = get_tmap_encoder();
  if (!enc) return SD_ERR_DRIVER;
  const char* p = static_cast<const char*>(base) + static_cast<size_t>(H) * ld * 2;
  if ((reinterpret_cast<uintptr_t>(p) & 15) || (ld * 2) % 16 || (box_cols * 2) % 16 || box_rows < 1 || box_rows > 256)
    return SD_ERR_ARG;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(ld), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(Tp) * ld * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<char*>(p), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SD_OK : SD_ERR_DRIVER;
}

// cudaFuncSetAttribute is per device: remember which devices already have it for a given kernel
inline bool attr_needed(bool (&done)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

// Programmatic dependent launch (sd_ptx.cuh: pdl_wait / pdl_trigger): while this flag is set, launch_pdl()
// adds cudaLaunchAttributeProgrammaticStreamSerialization, so a kernel's prologue overlaps its predecessor's
// tail.  ecapa.cu sets it around the trunk (every kernel there calls pdl_wait before touching global memory).
inline bool& pdl_flag() {
  static thread_local bool on = false;
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_flag() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int EPI, int MAX_BN>
inline int launch_gemm_t(const GemmParams& P, cudaStream_t stream) {
  using Cfg = GemmCfg<EPI, MAX_BN>;
  static bool attr_done[64] = {};
  if (attr_needed(attr_done) &&
      cudaFuncSetAttribute(gemm_tc_kernel<EPI, MAX_BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           Cfg::SMEM_BYTES) != cudaSuccess)
    return SD_ERR_CUDA;
  const int tiles = P.num_m_blocks * P.num_n_blocks * (P.k_splits > 1 ? P.k_splits : 1);
  if (tiles <= 0) return SD_OK;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  cudaError_t e = launch_pdl(gemm_tc_kernel<EPI, MAX_BN>, dim3(grid), dim3(gemm_threads_of(EPI)), Cfg::SMEM_BYTES, stream, P);
  count_launch();
  if (e == cudaSuccess) e = cudaGetLastError();
  static const bool sync_debug = getenv("SD_SYNC_DEBUG") != nullptr;  // localise a faulting launch
  if (e == cudaSuccess && sync_debug) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess)
    return fail(SD_ERR_CUDA, "gemm_tc_kernel<EPI=%d,MAX_BN=%d> n_tile=%d kiters=%d tiles=%d flags=%d: %s", EPI, MAX_BN,
                P.n_tile, P.num_kiters, tiles, P.epi.flags, cudaGetErrorString(e));
  return SD_OK;
}

#if SD_EXPERIMENTS
// 2-CTA-cluster launch with the B tile multicast (gemm_tc_mc_kernel).  tmapB's box must hold n_tile / 2 rows.
template <int EPI, int MAX_BN>
inline int launch_gemm_mc_t(const GemmParams& P, cudaStream_t stream) {
  using Cfg = GemmCfg<EPI, MAX_BN>;
  static bool attr_done[64] = {};
  if (attr_needed(attr_done) &&
      cudaFuncSetAttribute(gemm_tc_mc_kernel<EPI, MAX_BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           Cfg::SMEM_BYTES) != cudaSuccess)
    return SD_ERR_CUDA;
  const int units = ((P.num_m_blocks + 1) / 2) * P.num_n_blocks;
  if (units <= 0) return SD_OK;
  int grid = 2 * units < num_sms() ? 2 * units : (num_sms() & ~1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_mc_kernel<EPI, MAX_BN>, P);
  count_launch();
  static const bool sync_debug = getenv("SD_SYNC_DEBUG") != nullptr;
  if (e == cudaSuccess && sync_debug) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess)
    return fail(SD_ERR_CUDA, "gemm_tc_mc_kernel<EPI=%d,MAX_BN=%d> n_tile=%d kiters=%d units=%d: %s", EPI, MAX_BN,
                P.n_tile, P.num_kiters, units, cudaGetErrorString(e));
  return SD_OK;
}

#endif  // SD_EXPERIMENTS

// cta_group::2 launch (gemm_tc_2sm_kernel): EPI_TDNN, n_tile = 256, idesc with M = 256, tmapB box = 128 rows.
inline int launch_gemm_2sm(const GemmParams& P, cudaStream_t stream) {
  using Cfg = Cfg2sm;
  if (P.n_tile != 256 || P.num_kiters < 1 || P.num_kiters > MAX_KITERS) return SD_ERR_ARG;
  static bool attr_done[64] = {};
  if (attr_needed(attr_done) &&
      cudaFuncSetAttribute(gemm_tc_2sm_kernel<EPI_TDNN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           Cfg::SMEM_BYTES) != cudaSuccess)
    return SD_ERR_CUDA;
  const int units = ((P.num_m_blocks + 1) / 2) * P.num_n_blocks;
  if (units <= 0) return SD_OK;
  const int grid = 2 * units < num_sms() ? 2 * units : (num_sms() & ~1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_flag() ? 2 : 1;
  static const char* const trace_path = SD_TRACE_ON ? getenv("SD_GEMM_TRACE") : nullptr;   // read once: this runs per launch
  if (const char* path = trace_path) {   // debug: clock stamps of CTA 0 for launches with num_kiters == SD_GEMM_TRACE_K
    static const char* const ks = getenv("SD_GEMM_TRACE_K");
    if (P.num_kiters == (ks ? atoi(ks) : 16)) {
      GemmParams T = P;
      long long* dev = nullptr;
      std::vector<long long> host(32 * 16, 0);
      if (cudaMalloc(&dev, host.size() * 8) != cudaSuccess) return SD_ERR_CUDA;
      cudaMemsetAsync(dev, 0, host.size() * 8, stream);
      T.trace = dev;
      cudaLaunchKernelEx(&cfg, gemm_tc_2sm_kernel<EPI_TDNN>, T);
      cudaStreamSynchronize(stream);
      cudaMemcpy(host.data(), dev, host.size() * 8, cudaMemcpyDeviceToHost);
      cudaFree(dev);
      if (FILE* f = fopen(path, "w")) {
        for (int t = 0; t < 32; ++t) {
          for (int k = 0; k < 16; ++k) fprintf(f, "%lld ", host[t * 16 + k] ? host[t * 16 + k] - host[0] : -1LL);
          fprintf(f, "\n");
        }
        fclose(f);
      }
      count_launch();
      return SD_OK;
    }
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_2sm_kernel<EPI_TDNN>, P);
  count_launch();
  static const bool sync_debug = getenv("SD_SYNC_DEBUG") != nullptr;
  if (e == cudaSuccess && sync_debug) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess)
    return fail(SD_ERR_CUDA, "gemm_tc_2sm_kernel n_tile=%d kiters=%d units=%d: %s", P.n_tile, P.num_kiters, units,
                cudaGetErrorString(e));
  return SD_OK;
}

// Chooses the shared-memory configuration from n_tile.
template <int EPI>
inline int launch_gemm(const GemmParams& P, cudaStream_t stream) {
  if (P.n_tile % 16 || P.n_tile < 16 || P.n_tile > 256 || P.num_kiters < 1 ||
      P.num_kiters > MAX_KITERS || P.acc_slots * P.n_tile > TMEM_COLS)
    return SD_ERR_ARG;
  if constexpr (EPI == EPI_CONV3) {
    if (P.n_tile != 128 || P.conv_taps != 3 || P.num_kiters != 2 || P.conv_dil < 1 || 2 * P.conv_dil > 16)
      return SD_ERR_ARG;
    return launch_gemm_t<EPI, 128>(P, stream);
  } else {
    if (P.n_tile <= 128) return launch_gemm_t<EPI, 128>(P, stream);
    return launch_gemm_t<EPI, 256>(P, stream);
  }
}

#if SD_EXPERIMENTS
// Cooperative launch of a chain of dependent GEMMs (gemm_chain_kernel); `dev_steps` is a device
// array of num_steps GemmParams.  n_tile of every step must fit MAX_BN.
template <int EPI, int MAX_BN>
inline int launch_gemm_chain(const GemmParams* dev_steps, int num_steps, cudaStream_t stream) {
  using Cfg = GemmCfg<EPI, MAX_BN>;
  static bool attr_done[64] = {};
  if (attr_needed(attr_done) &&
      cudaFuncSetAttribute(gemm_chain_kernel<EPI, MAX_BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           Cfg::SMEM_BYTES) != cudaSuccess)
    return SD_ERR_CUDA;
  void* args[] = {(void*)&dev_steps, (void*)&num_steps};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)gemm_chain_kernel<EPI, MAX_BN>, dim3(num_sms()),
                                              dim3(GEMM_THREADS), args, Cfg::SMEM_BYTES, stream);
  count_launch();
  if (e != cudaSuccess) return fail(SD_ERR_CUDA, "cooperative launch failed: %s", cudaGetErrorString(e));
  return SD_OK;
}

#endif  // SD_EXPERIMENTS

inline void init_params(GemmParams& P) {
  std::memset(&P, 0, sizeof(P));
  P.acc_slots = 1;
  P.n_sub = 1;
  P.k_splits = 1;
  P.epi.Tp = 1;
}

}  // namespace sd
