// api_core.cu — version / error strings and the raw tensor-core GEMM test hook.
#include "gemm_host.cuh"

using namespace sd;

extern "C" int sd_version(void) { return 100; }

extern "C" const char* sd_status_string(int s) {
  switch (s) {
    case SD_OK: return "ok";
    case SD_ERR_ARG: return "bad argument";
    case SD_ERR_CUDA: return "CUDA error";
    case SD_ERR_DRIVER: return "TMA descriptor encode failed";
    case SD_ERR_NOMEM: return "out of device memory";
    case SD_ERR_UNSUPPORTED: return "unsupported";
    case SD_ERR_MISSING: return "missing weight tensor";
    case SD_ERR_RANGE: return "activation out of f16 range";
  }
  return "unknown";
}

extern "C" const char* sd_last_error(void) { return err_buf(); }

extern "C" long sd_launch_count(void) { return launch_counter().load(); }

extern "C" int sd_debug_gemm_f16(const void* A, int M, int K, const void* B, int N, int taps,
                                 int dil, int n_tile, float* D, void* stream) {
  if (!A || !B || !D || M < 1 || N < 1 || K < 64 || K % 64 || taps < 1 || (taps & 1) == 0 ||
      (K / 64) * taps > MAX_KITERS)
    return fail(SD_ERR_ARG, "sd_debug_gemm_f16: bad shape M=%d N=%d K=%d taps=%d", M, N, K, taps);
  GemmParams P;
  init_params(P);
  SD_TRY(make_tmap_f16(&P.tmapA, A, M, K, K, BM));
  SD_TRY(make_tmap_f16(&P.tmapB, B, N, (long)taps * K, (long)taps * K, n_tile));
  P.num_m_blocks = (M + BM - 1) / BM;
  P.num_n_blocks = (N + n_tile - 1) / n_tile;
  P.n_tile = n_tile;
  P.idesc = make_idesc_f16(n_tile, 0);
  int ki = 0;
  for (int j = 0; j < taps; ++j)
    for (int c = 0; c < K / 64; ++c, ++ki) {
      P.kit[ki].a_col = c * 64;
      P.kit[ki].a_row_off = (j - taps / 2) * dil;
      P.kit[ki].b_col = j * K + c * 64;
      P.kit[ki].slot = 0;
      P.kit[ki].accum = ki > 0;
    }
  P.num_kiters = ki;
  P.epi.M_rows = M;
  P.epi.N_cols = N;
  P.epi.out = D;
  P.epi.ld_out = N;
  return launch_gemm<EPI_F32>(P, (cudaStream_t)stream);
}
