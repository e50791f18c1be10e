// gemm_tc.cuh — one persistent, warp-specialised tcgen05 GEMM kernel that carries
// every tensor-core contraction of the hot path:
//
//   D[128 x n_tile] (f32, TMEM)  =  sum over "k-iterations"  A_tile[128 x 64] * B_tile[n_tile x 64]^T
//
// A k-iteration is one 64-element K chunk; a small table (KIter) says, for each
// one, which column of the A / B tensors to fetch, which ROW OFFSET to apply to
// A (that is how the dilated k=3 / k=5 convolutions of ECAPA-TDNN become
// implicit GEMMs over a channels-last activation tensor — the taps are row
// shifts), and which accumulator slot it feeds (the affinity kernel keeps three
// partial products apart).  Operands are f16, K-major, fetched by TMA with the
// 128-byte swizzle; accumulators are f32 in TMEM, double-buffered so that the
// epilogue of tile i overlaps the MMAs of tile i+1.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue.
// An epilogue warp may only touch the TMEM lane quarter warp_idx % 4, so two warps share each
// quarter and split the tile's columns between them (one epilogue warp per scheduler was
// latency-bound: a single warp cannot hide the TMEM-load -> math -> store dependency chain).
//
// Epilogues (template parameter EPI):
//   EPI_F32   raw f32 store (unit tests)
//   EPI_TDNN  speechbrain TDNNBlock tail: +bias -> ReLU -> BatchNorm(eval) [-> tanh], f16
//             channels-last store, with the reflect halo the next dilated conv needs
//             materialised, the Res2Net "x_{i+1} + y_i" sum emitted on the side, and
//             the ASP per-utterance context bias            (reference: speechbrain
//             ECAPA_TDNN, call sites speech_encode.py:77, ecapa_annote.py:22)
//   EPI_POOL  attentive-statistics pooling: rows = channels, columns = the frames of ONE
//             utterance, so softmax over time and the weighted mean/std are thread-local
//             (SURVEY App. A.2 AttentiveStatisticsPooling)
//   EPI_AFF   cosine distance 1 - S from split-f16 partial products
//             (diar_diag.py:215,219; anti_stick_diarize.py:176-177)
#pragma once
#include <cooperative_groups.h>
#include <cuda.h>
#include "sd_ptx.cuh"

namespace sd {

constexpr int BM = 128;        // UMMA M: rows of the A tile = TMEM lanes
constexpr int BK = 64;         // f16 elements per K chunk = one 128-byte swizzle span
constexpr int UMMA_K = 16;     // K per tcgen05.mma for 16-bit operands
constexpr int MAX_KITERS = 96; // 6144 / 64 (the per-utterance dense layers after pooling)
constexpr int GEMM_THREADS = 320;
constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;
#ifndef SD_CPRE
#define SD_CPRE 1   // 0: A/B build that fetches the epilogue constants at the top of each tile (round-1 behaviour)
#endif
// Epilogue warps per CTA, per epilogue kind.  The attentive-pooling epilogue (EPI_POOL) does ~14 instructions of
// softmax / moment arithmetic per frame-channel; POOL_EPI_WARPS = 16 (four warps per TMEM lane quarter, 576
// threads) was built and measured: the launch got SLOWER (0.27-0.29 -> 0.355 ms), so 8 stays.
constexpr int POOL_EPI_WARPS = 8;
__host__ __device__ constexpr int epi_warps_of(int epi) { return epi == 2 /*EPI_POOL*/ ? POOL_EPI_WARPS : EPI_WARPS; }
__host__ __device__ constexpr int gemm_threads_of(int epi) { return 64 + 32 * epi_warps_of(epi); }

enum { EPI_F32 = 0, EPI_TDNN = 1, EPI_POOL = 2, EPI_AFF = 3, EPI_ATT = 4, EPI_CONV3 = 5 };
// EPI_CONV3 = EPI_TDNN's epilogue behind a different operand path for the small dilated k=3 convolutions
// (Res2Net, 128 -> 128 channels): the tcgen05 unit applies the 128-byte swizzle from ABSOLUTE shared-memory
// address bits, so an A descriptor may start at any row of a TMA-loaded box (tools/micro/umma_rowoff_test.cu).
// One (128 + 2d)-row A box per 64-channel chunk therefore serves all three taps (row offsets 0, d, 2d),
// and the 96 KB of weights stay resident in shared memory for the whole launch: L2 -> SM traffic per tile
// drops from 192 KB to 34 KB.
__host__ __device__ constexpr bool epi_is_tdnn(int epi) { return epi == EPI_TDNN || epi == EPI_CONV3; }
enum {
  EF_REFLECT = 2,   // store interior rows only and mirror them into the halo rows
  EF_TMA_OUT = 4,   // cta_group::2 kernel, pointwise layers: the staged tile leaves through TMA tensor stores
                    // (GemmParams::tmapH = the output, tmapO2 = the out2 copy) instead of 16 LDS + 16 STG per thread
};
// EPI_TDNN : x = relu(acc + bias) * scale + shift                       (TDNNBlock tail)
// EPI_ATT  : x = tanh(relu(acc + bias + utt_bias[b]) * scale + shift)   (ASP attention TDNN)

struct KIter {
  int a_col;      // element column in A's tensor map
  int a_row_off;  // row shift applied to the A tile (conv tap * dilation)
  int b_col;      // element column in B's tensor map
  int slot;       // accumulator slot (0 unless EPI_AFF)
  int accum;      // 0 = first k-iteration of this slot (overwrite), 1 = accumulate
};

struct EpiParams {
  int flags;
  int M_rows;   // valid rows (A side) — rows >= M_rows are never stored
  int N_cols;   // valid columns (B side)
  // utterance geometry of the channels-last activation tensors:
  // rows [b*Tp, (b+1)*Tp), interior frame t at row b*Tp + H + t, 0 <= t < T.
  int Tp, T, H;
  // main output
  void* out;
  int ld_out;       // elements
  int out_col_off;  // elements
  const float* bias;   // [N] (EPI_TDNN) — ignored with EF_UTT_BIAS
  const float* scale;  // [N]
  const float* shift;  // [N]
  // second copy of the first `out2_cols` columns (Res2Net chunk 0 passes through)
  __half* out2;
  int ld_out2;
  int out2_cols;
  // Res2Net chain: sum_out[r, c] = y[r, c] + add_src[r, add_col_off + c]
  const __half* add_src;
  int ld_add;
  int add_col_off;
  __half* sum_out;
  int ld_sum;
  // ASP context bias [num_utts, N]
  const float* utt_bias;
  // EPI_POOL: h activations, global mean, pooled output [B, 2*C]
  const __half* h;
  int ld_h;
  const float* gmean;  // [B, ld_gmean], first C entries of a row = mean
  int ld_gmean;
  float* pooled;       // [B, 2*C]: mean then std
  __half* pooled_h;    // same, f16 (A operand of the final FC GEMM)
  int C;
  // EPI_AFF
  double* out_f64;  // optional second copy as f64 (AHC working matrix), same ld
  // EPI_TDNN, n_tile = 256, no EF_REFLECT: per-window column statistics fused into the write-out
  // (tdnn2 -> SE squeeze mean, MFA -> ASP mean/std).  The rows of a tile are cut into groups of
  // cs_group = gcd(128, Tp) rows; Tp is a multiple of 16, so a group never straddles two windows and always
  // covers the same window-relative rows.  colsum[(m_blk*(128/cs_group) + g)*N_cols + c] = sum over the interior
  // frames in group g of x - k[c], added in an order that depends only on the window-relative row; colsq the
  // same of (x - k[c])^2; k[c] = f16(shift[c]), the value a channel takes wherever its ReLU is off.
  float* colsum;
  float* colsq;
  // rows per statistics group = gcd(128, Tp) (16 .. 128): window-relative, so a window's partial sums are the
  // same numbers whichever batch slot (and therefore whichever 128-row tile split) it lands in
  int cs_group;
  // overflow flag (device int, may be null): set when an activation exceeded the f16 range and was saturated
  int* oflow;
};

struct alignas(64) GemmParams {
  CUtensorMap tmapA;
  CUtensorMap tmapB;
  CUtensorMap tmapH;  // EPI_POOL: the h activations, box = 64 channels x n_tile rows.  EF_TMA_OUT: the output
                      // columns [out_col_off, out_col_off + N_cols) x M_rows, box = 64 channels x 128 rows
  CUtensorMap tmapO2; // EF_TMA_OUT with epi.out2: the first out2_cols columns of out2, same box
  int num_m_blocks, num_n_blocks, num_kiters;
  int n_tile;      // UMMA N and rows of the B box
  int a_row_base;  // added to every A row coordinate (row-block sharding)
  int acc_slots;   // 1, or 3 for EPI_AFF
  // EPI_POOL on long utterances: a tile's columns (the frames of one utterance) are processed in
  // n_sub chunks of n_tile rows; B/h rows of chunk s start at n_blk * b_row_stride + s * n_tile.
  int n_sub;         // >= 1
  // tile order of the persistent loop (gemm_tc_kernel): 0 = n blocks fastest, 1 = m blocks fastest.
  // Measured on EPI_POOL (m fastest = all 24 channel blocks of an utterance run together): DRAM reads
  // 564 -> 531 MB but the launch got 3 % slower (0.266 -> 0.275 ms), so every caller leaves it at 0.
  int m_fastest;
  // gemm_tc_kernel (not the EPI_CONV3 path): walk the m blocks from the last to the first.  For a consumer whose
  // producer wrote A in ascending row order: the producer's last rows are still in L2 (126 MB) when it starts.
  int m_reverse;
  int b_row_stride;  // 0 = n_tile (dense tiling)
  // EPI_CONV3: taps, dilation (rows), padded input channels per tap in B's K axis; kit[kc].a_col = A column
  // of 64-channel chunk kc, num_kiters = number of chunks.  tmapA's box has 128 + (taps-1)*dil rows.
  int conv_taps, conv_dil, conv_cin;
  // split-K (EPI_F32 only): the k-iterations are divided among k_splits CTAs per output tile; split s
  // stores its partial sums at out + s * split_stride (elements) and a tiny follow-up kernel adds them in a
  // fixed order (deterministic, unlike atomics).  For the skinny per-utterance dense layers (M = batch,
  // K = 6144) that would otherwise occupy 4 SMs.
  int k_splits;  // >= 1
  long split_stride;
  uint32_t idesc;
  int a_prefetch;    // cta_group::2 kernel: k-iterations of A pulled into L2 ahead of the TMA loads (0 = off)
  long long* trace;  // debug (SD_GEMM_TRACE, cta_group::2 kernel): clock64 stamps of CTA 0, [tile < 32][16]
  KIter kit[MAX_KITERS];
  EpiParams epi;
};

template <int EPI, int MAX_BN>
struct GemmCfg {
  static constexpr bool CONV = (EPI == EPI_CONV3);
  static constexpr int A_BYTES = CONV ? 144 * BK * 2 : BM * BK * 2;  // 16 KB; CONV3: up to 128+16 rows = 18 KB
  static constexpr int B_BYTES = MAX_BN * BK * 2;     // 16 / 32 KB
  static constexpr int STAGE_BYTES = CONV ? A_BYTES : A_BYTES + B_BYTES;   // CONV3: B is resident, not staged
  static constexpr int BRES_BYTES = CONV ? 6 * B_BYTES : 0;                // 3 taps x 2 chunks x [128 x 64] f16
  // EPI_POOL has only 2 k-iterations per tile and needs room for the staged h tiles
  // EPI_TDNN/ATT stage the f16 output tile (and the Res2Net sum tile) in 64 KB of shared memory for
  // a coalesced write-out, which leaves room for 3 (MAX_BN 256) / 4 (MAX_BN 128) operand stages.
  // EPI_AFF stages its f32 tile ([128 rows][128 cols] = 64 KB) the same way.
  static constexpr bool STAGED_OUT = (epi_is_tdnn(EPI) || EPI == EPI_ATT || EPI == EPI_AFF);
  static constexpr int OUT_STAGE_BYTES = STAGED_OUT ? 65536 : 0;
  static constexpr int STAGES = CONV ? 3 : (EPI == EPI_POOL) ? 2 : (STAGED_OUT ? ((MAX_BN == 256) ? 3 : 4) : ((MAX_BN == 256) ? 4 : 6));
  static constexpr int EPI_SMEM_FLOATS = 3 * 256 + 1024;  // EPI_TDNN: bias/scale/shift of one n block + colsum exchange
  // EPI_POOL: two buffers of [2 chunks][n_tile rows][128 B] (n_tile <= 256 -> 64 KB each)
  // EPI_POOL additionally needs 2 KB to combine the two column halves of the softmax statistics
  static constexpr int EPI_REGION_BYTES = (EPI == EPI_POOL) ? 2 * 2 * MAX_BN * 128 + 2048 : EPI_SMEM_FLOATS * 4;
  static_assert(STAGE_BYTES % 1024 == 0, "stages must keep the 1024-byte swizzle alignment");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BRES_BYTES + OUT_STAGE_BYTES + EPI_REGION_BYTES + 256;
  static_assert(!CONV || MAX_BN == 128, "EPI_CONV3 is built for 128 output channels");
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can have");
};

// ------------------------------------------------------------------ epilogues
template <int THREADS = EPI_THREADS>
__device__ __forceinline__ void epi_named_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
}

// raw f32
__device__ __forceinline__ void epilogue_f32(const GemmParams& P, int m_blk, int n_blk, int ksplit,
                                             uint32_t tmem_acc, int quarter, int half, int lane) {
  const EpiParams& E = P.epi;
  const int r = m_blk * BM + quarter * 32 + lane;
  float* out = reinterpret_cast<float*>(E.out);
  for (int c0 = half * 16; c0 < P.n_tile; c0 += 32) {
    uint32_t v[16];
    __syncwarp();
    tmem_ld16(tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + c0, v);
    tmem_ld_wait();
    if (r < E.M_rows) {
      const int col0 = n_blk * P.n_tile + c0;
      float* dst = out + static_cast<size_t>(r) * E.ld_out + E.out_col_off + col0 + ksplit * P.split_stride;
      if (col0 + 16 <= E.N_cols && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        // the thread's 16 consecutive floats as four 16-byte stores: a warp's store instruction touches 32 rows either
        // way, but with 16 instead of 4 useful bytes per sector (the scalar form made the skinny split-K layers —
        // context bias, FC — spend most of their time in L2 store transactions)
        const bool wb = E.bias != nullptr && ksplit == 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                 __uint_as_float(v[4 * q + 3]));
          if (wb) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(E.bias + col0) + q);
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
          }
          reinterpret_cast<float4*>(dst)[q] = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col0 + j < E.N_cols)
            dst[j] = __uint_as_float(v[j]) + ((E.bias && ksplit == 0) ? __ldg(E.bias + col0 + j) : 0.f);
      }
    }
  }
}

// Res2Net chain: fetch this thread's x_{i+1} values (add_src) for its (at most two) 32-column
// chunks BEFORE waiting for the accumulator, so the global-load latency overlaps the MMAs.
__device__ __forceinline__ void tdnn_prefetch(const GemmParams& P, int m_blk, int n_blk, int quarter,
                                              int half, int lane, uint4 (&pre)[2][4]) {
  const EpiParams& E = P.epi;
  if (E.sum_out == nullptr) return;
  const int r = m_blk * BM + quarter * 32 + lane;
  if (r >= E.M_rows) return;
#pragma unroll
  for (int ci = 0; ci < 2; ++ci) {
    const int c0 = half * 32 + ci * 64;
    if (c0 < P.n_tile) {
      const uint4* a4 = reinterpret_cast<const uint4*>(
          E.add_src + static_cast<size_t>(r) * E.ld_add + E.add_col_off + n_blk * P.n_tile + c0);
#pragma unroll
      for (int q = 0; q < 4; ++q) pre[ci][q] = a4[q];  // plain (coherent) loads: in a chain, add_src was
                                                        // written by an earlier step of the same kernel
    }
  }
}

// TDNN tail. `sp` = this n block's {bias, scale, shift} staged in shared memory
// (read as float4: 3/4 of a shared load per element).  ATT adds the per-utterance context
// bias and the tanh of the attention TDNN; everything per-element is branch-free.
template <bool ATT>
__device__ __forceinline__ void epilogue_tdnn(const GemmParams& P, int m_blk, int n_blk,
                                              uint32_t tmem_acc, int quarter, int half, int lane,
                                              const float* sp, const uint4 (&pre)[2][4],
                                              uint8_t* stage_out) {
  const EpiParams& E = P.epi;
  const int rl = quarter * 32 + lane;  // row within the tile
  const int r = m_blk * BM + rl;
  const float* ub = ATT ? E.utt_bias + static_cast<size_t>(r < E.M_rows ? r / E.Tp : 0) * E.N_cols : nullptr;
  // staging tile: [64-column chunk][128 rows][128 B], 16-byte pieces XOR-swizzled by (row & 7) —
  // the layout a SWIZZLE_128B tensor map expects, and conflict-free for one-row-per-lane writes.
  uint8_t* srow = stage_out + rl * 128;
  const int sw = rl & 7;
  uint8_t* sum_stage = stage_out + 32768;  // Res2Net sum tile (n_tile = 128: two 16 KB chunks each)
  float amax = 0.f;                        // largest |activation| this thread stores (overflow flag)
  for (int c0 = half * 32; c0 < P.n_tile; c0 += 64) {
    uint32_t v[32];
    __syncwarp();
    tmem_ld32(tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + c0, v);
    tmem_ld_wait();
#if SD_EXPERIMENTS
    if (E.flags & 64) continue;  // TIMING PROBE ONLY (SD_DEBUG_EPI=1): skip the epilogue math and stores
#endif
    const int col0 = n_blk * P.n_tile + c0;  // column within this layer's output
    float x[32];
    const float4* b4 = reinterpret_cast<const float4*>(sp + c0);
    const float4* s4 = reinterpret_cast<const float4*>(sp + 256 + c0);
    const float4* h4 = reinterpret_cast<const float4*>(sp + 512 + c0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 bb = b4[q];
      const float4 ss = s4[q], hh = h4[q];
      if (ATT) {
        const float4 uu = __ldg(reinterpret_cast<const float4*>(ub + col0) + q);
        bb.x += uu.x; bb.y += uu.y; bb.z += uu.z; bb.w += uu.w;
      }
      x[4 * q + 0] = fmaf(fmaxf(__uint_as_float(v[4 * q + 0]) + bb.x, 0.f), ss.x, hh.x);
      x[4 * q + 1] = fmaf(fmaxf(__uint_as_float(v[4 * q + 1]) + bb.y, 0.f), ss.y, hh.y);
      x[4 * q + 2] = fmaf(fmaxf(__uint_as_float(v[4 * q + 2]) + bb.z, 0.f), ss.z, hh.z);
      x[4 * q + 3] = fmaf(fmaxf(__uint_as_float(v[4 * q + 3]) + bb.w, 0.f), ss.w, hh.w);
    }
    if (ATT) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = tanhf(x[j]);
    } else if (kTrackOflow) {
#pragma unroll
      for (int j = 0; j < 32; ++j) amax = fmaxf(amax, fabsf(x[j]));
    }
    uint4 pk[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      pk[q].x = pack_half2(x[8 * q + 0], x[8 * q + 1]);
      pk[q].y = pack_half2(x[8 * q + 2], x[8 * q + 3]);
      pk[q].z = pack_half2(x[8 * q + 4], x[8 * q + 5]);
      pk[q].w = pack_half2(x[8 * q + 6], x[8 * q + 7]);
    }
#if SD_EXPERIMENTS
    if (E.flags & 128) continue;  // TIMING PROBE ONLY (SD_DEBUG_EPI=2): math but no stores
#endif
    // -> staging (every row, valid or not: tdnn_writeout skips what must not be written)
    const int chunk = c0 >> 6;           // 64-column chunk of the tile
    const int p0 = (c0 >> 5 & 1) * 4;    // first 16-byte piece of this 32-column run inside the 128 B row
    {
      uint8_t* d = srow + chunk * 16384;
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(d + (((p0 + q) ^ sw) << 4)) = pk[q];
    }
    if (E.sum_out != nullptr) {
      // next Res2Net input: x_{i+1} + y_i, built from the f16-rounded y_i the next conv
      // would otherwise have read back (keeps the sum bit-identical to an unfused chain).
      // the x_{i+1} values were prefetched before the accumulator was waited for (tdnn_prefetch)
      const int ci = (c0 >> 6) & 1;
      uint4 sk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 av = ci ? pre[1][q] : pre[0][q];
        const __half2* ah = reinterpret_cast<const __half2*>(&av);
        const __half2* yh = reinterpret_cast<const __half2*>(&pk[q]);
        __half2 sh[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 fa = __half22float2(ah[e]);
          float2 fy = __half22float2(yh[e]);
          // (not part of the overflow tracking: rows beyond M_rows carry uninitialised x_{i+1} registers here; a
          // sum that saturates shows up in the next convolution's accumulators, which are tracked)
          const uint32_t pks = pack_half2(fa.x + fy.x, fa.y + fy.y);
          sh[e] = *reinterpret_cast<const __half2*>(&pks);
        }
        sk[q] = *reinterpret_cast<uint4*>(sh);
      }
      {
        uint8_t* d = sum_stage + rl * 128 + chunk * 16384;
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(d + (((p0 + q) ^ sw) << 4)) = sk[q];
      }
    }
  }
  // rows beyond M_rows hold whatever the zero-filled operand rows produced (bias / shift only): finite
  if (kTrackOflow && !ATT && amax > kHalfMax && E.oflow != nullptr) atomicOr(E.oflow, 1);
}

// Write-out of the staged tile by all 256 epilogue threads: 8 lanes cover one 128-byte row chunk,
// so every store instruction writes four complete 128-byte lines (the direct one-row-per-thread
// stores issued 16-byte pieces of 32 different lines per instruction and cost ~25% of the whole
// forward in L2 transactions).  Halo / padding rows are skipped per row and the reflect mirrors
// are written from the same registers.  (TMA tensor stores were tried first: they clip at the
// upper tensor bound but FAULT on negative start coordinates, which the halo clipping needs —
// tools/micro/tma_store_test.cu.)
__device__ __forceinline__ void tdnn_writeout(const GemmParams& P, int m_blk, int n_blk,
                                              const uint8_t* stage_out, int et) {
  const EpiParams& E = P.epi;
#if SD_EXPERIMENTS
  if (E.flags & (64 | 128 | 256)) return;  // TIMING/DEBUG PROBES (SD_DEBUG_EPI)
#endif
  const int sub = et & 7;    // 16-byte piece of the 128-byte row chunk
  const int rsub = et >> 3;  // row within a 32-row pass
  const int nchunks = P.n_tile >> 6;
  __half* out = reinterpret_cast<__half*>(E.out);
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int rl = pass * 32 + rsub;
    const int r = m_blk * BM + rl;
    if (r >= E.M_rows) continue;
    bool valid = true;
    int r2 = -1, r3 = -1;
    if (E.flags & EF_REFLECT) {
      const int b = r / E.Tp;
      const int t = r - b * E.Tp - E.H;
      valid = t >= 0 && t < E.T;
      if (valid) {
        if (t >= 1 && t <= E.H) r2 = b * E.Tp + E.H - t;
        if (t >= E.T - 1 - E.H && t <= E.T - 2) r3 = b * E.Tp + E.H + 2 * (E.T - 1) - t;
      }
    }
    if (!valid) continue;
    const uint8_t* srow = stage_out + rl * 128 + ((sub ^ (rl & 7)) << 4);
    uint4 vals[4];   // the row's (at most four) 16-byte pieces first, then the stores: four shared loads in flight
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nchunks) vals[j] = *reinterpret_cast<const uint4*>(srow + j * 16384);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j >= nchunks) break;
      const int col = n_blk * P.n_tile + j * 64 + sub * 8;
      const uint4 val = vals[j];
      const size_t coff = static_cast<size_t>(E.out_col_off + col);
      *reinterpret_cast<uint4*>(out + static_cast<size_t>(r) * E.ld_out + coff) = val;
      if (r2 >= 0) *reinterpret_cast<uint4*>(out + static_cast<size_t>(r2) * E.ld_out + coff) = val;
      if (r3 >= 0) *reinterpret_cast<uint4*>(out + static_cast<size_t>(r3) * E.ld_out + coff) = val;
      if (E.out2 != nullptr && col < E.out2_cols)
        *reinterpret_cast<uint4*>(E.out2 + static_cast<size_t>(r) * E.ld_out2 + col) = val;
      if (E.sum_out != nullptr) {
        const uint4 sv = *reinterpret_cast<const uint4*>(srow + 32768 + j * 16384);
        *reinterpret_cast<uint4*>(E.sum_out + static_cast<size_t>(r) * E.ld_sum + col) = sv;
        if (r2 >= 0) *reinterpret_cast<uint4*>(E.sum_out + static_cast<size_t>(r2) * E.ld_sum + col) = sv;
        if (r3 >= 0) *reinterpret_cast<uint4*>(E.sum_out + static_cast<size_t>(r3) * E.ld_sum + col) = sv;
      }
    }
  }
}

// Write-out + per-window column statistics for 256-wide tiles (see EpiParams::colsum).  Warp `we` owns the
// 64-column chunk we>>1 and rows (we&1)*64 .. +63: a store instruction still covers four complete 128-byte
// lines, and a column's partial sum lives in one warp (shuffle over its 4 row lanes).  A group of
// G = cs_group rows is summed as: each lane its rows (every fourth) in ascending order, then the butterfly over
// the four row lanes, then (G = 128 only) the two warps of a chunk through `part` in a fixed order — an order
// that depends on the row's position inside its group only, hence not on the window's batch slot.
template <bool STORE = true>   // STORE = false: statistics only, the tile itself leaves through TMA (EF_TMA_OUT)
__device__ __forceinline__ void tdnn_writeout_colsum(const GemmParams& P, int m_blk, int n_blk,
                                                     const uint8_t* stage_out, float* part, int et) {
  const EpiParams& E = P.epi;
  const int we = et >> 5, lane = et & 31;
  const int j = we >> 1, hrow = (we & 1) * 64;
  const int sub = lane & 7, rlo = lane >> 3;
  const int col = n_blk * P.n_tile + j * 64 + sub * 8;  // first of this thread's 8 channels
  const bool want_sq = E.colsq != nullptr;
  float k[8];
#pragma unroll
  for (int e = 0; e < 8; ++e)
    k[e] = (E.shift != nullptr && col + e < E.N_cols) ? __half2float(__float2half_rn(__ldg(E.shift + col + e))) : 0.f;
  const int G = E.cs_group;               // 16, 32, 64 or 128
  const int GW = G < 64 ? G : 64;         // rows between two flushes of this warp's accumulators
  const int gpt = BM / G;                 // groups per tile
  const int row0 = m_blk * BM;
  int p = (row0 + hrow + rlo) % E.Tp;     // window-relative row of this lane's next row
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  __half* out = reinterpret_cast<__half*>(E.out);
  const uint8_t* sbase = stage_out + j * 16384;
  const bool col_ok = col < E.N_cols;
  for (int quad = 0; quad < 4; ++quad) {
#pragma unroll
    for (int pi = 0; pi < 4; ++pi) {
      const int rl = hrow + (quad * 4 + pi) * 4 + rlo;
      const int r = row0 + rl;
      const int t = p - E.H;
      p += 4;
      if (p >= E.Tp) p -= E.Tp;
      if (r >= E.M_rows) continue;
      const uint4 val = *reinterpret_cast<const uint4*>(sbase + rl * 128 + ((sub ^ (rl & 7)) << 4));
      if (STORE) *reinterpret_cast<uint4*>(out + static_cast<size_t>(r) * E.ld_out + E.out_col_off + col) = val;
      if (t < 0 || t >= E.T) continue;
      const __half2* vh = reinterpret_cast<const __half2*>(&val);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(vh[e]);
        const float d0 = f.x - k[2 * e], d1 = f.y - k[2 * e + 1];
        s[2 * e] += d0;
        s[2 * e + 1] += d1;
        if (want_sq) {
          q[2 * e] = fmaf(d0, d0, q[2 * e]);
          q[2 * e + 1] = fmaf(d1, d1, q[2 * e + 1]);
        }
      }
    }
    if (((quad + 1) * 16) % GW != 0) continue;   // warp-uniform: the group (or this warp's half of it) is complete
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s[e] += __shfl_xor_sync(0xffffffffu, s[e], 8);
      s[e] += __shfl_xor_sync(0xffffffffu, s[e], 16);
      if (want_sq) {
        q[e] += __shfl_xor_sync(0xffffffffu, q[e], 8);
        q[e] += __shfl_xor_sync(0xffffffffu, q[e], 16);
      }
    }
    if (G < 128) {
      if (rlo == 0 && col_ok) {
        const int g = (hrow + quad * 16) / G;
        float* g0 = E.colsum + (static_cast<size_t>(m_blk) * gpt + g) * E.N_cols + col;
        *reinterpret_cast<float4*>(g0) = make_float4(s[0], s[1], s[2], s[3]);
        *reinterpret_cast<float4*>(g0 + 4) = make_float4(s[4], s[5], s[6], s[7]);
        if (want_sq) {
          float* h0 = E.colsq + (static_cast<size_t>(m_blk) * gpt + g) * E.N_cols + col;
          *reinterpret_cast<float4*>(h0) = make_float4(q[0], q[1], q[2], q[3]);
          *reinterpret_cast<float4*>(h0 + 4) = make_float4(q[4], q[5], q[6], q[7]);
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
    }
  }
  if (G < 128) return;
  // G = 128: the chunk's two warps (rows 0-63, 64-127) combine through shared memory, lower half first
  float* ps = part + j * 64 + sub * 8;   // part: [quantity s|q][chunk j][64 columns]
  float* pq = ps + 256;
  if ((we & 1) && rlo == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      ps[e] = s[e];
      if (want_sq) pq[e] = q[e];
    }
  }
  epi_named_barrier();
  if (!(we & 1) && rlo == 0 && col_ok) {
    float* g0 = E.colsum + static_cast<size_t>(m_blk) * E.N_cols + col;
#pragma unroll
    for (int e = 0; e < 8; ++e) g0[e] = s[e] + ps[e];
    if (want_sq) {
      float* h0 = E.colsq + static_cast<size_t>(m_blk) * E.N_cols + col;
#pragma unroll
      for (int e = 0; e < 8; ++e) h0[e] = q[e] + pq[e];
    }
  }
}

// Attentive statistics pooling. Tile rows = 128 channels, tile columns = the Tp
// rows of utterance n_blk, so each thread owns one channel's logits over time.
// The conv bias of the attention's output layer is constant over time and
// cancels in the softmax, so it is never added.  The matching h tile
// ([Tp rows][128 channels] f16) was staged in shared memory by TMA (128-byte
// swizzle: 16-byte chunk index ^= row & 7), so the 151 per-thread reads of h are
// shared-memory reads instead of latency-bound 2-byte global loads.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// running softmax statistics of one channel (one thread), carried across the time chunks of a tile
struct PoolState {
  float mx = -INFINITY, se = 0.f, s1 = 0.f, s2 = 0.f;
};

// Register-resident form of the two softmax passes for utterances of <= NCH*32 padded frames handled in
// one time chunk: the warp's NCH 16-column runs of logits are read from TMEM once (all loads in flight
// together, one wait), frames outside [0, T) are set to -inf so both passes are branch-free, and the
// three running sums use two independent accumulators each.  The generic path below (two TMEM reads,
// one wait per 16 columns, per-element edge tests) was issue- and latency-bound: 17 instructions and
// ~70 cycles per element per warp in ncu (profiles/r01_ncu_head.txt).
template <int NCH, int NP>   // NP = epilogue warps per lane quarter; this warp owns runs part, part + NP, ...
__device__ __forceinline__ void pool_chunks_regs(const GemmParams& P, uint32_t tbase, int half,
                                                 const uint8_t* hchunk, int c16, float g, PoolState& st) {
  const EpiParams& E = P.epi;
  const float LOG2E = 1.4426950408889634f;
  uint32_t v[NCH][16];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c0 = (NP * i + half) * 16;
    if (c0 < P.n_tile) tmem_ld16(tbase + c0, v[i]);
  }
  tmem_ld_wait();
  float cm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int t0 = (NP * i + half) * 16 - E.H;  // frame index of the run's first column
    if (!(t0 >= 0 && t0 + 16 <= E.T)) {        // run touches the halo / padding (warp-uniform)
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (t0 + j < 0 || t0 + j >= E.T) v[i][j] = 0xff800000u;  // -inf: ignored by max, exp2 -> 0
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) cm[j & 3] = fmaxf(cm[j & 3], __uint_as_float(v[i][j]));
  }
  const float mx = fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3]));
  const float mxs = (mx == -INFINITY) ? 0.f : mx * LOG2E;
  float2 se2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, s12[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)},
         s22[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c0 = (NP * i + half) * 16;
    const int t0 = c0 - E.H;
    if (c0 < P.n_tile) {
      float xv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int p = c0 + j;
        xv[j] = __half2float(*reinterpret_cast<const __half*>(hchunk + p * 128 + ((c16 ^ (p & 7)) << 4))) - g;
      }
      if (!(t0 >= 0 && t0 + 16 <= E.T)) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (t0 + j < 0 || t0 + j >= E.T) xv[j] = 0.f;  // weight is exactly 0; keep 0 * garbage out
      }
      // two frames per instruction (sm_100's packed f32x2 FMA / ADD / MUL): the same IEEE operations in the same
      // order as the scalar loop — accumulator j & 3 is lane j & 1 of pair (j >> 1) & 1 — at half the issue slots
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float2 arg = fma2(make_float2(__uint_as_float(v[i][j]), __uint_as_float(v[i][j + 1])),
                                make_float2(LOG2E, LOG2E), make_float2(-mxs, -mxs));
        const float2 e = make_float2(ex2_approx(arg.x), ex2_approx(arg.y));
        const float2 x2 = make_float2(xv[j], xv[j + 1]);
        const int a = (j >> 1) & 1;
        se2[a] = add2(se2[a], e);
        const float2 ex = mul2(e, x2);
        s12[a] = add2(s12[a], ex);
        s22[a] = fma2(ex, x2, s22[a]);
      }
    }
  }
  st.mx = mx;
  st.se = (se2[0].x + se2[0].y) + (se2[1].x + se2[1].y);
  st.s1 = (s12[0].x + s12[0].y) + (s12[1].x + s12[1].y);
  st.s2 = (s22[0].x + s22[0].y) + (s22[1].x + s22[1].y);
}

__device__ __forceinline__ void epilogue_pool(const GemmParams& P, int m_blk, int n_blk, int sub,
                                              uint32_t tmem_acc, int quarter, int half, int lane,
                                              uint8_t* hbuf, PoolState& st, float g) {
  constexpr int NP = POOL_EPI_WARPS / 4;          // `half` is the warp's part index 0 .. NP-1
  constexpr int PT = POOL_EPI_WARPS * 32;
  const EpiParams& E = P.epi;
  const int chl = quarter * 32 + lane;  // channel within the tile
  const int ch = m_blk * BM + chl;
  const int b = n_blk;
  const uint32_t tbase = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16);
  const bool chv = ch < E.C;
  const float LOG2E = 1.4426950408889634f;
  const int f0 = sub * P.n_tile - E.H;  // frame index of this chunk's column 0
  const uint8_t* hchunk = hbuf + (chl >> 6) * (P.n_tile * 128) + (chl & 7) * 2;
  const int c16 = (chl & 63) >> 3;
  float se, s1, s2;
  constexpr int NCH = NP == 2 ? 5 : 3;            // 16-column runs a warp keeps in registers (windows <= 1.5 s)
  if (P.n_sub == 1 && P.n_tile <= NP * NCH * 16) {
    pool_chunks_regs<NCH, NP>(P, tbase, half, hchunk, c16, g, st);
    se = st.se; s1 = st.s1; s2 = st.s2;
  } else {
    // this warp's share of the frames: 16-column chunks  half, half+2, ...
    // pass 1: max over this chunk's interior frames
    float cm = -INFINITY;
    for (int c0 = half * 16; c0 < P.n_tile; c0 += NP * 16) {
      uint32_t v[16];
      tmem_ld16(tbase + c0, v);
      tmem_ld_wait();
      if (f0 + c0 >= 0 && f0 + c0 + 16 <= E.T) {
#pragma unroll
        for (int j = 0; j < 16; ++j) cm = fmaxf(cm, __uint_as_float(v[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int t = f0 + c0 + j;
          if (t >= 0 && t < E.T) cm = fmaxf(cm, __uint_as_float(v[j]));
        }
      }
    }
    // online softmax: rescale what earlier chunks accumulated to the new running maximum
    if (cm > st.mx) {
      const float f = (st.mx == -INFINITY) ? 0.f : ex2_approx((st.mx - cm) * LOG2E);
      st.se *= f;
      st.s1 *= f;
      st.s2 *= f;
      st.mx = cm;
    }
    // pass 2: softmax-weighted first/second moments about the global mean g
    const float mxs = (st.mx == -INFINITY) ? 0.f : st.mx * LOG2E;  // nothing interior seen yet
    se = st.se; s1 = st.s1; s2 = st.s2;
    for (int c0 = half * 16; c0 < P.n_tile; c0 += NP * 16) {
      uint32_t v[16];
      tmem_ld16(tbase + c0, v);
      float xv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int p = c0 + j;
        xv[j] = __half2float(*reinterpret_cast<const __half*>(hchunk + p * 128 + ((c16 ^ (p & 7)) << 4))) - g;
      }
      tmem_ld_wait();
      if (f0 + c0 >= 0 && f0 + c0 + 16 <= E.T) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float e = ex2_approx(fmaf(__uint_as_float(v[j]), LOG2E, -mxs));
          se += e;
          const float ex = e * xv[j];
          s1 += ex;
          s2 = fmaf(ex, xv[j], s2);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int t = f0 + c0 + j;
          if (t >= 0 && t < E.T) {
            const float e = ex2_approx(fmaf(__uint_as_float(v[j]), LOG2E, -mxs));
            se += e;
            const float ex = e * xv[j];
            s1 += ex;
            s2 = fmaf(ex, xv[j], s2);
          }
        }
      }
    }
  }
  st.se = se;
  st.s1 = s1;
  st.s2 = s2;
  if (sub + 1 < P.n_sub) return;  // more time chunks of this utterance follow
  // last chunk: combine the NP column parts (parts 1.. -> shared -> part 0) and finish.  The exchange area is
  // the h tile this epilogue has just finished reading: the producer cannot refill it before all the
  // epilogue warps have arrived on its "empty" barrier, which part 0 does only after reading the exchange.
  float4* xchg = reinterpret_cast<float4*>(hbuf);
  epi_named_barrier<PT>();            // every warp is done with the h tile
  if (half > 0) xchg[(half - 1) * BM + chl] = make_float4(st.mx, se, s1, s2);
  epi_named_barrier<PT>();
  if (half == 0) {
    float M = st.mx;
#pragma unroll
    for (int q = 0; q < NP - 1; ++q) M = fmaxf(M, xchg[q * BM + chl].x);
    const float fa = (st.mx == -INFINITY) ? 0.f : ex2_approx((st.mx - M) * LOG2E);
    se *= fa;
    s1 *= fa;
    s2 *= fa;
#pragma unroll
    for (int q = 0; q < NP - 1; ++q) {
      const float4 o = xchg[q * BM + chl];
      const float fb = (o.x == -INFINITY) ? 0.f : ex2_approx((o.x - M) * LOG2E);
      se = fmaf(o.y, fb, se);
      s1 = fmaf(o.z, fb, s1);
      s2 = fmaf(o.w, fb, s2);
    }
    if (chv) {
      const float inv = 1.f / se;
      const float m1 = s1 * inv;
      const float mean = g + m1;
      const float sd = sqrtf(fmaxf(s2 * inv - m1 * m1, 1e-12f));
      const size_t oo = static_cast<size_t>(b) * 2 * E.C + ch;
      E.pooled[oo] = mean;
      E.pooled[oo + E.C] = sd;
      if (E.pooled_h) {
        E.pooled_h[oo] = half_sat(mean);
        E.pooled_h[oo + E.C] = half_sat(sd);
      }
    }
  }
  // generic-proxy accesses to the h buffer are ordered before the TMA refill that follows the empty barrier
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  st = PoolState();
}

// Cosine distance from the three split-f16 partial products:
//   slot0 = hi.hi   slot1 = hi.lo'   slot2 = lo'.hi      (lo' = lo * 2^11)
//   S = slot0 + (slot1 + slot2) * 2^-11 ,  D = 1 - S      (f32, as sklearn computes it)
// slot1 + slot2 is commutative, so D is exactly symmetric.
__device__ __forceinline__ void epilogue_aff(const GemmParams& P, int m_blk, int n_blk,
                                             uint32_t tmem_acc, int quarter, int half, int lane,
                                             uint8_t* stage_out) {
  const EpiParams& E = P.epi;
  const int rl = quarter * 32 + lane;
  const int r = m_blk * BM + rl;
  const uint32_t tbase = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16);
  // staging: [128 rows][512 B], 16-byte pieces XOR-swizzled by (row & 7) -> conflict-free float4 writes
  uint8_t* srow = stage_out + rl * 512;
  const int sw = rl & 7;
  for (int c0 = half * 16; c0 < P.n_tile; c0 += 32) {
    uint32_t v0[16], v1[16], v2[16];
    __syncwarp();
    tmem_ld16(tbase + c0, v0);
    tmem_ld16(tbase + P.n_tile + c0, v1);
    tmem_ld16(tbase + 2 * P.n_tile + c0, v2);
    tmem_ld_wait();
    float d[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float s = __uint_as_float(v0[j]) +
                      (__uint_as_float(v1[j]) + __uint_as_float(v2[j])) * 4.8828125e-4f;
      d[j] = 1.0f - s;
    }
    const int p0 = c0 >> 2;  // first 16-byte piece of this 16-column run
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<float4*>(srow + (((p0 + q) ^ sw) << 4)) =
          make_float4(d[4 * q], d[4 * q + 1], d[4 * q + 2], d[4 * q + 3]);
    if (E.out_f64 != nullptr) {
      const int col0 = n_blk * P.n_tile + c0;
      if (r < E.M_rows) {
        const size_t off = static_cast<size_t>(r) * E.ld_out + col0;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col0 + j < E.N_cols) E.out_f64[off + j] = static_cast<double>(d[j]);
      }
    }
  }
}

// Coalesced write-out of the staged f32 distance tile: one warp instruction = one row's 512 bytes.
__device__ __forceinline__ void aff_writeout(const GemmParams& P, int m_blk, int n_blk,
                                             const uint8_t* stage_out, int et) {
  const EpiParams& E = P.epi;
  const int lane = et & 31, w = et >> 5;  // 8 warps x 16 rows
  float* out = reinterpret_cast<float*>(E.out);
  const int col = n_blk * P.n_tile + lane * 4;
  const bool vec_ok = (E.ld_out & 3) == 0;
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int rl = w * 16 + i;
    const int r = m_blk * BM + rl;
    if (r >= E.M_rows) break;
    const float4 val = *reinterpret_cast<const float4*>(stage_out + rl * 512 + ((lane ^ (rl & 7)) << 4));
    float* dst = out + static_cast<size_t>(r) * E.ld_out + col;
    if (vec_ok && col + 4 <= E.N_cols) {
      *reinterpret_cast<float4*>(dst) = val;
    } else {
      if (col < E.N_cols) dst[0] = val.x;
      if (col + 1 < E.N_cols) dst[1] = val.y;
      if (col + 2 < E.N_cols) dst[2] = val.z;
      if (col + 3 < E.N_cols) dst[3] = val.w;
    }
  }
}

// --------------------------------------------------------------------- kernel
struct GemmCtx {
  uint8_t* smem;
  uint8_t* epi_region;
  uint64_t *full_bar, *empty_bar, *tfull_bar, *tempty_bar, *hfull_bar, *hempty_bar;
  uint32_t tmem_base;
  int warp, lane;
};

// Ring / accumulator positions.  Each role keeps its own copy; they stay in step because every
// role walks the same tile and k-iteration sequence.  They persist across the steps of a chain.
struct PipeState {
  int stage = 0;
  uint32_t phase = 0;
  int as = 0;
  uint32_t aphase = 0;
  int hs = 0;
  uint32_t hphase = 0;
  uint32_t bphase = 0;  // EPI_CONV3: phase of the resident-weights barrier (one use per GEMM)
};

template <int EPI, int MAX_BN, bool MC = false>
__device__ __forceinline__ void gemm_setup(uint8_t* smem, GemmCtx& c) {
  using Cfg = GemmCfg<EPI, MAX_BN>;
  // 1024-byte alignment is required by the 128-byte swizzle atoms.  It is declared on the array
  // (instead of rounding the pointer up by hand) so every derived pointer stays in the shared
  // address space and the epilogues' reads compile to LDS rather than generic loads.
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  c.smem = smem;
  c.epi_region = smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::BRES_BYTES + Cfg::OUT_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(c.epi_region + Cfg::EPI_REGION_BYTES);
  c.full_bar = bars;                        // [STAGES]
  c.empty_bar = bars + Cfg::STAGES;         // [STAGES]
  c.tfull_bar = bars + 2 * Cfg::STAGES;     // [2]
  c.tempty_bar = c.tfull_bar + 2;           // [2]
  c.hfull_bar = c.tempty_bar + 2;           // [3]  EPI_POOL: staged h tile landed
  c.hempty_bar = c.hfull_bar + 3;           // [3]  EPI_POOL: epilogue done with it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c.hempty_bar + 3);
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  if (c.warp == 0) {
    if (c.lane == 0) {
      for (int s = 0; s < Cfg::STAGES; ++s) {
        mbar_init(&c.full_bar[s], 1);
        mbar_init(&c.empty_bar[s], MC ? 2 : 1);  // MC: both CTAs of the pair must release a stage
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&c.tfull_bar[s], 1);
        mbar_init(&c.tempty_bar[s], epi_warps_of(EPI));  // one arrive per epilogue warp
      }
      for (int s = 0; s < 3; ++s) {
        mbar_init(&c.hfull_bar[s], 1);
        mbar_init(&c.hempty_bar[s], epi_warps_of(EPI));
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();  // the peer's barriers exist before anything is multicast at them
  tc_fence_after();
  c.tmem_base = *tmem_slot;
}

template <bool MC = false>
__device__ __forceinline__ void gemm_teardown(const GemmCtx& c) {
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();  // the peer may still multicast into this CTA's smem / arrive on its barriers
  if (c.warp == 0) {
    tc_fence_after();
    tmem_dealloc(c.tmem_base, TMEM_COLS);
  }
}

// One GEMM: every role walks this CTA's tiles (tile = blockIdx.x, + gridDim.x, ...).
// MC = 2-CTA cluster with TMA multicast of the B (weight) tile: the two CTAs of a pair work on adjacent
// m blocks of the same n block in lockstep; each loads HALF of the B tile and multicasts it to both, so a
// CTA pulls 32 KB instead of 48 KB per k-iteration through L2 (the large GEMMs are L2 -> SM bound).
template <int EPI, int MAX_BN, bool MC = false>
__device__ __forceinline__ void gemm_run(const GemmParams& P, const GemmCtx& c, PipeState& ps) {
  using Cfg = GemmCfg<EPI, MAX_BN>;
  uint8_t* const smem = c.smem;
  uint8_t* const epi_region = c.epi_region;
  uint8_t* const bres = smem + Cfg::STAGES * Cfg::STAGE_BYTES;        // EPI_CONV3: resident weights (96 KB)
  uint8_t* const stage_out = bres + Cfg::BRES_BYTES;                  // EPI_TDNN/ATT/AFF output staging (64 KB)
  float* const epi_sp = reinterpret_cast<float*>(epi_region);  // EPI_TDNN: per-column constants
  const uint32_t hbuf_bytes = 2u * static_cast<uint32_t>(P.n_tile) * 128u;  // EPI_POOL: one h tile
  // h tiles in flight: three when they fit the region sized for two 256-frame tiles (windows <= 1.6 s), so the
  // tile after next is already streaming from HBM while the epilogue works
  const int h_bufs = (EPI == EPI_POOL && 3u * hbuf_bytes <= 2u * 2u * MAX_BN * 128u) ? 3 : 2;
  const int warp = c.warp, lane = c.lane;
  const int k_splits = P.k_splits < 1 ? 1 : P.k_splits;
  const int k_per = (P.num_kiters + k_splits - 1) / k_splits;
  const int crank = MC ? static_cast<int>(cluster_ctarank()) : 0;
  const int m_units = MC ? (P.num_m_blocks + 1) / 2 : P.num_m_blocks;      // MC: pairs of m blocks
  const int num_tiles = m_units * P.num_n_blocks * k_splits;               // split index fastest
  const int tile0 = MC ? (blockIdx.x >> 1) : blockIdx.x;
  const int tstep = MC ? (gridDim.x >> 1) : gridDim.x;
  const int acc_stages = (P.acc_slots * P.n_tile <= 256) ? 2 : 1;
  const uint32_t tmem_base = c.tmem_base;
  const int n_sub = P.n_sub < 1 ? 1 : P.n_sub;
  const int b_row_stride = P.b_row_stride ? P.b_row_stride : P.n_tile;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&P.tmapA);
      tma_prefetch_desc(&P.tmapB);
      if (EPI == EPI_POOL) tma_prefetch_desc(&P.tmapH);
      const uint32_t tx = Cfg::A_BYTES + static_cast<uint32_t>(P.n_tile) * BK * 2;
      if (EPI == EPI_CONV3) {
        // weights: taps x chunks boxes of [n_tile x 64], loaded once and kept for every tile
        const uint32_t bbox = static_cast<uint32_t>(P.n_tile) * BK * 2;
        mbar_arrive_expect_tx(&c.hfull_bar[0], bbox * P.conv_taps * P.num_kiters);
        for (int kc = 0; kc < P.num_kiters; ++kc)
          for (int j = 0; j < P.conv_taps; ++j)
            tma_load_2d(bres + (kc * P.conv_taps + j) * Cfg::B_BYTES, &P.tmapB, &c.hfull_bar[0],
                        j * P.conv_cin + kc * BK, 0);
        const int box_rows = BM + (P.conv_taps - 1) * P.conv_dil;
        const uint32_t txa = static_cast<uint32_t>(box_rows) * BK * 2;
        const int row_shift = -(P.conv_taps / 2) * P.conv_dil;
        for (int tile = tile0; tile < num_tiles; tile += tstep) {
          const int m_blk = tile / P.num_n_blocks;
          for (int kc = 0; kc < P.num_kiters; ++kc) {
            mbar_wait(&c.empty_bar[ps.stage], ps.phase ^ 1);
            mbar_arrive_expect_tx(&c.full_bar[ps.stage], txa);
            tma_load_2d(smem + ps.stage * Cfg::STAGE_BYTES, &P.tmapA, &c.full_bar[ps.stage], P.kit[kc].a_col,
                        P.a_row_base + m_blk * BM + row_shift);
            if (++ps.stage == Cfg::STAGES) { ps.stage = 0; ps.phase ^= 1; }
          }
        }
      } else
      for (int tile = tile0; tile < num_tiles; tile += tstep) {
        const int ot = tile / k_splits, ks = tile - ot * k_splits;
        const int m_unit0 = P.m_fastest ? ot % m_units : ot / P.num_n_blocks;
        const int n_blk = P.m_fastest ? ot / m_units : ot - m_unit0 * P.num_n_blocks;
        const int m_unit = P.m_reverse ? m_units - 1 - m_unit0 : m_unit0;
        const int m_blk = MC ? 2 * m_unit + crank : m_unit;
        const int k_begin = ks * k_per, k_end = min(P.num_kiters, k_begin + k_per);
        for (int sub = 0; sub < n_sub; ++sub) {
          const int b_row = n_blk * b_row_stride + sub * P.n_tile;
          if (EPI == EPI_POOL) {
            // the h tile this chunk's epilogue will weight: 128 channels x n_tile rows, two 64-channel boxes
            mbar_wait(&c.hempty_bar[ps.hs], ps.hphase ^ 1);
            mbar_arrive_expect_tx(&c.hfull_bar[ps.hs], hbuf_bytes);
            uint8_t* hb = epi_region + ps.hs * hbuf_bytes;
            tma_load_2d(hb, &P.tmapH, &c.hfull_bar[ps.hs], m_blk * BM, b_row);
            tma_load_2d(hb + hbuf_bytes / 2, &P.tmapH, &c.hfull_bar[ps.hs], m_blk * BM + BK, b_row);
            if (++ps.hs == h_bufs) { ps.hs = 0; ps.hphase ^= 1; }
          }
          for (int k = k_begin; k < k_end; ++k) {
            mbar_wait(&c.empty_bar[ps.stage], ps.phase ^ 1);
            mbar_arrive_expect_tx(&c.full_bar[ps.stage], tx);
            uint8_t* sa = smem + ps.stage * Cfg::STAGE_BYTES;
            tma_load_2d(sa, &P.tmapA, &c.full_bar[ps.stage], P.kit[k].a_col,
                        P.a_row_base + m_blk * BM + P.kit[k].a_row_off);
            if (MC) {
              // this CTA's half of the B tile, delivered to both CTAs of the pair
              const int half_rows = P.n_tile >> 1;
              tma_load_2d_mc(sa + Cfg::A_BYTES + crank * half_rows * (BK * 2), &P.tmapB, &c.full_bar[ps.stage],
                             P.kit[k].b_col, b_row + crank * half_rows, 0x3);
            } else {
              tma_load_2d(sa + Cfg::A_BYTES, &P.tmapB, &c.full_bar[ps.stage], P.kit[k].b_col, b_row);
            }
            if (++ps.stage == Cfg::STAGES) { ps.stage = 0; ps.phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0 && EPI == EPI_CONV3) {
      mbar_wait(&c.hfull_bar[0], ps.bphase);  // resident weights have landed
      ps.bphase ^= 1;
      tc_fence_after();
      const uint32_t bres_addr = smem_u32(bres);
      for (int tile = tile0; tile < num_tiles; tile += tstep) {
        mbar_wait(&c.tempty_bar[ps.as], ps.aphase ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + ps.as * 256;
        for (int kc = 0; kc < P.num_kiters; ++kc) {
          mbar_wait(&c.full_bar[ps.stage], ps.phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + ps.stage * Cfg::STAGE_BYTES);
          for (int j = 0; j < P.conv_taps; ++j) {
            // tap j = the same box, j*dil rows further down (row pitch 128 B)
            const uint64_t da = make_smem_desc_sw128(a_addr + j * P.conv_dil * 128);
            const uint64_t db = make_smem_desc_sw128(bres_addr + (kc * P.conv_taps + j) * Cfg::B_BYTES);
#pragma unroll
            for (int kk = 0; kk < BK / UMMA_K; ++kk)
              umma_f16(acc, da + 2 * kk, db + 2 * kk, P.idesc, (kc | j | kk) ? 1u : 0u);
          }
          umma_commit(&c.empty_bar[ps.stage]);
          if (++ps.stage == Cfg::STAGES) { ps.stage = 0; ps.phase ^= 1; }
        }
        umma_commit(&c.tfull_bar[ps.as]);
        if (++ps.as == acc_stages) { ps.as = 0; ps.aphase ^= 1; }
      }
    } else if (lane == 0) {
      for (int tile = tile0; tile < num_tiles; tile += tstep)
      for (int sub = 0; sub < n_sub; ++sub) {
        const int ks = tile % k_splits;
        const int k_begin = ks * k_per, k_end = min(P.num_kiters, k_begin + k_per);
        mbar_wait(&c.tempty_bar[ps.as], ps.aphase ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + ps.as * 256;
        for (int k = k_begin; k < k_end; ++k) {
          mbar_wait(&c.full_bar[ps.stage], ps.phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + ps.stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
          const uint64_t da = make_smem_desc_sw128(a_addr);
          const uint64_t db = make_smem_desc_sw128(b_addr);
          const uint32_t d_addr = acc + P.kit[k].slot * P.n_tile;
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk) {
            // advance 16 elements = 32 bytes inside the swizzle span: +2 in the (addr>>4) field
            umma_f16(d_addr, da + 2 * kk, db + 2 * kk, P.idesc,
                     ((k > k_begin && P.kit[k].accum) || kk) ? 1u : 0u);
          }
          if (MC) umma_commit_mc(&c.empty_bar[ps.stage], 0x3);  // the slot is free once BOTH CTAs' MMAs retire
          else umma_commit(&c.empty_bar[ps.stage]);            // frees the smem slot once these MMAs retire
          if (++ps.stage == Cfg::STAGES) { ps.stage = 0; ps.phase ^= 1; }
        }
        umma_commit(&c.tfull_bar[ps.as]);  // accumulator complete -> epilogue
        if (++ps.as == acc_stages) { ps.as = 0; ps.aphase ^= 1; }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;    // which part of the tile's column chunks it handles (0..1; EPI_POOL 0..3)
    const int et = threadIdx.x - 64;
    int last_n_blk = -1;
    PoolState pool_state;
    for (int tile = tile0; tile < num_tiles; tile += tstep)
    for (int sub = 0; sub < n_sub; ++sub) {
      const int ot = tile / k_splits, ksplit = tile - ot * k_splits;
      const int m_unit0 = P.m_fastest ? ot % m_units : ot / P.num_n_blocks;
      const int n_blk = P.m_fastest ? ot / m_units : ot - m_unit0 * P.num_n_blocks;
      const int m_unit = P.m_reverse ? m_units - 1 - m_unit0 : m_unit0;
      const int m_blk = MC ? 2 * m_unit + crank : m_unit;
      if ((epi_is_tdnn(EPI) || EPI == EPI_ATT) && n_blk != last_n_blk) {
        // stage this n block's per-column constants (the previous tile's readers all passed the
        // "staging complete" barrier below, so the table may be overwritten)
        last_n_blk = n_blk;
        for (int i = et; i < P.n_tile; i += EPI_THREADS) {
          const int col = n_blk * P.n_tile + i;
          const bool ok = col < P.epi.N_cols;
          epi_sp[i] = (ok && P.epi.bias) ? P.epi.bias[col] : 0.f;
          epi_sp[256 + i] = (ok && P.epi.scale) ? P.epi.scale[col] : 1.f;
          epi_sp[512 + i] = (ok && P.epi.shift) ? P.epi.shift[col] : 0.f;
        }
        epi_named_barrier();
      }
      uint4 pre[2][4];
      if (epi_is_tdnn(EPI)) tdnn_prefetch(P, m_blk, n_blk, quarter, half, lane, pre);
      float pool_g = 0.f;  // EPI_POOL: the channel's global mean, fetched before the accumulator wait
      if (EPI == EPI_POOL) {
        const int ch = m_blk * BM + quarter * 32 + lane;
        if (ch < P.epi.C) pool_g = __ldg(P.epi.gmean + static_cast<size_t>(n_blk) * P.epi.ld_gmean + ch);
      }
      mbar_wait(&c.tfull_bar[ps.as], ps.aphase);
      tc_fence_after();
      const uint32_t acc = tmem_base + ps.as * 256;
      if (EPI == EPI_F32) epilogue_f32(P, m_blk, n_blk, ksplit, acc, quarter, half, lane);
      if (epi_is_tdnn(EPI) || EPI == EPI_ATT) {
        epi_named_barrier();  // every thread has finished writing out the previous tile's staging
        if (epi_is_tdnn(EPI)) epilogue_tdnn<false>(P, m_blk, n_blk, acc, quarter, half, lane, epi_sp, pre, stage_out);
        if (EPI == EPI_ATT) epilogue_tdnn<true>(P, m_blk, n_blk, acc, quarter, half, lane, epi_sp, pre, stage_out);
      }
      if (EPI == EPI_POOL) {
        mbar_wait(&c.hfull_bar[ps.hs], ps.hphase);
        epilogue_pool(P, m_blk, n_blk, sub, acc, quarter, half, lane, epi_region + ps.hs * hbuf_bytes, pool_state,
                      pool_g);
        __syncwarp();
        if (lane == 0) mbar_arrive(&c.hempty_bar[ps.hs]);
        if (++ps.hs == h_bufs) { ps.hs = 0; ps.hphase ^= 1; }
      }
      if (EPI == EPI_AFF) {
        epi_named_barrier();  // previous tile's staging has been written out
        epilogue_aff(P, m_blk, n_blk, acc, quarter, half, lane, stage_out);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&c.tempty_bar[ps.as]);  // TMEM drained: the next tile's MMAs may start
      if (++ps.as == acc_stages) { ps.as = 0; ps.aphase ^= 1; }
      if (epi_is_tdnn(EPI) || EPI == EPI_ATT) {
        epi_named_barrier();   // staging tile complete
        if (epi_is_tdnn(EPI) && P.epi.colsum != nullptr) tdnn_writeout_colsum(P, m_blk, n_blk, stage_out, epi_sp + 768, et);
        else tdnn_writeout(P, m_blk, n_blk, stage_out, et);
      }
      if (EPI == EPI_AFF) {
        epi_named_barrier();
        aff_writeout(P, m_blk, n_blk, stage_out, et);
      }
    }
  }
}

template <int EPI, int MAX_BN>
__global__ void __launch_bounds__(gemm_threads_of(EPI), 1)
gemm_tc_kernel(const __grid_constant__ GemmParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  GemmCtx c;
  PipeState ps;
  pdl_trigger();
  gemm_setup<EPI, MAX_BN>(smem, c);
  pdl_wait();   // barriers, TMEM and descriptors are set up under the previous kernel's tail
  gemm_run<EPI, MAX_BN>(P, c, ps);
  gemm_teardown(c);
}

#if SD_EXPERIMENTS
// The same GEMM on 2-CTA clusters with the B tile multicast (launched with cluster dims {2,1,1}).
template <int EPI, int MAX_BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_mc_kernel(const __grid_constant__ GemmParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  GemmCtx c;
  PipeState ps;
  gemm_setup<EPI, MAX_BN, true>(smem, c);
  gemm_run<EPI, MAX_BN, true>(P, c, ps);
  gemm_teardown<true>(c);
}

#endif  // SD_EXPERIMENTS

// ------------------------------------------------------------------------------------------------
// cta_group::2 variant for the 256-wide TDNN GEMMs (block0, tdnn1/2, MFA): two CTAs (a cluster) own a
// 256 x 256 output tile.  Each loads ITS 128 rows of A and ITS 128 rows of B per k-iteration (32 KB
// instead of 48 KB of operand ingest per SM — the 1-CTA kernel is bound by that ingest, not by the
// tensor pipe), the LEADER's single thread issues tcgen05.mma.cta_group::2 (M = 256), and each CTA's
// TMEM receives the accumulators of its own 128 rows, which its own epilogue warps drain as before.
//   full[s]   leader only   <- TMA complete_tx from BOTH CTAs (+ the leader's expect_tx arrive)
//   empty[s]  both CTAs     <- leader's tcgen05.commit (multicast to both)
//   tfull[a]  both CTAs     <- leader's tcgen05.commit after a tile's last MMA (multicast)
//   tempty[a] leader only   <- 16 arrivals: the 8 epilogue warps of each CTA
struct Cfg2sm {
  static constexpr int A_BYTES = BM * BK * 2;             // 16 KB: this CTA's 128 rows of A
  static constexpr int B_BYTES = 128 * BK * 2;            // 16 KB: this CTA's half of the 256-row B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 32 KB
// Clock-stamp trace of the cta_group::2 kernel (GemmParams::trace, null in production; tools/gemm_trace.py).  It stays
// compiled in although SD_EXPERIMENTS is off: the build WITH the (predicated-off) stamps is the faster one — re-measured
// this round, three alternating runs each on one box: MFA 1.060-1.086 ms with, 1.112-1.132 ms without; tdnn1 + tdnn2
// 0.97-0.99 vs 1.01-1.03 ms; block0 0.132-0.138 vs 0.124-0.126 ms (ptxas allocates 165 vs 163 registers and schedules
// the epilogue differently).  Net 0.07 ms per step in favour of keeping them.
#ifndef SD_TRACE_ON
#define SD_TRACE_ON 1
#endif
#ifndef SD_2SM_STAGES
#define SD_2SM_STAGES 4
#endif
  static constexpr int STAGES = SD_2SM_STAGES;
  static constexpr int OUT_STAGE_BYTES = 65536;
  static constexpr int EPI_BYTES = (3 * 256 + 1024) * 4;   // per-column constants + colsum exchange
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_STAGE_BYTES + EPI_BYTES + 256;
};

template <int EPI>  // EPI_TDNN only (a template so the header can be included from several translation units)
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_2sm_kernel(const __grid_constant__ GemmParams P) {
  static_assert(EPI == EPI_TDNN, "the cta_group::2 kernel carries the TDNN epilogue");
  using Cfg = Cfg2sm;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  pdl_trigger();
  uint8_t* const stage_out = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  float* const epi_sp = reinterpret_cast<float*>(stage_out + Cfg::OUT_STAGE_BYTES);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(stage_out + Cfg::OUT_STAGE_BYTES + Cfg::EPI_BYTES);
  uint64_t* const full_bar = bars;                    // [STAGES]  (used in the leader)
  uint64_t* const empty_bar = bars + Cfg::STAGES;     // [STAGES]
  uint64_t* const tfull_bar = bars + 2 * Cfg::STAGES; // [2]
  uint64_t* const tempty_bar = tfull_bar + 2;         // [2]       (used in the leader)
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = static_cast<int>(cluster_ctarank());
  const bool leader = crank == 0;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&P.tmapA);
      tma_prefetch_desc(&P.tmapB);
      for (int s = 0; s < Cfg::STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tfull_bar[s], 1);
        mbar_init(&tempty_bar[s], 2 * EPI_WARPS);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above ran under the previous kernel's tail (programmatic dependent launch)

  const int m_units = (P.num_m_blocks + 1) / 2;
  const int num_tiles = m_units * P.num_n_blocks;
  const int tile0 = blockIdx.x >> 1, tstep = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // L2 prefetch of the A boxes `pf` k-iterations ahead of the loads (GemmParams::a_prefetch): the four-stage
      // ring holds 128 KB in flight per SM, which covers ~2000 cycles of load latency at one k-iteration per 512
      // cycles only just (SD_GEMM_TRACE: the MMA thread waits on full[] — 650 cycles per k-iteration); A streams from
      // HBM, so the ring is effectively deepened in L2, where it costs no shared memory.  (B, the weights, is L2-resident.)
      const int pf = SD_EXPERIMENTS ? P.a_prefetch : 0;
      int pf_tile = tile0, pf_k = 0;
      auto prefetch_next = [&]() {
        if (pf_tile >= num_tiles) return;
        const int pm = pf_tile / P.num_n_blocks;
        tma_prefetch_l2_2d(&P.tmapA, P.kit[pf_k].a_col, P.a_row_base + (2 * pm + crank) * BM + P.kit[pf_k].a_row_off);
        if (++pf_k == P.num_kiters) { pf_k = 0; pf_tile += tstep; }
      };
      for (int i = 0; i < pf; ++i) prefetch_next();
      for (int tile = tile0; tile < num_tiles; tile += tstep) {
        const int m_unit = tile / P.num_n_blocks;
        const int n_blk = tile - m_unit * P.num_n_blocks;
        const int m_blk = 2 * m_unit + crank;
        // everything the two TMA instructions need is in registers BEFORE the wait for the free stage: the time
        // from "stage free" to "loads issued" adds to the load latency the four-stage ring has to cover
        const int a_row0 = P.a_row_base + m_blk * BM, b_row = n_blk * P.n_tile + crank * 128;
        int a_col = P.kit[0].a_col, a_off = P.kit[0].a_row_off, b_col = P.kit[0].b_col;
        for (int k = 0; k < P.num_kiters; ++k) {
          if (pf > 0) prefetch_next();
          const int kn = k + 1 < P.num_kiters ? k + 1 : 0;
          const int na_col = P.kit[kn].a_col, na_off = P.kit[kn].a_row_off, nb_col = P.kit[kn].b_col;
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint64_t* const fb = &full_bar[stage];
          const int a_row = a_row0 + a_off;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(fb, 2 * Cfg::STAGE_BYTES);  // both CTAs' bytes
          tma_load_2d_2sm(sa, &P.tmapA, fb, a_col, a_row);
          tma_load_2d_2sm(sa + Cfg::A_BYTES, &P.tmapB, fb, b_col, b_row);
          a_col = na_col; a_off = na_off; b_col = nb_col;
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      const uint32_t idesc = P.idesc;  // encoded with M = 256 by the host
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tstep) {
        const int tl = (tile - tile0) / tstep;   // this CTA pair's tile counter (trace slot)
        if (SD_TRACE_ON && P.trace && blockIdx.x == 0 && tl < 32) P.trace[tl * 16 + 0] = clock64();
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        if (SD_TRACE_ON && P.trace && blockIdx.x == 0 && tl < 32) P.trace[tl * 16 + 1] = clock64();
        const uint32_t acc = tmem_base + as * 256;
        for (int k = 0; k < P.num_kiters; ++k) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t da = make_smem_desc_sw128(a_addr);
          const uint64_t db = make_smem_desc_sw128(a_addr + Cfg::A_BYTES);
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk)
            umma_f16_2sm(acc, da + 2 * kk, db + 2 * kk, idesc, (k | kk) ? 1u : 0u);  // kit[k].accum == (k > 0) for every TDNN layer
          umma_commit_2sm(&empty_bar[stage], 0x3);  // frees this stage in BOTH CTAs
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tfull_bar[as], 0x3);       // accumulators complete in both CTAs' TMEM
        if (SD_TRACE_ON && P.trace && blockIdx.x == 0 && tl < 32) P.trace[tl * 16 + 2] = clock64();
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;
    const bool tma_out = (P.epi.flags & EF_TMA_OUT) != 0;
    if (tma_out && et == 0) {
      tma_prefetch_desc(&P.tmapH);
      if (P.epi.out2 != nullptr) tma_prefetch_desc(&P.tmapO2);
    }
    int last_n_blk = -1, as = 0;
    uint32_t aphase = 0;
    float cpre[3] = {0.f, 1.f, 0.f};   // next tile's {bias, scale, shift} of column `et`, requested a tile ahead
    int cpre_blk = -1;
    for (int tile = tile0; tile < num_tiles; tile += tstep) {
      const int m_unit = tile / P.num_n_blocks;
      const int n_blk = tile - m_unit * P.num_n_blocks;
      const int m_blk = 2 * m_unit + crank;
      if (n_blk != last_n_blk) {
        // this n block's per-column constants -> shared.  With 74 CTA pairs on 4 (or 12) n blocks the block changes
        // on every tile, so the three values were requested during the previous tile's write-out (cpre below):
        // fetched here they cost an L2 round trip per tile on the epilogue's critical path (block0: 2.1 k of 9 k cycles).
        last_n_blk = n_blk;
        if (cpre_blk == n_blk) {
          if (et < P.n_tile) {
            epi_sp[et] = cpre[0];
            epi_sp[256 + et] = cpre[1];
            epi_sp[512 + et] = cpre[2];
          }
        } else {
          for (int i = et; i < P.n_tile; i += EPI_THREADS) {
            const int col = n_blk * P.n_tile + i;
            const bool ok = col < P.epi.N_cols;
            epi_sp[i] = (ok && P.epi.bias) ? P.epi.bias[col] : 0.f;
            epi_sp[256 + i] = (ok && P.epi.scale) ? P.epi.scale[col] : 1.f;
            epi_sp[512 + i] = (ok && P.epi.shift) ? P.epi.shift[col] : 0.f;
          }
        }
        epi_named_barrier();
      }
      uint4 pre[2][4];
      tdnn_prefetch(P, m_blk, n_blk, quarter, half, lane, pre);
      const int tl = (tile - tile0) / tstep;
      const bool tr = SD_TRACE_ON && P.trace != nullptr && blockIdx.x == 0 && tl < 32 && lane == 0 && (warp == 2 || warp == 9);
      long long* const tp = P.trace + tl * 16 + (warp == 2 ? 3 : 9);
      if (tr) tp[0] = clock64();
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      if (tr) tp[1] = clock64();
      const uint32_t acc = tmem_base + as * 256;
      if (tma_out && et == 0) tma_store_wait_read();  // the previous tile's tensor stores have read the staging
      epi_named_barrier();  // every thread has finished writing out the previous tile's staging
      if (tr) tp[2] = clock64();
      epilogue_tdnn<false>(P, m_blk, n_blk, acc, quarter, half, lane, epi_sp, pre, stage_out);
      if (tr) tp[3] = clock64();
      tc_fence_before();
      if (tma_out) fence_proxy_async();  // this thread's staging writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);  // this CTA's share of the accumulator is drained
      if (++as == 2) { as = 0; aphase ^= 1; }
      epi_named_barrier();  // staging tile complete
      if (tr) tp[4] = clock64();
      {
        // request the next tile's constants now: the loads complete under this tile's write-out
        const int nt = tile + tstep;
        cpre_blk = -1;
        if (SD_CPRE && nt < num_tiles && P.n_tile <= EPI_THREADS) {
          const int nb = nt - (nt / P.num_n_blocks) * P.num_n_blocks;
          if (nb != n_blk) {
            cpre_blk = nb;
            const int col = nb * P.n_tile + et;
            const bool ok = et < P.n_tile && col < P.epi.N_cols;
            cpre[0] = (ok && P.epi.bias) ? P.epi.bias[col] : 0.f;
            cpre[1] = (ok && P.epi.scale) ? P.epi.scale[col] : 1.f;
            cpre[2] = (ok && P.epi.shift) ? P.epi.shift[col] : 0.f;
          }
        }
      }
      if (tma_out && (P.epi.flags & EF_REFLECT)) {
        // k > 1 layer leaving through TMA: the tile's halo rows were computed from rows outside the window, so they
        // are replaced IN THE STAGING TILE by their mirror images (interior row t = j for halo row t = -j, row
        // T-1-j for T-1+j; the host enables this path only when a window's halo and its sources share a tile) and
        // the pad rows behind the halo by zeros — what the thread write-out does with r2 / r3 and `valid`.
        const EpiParams& E = P.epi;
        const int we = et >> 5, ln = et & 31;
        const int cj = ln >> 3, sub = ln & 7;
        uint8_t* sc = stage_out + cj * 16384;
        int pw = (m_blk * BM + we) % E.Tp;
        for (int rl = we; rl < BM; rl += EPI_WARPS) {
          const int t = pw - E.H;
          int src = -2;                               // -2 keep, -1 zero, >= 0 copy that staging row
          if (t < 0) src = rl - 2 * t;
          else if (t >= E.T) src = (t - (E.T - 1) <= E.H) ? rl - 2 * (t - (E.T - 1)) : -1;
          if (src != -2) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (src >= 0 && src < BM) v = *reinterpret_cast<const uint4*>(sc + src * 128 + ((sub ^ (src & 7)) << 4));
            *reinterpret_cast<uint4*>(sc + rl * 128 + ((sub ^ (rl & 7)) << 4)) = v;
          }
          pw += EPI_WARPS;
          if (pw >= E.Tp) pw -= E.Tp;
        }
        fence_proxy_async();
        epi_named_barrier();
      }
      if (tma_out) {
        if (et == 0) {
          // four [128 rows x 64 channels] boxes, already in the 128-byte-swizzled layout; rows >= M_rows are clipped
          const int col0 = n_blk * 256, row0 = m_blk * BM;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            tma_store_2d(&P.tmapH, stage_out + j * 16384, col0 + j * 64, row0);
            if (P.epi.out2 != nullptr && col0 + j * 64 < P.epi.out2_cols)
              tma_store_2d(&P.tmapO2, stage_out + j * 16384, col0 + j * 64, row0);
          }
          tma_store_commit();
        }
        if (P.epi.colsum != nullptr) tdnn_writeout_colsum<false>(P, m_blk, n_blk, stage_out, epi_sp + 768, et);
      } else if (P.epi.colsum != nullptr) {
        tdnn_writeout_colsum(P, m_blk, n_blk, stage_out, epi_sp + 768, et);
      } else {
        tdnn_writeout(P, m_blk, n_blk, stage_out, et);
      }
      if (tr) tp[5] = clock64();
    }
    if (tma_out && et == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves while the peer may still signal its barriers
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

#if SD_EXPERIMENTS
// A CHAIN of dependent GEMMs in one cooperative launch: step s+1 reads what step s wrote
// (Res2Net: y_i = TDNN_i(x_i + y_{i-1}); around it the block's two 1x1 TDNNs).  Between steps
// the whole grid synchronises; the generic-proxy stores of the epilogues are made visible to
// the next step's TMA loads (async proxy) by fence.proxy.async on both sides of the barrier.
// Barriers, TMEM and the smem ring are set up once; the pipeline positions carry over.
// `steps` lives in global memory (TMA descriptors included; 64-byte aligned).
template <int EPI, int MAX_BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_chain_kernel(const GemmParams* __restrict__ steps, int num_steps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  GemmCtx c;
  PipeState ps;
  gemm_setup<EPI, MAX_BN>(smem, c);
  for (int s = 0; s < num_steps; ++s) {
    gemm_run<EPI, MAX_BN>(steps[s], c, ps);
    if (s + 1 < num_steps) {
      asm volatile("fence.proxy.async;" ::: "memory");
      __threadfence();
      grid.sync();
      asm volatile("fence.proxy.async;" ::: "memory");
    }
  }
  gemm_teardown(c);
}

#endif  // SD_EXPERIMENTS

}  // namespace sd
