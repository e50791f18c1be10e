// res2net_fused.cuh — the seven dependent dilated k=3 convolutions of one SERes2Net block as ONE launch.
//
// Reference behaviour: speechbrain Res2NetBlock inside SERes2NetBlock (reached from
// speech_encode.py:64-78 through EncoderClassifier.encode_batch): the 1024 channels of tdnn1's output u
// are split into 8 sub-bands x_0..x_7 of 128 channels; y_0 = x_0, y_1 = TDNN_1(x_1),
// y_i = TDNN_i(x_i + y_{i-1}); every TDNN is Conv1d(128 -> 128, k = 3, dilation d, reflect padding) -> ReLU
// -> BatchNorm.  The concatenation v = [y_0 .. y_7] feeds tdnn2.
//
// Why one kernel: launched one convolution at a time (gemm_tc_kernel<EPI_CONV3>) each step is ~30 us of
// which < 2 us is tensor-core work — the chain is serial, every step round-trips its 128-channel slice
// through L2/HBM and pays a launch plus a pipeline fill.  Reflect padding makes every WINDOW independent
// of its neighbours at every layer, so a CTA can carry one window through all seven convolutions with the
// running input (x_{i+1} + y_i) never leaving shared memory:
//
//   A buffer  [2 x 64-channel chunks][T + 2d rows][128 B], 128-byte swizzled exactly as TMA would write it.
//             Conv 1's input (+ its reflect halo, which tdnn1 produced for free) arrives by TMA; for the
//             later convs the epilogue writes x_{i+1} + y_i (and the mirrored halo rows) straight into it.
//             tcgen05 applies the swizzle from absolute address bits, so tap j is the same buffer read
//             j*d rows further down (tools/micro/umma_rowoff_test.cu).
//   weights   streamed through a 3-slot ring of [128 out x 64 in] boxes (L2-resident: 672 KB per block).
//   TMEM      two 128-row accumulators (T <= 160 -> two M tiles; the second one's rows >= T are never read).
//   epilogue  8 warps: TMEM -> bias/ReLU/BN -> f16 y_i -> (a) per-warp staging -> coalesced 64-byte row
//             segments of v (+ mirrored halo rows), (b) x_{i+1} + y_i -> A buffer.
//
// Two CTAs are resident per SM (110 KB of shared memory and 256 TMEM columns each), so one CTA's
// MMAs run under the other's epilogue.  The arithmetic (operand order, f16 rounding points) is the same
// as the per-conv path, so both produce bit-identical v (tests/test_gpu_ecapa.py).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "sd_ptx.cuh"

namespace sd {

constexpr int R2_THREADS = 320;             // warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue
constexpr int R2_CONVS = 7;
constexpr int R2_SUB = 128;                 // channels per sub-band
constexpr int R2_RA_MAX = 168;              // rows of the A buffer: T + 2*dil <= 168
constexpr int R2_A_CHUNK = R2_RA_MAX * 128; // bytes per 64-channel chunk (a multiple of 1024)
constexpr int R2_A_BYTES = 2 * R2_A_CHUNK;
constexpr int R2_WSLOTS = 3;
constexpr int R2_WBOX = 128 * 128;          // [128 out][64 in] f16
constexpr int R2_YSTAGE = 8 * 2048;         // per epilogue warp: [32 rows][64 B]
constexpr int R2_CONST = 2 * 384 * 4;       // double-buffered {bias, scale, shift}[128]
constexpr int R2_SMEM = R2_A_BYTES + R2_WSLOTS * R2_WBOX + R2_YSTAGE + R2_CONST + 256;
static_assert(R2_A_CHUNK % 1024 == 0, "A chunks must keep the swizzle alignment");
static_assert(2 * (R2_SMEM + 1024) <= 233472, "two CTAs per SM");

struct alignas(64) Res2Params {
  CUtensorMap tmapU;              // u [rows, ld] f16, box = 64 channels x (T + 2*dil) rows
  CUtensorMap tmapW[R2_CONVS];    // conv i weights [128 out, 3 taps * 128 in] f16, box = 64 x 128
  CUtensorMap tmapWh[R2_CONVS];   // the same tensor, box = 64 x 64: half a weight box (2-CTA multicast, res2net_pipe.cuh)
  CUtensorMap tmapV;              // v as {channels, T frames, windows}, unswizzled box 32 x 16 x 1 (res2net_pipe.cuh's y stores)
  const float* bias[R2_CONVS];
  const float* scale[R2_CONVS];
  const float* shift[R2_CONVS];
  const __half* u;                // tdnn1 output  [rows, ld]
  __half* v;                      // Res2Net output [rows, ld] (sub-band 0 is written by tdnn1)
  int ld;
  int B, T, Tp, H, dil;
  uint32_t idesc;                 // M = 128, N = 128, f16
  uint32_t idesc_t1;              // M = 128, N = 32, f16: MODE 3's transposed second tile
  long long* trace;               // debug (SD_R2_TRACE): CTA 0's per-conv clock64 stamps, [conv][18]
  int* oflow;                     // overflow flag (may be null): a y_i left the f16 range and was saturated.  A sum
                                  // x_{i+1} + y_i that overflows becomes +-inf in the A buffer; the next convolution's
                                  // accumulators then carry it into this check
};

// finer stamps inside the chunks of epilogue warp 4 (quarter 0: two M-tile passes)
__device__ __forceinline__ void r2_stamp2(const Res2Params& P, int n, int k, int s) {
  if (SD_EXPERIMENTS && P.trace != nullptr && blockIdx.x == 0 && n < 32 && (threadIdx.x >> 5) == 4 && (threadIdx.x & 31) == 0)
    P.trace[32 * 18 + (n * 4 + k) * 8 + s] = clock64();
}
__device__ __forceinline__ void r2_stamp(const Res2Params& P, int n, int slot) {
  if (SD_EXPERIMENTS && P.trace != nullptr && blockIdx.x == 0 && n < 32) P.trace[n * 18 + slot] = clock64();
}

// MODE 0: y_i goes through a per-warp staging tile and leaves as 64-byte row segments, 8 rows per store
//         instruction; this chunk's x_{i+1} is loaded right behind the TMEM load.
// MODE 2: (default) the x_{i+1} values of chunk 0 are requested BEFORE the wait for the accumulator, those of
//         chunk k+1 as soon as chunk k has consumed its own: 0.174 -> 0.163 ms per block at B = 512.
// MODE 1: MODE 2 + y_i stored straight from the registers (one row per lane, 4 x 16 B) without the staging round
//         trip.  Measured slower (0.192 ms): every store instruction then touches 32 different lines.
// MODE 3: MODE 2 + the frames beyond 127 (23 of them at T = 151) are computed TRANSPOSED: D1^T[channel, frame] =
//         W . X^T with M = 128 channels (TMEM lanes) and N = 32 frames (TMEM columns 128-159) instead of a second
//         128-row M tile of which 105 rows are waste.  A quarter of the tensor work of that tile, and — the point —
//         its epilogue is spread over all eight warps (32 channels x 16 frames each) instead of sitting on the two
//         warps that own TMEM lanes 0-31: the serial conv-to-conv path shrinks from 4 to 2.5 chunk units.
template <int MODE>
__global__ void __launch_bounds__(R2_THREADS, 2)
res2net_fused_kernel(const __grid_constant__ Res2Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  pdl_trigger();
  uint8_t* const abuf = smem;
  uint8_t* const wring = smem + R2_A_BYTES;
  uint8_t* const ystage = wring + R2_WSLOTS * R2_WBOX;
  float* const consts = reinterpret_cast<float*>(ystage + R2_YSTAGE);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(consts) + R2_CONST);
  uint64_t* const w_full = bars;         // [3] weight box landed
  uint64_t* const w_empty = bars + 3;    // [3] its MMAs retired
  uint64_t* const a_full = bars + 6;     // conv 1's input landed (TMA)
  uint64_t* const t_full = bars + 7;     // a conv's accumulators are complete
  uint64_t* const acc_free = bars + 8;   // epilogue drained TMEM and wrote the next conv's input
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = P.T, Tp = P.Tp, H = P.H, d = P.dil, ld = P.ld;
  const int RA = T + 2 * d;
  const int n_mt = (T + 127) >> 7;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < R2_WSLOTS; ++s) {
        mbar_init(&w_full[s], 1);
        mbar_init(&w_empty[s], 1);
      }
      mbar_init(a_full, 1);
      mbar_init(t_full, 1);
      mbar_init(acc_free, 8);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // set-up above overlaps the previous kernel's tail (programmatic dependent launch)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&P.tmapU);
      for (int i = 0; i < R2_CONVS; ++i) tma_prefetch_desc(&P.tmapW[i]);
      int slot = 0;
      uint32_t ph = 0;
      int n = 0;  // convolutions this CTA has started (all roles count the same sequence)
      for (int ws = blockIdx.x; ws < P.B; ws += gridDim.x) {
        const int w = P.B - 1 - ws;  // last windows first: tdnn1 wrote them last, so they are still in L2
        if (n > 0) mbar_wait(t_full, (n - 1) & 1);  // the previous window's last conv is done reading A
        mbar_arrive_expect_tx(a_full, static_cast<uint32_t>(2 * RA * 128));
        const int row0 = w * Tp + H - d;
        tma_load_2d(abuf, &P.tmapU, a_full, R2_SUB, row0);
        tma_load_2d(abuf + R2_A_CHUNK, &P.tmapU, a_full, R2_SUB + 64, row0);
        for (int i = 0; i < R2_CONVS; ++i)
          for (int kc = 0; kc < 2; ++kc)
            for (int j = 0; j < 3; ++j) {
              mbar_wait(&w_empty[slot], ph ^ 1);
              mbar_arrive_expect_tx(&w_full[slot], R2_WBOX);
              tma_load_2d(wring + slot * R2_WBOX, &P.tmapW[i], &w_full[slot], j * R2_SUB + kc * 64, 0);
              if (++slot == R2_WSLOTS) { slot = 0; ph ^= 1; }
            }
        n += R2_CONVS;
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      int n = 0, wi = 0;
      const uint32_t a_addr0 = smem_u32(abuf);
      for (int ws = blockIdx.x; ws < P.B; ws += gridDim.x, ++wi) {
        for (int i = 0; i < R2_CONVS; ++i, ++n) {
          if (i == 0) mbar_wait(a_full, wi & 1);
          if (n > 0) mbar_wait(acc_free, (n - 1) & 1);  // TMEM drained, next input written (i > 0)
          tc_fence_after();
          r2_stamp(P, n, 0);
          for (int kc = 0; kc < 2; ++kc)
            for (int j = 0; j < 3; ++j) {
              if (SD_EXPERIMENTS && P.trace != nullptr && blockIdx.x == 0 && n < 32) P.trace[1600 + n * 12 + (kc * 3 + j) * 2] = clock64();
              mbar_wait(&w_full[slot], ph);
              tc_fence_after();
              if (SD_EXPERIMENTS && P.trace != nullptr && blockIdx.x == 0 && n < 32) P.trace[1600 + n * 12 + (kc * 3 + j) * 2 + 1] = clock64();
              const uint64_t db = make_smem_desc_sw128(smem_u32(wring + slot * R2_WBOX));
              if (MODE == 3 && n_mt > 1) {
                // frames 128.. transposed: "A" = the weight box (128 output channels), "B" = 32 rows of the input
                const uint64_t dx = make_smem_desc_sw128(a_addr0 + kc * R2_A_CHUNK + (128 + j * d) * 128);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_f16(tmem_base + 128, db + 2 * kk, dx + 2 * kk, P.idesc_t1, (kc | j | kk) ? 1u : 0u);
              }
              for (int m = 0; m < (MODE == 3 ? 1 : n_mt); ++m) {
                // tap j of M tile m: rows m*128 + j*d .. of the same buffer (row pitch 128 B)
                const uint64_t da = make_smem_desc_sw128(a_addr0 + kc * R2_A_CHUNK + (m * 128 + j * d) * 128);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_f16(tmem_base + m * 128, da + 2 * kk, db + 2 * kk, P.idesc, (kc | j | kk) ? 1u : 0u);
              }
              umma_commit(&w_empty[slot]);
              if (++slot == R2_WSLOTS) { slot = 0; ph ^= 1; }
            }
          umma_commit(t_full);
          r2_stamp(P, n, 1);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------------- epilogue
    const int quarter = warp & 3;       // TMEM lane quarter
    const int half = (warp - 2) >> 2;   // which 64 of the 128 output channels
    const int et = threadIdx.x - 64;
    uint8_t* const my_stage = ystage + (warp - 2) * 2048;
    // this thread's slots of the {bias, scale, shift} table (384 floats, 256 threads)
    const int k0 = et, k1 = et + 256;
    auto const_src = [&](int conv, int idx) -> const float* {
      const int which = idx >> 7, c = idx & 127;
      return (which == 0 ? P.bias[conv] : which == 1 ? P.scale[conv] : P.shift[conv]) + c;
    };
    if (blockIdx.x < P.B) {
      consts[k0] = __ldg(const_src(0, k0));
      if (k1 < 384) consts[k1] = __ldg(const_src(0, k1));
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // number of (M tile, 32-column) chunks this warp works through per conv
    const int my_mt = (quarter * 32 < T ? 1 : 0) + ((MODE != 3 && 128 + quarter * 32 < T) ? 1 : 0);
    const int nk = 2 * my_mt;
    // MODE 3: this warp's share of the transposed tile: channels quarter*32 + lane, frames 128 + half*16 .. +15
    // (the frames beyond 127 are split evenly between the two warps of a quarter: 12 + 11 at T = 151)
    const bool has_t1 = MODE == 3 && T > 128;
    const int t1_h0 = (T - 128 + 1) >> 1;
    const int t1_f0 = 128 + half * t1_h0;
    const int t1_end = half == 0 ? 128 + t1_h0 : T;
    const int t1_c = quarter * 32 + lane;
    int n = 0;
    float amax = 0.f;   // largest |y| this thread produced
    for (int ws = blockIdx.x; ws < P.B; ws += gridDim.x) {
      const int w = P.B - 1 - ws;
      const size_t wrow = static_cast<size_t>(w) * Tp;
      const bool more_windows = ws + static_cast<int>(gridDim.x) < P.B;
      for (int i = 0; i < R2_CONVS; ++i, ++n) {
        const float* const cs = consts + (n & 1) * 384;
        // the next conv's constants go to the other buffer right away: its last readers finished a whole
        // conv ago (end-of-conv barrier below), and the load latency overlaps the wait for this conv's MMAs
        const bool next_conv = i + 1 < R2_CONVS || more_windows;
        const int ni = i + 1 < R2_CONVS ? i + 1 : 0;
        if (next_conv) {
          float* cw = consts + ((n + 1) & 1) * 384;
          cw[k0] = __ldg(const_src(ni, k0));
          if (k1 < 384) cw[k1] = __ldg(const_src(ni, k1));
        }
        const bool has_next = i + 1 < R2_CONVS;
        const int out_col = (i + 1) * R2_SUB + half * 64;  // this warp's columns of v
        const __half* xnext = P.u + (i + 2) * R2_SUB + half * 64 + (wrow + H) * ld;
        // x_{i+1}: pull this warp's rows (one 128-byte line per thread and M tile) towards L2 now, while the
        // MMAs run; the loads proper are issued per chunk right behind the TMEM load, so only one chunk's
        // worth of x is ever live in registers
        if (has_next) {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const int t = mt * 128 + quarter * 32 + lane;
            if (t < T)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(xnext + static_cast<size_t>(t) * ld));
          }
        }
        uint4 xc[4];
        auto load_xc = [&](int k) {
          const int t = (k >> 1) * 128 + quarter * 32 + lane;
          if (t < T) {
            const uint4* a4 = reinterpret_cast<const uint4*>(xnext + static_cast<size_t>(t) * ld + (k & 1) * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) xc[q] = __ldg(a4 + q);
          }
        };
        if (MODE != 0 && has_next && nk > 0) load_xc(0);
        uint32_t xt[8];   // MODE 3: x_{i+1}[frame, this channel] for the 16 transposed frames, two f16 per register
        if (has_t1 && has_next) {
          const unsigned short* xs = reinterpret_cast<const unsigned short*>(xnext) + t1_c - half * 64;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int t = t1_f0 + 2 * jj;
            const uint32_t lo = t < t1_end ? __ldg(xs + static_cast<size_t>(t) * ld) : 0u;
            const uint32_t hi = t + 1 < t1_end ? __ldg(xs + static_cast<size_t>(t + 1) * ld) : 0u;
            xt[jj] = lo | (hi << 16);
          }
        }
        mbar_wait(t_full, n & 1);
        tc_fence_after();
        if (lane == 0) r2_stamp(P, n, 2 + 2 * (warp - 2));
        if (has_t1) {
          uint32_t acc[16];
          __syncwarp();
          tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t1_f0, acc);
          tmem_ld_wait();
          const float cb = cs[t1_c], csc = cs[128 + t1_c], csh = cs[256 + t1_c];
          __half* const vcol = P.v + (i + 1) * R2_SUB + t1_c + (wrow + H) * ld;   // this channel's column of y_i
          uint8_t* const acol = abuf + (t1_c >> 6) * R2_A_CHUNK + (t1_c & 7) * 2;
          const int piece = (t1_c & 63) >> 3;
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const int t = t1_f0 + jj;
            if (t < t1_end) {
              const float yf = fmaf(fmaxf(__uint_as_float(acc[jj]) + cb, 0.f), csc, csh);
              if (kTrackOflow) amax = fmaxf(amax, fabsf(yf));
              const __half y = half_sat(yf);
              vcol[static_cast<long>(t) * ld] = y;
              if (t >= T - 1 - H && t <= T - 2) vcol[static_cast<long>(2 * (T - 1) - t) * ld] = y;
              if (has_next) {
                const __half xv = __ushort_as_half(static_cast<unsigned short>(jj & 1 ? xt[jj >> 1] >> 16 : xt[jj >> 1] & 0xffffu));
                const __half sv = __hadd(xv, y);
                const int p = t + d;
                *reinterpret_cast<__half*>(acol + p * 128 + ((piece ^ (p & 7)) << 4)) = sv;
                if (t >= T - 1 - d && t <= T - 2) {
                  const int p2 = d + 2 * (T - 1) - t;
                  *reinterpret_cast<__half*>(acol + p2 * 128 + ((piece ^ (p2 & 7)) << 4)) = sv;
                }
              }
            }
          }
          if (nk == 0) {   // no tile-0 rows for this quarter (cannot happen with T > 128, kept for symmetry)
            tc_fence_before();
            if (has_next) fence_proxy_async();
          }
        }
        for (int k = 0; k < nk; ++k) {
          const int mt = k >> 1, ci = k & 1;
          const int tw0 = mt * 128 + quarter * 32;  // first frame of this warp's 32 rows
          const int t = tw0 + lane;
          const bool valid = t < T;
          const int c0 = half * 64 + ci * 32;  // column within the conv's 128 outputs
          uint32_t acc[32];
          __syncwarp();
          r2_stamp2(P, n, k, 0);
          tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + mt * 128 + c0, acc);
          // MODE 0: this chunk's x_{i+1} while the accumulator is in flight
          if (MODE == 0 && has_next) load_xc(k);
          tmem_ld_wait();
          r2_stamp2(P, n, k, 1);
          const float4* b4 = reinterpret_cast<const float4*>(cs + c0);
          const float4* s4 = reinterpret_cast<const float4*>(cs + 128 + c0);
          const float4* h4 = reinterpret_cast<const float4*>(cs + 256 + c0);
          uint4 pk[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float x[8];
#pragma unroll
            for (int e2 = 0; e2 < 2; ++e2) {
              const float4 bb = b4[2 * q + e2], ss = s4[2 * q + e2], hh = h4[2 * q + e2];
              const int o = 8 * q + 4 * e2;
              x[4 * e2 + 0] = fmaf(fmaxf(__uint_as_float(acc[o + 0]) + bb.x, 0.f), ss.x, hh.x);
              x[4 * e2 + 1] = fmaf(fmaxf(__uint_as_float(acc[o + 1]) + bb.y, 0.f), ss.y, hh.y);
              x[4 * e2 + 2] = fmaf(fmaxf(__uint_as_float(acc[o + 2]) + bb.z, 0.f), ss.z, hh.z);
              x[4 * e2 + 3] = fmaf(fmaxf(__uint_as_float(acc[o + 3]) + bb.w, 0.f), ss.w, hh.w);
            }
            if (kTrackOflow)
              amax = fmaxf(fmaxf(amax, fmaxf(fabsf(x[0]), fabsf(x[1]))), fmaxf(fmaxf(fabsf(x[2]), fabsf(x[3])),
                           fmaxf(fmaxf(fabsf(x[4]), fabsf(x[5])), fmaxf(fabsf(x[6]), fabsf(x[7])))));
            pk[q].x = pack_half2(x[0], x[1]);
            pk[q].y = pack_half2(x[2], x[3]);
            pk[q].z = pack_half2(x[4], x[5]);
            pk[q].w = pack_half2(x[6], x[7]);
          }
          r2_stamp2(P, n, k, 2);
          // (a) MODE 1: y_i straight to v (+ the mirrored halo rows) from the registers
          if (MODE == 1 && valid) {
            __half* dst = P.v + out_col + ci * 32 + (wrow + H) * ld;
            uint4* d0 = reinterpret_cast<uint4*>(dst + static_cast<long>(t) * ld);
#pragma unroll
            for (int q = 0; q < 4; ++q) d0[q] = pk[q];
            if (t >= 1 && t <= H) {
              uint4* d1 = reinterpret_cast<uint4*>(dst - static_cast<long>(t) * ld);
#pragma unroll
              for (int q = 0; q < 4; ++q) d1[q] = pk[q];
            }
            if (t >= T - 1 - H && t <= T - 2) {
              uint4* d2 = reinterpret_cast<uint4*>(dst + static_cast<long>(2 * (T - 1) - t) * ld);
#pragma unroll
              for (int q = 0; q < 4; ++q) d2[q] = pk[q];
            }
          }
          // (a) MODE 0: y_i -> per-warp staging, 16-byte pieces XOR-swizzled so both sides are conflict-free
          if (MODE != 1) {
            uint8_t* srow = my_stage + lane * 64;
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(srow + ((q ^ sw) << 4)) = pk[q];
          }
          // (b) next conv's input x_{i+1} + y_i, from the f16-rounded y_i as an unfused chain would read it
          // back.  One f16 add: the f32 sum of two f16 values is exact whenever it matters for the f16
          // rounding, so this equals round_f16(float(x) + float(y)) bit for bit.
          r2_stamp2(P, n, k, 3);
          if (has_next && valid) {
            // rows of the A buffer this thread fills: its own and (near the ends) its mirror image
            const int p = t + d;
            int p2 = -1;
            if (t >= 1 && t <= d) p2 = d - t;
            else if (t >= T - 1 - d && t <= T - 2) p2 = d + 2 * (T - 1) - t;
            uint8_t* arow = abuf + (c0 >> 6) * R2_A_CHUNK;
            const int p0 = ci * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 sk;
              const __half2* ah = reinterpret_cast<const __half2*>(&xc[q]);
              const __half2* yh = reinterpret_cast<const __half2*>(&pk[q]);
              __half2* sh = reinterpret_cast<__half2*>(&sk);
#pragma unroll
              for (int e = 0; e < 4; ++e) sh[e] = __hadd2(ah[e], yh[e]);
              *reinterpret_cast<uint4*>(arow + p * 128 + (((p0 + q) ^ (p & 7)) << 4)) = sk;
              if (p2 >= 0) *reinterpret_cast<uint4*>(arow + p2 * 128 + (((p0 + q) ^ (p2 & 7)) << 4)) = sk;
            }
          }
          if (k + 1 == nk) {
            // last chunk: TMEM is drained and the next input is complete — release the MMA warp now; the
            // write-out below only touches the staging buffer and global memory
            tc_fence_before();
            if (has_next) fence_proxy_async();  // A-buffer writes -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) r2_stamp(P, n, 3 + 2 * (warp - 2));
            if (lane == 0) mbar_arrive(acc_free);
          }
          if (MODE != 0 && has_next && k + 1 < nk) load_xc(k + 1);  // xc is free again: the next chunk's x_{i+1}
          __syncwarp();
          r2_stamp2(P, n, k, 4);
          // MODE 0 coalesced write-out: 4 lanes cover one row's 64 bytes, 8 rows per instruction
          if (MODE != 1) {
            const int piece = lane & 3, rsub = lane >> 2;
            __half* dst = P.v + out_col + ci * 32 + piece * 8 + (wrow + H) * ld;
            uint4 val[4];
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {
              const int rl = pass * 8 + rsub;
              val[pass] = *reinterpret_cast<const uint4*>(my_stage + rl * 64 + ((piece ^ ((rl >> 1) & 3)) << 4));
            }
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {
              const int tt = tw0 + pass * 8 + rsub;
              if (tt < T) {
                *reinterpret_cast<uint4*>(dst + static_cast<long>(tt) * ld) = val[pass];
                if (tt >= 1 && tt <= H) *reinterpret_cast<uint4*>(dst - static_cast<long>(tt) * ld) = val[pass];
                if (tt >= T - 1 - H && tt <= T - 2)
                  *reinterpret_cast<uint4*>(dst + static_cast<long>(2 * (T - 1) - tt) * ld) = val[pass];
              }
            }
          }
          r2_stamp2(P, n, k, 5);
        }
        if (nk == 0) {   // a warp without rows (short windows) still owes its arrival
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_free);
        }
        // meet the other epilogue warps (constants for the next conv are complete) — off the critical
        // path: the MMAs of the next conv are running now
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
    if (kTrackOflow && amax > kHalfMax && P.oflow != nullptr) atomicOr(P.oflow, 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace sd
