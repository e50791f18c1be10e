// res2net_pipe.cuh — the Res2Net chain of one SERes2Net block, FOUR windows in flight per SM, computed TRANSPOSED.
//
// Same arithmetic and the same dataflow idea as res2net_fused.cuh (one window is carried through the seven
// dependent dilated k = 3 convolutions with its running input x_{i+1} + y_i held in shared memory; reference:
// speechbrain Res2NetBlock reached from /root/reference/speech_encode.py:64-78), reorganised around what the
// clock-stamp traces of that kernel and of the first versions of this one showed (profiles/r01_r2_trace_mode3.txt,
// profiles/r02_r2p_trace_*.txt):
//
//   * a convolution is ~2 k cycles of tensor work and ~10 k cycles of epilogue on the window's critical path, so
//     with two windows per SM both pipes idle most of the time.  Here ONE CTA per SM owns four window slots
//     (4 x 40 KB input buffers) and walks the jobs (window slot s, convolution i) round-robin
//     (s0,i) (s1,i) (s2,i) (s3,i) (s0,i+1) ...; job j accumulates into TMEM accumulator j % 3 and is finished by
//     epilogue group j % 2 (8 warps each): the MMAs of the next jobs run under the epilogues of the previous ones,
//     and between two convolutions of one window lie the jobs of the three other windows.  148 CTAs x 4 slots
//     >= 512 windows: the BASELINE batch is one round without a tail (the one-window-per-CTA kernel lost 14 % to
//     512 windows over 296 CTAs).
//   * the whole window is computed TRANSPOSED:  D^T[channel, frame] = W . X^T  with M = 128 output channels on the
//     TMEM lanes and N = roundup(T, 16) <= 160 frames in the columns (the weight box is the A operand, the input
//     buffer the B operand; both K-major, and tcgen05 applies the swizzle from absolute address bits, so tap j is
//     the same buffer read j*dil rows further down).  One MMA per (box, k-step) instead of a 128-frame tile plus a
//     32-frame tail: 30 % less shared-memory operand traffic, which is what the tensor core competes for with the
//     epilogue warps in this kernel (8 KB per 64-cycle M128 x N128 MMA = the SM's whole 128 B/clk).
//   * a thread of the epilogue therefore owns ONE CHANNEL: its bias / scale / shift are three registers (no
//     per-column constant table in shared memory), and the 8-channel x 16-frame block of every eight lanes is
//     transposed in registers (three xor-shuffle stages) so that x_{i+1} arrives and y_i / x_{i+1} + y_i leave as
//     16-byte pieces, 64 contiguous bytes per frame per four lanes — no staging tile, no 2-byte stores (sixteen
//     STS.U16 + sixteen STG.U16 per thread cost ~5.5 k of the first version's ~10 k-cycle epilogue).
// Weights stream through a 4-slot TMA ring (L2-resident, 96 KB per convolution).  MC = true (default): the kernel runs
// on 2-CTA clusters and every weight box is fetched ONCE per pair — each CTA loads the half of the box its rank owns
// and multicasts it into both CTAs' ring slots.  With every SM streaming 96 KB of weights per job the launch sat at
// the L2 bandwidth cap (~4.3 KB/clk of weights alone, SD_R2P_DBG probes: without any x / y traffic a job was still
// bound by weight delivery, 550 cycles per 16 KB box against 320 of tensor work), and all other global accesses
// queued behind them.  The two CTAs walk the same sequence of weight boxes: a CTA with one window less than its
// partner consumes (and releases) the boxes of the missing job without issuing MMAs.
//
// The products, their order inside every accumulator and the f16 rounding points are those of the per-convolution
// path, so v is bit-identical to it (tests/test_gpu_ecapa.py).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "res2net_fused.cuh"   // Res2Params, R2_CONVS, R2_SUB, R2_WBOX
#include "sd_ptx.cuh"

#ifndef SD_R2P_DBG
#define SD_R2P_DBG 0     // timing probes (wrong results): 1 no global stores of y, 2 no loads of x_{i+1}, 4 no input-buffer stores
#endif
#ifndef SD_R2P_TRACE
#define SD_R2P_TRACE 0   // 1: compile the clock-stamp trace in (Res2Params::trace, tools/r2p_trace.py); costs registers
#endif

namespace sd {

constexpr int R2P_SLOTS = 4;                      // windows resident per CTA
constexpr int R2P_GROUPS = 2;                     // epilogue groups (8 warps each)
constexpr int R2P_ACCS = 3;                       // TMEM accumulators (160 columns each): with two, handing the accumulator
                                                  // back serialised MMA -> epilogue per group (measured 10.8 k cycles per job)
constexpr int R2P_ACC_COLS = 160;
constexpr int R2P_THREADS = 64 + R2P_GROUPS * 256;   // warp 0 TMA, warp 1 MMA, 2 x 8 epilogue warps
constexpr int R2P_RA = 160;                       // rows of an input buffer: T + 2 * dil <= 160
constexpr int R2P_A_CHUNK = R2P_RA * 128;         // bytes per 64-channel chunk (20 x 1024)
constexpr int R2P_A_BYTES = 2 * R2P_A_CHUNK;
constexpr int R2P_WSLOTS = 3;
constexpr int R2P_YSTAGE = 16 * 1024;             // per epilogue warp: [16 frames][32 channels] f16, the source of a TMA store
constexpr int R2P_SMEM = R2P_SLOTS * R2P_A_BYTES + R2P_WSLOTS * R2_WBOX + R2P_YSTAGE + 256;
static_assert(R2P_A_CHUNK % 1024 == 0, "input chunks must keep the swizzle alignment");
static_assert(R2P_SMEM <= 232448, "exceeds the 227 KB of shared memory a CTA can have");

// 18 warps land 5 + 5 + 4 + 4 on the four sub-partitions (16 384 registers each), so a thread can have at most
// 16384 / (5 * 32) = 102 -> 96 registers.  The epilogue is written to stay under that: with 222 KB of shared memory
// carved out there is hardly any L1 left, and every spilled access would be an L2 round trip.
template <bool MC>
__global__ void __launch_bounds__(R2P_THREADS, 1)
res2net_pipe_kernel(const __grid_constant__ Res2Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  pdl_trigger();
  uint8_t* const abuf = smem;                                        // [slot][chunk][R2P_RA rows][128 B]
  uint8_t* const wring = smem + R2P_SLOTS * R2P_A_BYTES;
  uint8_t* const ystage = wring + R2P_WSLOTS * R2_WBOX;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(ystage + R2P_YSTAGE);
  uint64_t* const w_full = bars;            // [3] weight box landed        (room for 4)
  uint64_t* const w_empty = bars + 4;       // [3] its MMAs retired
  uint64_t* const x_full = bars + 8;        // [4] a slot's x tile landed (TMA): x_1 of a new window, or x_{i+2} after conv i
  uint64_t* const a_ready = bars + 12;      // [4] the epilogue added y_i onto it: the next convolution's input is complete (8 arrivals)
  uint64_t* const b_done = bars + 16;       // [4] a convolution's MMAs have finished reading the slot's buffer
  uint64_t* const t_full = bars + 20;       // [3] accumulator complete
  uint64_t* const t_empty = bars + 23;      // [3] accumulator drained (8 arrivals)
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = P.T, Tp = P.Tp, H = P.H, d = P.dil, ld = P.ld;
  const int RA = T + 2 * d;
  const int G = gridDim.x;
  const int n_mine = (P.B - static_cast<int>(blockIdx.x) + G - 1) / G;   // windows blockIdx.x, + G, ...
  const int rounds = (n_mine + R2P_SLOTS - 1) / R2P_SLOTS;
  // the weight-box sequence is the pair's: that of its even CTA, which never has fewer windows than the odd one
  const int crank = MC ? static_cast<int>(cluster_ctarank()) : 0;
  const int n_pair = MC ? (P.B - (static_cast<int>(blockIdx.x) & ~1) + G - 1) / G : n_mine;
  const int rounds_w = (n_pair + R2P_SLOTS - 1) / R2P_SLOTS;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < R2P_WSLOTS; ++s) {
        mbar_init(&w_full[s], 1);
        mbar_init(&w_empty[s], MC ? 2 : 1);     // MC: both CTAs of the pair must release a slot
      }
      for (int s = 0; s < R2P_SLOTS; ++s) {
        mbar_init(&x_full[s], 1);
        mbar_init(&a_ready[s], 8);
        mbar_init(&b_done[s], 1);
      }
      for (int a = 0; a < R2P_ACCS; ++a) {
        mbar_init(&t_full[a], 1);
        mbar_init(&t_empty[a], 8);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();   // the peer's barriers exist before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // window of (round r, slot s): the last windows first — tdnn1 wrote them last, so they are still in L2
  auto window_of = [&](int r, int s) { return P.B - 1 - (static_cast<int>(blockIdx.x) + (r * R2P_SLOTS + s) * G); };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producers
    // lane 0: the weight ring.  lane 1: the x tiles — x_1 of a window when its slot is taken over, and x_{i+2} into the
    // slot's buffer as soon as convolution i's MMAs have finished reading it; the epilogue then only ADDS y_{i+1} onto
    // it in shared memory.  (The first versions loaded x into registers in the epilogue: with one 16-frame block of
    // look-ahead that was an exposed L2 / HBM round trip per block, 8.5 k of the epilogue's 17.5 k cycles.)
    if (lane == 0) {
      for (int i = 0; i < R2_CONVS; ++i) tma_prefetch_desc(MC ? &P.tmapWh[i] : &P.tmapW[i]);
      int slot = 0;
      uint32_t ph = 0;
      for (int r = 0; r < rounds_w; ++r) {
        const int nact_w = min(R2P_SLOTS, n_pair - r * R2P_SLOTS);
        for (int i = 0; i < R2_CONVS; ++i)
          for (int s = 0; s < nact_w; ++s)
            for (int kc = 0; kc < 2; ++kc)
              for (int j = 0; j < 3; ++j) {
                mbar_wait(&w_empty[slot], ph ^ 1);
                mbar_arrive_expect_tx(&w_full[slot], R2_WBOX);   // the whole box: this CTA's half + the peer's
                if (MC)
                  tma_load_2d_mc(wring + slot * R2_WBOX + crank * (R2_WBOX / 2), &P.tmapWh[i], &w_full[slot],
                                 j * R2_SUB + kc * 64, crank * 64, 0x3);
                else
                  tma_load_2d(wring + slot * R2_WBOX, &P.tmapW[i], &w_full[slot], j * R2_SUB + kc * 64, 0);
                if (++slot == R2P_WSLOTS) { slot = 0; ph ^= 1; }
              }
      }
    } else if (lane == 1) {
      tma_prefetch_desc(&P.tmapU);
      for (int r = 0; r < rounds; ++r) {
        const int nact = min(R2P_SLOTS, n_mine - r * R2P_SLOTS);
        for (int i = 0; i < R2_CONVS; ++i)
          for (int s = 0; s < nact; ++s) {
            const int n = r * R2_CONVS + i;            // the slot's n-th tile: input base of its n-th convolution
            if (n > 0) mbar_wait(&b_done[s], (n - 1) & 1);   // the slot's previous convolution is done reading
            mbar_arrive_expect_tx(&x_full[s], static_cast<uint32_t>(2 * RA * 128));
            const int row0 = window_of(r, s) * Tp + H - d;
            tma_load_2d(abuf + s * R2P_A_BYTES, &P.tmapU, &x_full[s], (i + 1) * R2_SUB, row0);
            tma_load_2d(abuf + s * R2P_A_BYTES + R2P_A_CHUNK, &P.tmapU, &x_full[s], (i + 1) * R2_SUB + 64, row0);
          }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      int job = 0;
      uint32_t ready_ph = 0, empty_ph = 0;   // phase bits: bit s of a_ready[s], bit a of t_empty[a]
      int acc = 0;
      const uint32_t idesc = make_idesc_f16(((T + 15) >> 4) << 4, 0);   // M = 128 channels, N = frames
      for (int r = 0; r < rounds_w; ++r) {
        const int nact = max(0, min(R2P_SLOTS, n_mine - r * R2P_SLOTS));
        const int nact_w = min(R2P_SLOTS, n_pair - r * R2P_SLOTS);
        for (int i = 0; i < R2_CONVS; ++i)
          for (int s = 0; s < nact_w; ++s) {
            if (s >= nact) {
              // the partner's job without a counterpart here: consume and release its weight boxes
              for (int b = 0; b < 6; ++b) {
                mbar_wait(&w_full[slot], ph);
                umma_commit_mc(&w_empty[slot], 0x3);
                if (++slot == R2P_WSLOTS) { slot = 0; ph ^= 1; }
              }
              continue;
            }
            if (i == 0) mbar_wait(&x_full[s], (r * R2_CONVS) & 1);   // x_1 of a new window is the whole input
            else { mbar_wait(&a_ready[s], (ready_ph >> s) & 1u); ready_ph ^= 1u << s; }
            if (job >= R2P_ACCS) { mbar_wait(&t_empty[acc], (empty_ph >> acc) & 1u); empty_ph ^= 1u << acc; }
            tc_fence_after();
            if (SD_R2P_TRACE && P.trace != nullptr && blockIdx.x == 0 && job < 32) P.trace[job * 18 + 0] = clock64();
            const uint32_t d0 = tmem_base + acc * R2P_ACC_COLS;
            const uint32_t a_addr0 = smem_u32(abuf + s * R2P_A_BYTES);
            for (int kc = 0; kc < 2; ++kc)
              for (int j = 0; j < 3; ++j) {
                mbar_wait(&w_full[slot], ph);
                tc_fence_after();
                // "A" = the weight box (128 output channels x 64 inputs), "B" = the input rows j*dil .. of chunk kc
                const uint64_t dw = make_smem_desc_sw128(smem_u32(wring + slot * R2_WBOX));
                const uint64_t dx = make_smem_desc_sw128(a_addr0 + kc * R2P_A_CHUNK + (j * d) * 128);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_f16(d0, dw + 2 * kk, dx + 2 * kk, idesc, (kc | j | kk) ? 1u : 0u);
                if (MC) umma_commit_mc(&w_empty[slot], 0x3);   // the slot is free once BOTH CTAs' MMAs have read it
                else umma_commit(&w_empty[slot]);
                if (++slot == R2P_WSLOTS) { slot = 0; ph ^= 1; }
              }
            umma_commit(&t_full[acc]);
            umma_commit(&b_done[s]);
            if (SD_R2P_TRACE && P.trace != nullptr && blockIdx.x == 0 && job < 32) P.trace[job * 18 + 1] = clock64();
            if (++acc == R2P_ACCS) acc = 0;
            ++job;
          }
      }
    }
  } else {
    // ---------------------------------------------------------------------- epilogue
    const int grp = (warp - 2) >> 3;              // epilogue group: jobs with job % 2 == grp
    const int gw = (warp - 2) & 7;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access = 32 output channels
    const int half = gw >> 2;                     // frames [80 half, 80 half + 80)
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int ch = quarter * 32 + lane;           // this thread's output channel while the data is channel-per-lane
    // after the in-warp transpose lane (tg, tr) owns the channel octet tcol .. tcol + 7 of frames fb + tr, fb + tr + 8
    const int tg = lane >> 3, tr = lane & 7;
    const int tcol = quarter * 32 + 8 * tg;
    const int piece = (quarter & 1) * 4 + tg;     // 16-byte piece of the 128-byte input-buffer row
    const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4;
    uint8_t* const my_stage = ystage + (warp - 2) * 1024;
    bool store_pending = false;                   // lane 0: a TMA store of this warp may still be reading my_stage
    if (lane == 0) tma_prefetch_desc(&P.tmapV);
    const int f_lo = 80 * half;
    const int n_blk = f_lo < T ? (min(T, f_lo + 80) - f_lo + 15) >> 4 : 0;   // 16-frame blocks this warp works through
    float amax = 0.f;
    int job = 0;
    for (int r = 0; r < rounds; ++r) {
      const int nact = min(R2P_SLOTS, n_mine - r * R2P_SLOTS);
      for (int i = 0; i < R2_CONVS; ++i)
        for (int s = 0; s < nact; ++s, ++job) {
          if ((job & 1) != grp) continue;
          const int ai = job % R2P_ACCS;                 // accumulator of this job
          const uint32_t tacc = tlane + ai * R2P_ACC_COLS;
          const int w = window_of(r, s);
          const size_t wrow = static_cast<size_t>(w) * Tp;
          uint8_t* const acol = abuf + s * R2P_A_BYTES + (quarter >> 1) * R2P_A_CHUNK;
          const bool has_next = i + 1 < R2_CONVS;
          const float cb = __ldg(P.bias[i] + ch), csc = __ldg(P.scale[i] + ch), csh = __ldg(P.shift[i] + ch);
          __half* const vrow = P.v + (i + 1) * R2_SUB + tcol + (wrow + H) * ld;        // y_i, this lane's octet
          const bool trc = SD_R2P_TRACE && P.trace != nullptr && blockIdx.x == 0 && job < 32 && lane == 0 && (gw == 0 || gw == 7);
          long long* const tp = P.trace + job * 18 + (gw == 0 ? 2 : 9);
          if (trc) tp[0] = clock64();
          bool x_seen = false;                            // this job's x tile has been waited for
          mbar_wait(&t_full[ai], (job / R2P_ACCS) & 1);   // the accumulator's (job / 3)-th use (both groups use all three)
          tc_fence_after();
          if (trc) tp[1] = clock64();
          for (int k = 0; k < n_blk; ++k) {
            const int fb = f_lo + 16 * k;
            uint32_t acc[16];
            __syncwarp();
            tmem_ld16(tacc + fb, acc);
            tmem_ld_wait();
            // the staging tile of the previous block's TMA store must have been read before it is rewritten
            if (lane == 0 && store_pending) tma_store_wait_read();
            __syncwarp();
            uint32_t wv[8];                              // wv[q] = {frame fb + 2q, frame fb + 2q + 1} of channel ch
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float y0 = fmaf(fmaxf(__uint_as_float(acc[2 * q]) + cb, 0.f), csc, csh);
              const float y1 = fmaf(fmaxf(__uint_as_float(acc[2 * q + 1]) + cb, 0.f), csc, csh);
              if (kTrackOflow) {
                if (fb + 2 * q < T) amax = fmaxf(amax, fabsf(y0));
                if (fb + 2 * q + 1 < T) amax = fmaxf(amax, fabsf(y1));
              }
              wv[q] = pack_half2(y0, y1);
            }
            // 8 channels x 16 frames of every eight lanes, transposed in registers: three xor-shuffle stages per
            // 8-frame half; afterwards lane (tg, tr) holds channels tcol .. tcol+7 of frames fb + tr (z = 0), + 8 (z = 1)
#pragma unroll
            for (int z = 0; z < 2; ++z) {
              uint32_t u[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {               // lanes c, c^1 -> one frame parity each, two channels per word
                const uint32_t own = wv[4 * z + q];
                const uint32_t got = __shfl_xor_sync(0xffffffffu, own, 1);
                u[q] = b0 ? __byte_perm(own, got, 0x3276) : __byte_perm(own, got, 0x5410);
              }
              uint32_t x2[2][2];
#pragma unroll
              for (int j = 0; j < 2; ++j) {               // lanes c, c^2 -> frames = lane mod 4, four channels
                const uint32_t keep = b1 ? u[2 * j + 1] : u[2 * j];
                const uint32_t give = b1 ? u[2 * j] : u[2 * j + 1];
                const uint32_t got = __shfl_xor_sync(0xffffffffu, give, 2);
                x2[j][0] = b1 ? got : keep;
                x2[j][1] = b1 ? keep : got;
              }
              const uint32_t k0 = b2 ? x2[1][0] : x2[0][0], k1 = b2 ? x2[1][1] : x2[0][1];
              const uint32_t g0 = b2 ? x2[0][0] : x2[1][0], g1 = b2 ? x2[0][1] : x2[1][1];
              const uint32_t r0 = __shfl_xor_sync(0xffffffffu, g0, 4), r1 = __shfl_xor_sync(0xffffffffu, g1, 4);
              const uint4 y = b2 ? make_uint4(r0, r1, k0, k1) : make_uint4(k0, k1, r0, r1);   // lanes c, c^4 -> eight channels
              const int f = fb + tr + 8 * z;
              // y_i -> the warp's staging tile [16 frames][64 B]; it leaves as ONE TMA store per block below (per-lane
              // 16-byte global stores cost ~6 k of the epilogue's ~9 k cycles).  Only the few mirrored halo rows the
              // next layer's taps read are stored from the registers.
              *reinterpret_cast<uint4*>(my_stage + (tr + 8 * z) * 64 + tg * 16) = y;
              if (f < T) {
                if (!(SD_R2P_DBG & 1)) {
                  if (f >= 1 && f <= H) *reinterpret_cast<uint4*>(vrow - static_cast<long>(f) * ld) = y;
                  if (f >= T - 1 - H && f <= T - 2) *reinterpret_cast<uint4*>(vrow + static_cast<long>(2 * (T - 1) - f) * ld) = y;
                }
                if (has_next && !(SD_R2P_DBG & 4)) {
                  // next conv's input: y_i added onto the x tile the producer put into the slot's buffer, from the
                  // f16-rounded y_i as an unfused chain would read it back (one f16 add equals
                  // round_f16(float(x) + float(y)) bit for bit).  The mirrored halo rows hold u's own reflect halo,
                  // bit-identical to the frames they mirror, so adding the same y gives the same sums.
                  if (!x_seen) {
                    mbar_wait(&x_full[s], (r * R2_CONVS + i + 1) & 1);
                    x_seen = true;
                  }
                  auto add_y = [&](int p) {
                    uint4* const q4 = reinterpret_cast<uint4*>(acol + p * 128 + ((piece ^ (p & 7)) << 4));
                    uint4 sk = *q4;
                    __half2* sh = reinterpret_cast<__half2*>(&sk);
                    const __half2* yh = reinterpret_cast<const __half2*>(&y);
#pragma unroll
                    for (int e = 0; e < 4; ++e) sh[e] = __hadd2(sh[e], yh[e]);
                    *q4 = sk;
                  };
                  add_y(f + d);
                  if (f >= 1 && f <= d) add_y(d - f);
                  else if (f >= T - 1 - d && f <= T - 2) add_y(d + 2 * (T - 1) - f);
                }
              }
            }
            // the block's 16 frames x 32 channels of y_i: staging tile -> v (frames >= T are clipped by the tensor map)
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && !(SD_R2P_DBG & 1)) {
              tma_store_3d(&P.tmapV, my_stage, (i + 1) * R2_SUB + quarter * 32, fb, w);
              tma_store_commit();
              store_pending = true;
            }
          }
          if (trc) tp[2] = clock64();
          // the accumulator is drained and the next input complete: hand both to the MMA warp
          tc_fence_before();
          if (has_next) fence_proxy_async();  // input-buffer writes -> visible to the tensor core's reads
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&t_empty[ai]);
            if (has_next) mbar_arrive(&a_ready[s]);
          }
          if (trc) tp[3] = clock64();
        }
    }
    if (lane == 0 && store_pending) tma_store_wait_all();   // the staging tile must outlive the last store's read
    if (kTrackOflow && amax > kHalfMax && P.oflow != nullptr) atomicOr(P.oflow, 1);
  }

  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();   // the peer may still multicast into this CTA's ring / arrive on its barriers
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace sd
