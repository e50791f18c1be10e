// res2net_pipe.cuh — the Res2Net chain of one SERes2Net block, FOUR windows in flight per SM.
//
// Same arithmetic and the same dataflow idea as res2net_fused.cuh (one window is carried through the seven
// dependent dilated k = 3 convolutions with its running input x_{i+1} + y_i held in shared memory; reference:
// speechbrain Res2NetBlock reached from /root/reference/speech_encode.py:64-78), but organised around what the
// first version's clock-stamp traces showed (profiles/r01_r2_trace_mode3.txt): a convolution is ~2 k cycles of
// tensor work followed by ~9.5 k cycles of epilogue on the critical path, and with two windows per SM (two CTAs of
// one window each) both pipes idle most of the time — 11.6 k cycles per window-convolution per SM, plus a 14 %
// quantisation loss (512 windows over 296 one-window CTAs).
//
// Here ONE CTA per SM owns four window slots (4 x 40 KB input buffers) and walks the jobs (window slot s,
// convolution i) in a fixed round-robin order  (s0,i) (s1,i) (s2,i) (s3,i) (s0,i+1) ...  :
//   * job j accumulates into TMEM accumulator j & 1 and is finished by epilogue GROUP j & 1 (8 warps each), so the
//     MMAs of job j+1 run under the epilogue of job j, and two epilogues are always in flight;
//   * between two convolutions of the same window lie the jobs of the three other windows, which hides the
//     ~10 k-cycle epilogue latency that used to sit between dependent MMA phases;
//   * 148 CTAs x 4 slots >= 512 windows: the BASELINE batch is one round without a tail.
// Weights stream through the same 3-slot TMA ring (L2-resident, 96 KB per convolution).  y_i leaves through a
// 512-byte per-warp staging tile as 32-byte row segments; the per-channel constants sit in shared memory per
// epilogue group (read straight from global they cost an L2 round trip per use: with 222 KB of shared memory
// carved out there is next to no L1 left — measured, the epilogue went from ~10 k to ~15 k cycles).
// Budget: 4 x 40 + 3 x 16 + 8 + 6 KB of the 227 KB.
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

namespace sd {

constexpr int R2P_SLOTS = 4;                      // windows resident per CTA
constexpr int R2P_GROUPS = 2;                     // epilogue groups (8 warps each)
constexpr int R2P_ACCS = 3;                       // TMEM accumulators (160 columns each): the MMAs of job j+2 do not wait
                                                  // for the epilogue of job j (measured with two: the accumulator hand-back
                                                  // serialised MMA -> epilogue per group, 10.8 k cycles per job)
constexpr int R2P_ACC_COLS = 160;
constexpr int R2P_THREADS = 64 + R2P_GROUPS * 256;   // warp 0 TMA, warp 1 MMA, 2 x 8 epilogue warps
constexpr int R2P_RA = 160;                       // rows of an input buffer: T + 2 * dil <= 160
constexpr int R2P_A_CHUNK = R2P_RA * 128;         // bytes per 64-channel chunk (20 x 1024)
constexpr int R2P_A_BYTES = 2 * R2P_A_CHUNK;
constexpr int R2P_WSLOTS = 3;
constexpr int R2P_YSTAGE = 16 * 512;              // 512 B per epilogue warp: [16 rows][32 B]
constexpr int R2P_CONST = R2P_GROUPS * 2 * 384 * 4;   // per group, double-buffered {bias, scale, shift}[128]
constexpr int R2P_SMEM = R2P_SLOTS * R2P_A_BYTES + R2P_WSLOTS * R2_WBOX + R2P_YSTAGE + R2P_CONST + 256;
static_assert(R2P_A_CHUNK % 1024 == 0, "input chunks must keep the swizzle alignment");
static_assert(R2P_SMEM <= 232448, "exceeds the 227 KB of shared memory a CTA can have");

__global__ void __launch_bounds__(R2P_THREADS, 1)
res2net_pipe_kernel(const __grid_constant__ Res2Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  pdl_trigger();
  uint8_t* const abuf = smem;                                        // [slot][chunk][R2P_RA rows][128 B]
  uint8_t* const wring = smem + R2P_SLOTS * R2P_A_BYTES;
  uint8_t* const ystage = wring + R2P_WSLOTS * R2_WBOX;
  float* const consts = reinterpret_cast<float*>(ystage + R2P_YSTAGE);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(ystage + R2P_YSTAGE + R2P_CONST);
  uint64_t* const w_full = bars;            // [3] weight box landed
  uint64_t* const w_empty = bars + 3;       // [3] its MMAs retired
  uint64_t* const a_full = bars + 6;        // [4] a window's first input landed (TMA)
  uint64_t* const a_ready = bars + 10;      // [4] the epilogue wrote the next convolution's input (8 arrivals)
  uint64_t* const a_free = bars + 14;       // [4] the window's last convolution has read its input
  uint64_t* const t_full = bars + 18;       // [3] accumulator complete
  uint64_t* const t_empty = bars + 21;      // [3] accumulator drained (8 arrivals)
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = P.T, Tp = P.Tp, H = P.H, d = P.dil, ld = P.ld;
  const int RA = T + 2 * d;
  const bool tail = T > 128;                // frames 128.. are computed transposed (see res2net_fused.cuh, MODE 3)
  const int G = gridDim.x;
  const int n_mine = (P.B - static_cast<int>(blockIdx.x) + G - 1) / G;   // windows blockIdx.x, + G, ...
  const int rounds = (n_mine + R2P_SLOTS - 1) / R2P_SLOTS;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < R2P_WSLOTS; ++s) {
        mbar_init(&w_full[s], 1);
        mbar_init(&w_empty[s], 1);
      }
      for (int s = 0; s < R2P_SLOTS; ++s) {
        mbar_init(&a_full[s], 1);
        mbar_init(&a_ready[s], 8);
        mbar_init(&a_free[s], 1);
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
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // window of (round r, slot s): the last windows first — tdnn1 wrote them last, so they are still in L2
  auto window_of = [&](int r, int s) { return P.B - 1 - (static_cast<int>(blockIdx.x) + (r * R2P_SLOTS + s) * G); };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&P.tmapU);
      for (int i = 0; i < R2_CONVS; ++i) tma_prefetch_desc(&P.tmapW[i]);
      int slot = 0;
      uint32_t ph = 0;
      for (int r = 0; r < rounds; ++r) {
        const int nact = min(R2P_SLOTS, n_mine - r * R2P_SLOTS);
        for (int s = 0; s < nact; ++s) {
          if (r > 0) mbar_wait(&a_free[s], (r - 1) & 1);   // the previous window's last convolution is done reading
          mbar_arrive_expect_tx(&a_full[s], static_cast<uint32_t>(2 * RA * 128));
          const int row0 = window_of(r, s) * Tp + H - d;
          tma_load_2d(abuf + s * R2P_A_BYTES, &P.tmapU, &a_full[s], R2_SUB, row0);
          tma_load_2d(abuf + s * R2P_A_BYTES + R2P_A_CHUNK, &P.tmapU, &a_full[s], R2_SUB + 64, row0);
        }
        for (int i = 0; i < R2_CONVS; ++i)
          for (int s = 0; s < nact; ++s)
            for (int kc = 0; kc < 2; ++kc)
              for (int j = 0; j < 3; ++j) {
                mbar_wait(&w_empty[slot], ph ^ 1);
                mbar_arrive_expect_tx(&w_full[slot], R2_WBOX);
                tma_load_2d(wring + slot * R2_WBOX, &P.tmapW[i], &w_full[slot], j * R2_SUB + kc * 64, 0);
                if (++slot == R2P_WSLOTS) { slot = 0; ph ^= 1; }
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
      for (int r = 0; r < rounds; ++r) {
        const int nact = min(R2P_SLOTS, n_mine - r * R2P_SLOTS);
        for (int i = 0; i < R2_CONVS; ++i)
          for (int s = 0; s < nact; ++s, ++job) {
            if (i == 0) mbar_wait(&a_full[s], r & 1);
            else { mbar_wait(&a_ready[s], (ready_ph >> s) & 1u); ready_ph ^= 1u << s; }
            if (job >= R2P_ACCS) { mbar_wait(&t_empty[acc], (empty_ph >> acc) & 1u); empty_ph ^= 1u << acc; }
            tc_fence_after();
            if (P.trace != nullptr && blockIdx.x == 0 && job < 32) P.trace[job * 18 + 0] = clock64();
            const uint32_t d0 = tmem_base + acc * R2P_ACC_COLS;
            const uint32_t a_addr0 = smem_u32(abuf + s * R2P_A_BYTES);
            for (int kc = 0; kc < 2; ++kc)
              for (int j = 0; j < 3; ++j) {
                mbar_wait(&w_full[slot], ph);
                tc_fence_after();
                const uint64_t db = make_smem_desc_sw128(smem_u32(wring + slot * R2_WBOX));
                if (tail) {
                  // frames 128..: "A" = the weight box (128 output channels), "B" = 32 rows of the input
                  const uint64_t dx = make_smem_desc_sw128(a_addr0 + kc * R2P_A_CHUNK + (128 + j * d) * 128);
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk)
                    umma_f16(d0 + 128, db + 2 * kk, dx + 2 * kk, P.idesc_t1, (kc | j | kk) ? 1u : 0u);
                }
                // tap j of frames 0..127: rows j*d .. of the same buffer (row pitch 128 B)
                const uint64_t da = make_smem_desc_sw128(a_addr0 + kc * R2P_A_CHUNK + (j * d) * 128);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_f16(d0, da + 2 * kk, db + 2 * kk, P.idesc, (kc | j | kk) ? 1u : 0u);
                umma_commit(&w_empty[slot]);
                if (++slot == R2P_WSLOTS) { slot = 0; ph ^= 1; }
              }
            umma_commit(&t_full[acc]);
            if (i == R2_CONVS - 1) umma_commit(&a_free[s]);
            if (P.trace != nullptr && blockIdx.x == 0 && job < 32) P.trace[job * 18 + 1] = clock64();
            if (++acc == R2P_ACCS) acc = 0;
          }
      }
    }
  } else {
    // ---------------------------------------------------------------------- epilogue
    const int grp = (warp - 2) >> 3;              // epilogue group = accumulator
    const int gw = (warp - 2) & 7;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = gw >> 2;                     // which 64 of the 128 output channels
    uint8_t* const my_stage = ystage + (warp - 2) * 512;
    const int gt = threadIdx.x - 64 - grp * 256;  // thread within the group
    int njob = 0;                                 // jobs this group has started (constant-buffer parity)
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t full_ph = 0;                         // phase bits of t_full[a]
    const int nk = quarter * 32 < T ? 2 : 0;      // 32-column chunks of frames 0..127 this warp works through
    // this warp's share of the transposed tile: channel quarter*32 + lane, frames 128 + half*h0 .. (split evenly)
    const int t1_h0 = (T - 128 + 1) >> 1;
    const int t1_f0 = 128 + half * t1_h0;
    const int t1_end = half == 0 ? 128 + t1_h0 : T;
    const int t1_c = quarter * 32 + lane;
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
          uint8_t* const ab = abuf + s * R2P_A_BYTES;
          const bool has_next = i + 1 < R2_CONVS;
          // this job's {bias, scale, shift} -> the group's other constant buffer (its last readers were two jobs
          // ago and every warp of the group has passed the previous job's barrier since)
          float* const cst = consts + (grp * 2 + (njob & 1)) * 384;
          ++njob;
          {
            const int which = gt >> 7, c = gt & 127;
            cst[gt] = __ldg((which == 0 ? P.bias[i] : P.scale[i]) + c);
            if (gt < 128) cst[256 + gt] = __ldg(P.shift[i] + gt);
          }
          const float* const cb_ = cst;
          const float* const cs_ = cst + 128;
          const float* const ch_ = cst + 256;
          const int out_col = (i + 1) * R2_SUB + half * 64;  // this warp's columns of v
          const __half* xnext = P.u + (i + 2) * R2_SUB + half * 64 + (wrow + H) * ld;
          uint4 xc[4];
          auto load_xc = [&](int k) {
            const int t = quarter * 32 + lane;
            if (t < T) {
              const uint4* a4 = reinterpret_cast<const uint4*>(xnext + static_cast<size_t>(t) * ld + (k & 1) * 32);
#pragma unroll
              for (int q = 0; q < 4; ++q) xc[q] = __ldg(a4 + q);
            }
          };
          if (has_next && nk > 0) load_xc(0);      // requested before the accumulator wait
          uint32_t xt[8];   // x_{i+1}[frame, this channel] for the 16 transposed frames, two f16 per register
          if (tail && has_next) {
            const unsigned short* xs = reinterpret_cast<const unsigned short*>(xnext) + t1_c - half * 64;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const int t = t1_f0 + 2 * jj;
              const uint32_t lo = t < t1_end ? __ldg(xs + static_cast<size_t>(t) * ld) : 0u;
              const uint32_t hi = t + 1 < t1_end ? __ldg(xs + static_cast<size_t>(t + 1) * ld) : 0u;
              xt[jj] = lo | (hi << 16);
            }
          }
          asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory");   // the group's constants are in place
          const bool tr = P.trace != nullptr && blockIdx.x == 0 && job < 32 && lane == 0 && (gw == 0 || gw == 7);
          long long* const tp = P.trace + job * 18 + (gw == 0 ? 2 : 8);
          if (tr) tp[0] = clock64();
          mbar_wait(&t_full[ai], (full_ph >> ai) & 1u);
          full_ph ^= 1u << ai;
          tc_fence_after();
          if (tr) tp[1] = clock64();
          if (tail) {
            uint32_t acc[16];
            __syncwarp();
            tmem_ld16(tacc + 128 + (t1_f0 - 128), acc);
            tmem_ld_wait();
            const float cb = cb_[t1_c], csc = cs_[t1_c], csh = ch_[t1_c];
            __half* const vcol = P.v + (i + 1) * R2_SUB + t1_c + (wrow + H) * ld;   // this channel's column of y_i
            uint8_t* const acol = ab + (t1_c >> 6) * R2P_A_CHUNK + (t1_c & 7) * 2;
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
          }
          if (tr) tp[2] = clock64();
          if (nk == 0) {   // a warp without rows of frames 0..127 (short windows) still owes its arrivals
            tc_fence_before();
            if (has_next) fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(&t_empty[ai]);
              if (has_next) mbar_arrive(&a_ready[s]);
            }
          }
          for (int k = 0; k < nk; ++k) {
            const int ci = k & 1;
            const int tw0 = quarter * 32;        // first frame of this warp's 32 rows
            const int t = tw0 + lane;
            const bool valid = t < T;
            const int c0 = half * 64 + ci * 32;  // column within the conv's 128 outputs
            uint32_t acc[32];
            __syncwarp();
            tmem_ld32(tacc + c0, acc);
            tmem_ld_wait();
            const float4* b4 = reinterpret_cast<const float4*>(cb_ + c0);
            const float4* s4 = reinterpret_cast<const float4*>(cs_ + c0);
            const float4* h4 = reinterpret_cast<const float4*>(ch_ + c0);
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
            // next conv's input x_{i+1} + y_i, from the f16-rounded y_i as an unfused chain would read it back
            // (one f16 add equals round_f16(float(x) + float(y)) bit for bit)
            if (has_next && valid) {
              const int p = t + d;
              int p2 = -1;
              if (t >= 1 && t <= d) p2 = d - t;
              else if (t >= T - 1 - d && t <= T - 2) p2 = d + 2 * (T - 1) - t;
              uint8_t* arow = ab + (c0 >> 6) * R2P_A_CHUNK;
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
              // last chunk: the accumulator is drained and the next input complete — release the MMA warp now; the
              // write-out below only touches the staging tile and global memory
              tc_fence_before();
              if (has_next) fence_proxy_async();  // input-buffer writes -> visible to the tensor core's reads
              __syncwarp();
              if (lane == 0) {
                mbar_arrive(&t_empty[ai]);
                if (has_next) mbar_arrive(&a_ready[s]);
              }
              if (tr) tp[4] = clock64();
            }
            if (k == 0 && tr) tp[3] = clock64();
            if (has_next && k + 1 < nk) load_xc(k + 1);  // xc is free again: the next chunk's x_{i+1}
            // y_i -> v through the 512-byte staging tile ([16 rows][32 B], 16-byte pieces XOR-swizzled so both sides
            // are conflict-free): four passes (16 channels x 16 rows), each leaving as 32-byte row segments, 16 rows
            // per store instruction
            __half* const dst0 = P.v + out_col + ci * 32 + (wrow + H) * ld;
#pragma unroll
            for (int hp = 0; hp < 2; ++hp) {
#pragma unroll
              for (int rh = 0; rh < 2; ++rh) {
                __syncwarp();
                if ((lane >> 4) == rh) {
                  const int rl = lane & 15;
                  uint8_t* srow = my_stage + rl * 32;
                  const int sw = (rl >> 2) & 1;
                  *reinterpret_cast<uint4*>(srow + ((0 ^ sw) << 4)) = pk[2 * hp];
                  *reinterpret_cast<uint4*>(srow + ((1 ^ sw) << 4)) = pk[2 * hp + 1];
                }
                __syncwarp();
                const int piece = lane & 1, rsub = lane >> 1;
                const uint4 val = *reinterpret_cast<const uint4*>(my_stage + rsub * 32 + ((piece ^ ((rsub >> 2) & 1)) << 4));
                __half* const dst = dst0 + hp * 16 + piece * 8;
                const int tt = tw0 + rh * 16 + rsub;
                if (tt < T) {
                  *reinterpret_cast<uint4*>(dst + static_cast<long>(tt) * ld) = val;
                  if (tt >= 1 && tt <= H) *reinterpret_cast<uint4*>(dst - static_cast<long>(tt) * ld) = val;
                  if (tt >= T - 1 - H && tt <= T - 2)
                    *reinterpret_cast<uint4*>(dst + static_cast<long>(2 * (T - 1) - tt) * ld) = val;
                }
              }
            }
          }
          if (tr) tp[5] = clock64();
        }
    }
    if (kTrackOflow && amax > kHalfMax && P.oflow != nullptr) atomicOr(P.oflow, 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace sd
