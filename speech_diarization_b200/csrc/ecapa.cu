// ecapa.cu — the ECAPA-TDNN (C = 1024) embedding plan: weight repacking, activation
// workspace, the cached per-shape launch programme, and the forward pass.
//
// Replaces the model behind using_ecapa_encoder / EncoderClassifier.encode_batch
// (/root/reference/speech_encode.py:64-78, ecapa_annote.py:6-22, diar_diag.py:161-170);
// the arithmetic is speechbrain's ECAPA_TDNN.
//
// Data layout in HBM: every activation is f16, channels-last, [B*Tp, C] with
//   Tp = roundup(T + 2H, 16), H = 4 halo rows on each side of an utterance's T frames
//   holding the REFLECT padding of the dilated convolutions (max pad = dilation 4).
// A convolution tap is then a row shift of the TMA box, a 1x1 convolution a plain GEMM,
// and pointwise layers keep the halo valid for free (they commute with reflection).
//
// Forward programme for one batch (all launches on the caller's stream, no host sync):
//   fbank (2 kernels) -> block0 GEMM(k5) -> 3 x [tdnn1 GEMM, 7 x Res2Net GEMM(k3, fused
//   "x_{i+1}+y_i"), tdnn2 GEMM, SE mean, SE MLP, SE scale+residual] -> MFA GEMM ->
//   ASP stats -> context bias -> attention GEMM(tanh) -> pooling GEMM (softmax + weighted
//   mean/std fused in the epilogue) -> BN-folded FC -> optional L2 norm.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <chrono>
#include <map>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>
#include "ecapa_kernels.cuh"
#include "fbank.cuh"
#include "gemm_host.cuh"
#include "host_stage.cuh"
#include "res2net_fused.cuh"
#include "res2net_pipe.cuh"
#include "sd_status.h"

using namespace sd;

namespace {

constexpr int C1 = 1024;   // trunk channels
constexpr int C3 = 3072;   // MFA / ASP channels
constexpr int ATT = 128;   // attention channels
constexpr int SE = 128;    // squeeze-excitation bottleneck
constexpr int SUB = 128;   // Res2Net sub-band width (C1 / scale 8)
constexpr int EMB = 192;
constexpr int FEAT_P = 128; // 80 mel channels padded to two 64-element K chunks
constexpr int HALO = 4;
constexpr int KSPLIT_MAX = 32;  // split-K factor of the two per-utterance dense layers (context bias, FC): SdEcapaPlan::ksplit

struct TdnnW {       // one TDNNBlock: conv weight (f16, [Cout, taps*CinP]) + folded epilogue constants
  __half* W = nullptr;
  float* bias = nullptr;
  float* scale = nullptr;
  float* shift = nullptr;
};

struct BlockW {
  TdnnW tdnn1, res[7], tdnn2;
  __half *se_w1h = nullptr, *se_w2th = nullptr;   // SE weights as f16: conv1 [SE][C1], conv2 transposed [SE][C1]
  float *se_b1 = nullptr, *se_b2 = nullptr;
  int dil = 2;
};

struct Program {     // launch parameters for one (B, T) shape
  int B = 0, T = 0, Tp = 0;
  long rows = 0;
  GemmParams block0, tdnn1[3], res[3][7], resc[3][7], tdnn2[3], mfa, ctx, att, pool, fc;  // resc: EPI_CONV3 form of res
  GemmParams* chain_dev = nullptr;  // device copy of [3][9]: tdnn1, 7 x Res2Net, tdnn2 per block
  Res2Params r2[3];                 // the 7 Res2Net convs of a block as one launch (res2net_fused.cuh)
  bool r2_ok = false;               // shape fits the fused kernel (16 <= T, T + 2*dil <= 168)
  bool colsum_ok = false;           // SE / ASP time statistics come out of the GEMM write-outs
  int cs_group = 128;               // rows per statistics group: gcd(128, Tp)
  cudaGraphExec_t graph = nullptr;  // captured trunk (block0 .. FC) for this shape
  int graph_launches = 0;
  int runs = 0;
  // sd_ecapa_embed_host pipeline: block0 -> b1.tdnn1 -> b1.res2net -> b1.tdnn2 (everything before the first
  // whole-window reduction, the SE squeeze) per upload chunk of B / NFRONT windows, so the front of chunk c runs
  // while chunk c+1 is still crossing PCIe; the rest of the trunk ("tail") follows once for the whole batch.
  static constexpr int NFRONT = 4;
  bool front_built = false, front_ok = false;
  GemmParams f_block0[NFRONT], f_tdnn1[NFRONT], f_tdnn2[NFRONT];
  Res2Params f_r2[NFRONT];
  cudaGraphExec_t graph_tail = nullptr;  // captured tail (b1.se .. FC)
  int graph_tail_launches = 0;
  int runs_tail = 0;
};

}  // namespace

struct SdEcapaPlan {
  int max_batch = 0, max_samples = 0;
  long max_rows = 0;
  std::vector<void*> allocs;
  // weights
  TdnnW w0;
  BlockW blk[3];
  TdnnW wmfa, watt;
  __half* Wams = nullptr; // asp.tdnn weight, mean|std columns [ATT, 2*C3] f16
  __half* Wa2 = nullptr;  // asp.conv weight [C3, ATT] f16
  __half* Wfc = nullptr;  // fc with asp_bn folded in, [EMB, 2*C3] f16
  float* bfc = nullptr;
  // activations
  __half *feats = nullptr, *x0 = nullptr, *cat = nullptr, *u = nullptr, *v = nullptr, *w = nullptr;
  __half *s[2] = {nullptr, nullptr}, *h = nullptr, *attn = nullptr;
  float *raw = nullptr, *raw2 = nullptr, *se_mean = nullptr, *se_hid = nullptr, *se_scale = nullptr, *stats = nullptr;
  float *cs_se = nullptr, *cs_mfa = nullptr, *cq_mfa = nullptr;  // per (m block, window slot) column sums from the GEMM write-outs
  float *uttbias = nullptr, *pooled = nullptr, *emb_tmp = nullptr, *ctx_part = nullptr;
  int* oflow = nullptr;       // [0] an activation of the current forward left the f16 range (reset per forward),
                              // [1] sticky copy for the host (sd_ecapa_overflow / sd_ecapa_embed_host)
  int* oflow_host = nullptr;  // page-locked landing word for [1]
  __half *stats_h = nullptr, *pooled_h = nullptr;
  std::map<std::pair<int, int>, Program> programs;
  Program* last = nullptr;
  // optional per-stage timing (CUDA events on the caller's stream; bench.py's roofline leg)
  bool profile = false;
  // SD_ECAPA_CHAIN=1 runs each block's tdnn1 -> 7 x Res2Net -> tdnn2 as ONE cooperative launch with grid
  // barriers between the steps.  Measured slower than nine stream-ordered launches (0.735 vs 0.623 ms per
  // block at B=512: the per-step pipeline fill/drain costs more than the launch gaps), so it is off.
  bool use_chain = false;
  // SD_ECAPA_MC=1: 256-wide GEMMs on 2-CTA clusters with the weight tile multicast.  Correct, but only ~3% faster
  // on the MFA layer and neutral elsewhere (the GEMMs are bound by per-SM operand ingest, not by L2), so off.
  bool use_mc = false;
  bool use_2sm = true;     // SD_ECAPA_2SM=0: 256-wide GEMMs with cta_group::1 instead of CTA pairs
  bool use_r2fused = true; // SD_ECAPA_R2FUSED=0: Res2Net chain as 7 launches per block instead of one
  bool use_r2pipe = true;  // SD_R2_PIPE=0: res2net_fused_kernel (one window per CTA) instead of res2net_pipe_kernel
  bool use_pdl = false;    // SD_ECAPA_PDL=1: programmatic dependent launch between the trunk's kernels (measured
                           // 2 % SLOWER inside the replayed graph: 3.70-3.74 vs 3.64-3.66 ms per step)
  bool use_tma_out_reflect = true;  // SD_ECAPA_TMAOUT0=0: block0 writes its tile with thread stores like the other layers
  bool use_tma_out = false; // SD_ECAPA_TMAOUT=1: the cta_group::2 GEMMs of the pointwise layers hand their staged tile to TMA
                            // tensor stores.  Measured 3 % SLOWER than the per-thread write-out (tdnn1 0.144 -> 0.148 ms): the
                            // launch is bound by operand ingest next to the store traffic, not by the epilogue threads' time
  bool use_l2_order = true; // SD_ECAPA_L2ORDER=0: se_apply and the attention GEMM walk the rows in ascending order like
                            // their producers (tdnn2, MFA) instead of starting with the rows those left in L2
  int ksplit = 8;          // SD_ECAPA_KSPLIT (1..32, a divisor of 96 keeps the splits even)
  bool use_colsum = true;  // SD_ECAPA_COLSUM=0: separate passes over the activations for the SE mean and ASP mean/std
  bool use_conv3 = true;   // SD_ECAPA_CONV3=0: Res2Net convs through the generic tap-per-k-iteration path
  bool use_graph = true;   // SD_ECAPA_GRAPH=0 disables CUDA-graph replay of the trunk
  cudaStream_t cap_stream = nullptr;
  // sd_ecapa_embed_host: device staging of the caller's host audio, a copy stream and per-chunk events so the
  // host->device copy of chunk c+1 runs under the fbank kernels of chunk c
  float* h2d_buf = nullptr;
  size_t h2d_cap = 0;       // samples
  float* emb_stage = nullptr;
  float* emb_pinned = nullptr;  // page-locked landing buffer: a device->pageable copy is staged by the driver and blocks
  int emb_cap = 0;          // windows
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_ev[8] = {};
  cudaEvent_t start_ev = nullptr;
  StagingRing* ring = nullptr;   // pinned staging ring + copy threads for pageable callers (host_stage.cuh)
  bool ring_failed = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  int forwards_profiled = 0;
};

namespace {

int dev_alloc(SdEcapaPlan* p, void** out, size_t bytes, bool zero) {
  void* d = nullptr;
  if (cudaMalloc(&d, bytes) != cudaSuccess) {
    cudaGetLastError();
    return fail(SD_ERR_NOMEM, "cudaMalloc(%zu bytes) failed", bytes);
  }
  p->allocs.push_back(d);
  if (zero) SD_CUDA_OK(cudaMemset(d, 0, bytes));
  *out = d;
  return SD_OK;
}

template <typename T>
int upload(SdEcapaPlan* p, T** out, const std::vector<T>& host) {
  void* d = nullptr;
  SD_TRY(dev_alloc(p, &d, host.size() * sizeof(T), false));
  SD_CUDA_OK(cudaMemcpy(d, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = static_cast<T*>(d);
  return SD_OK;
}

struct StateDict {
  std::unordered_map<std::string, std::pair<const float*, int64_t>> m;
  int get(const std::string& key, int64_t numel, const float** out) const {
    auto it = m.find(key);
    if (it == m.end()) return fail(SD_ERR_MISSING, "state dict has no tensor '%s'", key.c_str());
    if (it->second.second != numel)
      return fail(SD_ERR_MISSING, "tensor '%s' has %lld elements, expected %lld", key.c_str(),
                  (long long)it->second.second, (long long)numel);
    *out = it->second.first;
    return SD_OK;
  }
};

// speechbrain TDNNBlock "<prefix>.conv.conv.{weight,bias}" + "<prefix>.norm.norm.*" ->
// f16 weight [cout, taps*cin_p] (tap-major K) and bias / BN scale / BN shift vectors.
// w_cols restricts the repack to the first w_cols input channels (ASP: x part only).
int load_tdnn(SdEcapaPlan* p, const StateDict& sd, const std::string& prefix, int cout, int cin,
              int taps, int cin_p, int w_cols, TdnnW* out) {
  const float *W = nullptr, *b = nullptr, *g = nullptr, *be = nullptr, *rm = nullptr, *rv = nullptr;
  SD_TRY(sd.get(prefix + ".conv.conv.weight", (int64_t)cout * cin * taps, &W));
  SD_TRY(sd.get(prefix + ".conv.conv.bias", cout, &b));
  SD_TRY(sd.get(prefix + ".norm.norm.weight", cout, &g));
  SD_TRY(sd.get(prefix + ".norm.norm.bias", cout, &be));
  SD_TRY(sd.get(prefix + ".norm.norm.running_mean", cout, &rm));
  SD_TRY(sd.get(prefix + ".norm.norm.running_var", cout, &rv));
  std::vector<__half> wh((size_t)cout * taps * cin_p, __float2half(0.f));
  for (int o = 0; o < cout; ++o)
    for (int c = 0; c < w_cols; ++c)
      for (int j = 0; j < taps; ++j)
        wh[((size_t)o * taps + j) * cin_p + c] = __float2half_rn(W[((size_t)o * cin + c) * taps + j]);
  std::vector<float> bias(b, b + cout), scale(cout), shift(cout);
  for (int o = 0; o < cout; ++o) {
    const double sc = (double)g[o] / sqrt((double)rv[o] + 1e-5);
    scale[o] = (float)sc;
    shift[o] = (float)((double)be[o] - (double)rm[o] * sc);
  }
  SD_TRY(upload(p, &out->W, wh));
  SD_TRY(upload(p, &out->bias, bias));
  SD_TRY(upload(p, &out->scale, scale));
  SD_TRY(upload(p, &out->shift, shift));
  return SD_OK;
}

inline int tp_of(int T) { return ((T + 2 * HALO + 15) / 16) * 16; }

// Fills the common fields of a TDNN-epilogue GEMM over `rows` activation rows.
int setup_tdnn_gemm(GemmParams& P, const __half* A, long rows, int a_cols, int ld_a,
                    const TdnnW& W, int cout, int k_total, int n_tile, int cin_p, int taps,
                    int dil, int a_col0, const Program& pr, void* out, int ld_out, int out_col0,
                    int flags, bool mc = false) {
  init_params(P);
  SD_TRY(make_tmap_f16(&P.tmapA, A, rows, a_cols, ld_a, BM));
  // mc: launched on 2-CTA clusters, each CTA fetches (and multicasts) half of the B tile
  SD_TRY(make_tmap_f16(&P.tmapB, W.W, cout, k_total, k_total, mc ? n_tile / 2 : n_tile));
  P.num_m_blocks = (int)((rows + BM - 1) / BM);
  P.num_n_blocks = cout / n_tile;
  P.n_tile = n_tile;
  P.idesc = make_idesc_f16(n_tile, 0);
  int ki = 0;  // (the cta_group::2 launch re-encodes idesc with M = 256, see build_program)
  for (int j = 0; j < taps; ++j)
    for (int c = 0; c < cin_p / BK; ++c, ++ki) {
      P.kit[ki].a_col = a_col0 + c * BK;
      P.kit[ki].a_row_off = (j - taps / 2) * dil;
      P.kit[ki].b_col = j * cin_p + c * BK;
      P.kit[ki].slot = 0;
      P.kit[ki].accum = ki > 0;
    }
  P.num_kiters = ki;
  EpiParams& E = P.epi;
  E.flags = flags;
#if SD_EXPERIMENTS
  if (const char* e = getenv("SD_ECAPA_APF")) P.a_prefetch = atoi(e) < 0 ? 0 : atoi(e) > 64 ? 64 : atoi(e);
  if (const char* e = getenv("SD_DEBUG_EPI")) E.flags |= (atoi(e) == 1 ? 64 : atoi(e) == 2 ? 128 : atoi(e) == 3 ? 256 : 0);
#endif
  E.M_rows = (int)rows;
  E.N_cols = cout;
  E.Tp = pr.Tp;
  E.T = pr.T;
  E.H = HALO;
  E.out = out;
  E.ld_out = ld_out;
  E.out_col_off = out_col0;
  E.bias = W.bias;
  E.scale = W.scale;
  E.shift = W.shift;
  E.cs_group = pr.cs_group;
  return SD_OK;
}

// Pointwise layer on the cta_group::2 kernel: hand the staged tile to TMA tensor stores (EF_TMA_OUT).  The maps
// cover exactly the columns this layer owns, so clipping at the tensor bounds replaces the row / column guards.
int enable_tma_out(SdEcapaPlan* p, GemmParams& P) {
  if (!p->use_2sm || P.n_tile != 256) return SD_OK;
  EpiParams& E = P.epi;
  if (E.flags & EF_REFLECT) {
    // block0 (k = 5): the one 256-wide layer whose EPILOGUE sets the pace (10 k-iterations per tile), so freeing
    // its threads from the 64 KB write-out pays.  The halo rows are mirrored inside the staging tile, which needs
    // every window's halo rows and their mirror sources in the same 128-row tile.
    if (!p->use_tma_out_reflect || E.out2 != nullptr || E.sum_out != nullptr || E.colsum != nullptr) return SD_OK;
    const int H = E.H, T = E.T, Tp = E.Tp;
    if (Tp % EPI_WARPS != 0) return SD_OK;
    for (long b = 0; b < 128 && b * Tp < E.M_rows; ++b) {   // the (window start mod 128) pattern repeats within 128 windows
      const long s0 = b * Tp;
      if (s0 / BM != (s0 + 2 * H) / BM) return SD_OK;                       // rows t = -H .. H
      if ((s0 + H + T - 1 - H) / BM != (s0 + H + T - 1 + H) / BM) return SD_OK;   // rows t = T-1-H .. T-1+H
    }
  } else if (!p->use_tma_out) {
    return SD_OK;
  }
  SD_TRY(make_tmap_f16(&P.tmapH, static_cast<__half*>(E.out) + E.out_col_off, E.M_rows, E.N_cols, E.ld_out, BM));
  if (E.out2 != nullptr) SD_TRY(make_tmap_f16(&P.tmapO2, E.out2, E.M_rows, E.out2_cols, E.ld_out2, BM));
  E.flags |= EF_TMA_OUT;
  return SD_OK;
}

int build_program(SdEcapaPlan* p, int B, int T, Program** out) {
  auto key = std::make_pair(B, T);
  auto it = p->programs.find(key);
  if (it != p->programs.end()) {
    *out = &it->second;
    return SD_OK;
  }
  if (p->programs.size() > 256) {   // variable-length callers (embed_segments) produce many (B, T) shapes
    cudaDeviceSynchronize();  // launches that reference the cached descriptors may still be in flight
    for (auto& kv : p->programs) {
      if (kv.second.graph) cudaGraphExecDestroy(kv.second.graph);
      if (kv.second.graph_tail) cudaGraphExecDestroy(kv.second.graph_tail);
    }
    p->programs.clear();
    p->last = nullptr;
  }
  Program pr;
  pr.B = B;
  pr.T = T;
  pr.Tp = tp_of(T);
  pr.rows = (long)B * pr.Tp;
  pr.cs_group = 128;
  while (pr.Tp % pr.cs_group) pr.cs_group >>= 1;   // gcd(128, Tp); Tp is a multiple of 16
  const long R = pr.rows;
  // block0: k = 5 over the 128-padded mel channels
  SD_TRY(setup_tdnn_gemm(pr.block0, p->feats, R, FEAT_P, FEAT_P, p->w0, C1, 5 * FEAT_P, 256, FEAT_P,
                         5, 1, 0, pr, p->x0, C1, 0, EF_REFLECT, p->use_mc || p->use_2sm));
  SD_TRY(enable_tma_out(p, pr.block0));
  for (int b = 0; b < 3; ++b) {
    const __half* in = b == 0 ? p->x0 : p->cat + (size_t)(b - 1) * C1;
    const int ld_in = b == 0 ? C1 : C3;
    const BlockW& bw = p->blk[b];
    // tdnn1: 1x1, also copies sub-band 0 into v (Res2Net passes it through)
    SD_TRY(setup_tdnn_gemm(pr.tdnn1[b], in, R, C1, ld_in, bw.tdnn1, C1, C1, 256, C1, 1, 1, 0, pr,
                           p->u, C1, 0, 0, p->use_mc || p->use_2sm));
    pr.tdnn1[b].epi.out2 = p->v;
    pr.tdnn1[b].epi.ld_out2 = C1;
    pr.tdnn1[b].epi.out2_cols = SUB;
    SD_TRY(enable_tma_out(p, pr.tdnn1[b]));
    // Res2Net chain: y_i = TDNN_i(x_i + y_{i-1}), i = 1..7 (y_0 = x_0 passes through)
    for (int i = 1; i <= 7; ++i) {
      const __half* A = i == 1 ? p->u : p->s[i & 1];
      const int a_cols = i == 1 ? C1 : SUB;
      const int a_col0 = i == 1 ? SUB : 0;
      GemmParams& G = pr.res[b][i - 1];
      SD_TRY(setup_tdnn_gemm(G, A, R, a_cols, a_cols, bw.res[i - 1], SUB, 3 * SUB, 128, SUB, 3,
                             bw.dil, a_col0, pr, p->v, C1, i * SUB, EF_REFLECT));
      if (i < 7) {
        G.epi.add_src = p->u;
        G.epi.ld_add = C1;
        G.epi.add_col_off = (i + 1) * SUB;
        G.epi.sum_out = p->s[(i + 1) & 1];
        G.epi.ld_sum = SUB;
      }
      // the same convolution on the fused-tap / resident-weights operand path (EPI_CONV3)
      GemmParams& Cv = pr.resc[b][i - 1];
      Cv = G;
      SD_TRY(make_tmap_f16(&Cv.tmapA, A, R, a_cols, a_cols, BM + 2 * bw.dil));
      Cv.num_kiters = SUB / BK;
      for (int kc = 0; kc < SUB / BK; ++kc) Cv.kit[kc].a_col = a_col0 + kc * BK;
      Cv.conv_taps = 3;
      Cv.conv_dil = bw.dil;
      Cv.conv_cin = SUB;
    }
    pr.r2_ok = T >= 16 && T + 2 * 4 <= R2_RA_MAX;
    if (pr.r2_ok) {
      Res2Params& Q = pr.r2[b];
      memset(&Q, 0, sizeof(Q));
      SD_TRY(make_tmap_f16(&Q.tmapU, p->u, R, C1, C1, T + 2 * bw.dil));
      SD_TRY(make_tmap_f16_interior_plain(&Q.tmapV, p->v, C1, pr.Tp, T, HALO, B, 32, 16));
      for (int i = 0; i < 7; ++i) {
        SD_TRY(make_tmap_f16(&Q.tmapW[i], bw.res[i].W, SUB, 3 * SUB, 3 * SUB, SUB));
        SD_TRY(make_tmap_f16(&Q.tmapWh[i], bw.res[i].W, SUB, 3 * SUB, 3 * SUB, SUB / 2));
        Q.bias[i] = bw.res[i].bias;
        Q.scale[i] = bw.res[i].scale;
        Q.shift[i] = bw.res[i].shift;
      }
      Q.u = p->u;
      Q.v = p->v;
      Q.ld = C1;
      Q.B = B;
      Q.T = T;
      Q.Tp = pr.Tp;
      Q.H = HALO;
      Q.dil = bw.dil;
      Q.idesc = make_idesc_f16(SUB, 0);
      Q.idesc_t1 = make_idesc_f16(32, 0);
      Q.oflow = p->oflow;
    }
    SD_TRY(setup_tdnn_gemm(pr.tdnn2[b], p->v, R, C1, C1, bw.tdnn2, C1, C1, 256, C1, 1, 1, 0, pr,
                           p->w, C1, 0, 0, p->use_mc || p->use_2sm));
    pr.colsum_ok = p->use_colsum;
    if (pr.colsum_ok) pr.tdnn2[b].epi.colsum = p->cs_se;
    SD_TRY(enable_tma_out(p, pr.tdnn2[b]));
  }
  SD_TRY(setup_tdnn_gemm(pr.mfa, p->cat, R, C3, C3, p->wmfa, C3, C3, 256, C3, 1, 1, 0, pr, p->h, C3,
                         0, 0, p->use_mc || p->use_2sm));
  if (pr.colsum_ok) {
    pr.mfa.epi.colsum = p->cs_mfa;
    pr.mfa.epi.colsq = p->cq_mfa;
  }
  SD_TRY(enable_tma_out(p, pr.mfa));
  SD_TRY(setup_tdnn_gemm(pr.att, p->h, R, C3, C3, p->watt, ATT, C3, 128, C3, 1, 1, 0, pr, p->attn,
                         ATT, 0, 0));
  pr.att.epi.utt_bias = p->uttbias;
  pr.att.m_reverse = p->use_l2_order ? 1 : 0;  // MFA wrote h in ascending row order
  // context bias: uttbias[b, :] = W_{mean|std} . stats[b]   (per-utterance dense layer, M = B rows)
  {
    GemmParams& P = pr.ctx;
    init_params(P);
    SD_TRY(make_tmap_f16(&P.tmapA, p->stats_h, B, 2 * C3, 2 * C3, BM));
    SD_TRY(make_tmap_f16(&P.tmapB, p->Wams, ATT, 2 * C3, 2 * C3, ATT));
    P.num_m_blocks = (B + BM - 1) / BM;
    P.num_n_blocks = 1;
    P.n_tile = ATT;
    P.idesc = make_idesc_f16(ATT, 0);
    P.num_kiters = 2 * C3 / BK;
    for (int c = 0; c < P.num_kiters; ++c) {
      P.kit[c].a_col = c * BK;
      P.kit[c].b_col = c * BK;
      P.kit[c].accum = c > 0;
    }
    P.epi.M_rows = B;
    P.epi.N_cols = ATT;
    P.epi.out = p->ctx_part;
    P.epi.ld_out = ATT;
    P.k_splits = p->ksplit;   // 96 k-iterations over ksplit CTAs per output tile
    P.split_stride = (long)B * ATT;
  }
  // final FC (asp_bn folded): emb[b, :] = Wfc . pooled[b] + bfc
  {
    GemmParams& P = pr.fc;
    init_params(P);
    SD_TRY(make_tmap_f16(&P.tmapA, p->pooled_h, B, 2 * C3, 2 * C3, BM));
    SD_TRY(make_tmap_f16(&P.tmapB, p->Wfc, EMB, 2 * C3, 2 * C3, EMB));
    P.num_m_blocks = (B + BM - 1) / BM;
    P.num_n_blocks = 1;
    P.n_tile = EMB;
    P.idesc = make_idesc_f16(EMB, 0);
    P.num_kiters = 2 * C3 / BK;
    for (int c = 0; c < P.num_kiters; ++c) {
      P.kit[c].a_col = c * BK;
      P.kit[c].b_col = c * BK;
      P.kit[c].accum = c > 0;
    }
    P.epi.M_rows = B;
    P.epi.N_cols = EMB;
    P.epi.ld_out = EMB;
    P.epi.bias = p->bfc;
    P.epi.out = p->emb_tmp;   // [KSPLIT][B][EMB] partial sums, reduced by fc_finish_kernel
    P.k_splits = p->ksplit;
    P.split_stride = (long)B * EMB;
  }
  // pooling GEMM: rows = channels of asp.conv, columns = the Tp rows of one utterance
  {
    GemmParams& P = pr.pool;
    init_params(P);
    // frames per chunk: the whole utterance when it fits one UMMA N (<= 256), else 256-row chunks
    // with the softmax statistics carried across chunks (online softmax)
    const int chunk = pr.Tp <= 256 ? pr.Tp : 256;
    SD_TRY(make_tmap_f16(&P.tmapA, p->Wa2, C3, ATT, ATT, BM));
    SD_TRY(make_tmap_f16(&P.tmapB, p->attn, R, ATT, ATT, chunk));
    SD_TRY(make_tmap_f16(&P.tmapH, p->h, R, C3, C3, chunk));
    P.num_m_blocks = C3 / BM;
    P.num_n_blocks = B;
    P.n_tile = chunk;
    P.n_sub = (pr.Tp + chunk - 1) / chunk;
    P.b_row_stride = pr.Tp;
    P.idesc = make_idesc_f16(chunk, 0);
    for (int c = 0; c < ATT / BK; ++c) {
      P.kit[c].a_col = c * BK;
      P.kit[c].b_col = c * BK;
      P.kit[c].accum = c > 0;
    }
    P.num_kiters = ATT / BK;
    P.epi.M_rows = C3;
    P.epi.N_cols = (int)R;
    P.epi.Tp = pr.Tp;
    P.epi.T = T;
    P.epi.H = HALO;
    P.epi.h = p->h;
    P.epi.ld_h = C3;
    P.epi.gmean = p->stats;   // [B, 2*C3]: mean | std
    P.epi.ld_gmean = 2 * C3;
    P.epi.pooled = p->pooled;
    P.epi.pooled_h = p->pooled_h;
    P.epi.C = C3;
  }
  {
    GemmParams* all[] = {&pr.block0, &pr.mfa, &pr.att};
    for (GemmParams* g : all) g->epi.oflow = p->oflow;
    for (int b = 0; b < 3; ++b) {
      pr.tdnn1[b].epi.oflow = pr.tdnn2[b].epi.oflow = p->oflow;
      for (int i = 0; i < 7; ++i) pr.res[b][i].epi.oflow = pr.resc[b][i].epi.oflow = p->oflow;
    }
  }
  if (p->use_2sm) {
    pr.block0.idesc = make_idesc_f16(256, 0, 256);
    pr.mfa.idesc = make_idesc_f16(256, 0, 256);
    for (int b = 0; b < 3; ++b) pr.tdnn1[b].idesc = pr.tdnn2[b].idesc = make_idesc_f16(256, 0, 256);
  }
  if (SD_EXPERIMENTS && p->use_chain) {   // device copy of the step table, only for the cooperative chain variant
    std::vector<GemmParams> chain;
    for (int b = 0; b < 3; ++b) {
      chain.push_back(pr.tdnn1[b]);
      for (int i = 0; i < 7; ++i) chain.push_back(pr.res[b][i]);
      chain.push_back(pr.tdnn2[b]);
    }
    void* d = nullptr;
    SD_TRY(dev_alloc(p, &d, chain.size() * sizeof(GemmParams), false));
    SD_CUDA_OK(cudaMemcpy(d, chain.data(), chain.size() * sizeof(GemmParams), cudaMemcpyHostToDevice));
    pr.chain_dev = static_cast<GemmParams*>(d);
  }
  auto ins = p->programs.emplace(key, pr);
  *out = &ins.first->second;
  return SD_OK;
}

// The same GEMM over the row range [row0, row0 + nrows) of its activation tensors (row0 a multiple of 128 and of
// Tp): A is addressed through a_row_base, the outputs through offset pointers, the per-(m block, slot) column
// sums keep their global numbering because the chunk starts on both an m-block and a window boundary.
GemmParams sub_rows(const GemmParams& G, long row0, long nrows) {
  GemmParams S = G;
  S.a_row_base = G.a_row_base + static_cast<int>(row0);
  S.num_m_blocks = static_cast<int>((nrows + BM - 1) / BM);
  EpiParams& E = S.epi;
  E.M_rows = static_cast<int>(nrows);
  E.out = static_cast<__half*>(E.out) + row0 * E.ld_out;
  if (E.out2) E.out2 += row0 * E.ld_out2;
  if (E.colsum) E.colsum += (row0 / BM) * (BM / E.cs_group) * E.N_cols;
  if (E.colsq) E.colsq += (row0 / BM) * (BM / E.cs_group) * E.N_cols;
  E.flags &= ~EF_TMA_OUT;   // the output tensor maps describe the whole tensor
  return S;
}

int build_front(SdEcapaPlan* p, Program& pr) {
  pr.front_built = true;
  constexpr int NF = Program::NFRONT;
  const int nb = pr.B / NF;
  const long nrows = static_cast<long>(nb) * pr.Tp;
  pr.front_ok = p->use_2sm && !p->use_chain && p->use_r2fused && pr.r2_ok && pr.colsum_ok && pr.B % NF == 0 &&
                nb >= 16 && nrows % BM == 0;
  if (!pr.front_ok) return SD_OK;
  for (int c = 0; c < NF; ++c) {
    const long row0 = c * nrows;
    pr.f_block0[c] = sub_rows(pr.block0, row0, nrows);
    pr.f_tdnn1[c] = sub_rows(pr.tdnn1[0], row0, nrows);
    pr.f_tdnn2[c] = sub_rows(pr.tdnn2[0], row0, nrows);
    Res2Params& Q = pr.f_r2[c];
    Q = pr.r2[0];
    Q.B = nb;
    Q.u = pr.r2[0].u + row0 * C1;
    Q.v = pr.r2[0].v + row0 * C1;
    SD_TRY(make_tmap_f16(&Q.tmapU, Q.u, nrows, C1, C1, pr.T + 2 * Q.dil));
    SD_TRY(make_tmap_f16_interior_plain(&Q.tmapV, Q.v, C1, pr.Tp, pr.T, HALO, nb, 32, 16));
  }
  return SD_OK;
}

// Stage boundaries recorded when profiling is on (one event BEFORE each stage + one at the end).
const char* const kStageNames[] = {"fbank", "block0",
                                   "b1.tdnn1", "b1.res2net", "b1.tdnn2", "b1.se",
                                   "b2.tdnn1", "b2.res2net", "b2.tdnn2", "b2.se",
                                   "b3.tdnn1", "b3.res2net", "b3.tdnn2", "b3.se",
                                   "mfa", "asp.context", "asp.attn", "asp.pool", "fc"};
constexpr int kNumStages = sizeof(kStageNames) / sizeof(kStageNames[0]);

void mark(SdEcapaPlan* p, cudaStream_t st) {
  if (!p->profile) return;
  if (p->ev_used == p->ev_pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) { p->profile = false; return; }
    p->ev_pool.push_back(e);
  }
  cudaEventRecord(p->ev_pool[p->ev_used++], st);
}

int launch_res2net_fused(const Res2Params& Q, cudaStream_t st, bool use_pipe) {
  static bool attr_done[64] = {};
  // SD_R2_MODE=0: x_{i+1} loaded behind each chunk's TMEM load (the first version, 0.174 ms per block at B = 512);
  // 2: requested before the accumulator wait / one chunk ahead (0.163 ms); 3 (default): 2 + the frames beyond 127
  // computed transposed, their epilogue spread over all eight warps (0.148 ms); 1: 2 + direct row-per-lane stores of
  // y_i without the staging tile (measured SLOWER: 0.192 ms — the 16-byte pieces of 32 different lines per store)
#if SD_EXPERIMENTS
  static const int mode = [] { const char* e = getenv("SD_R2_MODE"); return e ? atoi(e) : 3; }();
  void (*const kern)(const Res2Params) = mode == 0 ? res2net_fused_kernel<0> : mode == 2 ? res2net_fused_kernel<2> : mode == 3 ? res2net_fused_kernel<3> : res2net_fused_kernel<1>;
  const auto all_modes = {res2net_fused_kernel<0>, res2net_fused_kernel<1>, res2net_fused_kernel<2>, res2net_fused_kernel<3>};
#else
  void (*const kern)(const Res2Params) = res2net_fused_kernel<3>;
  const auto all_modes = {res2net_fused_kernel<3>};
#endif
  if (attr_needed(attr_done)) {
    for (auto k : all_modes)
      if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, R2_SMEM) != cudaSuccess ||
          cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared) != cudaSuccess)
        return fail(SD_ERR_CUDA, "res2net_fused_kernel attributes: %s", cudaGetErrorString(cudaGetLastError()));
  }
  static const char* const r2_trace_path = getenv("SD_R2_TRACE");   // read once: this runs per launch
  // use_pipe (SD_R2_PIPE, read at plan creation): the four-window pipeline; else the first fused kernel (one window
  // per CTA, two CTAs per SM), which also takes T + 2 dil = 161 .. 168
  if (use_pipe && Q.T + 2 * Q.dil <= R2P_RA && Q.B > 0) {
    static bool pipe_attr[64] = {};
    if (attr_needed(pipe_attr) &&
        (cudaFuncSetAttribute(res2net_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, R2P_SMEM) != cudaSuccess ||
         cudaFuncSetAttribute(res2net_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, R2P_SMEM) != cudaSuccess))
      return fail(SD_ERR_CUDA, "res2net_pipe_kernel attributes: %s", cudaGetErrorString(cudaGetLastError()));
    // SD_R2_MC=1: 2-CTA clusters with the weight boxes multicast (each SM fetches half of them).  Correct, measured
    // neutral (0.511 vs 0.513 ms for the three blocks): the launch is bound by the epilogue's own global loads and
    // stores, not by weight delivery, so the default stays the unpaired launch
    static const bool use_mc = [] { const char* e = getenv("SD_R2_MC"); return e && atoi(e) != 0; }();
    int pgrid = Q.B < num_sms() ? Q.B : num_sms();
    if (use_mc) pgrid = (pgrid + 1) & ~1;
    if (use_mc && pgrid > (num_sms() & ~1)) pgrid = num_sms() & ~1;
    if (const char* path = r2_trace_path) {   // debug: CTA 0's per-job clock stamps (no graph capture)
      Res2Params TQ = Q;
      long long* dev = nullptr;
      std::vector<long long> host(32 * 18, 0);
      if (cudaMalloc(&dev, host.size() * 8) != cudaSuccess) return SD_ERR_CUDA;
      cudaMemsetAsync(dev, 0, host.size() * 8, st);
      TQ.trace = dev;
      res2net_pipe_kernel<false><<<Q.B < num_sms() ? Q.B : num_sms(), R2P_THREADS, R2P_SMEM, st>>>(TQ);
      cudaStreamSynchronize(st);
      cudaMemcpy(host.data(), dev, host.size() * 8, cudaMemcpyDeviceToHost);
      cudaFree(dev);
      if (FILE* f = fopen(path, "w")) {
        for (int n = 0; n < 32; ++n) {
          for (int k = 0; k < 18; ++k) fprintf(f, "%lld ", host[n * 18 + k] ? host[n * 18 + k] - host[0] : -1LL);
          fprintf(f, "\n");
        }
        fclose(f);
      }
      count_launch();
      return SD_OK;
    }
    cudaError_t e;
    if (use_mc) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(pgrid);
      cfg.blockDim = dim3(R2P_THREADS);
      cfg.dynamicSmemBytes = R2P_SMEM;
      cfg.stream = st;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at;
      cfg.numAttrs = pdl_flag() ? 2 : 1;
      e = cudaLaunchKernelEx(&cfg, res2net_pipe_kernel<true>, Q);
    } else {
      e = launch_pdl(res2net_pipe_kernel<false>, dim3(pgrid), dim3(R2P_THREADS), R2P_SMEM, st, Q);
    }
    count_launch();
    if (e == cudaSuccess) e = cudaGetLastError();
    static const bool sync_dbg = getenv("SD_SYNC_DEBUG") != nullptr;
    if (e == cudaSuccess && sync_dbg) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess)
      return fail(SD_ERR_CUDA, "res2net_pipe_kernel B=%d T=%d dil=%d: %s", Q.B, Q.T, Q.dil, cudaGetErrorString(e));
    return SD_OK;
  }
  const int grid = Q.B < 2 * num_sms() ? Q.B : 2 * num_sms();
  if (grid <= 0) return SD_OK;
  if (const char* path = r2_trace_path) {   // debug: dump CTA 0's per-conv clock stamps (no graph capture)
    Res2Params T = Q;
    long long* dev = nullptr;
    std::vector<long long> host(32 * 18 + 32 * 4 * 8 + 32 * 12, 0);
    if (cudaMalloc(&dev, host.size() * 8) != cudaSuccess) return SD_ERR_CUDA;
    cudaMemsetAsync(dev, 0, host.size() * 8, st);
    T.trace = dev;
    kern<<<grid, R2_THREADS, R2_SMEM, st>>>(T);
    cudaStreamSynchronize(st);
    cudaMemcpy(host.data(), dev, host.size() * 8, cudaMemcpyDeviceToHost);
    cudaFree(dev);
    if (FILE* f = fopen(path, "w")) {
      for (int n = 0; n < 32; ++n) {
        for (int k = 0; k < 18; ++k) fprintf(f, "%lld ", host[n * 18 + k] ? host[n * 18 + k] - host[0] : -1LL);
        fprintf(f, "\n");
      }
      for (int n = 0; n < 32; ++n) {
        for (int k = 0; k < 12; ++k) fprintf(f, "%lld ", host[1600 + n * 12 + k] ? host[1600 + n * 12 + k] - host[0] : -1LL);
        fprintf(f, "\n");
      }
      for (int n = 0; n < 32 * 4; ++n) {
        for (int k = 0; k < 8; ++k)
          fprintf(f, "%lld ", host[32 * 18 + n * 8 + k] ? host[32 * 18 + n * 8 + k] - host[0] : -1LL);
        fprintf(f, "\n");
      }
      fclose(f);
    }
    count_launch();
    return SD_OK;
  }
  cudaError_t e = launch_pdl(kern, dim3(grid), dim3(R2_THREADS), R2_SMEM, st, Q);
  count_launch();
  if (e == cudaSuccess) e = cudaGetLastError();
  static const bool sync_debug = getenv("SD_SYNC_DEBUG") != nullptr;
  if (e == cudaSuccess && sync_debug) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess)
    return fail(SD_ERR_CUDA, "res2net_fused_kernel B=%d T=%d dil=%d: %s", Q.B, Q.T, Q.dil, cudaGetErrorString(e));
  return SD_OK;
}

// the 256-wide TDNN GEMMs (block0, tdnn1/2, MFA): 2-CTA multicast variant unless disabled
int launch_big(SdEcapaPlan* p, const GemmParams& P, cudaStream_t st) {
  if (p->use_2sm) return launch_gemm_2sm(P, st);
#if SD_EXPERIMENTS
  if (p->use_mc) return launch_gemm_mc_t<EPI_TDNN, 256>(P, st);
#endif
  return launch_gemm<EPI_TDNN>(P, st);
}

// The fixed-pointer part of the forward: block0 ... FC, reading p->feats and writing p->emb_tmp.
struct PdlScope {   // programmatic dependent launch for every kernel launched inside the scope
  explicit PdlScope(bool on) { pdl_flag() = on; }
  ~PdlScope() { pdl_flag() = false; }
};

int front_body(SdEcapaPlan* p, Program& pr, int c, cudaStream_t st) {
  SD_TRY(launch_big(p, pr.f_block0[c], st));
  SD_TRY(launch_big(p, pr.f_tdnn1[c], st));
  SD_TRY(launch_res2net_fused(pr.f_r2[c], st, p->use_r2pipe));
  SD_TRY(launch_big(p, pr.f_tdnn2[c], st));
  return SD_OK;
}

// skip_front: block0 and block 1's tdnn1 / Res2Net / tdnn2 already ran per upload chunk (front_body)
int trunk_body(SdEcapaPlan* p, Program& pr, cudaStream_t st, bool skip_front = false) {
  PdlScope pdl(p->use_pdl && !p->use_chain);
  const int B = pr.B, T = pr.T, Tp = pr.Tp;
  const long R = pr.rows;
  mark(p, st);  // end of fbank / start of block0
  if (!skip_front) {
    SD_CUDA_OK(cudaMemsetAsync(p->oflow, 0, sizeof(int), st));   // per-forward overflow flag (a graph node on replay)
    SD_TRY(launch_big(p, pr.block0, st));
  }
  for (int b = 0; b < 3; ++b) {
    const __half* in = b == 0 ? p->x0 : p->cat + (size_t)(b - 1) * C1;
    const int ld_in = b == 0 ? C1 : C3;
    mark(p, st);
    if (skip_front && b == 0) {
      mark(p, st);
      mark(p, st);
#if SD_EXPERIMENTS
    } else if (p->use_chain) {
      // tdnn1 -> 7 dependent Res2Net convs -> tdnn2 in ONE cooperative launch (grid barrier between steps)
      SD_TRY((launch_gemm_chain<EPI_TDNN, 256>(pr.chain_dev + 9 * b, 9, st)));
      mark(p, st);
      mark(p, st);
#endif
    } else {
      SD_TRY(launch_big(p, pr.tdnn1[b], st));
      mark(p, st);
      if (p->use_r2fused && pr.r2_ok) {
        SD_TRY(launch_res2net_fused(pr.r2[b], st, p->use_r2pipe));
      } else {
        for (int i = 0; i < 7; ++i) {
          if (p->use_conv3) SD_TRY(launch_gemm<EPI_CONV3>(pr.resc[b][i], st));
          else SD_TRY(launch_gemm<EPI_TDNN>(pr.res[b][i], st));
        }
      }
      mark(p, st);
      SD_TRY(launch_big(p, pr.tdnn2[b], st));
    }
    mark(p, st);
    // squeeze-excitation gate in one launch: finish the per-window column sums tdnn2's write-out left (or take
    // the mean of the separate pass), 1024 -> 128 -> 1024 MLP, sigmoid
    if (!pr.colsum_ok)
      SD_CUDA_OK(launch_pdl(time_mean_kernel, dim3(C1 / 256, B), dim3(128), 0, st, p->w, C1, Tp, T, HALO, C1, p->se_mean));
    SD_CUDA_OK(launch_pdl(se_gate_kernel, dim3((B + SEG - 1) / SEG), dim3(SE_THREADS), 0, st,
                          pr.colsum_ok ? p->cs_se : static_cast<const float*>(nullptr), p->blk[b].tdnn2.shift, Tp, T,
                          pr.cs_group, pr.colsum_ok ? static_cast<const float*>(nullptr) : p->se_mean, p->blk[b].se_w1h,
                          p->blk[b].se_b1, p->blk[b].se_w2th, p->blk[b].se_b2, B,
                          pr.colsum_ok ? p->se_mean : static_cast<float*>(nullptr), p->se_scale));
    const long vecs = R * (C1 / 8);
    const int grid = (int)((vecs + 255) / 256 < 148L * 16 ? (vecs + 255) / 256 : 148L * 16);
    SD_CUDA_OK(launch_pdl(se_apply_kernel, dim3(grid), dim3(256), 0, st, p->w, C1, p->se_scale, in, ld_in,
                          p->cat + (size_t)b * C1, C3, R, Tp, C1, p->use_l2_order ? 1 : 0, p->oflow));
    SD_CUDA_OK(cudaGetLastError());
    count_launch(pr.colsum_ok ? 2 : 3);
  }
  mark(p, st);
  SD_TRY(launch_big(p, pr.mfa, st));
  mark(p, st);
  if (pr.colsum_ok)
    SD_CUDA_OK(launch_pdl(colstats_finish_kernel, dim3(B), dim3(256), 0, st, p->cs_mfa, p->cq_mfa, p->wmfa.shift,
                          C3, Tp, T, pr.cs_group, p->stats, 2 * C3, p->stats + C3, p->stats_h));
  else
    SD_CUDA_OK(launch_pdl(time_mean_std_kernel, dim3(C3 / 256, B), dim3(128), 0, st, p->h, C3, Tp, T, HALO, C3, p->stats,
                          p->stats_h));
  SD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  // context bias: W_mean . mean + W_std . std  (the 2/3 of asp.tdnn that is constant over time)
  SD_TRY(launch_gemm<EPI_F32>(pr.ctx, st));
  SD_CUDA_OK(launch_pdl(sum_splits_kernel, dim3((B * ATT + 255) / 256), dim3(256), 0, st, p->ctx_part, p->ksplit, (long)B * ATT,
                        (long)B * ATT, p->uttbias));
  SD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  mark(p, st);
  SD_TRY(launch_gemm<EPI_ATT>(pr.att, st));
  mark(p, st);
  SD_TRY(launch_gemm<EPI_POOL>(pr.pool, st));
  mark(p, st);
  SD_TRY(launch_gemm<EPI_F32>(pr.fc, st));
  return SD_OK;
}

// Runs the trunk and delivers the embeddings.  After a shape has run once eagerly, its ~45 launches
// are captured into a CUDA graph and replayed (every pointer inside is a plan-owned buffer, so the
// graph is static); the caller-dependent ends — fbank reading the caller's audio before it, the
// L2-norm / copy into the caller's buffer after it — stay ordinary launches.  Replay removes
// ~0.2 ms of launch gaps per 512-window batch (4.28 -> 4.09 ms measured).
int run_trunk(SdEcapaPlan* p, Program& pr, int l2_normalize, float* emb, cudaStream_t st, bool skip_front = false) {
  const int B = pr.B;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  SD_CUDA_OK(cudaStreamIsCapturing(st, &cap));
  const bool can_graph = p->use_graph && !p->profile && cap == cudaStreamCaptureStatusNone;
  cudaGraphExec_t& graph = skip_front ? pr.graph_tail : pr.graph;
  int& graph_launches = skip_front ? pr.graph_tail_launches : pr.graph_launches;
  int& runs = skip_front ? pr.runs_tail : pr.runs;
  if (can_graph && graph) {
    SD_CUDA_OK(cudaGraphLaunch(graph, st));
    count_launch(graph_launches);
  } else if (can_graph && runs >= 2) {   // capture on the third use: one-off shapes are not worth ~1 ms
    const long before = launch_counter().load();
    cudaGraph_t g = nullptr;
    // capture on a plan-owned stream (the caller's may be the legacy default stream, which cannot
    // be captured); the instantiated graph is then launched into the caller's stream
    if (!p->cap_stream) SD_CUDA_OK(cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking));
    SD_CUDA_OK(cudaStreamBeginCapture(p->cap_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = trunk_body(p, pr, p->cap_stream, skip_front);
    const cudaError_t ec = cudaStreamEndCapture(p->cap_stream, &g);
    if (rc != SD_OK) { if (g) cudaGraphDestroy(g); return rc; }
    if (ec != cudaSuccess) return fail(SD_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ec));
    const cudaError_t ei = cudaGraphInstantiate(&graph, g, 0);
    cudaGraphDestroy(g);
    if (ei != cudaSuccess) { graph = nullptr; return fail(SD_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ei)); }
    graph_launches = (int)(launch_counter().load() - before);
    SD_CUDA_OK(cudaGraphLaunch(graph, st));
  } else {
    SD_TRY(trunk_body(p, pr, st, skip_front));
  }
  ++runs;
  fc_finish_kernel<<<(B + 7) / 8, 256, 0, st>>>(p->emb_tmp, p->ksplit, (long)B * EMB, B, EMB, l2_normalize, 1e-8f, emb,
                                                p->oflow, p->oflow + 1);
  SD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  mark(p, st);  // end of fc
  if (p->profile) ++p->forwards_profiled;
  p->last = &pr;
  return SD_OK;
}

int check_shape(SdEcapaPlan* p, int B, int T) {
  if (B < 1 || T < 2 * HALO + 2)
    return fail(SD_ERR_ARG, "ecapa: need B >= 1 and T >= %d frames (B=%d T=%d)", 2 * HALO + 2, B, T);
  if ((long)B * tp_of(T) > p->max_rows)
    return fail(SD_ERR_ARG, "ecapa: B=%d x T=%d exceeds the plan's workspace (%ld rows)", B, T, p->max_rows);
  return SD_OK;
}

}  // namespace

extern "C" int sd_ecapa_plan_create(const char* const* names, const float* const* tensors,
                                    const int64_t* numels, int n_tensors, int max_batch,
                                    int max_samples, SdEcapaPlan** plan_out) {
  if (!names || !tensors || !numels || !plan_out || n_tensors < 1 || max_batch < 1 || max_samples < 400)
    return fail(SD_ERR_ARG, "sd_ecapa_plan_create: bad arguments");
  StateDict sd;
  for (int i = 0; i < n_tensors; ++i) sd.m[names[i]] = {tensors[i], numels[i]};
  SdEcapaPlan* p = new SdEcapaPlan;
  p->max_batch = max_batch;
  if (const char* e = getenv("SD_ECAPA_CHAIN")) p->use_chain = SD_EXPERIMENTS && atoi(e) != 0;
  if (const char* e = getenv("SD_ECAPA_GRAPH")) p->use_graph = atoi(e) != 0;
  if (const char* e = getenv("SD_ECAPA_CONV3")) p->use_conv3 = atoi(e) != 0;
  if (const char* e = getenv("SD_ECAPA_R2FUSED")) p->use_r2fused = atoi(e) != 0;
  if (const char* e = getenv("SD_R2_PIPE")) p->use_r2pipe = atoi(e) != 0;
  if (const char* e = getenv("SD_ECAPA_COLSUM")) p->use_colsum = atoi(e) != 0;
  if (const char* e = getenv("SD_ECAPA_TMAOUT")) p->use_tma_out = atoi(e) != 0;
  if (const char* e = getenv("SD_ECAPA_TMAOUT0")) p->use_tma_out_reflect = atoi(e) != 0;
  if (const char* e = getenv("SD_ECAPA_KSPLIT")) p->ksplit = atoi(e) < 1 ? 1 : atoi(e) > KSPLIT_MAX ? KSPLIT_MAX : atoi(e);
  if (const char* e = getenv("SD_ECAPA_L2ORDER")) p->use_l2_order = atoi(e) != 0;
  if (const char* e = getenv("SD_ECAPA_MC")) p->use_mc = SD_EXPERIMENTS && atoi(e) != 0;
  if (const char* e = getenv("SD_ECAPA_2SM")) p->use_2sm = atoi(e) != 0;
  if (const char* e = getenv("SD_ECAPA_PDL")) p->use_pdl = atoi(e) != 0;
  if (p->use_chain) p->use_mc = p->use_2sm = false;  // the cooperative chain uses the plain kernels
  if (p->use_2sm) p->use_mc = false;
  p->max_samples = max_samples;
  const int maxT = 1 + max_samples / 160;
  p->max_rows = (long)max_batch * tp_of(maxT);
  int st = SD_OK;
  auto body = [&]() -> int {
    SD_TRY(load_tdnn(p, sd, "blocks.0", C1, 80, 5, FEAT_P, 80, &p->w0));
    for (int b = 0; b < 3; ++b) {
      const std::string pre = "blocks." + std::to_string(b + 1);
      BlockW& bw = p->blk[b];
      bw.dil = b + 2;
      SD_TRY(load_tdnn(p, sd, pre + ".tdnn1", C1, C1, 1, C1, C1, &bw.tdnn1));
      for (int i = 0; i < 7; ++i)
        SD_TRY(load_tdnn(p, sd, pre + ".res2net_block.blocks." + std::to_string(i), SUB, SUB, 3, SUB, SUB,
                         &bw.res[i]));
      SD_TRY(load_tdnn(p, sd, pre + ".tdnn2", C1, C1, 1, C1, C1, &bw.tdnn2));
      const float *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr;
      SD_TRY(sd.get(pre + ".se_block.conv1.conv.weight", (int64_t)SE * C1, &w1));
      SD_TRY(sd.get(pre + ".se_block.conv1.conv.bias", SE, &b1));
      SD_TRY(sd.get(pre + ".se_block.conv2.conv.weight", (int64_t)C1 * SE, &w2));
      SD_TRY(sd.get(pre + ".se_block.conv2.conv.bias", C1, &b2));
      std::vector<__half> w1h((size_t)SE * C1), w2th((size_t)SE * C1);
      for (size_t i = 0; i < w1h.size(); ++i) w1h[i] = __float2half_rn(w1[i]);
      for (int c = 0; c < C1; ++c)
        for (int j = 0; j < SE; ++j) w2th[(size_t)j * C1 + c] = __float2half_rn(w2[(size_t)c * SE + j]);
      SD_TRY(upload(p, &bw.se_w1h, w1h));
      SD_TRY(upload(p, &bw.se_b1, std::vector<float>(b1, b1 + SE)));
      SD_TRY(upload(p, &bw.se_w2th, w2th));
      SD_TRY(upload(p, &bw.se_b2, std::vector<float>(b2, b2 + C1)));
    }
    SD_TRY(load_tdnn(p, sd, "mfa", C3, C3, 1, C3, C3, &p->wmfa));
    // asp.tdnn: [ATT, 3*C3, 1]; columns [0,C3) act on x (tensor cores), [C3, 3*C3) on mean|std
    SD_TRY(load_tdnn(p, sd, "asp.tdnn", ATT, 3 * C3, 1, C3, C3, &p->watt));
    {
      const float* W;
      SD_TRY(sd.get("asp.tdnn.conv.conv.weight", (int64_t)ATT * 3 * C3, &W));
      std::vector<__half> ms((size_t)ATT * 2 * C3);
      for (int o = 0; o < ATT; ++o)
        for (int c = 0; c < 2 * C3; ++c)
          ms[(size_t)o * 2 * C3 + c] = __float2half_rn(W[(size_t)o * 3 * C3 + C3 + c]);
      SD_TRY(upload(p, &p->Wams, ms));
      const float* W2;
      SD_TRY(sd.get("asp.conv.conv.weight", (int64_t)C3 * ATT, &W2));
      const float* b2;  // constant over time -> cancels in the softmax; only validated
      SD_TRY(sd.get("asp.conv.conv.bias", C3, &b2));
      std::vector<__half> w2h((size_t)C3 * ATT);
      for (size_t i = 0; i < w2h.size(); ++i) w2h[i] = __float2half_rn(W2[i]);
      SD_TRY(upload(p, &p->Wa2, w2h));
    }
    {
      // fold asp_bn (eval) into fc: W' = W diag(sc), b' = b + W sh
      const float *g, *be, *rm, *rv, *W, *b;
      SD_TRY(sd.get("asp_bn.norm.weight", 2 * C3, &g));
      SD_TRY(sd.get("asp_bn.norm.bias", 2 * C3, &be));
      SD_TRY(sd.get("asp_bn.norm.running_mean", 2 * C3, &rm));
      SD_TRY(sd.get("asp_bn.norm.running_var", 2 * C3, &rv));
      SD_TRY(sd.get("fc.conv.weight", (int64_t)EMB * 2 * C3, &W));
      SD_TRY(sd.get("fc.conv.bias", EMB, &b));
      std::vector<__half> wf((size_t)EMB * 2 * C3);
      std::vector<float> bf(EMB);
      for (int o = 0; o < EMB; ++o) {
        double acc = b[o];
        for (int k = 0; k < 2 * C3; ++k) {
          const double sc = (double)g[k] / sqrt((double)rv[k] + 1e-5);
          const double sh = (double)be[k] - (double)rm[k] * sc;
          wf[(size_t)o * 2 * C3 + k] = __float2half_rn((float)((double)W[(size_t)o * 2 * C3 + k] * sc));
          acc += (double)W[(size_t)o * 2 * C3 + k] * sh;
        }
        bf[o] = (float)acc;
      }
      SD_TRY(upload(p, &p->Wfc, wf));
      SD_TRY(upload(p, &p->bfc, bf));
    }
    // activation workspace (zero-initialised once: padding rows are never written afterwards)
    const size_t R = (size_t)p->max_rows;
    SD_TRY(dev_alloc(p, (void**)&p->feats, R * FEAT_P * 2, true));
    SD_TRY(dev_alloc(p, (void**)&p->x0, R * C1 * 2, true));
    SD_TRY(dev_alloc(p, (void**)&p->cat, R * C3 * 2, true));
    SD_TRY(dev_alloc(p, (void**)&p->u, R * C1 * 2, true));
    SD_TRY(dev_alloc(p, (void**)&p->v, R * C1 * 2, true));
    SD_TRY(dev_alloc(p, (void**)&p->w, R * C1 * 2, true));
    SD_TRY(dev_alloc(p, (void**)&p->s[0], R * SUB * 2, true));
    SD_TRY(dev_alloc(p, (void**)&p->s[1], R * SUB * 2, true));
    SD_TRY(dev_alloc(p, (void**)&p->h, R * C3 * 2, true));
    SD_TRY(dev_alloc(p, (void**)&p->attn, R * ATT * 2, true));
    SD_TRY(dev_alloc(p, (void**)&p->raw, R * 80 * 4, true));
    SD_TRY(dev_alloc(p, (void**)&p->raw2, R * 80 * 4, true));
    const size_t MB = (size_t)p->max_rows / tp_of(2 * HALO + 2) + 1;  // most utterances any shape can have
    SD_TRY(dev_alloc(p, (void**)&p->se_mean, MB * C1 * 4, true));
    SD_TRY(dev_alloc(p, (void**)&p->se_scale, MB * C1 * 4, true));
    SD_TRY(dev_alloc(p, (void**)&p->se_hid, MB * SE * 4, true));
    SD_TRY(dev_alloc(p, (void**)&p->oflow, 2 * sizeof(int), true));
    {
      const size_t mblocks = (size_t)(p->max_rows + BM - 1) / BM + 1;
      // one partial sum per group of gcd(128, Tp) >= 16 rows: at most 8 groups per 128-row block
      SD_TRY(dev_alloc(p, (void**)&p->cs_se, mblocks * 8 * C1 * 4, true));
      SD_TRY(dev_alloc(p, (void**)&p->cs_mfa, mblocks * 8 * C3 * 4, true));
      SD_TRY(dev_alloc(p, (void**)&p->cq_mfa, mblocks * 8 * C3 * 4, true));
    }
    SD_TRY(dev_alloc(p, (void**)&p->stats, MB * 2 * C3 * 4, true));
    SD_TRY(dev_alloc(p, (void**)&p->uttbias, MB * ATT * 4, true));
    SD_TRY(dev_alloc(p, (void**)&p->pooled, MB * 2 * C3 * 4, true));
    SD_TRY(dev_alloc(p, (void**)&p->emb_tmp, MB * EMB * 4 * KSPLIT_MAX, true));
    SD_TRY(dev_alloc(p, (void**)&p->ctx_part, MB * ATT * 4 * KSPLIT_MAX, true));
    SD_TRY(dev_alloc(p, (void**)&p->stats_h, MB * 2 * C3 * 2, true));
    SD_TRY(dev_alloc(p, (void**)&p->pooled_h, MB * 2 * C3 * 2, true));
    SD_CUDA_OK(cudaDeviceSynchronize());
    return SD_OK;
  };
  st = body();
  if (st != SD_OK) {
    for (void* d : p->allocs) cudaFree(d);
    delete p;
    return st;
  }
  *plan_out = p;
  return SD_OK;
}

extern "C" int sd_ecapa_plan_destroy(SdEcapaPlan* p) {
  if (!p) return SD_OK;
  cudaDeviceSynchronize();
  for (cudaEvent_t e : p->ev_pool) cudaEventDestroy(e);
  if (p->cap_stream) cudaStreamDestroy(p->cap_stream);
  if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
  for (cudaEvent_t e : p->copy_ev)
    if (e) cudaEventDestroy(e);
  if (p->start_ev) cudaEventDestroy(p->start_ev);
  if (p->h2d_buf) cudaFree(p->h2d_buf);
  if (p->emb_stage) cudaFree(p->emb_stage);
  if (p->emb_pinned) cudaFreeHost(p->emb_pinned);
  if (p->oflow_host) cudaFreeHost(p->oflow_host);
  delete p->ring;
  for (auto& kv : p->programs) {
    if (kv.second.graph) cudaGraphExecDestroy(kv.second.graph);
    if (kv.second.graph_tail) cudaGraphExecDestroy(kv.second.graph_tail);
  }
  for (void* d : p->allocs) cudaFree(d);
  delete p;
  return SD_OK;
}

extern "C" int sd_ecapa_embed(SdEcapaPlan* p, const float* wav_dev, long wav_stride, int B,
                              int n_samples, int l2_normalize, float* emb_dev, void* stream) {
  if (!p || !wav_dev || !emb_dev) return fail(SD_ERR_ARG, "sd_ecapa_embed: NULL argument");
  if (n_samples < 400) return fail(SD_ERR_ARG, "sd_ecapa_embed: n_samples=%d < 400", n_samples);
  const int T = 1 + n_samples / 160;
  SD_TRY(check_shape(p, B, T));
  Program* pr = nullptr;
  SD_TRY(build_program(p, B, T, &pr));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mark(p, st);  // start of fbank
  SD_TRY(fbank_launch(wav_dev, wav_stride, B, n_samples, SD_FBANK_SPEECHBRAIN, 1, p->raw, nullptr,
                      p->feats, pr->Tp, HALO, st, nullptr, p->raw2));
  return run_trunk(p, *pr, l2_normalize, emb_dev, st);
}

extern "C" int sd_ecapa_embed_offsets(SdEcapaPlan* p, const float* wav_dev, const long* offsets_dev, int B,
                                      int n_samples, int l2_normalize, float* emb_dev, void* stream) {
  if (!p || !wav_dev || !offsets_dev || !emb_dev) return fail(SD_ERR_ARG, "sd_ecapa_embed_offsets: NULL argument");
  if (n_samples < 400) return fail(SD_ERR_ARG, "sd_ecapa_embed_offsets: n_samples=%d < 400", n_samples);
  const int T = 1 + n_samples / 160;
  SD_TRY(check_shape(p, B, T));
  Program* pr = nullptr;
  SD_TRY(build_program(p, B, T, &pr));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mark(p, st);  // start of fbank
  SD_TRY(fbank_launch(wav_dev, 0, B, n_samples, SD_FBANK_SPEECHBRAIN, 1, p->raw, nullptr, p->feats, pr->Tp, HALO, st,
                      offsets_dev, p->raw2));
  return run_trunk(p, *pr, l2_normalize, emb_dev, st);
}

extern "C" int sd_ecapa_embed_host(SdEcapaPlan* p, const float* wav_host, long wav_stride, int B, int n_samples,
                                   int l2_normalize, float* emb_host, void* stream) {
  if (!p || !wav_host || !emb_host) return fail(SD_ERR_ARG, "sd_ecapa_embed_host: NULL argument");
  if (n_samples < 400 || wav_stride < 1)
    return fail(SD_ERR_ARG, "sd_ecapa_embed_host: n_samples=%d stride=%ld", n_samples, wav_stride);
  const int T = 1 + n_samples / 160;
  // SD_HOST_TRACE=1: host-side timeline of each call on stderr (us since entry): queued-uploads, queued-trunk, synced, exit
  static const bool host_trace = getenv("SD_HOST_TRACE") != nullptr;
  const auto ht0 = std::chrono::steady_clock::now();
  auto ht_us = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - ht0).count(); };
  double ht[4] = {0, 0, 0, 0};
  SD_TRY(check_shape(p, B, T));
  Program* pr = nullptr;
  SD_TRY(build_program(p, B, T, &pr));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  constexpr int NCH_MAX = 8;   // = sizeof(copy_ev) / sizeof(copy_ev[0])
  // upload chunks (SD_ECAPA_UPCHUNKS = 1..8): more chunks expose less of the first one, each costs an event round trip
  static const int NCH = [] { const char* e = getenv("SD_ECAPA_UPCHUNKS"); const int v = e ? atoi(e) : 4; return v < 1 ? 1 : v > NCH_MAX ? NCH_MAX : v; }();
  if (!p->copy_stream) {
    SD_CUDA_OK(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    for (int c = 0; c < NCH_MAX; ++c) SD_CUDA_OK(cudaEventCreateWithFlags(&p->copy_ev[c], cudaEventDisableTiming));
    SD_CUDA_OK(cudaEventCreateWithFlags(&p->start_ev, cudaEventDisableTiming));
  }
  if (B > p->emb_cap) {
    SD_CUDA_OK(cudaStreamSynchronize(st));
    if (p->emb_stage) cudaFree(p->emb_stage);
    if (p->emb_pinned) cudaFreeHost(p->emb_pinned);
    p->emb_stage = p->emb_pinned = nullptr;
    p->emb_cap = 0;
    SD_CUDA_OK(cudaMalloc(&p->emb_stage, static_cast<size_t>(B) * EMB * sizeof(float)));
    if (cudaHostAlloc(&p->emb_pinned, static_cast<size_t>(B) * EMB * sizeof(float), cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      p->emb_pinned = nullptr;   // fall back to copying straight into the caller's buffer
    }
    p->emb_cap = B;
  }
  const size_t span = static_cast<size_t>(B - 1) * wav_stride + n_samples;
  if (span > p->h2d_cap) {
    SD_CUDA_OK(cudaStreamSynchronize(st));
    if (p->h2d_buf) cudaFree(p->h2d_buf);
    p->h2d_buf = nullptr;
    p->h2d_cap = 0;
    SD_CUDA_OK(cudaMalloc(&p->h2d_buf, span * sizeof(float)));
    p->h2d_cap = span;
  }
  // the staging buffer may still be read by work queued earlier on `st`
  SD_CUDA_OK(cudaEventRecord(p->start_ev, st));
  SD_CUDA_OK(cudaStreamWaitEvent(p->copy_stream, p->start_ev, 0));
  mark(p, st);  // start of fbank
  const int chunks = B >= 64 ? NCH : 1;
  // front of the trunk per upload chunk (Program::NFRONT): needs chunk boundaries on m-block boundaries
  // Measured and rejected (off unless SD_ECAPA_PIPE=1): per-call wall clock at B = 512 went 4.01 -> 4.13 ms.  The
  // upload costs only 0.18 ms of a call (3.83 ms with no copy at all), while a quarter-batch front is 0.16 ms
  // slower in total than one launch per layer (128 windows leave the Res2Net kernel one CTA per SM, the GEMMs 4.3
  // waves of tiles).
  const char* const pipe_str = getenv("SD_ECAPA_PIPE");
  const bool pipe_env = pipe_str && atoi(pipe_str) != 0;
  if (!pr->front_built) SD_TRY(build_front(p, *pr));
  const bool piped = pipe_env && chunks == Program::NFRONT && pr->front_ok && !p->profile;
  if (piped) SD_CUDA_OK(cudaMemsetAsync(p->oflow, 0, sizeof(int), st));   // the front runs before trunk_body's own reset
  static const bool nocopy = getenv("SD_DEBUG_NOCOPY") != nullptr;  // TIMING PROBE ONLY: leaves the staging as it is
  // chunk c = windows [B*c/chunks, B*(c+1)/chunks): once its samples are on the device, its fbank kernels (and, when
  // piped, the front of the trunk) are queued behind the copy event
  auto chunk_range = [&](int c, int* b0, int* b1) {
    *b0 = static_cast<int>(static_cast<long>(B) * c / chunks);
    *b1 = static_cast<int>(static_cast<long>(B) * (c + 1) / chunks);
  };
  auto launch_chunk = [&](int c) -> int {
    int b0, b1;
    chunk_range(c, &b0, &b1);
    if (b1 <= b0) return SD_OK;
    SD_CUDA_OK(cudaEventRecord(p->copy_ev[c], p->copy_stream));
    SD_CUDA_OK(cudaStreamWaitEvent(st, p->copy_ev[c], 0));
    SD_TRY(fbank_launch(p->h2d_buf + static_cast<size_t>(b0) * wav_stride, wav_stride, b1 - b0, n_samples,
                        SD_FBANK_SPEECHBRAIN, 1, p->raw + static_cast<size_t>(b0) * T * 80, nullptr,
                        p->feats + static_cast<size_t>(b0) * pr->Tp * FEAT_P, pr->Tp, HALO, st, nullptr,
                        p->raw2 + static_cast<size_t>(b0) * T * 80));
    if (piped) SD_TRY(front_body(p, *pr, c, st));
    return SD_OK;
  };
  // Page-locked caller memory is read by the copy engine directly.  Pageable memory (what the reference's callers
  // pass) goes through the pinned staging ring filled by copy threads, so the transfer is asynchronous and
  // pipelined instead of a driver-staged blocking cudaMemcpy.  SD_ECAPA_STAGING=0 restores the plain copy.
  bool staged = false;
  if (!nocopy && wav_stride <= n_samples) {
    static const bool staging_on = [] { const char* e = getenv("SD_ECAPA_STAGING"); return !e || atoi(e) != 0; }();
    cudaPointerAttributes pa;
    const bool pinned = cudaPointerGetAttributes(&pa, wav_host) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (!pinned && staging_on && !p->ring_failed && span * sizeof(float) >= (size_t(1) << 20)) {
      if (!p->ring) {
        static const int n_thr = [] { const char* e = getenv("SD_ECAPA_HOST_THREADS"); return e ? atoi(e) : 6; }();
        p->ring = new StagingRing;
        if (!p->ring->init(n_thr)) {
          delete p->ring;
          p->ring = nullptr;
          p->ring_failed = true;
        }
      }
      staged = p->ring != nullptr;
    }
  }
  if (staged) {
    int next = 0, rc = SD_OK;
    const cudaError_t e = p->ring->upload(
        reinterpret_cast<const char*>(wav_host), reinterpret_cast<char*>(p->h2d_buf), span * sizeof(float), p->copy_stream,
        [&](size_t done_bytes) {
          while (rc == SD_OK && next < chunks) {
            int b0, b1;
            chunk_range(next, &b0, &b1);
            const size_t end = b1 > b0 ? (static_cast<size_t>(b1 - 1) * wav_stride + n_samples) * sizeof(float) : 0;
            if (end > done_bytes) break;
            rc = launch_chunk(next++);
          }
        });
    if (e != cudaSuccess) return fail(SD_ERR_CUDA, "staged upload failed: %s", cudaGetErrorString(e));
    SD_TRY(rc);
  } else {
    size_t copied = 0;  // samples of the span already queued (windows overlap when stride < n)
    for (int c = 0; c < chunks; ++c) {
      int b0, b1;
      chunk_range(c, &b0, &b1);
      if (b1 <= b0) continue;
      if (nocopy) {
      } else if (wav_stride <= n_samples) {
        const size_t end = static_cast<size_t>(b1 - 1) * wav_stride + n_samples;
        SD_CUDA_OK(cudaMemcpyAsync(p->h2d_buf + copied, wav_host + copied, (end - copied) * sizeof(float),
                                   cudaMemcpyHostToDevice, p->copy_stream));
        copied = end;
      } else {   // gaps between windows belong to the caller: copy the windows only
        SD_CUDA_OK(cudaMemcpy2DAsync(p->h2d_buf + static_cast<size_t>(b0) * wav_stride, wav_stride * sizeof(float),
                                     wav_host + static_cast<size_t>(b0) * wav_stride, wav_stride * sizeof(float),
                                     static_cast<size_t>(n_samples) * sizeof(float), b1 - b0, cudaMemcpyHostToDevice,
                                     p->copy_stream));
      }
      SD_TRY(launch_chunk(c));
    }
  }
  ht[0] = ht_us();
  static cudaEvent_t ht_ev[2] = {nullptr, nullptr};
  if (host_trace) {
    if (!ht_ev[0]) { cudaEventCreate(&ht_ev[0]); cudaEventCreate(&ht_ev[1]); }
    cudaEventRecord(ht_ev[0], st);
  }
  SD_TRY(run_trunk(p, *pr, l2_normalize, p->emb_stage, st, piped));
  if (host_trace) cudaEventRecord(ht_ev[1], st);
  ht[1] = ht_us();
  const size_t emb_bytes = static_cast<size_t>(B) * EMB * sizeof(float);
  SD_CUDA_OK(cudaMemcpyAsync(p->emb_pinned ? p->emb_pinned : emb_host, p->emb_stage, emb_bytes, cudaMemcpyDeviceToHost, st));
  if (!p->oflow_host && cudaHostAlloc(reinterpret_cast<void**>(&p->oflow_host), sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    p->oflow_host = nullptr;
  }
  if (p->oflow_host) SD_CUDA_OK(cudaMemcpyAsync(p->oflow_host, p->oflow + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  SD_CUDA_OK(cudaStreamSynchronize(st));
  ht[2] = ht_us();
  if (p->emb_pinned) memcpy(emb_host, p->emb_pinned, emb_bytes);
  ht[3] = ht_us();
  if (host_trace) {
    float trunk_ms = 0.f;
    cudaEventElapsedTime(&trunk_ms, ht_ev[0], ht_ev[1]);
    fprintf(stderr, "[sd host trace] B=%d uploads+fbank queued %.0f us, trunk queued %.0f, synced %.0f, copied out %.0f; "
            "trunk on the device %.0f us (from the last fbank to the embeddings)\n", B, ht[0], ht[1], ht[2], ht[3], 1e3f * trunk_ms);
  }
  if (p->oflow_host && *p->oflow_host != 0) {
    cudaMemsetAsync(p->oflow + 1, 0, sizeof(int), st);
    return fail(SD_ERR_RANGE, "sd_ecapa_embed_host: an activation left the f16 range (|x| > 65504) and was saturated; "
                              "the embeddings of this call are NaN");
  }
  return SD_OK;
}

extern "C" int sd_ecapa_overflow(SdEcapaPlan* p, int reset, int* flag_out, void* stream) {
  if (!p || !flag_out) return fail(SD_ERR_ARG, "sd_ecapa_overflow: NULL argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int v = 0;
  SD_CUDA_OK(cudaMemcpyAsync(&v, p->oflow + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  SD_CUDA_OK(cudaStreamSynchronize(st));
  if (reset && v) SD_CUDA_OK(cudaMemsetAsync(p->oflow + 1, 0, sizeof(int), st));
  *flag_out = v;
  return SD_OK;
}

extern "C" int sd_ecapa_forward_feats(SdEcapaPlan* p, const float* feats_dev, int B, int T,
                                      int l2_normalize, float* emb_dev, void* stream) {
  if (!p || !feats_dev || !emb_dev) return fail(SD_ERR_ARG, "sd_ecapa_forward_feats: NULL argument");
  SD_TRY(check_shape(p, B, T));
  Program* pr = nullptr;
  SD_TRY(build_program(p, B, T, &pr));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mark(p, st);  // start of the feature repack (reported under "fbank")
  SD_TRY(feats_to_padded_f16(feats_dev, B, T, p->feats, pr->Tp, HALO, st));
  return run_trunk(p, *pr, l2_normalize, emb_dev, st);
}

extern "C" int sd_ecapa_debug_fetch(SdEcapaPlan* p, const char* name, float* out_dev, int* C_out,
                                    void* stream) {
  if (!p || !name || !out_dev || !p->last) return fail(SD_ERR_ARG, "sd_ecapa_debug_fetch: no forward yet");
  const Program& pr = *p->last;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const std::string n(name);
  const __half* src = nullptr;
  int ld = 0, off = 0, C = 0;
  if (n == "feats") { src = p->feats; ld = FEAT_P; C = 80; }
  else if (n == "block0") { src = p->x0; ld = C1; C = C1; }
  else if (n == "b1.out") { src = p->cat; ld = C3; C = C1; }
  else if (n == "b2.out") { src = p->cat; ld = C3; off = C1; C = C1; }
  else if (n == "b3.out") { src = p->cat; ld = C3; off = 2 * C1; C = C1; }
  else if (n == "b3.tdnn1") { src = p->u; ld = C1; C = C1; }
  else if (n == "b3.res2net") { src = p->v; ld = C1; C = C1; }
  else if (n == "b3.tdnn2") { src = p->w; ld = C1; C = C1; }
  else if (n == "mfa") { src = p->h; ld = C3; C = C3; }
  else if (n == "asp.attn") { src = p->attn; ld = ATT; C = ATT; }
  if (src) {
    const long total = (long)pr.B * pr.T * C;
    fetch_interior_kernel<<<(int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096), 256, 0, st>>>(
        src, ld, off, pr.Tp, pr.T, HALO, C, total, out_dev);
    SD_CUDA_OK(cudaGetLastError());
    count_launch();
    if (C_out) *C_out = C;
    return SD_OK;
  }
  const float* fsrc = nullptr;
  if (n == "b3.se") { fsrc = p->se_scale; C = C1; }
  else if (n == "asp.stats") { fsrc = p->stats; C = 2 * C3; }
  else if (n == "asp.uttbias") { fsrc = p->uttbias; C = ATT; }
  else if (n == "pooled") { fsrc = p->pooled; C = 2 * C3; }
  if (!fsrc) return fail(SD_ERR_ARG, "sd_ecapa_debug_fetch: unknown tensor '%s'", name);
  SD_CUDA_OK(cudaMemcpyAsync(out_dev, fsrc, (size_t)pr.B * C * 4, cudaMemcpyDeviceToDevice, st));
  if (C_out) *C_out = C;
  return SD_OK;
}

extern "C" double sd_ecapa_flops_per_window(int T) {
  // MACs per frame / per utterance of the contractions issued (ASP decomposed)
  const double per_frame = 80.0 * 5 * C1 + 3.0 * (2.0 * C1 * C1 + 7.0 * 3 * SUB * SUB) +
                           (double)C3 * C3 + (double)C3 * ATT + (double)ATT * C3;
  const double per_utt = 3.0 * 2 * C1 * SE + 2.0 * C3 * ATT + 2.0 * C3 * EMB;
  return 2.0 * (per_frame * T + per_utt);
}

extern "C" int sd_l2norm_f32(const float* x_dev, int N, int D, float eps, float* out_dev, void* stream) {
  if (!x_dev || !out_dev || N < 0 || D < 1) return fail(SD_ERR_ARG, "sd_l2norm_f32: bad arguments");
  if (N == 0) return SD_OK;
  l2norm_rows_kernel<<<(N + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x_dev, N, D, eps, out_dev);
  SD_CUDA_OK(cudaGetLastError());
  count_launch();
  return SD_OK;
}

extern "C" int sd_ecapa_profile(SdEcapaPlan* p, int enable) {
  if (!p) return fail(SD_ERR_ARG, "sd_ecapa_profile: NULL plan");
  p->profile = enable != 0;
  p->ev_used = 0;
  p->forwards_profiled = 0;
  return SD_OK;
}

extern "C" int sd_ecapa_profile_read(SdEcapaPlan* p, int max_stages, char* names, float* total_ms,
                                     int* n_forwards) {
  if (!p || !names || !total_ms || !n_forwards || max_stages < kNumStages)
    return fail(SD_ERR_ARG, "sd_ecapa_profile_read: need room for %d stages", kNumStages);
  const int per = kNumStages + 1;  // events per forward
  const int nf = p->forwards_profiled;
  if ((size_t)nf * per != p->ev_used) return fail(SD_ERR_ARG, "profile events out of step (%zu events, %d forwards)", p->ev_used, nf);
  for (int s = 0; s < kNumStages; ++s) {
    total_ms[s] = 0.f;
    snprintf(names + 32 * s, 32, "%s", kStageNames[s]);
  }
  if (nf > 0) SD_CUDA_OK(cudaEventSynchronize(p->ev_pool[p->ev_used - 1]));
  for (int f = 0; f < nf; ++f)
    for (int s = 0; s < kNumStages; ++s) {
      float ms = 0.f;
      SD_CUDA_OK(cudaEventElapsedTime(&ms, p->ev_pool[(size_t)f * per + s], p->ev_pool[(size_t)f * per + s + 1]));
      total_ms[s] += ms;
    }
  *n_forwards = nf;
  return kNumStages == 0 ? SD_OK : SD_OK;
}

extern "C" int sd_ecapa_num_stages(void) { return kNumStages; }
