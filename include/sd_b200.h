/* sd_b200.h — C ABI of libsd_b200.so: the B200 (sm_100a) implementation of the
 * speaker-diarization hot path of hzane/speech-diarization
 * (log-mel fbank -> ECAPA-TDNN -> cosine affinity -> AHC).
 *
 * The reference reaches this path through Python callables only
 * (SURVEY.md §8b); there is no native FFI in the reference.  Each entry point
 * below therefore cites the reference callable whose arithmetic it replaces,
 * and the modules of speech_diarization_b200 re-export those callables on top of this
 * ABI through ctypes (INTEGRATION.md shows the binding).
 *
 * Conventions
 *  - every function returns an int status, SD_OK (0) on success; it never
 *    throws and never falls back to a CPU path.  sd_last_error() returns a
 *    thread-local detail string for the last failure.
 *  - pointers named *_dev are device pointers on the current CUDA device;
 *    `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - calls are asynchronous with respect to the host unless stated otherwise.
 *  - all matrices are row-major, f32 unless stated otherwise.
 */
#ifndef SD_B200_H_
#define SD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_OK 0
#define SD_ERR_ARG 1         /* bad argument (shape, alignment, NULL) */
#define SD_ERR_CUDA 2        /* a CUDA runtime call or kernel launch failed */
#define SD_ERR_DRIVER 3      /* cuTensorMapEncodeTiled unavailable / failed */
#define SD_ERR_NOMEM 4       /* device allocation failed */
#define SD_ERR_UNSUPPORTED 5 /* valid request outside what this build implements */
#define SD_ERR_MISSING 6     /* a required weight tensor is missing or mis-sized */
#define SD_ERR_RANGE 7       /* an activation left the f16 range: results of the call are NaN, not silently wrong */

#define SD_EMB_DIM 192  /* ecapa_annote.py:11 (self.dimension = 192) */
#define SD_N_MELS 80

/* Library version (major*10000 + minor*100 + patch). */
int sd_version(void);
const char* sd_status_string(int status);
const char* sd_last_error(void);
/* Number of CUDA kernels this library has launched since it was loaded. */
long sd_launch_count(void);

/* ------------------------------------------------------------------ fbank ---
 * Number of frames a centre-padded STFT (n_fft = win = 400, hop = 160) yields:
 * T = 1 + n_samples / 160.  (torch.stft via torchaudio MelSpectrogram,
 * speech_encode.py:17-30.) */
int sd_fbank_num_frames(int n_samples);

/* Log-mel filterbank features of B windows.
 * Replaces fbank_batch (speech_encode.py:10-38) when variant == SD_FBANK_TORCHAUDIO:
 *   reflect centre pad 200, periodic Hann(400), |rFFT|^2, 80 HTK triangles on
 *   [20, 7900] Hz, log(x + 1e-6), and, if mean_norm, minus the per-(window, mel)
 *   time mean (speech_encode.py:35-36).
 * variant == SD_FBANK_SPEECHBRAIN is the front end inside
 *   EncoderClassifier.encode_batch (call sites speech_encode.py:77,
 *   ecapa_annote.py:22, diar_diag.py:169): zero centre pad, periodic Hamming(400),
 *   |rFFT|^2, speechbrain triangular filters on [0, 8000] Hz, 10 log10(max(x, 1e-10)),
 *   floor at (window max - 80 dB), then sentence mean normalisation (mean_norm).
 * Window b starts at wav_dev + b * wav_stride (elements), so overlapping windows of
 * one audio buffer are addressed in place (wav_stride = hop) instead of being
 * materialised as vad.py:9-16 / frame_audio does.
 * out_dev: [B, T, 80] f32.  Requires n_samples >= 400 (201 for the reflect pad). */
#define SD_FBANK_TORCHAUDIO 0
#define SD_FBANK_SPEECHBRAIN 1
int sd_fbank_f32(const float* wav_dev, long wav_stride, int B, int n_samples, int variant,
                 int mean_norm, float* out_dev, void* stream);
/* Measurement switch: which = 1 selects the tensor-core DFT frames kernel (default; $SD_FBANK_TC=0 starts with the
 * other), which = 0 the FFT kernel on the FP32 pipe, which < 0 only queries.  Returns the kernel now selected.  Both
 * implement the same operator; a change applies to later calls (a captured ECAPA plan graph keeps what it recorded). */
int sd_fbank_kernel(int which);

/* ------------------------------------------------------------- ECAPA-TDNN ---
 * A plan owns the f16-repacked weights (speechbrain ECAPA_TDNN, C = 1024,
 * attention 128, embedding 192), folded BatchNorm constants, activation
 * workspace for up to max_batch windows of max_samples samples, and cached TMA
 * descriptors.  Replaces the model behind using_ecapa_encoder
 * (speech_encode.py:64-70).
 *
 * Weights are passed as a speechbrain-keyed state dict: names[i] is the key
 * (e.g. "blocks.0.conv.conv.weight", "blocks.1.res2net_block.blocks.0.norm.norm.running_var",
 * "mfa.conv.conv.weight", "asp.tdnn.conv.conv.weight", "asp.conv.conv.weight",
 * "asp_bn.norm.weight", "fc.conv.weight"), tensors[i] a HOST pointer to the f32
 * data in PyTorch layout, numels[i] its element count.  Synchronous. */
typedef struct SdEcapaPlan SdEcapaPlan;
int sd_ecapa_plan_create(const char* const* names, const float* const* tensors,
                         const int64_t* numels, int n_tensors, int max_batch, int max_samples,
                         SdEcapaPlan** plan_out);
int sd_ecapa_plan_destroy(SdEcapaPlan* plan);

/* Embeddings of B windows: speechbrain fbank + sentence-mean norm + ECAPA-TDNN.
 * Replaces EncoderClassifier.encode_batch(wavs).squeeze(1) as called by
 * ecapa_encode_batch (speech_encode.py:73-78) and ECAPAEncoder.forward
 * (ecapa_annote.py:13-22); like them it uses no wav_lens (zero padding is
 * signal) and does not L2-normalise unless l2_normalize != 0
 * (then e / (||e|| + 1e-8), the consumers' convention,
 * anti_stick_diarize.py:176,430).
 * emb_dev: [B, 192] f32.  B <= max_batch, 400 <= n_samples <= max_samples. */
int sd_ecapa_embed(SdEcapaPlan* plan, const float* wav_dev, long wav_stride, int B, int n_samples,
                   int l2_normalize, float* emb_dev, void* stream);

/* The same for windows that start at ARBITRARY sample offsets of one device-resident recording: window b =
 * wav_dev[offsets_dev[b] : offsets_dev[b] + n_samples] (offsets_dev: B x int64 on the device; the caller
 * guarantees the windows lie inside the buffer).  One launch sequence for all the speech windows of a recording
 * (anti_stick_diarize.py:420-427 gathers them with a Python list comprehension per batch of 128) or all the sliding
 * windows of all the segments of scd_split_segments (:97-100, one encoder call per segment). */
int sd_ecapa_embed_offsets(SdEcapaPlan* plan, const float* wav_dev, const long* offsets_dev, int B, int n_samples,
                           int l2_normalize, float* emb_dev, void* stream);

/* The same from HOST memory — the call behind ecapa_encode_batch(np.ndarray) (speech_encode.py:73-78) and
 * `emb = encode_batch(wav).cpu().numpy()`: wav_host / emb_host are host pointers (pinned or pageable), window
 * b = wav_host[b * wav_stride : b * wav_stride + n_samples], emb_host [B, 192] f32.  The host->device copy is
 * issued in four chunks on a plan-owned copy stream and each chunk's fbank kernels start as soon as it has
 * landed, so all but the first quarter of the upload runs under compute; the embeddings are copied back and
 * the stream is synchronised before returning (the reference's call ends with the same implicit sync). */
int sd_ecapa_embed_host(SdEcapaPlan* plan, const float* wav_host, long wav_stride, int B, int n_samples,
                        int l2_normalize, float* emb_host, void* stream);
/* Pageable wav_host (ordinary numpy memory, what the reference's callers pass) is staged through a page-locked
 * ring filled by a few copy threads (SD_ECAPA_HOST_THREADS, default 6), so the transfer is pipelined instead of
 * a blocking driver-staged copy; page-locked memory is read by the copy engine directly.
 * Returns SD_ERR_RANGE (and NaN embeddings) when an activation left the f16 range — see sd_ecapa_overflow. */

/* Activations are stored as f16.  A value beyond +-65504 (a checkpoint with an unusually large BatchNorm scale) is
 * saturated, never turned into inf, and raises a flag: the embeddings of that forward are written as NaN so the
 * failure is loud on the asynchronous device path too.  This call synchronises `stream` and reports (and, with
 * reset != 0, clears) the sticky flag. */
int sd_ecapa_overflow(SdEcapaPlan* plan, int reset, int* flag_out, void* stream);

/* Same trunk from precomputed features feats_dev [B, T, 80] f32 (already
 * normalised); used by the parity tests to isolate the trunk. */
int sd_ecapa_forward_feats(SdEcapaPlan* plan, const float* feats_dev, int B, int T,
                           int l2_normalize, float* emb_dev, void* stream);

/* Test hook: copy an internal activation of the LAST forward to out_dev as f32
 * [B, T, C] (interior frames, channels-last) for name in {"feats" (C=80), "block0",
 * "b1.out", "b2.out", "b3.out", "b3.tdnn1", "b3.res2net", "b3.tdnn2" (C=1024),
 * "mfa" (C=3072), "asp.attn" (C=128)}, or [B, C] for {"b3.se" (1024), "asp.stats"
 * (6144 = mean|std), "asp.uttbias" (128), "pooled" (6144)}.  *C_out receives C. */
int sd_ecapa_debug_fetch(SdEcapaPlan* plan, const char* name, float* out_dev, int* C_out,
                         void* stream);

/* Per-stage device timing of the forward pass: when enabled, CUDA events are recorded on the
 * caller's stream at every stage boundary of each subsequent sd_ecapa_embed /
 * sd_ecapa_forward_feats.  sd_ecapa_profile(plan, 0/1) also resets the record.
 * sd_ecapa_profile_read waits for the last event and returns, for each of
 * sd_ecapa_num_stages() stages, its name (32 bytes per entry in `names`) and the elapsed
 * milliseconds SUMMED over the *n_forwards forwards recorded. */
int sd_ecapa_profile(SdEcapaPlan* plan, int enable);
int sd_ecapa_num_stages(void);
int sd_ecapa_profile_read(SdEcapaPlan* plan, int max_stages, char* names, float* total_ms,
                          int* n_forwards);

/* FLOPs (2 * MACs of the dense contractions actually issued, padding excluded)
 * of one window of T frames — the figure bench.py's roofline uses. */
double sd_ecapa_flops_per_window(int T);

/* ------------------------------------------------- affinity and clustering ---
 * out[i, :] = x[i, :] / (||x[i, :]|| + eps)   (anti_stick_diarize.py:176,203,430) */
int sd_l2norm_f32(const float* x_dev, int N, int D, float eps, float* out_dev, void* stream);

/* Cosine DISTANCE row block:  out[i - row0, j] = 1 - cos(x_i, x_j), row0 <= i < row0 + rows,
 * 0 <= j < N — the `D = 1 - cosine_similarity(embs)` of cluster_embeddings
 * (diar_diag.py:215,219) and cluster_hdbscan (anti_stick_diarize.py:177), with
 * sklearn's internal row normalisation (zero rows stay zero).  Tensor-core GEMM
 * on a split-f16 (hi + lo*2^-11) representation; |error| <= 1e-5 absolute.
 * emb_dev [N, D] f32, D % 64 == 0, D <= 512.  out_dev [rows, N] f32 (ld = N).
 * out_f64_dev, if not NULL, receives the same values widened to f64 (the AHC
 * working matrix).  workspace_dev: sd_affinity_workspace_bytes(N, D) bytes. */
size_t sd_affinity_workspace_bytes(int N, int D);
int sd_cosine_distance_rowblock(const float* emb_dev, int N, int D, int row0, int rows,
                                float* out_dev, double* out_f64_dev, void* workspace_dev,
                                void* stream);

/* Average-linkage agglomerative clustering of a precomputed distance matrix with
 * a distance threshold: AgglomerativeClustering(n_clusters=None,
 * linkage="average", metric="precomputed", distance_threshold=threshold)
 * .fit_predict(D) (diar_diag.py:221-226).  Merges every pair whose linkage
 * distance is < threshold; Lance-Williams updates in f64 on the f32-rounded
 * input, as scipy does.  Labels are numbered by each cluster's smallest member
 * (0, 1, ... in order of first appearance) — equal to sklearn's up to a
 * permutation.
 * dist_dev: [N, N] f32 symmetric (only read).  labels_dev: [N] int32.
 * n_clusters_dev: [1] int32.  workspace_dev: sd_ahc_workspace_bytes(N) bytes. */
size_t sd_ahc_workspace_bytes(int N);
int sd_ahc_average_f32(const float* dist_dev, int N, double threshold, int32_t* labels_dev,
                       int32_t* n_clusters_dev, void* workspace_dev, void* stream);

/* Diagnostics of the LAST sd_ahc_average_f32 run on this workspace (synchronous D2H read):
 * number of reciprocal-nearest-neighbour rounds and of merges performed. */
int sd_ahc_read_stats(const void* workspace_dev, int N, int32_t* rounds, int32_t* merges);

/* Centroid-linkage agglomerative clustering: the linkage matrix
 * scipy.cluster.hierarchy.linkage(X, method="centroid", metric="euclidean") computes inside pyannote's
 * AgglomerativeClustering, which diarization_baseline.py:176-180,252-257 drives through
 * `pipeline.clustering.threshold`.  x_dev [N, D] f32 (pyannote passes unit-normalised embeddings).
 * Z_dev [N-1, 4] f64 in scipy's layout: (cluster id a < b, cluster id b, centroid distance, size); ids
 * >= N name earlier merges.  Merges are taken strictly in order of the global minimum (centroid linkage is
 * not monotone; inversions are kept).  workspace_dev: sd_centroid_linkage_workspace_bytes(N, D) bytes. */
size_t sd_centroid_linkage_workspace_bytes(int N, int D);
int sd_centroid_linkage_f64(const float* x_dev, int N, int D, double* Z_dev, void* workspace_dev, void* stream);

/* best[i] = argmax_k <x_i, c_k>, score[i] = that maximum (frame_reassign,
 * anti_stick_diarize.py:433-434).  x_dev [N, D], cent_dev [K, D], K <= 64.
 * First maximum wins, as numpy.argmax. score_dev may be NULL. */
int sd_window_argmax(const float* x_dev, const float* cent_dev, int N, int K, int D,
                     int32_t* best_dev, float* score_dev, void* stream);

/* sims[i] = <x_i, x_{i+1}> / (||x_i|| ||x_{i+1}|| + 1e-8), 0 <= i < N - 1
 * (scd_split_segments, anti_stick_diarize.py:102-104). */
int sd_adjacent_cosine(const float* x_dev, int N, int D, float* sims_dev, void* stream);

/* ------------------------------------- score post-processing and VAD masks ---
 * (the small stages either side of the embedding / clustering path.)
 *
 * Viterbi decoding of a sticky K-state HMM over per-step scores — viterbi_hmm(scores, alpha)
 * (diar_diag.py:231-247; call site :393).  scores_dev [T, K] f32 (scores_f64 = 0) or f64 (= 1);
 * log_stay = f32(log(alpha + 1e-8)), log_move = f32(log((1 - alpha) / (K - 1) + 1e-8)) as the reference
 * builds logA.  The recursion is the reference's float32 one, operation for operation (one rounded add
 * per candidate, first maximum wins), so path_dev [T] int32 is identical, ties included.  1 <= K <= 32.
 * workspace_dev: sd_viterbi_workspace_bytes(T, K) bytes. */
size_t sd_viterbi_workspace_bytes(int T, int K);
int sd_viterbi_hmm(const void* scores_dev, int scores_f64, int T, int K, float log_stay, float log_move,
                   int32_t* path_dev, void* workspace_dev, void* stream);

/* Adaptive symmetric score normalisation — asnorm_scores(query_embs, ref_centers, cohort_embs, topk)
 * (diar_diag.py:196-208; call site :389 passes the segment embeddings as their own cohort).
 * q_dev [nq, D], r_dev [nr, D], c_dev [nc, D] f32; out_dev [nq, nr] f32 =
 * 0.5 * ((raw - mu_q) / sigma_q + (raw - mu_r) / sigma_r) with raw = Qn Rn^T, and mu / sigma (+1e-6) the
 * mean / population std of each row's min(topk, nc) largest cohort similarities.  Rows are normalised as
 * x / (||x|| + 1e-9).  Cohort similarities use the split-f16 tensor-core kernel of
 * sd_cosine_distance_rowblock; statistics are accumulated in f64.  D % 64 == 0, D <= 512.
 * workspace_dev: sd_asnorm_workspace_bytes(nq, nr, nc, D) bytes. */
size_t sd_asnorm_workspace_bytes(int nq, int nr, int nc, int D);
int sd_asnorm_scores(const float* q_dev, const float* r_dev, const float* c_dev, int nq, int nr, int nc,
                     int D, int topk, float* out_dev, void* workspace_dev, void* stream);

/* ZCA whitening + L2 normalisation of segment embeddings — whiten_l2(embs) (diar_diag.py:187-194; call site
 * :352, between embedding and clustering): X = embs - mean (f32); C = cov(X) (f64); W = (C + 1e-6 I)^(-1/2)
 * through the eigendecomposition of C (one-sided Jacobi on the device; the reference's SVD of the symmetric
 * PSD C is the same decomposition); out = X W, rows divided by (norm + 1e-9).  x_dev [N, D] f32, out_dev [N, D]
 * f64.  N >= 2, D % 32 == 0, D <= 192.  sweeps_host, if not NULL, receives the number of Jacobi sweeps
 * (synchronises the stream).  workspace_dev: sd_whiten_workspace_bytes(N, D) bytes. */
size_t sd_whiten_workspace_bytes(int N, int D);
int sd_whiten_l2_f64(const float* x_dev, int N, int D, double* out_dev, void* workspace_dev, int32_t* sweeps_host,
                     void* stream);

/* Unit-norm cluster centres — `m = embs[labels == k].mean(0); m /= norm(m) + 1e-9` for k = 0 .. K-1
 * (diar_diag.py:377-383).  x_dev [N, D] f64 (the whitened embeddings), labels_dev [N] int32 in [0, K);
 * out_f64_dev / out_f32_dev [K, D] (either may be NULL).  D <= 256. */
int sd_cluster_centers_f64(const double* x_dev, const int32_t* labels_dev, int N, int D, int K,
                           double* out_f64_dev, float* out_f32_dev, void* stream);

/* scores[i, k] = <x_i, c_k> — `scores = embs @ centers.T` (diar_diag.py:386).  x_dev [N, D], cent_dev [K, D],
 * out_dev [N, K], all f32. */
int sd_dot_scores(const float* x_dev, const float* cent_dev, int N, int K, int D, float* out_dev, void* stream);

/* hysteresis_binarize(probs, on, off) (vad.py:59-74; diar_diag.py:331): mask[i] = talking after frame i,
 * where a silent state turns on at p >= on and a talking state turns off at p < off (compared in f64, as
 * numba does).  probs_dev [n] f32 (probs_f64 = 0) or f64 (= 1); mask_dev [n] u8 (0 / 1).  A parallel scan
 * over state maps; bit-exact. */
int sd_hysteresis_u8(const void* probs_dev, int probs_f64, int n, double on, double off,
                     uint8_t* mask_dev, void* stream);

/* morph_open_close(mask, hop_ms, open_ms, close_ms) (vad.py:77-87): scipy.ndimage binary_opening then
 * binary_closing with flat structures of open_w / close_w frames (0 = skip), border value 0.
 * mask_dev, out_dev, tmp_dev: [n] u8; out_dev and tmp_dev must not alias mask_dev.  Bit-exact. */
int sd_morph_open_close_u8(const uint8_t* mask_dev, int n, int open_w, int close_w, uint8_t* out_dev,
                           uint8_t* tmp_dev, void* stream);

/* The frame-index part of mask_to_segments (vad.py:90-151): runs of ones, runs shorter than
 * min_speech_frames dropped, neighbours with a gap <= min_gap_frames merged.  seg_dev receives
 * [count][2] int32 (start frame, end frame exclusive), capacity n/2 + 1 pairs; count_dev [1] int32.
 * Padding and the conversion to rounded seconds (vad.py:153-161) stay with the caller.
 * workspace_dev: sd_mask_segments_workspace_bytes(n) bytes. */
size_t sd_mask_segments_workspace_bytes(int n);
int sd_mask_segments_i32(const uint8_t* mask_dev, int n, int min_speech_frames, int min_gap_frames,
                         int32_t* seg_dev, int32_t* count_dev, void* workspace_dev, void* stream);

/* ---------------------------------------------- dense-pass operators (f2) ---
 * The operators either side of the embedding kernels in anti_stick_diarize.py's SCD and frame-reassignment passes. */

/* scd_split_segments (anti_stick_diarize.py:102-116) for ALL segments at once.  emb_dev [total, D] f32 holds the
 * sliding-window embeddings of segment s in rows seg_off_dev[s] .. seg_off_dev[s+1] (int32, n_segments + 1
 * entries).  Per segment: d_i = 1 - cos(e_i, e_{i+1}) (denominator + 1e-8), z = (d - mean(d)) / std(d) when
 * std > 1e-6 else d, and peak_dev[seg_off[s] + i] = 1 where scipy.signal.find_peaks(z, height=thr) has a peak
 * (strict local maxima and the floor-middle of flat tops; never the first or last distance).  z_dev [total] f32
 * receives z (scratch the caller may inspect).  One CTA per segment. */
int sd_scd_peaks(const float* emb_dev, int D, const int32_t* seg_off_dev, int n_segments, float thr,
                 float* z_dev, uint8_t* peak_dev, void* stream);

/* speaker_centroids (anti_stick_diarize.py:333-349): out_dev[k] = mean of the rows with labels_dev[i] ==
 * spk_ids_dev[k], divided by (its norm + 1e-8).  D <= 256. */
int sd_speaker_centroids(const float* emb_dev, const int32_t* labels_dev, int N, int D,
                         const int32_t* spk_ids_dev, int K, float* out_dev, void* stream);

/* full_dev[valid_dev[i]] = label_map_dev ? label_map_dev[labels_dev[i]] : labels_dev[i]  for i < m
 * (anti_stick_diarize.py:374-375, :435; the caller pre-fills full_dev with -1). */
int sd_scatter_labels(const int32_t* valid_dev, const int32_t* labels_dev, const int32_t* label_map_dev, int m,
                      int32_t* full_dev, void* stream);

/* _labels_to_segments (anti_stick_diarize.py:377-386): run-length encoding of labels_dev[0, n) (-1 = no speech).
 * A run [i, j) of speaker k becomes start = window_starts_dev[i] / sr, end = j < n ? window_starts_dev[j] / sr :
 * max_t (f64, the reference's own expressions) and is kept when k != -1 and end > start.
 * run_idx_dev [count][3] int32 = (i, j, k); run_t_dev [count][2] f64 = (start, end); count_dev [1];
 * scratch_dev: n int32.  Capacity n rows each.  Single CTA of block scans (n <= a few 10^5 windows). */
int sd_label_runs(const int32_t* labels_dev, int n, const int64_t* window_starts_dev, double sr, double max_t,
                  int32_t* scratch_dev, int32_t* run_idx_dev, double* run_t_dev, int32_t* count_dev, void* stream);

/* merge_adjacent (anti_stick_diarize.py:464-475) on segments (seg_t_dev [n][2] f64 start/end, speaker of segment k
 * at spk_dev[k * spk_stride]): segment k joins its predecessor when the speakers are equal and
 * start[k] - end[k-1] <= gap.  group_dev [count][2] int32 = (first, last) segment of each merged group.
 * n is taken from *n_dev when n_dev != NULL (chained behind sd_label_runs without a host round trip; pass the
 * capacity as n). */
int sd_merge_adjacent(const double* seg_t_dev, const int32_t* spk_dev, int spk_stride, int n, const int32_t* n_dev,
                      double gap, int32_t* group_dev, int32_t* count_dev, void* stream);

/* Zero-padded batch of variable-length snippets of one device-resident recording (embed_segments,
 * anti_stick_diarize.py:162-166): out_dev[b, i] = audio_dev[start_dev[b] + i] for i < len_dev[b], else 0. */
int sd_gather_pad_f32(const float* audio_dev, const int64_t* start_dev, const int32_t* len_dev, int B, int max_len,
                      float* out_dev, void* stream);

/* ------------------------------------------------------------ debug / test ---
 * Raw tensor-core GEMM used by the unit tests of the tcgen05 kernel:
 *   D[m, n] = sum_{j < taps} sum_{k < K} A[m + (j - taps/2) * dil, k] * B[n, j*K + k]
 * A [M, K] f16 (rows outside [0, M) read as zero), B [N, taps*K] f16, D [M, N] f32.
 * K % 64 == 0, n_tile % 16 == 0 and <= 256. */
int sd_debug_gemm_f16(const void* A_dev, int M, int K, const void* B_dev, int N, int taps, int dil,
                      int n_tile, float* D_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SD_B200_H_ */
