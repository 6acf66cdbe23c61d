/* b200sep.h -- C ABI of libb200sep.so: B200-native (sm_100a) kernels for the speech-separation
 * hot path of fgnt/padertorch (STFT / iSTFT front-end, mask (*) spectrogram, permutation-invariant
 * MSE, deep-clustering affinity loss, SI-SDR / SDR / log-MSE pair statistics).
 *
 * The reference has no FFI on this path: its boundary is the Python call surface
 * `padertorch.ops.*`.  Every entry point below names the reference
 * function (file:line under the padertorch repository) whose arithmetic it replaces; the Python
 * mirror of that call surface lives in `padertorch_b200/ops` and binds these symbols with ctypes
 * (INTEGRATION.md shows the stub a padertorch maintainer would add).
 *
 * Conventions
 *   - all data pointers are DEVICE pointers owned by the caller (PyTorch's caching allocator in the
 *     Python host); the library never frees or retains them.  `meta` arrays are device int64.
 *   - `stream` is a cudaStream_t passed as void*; nothing here synchronises the device.
 *   - every function returns 0 on success, a negative code on failure; the message of the last
 *     failure on the calling thread is returned by b2s_last_error().
 *   - no CPU fallback exists: without a CUDA device every compute entry point fails.
 *   - reductions use fixed trees (no floating-point atomics): results are run-to-run deterministic,
 *     which padertorch's Trainer.test_run requires (padertorch/train/runtime_tests.py:317-330).
 */
#ifndef B200SEP_H_
#define B200SEP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_VERSION 100

#define B2S_OK 0
#define B2S_ERR_ARGUMENT (-1)
#define B2S_ERR_CUDA (-2)
#define B2S_ERR_UNSUPPORTED (-3)

#if defined(__GNUC__)
#define B2S_API __attribute__((visibility("default")))
#else
#define B2S_API
#endif

typedef void* b2s_stream; /* cudaStream_t */
typedef struct b2s_stft_plan b2s_stft_plan;

B2S_API int b2s_version(void);
B2S_API const char* b2s_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Frame arithmetic (host, integer, bit exact).  fading: 0 = None/False, 1 = 'full'/True, 2 = 'half'.
 * Replaces STFT.samples_to_frames / frames_to_samples / sample_index_to_frame_index
 * (padertorch/ops/_stft.py:265-307, which delegate to paderbox _samples_to_stft_frames etc.). */
B2S_API int64_t b2s_stft_frames(int64_t samples, int window_length, int shift, int pad, int fading);
B2S_API int64_t b2s_stft_samples(int64_t frames, int window_length, int shift, int fading);
B2S_API int64_t b2s_stft_frame_index(int64_t sample_index, int window_length, int shift, int fading);

/* ---------------------------------------------------------------------------------------------
 * STFT plans: immutable per-(device, size, shift, window) tables (fp64 -> fp32 once, like
 * get_stft_kernel / get_istft_kernel build their matrices in fp64, padertorch/ops/_stft.py:11-43).
 * analysis_window / synthesis_window: HOST double[window_length]; the synthesis window is the
 * biorthogonal window divided by `size` (ops/_stft.py:27-28).                                    */
B2S_API int b2s_stft_plan_create(b2s_stft_plan** plan, int device, int size, int shift, int window_length,
                         const double* analysis_window, const double* synthesis_window);
B2S_API int b2s_stft_plan_destroy(b2s_stft_plan* plan);
/* 1 if the plan runs the register-resident warp FFT (size 1024), 0 for the table-driven DFT.    */
B2S_API int b2s_stft_plan_is_fast(const b2s_stft_plan* plan);
/* bytes of scratch the inverse-type calls (b2s_istft_forward, b2s_stft_backward) need for (rows, frames).
 * Fast plans with shift 256 (the ring kernel): [ticket counters | partial sums of the chunk boundaries] -- the
 * buffer must be ZERO-FILLED ONCE before its first use; every call leaves the counters at zero, so one buffer
 * serves all later calls of any geometry on the same stream (not two calls that run concurrently).  Other fast
 * plans: 0.  Generic plans: the windowed frames [rows][frames][window_length], no initialisation needed.      */
B2S_API int64_t b2s_stft_scratch_bytes(const b2s_stft_plan* plan, int64_t rows, int64_t frames);

/* spectrum layouts (last axes of the output), F = size/2 + 1 */
#define B2S_SPEC_INTERLEAVED 0 /* [rows, frames, F, 2]  == torch complex64 == 'stacked'           */
#define B2S_SPEC_CONCAT 1      /* [rows, frames, 2F]    real bins then imaginary bins ('concat')  */
#define B2S_SPEC_ABS 2         /* [rows, frames, F]     |Y|          (fused magnitude epilogue)  */
#define B2S_SPEC_LOG1P_ABS 3   /* [rows, frames, F]     log1p(|Y|)   (fused PIT-model feature)    */
#define B2S_SPEC_FEATURE 4     /* internal: the parameterised feature epilogue of b2s_stft_features */

/* STFT.__call__ (padertorch/ops/_stft.py:103-174).  signal [rows, samples] (row stride given in
 * floats); frame m covers padded samples m*shift .. m*shift+window_length-1 where padded index p
 * maps to signal index p - pad_left and everything outside [0, samples) reads as zero (this supplies
 * the fading pads :137-146 and the tail pad :148-154 without materialising them).                 */
B2S_API int b2s_stft_forward(const b2s_stft_plan* plan, const float* signal, int64_t rows, int64_t samples,
                     int64_t row_stride, int64_t pad_left, int64_t frames, int layout, float* spec,
                     b2s_stream stream);
/* autograd adjoint of b2s_stft_forward for layouts INTERLEAVED / CONCAT (the reference gets it from
 * conv1d's backward): grad_signal [rows, samples] dense.                                           */
B2S_API int b2s_stft_backward(const b2s_stft_plan* plan, const float* grad_spec, int64_t rows,
                      int64_t frames, int layout, int64_t pad_left, int64_t samples,
                      float* grad_signal, float* scratch, b2s_stream stream);
/* STFT.inverse (padertorch/ops/_stft.py:176-263): Hermitian extension + synthesis-windowed inverse
 * DFT + overlap-add; output sample n is padded sample n + crop_left (the cropped fading :257-262).  */
B2S_API int b2s_istft_forward(const b2s_stft_plan* plan, const float* spec, int64_t rows, int64_t frames,
                      int layout, int64_t crop_left, int64_t samples_out, float* signal,
                      float* scratch, b2s_stream stream);
/* autograd adjoint of b2s_istft_forward (conv_transpose1d's backward in the reference).            */
B2S_API int b2s_istft_backward(const b2s_stft_plan* plan, const float* grad_signal, int64_t rows,
                       int64_t samples_out, int64_t crop_left, int64_t frames, int layout,
                       float* grad_spec, b2s_stream stream);

/* ---------------------------------------------------------------------------------------------
 * Feature epilogues fused behind the transform, replacing the ATen chains of
 * padertorch/contrib/mk/modules/features/timefreq.py:
 *   to_spectrogram (:171-183)   x = |Y|^power * scale      (scale = 1/size for scale_spec, else 1)
 *   Logarithm (:37-77)          log_b(max(eps, x)), b in {none, e, 10, 2}
 *   MelTransform.forward (:398-470, :444)   x @ mel_basis before the logarithm.
 * Without a filterbank features are [rows, frames, F]; with one [rows, frames, filters] (the spectrogram never
 * reaches HBM: 4T + 4 M filters algorithmic bytes per utterance).  Fast plans only (size 1024).
 * b2s_mel_create: basis = HOST float [bins][filters] row-major (MelTransform.mel_basis, :338).             */
typedef struct b2s_mel b2s_mel;
#define B2S_LOG_NONE 0
#define B2S_LOG_E 1
#define B2S_LOG_10 2
#define B2S_LOG_2 3
B2S_API int b2s_mel_create(b2s_mel** mel, int device, int bins, int filters, const float* basis);
B2S_API int b2s_mel_destroy(b2s_mel* mel);
B2S_API int b2s_stft_features(const b2s_stft_plan* plan, const float* signal, int64_t rows, int64_t samples,
                      int64_t row_stride, int64_t pad_left, int64_t frames, float power, float scale,
                      int log_kind, float eps, const b2s_mel* mel, float* features, b2s_stream stream);

/* ---------------------------------------------------------------------------------------------
 * Permutation-invariant MSE over spectrogram-shaped data.
 * Replaces the per-example loop of PermutationInvariantTrainingModel.review
 * (padertorch/contrib/examples/source_separation/pit/model.py:117-135) around
 * pit_loss(..., loss_fn=mse_loss) (padertorch/ops/losses/source_separation.py:34-124):
 *     estimate[t,i,f] = mask[t,i,f] * observation[t,f]          (observation may be NULL: factor 1)
 *     SSE[i][j]       = sum_{t,f} (estimate[t,i,f] - target[t,j,f] * scale[t,j,f])^2   (scale NULL: 1)
 *     loss            = min_perm  sum_k SSE[perm[k]][k] / (frames*K*F);  first minimum in
 *                       itertools.permutations order wins (torch.min on CPU, :119)
 * meta: device int64 [B][B2S_PIT_META] = {frames, off_mask, off_observation, off_target, off_scale,
 * off_grad} with offsets in floats relative to the respective base pointers (ragged lists, padded
 * batches and separately allocated per-example tensors all fit; off_grad places the example's block in
 * the gradient buffers of the backward call).  mask/target blocks are [frames, K, F], observation
 * blocks [frames, F], contiguous.
 * dual != 0 additionally evaluates the loss against target*scale in the same pass (the model's
 * pit_mse_loss and pit_ips_loss): outputs then hold [2][B] losses, [2][B][K] permutations,
 * [B][2][K][K] SSE matrices (slot 0: plain target, slot 1: scaled target).
 * workspace: b2s_pit_workspace_bytes() bytes, zero-filled once by the caller; calls leave it zeroed. */
#define B2S_PIT_META 6
#define B2S_MAX_SOURCES 8
B2S_API int64_t b2s_pit_workspace_bytes(int64_t batch, int64_t max_frames, int64_t bins, int sources, int dual);
B2S_API int b2s_pit_sse_forward(const float* mask, const float* observation, const float* target,
                        const float* scale, const int64_t* meta, int64_t batch, int64_t max_frames,
                        int sources, int64_t bins, int dual, float* loss, int32_t* perm,
                        double* sse, void* workspace, b2s_stream stream);
/* gradient of sum_b grad_loss[.][b] * loss[.][b] w.r.t. mask (and, when grad_target != NULL, w.r.t.
 * target; dual must then be 0).  Example b's gradient block starts at off_grad in both buffers.     */
B2S_API int b2s_pit_sse_backward(const float* mask, const float* observation, const float* target,
                         const float* scale, const int64_t* meta, int64_t batch, int64_t max_frames,
                         int sources, int64_t bins, int dual, const int32_t* perm,
                         const float* grad_loss, float* grad_mask, float* grad_target,
                         b2s_stream stream);
/* The same with the upstream gradient grad_scale * grad_loss[...]: grad_loss_stride 1 = [slots][batch] as above,
 * 0 = ONE value per slot ([slots]; the gradients of the batch means of b2s_pit_sse_forward_mean, scale 1 / batch).   */
B2S_API int b2s_pit_sse_backward_scaled(const float* mask, const float* observation, const float* target,
                         const float* scale, const int64_t* meta, int64_t batch, int64_t max_frames,
                         int sources, int64_t bins, int dual, const int32_t* perm,
                         const float* grad_loss, int64_t grad_loss_stride, double grad_scale,
                         float* grad_mask, float* grad_target, b2s_stream stream);
/* b2s_pit_sse_forward followed by mean[slot] = mean_b loss[slot][b] (the `losses` entry of
 * PermutationInvariantTrainingModel.review, pit/model.py:137-140): one extra one-warp-per-slot launch instead of a
 * reduction per loss; fixed summation order.  batch >= 1.                                                        */
B2S_API int b2s_pit_sse_forward_mean(const float* mask, const float* observation, const float* target,
                         const float* scale, const int64_t* meta, int64_t batch, int64_t max_frames,
                         int sources, int64_t bins, int dual, float* loss, float* mean, int32_t* perm,
                         double* sse, void* workspace, b2s_stream stream);

/* ---------------------------------------------------------------------------------------------
 * Pair statistics for the time-domain regression losses
 * (padertorch/ops/losses/regression.py:4-376) and their PIT wrapper as TasNet.loss uses it
 * (padertorch/contrib/examples/source_separation/tasnet/model.py:154-176).
 * For group g (an example, or an example x middle index) with rows e_i, t_j of `length` samples:
 *     stats[g] = { D[i][j] = <e_i,t_j> (K*K), Ee[i] = |e_i|^2, Tt[j] = |t_j|^2, Se[i] = sum e_i,
 *                  St[j] = sum t_j }                       (double, K*K + 4K values per group)
 * meta: device int64 [G][B2S_PAIR_META] = {length, off_estimate, off_target}; row i of a group
 * starts at off + i*source_stride.  workspace as for PIT (b2s_pair_workspace_bytes).               */
#define B2S_PAIR_META 3
B2S_API int64_t b2s_pair_workspace_bytes(int64_t groups, int64_t max_length, int sources);
B2S_API int b2s_pair_stats_forward(const float* estimate, const float* target, const int64_t* meta,
                           int64_t groups, int64_t max_length, int sources,
                           int64_t estimate_source_stride, int64_t target_source_stride,
                           double* stats, void* workspace, b2s_stream stream);

#define B2S_LOSS_MSE 0       /* mse_loss        regression.py:47   mean_t |e-t|^2                   */
#define B2S_LOSS_LOG_MSE 1   /* log_mse_loss    regression.py:71   log10(mse [+ tau*mean t^2])      */
#define B2S_LOSS_LOG1P_MSE 2 /* log1p_mse_loss  regression.py:299  log10(1 + mse)                   */
#define B2S_LOSS_SDR 3       /* sdr_loss        regression.py:131  -10 log10(|t|^2/(|e-t|^2[+tau|t|^2])) */
#define B2S_LOSS_SI_SDR 4    /* si_sdr_loss     regression.py:178  sdr(e, <e,t>/|t|^2 t)            */
#define B2S_LOSS_SA_SDR 5    /* source_aggregated_sdr_loss regression.py:344 (pit == 0 only; one value
                                per example, aggregated over all its rows)                          */
#define B2S_FLAG_OFFSET_INVARIANT 1 /* si_sdr_loss(offset_invariant=True), regression.py:280-282   */
#define B2S_FLAG_GRAD_STOP 2        /* si_sdr_loss(grad_stop=True),        regression.py:286-287   */
#define B2S_REDUCE_NONE 0
#define B2S_REDUCE_SUM 1
#define B2S_REDUCE_MEAN 2
/* Evaluate one loss from the statistics.  `inner` consecutive groups form one example (the middle
 * axes of a [K, ..., T] estimate); examples = groups / inner.
 *   pit != 0: loss[example] = min over permutations of reduce_{k,inner} l(e_perm[k], t_k) with
 *             `reduction` SUM or MEAN (the loss_fn's own default, source_separation.py:115-118),
 *             perm [examples][K] (first minimum wins).
 *   pit == 0: loss[group*K + k] = l(e_k, t_k) (reduction 'none'; perm unused).
 * tau < 0 disables soft_sdr_max, otherwise tau = 10^(-soft_sdr_max/10) (regression.py:39-44).      */
B2S_API int b2s_pair_loss(const double* stats, const int64_t* meta, int64_t groups, int64_t inner,
                  int sources, int kind, int flags, double tau, int reduction, int pit, float* loss,
                  int32_t* perm, b2s_stream stream);
/* `count` (<= B2S_MAX_LOSS_SET) PIT losses of the same statistics in one launch, plus their batch means
 * (TasNet.loss, tasnet/model.py:154-176: three loss functions per step, each averaged over the batch):
 * loss [count][examples], perm [count][examples][K], mean [count] = mean over examples of loss[c].
 * kinds / reductions: HOST arrays.                                                                 */
#define B2S_MAX_LOSS_SET 8
B2S_API int b2s_pair_loss_set(const double* stats, const int64_t* meta, int64_t groups, int64_t inner,
                      int sources, int count, const int* kinds, const int* reductions, int flags,
                      double tau, float* loss, int32_t* perm, float* mean, b2s_stream stream);
/* b2s_pair_stats_forward followed by b2s_pair_loss_set with inner == 1 (groups = examples: TasNet.loss,
 * tasnet/model.py:154-176) in ONE launch: the CTA that completes an example's statistics evaluates the loss set
 * from them, the CTA that completes the last example folds the batch means.  Same outputs as the two calls
 * (stats is written as well: the backward reads it).  Rows that cannot be read with 16-byte loads and
 * sources > 4 run the two launches internally.                                                       */
B2S_API int b2s_pair_stats_loss_set(const float* estimate, const float* target, const int64_t* meta,
                            int64_t groups, int64_t max_length, int sources,
                            int64_t estimate_source_stride, int64_t target_source_stride, int count,
                            const int* kinds, const int* reductions, int flags, double tau, double* stats,
                            float* loss, int32_t* perm, float* mean, void* workspace, b2s_stream stream);
/* grad_estimate[row i of group g] = grad_scale * grad_loss[. * grad_loss_stride] * dl/de_i for the
 * pairing the forward chose (perm NULL: identity).  grad_loss indexes examples (pit) or rows
 * (pit == 0); stride 0 broadcasts one upstream value (the gradient of a batch mean: scale 1/B).    */
B2S_API int b2s_pair_backward(const float* estimate, const float* target, const int64_t* meta,
                      int64_t groups, int64_t inner, int64_t max_length, int sources,
                      int64_t estimate_source_stride, int64_t target_source_stride,
                      const double* stats, int kind, int flags, double tau, int reduction, int pit,
                      const int32_t* perm, const float* grad_loss, int64_t grad_loss_stride,
                      double grad_scale, float* grad_estimate, b2s_stream stream);

/* compute_pairwise_losses (padertorch/ops/losses/source_separation.py:127-241) from the statistics:
 * matrix[example][i][j] = reduce over the example's `inner` groups of l(e_i, t_j), `reduction` SUM or MEAN
 * (the loss_fn's own default over the leading axes of estimate[i], :230-241).  float [examples][K][K].        */
B2S_API int b2s_pair_loss_matrix(const double* stats, const int64_t* meta, int64_t groups, int64_t inner,
                         int sources, int kind, int flags, double tau, int reduction, float* matrix,
                         b2s_stream stream);
/* grad_estimate = d/d estimate of sum_{example,i,j} grad_matrix[example][i][j] * matrix[example][i][j]
 * (autograd of the matrix above; the reference differentiates K^2 separate loss_fn calls).                    */
B2S_API int b2s_pair_matrix_backward(const float* estimate, const float* target, const int64_t* meta,
                             int64_t groups, int64_t inner, int64_t max_length, int sources,
                             int64_t estimate_source_stride, int64_t target_source_stride,
                             const double* stats, int kind, int flags, double tau, int reduction,
                             const float* grad_matrix, float* grad_estimate, b2s_stream stream);

/* ---------------------------------------------------------------------------------------------
 * Assignment on the device for a batch of K x K cost matrices (float [batch][K][K], K <= B2S_MAX_SOURCES),
 * replacing the host round trip of pit_loss_from_loss_matrix (padertorch/ops/losses/source_separation.py:
 * 285-298: to_numpy + scipy.optimize.linear_sum_assignment / pb_bss greedy).  Exhaustive search, exact.
 *   orientation 0: assignment[k] = row matched to column k, candidates in itertools.permutations order, first
 *                  minimum wins (pit_loss, :112-122);
 *   orientation 1: assignment[i] = column matched to row i (scipy's col_ind), first minimum in lexicographic order;
 *   greedy != 0 (orientation 1): repeatedly the smallest remaining entry (algorithm='greedy', parity unpinned:
 *                  pb_bss is not part of the reference tree).
 * value (may be NULL): float [batch], the assignment's total cost.                                            */
B2S_API int b2s_assign(const float* cost, int64_t batch, int sources, int orientation, int greedy,
               int32_t* assignment, float* value, b2s_stream stream);

/* ---------------------------------------------------------------------------------------------
 * Deep-clustering affinity loss, deep_clustering_loss (padertorch/ops/losses/source_separation.py:
 * 13-31) batched over the per-example loop of DeepClusteringModel.review (padertorch/contrib/tcl/
 * dc.py:76-84):   loss = (|V^T V|_F^2 - 2 |V^T Y|_F^2 + |Y^T Y|_F^2) / N^2,  N = frames*bins.
 * Element (t, c, f) of a block lives at off + t*frame_stride + c*channel_stride + f*bin_stride, so the
 * model's native 't e f' layout (no transpose copy, unlike dc.py:80-81) and the op's [N, E] layout
 * are both accepted.  meta: device int64 [B][B2S_DC_META] = {frames, off_embedding, off_target,
 * off_grad} (off_grad: start of the example's block in grad_embedding of the backward call).
 * gram: [B][C][C] double with C = E + K (row-major, symmetric), kept for the backward.             */
#define B2S_DC_META 4
#define B2S_DC_MAX_CHANNELS 64
B2S_API int64_t b2s_dc_workspace_bytes(int64_t batch, int64_t max_frames, int64_t bins, int channels);
B2S_API int b2s_dc_forward(const float* embedding, const float* target, const int64_t* meta, int64_t batch,
                   int64_t max_frames, int64_t bins, int embedding_dim, int sources,
                   const int64_t* embedding_strides /*host [3]: frame, channel, bin*/,
                   const int64_t* target_strides /*host [3]*/, float* loss, double* gram,
                   void* workspace, b2s_stream stream);
/* b2s_dc_forward plus mean[0] = mean_b loss[b] (dc_loss of DeepClusteringModel.review, tcl/dc.py:83-84) folded by
 * the same launch: the CTA that finishes the last example adds the losses in a fixed order.  batch >= 1.          */
B2S_API int b2s_dc_forward_mean(const float* embedding, const float* target, const int64_t* meta, int64_t batch,
                        int64_t max_frames, int64_t bins, int embedding_dim, int sources,
                        const int64_t* embedding_strides, const int64_t* target_strides, float* loss,
                        float* mean, double* gram, void* workspace, b2s_stream stream);
/* grad_embedding = grad_loss[b] * 4/N^2 (V V^T V - Y Y^T V), written with embedding's strides.      */
B2S_API int b2s_dc_backward(const float* embedding, const float* target, const int64_t* meta, int64_t batch,
                    int64_t max_frames, int64_t bins, int embedding_dim, int sources,
                    const int64_t* embedding_strides, const int64_t* target_strides,
                    const double* gram, const float* grad_loss, float* grad_embedding,
                    b2s_stream stream);
/* The same with the upstream gradient grad_scale * grad_loss[b * grad_loss_stride]: stride 0 broadcasts ONE value
 * (the gradient of a batch mean: scale 1 / batch) -- no expand / divide kernels in front of the backward.         */
B2S_API int b2s_dc_backward_scaled(const float* embedding, const float* target, const int64_t* meta, int64_t batch,
                           int64_t max_frames, int64_t bins, int embedding_dim, int sources,
                           const int64_t* embedding_strides, const int64_t* target_strides,
                           const double* gram, const float* grad_loss, int64_t grad_loss_stride,
                           double grad_scale, float* grad_embedding, b2s_stream stream);

/* ---------------------------------------------------------------------------------------------
 * The fused north-star kernel: STFT -> mask (*) |Y| -> PIT-MSE without materialising the target
 * spectra.  Per example b:  X_k = |STFT(sources[b,k])| is recomputed in registers, |Y| is read from
 * `observation_abs` (the front-end feature already in HBM) or, when that is NULL, recomputed from
 * `mixture`; then the SSE matrix / permutation search of b2s_pit_sse_forward.
 * Equals stft -> abs -> pit_loss(mask*Y_abs[:,None,:], X_abs, axis=-2) of pit/data.py:49-77 and
 * pit/model.py:117-128.  Fast plans (size 1024) only.
 * mixture [B, samples]; sources [B, K, samples]; mask [B, frames, K, F]; observation_abs
 * [B, frames, F]; meta: device int64 [B][2] = {samples_b, frames_b} or NULL (all full length).     */
B2S_API int64_t b2s_stft_pit_workspace_bytes(int64_t batch, int64_t frames, int sources);
B2S_API int b2s_stft_pit_forward(const b2s_stft_plan* plan, const float* mixture,
                         const float* observation_abs, const float* sources, const float* mask,
                         const int64_t* meta, int64_t batch, int64_t samples, int sources_k,
                         int64_t frames, int64_t pad_left, float* loss, int32_t* perm, double* sse,
                         void* workspace, b2s_stream stream);

/* Backward of b2s_stft_pit_forward: grad_mask [B, frames, K, F] = grad_loss[b] * d loss[b] / d mask for the
 * permutation the forward pass chose (perm [B, K] as written by it) -- what autograd derives for
 * pit_loss(mask * Y_abs[:, None, :], X_abs, axis=-2) (pit/model.py:117-128, source_separation.py:112-119):
 *     2 / (frames_b K F) * (mask |Y| - |X_j(i)|) * |Y|,   perm[j(i)] == i,
 * with |X_k| recomputed from the waveforms in registers (no 4MFK read of materialised targets).  Frames beyond
 * an example's length (meta) are not written: pass a zero-filled grad_mask for ragged batches.               */
B2S_API int b2s_stft_pit_backward(const b2s_stft_plan* plan, const float* mixture,
                          const float* observation_abs, const float* sources, const float* mask,
                          const int64_t* meta, int64_t batch, int64_t samples, int sources_k,
                          int64_t frames, int64_t pad_left, const int32_t* perm, const float* grad_loss,
                          float* grad_mask, b2s_stream stream);

/* ---------------------------------------------------------------------------------------------
 * Batched target / feature preparation of the PIT example on the device:
 * pre_batch_transform, padertorch/contrib/examples/source_separation/pit/data.py:49-77, which runs per
 * example in numpy on the data-loader workers:
 *     y_abs [B, M, F] = |Y|,  x_abs [B, M, K, F] = |X| ('k t f -> t k f'),
 *     cos_phase_difference [B, M, K, F] = cos(angle(Y)[:, None, :] - angle(X))   (angle(0) = 0)
 * spec_mixture [B, M, F, 2] and spec_sources [B, K, M, F, 2]: interleaved complex spectra as written by
 * b2s_stft_forward (layout B2S_SPEC_INTERLEAVED) for the rows [B] and [B * K].  16-byte aligned buffers. */
B2S_API int b2s_pit_targets(const float* spec_mixture, const float* spec_sources, int64_t batch, int sources,
                    int64_t frames, int64_t bins, float* y_abs, float* x_abs,
                    float* cos_phase_difference, b2s_stream stream);

/* Masked spectra of the evaluation path (padertorch/contrib/examples/source_separation/pit/evaluate.py:147-152):
 * masked [B, K, frames, F, 2] = mask [B, frames, K, F] * spec_mixture [B, frames, F, 2] -- the rows b2s_istft_forward
 * consumes ('t k f -> k t f' included).                                                                       */
B2S_API int b2s_mask_spectrum(const float* mask, const float* spec_mixture, int64_t batch, int sources,
                      int64_t frames, int64_t bins, float* masked, b2s_stream stream);

/* The same preparation straight from the waveforms, the complex spectra never leaving the registers (fast
 * plans: size 1024 / window_length 1024 / shift % 4 == 0; 1..4 sources):
 * mixture [B, samples], sources [B, K, samples] -> y_abs [B, frames, F], x_abs / cos_phase_difference
 * [B, frames, K, F].  frames / pad_left as for b2s_stft_forward.
 * meta: NULL (all examples full length) or device int64 [B][2] = {samples_b, frames_b} of a padded batch: samples
 * beyond samples_b are never read, rows of frames >= frames_b are written as zeros (what collating the
 * per-example results of pit/data.py:49-77 with zero padding gives).                                          */
B2S_API int b2s_stft_pit_targets(const b2s_stft_plan* plan, const float* mixture, const float* sources,
                         const int64_t* meta, int64_t batch, int64_t samples, int sources_k, int64_t frames,
                         int64_t pad_left, float* y_abs, float* x_abs, float* cos_phase_difference,
                         b2s_stream stream);

/* ---------------------------------------------------------------------------------------------
 * Dense projections of the mask networks on the tcgen05 tensor cores:
 *     c [m, n] = act(a [m, k] . weight [n, k]^T + bias [n])            == torch.nn.functional.linear
 * for Linear(1200, 1200) + ReLU, Linear(1200, F K) + sigmoid and the LSTM input projections of
 * padertorch/contrib/examples/source_separation/pit/model.py:60-73, 96-102, and the 1 x 1 convolutions of
 * padertorch/modules/convnet.py:120-167.  TMA tensor maps -> shared memory -> tcgen05.mma.kind::tf32 with the
 * accumulator in tensor memory -> tcgen05.ld -> bias / activation -> global.
 * a_lo / weight_lo: x - tf32(x) of the operands (b2s_tf32_split, or `c_lo` of the preceding projection): with
 * them the kernel evaluates a_hi w_hi + a_hi w_lo + a_lo w_hi (fp32-faithful: ~2^-21 relative per product,
 * the reference runs these GEMMs in fp32); without them (both NULL) one TF32 product (~1e-3).
 * Row strides in floats, multiples of 4 (TMA: 16-byte pitches); base pointers 16-byte aligned.
 * c_lo (may be NULL): additionally writes c - tf32(c) for a following projection.                           */
#define B2S_ACT_NONE 0
#define B2S_ACT_RELU 1
#define B2S_ACT_SIGMOID 2
B2S_API int b2s_tf32_split(const float* x, int64_t count, float* lo, b2s_stream stream);
B2S_API int b2s_linear_forward(const float* a, const float* a_lo, const float* weight, const float* weight_lo,
                       const float* bias, int64_t m, int64_t n, int64_t k, int64_t a_row_stride,
                       int64_t weight_row_stride, int activation, float* c, float* c_lo, b2s_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SEP_H_ */
