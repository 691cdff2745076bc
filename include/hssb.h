/* hssb.h -- C-ABI of libhssb.so: the B200 (sm_100a) FSST -> BiLSTM hot path.
 *
 * Drop-in boundary for alvgaona/heart-sounds-segmentation.  Every entry point names the reference
 * interface it replaces (file:line relative to the reference repository):
 *
 *   FSST      ssq.fsst(x, fs, window) -> (s, f, t)       hss/transforms/synchrosqueeze.py:48
 *                                                          scripts/visualize_signals.py:14
 *             band mask / abs / z-score+stack              hss/transforms/synchrosqueeze.py:56-111
 *             streaming mean / M2 recurrences              hss/moments/__init__.py:16,35-36
 *   BiLSTM    HeartSoundSegmenter.forward                  hss/model/segmenter.py:70-87
 *   training  nn.LSTM forward/backward under autograd      main.py:67-82 (segmenter.py:80-83)
 *   metrics   multiclass confusion counts                  main.py:36-62 (torchmetrics)
 *             one-vs-rest score histograms (AUROC)         main.py:48,60
 *
 * Conventions
 *   - plain C: pointers, sizes, no torch / C++ types.
 *   - "device" pointers are caller-owned CUDA device memory; the library never frees them and all
 *     work is ordered on the caller's stream (a cudaStream_t passed as void*).  It never
 *     synchronises unless the entry point says "host".
 *   - "host" entry points take host buffers, do the H2D / D2H copies themselves and synchronise
 *     before returning (this is what a binding replacing ssq.fsst would call).
 *   - return 0 on success; <0 = argument error (HSSB_E_*); >0 = cudaError_t.  The message of the
 *     last failure on the calling thread is returned by hssb_last_error().
 *   - there is no CPU fallback: without a usable sm_100 device every compute entry point fails.
 */
#ifndef HSSB_H
#define HSSB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSSB_VERSION 1

enum {
    HSSB_OK = 0,
    HSSB_E_NULL = -1,      /* required pointer is NULL */
    HSSB_E_SHAPE = -2,     /* negative / inconsistent sizes */
    HSSB_E_NWIN = -3,      /* unsupported window length (supported: 4 .. 1024; 128 and 256 run on the radix kernels) */
    HSSB_E_BAND = -4,      /* k_lo / k_hi outside [0, nwin/2] or k_hi < k_lo */
    HSSB_E_MODE = -5,      /* unknown output mode */
    HSSB_E_WORKSPACE = -6, /* workspace too small / misaligned */
    HSSB_E_MODEL = -7,     /* unsupported model geometry */
    HSSB_E_DEVICE = -8,    /* no sm_100 device / wrong architecture */
    HSSB_E_IO = -9         /* file cannot be read / malformed recording file */
};

/* output modes of the FSST wrapper (branch precedence of synchrosqueeze.py:56-65) */
enum { HSSB_MODE_RAW = 0, HSSB_MODE_ABS = 1, HSSB_MODE_STACK = 2 };

typedef struct { float re, im; } hssb_c32; /* complex64, layout of torch.complex64 / numpy.complex64 */

int hssb_version(void);
const char *hssb_last_error(void);

/* ------------------------------------------------------------------------------------------
 * FSST: three kernels + one fused convenience call.
 * Shapes: x [B,N] f32; g, dg [nwin] f32 (window and its dtwin derivative, computed by the host
 * in float64 and rounded); nfft = nwin; K = nwin/2+1; Kt = k_hi-k_lo+1.
 * ------------------------------------------------------------------------------------------ */

/* K1  replaces the STFT half of ssq.fsst (synchrosqueeze.py:48): hop-1 STFT of every window with
 * g and g'.  Sg, Sdg: [B,K,N] complex64 (frequency-major, time contiguous). */
int hssb_fsst_stft(const float *x, int64_t B, int64_t N, const float *g, const float *dg, int nwin,
                   hssb_c32 *Sg, hssb_c32 *Sdg, void *stream);

/* number of float64 words of the per-tile moment partials written by hssb_fsst_reassign */
size_t hssb_fsst_stats_words(int64_t B, int64_t N);

/* K2  replaces the IF-estimate + reassignSpectrum half of ssq.fsst (synchrosqueeze.py:48) fused
 * with the band mask (synchrosqueeze.py:107-111): T [B,Kt,N] complex64 holds rows k_lo..k_hi of
 * the synchrosqueezed transform.  stats (nullable) receives per-tile (n, mean, M2) partials of Re
 * and Im -- the hss/moments/__init__.py recurrences, parallelised -- for hssb_fsst_finish. */
int hssb_fsst_reassign(const hssb_c32 *Sg, const hssb_c32 *Sdg, int64_t B, int64_t N, int nwin,
                       float fs, int k_lo, int k_hi, hssb_c32 *T, double *stats, void *stream);

/* K3  replaces synchrosqueeze.py:59-60 (mode ABS: out f32 [B,N,Kt]) and synchrosqueeze.py:78-89
 * (mode STACK: out f32 [B,N,2*Kt] = [(Re-mean)/std , (Im-mean)/std], unbiased std). */
int hssb_fsst_finish(const hssb_c32 *T, const double *stats, int64_t B, int64_t N, int Kt, int mode,
                     float *out, void *stream);

/* K1+K2(+K3) on device buffers.  out: RAW -> hssb_c32 [B,Kt,N]; ABS -> f32 [B,N,Kt];
 * STACK -> f32 [B,N,2*Kt].  workspace: hssb_fsst_workspace_bytes() bytes, 256-byte aligned. */
size_t hssb_fsst_workspace_bytes(int64_t B, int64_t N, int nwin, int k_lo, int k_hi, int mode);
int hssb_fsst_forward(const float *x, int64_t B, int64_t N, const float *g, const float *dg, int nwin,
                      float fs, int k_lo, int k_hi, int mode, void *out, void *workspace,
                      size_t workspace_bytes, void *stream);

/* HOST entry point: the call a binding replacing ssq.fsst makes.  x, window, dwindow and out are
 * host buffers; copies + kernels + synchronisation happen inside.  Same `out` layouts as above. */
int hssb_fsst_host(const float *x, int64_t B, int64_t N, double fs, const double *window,
                   const double *dwindow, int nwin, int k_lo, int k_hi, int mode, void *out);

/* ------------------------------------------------------------------------------------------
 * BiLSTM segmenter: replaces HeartSoundSegmenter.forward (segmenter.py:70-87), eval mode.
 * ------------------------------------------------------------------------------------------ */
typedef struct hssb_model hssb_model; /* opaque: packed weights resident in HBM */

/* Parameters in torch layout, HOST or DEVICE pointers are both accepted (cudaMemcpyDefault):
 * per layer l in {0,1} and direction d in {0 fwd, 1 reverse}: w_ih [4H,Kin_l], w_hh [4H,H],
 * b_ih [4H], b_hh [4H] (gate order i,f,g,o; segmenter.py:43-58); lin_w [4,2H], lin_b [4]
 * (segmenter.py:61-67).  input_size = Kin_0, Kin_1 = 2H. */
typedef struct {
    int input_size;
    int hidden_size;
    const float *w_ih[2][2];
    const float *w_hh[2][2];
    const float *b_ih[2][2];
    const float *b_hh[2][2];
    const float *lin_w;
    const float *lin_b;
} hssb_model_params;

int hssb_model_create(const hssb_model_params *params, hssb_model **out, void *stream);
void hssb_model_destroy(hssb_model *m);

/* The first kernel of hssb_model_forward ahead of time: x[B][T][F] -> the fp16 hi / lo operand planes of the layer-1
 * recurrence (+ the input-range flag) in a caller-owned, 256-byte aligned device buffer of hssb_model_split_bytes bytes.  A
 * pipelined caller runs it behind the next batch's FSST on a side stream; hssb_model_forward_split (same arguments and results
 * as hssb_model_forward with impl 0; x is still needed: inputs beyond the fp16-split range are re-read from it) then starts
 * with the recurrence.  The buffer is consumed (and may be rewritten) by that forward.  No-op / ignored for models on the
 * generic kernels.  Part of `self.lstm_1(x, ...)` of segmenter.py:80. */
size_t hssb_model_split_bytes(const hssb_model *m, int64_t B, int64_t T);
int hssb_model_split_input(const hssb_model *m, const float *x, int64_t B, int64_t T, void *planes, size_t planes_bytes,
                           void *stream);
int hssb_model_forward_split(const hssb_model *m, const float *x, void *planes, int64_t B, int64_t T, const float *h0,
                             const float *c0, float *logp, int32_t *labels, void *workspace, size_t workspace_bytes,
                             void *stream);

/* Pipelining hook (no reference counterpart: the reference transforms and segments strictly one after the other,
 * heart_sounds.py:199-201 then main.py:64-65).  Enqueues on side_stream a wait until the layer-1 recurrence of the forward most
 * recently enqueued with (m, workspace, B, T) holds its SMs (benign 20 ms time-out).  Work queued on side_stream behind it --
 * the NEXT batch's FSST -- then runs on the ~50 SMs the latency-bound recurrences leave idle instead of delaying the placement
 * of their clusters.  Call right after hssb_model_forward; a no-op for models on the generic kernels. */
int hssb_model_side_gate(const hssb_model *m, int64_t B, int64_t T, void *workspace, void *side_stream);

/* Re-packs changed parameters into an existing model (same input_size / hidden_size, same device): what an optimiser step
 * (reference main.py:81-82) does to the nn.LSTM weights the next forward reads.  Synchronises `stream` once.  Not to be
 * called while a forward of the same model is being enqueued from another thread. */
int hssb_model_update(hssb_model *m, const hssb_model_params *params, void *stream);

/* 1 when the model runs on the tcgen05 kernels (hidden_size 240, input_size <= 64, |w| <= 2^15), 0 on the generic kernels. */
int hssb_model_uses_tensor_cores(const hssb_model *m);

size_t hssb_model_workspace_bytes(const hssb_model *m, int64_t B, int64_t T);

/* x [B,T,input_size] f32 device; h0, c0 [2,B,H] f32 device (segmenter.py:38-41);
 * logp [B,T,4] f32 device (nullable); labels [B,T] int32 device (nullable) = argmax over classes.
 * impl: 0 = default (tcgen05 kernels when hidden_size == 240, else the generic SIMT CUDA kernels),
 *       1 = force the SIMT fp32 kernels (on-device validation path). */
int hssb_model_forward(const hssb_model *m, const float *x, int64_t B, int64_t T, const float *h0,
                       const float *c0, float *logp, int32_t *labels, void *workspace,
                       size_t workspace_bytes, int impl, void *stream);

/* The same two calls under the names SURVEY 8b sketches for this boundary (`hssb_weights` = the packed model handle):
 * F must equal the model's input_size; impl = 0. */
typedef hssb_model hssb_weights;
size_t hssb_lstm_workspace_bytes(const hssb_weights *w, int64_t B, int64_t T);
int hssb_lstm_forward(const hssb_weights *w, const float *x, int64_t B, int64_t T, int F, const float *h0, const float *c0,
                      float *logp, int32_t *labels, void *workspace, size_t workspace_bytes, void *stream);

/* Diagnostic: layer-1 input projection only (kernel-level parity tests of K4).  impl 0 = tcgen05
 * kernel, 1 = SIMT kernel.  xproj [2,B*T,4H] f32 device, torch gate order.  workspace >=
 * 2*T*(B+1)*4H*4 + 2*B*T*256 bytes. */
int hssb_debug_inproj(const hssb_model *m, const float *x, int64_t B, int64_t T, int impl, float *xproj,
                      void *workspace, size_t workspace_bytes, void *stream);

/* Diagnostic: clock64 stamps of the recurrence kernel's roles (cluster 0, CTA rank 0) for the first
 * `steps` steps of every following recurrence launch.  buf: steps*4*16 uint64 on the device; NULL
 * disables.  Read by scripts/trace_recurrent.py. */
int hssb_debug_trace(unsigned long long *buf, int steps);
/* Diagnostic: byte offset of the flag area of the overlapped layer-2 projection inside the model workspace
 * (next_item[8], timeout flag, resident counter | chunk_done[8192] | tile_done[8192], all 32-bit). */
long long hssb_debug_sync_offset(long long B, long long T);
/* Diagnostic: co-resident 8-CTA recurrence clusters (geometry 32x3) on the current device; <0 on error. */
int hssb_debug_max_clusters(void);

/* ------------------------------------------------------------------------------------------
 * Training recurrences of one bidirectional LSTM layer (what autograd does for nn.LSTM in the reference's
 * training step, main.py:67-82 over segmenter.py:80-83).  fp32.  The plain GEMMs either side (x W_ih^T + b,
 * dG^T x, dG^T h_prev, dG W_ih) are library GEMMs issued by the caller (hss/model/_train.py).
 * Layouts: gates [2][B*T][4H] (row b*T+t, gate order i,f,g,o), cells [2][B*T][H], out / d_out [B][T][2H],
 * states [2][B][H]; dir 0 = forward, 1 = reverse.
 * ------------------------------------------------------------------------------------------ */

/* gates in: x W_ih^T + b_ih + b_hh per direction; out: the activated gates (kept for the backward).
 * w_hhT_*: W_hh transposed, [H][4H].  Writes out (raw h_t, no ReLU), cells (c_t), hn, cn. */
int hssb_lstm_train_forward(float *gates, const float *w_hhT_fwd, const float *w_hhT_rev, const float *h0, const float *c0,
                            int64_t B, int64_t T, int H, float *out, float *cells, float *hn, float *cn, void *stream);

/* The same forward for the reference geometry (hssb_model_uses_tensor_cores): input projection and recurrence of layer
 * `layer` (0: x[B][T][input_size], 1: x[B][T][2H] = dropout(relu(out of layer 0))) on the tcgen05 kernels of the inference
 * path (split-fp16 operands, fp32 accumulation), packed weights taken from `m` (hssb_model_update after each optimiser
 * step).  Writes the ACTIVATED gates, out (raw h_t), cells, hn, cn in the layouts above.  workspace: hssb_model_workspace_bytes. */
int hssb_lstm_train_forward_tc(const hssb_model *m, int layer, const float *x, int64_t B, int64_t T, const float *h0,
                               const float *c0, float *gates, float *out, float *cells, float *hn, float *cn,
                               void *workspace, size_t workspace_bytes, void *stream);

/* gates in: activated gates from the forward; out: dG = dL/d(gate pre-activations).  w_hh_*: [4H][H] (torch layout).
 * d_hn, d_cn nullable (zero).  Writes dh0, dc0 (gradient w.r.t. the initial state). */
int hssb_lstm_train_backward(float *gates, const float *cells, const float *w_hh_fwd, const float *w_hh_rev, const float *c0,
                             const float *d_out, const float *d_hn, const float *d_cn, int64_t B, int64_t T, int H,
                             float *dh0, float *dc0, void *stream);

/* The same back-propagation through time for hidden_size 240 on the tcgen05 kernels (W_hh^T slices resident in TMEM, the
 * gate gradients as split-fp16 operands scaled by a power of two taken from max|d_out|, |d_hn|; reduce-scatter of the
 * partial dL/dh between the 8 CTAs of a cluster by bulk copies).  Same layouts; the gate gradients go to dG, which may be
 * the gates buffer itself (in place, like the fp32 entry point) or a separate one (the saved activations stay intact),
 * and / or to the pair (dG_hi, dG_lo) = hssb_split_tf32 of the same values, laid out [B*T][2][4H] (both directions of a row
 * side by side: dG^T x and dG W_ih are then one GEMM each for the two directions) -- dG and the pair may each be NULL;
 * db (nullable) [2][4H]: the sum of dG over batch and time = the gradient of b_ih and of b_hh;
 * weights must be in the fp16-split range (hssb_model_uses_tensor_cores of a model holding them).  workspace: 256-byte
 * aligned device scratch of hssb_lstm_train_backward_tc_workspace_bytes() bytes. */
size_t hssb_lstm_train_backward_tc_workspace_bytes(void);
int hssb_lstm_train_backward_tc(const float *gates, float *dG, float *dG_hi, float *dG_lo, float *db, const float *cells, const float *w_hh_fwd, const float *w_hh_rev, const float *c0,
                                const float *d_out, const float *d_hn, const float *d_cn, int64_t B, int64_t T,
                                float *dh0, float *dc0, void *workspace, size_t workspace_bytes, void *stream);

/* Fused head + loss of the training step: linear(2H -> 4) + log_softmax (segmenter.py:86-87) + nn.CrossEntropyLoss on the
 * permuted output (main.py:69-70).  act [M,K] f32 (K = 2H <= 512, after ReLU / dropout), w [4,K], b [4], target [M] int64.
 * forward : logp [M,4]; *loss_sum (float64 device, caller zeroes) += sum over rows of (lse(logp) - logp[target]).
 * backward: with scale (x *upstream when upstream != NULL: a device scalar, so autograd's dL/d(loss) needs no host read)
 *           = upstream gradient / number of rows (mean reduction): d_act [M,K] (written), d_w [4,K] and d_b [4]
 *           (accumulated: caller zeroes). */
int hssb_ce_head_forward(const float *act, int64_t M, int K, const float *w, const float *b, const int64_t *target,
                         float *logp, double *loss_sum, void *stream);
int hssb_ce_head_backward(const float *act, const float *logp, int64_t M, int K, const float *w, const int64_t *target,
                          float scale, const float *upstream, float *d_act, float *d_w, float *d_b, void *stream);

/* a[n] = hi[n] + lo[n], hi exactly representable in TF32 (round to nearest, ties away; inf / NaN pass through).  Operand
 * preparation for running the fp32 GEMMs of back-propagation (autograd's mm kernels under nn.LSTM, reference main.py:72) as
 * three TF32 tensor-core GEMMs hi.hi + hi.lo + lo.hi.  Pointers 16-byte aligned. */
int hssb_split_tf32(const float *a, int64_t n, float *hi, float *lo, void *stream);

/* Gradient clipping by global norm (pl.Trainer(gradient_clip_val=1), main.py:226; torch.nn.utils.clip_grad_norm_ semantics:
 * coef = min(1, max_norm / (norm + 1e-6)); max_norm <= 0 disables) fused with one torch.optim.Adam step (main.py:130: default
 * betas / eps, no weight decay) over all n_tensors (<= 32) parameter tensors: two launches.  params / grads / exp_avg /
 * exp_avg_sq: HOST arrays of device pointers, numel: host array; lr = the step's learning rate (the LambdaLR 0.9^epoch schedule of
 * main.py:133-134 is applied by the caller); step counts from 1; norm2_scratch: 8 bytes of device scratch; grad_norm_out
 * (nullable, device): the un-clipped global norm. */
int hssb_clip_adam_step(int n_tensors, float *const *params, const float *const *grads, float *const *exp_avg,
                        float *const *exp_avg_sq, const int64_t *numel, float lr, float beta1, float beta2, float eps,
                        int64_t step, float max_norm, double *norm2_scratch, float *grad_norm_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * Metric counters: replaces the torchmetrics confusion statistics of main.py:36-62.
 * cm16 [4,4] int64 device, cm[target][pred] += 1 (accumulates; caller zeroes).
 * ------------------------------------------------------------------------------------------ */
int hssb_confusion(const int32_t *pred, const int64_t *target, int64_t n, int64_t *cm16, void *stream);

/* The complete metric state of one evaluation step (main.py:36-62 counters + the loss logged at main.py:69-70,91-92,
 * 112-117) in <= 32 scalars: state18 [18] float64 device = 16 confusion counts cm[target][pred] (exact integers), the summed
 * nn.CrossEntropyLoss terms of the permuted log-probabilities (lse(logp) - logp[target]) and the number of elements.
 * Accumulates (caller zeroes); one all-reduce(SUM) of the 18 doubles merges ranks; mean loss = state[16] / state[17].
 * logp [n,4] f32 device, 16-byte aligned; pred [n] int32 device or NULL (labels = first maximum of logp); rows with a target
 * outside 0..3 are skipped. */
int hssb_metrics_update(const float *logp, const int32_t *pred, const int64_t *target, int64_t n, double *state18, void *stream);

/* One-vs-rest score histograms for the binned AUROC of main.py:48,60 (torchmetrics multiclass AUROC on the class
 * probabilities).  logp [n,4] f32 device (log-probabilities, 16-byte aligned), target [n] int64 device (rows with a
 * target outside 0..3 are skipped); hist [4][2][nbins] int64 device,
 *   hist[c][target == c][min(nbins-1, floor(exp(logp[c]) * nbins))] += 1   (accumulates; caller zeroes; 2 <= nbins <= 4096).
 * Counters add across shards / ranks, so the job-wide AUROC comes from the all-reduced histogram. */
int hssb_auroc_hist(const float *logp, const int64_t *target, int64_t n, int nbins, int64_t *hist, void *stream);

/* ------------------------------------------------------------------------------------------
 * Recording ingest: replaces the per-file pandas.read_csv of hss/datasets/heart_sounds.py:193-197 (_load_file) for a
 * batch of files.  On-disk format: two-column CSV, first row a header (skipped), column 0 the PCG signal (decimal float),
 * column 1 the state label (integer 1..4).  HOST entry points (no CUDA): files are parsed by `threads` host threads
 * (<= 0: all cores) into caller-owned buffers -- pinned staging memory, from which the caller issues the async H2D copies.
 *   hssb_csv_scan : rows_out[i] = data rows of paths[i]
 *   hssb_csv_parse: file i -> signal / labels [offsets[i], offsets[i+1])  (offsets = prefix sums of the scanned rows;
 *                   decimal -> float64 correctly rounded, narrowed to float32 as the reference's torch.tensor(..., dtype) does)
 * ------------------------------------------------------------------------------------------ */
int hssb_csv_scan(const char *const *paths, int n_files, int threads, int64_t *rows_out);
int hssb_csv_parse(const char *const *paths, int n_files, int threads, const int64_t *offsets, float *signal, int64_t *labels);

/* ------------------------------------------------------------------------------------------
 * Per-launch timing for bench.py: when enabled, every kernel launch is bracketed by CUDA events on
 * the launching stream.  hssb_prof_read() synchronises on them, writes one line per kernel
 * ("<name> <launches> <total_ms>\n") into buf, clears the records and returns the text length.
 * ------------------------------------------------------------------------------------------ */
int hssb_prof_enable(int on);
int hssb_prof_read(char *buf, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* HSSB_H */
