// Packed, HBM-resident weights of the BiLSTM segmenter and the internal kernel entry points.
#pragma once
#include "hssb_common.cuh"
#include <cuda_fp16.h>
#include <mutex>
#include <atomic>

struct hssb_model {
    int F;          // input features (44)
    int H;          // hidden units (240)
    int device;
    // ---- SIMT fp32 validation path: transposed weights, folded biases -------------------------
    float *w_ihT[2][2];   // [Kin][4H]   (Kin = F for layer 0, 2H for layer 1)
    float *w_hhT[2][2];   // [H][4H]
    float *bias[2][2];    // [4H] = b_ih + b_hh
    float *lin_w;         // [4][2H]
    float *lin_b;         // [4]
    // ---- tcgen05 path (H == 240 only): split-fp16 operands in "cluster gate order" -------------
    // gate row index g' = rank*128 + gate*32 + u  (rank 0..7, gate 0..3 = i,f,g,o, u 0..31; u >= 30 is padding)
    bool tc_ready;
    __half *tc_wih[2];        // layer l: [2 planes hi/lo][2 dirs][1024 g'][KinP]   KinP = 64 (l=0) / 512 (l=1)
    __half *tc_whh[2];        // layer l: [2 planes][2 dirs][1024 g'][256]  (k index = rank*32+u order, padded)
    __half *tc_whh_frag[2];   // the same W_hh planes with the TMEM lanes in fragment order (see pack_whh_kernel)
    float *tc_bias[2];        // layer l: [2 dirs][960]
    float *tc_lin_w;          // [4][512] linear weights in slot layout
    __half *tc_wih0_frag;     // layer-1 W_ih slices [dir][rank][plane][128 rows in fragment order][64] (fused projection)
    float *tc_bias0_frag;     // their folded biases [dir][rank][128]
    void *all;                // single allocation backing everything above
    size_t all_bytes;
    float *stage;             // staging for the raw torch tensors (they may be host pointers); used by create and update
    void *tc_base;            // start of the tcgen05 operands inside `all`
    // overlapped layer-2 projection (tc_forward): the layer-2 recurrence runs on a high-priority internal stream while the tail of
    // the projection GEMM keeps the SMs it leaves idle busy; a second internal stream hosts the middle-out part of the GEMM that
    // runs under the layer-1 recurrence.  Everything is joined back into the caller's stream before tc_forward's last kernel.
    cudaStream_t hi_stream, side_stream;
    cudaEvent_t ev[6];
    int sm_count;
    std::mutex *enqueue_mu;   // the internal streams / events belong to the model: forwards on one model are ENQUEUED one at a time
    std::atomic<int> *pipelined;   // set by hssb_model_side_gate: the caller overlaps other work with the forwards, so the middle-out
                              // projection launch is held back until its first tile exists (its CTAs would otherwise poll on idle SMs)
};

namespace hssb {

// SIMT path (lstm_simt.cu)
int simt_inproj(const float *A, int64_t M, int K, const float *Wt, const float *bias, int N, float *C, cudaStream_t st);
int simt_recurrent(const float *xproj /*[2][B*T][4H]*/, const float *const w_hhT[2], const float *h0, const float *c0,
                   int64_t B, int64_t T, int H, float *out /*[B,T,2H] relu'd*/, float *hn, float *cn, cudaStream_t st);
int head_forward(const float *act /*[M,2H] already relu'd*/, int64_t M, int H2, const float *lin_w, const float *lin_b,
                 float *logp, int32_t *labels, cudaStream_t st, const int *poison = nullptr, int64_t T = 0, int64_t Tp = 0);

// tcgen05 path (lstm_tc.cu)
size_t tc_pack_bytes(int F, int H);
int tc_pack(hssb_model *m, const hssb_model_params *p, void *dst, cudaStream_t st);
size_t tc_workspace_bytes(const hssb_model *m, int64_t B, int64_t T);
int tc_forward(const hssb_model *m, const float *x, int64_t B, int64_t T, const float *h0, const float *c0,
               float *logp, int32_t *labels, void *ws, size_t ws_bytes, cudaStream_t st, void *presplit = nullptr);
size_t tc_split_bytes(const hssb_model *m, int64_t B, int64_t T);
int tc_split_input(const hssb_model *m, const float *x, int64_t B, int64_t T, void *planes, size_t planes_bytes, cudaStream_t st);
int tc_side_gate(const hssb_model *m, int64_t B, int64_t T, void *ws, cudaStream_t side);
// training forward of one layer on the tcgen05 kernels: activated gates / cell states / raw h kept for back-propagation
int tc_train_forward(const hssb_model *m, int layer, const float *x, int64_t B, int64_t T, const float *h0, const float *c0,
                     float *gates, float *out, float *cells, float *hn, float *cn, void *ws, size_t ws_bytes, cudaStream_t st);

// training recurrences with cluster-resident weights (lstm_train_cluster.cu); the generic ones are in lstm_train.cu
bool train_cluster_supported(int H);
int train_fwd_cluster_launch(float *gates, const float *w0T, const float *w1T, const float *h0, const float *c0, int64_t B, int64_t T,
                             float *out, float *cells, float *hn, float *cn, cudaStream_t st);
int train_bwd_cluster_launch(float *gates, const float *cells, const float *w0, const float *w1, const float *c0, const float *d_out,
                             const float *d_hn, const float *d_cn, int64_t B, int64_t T, float *dh0, float *dc0, bool all_gather,
                             cudaStream_t st);

// back-propagation through time on the tensor cores (lstm_bptt_tc.cu): hidden_size 240, weights in the fp16-split range
size_t bptt_tc_workspace_bytes();
int bptt_tc_prepare();
int bptt_tc_backward(const float *gates, float *dg, float *dg_hi, float *dg_lo, float *db, const float *cells, const float *w_fwd, const float *w_rev, const float *c0, const float *d_out,
                     const float *d_hn, const float *d_cn, int64_t B, int64_t T, float *dh0, float *dc0, void *ws, size_t ws_bytes,
                     cudaStream_t st);

}  // namespace hssb
