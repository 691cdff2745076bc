// hssb_model: packing of the torch-layout parameters into HBM-resident operands, workspace
// planning and the forward dispatch.  Replaces HeartSoundSegmenter.forward (reference
// hss/model/segmenter.py:70-87), eval mode.
#include "model.cuh"
#include <vector>
#include <cstring>

using namespace hssb;

namespace {

__global__ void transpose_kernel(const float *__restrict__ src, int rows, int cols, float *__restrict__ dst)
{
    // dst[c][r] = src[r][c]
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long long)rows * cols) {
        const int r = (int)(i / cols), c = (int)(i % cols);
        dst[(size_t)c * rows + r] = src[i];
    }
}

__global__ void add_kernel(const float *__restrict__ a, const float *__restrict__ b, int n, float *__restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = a[i] + b[i];
}

struct SimtWs { size_t xproj, out1, out2, hn, cn, total; };
SimtWs simt_ws_layout(const hssb_model *m, int64_t B, int64_t T)
{
    const size_t H = m->H, G = 4 * H;
    SimtWs w{};
    size_t off = 0;
    w.xproj = off; off += align_up(sizeof(float) * 2 * (size_t)B * T * G, 256);
    w.out1 = off;  off += align_up(sizeof(float) * (size_t)B * T * 2 * H, 256);
    w.out2 = off;  off += align_up(sizeof(float) * (size_t)B * T * 2 * H, 256);
    w.hn = off;    off += align_up(sizeof(float) * 2 * (size_t)B * H, 256);
    w.cn = off;    off += align_up(sizeof(float) * 2 * (size_t)B * H, 256);
    w.total = off;
    return w;
}

int simt_forward(const hssb_model *m, const float *x, int64_t B, int64_t T, const float *h0, const float *c0,
                 float *logp, int32_t *labels, void *ws, size_t ws_bytes, cudaStream_t st)
{
    const SimtWs w = simt_ws_layout(m, B, T);
    if (!ws || ws_bytes < w.total) return fail(HSSB_E_WORKSPACE, "model workspace %zu < %zu", ws_bytes, w.total);
    char *base = static_cast<char *>(ws);
    float *xproj = reinterpret_cast<float *>(base + w.xproj);
    float *out1 = reinterpret_cast<float *>(base + w.out1), *out2 = reinterpret_cast<float *>(base + w.out2);
    float *hn = reinterpret_cast<float *>(base + w.hn), *cn = reinterpret_cast<float *>(base + w.cn);
    const int H = m->H, G = 4 * H;
    const int64_t M = B * T;
    // layer 1
    for (int d = 0; d < 2; ++d)
        if (int rc = simt_inproj(x, M, m->F, m->w_ihT[0][d], m->bias[0][d], G, xproj + (size_t)d * M * G, st)) return rc;
    if (int rc = simt_recurrent(xproj, m->w_hhT[0], h0, c0, B, T, H, out1, hn, cn, st)) return rc;
    // layer 2 (initial state = layer-1 final state, segmenter.py:83)
    for (int d = 0; d < 2; ++d)
        if (int rc = simt_inproj(out1, M, 2 * H, m->w_ihT[1][d], m->bias[1][d], G, xproj + (size_t)d * M * G, st)) return rc;
    if (int rc = simt_recurrent(xproj, m->w_hhT[1], hn, cn, B, T, H, out2, hn, cn, st)) return rc;
    return head_forward(out2, M, 2 * H, m->lin_w, m->lin_b, logp, labels, st);
}

int check_params(const hssb_model_params *p, const char *who)
{
    for (int l = 0; l < 2; ++l)
        for (int d = 0; d < 2; ++d)
            if (!p->w_ih[l][d] || !p->w_hh[l][d] || !p->b_ih[l][d] || !p->b_hh[l][d])
                return fail(HSSB_E_NULL, "%s: null parameter (layer %d dir %d)", who, l, d);
    if (!p->lin_w || !p->lin_b) return fail(HSSB_E_NULL, "%s: null linear parameter", who);
    return 0;
}

// torch-layout parameters -> the operands of both kernel families, in the buffers hssb_model_create laid out.  Stream-ordered
// except for one synchronisation at the end (the staging buffer is reused per tensor; tc_pack reads the weight range back).
int pack_parameters(hssb_model *m, const hssb_model_params *p, cudaStream_t st)
{
    const int F = m->F, H = m->H;
    const size_t G = 4 * (size_t)H;
    const int kin[2] = {F, 2 * H};
    cudaError_t e;
    for (int l = 0; l < 2; ++l)
        for (int d = 0; d < 2; ++d) {
            float *s_wih = m->stage, *s_whh = s_wih + kin[l] * G, *s_bi = s_whh + H * G, *s_bh = s_bi + G;
            if ((e = cudaMemcpyAsync(s_wih, p->w_ih[l][d], sizeof(float) * kin[l] * G, cudaMemcpyDefault, st)) != cudaSuccess ||
                (e = cudaMemcpyAsync(s_whh, p->w_hh[l][d], sizeof(float) * H * G, cudaMemcpyDefault, st)) != cudaSuccess ||
                (e = cudaMemcpyAsync(s_bi, p->b_ih[l][d], sizeof(float) * G, cudaMemcpyDefault, st)) != cudaSuccess ||
                (e = cudaMemcpyAsync(s_bh, p->b_hh[l][d], sizeof(float) * G, cudaMemcpyDefault, st)) != cudaSuccess)
                return cuda_fail(e, "cudaMemcpyAsync(model parameter)");
            const long long n_ih = (long long)kin[l] * G, n_hh = (long long)H * G;
            transpose_kernel<<<(unsigned)((n_ih + 255) / 256), 256, 0, st>>>(s_wih, (int)G, kin[l], m->w_ihT[l][d]);
            transpose_kernel<<<(unsigned)((n_hh + 255) / 256), 256, 0, st>>>(s_whh, (int)G, H, m->w_hhT[l][d]);
            add_kernel<<<(unsigned)((G + 255) / 256), 256, 0, st>>>(s_bi, s_bh, (int)G, m->bias[l][d]);
            if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "model pack kernels");
        }
    if ((e = cudaMemcpyAsync(m->lin_w, p->lin_w, sizeof(float) * 4 * 2 * H, cudaMemcpyDefault, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(m->lin_b, p->lin_b, sizeof(float) * 4, cudaMemcpyDefault, st)) != cudaSuccess)
        return cuda_fail(e, "cudaMemcpyAsync(linear)");
    if (tc_pack_bytes(F, H) > 0)
        if (int rc = tc_pack(m, p, m->tc_base, st)) return rc;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize(model pack)");
    return 0;
}

}  // namespace

extern "C" int hssb_model_create(const hssb_model_params *p, hssb_model **out, void *stream)
{
    if (!p || !out) return fail(HSSB_E_NULL, "hssb_model_create: null pointer");
    *out = nullptr;
    const int F = p->input_size, H = p->hidden_size;
    if (F < 1 || H < 1 || F > 4096 || H > 2048) return fail(HSSB_E_MODEL, "input_size=%d hidden_size=%d unsupported", F, H);
    if (int rc = check_params(p, "hssb_model_create")) return rc;
    if (int rc = require_sm100()) return rc;
    cudaStream_t st = as_stream(stream);

    const size_t G = 4 * (size_t)H;
    const int kin[2] = {F, 2 * H};
    // allocation plan: [SIMT operands][tcgen05 operands][staging for the raw torch tensors]
    size_t off = 0;
    size_t o_wih[2][2], o_whh[2][2], o_bias[2][2];
    for (int l = 0; l < 2; ++l)
        for (int d = 0; d < 2; ++d) {
            o_wih[l][d] = off;  off += align_up(sizeof(float) * kin[l] * G, 256);
            o_whh[l][d] = off;  off += align_up(sizeof(float) * H * G, 256);
            o_bias[l][d] = off; off += align_up(sizeof(float) * G, 256);
        }
    const size_t o_linw = off; off += align_up(sizeof(float) * 4 * 2 * H, 256);
    const size_t o_linb = off; off += align_up(sizeof(float) * 4, 256);
    const size_t o_tc = off;   off += align_up(tc_pack_bytes(F, H), 256);
    const size_t o_stage = off;
    size_t stage_bytes = 0;
    for (int l = 0; l < 2; ++l) stage_bytes = std::max(stage_bytes, sizeof(float) * (kin[l] * G + H * G + 2 * G));
    off += align_up(stage_bytes + 64, 256);         // + the two weight-range words tc_pack keeps behind its staged tensor group

    hssb_model *m = new hssb_model();
    std::memset(m, 0, sizeof(*m));
    m->F = F; m->H = H;
    cudaGetDevice(&m->device);
    cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, m->device);
    m->enqueue_mu = new std::mutex();
    m->pipelined = new std::atomic<int>(0);
    {
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        cudaError_t es = cudaStreamCreateWithPriority(&m->hi_stream, cudaStreamNonBlocking, greatest);
        if (es == cudaSuccess) es = cudaStreamCreateWithPriority(&m->side_stream, cudaStreamNonBlocking, least);
        for (int i = 0; i < 6 && es == cudaSuccess; ++i) es = cudaEventCreateWithFlags(&m->ev[i], cudaEventDisableTiming);
        if (es != cudaSuccess) { hssb_model_destroy(m); return cuda_fail(es, "hssb_model_create: internal streams"); }
    }
    cudaError_t e = cudaMalloc(&m->all, off);
    if (e != cudaSuccess) { m->all = nullptr; hssb_model_destroy(m); return cuda_fail(e, "cudaMalloc(model)"); }
    m->all_bytes = off;
    char *base = static_cast<char *>(m->all);
    m->stage = reinterpret_cast<float *>(base + o_stage);
    m->tc_base = base + o_tc;
    for (int l = 0; l < 2; ++l)
        for (int d = 0; d < 2; ++d) {
            m->w_ihT[l][d] = reinterpret_cast<float *>(base + o_wih[l][d]);
            m->w_hhT[l][d] = reinterpret_cast<float *>(base + o_whh[l][d]);
            m->bias[l][d] = reinterpret_cast<float *>(base + o_bias[l][d]);
        }
    m->lin_w = reinterpret_cast<float *>(base + o_linw);
    m->lin_b = reinterpret_cast<float *>(base + o_linb);
    if (int rc = pack_parameters(m, p, st)) { hssb_model_destroy(m); return rc; }
    *out = m;
    return 0;
}

extern "C" int hssb_model_update(hssb_model *m, const hssb_model_params *p, void *stream)
{
    if (!m || !p) return fail(HSSB_E_NULL, "hssb_model_update: null pointer");
    if (p->input_size != m->F || p->hidden_size != m->H)
        return fail(HSSB_E_MODEL, "hssb_model_update: the model was created for input_size=%d hidden_size=%d", m->F, m->H);
    if (int rc = check_params(p, "hssb_model_update")) return rc;
    int dev = -1;
    HSSB_CUDA_OK(cudaGetDevice(&dev));
    if (dev != m->device) return fail(HSSB_E_DEVICE, "hssb_model_update: model lives on device %d, current device is %d", m->device, dev);
    std::lock_guard<std::mutex> enqueue_lock(*m->enqueue_mu);
    return pack_parameters(m, p, as_stream(stream));
}

extern "C" void hssb_model_destroy(hssb_model *m)
{
    if (!m) return;
    if (m->all) cudaFree(m->all);
    if (m->hi_stream) cudaStreamDestroy(m->hi_stream);
    if (m->side_stream) cudaStreamDestroy(m->side_stream);
    for (int i = 0; i < 6; ++i) if (m->ev[i]) cudaEventDestroy(m->ev[i]);
    delete m->enqueue_mu;
    delete m->pipelined;
    delete m;
}

extern "C" size_t hssb_model_workspace_bytes(const hssb_model *m, int64_t B, int64_t T)
{
    if (!m || B <= 0 || T <= 0) return 0;
    const size_t a = simt_ws_layout(m, B, T).total;
    const size_t b = m->tc_ready ? tc_workspace_bytes(m, B, T) : 0;
    return a > b ? a : b;
}

static int model_forward(const hssb_model *m, const float *x, void *presplit, int64_t B, int64_t T, const float *h0,
                         const float *c0, float *logp, int32_t *labels, void *workspace, size_t workspace_bytes, int impl, void *stream)
{
    if (!m || !x || !h0 || !c0) return fail(HSSB_E_NULL, "hssb_model_forward: null pointer");
    if (!logp && !labels) return fail(HSSB_E_NULL, "hssb_model_forward: need logp and/or labels");
    if (B < 0 || T < 0) return fail(HSSB_E_SHAPE, "hssb_model_forward: B=%lld T=%lld", (long long)B, (long long)T);
    if (B == 0 || T == 0) return 0;
    if (B > 65535 * 4) return fail(HSSB_E_SHAPE, "hssb_model_forward: B=%lld too large for one call", (long long)B);
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(HSSB_E_WORKSPACE, "model workspace must be 256-byte aligned");
    if (reinterpret_cast<uintptr_t>(presplit) & 255) return fail(HSSB_E_WORKSPACE, "split buffer must be 256-byte aligned");
    if (int rc = require_sm100()) return rc;
    int dev = -1;
    HSSB_CUDA_OK(cudaGetDevice(&dev));
    if (dev != m->device)
        return fail(HSSB_E_DEVICE, "hssb_model_forward: model lives on device %d, current device is %d", m->device, dev);
    cudaStream_t st = as_stream(stream);
    if (impl == 1) return simt_forward(m, x, B, T, h0, c0, logp, labels, workspace, workspace_bytes, st);
    if (impl != 0) return fail(HSSB_E_MODE, "hssb_model_forward: impl %d", impl);
    // geometries the tcgen05 kernels are not specialised for run on the generic SIMT CUDA kernels
    if (!m->tc_ready) return simt_forward(m, x, B, T, h0, c0, logp, labels, workspace, workspace_bytes, st);
    return tc_forward(m, x, B, T, h0, c0, logp, labels, workspace, workspace_bytes, st, presplit);
}

extern "C" int hssb_model_forward(const hssb_model *m, const float *x, int64_t B, int64_t T, const float *h0,
                                  const float *c0, float *logp, int32_t *labels, void *workspace,
                                  size_t workspace_bytes, int impl, void *stream)
{
    return model_forward(m, x, nullptr, B, T, h0, c0, logp, labels, workspace, workspace_bytes, impl, stream);
}

extern "C" size_t hssb_model_split_bytes(const hssb_model *m, int64_t B, int64_t T)
{
    if (!m || B <= 0 || T <= 0) return 0;
    return tc_split_bytes(m, B, T);
}

extern "C" int hssb_model_split_input(const hssb_model *m, const float *x, int64_t B, int64_t T, void *planes, size_t planes_bytes,
                                      void *stream)
{
    if (!m || !x || !planes) return fail(HSSB_E_NULL, "hssb_model_split_input: null pointer");
    if (B <= 0 || T <= 0) return fail(HSSB_E_SHAPE, "hssb_model_split_input: B=%lld T=%lld", (long long)B, (long long)T);
    if (reinterpret_cast<uintptr_t>(planes) & 255) return fail(HSSB_E_WORKSPACE, "split buffer must be 256-byte aligned");
    if (int rc = require_sm100()) return rc;
    int dev = -1;
    HSSB_CUDA_OK(cudaGetDevice(&dev));
    if (dev != m->device)
        return fail(HSSB_E_DEVICE, "hssb_model_split_input: model lives on device %d, current device is %d", m->device, dev);
    if (!m->tc_ready) return 0;            // the generic kernels read x itself
    return tc_split_input(m, x, B, T, planes, planes_bytes, as_stream(stream));
}

extern "C" int hssb_model_forward_split(const hssb_model *m, const float *x, void *planes, int64_t B, int64_t T, const float *h0,
                                        const float *c0, float *logp, int32_t *labels, void *workspace, size_t workspace_bytes,
                                        void *stream)
{
    if (!planes) return fail(HSSB_E_NULL, "hssb_model_forward_split: null split buffer");
    return model_forward(m, x, planes, B, T, h0, c0, logp, labels, workspace, workspace_bytes, 0, stream);
}

extern "C" size_t hssb_lstm_workspace_bytes(const hssb_weights *w, int64_t B, int64_t T) { return hssb_model_workspace_bytes(w, B, T); }

extern "C" int hssb_lstm_forward(const hssb_weights *w, const float *x, int64_t B, int64_t T, int F, const float *h0, const float *c0,
                                 float *logp, int32_t *labels, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!w) return fail(HSSB_E_NULL, "hssb_lstm_forward: null weights");
    if (F != w->F) return fail(HSSB_E_SHAPE, "hssb_lstm_forward: F=%d but the weights were packed for input_size %d", F, w->F);
    return hssb_model_forward(w, x, B, T, h0, c0, logp, labels, workspace, workspace_bytes, 0, stream);
}

extern "C" int hssb_lstm_train_forward_tc(const hssb_model *m, int layer, const float *x, int64_t B, int64_t T, const float *h0,
                                          const float *c0, float *gates, float *out, float *cells, float *hn, float *cn,
                                          void *workspace, size_t workspace_bytes, void *stream)
{
    if (!m || !x || !h0 || !c0 || !gates || !out || !cells || !hn || !cn) return fail(HSSB_E_NULL, "hssb_lstm_train_forward_tc: null pointer");
    if (layer != 0 && layer != 1) return fail(HSSB_E_MODE, "hssb_lstm_train_forward_tc: layer %d", layer);
    if (B < 0 || T < 0) return fail(HSSB_E_SHAPE, "hssb_lstm_train_forward_tc: B=%lld T=%lld", (long long)B, (long long)T);
    if (B == 0 || T == 0) return 0;
    if (B > 65535 * 4) return fail(HSSB_E_SHAPE, "hssb_lstm_train_forward_tc: B=%lld too large for one call", (long long)B);
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(HSSB_E_WORKSPACE, "model workspace must be 256-byte aligned");
    if (int rc = require_sm100()) return rc;
    if (!m->tc_ready)
        return fail(HSSB_E_MODEL, "hssb_lstm_train_forward_tc: this model runs on the generic kernels (hidden_size %d or its weight range); "
                                  "use hssb_lstm_train_forward", m->H);
    int dev = -1;
    HSSB_CUDA_OK(cudaGetDevice(&dev));
    if (dev != m->device)
        return fail(HSSB_E_DEVICE, "hssb_lstm_train_forward_tc: model lives on device %d, current device is %d", m->device, dev);
    return tc_train_forward(m, layer, x, B, T, h0, c0, gates, out, cells, hn, cn, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int hssb_model_uses_tensor_cores(const hssb_model *m) { return m && m->tc_ready ? 1 : 0; }

extern "C" int hssb_model_side_gate(const hssb_model *m, int64_t B, int64_t T, void *workspace, void *side_stream)
{
    if (!m || !workspace) return fail(HSSB_E_NULL, "hssb_model_side_gate: null pointer");
    if (B <= 0 || T <= 0) return fail(HSSB_E_SHAPE, "hssb_model_side_gate: B=%lld T=%lld", (long long)B, (long long)T);
    if (!m->tc_ready) return 0;            // the generic kernels have no idle SMs to hand out: nothing to wait for
    int dev = -1;
    HSSB_CUDA_OK(cudaGetDevice(&dev));
    if (dev != m->device) return fail(HSSB_E_DEVICE, "hssb_model_side_gate: model lives on device %d, current device is %d", m->device, dev);
    return tc_side_gate(m, B, T, workspace, as_stream(side_stream));
}
