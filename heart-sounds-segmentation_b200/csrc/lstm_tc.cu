// tcgen05 path of the BiLSTM segmenter (placeholder until the kernels land).
#include "model.cuh"
namespace hssb {
size_t tc_pack_bytes(int, int) { return 0; }
int tc_pack(hssb_model *m, const hssb_model_params *, void *, cudaStream_t) { m->tc_ready = false; return 0; }
size_t tc_workspace_bytes(const hssb_model *, int64_t, int64_t) { return 0; }
int tc_forward(const hssb_model *, const float *, int64_t, int64_t, const float *, const float *, float *, int32_t *,
               void *, size_t, cudaStream_t)
{
    return fail(HSSB_E_MODEL, "tcgen05 path not built");
}
}  // namespace hssb
