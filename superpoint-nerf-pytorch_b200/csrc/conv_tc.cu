// placeholder until the tcgen05 implicit-GEMM path lands
#include "spn_common.cuh"
int spn_tc_pack_layer(spn_ctx*, int, const float*, const float*, cudaStream_t) { return SPN_OK; }
int spn_tc_encoder(spn_ctx*, const float*, int, int, int, int, cudaStream_t) { spn_set_error("tcgen05 path not built"); return SPN_E_STATE; }
int spn_tc_detector_head(spn_ctx*, int, int, int, int, float*, cudaStream_t) { spn_set_error("tcgen05 path not built"); return SPN_E_STATE; }
int spn_tc_descriptor_head(spn_ctx*, int, int, int, int, float*, cudaStream_t) { spn_set_error("tcgen05 path not built"); return SPN_E_STATE; }
void spn_tc_destroy(spn_ctx*) {}
