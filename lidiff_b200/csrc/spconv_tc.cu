// K4 (variant B) placeholder until the tcgen05 kernel lands.
#include "common.cuh"
bool lb2_spconv_tc_supported(const lb2_conv_desc*) { return false; }
int lb2_spconv_tc_launch(Lb2Handle* h, cudaStream_t, const lb2_conv_desc*) { return lb2_fail(h, LB2_ERR_UNSUP, "tc not built%s", ""); }
extern "C" size_t lb2_packed_weight_bytes(int32_t, int32_t, int32_t) { return 0; }
extern "C" int lb2_pack_weights(void* h, void*, const float*, int32_t, int32_t, int32_t, void*) { return lb2_fail((Lb2Handle*)h, LB2_ERR_UNSUP, "tc not built%s", ""); }
