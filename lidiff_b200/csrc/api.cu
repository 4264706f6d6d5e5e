// Dispatch of lb2_spconv_forward to the CUDA-core (spconv_ffma.cu) or tcgen05 (spconv_tc.cu) variant.
#include "common.cuh"

int lb2_spconv_ffma_launch(Lb2Handle* h, cudaStream_t s, const lb2_conv_desc* d);
int lb2_spconv_tc_launch(Lb2Handle* h, cudaStream_t s, const lb2_conv_desc* d, bool persistent);
bool lb2_spconv_tc_supported(const lb2_conv_desc* d);

extern "C" int lb2_spconv_forward(void* handle, void* stream, const lb2_conv_desc* d, int algo) {
    Lb2Handle* h = (Lb2Handle*)handle;
    LB2_REQUIRE(h, h && d, "spconv_forward null");
    LB2_REQUIRE(h, d->npass == 1 || d->npass == 2, "npass must be 1 or 2");
    LB2_REQUIRE(h, d->c1 > 0 && d->c2 >= 0 && d->cout > 0 && d->kvol > 0 && d->mout_cap > 0, "shape");
    LB2_REQUIRE(h, d->weight != nullptr, "weight");
    LB2_REQUIRE(h, (d->scale == nullptr) == (d->shift == nullptr), "scale/shift must come together");
    LB2_REQUIRE(h, d->nbr != nullptr || d->kvol == 1, "identity map only for kvol == 1");
    LB2_REQUIRE(h, d->nbr == nullptr || d->nbr_stride >= d->mout_cap, "nbr_stride");
    for (int p = 0; p < d->npass; ++p) {
        LB2_REQUIRE(h, d->io[p].in1 != nullptr, "in1");
        LB2_REQUIRE(h, (d->c2 > 0) == (d->io[p].in2 != nullptr), "in2 / c2 mismatch");
        LB2_REQUIRE(h, d->io[p].out || d->io[p].out_gated, "no output");
    }
    LB2_REQUIRE(h, d->c2 == 0 || d->c1 % 16 == 0, "c1 must be a multiple of 16 when in2 is given");
    cudaStream_t s = (cudaStream_t)stream;
    if (algo == LB2_ALGO_TC || algo == LB2_ALGO_TC_TILE) {
        if (!d->weight_packed || !lb2_spconv_tc_supported(d))
            return lb2_fail(h, LB2_ERR_UNSUP, "tensor-core variant does not support this layer%s", "");
        return lb2_spconv_tc_launch(h, s, d, algo == LB2_ALGO_TC);
    }
    if (algo == LB2_ALGO_AUTO && d->weight_packed && lb2_spconv_tc_supported(d)) return lb2_spconv_tc_launch(h, s, d, true);
    return lb2_spconv_ffma_launch(h, s, d);
}
