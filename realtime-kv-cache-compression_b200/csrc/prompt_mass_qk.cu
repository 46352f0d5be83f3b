// prompt_mass_qk.cu — K1b: fused Q*K_prompt^T scoring on tcgen05 (placeholder until the
// tensor-core kernel lands; the entry point reports "unsupported" instead of guessing).
#include "common.cuh"

extern "C" int rtkv_prompt_mass_qk(const void *, const void *, const float *, int, int, int, int, int, int,
                                   float, int, float *, void *)
{
    rtkv::set_error("rtkv_prompt_mass_qk: not built in this version");
    return RTKV_EUNSUPPORTED;
}
