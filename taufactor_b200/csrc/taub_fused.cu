// taub_fused.cu -- temporally blocked two-colour sweep (placeholder until the kernel lands).
#include "taub_common.cuh"

extern "C" {

int taub_can_fuse(const taub_problem *p)
{
    (void)p;
    return 0;
}

int taub_fused_sweep2(const taub_problem *p, int64_t iter, int i_lo, int i_hi, void *stream)
{
    (void)p; (void)iter; (void)i_lo; (void)i_hi; (void)stream;
    taub::set_error("taub_fused_sweep2: not available for this problem");
    return TAUB_ERR_UNSUPPORTED;
}

}  // extern "C"
