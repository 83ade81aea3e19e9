// stokes3d_fused.cu — placeholder until the fused kernel lands (next commit).
#include "common.cuh"
int jr_stokes3d_VA_fused_supported(const jr_fields *, const jr_stokes_opts *) { return JR_ERR_UNSUPPORTED; }
int jr_stokes3d_VA_fused_iter(jr_context *, const jr_fields *, const jr_stokes_opts *, int, int)
{
    jr_set_error("fused 3D-VA kernel not built");
    return JR_ERR_UNSUPPORTED;
}
