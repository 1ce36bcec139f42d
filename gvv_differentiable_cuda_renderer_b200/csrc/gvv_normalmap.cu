// gvv_normalmap.cu -- UV-space normal map (compute_normal_map), SURVEY.md 8(f) row 2.  Placeholder
// launchers; filled in once the raster path is parity-green.
#include "gvv_internal.h"
namespace gvv {
int launch_build_texel_table(const float*, int, int, int, float4*, cudaStream_t) { return -1; }
int launch_normal_map(const FwdArgs&, const float4*, float*, cudaStream_t) { return -1; }
}  // namespace gvv
