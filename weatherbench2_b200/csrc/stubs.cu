// Entry points whose kernels are not built yet return WB2_EUNSUPPORTED loudly.
#include "common.cuh"
using namespace wb2;
extern "C" {
int wb2_regrid_conservative(wb2_ctx*, const float*, float*, int64_t, int64_t, int64_t,
                            const wb2_csr*, const wb2_csr*) {
  set_error("wb2_regrid_conservative: not implemented in this build");
  return WB2_EUNSUPPORTED;
}
int wb2_zonal_spectrum(wb2_ctx*, const float*, int64_t, int32_t, int32_t, const double*, float*,
                       int32_t, int64_t) {
  set_error("wb2_zonal_spectrum: not implemented in this build");
  return WB2_EUNSUPPORTED;
}
}
