// Entry points of stages that have not landed yet; each is replaced by its real
// implementation file (scanreg.cu, odom.cu, color.cu) as the stage is built.
#include "common.cuh"
extern "C" int lmono_project_color(lmono_ctx*, lmono_cloud_view, const uint8_t*, int32_t, const lmono_pinhole*, const lmono_pose*,
                                   uint8_t*, uint8_t*, float*, float*, uint8_t*, int32_t, int32_t*) { return LMONO_E_STATE; }
