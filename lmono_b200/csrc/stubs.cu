// Entry points of stages that have not landed yet; each is replaced by its real
// implementation file (scanreg.cu, odom.cu, color.cu) as the stage is built.
#include "common.cuh"
extern "C" int lmono_scan_register(lmono_ctx*, lmono_cloud_view, lmono_cloud_out*, lmono_cloud_out*, lmono_cloud_out*,
                                   lmono_cloud_out*, lmono_cloud_out*, int32_t*, lmono_scan_report*) { return LMONO_E_STATE; }
extern "C" int lmono_odom_step(lmono_ctx*, lmono_cloud_view, lmono_cloud_view, lmono_cloud_view, lmono_cloud_view,
                               lmono_pose*, lmono_pose*, lmono_odom_report*) { return LMONO_E_STATE; }
extern "C" int lmono_odom_reset(lmono_ctx*) { return LMONO_E_STATE; }
extern "C" int lmono_project_color(lmono_ctx*, lmono_cloud_view, const uint8_t*, int32_t, const lmono_pinhole*, const lmono_pose*,
                                   uint8_t*, uint8_t*, float*, float*, uint8_t*, int32_t, int32_t*) { return LMONO_E_STATE; }
