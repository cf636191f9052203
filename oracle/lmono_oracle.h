/*
 * lmono_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the LiDAR registration hot path of bobocode/lmono
 * (A-LOAM scanRegistration -> laserOdometry -> laserMapping, plus the
 * mono_lidar_mapping colour projection).  It exists only so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * can check and time the CUDA path against it.  Nothing under lmono_b200/
 * includes, links or calls anything in this directory.
 *
 * PARITY PIN.  The reference ships no tests, fixtures or golden vectors
 * (SURVEY.md section 4), and its third-party dependencies (ROS, PCL, FLANN,
 * Ceres, Eigen, OpenCV) are not installable here.  What pins this oracle:
 * (1) THE REFERENCE'S OWN SOURCES RUN HERE: `make ref` compiles the translation
 * units Aloam/src/scanRegistration.cpp, laserOdometry.cpp, laserMapping.cpp,
 * lidarFactor.hpp, mono_lidar_mapping/src/map_build_node.cc +
 * map_builder/Map_Builder.cc and camera_models/src/camera_models/Camera.cc +
 * PinholeCamera.cc where they lie under /root/reference (a driver #includes
 * them, nothing is copied) against functional stand-ins for those libraries
 * (refstubs/), into oracle/_ref/; tests/test_oracle_vs_ref.py runs the nodes'
 * own callbacks / main loops beside this oracle: full cloud, curvature, labels
 * and feature clouds bit for bit; odometry correspondences and poses; mapping
 * poses, registered clouds and every cube of the map bit for bit through all
 * six window shifts; solver traces with the reference cost functors on dual
 * numbers; the colour mapper's raster, filled depth image and lifted clouds
 * bit for bit.  Everything between the library calls is therefore pinned against
 * reference code compiled by this toolchain.
 * (2) THE LIBRARY ALGORITHMS remain restatements of the published code of the
 * un-vendored dependencies (PCL 1.8 VoxelGrid / KdTreeFLANN, FLANN 1.8/1.9
 * KDTreeSingleIndex, Ceres 1.14 trust-region LM, Eigen 3.3
 * SelfAdjointEigenSolver / ColPivHouseholderQR, OpenCV 3.2 morphology / blur),
 * unpinned against those libraries' binaries; they are cross-checked against
 * scipy / numpy / cv2 and independent Python restatements in tests/
 * (test_oracle_primitives.py, test_oracle_*_python.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#ifndef LMONO_ORACLE_H
#define LMONO_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Point layout used by the oracle API: packed XYZI, 16 bytes. */
typedef struct { float x, y, z, i; } o_pt;

typedef struct { double q[4]; /* x,y,z,w */ double t[3]; } o_pose;

/* ---- PCL VoxelGrid<PointXYZI> ---------------------- */
/* order_mode 0: members of a voxel are summed in input-index order (the
 *               canonical order the CUDA path reproduces bit-exactly);
 * order_mode 1: members are summed in the order libstdc++ std::sort leaves
 *               them, exactly as PCL does (unstable sort artefact). */
int lmono_cpu_voxel_grid(const o_pt* in, int n, float leaf, int order_mode,
                         o_pt* out, int* n_out);

/* ---- k-NN (PCL KdTreeFLANN / FLANN KDTreeSingleIndex, App. B.1-B.2) ------ */
/* brute force, (d2, index) ascending; d2 = ((dx*dx)+dy*dy)+dz*dz in fp32 */
int lmono_cpu_knn_brute(const o_pt* pts, int n, const o_pt* queries, int nq,
                        int k, int32_t* idx, float* d2);
typedef struct o_kdtree o_kdtree;
o_kdtree* lmono_cpu_kdtree_build(const o_pt* pts, int n);
void      lmono_cpu_kdtree_free(o_kdtree*);
int       lmono_cpu_kdtree_knn(const o_kdtree*, const o_pt* queries, int nq,
                               int k, int32_t* idx, float* d2);

/* ---- small dense algebra (Eigen 3.3 restatements, App. B.5) -------------- */
/* ascending eigenvalues w[3], eigenvectors in columns of V (row-major V[r*3+c]) */
int lmono_cpu_eigh3(const double A[9], double w[3], double V[9]);
/* least squares A(5x3, row-major) x = b via column-pivoted Householder QR */
int lmono_cpu_colpiv_qr_solve_5x3(const double A[15], const double b[5], double x[3]);

/* ---- factors (lidarFactor.hpp) ------------------------------------------- */
enum { O_FACTOR_EDGE = 0, O_FACTOR_PLANE = 1, O_FACTOR_PLANE_NORM = 2 };
typedef struct {
  int32_t type;
  int32_t pad;
  double p[3];   /* curr_point (sensor frame) */
  double a[3];   /* EDGE: last_point_a ; PLANE: last_point_j ; PLANE_NORM: unit normal */
  double b[3];   /* EDGE: last_point_b ; PLANE: ljm_norm (unit) ; PLANE_NORM: b[0] = negative_OA_dot_norm */
  double s;      /* EDGE / PLANE: interpolation ratio (lidarFactor.hpp:14,60; 1.0 with DISTORTION 0); 0.0 is read as 1.0 (zero-initialised records) */
} o_factor;

/* cost = sum 0.5*rho(|r|^2) with HuberLoss(0.1); H = J^T J (6x6, row-major),
 * g = J^T r, both with the Ceres loss "corrector" applied, in the 6-dim
 * tangent space (rotation first, translation second). */
int lmono_cpu_normal_eq(const o_factor* f, int nf, const o_pose* x,
                        double H[36], double g[6], double* cost);

typedef struct {
  int32_t iterations;        /* LM iterations taken (successful + unsuccessful) */
  int32_t num_successful;
  int32_t termination;       /* 0 max-iter, 1 gradient tol, 2 parameter tol, 3 function tol, 4 radius, 5 failure, 6 no residuals */
  int32_t num_factors;
  double initial_cost;
  double final_cost;
} o_solve_summary;

/* Ceres 1.14 TrustRegionMinimizer + LevenbergMarquardtStrategy + DENSE_QR,
 * max_num_iterations as given (reference uses 4), Huber(0.1),
 * EigenQuaternionParameterization (App. B.4). x is updated in place. */
int lmono_cpu_lm_solve(const o_factor* f, int nf, o_pose* x, int max_iter,
                       o_solve_summary* sum);

/* test hook for the oracle/_ref builds: residuals r[<=3] and the Jacobians d r / d q (x,y,z,w) and d r / d t of block i
 * at (q, t) come from the callback (returns the residual count) instead of the factor record */
typedef int (*o_block_hook)(void* user, int i, const double q[4], const double t[3], double r[3], double Jq[3][4], double Jt[3][3], int want_jac);
void lmono_cpu_lm_set_block_hook(o_block_hook fn, void* user);

/* ---- laserMapping (Aloam/src/laserMapping.cpp) --------------------------- */
typedef struct o_mapper o_mapper;
typedef struct {
  int32_t corner_from_map, surf_from_map;     /* laserMapping.cpp:538-539 */
  int32_t corner_stack, surf_stack;           /* :545,550 */
  int32_t corner_num[2], surf_num[2];         /* factors per outer iteration :620,685 */
  int32_t optimized;                          /* :554 gate */
  int32_t center_cube[3];
  int32_t cen[3];
  o_solve_summary solve[2];
  double ms_shift, ms_tree, ms_assoc, ms_solver, ms_add, ms_filter, ms_whole; /* TicToc names of the reference */
} o_map_report;

o_mapper* lmono_cpu_mapper_create(float line_res, float plane_res, int voxel_order_mode, int use_kdtree);
void      lmono_cpu_mapper_destroy(o_mapper*);
/* append points to the cube they fall in (no transform), then VoxelGrid every
 * non-empty cube: produces the "filtered" state the reference reaches after
 * :788-801.  which: 0 corner, 1 surf. */
int lmono_cpu_mapper_import(o_mapper*, int which, const o_pt* pts, int n);
/* which: 0 corner, 1 surf; scope 0: valid 5x5x3 window in laserMapping.cpp:512-537 order,
 * scope 1: all 4851 cubes in :826-830 order.  Returns count; writes up to cap. */
int lmono_cpu_mapper_export(o_mapper*, int which, int scope, o_pt* out, int cap);
void lmono_cpu_mapper_get_state(const o_mapper*, o_pose* wmap_wodom, int32_t cen[3]);
void lmono_cpu_mapper_set_state(o_mapper*, const o_pose* wmap_wodom);
/* one pass of process() :307-801 (+ :838-842 if full_res given) */
int lmono_cpu_map_step(o_mapper*, const o_pt* corner_last, int nc, const o_pt* surf_last, int ns,
                       const o_pose* wodom_curr, o_pose* w_curr, o_map_report* rep,
                       o_pt* full_res_inout, int nfull);
/* test hook: window shift + valid list + concatenation (:312-539) for a pose translation */
int lmono_cpu_mapper_prepare_window(o_mapper*, const double t_w_curr[3]);
/* test hook: 5-NN of world-frame queries against the current valid window
 * (indices in :533-537 concatenation order); call prepare_window first. */
int lmono_cpu_mapper_knn5(o_mapper*, int which, const o_pt* queries_world, int nq,
                          int32_t* idx, float* d2);
/* test hook: factors built by one association pass at pose x (no solve).
 * Returns number of factors (corner first, then surf), up to cap. */
int lmono_cpu_mapper_associate(o_mapper*, const o_pt* corner_stack, int nc, const o_pt* surf_stack, int ns,
                               const o_pose* w_curr, o_factor* out, int cap, int32_t* n_corner, int32_t* n_surf);

/* ---- scanRegistration (Aloam/src/scanRegistration.cpp) ------------------- */
typedef struct {
  int32_t n_in, n_kept;                 /* after NaN / range / ring filters */
  int32_t n_sharp, n_less_sharp, n_flat, n_less_flat;
  int32_t ring_start[64], ring_end[64]; /* scanStartInd / scanEndInd :249-251 */
  float start_ori, end_ori;
  int32_t min_margin_ok;               /* 1 if every kept point's ring id is stable under +-1e-3 deg */
} o_scan_report;
int lmono_cpu_scan_register(const float* xyz_in, int n_in, int stride_floats,
                            int n_scans, float minimum_range, int voxel_order_mode,
                            int sort_mode,  /* 0 stable (curv,index), 1 libstdc++ std::sort */
                            o_pt* full, o_pt* sharp, o_pt* less_sharp, o_pt* flat, o_pt* less_flat,
                            int32_t* labels, float* curvature, int32_t* src_index,
                            o_scan_report* rep);

/* 0: double atan/sqrt (GCC 5.4-era headers, default), 1: float overloads (scanRegistration.cpp:166) */
void lmono_cpu_scan_set_trig_mode(int mode);

/* ---- laserOdometry (Aloam/src/laserOdometry.cpp) ------------------------- */
typedef struct o_odom o_odom;
typedef struct {
  int32_t inited;                       /* 0 on the first frame :267-271 */
  int32_t corner_corr[2], plane_corr[2];
  o_solve_summary solve[2];
  double ms_assoc, ms_solver, ms_whole;
} o_odom_report;
o_odom* lmono_cpu_odom_create(void);
void    lmono_cpu_odom_destroy(o_odom*);
/* #define DISTORTION (Aloam/src/laserOdometry.cpp:59): 1 = every point is interpolated to the sweep start with
 * s = (intensity - int(intensity)) / SCAN_PERIOD (:111-129) and the factors carry that s (:374-381,472-479) */
void    lmono_cpu_odom_set_distortion(o_odom*, int on);
int lmono_cpu_odom_step(o_odom*, const o_pt* sharp, int n_sharp, const o_pt* less_sharp, int n_less_sharp,
                        const o_pt* flat, int n_flat, const o_pt* less_flat, int n_less_flat,
                        o_pose* last_curr /*out*/, o_pose* w_curr /*out*/, o_odom_report* rep);
/* test hook: correspondences of one association pass.
 * corner_idx: n_sharp x 2 (closest, min2) ; plane_idx: n_flat x 3 (closest, min2, min3); -1 = none */
int lmono_cpu_odom_associate(const o_pt* sharp, int n_sharp, const o_pt* flat, int n_flat,
                             const o_pt* corner_last, int n_cl, const o_pt* surf_last, int n_sl,
                             const o_pose* last_curr, int32_t* corner_idx, int32_t* plane_idx);

/* ---- colour projection (Map_Builder.cc:213-416, PinholeCamera.cc) -------- */
typedef struct {
  double fx, fy, cx, cy, k1, k2, p1, p2;
  int32_t width, height;
  int32_t kernel_type;   /* 0 FULL(rect) 1 CROSS 2 ELLIPSE */
  int32_t kernel_size;
  int32_t blur_type;     /* 0 bilateral 1 gaussian */
} o_camera;
/* pts_cam: camera-frame XYZ (after map_build_node.cc:216-225), bgr: h x w x 3.
 * depth_raw: raster before depthFill, depth_u8: after. rgb_cloud: 8 floats per point
 * (x,y,z,pad,r,g,b,pad as floats) in camera frame and world frame. */
/* D1: pcl::transformPointCloud with a double 3x4 (row-major) matrix, float store */
int lmono_cpu_transform_cloud(const float* in, int n, int stride_floats, const double T[12], float* out_xyz);
int lmono_cpu_project_raster(const float* pts_cam, int n, int stride_floats, const o_camera* cam,
                             uint8_t* depth_raw);
int lmono_cpu_depth_fill(const uint8_t* depth_raw, const o_camera* cam, uint8_t* depth_out);
int lmono_cpu_lift_cloud(const uint8_t* depth, const uint8_t* bgr, const o_camera* cam,
                         const o_pose* QT, float* cloud_cam_xyz, float* cloud_world_xyz,
                         uint8_t* cloud_rgb, int cap, int* n_out);

/* test hooks for the oracle/_ref builds: the OpenCV 8UC1 restatements of color.c one by one (kernel type 0 RECT 1 CROSS 2 ELLIPSE) */
void lmono_cpu_cv_kernel(int type, int ks, uint8_t* k);
void lmono_cpu_cv_morph(const uint8_t* src, uint8_t* dst, int W, int H, const uint8_t* k, int ks, int is_erode);
void lmono_cpu_cv_median5(const uint8_t* src, uint8_t* dst, int W, int H);
void lmono_cpu_cv_bilateral5(const uint8_t* src, uint8_t* dst, int W, int H, double sigma_color, double sigma_space);
void lmono_cpu_cv_gaussian5(const uint8_t* src, uint8_t* dst, int W, int H);

/* libstdc++ std::sort shims (stdsort_shim.cpp): reproduce the reference's unstable sorts */
void lmono_cpu_stdsort_voxel_pairs(uint32_t* idx, uint32_t* pt, int n);          /* PCL cloud_point_index_idx operator< */
void lmono_cpu_stdsort_by_curvature(int32_t* ind, int n, const float* curvature); /* scanRegistration.cpp:71,288 */

#ifdef __cplusplus
}
#endif
#endif
