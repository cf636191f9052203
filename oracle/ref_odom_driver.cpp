// TEST INFRASTRUCTURE.  Compiles the reference's laserOdometry node AS IT LIES under /root/reference (the translation
// unit, with its lidarFactor.hpp, is #included below; nothing is copied) against the functional stand-ins of
// oracle/refstubs/.  The node's main loop (Aloam/src/laserOdometry.cpp:220-598) runs unchanged: the stand-in spinOnce()
// hands it one sweep's five clouds per turn, ok() ends the loop when the deliveries are used up.  Correspondence search,
// TransformToStart, factor construction and the pose accumulation are reference code; kd-tree, Eigen arithmetic and the
// Ceres minimiser are the stand-ins (the minimiser being oracle/lm.c driven by the reference's cost functors).
// Also exports the three cost functors of lidarFactor.hpp evaluated on dual numbers, for direct comparison.
// Built by `make -C oracle ref` into oracle/_ref/libref_odom.so only where /root/reference exists.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <ctime>
#include <iostream>
#include <mutex>
#include <queue>
#include <string>
#include <vector>
#include <ros/ros.h>
#define printf(...) ((void)0)                 /* the node reports timings on stdout */
#define main ref_odom_main
#include "laserOdometry.cpp"                  /* -I/root/reference/Aloam/src */
#undef main
#undef printf

static sensor_msgs::PointCloud2ConstPtr make_msg(const float* xyzi, int n, double stamp) {
  pcl::PointCloud<PointType> c;
  c.points.resize((size_t)n);
  for (int i = 0; i < n; ++i) { c.points[i].x = xyzi[4 * i]; c.points[i].y = xyzi[4 * i + 1]; c.points[i].z = xyzi[4 * i + 2]; c.points[i].intensity = xyzi[4 * i + 3]; }
  sensor_msgs::PointCloud2Ptr m(new sensor_msgs::PointCloud2());
  pcl::toROSMsg(c, *m);
  m->header.stamp = ros::Time().fromSec(stamp);
  return m;
}

// one call per sweep; the node's globals carry the state from sweep to sweep like the running node does.
// pose_out: q_w_curr (x y z w), t_w_curr, then q_last_curr (x y z w), t_last_curr; counts: corner / plane correspondences of the last pass
extern "C" int ref_odom_step(const float* sharp, int n_sharp, const float* less_sharp, int n_less_sharp,
                             const float* flat, int n_flat, const float* less_flat, int n_less_flat,
                             const float* full, int n_full, double stamp, double pose_out[14], int32_t counts[2]) {
  refstub::State& S = refstub::state();
  S.params["mapping_skip_frame"] = 1;
  S.clouds.clear(); S.odoms.clear(); S.deliveries.clear(); S.next = 0;
  refstub::Delivery d;
  d.clouds.push_back({ "/laser_cloud_sharp", make_msg(sharp, n_sharp, stamp) });
  d.clouds.push_back({ "/laser_cloud_less_sharp", make_msg(less_sharp, n_less_sharp, stamp) });
  d.clouds.push_back({ "/laser_cloud_flat", make_msg(flat, n_flat, stamp) });
  d.clouds.push_back({ "/laser_cloud_less_flat", make_msg(less_flat, n_less_flat, stamp) });
  d.clouds.push_back({ "/velodyne_cloud_2", make_msg(full, n_full, stamp) });
  S.deliveries.push_back(d);
  int argc = 1; char arg0[] = "alaserOdometry"; char* argv[] = { arg0, nullptr };
  ref_odom_main(argc, argv);                   // subscribes, then runs the node's loop until the delivery is processed
  auto it = S.odoms.find("/laser_odom_to_init");
  if (it == S.odoms.end() || it->second.size() != 1) return -1;
  const nav_msgs::Odometry& o = it->second[0];
  pose_out[0] = o.pose.pose.orientation.x; pose_out[1] = o.pose.pose.orientation.y; pose_out[2] = o.pose.pose.orientation.z; pose_out[3] = o.pose.pose.orientation.w;
  pose_out[4] = o.pose.pose.position.x; pose_out[5] = o.pose.pose.position.y; pose_out[6] = o.pose.pose.position.z;
  for (int k = 0; k < 4; ++k) pose_out[7 + k] = para_q[k];
  for (int k = 0; k < 3; ++k) pose_out[11 + k] = para_t[k];
  counts[0] = corner_correspondence; counts[1] = plane_correspondence;
  return 0;
}

// a fresh node: the globals back to their initial values (laserOdometry.cpp:66-101)
extern "C" void ref_odom_reset(void) {
  std::cout.setstate(std::ios_base::failbit);          // "Initialization finished" (:265)
  systemInited = false;
  q_w_curr = Eigen::Quaterniond(1, 0, 0, 0); t_w_curr = Eigen::Vector3d(0, 0, 0);
  para_q[0] = para_q[1] = para_q[2] = 0; para_q[3] = 1; para_t[0] = para_t[1] = para_t[2] = 0;
  laserCloudCornerLast.reset(new pcl::PointCloud<PointType>()); laserCloudSurfLast.reset(new pcl::PointCloud<PointType>());
  laserCloudCornerLastNum = laserCloudSurfLastNum = 0;
  corner_correspondence = plane_correspondence = 0;
  while (!cornerSharpBuf.empty()) cornerSharpBuf.pop();
  while (!cornerLessSharpBuf.empty()) cornerLessSharpBuf.pop();
  while (!surfFlatBuf.empty()) surfFlatBuf.pop();
  while (!surfLessFlatBuf.empty()) surfLessFlatBuf.pop();
  while (!fullPointsBuf.empty()) fullPointsBuf.pop();
}

// lidarFactor.hpp on dual numbers: type 0 LidarEdgeFactor (p, a, b, s), 1 LidarPlaneFactor (p, j = a, l = b, m = c, s),
// 2 LidarPlaneNormFactor (p, unit normal = a, negative_OA_dot_norm = b[0]).  r[3], Jq[3][4] (x y z w), Jt[3][3]; returns the residual count.
extern "C" int ref_factor_eval(int type, const double p[3], const double a[3], const double b[3], const double c[3], double s,
                               const double q[4], const double t[3], double* r, double* Jq, double* Jt) {
  const Eigen::Vector3d P(p[0], p[1], p[2]), A(a[0], a[1], a[2]), B(b[0], b[1], b[2]);
  ceres::CostFunction* f = type == 0 ? LidarEdgeFactor::Create(P, A, B, s)
                         : type == 1 ? LidarPlaneFactor::Create(P, A, B, Eigen::Vector3d(c[0], c[1], c[2]), s)
                                     : LidarPlaneNormFactor::Create(P, A, b[0]);
  double const* params[2] = { q, t };
  double* jac[2] = { Jq, Jt };
  const int n = f->num_residuals();
  const bool ok = f->Evaluate(params, r, jac);
  delete f;
  return ok ? n : -1;
}

// A Ceres problem assembled the way the reference assembles it (laserOdometry.cpp:284-291,380-381,494-499 /
// laserMapping.cpp:563-571,618-619,683-684,713-720) from factor records of type 0 (LidarEdgeFactor) and 2
// (LidarPlaneNormFactor): HuberLoss(0.1), EigenQuaternionParameterization, DENSE_QR, max_num_iterations as given.
static int build_problem(ceres::Problem& problem, const o_factor* f, int nf, double* pq, double* pt) {
  ceres::LossFunction* loss_function = new ceres::HuberLoss(0.1);
  ceres::LocalParameterization* q_parameterization = new ceres::EigenQuaternionParameterization();
  problem.AddParameterBlock(pq, 4, q_parameterization);
  problem.AddParameterBlock(pt, 3);
  for (int i = 0; i < nf; ++i) {
    const Eigen::Vector3d P(f[i].p[0], f[i].p[1], f[i].p[2]), A(f[i].a[0], f[i].a[1], f[i].a[2]), B(f[i].b[0], f[i].b[1], f[i].b[2]);
    ceres::CostFunction* c;
    if (f[i].type == O_FACTOR_EDGE) c = LidarEdgeFactor::Create(P, A, B, f[i].s != 0.0 ? f[i].s : 1.0);
    else if (f[i].type == O_FACTOR_PLANE_NORM) c = LidarPlaneNormFactor::Create(P, A, f[i].b[0]);
    else { delete loss_function; return -1; }
    problem.AddResidualBlock(c, loss_function, pq, pt);
  }
  if (nf == 0) delete loss_function;
  return 0;
}
extern "C" int ref_lm_solve(const o_factor* f, int nf, o_pose* x, int max_iter, o_solve_summary* sum) {
  ceres::Problem::Options problem_options;
  ceres::Problem problem(problem_options);
  if (build_problem(problem, f, nf, x->q, x->t)) return -1;
  ceres::Solver::Options options;
  options.linear_solver_type = ceres::DENSE_QR;
  options.max_num_iterations = max_iter;
  options.minimizer_progress_to_stdout = false;
  ceres::Solver::Summary summary;
  ceres::Solve(options, &problem, &summary);
  *sum = summary.oracle;
  return 0;
}
extern "C" int ref_normal_eq(const o_factor* f, int nf, const o_pose* x, double H[36], double g[6], double* cost) {
  ceres::Problem problem;
  o_pose xc = *x;
  if (build_problem(problem, f, nf, xc.q, xc.t)) return -1;
  lmono_cpu_lm_set_block_hook(ceres::refstub_detail::block_hook, &problem);
  const int rc = lmono_cpu_normal_eq(f, nf, x, H, g, cost);
  lmono_cpu_lm_set_block_hook(nullptr, nullptr);
  return rc;
}
