#pragma once
// ceres/rotation.h: only named inside camodocal's templated projection helpers, which the tests do not instantiate
namespace ceres { template <class T> inline void QuaternionRotatePoint(const T q[4], const T pt[3], T result[3]) {
  const T t2 = q[0] * q[1], t3 = q[0] * q[2], t4 = q[0] * q[3], t5 = -q[1] * q[1], t6 = q[1] * q[2], t7 = q[1] * q[3], t8 = -q[2] * q[2], t9 = q[2] * q[3], t1 = -q[3] * q[3];
  result[0] = T(2) * ((t8 + t1) * pt[0] + (t6 - t4) * pt[1] + (t3 + t7) * pt[2]) + pt[0];
  result[1] = T(2) * ((t4 + t6) * pt[0] + (t5 + t1) * pt[1] + (t9 - t2) * pt[2]) + pt[1];
  result[2] = T(2) * ((t7 - t3) * pt[0] + (t2 + t9) * pt[1] + (t5 + t8) * pt[2]) + pt[2]; } }
