#pragma once
namespace ros {
struct Time { double t = 0; Time() {} explicit Time(double s) : t(s) {} double toSec() const { return t; } Time& fromSec(double s) { t = s; return *this; }
  static Time now() { return Time(); } bool operator==(const Time& o) const { return t == o.t; } bool operator!=(const Time& o) const { return t != o.t; } };
struct Rate { explicit Rate(double) {} bool sleep() { return true; } };
struct Duration { explicit Duration(double) {} bool sleep() { return true; } };
}
