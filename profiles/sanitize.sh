#!/bin/bash
# compute-sanitizer over one pass of every stage (profiles/one_sweep.py: plain launches, every kernel of the library once
# or more: registrations in both kNN forms, a fused HDL-64 sweep, an HDL-32 odometry step, a colour frame).
#   bash profiles/sanitize.sh [tools...]      default: memcheck synccheck racecheck
# memcheck / synccheck look at every kernel of the process (torch's workload generator included); racecheck is restricted
# to the library's kernels (names k_*), it takes minutes otherwise.  Logs: gpurun_out/sanitize_<tool>.log
# (initcheck is not usable here: with the filter it cannot see torch's writes, without it torch's caching allocator
#  hands out recycled blocks that it reports as uninitialised.)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=${@:-memcheck synccheck racecheck}
for t in $TOOLS; do
  extra=""
  [ $t = memcheck ] && extra="--leak-check no"
  [ $t = racecheck ] && extra="--kernel-name kns=k_"
  timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool $t $extra --print-limit 30 --error-exitcode 9 \
      --log-file gpurun_out/sanitize_$t.log python profiles/one_sweep.py > gpurun_out/sanitize_$t.out 2>&1
  rc=$?
  echo "== $t: exit $rc; $(grep "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/sanitize_$t.log | tail -1)"
  tail -1 gpurun_out/sanitize_$t.out
done
