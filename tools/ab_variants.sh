# A/B of experiment builds (libhm_matcher_<name>.so) against the shipped library, in ONE job: C4 and the 8-GPU shard shape
cd "$(dirname "$0")/.."
for rep in 1 2; do
for so in libhm_matcher.so $(cd slam_experiments_b200 && ls libhm_matcher_*.so | grep -v "trace\|r01"); do
  for sz in 8192000 1024000; do
    HM_TP_DIST=M HM_TP_ITERS=100 HM_MATCHER_SO=$PWD/slam_experiments_b200/$so python tools/time_prepared.py 2000 $sz | tail -1
  done
done
done
