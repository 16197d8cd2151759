cd /root/repo
for rep in 1 2; do
for so in libhm_matcher_r01.so libhm_matcher.so; do
  for dist in U M; do
    echo -n "$dist iters=100 "; HM_TP_DIST=$dist HM_TP_ITERS=100 HM_MATCHER_SO=$PWD/slam_experiments_b200/$so python tools/time_prepared.py 2000 8192000 | tail -1
  done
done
done
for so in libhm_matcher_r01.so libhm_matcher.so; do
  echo -n "M shard iters=200 "; HM_TP_DIST=M HM_TP_ITERS=200 HM_MATCHER_SO=$PWD/slam_experiments_b200/$so python tools/time_prepared.py 2000 1024000 | tail -1
done
