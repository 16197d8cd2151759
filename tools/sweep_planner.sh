# planner sweep: device time of the prepared k-NN call over HM_PROLOGUE_TILES x shapes (one process per point)
cd "$(dirname "$0")/.."
for shape in "2000 1024000" "2000 2048000" "2000 4096000" "16384 16384" "32768 32768" "65536 65536" "10000 10000" "4096 4096"; do
  for p in 10 25 45 90; do
    echo -n "prologue=$p "; HM_PROLOGUE_TILES=$p HM_TP_DIST=M HM_TP_ITERS=50 python tools/time_prepared.py $shape | tail -1
  done
done
