#!/bin/bash
# Round-2 evidence run (one B200): bench lines of every config, the ncu launch list of the default bench command, DRAM traffic of
# one step per config (cache control off: L2 state carries across kernels as in the real step) and full ncu captures of the
# dominant kernels.  Everything lands in gpurun_out/; the summaries are copied into profiles/ afterwards.
set -x
O=gpurun_out
python bench.py --steps 20 --warmup 3 > $O/bench_r02_n1.json 2> $O/bench_r02_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_r02_reference.json 2> $O/bench_r02_reference.err
for c in 2rhs 2general 1 3 4 5; do
  python bench.py --config $c --steps 5 --warmup 3 > $O/bench_r02_c$c.json 2> $O/bench_r02_c$c.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/launches_r02.out 2>&1
for c in 2 3 4 5; do
  n=$(python -c "print({'2':256,'3':128,'4':70,'5':192}['$c'])")
  GB200_GRAPH=0 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file $O/traffic_r02_c$c.csv python scripts/traffic_headline.py $c $n > $O/traffic_r02_c$c.out 2>&1
  python scripts/ncu_traffic_sum.py $O/traffic_r02_c$c.csv > $O/traffic_r02_c$c.json
done
GB200_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"q1hex_gather|cell_geom" -s 4 -c 2 -o $O/r02_headline -f python scripts/traffic_headline.py 2 256 > $O/r02_headline.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:bog_gather -s 2 -c 1 -o $O/r02_bog_c3 -f python scripts/traffic_headline.py 3 64 > $O/r02_bog_c3.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:bog_gather -s 2 -c 1 -o $O/r02_bog_c4 -f python scripts/traffic_headline.py 4 70 > $O/r02_bog_c4.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nh_q1|bog_gather" -s 4 -c 2 -o $O/r02_staged_c5 -f python scripts/traffic_headline.py 5 128 > $O/r02_staged_c5.out 2>&1
ls -la $O | tail -30
