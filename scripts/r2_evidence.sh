#!/bin/bash
# Round-2 evidence on one B200: GPU tests, the bench line (both arms), ncu launch list of the bench command, DRAM traffic of one
# step, `ncu --set full` on the dominant kernel, DRAM rows of the HBM-bound kernels.  Everything lands in gpurun_out/.
T=${1:-r2}
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log; tail -3 $O/${T}_pytest.log
( time python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err ) 2> $O/${T}_bench_time.txt; echo "bench rc=$?"
python bench.py --impl reference --steps 8 --warmup 2 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "reference rc=$?"
python bench.py --steps 2 --warmup 3 > $O/${T}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${T}_launches_bench.csv python bench.py --steps 2 --warmup 3 > $O/${T}_ncu_bench.log 2>&1
python scripts/profile_batch.py 256 > $O/${T}_plain_batch.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1300 --csv --log-file $O/${T}_traffic_launches.csv python scripts/profile_batch.py 256 > $O/${T}_ncu_traffic.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dmma_ws_kernel -s 70 -c 6 -o $O/${T}_full_dmma_ws -f python scripts/profile_batch.py 256 > $O/${T}_ncu_full.log 2>&1
python scripts/profile_elementwise.py > $O/${T}_plain_elem.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'cov_finish|gram_diag|alpha_kernel|wk_kernel|predict_finish|adjust_kernel|nlml_finish|rs_reduce|fill_aug' --csv --log-file $O/${T}_elementwise.csv python scripts/profile_elementwise.py > $O/${T}_ncu_elem.log 2>&1
tail -2 $O/${T}_plain_elem.log; ls -la $O | grep ${T}_ | head -30
