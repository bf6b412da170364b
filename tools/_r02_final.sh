# final round-2 measurements on one B200: bench line, launch list, ncu --set full captures
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 300 gpurun_out/r02_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_ncu_launches.log 2>&1
bash tools/ncu_capture.sh r02_picard_C4 k_krotov_picard 4 --no-configs
bash tools/ncu_capture.sh r02_dpsweep_C2 k_dp_sweep 4 --no-configs --workload C2
bash tools/ncu_capture.sh r02_dpsweep_C5 k_dp_sweep 4 --no-configs --workload C5
bash tools/ncu_capture.sh r02_dpbuild_C5 k_dp_build 0 --no-configs --workload C5
