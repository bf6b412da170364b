timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload C4sat 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})); print(json.dumps(d['roofline'])); print(json.dumps(d['e2e']))"
tools/ncu_capture.sh bw_r1_e k_prop_spec 2
tools/ncu_capture.sh fwsat_r1_e k_fwupd_spec 2 --workload C4sat
tools/ncu_capture.sh bwsat_r1_e k_prop_spec 2 --workload C4sat
