timeout 60 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for w in ${WORKLOADS:-C4 C1 C2 C3 C5}; do
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --workload $w 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', 'it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'fw_ms', round(d['roofline']['fw_sweep_ms'],3), 'bw_ms', round(d['roofline']['bw_sweep_ms'],3), 'ns/step fw', round(d['roofline']['ns_per_time_step_fw']), 'bw', round(d['roofline']['ns_per_time_step_bw']))"
done
