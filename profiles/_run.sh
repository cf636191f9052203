cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r02_gputests_u.log 2>&1
tail -4 gpurun_out/r02_gputests_u.log
timeout 400 python profiles/quick.py u single,batch,lm,sweep 2>&1 | grep -v "^\[lmono" | head -30
timeout 1200 python bench.py > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err; tail -c 1500 gpurun_out/r02_bench_f.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_f.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'single',d['single_sequence'],'more',d.get('more_sequences',{}).get('value'))
print('roofline', d['roofline']['kernel'][:60], d['roofline']['frac'], d['roofline']['whole_step'])
for k in ('fused_sweep','c2_odometry','colour_frame','c1_cpu_pipeline','c4_fused_batch'):
    v=d.get(k,{}); print(k, {a:v[a] for a in v if a in ('value','unit','ms_per_sweep','sweeps_per_s','error','ms_per_batch_step','gpu_ms_per_sweep_device','cpu_ms_per_sweep','ms_per_sweep_kernels','ms_per_frame_kernels')})
print('cpu', d.get('cpu_baseline'))
P
