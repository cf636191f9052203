cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r02_gputests_n.log 2>&1
tail -4 gpurun_out/r02_gputests_n.log
timeout 400 python profiles/quick.py n 2>&1 | grep -v "^\[lmono" | tail -60
LMONO_ODOM_NN=brute timeout 400 python profiles/quick.py n_brute sweep 2>&1 | grep -v "^\[lmono" | tail -30
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --legs fused_sweep,c2_odometry,c4_fused_batch > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_e.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['single_sequence']['ms_per_registration'])
for k in ('fused_sweep','c2_odometry','c4_fused_batch'):
    v=d.get(k,{}); print(k, {a:v[a] for a in v if a in ('value','unit','ms_per_sweep','error','ms_per_batch_step','stage_ms','single_sequence_value')})
P
