cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python profiles/quick.py q sweep 2>&1 | grep -v "^\[lmono" | tail -45
LMONO_RF_INPLACE=0 timeout 400 python profiles/quick.py q0 sweep 2>&1 | grep -v "^\[lmono" | tail -45
timeout 300 python -m pytest tests/test_gpu_scanreg.py tests/test_gpu_pipeline.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -5
