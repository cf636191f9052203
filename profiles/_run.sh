cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_odom.py tests/test_gpu_pipeline.py tests/test_gpu_consumer.py -m gpu -x -q 2>&1 | tail -5
timeout 400 python profiles/quick.py t sweep 2>&1 | grep -v "^\[lmono" | grep "fused sweep\|k_scan_ring\|k_odom_nn\|mapping ms"
