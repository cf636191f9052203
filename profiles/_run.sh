cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$GRAFT_REPO_ROOT/lmono_b200/csrc/variants
for m in 1 2 3 0; do
  echo "=== variant $m"
  LMONO_SO=$V/liblmono_pdl$m.so timeout 300 python -m pytest tests/test_golden.py tests/test_gpu_mapping.py -m gpu -x -q 2>&1 | tail -3
  LMONO_SO=$V/liblmono_pdl$m.so timeout 300 python profiles/quick.py pdlv$m single,batch 2>&1 | grep -v "^\[lmono" | tail -4
done
echo "=== variant 0 nograph"
LMONO_NO_GRAPH=1 LMONO_SO=$V/liblmono_pdl0.so timeout 300 python -m pytest tests/test_golden.py tests/test_gpu_mapping.py -m gpu -x -q 2>&1 | tail -3
