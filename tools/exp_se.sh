cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "seedext or reference_runs or random_shapes" 2>&1 | tail -3
for cfg in "1000 0 1" "4 3 296" "2 8 296" "3 6 592" "2 12 296" "3 9 296"; do set -- $cfg; echo "== round sweeps $1 rounds $2 dense $3"; RBG_SE_ROUND_SWEEPS=$1 RBG_SE_ROUNDS=$2 RBG_SE_DENSE_WARPS=$3 python tools/run_seedext.py 5 2>&1 | grep "seedext"; done
