cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "rollout or interleave or reference_env_composition or fixtures_on_gpu or mixed_api or dataset" 2>&1 | tail -3
run() { echo "== $*"; env "$@" timeout 300 python tools/run_rollout.py 10 5 65536 20 40 10 2>&1 | tail -1; }
run NOTIMING=1
run NOTIMING=1 RBG_ROLLOUT_OVERLAP=0
run NOTIMING=1 RBG_GEN_WARPS=1
