cd /root/repo
run() { echo "== $*"; env "${@:7}" timeout 300 python tools/run_rollout.py $1 $2 $3 $4 $5 $6 2>&1 | tail -1; }
for cfg in "32 16 8192 20 12 6" "20 10 32768 20 20 6" "16 8 65536 20 20 8" "24 12 16384 20 12 6" "40 32 4096 20 8 4" "10 5 8192 20 40 10" "10 5 128 20 40 10"; do
  for v in "NOTIMING=1" "NOTIMING=1 RBG_GEN_WARPS=1" "NOTIMING=1 RBG_ROLLOUT_IMPL=legacy"; do run $cfg $v; done
done
