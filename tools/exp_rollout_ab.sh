cd /root/repo
run() { echo "== $*"; env "${@:7}" timeout 300 python tools/run_rollout.py $1 $2 $3 $4 $5 $6 2>&1 | tail -1 | cut -c1-80; }
for cfg in "8 4 65536 20 30 8" "10 4 65536 20 30 8" "10 8 65536 20 30 8" "12 6 65536 20 30 8" "14 7 65536 20 20 8" "8 8 65536 20 30 8" "10 5 16384 20 40 10" "10 5 32768 20 40 10"; do
  for v in "NOTIMING=1" "NOTIMING=1 RBG_PERSIST_ENV_WARPS=4 RBG_GEN_WARPS=2 RBG_ROLLOUT_CTAS=8" "NOTIMING=1 RBG_ROLLOUT_CTAS=8"; do run $cfg $v; done
done
