for m in 0 1 2; do
  echo "== B200SP_TCG_WAIT=$m"
  B200SP_TCG_WAIT=$m python tools/gemm_bench.py --reps 5 --shapes 602112,16,32 150528,144,24 9408,64,192 9408,384,64 9408,96,576 2352,160,960 2352,1024,1280
done
