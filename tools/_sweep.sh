for cfg in "0 0" "16 0" "16 60" "16 50" "12 50" "12 40" "20 75" "20 0" "8 30"; do
  set -- $cfg
  if [ "$2" != "0" ]; then export SEEKR_B200_COUNT_CARVEOUT=$2; else unset SEEKR_B200_COUNT_CARVEOUT; fi
  export SEEKR_B200_COUNT_CTAS_PER_SM=$1
  echo "== ctas/SM $1 carveout $2"
  timeout 100 python tools/microbench_count.py --ks 6 --only-count 2>&1 | grep -E "count raw|fused -mean /std \+min|\+post"
done
