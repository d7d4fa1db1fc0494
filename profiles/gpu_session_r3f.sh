set -x
for cfg in "0 0" "200 0" "0 128" "0 112" "160 128"; do
set -- $cfg
for i in 7 1 4 3; do UAVNET_GEMM_BUDGET_KB=$1 UAVNET_GEMM_BN_MAX=$2 python profiles/gemm_bench.py --only $i 2>/dev/null | grep "^{'case" | sed "s/^/budget=$1 bn=$2 /"; done
done
