# A/B of the single-launch BatchNorm grid heuristics (bench.py ms/step, device-timed and e2e)
for cfg in ${AB_CFGS:-"0 4 148" "1 8 64" "1 8 96" "1 8 80" "1 8 48" "1 12 64" "1 6 64" "1 6 96"}; do
set -- $cfg
echo "FUSED=$1 VPT=$2 MAXG=$3: $(B200SCN_BN_FUSED=$1 B200SCN_BN_FUSED_VPT=$2 B200SCN_BN_FUSED_MAXG=$3 timeout 200 python bench.py --no-cpu-baseline --no-roofline 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"])')"
done
