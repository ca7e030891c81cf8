# A/B of programmatic dependent launch (bench.py ms/step, device-timed and e2e)
for v in 0 1 0 1; do
echo "PDL=$v: $(B200SCN_PDL=$v timeout 200 python bench.py --no-cpu-baseline --no-roofline 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"])')"
done
