timeout 900 python -m pytest tests/test_taxi_gpu.py tests/test_wrappers_gpu.py -m gpu -x -q --timeout=400 2>&1 | tail -4
for w in taxi taxi_hansen; do for f in 0 1; do
  if [ $f = 1 ]; then export GPT_NO_FUSED_STEPS=1; else unset GPT_NO_FUSED_STEPS; fi
  echo "$w nofuse=$f"
  timeout 200 python bench.py --quick --workload $w --steps 2000 --warmup 20 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']/1e9,1), round(d['ms_per_step']*1e3,2), r['alg_bytes_per_env_step'], round(r['frac'],3), d['gpu_launches'])"
done; done
