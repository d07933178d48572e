timeout 900 python -m pytest tests/test_taxi_gpu.py -m gpu -x -q --timeout=400 -k "fused or stat" 2>&1 | tail -3
for v in libgpt_b200 variant_t7 variant_t4; do for sh in 2128 4128 1128; do
  echo "$v shape=$sh"
  GPT_TAXI_SHAPE=$sh GPT_B200_LIB=$PWD/gym-po-taxi_b200/lib/$v.so timeout 200 python bench.py --quick --workload taxi --steps 2000 --warmup 20 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']/1e9,1), round(d['ms_per_step']*1e3,2), r['alg_bytes_per_env_step'], round(r['frac'],3), d['gpu_launches'])"
done; done
