for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_r02_${n}gpu.json 2> gpurun_out/bench_r02_${n}gpu.err
  cat gpurun_out/bench_r02_${n}gpu.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['identify_seconds'], d['e2e']['solver'])"
done
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_r02_1gpu.json 2>/dev/null; cat gpurun_out/bench_r02_1gpu.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['identify_seconds'], d['e2e']['solver'])"
