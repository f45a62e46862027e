for i in 1 2 3; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('clocks on ', d['roofline']['kernel_ms_each_rank0'], d['clocks']['samples'])"
BENCH_NO_CLOCKS=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('clocks off', d['roofline']['kernel_ms_each_rank0'])"
done
