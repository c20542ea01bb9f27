#!/bin/bash
# bench.py lines for the BASELINE shapes other than the headline, N = 1 (GPU box; results into gpurun_out/)
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py --workload reddit --K 256 --unvalued --steps 20 --warmup 5 --no-cpu > $O/bench_reddit_K256.json 2> $O/bench_reddit.err
timeout 600 python bench.py --workload reddit --K 128 --steps 20 --warmup 5 --no-cpu > $O/bench_reddit_K128.json 2>> $O/bench_reddit.err
timeout 600 python bench.py --workload products --K 128 --steps 20 --warmup 5 --no-cpu > $O/bench_products_K128.json 2> $O/bench_products.err
timeout 600 python bench.py --workload products --K 32 --steps 20 --warmup 5 --no-cpu > $O/bench_products_K32.json 2>> $O/bench_products.err
timeout 900 python bench.py --workload rmat --K 128 --steps 10 --warmup 3 --no-cpu > $O/bench_rmat_K128.json 2> $O/bench_rmat.err
