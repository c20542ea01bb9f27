#!/bin/bash
# One bounded GPU session (run through gpurun from the repo root): every step under its own timeout, most
# important first, everything into gpurun_out/.  Usage: bash scripts/gpu_session.sh [steps...]
#   steps (round 1): subwarp sweep ksweep cli opreddit suite seq seqsuite oddk rows auto64 bench smoke sanitize launches ncu128 ncu
#   steps (round 2): uniform regreddit newtests l2sweep bulksweep l2ncu sanitize_bulk sanitize_rg sanitize_odd probel2 ncurg hotsweep
#                    gcnlayer benchdrv multi (multi: under gpurun --gpus N -- NCCL test, bench.py at N GPUs with its R-MAT record)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
mkdir -p $O
STEPS=${*:-subwarp sweep suite auto64 bench smoke sanitize ncu}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/gpu.txt 2>&1
note() { echo "$(date +%T) $*" | tee -a $O/status.txt; }
note "session start: $STEPS"
for step in $STEPS; do
  case $step in
    subwarp)  # first contact with the sub-warp walker (GESPMM_VARIANT=2 inside the test)
      timeout 300 python -m pytest tests/test_spmm_gpu.py -q -m gpu -k "subwarp" > $O/t_subwarp.log 2>&1; note "subwarp rc=$?" ;;
    sweep)    # ring vs sub-warp walker, K <= 64, four graph shapes
      timeout 420 python scripts/sweep_narrow.py > $O/sweep_narrow.txt 2> $O/sweep_narrow.err; note "sweep rc=$?" ;;
    ksweep)   # the shipped configuration next to the reference kernel, every BASELINE shape
      timeout 600 python scripts/sweep_narrow.py --workloads products --Ks 32,64,128,256,512 --variants -1 --tasks 0 --valued 1 --ref > $O/ksweep.txt 2> $O/ksweep.err
      timeout 600 python scripts/sweep_narrow.py --workloads reddit,citpatents,rmat --rmat-scale 1.0 --Ks 128,256 --variants -1 --tasks 0 --valued 1 --ref >> $O/ksweep.txt 2>> $O/ksweep.err
      note "ksweep rc=$?" ;;
    cli)      # BASELINE.json configs[1] through the CLI: cit-Patents shape as a .mtx, reference kernel as the baseline cell
      timeout 600 python scripts/cli_synthetic.py --workload citpatents --validate > $O/cli_citpatents.txt 2> $O/cli_citpatents.err; note "cli rc=$?" ;;
    opreddit) # BASELINE.json configs[2]: Reddit shape, K = 256, SPMMFunction forward + backward next to the reference extension
      timeout 600 python scripts/op_reddit.py > $O/op_reddit.json 2> $O/op_reddit.err; note "opreddit rc=$?" ;;
    suite)    # the whole GPU suite with default settings
      timeout 900 python -m pytest tests -q -m gpu > $O/t_default.log 2>&1; note "suite rc=$?" ;;
    seq)      # the SpMM suite with the sequential ring walker forced for every K
      GESPMM_VARIANT=0 timeout 600 python -m pytest tests/test_spmm_gpu.py -q -m gpu > $O/t_seq.log 2>&1; note "seq rc=$?" ;;
    seqsuite) # the SpMM suite with GESPMM_SEQUENTIAL=1 (row-parallel walker for K <= 64, ring walker above)
      GESPMM_SEQUENTIAL=1 timeout 600 python -m pytest tests/test_spmm_gpu.py -q -m gpu > $O/t_sequential.log 2>&1; note "seqsuite rc=$?" ;;
    oddk)     # class-count widths (K % 4 != 0: the scalar walker) next to the reference's kernels for K < 64
      timeout 600 python scripts/sweep_narrow.py --workloads products,citpatents --Ks 3,7,41,47 --variants -1 --tasks 0 --valued 1,0 --ref > $O/sweep_oddk.txt 2> $O/sweep_oddk.err
      note "oddk rc=$?" ;;
    uniform)  # the uniformly-random variant of the cit-Patents shape: time + DRAM bytes per launch
      timeout 300 python bench.py --workload citpatents_uniform --no-cpu --no-e2e --steps 50 > $O/bench_uniform.json 2> $O/bench_uniform.err; note "uniform rc=$?"
      for wl in citpatents citpatents_uniform; do
        timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum \
            --clock-control none -k regex:spmm_ --launch-skip 6 -c 2 --csv --log-file $O/ncu_dram_$wl.csv \
            python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-e2e --no-ref-kernel > $O/ncu_dram_$wl.log 2>&1
        note "ncu dram $wl rc=$?"
      done ;;
    regreddit) # register-staged walker (GESPMM_VARIANT=1) next to the ring walker where B is L2-resident
      timeout 400 python scripts/sweep_narrow.py --workloads reddit,products --Ks 128,256 --variants=-1,1 --tasks 0 --valued 0,1 > $O/sweep_regreddit.txt 2> $O/sweep_regreddit.err; note "regreddit rc=$?" ;;
    newtests) # round 2: per-call options, fused scaling, row-group kernel for tiny odd K, L2 steering
      timeout 900 python -m pytest tests/test_spmm_gpu.py -q -m gpu -x -k "bulk_walker or sequential_order or fused or longest_row or l2_priority or k_sweep or gcnconv or degenerate or max_reduce or c_abi_strides" > $O/t_new.log 2>&1; note "newtests rc=$?" ;;
    l2sweep)  # ring (cp.async) vs bulk (TMA) walker, and L2 eviction-priority steering of the bulk gathers, cit-Patents shape
      timeout 600 python scripts/sweep_l2.py > $O/sweep_l2.txt 2> $O/sweep_l2.err; note "l2sweep rc=$?"
      timeout 300 python scripts/sweep_l2.py --workloads citpatents --walkers 0,5 --policies 0 --tasks 64,96,128,192 --pads 0,2048,4096 >> $O/sweep_l2.txt 2>> $O/sweep_l2.err; note "l2sweep2 rc=$?" ;;
    bulksweep) # ring vs bulk walker on the other shapes (Reddit / ogbn-products / R-MAT at half size), K = 128, 256
      for wl in reddit products rmat; do
        timeout 300 python scripts/sweep_l2.py --workloads $wl --K 128 --walkers 0,5 --policies 0 --iters 5 --batches 3 >> $O/sweep_bulk.txt 2>> $O/sweep_bulk.err
        timeout 300 python scripts/sweep_l2.py --workloads $wl --K 256 --walkers 0,5 --policies 0 --iters 5 --batches 3 >> $O/sweep_bulk.txt 2>> $O/sweep_bulk.err
      done; note "bulksweep rc=$?" ;;
    l2ncu)    # DRAM bytes per launch of chosen configurations: L2NCU="walker,policy,window,task,pad ..." (default below)
      for cfg in ${L2NCU:-0,0,0,0,0 5,0,0,0,0 5,22,131072,0,0 5,20,131072,0,0}; do
        timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
            --clock-control none -k regex:spmm_flat --launch-skip 3 -c 2 --csv --log-file $O/ncu_l2_$cfg.csv \
            python scripts/sweep_l2.py --workloads citpatents --one $cfg > $O/ncu_l2_$cfg.log 2>&1
        note "l2ncu $cfg rc=$?"
      done ;;
    sanitize_bulk) # compute-sanitizer over the bulk walker through the CLI
      for tool in memcheck racecheck; do
        echo "== $tool (GESPMM_VARIANT=5, K=128,200)" >> $O/sanitize_bulk.txt
        GESPMM_VARIANT=5 timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 ge-spmm_b200/bin/spmm_test $O/sanitize.mtx 0 \
            --K 128,200 --iters 2 --validate --out $O/sanitize.csv 2>&1 | grep -E "SUMMARY|validate|WA|Error|error" | head -12 >> $O/sanitize_bulk.txt
      done
      note "sanitize_bulk done" ;;
    sanitize_rg) # compute-sanitizer over the row-group kernel (K = 3, 7, 13) through the CLI
      python - > $O/sanitize_gen.log 2>&1 <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import __graft_entry__ as e
e.load_package()
from gespmm_b200 import graphs
rng = np.random.default_rng(0)
M = 3000
deg = rng.integers(0, 9, M); deg[rng.random(M) < 0.3] = 0
deg[[5, 700, 701, 2999]] = [40000, 5000, 4097, 9000]
deg[100:140] = rng.integers(30, 300, 40)
rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
colind = rng.integers(0, M, rowptr[-1]).astype(np.int32)
graphs.write_mtx('gpurun_out/sanitize.mtx', rowptr, colind)
PY
      for tool in memcheck racecheck synccheck; do
        echo "== $tool (K=3,7,13,128)" >> $O/sanitize_rowgroup.txt
        timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 ge-spmm_b200/bin/spmm_test $O/sanitize.mtx 0 \
            --K 3,7,13,128 --iters 2 --validate --out $O/sanitize.csv 2>&1 | grep -E "SUMMARY|validate|WA|Error|error" | head -12 >> $O/sanitize_rowgroup.txt
      done
      note "sanitize_rg done" ;;
    probel2)  # which hinted instruction does the hardware refuse?  store only / gathers only / both
      for pol in 16 32 1 4 5 21; do
        timeout 120 python scripts/probe_l2.py $pol >> $O/probe_l2.txt 2>&1; echo "policy $pol rc=$?" >> $O/probe_l2.txt
      done
      timeout 200 compute-sanitizer --tool memcheck python scripts/probe_l2.py 21 2>&1 | grep -v "^$" | head -40 > $O/probe_l2_sanitizer.txt
      note "probel2 done" ;;
    ncurg)    # full capture of the narrow-B kernels at K = 3: sub-warp walker on 4-byte slices (default) and row-group kernel
      for wl in products citpatents; do
        for v in -1 4; do
          timeout 400 ncu --set full --clock-control none --import-source on -k regex:"spmm_rowgroup|spmm_flat" --launch-skip 3 -c 1 \
              -f -o /tmp/ncu_rg python scripts/sweep.py --workload $wl --K 3 --variants=$v --iters 1 --unvalued > $O/ncu_rg_${wl}_v$v.log 2>&1
          note "ncurg $wl v$v rc=$?"
          ncu -i /tmp/ncu_rg.ncu-rep --page raw --csv > $O/ncu_rg_${wl}_v${v}_raw.csv 2>/dev/null
          ncu -i /tmp/ncu_rg.ncu-rep --page source --csv > $O/ncu_rg_${wl}_v${v}_source.csv 2>/dev/null
          rm -f /tmp/ncu_rg.ncu-rep
        done
      done ;;
    multi)    # N GPUs (gpurun --gpus N): NCCL sharding test, then the bench line with its R-MAT record
      NG=$(nvidia-smi -L | wc -l)
      timeout 600 python -m pytest tests/test_sharding_nccl_gpu.py -x -q -m gpu > $O/t_nccl.log 2>&1; note "nccl test rc=$?"
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
          bench.py --gpus $NG --steps 20 --warmup 5 > $O/bench_n$NG.json 2> $O/bench_n$NG.err; note "bench n=$NG rc=$?"
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 \
          bench.py --impl reference --gpus $NG --steps 20 --warmup 5 > $O/bench_n${NG}_reference.json 2> $O/bench_n${NG}_reference.err; note "bench reference n=$NG rc=$?" ;;
    hotsweep) # hot-column pinning in the L2 (bulk walker + gespmm_opts.hot_columns) on R-MAT 10M/200M
      timeout 600 python scripts/sweep_hot.py > $O/sweep_hot.txt 2> $O/sweep_hot.err; note "hotsweep rc=$?"
      for cfg in ${HOTNCU:-0,0,0 5,22,100000 5,22,150000}; do
        timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
            --clock-control none -k regex:spmm_flat --launch-skip 3 -c 1 --csv --log-file $O/ncu_hot_$cfg.csv \
            python scripts/sweep_hot.py --one $cfg > $O/ncu_hot_$cfg.log 2>&1
        note "hotncu $cfg rc=$?"
      done ;;
    sanitize_odd) # compute-sanitizer: 4-byte-slice ring walker (CLI, K = 41, 100) and the padded route + fused epilogue (operator)
      python - > $O/sanitize_gen.log 2>&1 <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import __graft_entry__ as e
e.load_package()
from gespmm_b200 import graphs
rng = np.random.default_rng(0)
M = 3000
deg = rng.integers(0, 9, M); deg[rng.random(M) < 0.3] = 0
deg[[5, 700, 701, 2999]] = [40000, 5000, 4097, 9000]
deg[100:140] = rng.integers(30, 300, 40)
rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
colind = rng.integers(0, M, rowptr[-1]).astype(np.int32)
graphs.write_mtx('gpurun_out/sanitize.mtx', rowptr, colind)
PY
      for tool in memcheck racecheck; do
        echo "== $tool, CLI (bare C ABI: 4-byte-slice ring walker), K=41,47,127" >> $O/sanitize_odd.txt
        timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 ge-spmm_b200/bin/spmm_test $O/sanitize.mtx 0 \
            --K 41,47,127 --iters 2 --validate --out $O/sanitize.csv 2>&1 | grep -E "SUMMARY|validate|WA|Error|error" | head -12 >> $O/sanitize_odd.txt
        echo "== $tool, operator (padded route, sequential + fused epilogue), K=41" >> $O/sanitize_odd.txt
        timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_spmm_gpu.py -q -m gpu -x \
            -k "widths_that_are_not and 41 or fused_scales and 41" 2>&1 | grep -E "SUMMARY|passed|failed|Error|error" | head -12 >> $O/sanitize_odd.txt
      done
      note "sanitize_odd done" ;;
    gcnlayer) # fused vs unfused aggregation half of a GCN layer, and the training loop per epoch
      timeout 600 python scripts/gcn_layer_bench.py > $O/gcn_layer.txt 2> $O/gcn_layer.err; note "gcnlayer rc=$?" ;;
    fuzz)     # randomised differential test of the C ABI against the oracle
      timeout 900 python scripts/fuzz_gpu.py --cases ${FUZZ_CASES:-800} > $O/fuzz.txt 2>&1; note "fuzz rc=$?" ;;
    rows)     # sub-warp (2) vs row-parallel (4) narrow walkers on every shape
      timeout 600 python scripts/sweep_narrow.py --variants 2,4 --tasks 0 > $O/sweep_v24.txt 2> $O/sweep_v24.err; note "rows rc=$?" ;;
    auto64)   # the SpMM suite with the sub-warp walker chosen automatically for K <= 64
      GESPMM_SUBWARP_MAX_K=64 timeout 600 python -m pytest tests/test_spmm_gpu.py -q -m gpu > $O/t_auto64.log 2>&1; note "auto64 rc=$?" ;;
    bench)
      timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; note "bench rc=$?"
      timeout 300 python bench.py --impl reference > $O/bench_n1_reference.json 2> $O/bench_n1_reference.err; note "bench reference rc=$?" ;;
    benchdrv) # the driver's own invocation at N = 1
      timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1_drv.json 2> $O/bench_n1_drv.err; note "benchdrv rc=$?"
      timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_n1_drv_reference.json 2> $O/bench_n1_drv_reference.err; note "benchdrv reference rc=$?" ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; note "smoke rc=$?" ;;
    sanitize) # compute-sanitizer over the sub-warp kernels through the CLI
      python - > $O/sanitize_gen.log 2>&1 <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import __graft_entry__ as e
e.load_package()
from gespmm_b200 import graphs
rng = np.random.default_rng(0)
M = 3000
deg = rng.integers(0, 9, M); deg[rng.random(M) < 0.3] = 0
deg[[5, 700, 701, 2999]] = [40000, 5000, 4097, 9000]
rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
colind = rng.integers(0, M, rowptr[-1]).astype(np.int32)
graphs.write_mtx('gpurun_out/sanitize.mtx', rowptr, colind)
PY
      for tool in memcheck racecheck; do
        echo "== $tool (GESPMM_VARIANT=2, K=16,32,64)" >> $O/sanitize_subwarp.txt
        GESPMM_VARIANT=2 timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 ge-spmm_b200/bin/spmm_test $O/sanitize.mtx 0 \
            --K 16,32,64 --iters 2 --validate --out $O/sanitize.csv 2>&1 | grep -E "SUMMARY|validate|WA|Error|error" | head -12 >> $O/sanitize_subwarp.txt
      done
      note "sanitize done" ;;
    launches) # launch list of the bench command (per-launch times are cold-cache and serialised: compare shares)
      timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
          python bench.py --steps 3 --warmup 3 --no-cpu > $O/launches_bench.log 2>&1; note "launches rc=$?"
      python scripts/summarize_launches.py $O/launches.csv > $O/launches_summary.txt 2>&1 ;;
    ncu128)   # one full capture of the dominant kernel of the bench command (cit-Patents shape, K = 128); exported as CSV pages
      timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmm_flat_kernel --launch-skip 3 -c 1 \
          -f -o /tmp/ncu_citpatents_K128 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-ref-kernel > $O/ncu128.log 2>&1
      note "ncu128 rc=$?"
      ncu -i /tmp/ncu_citpatents_K128.ncu-rep --page raw --csv > $O/ncu_citpatents_K128_raw.csv 2>/dev/null
      ncu -i /tmp/ncu_citpatents_K128.ncu-rep --page source --csv > $O/ncu_citpatents_K128_source.csv 2>/dev/null
      rm -f /tmp/ncu_citpatents_K128.ncu-rep ;;
    ncu)      # one full capture of each walker's kernel A on the ogbn-products shape, K = 32
      for v in 2 0; do
        GESPMM_VARIANT=$v timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmm_flat_kernel --launch-skip 3 -c 1 \
            -f -o $O/ncu_products_K32_v$v python scripts/sweep.py --workload products --K 32 --variants $v --iters 1 > $O/ncu_v$v.log 2>&1
        note "ncu v$v rc=$?"
      done ;;
  esac
done
note "session end"
