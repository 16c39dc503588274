#!/bin/bash
# One gpurun call's worth of checks on a 1-GPU box: parity suite, manyTarg kernel timings, configs 3-5, one ncu capture
# of the manyTarg kernels, the contract bench line. Everything lands in gpurun_out/ (merged back by gpurun).
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_session.sh [tag] [steps...]'   steps default: test mt configs ncu bench
#        gpurun --gpus N ... 'NP=N bash tools/gpu_session.sh tag configs_mp bench_mp'   (torchrun, one rank per GPU)
set -u
TAG=${1:-s}
shift || true
STEPS=${*:-test mt configs ncu bench}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
for step in $STEPS; do
  case $step in
    test)
      timeout 900 python -m pytest tests -m gpu -q --durations=8 > $OUT/${TAG}_pytest.log 2>&1
      echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
      tail -n 15 $OUT/${TAG}_pytest.log ;;
    mt)
      timeout 300 python tools/bench_manytarg.py 30 > $OUT/${TAG}_manytarg.jsonl 2> $OUT/${TAG}_manytarg.err
      cat $OUT/${TAG}_manytarg.jsonl; tail -n 5 $OUT/${TAG}_manytarg.err ;;
    configs)
      timeout 400 python tools/bench_configs.py > $OUT/${TAG}_configs_n1.jsonl 2> $OUT/${TAG}_configs.err
      cat $OUT/${TAG}_configs_n1.jsonl; tail -n 5 $OUT/${TAG}_configs.err ;;
    micro)
      timeout 300 python tools/microbench.py 30 > $OUT/${TAG}_microbench.jsonl 2> $OUT/${TAG}_microbench.err
      cat $OUT/${TAG}_microbench.jsonl; tail -n 5 $OUT/${TAG}_microbench.err ;;
    ncu)
      timeout 400 ncu --set full --clock-control none --import-source on -k regex:manyTarg -c 8 -f -o $OUT/${TAG}_manytarg \
          python tools/prof_manytarg.py 26 3,4,5 > $OUT/${TAG}_ncu.log 2>&1
      ncu -i $OUT/${TAG}_manytarg.ncu-rep --page raw --csv > $OUT/${TAG}_manytarg_raw.csv 2>> $OUT/${TAG}_ncu.log
      tail -n 3 $OUT/${TAG}_ncu.log ;;
    mt5)
      timeout 120 python -m pytest tests -m gpu -q -k many_targ > $OUT/${TAG}_pytest_mt.log 2>&1; tail -n 4 $OUT/${TAG}_pytest_mt.log
      MT_ONLY=${MT_ONLY:-3,4,5} timeout 300 python tools/bench_manytarg.py 30 > $OUT/${TAG}_manytarg.jsonl 2> $OUT/${TAG}_manytarg.err
      cat $OUT/${TAG}_manytarg.jsonl; tail -n 5 $OUT/${TAG}_manytarg.err ;;
    ncu5)
      for t in ${NCU_T:-5}; do for pl in ${NCU_PLACEMENTS:-mid low}; do
        timeout 300 ncu --set full --clock-control none --import-source on -k regex:manyTarg -c 2 -f -o $OUT/${TAG}_mt${t}_${pl} \
            python tools/prof_manytarg.py 26 $t $pl > $OUT/${TAG}_ncu_mt${t}_${pl}.log 2>&1
        ncu -i $OUT/${TAG}_mt${t}_${pl}.ncu-rep --page raw --csv > $OUT/${TAG}_mt${t}_${pl}_raw.csv 2>> $OUT/${TAG}_ncu_mt${t}_${pl}.log
        tail -n 2 $OUT/${TAG}_ncu_mt${t}_${pl}.log
      done; done ;;
    launches)
      timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_n1_30q.csv \
          python bench.py --steps 2 --warmup 1 --qubits 30 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1
      tail -n 2 $OUT/${TAG}_launches.log ;;
    bench)
      timeout 900 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench.err
      cat $OUT/${TAG}_bench_n1.json; tail -n 5 $OUT/${TAG}_bench.err ;;
    configs_mp)
      timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NP:-2} --master-addr 127.0.0.1 --master-port 29531 \
          tools/bench_configs.py --reps 2 > $OUT/${TAG}_configs_n${NP:-2}.jsonl 2> $OUT/${TAG}_configs_n${NP:-2}.err
      cat $OUT/${TAG}_configs_n${NP:-2}.jsonl; tail -n 3 $OUT/${TAG}_configs_n${NP:-2}.err ;;
    bench_mp)
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NP:-2} --master-addr 127.0.0.1 --master-port 29532 \
          bench.py --gpus ${NP:-2} --steps 2 --warmup 3 > $OUT/${TAG}_bench_n${NP:-2}.json 2> $OUT/${TAG}_bench_n${NP:-2}.err
      cat $OUT/${TAG}_bench_n${NP:-2}.json; tail -n 3 $OUT/${TAG}_bench_n${NP:-2}.err ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
      tail -n 3 $OUT/${TAG}_smoke.log ;;
  esac
done
