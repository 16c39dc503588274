#!/bin/bash
# Round-2 GPU sessions (one gpurun call each; everything lands in gpurun_out/ and the keepers are copied to profiles/):
#   gpurun --timeout 1500 -- 'bash tools/gpu_session_r02.sh one TAG'            1 GPU: selected parity tests, dense-gate timings, ncu source-level
#                                                                              capture of the t = 5 kernel, compute-sanitizer passes, the bench line
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_session_r02.sh link TAG'   2 GPUs: NVLink sweep over DFSA_REMOTE_INFLIGHT
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_session_r02.sh eight TAG'  8 GPUs: multi-rank parity on real NVLink, bench + per-gate times
set -u
MODE=${1:-one}
TAG=${2:-r02}
OUT=gpurun_out
mkdir -p $OUT
torchrun_n() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
case $MODE in
  one)
    timeout 600 python -m pytest tests -m gpu -q --durations=8 -k "many_targ or kraus or expec or partial or single_rank or dm_ops or comparator or 13_qubits" > $OUT/${TAG}_pytest_sel.log 2>&1
    tail -n 12 $OUT/${TAG}_pytest_sel.log
    MT_ONLY=${MT_ONLY:-5,6,7,8} timeout 400 python tools/bench_manytarg.py 30 > $OUT/${TAG}_manytarg_30q.jsonl 2> $OUT/${TAG}_manytarg.err
    cat $OUT/${TAG}_manytarg_30q.jsonl; tail -n 3 $OUT/${TAG}_manytarg.err
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:manyTargSpec -c 1 -f -o $OUT/${TAG}_mt5_mixed python tools/prof_manytarg.py 26 5 mixed > $OUT/${TAG}_ncu_mt5.log 2>&1
    ncu -i $OUT/${TAG}_mt5_mixed.ncu-rep --page raw --csv > $OUT/${TAG}_mt5_mixed_raw.csv 2>> $OUT/${TAG}_ncu_mt5.log
    ncu -i $OUT/${TAG}_mt5_mixed.ncu-rep --page source --csv > $OUT/${TAG}_mt5_mixed_source.csv 2>> $OUT/${TAG}_ncu_mt5.log
    tail -n 2 $OUT/${TAG}_ncu_mt5.log
    timeout 300 ncu --set full --clock-control none -k regex:manyTargGemm -c 1 -f -o $OUT/${TAG}_gemm8 python tools/prof_manytarg.py 26 8 mixed8 > $OUT/${TAG}_ncu_gemm.log 2>&1
    ncu -i $OUT/${TAG}_gemm8.ncu-rep --page raw --csv > $OUT/${TAG}_gemm8_raw.csv 2>> $OUT/${TAG}_ncu_gemm.log
    tail -n 2 $OUT/${TAG}_ncu_gemm.log
    for tool in memcheck racecheck synccheck; do
      timeout 400 compute-sanitizer --tool $tool python tools/sanitize_cases.py > $OUT/${TAG}_sanitizer_${tool}_np1.txt 2>&1; tail -n 3 $OUT/${TAG}_sanitizer_${tool}_np1.txt
    done
    DFSA_NP=4 timeout 500 compute-sanitizer --tool memcheck --target-processes all python tools/sanitize_cases.py > $OUT/${TAG}_sanitizer_memcheck_np4.txt 2>&1; tail -n 3 $OUT/${TAG}_sanitizer_memcheck_np4.txt
    timeout 600 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"; tail -n 3 $OUT/${TAG}_bench_n1.err; cat $OUT/${TAG}_bench_n1.json
    ;;
  fused)
    timeout 900 python -m pytest tests -m gpu -q --durations=10 --maxfail=10 > $OUT/${TAG}_pytest.log 2>&1; tail -n 16 $OUT/${TAG}_pytest.log
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:fusedGateTile -c 2 -f -o $OUT/${TAG}_fused python tools/prof_fused.py 28 > $OUT/${TAG}_ncu_fused.log 2>&1
    ncu -i $OUT/${TAG}_fused.ncu-rep --page raw --csv > $OUT/${TAG}_fused_raw.csv 2>> $OUT/${TAG}_ncu_fused.log
    ncu -i $OUT/${TAG}_fused.ncu-rep --page source --csv > $OUT/${TAG}_fused_source.csv 2>> $OUT/${TAG}_ncu_fused.log
    tail -n 2 $OUT/${TAG}_ncu_fused.log
    timeout 700 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"; tail -n 3 $OUT/${TAG}_bench_n1.err; cat $OUT/${TAG}_bench_n1.json
    ;;
  fused2)
    timeout 400 python -m pytest tests -m gpu -q -k "fused or every_target or config1" > $OUT/${TAG}_pytest_fused.log 2>&1; tail -n 5 $OUT/${TAG}_pytest_fused.log
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:fusedGateTile -c 2 -f -o $OUT/${TAG}_fused python tools/prof_fused.py 28 > $OUT/${TAG}_ncu_fused.log 2>&1
    ncu -i $OUT/${TAG}_fused.ncu-rep --page raw --csv > $OUT/${TAG}_fused_raw.csv 2>> $OUT/${TAG}_ncu_fused.log
    ncu -i $OUT/${TAG}_fused.ncu-rep --page source --csv > $OUT/${TAG}_fused_source.csv 2>> $OUT/${TAG}_ncu_fused.log
    tail -n 2 $OUT/${TAG}_ncu_fused.log
    timeout 300 python tools/bench_dm_extras.py 14 > $OUT/${TAG}_dm_extras_14q.jsonl 2> $OUT/${TAG}_dm_extras.err; cat $OUT/${TAG}_dm_extras_14q.jsonl; tail -n 3 $OUT/${TAG}_dm_extras.err
    timeout 700 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"; tail -n 3 $OUT/${TAG}_bench_n1.err; cat $OUT/${TAG}_bench_n1.json
    ;;
  link)
    NP=${NP:-2}
    for f in ${INFLIGHTS:-2048 4096 8192 16384}; do
      DFSA_REMOTE_INFLIGHT=$f timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
          tools/link_sweep.py 2>> $OUT/${TAG}_link_sweep_n${NP}.err | grep '^{' >> $OUT/${TAG}_link_sweep_n${NP}.jsonl
    done
    cat $OUT/${TAG}_link_sweep_n${NP}.jsonl; tail -n 5 $OUT/${TAG}_link_sweep_n${NP}.err
    ;;
  eight)
    NP=${NP:-8}
    timeout 500 python -m pytest tests -m gpu -q --durations=8 -k "multi_rank or relocation or lazy or corrected or chunk or config1 or catch" > $OUT/${TAG}_pytest_n${NP}.log 2>&1
    tail -n 14 $OUT/${TAG}_pytest_n${NP}.log
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $NP --steps 3 --warmup 3 --per-gate \
        > $OUT/${TAG}_bench_n${NP}.json 2> $OUT/${TAG}_bench_n${NP}.err; echo "bench rc=$?"
    grep "^gate" $OUT/${TAG}_bench_n${NP}.err > $OUT/${TAG}_bench_n${NP}_per_gate.txt; tail -n 12 $OUT/${TAG}_bench_n${NP}_per_gate.txt; grep -v "^gate\|^\[" $OUT/${TAG}_bench_n${NP}.err | tail -n 5
    cat $OUT/${TAG}_bench_n${NP}.json
    DFSA_REMOTE_INFLIGHT=${BEST_INFLIGHT:-8192} timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29578 tools/link_sweep.py \
        2> $OUT/${TAG}_link_sweep_n${NP}.err | grep '^{' > $OUT/${TAG}_link_sweep_n${NP}.jsonl; cat $OUT/${TAG}_link_sweep_n${NP}.jsonl
    ;;
  next)
    # FIRST thing to run in the next GPU session: what was written after the round-2 GPU budget was spent --
    # the grouped swap-in of rank-bit qubits (multi-pair relocation in the default-mode sweep), the staggered multi-pair gather,
    # the NUMA binding of the end-to-end leg. NP = 4 or 8.
    NP=${NP:-8}
    timeout 500 python -m pytest tests -m gpu -q --durations=8 -k "fused or relocation or lazy or multi_rank_circuit" > $OUT/${TAG}_pytest_next_n${NP}.log 2>&1; tail -n 8 $OUT/${TAG}_pytest_next_n${NP}.log
    timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29579 bench.py --gpus $NP --steps 5 --warmup 3 \
        > $OUT/${TAG}_bench_n${NP}.json 2> $OUT/${TAG}_bench_n${NP}.err; echo "bench rc=$?"; grep -v "^\[" $OUT/${TAG}_bench_n${NP}.err | tail -n 5; cat $OUT/${TAG}_bench_n${NP}.json
    # A/B: rank-bit qubits brought in one pair per relocation step (the version measured in round 2)
    DFSA_GROUP_SWAPIN=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $NP --steps 5 --warmup 3 \
        --skip-configs --skip-parity > $OUT/${TAG}_bench_n${NP}_single_swapin.json 2> $OUT/${TAG}_bench_n${NP}_single_swapin.err; echo "bench (single swap-in) rc=$?"; cat $OUT/${TAG}_bench_n${NP}_single_swapin.json
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29580 tools/link_sweep.py \
        2> $OUT/${TAG}_link_sweep_n${NP}.err | grep '^{' > $OUT/${TAG}_link_sweep_n${NP}.jsonl; cat $OUT/${TAG}_link_sweep_n${NP}.jsonl
    ;;
esac
