#!/bin/bash
# One parameterised GPU-box script (replaces the per-call scripts of round 1).  Usage, under gpurun:
#   bash tools/gpu_run.sh TAG step [step ...]
# Steps:  tests | tests:<pytest -k expr> | smoke | bench:<WL> | benchq:<WL> (1 step, no CPU leg) | diag:<WL>[,<WL>]
#         | dtau:<WL>:<K> | bisect:<WL>:<K> | bias | benchn:<WL>:<N>:<potrf_dist> | benchfull:<N> | time:<m>[,<m>] | syrk[:<m>,<n>] | create:<WL> | launches:<WL> | ncu:<WL>:<kernel regex>:<skip>:<count> | ref:<WL>
# Everything is written under gpurun_out/ with TAG in the name; a failing step does not stop the later ones.
set -u
TAG=$1; shift
mkdir -p gpurun_out
for step in "$@"; do
  kind=${step%%:*}; arg=${step#*:}; [ "$arg" == "$step" ] && arg=""
  t0=$(date +%s)
  case $kind in
    tests)
      if [ -n "$arg" ]; then
        timeout 1500 python -m pytest tests -m gpu -x -q -rP -k "$arg" > gpurun_out/pytest_${TAG}.log 2>&1
      else
        timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1
      fi
      echo "[$step] rc=$? $(tail -1 gpurun_out/pytest_${TAG}.log)";;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1
      echo "[$step] rc=$? $(tail -1 gpurun_out/smoke_${TAG}.log)";;
    bench)
      timeout 900 python bench.py --workload $arg > gpurun_out/bench_${arg}_${TAG}.json 2> gpurun_out/bench_${arg}_${TAG}.err
      echo "[$step] rc=$? $(head -c 600 gpurun_out/bench_${arg}_${TAG}.json)";;
    benchq)
      timeout 600 python bench.py --workload $arg --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/benchq_${arg}_${TAG}.json 2> gpurun_out/benchq_${arg}_${TAG}.err
      echo "[$step] rc=$? $(python tools/phase_line.py gpurun_out/benchq_${arg}_${TAG}.json)";;
    benchn)   # benchn:<WL>:<N>:<potrf_dist>  -- N ranks on one box through torchrun, 1 step, no CPU leg
      IFS=: read -r wl ng pd <<< "$arg"
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $ng --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $ng --workload $wl --steps 1 --warmup 1 --no-cpu-baseline --potrf-dist $pd --no-c5 \
        > gpurun_out/benchn_${wl}_n${ng}_d${pd}_${TAG}.json 2> gpurun_out/benchn_${wl}_n${ng}_d${pd}_${TAG}.err
      echo "[$step] rc=$? $(python tools/phase_line.py gpurun_out/benchn_${wl}_n${ng}_d${pd}_${TAG}.json)";;
    benchfull)   # benchfull:<N>  -- the driver's own command line at N ranks (C3 headline + C5 record)
      timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $arg --master-addr 127.0.0.1 --master-port 29512 \
        bench.py --gpus $arg --steps 2 --warmup 1 > gpurun_out/benchfull_n${arg}_${TAG}.json 2> gpurun_out/benchfull_n${arg}_${TAG}.err
      echo "[$step] rc=$? $(python tools/phase_line.py gpurun_out/benchfull_n${arg}_${TAG}.json)";;
    ref)
      timeout 900 python bench.py --impl reference --workload $arg --steps 1 --warmup 0 > gpurun_out/ref_${arg}_${TAG}.json 2> gpurun_out/ref_${arg}_${TAG}.err
      echo "[$step] rc=$? $(head -c 700 gpurun_out/ref_${arg}_${TAG}.json)";;
    diag)
      timeout 1500 python tools/diag_accuracy.py ${arg//,/ } > gpurun_out/diag_${TAG}.log 2>&1
      echo "[$step] rc=$? $(wc -l < gpurun_out/diag_${TAG}.log) lines";;
    dtau)
      IFS=: read -r wl k <<< "$arg"
      timeout 900 python tools/diag_dtau.py $wl $k > gpurun_out/dtau_${wl}_${k}_${TAG}.log 2>&1
      echo "[$step] rc=$?"; cat gpurun_out/dtau_${wl}_${k}_${TAG}.log | tail -40;;
    bisect)
      IFS=: read -r wl k <<< "$arg"
      timeout 900 python tools/diag_bisect_host.py $wl $k > gpurun_out/bisect_${wl}_${k}_${TAG}.log 2>&1
      echo "[$step] rc=$?"; cat gpurun_out/bisect_${wl}_${k}_${TAG}.log | tail -30;;
    bias)
      timeout 600 python tools/dmma_bias.py > gpurun_out/dmma_bias_${TAG}.log 2>&1
      echo "[$step] rc=$?"; cat gpurun_out/dmma_bias_${TAG}.log;;
    time)
      timeout 600 python tools/time_kernels.py ${arg//,/ } > gpurun_out/time_${TAG}.log 2>&1
      echo "[$step] rc=$?"; cat gpurun_out/time_${TAG}.log | tail -4;;
    syrk)
      timeout 600 python tools/time_syrk.py ${arg//,/ } > gpurun_out/syrk_${TAG}.log 2>&1
      echo "[$step] rc=$?"; cat gpurun_out/syrk_${TAG}.log | tail -12;;
    create)
      LPB_TIME_CREATE=1 timeout 600 python tools/time_create.py $arg > gpurun_out/create_${arg}_${TAG}.log 2>&1
      echo "[$step] rc=$?"; tail -12 gpurun_out/create_${arg}_${TAG}.log;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
        --log-file gpurun_out/launches_${arg}_${TAG}.csv python bench.py --workload $arg --steps 1 --warmup 0 --no-cpu-baseline --no-e2e \
        > gpurun_out/launches_${arg}_${TAG}.out 2>&1
      echo "[$step] rc=$?"; python tools/launch_summary.py gpurun_out/launches_${arg}_${TAG}.csv > gpurun_out/launches_${arg}_${TAG}.txt 2>&1; head -30 gpurun_out/launches_${arg}_${TAG}.txt;;
    ncu)
      IFS=: read -r wl regex skip count <<< "$arg"
      name=$(echo "$regex" | tr -c 'A-Za-z0-9_' '_')
      timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$regex" --launch-skip $skip --launch-count $count \
        -f -o gpurun_out/ncu_${name}_${wl}_${TAG} python bench.py --workload $wl --steps 1 --warmup 0 --no-cpu-baseline --no-e2e \
        > gpurun_out/ncu_${name}_${wl}_${TAG}.out 2>&1
      echo "[$step] rc=$? $(ls -la gpurun_out/ncu_${name}_${wl}_${TAG}.ncu-rep 2>/dev/null | awk '{print $5}') bytes";;
    *) echo "unknown step $step";;
  esac
  echo "   ($step took $(( $(date +%s) - t0 )) s)"
done
