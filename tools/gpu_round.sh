#!/bin/bash
# One GPU-box visit: parity tests, the bench lines, the ncu launch list and the --set full captures.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [stages...]   stages: test smoke bench ref launches full matchprof
set -u
TAG=${1:-r2}; shift || true
STAGES=${*:-"test smoke bench ref launches full"}
OUT=gpurun_out; mkdir -p $OUT
has() { case " $STAGES " in *" $1 "*) return 0;; esac; return 1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
if has test; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
  tail -3 $OUT/pytest_gpu_$TAG.log
fi
if has smoke; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log
fi
if has bench; then      # the default line: cfg4 headline + cfg2 + matching + pose refinement
  timeout 1200 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; tail -c 400 $OUT/bench_${TAG}.json; tail -3 $OUT/bench_${TAG}.err
fi
if has ref; then        # the CPU arm on the same box
  timeout 900 python bench.py --impl reference > $OUT/bench_${TAG}_ref.json 2>> $OUT/bench_${TAG}.err; tail -c 300 $OUT/bench_${TAG}_ref.json
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-cfg2 > $OUT/launches_$TAG.log 2>&1
fi
if has full; then
  # one instance of every BA kernel on cfg2 and on cfg4 (first LM iteration), and the matcher kernels.  The reports are
  # converted to the raw-metric CSV on the box and dropped (gpurun brings back at most 64 MiB).
  for cfg in cfg2 cfg4; do
    MM_NO_TIME_KERNEL=1 timeout 900 ncu --set full --clock-control none \
        -k regex:'k_residual_jacobian|k_schur_point|k_schur_blocks|k_schur_cam|k_backsub|k_tc_factor|k_tc_apply|k_dpcg_spmv' -c 12 -f -o $OUT/prof_ba_${cfg}_$TAG \
        python tools/time_ba.py $cfg 1 > $OUT/prof_ba_${cfg}_$TAG.log 2>&1
    ncu -i $OUT/prof_ba_${cfg}_$TAG.ncu-rep --page raw --csv > $OUT/prof_ba_${cfg}_$TAG.csv 2>/dev/null; rm -f $OUT/prof_ba_${cfg}_$TAG.ncu-rep
  done
fi
if has full || has matchprof; then
  timeout 600 ncu --set full --clock-control none \
      -k regex:'k_match_tc|k_rerank|k_tc_prep' -c 4 -f -o $OUT/prof_match_$TAG \
      python tools/time_match.py 64 tc > $OUT/prof_match_$TAG.log 2>&1
  ncu -i $OUT/prof_match_$TAG.ncu-rep --page raw --csv > $OUT/prof_match_$TAG.csv 2>/dev/null; rm -f $OUT/prof_match_$TAG.ncu-rep
fi
du -sh $OUT
echo done
