#!/bin/bash
# Everything one GPU call should bring back: parity tests, smoke, a bench line, the ncu launch list of
# the same bench command and one full capture of the hot kernels.  Run under gpurun from the repo root:
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag]'
# Environment: SKIP_TESTS=1, SKIP_NCU=1, SKIP_BENCH=1, TESTS="<pytest selection>", BENCH_ARGS="..."
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
if [ "$SKIP_TESTS" != "1" ]; then
python -m pytest ${TESTS:-tests} -m gpu -x -q --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
fi
if [ "$SKIP_BENCH" != "1" ]; then
python bench.py --steps 20 --warmup 5 $BENCH_ARGS > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
fi
if [ "$SKIP_NCU" != "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-config3 > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ssg_|row_loss_t' -s 9 -c 3 -f -o $OUT/prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-config3 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
ls -la $OUT
