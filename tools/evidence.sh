#!/bin/bash
# The evidence set of one build on one B200 box: tools/evidence.sh <tag>  ->  gpurun_out/<tag>_*  (copy what is to be judged into profiles/)
#   GPU tests, bench.py both arms, the 60 000-column command plain and under the ncu launch-list pass, ncu --set full of the four hot kernels.
tag=${1:-rX}
o=gpurun_out
mkdir -p $o
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $o/${tag}_pytest_gpu.log
python bench.py 2> $o/${tag}_bench_n1.err | grep '^{' > $o/${tag}_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 2> /dev/null | grep '^{' > $o/${tag}_bench_ref.json
python bench.py --columns 60000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2> /dev/null | grep '^{' > $o/${tag}_bench_60k_plain_run.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_ncu_launch_list.csv \
    python bench.py --columns 60000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-check > $o/${tag}_ncu_launch.log 2>&1
python tools/prof_case.py 30000 137 1 > $o/${tag}_prof_case_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -o $o/${tag}_kernels -f python tools/prof_case.py 30000 137 1 > $o/${tag}_ncu_full.log 2>&1
ls -la $o | tail -12
