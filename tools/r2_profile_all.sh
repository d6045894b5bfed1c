#!/bin/bash
# Round-end evidence run on one B200: GPU tests, smoke, the default bench, ncu launch lists, ncu --set full condensed to JSON.
# Everything lands in gpurun_out/; profiles/ is filled from there by hand (see profiles/README.md).
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2_pytest_gpu.log
tail -3 $O/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2_smoke.log 2>&1; tail -1 $O/r2_smoke.log
timeout 900 python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/r2_bench_ref.json 2> $O/r2_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-bow --no-flow --no-check --no-stereo --no-submap > $O/r2_ncu_bench.log 2>&1; echo "ncu bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_prof_run.csv \
    python tools/prof_run.py > $O/r2_ncu_prof_run.log 2>&1; echo "ncu prof_run rc=$?"
PROF_LIGHT=1 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:pyramid|fast_|octree|blur|describe|assign_slots|hamming|umma|top2|flow_|stereo' -o /tmp/r2_full -f \
    python tools/prof_run.py > $O/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_to_json.py /tmp/r2_full.ncu-rep $O/r2_ncu_kernels.json 64 > $O/r2_ncu_kernels.md 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hamming_top2_umma -c 2 -o $O/r2_umma -f \
    python tools/umma_prof.py > $O/r2_ncu_umma.log 2>&1; echo "ncu umma rc=$?"
python tools/ncu_keys.py $O/r2_umma.ncu-rep tensor utcimma smem shared > $O/r2_ncu_umma_keys.txt 2>&1
ls -la $O | grep r2_
