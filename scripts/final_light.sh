#!/bin/bash
# Round-end evidence run without the ncu --set full captures (kernels unchanged since the last full capture):
# tests, smoke, both bench arms, launch lists, linear-AE timing.   bash scripts/final_light.sh <tag>
tag=${1:-final}; out=gpurun_out/$tag; mkdir -p $out
python -m pytest tests -m gpu -q --timeout 900 > $out/pytest_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1
SECONDS=0
python bench.py --steps 20 --warmup 5 > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench.py wall clock: $SECONDS s" > $out/bench_wall.txt
SECONDS=0
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
echo "bench.py --impl reference wall clock: $SECONDS s" >> $out/bench_wall.txt
python scripts/time_linae.py > $out/linae.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_cae_2steps.csv python scripts/prof_cae.py cae 2 > $out/prof_cae.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_arhmm_2steps.csv python scripts/prof_cae.py hmm 2 > $out/prof_hmm.log 2>&1
tail -n 3 $out/pytest_gpu.txt; tail -n 2 $out/smoke.txt; head -c 300 $out/bench_n1.json; echo; head -c 300 $out/bench_reference.json; echo; cat $out/bench_wall.txt $out/linae.txt
