#!/bin/bash
# Round-end evidence run on one B200: tests, smoke, both bench arms, launch lists and ncu captures.
#   bash scripts/final_profile.sh <tag>     (writes gpurun_out/<tag>/...)
tag=${1:-final}; out=gpurun_out/$tag; mkdir -p $out
python -m pytest tests -m gpu -q --timeout 900 > $out/pytest_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1
python bench.py --steps 20 --warmup 5 > $out/bench_n1.json 2> $out/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_cae_2steps.csv python scripts/prof_cae.py cae 2 > $out/prof_cae.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_arhmm_2steps.csv python scripts/prof_cae.py hmm 2 > $out/prof_hmm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"igemm_|wgrad_|_halo|thin_" -s 0 -c 30 -o $out/cae_full python scripts/prof_cae.py cae 1 > $out/ncu_cae.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"scan2|emission" -s 0 -c 2 -o $out/hmm_full python scripts/prof_cae.py hmm 1 > $out/ncu_hmm.log 2>&1
ncu -i $out/cae_full.ncu-rep --page raw --csv > $out/cae_full.raw.csv 2>/dev/null
ncu -i $out/hmm_full.ncu-rep --page raw --csv > $out/hmm_full.raw.csv 2>/dev/null
rm -f $out/cae_full.ncu-rep $out/hmm_full.ncu-rep
tail -n 3 $out/pytest_gpu.txt; cat $out/smoke.txt | tail -n 2; head -c 300 $out/bench_n1.json; echo; head -c 300 $out/bench_reference.json
