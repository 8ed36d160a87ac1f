#!/bin/bash
# Copy the evidence of `scripts/final_profile.sh <tag>` from gpurun_out/<tag>/ into profiles/ under the name r02_<label>_*:
#   bash scripts/collect_profile.sh r02y y
tag=$1; lab=$2; U=gpurun_out/$tag; P=profiles/r02_${lab}
python scripts/ncu_traffic.py $U/cae_full.raw.csv --estep $U/hmm_full.raw.csv > profiles/r02_kernel_traffic.json
python scripts/ncu_keymetrics.py $U/cae_full.raw.csv $U/hmm_full.raw.csv > ${P}_ncu_full.txt
python scripts/sass_summary.py > ${P}_sass_summary.txt
cp $U/bench_n1.json ${P}_bench_n1.json
cp $U/bench_reference.json ${P}_bench_reference.json
cp $U/launches_cae_2steps.csv ${P}_launches_cae_2steps.csv
cp $U/launches_arhmm_2steps.csv ${P}_launches_arhmm_2steps.csv
(python scripts/summarize_launches.py $U/launches_cae_2steps.csv; python scripts/summarize_launches.py $U/launches_arhmm_2steps.csv) > ${P}_launches_summary.txt
(tail -n 4 $U/pytest_gpu.txt; cat $U/smoke.txt) > ${P}_pytest_smoke.txt
