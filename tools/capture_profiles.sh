#!/bin/bash
# Evidence run on one B200 (under gpurun): bench, ncu launch list, ncu --set full of the train kernel (148 trials, one CTA per
# trial) and of the 8-CTA-cluster build (one trial), in-kernel stage timers.  usage: tools/capture_profiles.sh r02 [a|b]
# (gpurun returns at most 64 MiB per call and the two ncu reports together exceed it: part a = everything but the cluster
# capture, part b = the cluster capture; no second argument = both)
TAG=${1:-r02}
PART=${2:-ab}
O=gpurun_out
mkdir -p $O
if [[ $PART == *b* && $PART != *a* ]]; then
  ncu --set full --clock-control none --import-source on -k regex:raae_train_kernel -s 4 -c 1 -f -o $O/prof_${TAG}_cluster8 \
      python tools/cluster_bench.py "1:8" 3 > $O/ncu_cluster_$TAG.log 2>&1
  ls -la $O/*.ncu-rep
  exit 0
fi
python bench.py --steps 10 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-single --no-configs --no-peak > $O/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raae_train_kernel -s 3 -c 1 -f -o $O/prof_${TAG}_final \
    python bench.py --steps 1 --warmup 3 --trials 148 --no-cpu --no-single --no-configs --no-peak > $O/ncu_full_$TAG.log 2>&1
if [[ $PART == *b* ]]; then
  ncu --set full --clock-control none --import-source on -k regex:raae_train_kernel -s 4 -c 1 -f -o $O/prof_${TAG}_cluster8 \
      python tools/cluster_bench.py "1:8" 3 > $O/ncu_cluster_$TAG.log 2>&1
fi
python tools/stage_profile.py 148 3 > $O/stage_profile_t148.txt 2>&1
python tools/stage_profile.py 1 3 > $O/stage_profile_t1.txt 2>&1
RAAE_CTAS_PER_TRIAL=8 python tools/stage_profile.py 1 3 > $O/stage_profile_t1_c8.txt 2>&1
tail -3 $O/stage_profile_t1_c8.txt
ls -la $O/*.ncu-rep
