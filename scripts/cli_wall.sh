#!/bin/bash
# whole-process wall time of the drop-in CLI vs the reference binary on the same FASTA (BASELINE configs[1] shape),
# outputs compared byte for byte.  usage: cli_wall.sh [amplicons] [extra swarm options...]
cd "$(dirname "$0")/.."; N=${1:-10000000}; shift; O=gpurun_out; mkdir -p $O
FA=/dev/shm/swb200_${N}x150_s42.fa
[ -f $FA ] || python -c "
import sys; sys.path.insert(0,'tests'); import helpers; helpers.make_fasta('$FA', $N, 150, 42)"
T=$(nproc)
s=$(date +%s.%N); bin/swarm_b200 "$@" -o /dev/shm/mine.o -s /dev/shm/mine.s -l /dev/shm/mine.log $FA; rc1=$?; e=$(date +%s.%N)
mine=$(python -c "print(round($e-$s,2))")
s=$(date +%s.%N); oracle/_ref/swarm -t $T "$@" -o /dev/shm/ref.o -s /dev/shm/ref.s -l /dev/shm/ref.log $FA; rc2=$?; e=$(date +%s.%N)
ref=$(python -c "print(round($e-$s,2))")
cmp -s /dev/shm/mine.o /dev/shm/ref.o && so=identical || so=DIFFERENT
cmp -s /dev/shm/mine.s /dev/shm/ref.s && ss=identical || ss=DIFFERENT
echo "{\"amplicons\": $N, \"options\": \"$*\", \"swarm_b200_wall_s\": $mine, \"reference_wall_s\": $ref, \"reference_threads\": $T, \"rc\": [$rc1, $rc2], \"swarms_file\": \"$so\", \"stats_file\": \"$ss\", \"bytes\": $(stat -c %s /dev/shm/ref.o)}"
rm -f /dev/shm/mine.* /dev/shm/ref.*
