#!/bin/bash
# per-site norm-block timings (forward / backward) at the training batch; see tools/prof_nb.py
B=${1:-512}
for s in "32 192 30 self" "64 96 60 plain" "64 192 30 self" "64 96 60 self" "128 96 15 self" "128 48 30 plain" "128 48 30 self" \
         "256 48 8 self" "256 24 15 self" "128 24 15 self" "32 48 30 self" "64 48 30 self"; do
  set -- $s
  printf "C%-4s %3sx%-3s %-5s " $1 $2 $3 $4
  python tools/prof_nb.py $1 $2 $3 $B $4
done
