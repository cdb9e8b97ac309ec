#!/bin/bash
# usage: run_ranks.sh <nproc> <timeout_s> <logfile> <script> [args...]
# Launches torchrun in its own process group and kills the WHOLE group on timeout
# (a plain `timeout torchrun` orphans the workers).
n=$1; t=$2; log=$3; shift 3
mkdir -p "$(dirname "$log")"
set -m
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@" > "$log" 2>&1 &
pid=$!
( sleep $t; kill -9 -- -$pid 2>/dev/null ) &
watcher=$!
wait $pid; rc=$?
kill $watcher 2>/dev/null
echo "run_ranks rc=$rc" >> "$log"
exit $rc
