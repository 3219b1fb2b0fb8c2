#!/bin/bash
# parameter sweep with unit skipping on (target workload): unit height, env groups, sweep blocks per SM,
# and where the end-to-end time goes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/shot2; mkdir -p $O
B="python bench.py --no-cpu-baseline --steps 100 --e2e-steps 10 --roofline-steps 10"
run() { name=$1; shift; ( timeout 90 "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err ); }
run base $B
run rpc14 $B --rows-per-chunk 14
run rpc30 $B --rows-per-chunk 30
run groups2 $B --env-groups 2
run groups8 $B --env-groups 8
SFB_SWEEP_BLOCKS_PER_SM=1 run bps1 $B
SFB_SWEEP_BLOCKS_PER_SM=4 run bps4 $B
SFB_DEBUG_TIMING=1 run dbg $B
run cfg3_groups8 $B --workload cfg3 --env-groups 8
run cfg3_rpc14 $B --workload cfg3 --rows-per-chunk 14
nproc > $O/nproc.txt
for f in $O/*.json; do echo $f; cut -c1-160 $f; done
