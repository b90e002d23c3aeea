#!/bin/bash
# nsb_launch.sh N <binary> [args...] -- start a driver once per GPU, the way the reference runs under
# `mpirun -n N` (Navier-Stokes/README.md): exports RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR /
# NSB_RDV_PORT (csrc/host/rendezvous.hpp) and waits for all ranks; rank 0's output goes to the terminal,
# the other ranks' to rank<r>.log.  Exit code: the first non-zero exit code of any rank.
set -u
n=${1:?usage: nsb_launch.sh N binary [args...]}; shift
port=${NSB_RDV_PORT:-$((20000 + RANDOM % 20000))}
pids=()
for ((r = 0; r < n; ++r)); do
  if ((r == 0)); then
    RANK=$r WORLD_SIZE=$n LOCAL_RANK=$r MASTER_ADDR=127.0.0.1 NSB_RDV_PORT=$port "$@" &
  else
    RANK=$r WORLD_SIZE=$n LOCAL_RANK=$r MASTER_ADDR=127.0.0.1 NSB_RDV_PORT=$port "$@" > "rank$r.log" 2>&1 &
  fi
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || { c=$?; ((rc == 0)) && rc=$c; }; done
exit $rc
