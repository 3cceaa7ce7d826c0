#!/bin/bash
mkdir -p gpurun_out
( timeout 90 python __graft_entry__.py --smoke ) > gpurun_out/r2_dbg_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_dbg_smoke.log
( PPRTS_B200_LOOKUP_PER_THREAD=1 timeout 90 python __graft_entry__.py --smoke ) > gpurun_out/r2_dbg_smoke_old.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_dbg_smoke_old.log
( timeout 240 compute-sanitizer --tool memcheck python __graft_entry__.py --smoke ) 2>&1 | tail -40 > gpurun_out/r2_dbg_memcheck.log
tail -3 gpurun_out/r2_dbg_smoke.log gpurun_out/r2_dbg_smoke_old.log; tail -15 gpurun_out/r2_dbg_memcheck.log
