#!/bin/bash
# round 2, pass aj: racecheck + memcheck over tools/sanitize_small.py after the tcgen05 kernel changes (fir_umma32t_kernel included)
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 compute-sanitizer --tool racecheck --target-processes all python tools/sanitize_small.py > $O/r02aj_sanitize_racecheck.log 2>&1 ) 2>&1 | grep real
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $O/r02aj_sanitize_racecheck.log | sort | uniq -c | head
( time timeout 900 compute-sanitizer --tool memcheck --target-processes all python tools/sanitize_small.py > $O/r02aj_sanitize_memcheck.log 2>&1 ) 2>&1 | grep real
grep -E "ERROR SUMMARY" $O/r02aj_sanitize_memcheck.log | sort | uniq -c
grep -c "fir_umma32t_kernel" $O/r02aj_sanitize_memcheck.log
