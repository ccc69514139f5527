#!/bin/bash
mkdir -p gpurun_out
{
which gdb strace
LD_PRELOAD=$PWD/tools/dbg/exit_trace.so NB200_TRACE=1 python -X faulthandler -u -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -60; echo "rc=${PIPESTATUS[0]}"
} > gpurun_out/dbg.log 2>&1
tail -c 6000 gpurun_out/dbg.log
