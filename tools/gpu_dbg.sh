#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_dbg2.py > gpurun_out/dbg.log 2>&1
