#!/bin/bash
TAG=${1:-r02z}
export B200BA_PANEL=2
timeout 200 ncu --set full --clock-control none --import-source on -k regex:potrf_trinv_tile -s 3 -c 1 -o gpurun_out/${TAG}_prof_tile -f python scripts/dense_timing.py 2048 > gpurun_out/${TAG}_prof_tile.log 2>&1; tail -3 gpurun_out/${TAG}_prof_tile.log; ls -la gpurun_out/${TAG}_prof_tile.ncu-rep
