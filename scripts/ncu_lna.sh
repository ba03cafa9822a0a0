#!/bin/bash
# ncu full capture of one LNA launch at the config-2 model (run under gpurun)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lna_f32 -s 1 -c 1 -o gpurun_out/$1 -f python scripts/ncu_gmm.py 0 30 2>&1 | tail -2
