#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/prof_step.py pyg 3 > gpurun_out/step_profile_v5.txt 2> gpurun_out/step_profile_v5.err; echo "exit $?"
head -90 gpurun_out/step_profile_v5.txt | cut -c1-200; tail -3 gpurun_out/step_profile_v5.err
