#!/bin/bash
# Round-2 GPU job 35: one shading step of the 300-sphere random scene against the oracle, grouped by material
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/random_scene_probe.py 106 300 > gpurun_out/random_probe.log 2>&1; cat gpurun_out/random_probe.log | cut -c1-260
