#!/bin/bash
# Round-2 GPU job 38: the complete GPU suite of HEAD in one invocation (as the driver runs it), then smoke()
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
RT_PARITY_LOG=gpurun_out/r02f_parity.jsonl timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02f_pytest_gpu.log 2>&1; tail -n 4 gpurun_out/r02f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 1
