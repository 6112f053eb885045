#!/bin/bash
mkdir -p gpurun_out
echo "=== umma check"; timeout 600 python tools/probe_umma.py all > gpurun_out/umma_probe.txt 2>&1; echo "rc=$?"; tail -40 gpurun_out/umma_probe.txt
bash tools/run_tk_diag.sh
