#!/bin/bash
mkdir -p gpurun_out
python tools/e2e_probe.py 32 2>&1 | tail -13
python tools/e2e_probe.py 64 2>&1 | tail -6
