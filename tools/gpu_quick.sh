#!/bin/bash
timeout 300 python tools/opbench.py unpack_filter --types 32,64
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
