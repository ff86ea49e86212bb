#!/bin/bash
mkdir -p gpurun_out
for m in warp slice; do echo "FLB_U8_PACK=$m"; FLB_U8_PACK=$m timeout 300 python tools/opbench.py pack,for_pack --types 8; done 2>&1 | tee gpurun_out/opbench_u8pack.log
FLB_U8_PACK=slice timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "pack_every_width or for_family" 2>&1 | tail -2
