#!/bin/bash
# round 2, batch d: file-chain tests + whole GPU suite; then (N GPUs) the quick strong-scaling line with the NVML sampler
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_pytest_d.log
