#!/bin/bash
# round 2, call T (1 GPU): the reference's GPUEngine over the drop-in GPUEngineCuda binding, on the device
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_integration_binding.py -m gpu -q -x > gpurun_out/r2t_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log; tail -25 gpurun_out/r2t_pytest.log
