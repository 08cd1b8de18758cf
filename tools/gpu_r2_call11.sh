#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "bn_act or critic or colsum" > gpurun_out/r2m_bn.log 2>&1; echo "bn tests rc=$?"; tail -2 gpurun_out/r2m_bn.log
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "gan or critic or Gan or GAN" > gpurun_out/r2m_gan.log 2>&1; echo "gan tests rc=$?"; tail -2 gpurun_out/r2m_gan.log
python tools/d_kernels.py > gpurun_out/r2m_d_kernels.txt 2>&1; grep -v Warn gpurun_out/r2m_d_kernels.txt | head -24
