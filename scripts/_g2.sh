python -m pytest tests/test_gpu_parity.py -q -x -k "gemm_kernel" > gpurun_out/t2_gemm.log 2>&1; echo "gemm rc=$?"; tail -15 gpurun_out/t2_gemm.log
python -m pytest tests -m gpu -q > gpurun_out/t2_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/t2_pytest.log
