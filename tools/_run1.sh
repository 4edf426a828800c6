ETPGT_GEMM_BN=256 timeout 120 python tools/exp_gemm.py > gpurun_out/r02w_gemm_bn256.txt 2>&1
ETPGT_GEMM_BN=128 timeout 120 python tools/exp_gemm.py > gpurun_out/r02w_gemm_bn128.txt 2>&1
timeout 120 python tools/exp_gemm.py > gpurun_out/r02w_gemm_default.txt 2>&1
tail -n 2 gpurun_out/r02w_gemm_bn256.txt gpurun_out/r02w_gemm_bn128.txt gpurun_out/r02w_gemm_default.txt
