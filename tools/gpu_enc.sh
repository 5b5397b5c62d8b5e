set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q -s -k "vae_encode" > gpurun_out/r2_enc_tests.log 2>&1; echo "enc tests rc $?"; tail -n 25 gpurun_out/r2_enc_tests.log
