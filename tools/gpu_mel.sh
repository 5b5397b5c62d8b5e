set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q -s -k "mel_front" > gpurun_out/r2_mel_tests.log 2>&1; echo "mel tests rc $?"; tail -n 25 gpurun_out/r2_mel_tests.log
