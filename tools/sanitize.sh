# compute-sanitizer passes over a small but complete slice of the GPU suite (run under gpurun)
set -x
compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "se_ard_small or one_observation or n129 or d40 or badly_scaled_se or factor" 2>&1 | tail -8
echo "memcheck rc=$?"
compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "se_iso and py" 2>&1 | tail -8
echo "racecheck rc=$?"
compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_golden.py -x -q -m gpu -k "acqmaxGP and branin_ard50" 2>&1 | tail -6
