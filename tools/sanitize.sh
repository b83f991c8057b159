# compute-sanitizer passes over a small but complete slice of the GPU suite (run under gpurun)
set -x
compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "se_ard_small or one_observation or n129 or d40 or badly_scaled_se or factor" 2>&1 | tail -8
echo "memcheck parity rc=$?"
compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "se_iso and py" 2>&1 | tail -8
echo "racecheck rc=$?"
compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_golden.py -x -q -m gpu -k "acqmaxGP and branin_ard50" 2>&1 | tail -6
echo "memcheck golden rc=$?"
# session 4 kernels: fused small-model kernel (both sides), marginal likelihood, append, Laplace fit
compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_tiny.py -x -q -m gpu -k "n1 or n5 or n127 or n128 or prior or d40 or direct" 2>&1 | tail -6
echo "memcheck tiny rc=$?"
compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_api.py -x -q -m gpu -k "aug_variance" 2>&1 | tail -6
echo "memcheck aug rc=$?"
compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_hyper.py -x -q -m gpu -k "known_answers or derivative_matrices or (nlml_and_gradient and 0.1)" 2>&1 | tail -6
echo "memcheck hyper rc=$?"
compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_tiny.py -x -q -m gpu -k "n5 and cpp" 2>&1 | tail -6
echo "racecheck tiny rc=$?"
compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_append.py tests/test_gpu_laplace.py -x -q -m gpu -k "127 or small" 2>&1 | tail -6
echo "memcheck append/laplace rc=$?"
# round 2: INT8 wide-batch path (K1 / K2 / guard pass), device-side Laplace matrix, look-ahead factorisation (every model build above)
compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_int8.py -x -q -m gpu -k "(default_wide and (300-3 or 129-1 or 200-2)) or guard or batch_size_classes" 2>&1 | tail -6
echo "memcheck int8 rc=$?"
compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_laplace.py -x -q -m gpu -k "assembled" 2>&1 | tail -6
echo "memcheck pref-C rc=$?"
# model build rewritten in round 2 (tensor-core diagonal block kernel, half-height tiles, paired block columns, four streams)
compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "factor_schedules and 640" 2>&1 | tail -6
echo "racecheck model build rc=$?"
compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "factor_schedules" 2>&1 | tail -6
echo "memcheck model build rc=$?"
# resident batch server of the small-model kernel (mailbox in mapped host memory); under the sanitizer its idle timeout also fires
compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_tiny.py -x -q -m gpu -k "batch_server" 2>&1 | tail -6
echo "memcheck batch server rc=$?"
