#!/bin/bash
# round 2: everything that needs 2 GPUs — the multi-rank parity tests (kept as a log: the driver's GPUTEST runs on 1 GPU and skips them)
# and the 2-rank bench lines (NCCL allreduce / fused peer-memory reduce) with the in-bench parity check
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2m2_gpus.txt
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_wide.py tests/test_gpu_estimators.py -m gpu -q -rs -k "data_parallel or rank or two_ranks or two_tasks" > gpurun_out/r2m2_pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2m2_pytest_multi.log
tail -15 gpurun_out/r2m2_pytest_multi.log
for fused in 0 1; do
  B200GBM_FUSED_REDUCE=$fused timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2m2_bench_n2_fused$fused.json 2> gpurun_out/r2m2_bench_n2_fused$fused.err
  tail -2 gpurun_out/r2m2_bench_n2_fused$fused.err; cat gpurun_out/r2m2_bench_n2_fused$fused.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step','histogram_reduce','parity_check','bins_sample_check','hist_conservation_check','timed_model')}, d['roofline']['frac'], d['e2e']['value'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/r2m2_bench_n2_reference.json 2> gpurun_out/r2m2_bench_n2_reference.err; cat gpurun_out/r2m2_bench_n2_reference.json | cut -c1-600
