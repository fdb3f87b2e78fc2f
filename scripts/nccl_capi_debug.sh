#!/bin/bash
# Run the torch-free NCCL workers directly (2 ranks) with NCCL_DEBUG=INFO, logs under gpurun_out/.
mkdir -p gpurun_out; rm -f /tmp/mlbsim_nccl_id
for r in 0 1; do
  NCCL_DEBUG=INFO timeout 70 python tests/nccl_capi_worker.py $r 2 /tmp/mlbsim_nccl_id > gpurun_out/nccl_capi_rank$r.log 2>&1 &
done
wait
for r in 0 1; do echo "== rank $r"; grep -v "NCCL INFO" gpurun_out/nccl_capi_rank$r.log | tail -12; grep "NCCL INFO" gpurun_out/nccl_capi_rank$r.log | tail -4; done
