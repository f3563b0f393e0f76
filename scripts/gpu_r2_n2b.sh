#!/bin/bash
# 2 GPUs: bench under torchrun after the batch-kernel work, and the socket rendezvous of the row partition
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2ax_bench_n2.json 2> gpurun_out/r2ax_bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ax_bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value %.4g e2e %.4g hardi %s" % (d["value"], d["e2e"]["value"], d.get("hardi")))
print("partitioned", d.get("partitioned"))
PY
tail -3 gpurun_out/r2ax_bench_n2.err
# torch-free plumbing: two ranks started by hand, SocketComm
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29733 WORLD_SIZE=2
RANK=1 LOCAL_RANK=1 timeout 300 python scripts/dist_solve.py --ecs 400 --comm socket > gpurun_out/r2ax_socket_r1.log 2>&1 &
RANK=0 LOCAL_RANK=0 timeout 300 python scripts/dist_solve.py --ecs 400 --comm socket --check > gpurun_out/r2ax_socket_r0.log 2>&1
wait
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2ax_socket_r0.log") if l.startswith("{")][-1])
    print("socket comm, 2 GPUs:", {k: d[k] for k in ("workload","world","us_per_iteration","rel_signal_err","speedup_vs_single_gpu")})
except Exception as e:
    print("socket run failed", e); print(open("gpurun_out/r2ax_socket_r0.log").read()[-1500:])
PY
