#!/bin/bash
# round 2, run F (2 GPUs): bench under torchrun — shared frame on GPU 0 over IPC, parity at N=2; progressive8k at N=2; multi-GPU tests
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_f_topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_f_n2_teapot.json 2> gpurun_out/r2_f_n2_teapot.err
echo "n2 teapot rc=$?"; tail -3 gpurun_out/r2_f_n2_teapot.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus 2 --steps 3 --warmup 1 --workload progressive8k > gpurun_out/r2_f_n2_prog.json 2> gpurun_out/r2_f_n2_prog.err
echo "n2 prog rc=$?"; tail -3 gpurun_out/r2_f_n2_prog.err
timeout 600 python -m pytest tests -x -q -m gpu --timeout 300 -k "multi_gpu or progressive_psnr" > gpurun_out/r2_f_tests.log 2>&1
tail -3 gpurun_out/r2_f_tests.log
python - <<PY
import json
for f in ("r2_f_n2_teapot","r2_f_n2_prog"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["ms_per_frame"], d["e2e"].get("d2h_floor_ms"), d["e2e"].get("ms_reduce"))
        print("  parity", d["parity"])
    except Exception as e: print(f, "no json", e)
PY
