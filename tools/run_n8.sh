TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1; lscpu | head -40 >> gpurun_out/r2_topo.txt; (numactl -H || cat /sys/devices/system/node/online) >> gpurun_out/r2_topo.txt 2>&1; free -g >> gpurun_out/r2_topo.txt
timeout 600 $TR --nproc-per-node 8 --master-port 29511 tools/bench_multi.py > gpurun_out/r2_multi_n8.jsonl 2> gpurun_out/r2_multi_n8.err; echo multi8 rc $?
timeout 400 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo bench8 rc $?
for n in 2 4; do timeout 300 $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --steps 3 --warmup 3 --no-parity > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err; echo bench$n rc $?; timeout 200 $TR --nproc-per-node $n --master-port 2953$n tools/bench_multi.py --what h2d >> gpurun_out/r2_multi_h2d.jsonl 2>/dev/null; done
tail -3 gpurun_out/r2_multi_n8.err; cat gpurun_out/r2_multi_h2d.jsonl
