# One-GPU evidence pass of a round (run on the B200 box: gpurun --timeout 1500 -- 'bash tools/run_evidence.sh r2c'):
# the whole GPU suite, smoke(), bench.py (both arms), every secondary config, the ncu launch list of the bench command and
# one `ncu --set full` capture each of mfcc_kernel and cnn_tc_kernel.  Everything lands in gpurun_out/<tag>_*; what is judged is copied to profiles/.
TAG=${1:-rX}
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo bench rc $?
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_reference_arm.err; echo ref rc $?
(python tools/bench_configs.py; python tools/bench_configs.py --frontdsp) > $O/${TAG}_configs.jsonl 2> $O/${TAG}_configs.err; echo configs rc $?
python tools/time_modes.py > $O/${TAG}_frontend_modes.txt 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-handoff > $O/${TAG}_bench_profiled_command.json 2>/dev/null &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'mfcc_kernel|cnn_|ctc_|fused_clip' -c 600 --csv --log-file $O/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-handoff > $O/${TAG}_ncu_launch.log 2>&1
python tools/prof_driver.py 65536 mfcc > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on --kernel-name regex:mfcc_kernel --launch-skip 2 -c 1 -f \
      -o $O/${TAG}_mfcc python tools/prof_driver.py 65536 mfcc > $O/${TAG}_ncu_mfcc.log 2>&1
python tools/prof_driver.py 131072 fused tensor > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on --kernel-name regex:cnn_tc_kernel --launch-skip 2 -c 1 -f \
      -o $O/${TAG}_cnn_tc python tools/prof_driver.py 131072 fused tensor > $O/${TAG}_ncu_cnn_tc.log 2>&1
for s in 1 2 3; do python tools/fuzz_gpu.py $s 2>&1 | tail -1; done
python - <<PY
import json
d = json.load(open("$O/${TAG}_bench_n1.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"], d["parity"]["ok"], d["parity"]["decision_mismatches"])
PY
