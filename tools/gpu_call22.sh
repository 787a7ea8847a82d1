#!/bin/bash
# GPU call 22: recurrence v6 (multicast h tiles + pipelined gate loads) vs v5; GEMM epilogue with shuffled bias; pure-write bandwidth.
mkdir -p gpurun_out
LOG=gpurun_out/call22_lstm_v6.log
: > $LOG
P="timeout 120 python tools/prof_lstm.py"
for sl in 1 2 3; do
  $P --B 12 --T 40 --K 34 --axis time --slots $sl --check --reps 1 >> $LOG 2>&1 || echo "FAILED time slots=$sl rc=$?" >> $LOG
  $P --B 3 --T 300 --K 34 --axis freq --slots $sl --check --reps 1 >> $LOG 2>&1 || echo "FAILED freq slots=$sl rc=$?" >> $LOG
done
$P --B 40 --T 60 --K 34 --axis time --slots 3 --maxcl 3 --check --reps 1 >> $LOG 2>&1 || echo "FAILED multi-group time" >> $LOG
$P --B 40 --T 60 --K 34 --axis freq --slots 3 --maxcl 2 --check --reps 1 >> $LOG 2>&1 || echo "FAILED multi-group freq" >> $LOG
for ax in time freq; do
  BSRNN_LSTM_VER=5 $P --B 64 --T 1001 --K 34 --axis $ax --slots 3 --reps 3 >> $LOG 2>&1
  $P --B 64 --T 1001 --K 34 --axis $ax --slots 3 --reps 3 --trace >> $LOG 2>&1
done
grep -E "CHECK|FAILED|ms,|cycles|producer|mma  |epilogue|rror" $LOG | tail -40
if grep -q "FAIL\|rror" $LOG; then export BSRNN_LSTM_VER=5; echo "v6 FAILED -> v5 for the rest"; fi
G=gpurun_out/call22_gemm.log; : > $G
for w in inproj fc; do timeout 120 python tools/prof_gemm.py --which $w --axis time --reps 3 >> $G 2>&1; done
timeout 600 python tools/gpu_check_tc.py 2>&1 | grep -E "gemm|rror" >> $G
python - >> $G 2>&1 <<'PY'
import torch
x = torch.empty(7 * (1 << 30), dtype=torch.float16, device="cuda")      # 14 GiB
for name, fn in (("fill_", lambda: x.fill_(1.0)), ("zero_", lambda: x.zero_())):
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(f"pure write {name}: {x.numel() * 2 / e0.elapsed_time(e1) / 1e6:.0f} GB/s")
y = torch.empty_like(x)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
print(f"copy (r+w bytes): {2 * x.numel() * 2 / e0.elapsed_time(e1) / 1e6:.0f} GB/s")
s = x.view(torch.float32)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.sum(); e1.record(); torch.cuda.synchronize()
print(f"pure read sum: {x.numel() * 2 / e0.elapsed_time(e1) / 1e6:.0f} GB/s")
PY
cat $G
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/call22_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/call22_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/call22_bench.json 2> gpurun_out/call22_bench.err; echo "bench rc=$?"; cat gpurun_out/call22_bench.json; tail -3 gpurun_out/call22_bench.err
