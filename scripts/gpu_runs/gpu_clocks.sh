mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader,nounits -lms 50 > gpurun_out/clocks.csv &
SMI=$!
timeout 900 python bench.py --steps 60 --warmup 5 --no-kernel-roofline --no-cpu-baseline > gpurun_out/bench_long.log 2>&1
kill $SMI
grep -v Warning gpurun_out/bench_long.log | tail -1 | cut -c1-250
python - <<'PY'
rows=[l.strip().split(', ') for l in open('gpurun_out/clocks.csv') if l.strip()]
import statistics
sm=[float(r[0]) for r in rows]; pw=[float(r[2]) for r in rows]
busy=[(s,p) for s,p in zip(sm,pw) if p>500]
print("samples",len(rows),"busy",len(busy))
if busy:
    print("busy sm MHz median %.0f min %.0f max %.0f; power median %.0f max %.0f"%(statistics.median([b[0] for b in busy]),min(b[0] for b in busy),max(b[0] for b in busy),statistics.median([b[1] for b in busy]),max(b[1] for b in busy)))
PY
