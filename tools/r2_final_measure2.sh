# Round 2, after the traceback rewrite: ncu launch list of the driver's command, ncu --set full of the traceback kernel
# on a small config-2-shaped workload, smoke.
mkdir -p gpurun_out/r2
python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2/plain_list.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(score|trace|rows|null|lrt|collect|gather|pack|alu)' -c 400 --csv --log-file gpurun_out/r2/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2/ncu_list.log 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary --profiles 50 --reads 4000"
$B > gpurun_out/r2/tr_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -f -o gpurun_out/r2/k_trace_final $B > gpurun_out/r2/tr_ncu.log 2>&1
python __graft_entry__.py smoke 2>&1 | tail -2
