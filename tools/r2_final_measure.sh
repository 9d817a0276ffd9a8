# Final single-GPU measurements of round 2 (run on the GPU box through gpurun): tests, class sweep, the driver's bench
# command and its reference arm, the ncu launch list, and ncu --set full captures of the top kernels.
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q > gpurun_out/r2/pytest_final.log 2>&1; tail -3 gpurun_out/r2/pytest_final.log
python tools/class_sweep.py --table > gpurun_out/r2/table_sweep_final.jsonl 2> gpurun_out/r2/table_sweep_final.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2/bench_n1_final.json 2> gpurun_out/r2/bench_n1_final.err; tail -2 gpurun_out/r2/bench_n1_final.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2/bench_ref_final.json 2> gpurun_out/r2/bench_ref_final.err
python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2/plain_list.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(score|trace|rows|null|lrt|collect|gather|walk|alu)' -c 400 --csv --log-file gpurun_out/r2/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2/ncu_list.log 2>&1
python bench.py --core 128 --profiles 64 --reads 1000 --steps 1 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2/h8_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_score_h -s 1 -c 1 -o gpurun_out/r2/h8_final python bench.py --core 128 --profiles 64 --reads 1000 --steps 1 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2/h8_ncu.log 2>&1
python bench.py --core 512 --profiles 32 --reads 400 --steps 1 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2/mw512_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_score_mw -s 1 -c 1 -o gpurun_out/r2/mw512_final python bench.py --core 512 --profiles 32 --reads 400 --steps 1 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2/mw512_ncu.log 2>&1
ls gpurun_out/r2 | wc -l
