mkdir -p gpurun_out/r2
python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2/plain_list.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(score|trace|rows|null|lrt|collect|gather|walk|alu)' -c 400 --csv --log-file gpurun_out/r2/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > gpurun_out/r2/ncu_list.log 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2/bench_n1_driver.json 2> gpurun_out/r2/bench_n1_driver.err; tail -2 gpurun_out/r2/bench_n1_driver.err
python __graft_entry__.py smoke 2>&1 | tail -2
